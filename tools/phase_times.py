"""Per-phase times of the persistent sweep kernel (DOTS_PHASE_CLOCK=1).  Usage: python tools/phase_times.py [workload]"""
import os, sys
os.environ["DOTS_PHASE_CLOCK"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import WORKLOADS
from dots_socp_b200 import synth, nested
from dots_socp_b200.engine import Engine
w = sys.argv[1] if len(sys.argv) > 1 else "icosphere7_nt63"
ex, n_time, cong, _ = WORKLOADS[w]
geo, _ = synth.example(ex)
eng = Engine(n_time, geo, congestion=cong, sweep_mode=1)
eng.use_graphs = False
eng.scale_z(2.0)
eng.iterate(5)
torch.cuda.synchronize()
t = eng._keep["phase_clock"].cpu().numpy().astype(np.int64)
d = np.diff(t) / 1e3
L = eng.sym.n_levels
lev = nested.level_schedule(eng.sym)
M = eng.m_pad
for p, us in enumerate(d):
    lv = p if p < L else 2 * L - 1 - p
    nodes = lev[lv]
    ent = int(sum(nested.panel_size(int(eng.sym.s[n]), int(eng.sym.b[n])) for n in nodes))
    items = np.diff(eng.plan["fwd_ptr" if p < L else "bwd_ptr"])[lv]
    print(f"phase {p:2d} {'fwd' if p < L else 'bwd'} level {lv:2d} nodes {len(nodes):6d} items {items:6d} bytes {ent*M*8/1e6:8.1f} MB  {us:8.1f} us  {ent*M*8/us/1e3:7.0f} GB/s")
print("total us", d.sum())
