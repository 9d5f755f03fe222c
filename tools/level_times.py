"""In-situ start-to-start times of the per-level sweep launches (DOTS_PHASE_CLOCK=1, %globaltimer stamps).
Usage: python tools/level_times.py [workload] [leaf]"""
import os, sys
os.environ["DOTS_PHASE_CLOCK"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import WORKLOADS
from dots_socp_b200 import synth, nested
from dots_socp_b200.engine import Engine
w = sys.argv[1] if len(sys.argv) > 1 else "icosphere7_nt63"
leaf = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ex, n_time, cong, _ = WORKLOADS[w]
geo, _ = synth.example(ex)
eng = Engine(n_time, geo, congestion=cong, sweep_mode=0, leaf_size=leaf)
eng.scale_z(2.0)
eng.iterate(6)
torch.cuda.synchronize()
t = eng._keep["phase_clock"].cpu().numpy().astype(np.int64)
L = eng.sym.n_levels
d = np.diff(t[:2 * L]) / 1e3
lev = nested.level_schedule(eng.sym)
M = eng.m_pad
for p, us in enumerate(d):
    lv = p if p < L else 2 * L - 1 - p
    nodes = lev[lv]
    ent = int(sum(nested.panel_size(int(eng.sym.s[n]), int(eng.sym.b[n])) for n in nodes))
    key = "fwd" if p < L else "bwd"
    items = np.diff(eng.plan[key + "_ptr"])[lv]
    wpr = int(eng.plan["wpr" if p < L else "cw"][lv])
    wpr = f"{wpr & 15}{'f' if wpr & 16 else ' '}"
    print(f"{key} level {lv:2d} nodes {len(nodes):6d} items {items:6d} wpr {wpr} bytes {ent*M*8/1e6:8.1f} MB  {us:8.1f} us  {ent*M*8/us/1e3:7.0f} GB/s")
print("sum us (all but the last backward level)", d.sum())
