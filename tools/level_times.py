"""Per-launch times of one pair of ring sweeps (sweep_mode 4) measured with CUDA events between the launches
(dots_ring_level_times), next to the panel bytes each launch streams.

    python tools/level_times.py [workload] [key=value ...]      keys as in tools/sweep_ab.py (stages, pdl, split, tasks, tmin, tmax, wpr)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ENV = dict(stages="DOTS_RING_STAGES", pdl="DOTS_RING_PDL", split="DOTS_RING_SPLIT_KB", tasks="DOTS_RING_TASKS_PER_SM",
           tmin="DOTS_RING_TASK_MIN_KB", tmax="DOTS_RING_TASK_MAX_KB", wpr="DOTS_RING_WPR_MAX",
           sb="DOTS_RING_STAGE_BYTES")
args = [a for a in sys.argv[1:] if "=" not in a]
for kv in (a for a in sys.argv[1:] if "=" in a):
    k, v = kv.split("=")
    os.environ[ENV[k]] = v
import numpy as np          # noqa: E402
import torch                # noqa: E402

from bench import WORKLOADS                      # noqa: E402
from dots_socp_b200 import capi, nested, synth   # noqa: E402
from dots_socp_b200.engine import Engine         # noqa: E402

w = args[0] if args else "icosphere7_nt63"
ex, n_time, cong, _ = WORKLOADS[w]
geo, _ = synth.example(ex)
eng = Engine(n_time, geo, congestion=cong, sweep_mode=4)
eng.scale_z(2.0)
eng.iterate(4)
torch.cuda.synchronize()
cap = 200
ms, tag, n = (C.c_float * cap)(), (C.c_int32 * cap)(), C.c_int(0)
acc = None
reps = 5
for _ in range(reps):
    capi.check(eng.lib.dots_ring_level_times(eng._ctxp, eng.stream, ms, tag, cap, C.byref(n)), "dots_ring_level_times")
    cur = np.array(ms[:n.value], dtype=np.float64)
    acc = cur if acc is None else acc + cur
acc /= reps
lev = nested.level_schedule(eng.sym)
M, sym, rp = eng.m_pad, eng.sym, eng.ring
total = 0.0
for i in range(n.value):
    t, us = tag[i], acc[i] * 1e3
    total += us
    if 1000 <= t < 2000:
        lv = t - 1000
        nv = int(rp["gv_ptr"][lv + 1] - rp["gv_ptr"][lv])
        print(f"gather lvl {lv:2d} vertices {nv:7d}                                  {us:8.1f} us")
        continue
    d, lv = ("bwd", t - 2000) if t >= 2000 else ("fwd", t)
    nodes = lev[lv]
    ent = int(sum(nested.panel_size(int(sym.s[k]), int(sym.b[k])) for k in nodes))
    items = int(rp[d + "_ptr"][lv + 1] - rp[d + "_ptr"][lv])
    print(f"{d} level {lv:2d} nodes {len(nodes):6d} items {items:6d} wpr {int(rp[d + '_wpr'][lv])} bytes {ent * M * 8 / 1e6:8.1f} MB "
          f"{us:8.1f} us {ent * M * 8 / us / 1e3:7.0f} GB/s")
print(f"sum {total:.1f} us; panel bytes {2 * 8 * M * sym.panel_entries / 1e9:.2f} GB -> {2 * 8 * M * sym.panel_entries / total / 1e3:.0f} GB/s")
