"""ms per ALM iteration (state resident, CUDA events) for any stand-in surface and time grid, including grids beyond 128 levels
(mode groups): python tools/time_iter.py <example> <n_time> [iterations]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                     # noqa: E402

from dots_socp_b200 import synth                 # noqa: E402
from dots_socp_b200.engine import Engine         # noqa: E402

example, n_time = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
geo, _ = synth.example(example)
tm = {}
eng = Engine(n_time, geo, timings=tm)
eng.scale_z(2.0)
eng.iterate(3)
eng.iterate(1)                                   # the first call after the eager warm-up captures the CUDA graph
eng.iterate(2)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.iterate(reps)
b.record()
torch.cuda.synchronize()
print(json.dumps({"example": example, "n_time": n_time, "vertices": eng.V, "mode_groups": eng.n_groups, "m_pad": eng.m_pad,
                  "sweep_mode": eng.sweep_mode, "ms_per_iteration": round(a.elapsed_time(b) / reps, 4),
                  "launches_per_iteration": eng.launches_per_iteration(), "setup_s": {k: round(v, 3) for k, v in tm.items()}}))
