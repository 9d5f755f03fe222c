"""Small driver for ncu: build the engine for a workload, run a few ALM iterations.  Usage:
    python tools/profile_iter.py [workload] [iterations]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS
from dots_socp_b200 import synth
from dots_socp_b200.engine import Engine

workload = sys.argv[1] if len(sys.argv) > 1 else "icosphere7_nt63"
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ex, n_time, cong, _ = WORKLOADS[workload]
geo, _ = synth.example(ex)
eng = Engine(n_time, geo, congestion=cong)
eng.scale_z(2.0)
torch.cuda.synchronize()
print("launches per iteration", eng.launches_per_iteration(), "levels", eng.sym.n_levels, flush=True)
eng.iterate(n_it, write_z=False)
torch.cuda.synchronize()
print("done", float(eng.slab["phi"].data.abs().max()))
