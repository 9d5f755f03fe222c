"""Cost of the pieces a solve adds to the plain iteration (headline workload, one GPU): iteration with / without z_mid store,
the fused KKT pass for the masks the solver uses, the penalty update.  CUDA events, 10 repetitions each."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np          # noqa: E402
import torch                # noqa: E402

from bench import WORKLOADS                      # noqa: E402
from dots_socp_b200 import synth                 # noqa: E402
from dots_socp_b200.engine import Engine         # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "icosphere7_nt63"
ex, n_time, cong, _ = WORKLOADS[workload]
geo, _ = synth.example(ex)
eng = Engine(n_time, geo, congestion=cong)
eng.scale_z(2.0)
eng.iterate(3, write_z=True)
eng.iterate(3, write_z=False)
torch.cuda.synchronize()


def timed(fn, reps=10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    for i in range(reps):
        ev[i].record()
        fn()
    ev[reps].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[reps]) / reps


def sums(conds):
    def go():
        eng._state_changed()
        eng.prefetch_sums(conds)
    return go


out = {"workload": workload, "setup": {k: round(v, 3) for k, v in eng.timings.items()}}
out["iterate"] = timed(lambda: eng.iterate(1, write_z=False))
out["iterate_write_z"] = timed(lambda: eng.iterate(1, write_z=True))
eng.iterate(1, write_z=True)
for name, conds in (("kkt_cond2", [2]), ("kkt_cond0", [0]), ("kkt_cond1", [1]), ("kkt_cond3", [3]), ("kkt_prim_dual_0123", [0, 1, 2, 3]),
                    ("kkt_4", [4]), ("kkt_5", [5]), ("kkt_6", [6]), ("kkt_all8", list(range(8)))):
    out[name] = timed(sums(conds))
out["adjust_penalty_pair"] = timed(lambda: (eng.adjust_penalty(1.25), eng.adjust_penalty(0.8))) / 2
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()}))
