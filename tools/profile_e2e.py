"""cProfile of one end-to-end solve through the public plug-in call (headline workload): where the host time goes."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                     # noqa: E402

import dots_socp_b200 as b200                    # noqa: E402
from bench import WORKLOADS                      # noqa: E402
from dots_socp_b200 import synth                 # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "icosphere7_nt63"
ex, n_time, cong, _ = WORKLOADS[workload]
geo, _ = synth.example(ex)
warm, _ = synth.example("icosphere2")
b200.solver_socp(15, warm, tol=1e-3, nit=12)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
sol, hist = b200.solver(n_time, geo, congestion=cong, tol=1e-3, nit=1000)
pr.disable()
st = pstats.Stats(pr, stream=sys.stdout)
st.sort_stats("cumulative").print_stats(45)
