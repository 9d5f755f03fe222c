"""Small target for `ncu --set full`: one eager ALM iteration (with z_mid stored) and one fused pass over all KKT sums on the
headline workload, nothing else, so that a kernel-name filter captures every hot kernel once.

    python tools/ncu_target.py [workload]
    ncu --set full --clock-control none --import-source on -k "regex:^k_" --launch-skip <setup launches> ... python tools/ncu_target.py

The setup (factorisation) kernels are named k_front_* / k_fl_*: filter with regex:^k_(phi|time|ring|sweep|vertex|tri|kkt|reduce)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                     # noqa: E402

from bench import WORKLOADS                      # noqa: E402
from dots_socp_b200 import capi, synth           # noqa: E402
from dots_socp_b200.engine import Engine         # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "icosphere7_nt63"
ex, n_time, cong, _ = WORKLOADS[workload]
geo, _ = synth.example(ex)
eng = Engine(n_time, geo, congestion=cong)
eng.scale_z(2.0)
eng.use_graphs = False
capi.check(eng.lib.dots_iterate(eng._ctxp, 2, 1, eng.stream))          # two eager iterations, the second stores z_mid
torch.cuda.synchronize()
eng.z_valid = True
eng._state_changed()
eng.prefetch_sums(range(9))                                            # all 7 conditions + objective + variable norms, one pass
print("kkt", [eng.kkt(i)[0] for i in range(7)], "launches", eng.launches)
