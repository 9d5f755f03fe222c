"""Where does the variance of the setup time come from?  Engine setup repeated in one process, with and without a busy GPU
before it, plus the raw cost of a large fresh device allocation."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                     # noqa: E402

from dots_socp_b200 import synth                 # noqa: E402
from dots_socp_b200.engine import Engine         # noqa: E402

geo, _ = synth.example("icosphere7")
torch.cuda.init()
torch.zeros(1, device="cuda")
torch.cuda.synchronize()
out = {}


def alloc(gb):
    t0 = time.perf_counter()
    x = torch.empty(int(gb * 2 ** 30) // 8, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    x.zero_()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    del x
    torch.cuda.empty_cache()
    return round(t1 - t0, 4), round(t2 - t1, 4)


out["alloc_10GB_cold"] = alloc(10)
out["alloc_10GB_again"] = alloc(10)
out["alloc_40GB"] = alloc(40)


def setup(tag, spin=False):
    if spin:
        a = torch.randn(8192, 8192, device="cuda")
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.5:
            a = (a @ a).clamp_(-1, 1)
        torch.cuda.synchronize()
        del a
    tm = {}
    eng = Engine(63, geo, timings=tm)
    out[tag] = {k: round(v, 3) for k, v in tm.items()}
    eng.close()
    del eng
    torch.cuda.empty_cache()


setup("setup_1")
setup("setup_2")
time.sleep(3.0)
setup("setup_3_after_idle")
time.sleep(3.0)
setup("setup_4_after_idle_with_spin", spin=True)
print(json.dumps(out))
