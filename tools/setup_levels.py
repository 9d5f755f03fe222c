"""Row f1 on the GPU: where the numeric factorisation spends its time.  The two factorisation entry points of the library
(dots_factor_small_fronts: k_front_small; dots_factor_large_fronts: the k_fl_* kernels of csrc/front_large.cu) are bracketed
with CUDA events per call, i.e. per tree level (and per chunk of large fronts), for the headline mesh; run twice: right after
the host analysis (GPU clocks down) and after a 0.5 s busy loop (clocks up).  Prints one JSON object.

    python tools/setup_levels.py [workload]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                               # noqa: E402
import torch                                     # noqa: E402

from bench import WORKLOADS                      # noqa: E402
from dots_socp_b200 import capi, nested, surface, synth   # noqa: E402
from dots_socp_b200.engine import time_basis     # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "icosphere7_nt63"
ex, n_time, _, _ = WORKLOADS[workload]
geo, _ = synth.example(ex)
lib = capi.load()
dev = torch.device("cuda:0")
torch.zeros(1, device=dev)
v = np.ascontiguousarray(geo["vertices"], dtype=np.float64)
tri = np.ascontiguousarray(geo["triangles"]).astype(np.int64)


class Timed:
    """The library with the two factorisation calls bracketed by events on the launch stream."""

    def __init__(self, lib_):
        self._lib, self.calls = lib_, []

    def __getattr__(self, name):
        return getattr(self._lib, name)

    def _wrap(self, kind, fn, args, n_fronts, n_max):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        self.calls.append((kind, int(n_fronts), int(n_max), e0, e1))
        return rc

    def dots_factor_small_fronts(self, *args):
        return self._wrap("small", self._lib.dots_factor_small_fronts, args, args[1], args[2])

    def dots_factor_large_fronts(self, *args):
        return self._wrap("large", self._lib.dots_factor_large_fronts, args, args[1], args[2])


def run(tag, spin):
    t0 = time.perf_counter()
    mesh = surface.mesh_operators_native(lib, v, tri)
    sym = nested.analyse_native(lib, v, mesh["K"], leaf_size=16)
    host_s = time.perf_counter() - t0
    if spin:
        a = torch.randn(8192, 8192, device=dev)
        t1 = time.perf_counter()
        while time.perf_counter() - t1 < 0.5:
            a = (a @ a).clamp_(-1, 1)
        torch.cuda.synchronize()
        del a
    _, lam = time_basis(n_time)
    timed = Timed(lib)
    t1 = time.perf_counter()
    panels, panels_t = nested.factor_hybrid_device(sym, mesh["K"], mesh["area_sum"] / 3.0, -lam, n_time + 1, dev, timed,
                                                   lambda: torch.cuda.current_stream(dev).cuda_stream)
    enqueue_s = time.perf_counter() - t1
    torch.cuda.synchronize()
    wall_s = time.perf_counter() - t1
    calls = [dict(kind=k, fronts=n, rows_max=m, ms=round(a.elapsed_time(b), 3)) for k, n, m, a, b in timed.calls]
    span = timed.calls[0][3].elapsed_time(timed.calls[-1][4])
    out = dict(host_analysis_s=round(host_s, 3), enqueue_s=round(enqueue_s, 3), wall_s=round(wall_s, 3), gpu_span_ms=round(span, 2),
               small_ms=round(sum(c["ms"] for c in calls if c["kind"] == "small"), 2),
               large_ms=round(sum(c["ms"] for c in calls if c["kind"] == "large"), 2),
               panel_gb=round(2 * panels.numel() * 8 / 1e9, 2), calls=calls)
    del panels, panels_t
    torch.cuda.empty_cache()
    return tag, out


res = dict(workload=workload, n_vertices=int(v.shape[0]), modes=n_time + 1)
for tag, spin in (("after_host_analysis", False), ("after_busy_loop", True), ("after_busy_loop_2", True)):
    k, o = run(tag, spin)
    res[k] = o
print(json.dumps(res))
