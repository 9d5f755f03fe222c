"""A/B the sweep implementations on one GPU: time of dots_mode_solves per call and agreement of the solve with the first variant.

    python tools/sweep_ab.py [workload] [variant ...]        (default: icosphere7_nt63, variants 0 4)

A variant is ``mode[:key=value,...]`` with mode 0 (k_sweep_run) or 4 (ring-streamed, csrc/sweep_ring.cu) and the keys
stages, pdl, split (KB), tasks (per SM), tmin / tmax (KB per contiguous task), wpr (max warps per output), l2hint, e.g.
``4:stages=4,pdl=1,split=64``.  The factorisation is done once per variant (the plan is built in the Engine constructor).
Wrap in `timeout`."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np          # noqa: E402
import torch                # noqa: E402

from bench import WORKLOADS                      # noqa: E402
from dots_socp_b200 import capi, synth           # noqa: E402
from dots_socp_b200.engine import Engine         # noqa: E402

ENV = dict(stages="DOTS_RING_STAGES", pdl="DOTS_RING_PDL", split="DOTS_RING_SPLIT_KB", tasks="DOTS_RING_TASKS_PER_SM",
           tmin="DOTS_RING_TASK_MIN_KB", tmax="DOTS_RING_TASK_MAX_KB", wpr="DOTS_RING_WPR_MAX",
           sb="DOTS_RING_STAGE_BYTES", l2hint="DOTS_RING_L2HINT")

workload = sys.argv[1] if len(sys.argv) > 1 else "icosphere7_nt63"
variants = sys.argv[2:] or ["0", "4"]
ex, n_time, cong, _ = WORKLOADS[workload]
geo, _ = synth.example(ex)
rng = np.random.default_rng(5)
rhs_host = None
ref = None
for spec in variants:
    mode, _, opts = spec.partition(":")
    for name in ENV.values():
        os.environ.pop(name, None)
    for kv in filter(None, opts.split(",")):
        k, v = kv.split("=")
        os.environ[ENV[k]] = v
    eng = Engine(n_time, geo, congestion=cong, sweep_mode=int(mode))
    if rhs_host is None:
        rhs_host = rng.standard_normal((eng.V, eng.m_pad))
    rhs = torch.from_numpy(rhs_host).to(eng.device)

    def solve():
        eng.t["hat"].copy_(rhs)
        capi.check(eng.lib.dots_mode_solves(eng._ctxp, eng.stream))

    solve()
    torch.cuda.synchronize()
    x = eng.t["hat"].cpu().numpy()[:, 1:n_time + 1]                    # mode 0 is singular: pinned, compared through phi elsewhere
    if ref is None:
        ref = x
    err = float(np.abs(x - ref).max() / np.abs(ref).max())
    for _ in range(3):
        solve()
    torch.cuda.synchronize()
    reps = 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    t_copy = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    total = 0.0
    for i in range(reps):
        eng.t["hat"].copy_(rhs)
        ev[i].record()
        capi.check(eng.lib.dots_mode_solves(eng._ctxp, eng.stream))
        t_copy[i].record()
    torch.cuda.synchronize()
    ms = float(np.mean([ev[i].elapsed_time(t_copy[i]) for i in range(reps)]))
    panel_bytes = 2 * 8 * eng.m_pad * eng.sym.panel_entries
    out = {"workload": workload, "variant": spec, "ms_per_solve": round(ms, 4), "panel_gbs": round(panel_bytes / ms / 1e6, 1),
           "rel_diff_vs_first_variant": err, "finite": bool(np.isfinite(x).all()),
           "launches": eng.launches_per_iteration() - 5}
    if eng.ring is not None:
        out["fwd_wpr"] = eng.ring["fwd_wpr"].tolist()
        out["bwd_wpr"] = eng.ring["bwd_wpr"].tolist()
        out["tasks"] = [int(eng.ring["fwd_ptr"][-1]), int(eng.ring["bwd_ptr"][-1])]
    print(json.dumps(out), flush=True)
    eng.close()
    del eng, rhs
    torch.cuda.empty_cache()
