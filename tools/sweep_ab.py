"""A/B the sweep implementations on one GPU: time of dots_mode_solves per call and agreement of the iterates with mode 0.

    python tools/sweep_ab.py [workload] [mode ...]        (default: icosphere7_nt63, modes 0 2 3)

Modes: 0 per-level k_sweep_run (default path), 1 persistent cooperative TMA kernel, 2 tile-streamed, 3 tile-streamed with
programmatic dependent launch (include/dots_b200.h: sweep_mode).  Wrap in `timeout`: modes 2 / 3 are experimental."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np          # noqa: E402
import torch                # noqa: E402

from bench import WORKLOADS                      # noqa: E402
from dots_socp_b200 import capi, synth           # noqa: E402
from dots_socp_b200.engine import Engine         # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "icosphere7_nt63"
modes = [int(m) for m in sys.argv[2:]] or [0, 2, 3]
ex, n_time, cong, _ = WORKLOADS[workload]
geo, _ = synth.example(ex)
ref_phi = None
for mode in modes:
    eng = Engine(n_time, geo, congestion=cong, sweep_mode=mode)
    eng.scale_z(2.0)
    eng.iterate(4, write_z=True)
    torch.cuda.synchronize()
    phi = eng.from_internal("phi").cpu().numpy()
    phi -= phi.mean()
    if ref_phi is None:
        ref_phi = phi
    err = float(np.abs(phi - ref_phi).max() / np.abs(ref_phi).max())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        capi.check(eng.lib.dots_mode_solves(eng._ctxp, eng.stream))
    torch.cuda.synchronize()
    reps = 20
    ev[0].record()
    for _ in range(reps):
        capi.check(eng.lib.dots_mode_solves(eng._ctxp, eng.stream))
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    panel_bytes = 2 * 8 * eng.m_pad * eng.sym.panel_entries
    print(json.dumps({"workload": workload, "sweep_mode": mode, "ms_per_solve": round(ms, 4),
                      "panel_gbs": round(panel_bytes / ms / 1e6, 1), "rel_diff_phi_vs_first_mode": err,
                      "finite": bool(np.isfinite(phi).all())}), flush=True)
    eng.close()
    del eng
    torch.cuda.empty_cache()
