#!/usr/bin/env python
"""bench.py - ALM iterations/s of the DOTs-SOCP hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one inexact semi-proximal ALM iteration (phi solve + cone projection + q/lambda + multiplier
update; reference socp/solver_socp.py:674-722) on a synthetic problem of BASELINE.json's headline size:
subdivided icosphere level 7 (V = 163 842, T = 327 680) x nT = 63 with seeded Gaussian-bump densities.

Own arm (default)
  value    : iterations/s, state resident in HBM, K steps enqueued back to back, CUDA events on the launch
             stream, barrier + synchronize on both sides, max over ranks.  The per-iteration working set
             (>10 GB) is far larger than the 126 MB L2, so no explicit L2 flush is needed.
  e2e      : the same metric through the public plug-in call solver(n_time, geometry) (the callable handed to
             run_dot_surface) with HOST numpy geometry in and the HOST numpy DOT solution (mu, E) out, solved to
             tol=1e-3: iterations / (host->device upload of the inputs + loop time incl. the lazy KKT reductions and
             the device->host read of their scalars + final solution download); the one-off analysis
             (mesh operators, ordering, batched factorisation) is reported beside it and inside
             ``value_incl_setup``, as the reference's own timers keep it apart.
  roofline : the dominant kernel's unique bytes per launch / its mean duration measured live with CUDA events.
  cpu_baseline : the oracle port (numpy/scipy restatement of the reference) on a bounded sample, host cores.

  parity   : measured on the bench workload before the timed region: residual of the phi-solve per time mode
             (space-time operator applied with the stand-alone operator kernels) after 3 iterations, and checksums of the
             iterate compared with the single-GPU record in profiles/.
  secondary: the knots_5-class configurations of BASELINE.json (configs[0..2]): ms per iteration and iterations to
             tol 1e-3, checked against the counts of the unmodified reference (fixtures).

Reference arm (--impl reference): the oracle port (the reference is pure Python and its tree does not travel to the
GPU box) on the host cores, ON THE CONFIGURATION IT PRINTS: all nT+1 modes of the full-size mesh are factorised with
SuperLU (threaded over the modes), then ALM iterations are timed; `steps` is the number of iterations actually timed
(bounded by --ref-seconds), nothing is extrapolated.  NCCL_DEBUG is left as the caller set it: the JSON line is the
last line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "alm_iterations_per_second"
UNIT = "iter/s"
WORKLOADS = {
    # name: (synth example, n_time, congestion, cpu sample example)
    "icosphere7_nt63": ("icosphere7", 63, 0.0, "icosphere5"),
    "icosphere6_nt63": ("icosphere6", 63, 0.0, "icosphere4"),
    "knots5class_nt31_c01": ("knot", 31, 0.1, "knot"),
    "knots5class_nt31": ("knot", 31, 0.0, "knot"),
    "knots5class_nt63": ("knot", 63, 0.0, "knot"),                 # BASELINE configs[2]: time-direction scaling
    "knots5class_nt127": ("knot", 127, 0.0, "knot"),
    "icosphere3_nt31": ("icosphere3", 31, 0.0, "icosphere3"),
}


FULL_SIZE = {"icosphere7": (163842, 327680), "icosphere6": (40962, 81920), "icosphere5": (10242, 20480),
             "icosphere3": (642, 1280), "knot": (4300, 8600)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def sizes(V, T, nT):
    return dict(a=nT * V, c=(nT + 1) * V, b=3 * (nT + 1) * T, z=18 * nT * T)


def reference_bytes_per_iteration(V, T, nT, factor_entries, m):
    """8(27a + 10b + 8c + 7z) + F with F = 2 * factor bytes streamed by the sweeps."""
    s = sizes(V, T, nT)
    return 8 * (27 * s["a"] + 10 * s["b"] + 8 * s["c"] + 7 * s["z"]) + 2 * factor_entries * m * 8


def kernel_bytes(V, T, nT, m_pad, sym):
    """Unique (compulsory) bytes per launch of each kernel group of the fused iteration."""
    a, c = nT * V, (nT + 1) * V
    tri = 8 * ((nT + 1) * T * (18 + 3 + 3) * 2 - 2 * 2 * 9 * T          # b_mid, B, E read + written (two side slots absent)
               + (nT + 1) * T * (6 + 3)                                  # corner_nrm, corner_div written
               + c + a                                                   # phi, lam gathered
               + T * 13) + 4 * 3 * T                                     # mesh constants
    vert = 8 * (c + (nT + 1) * 6 * T - 2 * 3 * T + 4 * a + 8 * a + V) + 4 * (V + 1 + 3 * T)
    rhs = 8 * (3 * a + (nT + 1) * 3 * T + 2 * V + c + V) + 4 * (V + 1 + 3 * T)
    ttr = 8 * (c + V * m_pad) * 2
    # SURVEY 8(d), strictly: the solve pass reads and writes the (nT+1) x V vector once (2c) and streams the factor twice (F);
    # the y / update vectors the two sweeps exchange (3c + 2 * update rows) are counted apart
    sweeps = 8 * m_pad * (2 * sym.panel_entries + 2 * V)
    return {"k_tri_tma": tri, "k_vertex": vert, "k_phi_rhs": rhs, "k_time_mma": ttr, "sweeps": sweeps,
            "sweeps_incl_work_vectors": sweeps + 8 * m_pad * (3 * V + 2 * int(sym.upd_off[-1]))}


class ClockSampler:
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index=0, period=0.02):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.period = period
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self.th.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.nv is not None:
            self.th.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_sample(workload, steps=3, warmup=1, threads=None):
    """cpu_baseline of the own arm: oracle port on the host cores, `steps` ALM iterations on a SMALLER sample mesh of the
    same family (bounded to ~10-30 s of CPU work), scaled linearly in V to the workload's size and labelled as such."""
    from dots_socp_b200 import synth
    from oracle import alm_oracle as orc
    ex, n_time, cong, sample_ex = WORKLOADS[workload]
    threads = threads or (os.cpu_count() or 1)
    geo, _ = synth.example(sample_ex)
    full_v = FULL_SIZE.get(ex, (geo["vertices"].shape[0], None))[0]
    t0 = time.perf_counter()
    ops = orc.MeshOps(n_time, geo, n_threads=threads)
    setup = time.perf_counter() - t0
    alm = orc.OracleALM(n_time, geo, congestion=cong, ops=ops)
    alm.two_threads = True
    for _ in range(warmup):
        alm.iterate()
    t0 = time.perf_counter()
    for _ in range(steps):
        alm.iterate()
    dt = (time.perf_counter() - t0) / steps
    v_s = geo["vertices"].shape[0]
    scale = v_s / full_v
    return dict(value=(1.0 / dt) * scale, unit=UNIT, cores=threads, kind="port, threaded", extrapolated=scale != 1.0,
                sample_value=1.0 / dt, sample_workload=f"{sample_ex}_nt{n_time}",
                sample=(f"oracle port (numpy + SuperLU, per-mode solves on {threads} threads, Laplacian || projection), {steps} ALM "
                        f"iterations on {sample_ex} (V={v_s}) x nT={n_time}: {dt * 1e3:.0f} ms/iteration, scaled by V_sample/V = "
                        f"{scale:.4f} (linear in V: favours the CPU, its sparse LU solves grow faster than V)"),
                ms_per_step_sample=dt * 1e3, setup_s_sample=setup)


def run_reference(args):
    """The CPU implementation of the path on the configuration it prints (no extrapolation): threaded SuperLU factorisation
    of all modes of the full-size mesh (setup, reported), then ALM iterations of the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dots_socp_b200 import synth
    from oracle import alm_oracle as orc
    workload, fallback = args.workload, None
    need_gb = {"icosphere7_nt63": 110.0, "icosphere6_nt63": 28.0}.get(workload, 0.0)   # SuperLU L+U of all modes + oracle state
    try:
        import psutil
        avail_gb = psutil.virtual_memory().available / 2 ** 30
    except Exception:
        avail_gb = float("inf")
    if avail_gb < need_gb and workload == "icosphere7_nt63":
        fallback = (f"{workload} needs ~{need_gb:.0f} GB of host memory for the {WORKLOADS[workload][1] + 1} SuperLU factors, "
                    f"{avail_gb:.0f} GB available: measured on icosphere6_nt63 instead (the own arm reports the same workload "
                    f"under secondary.icosphere6_nt63)")
        workload = "icosphere6_nt63"
    ex, n_time, cong, _ = WORKLOADS[workload]
    threads = os.cpu_count() or 1
    geo, _ = synth.example(ex)
    V, T = geo["vertices"].shape[0], geo["triangles"].shape[0]
    t0 = time.perf_counter()
    ops = orc.MeshOps(n_time, geo, n_threads=threads)
    setup = time.perf_counter() - t0
    alm = orc.OracleALM(n_time, geo, congestion=cong, ops=ops)
    alm.two_threads = True
    t0 = time.perf_counter()
    alm.iterate()                                                  # first iteration: also sizes the time budget
    first = time.perf_counter() - t0
    warm = max(0, min(args.warmup, 2) - 1) if first * 3 < args.ref_seconds else 0
    for _ in range(warm):
        alm.iterate()
    steps = int(max(1, min(args.steps, (args.ref_seconds - first * (1 + warm)) // max(first, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(steps):
        alm.iterate()
    dt = (time.perf_counter() - t0) / steps
    cb = dict(value=1.0 / dt, unit=UNIT, cores=threads, kind="port, threaded", extrapolated=False,
              sample=(f"oracle port (numpy + SuperLU; per-mode solves on {threads} threads, Laplacian || projection as in the "
                      f"reference's is_multi_threads) on the full workload {workload} (V={V}, nT={n_time}): {steps} ALM "
                      f"iterations timed after {1 + warm} warm-up, {dt * 1e3:.0f} ms/iteration; factorisation of the {n_time + 1} "
                      f"modes (setup, not in the value): {setup:.1f} s"),
              setup_s=setup, steps_measured=steps, steps_requested=args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": 1 + warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "same_config": fallback is None,
            "extrapolated": False, "fallback": fallback,
            "config": {"workload": workload, "n_vertices": V, "n_triangles": T, "n_time": n_time, "congestion": cong,
                       "tol": 1e-3, "time_modes": n_time + 1},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


PARITY_RECORD = os.path.join(ROOT, "profiles", "r2_parity_record.json")
SECONDARY = {   # workload -> fixture written by the unmodified reference (tests/golden/make_golden.py); None: no reference run
    "knots5class_nt31": "knots5class_nt31_c0", "knots5class_nt31_c01": "knots5class_nt31_c01",
    "knots5class_nt63": "knots5class_nt63_c0", "knots5class_nt127": "knots5class_nt127_c0",
    "icosphere6_nt63": None}     # the size the reference arm falls back to when the host cannot hold the level-7 factors


def reference_iterations(fixture):
    """Number of ALM iterations of the unmodified reference (its log prints the last 0-based index: +1)."""
    if fixture is None:
        return None
    z = np.load(os.path.join(ROOT, "tests", "golden", fixture + ".npz"), allow_pickle=False)
    return int(z["iterations"]) + 1


def parity_block(eng, workload, world):
    """3 iterations from the start state on the bench workload: phi-solve residual per time mode (single GPU) and checksums
    of the iterate, compared (rtol 1e-9) with the single-GPU record kept under profiles/."""
    import torch
    eng.reset_state()
    eng.scale_z(2.0)
    eng.iterate(3)
    out = {"iterations": 3}
    sums = eng.state_checksums()                      # the iterate after exactly 3 iterations, at every rank count
    out["state_checksums"] = sums
    if world == 1:
        from dots_socp_b200 import capi
        capi.check(eng.lib.dots_step_phi(eng._ctxp, eng.stream), "dots_step_phi")       # the phi-step of iteration 4
        res = eng.phi_residual()
        out["phi_solve_residual_max_over_modes"] = float(res.max())
        out["phi_solve_residual_ok"] = bool(np.isfinite(res).all() and res.max() < 1e-9)
    key = workload
    rec = {}
    if os.path.exists(PARITY_RECORD):
        with open(PARITY_RECORD) as f:
            rec = json.load(f)
    if key in rec:
        worst = 0.0
        for name, (a, b) in rec[key]["state_checksums"].items():
            for x, y in zip((a, b), sums[name]):
                worst = max(worst, abs(x - y) / max(abs(x), 1e-300))
        out["checksums_vs_single_gpu_record"] = worst
        out["checksums_ok"] = bool(worst < 1e-9)
    else:
        out["checksums_vs_single_gpu_record"] = None
    ok = out.get("phi_solve_residual_ok", True) and out.get("checksums_ok", True)
    pinned = ("phi_solve_residual_ok" in out) or ("checksums_ok" in out)
    out["status"] = ("green" if ok else "MISMATCH") if pinned else "unpinned"
    torch.cuda.synchronize()
    return out


def secondary_block(args):
    """BASELINE configs[0..2] (knots_5-class stand-in): iterations to tol 1e-3 through the public solver and ms per
    iteration of the resident loop (CUDA events, `steps` iterations); the working set fits in L2, so an L2 flush (a 256 MB
    fill) runs between the timed batches."""
    import torch
    import dots_socp_b200 as b200
    from dots_socp_b200 import synth
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, fixture in SECONDARY.items():
        want = reference_iterations(fixture)
        ex, n_time, cong, _ = WORKLOADS[name]
        geo, _ = synth.example(ex)
        sol, hist, eng = b200.solver(n_time, geo, congestion=cong, tol=1e-3, nit=3000, return_engine=True, leaf_size=args.leaf)
        iters = int(hist.kkt_iteration[-1]) + 1
        for _ in range(5):
            eng.iterate(1)
        times = []
        for _ in range(5):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                eng.iterate(1)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 50)
        ms = float(np.median(times))
        V, T = geo["vertices"].shape[0], geo["triangles"].shape[0]
        kb = kernel_bytes(V, T, n_time, eng.m_pad, eng.sym)
        fused = sum(v for k, v in kb.items() if k != "sweeps") 
        out[name] = {"ms_per_step": round(ms, 5), "iterations_to_tol": iters, "reference_iterations": want,
                     "iterations_match": (iters == want) if want is not None else None, "time_to_tol_s": round(float(hist.running_time), 4),
                     "fused_bytes": fused, "frac_hbm_peak": round(fused / ms / 1e6 / peaks()[0], 4),
                     "launches_per_iteration": eng.launches_per_iteration()}
        eng.close()
        del eng, sol
    return out


def run_own(args):
    import torch
    import torch.distributed as dist
    from dots_socp_b200 import synth
    from dots_socp_b200 import capi
    import dots_socp_b200 as b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ex, n_time, cong, _ = WORKLOADS[args.workload]
    geo, scale = synth.example(ex)
    V, T = geo["vertices"].shape[0], geo["triangles"].shape[0]

    # warm-up outside every timed region, through the same public call that is measured below: CUDA context, first-use kernel
    # attributes, NCCL channels (collectives + P2P) and the once-per-process host side of the hand-off (192 MB of pinned
    # staging, copy threads, the solution gather of sharded runs)
    warm_geo, _ = synth.example("icosphere2")
    b200.solver(15, warm_geo, tol=1e-3, nit=12, solution_root=0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    # ---- end to end through the public plug-in API (host buffers in, host buffers out) -------------
    t0 = time.perf_counter()
    # the plug-in callable run_dot_surface(opts, solver=...) receives: DOT-unit mu (centred grid) and E come back
    sol, hist, eng = b200.solver(n_time, geo, congestion=cong, tol=1e-3, nit=args.e2e_nit, return_engine=True,
                                 leaf_size=args.leaf, solution_root=0)          # sharded: rank 0 assembles and downloads
    wall = time.perf_counter() - t0
    iters = int(hist.kkt_iteration[-1]) + 1
    setup_s = eng.timings["setup_total"]
    loop_s = wall - setup_s                               # loop + KKT syncs + solution download to host numpy
    upload_s = float(eng.timings.get("upload", 0.0))      # host -> device copy of the mesh constants, index maps, masses + state allocation
    timed_s = upload_s + loop_s                           # e2e timed region: H2D of the inputs + loop + D2H of the result
    h2d = sum(t.numel() * t.element_size() for k, t in eng._keep.items() if k not in ("panels", "panels_t", "phase_clock"))
    d2h_solution = sum(v.nbytes for v in sol.values() if isinstance(v, np.ndarray))      # rank 0 (the only downloader)
    d2h = d2h_solution + 64 * (sum(hist.evaluations) + 8)
    e2e = {"value": iters / timed_s, "unit": UNIT, "h2d_bytes_per_step": h2d / iters, "d2h_bytes_per_step": d2h / iters,
           "iterations_to_tol": iters, "time_to_tol_s": hist.running_time, "timed_s": timed_s, "upload_s": upload_s,
           "loop_plus_download_s": loop_s,
           "setup_s": setup_s, "setup_breakdown_s": {k: round(v, 3) for k, v in eng.timings.items()},
           "value_incl_setup": iters / wall, "transport_cost": float(hist.history["Transportation cost"][-1]) / scale ** 2,
           "kkt_evaluations": hist.evaluations, "converged": bool(np.nanmax(hist.kkt_errors[-1]) < 1e-3)}

    # ---- parity on the bench workload itself, before the timed region -------------------------------
    parity = parity_block(eng, args.workload, world)

    # ---- device-resident timed region ----------------------------------------------------------------
    lib, ctxp = eng.lib, eng._ctxp
    stream = eng.stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        eng.iterate(1)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            eng.iterate(1)                                 # one cudaGraphLaunch per ALM iteration
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step

    # ---- per-kernel-group durations, measured live (second pass, events between the step calls) ------
    sweep_name = "k_ring_run+k_ring_split+k_ring_gather" if eng.sweep_mode == 4 else "k_sweep_run+k_sweep_gather"
    names = ["k_phi_rhs", "comm:all_gather(rhs)", "k_time_fwd", sweep_name, "comm:all_gather(hat)", "k_time_bwd",
             "k_vertex", "comm:halo(vertex)", "k_tri", "comm:halo(corner)"]
    part, comm, tt = eng.part, eng.comm, eng.t
    steps_fn = [
        lambda: capi.check(lib.dots_phi_rhs(ctxp, stream)),
        lambda: (eng.fence() if eng.peers else
                 comm.all_gather_into(tt["rhs"], tt["rhs"][part.rank * part.chunk:(part.rank + 1) * part.chunk])),
        lambda: capi.check(lib.dots_time_transform(ctxp, 0, stream)),
        lambda: capi.check(lib.dots_mode_solves(ctxp, stream)),
        lambda: (eng.fence() if eng.peers else comm.all_gather_into(tt["hat_all"], tt["hat"])),
        lambda: capi.check(lib.dots_time_transform(ctxp, 1, stream)),
        lambda: capi.check(lib.dots_step_vertex(ctxp, stream)),
        lambda: eng.exchange_vertex_halo(pushed=True),
        lambda: capi.check(lib.dots_step_tri(ctxp, 0, stream)),
        lambda: eng.exchange_corner_halo(fence=False),
    ]
    n_probe = min(args.steps, 20)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)] for _ in range(n_probe)]
    barrier()
    for i in range(n_probe):
        for j, fn in enumerate(steps_fn):
            evs[i][j].record()
            if world > 1 or not names[j].startswith("comm:"):
                fn()
        evs[i][-1].record()
    torch.cuda.synchronize()
    raw = {n: float(np.mean([e[j].elapsed_time(e[j + 1]) for e in evs])) for j, n in enumerate(names)}
    if world > 1:
        tmax = torch.tensor([raw[n] for n in names], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        raw = {n: float(v) for n, v in zip(names, tmax.tolist())}
    groups = {"k_phi_rhs": raw["k_phi_rhs"], "k_time_mma": raw["k_time_fwd"] + raw["k_time_bwd"],
              sweep_name: raw[sweep_name], "k_vertex": raw["k_vertex"], "k_tri_tma": raw["k_tri"]}
    comm_ms = {n: round(raw[n], 4) for n in names if n.startswith("comm:")} if world > 1 else {}
    peak, peak_src = peaks()
    kb_all = kernel_bytes(V, T, n_time, n_time + 1, eng.sym)
    sweeps_incl = kb_all.pop("sweeps_incl_work_vectors") / world
    kb_all[sweep_name] = kb_all.pop("sweeps")
    kb = {k: v / world for k, v in kb_all.items()}                                               # per GPU
    kernels = {}
    for name, t_ms in groups.items():
        kernels[name] = {"ms": round(t_ms, 4), "bytes": kb[name], "gbs": round(kb[name] / t_ms / 1e6, 1),
                         "frac": round(kb[name] / t_ms / 1e6 / peak, 4)}
    total_ms = sum(k["ms"] for k in kernels.values()) + sum(comm_ms.values())
    for k in kernels.values():
        k["share"] = round(k["ms"] / total_ms, 4)
    dom = max(kernels, key=lambda n: kernels[n]["ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic if world == 1 else None, "peak_source": peak_src,
                "bytes_per_launch": kb[dom], "ms_per_launch": kernels[dom]["ms"], "kernels": kernels,
                "per_gpu": True, "comm_ms": comm_ms,
                "bytes_definition": ("sweeps: SURVEY 8(d) strictly, 8 M (2 panel entries + 2 V) = factor streamed twice + the "
                                     "solve's vector read and written once; with the y / update work vectors of the two sweeps: "
                                     f"{sweeps_incl:.4g} B per launch")}
    ref_bytes = reference_bytes_per_iteration(V, T, n_time, eng.sym.panel_entries, n_time + 1)
    own_bytes = sum(kb.values()) * world
    roofline["iteration"] = {
        "reference_algorithmic_bytes": ref_bytes, "achieved_vs_reference_bytes_gbs": round(ref_bytes / ms_per_step / 1e6, 1),
        "frac_vs_reference_bytes": round(ref_bytes / ms_per_step / 1e6 / (peak * world), 4),
        "fused_unique_bytes": own_bytes, "achieved_fused_gbs": round(own_bytes / ms_per_step / 1e6, 1),
        "frac_fused": round(own_bytes / ms_per_step / 1e6 / (peak * world), 4), "peak_all_gpus_gbs": peak * world,
        "note": "reference bytes = SURVEY 8(d) formula 8(27a+10b+8c+7z)+F on the reference's data structures; the fused "
                "iteration moves fewer bytes, so its fraction of that figure may exceed 1"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n_vertices": V, "n_triangles": T, "n_time": n_time, "congestion": cong,
                       "tol": 1e-3, "time_modes": n_time + 1, "leaf_size": args.leaf,
                       "sharding": (f"{world} ranks: time slabs of {eng.part.chunk} levels for the streaming kernels, "
                                    f"{eng.part.chunk} time modes per rank for the sweeps" if world > 1 else "single GPU"),
                       "l2": "working set per iteration >> 126 MB L2 (no flush needed)" if V > 20000 else
                             "working set fits in L2 (small config): numbers are launch/latency bound"},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": args.steps * eng.launches_per_iteration(),
            "graph": ("cuda graph (one launch per iteration)" if world == 1 else
                      ("torch CUDA graph incl. NCCL ops" if eng.use_sharded_graphs else "eager launches")),
            "exchange": ("single GPU" if world == 1 else
                         ("peer memory over NVLink (stores for the rhs slabs and halos, loads for the solutions) + one-element all-reduce fences"
                          if eng.peers else f"NCCL all_gather + isend/irecv ({eng.peer_error})")),
            "roofline": roofline, "parity": parity, "sweep_mode": eng.sweep_mode}
    eng.close()
    del eng, sol
    torch.cuda.empty_cache()
    if world == 1 and not args.no_secondary and args.workload == "icosphere7_nt63":
        line["secondary"] = secondary_block(args)
    if rank == 0:
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_sample(args.workload)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        sys.stdout.flush()
        print(json.dumps(line), flush=True)
    return


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="icosphere7_nt63", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-nit", type=int, default=1000, dest="e2e_nit")
    ap.add_argument("--leaf", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    ap.add_argument("--no-secondary", action="store_true", dest="no_secondary")
    ap.add_argument("--ref-seconds", type=float, default=150.0, dest="ref_seconds",
                    help="reference arm: wall-clock budget of the timed ALM iterations")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
