"""Drop-in solver callables for ``dot_surface_socp.interface.run_dot_surface(opts, solver=...)``.

Mirrors the reference's solver plug-in surface:

* ``solver_socp(n_time, geometry, **kw) -> (SolutionSocpData-shaped dict, run history)``
      same keyword names, defaults, error behaviour and output keys/shapes as
      dot_surface_socp/socp/solver_socp.py:25-41,855-871;
* ``solver_raw``  = SOCP units -> DOT units            (socp/solver_decorator.py:10-26, utils/type.py:48-65)
* ``solver``      = ... on the time-centred grid       (socp/solver_decorator.py:29-54)

The outer loop below restates solver_socp.py:565-871 on host scalars; every array operation of an
iteration is a CUDA kernel launched through libdots_b200.so (``engine.Engine``).  The host only syncs
on the iterations whose KKT residuals the reference's lazy validator would evaluate; the iterations
in between are enqueued back to back.
"""
from __future__ import annotations

import logging
import time

import numpy as np
import torch

from .engine import Engine
from .history import RunHistory, LOG_KKT, LOG_SCALING, LOG_INFO
from .schedule import LazyResidualCheck, PenaltySchedule, max_or_none

KKT_LABELS = ["SOC & Org : Primal Feasibility (q)", "SOC       : Primal Feasibility (z)",
              "SOC & Org : Dual Feasibility (alpha)", "SOC       : Dual Feasibility (beta)",
              "      Org : ||rho - Pi+(rho + Fq)||", "      Org : ||m - rho o B||",
              "      Org : ||cong. rho - lambda_c||"]
KKT_SHORT = ["Prim(phi, q)", "Prim(q, z)", "Dual(alpha)", "Dual(beta)", "Comp(rho, f(q))", "Comp(m, rho o B)",
             "Comp(rho, cong.)"]
STEP_TAG = "Step 1-3 (fused Lap / SOC-Proj / Q & Lambda / Multiplier, GPU)"
STOP_SET, PRIM_SET, DUAL_SET = (0, 2, 4, 5), (0, 1), (2, 3)


def _validate_checkpoints(tol_checkpoints, tol):
    """solver_socp.py:85-94."""
    if tol_checkpoints is None:
        return None
    if not isinstance(tol_checkpoints, list) or not tol_checkpoints:
        raise ValueError("tol_checkpoints must be a non-empty list")
    for i, c in enumerate(tol_checkpoints):
        if not (isinstance(c, (int, float)) and 0 < c < 1):
            raise ValueError(f"Invalid checkpoint value at index {i}: {c}. Must be between 0 and 1")
        if c < tol:
            raise ValueError(f"Checkpoint value must be greater than tol. However, checkpoint ({c}) < tol ({tol})")
    return sorted(tol_checkpoints, reverse=True)


def _warm_start(eng: Engine, init):
    """solver_socp.py:239-250: any subset of the solution keys; missing ones get the reference's defaults."""
    nT, V, T = eng.nT, eng.V, eng.T
    dev = eng.device
    get = lambda k, shape: (torch.as_tensor(init[k], dtype=torch.float64, device=dev) if k in init
                            else torch.zeros(shape, dtype=torch.float64, device=dev))
    phi = get("phi", (nT + 1, V))
    bf, be = get("beta_fst", (nT, V)), get("beta_end", (nT, V))
    st = dict(phi=phi, lam_c=get("lambda_c", (nT, V)), z_fst=get("z_fst", (nT, V)), z_end=get("z_end", (nT, V)),
              z_mid=get("z_mid", (nT, 2, 3, T, 3)), b_fst=bf, b_end=be, b_mid=get("beta_mid", (nT, 2, 3, T, 3)),
              A=get("A", None) if "A" in init else torch.diff(phi, dim=0) / eng.dt,
              mu=get("mu", None) if "mu" in init else bf - be)
    eng.set_state(**st)
    if "B" in init:
        eng.set_state(B=init["B"])
    else:
        eng.grad_space_into("phi", "B")
    if "E" in init:
        eng.set_state(E=init["E"])
    else:
        eng.E_from_beta(1.0)
    eng.refresh()
    eng.z_valid = True


def solver_socp(n_time, geometry, congestion=0.0, nit=1000, eps=0.0 * 10 ** (-8), tol=1e-4, tau=1.90,
                is_palm=False, is_multi_threads=True, is_z_scaling=True, is_constant_scaling=False,
                check_kkt_step_by_step=False, init_solution=None, tol_checkpoints=None, time_limit=1000,
                device=None, leaf_size=16, show_progress=False, return_engine=False, comm=None, solution_keys=None,
                dot_units=None, solution_root=None):
    """B200 implementation of ``dot_surface_socp.socp.solver_socp.solver_socp``.

    Extra keyword arguments (``device``, ``leaf_size``, ``show_progress``, ``return_engine``, ``comm``) are additions;
    with torch.distributed initialised (one process per GPU) the problem is sharded over the ranks of ``comm``
    (default: the world group) and every rank returns the full solution; ``solution_keys`` limits which of the twelve
    solution arrays are converted and copied to the host (default: all, as the reference returns them); ``dot_units``
    ("staggered" / "centred", set by ``solver_raw`` / ``solver``) returns the DOT-unit ``mu``, ``E`` formed on the device;
    ``solution_root`` (sharded ``solver`` / ``solver_raw`` runs only): the rank that assembles and downloads the DOT-unit solution,
    the other ranks return ``mu = E = None`` (default None: every rank returns it, as a single process would);
    ``is_multi_threads`` is accepted and ignored (the two reference threads become stream order).
    ``is_palm=True`` and ``is_constant_scaling=True`` are solver-only knobs that the reference's CLI / interface cannot
    reach (interface.py:275-284); both are built: ``is_palm`` as two fused kernels before every iteration (``Engine.step_q0``,
    single GPU only), ``is_constant_scaling`` as the reference's primal / dual rescaling (``Engine.scale_prim_dual``)."""
    logging.basicConfig(level=LOG_INFO, format="%(message)s")
    tol_checkpoints = _validate_checkpoints(tol_checkpoints, tol)
    checkpoints = []

    eng = Engine(n_time, geometry, congestion=congestion, eps=eps, tau=tau, device=device, leaf_size=leaf_size, comm=comm)
    logging.log(LOG_KKT, f"---- Experiment info ".ljust(42, "-") + "\n"
                f"Congestion parameter: {congestion}Number of discretization points in time: {n_time}\n"
                f"Number of discretization vertices: {eng.V}\nNumber of discretization triangles: {eng.T}\nStepsize: {tau}")
    if init_solution:
        _warm_start(eng, init_solution)

    hist = RunHistory(max_record_numbers=max(nit, 1), kkt_labels=KKT_LABELS, kkt_short_labels=KKT_SHORT, name="SOCP",
                      show_progress=show_progress)
    sched = PenaltySchedule()
    lazy = LazyResidualCheck([(lambda i=i: eng.kkt(i)) for i in range(7)], tol)
    hist.setup_time = dict(eng.timings)

    if dot_units is not None and (not eng.comm.enabled or solution_root is None or eng.comm.rank == solution_root):
        # the host side of the hand-off is made ready while the GPU iterates: pinned staging, copy threads, the result arrays
        eng.prepare_download({"mu": (n_time + 1 if dot_units == "centred" else n_time, eng.V), "E": (n_time + 1, eng.T, 3)})
    hist.start()                                                                          # :565
    hist.create_tol_progress(target_tol=tol)
    prim_gap = 1.0 + 1.0 * np.exp(-100 * congestion)                                      # :568
    if is_z_scaling:
        logging.log(LOG_SCALING, "Initially scale z with z factor: 2.0")
        eng.scale_z(2.0)                                                                  # :571-572
    if is_constant_scaling:
        eng.initial_constant_scaling()                                                    # :574-587
    to_scale = lambda k: is_constant_scaling and (k == 10 or k == 50 or k % 100 == 50)    # admm_tools.is_to_scale :98-104
    use_org = False
    z_free_checks = dot_units is not None or (solution_keys is not None and "z_mid" not in solution_keys)
    it, passed = -1, False
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_a.record()
    pending = False                                  # GPU work enqueued since ev_a
    time_up, synced = False, False
    kkt_seconds = 0.0
    start = time.perf_counter()                                                           # :655
    for it in range(nit):                                                                 # :656
        if to_scale(it):                                                                  # :657-659
            if eng.scale_prim_dual():
                logging.log(LOG_SCALING, f"Var Norm at iteration {it}: rescaled, (prim, dual) scale = {eng.ps}, {eng.ds}")
        if is_z_scaling and sched.z_rescale_due(it, hist.get_current_kkt_errors()):       # :661-666
            row = hist.get_current_kkt_errors()
            with np.errstate(all="ignore"):
                f = prim_gap * float(np.sqrt(np.float64(row[1]) / np.float64(row[0])))
            if f > 1.25:
                logging.log(LOG_SCALING, f"Rescale z at iteration {it} with z factor: {f}")
                eng.scale_z(f)
        # The wall-clock limit is sampled before the step is enqueued (the reference samples it after
        # its synchronous steps, :725): the decision whether z_mid must be stored has to precede the launch.
        # On the sharded path the ranks must take the same branch (it feeds the KKT collectives and the break), so the
        # flag is agreed with an all-reduce, and only on iterations that follow a host synchronisation anyway.
        if not eng.comm.enabled:
            time_up = (time.perf_counter() - start) > time_limit
        elif it == 0 or synced:
            time_up = eng.comm.any_true((time.perf_counter() - start) > time_limit, eng.device)
        adjust = sched.due(it) or time_up                                                 # :726
        required = list(PRIM_SET + DUAL_SET) if adjust else None                          # :728-731
        if adjust:
            lazy.restart_ticks()                                                          # :735-736
        will_check = check_kkt_step_by_step or lazy.will_fire(required) or it == nit - 1
        if is_palm:
            eng.step_q0()                                                                 # Step 0, :668-672
        # z_mid itself is stored when something reads it back: Step 0 of is_palm, the variable norms of the re-scaling, the
        # returned SOCP-unit solution.  A check iteration of a DOT-unit run (solver / solver_raw return mu and E only) needs just
        # the triangle term of KKT #1, which the triangle kernel then accumulates on the fly (dots_step_tri mode 2).
        store_z = is_palm or to_scale(it + 1) or (will_check and not z_free_checks)
        eng.iterate(1, write_z=store_z, kkt1=will_check and not store_z)                  # Steps 1-3, :674-722
        pending = True

        cost = lagr = None
        if will_check and pending:
            ev_b.record()
        t_k = time.perf_counter()
        if not check_kkt_step_by_step:
            if required:
                eng.prefetch_sums(required)           # the forced conditions are all evaluated (no early exit): one fused pass
            passed, _ = lazy.evaluate(required)                                           # :738
        else:
            eng.prefetch_sums(range(8))               # all 7 residuals + the objective: one fused multi-reduction pass
            passed, _ = lazy.evaluate(list(range(7)))                                     # :770
            cost, lagr = eng.objective()                                                  # :773-775
        errs = lazy.collect()                                                             # :739-740
        org = [e[0] for e in errs]
        sec = [e[1] for e in errs]
        if any(v is not None and v != v for v in org):
            raise FloatingPointError(f"non-finite KKT residual at iteration {it}: {org} (device state is corrupt)")
        synced = bool(will_check)
        if will_check:
            torch.cuda.current_stream(eng.device).synchronize()
            hist.add_time(STEP_TAG, ev_a.elapsed_time(ev_b) * 1e-3)
            kkt_seconds += time.perf_counter() - t_k
            ev_a.record()
            pending = False
        if adjust and not check_kkt_step_by_step:
            lazy.restart_ticks()                                                          # :742-743
        hist.record(current_it=it, kkt_errors=org,
                    history=None if cost is None else {"Transportation cost": cost, "Objective value": lagr})
        error = max_or_none([org[k] for k in STOP_SET])                                   # :751
        if error is not None and not check_kkt_step_by_step:
            lazy.adapt(error)                                                             # :753-754
        if error is not None and (check_kkt_step_by_step or not adjust):
            hist.show_tol_progress(it, error)                                             # :757-768

        if tol_checkpoints and error is not None and error <= tol_checkpoints[0]:         # :790-801
            checkpoints.append(dict(mu=((eng.r * eng.ds) * eng.from_internal("mu")).cpu().numpy(),
                                    E=((eng.r * eng.ds) * eng.from_internal("E")).cpu().numpy(),
                                    iteration=it, time=hist.get_running_time(), kkt=np.array(org, dtype=object)))
            tol_checkpoints.pop(0)

        if passed or time_up:                                                             # :804
            break
        mx = max_or_none(sec)                                                             # :808-810
        if mx is not None and mx < 5 * tol:
            use_org = True
        if adjust:                                                                        # :813-823
            src = org if use_org else sec
            gap = max_or_none([src[k] for k in PRIM_SET]) / max_or_none([src[k] for k in DUAL_SET])
            eng.adjust_penalty(sched.next_penalty(eng.r, gap) / eng.r)

    eng.prefetch_sums(range(8))
    final = [eng.kkt(i)[0] for i in range(7)]                                             # :826-828
    cost, lagr = eng.objective()                                                          # :829-831
    hist.record(current_it=it, kkt_errors=final, history={"Transportation cost": cost, "Objective value": lagr})
    hist.end()                                                                            # :844
    hist.kkt_seconds = kkt_seconds
    hist.evaluations = list(lazy.evaluations)
    hist.gpu_launches = eng.launches
    if dot_units is None:
        solution = eng.solution(solution_keys)                                            # :845, :855-869
    else:                                                  # decorators' translate / centring done on the device
        solution = eng.dot_solution(geometry, centred=(dot_units == "centred"),
                                    root=solution_root if eng.comm.enabled else None)
    solution["checkpoints"] = checkpoints if checkpoints else None
    logging.log(LOG_INFO, "---- Overview of solution ".ljust(42, "-") + "\n"
                f"Congestion norm: {eng.congestion_norm():.2f}\n"
                f"Number of iterations: {it}\nIteration time: {hist.running_time:.2f}")
    if return_engine:
        return solution, hist, eng
    return solution, hist.as_reference_history()


# ----------------------------------------------------------------------------- decorators
def translate_solution_socp_to_dot(solution_socp, geom):
    """utils/type.py:48-65: densities -> masses (mu * area_v / 3, E * area_f)."""
    av = np.asarray(geom["area_vertices"])[None, :] / 3.0
    af = np.asarray(geom["area_triangles"])[None, :, None]
    out = dict(mu=solution_socp["mu"] * av, E=solution_socp["E"] * af)
    if solution_socp.get("checkpoints"):
        out["checkpoints"] = [dict(mu=c["mu"] * av, E=c["E"] * af, iteration=c["iteration"], time=c["time"], kkt=c["kkt"])
                              for c in solution_socp["checkpoints"]]
    return out


def _centre_in_time(sol, mu0, mu1):
    """socp/solver_decorator.py:32-34."""
    mid = 0.5 * (sol["mu"][:-1] + sol["mu"][1:])
    sol["mu"] = np.concatenate([mu0[None, :], mid, mu1[None, :]], axis=0)


def _dot_checkpoints(sol, geometry, centred):
    """Checkpoints (few, small) keep the host path of utils/type.py:54-63 / solver_decorator.py:50-52."""
    cps = sol.get("checkpoints")
    if not cps:
        sol.pop("checkpoints", None)
        return
    av = np.asarray(geometry["area_vertices"])[None, :] / 3.0
    af = np.asarray(geometry["area_triangles"])[None, :, None]
    out = [dict(mu=c["mu"] * av, E=c["E"] * af, iteration=c["iteration"], time=c["time"], kkt=c["kkt"]) for c in cps]
    if centred:
        for c in out:
            _centre_in_time(c, np.asarray(geometry["mu0"]), np.asarray(geometry["mu1"]))
    sol["checkpoints"] = out


def solver_raw(n_time, geometry, **kwargs):
    """DOT-unit solution on the staggered time grid (reference name ``dot_solver_socp``): solver_socp followed by
    translate_solution_socp_to_dot (socp/solver_decorator.py:10-26), the translation running on the device."""
    res = solver_socp(n_time, geometry, dot_units="staggered", **kwargs)
    _dot_checkpoints(res[0], geometry, centred=False)
    return res


def solver(n_time, geometry, **kwargs):
    """DOT-unit solution on the time-centred grid incl. mu0 / mu1 (reference name ``dot_solver_socp_center``,
    socp/solver_decorator.py:29-54)."""
    res = solver_socp(n_time, geometry, dot_units="centred", **kwargs)
    _dot_checkpoints(res[0], geometry, centred=True)
    return res


solver_raw.__name__ = "dot_solver_socp"
solver.__name__ = "dot_solver_socp_center"
