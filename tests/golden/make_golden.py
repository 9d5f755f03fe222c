"""Generate the golden fixtures by running the UNMODIFIED reference (``/root/reference``).

Run in the build container only:  ``python tests/golden/make_golden.py``  (writes tests/golden/*.npz).
The reference is imported through ``refshim`` (numexpr / matplotlib / trimesh stand-ins); nothing of
it is copied.  Live (scaled) solver state is captured non-invasively: ``RunningHistory.record`` is
called exactly once per iteration from ``solver_socp``'s own frame (socp/solver_socp.py:746), so the
caller's locals expose every iterate.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from dots_socp_b200 import synth  # noqa: E402
import refshim  # noqa: E402

STATE = ("phi", "A", "B", "lambda_c", "mu", "E", "z_fst", "z_mid", "z_end", "beta_fst", "beta_mid", "beta_end")
SCALARS = ("r", "scale_factor_z", "constant_d", "prim_scale", "dual_scale", "congestion", "norm_constant_d", "norm_boundary")


def run_reference(geo, n_time, snap_its=(), **kw):
    refshim.load()
    from dot_surface_socp.socp.solver_socp import solver_socp
    from dot_surface_socp.utils import admm_tools

    snaps, r_hist = {}, {}
    orig = admm_tools.RunningHistory.record

    def spy(self, current_it=None, kkt_errors=None, history=None):
        loc = sys._getframe(1).f_locals
        if current_it not in r_hist and "counter_main" in loc:
            r_hist[current_it] = float(loc["r"])
            if current_it in snap_its:
                d = {k: np.array(loc[k], copy=True) for k in STATE}
                d.update({k: float(loc[k]) for k in SCALARS})
                snaps[current_it] = d
        return orig(self, current_it=current_it, kkt_errors=kkt_errors, history=history)

    admm_tools.RunningHistory.record = spy
    cwd = os.getcwd()
    try:
        sol, hist = solver_socp(n_time, geo, **kw)
    finally:
        admm_tools.RunningHistory.record = orig
        os.chdir(cwd)
    return sol, hist, snaps, r_hist


CASES = {
    # name: (example, example kwargs, n_time, solver kwargs, snapshot iterations, keep full final solution)
    "ico2_nt7_c0": ("icosphere2", {}, 7, dict(tol=1e-3, nit=1000), (0, 1, 4, 49), True),
    "ico2_nt7_c01": ("icosphere2", {}, 7, dict(tol=1e-3, nit=1000, congestion=0.1), (0, 1, 4, 49), True),
    "plane8_nt6_c0": ("plane8", {}, 6, dict(tol=1e-3, nit=1000), (0, 9), False),
    "knot_small_nt8_c005": ("knot", dict(n_u=40, n_v=6), 8, dict(tol=1e-3, nit=600, congestion=0.05), (0, 9), False),
    "ico2_nt15_tol1e-4": ("icosphere2", {}, 15, dict(tol=1e-4, nit=3000), (), False),
    "ico3_nt31_c0": ("icosphere3", {}, 31, dict(tol=1e-3, nit=1000), (), False),
    "ico3_nt31_c01": ("icosphere3", {}, 31, dict(tol=1e-3, nit=1000, congestion=0.1), (), False),
    # BASELINE.json configs[0] / configs[1]: the knots_5-class stand-in (V=4300, T=8600), nT=31, tol 1e-3
    # --detail_runhist mode (check_kkt_step_by_step, solver_socp.py:769-787): all 7 residuals + objective every iteration
    "ico2_nt7_stepwise": ("icosphere2", {}, 7, dict(tol=1e-3, nit=400, congestion=0.05, check_kkt_step_by_step=True), (0, 9), False),
    # the reference's OWN example pipeline: load_example('plane', n_space=20) + normalize_geometry (survey: 328 iterations)
    "refplane20_nt15": ("@reference:plane:20", {}, 15, dict(tol=1e-3, nit=1000), (0, 9), False),
    "knots5class_nt31_c0": ("knot", {}, 31, dict(tol=1e-3, nit=1000), (), False),
    "knots5class_nt31_c01": ("knot", {}, 31, dict(tol=1e-3, nit=1000, congestion=0.1), (), False),
    # edge of the time grid: a single time step (two time levels / modes) and two steps, smallest closed mesh (V = 42)
    "ico1_nt1_c005": ("icosphere1", {}, 1, dict(tol=1e-3, nit=500, congestion=0.05), (0, 4), True),
    "ico1_nt2_c0": ("icosphere1", {}, 2, dict(tol=1e-3, nit=500), (0, 4), True),
    # eps > 0: regularised Laplacian (rhs term -eps*area*phi, shift lambda - eps; solver_socp.py:976-986, laplacian_inverse_socp.py:37)
    "ico2_nt7_eps1e-2": ("icosphere2", {}, 7, dict(tol=1e-3, nit=1000, eps=1e-2, congestion=0.05), (0, 4, 49), False),
    # time limit already exceeded at the first check: one iteration, full KKT row, un-converged solution returned (:725-731, :804)
    "ico2_nt7_tl0": ("icosphere2", {}, 7, dict(tol=1e-3, nit=1000, time_limit=0.0), (0,), True),
    # iteration cap reached before convergence: the last iteration is fully checked and returned (:656, :826-871)
    "ico2_nt7_nit20": ("icosphere2", {}, 7, dict(tol=1e-6, nit=20, congestion=0.05), (0, 19), True),
    # is_palm=True (solver-only knob, solver_socp.py:668-672): an extra q / lambda solve opens every iteration
    "ico2_nt7_palm": ("icosphere2", {}, 7, dict(tol=1e-3, nit=1000, congestion=0.05, is_palm=1.0), (0, 1, 4, 49), False),
    # BASELINE.json configs[2]: the same surface at nT = 63 and nT = 127 (time-direction scaling)
    "knots5class_nt63_c0": ("knot", {}, 63, dict(tol=1e-3, nit=2000), (), False),
    "knots5class_nt127_c0": ("knot", {}, 127, dict(tol=1e-3, nit=2000), (), False),
    # >= 10k vertices: fronts beyond the small-front factorisation kernel and split sweep items (round 2; ~10 min of reference time)
    "ico5_nt31_c0": ("icosphere5", {}, 31, dict(tol=1e-3, nit=1000), (), False),
    # is_constant_scaling=True (solver-only knob, solver_socp.py:324-365, 574-587, 657-659): initial primal / dual scaling
    # and re-scaling at iterations 10, 50, 150, ...; snapshots around the scaling iterations
    "ico2_nt7_cscale": ("icosphere2", {}, 7, dict(tol=1e-3, nit=1000, congestion=0.05, is_constant_scaling=1.0), (0, 9, 10, 49, 50, 51), True),
    "ico3_nt15_cscale_c0": ("icosphere3", {}, 15, dict(tol=1e-3, nit=1000, is_constant_scaling=1.0), (0, 10), False),
    # more than 128 time levels: the GPU path solves the modes in groups (160 levels -> 2 groups of 80; 300 -> 3 of 100)
    "ico2_nt159_c0": ("icosphere2", {}, 159, dict(tol=1e-3, nit=1500), (), False),
    "ico2_nt299_c005": ("icosphere2", {}, 299, dict(tol=1e-3, nit=1500, congestion=0.05), (), False),
}


def main(only=None):
    for name, (ex, exkw, n_time, kw, snap_its, keep_full) in CASES.items():
        if only and name not in only:
            continue
        if ex.startswith("@reference:"):
            _, ex_name, n_space = ex.split(":")
            refshim.load()
            cwd = os.getcwd()
            os.chdir(refshim.REFERENCE_ROOT)
            from dot_surface_socp.data.load_example import load_example
            from dot_surface_socp.socp.data_preprocessing import normalize_geometry
            _, raw_geo, _ = load_example(example_name=ex_name, kwargs_generating_mesh={"n": int(n_space)})
            geo, scale = normalize_geometry(raw_geo)
            os.chdir(cwd)
        else:
            geo, scale = synth.example(ex, **exkw)
        sol, hist, snaps, r_hist = run_reference(geo, n_time, snap_its, **kw)
        out = dict(
            vertices=geo["vertices"], triangles=geo["triangles"], mu0=geo["mu0"], mu1=geo["mu1"],
            n_time=n_time, scale_factor=scale,
            kw_keys=np.array(list(kw.keys())), kw_vals=np.array([float(v) for v in kw.values()]),
            iterations=int(hist.kkt_iteration[-1]),
            kkt_rows=hist.kkt_errors, kkt_iteration=hist.kkt_iteration,
            r_history=np.array([r_hist[i] for i in sorted(r_hist)]),
            cost=hist.history["Transportation cost"][-1], objective=hist.history["Objective value"][-1],
            cost_history=hist.history["Transportation cost"], objective_history=hist.history["Objective value"],
            sol_mu=sol["mu"], sol_phi_grad_t=np.diff(sol["phi"], axis=0),
            ref_running_time=hist.running_time, ref_steps_time=float(sum(hist.steps_time.values())),
        )
        if geo["vertices"].shape[0] > 2000:          # keep the big fixtures small: drop the phi gradient, store mu in f32-exact chunks
            out.pop("sol_phi_grad_t")
        if keep_full:
            for k in STATE:
                out["sol_" + k] = sol[k]
        for it, d in snaps.items():
            for k, v in d.items():
                out[f"it{it}_{k}"] = v
        out["snap_its"] = np.array(sorted(snaps), dtype=np.int64)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: V={geo['vertices'].shape[0]} iterations={out['iterations']} cost={out['cost']:.12e} "
              f"-> {os.path.getsize(path) / 1e6:.2f} MB")




def exact_study_fixture():
    """``refplane20_exact.npz``: the reference's analytic transport on ``plane --n_space=20 --ntime=15`` (centred grid,
    data/load_example.py:153-200) and its error functional (utils/evaluate_solution.py:47-69) evaluated on the DOT-unit,
    time-centred solution of the ``refplane20_nt15`` fixture (socp/solver_decorator.py:29-54, utils/type.py:48-65)."""
    refshim.load()
    cwd = os.getcwd()
    os.chdir(refshim.REFERENCE_ROOT)
    try:
        from dot_surface_socp.data.load_example import load_example, load_exact_transportation
        from dot_surface_socp.utils.evaluate_solution import (check_mass_conservation, check_negative_mass,
                                                              compare_with_exact_transportation)
        from dot_surface_socp.socp.data_preprocessing import normalize_geometry
        _, raw_geo, _ = load_example(example_name="plane", kwargs_generating_mesh={"n": 20})
        _, exact = load_exact_transportation(t_array=np.linspace(0.0, 1.0, 16), example_name="plane",
                                             kwargs_generating_mesh={"n": 20})
        norm_geo, _ = normalize_geometry(raw_geo)
    finally:
        os.chdir(cwd)
    fx = np.load(os.path.join(HERE, "refplane20_nt15.npz"))
    mu = fx["sol_mu"] * (norm_geo["area_vertices"][None, :] / 3.0)
    mu = np.concatenate([norm_geo["mu0"][None], 0.5 * (mu[:-1] + mu[1:]), norm_geo["mu1"][None]], axis=0)
    err = compare_with_exact_transportation(mu=mu, mu_exact=exact, geometry=raw_geo, verbose=False)
    neg, neg_layers = check_negative_mass(mu, verbose=False)
    path = os.path.join(HERE, "refplane20_exact.npz")
    np.savez_compressed(path, exact=exact, mu_centred=mu, l1=err["l1"], l2=err["l2"], linf=err["linf"],
                        mass_violation=check_mass_conservation(mu, verbose=False), negative_mass=neg,
                        negative_layers=neg_layers, raw_area_vertices=raw_geo["area_vertices"])
    print(f"refplane20_exact: l1={err['l1']:.6e} l2={err['l2']:.6e} linf={err['linf']:.6e} -> {os.path.getsize(path) / 1e3:.0f} kB")


def warm_start_fixture():
    """``ico2_nt7_warm.npz``: ``init_solution`` (socp/solver_socp.py:239-250).  A coarse reference solve (tol 1e-2) provides
    the start; the reference is then restarted (tol 1e-3) from (a) the full solution dict and (b) a partial dict
    (phi, beta_fst, beta_end, beta_mid only: A, B, mu, E take their defaults)."""
    geo, scale = synth.example("icosphere2")
    kw = dict(congestion=0.05, nit=1000)
    sol0, hist0, _, _ = run_reference(geo, 7, (), tol=1e-2, **kw)
    full = {k: np.array(sol0[k], copy=True) for k in STATE}
    part = {k: np.array(sol0[k], copy=True) for k in ("phi", "beta_fst", "beta_end", "beta_mid")}
    out = dict(vertices=geo["vertices"], triangles=geo["triangles"], mu0=geo["mu0"], mu1=geo["mu1"], n_time=7,
               scale_factor=scale, kw_keys=np.array(list(kw) + ["tol"]), kw_vals=np.array([float(v) for v in kw.values()] + [1e-3]),
               coarse_iterations=int(hist0.kkt_iteration[-1]))
    for k, v in full.items():
        out["init_" + k] = v
    for tag, init in (("full", full), ("part", part)):
        sol, hist, _, _ = run_reference(geo, 7, (), tol=1e-3, init_solution={k: v.copy() for k, v in init.items()}, **kw)
        out[tag + "_iterations"] = int(hist.kkt_iteration[-1])
        out[tag + "_kkt_rows"] = hist.kkt_errors
        out[tag + "_cost"] = hist.history["Transportation cost"][-1]
        out[tag + "_mu"] = sol["mu"]
        out[tag + "_phi_grad_t"] = np.diff(sol["phi"], axis=0)
        out[tag + "_beta_mid"] = sol["beta_mid"]
        print(f"ico2_nt7_warm[{tag}]: iterations={out[tag + '_iterations']} cost={out[tag + '_cost']:.12e}")
    path = os.path.join(HERE, "ico2_nt7_warm.npz")
    np.savez_compressed(path, **out)
    print(f"ico2_nt7_warm: coarse iterations={out['coarse_iterations']} -> {os.path.getsize(path) / 1e6:.2f} MB")


def plugin_fixture():
    """``ico2_nt7_plugin.npz``: the reference's own plug-in callables ``solver`` (dot_solver_socp_center) and ``solver_raw``
    (socp/solver_decorator.py:10-72) on a geometry with areas, with tolerance checkpoints (solver_socp.py:790-801):
    DOT-unit mu / E on the centred and on the staggered grid, and every checkpoint's iteration, kkt row, mu, E."""
    from dots_socp_b200 import surface
    refshim.load()
    from dot_surface_socp.socp.solver_decorator import solver as ref_solver, solver_raw as ref_raw
    geo, scale = synth.example("icosphere2")
    assert "area_vertices" in geo and "area_triangles" in geo
    kw = dict(tol=1e-3, nit=1000, congestion=0.1)
    cps = [1e-1, 1e-2]
    cwd = os.getcwd()
    try:
        sol_c, hist = ref_solver(7, geo, tol_checkpoints=list(cps), **kw)
        sol_r, _ = ref_raw(7, geo, **kw)
    finally:
        os.chdir(cwd)
    out = dict(vertices=geo["vertices"], triangles=geo["triangles"], mu0=geo["mu0"], mu1=geo["mu1"], n_time=7,
               scale_factor=scale, kw_keys=np.array(list(kw)), kw_vals=np.array([float(v) for v in kw.values()]),
               area_vertices=geo["area_vertices"], area_triangles=geo["area_triangles"], tol_checkpoints=np.array(cps),
               iterations=int(hist.kkt_iteration[-1]), centre_mu=sol_c["mu"], centre_E=sol_c["E"],
               raw_mu=sol_r["mu"], raw_E=sol_r["E"], n_checkpoints=len(sol_c["checkpoints"]))
    for i, cp in enumerate(sol_c["checkpoints"]):
        out[f"cp{i}_mu"], out[f"cp{i}_E"] = cp["mu"], cp["E"]
        out[f"cp{i}_iteration"] = int(cp["iteration"])
        out[f"cp{i}_kkt"] = np.array([np.nan if k is None else float(k) for k in cp["kkt"]])
    path = os.path.join(HERE, "ico2_nt7_plugin.npz")
    np.savez_compressed(path, **out)
    print(f"ico2_nt7_plugin: iterations={out['iterations']} checkpoints at {[out[f'cp{i}_iteration'] for i in range(len(cps))]}"
          f" -> {os.path.getsize(path) / 1e6:.2f} MB")


EXTRA = {"refplane20_exact": exact_study_fixture, "ico2_nt7_warm": warm_start_fixture, "ico2_nt7_plugin": plugin_fixture}

if __name__ == "__main__":
    names = sys.argv[1:]
    main([n for n in names if n not in EXTRA] or (None if not names else ["-"]))
    for n, fn in EXTRA.items():
        if not names or n in names:
            fn()
