"""The host control flow of the product's ``solver_socp`` (dots_socp_b200/solver.py) on the CPU.

On a GPU box that loop drives ``engine.Engine`` (CUDA kernels).  Here the engine is replaced by a test double with the
same interface whose arithmetic is the CPU oracle (oracle/alm_oracle.py - test infrastructure, allowed in tests/ only), so
everything ELSE the loop owns is exercised without a GPU and compared with the fixtures of the unmodified reference:
lazy KKT scheduling (which residual on which iteration), penalty path, z rescale, stopping, time limit, step-by-step mode,
tolerance checkpoints, warm start wiring, history rows, solution hand-off (SOCP units and DOT units)."""
import importlib
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import alm_oracle as orc

solver_mod = importlib.import_module("dots_socp_b200.solver")

E2O = dict(phi="phi", A="A", B="B", lam_c="lam_c", mu="mu", E="E", z_fst="z_fst", z_mid="z_mid", z_end="z_end",
           b_fst="b_fst", b_mid="b_mid", b_end="b_end")


class OracleEngine:
    """``engine.Engine``'s public surface, reference layout inside, numpy arithmetic from the oracle."""

    def __init__(self, n_time, geometry, congestion=0.0, eps=0.0, tau=1.9, device=None, leaf_size=16, comm=None, **_):
        self.alm = orc.OracleALM(n_time, geometry, congestion=congestion, eps=eps, tau=tau, is_z_scaling=False)
        o = self.alm.ops
        self.nT, self.V, self.T, self.dt = o.nT, o.V, o.T, o.dt
        self.cong = congestion
        self.device = torch.device("cpu")
        self.timings = {"setup_total": 0.0}
        self.launches = 0
        self.z_valid = True                              # z_mid = 0 is the true initial z_mid (engine.py does the same)
        self.calls = []
        from dots_socp_b200 import dist as dd
        self.comm = dd.Comm()                            # single rank

    r = property(lambda self: self.alm.r)
    ps = property(lambda self: self.alm.ps)
    ds = property(lambda self: self.alm.ds)

    def initial_constant_scaling(self):
        self.alm.initial_constant_scaling()

    def scale_prim_dual(self, factors=None):
        assert self.z_valid, "scale_prim_dual needs z_mid of the previous iteration"
        return self.alm.scale_prim_dual(factors)

    def scale_z(self, f):
        self.alm.scale_z(f)

    def adjust_penalty(self, f):
        self.alm.adjust_penalty(f)

    def step_q0(self):                                   # is_palm: Step 0 from the gradients of the current phi
        assert self.z_valid, "Step 0 needs z_mid of the previous iteration"
        self.alm.dt_phi = orc.grad_time(self.alm.ops.dt, self.alm.phi)
        self.alm.dx_phi = orc.grad_space(self.alm.ops.G, self.alm.phi)
        self.alm.step_q()
        self.calls_q0 = getattr(self, "calls_q0", 0) + 1

    def iterate(self, n=1, write_z=False, kkt1=False):
        for _ in range(n):
            self.alm.iterate()
        self.z_valid = bool(write_z)
        self.kkt1_valid = bool(kkt1) and not write_z      # the triangle kernel accumulated the triangle term of KKT #1 instead
        self.calls.append(bool(write_z))
        self.calls_kkt1 = getattr(self, "calls_kkt1", []) + [self.kkt1_valid]

    def prepare_download(self, shapes=None):             # host-side staging of the CUDA engine's hand-off: nothing to do here
        self.prepared = dict(shapes or {})

    def prefetch_sums(self, conditions):                 # the CUDA engine batches these reductions; nothing to do here
        if 1 in set(conditions):
            assert self.z_valid or getattr(self, "kkt1_valid", False), "KKT #1 prefetched on an iteration that neither stored z_mid nor accumulated its sums"

    def kkt(self, i):
        if i == 1:
            assert self.z_valid or getattr(self, "kkt1_valid", False), "KKT #1 requested on an iteration that neither stored z_mid nor accumulated its sums"
        return self.alm.kkt(i)

    def objective(self):
        return self.alm.objective()

    def from_internal(self, name):
        return torch.from_numpy(np.array(getattr(self.alm, E2O[name]), copy=True))

    def set_state(self, **arrays):
        for name, val in arrays.items():
            setattr(self.alm, E2O[name], np.array(torch.as_tensor(val).numpy(), dtype=np.float64, copy=True))

    def grad_space_into(self, src, dst):
        assert (src, dst) == ("phi", "B")
        self.alm.B = orc.grad_space(self.alm.ops.G, self.alm.phi)

    def E_from_beta(self, scale):
        self.alm.E = -orc.decouple_adjoint(self.alm.b_mid, scale)

    def refresh(self):
        pass

    def solution(self, keys=None):
        sol = self.alm.solution()
        return sol if keys is None else {k: sol[k] for k in keys}

    def dot_solution(self, geometry, centred, root=None):
        av = np.asarray(geometry["area_vertices"])[None, :] / 3.0
        mu = (self.alm.mu * (self.alm.r * self.alm.ds)) * av
        if centred:
            mu = np.concatenate([geometry["mu0"][None], 0.5 * (mu[:-1] + mu[1:]), geometry["mu1"][None]], axis=0)
        return dict(mu=mu, E=(self.alm.E * (self.alm.r * self.alm.ds)) * np.asarray(geometry["area_triangles"])[None, :, None])

    def congestion_norm(self):
        return float(np.linalg.norm(self.alm.lam_c - self.cong * self.alm.r * self.alm.mu))


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self):
        pass

    def elapsed_time(self, other):
        return 0.0


class _TorchShim:
    """Real torch, except the CUDA events / stream the loop uses for its timers."""
    cuda = SimpleNamespace(Event=_Event, current_stream=lambda dev=None: SimpleNamespace(synchronize=lambda: None))

    def __getattr__(self, name):
        return getattr(torch, name)


@pytest.fixture
def cpu_loop(monkeypatch):
    made = []

    def factory(*a, **kw):
        made.append(OracleEngine(*a, **kw))
        return made[-1]

    monkeypatch.setattr(solver_mod, "Engine", factory)
    monkeypatch.setattr(solver_mod, "torch", _TorchShim())
    return made


@pytest.mark.parametrize("name", ["ico2_nt7_c0", "ico2_nt7_c01", "plane8_nt6_c0", "knot_small_nt8_c005", "ico2_nt15_tol1e-4",
                                  "ico2_nt7_stepwise", "ico1_nt1_c005", "ico1_nt2_c0", "ico2_nt7_eps1e-2", "ico2_nt7_tl0",
                                  "ico2_nt7_nit20", "ico2_nt7_palm", "ico2_nt7_cscale", "ico3_nt15_cscale_c0"])
def test_loop_reproduces_reference_runs(cpu_loop, golden, name):
    z, geo, n_time, kw = golden(name)
    sol, hist = solver_mod.solver_socp(n_time, geo, **kw)
    assert int(hist.kkt_iteration[-1]) == int(z["iterations"])
    ref_rows = z["kkt_rows"]
    assert hist.kkt_errors.shape == ref_rows.shape
    assert np.array_equal(np.isnan(hist.kkt_errors), np.isnan(ref_rows))              # same residual on the same iteration
    m = ~np.isnan(ref_rows)
    assert np.allclose(hist.kkt_errors[m], ref_rows[m], rtol=1e-6, atol=1e-12)
    assert hist.history["Transportation cost"][-1] == pytest.approx(float(z["cost"]), rel=1e-8)
    if kw.get("check_kkt_step_by_step"):
        assert np.allclose(hist.history["Transportation cost"], z["cost_history"], rtol=1e-8)
    assert np.abs(sol["mu"] - z["sol_mu"]).max() <= 1e-7 * np.abs(z["sol_mu"]).max()
    eng = cpu_loop[-1]
    # z_mid is only materialised on iterations that check (or may check) KKT #1; most iterations must not ask for it
    if kw.get("is_palm"):
        assert eng.calls_q0 == len(eng.calls) and all(eng.calls)       # Step 0 before every iteration, z_mid always stored
    elif not kw.get("check_kkt_step_by_step") and len(eng.calls) > 50 and name != "ico2_nt7_cscale":
        # (ico2_nt7_cscale hovers just above the tolerance for 1000 iterations, so its validator fires almost every time)
        assert 0 < sum(eng.calls) < 0.8 * len(eng.calls)


@pytest.mark.parametrize("tag,keys", [("full", None), ("part", ("phi", "beta_fst", "beta_end", "beta_mid"))])
def test_loop_warm_start_wiring(cpu_loop, golden, tag, keys):
    z, geo, n_time, kw = golden("ico2_nt7_warm")
    init = {k[5:]: z[k] for k in z.files if k.startswith("init_") and (keys is None or k[5:] in keys)}
    sol, hist = solver_mod.solver_socp(n_time, geo, init_solution=init, **kw)
    assert int(hist.kkt_iteration[-1]) == int(z[tag + "_iterations"])
    assert hist.history["Transportation cost"][-1] == pytest.approx(float(z[tag + "_cost"]), rel=1e-8)
    assert np.abs(sol["mu"] - z[tag + "_mu"]).max() <= 1e-7 * np.abs(z[tag + "_mu"]).max()


def test_loop_checkpoints_and_dot_units(cpu_loop, golden):
    from dots_socp_b200 import surface
    z, geo, n_time, kw = golden("ico2_nt7_c01")
    af = surface.triangle_areas(geo["vertices"], geo["triangles"])
    geo = dict(geo, area_triangles=af, area_vertices=surface.incident_area_sum(geo["vertices"].shape[0], geo["triangles"], af))
    sol, hist = solver_mod.solver(n_time, geo, tol_checkpoints=[1e-1, 1e-2], **kw)
    assert int(hist.kkt_iteration[-1]) == int(z["iterations"])                         # checkpoints do not perturb the run
    av = geo["area_vertices"][None, :] / 3.0
    mid = z["sol_mu"] * av
    assert np.allclose(sol["mu"][1:-1], 0.5 * (mid[:-1] + mid[1:]), rtol=1e-7, atol=1e-14)
    assert np.array_equal(sol["mu"][0], geo["mu0"]) and np.array_equal(sol["mu"][-1], geo["mu1"])
    cps = sol["checkpoints"]
    assert len(cps) == 2 and cps[0]["iteration"] < cps[1]["iteration"] <= int(z["iterations"])
    for cp, level in zip(cps, (1e-1, 1e-2)):
        assert cp["mu"].shape == sol["mu"].shape and cp["E"].shape == sol["E"].shape
        assert max(k for k in cp["kkt"] if k is not None) <= level
    eng = cpu_loop[-1]
    # a DOT-unit run returns mu and E only: its check iterations accumulate KKT #1 on the fly and never store z_mid
    assert not any(eng.calls) and 0 < sum(eng.calls_kkt1) < len(eng.calls)
    raw, _ = solver_mod.solver_raw(n_time, geo, **kw)
    assert np.allclose(raw["mu"], mid, rtol=1e-7, atol=1e-14)


def test_loop_raises_on_non_finite_residual(cpu_loop, golden, monkeypatch):
    z, geo, n_time, kw = golden("ico2_nt7_c0")
    monkeypatch.setattr(OracleEngine, "kkt", lambda self, i: [float("nan"), float("nan")])
    with pytest.raises(FloatingPointError, match="non-finite KKT residual"):
        solver_mod.solver_socp(n_time, geo, **kw)


def _plugin_fixture():
    import os
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "ico2_nt7_plugin.npz"))
    geo = dict(vertices=z["vertices"], triangles=z["triangles"], mu0=z["mu0"], mu1=z["mu1"],
               area_vertices=z["area_vertices"], area_triangles=z["area_triangles"])
    kw = {str(k): float(v) for k, v in zip(z["kw_keys"], z["kw_vals"])}
    kw["nit"] = int(kw["nit"])
    return z, geo, int(z["n_time"]), kw


def check_plugin_outputs(z, sol_c, hist, sol_r, tol):
    """Shared with the GPU test: the plug-in callables against what the reference's own decorators returned."""
    close = lambda a, b: np.abs(a - b).max() <= tol * np.abs(b).max()
    assert int(hist.kkt_iteration[-1]) == int(z["iterations"])
    assert close(sol_c["mu"], z["centre_mu"]) and close(sol_c["E"], z["centre_E"])
    assert close(sol_r["mu"], z["raw_mu"]) and close(sol_r["E"], z["raw_E"])
    assert sol_r.get("checkpoints") is None
    cps = sol_c["checkpoints"]
    assert len(cps) == int(z["n_checkpoints"])
    for i, cp in enumerate(cps):
        assert int(cp["iteration"]) == int(z[f"cp{i}_iteration"])
        assert close(cp["mu"], z[f"cp{i}_mu"]) and close(cp["E"], z[f"cp{i}_E"])
        ref_kkt = z[f"cp{i}_kkt"]
        got = np.array([np.nan if k is None else float(k) for k in cp["kkt"]])
        assert np.array_equal(np.isnan(got), np.isnan(ref_kkt))
        assert np.allclose(got[~np.isnan(got)], ref_kkt[~np.isnan(ref_kkt)], rtol=1e-6, atol=1e-12)


def test_plugin_callables_match_the_reference_decorators(cpu_loop):
    """``solver`` / ``solver_raw`` (the callables handed to run_dot_surface) against the reference's own
    dot_solver_socp_center / dot_solver_socp incl. the tolerance checkpoints (fixture ico2_nt7_plugin)."""
    z, geo, n_time, kw = _plugin_fixture()
    sol_c, hist = solver_mod.solver(n_time, geo, tol_checkpoints=[float(t) for t in z["tol_checkpoints"]], **kw)
    sol_r, _ = solver_mod.solver_raw(n_time, geo, **kw)
    check_plugin_outputs(z, sol_c, hist, sol_r, 1e-7)
