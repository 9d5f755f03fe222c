"""GPU parity: the CUDA hot path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances (fp64, north_star): per-iterate primal/dual iterates within 1e-8 relative (phi modulo its
additive constant), identical iteration counts, transport cost within 1e-6 relative.  Operator-level
checks use 1e-11 or tighter.  Nothing here reads /root/reference (it does not exist on the GPU box);
the reference's own outputs enter through the committed fixtures in tests/golden."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import alm_oracle as orc          # noqa: E402  (checker only)
from dots_socp_b200 import synth              # noqa: E402
from dots_socp_b200.engine import Engine      # noqa: E402
import dots_socp_b200 as b200                  # noqa: E402

O2E = dict(phi="phi", A="A", B="B", lam_c="lam_c", mu="mu", E="E", z_fst="z_fst", z_mid="z_mid", z_end="z_end",
           b_fst="b_fst", b_mid="b_mid", b_end="b_end")


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def centred(phi, w):
    return phi - (phi * w[None, :]).sum() / (w.sum() * phi.shape[0])


def make_pair(example, n_time, congestion=0.0, leaf=8, sweep_mode=None, **exkw):
    geo, _ = synth.example(example, **exkw)
    alm = orc.OracleALM(n_time, geo, congestion=congestion)
    eng = Engine(n_time, geo, congestion=congestion, leaf_size=leaf, sweep_mode=sweep_mode)
    eng.scale_z(2.0)
    return geo, alm, eng


def push_state(alm, eng):
    st = alm.state()
    eng.set_scalars(r=alm.r, s=alm.s, d=alm.d, norm_d=alm.norm_d)
    eng.set_state(**{O2E[k]: v for k, v in st.items()})
    b = alm.bnd
    eng.t["bnd0"].copy_(torch.from_numpy(b[0][eng.perm_v]))
    eng.t["bnd1"].copy_(torch.from_numpy(b[-1][eng.perm_v]))


def compare_states(alm, eng, tol, label=""):
    got = eng.get_state()
    ref = alm.state()
    worst = {}
    for k, v in ref.items():
        a, b = got[O2E[k]], v
        if k == "phi":
            a, b = centred(a, alm.ops.area_v), centred(b, alm.ops.area_v)
        worst[k] = rel(a, b)
    bad = {k: e for k, e in worst.items() if not e < tol}
    assert not bad, f"{label} mismatch {bad} (all: {worst})"
    return worst


# ---------------------------------------------------------------------------------------------- operators
@pytest.mark.parametrize("example,n_time,leaf", [("icosphere2", 7, 8), ("plane8", 6, 6), ("icosphere3", 31, 16),
                                                 ("icosphere2", 40, 8),
                                                 ("icosphere2", 95, 8), ("icosphere3", 127, 16)])   # 96 / 128 time modes
def test_laplacian_inverse_rows_a7_a8(example, n_time, leaf):
    """rhs assembly, time transform, per-mode solves, inverse transform vs the oracle's SuperLU path."""
    geo, alm, eng = make_pair(example, n_time, congestion=0.05, leaf=leaf)
    rng = np.random.default_rng(3)
    o = alm.ops
    alm.A, alm.lam_c, alm.mu = (rng.standard_normal((n_time, o.V)) for _ in range(3))
    alm.B, alm.E = (rng.standard_normal((n_time + 1, o.T, 3)) for _ in range(2))
    push_state(alm, eng)
    rhs_ref = o.phi_rhs(alm.A, alm.B, alm.lam_c, alm.mu, alm.E, alm.bnd, alm.phi)
    from dots_socp_b200 import capi
    capi.check(eng.lib.dots_phi_rhs(eng._ctxp, eng.stream))
    rhs = eng.from_internal("rhs", eng.t["rhs"]).cpu().numpy()
    assert rel(rhs, rhs_ref) < 1e-12
    capi.check(eng.lib.dots_step_phi(eng._ctxp, eng.stream))
    phi = eng.from_internal("phi").cpu().numpy()
    phi_ref = o.lap_inv(rhs_ref)
    assert rel(centred(phi, o.area_v), centred(phi_ref, o.area_v)) < 1e-9
    # what the iteration consumes are the gradients of phi
    assert rel(orc.grad_time(o.dt, phi), orc.grad_time(o.dt, phi_ref)) < 1e-9
    assert rel(orc.grad_space(o.G, phi), orc.grad_space(o.G, phi_ref)) < 1e-9
    # residual of the space-time operator itself
    lap = orc.div_time(o.dt, o.area_v[None] * orc.grad_time(o.dt, phi)) + orc.div_space(o.D, o.area_f[None, :, None] * orc.grad_space(o.G, phi))
    assert rel(lap, rhs_ref) < 1e-9


def test_grad_div_space_rows_a5_a6():
    geo, alm, eng = make_pair("icosphere2", 5)
    o = alm.ops
    rng = np.random.default_rng(4)
    phi = rng.standard_normal((6, o.V))
    x = rng.standard_normal((6, o.T, 3))
    from dots_socp_b200 import capi
    phi_i, x_i = eng.to_internal("phi", phi), eng.to_internal("B", x)
    g_out = torch.empty((6, 3, o.T), dtype=torch.float64, device=eng.device)
    d_out = torch.empty((6, o.V), dtype=torch.float64, device=eng.device)
    capi.check(eng.lib.dots_grad_space(eng._ctxp, phi_i.data_ptr(), g_out.data_ptr(), eng.stream))
    capi.check(eng.lib.dots_div_space(eng._ctxp, x_i.data_ptr(), d_out.data_ptr(), eng.stream))
    assert rel(eng.from_internal("B", g_out).cpu().numpy(), orc.grad_space(o.G, phi)) < 1e-13
    assert rel(eng.from_internal("phi", d_out).cpu().numpy(), orc.div_space(o.D, x)) < 1e-13


@pytest.mark.parametrize("congestion", [0.0, 0.1])
def test_fused_step_rows_a1_a4(congestion):
    """Projection + q/lambda + multiplier update from a random state with a prescribed phi."""
    n_time = 6
    geo, alm, eng = make_pair("icosphere2", n_time, congestion=congestion)
    o = alm.ops
    rng = np.random.default_rng(5)
    for name in ("A", "lam_c", "mu", "b_fst", "b_end"):
        setattr(alm, name, rng.standard_normal((n_time, o.V)))
    for name in ("B", "E"):
        setattr(alm, name, rng.standard_normal((n_time + 1, o.T, 3)))
    alm.b_mid = rng.standard_normal((n_time, 2, 3, o.T, 3))
    alm.phi = rng.standard_normal((n_time + 1, o.V))
    alm.r = 1.7
    push_state(alm, eng)
    # oracle: Step 1-2, Step 2, Step 3 with this phi (no Laplacian solve)
    s, d, tau = alm.s, alm.d, alm.tau
    alm.z_fst, alm.z_mid, alm.z_end = o.proj_soc(alm.A, alm.B, alm.b_fst, alm.b_mid, alm.b_end, d, s)
    alm.dt_phi, alm.dx_phi = orc.grad_time(o.dt, alm.phi), orc.grad_space(o.G, alm.phi)
    alm.step_q()
    alm.Bd_new = orc.decouple(alm.B, s)
    alm.mu = alm.mu + tau * (alm.dt_phi - alm.A - alm.lam_c)
    alm.E = alm.E + tau * (alm.dx_phi - alm.B)
    alm.b_fst = alm.b_fst + tau * (alm.z_fst + s * alm.A - d)
    alm.b_mid = alm.b_mid + tau * (alm.z_mid - alm.Bd_new)
    alm.b_end = alm.b_end + tau * (alm.z_end - s * alm.A - d)
    from dots_socp_b200 import capi
    capi.check(eng.lib.dots_step_vertex(eng._ctxp, eng.stream))
    capi.check(eng.lib.dots_step_tri(eng._ctxp, 1, eng.stream))
    compare_states(alm, eng, 1e-12, "fused step")


# ---------------------------------------------------------------------------------------------- iterates
@pytest.mark.parametrize("example,n_time,congestion,exkw", [
    ("icosphere2", 7, 0.0, {}), ("icosphere2", 7, 0.1, {}), ("plane8", 6, 0.0, {}),
    ("knot", 8, 0.05, dict(n_u=40, n_v=6)), ("icosphere3", 31, 0.0, {}), ("icosphere2", 127, 0.0, {}),
    ("icosphere2", 159, 0.0, {}), ("icosphere3", 299, 0.05, {}),       # > 128 time levels: 2 / 3 mode groups, plain time transforms
    ("plane20", 15, 0.05, {})])                                        # odd triangle count (909): 8-byte-misaligned planes in the TMA kernel
@pytest.mark.parametrize("sweep_mode", [None, 4])          # None: the engine's choice (small factors: k_sweep_run); 4: ring-streamed
def test_iterates_match_oracle(example, n_time, congestion, exkw, sweep_mode):
    geo, alm, eng = make_pair(example, n_time, congestion=congestion, sweep_mode=sweep_mode, **exkw)
    done = 0
    for k in (1, 2, 5, 50):
        for _ in range(k - done):
            alm.iterate()
        eng.iterate(k - done, write_z=True)
        done = k
        compare_states(alm, eng, 1e-8, f"{example} nT={n_time} k={k}")
        if k == 5:                                   # penalty update + z rescale in the middle of the run
            alm.adjust_penalty(1.35)
            eng.adjust_penalty(1.35)
            alm.scale_z(1.3)
            eng.scale_z(1.3)
            compare_states(alm, eng, 1e-8, "after rescale")


@pytest.mark.parametrize("n_time,groups,m_pad", [(128, 2, 96), (159, 2, 96), (255, 2, 128), (256, 3, 96)])
def test_mode_groups_beyond_128_time_levels_solve_the_space_time_operator(n_time, groups, m_pad):
    """More than 128 time levels on one GPU: the phi-step (rhs, plain forward transform, sweeps per mode group, accumulated
    inverse transform) against the space-time operator formed from the stand-alone gradient / divergence kernels, per mode."""
    geo, _ = synth.example("icosphere3")
    eng = Engine(n_time, geo, congestion=0.05, leaf_size=12)
    assert eng.n_groups == groups and eng.m_pad == m_pad and len(eng.groups) == groups
    eng.scale_z(2.0)
    eng.iterate(3, write_z=True)
    eng.step_phi()
    res = eng.phi_residual()
    assert np.isfinite(res).all() and res.max() < 1e-9, res.max()


def test_kkt_and_objective_rows_a9_a11():
    geo, alm, eng = make_pair("icosphere2", 7, congestion=0.1)
    for _ in range(12):
        alm.iterate()
    eng.iterate(12, write_z=True)
    for i in range(7):
        got, ref = eng.kkt(i), alm.kkt(i)
        assert got[0] == pytest.approx(ref[0], rel=1e-8), (i, got, ref)
        assert (got[1] is None) == (ref[1] is None)
    c_got, c_ref = eng.objective(), alm.objective()
    assert c_got[0] == pytest.approx(c_ref[0], rel=1e-8) and c_got[1] == pytest.approx(c_ref[1], rel=1e-8)


@pytest.mark.parametrize("example,n_time", [("icosphere3", 15), ("knots_5", 31)])
def test_specialised_kkt_passes_equal_the_generic_one(example, n_time):
    """The instantiations of the KKT kernels for the sets the solver asks for (#2 alone, #0-#3) against the generic kernel (same
    sets plus the objective).  Compile-time masks keep the mapping of items to threads: those sums agree bit for bit; Dual(alpha)
    has its own kernel (thread = vertex x 8 time levels): same per-(t, v) terms, another summation order."""
    import ctypes as C
    from dots_socp_b200 import capi
    geo, _ = synth.example(example)
    eng = Engine(n_time, geo, congestion=0.05)
    eng.scale_z(2.0)
    eng.iterate(9, write_z=True)

    def multi(mask):
        out = np.zeros(72)
        capi.check(eng.lib.dots_kkt_sums_multi(eng._ctxp, C.c_uint(mask), out.ctypes.data, eng.stream))
        return out.reshape(9, 8)

    for mask, conds in ((4, [2]), (15, [0, 1, 2, 3])):
        special, generic = multi(mask), multi(mask | 128)
        for w in conds:
            if w == 2:
                assert np.array_equal(special[w][1:], generic[w][1:]) and abs(special[w][0] - generic[w][0]) <= 1e-13 * generic[w][0]
            else:
                assert np.array_equal(special[w], generic[w]), (mask, w)
            assert np.abs(special[w]).max() > 0.0


@pytest.mark.parametrize("example,n_time", [("plane20", 15), ("plane100", 7), ("plane8", 6)])
def test_tma_triangle_kernel_equals_the_plain_one_for_odd_and_even_triangle_counts(monkeypatch, example, n_time):
    """k_tri_tma (bulk copies; with an odd triangle count every other plane starts 8 bytes off a 16-byte boundary and is fetched
    from one element earlier) against the plain-load k_tri (DOTS_TRI_PLAIN=1): same arithmetic, bit-identical state, with and
    without the z_mid store."""
    geo, _ = synth.example(example)
    out = {}
    for plain in ("1", "0"):
        monkeypatch.setenv("DOTS_TRI_PLAIN", plain)
        eng = Engine(n_time, geo, congestion=0.05)
        assert bool(eng.ctx.ring_flags & 4) == (plain == "1")
        eng.scale_z(2.0)
        eng.iterate(4)
        eng.iterate(3, write_z=True)
        out[plain] = eng.get_state()
        eng.close()
    for k, v in out["1"].items():
        assert np.isfinite(v).all(), k
        assert np.array_equal(v, out["0"][k]), k


@pytest.mark.parametrize("example,n_time", [("icosphere4", 31), ("knots_5", 15), ("icosphere2", 7), ("icosphere3", 159), ("plane20", 15)])
def test_kkt1_accumulated_in_the_triangle_kernel_equals_the_stored_z_path(example, n_time):
    """iterate(kkt1=True) (dots_step_tri mode 2: no z_mid store, the triangle term of KKT #1 accumulated per block) against
    iterate(write_z=True) + the KKT pass over the stored z_mid: same iterates bit for bit, same per-element terms, only the
    summation order of the one sum differs."""
    geo, _ = synth.example(example)
    out = {}
    for tag in ("stored", "fused"):
        eng = Engine(n_time, geo, congestion=0.05)
        assert eng.can_fuse_kkt1
        eng.scale_z(2.0)
        eng.iterate(4)                                                   # eager first call, then graph replays
        if tag == "stored":
            eng.iterate(3, write_z=True)
        else:
            eng.iterate(3, kkt1=True)
            assert eng.kkt1_valid and not eng.z_valid
        eng.prefetch_sums([0, 1, 2, 3])                                  # the forced set of a penalty-update iteration
        sums = {i: eng.sums(i).copy() for i in range(4)}
        eng.adjust_penalty(1.3)                                          # keeps the accumulated term valid (z, B, s untouched)
        single = eng.sums(1).copy()
        out[tag] = (sums, single, eng.kkt(1), eng.get_state(("phi", "mu", "B", "b_mid")))
        eng.close()
    for k, v in out["stored"][3].items():
        assert np.array_equal(v, out["fused"][3][k]), k
    for i in (0, 2, 3):
        assert np.array_equal(out["stored"][0][i], out["fused"][0][i]), i
    a, b = out["stored"][0][1], out["fused"][0][1]
    assert np.array_equal(a[:4], b[:4]) and np.array_equal(a[5:], b[5:])  # vertex slots and unused slots: identical
    assert a[4] > 0.0 and abs(a[4] - b[4]) <= 1e-13 * a[4]
    assert np.allclose(out["stored"][1], out["fused"][1], rtol=1e-13, atol=0.0)
    assert out["stored"][2][0] == pytest.approx(out["fused"][2][0], rel=1e-13)


# ---------------------------------------------------------------------------------------------- end to end
@pytest.mark.parametrize("name", ["ico2_nt7_c0", "ico2_nt7_c01", "plane8_nt6_c0", "knot_small_nt8_c005",
                                  "ico2_nt15_tol1e-4", "ico3_nt31_c0", "ico3_nt31_c01",
                                  "knots5class_nt31_c0", "knots5class_nt31_c01",        # BASELINE configs[0], [1]
                                  "knots5class_nt63_c0", "knots5class_nt127_c0",        # BASELINE configs[2]
                                  "ico2_nt7_stepwise", "refplane20_nt15",
                                  "ico1_nt1_c005", "ico1_nt2_c0",                       # smallest time grids
                                  "ico2_nt7_eps1e-2",                                   # regularised Laplacian (eps > 0)
                                  "ico2_nt7_tl0",                                       # time limit hit on the first iteration
                                  "ico5_nt31_c0",                                       # 10 242 vertices: large fronts, split sweep items
                                  "ico2_nt7_cscale", "ico3_nt15_cscale_c0",            # is_constant_scaling=True (primal / dual rescaling)
                                  "ico2_nt159_c0", "ico2_nt299_c005"])                  # > 128 time levels (mode groups)
def test_solver_matches_reference_fixture(golden, name):
    """Through the public solver_socp: iteration count, KKT schedule (which residual on which iteration), penalty
    path, transport cost and the returned mu against the fixtures generated by the unmodified reference."""
    _check_against_fixture(golden, name)


def _check_against_fixture(golden, name):
    z, geo, n_time, kw = golden(name)
    sol, hist, eng = b200.solver_socp(n_time, geo, leaf_size=8 if geo["vertices"].shape[0] < 2000 else 24, return_engine=True, **kw)
    assert int(hist.kkt_iteration[-1]) == int(z["iterations"])
    ref_rows = z["kkt_rows"]
    assert hist.kkt_errors.shape == ref_rows.shape
    assert np.array_equal(np.isnan(hist.kkt_errors), np.isnan(ref_rows))
    m = ~np.isnan(ref_rows)
    assert np.allclose(hist.kkt_errors[m], ref_rows[m], rtol=1e-6, atol=1e-12)
    cost = hist.history["Transportation cost"][-1]
    assert cost == pytest.approx(float(z["cost"]), rel=1e-6)
    if kw.get("check_kkt_step_by_step"):                   # --detail_runhist: the cost is recorded every iteration
        assert np.allclose(hist.history["Transportation cost"], z["cost_history"], rtol=1e-6)
    assert math.sqrt(2 * cost) == pytest.approx(math.sqrt(2 * float(z["cost"])), rel=1e-6)          # W2
    assert rel(sol["mu"], z["sol_mu"]) < 1e-6
    assert sol["z_mid"].shape == (n_time, 2, 3, geo["triangles"].shape[0], 3)
    if "sol_z_mid" in z:
        for key in ("A", "B", "lambda_c", "E", "z_fst", "z_mid", "z_end", "beta_fst", "beta_mid", "beta_end"):
            assert rel(sol[key], z["sol_" + key]) < 1e-6, key


@pytest.mark.parametrize("tag,keys", [("full", None), ("part", ("phi", "beta_fst", "beta_end", "beta_mid"))])
def test_warm_start_matches_reference_fixture(golden, tag, keys):
    """``init_solution`` (solver_socp.py:239-250): restart from a coarse reference solution, all keys or a subset."""
    z, geo, n_time, kw = golden("ico2_nt7_warm")
    init = {k[5:]: z[k] for k in z.files if k.startswith("init_") and (keys is None or k[5:] in keys)}
    sol, hist = b200.solver_socp(n_time, geo, leaf_size=8, init_solution=init, **kw)
    assert int(hist.kkt_iteration[-1]) == int(z[tag + "_iterations"])
    ref_rows = z[tag + "_kkt_rows"]
    assert np.array_equal(np.isnan(hist.kkt_errors), np.isnan(ref_rows))
    m = ~np.isnan(ref_rows)
    assert np.allclose(hist.kkt_errors[m], ref_rows[m], rtol=1e-6, atol=1e-12)
    assert hist.history["Transportation cost"][-1] == pytest.approx(float(z[tag + "_cost"]), rel=1e-6)
    assert rel(sol["mu"], z[tag + "_mu"]) < 1e-6
    assert rel(sol["beta_mid"], z[tag + "_beta_mid"]) < 1e-6
    assert rel(np.diff(sol["phi"], axis=0), z[tag + "_phi_grad_t"]) < 1e-6


def test_drop_in_decorators_and_checkpoints():
    geo, _ = synth.example("icosphere2")
    sol, hist = b200.solver(7, geo, tol=1e-3, nit=400, tol_checkpoints=[1e-1, 1e-2], leaf_size=8)
    V = geo["vertices"].shape[0]
    assert sol["mu"].shape == (8, V)                                  # centred grid incl. mu0, mu1
    assert np.allclose(sol["mu"][0], geo["mu0"]) and np.allclose(sol["mu"][-1], geo["mu1"])
    assert np.allclose(sol["mu"].sum(axis=1), 1.0, atol=5e-3)         # mass conservation per time layer
    assert sol["checkpoints"] and sol["checkpoints"][0]["mu"].shape == (8, V)
    with pytest.raises(ValueError):
        b200.solver_socp(7, geo, tol=1e-3, tol_checkpoints=[1e-4])
    # the device-side translation of solver / solver_raw equals the reference decorators applied on the host
    from dots_socp_b200.solver import translate_solution_socp_to_dot
    full, _ = b200.solver_socp(7, geo, tol=1e-3, nit=400, leaf_size=8)
    raw, _ = b200.solver_raw(7, geo, tol=1e-3, nit=400, leaf_size=8)
    host = translate_solution_socp_to_dot(full, geo)
    assert rel(raw["mu"], host["mu"]) < 1e-13 and rel(raw["E"], host["E"]) < 1e-13
    mid = 0.5 * (host["mu"][:-1] + host["mu"][1:])
    assert rel(sol["mu"][1:-1], mid) < 1e-13 and rel(sol["E"], host["E"]) < 1e-13
    # mass diagnostics formed on the device (row f2) == utils/evaluate_solution.py:7-45 applied to the returned mu
    from dots_socp_b200 import replication as rep
    for s_ in (sol, raw):
        d = s_["diagnostics"]
        assert d["mass_time_layers"].shape == (s_["mu"].shape[0],)
        assert d["mass_conservation"] == pytest.approx(rep.mass_conservation(s_["mu"], verbose=False), rel=1e-9, abs=1e-15)
        neg, neg_layers = rep.negative_mass(s_["mu"], verbose=False)
        assert d["negative_mass"] == pytest.approx(neg, rel=1e-9, abs=1e-18)
        assert np.allclose(d["negative_mass_time_layers"], neg_layers, rtol=1e-9, atol=1e-18)


def test_replication_flow_reproduces_the_reference_plane_run_row_f4(golden):
    """The caller's flow (run_dot_surface_versus_exact, interface.py:386-480) around the plug-in on the reference's own
    ``plane --n_space=20 --ntime=15`` example: iteration count, de-scaled cost, centred DOT-unit mu and the error against
    the analytic transport equal what the unmodified reference produced (fixtures refplane20_nt15 / refplane20_exact)."""
    import os
    from conftest import GOLDEN_DIR
    from dots_socp_b200 import replication as rep
    z, _, n_time, kw = golden("refplane20_nt15")
    ex = np.load(os.path.join(GOLDEN_DIR, "refplane20_exact.npz"))
    opts = rep.options(example="plane", n_space=20, ntime=n_time, tol=kw["tol"], nit=kw["nit"], checkpoints=[1e-1, 1e-2])
    sol, geo, hist, err, cps = rep.run_versus_exact(opts, solver=b200.solver)
    assert int(hist.kkt_iteration[-1]) == int(z["iterations"])
    scale = float(z["scale_factor"])
    assert hist.history["Transportation cost"][-1] == pytest.approx(float(z["cost"]) / scale ** 2, rel=1e-6)
    assert rel(sol["mu"], ex["mu_centred"]) < 1e-6
    for k in ("l1", "l2", "linf"):
        assert err[k] == pytest.approx(float(ex[k]), rel=1e-5)
    assert rep.mass_conservation(sol["mu"], verbose=False) == pytest.approx(float(ex["mass_violation"]), rel=1e-4, abs=1e-9)
    assert [c["kkt_error"] <= t for c, t in zip(cps, (1e-1, 1e-2))] == [True, True]
    assert cps[0]["iteration"] < cps[1]["iteration"] <= int(z["iterations"])
    assert cps[1]["error"]["l2"] < cps[0]["error"]["l2"]


def operator_residual_per_mode(ops, n_time, phi, rhs):
    """max-norm of  div_t(area_v grad_t phi) + D(area_f G phi) - rhs  per time mode, relative to the mode's rhs
    (space-time operator of socp/solver_socp.py:976-986 / utils/laplacian_inverse_socp.py:52-61, oracle's sparse G / D)."""
    from dots_socp_b200.engine import time_basis
    lap = orc.div_time(ops.dt, ops.area_v[None] * orc.grad_time(ops.dt, phi)) \
        + orc.div_space(ops.D, ops.area_f[None, :, None] * orc.grad_space(ops.G, phi))
    Q, _ = time_basis(n_time)
    res_hat, rhs_hat = Q.T @ (lap - rhs), Q.T @ rhs
    return np.abs(res_hat).max(axis=1) / np.maximum(np.abs(rhs_hat).max(axis=1), 1e-300)


def test_large_problem_properties():
    """Full-size style check (size-independent properties): solve residual of the space-time operator and
    feasibility of the cone projection on a 10k-vertex icosphere with nT=63."""
    from dots_socp_b200 import capi
    geo, _ = synth.example("icosphere5")
    n_time = 63
    eng = Engine(n_time, geo, leaf_size=24)
    eng.scale_z(2.0)
    eng.iterate(3, write_z=True)
    st = eng.get_state(("phi", "z_fst", "z_end", "z_mid", "A", "lam_c", "mu", "B", "E"))
    ops = orc.MeshOps(n_time, geo, build_inverse=False)
    # cone feasibility in the area-weighted metric: ||u|| <= z_fst   
    zsq = (ops.diag_soc[None, None, :, :, None] * st["z_mid"]) ** 2
    corner = zsq.sum(axis=(1, 4)).reshape(n_time, 3 * ops.T)
    nrm = np.sqrt(ops.M_oneT.dot(corner.T).T + st["z_end"] ** 2)
    assert (nrm <= st["z_fst"] * (1 + 1e-12) + 1e-12).all()
    assert np.isfinite(st["phi"]).all()
    # the phi-step of the NEXT iteration: residual of the space-time operator against the rhs it was solved for
    capi.check(eng.lib.dots_step_phi(eng._ctxp, eng.stream))
    rhs = eng.from_internal("rhs", eng.t["rhs"][:n_time + 1]).cpu().numpy()
    phi = eng.from_internal("phi").cpu().numpy()
    assert operator_residual_per_mode(ops, n_time, phi, rhs).max() < 1e-9


# ---------------------------------------------------------------------------------------------- headline code path
@pytest.mark.parametrize("example", ["icosphere6", "icosphere7"])
def test_headline_laplacian_solve_residual(example):
    """The bench configuration itself (icosphere level 7 x nT = 63, leaf 16: fronts up to 1 277 rows, split sweep items with
    2 / 4 / 8 warps per output, the large-front factorisation) and the level below it: after 3 real iterations the phi-step
    is checked against the oracle's sparse space-time operator, per time mode.  No CPU factorisation is needed, so
    this runs at the full size in seconds."""
    from dots_socp_b200 import capi
    n_time = 63
    geo, _ = synth.example(example)
    eng = Engine(n_time, geo, leaf_size=16)
    eng.scale_z(2.0)
    eng.iterate(3)
    capi.check(eng.lib.dots_step_phi(eng._ctxp, eng.stream))
    rhs = eng.from_internal("rhs", eng.t["rhs"][:n_time + 1]).cpu().numpy()
    phi = eng.from_internal("phi").cpu().numpy()
    assert np.isfinite(phi).all()
    ops = orc.MeshOps(n_time, geo, build_inverse=False)
    res = operator_residual_per_mode(ops, n_time, phi, rhs)
    assert res.max() < 1e-9, res
    if eng.sweep_mode == 4:                         # the plan really contains the split kernels this test is meant to cover
        assert {int(w) for w in eng.ring["fwd_wpr"]} | {int(w) for w in eng.ring["bwd_wpr"]} >= {1, 2, 4}


@pytest.mark.parametrize("example,iters", [("icosphere5", (1, 2, 5)), ("icosphere6", (1, 2))])
def test_headline_size_iterates_match_oracle(example, iters):
    """Per-iterate parity (1e-8) at nT = 63 on the 10k- and 41k-vertex icospheres: the oracle factorises its 64 modes with
    SuperLU on all host cores (seconds / about a minute)."""
    import os
    n_time = 63
    geo, _ = synth.example(example)
    ops = orc.MeshOps(n_time, geo, n_threads=os.cpu_count() or 1)
    alm = orc.OracleALM(n_time, geo, ops=ops)
    eng = Engine(n_time, geo, leaf_size=16)
    eng.scale_z(2.0)
    done = 0
    for k in iters:
        for _ in range(k - done):
            alm.iterate()
        eng.iterate(k - done, write_z=True)
        done = k
        compare_states(alm, eng, 1e-8, f"{example} nT={n_time} k={k}")


@pytest.mark.parametrize("mode", [0, 4])
@pytest.mark.parametrize("example,n_time,leaf", [("icosphere3", 31, 16), ("icosphere2", 40, 8), ("icosphere3", 95, 16),
                                                 ("icosphere4", 127, 16)])
def test_both_sweep_kernels_match_the_sparse_solve(mode, example, n_time, leaf):
    """dots_mode_solves alone, both implementations (0: k_sweep_run, 4: ring-streamed), 32 / 64 / 96 / 128 modes, against
    scipy's sparse LU of the same shifted matrices (utils/laplacian_inverse_socp.py:34-41,58-59)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from dots_socp_b200 import capi, surface
    geo, _ = synth.example(example)
    eng = Engine(n_time, geo, leaf_size=leaf, sweep_mode=mode)
    rng = np.random.default_rng(17)
    rhs = rng.standard_normal((eng.V, eng.m_pad))
    eng.t["hat"].copy_(torch.from_numpy(rhs))
    capi.check(eng.lib.dots_mode_solves(eng._ctxp, eng.stream))
    x = eng.t["hat"].cpu().numpy()
    K = surface.stiffness_matrix(geo["vertices"], geo["triangles"])
    Kp = K[eng.perm_v][:, eng.perm_v].tocsc()
    M = sp.diags(eng.area_v_new)
    for m in range(1, n_time + 1, max(1, n_time // 6)):            # column 0 = mode 0 is singular (pinned): covered through phi elsewhere
        ref = -spla.spsolve(Kp + (-eng.lam_t[eng.mode_order[m]] + eng.eps) * M, rhs[:, m])        # column m holds mode mode_order[m]
        assert rel(x[:, m], ref) < 1e-10, (m, rel(x[:, m], ref))


@pytest.mark.parametrize("n_time", [15, 31, 63, 95, 127])
def test_symmetric_time_transforms_match_the_full_ones(monkeypatch, n_time):
    """k_time_sym (even / odd split of the DCT-II basis, modes stored even-first) against k_time_mma on the same data:
    forward transform of a random rhs (columns compared through the mode order) and inverse transform back; and against
    numpy's Q^T rhs (laplacian_inverse_socp.py:54,61)."""
    from dots_socp_b200 import capi
    geo, _ = synth.example("icosphere2")
    rng = np.random.default_rng(21)
    out = {}
    for sym in ("1", "0"):
        monkeypatch.setenv("DOTS_TT_SYM", sym)
        eng = Engine(n_time, geo, leaf_size=8)
        assert eng.tt_sym == (sym == "1" and (n_time + 1) % 16 == 0 and eng.m_pad == n_time + 1)
        rhs = rng.standard_normal((n_time + 1, eng.V)) if "rhs" not in out else out["rhs"]
        out["rhs"] = rhs
        eng.t["rhs"][:n_time + 1].copy_(torch.from_numpy(rhs))
        capi.check(eng.lib.dots_time_transform(eng._ctxp, 0, eng.stream))
        hat = eng.t["hat"].cpu().numpy()[:, :n_time + 1]
        nat = np.empty_like(hat)
        nat[:, eng.mode_order] = hat                                   # natural mode order
        capi.check(eng.lib.dots_time_transform(eng._ctxp, 1, eng.stream))
        out[sym] = (nat, eng.slab["phi"].levels(0, n_time + 1).cpu().numpy(), eng.Q)
    ref_hat = (out["0"][2].T @ out["rhs"]).T
    for sym in ("1", "0"):
        assert rel(out[sym][0], ref_hat) < 1e-13
        assert rel(out[sym][1], out["rhs"]) < 1e-12                    # Q Q^T = I
    assert rel(out["1"][0], out["0"][0]) < 1e-13


@pytest.mark.parametrize("stages,pdl", [(2, 0), (4, 1), (3, 1)])
def test_ring_sweep_variants_match(monkeypatch, stages, pdl):
    """Ring depth and programmatic dependent launch change scheduling only: bit-identical iterates."""
    geo, _ = synth.example("icosphere4")
    outs = []
    for st_, pd in ((3, 0), (stages, pdl)):
        monkeypatch.setenv("DOTS_RING_STAGES", str(st_))
        monkeypatch.setenv("DOTS_RING_PDL", str(pd))
        monkeypatch.setenv("DOTS_RING_SPLIT_KB", "24")               # small mesh: force split items too
        eng = Engine(63, geo, leaf_size=16, sweep_mode=4)
        eng.scale_z(2.0)
        eng.iterate(5, write_z=True)
        outs.append(eng.get_state(("phi", "mu", "B")))
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k

@pytest.mark.parametrize("example,n_time", [("icosphere4", 31), ("knots_5", 15), ("icosphere3", 7)])
def test_chained_level_launches_of_the_small_sweep_kernel_match(monkeypatch, example, n_time):
    """k_sweep_run with its level launches chained by programmatic dependent launch (the default) against plain
    stream-ordered launches: scheduling only, bit-identical iterates (graph replay and eager launches)."""
    from dots_socp_b200 import capi
    geo, _ = synth.example(example)
    outs = []
    for pd in (0, 1):
        monkeypatch.setenv("DOTS_RING_PDL", str(pd))
        eng = Engine(n_time, geo, leaf_size=16, sweep_mode=0)
        assert eng.ctx.ring_pdl == pd
        eng.scale_z(2.0)
        eng.iterate(6, write_z=True)                      # graph replays after the warm-up launches
        capi.check(eng.lib.dots_iterate(eng._ctxp, 3, 1, eng.stream))     # eager, several iterations in one chain
        outs.append(eng.get_state(("phi", "mu", "B")))
        eng.close()
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("example,leaf,n_time", [("icosphere3", 8, 7), ("icosphere5", 24, 31), ("plane8", 6, 6), ("icosphere5", 16, 63)])
def test_setup_factorisation_matches_numpy_multifrontal_row_f1(example, leaf, n_time):
    """GPU setup factorisation (hand-written front kernels) against the independent numpy multifrontal checker
    (tests/host_multifrontal.py: numpy Cholesky / inverse per front), both panel layouts; plus a solve through the panels
    against scipy's sparse LU (what the reference factorises with, utils/laplacian_inverse_socp.py:34-41)."""
    import host_multifrontal as hm
    from ring_emulation import transpose_panels
    from dots_socp_b200 import nested, surface, capi
    from dots_socp_b200.engine import time_basis
    geo, _ = synth.example(example)
    v, t = geo["vertices"], geo["triangles"]
    K = surface.stiffness_matrix(v, t)
    mass = surface.incident_area_sum(v.shape[0], t, surface.triangle_areas(v, t)) / 3.0
    sym = nested.analyse(v, K, leaf_size=leaf)
    _, lam = time_basis(n_time)
    shifts = -lam
    m_pad = 8 if n_time + 1 <= 8 else (32 if n_time + 1 <= 32 else 64)
    dev = torch.device("cuda:0")
    stats = {}
    got, got_t = nested.factor_hybrid_device(sym, K, mass, shifts, m_pad, dev, capi.load(),
                                             lambda: torch.cuda.current_stream(dev).cuda_stream, stats=stats)
    torch.cuda.synchronize()
    assert stats["small_fronts"] > 0
    ref = hm.factor_batched(sym, K, mass, shifts, m_pad=m_pad)
    ref_t = transpose_panels(sym, ref)
    got, got_t = got.cpu().numpy(), got_t.cpu().numpy()
    # node by node, relative to the node's largest entry (mode 0 is pinned: its panels are O(1 / pin) in some nodes)
    for name, a, b in (("panels", got, ref), ("panels_t", got_t, ref_t)):
        for i in range(sym.n_nodes):
            p0, p1 = int(sym.panel_off[i]), int(sym.panel_off[i + 1])
            if p1 > p0:
                scale = np.abs(b[p0:p1]).max(axis=0)
                err = (np.abs(a[p0:p1] - b[p0:p1]).max(axis=0) / np.maximum(scale, 1e-300)).max()
                assert err < 1e-9, (name, i, int(sym.s[i]), int(sym.b[i]), err)


def test_small_root_front_is_pinned_deterministically():
    """Meshes whose ROOT separator front fits the small-front kernel (<= 96 rows): the singular time mode must be
    pinned by exactly one thread (regression test for a lost-update race that produced NaN in mode 0)."""
    from dots_socp_b200 import nested, surface, capi
    from dots_socp_b200.engine import time_basis
    geo, _ = synth.example("icosphere3")
    v, t = geo["vertices"], geo["triangles"]
    K = surface.stiffness_matrix(v, t)
    mass = surface.incident_area_sum(v.shape[0], t, surface.triangle_areas(v, t)) / 3.0
    sym = nested.analyse(v, K, leaf_size=8)
    assert int(sym.s[-1] + sym.b[-1]) <= capi.load().dots_front_nmax()
    _, lam = time_basis(31)
    dev = torch.device("cuda:0")
    first = None
    for _ in range(25):
        p, _pt = nested.factor_hybrid_device(sym, K, mass, -lam, 32, dev, capi.load(),
                                             lambda: torch.cuda.current_stream(dev).cuda_stream)
        assert torch.isfinite(p).all()
        first = p if first is None else first
        assert torch.equal(p, first)


def test_is_palm_matches_reference_fixture(golden):
    """is_palm=True (solver_socp.py:668-672): the fused Step-0 kernels (dots_step_q0) against the unmodified reference."""
    _check_against_fixture(golden, "ico2_nt7_palm")


def test_is_palm_step_matches_oracle():
    """One Step 0 from a random state: A, B, lam_c and the refreshed corner terms (through the next fused iteration)."""
    n_time = 6
    geo, alm, eng = make_pair("icosphere2", n_time, congestion=0.07)
    o = alm.ops
    rng = np.random.default_rng(8)
    for name in ("A", "lam_c", "mu", "b_fst", "b_end", "z_fst", "z_end"):
        setattr(alm, name, rng.standard_normal((n_time, o.V)))
    for name in ("B", "E"):
        setattr(alm, name, rng.standard_normal((n_time + 1, o.T, 3)))
    alm.b_mid = rng.standard_normal((n_time, 2, 3, o.T, 3))
    alm.z_mid = rng.standard_normal((n_time, 2, 3, o.T, 3))
    alm.phi = rng.standard_normal((n_time + 1, o.V))
    alm.r = 1.3
    push_state(alm, eng)
    alm.dt_phi, alm.dx_phi = orc.grad_time(o.dt, alm.phi), orc.grad_space(o.G, alm.phi)
    alm.step_q()
    eng.step_q0()
    compare_states(alm, eng, 1e-12, "Step 0")
    alm.iterate()                                   # the corner terms Step 0 refreshed feed the next projection / rhs
    eng.iterate(1, write_z=True)
    compare_states(alm, eng, 1e-8, "iteration after Step 0")


def test_plugin_callables_match_the_reference_decorators_on_gpu():
    """Last in the file on purpose (added after the round's final GPU run): ``solver`` / ``solver_raw`` against the
    reference's own decorators incl. checkpoint iterations and values (fixture ico2_nt7_plugin; same check as the CPU
    loop test, here with the CUDA engine underneath)."""
    from test_solver_loop_cpu import _plugin_fixture, check_plugin_outputs
    z, geo, n_time, kw = _plugin_fixture()
    sol_c, hist = b200.solver(n_time, geo, leaf_size=8, tol_checkpoints=[float(t) for t in z["tol_checkpoints"]], **kw)
    sol_r, _ = b200.solver_raw(n_time, geo, leaf_size=8, **kw)
    check_plugin_outputs(z, sol_c, hist, sol_r, 1e-6)


@pytest.mark.parametrize("name", ["ico2_nt7_nit20"])          # iteration cap reached before convergence
def test_solver_matches_late_reference_fixtures(golden, name):
    """Fixtures added after the round's final GPU run (kept last so that they cannot mask the validated tests above)."""
    _check_against_fixture(golden, name)
