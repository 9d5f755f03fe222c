"""Pin the CPU oracle (oracle/alm_oracle.py) to outputs of the unmodified reference.

Fixtures: tests/golden/*.npz, written by tests/golden/make_golden.py from /root/reference."""
import numpy as np
import pytest

from oracle import alm_oracle as orc

ORACLE_NAMES = dict(phi="phi", A="A", B="B", lambda_c="lam_c", mu="mu", E="E", z_fst="z_fst", z_mid="z_mid",
                    z_end="z_end", beta_fst="b_fst", beta_mid="b_mid", beta_end="b_end")


def rel_err(a, b):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


def phi_mod_const(phi):
    return phi - phi.mean()


@pytest.mark.parametrize("name", ["ico2_nt7_c0", "ico2_nt7_c01", "plane8_nt6_c0", "knot_small_nt8_c005",
                                  "ico2_nt7_stepwise", "refplane20_nt15", "ico1_nt1_c005", "ico1_nt2_c0",
                                  "ico2_nt7_eps1e-2", "ico2_nt7_tl0", "ico2_nt7_nit20", "ico2_nt7_palm",
                                  "ico2_nt7_cscale", "ico3_nt15_cscale_c0"])        # is_constant_scaling=True
def test_iterates_match_reference(golden, name):
    z, geo, n_time, kw = golden(name)
    snap_its = [int(i) for i in z["snap_its"]]
    got = {}

    def trace(it, alm):
        if it in snap_its:
            got[it] = (alm.state(), alm.r, alm.s, alm.d, alm.ps, alm.ds, alm.cong)

    sol, info = orc.solve(n_time, geo, trace=trace, **kw)
    assert info["iterations"] == int(z["iterations"])
    for it in snap_its:
        st, r, s, d, ps, ds, cong = got[it]
        assert r == pytest.approx(float(z[f"it{it}_r"]), rel=1e-12)
        assert s == float(z[f"it{it}_scale_factor_z"]) and d == pytest.approx(float(z[f"it{it}_constant_d"]), rel=1e-13)
        if f"it{it}_prim_scale" in z:                      # fixtures written since the constant-scaling knob is covered
            assert ps == pytest.approx(float(z[f"it{it}_prim_scale"]), rel=1e-12)
            assert ds == pytest.approx(float(z[f"it{it}_dual_scale"]), rel=1e-12)
            assert cong == pytest.approx(float(z[f"it{it}_congestion"]), rel=1e-12, abs=0.0)
        for ref_name, my_name in ORACLE_NAMES.items():
            a, b = st[my_name], z[f"it{it}_{ref_name}"]
            if ref_name == "phi":
                a, b = phi_mod_const(a), phi_mod_const(b)
            assert rel_err(a, b) < 1e-9, (it, ref_name, rel_err(a, b))
    assert info["cost"] == pytest.approx(float(z["cost"]), rel=1e-9)
    assert info["objective"] == pytest.approx(float(z["objective"]), rel=1e-9)
    assert rel_err(sol["mu"], z["sol_mu"]) < 1e-8
    if "sol_z_mid" in z:
        for ref_name in ORACLE_NAMES:
            a, b = sol[ref_name], z["sol_" + ref_name]
            if ref_name == "phi":
                a, b = phi_mod_const(a), phi_mod_const(b)
            assert rel_err(a, b) < 1e-8, ref_name


@pytest.mark.parametrize("name", ["ico2_nt15_tol1e-4", "ico3_nt31_c0", "ico3_nt31_c01",
                                  "ico2_nt159_c0", "ico2_nt299_c005"])                   # more than 128 time levels
def test_schedule_and_history_match_reference(golden, name):
    """Iteration count, which KKT conditions were evaluated when (nan pattern), their values, penalty path, cost."""
    z, geo, n_time, kw = golden(name)
    sol, info = orc.solve(n_time, geo, **kw)
    assert info["iterations"] == int(z["iterations"])
    ref_rows = z["kkt_rows"]
    my_rows = info["kkt_rows"].copy()
    my_rows[-1] = info["final_kkt"]                       # the reference overwrites the last row with the full final check
    assert my_rows.shape == ref_rows.shape
    assert np.array_equal(np.isnan(my_rows), np.isnan(ref_rows))
    m = ~np.isnan(ref_rows)
    assert np.allclose(my_rows[m], ref_rows[m], rtol=1e-6, atol=1e-12)
    assert np.allclose(info["r_history"], z["r_history"], rtol=1e-12)
    assert info["cost"] == pytest.approx(float(z["cost"]), rel=1e-8)
    assert rel_err(sol["mu"], z["sol_mu"]) < 1e-7


def test_stepwise_mode_records_every_residual_and_the_cost_each_iteration(golden):
    z, geo, n_time, kw = golden("ico2_nt7_stepwise")
    sol, info = orc.solve(n_time, geo, **kw)
    assert info["iterations"] == int(z["iterations"])
    assert not np.isnan(info["kkt_rows"]).any()
    assert np.allclose(info["kkt_rows"][:-1], z["kkt_rows"][:-1], rtol=1e-6, atol=1e-12)
    ref_cost = z["cost_history"]
    assert np.allclose(np.array(info["cost_history"])[:-1, 0], ref_cost[:-1], rtol=1e-8)


def test_synthetic_plane_is_the_references_plane_generator(golden):
    """dots_socp_b200.synth.hex_plane + normalize_geometry reproduce data/meshes/plane.py + data_preprocessing.py."""
    from dots_socp_b200 import synth
    z, geo, n_time, kw = golden("refplane20_nt15")
    mine, scale = synth.example("plane20")
    assert np.array_equal(mine["triangles"], geo["triangles"])
    assert np.allclose(mine["vertices"], geo["vertices"], atol=1e-14)
    assert np.allclose(mine["mu0"], geo["mu0"], rtol=1e-12) and np.allclose(mine["mu1"], geo["mu1"], rtol=1e-12)
    assert scale == pytest.approx(float(z["scale_factor"]), rel=1e-14)


def test_operator_identities():
    """Adjointness / structure checks the survey lists on a small icosphere."""
    from dots_socp_b200 import synth
    geo, _ = synth.example("icosphere1")
    ops = orc.MeshOps(5, geo, build_inverse=False)
    rng = np.random.default_rng(1)
    b = rng.standard_normal((6, ops.T, 3))
    x = rng.standard_normal((5, 2, 3, ops.T, 3))
    assert np.sum(orc.decouple(b, 1.7) * x) == pytest.approx(np.sum(b * orc.decouple_adjoint(x, 1.7)), rel=1e-12)
    assert np.allclose(orc.adjoint_time_average(np.array([[1.], [2.], [3.], [4.]]))[:, 0], [0.5, 1.5, 2.5, 3.5, 2.0])
    m = rng.standard_normal((5, ops.V)); p = rng.standard_normal((6, ops.V))
    assert np.sum(orc.grad_time(ops.dt, p) * m) == pytest.approx(-np.sum(p * orc.div_time(ops.dt, m)), rel=1e-12)
    # L = D diag(area_f x 1_3) G, symmetric, zero row sums
    W = np.repeat(ops.area_f, 3)
    L2 = ops.D @ (ops.G.multiply(W[:, None])).tocsr()
    assert abs(L2 - ops.L).max() < 1e-13
    assert abs(ops.L @ np.ones(ops.V)).max() < 1e-12


@pytest.mark.parametrize("tag,keys", [("full", None), ("part", ("phi", "beta_fst", "beta_end", "beta_mid"))])
def test_warm_start_matches_reference(golden, tag, keys):
    """``init_solution`` (solver_socp.py:239-250): restart from a coarse reference solution, all keys or a subset
    (the missing ones take the reference's defaults)."""
    z, geo, n_time, kw = golden("ico2_nt7_warm")
    init = {k[5:]: z[k] for k in z.files if k.startswith("init_") and (keys is None or k[5:] in keys)}
    sol, info = orc.solve(n_time, geo, init_solution=init, **kw)
    assert info["iterations"] == int(z[tag + "_iterations"])
    assert info["cost"] == pytest.approx(float(z[tag + "_cost"]), rel=1e-8)
    assert rel_err(sol["mu"], z[tag + "_mu"]) < 1e-7
    assert rel_err(sol["beta_mid"], z[tag + "_beta_mid"]) < 1e-7
