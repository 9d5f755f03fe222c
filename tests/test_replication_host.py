"""Host side of the replication harness (dots_socp_b200/replication.py), checked without a GPU:
the stand-in ``plane`` reproduces the reference's example bit for bit, the analytic transport / error functional / mass
diagnostics agree with values produced by the reference (fixture ``refplane20_exact``, tests/golden/make_golden.py), the
OFF writer round-trips through the reference's dialect, and a run's log parses into the replication table."""
import logging

import os

import numpy as np
import pytest
from conftest import GOLDEN_DIR

from dots_socp_b200 import replication as rep
from dots_socp_b200.history import RunHistory, LOG_INFO


@pytest.fixture
def plane20(golden):
    return golden("refplane20_nt15")[0], np.load(os.path.join(GOLDEN_DIR, "refplane20_exact.npz"))


def test_plane_standin_is_the_reference_example(plane20):
    fx, _ = plane20
    raw = rep.load_standin("plane", n_space=20)
    geo, scale = rep.synth.normalize_geometry(raw)
    np.testing.assert_array_equal(geo["triangles"], fx["triangles"])
    np.testing.assert_allclose(geo["vertices"], fx["vertices"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(geo["mu0"], fx["mu0"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(geo["mu1"], fx["mu1"], rtol=1e-13, atol=0)
    assert abs(scale - float(fx["scale_factor"])) < 1e-15


def test_exact_transport_and_error_functional_match_reference_values(plane20):
    _, ex = plane20
    raw = rep.load_standin("plane", n_space=20)
    np.testing.assert_allclose(raw["area_vertices"], ex["raw_area_vertices"], rtol=1e-13)
    exact = rep.exact_plane(np.linspace(0.0, 1.0, 16), raw["vertices"], raw["area_vertices"])
    np.testing.assert_allclose(exact, ex["exact"], rtol=1e-12, atol=1e-300)
    err = rep.compare_with_exact(ex["mu_centred"], exact, raw, verbose=False)
    for k in ("l1", "l2", "linf"):
        assert abs(err[k] - float(ex[k])) <= 1e-12 * float(ex[k])
    assert abs(rep.mass_conservation(ex["mu_centred"], verbose=False) - float(ex["mass_violation"])) < 1e-15
    neg, layers = rep.negative_mass(ex["mu_centred"], verbose=False)
    np.testing.assert_allclose(layers, ex["negative_layers"], rtol=1e-12, atol=1e-18)
    assert abs(neg - float(ex["negative_mass"])) < 1e-15


def test_automatic_checkpoints():
    assert rep.automatic_checkpoints(1e-5) == [10 ** (-i - 1) for i in range(5)]
    assert rep.automatic_checkpoints(1e-3) == [1e-1, 1e-2, 1e-3]
    assert rep.automatic_checkpoints(5e-4) == [1e-1, 1e-2, 1e-3]


def test_off_round_trip_is_bit_exact(tmp_path):
    v, t = rep.synth.deformed_sphere(2, (1.0, 0.5, 0.3), amp=0.1)
    path = tmp_path / "m.off"
    rep.write_off(path, v, t)
    v2, t2, e2 = rep.read_off(path)
    np.testing.assert_array_equal(v2, v)
    np.testing.assert_array_equal(t2, t)
    np.testing.assert_array_equal(e2, t[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2))


@pytest.mark.parametrize("text,msg", [("PLY\n", "Not a valid"), ("OFF\n3\n", "missing vertex/triangle"),
                                      ("OFF\n3 1 0\n0 0 0\n1 0 0\n3 0 1 2\n", "Expected 3 vertices"),
                                      ("OFF\n3 2 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n", "Expected 2 triangles"),
                                      ("OFF\n3 1 0\n0 0 0\n1 0\n", "Invalid vertex data")])
def test_off_reader_errors(tmp_path, text, msg):
    path = tmp_path / "bad.off"
    path.write_text(text)
    with pytest.raises(ValueError, match=msg):
        rep.read_off(path)


def test_every_standin_is_a_manifold_surface_of_the_intended_size():
    sizes = {}
    for name in rep.STANDINS:
        g = rep.load_standin(name)
        v, t = g["vertices"], g["triangles"]
        sizes[name] = v.shape[0]
        assert np.unique(t).size == v.shape[0], name                              # no unused vertices
        e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]), axis=1)
        _, cnt = np.unique(e, axis=0, return_counts=True)
        assert cnt.max() <= 2, name                                               # manifold edges
        assert (g["area_triangles"] > 0).all(), name
        for m in (g["mu0"], g["mu1"]):
            assert abs(m.sum() - 1.0) < 1e-12 and (m >= 0).all(), name
    assert min(sizes.values()) > 1500 and max(sizes.values()) == 68000
    assert sizes["knots_5"] == 4300 and sizes["plane"] == 11716


def _fake_solver(n_time, geometry, **kw):
    """CPU stand-in for the plug-in: linear interpolation of the masses and a two-row history."""
    V = geometry["vertices"].shape[0]
    w = np.linspace(0.0, 1.0, n_time + 1)[:, None]
    mu = (1 - w) * geometry["mu0"][None, :] + w * geometry["mu1"][None, :]
    hist = RunHistory(max_record_numbers=4, kkt_labels=[f"c{i}" for i in range(7)], name="SOCP", show_progress=False)
    hist.start()
    hist.record(current_it=0, kkt_errors=[1.0] * 7, history={"Transportation cost": 2.0, "Objective value": 2.0})
    hist.record(current_it=41, kkt_errors=[1e-4] * 7, history={"Transportation cost": 0.125, "Objective value": 0.125})
    hist.add_time("Step 1-3", 0.5)
    hist.end()
    cps = [dict(mu=mu.copy(), E=np.zeros((n_time + 1, geometry["triangles"].shape[0], 3)), iteration=7, time=0.1,
                kkt=np.array([1e-2, None, 5e-2, None, 1e-3, 1e-3, None], dtype=object))] if kw.get("tol_checkpoints") else None
    _fake_solver.last_kwargs = kw
    return dict(mu=mu, E=np.zeros((n_time + 1, 1, 3)), checkpoints=cps), hist


def test_run_log_parses_into_the_replication_table(tmp_path):
    info = tmp_path / "info.log"
    root = logging.getLogger()
    old_handlers, old_level = list(root.handlers), root.level
    fh = logging.FileHandler(info)
    fh.setFormatter(logging.Formatter("%(message)s"))
    root.handlers[:] = [fh]
    root.setLevel(LOG_INFO)
    try:
        for ex in ("ring", "knots_5"):
            opts = rep.options(example=ex, congestion=0.01, **rep.MAIN_FLAGS)
            rep.print_example_info(opts)
            sol, geo, hist = rep.run_example(opts, solver=_fake_solver)
        opts = rep.options(ntime=15, n_space=20, **rep.TRUE_ERROR_FLAGS)
        rep.print_example_info(opts)
        _, _, _, err, cps = rep.run_versus_exact(opts, solver=_fake_solver)
    finally:
        fh.close()
        root.handlers[:] = old_handlers
        root.setLevel(old_level)
    assert _fake_solver.last_kwargs["tol_checkpoints"] == rep.automatic_checkpoints(1e-5)
    assert set(_fake_solver.last_kwargs) == {"eps", "nit", "tol", "congestion", "tol_checkpoints", "time_limit",
                                             "check_kkt_step_by_step"}
    assert len(cps) == 1 and cps[0]["kkt_error"] == 5e-2 and cps[0]["iteration"] == 7
    assert 0 < err["l1"] < 1 and 0 < err["linf"] < 1
    rows = rep.table_rows(rep.parse_log(info))
    assert [r["Example"] for r in rows] == ["Knots 5", "Plane", "Ring"]
    knots = rows[0]
    assert knots["Vertices"] == 4300 and knots["Triangles"] == 8600 and knots["Iterations"] == 41
    assert knots["Time [seconds]"] == 0.5
    g = rep.load_standin("knots_5")
    _, scale = rep.synth.normalize_geometry(g)
    assert knots["Transport Cost"] == round(0.125 / scale ** 2, 4)                # cost de-scaled, interface.py:302-308
    text = info.read_text()
    for needle in ("Mass Conservation Violation:", "Non-Negative Mass Violation:", "L_Inf Error:", "Example name: plane"):
        assert needle in text
    assert "| Knots 5 | 4300 | 8600 | 41 |" in rep.markdown_table(rows)


@pytest.mark.parametrize("kw,msg", [(dict(ntime=0), "ntime"), (dict(tau=2.5), "tau"), (dict(tol=0.0), "tol"),
                                    (dict(congestion=-1.0), "congestion"), (dict(nit=0), "nit"), (dict(eps=-1e-9), "eps")])
def test_option_validation_follows_the_reference_caller(kw, msg):
    with pytest.raises(ValueError, match=msg):
        rep.run_example(rep.options(example="ring", **kw), solver=_fake_solver)
    with pytest.raises(TypeError):
        rep.run_example(rep.options(example="ring"), solver=3)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_written_off_file_is_read_identically_by_the_reference_reader(tmp_path):
    import sys
    sys.path.insert(0, GOLDEN_DIR)
    import refshim
    refshim.load()
    from dot_surface_socp.data.util import read_mesh_off
    v, t = rep.synth.hills(12)
    path = tmp_path / "hills.off"
    rep.write_off(path, v, t)
    rv, rt, re_ = read_mesh_off(str(path))
    mv, mt, me = rep.read_off(path)
    np.testing.assert_array_equal(rv, v)
    np.testing.assert_array_equal(rt, t)
    np.testing.assert_array_equal(mv, rv)
    np.testing.assert_array_equal(mt, rt)
    np.testing.assert_array_equal(me, re_)
