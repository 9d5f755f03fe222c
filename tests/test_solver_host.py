"""Host-side behaviour of the solver plug-in that can be checked without a GPU: argument validation in the order and
wording of the reference (socp/solver_socp.py:85-94), the solver-only knobs, the loud failure
when no CUDA device exists, and the numpy DOT-unit translation used for checkpoints (utils/type.py:48-65,
socp/solver_decorator.py:32-34)."""
import numpy as np
import pytest
import torch

import dots_socp_b200 as b200
import importlib

from dots_socp_b200 import capi, synth

solver_mod = importlib.import_module("dots_socp_b200.solver")      # the package re-exports the function `solver`

no_gpu = pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a box without a GPU")


@pytest.fixture(scope="module")
def geo():
    return synth.example("icosphere1")[0]


@no_gpu
def test_solver_only_knobs_reach_the_engine(geo):
    """is_palm / is_constant_scaling are built (no NotImplementedError any more): without a GPU the call gets as far as the
    engine, which refuses loudly (no CPU fallback)."""
    for kw in (dict(is_constant_scaling=True), dict(is_palm=True)):
        with pytest.raises(capi.DotsError, match="CUDA"):
            b200.solver_socp(3, geo, **kw)


@pytest.mark.parametrize("cps,msg", [([], "non-empty list"), ((1e-2,), "non-empty list"), ([1.5], "between 0 and 1"),
                                     ([0], "between 0 and 1"), (["a"], "between 0 and 1"), ([1e-5], "greater than tol")])
def test_checkpoint_validation(geo, cps, msg):
    with pytest.raises(ValueError, match=msg):
        b200.solver_socp(3, geo, tol=1e-4, tol_checkpoints=cps)


def test_checkpoints_are_sorted_descending():
    assert solver_mod._validate_checkpoints([1e-3, 1e-1, 1e-2], 1e-4) == [1e-1, 1e-2, 1e-3]
    assert solver_mod._validate_checkpoints(None, 1e-4) is None


@no_gpu
@pytest.mark.parametrize("fn", ["solver_socp", "solver_raw", "solver"])
def test_every_entry_point_fails_loudly_without_cuda(geo, fn):
    with pytest.raises(capi.DotsError, match="no CPU fallback"):
        getattr(b200, fn)(3, geo)


def test_plugin_names_match_the_reference_decorators():
    assert b200.solver_raw.__name__ == "dot_solver_socp"
    assert b200.solver.__name__ == "dot_solver_socp_center"


def test_host_translation_of_checkpoints(geo):
    rng = np.random.default_rng(3)
    nT, V, T = 4, geo["vertices"].shape[0], geo["triangles"].shape[0]
    cp = dict(mu=rng.random((nT, V)), E=rng.random((nT, T, 3)), iteration=7, time=0.5, kkt=np.arange(7, dtype=object))
    sol = dict(mu=cp["mu"].copy(), E=cp["E"].copy(), checkpoints=[cp])
    dot = solver_mod.translate_solution_socp_to_dot(sol, geo)
    np.testing.assert_array_equal(dot["mu"], cp["mu"] * (geo["area_vertices"][None, :] / 3.0))      # reference order
    np.testing.assert_array_equal(dot["E"], cp["E"] * geo["area_triangles"][None, :, None])
    np.testing.assert_array_equal(dot["checkpoints"][0]["mu"], dot["mu"])
    assert dot["checkpoints"][0]["iteration"] == 7

    sol2 = dict(checkpoints=[dict(cp)])
    solver_mod._dot_checkpoints(sol2, geo, centred=True)
    c = sol2["checkpoints"][0]
    assert c["mu"].shape == (nT + 1, V)
    np.testing.assert_array_equal(c["mu"][0], geo["mu0"])
    np.testing.assert_array_equal(c["mu"][-1], geo["mu1"])
    np.testing.assert_allclose(c["mu"][1:-1], 0.5 * (dot["mu"][:-1] + dot["mu"][1:]), rtol=0, atol=0)

    sol3 = dict(checkpoints=None)
    solver_mod._dot_checkpoints(sol3, geo, centred=False)
    assert "checkpoints" not in sol3


@pytest.mark.parametrize("n,bound", [(100, 50), (1 << 14, 1 << 10), (40000, 1 << 16), (50000, 70000), (330000, 163842),
                                     (20000, (1 << 32) + 5)])
def test_triangle_renumbering_sort_is_numpys_stable_argsort(n, bound):
    """Engine orders the triangles by their smallest new vertex id with two radix passes over 16-bit halves; the triangle
    numbering (and with it every triangle-indexed array) must be exactly that of ``np.argsort(kind="stable")``."""
    from dots_socp_b200.engine import _stable_argsort
    keys = np.random.default_rng(n).integers(0, bound, n)
    keys[::7] = keys[0]                                               # many ties: stability matters
    assert np.array_equal(_stable_argsort(keys, bound), np.argsort(keys, kind="stable"))
