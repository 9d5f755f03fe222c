"""Host-side checks of the nested-dissection multifrontal setup (dots_socp_b200/nested.py).

The numpy sweep below walks the same panel layout the CUDA kernels stream, so a layout or index-map
bug shows up here without a GPU."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from dots_socp_b200 import nested, surface, synth
from host_multifrontal import factor_batched


def panel_rows(sym, panels, i):
    """Dense (s+b, s, M) view of node i's panel (upper triangle of the first block zero)."""
    s, b = int(sym.s[i]), int(sym.b[i])
    p0 = int(sym.panel_off[i])
    out = np.zeros((s + b, s, panels.shape[1]))
    tri = np.tril_indices(s)
    out[tri[0], tri[1]] = panels[p0:p0 + s * (s + 1) // 2]
    if b and s:
        out[s:] = panels[p0 + s * (s + 1) // 2:p0 + nested.panel_size(s, b)].reshape(b, s, -1)
    return out


def sweep_solve(sym, panels, rhs):
    """rhs: (n, M) in NEW ordering. Forward (leaves->root) then backward, mirroring the kernels."""
    n, M = rhs.shape
    y = np.zeros_like(rhs)
    upd = [None] * sym.n_nodes
    for i in np.argsort(sym.level, kind="stable"):
        s, b = int(sym.s[i]), int(sym.b[i])
        f0 = sym.front_off[i]
        P = panel_rows(sym, panels, i)
        r = np.zeros((s + b, M))
        r[:s] = rhs[sym.off[i]:sym.off[i] + s]
        for slot in range(2):
            k = sym.child[i, slot]
            if k >= 0 and sym.b[k]:
                cp = sym.child_pos[slot, f0:f0 + s + b]
                hit = cp >= 0
                r[hit] += upd[k][cp[hit]]
        out = np.einsum("ijm,jm->im", P, r[:s])
        y[sym.off[i]:sym.off[i] + s] = out[:s]
        upd[i] = r[s:] - out[s:]
    x = np.zeros_like(rhs)
    for i in np.argsort(-sym.level, kind="stable"):
        s, b = int(sym.s[i]), int(sym.b[i])
        f0 = sym.front_off[i]
        P = panel_rows(sym, panels, i)
        v = np.concatenate([y[sym.off[i]:sym.off[i] + s], -x[sym.front_idx[f0 + s:f0 + s + b]]], axis=0)
        x[sym.off[i]:sym.off[i] + s] = np.einsum("ijm,im->jm", P, v)
    return x


@pytest.mark.parametrize("example,leaf", [("icosphere2", 8), ("icosphere3", 24), ("plane8", 6), ("knot", 16)])
def test_batched_factor_solves_all_modes(example, leaf):
    kw = dict(n_u=60, n_v=6) if example == "knot" else {}
    geo, _ = synth.example(example, **kw)
    v, t = geo["vertices"], geo["triangles"]
    K = surface.stiffness_matrix(v, t)
    mass = surface.incident_area_sum(v.shape[0], t, surface.triangle_areas(v, t)) / 3.0
    sym = nested.analyse(v, K, leaf_size=leaf)
    assert sorted(sym.perm.tolist()) == list(range(v.shape[0]))
    n_time = 7
    k = np.arange(n_time + 1)
    shifts = 4.0 * n_time ** 2 * np.sin(np.pi * k / (2 * (n_time + 1))) ** 2
    panels = factor_batched(sym, K, mass, shifts, m_pad=8)
    rng = np.random.default_rng(0)
    rhs_old = rng.standard_normal((v.shape[0], 8))
    rhs_old[:, 0] -= rhs_old[:, 0].mean()                  # compatible rhs for the singular mode
    x_new = sweep_solve(sym, panels, rhs_old[sym.perm])
    x_old = np.empty_like(x_new)
    x_old[sym.perm] = x_new
    for m in range(1, 8):
        A = (K + shifts[m] * sp.diags(mass)).tocsc()
        ref = spla.spsolve(A, rhs_old[:, m])
        assert np.abs(x_old[:, m] - ref).max() / np.abs(ref).max() < 1e-11
    # mode 0: K x = b up to the pinned constant
    res = K @ x_old[:, 0] - rhs_old[:, 0]
    assert np.abs(res).max() / np.abs(rhs_old[:, 0]).max() < 1e-10


def test_levels_respect_tree():
    geo, _ = synth.example("icosphere3")
    K = surface.stiffness_matrix(geo["vertices"], geo["triangles"])
    sym = nested.analyse(geo["vertices"], K, leaf_size=16)
    for i in range(sym.n_nodes):
        for k in sym.child[i]:
            if k >= 0:
                assert sym.level[k] < sym.level[i] and sym.parent[k] == i
    assert sym.parent[sym.n_nodes - 1] == -1


def test_device_factorisation_matches_host_version():
    """factor_batched_device is the same algebra through torch (run on the CPU device here)."""
    geo, _ = synth.example("icosphere2")
    v, t = geo["vertices"], geo["triangles"]
    K = surface.stiffness_matrix(v, t)
    mass = surface.incident_area_sum(v.shape[0], t, surface.triangle_areas(v, t)) / 3.0
    sym = nested.analyse(v, K, leaf_size=8)
    shifts = np.array([0.0, 3.0, 11.0, 40.0, 170.0])
    a = factor_batched(sym, K, mass, shifts, m_pad=32)
    b = nested.factor_batched_device(sym, K, mass, shifts, m_pad=32, device="cpu").numpy()
    assert a.shape == b.shape
    assert np.abs(a - b).max() / np.abs(a).max() < 1e-12


def test_transposed_panel_copy_is_column_major_view_of_the_same_entries():
    geo, _ = synth.example("icosphere2")
    v, t = geo["vertices"], geo["triangles"]
    K = surface.stiffness_matrix(v, t)
    mass = surface.incident_area_sum(v.shape[0], t, surface.triangle_areas(v, t)) / 3.0
    sym = nested.analyse(v, K, leaf_size=8)
    shifts = np.array([0.0, 3.0, 11.0])
    p, pt = nested.factor_batched_device(sym, K, mass, shifts, m_pad=32, device="cpu", transposed=True)
    p, pt = p.numpy(), pt.numpy()
    for i in range(sym.n_nodes):
        s, b = int(sym.s[i]), int(sym.b[i])
        if s == 0:
            continue
        P = panel_rows(sym, p, i)                       # (s+b, s, M), upper triangle zero
        p0 = int(sym.panel_off[i])
        for j in range(s):
            col_off = j * (s + b) - j * (j - 1) // 2
            got = pt[p0 + col_off:p0 + col_off + (s + b - j)]
            assert np.array_equal(got, P[j:, j])


def test_front_maps_agree_with_the_per_node_search():
    geo, _ = synth.example("icosphere3")
    v, t = geo["vertices"], geo["triangles"]
    K = surface.stiffness_matrix(v, t)
    sym = nested.analyse(v, K, leaf_size=12)
    Kp = K[sym.perm][:, sym.perm].tocsr()
    Kp.sort_indices()
    a_pos, parent_pos = nested.front_maps(sym, Kp)
    for i in range(sym.n_nodes):
        s, b, lo, f0 = int(sym.s[i]), int(sym.b[i]), int(sym.off[i]), int(sym.front_off[i])
        rows = sym.front_idx[f0:f0 + s + b]
        q0, q1 = Kp.indptr[lo], Kp.indptr[lo + s]
        cols = Kp.indices[q0:q1]
        want = np.where(cols >= lo, np.searchsorted(rows, cols), -1)
        assert np.array_equal(a_pos[q0:q1], want)
        for slot in range(2):
            k = int(sym.child[i, slot])
            if k >= 0 and sym.b[k]:
                cp = sym.child_pos[slot, f0:f0 + s + b]
                pp = parent_pos[sym.upd_off[k]:sym.upd_off[k + 1]]
                assert np.array_equal(cp[pp], np.arange(sym.b[k]))


@pytest.mark.parametrize("example,leaf", [("icosphere2", 8), ("icosphere4", 16), ("plane20", 8), ("knot", 16), ("icosphere3", 1)])
def test_native_front_maps_equal_the_numpy_statement(example, leaf):
    """dots_front_maps (C++ merges) against nested.front_maps (global binary searches): identical index maps."""
    geo, _ = synth.example(example)
    v, t = geo["vertices"], geo["triangles"]
    K = surface.stiffness_matrix(v, t)
    sym = nested.analyse(v, K, leaf_size=leaf)
    Kp = K[sym.perm][:, sym.perm].tocsr()
    Kp.sort_indices()
    a_pos, parent_pos = nested.front_maps(sym, Kp)
    from dots_socp_b200 import capi
    a_nat, p_nat = nested.front_maps_native(capi.load(), sym, Kp)
    assert a_nat.dtype == a_pos.dtype and np.array_equal(a_nat, a_pos)
    assert p_nat.dtype == parent_pos.dtype and np.array_equal(p_nat, parent_pos)


def _same_symbolic(a, b):
    import dataclasses
    for f in dataclasses.fields(a):
        x, y = getattr(a, f.name), getattr(b, f.name)
        if isinstance(x, np.ndarray):
            assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y), f.name
        else:
            assert x == y, f.name


@pytest.mark.parametrize("example,leaf", [("icosphere1", 8), ("icosphere2", 8), ("icosphere3", 16), ("icosphere5", 24),
                                          ("plane8", 6), ("plane20", 8), ("knot", 16), ("icosphere4", 1)])
def test_native_ordering_is_identical_to_the_python_statement(example, leaf):
    """csrc/host_order.cpp (what the engine calls) against nested.dissect + nested.symbolic: same permutation, same
    separator tree, same front rows and child maps, bit for bit - so everything downstream sees the same inputs."""
    from dots_socp_b200 import capi
    geo, _ = synth.example(example)
    K = surface.stiffness_matrix(geo["vertices"], geo["triangles"])
    _same_symbolic(nested.analyse(geo["vertices"], K, leaf_size=leaf),
                   nested.analyse_native(capi.load(), geo["vertices"], K, leaf_size=leaf))


def test_native_ordering_on_open_and_higher_genus_surfaces():
    from dots_socp_b200 import capi
    for v, t in (synth.hills(24), synth.punctured_sphere(3), synth.torus(40, 12), synth.knot_tube(p=2, q=3, n_u=90, n_v=8)):
        K = surface.stiffness_matrix(v, t)
        _same_symbolic(nested.analyse(v, K, leaf_size=12), nested.analyse_native(capi.load(), v, K, leaf_size=12))


def test_threaded_host_analysis_does_not_depend_on_the_thread_count(monkeypatch):
    """csrc/host_order.cpp runs the bisection of the two halves of a part, the boundary sets of a tree level and the
    per-triangle / per-vertex passes of the mesh operators on DOTS_HOST_THREADS host threads once the mesh is large enough
    (V = 40 962 here: 5 threads at most).  Ordering, front structure, areas, hat gradients, stiffness matrix and corner lists
    must be the same bit for bit at every thread count, and the single-thread result is the numpy statement's."""
    from dots_socp_b200 import capi
    lib = capi.load()
    geo, _ = synth.example("icosphere6")
    v = np.ascontiguousarray(geo["vertices"], dtype=np.float64)
    t = np.ascontiguousarray(geo["triangles"], dtype=np.int64)
    results = {}
    for threads in (1, 2, 5, 16):
        monkeypatch.setenv("DOTS_HOST_THREADS", str(threads))
        m = surface.mesh_operators_native(lib, v, t)
        results[threads] = (m, nested.analyse_native(lib, v, m["K"], leaf_size=16))
    m1, sym1 = results[1]
    for threads, (m, sym) in results.items():
        _same_symbolic(sym1, sym)
        for key in ("area_f", "hat", "area_sum"):
            assert np.array_equal(m1[key], m[key]), (threads, key)
        assert all(np.array_equal(getattr(m1["K"], a), getattr(m["K"], a)) for a in ("indptr", "indices", "data")), threads
        assert all(np.array_equal(a, b) for a, b in zip(m1["corners"], m["corners"])), threads
    # the permuted matrix and the assembly index maps: threads over row / node ranges, same output
    maps = {}
    for threads in (1, 3, 16):
        monkeypatch.setenv("DOTS_HOST_THREADS", str(threads))
        Kp = nested.permuted_matrix(lib, sym1, m1["K"])
        maps[threads] = (Kp, nested.front_maps_native(lib, sym1, Kp))
    Kp_ref = nested.permuted_matrix(None, sym1, m1["K"])                       # scipy
    for threads, (Kp, (a_pos, parent_pos)) in maps.items():
        assert all(np.array_equal(getattr(Kp, a), getattr(Kp_ref, a)) for a in ("indptr", "indices", "data")), threads
        assert np.array_equal(a_pos, maps[1][1][0]) and np.array_equal(parent_pos, maps[1][1][1]), threads
    a_ref, p_ref = nested.front_maps(sym1, Kp_ref)
    assert np.array_equal(a_ref, maps[1][1][0]) and np.array_equal(p_ref, maps[1][1][1])
    K = surface.stiffness_matrix(v, t)
    _same_symbolic(nested.analyse(v, K, leaf_size=16), sym1)
    assert np.array_equal(m1["area_sum"], surface.incident_area_sum(v.shape[0], t, surface.triangle_areas(v, t)))
    assert np.array_equal(m1["K"].indices, K.indices) and np.abs(m1["K"].data - K.data).max() <= 1e-14 * np.abs(K.data).max()


@pytest.mark.parametrize("example,leaf", [("icosphere2", 8), ("plane8", 6), ("knot", 16)])
def test_native_matrix_permutation_equals_scipy(example, leaf):
    """dots_csr_permute against scipy's K[perm][:, perm] + sort_indices: same pattern, same values, sorted columns."""
    from dots_socp_b200 import capi
    geo, _ = synth.example(example)
    K = surface.stiffness_matrix(geo["vertices"], geo["triangles"])
    sym = nested.analyse(geo["vertices"], K, leaf_size=leaf)
    got, ref = nested.permuted_matrix(capi.load(), sym, K), nested.permuted_matrix(None, sym, K)
    assert all(np.array_equal(getattr(got, a), getattr(ref, a)) for a in ("indptr", "indices", "data"))
    assert np.all(np.diff(got.indices)[np.setdiff1d(np.arange(got.nnz - 1), got.indptr[1:-1] - 1)] > 0)
    bad = sym.perm.copy()
    bad[0] = -1
    with pytest.raises(capi.DotsError, match="perm"):
        nested.permuted_matrix(capi.load(), sym._replace(perm=bad) if hasattr(sym, "_replace") else
                               __import__("dataclasses").replace(sym, perm=bad), K)


def test_native_ordering_rejects_bad_arguments():
    from dots_socp_b200 import capi
    geo, _ = synth.example("icosphere1")
    K = surface.stiffness_matrix(geo["vertices"], geo["triangles"])
    with pytest.raises(capi.DotsError, match="leaf_size"):
        nested.analyse_native(capi.load(), geo["vertices"], K, leaf_size=0)
    with pytest.raises(ValueError):
        nested.analyse_native(capi.load(), geo["vertices"][:, :2], K, leaf_size=8)


@pytest.mark.parametrize("example,leaf,n_time", [("icosphere3", 8, 7), ("plane8", 6, 6), ("knot_small", 12, 15)])
def test_hybrid_driver_library_branch_matches_the_library_factorisation(example, leaf, n_time):
    """The library cross-check branch of ``factor_hybrid_device`` (``use_library=True``, DOTS_FACTOR=mixed; here forced for
    every front with front_nmax=0 so that no CUDA kernel is needed) against ``factor_batched_device``: same panels, both
    layouts.  Its index data is uploaded once and sliced per front; this pins that bookkeeping on the CPU."""
    from dots_socp_b200.engine import time_basis
    if example == "knot_small":
        v, t = synth.knot_tube(n_u=40, n_v=6)
    else:
        geo, _ = synth.example(example)
        v, t = geo["vertices"], geo["triangles"]
    K = surface.stiffness_matrix(v, t)
    mass = surface.incident_area_sum(v.shape[0], t, surface.triangle_areas(v, t)) / 3.0
    _, lam = time_basis(n_time)
    sym = nested.analyse(v, K, leaf_size=leaf)
    m_pad = 8 if n_time + 1 <= 8 else 16
    ref, ref_t = nested.factor_batched_device(sym, K, mass, -lam, m_pad, "cpu", transposed=True)
    stats = {}
    got, got_t = nested.factor_hybrid_device(sym, K, mass, -lam, m_pad, "cpu", None, lambda: 0, stats=stats, front_nmax=0,
                                             use_library=True)
    assert stats["small_fronts"] == 0 and stats["large_fronts"] == sym.n_nodes
    scale = np.abs(ref.numpy()).max()
    assert np.abs(got.numpy() - ref.numpy()).max() <= 1e-12 * scale
    assert np.abs(got_t.numpy() - ref_t.numpy()).max() <= 1e-12 * scale


@pytest.mark.parametrize("example", ["icosphere3", "plane8", "knot"])
def test_native_mesh_operators_equal_the_numpy_statement(example):
    """dots_mesh_* / dots_corner_lists (C++) against dots_socp_b200/surface.py (numpy): areas, hat gradients and incident
    area sums bit for bit, the cotan stiffness matrix with the same pattern and values to rounding (different summation
    order of the duplicates), the corner lists identical."""
    from dots_socp_b200 import capi
    lib = capi.load()
    geo, _ = synth.example(example)
    v, t = np.ascontiguousarray(geo["vertices"], dtype=np.float64), np.ascontiguousarray(geo["triangles"], dtype=np.int64)
    m = surface.mesh_operators_native(lib, v, t)
    af = surface.triangle_areas(v, t)
    assert np.array_equal(m["area_f"], af)
    assert np.array_equal(m["hat"], surface.hat_gradients(v, t))
    assert np.array_equal(m["area_sum"], surface.incident_area_sum(v.shape[0], t, af))
    K = surface.stiffness_matrix(v, t)
    assert np.array_equal(m["K"].indptr, K.indptr) and np.array_equal(m["K"].indices, K.indices)
    assert np.abs(m["K"].data - K.data).max() <= 1e-14 * np.abs(K.data).max()
    ptr, tri_of, corner_of = surface.corner_adjacency(v.shape[0], t)
    assert all(np.array_equal(a, b) for a, b in zip(m["corners"], (ptr, tri_of, corner_of)))
    cp, ci = np.empty(v.shape[0] + 1, np.int32), np.empty(3 * t.shape[0], np.int32)
    assert lib.dots_corner_lists(v.shape[0], t.shape[0], t.ctypes.data, cp.ctypes.data, ci.ctypes.data) == 0
    assert np.array_equal(cp, ptr) and np.array_equal(ci, corner_of * t.shape[0] + tri_of)


def test_host_analysis_is_clean_under_thread_sanitizer(tmp_path):
    """csrc/host_order.cpp + tests/host_tsan_main.cpp built with g++ -fsanitize=thread: no data race reported at 1, 5 and 16
    requested threads, and the same output hash at every thread count."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "tsan_host")
    cmd = [gxx, "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-I", os.path.join(root, "include"),
           os.path.join(root, "tests", "host_tsan_main.cpp"), os.path.join(root, "dots_socp_b200", "csrc", "host_order.cpp"), "-o", exe]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0:
        pytest.skip("thread sanitizer runtime not available: " + built.stderr[-300:])
    hashes = set()
    for threads in ("1", "5", "16"):
        res = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, DOTS_HOST_THREADS=threads), timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        assert "ThreadSanitizer" not in res.stderr, res.stderr[-2000:]
        hashes.add(res.stdout.strip().split("hash=")[1])
    assert len(hashes) == 1
