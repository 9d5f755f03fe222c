"""Mesh operators of the hot path, vectorised (host-side setup; the per-iteration work is CUDA).

Replaces the Python per-triangle loops of the reference's
``utils/surface_pre_computations_socp.py`` (geometricQuantities :11-39, geometricMatrices :42-86,
trianglesToVertices :88-132) with whole-array numpy:

* ``triangle_areas``        |f|                                             (ref :24)
* ``hat_gradients``         P1 basis gradients g[f,k,:] = grad of the hat function of corner k (ref :30-37)
* ``corner_cotangents``     cot of the angle at corner k                    (ref :26-28, :68)
* ``stiffness_matrix``      K = -L, the positive semidefinite cotan matrix  (ref :68-84)
* ``incident_area_sum``     sum of |f| over the triangles around a vertex   (ref :121-124)
* ``corner_adjacency``      CSR vertex -> incident (triangle, corner) list  (ref :112-127, the two incidence maps)
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _edge_vectors(vertices, triangles):
    p0, p1, p2 = vertices[triangles[:, 0]], vertices[triangles[:, 1]], vertices[triangles[:, 2]]
    return p1 - p0, p2 - p1, p0 - p2          # e01, e12, e20


def triangle_areas(vertices, triangles):
    e01, e12, _ = _edge_vectors(vertices, triangles)
    return 0.5 * np.linalg.norm(np.cross(e01, e12), axis=1)


def hat_gradients(vertices, triangles):
    """g[f,k,:]: gradient (a 3-vector in the triangle's plane) of the P1 hat function of corner k.

    It is the altitude vector from the opposite edge to corner k divided by its squared length."""
    e01, e12, e20 = _edge_vectors(vertices, triangles)

    def altitude(into, along):
        # component of -into orthogonal to along
        coef = np.sum(into * along, axis=1) / np.sum(along * along, axis=1)
        h = -into + along * coef[:, None]
        return h / np.sum(h * h, axis=1)[:, None]

    return np.stack([altitude(e01, e12), altitude(e12, e20), altitude(e20, e01)], axis=1)


def corner_cotangents(vertices, triangles):
    e01, e12, e20 = _edge_vectors(vertices, triangles)

    def cot(a, b):
        return np.sum(a * b, axis=1) / np.linalg.norm(np.cross(a, b), axis=1)

    return np.stack([cot(e01, -e20), cot(e12, -e01), cot(e20, -e12)], axis=1)


def stiffness_matrix(vertices, triangles):
    """K (V x V, CSR, symmetric PSD, zero row sums): K = -L with L the reference's cotan Laplacian."""
    n_v = vertices.shape[0]
    w = 0.5 * corner_cotangents(vertices, triangles)
    rows, cols, vals = [], [], []
    for k in range(3):                       # the angle at corner k weights the opposite edge (a, b)
        a, b = triangles[:, (k + 1) % 3], triangles[:, (k + 2) % 3]
        rows += [a, b, a, b]
        cols += [b, a, a, b]
        vals += [-w[:, k], -w[:, k], w[:, k], w[:, k]]
    K = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_v, n_v))
    K = K.tocsr()
    K.sum_duplicates()
    K.sort_indices()
    return K


def incident_area_sum(n_vertices, triangles, area_f):
    return np.bincount(triangles.reshape(-1), weights=np.repeat(area_f, 3), minlength=n_vertices)


def corner_adjacency(n_vertices, triangles):
    """CSR lists of the corners around each vertex.

    Returns (ptr (V+1,), tri_of (3T,), corner_of (3T,)): for vertex v the incident corners are
    entries ptr[v]:ptr[v+1], sorted by (corner, triangle) like the rows of the reference's
    vertex<-corner map (column index k*T+f)."""
    n_t = triangles.shape[0]
    col = np.arange(3 * n_t)                     # column k*T+f
    vert = triangles.T.reshape(-1)               # vertex of corner (k, f)
    order = np.argsort(vert, kind="stable")
    counts = np.bincount(vert, minlength=n_vertices)
    ptr = np.zeros(n_vertices + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    sorted_cols = col[order]
    return ptr, (sorted_cols % n_t).astype(np.int64), (sorted_cols // n_t).astype(np.int64)


def mesh_operators_native(lib, vertices, triangles):
    """All of the above in one pass of the C++ helper ``dots_mesh_*`` (csrc/host_order.cpp): returns a dict with
    ``area_f`` (T,), ``hat`` (T,3,3), ``area_sum`` (V,), ``K`` (CSR, sorted columns) and the corner lists
    ``(c_ptr, c_tri, c_corner)`` of the GIVEN triangle numbering.  The numpy functions of this module are the readable
    statement and the checker (tests/test_nested_host.py)."""
    import ctypes as C
    from . import capi
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    t = np.ascontiguousarray(triangles, dtype=np.int64)
    n_v, n_t = v.shape[0], t.shape[0]
    handle = C.c_void_p()
    capi.check(lib.dots_mesh_create(n_v, n_t, v.ctypes.data, t.ctypes.data, C.byref(handle)), "dots_mesh_create")
    try:
        nnz = C.c_int64()
        capi.check(lib.dots_mesh_sizes(handle, C.byref(nnz)), "dots_mesh_sizes")
        area_f, hat, area_sum = np.empty(n_t), np.empty((n_t, 3, 3)), np.empty(n_v)
        k_ptr, k_idx, k_val = np.empty(n_v + 1, np.int64), np.empty(nnz.value, np.int64), np.empty(nnz.value)
        c_ptr, c_tri, c_corner = np.empty(n_v + 1, np.int64), np.empty(3 * n_t, np.int64), np.empty(3 * n_t, np.int64)
        capi.check(lib.dots_mesh_export(handle, *(a.ctypes.data for a in (area_f, hat, area_sum, k_ptr, k_idx, k_val, c_ptr, c_tri, c_corner))),
                   "dots_mesh_export")
    finally:
        lib.dots_mesh_destroy(handle)
    K = sp.csr_matrix((k_val, k_idx, k_ptr), shape=(n_v, n_v))
    K.has_sorted_indices = True
    return dict(area_f=area_f, hat=hat, area_sum=area_sum, K=K, corners=(c_ptr, c_tri, c_corner))
