"""Synthetic surfaces and densities used by the tests and by bench.py.

The reference's bundled ``.off`` meshes are git-LFS stubs, so every
configuration in BASELINE.json is run on generated stand-ins:

* ``icosphere(level)``        - subdivided icosahedron, V = 10*4**level + 2.
* ``knot_tube()``             - (2,5) torus-knot tube, V = 4300 / T = 8600: the "knots_5-class" mesh.
* ``hex_plane(n)``            - flat hexagonal grid on [0,1]^2, the shape of the reference's
                                ``data/meshes/plane.py`` (reference dot_surface_socp/data/meshes/plane.py:3).

``make_geometry`` assembles the GeometryData dict (reference utils/type.py:6) and
``normalize_geometry`` mirrors what socp/data_preprocessing.py:5 hands to the solver
(centroid to origin, longest bounding-box edge scaled to 1, un-divided vertex areas).
All of it is host-side numpy: it feeds the hot path, it is not part of it.
"""
from __future__ import annotations

import numpy as np

from . import surface


# ----------------------------------------------------------------------------- meshes
def icosphere(level: int):
    """Unit-sphere icosphere: returns (vertices (V,3) f64, triangles (T,3) int64)."""
    g = (1.0 + 5.0 ** 0.5) / 2.0
    verts = np.array(
        [[-1, g, 0], [1, g, 0], [-1, -g, 0], [1, -g, 0],
         [0, -1, g], [0, 1, g], [0, -1, -g], [0, 1, -g],
         [g, 0, -1], [g, 0, 1], [-g, 0, -1], [-g, 0, 1]], dtype=np.float64)
    verts /= np.linalg.norm(verts, axis=1, keepdims=True)
    tris = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
         [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
         [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
         [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        n_v = verts.shape[0]
        # every undirected edge gets one midpoint vertex
        e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]], axis=0)
        e.sort(axis=1)
        key = e[:, 0] * n_v + e[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        mid = verts[uniq // n_v] + verts[uniq % n_v]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        verts = np.concatenate([verts, mid], axis=0)
        n_t = tris.shape[0]
        m01, m12, m20 = n_v + inv[:n_t], n_v + inv[n_t:2 * n_t], n_v + inv[2 * n_t:]
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        tris = np.concatenate([
            np.stack([a, m01, m20], axis=1),
            np.stack([b, m12, m01], axis=1),
            np.stack([c, m20, m12], axis=1),
            np.stack([m01, m12, m20], axis=1)], axis=0)
    return verts, tris


def knot_tube(p: int = 2, q: int = 5, big_r: float = 2.0, amp: float = 1.0,
              n_u: int = 430, n_v: int = 10, tube_r: float = 0.35):
    """Closed tube around a (p,q) torus knot; defaults give V=4300, T=8600."""
    th = 2.0 * np.pi * np.arange(n_u) / n_u

    def centre(t):
        rad = big_r + amp * np.cos(q * t)
        return np.stack([rad * np.cos(p * t), rad * np.sin(p * t), amp * np.sin(q * t)], axis=1)

    h = 1e-4
    d1 = (centre(th + h) - centre(th - h)) / (2 * h)
    d2 = (centre(th + h) - 2 * centre(th) + centre(th - h)) / (h * h)
    tan = d1 / np.linalg.norm(d1, axis=1, keepdims=True)
    nrm = d2 - np.sum(d2 * tan, axis=1, keepdims=True) * tan
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    bin_ = np.cross(tan, nrm)
    ang = 2.0 * np.pi * np.arange(n_v) / n_v
    c = centre(th)
    verts = (c[:, None, :]
             + tube_r * np.cos(ang)[None, :, None] * nrm[:, None, :]
             + tube_r * np.sin(ang)[None, :, None] * bin_[:, None, :]).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(n_u), np.arange(n_v), indexing="ij")
    i1, j1 = (i + 1) % n_u, (j + 1) % n_v
    vid = lambda a, b: (a * n_v + b).reshape(-1)
    t0 = np.stack([vid(i, j), vid(i1, j), vid(i1, j1)], axis=1)
    t1 = np.stack([vid(i, j), vid(i1, j1), vid(i, j1)], axis=1)
    return verts, np.concatenate([t0, t1], axis=0).astype(np.int64)


def hex_plane(n: int):
    """Hexagonal triangulation of (roughly) the unit square, n cells along x."""
    dx = 1.0 / n
    dy = dx * np.sqrt(3.0) / 2.0
    rows = int(1.0 / dy) + 1
    cols = n + 1
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    x = jj * dx + np.where(ii % 2 == 1, dx / 2.0, 0.0)
    verts = np.stack([x, ii * dy, np.zeros_like(x)], axis=-1).reshape(-1, 3)
    vid = lambda a, b: a * cols + b
    tris = []
    j = np.arange(cols - 1)
    for r in range(rows - 1):                # triangle order of the reference generator: per cell, first / second triangle
        if r % 2 == 0:
            first = np.stack([vid(r, j), vid(r, j + 1), vid(r + 1, j)], axis=1)
            second = np.stack([vid(r, j + 1), vid(r + 1, j + 1), vid(r + 1, j)], axis=1)
            tris.append(np.stack([first, second], axis=1).reshape(-1, 3))
        else:
            first = np.stack([vid(r, j), vid(r + 1, j + 1), vid(r + 1, j)], axis=1)
            second = np.stack([vid(r, j - 1), vid(r, j), vid(r + 1, j)], axis=1)
            both = np.stack([first, second], axis=1).reshape(-1, 3)
            tris.append(np.concatenate([first[:1], both[2:]], axis=0))      # cell j = 0 has no second triangle
    return verts, np.concatenate(tris, axis=0).astype(np.int64)


def torus(n_u: int = 100, n_v: int = 40, big_r: float = 1.0, small_r: float = 0.35):
    """Closed torus, V = n_u * n_v, T = 2V (stand-in for the reference's genus-1 example ``ring``)."""
    return knot_tube(p=1, q=0, big_r=big_r, amp=0.0, n_u=n_u, n_v=n_v, tube_r=small_r)


def deformed_sphere(level: int, axes=(1.0, 1.0, 1.0), amp: float = 0.0, freq: int = 3):
    """Icosphere pushed to an ellipsoid with semi-axes ``axes`` and a smooth radial ripple ``1 + amp*sin*sin*cos``:
    closed genus-0 stand-ins of the bundled scanned models (airplane, armadillo, hand, bunny) at matching sizes."""
    v, t = icosphere(level)
    rad = 1.0 + amp * np.sin(freq * np.pi * v[:, 0]) * np.sin(freq * np.pi * v[:, 1]) * np.cos(freq * np.pi * v[:, 2])
    return v * rad[:, None] * np.asarray(axes, dtype=np.float64)[None, :], t


def punctured_sphere(level: int, z_cut: float = 0.8):
    """Icosphere with the polar cap z > z_cut removed (one boundary loop; stand-in for ``punctured_ball``)."""
    v, t = icosphere(level)
    keep_t = (v[t][:, :, 2] <= z_cut).all(axis=1)
    t = t[keep_t]
    used = np.zeros(v.shape[0], dtype=bool)
    used[t.reshape(-1)] = True
    new_id = np.cumsum(used) - 1
    return v[used], new_id[t]


def hills(n: int = 60, amp: float = 0.15):
    """Height field over the hexagonal plane grid (open surface with boundary; stand-in for ``hills``)."""
    v, t = hex_plane(n)
    v = v.copy()
    v[:, 2] = amp * (np.sin(3 * np.pi * v[:, 0]) * np.sin(2 * np.pi * v[:, 1]) + 0.5 * np.cos(5 * np.pi * v[:, 0] * v[:, 1]))
    return v, t


# ----------------------------------------------------------------------------- densities
def _bump_mass(vertices, area_vertices, centres, width, cutoff=None):
    rho = np.zeros(vertices.shape[0])
    for c in np.atleast_2d(centres):
        d = np.linalg.norm(vertices - c[None, :], axis=1)
        g = np.exp(-d ** 2 / width)
        if cutoff is not None:
            g = np.where(d < cutoff, g, 0.0)
        rho += g
    m = rho * area_vertices
    return m / m.sum()


def gaussian_bump_masses(vertices, area_vertices, seed: int = 0, width: float = 0.1):
    """One bump for mu0, two for mu1, centres drawn from ONE default_rng(seed) and pushed to the unit sphere."""
    rng = np.random.default_rng(seed)
    c0 = rng.standard_normal((1, 3))
    c1 = rng.standard_normal((2, 3))
    c0 /= np.linalg.norm(c0, axis=1, keepdims=True)
    c1 /= np.linalg.norm(c1, axis=1, keepdims=True)
    scale = np.abs(vertices).max()
    return (_bump_mass(vertices, area_vertices, c0 * scale, width),
            _bump_mass(vertices, area_vertices, c1 * scale, width))


def vertex_bump_masses(vertices, area_vertices, seed: int = 0, rel_width: float = 0.02):
    """One bump for mu0, two for mu1, centred AT mesh vertices picked by default_rng(seed) (so they sit on the surface
    whatever its shape); width = rel_width * (bounding-box diagonal)^2."""
    rng = np.random.default_rng(seed)
    ids = rng.choice(vertices.shape[0], size=3, replace=False)
    width = rel_width * float(((vertices.max(axis=0) - vertices.min(axis=0)) ** 2).sum())
    return (_bump_mass(vertices, area_vertices, vertices[ids[:1]], width),
            _bump_mass(vertices, area_vertices, vertices[ids[1:]], width))


def knot_masses(vertices, area_vertices, ids=(2786, 1232, 406)):
    """Compact bumps centred at fixed vertex ids: the recipe of the reference's data/settings/knots_5.py:5-22."""
    n = vertices.shape[0]
    i0, i1, i2 = (k % n for k in ids)
    mu0 = _bump_mass(vertices, area_vertices, vertices[[i0]], 0.5, cutoff=0.5)
    mu1 = _bump_mass(vertices, area_vertices, vertices[[i1, i2]], 0.5, cutoff=0.5)
    return mu0, mu1


def plane_masses(vertices, area_vertices):
    """Gaussians of data/settings/plane.py:5-26 (centres (.4,.4) / (.6,.6), scale 2*0.1^2)."""
    mu0 = _bump_mass(vertices, area_vertices, np.array([[0.4, 0.4, 0.0]]), 2 * 0.1 ** 2)
    mu1 = _bump_mass(vertices, area_vertices, np.array([[0.6, 0.6, 0.0]]), 2 * 0.1 ** 2)
    return mu0, mu1


# ----------------------------------------------------------------------------- geometry dicts
def _edges_of(triangles):
    return triangles[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2)


def make_geometry(vertices, triangles, mu0=None, mu1=None, masses="gaussian", seed=0):
    """GeometryData-shaped dict (reference utils/type.py:6-13); ``area_vertices`` is the UN-divided incident-area sum."""
    area_f = surface.triangle_areas(vertices, triangles)
    area_v_sum = surface.incident_area_sum(vertices.shape[0], triangles, area_f)
    if mu0 is None:
        if masses == "gaussian":
            mu0, mu1 = gaussian_bump_masses(vertices, area_v_sum, seed)
        elif masses == "knot":
            mu0, mu1 = knot_masses(vertices, area_v_sum)
        elif masses == "plane":
            mu0, mu1 = plane_masses(vertices, area_v_sum)
        elif masses == "vertex":
            mu0, mu1 = vertex_bump_masses(vertices, area_v_sum, seed)
        else:
            raise ValueError(f"unknown masses recipe {masses!r}")
    return dict(vertices=np.ascontiguousarray(vertices, dtype=np.float64),
                triangles=np.ascontiguousarray(triangles, dtype=np.int64),
                edges=_edges_of(triangles), mu0=mu0, mu1=mu1,
                area_triangles=area_f, area_vertices=area_v_sum)


def normalize_geometry(geometry):
    """What the reference's caller does before invoking the solver (socp/data_preprocessing.py:5-43).

    Returns (normalized_geometry, scale_factor). The centroid is the area-weighted mean of the
    triangle centroids (trimesh's ``Trimesh.centroid``)."""
    v = np.array(geometry["vertices"], dtype=np.float64)
    t = np.asarray(geometry["triangles"])
    area_f = surface.triangle_areas(v, t)
    tri_c = v[t].mean(axis=1)
    centroid = (tri_c * area_f[:, None]).sum(axis=0) / area_f.sum()
    v = v - centroid
    scale = 1.0 / (v.max(axis=0) - v.min(axis=0)).max()
    v = v * scale
    v = v - v.min(axis=0)
    area_f = surface.triangle_areas(v, t)
    out = dict(vertices=v, triangles=t.copy(), edges=_edges_of(t),
               mu0=geometry["mu0"], mu1=geometry["mu1"], area_triangles=area_f,
               area_vertices=surface.incident_area_sum(v.shape[0], t, area_f))
    return out, scale


def example(name: str, **kw):
    """Named stand-ins: 'icosphere<L>', 'knot', 'plane<n>' -> normalized geometry + scale factor."""
    if name.startswith("icosphere"):
        v, t = icosphere(int(name[len("icosphere"):] or kw.get("level", 3)))
        g = make_geometry(v, t, masses="gaussian", seed=kw.get("seed", 0))
    elif name.startswith("knot"):
        v, t = knot_tube(n_u=kw.get("n_u", 430), n_v=kw.get("n_v", 10))
        g = make_geometry(v, t, masses="knot")
    elif name.startswith("plane"):
        v, t = hex_plane(int(name[len("plane"):] or kw.get("n", 20)))
        g = make_geometry(v, t, masses="plane")
    else:
        raise ValueError(f"unknown example {name!r}")
    return normalize_geometry(g)
