"""Host-side setup of the per-mode shifted surface-Laplacian solves (row a8 of SURVEY.md section 8).

The reference factorises ``L + (lambda_a - eps) diag(area_v)`` once per time mode with SuperLU
(utils/laplacian_inverse_socp.py:34-41) and calls the ``nT+1`` solve closures one after another
(:58-59).  All those matrices share one sparsity pattern and differ by a diagonal shift, so here
they are factorised TOGETHER, batched over the mode index, with one nested-dissection ordering:

* ``dissect``        geometric nested dissection (recursive coordinate bisection, vertex separators)
                     -> elimination order + separator tree.
* ``symbolic``       multifrontal structure: every tree node owns a contiguous block S of the new
                     ordering and a boundary set B of ancestor vertices; its front is dense
                     (|S|+|B|)^2.
* ``factor_hybrid_device`` / ``factor_batched_device``
                     numeric multifrontal Cholesky of ``K + shift_m * diag(mass)`` for all modes m at
                     once on the GPU (K = -L is the PSD cotan stiffness matrix; small fronts in the
                     hand-written ``k_front_small``, large ones through batched library Cholesky; a plain
                     numpy version of the same algebra lives in tests/host_multifrontal.py as the
                     checker).  What is stored per node is the
                     *solve-ready* panel  P = [ inv(L11) ; L21 inv(L11) ]  (lower triangle of the first
                     block, |B| x |S| second block) laid out ``[row][col][mode]`` with the mode index
                     fastest, which is the layout the CUDA level-scheduled kernels stream.

With that panel the forward sweep is one dense mat-vec per node, ``[y_S ; -du_B] = P r_S`` and the
backward sweep is its transpose: no sequential triangular recurrences are left inside a node, the
only dependencies are parent/child ones (tree levels).

Mode 0 (shift 0, eps = 0) is singular along the constant vector: the last pivot is pinned, which
returns the solution with that vertex at 0 (the reference returns an arbitrary constant, SURVEY.md
appendix B); everything downstream only uses differences of phi.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp


# ----------------------------------------------------------------------------- ordering
@dataclass
class _Node:
    own: np.ndarray                       # old vertex ids owned (eliminated) by this node
    kids: list = field(default_factory=list)


def dissect(coords: np.ndarray, adj: sp.csr_matrix, leaf_size: int = 24):
    """Recursive coordinate bisection with vertex separators.  Returns the root ``_Node``.

    Iterative (explicit stack) so deep trees do not hit the recursion limit."""
    n = coords.shape[0]
    side = np.full(n, -1, dtype=np.int8)          # scratch: -1 outside the current subset
    indptr, indices = adj.indptr, adj.indices
    root = _Node(own=np.empty(0, dtype=np.int64))
    stack = [(root, np.arange(n, dtype=np.int64))]
    while stack:
        node, verts = stack.pop()
        if verts.size <= leaf_size:
            node.own = verts
            continue
        x = coords[verts]
        axis = int(np.argmax(x.max(axis=0) - x.min(axis=0)))
        order = np.argsort(x[:, axis], kind="stable")
        half = verts.size // 2
        left, right = verts[order[:half]], verts[order[half:]]
        side[left], side[right] = 0, 1
        # boundary of each half: vertices with a neighbour in the other half
        def frontier(part, other_side):
            starts, ends = indptr[part], indptr[part + 1]
            counts = ends - starts
            rows = np.repeat(np.arange(part.size), counts)
            offs = np.arange(counts.sum()) - np.repeat(np.cumsum(counts) - counts, counts)
            nbr = indices[np.repeat(starts, counts) + offs]
            hit = side[nbr] == other_side
            mask = np.zeros(part.size, dtype=bool)
            mask[rows[hit]] = True
            return mask
        fl, fr = frontier(left, 1), frontier(right, 0)
        if fl.sum() <= fr.sum():
            sep, left = left[fl], left[~fl]
        else:
            sep, right = right[fr], right[~fr]
        side[verts] = -1
        node.own = sep
        for part in (left, right):
            if part.size:
                kid = _Node(own=np.empty(0, dtype=np.int64))
                node.kids.append(kid)
                stack.append((kid, part))
    return root


# ----------------------------------------------------------------------------- symbolic structure
@dataclass
class Symbolic:
    n: int
    perm: np.ndarray            # new -> old vertex id
    iperm: np.ndarray           # old -> new
    n_nodes: int
    off: np.ndarray             # first new index owned by node
    s: np.ndarray               # |S|
    b: np.ndarray               # |B|
    level: np.ndarray           # 0 = leaves ... root = max
    parent: np.ndarray
    child: np.ndarray           # (n_nodes, 2) ids or -1
    front_off: np.ndarray       # offset of the node's front rows in ``front_idx``   (n_nodes+1)
    front_idx: np.ndarray       # new vertex index of every front row (S rows then B rows, ascending)
    child_pos: np.ndarray       # (2, sum n_i): row of child slot's update vector feeding this front row, or -1
    panel_off: np.ndarray       # offset of the node's panel, in entries (n_nodes+1)
    upd_off: np.ndarray         # offset of the node's update vector, in rows      (n_nodes+1)

    @property
    def n_levels(self):
        return int(self.level.max()) + 1

    @property
    def panel_entries(self):
        return int(self.panel_off[-1])


def panel_size(s, b):
    return s * (s + 1) // 2 + b * s


def symbolic(root: _Node, adj: sp.csr_matrix) -> Symbolic:
    n = adj.shape[0]
    # post-order (children before parents), iterative
    order, stack = [], [(root, False)]
    while stack:
        node, seen = stack.pop()
        if seen:
            order.append(node)
        else:
            stack.append((node, True))
            for kid in reversed(node.kids):
                stack.append((kid, False))
    ids = {id(nd): i for i, nd in enumerate(order)}
    n_nodes = len(order)
    perm = np.concatenate([nd.own for nd in order]).astype(np.int64)
    assert perm.size == n and np.unique(perm).size == n, "dissection must cover every vertex exactly once"
    iperm = np.empty(n, dtype=np.int64)
    iperm[perm] = np.arange(n)
    s = np.array([nd.own.size for nd in order], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(s)[:-1]]).astype(np.int64)
    parent = np.full(n_nodes, -1, dtype=np.int64)
    child = np.full((n_nodes, 2), -1, dtype=np.int64)
    level = np.zeros(n_nodes, dtype=np.int64)
    for i, nd in enumerate(order):
        for slot, kid in enumerate(nd.kids):
            k = ids[id(kid)]
            parent[k] = i
            child[i, slot] = k
            level[i] = max(level[i], level[k] + 1)
    adj_new = adj[perm][:, perm].tocsr()
    adj_new.sort_indices()
    indptr, indices = adj_new.indptr, adj_new.indices
    bsets = [None] * n_nodes
    b = np.zeros(n_nodes, dtype=np.int64)
    for i in range(n_nodes):                                    # post-order: children are ready
        lo, hi = off[i], off[i] + s[i]
        nbr = indices[indptr[lo]:indptr[hi]]
        parts = [nbr[nbr >= hi]]
        for k in child[i]:
            if k >= 0:
                bk = bsets[k]
                parts.append(bk[bk >= hi])
        bsets[i] = np.unique(np.concatenate(parts)) if parts else np.empty(0, dtype=np.int64)
        b[i] = bsets[i].size
    nfront = s + b
    front_off = np.concatenate([[0], np.cumsum(nfront)]).astype(np.int64)
    front_idx = np.empty(front_off[-1], dtype=np.int64)
    child_pos = np.full((2, front_off[-1]), -1, dtype=np.int64)
    for i in range(n_nodes):
        f0 = front_off[i]
        rows = np.concatenate([np.arange(off[i], off[i] + s[i]), bsets[i]])
        front_idx[f0:f0 + nfront[i]] = rows
        for slot, k in enumerate(child[i]):
            if k >= 0 and b[k]:
                where = np.searchsorted(rows, bsets[k])
                assert np.array_equal(rows[where], bsets[k]), "child boundary must embed in the parent front"
                child_pos[slot, f0 + where] = np.arange(b[k])
    panel_off = np.concatenate([[0], np.cumsum([panel_size(int(a), int(c)) for a, c in zip(s, b)])]).astype(np.int64)
    upd_off = np.concatenate([[0], np.cumsum(b)]).astype(np.int64)
    return Symbolic(n=n, perm=perm, iperm=iperm, n_nodes=n_nodes, off=off, s=s, b=b, level=level, parent=parent,
                    child=child, front_off=front_off, front_idx=front_idx, child_pos=child_pos,
                    panel_off=panel_off, upd_off=upd_off)


def _pattern(K: sp.csr_matrix) -> sp.csr_matrix:
    adj = K.copy().tocsr()
    adj.setdiag(0)
    adj.eliminate_zeros()
    return adj


def analyse(vertices: np.ndarray, K: sp.csr_matrix, leaf_size: int = 24) -> Symbolic:
    """Ordering + symbolic structure, plain numpy (the readable statement of the algorithm and the checker of
    ``analyse_native``, which is what the engine calls)."""
    adj = _pattern(K)
    return symbolic(dissect(np.asarray(vertices, dtype=np.float64), adj, leaf_size), adj)


def analyse_native(lib, vertices: np.ndarray, K: sp.csr_matrix, leaf_size: int = 24) -> Symbolic:
    """Same result as ``analyse`` (bit for bit) from the C++ implementation in libdots_b200.so
    (csrc/host_order.cpp, ``dots_order_*`` of include/dots_b200.h): ~20x faster at V = 164k."""
    import ctypes as C
    from . import capi
    adj = _pattern(K)
    n = adj.shape[0]
    xyz = np.ascontiguousarray(vertices, dtype=np.float64)
    if xyz.shape != (n, 3):
        raise ValueError(f"vertices must be ({n}, 3), got {xyz.shape}")
    ptr = np.ascontiguousarray(adj.indptr, dtype=np.int64)
    idx = np.ascontiguousarray(adj.indices, dtype=np.int64)
    handle = C.c_void_p()
    capi.check(lib.dots_order_create(n, xyz.ctypes.data, ptr.ctypes.data, idx.ctypes.data, int(leaf_size), C.byref(handle)),
               "dots_order_create")
    try:
        n_nodes, front_total = C.c_int64(), C.c_int64()
        capi.check(lib.dots_order_sizes(handle, C.byref(n_nodes), C.byref(front_total)), "dots_order_sizes")
        n_nodes, front_total = n_nodes.value, front_total.value
        i64 = lambda *shape: np.empty(shape, dtype=np.int64)
        perm, s, b, level, parent = i64(n), i64(n_nodes), i64(n_nodes), i64(n_nodes), i64(n_nodes)
        child, front_idx, child_pos = i64(n_nodes, 2), i64(front_total), i64(2, front_total)
        capi.check(lib.dots_order_export(handle, *(a.ctypes.data for a in (perm, s, b, level, parent, child, front_idx, child_pos))),
                   "dots_order_export")
    finally:
        lib.dots_order_destroy(handle)
    iperm = np.empty(n, dtype=np.int64)
    iperm[perm] = np.arange(n)
    cum = lambda a: np.concatenate([[0], np.cumsum(a)]).astype(np.int64)
    return Symbolic(n=n, perm=perm, iperm=iperm, n_nodes=int(n_nodes), off=cum(s)[:-1], s=s, b=b, level=level, parent=parent,
                    child=child, front_off=cum(s + b), front_idx=front_idx, child_pos=child_pos,
                    panel_off=cum(s * (s + 1) // 2 + b * s), upd_off=cum(b))


# ----------------------------------------------------------------------------- level schedule for the kernels
def level_schedule(sym: Symbolic):
    """Nodes grouped by tree level (leaves first).  Returns list of int64 arrays of node ids."""
    return [np.nonzero(sym.level == lv)[0].astype(np.int64) for lv in range(sym.n_levels)]


def factor_batched_device(sym: Symbolic, K: sp.csr_matrix, mass: np.ndarray, shifts: np.ndarray, m_pad: int, device,
                          pin_singular: bool = True, transposed: bool = False):
    """Numeric factorisation of all modes with the dense front algebra in batched torch calls on ``device``.

    SETUP path (row (f)1 of SURVEY.md section 8), not the per-iteration hot path: the batched dense
    Cholesky / triangular-solve / GEMM calls go through torch.linalg (cuSOLVER / cuBLAS) on fp64 tensors
    of shape (modes, n, n).  Returns the panel tensor (panel_entries, m_pad) on ``device``; with
    ``transposed=True`` also the column-major copy (per node: column j holds rows j..s+b-1 of the stacked panel)
    that the backward sweep streams."""
    import torch

    shifts_t = torch.as_tensor(np.asarray(shifts, dtype=np.float64), device=device)
    n_modes = shifts_t.numel()
    Kp = K[sym.perm][:, sym.perm].tocsr()
    Kp.sort_indices()
    massp = torch.as_tensor(np.asarray(mass, dtype=np.float64)[sym.perm], device=device)
    indptr, indices, data = Kp.indptr, Kp.indices, Kp.data
    panels = torch.zeros((sym.panel_entries, m_pad), dtype=torch.float64, device=device)
    panels_t = torch.zeros((sym.panel_entries, m_pad), dtype=torch.float64, device=device) if transposed else None
    triu_cache = {}
    singular = [m for m in range(n_modes) if float(shifts[m]) == 0.0] if pin_singular else []
    pin_value = float(Kp.diagonal().mean())
    updates = [None] * sym.n_nodes
    tril_cache = {}
    dev_i64 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int64), device=device)
    for i in range(sym.n_nodes):
        s, b = int(sym.s[i]), int(sym.b[i])
        nf = s + b
        lo = int(sym.off[i])
        f0 = int(sym.front_off[i])
        rows = sym.front_idx[f0:f0 + nf]
        F = torch.zeros((n_modes, nf, nf), dtype=torch.float64, device=device)
        if s:
            a0, a1 = indptr[lo], indptr[lo + s]
            r = np.repeat(np.arange(s), np.diff(indptr[lo:lo + s + 1]))
            c_new, v = indices[a0:a1], data[a0:a1]
            keep = c_new >= lo
            r, c_new, v = r[keep], c_new[keep], v[keep]
            c = np.searchsorted(rows, c_new)
            rt, ct = dev_i64(r), dev_i64(c)
            vt = torch.as_tensor(v, device=device)
            F[:, rt, ct] = vt
            F[:, ct, rt] = vt
            d = torch.arange(s, device=device)
            F[:, d, d] += shifts_t[:, None] * massp[None, lo:lo + s]
            if i == sym.n_nodes - 1:
                for m in singular:
                    F[m, s - 1, s - 1] += pin_value
        for slot in range(2):
            k = int(sym.child[i, slot])
            if k >= 0 and sym.b[k]:
                cp = sym.child_pos[slot, f0:f0 + nf]
                where = np.nonzero(cp >= 0)[0]
                where = dev_i64(where[np.argsort(cp[where])])
                F[:, where[:, None], where[None, :]] += updates[k]
                updates[k] = None
        if s == 0:
            updates[i] = F if b else None
            continue
        L11 = torch.linalg.cholesky(F[:, :s, :s])
        eye = torch.eye(s, dtype=torch.float64, device=device).expand(n_modes, s, s)
        Linv = torch.linalg.solve_triangular(L11, eye, upper=False)
        if s not in tril_cache:
            tril_cache[s] = torch.tril_indices(s, s, device=device)
        tri = tril_cache[s]
        p0 = int(sym.panel_off[i])
        ntri = s * (s + 1) // 2
        panels[p0:p0 + ntri, :n_modes] = Linv[:, tri[0], tri[1]].T
        if m_pad > n_modes:
            dd = torch.arange(s, device=device)
            panels[p0 + dd * (dd + 1) // 2 + dd, n_modes:] = 1.0
        W21 = None
        if b:
            L21 = F[:, s:, :s] @ Linv.mT
            W21 = L21 @ Linv
            panels[p0 + ntri:p0 + ntri + b * s, :n_modes] = W21.reshape(n_modes, b * s).T
            updates[i] = F[:, s:, s:] - L21 @ L21.mT
        if transposed:
            stacked = Linv if W21 is None else torch.cat([Linv, W21], dim=1)          # (modes, s+b, s)
            if (s, b) not in triu_cache:
                triu_cache[(s, b)] = torch.ones((s, s + b), dtype=torch.bool, device=device).triu()
            mask = triu_cache[(s, b)]                                                   # (col j, row i) with i >= j
            panels_t[p0:p0 + ntri + b * s, :n_modes] = stacked.mT[:, mask].T
            if m_pad > n_modes:
                jj = torch.arange(s, device=device)
                panels_t[p0 + jj * (s + b) - jj * (jj - 1) // 2, n_modes:] = 1.0
    return (panels, panels_t) if transposed else panels


# ----------------------------------------------------------------------------- hybrid device factorisation (row f1)
def front_maps(sym: Symbolic, Kp: sp.csr_matrix):
    """Index maps of the assembly, fully vectorised:

    a_pos[q]      front position (in the front of the node that owns row(q)) of CSR entry q of the permuted matrix,
                  -1 when the column lies below the owner's first vertex (already eliminated);
    parent_pos[p] front row, in the PARENT's front, of boundary row p of every node (concatenated like ``upd_off``)."""
    n = sym.n
    node_of = np.repeat(np.arange(sym.n_nodes), sym.s)                       # owner node of every vertex (new ids)
    nfront = (sym.s + sym.b).astype(np.int64)
    front_node = np.repeat(np.arange(sym.n_nodes), nfront)
    keys = front_node * np.int64(n) + sym.front_idx                           # globally sorted: nodes ascending, rows ascending
    rows = np.repeat(np.arange(n), np.diff(Kp.indptr))
    owner = node_of[rows]
    cols = Kp.indices.astype(np.int64)
    hit = np.searchsorted(keys, owner * np.int64(n) + cols)
    hit = np.minimum(hit, keys.size - 1)
    ok = (cols >= sym.off[owner]) & (keys[hit] == owner * np.int64(n) + cols)
    a_pos = np.where(ok, hit - sym.front_off[owner], -1).astype(np.int32)
    # boundary rows of every node inside the parent's front
    bnd_node = np.repeat(np.arange(sym.n_nodes), sym.b)
    within = np.arange(bnd_node.size) - np.repeat(sym.upd_off[:-1], sym.b)
    bnd_vertex = sym.front_idx[sym.front_off[bnd_node] + sym.s[bnd_node] + within]
    par = sym.parent[bnd_node]
    hit = np.searchsorted(keys, par * np.int64(n) + bnd_vertex)
    hit = np.minimum(hit, keys.size - 1)
    assert np.all((par < 0) | (keys[hit] == par * np.int64(n) + bnd_vertex)), "child boundary must embed in the parent front"
    parent_pos = np.where(par >= 0, hit - sym.front_off[np.maximum(par, 0)], -1).astype(np.int32)
    return a_pos, parent_pos


def front_maps_native(lib, sym: Symbolic, Kp: sp.csr_matrix):
    """``front_maps`` by the C++ helper ``dots_front_maps`` (csrc/host_order.cpp; linear-time merges instead of two global
    binary searches).  ``Kp`` must have sorted column indices."""
    from . import capi
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    s_, b_, off, f_off, f_idx, u_off, par = (i64(a) for a in (sym.s, sym.b, sym.off, sym.front_off, sym.front_idx, sym.upd_off, sym.parent))
    a_ptr, a_idx = i64(Kp.indptr), i64(Kp.indices)
    a_pos = np.empty(a_idx.size, dtype=np.int32)
    parent_pos = np.empty(int(u_off[-1]), dtype=np.int32)
    capi.check(lib.dots_front_maps(sym.n, sym.n_nodes, *(a.ctypes.data for a in (s_, b_, off, f_off, f_idx, u_off, par, a_ptr, a_idx,
                                                                                   a_pos, parent_pos))), "dots_front_maps")
    return a_pos, parent_pos


def permuted_matrix(lib, sym: Symbolic, K: sp.csr_matrix) -> sp.csr_matrix:
    """``K[perm][:, perm]`` with sorted column indices: by ``dots_csr_permute`` (csrc/host_order.cpp) when the library is
    there, by scipy otherwise (the checker, tests/test_nested_host.py)."""
    if lib is None:
        Kp = K[sym.perm][:, sym.perm].tocsr()
        Kp.sort_indices()
        return Kp
    from . import capi
    K = K.tocsr()
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    a_ptr, a_idx, a_val = i64(K.indptr), i64(K.indices), np.ascontiguousarray(K.data, dtype=np.float64)
    perm, iperm = i64(sym.perm), i64(sym.iperm)
    o_ptr, o_idx, o_val = np.empty(sym.n + 1, np.int64), np.empty(a_idx.size, np.int64), np.empty(a_idx.size)
    capi.check(lib.dots_csr_permute(sym.n, *(a.ctypes.data for a in (a_ptr, a_idx, a_val, perm, iperm, o_ptr, o_idx, o_val))),
               "dots_csr_permute")
    Kp = sp.csr_matrix((o_val, o_idx, o_ptr), shape=K.shape)
    Kp.has_sorted_indices = True
    return Kp


def factor_hybrid_device(sym: Symbolic, K: sp.csr_matrix, mass: np.ndarray, shifts: np.ndarray, m_pad: int, device, lib,
                         stream_fn, stats: dict | None = None, front_nmax: int | None = None, use_library: bool = False):
    """Numeric factorisation on the GPU, level by level, with hand-written kernels only: small fronts by ``k_front_small``
    (one block per node and mode, front in shared memory), the large fronts near the root by the blocked kernels of
    csrc/front_large.cu (``dots_factor_large_fronts``, one call per level).  ``use_library=True`` (DOTS_FACTOR=mixed) routes
    the large fronts through batched torch.linalg calls instead (the path the kernels replaced; kept as a cross-check).
    Returns (panels, panels_t), both (panel_entries, m_pad) on ``device``."""
    import ctypes as C
    import torch
    from . import capi

    n_modes = int(len(shifts))
    nmax = int(lib.dots_front_nmax()) if front_nmax is None else int(front_nmax)   # 0: every front through the library path
    Kp = permuted_matrix(lib, sym, K)
    a_pos, parent_pos = front_maps_native(lib, sym, Kp) if lib is not None else front_maps(sym, Kp)
    dev = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dtype=dt), device=device)
    shifts_t = dev(shifts, np.float64)
    massp = dev(np.asarray(mass, dtype=np.float64)[sym.perm], np.float64)
    t = dict(nd_off=dev(sym.off, np.int32), nd_s=dev(sym.s, np.int32), nd_b=dev(sym.b, np.int32),
             nd_child=dev(sym.child, np.int32), nd_panel=dev(sym.panel_off[:-1], np.int64),
             nd_front=dev(sym.front_off[:-1], np.int64), nd_upd=dev(sym.upd_off[:-1], np.int64),
             parent_pos=dev(parent_pos, np.int32), a_ptr=dev(Kp.indptr, np.int64), a_pos=dev(a_pos, np.int32),
             a_val=dev(Kp.data, np.float64))
    panels = torch.zeros((sym.panel_entries, m_pad), dtype=torch.float64, device=device)
    panels_t = torch.zeros((sym.panel_entries, m_pad), dtype=torch.float64, device=device)
    u_ptr_host = np.zeros(sym.n_nodes, dtype=np.int64)
    u_ptr_dev = torch.zeros(sym.n_nodes, dtype=torch.int64, device=device)
    u_view = {}                                           # node -> (b, b, m_pad) view of its update matrix
    level_buf = {}
    levels = level_schedule(sym)
    # One pinned snapshot of the pointer table per level: a pageable copy of this size (> 64 KB) blocks the host until the
    # stream has drained, which serialised the enqueue of a level with the factorisation of the one before (measured with
    # tools/setup_levels.py: 0.19 s of enqueue for 0.185 s of kernels, so nothing of the caller's host planning overlapped).
    u_ptr_stage = (torch.zeros((len(levels), sym.n_nodes), dtype=torch.int64).pin_memory()
                   if torch.device(device).type == "cuda" else None)
    last_use = {}
    for lv, nodes in enumerate(levels):
        par = sym.parent[nodes]
        last_use[lv] = int(sym.level[par[par >= 0]].max()) if (par >= 0).any() else lv
    pin_node = sym.n_nodes - 1
    pin_value = float(Kp.diagonal().mean())
    singular = [m for m in range(n_modes) if float(shifts[m]) == 0.0]
    tril_cache, triu_cache = {}, {}
    dev_i64 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int64), device=device)
    n_small = n_large = 0
    # Index data of the library path, uploaded ONCE (a host->device copy per large front used to cost more than its algebra):
    # the CSR entries whose column lies inside their owner's front ("kept"), as (row inside the owner's S block, front
    # position, value), in CSR order; kept_ptr[v] = number of kept entries in the rows before vertex v.
    if use_library:
        row_of = np.repeat(np.arange(sym.n), np.diff(Kp.indptr))
        kept = a_pos >= 0
        kept_ptr = np.concatenate([[0], np.cumsum(np.bincount(row_of[kept], minlength=sym.n))]).astype(np.int64)
        node_of = np.repeat(np.arange(sym.n_nodes), sym.s)
        kept_r = dev_i64((row_of - sym.off[node_of[row_of]])[kept])
        kept_c = dev_i64(a_pos[kept])
        kept_v = torch.as_tensor(np.ascontiguousarray(Kp.data[kept], dtype=np.float64), device=device)
        parent_pos64 = dev_i64(parent_pos)

    args = capi.FrontArgs()
    args.n_modes, args.m_pad = n_modes, m_pad
    for k in ("nd_off", "nd_s", "nd_b", "nd_child", "nd_panel", "nd_front", "nd_upd", "parent_pos", "a_ptr", "a_pos", "a_val"):
        setattr(args, k, t[k].data_ptr())
    args.mass, args.shifts = massp.data_ptr(), shifts_t.data_ptr()
    args.panels, args.panels_t = panels.data_ptr(), panels_t.data_ptr()
    args.u_ptr = u_ptr_dev.data_ptr()
    args.pin_node, args.pin_value = pin_node, pin_value

    def _library_large_fronts(large):
        for i in large:                                                    # batched dense algebra over the modes
            i = int(i)
            s, b = int(sym.s[i]), int(sym.b[i])
            nf, lo, f0 = s + b, int(sym.off[i]), int(sym.front_off[i])
            F = torch.zeros((n_modes, nf, nf), dtype=torch.float64, device=device)
            if s:
                k0, k1 = int(kept_ptr[lo]), int(kept_ptr[lo + s])        # this node's kept CSR entries, in CSR order
                rt, ct, vt = kept_r[k0:k1], kept_c[k0:k1], kept_v[k0:k1]
                F[:, rt, ct] = vt
                F[:, ct, rt] = vt
                d = torch.arange(s, device=device)
                F[:, d, d] += shifts_t[:, None] * massp[None, lo:lo + s]
                if i == pin_node:
                    for m in singular:
                        F[m, s - 1, s - 1] += pin_value
            for slot in range(2):
                k = int(sym.child[i, slot])
                if k >= 0 and sym.b[k]:
                    pp = parent_pos64[int(sym.upd_off[k]):int(sym.upd_off[k + 1])]
                    F[:, pp[:, None], pp[None, :]] += u_view[k][:, :, :n_modes].permute(2, 0, 1)
            if s == 0:
                if b:
                    u_view[i][:, :, :n_modes] = F.permute(1, 2, 0)
                continue
            L11 = torch.linalg.cholesky(F[:, :s, :s])
            eye = torch.eye(s, dtype=torch.float64, device=device).expand(n_modes, s, s)
            Linv = torch.linalg.solve_triangular(L11, eye, upper=False)
            if s not in tril_cache:
                tril_cache[s] = torch.tril_indices(s, s, device=device)
            tri = tril_cache[s]
            p0, ntri = int(sym.panel_off[i]), s * (s + 1) // 2
            panels[p0:p0 + ntri, :n_modes] = Linv[:, tri[0], tri[1]].T
            W21 = None
            if b:
                L21 = F[:, s:, :s] @ Linv.mT
                W21 = L21 @ Linv
                panels[p0 + ntri:p0 + ntri + b * s, :n_modes] = W21.reshape(n_modes, b * s).T
                u_view[i][:, :, :n_modes] = (F[:, s:, s:] - L21 @ L21.mT).permute(1, 2, 0)
            stacked = Linv if W21 is None else torch.cat([Linv, W21], dim=1)
            if (s, b) not in triu_cache:
                triu_cache[(s, b)] = torch.ones((s, s + b), dtype=torch.bool, device=device).triu()
            panels_t[p0:p0 + ntri + b * s, :n_modes] = stacked.mT[:, triu_cache[(s, b)]].T
            if m_pad > n_modes:
                jj = torch.arange(s, device=device)
                panels[p0 + jj * (jj + 1) // 2 + jj, n_modes:] = 1.0
                panels_t[p0 + jj * (s + b) - jj * (jj - 1) // 2, n_modes:] = 1.0

    for lv, nodes in enumerate(levels):
        b_l = sym.b[nodes].astype(np.int64)
        offs = np.concatenate([[0], np.cumsum(b_l * b_l)])
        buf = torch.zeros((max(1, int(offs[-1])), m_pad), dtype=torch.float64, device=device)
        level_buf[lv] = buf
        has_b = b_l > 0
        u_ptr_host[nodes[has_b]] = buf.data_ptr() + offs[:-1][has_b] * (m_pad * 8)
        if use_library:                                                    # per-node tensor views: only the cross-check path reads them
            for nd, o0, bb in zip(nodes, offs[:-1], b_l):
                if bb:
                    u_view[int(nd)] = buf[int(o0):int(o0 + bb * bb)].view(int(bb), int(bb), m_pad)
        if u_ptr_stage is not None:
            u_ptr_stage[lv].copy_(torch.from_numpy(u_ptr_host))
            u_ptr_dev.copy_(u_ptr_stage[lv], non_blocking=True)
        else:
            u_ptr_dev.copy_(torch.from_numpy(u_ptr_host))
        nfront = sym.s[nodes] + sym.b[nodes]
        small = nodes[nfront <= nmax]
        large = nodes[nfront > nmax]
        if small.size:
            nodes_dev = dev(small, np.int32)
            args.nodes = nodes_dev.data_ptr()
            capi.check(lib.dots_factor_small_fronts(C.byref(args), int(small.size), int(nfront[nfront <= nmax].max()), stream_fn()),
                       "dots_factor_small_fronts")
            n_small += int(small.size)
        if large.size and use_library:
            _library_large_fronts(large)
            n_large += int(large.size)
        elif large.size:                                                   # hand-written blocked factorisation (csrc/front_large.cu)
            cap = max(1, 65535 // n_modes)
            for c0 in range(0, large.size, cap):
                part = large[c0:c0 + cap]
                n_l = (sym.s[part] + sym.b[part]).astype(np.int64)
                s_l = sym.s[part].astype(np.int64)
                foff = np.concatenate([[0], np.cumsum(n_l * n_l * n_modes)])
                goff = np.concatenate([[0], np.cumsum(n_l * s_l * n_modes)])
                Fw = torch.empty(int(foff[-1]), dtype=torch.float64, device=device)
                Gw = torch.empty(max(1, int(goff[-1])), dtype=torch.float64, device=device)
                nodes_dev, foff_d, goff_d = dev(part, np.int32), dev_i64(foff[:-1]), dev_i64(goff[:-1])
                args.nodes = nodes_dev.data_ptr()
                capi.check(lib.dots_factor_large_fronts(C.byref(args), int(part.size), int(n_l.max()), int(s_l.max()),
                                                        int(sym.b[part].max()), foff_d.data_ptr(), goff_d.data_ptr(),
                                                        Fw.data_ptr(), Gw.data_ptr(), stream_fn()), "dots_factor_large_fronts")
                n_large += int(part.size)
                del Fw, Gw
        for old in [k for k, lu in last_use.items() if lu <= lv and k in level_buf and k < lv]:
            if use_library:
                for nd in levels[old]:
                    u_view.pop(int(nd), None)
            del level_buf[old]
    if stats is not None:
        stats.update(small_fronts=n_small, large_fronts=n_large, front_nmax=nmax)
    return panels, panels_t
