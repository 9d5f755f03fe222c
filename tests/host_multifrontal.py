"""Plain numpy multifrontal factorisation used by the host tests as the checker of the GPU setup path.

Same algebra and the same solve-ready panel layout as ``dots_socp_b200.nested.factor_batched_device`` /
``factor_hybrid_device`` (P = [inv(L11); L21 inv(L11)], ``[row][col][mode]``, mode fastest), written with dense numpy
calls per front.  Test infrastructure only: nothing in the package imports it."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def factor_batched(sym, K: sp.csr_matrix, mass: np.ndarray, shifts: np.ndarray, m_pad: int | None = None,
                   pin_singular: bool = True, out: np.ndarray | None = None) -> np.ndarray:
    """Solve-ready panels of ``K + shifts[m] * diag(mass)`` for every mode m.

    Returns ``panels`` of shape (panel_entries, m_pad) float64, C-contiguous (mode fastest).  Modes
    ``>= len(shifts)`` (padding) hold identity-like data (inverse diagonal 1, zeros elsewhere)."""
    shifts = np.asarray(shifts, dtype=np.float64)
    n_modes = shifts.size
    m_pad = m_pad or n_modes
    Kp = K[sym.perm][:, sym.perm].tocsr()
    Kp.sort_indices()
    massp = np.asarray(mass, dtype=np.float64)[sym.perm]
    indptr, indices, data = Kp.indptr, Kp.indices, Kp.data
    panels = out if out is not None else np.zeros((sym.panel_entries, m_pad))
    singular = [m for m in range(n_modes) if shifts[m] == 0.0] if pin_singular else []
    pin_value = float(Kp.diagonal().mean())
    updates = [None] * sym.n_nodes
    for i in range(sym.n_nodes):
        s, b = int(sym.s[i]), int(sym.b[i])
        nf = s + b
        lo = int(sym.off[i])
        rows = sym.front_idx[sym.front_off[i]:sym.front_off[i] + nf]
        F = np.zeros((n_modes, nf, nf))
        # original entries with a row in S (upper part col >= lo); mirrored
        a0, a1 = indptr[lo], indptr[lo + s]
        r = np.repeat(np.arange(s), np.diff(indptr[lo:lo + s + 1]))
        c_new, v = indices[a0:a1], data[a0:a1]
        keep = c_new >= lo
        r, c_new, v = r[keep], c_new[keep], v[keep]
        c = np.searchsorted(rows, c_new)
        F[:, r, c] = v[None, :]
        F[:, c, r] = v[None, :]
        d = np.arange(s)
        F[:, d, d] += shifts[:, None] * massp[None, lo:lo + s]
        if i == sym.n_nodes - 1:
            for m in singular:
                F[m, s - 1, s - 1] += pin_value
        for slot in range(2):
            k = int(sym.child[i, slot])
            if k >= 0 and sym.b[k]:
                cp = sym.child_pos[slot, sym.front_off[i]:sym.front_off[i] + nf]
                where = np.nonzero(cp >= 0)[0]
                where = where[np.argsort(cp[where])]
                F[:, where[:, None], where[None, :]] += updates[k]
                updates[k] = None
        if s == 0:                                              # empty separator: just forward the children's updates
            updates[i] = F if b else None
            continue
        L11 = np.linalg.cholesky(F[:, :s, :s])
        Linv = np.linalg.inv(L11)
        tri = np.tril_indices(s)
        p0 = int(sym.panel_off[i])
        ntri = s * (s + 1) // 2
        panels[p0:p0 + ntri, :n_modes] = Linv[:, tri[0], tri[1]].T
        if m_pad > n_modes:
            panels[p0 + np.arange(s) * (np.arange(s) + 1) // 2 + np.arange(s), n_modes:] = 1.0
        if b:
            L21 = F[:, s:, :s] @ np.swapaxes(Linv, 1, 2)
            W21 = L21 @ Linv
            panels[p0 + ntri:p0 + ntri + b * s, :n_modes] = W21.reshape(n_modes, b * s).T
            updates[i] = F[:, s:, s:] - L21 @ np.swapaxes(L21, 1, 2)
    return panels
