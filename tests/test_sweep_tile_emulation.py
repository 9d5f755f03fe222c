"""Index logic of the experimental tile-streamed sweep (csrc/sweep_tile.cu, sweep_mode=2), checked on the CPU.

The kernel has not run on hardware yet, so its error-prone part - the (group, chunk) cursor, the
segment offsets of the bulk copies into the row-major panel / the column-major copy, the staged input-vector chunks and the
output writes - is transcribed statement by statement into numpy here and run against a direct sparse solve.  What this
cannot cover is the CUDA-specific part (mbarrier phases, barriers), which follows the validated k_tri_tma / k_sweeps pattern."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from dots_socp_b200 import engine, nested, surface, synth

WARPS, STAGES = 8, 3


def row_off(row, s):                                   # st_row_off
    return row * (row + 1) // 2 if row < s else s * (s + 1) // 2 + (row - s) * s


def col_off(col, s, b):                                # st_col_off
    return col * (s + b) - col * (col - 1) // 2


def span(d, o, s, b):                                  # st_span
    return (0, min(o + 1, s)) if d == 0 else (o, s + b)


def group_range(d, o0, n_o, g, s, b, C):               # st_group_range
    first = o0 + WARPS * g
    if first >= o0 + n_o:
        return None
    lastq = min(first + WARPS, o0 + n_o) - 1
    return (0, min(lastq + 1, s)) if d == 0 else ((first // C) * C, s + b)


def steps(d, o0, n_o, s, b, C):                        # st_first / st_next, statement by statement
    rng = group_range(d, o0, n_o, 0, s, b, C)          # st_first: the first step exists even if its chunk range is empty
    if rng is None:                                     # (a node with an empty separator still has to forward its
        return                                          #  children's updates: one step without data per output group)
    g, (cb, cb_end) = 0, rng
    while True:
        yield g, cb, cb_end
        cb += C                                         # st_next
        if cb < cb_end:
            continue
        g += 1
        rng = group_range(d, o0, n_o, g, s, b, C)
        if rng is None:
            return
        cb, cb_end = rng


def run_item(d, sym, item, panels, panels_t, hat, ywork, upd, M, C):
    node, o0, n_o = (int(x) for x in item)
    s, b, off = int(sym.s[node]), int(sym.b[node]), int(sym.off[node])
    f0 = int(sym.front_off[node])
    ch = [int(k) for k in sym.child[node]]
    u = [upd[int(sym.upd_off[k]):int(sym.upd_off[k + 1])] if (d == 0 and k >= 0) else None for k in ch]
    cp = [sym.child_pos[slot, f0:f0 + s + b] for slot in range(2)]
    fidx = sym.front_idx[f0:f0 + s + b]
    panel = (panels if d == 0 else panels_t)[int(sym.panel_off[node]):]
    myupd = upd[int(sym.upd_off[node]):int(sym.upd_off[node + 1])]
    ring = np.full((STAGES, WARPS, C, M), np.nan)       # what the bulk copies deposit
    walk = list(steps(d, o0, n_o, s, b, C))

    def issue(k, slot):                                 # the producer lambda
        g, cb, _ = walk[k]
        first = o0 + WARPS * g
        ring[slot] = np.nan
        for q in range(WARPS):
            o = first + q
            if o >= o0 + n_o:
                break
            lo, hi = span(d, o, s, b)
            e_lo = max(lo, cb)
            ne = min(hi, cb + C) - e_lo
            if ne <= 0:
                continue
            ent = row_off(o, s) + e_lo if d == 0 else col_off(o, s, b) + (e_lo - o)
            ring[slot, q, :ne] = panel[ent:ent + ne]

    def vec_chunk(cb, cb_end):                          # vec_fetch + vec_store
        v = np.zeros((C, M))
        for jj in range(C):
            j = cb + jj
            if j >= cb_end:
                continue
            if d == 0:
                r = hat[off + j].copy()
                for slot in range(2):
                    if u[slot] is not None and cp[slot][j] >= 0:
                        r += u[slot][cp[slot][j]]
                v[jj] = r
            else:
                v[jj] = -(ywork[off + j] if j < s else hat[fidx[j]])
        return v

    for k in range(min(STAGES, len(walk))):
        issue(k, k)
    issued = min(STAGES, len(walk))
    acc = np.zeros((WARPS, M))
    for step, (g, cb, cb_end) in enumerate(walk):
        slot = step % STAGES
        vec = vec_chunk(cb, cb_end)
        more = step + 1 < len(walk)
        for warp in range(WARPS):
            o = o0 + WARPS * g + warp
            if o >= o0 + n_o:
                continue
            lo, hi = span(d, o, s, b)
            e_lo = max(lo, cb)
            ne = min(hi, cb + C) - e_lo
            for e in range(max(ne, 0)):
                acc[warp] += ring[slot, warp, e] * vec[e_lo - cb + e]
            if not more or walk[step + 1][0] != g:
                if d == 1:
                    hat[off + o] = acc[warp]
                elif o < s:
                    ywork[off + o] = acc[warp]
                else:
                    val = np.zeros(M)
                    for sl in range(2):
                        if u[sl] is not None and cp[sl][o] >= 0:
                            val += u[sl][cp[sl][o]]
                    myupd[o - s] = val - acc[warp]
                acc[warp] = 0.0
        if issued < len(walk):
            issue(issued, slot)
            issued += 1
    assert not np.isnan(acc).any()


@pytest.mark.parametrize("example,leaf,n_sm", [("icosphere2", 8, 148), ("plane8", 6, 2), ("knot_small", 12, 4),
                                               ("knot", 16, 148)])          # the last one has an empty separator (s = 0, b > 0)
def test_tile_sweep_walk_solves_every_mode(example, leaf, n_sm):
    if example == "knot_small":
        v, t = synth.knot_tube(n_u=40, n_v=6)
    else:
        geo, _ = synth.example(example)
        v, t = geo["vertices"], geo["triangles"]
    K = surface.stiffness_matrix(v, t)
    mass = surface.incident_area_sum(v.shape[0], t, surface.triangle_areas(v, t)) / 3.0
    M = 32
    shifts = np.concatenate([[0.0], np.linspace(0.3, 40.0, M - 1)])
    sym = nested.analyse(v, K, leaf_size=leaf)
    if example == "knot":
        assert ((sym.s == 0) & (sym.b > 0)).any()
    p, pt = nested.factor_batched_device(sym, K, mass, shifts, m_pad=M, device="cpu", transposed=True)
    panels, panels_t = p.numpy(), pt.numpy()
    plan = engine._sweep_items_tile(sym, n_sm)
    rng = np.random.default_rng(5)
    rhs = rng.standard_normal((sym.n, M))
    rhs[:, 0] -= (rhs[:, 0]).mean()                    # singular mode: compatible right-hand side
    hat, ywork, upd = rhs.copy(), np.zeros_like(rhs), np.zeros((int(sym.upd_off[-1]), M))
    C = 512 // M
    n_lv = len(plan["fwd_ptr"]) - 1
    for lv in range(n_lv):
        for item in plan["fwd_items"][plan["fwd_ptr"][lv]:plan["fwd_ptr"][lv + 1]]:
            run_item(0, sym, item, panels, panels_t, hat, ywork, upd, M, C)
    for lv in reversed(range(n_lv)):
        for item in plan["bwd_items"][plan["bwd_ptr"][lv]:plan["bwd_ptr"][lv + 1]]:
            run_item(1, sym, item, panels, panels_t, hat, ywork, upd, M, C)
    Kp = K[sym.perm][:, sym.perm].tocsc()
    massp = mass[sym.perm]
    for m in range(1, M):                              # what the kernels leave in `hat` is xt = -(K + shift M)^-1 rhs
        x = spla.spsolve(Kp + shifts[m] * sp.diags(massp), rhs[:, m])
        assert np.abs(hat[:, m] + x).max() <= 1e-9 * np.abs(x).max()
    resid = Kp @ (-hat[:, 0]) - rhs[:, 0]              # mode 0: pinned solution of the singular system
    assert np.abs(resid[:-1]).max() <= 1e-8 * np.abs(rhs[:, 0]).max()
