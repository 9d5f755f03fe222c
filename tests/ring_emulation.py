"""Numpy statement of what the kernels of dots_socp_b200/csrc/sweep_ring.cu compute from a ring plan, entry by entry.

Test infrastructure (host tests of the PLAN: exact cover of the panels, pull lists, index maps, flush targets); the
kernels themselves are checked on the GPU in tests/test_gpu_parity.py."""
import numpy as np


def transpose_panels(sym, panels):
    """Column-major copy of the solve-ready panels (per node: column j holds rows j..s+b-1), as nested.py builds it."""
    out = np.zeros_like(panels)
    for i in range(sym.n_nodes):
        s, b = int(sym.s[i]), int(sym.b[i])
        if s == 0:
            continue
        p0 = int(sym.panel_off[i])
        full = np.zeros((s + b, s, panels.shape[1]))
        tri = np.tril_indices(s)
        ntri = s * (s + 1) // 2
        full[tri[0], tri[1]] = panels[p0:p0 + ntri]
        if b:
            full[s:] = panels[p0 + ntri:p0 + ntri + b * s].reshape(b, s, panels.shape[1])
        k = p0
        for j in range(s):
            out[k:k + s + b - j] = full[j:, j]
            k += s + b - j
    return out


def emulate(sym, plan, panels, panels_t, rhs):
    """What the kernels of csrc/sweep_ring.cu compute, statement by statement, in numpy (host tests of the PLAN: exact
    cover, pull lists, index maps).  ``panels`` / ``panels_t``: (panel_entries, M); ``rhs``: (V, M).  Returns x (V, M)."""
    V, M = rhs.shape
    z = np.concatenate([rhs.astype(np.float64), np.zeros((V, M))])         # Z = [hat | ywork]
    upd = np.zeros((max(1, int(sym.upd_off[-1])), M))

    def flush(t, o, acc, forward):
        if not forward:
            z[t["off"] + o] = -acc
        elif o < t["s"]:
            z[V + t["off"] + o] = acc
        else:
            upd[t["ubase"] + o - t["s"]] = -acc

    def run(t, forward):                                                   # k_ring_run: the per-entry codes drive everything
        src = panels if forward else panels_t
        codes = plan["erow_fwd"] if forward else plan["erow_bwd"]
        o, acc = int(t["oa"]), np.zeros(M)
        for e in range(int(t["n_ent"])):
            code = int(codes[t["pbase"] + e])
            acc = acc + src[t["pbase"] + e] * z[code & 0x7fffffff]
            if code < 0:
                flush(t, o, acc, forward)
                acc = np.zeros(M)
                o += 1
        assert o == t["oa"] + t["n_out"], "task does not end on an output boundary"

    def split(t, forward, wpr):                                            # k_ring_split
        src = panels if forward else panels_t
        codes = plan["erow_fwd"] if forward else plan["erow_bwd"]
        s, b = int(t["s"]), int(t["b"])
        for o in range(int(t["oa"]), int(t["oa"] + t["n_out"])):
            lo, hi = (0, min(o + 1, s)) if forward else (o, s + b)
            base = (o * (o + 1) // 2 if o < s else s * (s + 1) // 2 + (o - s) * s) if forward else (o * (s + b) - o * (o - 1) // 2)
            plen = -(-(hi - lo) // wpr)
            total = np.zeros(M)
            for w in range(wpr):
                xa, xb = lo + w * plen, min(hi, lo + (w + 1) * plen)
                part = np.zeros(M)
                for x in range(xa, xb):
                    ent = t["pbase"] + base + (x - lo)
                    part = part + src[ent] * z[int(codes[ent]) & 0x7fffffff]
                total = total + part
            flush(t, o, total, forward)

    n_levels = len(plan["fwd_wpr"])
    for lv in range(n_levels):
        for v in plan["gverts"][plan["gv_ptr"][lv]:plan["gv_ptr"][lv + 1]]:   # k_ring_gather
            for g in range(plan["gptr"][v], plan["gptr"][v + 1]):
                z[v] = z[v] + upd[plan["gidx"][g]]
        for t in plan["rt_fwd"][plan["fwd_ptr"][lv]:plan["fwd_ptr"][lv + 1]]:
            run(t, True) if plan["fwd_wpr"][lv] == 1 else split(t, True, int(plan["fwd_wpr"][lv]))
    for lv in range(n_levels - 1, -1, -1):
        for t in plan["rt_bwd"][plan["bwd_ptr"][lv]:plan["bwd_ptr"][lv + 1]]:
            run(t, False) if plan["bwd_wpr"][lv] == 1 else split(t, False, int(plan["bwd_wpr"][lv]))
    return z[:V]
