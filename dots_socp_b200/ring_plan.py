"""Host-side launch plan of the ring-streamed sweeps (csrc/sweep_ring.cu, ``sweep_mode`` 4; row a8 of SURVEY.md section 8,
reference utils/laplacian_inverse_socp.py:58-59).

Vocabulary: an *output* is a panel row in the forward sweep (columns ``[0, min(o+1, s))`` of the row-major panel) and a
panel column in the backward sweep (rows ``[o, s+b)`` of the column-major copy).  The entries of consecutive outputs of one
node are contiguous in memory in both directions.

* contiguous tasks (``wpr`` 1): one warp streams the outputs ``[oa, oa+n_out)`` of a node = ``n_ent`` consecutive panel
  entries starting at ``pbase``; tasks are cut so that a level has enough of them to fill the machine and none is much
  longer than ``target`` entries;
* split items (``wpr`` 2/4/8): near the root the runs of single outputs are hundreds of KB; there a block takes a group of
  outputs and ``wpr`` warps share each of them;
* pull lists: every boundary row of every node contributes to exactly one vertex of an ancestor; ``gidx[gptr[v]:gptr[v+1]]``
  are the rows of ``upd`` (producer order) that land on vertex ``v``, producers in post-order, and ``gverts`` lists, level by
  level, the vertices that receive anything;
* ``bidx``: row of ``Z = [hat | ywork]`` the backward sweep reads for every front row.
"""
from __future__ import annotations

import numpy as np

from . import nested

TASK_DTYPE = np.dtype([("pbase", "<i8"), ("n_ent", "<i4"), ("oa", "<i4"), ("n_out", "<i4"), ("s", "<i4"), ("b", "<i4"),
                       ("off", "<i4"), ("ubase", "<i4"), ("fbase", "<i4"), ("pad", "<i4", (2,))])
assert TASK_DTYPE.itemsize == 48


def _out_lengths(s, b, forward):
    """Concatenated output lengths of the given nodes and the index of each node's first output."""
    n_out = (s + b) if forward else s
    first = np.concatenate([[0], np.cumsum(n_out)]).astype(np.int64)
    owner = np.repeat(np.arange(s.size), n_out)
    o = np.arange(first[-1]) - first[owner]
    lens = np.minimum(o + 1, s[owner]) if forward else (s[owner] + b[owner] - o)
    return lens.astype(np.int64), owner, o, first


def _tasks(sym, nodes, forward, target):
    """Contiguous warp tasks of one level and direction (structured array, node order = memory order)."""
    nodes = nodes[sym.s[nodes] > 0]
    if nodes.size == 0:
        return np.zeros(0, dtype=TASK_DTYPE)
    s, b = sym.s[nodes].astype(np.int64), sym.b[nodes].astype(np.int64)
    lens, owner, o, first = _out_lengths(s, b, forward)
    cum = np.concatenate([[0], np.cumsum(lens)])                      # entries before every output (level-wide)
    node_cum0 = cum[first[:-1]]
    tot = cum[first[1:]] - node_cum0                                  # panel entries per node
    n_t = np.clip(np.rint(tot / float(target)).astype(np.int64), 1, first[1:] - first[:-1])
    t_owner = np.repeat(np.arange(nodes.size), n_t)
    k = np.arange(t_owner.size) - np.repeat(np.cumsum(n_t) - n_t, n_t)
    want = node_cum0[t_owner] + (k * tot[t_owner]) // n_t[t_owner]    # entry where task k should start
    start = np.searchsorted(cum, want, side="right") - 1              # output containing that entry (level-wide index)
    start = np.maximum(start, first[t_owner])
    keep = np.ones(start.size, dtype=bool)
    keep[1:] = start[1:] != start[:-1]                                # an output longer than a step: merge the duplicates
    start, t_owner = start[keep], t_owner[keep]
    end = np.concatenate([start[1:], [first[-1]]])
    end = np.minimum(end, first[t_owner + 1])                         # a task never crosses its node
    out = np.zeros(start.size, dtype=TASK_DTYPE)
    nd = nodes[t_owner]
    out["pbase"] = sym.panel_off[nd] + (cum[start] - node_cum0[t_owner])
    out["n_ent"] = cum[end] - cum[start]
    out["oa"] = o[start]
    out["n_out"] = end - start
    out["s"], out["b"], out["off"] = sym.s[nd], sym.b[nd], sym.off[nd]
    out["ubase"], out["fbase"] = sym.upd_off[nd], sym.front_off[nd]
    return out


def _split_items(sym, nodes, forward, group):
    """Split items of one level: groups of ``group`` outputs per block."""
    nodes = nodes[sym.s[nodes] > 0]
    n_out = (sym.s[nodes] + sym.b[nodes]) if forward else sym.s[nodes]
    n_items = -(-n_out // group)
    owner = np.repeat(np.arange(nodes.size), n_items)
    k = np.arange(owner.size) - np.repeat(np.cumsum(n_items) - n_items, n_items)
    out = np.zeros(owner.size, dtype=TASK_DTYPE)
    nd = nodes[owner]
    out["pbase"] = sym.panel_off[nd]
    out["oa"] = k * group
    out["n_out"] = np.minimum(group, n_out[owner] - k * group)
    out["s"], out["b"], out["off"] = sym.s[nd], sym.b[nd], sym.off[nd]
    out["ubase"], out["fbase"] = sym.upd_off[nd], sym.front_off[nd]
    return out


def pull_lists(sym):
    """(gptr, gidx): for every vertex the rows of ``upd`` that land on it (nodes with an empty separator contribute nothing)."""
    b = sym.b.astype(np.int64)
    bnd_node = np.repeat(np.arange(sym.n_nodes), b)
    within = np.arange(bnd_node.size) - np.repeat(sym.upd_off[:-1], b)
    vertex = sym.front_idx[sym.front_off[bnd_node] + sym.s[bnd_node] + within]
    rows = np.arange(bnd_node.size, dtype=np.int64)
    live = sym.s[bnd_node] > 0
    vertex, rows = vertex[live], rows[live]
    order = np.argsort(vertex, kind="stable")                          # producers stay in post-order inside a vertex
    gidx = rows[order].astype(np.int32)
    gptr = np.concatenate([[0], np.cumsum(np.bincount(vertex, minlength=sym.n))]).astype(np.int32)
    return gptr, gidx


def build(sym, n_sm: int, m_pad: int, split_bytes: int = 96 * 1024, tasks_per_sm: int = 64,
          task_bytes=(16 * 1024, 96 * 1024), wpr_max: int = 8):
    """Plan of both sweeps.  Returns a dict of contiguous numpy arrays (see the module docstring)."""
    ent_bytes = 8 * m_pad
    lo_ent, hi_ent = max(1, task_bytes[0] // ent_bytes), max(1, task_bytes[1] // ent_bytes)
    fwd, bwd, fptr, bptr, fw, bw = [], [], [0], [0], [], []
    gverts, gv_ptr = [], [0]
    gptr, gidx = pull_lists(sym)
    has = np.diff(gptr) > 0
    for nodes in nested.level_schedule(sym):
        live = nodes[sym.s[nodes] > 0]
        s_l, b_l = sym.s[live].astype(np.int64), sym.b[live].astype(np.int64)
        for forward, acc, ptr, wl in ((True, fwd, fptr, fw), (False, bwd, bptr, bw)):
            if live.size == 0:
                acc.append(np.zeros(0, dtype=TASK_DTYPE)); ptr.append(ptr[-1]); wl.append(1)
                continue
            work = s_l * (s_l + 1) // 2 + s_l * b_l
            n_out = int((s_l + b_l).sum() if forward else s_l.sum())
            # work-weighted mean run length of an output
            length = s_l.astype(float) if forward else (s_l / 2.0 + b_l)
            len_eff = float((length * work).sum() / work.sum())
            run_bytes = len_eff * ent_bytes
            wpr = 1
            while wpr < wpr_max and run_bytes / wpr > split_bytes:
                wpr *= 2
            if wpr == 1:
                target = int(np.clip(work.sum() // (n_sm * tasks_per_sm), lo_ent, hi_ent))
                items = _tasks(sym, live, forward, target)
            else:
                rows = 8 // wpr
                passes = int(min(8, max(1, n_out // (rows * 4 * n_sm))))
                items = _split_items(sym, live, forward, rows * passes)
            acc.append(items); ptr.append(ptr[-1] + items.size); wl.append(wpr)
        mine = np.repeat(sym.off[live], s_l) + (np.arange(int(s_l.sum())) - np.repeat(np.cumsum(s_l) - s_l, s_l))
        mine = mine[has[mine]] if mine.size else mine.astype(np.int64)
        gverts.append(mine.astype(np.int32)); gv_ptr.append(gv_ptr[-1] + mine.size)
    nfront = (sym.s + sym.b).astype(np.int64)
    pos = np.arange(int(sym.front_off[-1])) - np.repeat(sym.front_off[:-1], nfront)
    bidx = (sym.front_idx + sym.n * (pos < np.repeat(sym.s, nfront))).astype(np.int32)
    cat = lambda parts: np.ascontiguousarray(np.concatenate(parts)) if sum(p.size for p in parts) else np.zeros(1, dtype=parts[0].dtype if parts else np.int32)
    i32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    return dict(rt_fwd=cat(fwd), rt_bwd=cat(bwd), fwd_ptr=i32(fptr), bwd_ptr=i32(bptr), fwd_wpr=i32(fw), bwd_wpr=i32(bw),
                bidx=np.ascontiguousarray(bidx), gptr=np.ascontiguousarray(gptr),
                gidx=gidx if gidx.size else np.zeros(1, np.int32), gverts=cat(gverts), gv_ptr=i32(gv_ptr))


def entry_rows(lib, sym, bidx):
    """(erow_fwd, erow_bwd): for every panel entry in streaming order the row of Z it multiplies (bit 31: last entry of its
    output), built by the C++ helper ``dots_ring_entry_rows`` (a Python loop over 10 M entries would dominate the setup)."""
    from . import capi
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    s, b, off, fo, po = i64(sym.s), i64(sym.b), i64(sym.off), i64(sym.front_off), i64(sym.panel_off)
    bi = np.ascontiguousarray(bidx, dtype=np.int32)
    ef, eb = np.zeros(max(1, sym.panel_entries), dtype=np.int32), np.zeros(max(1, sym.panel_entries), dtype=np.int32)
    capi.check(lib.dots_ring_entry_rows(sym.n_nodes, s.ctypes.data, b.ctypes.data, off.ctypes.data, fo.ctypes.data, po.ctypes.data,
                                        bi.ctypes.data, ef.ctypes.data, eb.ctypes.data), "dots_ring_entry_rows")
    return ef, eb


def entry_rows_numpy(sym, bidx):
    """Plain statement of ``entry_rows`` (tests)."""
    last = np.int32(-2 ** 31)
    ef, eb = np.zeros(max(1, sym.panel_entries), dtype=np.int32), np.zeros(max(1, sym.panel_entries), dtype=np.int32)
    for nd in range(sym.n_nodes):
        s, b, off, p = int(sym.s[nd]), int(sym.b[nd]), int(sym.off[nd]), int(sym.panel_off[nd])
        if s == 0:
            continue
        bi = bidx[sym.front_off[nd]:sym.front_off[nd] + s + b]
        k = p
        for i in range(s + b):
            n = min(i + 1, s)
            ef[k:k + n] = off + np.arange(n)
            ef[k + n - 1] |= last
            k += n
        k = p
        for j in range(s):
            n = s + b - j
            eb[k:k + n] = bi[j:]
            eb[k + n - 1] |= last
            k += n
    return ef, eb


def launches(plan) -> int:
    """Kernel launches of one pair of sweeps with this plan."""
    return int(np.count_nonzero(np.diff(plan["fwd_ptr"])) + np.count_nonzero(np.diff(plan["bwd_ptr"]))
               + np.count_nonzero(np.diff(plan["gv_ptr"])))
