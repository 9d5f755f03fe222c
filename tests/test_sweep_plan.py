"""Host-side invariants of the launch plans of the batched sweeps (dots_socp_b200/engine.py:_sweep_items for k_sweep_run,
dots_socp_b200/ring_plan.py for the ring-streamed kernels of csrc/sweep_ring.cu).

The kernels trust the plan blindly: every panel row (forward) and every panel column (backward) must be covered by
exactly one work item, a level may only contain nodes of that level, and the fused child gather may only be selected
where the forward block's shared-memory staging fits.  These are checked here without a GPU."""
import numpy as np
import pytest

from dots_socp_b200 import engine, nested, ring_plan, surface, synth


def _sym(example, leaf):
    geo, _ = synth.example(example)
    K = surface.stiffness_matrix(geo["vertices"], geo["triangles"])
    return nested.analyse(geo["vertices"], K, leaf_size=leaf)


def _coverage(items, lens):
    seen = {int(nd): np.zeros(int(n), dtype=np.int32) for nd, n in lens.items()}
    for nd, o0, cnt in items:
        assert cnt >= 1
        seen[int(nd)][o0:o0 + cnt] += 1
    return seen


@pytest.mark.parametrize("example,leaf,m_pad,n_sm", [("icosphere3", 16, 32, 148), ("icosphere4", 16, 64, 148),
                                                     ("knot", 16, 8, 148), ("plane8", 6, 16, 4)])
def test_every_output_is_covered_exactly_once(example, leaf, m_pad, n_sm):
    sym = _sym(example, leaf)
    plan = engine._sweep_items(sym, n_sm, m_pad)
    levels = nested.level_schedule(sym)
    assert len(plan["fwd_ptr"]) == len(levels) + 1 == len(plan["bwd_ptr"]) == len(plan["node_ptr"])
    assert len(plan["wpr"]) == len(levels) == len(plan["cw"])
    for lv, nodes in enumerate(levels):
        fwd = plan["fwd_items"][plan["fwd_ptr"][lv]:plan["fwd_ptr"][lv + 1]]
        bwd = plan["bwd_items"][plan["bwd_ptr"][lv]:plan["bwd_ptr"][lv + 1]]
        node_set = {int(n) for n in nodes}
        live_rows = {n for n in node_set if sym.s[n] + sym.b[n] > 0}
        live_cols = {n for n in node_set if sym.s[n] > 0}
        assert {int(n) for n in fwd[:, 0]} == live_rows
        assert {int(n) for n in bwd[:, 0]} == live_cols
        rows = _coverage(fwd, {n: sym.s[n] + sym.b[n] for n in node_set})
        cols = _coverage(bwd, {n: sym.s[n] for n in node_set})
        for n in node_set:
            assert (rows[n] == 1).all() and (cols[n] == 1).all()


@pytest.mark.parametrize("example,leaf", [("icosphere4", 16), ("knot", 16)])
def test_gather_is_either_fused_or_listed(example, leaf):
    sym = _sym(example, leaf)
    plan = engine._sweep_items(sym, 148, 64)
    for lv, nodes in enumerate(nested.level_schedule(sym)):
        code = int(plan["wpr"][lv])
        fused, wpr = code >= 16, code & 15
        assert wpr in (1, 2, 4, 8) and int(plan["cw"][lv]) in (1, 2, 4, 8)
        listed = plan["nodes"][plan["node_ptr"][lv]:plan["node_ptr"][lv + 1]]
        parents = {int(n) for n in nodes if (sym.child[n] >= 0).any() and sym.s[n] > 0}
        if fused:
            assert len(listed) == 0
            assert wpr <= 2 and int(sym.s[nodes].max()) <= 64          # SWEEP_FG_SMAX of lap_kernels.cu
        else:
            assert {int(n) for n in listed[:, 0]} == parents
            cols = _coverage(listed, {n: sym.s[n] for n in parents})
            assert all((c == 1).all() for c in cols.values())
            assert (listed[:, 2] <= 32).all()


def test_children_sit_on_strictly_lower_levels():
    sym = _sym("icosphere3", 8)
    for n in range(sym.n_nodes):
        for k in sym.child[n]:
            if k >= 0:
                assert sym.level[k] < sym.level[n]


# ------------------------------------------------------------------------------------------------ ring plan (sweep_mode 4)
RING_CASES = [("icosphere3", 16, 32, 148, {}), ("icosphere4", 16, 64, 148, {}), ("knot", 16, 128, 148, {}),
              ("plane8", 6, 96, 4, {}),
              # tiny thresholds: every kind of item (many short tasks, split outputs with 2 / 4 / 8 warps) on a small mesh
              ("icosphere3", 8, 64, 2, dict(split_bytes=2048, task_bytes=(512, 2048))),
              ("icosphere2", 8, 32, 2, dict(split_bytes=1024, task_bytes=(256, 1024), wpr_max=4))]


@pytest.mark.parametrize("example,leaf,m_pad,n_sm,kw", RING_CASES)
def test_ring_plan_streams_every_panel_entry_exactly_once(example, leaf, m_pad, n_sm, kw):
    """Contiguous tasks: their entry runs tile each node's panel exactly, start and end on output boundaries and carry the
    node data the kernel reads.  Split items: every output of the node is in exactly one item."""
    sym = _sym(example, leaf)
    plan = ring_plan.build(sym, n_sm, m_pad, **kw)
    levels = nested.level_schedule(sym)
    assert len(plan["fwd_ptr"]) == len(plan["bwd_ptr"]) == len(plan["gv_ptr"]) == len(levels) + 1
    for lv, nodes in enumerate(levels):
        for d, forward in (("fwd", True), ("bwd", False)):
            items = plan["rt_" + d][plan[d + "_ptr"][lv]:plan[d + "_ptr"][lv + 1]]
            wpr = int(plan[d + "_wpr"][lv])
            assert wpr in (1, 2, 4, 8)
            live = [int(n) for n in nodes if sym.s[n] > 0]
            by_node = {n: [] for n in live}
            for t in items:
                nd = int(np.searchsorted(sym.off, t["off"], side="right") - 1)
                while sym.s[nd] == 0:                                    # empty separators share their offset with a neighbour
                    nd -= 1
                assert nd in by_node, "item of a node from another level"
                assert (t["s"], t["b"], t["ubase"], t["fbase"]) == (sym.s[nd], sym.b[nd], sym.upd_off[nd], sym.front_off[nd])
                by_node[nd].append(t)
            for nd, ts in by_node.items():
                s, b = int(sym.s[nd]), int(sym.b[nd])
                n_out = s + b if forward else s
                row_off = lambda o: (o * (o + 1) // 2 if o < s else s * (s + 1) // 2 + (o - s) * s) if forward else (o * (s + b) - o * (o - 1) // 2)
                nxt = 0
                for t in ts:
                    assert t["oa"] == nxt and t["n_out"] >= 1
                    nxt = int(t["oa"] + t["n_out"])
                    if wpr == 1:
                        assert t["pbase"] == sym.panel_off[nd] + row_off(int(t["oa"]))
                        assert t["n_ent"] == row_off(nxt) - row_off(int(t["oa"])) > 0
                    else:
                        assert t["pbase"] == sym.panel_off[nd]
                assert nxt == n_out


@pytest.mark.parametrize("example,leaf", [("icosphere3", 8), ("knot", 16), ("plane8", 6)])
def test_pull_lists_route_every_boundary_row_to_its_vertex_once(example, leaf):
    sym = _sym(example, leaf)
    plan = ring_plan.build(sym, 148, 64)
    gptr, gidx = plan["gptr"], plan["gidx"]
    total = int(sym.upd_off[-1])
    live_rows = np.repeat(sym.s > 0, sym.b)
    assert gptr[-1] == live_rows.sum() and gptr[0] == 0
    assert sorted(gidx[:gptr[-1]].tolist()) == np.nonzero(live_rows)[0].tolist()         # every live row exactly once
    row_node = np.repeat(np.arange(sym.n_nodes), sym.b)
    for v in range(sym.n):
        rows = gidx[gptr[v]:gptr[v + 1]]
        prod = row_node[rows]
        assert (np.diff(prod) > 0).all()                                                # producers in post-order
        for r, nd in zip(rows, prod):
            assert sym.front_idx[sym.front_off[nd] + sym.s[nd] + (r - sym.upd_off[nd])] == v
    # gverts: level by level exactly the vertices with a non-empty list
    node_of = np.repeat(np.arange(sym.n_nodes), sym.s)
    for lv in range(sym.n_levels):
        gv = plan["gverts"][plan["gv_ptr"][lv]:plan["gv_ptr"][lv + 1]]
        want = [v for v in range(sym.n) if sym.level[node_of[v]] == lv and gptr[v + 1] > gptr[v]]
        assert gv.tolist() == want
    assert total == 0 or plan["bidx"].max() < 2 * sym.n


@pytest.mark.parametrize("example,leaf,m_pad,n_sm,kw", RING_CASES[:1] + RING_CASES[4:])
def test_ring_plan_reproduces_the_sparse_solve(example, leaf, m_pad, n_sm, kw):
    """The numpy statement of the ring kernels (tests/ring_emulation.py), run on this plan with panels from the numpy
    multifrontal checker, solves (K + shift M) x = -rhs for every mode."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import host_multifrontal as hm
    from ring_emulation import emulate, transpose_panels
    geo, _ = synth.example(example)
    v, tri = geo["vertices"], geo["triangles"]
    K = surface.stiffness_matrix(v, tri)
    area_f = surface.triangle_areas(v, tri)
    mass = surface.incident_area_sum(v.shape[0], tri, area_f) / 3.0
    sym = nested.analyse(v, K, leaf_size=leaf)
    shifts = np.array([0.7, 31.0, 900.0])
    panels = hm.factor_batched(sym, K, mass, shifts)
    plan = ring_plan.build(sym, n_sm, m_pad, **kw)
    plan["erow_fwd"], plan["erow_bwd"] = ring_plan.entry_rows_numpy(sym, plan["bidx"])
    rng = np.random.default_rng(11)
    rhs = rng.standard_normal((sym.n, shifts.size))
    x = emulate(sym, plan, panels, transpose_panels(sym, panels), rhs)
    Kp = K[sym.perm][:, sym.perm].tocsc()
    for m, sh in enumerate(shifts):
        ref = -spla.spsolve(Kp + sh * sp.diags(mass[sym.perm]), rhs[:, m])
        assert np.abs(x[:, m] - ref).max() <= 1e-10 * np.abs(ref).max()


@pytest.mark.parametrize("example,leaf", [("icosphere3", 8), ("knot", 16), ("plane8", 6)])
def test_entry_rows_native_equals_numpy_statement(example, leaf):
    """dots_ring_entry_rows (C++) == ring_plan.entry_rows_numpy: operand row of every panel entry in both streaming orders,
    with the end-of-output flag on exactly one entry per output."""
    from dots_socp_b200 import capi
    sym = _sym(example, leaf)
    plan = ring_plan.build(sym, 148, 64)
    ef, eb = ring_plan.entry_rows(capi.load(), sym, plan["bidx"])
    rf, rb = ring_plan.entry_rows_numpy(sym, plan["bidx"])
    assert np.array_equal(ef, rf) and np.array_equal(eb, rb)
    live = sym.s > 0
    assert int((ef[:sym.panel_entries] < 0).sum()) == int((sym.s + sym.b)[live].sum())      # one flag per panel row
    assert int((eb[:sym.panel_entries] < 0).sum()) == int(sym.s[live].sum())                # one flag per panel column
    assert (ef & 0x7fffffff).max() < sym.n and (eb & 0x7fffffff).max() < 2 * sym.n
