"""Host-side invariants of the per-level launch plan of the batched sweeps (dots_socp_b200/engine.py:_sweep_items).

The kernels trust the plan blindly: every panel row (forward) and every panel column (backward) must be covered by
exactly one work item, a level may only contain nodes of that level, and the fused child gather may only be selected
where the forward block's shared-memory staging fits.  These are checked here without a GPU."""
import numpy as np
import pytest

from dots_socp_b200 import engine, nested, surface, synth


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


@pytest.mark.parametrize("example,leaf,n_sm", [("icosphere3", 16, 148), ("icosphere5", 16, 148), ("knot", 16, 148), ("plane8", 6, 4)])
def test_tile_plan_covers_every_output_once_in_whole_groups(example, leaf, n_sm):
    """Plan of the experimental tile-streamed sweep (sweep_mode=2): exact cover, items start on their node's output 0 + k*per
    and every item but a node's last one is a multiple of 8 outputs long."""
    sym = _sym(example, leaf)
    plan = engine._sweep_items_tile(sym, n_sm)
    for lv, nodes in enumerate(nested.level_schedule(sym)):
        for key_ptr, key_items, lens in (("fwd_ptr", "fwd_items", {int(n): int(sym.s[n] + sym.b[n]) for n in nodes}),
                                         ("bwd_ptr", "bwd_items", {int(n): int(sym.s[n]) for n in nodes})):
            items = plan[key_items][plan[key_ptr][lv]:plan[key_ptr][lv + 1]]
            seen = _coverage(items, lens)
            assert all((c == 1).all() for c in seen.values())
            assert {int(n) for n in items[:, 0]} == {n for n, ln in lens.items() if ln > 0}
            for nd, o0, cnt in items:
                assert cnt <= 32 and (cnt % 8 == 0 or o0 + cnt == lens[int(nd)])
