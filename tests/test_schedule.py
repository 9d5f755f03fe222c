"""CPU-only: the product's host control logic (dots_socp_b200/schedule.py, history.py).

* replayed against the reference's recorded decisions (golden KKT tables: which residual was evaluated on
  which iteration, penalty path, stopping iteration);
* cross-checked against the oracle's independent restatement on random residual streams;
* and, when /root/reference is present (build container only), against the reference's own classes."""
import os

import numpy as np
import pytest

from dots_socp_b200.schedule import LazyResidualCheck, PenaltySchedule, max_or_none
from dots_socp_b200.history import RunHistory
from oracle import alm_oracle as orc


def replay(rows, tol, nit):
    """Drive the product's control flow with residual values taken from a recorded run; returns what it asked for."""
    sched = PenaltySchedule()
    it_box = [0]

    def residual(i):
        v = rows[it_box[0], i]
        assert not np.isnan(v), f"iteration {it_box[0]}: condition {i} requested but the reference did not evaluate it"
        return [v, v] if i < 4 else [v, None]

    lazy = LazyResidualCheck([(lambda i=i: residual(i)) for i in range(7)], tol)
    r, r_hist, use_org, asked = 1.0, [], False, []
    for it in range(nit):
        it_box[0] = it
        adjust = sched.due(it)
        required = [0, 1, 2, 3] if adjust else None
        if adjust:
            lazy.restart_ticks()
        passed, _ = lazy.evaluate(required)
        errs = lazy.collect()
        if adjust:
            lazy.restart_ticks()
        org, sec = [e[0] for e in errs], [e[1] for e in errs]
        asked.append([v is not None for v in org])
        r_hist.append(r)
        err = max_or_none([org[k] for k in (0, 2, 4, 5)])
        if err is not None:
            lazy.adapt(err)
        if passed:
            return it, np.array(asked), np.array(r_hist)
        mx = max_or_none(sec)
        if mx is not None and mx < 5 * tol:
            use_org = True
        if adjust:
            src = org if use_org else sec
            r = sched.next_penalty(r, max_or_none(src[0:2]) / max_or_none(src[2:4]))
    return nit - 1, np.array(asked), np.array(r_hist)


@pytest.mark.parametrize("name", ["ico2_nt7_c0", "ico2_nt15_tol1e-4", "ico3_nt31_c0", "ico3_nt31_c01", "knot_small_nt8_c005",
                                  "knots5class_nt31_c0", "knots5class_nt31_c01"])
def test_replay_of_reference_decisions(golden, name):
    z, geo, n_time, kw = golden(name)
    rows = z["kkt_rows"].copy()
    last = int(z["iterations"])
    # the reference overwrites the last row with the full final check: every entry is valid there
    stop, asked, r_hist = replay(rows, kw["tol"], kw["nit"])
    assert stop == last
    assert np.array_equal(asked[:-1], ~np.isnan(rows[:-1]))
    assert np.allclose(r_hist, z["r_history"], rtol=1e-12)


def test_against_oracle_restatement_on_random_streams():
    rng = np.random.default_rng(0)
    for trial in range(20):
        tol = 1e-3
        vals = 10 ** rng.uniform(-4.5, -1.0, size=(400, 7))
        vals *= np.linspace(3.0, 0.2, 400)[:, None]
        box = [0]
        fa = [(lambda i=i: [vals[box[0], i], vals[box[0], i]]) for i in range(7)]
        a, b = LazyResidualCheck(fa, tol), orc.LazyKKT(fa, tol)
        sched, last = PenaltySchedule(), -1
        for it in range(400):
            box[0] = it
            due_a = sched.due(it)
            due_b = orc.penalty_due(it, last)
            if due_b:
                last = it
            assert due_a == due_b
            req = [0, 1, 2, 3] if due_a else None
            if due_a:
                a.restart_ticks(); b.reset_counter()
            ra, rb = a.evaluate(req), b.validate(req)
            assert ra == rb
            ea, eb = a.collect(), b.pop_errors()
            assert ea == eb
            err = max_or_none([e[0] for e in ea][k] for k in (0, 2, 4, 5))
            if err is not None:
                a.adapt(err); b.retune(err)
            assert a.interval == b.interval and a.head == b.front
    for gap in 10 ** np.linspace(-3, 3, 200):
        assert PenaltySchedule().next_penalty(0.7, gap) == orc.penalty_new_value(0.7, gap)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_against_reference_classes():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import refshim
    cwd = os.getcwd()
    try:
        refshim.load()
        from dot_surface_socp.utils import admm_tools
        from dot_surface_socp.utils.condition_validator import create_convergence_condition_validator
        from dot_surface_socp.utils.condition_validator_wrapper import AdaptiveValidatorWrapper
    finally:
        os.chdir(cwd)
    rng = np.random.default_rng(1)
    tol = 1e-3
    vals = 10 ** rng.uniform(-4.5, -1.5, size=(600, 7)) * np.linspace(3.0, 0.3, 600)[:, None]
    box = [0]
    fns = [(lambda i=i: [vals[box[0], i], vals[box[0], i]]) for i in range(7)]
    mine = LazyResidualCheck(fns, tol)
    val, col = create_convergence_condition_validator([(f"c{i}", fns[i]) for i in range(7)], tolerance=tol, num_values=2)
    val.optimize_queue_order(queue_order=[6, 2, 0, 3, 1, 4, 5])
    val = AdaptiveValidatorWrapper(val)
    ref_sched, my_sched = admm_tools.AdjustAdmmParam(), PenaltySchedule()
    for it in range(600):
        box[0] = it
        due = ref_sched.is_to_adjust(it)
        assert my_sched.due(it) == due
        req = [0, 1, 2, 3] if due else None
        if due:
            val.reset_counter(); mine.restart_ticks()
        p_ref, _ = val.validate(required_conditions=req)
        p_me, _ = mine.evaluate(req)
        assert p_ref == p_me
        e_ref = col.get_errors()
        e_me = mine.collect()
        assert [list(x) for x in e_ref] == e_me
        err = max_or_none([e[0] for e in e_me][k] for k in (0, 2, 4, 5))
        if err is not None:
            val.set_error_and_tolerance(err, tol); mine.adapt(err)
        assert val.current_interval == mine.interval
    for gap in 10 ** np.linspace(-3, 3, 300):
        assert ref_sched.get_updated_value(0.9, gap) == my_sched.next_penalty(0.9, gap)
    row = np.array([np.nan, 1e-3, 1e-4, np.nan, 1e-4, 1e-4, 1e-4])
    for r in (row, row[::-1].copy(), np.full(7, 1e-4)):
        assert admm_tools.AdjustAdmmParam().is_to_scale_matrix(150, r) == PenaltySchedule().z_rescale_due(150, r)


def test_history_report_is_parseable_by_the_replication_regexes(caplog):
    import logging
    import re
    h = RunHistory(10, ["a", "b"], "SOCP", show_progress=False)
    h.start()
    h.add_time("Step 1", 0.5)
    h.record(0, [1e-2, None])
    h.record(3, [1e-3, 2e-3], history={"Transportation cost": 0.2077944, "Objective value": 0.2})
    h.record(3, [1e-4, 2e-4], history={"Transportation cost": 0.2077945, "Objective value": 0.2})
    h.end()
    assert h.kkt_errors.shape == (2, 2) and np.isnan(h.kkt_errors[0, 1]) and h.kkt_iteration[-1] == 3
    with caplog.at_level(logging.INFO):
        h.print_end_history()
        h.print_steps_time()
    text = caplog.text
    assert re.search(r"^Transportation cost:\s*([-+]?\d+\.\d+e[-+]?\d+)", text, re.M)      # replication/log2table.py:101
    assert re.search(r"^Time of steps\s*:\s*(\d+\.?\d*)\s*sec", text, re.M)               # :102
    assert re.search(r"^Total Iteration(?:\s*\(l\.l\.\))?\s*:\s*(\d+) iterations", text, re.M)   # :103
