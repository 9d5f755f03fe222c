"""Host-side control logic of the ALM driver: when the penalty moves, which KKT residuals are
evaluated on which iteration, how often the lazy check fires.

These decisions fix the iteration count, so they restate the reference's behaviour exactly
(including its quirks) while living on host scalars only:

* ``PenaltySchedule``   utils/admm_tools.py:19-114  (cadence :30-52, factor table :64-95, z-rescale trigger :107-114)
* ``LazyResidualCheck`` utils/condition_validator.py:194-331 (circular sweep with early exit) wrapped by
                        utils/condition_validator_wrapper.py:9-134 (adaptive checking interval 1..37)

The residual functions themselves are supplied by the caller (they launch CUDA reductions).
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence


class PenaltySchedule:
    R_MIN, R_MAX = 1e-3, 1e3
    #: (iteration bound, minimum gap) pairs of admm_tools.py:43-48; beyond the last bound the gap is 43
    CADENCE = ((20, 3), (50, 7), (100, 11), (200, 17), (500, 31))
    LATE_GAP = 43
    #: thresholds on max(gap, 1/gap) -> multiplicative step, admm_tools.py:78-89
    STEPS = ((50.0, 2.00), (35.0, 1.75), (20.0, 1.60), (10.0, 1.40), (5.0, 1.35), (3.0, 1.32),
             (2.5, 1.28), (2.0, 1.26), (1.5, 1.20), (1.2, 1.10))

    def __init__(self):
        self.last_update = -1
        self.z_rescale_attempts = 0

    def due(self, it: int) -> bool:
        """True (and remember ``it``) when a penalty update falls on this iteration."""
        waited = it - self.last_update
        hit = waited >= self.LATE_GAP or any(it < bound and waited >= gap for bound, gap in self.CADENCE)
        if hit:
            self.last_update = it
        return hit

    @classmethod
    def step_factor(cls, prim_dual_gap: float) -> float:
        shrink = prim_dual_gap < 1.0
        ratio = 1.0 / prim_dual_gap if shrink else prim_dual_gap
        factor = next((f for thr, f in cls.STEPS if ratio > thr), 1.0)
        return 1.0 / factor if shrink else factor

    def next_penalty(self, r: float, prim_dual_gap: float) -> float:
        return max(min(r * self.step_factor(prim_dual_gap), self.R_MAX), self.R_MIN)

    def z_rescale_due(self, it: int, last_kkt_row, min_it: int = 100, max_times: int = 1, tol: float = 5e-3) -> bool:
        """admm_tools.py:107-114.  ``max`` is Python's builtin on purpose: with NaN entries (conditions not
        evaluated on the last recorded iteration) its result depends on the position of the NaNs, and
        the reference's trigger inherits exactly that."""
        if it >= min_it and self.z_rescale_attempts < max_times and max(last_kkt_row) < tol:
            self.z_rescale_attempts += 1
            return True
        return False


def max_or_none(values) -> Optional[float]:
    vals = [v for v in values if v is not None]
    return max(vals) if vals else None


class LazyResidualCheck:
    """Seven two-valued residual functions behind a circular queue + an adaptive interval.

    ``evaluate(required)`` returns (all_passed, n_evaluated); ``collect()`` hands out and clears the
    values computed since the last call (``None`` for conditions that were not touched)."""

    MIN_INTERVAL, MAX_INTERVAL = 1, 37

    def __init__(self, residuals: Sequence[Callable[[], List[Optional[float]]]], tol: float,
                 queue: Sequence[int] = (6, 2, 0, 3, 1, 4, 5)):
        self.residuals = list(residuals)
        self.size = len(self.residuals)
        self.tol = tol
        self.queue = list(queue)                       # queue slot -> condition id (solver_socp.py:644)
        self.slot_of = {cond: slot for slot, cond in enumerate(self.queue)}
        self.head = 0                                  # persists between sweeps
        self.interval = 1
        self.ticks = 0
        self._fresh = [[None, None] for _ in range(self.size)]
        self.evaluations = [0] * self.size

    # -- one residual ------------------------------------------------------------------------------
    def _run(self, cond: int) -> bool:
        values = list(self.residuals[cond]())
        self._fresh[cond] = values
        self.evaluations[cond] += 1
        return values[0] < self.tol

    # -- condition_validator.py:236-331 ---------------------------------------------------------------
    def _sweep(self, required: Optional[Sequence[int]]):
        done_slots, count = set(), 0
        required_ok = True
        for cond in (required or ()):
            slot = self.slot_of[cond]
            if slot in done_slots:
                continue
            done_slots.add(slot)
            count += 1
            if not self._run(cond):                    # required conditions are ALL evaluated, no early exit
                required_ok = False
        if not required_ok:
            return False, count
        if count >= self.size:
            return True, count
        first = self.head
        while count < self.size:
            slot = self.head % self.size
            if slot not in done_slots:
                done_slots.add(slot)
                count += 1
                if not self._run(self.queue[slot]):
                    return False, count                # the head stays on the failing condition
            self.head = (self.head + 1) % self.size
            if self.head == first:
                return True, count
        return False, count                            # budget exhausted before the head wrapped (reference quirk)

    # -- condition_validator_wrapper.py:99-125 ----------------------------------------------------------
    def will_fire(self, required: Optional[Sequence[int]]) -> bool:
        """Pure look-ahead: would ``evaluate(required)`` run a sweep on this tick?"""
        return (self.ticks % self.interval) == 0 or bool(required)

    def evaluate(self, required: Optional[Sequence[int]] = None):
        fire = (self.ticks % self.interval) == 0
        self.ticks += 1
        if fire or required:
            return self._sweep(required)
        return False, 0

    def restart_ticks(self):
        self.ticks = 0

    def collect(self):
        out, self._fresh = self._fresh, [[None, None] for _ in range(self.size)]
        return out

    # -- condition_validator_wrapper.py:44-97 -------------------------------------------------------------
    def adapt(self, error: float):
        ratio = error / max(self.tol, 1e-10)
        if ratio <= 1.0:
            self.interval = self.MIN_INTERVAL
            return
        decades = math.log10(ratio)
        if decades > 1.0:
            self.interval = self.MAX_INTERVAL
        else:
            self.interval = max(self.MIN_INTERVAL,
                                int(self.MIN_INTERVAL + decades * (self.MAX_INTERVAL - self.MIN_INTERVAL)))
