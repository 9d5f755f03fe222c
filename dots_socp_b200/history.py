"""Run history of the ALM loop: KKT rows, timers, end-of-run report.

Same public surface and the same report text as the reference's ``RunningHistory``
(utils/admm_tools.py:174-562) because the text is an interface: ``replication/log2table.py:98-106``
parses "Transportation cost:", "Time of steps :" and "Total Iteration :" out of it.
``as_reference_history()`` re-homes the data in the reference's own class when that package is
importable, so ``interface.run_dot_surface`` (which does an isinstance check, interface.py:320-324)
prints its usual report.
"""
from __future__ import annotations

import logging
import sys
import time
from contextlib import contextmanager

import numpy as np

LOG_KKT, LOG_SCALING, LOG_INFO = 11, 12, 20        # config/logging_config.toml:3-10
_RULE = 42


def _banner(text: str) -> str:
    return f"---- {text} ".ljust(_RULE, "-")


class RunHistory:
    def __init__(self, max_record_numbers, kkt_labels, name, kkt_short_labels=None, show_progress=True):
        self.kkt_entry_num = len(kkt_labels)
        self.kkt_labels = list(kkt_labels)
        self.kkt_short_labels = list(kkt_short_labels) if kkt_short_labels is not None else list(kkt_labels)
        self.name = name
        self._cap = int(max_record_numbers)
        self._rows = 0
        self.kkt_errors = np.full((self._cap, self.kkt_entry_num), np.inf)
        self.kkt_iteration = np.full(self._cap, np.inf)
        self.kkt_time = np.full(self._cap, np.inf)
        self.running_time = np.inf
        self.last_record_it = -1
        self.steps_time = {}
        self.history = {}
        self._t0 = np.inf
        self._show_progress = show_progress
        self._target_tol = None
        self._last_progress = 0.0

    # ---- clock ----------------------------------------------------------------------------------
    def start(self):
        self._t0 = time.perf_counter()

    def get_running_time(self):
        return time.perf_counter() - self._t0

    def end(self):
        self.running_time = time.perf_counter() - self._t0
        n = self._rows
        self.kkt_errors, self.kkt_iteration, self.kkt_time = self.kkt_errors[:n], self.kkt_iteration[:n], self.kkt_time[:n]
        for key in self.history:
            self.history[key] = self.history[key][:n]
        if self._target_tol is not None and self._show_progress:
            print(_banner("Finish performing"))
            sys.stdout.flush()

    @contextmanager
    def timer(self, tag):
        t = time.perf_counter()
        yield
        self.add_time(tag, time.perf_counter() - t)

    def add_time(self, tag, seconds):
        self.steps_time[tag] = self.steps_time.get(tag, 0.0) + seconds

    # ---- records ----------------------------------------------------------------------------------
    def record(self, current_it=None, kkt_errors=None, history=None):
        if kkt_errors is None or current_it is None:
            raise ValueError("Argument `kkt_errors` or `current_it` must be provided.")
        if current_it < self.last_record_it:
            raise ValueError(f"Current iteration {current_it} is smaller than last recorded iteration {self.last_record_it}.")
        if current_it == self.last_record_it:
            self._rows -= 1                                   # same iteration again: overwrite
        if self._rows >= self._cap:
            raise ValueError(f"No space left to store the running history ({self._rows} >= {self._cap}).")
        self.last_record_it = current_it
        self.kkt_errors[self._rows] = [np.nan if v is None else v for v in kkt_errors]
        self.kkt_iteration[self._rows] = current_it
        self.kkt_time[self._rows] = time.perf_counter() - self._t0
        for key, val in (history or {}).items():
            if key not in self.history:
                self.history[key] = np.full(self._cap, np.inf)
            self.history[key][self._rows] = val
        self._rows += 1

    def get_current_kkt_errors(self):
        if self._rows == 0:
            return np.full(self.kkt_entry_num, np.inf)
        return self.kkt_errors[self._rows - 1]

    # ---- progress (plain log lines instead of a tqdm bar) ---------------------------------------------
    def create_tol_progress(self, target_tol):
        self._target_tol = target_tol
        if self._show_progress:
            print(_banner("Starting to perform ..."))
        logging.log(LOG_KKT, _banner("Iteration Start")[:-1])

    def show_tol_progress(self, current_it, current_err, active_idx=None, converged_idx=None):
        if converged_idx and self._show_progress:
            names = ", ".join(self.kkt_short_labels[i] for i in converged_idx)
            print(f"Conditions converged at iteration {current_it}: {names}")
        idx = self._rows - 1
        row = " ".join(f"{e:6.2e}" for e in self.kkt_errors[idx])
        logging.log(LOG_KKT, f"Iteration: {self.kkt_iteration[idx]:4.0f} - KKT: {row}")

    # ---- reports (text identical to admm_tools.py:505-562) --------------------------------------------
    def print_steps_time(self, tag_tips="Time of each step", tag_step_time="Time of steps",
                         tag_total_time="Total Time", tag_total_iteration="Total Iteration"):
        total_time, total_it = self.running_time, self.kkt_iteration[-1]
        labels, secs = list(self.steps_time.keys()), list(self.steps_time.values())
        sum_steps = sum(secs)
        width = max(len(x) for x in labels + [tag_step_time, tag_total_time, tag_total_iteration])
        per_step = "\n".join(
            f"{lab:<{width}}: {sec:>7.2f} sec ({100.0 * sec / total_time:5.2f}%) "
            f"({100.0 * sec / total_it:<5.2f} sec/100-iterations)" for lab, sec in zip(labels, secs))
        totals = (f"{tag_step_time.ljust(width)}: {sum_steps:>7.2f} sec ({100.0 * sum_steps / total_time:5.2f}%) "
                  f"({100.0 * sum_steps / total_it:<5.2f} sec/100-iterations)\n"
                  f"{tag_total_time.ljust(width)}: {total_time:>7.2f} sec ({100.0:5.2f}%)\n"
                  f"{tag_total_iteration.ljust(width)}: {total_it:>7.0f} iterations")
        logging.log(LOG_INFO, f"{_banner(tag_tips)}\n{per_step}\n{'-' * _RULE}\n{totals}")

    def print_end_history(self):
        width = max(len(x) for x in self.kkt_labels)
        rows = "\n".join(f"{lab:<{width}}: {err:>6.2e}" for err, lab in zip(self.kkt_errors[-1], self.kkt_labels))
        logging.log(LOG_INFO, f"{_banner('The kkt errors at end')}\n{rows}")
        if self.history:
            extra = "\n".join(f"{key}: {val[-1]:.6e}" for key, val in self.history.items())
            logging.log(LOG_INFO, f"{_banner('Other history at end')}\n{extra}")

    # ---- hand-over to the reference's class -------------------------------------------------------------
    def as_reference_history(self):
        """Same data inside ``dot_surface_socp.utils.admm_tools.RunningHistory`` if that package is importable
        (so the caller's isinstance check passes); otherwise ``self``."""
        try:
            from dot_surface_socp.utils.admm_tools import RunningHistory
        except Exception:
            return self
        ref = RunningHistory(max_record_numbers=max(1, len(self.kkt_iteration)), kkt_labels=self.kkt_labels,
                             name=self.name, kkt_short_labels=self.kkt_short_labels)
        for attr in ("kkt_errors", "kkt_iteration", "kkt_time", "running_time", "last_record_it", "steps_time", "history"):
            setattr(ref, attr, getattr(self, attr))
        ref._kkt_num = len(self.kkt_iteration)
        return ref
