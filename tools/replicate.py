#!/usr/bin/env python
"""Replication sweep and exact-solution study on the B200 path (BASELINE.json configs[4]).

    python tools/replicate.py --suite main        --outdir gpurun_out/replication
    python tools/replicate.py --suite true_error  --outdir gpurun_out/replication

``main``       = the protocol of the reference's ``make main`` (Makefile:50-91): 14 examples x congestion {0, 0.01, 0.05},
                 ``--ntime=31 --nit=10000 --time_limit=5000 --tol=1e-4``, one ``info.log`` per congestion value in the
                 reference's log format, parsed into ``comparison_table.{md,csv}`` with the regexes of
                 replication/log2table.py:98-106.  The bundled meshes are git-LFS stubs, so generated stand-in surfaces of
                 comparable sizes are used (dots_socp_b200/replication.py:STANDINS).
``true_error`` = ``make true_error`` (Makefile:101-108): ``plane``, tol 1e-5, nit 20000, tolerance checkpoints
                 1e-1 ... 1e-5, L1 / L2 / Linf error of the density against the analytic transport.

Needs a B200; there is no CPU fallback."""
from __future__ import annotations

import argparse
import csv
import json
import logging
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dots_socp_b200 as b200                                    # noqa: E402
from dots_socp_b200 import replication as rep                    # noqa: E402
from dots_socp_b200.history import LOG_INFO                      # noqa: E402


def log_to(path):
    """Root logger -> ``path`` (append) at level info with bare messages, like set_logging_level (interface.py:76-103)."""
    root = logging.getLogger()
    for h in list(root.handlers):
        root.removeHandler(h)
    root.setLevel(LOG_INFO)
    fh = logging.FileHandler(path, mode="a")
    fh.setFormatter(logging.Formatter("%(message)s"))
    root.addHandler(fh)
    return fh


def save(outdir, summary):
    with open(os.path.join(outdir, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1)


def run_main(outdir, examples, congestions, flags, summary):
    for c in congestions:
        sub = os.path.join(outdir, "congestion_" + f"{c:.2f}".replace(".", "_"))
        os.makedirs(sub, exist_ok=True)
        info = os.path.join(sub, "info.log")
        if os.path.exists(info):
            os.remove(info)
        fh = log_to(info)
        for ex in examples:
            opts = rep.options(example=ex, congestion=c, **flags)
            rep.print_example_info(opts)
            t0 = time.perf_counter()
            try:
                sol, geo, hist = rep.run_example(opts, solver=b200.solver)
            except Exception as exc:                     # keep the sweep going; the failure is part of the record
                row = dict(example=ex, congestion=c, error=f"{type(exc).__name__}: {exc}")
                summary["main"].append(row)
                print(json.dumps(row), flush=True)
                continue
            wall = time.perf_counter() - t0
            row = dict(example=ex, congestion=c, n_vertices=int(geo["vertices"].shape[0]),
                       n_triangles=int(geo["triangles"].shape[0]), iterations=int(hist.kkt_iteration[-1]) + 1,
                       loop_seconds=float(hist.running_time), steps_seconds=float(sum(hist.steps_time.values())),
                       wall_seconds_incl_setup=wall, transport_cost=float(hist.history["Transportation cost"][-1]),
                       final_kkt=max(float(hist.kkt_errors[-1][k]) for k in (0, 2, 4, 5)),
                       mass_violation=rep.mass_conservation(sol["mu"], verbose=False),
                       negative_mass=rep.negative_mass(sol["mu"], verbose=False)[0])
            summary["main"].append(row)
            print(json.dumps(row), flush=True)
            save(outdir, summary)
        fh.close()
        rows = rep.table_rows(rep.parse_log(info))
        with open(os.path.join(sub, "comparison_table.md"), "w") as f:
            f.write(rep.markdown_table(rows) + "\n")
        with open(os.path.join(sub, "comparison_table.csv"), "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=list(rows[0].keys()) if rows else ["Example"])
            w.writeheader()
            w.writerows(rows)


def run_true_error(outdir, flags, summary):
    sub = os.path.join(outdir, "true_error")
    os.makedirs(sub, exist_ok=True)
    info = os.path.join(sub, "info.log")
    if os.path.exists(info):
        os.remove(info)
    fh = log_to(info)
    opts = rep.options(**flags)
    rep.print_example_info(opts)
    sol, geo, hist, err, rows = rep.run_versus_exact(opts, solver=b200.solver)
    fh.close()
    out = dict(n_vertices=int(geo["vertices"].shape[0]), n_time=opts.ntime, iterations=int(hist.kkt_iteration[-1]) + 1,
               loop_seconds=float(hist.running_time), transport_cost=float(hist.history["Transportation cost"][-1]),
               exact_cost=0.5 * 2 * 0.2 ** 2, error=err,
               checkpoints=[dict(kkt_error=float(r["kkt_error"]), iteration=int(r["iteration"]), time=float(r["time"]),
                                 **r["error"]) for r in rows])
    summary["true_error"] = out
    with open(os.path.join(sub, "error_versus_exact.md"), "w") as f:
        f.write("| KKT error | iteration | time [s] | L1 | L2 | Linf |\n|---|---|---|---|---|---|\n")
        for r in out["checkpoints"]:
            f.write(f"| {r['kkt_error']:.2e} | {r['iteration']} | {r['time']:.3f} | {r['l1']:.2e} | {r['l2']:.2e} | {r['linf']:.2e} |\n")
        f.write(f"| final | {out['iterations']} | {out['loop_seconds']:.3f} | {err['l1']:.2e} | {err['l2']:.2e} | {err['linf']:.2e} |\n")
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    ap.add_argument("--suite", choices=["main", "true_error", "all"], default="all")
    ap.add_argument("--outdir", default="gpurun_out/replication")
    ap.add_argument("--examples", nargs="+", default=[e for e in rep.STANDINS if e != "plane"])
    ap.add_argument("--congestions", nargs="+", type=float, default=list(rep.CONGESTIONS))
    ap.add_argument("--tol", type=float, default=None, help="override the protocol's tolerance")
    ap.add_argument("--ntime", type=int, default=None)
    ap.add_argument("--n_space", type=int, default=100, help="plane resolution of the true_error suite")
    args = ap.parse_args()
    os.makedirs(args.outdir, exist_ok=True)
    summary = {"main": [], "true_error": None}
    if args.suite in ("main", "all"):
        flags = dict(rep.MAIN_FLAGS)
        if args.tol:
            flags["tol"] = args.tol
        if args.ntime:
            flags["ntime"] = args.ntime
        run_main(args.outdir, args.examples, args.congestions, flags, summary)
    if args.suite in ("true_error", "all"):
        flags = dict(rep.TRUE_ERROR_FLAGS, n_space=args.n_space)
        if args.tol:
            flags["tol"] = args.tol
        if args.ntime:
            flags["ntime"] = args.ntime
        run_true_error(args.outdir, flags, summary)
    save(args.outdir, summary)


if __name__ == "__main__":
    main()
