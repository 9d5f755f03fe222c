"""Replication harness around the solver plug-in (BASELINE.json configs[4]).

What the reference does with ``make main`` / ``make true_error`` (Makefile:50-108, replication/main.py,
replication/main_versus_exact.py, replication/log2table.py) is a *caller* of the hot path: load an example, normalise it,
call ``solver(n_time, geometry, **kw)``, de-scale the cost, print the mass diagnostics and the run history into one
``info.log``, then regex the log into a table.  The reference package does not exist on the GPU box and its bundled
``.off`` meshes are git-LFS stubs, so this module restates that caller for generated stand-in surfaces and writes logs in
the same text format (the format is the interface: log2table.py:98-106):

* ``STANDINS``                    - one generated surface per ``Makefile:56-61`` example name (sizes 1.6k ... 68k vertices);
* ``write_off`` / ``read_off``    - the OFF dialect ``data/util.py:73-143`` reads;
* ``run_example``                 - the ``run_dot_surface`` flow (interface.py:223-335) without rendering;
* ``exact_plane`` / ``compare_with_exact`` / ``run_versus_exact`` - the analytic ``plane`` study
  (data/settings/plane.py:29-46, data/load_example.py:153-200, utils/evaluate_solution.py:47-69, interface.py:386-480);
* ``parse_log`` / ``table_rows``  - the log -> table step.

Host-side numpy only; the solver call is the B200 path."""
from __future__ import annotations

import logging
import math
import re
from types import SimpleNamespace

import numpy as np

from . import synth
from .history import LOG_INFO, _banner

CONGESTIONS = (0.00, 0.01, 0.05)                                                   # Makefile:62
MAIN_FLAGS = dict(ntime=31, nit=10000, time_limit=5000, tol=1e-4)                  # Makefile:50,65-66
TRUE_ERROR_FLAGS = dict(example="plane", tol=1e-5, nit=20000)                      # Makefile:101-108

# example name (Makefile:56-61) -> (generator, kwargs, masses recipe).  The scanned models are replaced by smooth closed
# surfaces of comparable vertex counts; "refined_*" is one subdivision level finer, like the bundled refined meshes.
STANDINS = {
    "airplane": (synth.deformed_sphere, dict(level=4, axes=(1.0, 0.9, 0.25)), "vertex"),
    "refined_airplane": (synth.deformed_sphere, dict(level=5, axes=(1.0, 0.9, 0.25)), "vertex"),
    "armadillo": (synth.deformed_sphere, dict(level=4, axes=(0.8, 1.0, 0.7), amp=0.15, freq=3), "vertex"),
    "refined_armadillo": (synth.deformed_sphere, dict(level=5, axes=(0.8, 1.0, 0.7), amp=0.15, freq=3), "vertex"),
    "hand": (synth.deformed_sphere, dict(level=4, axes=(1.0, 0.6, 0.3), amp=0.1, freq=2), "vertex"),
    "refined_hand": (synth.deformed_sphere, dict(level=5, axes=(1.0, 0.6, 0.3), amp=0.1, freq=2), "vertex"),
    "punctured_ball": (synth.punctured_sphere, dict(level=4), "vertex"),
    "refined_punctured_ball": (synth.punctured_sphere, dict(level=5), "vertex"),
    "bunny": (synth.deformed_sphere, dict(level=5, axes=(0.8, 0.7, 1.0), amp=0.08, freq=2), "vertex"),
    "refined_bunny": (synth.torus, dict(n_u=340, n_v=200, big_r=1.0, small_r=0.55), "vertex"),
    "ring": (synth.torus, dict(n_u=120, n_v=40), "vertex"),
    "knots_3": (synth.knot_tube, dict(p=2, q=3, n_u=300, n_v=10), "vertex"),
    "knots_5": (synth.knot_tube, dict(), "knot"),
    "hills": (synth.hills, dict(n=60), "vertex"),
    "plane": (synth.hex_plane, dict(n=100), "plane"),
}


def load_standin(name: str, n_space=None, seed: int = 0):
    """``load_example`` (data/load_example.py:100-151) for the stand-ins: the RAW (un-normalised) GeometryData dict."""
    if name not in STANDINS:
        raise ValueError(f"unknown example {name!r}; known: {sorted(STANDINS)}")
    gen, kw, masses = STANDINS[name]
    kw = dict(kw)
    if name == "plane" and n_space is not None:
        kw["n"] = int(n_space)
    v, t = gen(**kw)
    return synth.make_geometry(v, t, masses=masses, seed=seed)


# ----------------------------------------------------------------------------- OFF files
def write_off(path, vertices, triangles):
    """The dialect ``read_mesh_off`` (data/util.py:73-143) accepts: 'OFF', 'nV nT nE', vertex lines, '3 a b c' lines.
    Coordinates are written with 17 significant digits so that a round trip is bit-exact."""
    v, t = np.asarray(vertices, dtype=np.float64), np.asarray(triangles)
    with open(path, "w") as f:
        f.write(f"OFF\n{v.shape[0]} {t.shape[0]} 0\n")
        for x, y, z in v:
            f.write(f"{float(x)!r} {float(y)!r} {float(z)!r}\n")
        for a, b, c in t:
            f.write(f"3 {int(a)} {int(b)} {int(c)}\n")


def read_off(path):
    """Returns (vertices (V,3) f64, triangles (T,3) int64, edges (3T,2)) with the reference reader's conventions and
    errors (data/util.py:73-143): a line starting with the token '3' is a triangle, anything else a vertex."""
    with open(path, "r") as f:
        if f.readline().strip() != "OFF":
            raise ValueError("Not a valid .off file")
        head = f.readline().split()
        if len(head) < 2:
            raise ValueError("Invalid file format: missing vertex/triangle counts")
        n_v, n_t = int(head[0]), int(head[1])
        verts, tris = [], []
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "3":
                if len(tok) < 4:
                    raise ValueError(f"Invalid triangle data at line {len(tris) + 1}")
                tris.append((int(tok[1]), int(tok[2]), int(tok[3])))
            else:
                if len(tok) < 3:
                    raise ValueError(f"Invalid vertex data at line {len(verts) + 1}")
                verts.append((float(tok[0]), float(tok[1]), float(tok[2])))
    if len(verts) != n_v:
        raise ValueError(f"Expected {n_v} vertices but found {len(verts)}")
    if len(tris) != n_t:
        raise ValueError(f"Expected {n_t} triangles but found {len(tris)}")
    v = np.array(verts, dtype=np.float64).reshape(-1, 3)
    t = np.array(tris, dtype=np.int64).reshape(-1, 3)
    return v, t, t[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2)


# ----------------------------------------------------------------------------- diagnostics on the returned mu
def mass_conservation(mu, verbose=True):
    """utils/evaluate_solution.py:7-22: rms deviation of the per-layer mass from 1."""
    layers = np.asarray(mu).sum(axis=1)
    err = float(np.linalg.norm(layers - 1.0) / np.sqrt(layers.shape[0]))
    if verbose:
        with np.printoptions(precision=4, suppress=True):
            logging.log(LOG_INFO, f"{_banner('Mass Conservation')}\nSum of Mass at each time layer:\n{layers}\n"
                                  f"Mass Conservation Violation: {err:.2e}")
    return err


def negative_mass(mu, verbose=True):
    """utils/evaluate_solution.py:24-45: rms of the per-layer sums of the negative entries."""
    mu = np.asarray(mu)
    layers = np.where(mu < 0, mu, 0.0).sum(axis=1)
    err = float(np.linalg.norm(layers) / np.sqrt(layers.shape[0]))
    if verbose:
        with np.printoptions(precision=4, suppress=True):
            logging.log(LOG_INFO, f"{_banner('Negative Mass')}\nSum of Negative Mass at each time layer:\n{layers}\n"
                                  f"Non-Negative Mass Violation: {err:.2e}")
    return err, layers


# ----------------------------------------------------------------------------- the run_dot_surface flow
def print_example_info(opts):
    """interface.py:25-73: the block log2table keys its rows on."""
    fields = ["example", "mesh_file", "setting_file", "congestion", "ntime", "tol", "tau", "eps", "nit", "power_perceptual"]
    lines = [f"{k}: {getattr(opts, k)}" for k in fields if getattr(opts, k, None) is not None]
    logging.log(LOG_INFO, "")
    logging.log(LOG_INFO, _banner("Info: Experiment Setting") + "\n" + "\n".join(lines))


def run_example(opts, solver=None, geometry=None):
    """``run_dot_surface`` (interface.py:106-383) minus rendering, for ``opts.example`` in STANDINS (or a given RAW
    geometry).  ``opts``: any namespace with ``example, ntime`` and optionally ``congestion, tol, nit, tau, eps,
    time_limit, checkpoints, detail_runhist, n_space``.  Returns (solution, raw geometry, run history)."""
    if solver is None:
        from .solver import solver
    if not callable(solver):
        raise TypeError("Solver must be a callable function")
    if not hasattr(opts, "ntime") or opts.ntime <= 0:
        raise ValueError("'ntime' must be a positive integer")
    tau = getattr(opts, "tau", None)
    if tau is not None and (tau <= 0 or tau > 2):
        raise ValueError("'tau' must be in range (0, 2]")
    for key, msg in (("tol", "'tol' must be positive"), ("nit", "'nit' must be positive"),
                     ("time_limit", "'time_limit' must be positive")):
        val = getattr(opts, key, None)
        if val is not None and val <= 0:
            raise ValueError(msg)
    for key, msg in (("congestion", "'congestion' must be non-negative"), ("eps", "'eps' must be non-negative")):
        val = getattr(opts, key, None)
        if val is not None and val < 0:
            raise ValueError(msg)

    if geometry is None:
        geometry = load_standin(opts.example, n_space=getattr(opts, "n_space", None))
    logging.log(LOG_INFO, _banner("Discretization") + "\n"
                f"Example name: {opts.example}\nNumber of points in time: {opts.ntime}\n"
                f"Number of vertices: {geometry['vertices'].shape[0]}\n"
                f"Number of triangles: {geometry['triangles'].shape[0]}\n"
                f"Area of the vertices: {np.sum(geometry['area_vertices'] / 3.0)}\n"
                f"Area of the triangles: {np.sum(geometry['area_triangles'])}")
    key_mapping = dict(eps="eps", tau="tau", nit="nit", tol="tol", congestion="congestion", checkpoints="tol_checkpoints",
                       time_limit="time_limit", detail_runhist="check_kkt_step_by_step")          # interface.py:275-284
    kwargs = {dst: getattr(opts, src) for src, dst in key_mapping.items() if getattr(opts, src, None) is not None}
    for extra in ("device", "comm", "leaf_size"):                                                # additions of this package
        if getattr(opts, extra, None) is not None:
            kwargs[extra] = getattr(opts, extra)

    normalized, scale = synth.normalize_geometry(geometry)
    solution, hist = solver(opts.ntime, normalized, **kwargs)
    if not isinstance(solution, dict) or "mu" not in solution:
        raise ValueError("Solver must return a solution dictionary containing 'mu' key")
    for key in ("Transportation cost", "Objective value"):                                        # interface.py:302-308
        if key in hist.history:
            hist.history[key] = hist.history[key] / scale ** 2
    mass_conservation(solution["mu"])
    negative_mass(solution["mu"])
    hist.print_end_history()
    hist.print_steps_time()
    return solution, geometry, hist


# ----------------------------------------------------------------------------- analytic plane study
PLANE_C0, PLANE_C1 = np.array([0.4, 0.4, 0.0]), np.array([0.6, 0.6, 0.0])          # data/settings/plane.py:5-11
PLANE_S0 = PLANE_S1 = 2 * (0.1 ** 2)


def exact_plane(t_array, vertices, area_vertices):
    """Displacement interpolation of two Gaussians (data/settings/plane.py:29-46), normalised like
    data/load_example.py:190-194 (by the mean of the first and last layer masses).  Vectorised over vertices."""
    out = np.zeros((len(t_array), vertices.shape[0]))
    q0, q1 = PLANE_S0 ** 0.25, PLANE_S1 ** 0.25
    for i, t in enumerate(t_array):
        sigma = ((1 - t) * q0 + t * q1) ** 4
        centre = (1 - t) * PLANE_C0 + t * PLANE_C1
        out[i] = area_vertices * np.exp(-np.linalg.norm(vertices - centre[None, :], axis=1) ** 2 / sigma)
    return out / (0.5 * (out[0].sum() + out[-1].sum()))


def compare_with_exact(mu, mu_exact, geometry, verbose=True):
    """utils/evaluate_solution.py:47-69 with utils/util.py:31-68: relative L1 / L2 / Linf error of the DENSITY."""
    w = np.asarray(geometry["area_vertices"])[None, :] / 3.0
    a, b = np.asarray(mu) / w, np.asarray(mu_exact) / w
    d = a - b
    h = 1.0 / d.shape[0]
    l1 = lambda x: float(np.sum(np.abs(x) * w) * h)
    l2 = lambda x: float(np.sqrt(np.sum(np.square(x) * w) * h))
    linf = lambda x: float(np.max(np.abs(x)))
    err = dict(l1=l1(d) / (1.0 + l1(b)), l2=l2(d) / (1.0 + l2(b)), linf=linf(d) / (1.0 + linf(b)))
    if verbose:
        logging.log(LOG_INFO, _banner("Versus exact transportation") + "\n"
                    f"L_1 Error: {err['l1']:.2e}\nL_2 Error: {err['l2']:.2e}\nL_Inf Error: {err['linf']:.2e}")
    return err


def automatic_checkpoints(tol: float):
    """replication/main_versus_exact.py:43-50: 1e-1, 1e-2, ... down to tol."""
    raw = -math.log(tol, 10)
    n = int(round(raw, 12) if abs(raw - round(raw)) < 1e-12 else raw)
    return [10 ** (-i - 1) for i in range(n)]


def run_versus_exact(opts, solver=None):
    """``run_dot_surface_versus_exact`` (interface.py:386-480) on the centred time grid for the ``plane`` example."""
    if opts.example != "plane":
        raise ValueError("only 'plane' defines an exact transportation (data/settings/plane.py:29)")
    if not getattr(opts, "checkpoints", None):
        opts.checkpoints = automatic_checkpoints(opts.tol)
    geometry = load_standin("plane", n_space=getattr(opts, "n_space", None))
    exact = exact_plane(np.linspace(0.0, 1.0, opts.ntime + 1), geometry["vertices"], geometry["area_vertices"])
    solution, geometry, hist = run_example(opts, solver=solver, geometry=geometry)
    error = compare_with_exact(solution["mu"], exact, geometry)
    rows = []
    for cp in solution.get("checkpoints") or []:
        rows.append(dict(error=compare_with_exact(cp["mu"], exact, geometry, verbose=False),
                         kkt_error=max(k for k in cp["kkt"] if k is not None), iteration=cp["iteration"], time=cp["time"]))
    return solution, geometry, hist, error, rows


# ----------------------------------------------------------------------------- log -> table
_BLOCK = re.compile(r".*Info: Experiment Setting.*")
_FIELDS = (("Example", re.compile(r"^Example name:\s*(\S+)")),                                       # log2table.py:98-106
           ("Vertices", re.compile(r"^Number of vertices:\s*(\d+)")),
           ("Triangles", re.compile(r"^Number of triangles:\s*(\d+)")),
           ("Transport Cost", re.compile(r"^Transportation cost:\s*([-+]?\d+\.\d+e[-+]?\d+)")),
           ("Time [seconds]", re.compile(r"^Time of steps\s*:\s*(\d+\.?\d*)\s*sec")),
           ("Iterations", re.compile(r"^Total Iteration(?:\s*\(l\.l\.\))?\s*:\s*(\d+) iterations")))


def parse_log(path):
    """One dict per 'Info: Experiment Setting' block in which all six fields were found (log2table.py:40-88)."""
    with open(path, "r") as f:
        lines = f.readlines()
    starts = [i for i, ln in enumerate(lines) if _BLOCK.match(ln)] + [len(lines)]
    rows = []
    for lo, hi in zip(starts[:-1], starts[1:]):
        row = {}
        for name, pat in _FIELDS:
            for ln in lines[lo + 1:hi]:
                m = pat.search(ln)
                if m:
                    row[name] = m.group(1)
                    break
        if len(row) == len(_FIELDS):
            rows.append(row)
    return rows


def table_rows(rows):
    """First run per example, typed and titled like the reference table (log2table.py:122-131)."""
    seen, out = set(), []
    for r in rows:
        if r["Example"] in seen:
            continue
        seen.add(r["Example"])
        out.append({"Example": r["Example"].replace("_", " ").title(), "Vertices": int(r["Vertices"]),
                    "Triangles": int(r["Triangles"]), "Iterations": int(r["Iterations"]),
                    "Time [seconds]": float(r["Time [seconds]"]), "Transport Cost": round(float(r["Transport Cost"]), 4)})
    return sorted(out, key=lambda r: r["Example"])


def markdown_table(rows):
    cols = ["Example", "Vertices", "Triangles", "Iterations", "Time [seconds]", "Transport Cost"]
    lines = ["| " + " | ".join(cols) + " |", "|" + "---|" * len(cols)]
    lines += ["| " + " | ".join(str(r[c]) for c in cols) + " |" for r in rows]
    return "\n".join(lines)


def options(**kw):
    """Namespace with the CLI's defaults that matter here (cli.py:27-139)."""
    base = dict(example=None, mesh_file=None, setting_file=None, congestion=0.0, ntime=31, tol=1e-3, tau=None, eps=0.0,
                nit=1000, time_limit=float("inf"), checkpoints=None, detail_runhist=False, power_perceptual=1.0,
                n_space=None)
    base.update(kw)
    return SimpleNamespace(**base)
