"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm (the oracle port on the
host cores) prints ONE JSON line with the agreed keys on the configuration it names, the ranks other than 0 of a torchrun
launch exit quietly, the own arm has no CPU fallback, and the byte counts of the roofline follow SURVEY.md section 8(d)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line_on_the_configuration_it_names():
    res = _run(["--impl", "reference", "--workload", "icosphere3_nt31", "--steps", "2", "--warmup", "1", "--ref-seconds", "30"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "alm_iterations_per_second" and line["unit"] == "iter/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["same_config"] is True and line["extrapolated"] is False and line["fallback"] is None
    cfg = line["config"]
    assert cfg["workload"] == "icosphere3_nt31" and cfg["n_vertices"] == 642 and cfg["n_triangles"] == 1280 and cfg["n_time"] == 31
    assert 1 <= line["steps"] <= 2 and line["value"] > 0
    assert abs(line["value"] - 1e3 / line["ms_per_step"]) <= 1e-9 * line["value"]
    cb = line["cpu_baseline"]
    assert cb["kind"].startswith("port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["unit"] == "iter/s"
    assert cb["steps_measured"] == line["steps"] and "icosphere3_nt31" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_zero_only():
    res = _run(["--impl", "reference", "--workload", "icosphere3_nt31"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_own_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    res = _run(["--workload", "icosphere3_nt31", "--steps", "2", "--no-cpu", "--no-secondary"], env={"CUDA_VISIBLE_DEVICES": ""})
    assert res.returncode != 0 and res.stdout.strip() == ""


def test_roofline_byte_counts_follow_survey_8d():
    sys.path.insert(0, ROOT)
    import bench
    V, T, nT, m = 642, 1280, 31, 32
    a, c, b, z = nT * V, (nT + 1) * V, 3 * (nT + 1) * T, 18 * nT * T
    assert bench.sizes(V, T, nT) == dict(a=a, c=c, b=b, z=z)
    entries = 12345
    assert bench.reference_bytes_per_iteration(V, T, nT, entries, m) == 8 * (27 * a + 10 * b + 8 * c + 7 * z) + 2 * entries * m * 8

    class Sym:
        panel_entries = entries
        upd_off = np.array([0, 10, 25])
    kb = bench.kernel_bytes(V, T, nT, m, Sym)
    assert kb["sweeps"] == 8 * m * (2 * entries + 2 * V)                         # factor streamed twice + the vector in and out
    assert kb["sweeps_incl_work_vectors"] == kb["sweeps"] + 8 * m * (3 * V + 2 * 25)
    assert kb["k_time_mma"] == 2 * 8 * (c + V * m)
    assert all(v > 0 for v in kb.values())
    # every fixture the secondary block compares with exists and carries the reference's iteration count
    for fixture in bench.SECONDARY.values():
        if fixture is not None:
            assert bench.reference_iterations(fixture) > 0
