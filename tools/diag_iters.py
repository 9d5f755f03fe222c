"""Where does a GPU run leave the reference's KKT history?  python tools/diag_iters.py <fixture> [leaf] [sweep_mode]
Prints iteration counts, the first iteration whose evaluation pattern differs and the largest relative deviation of the
recorded residuals before that point (fixtures: tests/golden/*.npz, written by the unmodified reference)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np                       # noqa: E402
from conftest import load_golden         # noqa: E402
import dots_socp_b200 as b200            # noqa: E402

name = sys.argv[1]
leaf = int(sys.argv[2]) if len(sys.argv) > 2 else 16
if len(sys.argv) > 3:
    os.environ["DOTS_SWEEP_MODE"] = sys.argv[3]
z, geo, n_time, kw = load_golden(name)
sol, hist, eng = b200.solver_socp(n_time, geo, leaf_size=leaf, return_engine=True, **kw)
ref, got = z["kkt_rows"], hist.kkt_errors
n = min(len(ref), len(got))
print(f"{name} leaf={leaf} sweep_mode={eng.sweep_mode} m_pad={eng.m_pad}: iterations {int(hist.kkt_iteration[-1])} (reference {int(z['iterations'])})")
pat = np.isnan(ref[:n]) != np.isnan(got[:n])
first = int(np.argmax(pat.any(axis=1))) if pat.any() else None
print("first row with a different evaluation pattern:", first, "(reference iteration", None if first is None else int(z["kkt_iteration"][first]), ")")
upto = n if first is None else first
m = ~np.isnan(ref[:upto])
rel = np.abs(got[:upto][m] - ref[:upto][m]) / np.abs(ref[:upto][m])
print(f"rows compared {upto}; max rel deviation of the residuals {rel.max():.3e}; median {np.median(rel):.3e}")
worst = np.unravel_index(np.nanargmax(np.where(m, np.abs(got[:upto] - ref[:upto]) / np.abs(ref[:upto]), 0)), ref[:upto].shape)
print("worst at row", worst, "ref", ref[worst], "got", got[worst])
if first is not None:
    lo = max(0, first - 2)
    for r in range(lo, min(n, first + 2)):
        print("row", r, "ref", ref[r], "\n      got", got[r])
r_ref = z["r_history"]
print("penalty path equal up to", int(np.argmax(~np.isclose(r_ref[:min(len(r_ref), n)], r_ref[:min(len(r_ref), n)]))) if False else "n/a")
