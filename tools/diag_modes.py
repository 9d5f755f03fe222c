import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dots_socp_b200 import synth
from dots_socp_b200.engine import Engine
geo, _ = synth.example("icosphere4")
res = {}
for tag, kw in (("plain", {}), ("stored", dict(write_z=True)), ("fused", dict(kkt1=True))):
    eng = Engine(31, geo, congestion=0.05)
    eng.use_graphs = bool(int(os.environ.get("GRAPHS", "1")))
    eng.scale_z(2.0)
    eng.iterate(4)
    eng.iterate(3, **kw)
    res[tag] = eng.get_state(("phi", "mu", "B", "E", "b_mid", "b_fst"))
    eng.close()
for a, b in (("plain", "stored"), ("plain", "fused"), ("stored", "fused")):
    for k in res[a]:
        x, y = res[a][k], res[b][k]
        d = np.abs(x - y)
        print(a, b, k, "n_diff", int((d > 0).sum()), "of", d.size, "max abs", d.max(), "max rel", (d / np.maximum(np.abs(x), 1e-300)).max() if d.max() > 0 else 0.0)
