"""Import the UNMODIFIED reference (``/root/reference``) in the build container.

Test infrastructure only (never imported by the product, bench.py or the ``-m gpu`` tests: the
reference tree does not exist on the GPU box).  Three of the reference's imports are not installed
here, so tiny stand-ins are registered in ``sys.modules`` before the import:

* ``numexpr``            - ``evaluate`` runs the expression with numpy in the caller's frame
* ``matplotlib.pyplot``  - empty module (only used for optional plots)
* ``trimesh``            - ``Trimesh(vertices, faces)`` with ``centroid`` / ``vertices`` / ``faces`` / ``edges``

``load()`` returns the imported ``dot_surface_socp`` package (cwd is switched to the reference
root because its path_config.toml is CWD-relative).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("DOTS_REFERENCE_ROOT", "/root/reference")


def _numexpr_module():
    m = types.ModuleType("numexpr")
    ns = {k: getattr(np, k) for k in ("sqrt", "exp", "log", "abs", "where", "sin", "cos", "tan", "arctan2")}

    def evaluate(expr, local_dict=None, global_dict=None, out=None, **_):
        frame = sys._getframe(1)
        scope = {}
        if local_dict is None:
            scope.update(frame.f_globals)
            scope.update(frame.f_locals)
        else:
            scope.update(local_dict)
        scope.update(ns)          # numexpr's own function names win (the reference also imports math.sqrt)
        val = eval(expr, {"__builtins__": {}}, scope)
        if out is not None:
            out[...] = val
            return out
        return np.asarray(val)

    m.evaluate = evaluate
    m.set_num_threads = lambda n: n
    m.detect_number_of_cores = lambda: os.cpu_count() or 1
    m.__version__ = "shim"
    return m


def _trimesh_module():
    m = types.ModuleType("trimesh")

    class Trimesh:
        def __init__(self, vertices=None, faces=None, process=False, **_):
            self.vertices = np.array(vertices, dtype=np.float64)
            self.faces = np.array(faces)

        @property
        def centroid(self):
            tri = self.vertices[self.faces]
            area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
            return (tri.mean(axis=1) * area[:, None]).sum(axis=0) / area.sum()

        @property
        def edges(self):
            return self.faces[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2)

    m.Trimesh = Trimesh
    return m


def load():
    if not os.path.isdir(REFERENCE_ROOT):
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    if "numexpr" not in sys.modules:
        try:
            import numexpr  # noqa: F401
        except ImportError:
            sys.modules["numexpr"] = _numexpr_module()
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except ImportError:
            mpl = types.ModuleType("matplotlib")
            mpl.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = mpl.pyplot
    if "trimesh" not in sys.modules:
        try:
            import trimesh  # noqa: F401
        except ImportError:
            sys.modules["trimesh"] = _trimesh_module()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    os.chdir(REFERENCE_ROOT)
    import dot_surface_socp
    return dot_surface_socp
