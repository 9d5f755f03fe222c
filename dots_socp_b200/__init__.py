"""dots_socp_b200 - B200-native (sm_100a) implementation of the inner ALM iteration of DOTs-SOCP.

Public surface = the reference's solver plug-in API (dot_surface_socp/__init__.py:9-25):
``solver`` / ``solver_raw`` / ``solver_socp``; pass ``solver`` as ``run_dot_surface(opts, solver=...)``.
Importing the package needs neither a GPU nor the CUDA library; calling a solver does."""
from .solver import solver, solver_raw, solver_socp  # noqa: F401

__all__ = ["solver", "solver_raw", "solver_socp"]
__version__ = "0.1.0"
