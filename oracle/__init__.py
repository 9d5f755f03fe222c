"""CPU oracle of the DOTs-SOCP ALM hot path - TEST INFRASTRUCTURE, never imported by the product.

See ``alm_oracle.py`` for the parity status ("pinned against reference outputs generated in the
build container", fixtures under ``tests/golden``)."""
from .alm_oracle import MeshOps, OracleALM, solve  # noqa: F401
