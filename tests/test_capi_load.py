"""CPU-only: the C-ABI library builds/loads and exports every symbol include/dots_b200.h declares."""
import ctypes
import os
import re

from dots_socp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dots_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dots_[a-z_0-9]+)\s*\(", text)))


def test_library_loads_and_exports_header_symbols():
    lib = capi.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(capi.EXPORTS)
    assert lib.dots_abi_version() == capi.ABI_VERSION
    assert lib.dots_ctx_sizeof() == ctypes.sizeof(capi.DotsCtx)


def test_bad_context_is_rejected_without_touching_the_gpu():
    lib = capi.load()
    ctx = capi.DotsCtx()
    ctx.abi_version = 999
    assert lib.dots_step_phi(ctypes.byref(ctx), None) != 0
    assert b"abi" in lib.dots_last_error()


def test_engine_refuses_to_run_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dots_socp_b200 import synth
    from dots_socp_b200.engine import Engine
    geo, _ = synth.example("icosphere1")
    with pytest.raises(capi.DotsError):
        Engine(3, geo)
