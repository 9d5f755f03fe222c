"""ctypes binding of libdots_b200.so (C-ABI declared in include/dots_b200.h).

The library is the product: there is NO fallback.  ``load()`` raises if the shared object is missing
or its ABI does not match this mirror of ``dots_ctx_t``."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

ABI_VERSION = 19
P_R, P_S, P_D, P_CONG, P_TAU, P_EPS, P_PS, P_DS, P_COUNT = 0, 1, 2, 3, 4, 5, 6, 7, 8

_i32p, _i64p, _f64p = C.c_void_p, C.c_void_p, C.c_void_p      # raw device/host addresses


class DotsCtx(C.Structure):
    """Field order mirrors ``struct dots_ctx`` exactly (checked against dots_ctx_sizeof() at load)."""
    _fields_ = (
        [(n, C.c_int32) for n in ("abi_version", "n_time", "n_vert", "n_tri", "m_pad", "n_nodes", "n_levels", "n_sm")]
        + [(n, C.c_void_p) for n in (
            "tri", "hat_grad", "area_f", "area_v", "diag_soc", "vc_ptr", "vc_idx", "vc_ell", "qf", "qb",
            "panels", "panels_t", "nd_off", "nd_s", "nd_b", "nd_child", "nd_panel", "nd_front", "nd_upd",
            "front_idx", "child_pos", "lvl_ptr", "lvl_items", "lvb_ptr", "lvb_items", "lvn_nodes", "h_lvl_ptr", "h_lvb_ptr",
            "h_lvn_ptr", "h_lvl_wpr", "h_lvb_cw")]
        + [("front_total", C.c_int64)]
        + [(n, C.c_int32) for n in ("lvl_begin", "lvl_end", "n_ranks", "tt_kf", "tt_kb", "tt_nb", "tt_nout", "tt_sym")]
        + [(n, C.c_void_p) for n in (
            "params", "phi", "A", "lam_c", "mu", "z_fst", "z_end", "b_fst", "b_end", "lam",
            "bnd0", "bnd1", "B", "E", "b_mid", "z_mid", "corner_nrm", "corner_div",
            "rhs", "hat", "hat_all", "ywork", "upd", "red_part", "red_out")]
        + [("red_blocks", C.c_int32), ("sweep_mode", C.c_int32), ("sweep_grid", C.c_int32), ("reserved0", C.c_int32),
           ("phase_clock", C.c_void_p), ("peer_vertex", C.c_void_p * 4), ("peer_corner", C.c_void_p),
           ("peer_rhs", C.c_void_p * 8), ("peer_hat", C.c_void_p * 8)]
        + [(n, C.c_void_p) for n in ("rt_fwd", "rt_bwd", "h_rt_fwd_ptr", "h_rt_bwd_ptr", "h_rt_fwd_wpr", "h_rt_bwd_wpr",
                                     "bidx", "erow_fwd", "erow_bwd", "gptr", "gidx", "gverts", "h_gv_ptr")]
        + [("ring_stages", C.c_int32), ("ring_pdl", C.c_int32), ("ring_stage_bytes", C.c_int32), ("ring_flags", C.c_int32),
           ("kkt1_part", C.c_void_p), ("kkt1_blocks", C.c_int32), ("reserved3", C.c_int32)]
    )


class FrontArgs(C.Structure):
    """Mirror of ``dots_front_args_t`` (setup: small-front factorisation launch)."""
    _fields_ = ([("nodes", C.c_void_p), ("n_modes", C.c_int32), ("m_pad", C.c_int32)]
                + [(n, C.c_void_p) for n in ("nd_off", "nd_s", "nd_b", "nd_child", "nd_panel", "nd_front", "nd_upd", "parent_pos",
                                             "u_ptr", "a_ptr", "a_pos", "a_val", "mass", "shifts", "panels", "panels_t")]
                + [("pin_node", C.c_int32), ("reserved", C.c_int32), ("pin_value", C.c_double)])


class DotsError(RuntimeError):
    pass


_lib = None


def lib_path() -> str:
    return os.environ.get("DOTS_LIB", _build.LIB)            # DOTS_LIB: A/B-test an alternative build of the same sources


def load(build_if_missing: bool = True):
    """Load (building with nvcc first when the .so is absent or stale and ``build_if_missing``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and "DOTS_LIB" not in os.environ and _build.needs_build():
        try:
            _build.build()
        except Exception as exc:                      # no nvcc on the box: use the shipped .so if there is one
            if not os.path.exists(path):
                raise DotsError(f"libdots_b200.so is missing and could not be built: {exc}") from exc
    if not os.path.exists(path):
        raise DotsError(f"{path} not found: run `python -m dots_socp_b200.build` (the CUDA library is required)")
    lib = C.CDLL(path)
    lib.dots_last_error.restype = C.c_char_p
    if lib.dots_abi_version() != ABI_VERSION:
        raise DotsError(f"ABI mismatch: library {lib.dots_abi_version()} binding {ABI_VERSION}")
    if lib.dots_ctx_sizeof() != C.sizeof(DotsCtx):
        raise DotsError(f"dots_ctx_t size mismatch: library {lib.dots_ctx_sizeof()} binding {C.sizeof(DotsCtx)}")
    ctxp, vp, i, d = C.POINTER(DotsCtx), C.c_void_p, C.c_int, C.c_double
    protos = {
        "dots_step_phi": (ctxp, vp), "dots_step_vertex": (ctxp, vp), "dots_step_q0": (ctxp, vp), "dots_step_tri": (ctxp, i, vp),
        "dots_iterate": (ctxp, i, i, vp), "dots_refresh_corner_terms": (ctxp, vp),
        "dots_scale_dual": (ctxp, d, vp), "dots_scale_z": (ctxp, d, vp), "dots_scale_prim_dual": (ctxp, d, d, vp), "dots_set_params": (ctxp, vp, vp),
        "dots_kkt_sums": (ctxp, i, vp, vp), "dots_kkt_sums_multi": (ctxp, C.c_uint, vp, vp), "dots_phi_rhs": (ctxp, vp), "dots_time_transform": (ctxp, i, vp),
        "dots_time_transform_plain": (ctxp, i, vp, i, vp),
        "dots_mode_solves": (ctxp, vp), "dots_grad_space": (ctxp, vp, vp, vp), "dots_div_space": (ctxp, vp, vp, vp),
        "dots_graph_create": (ctxp, i, vp, C.POINTER(vp)), "dots_graph_launch": (vp, vp), "dots_graph_destroy": (vp,),
        "dots_factor_small_fronts": (C.POINTER(FrontArgs), i, i, vp), "dots_front_nmax": (),
        "dots_factor_large_fronts": (C.POINTER(FrontArgs), i, i, i, i, vp, vp, vp, vp, vp), "dots_enable_peer": (i,),
        "dots_ipc_export": (vp, vp, C.POINTER(C.c_ulonglong)), "dots_ipc_import": (vp, C.c_ulonglong, C.POINTER(vp)),
        "dots_ring_level_times": (ctxp, vp, vp, vp, i, C.POINTER(C.c_int)),
    }
    i64 = C.c_int64
    protos.update({
        "dots_order_create": (i64, vp, vp, vp, i64, C.POINTER(vp)), "dots_order_sizes": (vp, C.POINTER(i64), C.POINTER(i64)),
        "dots_order_export": (vp,) * 9, "dots_order_destroy": (vp,),
        "dots_ring_entry_rows": (i64,) + (vp,) * 8,
        "dots_mesh_create": (i64, i64, vp, vp, C.POINTER(vp)), "dots_mesh_sizes": (vp, C.POINTER(i64)),
        "dots_mesh_export": (vp,) * 10, "dots_mesh_destroy": (vp,), "dots_corner_lists": (i64, i64, vp, vp, vp),
        "dots_front_maps": (i64, i64) + (vp,) * 11, "dots_csr_permute": (i64,) + (vp,) * 8,
    })
    for name, args in protos.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = list(args), C.c_int
    _lib = lib
    return lib


EXPORTS = ("dots_abi_version", "dots_ctx_sizeof", "dots_last_error", "dots_step_phi", "dots_step_vertex",
           "dots_step_tri", "dots_step_q0", "dots_iterate", "dots_refresh_corner_terms", "dots_scale_dual", "dots_scale_z", "dots_scale_prim_dual",
           "dots_set_params", "dots_kkt_sums", "dots_kkt_sums_multi", "dots_phi_rhs", "dots_time_transform", "dots_time_transform_plain", "dots_mode_solves",
           "dots_grad_space", "dots_div_space", "dots_graph_create", "dots_graph_launch", "dots_graph_destroy",
           "dots_factor_small_fronts", "dots_front_nmax", "dots_factor_large_fronts", "dots_enable_peer", "dots_ipc_export", "dots_ipc_import",
           "dots_ring_level_times", "dots_ring_entry_rows",
           "dots_order_create", "dots_order_sizes", "dots_order_export", "dots_order_destroy",
           "dots_mesh_create", "dots_mesh_sizes", "dots_mesh_export", "dots_mesh_destroy", "dots_corner_lists", "dots_front_maps",
           "dots_csr_permute")


def check(code: int, what: str = ""):
    if code != 0:
        msg = load().dots_last_error()
        raise DotsError(f"{what} failed with code {code}: {msg.decode() if msg else ''}")
