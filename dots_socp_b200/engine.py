"""Device-resident ALM state + the calls into libdots_b200.so.

``Engine`` owns what ``solver_socp`` sets up before its loop (reference socp/solver_socp.py:97-270):
mesh operators, the space-time Laplacian inverse, the state arrays and the scalars (r, scale_factor_z,
constant_d).  Everything numeric that happens per iteration is a CUDA kernel behind the C-ABI
(include/dots_b200.h); PyTorch only provides device memory and the stream.  There is no CPU path:
constructing an Engine without a CUDA device or without the library raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time

import numpy as np
import torch

from . import capi, nested, ring_plan, surface
from . import dist as dd

STATE_VERTEX = ("phi", "A", "lam_c", "mu", "z_fst", "z_end", "b_fst", "b_end")
STATE_TRI = ("B", "E")
STATE_CORNER = ("b_mid", "z_mid")
#: reference key (utils/type.py:22-38)  ->  engine field
REF_KEYS = dict(phi="phi", A="A", B="B", lambda_c="lam_c", mu="mu", E="E", z_fst="z_fst", z_mid="z_mid",
                z_end="z_end", beta_fst="b_fst", beta_mid="b_mid", beta_end="b_end")


def time_basis(n_time: int):
    """Eigen-decomposition of the Neumann time Laplacian (laplacian_inverse_socp.py:15-31), analytically:
    Q[t,k] = c_k cos(pi k (t+1/2)/(nT+1)),  lambda_k = -4 nT^2 sin^2(pi k / (2 (nT+1)))."""
    n = n_time + 1
    t = np.arange(n)[:, None]
    k = np.arange(n)[None, :]
    Q = np.cos(np.pi * k * (t + 0.5) / n)
    Q[:, 0] = 1.0 / math.sqrt(n)
    Q[:, 1:] *= math.sqrt(2.0 / n)
    lam = -4.0 * n_time ** 2 * np.sin(np.pi * np.arange(n) / (2.0 * n)) ** 2
    return Q, lam


def _stable_argsort(keys, bound):
    """``np.argsort(keys, kind="stable")`` for integer keys in [0, bound): numpy sorts 16-bit keys with a radix sort, so two
    stable passes over the 16-bit halves (4x faster than its merge sort of 64-bit keys at 330k triangles)."""
    keys = np.asarray(keys)
    if bound > 1 << 32 or keys.size < 1 << 14:
        return np.argsort(keys, kind="stable")
    low = np.argsort((keys & 0xffff).astype(np.uint16), kind="stable")
    if bound <= 1 << 16:
        return low
    return low[np.argsort((keys >> 16).astype(np.uint16)[low], kind="stable")]


def _cut(nodes, lengths, step):
    """Work items (node, first output, n outputs <= step) covering ``lengths[i]`` outputs of ``nodes[i]``, node by node."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n_items = -(-lengths // step)
    total = int(n_items.sum())
    if total == 0:
        return np.zeros((0, 3), dtype=np.int32)
    owner = np.repeat(np.arange(len(nodes)), n_items)
    first = (np.arange(total) - np.repeat(np.cumsum(n_items) - n_items, n_items)) * step
    count = np.minimum(step, lengths[owner] - first)
    return np.stack([np.asarray(nodes, dtype=np.int64)[owner], first, count], axis=1).astype(np.int32)


def _sweep_items(sym: nested.Symbolic, n_sm: int, m_pad: int):
    """Per-level launch plan of the two sweeps (one launch per level and direction).

    Items are (node, first output, n outputs): outputs are panel rows in the forward sweep and panel columns (of the
    column-major copy) in the backward sweep.  ``wpr`` warps share one output: 1 for the short runs of the leaf
    fronts, up to 8 for the long runs of the top separators, so every level exposes >= ~8 blocks per SM whenever it
    has the outputs for it."""
    fwd_ptr, bwd_ptr, node_ptr, fwd, bwd, gather, wprs, cws = [0], [0], [0], [], [], [], [], []

    def pick(len_eff, total):
        wpr = 1 if len_eff < 32 else 2 if len_eff < 96 else 4 if len_eff < 256 else 8
        per_pass = 8 // wpr
        passes = int(min(8, max(1, total // (per_pass * 8 * n_sm))))
        return wpr, per_pass * passes

    has_child = (sym.child >= 0).any(axis=1)
    for nodes in nested.level_schedule(sym):
        s_l, b_l = sym.s[nodes].astype(np.int64), sym.b[nodes].astype(np.int64)
        work = np.maximum(1, s_l * (s_l + 1) // 2 + s_l * b_l)
        s_eff = float((s_l * work).sum() / work.sum())
        col_eff = float(((s_l / 2 + b_l) * work).sum() / work.sum())
        wpr, rb = pick(s_eff, int((s_l + b_l).sum()))
        cw, cb = pick(col_eff, int(s_l.sum()))
        kids = has_child[nodes]
        fuse = bool(kids.any()) and wpr <= 2 and int(s_l.max()) <= 64 and os.environ.get("DOTS_FUSE_GATHER", "1") == "1"
        fwd.append(_cut(nodes, s_l + b_l, rb))
        bwd.append(_cut(nodes, s_l, cb))
        if not fuse:
            gather.append(_cut(nodes[kids], s_l[kids], 32))
        fwd_ptr.append(fwd_ptr[-1] + len(fwd[-1]))
        bwd_ptr.append(bwd_ptr[-1] + len(bwd[-1]))
        node_ptr.append(node_ptr[-1] + (len(gather[-1]) if not fuse else 0))
        wprs.append(wpr + (16 if fuse else 0))
        cws.append(cw)
    cat = lambda parts: (np.ascontiguousarray(np.concatenate(parts, axis=0)) if sum(len(p) for p in parts)
                         else np.zeros((1, 3), np.int32))
    i32 = lambda a: np.ascontiguousarray(np.array(a, dtype=np.int32))
    return dict(fwd_ptr=i32(fwd_ptr), fwd_items=cat(fwd), bwd_ptr=i32(bwd_ptr), bwd_items=cat(bwd),
                node_ptr=i32(node_ptr), nodes=cat(gather), wpr=i32(wprs), cw=i32(cws))


class Engine:
    """One rank's share of the problem.  ``comm`` (dist.Comm) spans the ranks; None / single rank = whole problem."""

    def __init__(self, n_time, geometry, congestion=0.0, eps=0.0, tau=1.9, device=None, leaf_size=16, timings=None,
                 sweep_mode=None, comm=None):
        if not torch.cuda.is_available():
            raise capi.DotsError("dots_socp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = capi.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        torch.cuda.set_device(self.device)
        t_start = time.perf_counter()
        tm = timings if timings is not None else {}
        self.comm = comm if comm is not None else dd.Comm()

        v = np.ascontiguousarray(geometry["vertices"], dtype=np.float64)
        tri_old = np.ascontiguousarray(geometry["triangles"]).astype(np.int64)
        self.nT, self.V, self.T = int(n_time), v.shape[0], tri_old.shape[0]
        nT, V, T = self.nT, self.V, self.T
        self.dt = 1.0 / nT
        self.cong, self.tau, self.eps = float(congestion), float(tau), float(eps)

        # ---- mesh operators (host, vectorised) --------------------------------------------------
        mesh = surface.mesh_operators_native(self.lib, v, tri_old)                      # C++ (csrc/host_order.cpp, dots_mesh_*)
        area_f, hat, K = mesh["area_f"], mesh["hat"], mesh["K"]                         # hat: (T,3,3) [f,k,xyz]
        area_v = mesh["area_sum"] / 3.0                                                 # solver_socp.py:112
        tm["mesh_operators"] = time.perf_counter() - t_start

        # ---- ordering + batched factorisation (this rank's time modes only) -------------------------
        t0 = time.perf_counter()
        sym = nested.analyse_native(self.lib, v, K, leaf_size=leaf_size)               # C++ (csrc/host_order.cpp)
        self.sym = sym
        self.perm_v = sym.perm                                                          # new -> old
        tri_new = sym.iperm[tri_old]                                                    # (T,3) in new vertex ids
        self.perm_f = _stable_argsort(tri_new.min(axis=1), V)                           # new -> old triangle
        tri_new = tri_new[self.perm_f]
        tm["ordering"] = time.perf_counter() - t0
        # ---- which sweep kernels, how many padded modes ---------------------------------------------------
        # 4: ring-streamed sweeps (csrc/sweep_ring.cu): whole warps of modes, so a single rank pads small time grids up to 32
        #    modes (identity padding).  0: register-staged k_sweep_run: any mode count; kept for mode-sharded ranks with fewer than
        #    32 modes and for factors of a few hundred MB, where every level launch is latency bound and its shorter
        #    dependent chain wins (knots_5-class, nT = 31: 0.25 ms / iteration against 0.34; profiles/README.md).
        # More than 128 time levels on one GPU: the modes are solved in n_groups groups of at most 128 (own factor panels per
        # group, shared work vectors), with the plain time transforms (dots_time_transform_plain) around each group's sweeps.
        self.n_groups = -(-(nT + 1) // 128) if self.comm.world == 1 else 1
        per_rank = -(-(nT + 1) // (self.comm.world * self.n_groups))          # modes per rank (per group)
        if sweep_mode is None:
            sweep_mode = int(os.environ.get("DOTS_SWEEP_MODE", "-1"))
        if sweep_mode == -1:
            n_pad32 = max(32, dd.pad_modes(per_rank))
            small = 2 * 8 * n_pad32 * sym.panel_entries / 2 ** 20 < float(os.environ.get("DOTS_RING_MIN_MB", 300))
            sweep_mode = 0 if (small or (self.comm.world > 1 and dd.pad_modes(per_rank) % 32)) else 4
        if sweep_mode not in (0, 4):
            raise capi.DotsError(f"sweep_mode={sweep_mode} unsupported (0: k_sweep_run, 4: ring-streamed)")
        if self.n_groups > 1:
            self.part = dd.Partition(nT, 0, 1, nT + 1, 0, nT + 1, max(dd.pad_modes(per_rank), 32 if sweep_mode == 4 else 0))
        else:
            self.part = dd.partition(nT, self.comm.rank, self.comm.world, min_pad=32 if sweep_mode == 4 else 0)
        part = self.part
        if sweep_mode == 4 and part.m_pad % 32:
            raise capi.DotsError(f"sweep_mode=4 needs a multiple of 32 time modes per rank, got {part.m_pad}")
        self.m_pad = part.m_pad
        t0 = time.perf_counter()
        Q, lam_t = time_basis(nT)
        self.Q, self.lam_t = Q, lam_t
        # Mode storage order.  One rank with an even, warp-aligned level count: even modes first, so that the transforms can
        # use the symmetry Q[n-1-t][k] = (-1)^k Q[t][k] of the DCT-II basis (k_time_sym: half the tensor work).
        n_lv = nT + 1
        self.tt_sym = bool(part.world == 1 and n_lv % 16 == 0 and part.m_pad == n_lv and os.environ.get("DOTS_TT_SYM", "1") == "1")
        self.mode_order = (np.concatenate([np.arange(0, n_lv, 2), np.arange(1, n_lv, 2)]) if self.tt_sym else np.arange(n_lv))
        Q_st, lam_st = Q[:, self.mode_order], lam_t[self.mode_order]            # column j of `hat` holds mode mode_order[j]
        shifts = (-lam_st + self.eps)[part.lvl_begin:part.lvl_end]   # (L + (lam - eps) M) = -(K + (|lam| + eps) M)   (laplacian_inverse_socp.py:37-38)
        self.sweep_mode = int(sweep_mode)
        self.factor_stats = {}

        def factorise(shifts_g):
            if os.environ.get("DOTS_FACTOR", "hybrid") == "library":
                return nested.factor_batched_device(sym, K, area_v, shifts_g, m_pad=self.m_pad, device=self.device, transposed=True)
            return nested.factor_hybrid_device(sym, K, area_v, shifts_g, self.m_pad, self.device, self.lib,
                                               lambda: torch.cuda.current_stream(self.device).cuda_stream,
                                               stats=self.factor_stats, use_library=os.environ.get("DOTS_FACTOR") == "mixed")

        self.group_modes = [(g * per_rank, min(nT + 1, (g + 1) * per_rank)) for g in range(self.n_groups)] if self.n_groups > 1 else []
        if self.n_groups > 1:
            group_panels = [factorise(shifts[k0:k1]) for k0, k1 in self.group_modes]
            panels, panels_t = group_panels[0]
        else:
            panels, panels_t = factorise(shifts)
        t_fact0, t_fact_enqueued = t0, time.perf_counter()    # the last (largest) fronts are still being factorised on the GPU

        # ---- launch plans and index maps (host; overlaps the tail of the factorisation), then upload ---------
        t0 = time.perf_counter()
        dev = self.device
        self._keep = {}

        def up(name, arr, dtype):
            ten = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(dev)
            self._keep[name] = ten
            return ten

        props = torch.cuda.get_device_properties(dev)
        self.n_sm = props.multi_processor_count
        area_f_n, area_v_n = area_f[self.perm_f], area_v[self.perm_v]
        self.area_f_new, self.area_v_new = area_f_n, area_v_n
        self._host_mesh = dict(tri=tri_new, hat=hat[self.perm_f])                        # internal numbering, (T,3) / (T,3,3)
        hat_n = hat[self.perm_f]                                                         # (T,3,3)
        diag = np.sqrt(area_f_n[None, :] / area_v_n[tri_new.T])                          # (3,T)  solver_socp.py:172-180
        vc_ptr, vc_idx = np.empty(V + 1, dtype=np.int32), np.empty(3 * T, dtype=np.int32)
        tri_c = np.ascontiguousarray(tri_new, dtype=np.int64)
        capi.check(self.lib.dots_corner_lists(V, T, tri_c.ctypes.data, vc_ptr.ctypes.data, vc_idx.ctypes.data), "dots_corner_lists")
        deg = np.diff(vc_ptr)
        vc_ell = np.full((V, 8), -1, dtype=np.int32)                                      # ELL form of the same lists (iter_kernels.cu: vc_load)
        slot = np.arange(3 * T) - np.repeat(vc_ptr[:-1], deg)
        keep = slot < 8
        vc_ell[np.repeat(np.arange(V), deg)[keep], slot[keep]] = vc_idx[keep]
        vc_ell[deg > 8, 7] = -2                                                          # long lists: the kernels walk the CSR form
        if self.n_groups > 1:                                                            # the grouped path passes its bases per call
            qf, qb, n_phi_out = np.zeros((4, 8)), np.zeros((8, 8)), nT + 1
        else:
            qf, qb, n_phi_out = dd.transform_matrices(Q_st, part)
        self.sweep_grid = 0
        plan = _sweep_items(sym, self.n_sm, self.m_pad)
        self.plan = plan                                                                 # host arrays stay alive
        env_kb = lambda name, default: int(float(os.environ.get(name, default)) * 1024)
        self.ring = None
        if self.sweep_mode == 4:
            self.ring = ring_plan.build(
                sym=sym, n_sm=self.n_sm, m_pad=self.m_pad, split_bytes=env_kb("DOTS_RING_SPLIT_KB", 64),
                tasks_per_sm=int(os.environ.get("DOTS_RING_TASKS_PER_SM", 64)),
                task_bytes=(env_kb("DOTS_RING_TASK_MIN_KB", 16), env_kb("DOTS_RING_TASK_MAX_KB", 48)),
                wpr_max=int(os.environ.get("DOTS_RING_WPR_MAX", 8)))
            self.ring["erow_fwd"], self.ring["erow_bwd"] = ring_plan.entry_rows(self.lib, sym, self.ring["bidx"])
        fwd_ptr, fwd_items, bwd_ptr, bwd_items = plan["fwd_ptr"], plan["fwd_items"], plan["bwd_ptr"], plan["bwd_items"]
        self._h_fwd_ptr, self._h_bwd_ptr = fwd_ptr, bwd_ptr
        tm["plan"] = time.perf_counter() - t0            # host-side launch plans and index maps (analysis, like the ordering)
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        tm["factorization"] = (t_fact_enqueued - t_fact0) + (t0 - t_fact_enqueued - tm["plan"])   # enqueue + wait after the plans

        ctx = capi.DotsCtx()
        ctx.abi_version, ctx.n_time, ctx.n_vert, ctx.n_tri = capi.ABI_VERSION, nT, V, T
        ctx.m_pad, ctx.n_nodes, ctx.n_levels, ctx.n_sm = self.m_pad, sym.n_nodes, sym.n_levels, self.n_sm
        ctx.lvl_begin, ctx.lvl_end, ctx.n_ranks = part.lvl_begin, part.lvl_end, part.world
        ctx.tt_kf, ctx.tt_kb, ctx.tt_nb, ctx.tt_nout = qf.shape[0], qb.shape[0], qb.shape[1], n_phi_out
        ctx.tt_sym = int(self.tt_sym)
        const = dict(
            tri=up("tri", tri_new.T, np.int32), hat_grad=up("hat_grad", hat_n.transpose(1, 2, 0), np.float64),
            area_f=up("area_f", area_f_n, np.float64), area_v=up("area_v", area_v_n, np.float64),
            diag_soc=up("diag_soc", diag, np.float64), vc_ptr=up("vc_ptr", vc_ptr, np.int32),
            vc_idx=up("vc_idx", vc_idx, np.int32), vc_ell=up("vc_ell", vc_ell, np.int32), qf=up("qf", qf, np.float64), qb=up("qb", qb, np.float64),
            panels=panels, panels_t=panels_t,
            nd_off=up("nd_off", sym.off, np.int32), nd_s=up("nd_s", sym.s, np.int32), nd_b=up("nd_b", sym.b, np.int32),
            nd_child=up("nd_child", sym.child, np.int32), nd_panel=up("nd_panel", sym.panel_off[:-1], np.int64),
            nd_front=up("nd_front", sym.front_off[:-1], np.int64), nd_upd=up("nd_upd", sym.upd_off[:-1], np.int64),
            front_idx=up("front_idx", sym.front_idx, np.int32), child_pos=up("child_pos", sym.child_pos, np.int32),
            lvl_ptr=up("lvl_ptr", fwd_ptr, np.int32), lvl_items=up("lvl_items", fwd_items, np.int32),
            lvb_ptr=up("lvb_ptr", bwd_ptr, np.int32), lvb_items=up("lvb_items", bwd_items, np.int32),
            lvn_nodes=up("lvn_nodes", plan["nodes"], np.int32))
        self._keep["panels"], self._keep["panels_t"] = panels, panels_t
        for k, ten in const.items():
            setattr(ctx, k, ten.data_ptr())
        ctx.h_lvl_ptr = self._h_fwd_ptr.ctypes.data
        ctx.h_lvb_ptr = self._h_bwd_ptr.ctypes.data
        ctx.h_lvn_ptr = plan["node_ptr"].ctypes.data
        ctx.h_lvl_wpr = plan["wpr"].ctypes.data
        ctx.h_lvb_cw = plan["cw"].ctypes.data
        ctx.front_total = int(sym.front_off[-1])
        ctx.sweep_mode, ctx.sweep_grid = self.sweep_mode, self.sweep_grid
        if self.ring is not None:
            rp = self.ring
            for name in ("rt_fwd", "rt_bwd"):
                ten = torch.from_numpy(rp[name].view(np.uint8).reshape(-1).copy()).to(dev)
                self._keep[name] = ten
                setattr(ctx, name, ten.data_ptr())
            for name in ("bidx", "gptr", "gidx", "gverts"):
                setattr(ctx, name, up("ring_" + name, rp[name], np.int32).data_ptr())
            for name in ("erow_fwd", "erow_bwd"):
                setattr(ctx, name, up("ring_" + name, rp[name], np.int32).data_ptr())
            ctx.h_rt_fwd_ptr, ctx.h_rt_bwd_ptr = rp["fwd_ptr"].ctypes.data, rp["bwd_ptr"].ctypes.data
            ctx.h_rt_fwd_wpr, ctx.h_rt_bwd_wpr = rp["fwd_wpr"].ctypes.data, rp["bwd_wpr"].ctypes.data
            ctx.h_gv_ptr = rp["gv_ptr"].ctypes.data
            ctx.ring_stages = int(os.environ.get("DOTS_RING_STAGES", 2))      # measured: 2 x 4 KB stages, 3 blocks / SM
            ctx.ring_stage_bytes = int(os.environ.get("DOTS_RING_STAGE_BYTES", 4096))
        ctx.ring_pdl = int(os.environ.get("DOTS_RING_PDL", 1))                # both sweep kernels chain their level launches
        ctx.ring_flags = int(os.environ.get("DOTS_RING_L2HINT", 1))           # bit 0: the panel stream is marked evict_first in L2 (-2..4 % sweep time)
        self._keep["phase_clock"] = torch.zeros(2 * sym.n_levels + 1, dtype=torch.int64, device=dev)
        if os.environ.get("DOTS_PHASE_CLOCK"):
            ctx.phase_clock = self._keep["phase_clock"].data_ptr()

        torch.cuda.synchronize(dev)
        tm["upload_constants"] = time.perf_counter() - t0
        # ---- state: level-indexed slabs (+ halo levels), addressed through virtual bases -------------------
        l0, l1, te = part.lvl_begin, part.lvl_end, part.t_end
        S = lambda lo, hi, *row: dd.SlabStore(lo, hi, row, dev)
        self.slab = dict(
            phi=S(l0, l1 + 1, V), lam=S(l0 - 1, te, V), A=S(l0 - 1, te, V), lam_c=S(l0 - 1, te, V), mu=S(l0 - 1, te, V),
            z_fst=S(l0, te, V), z_end=S(l0, te, V), b_fst=S(l0, te, V), b_end=S(l0, te, V),
            B=S(l0, l1 + 1, 3, T), E=S(l0, l1, 3, T), b_mid=S(l0, l1, 2, 3, 3, T), z_mid=S(l0, l1, 2, 3, 3, T),
            corner_nrm=S(l0, l1 + 1, 2, 3, T), corner_div=S(l0, l1, 3, T))
        for name, st_ in self.slab.items():
            setattr(ctx, name, st_.base_ptr)
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)
        self.red_blocks = self.n_sm * 4
        z_all = z(2 * V, self.m_pad)                     # Z = [hat | ywork]: one allocation, the ring sweeps index both halves
        hat_local = z_all[:V]
        self._keep["z_all"] = z_all
        self.t = dict(params=z(capi.P_COUNT), bnd0=z(V), bnd1=z(V), rhs=z(part.world * part.chunk, V),
                      hat=hat_local, hat_all=hat_local if part.world == 1 else z(part.world, V, self.m_pad),
                      ywork=z_all[V:], upd=z(max(1, int(sym.upd_off[-1])), self.m_pad),
                      red_part=z(self.red_blocks, 72), red_out=z(72))
        for k, ten in self.t.items():
            setattr(ctx, k, ten.data_ptr())
        ctx.red_blocks = self.red_blocks
        # per-block partials of the triangle term of KKT #1 (iterate(kkt1=True)): one per block of k_tri_tma's grid, whose
        # time chunks hold at least 2 levels; the plain-load triangle kernel (DOTS_TRI_PLAIN=1, diagnostics) has no such mode
        if os.environ.get("DOTS_TRI_PLAIN", "0") == "1":
            ctx.ring_flags |= 4                          # diagnostics: plain-load triangle kernel instead of the TMA-staged one
        self.can_fuse_kkt1 = not (ctx.ring_flags & 4)
        n_k1 = -(-T // 128) * -(-(l1 - l0) // 2)
        self.t["kkt1_part"] = z(max(1, n_k1))
        ctx.kkt1_part, ctx.kkt1_blocks = self.t["kkt1_part"].data_ptr(), n_k1
        self.kkt1_valid = False
        self.ctx = ctx
        self._ctxp = C.byref(ctx)
        self._host_out = np.zeros(72)
        self._sum_cache = {}                             # condition -> raw sums of the CURRENT state (prefetch_sums)
        self._host_params = np.zeros(capi.P_COUNT)
        self._halo_v = z(4, V)
        self._halo_c = z(3, T)

        # ---- scalars of the reference driver (:97, :267-270, :296-313, :318-321) ------------------
        self.r, self.s, self.d = 1.0, 1.0, 1.0
        self.ps, self.ds = 1.0, 1.0                      # prim_scale, dual_scale (:318-319; change only under is_constant_scaling)
        self.cong0 = self.cong                           # the caller's congestion (self.cong is rescaled with the variables)
        mu0 = np.asarray(geometry["mu0"], dtype=np.float64)[self.perm_v]
        mu1 = np.asarray(geometry["mu1"], dtype=np.float64)[self.perm_v]
        b0, b1 = -mu0 / (self.r * self.dt), mu1 / (self.r * self.dt)
        self.t["bnd0"].copy_(torch.from_numpy(b0))
        self.t["bnd1"].copy_(torch.from_numpy(b1))
        self._bnd_init = (self.t["bnd0"].clone(), self.t["bnd1"].clone())
        self.norm_bnd = self.r * self.dt * math.sqrt((np.sum((b0 / area_v_n) ** 2 * area_v_n)
                                                      + np.sum((b1 / area_v_n) ** 2 * area_v_n)) / (nT + 1))
        self._norm_bnd_init = self.norm_bnd
        self._bnd_host = (b0, b1)
        self.area_mesh = float(np.sum(area_f))
        self.norm_d = math.sqrt(2 * self.area_mesh)
        ma_v, ma_f = float(np.mean(area_v)), float(np.mean(area_f))
        self.k_prim_q = np.mean([ma_v, ma_f])
        self.k_prim_z = np.mean([ma_v, ma_f, ma_v])
        self.k_dual_a = ma_v
        self.k_dual_b = np.mean([ma_v, ma_f])
        self.k_comp_rho, self.k_comp_m = ma_v, ma_f
        self.z_valid = True                              # z_mid = 0 is the reference's initial z_mid (:245)
        self.use_graphs, self._warm, self._graphs, self._cap_stream = part.world == 1, False, {}, None
        # capturing NCCL collectives in a CUDA graph gained only ~4 % at N=2 and hung at process-group teardown: opt-in
        self.use_sharded_graphs = os.environ.get("DOTS_SHARDED_GRAPHS", "0") == "1"
        self._tgraphs, self._tgraph_warm, self.graph_error = {}, {}, None
        self.launches = 0
        self.peers = False
        self.peer_error = None
        if self.comm.enabled and os.environ.get("DOTS_PEER", "1") == "1" and part.world <= 8:
            self._setup_peers()
        self.groups = []                                 # (context with the group's panels, qf [levels][m_pad], qb [m_pad][levels])
        for g, (k0, k1) in enumerate(self.group_modes):
            cg = capi.DotsCtx()
            C.memmove(C.byref(cg), C.byref(ctx), C.sizeof(ctx))
            pg, pgt = group_panels[g]
            self._keep[f"panels_g{g}"], self._keep[f"panels_t_g{g}"] = pg, pgt
            cg.panels, cg.panels_t = pg.data_ptr(), pgt.data_ptr()
            qf_g = np.zeros((nT + 1, self.m_pad))
            qf_g[:, :k1 - k0] = Q[:, k0:k1]
            self.groups.append((cg, up(f"qf_g{g}", qf_g, np.float64), up(f"qb_g{g}", qf_g.T, np.float64)))
        if self.groups:
            self.use_graphs = False                      # eager launches: the group loop lives on the host
        self._push_params()
        torch.cuda.synchronize(dev)
        tm["upload"] = time.perf_counter() - t0
        tm["setup_total"] = time.perf_counter() - t_start
        self.timings = tm

    # ------------------------------------------------------------------ plumbing
    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _push_params(self):
        p = self._host_params
        p[capi.P_R], p[capi.P_S], p[capi.P_D] = self.r, self.s, self.d
        p[capi.P_CONG], p[capi.P_TAU], p[capi.P_EPS] = self.cong, self.tau, self.eps
        p[capi.P_PS], p[capi.P_DS] = self.ps, self.ds
        capi.check(self.lib.dots_set_params(self._ctxp, p.ctypes.data, self.stream), "dots_set_params")

    def launches_per_iteration(self):
        """Kernel launches of one iteration: rhs, 2 transforms, the sweep launches, vertex, triangle."""
        if self.sweep_mode == 4:
            sweeps = ring_plan.launches(self.ring)
            return 3 + max(1, self.n_groups) * (2 + sweeps)
        n_f = int(np.count_nonzero(np.diff(self._h_fwd_ptr)))
        n_b = int(np.count_nonzero(np.diff(self._h_bwd_ptr)))
        n_g = int(np.count_nonzero(np.diff(self.plan["node_ptr"])[1:]))
        return 3 + max(1, self.n_groups) * (2 + n_f + n_b + n_g)

    def _call(self, fn, *args):
        capi.check(getattr(self.lib, fn)(self._ctxp, *args, self.stream), fn)

    # ------------------------------------------------------------------ peer memory (NVLink stores instead of NCCL copies)
    def _setup_peers(self):
        """Map the neighbours' halo rows and every rank's rhs buffer into this process (CUDA IPC through torch's tensor
        sharing), so that k_vertex / k_tri / k_phi_rhs store their boundary data straight into the consumers' memory.
        Any failure leaves the NCCL exchanges in place."""
        try:
            part, comm, ctx = self.part, self.comm, self.ctx
            torch.cuda.set_device(self.device)

            def export(ten):
                handle, off = C.create_string_buffer(64), C.c_ulonglong(0)
                capi.check(self.lib.dots_ipc_export(ten.data_ptr(), handle, C.byref(off)), "dots_ipc_export")
                return bytes(handle.raw), int(off.value), self.device.index

            mine = {n: export(self.slab[n].data) for n in ("lam", "A", "lam_c", "mu", "corner_nrm")}
            mine["rhs"] = export(self.t["rhs"])
            mine["hat"] = export(self.t["hat"])
            everyone = [None] * part.world
            comm.dist.all_gather_object(everyone, mine, group=comm.group)
            self._ipc_cache = {}

            def open_(rank, name):
                handle, off, dev_index = everyone[rank][name]
                if handle not in self._ipc_cache:               # one mapping per exporting allocation
                    capi.check(self.lib.dots_enable_peer(int(dev_index)), "dots_enable_peer")
                    base = C.c_void_p()
                    capi.check(self.lib.dots_ipc_import(handle, 0, C.byref(base)), "dots_ipc_import")
                    self._ipc_cache[handle] = base.value
                return self._ipc_cache[handle] + off

            V, T = self.V, self.T
            if part.rank + 1 < part.world:
                for i, n in enumerate(("lam", "A", "lam_c", "mu")):
                    ctx.peer_vertex[i] = open_(part.rank + 1, n)                      # row 0 = its step lvl_begin-1
            if part.rank > 0:
                prev = dd.partition(self.nT, part.rank - 1, part.world)
                halo_elems = (prev.n_levels * 2 + 1) * 3 * T                          # its halo level lvl_end, side 1
                ctx.peer_corner = open_(part.rank - 1, "corner_nrm") + halo_elems * 8
            for r in range(part.world):
                ctx.peer_rhs[r] = self.t["rhs"].data_ptr() if r == part.rank else open_(r, "rhs")
                ctx.peer_hat[r] = self.t["hat"].data_ptr() if r == part.rank else open_(r, "hat")
            self._fence_t = torch.zeros(1, dtype=torch.float64, device=self.device)
            ok = torch.ones(1, dtype=torch.float64, device=self.device)
            comm.dist.all_reduce(ok, op=comm.dist.ReduceOp.MIN, group=comm.group)
            self.peers = bool(ok.item() == 1.0)
        except Exception as exc:                                   # pragma: no cover - depends on the node's IPC support
            self.peer_error = repr(exc)
            self.peers = False
            try:
                bad = torch.zeros(1, dtype=torch.float64, device=self.device)
                self.comm.dist.all_reduce(bad, op=self.comm.dist.ReduceOp.MIN, group=self.comm.group)
            except Exception:
                pass
        if not self.peers:
            ctx = self.ctx
            for i in range(4):
                ctx.peer_vertex[i] = None
            ctx.peer_corner = None
            for r in range(8):
                ctx.peer_rhs[r] = None
                ctx.peer_hat[r] = None

    def fence(self):
        """Cross-rank, stream-ordered fence: a one-element all-reduce completes on a rank only after every rank's
        preceding kernels (and their peer stores) have finished."""
        self.comm.dist.all_reduce(self._fence_t, group=self.comm.group)

    # ------------------------------------------------------------------ halo exchanges (no-ops on one rank)
    def exchange_vertex_halo(self, pushed=False):
        """(lam, A, lam_c, mu) of the last owned step -> step lvl_begin-1 of the next rank.  ``pushed``: the producing
        kernel already stored them through peer memory, only the fence is needed."""
        if not self.comm.enabled:
            return
        if pushed and self.peers:
            return self.fence()
        part, sl = self.part, self.slab
        send = None
        if part.rank + 1 < part.world and part.n_steps > 0:
            send = torch.stack([sl[n].level(part.t_end - 1) for n in ("lam", "A", "lam_c", "mu")])
        recv = self._halo_v if part.rank > 0 else None
        self.comm.shift(send_next=send, recv_prev=recv)
        if recv is not None:
            for i, n in enumerate(("lam", "A", "lam_c", "mu")):
                sl[n].level(part.lvl_begin - 1).copy_(recv[i])

    def exchange_corner_halo(self, fence=True):
        """side-1 corner norms of the first owned level -> level lvl_end of the previous rank (k_tri stores them through
        peer memory when that is set up; then only a fence is left, and inside the iteration not even that: the fence
        after the next k_phi_rhs orders them before their only reader, k_vertex)."""
        if not self.comm.enabled:
            return
        if self.peers:
            return self.fence() if fence else None
        part, sl = self.part, self.slab
        send = sl["corner_nrm"].level(part.lvl_begin)[1].contiguous() if part.rank > 0 else None
        recv = self._halo_c if part.rank + 1 < part.world else None
        self.comm.shift(send_prev=send, recv_next=recv)
        if recv is not None:
            sl["corner_nrm"].level(part.lvl_end)[1].copy_(recv)

    def exchange_B_halo(self):
        """B of the first owned level -> level lvl_end of the previous rank (only KKT #4 reads it)."""
        if not self.comm.enabled:
            return
        part, sl = self.part, self.slab
        send = sl["B"].level(part.lvl_begin).contiguous() if part.rank > 0 else None
        recv = self._halo_c if part.rank + 1 < part.world else None
        self.comm.shift(send_prev=send, recv_next=recv)
        if recv is not None:
            sl["B"].level(part.lvl_end).copy_(recv)

    # ------------------------------------------------------------------ the iteration
    def iterate(self, n=1, write_z=False, kkt1=False):
        """n ALM iterations (Steps 1-3); ``write_z`` stores z_mid on the last one; ``kkt1`` (without ``write_z``) makes the last
        one accumulate the triangle term of KKT #1 instead (dots_step_tri mode 2): what a check iteration needs when z_mid
        itself will not be returned (3 GB less to write, 3 GB less to read back at the headline size).  On one GPU the
        iteration is replayed from a captured CUDA graph after the first call (one per flavour)."""
        n, write_z = int(n), bool(write_z)
        if kkt1 and not write_z and not self.can_fuse_kkt1:
            write_z = True                               # no accumulating mode in the plain-load kernel: store z_mid instead
        mode = 1 if write_z else (2 if kkt1 else 0)
        st = self.stream
        if self.comm.enabled:
            for i in range(n):
                self._iterate_sharded_graphed(mode if i == n - 1 else 0)
        elif self.groups:
            for i in range(n):
                self.step_phi()
                self._call("dots_step_vertex")
                self._call("dots_step_tri", mode if i == n - 1 else 0)
        elif not self.use_graphs or not self._warm:
            capi.check(self.lib.dots_iterate(self._ctxp, n, mode, st), "dots_iterate")
            self._warm = True
        else:
            for i in range(n):
                wz = mode if i == n - 1 else 0
                key = (int(wz), st)
                if key not in self._graphs:
                    # capture on a private stream (the legacy default stream cannot be captured); capturing does
                    # not execute anything, and the instantiated graph is launched on the caller's stream
                    if self._cap_stream is None:
                        self._cap_stream = torch.cuda.Stream(self.device)
                    h = C.c_void_p()
                    rc = self.lib.dots_graph_create(self._ctxp, int(wz), self._cap_stream.cuda_stream, C.byref(h))
                    if rc != 0:                      # capture refused (e.g. a launch kind the driver cannot capture): stay eager
                        msg = self.lib.dots_last_error()
                        self.graph_error = msg.decode() if msg else f"dots_graph_create -> {rc}"
                        self.use_graphs = False
                        torch.cuda.synchronize(self.device)
                        capi.check(self.lib.dots_iterate(self._ctxp, n - i, mode, st), "dots_iterate")
                        break
                    self._graphs[key] = h
                capi.check(self.lib.dots_graph_launch(self._graphs[key], st), "dots_graph_launch")
        self.launches += n * self.launches_per_iteration()
        self.z_valid = write_z
        self._state_changed()
        self.kkt1_valid = mode == 2

    def step_phi(self):
        """Step 1 alone (single GPU): rhs, time transform, mode solves, inverse transform; with more than 128 time levels the
        modes go through in groups (own panels, shared work vectors, phi accumulated group by group in a fixed order)."""
        if not self.groups:
            return self._call("dots_step_phi")
        st = self.stream
        self._call("dots_phi_rhs")
        for g, (cg, qf_g, qb_g) in enumerate(self.groups):
            capi.check(self.lib.dots_time_transform_plain(C.byref(cg), 0, qf_g.data_ptr(), 0, st), "dots_time_transform_plain")
            capi.check(self.lib.dots_mode_solves(C.byref(cg), st), "dots_mode_solves")
            capi.check(self.lib.dots_time_transform_plain(C.byref(cg), 1, qb_g.data_ptr(), int(g > 0), st), "dots_time_transform_plain")

    def _iterate_sharded(self, write_z):
        """One iteration across ranks: the same kernels on this rank's slab / modes + the exchanges of dist.py."""
        part, comm, t = self.part, self.comm, self.t
        self._call("dots_phi_rhs")                                                     # own levels of rhs
        if self.peers:
            self.fence()                                                               # slabs were stored into every rank's rhs
        else:
            comm.all_gather_into(t["rhs"], t["rhs"][part.rank * part.chunk:(part.rank + 1) * part.chunk])
        self._call("dots_time_transform", 0)                                           # all levels -> own modes
        self._call("dots_mode_solves")
        if self.peers:
            self.fence()                                                               # the inverse transform reads the peers' `hat`
        else:
            comm.all_gather_into(t["hat_all"], t["hat"])
        self._call("dots_time_transform", 1)                                           # all modes -> own levels (+halo)
        self._call("dots_step_vertex")
        self.exchange_vertex_halo(pushed=True)
        self._call("dots_step_tri", int(write_z))
        self.exchange_corner_halo(fence=False)

    def _iterate_sharded_graphed(self, write_z):
        """Replay the sharded iteration (kernels + NCCL collectives) from a torch CUDA graph; the first two calls of each
        flavour run eagerly (NCCL connection set-up, first-use kernel attributes), a failed capture falls back to eager."""
        key = int(write_z)
        if not self.use_sharded_graphs or self._tgraph_warm.get(key, 0) < 2:
            self._tgraph_warm[key] = self._tgraph_warm.get(key, 0) + 1
            return self._iterate_sharded(write_z)
        if key not in self._tgraphs:
            try:
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._iterate_sharded(write_z)
                self._tgraphs[key] = g
            except Exception as exc:                     # capture unsupported for some op: stay eager
                self.use_sharded_graphs = False
                self.graph_error = repr(exc)
                torch.cuda.synchronize(self.device)
                return self._iterate_sharded(write_z)
        self._tgraphs[key].replay()

    def close(self):
        for h in self._graphs.values():
            self.lib.dots_graph_destroy(h)
        self._graphs = {}
        self._tgraphs = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step_q0(self):
        """``is_palm=True`` only (solver_socp.py:668-672): the q / lambda solve that opens an iteration, from the gradients
        of the current phi and the current z, beta, mu, E: two fused kernels (``dots_step_q0``: vertex half -> A, lam_c;
        triangle half -> B and the corner terms of the new B).  Needs z_mid of the previous iteration (write_z=True)."""
        if self.comm.enabled:
            raise NotImplementedError("is_palm is not available on the sharded (multi-GPU) path")
        if not self.z_valid:
            raise capi.DotsError("is_palm needs z_mid of the previous iteration: call iterate(..., write_z=True)")
        self._state_changed()
        capi.check(self.lib.dots_step_q0(self._ctxp, self.stream), "dots_step_q0")
        self.launches += 2

    def adjust_penalty(self, f):                                                         # :367-371
        self._state_changed(keep_kkt1=True)              # KKT #1 involves z, B, A and s only: untouched by the dual rescaling
        self.r *= f
        self._push_params()
        capi.check(self.lib.dots_scale_dual(self._ctxp, float(f), self.stream), "dots_scale_dual")
        self.exchange_corner_halo()
        self.launches += 8

    def scale_z(self, f):                                                                # :373-395
        self._state_changed()
        self.s *= f
        self.d *= f
        self.norm_d *= f
        capi.check(self.lib.dots_scale_z(self._ctxp, float(self.s), self.stream), "dots_scale_z")
        self._push_params()
        self.exchange_vertex_halo()
        self.refresh()
        self.launches += 8

    # ------------------------------------------------------------------ is_constant_scaling (solver_socp.py:324-365, :574-587)
    def variable_norms(self):
        """The five norms scale_prim_dual compares (:329-340): ([prim x3], [dual x2]) of the scaled variables, from ONE fused
        reduction pass (conditions 0 and 3 carry dt_phi, dx_phi, A, B, mu, E; slot 8 the z / beta norms).  Needs z_mid of the
        current iteration."""
        if not self.z_valid:
            raise capi.DotsError("scale_prim_dual needs z_mid of the current iteration: call iterate(..., write_z=True)")
        self.prefetch_sums((0, 3, 8))
        c0, c3, c8 = self._sum_cache[0], self._sum_cache[3], self._sum_cache[8]
        tn, sn, sq = 1.0 / self.nT, 1.0 / (self.nT + 1), math.sqrt
        prim = [sq(c0[1] * tn + c0[5] * sn), sq(c0[2] * tn + c0[6] * sn), sq(c8[0] * tn + c8[4] * tn + c8[1] * tn)]
        dual = [self.r * sq(c3[0] * tn + c3[4] * sn), self.r * sq(c8[2] * tn + c8[5] * tn + c8[3] * tn)]
        return prim, dual

    def scale_prim_dual(self, factors=None):
        """scale_prim_dual (:324-365): rescale primal variables by 1/p and dual ones by p/d^2 when the two factors differ by
        more than 2x; ``factors`` None = from the current variable norms (admm_tools.compute_scale_factor)."""
        if factors is None:
            prim, dual = self.variable_norms()
            p, d = max(prim) / 1.0, max(dual) / 1.0
        else:
            p, d = float(factors[0]), float(factors[1])
        if not max(p, d) / min(p, d) > 2.0:                                              # :344
            return False
        self._state_changed()
        self.ps *= p
        self.ds *= d
        capi.check(self.lib.dots_scale_prim_dual(self._ctxp, p, d ** 2 / p, self.stream), "dots_scale_prim_dual")
        self.r *= d / p
        self.cong *= d / p
        self.d /= p
        self.norm_d /= p
        self.norm_bnd /= d
        self._push_params()
        self.exchange_corner_halo()
        self.launches += 15
        return True

    def initial_constant_scaling(self):
        """The initial scaling of :574-587 (after the initial z scaling): factors from the boundary data, then r -> 1."""
        nT, dt, sq = self.nT, self.dt, math.sqrt
        av, af = self.area_v_new, self.area_f_new
        tri, hat = self._host_mesh["tri"], self._host_mesh["hat"]
        b0, b1 = (self.r * self._bnd_host[0]) / av, (self.r * self._bnd_host[1]) / av     # _boundary_time rows 0 and nT
        norm_c = sq((np.sum(b0 ** 2 * av) + np.sum(b1 ** 2 * av)) / (nT + 1))
        if nT == 1:
            gt = np.sum(((b1 - b0) / dt) ** 2 * av)
        else:
            gt = np.sum((b0 / dt) ** 2 * av) + np.sum((b1 / dt) ** 2 * av)                 # rows 0 and nT-1 of grad_time
        gs = 0.0
        for bt in (b0, b1):
            g = np.einsum("fkx,fk->fx", hat, bt[tri])                                     # vanilla_grad_space (:898-907)
            gs += np.sum(g ** 2 * af[:, None])
        norm_ac = sq(gt / nT + gs / (nT + 1))
        self.scale_prim_dual((self.norm_d, sq(nT) * norm_c ** 2 / norm_ac))
        self.adjust_penalty(1.0 / self.r)

    def reset_state(self):
        """Back to the state solver_socp starts from (socp/solver_socp.py:239-270: all arrays zero, r = 1, unscaled z)."""
        for st_ in self.slab.values():
            st_.data.zero_()
        self.r, self.s, self.d = 1.0, 1.0, 1.0
        self.ps, self.ds, self.cong = 1.0, 1.0, self.cong0
        self.norm_d = math.sqrt(2 * self.area_mesh)
        self.norm_bnd = self._norm_bnd_init
        self.t["bnd0"].copy_(self._bnd_init[0])
        self.t["bnd1"].copy_(self._bnd_init[1])
        self.z_valid = True
        self._push_params()
        self.refresh()

    def phi_residual(self):
        """Residual of the space-time operator for the CURRENT phi against the CURRENT rhs, per time mode:
        ``max |Q^T (div_t(area_v grad_t phi) + D(area_f G phi) - rhs)| / max |Q^T rhs|`` (eps = 0; the operator the reference
        inverts in utils/laplacian_inverse_socp.py:52-61), formed with the stand-alone operator kernels dots_grad_space /
        dots_div_space (rows a5, a6) and whole-array device arithmetic.  Single GPU.  Returns a numpy array (nT + 1)."""
        if self.comm.enabled:
            raise NotImplementedError("phi_residual is a single-GPU diagnostic")
        nT, V, T, dev = self.nT, self.V, self.T, self.device
        phi = self.slab["phi"].levels(0, nT + 1)
        g = torch.empty((nT + 1, 3, T), dtype=torch.float64, device=dev)
        capi.check(self.lib.dots_grad_space(self._ctxp, self.slab["phi"].base_ptr, g.data_ptr(), self.stream), "dots_grad_space")
        g *= self._keep["area_f"][None, None, :]
        lap = torch.empty((nT + 1, V), dtype=torch.float64, device=dev)
        capi.check(self.lib.dots_div_space(self._ctxp, g.data_ptr(), lap.data_ptr(), self.stream), "dots_div_space")
        m = (phi[1:] - phi[:-1]) / self.dt * self._keep["area_v"][None, :]
        lap[0] += m[0] / self.dt
        lap[1:-1] += (m[1:] - m[:-1]) / self.dt
        lap[-1] -= m[-1] / self.dt
        q = torch.as_tensor(self.Q, device=dev)
        rhs = self.t["rhs"][:nT + 1]
        res, den = q.T @ (lap - rhs), q.T @ rhs
        self.launches += 2
        return (res.abs().amax(dim=1) / den.abs().amax(dim=1).clamp_min(1e-300)).cpu().numpy()

    def state_checksums(self):
        """A few order-independent sums of the iterate (for cross-run / cross-rank-count parity records)."""
        out = {}
        for name in ("mu", "A", "B", "E"):
            x = self.full(name)
            out[name] = [float(x.sum()), float((x * x).sum())]
        g = torch.diff(self.full("phi"), dim=0)
        out["dt_phi"] = [float(g.sum()), float((g * g).sum())]
        return out

    def set_scalars(self, r=None, s=None, d=None, norm_d=None):
        """Overwrite the driver scalars (tests / warm starts) and push them to the device."""
        self._state_changed()
        if r is not None: self.r = float(r)
        if s is not None: self.s = float(s)
        if d is not None: self.d = float(d)
        if norm_d is not None: self.norm_d = float(norm_d)
        self._push_params()

    def grad_space_into(self, src, dst):
        """slab[dst] = G slab[src] on the owned levels (vanilla_grad_space, solver_socp.py:898-907)."""
        capi.check(self.lib.dots_grad_space(self._ctxp, self.slab[src].base_ptr, self.slab[dst].base_ptr, self.stream), "dots_grad_space")
        self.launches += 1

    def E_from_beta(self, scale):
        """E = -decouple_adjoin_spacial(b_mid, scale) (warm-start default, solver_socp.py:250); rare, so plain tensor ops."""
        part = self.part
        bm = self.slab["b_mid"].levels(part.lvl_begin, part.lvl_end)
        self.slab["E"].levels(part.lvl_begin, part.lvl_end).copy_(-(scale / math.sqrt(3.0)) * bm.sum(dim=2).sum(dim=1))

    def refresh(self):
        self._state_changed()
        capi.check(self.lib.dots_refresh_corner_terms(self._ctxp, self.stream), "dots_refresh_corner_terms")
        self.exchange_corner_halo()
        self.launches += 1

    # ------------------------------------------------------------------ residuals
    def sums(self, which):
        which = int(which)
        if which in self._sum_cache:
            return self._sum_cache[which]
        if which == 1 and not self.z_valid and self.kkt1_valid:
            self.prefetch_sums([1])
            return self._sum_cache[1]
        capi.check(self.lib.dots_kkt_sums(self._ctxp, which, self._host_out.ctypes.data, self.stream), "dots_kkt_sums")
        self.launches += 3
        return self.comm.sum_in_rank_order(self._host_out[:8].copy(), self.device)

    def prefetch_sums(self, conditions):
        """Raw sums of several conditions (0..6, 7 = objective) in ONE fused pass per side and ONE host synchronisation
        (dots_kkt_sums_multi); ``kkt`` / ``objective`` then read them from the cache until the state changes.  Used where
        the reference is known to evaluate a whole set: the forced conditions of a penalty-update iteration
        (solver_socp.py:728-729) and the step-by-step mode (:769-787)."""
        conds = sorted({int(i) for i in conditions})
        if not conds:
            return
        if 1 in conds and not (self.z_valid or self.kkt1_valid):
            raise capi.DotsError("KKT #1 needs z_mid of the current iteration: call iterate(..., write_z=True) or iterate(..., kkt1=True)")
        if 4 in conds:
            self.exchange_B_halo()
        mask = sum(1 << i for i in conds)
        if 1 in conds and not self.z_valid:
            mask |= 0x200                                # triangle term of #1 from the partials the iteration accumulated
        capi.check(self.lib.dots_kkt_sums_multi(self._ctxp, mask, self._host_out.ctypes.data, self.stream), "dots_kkt_sums_multi")
        self.launches += 4
        total = self.comm.sum_in_rank_order(self._host_out.copy(), self.device)
        self._sum_cache.update({i: total[8 * i:8 * i + 8].copy() for i in conds})

    def _state_changed(self, keep_kkt1=False):
        self._sum_cache = {}
        if not keep_kkt1:
            self.kkt1_valid = False                      # the accumulated triangle term of KKT #1 belongs to the state it was formed from

    def kkt(self, i):
        """Relative KKT residual i as [value, value] (conditions 0-3) or [value, None] (4-6):
        the two-valued convention of solver_socp.py:589-643 with prim_scale = dual_scale = 1."""
        nT, sq = self.nT, math.sqrt
        if i == 1 and not (self.z_valid or self.kkt1_valid):
            raise capi.DotsError("KKT #1 needs z_mid of the current iteration: call iterate(..., write_z=True) or iterate(..., kkt1=True)")
        if i == 4 and i not in self._sum_cache:
            self.exchange_B_halo()
        o = self.sums(i)
        v, t = o[0:4], o[4:8]
        tn, sn = 1.0 / nT, 1.0 / (nT + 1)
        if i == 0:                                                                       # :433-450
            norm_sum = sq(v[1] * tn + t[1] * sn) + sq(v[2] * tn + t[2] * sn) + sq(v[3] * tn)
            res = sq(v[0] * tn + t[0] * sn)
            return [res / (self.k_prim_q / self.ps + norm_sum), res / (self.k_prim_q / 1.0 + norm_sum)]
        if i == 1:                                                                       # :452-464
            res = sq(v[0] * tn + v[1] * tn + t[0] * tn)
            return [res / (self.k_prim_z / self.ps + self.norm_d), res / (self.k_prim_z / 1.0 + self.norm_d)]
        if i == 2:                                                                       # :466-482
            res = sq(v[0] * sn)
            return [res / (self.k_dual_a / self.ds + self.norm_bnd), res / (self.k_dual_a / 1.0 + self.norm_bnd)]
        if i == 3:                                                                       # :484-503
            norm_sum = self.r * (sq(v[0] * tn + t[0] * sn) + sq(v[1] * tn + t[1] * sn))
            res = self.r * sq(v[2] * tn + t[2] * sn)
            return [res / (self.k_dual_b / self.ds + norm_sum), res / (self.k_dual_b / 1.0 + norm_sum)]
        if i == 4:                                                                       # :505-526
            return [sq(v[2] * tn) / (self.k_comp_rho + sq(v[0] * tn) + sq(v[1] * tn)), None]
        if i == 5:                                                                       # :528-547
            return [sq(t[2] * sn) / (self.k_comp_m + sq(t[0] * sn) + sq(t[1] * sn)), None]
        if i == 6:                                                                       # :549-559
            return [sq(v[2] * tn) / (self.k_comp_rho + sq(v[0] * tn) + sq(v[1] * tn)), None]
        raise IndexError(i)

    def objective(self):                                                                 # :417-431
        o = self.sums(7)
        cost = self.dt * (o[0] + o[1])
        cong = self.cong * self.ps / self.ds                                             # :830
        if cong > 10 ** (-10):
            return cost, cost - 1. / (2. * cong) * (o[2] / self.nT)
        return cost, cost

    # ------------------------------------------------------------------ layout conversion (host <-> device)
    def _perm_v_t(self):
        if "perm_v_t" not in self._keep:
            self._keep["perm_v_t"] = torch.from_numpy(self.perm_v.astype(np.int64)).to(self.device)
            self._keep["perm_f_t"] = torch.from_numpy(self.perm_f.astype(np.int64)).to(self.device)
        return self._keep["perm_v_t"], self._keep["perm_f_t"]

    def to_internal(self, name, ref):
        """Reference-layout array (numpy or torch, ALL time levels) -> internal-layout device tensor (all levels)."""
        pv, pf = self._perm_v_t()
        x = torch.as_tensor(ref, dtype=torch.float64).to(self.device)
        if name in STATE_VERTEX or name in ("lam", "rhs"):
            return x[:, pv].contiguous()
        if name in STATE_TRI:
            return x[:, pf, :].permute(0, 2, 1).contiguous()
        if name in STATE_CORNER:
            nT, T = self.nT, self.T
            out = torch.zeros((nT + 1, 2, 3, 3, T), dtype=torch.float64, device=self.device)
            out[:nT, 0] = x[:, 0][:, :, pf, :].permute(0, 1, 3, 2)
            out[1:, 1] = x[:, 1][:, :, pf, :].permute(0, 1, 3, 2)
            return out
        raise KeyError(name)

    def full(self, name, root=None):
        """All time levels of a state field in the internal layout (gathered from the ranks when sharded: on every rank, or
        on rank ``root`` only - the others get None)."""
        n_total = self.nT + 1 if name in ("phi", "rhs") + STATE_TRI + STATE_CORNER else self.nT
        if name == "rhs":
            return self.t["rhs"][:n_total]
        st_ = self.slab[name]
        hi = self.part.lvl_end if n_total == self.nT + 1 else self.part.t_end
        if not self.comm.enabled:
            return st_.levels(0, n_total)
        return dd.gather_levels(self.comm, self.part, st_, n_total, owned_hi=hi, root=root)

    def from_internal(self, name, ten=None, root=None):
        """Internal tensor with all time levels (default: the current state field) -> reference-layout device tensor
        (None on the ranks other than ``root`` when a root is given)."""
        pv, pf = self._perm_v_t()
        x = self.full(name, root=root) if ten is None else ten
        if x is None:
            return None
        if name in STATE_VERTEX or name in ("lam", "rhs"):
            out = torch.empty_like(x)
            out[:, pv] = x
            return out
        if name in STATE_TRI:
            out = torch.empty((x.shape[0], self.T, 3), dtype=torch.float64, device=self.device)
            out[:, pf, :] = x.permute(0, 2, 1)
            return out
        if name in STATE_CORNER:
            nT, T = self.nT, self.T
            out = torch.empty((nT, 2, 3, T, 3), dtype=torch.float64, device=self.device)
            out[:, 0][:, :, pf, :] = x[:nT, 0].permute(0, 1, 3, 2)
            out[:, 1][:, :, pf, :] = x[1:, 1].permute(0, 1, 3, 2)
            return out
        raise KeyError(name)

    def _store(self, name, internal_full):
        """Copy this rank's levels (incl. the halo levels it backs) of a full internal array into its slab."""
        st_ = self.slab[name]
        lo, hi = max(st_.lo, 0), min(st_.hi, internal_full.shape[0])
        if hi > lo:
            st_.levels(lo, hi).copy_(internal_full[lo:hi])

    def set_state(self, **arrays):
        """Overwrite state fields from reference-layout arrays (all levels), then rebuild the derived corner terms."""
        for name, ref in arrays.items():
            self._store(name, self.to_internal(name, ref))
        self.z_valid = "z_mid" in arrays
        self.refresh()

    def get_state(self, names=None):
        names = names or (STATE_VERTEX + STATE_TRI + STATE_CORNER)
        return {n: self.from_internal(n).cpu().numpy() for n in names}

    _STAGE_BYTES = 64 << 20
    _STAGE_BUFS = 3
    _COPY_THREADS = 4
    _stage_cache = {}                                  # device index -> pinned staging buffers, reused by every download
    _copy_pool = None

    def prepare_download(self, shapes=None):
        """Everything a later ``_download`` needs that does not depend on the result: the pinned staging buffers and the copy
        threads (once per process and device) and, for ``shapes`` = {name: shape}, the host arrays themselves, allocated and
        page-faulted by a background thread while the GPU iterates (first touch of 0.6 GB of fresh memory costs ~0.1 s)."""
        from concurrent.futures import ThreadPoolExecutor
        key = self.device.index
        if key not in Engine._stage_cache:
            Engine._stage_cache[key] = [torch.empty(Engine._STAGE_BYTES // 8, dtype=torch.float64, pin_memory=True)
                                        for _ in range(Engine._STAGE_BUFS)]
        if Engine._copy_pool is None:
            Engine._copy_pool = ThreadPoolExecutor(Engine._COPY_THREADS, thread_name_prefix="dots-d2h")
        if shapes:
            def make(shape):
                a = np.empty(shape, dtype=np.float64)
                a.reshape(-1)[::512] = 0.0               # touch every 4 KB page
                return a
            self._host_ready = {name: Engine._copy_pool.submit(make, tuple(shape)) for name, shape in shapes.items()}

    def _download(self, ten, name=None):
        """Device tensor -> numpy array through cached pinned staging buffers (3 x 64 MB): the copy of chunk i + 1 over PCIe
        overlaps the host memcpy of chunk i, which a few threads share (numpy releases the GIL while copying; first-touch page
        faults of a fresh array dominate a single-threaded copy), and no 0.5 GB pinned allocation is made per call.  ``name``:
        use the array ``prepare_download`` made ready under that name when its shape fits."""
        ten = ten.contiguous()
        self.prepare_download()
        out = None
        ready = getattr(self, "_host_ready", {}).pop(name, None) if name is not None else None
        if ready is not None:
            out = ready.result()
            if out.shape != tuple(ten.shape):
                out = None
        if out is None:
            out = np.empty(tuple(ten.shape), dtype=np.float64)
        flat_d, flat_h = ten.view(-1), out.reshape(-1)
        key = self.device.index
        nb = Engine._STAGE_BUFS
        bufs, pool = Engine._stage_cache[key], Engine._copy_pool
        host = [b.numpy() for b in bufs]
        step = bufs[0].numel()
        stream = torch.cuda.current_stream(self.device)
        evs = [torch.cuda.Event() for _ in range(nb)]
        pending = [[] for _ in range(nb)]                # host copies still reading staging buffer j
        n = flat_d.numel()
        chunks = [(o, min(step, n - o)) for o in range(0, n, step)]

        def drain(i):                                    # chunk i has arrived: hand its slices to the copy threads
            o, m = chunks[i]
            j = i % nb
            evs[j].synchronize()
            piece = -(-m // Engine._COPY_THREADS)
            for q in range(0, m, piece):
                e = min(m, q + piece)
                pending[j].append(pool.submit(np.copyto, flat_h[o + q:o + e], host[j][q:e]))

        for i, (o, m) in enumerate(chunks):
            j = i % nb
            for f in pending[j]:
                f.result()
            pending[j] = []
            bufs[j][:m].copy_(flat_d[o:o + m], non_blocking=True)
            evs[j].record(stream)
            if i:
                drain(i - 1)
        if chunks:
            drain(len(chunks) - 1)
        for lst in pending:
            for f in lst:
                f.result()
        return out

    def dot_solution(self, geometry, centred, root=None):
        """The DOT-unit solution of utils/type.py:48-65 (mu * area_v / 3, E * area_f) and, if ``centred``, the time-centred
        mu of socp/solver_decorator.py:32-34, formed on the device so that only the final mu and E cross PCIe.  The mass
        diagnostics the caller prints (interface.py:313-314) come along as ``diagnostics`` (2 x (nT [+1]) doubles).
        Sharded runs: ``root`` = the rank that assembles and downloads the solution (the other ranks return mu = E = None);
        None = every rank gets the full solution."""
        dev = self.device
        mu_i, E_i = self.from_internal("mu", root=root), self.from_internal("E", root=root)
        if mu_i is None:
            return dict(mu=None, E=None, diagnostics=None, root=root)
        av = torch.as_tensor(np.asarray(geometry["area_vertices"], dtype=np.float64), device=dev)[None, :] / 3.0
        af = torch.as_tensor(np.asarray(geometry["area_triangles"], dtype=np.float64), device=dev)[None, :, None]
        mu = (mu_i * (self.r * self.ds)) * av
        if centred:
            mu0 = torch.as_tensor(np.asarray(geometry["mu0"], dtype=np.float64), device=dev)[None, :]
            mu1 = torch.as_tensor(np.asarray(geometry["mu1"], dtype=np.float64), device=dev)[None, :]
            mu = torch.cat([mu0, 0.5 * (mu[:-1] + mu[1:]), mu1], dim=0)
        E = (E_i * (self.r * self.ds)) * af
        # utils/evaluate_solution.py:7-45 on the device: per-layer mass and per-layer negative mass of the returned mu
        layers = torch.stack([mu.sum(dim=1), torch.where(mu < 0, mu, torch.zeros_like(mu)).sum(dim=1)]).cpu().numpy()
        n = layers.shape[1]
        diagnostics = dict(mass_time_layers=layers[0], negative_mass_time_layers=layers[1],
                           mass_conservation=float(np.linalg.norm(layers[0] - 1.0) / np.sqrt(n)),
                           negative_mass=float(np.linalg.norm(layers[1]) / np.sqrt(n)))
        return dict(mu=self._download(mu, "mu"), E=self._download(E, "E"), diagnostics=diagnostics)

    def congestion_norm(self):
        """``||lambda_c - congestion * mu||_2`` of the un-scaled solution (the solver's closing log line,
        socp/solver_socp.py:846-853), formed on the device from the owned steps (+ a rank-ordered sum of the partial squares);
        a norm does not care about the vertex ordering."""
        part = self.part
        lc = self.slab["lam_c"].levels(part.lvl_begin, max(part.t_end, part.lvl_begin))
        mu = self.slab["mu"].levels(part.lvl_begin, max(part.t_end, part.lvl_begin))
        # on the un-scaled solution, with the congestion as the reference holds it at that point (:846-853)
        diff = self.ps * lc - self.cong * (self.r * self.ds) * mu
        local = np.array([float((diff * diff).sum())])
        return float(math.sqrt(self.comm.sum_in_rank_order(local, self.device)[0]))

    def solution(self, keys=None):
        """Un-scaled solution dict with the reference's keys and layouts (:397-405, :855-869).

        ``keys`` restricts what is converted and downloaded (the DOT-unit decorators only need ``mu`` and ``E``; the two
        18-wide corner arrays are 3 GB each at the headline size)."""
        keys = tuple(REF_KEYS) if keys is None else tuple(keys)
        if "z_mid" in keys and not self.z_valid:
            raise capi.DotsError("z_mid was not materialised on the last iteration")
        r, s, ps, ds = self.r, self.s, self.ps, self.ds                                 # :397-405
        scale = dict(phi=ps, A=ps, B=ps, lam_c=ps, z_fst=ps / s, z_mid=ps / s, z_end=ps / s,
                     mu=r * ds, E=r * ds, b_fst=r * s * ds, b_mid=r * s * ds, b_end=r * s * ds)
        out = {}
        for key in keys:
            name = REF_KEYS[key]
            out[key] = (self.from_internal(name) * scale[name]).cpu().numpy()
        return out
