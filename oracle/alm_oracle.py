"""CPU oracle: a numpy/scipy restatement of the reference's inexact semi-proximal ALM iteration.

TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package.  The product
(``dots_socp_b200``) never does; it fails loudly when its CUDA library is missing.

Parity status: PINNED against outputs of the unmodified reference run in the build container
(``tests/golden/make_golden.py`` imports ``/root/reference`` through three import shims and stores
iterates / KKT histories / iteration counts under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
replays them).  The reference itself ships no tests or golden vectors.

Every function cites the reference lines it restates (paths relative to
``/root/reference/dot_surface_socp``).  Arithmetic follows the reference's expression order so the
iterates agree to rounding; only default-flag semantics that the CLI can reach plus ``is_palm`` are
covered (``is_constant_scaling=False``: prim_scale = dual_scale = 1 throughout).
"""
from __future__ import annotations

import math
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

SQRT3 = math.sqrt(3.0)


# =============================================================================
# mesh operators  (utils/surface_pre_computations_socp.py)
# =============================================================================
def mesh_quantities(vertices, triangles):
    """area_f (T,), corner angles (T,3), P1 basis gradients (T,3,3).  Ref :11-39, loop body vectorised."""
    p0, p1, p2 = (vertices[triangles[:, k]] for k in range(3))
    v01, v12, v20 = p1 - p0, p2 - p1, p0 - p2
    nrm = lambda a: np.sqrt(np.sum(a * a, axis=1))
    dot = lambda a, b: np.sum(a * b, axis=1)
    area = nrm(np.cross(v01, v12)) / 2                                            # :24
    ang = np.stack([np.arccos(dot(v01, -v20) / (nrm(v01) * nrm(v20))),            # :26-28
                    np.arccos(dot(v12, -v01) / (nrm(v12) * nrm(v01))),
                    np.arccos(dot(v20, -v12) / (nrm(v20) * nrm(v12)))], axis=1)
    base = np.stack([-v01 + v12 * (dot(v01, v12) / dot(v12, v12))[:, None],       # :30-32
                     -v12 + v20 * (dot(v12, v20) / dot(v20, v20))[:, None],
                     -v20 + v01 * (dot(v20, v01) / dot(v01, v01))[:, None]], axis=1)
    base = base / (np.sqrt(np.sum(base * base, axis=2)) ** 2)[:, :, None]         # :34-36  (/ norm**2)
    return area, ang, base


def mesh_matrices(vertices, triangles, area, ang, base):
    """G (3T x V), D = -G^T, cotan Laplacian L (V x V).  Ref :42-86."""
    n_v, n_t = vertices.shape[0], triangles.shape[0]
    rows = (3 * np.arange(n_t)[:, None, None] + np.arange(3)[None, None, :])      # row 3f+xyz  (:57)
    cols = np.broadcast_to(triangles[:, :, None], (n_t, 3, 3))                    # col tri[f,k] (:58)
    G = sp.coo_matrix((base.reshape(-1), (np.broadcast_to(rows, (n_t, 3, 3)).reshape(-1), cols.reshape(-1))),
                      shape=(3 * n_t, n_v)).tocsr()
    D = (-G.transpose()).tocsr()                                                   # :65
    w = 0.5 * np.cos(ang) / np.sin(ang)                                            # :68
    L = sp.csr_matrix((n_v, n_v))
    for k in range(3):                                                             # :70-84
        a, b = triangles[:, (k + 1) % 3], triangles[:, (k + 2) % 3]
        for r, c, s in ((a, b, 1.0), (a, a, -1.0), (b, a, 1.0), (b, b, -1.0)):
            L = L + sp.coo_matrix((s * w[:, k], (r, c)), shape=(n_v, n_v)).tocsr()
    return G, D, L


def corner_maps(n_v, triangles, area):
    """Incidence maps of ref :88-132: corner index i = k*T+f.

    Returns (M_area (V x 3T, weights |f|), area_v_sum (V,), M_one (3T x V, ones), area_v_sum at corners (3T,))."""
    n_t = triangles.shape[0]
    vert = triangles.T.reshape(-1)
    corner = np.arange(3 * n_t)
    w = np.tile(area, 3)
    M_area = sp.coo_matrix((w, (vert, corner)), shape=(n_v, 3 * n_t)).tocsr()
    area_v = M_area.dot(np.ones(3 * n_t))
    M_one = sp.coo_matrix((np.ones(3 * n_t), (corner, vert)), shape=(3 * n_t, n_v)).tocsr()
    return M_area, area_v, M_one, area_v[vert]


# =============================================================================
# space-time Laplacian inverse  (utils/laplacian_inverse_socp.py)
# =============================================================================
def build_laplacian_inverse(n_time, dt, area_v, L, eps=0.0, n_threads=1):
    """Ref :11-50: dense eigh of the Neumann time Laplacian, one sparse LU per time mode.

    ``n_threads > 1`` (bench.py's CPU arm only) factorises / solves the independent modes from a thread pool
    (SuperLU releases the GIL); the arithmetic per mode is unchanged."""
    n = n_time + 1
    Lt = np.zeros((n, n))
    i = np.arange(1, n_time)
    Lt[i, i], Lt[i, i + 1], Lt[i, i - 1] = -2.0, 1.0, 1.0
    Lt[0, 0], Lt[0, 1], Lt[-1, -1], Lt[-1, -2] = -1.0, 1.0, -1.0, 1.0
    Lt *= 1 / (dt ** 2)
    lam, Q = np.linalg.eigh(Lt)                                                    # :31
    mass = sp.diags([area_v], [0])
    make = lambda a: spla.splu((L + (lam[a] - eps) * mass).tocsc()).solve            # :35-41 (factorized == splu.solve)
    pool = None
    if n_threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(n_threads)
        solves = list(pool.map(make, range(n)))
    else:
        solves = [make(a) for a in range(n)]

    def invert(rhs):                                                               # :52-61
        hat = np.array(np.dot(Q.T, rhs))
        sol = np.zeros_like(hat)
        if pool is not None:
            for a, x in enumerate(pool.map(lambda a: solves[a](hat[a, :]), range(n))):
                sol[a, :] = x
        else:
            for a in range(n):
                sol[a, :] = solves[a](hat[a, :])
        return np.array(np.dot(Q, sol))

    return invert, lam, Q


# =============================================================================
# operators  (socp/solver_socp.py:875-1065)
# =============================================================================
def weighted_norm_sq(weight, n_avg, a):                    # :875-878
    return np.sum(a ** 2 * weight) / n_avg


def grad_time(dt, x):                                      # :881-884
    return np.diff(x, axis=0) / dt


def div_time(dt, m):                                       # :886-896
    out = np.zeros((m.shape[0] + 1, m.shape[1]))
    out[1:-1] = np.diff(m, axis=0) / dt
    out[0] = m[0] / dt
    out[-1] = -m[-1] / dt
    return out


def grad_space(G, x):                                      # :898-907
    return G.dot(x.T).T.reshape(x.shape[0], -1, 3)


def div_space(D, x):                                       # :909-921
    return D.dot(x.reshape(x.shape[0], -1).T).T


def decouple(b, scale_z=1.0):                              # :923-942
    n_time, n_t = b.shape[0] - 1, b.shape[1]
    aux = (scale_z / SQRT3) * b
    out = np.empty((n_time, 2, 3, n_t, 3))
    out[:, 0] = aux[:-1, None]
    out[:, 1] = aux[1:, None]
    return out


def decouple_adjoint(x, scale_z=1.0):                      # :944-959
    aux = (scale_z / SQRT3) * np.sum(x, axis=2)
    out = np.zeros((x.shape[0] + 1, x.shape[3], 3))
    out[:-1] = aux[:, 0]
    out[1:] += aux[:, 1]
    return out


def adjoint_time_average(x):                               # :961-974  correlate1d([.5,.5]) on the zero-padded array
    pad = np.concatenate([x, np.zeros((1,) + x.shape[1:])], axis=0)
    out = 0.5 * pad
    out[1:] += 0.5 * pad[:-1]
    return out


class MeshOps:
    """Everything ``solver_socp`` prepares before its loop (ref :97-236)."""

    def __init__(self, n_time, geometry, eps=0.0, build_inverse=True, n_threads=1):
        v = np.asarray(geometry["vertices"], dtype=np.float64)
        t = np.asarray(geometry["triangles"])
        self.nT, self.V, self.T = n_time, v.shape[0], t.shape[0]
        self.dt = 1.0 / n_time
        self.tri = t
        self.area_f, ang, self.base = mesh_quantities(v, t)
        self.G, self.D, self.L = mesh_matrices(v, t, self.area_f, ang, self.base)
        M_area, area_v, M_one, area_v_corner = corner_maps(self.V, t, self.area_f)
        self.area_v = area_v / 3.0                                                  # :112
        area_v_corner = area_v_corner / 3.0                                         # :113
        nT, V, T = self.nT, self.V, self.T
        self.w_space = np.broadcast_to(self.area_f[None, :, None], (nT + 1, T, 3))                # :139-142
        self.w_dec = np.broadcast_to(self.area_f[None, None, None, :, None], (nT, 2, 3, T, 3))    # :144-147
        self.w_time = np.broadcast_to(self.area_v[None, :], (nT, V))                              # :149-152
        self.w_center = np.broadcast_to(self.area_v[None, :], (nT + 1, V))                        # :154-157
        self.M_one = M_one                     # (3T x V): vertex -> corners                      (:161)
        self.M_oneT = M_one.transpose().tocsr()  # (V x 3T): corners -> vertex, unit weights        (:170)
        self.M_area = M_area                   # (V x 3T): corners -> vertex, weights |f|         (:168)
        third = (1.0 / 3.0) * (M_one[:T] + M_one[T:2 * T] + M_one[2 * T:])                        # :163-166
        self.V2T_third = third.tocsr()
        self.diag_soc = np.sqrt(np.tile(self.area_f, 3) / area_v_corner).reshape(3, T)            # :172-192
        self.eps = eps
        self.area_mesh = np.sum(self.area_f)
        if build_inverse:
            self.lap_inv, self.lam_t, self.Q = build_laplacian_inverse(n_time, self.dt, self.area_v, self.L, eps, n_threads)

    # weighted norms (:215-218)
    def nsq_center(self, a): return weighted_norm_sq(self.w_center, self.nT + 1, a)
    def nsq_time(self, a): return weighted_norm_sq(self.w_time, self.nT, a)
    def nsq_space(self, a): return weighted_norm_sq(self.w_space, self.nT + 1, a)
    def nsq_dec(self, a): return weighted_norm_sq(self.w_dec, self.nT, a)

    def diag_b(self, s):                                   # :194-202
        d = 1.0 + (2.0 * s ** 2) * np.ones(self.nT + 1)
        d[0] = d[-1] = 1.0 + s ** 2
        return d[:, None, None]

    # ---- the four steps ------------------------------------------------------
    def phi_rhs(self, A, B, lam_c, mu, E, bnd, phi_old):   # :976-986 (the argument of laplacian_invert)
        return (div_time(self.dt, (A + lam_c - mu) * self.w_time)
                + div_space(self.D, (B - E) * self.w_space)
                - bnd - self.eps * self.w_center * phi_old)

    def solve_laplacian(self, A, B, lam_c, mu, E, bnd, phi_old):
        return self.lap_inv(self.phi_rhs(A, B, lam_c, mu, E, bnd, phi_old))

    def proj_soc(self, A, B, b_fst, b_mid, b_end, d, s):   # :988-1042
        nT, V, T = self.nT, self.V, self.T
        Bd = decouple(B, s)
        p = d - s * A - b_fst                               # :997
        w = self.diag_soc[None, None, :, :, None] * (Bd - b_mid)   # :998
        e = d + s * A - b_end                               # :999
        sq = w ** 2                                         # :1003
        corner = (sq[:, 0, :, :, 0] + sq[:, 0, :, :, 1] + sq[:, 0, :, :, 2]
                  + sq[:, 1, :, :, 0] + sq[:, 1, :, :, 1] + sq[:, 1, :, :, 2])        # :1005-1014  (nT,3,T)
        nrm = self.M_oneT.dot(corner.reshape(nT, 3 * T).T).T                          # :1004-1016
        nrm = np.sqrt(nrm + e ** 2)                         # :1017
        with np.errstate(divide="ignore", invalid="ignore"):
            lam = np.clip(0.5 * (1.0 + p / nrm), 0.0, 1.0)  # :1018
        ind = lam >= 1.0                                    # :1019
        lam_corner = self.M_one.dot(lam.T).T.reshape(nT, 3, T) / self.diag_soc[None]  # :1020-1030
        z_fst = np.where(ind, p, lam * nrm)                 # :1040
        z_mid = lam_corner[:, None, :, :, None] * w         # :1041
        z_end = lam * e                                     # :1042
        return z_fst, z_mid, z_end

    def solve_q_lambda(self, s, cong, r, dt_phi, dx_phi, mu, E, z_fst, z_mid, z_end, b_fst, b_mid, b_end):   # :1044-1065
        c1 = s * (1.0 + cong * r)
        c2 = 1.0 + 2.0 * s * c1
        memo_a = dt_phi + mu
        memo_b = decouple_adjoint(z_mid + b_mid, s)
        A = (1.0 / c2) * memo_a + (c1 / c2) * (z_end + b_end - z_fst - b_fst)
        B = (dx_phi + E + memo_b) / self.diag_b(s)
        lam_c = (cong * r / (1. + cong * r)) * (memo_a - A)
        return A, B, lam_c


# =============================================================================
# control logic  (utils/admm_tools.py:19-114, utils/condition_validator*.py) - compact restatement
# =============================================================================
def penalty_due(it, last_it):
    """admm_tools.py:30-52: returns True when the penalty update is due at iteration ``it``."""
    gap = it - last_it
    return ((it < 20 and gap >= 3) or (it < 50 and gap >= 7) or (it < 100 and gap >= 11)
            or (it < 200 and gap >= 17) or (it < 500 and gap >= 31) or gap >= 43)


_FACTOR_TABLE = ((50, 2.00), (35, 1.75), (20, 1.60), (10, 1.40), (5, 1.35), (3, 1.32),
                 (2.5, 1.28), (2, 1.26), (1.5, 1.20), (1.2, 1.10))


def penalty_new_value(r, gap):
    """admm_tools.py:54-95."""
    inv = gap < 1.0
    g = 1.0 / gap if inv else gap
    f = 1.0
    for thr, val in _FACTOR_TABLE:
        if g > thr:
            f = val
            break
    if inv:
        f = 1.0 / f
    return max(min(r * f, 10 ** 3), 10 ** (-3))


def _max_skip_none(vals):
    vals = [v for v in vals if v is not None]
    return max(vals) if vals else None


class LazyKKT:
    """condition_validator.py:194-331 + condition_validator_wrapper.py:9-134, for 7 two-valued conditions."""

    def __init__(self, funcs, tol, order=(6, 2, 0, 3, 1, 4, 5)):
        self.funcs, self.tol = funcs, tol
        self.n = len(funcs)
        self.slots = list(order)                 # queue position -> condition id (after optimize_queue_order)
        self.pos = {c: i for i, c in enumerate(self.slots)}
        self.front = 0
        self.last = [[None, None] for _ in range(self.n)]
        self.interval, self.counter = 1, 0

    def _eval(self, cond):
        vals = self.funcs[cond]()
        self.last[cond] = list(vals)
        return vals[0] < self.tol

    def sweep(self, required):
        """ConditionValidator.validate (:236-331). Returns (all_passed, n_checked)."""
        seen, n_checked = set(), 0
        req_ok = []
        for c in (required or []):
            q = self.pos[c]
            if q in seen:
                req_ok.append(True)
                continue
            seen.add(q); n_checked += 1
            req_ok.append(self._eval(c))
        done = False
        if all(req_ok) and n_checked < self.n:
            start = self.front
            while n_checked < self.n:
                q = self.front % self.n
                if q not in seen:
                    seen.add(q); n_checked += 1
                    if not self._eval(self.slots[q]):
                        break
                self.front = (self.front + 1) % self.n
                if self.front == start:
                    done = True
                    break
        elif all(req_ok):
            done = True
        return done, n_checked

    def validate(self, required=None):
        """AdaptiveValidatorWrapper.validate (wrapper :99-125)."""
        fire = (self.counter % self.interval) == 0
        self.counter += 1
        if fire or required:
            return self.sweep(required)
        return False, 0

    def reset_counter(self):
        self.counter = 0

    def pop_errors(self):
        out, self.last = self.last, [[None, None] for _ in range(self.n)]
        return out

    def retune(self, err):
        """wrapper :44-97 with min 1 / max 37."""
        ratio = err / max(self.tol, 1e-10)
        if ratio <= 1.0:
            self.interval = 1
            return
        lg = np.log10(ratio)
        self.interval = 37 if lg > 1.0 else max(1, int(1 + lg * 36))


# =============================================================================
# the solver  (socp/solver_socp.py:25-871)
# =============================================================================
class OracleALM:
    """State + one ALM iteration, written against MeshOps.  Default-flag semantics of ``solver_socp``."""

    def __init__(self, n_time, geometry, congestion=0.0, eps=0.0, tau=1.9, is_palm=False, is_z_scaling=True,
                 ops: MeshOps | None = None, init_solution=None):
        self.ops = ops or MeshOps(n_time, geometry, eps=eps)
        o = self.ops
        nT, V, T = o.nT, o.V, o.T
        self.cong, self.tau, self.is_palm = congestion, tau, is_palm
        self.r = 1.0
        self.s = 1.0            # scale_factor_z
        self.d = 1.0            # constant_d
        self.ps = 1.0           # prim_scale  (:318, is_constant_scaling only)
        self.ds = 1.0           # dual_scale  (:319)
        self.norm_d = math.sqrt(2 * o.area_mesh)                                   # :297
        z = np.zeros
        self.phi = z((nT + 1, V))
        self.A, self.lam_c, self.mu = z((nT, V)), z((nT, V)), z((nT, V))
        self.B, self.E = z((nT + 1, T, 3)), z((nT + 1, T, 3))
        self.z_fst, self.z_end, self.b_fst, self.b_end = z((nT, V)), z((nT, V)), z((nT, V)), z((nT, V))
        self.z_mid, self.b_mid = z((nT, 2, 3, T, 3)), z((nT, 2, 3, T, 3))
        if init_solution:                                                          # :239-250 (r = 1 here)
            g = lambda k, default: np.array(init_solution[k], dtype=np.float64, copy=True) if k in init_solution else default()
            self.phi = g("phi", lambda: self.phi)
            self.A = g("A", lambda: grad_time(o.dt, self.phi))
            self.B = g("B", lambda: grad_space(o.G, self.phi))
            self.lam_c = g("lambda_c", lambda: self.lam_c)
            self.z_fst, self.z_end, self.z_mid = g("z_fst", lambda: self.z_fst), g("z_end", lambda: self.z_end), g("z_mid", lambda: self.z_mid)
            self.b_fst, self.b_end, self.b_mid = g("beta_fst", lambda: self.b_fst), g("beta_end", lambda: self.b_end), g("beta_mid", lambda: self.b_mid)
            self.mu = g("mu", lambda: self.b_fst - self.b_end)
            self.E = g("E", lambda: -decouple_adjoint(self.b_mid, 1.0))
        self.Bd_new = z((nT, 2, 3, T, 3))       # memo_z_mid (:262, :717)
        self.dt_phi = np.array(0.0) if not is_palm else grad_time(o.dt, self.phi)  # :253-257
        self.dx_phi = np.array(0.0) if not is_palm else grad_space(o.G, self.phi)
        self.bnd = z((nT + 1, V))                                                  # :267-270
        self.bnd[0] = -np.asarray(geometry["mu0"]) / (self.r * o.dt)
        self.bnd[-1] = np.asarray(geometry["mu1"]) / (self.r * o.dt)
        self.norm_bnd = self.r * o.dt * math.sqrt(o.nsq_center(self.bnd / o.w_center))   # :296
        ma_c, ma_t = np.mean(o.area_v), np.mean(o.area_v)                          # :303-313
        ma_s = np.mean(o.area_f)
        self.k_prim_q = np.mean([ma_t, ma_s])
        self.k_prim_z = np.mean([ma_t, ma_s, ma_t])
        self.k_dual_a = ma_c
        self.k_dual_b = np.mean([ma_t, ma_s])
        self.k_comp_rho = ma_t
        self.k_comp_m = ma_s
        if is_z_scaling:
            self.scale_z(2.0)                                                      # :571-572

    # ---- scaling (:367-405) ---------------------------------------------------
    def adjust_penalty(self, f):
        self.r *= f
        for a in (self.mu, self.E, self.bnd, self.b_fst, self.b_mid, self.b_end):
            a /= f

    def scale_prim_dual(self, factors=None):                                       # :324-365
        o, sq = self.ops, math.sqrt
        if factors is None:                                                        # admm_tools.compute_scale_factor :171-174
            prim = [sq(o.nsq_time(self.dt_phi) + o.nsq_space(self.dx_phi)), sq(o.nsq_time(self.A) + o.nsq_space(self.B)),
                    sq(o.nsq_time(self.z_fst) + o.nsq_dec(self.z_mid) + o.nsq_time(self.z_end))]
            dual = [self.r * sq(o.nsq_time(self.mu) + o.nsq_space(self.E)),
                    self.r * sq(o.nsq_time(self.b_fst) + o.nsq_dec(self.b_mid) + o.nsq_time(self.b_end))]
            p, d = max(prim) / 1.0, max(dual) / 1.0
        else:
            p, d = factors
        if max(p, d) / min(p, d) > 2.0:                                            # :344
            self.ps *= p
            self.ds *= d
            for name in ("phi", "A", "B", "lam_c", "dt_phi", "dx_phi", "z_fst", "z_mid", "z_end"):
                setattr(self, name, getattr(self, name) / p)
            f = d ** 2 / p
            for name in ("bnd", "mu", "E", "b_fst", "b_mid", "b_end"):
                setattr(self, name, getattr(self, name) / f)
            self.r *= d / p
            self.cong *= d / p
            self.d /= p
            self.norm_d /= p
            self.norm_bnd /= d
            return True
        return False

    def initial_constant_scaling(self):                                            # :574-587
        o, sq = self.ops, math.sqrt
        bt = self.r * np.divide(self.bnd, o.w_center)
        norm_c = sq(o.nsq_center(bt))
        norm_ac = sq(o.nsq_time(grad_time(o.dt, bt)) + o.nsq_space(grad_space(o.G, bt)))
        self.scale_prim_dual((self.norm_d, sq(o.nT) * norm_c ** 2 / norm_ac))
        self.adjust_penalty(1.0 / self.r)

    def scale_z(self, f):
        self.s *= f; self.d *= f; self.norm_d *= f
        for a in (self.z_fst, self.z_mid, self.z_end):
            a *= self.s
        for a in (self.b_fst, self.b_mid, self.b_end):
            a *= 1.0 / self.s
        self.mu = self.s * (self.b_fst - self.b_end)
        self.E = -decouple_adjoint(self.b_mid, self.s)

    # ---- one iteration (:668-722) --------------------------------------------
    def step_q(self):
        o = self.ops
        self.A, self.B, self.lam_c = o.solve_q_lambda(self.s, self.cong, self.r, self.dt_phi, self.dx_phi, self.mu,
                                                      self.E, self.z_fst, self.z_mid, self.z_end,
                                                      self.b_fst, self.b_mid, self.b_end)

    def iterate(self):
        o, tau, s, d = self.ops, self.tau, self.s, self.d
        if self.is_palm:
            self.step_q()
        if getattr(self, "two_threads", False):          # the reference's default is_multi_threads=True (:674-696)
            import threading
            box = {}
            th = threading.Thread(target=lambda: box.setdefault("phi", o.solve_laplacian(
                self.A, self.B, self.lam_c, self.mu, self.E, self.bnd, self.phi)))
            th.start()
            self.z_fst, self.z_mid, self.z_end = o.proj_soc(self.A, self.B, self.b_fst, self.b_mid, self.b_end, d, s)
            th.join()
            phi = box["phi"]
        else:
            phi = o.solve_laplacian(self.A, self.B, self.lam_c, self.mu, self.E, self.bnd, self.phi)
            self.z_fst, self.z_mid, self.z_end = o.proj_soc(self.A, self.B, self.b_fst, self.b_mid, self.b_end, d, s)
        self.phi = phi
        self.dt_phi = grad_time(o.dt, self.phi)
        self.dx_phi = grad_space(o.G, self.phi)
        self.step_q()
        self.Bd_new = decouple(self.B, s)                                                  # :717
        self.mu = self.mu + tau * (self.dt_phi - self.A - self.lam_c)                      # :718
        self.E = self.E + tau * (self.dx_phi - self.B)                                     # :719
        self.b_fst = self.b_fst + tau * (self.z_fst + s * self.A - d)                      # :720
        self.b_mid = self.b_mid + tau * (self.z_mid - self.Bd_new)                         # :721
        self.b_end = self.b_end + tau * (self.z_end - s * self.A - d)                      # :722

    # ---- KKT residuals (:433-559, wiring :589-643); each returns [value, value] or [value, None] ----
    def kkt(self, i):
        o, s, r = self.ops, self.s, self.r
        sq = math.sqrt
        if i == 0:                                                                         # :433-450, :591-596
            norm_sum = (sq(o.nsq_time(self.dt_phi) + o.nsq_space(self.dx_phi))
                        + sq(o.nsq_time(self.A) + o.nsq_space(self.B)) + sq(o.nsq_time(self.lam_c)))
            res = sq(o.nsq_time(self.dt_phi - self.A - self.lam_c) + o.nsq_space(self.dx_phi - self.B))
            return [res / (self.k_prim_q / self.ps + norm_sum), res / (self.k_prim_q / 1.0 + norm_sum)]
        if i == 1:                                                                         # :452-464, :597-603
            res = sq(o.nsq_time(self.z_fst + s * self.A - self.d) + o.nsq_time(self.z_end - s * self.A - self.d)
                     + o.nsq_dec(s * (self.z_mid - self.Bd_new)))
            return [res / (self.k_prim_z / self.ps + self.norm_d), res / (self.k_prim_z / 1.0 + self.norm_d)]
        if i == 2:                                                                         # :466-482
            aux = (r * o.dt) * np.divide(self.bnd + div_time(o.dt, self.mu * o.w_time)
                                         + div_space(o.D, self.E * o.w_space), o.w_center)
            res = sq(o.nsq_center(aux))
            return [res / (self.k_dual_a / self.ds + self.norm_bnd), res / (self.k_dual_a / 1.0 + self.norm_bnd)]
        if i == 3:                                                                         # :484-503
            a1 = s * (self.b_end - self.b_fst)
            a2 = decouple_adjoint(self.b_mid, s)
            norm_sum = r * (sq(o.nsq_time(self.mu) + o.nsq_space(self.E)) + sq(o.nsq_time(a1) + o.nsq_space(a2)))
            res = r * sq(o.nsq_time(self.mu + a1) + o.nsq_space(self.E + a2))
            return [res / (self.k_dual_b / self.ds + norm_sum), res / (self.k_dual_b / 1.0 + norm_sum)]
        rho = (self.ds * r) * self.mu                                                      # :619-637: un-scaled arguments
        qA, qB, lam_c = self.ps * self.A, self.ps * self.B, self.ps * self.lam_c
        if i == 4:                                                                         # :505-526
            corner = np.sum(np.square(decouple(qB)), axis=(1, 4)).reshape(o.nT, 3 * o.T)
            aux = qA + .25 * np.divide(o.M_area.dot(corner.T).T, o.w_time)
            norm_sum = sq(o.nsq_time(rho)) + sq(o.nsq_time(aux))
            res = sq(o.nsq_time(np.maximum(0., aux + rho) - rho))
            return [res / (self.k_comp_rho + norm_sum), None]
        if i == 5:                                                                         # :528-547
            m = (self.ds * r) * self.E
            avg = o.V2T_third.dot(adjoint_time_average(rho).T).T[:, :, None] * qB
            norm_sum = sq(o.nsq_space(m)) + sq(o.nsq_space(avg))
            res = sq(o.nsq_space(avg - m))
            return [res / (self.k_comp_m + norm_sum), None]
        if i == 6:                                                                         # :549-559
            norm_sum = sq(o.nsq_time(rho)) + sq(o.nsq_time(lam_c))
            res = sq(o.nsq_time(self.cong * rho - lam_c))                                  # the (scaled) congestion, as the reference has it
            return [res / (self.k_comp_rho + norm_sum), None]
        raise IndexError(i)

    def objective(self):                                                                   # :417-431, :829-831
        o = self.ops
        phi, lam_c, bnd = self.ps * self.phi, self.ps * self.lam_c, (self.ds * self.r) * self.bnd
        cong = self.cong * self.ps / self.ds
        cost = o.dt * (np.dot(phi[0], bnd[0]) + np.dot(phi[-1], bnd[-1]))
        if cong > 10 ** (-10):
            return cost, cost - 1. / (2. * cong) * o.nsq_time(lam_c)
        return cost, cost

    def state(self, copy=True):
        names = ("phi", "A", "B", "lam_c", "mu", "E", "z_fst", "z_mid", "z_end", "b_fst", "b_mid", "b_end")
        return {n: (getattr(self, n).copy() if copy else getattr(self, n)) for n in names}

    def solution(self):
        """Un-scaled output dict, keys of utils/type.py:22-38 (ref :397-405, :855-869)."""
        r, s, ps, ds = self.r, self.s, self.ps, self.ds
        return dict(phi=ps * self.phi, A=ps * self.A, B=ps * self.B, lambda_c=ps * self.lam_c,
                    mu=(r * ds) * self.mu, E=(r * ds) * self.E,
                    z_fst=(ps / s) * self.z_fst, z_mid=(ps / s) * self.z_mid, z_end=(ps / s) * self.z_end,
                    beta_fst=(r * s * ds) * self.b_fst, beta_mid=(r * s * ds) * self.b_mid, beta_end=(r * s * ds) * self.b_end)


def solve(n_time, geometry, congestion=0.0, nit=1000, eps=0.0, tol=1e-4, tau=1.9, is_palm=False,
          is_z_scaling=True, time_limit=1000, trace=None, ops=None, check_kkt_step_by_step=False, init_solution=None,
          is_constant_scaling=False):
    """The reference's outer loop (:565-871) around OracleALM.  Returns (solution, info).

    ``info``: iterations (= last 0-based index, what the reference prints), kkt rows (nan = not
    evaluated), r per iteration, costs.  ``trace(it, alm)`` is called after every iteration."""
    alm = OracleALM(n_time, geometry, congestion, eps, tau, is_palm, is_z_scaling, ops=ops, init_solution=init_solution)
    if is_constant_scaling:
        alm.initial_constant_scaling()                                                     # :574-587
    t0 = time.perf_counter()
    prim_gap = 1.0 + 1.0 * np.exp(-100 * congestion)                                       # :568
    lazy = LazyKKT([(lambda i=i: alm.kkt(i)) for i in range(7)], tol)
    rows, its, r_hist, costs = [], [], [], []
    last_adjust, z_rescales, use_org = -1, 0, False
    last_row = np.full(7, np.inf)
    it, passed = -1, False
    for it in range(nit):
        if is_constant_scaling and (it == 10 or it == 50 or it % 100 == 50):              # :657-659, admm_tools :98-104
            alm.scale_prim_dual()
        if is_z_scaling and it >= 100 and z_rescales < 1 and max(last_row) < 5e-3:        # :661-666, admm_tools :107-114
            z_rescales += 1
            f = prim_gap * math.sqrt(last_row[1] / last_row[0])
            if f > 1.25:
                alm.scale_z(f)
        alm.iterate()
        time_up = (time.perf_counter() - t0) > time_limit                                  # :725
        due = penalty_due(it, last_adjust)                                                 # :726
        if due:
            last_adjust = it
        adjust = due or time_up
        required = [0, 1, 2, 3] if adjust else None
        if check_kkt_step_by_step:                                                         # :769-787
            passed, _ = lazy.validate(list(range(7)))
            costs.append(alm.objective())
        else:
            if adjust:
                lazy.reset_counter()
            passed, _ = lazy.validate(required)                                            # :738
        errs = lazy.pop_errors()
        org = [e[0] for e in errs]
        sec = [e[1] for e in errs]
        if adjust and not check_kkt_step_by_step:
            lazy.reset_counter()
        last_row = np.array([np.nan if v is None else v for v in org], dtype=float)        # record(): None -> nan
        rows.append(last_row); its.append(it); r_hist.append(alm.r)
        err = _max_skip_none([org[k] for k in (0, 2, 4, 5)])                               # :751
        if err is not None and not check_kkt_step_by_step:
            lazy.retune(err)
        if trace is not None:
            trace(it, alm)
        if passed or time_up:                                                              # :804
            break
        mx = _max_skip_none(sec)
        if mx is not None and mx < 5 * tol:                                                # :808-810
            use_org = True
        if adjust:                                                                         # :813-823
            src = org if use_org else sec
            gap = _max_skip_none(src[0:2]) / _max_skip_none(src[2:4])
            alm.adjust_penalty(penalty_new_value(alm.r, gap) / alm.r)
    final = [alm.kkt(i)[0] for i in range(7)]                                              # :826-828
    cost, lagr = alm.objective()
    info = dict(iterations=it, converged=bool(passed), kkt_rows=np.array(rows), kkt_iteration=np.array(its),
                r_history=np.array(r_hist), final_kkt=np.array(final), cost=cost, objective=lagr, cost_history=np.array(costs),
                running_time=time.perf_counter() - t0, r=alm.r, scale_z=alm.s)
    return alm.solution(), info
