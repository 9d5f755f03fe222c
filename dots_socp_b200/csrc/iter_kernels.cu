// Fused per-iteration kernels of the inexact semi-proximal ALM (reference: socp/solver_socp.py:656-823).
//
// One iteration of the reference touches the 18-wide corner arrays z_mid / beta_mid seven times
// (projection :988-1042, q-step :1044-1065, multiplier step :716-722).  Here the three steps are
// re-associated by *where* their data lives:
//
//   k_phi_rhs  per (t,v)   : right-hand side of the space-time Laplacian (:976-986)
//   k_vertex   per (t,v)   : everything that is local to a (time, vertex) pair once the cone norm is
//                            known: dt_phi, lam, z_fst, z_end (:1017-1042), A, lam_c (:1056-1065),
//                            mu, beta_fst, beta_end (:718-722)
//   k_tri      per (tau,f) : everything local to a (time level, triangle) pair: dx_phi (:898-907),
//                            z_mid (:1041), the adjoint sum (:944-959), B (:1064), E, beta_mid (:719-721)
//                            and the two corner by-products the NEXT iteration needs - the squared cone
//                            norms per corner (:998-1014) and the divergence terms per corner (:980).
//
// so beta_mid is read once and written once per iteration.  All arithmetic is fp64 and keeps the
// reference's expression order inside each formula.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// rhs[t][v] = div_t((A + lam_c - mu) area_v) + D((B - E) area_f) - boundary - eps area_v phi      (:979-986)
__global__ void __launch_bounds__(256) k_phi_rhs(dots_ctx_t c)
{
    pdl_launch_dependents();
    pdl_wait();
    const int V = c.n_vert, nT = c.n_time, T = c.n_tri;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = c.lvl_begin + blockIdx.y;
    if (v >= V) return;
    const double dt = 1.0 / nT;
    const double av = c.area_v[v];
    const double eps = c.params[DOTS_P_EPS];
    double divt;
    if (t == 0) {
        const size_t i = v;
        divt = ((c.A[i] + c.lam_c[i] - c.mu[i]) * av) / dt;                                   // :893
    } else if (t == nT) {
        const size_t i = (size_t)(nT - 1) * V + v;
        divt = -((c.A[i] + c.lam_c[i] - c.mu[i]) * av) / dt;                                  // :894
    } else {
        const size_t i1 = (size_t)t * V + v, i0 = i1 - V;
        divt = ((c.A[i1] + c.lam_c[i1] - c.mu[i1]) * av - (c.A[i0] + c.lam_c[i0] - c.mu[i0]) * av) / dt;   // :892
    }
    const double *cd = c.corner_div + (size_t)t * 3 * T;
    double divx = 0.0;
    const vc_row vr = vc_load(c, v);
    if (vr.id[7] != -2) {
        double val[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) val[q] = vr.id[q] >= 0 ? cd[vr.id[q]] : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) divx += val[q];
    } else {
        for (int q = c.vc_ptr[v], qe = c.vc_ptr[v + 1]; q < qe; ++q) divx += cd[c.vc_idx[q]];
    }
    const double bnd = (t == 0) ? c.bnd0[v] : ((t == nT) ? c.bnd1[v] : 0.0);
    const size_t o = (size_t)t * V + v;
    const double val = divt + divx - bnd - eps * av * c.phi[o];
    if (c.peer_rhs[0]) {                              // peer memory: the slab lands in every rank's rhs buffer directly
        for (int p = 0; p < c.n_ranks; ++p) c.peer_rhs[p][o] = val;
    } else {
        c.rhs[o] = val;
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vertex(dots_ctx_t c)
{
    pdl_launch_dependents();
    pdl_wait();
    const int V = c.n_vert, nT = c.n_time, T = c.n_tri;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = c.lvl_begin + blockIdx.y;         // owned staggered steps
    if (v >= V) return;
    const double *prm = c.params;
    const double r = prm[DOTS_P_R], s = prm[DOTS_P_S], d = prm[DOTS_P_D], cong = prm[DOTS_P_CONG], tau = prm[DOTS_P_TAU];
    const double dt = 1.0 / nT;
    const size_t i = (size_t)t * V + v;

    const double dtphi = (c.phi[i + V] - c.phi[i]) / dt;                                      // :884

    // cone norm: corners contribute their side-0 part at level t and their side-1 part at level t+1
    const double *n0 = c.corner_nrm + ((size_t)t * 2 + 0) * 3 * T;
    const double *n1 = c.corner_nrm + ((size_t)(t + 1) * 2 + 1) * 3 * T;
    double nsq = 0.0;
    const vc_row vr = vc_load(c, v);
    if (vr.id[7] != -2) {
        double v0[8], v1[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { v0[q] = vr.id[q] >= 0 ? n0[vr.id[q]] : 0.0; v1[q] = vr.id[q] >= 0 ? n1[vr.id[q]] : 0.0; }
#pragma unroll
        for (int q = 0; q < 8; ++q) nsq += v0[q] + v1[q];                                      // :1004-1016
    } else {
        for (int q = c.vc_ptr[v], qe = c.vc_ptr[v + 1]; q < qe; ++q) {
            const int cid = c.vc_idx[q];
            nsq += n0[cid] + n1[cid];                                                         // :1004-1016
        }
    }
    const double A0 = c.A[i], bf = c.b_fst[i], be = c.b_end[i], mu0 = c.mu[i];
    const double p = d - s * A0 - bf;                                                         // :997
    const double e = d + s * A0 - be;                                                         // :999
    const double nrm = sqrt(nsq + e * e);                                                     // :1017
    const double lam = clip01(0.5 * (1.0 + p / nrm));                                         // :1018
    const double zf = (lam >= 1.0) ? p : lam * nrm;                                           // :1040
    const double ze = lam * e;                                                                // :1042

    const double c1 = s * (1.0 + cong * r);                                                   // :1049
    const double c2 = 1.0 + 2.0 * s * c1;                                                     // :1050
    const double memo_a = dtphi + mu0;                                                        // :1051
    const double An = (1.0 / c2) * memo_a + (c1 / c2) * (ze + be - zf - bf);                  // :1062
    const double lc = (cong * r / (1. + cong * r)) * (memo_a - An);                           // :1065

    const double mun = mu0 + tau * (dtphi - An - lc);                                         // :718
    c.lam[i] = lam;
    c.z_fst[i] = zf;
    c.z_end[i] = ze;
    c.A[i] = An;
    c.lam_c[i] = lc;
    c.mu[i] = mun;
    if (c.peer_vertex[0] && t == min(c.lvl_end, nT) - 1) {      // last owned step: halo rows of the next rank (peer memory)
        c.peer_vertex[0][v] = lam;
        c.peer_vertex[1][v] = An;
        c.peer_vertex[2][v] = lc;
        c.peer_vertex[3][v] = mun;
    }
    c.b_fst[i] = bf + tau * (zf + s * An - d);                                                // :720
    c.b_end[i] = be + tau * (ze - s * An - d);                                                // :722
}

// ------------------------------------------------------------------------------------------------
// MODE 0: full step.  MODE 1: full step + store z_mid.  MODE 2: only recompute corner_nrm / corner_div.
// MODE 3: the triangle half of the is_palm Step 0 (solver_socp.py:668-672 = vanilla_solve_q_lambda :1044-1065 with the
//         STORED z_mid): B = (dx_phi + E + adj(z_mid + b_mid)) / diag_b, then the corner terms of the new B (E, b_mid unchanged).
// One thread owns triangle f for TRI_TCH consecutive time levels, so the mesh constants (hat gradients, cone
// diagonal, vertex ids) are loaded once per chunk and lam[tau] is reused as lam[tau-1] of the next level.
// beta_mid is streamed twice per level from the same thread (second time out of L1) instead of being held in
// 36 registers: that keeps the kernel at 4 CTAs/SM.
#define TRI_TCH 8
template <int MODE>
__global__ void __launch_bounds__(128, 4) k_tri(dots_ctx_t c)
{
    const int V = c.n_vert, nT = c.n_time;
    const size_t T = (size_t)c.n_tri;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= c.n_tri) return;
    const int tau_begin = c.lvl_begin + blockIdx.y * TRI_TCH;
    const int tau_end = min(tau_begin + TRI_TCH, c.lvl_end);
    const double *prm = c.params;
    const double s = prm[DOTS_P_S], step = prm[DOTS_P_TAU];
    const double cs = s / sqrt(3.0);                                                          // :932, :953

    double g[3][3], dg[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        dg[k] = c.diag_soc[k * T + f];
#pragma unroll
        for (int x = 0; x < 3; ++x) g[k][x] = c.hat_grad[(k * 3 + x) * T + f];
    }
    const double af = c.area_f[f];
    int vk[3] = {0, 0, 0};
    double lam_prev[3] = {0.0, 0.0, 0.0};
    if (MODE != 2) {
        vk[0] = c.tri[f]; vk[1] = c.tri[T + f]; vk[2] = c.tri[2 * T + f];
        if (MODE != 3 && tau_begin > 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) lam_prev[k] = c.lam[(size_t)(tau_begin - 1) * V + vk[k]];
        }
    }

    for (int tau = tau_begin; tau < tau_end; ++tau) {
        const bool has0 = tau < nT, has1 = tau > 0;
        double *Bp = c.B + (size_t)tau * 3 * T + f;
        double *Ep = c.E + (size_t)tau * 3 * T + f;
        double *bm = c.b_mid + (size_t)tau * 18 * T + f;
        double Bn[3], En[3];
        double lamk[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
        double bs[3] = {0.0, 0.0, 0.0};

        if (MODE == 2) {
#pragma unroll
            for (int x = 0; x < 3; ++x) { Bn[x] = Bp[x * T]; En[x] = Ep[x * T]; }
        } else if (MODE == 3) {
            const double *ph = c.phi + (size_t)tau * V;
            const double p0 = ph[vk[0]], p1 = ph[vk[1]], p2 = ph[vk[2]];
            const double *zq = c.z_mid + (size_t)tau * 18 * T + f;
            const double db = (tau == 0 || tau == nT) ? (1.0 + s * s) : (1.0 + (2.0 * s * s));    // :195-197
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                const double dx = g[0][x] * p0 + g[1][x] * p1 + g[2][x] * p2;                 // :902-906
                double adj = 0.0;
                if (has0) adj = cs * (((zq[(0 * 3 + x) * T] + bm[(0 * 3 + x) * T]) + (zq[(1 * 3 + x) * T] + bm[(1 * 3 + x) * T]))
                                      + (zq[(2 * 3 + x) * T] + bm[(2 * 3 + x) * T]));          // :1052, :953
                if (has1) {
                    const double s1 = cs * (((zq[(9 + 0 * 3 + x) * T] + bm[(9 + 0 * 3 + x) * T]) + (zq[(9 + 1 * 3 + x) * T] + bm[(9 + 1 * 3 + x) * T]))
                                            + (zq[(9 + 2 * 3 + x) * T] + bm[(9 + 2 * 3 + x) * T]));
                    adj = has0 ? adj + s1 : s1;                                               // :955-957
                }
                En[x] = Ep[x * T];
                Bn[x] = (dx + En[x] + adj) / db;                                              // :1064
                Bp[x * T] = Bn[x];
            }
        } else {
            const double *ph = c.phi + (size_t)tau * V;
            const double p0 = ph[vk[0]], p1 = ph[vk[1]], p2 = ph[vk[2]];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                lamk[0][k] = has0 ? c.lam[(size_t)tau * V + vk[k]] : 0.0;
                lamk[1][k] = lam_prev[k];
            }
            double Eo[3], dx[3];
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                const double Bo = Bp[x * T];
                Eo[x] = Ep[x * T];
                dx[x] = g[0][x] * p0 + g[1][x] * p1 + g[2][x] * p2;                           // :902-906
                bs[x] = cs * Bo;                                                              // :932
            }
            double adj[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                if (sd == 0 ? !has0 : !has1) continue;
                double sum[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double lt = lamk[sd][k] / dg[k];                                    // :1023
#pragma unroll
                    for (int x = 0; x < 3; ++x) {
                        const double b = bm[((sd * 3 + k) * 3 + x) * T];
                        const double w = dg[k] * (bs[x] - b);                                 // :998
                        const double zb = lt * w + b;                                         // :1041, :1052
                        sum[x] = (k == 0) ? zb : sum[x] + zb;                                 // np.sum(axis=2) :953
                    }
                }
#pragma unroll
                for (int x = 0; x < 3; ++x) adj[x] = (sd == 0 || !has0) ? cs * sum[x] : adj[x] + cs * sum[x];   // :953-957
            }
            const double db = (tau == 0 || tau == nT) ? (1.0 + s * s) : (1.0 + (2.0 * s * s));    // :195-197
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                Bn[x] = (dx[x] + Eo[x] + adj[x]) / db;                                        // :1064
                En[x] = Eo[x] + step * (dx[x] - Bn[x]);                                       // :719
                Bp[x * T] = Bn[x];
                Ep[x * T] = En[x];
            }
        }

        // second pass over beta_mid: multiplier update (:721) + the per-corner terms of the NEXT iteration
        double *zm = c.z_mid + (size_t)tau * 18 * T + f;
        double *cn = c.corner_nrm + (size_t)tau * 6 * T + f;
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            const bool has = (sd == 0) ? has0 : has1;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double acc = 0.0;
                if (has) {
                    const double lt = lamk[sd][k] / dg[k];
#pragma unroll
                    for (int x = 0; x < 3; ++x) {
                        double b = bm[((sd * 3 + k) * 3 + x) * T];
                        if (MODE == 0 || MODE == 1) {
                            const double zz = lt * (dg[k] * (bs[x] - b));                     // same arithmetic as pass one
                            // explicit fused operations: the rounding of the update must not depend on what else a MODE does with
                            // zz (the compiler contracted this line differently with and without the z_mid store: 1 ulp in b_mid)
                            b = __fma_rn(step, __fma_rn(-cs, Bn[x], zz), b);                  // b + step (z - cs B)   :717, :721
                            bm[((sd * 3 + k) * 3 + x) * T] = b;
                            if (MODE == 1) zm[((sd * 3 + k) * 3 + x) * T] = zz;
                        }
                        const double w = __dmul_rn(dg[k], __fma_rn(cs, Bn[x], -b));           // dg (cs B - b)   :995-998
                        acc += w * w;                                                         // :1003-1014
                    }
                }
                cn[(sd * 3 + k) * T] = acc;
                if (sd == 1 && c.peer_corner && tau == c.lvl_begin && has) c.peer_corner[k * T + f] = acc;   // halo of the previous rank
            }
        }
        double *cd = c.corner_div + (size_t)tau * 3 * T + f;
        double y[3];
#pragma unroll
        for (int x = 0; x < 3; ++x) y[x] = (Bn[x] - En[x]) * af;                              // :980
#pragma unroll
        for (int k = 0; k < 3; ++k) cd[k * T] = -(g[k][0] * y[0] + g[k][1] * y[1] + g[k][2] * y[2]);   // D = -G^T
#pragma unroll
        for (int k = 0; k < 3; ++k) lam_prev[k] = lamk[0][k];
    }
}

// ------------------------------------------------------------------------------------------------
// k_tri with TMA-staged input (MODE 0 / 1 as k_tri; 2: accumulate the triangle term of KKT #1 instead of storing z_mid; 3: the
// dual rescaling of a penalty update: E and beta_mid divided by `factor`, corner terms recomputed).  Same arithmetic as k_tri<0/1>; what changes is how the 24 input planes of a
// (time level, 128-triangle tile) reach the SM: thread 0 issues one 1-D bulk async copy (cp.async.bulk, 1 KB)
// per plane into a 3-stage shared-memory ring guarded by mbarriers, two time levels ahead of the math.  The
// bytes in flight per SM (up to 3 blocks x 2 stages x 24 KB) no longer depend on registers or occupancy, which is
// what the plain-load version was limited by (ncu: 12% warps active, 38% DRAM).  beta_mid is then read twice out of
// shared memory (projection pass, update pass) at no register cost.  Needs T even (16-byte plane alignment).
#define TRI_TILE 128
#define TRI_STAGES 3
#define TRI_PLANES 24
#define TRI_TMA_TCH 16
// ODD: the triangle count is odd (planes may start 8 bytes off a 16-byte boundary): shared-memory rows get room for the
// leading element of a misaligned run and every read adds the plane's shift; with an even count both compile away.
#define TRI_ROW_OF(ODD) (TRI_TILE + ((ODD) ? 2 : 0))
#define TRI_TMA_SMEM_OF(ODD) (TRI_STAGES * TRI_PLANES * TRI_ROW_OF(ODD) * 8 + 64)
template <int MODE, bool ODD>
__global__ void __launch_bounds__(TRI_TILE, 3) k_tri_tma(dots_ctx_t c, int tch, double factor)
{
    constexpr int TRI_ROW = TRI_ROW_OF(ODD);
    extern __shared__ __align__(128) unsigned char smraw[];
    double *tile = reinterpret_cast<double *>(smraw);                       // [stage][plane][TRI_ROW]
    uint64_t *bar = reinterpret_cast<uint64_t *>(smraw + TRI_STAGES * TRI_PLANES * TRI_ROW * 8);
    const int V = c.n_vert, nT = c.n_time;
    const size_t T = (size_t)c.n_tri;
    const int tid = threadIdx.x;
    const int f0 = blockIdx.x * TRI_TILE;
    const int nf = min(TRI_TILE, c.n_tri - f0);
    const int f = f0 + tid;
    const bool active = tid < nf;
    const int tau_begin = c.lvl_begin + blockIdx.y * tch;
    const int tau_end = min(tau_begin + tch, c.lvl_end);
    pdl_launch_dependents();

    if (tid == 0) {
        for (int i = 0; i < TRI_STAGES; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    double g[3][3], dg[3], af = 0.0;                                        // mesh constants: before the wait
    int vk[3] = {0, 0, 0};
    if (active) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dg[k] = c.diag_soc[k * T + f];
            vk[k] = c.tri[k * T + f];
#pragma unroll
            for (int x = 0; x < 3; ++x) g[k][x] = c.hat_grad[(k * 3 + x) * T + f];
        }
        af = c.area_f[f];
    }
    pdl_wait();                                                             // below: state written by the previous launches
    const double *prm = c.params;
    const double s = prm[DOTS_P_S], step = prm[DOTS_P_TAU];
    const double cs = s / sqrt(3.0);

    // A bulk copy needs 16-byte aligned addresses and sizes.  With an odd triangle count every other plane starts 8 bytes off:
    // such a run is fetched from one element earlier (`mis` = 1) and its data sits one slot later in the shared-memory row;
    // sizes are rounded up to an even element count (the element after a plane is the next plane's first, or the padding
    // the engine allocates behind the arrays).  Even counts: mis = 0 everywhere, the copies are exactly the planes.
    auto misaligned = [](const double *p) -> unsigned { return ODD ? ((unsigned)(reinterpret_cast<uintptr_t>(p) >> 3) & 1u) : 0u; };
    const unsigned todd = ODD ? 1u : 0u;
    auto issue = [&](int tau, int stage) {                                  // thread 0 only
        const bool h0 = tau < nT, h1 = tau > 0;
        double *dst = tile + (size_t)stage * TRI_PLANES * TRI_ROW;
        const double *bm = c.b_mid + (size_t)tau * 18 * T + f0;
        const double *Bp = c.B + (size_t)tau * 3 * T + f0, *Ep = c.E + (size_t)tau * 3 * T + f0;
        auto count = [&](const double *src) -> uint32_t { return ((uint32_t)nf + misaligned(src) + 1u) & ~1u; };
        uint32_t total = 0;
        for (int p = 0; p < 18; ++p)
            if ((p < 9) ? h0 : h1) total += count(bm + (size_t)p * T);
        for (int x = 0; x < 3; ++x) total += count(Bp + (size_t)x * T) + count(Ep + (size_t)x * T);
        mbar_expect_tx(&bar[stage], total * 8u);
        auto fetch = [&](int row, const double *src) {
            tma_load_1d(dst + row * TRI_ROW, src - misaligned(src), count(src) * 8u, &bar[stage]);
        };
        for (int p = 0; p < 18; ++p) {                                      // (an L2 evict_first hint on these copies measured 0.8 % slower)
            if ((p < 9) ? h0 : h1) fetch(p, bm + (size_t)p * T);
        }
        for (int x = 0; x < 3; ++x) {
            fetch(18 + x, Bp + (size_t)x * T);
            fetch(21 + x, Ep + (size_t)x * T);
        }
    };
    if (tid == 0) {
        for (int i = 0; i < TRI_STAGES && tau_begin + i < tau_end; ++i) issue(tau_begin + i, i);
    }

    double lam_prev[3] = {0.0, 0.0, 0.0};
    if (active && tau_begin > 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) lam_prev[k] = c.lam[(size_t)(tau_begin - 1) * V + vk[k]];
    }
    double kkt1 = 0.0;                                                      // MODE 2: this thread's share of KKT #1 (inactive threads: 0)

    for (int tau = tau_begin, it = 0; tau < tau_end; ++tau, ++it) {
        const int stage = it % TRI_STAGES;
        const uint32_t parity = (uint32_t)((it / TRI_STAGES) & 1);
        const bool has0 = tau < nT, has1 = tau > 0;
        // gathers first: they overlap the wait for the bulk copies
        double ph[3] = {0.0, 0.0, 0.0}, lamk[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
        if (active && MODE != 3) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ph[k] = c.phi[(size_t)tau * V + vk[k]];
                lamk[0][k] = has0 ? c.lam[(size_t)tau * V + vk[k]] : 0.0;
                lamk[1][k] = lam_prev[k];
            }
        }
        mbar_wait(&bar[stage], parity);
        if (active) {
            const double *srow = tile + (size_t)stage * TRI_PLANES * TRI_ROW + tid;
            double *Bp = c.B + (size_t)tau * 3 * T + f;
            double *Ep = c.E + (size_t)tau * 3 * T + f;
            double *bm = c.b_mid + (size_t)tau * 18 * T + f;
            // slot of this triangle in the row of plane p: + 1 where the plane's run starts 8 bytes off (f - tid = f0 is even)
            const unsigned mb = misaligned(bm - tid), mB = misaligned(Bp - tid), mE = misaligned(Ep - tid);
            auto sm_b = [&](int p) -> double { return srow[p * TRI_ROW + ((mb + (unsigned)p * todd) & 1u)]; };
            auto sm_B = [&](int x) -> double { return srow[(18 + x) * TRI_ROW + ((mB + (unsigned)x * todd) & 1u)]; };
            auto sm_E = [&](int x) -> double { return srow[(21 + x) * TRI_ROW + ((mE + (unsigned)x * todd) & 1u)]; };
            double Bn[3], En[3], Eo[3], dx[3], bs[3];
            if (MODE == 3) {                                                // penalty update: E / factor, B as it is (:370)
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    Bn[x] = sm_B(x);
                    En[x] = sm_E(x) / factor;
                    Ep[x * T] = En[x];
                    Eo[x] = dx[x] = bs[x] = 0.0;
                }
            } else {
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                const double Bo = sm_B(x);
                Eo[x] = sm_E(x);
                dx[x] = g[0][x] * ph[0] + g[1][x] * ph[1] + g[2][x] * ph[2];                  // :902-906
                bs[x] = cs * Bo;                                                              // :932
            }
            double adj[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                if (sd == 0 ? !has0 : !has1) continue;
                double sum[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double lt = lamk[sd][k] / dg[k];                                    // :1023
#pragma unroll
                    for (int x = 0; x < 3; ++x) {
                        const double b = sm_b((sd * 3 + k) * 3 + x);
                        const double w = dg[k] * (bs[x] - b);                                 // :998
                        const double zb = lt * w + b;                                         // :1041, :1052
                        sum[x] = (k == 0) ? zb : sum[x] + zb;                                 // np.sum(axis=2) :953
                    }
                }
#pragma unroll
                for (int x = 0; x < 3; ++x) adj[x] = (sd == 0 || !has0) ? cs * sum[x] : adj[x] + cs * sum[x];   // :953-957
            }
            const double db = (tau == 0 || tau == nT) ? (1.0 + s * s) : (1.0 + (2.0 * s * s));    // :195-197
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                Bn[x] = (dx[x] + Eo[x] + adj[x]) / db;                                        // :1064
                En[x] = Eo[x] + step * (dx[x] - Bn[x]);                                       // :719
                Bp[x * T] = Bn[x];
                Ep[x * T] = En[x];
            }
            }
            double *zm = c.z_mid + (size_t)tau * 18 * T + f;
            double *cn = c.corner_nrm + (size_t)tau * 6 * T + f;
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                const bool has = (sd == 0) ? has0 : has1;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double acc = 0.0;
                    if (has) {
                        const double lt = lamk[sd][k] / dg[k];
#pragma unroll
                        for (int x = 0; x < 3; ++x) {
                            double b = sm_b((sd * 3 + k) * 3 + x);
                            if (MODE == 3) {                                                  // penalty update: beta_mid / factor (:370)
                                b = b / factor;
                                bm[((sd * 3 + k) * 3 + x) * T] = b;
                                const double w3 = __dmul_rn(dg[k], __fma_rn(cs, Bn[x], -b));
                                acc += w3 * w3;
                                continue;
                            }
                            const double zz = lt * (dg[k] * (bs[x] - b));                     // same arithmetic as pass one
                            // explicit fused operations: the rounding of the update must not depend on what else a MODE does with
                            // zz (the compiler contracted this line differently with and without the z_mid store: 1 ulp in b_mid)
                            const double dz = __fma_rn(-cs, Bn[x], zz);                       // z - cs B: the residual of KKT #1 (:599)
                            b = __fma_rn(step, dz, b);                                        // :717, :721
                            bm[((sd * 3 + k) * 3 + x) * T] = b;
                            if (MODE == 1) zm[((sd * 3 + k) * 3 + x) * T] = zz;
                            if (MODE == 2) {                                                  // KKT #1, triangle term, as k_kkt_tri forms it
                                const double res = __dmul_rn(s, dz);                          // from the stored z_mid and B
                                kkt1 = __fma_rn(__dmul_rn(res, res), af, kkt1);
                            }
                            const double w = __dmul_rn(dg[k], __fma_rn(cs, Bn[x], -b));       // dg (cs B - b)   :995-998
                            acc += w * w;                                                     // :1003-1014
                        }
                    }
                    cn[(sd * 3 + k) * T] = acc;
                    if (sd == 1 && c.peer_corner && tau == c.lvl_begin && has) c.peer_corner[k * T + f] = acc;   // halo of the previous rank
                }
            }
            double *cd = c.corner_div + (size_t)tau * 3 * T + f;
            double y[3];
#pragma unroll
            for (int x = 0; x < 3; ++x) y[x] = (Bn[x] - En[x]) * af;                          // :980
#pragma unroll
            for (int k = 0; k < 3; ++k) cd[k * T] = -(g[k][0] * y[0] + g[k][1] * y[1] + g[k][2] * y[2]);   // D = -G^T
#pragma unroll
            for (int k = 0; k < 3; ++k) lam_prev[k] = lamk[0][k];
        }
        __syncthreads();                                                    // everyone is done reading this stage
        if (tid == 0 && tau + TRI_STAGES < tau_end) issue(tau + TRI_STAGES, stage);
    }
    if (MODE == 2) {                                                        // fixed-order block sum -> one partial per block
        __shared__ double kk[TRI_TILE / 32];
        const double w = warp_sum(kkt1);
        if ((tid & 31) == 0) kk[tid >> 5] = w;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
#pragma unroll
            for (int i = 0; i < TRI_TILE / 32; ++i) v += kk[i];
            c.kkt1_part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// is_palm Step 0, vertex half (solver_socp.py:668-672): A, lam_c from the gradient of the CURRENT phi and the current
// z, beta, mu (vanilla_solve_q_lambda :1049-1065).
__global__ void __launch_bounds__(256) k_palm_vertex(dots_ctx_t c)
{
    const int V = c.n_vert, nT = c.n_time;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = c.lvl_begin + blockIdx.y;
    if (v >= V) return;
    const double *prm = c.params;
    const double r = prm[DOTS_P_R], s = prm[DOTS_P_S], cong = prm[DOTS_P_CONG];
    const double dt = 1.0 / nT;
    const size_t i = (size_t)t * V + v;
    const double dtphi = (c.phi[i + V] - c.phi[i]) / dt;                                      // :884
    const double c1 = s * (1.0 + cong * r);                                                   // :1049
    const double c2 = 1.0 + 2.0 * s * c1;                                                     // :1050
    const double memo_a = dtphi + c.mu[i];                                                    // :1051
    const double An = (1.0 / c2) * memo_a + (c1 / c2) * (c.z_end[i] + c.b_end[i] - c.z_fst[i] - c.b_fst[i]);   // :1062
    c.A[i] = An;
    c.lam_c[i] = (cong * r / (1. + cong * r)) * (memo_a - An);                                // :1065
}

// ------------------------------------------------------------------------------------------------
// rescaling helpers (row a12)
__global__ void k_div_scalar(double *__restrict__ a, size_t n, double f)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = a[i] / f;
}
__global__ void k_mul_scalar(double *__restrict__ a, size_t n, double f)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] *= f;
}
// mu = s (b_fst - b_end)                                                                      (:387)
__global__ void k_mu_from_beta(dots_ctx_t c, double s, size_t i0, size_t n)
{
    for (size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        c.mu[i] = s * (c.b_fst[i] - c.b_end[i]);
}
// E = -decouple_adjoin_spacial(b_mid, s)                                                      (:388)
__global__ void k_E_from_beta(dots_ctx_t c, double s)
{
    const size_t T = (size_t)c.n_tri;
    const int nT = c.n_time;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int tau = c.lvl_begin + blockIdx.y;
    if (f >= c.n_tri) return;
    const double cs = s / sqrt(3.0);
    const double *bm = c.b_mid + (size_t)tau * 18 * T + f;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        double out = 0.0;
        if (tau < nT) out = cs * ((bm[(0 * 3 + x) * T] + bm[(1 * 3 + x) * T]) + bm[(2 * 3 + x) * T]);
        if (tau > 0) {
            const double s1 = cs * ((bm[(9 + 0 * 3 + x) * T] + bm[(9 + 1 * 3 + x) * T]) + bm[(9 + 2 * 3 + x) * T]);
            out = (tau < nT) ? out + s1 : s1;
        }
        c.E[((size_t)tau * 3 + x) * T + f] = -out;
    }
}

// ------------------------------------------------------------------------------------------------
// operator-level kernels on the internal layout (rows a5/a6): G phi and D x for all time levels
__global__ void k_grad_space(dots_ctx_t c, const double *__restrict__ phi, double *__restrict__ out)
{
    const size_t T = (size_t)c.n_tri;
    const int f = blockIdx.x * blockDim.x + threadIdx.x, tau = c.lvl_begin + blockIdx.y;
    if (f >= c.n_tri) return;
    const double *ph = phi + (size_t)tau * c.n_vert;
    const double p0 = ph[c.tri[f]], p1 = ph[c.tri[T + f]], p2 = ph[c.tri[2 * T + f]];
#pragma unroll
    for (int x = 0; x < 3; ++x)
        out[((size_t)tau * 3 + x) * T + f] = c.hat_grad[(0 * 3 + x) * T + f] * p0 + c.hat_grad[(1 * 3 + x) * T + f] * p1 + c.hat_grad[(2 * 3 + x) * T + f] * p2;
}
__global__ void k_div_space(dots_ctx_t c, const double *__restrict__ x, double *__restrict__ out)
{
    const size_t T = (size_t)c.n_tri;
    const int v = blockIdx.x * blockDim.x + threadIdx.x, tau = c.lvl_begin + blockIdx.y;
    if (v >= c.n_vert) return;
    const double *xt = x + (size_t)tau * 3 * T;
    double acc = 0.0;
    for (int q = c.vc_ptr[v], qe = c.vc_ptr[v + 1]; q < qe; ++q) {
        const int cid = c.vc_idx[q];
        const size_t k = cid / T, f = cid - k * T;
        acc += -(c.hat_grad[(k * 3 + 0) * T + f] * xt[f] + c.hat_grad[(k * 3 + 1) * T + f] * xt[T + f] + c.hat_grad[(k * 3 + 2) * T + f] * xt[2 * T + f]);
    }
    out[(size_t)tau * c.n_vert + v] = acc;
}

// ------------------------------------------------------------------------------------------------
extern "C" int dots_phi_rhs(const dots_ctx_t *c, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    return pdl_launch2(k_phi_rhs, dim3(ceil_div(c->n_vert, 256), c->lvl_end - c->lvl_begin), 256, (cudaStream_t)stream, c->ring_pdl != 0, *c);
}

extern "C" int dots_step_vertex(const dots_ctx_t *c, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    if (dots_t_end(c) <= c->lvl_begin) return 0;
    return pdl_launch2(k_vertex, dim3(ceil_div(c->n_vert, 256), dots_t_end(c) - c->lvl_begin), 256, (cudaStream_t)stream, c->ring_pdl != 0, *c);
}

// mode 0 / 1 / 2: the triangle half of an iteration (write_z); 3: E, beta_mid /= factor + corner terms (dots_scale_dual)
static int launch_tri_tma(const dots_ctx_t *c, int mode, double factor, cudaStream_t st)
{
    static bool configured[64] = {false};                  // per device (a second engine on another GPU of the process)
    int dev = 0;
    DOTS_CUDA(cudaGetDevice(&dev));
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    if (!configured[dev]) {
        DOTS_CUDA(cudaFuncSetAttribute(k_tri_tma<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_TMA_SMEM_OF(false)));
        DOTS_CUDA(cudaFuncSetAttribute(k_tri_tma<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_TMA_SMEM_OF(false)));
        DOTS_CUDA(cudaFuncSetAttribute(k_tri_tma<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_TMA_SMEM_OF(false)));
        DOTS_CUDA(cudaFuncSetAttribute(k_tri_tma<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_TMA_SMEM_OF(false)));
        DOTS_CUDA(cudaFuncSetAttribute(k_tri_tma<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_TMA_SMEM_OF(true)));
        DOTS_CUDA(cudaFuncSetAttribute(k_tri_tma<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_TMA_SMEM_OF(true)));
        DOTS_CUDA(cudaFuncSetAttribute(k_tri_tma<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_TMA_SMEM_OF(true)));
        DOTS_CUDA(cudaFuncSetAttribute(k_tri_tma<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_TMA_SMEM_OF(true)));
        configured[dev] = true;
    }
    int tch = 0;
    const int n_blocks = dots_tri_tma_blocks(c, &tch);
    const unsigned gx = (unsigned)ceil_div(c->n_tri, TRI_TILE), gy = (unsigned)(n_blocks / (int)gx);
    if (mode == 2 && (!c->kkt1_part || c->kkt1_blocks < n_blocks)) {
        dots_set_error("dots_step_tri(write_z = 2): kkt1_part holds %d partials, the grid has %d blocks", c->kkt1_part ? c->kkt1_blocks : 0, n_blocks);
        return DOTS_ERR_BAD_ARG;
    }
    const bool odd = (c->n_tri & 1) != 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, gy);
    cfg.blockDim = dim3(TRI_TILE);
    cfg.dynamicSmemBytes = odd ? TRI_TMA_SMEM_OF(true) : TRI_TMA_SMEM_OF(false);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (c->ring_pdl && mode != 3) ? 1 : 0;     // the rescaling follows plain launches: ordinary stream order
    if (odd) {
        switch (mode) {
        case 3: DOTS_CUDA(cudaLaunchKernelEx(&cfg, k_tri_tma<3, true>, *c, tch, factor)); break;
        case 2: DOTS_CUDA(cudaLaunchKernelEx(&cfg, k_tri_tma<2, true>, *c, tch, factor)); break;
        case 1: DOTS_CUDA(cudaLaunchKernelEx(&cfg, k_tri_tma<1, true>, *c, tch, factor)); break;
        default: DOTS_CUDA(cudaLaunchKernelEx(&cfg, k_tri_tma<0, true>, *c, tch, factor)); break;
        }
    } else {
        switch (mode) {
        case 3: DOTS_CUDA(cudaLaunchKernelEx(&cfg, k_tri_tma<3, false>, *c, tch, factor)); break;
        case 2: DOTS_CUDA(cudaLaunchKernelEx(&cfg, k_tri_tma<2, false>, *c, tch, factor)); break;
        case 1: DOTS_CUDA(cudaLaunchKernelEx(&cfg, k_tri_tma<1, false>, *c, tch, factor)); break;
        default: DOTS_CUDA(cudaLaunchKernelEx(&cfg, k_tri_tma<0, false>, *c, tch, factor)); break;
        }
    }
    return 0;
}

extern "C" int dots_step_tri(const dots_ctx_t *c, int write_z, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    if (write_z < 0 || write_z > 2) { dots_set_error("dots_step_tri: write_z = %d (0, 1 or 2)", write_z); return DOTS_ERR_BAD_ARG; }
    if (!(c->ring_flags & 4)) return launch_tri_tma(c, write_z, 1.0, (cudaStream_t)stream);   // TMA-staged kernel
    // ring_flags bit 2: the plain-load kernel (diagnostics)
    if (write_z == 2) { dots_set_error("dots_step_tri(write_z = 2) needs the TMA triangle kernel (ring_flags bit 2 is set)"); return DOTS_ERR_BAD_ARG; }
    dim3 grid(ceil_div(c->n_tri, 128), ceil_div(c->lvl_end - c->lvl_begin, TRI_TCH));
    if (write_z) k_tri<1><<<grid, 128, 0, (cudaStream_t)stream>>>(*c);
    else k_tri<0><<<grid, 128, 0, (cudaStream_t)stream>>>(*c);
    DOTS_LAUNCH_CHECK();
    return 0;
}

extern "C" int dots_step_q0(const dots_ctx_t *c, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    if (c->n_ranks != 1) { dots_set_error("dots_step_q0 (is_palm) is single-GPU"); return DOTS_ERR_BAD_ARG; }
    if (dots_t_end(c) > c->lvl_begin) {
        dim3 gv(ceil_div(c->n_vert, 256), dots_t_end(c) - c->lvl_begin);
        k_palm_vertex<<<gv, 256, 0, (cudaStream_t)stream>>>(*c);
        DOTS_LAUNCH_CHECK();
    }
    dim3 grid(ceil_div(c->n_tri, 128), ceil_div(c->lvl_end - c->lvl_begin, TRI_TCH));
    k_tri<3><<<grid, 128, 0, (cudaStream_t)stream>>>(*c);
    DOTS_LAUNCH_CHECK();
    return 0;
}

extern "C" int dots_refresh_corner_terms(const dots_ctx_t *c, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    dim3 grid(ceil_div(c->n_tri, 128), ceil_div(c->lvl_end - c->lvl_begin, TRI_TCH));
    k_tri<2><<<grid, 128, 0, (cudaStream_t)stream>>>(*c);
    DOTS_LAUNCH_CHECK();
    return 0;
}

static int launch_div(double *a, size_t n, double f, int n_sm, cudaStream_t st)
{
    if (!n) return 0;
    int blocks = (int)((n + 255) / 256);
    if (blocks > n_sm * 16) blocks = n_sm * 16;
    k_div_scalar<<<blocks, 256, 0, st>>>(a, n, f);
    DOTS_LAUNCH_CHECK();
    return 0;
}
static int launch_mul(double *a, size_t n, double f, int n_sm, cudaStream_t st)
{
    if (!n) return 0;
    int blocks = (int)((n + 255) / 256);
    if (blocks > n_sm * 16) blocks = n_sm * 16;
    k_mul_scalar<<<blocks, 256, 0, st>>>(a, n, f);
    DOTS_LAUNCH_CHECK();
    return 0;
}

extern "C" int dots_scale_dual(const dots_ctx_t *c, double factor, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t V = c->n_vert, T = c->n_tri;
    const size_t l0 = c->lvl_begin, nl = c->lvl_end - c->lvl_begin, nt = dots_t_end(c) - c->lvl_begin;
    int e;
    // mu carries one halo step in front (t = lvl_begin-1), scaled here so that it stays consistent without an exchange
    if ((e = launch_div(c->mu + ((long long)l0 - 1) * (long long)V, (nt + 1) * V, factor, c->n_sm, st))) return e;     // :370
    if ((e = launch_div(c->bnd0, V, factor, c->n_sm, st))) return e;
    if ((e = launch_div(c->bnd1, V, factor, c->n_sm, st))) return e;
    if ((e = launch_div(c->b_fst + l0 * V, nt * V, factor, c->n_sm, st))) return e;
    if ((e = launch_div(c->b_end + l0 * V, nt * V, factor, c->n_sm, st))) return e;
    if (!(c->ring_flags & 4))                              // E, beta_mid / factor inside the TMA-staged pass that recomputes the corner terms
        return launch_tri_tma(c, 3, factor, st);
    if ((e = launch_div(c->E + l0 * 3 * T, nl * 3 * T, factor, c->n_sm, st))) return e;
    if ((e = launch_div(c->b_mid + l0 * 18 * T, nl * 18 * T, factor, c->n_sm, st))) return e;
    return dots_refresh_corner_terms(c, stream);
}

extern "C" int dots_scale_z(const dots_ctx_t *c, double s_cum, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t V = c->n_vert, T = c->n_tri;
    const size_t l0 = c->lvl_begin, nl = c->lvl_end - c->lvl_begin, nt = dots_t_end(c) - c->lvl_begin;
    const size_t a = nt * V, z = nl * 18 * T;
    int e;
    if ((e = launch_mul(c->z_fst + l0 * V, a, s_cum, c->n_sm, st))) return e;                 // :383
    if ((e = launch_mul(c->z_mid + l0 * 18 * T, z, s_cum, c->n_sm, st))) return e;
    if ((e = launch_mul(c->z_end + l0 * V, a, s_cum, c->n_sm, st))) return e;
    const double inv = 1.0 / s_cum;                                                           // :384
    if ((e = launch_mul(c->b_fst + l0 * V, a, inv, c->n_sm, st))) return e;
    if ((e = launch_mul(c->b_mid + l0 * 18 * T, z, inv, c->n_sm, st))) return e;
    if ((e = launch_mul(c->b_end + l0 * V, a, inv, c->n_sm, st))) return e;
    if (a) {
        int blocks = ceil_div((long long)a, 256);
        if (blocks > c->n_sm * 16) blocks = c->n_sm * 16;
        k_mu_from_beta<<<blocks, 256, 0, st>>>(*c, s_cum, l0 * V, l0 * V + a);                // halo step of mu: host exchange
        DOTS_LAUNCH_CHECK();
    }
    dim3 grid(ceil_div(c->n_tri, 128), c->lvl_end - c->lvl_begin);
    k_E_from_beta<<<grid, 128, 0, st>>>(*c, s_cum);
    DOTS_LAUNCH_CHECK();
    return 0;   // caller sets params (s, d) and then calls dots_refresh_corner_terms
}

extern "C" int dots_scale_prim_dual(const dots_ctx_t *c, double prim_div, double dual_div, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t V = c->n_vert, T = c->n_tri;
    const long long l0 = c->lvl_begin;
    const size_t nl = c->lvl_end - c->lvl_begin, nt = dots_t_end(c) - c->lvl_begin;
    int e;
    // primal variables (:347-349); A, lam_c, lam carry a halo step in front (t = lvl_begin-1), phi a halo level behind
    if ((e = launch_div(c->phi + l0 * (long long)V, (nl + 1) * V, prim_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->A + (l0 - 1) * (long long)V, (nt + 1) * V, prim_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->lam_c + (l0 - 1) * (long long)V, (nt + 1) * V, prim_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->B + l0 * 3 * (long long)T, nl * 3 * T, prim_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->z_fst + l0 * (long long)V, nt * V, prim_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->z_mid + l0 * 18 * (long long)T, nl * 18 * T, prim_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->z_end + l0 * (long long)V, nt * V, prim_div, c->n_sm, st))) return e;
    // dual variables (:352-354)
    if ((e = launch_div(c->mu + (l0 - 1) * (long long)V, (nt + 1) * V, dual_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->E + l0 * 3 * (long long)T, nl * 3 * T, dual_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->bnd0, V, dual_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->bnd1, V, dual_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->b_fst + l0 * (long long)V, nt * V, dual_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->b_mid + l0 * 18 * (long long)T, nl * 18 * T, dual_div, c->n_sm, st))) return e;
    if ((e = launch_div(c->b_end + l0 * (long long)V, nt * V, dual_div, c->n_sm, st))) return e;
    return dots_refresh_corner_terms(c, stream);
}

extern "C" int dots_set_params(const dots_ctx_t *c, const double *host_params, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    DOTS_CUDA(cudaMemcpyAsync(c->params, host_params, sizeof(double) * DOTS_P_COUNT, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
}

extern "C" int dots_grad_space(const dots_ctx_t *c, const double *phi, double *out, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    dim3 grid(ceil_div(c->n_tri, 128), c->lvl_end - c->lvl_begin);
    k_grad_space<<<grid, 128, 0, (cudaStream_t)stream>>>(*c, phi, out);
    DOTS_LAUNCH_CHECK();
    return 0;
}

extern "C" int dots_div_space(const dots_ctx_t *c, const double *x, double *out, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    dim3 grid(ceil_div(c->n_vert, 128), c->lvl_end - c->lvl_begin);
    k_div_space<<<grid, 128, 0, (cudaStream_t)stream>>>(*c, x, out);
    DOTS_LAUNCH_CHECK();
    return 0;
}
