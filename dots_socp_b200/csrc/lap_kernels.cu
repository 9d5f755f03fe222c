// Space-time Laplacian inverse (reference: utils/laplacian_inverse_socp.py:52-61).
//
//   hat = Q^T rhs           dense time transform, (nT+1)^2 x V                       (:54)
//   hat[mode] <- (K + shift_mode diag(area_v))^-1 hat[mode]   for every mode at once   (:58-59)
//   phi = Q hat                                                                       (:61)
//
// The reference keeps one SuperLU factor per mode and solves them one after another.  Here all modes
// share one nested-dissection multifrontal factor structure (dots_socp_b200/nested.py) whose numeric
// values are stored "solve ready": per separator-tree node a dense panel
//        P = [ inv(L11) ; L21 inv(L11) ]     laid out [row][col][mode], mode fastest,
// so both sweeps are batched dense mat-vecs that stream every panel entry exactly once with
// 256/512/1024-byte coalesced rows (32/64/128 modes x 8 B), and the only ordering constraints are the
// parent/child ones between tree levels (one launch per level).
#include "common.cuh"

#define SWEEP_THREADS 256
#define SWEEP_WARPS 8

// ------------------------------------------------------------------------------------------------
// Time transforms as fp64 tensor-core GEMMs (DMMA, mma.sync.m8n8k4.f64: the only fp64 tensor path - tcgen05 has
// no f64 kind).  Per 64-vertex tile:   C[64][N] = A[64][K] * B[K][N]
//   DIR 0 (laplacian_inverse_socp.py:54)  hat[v][k] = sum_t rhs[t][v] Q[t][k]   A = rhs^T, B = qf (all levels x own modes)
//   DIR 1 (laplacian_inverse_socp.py:61)  phi[t][v] = sum_k hat[v][k] Q[t][k]   A = hat_all, B = qb (all modes x own levels)
// B (a slice of the time eigenbasis prepared by the host: this rank's modes / levels) is loaded into shared memory once
// per block; blocks are persistent over vertex tiles.
// Leading dimensions lda = K+4, ldb = N+8 make both fragment loads bank-conflict free.
#define TT_VT 64
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int DIR>
__global__ void __launch_bounds__(256) k_time_mma(dots_ctx_t c, int n_tiles)
{
    extern __shared__ double sm[];
    const int nt1 = c.n_time + 1, M = c.m_pad, V = c.n_vert;
    const int K = (DIR == 0) ? c.tt_kf : c.tt_kb;
    const int N = (DIR == 0) ? M : c.tt_nb;
    const double *Bg = (DIR == 0) ? c.qf : c.qb;           // [K][N], built by the host for this rank
    const int lda = K + 4, ldb = N + 8, NT = N / 8;
    double *Bs = sm;                       // [K][ldb]
    double *As = sm + (size_t)K * ldb;     // [TT_VT][lda]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    for (int i = tid; i < K * N; i += 256) {
        const int p = i / N, j = i - p * N;
        Bs[p * ldb + j] = Bg[i];
    }
    pdl_wait();                                            // Q is constant; the tiles are the previous launch's output
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int v0 = tile * TT_VT;
        __syncthreads();
        // tile load in batches of 8 independent global loads per thread (all in flight before the first smem store)
        for (int i0 = tid; i0 < K * TT_VT; i0 += 256 * 8) {
            double val[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * 256;
                val[u] = 0.0;
                if (i < K * TT_VT) {
                    if (DIR == 0) {
                        const int p = i / TT_VT, vv = i - p * TT_VT;
                        if (p < nt1 && v0 + vv < V) val[u] = c.rhs[(size_t)p * V + v0 + vv];
                    } else {
                        const int vv = i / K, p = i - vv * K;
                        if (v0 + vv < V) {
                            if (c.n_ranks == 1) {                   // K == M: the tile is a contiguous run of `hat`
                                val[u] = c.hat[(size_t)(v0 + vv) * M + p];
                            } else {
                                const int rk = p / M, pos = p - rk * M;     // gathered layout [rank][v][m_pad]
                                val[u] = c.peer_hat[0] ? c.peer_hat[rk][(size_t)(v0 + vv) * M + pos]      // peer memory (NVLink loads)
                                                       : c.hat_all[((size_t)rk * V + v0 + vv) * M + pos];
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * 256;
                if (i < K * TT_VT) {
                    if (DIR == 0) { const int p = i / TT_VT, vv = i - p * TT_VT; As[vv * lda + p] = val[u]; }
                    else { const int vv = i / K, p = i - vv * K; As[vv * lda + p] = val[u]; }
                }
            }
        }
        __syncthreads();
        double acc[16][2];
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        const double *ap = As + (8 * warp + (lane >> 2)) * lda + (lane & 3);
        const double *bp = Bs + (lane & 3) * ldb + (lane >> 2);
        for (int p0 = 0; p0 < K; p0 += 4) {
            const double a = ap[p0];
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) {
                if (nt < NT) dmma_m8n8k4(acc[nt][0], acc[nt][1], a, bp[(size_t)p0 * ldb + nt * 8]);
            }
        }
        const int vrow = v0 + 8 * warp + (lane >> 2);
        if (vrow < V) {
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) {
                if (nt < NT) {
                    const int j = nt * 8 + 2 * (lane & 3);
                    if (DIR == 0) {
                        *reinterpret_cast<double2 *>(c.hat + (size_t)vrow * M + j) = make_double2(acc[nt][0], acc[nt][1]);
                    } else {
                        if (j < c.tt_nout) c.phi[(size_t)(c.lvl_begin + j) * V + vrow] = acc[nt][0];
                        if (j + 1 < c.tt_nout) c.phi[(size_t)(c.lvl_begin + j + 1) * V + vrow] = acc[nt][1];
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same transforms with HALF the tensor work, for an even number of time levels n on one GPU.  The time eigenbasis is the
// DCT-II basis, Q[t][k] = c_k cos(pi k (t + 1/2) / n), so Q[n-1-t][k] = (-1)^k Q[t][k].  With the modes stored even-first
// (the engine permutes them: [k = 0, 2, 4, ... | 1, 3, 5, ...], tt_sym = 1) both GEMMs split into two of half the depth:
//   DIR 0:  hat[v][even] = sum_{t < n/2} (rhs[t] + rhs[n-1-t]) Q[t][even],   hat[v][odd] = sum_{t < n/2} (rhs[t] - rhs[n-1-t]) Q[t][odd]
//   DIR 1:  E[t] = sum_even hat Q[t][even], O[t] = sum_odd hat Q[t][odd]  (t < n/2);   phi[t] = E + O,  phi[n-1-t] = E - O
// (ncu on k_time_mma: tensor pipe 71 % / 36 % busy at 0.14 ms each: the transforms were DMMA bound, not HBM bound).
template <int DIR>
__global__ void __launch_bounds__(256) k_time_sym(dots_ctx_t c, int n_tiles)
{
    extern __shared__ double sm[];
    const int n = c.n_time + 1, h = n / 2, M = c.m_pad, V = c.n_vert;      // M == n (checked by the host)
    const int NT = n / 8, HT = NT / 2;                                     // 8-column tiles: HT per half
    const double *Bg = (DIR == 0) ? c.qf : c.qb;                           // qf [tt_kf][M]: rows t;  qb [tt_kb][tt_nb]: rows = modes, columns t
    const int ldg = (DIR == 0) ? M : c.tt_nb;
    const int lda = n + 4, ldb = n + 8;
    double *Bs = sm;                                                       // DIR 0: [h][ldb] rows t < h; DIR 1: [n][ldb] rows = modes, columns t < h used
    const int b_rows = (DIR == 0) ? h : n;
    double *As = sm + (size_t)b_rows * ldb;                                // [TT_VT][lda]: DIR 0: [sum | difference] halves; DIR 1: hat row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    for (int i = tid; i < b_rows * n; i += 256) {
        const int p = i / n, j = i - p * n;
        Bs[p * ldb + j] = (DIR == 0 || j < c.tt_nb) ? Bg[(size_t)p * ldg + j] : 0.0;
    }
    pdl_wait();                                                            // Q is constant; the tiles are the previous launch's output
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int v0 = tile * TT_VT;
        __syncthreads();
        if (DIR == 0) {                                                    // pairs (t, n-1-t): 8 pairs per thread in flight
            for (int i0 = tid; i0 < h * TT_VT; i0 += 256 * 8) {
                double lo[8], hi[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * 256;
                    lo[u] = hi[u] = 0.0;
                    if (i < h * TT_VT) {
                        const int p = i / TT_VT, vv = i - p * TT_VT;
                        if (v0 + vv < V) { lo[u] = c.rhs[(size_t)p * V + v0 + vv]; hi[u] = c.rhs[(size_t)(n - 1 - p) * V + v0 + vv]; }
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * 256;
                    if (i < h * TT_VT) {
                        const int p = i / TT_VT, vv = i - p * TT_VT;
                        As[vv * lda + p] = lo[u] + hi[u];
                        As[vv * lda + h + p] = lo[u] - hi[u];
                    }
                }
            }
        } else {
            for (int i0 = tid; i0 < n * TT_VT; i0 += 256 * 8) {
                double val[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * 256;
                    val[u] = 0.0;
                    if (i < n * TT_VT) {
                        const int vv = i / n, p = i - vv * n;
                        if (v0 + vv < V) val[u] = c.hat[(size_t)(v0 + vv) * M + p];
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * 256;
                    if (i < n * TT_VT) { const int vv = i / n, p = i - vv * n; As[vv * lda + p] = val[u]; }
                }
            }
        }
        __syncthreads();
        double acc[16][2];
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        const double *ap = As + (8 * warp + (lane >> 2)) * lda + (lane & 3);
        const double *bp = Bs + (lane & 3) * ldb + (lane >> 2);
        for (int p0 = 0; p0 < h; p0 += 4) {
            const double a0 = ap[p0], a1 = ap[h + p0];                     // DIR 0: sum / difference;  DIR 1: even / odd modes
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                if (nt < HT) {
                    if (DIR == 0) {
                        dmma_m8n8k4(acc[nt][0], acc[nt][1], a0, bp[(size_t)p0 * ldb + nt * 8]);                    // even modes
                        dmma_m8n8k4(acc[8 + nt][0], acc[8 + nt][1], a1, bp[(size_t)p0 * ldb + (HT + nt) * 8]);     // odd modes
                    } else {
                        dmma_m8n8k4(acc[nt][0], acc[nt][1], a0, bp[(size_t)p0 * ldb + nt * 8]);                    // E[t]
                        dmma_m8n8k4(acc[8 + nt][0], acc[8 + nt][1], a1, bp[(size_t)(h + p0) * ldb + nt * 8]);      // O[t]
                    }
                }
            }
        }
        const int vrow = v0 + 8 * warp + (lane >> 2);
        if (vrow < V) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                if (nt < HT) {
                    const int j = nt * 8 + 2 * (lane & 3);
                    if (DIR == 0) {
                        *reinterpret_cast<double2 *>(c.hat + (size_t)vrow * M + j) = make_double2(acc[nt][0], acc[nt][1]);
                        *reinterpret_cast<double2 *>(c.hat + (size_t)vrow * M + h + j) = make_double2(acc[8 + nt][0], acc[8 + nt][1]);
                    } else {
                        c.phi[(size_t)j * V + vrow] = acc[nt][0] + acc[8 + nt][0];
                        c.phi[(size_t)(j + 1) * V + vrow] = acc[nt][1] + acc[8 + nt][1];
                        c.phi[(size_t)(n - 1 - j) * V + vrow] = acc[nt][0] - acc[8 + nt][0];
                        c.phi[(size_t)(n - 2 - j) * V + vrow] = acc[nt][1] - acc[8 + nt][1];
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t panel_row_off(int row, int s)
{
    return (row < s) ? (size_t)row * (row + 1) / 2 : (size_t)s * (s + 1) / 2 + (size_t)(row - s) * s;
}

// optional profiling: first thread of the first block records %globaltimer at kernel start (tools/level_times.py)
__device__ __forceinline__ void sweep_stamp(const dots_ctx_t &c, int idx)
{
    if (c.phase_clock && idx >= 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        c.phase_clock[idx] = t;
    }
}

// Gather step of the forward sweep, one tree level: r_S = hat_S + (children's updates landing on S), in place.
// Block = one item (node, first S row, n rows) of the level's gather list (leaves have no children: no items).
__global__ void __launch_bounds__(256) k_sweep_gather(dots_ctx_t c, int item0, int stamp)
{
    pdl_launch_dependents();
    sweep_stamp(c, stamp);
    const int M = c.m_pad;
    const int *it = c.lvn_nodes + 3 * (size_t)(item0 + blockIdx.x);
    const int node = it[0], j0 = it[1], nj = it[2];
    const int off = c.nd_off[node];
    const int ch0 = c.nd_child[2 * node], ch1 = c.nd_child[2 * node + 1];
    const double *u0 = (ch0 >= 0) ? c.upd + (size_t)c.nd_upd[ch0] * M : nullptr;
    const double *u1 = (ch1 >= 0) ? c.upd + (size_t)c.nd_upd[ch1] * M : nullptr;
    const int32_t *cp0 = c.child_pos + c.nd_front[node];
    const int32_t *cp1 = cp0 + c.front_total;
    pdl_wait();
    for (int i = threadIdx.x; i < nj * M; i += blockDim.x) {
        const int j = j0 + i / M, m = i % M;
        const int a = cp0[j], b = cp1[j];
        double r = c.hat[(size_t)(off + j) * M + m];
        if (u0 && a >= 0) r += u0[(size_t)a * M + m];
        if (u1 && b >= 0) r += u1[(size_t)b * M + m];
        c.hat[(size_t)(off + j) * M + m] = r;
    }
}

__device__ __forceinline__ size_t panel_col_off(int col, int s, int b)     // column-major copy: offset of column `col`
{
    return (size_t)col * (s + b) - (size_t)col * (col - 1) / 2;
}
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes)     // bytes % 16 == 0; SASS: UBLKPF.L2
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// One tree level of a sweep.  Block = one work item (node, first output, n outputs).
//   DIR 0, forward : output = panel row i of the row-major panel  (columns [0, min(i+1,s)))
//        y_S   = inv(L11) r_S                                               -> ywork
//        upd_B = (children's updates landing on B) - (L21 inv(L11)) r_S     -> this node's update vector
//   DIR 1, backward: output = panel column j of the column-major copy (rows [j, s+b))
//        xt_S = P^T (-[y_S ; xt_B])      (x of the ancestors is final in `hat`; the reference's per-mode matrix is
//        L + (lambda - eps) M = -(K + shift M), laplacian_inverse_socp.py:37-38, so what is stored is xt = -x)
// In both directions an output's data is ONE contiguous run of len*M doubles.  WPR warps share an output (interleaved
// entries, partial sums combined through shared memory in a fixed order), 8/WPR outputs are in flight per pass, and
// the run of the pass after next is pulled into L2 with one bulk prefetch (cp.async.bulk.prefetch.L2) so that the
// demand loads mostly see L2 latency instead of DRAM latency (ncu before: >80 % long_scoreboard).
// ML = modes of this rank (8, 16, 32, 64, 96, 128).  ML >= 32: lane = mode (+32 per register); ML < 32: a warp covers
// G = 32/ML consecutive entries at once (lane = entry-in-group * ML + mode) and folds the groups with shuffles.
// FG (forward only): the block folds the children's updates into r_S itself and keeps r_S in shared memory (levels
// whose separators have <= SWEEP_FG_SMAX rows), which removes the separate gather launch of the level.
#define SWEEP_FG_SMAX 64
#ifndef SWEEP_MIN_BLOCKS
#define SWEEP_MIN_BLOCKS 1
#endif
template <int ML, int WPR, int DIR, bool FG>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS) k_sweep_run(dots_ctx_t c, int item0, int stamp)
{
    pdl_launch_dependents();
    sweep_stamp(c, stamp);
    extern __shared__ double rsm[];                    // [s][M] staged r_S (FG only)
    constexpr int M = ML;
    constexpr int MP = (ML >= 32) ? ML / 32 : 1;
    constexpr int G = (ML >= 32) ? 1 : 32 / ML;
    constexpr int LSTRIDE = (ML >= 32) ? 32 : 0;       // register m of a lane is mode lane + 32 m (only ML >= 32)
    constexpr int ROWS = SWEEP_WARPS / WPR;
    // measured: two outputs per warp or 8-deep unrolling cost more in occupancy (32-48 -> 90-114 registers) than they
    // gain in loads in flight, so one output at a time and 4 entries in flight (profiles/README.md)
    constexpr int NR = 1;                              // outputs a warp works on at a time
    constexpr int UN = 4;                              // entries per output in flight
    __shared__ double red[(WPR > 1) ? SWEEP_WARPS * M : 1];
    const int *it = (DIR == 0 ? c.lvl_items : c.lvb_items) + 3 * (size_t)(item0 + blockIdx.x);
    const int node = it[0], o0 = it[1], n_o = it[2];
    const int s = c.nd_s[node], off = c.nd_off[node], b_rows = c.nd_b[node];
    const int ch0 = c.nd_child[2 * node], ch1 = c.nd_child[2 * node + 1];
    const double *u0 = (DIR == 0 && ch0 >= 0) ? c.upd + (size_t)c.nd_upd[ch0] * M : nullptr;
    const double *u1 = (DIR == 0 && ch1 >= 0) ? c.upd + (size_t)c.nd_upd[ch1] * M : nullptr;
    const int32_t *cp0 = c.child_pos + c.nd_front[node];
    const int32_t *cp1 = cp0 + c.front_total;
    const int32_t *fidx = c.front_idx + c.nd_front[node];
    const double *panel = (DIR == 0 ? c.panels : c.panels_t) + (size_t)c.nd_panel[node] * M;
    double *myupd = c.upd + (size_t)c.nd_upd[node] * M;
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = (G > 1) ? lane_id / ML : 0;        // entry slot inside the warp
    const int lane = (G > 1) ? lane_id % ML : lane_id; // mode index of register 0
    const int rslot = warp / WPR, cslot = warp % WPR;
    const int estep = WPR * G, e_first = cslot * G + grp;
    const int last = o0 + n_o - 1;

    auto out_len = [&](int o) { return DIR == 0 ? min(o + 1, s) : s + b_rows - o; };
    auto out_off = [&](int o) { return DIR == 0 ? panel_row_off(o, s) : panel_col_off(o, s, b_rows); };
    auto prefetch = [&](int o) {
        if (cslot == 0 && lane_id == 0 && o <= last) {
            const int len = out_len(o);
            if (len > 0) l2_prefetch_bulk(panel + out_off(o) * M, (uint32_t)len * M * 8u);
        }
    };
#pragma unroll
    for (int k = 0; k < 2 * NR; ++k) prefetch(o0 + k * ROWS + rslot);
    pdl_wait();                                        // below: vectors written by the previous launches

    if (FG) {                                          // r_S = hat_S + children's updates landing on S
        const int ncol = min(s, last + 1);
        for (int i = threadIdx.x; i < ncol * M; i += SWEEP_THREADS) {
            const int j = i / M, m = i - j * M;
            double r = c.hat[(size_t)(off + j) * M + m];
            const int a = cp0[j], b = cp1[j];
            if (u0 && a >= 0) r += u0[(size_t)a * M + m];
            if (u1 && b >= 0) r += u1[(size_t)b * M + m];
            rsm[i] = r;
        }
        __syncthreads();
    }

    for (int base = o0; base <= last; base += ROWS * NR) {
        int o[NR], len[NR];
        const double *pr[NR];
        int lenmax = 0;
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            o[q] = base + q * ROWS + rslot;
            len[q] = (o[q] <= last) ? out_len(o[q]) : 0;
            pr[q] = panel + ((o[q] <= last) ? out_off(o[q]) : 0) * M + lane;
            lenmax = max(lenmax, len[q]);
            prefetch(o[q] + 2 * NR * ROWS);
        }
        double acc[NR][MP];
#pragma unroll
        for (int q = 0; q < NR; ++q)
#pragma unroll
            for (int m = 0; m < MP; ++m) acc[q][m] = 0.0;
        auto vptr = [&](int q, int e) -> const double * {
            if (DIR == 0) return (FG ? rsm + (size_t)e * M : c.hat + (size_t)(off + e) * M) + lane;
            const int row = o[q] + e;
            return (row < s ? c.ywork + (size_t)(off + row) * M : c.hat + (size_t)fidx[row] * M) + lane;
        };
        for (int e = e_first; e < lenmax; e += UN * estep) {
            double p[NR][UN][MP], r[NR][UN][MP];
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int q = 0; q < NR; ++q) {
                    const int ee = e + u * estep;
                    const bool on = ee < len[q];
                    const double *vp = on ? vptr(q, ee) : nullptr;
#pragma unroll
                    for (int m = 0; m < MP; ++m) {
                        p[q][u][m] = on ? __ldcs(pr[q] + (size_t)ee * M + LSTRIDE * m) : 0.0;
                        r[q][u][m] = on ? vp[LSTRIDE * m] : 0.0;
                    }
                }
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int q = 0; q < NR; ++q)
#pragma unroll
                    for (int m = 0; m < MP; ++m) acc[q][m] += p[q][u][m] * r[q][u][m];
        }
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const bool valid = o[q] <= last;
            if (G > 1) {                                       // fold the entry slots of the warp (fixed order)
#pragma unroll
                for (int o2 = ML; o2 < 32; o2 <<= 1) acc[q][0] += __shfl_xor_sync(0xffffffffu, acc[q][0], o2);
            }
            const bool writer = (G > 1) ? (grp == 0) : true;
            if (WPR > 1) {
                __syncthreads();
#pragma unroll
                for (int m = 0; m < MP; ++m) if (writer) red[warp * M + LSTRIDE * m + lane] = acc[q][m];
                __syncthreads();
                if (cslot == 0) {
#pragma unroll
                    for (int m = 0; m < MP; ++m) {
                        double v = 0.0;
#pragma unroll
                        for (int w = 0; w < WPR; ++w) v += red[(rslot * WPR + w) * M + LSTRIDE * m + lane];
                        acc[q][m] = v;
                    }
                }
            }
            if (valid && cslot == 0 && writer) {
                const int oo = o[q];
                if (DIR == 1) {
#pragma unroll
                    for (int m = 0; m < MP; ++m) c.hat[(size_t)(off + oo) * M + LSTRIDE * m + lane] = -acc[q][m];
                } else if (oo < s) {
#pragma unroll
                    for (int m = 0; m < MP; ++m) c.ywork[(size_t)(off + oo) * M + LSTRIDE * m + lane] = acc[q][m];
                } else if (oo - s < b_rows) {
                    const int a = cp0[oo], b = cp1[oo];
#pragma unroll
                    for (int m = 0; m < MP; ++m) {
                        double val = 0.0;
                        if (u0 && a >= 0) val += u0[(size_t)a * M + LSTRIDE * m + lane];
                        if (u1 && b >= 0) val += u1[(size_t)b * M + LSTRIDE * m + lane];
                        myupd[(size_t)(oo - s) * M + LSTRIDE * m + lane] = val - acc[q][m];
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <int ML, int DIR>
static int launch_level(const dots_ctx_t *c, int wpr_code, int i0, int n, int stamp, bool pdl, cudaStream_t st)
{
    const bool fuse = (DIR == 0) && (wpr_code & 16);
    const size_t smem = fuse ? (size_t)SWEEP_FG_SMAX * ML * sizeof(double) : 0;
    if (fuse) {
        // dynamic r_S staging + the static combine buffer of the WPR = 2 instantiation must fit: opt in whenever the sum can
        // exceed the 48 KB default (ML = 96: 48 KB + 6 KB), once per device
        static bool configured[64] = {false};
        int dev = 0;
        DOTS_CUDA(cudaGetDevice(&dev));
        if (smem + (size_t)SWEEP_WARPS * ML * sizeof(double) > 48 * 1024 && dev >= 0 && dev < 64 && !configured[dev]) {
            DOTS_CUDA(cudaFuncSetAttribute(k_sweep_run<ML, 1, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            DOTS_CUDA(cudaFuncSetAttribute(k_sweep_run<ML, 2, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[dev] = true;
        }
        switch (wpr_code & 15) {
        case 1: return pdl_launch(k_sweep_run<ML, 1, 0, true>, n, SWEEP_THREADS, smem, st, pdl, *c, i0, stamp);
        default: return pdl_launch(k_sweep_run<ML, 2, 0, true>, n, SWEEP_THREADS, smem, st, pdl, *c, i0, stamp);
        }
    } else {
        switch (wpr_code & 15) {
        case 1: return pdl_launch(k_sweep_run<ML, 1, DIR, false>, n, SWEEP_THREADS, (size_t)0, st, pdl, *c, i0, stamp);
        case 2: return pdl_launch(k_sweep_run<ML, 2, DIR, false>, n, SWEEP_THREADS, (size_t)0, st, pdl, *c, i0, stamp);
        case 4: return pdl_launch(k_sweep_run<ML, 4, DIR, false>, n, SWEEP_THREADS, (size_t)0, st, pdl, *c, i0, stamp);
        default: return pdl_launch(k_sweep_run<ML, 8, DIR, false>, n, SWEEP_THREADS, (size_t)0, st, pdl, *c, i0, stamp);
        }
    }
}

template <int ML>
static int launch_sweeps(const dots_ctx_t *c, cudaStream_t st)
{
    const bool pdl = c->ring_pdl != 0;
    bool chain = true;                                 // the first launch chains to the time transform
    for (int lv = 0; lv < c->n_levels; ++lv) {
        const int g0 = c->h_lvn_ptr[lv], gn = c->h_lvn_ptr[lv + 1] - g0;
        // stamps: slot lv = start of forward level lv (its gather when there is one), slot L + k = start of the k-th backward level
        const bool fused = (c->h_lvl_wpr[lv] & 16) != 0;      // the level's blocks fold the children's updates themselves
        const bool gather = lv > 0 && gn > 0 && !fused;
        if (gather) {
            if (int e = pdl_launch(k_sweep_gather, gn, 256, (size_t)0, st, pdl && chain, *c, g0, lv)) return e;
            chain = true;
        }
        const int i0 = c->h_lvl_ptr[lv], n = c->h_lvl_ptr[lv + 1] - i0;
        if (n <= 0) continue;
        if (int e = launch_level<ML, 0>(c, c->h_lvl_wpr[lv], i0, n, gather ? -1 : lv, pdl && chain, st)) return e;
        chain = true;
    }
    for (int lv = c->n_levels - 1; lv >= 0; --lv) {
        const int i0 = c->h_lvb_ptr[lv], n = c->h_lvb_ptr[lv + 1] - i0;
        if (n <= 0) continue;
        if (int e = launch_level<ML, 1>(c, c->h_lvb_cw[lv], i0, n, 2 * c->n_levels - 1 - lv, pdl && chain, st)) return e;
        chain = true;
    }
    return 0;
}

int dots_mode_solves_ring(const dots_ctx_t *c, void *stream);         // sweep_ring.cu

extern "C" int dots_mode_solves(const dots_ctx_t *c, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    if (c->sweep_mode == 4) return dots_mode_solves_ring(c, stream);
    if (c->sweep_mode != 0) { dots_set_error("sweep_mode=%d unsupported (0 or 4)", c->sweep_mode); return DOTS_ERR_BAD_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    switch (c->m_pad) {
    case 8: return launch_sweeps<8>(c, st);
    case 16: return launch_sweeps<16>(c, st);
    case 32: return launch_sweeps<32>(c, st);
    case 64: return launch_sweeps<64>(c, st);
    case 96: return launch_sweeps<96>(c, st);
    case 128: return launch_sweeps<128>(c, st);
    }
    dots_set_error("m_pad=%d unsupported", c->m_pad);
    return DOTS_ERR_BAD_ARG;
}

extern "C" int dots_time_transform(const dots_ctx_t *c, int inverse, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    if (c->tt_sym) {                                                       // even level count on one GPU, modes stored even-first
        const int n = c->n_time + 1;
        if (c->n_ranks != 1 || n % 16 || n != c->m_pad || n > 128 || c->tt_nout != n || c->lvl_begin != 0) { dots_set_error("tt_sym needs one rank and n_time + 1 = m_pad in {16, 32, ..., 128}"); return DOTS_ERR_BAD_ARG; }
        const size_t smem_s = ((size_t)(inverse ? n : n / 2) * (n + 8) + (size_t)TT_VT * (n + 4)) * sizeof(double);
        const int tiles = ceil_div(c->n_vert, TT_VT);
        const int per = (smem_s <= 72 * 1024) ? 3 : (smem_s <= 110 * 1024 ? 2 : 1);
        const int grid_s = tiles < c->n_sm * per ? tiles : c->n_sm * per;
        static size_t conf[64][2] = {{0, 0}};
        int dv = 0;
        DOTS_CUDA(cudaGetDevice(&dv));
        dv = (dv >= 0 && dv < 64) ? dv : 0;
        if (smem_s > 48 * 1024 && smem_s > conf[dv][inverse ? 1 : 0]) {
            if (inverse) DOTS_CUDA(cudaFuncSetAttribute(k_time_sym<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
            else DOTS_CUDA(cudaFuncSetAttribute(k_time_sym<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
            conf[dv][inverse ? 1 : 0] = smem_s;
        }
        const bool pdl_s = c->ring_pdl != 0;
        if (inverse) return pdl_launch(k_time_sym<1>, grid_s, 256, smem_s, st, pdl_s, *c, tiles);
        return pdl_launch(k_time_sym<0>, grid_s, 256, smem_s, st, pdl_s, *c, tiles);
    }
    const int K = inverse ? c->tt_kb : c->tt_kf, N = inverse ? c->tt_nb : c->m_pad;
    if (K % 4 || N % 8 || N > 128 || K <= 0) { dots_set_error("time transform shape K=%d N=%d unsupported", K, N); return DOTS_ERR_BAD_ARG; }
    const size_t smem = ((size_t)K * (N + 8) + (size_t)TT_VT * (K + 4)) * sizeof(double);
    const int n_tiles = ceil_div(c->n_vert, TT_VT);
    const int per_sm = (smem <= 72 * 1024) ? 3 : (smem <= 110 * 1024 ? 2 : 1);
    // a small B (sharded ranks: few modes / levels) is cheap to reload, so one tile per block lets the hardware overlap
    // the tile loads of different blocks; the single-GPU B (36-140 KB) is loaded once per persistent block instead
    const bool small_b = (size_t)K * N * sizeof(double) <= 16 * 1024;
    const int grid = (small_b || n_tiles < c->n_sm * per_sm) ? n_tiles : c->n_sm * per_sm;
    static size_t configured[64][2] = {{0, 0}};                    // per device: the attribute belongs to the device's copy of the function
    int dev = 0;
    DOTS_CUDA(cudaGetDevice(&dev));
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    if (smem > 48 * 1024 && smem > configured[dev][inverse ? 1 : 0]) {
        if (inverse) DOTS_CUDA(cudaFuncSetAttribute(k_time_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else DOTS_CUDA(cudaFuncSetAttribute(k_time_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev][inverse ? 1 : 0] = smem;
    }
    const bool pdl = c->ring_pdl != 0;
    if (inverse) return pdl_launch(k_time_mma<1>, grid, 256, smem, st, pdl, *c, n_tiles);
    return pdl_launch(k_time_mma<0>, grid, 256, smem, st, pdl, *c, n_tiles);
}

// ------------------------------------------------------------------------------------------------
// Plain (CUDA-core) time transforms for time grids beyond the tensor-core kernels' shared-memory budget (n_time + 1 > 128).
// The modes are then solved in groups of m_pad <= 128 (the engine loops: transform -> sweeps -> inverse transform per
// group), so one call handles ONE group:
//   forward  hat[v][j]     = sum_t q[t][j] rhs[t][v]          q: [n_time + 1][m_pad], this group's columns of the DCT basis
//   inverse  phi[t][v] (+)= sum_j q[j][t] hat[v][j]           q: [m_pad][n_time + 1]; accumulate: add to phi (groups after the first)
// Block = 32 vertices; q is staged through shared memory in chunks of 32 rows / columns; sums run in index order.
#define TP_VT 32
#define TP_KC 32
template <int DIR>
__global__ void __launch_bounds__(256) k_time_plain(dots_ctx_t c, const double *__restrict__ q, int accumulate)
{
    extern __shared__ double tp_sm[];
    const int n = c.n_time + 1, M = c.m_pad, V = c.n_vert;
    const int v0 = blockIdx.x * TP_VT;
    const int tid = threadIdx.x, vv = tid & 31, grp = tid >> 5;          // 8 groups of 32 lanes
    if (DIR == 0) {
        double *As = tp_sm;                                              // [TP_KC][TP_VT]   rhs chunk
        double *Qs = tp_sm + TP_KC * TP_VT;                              // [TP_KC][M]       q chunk; reused as the output tile [TP_VT][M]
        double acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.0;
        for (int t0 = 0; t0 < n; t0 += TP_KC) {
            const int kc = min(TP_KC, n - t0);
            __syncthreads();
            for (int i = tid; i < kc * TP_VT; i += 256) {
                const int p = i >> 5, w = i & 31;
                As[i] = (v0 + w < V) ? c.rhs[(size_t)(t0 + p) * V + v0 + w] : 0.0;
            }
            for (int i = tid; i < kc * M; i += 256) Qs[i] = q[(size_t)t0 * M + i];
            __syncthreads();
            for (int p = 0; p < kc; ++p) {
                const double a = As[p * TP_VT + vv];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (grp + 8 * i < M) acc[i] += a * Qs[p * M + grp + 8 * i];
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (grp + 8 * i < M) Qs[vv * M + grp + 8 * i] = acc[i];
        __syncthreads();
        for (int i = tid; i < TP_VT * M; i += 256) {
            const int w = i / M;
            if (v0 + w < V) c.hat[(size_t)v0 * M + i] = Qs[i];
        }
    } else {
        double *Hs = tp_sm;                                              // [TP_VT][M + 1]   hat rows of the block's vertices
        double *Qs = tp_sm + TP_VT * (M + 1);                            // [M][TP_KC]       q chunk (columns t0 .. t0 + 31)
        for (int i = tid; i < TP_VT * M; i += 256) {
            const int w = i / M, j = i - w * M;
            Hs[w * (M + 1) + j] = (v0 + w < V) ? c.hat[(size_t)v0 * M + i] : 0.0;
        }
        for (int t0 = 0; t0 < n; t0 += TP_KC) {
            const int kc = min(TP_KC, n - t0);
            __syncthreads();
            for (int i = tid; i < M * TP_KC; i += 256) {
                const int j = i >> 5, p = i & 31;
                Qs[i] = (p < kc) ? q[(size_t)j * n + t0 + p] : 0.0;
            }
            __syncthreads();
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int j = 0; j < M; ++j) {
                const double h = Hs[vv * (M + 1) + j];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] += h * Qs[j * TP_KC + grp + 8 * i];
            }
            if (v0 + vv < V) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int t = t0 + grp + 8 * i;
                    if (t < n) {
                        double *dst = c.phi + (size_t)t * V + v0 + vv;
                        *dst = accumulate ? *dst + acc[i] : acc[i];
                    }
                }
            }
        }
    }
}

extern "C" int dots_time_transform_plain(const dots_ctx_t *c, int inverse, const double *q, int accumulate, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    if (!q) { dots_set_error("dots_time_transform_plain: null basis"); return DOTS_ERR_BAD_ARG; }
    if (c->n_ranks != 1 || c->lvl_begin != 0 || c->lvl_end != c->n_time + 1) { dots_set_error("dots_time_transform_plain is single-GPU (whole time grid)"); return DOTS_ERR_BAD_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const int M = c->m_pad, grid = ceil_div(c->n_vert, TP_VT);
    const size_t smem = inverse ? ((size_t)TP_VT * (M + 1) + (size_t)M * TP_KC) * sizeof(double)
                                : ((size_t)TP_KC * TP_VT + (size_t)(TP_KC > TP_VT ? TP_KC : TP_VT) * M) * sizeof(double);
    static size_t configured[64][2] = {{0, 0}};
    int dev = 0;
    DOTS_CUDA(cudaGetDevice(&dev));
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    if (smem > 48 * 1024 && smem > configured[dev][inverse ? 1 : 0]) {
        if (inverse) DOTS_CUDA(cudaFuncSetAttribute(k_time_plain<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else DOTS_CUDA(cudaFuncSetAttribute(k_time_plain<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev][inverse ? 1 : 0] = smem;
    }
    if (inverse) k_time_plain<1><<<grid, 256, smem, st>>>(*c, q, accumulate);
    else k_time_plain<0><<<grid, 256, smem, st>>>(*c, q, accumulate);
    DOTS_LAUNCH_CHECK();
    return 0;
}

extern "C" int dots_step_phi(const dots_ctx_t *c, void *stream)
{
    int e;
    if ((e = dots_phi_rhs(c, stream))) return e;
    if ((e = dots_time_transform(c, 0, stream))) return e;
    if ((e = dots_mode_solves(c, stream))) return e;
    return dots_time_transform(c, 1, stream);
}

extern "C" int dots_iterate(const dots_ctx_t *c, int n_iter, int write_z, void *stream)
{
    for (int i = 0; i < n_iter; ++i) {
        int e;
        if ((e = dots_step_phi(c, stream))) return e;
        if ((e = dots_step_vertex(c, stream))) return e;
        if ((e = dots_step_tri(c, (i == n_iter - 1) ? write_z : 0, stream))) return e;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// CUDA-graph form of one iteration: the ~45 launches of dots_iterate(1) are captured once and replayed with a
// single cudaGraphLaunch (launch latency matters for the small meshes: knots_5-class is ~60 us of HBM traffic).
// Scalars (r, s, d, ...) live in device memory (ctx->params), so the graph stays valid across penalty updates.
struct dots_graph {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
};

extern "C" int dots_graph_create(const dots_ctx_t *c, int write_z, void *stream, dots_graph_t **out)
{
    if (int e = dots_check_ctx(c)) return e;
    if (!out) { dots_set_error("null output handle"); return DOTS_ERR_BAD_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    dots_graph *g = new dots_graph{nullptr, nullptr};
    DOTS_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int e = dots_iterate(c, 1, write_z, stream);
    cudaError_t ce = cudaStreamEndCapture(st, &g->graph);
    if (e || ce != cudaSuccess) {
        if (!e) dots_set_error("cudaStreamEndCapture -> %s", cudaGetErrorString(ce));
        delete g;
        return e ? e : (int)ce;
    }
    DOTS_CUDA(cudaGraphInstantiate(&g->exec, g->graph, 0));
    *out = g;
    return 0;
}

extern "C" int dots_graph_launch(dots_graph_t *g, void *stream)
{
    if (!g || !g->exec) { dots_set_error("null graph"); return DOTS_ERR_BAD_ARG; }
    DOTS_CUDA(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
    return 0;
}

extern "C" int dots_graph_destroy(dots_graph_t *g)
{
    if (!g) return 0;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return 0;
}
