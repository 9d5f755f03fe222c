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
// forward transform: hat[v][k] = sum_t Q[t][k] rhs[t][v]
#define TT_TILE 32
__global__ void __launch_bounds__(256) k_time_fwd(dots_ctx_t c)
{
    extern __shared__ double sm[];
    const int nt1 = c.n_time + 1, M = c.m_pad, V = c.n_vert;
    double *Qs = sm;                       // [nt1][M]
    double *Xs = sm + (size_t)nt1 * M;     // [nt1][TT_TILE]
    const int v0 = blockIdx.x * TT_TILE;
    for (int i = threadIdx.x; i < nt1 * M; i += blockDim.x) Qs[i] = c.qmat[i];
    for (int i = threadIdx.x; i < nt1 * TT_TILE; i += blockDim.x) {
        const int t = i / TT_TILE, vv = i % TT_TILE;
        Xs[i] = (v0 + vv < V) ? c.rhs[(size_t)t * V + v0 + vv] : 0.0;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < TT_TILE * M; o += blockDim.x) {
        const int vv = o / M, k = o % M;
        if (v0 + vv >= V) continue;
        double acc = 0.0;
        for (int t = 0; t < nt1; ++t) acc += Qs[t * M + k] * Xs[t * TT_TILE + vv];
        c.hat[(size_t)(v0 + vv) * M + k] = acc;
    }
}

// inverse transform: phi[t][v] = sum_k Q[t][k] hat[v][k]
__global__ void __launch_bounds__(256) k_time_bwd(dots_ctx_t c)
{
    extern __shared__ double sm[];
    const int nt1 = c.n_time + 1, M = c.m_pad, V = c.n_vert, Mp = M + 1;
    double *Qs = sm;                       // [nt1][M]
    double *Hs = sm + (size_t)nt1 * M;     // [TT_TILE][M+1]
    const int v0 = blockIdx.x * TT_TILE;
    for (int i = threadIdx.x; i < nt1 * M; i += blockDim.x) Qs[i] = c.qmat[i];
    for (int i = threadIdx.x; i < TT_TILE * M; i += blockDim.x) {
        const int vv = i / M, k = i % M;
        Hs[vv * Mp + k] = (v0 + vv < V) ? c.hat[(size_t)(v0 + vv) * M + k] : 0.0;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < nt1 * TT_TILE; o += blockDim.x) {
        const int t = o / TT_TILE, vv = o % TT_TILE;
        if (v0 + vv >= V) continue;
        double acc = 0.0;
        for (int k = 0; k < nt1; ++k) acc += Qs[t * M + k] * Hs[vv * Mp + k];
        c.phi[(size_t)t * V + v0 + vv] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t panel_row_off(int row, int s)
{
    return (row < s) ? (size_t)row * (row + 1) / 2 : (size_t)s * (s + 1) / 2 + (size_t)(row - s) * s;
}

// Gather step of the forward sweep, one tree level: r_S = hat_S + (children's updates landing on S), in place.
// Block = one node of the level (leaves have no children and are skipped by the host).
__global__ void __launch_bounds__(256) k_sweep_gather(dots_ctx_t c, int node0)
{
    const int M = c.m_pad;
    const int node = c.lvn_nodes[node0 + blockIdx.x];
    const int s = c.nd_s[node], off = c.nd_off[node];
    const int ch0 = c.nd_child[2 * node], ch1 = c.nd_child[2 * node + 1];
    const double *u0 = (ch0 >= 0) ? c.upd + (size_t)c.nd_upd[ch0] * M : nullptr;
    const double *u1 = (ch1 >= 0) ? c.upd + (size_t)c.nd_upd[ch1] * M : nullptr;
    const int32_t *cp0 = c.child_pos + c.nd_front[node];
    const int32_t *cp1 = cp0 + c.front_total;
    for (int i = threadIdx.x; i < s * M; i += blockDim.x) {
        const int j = i / M, m = i - j * M;
        const int a = cp0[j], b = cp1[j];
        double r = c.hat[(size_t)(off + j) * M + m];
        if (u0 && a >= 0) r += u0[(size_t)a * M + m];
        if (u1 && b >= 0) r += u1[(size_t)b * M + m];
        c.hat[(size_t)(off + j) * M + m] = r;
    }
}

// Forward sweep, one tree level.  Block = one work item (node, first row, n rows).
//   y_S   = inv(L11) r_S                                               -> ywork
//   upd_B = (children's updates landing on B) - (L21 inv(L11)) r_S     -> this node's update vector
// WPR warps share one panel row (interleaved columns, partial sums combined through shared memory in a fixed
// order); 8/WPR rows are in flight per pass.  Every panel entry is streamed exactly once, as M*8-byte rows.
template <int MP, int WPR>
__global__ void __launch_bounds__(SWEEP_THREADS) k_sweep_fwd(dots_ctx_t c, int item0)
{
    constexpr int M = 32 * MP;
    constexpr int ROWS = SWEEP_WARPS / WPR;
    __shared__ double red[(WPR > 1) ? SWEEP_WARPS * M : 1];
    const int *it = c.lvl_items + 3 * (size_t)(item0 + blockIdx.x);
    const int node = it[0], row0 = it[1], nrows = it[2];
    const int s = c.nd_s[node], off = c.nd_off[node], b_rows = c.nd_b[node];
    const int ch0 = c.nd_child[2 * node], ch1 = c.nd_child[2 * node + 1];
    const double *u0 = (ch0 >= 0) ? c.upd + (size_t)c.nd_upd[ch0] * M : nullptr;
    const double *u1 = (ch1 >= 0) ? c.upd + (size_t)c.nd_upd[ch1] * M : nullptr;
    const int32_t *cp0 = c.child_pos + c.nd_front[node];
    const int32_t *cp1 = cp0 + c.front_total;
    const double *panel = c.panels + (size_t)c.nd_panel[node] * M;
    const double *rvec = c.hat + (size_t)off * M;
    double *myupd = c.upd + (size_t)c.nd_upd[node] * M;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rslot = warp / WPR, cslot = warp % WPR;
    const int last_row = row0 + nrows - 1;

    for (int base = row0; base <= last_row; base += ROWS) {
        const int row = base + rslot;
        const bool valid = row <= last_row;
        double acc[MP];
#pragma unroll
        for (int m = 0; m < MP; ++m) acc[m] = 0.0;
        if (valid) {
            const int len = min(row + 1, s);
            const double *pr = panel + panel_row_off(row, s) * M + lane;
            const double *rv = rvec + lane;
            int j = cslot;
            for (; j + 3 * WPR < len; j += 4 * WPR) {
                double p[4][MP], r[4][MP];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int m = 0; m < MP; ++m) {
                        p[u][m] = __ldcs(pr + (size_t)(j + u * WPR) * M + 32 * m);
                        r[u][m] = rv[(size_t)(j + u * WPR) * M + 32 * m];
                    }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int m = 0; m < MP; ++m) acc[m] += p[u][m] * r[u][m];
            }
            for (; j < len; j += WPR) {
#pragma unroll
                for (int m = 0; m < MP; ++m) acc[m] += __ldcs(pr + (size_t)j * M + 32 * m) * rv[(size_t)j * M + 32 * m];
            }
        }
        if (WPR > 1) {
            __syncthreads();
#pragma unroll
            for (int m = 0; m < MP; ++m) red[warp * M + 32 * m + lane] = acc[m];
            __syncthreads();
            if (cslot == 0) {
#pragma unroll
                for (int m = 0; m < MP; ++m) {
                    double v = 0.0;
#pragma unroll
                    for (int w = 0; w < WPR; ++w) v += red[(rslot * WPR + w) * M + 32 * m + lane];
                    acc[m] = v;
                }
            }
        }
        if (valid && cslot == 0) {
            if (row < s) {
#pragma unroll
                for (int m = 0; m < MP; ++m) c.ywork[(size_t)(off + row) * M + 32 * m + lane] = acc[m];
            } else if (row - s < b_rows) {
                const int a = cp0[row], b = cp1[row];
#pragma unroll
                for (int m = 0; m < MP; ++m) {
                    double val = 0.0;
                    if (u0 && a >= 0) val += u0[(size_t)a * M + 32 * m + lane];
                    if (u1 && b >= 0) val += u1[(size_t)b * M + 32 * m + lane];
                    myupd[(size_t)(row - s) * M + 32 * m + lane] = val - acc[m];
                }
            }
        }
    }
}

// Backward sweep, one tree level.  Block = one work item (node, first column, n columns <= CW).
//   x_S = inv(L11)^T y_S - (L21 inv(L11))^T x_B          (x of the ancestors is already final in `hat`)
// The reference's per-mode matrix is L + (lambda - eps) M = -(K + shift M)  (laplacian_inverse_socp.py:37-38), so what
// is stored in `hat` is xt = -x; substituting gives  xt_S = P^T (-[y_S ; xt_B]):  the sign costs nothing.
// The 8 warps of a block interleave over the panel rows (each row contributes n_cols*M*8 contiguous bytes) and
// their partial column sums are combined through shared memory in a fixed order.
template <int MP, int CW>
__global__ void __launch_bounds__(SWEEP_THREADS) k_sweep_bwd(dots_ctx_t c, int item0)
{
    constexpr int M = 32 * MP;
    __shared__ double red[SWEEP_WARPS * CW * M];
    const int *it = c.lvb_items + 3 * (size_t)(item0 + blockIdx.x);
    const int node = it[0], col0 = it[1], ncols = it[2];
    const int s = c.nd_s[node], b = c.nd_b[node], off = c.nd_off[node];
    const int32_t *fidx = c.front_idx + c.nd_front[node];
    const double *panel = c.panels + (size_t)c.nd_panel[node] * M;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nrows = s + b;

    double acc[CW][MP];
#pragma unroll
    for (int q = 0; q < CW; ++q)
#pragma unroll
        for (int m = 0; m < MP; ++m) acc[q][m] = 0.0;

#pragma unroll 4
    for (int row = col0 + warp; row < nrows; row += SWEEP_WARPS) {     // rows < col0 never touch these columns
        const double *vsrc = (row < s) ? c.ywork + (size_t)(off + row) * M : c.hat + (size_t)fidx[row] * M;
        const double *pr = panel + (panel_row_off(row, s) + col0) * M + lane;
        const int qe = (row < s) ? min(ncols, row - col0 + 1) : ncols;    // lower triangle of inv(L11) only
        double v[MP];
#pragma unroll
        for (int m = 0; m < MP; ++m) v[m] = -vsrc[32 * m + lane];
#pragma unroll
        for (int q = 0; q < CW; ++q) {
            if (q < qe) {
#pragma unroll
                for (int m = 0; m < MP; ++m) acc[q][m] += __ldcs(pr + (size_t)q * M + 32 * m) * v[m];
            }
        }
    }
#pragma unroll
    for (int q = 0; q < CW; ++q)
#pragma unroll
        for (int m = 0; m < MP; ++m) red[(warp * CW + q) * M + 32 * m + lane] = acc[q][m];
    __syncthreads();
    for (int o = threadIdx.x; o < ncols * M; o += SWEEP_THREADS) {
        const int q = o / M, m = o - q * M;
        double vsum = 0.0;
#pragma unroll
        for (int w = 0; w < SWEEP_WARPS; ++w) vsum += red[(w * CW + q) * M + m];
        c.hat[(size_t)(off + col0 + q) * M + m] = vsum;
    }
}

// ------------------------------------------------------------------------------------------------
template <int MP>
static int launch_sweeps(const dots_ctx_t *c, cudaStream_t st)
{
    for (int lv = 0; lv < c->n_levels; ++lv) {
        const int g0 = c->h_lvn_ptr[lv], gn = c->h_lvn_ptr[lv + 1] - g0;
        if (lv > 0 && gn > 0) { k_sweep_gather<<<gn, 256, 0, st>>>(*c, g0); DOTS_LAUNCH_CHECK(); }
        const int i0 = c->h_lvl_ptr[lv], n = c->h_lvl_ptr[lv + 1] - i0;
        if (n <= 0) continue;
        switch (c->h_lvl_wpr[lv]) {
        case 1: k_sweep_fwd<MP, 1><<<n, SWEEP_THREADS, 0, st>>>(*c, i0); break;
        case 2: k_sweep_fwd<MP, 2><<<n, SWEEP_THREADS, 0, st>>>(*c, i0); break;
        case 4: k_sweep_fwd<MP, 4><<<n, SWEEP_THREADS, 0, st>>>(*c, i0); break;
        default: k_sweep_fwd<MP, 8><<<n, SWEEP_THREADS, 0, st>>>(*c, i0); break;
        }
        DOTS_LAUNCH_CHECK();
    }
    for (int lv = c->n_levels - 1; lv >= 0; --lv) {
        const int i0 = c->h_lvb_ptr[lv], n = c->h_lvb_ptr[lv + 1] - i0;
        if (n <= 0) continue;
        switch (c->h_lvb_cw[lv]) {
        case 1: k_sweep_bwd<MP, 1><<<n, SWEEP_THREADS, 0, st>>>(*c, i0); break;
        case 2: k_sweep_bwd<MP, 2><<<n, SWEEP_THREADS, 0, st>>>(*c, i0); break;
        case 4: k_sweep_bwd<MP, 4><<<n, SWEEP_THREADS, 0, st>>>(*c, i0); break;
        default:
            if (MP * 8 * SWEEP_WARPS * 32 * 8 <= 48 * 1024) { k_sweep_bwd<MP, (MP <= 3 ? 8 : 4)><<<n, SWEEP_THREADS, 0, st>>>(*c, i0); }
            else { dots_set_error("column block 8 unsupported for m_pad=%d", 32 * MP); return DOTS_ERR_BAD_ARG; }
            break;
        }
        DOTS_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int dots_mode_solves(const dots_ctx_t *c, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    switch (c->m_pad / 32) {
    case 1: return launch_sweeps<1>(c, st);
    case 2: return launch_sweeps<2>(c, st);
    case 3: return launch_sweeps<3>(c, st);
    case 4: return launch_sweeps<4>(c, st);
    }
    dots_set_error("m_pad=%d unsupported", c->m_pad);
    return DOTS_ERR_BAD_ARG;
}

extern "C" int dots_time_transform(const dots_ctx_t *c, int inverse, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int nt1 = c->n_time + 1, M = c->m_pad;
    const int grid = ceil_div(c->n_vert, TT_TILE);
    if (!inverse) {
        const size_t smem = ((size_t)nt1 * M + (size_t)nt1 * TT_TILE) * sizeof(double);
        static size_t configured = 0;
        if (smem > 48 * 1024 && smem > configured) {
            DOTS_CUDA(cudaFuncSetAttribute(k_time_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        k_time_fwd<<<grid, 256, smem, st>>>(*c);
    } else {
        const size_t smem = ((size_t)nt1 * M + (size_t)TT_TILE * (M + 1)) * sizeof(double);
        static size_t configured = 0;
        if (smem > 48 * 1024 && smem > configured) {
            DOTS_CUDA(cudaFuncSetAttribute(k_time_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        k_time_bwd<<<grid, 256, smem, st>>>(*c);
    }
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
        if ((e = dots_step_tri(c, write_z && i == n_iter - 1, stream))) return e;
    }
    return 0;
}
