// Persistent, TMA-fed form of the batched multifrontal sweeps (row a8; reference utils/laplacian_inverse_socp.py:58-59).
//
// One cooperative launch runs BOTH sweeps over ALL separator-tree levels:
//     phases 0..L-1      forward  level 0..L-1   [y_S ; -du_B] = P r_S                (leaves -> root)
//     phases L..2L-1     backward level L-1..0   xt_S = Pt^T-form of the same panel   (root -> leaves)
// with a grid-wide barrier between phases (parent/child dependencies are the only ordering constraints).  Compared
// with one launch per level and direction this removes ~40 launch boundaries per iteration (what the small
// knots_5-class meshes are bound by) and, more importantly, decouples the bytes in flight from occupancy: thread 0
// of every block streams the block's panel segments with 1-D bulk async copies (cp.async.bulk -> UBLKCP) into a
// 4-slot shared-memory ring guarded by mbarriers, running SW_STAGES segments ahead of the math - also ACROSS the
// grid barriers, because the factor panels are read-only.  (ncu on the per-level kernels: >80 % long_scoreboard.)
//
// Work decomposition.  An item is (node, first output, n outputs <= 8).  Forward outputs are panel rows (row i
// owns columns [0, min(i+1,s)) of the row-major panel), backward outputs are panel columns (column j owns rows
// [j, s+b) of the column-major copy `panels_t`), so in both directions an output's data is one contiguous run and
// is cut into segments of <= CH entries aligned to multiples of CH in the INPUT index; the input-vector chunk of a
// segment (r_S with the children's updates folded in, or -[y_S ; xt_B]) is staged once per (item, chunk) in shared
// memory and shared by the item's outputs.  Inside a segment the 8 warps interleave over the entries (lane = time
// mode), and their partial sums are combined once per item through shared memory in a fixed order.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define SW_THREADS 256
#define SW_WARPS 8
#define SW_STAGES 4
#define SW_MAXOUT 8

__host__ __device__ __forceinline__ int sw_chunk(int M) { return M <= 32 ? 64 : (M <= 64 ? 32 : 16); }

__device__ __forceinline__ size_t row_off_f(int row, int s)            // row-major panel: offset of row `row`
{
    return (row < s) ? (size_t)row * (row + 1) / 2 : (size_t)s * (s + 1) / 2 + (size_t)(row - s) * s;
}
__device__ __forceinline__ size_t col_off_t(int col, int s, int b)     // column-major copy: offset of column `col`
{
    return (size_t)col * (s + b) - (size_t)col * (col - 1) / 2;
}

struct SwItem {
    int node, o0, n_o, s, b, off;
    int cb_lo, cb_end;                 // chunk range of the input index
    size_t panel;                      // panel offset of the node (entries)
};

__device__ __forceinline__ void sw_load_item(const dots_ctx_t &c, int dir, const int *items, int it, int CH, SwItem &I)
{
    const int *p = items + 3 * (size_t)it;
    I.node = p[0]; I.o0 = p[1]; I.n_o = p[2];
    I.s = c.nd_s[I.node]; I.b = c.nd_b[I.node]; I.off = c.nd_off[I.node];
    I.panel = (size_t)c.nd_panel[I.node];
    if (dir == 0) { I.cb_lo = 0; I.cb_end = min(I.o0 + I.n_o, I.s); }
    else { I.cb_lo = (I.o0 / CH) * CH; I.cb_end = I.s + I.b; }
}

// entries of output `o` that fall in the input chunk [cb, cb+CH): returns count, sets first entry index
__device__ __forceinline__ int sw_seg(int dir, int o, int s, int b, int cb, int CH, int &e_lo)
{
    const int lo = dir == 0 ? 0 : o, hi = dir == 0 ? min(o + 1, s) : s + b;
    e_lo = max(lo, cb);
    return min(hi, cb + CH) - e_lo;
}

// flattened walk over (phase, item, chunk, output) used by thread 0 to run the bulk copies ahead of the math
struct SwProducer {
    int phase, it, cb, q, n_phases, L, CH;
    SwItem I;
    bool item_ok;
};

__device__ __forceinline__ bool sw_prod_item(const dots_ctx_t &c, SwProducer &P, int first_it)
{
    // position on item `first_it` of the current phase or on the first item of a later phase
    int it = first_it;
    while (P.phase < P.n_phases) {
        const int dir = P.phase < P.L ? 0 : 1;
        const int lv = dir == 0 ? P.phase : 2 * P.L - 1 - P.phase;
        const int *ptr = dir == 0 ? c.lvl_ptr : c.lvb_ptr;
        if (it < 0) it = ptr[lv] + blockIdx.x;
        if (it < ptr[lv + 1]) {
            sw_load_item(c, dir, dir == 0 ? c.lvl_items : c.lvb_items, it, P.CH, P.I);
            P.it = it; P.cb = P.I.cb_lo; P.q = 0;
            return true;
        }
        ++P.phase;
        it = -1;
    }
    return false;
}

// next non-empty segment: global source address + byte count; false when the walk is exhausted
__device__ __forceinline__ bool sw_prod_next(const dots_ctx_t &c, SwProducer &P, const double *&src, uint32_t &bytes, int M)
{
    while (P.item_ok) {
        const int dir = P.phase < P.L ? 0 : 1;
        while (P.cb < P.I.cb_end) {
            while (P.q < P.I.n_o) {
                const int o = P.I.o0 + P.q;
                int e_lo;
                const int ne = sw_seg(dir, o, P.I.s, P.I.b, P.cb, P.CH, e_lo);
                ++P.q;
                if (ne > 0) {
                    const size_t ent = dir == 0 ? P.I.panel + row_off_f(o, P.I.s) + e_lo
                                                : P.I.panel + col_off_t(o, P.I.s, P.I.b) + (e_lo - o);
                    src = (dir == 0 ? c.panels : c.panels_t) + ent * M;
                    bytes = (uint32_t)ne * (uint32_t)M * 8u;
                    return true;
                }
            }
            P.q = 0;
            P.cb += P.CH;
        }
        P.item_ok = sw_prod_item(c, P, P.it + (int)gridDim.x);
    }
    return false;
}

template <int MP>
__global__ void __launch_bounds__(SW_THREADS, 2) k_sweeps(dots_ctx_t c)
{
    constexpr int M = 32 * MP;
    extern __shared__ __align__(128) unsigned char smraw[];
    const int CH = sw_chunk(M);
    double *ring = reinterpret_cast<double *>(smraw);                     // [SW_STAGES][CH][M]
    double *vec = ring + (size_t)SW_STAGES * CH * M;                      // [CH][M]
    double *red = vec + (size_t)CH * M;                                   // [SW_WARPS][SW_MAXOUT][M]
    uint64_t *full = reinterpret_cast<uint64_t *>(red + (size_t)SW_WARPS * SW_MAXOUT * M);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = c.n_levels, n_phases = 2 * L;
    cg::grid_group grid = cg::this_grid();

    if (tid == 0) {
        for (int i = 0; i < SW_STAGES; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    SwProducer P;
    P.phase = 0; P.n_phases = n_phases; P.L = L; P.CH = CH; P.it = 0; P.cb = 0; P.q = 0;
    uint32_t n_issued = 0, n_used = 0;
    if (tid == 0) {
        P.item_ok = sw_prod_item(c, P, -1);
        for (int i = 0; i < SW_STAGES; ++i) {
            const double *src; uint32_t bytes;
            if (!sw_prod_next(c, P, src, bytes, M)) break;
            mbar_expect_tx(&full[i], bytes);
            tma_load_1d(ring + (size_t)i * CH * M, src, bytes, &full[i]);
            ++n_issued;
        }
    }

    auto stamp = [&](int slot) {
        if (c.phase_clock && blockIdx.x == 0 && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            c.phase_clock[slot] = t;
        }
    };
    stamp(0);
    for (int phase = 0; phase < n_phases; ++phase) {
        const int dir = phase < L ? 0 : 1;
        const int lv = dir == 0 ? phase : 2 * L - 1 - phase;
        const int *ptr = dir == 0 ? c.lvl_ptr : c.lvb_ptr;
        const int *items = dir == 0 ? c.lvl_items : c.lvb_items;
        for (int it = ptr[lv] + blockIdx.x; it < ptr[lv + 1]; it += gridDim.x) {
            SwItem I;
            sw_load_item(c, dir, items, it, CH, I);
            const int ch0 = c.nd_child[2 * I.node], ch1 = c.nd_child[2 * I.node + 1];
            const double *u0 = (ch0 >= 0) ? c.upd + (size_t)c.nd_upd[ch0] * M : nullptr;
            const double *u1 = (ch1 >= 0) ? c.upd + (size_t)c.nd_upd[ch1] * M : nullptr;
            const int32_t *cp0 = c.child_pos + c.nd_front[I.node];
            const int32_t *cp1 = cp0 + c.front_total;
            const int32_t *fidx = c.front_idx + c.nd_front[I.node];

            double acc[SW_MAXOUT][MP];
#pragma unroll
            for (int q = 0; q < SW_MAXOUT; ++q)
#pragma unroll
                for (int m = 0; m < MP; ++m) acc[q][m] = 0.0;

            for (int cb = I.cb_lo; cb < I.cb_end; cb += CH) {
                // stage the input-vector chunk [cb, cb+CH) (previous readers are past the post-segment barrier)
                const int nvec = min(CH, I.cb_end - cb);
                for (int i = tid; i < nvec * M; i += SW_THREADS) {
                    const int jj = i / M, m = i - jj * M, j = cb + jj;
                    double v;
                    if (dir == 0) {
                        v = c.hat[(size_t)(I.off + j) * M + m];
                        const int a = cp0[j], bb = cp1[j];
                        if (u0 && a >= 0) v += u0[(size_t)a * M + m];
                        if (u1 && bb >= 0) v += u1[(size_t)bb * M + m];
                    } else {
                        v = -((j < I.s) ? c.ywork[(size_t)(I.off + j) * M + m] : c.hat[(size_t)fidx[j] * M + m]);
                    }
                    vec[jj * M + m] = v;
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < SW_MAXOUT; ++q) {
                    if (q >= I.n_o) break;
                    int e_lo;
                    const int ne = sw_seg(dir, I.o0 + q, I.s, I.b, cb, CH, e_lo);
                    if (ne <= 0) continue;
                    const int slot = n_used % SW_STAGES;
                    mbar_wait(&full[slot], (n_used / SW_STAGES) & 1);
                    const double *seg = ring + (size_t)slot * CH * M + lane;
                    const double *vv = vec + (size_t)(e_lo - cb) * M + lane;
                    for (int e = warp; e < ne; e += SW_WARPS) {
#pragma unroll
                        for (int m = 0; m < MP; ++m) acc[q][m] += seg[(size_t)e * M + 32 * m] * vv[(size_t)e * M + 32 * m];
                    }
                    ++n_used;
                    __syncthreads();                                        // slot (and later: vec) free again
                    if (tid == 0) {
                        const double *src; uint32_t bytes;
                        if (sw_prod_next(c, P, src, bytes, M)) {
                            mbar_expect_tx(&full[slot], bytes);
                            tma_load_1d(ring + (size_t)slot * CH * M, src, bytes, &full[slot]);
                            ++n_issued;
                        }
                    }
                }
            }
            // combine the 8 warps' partial sums in a fixed order and write the item's outputs
#pragma unroll
            for (int q = 0; q < SW_MAXOUT; ++q)
#pragma unroll
                for (int m = 0; m < MP; ++m) red[((size_t)warp * SW_MAXOUT + q) * M + 32 * m + lane] = acc[q][m];
            __syncthreads();
            double *myupd = c.upd + (size_t)c.nd_upd[I.node] * M;
            for (int i = tid; i < I.n_o * M; i += SW_THREADS) {
                const int q = i / M, m = i - q * M, o = I.o0 + q;
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < SW_WARPS; ++w) v += red[((size_t)w * SW_MAXOUT + q) * M + m];
                if (dir == 0) {
                    if (o < I.s) {
                        c.ywork[(size_t)(I.off + o) * M + m] = v;
                    } else {
                        const int a = cp0[o], bb = cp1[o];
                        double val = 0.0;
                        if (u0 && a >= 0) val += u0[(size_t)a * M + m];
                        if (u1 && bb >= 0) val += u1[(size_t)bb * M + m];
                        myupd[(size_t)(o - I.s) * M + m] = val - v;
                    }
                } else {
                    c.hat[(size_t)(I.off + o) * M + m] = v;
                }
            }
            __syncthreads();
        }
        grid.sync();
        stamp(phase + 1);
    }
}

template <int MP>
static int launch_persistent(const dots_ctx_t *c, cudaStream_t st)
{
    constexpr int M = 32 * MP;
    const int CH = sw_chunk(M);
    const size_t smem = ((size_t)(SW_STAGES + 1) * CH * M + (size_t)SW_WARPS * SW_MAXOUT * M) * sizeof(double) + 64;
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        DOTS_CUDA(cudaFuncSetAttribute(k_sweeps<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DOTS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_sweeps<MP>, SW_THREADS, smem));
        if (blocks_per_sm < 1) { dots_set_error("k_sweeps does not fit on an SM (smem %zu)", smem); blocks_per_sm = 0; return DOTS_ERR_BAD_ARG; }
        if (blocks_per_sm > 2) blocks_per_sm = 2;
    }
    int grid = c->sweep_grid > 0 ? c->sweep_grid : c->n_sm * blocks_per_sm;
    if (grid > c->n_sm * blocks_per_sm) grid = c->n_sm * blocks_per_sm;
    dots_ctx_t ctx = *c;
    void *args[] = {&ctx};
    DOTS_CUDA(cudaLaunchCooperativeKernel((const void *)k_sweeps<MP>, dim3(grid), dim3(SW_THREADS), args, smem, st));
    return 0;
}

int dots_mode_solves_persistent(const dots_ctx_t *c, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (c->m_pad % 32) { dots_set_error("persistent sweep needs m_pad >= 32 (got %d)", c->m_pad); return DOTS_ERR_BAD_ARG; }
    switch (c->m_pad / 32) {
    case 1: return launch_persistent<1>(c, st);
    case 2: return launch_persistent<2>(c, st);
    case 3: return launch_persistent<3>(c, st);
    case 4: return launch_persistent<4>(c, st);
    }
    dots_set_error("m_pad=%d unsupported", c->m_pad);
    return DOTS_ERR_BAD_ARG;
}
