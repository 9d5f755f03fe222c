// EXPERIMENTAL (sweep_mode = 2, or 3 = the same with programmatic dependent launch; off by default, NOT yet run on hardware - written at the end of round 1 from the analysis in
// DESIGN.md section 8; validate with DOTS_TEST_EXPERIMENTAL=1 before use).
//
// Tile-streamed form of the batched multifrontal sweeps (row a8; reference utils/laplacian_inverse_socp.py:58-59).
// Same math and same per-level launches as k_sweep_run (lap_kernels.cu), different data movement:
//
//   * a work item is (node, first output, n outputs); a block walks it in groups of 8 outputs, ONE WARP PER OUTPUT, so an
//     output's dot product is accumulated by one warp in a fixed order (no cross-warp combine, deterministic);
//   * the input index (columns of the row-major panel in the forward sweep, rows of the column-major copy in the
//     backward sweep) is cut into chunks of C = 512 / M entries; the 8 outputs' segments of one chunk form a 2-D tile of
//     <= 32 KB that the producer streams with 8 bulk async copies (cp.async.bulk, one mbarrier per stage) into a 3-stage
//     shared-memory ring, running 3 tiles ahead of the math.  Bytes in flight per SM = 2 blocks x 3 x 32 KB, independent
//     of occupancy and registers (k_sweep_run: ~48 KB requested, ~half of it in flight, 66 % of the HBM peak);
//   * the chunk of the input vector (r_S with the children's updates folded in / -[y_S ; xt_B]) is fetched one step ahead
//     into registers and double-buffered in shared memory; it is shared by the 8 outputs of the tile;
//   * a ninth warp is the producer: after the per-tile barrier it refills the freed stage while the math warps are already
//     on the next tile, so the address arithmetic of the 8 copies is off the critical path;
//   * one __syncthreads per tile.  The persistent variant (sweep_tma.cu) paid that per <= 16 KB (row, chunk) segment and a
//     single-thread address walk on top, which is why it lost.
//
// The children's updates are folded while staging r_S, so no k_sweep_gather launches are needed.
#include "common.cuh"

#define ST_WARPS 8                          // math warps: one output each
#define ST_MATH (32 * ST_WARPS)
#define ST_THREADS (ST_MATH + 32)           // + one producer warp (its lane 0 issues the bulk copies)
#define ST_STAGES 3

__host__ __device__ __forceinline__ int st_chunk(int M) { return 512 / M; }          // 16, 8, 5, 4 for M = 32, 64, 96, 128

__device__ __forceinline__ size_t st_row_off(int row, int s)            // row-major panel: offset of row `row` (entries)
{
    return (row < s) ? (size_t)row * (row + 1) / 2 : (size_t)s * (s + 1) / 2 + (size_t)(row - s) * s;
}
__device__ __forceinline__ size_t st_col_off(int col, int s, int b)     // column-major copy: offset of column `col`
{
    return (size_t)col * (s + b) - (size_t)col * (col - 1) / 2;
}

// input-index range [lo, hi) an output touches
template <int DIR>
__device__ __forceinline__ void st_span(int o, int s, int b, int &lo, int &hi)
{
    if (DIR == 0) { lo = 0; hi = min(o + 1, s); }
    else { lo = o; hi = s + b; }
}

// walk over the (group, chunk) steps of an item, shared by the math and by the producer (which runs ST_STAGES ahead)
struct StCursor {
    int g;          // output group (8 outputs)
    int cb;         // first input index of the chunk
    int cb_end;     // end of the chunk range of group g
};

template <int DIR>
__device__ __forceinline__ bool st_group_range(int o0, int n_o, int g, int s, int b, int C, int &cb_lo, int &cb_end)
{
    const int first = o0 + ST_WARPS * g;
    if (first >= o0 + n_o) return false;
    const int lastq = min(first + ST_WARPS, o0 + n_o) - 1;
    if (DIR == 0) { cb_lo = 0; cb_end = min(lastq + 1, s); }            // longest row of the group
    else { cb_lo = (first / C) * C; cb_end = s + b; }                    // longest column of the group
    return true;
}

template <int DIR>
__device__ __forceinline__ bool st_first(StCursor &k, int o0, int n_o, int s, int b, int C)
{
    k.g = 0;
    int lo;
    if (!st_group_range<DIR>(o0, n_o, 0, s, b, C, lo, k.cb_end)) return false;
    k.cb = lo;
    return true;
}

template <int DIR>
__device__ __forceinline__ bool st_next(StCursor &k, int o0, int n_o, int s, int b, int C)
{
    k.cb += C;
    if (k.cb < k.cb_end) return true;
    ++k.g;
    int lo;
    if (!st_group_range<DIR>(o0, n_o, k.g, s, b, C, lo, k.cb_end)) return false;
    k.cb = lo;
    return true;
}

// PDL (sweep_mode = 3): the launches of consecutive tree levels are chained with programmatic dependent launch.  A block
// lets the next level start as soon as it runs (griddepcontrol.launch_dependents) and, because the factor panels are
// read-only, streams its first ST_STAGES tiles BEFORE it waits for the previous level (griddepcontrol.wait): the tail of
// one level overlaps the first loads of the next instead of an idle launch boundary.
template <int MP, int DIR, bool PDL>
__global__ void __launch_bounds__(ST_THREADS, 2) k_sweep_tile(dots_ctx_t c, int item0)
{
    if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr int M = 32 * MP;
    constexpr int C = 512 / M;
    constexpr int VPT = (C * M + ST_MATH - 1) / ST_MATH;                  // input-vector elements per math thread and chunk
    extern __shared__ __align__(128) unsigned char smraw[];
    double *ring = reinterpret_cast<double *>(smraw);                     // [ST_STAGES][8][C][M]
    double *vec = ring + (size_t)ST_STAGES * ST_WARPS * C * M;            // [2][C][M]
    uint64_t *full = reinterpret_cast<uint64_t *>(vec + (size_t)2 * C * M);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool math = warp < ST_WARPS, producer = tid == ST_MATH;

    const int *it = (DIR == 0 ? c.lvl_items : c.lvb_items) + 3 * (size_t)(item0 + blockIdx.x);
    const int node = it[0], o0 = it[1], n_o = it[2];
    const int s = c.nd_s[node], b = c.nd_b[node], off = c.nd_off[node];
    const int ch0 = c.nd_child[2 * node], ch1 = c.nd_child[2 * node + 1];
    const double *u0 = (DIR == 0 && ch0 >= 0) ? c.upd + (size_t)c.nd_upd[ch0] * M : nullptr;
    const double *u1 = (DIR == 0 && ch1 >= 0) ? c.upd + (size_t)c.nd_upd[ch1] * M : nullptr;
    const int32_t *cp0 = c.child_pos + c.nd_front[node];
    const int32_t *cp1 = cp0 + c.front_total;
    const int32_t *fidx = c.front_idx + c.nd_front[node];
    const double *panel = (DIR == 0 ? c.panels : c.panels_t) + (size_t)c.nd_panel[node] * M;
    double *myupd = c.upd + (size_t)c.nd_upd[node] * M;

    if (tid == 0) {
        for (int i = 0; i < ST_STAGES; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // producer (lane 0 of the ninth warp): one stage = the segments of the group's outputs that fall into the chunk
    auto issue = [&](const StCursor &k, int slot) {
        uint32_t total = 0;
        const int first = o0 + ST_WARPS * k.g;
#pragma unroll
        for (int q = 0; q < ST_WARPS; ++q) {
            const int o = first + q;
            if (o >= o0 + n_o) break;
            int lo, hi;
            st_span<DIR>(o, s, b, lo, hi);
            const int e_lo = max(lo, k.cb), ne = min(hi, k.cb + C) - e_lo;
            if (ne > 0) total += (uint32_t)ne * M * 8u;
        }
        mbar_expect_tx(&full[slot], total);
#pragma unroll
        for (int q = 0; q < ST_WARPS; ++q) {
            const int o = first + q;
            if (o >= o0 + n_o) break;
            int lo, hi;
            st_span<DIR>(o, s, b, lo, hi);
            const int e_lo = max(lo, k.cb), ne = min(hi, k.cb + C) - e_lo;
            if (ne <= 0) continue;
            const size_t ent = (DIR == 0) ? st_row_off(o, s) + e_lo : st_col_off(o, s, b) + (e_lo - o);
            tma_load_1d(ring + ((size_t)slot * ST_WARPS + q) * C * M, panel + ent * M, (uint32_t)ne * M * 8u, &full[slot]);
        }
    };
    // input vector of chunk [cb, cb + C): element i of the thread's share
    auto vec_fetch = [&](int cb, int cb_end, double (&reg)[VPT]) {
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int i = tid + k * ST_MATH;
            const int jj = i / M, m = i - jj * M, j = cb + jj;
            double v = 0.0;
            if (math && i < C * M && j < cb_end) {
                if (DIR == 0) {
                    v = c.hat[(size_t)(off + j) * M + m];
                    const int a = cp0[j], bb = cp1[j];
                    if (u0 && a >= 0) v += u0[(size_t)a * M + m];
                    if (u1 && bb >= 0) v += u1[(size_t)bb * M + m];
                } else {
                    v = -((j < s) ? c.ywork[(size_t)(off + j) * M + m] : c.hat[(size_t)fidx[j] * M + m]);
                }
            }
            reg[k] = v;
        }
    };
    auto vec_store = [&](int buf, const double (&reg)[VPT]) {
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int i = tid + k * ST_MATH;
            if (math && i < C * M) vec[(size_t)buf * C * M + i] = reg[k];
        }
    };

    StCursor cur, prod;
    if (!st_first<DIR>(cur, o0, n_o, s, b, C)) return;                   // empty item (never produced by the plan)
    prod = cur;
    bool prod_ok = true;
    if (producer) {
        for (int i = 0; i < ST_STAGES && prod_ok; ++i) {
            issue(prod, i);
            prod_ok = st_next<DIR>(prod, o0, n_o, s, b, C);
        }
    }
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");           // everything below reads the previous level's output
    double reg[VPT];
    vec_fetch(cur.cb, cur.cb_end, reg);
    vec_store(0, reg);
    __syncthreads();

    double acc[MP];
#pragma unroll
    for (int m = 0; m < MP; ++m) acc[m] = 0.0;
    bool more = true;
    for (uint32_t step = 0; more; ++step) {
        const int slot = step % ST_STAGES, buf = step & 1;
        StCursor nxt = cur;
        more = st_next<DIR>(nxt, o0, n_o, s, b, C);
        if (more) vec_fetch(nxt.cb, nxt.cb_end, reg);                    // next chunk's input vector: loads in flight during the math
        const int o = o0 + ST_WARPS * cur.g + warp;
        if (math && o < o0 + n_o) {
            mbar_wait(&full[slot], (step / ST_STAGES) & 1);
            int lo, hi;
            st_span<DIR>(o, s, b, lo, hi);
            const int e_lo = max(lo, cur.cb), ne = min(hi, cur.cb + C) - e_lo;
            const double *seg = ring + ((size_t)slot * ST_WARPS + warp) * C * M + lane;
            const double *vv = vec + (size_t)buf * C * M + (size_t)(e_lo - cur.cb) * M + lane;
            for (int e = 0; e < ne; ++e) {
#pragma unroll
                for (int m = 0; m < MP; ++m) acc[m] += seg[(size_t)e * M + 32 * m] * vv[(size_t)e * M + 32 * m];
            }
            const bool group_done = !more || nxt.g != cur.g;
            if (group_done) {                                            // the output is complete: write it
                if (DIR == 1) {
#pragma unroll
                    for (int m = 0; m < MP; ++m) c.hat[(size_t)(off + o) * M + 32 * m + lane] = acc[m];
                } else if (o < s) {
#pragma unroll
                    for (int m = 0; m < MP; ++m) c.ywork[(size_t)(off + o) * M + 32 * m + lane] = acc[m];
                } else {
                    const int a = cp0[o], bb = cp1[o];
#pragma unroll
                    for (int m = 0; m < MP; ++m) {
                        double val = 0.0;
                        if (u0 && a >= 0) val += u0[(size_t)a * M + 32 * m + lane];
                        if (u1 && bb >= 0) val += u1[(size_t)bb * M + 32 * m + lane];
                        myupd[(size_t)(o - s) * M + 32 * m + lane] = val - acc[m];
                    }
                }
#pragma unroll
                for (int m = 0; m < MP; ++m) acc[m] = 0.0;
            }
        }
        if (more) vec_store(buf ^ 1, reg);
        __syncthreads();                                                 // ring[slot] and vec[buf] are free, vec[buf^1] is visible
        if (producer && prod_ok) {
            issue(prod, slot);
            prod_ok = st_next<DIR>(prod, o0, n_o, s, b, C);
        }
        cur = nxt;
    }
}

template <int MP, int DIR, bool PDL>
static int launch_tile_level(const dots_ctx_t *c, int i0, int n, size_t smem, cudaStream_t st)
{
    if (!PDL) {
        k_sweep_tile<MP, DIR, false><<<n, ST_THREADS, smem, st>>>(*c, i0);
        DOTS_LAUNCH_CHECK();
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n);
    cfg.blockDim = dim3(ST_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DOTS_CUDA(cudaLaunchKernelEx(&cfg, k_sweep_tile<MP, DIR, true>, *c, i0));
    return 0;
}

template <int MP, bool PDL>
static int launch_tile(const dots_ctx_t *c, cudaStream_t st)
{
    constexpr int M = 32 * MP;
    const int C = st_chunk(M);
    const size_t smem = ((size_t)ST_STAGES * ST_WARPS * C * M + (size_t)2 * C * M) * sizeof(double) + 64;
    static bool configured = false;
    if (!configured) {
        DOTS_CUDA(cudaFuncSetAttribute(k_sweep_tile<MP, 0, PDL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DOTS_CUDA(cudaFuncSetAttribute(k_sweep_tile<MP, 1, PDL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    for (int lv = 0; lv < c->n_levels; ++lv) {
        const int i0 = c->h_lvl_ptr[lv], n = c->h_lvl_ptr[lv + 1] - i0;
        if (n <= 0) continue;
        if (int e = launch_tile_level<MP, 0, PDL>(c, i0, n, smem, st)) return e;
    }
    for (int lv = c->n_levels - 1; lv >= 0; --lv) {
        const int i0 = c->h_lvb_ptr[lv], n = c->h_lvb_ptr[lv + 1] - i0;
        if (n <= 0) continue;
        if (int e = launch_tile_level<MP, 1, PDL>(c, i0, n, smem, st)) return e;
    }
    return 0;
}

int dots_mode_solves_tile(const dots_ctx_t *c, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (c->m_pad % 32) { dots_set_error("tile sweep needs m_pad >= 32 (got %d)", c->m_pad); return DOTS_ERR_BAD_ARG; }
    const bool pdl = c->sweep_mode == 3;
    switch (c->m_pad / 32) {
    case 1: return pdl ? launch_tile<1, true>(c, st) : launch_tile<1, false>(c, st);
    case 2: return pdl ? launch_tile<2, true>(c, st) : launch_tile<2, false>(c, st);
    case 3: return pdl ? launch_tile<3, true>(c, st) : launch_tile<3, false>(c, st);
    case 4: return pdl ? launch_tile<4, true>(c, st) : launch_tile<4, false>(c, st);
    }
    dots_set_error("m_pad=%d unsupported", c->m_pad);
    return DOTS_ERR_BAD_ARG;
}
