// Ring-streamed batched multifrontal sweeps (sweep_mode 4; row a8, reference utils/laplacian_inverse_socp.py:58-59:
// the nT+1 per-mode solves  x_a = (L + (lambda_a - eps) M)^-1 b_a ).
//
// Same factor data as k_sweep_run (lap_kernels.cu): per separator-tree node a solve-ready panel
// P = [inv(L11) ; L21 inv(L11)], `[row][col][mode]` (forward) and its column-major copy (backward), mode fastest.
// What changes is how the panels reach the SMs and how the children's updates are passed on:
//
//  * every WARP owns a small shared-memory ring (ring_stages x 4 KB) that its lane 0 feeds with 1-D bulk async copies
//    (cp.async.bulk + mbarrier, SASS UBLKCP / SYNCS).  A warp's work is one contiguous run of panel bytes, so the copies
//    are full 4 KB pieces regardless of where the (short) panel rows begin and end, the bytes in flight per SM
//    (2 blocks x 8 warps x (stages-1) x 4 KB) no longer depend on registers, and warps never wait for each other:
//    there is no block-wide barrier in the contiguous kernel (k_sweep_run at the headline size: 31-35 % warps active,
//    60-80 % long_scoreboard, 0.61-0.67 of the HBM peak);
//  * the vector operand (r_S forward, [y_S ; x_B] backward) of a stage is fetched into registers before the warp waits
//    for the stage's mbarrier, so its (L1 / L2) latency overlaps the wait; WHICH row of Z an entry multiplies, and where an
//    output ends, comes from a per-entry code array built once on the host (dots_ring_entry_rows), so the kernels carry no
//    row / column cursor arithmetic and no index indirection (ncu on the first version: 53 warp instructions per 512-byte
//    entry, issue slots 50-57 % busy - the kernel was instruction bound, not memory bound); lanes hold 2 consecutive
//    modes and move them with 16-byte accesses;
//  * "pull" extend-add: a node stores only ITS OWN contribution  -(L21 inv(L11)) r_S  to its boundary rows (`upd`, producer
//    order).  Before a level is swept, k_ring_gather adds to every vertex of that level the contributions of ALL its
//    descendants (fixed order: gidx lists them in post-order), in place in `hat`.  The pass-through gather at the end of
//    every boundary row of k_sweep_run (two dependent loads per 4-20 streamed entries on the lower levels) is gone;
//  * outputs whose run is long (top of the tree) are shared by WPR warps as contiguous pieces and combined in a fixed
//    order through shared memory (k_ring_split);
//  * optionally (ring_pdl) the level launches are chained with programmatic dependent launch: a block streams its first
//    panel stages (read-only data) before griddepcontrol.wait, so the tail of one level overlaps the ramp-up of the next.
//
// Needs m_pad % 32 == 0 and Z = [hat | ywork] in one allocation (bidx addresses both).
#include "common.cuh"

#define SR_WARPS 8
#define SR_THREADS (32 * SR_WARPS)
#define SR_STAGE_BYTES 4096

__device__ __forceinline__ size_t sr_row_off(int row, int s)             // row-major panel: entries before row `row`
{
    return (row < s) ? (size_t)row * (row + 1) / 2 : (size_t)s * (s + 1) / 2 + (size_t)(row - s) * s;
}
__device__ __forceinline__ size_t sr_col_off(int col, int s, int b)      // column-major copy: entries before column `col`
{
    return (size_t)col * (s + b) - (size_t)col * (col - 1) / 2;
}
// range [lo, hi) of the shared index (panel column, forward / panel row, backward) an output runs over
template <int DIR> __device__ __forceinline__ int sr_lo(int o) { return DIR == 0 ? 0 : o; }
template <int DIR> __device__ __forceinline__ int sr_hi(int o, int s, int b) { return DIR == 0 ? min(o + 1, s) : s + b; }

// Lane layout of one panel entry (ML doubles): VW consecutive modes per lane and access, NA accesses per entry.
//   ML = 64 : 1 x 16-byte access (lane l: modes 2l, 2l+1)          ML = 128: 2 x 16-byte accesses
//   ML = 32 : 1 x  8-byte access                                    ML = 96 : 3 x  8-byte accesses
template <int ML> struct sr_lane {
    static constexpr int VW = (ML % 64 == 0) ? 2 : 1;
    static constexpr int NA = ML / (32 * VW);
};
template <int VW> struct sr_vec;
template <> struct sr_vec<1> {
    double x;
    __device__ __forceinline__ void zero() { x = 0.0; }
    __device__ __forceinline__ void fma(const sr_vec<1> &p, const sr_vec<1> &r) { x += p.x * r.x; }
    __device__ __forceinline__ void add(const sr_vec<1> &p) { x += p.x; }
    __device__ __forceinline__ sr_vec<1> scaled(double f) const { sr_vec<1> o; o.x = f * x; return o; }
    static __device__ __forceinline__ sr_vec<1> ldcg(const double *p) { sr_vec<1> o; o.x = __ldcg(p); return o; }
};
template <> struct __align__(16) sr_vec<2> {
    double x, y;
    __device__ __forceinline__ void zero() { x = y = 0.0; }
    __device__ __forceinline__ void fma(const sr_vec<2> &p, const sr_vec<2> &r) { x += p.x * r.x; y += p.y * r.y; }
    __device__ __forceinline__ void add(const sr_vec<2> &p) { x += p.x; y += p.y; }
    __device__ __forceinline__ sr_vec<2> scaled(double f) const { sr_vec<2> o; o.x = f * x; o.y = f * y; return o; }
    static __device__ __forceinline__ sr_vec<2> ldcg(const double *p)                     // L2 only: data written by other SMs in this launch
    {
        const double2 d = __ldcg(reinterpret_cast<const double2 *>(p));
        sr_vec<2> o; o.x = d.x; o.y = d.y; return o;
    }
};

template <int ML, int DIR>
__device__ __forceinline__ void sr_flush(const dots_ctx_t &c, const dots_ring_task_t &t, int o,
                                         const sr_vec<sr_lane<ML>::VW> (&acc)[sr_lane<ML>::NA], int lane)
{
    constexpr int VW = sr_lane<ML>::VW, NA = sr_lane<ML>::NA;
    double *dst;
    double sign;
    if (DIR == 1) { dst = c.hat + (size_t)(t.off + o) * ML; sign = -1.0; }                 // x_S = -P^T [y_S ; x_B]
    else if (o < t.s) { dst = c.ywork + (size_t)(t.off + o) * ML; sign = 1.0; }            // y_S = inv(L11) r_S
    else { dst = c.upd + (size_t)(t.ubase + o - t.s) * ML; sign = -1.0; }                  // own contribution to boundary row o - s
#pragma unroll
    for (int a = 0; a < NA; ++a) *reinterpret_cast<sr_vec<VW> *>(dst + a * 32 * VW + lane * VW) = acc[a].scaled(sign);
}

// One ring stage: `ne` (<= EC) consecutive panel entries in shared memory, `code` = this lane's entry code (lane e holds
// the operand row of entry e; bit 31: last entry of its output).  The operand rows are fetched BEFORE the stage's
// mbarrier is waited for, so their L1 / L2 latency overlaps the wait.  FLUSH: outputs may end inside the stage.
template <int ML, int DIR, int EC, bool FLUSH>
__device__ __forceinline__ void sr_stage(const dots_ctx_t &c, const dots_ring_task_t &t, const double *z, const double *sp,
                                         uint64_t *bar, uint32_t phase, int ne, int code, int lane, int &o,
                                         sr_vec<sr_lane<ML>::VW> (&acc)[sr_lane<ML>::NA])
{
    constexpr int VW = sr_lane<ML>::VW, NA = sr_lane<ML>::NA;
    typedef sr_vec<VW> vec_t;
    vec_t rv[EC][NA];
    int cd[EC];
#pragma unroll
    for (int e = 0; e < EC; ++e) {                           // lanes >= ne hold code 0: row 0 is fetched and never used (no predication)
        cd[e] = __shfl_sync(0xffffffffu, code, e);
        unsigned long long va;                               // z + row * ML doubles in ONE instruction (IMAD.WIDE.U32)
#ifdef SR_DEBUG_VEC0
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(va) : "r"(0u), "r"((unsigned)(ML * 8)), "l"((unsigned long long)z));   // timing experiment: WRONG results
#else
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(va) : "r"((unsigned)cd[e] & 0x7fffffffu), "r"((unsigned)(ML * 8)), "l"((unsigned long long)z));
#endif
        const double *vp = reinterpret_cast<const double *>(va);
#pragma unroll
        for (int a = 0; a < NA; ++a) rv[e][a] = *reinterpret_cast<const vec_t *>(vp + a * 32 * VW);
    }
    mbar_wait(bar, phase);
#pragma unroll
    for (int e = 0; e < EC; ++e) {
        if (e < ne) {
#pragma unroll
            for (int a = 0; a < NA; ++a) acc[a].fma(*reinterpret_cast<const vec_t *>(sp + e * ML + a * 32 * VW), rv[e][a]);
            if (FLUSH && cd[e] < 0) {
                sr_flush<ML, DIR>(c, t, o, acc, lane);
#pragma unroll
                for (int a = 0; a < NA; ++a) acc[a].zero();
                ++o;
            }
        }
    }
}

// Contiguous tasks: one warp streams the outputs [oa, oa + n_out) of a node, i.e. n_ent consecutive panel entries.
// SB = bytes per ring stage.  No cursor arithmetic: the per-entry codes (erow_fwd / erow_bwd, built once on the host) say
// which row of Z an entry multiplies and where an output ends; the codes of stage k + 1 are requested during stage k.
template <int ML, int DIR, int SB>
__global__ void __launch_bounds__(SR_THREADS, 3) k_ring_run(dots_ctx_t c, int task0, int task_end)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // no-ops unless launched as a programmatic dependent
    constexpr int VW = sr_lane<ML>::VW, NA = sr_lane<ML>::NA;
    constexpr int EC = SB / (8 * ML);                                    // panel entries per stage
    extern __shared__ __align__(128) unsigned char sr_smem[];
    const int nst = c.ring_stages;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ring = reinterpret_cast<double *>(sr_smem) + (size_t)warp * nst * EC * ML;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sr_smem + (size_t)SR_WARPS * nst * EC * ML * 8) + warp * nst;
    const int ti = task0 + blockIdx.x * SR_WARPS + warp;
    if (ti >= task_end) return;                                          // warps are independent: no block barrier below
    if (lane == 0) {
        for (int i = 0; i < nst; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const dots_ring_task_t t = (DIR == 0 ? c.rt_fwd : c.rt_bwd)[ti];
    const double *src = (DIR == 0 ? c.panels : c.panels_t) + (size_t)t.pbase * ML;
    const int32_t *codes = (DIR == 0 ? c.erow_fwd : c.erow_bwd) + t.pbase;
    const int n_ent = t.n_ent, n_stage = (n_ent + EC - 1) / EC;
    const bool evict_first = (c.ring_flags & 1) != 0;
    const uint64_t policy = l2_policy_evict_first();

    auto issue = [&](int k, int slot) {                                  // lane 0: stage k of the run -> ring slot
        const uint32_t bytes = (uint32_t)min(EC, n_ent - k * EC) * (uint32_t)(ML * 8);
        mbar_expect_tx(&bar[slot], bytes);
        if (evict_first) tma_load_1d_hint(ring + (size_t)slot * EC * ML, src + (size_t)k * EC * ML, bytes, &bar[slot], policy);
        else tma_load_1d(ring + (size_t)slot * EC * ML, src + (size_t)k * EC * ML, bytes, &bar[slot]);
    };
    if (lane == 0) {
        for (int k = 0; k < nst && k < n_stage; ++k) issue(k, k);        // panels and codes are read-only: stream before the wait
    }
    int code_next = (lane < EC && lane < n_ent) ? codes[lane] : 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");                   // below: data written by the previous launches

    const double *z = c.hat + lane * VW;                                 // Z = [hat | ywork]
    int o = t.oa;
    sr_vec<VW> acc[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) acc[a].zero();
    int slot = 0;
    uint32_t phase = 0;
    for (int k = 0; k < n_stage; ++k) {
        const int ne = min(EC, n_ent - k * EC);
        const int code = code_next;
        const int nxt = (k + 1) * EC + lane;
        code_next = (lane < EC && nxt < n_ent) ? codes[nxt] : 0;
        sr_stage<ML, DIR, EC, true>(c, t, z, ring + (size_t)slot * EC * ML + lane * VW, &bar[slot], phase, ne, code, lane, o, acc);
        __syncwarp();                                                    // every lane is done reading the slot
        if (lane == 0 && k + nst < n_stage) issue(k + nst, slot);
        if (++slot == nst) { slot = 0; phase ^= 1u; }
    }
}

// ------------------------------------------------------------------------------------------------
// Split items: block = (node, outputs [oa, oa + n_out)); ROWS = 8 / WPR outputs per pass, the run of an output is cut into
// WPR contiguous pieces (one per warp), partial sums combined in warp order through shared memory.  A warp's chunk
// sequence runs over all passes of the item; two cursors walk it: the bulk copies (ring_stages chunks ahead) and the math.
template <int ML, int WPR, int DIR, int SB>
__global__ void __launch_bounds__(SR_THREADS, 2) k_ring_split(dots_ctx_t c, int item0)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr int VW = sr_lane<ML>::VW, NA = sr_lane<ML>::NA;
    constexpr int EC = SB / (8 * ML);
    constexpr int ROWS = SR_WARPS / WPR;
    extern __shared__ __align__(128) unsigned char sr_smem[];
    const int nst = c.ring_stages;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ring = reinterpret_cast<double *>(sr_smem) + (size_t)warp * nst * EC * ML;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sr_smem + (size_t)SR_WARPS * nst * EC * ML * 8) + warp * nst;
    double *red = reinterpret_cast<double *>(sr_smem + (((size_t)SR_WARPS * nst * (EC * ML * 8 + 8) + 15) & ~(size_t)15));   // [SR_WARPS][ML]
    if (lane == 0) {
        for (int i = 0; i < nst; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const dots_ring_task_t t = (DIR == 0 ? c.rt_fwd : c.rt_bwd)[item0 + blockIdx.x];
    const double *pan = (DIR == 0 ? c.panels : c.panels_t) + (size_t)t.pbase * ML;
    const int32_t *codes = (DIR == 0 ? c.erow_fwd : c.erow_bwd) + t.pbase;
    const int s = t.s, b = t.b, last = t.oa + t.n_out - 1;
    const int npass = (t.n_out + ROWS - 1) / ROWS;
    const int rslot = warp / WPR, cslot = warp % WPR;

    struct Cur { int pass, x, xb; size_t ent; };                         // chunk [x, min(x + EC, xb)) of the piece; ent: panel entry of x
    // this warp's piece of the output it shares in pass `pass`
    auto piece = [&](Cur &q) -> bool {
        const int o = t.oa + q.pass * ROWS + rslot;
        if (o > last) return false;
        const int lo = sr_lo<DIR>(o), hi = sr_hi<DIR>(o, s, b);
        const int plen = ((hi - lo + WPR - 1) / WPR + EC - 1) / EC * EC;
        q.x = lo + cslot * plen;
        q.xb = min(hi, q.x + plen);
        q.ent = (DIR == 0 ? sr_row_off(o, s) : sr_col_off(o, s, b)) + (size_t)(q.x - lo);
        return q.x < q.xb;
    };
    auto next = [&](Cur &q) -> bool {                                    // advance to the warp's next chunk; false: no more
        q.x += EC;
        q.ent += EC;
        while (q.x >= q.xb) {
            if (++q.pass >= npass) return false;
            if (!piece(q)) q.x = q.xb = 0;
        }
        return true;
    };
    const bool evict_first = (c.ring_flags & 1) != 0;
    const uint64_t policy = l2_policy_evict_first();
    auto issue = [&](const Cur &q, int slot) {                           // lane 0
        const uint32_t bytes = (uint32_t)min(EC, q.xb - q.x) * (uint32_t)(ML * 8);
        mbar_expect_tx(&bar[slot], bytes);
        if (evict_first) tma_load_1d_hint(ring + (size_t)slot * EC * ML, pan + q.ent * ML, bytes, &bar[slot], policy);
        else tma_load_1d(ring + (size_t)slot * EC * ML, pan + q.ent * ML, bytes, &bar[slot]);
    };
    Cur pc{-1, 0, 0, 0}, cc{-1, 0, 0, 0};
    bool pvalid = next(pc), cvalid = next(cc);
    for (int i = 0; i < nst && pvalid; ++i) {
        if (lane == 0) issue(pc, i);
        pvalid = next(pc);
    }
    int code_next = (cvalid && lane < min(EC, cc.xb - cc.x)) ? codes[cc.ent + lane] : 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const double *z = c.hat + lane * VW;
    int slot = 0, o_unused = 0;
    uint32_t phase = 0;
    for (int pass = 0; pass < npass; ++pass) {
        sr_vec<VW> acc[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) acc[a].zero();
        while (cvalid && cc.pass == pass) {
            const int ne = min(EC, cc.xb - cc.x);
            const int code = code_next;
            const double *sp = ring + (size_t)slot * EC * ML + lane * VW;
            uint64_t *bp = &bar[slot];
            const uint32_t ph = phase;
            cvalid = next(cc);                                           // the codes of the next chunk are requested now
            code_next = (cvalid && lane < min(EC, cc.xb - cc.x)) ? codes[cc.ent + lane] : 0;
            sr_stage<ML, DIR, EC, false>(c, t, z, sp, bp, ph, ne, code, lane, o_unused, acc);
            __syncwarp();
            if (pvalid) {
                if (lane == 0) issue(pc, slot);
                pvalid = next(pc);
            }
            if (++slot == nst) { slot = 0; phase ^= 1u; }
        }
#pragma unroll
        for (int a = 0; a < NA; ++a) *reinterpret_cast<sr_vec<VW> *>(red + warp * ML + a * 32 * VW + lane * VW) = acc[a];
        __syncthreads();
        const int o = t.oa + pass * ROWS + rslot;
        if (cslot == 0 && o <= last) {
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                sr_vec<VW> v;
                v.zero();
#pragma unroll
                for (int w = 0; w < WPR; ++w) v.add(*reinterpret_cast<const sr_vec<VW> *>(red + (rslot * WPR + w) * ML + a * 32 * VW + lane * VW));
                acc[a] = v;
            }
            sr_flush<ML, DIR>(c, t, o, acc, lane);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// r_S of one tree level, in place in `hat`:  hat[v] += sum of the descendants' contributions landing on v (fixed order).
// One warp per vertex: the warp reads up to 32 list entries with ONE coalesced load and hands them round with shuffles, so
// the dependent chain is (list -> rows) once per 32 contributions instead of once per 4, and the rows of a batch of 8 are in
// flight together; lanes hold VW consecutive modes (16-byte accesses where ML allows).
template <int ML>
__global__ void __launch_bounds__(256) k_ring_gather(dots_ctx_t c, int v0, int v_end)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr int VW = sr_lane<ML>::VW, NA = sr_lane<ML>::NA;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = v0 + blockIdx.x * SR_WARPS + warp;
    const bool on = i < v_end;
    int v = 0, g0 = 0, g1 = 0, idx = 0;
    if (on) {                                                            // structure: constant, safe before the wait
        v = c.gverts[i];
        g0 = c.gptr[v];
        g1 = c.gptr[v + 1];
        if (g0 + lane < g1) idx = c.gidx[g0 + lane];
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!on) return;
    double *h = c.hat + (size_t)v * ML + lane * VW;
    sr_vec<VW> r[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) r[a] = *reinterpret_cast<const sr_vec<VW> *>(h + a * 32 * VW);
    for (int g = g0; g < g1; g += 32) {
        if (g > g0) idx = (g + lane < g1) ? c.gidx[g + lane] : 0;
        const int n = min(32, g1 - g);
        for (int j = 0; j < n; j += 8) {
            sr_vec<VW> t[8][NA];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int row = __shfl_sync(0xffffffffu, idx, (j + u) & 31);
                if (j + u < n) {
                    const double *p = c.upd + (size_t)row * ML + lane * VW;
#pragma unroll
                    for (int a = 0; a < NA; ++a) t[u][a] = *reinterpret_cast<const sr_vec<VW> *>(p + a * 32 * VW);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (j + u < n) {
#pragma unroll
                    for (int a = 0; a < NA; ++a) r[a].add(t[u][a]);
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) *reinterpret_cast<sr_vec<VW> *>(h + a * 32 * VW) = r[a];
}

// ------------------------------------------------------------------------------------------------
template <typename... Args>
static int sr_launch(void (*kern)(Args...), int grid, int threads, size_t smem, cudaStream_t st, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    DOTS_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    return 0;
}

static size_t sr_smem_bytes(int ML, int SB, int nst, bool split)
{
    const int EC = SB / (8 * ML);
    return (((size_t)SR_WARPS * nst * ((size_t)EC * ML * 8 + 8) + 15) & ~(size_t)15) + (split ? (size_t)SR_WARPS * ML * 8 : 0) + 16;
}

template <int ML, int SB>
static int sr_configure(int nst)
{
    // the opt-in is per device and per function: keyed by (device, stages) so that a second engine on another GPU of the
    // same process configures its own copy
    static int done[64] = {0};
    int dev = 0;
    DOTS_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && done[dev] == nst) return 0;
    const int run = (int)sr_smem_bytes(ML, SB, nst, false), split = (int)sr_smem_bytes(ML, SB, nst, true);
    DOTS_CUDA(cudaFuncSetAttribute(k_ring_run<ML, 0, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, run));
    DOTS_CUDA(cudaFuncSetAttribute(k_ring_run<ML, 1, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, run));
    DOTS_CUDA(cudaFuncSetAttribute(k_ring_split<ML, 2, 0, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, split));
    DOTS_CUDA(cudaFuncSetAttribute(k_ring_split<ML, 4, 0, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, split));
    DOTS_CUDA(cudaFuncSetAttribute(k_ring_split<ML, 8, 0, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, split));
    DOTS_CUDA(cudaFuncSetAttribute(k_ring_split<ML, 2, 1, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, split));
    DOTS_CUDA(cudaFuncSetAttribute(k_ring_split<ML, 4, 1, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, split));
    DOTS_CUDA(cudaFuncSetAttribute(k_ring_split<ML, 8, 1, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, split));
    if (dev >= 0 && dev < 64) done[dev] = nst;
    return 0;
}

template <int ML, int DIR, int SB>
static int sr_level(const dots_ctx_t *c, int lv, bool chain, cudaStream_t st)
{
    const int32_t *ptr = DIR == 0 ? c->h_rt_fwd_ptr : c->h_rt_bwd_ptr;
    const int wpr = (DIR == 0 ? c->h_rt_fwd_wpr : c->h_rt_bwd_wpr)[lv];
    const int i0 = ptr[lv], n = ptr[lv + 1] - i0;
    if (n <= 0) return 0;
    const int nst = c->ring_stages;
    if (wpr == 1) return sr_launch(k_ring_run<ML, DIR, SB>, ceil_div(n, SR_WARPS), SR_THREADS, sr_smem_bytes(ML, SB, nst, false), st, chain, *c, i0, i0 + n);
    const size_t smem = sr_smem_bytes(ML, SB, nst, true);
    switch (wpr) {
    case 2: return sr_launch(k_ring_split<ML, 2, DIR, SB>, n, SR_THREADS, smem, st, chain, *c, i0);
    case 4: return sr_launch(k_ring_split<ML, 4, DIR, SB>, n, SR_THREADS, smem, st, chain, *c, i0);
    case 8: return sr_launch(k_ring_split<ML, 8, DIR, SB>, n, SR_THREADS, smem, st, chain, *c, i0);
    }
    dots_set_error("ring sweep: wpr=%d unsupported", wpr);
    return DOTS_ERR_BAD_ARG;
}

// `marks` (optional, profiling): an event is recorded before every launch and after the last one; tags[i] = level of launch
// i, +1000 for a gather, +2000 for a backward level.
struct sr_marks {
    cudaEvent_t *ev;
    int *tag;
    int cap, n;
};
static int sr_mark(sr_marks *mk, int tag, cudaStream_t st)
{
    if (!mk || mk->n >= mk->cap) return 0;
    DOTS_CUDA(cudaEventRecord(mk->ev[mk->n], st));
    mk->tag[mk->n++] = tag;
    return 0;
}

template <int ML, int SB>
static int sr_sweeps(const dots_ctx_t *c, cudaStream_t st, sr_marks *mk)
{
    if (int e = sr_configure<ML, SB>(c->ring_stages)) return e;
    const bool pdl = c->ring_pdl != 0;
    bool chain = true;                                                    // the first launch chains to the time transform
    for (int lv = 0; lv < c->n_levels; ++lv) {
        const int g0 = c->h_gv_ptr[lv], gn = c->h_gv_ptr[lv + 1] - g0;
        if (gn > 0) {
            if (int e = sr_mark(mk, 1000 + lv, st)) return e;
            if (int e = sr_launch(k_ring_gather<ML>, ceil_div(gn, SR_WARPS), 256, 0, st, pdl && chain, *c, g0, g0 + gn)) return e;
            chain = true;
        }
        const int n = c->h_rt_fwd_ptr[lv + 1] - c->h_rt_fwd_ptr[lv];
        if (n > 0) { if (int e = sr_mark(mk, lv, st)) return e; }
        if (int e = sr_level<ML, 0, SB>(c, lv, pdl && chain, st)) return e;
        if (n > 0) chain = true;
    }
    for (int lv = c->n_levels - 1; lv >= 0; --lv) {
        const int n = c->h_rt_bwd_ptr[lv + 1] - c->h_rt_bwd_ptr[lv];
        if (n > 0) { if (int e = sr_mark(mk, 2000 + lv, st)) return e; }
        if (int e = sr_level<ML, 1, SB>(c, lv, pdl && chain, st)) return e;
        if (n > 0) chain = true;
    }
    return sr_mark(mk, -1, st);
}

static int sr_dispatch(const dots_ctx_t *c, void *stream, sr_marks *mk)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (c->m_pad % 32 || c->m_pad > 128) { dots_set_error("ring sweeps need m_pad in {32, 64, 96, 128} (got %d)", c->m_pad); return DOTS_ERR_BAD_ARG; }
    if (c->ywork != c->hat + (size_t)c->n_vert * c->m_pad) { dots_set_error("ring sweeps need ywork == hat + n_vert * m_pad"); return DOTS_ERR_BAD_ARG; }
    if (c->ring_stages < 2 || c->ring_stages > 6) { dots_set_error("ring_stages=%d outside 2..6", c->ring_stages); return DOTS_ERR_BAD_ARG; }
    if (!c->rt_fwd || !c->rt_bwd || !c->erow_fwd || !c->erow_bwd || !c->gptr || !c->gidx) { dots_set_error("ring sweep plan missing from the context"); return DOTS_ERR_BAD_ARG; }
    const bool small = c->ring_stage_bytes == 2048;
    if (!small && c->ring_stage_bytes != 4096) { dots_set_error("ring_stage_bytes=%d: 2048 or 4096", c->ring_stage_bytes); return DOTS_ERR_BAD_ARG; }
    switch (c->m_pad) {
    case 32: return small ? sr_sweeps<32, 2048>(c, st, mk) : sr_sweeps<32, 4096>(c, st, mk);
    case 64: return small ? sr_sweeps<64, 2048>(c, st, mk) : sr_sweeps<64, 4096>(c, st, mk);
    case 96: return small ? sr_sweeps<96, 2048>(c, st, mk) : sr_sweeps<96, 4096>(c, st, mk);
    case 128: return small ? sr_sweeps<128, 2048>(c, st, mk) : sr_sweeps<128, 4096>(c, st, mk);
    }
    return DOTS_ERR_BAD_ARG;
}

int dots_mode_solves_ring(const dots_ctx_t *c, void *stream) { return sr_dispatch(c, stream, nullptr); }

// Profiling aid (tools/level_times.py): one pair of ring sweeps with a CUDA event between the launches.  ms_out[i] = time from
// the start of launch i to the start of launch i + 1 (the end of the sweeps for the last one), tag_out[i] as in sr_marks.
// Synchronises the stream.  Returns the number of launches in *n_out.
extern "C" int dots_ring_level_times(const dots_ctx_t *c, void *stream, float *ms_out, int32_t *tag_out, int cap, int *n_out)
{
    if (int e = dots_check_ctx(c)) return e;
    if (c->sweep_mode != 4 || cap < 2 || cap > 256) { dots_set_error("dots_ring_level_times: sweep_mode 4 and 2 <= cap <= 256"); return DOTS_ERR_BAD_ARG; }
    cudaEvent_t ev[256];
    int tag[256];
    for (int i = 0; i < cap; ++i) DOTS_CUDA(cudaEventCreate(&ev[i]));
    sr_marks mk{ev, tag, cap, 0};
    int e = sr_dispatch(c, stream, &mk);
    cudaError_t ce = cudaStreamSynchronize((cudaStream_t)stream);
    int n = 0;
    if (!e && ce == cudaSuccess) {
        for (int i = 0; i + 1 < mk.n; ++i) {
            if (cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]) != cudaSuccess) break;
            tag_out[i] = tag[i];
            n = i + 1;
        }
    }
    for (int i = 0; i < cap; ++i) cudaEventDestroy(ev[i]);
    if (n_out) *n_out = n;
    if (!e && ce != cudaSuccess) { dots_set_error("dots_ring_level_times: %s", cudaGetErrorString(ce)); return (int)ce; }
    return e;
}
