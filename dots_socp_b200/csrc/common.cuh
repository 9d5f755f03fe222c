// Shared helpers for the dots_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dots_b200.h"

#define DOTS_ERR_BAD_ARG (-1)
#define DOTS_ERR_ABI (-2)

void dots_set_error(const char *fmt, ...);

#define DOTS_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            dots_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return (int)e__;                                                                     \
        }                                                                                        \
    } while (0)

#define DOTS_LAUNCH_CHECK() DOTS_CUDA(cudaGetLastError())

static inline int dots_check_ctx(const dots_ctx_t *c)
{
    if (!c) { dots_set_error("null context"); return DOTS_ERR_BAD_ARG; }
    if (c->abi_version != DOTS_ABI_VERSION) { dots_set_error("abi mismatch: ctx %d lib %d", c->abi_version, DOTS_ABI_VERSION); return DOTS_ERR_ABI; }
    if (!(c->m_pad == 8 || c->m_pad == 16 || (c->m_pad % 32 == 0 && c->m_pad >= 32 && c->m_pad <= 128))) { dots_set_error("m_pad=%d unsupported (8, 16, 32, 64, 96, 128)", c->m_pad); return DOTS_ERR_BAD_ARG; }
    if (c->lvl_begin < 0 || c->lvl_end > c->n_time + 1 || c->lvl_begin >= c->lvl_end) { dots_set_error("bad level range [%d, %d)", c->lvl_begin, c->lvl_end); return DOTS_ERR_BAD_ARG; }
    return 0;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Grid of the TMA-staged triangle kernel (iter_kernels.cu: k_tri_tma, 128 triangles per block): time levels per block and
// the number of blocks.  16 levels amortise the per-triangle constants; small meshes (fewer blocks than two waves of
// 3 blocks / SM) take the shortest chunk that still fits ONE wave: a knots_5-class mesh (68 triangle tiles, 32 levels) runs
// 6 levels per block on 408 blocks instead of 16 on 136.  Shared with the reduction of the per-block KKT #1 partials.
static inline int dots_tri_tma_blocks(const dots_ctx_t *c, int *tch_out)
{
    const int tiles = ceil_div(c->n_tri, 128), levels = c->lvl_end - c->lvl_begin;
    const int wave = 3 * (c->n_sm > 0 ? c->n_sm : 148);
    int tch = 16;
    if ((long long)tiles * ceil_div(levels, 16) < 2LL * wave) {
        tch = 2;
        while (tch < 16 && (long long)tiles * ceil_div(levels, tch) > wave) ++tch;
    }
    if (tch_out) *tch_out = tch;
    return tiles * ceil_div(levels, tch);
}
static inline int dots_t_end(const dots_ctx_t *c) { return c->lvl_end < c->n_time ? c->lvl_end : c->n_time; }   // staggered steps owned: [lvl_begin, t_end)

// Programmatic dependent launch: the launches of one iteration are chained (the attribute is passed when ctx.ring_pdl is
// set).  A block announces itself at once, reads only constant data (index lists, mesh constants, Q, factor panels) and then
// waits for the previous launch to finish and flush; every block executes the wait, so "this grid has completed" still
// implies "all earlier grids have completed" for whatever follows (normal launches, NCCL kernels, graph ends).  Nothing is
// written before the wait.  The two instructions are no-ops in a kernel that was launched normally.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... Args>
static inline int pdl_launch2(void (*kern)(Args...), dim3 grid, int threads, cudaStream_t st, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)threads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    DOTS_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    return 0;
}

template <typename... Args>
static inline int pdl_launch(void (*kern)(Args...), int grid, int threads, size_t smem, cudaStream_t st, bool pdl, Args... args)
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

// Corners around a vertex, ELL form: vc_ell[v][8] = the first 8 corner ids (k*T + f) of the CSR list, -1 padded; row[7] == -2
// marks a vertex with more than 8 corners (the CSR list is walked instead).  Two aligned 16-byte loads give a thread all its
// indices at once, so the 6-16 value gathers of a (t, v) pair are independent loads in flight together instead of a
// pointer-chasing loop (ncu on the CSR form: long_scoreboard-bound, k_vertex 0.76 / k_phi_rhs 0.65 of the HBM peak).
// The summation order is the CSR order (padding adds +0.0 at the end): bit-identical results.
struct vc_row {
    int id[8];
};
__device__ __forceinline__ vc_row vc_load(const dots_ctx_t &c, int v)
{
    const int4 *p = reinterpret_cast<const int4 *>(c.vc_ell) + 2 * (size_t)v;
    const int4 a = p[0], b = p[1];
    vc_row r;
    r.id[0] = a.x; r.id[1] = a.y; r.id[2] = a.z; r.id[3] = a.w; r.id[4] = b.x; r.id[5] = b.y; r.id[6] = b.z; r.id[7] = b.w;
    return r;
}

// np.clip semantics: NaN passes through (fmin/fmax would drop it)
__device__ __forceinline__ double clip01(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Fixed-order block reduction of NS partial sums per thread; result row written to part[blockIdx.x*8 + slot0 + i].
template <int NS>
__device__ __forceinline__ void block_reduce_store(double (&acc)[NS], double *part, int slot0)
{
    __shared__ double sm[32][NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        double v = warp_sum(acc[i]);
        if (lane == 0) sm[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double v = 0.0;
        for (int w = 0; w < nw; ++w) v += sm[w][threadIdx.x];
        part[(size_t)blockIdx.x * 8 + slot0 + threadIdx.x] = v;
    }
}

// ---- mbarrier + TMA (1-D bulk async copy) helpers: cp.async.bulk -> SASS UBLKCP -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
// L2 eviction policy for data that is streamed once (factor panels): keeps the vectors the sweeps re-read resident
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// global -> shared bulk copy; bytes and both addresses must be multiples of 16
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
