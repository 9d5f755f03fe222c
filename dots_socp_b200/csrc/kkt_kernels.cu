// Weighted-norm reductions behind the seven KKT residuals and the objective
// (reference: socp/solver_socp.py:417-559, wiring :589-643, norm_square_weight :875-878).
//
// The reference evaluates the conditions lazily (utils/condition_validator.py:236-331), one closure at
// a time, and every closure materialises its residual arrays before reducing them.  Here each
// condition is one pass of a vertex kernel and/or a triangle kernel that forms the residual in
// registers and reduces `a^2 * weight` on the fly; block partials are combined in a fixed order so the
// result is bit-reproducible for a given launch shape.  Only the raw sums travel to the host (<= 8
// doubles); the square roots / normalisations of :433-559 are host arithmetic on scalars.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <dlfcn.h>

static thread_local char g_err[512] = "";
void dots_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char *dots_last_error(void) { return g_err; }
extern "C" int dots_abi_version(void) { return DOTS_ABI_VERSION; }
extern "C" int dots_ctx_sizeof(void) { return (int)sizeof(dots_ctx_t); }

// Let kernels of the CURRENT device dereference memory of `peer_device` (CUDA IPC mappings of other ranks' buffers).
extern "C" int dots_enable_peer(int peer_device)
{
    int me = -1, can = 0;
    DOTS_CUDA(cudaGetDevice(&me));
    if (me == peer_device) return 0;
    DOTS_CUDA(cudaDeviceCanAccessPeer(&can, me, peer_device));
    if (!can) { dots_set_error("device %d cannot access device %d", me, peer_device); return DOTS_ERR_BAD_ARG; }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) { dots_set_error("cudaDeviceEnablePeerAccess(%d) -> %s", peer_device, cudaGetErrorString(e)); return (int)e; }
    return 0;
}

// CUDA IPC export / import of a device buffer (peer-memory exchange between the ranks of one node).
// export: handle of the allocation that contains `dev_ptr` + the byte offset of `dev_ptr` inside it.
extern "C" int dots_ipc_export(const void *dev_ptr, void *handle_out_64, unsigned long long *offset_out)
{
    // the driver entry point is resolved at run time so that the library itself does not link against libcuda
    typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
    static range_fn get_range = nullptr;
    if (!get_range) {
        void *drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (drv) get_range = (range_fn)dlsym(drv, "cuMemGetAddressRange_v2");
        if (!get_range) { dots_set_error("cuMemGetAddressRange_v2 not available"); return DOTS_ERR_BAD_ARG; }
    }
    unsigned long long base = 0;
    size_t size = 0;
    if (get_range(&base, &size, (unsigned long long)dev_ptr) != 0) { dots_set_error("cuMemGetAddressRange failed"); return DOTS_ERR_BAD_ARG; }
    cudaIpcMemHandle_t h;
    DOTS_CUDA(cudaIpcGetMemHandle(&h, (void *)base));
    memcpy(handle_out_64, &h, sizeof(h));
    *offset_out = (unsigned long long)dev_ptr - base;
    return 0;
}
// import in the CURRENT device's context (lazy peer access), so kernels of this device may dereference the result
extern "C" int dots_ipc_import(const void *handle_64, unsigned long long offset, void **dev_ptr_out)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64, sizeof(h));
    void *base = nullptr;
    DOTS_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr_out = (char *)base + offset;
    return 0;
}

#define KKT_THREADS 256
#define KKT_CONDS 9            // 7 KKT conditions + the objective (7) + the variable norms of scale_prim_dual (8)

// ---- vertex-side terms of condition W at (t, v): acc[0..3] += ...  (slots 0..3 of the condition's row) ---------------
// ps / ds: prim_scale / dual_scale (1 unless is_constant_scaling): conditions 4-6 and the objective are formed from the
// UN-scaled variables (dual_scale r) mu, prim_scale A, ... (solver_socp.py:617-637, :829-831)
template <int W>
__device__ __forceinline__ void kkt_vertex_terms(const dots_ctx_t &c, size_t i, int t, int v, double av, double r, double s, double d,
                                                 double cong, double dt, double ps, double ds, double (&acc)[4])
{
    const int V = c.n_vert, nT = c.n_time;
    const size_t T = (size_t)c.n_tri;
    if (W != 2 && t >= nT) return;                                         // staggered arrays have nT steps; #2 lives on the nT+1 levels
    if (W == 0) {                                                          // Prim(phi, q)  :433-450, :591-596
        const double dtp = (c.phi[i + V] - c.phi[i]) / dt;
        const double A = c.A[i], lc = c.lam_c[i];
        const double res = dtp - A - lc;
        acc[0] += res * res * av; acc[1] += dtp * dtp * av; acc[2] += A * A * av; acc[3] += lc * lc * av;
    } else if (W == 1) {                                                   // Prim(q, z)    :452-464, :597-603
        const double A = c.A[i];
        const double r1 = c.z_fst[i] + s * A - d, r2 = c.z_end[i] - s * A - d;
        acc[0] += r1 * r1 * av; acc[1] += r2 * r2 * av;
    } else if (W == 2) {                                                   // Dual(alpha)   :466-482
        double divt;
        if (t == 0) divt = (c.mu[v] * av) / dt;
        else if (t == nT) divt = -(c.mu[(size_t)(nT - 1) * V + v] * av) / dt;
        else divt = (c.mu[(size_t)t * V + v] * av - c.mu[(size_t)(t - 1) * V + v] * av) / dt;
        const double *Et = c.E + (size_t)t * 3 * T;
        double divx = 0.0;
        auto corner = [&](int cid) -> double {                             // -(g_k . (E af)) of corner id = k T + f
            const size_t k = (size_t)(cid >= (int)T) + (size_t)(cid >= 2 * (int)T), f = cid - k * T;
            const double af = c.area_f[f];
            return -(c.hat_grad[(k * 3 + 0) * T + f] * (Et[f] * af) + c.hat_grad[(k * 3 + 1) * T + f] * (Et[T + f] * af)
                     + c.hat_grad[(k * 3 + 2) * T + f] * (Et[2 * T + f] * af));
        };
        const vc_row vr = vc_load(c, v);                                   // all corner ids at once: independent gathers
        if (vr.id[7] != -2) {
            double val[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) val[q] = vr.id[q] >= 0 ? corner(vr.id[q]) : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) if (vr.id[q] >= 0) divx += val[q];
        } else {
            for (int q = c.vc_ptr[v], qe = c.vc_ptr[v + 1]; q < qe; ++q) divx += corner(c.vc_idx[q]);
        }
        const double bnd = (t == 0) ? c.bnd0[v] : ((t == nT) ? c.bnd1[v] : 0.0);
        const double aux = (r * dt) * ((bnd + divt + divx) / av);
        acc[0] += aux * aux * av;
    } else if (W == 3) {                                                   // Dual(beta)    :484-503
        const double mu = c.mu[i];
        const double a1 = s * (c.b_end[i] - c.b_fst[i]);
        const double sum = mu + a1;
        acc[0] += mu * mu * av; acc[1] += a1 * a1 * av; acc[2] += sum * sum * av;
    } else if (W == 4) {                                                   // Comp(rho, f(q)) :505-526
        const double c3 = 1.0 / sqrt(3.0);
        const double *B0 = c.B + (size_t)t * 3 * T, *B1 = B0 + 3 * T;
        double q = 0.0;
        for (int p = c.vc_ptr[v], pe = c.vc_ptr[v + 1]; p < pe; ++p) {
            const int cid = c.vc_idx[p];
            const size_t k = (size_t)(cid >= (int)T) + (size_t)(cid >= 2 * (int)T), f = cid - k * T;
            double sq = 0.0;
#pragma unroll
            for (int x = 0; x < 3; ++x) { const double a = c3 * (ps * B0[x * T + f]); sq += a * a; }
#pragma unroll
            for (int x = 0; x < 3; ++x) { const double a = c3 * (ps * B1[x * T + f]); sq += a * a; }
            q += c.area_f[f] * sq;
        }
        const double rho = (ds * r) * c.mu[i];
        const double aux = ps * c.A[i] + .25 * (q / av);
        double pos = aux + rho;
        pos = (pos < 0.) ? 0. : pos;                                       // np.maximum(0., .) keeps NaN
        const double res = pos - rho;
        acc[0] += rho * rho * av; acc[1] += aux * aux * av; acc[2] += res * res * av;
    } else if (W == 6) {                                                   // Comp(rho, cong.) :549-559
        const double rho = (ds * r) * c.mu[i], lc = ps * c.lam_c[i];
        const double res = cong * rho - lc;
        acc[0] += rho * rho * av; acc[1] += lc * lc * av; acc[2] += res * res * av;
    } else if (W == 7) {                                                   // objective :417-431
        if (t == 0) acc[0] += (ps * c.phi[v]) * ((ds * r) * c.bnd0[v]);
        if (t == nT - 1) acc[1] += (ps * c.phi[(size_t)nT * V + v]) * ((ds * r) * c.bnd1[v]);
        const double lc = ps * c.lam_c[i];
        acc[2] += lc * lc * av;
    } else if (W == 8) {                                                   // variable norms of scale_prim_dual :331-340
        const double zf = c.z_fst[i], ze = c.z_end[i], bf = c.b_fst[i], be = c.b_end[i];
        acc[0] += zf * zf * av; acc[1] += ze * ze * av; acc[2] += bf * bf * av; acc[3] += be * be * av;
    }
}

// ---- triangle-side terms of condition W at (tau, f): slots 4..7 of the condition's row ------------------------------
template <int W>
__device__ __forceinline__ void kkt_tri_terms(const dots_ctx_t &c, int tau, size_t f, double af, double r, double s, double cs,
                                              double ps, double ds, double (&acc)[4])
{
    const int V = c.n_vert, nT = c.n_time;
    const size_t T = (size_t)c.n_tri;
    const double *Bp = c.B + (size_t)tau * 3 * T + f;
    if (W == 0) {                                                          // ||dx_phi - B||, ||dx_phi||, ||B||
        const double *ph = c.phi + (size_t)tau * V;
        const double p0 = ph[c.tri[f]], p1 = ph[c.tri[T + f]], p2 = ph[c.tri[2 * T + f]];
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            const double dx = c.hat_grad[(0 * 3 + x) * T + f] * p0 + c.hat_grad[(1 * 3 + x) * T + f] * p1 + c.hat_grad[(2 * 3 + x) * T + f] * p2;
            const double B = Bp[x * T], res = dx - B;
            acc[0] += res * res * af; acc[1] += dx * dx * af; acc[2] += B * B * af;
        }
    } else if (W == 1) {                                                   // ||s (z_mid - Bd(B))||_dec  :599
        const double *zm = c.z_mid + (size_t)tau * 18 * T + f;
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            if (sd == 0 ? tau >= nT : tau <= 0) continue;
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    const double res = s * (zm[((sd * 3 + k) * 3 + x) * T] - cs * Bp[x * T]);
                    acc[0] += res * res * af;
                }
        }
    } else if (W == 3) {                                                   // ||E||, ||adj(b_mid)||, ||E + adj(b_mid)||
        const double *bm = c.b_mid + (size_t)tau * 18 * T + f;
        const double *Ep = c.E + (size_t)tau * 3 * T + f;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            double a2 = 0.0;
            if (tau < nT) a2 = cs * ((bm[(0 * 3 + x) * T] + bm[(1 * 3 + x) * T]) + bm[(2 * 3 + x) * T]);
            if (tau > 0) {
                const double s1 = cs * ((bm[(9 + 0 * 3 + x) * T] + bm[(9 + 1 * 3 + x) * T]) + bm[(9 + 2 * 3 + x) * T]);
                a2 = (tau < nT) ? a2 + s1 : s1;
            }
            const double E = Ep[x * T], sum = E + a2;
            acc[0] += E * E * af; acc[1] += a2 * a2 * af; acc[2] += sum * sum * af;
        }
    } else if (W == 5) {                                                   // Comp(m, rho o B) :528-547
        double avg = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int v = c.tri[k * T + f];
            const double cur = (tau < nT) ? (ds * r) * c.mu[(size_t)tau * V + v] : 0.0;
            const double prv = (tau > 0) ? (ds * r) * c.mu[(size_t)(tau - 1) * V + v] : 0.0;
            avg += (1.0 / 3.0) * (0.5 * cur + 0.5 * prv);                  // :961-974, :163-166
        }
        const double *Ep = c.E + (size_t)tau * 3 * T + f;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            const double m = (ds * r) * Ep[x * T], aux = avg * (ps * Bp[x * T]), res = aux - m;
            acc[0] += m * m * af; acc[1] += aux * aux * af; acc[2] += res * res * af;
        }
    } else if (W == 8) {                                                   // ||z_mid||^2_dec, ||b_mid||^2_dec (absent side slots are zero)
        const double *zm = c.z_mid + (size_t)tau * 18 * T + f;
        const double *bm = c.b_mid + (size_t)tau * 18 * T + f;
#pragma unroll
        for (int p = 0; p < 18; ++p) {
            if (p < 9 ? tau >= nT : tau <= 0) continue;
            const double zz = zm[p * T], bb = bm[p * T];
            acc[0] += zz * zz * af; acc[1] += bb * bb * af;
        }
    }
}

// Fixed-order block reduction of the 4 partial sums of every condition in `mask`; row blockIdx.x of part[.][64].
__device__ __forceinline__ void kkt_block_store(double (&acc)[KKT_CONDS][4], unsigned mask, double *part, int slot0)
{
    __shared__ double sm[KKT_THREADS / 32][KKT_CONDS * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int w = 0; w < KKT_CONDS; ++w) {
        if (!(mask >> w & 1u)) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double v = warp_sum(acc[w][k]);
            if (lane == 0) sm[warp][w * 4 + k] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x < KKT_CONDS * 4 && (mask >> (threadIdx.x >> 2) & 1u)) {
        double v = 0.0;
        for (int w = 0; w < KKT_THREADS / 32; ++w) v += sm[w][threadIdx.x];
        part[(size_t)blockIdx.x * (KKT_CONDS * 8) + (threadIdx.x >> 2) * 8 + slot0 + (threadIdx.x & 3)] = v;
    }
}

// One pass over the (time, vertex) pairs for ALL conditions in `mask`: shared operands (A, mu, lam_c, phi, ...) are loaded once.
// CMASK != 0: the mask is a compile-time constant (the sets the solver asks for most: #2 alone on the lazy checks, #0-#3 on
// the forced ones), so the untaken conditions cost neither registers nor branches; the mapping of items to threads and
// blocks, hence the summation order and the sums, are the same in every instantiation.
template <unsigned CMASK>
__global__ void __launch_bounds__(KKT_THREADS) k_kkt_vertex(dots_ctx_t c, unsigned mask_rt)
{
    const unsigned mask = CMASK ? CMASK : mask_rt;
    const int V = c.n_vert, nT = c.n_time;
    const double *prm = c.params;
    const double r = prm[DOTS_P_R], s = prm[DOTS_P_S], d = prm[DOTS_P_D], cong = prm[DOTS_P_CONG];
    const double ps = prm[DOTS_P_PS], ds = prm[DOTS_P_DS];
    const double dt = 1.0 / nT;
    double acc[KKT_CONDS][4];
#pragma unroll
    for (int w = 0; w < KKT_CONDS; ++w)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[w][k] = 0.0;
    // owned range: levels [lvl_begin, lvl_end) when the centred residual (#2) is wanted, staggered steps otherwise
    const int t_end = (mask >> 2 & 1u) ? c.lvl_end : min(c.lvl_end, nT);
    const size_t i0 = (size_t)c.lvl_begin * V, n = (size_t)max(t_end, c.lvl_begin) * V;
    for (size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / V), v = (int)(i - (size_t)t * V);
        const double av = c.area_v[v];
        if (mask & 1u) kkt_vertex_terms<0>(c, i, t, v, av, r, s, d, cong, dt, ps, ds, acc[0]);
        if (mask & 2u) kkt_vertex_terms<1>(c, i, t, v, av, r, s, d, cong, dt, ps, ds, acc[1]);
        if (mask & 4u) kkt_vertex_terms<2>(c, i, t, v, av, r, s, d, cong, dt, ps, ds, acc[2]);
        if (mask & 8u) kkt_vertex_terms<3>(c, i, t, v, av, r, s, d, cong, dt, ps, ds, acc[3]);
        if (mask & 16u) kkt_vertex_terms<4>(c, i, t, v, av, r, s, d, cong, dt, ps, ds, acc[4]);
        if (mask & 64u) kkt_vertex_terms<6>(c, i, t, v, av, r, s, d, cong, dt, ps, ds, acc[6]);
        if (mask & 128u) kkt_vertex_terms<7>(c, i, t, v, av, r, s, d, cong, dt, ps, ds, acc[7]);
        if (mask & 256u) kkt_vertex_terms<8>(c, i, t, v, av, r, s, d, cong, dt, ps, ds, acc[8]);
    }
    kkt_block_store(acc, mask & 0x1dfu, c.red_part, 0);
}

// Dual(alpha) (#2, :466-482) alone, the residual the lazy checks ask for most.  Thread = (vertex, KKT_TC consecutive time
// levels): the corner ids, the hat gradients and the areas of a vertex are gathered once and serve all levels of the chunk,
// so only E is gathered per level (the per-(t, v) kernel re-gathers 4 of 7 values per corner on every level and is bound by the
// L2 sector traffic of those gathers).  Per (t, v) the arithmetic and the corner order are those of kkt_vertex_terms<2>.
#define KKT_TC 8
__global__ void __launch_bounds__(KKT_THREADS, 4) k_kkt_dual_alpha(dots_ctx_t c)      // 4 blocks per SM = the persistent grid in one wave
{
    const int V = c.n_vert, nT = c.n_time;
    const size_t T = (size_t)c.n_tri;
    const double r = c.params[DOTS_P_R];
    const double dt = 1.0 / nT;
    const int l0 = c.lvl_begin, levels = c.lvl_end - c.lvl_begin;
    const int n_chunks = (levels + KKT_TC - 1) / KKT_TC;
    const size_t n_items = (size_t)n_chunks * V;
    double total = 0.0;
    for (size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(item / V), v = (int)(item - (size_t)ch * V);
        const int t0 = l0 + ch * KKT_TC, nt = min(KKT_TC, c.lvl_end - t0);
        const double av = c.area_v[v];
        double divx[KKT_TC];
#pragma unroll
        for (int j = 0; j < KKT_TC; ++j) divx[j] = 0.0;
        auto corner = [&](int cid) {                                       // adds -(g_k . (E af)) of corner id = k T + f to every level
            const size_t k = (size_t)(cid >= (int)T) + (size_t)(cid >= 2 * (int)T), f = cid - k * T;
            const double af = c.area_f[f];
            const double g0 = c.hat_grad[(k * 3 + 0) * T + f], g1 = c.hat_grad[(k * 3 + 1) * T + f], g2 = c.hat_grad[(k * 3 + 2) * T + f];
#pragma unroll
            for (int j = 0; j < KKT_TC; ++j) {
                if (j < nt) {
                    const double *Et = c.E + (size_t)(t0 + j) * 3 * T;
                    divx[j] += -(g0 * (Et[f] * af) + g1 * (Et[T + f] * af) + g2 * (Et[2 * T + f] * af));
                }
            }
        };
        const vc_row vr = vc_load(c, v);
        if (vr.id[7] != -2) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (vr.id[q] >= 0) corner(vr.id[q]);
        } else {
            for (int q = c.vc_ptr[v], qe = c.vc_ptr[v + 1]; q < qe; ++q) corner(c.vc_idx[q]);
        }
        double mu_prev = (t0 > 0) ? c.mu[(size_t)(t0 - 1) * V + v] : 0.0;
#pragma unroll
        for (int j = 0; j < KKT_TC; ++j) {
            if (j < nt) {
                const int t = t0 + j;
                const double mu_cur = (t < nT) ? c.mu[(size_t)t * V + v] : 0.0;
                double divt;
                if (t == 0) divt = (mu_cur * av) / dt;
                else if (t == nT) divt = -(mu_prev * av) / dt;
                else divt = (mu_cur * av - mu_prev * av) / dt;
                const double bnd = (t == 0) ? c.bnd0[v] : ((t == nT) ? c.bnd1[v] : 0.0);
                const double aux = (r * dt) * ((bnd + divt + divx[j]) / av);
                total += aux * aux * av;
                mu_prev = mu_cur;
            }
        }
    }
    __shared__ double sm[KKT_THREADS / 32];
    const double w = warp_sum(total);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        double vsum = 0.0;
        for (int i = 0; i < KKT_THREADS / 32; ++i) vsum += sm[i];
        c.red_part[(size_t)blockIdx.x * (KKT_CONDS * 8) + 2 * 8 + 0] = vsum;
    }
}

template <unsigned CMASK>
__global__ void __launch_bounds__(KKT_THREADS) k_kkt_tri(dots_ctx_t c, unsigned mask_rt)
{
    const unsigned mask = CMASK ? CMASK : mask_rt;
    const size_t T = (size_t)c.n_tri;
    const double *prm = c.params;
    const double r = prm[DOTS_P_R], s = prm[DOTS_P_S], ps = prm[DOTS_P_PS], ds = prm[DOTS_P_DS];
    const double cs = s / sqrt(3.0);
    double acc[KKT_CONDS][4];
#pragma unroll
    for (int w = 0; w < KKT_CONDS; ++w)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[w][k] = 0.0;
    const size_t i0 = (size_t)c.lvl_begin * T, n = (size_t)c.lvl_end * T;
    for (size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int tau = (int)(i / T);
        const size_t f = i - (size_t)tau * T;
        const double af = c.area_f[f];
        if (mask & 1u) kkt_tri_terms<0>(c, tau, f, af, r, s, cs, ps, ds, acc[0]);
        if (mask & 2u) kkt_tri_terms<1>(c, tau, f, af, r, s, cs, ps, ds, acc[1]);
        if (mask & 8u) kkt_tri_terms<3>(c, tau, f, af, r, s, cs, ps, ds, acc[3]);
        if (mask & 32u) kkt_tri_terms<5>(c, tau, f, af, r, s, cs, ps, ds, acc[5]);
        if (mask & 256u) kkt_tri_terms<8>(c, tau, f, af, r, s, cs, ps, ds, acc[8]);
    }
    kkt_block_store(acc, mask & 0x12bu, c.red_part, 4);
}

// one warp per (condition, slot) pair, 8 pairs per block; fixed order over blocks
__global__ void k_reduce_final(const double *__restrict__ part, int nblocks, double *__restrict__ out)
{
    const int slot = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    double v = 0.0;
    for (int b = lane; b < nblocks; b += 32) v += part[(size_t)b * (KKT_CONDS * 8) + slot];
    v = warp_sum(v);
    if (lane == 0) out[slot] = v;
}

// Triangle term of KKT #1 from the per-block partials an iteration with write_z = 2 left in kkt1_part: one block, fixed order.
__global__ void __launch_bounds__(256) k_reduce_kkt1(const double *__restrict__ part, int n, double *__restrict__ out)
{
    __shared__ double sm[8];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) v += part[i];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        out[1 * 8 + 4] = t;                                               // condition 1, first triangle slot
    }
}

// Raw sums of every condition in `mask` (bit i = condition i, bit 7 = objective) in ONE pass per side: host_out[8 * i + k].
extern "C" int dots_kkt_sums_multi(const dots_ctx_t *c, unsigned mask, double *host_out, void *stream)
{
    if (int e = dots_check_ctx(c)) return e;
    if (!mask || mask > 0x3ffu) { dots_set_error("kkt mask %u out of range", mask); return DOTS_ERR_BAD_ARG; }
    const bool fused1 = (mask & 0x200u) != 0;                              // triangle term of #1 from kkt1_part
    mask &= 0x1ffu;
    int n_part = 0;
    if (fused1) {
        if (!(mask & 2u)) { dots_set_error("kkt mask: bit 9 without condition 1"); return DOTS_ERR_BAD_ARG; }
        n_part = dots_tri_tma_blocks(c, nullptr);
        if (!c->kkt1_part || c->kkt1_blocks < n_part || (c->ring_flags & 4)) { dots_set_error("kkt mask: bit 9 needs kkt1_part from dots_step_tri(write_z = 2)"); return DOTS_ERR_BAD_ARG; }
    }
    const unsigned tmask = (mask & 0x12bu) & ~(fused1 ? 2u : 0u);
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = c->red_blocks;
    DOTS_CUDA(cudaMemsetAsync(c->red_part, 0, sizeof(double) * KKT_CONDS * 8 * nb, st));
    if (mask == 4u) k_kkt_dual_alpha<<<nb, KKT_THREADS, 0, st>>>(*c);                           // lazy check of Dual(alpha) alone
    else if (mask == 15u) {                                                                     // forced primal + dual set
        k_kkt_vertex<11u><<<nb, KKT_THREADS, 0, st>>>(*c, 11u);
        k_kkt_dual_alpha<<<nb, KKT_THREADS, 0, st>>>(*c);
    } else if (mask & 0x1dfu) k_kkt_vertex<0u><<<nb, KKT_THREADS, 0, st>>>(*c, mask);
    DOTS_LAUNCH_CHECK();
    if (tmask == 11u) k_kkt_tri<11u><<<nb, KKT_THREADS, 0, st>>>(*c, tmask);                     // #0, #1, #3 have triangle terms
    else if (tmask == 9u) k_kkt_tri<9u><<<nb, KKT_THREADS, 0, st>>>(*c, tmask);                  // the same set with #1 taken from kkt1_part
    else if (tmask) k_kkt_tri<0u><<<nb, KKT_THREADS, 0, st>>>(*c, tmask);
    DOTS_LAUNCH_CHECK();
    k_reduce_final<<<KKT_CONDS, 256, 0, st>>>(c->red_part, nb, c->red_out);
    DOTS_LAUNCH_CHECK();
    if (fused1) { k_reduce_kkt1<<<1, 256, 0, st>>>(c->kkt1_part, n_part, c->red_out); DOTS_LAUNCH_CHECK(); }
    DOTS_CUDA(cudaMemcpyAsync(host_out, c->red_out, sizeof(double) * KKT_CONDS * 8, cudaMemcpyDeviceToHost, st));
    DOTS_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int dots_kkt_sums(const dots_ctx_t *c, int which, double *host_out, void *stream)
{
    if (which < 0 || which > 7) { dots_set_error("kkt condition %d out of range", which); return DOTS_ERR_BAD_ARG; }
    double all[KKT_CONDS * 8];
    if (int e = dots_kkt_sums_multi(c, 1u << which, all, stream)) return e;
    for (int k = 0; k < 8; ++k) host_out[k] = all[which * 8 + k];
    return 0;
}
