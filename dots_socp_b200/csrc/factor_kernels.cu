// Setup path (row f1 of SURVEY.md section 8): numeric multifrontal factorisation of the small fronts on the GPU.
//
// The reference factorises L + (lambda_a - eps) diag(area_v) once per time mode with SuperLU on the host
// (utils/laplacian_inverse_socp.py:34-41, ~8.5 min at V = 164k).  Here every separator-tree node whose front fits in
// shared memory (n = |S| + |B| <= FRONT_NMAX; ~93 % of the nodes, everything near the leaves) is handled by one block
// per (node, time mode):
//     assemble  F = K[front, front] restricted to rows/cols touching S  +  shift_mode * diag(mass_S)  +  children's updates
//     partial Cholesky of the first s columns (right-looking, in shared memory)   F11 = L11 L11^T, L21 = F21 L11^-T
//     in-place inverse of L11, W21 = L21 inv(L11)
//     write the solve-ready panel  [inv(L11) ; W21]  in both layouts (row-major `panels`, column-major `panels_t`)
//     write the update matrix  U = F22 - L21 L21^T  for the parent.
// The few large fronts near the root go through batched dense library calls (dots_socp_b200/nested.py).
#include "common.cuh"

#define FRONT_NMAX 96
#define FRONT_THREADS 256

__global__ void __launch_bounds__(FRONT_THREADS) k_front_small(dots_front_args_t a)
{
    extern __shared__ double F[];                              // [n][ld]
    const int node = a.nodes[blockIdx.x];
    const int mode = blockIdx.y;
    const int s = a.nd_s[node], b = a.nd_b[node], n = s + b, ld = n + 1;
    const int off = a.nd_off[node];
    const int M = a.m_pad;
    const int tid = threadIdx.x;
    const size_t pbase = (size_t)a.nd_panel[node];
    const size_t ntri = (size_t)s * (s + 1) / 2;

    if (mode >= a.n_modes) {                                   // padding modes: identity-like panel
        for (int j = tid; j < s; j += FRONT_THREADS) {
            a.panels[(pbase + (size_t)j * (j + 1) / 2 + j) * M + mode] = 1.0;
            a.panels_t[(pbase + (size_t)j * (s + b) - (size_t)j * (j - 1) / 2) * M + mode] = 1.0;
        }
        return;
    }
    for (int i = tid; i < n * ld; i += FRONT_THREADS) F[i] = 0.0;
    __syncthreads();
    // ---- assemble: matrix entries of the rows owned by this node (mirrored), shifted mass on the diagonal
    const double shift = a.shifts[mode];
    for (int r = 0; r < s; ++r) {
        const int64_t q0 = a.a_ptr[off + r], q1 = a.a_ptr[off + r + 1];
        for (int64_t q = q0 + tid; q < q1; q += FRONT_THREADS) {
            const int c = a.a_pos[q];
            if (c >= 0) {
                const double v = a.a_val[q];
                F[r * ld + c] = v;
                F[c * ld + r] = v;
            }
        }
    }
    __syncthreads();
    // the pin of the singular mode goes through the SAME thread that owns the last diagonal entry: a second thread
    // updating F[s-1][s-1] here raced with it (lost update -> zero pivot -> NaN in mode 0 on small meshes)
    const bool pinned = node == a.pin_node && shift == 0.0;
    for (int r = tid; r < s; r += FRONT_THREADS)
        F[r * ld + r] += shift * a.mass[off + r] + ((pinned && r == s - 1) ? a.pin_value : 0.0);
    __syncthreads();
    // ---- extend-add the children's update matrices
    for (int slot = 0; slot < 2; ++slot) {
        const int ch = a.nd_child[2 * node + slot];
        if (ch < 0) continue;
        const int bc = a.nd_b[ch];
        const double *U = reinterpret_cast<const double *>(a.u_ptr[ch]);
        if (!bc || !U) continue;
        const int32_t *pp = a.parent_pos + a.nd_upd[ch];
        for (int i = tid; i < bc * bc; i += FRONT_THREADS) {
            const int p = i / bc, q = i - p * bc;
            F[pp[p] * ld + pp[q]] += U[(size_t)i * M + mode];
        }
        __syncthreads();
    }
    // ---- right-looking partial Cholesky of the first s columns (lower triangle + full trailing block)
    for (int k = 0; k < s; ++k) {
        const double d = sqrt(F[k * ld + k]);
        __syncthreads();
        for (int i = k + tid; i < n; i += FRONT_THREADS) F[i * ld + k] = (i == k) ? d : F[i * ld + k] / d;
        __syncthreads();
        const int rem = n - k - 1;
        for (int i = tid; i < rem * rem; i += FRONT_THREADS) {
            const int r = k + 1 + i / rem, c = k + 1 + i % rem;
            if (c <= r || r >= s) F[r * ld + c] -= F[r * ld + k] * F[c * ld + k];   // lower part of the pivot block, all of the rest
        }
        __syncthreads();
    }
    // ---- in-place inverse of the lower-triangular L11 (column by column)
    for (int j = tid; j < s; j += FRONT_THREADS) {
        // column j of inv(L11): x_j = 1/L_jj, x_i = -(sum_{k=j..i-1} L_ik x_k) / L_ii   (uses only column j's own results)
        // done sequentially per thread in registers-free form: results are written to the strictly upper part as scratch
        const double xj = 1.0 / F[j * ld + j];
        // scratch row: upper part of row j (columns j+1..s-1) holds x_i for i > j
        for (int i = j + 1; i < s; ++i) {
            double acc = F[i * ld + j] * xj;
            for (int k = j + 1; k < i; ++k) acc += F[i * ld + k] * F[j * ld + k];
            F[j * ld + i] = -acc / F[i * ld + i];
        }
    }
    __syncthreads();
    // move the scratch (upper part, row j = column j of the inverse) into place: Linv[i][j] for i > j, then the diagonal
    for (int i = tid; i < s * s; i += FRONT_THREADS) {
        const int r = i / s, c = i - r * s;
        if (r > c) F[r * ld + c] = F[c * ld + r];
    }
    __syncthreads();
    for (int j = tid; j < s; j += FRONT_THREADS) F[j * ld + j] = 1.0 / F[j * ld + j];
    __syncthreads();
    // ---- update matrix for the parent: U = trailing block (already F22 - L21 L21^T), written before L21 is overwritten
    if (b) {
        double *U = reinterpret_cast<double *>(a.u_ptr[node]);
        for (int i = tid; i < b * b; i += FRONT_THREADS) {
            const int p = i / b, q = i - p * b;
            const double v = (q <= p) ? F[(s + p) * ld + s + q] : F[(s + q) * ld + s + p];
            U[(size_t)i * M + mode] = v;
        }
    }
    // ---- W21 = L21 inv(L11), row by row in place (row i only needs its own old values)
    for (int i = tid; i < b; i += FRONT_THREADS) {
        double *row = F + (s + i) * ld;
        for (int j = 0; j < s; ++j) {
            double acc = 0.0;
            for (int k = j; k < s; ++k) acc += row[k] * F[k * ld + j];
            row[j] = acc;                                      // entries k > j of the row are still the old L21 values
        }
    }
    __syncthreads();
    // ---- solve-ready panel in both layouts
    for (int i = tid; i < n * s; i += FRONT_THREADS) {
        const int r = i / s, c = i - r * s;
        if (r < s && c > r) continue;
        const double v = F[r * ld + c];
        const size_t fo = (r < s) ? (size_t)r * (r + 1) / 2 + c : ntri + (size_t)(r - s) * s + c;
        const size_t to = (size_t)c * (s + b) - (size_t)c * (c - 1) / 2 + (r - c);
        a.panels[(pbase + fo) * M + mode] = v;
        a.panels_t[(pbase + to) * M + mode] = v;
    }
}

extern "C" int dots_factor_small_fronts(const dots_front_args_t *a, int n_launch, int max_front, void *stream)
{
    if (!a || n_launch <= 0) return 0;
    if (max_front > FRONT_NMAX) { dots_set_error("front of %d rows exceeds FRONT_NMAX=%d", max_front, FRONT_NMAX); return DOTS_ERR_BAD_ARG; }
    const size_t smem = (size_t)max_front * (max_front + 1) * sizeof(double);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        DOTS_CUDA(cudaFuncSetAttribute(k_front_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(n_launch, a->m_pad);
    k_front_small<<<grid, FRONT_THREADS, smem, (cudaStream_t)stream>>>(*a);
    DOTS_LAUNCH_CHECK();
    return 0;
}

extern "C" int dots_front_nmax(void) { return FRONT_NMAX; }
