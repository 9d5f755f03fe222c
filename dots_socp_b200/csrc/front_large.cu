// Setup path (row f1 of SURVEY.md section 8): numeric multifrontal factorisation of the LARGE fronts (more rows than
// k_front_small keeps in shared memory; 551 fronts of up to 1 277 rows on the 164k-vertex headline mesh), batched over the
// nodes of one tree level and all time modes.  Replaces the per-mode SuperLU factorisations of the reference
// (utils/laplacian_inverse_socp.py:34-41) and the batched library calls (cuSOLVER / cuBLAS through torch.linalg) this
// setup step used before.
//
// Every (node, mode) pair owns a dense front F (n x n, row-major, n = |S| + |B|) and a panel buffer G (n x s) in two
// level-wide work arrays.  One level = a short sequence of launches, each with a block grid over (tile, front):
//     k_fl_assemble                         F = K[front, front] rows touching S + shift_mode diag(mass_S) + children's updates
//     for k0 = 0, 32, ... < s:              blocked right-looking partial Cholesky of the first s columns
//         k_fl_panel(k0)                        diagonal block (32 x 32, shared memory) + triangular solve of the rows below
//         k_fl_syrk(k0)                         trailing update, 64 x 64 tiles, 4 x 4 register blocks, lower triangle only
//     k_fl_trtri                            G[0:s] = inv(L11), one block per (front, 32-column block), forward substitution
//     k_fl_w21                              G[s:n] = L21 inv(L11), 64 x 64 tiles
//     k_fl_emit                             solve-ready panel [inv(L11) ; L21 inv(L11)] in both layouts (mode fastest) and the
//                                           update matrix F22 - L21 L21^T for the parent, in the layout k_front_small uses
// Plain fp64 FMA arithmetic: the tensor path (DMMA m8n8k4) has the same peak on B200 and the factorisation is a
// one-off of ~0.5 TFLOP.
#include "common.cuh"

#define FL_NB 32                 // panel width
#define FL_TILE 64               // trailing-update tile
#define FL_THREADS 256

struct fl_front {                // geometry of the front handled by a block
    int node, mode, n, s, b;
    double *F, *G;               // row-major, ld = n (F) and ld = s (G)
};

__device__ __forceinline__ fl_front fl_get(const dots_front_args_t &a, const int64_t *foff, const int64_t *goff, double *Fbuf, double *Gbuf,
                                           int front)
{
    fl_front f;
    const int il = front / a.n_modes;
    f.mode = front - il * a.n_modes;
    f.node = a.nodes[il];
    f.s = a.nd_s[f.node];
    f.b = a.nd_b[f.node];
    f.n = f.s + f.b;
    f.F = Fbuf + foff[il] + (size_t)f.mode * f.n * f.n;
    f.G = Gbuf + goff[il] + (size_t)f.mode * f.n * f.s;
    return f;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FL_THREADS) k_fl_assemble(dots_front_args_t a, const int64_t *foff, const int64_t *goff, double *Fbuf, double *Gbuf)
{
    const fl_front f = fl_get(a, foff, goff, Fbuf, Gbuf, blockIdx.x);
    const int n = f.n, s = f.s, tid = threadIdx.x, M = a.m_pad;
    const int off = a.nd_off[f.node];
    double *F = f.F;
    for (size_t i = tid; i < (size_t)n * n; i += FL_THREADS) F[i] = 0.0;
    __syncthreads();
    for (int r = 0; r < s; ++r) {                                           // matrix rows owned by the node, mirrored
        const int64_t q0 = a.a_ptr[off + r], q1 = a.a_ptr[off + r + 1];
        for (int64_t q = q0 + tid; q < q1; q += FL_THREADS) {
            const int c = a.a_pos[q];
            if (c >= 0) {
                const double v = a.a_val[q];
                F[(size_t)r * n + c] = v;
                F[(size_t)c * n + r] = v;
            }
        }
    }
    __syncthreads();
    const double shift = a.shifts[f.mode];
    const bool pinned = f.node == a.pin_node && shift == 0.0;               // singular mode: pin the last pivot (one owner thread)
    for (int r = tid; r < s; r += FL_THREADS)
        F[(size_t)r * n + r] += shift * a.mass[off + r] + ((pinned && r == s - 1) ? a.pin_value : 0.0);
    __syncthreads();
    for (int slot = 0; slot < 2; ++slot) {                                  // extend-add the children's update matrices
        const int ch = a.nd_child[2 * f.node + slot];
        if (ch < 0) continue;
        const int bc = a.nd_b[ch];
        const double *U = reinterpret_cast<const double *>(a.u_ptr[ch]);
        if (!bc || !U) continue;
        const int32_t *pp = a.parent_pos + a.nd_upd[ch];
        for (size_t i = tid; i < (size_t)bc * bc; i += FL_THREADS) {
            const int p = (int)(i / bc), q = (int)(i - (size_t)p * bc);
            F[(size_t)pp[p] * n + pp[q]] += U[i * M + f.mode];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Panel step k0: Cholesky of the diagonal block D = F[k0:k0+nb, k0:k0+nb] in shared memory, then the rows below:
// F[i, k0:k0+nb] <- F[i, k0:k0+nb] D^-T  (one thread per row, forward substitution in registers).
__global__ void __launch_bounds__(FL_THREADS) k_fl_panel(dots_front_args_t a, const int64_t *foff, const int64_t *goff, double *Fbuf, double *Gbuf, int k0)
{
    const fl_front f = fl_get(a, foff, goff, Fbuf, Gbuf, blockIdx.x);
    if (k0 >= f.s) return;
    __shared__ double D[FL_NB][FL_NB + 1];
    const int n = f.n, tid = threadIdx.x;
    const int nb = min(FL_NB, f.s - k0);
    double *F = f.F;
    for (int i = tid; i < FL_NB * FL_NB; i += FL_THREADS) {
        const int r = i / FL_NB, c = i - r * FL_NB;
        D[r][c] = (r < nb && c < nb) ? F[(size_t)(k0 + r) * n + k0 + c] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int k = 0; k < nb; ++k) {                                          // unblocked right-looking Cholesky of D (lower)
        if (tid == 0) D[k][k] = sqrt(D[k][k]);
        __syncthreads();
        if (tid > k && tid < nb) D[tid][k] /= D[k][k];
        __syncthreads();
        for (int i = tid; i < nb * nb; i += FL_THREADS) {
            const int r = i / nb, c = i - r * nb;
            if (r > k && c > k && c <= r) D[r][c] -= D[r][k] * D[c][k];
        }
        __syncthreads();
    }
    for (int i = tid; i < nb * nb; i += FL_THREADS) {                       // L_D back into the front (lower part)
        const int r = i / nb, c = i - r * nb;
        if (c <= r) F[(size_t)(k0 + r) * n + k0 + c] = D[r][c];
    }
    for (int i = k0 + nb + tid; i < n; i += FL_THREADS) {                   // x D^T = row  =>  x_c = (row_c - sum_{p<c} x_p D[c][p]) / D[c][c]
        double *row = F + (size_t)i * n + k0;
        double x[FL_NB];
#pragma unroll
        for (int c = 0; c < FL_NB; ++c) {
            double v = (c < nb) ? row[c] : 0.0;
#pragma unroll
            for (int p = 0; p < c; ++p) v -= x[p] * D[c][p];
            x[c] = v / D[c][c];
        }
#pragma unroll
        for (int c = 0; c < FL_NB; ++c)
            if (c < nb) row[c] = x[c];
    }
}

// ------------------------------------------------------------------------------------------------
// Trailing update of panel step k0, lower triangle: C[i][j] -= sum_k P[i][k] P[j][k], P = F[:, k0:k0+nb], i >= j >= k0+nb.
__global__ void __launch_bounds__(FL_THREADS) k_fl_syrk(dots_front_args_t a, const int64_t *foff, const int64_t *goff, double *Fbuf, double *Gbuf, int k0)
{
    const fl_front f = fl_get(a, foff, goff, Fbuf, Gbuf, blockIdx.z);
    if (k0 >= f.s) return;
    const int nb = min(FL_NB, f.s - k0);
    const int base = k0 + nb;
    const int i0 = base + blockIdx.y * FL_TILE, j0 = base + blockIdx.x * FL_TILE;
    if (i0 >= f.n || j0 >= f.n || j0 > i0) return;
    __shared__ double A[FL_NB][FL_TILE + 4], B[FL_NB][FL_TILE + 4];        // [k][row]: conflict-free 4-wide reads
    const int n = f.n, tid = threadIdx.x;
    double *F = f.F;
    for (int i = tid; i < FL_TILE * FL_NB; i += FL_THREADS) {
        const int r = i / FL_NB, k = i - r * FL_NB;
        A[k][r] = (i0 + r < n && k < nb) ? F[(size_t)(i0 + r) * n + k0 + k] : 0.0;
        B[k][r] = (j0 + r < n && k < nb) ? F[(size_t)(j0 + r) * n + k0 + k] : 0.0;
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    double acc[4][4] = {};
#pragma unroll 8
    for (int k = 0; k < FL_NB; ++k) {
        double av[4], bv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { av[q] = A[k][4 * ty + q]; bv[q] = B[k][4 * tx + q]; }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[p][q] += av[p] * bv[q];
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int i = i0 + 4 * ty + p;
        if (i >= n) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + 4 * tx + q;
            if (j <= i) F[(size_t)i * n + j] -= acc[p][q];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// G[0:s, c0:c0+w] = column block of inv(L11): X_jj = inv(L_jj);  X_ij = -inv(L_ii) sum_{k=j}^{i-1} L_ik X_kj  (i > j).
__global__ void __launch_bounds__(FL_THREADS) k_fl_trtri(dots_front_args_t a, const int64_t *foff, const int64_t *goff, double *Fbuf, double *Gbuf)
{
    const fl_front f = fl_get(a, foff, goff, Fbuf, Gbuf, blockIdx.y);
    const int s = f.s, n = f.n, tid = threadIdx.x;
    const int jb = blockIdx.x, c0 = jb * FL_NB;
    if (c0 >= s) return;
    const int w = min(FL_NB, s - c0);
    __shared__ double Ls[FL_NB][FL_NB + 1], Xs[FL_NB][FL_NB + 1], Ts[FL_NB][FL_NB + 1];
    const double *F = f.F;
    double *G = f.G;
    const int r = tid >> 3, cg = tid & 7;                                   // thread -> row r, columns cg, cg+8, cg+16, cg+24
    // zero the strictly upper part of this column block (rows above c0) so that the emit step can read G freely
    for (int i = tid; i < c0 * w; i += FL_THREADS) G[(size_t)(i / w) * s + c0 + i % w] = 0.0;
    const int nblk = (s + FL_NB - 1) / FL_NB;
    for (int ib = jb; ib < nblk; ++ib) {
        const int r0 = ib * FL_NB, h = min(FL_NB, s - r0);
        double t[4] = {0.0, 0.0, 0.0, 0.0};                                 // T = sum_k L_ik X_kj for (r, cg + 8q)
        for (int kb = jb; kb < ib; ++kb) {
            const int k0 = kb * FL_NB;
            __syncthreads();
            for (int i = tid; i < FL_NB * FL_NB; i += FL_THREADS) {
                const int rr = i / FL_NB, cc = i - rr * FL_NB;
                Ls[rr][cc] = (rr < h) ? F[(size_t)(r0 + rr) * n + k0 + cc] : 0.0;          // L_ik (k block is full: kb < ib)
                Xs[rr][cc] = (cc < w) ? G[(size_t)(k0 + rr) * s + c0 + cc] : 0.0;          // X_kj
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < FL_NB; ++k) {
                const double l = Ls[r][k];
#pragma unroll
                for (int q = 0; q < 4; ++q) t[q] += l * Xs[k][cg + 8 * q];
            }
        }
        __syncthreads();
        for (int i = tid; i < FL_NB * FL_NB; i += FL_THREADS) {             // L_ii (identity padding) and the right-hand side
            const int rr = i / FL_NB, cc = i - rr * FL_NB;
            Ls[rr][cc] = (rr < h && cc < h) ? F[(size_t)(r0 + rr) * n + r0 + cc] : (rr == cc ? 1.0 : 0.0);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) Ts[r][cg + 8 * q] = (ib == jb) ? ((r == cg + 8 * q) ? 1.0 : 0.0) : -t[q];
        __syncthreads();
        if (tid < FL_NB) {                                                   // column tid of X_ij: forward substitution with L_ii
            const int c = tid;
            double x[FL_NB];
#pragma unroll
            for (int rr = 0; rr < FL_NB; ++rr) {
                double v = Ts[rr][c];
#pragma unroll
                for (int p = 0; p < rr; ++p) v -= Ls[rr][p] * x[p];
                x[rr] = v / Ls[rr][rr];
            }
#pragma unroll
            for (int rr = 0; rr < FL_NB; ++rr) Xs[rr][c] = x[rr];
        }
        __syncthreads();
        for (int i = tid; i < h * w; i += FL_THREADS) {
            const int rr = i / w, cc = i - rr * w;
            G[(size_t)(r0 + rr) * s + c0 + cc] = Xs[rr][cc];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// G[s + i][j] = sum_{k >= j} L21[i][k] X[k][j]   (X = inv(L11) lower triangular, rows 0..s-1 of G), 64 x 64 tiles.
__global__ void __launch_bounds__(FL_THREADS) k_fl_w21(dots_front_args_t a, const int64_t *foff, const int64_t *goff, double *Fbuf, double *Gbuf)
{
    const fl_front f = fl_get(a, foff, goff, Fbuf, Gbuf, blockIdx.z);
    const int s = f.s, n = f.n, b = f.b, tid = threadIdx.x;
    const int i0 = blockIdx.y * FL_TILE, j0 = blockIdx.x * FL_TILE;
    if (i0 >= b || j0 >= s) return;
    __shared__ double A[FL_NB][FL_TILE + 4], B[FL_NB][FL_TILE + 4];        // A[k][row of L21], B[k][column of X]
    const double *F = f.F;
    double *G = f.G;
    const int ty = tid >> 4, tx = tid & 15;
    double acc[4][4] = {};
    for (int k0 = (j0 / FL_NB) * FL_NB; k0 < s; k0 += FL_NB) {
        __syncthreads();
        for (int i = tid; i < FL_TILE * FL_NB; i += FL_THREADS) {
            const int rr = i / FL_NB, k = i - rr * FL_NB;
            A[k][rr] = (i0 + rr < b && k0 + k < s) ? F[(size_t)(s + i0 + rr) * n + k0 + k] : 0.0;
        }
        for (int i = tid; i < FL_NB * FL_TILE; i += FL_THREADS) {
            const int k = i / FL_TILE, cc = i - k * FL_TILE;
            B[k][cc] = (k0 + k < s && j0 + cc < s && j0 + cc <= k0 + k) ? G[(size_t)(k0 + k) * s + j0 + cc] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < FL_NB; ++k) {
            double av[4], bv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { av[q] = A[k][4 * ty + q]; bv[q] = B[k][4 * tx + q]; }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] += av[p] * bv[q];
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int i = i0 + 4 * ty + p;
        if (i >= b) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + 4 * tx + q;
            if (j < s) G[(size_t)(s + i) * s + j] = acc[p][q];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Solve-ready panel in both layouts (mode fastest: 64 consecutive threads write 64 consecutive modes of one entry) and the
// update matrix for the parent.  blockIdx.y = launched node, blockIdx.x = chunk of work items.
__global__ void __launch_bounds__(FL_THREADS) k_fl_emit(dots_front_args_t a, const int64_t *foff, const int64_t *goff, double *Fbuf, double *Gbuf)
{
    const int il = blockIdx.y, node = a.nodes[il];
    const int s = a.nd_s[node], b = a.nd_b[node], n = s + b, M = a.m_pad, nm = a.n_modes;
    const size_t pbase = (size_t)a.nd_panel[node];
    const size_t ntri = (size_t)s * (s + 1) / 2;
    const size_t n_panel = (size_t)n * s, n_upd = (size_t)b * b;
    const double *Fn = Fbuf + foff[il], *Gn = Gbuf + goff[il];
    double *U = reinterpret_cast<double *>(a.u_ptr[node]);
    const int mode = threadIdx.x % M, sub = threadIdx.x / M, per = FL_THREADS / M;       // M in {8 .. 128}: per >= 2
    if (sub >= per) return;                                                               // M = 96: 64 idle threads
    for (size_t e = (size_t)blockIdx.x * per + sub; e < n_panel + n_upd; e += (size_t)gridDim.x * per) {
        if (e < n_panel) {
            const int r = (int)(e / s), c = (int)(e - (size_t)r * s);
            if (r < s && c > r) continue;
            const size_t fo = (r < s) ? (size_t)r * (r + 1) / 2 + c : ntri + (size_t)(r - s) * s + c;
            const size_t to = (size_t)c * n - (size_t)c * (c - 1) / 2 + (r - c);
            const double v = (mode < nm) ? Gn[(size_t)mode * n * s + e] : ((r == c) ? 1.0 : 0.0);   // padding modes: identity
            a.panels[(pbase + fo) * M + mode] = v;
            a.panels_t[(pbase + to) * M + mode] = v;
        } else if (U && mode < nm) {
            const size_t i = e - n_panel;
            const int p = (int)(i / b), q = (int)(i - (size_t)p * b);
            const double *F = Fn + (size_t)mode * n * n;
            U[i * M + mode] = (q <= p) ? F[(size_t)(s + p) * n + s + q] : F[(size_t)(s + q) * n + s + p];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Factorise the `n_launch` large fronts a->nodes of one tree level for all modes.  foff / goff (device, [n_launch]): offsets
// of every node's block of fronts inside the work arrays Fwork (sum n^2 * n_modes doubles) / Gwork (sum n * s * n_modes);
// n_max / s_max / b_max: largest front of the launch (grid sizing).
extern "C" int dots_factor_large_fronts(const dots_front_args_t *a, int n_launch, int n_max, int s_max, int b_max,
                                        const int64_t *foff, const int64_t *goff, double *Fwork, double *Gwork, void *stream)
{
    if (!a || n_launch <= 0) return 0;
    if (!foff || !goff || !Fwork || !Gwork || a->m_pad > FL_THREADS / 2 || a->m_pad < 8) { dots_set_error("dots_factor_large_fronts: bad arguments"); return DOTS_ERR_BAD_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const int fronts = n_launch * a->n_modes;
    if (fronts > 65535) { dots_set_error("dots_factor_large_fronts: %d fronts exceed the grid (split the launch)", fronts); return DOTS_ERR_BAD_ARG; }
    k_fl_assemble<<<fronts, FL_THREADS, 0, st>>>(*a, foff, goff, Fwork, Gwork);
    DOTS_LAUNCH_CHECK();
    for (int k0 = 0; k0 < s_max; k0 += FL_NB) {
        k_fl_panel<<<fronts, FL_THREADS, 0, st>>>(*a, foff, goff, Fwork, Gwork, k0);
        DOTS_LAUNCH_CHECK();
        const int rem = n_max - k0 - 1;                                    // the smallest trailing block any front can have starts after k0
        if (rem > 0) {
            const int tiles = ceil_div(n_max - k0, FL_TILE);
            k_fl_syrk<<<dim3(tiles, tiles, fronts), FL_THREADS, 0, st>>>(*a, foff, goff, Fwork, Gwork, k0);
            DOTS_LAUNCH_CHECK();
        }
    }
    k_fl_trtri<<<dim3(ceil_div(s_max, FL_NB), fronts), FL_THREADS, 0, st>>>(*a, foff, goff, Fwork, Gwork);
    DOTS_LAUNCH_CHECK();
    if (b_max > 0) {
        k_fl_w21<<<dim3(ceil_div(s_max, FL_TILE), ceil_div(b_max, FL_TILE), fronts), FL_THREADS, 0, st>>>(*a, foff, goff, Fwork, Gwork);
        DOTS_LAUNCH_CHECK();
    }
    k_fl_emit<<<dim3(1024, n_launch), FL_THREADS, 0, st>>>(*a, foff, goff, Fwork, Gwork);
    DOTS_LAUNCH_CHECK();
    return 0;
}
