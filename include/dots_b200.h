/*
 * dots_b200.h - C-ABI of the B200-native DOTs-SOCP hot path (libdots_b200.so).
 *
 * The reference (chlhnu/DOTs-SOCP) is pure Python and has no FFI of its own; the boundary it offers
 * is the solver plug-in  solver(n_time, geometry, **kwargs) -> (solution, run_history)
 * (dot_surface_socp/interface.py:106-122,295-299).  This header is what a host language binds to run
 * the inner ALM iteration of dot_surface_socp/socp/solver_socp.py:656-823 on one B200; the Python
 * mirror of the reference interface that sits on top of it is dots_socp_b200/solver.py, and the
 * ctypes binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer to C-contiguous fp64 / int32 / int64 storage owned by the
 *     caller (PyTorch tensors in the shipped host code); nothing is allocated or freed here except
 *     CUDA graphs cached inside a dots_graph_t;
 *   - every call takes the CUDA stream to enqueue on (a cudaStream_t passed as void*) and returns
 *     immediately (no hidden synchronisation) unless its comment says "synchronises";
 *   - return value: 0 on success, otherwise a negative dots error or a positive cudaError_t;
 *     dots_last_error() gives the message of the last failure on the calling thread;
 *   - INTERNAL LAYOUT.  Vertices are numbered in the nested-dissection elimination order, triangles
 *     are renumbered for locality, and per-triangle data is structure-of-arrays with the triangle
 *     index fastest:
 *         vertex fields   phi[t][v]                 t = 0..nT      (reference: (nT+1, V))
 *                         A, lam_c, mu, z_fst, z_end, b_fst, b_end, lam  [t][v], t = 0..nT-1
 *         triangle fields B, E  [tau][xyz][f]        tau = 0..nT    (reference: (nT+1, T, 3))
 *         corner fields   b_mid, z_mid [tau][side][k][xyz][f]       (reference: (nT, 2, 3, T, 3) indexed
 *                         [t][s][k][f][xyz] with tau = t + s, side = s; the slots (tau=0, side=1) and
 *                         (tau=nT, side=0) do not exist in the reference and stay zero)
 *     The host side (dots_socp_b200/engine.py: to_internal / from_internal) converts between the reference
 *     layout and this one.
 */
#ifndef DOTS_B200_H
#define DOTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOTS_ABI_VERSION 19

/* scalar block read by the kernels from device memory (so CUDA graphs stay valid across penalty updates) */
enum {
    DOTS_P_R = 0,        /* penalty r                      (solver_socp.py:97, :367-371)              */
    DOTS_P_S,            /* scale_factor_z                 (:321, :373-395)                           */
    DOTS_P_D,            /* constant_d                     (:320)                                     */
    DOTS_P_CONG,         /* congestion                     (:28)                                      */
    DOTS_P_TAU,          /* multiplier step tau            (:32)                                      */
    DOTS_P_EPS,          /* Laplacian regularisation eps   (:30)                                      */
    DOTS_P_PS,           /* prim_scale (1 unless is_constant_scaling; :318, :324-365): read by KKT #4-#6 / objective */
    DOTS_P_DS,           /* dual_scale (:319)                                                         */
    DOTS_P_COUNT = 8
};

/* Work item of the ring-streamed sweeps (sweep_mode 4, csrc/sweep_ring.cu).  "Output" = a panel row in the forward
 * sweep, a panel column (of the column-major copy) in the backward sweep.
 *   contiguous task (one warp): the outputs [oa, oa + n_out) of one node, whose n_ent panel entries form ONE contiguous
 *                               run starting at entry pbase of panels / panels_t;
 *   split item (one block)    : the outputs [oa, oa + n_out) of one node whose runs are long enough to be shared by several
 *                               warps; pbase = the node's panel base, n_ent unused.                                        */
typedef struct dots_ring_task {
    int64_t pbase;
    int32_t n_ent;
    int32_t oa;
    int32_t n_out;
    int32_t s, b;              /* |S|, |B| of the node                                                 */
    int32_t off;               /* first vertex owned by the node                                       */
    int32_t ubase;             /* first row of the node's contribution block in `upd`                  */
    int32_t fbase;             /* first row of the node's front in `bidx`                              */
    int32_t pad[2];
} dots_ring_task_t;

/* Problem description + device storage.  Filled by the host side once; passed to every call. */
typedef struct dots_ctx {
    int32_t abi_version;
    int32_t n_time;            /* nT: number of time intervals                                        */
    int32_t n_vert;            /* V                                                                   */
    int32_t n_tri;             /* T                                                                   */
    int32_t m_pad;             /* time modes solved by THIS rank, padded to 8, 16, 32, 64, 96 or 128   */
    int32_t n_nodes;           /* separator-tree nodes                                                */
    int32_t n_levels;          /* separator-tree levels                                               */
    int32_t n_sm;              /* multiprocessor count (grid sizing)                                  */

    /* ---- mesh constants (read-only) ---- */
    const int32_t *tri;        /* [3][T]    vertex of corner k of triangle f                          */
    const double  *hat_grad;   /* [3][3][T] hat-function gradients g[k][xyz][f]  (surface_pre_computations_socp.py:30-37) */
    const double  *area_f;     /* [T]                                                                 */
    const double  *area_v;     /* [V]       (sum of incident |f|)/3               (solver_socp.py:112) */
    const double  *diag_soc;   /* [3][T]    sqrt(|f| / area_v[tri[k][f]])         (:172-192)           */
    const int32_t *vc_ptr;     /* [V+1]     CSR vertex -> incident corners                            */
    const int32_t *vc_idx;     /* [3T]      corner ids k*T+f, ascending per vertex                    */
    const int32_t *vc_ell;     /* [V][8]    the same lists in ELL form: first 8 corner ids, -1 padded; [v][7] == -2: more than 8
                                  corners, walk the CSR list (32-byte rows, 16-byte aligned)                                */
    const double  *qf;         /* [tt_kf][m_pad]   forward transform matrix: rows = time levels (all), columns = this rank's modes
                                  (Q[t][mode] of laplacian_inverse_socp.py:31, zero padded)             */
    const double  *qb;         /* [tt_kb][tt_nb]   inverse transform matrix: rows = gathered modes (rank-major, padded),
                                  columns = the phi levels this rank needs (lvl_begin .. lvl_begin+tt_nout-1) */

    /* ---- batched multifrontal factor of K + shift_mode*diag(area_v) (dots_socp_b200/nested.py) ---- */
    const double  *panels;     /* [panel_entries][m_pad]  per node row-major: rows of [inv(L11) ; L21 inv(L11)]  */
    const double  *panels_t;   /* same entries, per node column-major (column j holds rows j..s+b-1): backward sweep */
    const int32_t *nd_off;     /* [n_nodes] first vertex owned                                        */
    const int32_t *nd_s;       /* [n_nodes] |S|                                                       */
    const int32_t *nd_b;       /* [n_nodes] |B|                                                       */
    const int32_t *nd_child;   /* [n_nodes][2] child node ids or -1                                   */
    const int64_t *nd_panel;   /* [n_nodes] panel offset (entries)                                    */
    const int64_t *nd_front;   /* [n_nodes] offset into front_idx / child_pos                         */
    const int64_t *nd_upd;     /* [n_nodes] offset of the update vector (rows)                        */
    const int32_t *front_idx;  /* [sum(s+b)] vertex of every front row                                */
    const int32_t *child_pos;  /* [2][sum(s+b)] row in the child's update vector or -1                */
    const int32_t *lvl_ptr;    /* [n_levels+1] ranges into lvl_items                                  */
    const int32_t *lvl_items;  /* work items (node, first row, n rows) as int32 triples, forward      */
    const int32_t *lvb_ptr;    /* [n_levels+1] ranges into lvb_items                                  */
    const int32_t *lvb_items;  /* work items (node, first col, n cols) as int32 triples, backward     */
    const int32_t *lvn_nodes;  /* gather work items (node, first S row, n rows) grouped by level      */
    const int32_t *h_lvl_ptr;  /* HOST copies of lvl_ptr / lvb_ptr (grid sizing of the per-level launches) */
    const int32_t *h_lvb_ptr;
    const int32_t *h_lvn_ptr;  /* HOST [n_levels+1] ranges into lvn_nodes                             */
    const int32_t *h_lvl_wpr;  /* HOST [n_levels] warps sharing one panel row in the forward sweep (1,2,4,8); +16: the level's
                                  blocks fold the children's updates into r_S themselves (no separate gather launch) */
    const int32_t *h_lvb_cw;   /* HOST [n_levels] warps sharing one panel column in the backward sweep (1,2,4,8) */
    int64_t front_total;       /* sum(s+b)                                                            */

    /* ---- sharding (one process per GPU; all 0 / full range on a single GPU) ----
     * Time-slab: this rank owns the time levels [lvl_begin, lvl_end) of every level-indexed array and the staggered
     * steps [lvl_begin, min(lvl_end, nT)).  State pointers below are VIRTUAL bases: element (level, ...) of an array
     * is at base + level*stride exactly as on one GPU, but only the owned levels (+ the halo levels named in
     * DESIGN.md section 5) are backed by memory.  Time-mode: the sweeps run on this rank's m_pad modes only.       */
    int32_t lvl_begin, lvl_end;
    int32_t n_ranks;
    int32_t tt_kf;             /* rows of qf (nT+1 rounded up to 4)                                   */
    int32_t tt_kb;             /* rows of qb (n_ranks * m_pad)                                        */
    int32_t tt_nb;             /* columns of qb (tt_nout rounded up to 8)                             */
    int32_t tt_nout;           /* phi levels written by the inverse transform                         */
    int32_t tt_sym;            /* 1: one rank, n_time + 1 = m_pad a multiple of 16, modes stored even-first (k = 0, 2, ... | 1, 3, ...):
                                  the transforms use the even / odd symmetry of the DCT-II basis (half the tensor work)      */

    /* ---- ALM state (read-write) ---- */
    double *params;            /* [DOTS_P_COUNT]                                                      */
    double *phi, *A, *lam_c, *mu, *z_fst, *z_end, *b_fst, *b_end, *lam;
    double *bnd0, *bnd1;       /* [V] rows t=0 and t=nT of boundary_time_with_area (:267-270)         */
    double *B, *E;
    double *b_mid, *z_mid;
    double *corner_nrm;        /* [nT+1][2][3][T] per-corner squared norms feeding the next projection */
    double *corner_div;        /* [nT+1][3][T]    per-corner divergence terms feeding the next rhs     */

    /* ---- work space ---- */
    double *rhs;               /* [nT+1 (padded to n_ranks equal slabs)][V], NOT virtual: every rank holds all levels */
    double *hat;               /* [V][m_pad]  transformed rhs, then solution (this rank's modes)      */
    double *hat_all;           /* [n_ranks][V][m_pad] all ranks' solutions (== hat on one GPU)        */
    double *ywork;             /* [V][m_pad]  forward-sweep result                                    */
    double *upd;               /* [sum b][m_pad] update vectors                                       */
    double *red_part;          /* [red_blocks][72] block partial sums (9 conditions x 8 slots)        */
    double *red_out;           /* [72] reduced sums (device)                                          */
    int32_t red_blocks;
    int32_t sweep_mode;        /* 0: k_sweep_run, register-staged loads (any m_pad); 4: ring-streamed sweeps, every warp feeds its
                                  own shared-memory ring with bulk async copies (m_pad a multiple of 32), one launch per level   */
    int32_t sweep_grid;        /* unused (kept for layout stability)                                   */
    int32_t reserved0;
    uint64_t *phase_clock;     /* optional [2*n_levels+1]: %globaltimer (ns) at the start and after each sweep phase */

    /* ---- peer memory (multi-GPU, optional; all NULL = exchanges go through the host-side collectives) ----
     * Pointers into OTHER ranks' device memory (CUDA IPC mappings over NVLink).  When set, the producing kernels store
     * the neighbour halos / the rhs slab straight into the consumers' buffers, and the host only issues a tiny
     * stream-ordered all-reduce as the cross-rank fence (dist.py).                                                */
    double *peer_vertex[4];    /* next rank's halo rows (step lvl_end-1 there = its lvl_begin-1) of lam, A, lam_c, mu  */
    double *peer_corner;       /* previous rank's corner_nrm halo level (its lvl_end = my lvl_begin), side 1: [3][T]   */
    double *peer_rhs[8];       /* every rank's rhs buffer (own included) when n_ranks <= 8                             */
    const double *peer_hat[8]; /* every rank's solution buffer `hat` (own included): the inverse transform reads the other
                                  ranks' modes straight from their memory instead of from a gathered copy             */

    /* ---- ring-streamed sweeps (sweep_mode 4).  Needs ywork == hat + n_vert*m_pad (one allocation Z = [hat | ywork]). ---- */
    const dots_ring_task_t *rt_fwd, *rt_bwd;    /* device task arrays, grouped by tree level                              */
    const int32_t *h_rt_fwd_ptr, *h_rt_bwd_ptr; /* HOST [n_levels+1] ranges into rt_fwd / rt_bwd                           */
    const int32_t *h_rt_fwd_wpr, *h_rt_bwd_wpr; /* HOST [n_levels] 1: contiguous warp tasks; 2, 4, 8: split items, that many
                                                   warps share one output                                                  */
    const int32_t *bidx;       /* [front_total] row of Z read by the backward sweep for every front row: n_vert + vertex for
                                  the S rows (y in ywork), vertex for the B rows (x of the ancestors in hat)               */
    const int32_t *erow_fwd;   /* [panel_entries] per panel entry (row-major order): row of Z it multiplies | last-of-output << 31 */
    const int32_t *erow_bwd;   /* [panel_entries] the same for the column-major copy                                     */
    const int32_t *gptr;       /* [V+1] ranges into gidx: the contributions landing on vertex v                           */
    const int32_t *gidx;       /* [sum b] rows of `upd` (producer order: nd_upd[node] + boundary row), grouped by the vertex
                                  they land on, producers in post-order (fixed summation order)                           */
    const int32_t *gverts;     /* vertices that receive contributions, grouped by the tree level of their owner node      */
    const int32_t *h_gv_ptr;   /* HOST [n_levels+1] ranges into gverts                                                    */
    int32_t ring_stages;       /* shared-memory stages per warp (2..6)                                                    */
    int32_t ring_pdl;          /* 1: chain the launches of an iteration with programmatic dependent launch                */
    int32_t ring_stage_bytes;  /* bytes per ring stage: 2048 or 4096                                                      */
    int32_t ring_flags;        /* bit 0: the panel copies of the ring-streamed sweeps carry an L2 evict_first hint;
                                  bit 2: dots_step_tri uses the plain-load kernel instead of the TMA-staged one (diagnostics).
                                  The TMA kernel may read up to 8 bytes past the end of B, E and b_mid: pad them by 16.     */
    double *kkt1_part;         /* [kkt1_blocks] per-block partial sums of the triangle term of KKT #1, written by
                                  dots_step_tri(write_z = 2) (one per block of its grid, fixed order)                     */
    int32_t kkt1_blocks;       /* capacity of kkt1_part: >= ceil(n_tri / 128) * ceil(owned levels / 2)                     */
    int32_t reserved3;

} dots_ctx_t;

/* ------------------------------------------------------------------------------------------------ */
int  dots_enable_peer(int peer_device);         /* cudaDeviceEnablePeerAccess(current -> peer), idempotent */
/* CUDA IPC hand-over of a device buffer between the ranks of one node (64-byte handle + offset inside the allocation) */
int  dots_ipc_export(const void *dev_ptr, void *handle_out_64, unsigned long long *offset_out);
int  dots_ipc_import(const void *handle_64, unsigned long long offset, void **dev_ptr_out);
int  dots_abi_version(void);
int  dots_ctx_sizeof(void);                      /* sizeof(dots_ctx_t): the binding checks its mirror  */
const char *dots_last_error(void);

/* ---- one ALM iteration, fused (rows a1-a8 of SURVEY.md section 8) ---------------------------------
 * dots_step_phi   : rhs assembly + time transform + per-mode solves + inverse transform -> phi
 *                   (vanilla_solve_laplacian solver_socp.py:976-986, __laplacian_invert laplacian_inverse_socp.py:52-61)
 * dots_step_vertex: per (t,v): dt_phi, cone multiplier lam, z_fst/z_end, A, lam_c, mu, b_fst, b_end
 *                   (vertex halves of vanilla_solve_proj_soc :988-1042, vanilla_solve_q_lambda :1044-1065, Step 3 :716-722)
 * dots_step_tri   : per (tau,f): dx_phi, z_mid, B, E, b_mid and the corner terms of the next iteration
 *                   (triangle halves of the same three steps + decouple_spacial :923-942 / adjoint :944-959)
 *                   write_z = 1 also stores z_mid (needed by KKT #1, is_palm, the variable norms and the returned solution);
 *                   write_z = 2 stores no z_mid but accumulates the triangle term of KKT #1, sum of area_f (s (z_mid -
 *                   s/sqrt3 B))^2, per block into kkt1_part (TMA kernel only): what a check iteration needs when z_mid
 *                   itself is not going to be returned.
 * dots_iterate    : n_iter x (phi, vertex, tri); write_z applies to the last one.                    */
int dots_step_phi(const dots_ctx_t *c, void *stream);
int dots_step_vertex(const dots_ctx_t *c, void *stream);
int dots_step_tri(const dots_ctx_t *c, int write_z, void *stream);
int dots_iterate(const dots_ctx_t *c, int n_iter, int write_z, void *stream);
/* is_palm=True only (solver_socp.py:253-257, :668-672): the extra q / lambda solve that opens an iteration, from the gradients
 * of the current phi and the STORED z (z_mid of the previous iteration must have been written): A, lam_c (vertex kernel), B and
 * the corner terms of the new B (triangle kernel).  Single GPU.                                                           */
int dots_step_q0(const dots_ctx_t *c, void *stream);

/* One iteration captured as a CUDA graph (same work as dots_iterate(c, 1, write_z)); the context must outlive it and
 * must not be modified afterwards (scalars go through dots_set_params, which the graph picks up).  Call
 * dots_iterate once before creating a graph (first-use kernel attributes are set outside of stream capture). */
typedef struct dots_graph dots_graph_t;
int dots_graph_create(const dots_ctx_t *c, int write_z, void *stream, dots_graph_t **out);
int dots_graph_launch(dots_graph_t *g, void *stream);
int dots_graph_destroy(dots_graph_t *g);

/* recompute corner_nrm / corner_div from (B, E, b_mid) after the state was set or rescaled */
int dots_refresh_corner_terms(const dots_ctx_t *c, void *stream);

/* ---- rescaling (row a12) --------------------------------------------------------------------------
 * dots_scale_dual: adjust_penalty :367-371  (mu, E, boundary, b_* divided by factor) + refresh.
 * dots_scale_z   : scale_variable_z :373-395 with cumulative factor s_cum (z_* *= s_cum, b_* /= s_cum,
 *                  mu = s_cum (b_fst - b_end), E = -adjoint(b_mid; s_cum)) + refresh.
 * The caller updates params[] itself (dots_set_params).                                              */
int dots_scale_dual(const dots_ctx_t *c, double factor, void *stream);
int dots_scale_z(const dots_ctx_t *c, double s_cum, void *stream);
/* dots_scale_prim_dual: scale_prim_dual :324-365 (is_constant_scaling): phi, A, B, lam_c, z_* divided by prim_div and
 *                  boundary, mu, E, b_* divided by dual_div (= dual_rescale^2 / prim_rescale), incl. the halo rows, + refresh.
 *                  The caller updates r, congestion, constant_d, the two scales in params[] itself.                       */
int dots_scale_prim_dual(const dots_ctx_t *c, double prim_div, double dual_div, void *stream);
int dots_set_params(const dots_ctx_t *c, const double *host_params, void *stream);

/* ---- residuals (rows a9-a11).  Writes the raw weighted sums (un-normalised, un-rooted) of KKT condition
 * `which` (0..6, order of solver_socp.py:591-639) or of the objective (which = 7) into host_out[8].
 * Synchronises the stream.  Slot meaning per condition: csrc/kkt_kernels.cu (k_kkt_vertex / k_kkt_tri) and
 * dots_socp_b200/engine.py (Engine.kkt, which forms the relative residuals of solver_socp.py:433-559 from them). */
int dots_kkt_sums(const dots_ctx_t *c, int which, double *host_out, void *stream);
/* All conditions of `mask` (bit i = condition i, bit 7 = objective) in ONE pass over the vertex arrays and ONE over the
 * triangle arrays (the --detail_runhist mode of solver_socp.py:769-787 evaluates all 7 + the objective every iteration; the
 * penalty-update iterations force conditions 0-3, :728-729): host_out[8 * i + k] = slot k of condition i as in
 * dots_kkt_sums.  Bit 8: the variable norms of scale_prim_dual (:331-340): vertex slots z_fst^2, z_end^2, b_fst^2,
 * b_end^2 (x area_v), triangle slots z_mid^2, b_mid^2 (x area_f).  Bit 9: the triangle term of condition 1 is taken from
 * kkt1_part (valid after dots_step_tri / dots_iterate with write_z = 2) instead of from z_mid.  red_part must hold
 * red_blocks x 72 doubles, red_out and host_out 72.  Synchronises the stream.                                             */
int dots_kkt_sums_multi(const dots_ctx_t *c, unsigned mask, double *host_out, void *stream);

/* ---- setup (row f1): numeric factorisation of the small fronts (n <= dots_front_nmax()) of one tree level, one block
 * per (node, time mode); replaces the per-mode SuperLU factorisations of utils/laplacian_inverse_socp.py:34-41 for the
 * ~93 % of the separator-tree nodes near the leaves.  Large fronts: batched dense library calls (nested.py).        */
typedef struct dots_front_args {
    const int32_t *nodes;          /* [n_launch] node ids of this launch                                    */
    int32_t n_modes, m_pad;
    const int32_t *nd_off, *nd_s, *nd_b, *nd_child;
    const int64_t *nd_panel, *nd_front, *nd_upd;
    const int32_t *parent_pos;     /* [sum b] front row in the PARENT of every boundary row of a node        */
    const int64_t *u_ptr;          /* [n_nodes] device address of the node's update matrix [b][b][m_pad]     */
    const int64_t *a_ptr;          /* [V+1] CSR row pointers of K in the elimination ordering               */
    const int32_t *a_pos;          /* [nnz] front position (in the row owner's front) of every entry or -1  */
    const double  *a_val;          /* [nnz]                                                                  */
    const double  *mass;           /* [V]                                                                    */
    const double  *shifts;         /* [n_modes]                                                              */
    double *panels, *panels_t;
    int32_t pin_node, reserved;
    double pin_value;
} dots_front_args_t;
int dots_factor_small_fronts(const dots_front_args_t *a, int n_launch, int max_front, void *stream);
int dots_front_nmax(void);
/* The fronts with more rows than dots_front_nmax() (near the root: up to 1 277 rows at V = 164k), batched over the `n_launch`
 * nodes a->nodes of one tree level and all modes: blocked right-looking partial Cholesky, triangular inverse,
 * L21 inv(L11), panels in both layouts and the update matrices for the parents (csrc/front_large.cu).  foff / goff
 * (device, [n_launch]): offsets, in doubles, of every node's block of n_modes fronts (n x n each) / panel buffers (n x s each)
 * inside the work arrays Fwork / Gwork; n_max, s_max, b_max: largest sizes in the launch.  At most 65535 / n_modes nodes
 * per call.                                                                                                             */
int dots_factor_large_fronts(const dots_front_args_t *a, int n_launch, int n_max, int s_max, int b_max,
                             const int64_t *foff, const int64_t *goff, double *Fwork, double *Gwork, void *stream);

/* ---- setup (row f1), host side: ONE nested-dissection ordering + symbolic multifrontal analysis shared by all time
 * modes (the reference lets SuperLU order and analyse each of its nT+1 matrices, utils/laplacian_inverse_socp.py:34-41).
 * HOST pointers, no CUDA.  vertices [V][3]; adj_ptr [V+1] / adj_idx: CSR pattern of the Laplacian WITHOUT the diagonal.
 * dots_order_export fills caller-allocated int64 arrays: perm [V] (new -> old vertex), s / b / level / parent [n_nodes]
 * (|S|, |B|, tree level with leaves = 0, parent node or -1; nodes in post-order), child [n_nodes][2] (-1 = none),
 * front_idx [front_total] (rows of every front: S rows then B rows, new vertex ids ascending) and
 * child_pos [2][front_total] (row of child slot's update vector that feeds this front row, or -1).
 * Semantics and tie breaks are those of dots_socp_b200/nested.py (dissect, symbolic): identical output.            */
typedef struct dots_order dots_order_t;
int dots_order_create(int64_t n_vert, const double *vertices, const int64_t *adj_ptr, const int64_t *adj_idx,
                      int64_t leaf_size, dots_order_t **out);
int dots_order_sizes(const dots_order_t *o, int64_t *n_nodes, int64_t *front_total);
int dots_order_export(const dots_order_t *o, int64_t *perm, int64_t *s, int64_t *b, int64_t *level, int64_t *parent,
                      int64_t *child, int64_t *front_idx, int64_t *child_pos);
int dots_order_destroy(dots_order_t *o);

/* ---- setup (row f1), host side: the mesh operators of the hot path (reference utils/surface_pre_computations_socp.py:11-132,
 * Python loops over the triangles, ~25 s each at T = 328k).  HOST pointers.  vertices [V][3] f64, triangles [T][3] int64.
 * dots_mesh_export fills caller arrays (NULL = skip): area_f [T]; hat [T][3][3] P1 hat-function gradients g[f][k][xyz] (:30-37);
 * area_sum [V] UN-divided sum of incident |f| (:121-124); the cotan stiffness matrix K = -L (:68-84) as CSR with ascending
 * columns (k_ptr [V+1], k_idx / k_val [nnz], nnz from dots_mesh_sizes); c_ptr [V+1], c_tri / c_corner [3T]: corners around every
 * vertex ordered by (corner, triangle), the rows of the reference's vertex <- corner incidence map (:112-127).            */
typedef struct dots_mesh dots_mesh_t;
int dots_mesh_create(int64_t n_vert, int64_t n_tri, const double *vertices, const int64_t *triangles, dots_mesh_t **out);
int dots_mesh_sizes(const dots_mesh_t *m, int64_t *nnz);
int dots_mesh_export(const dots_mesh_t *m, double *area_f, double *hat, double *area_sum, int64_t *k_ptr, int64_t *k_idx,
                     double *k_val, int64_t *c_ptr, int64_t *c_tri, int64_t *c_corner);
int dots_mesh_destroy(dots_mesh_t *m);
/* CSR vertex -> incident corner ids k * T + f, ordered by (corner, triangle), for any triangle numbering: c_ptr [V+1], c_idx [3T]
 * (int32; dots_ctx_t.vc_ptr / vc_idx).                                                                                  */
int dots_corner_lists(int64_t n_vert, int64_t n_tri, const int64_t *triangles, int32_t *c_ptr, int32_t *c_idx);
/* Index maps of the numeric assembly: a_pos [nnz] = front position of every CSR entry of the PERMUTED matrix (a_ptr / a_idx, sorted
 * columns) inside the front of the node that owns its row (-1: already eliminated), parent_pos [upd_off[n_nodes]] = row of every
 * boundary row inside the parent's front (-1 at the root).  Arrays as in dots_order_export.  Replaces the symbolic part of
 * pre_factorize_* (utils/laplacian_inverse_socp.py:11-49).                                                                */
int dots_front_maps(int64_t n_vert, int64_t n_nodes, const int64_t *s, const int64_t *b, const int64_t *off, const int64_t *front_off,
                    const int64_t *front_idx, const int64_t *upd_off, const int64_t *parent, const int64_t *a_ptr,
                    const int64_t *a_idx, int32_t *a_pos, int32_t *parent_pos);

/* out = A[perm][:, perm] of a CSR matrix, ascending columns in every row (perm: new -> old vertex, iperm: old -> new; o_ptr [n+1],
 * o_idx / o_val [nnz] caller-allocated): the permuted stiffness matrix the numeric assembly reads.  HOST pointers.       */
int dots_csr_permute(int64_t n, const int64_t *a_ptr, const int64_t *a_idx, const double *a_val, const int64_t *perm,
                     const int64_t *iperm, int64_t *o_ptr, int64_t *o_idx, double *o_val);

/* ---- setup, host side: per-entry operand rows of the ring-streamed sweeps (sweep_mode 4).  HOST pointers.  For every
 * panel entry in streaming order (rows_fwd: row-major panels, rows_bwd: column-major copy) the row of Z = [hat | ywork]
 * the entry is multiplied with; bit 31 = last entry of its output.  s, b, off, front_off [n_nodes(+1)], panel_off
 * [n_nodes+1] as in dots_order_export / nested.Symbolic; bidx [front_total] as in dots_ctx_t.                            */
int dots_ring_entry_rows(int64_t n_nodes, const int64_t *s, const int64_t *b, const int64_t *off, const int64_t *front_off,
                         const int64_t *panel_off, const int32_t *bidx, int32_t *rows_fwd, int32_t *rows_bwd);

/* ---- operator-level entry points on the internal layout (rows a5, a6, a7, a8) ---------------------- */
int dots_phi_rhs(const dots_ctx_t *c, void *stream);                       /* -> c->rhs                */
int dots_time_transform(const dots_ctx_t *c, int inverse, void *stream);
/* One mode group of a time grid with n_time + 1 > 128 levels (single GPU; the engine loops transform -> dots_mode_solves ->
 * inverse transform over groups of ctx->m_pad modes, each with its own factor panels).  inverse = 0: ctx->hat[v][j] =
 * sum_t q[t][j] rhs[t][v], q [n_time+1][m_pad]; inverse = 1: phi[t][v] (+)= sum_j q[j][t] hat[v][j], q [m_pad][n_time+1],
 * accumulate != 0 adds to phi.  Same role as dots_time_transform (utils/laplacian_inverse_socp.py:52-61).               */
int dots_time_transform_plain(const dots_ctx_t *ctx, int inverse, const double *q, int accumulate, void *stream);   /* rhs -> hat / hat -> phi   */
int dots_mode_solves(const dots_ctx_t *c, void *stream);                   /* hat <- (K+shift M)^-1 hat */
/* profiling aid: one pair of ring sweeps (sweep_mode 4) with a CUDA event before every launch; ms_out[i] = start of launch i ->
 * start of launch i+1, tag_out[i] = tree level (+1000: gather, +2000: backward).  Synchronises the stream.              */
int dots_ring_level_times(const dots_ctx_t *c, void *stream, float *ms_out, int32_t *tag_out, int cap, int *n_out);
int dots_grad_space(const dots_ctx_t *c, const double *phi, double *out, void *stream);  /* [nT+1][3][T] */
int dots_div_space(const dots_ctx_t *c, const double *x, double *out, void *stream);     /* [nT+1][V]    */

#ifdef __cplusplus
}
#endif
#endif /* DOTS_B200_H */
