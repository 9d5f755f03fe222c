// Host-side setup (row f1): nested-dissection ordering + symbolic multifrontal structure of the surface Laplacian.
//
// The reference hands each of its nT+1 shifted Laplacians to SuperLU, which orders and analyses every one of them
// separately (utils/laplacian_inverse_socp.py:34-41).  Here ONE ordering / symbolic analysis serves all time modes
// (dots_socp_b200/nested.py explains the structure); this file is the native version of nested.dissect + nested.symbolic
// (a Python loop over ~18 000 tree nodes at V = 164k took 1 s of the 2.4 s setup).  It reproduces the Python reference
// implementation decision for decision (same split axis, same stable order, same tie breaks), so both produce the same
// permutation and the same front structure bit for bit (tests/test_nested_host.py checks that).
//
// Plain C++, no CUDA: geometric recursive bisection with vertex separators, then boundary sets by a post-order sweep.
// The bisection of the two halves of a part and the per-triangle / per-vertex passes of the mesh operators run on a few
// host threads (DOTS_HOST_THREADS, default: the hardware's share of this rank, at most 16); the results do not depend on the thread count
// (every sum keeps its order, the tree keeps its shape), which tests/test_nested_host.py checks bit for bit.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <new>
#include <numeric>
#include <system_error>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/dots_b200.h"

void dots_set_error(const char *fmt, ...);

namespace {

struct TreeNode {
    std::vector<int64_t> own;   // old vertex ids eliminated by this node, in elimination order
    int kid[2] = {-1, -1};
    int n_kids = 0;
};

}  // namespace

struct dots_order {
    int64_t n = 0;
    std::vector<int64_t> perm, s, b, level, parent, child, front_idx, child_pos;
    int64_t n_nodes = 0, front_total = 0;
};

namespace {

// Host threads for a pass over `items` work items (at least `grain` items per thread).
int host_threads(int64_t items, int64_t grain) {
    const char *e = std::getenv("DOTS_HOST_THREADS");                         // read per call: tests vary it within one process
    const char *lw = std::getenv("LOCAL_WORLD_SIZE");                         // torchrun: the ranks of a node share its cores
    const unsigned ranks = lw ? (unsigned)std::max(1, std::atoi(lw)) : 1u;
    const int want = e ? std::atoi(e) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency() / ranks));
    const int limit = std::max(1, std::min(want, 64));
    return (int)std::max<int64_t>(1, std::min<int64_t>(limit, items / std::max<int64_t>(grain, 1)));
}

// Stable sort of (column, value) pairs by column: insertion sort for the short rows of a surface mesh.
inline void sort_row_stable(std::vector<std::pair<int64_t, double>> &row) {
    if (row.size() > 256) {
        std::stable_sort(row.begin(), row.end(),
                         [](const std::pair<int64_t, double> &x, const std::pair<int64_t, double> &y) { return x.first < y.first; });
        return;
    }
    for (size_t i = 1; i < row.size(); ++i) {
        const std::pair<int64_t, double> x = row[i];
        size_t j = i;
        for (; j > 0 && row[j - 1].first > x.first; --j) row[j] = row[j - 1];
        row[j] = x;
    }
}

// fn(tid, lo, hi) over contiguous ranges of [0, n); the first exception of a worker is rethrown on the caller.
template <class F>
void parallel_ranges(int64_t n, int n_threads, F fn) {
    if (n_threads <= 1) { fn(0, (int64_t)0, n); return; }
    std::vector<std::thread> pool;
    std::vector<std::exception_ptr> err((size_t)n_threads);
    auto work = [&](int t) {
        try { fn(t, n * t / n_threads, n * (t + 1) / n_threads); } catch (...) { err[t] = std::current_exception(); }
    };
    pool.reserve((size_t)n_threads);
    for (int t = 1; t < n_threads; ++t) {
        try { pool.emplace_back(work, t); } catch (...) { work(t); }        // no thread to be had: this range runs here
    }
    work(0);
    for (auto &th : pool) th.join();
    for (auto &e : err)
        if (e) std::rethrow_exception(e);
}

// One bisection step of nested.dissect: split `verts` along the longest bounding-box axis at the median (stable order), take
// as separator the smaller of the two one-sided vertex frontiers.  `side` (one byte per vertex, -1 outside this call) is
// scratch; `sep`, `left`, `right` receive the three parts.
struct Bisector {
    const double *xyz;
    const int64_t *indptr, *indices;
    std::vector<int8_t> side;
    std::vector<std::pair<double, int64_t>> keyed;
    std::vector<char> fl, fr;
    std::vector<int64_t> keep;

    Bisector(int64_t n, const double *xyz_, const int64_t *indptr_, const int64_t *indices_)
        : xyz(xyz_), indptr(indptr_), indices(indices_), side((size_t)n, -1) {}

    void split(const std::vector<int64_t> &verts, std::vector<int64_t> &sep, std::vector<int64_t> &left, std::vector<int64_t> &right) {
        const int64_t m = (int64_t)verts.size();
        double lo[3], hi[3];
        for (int a = 0; a < 3; ++a) lo[a] = hi[a] = xyz[3 * verts[0] + a];
        for (int64_t i = 1; i < m; ++i)
            for (int a = 0; a < 3; ++a) {
                const double x = xyz[3 * verts[i] + a];
                if (x < lo[a]) lo[a] = x;
                if (x > hi[a]) hi[a] = x;
            }
        int axis = 0;
        for (int a = 1; a < 3; ++a)
            if (hi[a] - lo[a] > hi[axis] - lo[axis]) axis = a;                 // first maximum wins, like np.argmax
        keyed.resize((size_t)m);                                               // (coordinate, vertex) pairs: the sort then
        for (int64_t i = 0; i < m; ++i) keyed[i] = {xyz[3 * verts[i] + axis], verts[i]};   // runs over contiguous memory
        std::stable_sort(keyed.begin(), keyed.end(),
                         [](const std::pair<double, int64_t> &p, const std::pair<double, int64_t> &q) { return p.first < q.first; });
        const int64_t half = m / 2;
        left.resize((size_t)half);
        right.resize((size_t)(m - half));
        for (int64_t i = 0; i < half; ++i) { left[i] = keyed[i].second; side[left[i]] = 0; }
        for (int64_t i = half; i < m; ++i) { right[i - half] = keyed[i].second; side[right[i - half]] = 1; }
        auto frontier = [&](const std::vector<int64_t> &part, int8_t other, std::vector<char> &mask) {
            int64_t cnt = 0;
            mask.assign(part.size(), 0);
            for (size_t i = 0; i < part.size(); ++i) {
                const int64_t v = part[i];
                for (int64_t q = indptr[v]; q < indptr[v + 1]; ++q)
                    if (side[indices[q]] == other) { mask[i] = 1; ++cnt; break; }
            }
            return cnt;
        };
        const int64_t nl = frontier(left, 1, fl), nr = frontier(right, 0, fr);
        sep.clear();
        auto take = [&](std::vector<int64_t> &part, const std::vector<char> &mask) {
            keep.clear();
            for (size_t i = 0; i < part.size(); ++i) (mask[i] ? sep : keep).push_back(part[i]);
            part.swap(keep);
        };
        if (nl <= nr) take(left, fl); else take(right, fr);
        for (int64_t i = 0; i < m; ++i) side[verts[i]] = -1;
    }
};

// The subtree below `verts`, root at nodes[0] (kid indices local to `nodes`), on the calling thread.
void dissect_serial(Bisector &bis, std::vector<int64_t> verts, int64_t leaf_size, std::vector<TreeNode> &nodes) {
    struct Frame { int node; std::vector<int64_t> verts; };
    std::vector<Frame> stack;
    nodes.clear();
    nodes.emplace_back();
    stack.push_back({0, std::move(verts)});
    std::vector<int64_t> sep, left, right;
    while (!stack.empty()) {
        Frame fr_ = std::move(stack.back());
        stack.pop_back();
        const int me = fr_.node;
        if ((int64_t)fr_.verts.size() <= leaf_size) { nodes[me].own = std::move(fr_.verts); continue; }
        bis.split(fr_.verts, sep, left, right);
        nodes[me].own = sep;
        for (std::vector<int64_t> *part : {&left, &right}) {
            if (part->empty()) continue;
            const int kid = (int)nodes.size();
            nodes.emplace_back();
            nodes[me].kid[nodes[me].n_kids++] = kid;
            stack.push_back({kid, *part});
        }
    }
}

// nested.dissect: recurse on what is left of the halves until a part has <= leaf_size vertices.  While threads are left
// (`budget` > 1) and the part is large, the left half goes to a new thread (own scratch) and the right half stays here; the
// two subtrees are then appended behind their root.  The shape of the tree (left kid first) is that of the serial
// recursion, and the numbering the solver uses (post-order, `symbolic`) depends on nothing else.
void dissect_rec(int64_t n, const double *xyz, const int64_t *indptr, const int64_t *indices, int64_t leaf_size, int budget,
                 Bisector &bis, std::vector<int64_t> verts, std::vector<TreeNode> &nodes) {
    if (budget <= 1 || (int64_t)verts.size() <= std::max<int64_t>(leaf_size, 4096)) {
        dissect_serial(bis, std::move(verts), leaf_size, nodes);
        return;
    }
    std::vector<int64_t> sep, left, right;
    bis.split(verts, sep, left, right);
    std::vector<int64_t>().swap(verts);
    std::vector<TreeNode> sub[2];
    std::exception_ptr err;
    const int b_left = budget / 2;
    std::thread other;
    auto do_left = [&] {
        try {
            Bisector mine(n, xyz, indptr, indices);
            dissect_rec(n, xyz, indptr, indices, leaf_size, b_left, mine, std::move(left), sub[0]);
        } catch (...) { err = std::current_exception(); }
    };
    if (!left.empty()) {
        try { other = std::thread(do_left); } catch (const std::system_error &) { do_left(); }   // no thread to be had: inline
    }
    try {
        if (!right.empty()) dissect_rec(n, xyz, indptr, indices, leaf_size, budget - b_left, bis, std::move(right), sub[1]);
    } catch (...) {
        if (other.joinable()) other.join();
        throw;
    }
    if (other.joinable()) other.join();
    if (err) std::rethrow_exception(err);
    nodes.clear();
    nodes.reserve(1 + sub[0].size() + sub[1].size());
    nodes.emplace_back();
    nodes[0].own = std::move(sep);
    for (auto &part : sub) {
        if (part.empty()) continue;
        const int base = (int)nodes.size();
        nodes[0].kid[nodes[0].n_kids++] = base;
        for (TreeNode &nd : part) {
            for (int k = 0; k < nd.n_kids; ++k) nd.kid[k] += base;
            nodes.push_back(std::move(nd));
        }
    }
}

void dissect(int64_t n, const double *xyz, const int64_t *indptr, const int64_t *indices, int64_t leaf_size,
             std::vector<TreeNode> &nodes) {
    std::vector<int64_t> all((size_t)n);
    std::iota(all.begin(), all.end(), (int64_t)0);
    Bisector bis(n, xyz, indptr, indices);
    dissect_rec(n, xyz, indptr, indices, leaf_size, host_threads(n, 8192), bis, std::move(all), nodes);
}

// nested.symbolic: post-order numbering, boundary sets B(i) = (neighbours of S(i) and children's boundaries) beyond the
// node's own block, front row lists and the child -> parent row maps.
void symbolic(int64_t n, const int64_t *indptr, const int64_t *indices, const std::vector<TreeNode> &nodes, dots_order &o) {
    std::vector<int> post;
    post.reserve(nodes.size());
    {
        std::vector<std::pair<int, bool>> st;
        st.push_back({0, false});
        while (!st.empty()) {
            auto [nd, seen] = st.back();
            st.pop_back();
            if (seen) { post.push_back(nd); continue; }
            st.push_back({nd, true});
            for (int k = nodes[nd].n_kids - 1; k >= 0; --k) st.push_back({nodes[nd].kid[k], false});
        }
    }
    const int64_t N = (int64_t)post.size();
    std::vector<int64_t> id_of(nodes.size());
    for (int64_t i = 0; i < N; ++i) id_of[post[i]] = i;
    o.n = n;
    o.n_nodes = N;
    o.perm.clear();
    o.perm.reserve((size_t)n);
    o.s.assign(N, 0); o.b.assign(N, 0); o.level.assign(N, 0); o.parent.assign(N, -1); o.child.assign(2 * N, -1);
    std::vector<int64_t> off(N + 1, 0);
    for (int64_t i = 0; i < N; ++i) {
        const TreeNode &nd = nodes[post[i]];
        o.s[i] = (int64_t)nd.own.size();
        off[i + 1] = off[i] + o.s[i];
        o.perm.insert(o.perm.end(), nd.own.begin(), nd.own.end());
        for (int slot = 0; slot < nd.n_kids; ++slot) {
            const int64_t k = id_of[nd.kid[slot]];
            o.parent[k] = i;
            o.child[2 * i + slot] = k;
            o.level[i] = std::max(o.level[i], o.level[k] + 1);
        }
    }
    std::vector<int64_t> iperm((size_t)n);
    for (int64_t i = 0; i < n; ++i) iperm[o.perm[i]] = i;
    // Boundary sets, level by level (leaves first): the nodes of one level only read their children's sets, so threads
    // share a level's nodes; every set is the sorted union of the same candidates whatever the thread count.
    std::vector<std::vector<int64_t>> bset(N);
    int64_t n_levels = 0;
    for (int64_t i = 0; i < N; ++i) n_levels = std::max(n_levels, o.level[i] + 1);
    std::vector<int64_t> lvl_ptr(n_levels + 1, 0), lvl_nodes(N);
    for (int64_t i = 0; i < N; ++i) ++lvl_ptr[o.level[i] + 1];
    for (int64_t l = 0; l < n_levels; ++l) lvl_ptr[l + 1] += lvl_ptr[l];
    {
        std::vector<int64_t> at(lvl_ptr.begin(), lvl_ptr.end() - 1);
        for (int64_t i = 0; i < N; ++i) lvl_nodes[at[o.level[i]]++] = i;
    }
    const int nt = host_threads(n, 8192);
    for (int64_t l = 0; l < n_levels; ++l) {
        const int64_t *ids = lvl_nodes.data() + lvl_ptr[l];
        const int64_t cnt = lvl_ptr[l + 1] - lvl_ptr[l];
        parallel_ranges(cnt, (int)std::min<int64_t>(nt, std::max<int64_t>(1, cnt / 8)), [&](int, int64_t q0, int64_t q1) {
            std::vector<int64_t> cand;
            for (int64_t q = q0; q < q1; ++q) {
                const int64_t i = ids[q], hi = off[i + 1];
                cand.clear();
                for (int64_t r = off[i]; r < hi; ++r) {
                    const int64_t v = o.perm[r];
                    for (int64_t e = indptr[v]; e < indptr[v + 1]; ++e) {
                        const int64_t w = iperm[indices[e]];
                        if (w >= hi) cand.push_back(w);
                    }
                }
                for (int slot = 0; slot < 2; ++slot) {
                    const int64_t k = o.child[2 * i + slot];
                    if (k < 0) continue;
                    for (int64_t w : bset[k])
                        if (w >= hi) cand.push_back(w);
                }
                std::sort(cand.begin(), cand.end());
                cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
                bset[i] = cand;
                o.b[i] = (int64_t)cand.size();
            }
        });
    }
    std::vector<int64_t> front_off(N + 1, 0);
    for (int64_t i = 0; i < N; ++i) front_off[i + 1] = front_off[i] + o.s[i] + o.b[i];
    o.front_total = front_off[N];
    o.front_idx.assign((size_t)o.front_total, 0);
    o.child_pos.assign((size_t)(2 * o.front_total), -1);
    parallel_ranges(N, (int)std::min<int64_t>(nt, std::max<int64_t>(1, N / 256)), [&](int, int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; ++i) {
            int64_t *rows = o.front_idx.data() + front_off[i];
            const int64_t nf = o.s[i] + o.b[i];
            for (int64_t r = 0; r < o.s[i]; ++r) rows[r] = off[i] + r;
            std::copy(bset[i].begin(), bset[i].end(), rows + o.s[i]);
            for (int slot = 0; slot < 2; ++slot) {
                const int64_t k = o.child[2 * i + slot];
                if (k < 0 || o.b[k] == 0) continue;
                int64_t *cp = o.child_pos.data() + (size_t)slot * o.front_total + front_off[i];
                for (int64_t j = 0; j < o.b[k]; ++j) {
                    const int64_t *hit = std::lower_bound(rows, rows + nf, bset[k][j]);
                    cp[hit - rows] = j;                                 // the child's boundary always embeds in the parent's front
                }
            }
        }
    });
}

}  // namespace

extern "C" int dots_order_create(int64_t n_vert, const double *vertices, const int64_t *adj_ptr, const int64_t *adj_idx,
                                 int64_t leaf_size, dots_order_t **out) {
    if (!out || !vertices || !adj_ptr || !adj_idx || n_vert <= 0 || leaf_size < 1) {
        dots_set_error("dots_order_create: bad argument (n_vert=%lld, leaf_size=%lld)", (long long)n_vert, (long long)leaf_size);
        return -1;
    }
    *out = nullptr;
    try {
        std::unique_ptr<dots_order> o(new dots_order());
        std::vector<TreeNode> nodes;
        dissect(n_vert, vertices, adj_ptr, adj_idx, leaf_size, nodes);
        symbolic(n_vert, adj_ptr, adj_idx, nodes, *o);
        if ((int64_t)o->perm.size() != n_vert) {
            dots_set_error("dots_order_create: dissection covered %lld of %lld vertices", (long long)o->perm.size(), (long long)n_vert);
            return -1;
        }
        *out = o.release();
    } catch (const std::exception &) {
        dots_set_error("dots_order_create: out of host memory or threads");
        return -1;
    }
    return 0;
}

extern "C" int dots_order_sizes(const dots_order_t *o, int64_t *n_nodes, int64_t *front_total) {
    if (!o || !n_nodes || !front_total) { dots_set_error("dots_order_sizes: null argument"); return -1; }
    *n_nodes = o->n_nodes;
    *front_total = o->front_total;
    return 0;
}

extern "C" int dots_order_export(const dots_order_t *o, int64_t *perm, int64_t *s, int64_t *b, int64_t *level, int64_t *parent,
                                 int64_t *child, int64_t *front_idx, int64_t *child_pos) {
    if (!o || !perm || !s || !b || !level || !parent || !child || !front_idx || !child_pos) {
        dots_set_error("dots_order_export: null argument");
        return -1;
    }
    auto put = [](int64_t *dst, const std::vector<int64_t> &src) { if (!src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(int64_t)); };
    put(perm, o->perm); put(s, o->s); put(b, o->b); put(level, o->level); put(parent, o->parent); put(child, o->child);
    put(front_idx, o->front_idx); put(child_pos, o->child_pos);
    return 0;
}

extern "C" int dots_order_destroy(dots_order_t *o) {
    delete o;
    return 0;
}

// Per-entry operand rows of the ring-streamed sweeps (csrc/sweep_ring.cu): for every panel entry, in streaming order, the
// row of Z = [hat | ywork] it is multiplied with; bit 31 marks the last entry of an output.
//   forward  (row-major panel)       entry (row i, col j < min(i+1, s)):  off + j                 last: j == min(i+1, s) - 1
//   backward (column-major copy)     entry (col j, row i in [j, s+b)):    bidx[front_off + i]     last: i == s + b - 1
extern "C" int dots_ring_entry_rows(int64_t n_nodes, const int64_t *s, const int64_t *b, const int64_t *off, const int64_t *front_off,
                                    const int64_t *panel_off, const int32_t *bidx, int32_t *rows_fwd, int32_t *rows_bwd)
{
    if (!s || !b || !off || !front_off || !panel_off || !bidx || !rows_fwd || !rows_bwd) { dots_set_error("dots_ring_entry_rows: null argument"); return -1; }
    const int32_t LAST = (int32_t)0x80000000u;
    for (int64_t nd = 0; nd < n_nodes; ++nd) {
        const int64_t sn = s[nd], bn = b[nd], o = off[nd];
        if (sn <= 0) continue;
        int32_t *f = rows_fwd + panel_off[nd];
        for (int64_t i = 0; i < sn + bn; ++i) {
            const int64_t len = (i + 1 < sn) ? i + 1 : sn;
            for (int64_t j = 0; j < len; ++j) *f++ = (int32_t)(o + j) | (j == len - 1 ? LAST : 0);
        }
        int32_t *t = rows_bwd + panel_off[nd];
        const int32_t *bi = bidx + front_off[nd];
        for (int64_t j = 0; j < sn; ++j)
            for (int64_t i = j; i < sn + bn; ++i) *t++ = bi[i] | (i == sn + bn - 1 ? LAST : 0);
        if (f - rows_fwd != panel_off[nd + 1] || t - rows_bwd != panel_off[nd + 1]) { dots_set_error("dots_ring_entry_rows: panel size mismatch at node %lld", (long long)nd); return -1; }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Mesh operators of the hot path (row f1; reference utils/surface_pre_computations_socp.py:11-132, Python loops over the
// triangles): per triangle the area, the P1 hat-function gradients g[f][k][:] (altitude vector of corner k over its squared
// length, :30-37) and the corner cotangents; per vertex the incident area sum (:121-124); the cotan stiffness matrix
// K = -L (:68-84) in CSR with sorted columns; and the CSR vertex -> incident corner lists (:112-127).
//   mesh_create: computes everything, returns a handle;  mesh_sizes: nnz(K);  mesh_export: copies into caller arrays.
struct dots_mesh {
    int64_t V = 0, T = 0;
    std::vector<double> area_f, hat, area_sum, kval;
    std::vector<int64_t> kptr, kidx, cptr, ctri, ccorner;
};

namespace {
inline void sub3(const double *a, const double *b, double *o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void cross3(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
inline double norm3(const double *a) { return std::sqrt(dot3(a, a)); }
}  // namespace

extern "C" int dots_mesh_create(int64_t V, int64_t T, const double *vertices, const int64_t *triangles, dots_mesh_t **out)
{
    if (!vertices || !triangles || !out || V <= 0 || T <= 0) { dots_set_error("dots_mesh_create: bad arguments"); return -1; }
    *out = nullptr;
    dots_mesh *m = nullptr;
    try {
        m = new dots_mesh;
        m->V = V; m->T = T;
        m->area_f.assign(T, 0.0); m->hat.assign(9 * T, 0.0); m->area_sum.assign(V, 0.0);
        std::vector<double> w(3 * T);                                      // half cotangent of the angle at corner k
        // ---- per triangle (threads over triangle ranges): area, hat gradients, corner cotangents
        const int nt_tri = host_threads(T, 16384);
        std::vector<int64_t> bad((size_t)nt_tri, -1);                      // first triangle with a vertex id out of range
        parallel_ranges(T, nt_tri, [&](int tid, int64_t f0, int64_t f1) {
            for (int64_t f = f0; f < f1; ++f) {
                const int64_t *t = triangles + 3 * f;
                if (t[0] < 0 || t[0] >= V || t[1] < 0 || t[1] >= V || t[2] < 0 || t[2] >= V) { bad[tid] = f; return; }
                const double *p0 = vertices + 3 * t[0], *p1 = vertices + 3 * t[1], *p2 = vertices + 3 * t[2];
                double e[3][3], n[3];                                      // e01, e12, e20
                sub3(p1, p0, e[0]); sub3(p2, p1, e[1]); sub3(p0, p2, e[2]);
                cross3(e[0], e[1], n);
                m->area_f[f] = 0.5 * norm3(n);
                for (int k = 0; k < 3; ++k) {                              // altitude(into = e[k], along = e[(k+1)%3])
                    const double *into = e[k], *along = e[(k + 1) % 3];
                    const double coef = dot3(into, along) / dot3(along, along);
                    double h[3] = {-into[0] + along[0] * coef, -into[1] + along[1] * coef, -into[2] + along[2] * coef};
                    const double hh = dot3(h, h);
                    for (int x = 0; x < 3; ++x) m->hat[(size_t)f * 9 + k * 3 + x] = h[x] / hh;
                    const double *a = e[k], *bb = e[(k + 2) % 3];          // cot(e[k], -e[(k+2)%3]) = angle at corner k
                    double nb[3] = {-bb[0], -bb[1], -bb[2]}, c[3];
                    cross3(a, nb, c);
                    w[3 * f + k] = 0.5 * (dot3(a, nb) / norm3(c));
                }
            }
        });
        for (int64_t f : bad)
            if (f >= 0) {
                const int64_t *t = triangles + 3 * f;
                const int64_t v = (t[0] < 0 || t[0] >= V) ? t[0] : ((t[1] < 0 || t[1] >= V) ? t[1] : t[2]);
                dots_set_error("dots_mesh_create: triangle %lld has vertex %lld", (long long)f, (long long)v);
                delete m;
                return -1;
            }
        // ---- corners around every vertex, ordered by (corner, triangle): column k*T + f of the reference's incidence map.
        // Counts per (corner slot, vertex) give every list entry its place, so threads can fill disjoint vertex ranges.
        std::vector<int64_t> cnt_k(3 * (size_t)V, 0);
        for (int64_t f = 0; f < T; ++f)
            for (int k = 0; k < 3; ++k) ++cnt_k[(size_t)k * V + triangles[3 * f + k]];
        m->cptr.assign(V + 1, 0);
        for (int64_t v = 0; v < V; ++v) m->cptr[v + 1] = m->cptr[v] + cnt_k[v] + cnt_k[V + v] + cnt_k[2 * V + v];
        m->ctri.assign(3 * T, 0); m->ccorner.assign(3 * T, 0);
        const int nt_v = host_threads(V, 8192);
        // ---- per vertex (threads over vertex ranges): incident area sum and row v of the cotan stiffness matrix K = -L.
        // Every corner k of a triangle weights its opposite edge (a, b): K[a][b] -= w, K[b][a] -= w, K[a][a] += w, K[b][b] += w;
        // a row collects its terms triangle by triangle (ascending), corner by corner, then sums equal columns in that order.
        std::vector<std::vector<int64_t>> row_idx((size_t)nt_v);
        std::vector<std::vector<double>> row_val((size_t)nt_v);
        std::vector<int64_t> row_nnz((size_t)V, 0);
        parallel_ranges(V, nt_v, [&](int tid, int64_t v0, int64_t v1) {
            {   // this range's share of the corner lists: one pass over the triangles, three cursors per vertex
                std::vector<int64_t> pos(3 * (size_t)(v1 - v0));
                for (int64_t v = v0; v < v1; ++v) {
                    int64_t p = m->cptr[v];
                    for (int k = 0; k < 3; ++k) { pos[(size_t)k * (v1 - v0) + (v - v0)] = p; p += cnt_k[(size_t)k * V + v]; }
                }
                for (int64_t f = 0; f < T; ++f)
                    for (int k = 0; k < 3; ++k) {
                        const int64_t v = triangles[3 * f + k];
                        if (v < v0 || v >= v1) continue;
                        const int64_t p = pos[(size_t)k * (v1 - v0) + (v - v0)]++;
                        m->ctri[p] = f; m->ccorner[p] = k;
                    }
            }
            std::vector<int64_t> &kidx = row_idx[tid];
            std::vector<double> &kval = row_val[tid];
            kidx.reserve(8 * (size_t)(v1 - v0)); kval.reserve(8 * (size_t)(v1 - v0));
            std::vector<int64_t> tris;
            std::vector<std::pair<int64_t, double>> row;
            for (int64_t v = v0; v < v1; ++v) {
                tris.assign(m->ctri.begin() + m->cptr[v], m->ctri.begin() + m->cptr[v + 1]);
                std::sort(tris.begin(), tris.end());
                double asum = 0.0;
                for (int64_t f : tris) asum += m->area_f[f];               // once per incident corner, ascending triangles
                m->area_sum[v] = asum;
                tris.erase(std::unique(tris.begin(), tris.end()), tris.end());
                row.clear();
                for (int64_t f : tris) {
                    const int64_t *t = triangles + 3 * f;
                    for (int k = 0; k < 3; ++k) {
                        const int64_t a = t[(k + 1) % 3], b = t[(k + 2) % 3];
                        const double wk = w[3 * f + k];
                        if (a == v) { row.emplace_back(b, -wk); row.emplace_back(a, wk); }
                        if (b == v) { row.emplace_back(a, -wk); row.emplace_back(b, wk); }
                    }
                }
                sort_row_stable(row);                                      // by column; equal columns keep their order
                const size_t before = kidx.size();
                for (size_t q = 0; q < row.size();) {
                    size_t e2 = q;
                    double sum = 0.0;
                    while (e2 < row.size() && row[e2].first == row[q].first) sum += row[e2++].second;
                    kidx.push_back(row[q].first); kval.push_back(sum);
                    q = e2;
                }
                row_nnz[v] = (int64_t)(kidx.size() - before);
            }
        });
        m->kptr.assign(V + 1, 0);
        for (int64_t v = 0; v < V; ++v) m->kptr[v + 1] = m->kptr[v] + row_nnz[v];
        m->kidx.resize((size_t)m->kptr[V]); m->kval.resize((size_t)m->kptr[V]);
        for (int t = 0; t < nt_v; ++t) {
            const int64_t at = m->kptr[V * t / nt_v];
            std::copy(row_idx[t].begin(), row_idx[t].end(), m->kidx.begin() + at);
            std::copy(row_val[t].begin(), row_val[t].end(), m->kval.begin() + at);
        }
    } catch (const std::exception &) {
        delete m;
        dots_set_error("dots_mesh_create: out of host memory or threads");
        return -1;
    }
    *out = m;
    return 0;
}

extern "C" int dots_mesh_sizes(const dots_mesh_t *m, int64_t *nnz)
{
    if (!m || !nnz) { dots_set_error("dots_mesh_sizes: null argument"); return -1; }
    *nnz = (int64_t)m->kidx.size();
    return 0;
}

extern "C" int dots_mesh_export(const dots_mesh_t *m, double *area_f, double *hat, double *area_sum, int64_t *k_ptr, int64_t *k_idx,
                                double *k_val, int64_t *c_ptr, int64_t *c_tri, int64_t *c_corner)
{
    if (!m) { dots_set_error("dots_mesh_export: null handle"); return -1; }
    auto cp = [](auto *dst, const auto &src) { if (dst) std::copy(src.begin(), src.end(), dst); };
    cp(area_f, m->area_f); cp(hat, m->hat); cp(area_sum, m->area_sum); cp(k_ptr, m->kptr); cp(k_idx, m->kidx); cp(k_val, m->kval);
    cp(c_ptr, m->cptr); cp(c_tri, m->ctri); cp(c_corner, m->ccorner);
    return 0;
}

extern "C" int dots_mesh_destroy(dots_mesh_t *m) { delete m; return 0; }

// CSR vertex -> incident corner ids (k * T + f), ordered by (corner, triangle), for any triangle numbering (the engine calls it
// on the renumbered mesh).  c_ptr [V+1] int32, c_idx [3T] int32.
extern "C" int dots_corner_lists(int64_t V, int64_t T, const int64_t *triangles, int32_t *c_ptr, int32_t *c_idx)
{
    if (!triangles || !c_ptr || !c_idx || V <= 0 || T <= 0 || 3 * T > 0x7fffffffLL) { dots_set_error("dots_corner_lists: bad arguments"); return -1; }
    std::fill(c_ptr, c_ptr + V + 1, 0);
    for (int64_t i = 0; i < 3 * T; ++i) {
        if (triangles[i] < 0 || triangles[i] >= V) { dots_set_error("dots_corner_lists: vertex id out of range"); return -1; }
        ++c_ptr[triangles[i] + 1];
    }
    for (int64_t v = 0; v < V; ++v) c_ptr[v + 1] += c_ptr[v];
    std::vector<int32_t> pos(c_ptr, c_ptr + V);
    for (int k = 0; k < 3; ++k)
        for (int64_t f = 0; f < T; ++f) c_idx[pos[triangles[3 * f + k]]++] = (int32_t)(k * T + f);
    return 0;
}

// Index maps of the numeric assembly (nested.front_maps is the numpy statement and the checker).  All index lists are
// ascending (front_idx inside a node, CSR columns inside a row), so every map is a two-pointer merge:
//   a_pos[q]      position, in the front of the node that owns row(q), of CSR entry q of the permuted matrix; -1 when the
//                 column lies before the owner's first vertex (already eliminated) or outside its front;
//   parent_pos[p] row, in the PARENT's front, of boundary row p of every node (concatenated like upd_off); -1 at the root.
extern "C" int dots_front_maps(int64_t n_vert, int64_t n_nodes, const int64_t *s, const int64_t *b, const int64_t *off,
                               const int64_t *front_off, const int64_t *front_idx, const int64_t *upd_off, const int64_t *parent,
                               const int64_t *a_ptr, const int64_t *a_idx, int32_t *a_pos, int32_t *parent_pos)
{
    if (!s || !b || !off || !front_off || !front_idx || !upd_off || !parent || !a_ptr || !a_idx || !a_pos || !parent_pos || n_vert <= 0 ||
        n_nodes <= 0) {
        dots_set_error("dots_front_maps: bad arguments");
        return -1;
    }
    std::atomic<int64_t> failed{-1};                                       // a node whose boundary does not embed in its parent's front
    try {
        parallel_ranges(n_nodes, host_threads(n_nodes, 512), [&](int, int64_t k0, int64_t k1) {
            for (int64_t k = k0; k < k1; ++k) {
                const int64_t *fi = front_idx + front_off[k];
                const int64_t nf = s[k] + b[k];
                for (int64_t r = off[k]; r < off[k] + s[k]; ++r) {
                    int64_t j = 0;
                    for (int64_t q = a_ptr[r]; q < a_ptr[r + 1]; ++q) {
                        const int64_t col = a_idx[q];
                        int32_t pos = -1;
                        if (col >= off[k]) {
                            while (j < nf && fi[j] < col) ++j;
                            if (j < nf && fi[j] == col) pos = (int32_t)j;
                        }
                        a_pos[q] = pos;
                    }
                }
                const int64_t par = parent[k];
                int32_t *pp = parent_pos + upd_off[k];
                if (par < 0) {
                    for (int64_t p = 0; p < b[k]; ++p) pp[p] = -1;
                    continue;
                }
                const int64_t *pf = front_idx + front_off[par];
                const int64_t npf = s[par] + b[par];
                int64_t j = 0;
                for (int64_t p = 0; p < b[k]; ++p) {
                    const int64_t vtx = fi[s[k] + p];
                    while (j < npf && pf[j] < vtx) ++j;
                    if (j >= npf || pf[j] != vtx) { failed.store(k); return; }
                    pp[p] = (int32_t)j;
                }
            }
        });
    } catch (const std::exception &) { dots_set_error("dots_front_maps: out of host memory or threads"); return -1; }
    if (failed.load() >= 0) {
        dots_set_error("dots_front_maps: boundary of node %lld does not embed in its parent's front", (long long)failed.load());
        return -1;
    }
    return 0;
}

// Symmetric permutation of a CSR matrix: out = A[perm][:, perm] with ascending columns in every row (what the numeric
// assembly reads; scipy's two fancy-indexing passes + sort_indices took 45 ms at V = 164k).  perm: new -> old; iperm: old -> new.
extern "C" int dots_csr_permute(int64_t n, const int64_t *a_ptr, const int64_t *a_idx, const double *a_val, const int64_t *perm,
                                const int64_t *iperm, int64_t *o_ptr, int64_t *o_idx, double *o_val)
{
    if (!a_ptr || !a_idx || !a_val || !perm || !iperm || !o_ptr || !o_idx || !o_val || n <= 0) { dots_set_error("dots_csr_permute: bad arguments"); return -1; }
    o_ptr[0] = 0;
    for (int64_t r = 0; r < n; ++r) {
        if (perm[r] < 0 || perm[r] >= n) { dots_set_error("dots_csr_permute: perm[%lld] = %lld", (long long)r, (long long)perm[r]); return -1; }
        o_ptr[r + 1] = o_ptr[r] + (a_ptr[perm[r] + 1] - a_ptr[perm[r]]);
    }
    try {
        parallel_ranges(n, host_threads(n, 8192), [&](int, int64_t r0, int64_t r1) {
            std::vector<std::pair<int64_t, double>> row;
            for (int64_t r = r0; r < r1; ++r) {
                const int64_t old = perm[r];
                row.clear();
                for (int64_t q = a_ptr[old]; q < a_ptr[old + 1]; ++q) row.emplace_back(iperm[a_idx[q]], a_val[q]);
                sort_row_stable(row);
                int64_t at = o_ptr[r];
                for (const auto &e : row) { o_idx[at] = e.first; o_val[at++] = e.second; }
            }
        });
    } catch (const std::exception &) { dots_set_error("dots_csr_permute: out of host memory or threads"); return -1; }
    return 0;
}
