// Driver of tests/test_nested_host.py::test_host_analysis_is_clean_under_thread_sanitizer: the threaded host analysis of
// csrc/host_order.cpp (mesh operators, ordering, matrix permutation, index maps) on a perturbed grid surface, compiled with
// -fsanitize=thread; prints a hash of the outputs so that runs at different thread counts can be compared.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdarg>
#include "dots_b200.h"
void dots_set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
int main() {
    const int64_t n = 220;                                  // grid of n x n vertices: 48 400 vertices, 95 922 triangles
    std::vector<double> xyz; std::vector<int64_t> tri;
    for (int64_t i = 0; i < n; ++i) for (int64_t j = 0; j < n; ++j) { xyz.push_back(i * 1.0 + 0.1 * ((i * 7 + j * 3) % 5)); xyz.push_back(j * 1.0); xyz.push_back(0.01 * ((i * j) % 11)); }
    for (int64_t i = 0; i + 1 < n; ++i) for (int64_t j = 0; j + 1 < n; ++j) {
        int64_t a = i * n + j, b = a + 1, c = a + n, d = c + 1;
        tri.insert(tri.end(), {a, b, c}); tri.insert(tri.end(), {b, d, c});
    }
    const int64_t V = n * n, T = (int64_t)tri.size() / 3;
    dots_mesh_t *m = nullptr;
    if (dots_mesh_create(V, T, xyz.data(), tri.data(), &m)) return 1;
    int64_t nnz = 0; dots_mesh_sizes(m, &nnz);
    std::vector<double> af(T), hat(9 * T), as(V), kv(nnz); std::vector<int64_t> kp(V + 1), ki(nnz), cp(V + 1), ct(3 * T), cc(3 * T);
    dots_mesh_export(m, af.data(), hat.data(), as.data(), kp.data(), ki.data(), kv.data(), cp.data(), ct.data(), cc.data());
    dots_mesh_destroy(m);
    std::vector<int64_t> ap(V + 1, 0), ai;                  // pattern without the diagonal
    for (int64_t v = 0; v < V; ++v) { for (int64_t q = kp[v]; q < kp[v + 1]; ++q) if (ki[q] != v) ai.push_back(ki[q]); ap[v + 1] = (int64_t)ai.size(); }
    dots_order_t *o = nullptr;
    if (dots_order_create(V, xyz.data(), ap.data(), ai.data(), 16, &o)) return 2;
    int64_t nn = 0, ft = 0; dots_order_sizes(o, &nn, &ft);
    std::vector<int64_t> perm(V), s(nn), b(nn), lvl(nn), par(nn), ch(2 * nn), fi(ft), cpos(2 * ft);
    dots_order_export(o, perm.data(), s.data(), b.data(), lvl.data(), par.data(), ch.data(), fi.data(), cpos.data());
    dots_order_destroy(o);
    std::vector<int64_t> iperm(V); for (int64_t i = 0; i < V; ++i) iperm[perm[i]] = i;
    std::vector<int64_t> op(V + 1), oi(nnz); std::vector<double> ov(nnz);
    if (dots_csr_permute(V, kp.data(), ki.data(), kv.data(), perm.data(), iperm.data(), op.data(), oi.data(), ov.data())) return 3;
    std::vector<int64_t> off(nn + 1, 0), foff(nn + 1, 0), uoff(nn + 1, 0);
    for (int64_t i = 0; i < nn; ++i) { off[i + 1] = off[i] + s[i]; foff[i + 1] = foff[i] + s[i] + b[i]; uoff[i + 1] = uoff[i] + b[i]; }
    std::vector<int32_t> apos(nnz), ppos(uoff[nn] + 1);
    if (dots_front_maps(V, nn, s.data(), b.data(), off.data(), foff.data(), fi.data(), uoff.data(), par.data(), op.data(), oi.data(), apos.data(), ppos.data())) return 4;
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t bytes) { const unsigned char *c = (const unsigned char *)p; for (size_t i = 0; i < bytes; ++i) { h ^= c[i]; h *= 1099511628211ull; } };
    mix(perm.data(), perm.size() * 8); mix(fi.data(), fi.size() * 8); mix(cpos.data(), cpos.size() * 8); mix(kv.data(), kv.size() * 8); mix(ov.data(), ov.size() * 8);
    mix(apos.data(), apos.size() * 4); mix(ppos.data(), (size_t)uoff[nn] * 4); mix(as.data(), as.size() * 8); mix(ct.data(), ct.size() * 8);
    printf("V=%lld T=%lld nodes=%lld hash=%016llx\n", (long long)V, (long long)T, (long long)nn, (unsigned long long)h);
    return 0;
}
