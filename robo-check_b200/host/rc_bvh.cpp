// rc_bvh.cpp -- median-split AABB tree (see rc_bvh.h).
#include "rc_bvh.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <numeric>
#include <utility>

namespace rc {
namespace {

constexpr int kSahMaxPrims = 16384;  // nodes larger than this are split at the median

struct Builder {
    const float *plo, *phi;  // primitive boxes, 3 floats each
    float margin;
    int leaf;                // primitives per leaf (1 .. kLeafTris)
    std::vector<float> cen;  // 3 per primitive
    Bvh out;

    void tri_bounds(int t, float lo[3], float hi[3]) const
    {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], plo[3 * (size_t)t + k]);
            hi[k] = std::max(hi[k], phi[3 * (size_t)t + k]);
        }
    }

    static float bits(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

    int build(int *tris, int n, int depth)
    {
        const int me = out.n_nodes++;
        out.nodes.resize((size_t)out.n_nodes * 8);
        out.depth = std::max(out.depth, depth);
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = 0; i < n; i++) tri_bounds(tris[i], lo, hi);
        float node[8];
        for (int k = 0; k < 3; k++) {
            node[k] = 0.5f * (lo[k] + hi[k]);
            // half extent measured from the rounded centre, then inflated
            node[4 + k] = std::max(hi[k] - node[k], node[k] - lo[k]) + margin;
        }
        if (n <= leaf) {
            const int first = (int)out.order.size();
            for (int i = 0; i < n; i++) out.order.push_back(tris[i]);
            node[3] = bits(kLeafFlag | (uint32_t)first);
            node[7] = bits((uint32_t)n);
            std::memcpy(&out.nodes[(size_t)me * 8], node, sizeof node);
            return me;
        }
        int half = 0;
        if (n <= kSahMaxPrims) {
            // Surface-area heuristic over every split position that keeps the leaves full
            // (multiples of the leaf size), on each axis: minimise area(L) |L| + area(R) |R|.
            // ~35 % fewer box-pair tests in the dual descent than the median split.
            double best = DBL_MAX;
            int best_ax = 0;
            std::vector<int> sorted(tris, tris + n), best_sorted;
            std::vector<float> suf((size_t)n * 6);
            for (int ax = 0; ax < 3; ax++) {
                std::sort(sorted.begin(), sorted.end(), [&](int a, int b) {
                    const float ca = cen[3 * (size_t)a + ax], cb = cen[3 * (size_t)b + ax];
                    return ca < cb || (ca == cb && a < b);
                });
                float slo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, shi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                for (int i = n - 1; i >= 0; i--) {
                    tri_bounds(sorted[i], slo, shi);
                    for (int k = 0; k < 3; k++) { suf[(size_t)i * 6 + k] = slo[k]; suf[(size_t)i * 6 + 3 + k] = shi[k]; }
                }
                float plo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, phi2[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                auto area = [](const float *lo3, const float *hi3) {
                    const double dx = hi3[0] - lo3[0], dy = hi3[1] - lo3[1], dz = hi3[2] - lo3[2];
                    return 2.0 * (dx * dy + dy * dz + dx * dz);
                };
                for (int i = 0; i + 1 < n; i++) {
                    tri_bounds(sorted[i], plo, phi2);
                    const int k = i + 1;  // left gets k primitives
                    if (k % leaf) continue;
                    const double cost = area(plo, phi2) * k + area(&suf[(size_t)k * 6], &suf[(size_t)k * 6 + 3]) * (n - k);
                    if (cost < best) { best = cost; best_ax = ax; half = k; }
                }
                if (best_ax == ax) best_sorted = sorted;
            }
            if (half > 0) std::copy(best_sorted.begin(), best_sorted.end(), tris);
        }
        if (half <= 0 || half >= n) {
            // very large nodes: median split along the longest centroid axis (O(n))
            float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (int i = 0; i < n; i++)
                for (int k = 0; k < 3; k++) {
                    clo[k] = std::min(clo[k], cen[3 * (size_t)tris[i] + k]);
                    chi[k] = std::max(chi[k], cen[3 * (size_t)tris[i] + k]);
                }
            int ax = 0;
            if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
            if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
            // left half rounded up to a multiple of the leaf size, so leaves are full
            half = n / 2;
            half = std::max(leaf, (half + leaf - 1) / leaf * leaf);
            if (half >= n) half = n / 2;
            std::nth_element(tris, tris + half, tris + n, [&](int a, int b) {
                const float ca = cen[3 * (size_t)a + ax], cb = cen[3 * (size_t)b + ax];
                return ca < cb || (ca == cb && a < b);
            });
        }
        const int l = build(tris, half, depth + 1);
        const int r = build(tris + half, n - half, depth + 1);
        node[3] = bits((uint32_t)l);
        node[7] = bits((uint32_t)r);
        std::memcpy(&out.nodes[(size_t)me * 8], node, sizeof node);
        return me;
    }
};

}  // namespace

std::vector<int> Bvh::cut(int d) const
{
    std::vector<int> out;
    if (n_nodes == 0) return out;
    std::vector<std::pair<int, int>> todo{{0, 0}};
    while (!todo.empty()) {
        const auto [node, depth_now] = todo.back();
        todo.pop_back();
        uint32_t l0, l1;
        std::memcpy(&l0, &nodes[(size_t)node * 8 + 3], 4);
        std::memcpy(&l1, &nodes[(size_t)node * 8 + 7], 4);
        if ((l0 & kLeafFlag) || depth_now == d) {
            out.push_back(node);
        } else {
            todo.push_back({(int)l1, depth_now + 1});
            todo.push_back({(int)l0, depth_now + 1});
        }
    }
    return out;
}

Bvh build_bvh_boxes(const float *lo, const float *hi, const float *centre, int n, float margin, int leaf_tris)
{
    Builder b{lo, hi, margin, std::max(1, std::min(leaf_tris, kLeafTris)), {}, {}};
    b.cen.assign(centre, centre + 3 * (size_t)std::max(n, 0));
    std::vector<int> prims((size_t)std::max(n, 0));
    std::iota(prims.begin(), prims.end(), 0);
    if (n > 0) b.build(prims.data(), n, 0);
    return b.out;
}

Bvh build_bvh(const float *x, const float *y, const float *z, const int *v1, const int *v2,
              const int *v3, int n_t, float margin, int leaf_tris)
{
    std::vector<float> lo(3 * (size_t)std::max(n_t, 1)), hi(lo.size()), cen(lo.size());
    for (int t = 0; t < n_t; t++) {
        const int idx[3] = {v1[t], v2[t], v3[t]};
        const float *c[3] = {x, y, z};
        for (int k = 0; k < 3; k++) {
            const float a = c[k][idx[0]], b2 = c[k][idx[1]], d = c[k][idx[2]];
            lo[3 * (size_t)t + k] = std::min(a, std::min(b2, d));
            hi[3 * (size_t)t + k] = std::max(a, std::max(b2, d));
            cen[3 * (size_t)t + k] = (a + b2 + d) / 3.0f;
        }
    }
    return build_bvh_boxes(lo.data(), hi.data(), cen.data(), n_t, margin, leaf_tris);
}

void triangle_spheres(const float *x, const float *y, const float *z, const int *v1, const int *v2,
                      const int *v3, int n_t, float *out)
{
    for (int t = 0; t < n_t; t++) {
        const double p[3][3] = {{x[v1[t]], y[v1[t]], z[v1[t]]}, {x[v2[t]], y[v2[t]], z[v2[t]]},
                                {x[v3[t]], y[v3[t]], z[v3[t]]}};
        auto d2 = [&](int i, int j) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += (p[i][k] - p[j][k]) * (p[i][k] - p[j][k]);
            return s;
        };
        // longest edge (i, j), opposite vertex k
        int i = 0, j = 1, k = 2;
        if (d2(1, 2) > d2(i, j)) { i = 1; j = 2; k = 0; }
        if (d2(2, 0) > d2(i, j)) { i = 2; j = 0; k = 1; }
        double c[3];
        for (int q = 0; q < 3; q++) c[q] = 0.5 * (p[i][q] + p[j][q]);
        double r2 = 0.25 * d2(i, j), dk = 0;
        for (int q = 0; q < 3; q++) dk += (p[k][q] - c[q]) * (p[k][q] - c[q]);
        if (dk > r2) {
            // acute: circumcentre = a + (|ac|^2 (ab x ac) x ab + |ab|^2 ac x (ab x ac)) / (2 |ab x ac|^2)
            double ab[3], ac[3], n[3], t1[3], t2[3];
            for (int q = 0; q < 3; q++) { ab[q] = p[1][q] - p[0][q]; ac[q] = p[2][q] - p[0][q]; }
            n[0] = ab[1] * ac[2] - ab[2] * ac[1]; n[1] = ab[2] * ac[0] - ab[0] * ac[2]; n[2] = ab[0] * ac[1] - ab[1] * ac[0];
            t1[0] = n[1] * ab[2] - n[2] * ab[1]; t1[1] = n[2] * ab[0] - n[0] * ab[2]; t1[2] = n[0] * ab[1] - n[1] * ab[0];
            t2[0] = ac[1] * n[2] - ac[2] * n[1]; t2[1] = ac[2] * n[0] - ac[0] * n[2]; t2[2] = ac[0] * n[1] - ac[1] * n[0];
            const double ab2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
            const double ac2 = ac[0] * ac[0] + ac[1] * ac[1] + ac[2] * ac[2];
            const double nn = 2.0 * (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (nn > 0) {
                r2 = 0;
                for (int q = 0; q < 3; q++) c[q] = p[0][q] + (ac2 * t1[q] + ab2 * t2[q]) / nn;
                for (int v = 0; v < 3; v++) {
                    double dv = 0;
                    for (int q = 0; q < 3; q++) dv += (p[v][q] - c[q]) * (p[v][q] - c[q]);
                    r2 = std::max(r2, dv);
                }
            } else {
                r2 = dk;
            }
        }
        for (int q = 0; q < 3; q++) out[4 * (size_t)t + q] = (float)c[q];
        // radius from the rounded centre, rounded up
        double rr = 0;
        for (int v = 0; v < 3; v++) {
            double dv = 0;
            for (int q = 0; q < 3; q++) {
                const double e = p[v][q] - (double)out[4 * (size_t)t + q];
                dv += e * e;
            }
            rr = std::max(rr, dv);
        }
        out[4 * (size_t)t + 3] = (float)(std::sqrt(rr) * (1.0 + 1e-6)) + 1e-6f;
    }
}

}  // namespace rc
