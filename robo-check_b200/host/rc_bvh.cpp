// rc_bvh.cpp -- median-split AABB tree (see rc_bvh.h).
#include "rc_bvh.h"

#include <algorithm>
#include <cfloat>
#include <cstring>
#include <numeric>
#include <utility>

namespace rc {
namespace {

struct Builder {
    const float *x, *y, *z;
    const int *v1, *v2, *v3;
    float margin;
    std::vector<float> cen;  // 3 per triangle
    Bvh out;

    void tri_bounds(int t, float lo[3], float hi[3]) const
    {
        const int idx[3] = {v1[t], v2[t], v3[t]};
        for (int k = 0; k < 3; k++) {
            lo[0] = std::min(lo[0], x[idx[k]]); hi[0] = std::max(hi[0], x[idx[k]]);
            lo[1] = std::min(lo[1], y[idx[k]]); hi[1] = std::max(hi[1], y[idx[k]]);
            lo[2] = std::min(lo[2], z[idx[k]]); hi[2] = std::max(hi[2], z[idx[k]]);
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
        if (n <= kLeafTris) {
            const int first = (int)out.order.size();
            for (int i = 0; i < n; i++) out.order.push_back(tris[i]);
            node[3] = bits(kLeafFlag | (uint32_t)first);
            node[7] = bits((uint32_t)n);
            std::memcpy(&out.nodes[(size_t)me * 8], node, sizeof node);
            return me;
        }
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
        int half = n / 2;
        half = std::max(kLeafTris, (half + kLeafTris - 1) / kLeafTris * kLeafTris);
        if (half >= n) half = n / 2;
        std::nth_element(tris, tris + half, tris + n, [&](int a, int b) {
            const float ca = cen[3 * (size_t)a + ax], cb = cen[3 * (size_t)b + ax];
            return ca < cb || (ca == cb && a < b);
        });
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

Bvh build_bvh(const float *x, const float *y, const float *z, const int *v1, const int *v2,
              const int *v3, int n_t, float margin)
{
    Builder b{x, y, z, v1, v2, v3, margin, {}, {}};
    b.cen.resize(3 * (size_t)std::max(n_t, 1));
    for (int t = 0; t < n_t; t++) {
        b.cen[3 * (size_t)t + 0] = (x[v1[t]] + x[v2[t]] + x[v3[t]]) / 3.0f;
        b.cen[3 * (size_t)t + 1] = (y[v1[t]] + y[v2[t]] + y[v3[t]]) / 3.0f;
        b.cen[3 * (size_t)t + 2] = (z[v1[t]] + z[v2[t]] + z[v3[t]]) / 3.0f;
    }
    std::vector<int> tris((size_t)n_t);
    std::iota(tris.begin(), tris.end(), 0);
    if (n_t > 0) b.build(tris.data(), n_t, 0);
    return b.out;
}

}  // namespace rc
