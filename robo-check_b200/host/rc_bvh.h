// rc_bvh.h -- binary AABB tree over the triangles of one mesh, built once per context on the
// host (surface-area heuristic, leaves of <= 4 triangles kept full).  The reference has no acceleration structure in its live pipeline (its dead MegaKernel
// precomputes per-triangle obstacle AABBs, test/MegaKernel.cu:31-35); culling here must never
// drop a pair the reference arithmetic would report, so every box is inflated by `margin`.
#pragma once
#include <cstdint>
#include <vector>

namespace rc {

constexpr int kLeafTris = 4;           // triangles per leaf (<= 4)
constexpr uint32_t kLeafFlag = 0x80000000u;

struct Bvh {
    // Two float4 per node:  (cx, cy, cz, link0)  (ex, ey, ez, link1)
    //   internal: link0 = left child index, link1 = right child index
    //   leaf:     link0 = kLeafFlag | first triangle (position in `order`), link1 = count
    std::vector<float> nodes;          // 8 floats per node, root = node 0
    std::vector<int> order;            // order[p] = original triangle index at leaf position p
    int n_nodes = 0;
    int depth = 0;
    // nodes at depth `d` (plus leaves above it): a cut through the tree, used to seed traversals
    std::vector<int> cut(int d) const;
};

// leaf_tris: triangles per leaf, 1 .. kLeafTris (the robot tree of a context whose obstacle is much
// finer than the robot is built with one triangle per leaf: a leaf box is then a triangle box)
Bvh build_bvh(const float *x, const float *y, const float *z, const int *v1, const int *v2,
              const int *v3, int n_t, float margin, int leaf_tris = kLeafTris);
// Same tree over arbitrary primitives given by their boxes (3 floats each) and sort keys.
Bvh build_bvh_boxes(const float *lo, const float *hi, const float *centre, int n, float margin,
                    int leaf_tris = kLeafTris);

// Bounding sphere of every triangle (smallest enclosing: mid-point of the longest edge if the
// triangle is obtuse, circumsphere otherwise), radius rounded up.  out = n_t x (x, y, z, r).
void triangle_spheres(const float *x, const float *y, const float *z, const int *v1, const int *v2,
                      const int *v3, int n_t, float *out);

}  // namespace rc
