// rc_spheres.cuh -- sphere-approximated robot against the obstacle mesh (BASELINE config 3).
//
// The reference has no such path (only a two-unit-sphere FCL smoke test, test/CPU-Sphere-Test.cpp:
// 8-18), so the semantics are defined here: the robot is a set of spheres in its local frame
// (default: the smallest enclosing sphere of every robot triangle, so the set covers the surface);
// a configuration is invalid iff some sphere, its centre moved by the configuration with the same
// pinned arithmetic as the mesh path, comes within its radius of some obstacle triangle
// (squared distance <= r*r, rc_device.cuh: point_tri_dist2).  There is no reference result to
// compare with; the tests check it against an independent CPU restatement of this definition.
//
// Same machinery as rc_culled.cuh: one warp per configuration, persistent grid, dual-tree descent
// (sphere tree in the robot frame x obstacle tree), leaf pairs of <= 4 spheres x <= 4 triangles.
#pragma once

namespace rc {

struct SphereArgs {
    const float4 *sph;        // [n_s] (x, y, z, r) in sphere-tree leaf order
    const float4 *sph_bvh, *obs_bvh;
    const float4 *obs_a, *obs_b, *obs_c;
    int n_s, n_sph_nodes, n_obs_nodes, n_obs_t;
    rc_aabb obs_box;
    float bs_x, bs_y, bs_z, bs_r;  // bounding sphere of the whole set (local frame)
    const rc_config *confs;
    long n;
    uint8_t *valid;
    uint32_t *bitmask;
    unsigned long long *counters, *queue;
    long tail_begin;
    // work list written by classify_kernel (cost class 0 only): configurations the bounding sphere /
    // bounding box of the set could not clear, with their matrices
    const uint32_t *work_idx;
    const float4 *work_mat;
    unsigned long long *work_count;
};

constexpr int kSphereWarps = 24;

template <bool STATS, bool GOBS>
__global__ void __launch_bounds__(kSphereWarps * 32, 1) check_spheres_kernel(const SphereArgs a)
{
    constexpr int SH = GOBS ? 21 : 16;
    constexpr uint32_t LOW = (1u << SH) - 1u;
    constexpr int kStack = GOBS ? kStackCapGlobal : kStackCap;
    constexpr int kWarpWork = kStack + kLeafQCap;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    float4 *s_sph = reinterpret_cast<float4 *>(smem_raw + 16);
    float4 *s_sph_bvh = s_sph + a.n_s;
    float4 *s_obs_bvh = s_sph_bvh + 2 * a.n_sph_nodes;
    float4 *s_obs_a = s_obs_bvh + (GOBS ? 0 : 2 * a.n_obs_nodes);
    float4 *s_obs_b = s_obs_a + (GOBS ? 0 : a.n_obs_t);
    float4 *s_obs_c = s_obs_b + (GOBS ? 0 : a.n_obs_t);
    uint32_t *s_work = reinterpret_cast<uint32_t *>(s_obs_c + (GOBS ? 0 : a.n_obs_t));
    auto obs_node = [&](int nb, int half) -> float4 {
        if constexpr (GOBS) return __ldg(a.obs_bvh + 2 * nb + half);
        else return s_obs_bvh[2 * nb + half];
    };
    auto obs_A = [&](int k) -> float4 { if constexpr (GOBS) return __ldg(a.obs_a + k); else return s_obs_a[k]; };
    auto obs_B = [&](int k) -> float4 { if constexpr (GOBS) return __ldg(a.obs_b + k); else return s_obs_b[k]; };
    auto obs_C = [&](int k) -> float4 { if constexpr (GOBS) return __ldg(a.obs_c + k); else return s_obs_c[k]; };

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bs = a.n_s * 16u, bsn = a.n_sph_nodes * 32u, bon = a.n_obs_nodes * 32u,
                       bo = a.n_obs_t * 16u;
        mbar_expect_tx(bar, bs + bsn + (GOBS ? 0u : bon + 3u * bo));
        tma_load_1d(s_sph, a.sph, bs, bar);
        tma_load_1d(s_sph_bvh, a.sph_bvh, bsn, bar);
        if (!GOBS) {
            tma_load_1d(s_obs_bvh, a.obs_bvh, bon, bar);
            tma_load_1d(s_obs_a, a.obs_a, bo, bar);
            tma_load_1d(s_obs_b, a.obs_b, bo, bar);
            tma_load_1d(s_obs_c, a.obs_c, bo, bar);
        }
    }
    mbar_wait(bar, 0);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t *stack = s_work + warp * kWarpWork;
    uint32_t *leafq = stack + kStack;
    unsigned long long st_valid = 0, st_invalid = 0, st_nodes = 0, st_leaves = 0,
                       st_tris = 0, st_over = 0;
    const long n_work = (long)__ldcg(a.work_count);
    const int chunk = (int)max(1L, min(16L, n_work / ((long)gridDim.x * kSphereWarps * 32)));
    long seen = 0;
    for (;;) {
        const int take = (seen + a.tail_begin < n_work) ? chunk : 1;
        long base = 0;
        if (lane == 0) base = (long)atomicAdd(a.work_count + kWorkBins, (unsigned long long)take);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n_work) break;
        seen = base;
        const long end = min(base + (long)take, n_work);
        for (long w = base; w < end; w++) {
            const long i = (long)__ldcg(a.work_idx + w);
            Mat34 m;
            {
                const float4 q0 = __ldcg(a.work_mat + 3 * i), q1 = __ldcg(a.work_mat + 3 * i + 1),
                             q2 = __ldcg(a.work_mat + 3 * i + 2);
                m.r00 = q0.x; m.r01 = q0.y; m.r02 = q0.z; m.tx = q0.w;
                m.r10 = q1.x; m.r11 = q1.y; m.r12 = q1.z; m.ty = q1.w;
                m.r20 = q2.x; m.r21 = q2.y; m.r22 = q2.z; m.tz = q2.w;
            }
            bool hit = false;
            const bool done = false;  // step 1 (classify_kernel) has cleared what the bounds can clear
            if (!done) {
                int sp = 1, lq = 0;
                bool overflow = false;
                if (lane == 0) stack[0] = 0u;
                __syncwarp();
                while (sp > 0 || lq > 0) {
                    if (lq >= 8 || sp == 0) {
                        const int cnt = min(lq, 8);
                        lq -= cnt;
                        const int slot = lane >> 2, si = lane & 3;
                        bool act = slot < cnt;
                        const uint32_t e = leafq[lq + (act ? slot : 0)];
                        const int na = (int)(e >> SH), nb = (int)(e & LOW);
                        const uint32_t sfirst = __float_as_uint(s_sph_bvh[2 * na].w) & 0x7fffffffu;
                        const int scount = (int)__float_as_uint(s_sph_bvh[2 * na + 1].w);
                        const uint32_t ofirst = __float_as_uint(obs_node(nb, 0).w) & 0x7fffffffu;
                        const int ocount = (int)__float_as_uint(obs_node(nb, 1).w);
                        act = act && si < scount;
                        const float4 s = s_sph[sfirst + (act ? si : 0)];
                        float cx, cy, cz;
                        xform(m, s.x, s.y, s.z, cx, cy, cz);
                        const float r2 = __fmul_rn(s.w, s.w);
                        bool h = false;
#pragma unroll
                        for (int q = 0; q < kLeafTrisDev; q++) {
                            const int k = ofirst + min(q, ocount - 1);
                            const float4 oa = obs_A(k), ob = obs_B(k), oc = obs_C(k);
                            const Tri o = {oa.x, oa.y, oa.z, ob.x, ob.y, ob.z, oc.x, oc.y, oc.z};
                            if (act && q < ocount && point_tri_dist2(cx, cy, cz, o) <= r2) h = true;
                        }
                        if (STATS) { st_leaves += cnt; st_tris += (act ? ocount : 0); }
                        if (__any_sync(kFull, h)) { hit = true; break; }
                        __syncwarp();
                        continue;
                    }
                    const int cnt = min(sp, 32);
                    sp -= cnt;
                    const bool act = lane < cnt;
                    const uint32_t e = stack[sp + (act ? lane : 0)];
                    const int na = (int)(e >> SH), nb = (int)(e & LOW);
                    const float4 r0 = s_sph_bvh[2 * na], r1 = s_sph_bvh[2 * na + 1];
                    const float4 o0 = obs_node(nb, 0), o1 = obs_node(nb, 1);
                    const bool ov = act && boxes_overlap(m, r0, r1, o0, o1);
                    const uint32_t ra = __float_as_uint(r0.w), rb = __float_as_uint(r1.w);
                    const uint32_t oa = __float_as_uint(o0.w), ob = __float_as_uint(o1.w);
                    const bool la = (ra & 0x80000000u) != 0, lb = (oa & 0x80000000u) != 0;
                    if (STATS) st_nodes += cnt;
                    __syncwarp();
                    const bool ll = ov && la && lb;
                    const unsigned mll = __ballot_sync(kFull, ll);
                    if (ll) leafq[lq + __popc(mll & lt)] = e;
                    lq += __popc(mll);
                    const bool two = ov && (la != lb), four = ov && !la && !lb;
                    const unsigned m2 = __ballot_sync(kFull, two), m4 = __ballot_sync(kFull, four);
                    const int total = 2 * __popc(m2) + 4 * __popc(m4);
                    if (sp + total > kStack) { overflow = true; break; }
                    const int off = sp + 2 * __popc(m2 & lt) + 4 * __popc(m4 & lt);
                    if (two) {
                        if (la) {
                            stack[off] = ((uint32_t)na << SH) | oa;
                            stack[off + 1] = ((uint32_t)na << SH) | ob;
                        } else {
                            stack[off] = (ra << SH) | (uint32_t)nb;
                            stack[off + 1] = (rb << SH) | (uint32_t)nb;
                        }
                    } else if (four) {
                        stack[off] = (ra << SH) | oa;
                        stack[off + 1] = (ra << SH) | ob;
                        stack[off + 2] = (rb << SH) | oa;
                        stack[off + 3] = (rb << SH) | ob;
                    }
                    sp += total;
                    __syncwarp();
                }
                if (overflow) {
                    // exhaustive fallback: lane = sphere, all obstacle triangles
                    if (STATS) st_over++;
                    hit = false;
                    for (int sb = 0; sb < a.n_s && !hit; sb += 32) {
                        const float4 s = s_sph[min(sb + lane, a.n_s - 1)];
                        float cx, cy, cz;
                        xform(m, s.x, s.y, s.z, cx, cy, cz);
                        const float r2 = __fmul_rn(s.w, s.w);
                        bool h = false;
                        for (int k = 0; k < a.n_obs_t; k++) {
                            const float4 oa = obs_A(k), ob = obs_B(k), oc = obs_C(k);
                            const Tri o = {oa.x, oa.y, oa.z, ob.x, ob.y, ob.z, oc.x, oc.y, oc.z};
                            if (point_tri_dist2(cx, cy, cz, o) <= r2) h = true;
                        }
                        hit = __any_sync(kFull, h);
                    }
                }
                __syncwarp();
                if (STATS) { if (hit) st_invalid++; else st_valid++; }
            }
            if (lane == 0) {
                if (a.valid) a.valid[i] = hit ? 0 : 1;
                if (a.bitmask && !hit) atomicOr(a.bitmask + (i >> 5), 1u << (i & 31));
            }
        }
    }
    if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) st_tris += __shfl_xor_sync(kFull, st_tris, o);
        if (lane == 0) {
            atomicAdd(a.counters + CNT_NARROW_VALID, st_valid);
            atomicAdd(a.counters + CNT_NARROW_INVALID, st_invalid);
            atomicAdd(a.counters + CNT_NODE_PAIRS, st_nodes);
            atomicAdd(a.counters + CNT_LEAF_PAIRS, st_leaves);
            atomicAdd(a.counters + CNT_TRI_PAIRS, st_tris);
            atomicAdd(a.counters + CNT_OVERFLOW, st_over);
        }
    }
    // the last CTA to finish leaves the list counters at zero for the next call (as the mesh path does)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.work_count + kWorkBins + 1, 1ULL) == (unsigned long long)gridDim.x - 1ULL)
            for (int q = 0; q < kWorkBins + 2; q++) a.work_count[q] = 0;
    }
}

cudaError_t launch_check_spheres(const rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid,
                                 uint32_t *d_bitmask, int want_stats, cudaStream_t stream,
                                 unsigned long long *queue)
{
    const DeviceTables &t = ctx->t;
    if (n <= 0) return cudaSuccess;
    if (t.n_sph <= 0) return cudaErrorInvalidConfiguration;
    cudaError_t e;
    if (!queue) queue = ctx->d_counters + CNT_QUEUE;
    if (!ctx->d_work_idx || ctx->work_cap < n) return cudaErrorMemoryAllocation;  // rc_check_spheres sizes it
    auto smem_bytes = [&](bool gobs) {
        size_t b = 16 + (size_t)t.n_sph * 16 + (size_t)t.n_sph_nodes * 32;
        if (!gobs) b += (size_t)t.n_obs_nodes * 32 + (size_t)t.n_obs_t * 48;
        return b + (size_t)kSphereWarps * ((gobs ? kStackCapGlobal : kStackCap) + kLeafQCap) * 4;
    };
    bool gobs = ctx->knob_global_obstacle;
    if (!gobs && (smem_bytes(false) > (size_t)ctx->max_smem_optin || t.n_obs_nodes > 65535 ||
                  t.n_sph_nodes > 65535))
        gobs = true;
    const size_t smem = smem_bytes(gobs);
    if (smem > (size_t)ctx->max_smem_optin) return cudaErrorInvalidConfiguration;
    if (gobs && (t.n_sph_nodes >= (1 << 11) || t.n_obs_nodes >= (1 << 21) || t.n_obs_t >= (1 << 21)))
        return cudaErrorInvalidConfiguration;
    SphereArgs a;
    a.sph = t.sph; a.sph_bvh = t.sph_bvh; a.obs_bvh = t.obs_bvh;
    a.obs_a = t.obs_a_sorted; a.obs_b = t.obs_b_sorted; a.obs_c = t.obs_c_sorted;
    a.n_s = t.n_sph; a.n_sph_nodes = t.n_sph_nodes; a.n_obs_nodes = t.n_obs_nodes; a.n_obs_t = t.n_obs_t;
    a.obs_box = t.obs_box;
    a.bs_x = t.sph_bound[0]; a.bs_y = t.sph_bound[1]; a.bs_z = t.sph_bound[2]; a.bs_r = t.sph_bound[3];
    a.confs = d_confs; a.n = n; a.valid = d_valid; a.bitmask = d_bitmask;
    a.counters = ctx->d_counters; a.queue = queue;
    const int grid = ctx->sm_count;
    a.work_idx = ctx->d_work_idx; a.work_mat = ctx->d_work_mat; a.work_count = ctx->d_work_count;
    // step 1: the per-thread classify kernel of the mesh path with the bounds of the sphere set
    // (bounding sphere, rotated local box; no vertex groups, so everything lands in cost class 0);
    // it writes the flags and whole bitmask words of the configurations it clears
    ClassifyArgs k{};
    k.confs = d_confs; k.n = n; k.valid = d_valid; k.bitmask = d_bitmask; k.obs_box = t.obs_box;
    k.sph_x = a.bs_x; k.sph_y = a.bs_y; k.sph_z = a.bs_z; k.sph_r = a.bs_r;
    k.box_cx = t.sph_box_c[0]; k.box_cy = t.sph_box_c[1]; k.box_cz = t.sph_box_c[2];
    k.box_ex = t.sph_box_e[0]; k.box_ey = t.sph_box_e[1]; k.box_ez = t.sph_box_e[2];
    k.n_bound = 0;
    k.work_idx = ctx->d_work_idx; k.work_mat = ctx->d_work_mat; k.work_cap = ctx->work_cap;
    k.work_count = ctx->d_work_count; k.counters = ctx->d_counters;
    const int cgrid = (int)min((n + 63) / 64, (long)ctx->sm_count * 256);
    if (want_stats) classify_kernel<true><<<cgrid, 64, 0, stream>>>(k);
    else classify_kernel<false><<<cgrid, 64, 0, stream>>>(k);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    a.tail_begin = (long)grid * kSphereWarps * 4;
    void (*kern)(const SphereArgs) =
        gobs ? (want_stats ? check_spheres_kernel<true, true> : check_spheres_kernel<false, true>)
             : (want_stats ? check_spheres_kernel<true, false> : check_spheres_kernel<false, false>);
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
        return e;
    kern<<<grid, kSphereWarps * 32, smem, stream>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess)  // nobody will reset the list counters step 1 filled
        cudaMemsetAsync(ctx->d_work_count, 0, kWorkCounters * sizeof(unsigned long long), stream);
    return e;
}

}  // namespace rc
