// rc_kernels.cu -- sm_100a kernels of the per-configuration collision check.
//
//   check_brute_kernel   fused transform + AABB + broad + exhaustive narrow phase; one warp per
//                        configuration, persistent grid with an atomic work queue.  Lane = robot
//                        triangle, obstacle triangles are broadcast from shared memory (SoA of
//                        float4 staged with TMA bulk copies).  Exact by construction: it decides
//                        every one of the n_rob_t x n_obs_t pairs the reference decides
//                        (narrowPhaseKernel_coarse, narrow-phase.cu:586-677).
//   check_culled_kernel  same result, BVH-culled (rc_culled.cuh).
//   stage kernels        the un-fused stage API (transform / AABB / broad / narrow).
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "rc_device.cuh"
#include "rc_internal.h"

namespace rc {

constexpr int kWarpsPerCta = 12;  // x 2 CTAs per SM = 24 warps (<= 85 registers)
constexpr int kThreads = kWarpsPerCta * 32;
constexpr unsigned kFull = 0xffffffffu;
__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }

// ------------------------------------------------------------------------------------------
// Obstacle planes: configuration independent, computed once per context with the same
// arithmetic the reference re-executes per pair (compute_plane_sep, narrow-phase.cu:629).
__global__ void obstacle_planes_kernel(const float4 *a, const float4 *b, const float4 *c,
                                       float4 *a_out, float4 *b_out, float4 *c_out, float *d_out,
                                       int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 A = a[k], B = b[k], C = c[k];
    const Tri t = {A.x, A.y, A.z, B.x, B.y, B.z, C.x, C.y, C.z};
    const Plane p = plane_of(t);
    a_out[k] = make_float4(A.x, A.y, A.z, p.nx);
    b_out[k] = make_float4(B.x, B.y, B.z, p.ny);
    c_out[k] = make_float4(C.x, C.y, C.z, p.nz);
    d_out[k] = p.s;
}

cudaError_t launch_obstacle_planes(const float4 *a, const float4 *b, const float4 *c, float4 *a_out,
                                   float4 *b_out, float4 *c_out, float *d_out, int n,
                                   cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    obstacle_planes_kernel<<<(n + 127) / 128, 128, 0, stream>>>(a, b, c, a_out, b_out, c_out, d_out, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// shared helpers

__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

__device__ __forceinline__ Mat34 load_matrix(const rc_config *confs, long i)
{
    // 24-byte records: three 8-byte loads (always 8-byte aligned)
    const float2 *p = reinterpret_cast<const float2 *>(confs + i);
    const float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    return make_matrix(a.x, a.y, b.x, b.y, c.x, c.y);
}

// one configuration per warp (all lanes call it with the same i)
__device__ __forceinline__ Mat34 load_matrix_warp(const rc_config *confs, long i, int lane)
{
    const float2 *p = reinterpret_cast<const float2 *>(confs + i);
    const float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    return make_matrix_warp(a.x, a.y, b.x, b.y, c.x, c.y, lane);
}

// Transform every robot vertex of one configuration with the lanes striding over the
// (padded) vertex table, reduce the AABB with warp shuffles.  All lanes return the box.
__device__ __forceinline__ void warp_aabb(const Mat34 &m, const float4 *__restrict__ s_rob_v,
                                          int n_v_pad, int lane, float box[6])
{
    float lx = FLT_MAX, ly = FLT_MAX, lz = FLT_MAX, hx = -FLT_MAX, hy = -FLT_MAX, hz = -FLT_MAX;
    for (int v = lane; v < n_v_pad; v += 32) {
        const float4 p = s_rob_v[v];
        float x, y, z;
        xform(m, p.x, p.y, p.z, x, y, z);
        lx = fminf(lx, x); hx = fmaxf(hx, x);
        ly = fminf(ly, y); hy = fmaxf(hy, y);
        lz = fminf(lz, z); hz = fmaxf(hz, z);
    }
    box[0] = warp_min(lx); box[1] = warp_min(ly); box[2] = warp_min(lz);
    box[3] = warp_max(hx); box[4] = warp_max(hy); box[5] = warp_max(hz);
}

struct CheckArgs {
    const float4 *rob_v;   // [n_rob_v_pad]
    const uint2 *rob_t;    // [n_rob_t]
    const float4 *obs_a, *obs_b, *obs_c;  // [n_obs_t]
    const float *obs_d;                   // [n_obs_t rounded up to 4]
    int n_rob_v_pad, n_rob_t, n_obs_t;
    rc_aabb obs_box;
    const rc_config *confs;
    long n;
    uint8_t *valid;
    uint32_t *bitmask;
    unsigned long long *counters;  // statistics
    unsigned long long *queue;     // work-queue counter of this launch
    int chunk;
    uint32_t *hit_counts;          // statistics kernels: hitting pairs per configuration, no early exit
};

__device__ __forceinline__ void publish(const CheckArgs &a, long i, bool valid, int lane)
{
    if (lane == 0) {
        if (a.valid) a.valid[i] = valid ? 1 : 0;
        if (a.bitmask && valid) atomicOr(a.bitmask + (i >> 5), 1u << (i & 31));
    }
}

// ------------------------------------------------------------------------------------------
// Exhaustive narrow phase of one configuration by one warp: every (robot triangle, obstacle
// triangle) pair the reference decides (narrow-phase.cu:586-677).  A lane owns two robot
// triangles per pass (j and j + 32), obstacle triangles are broadcast from shared memory with
// 128-bit loads, so one pass costs 3 LDS for 64 first plane tests (18 flop each).  Lanes past
// the end re-test the last triangle (a duplicate cannot change "some pair hits").  The later
// stages run at most once per lane and triangle between two warp-wide flushes, so the rare path
// executes with several lanes instead of one.  `load_tri(j)` yields the world-space triangle.
// COUNT: no early exit, every hitting pair is counted into `tails`' sibling `n_hits` (audit mode).
constexpr int kSweepQueue = 128;  // words per warp: two queues of 64 (pairs that passed the first / both plane tests)

template <bool COP, bool COUNT, class LoadTri>
__device__ __forceinline__ bool exhaustive_sweep(LoadTri load_tri, int n_rob_t,
                                                 const float4 *__restrict__ s_obs_a,
                                                 const float4 *__restrict__ s_obs_b,
                                                 const float4 *__restrict__ s_obs_c,
                                                 const float *__restrict__ s_obs_d, int n_obs_t,
                                                 int lane, unsigned long long &tails, uint32_t &n_hits,
                                                 uint32_t *__restrict__ q1)
{
    // The ~5 % of the pairs that pass the first plane test are not finished where they are found (a
    // handful of lanes at a time: the kernel ran 23 of 32 lanes) but queued per warp as (robot triangle,
    // obstacle triangle) and taken up 32 at a time: second plane test with full warps, its survivors
    // queued again, intersection line and intervals with full warps.  The robot triangle of a queued
    // pair is produced again by load_tri (bit-identical), so the queues outlive the pass they were
    // filled in.  Which pairs are decided, and how, is unchanged: the order does not enter the result.
    uint32_t *q2 = q1 + kSweepQueue / 2;
    int n1 = 0, n2 = 0;
    bool hit = false;
    const unsigned lt = (1u << lane) - 1u;

    // the rest of the test for up to 32 queued pairs that passed both plane tests
    auto finish = [&](int cnt) {
        n2 -= cnt;
        const bool act = lane < cnt;
        const uint32_t e = q2[n2 + (act ? lane : 0)];
        const int k = (int)(e & 0xffffu);
        const Tri r = load_tri((int)(e >> 16));
        const Plane pr = plane_of(r);
        const float4 oa = s_obs_a[k], ob = s_obs_b[k], oc = s_obs_c[k];
        const Tri o = {oa.x, oa.y, oa.z, ob.x, ob.y, ob.z, oc.x, oc.y, oc.z};
        const Plane po = {oa.w, ob.w, oc.w, s_obs_d[k]};
        const float dR0 = sdist(po, r.ax, r.ay, r.az), dR1 = sdist(po, r.bx, r.by, r.bz),
                    dR2 = sdist(po, r.cx, r.cy, r.cz);
        const float dO0 = sdist(pr, o.ax, o.ay, o.az), dO1 = sdist(pr, o.bx, o.by, o.bz),
                    dO2 = sdist(pr, o.cx, o.cy, o.cz);
        const int verdict = act ? pair_tail(r, pr, dR0, dR1, dR2, o, po, dO0, dO1, dO2) : kPairMiss;
        bool h = verdict == kPairHit;
        if constexpr (COP) {  // RC_COPLANAR_DECIDE
            if (verdict == kPairCoplanar && coplanar_overlap(r, o, pr)) h = true;
        }
        const unsigned mh = __ballot_sync(kFull, h);
        if (lane == 0) {
            tails += (unsigned long long)cnt;
            if (COUNT) n_hits += (uint32_t)__popc(mh);
        }
        if (mh) hit = true;
        __syncwarp();
    };
    // the second plane test for up to 32 queued pairs that passed the first
    auto second = [&](int cnt) {
        n1 -= cnt;
        const bool act = lane < cnt;
        const uint32_t e = q1[n1 + (act ? lane : 0)];
        const int k = (int)(e & 0xffffu);
        const Tri r = load_tri((int)(e >> 16));
        const float4 oa = s_obs_a[k], ob = s_obs_b[k], oc = s_obs_c[k];
        const Plane po = {oa.w, ob.w, oc.w, s_obs_d[k]};
        const bool pass = act && straddles(sdist(po, r.ax, r.ay, r.az), sdist(po, r.bx, r.by, r.bz),
                                           sdist(po, r.cx, r.cy, r.cz));
        const unsigned mp = __ballot_sync(kFull, pass);
        __syncwarp();  // everybody has read its entry before the queues move
        if (pass) q2[n2 + __popc(mp & lt)] = e;
        n2 += __popc(mp);
        __syncwarp();
        if (n2 >= 32) finish(32);
    };

    for (int jb = 0; jb < n_rob_t && (COUNT || !hit); jb += 64) {
        const int j0 = min(jb + lane, n_rob_t - 1), j1 = min(jb + 32 + lane, n_rob_t - 1);
        const Tri r0 = load_tri(j0);
        const Tri r1 = load_tri(j1);
        const Plane p0 = plane_of(r0), p1 = plane_of(r1);
        // Obstacle triangles in blocks of 32: the first plane test of every pair sets a bit; the hot
        // loop has no vote and no branch.
        for (int kb = 0; kb < n_obs_t; kb += 32) {
            uint32_t m0 = 0, m1 = 0;
            const int kn = min(32, n_obs_t - kb);
#pragma unroll 4
            for (int q = 0; q < kn; q++) {
                const int k = kb + q;
                const float4 oa = s_obs_a[k], ob = s_obs_b[k], oc = s_obs_c[k];  // broadcast
                const bool pass0 = straddles(sdist(p0, oa.x, oa.y, oa.z), sdist(p0, ob.x, ob.y, ob.z),
                                             sdist(p0, oc.x, oc.y, oc.z));
                const bool pass1 = straddles(sdist(p1, oa.x, oa.y, oa.z), sdist(p1, ob.x, ob.y, ob.z),
                                             sdist(p1, oc.x, oc.y, oc.z));
                m0 |= (uint32_t)pass0 << q;
                m1 |= (uint32_t)pass1 << q;
            }
            // lanes past the end re-test the last triangle: a duplicate cannot change "some pair hits",
            // but it must not be counted twice
            if (COUNT) {
                if (jb + lane >= n_rob_t) m0 = 0u;
                if (jb + 32 + lane >= n_rob_t) m1 = 0u;
            }
            // queue the survivors, one per lane and step
            while (__any_sync(kFull, (m0 | m1) != 0u)) {
                const bool has = (m0 | m1) != 0u, first = m0 != 0u;
                const uint32_t mm = first ? m0 : m1;
                const int q = has ? __ffs(mm) - 1 : 0;
                if (first) m0 &= m0 - 1; else m1 &= m1 - 1;
                const unsigned mp = __ballot_sync(kFull, has);
                if (has) q1[n1 + __popc(mp & lt)] = ((uint32_t)(first ? j0 : j1) << 16) | (uint32_t)(kb + q);
                n1 += __popc(mp);
                __syncwarp();
                if (n1 >= 32) {
                    second(32);
                    if (!COUNT && hit) break;
                }
            }
            if (!COUNT && hit) break;
        }
    }
    while (n1 > 0 && (COUNT || !hit)) second(min(n1, 32));
    while (n2 > 0 && (COUNT || !hit)) finish(min(n2, 32));
    return hit;
}

// ------------------------------------------------------------------------------------------
// Brute-force fused kernel.

template <bool STATS, bool COP>
__global__ void __launch_bounds__(kThreads, 2) check_brute_kernel(const CheckArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    float4 *s_rob_v = reinterpret_cast<float4 *>(smem_raw + 16);
    float4 *s_obs_a = s_rob_v + a.n_rob_v_pad;
    float4 *s_obs_b = s_obs_a + a.n_obs_t;
    float4 *s_obs_c = s_obs_b + a.n_obs_t;
    float *s_obs_d = reinterpret_cast<float *>(s_obs_c + a.n_obs_t);
    uint2 *s_rob_t = reinterpret_cast<uint2 *>(s_obs_d + pad4(a.n_obs_t));
    uint32_t *s_queue = reinterpret_cast<uint32_t *>(s_rob_t + a.n_rob_t + (a.n_rob_t & 1)) + (threadIdx.x >> 5) * kSweepQueue;

    // Stage the mesh tables with TMA bulk copies (one elected thread issues them, everybody
    // waits on the mbarrier's transaction count).
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bv = a.n_rob_v_pad * 16u, bo = a.n_obs_t * 16u,
                       bt = (a.n_rob_t + (a.n_rob_t & 1)) * 8u;
        const uint32_t bd = pad4(a.n_obs_t) * 4u;
        mbar_expect_tx(bar, bv + 3u * bo + bd + bt);
        tma_load_1d(s_rob_v, a.rob_v, bv, bar);
        tma_load_1d(s_obs_a, a.obs_a, bo, bar);
        tma_load_1d(s_obs_b, a.obs_b, bo, bar);
        tma_load_1d(s_obs_c, a.obs_c, bo, bar);
        tma_load_1d(s_obs_d, a.obs_d, bd, bar);
        tma_load_1d(s_rob_t, a.rob_t, bt, bar);
    }
    mbar_wait(bar, 0);

    const int lane = threadIdx.x & 31;
    unsigned long long st_broad = 0, st_valid = 0, st_invalid = 0, st_tail = 0;

    for (;;) {
        long base = 0;
        if (lane == 0) base = (long)atomicAdd(a.queue, (unsigned long long)a.chunk);
        base = __shfl_sync(kFull, base, 0);
        if (base >= a.n) break;
        const long end = min(base + (long)a.chunk, a.n);
        for (long i = base; i < end; i++) {
            const Mat34 m = load_matrix(a.confs, i);
            float box[6];
            warp_aabb(m, s_rob_v, a.n_rob_v_pad, lane, box);
            if (broad_valid(box[0], box[1], box[2], box[3], box[4], box[5], a.obs_box.x_min,
                            a.obs_box.y_min, a.obs_box.z_min, a.obs_box.x_max, a.obs_box.y_max,
                            a.obs_box.z_max)) {
                publish(a, i, true, lane);
                if (STATS) {
                    st_broad++;
                    if (a.hit_counts && lane == 0) a.hit_counts[i] = 0u;
                }
                continue;
            }
            auto load_tri = [&](int j) {
                const uint2 ti = s_rob_t[j];
                const float4 A = s_rob_v[ti.x & 0xffffu], B = s_rob_v[ti.x >> 16], C = s_rob_v[ti.y];
                Tri r;
                xform(m, A.x, A.y, A.z, r.ax, r.ay, r.az);
                xform(m, B.x, B.y, B.z, r.bx, r.by, r.bz);
                xform(m, C.x, C.y, C.z, r.cx, r.cy, r.cz);
                return r;
            };
            bool hit;
            uint32_t n_hits = 0;
            if (STATS && a.hit_counts) {
                hit = exhaustive_sweep<COP, true>(load_tri, a.n_rob_t, s_obs_a, s_obs_b, s_obs_c, s_obs_d,
                                                  a.n_obs_t, lane, st_tail, n_hits, s_queue);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) n_hits += __shfl_xor_sync(kFull, n_hits, o);
                if (lane == 0) a.hit_counts[i] = n_hits;
            } else {
                hit = exhaustive_sweep<COP, false>(load_tri, a.n_rob_t, s_obs_a, s_obs_b, s_obs_c, s_obs_d,
                                                   a.n_obs_t, lane, st_tail, n_hits, s_queue);
            }
            publish(a, i, !hit, lane);
            if (STATS) { if (hit) st_invalid++; else st_valid++; }
        }
    }
    if (STATS) {
        // per-config counters are warp-uniform (count on lane 0); tail counts are per lane
        st_tail = st_tail + __shfl_xor_sync(kFull, st_tail, 16);
        st_tail = st_tail + __shfl_xor_sync(kFull, st_tail, 8);
        st_tail = st_tail + __shfl_xor_sync(kFull, st_tail, 4);
        st_tail = st_tail + __shfl_xor_sync(kFull, st_tail, 2);
        st_tail = st_tail + __shfl_xor_sync(kFull, st_tail, 1);
        if (lane == 0) {
            atomicAdd(a.counters + CNT_BROAD_CLEARED, st_broad);
            atomicAdd(a.counters + CNT_NARROW_VALID, st_valid);
            atomicAdd(a.counters + CNT_NARROW_INVALID, st_invalid);
            atomicAdd(a.counters + CNT_TAIL_PAIRS, st_tail);
            atomicAdd(a.counters + CNT_TRI_PAIRS,
                      (st_valid + st_invalid) * 0ull);  // exhaustive: derived on the host
        }
    }
}

}  // namespace rc

#include "rc_gather.cuh"
#include "rc_culled.cuh"
#include "rc_spheres.cuh"
#include "rc_stages.cuh"
#include "rc_kat.cuh"

namespace rc {

static size_t brute_smem_bytes(const DeviceTables &t)
{
    return 16 + (size_t)t.n_rob_v_pad * 16 + (size_t)t.n_obs_t * 48 + (size_t)pad4(t.n_obs_t) * 4 +
           (size_t)(t.n_rob_t + (t.n_rob_t & 1)) * 8 + (size_t)kWarpsPerCta * kSweepQueue * 4;
}

cudaError_t launch_check(const rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid,
                         uint32_t *d_bitmask, int mode, int want_stats, cudaStream_t stream,
                         unsigned long long *queue, const CheckExtras *extras)
{
    if (!queue) queue = ctx->d_counters + CNT_QUEUE;
    const DeviceTables &t = ctx->t;
    cudaError_t e;
    if (n <= 0) return cudaSuccess;
    // the culled path keeps its own self-resetting counters
    if (mode == RC_MODE_BRUTE &&
        (e = cudaMemsetAsync(queue, 0, sizeof(unsigned long long), stream)) != cudaSuccess)
        return e;
    if (d_bitmask && mode == RC_MODE_BRUTE &&
        (e = cudaMemsetAsync(d_bitmask, 0, (size_t)((n + 31) / 32) * 4, stream)) != cudaSuccess)
        return e;

    if (mode == RC_MODE_BRUTE) {
        CheckArgs a;
        a.rob_v = t.rob_v; a.rob_t = t.rob_t;
        a.obs_a = t.obs_a; a.obs_b = t.obs_b; a.obs_c = t.obs_c; a.obs_d = t.obs_d;
        a.n_rob_v_pad = t.n_rob_v_pad; a.n_rob_t = t.n_rob_t; a.n_obs_t = t.n_obs_t;
        a.obs_box = t.obs_box;
        a.confs = d_confs; a.n = n; a.valid = d_valid; a.bitmask = d_bitmask;
        a.counters = ctx->d_counters; a.queue = queue;
        a.hit_counts = (want_stats && extras) ? extras->hit_counts : nullptr;
        const size_t smem = brute_smem_bytes(t);
        if (smem > (size_t)ctx->max_smem_optin) return cudaErrorInvalidConfiguration;
        const int ctas_per_sm = (2 * smem + 2048 <= (size_t)ctx->max_smem_optin) ? 2 : 1;
        const int grid = ctx->sm_count * ctas_per_sm;
        const long warps = (long)grid * kWarpsPerCta;
        a.chunk = (int)max(1L, min(8L, n / (warps * 16)));
        void (*kern)(const CheckArgs);
        if (ctx->coplanar_mode) kern = want_stats ? check_brute_kernel<true, true> : check_brute_kernel<false, true>;
        else kern = want_stats ? check_brute_kernel<true, false> : check_brute_kernel<false, false>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem)) != cudaSuccess)
            return e;
        kern<<<grid, kThreads, smem, stream>>>(a);
        return cudaGetLastError();
    }
    return launch_check_culled(ctx, d_confs, n, d_valid, d_bitmask, want_stats, stream, queue, extras);
}

}  // namespace rc
