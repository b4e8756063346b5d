// rc_culled.cuh -- the headline kernel: fused transform + AABB + broad phase + BVH-culled narrow
// phase, one warp per configuration, persistent grid fed by an atomic work queue.
//
// Flag semantics are the reference's: a configuration is invalid iff SOME (robot triangle,
// obstacle triangle) pair passes the Moller test of narrowPhaseKernel_coarse
// (narrow-phase.cu:586-677); the order in which pairs are visited does not matter, so the pairs
// may be enumerated by a tree instead of two nested loops as long as no hitting pair is skipped.
// A pair is skipped only when an ancestor pair of boxes is separated by more than the inflation
// margin on one of six axes (three world axes, three robot axes), i.e. when the two triangles are
// geometrically disjoint with clearance; every surviving pair is decided with the exact reference
// arithmetic (rc_device.cuh).  RC_MODE_BRUTE keeps the exhaustive sweep for audits.
//
// Step 1 (classify_kernel, one thread per configuration): matrix, bounding-sphere and rotated-box
// tests against the obstacle AABB; cleared configurations are written out, the rest go onto a work
// list.  Step 2 (check_culled_kernel), per warp and listed configuration:
//   A   all robot vertices transformed by the lanes, warp-shuffle AABB, exact broad-phase flag
//   B   dual-tree descent: a LIFO of (robot node, obstacle node) pairs in shared memory, 32 pairs
//       tested per step (one per lane), children appended with ballot/popc compaction
//   C   leaf pairs (<= 4 x 4 triangles) are batched eight at a time: lane = (leaf pair, robot
//       triangle), loop over the obstacle triangles of the leaf: the two plane tests
//   D   the few pairs that pass both plane tests are queued and finished (coplanarity,
//       intersection line, interval test) 32 at a time, so the long rare path runs with full
//       warps instead of a handful of lanes; warp vote early exit.
// Single-launch form (FEED kernels, RC_FEED): step 1 runs inside step 2 -- a few warps per CTA classify
// batches of 32 configurations from wherever they lie (rc_check_host: the caller's pinned buffer, read
// over PCIe) while the others already descend; see FeedArgs / FeedShared / feed_next below.
// Warps per CTA follow the size of the call (16 / 24 / 32, launch_check_culled): short calls are bound
// by their longest descents, long ones by instruction issue.
#pragma once

namespace rc {

constexpr int kCulledWarpsDefault = 32;  // warps per CTA (one CTA per SM); 16/20/24/28/32 instantiated
constexpr int kStackCap = 1024;  // node pairs per warp (fewer with more warps per CTA); overflow falls back to the exhaustive sweep
constexpr int kLeafQCap = 64;
constexpr int kTailQCap = 160;   // pairs that passed both plane tests, finished 32 at a time
constexpr int kStackCapGlobal = 1536;  // obstacle tables in global memory: deeper trees
// the deepest stack seen on alpha1.0 is ~200 pairs, so more warps per CTA can live with smaller stacks
__host__ __device__ constexpr int stack_cap_of(int warps, bool gobs)
{
    // (rounds pop a single pair near the capacity, see kNarrowPop, so small stacks do not overflow)
    return gobs ? (warps > 24 ? 640 : warps > 16 ? 1024 : kStackCapGlobal) : (warps > 28 ? 512 : warps > 24 ? 768 : kStackCap);
}
constexpr int kLeafTrisDev = 4;  // must match rc::kLeafTris of the host builder (rc_bvh.h)
constexpr int kClassifyBoxes = 32;  // vertex groups of the robot bounded in classify_kernel
constexpr int kMailSlots = 16;   // per CTA (capacity; a.mail_slots are used): pieces of a deep descent handed to warps without work
constexpr int kDonateMin = 64;   // node pairs on a warp's stack before it hands 32 of them out
constexpr uint32_t kPendHit = 0x80000000u;
constexpr uint32_t kWatchdogSpins = 1u << 24;  // polls of 256 ns each that a warp without work EXECUTES (time the
                                                // context spends switched out does not count): several seconds
constexpr int kSplitMax = 64;    // parts one configuration may be cut into when the list is short (rc_culled.cuh: split)

struct __align__(16) Mail {
    float4 m[3];         // [R | t] of the configuration
    uint32_t pairs[32];  // node pairs
    uint32_t state;      // 0 empty, 1 being filled, 2 full, 3 being emptied
    uint32_t owner;      // warp that owns the configuration (index into s_pend)
    uint32_t cnt;
    uint32_t kind;       // 0 node pairs, 1 leaf pairs, 2 triangle pairs for the tail
};

constexpr int kNarrowPop = 192;  // node pairs of head room below which a node round pops one pair instead of 32

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t *p) { return *(const volatile uint32_t *)p; }
__device__ __forceinline__ void st_volatile(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }

// ---- step 1 of rc_check: one THREAD per configuration ------------------------------------------
// Matrix (six libdevice sin/cos), bounding-sphere and rotated-bounding-box tests against the
// obstacle AABB.  A configuration either is provably cleared by the broad phase -- flag byte and a
// whole bitmask word per warp are written here, 24 B read + 1 B written per configuration, so this
// step is HBM-bound -- or goes, with its matrix, onto the work list of the warp-per-configuration
// kernel below.  Doing this per thread instead of per warp computes the trigonometry once per
// configuration instead of 32 times.
struct ClassifyArgs {
    const rc_config *confs;
    long n;
    uint8_t *valid;
    uint32_t *bitmask;
    rc_aabb obs_box;
    float sph_x, sph_y, sph_z, sph_r;
    float box_cx, box_cy, box_cz, box_ex, box_ey, box_ez;
    float4 bound[2 * 32];  // (centre, half extent) of the vertex groups, local frame: kernel parameter =
                           // constant bank, the unrolled loop reads them as instruction operands
    int n_bound;           // kClassifyBoxes, or 0 to skip them
    uint32_t *work_idx;
    float4 *work_mat;
    long work_cap;
    unsigned long long *work_count;
    unsigned long long *counters;
    int optimistic;        // listed configurations are marked valid here; the descent only ever writes "invalid"
};


// Single-launch form of rc_check ("feed"): the descent kernel classifies the configurations itself --
// a few warps per CTA take batches of 32, one configuration per lane, straight from wherever they
// lie (device memory, or the caller's page-locked host buffer read over PCIe) -- and hands the ones
// that need a descent to per-class lists that every warp draws from while later batches are still
// arriving.  List slot = (call number, configuration index): a slot is filled once its tag equals the
// call number of this launch, so the lists need no clearing between calls.
struct FeedArgs {
    int on;              // this launch classifies in the kernel (kernels instantiated with FEED only)
    int gs;              // split shift (parts per configuration = 1 << gs), decided on the host from n
    int cls_warps;       // warps per CTA that classify batches before they draw entries
    uint32_t gen;        // call number (never 0)
    uint2 *slots;        // [kWorkBins][work_cap]
    uint32_t *ctr;       // kFeedStride words apart (one L2 line each): [0..7] configurations listed per class
                         // [8..15] entries drawn per class  [16] next batch  [17] batches done
};
constexpr int kFeedStride = 32;
constexpr int kFeedCounters = 18;

struct CulledArgs {
    CheckArgs c;                  // rob_t / obs_* point at the BVH-ordered tables
    const float4 *rob_bvh, *obs_bvh;
    int n_rob_nodes, n_obs_nodes;
    float sph_x, sph_y, sph_z, sph_r;
    float box_cx, box_cy, box_cz, box_ex, box_ey, box_ez;  // robot local AABB (centre, half extent)
    int probe_stride;                                      // probe vertices: lane * stride
    long tail_begin;                                       // configurations left when chunks shrink to 1
    uint32_t seed_r[8], seed_o[8];                   // traversal starts from this cut of both trees
    int n_seed_r, n_seed_o;
    int tail_thresh, leaf_thresh;
    int stack_cap;                                         // <= compiled capacity (smaller only in tests)
    int share;                                             // 0: no hand-over between the warps of a CTA
    int mail_slots, donate_min, donate_leaf;               // hand-over tuning (<= kMailSlots; node / leaf pairs queued before giving)
    // work list written by classify_kernel: configurations the cheap tests could not clear
    const uint32_t *work_idx;                              // [kWorkBins][work_cap] configuration indices
    const float4 *work_mat;                                // three rows of [R | t] per CONFIGURATION
    long work_cap;
    uint32_t *trace;                 // rc_stats_trace, STATS kernels only
    unsigned long long *work_count;  // [0..7] entries per cost class  [8] next entry  [9] CTAs finished (self-resetting)
    // Short lists (planner-sized batches): a configuration is cut into up to split_max PARTS, each a
    // subset of the seed pairs of the deeper cut below, each an entry of its own for the warps of the
    // whole grid.  Parts never meet again, which is why the flags are written optimistically:
    // classify_kernel has marked every listed configuration valid, a part that finds a hit writes
    // "invalid" (optimistic != 0; the statistics kernels keep the one-warp-one-flag protocol).
    int optimistic, split_max, split_load;
    float tri_margin;                // triangle-box clearance of step C = the two inflated tree boxes
    uint32_t seed_r3[16], seed_o3[16];   // the cut at depth 4 of both trees: up to 256 seed pairs, dealt out to the parts
    int n_seed_r3, n_seed_o3;
    uint32_t *hit_counts;            // STATS kernels: hitting pairs per configuration, no early exit (rc_hit_counts)
    const float4 *rob_plane;         // RL1 kernels: plane of every robot triangle in the robot frame, leaf order
    GatherArgs g;                    // g.world > 1: the last CTA pushes the bitmask slot to the peers (rc_gather.cuh)
    FeedArgs feed;                   // FEED kernels
    ClassifyArgs k;                  // FEED kernels: what classify_kernel gets in the two-launch form
};

__device__ __forceinline__ bool boxes_overlap(const Mat34 &m, const float4 r0, const float4 r1,
                                              const float4 o0, const float4 o1)
{
    // robot box centre in the world frame, relative to the obstacle box centre
    const float dx = fmaf(m.r02, r0.z, fmaf(m.r01, r0.y, fmaf(m.r00, r0.x, m.tx))) - o0.x;
    const float dy = fmaf(m.r12, r0.z, fmaf(m.r11, r0.y, fmaf(m.r10, r0.x, m.ty))) - o0.y;
    const float dz = fmaf(m.r22, r0.z, fmaf(m.r21, r0.y, fmaf(m.r20, r0.x, m.tz))) - o0.z;
    // world axes: |d| <= |R| e_r + e_o
    const float wx = fmaf(fabsf(m.r02), r1.z, fmaf(fabsf(m.r01), r1.y, fabsf(m.r00) * r1.x)) + o1.x;
    const float wy = fmaf(fabsf(m.r12), r1.z, fmaf(fabsf(m.r11), r1.y, fabsf(m.r10) * r1.x)) + o1.y;
    const float wz = fmaf(fabsf(m.r22), r1.z, fmaf(fabsf(m.r21), r1.y, fabsf(m.r20) * r1.x)) + o1.z;
    bool sep = fabsf(dx) > wx || fabsf(dy) > wy || fabsf(dz) > wz;
    // robot axes: |R^T d| <= e_r + |R^T| e_o
    const float lx = fmaf(m.r20, dz, fmaf(m.r10, dy, m.r00 * dx));
    const float ly = fmaf(m.r21, dz, fmaf(m.r11, dy, m.r01 * dx));
    const float lz = fmaf(m.r22, dz, fmaf(m.r12, dy, m.r02 * dx));
    const float ex = fmaf(fabsf(m.r20), o1.z, fmaf(fabsf(m.r10), o1.y, fabsf(m.r00) * o1.x)) + r1.x;
    const float ey = fmaf(fabsf(m.r21), o1.z, fmaf(fabsf(m.r11), o1.y, fabsf(m.r01) * o1.x)) + r1.y;
    const float ez = fmaf(fabsf(m.r22), o1.z, fmaf(fabsf(m.r12), o1.y, fabsf(m.r02) * o1.x)) + r1.z;
    sep = sep || fabsf(lx) > ex || fabsf(ly) > ey || fabsf(lz) > ez;
    return !sep;
}

// Exhaustive sweep for one configuration, used only when the pair stack overflows (never
// observed); kept small on purpose: lane = robot triangle, obstacle triangles broadcast.
// Exact broad-phase flag of the reference (broad-phase-fused.cu:186-200), one warp.  If ANY
// exactly-transformed robot vertex lies inside the obstacle AABB, the two AABBs overlap on every
// axis (robot.min <= v <= obstacle.max, obstacle.min <= v <= robot.max), so the broad phase cannot
// clear the configuration: 32 probe vertices are tried first.  Otherwise the lanes transform all
// vertices and reduce the exact AABB.  Not inlined, and the matrix is read again from the work list
// (L2) instead of being passed: the call sits between two configurations and must not weigh on the
// register allocation of the descent loop (passing the matrix made ptxas keep it in local memory).
__device__ __noinline__ bool exact_broad_clears(const float4 *work_mat, long i, const float4 *s_rob_v,
                                                int n_rob_v_pad, int probe_stride, const rc_aabb obs_box,
                                                int lane)
{
    Mat34 m;
    {
        const float4 q0 = __ldcg(work_mat + 3 * i), q1 = __ldcg(work_mat + 3 * i + 1), q2 = __ldcg(work_mat + 3 * i + 2);
        m.r00 = q0.x; m.r01 = q0.y; m.r02 = q0.z; m.tx = q0.w;
        m.r10 = q1.x; m.r11 = q1.y; m.r12 = q1.z; m.ty = q1.w;
        m.r20 = q2.x; m.r21 = q2.y; m.r22 = q2.z; m.tz = q2.w;
    }
    const float4 p = s_rob_v[lane * probe_stride];
    float x, y, z;
    xform(m, p.x, p.y, p.z, x, y, z);
    const bool inside = obs_box.x_min <= x && x <= obs_box.x_max && obs_box.y_min <= y &&
                        y <= obs_box.y_max && obs_box.z_min <= z && z <= obs_box.z_max;
    if (__any_sync(kFull, inside)) return false;
    float box[6];
    warp_aabb(m, s_rob_v, n_rob_v_pad, lane, box);
    return broad_valid(box[0], box[1], box[2], box[3], box[4], box[5], obs_box.x_min, obs_box.y_min,
                       obs_box.z_min, obs_box.x_max, obs_box.y_max, obs_box.z_max);
}

template <bool COP>
__device__ __noinline__ bool narrow_exhaustive(const Mat34 m, const float4 *s_rob_v,
                                               const uint2 *s_rob_t, int n_rob_t,
                                               const float4 *s_obs_a, const float4 *s_obs_b,
                                               const float4 *s_obs_c, const float *s_obs_d,
                                               int n_obs_t, int lane)
{
    bool hit = false;
    for (int jb = 0; jb < n_rob_t && !hit; jb += 32) {
        const uint2 ti = s_rob_t[min(jb + lane, n_rob_t - 1)];
        const float4 A = s_rob_v[ti.x & 0xffffu], B = s_rob_v[ti.x >> 16], C = s_rob_v[ti.y];
        Tri r;
        xform(m, A.x, A.y, A.z, r.ax, r.ay, r.az);
        xform(m, B.x, B.y, B.z, r.bx, r.by, r.bz);
        xform(m, C.x, C.y, C.z, r.cx, r.cy, r.cz);
        const Plane pr = plane_of(r);
        for (int k = 0; k < n_obs_t; k++) {
            const float4 oa = s_obs_a[k], ob = s_obs_b[k], oc = s_obs_c[k];
            const Tri o = {oa.x, oa.y, oa.z, ob.x, ob.y, ob.z, oc.x, oc.y, oc.z};
            const Plane po = {oa.w, ob.w, oc.w, s_obs_d[k]};
            if (pair_hits<COP>(r, pr, o, po)) hit = true;
        }
        hit = __any_sync(kFull, hit);
    }
    return hit;
}

// The three conservative tests of one configuration against the obstacle AABB (bounding sphere, rotated
// local box, union of the rotated vertex-group boxes).  clear: provably cleared by the reference's broad
// phase; inside: vertex groups whose box meets the obstacle AABB = the cost class of the descent.
__device__ __forceinline__ void classify_one(const ClassifyArgs &a, const Mat34 &m, bool &clear, int &inside)
{
    const float sx = fmaf(m.r02, a.sph_z, fmaf(m.r01, a.sph_y, fmaf(m.r00, a.sph_x, m.tx)));
    const float sy = fmaf(m.r12, a.sph_z, fmaf(m.r11, a.sph_y, fmaf(m.r10, a.sph_x, m.ty)));
    const float sz = fmaf(m.r22, a.sph_z, fmaf(m.r21, a.sph_y, fmaf(m.r20, a.sph_x, m.tz)));
    // the robot lies inside its bounding sphere; if the sphere's box clears the obstacle AABB
    // with margin, so does the exact vertex AABB of the reference
    const float hx = a.sph_r + 0.01f + 4e-6f * (fabsf(sx) + a.sph_r);
    const float hy = a.sph_r + 0.01f + 4e-6f * (fabsf(sy) + a.sph_r);
    const float hz = a.sph_r + 0.01f + 4e-6f * (fabsf(sz) + a.sph_r);
    clear = sx - hx > a.obs_box.x_max || sx + hx < a.obs_box.x_min ||
            sy - hy > a.obs_box.y_max || sy + hy < a.obs_box.y_min ||
            sz - hz > a.obs_box.z_max || sz + hz < a.obs_box.z_min;
    if (!clear) {
        // ... and inside the world AABB of its rotated local bounding box
        const float cx = fmaf(m.r02, a.box_cz, fmaf(m.r01, a.box_cy, fmaf(m.r00, a.box_cx, m.tx)));
        const float cy = fmaf(m.r12, a.box_cz, fmaf(m.r11, a.box_cy, fmaf(m.r10, a.box_cx, m.ty)));
        const float cz = fmaf(m.r22, a.box_cz, fmaf(m.r21, a.box_cy, fmaf(m.r20, a.box_cx, m.tz)));
        const float ex = fmaf(fabsf(m.r02), a.box_ez, fmaf(fabsf(m.r01), a.box_ey, fabsf(m.r00) * a.box_ex));
        const float ey = fmaf(fabsf(m.r12), a.box_ez, fmaf(fabsf(m.r11), a.box_ey, fabsf(m.r10) * a.box_ex));
        const float ez = fmaf(fabsf(m.r22), a.box_ez, fmaf(fabsf(m.r21), a.box_ey, fabsf(m.r20) * a.box_ex));
        const float gx = ex + 0.01f + 4e-6f * (fabsf(cx) + ex);
        const float gy = ey + 0.01f + 4e-6f * (fabsf(cy) + ey);
        const float gz = ez + 0.01f + 4e-6f * (fabsf(cz) + ez);
        clear = cx - gx > a.obs_box.x_max || cx + gx < a.obs_box.x_min ||
                cy - gy > a.obs_box.y_max || cy + gy < a.obs_box.y_min ||
                cz - gz > a.obs_box.z_max || cz + gz < a.obs_box.z_min;
    }
    inside = 0;
    if (!clear && a.n_bound == kClassifyBoxes) {
        // ... and inside the union of the rotated boxes of its vertex groups: close to the exact
        // vertex AABB, so all but a thin band of the configurations the reference's broad phase
        // clears are decided here, by one thread, without transforming a single vertex.
        float lox = FLT_MAX, loy = FLT_MAX, loz = FLT_MAX, hix = -FLT_MAX, hiy = -FLT_MAX, hiz = -FLT_MAX;
        const float ax = fabsf(m.r00), ay = fabsf(m.r01), az = fabsf(m.r02);
        const float bx = fabsf(m.r10), by = fabsf(m.r11), bz = fabsf(m.r12);
        const float cx = fabsf(m.r20), cy = fabsf(m.r21), cz = fabsf(m.r22);
#pragma unroll
        for (int r = 0; r < kClassifyBoxes; r++) {
            const float4 bc = a.bound[2 * r], be = a.bound[2 * r + 1];
            const float px = fmaf(m.r02, bc.z, fmaf(m.r01, bc.y, fmaf(m.r00, bc.x, m.tx)));
            const float py = fmaf(m.r12, bc.z, fmaf(m.r11, bc.y, fmaf(m.r10, bc.x, m.ty)));
            const float pz = fmaf(m.r22, bc.z, fmaf(m.r21, bc.y, fmaf(m.r20, bc.x, m.tz)));
            const float ex = fmaf(az, be.z, fmaf(ay, be.y, ax * be.x));
            const float ey = fmaf(bz, be.z, fmaf(by, be.y, bx * be.x));
            const float ez = fmaf(cz, be.z, fmaf(cy, be.y, cx * be.x));
            const float l0 = px - ex, h0 = px + ex, l1 = py - ey, h1 = py + ey, l2 = pz - ez, h2 = pz + ez;
            lox = fminf(lox, l0); hix = fmaxf(hix, h0);
            loy = fminf(loy, l1); hiy = fmaxf(hiy, h1);
            loz = fminf(loz, l2); hiz = fmaxf(hiz, h2);
            inside += (l0 <= a.obs_box.x_max && h0 >= a.obs_box.x_min && l1 <= a.obs_box.y_max &&
                       h1 >= a.obs_box.y_min && l2 <= a.obs_box.z_max && h2 >= a.obs_box.z_min) ? 1 : 0;
        }
        const float gx = 0.01f + 4e-6f * fmaxf(fabsf(lox), fabsf(hix));
        const float gy = 0.01f + 4e-6f * fmaxf(fabsf(loy), fabsf(hiy));
        const float gz = 0.01f + 4e-6f * fmaxf(fabsf(loz), fabsf(hiz));
        clear = lox - gx > a.obs_box.x_max || hix + gx < a.obs_box.x_min ||
                loy - gy > a.obs_box.y_max || hiy + gy < a.obs_box.y_min ||
                loz - gz > a.obs_box.z_max || hiz + gz < a.obs_box.z_min;
    }
}

// cost class from the share of the robot inside the obstacle AABB (edges ~ 1/64, 3/64, 7/64, 13/64,
// 21/64, 31/64, 43/64 of the vertex groups)
__device__ __forceinline__ int cost_class(int inside)
{
    const int f = inside * (64 / kClassifyBoxes);
    return (f >= 1) + (f >= 3) + (f >= 7) + (f >= 13) + (f >= 21) + (f >= 31) + (f >= 43);
}

template <bool STATS>
__global__ void __launch_bounds__(64) classify_kernel(const __grid_constant__ ClassifyArgs a)
{
    // the descent kernel is launched with programmatic stream serialisation: let its CTAs start
    // staging the mesh tables while this grid is still running (they wait before reading the list)
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const long n32 = (a.n + 31) & ~31L;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += (long)gridDim.x * blockDim.x) {
        const bool in = i < a.n;
        const Mat34 m = load_matrix(a.confs, in ? i : a.n - 1);
        bool clear;
        int inside;  // vertex groups whose box meets the obstacle AABB: the cost class of the descent
        classify_one(a, m, clear, inside);
        clear = clear && in;
        const bool work = in && !clear;
        const unsigned mc = __ballot_sync(kFull, clear), mw = __ballot_sync(kFull, work);
        const bool one = clear || (work && a.optimistic);
        if (one && a.valid) a.valid[i] = 1;
        // the warp covers exactly one bitmask word: write it (the second kernel ORs its bits in, or,
        // in optimistic mode, clears the bits of the colliding configurations)
        if (lane == 0 && a.bitmask) a.bitmask[i >> 5] = a.optimistic ? (mc | mw) : mc;
        if (mw) {
            // one list per cost class, warp-aggregated append
            const int bin = cost_class(inside);
            const unsigned peers = __match_any_sync(kFull, work ? bin : kWorkBins);
            if (work) {
                const int leader = __ffs(peers) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(a.work_count + bin, (unsigned long long)__popc(peers));
                base = __shfl_sync(peers, base, leader);
                a.work_idx[(size_t)bin * a.work_cap + base + __popc(peers & lt)] = (uint32_t)i;
                a.work_mat[3 * i + 0] = make_float4(m.r00, m.r01, m.r02, m.tx);
                a.work_mat[3 * i + 1] = make_float4(m.r10, m.r11, m.r12, m.ty);
                a.work_mat[3 * i + 2] = make_float4(m.r20, m.r21, m.r22, m.tz);
            }
        }
        if (STATS && lane == 0 && mc) atomicAdd(a.counters + CNT_SPHERE_CLEARED, (unsigned long long)__popc(mc));
    }
}

// ---- step 2: one WARP per remaining configuration ------------------------------------------------
// GOBS = obstacle tables (BVH + triangles) stay in global memory (L2-resident) instead of being
// staged into shared memory: meshes of any size (BASELINE config 5, the subdivided obstacle).
// Pair encodings then use 11 bits for the robot side and 21 bits for the obstacle side.
__device__ __forceinline__ void publish_invalid(const CheckArgs &a, long i, int lane)
{
    if (lane == 0) {
        if (a.valid) a.valid[i] = 0;
        if (a.bitmask) atomicAnd(a.bitmask + (i >> 5), ~(1u << (i & 31)));
    }
}

// FEED kernels: what a CTA keeps in shared memory.  Entries are drawn from the global lists a batch at
// a time by whichever warp holds the lock and found the ring low (feed_refill); every warp of the CTA
// pops its next entry from the ring.  Thousands of warps drawing one entry each from the same global
// counters would queue up on their L2 lines -- and the dearest-first order needs the lengths of all
// lists, which only one poller per CTA reads.
constexpr int kFeedRing = 64;
struct FeedShared {
    uint2 ring[kFeedRing];          // (configuration, part)
    uint32_t head, tail;            // popped (compare-and-swap by any warp), pushed (lock holder)
    uint32_t lock, done, bex, first;  // refill lock; lists final and empty; no batch left to classify; warps that took their own first batch
    uint32_t p_lo[kWorkBins], p_hi[kWorkBins];  // tickets drawn beyond the end of a list (a race with other CTAs while
                                                // the classifiers were still listing): entries once the list reaches them
};

// Called by the warp that holds f->lock.  Reads the global counters, takes up to 32 tickets of the
// dearest class that has entries (its own overshoot tickets first), waits for their slots, pushes the
// entries onto the ring.  Sets f->done when every batch is classified and nothing is left for this CTA.
__device__ __noinline__ void feed_refill(const CulledArgs &a, FeedShared *f, int lane)
{
    const uint32_t n_batches = (uint32_t)((a.k.n + 31) >> 5);
    const uint32_t gs = (uint32_t)a.feed.gs;
    const int cq = kWorkBins - 1 - lane;  // lane 0 looks at the dearest class
    uint32_t h = 0, r = 0, lo = 0, hi = 0, d = 0;
    for (int pass = 0; pass < 2; pass++) {
        if (lane < kWorkBins) {
            r = ld_volatile(a.feed.ctr + cq * kFeedStride) << gs;
            lo = f->p_lo[cq]; hi = f->p_hi[cq];
            h = (lo == hi) ? ld_volatile(a.feed.ctr + (kWorkBins + cq) * kFeedStride) : r;
        }
        if (pass || __any_sync(kFull, (lo < hi && lo < r) || (lo == hi && h < r))) break;
        // nothing to draw: are the lists final?  Batches done first, then the lengths once more -- if
        // every batch was done before they were read, they are final
        if (lane == 0) d = ld_volatile(a.feed.ctr + 17 * kFeedStride);
        d = __shfl_sync(kFull, d, 0);
        __threadfence();
    }
    uint32_t space = 0;
    if (lane == 0) space = kFeedRing - (f->tail - ld_volatile(&f->head));
    space = __shfl_sync(kFull, space, 0);
    uint32_t want = min(32u, space), t0 = 0, k = 0;
    int cl = -1;
    const unsigned pv = __ballot_sync(kFull, lo < hi && lo < r);
    const unsigned any_left = __ballot_sync(kFull, lo == hi && h < r);
    if (pv && want > 0) {
        const int l = __ffs(pv) - 1;
        cl = kWorkBins - 1 - l;
        t0 = __shfl_sync(kFull, lo, l);
        k = min(want, min(__shfl_sync(kFull, hi, l), __shfl_sync(kFull, r, l)) - t0);
        if (lane == 0) f->p_lo[cl] = t0 + k;
    } else if (want > 0) {
        const uint32_t mine = (lo == hi && h < r) ? r - h : 0u;
        uint32_t total = mine;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
        total = __shfl_sync(kFull, total, 0);
        const unsigned av = __ballot_sync(kFull, mine > 0u);
        if (av) {
            const int l = __ffs(av) - 1;
            cl = kWorkBins - 1 - l;
            const uint32_t rr = __shfl_sync(kFull, r, l);
            // near the end the CTAs take their share of what is left, so that none sits on 32 entries
            want = min(want, max(1u, total / gridDim.x));
            k = min(want, __shfl_sync(kFull, mine, l));
            if (lane == 0) t0 = atomicAdd(a.feed.ctr + (kWorkBins + cl) * kFeedStride, k);
            t0 = __shfl_sync(kFull, t0, 0);
            const uint32_t t1 = t0 + k, good = min(t1, max(t0, rr));
            if (good < t1 && lane == 0) { f->p_lo[cl] = good; f->p_hi[cl] = t1; }
            k = good - t0;
        }
    }
    if (k > 0) {
        uint32_t tail = 0;
        if (lane == 0) tail = f->tail;
        tail = __shfl_sync(kFull, tail, 0);
        if ((uint32_t)lane < k) {
            const uint32_t t = t0 + lane;
            const volatile unsigned long long *sl = reinterpret_cast<const volatile unsigned long long *>(
                a.feed.slots + (size_t)cl * a.work_cap + (t >> gs));
            unsigned long long v = 0;
            // the slot was written right after its position was reserved: a short wait at most
            for (uint32_t q = 0; q < kWatchdogSpins; q++) {
                v = *sl;
                if ((uint32_t)v == a.feed.gen) break;
            }
            if ((uint32_t)v != a.feed.gen) {  // cannot happen; counted.  The entry is pushed as a no-op (configuration n)
                atomicAdd(a.c.counters + CNT_WATCHDOG, 1ULL);
                atomicOr(a.c.counters + CNT_WATCHDOG_WHERE, 2ULL);
                v = (unsigned long long)0xffffffffu << 32;
            }
            f->ring[(tail + lane) & (kFeedRing - 1)] = make_uint2((uint32_t)(v >> 32), t & ((1u << gs) - 1u));
        }
        __threadfence();  // the matrices of these entries were written before their slots
        __syncwarp();
        if (lane == 0) st_volatile(&f->tail, tail + k);
    } else if (d == n_batches && !pv && !any_left) {
        // the lists are final and hold nothing this CTA could still draw (a ticket range it keeps is void)
        if (lane == 0) st_volatile(&f->done, 1u);
    }
}

// FEED kernels: one batch of 32 configurations classified by one warp of the descent kernel (what a
// warp of classify_kernel does in the two-launch form), flags written optimistically, the listed ones
// appended to the per-class lists.  Not inlined: it runs between two descents and must not weigh on
// the register allocation of the descent loop.
__device__ __noinline__ void classify_batch(const CulledArgs &a, uint32_t b, int lane)
{
    const ClassifyArgs &k = a.k;
    const unsigned lt = (1u << lane) - 1u;
    const long i = (long)b * 32 + lane;
    const bool in = i < k.n;
    const Mat34 m = load_matrix(k.confs, in ? i : k.n - 1);
    bool clear;
    int inside;
    classify_one(k, m, clear, inside);
    clear = clear && in;
    const bool work = in && !clear;
    const unsigned mc = __ballot_sync(kFull, clear), mw = __ballot_sync(kFull, work);
    if (in && k.valid) k.valid[i] = 1;  // optimistic: the descent only ever writes "invalid"
    if (lane == 0 && k.bitmask) k.bitmask[i >> 5] = mc | mw;
    if (mw) {
        const int bin = cost_class(inside);
        const unsigned peers = __match_any_sync(kFull, work ? bin : kWorkBins);
        uint32_t pos = 0;
        if (work) {
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(a.feed.ctr + bin * kFeedStride, (uint32_t)__popc(peers));
            pos = __shfl_sync(peers, base, leader) + __popc(peers & lt);
            k.work_mat[3 * i + 0] = make_float4(m.r00, m.r01, m.r02, m.tx);
            k.work_mat[3 * i + 1] = make_float4(m.r10, m.r11, m.r12, m.ty);
            k.work_mat[3 * i + 2] = make_float4(m.r20, m.r21, m.r22, m.tz);
        }
        // flag, bitmask word and matrix first, then the slot that lets a descent start on them
        __threadfence();
        __syncwarp();
        if (work) {
            const unsigned long long v = (unsigned long long)a.feed.gen | ((unsigned long long)(uint32_t)i << 32);
            *(volatile unsigned long long *)(a.feed.slots + (size_t)bin * k.work_cap + pos) = v;
        }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(a.feed.ctr + 17 * kFeedStride, 1u);  // the batch is done: its entries are visible before this count
}

// FEED kernels: the next entry of this warp -- (part << 32) | configuration -- or ~0 when the lists are
// final and empty.  Classifies batches (the first cls_warps warps of the CTA, while any are left), else
// pops an entry from the ring of the CTA, refilling it from the global lists when it runs low.
__device__ __noinline__ unsigned long long feed_next(const CulledArgs &a, FeedShared *s_feed, int warp, int lane)
{
    const uint32_t n_batches = (uint32_t)((a.k.n + 31) >> 5);
#ifdef RC_TIMELINE
    // diagnostic build: per warp [4] ns inside feed_next  [5] iterations that found the ring empty  [6] ns inside feed_refill  [7] refills
    uint32_t *tlf = a.trace ? a.trace + 8 * (size_t)(blockIdx.x * (blockDim.x >> 5) + warp) : nullptr;
    unsigned long long tl_in;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl_in));
#define TLF_LEAVE() do { if (tlf && lane == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tlf[4] += (uint32_t)(t_ - tl_in); } } while (0)
#else
#define TLF_LEAVE() do { } while (0)
#endif
    for (uint32_t spins = 0;; spins++) {
        // what other warps publish is read by one lane and broadcast: lanes reading the same
        // word at different moments could take different branches of this loop
        uint32_t fl = 0;
        if (lane == 0) fl = (ld_volatile(&s_feed->bex) ? 1u : 0u) | (ld_volatile(&s_feed->done) ? 2u : 0u);
        fl = __shfl_sync(kFull, fl, 0);
        const uint32_t done = fl & 2u;  // read BEFORE the pop: nothing is pushed once it is set
        if (warp < a.feed.cls_warps) {
            // the first batch of every classifying warp is its own (thousands of atomicAdds on one word at
            // the start of the kernel would queue up for microseconds), the rest are drawn
            const uint32_t first = gridDim.x * (uint32_t)a.feed.cls_warps;
            uint32_t b = 0xffffffffu;
            if (lane == 0) {
                if (!((ld_volatile(&s_feed->first) >> warp) & 1u)) {
                    atomicOr(&s_feed->first, 1u << warp);
                    b = blockIdx.x * (uint32_t)a.feed.cls_warps + (uint32_t)warp;
                    if (first >= n_batches) st_volatile(&s_feed->bex, 1u);  // nothing will be left to draw
                } else if (!(fl & 1u)) {
                    b = first + atomicAdd(a.feed.ctr + 16 * kFeedStride, 1u);
                    if (b >= n_batches) st_volatile(&s_feed->bex, 1u);
                }
            }
            b = __shfl_sync(kFull, b, 0);
            if (b < n_batches) {
                classify_batch(a, b, lane);
                continue;
            }
        }
        // pop: the entry is read before the compare-and-swap that claims it (a refill never
        // overwrites a position that has not been claimed)
        uint32_t st = 0, left = 0;
        unsigned long long e = 0;
        if (lane == 0) {
            for (;;) {
                const uint32_t hd = ld_volatile(&s_feed->head), tl = ld_volatile(&s_feed->tail);
                if (hd >= tl) break;
                e = *reinterpret_cast<volatile unsigned long long *>(&s_feed->ring[hd & (kFeedRing - 1)]);
                if (atomicCAS(&s_feed->head, hd, hd + 1u) == hd) {
                    st = 1u; left = tl - hd - 1u;
                    break;
                }
            }
        }
        st = __shfl_sync(kFull, st, 0);
        const bool low = !st || __shfl_sync(kFull, left, 0) < 24u;
        if (low && !done) {
            uint32_t got = 0;
            if (lane == 0) got = atomicCAS(&s_feed->lock, 0u, 1u) == 0u ? 1u : 0u;
            got = __shfl_sync(kFull, got, 0);
            if (got) {
#ifdef RC_TIMELINE
                unsigned long long t0_;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0_));
#endif
                feed_refill(a, s_feed, lane);
#ifdef RC_TIMELINE
                if (tlf && lane == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tlf[6] += (uint32_t)(t_ - t0_); tlf[7] += 1u; }
#endif
                __syncwarp();
                if (lane == 0) st_volatile(&s_feed->lock, 0u);
            }
        }
        if (st) {
            e = __shfl_sync(kFull, e, 0);
            if ((uint32_t)e == 0xffffffffu) continue;  // a slot that never arrived (counted in feed_refill)
            TLF_LEAVE();
            return e;
        }
        if (done || spins > kWatchdogSpins) {
            if (!done && lane == 0) { atomicAdd(a.c.counters + CNT_WATCHDOG, 1ULL); atomicOr(a.c.counters + CNT_WATCHDOG_WHERE, 4ULL); }
            TLF_LEAVE();
            return ~0ULL;
        }
#ifdef RC_TIMELINE
        if (tlf && lane == 0 && !st) tlf[5] += 1u;
#endif
        __nanosleep(spins < 8 ? 100 : 400);
    }
}

// RL1 (with GOBS) = every leaf of the robot tree holds ONE triangle and the kernel knows its plane in
// the robot frame: for an obstacle much finer than the robot (the subdivided obstacle: a robot triangle
// is a thousand times the area of an obstacle triangle) the box of a four-triangle robot leaf is a poor
// bound, while the slab of ONE robot triangle's plane culls every obstacle box that stays on one side
// of it -- exactly the pairs the reference's first plane test rejects.
template <bool STATS, int WARPS, bool GOBS, bool COP, bool RL1, bool FEED = false>
__global__ void __launch_bounds__(WARPS * 32, 1) check_culled_kernel(const __grid_constant__ CulledArgs a)
{
    static_assert(!RL1 || GOBS, "one-triangle robot leaves are used with the obstacle in global memory");
    static_assert(!FEED || !STATS, "the statistics kernels keep the two-launch form");
    constexpr int SH = GOBS ? 21 : 16;
    constexpr uint32_t LOW = (1u << SH) - 1u;
    constexpr int kStack = stack_cap_of(WARPS, GOBS);
    constexpr int kWarpWork = kStack + kLeafQCap + kTailQCap;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const CheckArgs &c = a.c;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    float4 *s_rob_v = reinterpret_cast<float4 *>(smem_raw + 16);
    float4 *s_rob_bvh = s_rob_v + c.n_rob_v_pad;
    float4 *s_obs_bvh = s_rob_bvh + 2 * a.n_rob_nodes;
    float4 *s_obs_a = s_obs_bvh + (GOBS ? 0 : 2 * a.n_obs_nodes);
    float4 *s_obs_b = s_obs_a + (GOBS ? 0 : c.n_obs_t);
    float4 *s_obs_c = s_obs_b + (GOBS ? 0 : c.n_obs_t);
    float *s_obs_d = reinterpret_cast<float *>(s_obs_c + (GOBS ? 0 : c.n_obs_t));
    uint2 *s_rob_t = reinterpret_cast<uint2 *>(s_obs_d + (GOBS ? 0 : pad4(c.n_obs_t)));
    float4 *s_rob_plane = reinterpret_cast<float4 *>(s_rob_t + c.n_rob_t + (c.n_rob_t & 1));  // RL1: (n, s) per triangle, robot frame
    uint32_t *s_work = reinterpret_cast<uint32_t *>(s_rob_plane + (RL1 ? c.n_rob_t : 0));
    Mail *s_mail = reinterpret_cast<Mail *>(s_work + WARPS * kWarpWork);
    uint32_t *s_pend = reinterpret_cast<uint32_t *>(s_mail + kMailSlots);
    uint32_t *s_idle = s_pend + WARPS;
    uint32_t *s_misc = s_idle + 1;   // [0] split shift  [1] this CTA finished last  [2..5] spare
    uint32_t *s_bin = s_misc + 6;    // [kWorkBins + 1] first list position of each cost class, dearest class first
    uint32_t *s_wait = s_bin + kWorkBins + 1;  // [2 * WARPS] configuration a warp has finished but for its pieces: index, hit
    uint32_t *s_fetch = s_wait + 2 * WARPS;    // [2 * WARPS] entries of the list a warp has drawn: next, end
    uint32_t *s_seed = s_fetch + 2 * WARPS;    // [64] seed pairs of a whole descent  [256] seed pairs dealt out to split descents
    static_assert(WARPS % 2 == 0, "an even number of words precedes s_feed, which holds 8-byte entries");
    FeedShared *s_feed = reinterpret_cast<FeedShared *>(s_seed + 320);  // FEED kernels
    // obstacle accessors: shared memory, or the read-only global path
    auto obs_node = [&](int nb, int half) -> float4 {
        if constexpr (GOBS) return __ldg(a.obs_bvh + 2 * nb + half);
        else return s_obs_bvh[2 * nb + half];
    };
    auto obs_A = [&](int k) -> float4 { if constexpr (GOBS) return __ldg(c.obs_a + k); else return s_obs_a[k]; };
    auto obs_B = [&](int k) -> float4 { if constexpr (GOBS) return __ldg(c.obs_b + k); else return s_obs_b[k]; };
    auto obs_C = [&](int k) -> float4 { if constexpr (GOBS) return __ldg(c.obs_c + k); else return s_obs_c[k]; };
    auto obs_D = [&](int k) -> float { if constexpr (GOBS) return __ldg(c.obs_d + k); else return s_obs_d[k]; };

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (threadIdx.x < kMailSlots) s_mail[threadIdx.x].state = 0u;
    if (FEED && threadIdx.x < sizeof(FeedShared) / 4) reinterpret_cast<uint32_t *>(s_feed)[threadIdx.x] = 0u;
    {
        // the seed pairs of every descent, encoded once per CTA (indexing the argument arrays per
        // configuration cost ~80 instructions each time)
        const int q = (int)threadIdx.x;
        if (q < a.n_seed_r * a.n_seed_o) s_seed[q] = (a.seed_r[q / a.n_seed_o] << SH) | a.seed_o[q % a.n_seed_o];
        if (q < a.n_seed_r3 * a.n_seed_o3) s_seed[64 + q] = (a.seed_r3[q / a.n_seed_o3] << SH) | a.seed_o3[q % a.n_seed_o3];
    }
    if (threadIdx.x <= WARPS + 6) s_pend[threadIdx.x] = 0u;  // s_pend, s_idle, s_misc[0..5]
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bv = c.n_rob_v_pad * 16u, bo = c.n_obs_t * 16u,
                       bt = (c.n_rob_t + (c.n_rob_t & 1)) * 8u;
        const uint32_t brn = a.n_rob_nodes * 32u, bon = a.n_obs_nodes * 32u;
        const uint32_t bd = pad4(c.n_obs_t) * 4u;
        mbar_expect_tx(bar, bv + bt + brn + (GOBS ? 0u : 3u * bo + bd + bon) + (RL1 ? c.n_rob_t * 16u : 0u));
        tma_load_1d(s_rob_v, c.rob_v, bv, bar);
        tma_load_1d(s_rob_bvh, a.rob_bvh, brn, bar);
        tma_load_1d(s_rob_t, c.rob_t, bt, bar);
        if (RL1) tma_load_1d(s_rob_plane, a.rob_plane, c.n_rob_t * 16u, bar);
        if (!GOBS) {
            tma_load_1d(s_obs_bvh, a.obs_bvh, bon, bar);
            tma_load_1d(s_obs_a, c.obs_a, bo, bar);
            tma_load_1d(s_obs_b, c.obs_b, bo, bar);
            tma_load_1d(s_obs_c, c.obs_c, bo, bar);
            tma_load_1d(s_obs_d, c.obs_d, bd, bar);
        }
    }
    // everything above reads only the static tables; the work list, the flags and the bitmask
    // words come from classify_kernel: wait for that grid to complete (programmatic dependency)
    if constexpr (!FEED) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (FEED && threadIdx.x == 0) s_misc[0] = (uint32_t)a.feed.gs;
    if (!FEED && threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int o = 0; o < kWorkBins; o++) {
            s_bin[o] = acc;
            acc += (uint32_t)__ldcg(a.work_count + (kWorkBins - 1 - o));
        }
        s_bin[kWorkBins] = acc;
        // split: as many parts per configuration (a power of two) as keep the entries at or below
        // split_load per warp of the grid -- 1 for the large sets the kernel is tuned for
        uint32_t gs = 0;
        if (!STATS)
            while ((2u << gs) <= (uint32_t)a.split_max &&
                   ((unsigned long long)acc << (gs + 1)) <= (unsigned long long)a.split_load * gridDim.x * WARPS)
                gs++;
        s_misc[0] = gs;
    }
    __syncthreads();
    // FEED kernels classify while the tables are still on their way: they wait before every descent
    if constexpr (!FEED) mbar_wait(bar, 0);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t *stack = s_work + warp * kWarpWork;
    uint32_t *leafq = stack + kStack;
    uint32_t *tailq = leafq + kLeafQCap;
    unsigned long long st_broad = 0, st_valid = 0, st_invalid = 0, st_nodes = 0,
                       st_leaves = 0, st_tris = 0, st_over = 0, st_tails = 0, st_nr = 0, st_lr = 0, st_tr = 0, st_shared = 0,
                       st_pre = 0, st_pl1 = 0;
    const bool counting = STATS && a.hit_counts != nullptr;  // audit: every hitting pair, no early exit
    // OPT: the production protocol.  Flags are optimistic (classify_kernel marked every listed
    // configuration valid, whoever finds a hit overwrites), so a piece of a descent that one warp
    // hands to another never reports back: one-way mail through the slots of the CTA.  The statistics
    // kernels (!OPT) keep the rendezvous protocol -- the owner of a configuration collects the verdict
    // of its pieces and writes the flag once -- because they count per configuration.
    // (Mail between CTAs through slots in global memory was built and measured in round 2: the end of
    // a launch is not where the time goes -- 100,000 configurations 0.2835 ms with it, 0.2796 without,
    // a million 2.474 / 2.471 -- so it was taken out again; profiles/ab_r02.txt.)
    constexpr bool OPT = !STATS;

    // Guided self-scheduling over the work list: chunks while plenty of work is left, single
    // entries near the end so that no warp is left holding a long chunk.  The cost of one
    // configuration is heavy-tailed (median 9 k cycles, 1 % above 120 k, maximum 280 k = 40 % of
    // the whole kernel), so once the list is dry the warps SHARE what is left: a warp whose
    // pair stack is deep hands its 32 oldest entries (the largest subtrees) to a mail slot, warps
    // without work of their own pick slots up and run the same descent on them.  No wait of either
    // protocol can give up on work, and no descent is ever abandoned: a piece that nobody picks up
    // is worked off by the warp that posted it once it has nothing else to do.
    // The bookkeeping of the list (what this warp has drawn, the list length, the chunk size) lives in
    // shared memory and is looked at once per configuration: the descent loop is at the 64-register
    // limit of a 1,024-thread CTA, and every value kept alive across it costs spills.
    if (lane == 0) { s_fetch[2 * warp] = 0u; s_fetch[2 * warp + 1] = 0u; }
    __syncwarp();
    bool dry = false, waiting = false;
#ifdef RC_TIMELINE
    // diagnostic build only (profiles/tools/timeline.py): per warp, when it started, found the list dry, left
    uint32_t *tl = (!STATS && a.trace) ? a.trace + 8 * (size_t)(blockIdx.x * WARPS + warp) : nullptr;
    uint32_t tl_entries = 0, tl_longest = 0, tl_rounds = 0, tl_given = 0, tl_got = 0, tl_t0 = 0, tl_nr = 0, tl_lr = 0, tl_tr = 0;
    auto tl_now = []() { unsigned long long ns; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns)); return (uint32_t)ns; };
    if (tl && lane == 0) tl[0] = tl_now();
#endif
    for (;;) {
        uint32_t w = 0;
        if constexpr (FEED) {
            if (!dry) {
                // ---- single-launch form: feed_next classifies batches, pops the ring of the CTA and
                // refills it (not inlined: the descent loop keeps the register allocation of the
                // two-launch kernel)
                const unsigned long long nx = feed_next(a, s_feed, warp, lane);
                if (nx == ~0ULL) {
                    dry = true;
                    if (lane == 0) atomicAdd(s_idle, 1u);
                } else {
                    w = (uint32_t)nx;                                          // configuration index
                    if (lane == 0) s_fetch[2 * warp] = (uint32_t)(nx >> 32);   // part
                    __syncwarp();
                }
            }
        }
        if (!FEED && !dry && !waiting) {
            const uint32_t gs0 = s_misc[0];
            const uint32_t n_work = s_bin[kWorkBins] << gs0;  // entries: (configuration, part)
            uint32_t w_next = s_fetch[2 * warp], w_end = s_fetch[2 * warp + 1];
            if (w_next >= w_end) {
                const uint32_t chunk = max(1u, min(16u, n_work / (gridDim.x * WARPS * 32u)));
                const uint32_t take = ((unsigned long long)w_end + a.tail_begin < n_work) ? chunk : 1u;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.work_count + kWorkBins, (unsigned long long)take);
                base = __shfl_sync(kFull, base, 0);
                if (base >= n_work) {
                    dry = true;
                    if (lane == 0) atomicAdd(s_idle, 1u);
#ifdef RC_TIMELINE
                    if (tl && lane == 0) tl[1] = tl_now();
#endif
                } else {
                    w_next = (uint32_t)base;
                    w_end = (uint32_t)min(base + take, (unsigned long long)n_work);
                }
            }
            if (!dry) {
                w = w_next;
                __syncwarp();
                if (lane == 0) { s_fetch[2 * warp] = w_next + 1u; s_fetch[2 * warp + 1] = w_end; }
            }
        }
        Mat34 m;
        uint32_t i = 0;
        int owner = -1;       // !OPT, >= 0: this warp works on a piece of that warp's configuration
        bool sharing = false;  // !OPT: pieces of the current configuration are (or were) out
        int sp = 0, lq = 0, tq = 0;
        long long t_begin = 0;
        uint32_t tr_nodes = 0, tr_leaves = 0, tr_tails = 0, tr_maxsp = 0;  // rounds of the owner itself
        uint32_t n_hits = 0;   // counting mode
        if (STATS) t_begin = clock64();
        auto trace = [&]() {
            if (STATS && a.trace && lane == 0) {
                unsigned long long ns;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
                a.trace[4 * (size_t)i] = (uint32_t)(clock64() - t_begin);
                a.trace[4 * (size_t)i + 1] = (uint32_t)ns;
                a.trace[4 * (size_t)i + 2] = (tr_nodes << 16) | min(tr_leaves, 0xffffu);
                a.trace[4 * (size_t)i + 3] = (tr_tails << 16) | min(tr_maxsp, 0xffffu);
            }
        };
        if (OPT && dry) {
            // ---- one-way mail: poll the slots of the CTA; leave when every warp of the CTA has found
            // the list dry (a warp that still works on a piece picks up whatever it posts itself)
            if (!a.share) break;  // nobody hands anything over: no point in polling beside the warps still at work
            int k;
            for (uint32_t spins = 0;; spins++) {
                k = -1;
                if (lane == 0) {
                    for (int q = 0; q < a.mail_slots; q++)
                        if (ld_volatile(&s_mail[q].state) == 2u && atomicCAS(&s_mail[q].state, 2u, 3u) == 2u) {
                            k = q;
                            break;
                        }
                    if (k < 0) {
                        if (ld_volatile(s_idle) == (uint32_t)WARPS) k = -2;
                        else if (spins > kWatchdogSpins) {
                            // cannot happen (every descent is finite).  Leaving costs parallelism only: a
                            // piece nobody picks up is worked off by the warp that posted it.
                            atomicAdd(c.counters + CNT_WATCHDOG, 1ULL);
                            atomicOr(c.counters + CNT_WATCHDOG_WHERE, 1ULL);
                            k = -2;
                        }
                    }
                }
                k = __shfl_sync(kFull, k, 0);
                if (k != -1) break;
                __nanosleep(200);
            }
            if (k == -2) break;
            __threadfence_block();
            Mail &b = s_mail[k];
            i = b.owner;  // the configuration the piece belongs to
            const int cnt = (int)b.cnt, kind = (int)b.kind;
            const uint32_t e = b.pairs[lane];
            m.r00 = b.m[0].x; m.r01 = b.m[0].y; m.r02 = b.m[0].z; m.tx = b.m[0].w;
            m.r10 = b.m[1].x; m.r11 = b.m[1].y; m.r12 = b.m[1].z; m.ty = b.m[1].w;
            m.r20 = b.m[2].x; m.r21 = b.m[2].y; m.r22 = b.m[2].z; m.tz = b.m[2].w;
            __syncwarp();
            __threadfence_block();
            if (lane == 0) st_volatile(&s_mail[k].state, 0u);
            {
                uint32_t *dst = kind == 0 ? stack : kind == 1 ? leafq : tailq;
                if (lane < cnt) dst[lane] = e;
                if (kind == 0) sp = cnt; else if (kind == 1) lq = cnt; else tq = cnt;
            }
            __syncwarp();
        } else if (!OPT && (dry || waiting)) {
            // ---- rendezvous mail (statistics kernels): serve the slots of the CTA.  A warp that
            // waits for the pieces of its configuration leaves when they are all back; a warp that
            // found the list dry leaves when every warp of the CTA is in that state.
            int k;
            for (uint32_t spins = 0;; spins++) {
                k = -1;
                if (lane == 0) {
                    if (waiting && (ld_volatile(&s_pend[warp]) & ~kPendHit) == 0u) {
                        k = -3;
                    } else {
                        for (int q = 0; q < a.mail_slots; q++)
                            if (ld_volatile(&s_mail[q].state) == 2u && atomicCAS(&s_mail[q].state, 2u, 3u) == 2u) {
                                k = q;
                                break;
                            }
                        if (k < 0 && !waiting) {
                            if (ld_volatile(s_idle) == (uint32_t)WARPS) k = -2;
                            else if (spins > kWatchdogSpins) {
                                // cannot happen (every descent is finite).  Leaving costs parallelism only:
                                // whoever still has pieces out works them off itself.
                                atomicAdd(c.counters + CNT_WATCHDOG, 1ULL);
                                atomicOr(c.counters + CNT_WATCHDOG_WHERE, 1ULL);
                                k = -2;
                            }
                        }
                    }
                }
                k = __shfl_sync(kFull, k, 0);
                if (k != -1) break;
                __nanosleep(waiting ? 64 : 256);
            }
            if (k == -2) break;
            if (k == -3) {
                // all pieces are back: take over their verdict, publish the configuration
                uint32_t v = 0, wi = 0, wh = 0;
                if (lane == 0) {
                    v = ld_volatile(&s_pend[warp]);
                    st_volatile(&s_pend[warp], 0u);
                    wi = s_wait[2 * warp];
                    wh = s_wait[2 * warp + 1];
                }
                v = __shfl_sync(kFull, v, 0);
                wi = __shfl_sync(kFull, wi, 0);
                wh = __shfl_sync(kFull, wh, 0);
                waiting = false;
                const bool h = wh != 0u || (v & kPendHit) != 0u;
                publish(c, (long)wi, !h, lane);
                if (h) st_invalid++; else st_valid++;
                continue;
            }
            __threadfence_block();
            Mail &b = s_mail[k];
            owner = (int)b.owner;
            sharing = true;
            {
                const int cnt = (int)b.cnt, kind = (int)b.kind;
                uint32_t *dst = kind == 0 ? stack : kind == 1 ? leafq : tailq;
                if (lane < cnt) dst[lane] = b.pairs[lane];
                if (kind == 0) sp = cnt; else if (kind == 1) lq = cnt; else tq = cnt;
            }
            m.r00 = b.m[0].x; m.r01 = b.m[0].y; m.r02 = b.m[0].z; m.tx = b.m[0].w;
            m.r10 = b.m[1].x; m.r11 = b.m[1].y; m.r12 = b.m[1].z; m.ty = b.m[1].w;
            m.r20 = b.m[2].x; m.r21 = b.m[2].y; m.r22 = b.m[2].z; m.tz = b.m[2].w;
            __syncwarp();
            __threadfence_block();
            if (lane == 0) st_volatile(&s_mail[k].state, 0u);
            __syncwarp();
        } else {
            const uint32_t gs = s_misc[0];
            uint32_t part;
            if constexpr (FEED) {
                i = w;
                part = s_fetch[2 * warp];
            } else {
                const uint32_t wl = w >> gs;
                part = w & ((1u << gs) - 1u);
                // list position -> (cost class, position in its list): dearest class first
                const uint32_t o = __popc(__ballot_sync(kFull, lane >= 1 && lane < kWorkBins && wl >= s_bin[lane & (kWorkBins - 1)]));
                i = __ldcg(a.work_idx + (size_t)(kWorkBins - 1 - o) * a.work_cap + (wl - s_bin[o]));
            }
            {
                const float4 q0 = __ldcg(a.work_mat + 3 * (size_t)i), q1 = __ldcg(a.work_mat + 3 * (size_t)i + 1),
                             q2 = __ldcg(a.work_mat + 3 * (size_t)i + 2);
                m.r00 = q0.x; m.r01 = q0.y; m.r02 = q0.z; m.tx = q0.w;
                m.r10 = q1.x; m.r11 = q1.y; m.r12 = q1.z; m.ty = q1.w;
                m.r20 = q2.x; m.r21 = q2.y; m.r22 = q2.z; m.tz = q2.w;
            }

            // A: exact broad-phase flag (exact_broad_clears).  The flag of a configuration is "the
            // broad phase clears it, or the narrow phase finds no hit", so the exact robot AABB only
            // matters when the descent DOES report a hit: the production build runs the descent
            // first and settles the broad phase afterwards, for colliding configurations only (a
            // configuration the broad phase clears usually ends after the 16 seed pairs).  The
            // statistics build keeps the reference's order, because it reports the two kinds of
            // valid configurations separately.
            if constexpr (STATS) {
                if (exact_broad_clears(a.work_mat, i, s_rob_v, c.n_rob_v_pad, a.probe_stride, c.obs_box, lane)) {
                    publish(c, i, true, lane);
                    if (counting && lane == 0) a.hit_counts[i] = 0u;
                    st_broad++;
                    trace();
                    continue;
                }
            }
            // the descent starts from a cut through both trees instead of (root, root): the top
            // levels almost always overlap here and would cost steps with a handful of lanes.
            // Split: part p of 2^gs takes every 2^gs-th pair of the deeper cut (depth 4: up to 256 pairs,
            // so that the parts of one configuration carry similar shares of its descent).
            if (gs == 0) {
                sp = a.n_seed_r * a.n_seed_o;
                for (int q = lane; q < sp; q += 32) stack[q] = s_seed[q];
            } else {
                const int ns = a.n_seed_r3 * a.n_seed_o3;
                sp = (int)part < ns ? (ns - (int)part + (1 << gs) - 1) >> gs : 0;
                for (int j = lane; j < sp; j += 32) stack[j] = s_seed[64 + (int)part + (j << gs)];
            }
            __syncwarp();
        }
        const int pend_slot = owner >= 0 ? owner : warp;
        if constexpr (FEED) mbar_wait(bar, 0);  // the tables have landed (immediate after the first time)
#ifdef RC_TIMELINE
        tl_entries++;
        if (dry) tl_got++;
        tl_t0 = tl_now(); tl_nr = tl_lr = tl_tr = 0;
#endif

        // B + C + D
        bool hit = false, overflow = false;
        while (sp > 0 || lq > 0 || tq > 0) {
            // ---- sharing (only once some warp of the CTA has run out of configurations): hand out
            // a tail round's worth of triangle pairs (the longest rounds), else a leaf round's
            // worth of leaf pairs, else up to half of the node pairs, oldest = largest subtrees
            // a warp always keeps work for itself, so what it receives is never passed on untouched
            // (a piece bouncing between idle warps would never be decided)
            // What other warps publish is agreed on by the whole warp (vote / broadcast: two lanes
            // reading the same word at different moments could otherwise take different branches of
            // this loop).  s_idle only ever grows and stays zero until the first warp of the CTA finds
            // the list dry, so the common case is one shared-memory read and one vote; what there is
            // to give is only worked out once somebody could take it.  With idle warps around, anything
            // beyond the round this warp is about to run is worth giving: a round takes a lone warp
            // ~0.8 us, a hand-over a fraction of that.
            int kind = -1;
            uint32_t seen = 0;
            if (a.share || (!OPT && sharing)) {
                if (__any_sync(kFull, ld_volatile(s_idle) != 0u)) {
                    if (a.share)
                        kind = (tq >= a.tail_thresh && (sp > 0 || lq > 0 || tq > 32)) ? 2
                               : (lq >= a.donate_leaf && (sp > 0 || lq >= 16)) ? 1
                               : sp >= a.donate_min ? 0 : -1;
                    if (lane == 0) {
                        seen = 1u;
                        if (!OPT && sharing && (ld_volatile(&s_pend[pend_slot]) & kPendHit) != 0u) seen = 3u;
                    }
                    seen = __shfl_sync(kFull, seen, 0);
                }
            }
            if (!OPT && (seen & 2u)) {
                hit = true;  // another piece of this configuration has found a hit
                break;
            }
            if (seen && kind >= 0) {
                int k = -1;
                if (lane == 0)
                    for (int q = 0; q < a.mail_slots; q++)
                        if (ld_volatile(&s_mail[q].state) == 0u && atomicCAS(&s_mail[q].state, 0u, 1u) == 0u) {
                            k = q;
                            break;
                        }
                k = __shfl_sync(kFull, k, 0);
                if (k >= 0) {
                    const int cnt = kind == 2 ? min(tq, 32) : kind == 1 ? 8 : min(32, sp >> 1);  // node pairs: at most half
                    const int src = lane < cnt ? lane : 0;
                    uint32_t out;
                    if (kind == 2) {
                        tq -= cnt;
                        out = tailq[tq + src];
                    } else if (kind == 1) {
                        lq -= cnt;
                        out = leafq[lq + src];
                    } else {
                        out = stack[src];
                        const uint32_t top = stack[sp - cnt + src];  // the newest take their place
                        __syncwarp();
                        if (lane < cnt) stack[lane] = top;
                        sp -= cnt;
                    }
                    {
                        Mail &b = s_mail[k];
                        if (lane < cnt) b.pairs[lane] = out;
                        if (lane == 0) {
                            b.m[0] = make_float4(m.r00, m.r01, m.r02, m.tx);
                            b.m[1] = make_float4(m.r10, m.r11, m.r12, m.ty);
                            b.m[2] = make_float4(m.r20, m.r21, m.r22, m.tz);
                            b.owner = OPT ? i : (uint32_t)pend_slot;
                            b.cnt = (uint32_t)cnt;
                            b.kind = (uint32_t)kind;
                            if (!OPT) atomicAdd(&s_pend[pend_slot], 1u);
                        }
                        if (STATS) st_shared++;
#ifdef RC_TIMELINE
                        tl_given++;
#endif
                        sharing = true;
                        __syncwarp();
                        __threadfence_block();
                        if (lane == 0) st_volatile(&s_mail[k].state, 2u);
                    }
                    __syncwarp();
                    continue;
                }
            }
            if (tq >= a.tail_thresh || (sp == 0 && lq == 0)) {
                // ---- D: finish up to 32 queued pairs, one per lane
                const int cnt = min(tq, 32);
                tq -= cnt;
                const bool act = lane < cnt;
                const uint32_t e = tailq[tq + (act ? lane : 0)];
                const uint2 ti = s_rob_t[e >> SH];
                const int k = (int)(e & LOW);
                const float4 A = s_rob_v[ti.x & 0xffffu], B = s_rob_v[ti.x >> 16], C = s_rob_v[ti.y];
                Tri r;
                xform(m, A.x, A.y, A.z, r.ax, r.ay, r.az);
                xform(m, B.x, B.y, B.z, r.bx, r.by, r.bz);
                xform(m, C.x, C.y, C.z, r.cx, r.cy, r.cz);
                const Plane pr = plane_of(r);
                const float4 oa = obs_A(k), ob = obs_B(k), oc = obs_C(k);
                const Tri o = {oa.x, oa.y, oa.z, ob.x, ob.y, ob.z, oc.x, oc.y, oc.z};
                const Plane po = {oa.w, ob.w, oc.w, obs_D(k)};
                const float dO0 = sdist(pr, o.ax, o.ay, o.az), dO1 = sdist(pr, o.bx, o.by, o.bz),
                            dO2 = sdist(pr, o.cx, o.cy, o.cz);
                const float dR0 = sdist(po, r.ax, r.ay, r.az), dR1 = sdist(po, r.bx, r.by, r.bz),
                            dR2 = sdist(po, r.cx, r.cy, r.cz);
                const int verdict = act ? pair_tail(r, pr, dR0, dR1, dR2, o, po, dO0, dO1, dO2) : kPairMiss;
                bool h = verdict == kPairHit;
                if constexpr (COP) {  // RC_COPLANAR_DECIDE: the undecided pairs are decided in the common plane
                    if (verdict == kPairCoplanar) h = coplanar_overlap(r, o, pr);
                }
                if (STATS) { st_tails += cnt; st_tr++; tr_tails++; }
#ifdef RC_TIMELINE
                tl_tr++;
#endif
                const unsigned mh = __ballot_sync(kFull, h);
                if (mh) {
                    hit = true;
                    if (STATS && counting) n_hits += __popc(mh);
                    else break;
                }
                __syncwarp();
                continue;
            }
            if (RL1 && (lq >= a.leaf_thresh || sp == 0)) {
                // ---- C (one-triangle robot leaves): up to eight leaf pairs, lane = (slot, obstacle
                // triangle of the leaf); the four lanes of a slot share the robot triangle
                const int cnt = min(lq, 8);
                lq -= cnt;
                const int slot = lane >> 2, q = lane & 3;
                bool act = slot < cnt;
                const uint32_t e = leafq[lq + (act ? slot : 0)];
                const int na = (int)(e >> SH), nb = (int)(e & LOW);
                const uint32_t j = __float_as_uint(s_rob_bvh[2 * na].w) & 0x7fffffffu;
                const uint32_t ofirst = __float_as_uint(obs_node(nb, 0).w) & 0x7fffffffu;
                const int ocount = (int)__float_as_uint(obs_node(nb, 1).w);
                act = act && q < ocount;
                const int k = (int)ofirst + (act ? q : 0);
                const float4 oa = obs_A(k), ob = obs_B(k), oc = obs_C(k);
                const uint2 ti = s_rob_t[j];
                const float4 A = s_rob_v[ti.x & 0xffffu], B = s_rob_v[ti.x >> 16], C = s_rob_v[ti.y];
                Tri r;
                xform(m, A.x, A.y, A.z, r.ax, r.ay, r.az);
                xform(m, B.x, B.y, B.z, r.bx, r.by, r.bz);
                xform(m, C.x, C.y, C.z, r.cx, r.cy, r.cz);
                const Plane pr = plane_of(r);
                bool pass = act &&
                            fminf(fminf(oa.x, ob.x), oc.x) <= fmaxf(fmaxf(r.ax, r.bx), r.cx) + a.tri_margin &&
                            fmaxf(fmaxf(oa.x, ob.x), oc.x) >= fminf(fminf(r.ax, r.bx), r.cx) - a.tri_margin &&
                            fminf(fminf(oa.y, ob.y), oc.y) <= fmaxf(fmaxf(r.ay, r.by), r.cy) + a.tri_margin &&
                            fmaxf(fmaxf(oa.y, ob.y), oc.y) >= fminf(fminf(r.ay, r.by), r.cy) - a.tri_margin &&
                            fminf(fminf(oa.z, ob.z), oc.z) <= fmaxf(fmaxf(r.az, r.bz), r.cz) + a.tri_margin &&
                            fmaxf(fmaxf(oa.z, ob.z), oc.z) >= fminf(fminf(r.az, r.bz), r.cz) - a.tri_margin;
                if (STATS) st_pre += __popc(__ballot_sync(kFull, pass));
                if (pass)
                    pass = straddles(sdist(pr, oa.x, oa.y, oa.z), sdist(pr, ob.x, ob.y, ob.z),
                                     sdist(pr, oc.x, oc.y, oc.z));
                if (STATS) st_pl1 += __popc(__ballot_sync(kFull, pass));
                if (pass) {
                    const Plane po = {oa.w, ob.w, oc.w, obs_D(k)};
                    pass = straddles(sdist(po, r.ax, r.ay, r.az), sdist(po, r.bx, r.by, r.bz),
                                     sdist(po, r.cx, r.cy, r.cz));
                }
                const unsigned mp = __ballot_sync(kFull, pass);
                if (pass) tailq[tq + __popc(mp & lt)] = (j << SH) | (uint32_t)k;
                tq += __popc(mp);
                if (STATS) { st_leaves += cnt; st_tris += (act ? 1 : 0); st_lr++; tr_leaves++; }
                __syncwarp();
                continue;
            }
            if (!RL1 && (lq >= a.leaf_thresh || sp == 0)) {
                // ---- C: up to eight leaf pairs, lane = (slot, robot triangle of the leaf)
                const int cnt = min(lq, 8);
                lq -= cnt;
                const int slot = lane >> 2, ri = lane & 3;
                bool act = slot < cnt;
                const uint32_t e = leafq[lq + (act ? slot : 0)];
                const int na = (int)(e >> SH), nb = (int)(e & LOW);
                const uint32_t rfirst = __float_as_uint(s_rob_bvh[2 * na].w) & 0x7fffffffu;
                const int rcount = (int)__float_as_uint(s_rob_bvh[2 * na + 1].w);
                const uint32_t ofirst = __float_as_uint(obs_node(nb, 0).w) & 0x7fffffffu;
                const int ocount = (int)__float_as_uint(obs_node(nb, 1).w);
                act = act && ri < rcount;
                const uint32_t j = rfirst + (act ? ri : 0);
                const uint2 ti = s_rob_t[j];
                const float4 A = s_rob_v[ti.x & 0xffffu], B = s_rob_v[ti.x >> 16], C = s_rob_v[ti.y];
                Tri r;
                xform(m, A.x, A.y, A.z, r.ax, r.ay, r.az);
                xform(m, B.x, B.y, B.z, r.bx, r.by, r.bz);
                xform(m, C.x, C.y, C.z, r.cx, r.cy, r.cz);
                const Plane pr = plane_of(r);
                // AABB of the robot triangle, inflated like the tree boxes: a pair whose
                // triangle boxes are disjoint with clearance is skipped like a culled node pair
                // (it cuts the queued pairs four-fold: most pairs of two overlapping leaves are
                // themselves apart)
                const float rlx = fminf(fminf(r.ax, r.bx), r.cx) - a.tri_margin,
                            rly = fminf(fminf(r.ay, r.by), r.cy) - a.tri_margin,
                            rlz = fminf(fminf(r.az, r.bz), r.cz) - a.tri_margin,
                            rhx = fmaxf(fmaxf(r.ax, r.bx), r.cx) + a.tri_margin,
                            rhy = fmaxf(fmaxf(r.ay, r.by), r.cy) + a.tri_margin,
                            rhz = fmaxf(fmaxf(r.az, r.bz), r.cz) + a.tri_margin;
#pragma unroll
                for (int q = 0; q < kLeafTrisDev; q++) {
                    const int k = ofirst + min(q, ocount - 1);
                    const float4 oa = obs_A(k), ob = obs_B(k), oc = obs_C(k);
                    bool pass = act && q < ocount &&
                                fminf(fminf(oa.x, ob.x), oc.x) <= rhx && fmaxf(fmaxf(oa.x, ob.x), oc.x) >= rlx &&
                                fminf(fminf(oa.y, ob.y), oc.y) <= rhy && fmaxf(fmaxf(oa.y, ob.y), oc.y) >= rly &&
                                fminf(fminf(oa.z, ob.z), oc.z) <= rhz && fmaxf(fmaxf(oa.z, ob.z), oc.z) >= rlz;
                    if (STATS) st_pre += __popc(__ballot_sync(kFull, pass));
                    if (pass)
                        pass = straddles(sdist(pr, oa.x, oa.y, oa.z), sdist(pr, ob.x, ob.y, ob.z),
                                         sdist(pr, oc.x, oc.y, oc.z));
                    if (STATS) st_pl1 += __popc(__ballot_sync(kFull, pass));
                    if (pass) {
                        const Plane po = {oa.w, ob.w, oc.w, obs_D(k)};
                        pass = straddles(sdist(po, r.ax, r.ay, r.az), sdist(po, r.bx, r.by, r.bz),
                                         sdist(po, r.cx, r.cy, r.cz));
                    }
                    const unsigned mp = __ballot_sync(kFull, pass);
                    if (pass) tailq[tq + __popc(mp & lt)] = (j << SH) | (uint32_t)k;
                    tq += __popc(mp);
                }
                if (STATS) { st_leaves += cnt; st_tris += (act ? ocount : 0); st_lr++; tr_leaves++; }
#ifdef RC_TIMELINE
                tl_lr++;
#endif
                __syncwarp();
                continue;
            }
            // ---- B: pop up to 32 node pairs, test, push children.  A round that pops 32 pairs can
            // push 128; close to the capacity of the stack the rounds pop ONE pair, i.e. turn into a
            // depth-first walk whose stack grows by at most three per tree level, so the capacity is
            // not reached in practice whatever the configuration (the exhaustive fallback stays).
            const int cnt = (sp + kNarrowPop <= a.stack_cap || a.stack_cap < 2 * kNarrowPop) ? min(sp, 32) : 1;
            sp -= cnt;
            const bool act = lane < cnt;
            const uint32_t e = stack[sp + (act ? lane : 0)];
            const int na = (int)(e >> SH), nb = (int)(e & LOW);
            const float4 r0 = s_rob_bvh[2 * na], r1 = s_rob_bvh[2 * na + 1];
            const float4 o0 = obs_node(nb, 0), o1 = obs_node(nb, 1);
            bool ov = act && boxes_overlap(m, r0, r1, o0, o1);
            const uint32_t ra = __float_as_uint(r0.w), rb = __float_as_uint(r1.w);
            const uint32_t oa = __float_as_uint(o0.w), ob = __float_as_uint(o1.w);
            const bool la = (ra & 0x80000000u) != 0, lb = (oa & 0x80000000u) != 0;
            if constexpr (RL1) {
                if (ov && la) {
                    // the robot node is ONE triangle: an obstacle box that stays on one side of its plane
                    // holds no triangle whose vertices straddle that plane (the reference's first test).
                    // World normal n = R n_local, n . p = s_local + n . t on the triangle; the box is cut
                    // by the plane iff |n . c - s| <= |n| . e.  e carries the box margin (the clearance);
                    // the slack covers the rounding of these few operations at scene scale.
                    const float4 pl = s_rob_plane[ra & 0x7fffffffu];
                    const float nx = fmaf(m.r02, pl.z, fmaf(m.r01, pl.y, m.r00 * pl.x));
                    const float ny = fmaf(m.r12, pl.z, fmaf(m.r11, pl.y, m.r10 * pl.x));
                    const float nz = fmaf(m.r22, pl.z, fmaf(m.r21, pl.y, m.r20 * pl.x));
                    const float sw = fmaf(nz, m.tz, fmaf(ny, m.ty, fmaf(nx, m.tx, pl.w)));
                    const float dist = fmaf(nz, o0.z, fmaf(ny, o0.y, nx * o0.x)) - sw;
                    const float rad = fmaf(fabsf(nz), o1.z, fmaf(fabsf(ny), o1.y, fabsf(nx) * o1.x));
                    const float mag = fmaf(fabsf(nz), fabsf(o0.z), fmaf(fabsf(ny), fabsf(o0.y), fabsf(nx * o0.x))) + fabsf(sw);
                    if (fabsf(dist) > fmaf(1e-5f, mag, rad)) ov = false;
                }
            }
            if (STATS) { st_nodes += cnt; st_nr++; tr_nodes++; tr_maxsp = max(tr_maxsp, (uint32_t)(sp + cnt)); }
#ifdef RC_TIMELINE
            tl_nr++;
#endif
            __syncwarp();  // everybody has read its entry before the slots are reused
            const bool ll = ov && la && lb;
            const unsigned mll = __ballot_sync(kFull, ll);
            if (ll) leafq[lq + __popc(mll & lt)] = e;
            lq += __popc(mll);
            const bool two = ov && (la != lb), four = ov && !la && !lb;
            const unsigned m2 = __ballot_sync(kFull, two), m4 = __ballot_sync(kFull, four);
            const int total = 2 * __popc(m2) + 4 * __popc(m4);
            if (sp + total > a.stack_cap) { overflow = true; break; }
            const int off = sp + 2 * __popc(m2 & lt) + 4 * __popc(m4 & lt);
            if (two) {
                if (la) {  // robot node is a leaf: descend the obstacle node
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
            if (STATS) st_over++;
            if constexpr (GOBS)
                hit = narrow_exhaustive<COP>(m, s_rob_v, s_rob_t, c.n_rob_t, c.obs_a, c.obs_b, c.obs_c,
                                        c.obs_d, c.n_obs_t, lane);
            else
                hit = narrow_exhaustive<COP>(m, s_rob_v, s_rob_t, c.n_rob_t, s_obs_a, s_obs_b, s_obs_c,
                                        s_obs_d, c.n_obs_t, lane);
            if (STATS && counting) n_hits = 0xffffffffu;  // the fallback does not count: the audit reports it
        }
        __syncwarp();
#ifdef RC_TIMELINE
        { const uint32_t d = tl_now() - tl_t0; if (d > tl_longest) { tl_longest = d; tl_rounds = (min(tl_nr, 1023u) << 20) | (min(tl_lr, 1023u) << 10) | min(tl_tr, 1023u); } }
#endif
        if constexpr (OPT) {
            // whoever finds a hit writes "invalid" -- unless the reference's broad phase, which comes
            // first, clears the configuration: then it is valid whatever the narrow phase says
            if (hit && !exact_broad_clears(a.work_mat, i, s_rob_v, c.n_rob_v_pad, a.probe_stride, c.obs_box, lane))
                publish_invalid(c, i, lane);
            continue;
        }
        if (owner >= 0) {
            // a piece: report to the owner of the configuration
            if (lane == 0) {
                if (hit) atomicOr(&s_pend[owner], kPendHit);
                __threadfence_block();
                atomicSub(&s_pend[owner], 1u);
            }
            continue;
        }
        trace();
        if (STATS && counting && lane == 0) a.hit_counts[i] = n_hits;
        if (sharing) {
            // pieces of this configuration went out: it is published when they are all back; until
            // then this warp serves mail slots (top of the loop)
            if (lane == 0) {
                s_wait[2 * warp] = i;
                s_wait[2 * warp + 1] = hit ? 1u : 0u;
                if (hit) atomicOr(&s_pend[warp], kPendHit);  // the pieces still out may stop
            }
            __syncwarp();
            waiting = true;
            continue;
        }
        publish(c, i, !hit, lane);
        if (hit) st_invalid++; else st_valid++;
    }
#ifdef RC_TIMELINE
    if (tl && lane == 0) { tl[2] = tl_now(); tl[3] = tl_entries; if (!FEED) { tl[4] = tl_longest; tl[5] = tl_rounds; tl[6] = tl_given; tl[7] = tl_got; } }
#endif
    if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) st_tris += __shfl_xor_sync(kFull, st_tris, o);
        if (lane == 0) {
            atomicAdd(c.counters + CNT_BROAD_CLEARED, st_broad);
            atomicAdd(c.counters + CNT_NARROW_VALID, st_valid);
            atomicAdd(c.counters + CNT_NARROW_INVALID, st_invalid);
            atomicAdd(c.counters + CNT_NODE_PAIRS, st_nodes);
            atomicAdd(c.counters + CNT_LEAF_PAIRS, st_leaves);
            atomicAdd(c.counters + CNT_TRI_PAIRS, st_tris);
            atomicAdd(c.counters + CNT_TAIL_PAIRS, st_tails);
            atomicAdd(c.counters + CNT_OVERFLOW, st_over);
            atomicAdd(c.counters + CNT_NODE_ROUNDS, st_nr);
            atomicAdd(c.counters + CNT_LEAF_ROUNDS, st_lr);
            atomicAdd(c.counters + CNT_TAIL_ROUNDS, st_tr);
            atomicAdd(c.counters + CNT_SHARED, st_shared);
            atomicAdd(c.counters + CNT_PREFILTER, st_pre);
            atomicAdd(c.counters + CNT_PLANE1, st_pl1);
        }
    }
    // the last CTA to finish leaves the list counters at zero for the next call (no memset per
    // call): every CTA has read the list length and drawn its last entry by the time it gets here.
    // With a gather attached (rc_check_gather) the same CTA then delivers the bitmask slot, which
    // is complete at this point, to the peers.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const bool last = atomicAdd(a.work_count + kWorkBins + 1, 1ULL) == (unsigned long long)gridDim.x - 1ULL;
        if (last)
            for (int q = 0; q < kWorkBins + 2; q++) a.work_count[q] = 0;
        __threadfence();
        s_misc[1] = last ? 1u : 0u;
    }
    if (a.g.world > 1) {
        __syncthreads();
        if (s_misc[1]) gather_tail(a.g);
    }
}

static size_t culled_smem_bytes(const DeviceTables &t, int warps, bool gobs)
{
    size_t b = 16 + (size_t)t.n_rob_v_pad * 16 + (size_t)t.n_rob_nodes * 32 +
               (size_t)(t.n_rob_t + (t.n_rob_t & 1)) * 8;
    if (!gobs)
        b += (size_t)t.n_obs_nodes * 32 + (size_t)t.n_obs_t * 48 + (size_t)pad4(t.n_obs_t) * 4;
    else if (t.rob_leaf == 1)
        b += (size_t)t.n_rob_t * 16;  // robot triangle planes
    b += (size_t)warps * (stack_cap_of(warps, gobs) + kLeafQCap + kTailQCap) * 4;
    b = (b + 15) / 16 * 16 + kMailSlots * sizeof(Mail) + (5 * (size_t)warps + 7 + kWorkBins + 1 + 2 + 320) * 4 + 8 + sizeof(FeedShared);  // + s_pend, s_idle, s_misc, s_bin, s_wait, s_fetch, s_seed
    return b;
}

template <bool GOBS, bool COP, bool RL1>
static void (*pick_culled(int warps, int want_stats))(const CulledArgs)
{
    if (warps == 16) return want_stats ? check_culled_kernel<true, 16, GOBS, COP, RL1> : check_culled_kernel<false, 16, GOBS, COP, RL1>;
    if (warps == 32) return want_stats ? check_culled_kernel<true, 32, GOBS, COP, RL1> : check_culled_kernel<false, 32, GOBS, COP, RL1>;
    return want_stats ? check_culled_kernel<true, 24, GOBS, COP, RL1> : check_culled_kernel<false, 24, GOBS, COP, RL1>;
}

template <bool COP>
static void (*pick_culled_feed(int warps))(const CulledArgs)
{
    // the single-launch form exists for the full-size CTA with the obstacle in shared memory
    return warps == 32 ? check_culled_kernel<false, 32, false, COP, false, true> : nullptr;
}

// Obstacle tables go to shared memory when everything fits with at least 16 warps per CTA; otherwise
// they stay in global memory (any obstacle size, the robot tables must still fit).  Decided once, when
// the context is built (the robot tree of a global-obstacle context has one triangle per leaf).
bool culled_obstacle_in_shared(const rc_ctx *ctx)
{
    const DeviceTables &t = ctx->t;
    if (ctx->knob_global_obstacle) return false;
    return culled_smem_bytes(t, 16, false) <= (size_t)ctx->max_smem_optin && t.n_obs_nodes <= 65535 &&
           t.n_rob_nodes <= 65535;
}

cudaError_t launch_check_culled(const rc_ctx *ctx, const rc_config *d_confs, long n,
                                uint8_t *d_valid, uint32_t *d_bitmask, int want_stats,
                                cudaStream_t stream, unsigned long long *queue, const CheckExtras *extras)
{
    const DeviceTables &t = ctx->t;
    const bool gobs = t.gobs, rl1 = t.gobs && t.rob_leaf == 1;
    int warps = ctx->knob_warps;
    if (warps != 16 && warps != 24 && warps != 32) {
        // Warps per CTA by batch size (obstacle in shared memory): a short call is bound by its longest
        // descents, and a warp that shares its scheduler with 3 or 5 others instead of 7 gets through one
        // faster (and has a deeper stack); a long call is bound by throughput and wants all 32.  Measured
        // (profiles/ab_r02.txt 6): 512 hard configurations 42.0 -> 35.8 us with 16 warps, 10,000: 103.4 -> 89.2
        // and 30,000: 154.7 -> 129.9 with 24, 100,000: 283.6 (32) against 287.8 (24) and 338.9 (16).
        // With the obstacle in global memory the longest descents take milliseconds and bound calls of any
        // size tried: 24 warps (one C5 shard of 1.25 M configurations 2.53 -> 2.16 ms, 10 M: 9.2-9.3 -> 8.9 ms).
        warps = gobs ? 24 : n <= 4096 ? 16 : n < 65536 ? 24 : 32;
    }
    while (warps > 16 && culled_smem_bytes(t, warps, gobs) > (size_t)ctx->max_smem_optin) warps -= 8;
    const size_t smem = culled_smem_bytes(t, warps, gobs);
    if (smem > (size_t)ctx->max_smem_optin || t.n_rob_v > 65535) return cudaErrorInvalidConfiguration;
    if (gobs && (t.n_rob_nodes >= (1 << 11) || t.n_rob_t >= (1 << 11) || t.n_obs_nodes >= (1 << 21) ||
                 t.n_obs_t >= (1 << 21)))
        return cudaErrorInvalidConfiguration;
    CulledArgs a;
    a.c.rob_v = t.rob_v; a.c.rob_t = t.rob_t_sorted;
    a.c.obs_a = t.obs_a_sorted; a.c.obs_b = t.obs_b_sorted; a.c.obs_c = t.obs_c_sorted;
    a.c.obs_d = t.obs_d_sorted;
    a.c.n_rob_v_pad = t.n_rob_v_pad; a.c.n_rob_t = t.n_rob_t; a.c.n_obs_t = t.n_obs_t;
    a.c.obs_box = t.obs_box;
    a.c.confs = d_confs; a.c.n = n; a.c.valid = d_valid; a.c.bitmask = d_bitmask;
    a.c.counters = ctx->d_counters; a.c.queue = queue; a.c.hit_counts = nullptr;
    a.rob_bvh = t.rob_bvh; a.obs_bvh = t.obs_bvh;
    a.n_rob_nodes = t.n_rob_nodes; a.n_obs_nodes = t.n_obs_nodes;
    a.sph_x = t.rob_center[0]; a.sph_y = t.rob_center[1]; a.sph_z = t.rob_center[2];
    a.sph_r = t.rob_radius;
    a.box_cx = t.rob_box_c[0]; a.box_cy = t.rob_box_c[1]; a.box_cz = t.rob_box_c[2];
    a.box_ex = t.rob_box_e[0]; a.box_ey = t.rob_box_e[1]; a.box_ez = t.rob_box_e[2];
    a.probe_stride = max(1, t.n_rob_v / 32);
    const int grid = ctx->sm_count;
    // small chunks: the per-configuration cost varies 1:8 (cleared / full sweep / colliding)
    a.c.chunk = 1;  // the list-driven kernel sizes its chunks from the work count on the device
    a.tail_begin = (long)grid * warps * 4;
    a.tail_thresh = gobs ? 32 : 16; a.leaf_thresh = 8;  // one-triangle robot leaves queue many finishing pairs: full rounds
    const int sd = ctx->knob_seed_depth < 0 ? 2 : min(3, ctx->knob_seed_depth);
    a.n_seed_r = t.n_seed_r[sd]; a.n_seed_o = t.n_seed_o[sd];
    for (int q = 0; q < 8; q++) { a.seed_r[q] = t.seed_r[sd][q]; a.seed_o[q] = t.seed_o[sd][q]; }
    a.stack_cap = stack_cap_of(warps, gobs);
    if (ctx->knob_stack_cap > 0) a.stack_cap = max(64, min(a.stack_cap, ctx->knob_stack_cap));
    if (ctx->knob_tail > 0) a.tail_thresh = min(32, ctx->knob_tail);
    if (ctx->knob_leaf > 0) a.leaf_thresh = min(8, ctx->knob_leaf);
    a.share = ctx->knob_share;
    // a call so short that every configuration is cut into 16 or more parts (the list cannot be longer
    // than n): the parts are small, handing pieces of them on costs more than it saves (64 hard
    // configurations 35.8 -> 33.8 us, 2,000: 62.4 -> 60.4 us; from 10,000 on the hand-over pays)
    if (((unsigned long long)n << 4) <= (unsigned long long)max(1, ctx->knob_split_load) * grid * warps &&
        ctx->knob_split_max >= 16)
        a.share = 0;
    a.hit_counts = (want_stats && extras) ? extras->hit_counts : nullptr;
    if (a.hit_counts) a.share = 0;  // counting: a configuration stays with one warp
    // flags: written once per configuration by its warp (statistics kernels), or optimistically --
    // marked valid by step 1, overwritten by whichever part of the descent finds a hit
    a.optimistic = want_stats ? 0 : 1;
    a.rob_plane = t.rob_plane;
    a.split_max = max(1, min(kSplitMax, ctx->knob_split_max));
    a.split_load = max(1, ctx->knob_split_load);
    a.tri_margin = 2.f * t.box_margin;
    a.n_seed_r3 = t.n_seed_r4; a.n_seed_o3 = t.n_seed_o4;
    for (int q = 0; q < 16; q++) { a.seed_r3[q] = t.seed_r4[q]; a.seed_o3[q] = t.seed_o4[q]; }
    if (extras && extras->gather) a.g = *extras->gather;
    else { memset(&a.g, 0, sizeof a.g); a.g.world = 1; }
    a.mail_slots = max(1, min(kMailSlots, ctx->knob_mail_slots));
    a.donate_min = max(4, ctx->knob_donate_min);
    a.donate_leaf = max(8, ctx->knob_donate_leaf);
#ifdef RC_TIMELINE
    a.trace = ctx->d_trace;
#else
    a.trace = want_stats ? ctx->d_trace : nullptr;
#endif
    a.work_idx = ctx->d_work_idx; a.work_mat = ctx->d_work_mat; a.work_count = ctx->d_work_count;
    a.work_cap = ctx->work_cap;
    if (!ctx->d_work_idx || ctx->work_cap < n) return cudaErrorMemoryAllocation;  // rc_check sizes it

    // step 1: classify (thread per configuration); writes every bitmask word, so no memset
    ClassifyArgs &k = a.k;
    k.confs = d_confs; k.n = n; k.valid = d_valid; k.bitmask = d_bitmask; k.obs_box = t.obs_box;
    k.sph_x = a.sph_x; k.sph_y = a.sph_y; k.sph_z = a.sph_z; k.sph_r = a.sph_r;
    k.box_cx = a.box_cx; k.box_cy = a.box_cy; k.box_cz = a.box_cz;
    k.box_ex = a.box_ex; k.box_ey = a.box_ey; k.box_ez = a.box_ez;
    for (int q = 0; q < 2 * kClassifyBoxes; q++) k.bound[q] = t.h_rob_bound[q];
    k.n_bound = ctx->knob_bound ? t.n_rob_bound : 0;
    k.work_idx = ctx->d_work_idx; k.work_mat = ctx->d_work_mat; k.work_count = ctx->d_work_count;
    k.work_cap = ctx->work_cap;
    k.counters = ctx->d_counters;
    k.optimistic = a.optimistic;
    memset(&a.feed, 0, sizeof a.feed);
    if (extras && extras->feed && !want_stats && !gobs && ctx->d_feed_slots && n <= ctx->work_cap && n < (1L << 28)) {
        // single launch: the descent kernel classifies the configurations itself (FeedArgs)
        void (*fk)(const CulledArgs) = ctx->coplanar_mode ? pick_culled_feed<true>(warps) : pick_culled_feed<false>(warps);
        if (fk) {
            a.feed.on = 1;
            int gs = 0;
            while ((2 << gs) <= a.split_max &&
                   ((unsigned long long)n << (gs + 1)) <= (unsigned long long)a.split_load * grid * warps)
                gs++;
            a.feed.gs = gs;
            a.feed.cls_warps = extras->feed_cls_warps > 0 ? min(warps, extras->feed_cls_warps) : min(warps, 8);
            a.feed.gen = extras->feed_gen;
            a.feed.slots = ctx->d_feed_slots;
            a.feed.ctr = ctx->d_feed_ctr;
            k.optimistic = 1;
            cudaError_t e = cudaMemsetAsync(ctx->d_feed_ctr, 0, kFeedCounters * kFeedStride * sizeof(uint32_t), stream);
            if (e != cudaSuccess) return e;
            if ((e = cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
            fk<<<grid, warps * 32, smem, stream>>>(a);
            return cudaGetLastError();
        }
    }
    void (*kern)(const CulledArgs);
    if (ctx->coplanar_mode)
        kern = rl1 ? pick_culled<true, true, true>(warps, want_stats)
               : gobs ? pick_culled<true, true, false>(warps, want_stats) : pick_culled<false, true, false>(warps, want_stats);
    else
        kern = rl1 ? pick_culled<true, false, true>(warps, want_stats)
               : gobs ? pick_culled<true, false, false>(warps, want_stats) : pick_culled<false, false, false>(warps, want_stats);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // small blocks: 100,000 configurations are only 21 warps per SM, spread them evenly
    const int cgrid = (int)min((n + 63) / 64, (long)ctx->sm_count * 256);
    if (want_stats) classify_kernel<true><<<cgrid, 64, 0, stream>>>(k);
    else classify_kernel<false><<<cgrid, 64, 0, stream>>>(k);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;

    // step 2: the remaining configurations, one warp each; programmatic dependent launch, so the
    // CTAs are placed and stage their tables while the tail of step 1 is still running
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)(warps * 32));
    cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = ctx->knob_pdl ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e != cudaSuccess)  // the list counters were filled by step 1 and nobody will reset them
        cudaMemsetAsync(ctx->d_work_count, 0, kWorkCounters * sizeof(unsigned long long), stream);
    return e;
}

}  // namespace rc
