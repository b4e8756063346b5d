// rc_internal.h -- context layout shared by the C-ABI translation units (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/robocheck.h"

namespace rc {

// Device-resident mesh tables, laid out for 128-bit loads and for staging into shared
// memory with one TMA bulk copy per table (every table starts 16-byte aligned and its
// byte size is a multiple of 16).
struct DeviceTables {
    // robot, local frame
    float4 *rob_v = nullptr;    // [n_rob_v]  (x, y, z, 0)
    uint2 *rob_t = nullptr;     // [n_rob_t]  (v1 | v2 << 16, v3)          indices < 65536
    rc_vec3 *rob_aos = nullptr; // [n_rob_v]  AoS copy for the genTransformedCopies stage
    // obstacle, world frame, gathered per triangle ("SoA of float4")
    float4 *obs_a = nullptr;    // [n_obs_t]  (p1.x, p1.y, p1.z, N.x)
    float4 *obs_b = nullptr;    // [n_obs_t]  (p2.x, p2.y, p2.z, N.y)
    float4 *obs_c = nullptr;    // [n_obs_t]  (p3.x, p3.y, p3.z, N.z)
    float *obs_d = nullptr;     // [n_obs_t↑4] s = N.p1 (the reference's d is -s)
    // BVHs for the culled path (see bvh.h): two float4 per node
    float4 *rob_bvh = nullptr;  // [2 * n_rob_nodes]
    float4 *obs_bvh = nullptr;  // [2 * n_obs_nodes]
    uint2 *rob_t_sorted = nullptr;   // robot triangles in BVH leaf order
    float4 *obs_a_sorted = nullptr, *obs_b_sorted = nullptr, *obs_c_sorted = nullptr;
    float *obs_d_sorted = nullptr;   // obstacle triangles in BVH leaf order
    // sphere-approximated robot (optional, built on first use): spheres in sphere-tree leaf order
    float4 *sph = nullptr, *sph_bvh = nullptr;
    int n_sph = 0, n_sph_nodes = 0;
    float sph_bound[4] = {0, 0, 0, 0};  // bounding sphere of the set, local frame
    float sph_box_c[3] = {0, 0, 0}, sph_box_e[3] = {0, 0, 0};  // its local AABB (radii included), rounded up
    int n_rob_v = 0, n_rob_t = 0, n_obs_v = 0, n_obs_t = 0;
    int n_rob_nodes = 0, n_obs_nodes = 0;
    int n_rob_v_pad = 0;        // n_rob_v rounded up to 32 (padding repeats vertex 0)
    float box_margin = 1e-3f;   // inflation of every tree box; scales with the scene (rc_create)
    bool gobs = false;          // obstacle tables stay in global memory (too large for shared memory)
    int rob_leaf = 4;           // triangles per leaf of the robot tree: 1 with a global-memory obstacle
    float4 *rob_plane = nullptr;  // [n_rob_t] plane (n, s = n . p1) of every robot triangle, robot frame, leaf order
    rc_aabb obs_box;
    float rob_center[3];        // bounding sphere of the robot in its local frame
    float rob_radius;
    float rob_box_c[3], rob_box_e[3];
    // boxes (local frame) of a k-d partition of ALL robot vertices: a much tighter outer bound of
    // the exact vertex AABB than the single rotated box; two float4 (centre, half extent) each
    float4 h_rob_bound[2 * 32];                // passed to classify_kernel by value (constant bank)
    int n_rob_bound = 0;                       // 32, or 0 if the mesh is empty
    // traversal seeds: node cuts at depth 0..3 of both trees (<= 8 nodes each)
    uint32_t seed_r[4][8], seed_o[4][8];
    int n_seed_r[4], n_seed_o[4];
    uint32_t seed_r4[16], seed_o4[16];  // the cut at depth 4 (<= 16 nodes each): seeds of a split descent
    int n_seed_r4 = 0, n_seed_o4 = 0;
};

}  // namespace rc

// classify_kernel sorts the configurations it cannot clear into this many cost classes (how much
// of the robot lies inside the obstacle AABB); the descent takes the dearest class first so that
// the long descents are under way early and the work list ends with cheap ones
constexpr int kWorkBins = 8;
constexpr int kHostWaves = 4;           // rc_check_host cuts a large batch into at most this many waves
constexpr int kWorkCounters = 16;  // [0..7] entries per class, [8] next entry, [9] CTAs finished

// one of the two batches rc_check_host_submit may have in flight
struct PipeLane {
    rc_config *d_confs = nullptr;   // device staging of the configurations
    rc_config *h_confs = nullptr;   // page-locked staging for pageable callers
    uint8_t *h_valid = nullptr;     // page-locked staging for pageable callers
    uint8_t *d_valid = nullptr;     // device flags when the host buffer cannot be mapped
    long cap = 0;
    cudaEvent_t copied = nullptr, done = nullptr;
    bool busy = false;
    int ticket = -1;
    uint8_t *user_valid = nullptr;  // where the flags go at rc_check_host_wait (nullptr: already there)
    long n = 0;
};

struct rc_ctx {
    PipeLane pipe[2];
    int pipe_next = 0;
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    rc::DeviceTables t;
    // host copies (for rc_get_* and for rebuilding)
    std::vector<float> rob_x, rob_y, rob_z, obs_x, obs_y, obs_z;
    std::vector<int> rob_v1, rob_v2, rob_v3, obs_v1, obs_v2, obs_v3;
    std::vector<float> spheres;                // robot sphere set as given / generated, n x 4
    // scratch owned by the context
    unsigned long long *d_counters = nullptr;  // [RC_NUM_COUNTERS] work queue + statistics
    rc_aabb *d_obs_box = nullptr;              // obstacle AABB on the device (broad-phase stage)
    int stats_on = 0;
    uint32_t gather_calls = 0;                 // number of the last rc_check_gather call on gather_buf (all ranks in lock-step)
    uint32_t *gather_buf = nullptr;            // the own gather buffer that numbering belongs to
    cudaEvent_t ev_wave[2 * 4] = {};           // rc_check_host waves: copied (first half), checked (second half)
    cudaEvent_t ev_scratch = nullptr;          // behind the last launch that used the scratch below
    cudaStream_t scratch_stream = nullptr;
    bool scratch_used = false;
    int knob_host_chunks = 0;                  // RC_HOST_CHUNKS: waves rc_check_host cuts a batch into (0 = automatic)
    int coplanar_mode = 0;                     // rc_set_coplanar_mode: 0 = skip like the reference, 1 = decide in the plane
    uint32_t *d_trace = nullptr;               // rc_stats_trace (caller's memory)
    unsigned long long launches = 0;           // kernels launched through this context
    rc_config *d_stage_confs = nullptr;        // device staging for rc_check_host
    uint8_t *d_stage_valid = nullptr;
    uint32_t *d_stage_mask = nullptr;          // bitmask words of the staged configurations (rc_check_multi)
    uint32_t *h_pinned_mask = nullptr;
    unsigned long long peers_enabled = 0;      // devices this context's device may copy to directly
    rc_config *h_pinned_confs = nullptr;
    uint8_t *h_pinned_valid = nullptr;
    long stage_capacity = 0;
    cudaStream_t own_stream = nullptr, own_stream2 = nullptr;
    cudaEvent_t ev_join = nullptr;
    // work list between the two kernels of rc_check, grown on demand to the largest n seen
    uint32_t *d_work_idx = nullptr;            // [kWorkBins][work_cap]: configuration indices per cost class
    float4 *d_work_mat = nullptr;
    unsigned long long *d_work_count = nullptr;
    long work_cap = 0;
    uint2 *d_feed_slots = nullptr;             // single-launch form: [kWorkBins][work_cap] (call number, configuration)
    uint32_t *d_feed_ctr = nullptr;            // its 32 counters
    uint32_t feed_gen = 0;                     // call number of the last single-launch check
    int knob_feed = 1;                         // RC_FEED: 0 never, 1 rc_check_host (configurations read in place over PCIe), 2 rc_check too
    int knob_feed_cls = 0;                     // RC_FEED_CLS: classifying warps per CTA (0 = automatic)
    uint32_t *d_block_counts = nullptr;        // scratch of rc_compact, grown on demand
    long block_counts_cap = 0;
    unsigned long long *d_queues = nullptr;    // [kHostChunks] work-queue counters of rc_check_host
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_kernel_ms = 0.f;
    // tuning knobs, read from the environment once at rc_create (RC_CULLED_WARPS, RC_SEED_DEPTH,
    // RC_TAIL_THRESH, RC_LEAF_THRESH, RC_CHUNK, RC_FORCE_GLOBAL_OBSTACLE); 0 / -1 = default
    int knob_warps = 0, knob_seed_depth = -1, knob_tail = 0, knob_leaf = 0, knob_chunk = 0;
    int knob_stack_cap = 0;  // RC_STACK_CAP: shrink the pair stack to exercise the overflow fallback
    bool knob_global_obstacle = false;
    int knob_pdl = 1;        // RC_PDL=0: no programmatic dependent launch of the descent step
    int knob_bound = 1;      // RC_BOUND=0: classify without the vertex-group boxes
    int knob_mail_slots = 4, knob_donate_min = 40, knob_donate_leaf = 16;  // RC_MAIL_SLOTS, RC_DONATE_MIN, RC_DONATE_LEAF
    int knob_host_direct = 0; // RC_HOST_DIRECT=1: rc_check_host stores the flags into the pinned buffer from the kernels instead of copying them back
    int knob_share = 1;      // RC_SHARE=0: warps of a CTA do not hand pieces of a deep descent to each other
    int knob_split_max = 64;  // RC_SPLIT_MAX: parts a configuration may be cut into when the work list is short (1 = never)
    int knob_split_load = 8;  // RC_SPLIT_LOAD: split while entries <= this many per warp of the grid
    std::string last_error;
    std::vector<void *> allocations;
};

namespace rc {

enum Counter {
    CNT_QUEUE = 0,        // next configuration index (persistent work queue)
    CNT_BROAD_CLEARED,    // statistics below are only updated when stats are requested
    CNT_SPHERE_CLEARED,
    CNT_NARROW_VALID,
    CNT_NARROW_INVALID,
    CNT_NODE_PAIRS,
    CNT_LEAF_PAIRS,
    CNT_TRI_PAIRS,
    CNT_TAIL_PAIRS,
    CNT_OVERFLOW,
    CNT_LAUNCHES,
    CNT_NODE_ROUNDS,
    CNT_LEAF_ROUNDS,
    CNT_TAIL_ROUNDS,
    CNT_SHARED,           // pieces of a descent handed from one warp to another
    CNT_WATCHDOG,         // spin loops of the hand-over protocol that gave up (always counted; must stay 0)
    CNT_WATCHDOG_WHERE,   // bit 0: serve loop of a warp without work, bit 3: gather wait
    CNT_PREFILTER,        // triangle pairs whose boxes overlap (leaf rounds)
    CNT_PLANE1,           // ... that also pass the first plane test
    RC_NUM_COUNTERS = 24
};
constexpr int kHostChunks = 8;

int set_error(rc_ctx *ctx, int code, const std::string &msg);
int check_cuda(rc_ctx *ctx, cudaError_t e, const char *what);

// all-gather of the bitmask through peer-mapped buffers (rc_gather.cuh)
constexpr int kMaxGatherRanks = 16;
constexpr int kGatherHeaderWords = 4;        // behind the signal row: calls, timeouts, 2 reserved
struct GatherArgs {
    const uint32_t *src;                     // this rank's slot in its own buffer
    uint32_t *peer_slot[kMaxGatherRanks];    // the same slot inside every rank's buffer
    uint32_t *peer_signal[kMaxGatherRanks];  // word [rank] of every rank's signal row
    uint32_t *my_row;                        // this rank's signal row; [world] calls, [world + 1] timeouts
    int rank, world, words;
    int wait;                                // 1: also wait for the peers' slots (one-call form)
    uint32_t call;                           // 1, 2, 3, ...
    unsigned long long *counters;
    unsigned long long timeout_ns;
};
// what a check launch may carry besides the flags
struct CheckExtras {
    const GatherArgs *gather = nullptr;      // culled mode: the last CTA pushes the bitmask slot to the peers
    uint32_t *hit_counts = nullptr;          // statistics kernels: hitting pairs per configuration, no early exit
    int feed = 0;                            // culled mode: single launch, the descent kernel classifies (rc_culled.cuh: FeedArgs)
    int feed_cls_warps = 0;                  // warps per CTA that classify (0 = all; 2 when the configurations are read over PCIe)
    uint32_t feed_gen = 0;
};

// kernels.cu launchers (all asynchronous on `stream`)
// `queue` = the work-queue counter of this launch (concurrent launches need distinct ones)
cudaError_t launch_check(const rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid,
                         uint32_t *d_bitmask, int mode, int want_stats, cudaStream_t stream,
                         unsigned long long *queue, const CheckExtras *extras = nullptr);
cudaError_t launch_check_culled(const rc_ctx *ctx, const rc_config *d_confs, long n,
                                uint8_t *d_valid, uint32_t *d_bitmask, int want_stats,
                                cudaStream_t stream, unsigned long long *queue,
                                const CheckExtras *extras = nullptr);
cudaError_t launch_check_spheres(const rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid,
                                 uint32_t *d_bitmask, int want_stats, cudaStream_t stream,
                                 unsigned long long *queue);
bool culled_obstacle_in_shared(const rc_ctx *ctx);
cudaError_t launch_gather_push(const rc_ctx *ctx, const GatherArgs &a, cudaStream_t stream);
cudaError_t launch_gather_wait(const rc_ctx *ctx, const GatherArgs &a, cudaStream_t stream);
// device known-answer tests of the per-pair helpers (rc_kat.cuh)
cudaError_t launch_kat_pairs(const float *d_r9, const float *d_o9, int n, float *d_out, int *d_verdict,
                             cudaStream_t stream);
cudaError_t launch_kat_la_solve(const float *d_in, int n, float *d_out, cudaStream_t stream);
cudaError_t launch_fma_peak(float *d_out, int grid, int iters, cudaStream_t stream);
cudaError_t launch_conf_generate(rc_config *d_out, long n, const float box[6], uint64_t seed,
                                 cudaStream_t stream);
cudaError_t launch_conf_generate_mix(rc_config *d_out, long first, long n, const float box_a[6],
                                     const float box_b[6], int block, int period, uint64_t seed,
                                     cudaStream_t stream);
cudaError_t launch_compact(const uint8_t *d_valid, const uint32_t *d_mask, long n, int want,
                           uint32_t *d_indices, unsigned long long *d_count, uint32_t *d_block_counts,
                           cudaStream_t stream);
cudaError_t launch_transform_soa(const rc_ctx *ctx, const rc_config *d_confs, long n, float *d_x,
                                 float *d_y, float *d_z, rc_aabb *d_aabb, uint8_t *d_valid,
                                 cudaStream_t stream);
cudaError_t launch_transform_aos(const rc_ctx *ctx, const rc_config *d_confs, long n,
                                 const rc_vec3 *d_base, int n_v, rc_vec3 *d_out, cudaStream_t stream);
cudaError_t launch_aabb_aos(const rc_vec3 *d_verts, unsigned n_v, unsigned n, rc_aabb *d_out,
                            cudaStream_t stream);
cudaError_t launch_aabb_soa(const float *d_x, const float *d_y, const float *d_z, unsigned n_v,
                            unsigned n, rc_aabb *d_out, cudaStream_t stream);
cudaError_t launch_broad(long n, const rc_aabb *d_robots, const rc_aabb *d_obstacle,
                         uint8_t *d_valid, cudaStream_t stream);
cudaError_t launch_narrow_soa(const rc_ctx *ctx, const float *d_x, const float *d_y,
                              const float *d_z, long n, uint8_t *d_valid, cudaStream_t stream);
cudaError_t launch_obstacle_planes(const float4 *a, const float4 *b, const float4 *c, float4 *a_out,
                                   float4 *b_out, float4 *c_out, float *d_out, int n,
                                   cudaStream_t stream);

}  // namespace rc
