/*
 * robocheck.h -- C ABI of the B200-native per-configuration collision check.
 *
 * Drop-in boundary for robo-check's hot path.  The reference has no FFI layer: its boundary
 * is the set of C++ declarations in the stage headers, called directly by its test drivers.
 * Each entry point below names the reference interface it replaces (paths relative to the
 * reference tree).  include/robocheck_compat.hpp re-exposes the reference's own C++
 * signatures on top of this ABI so existing call sites keep compiling.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types; every function returns an int
 *    status (RC_OK or a negative RC_ERR_*) and never exits or prints.
 *  - `valid` flags are bytes, 1 = collision free, 0 = in collision (reference: bool*,
 *    true = valid, narrow-phase/narrow-phase.cu:341).
 *  - pointers named d_* are DEVICE pointers on the context's device, h_* are host pointers.
 *  - `stream` is a cudaStream_t passed as void* (NULL = CUDA's legacy default stream, as in
 *    the runtime API); stage calls are asynchronous on it unless stated otherwise.  The *_host
 *    convenience calls use a private stream of the context and synchronise before returning.
 *  - one context per device; calls on one context are serialised by the caller (host side).  Device
 *    work the calls leave queued is ordered by the library: entry points of one context share its
 *    scratch (work lists, about 80 bytes per configuration of the largest batch seen), and a launch
 *    on another stream than the previous one first waits for it (an event), so rc_check on the
 *    caller's stream followed by rc_check_host on the context's own stream is safe.
 *  - there is NO CPU fallback: without a usable CUDA device rc_create fails with
 *    RC_ERR_NO_DEVICE.
 */
#ifndef ROBOCHECK_H
#define ROBOCHECK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RC_OK 0
#define RC_ERR_INVALID (-1)     /* bad argument                                   */
#define RC_ERR_CUDA (-2)        /* a CUDA runtime call failed (see rc_last_error) */
#define RC_ERR_NOMEM (-3)
#define RC_ERR_IO (-4)          /* file could not be opened                       */
#define RC_ERR_FORMAT (-5)      /* file could not be parsed                       */
#define RC_ERR_UNSUPPORTED (-6) /* mesh too large for the requested path          */
#define RC_ERR_NO_DEVICE (-7)   /* no CUDA device: there is no CPU fallback       */

/* POD types, layout-identical to the reference's (Utils_rai.h:51-58, 70-96). */
typedef struct { float x, y, z, pitch, yaw, roll; } rc_config;                  /* Configuration */
typedef struct { float x, y, z; } rc_vec3;                                      /* Vector3f      */
typedef struct { int v1, v2, v3; } rc_tri;                                      /* Triangle      */
typedef struct { float x_min, y_min, z_min, x_max, y_max, z_max; } rc_aabb;     /* AABB          */

typedef struct rc_ctx rc_ctx;
typedef struct rc_mesh rc_mesh;

/* rc_check modes -- the exactness contract.
 * RC_MODE_BRUTE decides every pair the reference decides, with the reference's arithmetic: its
 * flags are bit-identical to the reference pipeline on ANY input, by construction.
 * RC_MODE_CULLED decides with the same arithmetic every pair it looks at, but never looks at a pair
 * whose triangles are separated by more than twice the box margin (2e-3 on scenes up to 250 units,
 * proportionally more beyond: an ancestor pair of inflated boxes is disjoint).  Its flags equal the
 * reference's whenever the reference reports no such box-disjoint pair.  That is an EMPIRICAL
 * property of the scene, not a theorem: for two nearly coplanar triangles the reference's
 * intersection line is rounding noise and it can report a pair whose triangles lie far apart inside
 * their common plane (tests/coplanar_scene.py: one configuration in 3,000).  RC_MODE_CULLED calls
 * such a configuration valid, i.e. where the two modes differ the culled answer is the geometric
 * one.  On alpha1.0 no such pair exists in any set examined (every BASELINE workload).
 * rc_audit() below measures the property on the caller's own scene and configurations: flags AND
 * per-configuration hitting-pair counts of the two modes side by side. */
#define RC_MODE_CULLED 0 /* BVH-culled fused kernel (default, fastest)                        */
#define RC_MODE_BRUTE 1  /* fused kernel, all n_rob_t x n_obs_t pairs: the reference's sweep  */

/* ---- context ------------------------------------------------------------------------------
 * Replaces the per-call set-up inside broadPhaseFused_sep (broad-phase/broad-phase-fused.cu:
 * 365-476: OBJ parse, cudaMalloc, 12x cudaMemcpyToSymbol into the __constant__ mesh tables
 * declared at Utils_rai.h:118-129, obstacle AABB via generateAABBBaseline).  Meshes are SoA
 * host arrays exactly as the reference's SoA loadOBJFile overload produces them
 * (Utils_rai.cu:228-320).  Sizes are run-time values (the reference bakes 792/1008 in). */
int rc_create(rc_ctx **out, const float *rob_x, const float *rob_y, const float *rob_z, int n_rob_v,
              const int *rob_v1, const int *rob_v2, const int *rob_v3, int n_rob_t,
              const float *obs_x, const float *obs_y, const float *obs_z, int n_obs_v,
              const int *obs_v1, const int *obs_v2, const int *obs_v3, int n_obs_t, int device);
int rc_create_from_meshes(rc_ctx **out, const rc_mesh *robot, const rc_mesh *obstacle, int device);
void rc_destroy(rc_ctx *ctx);
const char *rc_strerror(int code);
const char *rc_last_error(const rc_ctx *ctx);
int rc_device(const rc_ctx *ctx);
int rc_mesh_sizes(const rc_ctx *ctx, int *n_rob_v, int *n_rob_t, int *n_obs_v, int *n_obs_t);
/* obstacle AABB as generateAABBBaseline(x,y,z) computes it (generate-AABB/generate-AABB.cu:183-201) */
int rc_obstacle_aabb(const rc_ctx *ctx, rc_aabb *out);

/* ---- fused entry point ("MegaKernel") -------------------------------------------------------
 * Replaces  MegaKernel<<<>>> / CallMegaKernel (test/MegaKernel.hu:6-9) and the live pipeline
 * broadPhaseFused_sep = broadPhaseFusedKernel_sep + narrowPhaseKernel_coarse
 * (broad-phase-fused.cu:365-520).  One persistent kernel: SE(3) transform, warp-shuffle AABB,
 * broad phase, narrow phase with early exit.  Flags: see the exactness contract of the two modes
 * above.  d_valid (n bytes) and d_bitmask (ceil(n/32) words, bit i%32 of word i/32 = valid)
 * are both optional but at least one must be given.  Unlike MegaKernel the outputs need no
 * pre-fill.  Asynchronous. */
int rc_check(rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid, uint32_t *d_bitmask,
             int mode, void *stream);
/* Same, host buffers, synchronised on return.  This is the call behind the
 * broadPhaseFused_sep(std::vector<Configuration>&, bool*) shim.  From 65,536 configurations on (below that the
 * configurations are copied first: the copy is short there) RC_MODE_CULLED runs as ONE launch
 * with no host-to-device copy in front of it: the kernel reads the configurations where they lie -- the
 * caller's buffer if it is page-locked (cudaMallocHost / cudaHostRegister), else the pinned staging
 * area a pageable buffer is copied to first -- over PCIe, a few warps per CTA classifying batches of
 * 32 as they come in while the other warps already run descents on the ones that need them; the flags
 * are copied back.  RC_FEED=0 in the environment restores copy + two launches. */
int rc_check_host(rc_ctx *ctx, const rc_config *h_confs, long n, uint8_t *h_valid, int mode);

/* Single-process multi-GPU check (north_star: configurations shard across the GPUs of one box, the
 * meshes replicated, only the validity bitmask gathered).  ctxs[g] = one context per device, built
 * from the same meshes.  Shard g = the rc_shard_size(n, n_gpus) configurations starting at
 * g * rc_shard_size (a multiple of 32, so every shard owns whole bitmask words; trailing shards may
 * be short or empty).  Every device stages and checks its shard concurrently.  Outputs, each
 * optional: h_bitmask ((n+31)/32 words, bit i%32 of word i/32 = configuration i is valid), h_valid
 * (n flag bytes), d_bitmask_all[g] = device buffer of (n+31)/32 words on ctxs[g]'s device: every
 * device ends up with the FULL bitmask -- each kernel writes its words into its own slot and the
 * slot is copied to the peers (cudaMemcpyPeerAsync, NVLink / NVSwitch), i.e. the all-gather of the
 * reference design without a communicator.  Blocking; returns when all outputs are complete.
 * (Multi-process callers -- one rank per GPU -- use rc_check + ncclAllGather on the bitmask
 * instead, see robo-check_b200/sharding.py and bench.py.) */
int rc_shard_size(long n, int n_gpus);
/* CUDA devices visible to this process (0 and RC_ERR_NO_DEVICE when there is none). */
int rc_device_count(int *count);
int rc_check_multi(rc_ctx **ctxs, int n_gpus, const rc_config *h_confs, long n, uint32_t *h_bitmask,
                   uint8_t *h_valid, uint32_t **d_bitmask_all);

/* ---- stage API ---------------------------------------------------------------------------- */
/* transformation/transform.hu:10-12 genTransformedCopies: AoS copies [conf][vertex].  d_base
 * may be NULL (use the context's robot).  Arithmetic is the live fused kernel's. */
int rc_transform_aos(rc_ctx *ctx, const rc_config *d_confs, long n, const rc_vec3 *d_base, int n_v,
                     rc_vec3 *d_out, void *stream);
/* generate-AABB/generate-AABB.hu:14-27 generateAABBPrimitiveKernel on device pointers ... */
int rc_generate_aabb(rc_ctx *ctx, const rc_vec3 *d_verts, unsigned n_v, unsigned n, rc_aabb *d_out,
                     void *stream);
/* ... and generateAABB(): host in/out, does its own H2D/D2H and synchronises. */
int rc_generate_aabb_host(rc_ctx *ctx, const rc_vec3 *h_verts, unsigned n_v, unsigned n,
                          rc_aabb *h_out);
/* SoA variant for the arrays the fused stage writes. */
int rc_generate_aabb_soa(rc_ctx *ctx, const float *d_x, const float *d_y, const float *d_z,
                         unsigned n_v, unsigned n, rc_aabb *d_out, void *stream);
/* broad-phase/broad-phase.hu:13 broadPhase: device pointers, asynchronous.  d_obstacle may be
 * NULL (use the context's obstacle AABB). */
int rc_broad_phase(rc_ctx *ctx, long n, const rc_aabb *d_robots, const rc_aabb *d_obstacle,
                   uint8_t *d_valid, void *stream);
/* broad-phase/broad-phase-fused.cu:157-207 broadPhaseFusedKernel_sep: transform + AABB + broad
 * phase; writes the transformed vertices as three SoA arrays indexed [i * n_rob_v + v]
 * (d_x/d_y/d_z may all be NULL to skip the store), optionally the per-config AABB, and
 * d_valid[i] = !overlap. */
int rc_broad_fused(rc_ctx *ctx, const rc_config *d_confs, long n, float *d_x, float *d_y, float *d_z,
                   rc_aabb *d_aabb, uint8_t *d_valid, void *stream);
/* narrow-phase/narrow-phase.hu:51-53 narrowPhaseKernel_coarse on the SoA arrays above:
 * d_valid is in/out, entries already 1 are skipped (narrow-phase.cu:578). */
int rc_narrow_phase(rc_ctx *ctx, const float *d_x, const float *d_y, const float *d_z, long n,
                    uint8_t *d_valid, void *stream);
/* narrow-phase/narrow-phase.hu:22-25 narrowPhase(): host AoS pointers, in/out flags, own
 * copies and synchronisation; meshes are the arguments, not the context's. */
int rc_narrow_phase_host(int n, int n_rob_t, int n_rob_v, int n_obs_t, int n_obs_v,
                         const rc_tri *rob_trs, const rc_vec3 *rob_pts, const rc_tri *obs_trs,
                         const rc_vec3 *obs_pts, uint8_t *valid, int device);

/* test/MegaKernel.hu:6-7 MegaKernel<<<>>> argument list on device pointers.  The obstacle AABB and
 * the per-triangle obstacle AABBs (generateTriAABBs, test/MegaKernel.cu:31-35) are accepted for
 * call-site compatibility and ignored: the context owns its own culling structures.  Like the
 * reference kernel this entry point only ever writes `false`: the caller pre-fills valid_confs
 * with true (MegaKernel.cu:448-452). */
int rc_megakernel(rc_ctx *ctx, const rc_config *d_configs, const rc_aabb *d_obs_aabb,
                  const rc_aabb *d_obs_tri_aabbs, uint8_t *d_valid_confs, int num_configs, void *stream);
/* generateTriAABBs (test/MegaKernel.cu:20-35): per-triangle AABBs of the context's obstacle, host
 * array of n_obs_t entries. */
int rc_obstacle_tri_aabbs(const rc_ctx *ctx, rc_aabb *h_out);

/* ---- planner loop helpers ----------------------------------------------------
 * Configurations generated on the device (no H2D copy; the reference spends a fixed 45-50 ms per
 * call on preparation, report 4.8): counter-based generator, uniform in the given box and in
 * [-pi, pi) for the angles like generateConfs (Utils_rai.cu:322-348) -- but NOT the mt19937 stream
 * of rc_conf_generate. */
int rc_conf_generate_device(rc_ctx *ctx, rc_config *d_out, long n, float x_min, float x_max,
                            float y_min, float y_max, float z_min, float z_max, uint64_t seed,
                            void *stream);
/* The same generator for ONE global set produced in pieces (one piece per rank, BASELINE config 5:
 * 10 M configurations, no 240 MB host-to-device copy): configuration i of this call is number
 * first + i of the set, so every split of the set over calls or ranks yields the same
 * configurations.  Blocks of `block` consecutive configurations whose block number is period - 1
 * modulo period are drawn from box_b, all others from box_a (period 0: box_a only); boxes are
 * {x_min, x_max, y_min, y_max, z_min, z_max}.  rc_alpha_box gives the reference's two boxes
 * (createAlphaBotConfigurations, Utils_rai.cu:111-136: hard = +-range/200, easy = +-10*range). */
int rc_conf_generate_device_mix(rc_ctx *ctx, rc_config *d_out, long first, long n, const float *box_a,
                                const float *box_b, int block, int period, uint64_t seed, void *stream);
int rc_alpha_box(int hard, float box[6]);
/* Order-preserving compaction: indices i < n whose flag equals `want` (1 = valid), ascending, and
 * their count.  Flags come from d_bitmask if it is not NULL, else from d_valid.  d_indices needs
 * room for n entries; d_count is one device word.  Asynchronous. */
int rc_compact(rc_ctx *ctx, const uint8_t *d_valid, const uint32_t *d_bitmask, long n, int want,
               uint32_t *d_indices, unsigned long long *d_count, void *stream);

/* ---- pipelined host calls (planner loops) ---------------------------------------
 * rc_check_host returns when the flags are in host memory, so the copy of the next batch cannot
 * start before the kernels of this one have finished (the reference has the same shape, plus its
 * 45-50 ms of per-call set-up, report 4.8).  rc_check_host_submit queues a batch and returns at
 * once; up to two batches are in flight per context: the configurations of batch k+1 travel on a
 * copy stream while the kernels of batch k run, the kernels of all batches run one after the other.
 * rc_check_host_wait(ticket) returns when the flags of that batch are in h_valid.  Buffers must stay
 * untouched between submit and wait; page-locked buffers are used in place, pageable ones are
 * staged.  Submitting a third batch first completes the oldest one.  Tickets count up from 0. */
int rc_check_host_submit(rc_ctx *ctx, const rc_config *h_confs, long n, uint8_t *h_valid, int mode,
                         int *ticket);
int rc_check_host_wait(rc_ctx *ctx, int ticket);

/* ---- one process per GPU: check + all-gather of the bitmask in one call -----------
 * The reference is single-GPU.  north_star shards the configurations over the GPUs of a box and
 * gathers only the validity bitmask (ncclAllGather); here the gather is fused INTO the check and
 * runs over peer-mapped memory (NVLink / NVSwitch) without a communicator call: every rank owns one
 * buffer of rc_gather_buffer_words(world, words_per_rank) 32-bit words, ZEROED once, mapped into the
 * address space of every rank (torch.distributed._symmetric_memory, cudaIpcOpenMemHandle, ...);
 * peer_buffers[r] is rank r's buffer as seen from the calling process (peer_buffers[rank] = the own
 * one).  rc_check_gather checks the rank's n configurations (n <= 32 * words_per_rank, mode
 * RC_MODE_CULLED) with the bitmask going straight into the rank's slot; the last CTA of the check
 * kernel to finish stores the slot into every peer's buffer, signals, and waits for the peers'
 * slots -- two launches per call, the same as rc_check.  All ranks must call it the same number of
 * times with the same world / words_per_rank (a collective).  After the call, in stream order,
 * *d_result points at world * words_per_rank words inside the own buffer: the bitmask of all ranks
 * (rank r's configurations start at bit 32 * words_per_rank * r).  It stays valid until the second
 * next call.  The call numbering lives in the buffer, so contexts may come and go.  A peer that never
 * arrives makes the wait give up after 4 s instead of hanging the device: the result is then
 * incomplete, and the only way to know is rc_gather_status (sticky `timeouts` word; also counted
 * in rc_stats.watchdog).  d_valid (optional) receives the rank's own flag bytes.
 *
 * Two-phase form: rc_check_gather_begin does everything but the wait (it returns, in stream order,
 * once the own slot has been delivered and signalled), rc_gather_end enqueues the wait for the
 * peers' slots of that call.  Work that does not need the peers' flags can sit between the two.
 * (It is also how the multi-rank path is tested on ONE device: begin for every emulated rank, then
 * end for every rank -- kernels that wait for each other must not share a GPU.) */
long rc_gather_buffer_words(int world, long words_per_rank);
int rc_check_gather(rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid,
                    void *const *peer_buffers, int rank, int world, long words_per_rank,
                    uint32_t **d_result, void *stream);
int rc_check_gather_begin(rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid,
                          void *const *peer_buffers, int rank, int world, long words_per_rank,
                          uint32_t **d_result, void *stream);
int rc_gather_end(rc_ctx *ctx, void *const *peer_buffers, int rank, int world, long words_per_rank,
                  void *stream);
/* Header words of the own gather buffer, read back synchronously: number of the last call this rank
 * delivered, and how many waits have timed out since the buffer was zeroed (must be 0). */
int rc_gather_status(rc_ctx *ctx, const void *own_buffer, int world, long words_per_rank,
                     uint32_t *calls, uint32_t *timeouts);

/* ---- sphere-approximated robot (BASELINE config 3) -----------------------------
 * The reference has no such path (test/CPU-Sphere-Test.cpp:8-18 only collides two unit spheres
 * with FCL), so the semantics are this library's: the robot is a set of spheres (x, y, z, r) in
 * its local frame; a configuration is valid iff no sphere, moved by the configuration, comes
 * within its radius of an obstacle triangle.  xyzr == NULL installs the default set: the smallest
 * enclosing sphere of every robot triangle. */
int rc_set_robot_spheres(rc_ctx *ctx, const float *xyzr, int n);
int rc_robot_spheres(const rc_ctx *ctx, float *xyzr_out, int *n); /* xyzr_out may be NULL */
int rc_check_spheres(rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid,
                     uint32_t *d_bitmask, void *stream);

/* ---- coplanar pairs -----------------------------------------------------------
 * A pair that passes both plane tests and whose planes is_coplanar_sep calls proportional
 * (narrow-phase/narrow-phase.cu:297-338) is not decided by the reference kernel: it sets
 * req_coplanar, prints "Error: require coplanar intersection" and goes on (:647-652, :680), so
 * such a pair never makes a configuration invalid.  RC_COPLANAR_SKIP (default) reproduces that,
 * flags stay bit-identical to the reference pipeline.  RC_COPLANAR_DECIDE decides the pair in the
 * common plane the way the reference's design model does (python/narrow.py:98-177: edge-edge
 * intersections, then containment): both triangles are projected along the dominant axis of the
 * robot triangle's normal, the nine edge pairs are tested (touching counts), then one vertex of
 * each triangle for containment in the other.  Applies to rc_check (both modes), rc_check_host,
 * rc_check_multi, rc_megakernel and rc_narrow_phase of this context; not to rc_check_spheres. */
#define RC_COPLANAR_SKIP 0
#define RC_COPLANAR_DECIDE 1
int rc_set_coplanar_mode(rc_ctx *ctx, int mode);
int rc_coplanar_mode(const rc_ctx *ctx);

/* ---- audits (closing rule) ------------------------------------------------
 * rc_hit_counts: the check of `mode` WITHOUT early exit; d_counts[i] = number of (robot triangle,
 * obstacle triangle) pairs of configuration i that the reference arithmetic reports -- all n_rob_t x
 * n_obs_t of them in RC_MODE_BRUTE, the ones the tree descent reaches in RC_MODE_CULLED; 0 for a
 * configuration the broad phase clears; 0xffffffff if the descent fell back to the exhaustive sweep
 * (pair stack overflow).  d_valid (optional) receives the flags.  Equal counts in both modes mean
 * the culling skipped no reported pair; equal counts against an independent CPU sweep mean no
 * compensating per-pair errors hide behind equal flags (a colliding alpha1.0 configuration has about
 * 100 hitting pairs).  Asynchronous; slower than rc_check (statistics kernels, no early exit). */
int rc_hit_counts(rc_ctx *ctx, const rc_config *d_confs, long n, int mode, uint8_t *d_valid,
                  uint32_t *d_counts, void *stream);
/* rc_audit: both modes on the same configurations; how many flags differ, how many hitting-pair
 * counts differ, and the first `cap` indices where either does (ascending).  Blocking. */
int rc_audit(rc_ctx *ctx, const rc_config *d_confs, long n, long *n_flags_differ, long *n_counts_differ,
             uint32_t *h_indices, long cap);
/* Device known-answer tests: the per-pair helpers the kernels are built from, run on n host-supplied
 * (robot triangle, obstacle triangle) pairs of 9 floats each (the role of the reference's host-side
 * unit tests, narrow-phase/test-narrow-phase.cu:7-626).  h_out32[32 * i ...]: 0-2 Nr, 3 dr, 4-6
 * signed distances of the obstacle vertices, 7-9 No, 10 do, 11-13 signed distances of the robot
 * vertices, 14-16 D, 17-19 O, 20-23 t_r01 t_r12 t_o01 t_o12 (14-23 unless the planes are coplanar).
 * h_verdict[i] = plane test O | plane test R << 1 | tail << 2, tail: 0 miss, 1 hit, 2 coplanar (what
 * the tail would say; the kernels only ask it when both plane tests pass).  rc_kat_la_solve: la_solve on (A1 A2 A3 A4 b1 b2) -> (x1 x2). */
int rc_kat_pairs(rc_ctx *ctx, const float *h_rob9, const float *h_obs9, int n, float *h_out32, int *h_verdict);
int rc_kat_la_solve(rc_ctx *ctx, const float *h_in6, int n, float *h_out2);

/* ---- measurement ------------------------------------------------------------------------- */
typedef struct {
    unsigned long long broad_cleared, sphere_cleared, narrow_valid, narrow_invalid;
    unsigned long long node_pairs, leaf_pairs, tri_pairs, tail_pairs, overflow, launches;
    unsigned long long node_rounds, leaf_rounds, tail_rounds; /* warp steps of each phase */
    unsigned long long shared_pieces;  /* parts of a descent one warp handed to another */
    unsigned long long watchdog, watchdog_where; /* waits that gave up: must be 0.  Inside the check kernels a
                                                  * give-up costs parallelism only, never a flag; a gather wait
                                                  * that gave up (bit 3) left *d_result incomplete */
    /* leaf rounds of RC_MODE_CULLED: of the tri_pairs looked at, those whose triangle boxes overlap, and
     * those that then pass the first plane test (tail_pairs pass both) */
    unsigned long long prefilter_pairs, plane1_pairs;
} rc_stats;
/* Statistics of rc_check calls since the last reset; counting is off by default. */
int rc_stats_enable(rc_ctx *ctx, int on);
int rc_stats_read(rc_ctx *ctx, rc_stats *out, int reset);
/* Per-configuration trace of the descent step of RC_MODE_CULLED, written while statistics are
 * enabled, four words per configuration i: [0] SM cycles its warp spent on it (0 if step 1 cleared
 * it), [1] low 32 bits of the device's nanosecond timer when it was decided, [2] node rounds << 16 |
 * leaf rounds, [3] tail rounds << 16 | deepest pair stack (rounds of the owning warp only).  The
 * caller zero-fills 4*n words of device memory; NULL switches the trace off. */
int rc_stats_trace(rc_ctx *ctx, uint32_t *d_trace);
/* CUDA-event time of the kernels of the last rc_check_host call, milliseconds. */
float rc_last_kernel_ms(const rc_ctx *ctx);
/* Sustained FP32 FMA rate of the device, TFLOP/s (2 flop per FMA), measured by a register-
 * resident FMA kernel: the denominator of the narrow-phase roofline. */
int rc_measure_fp32_peak(rc_ctx *ctx, double *tflops);

/* ---- host side: meshes, .conf files, configuration sets ------------------------------------ */
/* OBJ: `v x y z` and `f a b c ...` with `a//n` / `a/t/n` tokens, fan triangulation, 1-based ->
 * 0-based (Utils_rai.cu:139-320).  BYU (.g): header "parts verts polys edges", part ranges,
 * vertices, polygons terminated by a negative index (models/BYU_Model_Loader.cpp:33-89). */
int rc_mesh_load_obj(const char *path, rc_mesh **out);
int rc_mesh_load_byu(const char *path, rc_mesh **out);
int rc_mesh_from_arrays(const float *x, const float *y, const float *z, int n_v, const int *v1,
                        const int *v2, const int *v3, int n_t, rc_mesh **out);
int rc_mesh_save_obj(const rc_mesh *m, const char *path);
int rc_mesh_save_byu(const rc_mesh *m, const char *path);
int rc_mesh_counts(const rc_mesh *m, int *n_v, int *n_t);
const float *rc_mesh_x(const rc_mesh *m);
const float *rc_mesh_y(const rc_mesh *m);
const float *rc_mesh_z(const rc_mesh *m);
const int *rc_mesh_v1(const rc_mesh *m);
const int *rc_mesh_v2(const rc_mesh *m);
const int *rc_mesh_v3(const rc_mesh *m);
/* 1 -> 4 midpoint subdivision, `levels` times (BASELINE config 5: subdivided obstacle). */
int rc_mesh_subdivide(const rc_mesh *m, int levels, rc_mesh **out);
int rc_mesh_aabb(const rc_mesh *m, rc_aabb *out);
void rc_mesh_free(rc_mesh *m);

/* .env / .query of the motion-planning benchmark the meshes come from (models/alpha1.0/alpha.env,
 * alpha.query).  The environment file names the bodies and places the passive ones in the world
 * (alpha.env:13 puts the obstacle at 11.0899 9.00906 3.24427, no rotation); the query file holds
 * start and goal.  The reference ignores both and checks the OBJ files as they are, so the default
 * everywhere in this library is the unplaced mesh; a caller who wants the benchmark's world frame
 * reads the file and applies the placement with rc_mesh_place before rc_create_from_meshes.
 * Rotations of a placement are degrees about x, y, z (applied as Rz Ry Rx, the reference's own
 * matrix); the query's three rotations are fractions of a full turn (-> roll, pitch, yaw). */
#define RC_ENV_MAX_BODIES 16
typedef struct {
    char file[256];        /* geometry file as written, relative to the .env file      */
    int active;            /* 1 = robot (movable body), 0 = passive obstacle             */
    float color[4];        /* -c(r g b a); 0 0 0 1 when absent                           */
    int has_placement;     /* passive bodies: x y z rx ry rz follow the file name        */
    float x, y, z, rx, ry, rz;
} rc_env_body;
typedef struct {
    int has_box;           /* "Boundary Box [lo:hi ; lo:hi ; lo:hi]" present            */
    float box_lo[3], box_hi[3];
    int n_bodies;
    rc_env_body bodies[RC_ENV_MAX_BODIES];
} rc_env;
int rc_env_read(const char *path, rc_env *out);
int rc_query_read(const char *path, rc_config **confs, long *n); /* free *confs with rc_free */
int rc_mesh_place(const rc_mesh *m, float x, float y, float z, float rx_deg, float ry_deg, float rz_deg,
                  rc_mesh **out);

/* .conf: one header line, then `x y z pitch yaw roll tag` per line
 * (readConfigurationFromFile / writeConfigurationToFileTagged, Utils_rai.cu:11-97).
 * rc_conf_read allocates *confs (and *tags if tags != NULL) with malloc; free with rc_free. */
int rc_conf_read(const char *path, rc_config **confs, uint8_t **tags, long *n);
int rc_conf_write(const char *path, const rc_config *confs, const uint8_t *valid, long n);
void rc_free(void *p);
/* generateConfs (Utils_rai.cu:322-348) with a caller-supplied seed instead of
 * std::random_device: mt19937, draw order x, y, z, pitch, yaw, roll per configuration. */
int rc_conf_generate(rc_config *out, long n, float x_min, float x_max, float y_min, float y_max,
                     float z_min, float z_max, uint32_t seed);
/* createAlphaBotConfigurations (Utils_rai.cu:111-136): hard = +-range/200, easy = +-10*range. */
int rc_conf_generate_alpha(rc_config *out, long n, int hard, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif /* ROBOCHECK_H */
