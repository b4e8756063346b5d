// rc_api.cu -- the C ABI (include/robocheck.h): context, fused entry point, stage wrappers.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../host/rc_bvh.h"
#include "../host/rc_mesh.h"
#include "rc_internal.h"

namespace rc {

int set_error(rc_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->last_error = msg;
    return code;
}

int check_cuda(rc_ctx *ctx, cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return RC_OK;
    return set_error(ctx, RC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

cudaError_t launch_fma_peak(float *d_out, int grid, int iters, cudaStream_t stream);

namespace {

// culling boxes are inflated by this much on each side (scene extent is ~90 units)
constexpr float kBoxMargin = 1e-3f;
// groups of robot vertices whose rotated boxes bound the vertex AABB in classify_kernel
constexpr int kBoundBoxes = 32;  // = rc::kClassifyBoxes (rc_culled.cuh)

template <class T>
cudaError_t dev_alloc(rc_ctx *ctx, T **p, size_t count)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count * sizeof(T), 16));
    if (e == cudaSuccess) {
        ctx->allocations.push_back(q);
        cudaMemset(q, 0, std::max<size_t>(count * sizeof(T), 16));
    }
    *p = static_cast<T *>(q);
    return e;
}

template <class T>
cudaError_t upload(rc_ctx *ctx, T **p, const std::vector<T> &h, size_t padded_count = 0)
{
    cudaError_t e = dev_alloc(ctx, p, std::max(padded_count, h.size()));
    if (e != cudaSuccess) return e;
    if (h.empty()) return cudaSuccess;
    return cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

// k-d partition of the robot's vertices into at most `want` groups (median split on the longest
// axis); returns (centre, half extent) per group, extents measured from the rounded centre and
// rounded up.  Every vertex -- referenced by a triangle or not -- is inside exactly one box, so
// the union of the rotated boxes contains the exact vertex AABB of broadPhaseFusedKernel_sep.
std::vector<float4> vertex_bound_boxes(const float *x, const float *y, const float *z, int n, int want)
{
    std::vector<float4> out;
    if (n <= 0) return out;
    std::vector<int> idx((size_t)n);
    for (int i = 0; i < n; i++) idx[i] = i;
    struct Range { int lo, hi; };
    std::vector<Range> groups{{0, n}};
    const float *c[3] = {x, y, z};
    auto extent = [&](const Range &r, double lo[3], double hi[3]) {
        for (int k = 0; k < 3; k++) { lo[k] = 1e300; hi[k] = -1e300; }
        for (int i = r.lo; i < r.hi; i++)
            for (int k = 0; k < 3; k++) {
                lo[k] = std::min(lo[k], (double)c[k][idx[i]]);
                hi[k] = std::max(hi[k], (double)c[k][idx[i]]);
            }
    };
    while ((int)groups.size() < want) {
        // split the group with the longest box edge
        int best = -1, best_ax = 0;
        double best_len = 0;
        for (size_t g = 0; g < groups.size(); g++) {
            if (groups[g].hi - groups[g].lo < 2) continue;
            double lo[3], hi[3];
            extent(groups[g], lo, hi);
            for (int k = 0; k < 3; k++)
                if (hi[k] - lo[k] > best_len) { best_len = hi[k] - lo[k]; best = (int)g; best_ax = k; }
        }
        if (best < 0) break;
        const Range r = groups[best];
        const int mid = (r.lo + r.hi) / 2;
        const float *key = c[best_ax];
        std::nth_element(idx.begin() + r.lo, idx.begin() + mid, idx.begin() + r.hi,
                         [&](int a, int b) { return key[a] < key[b] || (key[a] == key[b] && a < b); });
        groups[best] = Range{r.lo, mid};
        groups.push_back(Range{mid, r.hi});
    }
    for (const Range &r : groups) {
        double lo[3], hi[3];
        extent(r, lo, hi);
        float cen[3], ext[3];
        for (int k = 0; k < 3; k++) {
            cen[k] = (float)(0.5 * (lo[k] + hi[k]));
            const double e = std::max(hi[k] - (double)cen[k], (double)cen[k] - lo[k]);
            ext[k] = (float)(e * (1.0 + 1e-6)) + 1e-5f;
        }
        out.push_back(make_float4(cen[0], cen[1], cen[2], 0.f));
        out.push_back(make_float4(ext[0], ext[1], ext[2], 0.f));
    }
    return out;
}

struct GatheredTris {
    std::vector<float4> a, b, c;
};

GatheredTris gather(const std::vector<float> &x, const std::vector<float> &y,
                    const std::vector<float> &z, const std::vector<int> &v1,
                    const std::vector<int> &v2, const std::vector<int> &v3,
                    const std::vector<int> *order)
{
    GatheredTris g;
    const size_t n = v1.size();
    g.a.resize(n); g.b.resize(n); g.c.resize(n);
    for (size_t p = 0; p < n; p++) {
        const size_t t = order ? (size_t)(*order)[p] : p;
        g.a[p] = make_float4(x[v1[t]], y[v1[t]], z[v1[t]], 0.f);
        g.b[p] = make_float4(x[v2[t]], y[v2[t]], z[v2[t]], 0.f);
        g.c[p] = make_float4(x[v3[t]], y[v3[t]], z[v3[t]], 0.f);
    }
    return g;
}

int upload_obstacle(rc_ctx *ctx, const GatheredTris &g, float4 **a, float4 **b, float4 **c, float **d)
{
    float4 *ra = nullptr, *rb = nullptr, *rcv = nullptr;
    const size_t n = g.a.size();
    cudaError_t e;
    if ((e = upload(ctx, &ra, g.a)) || (e = upload(ctx, &rb, g.b)) || (e = upload(ctx, &rcv, g.c)) ||
        (e = dev_alloc(ctx, a, n)) || (e = dev_alloc(ctx, b, n)) || (e = dev_alloc(ctx, c, n)) ||
        (e = dev_alloc(ctx, d, (n + 3) & ~(size_t)3)))
        return check_cuda(ctx, e, "obstacle upload");
    // dev_alloc clears with cudaMemset on the legacy default stream, which does NOT order against the
    // context's non-blocking stream: without this the clear of a large table could land after the
    // kernel below has filled it (seen once in a few hundred context builds on the 16,128-triangle
    // obstacle: part of the table read back as zeros, collisions missed)
    if ((e = cudaDeviceSynchronize())) return check_cuda(ctx, e, "obstacle upload sync");
    if ((e = launch_obstacle_planes(ra, rb, rcv, *a, *b, *c, *d, (int)n, ctx->own_stream)))
        return check_cuda(ctx, e, "obstacle planes");
    return check_cuda(ctx, cudaStreamSynchronize(ctx->own_stream), "obstacle planes sync");
}

// NULL is CUDA's legacy default stream (which is also torch's default stream), NOT the
// context's private stream: work must land on the stream the caller times and orders.
cudaStream_t stream_of(rc_ctx *, void *stream) { return static_cast<cudaStream_t>(stream); }

// shared argument check of the entry points that launch a check
int check_args(rc_ctx *ctx, long n, const void *confs, int mode)
{
    if (!ctx || n < 0 || (n > 0 && !confs)) return RC_ERR_INVALID;
    if (mode != RC_MODE_CULLED && mode != RC_MODE_BRUTE)
        return set_error(ctx, RC_ERR_INVALID, "mode must be RC_MODE_CULLED or RC_MODE_BRUTE");
    if (n > 0xffffffffL) return set_error(ctx, RC_ERR_INVALID, "more than 2^32 - 1 configurations in one call");
    return RC_OK;
}

// The scratch of a context (work lists, their self-resetting counters, the queue word) is shared by
// every entry point, and entry points enqueue on different streams (the caller's, or the context's
// own).  Device work of one call must not overlap the next call's on another stream: every launch
// that touches the scratch records an event behind itself, and a launch on a different stream waits
// for it first.  Calls on one stream pay one cudaEventRecord (about a microsecond of host time).
int scratch_acquire(rc_ctx *ctx, cudaStream_t s)
{
    if (ctx->scratch_used && ctx->scratch_stream != s) {
        const cudaError_t e = cudaStreamWaitEvent(s, ctx->ev_scratch, 0);
        if (e != cudaSuccess) return check_cuda(ctx, e, "scratch hand-over between streams");
    }
    return RC_OK;
}
void scratch_release(rc_ctx *ctx, cudaStream_t s)
{
    cudaEventRecord(ctx->ev_scratch, s);
    ctx->scratch_stream = s;
    ctx->scratch_used = true;
}

int ensure_work_list(rc_ctx *ctx, long n)
{
    if (n <= ctx->work_cap) return RC_OK;
    cudaDeviceSynchronize();  // an earlier asynchronous check may still use the old list
    if (ctx->d_work_idx) cudaFree(ctx->d_work_idx);
    if (ctx->d_work_mat) cudaFree(ctx->d_work_mat);
    if (ctx->d_feed_slots) cudaFree(ctx->d_feed_slots);
    ctx->d_work_idx = nullptr; ctx->d_work_mat = nullptr; ctx->d_feed_slots = nullptr; ctx->work_cap = 0;
    const long cap = std::max(n, 1L << 17);
    cudaError_t e;
    if ((e = cudaMalloc(&ctx->d_work_idx, (size_t)cap * kWorkBins * sizeof(uint32_t))) ||
        (e = cudaMalloc(&ctx->d_work_mat, (size_t)cap * 3 * sizeof(float4))))
        return check_cuda(ctx, e, "work list");
    ctx->work_cap = cap;
    return RC_OK;
}

// Lists of the single-launch form (rc_culled.cuh: FeedArgs), sized like the work list, allocated the
// first time that form is used.  Zeroed once: a slot counts as filled when its tag equals the call
// number, which starts at 1.
int ensure_feed(rc_ctx *ctx)
{
    cudaError_t e;
    if (!ctx->d_feed_ctr && (e = cudaMalloc(&ctx->d_feed_ctr, 1024 * sizeof(uint32_t)))) return check_cuda(ctx, e, "feed counters");
    if (!ctx->d_feed_slots) {
        const size_t bytes = (size_t)ctx->work_cap * kWorkBins * sizeof(uint2);
        if ((e = cudaMalloc(&ctx->d_feed_slots, bytes)) || (e = cudaMemset(ctx->d_feed_slots, 0, bytes)))
            return check_cuda(ctx, e, "feed lists");
    }
    return RC_OK;
}

int ensure_staging(rc_ctx *ctx, long n)
{
    if (n <= ctx->stage_capacity) return RC_OK;
    if (ctx->h_pinned_confs) cudaFreeHost(ctx->h_pinned_confs);
    if (ctx->h_pinned_valid) cudaFreeHost(ctx->h_pinned_valid);
    if (ctx->d_stage_confs) cudaFree(ctx->d_stage_confs);
    if (ctx->d_stage_valid) cudaFree(ctx->d_stage_valid);
    if (ctx->d_stage_mask) cudaFree(ctx->d_stage_mask);
    if (ctx->h_pinned_mask) cudaFreeHost(ctx->h_pinned_mask);
    ctx->h_pinned_confs = nullptr; ctx->h_pinned_valid = nullptr; ctx->h_pinned_mask = nullptr;
    ctx->d_stage_confs = nullptr; ctx->d_stage_valid = nullptr; ctx->d_stage_mask = nullptr;
    ctx->stage_capacity = 0;
    const long cap = std::max(n, 1L << 16);
    cudaError_t e;
    if ((e = cudaMallocHost(&ctx->h_pinned_confs, (size_t)cap * sizeof(rc_config))) ||
        (e = cudaMallocHost(&ctx->h_pinned_valid, (size_t)cap)) ||
        (e = cudaMalloc(&ctx->d_stage_confs, (size_t)cap * sizeof(rc_config))) ||
        (e = cudaMalloc(&ctx->d_stage_valid, (size_t)cap)) ||
        (e = cudaMalloc(&ctx->d_stage_mask, (size_t)((cap + 31) / 32) * sizeof(uint32_t))) ||
        (e = cudaMallocHost(&ctx->h_pinned_mask, (size_t)((cap + 31) / 32) * sizeof(uint32_t))))
        return check_cuda(ctx, e, "staging buffers");
    ctx->stage_capacity = cap;
    return RC_OK;
}

int ensure_lane(rc_ctx *ctx, PipeLane &l, long n)
{
    cudaError_t e;
    if (!l.copied && ((e = cudaEventCreateWithFlags(&l.copied, cudaEventDisableTiming)) ||
                      (e = cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming))))
        return check_cuda(ctx, e, "pipeline events");
    if (n <= l.cap) return RC_OK;
    if (l.d_confs) cudaFree(l.d_confs);
    if (l.d_valid) cudaFree(l.d_valid);
    if (l.h_confs) cudaFreeHost(l.h_confs);
    if (l.h_valid) cudaFreeHost(l.h_valid);
    l.d_confs = nullptr; l.d_valid = nullptr; l.h_confs = nullptr; l.h_valid = nullptr; l.cap = 0;
    const long cap = std::max(n, 1L << 16);
    if ((e = cudaMalloc(&l.d_confs, (size_t)cap * sizeof(rc_config))) || (e = cudaMalloc(&l.d_valid, (size_t)cap)) ||
        (e = cudaMallocHost(&l.h_confs, (size_t)cap * sizeof(rc_config))) ||
        (e = cudaMallocHost(&l.h_valid, (size_t)cap)))
        return check_cuda(ctx, e, "pipeline staging buffers");
    l.cap = cap;
    return RC_OK;
}

void free_lane(PipeLane &l)
{
    if (l.d_confs) cudaFree(l.d_confs);
    if (l.d_valid) cudaFree(l.d_valid);
    if (l.h_confs) cudaFreeHost(l.h_confs);
    if (l.h_valid) cudaFreeHost(l.h_valid);
    if (l.copied) cudaEventDestroy(l.copied);
    if (l.done) cudaEventDestroy(l.done);
    l = PipeLane();
}

int finish_lane(rc_ctx *ctx, PipeLane &l)
{
    if (!l.busy) return RC_OK;
    const cudaError_t e = cudaEventSynchronize(l.done);
    l.busy = false;
    if (e != cudaSuccess) return check_cuda(ctx, e, "rc_check_host_wait");
    if (l.user_valid) std::memcpy(l.user_valid, l.h_valid, (size_t)l.n);
    l.user_valid = nullptr;
    return RC_OK;
}

}  // namespace
}  // namespace rc

using namespace rc;

extern "C" {

const char *rc_strerror(int code)
{
    switch (code) {
    case RC_OK: return "ok";
    case RC_ERR_INVALID: return "invalid argument";
    case RC_ERR_CUDA: return "CUDA runtime error";
    case RC_ERR_NOMEM: return "out of memory";
    case RC_ERR_IO: return "cannot open file";
    case RC_ERR_FORMAT: return "malformed file";
    case RC_ERR_UNSUPPORTED: return "mesh too large for this path";
    case RC_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
    default: return "unknown error";
    }
}

const char *rc_last_error(const rc_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
int rc_device(const rc_ctx *ctx) { return ctx ? ctx->device : -1; }

int rc_create(rc_ctx **out, const float *rob_x, const float *rob_y, const float *rob_z, int n_rob_v,
              const int *rob_v1, const int *rob_v2, const int *rob_v3, int n_rob_t,
              const float *obs_x, const float *obs_y, const float *obs_z, int n_obs_v,
              const int *obs_v1, const int *obs_v2, const int *obs_v3, int n_obs_t, int device)
{
    if (!out) return RC_ERR_INVALID;
    *out = nullptr;
    if (n_rob_v <= 0 || n_rob_t <= 0 || n_obs_v <= 0 || n_obs_t <= 0 || !rob_x || !rob_y || !rob_z ||
        !rob_v1 || !rob_v2 || !rob_v3 || !obs_x || !obs_y || !obs_z || !obs_v1 || !obs_v2 || !obs_v3)
        return RC_ERR_INVALID;
    for (int t = 0; t < n_rob_t; t++)
        if (rob_v1[t] < 0 || rob_v1[t] >= n_rob_v || rob_v2[t] < 0 || rob_v2[t] >= n_rob_v ||
            rob_v3[t] < 0 || rob_v3[t] >= n_rob_v)
            return RC_ERR_INVALID;
    for (int t = 0; t < n_obs_t; t++)
        if (obs_v1[t] < 0 || obs_v1[t] >= n_obs_v || obs_v2[t] < 0 || obs_v2[t] >= n_obs_v ||
            obs_v3[t] < 0 || obs_v3[t] >= n_obs_v)
            return RC_ERR_INVALID;
    if (n_rob_v > 65535) return RC_ERR_UNSUPPORTED;  // robot triangle indices are packed in 16 bits

    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return RC_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) return RC_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return RC_ERR_CUDA;

    rc_ctx *ctx = new (std::nothrow) rc_ctx();
    if (!ctx) return RC_ERR_NOMEM;
    ctx->device = device;
    {
        auto knob = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
        ctx->knob_warps = knob("RC_CULLED_WARPS", 0);
        ctx->knob_seed_depth = knob("RC_SEED_DEPTH", -1);
        ctx->knob_tail = knob("RC_TAIL_THRESH", 0);
        ctx->knob_leaf = knob("RC_LEAF_THRESH", 0);
        ctx->knob_chunk = knob("RC_CHUNK", 0);
        ctx->knob_stack_cap = knob("RC_STACK_CAP", 0);
        ctx->knob_global_obstacle = getenv("RC_FORCE_GLOBAL_OBSTACLE") != nullptr;
        ctx->knob_share = knob("RC_SHARE", 1);
        ctx->knob_host_direct = knob("RC_HOST_DIRECT", 0);
        ctx->knob_pdl = knob("RC_PDL", 1);
        ctx->knob_bound = knob("RC_BOUND", 1);
        ctx->knob_mail_slots = knob("RC_MAIL_SLOTS", 4);
        ctx->knob_donate_min = knob("RC_DONATE_MIN", 40);
        ctx->knob_donate_leaf = knob("RC_DONATE_LEAF", 16);
        ctx->knob_feed = knob("RC_FEED", 1);
        ctx->knob_feed_cls = knob("RC_FEED_CLS", 0);
        ctx->knob_split_max = knob("RC_SPLIT_MAX", 64);
        ctx->knob_split_load = knob("RC_SPLIT_LOAD", 8);
        ctx->knob_host_chunks = knob("RC_HOST_CHUNKS", 0);
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    ctx->rob_x.assign(rob_x, rob_x + n_rob_v); ctx->rob_y.assign(rob_y, rob_y + n_rob_v);
    ctx->rob_z.assign(rob_z, rob_z + n_rob_v);
    ctx->rob_v1.assign(rob_v1, rob_v1 + n_rob_t); ctx->rob_v2.assign(rob_v2, rob_v2 + n_rob_t);
    ctx->rob_v3.assign(rob_v3, rob_v3 + n_rob_t);
    ctx->obs_x.assign(obs_x, obs_x + n_obs_v); ctx->obs_y.assign(obs_y, obs_y + n_obs_v);
    ctx->obs_z.assign(obs_z, obs_z + n_obs_v);
    ctx->obs_v1.assign(obs_v1, obs_v1 + n_obs_t); ctx->obs_v2.assign(obs_v2, obs_v2 + n_obs_t);
    ctx->obs_v3.assign(obs_v3, obs_v3 + n_obs_t);

    DeviceTables &t = ctx->t;
    t.n_rob_v = n_rob_v; t.n_rob_t = n_rob_t; t.n_obs_v = n_obs_v; t.n_obs_t = n_obs_t;
    t.n_rob_v_pad = (n_rob_v + 31) / 32 * 32;

    int rcode = RC_OK;
    cudaError_t e;
    do {
        if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) ||
            (e = cudaStreamCreateWithFlags(&ctx->own_stream2, cudaStreamNonBlocking)) ||
            (e = cudaEventCreate(&ctx->ev0)) || (e = cudaEventCreate(&ctx->ev1)) ||
            (e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) ||
            (e = cudaEventCreateWithFlags(&ctx->ev_scratch, cudaEventDisableTiming))) {
            rcode = check_cuda(ctx, e, "stream/event");
            break;
        }
        // obstacle AABB: generateAABBBaseline(x, y, z), generate-AABB.cu:183-201
        {
            float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (int v = 0; v < n_obs_v; v++) {
                lo[0] = std::min(lo[0], obs_x[v]); hi[0] = std::max(hi[0], obs_x[v]);
                lo[1] = std::min(lo[1], obs_y[v]); hi[1] = std::max(hi[1], obs_y[v]);
                lo[2] = std::min(lo[2], obs_z[v]); hi[2] = std::max(hi[2], obs_z[v]);
            }
            t.obs_box = rc_aabb{lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
        }
        // robot bounding sphere (local frame), radius rounded up
        {
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            for (int v = 0; v < n_rob_v; v++) {
                lo[0] = std::min(lo[0], (double)rob_x[v]); hi[0] = std::max(hi[0], (double)rob_x[v]);
                lo[1] = std::min(lo[1], (double)rob_y[v]); hi[1] = std::max(hi[1], (double)rob_y[v]);
                lo[2] = std::min(lo[2], (double)rob_z[v]); hi[2] = std::max(hi[2], (double)rob_z[v]);
            }
            for (int k = 0; k < 3; k++) t.rob_center[k] = (float)(0.5 * (lo[k] + hi[k]));
            for (int k = 0; k < 3; k++) {
                t.rob_box_c[k] = t.rob_center[k];
                const double e = std::max(hi[k] - (double)t.rob_center[k], (double)t.rob_center[k] - lo[k]);
                t.rob_box_e[k] = (float)(e * (1.0 + 1e-6)) + 1e-4f;
            }
            double r2 = 0;
            for (int v = 0; v < n_rob_v; v++) {
                const double dx = rob_x[v] - t.rob_center[0], dy = rob_y[v] - t.rob_center[1],
                             dz = rob_z[v] - t.rob_center[2];
                r2 = std::max(r2, dx * dx + dy * dy + dz * dz);
            }
            t.rob_radius = (float)(std::sqrt(r2) * (1.0 + 1e-6)) + 1e-4f;
        }
        // robot vertices (padded with copies of vertex 0) and packed triangles
        std::vector<float4> rv((size_t)t.n_rob_v_pad);
        for (int v = 0; v < t.n_rob_v_pad; v++) {
            const int s = v < n_rob_v ? v : 0;
            rv[v] = make_float4(rob_x[s], rob_y[s], rob_z[s], 0.f);
        }
        std::vector<rc_vec3> raos((size_t)n_rob_v);
        for (int v = 0; v < n_rob_v; v++) raos[v] = rc_vec3{rob_x[v], rob_y[v], rob_z[v]};
        std::vector<uint2> rt((size_t)n_rob_t);
        for (int k = 0; k < n_rob_t; k++)
            rt[k] = make_uint2((uint32_t)rob_v1[k] | ((uint32_t)rob_v2[k] << 16), (uint32_t)rob_v3[k]);
        // BVHs.  Every tree box is inflated by box_margin; a pair of boxes is culled only when it is
        // separated on an axis after that inflation, so the margin has to stay well above the rounding
        // of the box test itself, which grows with the size of the coordinates: about 5e-7 x |coordinate|
        // (four FMAs on values of that size, boxes stored in float).  Near-contact configurations -- the
        // only ones whose pairs are ever looked at -- keep the robot within its own diameter of the
        // obstacle, hence the scene extent below.  kBoxMargin (1e-3) holds up to extents of 250 units
        // (alpha1.0: about 130); larger scenes get a proportionally larger margin.
        {
            double ext = 0;
            for (int v = 0; v < n_obs_v; v++)
                ext = std::max(ext, (double)std::max(std::fabs(obs_x[v]), std::max(std::fabs(obs_y[v]), std::fabs(obs_z[v]))));
            ext += 2.0 * (double)t.rob_radius;
            for (int k = 0; k < 3; k++) ext = std::max(ext, (double)std::fabs(t.rob_center[k]) + t.rob_radius);
            t.box_margin = std::max(kBoxMargin, (float)(4e-6 * ext));
        }
        const Bvh ob = build_bvh(obs_x, obs_y, obs_z, obs_v1, obs_v2, obs_v3, n_obs_t, t.box_margin);
        Bvh rb = build_bvh(rob_x, rob_y, rob_z, rob_v1, rob_v2, rob_v3, n_rob_t, t.box_margin);
        t.n_rob_nodes = rb.n_nodes; t.n_obs_nodes = ob.n_nodes;
        // An obstacle that does not fit into shared memory stays in global memory; it is then much
        // finer than the robot, whose tree gets one triangle per leaf (node indices are 11 bits there)
        t.gobs = !culled_obstacle_in_shared(ctx);
        if (t.gobs && 2 * n_rob_t - 1 < (1 << 11)) {
            rb = build_bvh(rob_x, rob_y, rob_z, rob_v1, rob_v2, rob_v3, n_rob_t, t.box_margin, 1);
            t.rob_leaf = 1;
            t.n_rob_nodes = rb.n_nodes;
        }
        for (int d = 0; d < 4; d++) {
            const std::vector<int> cr = rb.cut(d), co = ob.cut(d);
            t.n_seed_r[d] = (int)std::min<size_t>(cr.size(), 8);
            t.n_seed_o[d] = (int)std::min<size_t>(co.size(), 8);
            for (int q = 0; q < 8; q++) {
                t.seed_r[d][q] = (uint32_t)(q < (int)cr.size() ? cr[q] : 0);
                t.seed_o[d][q] = (uint32_t)(q < (int)co.size() ? co[q] : 0);
            }
        }
        {
            const std::vector<int> cr = rb.cut(4), co = ob.cut(4);
            t.n_seed_r4 = (int)std::min<size_t>(cr.size(), 16);
            t.n_seed_o4 = (int)std::min<size_t>(co.size(), 16);
            for (int q = 0; q < 16; q++) {
                t.seed_r4[q] = (uint32_t)(q < (int)cr.size() ? cr[q] : 0);
                t.seed_o4[q] = (uint32_t)(q < (int)co.size() ? co[q] : 0);
            }
        }
        // plane of every robot triangle in the robot frame, leaf order: n = (b - a) x (c - b), s = n . a
        std::vector<float4> rpl((size_t)n_rob_t);
        for (int p = 0; p < n_rob_t; p++) {
            const int k = rb.order[p];
            const double a[3] = {rob_x[rob_v1[k]], rob_y[rob_v1[k]], rob_z[rob_v1[k]]};
            const double b[3] = {rob_x[rob_v2[k]], rob_y[rob_v2[k]], rob_z[rob_v2[k]]};
            const double c3[3] = {rob_x[rob_v3[k]], rob_y[rob_v3[k]], rob_z[rob_v3[k]]};
            const double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c3[0] - b[0], c3[1] - b[1], c3[2] - b[2]};
            const double nn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            rpl[p] = make_float4((float)nn[0], (float)nn[1], (float)nn[2],
                                 (float)(nn[0] * a[0] + nn[1] * a[1] + nn[2] * a[2]));
        }
        std::vector<uint2> rts((size_t)n_rob_t);
        for (int p = 0; p < n_rob_t; p++) rts[p] = rt[rb.order[p]];
        std::vector<float4> rbn((size_t)rb.n_nodes * 2), obn((size_t)ob.n_nodes * 2);
        std::memcpy(rbn.data(), rb.nodes.data(), rb.nodes.size() * sizeof(float));
        std::memcpy(obn.data(), ob.nodes.data(), ob.nodes.size() * sizeof(float));

        const size_t rt_pad = (size_t)n_rob_t + (n_rob_t & 1);
        const std::vector<float4> bound = vertex_bound_boxes(rob_x, rob_y, rob_z, n_rob_v, kBoundBoxes);
        // fewer groups than boxes (tiny meshes): repeat the first one, the union does not change
        t.n_rob_bound = bound.empty() ? 0 : kBoundBoxes;
        for (int q = 0; q < kBoundBoxes && !bound.empty(); q++) {
            const size_t src = (size_t)q < bound.size() / 2 ? (size_t)q : 0;
            t.h_rob_bound[2 * q] = bound[2 * src];
            t.h_rob_bound[2 * q + 1] = bound[2 * src + 1];
        }
        if ((e = upload(ctx, &t.rob_v, rv)) || (e = upload(ctx, &t.rob_aos, raos)) ||
            (e = upload(ctx, &t.rob_t, rt, rt_pad)) || (e = upload(ctx, &t.rob_t_sorted, rts, rt_pad)) ||
            (e = upload(ctx, &t.rob_bvh, rbn)) || (e = upload(ctx, &t.obs_bvh, obn)) ||
            (e = upload(ctx, &t.rob_plane, rpl)) ||
            (e = dev_alloc(ctx, &ctx->d_counters, RC_NUM_COUNTERS)) ||
            (e = dev_alloc(ctx, &ctx->d_queues, kHostChunks)) ||
            (e = dev_alloc(ctx, &ctx->d_work_count, kWorkCounters)) ||
            (e = dev_alloc(ctx, &ctx->d_obs_box, 1)) ||
            (e = cudaMemcpy(ctx->d_obs_box, &t.obs_box, sizeof(rc_aabb), cudaMemcpyHostToDevice))) {
            rcode = check_cuda(ctx, e, "table upload");
            break;
        }
        const GatheredTris g = gather(ctx->obs_x, ctx->obs_y, ctx->obs_z, ctx->obs_v1, ctx->obs_v2,
                                      ctx->obs_v3, nullptr);
        const GatheredTris gs = gather(ctx->obs_x, ctx->obs_y, ctx->obs_z, ctx->obs_v1, ctx->obs_v2,
                                       ctx->obs_v3, &ob.order);
        if ((rcode = upload_obstacle(ctx, g, &t.obs_a, &t.obs_b, &t.obs_c, &t.obs_d)) != RC_OK) break;
        if ((rcode = upload_obstacle(ctx, gs, &t.obs_a_sorted, &t.obs_b_sorted, &t.obs_c_sorted,
                                     &t.obs_d_sorted)) != RC_OK)
            break;
    } while (0);

    // every table and every cleared scratch word is in place before the first launch, whatever stream
    // (blocking or not) the caller brings
    if (rcode == RC_OK && (e = cudaDeviceSynchronize()) != cudaSuccess) rcode = check_cuda(ctx, e, "rc_create sync");
    if (rcode != RC_OK) {  // the context is discarded; the code says what failed (never prints)
        rc_destroy(ctx);
        return rcode;
    }
    *out = ctx;
    return RC_OK;
}

int rc_create_from_meshes(rc_ctx **out, const rc_mesh *r, const rc_mesh *o, int device)
{
    if (!r || !o) return RC_ERR_INVALID;
    return rc_create(out, r->x.data(), r->y.data(), r->z.data(), (int)r->x.size(), r->v1.data(),
                     r->v2.data(), r->v3.data(), (int)r->v1.size(), o->x.data(), o->y.data(),
                     o->z.data(), (int)o->x.size(), o->v1.data(), o->v2.data(), o->v3.data(),
                     (int)o->v1.size(), device);
}

void rc_destroy(rc_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (void *p : ctx->allocations) cudaFree(p);
    for (cudaEvent_t ev : ctx->ev_wave)
        if (ev) cudaEventDestroy(ev);
    if (ctx->h_pinned_confs) cudaFreeHost(ctx->h_pinned_confs);
    if (ctx->h_pinned_valid) cudaFreeHost(ctx->h_pinned_valid);
    if (ctx->d_stage_confs) cudaFree(ctx->d_stage_confs);
    if (ctx->d_stage_valid) cudaFree(ctx->d_stage_valid);
    if (ctx->d_stage_mask) cudaFree(ctx->d_stage_mask);
    if (ctx->h_pinned_mask) cudaFreeHost(ctx->h_pinned_mask);
    if (ctx->d_block_counts) cudaFree(ctx->d_block_counts);
    if (ctx->d_work_idx) cudaFree(ctx->d_work_idx);
    if (ctx->d_work_mat) cudaFree(ctx->d_work_mat);
    if (ctx->d_feed_slots) cudaFree(ctx->d_feed_slots);
    if (ctx->d_feed_ctr) cudaFree(ctx->d_feed_ctr);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    free_lane(ctx->pipe[0]); free_lane(ctx->pipe[1]);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_scratch) cudaEventDestroy(ctx->ev_scratch);
    if (ctx->own_stream2) cudaStreamDestroy(ctx->own_stream2);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int rc_mesh_sizes(const rc_ctx *ctx, int *n_rob_v, int *n_rob_t, int *n_obs_v, int *n_obs_t)
{
    if (!ctx) return RC_ERR_INVALID;
    if (n_rob_v) *n_rob_v = ctx->t.n_rob_v;
    if (n_rob_t) *n_rob_t = ctx->t.n_rob_t;
    if (n_obs_v) *n_obs_v = ctx->t.n_obs_v;
    if (n_obs_t) *n_obs_t = ctx->t.n_obs_t;
    return RC_OK;
}

int rc_obstacle_aabb(const rc_ctx *ctx, rc_aabb *out)
{
    if (!ctx || !out) return RC_ERR_INVALID;
    *out = ctx->t.obs_box;
    return RC_OK;
}

constexpr long kFeedMinHost = 1L << 16;  // rc_check_host: smallest batch read in place (below: copy + two launches)
constexpr long kFeedMax = 1L << 22;  // configurations per call of the single-launch form (64 B of list per configuration)

static int check_on_stream(rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid, uint32_t *d_bitmask,
                           int mode, cudaStream_t s, const rc::CheckExtras *extras, const char *what)
{
    if (mode == RC_MODE_CULLED) {
        const int rcode = ensure_work_list(ctx, n);
        if (rcode != RC_OK) return rcode;
    }
    int rcode = scratch_acquire(ctx, s);
    if (rcode != RC_OK) return rcode;
    const int stats = ctx->stats_on || (extras && extras->hit_counts);
    // single-launch form (the descent kernel classifies): asked for by the caller (rc_check_host reads the
    // configurations in place), or for every culled check with RC_FEED=2
    rc::CheckExtras x;
    if (extras) x = *extras;
    if (mode == RC_MODE_CULLED && !stats && !ctx->t.gobs && n <= kFeedMax && (x.feed || ctx->knob_feed >= 2)) {
        if ((rcode = ensure_feed(ctx)) != RC_OK) { scratch_release(ctx, s); return rcode; }
        x.feed = 1;
        if (ctx->knob_feed_cls > 0) x.feed_cls_warps = ctx->knob_feed_cls;
        if (++ctx->feed_gen == 0) ctx->feed_gen = 1;
        x.feed_gen = ctx->feed_gen;
    } else {
        x.feed = 0;
    }
    cudaError_t e = launch_check(ctx, d_confs, n, d_valid, d_bitmask, mode, stats, s, nullptr, &x);
    scratch_release(ctx, s);
    if (e == cudaErrorInvalidConfiguration)
        return set_error(ctx, RC_ERR_UNSUPPORTED, "mesh tables do not fit in shared memory");
    if (e == cudaSuccess) ctx->launches += mode == RC_MODE_CULLED ? 2 : 1;
    return check_cuda(ctx, e, what);
}

int rc_check(rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid, uint32_t *d_bitmask,
             int mode, void *stream)
{
    const int rcode = check_args(ctx, n, d_confs, mode);
    if (rcode != RC_OK) return rcode;
    if (!d_valid && !d_bitmask) return RC_ERR_INVALID;
    if (n == 0) return RC_OK;
    cudaSetDevice(ctx->device);
    return check_on_stream(ctx, d_confs, n, d_valid, d_bitmask, mode, stream_of(ctx, stream), nullptr,
                           "rc_check launch");
}

long rc_gather_buffer_words(int world, long words_per_rank)
{
    if (world <= 0 || world > rc::kMaxGatherRanks || words_per_rank <= 0) return RC_ERR_INVALID;
    return 2L * world * words_per_rank + world + rc::kGatherHeaderWords;
}

// wait = 1: one-call form (the last CTA of the descent also waits for the peers); 0: push only
static int check_gather(rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid,
                        void *const *peer_buffers, int rank, int world, long words_per_rank,
                        uint32_t **d_result, void *stream, int wait)
{
    if (!ctx || n < 0 || (n > 0 && !d_confs) || !peer_buffers || world <= 0 ||
        world > rc::kMaxGatherRanks || rank < 0 || rank >= world || words_per_rank <= 0 ||
        (n + 31) / 32 > words_per_rank || words_per_rank > 0x7fffffffL / 4)
        return RC_ERR_INVALID;
    for (int r = 0; r < world; r++)
        if (!peer_buffers[r]) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStream_t s = stream_of(ctx, stream);
    uint32_t *mine = static_cast<uint32_t *>(peer_buffers[rank]);
    uint32_t *row = mine + 2L * world * words_per_rank;
    cudaError_t e;
    // The call numbering lives in the buffer (header word [0], written by every push): a context
    // that sees this buffer for the first time picks it up from there, so a fresh context on a used
    // buffer continues the sequence of its peers instead of restarting at 1 (which would make every
    // wait pass at once on stale signals).
    if (ctx->gather_buf != mine) {
        uint32_t calls = 0;
        if ((e = cudaStreamSynchronize(s)) ||
            (e = cudaMemcpy(&calls, row + world, sizeof calls, cudaMemcpyDeviceToHost)))
            return check_cuda(ctx, e, "rc_check_gather: reading the call number of the buffer");
        ctx->gather_buf = mine;
        ctx->gather_calls = calls;
    }
    const uint32_t call = ctx->gather_calls + 1;  // committed only when every launch below went through
    const long area = (long)(call & 1u) * world * words_per_rank;  // the two gather areas alternate
    uint32_t *slot = mine + area + (long)rank * words_per_rank;
    rc::GatherArgs a{};
    a.src = slot;
    for (int r = 0; r < world; r++) {
        uint32_t *base = static_cast<uint32_t *>(peer_buffers[r]);
        a.peer_slot[r] = base + area + (long)rank * words_per_rank;
        a.peer_signal[r] = base + 2L * world * words_per_rank + rank;
    }
    a.my_row = row;
    a.rank = rank; a.world = world; a.words = (int)words_per_rank; a.call = call; a.wait = wait;
    a.counters = ctx->d_counters;
    a.timeout_ns = 4000000000ULL;
    const long used = (n + 31) / 32;
    // words of the slot past the last configuration (short or empty last shards) are defined as 0
    if (used < words_per_rank &&
        (e = cudaMemsetAsync(slot + used, 0, (size_t)(words_per_rank - used) * 4, s)) != cudaSuccess)
        return check_cuda(ctx, e, "rc_check_gather memset");
    if (n > 0) {
        // the bitmask goes straight into the rank's slot; the last CTA of the descent kernel pushes it
        rc::CheckExtras x;
        x.gather = &a;
        const int rcode = check_on_stream(ctx, d_confs, n, d_valid, slot, RC_MODE_CULLED, s, &x,
                                          "rc_check_gather launch");
        if (rcode != RC_OK) return rcode;
    } else if (world > 1) {
        if ((e = rc::launch_gather_push(ctx, a, s)) != cudaSuccess) return check_cuda(ctx, e, "rc_check_gather push");
        ctx->launches += 1;
    }
    ctx->gather_calls = call;
    if (d_result) *d_result = mine + area;
    return RC_OK;
}

int rc_check_gather(rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid,
                    void *const *peer_buffers, int rank, int world, long words_per_rank,
                    uint32_t **d_result, void *stream)
{
    return check_gather(ctx, d_confs, n, d_valid, peer_buffers, rank, world, words_per_rank, d_result, stream, 1);
}

int rc_check_gather_begin(rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid,
                          void *const *peer_buffers, int rank, int world, long words_per_rank,
                          uint32_t **d_result, void *stream)
{
    return check_gather(ctx, d_confs, n, d_valid, peer_buffers, rank, world, words_per_rank, d_result, stream, 0);
}

int rc_gather_end(rc_ctx *ctx, void *const *peer_buffers, int rank, int world, long words_per_rank, void *stream)
{
    if (!ctx || !peer_buffers || world <= 0 || world > rc::kMaxGatherRanks || rank < 0 || rank >= world ||
        words_per_rank <= 0 || !peer_buffers[rank])
        return RC_ERR_INVALID;
    uint32_t *mine = static_cast<uint32_t *>(peer_buffers[rank]);
    if (ctx->gather_buf != mine || ctx->gather_calls == 0)
        return set_error(ctx, RC_ERR_INVALID, "rc_gather_end without rc_check_gather_begin on this buffer");
    cudaSetDevice(ctx->device);
    rc::GatherArgs a{};
    a.my_row = mine + 2L * world * words_per_rank;
    a.rank = rank; a.world = world; a.words = (int)words_per_rank; a.call = ctx->gather_calls; a.wait = 1;
    a.counters = ctx->d_counters;
    a.timeout_ns = 4000000000ULL;
    if (world > 1) ctx->launches += 1;
    return check_cuda(ctx, rc::launch_gather_wait(ctx, a, stream_of(ctx, stream)), "rc_gather_end");
}

int rc_gather_status(rc_ctx *ctx, const void *own_buffer, int world, long words_per_rank, uint32_t *calls,
                     uint32_t *timeouts)
{
    if (!ctx || !own_buffer || world <= 0 || world > rc::kMaxGatherRanks || words_per_rank <= 0) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    uint32_t h[2] = {0, 0};
    const uint32_t *row = static_cast<const uint32_t *>(own_buffer) + 2L * world * words_per_rank + world;
    const cudaError_t e = cudaMemcpy(h, row, sizeof h, cudaMemcpyDeviceToHost);  // synchronises with the device
    if (e != cudaSuccess) return check_cuda(ctx, e, "rc_gather_status");
    if (calls) *calls = h[0];
    if (timeouts) *timeouts = h[1];
    return RC_OK;
}

int rc_check_host(rc_ctx *ctx, const rc_config *h_confs, long n, uint8_t *h_valid, int mode)
{
    int rcode = check_args(ctx, n, h_confs, mode);
    if (rcode != RC_OK) return rcode;
    if (n > 0 && !h_valid) return RC_ERR_INVALID;
    if (n == 0) return RC_OK;
    cudaSetDevice(ctx->device);
    rcode = ensure_staging(ctx, n);
    if (rcode != RC_OK) return rcode;
    cudaStream_t s = ctx->own_stream;
    cudaError_t e;
    // page-locked caller buffers are used in place; pageable ones go through the pinned staging
    cudaPointerAttributes pa_in{}, pa_out{};
    const bool in_pinned = cudaPointerGetAttributes(&pa_in, h_confs) == cudaSuccess &&
                           pa_in.type == cudaMemoryTypeHost;
    const bool out_pinned = cudaPointerGetAttributes(&pa_out, h_valid) == cudaSuccess &&
                            pa_out.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const rc_config *src = h_confs;
    if (!in_pinned) {
        std::memcpy(ctx->h_pinned_confs, h_confs, (size_t)n * sizeof(rc_config));
        src = ctx->h_pinned_confs;
    }
    uint8_t *dst = out_pinned ? h_valid : ctx->h_pinned_valid;
    // RC_HOST_DIRECT=1: the flags go straight into the page-locked host buffer -- the kernels store
    // one byte per configuration through the mapped address (posted PCIe writes spread over the
    // run) and there is no device-to-host copy at the end.  Off by default: measured in one process
    // (profiles/tools/e2e_breakdown.py) the copy is 12 us cheaper per 100,000 configurations than
    // the 100,000 small writes.
    uint8_t *d_out = ctx->d_stage_valid;
    bool direct = false;
    if (ctx->knob_host_direct) {
        void *mapped = nullptr;
        if (cudaHostGetDevicePointer(&mapped, dst, 0) == cudaSuccess && mapped) {
            d_out = static_cast<uint8_t *>(mapped);
            direct = true;
        }
        cudaGetLastError();
    }
    // Waves (RC_HOST_CHUNKS = 2 .. 4; default 1 = off): the batch is cut into contiguous pieces; the
    // configurations of piece k+1 travel on the copy stream while the kernels of piece k run, and the
    // flags of piece k travel back while piece k+1 is being checked.  Measured on 100,000 hard
    // configurations (profiles/host_waves_r02.txt): 208 M configs/s in one piece, 184 M in two, 130 M
    // in four -- every piece pays the two launches, the stream hand-overs and the fixed start-up of the
    // descent again, which is more than the 25-50 us of copy it hides.  A caller that wants the copy
    // hidden keeps two whole batches in flight instead (rc_check_host_submit / _wait: 0.96 of the
    // device-resident rate).
    int waves = ctx->knob_host_chunks > 0 ? std::min(ctx->knob_host_chunks, kHostWaves) : 1;
    if (direct || n < 4096) waves = 1;
    // Single launch, no host-to-device copy (RC_FEED >= 1, the default): the kernel reads the
    // configurations where they lie -- the caller's page-locked buffer, or the pinned staging area a
    // pageable one was copied to -- over PCIe, four warps per CTA classifying batches as they come in while
    // the other warps already run descents on the ones that need them.  The copy (2.4 MB, ~50 us for
    // 100,000 configurations), the classify step and the second launch no longer stand in front of the
    // descents.
    const rc_config *mapped_confs = nullptr;
    // ... for batches large enough that the copy is worth hiding: below ~64 k configurations the copy path
    // is quicker (2,000 configurations 88 us against 97, 30,000: 200-223 against 254-264; 100,000: 360-470
    // against 344-352, profiles/feed_r02.txt), RC_FEED=2 uses the form for any size
    if (ctx->knob_feed >= 1 && (n >= kFeedMinHost || ctx->knob_feed >= 2) && mode == RC_MODE_CULLED && !ctx->stats_on &&
        !ctx->t.gobs && n <= kFeedMax && waves <= 1) {
        void *mp = nullptr;
        if (cudaHostGetDevicePointer(&mp, const_cast<rc_config *>(src), 0) == cudaSuccess && mp)
            mapped_confs = static_cast<const rc_config *>(mp);
        cudaGetLastError();
    }
    if (mapped_confs) {
        rc::CheckExtras x;
        x.feed = 1;
        x.feed_cls_warps = 4;
        cudaEventRecord(ctx->ev0, s);
        if ((rcode = check_on_stream(ctx, mapped_confs, n, d_out, nullptr, mode, s, &x, "rc_check_host launch (single)")) != RC_OK)
            return rcode;
        cudaEventRecord(ctx->ev1, s);
        if ((!direct && (e = cudaMemcpyAsync(dst, ctx->d_stage_valid, (size_t)n, cudaMemcpyDeviceToHost, s))) ||
            (e = cudaStreamSynchronize(s)))
            return check_cuda(ctx, e, "D2H flags");
    } else if (waves <= 1) {
        if ((e = cudaMemcpyAsync(ctx->d_stage_confs, src, (size_t)n * sizeof(rc_config),
                                 cudaMemcpyHostToDevice, s)))
            return check_cuda(ctx, e, "H2D configurations");
        cudaEventRecord(ctx->ev0, s);
        if ((rcode = check_on_stream(ctx, ctx->d_stage_confs, n, d_out, nullptr, mode, s, nullptr,
                                     "rc_check_host launch")) != RC_OK)
            return rcode;
        cudaEventRecord(ctx->ev1, s);
        if ((!direct && (e = cudaMemcpyAsync(dst, ctx->d_stage_valid, (size_t)n, cudaMemcpyDeviceToHost, s))) ||
            (e = cudaStreamSynchronize(s)))
            return check_cuda(ctx, e, "D2H flags");
    } else {
        cudaStream_t copy = ctx->own_stream2;
        for (int w = 0; w < 2 * waves; w++)
            if (!ctx->ev_wave[w] && (e = cudaEventCreateWithFlags(&ctx->ev_wave[w], cudaEventDisableTiming)))
                return check_cuda(ctx, e, "wave events");
        const long per = ((n + waves - 1) / waves + 31) / 32 * 32;
        // the copy stream first waits for whatever the compute stream still reads from the staging area
        if ((e = cudaEventRecord(ctx->ev_join, s)) || (e = cudaStreamWaitEvent(copy, ctx->ev_join, 0)))
            return check_cuda(ctx, e, "wave join");
        for (int w = 0; w < waves; w++) {
            const long lo = std::min(n, w * per), cnt = std::min(n, lo + per) - lo;
            if (cnt <= 0) continue;
            if ((e = cudaMemcpyAsync(ctx->d_stage_confs + lo, src + lo, (size_t)cnt * sizeof(rc_config),
                                     cudaMemcpyHostToDevice, copy)) ||
                (e = cudaEventRecord(ctx->ev_wave[w], copy)))
                return check_cuda(ctx, e, "H2D configurations (wave)");
        }
        for (int w = 0; w < waves; w++) {
            const long lo = std::min(n, w * per), cnt = std::min(n, lo + per) - lo;
            if (cnt <= 0) continue;
            if ((e = cudaStreamWaitEvent(s, ctx->ev_wave[w], 0))) return check_cuda(ctx, e, "wave wait");
            if (w == 0) cudaEventRecord(ctx->ev0, s);
            if ((rcode = check_on_stream(ctx, ctx->d_stage_confs + lo, cnt, ctx->d_stage_valid + lo, nullptr, mode, s,
                                         nullptr, "rc_check_host launch (wave)")) != RC_OK)
                return rcode;
            if ((e = cudaEventRecord(ctx->ev_wave[waves + w], s)) ||
                (e = cudaStreamWaitEvent(copy, ctx->ev_wave[waves + w], 0)) ||
                (e = cudaMemcpyAsync(dst + lo, ctx->d_stage_valid + lo, (size_t)cnt, cudaMemcpyDeviceToHost, copy)))
                return check_cuda(ctx, e, "D2H flags (wave)");
        }
        cudaEventRecord(ctx->ev1, s);
        if ((e = cudaStreamSynchronize(copy)) || (e = cudaStreamSynchronize(s)))
            return check_cuda(ctx, e, "wave synchronise");
    }
    cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev0, ctx->ev1);
    if (!out_pinned) std::memcpy(h_valid, ctx->h_pinned_valid, (size_t)n);
    return RC_OK;
}

int rc_check_host_submit(rc_ctx *ctx, const rc_config *h_confs, long n, uint8_t *h_valid, int mode, int *ticket)
{
    int rcode = check_args(ctx, n, h_confs, mode);
    if (rcode != RC_OK) return rcode;
    if (!ticket || n <= 0 || !h_valid) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    PipeLane &l = ctx->pipe[ctx->pipe_next & 1];
    rcode = finish_lane(ctx, l);  // the batch submitted two calls ago, if nobody has waited for it yet
    if (rcode != RC_OK || (rcode = ensure_lane(ctx, l, n)) != RC_OK) return rcode;
    cudaStream_t copy = ctx->own_stream2, compute = ctx->own_stream;
    cudaPointerAttributes pa_in{}, pa_out{};
    const bool in_pinned = cudaPointerGetAttributes(&pa_in, h_confs) == cudaSuccess && pa_in.type == cudaMemoryTypeHost;
    const bool out_pinned = cudaPointerGetAttributes(&pa_out, h_valid) == cudaSuccess && pa_out.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const rc_config *src = h_confs;
    if (!in_pinned) {
        std::memcpy(l.h_confs, h_confs, (size_t)n * sizeof(rc_config));
        src = l.h_confs;
    }
    uint8_t *dst = out_pinned ? h_valid : l.h_valid;
    void *mapped = nullptr;
    const bool direct = ctx->knob_host_direct && cudaHostGetDevicePointer(&mapped, dst, 0) == cudaSuccess && mapped;
    cudaGetLastError();
    cudaError_t e;
    // copy stream: the configurations; compute stream: the kernels of all batches, one after the other
    // (one work list per context), each waiting only for its own copy
    if ((e = cudaMemcpyAsync(l.d_confs, src, (size_t)n * sizeof(rc_config), cudaMemcpyHostToDevice, copy)) ||
        (e = cudaEventRecord(l.copied, copy)) || (e = cudaStreamWaitEvent(compute, l.copied, 0)))
        return check_cuda(ctx, e, "rc_check_host_submit copy");
    if ((rcode = check_on_stream(ctx, l.d_confs, n, direct ? static_cast<uint8_t *>(mapped) : l.d_valid, nullptr,
                                 mode, compute, nullptr, "rc_check_host_submit launch")) != RC_OK)
        return rcode;
    if ((!direct && (e = cudaMemcpyAsync(dst, l.d_valid, (size_t)n, cudaMemcpyDeviceToHost, compute))) ||
        (e = cudaEventRecord(l.done, compute)))
        return check_cuda(ctx, e, "rc_check_host_submit");
    l.busy = true; l.n = n; l.user_valid = out_pinned ? nullptr : h_valid;
    l.ticket = ctx->pipe_next;
    *ticket = ctx->pipe_next++;
    return RC_OK;
}

int rc_check_host_wait(rc_ctx *ctx, int ticket)
{
    if (!ctx || ticket < 0 || ticket >= ctx->pipe_next) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    PipeLane &l = ctx->pipe[ticket & 1];
    if (l.ticket != ticket) return RC_OK;  // completed when its lane was taken over by a later batch
    return finish_lane(ctx, l);
}

int rc_device_count(int *count)
{
    if (!count) return RC_ERR_INVALID;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        *count = 0;
        return RC_ERR_NO_DEVICE;
    }
    *count = n;
    return RC_OK;
}

int rc_shard_size(long n, int n_gpus)
{
    if (n < 0 || n_gpus <= 0) return RC_ERR_INVALID;
    const long per = ((n + n_gpus - 1) / n_gpus + 31) / 32 * 32;
    return per > 0x7fffffffL ? RC_ERR_INVALID : (int)per;
}

int rc_check_multi(rc_ctx **ctxs, int n_gpus, const rc_config *h_confs, long n, uint32_t *h_bitmask,
                   uint8_t *h_valid, uint32_t **d_bitmask_all)
{
    if (!ctxs || n_gpus <= 0 || n < 0 || (n > 0 && (!h_confs || (!h_bitmask && !h_valid && !d_bitmask_all))))
        return RC_ERR_INVALID;
    for (int g = 0; g < n_gpus; g++) {
        if (!ctxs[g]) return RC_ERR_INVALID;
        if (d_bitmask_all && !d_bitmask_all[g]) return RC_ERR_INVALID;
    }
    if (n == 0) return RC_OK;
    if (n > 0xffffffffL) return RC_ERR_INVALID;
    const long per = rc_shard_size(n, n_gpus);
    if (per <= 0) return RC_ERR_INVALID;
    cudaPointerAttributes pa{};
    const bool pinned = cudaPointerGetAttributes(&pa, h_confs) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (d_bitmask_all)  // direct peer copies where the topology allows them (once per pair)
        for (int g = 0; g < n_gpus; g++)
            for (int p = 0; p < n_gpus; p++) {
                const int dg = ctxs[g]->device, dp = ctxs[p]->device;
                if (dg == dp || dp >= 64 || (ctxs[g]->peers_enabled >> dp & 1ULL)) continue;
                int can = 0;
                cudaSetDevice(dg);
                if (cudaDeviceCanAccessPeer(&can, dg, dp) == cudaSuccess && can) cudaDeviceEnablePeerAccess(dp, 0);
                cudaGetLastError();  // "already enabled" is fine
                ctxs[g]->peers_enabled |= 1ULL << dp;
            }
    // 1. every device: stage its shard, launch the check (asynchronous: the devices overlap)
    for (int g = 0; g < n_gpus; g++) {
        rc_ctx *ctx = ctxs[g];
        const long lo = std::min(n, (long)g * per), cnt = std::min(n, lo + per) - lo;
        if (cnt <= 0) continue;
        cudaSetDevice(ctx->device);
        int rcode = ensure_staging(ctx, cnt);
        if (rcode != RC_OK) return rcode;
        cudaStream_t s = ctx->own_stream;
        const rc_config *src = h_confs + lo;
        if (!pinned) {
            std::memcpy(ctx->h_pinned_confs, src, (size_t)cnt * sizeof(rc_config));
            src = ctx->h_pinned_confs;
        }
        cudaError_t e = cudaMemcpyAsync(ctx->d_stage_confs, src, (size_t)cnt * sizeof(rc_config),
                                        cudaMemcpyHostToDevice, s);
        if (e) return check_cuda(ctx, e, "rc_check_multi H2D");
        // with a gather buffer on this device the kernel writes its words straight into its slot
        uint32_t *mask = d_bitmask_all ? d_bitmask_all[g] + lo / 32 : ctx->d_stage_mask;
        if ((rcode = check_on_stream(ctx, ctx->d_stage_confs, cnt, h_valid ? ctx->d_stage_valid : nullptr, mask,
                                     RC_MODE_CULLED, s, nullptr, "rc_check_multi launch")) != RC_OK)
            return rcode;
        const size_t words = (size_t)(cnt + 31) / 32;
        if (h_bitmask && (e = cudaMemcpyAsync(ctx->h_pinned_mask, mask, words * 4, cudaMemcpyDeviceToHost, s)))
            return check_cuda(ctx, e, "rc_check_multi D2H bitmask");
        if (h_valid && (e = cudaMemcpyAsync(ctx->h_pinned_valid, ctx->d_stage_valid, (size_t)cnt,
                                            cudaMemcpyDeviceToHost, s)))
            return check_cuda(ctx, e, "rc_check_multi D2H flags");
        // 2. all-gather without a communicator: this shard's words go to every peer's buffer
        //    (cudaMemcpyPeerAsync: NVLink / NVSwitch when peer access is available, staged otherwise)
        if (d_bitmask_all)
            for (int p = 0; p < n_gpus; p++)
                if (p != g && (e = cudaMemcpyPeerAsync(d_bitmask_all[p] + lo / 32, ctxs[p]->device, mask,
                                                       ctx->device, words * 4, s)))
                    return check_cuda(ctx, e, "rc_check_multi peer copy");
    }
    // 3. join
    int rcode = RC_OK;
    for (int g = 0; g < n_gpus; g++) {
        rc_ctx *ctx = ctxs[g];
        const long lo = std::min(n, (long)g * per), cnt = std::min(n, lo + per) - lo;
        if (cnt <= 0) continue;
        cudaSetDevice(ctx->device);
        const cudaError_t e = cudaStreamSynchronize(ctx->own_stream);
        if (e) { rcode = check_cuda(ctx, e, "rc_check_multi join"); continue; }
        if (h_bitmask) std::memcpy(h_bitmask + lo / 32, ctx->h_pinned_mask, (size_t)((cnt + 31) / 32) * 4);
        if (h_valid) std::memcpy(h_valid + lo, ctx->h_pinned_valid, (size_t)cnt);
    }
    return rcode;
}

float rc_last_kernel_ms(const rc_ctx *ctx) { return ctx ? ctx->last_kernel_ms : 0.f; }

// ---- stage API -----------------------------------------------------------------------------

int rc_transform_aos(rc_ctx *ctx, const rc_config *d_confs, long n, const rc_vec3 *d_base, int n_v,
                     rc_vec3 *d_out, void *stream)
{
    if (!ctx || n < 0 || (n > 0 && (!d_confs || !d_out))) return RC_ERR_INVALID;
    if (!d_base) { d_base = ctx->t.rob_aos; n_v = ctx->t.n_rob_v; }
    if (n_v <= 0) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ctx->launches++;
    return check_cuda(ctx, launch_transform_aos(ctx, d_confs, n, d_base, n_v, d_out, stream_of(ctx, stream)),
                      "rc_transform_aos");
}

int rc_generate_aabb(rc_ctx *ctx, const rc_vec3 *d_verts, unsigned n_v, unsigned n, rc_aabb *d_out,
                     void *stream)
{
    if (!ctx || (n > 0 && (!d_verts || !d_out)) || n_v == 0) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ctx->launches++;
    return check_cuda(ctx, launch_aabb_aos(d_verts, n_v, n, d_out, stream_of(ctx, stream)),
                      "rc_generate_aabb");
}

int rc_generate_aabb_host(rc_ctx *ctx, const rc_vec3 *h_verts, unsigned n_v, unsigned n, rc_aabb *h_out)
{
    if (!ctx || (n > 0 && (!h_verts || !h_out)) || n_v == 0) return RC_ERR_INVALID;
    if (n == 0) return RC_OK;
    cudaSetDevice(ctx->device);
    rc_vec3 *dv = nullptr;
    rc_aabb *db = nullptr;
    const size_t bytes = (size_t)n * n_v * sizeof(rc_vec3);
    cudaError_t e;
    int rcode = RC_OK;
    if ((e = cudaMalloc(&dv, bytes)) || (e = cudaMalloc(&db, (size_t)n * sizeof(rc_aabb))) ||
        (e = cudaMemcpyAsync(dv, h_verts, bytes, cudaMemcpyHostToDevice, ctx->own_stream)) ||
        (e = launch_aabb_aos(dv, n_v, n, db, ctx->own_stream)) ||
        (e = cudaMemcpyAsync(h_out, db, (size_t)n * sizeof(rc_aabb), cudaMemcpyDeviceToHost,
                             ctx->own_stream)) ||
        (e = cudaStreamSynchronize(ctx->own_stream)))
        rcode = check_cuda(ctx, e, "rc_generate_aabb_host");
    ctx->launches++;
    cudaFree(dv);
    cudaFree(db);
    return rcode;
}

int rc_generate_aabb_soa(rc_ctx *ctx, const float *d_x, const float *d_y, const float *d_z,
                         unsigned n_v, unsigned n, rc_aabb *d_out, void *stream)
{
    if (!ctx || (n > 0 && (!d_x || !d_y || !d_z || !d_out)) || n_v == 0) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ctx->launches++;
    return check_cuda(ctx, launch_aabb_soa(d_x, d_y, d_z, n_v, n, d_out, stream_of(ctx, stream)),
                      "rc_generate_aabb_soa");
}

int rc_broad_phase(rc_ctx *ctx, long n, const rc_aabb *d_robots, const rc_aabb *d_obstacle,
                   uint8_t *d_valid, void *stream)
{
    if (!ctx || n < 0 || (n > 0 && (!d_robots || !d_valid))) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ctx->launches++;
    return check_cuda(ctx, launch_broad(n, d_robots, d_obstacle ? d_obstacle : ctx->d_obs_box, d_valid,
                                        stream_of(ctx, stream)),
                      "rc_broad_phase");
}

int rc_broad_fused(rc_ctx *ctx, const rc_config *d_confs, long n, float *d_x, float *d_y, float *d_z,
                   rc_aabb *d_aabb, uint8_t *d_valid, void *stream)
{
    if (!ctx || n < 0 || (n > 0 && !d_confs)) return RC_ERR_INVALID;
    const int given = (d_x != nullptr) + (d_y != nullptr) + (d_z != nullptr);
    if (given != 0 && given != 3) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ctx->launches++;
    return check_cuda(ctx, launch_transform_soa(ctx, d_confs, n, d_x, d_y, d_z, d_aabb, d_valid,
                                                stream_of(ctx, stream)),
                      "rc_broad_fused");
}

int rc_narrow_phase(rc_ctx *ctx, const float *d_x, const float *d_y, const float *d_z, long n,
                    uint8_t *d_valid, void *stream)
{
    if (!ctx || n < 0 || (n > 0 && (!d_x || !d_y || !d_z || !d_valid))) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaError_t e = launch_narrow_soa(ctx, d_x, d_y, d_z, n, d_valid, stream_of(ctx, stream));
    if (e == cudaErrorInvalidConfiguration)
        return set_error(ctx, RC_ERR_UNSUPPORTED, "mesh tables do not fit in shared memory");
    ctx->launches++;
    return check_cuda(ctx, e, "rc_narrow_phase");
}

int rc_narrow_phase_host(int n, int n_rob_t, int n_rob_v, int n_obs_t, int n_obs_v,
                         const rc_tri *rob_trs, const rc_vec3 *rob_pts, const rc_tri *obs_trs,
                         const rc_vec3 *obs_pts, uint8_t *valid, int device)
{
    if (n < 0 || n_rob_t <= 0 || n_rob_v <= 0 || n_obs_t <= 0 || n_obs_v <= 0 || !rob_trs ||
        !rob_pts || !obs_trs || !obs_pts || !valid)
        return RC_ERR_INVALID;
    if (n == 0) return RC_OK;
    // The robot vertices arrive already transformed, [conf][vertex]; the context only needs the
    // topology and the obstacle, so it is built from configuration 0's vertices.
    std::vector<float> rx(n_rob_v), ry(n_rob_v), rz(n_rob_v), ox(n_obs_v), oy(n_obs_v), oz(n_obs_v);
    std::vector<int> r1(n_rob_t), r2(n_rob_t), r3(n_rob_t), o1(n_obs_t), o2(n_obs_t), o3(n_obs_t);
    for (int v = 0; v < n_rob_v; v++) { rx[v] = rob_pts[v].x; ry[v] = rob_pts[v].y; rz[v] = rob_pts[v].z; }
    for (int v = 0; v < n_obs_v; v++) { ox[v] = obs_pts[v].x; oy[v] = obs_pts[v].y; oz[v] = obs_pts[v].z; }
    for (int t = 0; t < n_rob_t; t++) { r1[t] = rob_trs[t].v1; r2[t] = rob_trs[t].v2; r3[t] = rob_trs[t].v3; }
    for (int t = 0; t < n_obs_t; t++) { o1[t] = obs_trs[t].v1; o2[t] = obs_trs[t].v2; o3[t] = obs_trs[t].v3; }
    rc_ctx *ctx = nullptr;
    int rcode = rc_create(&ctx, rx.data(), ry.data(), rz.data(), n_rob_v, r1.data(), r2.data(),
                          r3.data(), n_rob_t, ox.data(), oy.data(), oz.data(), n_obs_v, o1.data(),
                          o2.data(), o3.data(), n_obs_t, device);
    if (rcode != RC_OK) return rcode;
    const size_t cnt = (size_t)n * n_rob_v;
    std::vector<float> hx(cnt), hy(cnt), hz(cnt);
    for (size_t e = 0; e < cnt; e++) { hx[e] = rob_pts[e].x; hy[e] = rob_pts[e].y; hz[e] = rob_pts[e].z; }
    float *dx = nullptr, *dy = nullptr, *dz = nullptr;
    uint8_t *dv = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&dx, cnt * 4)) || (e = cudaMalloc(&dy, cnt * 4)) || (e = cudaMalloc(&dz, cnt * 4)) ||
        (e = cudaMalloc(&dv, (size_t)n)) ||
        (e = cudaMemcpy(dx, hx.data(), cnt * 4, cudaMemcpyHostToDevice)) ||
        (e = cudaMemcpy(dy, hy.data(), cnt * 4, cudaMemcpyHostToDevice)) ||
        (e = cudaMemcpy(dz, hz.data(), cnt * 4, cudaMemcpyHostToDevice)) ||
        (e = cudaMemcpy(dv, valid, (size_t)n, cudaMemcpyHostToDevice)))
        rcode = check_cuda(ctx, e, "rc_narrow_phase_host copies");
    if (rcode == RC_OK) rcode = rc_narrow_phase(ctx, dx, dy, dz, n, dv, nullptr);
    if (rcode == RC_OK &&
        ((e = cudaStreamSynchronize(ctx->own_stream)) ||
         (e = cudaMemcpy(valid, dv, (size_t)n, cudaMemcpyDeviceToHost))))
        rcode = check_cuda(ctx, e, "rc_narrow_phase_host result");
    cudaFree(dx); cudaFree(dy); cudaFree(dz); cudaFree(dv);
    rc_destroy(ctx);
    return rcode;
}

int rc_megakernel(rc_ctx *ctx, const rc_config *d_configs, const rc_aabb *, const rc_aabb *,
                  uint8_t *d_valid_confs, int num_configs, void *stream)
{
    // writes every flag; on a buffer pre-filled with true that equals "only ever writes false"
    return rc_check(ctx, d_configs, num_configs, d_valid_confs, nullptr, RC_MODE_CULLED, stream);
}

int rc_obstacle_tri_aabbs(const rc_ctx *ctx, rc_aabb *h_out)
{
    if (!ctx || !h_out) return RC_ERR_INVALID;
    for (int t = 0; t < ctx->t.n_obs_t; t++) {
        const int idx[3] = {ctx->obs_v1[t], ctx->obs_v2[t], ctx->obs_v3[t]};
        rc_aabb b = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int k = 0; k < 3; k++) {
            b.x_min = std::min(b.x_min, ctx->obs_x[idx[k]]); b.x_max = std::max(b.x_max, ctx->obs_x[idx[k]]);
            b.y_min = std::min(b.y_min, ctx->obs_y[idx[k]]); b.y_max = std::max(b.y_max, ctx->obs_y[idx[k]]);
            b.z_min = std::min(b.z_min, ctx->obs_z[idx[k]]); b.z_max = std::max(b.z_max, ctx->obs_z[idx[k]]);
        }
        h_out[t] = b;
    }
    return RC_OK;
}

int rc_conf_generate_device(rc_ctx *ctx, rc_config *d_out, long n, float x_min, float x_max,
                            float y_min, float y_max, float z_min, float z_max, uint64_t seed,
                            void *stream)
{
    if (!ctx || n < 0 || (n > 0 && !d_out)) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    const float box[6] = {x_min, x_max, y_min, y_max, z_min, z_max};
    ctx->launches++;
    return check_cuda(ctx, launch_conf_generate(d_out, n, box, seed, stream_of(ctx, stream)),
                      "rc_conf_generate_device");
}

int rc_conf_generate_device_mix(rc_ctx *ctx, rc_config *d_out, long first, long n, const float *box_a,
                                const float *box_b, int block, int period, uint64_t seed, void *stream)
{
    if (!ctx || n < 0 || first < 0 || (n > 0 && !d_out) || !box_a || period < 0 || (period > 0 && (!box_b || block <= 0)))
        return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ctx->launches++;
    return check_cuda(ctx, launch_conf_generate_mix(d_out, first, n, box_a, box_b ? box_b : box_a, block, period,
                                                    seed, stream_of(ctx, stream)),
                      "rc_conf_generate_device_mix");
}

int rc_compact(rc_ctx *ctx, const uint8_t *d_valid, const uint32_t *d_bitmask, long n, int want,
               uint32_t *d_indices, unsigned long long *d_count, void *stream)
{
    if (!ctx || n < 0 || n > 0xffffffffL || !d_count || (n > 0 && (!d_indices || (!d_valid && !d_bitmask))))
        return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (n == 0)
        return check_cuda(ctx, cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream_of(ctx, stream)),
                          "rc_compact");
    const long n_blocks = (n + 4095) / 4096;
    if (n_blocks > ctx->block_counts_cap) {
        // grown on demand; the old buffer may still be in use by an earlier asynchronous call
        cudaDeviceSynchronize();
        if (ctx->d_block_counts) cudaFree(ctx->d_block_counts);
        ctx->d_block_counts = nullptr;
        ctx->block_counts_cap = 0;
        cudaError_t e = cudaMalloc(&ctx->d_block_counts, (size_t)std::max(n_blocks, 1024L) * 4);
        if (e) return check_cuda(ctx, e, "rc_compact scratch");
        ctx->block_counts_cap = std::max(n_blocks, 1024L);
    }
    ctx->launches += 3;
    return check_cuda(ctx, launch_compact(d_valid, d_bitmask, n, want ? 1 : 0, d_indices, d_count,
                                          ctx->d_block_counts, stream_of(ctx, stream)),
                      "rc_compact");
}

int rc_set_robot_spheres(rc_ctx *ctx, const float *xyzr, int n)
{
    if (!ctx || (xyzr && n <= 0)) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DeviceTables &t = ctx->t;
    std::vector<float> s;
    if (xyzr) {
        s.assign(xyzr, xyzr + 4 * (size_t)n);
        for (int k = 0; k < n; k++)
            if (!(s[4 * k + 3] >= 0.f)) return RC_ERR_INVALID;
    } else {
        n = t.n_rob_t;
        s.resize(4 * (size_t)n);
        triangle_spheres(ctx->rob_x.data(), ctx->rob_y.data(), ctx->rob_z.data(), ctx->rob_v1.data(),
                         ctx->rob_v2.data(), ctx->rob_v3.data(), n, s.data());
    }
    // tree over the spheres' boxes (robot frame), spheres stored in leaf order
    std::vector<float> lo(3 * (size_t)n), hi(lo.size()), cen(lo.size());
    double blo[3] = {1e300, 1e300, 1e300}, bhi[3] = {-1e300, -1e300, -1e300};
    for (int k = 0; k < n; k++)
        for (int q = 0; q < 3; q++) {
            // box of the sphere, widened by an ulp-scale slack so that it really contains it
            const float r = s[4 * k + 3] * (1.0f + 1e-6f) + 1e-6f;
            lo[3 * (size_t)k + q] = s[4 * k + q] - r;
            hi[3 * (size_t)k + q] = s[4 * k + q] + r;
            cen[3 * (size_t)k + q] = s[4 * k + q];
            blo[q] = std::min(blo[q], (double)lo[3 * (size_t)k + q]);
            bhi[q] = std::max(bhi[q], (double)hi[3 * (size_t)k + q]);
        }
    const Bvh tree = build_bvh_boxes(lo.data(), hi.data(), cen.data(), n, kBoxMargin);
    std::vector<float4> sorted((size_t)n), nodes((size_t)tree.n_nodes * 2);
    for (int p = 0; p < n; p++) {
        const int k = tree.order[p];
        sorted[p] = make_float4(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]);
    }
    std::memcpy(nodes.data(), tree.nodes.data(), tree.nodes.size() * sizeof(float));
    double r2 = 0;
    for (int q = 0; q < 3; q++) t.sph_bound[q] = (float)(0.5 * (blo[q] + bhi[q]));
    for (int k = 0; k < n; k++) {
        double d2 = 0;
        for (int q = 0; q < 3; q++) d2 += ((double)s[4 * k + q] - t.sph_bound[q]) * ((double)s[4 * k + q] - t.sph_bound[q]);
        r2 = std::max(r2, std::sqrt(d2) + (double)s[4 * k + 3]);
    }
    t.sph_bound[3] = (float)(r2 * (1.0 + 1e-6)) + 1e-4f;
    for (int q = 0; q < 3; q++) {
        t.sph_box_c[q] = t.sph_bound[q];
        const double e = std::max(bhi[q] - (double)t.sph_box_c[q], (double)t.sph_box_c[q] - blo[q]);
        t.sph_box_e[q] = (float)(e * (1.0 + 1e-6)) + 1e-4f;
    }
    cudaError_t e;
    cudaDeviceSynchronize();  // an earlier asynchronous check may still read the old tables
    for (void *old : {static_cast<void *>(t.sph), static_cast<void *>(t.sph_bvh)}) {  // replaced, not leaked
        if (!old) continue;
        auto it = std::find(ctx->allocations.begin(), ctx->allocations.end(), old);
        if (it != ctx->allocations.end()) ctx->allocations.erase(it);
        cudaFree(old);
    }
    t.sph = nullptr; t.sph_bvh = nullptr; t.n_sph = 0; t.n_sph_nodes = 0;
    if ((e = upload(ctx, &t.sph, sorted)) || (e = upload(ctx, &t.sph_bvh, nodes)))
        return check_cuda(ctx, e, "sphere tables");
    t.n_sph = n;
    t.n_sph_nodes = tree.n_nodes;
    ctx->spheres = s;
    return RC_OK;
}

int rc_robot_spheres(const rc_ctx *ctx, float *xyzr_out, int *n)
{
    if (!ctx || !n) return RC_ERR_INVALID;
    *n = (int)(ctx->spheres.size() / 4);
    if (xyzr_out && !ctx->spheres.empty())
        std::memcpy(xyzr_out, ctx->spheres.data(), ctx->spheres.size() * sizeof(float));
    return RC_OK;
}

int rc_check_spheres(rc_ctx *ctx, const rc_config *d_confs, long n, uint8_t *d_valid,
                     uint32_t *d_bitmask, void *stream)
{
    if (!ctx || n < 0 || (n > 0 && !d_confs) || (!d_valid && !d_bitmask)) return RC_ERR_INVALID;
    if (n == 0) return RC_OK;
    if (ctx->t.n_sph == 0) {
        const int rcode = rc_set_robot_spheres(ctx, nullptr, 0);
        if (rcode != RC_OK) return rcode;
    }
    cudaSetDevice(ctx->device);
    {
        const int rcode = ensure_work_list(ctx, n);
        if (rcode != RC_OK) return rcode;
    }
    {
        const int rcode = scratch_acquire(ctx, stream_of(ctx, stream));
        if (rcode != RC_OK) return rcode;
    }
    cudaError_t e = launch_check_spheres(ctx, d_confs, n, d_valid, d_bitmask, ctx->stats_on,
                                         stream_of(ctx, stream), nullptr);
    scratch_release(ctx, stream_of(ctx, stream));
    if (e == cudaErrorInvalidConfiguration)
        return set_error(ctx, RC_ERR_UNSUPPORTED, "sphere / obstacle tables do not fit");
    ctx->launches += 2;  // classify + sphere descent
    return check_cuda(ctx, e, "rc_check_spheres launch");
}

// ---- audits ------------------------------------------------------------------------------------

int rc_hit_counts(rc_ctx *ctx, const rc_config *d_confs, long n, int mode, uint8_t *d_valid,
                  uint32_t *d_counts, void *stream)
{
    const int rcode = check_args(ctx, n, d_confs, mode);
    if (rcode != RC_OK) return rcode;
    if (n > 0 && !d_counts) return RC_ERR_INVALID;
    if (n == 0) return RC_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t s = stream_of(ctx, stream);
    uint8_t *flags = d_valid;
    if (!flags) {  // the kernels need one of the two outputs: a scratch flag array of the context
        const int r2 = ensure_staging(ctx, n);
        if (r2 != RC_OK) return r2;
        flags = ctx->d_stage_valid;
    }
    // configurations the cheap bounds of step 1 clear never reach the counting kernel: their count is 0
    const cudaError_t e = cudaMemsetAsync(d_counts, 0, (size_t)n * sizeof(uint32_t), s);
    if (e != cudaSuccess) return check_cuda(ctx, e, "rc_hit_counts memset");
    rc::CheckExtras x;
    x.hit_counts = d_counts;
    return check_on_stream(ctx, d_confs, n, flags, nullptr, mode, s, &x, "rc_hit_counts launch");
}

int rc_audit(rc_ctx *ctx, const rc_config *d_confs, long n, long *n_flags_differ, long *n_counts_differ,
             uint32_t *h_indices, long cap)
{
    if (!ctx || n < 0 || (n > 0 && !d_confs) || cap < 0 || (cap > 0 && !h_indices)) return RC_ERR_INVALID;
    if (n_flags_differ) *n_flags_differ = 0;
    if (n_counts_differ) *n_counts_differ = 0;
    if (n == 0) return RC_OK;
    if (n > 0xffffffffL) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    uint8_t *d_flags = nullptr;
    uint32_t *d_cnt = nullptr;
    std::vector<uint8_t> f[2];
    std::vector<uint32_t> k[2];
    cudaError_t e = cudaSuccess;
    int rcode = RC_OK;
    if ((e = cudaMalloc(&d_flags, (size_t)n)) || (e = cudaMalloc(&d_cnt, (size_t)n * 4)))
        rcode = check_cuda(ctx, e, "rc_audit buffers");
    for (int mode = 0; mode < 2 && rcode == RC_OK; mode++) {
        f[mode].resize((size_t)n);
        k[mode].resize((size_t)n);
        rcode = rc_hit_counts(ctx, d_confs, n, mode == 0 ? RC_MODE_CULLED : RC_MODE_BRUTE, d_flags, d_cnt,
                              ctx->own_stream);
        if (rcode == RC_OK &&
            ((e = cudaMemcpyAsync(f[mode].data(), d_flags, (size_t)n, cudaMemcpyDeviceToHost, ctx->own_stream)) ||
             (e = cudaMemcpyAsync(k[mode].data(), d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->own_stream)) ||
             (e = cudaStreamSynchronize(ctx->own_stream))))
            rcode = check_cuda(ctx, e, "rc_audit copies");
    }
    cudaFree(d_flags);
    cudaFree(d_cnt);
    if (rcode != RC_OK) return rcode;
    long nf = 0, nc = 0, listed = 0;
    for (long i = 0; i < n; i++) {
        const bool df = f[0][i] != f[1][i], dc = k[0][i] != k[1][i];
        nf += df;
        nc += dc;
        if ((df || dc) && listed < cap) h_indices[listed++] = (uint32_t)i;
    }
    if (n_flags_differ) *n_flags_differ = nf;
    if (n_counts_differ) *n_counts_differ = nc;
    return RC_OK;
}

int rc_kat_pairs(rc_ctx *ctx, const float *h_rob9, const float *h_obs9, int n, float *h_out32, int *h_verdict)
{
    if (!ctx || n < 0 || (n > 0 && (!h_rob9 || !h_obs9 || !h_out32 || !h_verdict))) return RC_ERR_INVALID;
    if (n == 0) return RC_OK;
    cudaSetDevice(ctx->device);
    float *d_r = nullptr, *d_o = nullptr, *d_out = nullptr;
    int *d_v = nullptr;
    cudaError_t e;
    int rcode = RC_OK;
    if ((e = cudaMalloc(&d_r, (size_t)n * 36)) || (e = cudaMalloc(&d_o, (size_t)n * 36)) ||
        (e = cudaMalloc(&d_out, (size_t)n * 128)) || (e = cudaMalloc(&d_v, (size_t)n * 4)) ||
        (e = cudaMemcpy(d_r, h_rob9, (size_t)n * 36, cudaMemcpyHostToDevice)) ||
        (e = cudaMemcpy(d_o, h_obs9, (size_t)n * 36, cudaMemcpyHostToDevice)) ||
        (e = rc::launch_kat_pairs(d_r, d_o, n, d_out, d_v, ctx->own_stream)) ||
        (e = cudaStreamSynchronize(ctx->own_stream)) ||
        (e = cudaMemcpy(h_out32, d_out, (size_t)n * 128, cudaMemcpyDeviceToHost)) ||
        (e = cudaMemcpy(h_verdict, d_v, (size_t)n * 4, cudaMemcpyDeviceToHost)))
        rcode = check_cuda(ctx, e, "rc_kat_pairs");
    ctx->launches++;
    cudaFree(d_r); cudaFree(d_o); cudaFree(d_out); cudaFree(d_v);
    return rcode;
}

int rc_kat_la_solve(rc_ctx *ctx, const float *h_in6, int n, float *h_out2)
{
    if (!ctx || n < 0 || (n > 0 && (!h_in6 || !h_out2))) return RC_ERR_INVALID;
    if (n == 0) return RC_OK;
    cudaSetDevice(ctx->device);
    float *d_in = nullptr, *d_out = nullptr;
    cudaError_t e;
    int rcode = RC_OK;
    if ((e = cudaMalloc(&d_in, (size_t)n * 24)) || (e = cudaMalloc(&d_out, (size_t)n * 8)) ||
        (e = cudaMemcpy(d_in, h_in6, (size_t)n * 24, cudaMemcpyHostToDevice)) ||
        (e = rc::launch_kat_la_solve(d_in, n, d_out, ctx->own_stream)) ||
        (e = cudaStreamSynchronize(ctx->own_stream)) ||
        (e = cudaMemcpy(h_out2, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost)))
        rcode = check_cuda(ctx, e, "rc_kat_la_solve");
    ctx->launches++;
    cudaFree(d_in); cudaFree(d_out);
    return rcode;
}

// ---- coplanar pairs ------------------------------------------------------------

int rc_set_coplanar_mode(rc_ctx *ctx, int mode)
{
    if (!ctx) return RC_ERR_INVALID;
    if (mode != RC_COPLANAR_SKIP && mode != RC_COPLANAR_DECIDE)
        return set_error(ctx, RC_ERR_INVALID, "coplanar mode must be RC_COPLANAR_SKIP or RC_COPLANAR_DECIDE");
    ctx->coplanar_mode = mode;
    return RC_OK;
}

int rc_coplanar_mode(const rc_ctx *ctx) { return ctx ? ctx->coplanar_mode : RC_ERR_INVALID; }

// ---- measurement --------------------------------------------------------------------------

int rc_stats_enable(rc_ctx *ctx, int on)
{
    if (!ctx) return RC_ERR_INVALID;
    ctx->stats_on = on ? 1 : 0;
    return RC_OK;
}

int rc_stats_trace(rc_ctx *ctx, uint32_t *d_trace)
{
    if (!ctx) return RC_ERR_INVALID;
    ctx->d_trace = d_trace;
    return RC_OK;
}

int rc_stats_read(rc_ctx *ctx, rc_stats *out, int reset)
{
    if (!ctx || !out) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    unsigned long long h[RC_NUM_COUNTERS];
    cudaError_t e;
    if ((e = cudaDeviceSynchronize()) ||
        (e = cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost)))
        return check_cuda(ctx, e, "rc_stats_read");
    out->broad_cleared = h[CNT_BROAD_CLEARED]; out->sphere_cleared = h[CNT_SPHERE_CLEARED];
    out->narrow_valid = h[CNT_NARROW_VALID]; out->narrow_invalid = h[CNT_NARROW_INVALID];
    out->node_pairs = h[CNT_NODE_PAIRS]; out->leaf_pairs = h[CNT_LEAF_PAIRS];
    out->tri_pairs = h[CNT_TRI_PAIRS]; out->tail_pairs = h[CNT_TAIL_PAIRS];
    out->overflow = h[CNT_OVERFLOW]; out->launches = ctx->launches;
    out->node_rounds = h[CNT_NODE_ROUNDS]; out->leaf_rounds = h[CNT_LEAF_ROUNDS];
    out->tail_rounds = h[CNT_TAIL_ROUNDS];
    out->shared_pieces = h[CNT_SHARED];
    out->watchdog = h[CNT_WATCHDOG]; out->watchdog_where = h[CNT_WATCHDOG_WHERE];
    out->prefilter_pairs = h[CNT_PREFILTER]; out->plane1_pairs = h[CNT_PLANE1];
    if (reset) {
        ctx->launches = 0;
        if ((e = cudaMemset(ctx->d_counters, 0, sizeof h)) || (e = cudaDeviceSynchronize()))
            return check_cuda(ctx, e, "rc_stats reset");
    }
    return RC_OK;
}

int rc_measure_fp32_peak(rc_ctx *ctx, double *tflops)
{
    if (!ctx || !tflops) return RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    const int grid = ctx->sm_count * 8, iters = 4096;
    float *d = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&d, (size_t)grid * 256 * 4))) return check_cuda(ctx, e, "fp32 peak alloc");
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(ctx->ev0, ctx->own_stream);
        e = launch_fma_peak(d, grid, iters, ctx->own_stream);
        cudaEventRecord(ctx->ev1, ctx->own_stream);
        if (e || (e = cudaStreamSynchronize(ctx->own_stream))) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        if (rep > 0) best = std::min(best, ms);
    }
    cudaFree(d);
    if (e) return check_cuda(ctx, e, "fp32 peak");
    const double fmas = (double)grid * 256 * iters * 64;
    *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    return RC_OK;
}

}  // extern "C"
