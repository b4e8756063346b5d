// ref_harness.cu -- glue that lets the tests and bench.py call the UNMODIFIED
// reference (compiled from the sources where they lie under /root/reference by
// oracle/Makefile, outputs only into oracle/_ref/).  TEST INFRASTRUCTURE ONLY.
//
// The reference ships no library and no FFI: its entry points are C++ functions
// called from test drivers, and its mesh tables are `extern __constant__` symbols
// that the *test* translation unit has to define
// (test/full-integration-test.cu:8-19).  This file is that translation unit.
//
//   ref_gpu_check      -> broadPhaseFused_sep()            broad-phase-fused.cu:365-520
//   ref_gpu_transform  -> broadPhaseFusedKernel_sep<<<>>>  broad-phase-fused.cu:157-207
//   ref_cpu_narrow     -> narrowPhaseBaseline()            narrow-phase.cu:342-424
//
// broadPhaseFused_sep() opens "./models/alpha1.0/{robot,obstacle}.obj" relative to
// the working directory (Utils_rai.h:18-19), so callers chdir into a directory that
// holds OBJ files written from this repo's own assets (tests/refpipe.py).
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <unistd.h>
#include <vector>

#include "Utils_rai.h"
#include "broad-phase/broad-phase-fused.hu"
#include "narrow-phase/narrow-phase.hu"

__constant__ float base_rob_x[NUM_ROB_VERTICES];
__constant__ float base_rob_y[NUM_ROB_VERTICES];
__constant__ float base_rob_z[NUM_ROB_VERTICES];
__constant__ int base_rob_tri_v1[MAX_NUM_ROBOT_TRIANGLES];
__constant__ int base_rob_tri_v2[MAX_NUM_ROBOT_TRIANGLES];
__constant__ int base_rob_tri_v3[MAX_NUM_ROBOT_TRIANGLES];
__constant__ float base_obs_x[NUM_ROB_VERTICES];
__constant__ float base_obs_y[NUM_ROB_VERTICES];
__constant__ float base_obs_z[NUM_ROB_VERTICES];
__constant__ int base_obs_tri_v1[MAX_NUM_ROBOT_TRIANGLES];
__constant__ int base_obs_tri_v2[MAX_NUM_ROBOT_TRIANGLES];
__constant__ int base_obs_tri_v3[MAX_NUM_ROBOT_TRIANGLES];

// defined in broad-phase-fused.cu:157, not declared in its header
__global__ void broadPhaseFusedKernel_sep(Configuration *configs, const AABB *obstacle,
                                          float *rob_pts_x, float *rob_pts_y, float *rob_pts_z,
                                          bool *valid_conf, const int num_configs,
                                          const int num_robot_vertices);

namespace {
// the reference chats on std::cout from inside the entry point; keep the test logs clean
struct QuietCout {
    std::streambuf *old;
    std::ostringstream sink;
    QuietCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~QuietCout() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

// The reference indexes `i * num_rob_pts` in int (narrow-phase.cu:588): more than
// 2,711,541 configurations per call overflow.  Callers chunk; this guards.
int ref_gpu_max_confs_per_call() { return 2000000; }

// Whole live pipeline, exactly as test/full-integration-test.cu:688 calls it.
// `workdir` must contain models/alpha1.0/{robot,obstacle}.obj.  flags: 1 = valid.
int ref_gpu_check(const char *workdir, const Configuration *confs, long n, unsigned char *flags)
{
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return -1;
    if (chdir(workdir) != 0) return -2;
    int rc = 0;
    {
        QuietCout quiet;
        const long step = ref_gpu_max_confs_per_call();
        for (long off = 0; off < n; off += step) {
            const long m = (n - off < step) ? (n - off) : step;
            std::vector<Configuration> v(confs + off, confs + off + m);
            static_assert(sizeof(bool) == 1, "bool flags");
            broadPhaseFused_sep(v, reinterpret_cast<bool *>(flags + off));
            if (cudaGetLastError() != cudaSuccess) rc = -3;
        }
    }
    if (chdir(cwd) != 0) return -4;
    return rc;
}

// Transformed vertices + broad-phase flags from the reference's fused kernel with
// its own launch shape (broad-phase-fused.cu:478-484).  Mesh SoA arrays are passed
// in (n_v must be NUM_ROB_VERTICES).  Outputs are host arrays [n * n_v] / [n].
int ref_gpu_transform(const float *rx, const float *ry, const float *rz, int n_v,
                      const float obs_aabb[6], const Configuration *confs, int n, float *ox,
                      float *oy, float *oz, unsigned char *flags)
{
    if (n_v != NUM_ROB_VERTICES) return -1;
    cudaMemcpyToSymbol(base_rob_x, rx, n_v * sizeof(float));
    cudaMemcpyToSymbol(base_rob_y, ry, n_v * sizeof(float));
    cudaMemcpyToSymbol(base_rob_z, rz, n_v * sizeof(float));
    float *dx, *dy, *dz;
    bool *dv;
    AABB *dbox;
    Configuration *dc;
    const size_t vb = (size_t)n * n_v * sizeof(float);
    if (cudaMalloc(&dx, vb) || cudaMalloc(&dy, vb) || cudaMalloc(&dz, vb)) return -2;
    cudaMalloc(&dv, n);
    cudaMalloc(&dbox, sizeof(AABB));
    cudaMalloc(&dc, (size_t)n * sizeof(Configuration));
    cudaMemcpy(dbox, obs_aabb, sizeof(AABB), cudaMemcpyHostToDevice);
    cudaMemcpy(dc, confs, (size_t)n * sizeof(Configuration), cudaMemcpyHostToDevice);
    broadPhaseFusedKernel_sep<<<(n + 31) / 32, 32>>>(dc, dbox, dx, dy, dz, dv, n, n_v);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(ox, dx, vb, cudaMemcpyDeviceToHost);
    cudaMemcpy(oy, dy, vb, cudaMemcpyDeviceToHost);
    cudaMemcpy(oz, dz, vb, cudaMemcpyDeviceToHost);
    cudaMemcpy(flags, dv, n, cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dy); cudaFree(dz); cudaFree(dv); cudaFree(dbox); cudaFree(dc);
    return e == cudaSuccess ? 0 : -3;
}

// Reference CPU brute force.  AoS inputs as the reference declares them; rob_pts is
// [n][n_rob_v] already transformed; flags in/out (every entry is overwritten).
int ref_cpu_narrow(int n, int n_rob_t, int n_rob_v, int n_obs_t, int n_obs_v, const int *rob_tri,
                   const float *rob_pts, const int *obs_tri, const float *obs_pts,
                   unsigned char *flags)
{
    static_assert(sizeof(Triangle) == 12 && sizeof(Vector3f) == 12, "POD layout");
    narrowPhaseBaseline(n, n_rob_t, n_rob_v, n_obs_t, n_obs_v,
                        reinterpret_cast<const Triangle *>(rob_tri),
                        reinterpret_cast<const Vector3f *>(rob_pts),
                        reinterpret_cast<const Triangle *>(obs_tri),
                        reinterpret_cast<const Vector3f *>(obs_pts),
                        reinterpret_cast<bool *>(flags));
    return 0;
}

}  // extern "C"
