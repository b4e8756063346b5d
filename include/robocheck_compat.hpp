// robocheck_compat.hpp -- the reference's own C++ entry points, re-exposed on top of the C ABI
// (include/robocheck.h) so that the reference's call sites compile and link unchanged against
// librobocheck.so.  Each declaration cites the reference header it mirrors.  Types are
// layout-identical to the reference's PODs (Utils_rai.h:51-96).
#pragma once
#include <string>
#include <vector>

#include "robocheck.h"

#ifndef ROB_FILE  // Utils_rai.h:18-19
#define ROB_FILE "./models/alpha1.0/robot.obj"
#endif
#ifndef OBS_FILE
#define OBS_FILE "./models/alpha1.0/obstacle.obj"
#endif

struct Configuration { float x, y, z, pitch, yaw, roll; };                    // Utils_rai.h:51-58
struct ConfigurationTagged { float x, y, z, pitch, yaw, roll; bool valid; };  // Utils_rai.h:60-68
struct Vector3f {                                                              // Utils_rai.h:78-82
    float x, y, z;
    Vector3f(float _x, float _y, float _z) : x(_x), y(_y), z(_z) {}
    Vector3f() {}
};
struct Triangle { int v1, v2, v3; };                                           // Utils_rai.h:84-86
struct AABB { float x_min, y_min, z_min, x_max, y_max, z_max; };               // Utils_rai.h:88-96

// Utils_rai.h:131-148 -- host utilities
void writeConfigurationToFileTagged(const std::vector<ConfigurationTagged> &confs, const std::string &filename);
void readConfigurationFromFileTagged(const std::string &filename, std::vector<ConfigurationTagged> &confs);
void readConfigurationFromFile(const std::string &filename, std::vector<Configuration> &confs);
ConfigurationTagged makeTagged(const Configuration &conf);
void createAlphaBotConfigurations(std::vector<Configuration> &confs, int num_confs, bool hard);
void loadOBJFile(const char *filename, std::vector<Vector3f> &points, std::vector<Triangle> &triangles);
void loadOBJFile(const char *filename, std::vector<float> &x, std::vector<float> &y, std::vector<float> &z,
                 std::vector<int> &v1, std::vector<int> &v2, std::vector<int> &v3);
void generateConfs(std::vector<Configuration> &confs, float x_min, float x_max, float y_min, float y_max,
                   float z_min, float z_max, int num_confs);
// The reference seeds from std::random_device; set this (default 508) for reproducible sets.
void robocheck_set_seed(unsigned seed);

// broad-phase/broad-phase-fused.hu:9-11 -- THE live entry point: host vector in, caller-allocated
// bool[confs.size()] out (true = valid).  Meshes are ROB_FILE / OBS_FILE (OBJ or BYU .g),
// overridable at run time with the environment variables ROBOCHECK_ROB_FILE / ROBOCHECK_OBS_FILE;
// the device context is built on first use and kept (the reference re-parses both files and
// re-allocates everything on every call).
void broadPhaseFused_sep(std::vector<Configuration> &confs, bool *valid_conf);
void broadPhaseFused(std::vector<Configuration> &confs, bool *valid_conf);
// test/MegaKernel.hu:9
void CallMegaKernel(std::vector<Configuration> configs, bool *valid_confs);

// generate-AABB/generate-AABB.hu:14-21
void generateAABB(Vector3f *vertices, unsigned int numVertices, unsigned int numConfigs, AABB *botBounds);
void generateAABBBaseline(std::vector<float> &x, std::vector<float> &y, std::vector<float> &z, AABB *botBounds);
void generateAABBBaseline(Vector3f *vertices, unsigned int numVertices, unsigned int numConfigs, AABB *botBounds);
// broad-phase/broad-phase.hu:13 -- device pointers, asynchronous on the default stream
void broadPhase(int num_confs, const AABB *robots, const AABB *obstacle, bool *valid_conf);
// broad-phase/broad-phase.hu:17 -- host pointers.  The reference declares it and never defines it;
// here it is the inclusive per-axis overlap test of the kernel (broad-phase.cu:10-28) on the host.
void broadPhaseBaseline(int num_confs, const AABB *robots, const AABB *obstacle, bool *valid_conf);
// narrow-phase/narrow-phase.hu:22-33 -- host pointers, valid_conf in/out (true entries skipped)
void narrowPhase(int num_confs, int num_rob_trs, int num_rob_pts, int num_obs_trs, int num_obs_pts,
                 const Triangle *rob_trs, const Vector3f *rob_pts, const Triangle *obs_trs,
                 const Vector3f *obs_pts, bool *valid_conf);
void narrowPhase_unopt(int num_confs, int num_rob_trs, int num_rob_pts, int num_obs_trs, int num_obs_pts,
                       const Triangle *rob_trs, const Vector3f *rob_pts, const Triangle *obs_trs,
                       const Vector3f *obs_pts, bool *valid_conf);
// narrow-phase/narrow-phase.hu:37 -- the reference's CPU baseline: writes EVERY valid_conf[i] (no
// in/out skipping, narrow-phase.cu:346-423).  There is no CPU path in this library, so the call
// runs the exhaustive GPU sweep on all configurations.  The arithmetic is the live CUDA kernel's,
// which differs from the reference's host code in four documented places (FMA
// contraction, the float/double threshold of no_overlap, coplanarity tested after the second plane
// test, `>=` instead of `>` in the interval test); on alpha1.0 the two agree on every configuration
// of the golden sets.  Coplanar pairs are skipped silently (the reference prints a line per
// configuration, :419-420).
void narrowPhaseBaseline(int num_confs, int num_rob_trs, int num_rob_pts, int num_obs_trs, int num_obs_pts,
                         const Triangle *rob_trs, const Vector3f *rob_pts, const Triangle *obs_trs,
                         const Vector3f *obs_pts, bool *valid_conf);
