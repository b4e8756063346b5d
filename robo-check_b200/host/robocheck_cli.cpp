// robocheck_cli.cpp -- C++ harness with the role of the reference's only complete entry point,
// test/full-integration-test.cu:652-732 (main: read .conf -> GPU pipeline -> compare with the
// tags in the file -> report).  It is written against the reference's own function signatures
// (include/robocheck_compat.hpp), so it doubles as the proof that those call sites link against
// librobocheck.so.  FCL is not available in this image, so the comparison uses the FCL verdicts
// persisted in column 7 of the .conf file instead of a live FCL run.
//
//   robocheck_cli <file.conf> [robot mesh] [obstacle mesh] [--gpus N]
// With --gpus N the check goes through rc_check_multi of the C ABI (N contexts, one per device,
// device index modulo the number of devices present): shards, concurrent devices, gathered bitmask.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include <cstring>
#include <vector>

#include "../../include/robocheck.h"
#include "../../include/robocheck_compat.hpp"

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s <file.conf> [robot.obj|.g] [obstacle.obj|.g]\n", argv[0]);
        return 2;
    }
    int n_gpus = 0;
    std::vector<const char *> pos;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--gpus") && i + 1 < argc) n_gpus = atoi(argv[++i]);
        else pos.push_back(argv[i]);
    }
    if (pos.empty()) return 2;
    argv[1] = const_cast<char *>(pos[0]);
    if (pos.size() > 1) setenv("ROBOCHECK_ROB_FILE", pos[1], 1);
    if (pos.size() > 2) setenv("ROBOCHECK_OBS_FILE", pos[2], 1);
    std::vector<ConfigurationTagged> tagged;
    std::vector<Configuration> confs;
    try {
        readConfigurationFromFileTagged(argv[1], tagged);
        readConfigurationFromFile(argv[1], confs);
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    std::unique_ptr<bool[]> valid(new bool[confs.size() + 1]());
    auto t0 = std::chrono::high_resolution_clock::now(), t1 = t0;
    if (n_gpus > 0) {
        // the C ABI from C++: meshes, one context per device, sharded check, bitmask back
        const char *rob_path = getenv("ROBOCHECK_ROB_FILE"), *obs_path = getenv("ROBOCHECK_OBS_FILE");
        auto load = [](const char *path, rc_mesh **m) {
            const size_t len = strlen(path);
            return (len > 2 && !strcmp(path + len - 2, ".g")) ? rc_mesh_load_byu(path, m) : rc_mesh_load_obj(path, m);
        };
        rc_mesh *rob = nullptr, *obs = nullptr;
        if (!rob_path || !obs_path || load(rob_path, &rob) != RC_OK || load(obs_path, &obs) != RC_OK) {
            fprintf(stderr, "--gpus needs the robot and obstacle mesh files\n");
            return 1;
        }
        int n_dev = 0;
        if (rc_device_count(&n_dev) != RC_OK || n_dev <= 0) { fprintf(stderr, "no CUDA device\n"); return 1; }
        std::vector<rc_ctx *> ctxs((size_t)n_gpus, nullptr);
        for (int g = 0; g < n_gpus; g++) {
            const int code = rc_create_from_meshes(&ctxs[g], rob, obs, g % n_dev);
            if (code != RC_OK) { fprintf(stderr, "rc_create: %s\n", rc_strerror(code)); return 1; }
        }
        const long n = (long)confs.size();
        std::vector<uint32_t> mask((size_t)(n + 31) / 32);
        static_assert(sizeof(Configuration) == sizeof(rc_config), "layout");
        const rc_config *h = reinterpret_cast<const rc_config *>(confs.data());
        int code = rc_check_multi(ctxs.data(), n_gpus, h, n, mask.data(), nullptr, nullptr);  // warm-up
        t0 = std::chrono::high_resolution_clock::now();
        if (code == RC_OK) code = rc_check_multi(ctxs.data(), n_gpus, h, n, mask.data(), nullptr, nullptr);
        t1 = std::chrono::high_resolution_clock::now();
        if (code != RC_OK) { fprintf(stderr, "rc_check_multi: %s\n", rc_strerror(code)); return 1; }
        for (long i = 0; i < n; i++) valid[i] = (mask[i >> 5] >> (i & 31)) & 1u;
        for (rc_ctx *c : ctxs) rc_destroy(c);
        rc_mesh_free(rob); rc_mesh_free(obs);
        printf("rc_check_multi over %d contexts on %d device(s), shards of %d configurations\n", n_gpus, n_dev,
               rc_shard_size(n, n_gpus));
    } else {
        broadPhaseFused_sep(confs, valid.get());  // builds the context (one-time cost)
        t0 = std::chrono::high_resolution_clock::now();
        broadPhaseFused_sep(confs, valid.get());
        t1 = std::chrono::high_resolution_clock::now();
    }
    size_t n_valid = 0, mismatches = 0;
    for (size_t i = 0; i < confs.size(); i++) {
        n_valid += valid[i];
        mismatches += valid[i] != tagged[i].valid;
    }
    const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    printf("%zu configurations, %zu valid, %zu invalid, %.3f ms (%.1f configs/s)\n", confs.size(), n_valid,
           confs.size() - n_valid, ms, confs.size() / (ms * 1e-3));
    printf("%zu disagreements with the tags in the file%s\n", mismatches, mismatches ? "  <-- KERNEL BROKEN" : "");
    return mismatches ? 3 : 0;
}
