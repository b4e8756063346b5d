// robocheck_cli.cpp -- C++ harness with the role of the reference's only complete entry point,
// test/full-integration-test.cu:652-732 (main: read .conf -> GPU pipeline -> compare with the
// tags in the file -> report).  It is written against the reference's own function signatures
// (include/robocheck_compat.hpp), so it doubles as the proof that those call sites link against
// librobocheck.so.  FCL is not available in this image, so the comparison uses the FCL verdicts
// persisted in column 7 of the .conf file instead of a live FCL run.
//
//   robocheck_cli <file.conf> [robot mesh] [obstacle mesh]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "../../include/robocheck_compat.hpp"

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s <file.conf> [robot.obj|.g] [obstacle.obj|.g]\n", argv[0]);
        return 2;
    }
    if (argc > 2) setenv("ROBOCHECK_ROB_FILE", argv[2], 1);
    if (argc > 3) setenv("ROBOCHECK_OBS_FILE", argv[3], 1);
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
    broadPhaseFused_sep(confs, valid.get());  // builds the context (one-time cost)
    const auto t0 = std::chrono::high_resolution_clock::now();
    broadPhaseFused_sep(confs, valid.get());
    const auto t1 = std::chrono::high_resolution_clock::now();
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
