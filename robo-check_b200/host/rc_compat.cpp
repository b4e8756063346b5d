// rc_compat.cpp -- reference-signature shims over the C ABI (see include/robocheck_compat.hpp).
// Error behaviour follows the reference: I/O failures throw std::runtime_error where the
// reference throws (Utils_rai.cu:39,46,78) or print to stderr and return where it prints
// (Utils_rai.cu:141-145); CUDA failures are reported on stderr and leave the outputs untouched
// (the reference's checkCudaCall logs and continues, Utils_rai.h:29-36).
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "../../include/robocheck_compat.hpp"
#include "rc_mesh.h"

static_assert(sizeof(Configuration) == sizeof(rc_config), "Configuration layout");
static_assert(sizeof(Vector3f) == sizeof(rc_vec3), "Vector3f layout");
static_assert(sizeof(Triangle) == sizeof(rc_tri), "Triangle layout");
static_assert(sizeof(AABB) == sizeof(rc_aabb), "AABB layout");
static_assert(sizeof(bool) == 1, "bool flags are bytes");

namespace {

unsigned g_seed = 508;
rc_ctx *g_ctx = nullptr;
std::string g_ctx_key;

rc_mesh *load_any(const std::string &path)
{
    rc_mesh *m = nullptr;
    const bool obj = path.size() > 4 && path.compare(path.size() - 4, 4, ".obj") == 0;
    const int rc = obj ? rc_mesh_load_obj(path.c_str(), &m) : rc_mesh_load_byu(path.c_str(), &m);
    if (rc != RC_OK) {
        fprintf(stderr, "robocheck: cannot load mesh %s: %s\n", path.c_str(), rc_strerror(rc));
        return nullptr;
    }
    return m;
}

rc_ctx *default_context()
{
    const char *er = getenv("ROBOCHECK_ROB_FILE"), *eo = getenv("ROBOCHECK_OBS_FILE");
    const std::string rob = er ? er : ROB_FILE, obs = eo ? eo : OBS_FILE;
    const std::string key = rob + "\n" + obs;
    if (g_ctx && key == g_ctx_key) return g_ctx;
    if (g_ctx) { rc_destroy(g_ctx); g_ctx = nullptr; }
    rc_mesh *r = load_any(rob), *o = load_any(obs);
    if (r && o) {
        const int rc = rc_create_from_meshes(&g_ctx, r, o, 0);
        if (rc != RC_OK) {
            fprintf(stderr, "robocheck: rc_create failed: %s\n", rc_strerror(rc));
            g_ctx = nullptr;
        }
    }
    rc_mesh_free(r);
    rc_mesh_free(o);
    g_ctx_key = key;
    return g_ctx;
}

void report(const char *what, int rc, rc_ctx *ctx)
{
    if (rc != RC_OK)
        fprintf(stderr, "robocheck: %s failed: %s (%s)\n", what, rc_strerror(rc), ctx ? rc_last_error(ctx) : "");
}

}  // namespace

void robocheck_set_seed(unsigned seed) { g_seed = seed; }

void broadPhaseFused_sep(std::vector<Configuration> &confs, bool *valid_conf)
{
    rc_ctx *ctx = default_context();
    if (!ctx) return;
    report("broadPhaseFused_sep",
           rc_check_host(ctx, reinterpret_cast<const rc_config *>(confs.data()), (long)confs.size(),
                         reinterpret_cast<uint8_t *>(valid_conf), RC_MODE_CULLED),
           ctx);
}

void broadPhaseFused(std::vector<Configuration> &confs, bool *valid_conf) { broadPhaseFused_sep(confs, valid_conf); }

void CallMegaKernel(std::vector<Configuration> configs, bool *valid_confs) { broadPhaseFused_sep(configs, valid_confs); }

void generateAABB(Vector3f *vertices, unsigned int numVertices, unsigned int numConfigs, AABB *botBounds)
{
    rc_ctx *ctx = default_context();
    if (!ctx) return;
    report("generateAABB",
           rc_generate_aabb_host(ctx, reinterpret_cast<const rc_vec3 *>(vertices), numVertices, numConfigs,
                                 reinterpret_cast<rc_aabb *>(botBounds)),
           ctx);
}

void generateAABBBaseline(std::vector<float> &x, std::vector<float> &y, std::vector<float> &z, AABB *b)
{
    AABB a = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (size_t v = 0; v < x.size(); v++) {
        a.x_min = std::min(a.x_min, x[v]); a.x_max = std::max(a.x_max, x[v]);
        a.y_min = std::min(a.y_min, y[v]); a.y_max = std::max(a.y_max, y[v]);
        a.z_min = std::min(a.z_min, z[v]); a.z_max = std::max(a.z_max, z[v]);
    }
    *b = a;
}

void generateAABBBaseline(Vector3f *vertices, unsigned int numVertices, unsigned int numConfigs, AABB *botBounds)
{
    for (unsigned i = 0; i < numConfigs; i++) {
        AABB a = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (unsigned v = 0; v < numVertices; v++) {
            const Vector3f &p = vertices[(size_t)i * numVertices + v];
            a.x_min = std::min(a.x_min, p.x); a.x_max = std::max(a.x_max, p.x);
            a.y_min = std::min(a.y_min, p.y); a.y_max = std::max(a.y_max, p.y);
            a.z_min = std::min(a.z_min, p.z); a.z_max = std::max(a.z_max, p.z);
        }
        botBounds[i] = a;
    }
}

void broadPhase(int num_confs, const AABB *robots, const AABB *obstacle, bool *valid_conf)
{
    rc_ctx *ctx = default_context();
    if (!ctx) return;
    report("broadPhase",
           rc_broad_phase(ctx, num_confs, reinterpret_cast<const rc_aabb *>(robots),
                          reinterpret_cast<const rc_aabb *>(obstacle), reinterpret_cast<uint8_t *>(valid_conf),
                          nullptr),
           ctx);
}

void narrowPhase(int num_confs, int num_rob_trs, int num_rob_pts, int num_obs_trs, int num_obs_pts,
                 const Triangle *rob_trs, const Vector3f *rob_pts, const Triangle *obs_trs,
                 const Vector3f *obs_pts, bool *valid_conf)
{
    report("narrowPhase",
           rc_narrow_phase_host(num_confs, num_rob_trs, num_rob_pts, num_obs_trs, num_obs_pts,
                                reinterpret_cast<const rc_tri *>(rob_trs), reinterpret_cast<const rc_vec3 *>(rob_pts),
                                reinterpret_cast<const rc_tri *>(obs_trs), reinterpret_cast<const rc_vec3 *>(obs_pts),
                                reinterpret_cast<uint8_t *>(valid_conf), 0),
           nullptr);
}

void narrowPhase_unopt(int num_confs, int num_rob_trs, int num_rob_pts, int num_obs_trs, int num_obs_pts,
                       const Triangle *rob_trs, const Vector3f *rob_pts, const Triangle *obs_trs,
                       const Vector3f *obs_pts, bool *valid_conf)
{
    narrowPhase(num_confs, num_rob_trs, num_rob_pts, num_obs_trs, num_obs_pts, rob_trs, rob_pts, obs_trs,
                obs_pts, valid_conf);
}

void narrowPhaseBaseline(int num_confs, int num_rob_trs, int num_rob_pts, int num_obs_trs, int num_obs_pts,
                         const Triangle *rob_trs, const Vector3f *rob_pts, const Triangle *obs_trs,
                         const Vector3f *obs_pts, bool *valid_conf)
{
    // narrow-phase.cu:346-423 decides every configuration, whatever valid_conf held
    if (num_confs > 0 && valid_conf) memset(valid_conf, 0, (size_t)num_confs);
    narrowPhase(num_confs, num_rob_trs, num_rob_pts, num_obs_trs, num_obs_pts, rob_trs, rob_pts, obs_trs,
                obs_pts, valid_conf);
}

void broadPhaseBaseline(int num_confs, const AABB *robots, const AABB *obstacle, bool *valid_conf)
{
    if (!robots || !obstacle || !valid_conf) return;
    for (int i = 0; i < num_confs; i++) {
        const AABB &r = robots[i], &o = *obstacle;
        const bool hit = (o.x_min <= r.x_max && r.x_min <= o.x_max) && (o.y_min <= r.y_max && r.y_min <= o.y_max) &&
                         (o.z_min <= r.z_max && r.z_min <= o.z_max);
        valid_conf[i] = !hit;
    }
}

// ---- host utilities ---------------------------------------------------------------------------

void loadOBJFile(const char *filename, std::vector<float> &x, std::vector<float> &y, std::vector<float> &z,
                 std::vector<int> &v1, std::vector<int> &v2, std::vector<int> &v3)
{
    rc_mesh *m = nullptr;
    if (rc_mesh_load_obj(filename, &m) != RC_OK) {
        fprintf(stderr, "file not exist\n");  // Utils_rai.cu:232-236
        return;
    }
    x.insert(x.end(), m->x.begin(), m->x.end());
    y.insert(y.end(), m->y.begin(), m->y.end());
    z.insert(z.end(), m->z.begin(), m->z.end());
    v1.insert(v1.end(), m->v1.begin(), m->v1.end());
    v2.insert(v2.end(), m->v2.begin(), m->v2.end());
    v3.insert(v3.end(), m->v3.begin(), m->v3.end());
    rc_mesh_free(m);
}

void loadOBJFile(const char *filename, std::vector<Vector3f> &points, std::vector<Triangle> &triangles)
{
    rc_mesh *m = nullptr;
    if (rc_mesh_load_obj(filename, &m) != RC_OK) {
        fprintf(stderr, "file not exist\n");
        return;
    }
    for (size_t v = 0; v < m->x.size(); v++) points.push_back(Vector3f(m->x[v], m->y[v], m->z[v]));
    for (size_t t = 0; t < m->v1.size(); t++) triangles.push_back(Triangle{m->v1[t], m->v2[t], m->v3[t]});
    rc_mesh_free(m);
}

void readConfigurationFromFileTagged(const std::string &filename, std::vector<ConfigurationTagged> &confs)
{
    rc_config *c = nullptr;
    uint8_t *t = nullptr;
    long n = 0;
    if (rc_conf_read(filename.c_str(), &c, &t, &n) != RC_OK)
        throw std::runtime_error("Failed to open file " + filename);
    for (long i = 0; i < n; i++)
        confs.push_back(ConfigurationTagged{c[i].x, c[i].y, c[i].z, c[i].pitch, c[i].yaw, c[i].roll, t[i] != 0});
    rc_free(c);
    rc_free(t);
}

void readConfigurationFromFile(const std::string &filename, std::vector<Configuration> &confs)
{
    rc_config *c = nullptr;
    long n = 0;
    if (rc_conf_read(filename.c_str(), &c, nullptr, &n) != RC_OK)
        throw std::runtime_error("Failed to open file " + filename);
    for (long i = 0; i < n; i++)
        confs.push_back(Configuration{c[i].x, c[i].y, c[i].z, c[i].pitch, c[i].yaw, c[i].roll});
    rc_free(c);
}

void writeConfigurationToFileTagged(const std::vector<ConfigurationTagged> &confs, const std::string &filename)
{
    std::vector<rc_config> c(confs.size());
    std::vector<uint8_t> v(confs.size());
    for (size_t i = 0; i < confs.size(); i++) {
        c[i] = rc_config{confs[i].x, confs[i].y, confs[i].z, confs[i].pitch, confs[i].yaw, confs[i].roll};
        v[i] = confs[i].valid ? 1 : 0;
    }
    if (rc_conf_write(filename.c_str(), c.data(), v.data(), (long)c.size()) != RC_OK)
        throw std::runtime_error("Failed to open file " + filename);
}

ConfigurationTagged makeTagged(const Configuration &c)
{
    return ConfigurationTagged{c.x, c.y, c.z, c.pitch, c.yaw, c.roll, false};
}

void generateConfs(std::vector<Configuration> &confs, float x_min, float x_max, float y_min, float y_max,
                   float z_min, float z_max, int num_confs)
{
    // like the reference (Utils_rai.cu:347 `confs[i] = conf`), the vector must already hold
    // num_confs elements
    rc_conf_generate(reinterpret_cast<rc_config *>(confs.data()), num_confs, x_min, x_max, y_min, y_max,
                     z_min, z_max, g_seed);
}

void createAlphaBotConfigurations(std::vector<Configuration> &confs, int num_confs, bool hard)
{
    rc_conf_generate_alpha(reinterpret_cast<rc_config *>(confs.data()), num_confs, hard ? 1 : 0, g_seed);
}
