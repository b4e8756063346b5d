// rc_host.cpp -- host side of the hot path's callers: mesh loaders, .conf reader/writer and
// configuration generator.  Clean-room C++ with the behaviour of the reference's utilities:
//   OBJ loader     Utils_rai.cu:139-320   (v / vn / vt lines, `a//n` face tokens, fan split)
//   BYU loader     models/BYU_Model_Loader.cpp:33-89, models/byu-to-obj.cpp:27-88
//   .conf I/O      Utils_rai.cu:11-97
//   generators     Utils_rai.cu:111-136, 322-348
// No CUDA in this file.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/robocheck.h"
#include "rc_mesh.h"

namespace {

bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }

// split a line into whitespace separated tokens (the reference uses strtok with "\t \r\n")
std::vector<std::string> tokens_of(const std::string &line)
{
    std::vector<std::string> out;
    size_t i = 0;
    while (i < line.size()) {
        while (i < line.size() && is_blank(line[i])) i++;
        size_t j = i;
        while (j < line.size() && !is_blank(line[j])) j++;
        if (j > i) out.emplace_back(line.substr(i, j - i));
        i = j;
    }
    return out;
}

}  // namespace

extern "C" {

int rc_mesh_load_obj(const char *path, rc_mesh **out)
{
    if (!path || !out) return RC_ERR_INVALID;
    std::ifstream in(path, std::ios::binary);
    if (!in.good()) return RC_ERR_IO;
    rc_mesh *m = new rc_mesh();
    std::string line;
    while (std::getline(in, line)) {
        const std::vector<std::string> tok = tokens_of(line);
        if (tok.empty() || tok[0][0] == '#') continue;
        const std::string &head = tok[0];
        if (head[0] == 'v') {
            if (head.size() > 1 && (head[1] == 'n' || head[1] == 't')) continue;  // normals / uvs
            if (tok.size() < 4) { delete m; return RC_ERR_FORMAT; }
            // the reference converts with (float)atof(): decimal -> double -> float
            m->x.push_back((float)atof(tok[1].c_str()));
            m->y.push_back((float)atof(tok[2].c_str()));
            m->z.push_back((float)atof(tok[3].c_str()));
        } else if (head[0] == 'f') {
            // atoi() stops at the first '/', so `a`, `a/t`, `a//n`, `a/t/n` all yield a.
            // A polygon with k corners becomes the fan (0, t+1, t+2), t = 0..k-3.
            const int k = (int)tok.size() - 1;
            if (k < 3) { delete m; return RC_ERR_FORMAT; }
            const int a = atoi(tok[1].c_str()) - 1;
            for (int t = 0; t < k - 2; t++) {
                m->v1.push_back(a);
                m->v2.push_back(atoi(tok[2 + t].c_str()) - 1);
                m->v3.push_back(atoi(tok[3 + t].c_str()) - 1);
            }
        }
    }
    const int nv = (int)m->x.size();
    for (size_t t = 0; t < m->v1.size(); t++) {
        if (m->v1[t] < 0 || m->v1[t] >= nv || m->v2[t] < 0 || m->v2[t] >= nv || m->v3[t] < 0 ||
            m->v3[t] >= nv) {
            delete m;
            return RC_ERR_FORMAT;
        }
    }
    *out = m;
    return RC_OK;
}

int rc_mesh_load_byu(const char *path, rc_mesh **out)
{
    if (!path || !out) return RC_ERR_INVALID;
    std::ifstream in(path);
    if (!in.good()) return RC_ERR_IO;
    long parts = 0, nverts = 0, npolys = 0, nedges = 0;
    if (!(in >> parts >> nverts >> npolys >> nedges)) return RC_ERR_FORMAT;
    if (parts < 0 || nverts < 0 || npolys < 0 || parts > (1 << 24) || nverts > (1L << 30) ||
        npolys > (1L << 30))
        return RC_ERR_FORMAT;
    for (long p = 0; p < parts; p++) {
        long first, last;
        if (!(in >> first >> last)) return RC_ERR_FORMAT;
    }
    rc_mesh *m = new rc_mesh();
    m->x.reserve(nverts); m->y.reserve(nverts); m->z.reserve(nverts);
    for (long v = 0; v < nverts; v++) {
        std::string sx, sy, sz;
        if (!(in >> sx >> sy >> sz)) { delete m; return RC_ERR_FORMAT; }
        m->x.push_back((float)atof(sx.c_str()));
        m->y.push_back((float)atof(sy.c_str()));
        m->z.push_back((float)atof(sz.c_str()));
    }
    // polygons: 1-based indices, the last one of each polygon is negated.  The PMPL loader
    // insists on triangles; like the OBJ path, larger polygons are fan-split here.
    for (long p = 0; p < npolys; p++) {
        std::vector<int> idx;
        long id = 0;
        do {
            if (!(in >> id)) { delete m; return RC_ERR_FORMAT; }
            idx.push_back((int)(id < 0 ? -id : id) - 1);
        } while (id >= 0);
        if (idx.size() < 3) { delete m; return RC_ERR_FORMAT; }
        for (size_t t = 0; t + 2 < idx.size(); t++) {
            m->v1.push_back(idx[0]);
            m->v2.push_back(idx[t + 1]);
            m->v3.push_back(idx[t + 2]);
        }
    }
    for (size_t t = 0; t < m->v1.size(); t++) {
        if (m->v1[t] < 0 || m->v1[t] >= nverts || m->v2[t] < 0 || m->v2[t] >= nverts ||
            m->v3[t] < 0 || m->v3[t] >= nverts) {
            delete m;
            return RC_ERR_FORMAT;
        }
    }
    *out = m;
    return RC_OK;
}

int rc_mesh_from_arrays(const float *x, const float *y, const float *z, int n_v, const int *v1,
                        const int *v2, const int *v3, int n_t, rc_mesh **out)
{
    if (!out || n_v < 0 || n_t < 0 || (n_v && (!x || !y || !z)) || (n_t && (!v1 || !v2 || !v3)))
        return RC_ERR_INVALID;
    for (int t = 0; t < n_t; t++)
        if (v1[t] < 0 || v1[t] >= n_v || v2[t] < 0 || v2[t] >= n_v || v3[t] < 0 || v3[t] >= n_v)
            return RC_ERR_INVALID;
    rc_mesh *m = new rc_mesh();
    m->x.assign(x, x + n_v); m->y.assign(y, y + n_v); m->z.assign(z, z + n_v);
    m->v1.assign(v1, v1 + n_t); m->v2.assign(v2, v2 + n_t); m->v3.assign(v3, v3 + n_t);
    *out = m;
    return RC_OK;
}

// Writers print 9 significant digits so that every float survives the text round trip.
int rc_mesh_save_obj(const rc_mesh *m, const char *path)
{
    if (!m || !path) return RC_ERR_INVALID;
    FILE *f = fopen(path, "w");
    if (!f) return RC_ERR_IO;
    fprintf(f, "# robocheck-b200 mesh\n");
    for (size_t v = 0; v < m->x.size(); v++)
        fprintf(f, "v %.9g %.9g %.9g\n", m->x[v], m->y[v], m->z[v]);
    for (size_t t = 0; t < m->v1.size(); t++)
        fprintf(f, "f %d %d %d\n", m->v1[t] + 1, m->v2[t] + 1, m->v3[t] + 1);
    fclose(f);
    return RC_OK;
}

int rc_mesh_save_byu(const rc_mesh *m, const char *path)
{
    if (!m || !path) return RC_ERR_INVALID;
    FILE *f = fopen(path, "w");
    if (!f) return RC_ERR_IO;
    const size_t nt = m->v1.size();
    fprintf(f, "1 %zu %zu %zu\n1 %zu\n", m->x.size(), nt, 3 * nt, nt);
    for (size_t v = 0; v < m->x.size(); v++)
        fprintf(f, "%.9g %.9g %.9g\n", m->x[v], m->y[v], m->z[v]);
    for (size_t t = 0; t < nt; t++)
        fprintf(f, "%d %d -%d\n", m->v1[t] + 1, m->v2[t] + 1, m->v3[t] + 1);
    fclose(f);
    return RC_OK;
}

int rc_mesh_counts(const rc_mesh *m, int *n_v, int *n_t)
{
    if (!m) return RC_ERR_INVALID;
    if (n_v) *n_v = (int)m->x.size();
    if (n_t) *n_t = (int)m->v1.size();
    return RC_OK;
}
const float *rc_mesh_x(const rc_mesh *m) { return m ? m->x.data() : nullptr; }
const float *rc_mesh_y(const rc_mesh *m) { return m ? m->y.data() : nullptr; }
const float *rc_mesh_z(const rc_mesh *m) { return m ? m->z.data() : nullptr; }
const int *rc_mesh_v1(const rc_mesh *m) { return m ? m->v1.data() : nullptr; }
const int *rc_mesh_v2(const rc_mesh *m) { return m ? m->v2.data() : nullptr; }
const int *rc_mesh_v3(const rc_mesh *m) { return m ? m->v3.data() : nullptr; }
void rc_mesh_free(rc_mesh *m) { delete m; }

int rc_mesh_aabb(const rc_mesh *m, rc_aabb *out)
{
    // generateAABBBaseline(x, y, z, AABB*), generate-AABB/generate-AABB.cu:183-201
    if (!m || !out) return RC_ERR_INVALID;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (size_t v = 0; v < m->x.size(); v++) {
        lo[0] = std::min(lo[0], m->x[v]); hi[0] = std::max(hi[0], m->x[v]);
        lo[1] = std::min(lo[1], m->y[v]); hi[1] = std::max(hi[1], m->y[v]);
        lo[2] = std::min(lo[2], m->z[v]); hi[2] = std::max(hi[2], m->z[v]);
    }
    *out = rc_aabb{lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
    return RC_OK;
}

// 1 -> 4 midpoint subdivision.  Midpoints are shared between the two triangles of an edge
// (keyed on the sorted vertex pair), computed as 0.5f * (a + b) in float.
int rc_mesh_subdivide(const rc_mesh *src, int levels, rc_mesh **out)
{
    if (!src || !out || levels < 0 || levels > 8) return RC_ERR_INVALID;
    rc_mesh *cur = new rc_mesh(*src);
    for (int l = 0; l < levels; l++) {
        rc_mesh *nxt = new rc_mesh();
        nxt->x = cur->x; nxt->y = cur->y; nxt->z = cur->z;
        std::map<std::pair<int, int>, int> mid;
        auto midpoint = [&](int a, int b) {
            const std::pair<int, int> key(std::min(a, b), std::max(a, b));
            auto it = mid.find(key);
            if (it != mid.end()) return it->second;
            const int id = (int)nxt->x.size();
            nxt->x.push_back(0.5f * (cur->x[key.first] + cur->x[key.second]));
            nxt->y.push_back(0.5f * (cur->y[key.first] + cur->y[key.second]));
            nxt->z.push_back(0.5f * (cur->z[key.first] + cur->z[key.second]));
            mid.emplace(key, id);
            return id;
        };
        for (size_t t = 0; t < cur->v1.size(); t++) {
            const int a = cur->v1[t], b = cur->v2[t], c = cur->v3[t];
            const int ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
            const int tri[4][3] = {{a, ab, ca}, {ab, b, bc}, {ca, bc, c}, {ab, bc, ca}};
            for (auto &q : tri) {
                nxt->v1.push_back(q[0]); nxt->v2.push_back(q[1]); nxt->v3.push_back(q[2]);
            }
        }
        delete cur;
        cur = nxt;
    }
    *out = cur;
    return RC_OK;
}

// ---- .env / .query (placement of the bodies, start and goal) -----------------------------------
// The benchmark's environment file (models/alpha1.0/alpha.env) names the bodies and places the
// passive ones in the world: `obstacle.obj -c(1 1 0 1) 11.0899 9.00906 3.24427  0 0 0` (:13).  The
// reference never reads it -- its pipeline checks the two OBJ files as they are -- so nothing in
// this library applies a placement unless the caller asks for it with rc_mesh_place.
//
//   Boundary Box [lo:hi ; lo:hi ; lo:hi]
//   Multibodies / <count>
//   Active  / <count> / <file> [-c(r g b a)] <base type words> / Connections / <count>
//   Passive / <file> [-c(r g b a)] x y z  rx ry rz
// Tokens are read in order; a body line is any token ending in .obj or .g.

namespace {

bool ends_with(const std::string &s, const char *suffix)
{
    const size_t n = strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

bool parse_float(const std::string &tok, float *out)
{
    if (tok.empty()) return false;
    char *end = nullptr;
    const float v = strtof(tok.c_str(), &end);
    if (end == tok.c_str() || *end != '\0') return false;
    *out = v;
    return true;
}

}  // namespace

int rc_env_read(const char *path, rc_env *out)
{
    if (!path || !out) return RC_ERR_INVALID;
    std::ifstream in(path);
    if (!in.is_open()) return RC_ERR_IO;
    memset(out, 0, sizeof *out);
    for (int k = 0; k < 3; k++) { out->box_lo[k] = -FLT_MAX; out->box_hi[k] = FLT_MAX; }
    std::string line;
    int active = -1;  // section: 1 active, 0 passive, -1 none yet
    while (std::getline(in, line)) {
        // punctuation of the header and of the colour option becomes white space
        std::string flat = line;
        for (char &c : flat)
            if (c == '[' || c == ']' || c == ':' || c == ';' || c == '(' || c == ')' || c == ',') c = ' ';
        const std::vector<std::string> tok = tokens_of(flat);
        if (tok.empty() || tok[0][0] == '#') continue;
        if (tok[0] == "Boundary") {
            std::vector<float> v;
            for (const std::string &t : tok) { float f; if (parse_float(t, &f)) v.push_back(f); }
            if (v.size() != 6) return RC_ERR_FORMAT;
            for (int k = 0; k < 3; k++) { out->box_lo[k] = v[2 * k]; out->box_hi[k] = v[2 * k + 1]; }
            out->has_box = 1;
            continue;
        }
        if (tok[0] == "Active") { active = 1; continue; }
        if (tok[0] == "Passive") { active = 0; continue; }
        if (!(ends_with(tok[0], ".obj") || ends_with(tok[0], ".g"))) continue;  // counts, "Multibodies", "Connections"
        if (active < 0) return RC_ERR_FORMAT;
        if (out->n_bodies >= RC_ENV_MAX_BODIES) return RC_ERR_UNSUPPORTED;
        rc_env_body &b = out->bodies[out->n_bodies];
        if (tok[0].size() >= sizeof b.file) return RC_ERR_FORMAT;
        strcpy(b.file, tok[0].c_str());
        b.active = active;
        b.color[0] = b.color[1] = b.color[2] = 0.f; b.color[3] = 1.f;
        size_t k = 1;
        if (k < tok.size() && tok[k] == "-c") {
            for (int q = 0; q < 4; q++)
                if (k + 1 + q >= tok.size() || !parse_float(tok[k + 1 + q], &b.color[q])) return RC_ERR_FORMAT;
            k += 5;
        }
        // six numbers = placement (passive bodies; an active body carries its base type instead)
        std::vector<float> v;
        for (; k < tok.size(); k++) { float f; if (parse_float(tok[k], &f)) v.push_back(f); else break; }
        if (v.size() >= 6) {
            b.has_placement = 1;
            b.x = v[0]; b.y = v[1]; b.z = v[2]; b.rx = v[3]; b.ry = v[4]; b.rz = v[5];
        } else if (!active && !v.empty()) {
            return RC_ERR_FORMAT;  // a passive body with a truncated placement
        }
        out->n_bodies++;
    }
    return out->n_bodies > 0 ? RC_OK : RC_ERR_FORMAT;
}

// Lines of `<robot index> x y z r1 r2 r3` (models/alpha1.0/alpha.query: start, then goal).  The
// rotations are fractions of a full turn about x, y, z, i.e. roll, pitch, yaw of Configuration
// (Utils_rai.h:51-58) after scaling by 2 pi and wrapping into [-pi, pi).
int rc_query_read(const char *path, rc_config **confs, long *n)
{
    if (!path || !confs || !n) return RC_ERR_INVALID;
    std::ifstream in(path);
    if (!in.is_open()) return RC_ERR_IO;
    std::vector<rc_config> v;
    std::string line;
    const double two_pi = 6.283185307179586;
    while (std::getline(in, line)) {
        const std::vector<std::string> tok = tokens_of(line);
        if (tok.empty() || tok[0][0] == '#') continue;
        float f[7];
        if (tok.size() < 7) return RC_ERR_FORMAT;
        for (int k = 0; k < 7; k++)
            if (!parse_float(tok[k], &f[k])) return RC_ERR_FORMAT;
        auto angle = [&](float turns) {
            double a = std::fmod((double)turns, 1.0) * two_pi;
            if (a >= 3.141592653589793) a -= two_pi;
            if (a < -3.141592653589793) a += two_pi;
            return (float)a;
        };
        rc_config c;
        c.x = f[1]; c.y = f[2]; c.z = f[3];
        c.roll = angle(f[4]); c.pitch = angle(f[5]); c.yaw = angle(f[6]);
        v.push_back(c);
    }
    *n = (long)v.size();
    *confs = (rc_config *)malloc(sizeof(rc_config) * std::max<size_t>(v.size(), 1));
    if (!*confs) return RC_ERR_NOMEM;
    if (!v.empty()) memcpy(*confs, v.data(), sizeof(rc_config) * v.size());
    return RC_OK;
}

// The body moved to its place in the world: p' = Rz(rz) Ry(ry) Rx(rx) p + (x, y, z), angles in
// degrees, the matrix the reference builds for a configuration (broad-phase-fused.cu:26-55, A = yaw
// about z, B = pitch about y, C = roll about x), evaluated in double and rounded once per coordinate.
int rc_mesh_place(const rc_mesh *m, float x, float y, float z, float rx_deg, float ry_deg, float rz_deg,
                  rc_mesh **out)
{
    if (!m || !out) return RC_ERR_INVALID;
    const double d2r = 3.141592653589793 / 180.0;
    const double cA = std::cos(rz_deg * d2r), sA = std::sin(rz_deg * d2r);
    const double cB = std::cos(ry_deg * d2r), sB = std::sin(ry_deg * d2r);
    const double cC = std::cos(rx_deg * d2r), sC = std::sin(rx_deg * d2r);
    const double R[3][3] = {{cA * cB, cA * sB * sC - sA * cC, cA * sB * cC + sA * sC},
                            {sA * cB, sA * sB * sC + cA * cC, sA * sB * cC - cA * sC},
                            {-sB, cB * sC, cB * cC}};
    rc_mesh *r = new rc_mesh(*m);
    for (size_t v = 0; v < m->x.size(); v++) {
        const double px = m->x[v], py = m->y[v], pz = m->z[v];
        r->x[v] = (float)(R[0][0] * px + R[0][1] * py + R[0][2] * pz + x);
        r->y[v] = (float)(R[1][0] * px + R[1][1] * py + R[1][2] * pz + y);
        r->z[v] = (float)(R[2][0] * px + R[2][1] * py + R[2][2] * pz + z);
    }
    *out = r;
    return RC_OK;
}

// ---- .conf ---------------------------------------------------------------------------------

int rc_conf_read(const char *path, rc_config **confs, uint8_t **tags, long *n)
{
    if (!path || !confs || !n) return RC_ERR_INVALID;
    std::ifstream in(path);
    if (!in.is_open()) return RC_ERR_IO;
    std::string header;
    std::getline(in, header);  // "There are V valid configurations and I invalid configurations"
    std::vector<rc_config> v;
    std::vector<uint8_t> t;
    rc_config c;
    // operator>>(float) is what the reference uses (Utils_rai.cu:90): strtof semantics
    while (in >> c.x >> c.y >> c.z >> c.pitch >> c.yaw >> c.roll) {
        std::string tag;
        in >> tag;  // validity column (may be absent on the last line)
        v.push_back(c);
        t.push_back((uint8_t)(!tag.empty() && tag != "0"));
    }
    *n = (long)v.size();
    *confs = (rc_config *)malloc(sizeof(rc_config) * std::max<size_t>(v.size(), 1));
    if (!*confs) return RC_ERR_NOMEM;
    if (!v.empty()) memcpy(*confs, v.data(), sizeof(rc_config) * v.size());
    if (tags) {
        *tags = (uint8_t *)malloc(std::max<size_t>(t.size(), 1));
        if (!*tags) { free(*confs); *confs = nullptr; return RC_ERR_NOMEM; }
        if (!t.empty()) memcpy(*tags, t.data(), t.size());
    }
    return RC_OK;
}

int rc_conf_write(const char *path, const rc_config *confs, const uint8_t *valid, long n)
{
    if (!path || n < 0 || (n && !confs)) return RC_ERR_INVALID;
    std::ofstream out(path);
    if (!out.is_open()) return RC_ERR_IO;
    long nv = 0;
    for (long i = 0; i < n; i++) nv += (valid && valid[i]) ? 1 : 0;
    out << "There are " << nv << " valid configurations and " << (n - nv)
        << " invalid configurations\n";
    // default ostream formatting (6 significant digits), as the reference writes them
    for (long i = 0; i < n; i++)
        out << confs[i].x << " " << confs[i].y << " " << confs[i].z << " " << confs[i].pitch << " "
            << confs[i].yaw << " " << confs[i].roll << " " << (int)((valid && valid[i]) ? 1 : 0)
            << "\n";
    return out.good() ? RC_OK : RC_ERR_IO;
}

void rc_free(void *p) { free(p); }

int rc_conf_generate(rc_config *out, long n, float x_min, float x_max, float y_min, float y_max,
                     float z_min, float z_max, uint32_t seed)
{
    if (n < 0 || (n && !out)) return RC_ERR_INVALID;
    std::mt19937 gen(seed);
    std::uniform_real_distribution<float> dis_x(x_min, x_max), dis_y(y_min, y_max),
        dis_z(z_min, z_max), dis_rot(-M_PI, M_PI);
    for (long i = 0; i < n; i++) {
        rc_config c;
        c.x = dis_x(gen); c.y = dis_y(gen); c.z = dis_z(gen);
        c.pitch = dis_rot(gen); c.yaw = dis_rot(gen); c.roll = dis_rot(gen);
        out[i] = c;
    }
    return RC_OK;
}

int rc_alpha_box(int hard, float box[6])
{
    if (!box) return RC_ERR_INVALID;
    // extents of alpha1.0/robot.obj as hard-coded by the reference (Utils_rai.cu:113-118)
    const float x_min = 3.72119f, y_min = -11.0518f, z_min = -0.608012f;
    const float x_max = 65.9453f, y_max = 26.0984f, z_max = 18.6984f;
    const float r[3] = {x_max - x_min, y_max - y_min, z_max - z_min};
    for (int k = 0; k < 3; k++) {
        box[2 * k] = hard ? -r[k] / 200 : -r[k] * 10;
        box[2 * k + 1] = hard ? r[k] / 200 : r[k] * 10;
    }
    return RC_OK;
}

int rc_conf_generate_alpha(rc_config *out, long n, int hard, uint32_t seed)
{
    float b[6];
    rc_alpha_box(hard, b);
    return rc_conf_generate(out, n, b[0], b[1], b[2], b[3], b[4], b[5], seed);
}

}  // extern "C"
