/* oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE ONLY, see oracle.c). */
#ifndef RCO_ORACLE_H
#define RCO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RCO_VARIANT_GPU 0  /* broadPhaseFusedKernel_sep + narrowPhaseKernel_coarse */
#define RCO_VARIANT_HOST 1 /* narrowPhaseBaseline (reference CPU brute force)       */
/* or-ed into a variant: pairs the reference leaves undecided for coplanarity
 * (narrow-phase.cu:647-652) are decided in the common plane (python/narrow.py:98-177);
 * mirrors RC_COPLANAR_DECIDE of the product.  Not reference behaviour: parity of this
 * mode is pinned only against the product's own definition.                          */
#define RCO_FLAG_COPLANAR 0x10

typedef struct { float x, y, z, pitch, yaw, roll; } rco_config;            /* Utils_rai.h:51-58 */
typedef struct { float x_min, y_min, z_min, x_max, y_max, z_max; } rco_aabb_t; /* Utils_rai.h:87-95 */

typedef struct {
    const float *x, *y, *z;   /* SoA vertices, n_v each   */
    const int *v1, *v2, *v3;  /* SoA triangles, n_t each  */
    int n_v, n_t;
} rco_mesh;

typedef struct {
    long broad_cleared, narrow_valid, narrow_invalid;
    long pairs_stage1, pairs_full, hit_pairs, coplanar;
} rco_stats;

void rco_sincosf(float x, float *sn, float *cs);
void rco_matrix(const rco_config *c, float m[12]);
void rco_transform(const rco_config *confs, int n, const float *vx, const float *vy,
                   const float *vz, int nv, float *ox, float *oy, float *oz);
void rco_aabb(const float *x, const float *y, const float *z, int nv, rco_aabb_t *out);
int rco_broad_valid(const rco_aabb_t *rob, const rco_aabb_t *obs);

int rco_isclose(float v1, float v2);
void rco_plane(const float p[9], int fused, float N[3], float *d);
void rco_signed_dists(const float N[3], float d, const float p[9], int fused, float out[3]);
int rco_no_overlap(const float dist[3], int sep);
void rco_la_solve(float A1, float A2, float A3, float A4, float b1, float b2, float *x1, float *x2);
void rco_intersect_line(const float N1[3], float d1, const float N2[3], float d2, int fused,
                        float D[3], float O[3]);
float rco_project(const float V[3], const float D[3], const float O[3], int fused);
void rco_canonicalize(const float d[3], int idx[3]);
float rco_param(const float v0[3], const float v1[3], float d0, float d1, const float D[3],
                const float O[3], int fused);
int rco_is_coplanar(const float N1[3], float d1, const float N2[3], float d2);
int rco_coplanar_overlap(const float r[9], const float o[9], const float Nr[3]);
int rco_pair_hits(const float r[9], const float o[9], int variant, int *coplanar);

int rco_check(const rco_mesh *rob, const rco_mesh *obs, const rco_config *confs, long n,
              int variant, int nthreads, int early_exit, uint8_t *valid, rco_stats *stats);
/* No early exit: counts[i] = pairs of configuration i the reference arithmetic reports. */
int rco_check_counts(const rco_mesh *rob, const rco_mesh *obs, const rco_config *confs, long n,
                     int variant, int nthreads, uint8_t *valid, uint32_t *counts);
/* la_solve with the FMA placement of the CUDA kernel (the GPU variant's; rco_la_solve is the host's) */
void rco_la_solve_gpu(float A1, float A2, float A3, float A4, float b1, float b2, float *x1, float *x2);
/* Same flags as rco_check (early exit), pairs enumerated by a dual-tree descent (oracle_bvh.c):
 * whole BASELINE-size sets in seconds.  In its statistics pairs_stage1 counts box pairs tested
 * and pairs_full triangle pairs decided. */
int rco_check_bvh(const rco_mesh *rob, const rco_mesh *obs, const rco_config *confs, long n,
                  int variant, int nthreads, uint8_t *valid, rco_stats *stats);
int rco_narrow(const rco_mesh *rob, const rco_mesh *obs, const float *px, const float *py,
               const float *pz, long n, int variant, uint8_t *valid);

float rco_point_tri_dist2(const float p[3], const float t[9]);
int rco_check_spheres(const float *spheres, int n_s, const rco_mesh *obs, const rco_config *confs,
                      long n, int nthreads, uint8_t *valid);

#ifdef __cplusplus
}
#endif
#endif
