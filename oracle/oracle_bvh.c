/* oracle_bvh.c -- CPU checker with bounding-volume trees (TEST INFRASTRUCTURE ONLY).
 *
 * Same per-pair arithmetic and the same flag as rco_check (oracle.c), but the pairs are
 * enumerated by a dual-tree descent instead of the exhaustive sweep, so whole BASELINE-size sets
 * (100,000 / 1,000,000 configurations) can be checked on the host cores in seconds.  Two roles:
 *   - full-size parity checker for the GPU path (tests/), itself pinned against rco_check on
 *     every set the exhaustive oracle can finish (tests/test_oracle.py);
 *   - the same-box CPU comparator SURVEY 8(d) asks for next to the reference's brute force:
 *     "a self-written multithreaded BVH CPU checker using the identical per-pair arithmetic",
 *     standing in for FCL's BVH path (Utils.cu:116-147, full-integration-test.cu:519-554:
 *     fcl::BVHModel<OBBRSS> + fcl::collide per configuration), which cannot be installed here.
 *
 * Nothing here is shared with the product (robo-check_b200/host/rc_bvh.cpp is a different
 * build: SAH splits, other leaf sizes, float arithmetic); the only common idea is which pairs
 * may be skipped: a (robot node, obstacle node) pair is dropped when the robot node's box
 * (an oriented box in the world) and the obstacle node's box are separated along one of the
 * three world axes or the three robot axes, boxes inflated by RCO_BVH_MARGIN.  That no such pair
 * is ever reported by the reference arithmetic is the empirical statement of SURVEY 7 (0 of
 * 86,410 reported pairs had disjoint boxes); tests/test_oracle.py re-checks it against the
 * exhaustive sweep.
 *
 * Reference lines followed for the arithmetic: see oracle.c (transform broad-phase-fused.cu:
 * 26-55,84-105; AABB + broad phase :108-111,186-200; pair decision narrow-phase.cu:619-670).   */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define RCO_BVH_MARGIN 1e-3
#define LEAF_TRIS 4

typedef struct {
    double c[3], h[3];   /* box centre, half extent (inflated)        */
    int left, right;     /* children, or -1                           */
    int first, count;    /* leaf: range in the triangle permutation   */
} node_t;

typedef struct {
    node_t *nodes;
    int n_nodes, cap;
    int *perm;           /* triangle order                             */
} tree_t;

static void tri_box(const rco_mesh *m, int t, double lo[3], double hi[3])
{
    const int v[3] = {m->v1[t], m->v2[t], m->v3[t]};
    for (int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; }
    for (int k = 0; k < 3; k++) {
        const double p[3] = {m->x[v[k]], m->y[v[k]], m->z[v[k]]};
        for (int a = 0; a < 3; a++) {
            if (p[a] < lo[a]) lo[a] = p[a];
            if (p[a] > hi[a]) hi[a] = p[a];
        }
    }
}

static const double *g_cent;
static int g_axis;
static int cmp_cent(const void *pa, const void *pb)
{
    const double a = g_cent[3 * *(const int *)pa + g_axis], b = g_cent[3 * *(const int *)pb + g_axis];
    return (a > b) - (a < b);
}

static int build_node(tree_t *T, const rco_mesh *m, const double *cent, int first, int count)
{
    if (T->n_nodes == T->cap) {
        T->cap = T->cap ? 2 * T->cap : 256;
        T->nodes = (node_t *)realloc(T->nodes, sizeof(node_t) * (size_t)T->cap);
    }
    const int id = T->n_nodes++;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    double clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = first; i < first + count; i++) {
        double l[3], h[3];
        tri_box(m, T->perm[i], l, h);
        for (int a = 0; a < 3; a++) {
            if (l[a] < lo[a]) lo[a] = l[a];
            if (h[a] > hi[a]) hi[a] = h[a];
            const double c = cent[3 * T->perm[i] + a];
            if (c < clo[a]) clo[a] = c;
            if (c > chi[a]) chi[a] = c;
        }
    }
    node_t nd;
    for (int a = 0; a < 3; a++) {
        nd.c[a] = 0.5 * (lo[a] + hi[a]);
        nd.h[a] = 0.5 * (hi[a] - lo[a]) + RCO_BVH_MARGIN;
    }
    nd.left = nd.right = -1; nd.first = first; nd.count = count;
    if (count > LEAF_TRIS) {
        int axis = 0;
        for (int a = 1; a < 3; a++) if (chi[a] - clo[a] > chi[axis] - clo[axis]) axis = a;
        g_cent = cent; g_axis = axis;   /* build is single threaded */
        qsort(T->perm + first, (size_t)count, sizeof(int), cmp_cent);
        const int half = count / 2;
        T->nodes[id] = nd;
        const int l = build_node(T, m, cent, first, half);
        const int r = build_node(T, m, cent, first + half, count - half);
        nd.left = l; nd.right = r;
    }
    T->nodes[id] = nd;
    return id;
}

static void build_tree(tree_t *T, const rco_mesh *m)
{
    memset(T, 0, sizeof *T);
    const int nt = m->n_t;
    T->perm = (int *)malloc(sizeof(int) * (size_t)(nt > 0 ? nt : 1));
    double *cent = (double *)malloc(sizeof(double) * 3 * (size_t)(nt > 0 ? nt : 1));
    for (int t = 0; t < nt; t++) {
        double lo[3], hi[3];
        T->perm[t] = t;
        tri_box(m, t, lo, hi);
        for (int a = 0; a < 3; a++) cent[3 * t + a] = 0.5 * (lo[a] + hi[a]);
    }
    if (nt > 0) build_node(T, m, cent, 0, nt);
    free(cent);
}

static void free_tree(tree_t *T) { free(T->nodes); free(T->perm); }

/* robot node (box in the robot frame, placed by R | t) against obstacle node (world box):
 * 1 = possibly touching, 0 = separated along a world axis or a robot axis                 */
static inline int boxes_meet(const node_t *r, const node_t *o, const double R[3][3],
                             const double aR[3][3], const double t[3])
{
    double d[3];
    for (int i = 0; i < 3; i++)
        d[i] = R[i][0] * r->c[0] + R[i][1] * r->c[1] + R[i][2] * r->c[2] + t[i] - o->c[i];
    for (int i = 0; i < 3; i++)
        if (fabs(d[i]) > o->h[i] + aR[i][0] * r->h[0] + aR[i][1] * r->h[1] + aR[i][2] * r->h[2]) return 0;
    for (int k = 0; k < 3; k++) {
        const double p = d[0] * R[0][k] + d[1] * R[1][k] + d[2] * R[2][k];
        if (fabs(p) > r->h[k] + aR[0][k] * o->h[0] + aR[1][k] * o->h[1] + aR[2][k] * o->h[2]) return 0;
    }
    return 1;
}

typedef struct {
    const rco_mesh *rob, *obs;
    const tree_t *tr, *to;
    const float *obs_tri;      /* [n_obs_t][9] gathered obstacle triangles */
    const rco_config *confs;
    uint8_t *valid;
    rco_aabb_t obs_box;
    int variant;
    long n;
    long *next;                /* shared work counter (chunks of configurations) */
    rco_stats stats;
} bjob_t;

#define CHUNK 32

static void *bvh_worker(void *arg)
{
    bjob_t *J = (bjob_t *)arg;
    const rco_mesh *Rm = J->rob;
    const int nv = Rm->n_v;
    float *wx = (float *)malloc(sizeof(float) * 3 * (size_t)(nv > 0 ? nv : 1));
    float *wy = wx + nv, *wz = wy + nv;
    int cap = 1024, (*stack)[2] = malloc(sizeof(int[2]) * (size_t)cap);
    rco_stats st;
    memset(&st, 0, sizeof st);
    for (;;) {
        const long begin = __atomic_fetch_add(J->next, CHUNK, __ATOMIC_RELAXED);
        if (begin >= J->n) break;
        const long end = begin + CHUNK < J->n ? begin + CHUNK : J->n;
        for (long i = begin; i < end; i++) {
            float m[12];
            rco_aabb_t box;
            rco_matrix(&J->confs[i], m);
            rco_transform(&J->confs[i], 1, Rm->x, Rm->y, Rm->z, nv, wx, wy, wz);
            rco_aabb(wx, wy, wz, nv, &box);
            if (rco_broad_valid(&box, &J->obs_box)) {
                J->valid[i] = 1;
                st.broad_cleared++;
                continue;
            }
            double R[3][3], aR[3][3], t[3];
            for (int r = 0; r < 3; r++) {
                for (int c = 0; c < 3; c++) { R[r][c] = m[4 * r + c]; aR[r][c] = fabs(R[r][c]); }
                t[r] = m[4 * r + 3];
            }
            int sp = 0, hit = 0;
            if (J->tr->n_nodes && J->to->n_nodes) { stack[0][0] = 0; stack[0][1] = 0; sp = 1; }
            while (sp && !hit) {
                sp--;
                const int a = stack[sp][0], b = stack[sp][1];
                const node_t *ra = &J->tr->nodes[a], *ob = &J->to->nodes[b];
                st.pairs_stage1++;   /* here: box pairs tested */
                if (!boxes_meet(ra, ob, R, aR, t)) continue;
                const int rleaf = ra->left < 0, oleaf = ob->left < 0;
                if (rleaf && oleaf) {
                    for (int p = ra->first; p < ra->first + ra->count && !hit; p++) {
                        const int j = J->tr->perm[p];
                        const int v[3] = {Rm->v1[j], Rm->v2[j], Rm->v3[j]};
                        const float r9[9] = {wx[v[0]], wy[v[0]], wz[v[0]], wx[v[1]], wy[v[1]], wz[v[1]],
                                             wx[v[2]], wy[v[2]], wz[v[2]]};
                        for (int q = ob->first; q < ob->first + ob->count; q++) {
                            int cop = 0;
                            st.pairs_full++;   /* here: triangle pairs decided */
                            if (rco_pair_hits(r9, J->obs_tri + 9 * (size_t)J->to->perm[q], J->variant, &cop)) {
                                hit = 1;
                                st.hit_pairs++;
                                break;
                            }
                            st.coplanar += cop;
                        }
                    }
                    continue;
                }
                if (sp + 2 > cap) { cap *= 2; stack = realloc(stack, sizeof(int[2]) * (size_t)cap); }
                /* split the node with the larger box (or the only inner one) */
                const double sr = ra->h[0] + ra->h[1] + ra->h[2], so = ob->h[0] + ob->h[1] + ob->h[2];
                if (oleaf || (!rleaf && sr >= so)) {
                    stack[sp][0] = ra->left; stack[sp][1] = b; sp++;
                    stack[sp][0] = ra->right; stack[sp][1] = b; sp++;
                } else {
                    stack[sp][0] = a; stack[sp][1] = ob->left; sp++;
                    stack[sp][0] = a; stack[sp][1] = ob->right; sp++;
                }
            }
            J->valid[i] = (uint8_t)!hit;
            if (hit) st.narrow_invalid++; else st.narrow_valid++;
        }
    }
    free(wx);
    free(stack);
    J->stats = st;
    return NULL;
}

int rco_check_bvh(const rco_mesh *rob, const rco_mesh *obs, const rco_config *confs, long n,
                  int variant, int nthreads, uint8_t *valid, rco_stats *stats)
{
    if (!rob || !obs || !confs || !valid || n < 0) return -1;
    const int base = variant & ~RCO_FLAG_COPLANAR;
    if (base != RCO_VARIANT_GPU && base != RCO_VARIANT_HOST) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    tree_t tr, to;
    build_tree(&tr, rob);
    build_tree(&to, obs);
    float *tri = (float *)malloc(sizeof(float) * 9 * (size_t)(obs->n_t > 0 ? obs->n_t : 1));
    for (int k = 0; k < obs->n_t; k++) {
        const int idx[3] = {obs->v1[k], obs->v2[k], obs->v3[k]};
        for (int v = 0; v < 3; v++) {
            tri[9 * k + 3 * v + 0] = obs->x[idx[v]];
            tri[9 * k + 3 * v + 1] = obs->y[idx[v]];
            tri[9 * k + 3 * v + 2] = obs->z[idx[v]];
        }
    }
    rco_aabb_t obox;
    rco_aabb(obs->x, obs->y, obs->z, obs->n_v, &obox);
    long next = 0;
    bjob_t *jobs = (bjob_t *)calloc((size_t)nthreads, sizeof(bjob_t));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        jobs[t].rob = rob; jobs[t].obs = obs; jobs[t].tr = &tr; jobs[t].to = &to;
        jobs[t].obs_tri = tri; jobs[t].confs = confs; jobs[t].valid = valid; jobs[t].obs_box = obox;
        jobs[t].variant = variant; jobs[t].n = n; jobs[t].next = &next;
    }
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, bvh_worker, &jobs[t]);
    bvh_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int t = 0; t < nthreads; t++) {
            stats->broad_cleared += jobs[t].stats.broad_cleared;
            stats->narrow_valid += jobs[t].stats.narrow_valid;
            stats->narrow_invalid += jobs[t].stats.narrow_invalid;
            stats->pairs_stage1 += jobs[t].stats.pairs_stage1;
            stats->pairs_full += jobs[t].stats.pairs_full;
            stats->hit_pairs += jobs[t].stats.hit_pairs;
            stats->coplanar += jobs[t].stats.coplanar;
        }
    }
    free(jobs); free(th); free(tri);
    free_tree(&tr); free_tree(&to);
    return 0;
}
