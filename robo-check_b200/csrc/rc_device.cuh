// rc_device.cuh -- FP32 geometry shared by every kernel of the hot path.
//
// Every rounding step is pinned with __fmul_rn/__fadd_rn/__fsub_rn/__fmaf_rn/__fdiv_rn
// (ptxas never contracts or reassociates those), in the order the reference's own kernels
// execute them when compiled for sm_100a -- that order is decided by ptxas, not by the
// C++ text, and was read off the SASS of
//   broadPhaseFusedKernel_sep  (reference broad-phase/broad-phase-fused.cu:157-207)
//   narrowPhaseKernel_coarse   (reference narrow-phase/narrow-phase.cu:551-688)
// See DESIGN.md "Arithmetic ground truth".  Result: collision flags and transformed
// vertices are bit-identical to the reference pipeline, not merely close.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rc {

// narrow-phase.hu:14  `#define TOL 1e-6` is a DOUBLE.  The reference compares
//   (double)x >= TOL     <=>  x >= 0x358637BE   (first float above 1e-6)
//   (double)|x| <  TOL   <=>  |x| <= 0x358637BD (1e-6f, the float just below 1e-6)
//   x <= (float)(-TOL)   <=>  x <= -0x358637BD  (no_overlap_sep rounds -TOL to float)
#define RC_TOL_GE __int_as_float(0x358637BE)
#define RC_TOL_F  __int_as_float(0x358637BD)

struct Mat34 {  // rows of [R | t]; the 4th matrix row (0,0,0,1) is never used
    float r00, r01, r02, tx, r10, r11, r12, ty, r20, r21, r22, tz;
};

// createTransformationMatrix, broad-phase-fused.cu:26-55 (A = yaw, B = pitch, C = roll).
// sinf/cosf are the same libdevice routines the reference inlines.
__device__ __forceinline__ Mat34 matrix_from_trig(float x, float y, float z, float sA, float cA, float sB,
                                                  float cB, float sC, float cC)
{
    const float cAsB = __fmul_rn(cA, sB);
    const float sAsB = __fmul_rn(sA, sB);
    Mat34 m;
    m.r00 = __fmul_rn(cA, cB);
    m.r01 = __fmaf_rn(sC, cAsB, -__fmul_rn(sA, cC));
    m.r02 = __fmaf_rn(cC, cAsB, __fmul_rn(sA, sC));
    m.tx = x;
    m.r10 = __fmul_rn(sA, cB);
    m.r11 = __fmaf_rn(cA, cC, __fmul_rn(sC, sAsB));
    m.r12 = __fmaf_rn(cC, sAsB, -__fmul_rn(cA, sC));
    m.ty = y;
    m.r20 = -sB;
    m.r21 = __fmul_rn(cB, sC);
    m.r22 = __fmul_rn(cB, cC);
    m.tz = z;
    return m;
}

__device__ __forceinline__ Mat34 make_matrix(float x, float y, float z, float pitch, float yaw,
                                             float roll)
{
    const float cB = cosf(pitch), sB = sinf(pitch);
    const float cA = cosf(yaw), sA = sinf(yaw);
    const float cC = cosf(roll), sC = sinf(roll);
    return matrix_from_trig(x, y, z, sA, cA, sB, cB, sC, cC);
}

// The same matrix for a whole warp working on ONE configuration: lanes 0, 1, 2 evaluate the sine and
// cosine of yaw, pitch, roll (the other lanes repeat roll), six shuffles hand them round -- a third
// of the trigonometry of make_matrix in every lane, same values bit for bit.
__device__ __forceinline__ Mat34 make_matrix_warp(float x, float y, float z, float pitch, float yaw,
                                                  float roll, int lane)
{
    const float ang = lane == 0 ? yaw : lane == 1 ? pitch : roll;
    const float s = sinf(ang), c = cosf(ang);
    const float sA = __shfl_sync(0xffffffffu, s, 0), cA = __shfl_sync(0xffffffffu, c, 0);
    const float sB = __shfl_sync(0xffffffffu, s, 1), cB = __shfl_sync(0xffffffffu, c, 1);
    const float sC = __shfl_sync(0xffffffffu, s, 2), cC = __shfl_sync(0xffffffffu, c, 2);
    return matrix_from_trig(x, y, z, sA, cA, sB, cB, sC, cC);
}

// transformVector(Vector3f, Matrix4f), broad-phase-fused.cu:84-105: each row accumulates
// from +0 and the translation column is added LAST.
__device__ __forceinline__ void xform(const Mat34 &m, float x, float y, float z, float &ox,
                                      float &oy, float &oz)
{
    ox = __fadd_rn(m.tx, __fmaf_rn(m.r02, z, __fmaf_rn(m.r01, y, __fmaf_rn(m.r00, x, 0.0f))));
    oy = __fadd_rn(m.ty, __fmaf_rn(m.r12, z, __fmaf_rn(m.r11, y, __fmaf_rn(m.r10, x, 0.0f))));
    oz = __fadd_rn(m.tz, __fmaf_rn(m.r22, z, __fmaf_rn(m.r21, y, __fmaf_rn(m.r20, x, 0.0f))));
}

// dimensionCollides x3 and the flag, broad-phase-fused.cu:108-111,195-206.  1 = valid.
__device__ __forceinline__ bool broad_valid(float rxmin, float rymin, float rzmin, float rxmax,
                                            float rymax, float rzmax, float oxmin, float oymin,
                                            float ozmin, float oxmax, float oymax, float ozmax)
{
    const bool hit = (oxmin <= rxmax && rxmin <= oxmax) && (oymin <= rymax && rymin <= oymax) &&
                     (ozmin <= rzmax && rzmin <= ozmax);
    return !hit;
}

struct Tri {
    float ax, ay, az, bx, by, bz, cx, cy, cz;
};
struct Plane {  // N and s = N.p1; the reference's d is -s (compute_plane_sep, :62-76)
    float nx, ny, nz, s;
};

__device__ __forceinline__ Plane plane_of(const Tri &t)
{
    const float e1x = __fsub_rn(t.bx, t.ax), e1y = __fsub_rn(t.by, t.ay), e1z = __fsub_rn(t.bz, t.az);
    const float e2x = __fsub_rn(t.cx, t.bx), e2y = __fsub_rn(t.cy, t.by), e2z = __fsub_rn(t.cz, t.bz);
    Plane p;
    p.nx = __fmaf_rn(e1y, e2z, -__fmul_rn(e1z, e2y));
    p.ny = __fmaf_rn(e1z, e2x, -__fmul_rn(e1x, e2z));
    p.nz = __fmaf_rn(e1x, e2y, -__fmul_rn(e1y, e2x));
    p.s = __fmaf_rn(p.nz, t.az, __fmaf_rn(p.nx, t.ax, __fmul_rn(p.ny, t.ay)));
    return p;
}

// compute_signed_dists_sep, :87-91: (Ny*py) rounded, then two FMAs, then + d (= - s).
__device__ __forceinline__ float sdist(const Plane &p, float x, float y, float z)
{
    return __fsub_rn(__fmaf_rn(p.nz, z, __fmaf_rn(p.nx, x, __fmul_rn(p.ny, y))), p.s);
}

// !no_overlap_sep, :100-106: some distance >= TOL (double compare) and some <= (float)-TOL.
__device__ __forceinline__ bool straddles(float d0, float d1, float d2)
{
    const float mx = fmaxf(fmaxf(d0, d1), d2);
    const float mn = fminf(fminf(d0, d1), d2);
    return (mx >= RC_TOL_GE) && (mn <= -RC_TOL_F);
}

__device__ __forceinline__ bool close0(float v) { return fabsf(v) <= RC_TOL_F; }

// isclose, :24-35
__device__ __forceinline__ bool isclose(float v1, float v2)
{
    const bool z1 = close0(v1), z2 = close0(v2);
    if (z1 && z2) return true;
    if (z1 || z2) return false;
    return close0(__fdiv_rn(__fsub_rn(v1, v2), v1));
}

// is_coplanar_sep, :297-338
__device__ __forceinline__ bool is_coplanar(const Plane &a, const Plane &b)
{
    const float p1[4] = {a.nx, a.ny, a.nz, -a.s};
    const float p2[4] = {b.nx, b.ny, b.nz, -b.s};
    float ratio = 0.0f;
    bool started = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool z1 = close0(p1[i]), z2 = close0(p2[i]);
        if (z1 != z2) return false;
        if (z1) continue;
        const float q = __fdiv_rn(p1[i], p2[i]);
        if (!started) {
            ratio = q;
            started = true;
        } else if (!isclose(ratio, q)) {
            return false;
        }
    }
    return true;
}

// la_solve, :109-125, with the kernel's placement: (A1*b2) rounded, A3*b1 fused into the
// subtraction; the denominator is the cross-product component the caller already has.
__device__ __forceinline__ void la_solve(float A1, float A2, float A3, float A4, float b1, float b2,
                                         float den, float &x1, float &x2)
{
    if (close0(A1)) {
        x2 = __fdiv_rn(b1, A2);
        x1 = __fdiv_rn(__fmaf_rn(A4, -x2, b2), A3);
    } else {
        const float num = __fmaf_rn(b1, A3, -__fmul_rn(A1, b2));
        x2 = __fdiv_rn(num, den);
        x1 = __fdiv_rn(__fmaf_rn(-A2, x2, b1), A1);
    }
}

// project_vertex_sep, :188-190
__device__ __forceinline__ float project(float vx, float vy, float vz, float Dx, float Dy, float Dz,
                                         float Ox, float Oy, float Oz)
{
    return __fmaf_rn(Dz, __fsub_rn(vz, Oz),
                     __fmaf_rn(Dx, __fsub_rn(vx, Ox), __fmul_rn(Dy, __fsub_rn(vy, Oy))));
}

// canonicalize_triangle_sep, :220-236 -> which vertex goes first / middle / last
__device__ __forceinline__ void canonical_order(float d0, float d1, float d2, int &i0, int &i1,
                                                int &i2)
{
    if ((d0 > 0.f && d1 > 0.f) || (d0 < 0.f && d1 < 0.f)) {
        i0 = 0; i1 = 2; i2 = 1;
    } else if ((d0 > 0.f && d2 > 0.f) || (d0 < 0.f && d2 < 0.f)) {
        i0 = 0; i1 = 1; i2 = 2;
    } else {
        i0 = 1; i1 = 0; i2 = 2;
    }
}

// The two t-values of one triangle on the intersection line
// (compute_parametric_variable_sep x2, :246-251, narrow-phase.cu:653-656).
__device__ __forceinline__ void tri_interval(const Tri &t, float d0, float d1, float d2, float Dx,
                                             float Dy, float Dz, float Ox, float Oy, float Oz,
                                             float &t01, float &t12)
{
    int i0, i1, i2;
    canonical_order(d0, d1, d2, i0, i1, i2);
    const float p[3] = {project(t.ax, t.ay, t.az, Dx, Dy, Dz, Ox, Oy, Oz),
                        project(t.bx, t.by, t.bz, Dx, Dy, Dz, Ox, Oy, Oz),
                        project(t.cx, t.cy, t.cz, Dx, Dy, Dz, Ox, Oy, Oz)};
    const float d[3] = {d0, d1, d2};
    // select without dynamic register indexing
    const float pa = i0 == 0 ? p[0] : p[1];
    const float da = i0 == 0 ? d[0] : d[1];
    const float pb = i1 == 0 ? p[0] : (i1 == 1 ? p[1] : p[2]);
    const float db = i1 == 0 ? d[0] : (i1 == 1 ? d[1] : d[2]);
    const float pc = i2 == 1 ? p[1] : p[2];
    const float dc = i2 == 1 ? d[1] : d[2];
    t01 = __fadd_rn(pa, __fdiv_rn(__fmul_rn(__fsub_rn(pb, pa), da), __fsub_rn(da, db)));
    t12 = __fadd_rn(pb, __fdiv_rn(__fmul_rn(__fsub_rn(pc, pb), db), __fsub_rn(db, dc)));
}

// Everything after the two plane tests of narrowPhaseKernel_coarse (:639-670):
// coplanarity skip, intersection line, canonical order, parametric intervals, interval test.
// ---- optional decision of coplanar pairs ----------------------------------------
// The reference's kernel only prints "Error: require coplanar intersection" for a pair whose planes
// is_coplanar_sep calls proportional and moves on (narrow-phase.cu:647-652, :680); its Python design
// model decides such pairs in the common plane (python/narrow.py:98-177: edge-edge intersections,
// then containment).  With coplanar mode 1 the pair is decided here the same way -- both triangles
// projected along the dominant axis of the robot triangle's normal, the nine edge pairs tested for
// intersection (orientation signs, touching counts), then containment of a vertex of one triangle in
// the other (done exactly; the model's box comparison is only a sketch).  Every operation is rounded
// individually so that the CPU checker reproduces the decision bit for bit.  Default mode 0 = the
// reference's behaviour: flags are unchanged.
__device__ __forceinline__ float orient2(float ax, float ay, float bx, float by, float cx, float cy)
{
    return __fsub_rn(__fmul_rn(__fsub_rn(bx, ax), __fsub_rn(cy, ay)),
                     __fmul_rn(__fsub_rn(by, ay), __fsub_rn(cx, ax)));
}
__device__ __forceinline__ bool on_seg2(float ax, float ay, float bx, float by, float px, float py)
{
    return fminf(ax, bx) <= px && px <= fmaxf(ax, bx) && fminf(ay, by) <= py && py <= fmaxf(ay, by);
}
__device__ __forceinline__ bool seg_meet2(float p1x, float p1y, float p2x, float p2y, float p3x, float p3y,
                                          float p4x, float p4y)
{
    const float d1 = orient2(p3x, p3y, p4x, p4y, p1x, p1y), d2 = orient2(p3x, p3y, p4x, p4y, p2x, p2y);
    const float d3 = orient2(p1x, p1y, p2x, p2y, p3x, p3y), d4 = orient2(p1x, p1y, p2x, p2y, p4x, p4y);
    if (((d1 > 0.f && d2 < 0.f) || (d1 < 0.f && d2 > 0.f)) && ((d3 > 0.f && d4 < 0.f) || (d3 < 0.f && d4 > 0.f)))
        return true;
    if (d1 == 0.f && on_seg2(p3x, p3y, p4x, p4y, p1x, p1y)) return true;
    if (d2 == 0.f && on_seg2(p3x, p3y, p4x, p4y, p2x, p2y)) return true;
    if (d3 == 0.f && on_seg2(p1x, p1y, p2x, p2y, p3x, p3y)) return true;
    if (d4 == 0.f && on_seg2(p1x, p1y, p2x, p2y, p4x, p4y)) return true;
    return false;
}
__device__ __forceinline__ bool inside2(const float t[6], float px, float py)
{
    const float o1 = orient2(t[0], t[1], t[2], t[3], px, py), o2 = orient2(t[2], t[3], t[4], t[5], px, py),
                o3 = orient2(t[4], t[5], t[0], t[1], px, py);
    return (o1 >= 0.f && o2 >= 0.f && o3 >= 0.f) || (o1 <= 0.f && o2 <= 0.f && o3 <= 0.f);
}
__device__ __noinline__ bool coplanar_overlap(const Tri r, const Tri o, const Plane pr)
{
    const float ax = fabsf(pr.nx), ay = fabsf(pr.ny), az = fabsf(pr.nz);
    const int drop = (az >= ax && az >= ay) ? 2 : (ay >= ax ? 1 : 0);
    float a[6], b[6];
    if (drop == 2) {
        a[0] = r.ax; a[1] = r.ay; a[2] = r.bx; a[3] = r.by; a[4] = r.cx; a[5] = r.cy;
        b[0] = o.ax; b[1] = o.ay; b[2] = o.bx; b[3] = o.by; b[4] = o.cx; b[5] = o.cy;
    } else if (drop == 1) {
        a[0] = r.ax; a[1] = r.az; a[2] = r.bx; a[3] = r.bz; a[4] = r.cx; a[5] = r.cz;
        b[0] = o.ax; b[1] = o.az; b[2] = o.bx; b[3] = o.bz; b[4] = o.cx; b[5] = o.cz;
    } else {
        a[0] = r.ay; a[1] = r.az; a[2] = r.by; a[3] = r.bz; a[4] = r.cy; a[5] = r.cz;
        b[0] = o.ay; b[1] = o.az; b[2] = o.by; b[3] = o.bz; b[4] = o.cy; b[5] = o.cz;
    }
    // rare path: rolled loops over arrays in local memory keep the register footprint (which every
    // caller of pair_tail pays for around the call) small
#pragma unroll 1
    for (int i = 0; i < 3; i++) {
        const int i2 = (i + 1) % 3;
#pragma unroll 1
        for (int k = 0; k < 3; k++) {
            const int k2 = (k + 1) % 3;
            if (seg_meet2(a[2 * i], a[2 * i + 1], a[2 * i2], a[2 * i2 + 1], b[2 * k], b[2 * k + 1], b[2 * k2],
                          b[2 * k2 + 1]))
                return true;
        }
    }
    return inside2(b, a[0], a[1]) || inside2(a, b[0], b[1]);
}

// Verdict of pair_tail.  kPairCoplanar: is_coplanar_sep calls the two planes proportional; the
// reference leaves such a pair undecided (no collision, narrow-phase.cu:647-652), and so does every
// caller here unless the context is in RC_COPLANAR_DECIDE mode, in which case the caller runs
// coplanar_overlap on the pair.  The kernels carry the mode as a template parameter (COP): the
// descent kernel is register-bound (64 registers at 1,024 threads), any code added to its default
// instantiation costs spills, so the default build contains none of this.
constexpr int kPairMiss = 0, kPairHit = 1, kPairCoplanar = 2;

// DBG: the intermediate values go to dbg[0..9] = D, O, t_r01, t_r12, t_o01, t_o12 (device known-answer
// tests, rc_kat.cuh); the production instantiation is pair_tail below.
template <bool DBG>
__device__ __forceinline__ int pair_tail_core(const Tri &r, const Plane &pr, float dR0, float dR1, float dR2,
                                              const Tri &o, const Plane &po, float dO0, float dO1, float dO2,
                                              float *dbg)
{
    if (is_coplanar(pr, po)) return kPairCoplanar;

    // compute_intersect_line_sep(N1 = Nr, d1 = dr, N2 = No, d2 = do): Dx, Dz are two rounded
    // products and a subtraction (shared with la_solve's A2A3 - A1A4); Dy is fused.
    const float yz = __fmul_rn(pr.ny, po.nz), zy = __fmul_rn(pr.nz, po.ny);
    const float xy = __fmul_rn(pr.nx, po.ny), yx = __fmul_rn(pr.ny, po.nx);
    const float Dx = __fsub_rn(yz, zy);
    const float Dy = __fmaf_rn(pr.nz, po.nx, -__fmul_rn(pr.nx, po.nz));
    const float Dz = __fsub_rn(xy, yx);
    float Ox, Oy, Oz, x1, x2;
    // b1 = -d1 = pr.s, b2 = -d2 = po.s
    if (!close0(Dz)) {
        la_solve(pr.nx, pr.ny, po.nx, po.ny, pr.s, po.s, __fsub_rn(yx, xy), x1, x2);
        Ox = x1; Oy = x2; Oz = 0.f;
    } else if (!close0(Dy)) {
        la_solve(pr.nx, pr.nz, po.nx, po.nz, pr.s, po.s, Dy, x1, x2);
        Ox = x1; Oy = 0.f; Oz = x2;
    } else {
        la_solve(pr.ny, pr.nz, po.ny, po.nz, pr.s, po.s, __fsub_rn(zy, yz), x1, x2);
        Ox = 0.f; Oy = x1; Oz = x2;
    }
    float tr01, tr12, to01, to12;
    tri_interval(r, dR0, dR1, dR2, Dx, Dy, Dz, Ox, Oy, Oz, tr01, tr12);
    tri_interval(o, dO0, dO1, dO2, Dx, Dy, Dz, Ox, Oy, Oz, to01, to12);
    if (DBG) {
        dbg[0] = Dx; dbg[1] = Dy; dbg[2] = Dz; dbg[3] = Ox; dbg[4] = Oy; dbg[5] = Oz;
        dbg[6] = tr01; dbg[7] = tr12; dbg[8] = to01; dbg[9] = to12;
    }
    // :659-670 -- `>=`; a NaN makes both comparisons false, i.e. a collision, as in the reference
    if (fminf(tr01, tr12) >= fmaxf(to01, to12)) return kPairMiss;
    if (fminf(to01, to12) >= fmaxf(tr01, tr12)) return kPairMiss;
    return kPairHit;
}

__device__ __noinline__ int pair_tail(const Tri r, const Plane pr, float dR0, float dR1, float dR2,
                                      const Tri o, const Plane po, float dO0, float dO1, float dO2)
{
    return pair_tail_core<false>(r, pr, dR0, dR1, dR2, o, po, dO0, dO1, dO2, nullptr);
}

// One (robot triangle, obstacle triangle) decision given both planes (the obstacle plane is
// configuration independent, so the context precomputes it with the same arithmetic).
template <bool COP>
__device__ __forceinline__ bool pair_hits(const Tri &r, const Plane &pr, const Tri &o, const Plane &po)
{
    const float dO0 = sdist(pr, o.ax, o.ay, o.az), dO1 = sdist(pr, o.bx, o.by, o.bz),
                dO2 = sdist(pr, o.cx, o.cy, o.cz);
    if (!straddles(dO0, dO1, dO2)) return false;
    const float dR0 = sdist(po, r.ax, r.ay, r.az), dR1 = sdist(po, r.bx, r.by, r.bz),
                dR2 = sdist(po, r.cx, r.cy, r.cz);
    if (!straddles(dR0, dR1, dR2)) return false;
    const int verdict = pair_tail(r, pr, dR0, dR1, dR2, o, po, dO0, dO1, dO2);
    if constexpr (COP) {
        if (verdict == kPairCoplanar) return coplanar_overlap(r, o, pr);
    }
    return verdict == kPairHit;
}

// ---- small PTX wrappers: mbarrier + TMA bulk copy (global -> shared) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// cp.async.bulk: bytes must be a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace rc

// ---- sphere-approximated robot (BASELINE config 3) ---------------------------------------------
// Squared distance from point p to triangle (a, b, c): closest-point regions of Ericson,
// "Real-Time Collision Detection" 5.1.5.  Every operation is individually rounded (no FMA), dot
// products accumulate left to right, so a plain C restatement reproduces it bit for bit.
namespace rc {
__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

__device__ __forceinline__ float point_tri_dist2(float px, float py, float pz, const Tri &t)
{
    const float abx = __fsub_rn(t.bx, t.ax), aby = __fsub_rn(t.by, t.ay), abz = __fsub_rn(t.bz, t.az);
    const float acx = __fsub_rn(t.cx, t.ax), acy = __fsub_rn(t.cy, t.ay), acz = __fsub_rn(t.cz, t.az);
    const float apx = __fsub_rn(px, t.ax), apy = __fsub_rn(py, t.ay), apz = __fsub_rn(pz, t.az);
    const float d1 = dot3_rn(abx, aby, abz, apx, apy, apz), d2 = dot3_rn(acx, acy, acz, apx, apy, apz);
    float qx, qy, qz;  // closest point
    const float bpx = __fsub_rn(px, t.bx), bpy = __fsub_rn(py, t.by), bpz = __fsub_rn(pz, t.bz);
    const float d3 = dot3_rn(abx, aby, abz, bpx, bpy, bpz), d4 = dot3_rn(acx, acy, acz, bpx, bpy, bpz);
    const float cpx = __fsub_rn(px, t.cx), cpy = __fsub_rn(py, t.cy), cpz = __fsub_rn(pz, t.cz);
    const float d5 = dot3_rn(abx, aby, abz, cpx, cpy, cpz), d6 = dot3_rn(acx, acy, acz, cpx, cpy, cpz);
    const float vc = __fsub_rn(__fmul_rn(d1, d4), __fmul_rn(d3, d2));
    const float vb = __fsub_rn(__fmul_rn(d5, d2), __fmul_rn(d1, d6));
    const float va = __fsub_rn(__fmul_rn(d3, d6), __fmul_rn(d5, d4));
    if (d1 <= 0.f && d2 <= 0.f) {
        qx = t.ax; qy = t.ay; qz = t.az;
    } else if (d3 >= 0.f && d4 <= d3) {
        qx = t.bx; qy = t.by; qz = t.bz;
    } else if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        const float v = __fdiv_rn(d1, __fsub_rn(d1, d3));
        qx = __fadd_rn(t.ax, __fmul_rn(v, abx)); qy = __fadd_rn(t.ay, __fmul_rn(v, aby));
        qz = __fadd_rn(t.az, __fmul_rn(v, abz));
    } else if (d6 >= 0.f && d5 <= d6) {
        qx = t.cx; qy = t.cy; qz = t.cz;
    } else if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        const float w = __fdiv_rn(d2, __fsub_rn(d2, d6));
        qx = __fadd_rn(t.ax, __fmul_rn(w, acx)); qy = __fadd_rn(t.ay, __fmul_rn(w, acy));
        qz = __fadd_rn(t.az, __fmul_rn(w, acz));
    } else if (va <= 0.f && __fsub_rn(d4, d3) >= 0.f && __fsub_rn(d5, d6) >= 0.f) {
        const float e = __fsub_rn(d4, d3);
        const float w = __fdiv_rn(e, __fadd_rn(e, __fsub_rn(d5, d6)));
        qx = __fadd_rn(t.bx, __fmul_rn(w, __fsub_rn(t.cx, t.bx)));
        qy = __fadd_rn(t.by, __fmul_rn(w, __fsub_rn(t.cy, t.by)));
        qz = __fadd_rn(t.bz, __fmul_rn(w, __fsub_rn(t.cz, t.bz)));
    } else {
        const float denom = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(va, vb), vc));
        const float v = __fmul_rn(vb, denom), w = __fmul_rn(vc, denom);
        qx = __fadd_rn(__fadd_rn(t.ax, __fmul_rn(abx, v)), __fmul_rn(acx, w));
        qy = __fadd_rn(__fadd_rn(t.ay, __fmul_rn(aby, v)), __fmul_rn(acy, w));
        qz = __fadd_rn(__fadd_rn(t.az, __fmul_rn(abz, v)), __fmul_rn(acz, w));
    }
    const float ex = __fsub_rn(px, qx), ey = __fsub_rn(py, qy), ez = __fsub_rn(pz, qz);
    return dot3_rn(ex, ey, ez, ex, ey, ez);
}
}  // namespace rc
