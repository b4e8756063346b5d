// rc_kat.cuh -- device known-answer tests of the per-pair helpers in rc_device.cuh.
//
// The reference pins its helpers with host-side unit tests (narrow-phase/test-narrow-phase.cu:7-626:
// isclose, plane, signed distances, no_overlap, la_solve, intersection line, projection, canonical
// order, parametric values, coplanarity).  The kernels of this library use the SAME device
// functions as below -- plane_of, sdist, straddles, pair_tail_core (is_coplanar, la_solve,
// tri_interval, canonical_order, project) -- so running those inputs, and any number of real
// triangle pairs, through this kernel and comparing every intermediate value bit for bit with the
// CPU restatement pins the CUDA helpers themselves, not merely the flags they add up to.
#pragma once

namespace rc {

// One thread per (robot triangle, obstacle triangle) pair, 9 floats each (a, b, c).  out[32 * i ..]:
//   0-2 Nr  3 dr  4-6 distO   7-9 No  10 do  11-13 distR   14-16 D  17-19 O  20-23 t_r01 t_r12 t_o01 t_o12
// verdict[i] = straddleO | straddleR << 1 | tail << 2   (tail: 0 miss, 1 hit, 2 coplanar; the kernels
// evaluate it only when both plane tests pass, like narrowPhaseKernel_coarse, narrow-phase.cu:619-670)
__global__ void kat_pairs_kernel(const float *r9, const float *o9, int n, float *out, int *verdict)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = r9 + 9 * (size_t)i, *q = o9 + 9 * (size_t)i;
    const Tri r = {p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]};
    const Tri o = {q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8]};
    const Plane pr = plane_of(r), po = plane_of(o);
    const float dO0 = sdist(pr, o.ax, o.ay, o.az), dO1 = sdist(pr, o.bx, o.by, o.bz), dO2 = sdist(pr, o.cx, o.cy, o.cz);
    const float dR0 = sdist(po, r.ax, r.ay, r.az), dR1 = sdist(po, r.bx, r.by, r.bz), dR2 = sdist(po, r.cx, r.cy, r.cz);
    float *w = out + 32 * (size_t)i;
    for (int k = 0; k < 32; k++) w[k] = 0.f;
    w[0] = pr.nx; w[1] = pr.ny; w[2] = pr.nz; w[3] = -pr.s;
    w[4] = dO0; w[5] = dO1; w[6] = dO2;
    w[7] = po.nx; w[8] = po.ny; w[9] = po.nz; w[10] = -po.s;
    w[11] = dR0; w[12] = dR1; w[13] = dR2;
    const bool sO = straddles(dO0, dO1, dO2), sR = straddles(dR0, dR1, dR2);
    // the tail runs unconditionally here (the reference's unit tests exercise the intersection line and
    // the parametric values on pairs that do not pass both plane tests); the kernels only reach it when
    // sO && sR, and only then does the verdict's tail field mean anything
    const int tail = pair_tail_core<true>(r, pr, dR0, dR1, dR2, o, po, dO0, dO1, dO2, w + 14);
    verdict[i] = (sO ? 1 : 0) | (sR ? 2 : 0) | (tail << 2);
}

// la_solve (narrow-phase.cu:109-125) on in[6 * i ..] = A1 A2 A3 A4 b1 b2 -> out[2 * i ..] = x1 x2, with
// the denominator the kernels hand it: the rounded products' difference A2 A3 - A1 A4.
__global__ void kat_la_solve_kernel(const float *in, int n, float *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *v = in + 6 * (size_t)i;
    float x1, x2;
    la_solve(v[0], v[1], v[2], v[3], v[4], v[5], __fsub_rn(__fmul_rn(v[1], v[2]), __fmul_rn(v[0], v[3])), x1, x2);
    out[2 * i] = x1;
    out[2 * i + 1] = x2;
}

cudaError_t launch_kat_pairs(const float *d_r9, const float *d_o9, int n, float *d_out, int *d_verdict,
                             cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    kat_pairs_kernel<<<(n + 127) / 128, 128, 0, stream>>>(d_r9, d_o9, n, d_out, d_verdict);
    return cudaGetLastError();
}

cudaError_t launch_kat_la_solve(const float *d_in, int n, float *d_out, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    kat_la_solve_kernel<<<(n + 127) / 128, 128, 0, stream>>>(d_in, n, d_out);
    return cudaGetLastError();
}

}  // namespace rc
