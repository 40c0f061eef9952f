// mccba_f32x2.cuh -- the per-corner arithmetic of the residual / Jacobian pass in packed single precision: every lane
// evaluates TWO corners at once in Blackwell's f32x2 instructions (FFMA2 / FMUL2 / FADD2: one issue slot, two FMAs).
//
// Why (profiles/r2_pipe_peaks.txt, measured on the box): DFMA 63 lane-instr/clk/SM, FFMA2 63 lane-instr/clk/SM -- but an
// FFMA2 does two FMAs, so the same arithmetic costs half the pipe time, and the FP64 pipe (152 instructions per corner
// in the fp64 kernel) was the bound of the whole pass.  The reference itself evaluates the projection through float32
// (src/multicalib.cpp:742-749, 789-792: composed pose, projected points and residual are CV_32F), so float32 per-corner
// arithmetic is inside the reference's own noise; what is summed over many corners (the 28 per-edge sums) is promoted
// to double once per edge, and everything downstream (Schur complement, reduced solve, update, cost test) stays fp64.
//
// The same source compiles for the host (plain float pairs, fmaf) so that tests/harness can run the packed arithmetic
// without a GPU; the two agree except for the reciprocal seed (MUFU + one Newton step on the device, 1.0f / x on the host).
#pragma once
#include <math.h>

#include "mccba_math.cuh"

namespace mccba {

#if defined(__CUDA_ARCH__)
struct f2 {
    unsigned long long v;
};
MC_HD f2 f2_make(float lo, float hi)
{
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
MC_HD float f2_lo(f2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); (void)hi; return lo; }
MC_HD float f2_hi(f2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); (void)lo; return hi; }
MC_HD f2 operator+(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
MC_HD f2 operator-(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
MC_HD f2 operator*(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
MC_HD f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
// negation as "unpack, negate both halves, repack": ptxas folds exactly this form into the operand modifier of the consuming
// FFMA2 / FMUL2 (-R.F32x2.HI_LO); the 64-bit XOR of the sign bits it does not (two LOP3 per negation, 16 per step)
MC_HD f2 f2_neg(f2 a) { return f2_make(-f2_lo(a), -f2_hi(a)); }
MC_HD f2 f2_rcp(f2 a)   // positive, normal arguments (depths, distortion denominators); ~1 ulp
{
    float lo = f2_lo(a), hi = f2_hi(a), ylo, yhi;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ylo) : "f"(lo));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(yhi) : "f"(hi));
    const f2 y = f2_make(ylo, yhi);
    const f2 e = f2_fma(f2_neg(a), y, f2_make(1.0f, 1.0f));
    return f2_fma(y, e, y);
}
MC_HD f2 f2_rsqrt(f2 a)
{
    float lo = f2_lo(a), hi = f2_hi(a), ylo, yhi;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ylo) : "f"(lo));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yhi) : "f"(hi));
    const f2 y = f2_make(ylo, yhi);
    const f2 h = f2_make(0.5f, 0.5f) * a;
    return y * f2_fma(f2_neg(h) * y, y, f2_make(1.5f, 1.5f));
}
#else
struct f2 {
    float x, y;
};
MC_HD f2 f2_make(float lo, float hi) { f2 r; r.x = lo; r.y = hi; return r; }
MC_HD float f2_lo(f2 a) { return a.x; }
MC_HD float f2_hi(f2 a) { return a.y; }
MC_HD f2 operator+(f2 a, f2 b) { return f2_make(a.x + b.x, a.y + b.y); }
MC_HD f2 operator-(f2 a, f2 b) { return f2_make(a.x - b.x, a.y - b.y); }
MC_HD f2 operator*(f2 a, f2 b) { return f2_make(a.x * b.x, a.y * b.y); }
MC_HD f2 f2_fma(f2 a, f2 b, f2 c) { return f2_make(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
MC_HD f2 f2_neg(f2 a) { return f2_make(-a.x, -a.y); }
MC_HD f2 f2_rcp(f2 a) { return f2_make(1.0f / a.x, 1.0f / a.y); }
MC_HD f2 f2_rsqrt(f2 a) { return f2_make(1.0f / sqrtf(a.x), 1.0f / sqrtf(a.y)); }
#endif
MC_HD f2 f2_dup(float a) { return f2_make(a, a); }

// Intrinsics of one camera in single precision, with the constant factors the Jacobian needs folded in once.  Plain floats:
// a value duplicated into both halves at its use (f2_dup: mov.b64 {x, x}) is folded by ptxas into the scalar-broadcast
// operand form of the packed instruction (FFMA2 R, R.F32x2.HI_LO, R.F32, ...), so a constant costs one register and 4 bytes
// of shared memory instead of a register pair and 8 bytes.
struct alignas(16) CamF2 {
    float fx, fy, skew, xi;
    float cx, cy;
    float k1, k2, k3, k4, k5, k6, p1, p2;
    float k1_2, k2_4, k3_6;      // 2 k1, 4 k2, 6 k3 (d rad / d r2, doubled)
    float k4_1, k5_2, k6_3;      // k4, 2 k5, 3 k6
    float p1_2, p2_2, p1_6, p2_6;
    int model, rational;
};
MC_HD CamF2 make_cam_f2(const CamParams& c)
{
    CamF2 r;
    r.fx = (float)c.fx; r.fy = (float)c.fy; r.skew = (float)c.skew; r.xi = (float)c.xi;
    r.cx = (float)c.cx; r.cy = (float)c.cy;
    r.k1 = (float)c.k1; r.k2 = (float)c.k2; r.k3 = (float)c.k3;
    r.k4 = (float)c.k4; r.k5 = (float)c.k5; r.k6 = (float)c.k6;
    r.p1 = (float)c.p1; r.p2 = (float)c.p2;
    r.k1_2 = (float)(2.0 * c.k1); r.k2_4 = (float)(4.0 * c.k2); r.k3_6 = (float)(6.0 * c.k3);
    r.k4_1 = (float)c.k4; r.k5_2 = (float)(2.0 * c.k5); r.k6_3 = (float)(3.0 * c.k6);
    r.p1_2 = (float)(2.0 * c.p1); r.p2_2 = (float)(2.0 * c.p2);
    r.p1_6 = (float)(6.0 * c.p1); r.p2_6 = (float)(6.0 * c.p2);
    r.model = c.model; r.rational = c.rational;
    return r;
}

// Residual e = observed - projected and A = d(u,v)/dXc (2 x 3) for the corner pair; same formulas as pinhole_point /
// omnidir_point in mccba_math.cuh (cv::projectPoints per-point arithmetic; src/omnidir.cpp:146-165, 185-199).
// kNeedE == false: the caller has the residual already (MIXED policy) and the distorted point is not evaluated.
template <bool kRational, bool kNeedE>
MC_HD void pinhole_pair(const CamF2& c, const f2* Xc, f2 iu, f2 iv, f2* e, f2* A)
{
    const f2 one = f2_dup(1.0f);
    const f2 iz = f2_rcp(Xc[2]);
    const f2 x = Xc[0] * iz, y = Xc[1] * iz;
    const f2 xx = x * x, xy = x * y, yy = y * y;
    const f2 r2 = xx + yy;
    f2 rad = f2_fma(r2, f2_fma(r2, f2_fma(r2, f2_dup(c.k3), f2_dup(c.k2)), f2_dup(c.k1)), one);
    f2 dd2 = f2_fma(r2, f2_fma(r2, f2_dup(c.k3_6), f2_dup(c.k2_4)), f2_dup(c.k1_2));          // 2 d rad / d r2
    if (kRational) {
        const f2 den = f2_fma(r2, f2_fma(r2, f2_fma(r2, f2_dup(c.k6), f2_dup(c.k5)), f2_dup(c.k4)), one);
        const f2 dden2 = f2_fma(r2, f2_fma(r2, f2_dup(c.k6_3), f2_dup(c.k5_2)), f2_dup(c.k4_1));
        const f2 iden = f2_rcp(den);
        rad = rad * iden;
        dd2 = (dd2 - (rad + rad) * dden2) * iden;                      // 2 (drad - rad dden) / den
    }
    if (kNeedE) {
        const f2 xy2 = xy + xy;
        const f2 xd = f2_fma(x, rad, f2_fma(f2_dup(c.p1), xy2, f2_dup(c.p2) * f2_fma(xx, f2_dup(2.0f), r2)));
        const f2 yd = f2_fma(y, rad, f2_fma(f2_dup(c.p2), xy2, f2_dup(c.p1) * f2_fma(yy, f2_dup(2.0f), r2)));
        e[0] = f2_fma(f2_neg(f2_dup(c.fx)), xd, iu - f2_dup(c.cx));
        e[1] = f2_fma(f2_neg(f2_dup(c.fy)), yd, iv - f2_dup(c.cy));
    }
    const f2 t = f2_fma(f2_dup(c.p1_2), x, f2_dup(c.p2_2) * y);
    // (the tangential terms are summed on their own and added last: chaining them onto rad costs one more rounding at
    // magnitude 1 per entry, and the double-sided problem of tests/test_double_side.py then leaves the 1e-6 gate)
    const f2 dxdx = f2_fma(dd2, xx, rad) + f2_fma(f2_dup(c.p1_2), y, f2_dup(c.p2_6) * x);
    const f2 dxdy = f2_fma(dd2, xy, t);
    const f2 dydy = f2_fma(dd2, yy, rad) + f2_fma(f2_dup(c.p1_6), y, f2_dup(c.p2_2) * x);
    const f2 fxz = f2_dup(c.fx) * iz, fyz = f2_dup(c.fy) * iz;
    const f2 nx = f2_neg(x), ny = f2_neg(y);
    A[0] = fxz * dxdx; A[1] = fxz * dxdy; A[2] = f2_fma(A[0], nx, A[1] * ny);
    A[3] = fyz * dxdy; A[4] = fyz * dydy; A[5] = f2_fma(A[3], nx, A[4] * ny);
}

template <bool kNeedE>
MC_HD void omnidir_pair(const CamF2& c, const f2* Xc, f2 iu, f2 iv, f2* e, f2* A)
{
    const f2 one = f2_dup(1.0f);
    const f2 n2 = f2_fma(Xc[0], Xc[0], f2_fma(Xc[1], Xc[1], Xc[2] * Xc[2]));
    const f2 rn = f2_rsqrt(n2);
    const f2 s0 = Xc[0] * rn, s1 = Xc[1] * rn, s2 = Xc[2] * rn;
    const f2 id = f2_rcp(s2 + f2_dup(c.xi));
    const f2 x = s0 * id, y = s1 * id;
    const f2 xx = x * x, xy = x * y, yy = y * y;
    const f2 r2 = xx + yy;
    const f2 rad = f2_fma(r2, f2_fma(r2, f2_dup(c.k2), f2_dup(c.k1)), one);
    const f2 dd2 = f2_fma(r2, f2_dup(c.k2_4), f2_dup(c.k1_2));
    if (kNeedE) {
        const f2 xy2 = xy + xy;
        const f2 xd = f2_fma(x, rad, f2_fma(f2_dup(c.p1), xy2, f2_dup(c.p2) * f2_fma(xx, f2_dup(2.0f), r2)));
        const f2 yd = f2_fma(y, rad, f2_fma(f2_dup(c.p2), xy2, f2_dup(c.p1) * f2_fma(yy, f2_dup(2.0f), r2)));
        e[0] = f2_fma(f2_neg(f2_dup(c.fx)), xd, f2_fma(f2_neg(f2_dup(c.skew)), yd, iu - f2_dup(c.cx)));
        e[1] = f2_fma(f2_neg(f2_dup(c.fy)), yd, iv - f2_dup(c.cy));
    }
    const f2 t = f2_fma(f2_dup(c.p1_2), x, f2_dup(c.p2_2) * y);
    const f2 dxdx = f2_fma(dd2, xx, rad) + f2_fma(f2_dup(c.p1_2), y, f2_dup(c.p2_6) * x);
    const f2 dxdy = f2_fma(dd2, xy, t);
    const f2 dydy = f2_fma(dd2, yy, rad) + f2_fma(f2_dup(c.p1_6), y, f2_dup(c.p2_2) * x);
    const f2 m00 = f2_fma(f2_dup(c.fx), dxdx, f2_dup(c.skew) * dxdy), m01 = f2_fma(f2_dup(c.fx), dxdy, f2_dup(c.skew) * dydy);
    const f2 m10 = f2_dup(c.fy) * dxdy, m11 = f2_dup(c.fy) * dydy;
    const f2 k = rn * id;
    const f2 nx = f2_neg(x), ny = f2_neg(y);
    {
        const f2 g2 = f2_fma(m00, nx, m01 * ny);
        const f2 d = f2_fma(m00, s0, f2_fma(m01, s1, g2 * s2));
        const f2 nd = f2_neg(d);
        A[0] = k * f2_fma(nd, s0, m00); A[1] = k * f2_fma(nd, s1, m01); A[2] = k * f2_fma(nd, s2, g2);
    }
    {
        const f2 g2 = f2_fma(m10, nx, m11 * ny);
        const f2 d = f2_fma(m10, s0, f2_fma(m11, s1, g2 * s2));
        const f2 nd = f2_neg(d);
        A[3] = k * f2_fma(nd, s0, m10); A[4] = k * f2_fma(nd, s1, m11); A[5] = k * f2_fma(nd, s2, g2);
    }
}

// Residual of one corner in double (projection only, no Jacobian): what the MIXED policy feeds into the packed pass.
// A float32 projection of ~1000 px coordinates carries ~1e-5 px of rounding noise; harmless for the cost, but the tilt of
// a board that faces a camera almost squarely reacts to anisotropic noise like noise / tilt, and 1e-5 px then moves
// such a pose by more than the 1e-6 parity gate.  Evaluated in double and rounded ONCE (2^-24 |e| ~ 1e-8 px) it does not.
// The arithmetic is that of pinhole_point / omnidir_point with the products regrouped so that every operation is one FMA
// (translation folded into the rotation chain, tangential terms through the shared x^2, xy, y^2 and doubled constants):
// 37.5 FP64 instructions per pinhole corner instead of 41.5.  kPlanar: the object points have z = 0 (a flat board).
struct alignas(16) CamResid {      // the intrinsics the double residual reads
    double fx, fy, cx, cy, skew, xi;
    double k1, k2, k3, k4, k5, k6;
    double p1, p2, p1_2, p2_2;
};
MC_HD CamResid make_cam_resid(const CamParams& c)
{
    CamResid r;
    r.fx = c.fx; r.fy = c.fy; r.cx = c.cx; r.cy = c.cy; r.skew = c.skew; r.xi = c.xi;
    r.k1 = c.k1; r.k2 = c.k2; r.k3 = c.k3; r.k4 = c.k4; r.k5 = c.k5; r.k6 = c.k6;
    r.p1 = c.p1; r.p2 = c.p2; r.p1_2 = 2.0 * c.p1; r.p2_2 = 2.0 * c.p2;
    return r;
}
template <int kModel, bool kRational, bool kPlanar>
MC_HD void corner_residual(const CamResid& c, const double* R3, const double* T3, float ox, float oy, float oz, float iu, float iv,
                           double* e)
{
    const double X0 = (double)ox, X1 = (double)oy;
    double Xc[3];
    if (kPlanar) {
#pragma unroll
        for (int i = 0; i < 3; ++i) Xc[i] = fma(R3[3 * i], X0, fma(R3[3 * i + 1], X1, T3[i]));
    } else {
        const double X2 = (double)oz;
#pragma unroll
        for (int i = 0; i < 3; ++i) Xc[i] = fma(R3[3 * i], X0, fma(R3[3 * i + 1], X1, fma(R3[3 * i + 2], X2, T3[i])));
    }
    double x, y;
    if (kModel == kPinhole) {
        const double iz = proj_rcp(Xc[2]);
        x = Xc[0] * iz; y = Xc[1] * iz;
    } else {
        const double n2 = fma(Xc[0], Xc[0], fma(Xc[1], Xc[1], Xc[2] * Xc[2]));
        const double rn = proj_rsqrt(n2);
        const double id = proj_rcp(fma(Xc[2], rn, c.xi));
        const double k = rn * id;
        x = Xc[0] * k; y = Xc[1] * k;
    }
    const double xx = x * x, xy = x * y, yy = y * y;
    const double r2 = xx + yy;
    double rad;
    if (kModel == kPinhole) {
        rad = fma(r2, fma(r2, fma(r2, c.k3, c.k2), c.k1), 1.0);
        if (kRational) rad = rad * proj_rcp(fma(r2, fma(r2, fma(r2, c.k6, c.k5), c.k4), 1.0));
    } else {
        rad = fma(r2, fma(r2, c.k2, c.k1), 1.0);
    }
    // xd = x rad + 2 p1 xy + p2 (r2 + 2 x^2),  yd = y rad + p1 (r2 + 2 y^2) + 2 p2 xy
    const double xd = fma(x, rad, fma(c.p1_2, xy, fma(c.p2, r2, c.p2_2 * xx)));
    const double yd = fma(y, rad, fma(c.p2_2, xy, fma(c.p1, r2, c.p1_2 * yy)));
    if (kModel == kPinhole) e[0] = fma(-c.fx, xd, (double)iu - c.cx);
    else e[0] = fma(-c.fx, xd, fma(-c.skew, yd, (double)iu - c.cx));
    e[1] = fma(-c.fy, yd, (double)iv - c.cy);
}

// Two corners of one edge: Xc = R3 X + T3, residual, 2 x 6 Jacobian wrt the left perturbation of the composed pose,
// accumulated into acc[28] (pairs; the two halves are added at the end of the edge) in the layout of corner_accumulate:
// upper triangle of sum J^T J (21) | sum J^T e (6) | sum |e|^2.  w = 1 for a live corner, 0 for layout padding.
// kExactE: the residual pair (ex0, ex1) was evaluated in double by the caller (corner_residual) and replaces the float one.
template <int kModel, bool kRational, bool kExactE, bool kPlanar = false>
MC_HD void corner_pair_accumulate(const CamF2& c, const float* R3, const float* T3, f2 ox, f2 oy, f2 oz, f2 iu, f2 iv, f2 w, bool masked,
                                  f2* acc, f2 ex0, f2 ex1)
{
    f2 Q[3], Xc[3], e[2], A[6];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (kPlanar) Q[i] = f2_fma(f2_dup(R3[3 * i]), ox, f2_dup(R3[3 * i + 1]) * oy);      // flat board: oz = 0
        else Q[i] = f2_fma(f2_dup(R3[3 * i]), ox, f2_fma(f2_dup(R3[3 * i + 1]), oy, f2_dup(R3[3 * i + 2]) * oz));
        Xc[i] = Q[i] + f2_dup(T3[i]);
    }
    if (kModel == kPinhole) pinhole_pair<kRational, !kExactE>(c, Xc, iu, iv, e, A);
    else omnidir_pair<!kExactE>(c, Xc, iu, iv, e, A);
    if (kExactE) { e[0] = ex0; e[1] = ex1; }
    if (masked) {
        e[0] = e[0] * w; e[1] = e[1] * w;
#pragma unroll
        for (int i = 0; i < 6; ++i) A[i] = A[i] * w;
    }
    f2 j0[6], j1[6];
    const f2 nq0 = f2_neg(Q[0]), nq1 = f2_neg(Q[1]), nq2 = f2_neg(Q[2]);
    j0[0] = f2_fma(Q[1], A[2], nq2 * A[1]); j0[1] = f2_fma(Q[2], A[0], nq0 * A[2]); j0[2] = f2_fma(Q[0], A[1], nq1 * A[0]);
    j1[0] = f2_fma(Q[1], A[5], nq2 * A[4]); j1[1] = f2_fma(Q[2], A[3], nq0 * A[5]); j1[2] = f2_fma(Q[0], A[4], nq1 * A[3]);
    j0[3] = A[0]; j0[4] = A[1]; j0[5] = A[2];
    j1[3] = A[3]; j1[4] = A[4]; j1[5] = A[5];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = i; j < 6; ++j) acc[tri6(i, j)] = f2_fma(j1[i], j1[j], f2_fma(j0[i], j0[j], acc[tri6(i, j)]));
        acc[21 + i] = f2_fma(j1[i], e[1], f2_fma(j0[i], e[0], acc[21 + i]));
    }
    if (!kExactE) acc[27] = f2_fma(e[1], e[1], f2_fma(e[0], e[0], acc[27]));   // MIXED: the caller sums the cost in double
}

}  // namespace mccba
