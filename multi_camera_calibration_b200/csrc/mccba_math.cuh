// mccba_math.cuh -- per-thread arithmetic of the calibration bundle adjustment, written once as
// __host__ __device__ inline functions so that the CUDA kernels (mccba_kernels.cu) and the host-side unit test
// harness (tests/harness) exercise the very same code.
//
// Formulation (B200-first, NOT the reference's): the reference differentiates wrt additive Rodrigues vectors
// through compose_motion's 9x9 chain products (src/multicalib.cpp:1008-1056) and cv::projectPoints' dR/dom.  Here
// every rotation is perturbed on the left, R <- exp(psi) R, so that
//   * the per-corner Jacobian wrt the composed pose is  J = [ Q x a | a ]  (Q = R3 X, a = rows of d(u,v)/dXc),
//   * the chain matrices to camera / pattern-pose coordinates are rotations and cross products only,
//   * no log map is needed anywhere on the device.
// The normal equations are solved in these tangent coordinates; the step is mapped back to the reference's additive
// Rodrigues parametrisation with J_l(om)^-1 before the update, so iterates match the reference formulation to
// rounding (tests/test_parity_gpu.py, tests/test_math_host.py).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MC_HD __host__ __device__ __forceinline__
#else
#define MC_HD inline
#endif

namespace mccba {

// 1/x and 1/sqrt(x) for the per-corner projection (depths, distortion denominators: positive, far from the ends of the
// double range).  On the device: MUFU seed + Newton steps, ~1 ulp, no special-case branch -- the IEEE division and
// sqrt() carry a slow-path call that costs the residual kernel a divergence check per corner.  The host build of this
// header (tests/harness) keeps the library operations; the two agree to a few ulp.
MC_HD double proj_rcp(double d)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);   // seed error e0 -> e0^4 after two steps
    return fma(y, e, y);
#else
    return 1.0 / d;
#endif
}
MC_HD double proj_rsqrt(double d)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double hd = 0.5 * d;
    y = y * fma(-hd * y, y, 1.5);
    y = y * fma(-hd * y, y, 1.5);
    return y * fma(-hd * y, y, 1.5);
#else
    return 1.0 / sqrt(d);
#endif
}

constexpr int kPinhole = 0;
constexpr int kOmnidir = 1;
constexpr int kBlk = 28;  // per-edge block record: 21 (upper triangle of H6) + 6 (g6) + 1 (sum sq residual)

// Intrinsics of one camera, widened to double.  Mirrors _cameraMatrix/_distortCoeffs/_xi (multicalib.hpp:211-213).
struct alignas(16) CamParams {
    double fx, fy, cx, cy, skew, xi;
    double k1, k2, p1, p2, k3, k4, k5, k6;
    int model;     // kPinhole / kOmnidir
    int rational;  // pinhole only: any of k4,k5,k6 non-zero
};

MC_HD constexpr int tri6(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }  // i <= j

// ---------------------------------------------------------------------------------------------------------
// SO(3)
// ---------------------------------------------------------------------------------------------------------
MC_HD void cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// R = exp([om]x) = I + a K + b K^2, a = sin(th)/th, b = (1-cos th)/th^2   (cv::Rodrigues vec->mat)
MC_HD void rodrigues(const double* om, double* R)
{
    const double x = om[0], y = om[1], z = om[2];
    const double th2 = x * x + y * y + z * z;
    double a, b;
    if (th2 < 1e-8) {
        a = 1.0 - th2 * (1.0 / 6.0) + th2 * th2 * (1.0 / 120.0);
        b = 0.5 - th2 * (1.0 / 24.0) + th2 * th2 * (1.0 / 720.0);
    } else {
        const double th = sqrt(th2);
        double s, c;
#if defined(__CUDA_ARCH__)
        sincos(th, &s, &c);
        double sh = sin(0.5 * th);
#else
        s = sin(th); c = cos(th);
        double sh = sin(0.5 * th);
#endif
        (void)c;
        a = s / th;
        b = 2.0 * sh * sh / th2;
    }
    R[0] = 1.0 - b * (y * y + z * z); R[1] = b * x * y - a * z;         R[2] = b * x * z + a * y;
    R[3] = b * x * y + a * z;         R[4] = 1.0 - b * (x * x + z * z); R[5] = b * y * z - a * x;
    R[6] = b * x * z - a * y;         R[7] = b * y * z + a * x;         R[8] = 1.0 - b * (x * x + y * y);
}

// out = J_l(om)^-1 psi = psi - om x psi / 2 + D om x (om x psi),  D = 1/th^2 - (1+cos th)/(2 th sin th).
// Maps a left-perturbation rotation step psi to the additive Rodrigues-vector step the reference applies.
MC_HD void left_jacobian_inv_apply(const double* om, const double* psi, double* out)
{
    const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    double D;
    if (th2 < 0.0625) {
        D = 1.0 / 12 + th2 * (1.0 / 720 + th2 * (1.0 / 30240 + th2 * (1.0 / 1209600 + th2 * (1.0 / 47900160))));
    } else {
        const double th = sqrt(th2);
        D = 1.0 / th2 - (1.0 + cos(th)) / (2.0 * th * sin(th));
    }
    double c1[3], c2[3];
    cross3(om, psi, c1);
    cross3(om, c1, c2);
    out[0] = psi[0] - 0.5 * c1[0] + D * c2[0];
    out[1] = psi[1] - 0.5 * c1[1] + D * c2[1];
    out[2] = psi[2] - 0.5 * c1[2] + D * c2[2];
}

MC_HD void mat3_mul(const double* A, const double* B, double* C)
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
MC_HD void mat3_vec(const double* A, const double* x, double* y)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = A[i * 3] * x[0] + A[i * 3 + 1] * x[1] + A[i * 3 + 2] * x[2];
}
MC_HD void mat3t_vec(const double* A, const double* x, double* y)  // A^T x
{
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2];
}

// ---------------------------------------------------------------------------------------------------------
// camera models: residual e = observed - projected and A = d(u,v)/dXc (2x3) for a point Xc in camera coordinates
// ---------------------------------------------------------------------------------------------------------
// Pinhole + radtan (+ rational), the per-point arithmetic of cv::projectPoints (call sites src/multicalib.cpp:771,
// 947).  K(0,1) is ignored, as OpenCV does.
template <bool kRational, bool kJac>
MC_HD void pinhole_point(const CamParams& c, const double* Xc, double* uv, double* A)
{
    const double iz = proj_rcp(Xc[2]);
    const double x = Xc[0] * iz, y = Xc[1] * iz;
    const double r2 = x * x + y * y;
    double rad = 1.0 + r2 * (c.k1 + r2 * (c.k2 + r2 * c.k3));
    double drad = c.k1 + r2 * (2.0 * c.k2 + 3.0 * c.k3 * r2);  // d rad / d r2
    if (kRational) {
        const double den = 1.0 + r2 * (c.k4 + r2 * (c.k5 + r2 * c.k6));
        const double dden = c.k4 + r2 * (2.0 * c.k5 + 3.0 * c.k6 * r2);
        const double iden = proj_rcp(den);
        drad = (drad - rad * iden * dden) * iden;
        rad = rad * iden;
    }
    const double xy2 = 2.0 * x * y;
    const double xd = x * rad + c.p1 * xy2 + c.p2 * (r2 + 2.0 * x * x);
    const double yd = y * rad + c.p1 * (r2 + 2.0 * y * y) + c.p2 * xy2;
    uv[0] = c.fx * xd + c.cx;
    uv[1] = c.fy * yd + c.cy;
    if (kJac) {
        const double dd2 = 2.0 * drad;
        const double t = 2.0 * (c.p1 * x + c.p2 * y);
        const double dxdx = rad + dd2 * x * x + 2.0 * c.p1 * y + 6.0 * c.p2 * x;
        const double dxdy = dd2 * x * y + t;
        const double dydy = rad + dd2 * y * y + 6.0 * c.p1 * y + 2.0 * c.p2 * x;
        const double fxz = c.fx * iz, fyz = c.fy * iz;
        A[0] = fxz * dxdx; A[1] = fxz * dxdy; A[2] = -(A[0] * x + A[1] * y);
        A[3] = fyz * dxdy; A[4] = fyz * dydy; A[5] = -(A[3] * x + A[4] * y);
    }
}

// Mei unified model, the per-point arithmetic of cv::omnidir::projectPoints (src/omnidir.cpp:146-165, 185-199).
template <bool kJac>
MC_HD void omnidir_point(const CamParams& c, const double* Xc, double* uv, double* A)
{
    const double n2 = Xc[0] * Xc[0] + Xc[1] * Xc[1] + Xc[2] * Xc[2];
    const double rn = proj_rsqrt(n2);
    const double s0 = Xc[0] * rn, s1 = Xc[1] * rn, s2 = Xc[2] * rn;
    const double id = proj_rcp(s2 + c.xi);
    const double x = s0 * id, y = s1 * id;
    const double r2 = x * x + y * y;
    const double rad = 1.0 + r2 * (c.k1 + r2 * c.k2);
    const double xy2 = 2.0 * x * y;
    const double xd = x * rad + c.p1 * xy2 + c.p2 * (r2 + 2.0 * x * x);
    const double yd = y * rad + c.p1 * (r2 + 2.0 * y * y) + c.p2 * xy2;
    uv[0] = c.fx * xd + c.skew * yd + c.cx;
    uv[1] = c.fy * yd + c.cy;
    if (kJac) {
        const double dd2 = 2.0 * (c.k1 + 2.0 * c.k2 * r2);
        const double t = 2.0 * (c.p1 * x + c.p2 * y);
        const double dxdx = rad + dd2 * x * x + 2.0 * c.p1 * y + 6.0 * c.p2 * x;
        const double dxdy = dd2 * x * y + t;
        const double dydy = rad + dd2 * y * y + 6.0 * c.p1 * y + 2.0 * c.p2 * x;
        // d(u,v)/d(x,y)
        const double m00 = c.fx * dxdx + c.skew * dxdy, m01 = c.fx * dxdy + c.skew * dydy;
        const double m10 = c.fy * dxdy, m11 = c.fy * dydy;
        // d(x,y)/dXs = id * [1 0 -x; 0 1 -y];  dXs/dXc = rn (I - Xs Xs^T)
        // row g = (g0, g1, g2) in Xs-space  ->  rn * (g - (g.Xs) Xs)
        const double k = rn * id;
        {
            const double g0 = m00, g1 = m01, g2 = -(m00 * x + m01 * y);
            const double d = g0 * s0 + g1 * s1 + g2 * s2;
            A[0] = k * (g0 - d * s0); A[1] = k * (g1 - d * s1); A[2] = k * (g2 - d * s2);
        }
        {
            const double g0 = m10, g1 = m11, g2 = -(m10 * x + m11 * y);
            const double d = g0 * s0 + g1 * s1 + g2 * s2;
            A[3] = k * (g0 - d * s0); A[4] = k * (g1 - d * s1); A[5] = k * (g2 - d * s2);
        }
    }
}

// Mei model with the intrinsic block: A = d(u,v)/dXc (2x3) and Jin = d(u,v)/d(fx, fy, s, cx, cy, xi, k1, k2, p1, p2)
// (2x10, the parameter order of omnidir::internal::encodeParameters, src/omnidir.cpp:1559-1567; same order as
// Jacobian columns 6..15, src/omnidir.cpp:65-73, 209-241).
MC_HD void omnidir_point_full(const CamParams& c, const double* Xc, double* uv, double* A, double* Jin)
{
    const double n2 = Xc[0] * Xc[0] + Xc[1] * Xc[1] + Xc[2] * Xc[2];
    const double rn = 1.0 / sqrt(n2);
    const double s0 = Xc[0] * rn, s1 = Xc[1] * rn, s2 = Xc[2] * rn;
    const double id = 1.0 / (s2 + c.xi);
    const double x = s0 * id, y = s1 * id;
    const double r2 = x * x + y * y, r4 = r2 * r2;
    const double rad = 1.0 + r2 * (c.k1 + r2 * c.k2);
    const double xy2 = 2.0 * x * y;
    const double xd = x * rad + c.p1 * xy2 + c.p2 * (r2 + 2.0 * x * x);
    const double yd = y * rad + c.p1 * (r2 + 2.0 * y * y) + c.p2 * xy2;
    uv[0] = c.fx * xd + c.skew * yd + c.cx;
    uv[1] = c.fy * yd + c.cy;
    const double dd2 = 2.0 * (c.k1 + 2.0 * c.k2 * r2);
    const double t = 2.0 * (c.p1 * x + c.p2 * y);
    const double dxdx = rad + dd2 * x * x + 2.0 * c.p1 * y + 6.0 * c.p2 * x;
    const double dxdy = dd2 * x * y + t;
    const double dydy = rad + dd2 * y * y + 6.0 * c.p1 * y + 2.0 * c.p2 * x;
    const double m00 = c.fx * dxdx + c.skew * dxdy, m01 = c.fx * dxdy + c.skew * dydy;
    const double m10 = c.fy * dxdy, m11 = c.fy * dydy;
    const double k = rn * id;
    {
        const double g0 = m00, g1 = m01, g2 = -(m00 * x + m01 * y);
        const double d = g0 * s0 + g1 * s1 + g2 * s2;
        A[0] = k * (g0 - d * s0); A[1] = k * (g1 - d * s1); A[2] = k * (g2 - d * s2);
    }
    {
        const double g0 = m10, g1 = m11, g2 = -(m10 * x + m11 * y);
        const double d = g0 * s0 + g1 * s1 + g2 * s2;
        A[3] = k * (g0 - d * s0); A[4] = k * (g1 - d * s1); A[5] = k * (g2 - d * s2);
    }
    // intrinsic columns
    const double dxi0 = -x * id, dxi1 = -y * id;          // d(x,y)/dxi = -Xs.xy / (Xs.z + xi)^2
    Jin[0] = xd;  Jin[10] = 0.0;                          // fx
    Jin[1] = 0.0; Jin[11] = yd;                           // fy
    Jin[2] = yd;  Jin[12] = 0.0;                          // s
    Jin[3] = 1.0; Jin[13] = 0.0;                          // cx
    Jin[4] = 0.0; Jin[14] = 1.0;                          // cy
    Jin[5] = m00 * dxi0 + m01 * dxi1; Jin[15] = m10 * dxi0 + m11 * dxi1;   // xi
    const double kx[4] = {x * r2, x * r4, xy2, r2 + 2.0 * x * x};          // d xd / d(k1,k2,p1,p2)
    const double ky[4] = {y * r2, y * r4, r2 + 2.0 * y * y, xy2};          // d yd / d(k1,k2,p1,p2)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        Jin[6 + j] = c.fx * kx[j] + c.skew * ky[j];
        Jin[16 + j] = c.fy * ky[j];
    }
}

// J_l(om): left Jacobian of SO(3), psi = J_l(om) d_om (row-major 3x3).
MC_HD void left_jacobian(const double* om, double* J)
{
    const double x = om[0], y = om[1], z = om[2];
    const double th2 = x * x + y * y + z * z;
    double B, C;
    if (th2 < 0.0625) {
        B = 0.5 - th2 * (1.0 / 24 - th2 * (1.0 / 720 - th2 * (1.0 / 40320 - th2 * (1.0 / 3628800))));
        C = 1.0 / 6 - th2 * (1.0 / 120 - th2 * (1.0 / 5040 - th2 * (1.0 / 362880 - th2 * (1.0 / 39916800))));
    } else {
        const double th = sqrt(th2);
        const double sh = sin(0.5 * th);
        B = 2.0 * sh * sh / th2;
        C = (th - sin(th)) / (th2 * th);
    }
    // I + B K + C K^2
    J[0] = 1.0 - C * (y * y + z * z); J[1] = C * x * y - B * z;         J[2] = C * x * z + B * y;
    J[3] = C * x * y + B * z;         J[4] = 1.0 - C * (x * x + z * z); J[5] = C * y * z - B * x;
    J[6] = C * x * z - B * y;         J[7] = C * y * z + B * x;         J[8] = 1.0 - C * (x * x + y * y);
}

// One corner: Xc = R3 X + T3, residual, 2x6 Jacobian wrt the left perturbation (phi3, tau3) of the composed pose,
// accumulated into acc[28] = upper triangle of sum J^T J (21) | sum J^T e (6) | sum |e|^2.
// Replaces one pass of the per-corner loops inside cv::projectPoints / omnidir::projectPoints plus the rows of
// J^T J / J^T E this corner contributes (src/multicalib.cpp:688-689, 771-797; src/omnidir.cpp:141-244).
template <int kModel, bool kRational>
MC_HD void corner_accumulate(const CamParams& c, const double* R3, const double* T3, float ox, float oy, float oz,
                             float iu, float iv, double* acc)
{
    const double X[3] = {(double)ox, (double)oy, (double)oz};
    double Q[3], Xc[3], uv[2], A[6];
    mat3_vec(R3, X, Q);
    Xc[0] = Q[0] + T3[0]; Xc[1] = Q[1] + T3[1]; Xc[2] = Q[2] + T3[2];
    if (kModel == kPinhole) pinhole_point<kRational, true>(c, Xc, uv, A);
    else omnidir_point<true>(c, Xc, uv, A);
    const double e0 = (double)iu - uv[0], e1 = (double)iv - uv[1];
    double j0[6], j1[6];
    cross3(Q, A, j0);       // d u / d phi3 = Q x a0
    cross3(Q, A + 3, j1);
    j0[3] = A[0]; j0[4] = A[1]; j0[5] = A[2];
    j1[3] = A[3]; j1[4] = A[4]; j1[5] = A[5];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = i; j < 6; ++j) acc[tri6(i, j)] = fma(j1[i], j1[j], fma(j0[i], j0[j], acc[tri6(i, j)]));
        acc[21 + i] = fma(j1[i], e1, fma(j0[i], e0, acc[21 + i]));
    }
    acc[27] = fma(e1, e1, fma(e0, e0, acc[27]));
}

// Residual only: returns |e|^2 and |e| (for computeProjectError, src/multicalib.cpp:969-983).
template <int kModel, bool kRational>
MC_HD void corner_error(const CamParams& c, const double* R3, const double* T3, float ox, float oy, float oz, float iu,
                        float iv, double* sq, double* nrm)
{
    const double X[3] = {(double)ox, (double)oy, (double)oz};
    double Xc[3], uv[2];
    mat3_vec(R3, X, Xc);
    Xc[0] += T3[0]; Xc[1] += T3[1]; Xc[2] += T3[2];
    if (kModel == kPinhole) pinhole_point<kRational, false>(c, Xc, uv, nullptr);
    else omnidir_point<false>(c, Xc, uv, nullptr);
    const double e0 = (double)iu - uv[0], e1 = (double)iv - uv[1];
    const double s = e0 * e0 + e1 * e1;
    *sq += s;
    *nrm += sqrt(s);
}

// Composed pose of an edge: R3 = Rc Rp, T3 = Rc tp + tc  (compose_motion, src/multicalib.cpp:1030, 1045-1051).
MC_HD void compose_pose(const double* Rc, const double* tc, const double* Rp, const double* tp, double* R3, double* T3)
{
    mat3_mul(Rc, Rp, R3);
    double s[3];
    mat3_vec(Rc, tp, s);
    T3[0] = s[0] + tc[0]; T3[1] = s[1] + tc[1]; T3[2] = s[2] + tc[2];
}

// ---------------------------------------------------------------------------------------------------------
// lifting of a per-edge block (H6, g6) to pattern-pose and camera tangent coordinates
//   pattern pose p: (phi3, tau3) = blockdiag(Rc, Rc) (psi_p, dt_p)
//   camera c      : (phi3, tau3) = [[I, 0], [-[s]x, I]] (psi_c, dt_c),  s = Rc tp
// ---------------------------------------------------------------------------------------------------------
MC_HD void unpack_sym6(const double* tri, double* H)
{
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 6; ++j) {
            H[i * 6 + j] = tri[tri6(i, j)];
            H[j * 6 + i] = tri[tri6(i, j)];
        }
}

// Hpp (21, packed upper) += blockdiag(Rc,Rc)^T H blockdiag(Rc,Rc);  gp += blockdiag(Rc,Rc)^T g.
// H is the full 6x6; gauge != 0 means Rc = I.
MC_HD void lift_frame(const double* H, const double* g, const double* Rc, int gauge, double* Hpp, double* gp)
{
    if (gauge) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = i; j < 6; ++j) Hpp[tri6(i, j)] += H[i * 6 + j];
            gp[i] += g[i];
        }
        return;
    }
    // T = H * blockdiag(Rc,Rc)
    double T[36];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                T[i * 6 + 3 * b + j] = H[i * 6 + 3 * b] * Rc[j] + H[i * 6 + 3 * b + 1] * Rc[3 + j] + H[i * 6 + 3 * b + 2] * Rc[6 + j];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int r = 3 * a + i;
#pragma unroll
            for (int j = r; j < 6; ++j)
                Hpp[tri6(r, j)] += Rc[i] * T[(3 * a) * 6 + j] + Rc[3 + i] * T[(3 * a + 1) * 6 + j] + Rc[6 + i] * T[(3 * a + 2) * 6 + j];
            gp[r] += Rc[i] * g[3 * a] + Rc[3 + i] * g[3 * a + 1] + Rc[6 + i] * g[3 * a + 2];
        }
}

// Camera-side lift: M = H Cc, Hcc = Cc^T M (full 6x6), gc = Cc^T g, W = blockdiag(Rc,Rc)^T M (pattern rows x
// camera columns).
MC_HD void lift_camera(const double* H, const double* g, const double* Rc, const double* s, double* Hcc, double* gc, double* W)
{
    double M[36];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double c[3];
        cross3(s, H + i * 6 + 3, c);  // s x H_tau(i)
        M[i * 6 + 0] = H[i * 6 + 0] + c[0];
        M[i * 6 + 1] = H[i * 6 + 1] + c[1];
        M[i * 6 + 2] = H[i * 6 + 2] + c[2];
        M[i * 6 + 3] = H[i * 6 + 3];
        M[i * 6 + 4] = H[i * 6 + 4];
        M[i * 6 + 5] = H[i * 6 + 5];
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const double top[3] = {M[j], M[6 + j], M[12 + j]}, bot[3] = {M[18 + j], M[24 + j], M[30 + j]};
        double c[3], w[3];
        cross3(s, bot, c);
        Hcc[j] = top[0] + c[0]; Hcc[6 + j] = top[1] + c[1]; Hcc[12 + j] = top[2] + c[2];
        Hcc[18 + j] = bot[0]; Hcc[24 + j] = bot[1]; Hcc[30 + j] = bot[2];
        mat3t_vec(Rc, top, w);
        W[j] = w[0]; W[6 + j] = w[1]; W[12 + j] = w[2];
        mat3t_vec(Rc, bot, w);
        W[18 + j] = w[0]; W[24 + j] = w[1]; W[30 + j] = w[2];
    }
    double c[3];
    cross3(s, g + 3, c);
    gc[0] = g[0] + c[0]; gc[1] = g[1] + c[1]; gc[2] = g[2] + c[2];
    gc[3] = g[3]; gc[4] = g[4]; gc[5] = g[5];
}

// In-place Cholesky of a packed-upper SPD 6x6 (A = L L^T).  On return tri holds L^T (i.e. U = L^T, U[i][j], i<=j)
// with the RECIPROCAL of the diagonal on the diagonal.  Returns 0 if a pivot is not positive.
MC_HD int chol6_packed(double* U)
{
    int ok = 1;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double d = U[tri6(k, k)];
        if (!(d > 0.0)) { ok = 0; d = 1.0; }
        const double inv = 1.0 / sqrt(d);
        U[tri6(k, k)] = inv;
#pragma unroll
        for (int j = k + 1; j < 6; ++j) U[tri6(k, j)] *= inv;
#pragma unroll
        for (int i = k + 1; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) U[tri6(i, j)] -= U[tri6(k, i)] * U[tri6(k, j)];
    }
    return ok;
}
// y = L^-1 b (forward substitution with L = U^T, reciprocal diagonal), b overwritten; stride = element stride of b
MC_HD void chol6_forward(const double* U, double* b, int stride)
{
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = b[i * stride];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= U[tri6(k, i)] * b[k * stride];
        b[i * stride] = s * U[tri6(i, i)];
    }
}
// x = L^-T y (backward substitution), y overwritten
MC_HD void chol6_backward(const double* U, double* y)
{
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) s -= U[tri6(i, k)] * y[k];
        y[i] = s * U[tri6(i, i)];
    }
}

}  // namespace mccba
