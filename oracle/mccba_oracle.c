/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the calibration bundle-adjustment hot path.
 *
 * Nothing under oracle/ is part of the product.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product path
 * (multi_camera_calibration_b200/csrc) never links, loads or calls it.
 *
 * PARITY STATUS: "parity unpinned" by the reference's own tests -- the reference ships no tests, golden
 * vectors or expected outputs for this path, and it cannot be compiled in this image
 * (needs OpenCV C++ headers and Eigen, neither present).  The oracle is pinned instead against
 *   tier 0: OpenCV 4.13.0 through Python cv2 (projectPoints, Rodrigues, composeRT) -> tests/golden/ npz files
 *   tier 1: a numpy transcription of src/omnidir.cpp:126-243 (finite-difference checked)
 *   tier 2: a dense literal numpy re-enactment of the reference loops (oracle/dense_reenact.py)
 *
 * What is restated (citations relative to /root/reference):
 *   orc_rodrigues            cv::Rodrigues vec->mat + 3x9 Jacobian (third-party, OpenCV calib3d 4.x;
 *                            call sites src/multicalib.cpp:1023-1024, src/omnidir.cpp:126-128)
 *   orc_rodrigues_inv        cv::Rodrigues mat->vec (call site src/multicalib.cpp:1035)
 *   orc_compose_motion       src/multicalib.cpp:1008-1056 (dup src/omnidir.cpp:1023-1065)
 *   orc_project_pinhole      cv::projectPoints (third-party; call sites src/multicalib.cpp:771, 947)
 *   orc_project_omnidir      src/omnidir.cpp:84-245
 *   orc_rig_eval             src/multicalib.cpp:593-703 + 717-824, block-sparse instead of dense J
 *   orc_rig_solve            src/multicalib.cpp:462-514 (mode 0 = the reference's step-scaled Gauss-Newton)
 *                            + a true LM (mode 1) that the reference does not have
 *   orc_rig_error            src/multicalib.cpp:895-1006 and the fp64 RMS of src/omnidir.cpp:1794-1802
 *   orc_omni_solve           src/omnidir.cpp:851-935, 1119-1147, 2031-2076, 2138-2153
 *
 * The reference materialises a dense 2M x P Jacobian and solves the full P x P system with Eigen CG.
 * The normal equations have block structure (one 6x6 block per vertex, one 6x6 coupling per edge), so this
 * oracle accumulates the same sums block by block and eliminates the per-frame blocks with a Schur
 * complement; the solution of J^T J x = J^T E is the same vector.  tests/test_oracle_dense.py checks that
 * against the dense re-enactment.  All arithmetic stays in the reference's own parametrisation (additive
 * Rodrigues vectors) -- the CUDA path works in a different (tangent-space) formulation on purpose, so the
 * comparison between the two is between independent derivations.
 *
 * Precision policy: policy 0 = fp64 everywhere; policy 1 = "faithful_f32", the
 * reference's float32 round trips (parameters stored float32, composed pose rounded to float32, projected
 * points rounded to float32, float32 residual, float32 step and update); policy 2 = "fp64_direct": fp64 like
 * policy 0, except that the composed rotation R3 = R2 R1 reaches the projection AS A MATRIX instead of through the
 * reference's Rodrigues(R3) -> om3 -> Rodrigues(om3) round trip (src/multicalib.cpp:1035 followed by the
 * projectPoints call at :771/:787).  That round trip is inexact for edges whose composed rotation angle is close to
 * pi: for sin(theta3) < 1e-5 the axis is read off the diagonal of R3 (as cv::Rodrigues does), an approximation with
 * relative error ~(pi - theta3)^2 / (8 a_i^2), and just outside that branch the axis comes from R - R^T with error
 * eps / sin(theta3).  Among 200k edges some fall there and the reference's own residuals carry ~1e-9 relative noise,
 * which the weakly damped 64-camera system amplifies to ~1e-6..1e-5 in the parameters.  Policy 2 removes exactly that
 * noise and nothing else (the chain
 * rule still runs in Rodrigues-vector coordinates); tests use it to separate reference noise from build error.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PINHOLE 0
#define ORC_OMNIDIR 1
#define ORC_MAX_DIST 8

/* ------------------------------------------------------------------------------------------------ */
/* small dense helpers                                                                              */
/* ------------------------------------------------------------------------------------------------ */
static void mat3_mul(const double *A, const double *B, double *C)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
static void mat3_vec(const double *A, const double *x, double *y)
{
    for (int i = 0; i < 3; ++i) y[i] = A[i * 3] * x[0] + A[i * 3 + 1] * x[1] + A[i * 3 + 2] * x[2];
}
static void skew3(const double *v, double *K)
{
    K[0] = 0; K[1] = -v[2]; K[2] = v[1];
    K[3] = v[2]; K[4] = 0; K[5] = -v[0];
    K[6] = -v[1]; K[7] = v[0]; K[8] = 0;
}

/* coefficients of the SO(3) left Jacobian J_l = I + B K + C K^2 and its inverse I - K/2 + D K^2 */
static void so3_coeffs(double th2, double *B, double *C, double *D)
{
    double th = sqrt(th2);
    if (th < 0.25) {
        double t2 = th2;
        *B = 0.5 - t2 / 24 + t2 * t2 / 720 - t2 * t2 * t2 / 40320 + t2 * t2 * t2 * t2 / 3628800;
        *C = 1.0 / 6 - t2 / 120 + t2 * t2 / 5040 - t2 * t2 * t2 / 362880 + t2 * t2 * t2 * t2 / 39916800;
        *D = 1.0 / 12 + t2 / 720 + t2 * t2 / 30240 + t2 * t2 * t2 / 1209600 + t2 * t2 * t2 * t2 / 47900160;
    } else {
        double s = sin(th), c = cos(th), sh = sin(0.5 * th);
        *B = 2 * sh * sh / th2;
        *C = (th - s) / (th2 * th);
        *D = 1.0 / th2 - (1 + c) / (2 * th * s);
    }
}
static void so3_left_jacobian(const double *om, double *J)
{
    double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2], B, C, D, K[9], K2[9];
    so3_coeffs(th2, &B, &C, &D);
    skew3(om, K);
    mat3_mul(K, K, K2);
    for (int i = 0; i < 9; ++i) J[i] = B * K[i] + C * K2[i];
    J[0] += 1; J[4] += 1; J[8] += 1;
}
static void so3_left_jacobian_inv(const double *om, double *J)
{
    double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2], B, C, D, K[9], K2[9];
    so3_coeffs(th2, &B, &C, &D);
    skew3(om, K);
    mat3_mul(K, K, K2);
    for (int i = 0; i < 9; ++i) J[i] = -0.5 * K[i] + D * K2[i];
    J[0] += 1; J[4] += 1; J[8] += 1;
}

/* cv::Rodrigues, vector -> matrix.  R row-major 3x3.  dRdom (optional) is 3x9 like OpenCV's output:
 * row k = d(vec R)/d om_k with R vectorised row-major.  Closed form: dR/dom_k = [J_l e_k]_x R. */
int orc_rodrigues(const double *om, double *R, double *dRdom)
{
    double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    double th = sqrt(th2);
    double K[9], K2[9];
    skew3(om, K);
    mat3_mul(K, K, K2);
    double a, b; /* R = I + a K + b K^2, a = sin(th)/th, b = (1-cos th)/th^2 */
    if (th < 1e-4) {
        a = 1 - th2 / 6 + th2 * th2 / 120;
        b = 0.5 - th2 / 24 + th2 * th2 / 720;
    } else {
        double sh = sin(0.5 * th);
        a = sin(th) / th;
        b = 2 * sh * sh / th2;
    }
    for (int i = 0; i < 9; ++i) R[i] = a * K[i] + b * K2[i];
    R[0] += 1; R[4] += 1; R[8] += 1;
    if (dRdom) {
        double Jl[9];
        so3_left_jacobian(om, Jl);
        for (int k = 0; k < 3; ++k) {
            double col[3] = {Jl[k], Jl[3 + k], Jl[6 + k]}, S[9], dR[9];
            skew3(col, S);
            mat3_mul(S, R, dR);
            for (int i = 0; i < 9; ++i) dRdom[k * 9 + i] = dR[i];
        }
    }
    return 0;
}

/* cv::Rodrigues, matrix -> vector (R assumed orthonormal; OpenCV additionally re-orthonormalises by SVD). */
int orc_rodrigues_inv(const double *R, double *om)
{
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1 ? 1 : (c < -1 ? -1 : c);
    double th = atan2(s, c);
    if (s < 1e-5) {
        if (c > 0) { /* theta ~ 0: om = vee(R - R^T)/2 to first order */
            om[0] = 0.5 * rx; om[1] = 0.5 * ry; om[2] = 0.5 * rz;
        } else { /* theta ~ pi: take the axis from the diagonal, as OpenCV does */
            double t;
            t = (R[0] + 1) * 0.5; om[0] = sqrt(t > 0 ? t : 0);
            t = (R[4] + 1) * 0.5; om[1] = sqrt(t > 0 ? t : 0) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5; om[2] = sqrt(t > 0 ? t : 0) * (R[2] < 0 ? -1. : 1.);
            if (fabs(om[0]) < fabs(om[1]) && fabs(om[0]) < fabs(om[2]) && ((R[5] > 0) != (om[1] * om[2] > 0)))
                om[2] = -om[2];
            double n = th / sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
            om[0] *= n; om[1] *= n; om[2] *= n;
        }
        return 0;
    }
    double vth = th / (2 * s);
    om[0] = rx * vth; om[1] = ry * vth; om[2] = rz * vth;
    return 0;
}

/* src/multicalib.cpp:1008-1056.  (om3,T3) = (om2,T2) o (om1,T1), i.e. R3 = R2 R1, T3 = R2 T1 + T2.
 * d[8][9] (each row-major 3x3) in the reference's argument order:
 *   0 dom3dom1, 1 dom3dT1, 2 dom3dom2, 3 dom3dT2, 4 dT3dom1, 5 dT3dT1, 6 dT3dom2, 7 dT3dT2.
 * The reference chains dom3dR3 * dR3dR1 * dR1dom1 (9x9 products); because dR3dR1*dR1dom1 maps into the
 * tangent space of SO(3) at R3 the product equals J_l(om3)^-1 R2 J_l(om1) (verified
 * against cv2.composeRT in tests/test_oracle_primitives.py). */
static int compose_motion_ex(const double *om1, const double *T1, const double *om2, const double *T2,
                             double *om3, double *T3, double *d, double *R3);
int orc_compose_motion(const double *om1, const double *T1, const double *om2, const double *T2,
                       double *om3, double *T3, double *d)
{
    double R3[9];
    return compose_motion_ex(om1, T1, om2, T2, om3, T3, d, R3);
}
static int compose_motion_ex(const double *om1, const double *T1, const double *om2, const double *T2,
                             double *om3, double *T3, double *d, double *R3)
{
    double R1[9], R2[9], t[3];
    orc_rodrigues(om1, R1, 0);
    orc_rodrigues(om2, R2, 0);
    mat3_mul(R2, R1, R3);
    orc_rodrigues_inv(R3, om3);
    mat3_vec(R2, T1, t);
    T3[0] = t[0] + T2[0]; T3[1] = t[1] + T2[1]; T3[2] = t[2] + T2[2];
    if (d) {
        double Jl1[9], Jl2[9], Jl3i[9], tmp[9], S[9];
        so3_left_jacobian(om1, Jl1);
        so3_left_jacobian(om2, Jl2);
        so3_left_jacobian_inv(om3, Jl3i);
        memset(d, 0, sizeof(double) * 72);
        mat3_mul(R2, Jl1, tmp);
        mat3_mul(Jl3i, tmp, d + 0 * 9);          /* dom3dom1 */
        mat3_mul(Jl3i, Jl2, d + 2 * 9);          /* dom3dom2 */
        memcpy(d + 5 * 9, R2, sizeof(double) * 9); /* dT3dT1 */
        skew3(t, S);                             /* dT3dom2 = -[R2 T1]_x J_l(om2) */
        mat3_mul(S, Jl2, tmp);
        for (int i = 0; i < 9; ++i) d[6 * 9 + i] = -tmp[i];
        d[7 * 9 + 0] = d[7 * 9 + 4] = d[7 * 9 + 8] = 1; /* dT3dT2 = I */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* camera models: projection + d(u,v)/d(Xc) (2x3 "A") for one point given in camera coordinates      */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int model;
    double fx, fy, cx, cy, skew; /* skew used by the Mei model only: cv::projectPoints ignores K(0,1) */
    double k[ORC_MAX_DIST];      /* pinhole: k1 k2 p1 p2 k3 k4 k5 k6; omnidir: k1 k2 p1 p2 */
    double xi;
} orc_cam;

/* cv::projectPoints per-point arithmetic (OpenCV calib3d, rational model without thin prism / tilt). */
static void pinhole_point(const orc_cam *c, const double *Xc, double *uv, double *A)
{
    double X = Xc[0], Y = Xc[1], Z = Xc[2];
    double z = Z != 0 ? 1. / Z : 1.;
    double x = X * z, y = Y * z;
    double k1 = c->k[0], k2 = c->k[1], p1 = c->k[2], p2 = c->k[3], k3 = c->k[4], k4 = c->k[5], k5 = c->k[6], k6 = c->k[7];
    double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
    double cdist = 1 + k1 * r2 + k2 * r4 + k3 * r6;
    double icdist2 = 1. / (1 + k4 * r2 + k5 * r4 + k6 * r6);
    double xd0 = x * cdist * icdist2 + p1 * a1 + p2 * a2;
    double yd0 = y * cdist * icdist2 + p1 * a3 + p2 * a1;
    uv[0] = xd0 * c->fx + c->cx;
    uv[1] = yd0 * c->fy + c->cy;
    if (!A) return;
    double dxdt[3] = {z, 0, -x * z}, dydt[3] = {0, z, -y * z};
    for (int j = 0; j < 3; ++j) {
        double dr2dt = 2 * x * dxdt[j] + 2 * y * dydt[j];
        double dcdist_dt = k1 * dr2dt + 2 * k2 * r2 * dr2dt + 3 * k3 * r4 * dr2dt;
        double dicdist2_dt = -icdist2 * icdist2 * (k4 * dr2dt + 2 * k5 * r2 * dr2dt + 3 * k6 * r4 * dr2dt);
        double da1dt = 2 * (x * dydt[j] + y * dxdt[j]);
        double dmxdt = dxdt[j] * cdist * icdist2 + x * dcdist_dt * icdist2 + x * cdist * dicdist2_dt +
                       p1 * da1dt + p2 * (dr2dt + 4 * x * dxdt[j]);
        double dmydt = dydt[j] * cdist * icdist2 + y * dcdist_dt * icdist2 + y * cdist * dicdist2_dt +
                       p1 * (dr2dt + 4 * y * dydt[j]) + p2 * da1dt;
        A[j] = c->fx * dmxdt;
        A[3 + j] = c->fy * dmydt;
    }
}

/* src/omnidir.cpp:146-165 (projection) and :185-199 (dxpddXc); also the intrinsic columns :209-241.
 * jin (optional): 2x10 = d(u,v)/d(fx, fy, s, cx, cy, xi, k1, k2, p1, p2) in the parameter-vector order of
 * src/omnidir.cpp:1559-1567 (note the Jacobian's own column order is df ds dc dxi dkp = the same). */
static void omnidir_point(const orc_cam *c, const double *Xc, double *uv, double *A, double *jin)
{
    double xi = c->xi, k1 = c->k[0], k2 = c->k[1], p1 = c->k[2], p2 = c->k[3];
    double nrm = sqrt(Xc[0] * Xc[0] + Xc[1] * Xc[1] + Xc[2] * Xc[2]);
    double Xs[3] = {Xc[0] / nrm, Xc[1] / nrm, Xc[2] / nrm};
    double den = Xs[2] + xi;
    double xu0 = Xs[0] / den, xu1 = Xs[1] / den;
    double r2 = xu0 * xu0 + xu1 * xu1, r4 = r2 * r2;
    double xd0 = xu0 * (1 + k1 * r2 + k2 * r4) + 2 * p1 * xu0 * xu1 + p2 * (r2 + 2 * xu0 * xu0);
    double xd1 = xu1 * (1 + k1 * r2 + k2 * r4) + p1 * (r2 + 2 * xu1 * xu1) + 2 * p2 * xu0 * xu1;
    uv[0] = c->fx * xd0 + c->skew * xd1 + c->cx;
    uv[1] = c->fy * xd1 + c->cy;
    if (!A && !jin) return;
    double r_1 = 1.0 / nrm, r_3 = r_1 * r_1 * r_1;
    double dXsdXc[9] = {r_1 - Xc[0] * Xc[0] * r_3, -(Xc[0] * Xc[1]) * r_3, -(Xc[0] * Xc[2]) * r_3,
                        -(Xc[0] * Xc[1]) * r_3, r_1 - Xc[1] * Xc[1] * r_3, -(Xc[1] * Xc[2]) * r_3,
                        -(Xc[0] * Xc[2]) * r_3, -(Xc[1] * Xc[2]) * r_3, r_1 - Xc[2] * Xc[2] * r_3};
    double dxudXs[6] = {1 / den, 0, -Xs[0] / den / den, 0, 1 / den, -Xs[1] / den / den};
    double temp1 = 2 * k1 * xu0 + 4 * k2 * xu0 * r2;
    double temp2 = 2 * k1 * xu1 + 4 * k2 * xu1 * r2;
    double dxddxu[4] = {k2 * r4 + 6 * p2 * xu0 + 2 * p1 * xu1 + xu0 * temp1 + k1 * r2 + 1,
                        2 * p1 * xu0 + 2 * p2 * xu1 + xu0 * temp2,
                        2 * p1 * xu0 + 2 * p2 * xu1 + xu1 * temp1,
                        k2 * r4 + 2 * p2 * xu0 + 6 * p1 * xu1 + xu1 * temp2 + k1 * r2 + 1};
    double dxpddxd[4] = {c->fx, c->skew, 0, c->fy};
    double M22[4]; /* dxpddxd * dxddxu */
    M22[0] = dxpddxd[0] * dxddxu[0] + dxpddxd[1] * dxddxu[2];
    M22[1] = dxpddxd[0] * dxddxu[1] + dxpddxd[1] * dxddxu[3];
    M22[2] = dxpddxd[3] * dxddxu[2];
    M22[3] = dxpddxd[3] * dxddxu[3];
    if (A) {
        double M23[6]; /* M22 * dxudXs */
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 3; ++j) M23[i * 3 + j] = M22[i * 2] * dxudXs[j] + M22[i * 2 + 1] * dxudXs[3 + j];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 3; ++j)
                A[i * 3 + j] = M23[i * 3] * dXsdXc[j] + M23[i * 3 + 1] * dXsdXc[3 + j] + M23[i * 3 + 2] * dXsdXc[6 + j];
    }
    if (jin) {
        double dxudxi[2] = {-Xs[0] / den / den, -Xs[1] / den / den};
        double dkp[8] = {xu0 * r2, xu0 * r4, 2 * xu0 * xu1, r2 + 2 * xu0 * xu0,
                         xu1 * r2, xu1 * r4, r2 + 2 * xu1 * xu1, 2 * xu0 * xu1};
        memset(jin, 0, sizeof(double) * 20);
        jin[0] = xd0;            /* du/dfx */
        jin[10 + 1] = xd1;       /* dv/dfy */
        jin[2] = xd1;            /* du/ds  */
        jin[3] = 1;              /* du/dcx */
        jin[10 + 4] = 1;         /* dv/dcy */
        jin[5] = M22[0] * dxudxi[0] + M22[1] * dxudxi[1];
        jin[10 + 5] = M22[2] * dxudxi[0] + M22[3] * dxudxi[1];
        for (int j = 0; j < 4; ++j) {
            jin[6 + j] = dxpddxd[0] * dkp[j] + dxpddxd[1] * dkp[4 + j];
            jin[10 + 6 + j] = dxpddxd[3] * dkp[4 + j];
        }
    }
}

static void cam_from_arrays(orc_cam *c, int model, const double *K5, const double *dist, int ndist, double xi)
{
    memset(c, 0, sizeof(*c));
    c->model = model;
    c->fx = K5[0]; c->fy = K5[1]; c->cx = K5[2]; c->cy = K5[3]; c->skew = K5[4];
    for (int i = 0; i < ndist && i < ORC_MAX_DIST; ++i) c->k[i] = dist[i];
    c->xi = xi;
}

/* Projection of n world points with the 2n x 6 Jacobian wrt (om, T) exactly as both projectPoints flavours
 * form it: d/dT = A, d/dom = A * dXc/dom with dXc/dom_k = (dR/dom_k) Xw (cv::Rodrigues Jacobian). */
static void project_points(const orc_cam *c, int n, const double *obj, const double *om, const double *T,
                           double *proj, double *jac6, double *jin10)
{
    double R[9], dRdom[27];
    orc_rodrigues(om, R, jac6 ? dRdom : 0);
    for (int i = 0; i < n; ++i) {
        const double *Xw = obj + 3 * i;
        double Xc[3], A[6];
        mat3_vec(R, Xw, Xc);
        Xc[0] += T[0]; Xc[1] += T[1]; Xc[2] += T[2];
        if (c->model == ORC_PINHOLE)
            pinhole_point(c, Xc, proj + 2 * i, jac6 ? A : 0);
        else
            omnidir_point(c, Xc, proj + 2 * i, jac6 ? A : 0, jin10 ? jin10 + 20 * i : 0);
        if (jac6) {
            double *r0 = jac6 + 12 * i, *r1 = r0 + 6;
            for (int k = 0; k < 3; ++k) {
                double dX[3];
                mat3_vec(dRdom + 9 * k, Xw, dX);
                r0[k] = A[0] * dX[0] + A[1] * dX[1] + A[2] * dX[2];
                r1[k] = A[3] * dX[0] + A[4] * dX[1] + A[5] * dX[2];
            }
            r0[3] = A[0]; r0[4] = A[1]; r0[5] = A[2];
            r1[3] = A[3]; r1[4] = A[4]; r1[5] = A[5];
        }
    }
}

int orc_project_pinhole(int n, const double *obj, const double *om, const double *T, const double *K5,
                        const double *dist, int ndist, double *proj, double *jac6)
{
    orc_cam c;
    if (ndist != 0 && ndist != 4 && ndist != 5 && ndist != 8) return 1;
    cam_from_arrays(&c, ORC_PINHOLE, K5, dist, ndist, 0);
    project_points(&c, n, obj, om, T, proj, jac6, 0);
    return 0;
}

/* jac16 layout = src/omnidir.cpp:65-73: dom(3) dT(3) df(2) ds(1) dc(2) dxi(1) dkp(4), row-major 2n x 16 */
int orc_project_omnidir(int n, const double *obj, const double *om, const double *T, const double *K5,
                        double xi, const double *D4, double *proj, double *jac16)
{
    orc_cam c;
    cam_from_arrays(&c, ORC_OMNIDIR, K5, D4, 4, xi);
    if (!jac16) {
        project_points(&c, n, obj, om, T, proj, 0, 0);
        return 0;
    }
    double *j6 = (double *)malloc(sizeof(double) * 12 * (size_t)n);
    double *j10 = (double *)malloc(sizeof(double) * 20 * (size_t)n);
    project_points(&c, n, obj, om, T, proj, j6, j10);
    for (int i = 0; i < n; ++i)
        for (int r = 0; r < 2; ++r) {
            double *dst = jac16 + (size_t)(2 * i + r) * 16;
            memcpy(dst, j6 + 12 * i + 6 * r, sizeof(double) * 6);
            const double *q = j10 + 20 * i + 10 * r; /* fx fy s cx cy xi k1 k2 p1 p2 */
            dst[6] = q[0]; dst[7] = q[1]; dst[8] = q[2]; dst[9] = q[3]; dst[10] = q[4];
            dst[11] = q[5]; dst[12] = q[6]; dst[13] = q[7]; dst[14] = q[8]; dst[15] = q[9];
        }
    free(j6);
    free(j10);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* small SPD solves                                                                                 */
/* ------------------------------------------------------------------------------------------------ */
/* in-place lower Cholesky of an n x n row-major SPD matrix (lda = n); returns 0 on success */
static int chol_lower(double *A, int n)
{
    for (int j = 0; j < n; ++j) {
        double d = A[j * n + j];
        for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
        if (!(d > 0)) return j + 1;
        d = sqrt(d);
        A[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[i * n + j];
            for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = s / d;
        }
    }
    return 0;
}
static void chol_solve(const double *L, int n, double *b)
{
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k];
        b[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k];
        b[i] = s / L[i * n + i];
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* rig problem                                                                                      */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int n_cam, n_frame, n_edge;
    int64_t n_pts;
    int *edge_cam, *edge_pv;   /* cameraVertex in [0,nC); photoVertex in [nC, nC+F) (multicalib.hpp:86-103) */
    int64_t *edge_off;         /* E+1 offsets in points; the reference's pointsLocation is 2x this */
    float *obj, *img;          /* AoS xyz / uv float32, as the reference stores CV_32F points */
    orc_cam *cam;
    int *frame_off, *frame_edges; /* CSR frame -> edges */
    /* per-edge blocks from the last eval, in the reference's parametrisation (rvec, tvec) */
    double *Hcc, *Hpp, *Wpc, *gc, *gp, *cost_e, *sumnorm_e;
    double *H6, *g6;           /* per-edge blocks wrt the composed (om3,T3), for cross-checking the CUDA path */
    double *alt[9];            /* spare block set (LM keeps the accepted point's blocks across a rejection) */
} orc_rig;

void orc_rig_destroy(orc_rig *r)
{
    if (!r) return;
    free(r->edge_cam); free(r->edge_pv); free(r->edge_off); free(r->obj); free(r->img); free(r->cam);
    free(r->frame_off); free(r->frame_edges);
    free(r->Hcc); free(r->Hpp); free(r->Wpc); free(r->gc); free(r->gp); free(r->cost_e); free(r->sumnorm_e);
    free(r->H6); free(r->g6);
    for (int i = 0; i < 9; ++i) free(r->alt[i]);
    free(r);
}

orc_rig *orc_rig_create(int n_cam, int n_frame, int n_edge, const int *edge_cam, const int *edge_pv,
                        const int64_t *edge_off, const float *obj_xyz, const float *img_uv,
                        const int *cam_model, const double *cam_K5, const double *cam_dist8,
                        const int *cam_ndist, const double *cam_xi)
{
    orc_rig *r = (orc_rig *)calloc(1, sizeof(orc_rig));
    r->n_cam = n_cam; r->n_frame = n_frame; r->n_edge = n_edge;
    r->n_pts = edge_off[n_edge];
    size_t E = (size_t)n_edge, M = (size_t)r->n_pts;
    r->edge_cam = (int *)malloc(sizeof(int) * E);
    r->edge_pv = (int *)malloc(sizeof(int) * E);
    r->edge_off = (int64_t *)malloc(sizeof(int64_t) * (E + 1));
    memcpy(r->edge_cam, edge_cam, sizeof(int) * E);
    memcpy(r->edge_pv, edge_pv, sizeof(int) * E);
    memcpy(r->edge_off, edge_off, sizeof(int64_t) * (E + 1));
    r->obj = (float *)malloc(sizeof(float) * 3 * M);
    r->img = (float *)malloc(sizeof(float) * 2 * M);
    memcpy(r->obj, obj_xyz, sizeof(float) * 3 * M);
    memcpy(r->img, img_uv, sizeof(float) * 2 * M);
    r->cam = (orc_cam *)calloc((size_t)n_cam, sizeof(orc_cam));
    for (int c = 0; c < n_cam; ++c)
        cam_from_arrays(&r->cam[c], cam_model[c], cam_K5 + 5 * c, cam_dist8 + 8 * c, cam_ndist[c], cam_xi[c]);
    /* CSR frame -> edges, edges kept in ascending edge index within a frame */
    r->frame_off = (int *)calloc((size_t)n_frame + 1, sizeof(int));
    r->frame_edges = (int *)malloc(sizeof(int) * E);
    for (int e = 0; e < n_edge; ++e) {
        int f = edge_pv[e] - n_cam;
        if (f < 0 || f >= n_frame || edge_cam[e] < 0 || edge_cam[e] >= n_cam) { orc_rig_destroy(r); return 0; }
        r->frame_off[f + 1]++;
    }
    for (int f = 0; f < n_frame; ++f) r->frame_off[f + 1] += r->frame_off[f];
    int *fill = (int *)calloc((size_t)n_frame, sizeof(int));
    for (int e = 0; e < n_edge; ++e) {
        int f = edge_pv[e] - n_cam;
        r->frame_edges[r->frame_off[f] + fill[f]++] = e;
    }
    free(fill);
    r->Hcc = (double *)malloc(sizeof(double) * 36 * E);
    r->Hpp = (double *)malloc(sizeof(double) * 36 * E);
    r->Wpc = (double *)malloc(sizeof(double) * 36 * E);
    r->gc = (double *)malloc(sizeof(double) * 6 * E);
    r->gp = (double *)malloc(sizeof(double) * 6 * E);
    r->cost_e = (double *)malloc(sizeof(double) * E);
    r->sumnorm_e = (double *)malloc(sizeof(double) * E);
    r->H6 = (double *)malloc(sizeof(double) * 36 * E);
    r->g6 = (double *)malloc(sizeof(double) * 6 * E);
    return r;
}

static void rig_swap_blocks(orc_rig *r)
{
    double **cur[9] = {&r->Hcc, &r->Hpp, &r->Wpc, &r->gc, &r->gp, &r->cost_e, &r->sumnorm_e, &r->H6, &r->g6};
    size_t len[9] = {36, 36, 36, 6, 6, 1, 1, 36, 6};
    for (int i = 0; i < 9; ++i) {
        if (!r->alt[i]) r->alt[i] = (double *)malloc(sizeof(double) * len[i] * (size_t)r->n_edge);
        double *t = *cur[i]; *cur[i] = r->alt[i]; r->alt[i] = t;
    }
}

static void get_pose(const orc_rig *r, const double *params, int vertex, int policy, double *om, double *T)
{
    if (vertex == 0) { /* src/multicalib.cpp:636-640: camera 0 is the gauge */
        om[0] = om[1] = om[2] = T[0] = T[1] = T[2] = 0;
        return;
    }
    const double *p = params + 6 * (size_t)(vertex - 1);
    for (int i = 0; i < 3; ++i) {
        om[i] = policy == 1 ? (double)(float)p[i] : p[i];
        T[i] = policy == 1 ? (double)(float)p[3 + i] : p[3 + i];
    }
    (void)r;
}

/* One edge: src/multicalib.cpp:717-824.  Accumulates JcJc, JpJp, JpJc, Jc^T E, Jp^T E, |E|^2 over the
 * edge's corners.  The per-corner rows Jc = J6*Cc, Jp = J6*Cp are not materialised: sum_i J6_i^T J6_i is
 * formed once (H6) and multiplied by the 6x6 chain matrices per edge -- the same sums, re-associated. */
static void eval_edge(orc_rig *r, int e, const double *params, int policy, int want_blocks)
{
    int cam = r->edge_cam[e], pv = r->edge_pv[e];
    const orc_cam *c = &r->cam[cam];
    double omP[3], TP[3], omC[3], TC[3], om3[3], T3[3], d[72];
    get_pose(r, params, pv, policy, omP, TP);
    get_pose(r, params, cam, policy, omC, TC);
    double R3direct[9];
    compose_motion_ex(omP, TP, omC, TC, om3, T3, want_blocks ? d : 0, R3direct); /* call order of :734 */
    if (policy == 1) { /* :742-749 */
        for (int i = 0; i < 3; ++i) { om3[i] = (double)(float)om3[i]; T3[i] = (double)(float)T3[i]; }
    }
    double R3[9], dRdom[27];
    orc_rodrigues(om3, R3, want_blocks ? dRdom : 0);
    if (policy == 2) { /* fp64_direct: R3 = R2 R1 itself; dR/dom3_k = [J_l(om3) e_k]_x R3 with that R3 */
        memcpy(R3, R3direct, sizeof(R3));
        if (want_blocks) {
            double Jl[9];
            so3_left_jacobian(om3, Jl);
            for (int k = 0; k < 3; ++k) {
                double colk[3] = {Jl[k], Jl[3 + k], Jl[6 + k]}, S[9], dR[9];
                skew3(colk, S);
                mat3_mul(S, R3, dR);
                for (int i = 0; i < 9; ++i) dRdom[k * 9 + i] = dR[i];
            }
        }
    }
    double H[36], g[6], cost = 0, sumnorm = 0;
    memset(H, 0, sizeof(H));
    memset(g, 0, sizeof(g));
    for (int64_t i = r->edge_off[e]; i < r->edge_off[e + 1]; ++i) {
        double Xw[3] = {r->obj[3 * i], r->obj[3 * i + 1], r->obj[3 * i + 2]};
        double Xc[3], uv[2], A[6], E0, E1;
        mat3_vec(R3, Xw, Xc);
        Xc[0] += T3[0]; Xc[1] += T3[1]; Xc[2] += T3[2];
        if (c->model == ORC_PINHOLE) pinhole_point(c, Xc, uv, want_blocks ? A : 0);
        else omnidir_point(c, Xc, uv, want_blocks ? A : 0, 0);
        if (policy == 1) { /* projected points come back as float32 and the subtraction is float32 (:789-792) */
            E0 = (double)(float)(r->img[2 * i] - (float)uv[0]);
            E1 = (double)(float)(r->img[2 * i + 1] - (float)uv[1]);
        } else {
            E0 = (double)r->img[2 * i] - uv[0];
            E1 = (double)r->img[2 * i + 1] - uv[1];
        }
        cost += E0 * E0 + E1 * E1;
        sumnorm += sqrt(E0 * E0 + E1 * E1);
        if (!want_blocks) continue;
        double J[12];
        for (int k = 0; k < 3; ++k) {
            double dX[3];
            mat3_vec(dRdom + 9 * k, Xw, dX);
            J[k] = A[0] * dX[0] + A[1] * dX[1] + A[2] * dX[2];
            J[6 + k] = A[3] * dX[0] + A[4] * dX[1] + A[5] * dX[2];
        }
        J[3] = A[0]; J[4] = A[1]; J[5] = A[2];
        J[9] = A[3]; J[10] = A[4]; J[11] = A[5];
        for (int a = 0; a < 6; ++a) {
            for (int b = 0; b < 6; ++b) H[a * 6 + b] += J[a] * J[b] + J[6 + a] * J[6 + b];
            g[a] += J[a] * E0 + J[6 + a] * E1;
        }
    }
    r->cost_e[e] = cost;
    r->sumnorm_e[e] = sumnorm;
    if (!want_blocks) return;
    memcpy(r->H6 + 36 * (size_t)e, H, sizeof(H));
    memcpy(r->g6 + 6 * (size_t)e, g, sizeof(g));
    /* chain matrices: Cp = d(om3,T3)/d(omP,TP) (photo = pose 1), Cc = d(om3,T3)/d(omC,TC) (camera = pose 2) */
    double Cp[36], Cc[36];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Cp[i * 6 + j] = d[0 * 9 + i * 3 + j];           /* dom3dom1 */
            Cp[i * 6 + 3 + j] = d[1 * 9 + i * 3 + j];       /* dom3dT1  */
            Cp[(3 + i) * 6 + j] = d[4 * 9 + i * 3 + j];     /* dT3dom1  */
            Cp[(3 + i) * 6 + 3 + j] = d[5 * 9 + i * 3 + j]; /* dT3dT1   */
            Cc[i * 6 + j] = d[2 * 9 + i * 3 + j];           /* dom3dom2 */
            Cc[i * 6 + 3 + j] = d[3 * 9 + i * 3 + j];       /* dom3dT2  */
            Cc[(3 + i) * 6 + j] = d[6 * 9 + i * 3 + j];     /* dT3dom2  */
            Cc[(3 + i) * 6 + 3 + j] = d[7 * 9 + i * 3 + j]; /* dT3dT2   */
        }
    double HCp[36], HCc[36];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s1 = 0, s2 = 0;
            for (int k = 0; k < 6; ++k) { s1 += H[i * 6 + k] * Cp[k * 6 + j]; s2 += H[i * 6 + k] * Cc[k * 6 + j]; }
            HCp[i * 6 + j] = s1; HCc[i * 6 + j] = s2;
        }
    double *Hpp = r->Hpp + 36 * (size_t)e, *Hcc = r->Hcc + 36 * (size_t)e, *W = r->Wpc + 36 * (size_t)e;
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) {
            double s1 = 0, s2 = 0, s3 = 0;
            for (int k = 0; k < 6; ++k) {
                s1 += Cp[k * 6 + i] * HCp[k * 6 + j];
                s2 += Cc[k * 6 + i] * HCc[k * 6 + j];
                s3 += Cp[k * 6 + i] * HCc[k * 6 + j];
            }
            Hpp[i * 6 + j] = s1; Hcc[i * 6 + j] = s2; W[i * 6 + j] = s3;
        }
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 6; ++k) { s1 += Cp[k * 6 + i] * g[k]; s2 += Cc[k * 6 + i] * g[k]; }
        r->gp[6 * (size_t)e + i] = s1;
        r->gc[6 * (size_t)e + i] = s2;
    }
}

/* Evaluate all edges at params (length 6*(nC+F-1)).  Returns sum of squared residuals. */
double orc_rig_eval(orc_rig *r, const double *params, int policy, int want_blocks)
{
#pragma omp parallel for schedule(static)
    for (int e = 0; e < r->n_edge; ++e) eval_edge(r, e, params, policy, want_blocks);
    double cost = 0;
    for (int e = 0; e < r->n_edge; ++e) cost += r->cost_e[e];
    return cost;
}

/* copies of the per-edge blocks for tests: which = 0 H6(36) 1 g6(6) 2 Hcc 3 Hpp 4 Wpc 5 gc 6 gp 7 cost 8 sumnorm */
int orc_rig_get_blocks(const orc_rig *r, int which, double *out)
{
    size_t E = (size_t)r->n_edge;
    const double *src[9] = {r->H6, r->g6, r->Hcc, r->Hpp, r->Wpc, r->gc, r->gp, r->cost_e, r->sumnorm_e};
    size_t len[9] = {36, 6, 36, 36, 36, 6, 6, 1, 1};
    if (which < 0 || which > 8) return 1;
    memcpy(out, src[which], sizeof(double) * len[which] * E);
    return 0;
}

/* tangent transform T_v = blockdiag(J_l(om_v), I): psi = J_l(om) * d om (left perturbation of R_v). */
static void vertex_Tinv(const double *params, int vertex, double *Ti /*36*/)
{
    memset(Ti, 0, sizeof(double) * 36);
    double Jli[9];
    so3_left_jacobian_inv(params + 6 * (size_t)(vertex - 1), Jli);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Ti[i * 6 + j] = Jli[i * 3 + j];
    Ti[21] = Ti[28] = Ti[35] = 1;
}
static void congruence6(const double *A, const double *H, const double *B, double *out) /* A^T H B */
{
    double HB[36];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0;
            for (int k = 0; k < 6; ++k) s += H[i * 6 + k] * B[k * 6 + j];
            HB[i * 6 + j] = s;
        }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0;
            for (int k = 0; k < 6; ++k) s += A[k * 6 + i] * HB[k * 6 + j];
            out[i * 6 + j] = s;
        }
}

/* Solve the (optionally damped) normal equations built from the per-edge blocks of the last eval.
 * lambda == 0: plain J^T J x = J^T E in the reference's rvec parametrisation (src/multicalib.cpp:688-691).
 * lambda  > 0: LM system (H_t + lambda diag(H_t)) d = g_t posed in tangent coordinates psi = J_l(om) d_om
 *              (what the CUDA path uses); the returned step is converted back to additive rvec steps.
 * step: 6*(nV-1) output.  Returns 0, or >0 if a block is not SPD.  Also returns S and g_s (n_s = 6(nC-1)). */
int orc_rig_solve_normal(orc_rig *r, const double *params, double lambda, double *step, double *S_out, double *gs_out)
{
    int nC = r->n_cam, F = r->n_frame, ns = 6 * (nC - 1);
    int tangent = lambda > 0;
    double *S = (double *)calloc((size_t)(ns > 0 ? ns * ns : 1), sizeof(double));
    double *gs = (double *)calloc((size_t)(ns > 0 ? ns : 1), sizeof(double));
    double *Hpp_inv = (double *)malloc(sizeof(double) * 36 * (size_t)F); /* Cholesky factors of damped H_pp */
    double *gp = (double *)malloc(sizeof(double) * 6 * (size_t)F);
    double *W_t = (double *)malloc(sizeof(double) * 36 * (size_t)r->n_edge);
    int fail = 0;
    double *Tci = 0;
    if (tangent) {
        Tci = (double *)malloc(sizeof(double) * 36 * (size_t)nC);
        for (int c = 1; c < nC; ++c) vertex_Tinv(params, c, Tci + 36 * (size_t)c);
    }
    /* camera diagonal blocks first (their damping needs the complete diagonal) */
    for (int e = 0; e < r->n_edge; ++e) {
        int c = r->edge_cam[e];
        if (c == 0) continue;
        double blk[36];
        const double *Hcc = r->Hcc + 36 * (size_t)e;
        if (tangent) congruence6(Tci + 36 * (size_t)c, Hcc, Tci + 36 * (size_t)c, blk);
        else memcpy(blk, Hcc, sizeof(blk));
        int o = 6 * (c - 1);
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < 6; ++j) S[(o + i) * ns + o + j] += blk[i * 6 + j];
            double s = 0;
            for (int k = 0; k < 6; ++k)
                s += (tangent ? Tci[36 * (size_t)c + k * 6 + i] : (k == i ? 1. : 0.)) * r->gc[6 * (size_t)e + k];
            gs[o + i] += s;
        }
    }
    for (int i = 0; i < ns; ++i) S[i * ns + i] *= (1 + lambda);
    /* frames: independent given the camera blocks; every thread accumulates its own copy of the reduced system and
     * the copies are summed in thread order afterwards (deterministic for a fixed thread count) */
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    double *S_t = (double *)calloc((size_t)nthreads * (size_t)(ns > 0 ? ns * ns : 1), sizeof(double));
    double *gs_t = (double *)calloc((size_t)nthreads * (size_t)(ns > 0 ? ns : 1), sizeof(double));
#pragma omp parallel for schedule(static)
    for (int f = 0; f < F; ++f) {
        if (fail) continue;
        int tidx = 0;
#ifdef _OPENMP
        tidx = omp_get_thread_num();
#endif
        double *S = S_t + (size_t)tidx * (size_t)(ns > 0 ? ns * ns : 1);
        double *gs = gs_t + (size_t)tidx * (size_t)(ns > 0 ? ns : 1);
        double H[36], g[6], Tpi[36];
        memset(H, 0, sizeof(H));
        memset(g, 0, sizeof(g));
        int pv = nC + f;
        if (tangent) vertex_Tinv(params, pv, Tpi);
        for (int q = r->frame_off[f]; q < r->frame_off[f + 1]; ++q) {
            int e = r->frame_edges[q], c = r->edge_cam[e];
            double blk[36];
            if (tangent) congruence6(Tpi, r->Hpp + 36 * (size_t)e, Tpi, blk);
            else memcpy(blk, r->Hpp + 36 * (size_t)e, sizeof(blk));
            for (int i = 0; i < 36; ++i) H[i] += blk[i];
            for (int i = 0; i < 6; ++i) {
                double s = 0;
                for (int k = 0; k < 6; ++k)
                    s += (tangent ? Tpi[k * 6 + i] : (k == i ? 1. : 0.)) * r->gp[6 * (size_t)e + k];
                g[i] += s;
            }
            if (c > 0) {
                if (tangent) congruence6(Tpi, r->Wpc + 36 * (size_t)e, Tci + 36 * (size_t)c, W_t + 36 * (size_t)e);
                else memcpy(W_t + 36 * (size_t)e, r->Wpc + 36 * (size_t)e, sizeof(double) * 36);
            }
        }
        for (int i = 0; i < 6; ++i) H[i * 6 + i] *= (1 + lambda);
        if (chol_lower(H, 6)) { fail = 1000 + f; continue; }
        memcpy(Hpp_inv + 36 * (size_t)f, H, sizeof(H));
        memcpy(gp + 6 * (size_t)f, g, sizeof(g));
        /* Schur: S -= W_a^T H^-1 W_b, g_s -= W_a^T H^-1 g_p over the frame's non-gauge views */
        double Hig[6];
        memcpy(Hig, g, sizeof(g));
        chol_solve(H, 6, Hig);
        for (int qb = r->frame_off[f]; qb < r->frame_off[f + 1]; ++qb) {
            int eb = r->frame_edges[qb], cb = r->edge_cam[eb];
            if (cb == 0) continue;
            double HiW[36]; /* H^-1 W_b, column by column */
            for (int j = 0; j < 6; ++j) {
                double col[6];
                for (int i = 0; i < 6; ++i) col[i] = W_t[36 * (size_t)eb + i * 6 + j];
                chol_solve(H, 6, col);
                for (int i = 0; i < 6; ++i) HiW[i * 6 + j] = col[i];
            }
            for (int qa = r->frame_off[f]; qa < r->frame_off[f + 1]; ++qa) {
                int ea = r->frame_edges[qa], ca = r->edge_cam[ea];
                if (ca == 0) continue;
                const double *Wa = W_t + 36 * (size_t)ea;
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 6; ++j) {
                        double s = 0;
                        for (int k = 0; k < 6; ++k) s += Wa[k * 6 + i] * HiW[k * 6 + j];
                        S[(6 * (ca - 1) + i) * ns + 6 * (cb - 1) + j] -= s;
                    }
            }
            for (int i = 0; i < 6; ++i) {
                double s = 0;
                for (int k = 0; k < 6; ++k) s += W_t[36 * (size_t)eb + k * 6 + i] * Hig[k];
                gs[6 * (cb - 1) + i] -= s;
            }
        }
    }
    for (int t = 0; t < nthreads; ++t) {
        for (int i = 0; i < ns * ns; ++i) S[i] += S_t[(size_t)t * (size_t)(ns * ns) + i];
        for (int i = 0; i < ns; ++i) gs[i] += gs_t[(size_t)t * (size_t)ns + i];
    }
    free(S_t); free(gs_t);
    if (S_out && ns > 0) memcpy(S_out, S, sizeof(double) * (size_t)ns * ns);
    if (gs_out && ns > 0) memcpy(gs_out, gs, sizeof(double) * (size_t)ns);
    double *dc = (double *)calloc((size_t)(ns > 0 ? ns : 1), sizeof(double));
    if (!fail && ns > 0) {
        int rc = chol_lower(S, ns);
        if (rc) fail = rc;
        else {
            memcpy(dc, gs, sizeof(double) * (size_t)ns);
            chol_solve(S, ns, dc);
        }
    }
    if (!fail) {
        /* back-substitution: d_p = H^-1 (g_p - sum_v W_v d_c) */
#pragma omp parallel for schedule(static)
        for (int f = 0; f < F; ++f) {
            double rhs[6];
            memcpy(rhs, gp + 6 * (size_t)f, sizeof(rhs));
            for (int q = r->frame_off[f]; q < r->frame_off[f + 1]; ++q) {
                int e = r->frame_edges[q], c = r->edge_cam[e];
                if (c == 0) continue;
                for (int i = 0; i < 6; ++i) {
                    double s = 0;
                    for (int k = 0; k < 6; ++k) s += W_t[36 * (size_t)e + i * 6 + k] * dc[6 * (c - 1) + k];
                    rhs[i] -= s;
                }
            }
            chol_solve(Hpp_inv + 36 * (size_t)f, 6, rhs);
            int pv = nC + f;
            if (tangent) { /* psi -> d om */
                double Tpi[36], o[6];
                vertex_Tinv(params, pv, Tpi);
                for (int i = 0; i < 6; ++i) {
                    double s = 0;
                    for (int k = 0; k < 6; ++k) s += Tpi[i * 6 + k] * rhs[k];
                    o[i] = s;
                }
                memcpy(rhs, o, sizeof(o));
            }
            memcpy(step + 6 * (size_t)(pv - 1), rhs, sizeof(rhs));
        }
        for (int c = 1; c < nC; ++c) {
            double o[6];
            for (int i = 0; i < 6; ++i) {
                if (tangent) {
                    double s = 0;
                    for (int k = 0; k < 6; ++k) s += Tci[36 * (size_t)c + i * 6 + k] * dc[6 * (c - 1) + k];
                    o[i] = s;
                } else o[i] = dc[6 * (c - 1) + i];
            }
            memcpy(step + 6 * (size_t)(c - 1), o, sizeof(o));
        }
    }
    free(S); free(gs); free(Hpp_inv); free(gp); free(W_t); free(dc); free(Tci);
    return fail;
}

static double norm2(const double *x, size_t n)
{
    double s = 0;
    for (size_t i = 0; i < n; ++i) s += x[i] * x[i];
    return sqrt(s);
}

/* Outer loop.
 * mode 0: the reference's schedule, src/multicalib.cpp:473-507: x = solve(JTJ, JTE); G = 0.95^(iter+1) x;
 *         params += G; change = |G| / |params|.  Termination decoded from criteria.type as at :475-477.
 * mode 1: Levenberg-Marquardt (not in the reference): (H_t + lambda diag H_t) d = g_t; accept iff the
 *         cost decreases; lambda *= down on accept, *= up on reject; change from accepted steps only.
 * trace (optional): per iteration 5 doubles [cost at the linearisation point, change, lambda, accepted,
 *         cost at the trial point (mode 1) or 0].
 * report: [0] iterations run, [1] final change, [2] final cost (sum sq), [3] final lambda, [4] status. */
int orc_rig_solve(orc_rig *r, double *params, int mode, int crit_type, int max_count, double eps, int policy,
                  double lambda0, double lambda_up, double lambda_down, double *report, double *trace,
                  int trace_cap)
{
    size_t P = 6 * (size_t)(r->n_cam + r->n_frame - 1);
    double *step = (double *)malloc(sizeof(double) * P);
    double *trial = (double *)malloc(sizeof(double) * P);
    double change = 1, lambda = mode ? lambda0 : 0, cost = 0;
    int iter = 0, status = 0;
    if (policy == 1)
        for (size_t i = 0; i < P; ++i) params[i] = (double)(float)params[i]; /* CV_32F storage :426 */
    if (mode == 1) cost = orc_rig_eval(r, params, policy, 1);
    for (;; ++iter) {
        if ((crit_type == 1 && iter >= max_count) || (crit_type == 2 && change <= eps) ||
            (crit_type == 3 && (change <= eps || iter >= max_count)))
            break;
        if (mode == 0) {
            double alpha = pow(0.95, (double)iter + 1.0);
            cost = orc_rig_eval(r, params, policy, 1);
            status = orc_rig_solve_normal(r, params, 0, step, 0, 0);
            if (status) break;
            for (size_t i = 0; i < P; ++i) {
                double G = alpha * step[i];
                if (policy == 1) { /* :493-501 */
                    float Gf = (float)G;
                    step[i] = (double)Gf;
                    params[i] = (double)(float)((float)params[i] + Gf);
                } else {
                    step[i] = G;
                    params[i] += G;
                }
            }
            change = norm2(step, P) / norm2(params, P);
            if (trace && iter < trace_cap) {
                trace[5 * iter] = cost; trace[5 * iter + 1] = change; trace[5 * iter + 2] = 0;
                trace[5 * iter + 3] = 1; trace[5 * iter + 4] = 0;
            }
        } else {
            status = orc_rig_solve_normal(r, params, lambda, step, 0, 0);
            if (status) break;
            for (size_t i = 0; i < P; ++i) trial[i] = params[i] + step[i];
            /* evaluate the trial point with blocks into the spare block set; swap back on rejection */
            rig_swap_blocks(r);
            double cost_trial = orc_rig_eval(r, trial, policy, 1);
            int accept = cost_trial < cost;
            double lam_used = lambda, cost_before = cost;
            if (accept) {
                change = norm2(step, P) / norm2(trial, P);
                memcpy(params, trial, sizeof(double) * P);
                cost = cost_trial;
                lambda *= lambda_down;
                if (lambda < 1e-15) lambda = 1e-15;
            } else {
                rig_swap_blocks(r);
                lambda *= lambda_up;
                if (lambda > 1e15) lambda = 1e15;
            }
            if (trace && iter < trace_cap) {
                trace[5 * iter] = cost_before; trace[5 * iter + 1] = change; trace[5 * iter + 2] = lam_used;
                trace[5 * iter + 3] = accept; trace[5 * iter + 4] = cost_trial;
            }
        }
    }
    double final_cost = orc_rig_eval(r, params, policy, 0);
    if (report) {
        report[0] = iter; report[1] = change; report[2] = final_cost; report[3] = lambda; report[4] = status;
    }
    free(step);
    free(trial);
    return status;
}

/* src/multicalib.cpp:895-1006 + src/omnidir.cpp:1794-1802.
 * out[0] = the reference's meanReprojectError = sum ||e|| / totalNPoints with totalNPoints counting 2N per
 *          PINHOLE edge (error is N x 2 single channel, :983) and N per OMNIDIRECTIONAL edge;
 * out[1] = fp64 RMS sqrt(sum(ex^2+ey^2)/Npoints); out[2] = sum ||e||; out[3] = sum sq; out[4] = Npoints.
 * per_edge_mean (optional, E) = edge.reprojecterror (:979-980).
 * policy 1 accumulates in float32 like the reference; the compose there is done with float32 matrices
 * (:918-939) which this restatement approximates by rounding the composed pose to float32. */
int orc_rig_error(orc_rig *r, const double *params, int policy, double *out, double *per_edge_mean)
{
    orc_rig_eval(r, params, policy, 0);
    double tot = 0, sq = 0;
    float totf = 0;
    int64_t cnt_ref = 0, npts = 0;
    for (int e = 0; e < r->n_edge; ++e) {
        int64_t n = r->edge_off[e + 1] - r->edge_off[e];
        if (per_edge_mean) per_edge_mean[e] = n ? r->sumnorm_e[e] / (double)n : 0;
        tot += r->sumnorm_e[e];
        totf += (float)r->sumnorm_e[e];
        sq += r->cost_e[e];
        npts += n;
        cnt_ref += (r->cam[r->edge_cam[e]].model == ORC_PINHOLE) ? 2 * n : n;
    }
    out[0] = (policy == 1 ? (double)totf : tot) / (double)cnt_ref;
    out[1] = sqrt(sq / (double)npts);
    out[2] = tot; out[3] = sq; out[4] = (double)npts;
    return 0;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------ */
/* omnidir::calibrate loop (single camera, intrinsics + per-frame poses)                            */
/* ------------------------------------------------------------------------------------------------ */
/* src/omnidir.cpp:2031-2076: >= / subtract cascade over the flag bits, reproduced literally. */
void orc_omni_flags2idx(int flags, int n, int *idx /* 6n+10 */)
{
    for (int i = 0; i < 6 * n + 10; ++i) idx[i] = 1;
    int f = flags;
    if (f >= 256) { idx[6 * n + 3] = 0; idx[6 * n + 4] = 0; f -= 256; }
    if (f >= 128) { idx[6 * n] = 0; idx[6 * n + 1] = 0; f -= 128; }
    if (f >= 64) { idx[6 * n + 5] = 0; f -= 64; }
    if (f >= 32) { idx[6 * n + 9] = 0; f -= 32; }
    if (f >= 16) { idx[6 * n + 8] = 0; f -= 16; }
    if (f >= 8) { idx[6 * n + 7] = 0; f -= 8; }
    if (f >= 4) { idx[6 * n + 6] = 0; f -= 4; }
    if (f >= 2) { idx[6 * n + 2] = 0; }
}

/* One normal-equation build, src/omnidir.cpp:877-915, kept in arrow form:
 *   Hii (n x 36), Hi I (n x 60: 6 x 10), HII (100), gi (n x 6), gI (10).  Returns sum sq residual. */
static double omni_build(int n, const int64_t *off, const double *obj, const double *img, const double *param,
                         double *Hii, double *HiI, double *HII, double *gi, double *gI)
{
    orc_cam c;
    double K5[5] = {param[6 * n], param[6 * n + 1], param[6 * n + 3], param[6 * n + 4], param[6 * n + 2]};
    cam_from_arrays(&c, ORC_OMNIDIR, K5, param + 6 * n + 6, 4, param[6 * n + 5]);
    double cost = 0;
    memset(HII, 0, sizeof(double) * 100);
    memset(gI, 0, sizeof(double) * 10);
#pragma omp parallel
    {
        double HII_l[100], gI_l[10], cost_l = 0;
        memset(HII_l, 0, sizeof(HII_l));
        memset(gI_l, 0, sizeof(gI_l));
#pragma omp for schedule(static)
        for (int i = 0; i < n; ++i) {
            int np = (int)(off[i + 1] - off[i]);
            double *proj = (double *)malloc(sizeof(double) * 2 * (size_t)np);
            double *j6 = (double *)malloc(sizeof(double) * 12 * (size_t)np);
            double *j10 = (double *)malloc(sizeof(double) * 20 * (size_t)np);
            project_points(&c, np, obj + 3 * off[i], param + 6 * i, param + 6 * i + 3, proj, j6, j10);
            double *H = Hii + 36 * (size_t)i, *X = HiI + 60 * (size_t)i, *g = gi + 6 * (size_t)i;
            memset(H, 0, sizeof(double) * 36);
            memset(X, 0, sizeof(double) * 60);
            memset(g, 0, sizeof(double) * 6);
            for (int p = 0; p < np; ++p)
                for (int rr = 0; rr < 2; ++rr) {
                    const double *je = j6 + 12 * p + 6 * rr, *ji = j10 + 20 * p + 10 * rr;
                    double err = img[2 * (off[i] + p) + rr] - proj[2 * p + rr];
                    cost_l += err * err;
                    for (int a = 0; a < 6; ++a) {
                        for (int b = 0; b < 6; ++b) H[a * 6 + b] += je[a] * je[b];
                        for (int b = 0; b < 10; ++b) X[a * 10 + b] += je[a] * ji[b];
                        g[a] += je[a] * err;
                    }
                    for (int a = 0; a < 10; ++a) {
                        for (int b = 0; b < 10; ++b) HII_l[a * 10 + b] += ji[a] * ji[b];
                        gI_l[a] += ji[a] * err;
                    }
                }
            free(proj); free(j6); free(j10);
        }
#pragma omp critical
        {
            for (int a = 0; a < 100; ++a) HII[a] += HII_l[a];
            for (int a = 0; a < 10; ++a) gI[a] += gI_l[a];
            cost += cost_l;
        }
    }
    return cost;
}

/* Solve (JTJ_sub + eps 11^T) G = JTE_sub for the arrow-structured JTJ (src/omnidir.cpp:918-934, 1135-1137).
 * dense != 0: form the dense (6n+10)^2 matrix and invert it literally (toy sizes only);
 * dense == 0: Schur complement on the intrinsic block, the rank-one eps 11^T term carried as one extra bordered unknown.
 * G is returned zero-filled at fixed parameters (fillFixed). */
static int omni_solve_step(int n, const double *Hii, const double *HiI, const double *HII, const double *gi,
                           const double *gI, int flags, double epsilon, int dense, double *G)
{
    int P = 6 * n + 10;
    int *idx = (int *)malloc(sizeof(int) * (size_t)P);
    orc_omni_flags2idx(flags, n, idx);
    int map[10], m = 0; /* free intrinsic columns */
    for (int a = 0; a < 10; ++a)
        if (idx[6 * n + a]) map[m++] = a;
    memset(G, 0, sizeof(double) * (size_t)P);
    int rc = 0;
    if (dense) {
        int Q = 6 * n + m;
        double *A = (double *)calloc((size_t)Q * Q, sizeof(double));
        double *b = (double *)calloc((size_t)Q, sizeof(double));
        for (int i = 0; i < n; ++i) {
            for (int a = 0; a < 6; ++a) {
                for (int bb = 0; bb < 6; ++bb) A[(6 * i + a) * Q + 6 * i + bb] = Hii[36 * i + a * 6 + bb];
                for (int bb = 0; bb < m; ++bb) {
                    A[(6 * i + a) * Q + 6 * n + bb] = HiI[60 * i + a * 10 + map[bb]];
                    A[(6 * n + bb) * Q + 6 * i + a] = HiI[60 * i + a * 10 + map[bb]];
                }
                b[6 * i + a] = gi[6 * i + a];
            }
        }
        for (int a = 0; a < m; ++a) {
            for (int bb = 0; bb < m; ++bb) A[(6 * n + a) * Q + 6 * n + bb] = HII[map[a] * 10 + map[bb]];
            b[6 * n + a] = gI[map[a]];
        }
        for (size_t i = 0; i < (size_t)Q * Q; ++i) A[i] += epsilon; /* :934 scalar added to every element */
        /* Gaussian elimination with partial pivoting (the reference uses cv::Mat::inv = LU) */
        for (int k = 0; k < Q && !rc; ++k) {
            int piv = k;
            for (int i = k + 1; i < Q; ++i)
                if (fabs(A[i * Q + k]) > fabs(A[piv * Q + k])) piv = i;
            if (A[piv * Q + k] == 0) { rc = 1; break; }
            if (piv != k) {
                for (int j = 0; j < Q; ++j) { double t = A[k * Q + j]; A[k * Q + j] = A[piv * Q + j]; A[piv * Q + j] = t; }
                double t = b[k]; b[k] = b[piv]; b[piv] = t;
            }
            for (int i = k + 1; i < Q; ++i) {
                double f = A[i * Q + k] / A[k * Q + k];
                if (f == 0) continue;
                for (int j = k; j < Q; ++j) A[i * Q + j] -= f * A[k * Q + j];
                b[i] -= f * b[k];
            }
        }
        if (!rc) {
            for (int i = Q - 1; i >= 0; --i) {
                double s = b[i];
                for (int j = i + 1; j < Q; ++j) s -= A[i * Q + j] * b[j];
                b[i] = s / A[i * Q + i];
            }
            for (int i = 0; i < 6 * n; ++i) G[i] = b[i];
            for (int a = 0; a < m; ++a) G[6 * n + map[a]] = b[6 * n + a];
        }
        free(A); free(b);
    } else {
        /* JTJ + eps 11^T is dense, but [H, sqrt(eps) u; sqrt(eps) u^T, -1] [x; t] = [g; 0] (t = sqrt(eps) u^T x) is an
         * arrow matrix again with ONE more shared unknown.  Eliminating the per-frame 6x6 blocks leaves a symmetric
         * indefinite (m+1) x (m+1) system, solved by Gaussian elimination with partial pivoting.  Unlike
         * Sherman-Morrison this does not need H itself to be invertible (it is singular when focal length and xi
         * are not separately observable -- the case the reference's "in case JTJ is singular" comment is about). */
        double S[121], rg[11], ru[10], dsum = 0, csum = 0;
        const double se = sqrt(epsilon);
        for (int a = 0; a < m; ++a) {
            for (int b = 0; b < m; ++b) S[a * m + b] = HII[map[a] * 10 + map[b]];
            rg[a] = gI[map[a]];
            ru[a] = 1;
        }
        double *L = (double *)malloc(sizeof(double) * 36 * (size_t)n);
        double *zg = (double *)malloc(sizeof(double) * 6 * (size_t)n), *zu = (double *)malloc(sizeof(double) * 6 * (size_t)n);
        double *Y = (double *)malloc(sizeof(double) * 60 * (size_t)n); /* L^-1 H_pI, 6 x m */
        for (int i = 0; i < n && !rc; ++i) {
            double *Li = L + 36 * (size_t)i;
            memcpy(Li, Hii + 36 * (size_t)i, sizeof(double) * 36);
            if (chol_lower(Li, 6)) { rc = 1000 + i; break; }
            double *g6 = zg + 6 * (size_t)i, *u6 = zu + 6 * (size_t)i;
            for (int a = 0; a < 6; ++a) { g6[a] = gi[6 * (size_t)i + a]; u6[a] = 1; }
            for (int a = 0; a < 6; ++a) { /* forward substitution only: z = L^-1 b */
                double sg = g6[a], su = u6[a];
                for (int k = 0; k < a; ++k) { sg -= Li[a * 6 + k] * g6[k]; su -= Li[a * 6 + k] * u6[k]; }
                g6[a] = sg / Li[a * 6 + a]; u6[a] = su / Li[a * 6 + a];
            }
            double *Yi = Y + 60 * (size_t)i;
            for (int b = 0; b < m; ++b) {
                double col[6];
                for (int a = 0; a < 6; ++a) col[a] = HiI[60 * (size_t)i + a * 10 + map[b]];
                for (int a = 0; a < 6; ++a) {
                    double sv = col[a];
                    for (int k = 0; k < a; ++k) sv -= Li[a * 6 + k] * col[k];
                    col[a] = sv / Li[a * 6 + a];
                }
                for (int a = 0; a < 6; ++a) Yi[a * 10 + b] = col[a];
            }
            for (int a = 0; a < m; ++a) {
                double sg = 0, su = 0;
                for (int k = 0; k < 6; ++k) { sg += Yi[k * 10 + a] * g6[k]; su += Yi[k * 10 + a] * u6[k]; }
                rg[a] -= sg; ru[a] -= su;
                for (int b = 0; b < m; ++b) {
                    double sv = 0;
                    for (int k = 0; k < 6; ++k) sv += Yi[k * 10 + a] * Yi[k * 10 + b];
                    S[a * m + b] -= sv;
                }
            }
            for (int k = 0; k < 6; ++k) { dsum += u6[k] * u6[k]; csum += u6[k] * g6[k]; }
        }
        if (!rc) {
            const int Q = m + 1;
            double B[144], rhs[12];
            for (int a = 0; a < m; ++a) {
                for (int b = 0; b < m; ++b) B[a * Q + b] = S[a * m + b];
                B[a * Q + m] = se * ru[a];
                B[m * Q + a] = se * ru[a];
                rhs[a] = rg[a];
            }
            B[m * Q + m] = -(1 + epsilon * dsum);
            rhs[m] = -se * csum;
            for (int k = 0; k < Q && !rc; ++k) {
                int piv = k;
                for (int i = k + 1; i < Q; ++i)
                    if (fabs(B[i * Q + k]) > fabs(B[piv * Q + k])) piv = i;
                if (B[piv * Q + k] == 0) { rc = 2; break; }
                if (piv != k) {
                    for (int j = 0; j < Q; ++j) { double t = B[k * Q + j]; B[k * Q + j] = B[piv * Q + j]; B[piv * Q + j] = t; }
                    double t = rhs[k]; rhs[k] = rhs[piv]; rhs[piv] = t;
                }
                for (int i = k + 1; i < Q; ++i) {
                    double f = B[i * Q + k] / B[k * Q + k];
                    for (int j = k; j < Q; ++j) B[i * Q + j] -= f * B[k * Q + j];
                    rhs[i] -= f * rhs[k];
                }
            }
            if (!rc) {
                for (int i = Q - 1; i >= 0; --i) {
                    double sv = rhs[i];
                    for (int j = i + 1; j < Q; ++j) sv -= B[i * Q + j] * rhs[j];
                    rhs[i] = sv / B[i * Q + i];
                }
                const double t = rhs[m];
                for (int b = 0; b < m; ++b) G[6 * n + map[b]] = rhs[b];
                for (int i = 0; i < n; ++i) { /* x_p = L^-T (z_g - Y x_I - sqrt(eps) t z_u) */
                    const double *Li = L + 36 * (size_t)i, *Yi = Y + 60 * (size_t)i;
                    double r6[6];
                    for (int a = 0; a < 6; ++a) {
                        double sv = zg[6 * (size_t)i + a] - se * t * zu[6 * (size_t)i + a];
                        for (int b = 0; b < m; ++b) sv -= Yi[a * 10 + b] * rhs[b];
                        r6[a] = sv;
                    }
                    for (int a = 5; a >= 0; --a) {
                        double sv = r6[a];
                        for (int k = a + 1; k < 6; ++k) sv -= Li[k * 6 + a] * r6[k];
                        r6[a] = sv / Li[a * 6 + a];
                    }
                    for (int a = 0; a < 6; ++a) G[6 * i + a] = r6[a];
                }
            }
        }
        free(L); free(zg); free(zu); free(Y);
    }
    free(idx);
    return rc;
}

/* src/omnidir.cpp:1119-1147.  param (6n+10, in/out): [om_i,T_i]*n, fx, fy, s, cx, cy, xi, k1, k2, p1, p2.
 * obj (3 per point) / img (2 per point) are double like the reference's CV_64F copies (:1084-1093).
 * trace: per iteration [cost before, change, alpha, epsilon].  report: [iters, change, final sum sq, rms]. */
int orc_omni_solve(int n, const int64_t *off, const double *obj, const double *img, double *param, int flags,
                   int crit_type, int max_count, double eps, int dense, double *report, double *trace, int trace_cap)
{
    int P = 6 * n + 10;
    double *Hii = (double *)malloc(sizeof(double) * 36 * (size_t)n);
    double *HiI = (double *)malloc(sizeof(double) * 60 * (size_t)n);
    double *gi = (double *)malloc(sizeof(double) * 6 * (size_t)n);
    double *G = (double *)malloc(sizeof(double) * (size_t)P);
    double HII[100], gI[10], change = 1;
    int iter = 0, rc = 0;
    for (;; ++iter) {
        if ((crit_type == 1 && iter >= max_count) || (crit_type == 2 && change <= eps) ||
            (crit_type == 3 && (change <= eps || iter >= max_count)))
            break;
        double alpha = 1 - pow(1 - 0.01, (double)iter + 1.0);
        double epsilon = 0.01 * pow(0.9, (double)iter / 10);
        double cost = omni_build(n, off, obj, img, param, Hii, HiI, HII, gi, gI);
        rc = omni_solve_step(n, Hii, HiI, HII, gi, gI, flags, epsilon, dense, G);
        if (rc) break;
        for (int i = 0; i < P; ++i) G[i] *= alpha;
        change = norm2(G, (size_t)P) / norm2(param, (size_t)P); /* :1141 uses the parameters before the update */
        for (int i = 0; i < P; ++i) param[i] += G[i];
        if (trace && iter < trace_cap) {
            trace[4 * iter] = cost; trace[4 * iter + 1] = change; trace[4 * iter + 2] = alpha; trace[4 * iter + 3] = epsilon;
        }
    }
    double cost = omni_build(n, off, obj, img, param, Hii, HiI, HII, gi, gI);
    if (report) {
        report[0] = iter; report[1] = change; report[2] = cost;
        report[3] = sqrt(cost / (double)off[n]); /* src/omnidir.cpp:1794-1802 */
    }
    free(Hii); free(HiI); free(gi); free(G);
    return rc;
}

/* arrow blocks of one build, for cross-checking the CUDA path */
double orc_omni_build(int n, const int64_t *off, const double *obj, const double *img, const double *param,
                      double *Hii, double *HiI, double *HII, double *gi, double *gI)
{
    return omni_build(n, off, obj, img, param, Hii, HiI, HII, gi, gI);
}
