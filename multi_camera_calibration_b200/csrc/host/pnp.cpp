// pnp.cpp -- initial pattern pose of one image: cv::solvePnP(objectPoints, imagePoints, K, dist, rvec, tvec) with the
// default SOLVEPNP_ITERATIVE method, which the reference calls per corner file (src/mymulticalib.cpp:203-211).
// Restated from the published algorithm (OpenCV calib3d, cvFindExtrinsicCameraParams2): undistort the image points,
// initialise the pose from the plane-to-image homography (planar targets) or the 12-parameter DLT (general point sets),
// then minimise the reprojection error in pixels over the 6 pose parameters.  The refinement here iterates to the
// optimum (OpenCV stops after 20 LM iterations or a step below FLT_EPSILON); both land on the same least-squares pose.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../mccba_math.cuh"
#include "host_impl.hpp"

namespace mccba {

namespace {
// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 12): A = V diag(w) V^T, columns of V
void jacobi_eig(int n, std::vector<double>& A, std::vector<double>& V, std::vector<double>& w)
{
    V.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) off += A[(size_t)p * n + q] * A[(size_t)p * n + q];
        if (off < 1e-300) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[(size_t)p * n + q];
                if (std::fabs(apq) < 1e-300) continue;
                const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
                const double tau = (aqq - app) / (2 * apq);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1 + tau * tau));
                const double c = 1 / std::sqrt(1 + t * t), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
                    A[(size_t)k * n + p] = c * akp - s * akq;
                    A[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
                    A[(size_t)p * n + k] = c * apk - s * aqk;
                    A[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                    V[(size_t)k * n + p] = c * vkp - s * vkq;
                    V[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    w.resize(n);
    for (int i = 0; i < n; ++i) w[i] = A[(size_t)i * n + i];
}
// nearest rotation to M (3x3 row-major): polar decomposition by Newton iteration M <- (M + M^-T) / 2
void orthonormalise(double* M)
{
    for (int it = 0; it < 50; ++it) {
        const double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
        if (std::fabs(det) < 1e-300) return;
        double inv_t[9] = {(M[4] * M[8] - M[5] * M[7]) / det, (M[5] * M[6] - M[3] * M[8]) / det, (M[3] * M[7] - M[4] * M[6]) / det,
                           (M[2] * M[7] - M[1] * M[8]) / det, (M[0] * M[8] - M[2] * M[6]) / det, (M[1] * M[6] - M[0] * M[7]) / det,
                           (M[1] * M[5] - M[2] * M[4]) / det, (M[2] * M[3] - M[0] * M[5]) / det, (M[0] * M[4] - M[1] * M[3]) / det};
        double d = 0;
        for (int i = 0; i < 9; ++i) {
            const double v = 0.5 * (M[i] + inv_t[i]);
            d = std::max(d, std::fabs(v - M[i]));
            M[i] = v;
        }
        if (d < 1e-15) break;
    }
}
void undistort(const CamParams& c, double u, double v, double* xy)
{
    const double x0 = (u - c.cx) / c.fx, y0 = (v - c.cy) / c.fy;
    double x = x0, y = y0;
    for (int it = 0; it < 20; ++it) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((c.k6 * r2 + c.k5) * r2 + c.k4) * r2) / (1 + ((c.k3 * r2 + c.k2) * r2 + c.k1) * r2);
        const double dx = 2 * c.p1 * x * y + c.p2 * (r2 + 2 * x * x), dy = c.p1 * (r2 + 2 * y * y) + 2 * c.p2 * x * y;
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
    }
    xy[0] = x; xy[1] = y;
}
}  // namespace

bool solve_pnp(int n, const double* obj, const double* img, const double* K5, const double* dist8, int ndist, double* rvec,
               double* tvec)
{
    if (n < 4) return false;
    CamParams cam;
    std::memset(&cam, 0, sizeof(cam));
    cam.model = kPinhole;
    cam.fx = K5[0]; cam.fy = K5[1]; cam.cx = K5[2]; cam.cy = K5[3];
    double k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < ndist && i < 8; ++i) k[i] = dist8[i];
    cam.k1 = k[0]; cam.k2 = k[1]; cam.p1 = k[2]; cam.p2 = k[3]; cam.k3 = k[4]; cam.k4 = k[5]; cam.k5 = k[6]; cam.k6 = k[7];
    cam.rational = k[5] != 0 || k[6] != 0 || k[7] != 0;
    std::vector<double> m(2 * (size_t)n);
    for (int i = 0; i < n; ++i) undistort(cam, img[2 * i], img[2 * i + 1], &m[2 * (size_t)i]);
    // centroid and principal axes of the object points
    double Mc[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) Mc[a] += obj[3 * i + a];
    for (int a = 0; a < 3; ++a) Mc[a] /= n;
    std::vector<double> C(9, 0.0), V, w;
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) C[a * 3 + b] += (obj[3 * i + a] - Mc[a]) * (obj[3 * i + b] - Mc[b]);
    jacobi_eig(3, C, V, w);
    int ord[3] = {0, 1, 2};
    std::sort(ord, ord + 3, [&](int a, int b) { return w[a] > w[b]; });
    double R[9], t[3];
    const bool planar = w[ord[2]] <= 1e-3 * w[ord[1]] || n < 6;
    if (planar) {
        // plane frame: rows = principal axes (largest first), right-handed
        double Rp[9];
        for (int r = 0; r < 3; ++r)
            for (int a = 0; a < 3; ++a) Rp[r * 3 + a] = V[(size_t)a * 3 + ord[r]];
        const double det = Rp[0] * (Rp[4] * Rp[8] - Rp[5] * Rp[7]) - Rp[1] * (Rp[3] * Rp[8] - Rp[5] * Rp[6]) + Rp[2] * (Rp[3] * Rp[7] - Rp[4] * Rp[6]);
        if (det < 0)
            for (int a = 0; a < 3; ++a) Rp[6 + a] = -Rp[6 + a];
        // homography (X', Y') -> (x, y), Hartley-normalised DLT
        std::vector<double> P(2 * (size_t)n);
        double sx = 0, sy = 0;
        for (int i = 0; i < n; ++i) {
            const double d[3] = {obj[3 * i] - Mc[0], obj[3 * i + 1] - Mc[1], obj[3 * i + 2] - Mc[2]};
            P[2 * i] = Rp[0] * d[0] + Rp[1] * d[1] + Rp[2] * d[2];
            P[2 * i + 1] = Rp[3] * d[0] + Rp[4] * d[1] + Rp[5] * d[2];
            sx += std::fabs(P[2 * i]); sy += std::fabs(P[2 * i + 1]);
        }
        const double sP = (sx + sy) / (2.0 * n) > 0 ? (2.0 * n) / (sx + sy) : 1.0;
        double mx = 0, my = 0, sm = 0;
        for (int i = 0; i < n; ++i) { mx += m[2 * i]; my += m[2 * i + 1]; }
        mx /= n; my /= n;
        for (int i = 0; i < n; ++i) sm += std::fabs(m[2 * i] - mx) + std::fabs(m[2 * i + 1] - my);
        const double sM = sm > 0 ? (2.0 * n) / sm : 1.0;
        std::vector<double> A(81, 0.0), VV, ww;
        for (int i = 0; i < n; ++i) {
            const double X = P[2 * i] * sP, Y = P[2 * i + 1] * sP, x = (m[2 * i] - mx) * sM, y = (m[2 * i + 1] - my) * sM;
            const double r1[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x}, r2[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
            for (int a = 0; a < 9; ++a)
                for (int b = 0; b < 9; ++b) A[a * 9 + b] += r1[a] * r1[b] + r2[a] * r2[b];
        }
        jacobi_eig(9, A, VV, ww);
        int best = 0;
        for (int i = 1; i < 9; ++i)
            if (ww[i] < ww[best]) best = i;
        double Hn[9], H[9];
        for (int i = 0; i < 9; ++i) Hn[i] = VV[(size_t)i * 9 + best];
        // denormalise: m = Tm^-1 Hn Tp P
        const double Tp[9] = {sP, 0, 0, 0, sP, 0, 0, 0, 1}, Tmi[9] = {1 / sM, 0, mx, 0, 1 / sM, my, 0, 0, 1};
        double tmp[9];
        mat3_mul(Hn, Tp, tmp);
        mat3_mul(Tmi, tmp, H);
        double h1[3] = {H[0], H[3], H[6]}, h2[3] = {H[1], H[4], H[7]}, h3[3] = {H[2], H[5], H[8]};
        const double n1 = std::sqrt(h1[0] * h1[0] + h1[1] * h1[1] + h1[2] * h1[2]), n2 = std::sqrt(h2[0] * h2[0] + h2[1] * h2[1] + h2[2] * h2[2]);
        if (n1 < 1e-300 || n2 < 1e-300) return false;
        double sc = 2.0 / (n1 + n2);
        if (h3[2] * sc < 0) sc = -sc;                       // the target is in front of the camera
        for (int a = 0; a < 3; ++a) { h1[a] *= (sc > 0 ? 1 : -1) / n1; h2[a] *= (sc > 0 ? 1 : -1) / n2; h3[a] *= sc; }
        double c3[3];
        cross3(h1, h2, c3);
        double Rh[9] = {h1[0], h2[0], c3[0], h1[1], h2[1], c3[1], h1[2], h2[2], c3[2]};
        orthonormalise(Rh);
        mat3_mul(Rh, Rp, R);                                 // Xc = Rh Rp (X - Mc) + th
        double RM[3];
        mat3_vec(R, Mc, RM);
        for (int a = 0; a < 3; ++a) t[a] = h3[a] - RM[a];
    } else {
        // DLT: [x y 1]^T ~ [R | t] [X 1]^T, 12 unknowns
        std::vector<double> A(144, 0.0), VV, ww;
        for (int i = 0; i < n; ++i) {
            const double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2], x = m[2 * i], y = m[2 * i + 1];
            const double r1[12] = {X, Y, Z, 1, 0, 0, 0, 0, -x * X, -x * Y, -x * Z, -x}, r2[12] = {0, 0, 0, 0, X, Y, Z, 1, -y * X, -y * Y, -y * Z, -y};
            for (int a = 0; a < 12; ++a)
                for (int b = 0; b < 12; ++b) A[a * 12 + b] += r1[a] * r1[b] + r2[a] * r2[b];
        }
        jacobi_eig(12, A, VV, ww);
        int best = 0;
        for (int i = 1; i < 12; ++i)
            if (ww[i] < ww[best]) best = i;
        double L[12];
        for (int i = 0; i < 12; ++i) L[i] = VV[(size_t)i * 12 + best];
        double RR[9] = {L[0], L[1], L[2], L[4], L[5], L[6], L[8], L[9], L[10]};
        double det = RR[0] * (RR[4] * RR[8] - RR[5] * RR[7]) - RR[1] * (RR[3] * RR[8] - RR[5] * RR[6]) + RR[2] * (RR[3] * RR[7] - RR[4] * RR[6]);
        const double sgn = det < 0 ? -1.0 : 1.0;
        const double sc = sgn / std::cbrt(std::fabs(det) > 1e-300 ? std::fabs(det) : 1.0);
        for (int i = 0; i < 9; ++i) RR[i] *= sc;
        t[0] = L[3] * sc; t[1] = L[7] * sc; t[2] = L[11] * sc;
        orthonormalise(RR);
        std::memcpy(R, RR, sizeof(R));
    }
    // refinement: Levenberg-Marquardt on the pixel reprojection error, left perturbation of R
    double lambda = 1e-3, cost = -1;
    auto eval = [&](const double* Rm, const double* tm, double* H, double* g) {
        double c = 0;
        if (H) { std::fill(H, H + 36, 0.0); std::fill(g, g + 6, 0.0); }
        for (int i = 0; i < n; ++i) {
            const double X[3] = {obj[3 * i], obj[3 * i + 1], obj[3 * i + 2]};
            double Q[3], Xc[3], uv[2], A[6];
            mat3_vec(Rm, X, Q);
            for (int a = 0; a < 3; ++a) Xc[a] = Q[a] + tm[a];
            if (cam.rational) pinhole_point<true, true>(cam, Xc, uv, A);
            else pinhole_point<false, true>(cam, Xc, uv, A);
            const double e0 = img[2 * i] - uv[0], e1 = img[2 * i + 1] - uv[1];
            c += e0 * e0 + e1 * e1;
            if (!H) continue;
            double j0[6], j1[6];
            cross3(Q, A, j0);
            cross3(Q, A + 3, j1);
            for (int a = 0; a < 3; ++a) { j0[3 + a] = A[a]; j1[3 + a] = A[3 + a]; }
            for (int a = 0; a < 6; ++a) {
                for (int b = 0; b < 6; ++b) H[a * 6 + b] += j0[a] * j0[b] + j1[a] * j1[b];
                g[a] += j0[a] * e0 + j1[a] * e1;
            }
        }
        return c;
    };
    for (int it = 0; it < 100; ++it) {
        double H[36], g[6];
        cost = eval(R, t, H, g);
        bool improved = false;
        for (int tries = 0; tries < 12 && !improved; ++tries) {
            double U[21], d[6];
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b) U[tri6(a, b)] = H[a * 6 + b] * (a == b ? 1.0 + lambda : 1.0);
            for (int a = 0; a < 6; ++a) d[a] = g[a];
            if (!chol6_packed(U)) { lambda *= 10; continue; }
            chol6_forward(U, d, 1);
            chol6_backward(U, d);
            double dR[9], Rn[9], tn[3];
            rodrigues(d, dR);
            mat3_mul(dR, R, Rn);
            for (int a = 0; a < 3; ++a) tn[a] = t[a] + d[3 + a];
            const double cn = eval(Rn, tn, nullptr, nullptr);
            if (cn <= cost) {
                const double step = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) + std::sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]) / (1.0 + std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]));
                std::memcpy(R, Rn, sizeof(R));
                std::memcpy(t, tn, sizeof(t));
                lambda = std::max(lambda / 10, 1e-12);
                improved = true;
                if (step < 1e-13 || cost - cn <= 1e-16 * cost) it = 1000;
                cost = cn;
            } else lambda *= 10;
        }
        if (!improved) break;
    }
    log_so3_3x3(R, rvec);
    tvec[0] = t[0]; tvec[1] = t[1]; tvec[2] = t[2];
    return std::isfinite(cost);
}

}  // namespace mccba
