// mccba_stereo.cuh -- cv::omnidir::stereoCalibrate's optimisation loop and estimateUncertaintiesStereo on the device
// (SURVEY.md 8(f) row 3).  Replaces src/omnidir.cpp:937-1020 (computeJacobianStereo: dense J of 4 N n x (6 (n + 1) + 20),
// J^T J, (J^T J + eps)^-1), :1268-1296 (loop), :2078-2136 / :2155-2170 (flag masks) and :1804-1889 (uncertainties).
//
// Parameter vector (encodeParametersStereo, :1570-1620):
//   [om, T: camera 2 relative to camera 1 | om_i, T_i of the n frames in camera 1 | 10 intrinsics of camera 1 | of camera 2]
// Structure: every frame touches its own 6 pose parameters and the 26 shared ones (relative pose + 2 x 10 intrinsics),
// so -- exactly as in the single-camera loop (mccba_omni.cuh) -- the per-frame 6 x 6 blocks are eliminated, the reduced
// system is 26 x 26, and the reference's "+ eps on EVERY element" (a rank-one eps 11^T) is carried as one bordered
// unknown.  Additive Rodrigues coordinates, because that term is not invariant under a change of coordinates.
//   stereo_frame_kernel   warp per frame: rows [d/d(om_i,T_i) (6) | d/d(om,T) (6) | intrinsics 1 (10) | intrinsics 2 (10) | e] of
//                         the left and right image of every corner -> 33 x 33 Gram matrix; 6 x 6 Cholesky, Y = L^-1 H_pS,
//                         Schur record (406 numbers), residual moments for the uncertainties
//   stereo_reduce_kernel  fixed-order sums of the record entries over the frames (no atomics)
//   stereo_solve_kernel   bordered (m + 1) x (m + 1) system, m <= 26 free shared parameters, partial pivoting
//   stereo_update_kernel  back-substitution per frame, G = alpha x, parameter update, norm partials
//   stereo_decide_kernel  change = |G| / |param_old|, termination (:1271-1274)
//   stereo_cov_kernel(s)  diag((J^T J)^-1): inverse of the reduced system, then per frame L^-T (I + Y S^-1 Y^T) L^-1
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mccba_math.cuh"
#include "mccba_warp_solve.cuh"

namespace mccba {

constexpr int kStNS = 26;                  // shared parameters: 6 (relative pose) + 10 + 10
constexpr int kStW = 6 + kStNS + 1;        // row width: frame pose | shared | residual
constexpr int kStTri = kStW * (kStW + 1) / 2;   // 561
constexpr int kStThreads = 128;                // 4 warps = 4 frames per CTA
constexpr int kStSTri = kStNS * (kStNS + 1) / 2;   // 351
constexpr int kStRec = kStSTri + 2 * kStNS + 3 + 4;   // S | rg | ru | d, c, cost | sum ex, ey, ex^2, ey^2  = 410
constexpr int kStSave = 21 + 6 + 6 + 6 * kStNS;       // U | z_g | z_u | Y = 189

struct StereoState {
    int flags, crit_type, max_count, iter, done, status;
    double eps_crit, change, alpha, epsilon;
    double x_sh[kStNS];  // solution of the shared block (0 at fixed parameters)
    double t;            // bordered unknown sqrt(eps) 1^T x
};

struct StereoProblem {
    int n_frame, n_blocks_upd;
    int64_t n_pts;                 // corners (each seen by both cameras)
    const float *ox, *oy, *oz;     // object points, per frame ranges f_off
    const float *u1, *v1, *u2, *v2;
    const int* f_off;
    double* param;                 // 6 (n + 1) + 20
    double* rec;                   // kStRec x n_frame (SoA)
    double* save;                  // kStSave x n_frame (SoA)
    double* tot;                   // kStRec
    double* norm_part;             // 2 x n_blocks_upd
    double* sinv;                  // 26 x 26: inverse of the free part of the reduced system (uncertainties)
    double* diag;                  // 6 (n + 1) + 20: diag((J^T J)^-1), 0 at fixed parameters
    double* dump;                  // optional: n_frame x 33 x 33 Gram matrices (tests)
    StereoState* st;
};

__device__ __forceinline__ void stereo_tri(int t, int& i, int& j)   // t in [0, 561) -> (i <= j) of the 33 x 33 upper triangle
{
    int r = 0, rem = t;
    while (rem >= kStW - r) { rem -= kStW - r; ++r; }
    i = r; j = r + rem;
}
__device__ __forceinline__ int stereo_sidx(int a, int c)            // upper-triangle index of (a, c) in the 26 x 26 block
{
    if (a > c) { const int t = a; a = c; c = t; }
    return a * kStNS - (a * (a - 1)) / 2 + (c - a);
}
__device__ __forceinline__ void stereo_cam(const double* p, CamParams& cam)
{
    cam.model = kOmnidir; cam.rational = 0;
    cam.fx = p[0]; cam.fy = p[1]; cam.skew = p[2]; cam.cx = p[3]; cam.cy = p[4]; cam.xi = p[5];
    cam.k1 = p[6]; cam.k2 = p[7]; cam.p1 = p[8]; cam.p2 = p[9];
    cam.k3 = cam.k4 = cam.k5 = cam.k6 = 0.0;
}

// One WARP per frame (the layout of omni_frame_kernel): per pass of 16 corners lanes 0..15 evaluate the left image and
// lanes 16..31 the right image of the same corners, two 33-wide rows each, into the warp's shared-memory stage (64 rows);
// the 33 x 33 Gram matrix sum a^T a is accumulated by the FP64 tensor cores: 15 upper 8 x 8 tile pairs of the five
// 8-column slices, 16 k-steps of mma.m8n8k4 per pass (5 fragment loads for 15 MMAs).  The per-frame algebra (6 x 6
// Cholesky, 28 forward solves, the 410-entry Schur record) follows in the same warp.
constexpr int kStWarps = 4;                 // frames per CTA
constexpr int kStLd = 36;                   // row stride (doubles): 8 x 4 fragment loads in the minimal 2 wavefronts
constexpr int kStRows = 64;                 // rows staged per pass = 16 corners x (2 left + 2 right)
__global__ void __launch_bounds__(kStThreads) stereo_frame_kernel(StereoProblem P, int forced)
{
    extern __shared__ __align__(16) unsigned char st_smem[];
    double* s_rows_all = reinterpret_cast<double*>(st_smem);                        // [kStWarps][kStRows * kStLd]
    double* s_fac_all = s_rows_all + kStWarps * kStRows * kStLd;                    // [kStWarps][192]: U 21 | z_g 6 | z_u 6 | Y 6 x 26
    const StereoState* st = P.st;
    if (!forced && st->done) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, n = P.n_frame;
    const int f = blockIdx.x * kStWarps + w;
    if (f >= n) return;                          // whole warps leave; nothing below synchronises across warps
    double* rows = s_rows_all + w * kStRows * kStLd;
    double* sU = s_fac_all + w * 192;
    double* szg = sU + 21;
    double* szu = sU + 27;
    double (*sY)[kStNS] = reinterpret_cast<double (*)[kStNS]>(sU + 33);
    const double* par = P.param;
    const int o1 = 6 * (n + 1), o2 = o1 + 10;
    const bool right = lane >= 16;
    CamParams cam;
    stereo_cam(par + (right ? o2 : o1), cam);
    const double om[3] = {par[0], par[1], par[2]}, T[3] = {par[3], par[4], par[5]};
    const double om1[3] = {par[6 + 6 * f], par[7 + 6 * f], par[8 + 6 * f]}, T1[3] = {par[9 + 6 * f], par[10 + 6 * f], par[11 + 6 * f]};
    double R[9], R1[9], R2[9], Jl[9], Jl1[9], s[3], T2[3];
    rodrigues(om, R);
    rodrigues(om1, R1);
    left_jacobian(om, Jl);
    left_jacobian(om1, Jl1);
    mat3_mul(R, R1, R2);                     // compose_motion(om1, T1, om, T): R2 = R R1, T2 = R T1 + T (:989)
    mat3_vec(R, T1, s);
    T2[0] = s[0] + T[0]; T2[1] = s[1] + T[1]; T2[2] = s[2] + T[2];
    const int g = lane >> 2, t = lane & 3;
    double acc[15][2];
#pragma unroll
    for (int i = 0; i < 15; ++i) acc[i][0] = acc[i][1] = 0.0;
    double mom[4] = {0, 0, 0, 0};
    const int b = P.f_off[f], e = P.f_off[f + 1];
    for (int c0 = b; c0 < e; c0 += 16) {
        const int nc = min(16, e - c0);
        const int cl = lane & 15;                // corner of the pass this lane evaluates
        __syncwarp();                            // the previous pass has been consumed
        double* row0 = rows + (4 * cl + (right ? 2 : 0)) * kStLd;
        double* row1 = row0 + kStLd;
        if (cl < nc) {
            const int i = c0 + cl;
            const double X[3] = {(double)P.ox[i], (double)P.oy[i], (double)P.oz[i]};
            double Q[3], Xc[3], uv[2], A[6], Jin[20];
            if (!right) {                        // left image: camera 1 at the frame's pose
                mat3_vec(R1, X, Q);
                Xc[0] = Q[0] + T1[0]; Xc[1] = Q[1] + T1[1]; Xc[2] = Q[2] + T1[2];
            } else {                             // right image: camera 2 at the composed pose
                mat3_vec(R2, X, Q);
                Xc[0] = Q[0] + T2[0]; Xc[1] = Q[1] + T2[1]; Xc[2] = Q[2] + T2[2];
            }
            omnidir_point_full(cam, Xc, uv, A, Jin);
            const double err[2] = {(double)(right ? P.u2[i] : P.u1[i]) - uv[0], (double)(right ? P.v2[i] : P.v1[i]) - uv[1]};
            mom[0] += err[0]; mom[1] += err[1]; mom[2] += err[0] * err[0]; mom[3] += err[1] * err[1];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                double* row = r == 0 ? row0 : row1;
                const double* a = A + 3 * r;
                double jphi[3];
                cross3(Q, a, jphi);
                if (!right) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) row[k] = jphi[0] * Jl1[k] + jphi[1] * Jl1[3 + k] + jphi[2] * Jl1[6 + k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) row[3 + k] = a[k];
#pragma unroll
                    for (int k = 0; k < 6; ++k) row[6 + k] = 0.0;
#pragma unroll
                    for (int k = 0; k < 10; ++k) { row[12 + k] = Jin[10 * r + k]; row[22 + k] = 0.0; }
                } else {
                    // left perturbations: frame psi_1 -> phi_2 = R psi_1, tau_2 = R dT_1; relative pose psi -> phi_2 = psi,
                    // tau_2 = psi x (R T_1) + dT
                    double jf[3], jt[3], sc[3], jr[3];
                    mat3t_vec(R, jphi, jf);                   // d/d psi_1 = R^T jphi
                    mat3t_vec(R, a, jt);                      // d/d dT_1  = R^T a
                    cross3(s, a, sc);                         // a . (psi x s) = psi . (s x a)
                    jr[0] = jphi[0] + sc[0]; jr[1] = jphi[1] + sc[1]; jr[2] = jphi[2] + sc[2];
#pragma unroll
                    for (int k = 0; k < 3; ++k) row[k] = jf[0] * Jl1[k] + jf[1] * Jl1[3 + k] + jf[2] * Jl1[6 + k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) row[3 + k] = jt[k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) row[6 + k] = jr[0] * Jl[k] + jr[1] * Jl[3 + k] + jr[2] * Jl[6 + k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) row[9 + k] = a[k];
#pragma unroll
                    for (int k = 0; k < 10; ++k) { row[12 + k] = 0.0; row[22 + k] = Jin[10 * r + k]; }
                }
                row[32] = err[r];
            }
        } else {
#pragma unroll
            for (int k = 0; k < kStW; ++k) { row0[k] = 0.0; row1[k] = 0.0; }     // rows past the last corner add zeros
        }
        __syncwarp();
        const int ksteps = nc;                   // 4 rows per corner = one k-step per corner
        for (int ks = 0; ks < ksteps; ++ks) {
            const double* r = rows + (4 * ks + t) * kStLd + g;
            double fr[5];
            fr[0] = r[0]; fr[1] = r[8]; fr[2] = r[16]; fr[3] = r[24]; fr[4] = g == 0 ? r[32] : 0.0;
            int q = 0;
#pragma unroll
            for (int a = 0; a < 5; ++a)
#pragma unroll
                for (int c = a; c < 5; ++c) { dmma(acc[q][0], acc[q][1], fr[a], fr[c]); ++q; }
        }
    }
    __syncwarp();
    double (*M)[kStW] = reinterpret_cast<double (*)[kStW]>(rows);   // the Gram matrix takes the place of the staged rows
    {
        int q = 0;
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int c = a; c < 5; ++c) {
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int r = 8 * a + g, cc = 8 * c + 2 * t + h2;
                    if (r < kStW && cc < kStW && r <= cc) { M[r][cc] = acc[q][h2]; M[cc][r] = acc[q][h2]; }
                }
                ++q;
            }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mom[k] += __shfl_xor_sync(0xffffffffu, mom[k], o);
    }
    __syncwarp();
    if (P.dump)
        for (int k = lane; k < kStW * kStW; k += 32) P.dump[(int64_t)f * kStW * kStW + k] = M[k / kStW][k % kStW];
    int bad = 0;
    if (lane == 0) {
        double U[21];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) U[tri6(i, j)] = M[i][j];
        if (!chol6_packed(U)) bad = 1;
#pragma unroll
        for (int i = 0; i < 21; ++i) sU[i] = U[i];
    }
    __syncwarp();
    if (lane < kStNS + 2) {                      // Y columns (26), z_g, z_u
        double U[21], col[6];
#pragma unroll
        for (int i = 0; i < 21; ++i) U[i] = sU[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) col[i] = lane < kStNS ? M[i][6 + lane] : (lane == kStNS ? M[i][32] : 1.0);
        chol6_forward(U, col, 1);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (lane < kStNS) sY[i][lane] = col[i];
            else if (lane == kStNS) szg[i] = col[i];
            else szu[i] = col[i];
        }
    }
    __syncwarp();
    const int64_t nf = n;
    for (int tt = lane; tt < kStRec; tt += 32) {
        double v = 0.0;
        if (tt < kStSTri) {                      // S: upper triangle of H_SS - Y^T Y
            int a = 0, rem = tt;
            while (rem >= kStNS - a) { rem -= kStNS - a; ++a; }
            const int c = a + rem;
            v = M[6 + a][6 + c];
#pragma unroll
            for (int k = 0; k < 6; ++k) v -= sY[k][a] * sY[k][c];
        } else if (tt < kStSTri + kStNS) {       // r_g
            const int a = tt - kStSTri;
            v = M[6 + a][32];
#pragma unroll
            for (int k = 0; k < 6; ++k) v -= sY[k][a] * szg[k];
        } else if (tt < kStSTri + 2 * kStNS) {   // r_u
            const int a = tt - kStSTri - kStNS;
#pragma unroll
            for (int k = 0; k < 6; ++k) v -= sY[k][a] * szu[k];
        } else {
            const int q = tt - kStSTri - 2 * kStNS;
            if (q == 0) { for (int k = 0; k < 6; ++k) v += szu[k] * szu[k]; }
            else if (q == 1) { for (int k = 0; k < 6; ++k) v += szu[k] * szg[k]; }
            else if (q == 2) v = M[32][32];
            else v = mom[q - 3];
        }
        P.rec[(int64_t)tt * nf + f] = v;
    }
    for (int k = lane; k < kStSave; k += 32) P.save[(int64_t)k * nf + f] = sU[k];     // U | z_g | z_u | Y are contiguous
    if (lane == 0 && bad) P.st->status = 4;
}
constexpr int kStFrameSmem = (kStWarps * kStRows * kStLd + kStWarps * 192) * 8;

__global__ void __launch_bounds__(256) stereo_reduce_kernel(StereoProblem P, int forced)
{
    if (!forced && P.st->done) return;
    __shared__ double sm[8];
    const int v = blockIdx.x, n = P.n_frame;
    const double* src = P.rec + (int64_t)v * n;
    double s = 0.0;
    for (int f = threadIdx.x; f < n; f += blockDim.x) s += src[f];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        P.tot[v] = t;
    }
}

// shared-block index a in [0, 26) -> free? (flags2idxStereo, :2078-2136: the same >= / subtract cascade for both cameras)
__device__ __forceinline__ void stereo_flags2free(int flags, int* fr)
{
    for (int a = 0; a < kStNS; ++a) fr[a] = 1;
    int f = flags;
    auto fix = [&](int c) { fr[6 + c] = 0; fr[16 + c] = 0; };
    if (f >= 256) { fix(3); fix(4); f -= 256; }
    if (f >= 128) { fix(0); fix(1); f -= 128; }
    if (f >= 64) { fix(5); f -= 64; }
    if (f >= 32) { fix(9); f -= 32; }
    if (f >= 16) { fix(8); f -= 16; }
    if (f >= 8) { fix(7); f -= 8; }
    if (f >= 4) { fix(6); f -= 4; }
    if (f >= 2) { fix(2); }
}

// One warp, lane = row of the bordered (m + 1) x (m + 1) system (mccba_warp_solve.cuh); round 2 first ran this on one
// thread out of local memory, which was most of the iteration (368 us per iteration at 400 frames).
__global__ void __launch_bounds__(32) stereo_solve_kernel(StereoProblem P)
{
    constexpr int N = kStNS + 1;
    __shared__ double s_rows[2 * (N + 1)];
    __shared__ int s_map[kStNS];
    StereoState* st = P.st;
    // no early exit on the loaded state (the warp must provably stay converged): only the stores at the end are guarded
    const int done = st->done, status = st->status;
    const int lane = threadIdx.x;
    const double alpha = 1.0 - pow(1.0 - 0.01, (double)st->iter + 1.0);       // :1276
    const double epsilon = 0.01 * pow(0.9, (double)st->iter / 10.0);           // :1278
    int m = 0;
    {
        int fr[kStNS];
        stereo_flags2free(st->flags, fr);
        for (int a = 0; a < kStNS; ++a)
            if (fr[a]) { if (lane == 0) s_map[m] = a; ++m; }
    }
    __syncwarp();
    const double* tot = P.tot;
    const double se = sqrt(epsilon);
    const int Q = m + 1;
    const int mine = lane < m ? s_map[lane] : 0;
    double B[N], rhs = 0.0;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        double v = 0.0;
        if (lane < m) {
            if (c < m) v = tot[stereo_sidx(mine, s_map[c])];
            else if (c == m) v = se * (1.0 + tot[kStSTri + kStNS + mine]);
        } else if (lane == m) {
            if (c < m) v = se * (1.0 + tot[kStSTri + kStNS + s_map[c]]);
            else if (c == m) v = -(1.0 + epsilon * tot[kStSTri + 2 * kStNS]);
        } else if (lane == c) v = 1.0;          // identity padding of the rows and columns past the system
        B[c] = v;
    }
    if (lane < m) rhs = tot[kStSTri + mine];
    else if (lane == m) rhs = -se * tot[kStSTri + 2 * kStNS + 1];
    double x[N];
    const bool failed = warp_gauss_solve<N>(B, rhs, lane, Q, s_rows, x);
    if (lane == 0 && !done) {
        st->alpha = alpha; st->epsilon = epsilon;
        if (status) st->done = 1;
        else if (failed) { st->status = 4; st->done = 1; }
        else {
            for (int a = 0; a < kStNS; ++a) st->x_sh[a] = 0.0;
            double tv = 0.0;
#pragma unroll
            for (int a = 0; a < N; ++a) {
                if (a < m) st->x_sh[s_map[a]] = x[a];
                if (a == m) tv = x[a];
            }
            st->t = tv;
        }
    }
}

// shared index a -> position in the parameter vector
__device__ __forceinline__ int64_t stereo_param_of(int a, int n) { return a < 6 ? a : (int64_t)6 * (n + 1) + (a - 6); }

__global__ void __launch_bounds__(128) stereo_update_kernel(StereoProblem P)
{
    const StereoState* st = P.st;
    if (st->done) return;
    const int n = P.n_frame;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const double alpha = st->alpha, set = sqrt(st->epsilon) * st->t;
    double g2 = 0.0, p2 = 0.0;
    if (f < n) {
        const int64_t nf = n;
        double U[21], r[6];
#pragma unroll
        for (int k = 0; k < 21; ++k) U[k] = P.save[(int64_t)k * nf + f];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double s = P.save[(int64_t)(21 + i) * nf + f] - set * P.save[(int64_t)(27 + i) * nf + f];
            for (int a = 0; a < kStNS; ++a) s -= P.save[(int64_t)(33 + i * kStNS + a) * nf + f] * st->x_sh[a];
            r[i] = s;
        }
        chol6_backward(U, r);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double old = P.param[6 + 6 * (int64_t)f + i], G = alpha * r[i];
            g2 += G * G; p2 += old * old;
            P.param[6 + 6 * (int64_t)f + i] = old + G;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < kStNS) {   // shared block (fixed parameters get G = 0, fillFixedStereo)
        const int a = threadIdx.x;
        const int64_t pi = stereo_param_of(a, n);
        const double old = P.param[pi], G = alpha * st->x_sh[a];
        g2 += G * G; p2 += old * old;
        P.param[pi] = old + G;
    }
    __shared__ double sr[2][4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { g2 += __shfl_xor_sync(0xffffffffu, g2, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o); }
    if ((threadIdx.x & 31) == 0) { sr[0][threadIdx.x >> 5] = g2; sr[1][threadIdx.x >> 5] = p2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        P.norm_part[blockIdx.x] = sr[0][0] + sr[0][1] + sr[0][2] + sr[0][3];
        P.norm_part[P.n_blocks_upd + blockIdx.x] = sr[1][0] + sr[1][1] + sr[1][2] + sr[1][3];
    }
}

__global__ void __launch_bounds__(256) stereo_decide_kernel(StereoProblem P)
{
    StereoState* st = P.st;
    if (st->done) return;
    __shared__ double sm[2][8];
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < P.n_blocks_upd; k += blockDim.x) { a += P.norm_part[k]; b += P.norm_part[P.n_blocks_upd + k]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = a; sm[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double g2 = 0, p2 = 0;
        for (int w = 0; w < 8; ++w) { g2 += sm[0][w]; p2 += sm[1][w]; }
        st->change = sqrt(g2) / sqrt(p2);   // :1288, the norm of the parameters BEFORE the update
        st->iter += 1;
        const int t = st->crit_type;
        if ((t == 1 && st->iter >= st->max_count) || (t == 2 && st->change <= st->eps_crit) ||
            (t == 3 && (st->change <= st->eps_crit || st->iter >= st->max_count)))
            st->done = 1;
    }
}

__global__ void stereo_init_state_kernel(StereoState* st, int flags, int crit_type, int max_count, double eps)
{
    st->flags = flags; st->crit_type = crit_type; st->max_count = max_count; st->iter = 0; st->status = 0;
    st->eps_crit = eps; st->change = 1.0; st->alpha = 0; st->epsilon = 0; st->t = 0;
    for (int a = 0; a < kStNS; ++a) st->x_sh[a] = 0;
    st->done = (crit_type == 1 && 0 >= max_count) || (crit_type == 3 && 0 >= max_count) || (crit_type == 2 && 1.0 <= eps) ? 1 : 0;
}

// ---- uncertainties: diag((J^T J)^-1) through the Schur factor (estimateUncertaintiesStereo, :1873-1877, eps = 0) -------
// step 1 (one thread): invert the free part of the reduced system (<= 26 x 26, Gauss-Jordan with partial pivoting) into
// P.sinv (26 x 26, zero rows / columns at fixed parameters) and write the shared entries of the diagonal
__global__ void stereo_cov_shared_kernel(StereoProblem P, int flags)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int fr[kStNS], map[kStNS], m = 0;
    stereo_flags2free(flags, fr);
    for (int a = 0; a < kStNS; ++a)
        if (fr[a]) map[m++] = a;
    double A[kStNS * kStNS], I[kStNS * kStNS];
    for (int a = 0; a < m; ++a)
        for (int c = 0; c < m; ++c) { A[a * m + c] = P.tot[stereo_sidx(map[a], map[c])]; I[a * m + c] = a == c ? 1.0 : 0.0; }
    int fail = 0;
    for (int k = 0; k < m && !fail; ++k) {
        int piv = k;
        for (int i = k + 1; i < m; ++i)
            if (fabs(A[i * m + k]) > fabs(A[piv * m + k])) piv = i;
        if (A[piv * m + k] == 0.0 || !isfinite(A[piv * m + k])) { fail = 1; break; }
        if (piv != k)
            for (int j = 0; j < m; ++j) {
                double t = A[k * m + j]; A[k * m + j] = A[piv * m + j]; A[piv * m + j] = t;
                t = I[k * m + j]; I[k * m + j] = I[piv * m + j]; I[piv * m + j] = t;
            }
        const double d = 1.0 / A[k * m + k];
        for (int j = 0; j < m; ++j) { A[k * m + j] *= d; I[k * m + j] *= d; }
        for (int i = 0; i < m; ++i) {
            if (i == k) continue;
            const double fct = A[i * m + k];
            for (int j = 0; j < m; ++j) { A[i * m + j] -= fct * A[k * m + j]; I[i * m + j] -= fct * I[k * m + j]; }
        }
    }
    if (fail) { P.st->status = 4; return; }
    for (int a = 0; a < kStNS * kStNS; ++a) P.sinv[a] = 0.0;
    for (int a = 0; a < m; ++a)
        for (int c = 0; c < m; ++c) P.sinv[map[a] * kStNS + map[c]] = I[a * m + c];
    for (int a = 0; a < kStNS; ++a) P.diag[stereo_param_of(a, P.n_frame)] = fr[a] ? P.sinv[a * kStNS + a] : 0.0;
}
// step 2 (one thread per frame): the frame's diagonal of  H_pp^-1 + H_pp^-1 W S^-1 W^T H_pp^-1 = L^-T (I + Y S^-1 Y^T) L^-1
__global__ void __launch_bounds__(128) stereo_cov_frame_kernel(StereoProblem P)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x, n = P.n_frame;
    if (f >= n) return;
    const int64_t nf = n;
    double U[21], C[36];
#pragma unroll
    for (int k = 0; k < 21; ++k) U[k] = P.save[(int64_t)k * nf + f];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) C[i * 6 + j] = i == j ? 1.0 : 0.0;
    for (int a = 0; a < kStNS; ++a) {
        double t[6] = {0, 0, 0, 0, 0, 0};     // t = Y S^-1 (:, a)
        for (int c = 0; c < kStNS; ++c) {
            const double sv = P.sinv[c * kStNS + a];
            if (sv == 0.0) continue;
            for (int i = 0; i < 6; ++i) t[i] += P.save[(int64_t)(33 + i * kStNS + c) * nf + f] * sv;
        }
        for (int i = 0; i < 6; ++i) {
            const double ti = t[i];
            for (int j = 0; j < 6; ++j) C[i * 6 + j] += ti * P.save[(int64_t)(33 + j * kStNS + a) * nf + f];
        }
    }
    // (L^-T C L^-1)_kk = w^T C w with w = L^-1 e_k (forward substitution)
    for (int k = 0; k < 6; ++k) {
        double w[6] = {0, 0, 0, 0, 0, 0};
        w[k] = 1.0;
        chol6_forward(U, w, 1);
        double q = 0.0;
        for (int i = 0; i < 6; ++i) {
            double s = 0.0;
            for (int j = 0; j < 6; ++j) s += C[i * 6 + j] * w[j];
            q += w[i] * s;
        }
        P.diag[6 + 6 * (int64_t)f + k] = q;
    }
}

}  // namespace mccba
