// mccba_omni.cuh -- the cv::omnidir::calibrate optimisation loop (single Mei camera: per-frame poses + 10 intrinsics)
// on the device.  Replaces src/omnidir.cpp:1119-1147 (loop), :851-935 (internal::computeJacobian: arrow-structured
// J^T J / J^T E, flag masking, (JTJ + eps)^-1), :2031-2076 (flags2idx), :2138-2153 (fillFixed) and the RMS of
// :1794-1802.  The reference forms a dense (6n+10)^2 matrix and inverts it; here the per-frame 6x6 blocks are
// eliminated and the "+ eps on EVERY element" term (a rank-one eps 11^T) is carried as one extra bordered unknown, so
// the iterates are the reference's to rounding (this path works in the reference's additive Rodrigues coordinates,
// because the eps 11^T term is not invariant under a change of coordinates).
//
//   omni_frame_kernel    warp per frame: rows a = [J(16) | e] of every corner -> shared memory; the 17x17 Gram matrix
//                        sum a^T a (= all arrow blocks + gradient + cost of the frame) on the FP64 tensor cores; then
//                        the frame's 6x6 Cholesky, Y = L^-1 H_pI, z_g, z_u and its Schur record
//   omni_reduce_kernel   fixed-order sums of the 78 record entries over the frames (no atomics)
//   omni_solve_kernel    bordered (m+1) x (m+1) system, Gaussian elimination with partial pivoting (m <= 10), one warp
//   omni_update_kernel   back-substitution per frame, G = alpha * x, parameter update, norm partials
//   omni_decide_kernel   change = |G| / |param_old| (:1141), iteration count, termination (:1125-1127)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mccba_math.cuh"
#include "mccba_warp_solve.cuh"

namespace mccba {

#ifndef MCCBA_OMNI_MINBLOCKS
#define MCCBA_OMNI_MINBLOCKS 3
#endif
constexpr int kOmniWarps = 4;        // frames per CTA: one warp per frame
constexpr int kOmniThreads = 32 * kOmniWarps;
constexpr int kOmniChunk = 32;      // corners staged per pass: one per lane (64 rows)
constexpr int kOmniLd = 20;         // row stride (doubles) of the staged rows: the 8 x 4 MMA fragment loads take 2 wavefronts, the minimum
constexpr int kOmniRec = 78;        // 55 (S upper) + 10 (rg) + 10 (ru) + d, c, cost
constexpr int kOmniSave = 93;       // 21 (U) + 6 (z_g) + 6 (z_u) + 60 (Y)

struct OmniState {
    int flags, crit_type, max_count, iter, done, status;
    double eps_crit, change, alpha, epsilon;
    double x_intr[10];   // solution for the intrinsic block (0 at fixed parameters)
    double t;            // bordered unknown sqrt(eps) 1^T x
};

struct OmniProblem {
    int n_frame;
    int n_blocks_upd;
    int64_t n_pts;
    const float *ox, *oy, *oz, *iu, *iv;
    const int* f_off;     // n_frame + 1
    double* param;        // 6 n + 10
    double* rec;          // kOmniRec x n_frame (SoA)
    double* save;         // kOmniSave x n_frame (SoA)
    double* tot;          // kOmniRec
    double* norm_part;    // 2 x n_blocks_upd
    double* norm_tot;     // 2: |G|^2, |param_old|^2 (frame-sharded runs: summed over the ranks before the decision)
    double* dump;         // optional: n_frame x 289 Gram matrices (tests)
    OmniState* st;
    int count_intr;       // 1 on the rank that adds the 10 intrinsics to the norms (rank 0); every rank updates them
};

// One WARP per frame.  Per pass of 32 corners every lane evaluates one corner (projection + the two 17-wide rows
// a = [J(16) | e]) into the warp's shared-memory stage; the 17 x 17 Gram matrix sum a^T a is then accumulated by the
// FP64 tensor cores: with R the staged 64 x 17 row block, the six upper 8 x 8 tiles of R^T R are 16 k-steps of
// mma.m8n8k4 each, and the A and B fragments of a tile pair are the same three 8-column slices of R (3 loads per
// k-step per lane for 6 MMAs).  The per-frame algebra (6 x 6 Cholesky, 12 forward solves, Schur record) follows in the
// same warp.  (Round 1: one 160-thread CTA per frame, 153 threads x 108 scalar FMAs, ~200 registers for every thread of
// which 54 projected: 91 us per pass over 5000 frames.)
__global__ void __launch_bounds__(kOmniThreads, MCCBA_OMNI_MINBLOCKS) omni_frame_kernel(OmniProblem P, int forced)
{
    __shared__ __align__(16) double s_rows[kOmniWarps][2 * kOmniChunk * kOmniLd];
    __shared__ double s_fac[kOmniWarps][96];    // U 21 | z_g 6 | z_u 6 | Y 6 x 10
    const OmniState* st = P.st;
    if (!forced && st->done) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, n = P.n_frame;
    const int f = blockIdx.x * kOmniWarps + w;
    if (f >= n) return;                          // whole warps leave; nothing below synchronises across warps
    double* rows = s_rows[w];
    const double* par = P.param;
    CamParams cam;
    cam.model = kOmnidir; cam.rational = 0;
    cam.fx = par[6 * n]; cam.fy = par[6 * n + 1]; cam.skew = par[6 * n + 2]; cam.cx = par[6 * n + 3]; cam.cy = par[6 * n + 4];
    cam.xi = par[6 * n + 5]; cam.k1 = par[6 * n + 6]; cam.k2 = par[6 * n + 7]; cam.p1 = par[6 * n + 8]; cam.p2 = par[6 * n + 9];
    cam.k3 = cam.k4 = cam.k5 = cam.k6 = 0.0;
    const double om[3] = {par[6 * f], par[6 * f + 1], par[6 * f + 2]};
    const double T[3] = {par[6 * f + 3], par[6 * f + 4], par[6 * f + 5]};
    double R[9], Jl[9];
    rodrigues(om, R);
    left_jacobian(om, Jl);
    const int g = lane >> 2, t = lane & 3;
    double acc[6][2];
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[i][0] = acc[i][1] = 0.0;
    const int b = P.f_off[f], e = P.f_off[f + 1];
    for (int c0 = b; c0 < e; c0 += kOmniChunk) {
        const int nc = min(kOmniChunk, e - c0);
        __syncwarp();                            // the previous pass has been consumed
        double* row0 = rows + (2 * lane) * kOmniLd;
        double* row1 = row0 + kOmniLd;
        if (lane < nc) {
            const int i = c0 + lane;
            const double X[3] = {(double)P.ox[i], (double)P.oy[i], (double)P.oz[i]};
            double Q[3], Xc[3], uv[2], A[6], Jin[20];
            mat3_vec(R, X, Q);
            Xc[0] = Q[0] + T[0]; Xc[1] = Q[1] + T[1]; Xc[2] = Q[2] + T[2];
            omnidir_point_full(cam, Xc, uv, A, Jin);
            const double err[2] = {(double)P.iu[i] - uv[0], (double)P.iv[i] - uv[1]};
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                double jphi[3];
                cross3(Q, A + 3 * r, jphi);
                double* row = r == 0 ? row0 : row1;
#pragma unroll
                for (int k = 0; k < 3; ++k) row[k] = jphi[0] * Jl[k] + jphi[1] * Jl[3 + k] + jphi[2] * Jl[6 + k];   // d/d om
#pragma unroll
                for (int k = 0; k < 3; ++k) row[3 + k] = A[3 * r + k];
#pragma unroll
                for (int k = 0; k < 10; ++k) row[6 + k] = Jin[10 * r + k];
                row[16] = err[r];
            }
        } else {
#pragma unroll
            for (int k = 0; k < 17; ++k) { row0[k] = 0.0; row1[k] = 0.0; }     // rows past the last corner add zeros
        }
        __syncwarp();
        const int ksteps = (2 * nc + 3) >> 2;
        for (int ks = 0; ks < ksteps; ++ks) {
            const double* r = rows + (4 * ks + t) * kOmniLd + g;
            const double f0 = r[0], f1 = r[8], f2 = g == 0 ? r[16] : 0.0;
            dmma(acc[0][0], acc[0][1], f0, f0);
            dmma(acc[1][0], acc[1][1], f0, f1);
            dmma(acc[2][0], acc[2][1], f0, f2);
            dmma(acc[3][0], acc[3][1], f1, f1);
            dmma(acc[4][0], acc[4][1], f1, f2);
            dmma(acc[5][0], acc[5][1], f2, f2);
        }
    }
    __syncwarp();
    double (*M)[17] = reinterpret_cast<double (*)[17]>(rows);   // the Gram matrix takes the place of the staged rows
    {
        const int mt[6] = {0, 0, 0, 1, 1, 2}, nt[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = 8 * mt[i] + g, c = 8 * nt[i] + 2 * t + h;
                if (r < 17 && c < 17 && r <= c) { M[r][c] = acc[i][h]; M[c][r] = acc[i][h]; }
            }
    }
    __syncwarp();
    if (P.dump)
        for (int k = lane; k < 289; k += 32) P.dump[(int64_t)f * 289 + k] = M[k / 17][k % 17];
    // per-frame algebra: Cholesky of H_pp (lane 0), then z_g, z_u and the 10 columns of Y (lanes 0..11), Schur record
    double* sU = s_fac[w];
    double* szg = sU + 21;
    double* szu = sU + 27;
    double (*sY)[10] = reinterpret_cast<double (*)[10]>(sU + 33);
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
    if (lane < 12) {
        double U[21], col[6];
#pragma unroll
        for (int i = 0; i < 21; ++i) U[i] = sU[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) col[i] = lane < 10 ? M[i][6 + lane] : (lane == 10 ? M[i][16] : 1.0);
        chol6_forward(U, col, 1);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (lane < 10) sY[i][lane] = col[i];
            else if (lane == 10) szg[i] = col[i];
            else szu[i] = col[i];
        }
    }
    __syncwarp();
    const int64_t nf = n;
    for (int tid = lane; tid < 76; tid += 32) {
        if (tid < 55) {   // S record: upper triangle of H_II - Y^T Y
            int a = 0, rem = tid;
            while (rem >= 10 - a) { rem -= 10 - a; ++a; }
            const int c = a + rem;
            double s = M[6 + a][6 + c];
#pragma unroll
            for (int k = 0; k < 6; ++k) s -= sY[k][a] * sY[k][c];
            P.rec[(int64_t)tid * nf + f] = s;
        } else if (tid < 65) {
            const int a = tid - 55;
            double s = M[6 + a][16];
#pragma unroll
            for (int k = 0; k < 6; ++k) s -= sY[k][a] * szg[k];
            P.rec[(int64_t)tid * nf + f] = s;
        } else if (tid < 75) {
            const int a = tid - 65;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) s -= sY[k][a] * szu[k];
            P.rec[(int64_t)tid * nf + f] = s;
        } else {
            double d = 0, c = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) { d += szu[k] * szu[k]; c += szu[k] * szg[k]; }
            P.rec[(int64_t)75 * nf + f] = d;
            P.rec[(int64_t)76 * nf + f] = c;
            P.rec[(int64_t)77 * nf + f] = M[16][16];
        }
    }
    // saved factors for the back-substitution
    for (int k = lane; k < kOmniSave; k += 32) P.save[(int64_t)k * nf + f] = sU[k];    // U | z_g | z_u | Y are contiguous
    if (lane == 0 && bad) P.st->status = 4;   // benign race: every writer stores the same value
}

// one block per record entry: fixed-order sum over the frames
__global__ void __launch_bounds__(256) omni_reduce_kernel(OmniProblem P, int forced)
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

// src/omnidir.cpp:2031-2076, literal >= / subtract cascade; idx[a] = 1 if intrinsic a is free
__device__ __forceinline__ void omni_flags2free(int flags, int* fr)
{
    for (int a = 0; a < 10; ++a) fr[a] = 1;
    int f = flags;
    if (f >= 256) { fr[3] = 0; fr[4] = 0; f -= 256; }
    if (f >= 128) { fr[0] = 0; fr[1] = 0; f -= 128; }
    if (f >= 64) { fr[5] = 0; f -= 64; }
    if (f >= 32) { fr[9] = 0; f -= 32; }
    if (f >= 16) { fr[8] = 0; f -= 16; }
    if (f >= 8) { fr[7] = 0; f -= 8; }
    if (f >= 4) { fr[6] = 0; f -= 4; }
    if (f >= 2) { fr[2] = 0; }
}

// One warp: lane i holds row i of the bordered system [B | rhs] (at most 11 x 11); the elimination order, the pivot
// rule (first row of maximal modulus) and the order of every sum are those of a serial Gaussian elimination with partial
// pivoting, the rows just live in different lanes (round 1 ran it on one thread out of local memory: 37 us).
__global__ void __launch_bounds__(32) omni_solve_kernel(OmniProblem P)
{
    OmniState* st = P.st;
    // no early exit on the loaded state: everything is computed unconditionally and only the stores at the end are guarded,
    // so the warp provably stays converged
    const int done = st->done, status = st->status;
    const int lane = threadIdx.x;
    // schedule of this iteration (src/omnidir.cpp:1129-1131)
    const double alpha = 1.0 - pow(1.0 - 0.01, (double)st->iter + 1.0);
    const double epsilon = 0.01 * pow(0.9, (double)st->iter / 10.0);
    int fr[10], map[10], m = 0;
    omni_flags2free(st->flags, fr);
    for (int a = 0; a < 10; ++a)
        if (fr[a]) map[m++] = a;
    const double* tot = P.tot;
    const double se = sqrt(epsilon);
    const int Q = m + 1;
    auto Sfull = [&](int a, int c) {   // upper-triangle record index of (a, c)
        if (a > c) { const int t = a; a = c; c = t; }
        int idx = 0;
        for (int r = 0; r < a; ++r) idx += 10 - r;
        return tot[idx + (c - a)];
    };
    // rows and columns Q..10 are padded with the identity, so all 11 elimination steps run unconditionally: every shuffle
    // below sits in straight-line code (no WARPSYNC wrappers), and the padding never wins a pivot search
    double B[11], rhs = 0.0;
#pragma unroll
    for (int c = 0; c < 11; ++c) {
        double v = 0.0;
        if (lane < m) {
            if (c < m) v = Sfull(map[lane], map[c]);
            else if (c == m) v = se * (1.0 + tot[65 + map[lane]]);
        } else if (lane == m) {
            if (c < m) v = se * (1.0 + tot[65 + map[c]]);
            else if (c == m) v = -(1.0 + epsilon * tot[75]);
        } else if (lane == c) v = 1.0;
        B[c] = v;
    }
    if (lane < m) rhs = tot[55 + map[lane]];
    else if (lane == m) rhs = -se * tot[76];
    __shared__ double s_rows[2 * 12];
    double x[11];
    const bool failed = warp_gauss_solve<11>(B, rhs, lane, Q, s_rows, x);
    if (lane == 0 && !done) {
        st->alpha = alpha; st->epsilon = epsilon;
        if (status) st->done = 1;
        else if (failed) { st->status = 4; st->done = 1; }
        else {
            for (int a = 0; a < 10; ++a) st->x_intr[a] = 0.0;
#pragma unroll
            for (int a = 0; a < 10; ++a)
                if (a < m) st->x_intr[map[a]] = x[a];
            double tv = 0.0;
#pragma unroll
            for (int a = 0; a < 11; ++a)
                if (a == m) tv = x[a];
            st->t = tv;
        }
    }
}

__global__ void __launch_bounds__(128) omni_update_kernel(OmniProblem P)
{
    const OmniState* st = P.st;
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
#pragma unroll
            for (int a = 0; a < 10; ++a) s -= P.save[(int64_t)(33 + i * 10 + a) * nf + f] * st->x_intr[a];
            r[i] = s;
        }
        chol6_backward(U, r);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double old = P.param[6 * (int64_t)f + i], G = alpha * r[i];
            g2 += G * G; p2 += old * old;
            P.param[6 * (int64_t)f + i] = old + G;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 10) {   // intrinsic block (fixed parameters get G = 0, fillFixed)
        const int a = threadIdx.x;
        const double old = P.param[6 * (int64_t)n + a], G = alpha * st->x_intr[a];
        if (P.count_intr) { g2 += G * G; p2 += old * old; }
        P.param[6 * (int64_t)n + a] = old + G;
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

// phase 0: sum the norm partials and decide (single rank).  Frame-sharded runs split it around the collective:
// phase 1 writes the rank's two sums to norm_tot, phase 2 decides from the all-reduced norm_tot.
__global__ void __launch_bounds__(256) omni_decide_kernel(OmniProblem P, int phase)
{
    OmniState* st = P.st;
    if (st->done) return;
    __shared__ double sm[2][8];
    double a = 0.0, b = 0.0;
    if (phase != 2)
        for (int k = threadIdx.x; k < P.n_blocks_upd; k += blockDim.x) { a += P.norm_part[k]; b += P.norm_part[P.n_blocks_upd + k]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = a; sm[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double g2 = 0, p2 = 0;
        for (int w = 0; w < 8; ++w) { g2 += sm[0][w]; p2 += sm[1][w]; }
        if (phase == 1) { P.norm_tot[0] = g2; P.norm_tot[1] = p2; return; }
        if (phase == 2) { g2 = P.norm_tot[0]; p2 = P.norm_tot[1]; }
        st->change = sqrt(g2) / sqrt(p2);   // src/omnidir.cpp:1141: the norm of the parameters BEFORE the update
        st->iter += 1;
        const int t = st->crit_type;
        if ((t == 1 && st->iter >= st->max_count) || (t == 2 && st->change <= st->eps_crit) ||
            (t == 3 && (st->change <= st->eps_crit || st->iter >= st->max_count)))
            st->done = 1;
    }
}

__global__ void omni_init_state_kernel(OmniState* st, int flags, int crit_type, int max_count, double eps)
{
    st->flags = flags; st->crit_type = crit_type; st->max_count = max_count; st->iter = 0; st->status = 0;
    st->eps_crit = eps; st->change = 1.0; st->alpha = 0; st->epsilon = 0; st->t = 0;
    for (int a = 0; a < 10; ++a) st->x_intr[a] = 0;
    // loop-top test of src/omnidir.cpp:1125-1127 with iter = 0, change = 1
    st->done = (crit_type == 1 && 0 >= max_count) || (crit_type == 3 && 0 >= max_count) || (crit_type == 2 && 1.0 <= eps) ? 1 : 0;
}

}  // namespace mccba
