// mccba_dense.cuh -- dense SPD solve of the reduced camera system S dc = g (n_s = 6(nC-1); 378 for 64 cameras).
//
// Replaces the reference's Eigen ConjugateGradient on the full P x P system (src/multicalib.cpp:565-592): after
// the Schur complement only the camera unknowns are left, and a direct factorisation gives the exact solution the CG
// iterates converge to.
//
// Right-looking tiled Cholesky (tile 32) on the augmented matrix A = [S; g^T] ((n+1) x n, row-major, ld = n, exactly
// the all-reduce buffer layout), so y = L^-1 g falls out of the factorisation as the last row.  One block column per
// step, two kernels per step inside the iteration graph:
//   chol_panel_kernel   1 CTA : panel (rows c0..n) -> shared memory; diagonal tile factored in 8-column sub-blocks
//                               (one thread factors the 8x8 block in registers, the CTA does the sub-panel solve and
//                               rank-8 update); rows below: one thread per row, forward substitution in registers
//   chol_update_kernel  grid  : trailing tiles (i,j) -= L_ik L_jk^T, one CTA per 32x32 tile, spread over the SMs
// then chol_backward (inside camera_update_kernel, 1 CTA): blocked backward substitution L^T x = y.
// Only the lower triangle of S is read.  The 1/L_jj are kept in a global array for the triangular solves.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mccba {

constexpr int kCT = 32;        // tile
constexpr int kCLD = kCT + 1;  // padded shared-memory row stride (doubles)
constexpr int kPanelThreads = 512;
constexpr int kUpdThreads = 256;

// rows: the panel itself (n+1) and at least 40 so that the 8-wide register blocks may read (zero/garbage, unused)
// rows up to the next multiple of 8 past a ragged tile
__host__ __device__ inline size_t chol_panel_smem_bytes(int n) { return sizeof(double) * (size_t)(n + 2 > 40 ? n + 2 : 40) * kCLD; }
__host__ __device__ inline int chol_col_tiles(int n) { return (n + kCT - 1) / kCT; }
__host__ __device__ inline int chol_row_tiles(int n) { return (n + 1 + kCT - 1) / kCT; }

// Factor block column k.  fail: set to 1 if a pivot is not positive / finite.
__device__ inline void chol_panel(double* __restrict__ A, int n, int k, int* fail, double* __restrict__ rinv_g, double* panel)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ double s_rinv[kCT];
    const int c0 = k * kCT;
    const int w = min(kCT, n - c0);
    const int R = n + 1 - c0;  // panel rows c0 .. n (row n is g)
    {   // load: 4 independent global loads in flight per thread
        const int total = R * kCT;
        int idx = tid;
        for (; idx + 3 * nt < total; idx += 4 * nt) {
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int id = idx + q * nt, r = id >> 5, c = id & 31;
                v[q] = c < w ? A[(int64_t)(c0 + r) * n + c0 + c] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int id = idx + q * nt;
                panel[(id >> 5) * kCLD + (id & 31)] = v[q];
            }
        }
        for (; idx < total; idx += nt) {
            const int r = idx >> 5, c = idx & 31;
            panel[r * kCLD + c] = c < w ? A[(int64_t)(c0 + r) * n + c0 + c] : 0.0;
        }
    }
    __syncthreads();
    // diagonal tile, 8 columns at a time
    for (int kk = 0; kk < w; kk += 8) {
        const int bw = min(8, w - kk);
        if (tid == 0) {
            double b[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j)
                    b[i][j] = (i < bw) ? panel[(kk + i) * kCLD + kk + j] : (i == j ? 1.0 : 0.0);
            int bad = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double d = b[j][j];
                if (!(d > 0.0) || !isfinite(d)) bad = 1;
                const double rinv = rsqrt(d > 0.0 ? d : 1.0);
                b[j][j] = d * rinv;
                if (j < bw) { s_rinv[kk + j] = rinv; rinv_g[c0 + kk + j] = rinv; }
#pragma unroll
                for (int i = j + 1; i < 8; ++i) b[i][j] *= rinv;
#pragma unroll
                for (int i = j + 1; i < 8; ++i)
#pragma unroll
                    for (int m = j + 1; m <= i; ++m) b[i][m] -= b[i][j] * b[m][j];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j)
                    if (i < bw) panel[(kk + i) * kCLD + kk + j] = b[i][j];
            if (bad) *fail = 1;
        }
        __syncthreads();
        // rows of the tile below this sub-block: 8-column triangular solve, one thread per row
        const int below = w - (kk + 8);
        if (below > 0) {
            if (tid < below) {
                double* prow = panel + (kk + 8 + tid) * kCLD + kk;
                double x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    double s = prow[q];
#pragma unroll
                    for (int m = 0; m < q; ++m) s -= x[m] * panel[(kk + q) * kCLD + kk + m];
                    x[q] = s * s_rinv[kk + q];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) prow[q] = x[q];
            }
            __syncthreads();
            // rank-8 update of the remaining lower part of the tile
            for (int idx = tid; idx < below * below; idx += nt) {
                const int r = idx / below, c = idx % below;
                if (c <= r) {
                    const double* pr = panel + (kk + 8 + r) * kCLD + kk;
                    const double* pc = panel + (kk + 8 + c) * kCLD + kk;
                    double s = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) s += pr[q] * pc[q];
                    panel[(kk + 8 + r) * kCLD + kk + 8 + c] -= s;
                }
            }
            __syncthreads();
        }
    }
    // rows below the diagonal tile: x L_kk^T = p, one thread per row, 8 columns at a time in registers
    for (int r = w + tid; r < R; r += nt) {
        double* prow = panel + r * kCLD;
        for (int cb = 0; cb < w; cb += 8) {
            double s8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) s8[q] = prow[cb + q];  // cols >= w are zero-filled
            for (int m = 0; m < cb; ++m) {
                const double xm = prow[m];
#pragma unroll
                for (int q = 0; q < 8; ++q) s8[q] -= xm * panel[(cb + q) * kCLD + m];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int m = 0; m < q; ++m) s8[q] -= s8[m] * panel[(cb + q) * kCLD + cb + m];
                s8[q] *= (cb + q < w) ? s_rinv[cb + q] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (cb + q < w) prow[cb + q] = s8[q];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < R * kCT; idx += nt) {
        const int r = idx >> 5, c = idx & 31;
        if (c < w) A[(int64_t)(c0 + r) * n + c0 + c] = panel[r * kCLD + c];
    }
}

// Trailing update of step k for tile (ti, tj): A[ti][tj] -= L[ti][k] * L[tj][k]^T.  256 threads, one tile per CTA.
__device__ inline void chol_update_tile(double* __restrict__ A, int n, int k, int ti, int tj)
{
    __shared__ double sa[kCT][kCLD], sb[kCT][kCLD];
    const int tid = threadIdx.x;
    const int r0 = ti * kCT, q0 = tj * kCT, c0 = k * kCT;
    const int nrow = n + 1;  // rows of the augmented matrix
    for (int idx = tid; idx < kCT * kCT; idx += kUpdThreads) {
        const int r = idx >> 5, c = idx & 31;
        sa[r][c] = (r0 + r < nrow) ? A[(int64_t)(r0 + r) * n + c0 + c] : 0.0;   // c0 + c < n always (k is not the last tile)
        sb[r][c] = (q0 + r < n) ? A[(int64_t)(q0 + r) * n + c0 + c] : 0.0;
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;  // outputs rows 2ty, 2ty+1; cols tx, tx+16
    double acc[2][2] = {{0, 0}, {0, 0}};
#pragma unroll 8
    for (int m = 0; m < kCT; ++m) {
        const double a0 = sa[2 * ty][m], a1 = sa[2 * ty + 1][m];
        const double b0 = sb[tx][m], b1 = sb[tx + 16][m];
        acc[0][0] += a0 * b0; acc[0][1] += a0 * b1;
        acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r = r0 + 2 * ty + i, c = q0 + tx + 16 * j;
            if (r < nrow && c < n && c <= r) A[(int64_t)r * n + c] -= acc[i][j];
        }
}

// Backward substitution L^T x = y (y = row n of A), block columns from last to first.  One CTA (any multiple of 32
// threads).  smem: n + (blockDim/32) * kCLD doubles.
__device__ inline void chol_backward(const double* __restrict__ A, int n, const double* __restrict__ rinv_g,
                                     double* __restrict__ xout, double* smem)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = nt >> 5;
    double* y = smem;
    double* part = smem + n + (n & 1);
    for (int i = tid; i < n; i += nt) y[i] = A[(int64_t)n * n + i];
    __syncthreads();
    for (int k = chol_col_tiles(n) - 1; k >= 0; --k) {
        const int c0 = k * kCT, w = min(kCT, n - c0), c1 = c0 + w;
        // y[c0+c] -= sum_{i >= c1} L[i][c0+c] * x[i]; lanes over c (coalesced rows), warps over i
        double s = 0.0;
        if (lane < w) {
            int i = c1 + warp;
            for (; i + 3 * nwarps < n; i += 4 * nwarps) {
                const double l0 = A[(int64_t)i * n + c0 + lane], l1 = A[(int64_t)(i + nwarps) * n + c0 + lane];
                const double l2 = A[(int64_t)(i + 2 * nwarps) * n + c0 + lane], l3 = A[(int64_t)(i + 3 * nwarps) * n + c0 + lane];
                s += l0 * y[i] + l1 * y[i + nwarps] + l2 * y[i + 2 * nwarps] + l3 * y[i + 3 * nwarps];
            }
            for (; i < n; i += nwarps) s += A[(int64_t)i * n + c0 + lane] * y[i];
        }
        part[warp * kCLD + lane] = s;
        __syncthreads();
        if (warp == 0) {
            double t = 0.0;
            for (int q = 0; q < nwarps; ++q) t += part[q * kCLD + lane];
            double yv = lane < w ? y[c0 + lane] - t : 0.0;
            // x_c = (y_c - sum_{m>c} L[m][c] x_m) / L[c][c]; lane holds column `lane` of the tile
            double col[kCT];
#pragma unroll
            for (int c = 0; c < kCT; ++c) col[c] = (c < w && lane < c) ? A[(int64_t)(c0 + c) * n + c0 + lane] : 0.0;
            const double myrinv = lane < w ? rinv_g[c0 + lane] : 1.0;
#pragma unroll
            for (int c = kCT - 1; c >= 0; --c) {
                if (c < w) {
                    const double xc = __shfl_sync(0xffffffffu, yv * myrinv, c);
                    if (lane == c) yv = xc;
                    yv -= col[c] * xc;  // col[c] is zero for lanes >= c
                }
            }
            if (lane < w) { y[c0 + lane] = yv; xout[c0 + lane] = yv; }
        }
        __syncthreads();
    }
}

}  // namespace mccba
