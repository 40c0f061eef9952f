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

// 1/sqrt(d) for the pivot chain: float MUFU seed + two Newton steps in double (error ~1e-15 relative), about 4x
// shorter than the library rsqrt() on the critical path; falls back to rsqrt() outside the float range.
__device__ __forceinline__ double fast_rsqrt(double d)
{
    if (!(d > 1e-30 && d < 1e30)) return rsqrt(d);
    double y = (double)rsqrtf((float)d);
    const double hd = 0.5 * d;
    y = y * fma(-hd * y, y, 1.5);
    y = y * fma(-hd * y, y, 1.5);
    y = y * fma(-hd * y, y, 1.5);
    return y;
}

// ---- shared tile helpers (used by the panel kernel above and by the one-launch tile DAG below) ------------------
// In-place Cholesky of the w x w lower triangle held in C (stride kCLD), 8 columns at a time: one thread factors the
// 8x8 block in registers, the CTA solves the rows below it and applies the rank-8 update.  rinv[0..w) <- 1 / L_jj.
__device__ inline void tile_potrf(double* C, int w, double* rinv, int* bad_flag)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int kk = 0; kk < w; kk += 8) {
        const int bw = min(8, w - kk);
        if (tid == 0) {
            double b[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) b[i][j] = (i < bw) ? C[(kk + i) * kCLD + kk + j] : (i == j ? 1.0 : 0.0);
            int bad = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double d = b[j][j];
                if (!(d > 0.0) || !isfinite(d)) bad = 1;
                const double ri = fast_rsqrt(d > 0.0 ? d : 1.0);
                b[j][j] = d * ri;
                if (j < bw) rinv[kk + j] = ri;
#pragma unroll
                for (int i = j + 1; i < 8; ++i) b[i][j] *= ri;
#pragma unroll
                for (int i = j + 1; i < 8; ++i)
#pragma unroll
                    for (int m = j + 1; m <= i; ++m) b[i][m] -= b[i][j] * b[m][j];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j)
                    if (i < bw) C[(kk + i) * kCLD + kk + j] = b[i][j];
            if (bad) *bad_flag = 1;
        }
        __syncthreads();
        const int below = w - (kk + 8);
        if (below > 0) {
            if (tid < below) {
                double* prow = C + (kk + 8 + tid) * kCLD + kk;
                double x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    double sacc = prow[q];
#pragma unroll
                    for (int m = 0; m < q; ++m) sacc -= x[m] * C[(kk + q) * kCLD + kk + m];
                    x[q] = sacc * rinv[kk + q];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) prow[q] = x[q];
            }
            __syncthreads();
            for (int idx = tid; idx < below * below; idx += nt) {
                const int r = idx / below, c = idx % below;
                if (c <= r) {
                    const double* pr = C + (kk + 8 + r) * kCLD + kk;
                    const double* pc = C + (kk + 8 + c) * kCLD + kk;
                    double sacc = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) sacc += pr[q] * pc[q];
                    C[(kk + 8 + r) * kCLD + kk + 8 + c] -= sacc;
                }
            }
            __syncthreads();
        }
    }
}

// One row x of a tile: x L^T = p (forward substitution against the w x w factor L, stride kCLD), 8 columns at a time.
// Columns >= w of the row must be zero; rows of L up to the next multiple of 8 past w may hold anything finite.
__device__ inline void tile_trsm_row(double* prow, const double* L, const double* rinv, int w)
{
    for (int cb = 0; cb < w; cb += 8) {
        double s8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) s8[q] = prow[cb + q];
        for (int m = 0; m < cb; ++m) {
            const double xm = prow[m];
#pragma unroll
            for (int q = 0; q < 8; ++q) s8[q] -= xm * L[(cb + q) * kCLD + m];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
#pragma unroll
            for (int m = 0; m < q; ++m) s8[q] -= s8[m] * L[(cb + q) * kCLD + cb + m];
            s8[q] *= (cb + q < w) ? rinv[cb + q] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (cb + q < w) prow[cb + q] = s8[q];
    }
}

// ---- one-launch tile DAG ---------------------------------------------------------------------------------------
// One CTA per lower-triangle tile (i, j) of the augmented matrix, launched in column-major order so that every
// dependency points to a CTA with a smaller block index (progress is guaranteed even if not all CTAs are resident).
// The owner keeps its tile in shared memory, applies the updates of block columns k < j as their tiles are published,
// finalises it (diagonal: Cholesky; below: triangular solve) and publishes it once to global memory with a
// release flag.  The diagonal CTAs then run the backward substitution the same way.  Critical path per block column:
// POTRF -> TRSM -> one tile update, instead of a kernel boundary per phase.
struct CholDag {
    double* A;        // (n+1) x n
    int n;
    int* ready;       // ntr x ntc tile flags, then ntc flags for the backward substitution (zeroed before the launch)
    double* rinv;     // n
    double* x;        // n: solution
    const int* go;    // optional gate
    int* fail;
};

__device__ __forceinline__ void dag_wait(const int* flag)
{
    if (threadIdx.x == 0) {
        int v;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        } while (v == 0);
    }
    __syncthreads();
}
__device__ __forceinline__ void dag_publish(int* flag)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
    }
}

__device__ inline void chol_dag_tile(const CholDag& D)
{
    __shared__ double C[kCT][kCLD], A1[kCT + 8][kCLD], B1[kCT + 8][kCLD];
    __shared__ double s_rinv[kCT], s_part[8][kCT];
    __shared__ int s_bad;
    const int tid = threadIdx.x, n = D.n;
    const int ntc = chol_col_tiles(n), ntr = chol_row_tiles(n);
    // block index -> (i, j), column-major over the lower triangle
    int j = 0, rem = blockIdx.x;
    while (rem >= ntr - j) { rem -= ntr - j; ++j; }
    const int i = j + rem;
    const int r0 = i * kCT, c0 = j * kCT;
    const int h = min(kCT, n + 1 - r0), w = min(kCT, n - c0);
    if (tid == 0) s_bad = 0;
    for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) {
        const int r = idx >> 5, c = idx & 31;
        C[r][c] = (r < h && c < w) ? D.A[(int64_t)(r0 + r) * n + c0 + c] : 0.0;
    }
    for (int idx = tid; idx < 8 * kCLD; idx += blockDim.x) { (&A1[kCT][0])[idx] = 0.0; (&B1[kCT][0])[idx] = 0.0; }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;
    for (int k = 0; k < j; ++k) {
        dag_wait(D.ready + i * ntc + k);
        if (i != j) dag_wait(D.ready + j * ntc + k);
        const int k0 = k * kCT;
        for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) {
            const int r = idx >> 5, c = idx & 31;
            A1[r][c] = (r < h) ? __ldcg(D.A + (int64_t)(r0 + r) * n + k0 + c) : 0.0;
            if (i != j) B1[r][c] = (r < w) ? __ldcg(D.A + (int64_t)(c0 + r) * n + k0 + c) : 0.0;
        }
        __syncthreads();
        const double(*Bm)[kCLD] = (i != j) ? B1 : A1;
        double acc[2][2] = {{0, 0}, {0, 0}};
#pragma unroll 8
        for (int m = 0; m < kCT; ++m) {
            const double a0 = A1[2 * ty][m], a1 = A1[2 * ty + 1][m];
            const double b0 = Bm[tx][m], b1 = Bm[tx + 16][m];
            acc[0][0] += a0 * b0; acc[0][1] += a0 * b1;
            acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
        }
        C[2 * ty][tx] -= acc[0][0]; C[2 * ty][tx + 16] -= acc[0][1];
        C[2 * ty + 1][tx] -= acc[1][0]; C[2 * ty + 1][tx + 16] -= acc[1][1];
        __syncthreads();
    }
    if (i == j) {
        tile_potrf(&C[0][0], w, s_rinv, &s_bad);
        __syncthreads();
        if (h > w && tid == 0) tile_trsm_row(&C[w][0], &C[0][0], s_rinv, w);   // the g row lives in this tile (ragged last column)
        __syncthreads();
        for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) {
            const int r = idx >> 5, c = idx & 31;
            if (r < h && c < w && c <= r) D.A[(int64_t)(r0 + r) * n + c0 + c] = C[r][c];
        }
        if (tid < w) D.rinv[c0 + tid] = s_rinv[tid];
        if (tid == 0 && s_bad) *D.fail = 1;
        dag_publish(D.ready + i * ntc + j);
    } else {
        dag_wait(D.ready + j * ntc + j);
        for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) {
            const int r = idx >> 5, c = idx & 31;
            B1[r][c] = (r < w && c <= r) ? __ldcg(D.A + (int64_t)(c0 + r) * n + c0 + c) : 0.0;
        }
        if (tid < kCT) s_rinv[tid] = tid < w ? __ldcg(D.rinv + c0 + tid) : 0.0;
        __syncthreads();
        if (tid < h) tile_trsm_row(&C[tid][0], &B1[0][0], s_rinv, w);
        __syncthreads();
        for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) {
            const int r = idx >> 5, c = idx & 31;
            if (r < h && c < w) D.A[(int64_t)(r0 + r) * n + c0 + c] = C[r][c];
        }
        dag_publish(D.ready + i * ntc + j);
        return;
    }
    // ---- backward substitution, diagonal CTAs only: x_j = L_jj^-T (y_j - sum_{t>j} L_tj^T x_t) ------------------
    int* xready = D.ready + ntr * ntc;
    const int ig = n / kCT;                       // row tile that holds the g row (row n)
    if (ig != j) dag_wait(D.ready + ig * ntc + j);   // y_j = A[n][c0 ..] final
    const int lane = tid & 31, wp = tid >> 5;        // 8 warps over rows, lanes over columns
    double sacc = 0.0;
    for (int t = ntc - 1; t > j; --t) {
        dag_wait(xready + t);                        // also implies tile (t, j) is final (x_t needed it... see below)
        dag_wait(D.ready + t * ntc + j);
        const int rt = t * kCT, ht = min(kCT, n - rt);
        if (lane < w)
            for (int r = wp; r < ht; r += 8) sacc += __ldcg(D.A + (int64_t)(rt + r) * n + c0 + lane) * __ldcg(D.x + rt + r);
    }
    s_part[wp][lane] = sacc;
    __syncthreads();
    if (wp == 0) {
        double tsum = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) tsum += s_part[q][lane];
        double yv = 0.0;
        if (lane < w) yv = (ig == j ? C[w][lane] : __ldcg(D.A + (int64_t)n * n + c0 + lane)) - tsum;
        const double myrinv = lane < w ? s_rinv[lane] : 1.0;
        for (int c = w - 1; c >= 0; --c) {
            const double xc = __shfl_sync(0xffffffffu, yv * myrinv, c);
            if (lane == c) yv = xc;
            if (lane < c) yv -= C[c][lane] * xc;
        }
        if (lane < w) D.x[c0 + lane] = yv;
    }
    dag_publish(xready + j);
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
