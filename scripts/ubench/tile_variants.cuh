// Earlier variants of the tile kernels, kept only for the comparisons in lat.cu (not part of the product):
// tile_potrf (one thread factors 8 x 8 blocks), tile_trsm_row (one thread per row), tile_potrf_blocked (8-column panels,
// shuffle-based pivot exchange, DMMA trailing update).  Include after mccba_dense.cuh.
#pragma once
namespace mccba {

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

// Blocked Cholesky of the w x w lower triangle in C (stride kCLD), called by all 256 threads.  Per 8-column panel:
// warp 0 eliminates the 8 pivots with lane = row and the panel in registers -- the dependent chain per pivot is
// shuffle(d_j) -> 1/d_j -> a_ij/d_j -> fma, about 90 cycles -- then all warps apply the rank-8 update of the
// trailing tiles with DMMA.  Columns stay unscaled (R) next to their 1/d_j-scaled copy T until one final pass
// multiplies by 1/sqrt(d_j).  `T` = 32 x 8 doubles of shared memory.  rinv[0..w) <- 1 / L_jj; rows >= w of C are not
// touched.  A pivot outside [1e-200, 1e200] = failure.
__device__ inline void tile_potrf_blocked(double* C, int w, double* rinv, int* bad_flag, double* T)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    for (int p0 = 0; p0 < w; p0 += 8) {
        if (warp == 0) {
            double a[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = lane < w ? C[lane * kCLD + p0 + k] : (p0 + k == lane ? 1.0 : 0.0);
            int bad = 0;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = p0 + jj;
                if (j < w) {
                    const double d = __shfl_sync(0xffffffffu, a[jj], j);
                    double ck[8];
#pragma unroll
                    for (int k = jj + 1; k < 8; ++k) ck[k] = __shfl_sync(0xffffffffu, a[jj], p0 + k);
                    const double tt = a[jj] * pivot_rcp(d);
                    if (!(d > 1e-200 && d < 1e200)) bad = 1;
                    T[lane * 8 + jj] = lane > j ? tt : 0.0;
#pragma unroll
                    for (int k = jj + 1; k < 8; ++k) a[k] = fma(-tt, ck[k], a[k]);
                } else {
                    T[lane * 8 + jj] = 0.0;
                }
            }
            if (lane < w) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (p0 + k < w && lane >= p0 + k) C[lane * kCLD + p0 + k] = a[k];
            }
            if (bad) *bad_flag = 1;
        }
        __syncthreads();
        // trailing tiles (tr >= tc > b): C_tr,tc -= T_tr R_tc^T, R = the unscaled panel columns
        const int b = p0 >> 3, nt = 3 - b;
        if (warp < nt * (nt + 1) / 2 && p0 + 8 < w) {
            int tr = 0, rem = warp;
            while (rem > tr) { rem -= tr + 1; ++tr; }
            const int tc = rem + b + 1;
            tr += b + 1;
            double* cp = C + (8 * tr + g) * kCLD + 8 * tc + 2 * t4;
            double c0 = cp[0], c1 = cp[1];
            const double* tp = T + (8 * tr + g) * 8 + t4;
            const double* rp = C + (8 * tc + g) * kCLD + p0 + t4;
            dmma(c0, c1, -tp[0], rp[0]);
            dmma(c0, c1, -tp[4], rp[4]);
            cp[0] = c0; cp[1] = c1;
        }
        __syncthreads();
    }
    if (tid < kCT) rinv[tid] = tid < w ? rsqrt(C[tid * kCLD + tid]) : 0.0;
    __syncthreads();
    {
        const int r = tid >> 5, c = tid & 31;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (r + 8 * q < w && c <= r + 8 * q) C[(r + 8 * q) * kCLD + c] *= rinv[c];
    }
    __syncthreads();
}

}  // namespace mccba
