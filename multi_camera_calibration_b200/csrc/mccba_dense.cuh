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
constexpr int kCLD = kCT + 2;  // padded shared-memory row stride (doubles): 16-byte aligned rows, 2 wavefronts per fragment load
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

// 1/d for the pivot chain: MUFU.RCP64H seed + two Newton steps (relative error ~1e-16), ~50 cycles instead of ~80
// for the IEEE division.  Caller guarantees d is a normal positive number well inside the double range.
__device__ __forceinline__ double pivot_rcp(double d)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double r = fma(-d, y, 1.0);
    y = fma(y, r, y);
    r = fma(-d, y, 1.0);
    y = fma(y, r, y);
    return y;
}

// Warp-level Cholesky of the w x w lower triangle in C (stride kCLD), executed by ONE warp.  The SM issues in order,
// so everything between a producer and its consumer has to be filled by hand; measured on B200: DFMA 8 cycles
// (2-3 cycles issue), shuffle 25, shared-memory round trip 35, rsqrt() 74, 1/x 79, MUFU.RCP64H 16.
// Lane i keeps the not yet eliminated part of row i in registers, shifted by one column per pivot (the shift is the
// destination of the update fma, so it is free).  The pivot loop is software-pipelined: as soon as the first update
// fma has produced the next pivot column it goes to shared memory, and its broadcast + 1/d_{j+1} overlap the
// remaining 30 fmas of pivot j.  Columns are stored unscaled and multiplied by 1/sqrt(d_j) in one pass at the end.
// `colbuf` = 2 x 64 doubles of shared memory.  A pivot outside [1e-200, 1e200] = failure.  tile_potrf_scale() finishes.
#ifdef MCCBA_POTRF_DBG
__device__ long long g_potrf_ts[40];
#endif
template <int NW>
struct PotrfCol {
    double c[NW - 1];   // column entries below the pivot, broadcast to every lane
    double d, t;        // pivot, a_ij / d_j
};

// Order-pinned arithmetic for the pivot step: the SM issues in order and ptxas, left alone, parks the whole update
// block in front of the reciprocal chain; volatile asm keeps the hand-made interleave (chain op, a few independent
// update fmas in its shadow, next chain op ...).
__device__ __forceinline__ double pin_fma(double a, double b, double c)
{
    double d;
    asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
    return d;
}
__device__ __forceinline__ double pin_nfma(double a, double b, double c)   // c - a * b
{
    double d;
    asm volatile("{ .reg .f64 na; neg.f64 na, %1; fma.rn.f64 %0, na, %2, %3; }" : "=d"(d) : "d"(a), "d"(b), "d"(c));
    return d;
}
__device__ __forceinline__ double pin_nmul(double a, double b)            // -(a * b)
{
    double d;
    asm volatile("{ .reg .f64 na; neg.f64 na, %1; mul.rn.f64 %0, na, %2; }" : "=d"(d) : "d"(a), "d"(b));
    return d;
}
__device__ __forceinline__ double pin_rcp_seed(double a)
{
    double d;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(d) : "d"(a));
    return d;
}

// One pivot with NW live columns (a[k] = A[lane][j + k]); `t` fields hold -a_ij / d_j.  The column of pivot j + 1 is
// broadcast with 16-byte shared-memory loads from an even offset; kNextOdd says whether j + 1 is odd (then the pivot
// is the second word).  The NW - 2 remaining update fmas of pivot j are spread over the six latency gaps of the
// reciprocal chain of pivot j + 1 (MUFU, four Newton fmas, the final multiply).
template <int NW, bool kNextOdd>
__device__ __forceinline__ void potrf_step(double* C, double* colbuf, int lane, int j, int w, double (&a)[NW],
                                           const PotrfCol<NW>& cur, PotrfCol<NW>& nxt, int& bad)
{
#ifdef MCCBA_POTRF_DBG
    if (lane == 0) g_potrf_ts[j] = clock64();
#endif
    const double a0n = pin_fma(cur.t, cur.c[0], a[1]);
    double* cbn = colbuf + ((j + 1) & 1) * 2 * kCT;
    cbn[lane] = a0n;
    __syncwarp();
    const unsigned vaddr = (unsigned)__cvta_generic_to_shared(cbn + ((j + 1) & ~1));
    double v[NW];
#pragma unroll
    for (int p = 0; p < NW / 2; ++p)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[2 * p]), "=d"(v[2 * p + 1]) : "r"(vaddr + 16u * p) : "memory");
    if (lane >= j && lane < w) C[lane * kCLD + j] = a[0];
    if (!(cur.d > 1e-200 && cur.d < 1e200)) bad = 1;
    const double d = kNextOdd ? v[1] : v[0];
    constexpr int nF = NW - 2;
    int k = 1;   // compile-time after unrolling
    double y = pin_rcp_seed(d), r = 0.0, tn = 0.0;
#pragma unroll
    for (int gap = 0; gap < 6; ++gap) {
#pragma unroll
        for (int f = (nF * gap) / 6; f < (nF * (gap + 1)) / 6; ++f, ++k) a[k] = pin_fma(cur.t, cur.c[k], a[k + 1]);
        if (gap == 0) r = pin_nfma(d, y, 1.0);
        else if (gap == 1) y = pin_fma(y, r, y);
        else if (gap == 2) r = pin_nfma(d, y, 1.0);
        else if (gap == 3) y = pin_fma(y, r, y);
        else if (gap == 4) tn = pin_nmul(a0n, y);
    }
    a[0] = a0n;
    a[NW - 1] = 0.0;
    nxt.d = d;
    nxt.t = tn;
    if (!kNextOdd) {
#pragma unroll
        for (int q = 0; q < NW - 1; ++q) nxt.c[q] = v[1 + q];
    } else {
#pragma unroll
        for (int q = 0; q < NW - 2; ++q) nxt.c[q] = v[2 + q];
        nxt.c[NW - 2] = 0.0;   // column j + NW: past the tile
    }
}

__device__ inline void tile_potrf_warp(double* C, int w, int* bad_flag, double* colbuf)
{
    const int lane = threadIdx.x & 31;
    double a[kCT];   // a[k] = A[lane][j + k] at pivot j
    {
        const double2* row = reinterpret_cast<const double2*>(C + lane * kCLD);   // rows are 16-byte aligned
#pragma unroll
        for (int k = 0; k < kCT; k += 2) {
            const double2 q = row[k >> 1];
            a[k] = (lane < w && k <= lane) ? q.x : (k == lane ? 1.0 : 0.0);
            a[k + 1] = (lane < w && k + 1 <= lane) ? q.y : (k + 1 == lane ? 1.0 : 0.0);
        }
    }
    colbuf[kCT + lane] = 1.0;       // what pivots past the tile read
    colbuf[3 * kCT + lane] = 1.0;
    colbuf[lane] = a[0];
    __syncwarp();
    int bad = 0;
    PotrfCol<kCT> c0, c1;
    c0.d = colbuf[0];
#pragma unroll
    for (int k = 0; k < kCT - 1; ++k) c0.c[k] = colbuf[1 + k];
    c0.t = -a[0] * pivot_rcp(c0.d);
    const int wa = min(w, kCT / 2);
#pragma unroll 1
    for (int j = 0; j < wa; j += 2) {   // pivots 0..15: 32 live columns
        potrf_step<kCT, true>(C, colbuf, lane, j, w, a, c0, c1, bad);
        if (j + 1 < wa) potrf_step<kCT, false>(C, colbuf, lane, j + 1, w, a, c1, c0, bad);
    }
    if (w > kCT / 2) {                  // pivots 16..31: 16 live columns
        double b[kCT / 2];
        PotrfCol<kCT / 2> e0, e1;
#pragma unroll
        for (int k = 0; k < kCT / 2; ++k) b[k] = a[k];
        e0.d = c0.d; e0.t = c0.t;
#pragma unroll
        for (int k = 0; k < kCT / 2 - 1; ++k) e0.c[k] = c0.c[k];
#pragma unroll 1
        for (int j = kCT / 2; j < w; j += 2) {
            potrf_step<kCT / 2, true>(C, colbuf, lane, j, w, b, e0, e1, bad);
            if (j + 1 < w) potrf_step<kCT / 2, false>(C, colbuf, lane, j + 1, w, b, e1, e0, bad);
        }
    }
#ifdef MCCBA_POTRF_DBG
    if (lane == 0) g_potrf_ts[32] = clock64();
#endif
    if (bad) *bad_flag = 1;
}

// Second half of the factorisation, all threads (after a barrier): rinv_j = 1 / sqrt(d_j), L_ij = a_ij * rinv_j.
// Ends with a barrier.
__device__ __forceinline__ void tile_potrf_scale(double* C, int w, double* rinv)
{
    const int tid = threadIdx.x;
    if (tid < kCT) rinv[tid] = tid < w ? rsqrt(C[tid * kCLD + tid]) : 0.0;
    __syncthreads();
    const int r = tid >> 5, c = tid & 31;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (r + 8 * q < w && c <= r + 8 * q) C[(r + 8 * q) * kCLD + c] *= rinv[c];
    __syncthreads();
}

// FP64 tensor-core MMA (DMMA), D(8x8) += A(8x4, row) * B(4x8, col).  Lane = 4 g + t holds A[g][t], B[t][g] and
// D[g][2t], D[g][2t+1].  Measured on B200: 33 cycles per dependent op; a 32^3 tile update built on it takes ~1000
// cycles for 256 threads against ~2050 for the scalar DFMA version, which is bound by shared-memory loads.
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C -= A * B^T for 32 x 32 tiles in shared memory (stride kCLD), 256 threads: 8 warps x two 8 x 8 output blocks.
__device__ __forceinline__ void tile_gemm_sub(double* C, const double* A, const double* B)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int tr = warp >> 1, tc0 = (warp & 1) * 2;
    const double* ap = A + (8 * tr + g) * kCLD + t;
    const double* bp0 = B + (8 * tc0 + g) * kCLD + t;
    const double* bp1 = bp0 + 8 * kCLD;
    double c00 = 0, c01 = 0, c10 = 0, c11 = 0, d00 = 0, d01 = 0, d10 = 0, d11 = 0;
#pragma unroll
    for (int ks = 0; ks < 8; ks += 2) {
        const double a0 = ap[4 * ks], a1 = ap[4 * ks + 4];
        dmma(c00, c01, a0, bp0[4 * ks]);
        dmma(c10, c11, a0, bp1[4 * ks]);
        dmma(d00, d01, a1, bp0[4 * ks + 4]);
        dmma(d10, d11, a1, bp1[4 * ks + 4]);
    }
    double* cp = C + (8 * tr + g) * kCLD + 8 * tc0 + 2 * t;
    cp[0] -= c00 + d00; cp[1] -= c01 + d01; cp[8] -= c10 + d10; cp[9] -= c11 + d11;
}

// Rows of a tile against the w x w factor L (stride kCLD): X L^T = P in place for up to 32 rows of P.  Warps 0..3
// take 8 rows each (the other warps return); no block-wide barrier inside.  Each warp first inverts the four 8 x 8
// diagonal blocks of L (one lane per column, `winv` = 4 x 8 x 8 doubles of shared memory -- all warps write the same
// values), then walks the four column blocks right-looking with DMMA: X_b = S_b W_b^T, S_b' -= X_b L_b'b^T (b' > b).
// Only the lower triangle of L is read; rows/columns of L past w must be zero below the diagonal, and rinv[c] for
// c >= w is taken as zero, so columns past w come out as zero.
__device__ inline void tile_trsm_dmma(double* P, int nrows, const double* L, const double* rinv, int w, double* winv)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp * 8 >= nrows) return;
    {
        const int b0 = (lane >> 3) * 8, q = lane & 7;
        double acc[8], ri[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i] = (i == q) ? 1.0 : 0.0; ri[i] = (b0 + i < w) ? rinv[b0 + i] : 0.0; }
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const double wm = acc[m] * ri[m];
#pragma unroll
            for (int i = m + 1; i < 8; ++i) acc[i] = fma(-L[(b0 + i) * kCLD + b0 + m], wm, acc[i]);
            winv[(b0 + m) * 8 + q] = wm;   // W_b[m][q]
        }
    }
    __syncwarp();
    const int g = lane >> 2, t = lane & 3;
    double* prow = P + (8 * warp + g) * kCLD;
    double cf[4][2];
#pragma unroll
    for (int b = 0; b < 4; ++b) { cf[b][0] = prow[8 * b + 2 * t]; cf[b][1] = prow[8 * b + 2 * t + 1]; }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        if (8 * b < w) {
            // S_b: accumulator layout -> A-operand layout through shared memory
            if (b > 0) { prow[8 * b + 2 * t] = cf[b][0]; prow[8 * b + 2 * t + 1] = cf[b][1]; __syncwarp(); }
            const double s0 = prow[8 * b + t], s1 = prow[8 * b + 4 + t];
            const double* wr = winv + (8 * b + g) * 8 + t;    // B(k = m, n = c) = W_b[c][m]
            double x0 = 0.0, x1 = 0.0;
            dmma(x0, x1, s0, wr[0]);
            dmma(x0, x1, s1, wr[4]);
            __syncwarp();
            prow[8 * b + 2 * t] = x0; prow[8 * b + 2 * t + 1] = x1;
            __syncwarp();
            const double a0 = -prow[8 * b + t], a1 = -prow[8 * b + 4 + t];
#pragma unroll
            for (int b2 = b + 1; b2 < 4; ++b2) {
                const double* lr = L + (8 * b2 + g) * kCLD + 8 * b + t;   // B(k = m, n = c) = L[8 b2 + c][8 b + m]
                dmma(cf[b2][0], cf[b2][1], a0, lr[0]);
                dmma(cf[b2][0], cf[b2][1], a1, lr[4]);
            }
        }
    }
}

// ---- one-launch tile DAG ---------------------------------------------------------------------------------------
// One CTA per lower-triangle tile (i, j) of the augmented matrix [S; g^T], launched in column-major order; all CTAs
// are co-resident (checked by the host).  The input is read-only; results go to a separate buffer that the host
// fills with an all-ones bit pattern before the launch, and THE DATA ARE THE FLAGS: a consumer polls the very words
// it needs (relaxed gpu-scope loads) until none of them is the sentinel, so a hand-over costs one store -> L2 -> load
// trip instead of data + fence + flag + data.  (Arithmetic never produces the all-ones NaN; stores canonicalise it.)
//
// Off-diagonal CTA (i, j): applies L_ik L_jk^T for k < j as the tiles appear, solves against L_jj (DMMA block solve),
// publishes L_ij; later, when x_i appears, publishes the partial product L_ij^T x_i for the backward sweep.
// Diagonal CTA (j, j) keeps a private copy of its left neighbour (j, j-1) as well, so that the critical path
//   L_{j-1,j-1}  ->  solve (j, j-1)  ->  L_jj update  ->  Cholesky of the diagonal tile
// stays inside one CTA with a single hand-over per block column; in the backward sweep it solves
// L_jj^T x_j = y_j - sum_t part(t, j) in a fixed order (t descending: deterministic) and immediately produces
// part(j, j-1) from its private tile for the next diagonal CTA.
constexpr unsigned long long kDagSentinel = ~0ull;

struct CholDag {
    const double* A;   // (n+1) x n input: reduced matrix, then the gradient as row n (read only)
    double* L;         // sentinel-initialised, contiguous: L (n+1) x n | part ntr x ntc x 32 | rinv n | xs n
    int n;
    double* xout;      // n: solution (plain copy for the kernels that follow)
    const int* go;     // optional gate
    int* fail;
    unsigned long long* trace;   // optional: 8 globaltimer stamps per CTA (diagnostics), else nullptr
};
__host__ __device__ inline size_t chol_dag_words(int n)
{
    return (size_t)(n + 1) * n + (size_t)chol_row_tiles(n) * chol_col_tiles(n) * kCT + 2 * (size_t)n;
}

__device__ __forceinline__ void dag_stamp(const CholDag& D, int k)
{
    if (D.trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        D.trace[blockIdx.x * 8 + k] = t;
    }
}
__device__ __forceinline__ unsigned long long dag_ld(const double* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double dag_poll(const double* p)
{
    unsigned long long v;
    do { v = dag_ld(p); } while (v == kDagSentinel);
    return __longlong_as_double((long long)v);
}
__device__ __forceinline__ void dag_st(double* p, double x)
{
    unsigned long long v = (unsigned long long)__double_as_longlong(x);
    if (v == kDagSentinel) v = 0x7ff8000000000000ull;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// rows x cols of a published tile (row stride n) into T, zero elsewhere in the 32 x 32 area; 256 threads, 4 words each
template <bool kLower>
__device__ __forceinline__ void dag_load_tile(double (*T)[kCLD], const double* src, int n, int rows, int cols)
{
    const int r = threadIdx.x >> 5, c = threadIdx.x & 31;   // rows r, r + 8, r + 16, r + 24
    unsigned long long v[4] = {0, 0, 0, 0};
    bool ok;
    do {
        ok = true;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int rr = r + 8 * q;
            if (rr < rows && c < cols && (!kLower || c <= rr)) {
                v[q] = dag_ld(src + (int64_t)rr * n + c);
                ok = ok && v[q] != kDagSentinel;
            }
        }
    } while (!ok);
#pragma unroll
    for (int q = 0; q < 4; ++q) T[r + 8 * q][c] = __longlong_as_double((long long)v[q]);
}
__device__ __forceinline__ void dag_load_input(double (*T)[kCLD], const double* src, int n, int rows, int cols)
{
    const int r = threadIdx.x >> 5, c = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int rr = r + 8 * q;
        T[rr][c] = (rr < rows && c < cols) ? src[(int64_t)rr * n + c] : 0.0;
    }
}

__device__ inline void chol_dag_tile(const CholDag& D)
{
    __shared__ __align__(16) double C[kCT + 8][kCLD], S[kCT][kCLD], A1[kCT][kCLD], B1[kCT][kCLD];
    __shared__ __align__(16) double s_rinv[kCT], s_x[kCT];
    __shared__ __align__(16) double s_col[4 * kCT], s_winv[kCT * 8];
    __shared__ int s_bad;
    const int tid = threadIdx.x, n = D.n;
    const int ntc = chol_col_tiles(n), ntr = chol_row_tiles(n);
    double* const Lg = D.L;
    double* const part = Lg + (size_t)(n + 1) * n;
    double* const rinvg = part + (size_t)ntr * ntc * kCT;
    double* const xs = rinvg + n;
    // block index -> (i, j), column-major over the lower triangle
    int j = 0, rem = blockIdx.x;
    while (rem >= ntr - j) { rem -= ntr - j; ++j; }
    const int i = j + rem;
    const int r0 = i * kCT, c0 = j * kCT;
    const int h = min(kCT, n + 1 - r0), w = min(kCT, n - c0);   // rows (may include the g row), columns
    const int hm = min(kCT, max(n - r0, 0));                     // matrix rows of this tile row
    dag_stamp(D, 0);
    if (tid == 0) s_bad = 0;
    dag_load_input(C, D.A + (int64_t)r0 * n + c0, n, h, w);
    for (int idx = tid; idx < 8 * kCLD; idx += blockDim.x) (&C[kCT][0])[idx] = 0.0;
    if (i != j) {
        // ---------------- off-diagonal tile ----------------
        for (int k = 0; k < j; ++k) {
            __syncthreads();
            dag_load_tile<false>(A1, Lg + (int64_t)r0 * n + k * kCT, n, h, kCT);
            dag_load_tile<false>(B1, Lg + (int64_t)c0 * n + k * kCT, n, w, kCT);
            __syncthreads();
            tile_gemm_sub(&C[0][0], &A1[0][0], &B1[0][0]);
        }
        __syncthreads();
        dag_stamp(D, 1);
        dag_load_tile<true>(B1, Lg + (int64_t)c0 * n + c0, n, w, w);
        if (tid < kCT) s_rinv[tid] = tid < w ? dag_poll(rinvg + c0 + tid) : 0.0;
        __syncthreads();
        dag_stamp(D, 2);
        tile_trsm_dmma(&C[0][0], h, &B1[0][0], s_rinv, w, s_winv);
        __syncthreads();
        {
            const int r = tid >> 5, c = tid & 31;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (r + 8 * q < h && c < w) dag_st(Lg + (int64_t)(r0 + r + 8 * q) * n + c0 + c, C[r + 8 * q][c]);
        }
        dag_stamp(D, 3);
        // backward sweep: part(i, j) = L_ij^T x_i.  (i == j + 1 is produced by the diagonal CTA i from its private copy.)
        if (i >= j + 2 && i < ntc && tid < kCT) {
            s_x[tid] = tid < hm ? dag_poll(xs + r0 + tid) : 0.0;
            __syncwarp();
            double p0 = 0.0, p1 = 0.0;
#pragma unroll 4
            for (int r = 0; r < kCT; r += 2) {
                p0 = fma(C[r][tid], s_x[r], p0);
                p1 = fma(C[r + 1][tid], s_x[r + 1], p1);
            }
            if (tid < w) dag_st(part + ((size_t)i * ntc + j) * kCT + tid, p0 + p1);
        }
        return;
    }
    // ---------------- diagonal tile ----------------
    if (j > 0) dag_load_input(S, D.A + (int64_t)r0 * n + c0 - kCT, n, h, kCT);
    for (int k = 0; k + 1 < j; ++k) {
        __syncthreads();
        dag_load_tile<false>(A1, Lg + (int64_t)r0 * n + k * kCT, n, h, kCT);
        dag_load_tile<false>(B1, Lg + (int64_t)(c0 - kCT) * n + k * kCT, n, kCT, kCT);
        __syncthreads();
        tile_gemm_sub(&C[0][0], &A1[0][0], &A1[0][0]);
        tile_gemm_sub(&S[0][0], &A1[0][0], &B1[0][0]);
    }
    __syncthreads();
    dag_stamp(D, 1);
    if (j > 0) {
        dag_load_tile<true>(B1, Lg + (int64_t)(c0 - kCT) * n + c0 - kCT, n, kCT, kCT);
        if (tid < kCT) s_rinv[tid] = dag_poll(rinvg + c0 - kCT + tid);
        __syncthreads();
        dag_stamp(D, 2);
        tile_trsm_dmma(&S[0][0], h, &B1[0][0], s_rinv, kCT, s_winv);
        __syncthreads();
        tile_gemm_sub(&C[0][0], &S[0][0], &S[0][0]);
        __syncthreads();
    }
    if (tid < 32) tile_potrf_warp(&C[0][0], w, &s_bad, s_col);
    __syncthreads();
    tile_potrf_scale(&C[0][0], w, s_rinv);
    if (h > w) {   // the g row lives in this tile (ragged last block column)
        tile_trsm_dmma(&C[w][0], 1, &C[0][0], s_rinv, w, s_winv);
        __syncthreads();
    }
    {
        const int r = tid >> 5, c = tid & 31;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int rr = r + 8 * q;
            if (rr < h && c < w && (c <= rr)) dag_st(Lg + (int64_t)(r0 + rr) * n + c0 + c, C[rr][c]);
        }
        if (tid < w) dag_st(rinvg + c0 + tid, s_rinv[tid]);
        if (tid == 0 && s_bad) *D.fail = 1;
    }
    dag_stamp(D, 3);
    // While the factorisation moves on, invert the factor for the backward sweep: X = L_jj^-T (X L^T = I, same block
    // solve as the off-diagonal tiles), so that x_j is one 32 x 32 matrix-vector product instead of a 32-step
    // dependent chain.  Not for the last block column, whose backward step starts right away.
    const bool use_inv = j != ntc - 1;
    if (use_inv) {
        const int r = tid >> 5, c = tid & 31;
#pragma unroll
        for (int q = 0; q < 4; ++q) A1[r + 8 * q][c] = (r + 8 * q == c) ? 1.0 : 0.0;
        __syncthreads();
        tile_trsm_dmma(&A1[0][0], kCT, &C[0][0], s_rinv, w, s_winv);
        __syncthreads();
    }
    // ---- backward substitution: x_j = L_jj^-T (y_j - sum_{t > j} part(t, j)), warp 0 ----------------------------
    if (tid >= 32) return;
    const int lane = tid;
    const int ig = n / kCT;                       // row tile that holds the g row (row n)
    double yv = 0.0;
    if (lane < w) {
        yv = (ig == j) ? C[w][lane] : dag_poll(Lg + (int64_t)n * n + c0 + lane);
        for (int t = ntc - 1; t > j; --t) yv -= dag_poll(part + ((size_t)t * ntc + j) * kCT + lane);
    }
    dag_stamp(D, 5);
    if (use_inv) {
        s_x[lane] = yv;
        __syncwarp();
        const double2* xr = reinterpret_cast<const double2*>(&A1[lane][0]);
        const double2* yr = reinterpret_cast<const double2*>(s_x);
        double x0 = 0.0, x1 = 0.0;
#pragma unroll
        for (int c = 0; c < kCT / 2; ++c) {
            const double2 a = xr[c], b = yr[c];
            x0 = fma(a.x, b.x, x0);
            x1 = fma(a.y, b.y, x1);
        }
        yv = x0 + x1;
        __syncwarp();
    } else {
        const double myrinv = lane < w ? s_rinv[lane] : 1.0;
        for (int c = w - 1; c >= 0; --c) {
            const double xc = __shfl_sync(0xffffffffu, yv * myrinv, c);
            if (lane == c) yv = xc;
            if (lane < c) yv -= C[c][lane] * xc;
        }
    }
    s_x[lane] = lane < w ? yv : 0.0;
    __syncwarp();
    if (j > 0) {   // part(j, j-1) = L_{j,j-1}^T x_j from the private copy: the only hand-over on the backward chain
        double p0 = 0.0, p1 = 0.0;
#pragma unroll 4
        for (int r = 0; r < kCT; r += 2) {
            p0 = fma(S[r][lane], r < hm ? s_x[r] : 0.0, p0);
            p1 = fma(S[r + 1][lane], r + 1 < hm ? s_x[r + 1] : 0.0, p1);
        }
        dag_st(part + ((size_t)j * ntc + j - 1) * kCT + lane, p0 + p1);
    }
    if (lane < w) { dag_st(xs + c0 + lane, yv); D.xout[c0 + lane] = yv; }
    dag_stamp(D, 6);
}

// ---- banded solver ----------------------------------------------------------------------------------------------
// Camera rigs overlap with their neighbours only, so under the camera numbering the reduced system is banded: block
// (A, B) is structurally zero unless |A - B| <= m.  The host measures m from the record layout (and agrees on the
// maximum over the ranks); for 6 (m + 1) <= 32 the solve is an LDL^T factorisation of the band by ONE warp -- the
// pivot chain is the cost either way (n dependent pivots), but the band needs no tile hand-overs and touches
// (w + 1) n numbers instead of n^2 / 2.
//   * lane = row mod 32; an active row keeps its window  a[k] = A[row][j + k]  in registers, shifted by one column
//     per pivot exactly like tile_potrf_warp; rows enter from the shared-memory band w pivots before they become the
//     pivot row and retire after it;
//   * the right-hand side rides along as one more column (rhs_i -= (a_ij / d_j) rhs_j);
//   * no square roots: with u_ij the unscaled entries,  x_j = (rhs_j - sum_{i>j} u_ij x_i) / d_j;
//   * the backward sweep keeps one accumulator per lane (= row) and handles two rows per shuffle round trip.
// band: (n + 1) x NW doubles, band[r][k] = A[r][r - w + k] (zero for negative columns), w = NW - 1; overwritten with the
// unscaled factor (same positions, d_j on the diagonal).  rhs: n + 1 doubles, overwritten with the solution x.
// Returns non-zero (in every lane) if a pivot is outside [1e-200, 1e200].
#ifdef MCCBA_BAND_DBG
__device__ long long g_band_ts[4];
#endif
// predicated shared-memory loads (no branch: a divergent branch per pivot costs the whole warp ~30 cycles)
__device__ __forceinline__ void lds_v2_if(double& x0, double& x1, unsigned addr, int pred)
{
    asm volatile("{ .reg .pred p; setp.ne.s32 p, %3, 0; @p ld.shared.v2.f64 {%0, %1}, [%2]; }" : "+d"(x0), "+d"(x1) : "r"(addr), "r"(pred) : "memory");
}
__device__ __forceinline__ void lds_if(double& x0, unsigned addr, int pred)
{
    asm volatile("{ .reg .pred p; setp.ne.s32 p, %2, 0; @p ld.shared.f64 %0, [%1]; }" : "+d"(x0) : "r"(addr), "r"(pred) : "memory");
}

// Column j of the active rows as every lane sees it: pivot d, right-hand side of the pivot row, the entries of the w
// rows below the pivot, and this lane's multiplier t = -a_ij / d_j (0 unless its row lies below the pivot in the band).
template <int NW>
struct BandCol {
    double c[NW - 1];
    double d, rp, t;
};

// colbuf layout per parity buffer: [64 column entries | 64 right-hand sides], indexed by row & 31 and duplicated at +32
// so that a window never wraps; 16-byte loads from the even offset at or below the pivot slot.
template <int NW, bool kOdd>
__device__ __forceinline__ void band_fetch(const double* cb, int pj, BandCol<NW>& col)
{
    const unsigned vaddr = (unsigned)__cvta_generic_to_shared(cb + (pj & ~1));
    double v[NW + 2];
#pragma unroll
    for (int p = 0; p < (NW + 2) / 2; ++p)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[2 * p]), "=d"(v[2 * p + 1]) : "r"(vaddr + 16u * p) : "memory");
    col.d = kOdd ? v[1] : v[0];
#pragma unroll
    for (int k = 0; k < NW - 1; ++k) col.c[k] = v[(kOdd ? 2 : 1) + k];
    col.rp = cb[64 + pj];
}

// One pivot of the banded factorisation, software-pipelined like potrf_step: the column of pivot j + 1 is published as
// soon as its first update fma is done, and its broadcast and reciprocal chain overlap the remaining work of pivot j
// (w - 1 update fmas, the factor store, retiring the pivot row, fetching the row that becomes active at pivot j + 2
// one position to the right so that the idle shift of pivot j + 1 moves it into place).  kOdd = parity of j + 1.
template <int NW, bool kOdd>
__device__ __forceinline__ void band_step(double* band, double* rhs, double* colbuf, int lane, int j, int n, double (&a)[NW + 1],
                                          double& r, int& row, const BandCol<NW>& cur, BandCol<NW>& nxt, int& bad)
{
    constexpr int w = NW - 1;
    const double a0 = a[0];
    const double a0n = pin_fma(cur.t, cur.c[0], a[1]);
    const double rn = pin_fma(cur.t, cur.rp, r);
    double* cbn = colbuf + (kOdd ? 128 : 0);
    cbn[lane] = a0n; cbn[lane + 32] = a0n;
    cbn[64 + lane] = rn;
    __syncwarp();
    band_fetch<NW, kOdd>(cbn, (j + 1) & 31, nxt);
    double y = pin_rcp_seed(nxt.d);
    // ---- work of pivot j in the shadow of the chain ----
    const bool in_band = row >= j && row <= j + w && row < n;      // the pivot row and the rows below it inside the band
    if (in_band) band[row * NW + (j - (row - w))] = a0;             // unscaled factor entry u_{row, j}; d_j for row == j
    const bool piv = row == j;
    if (piv) rhs[j] = rn;                                           // the pivot row retires with its final right-hand side
    if (!(cur.d > 1e-200 && cur.d < 1e200)) bad = 1;
    double e = pin_nfma(nxt.d, y, 1.0);
    constexpr int nF = w - 1;
#pragma unroll
    for (int k = 1; k <= nF / 3; ++k) a[k] = pin_fma(cur.t, cur.c[k], a[k + 1]);
    y = pin_fma(y, e, y);
#pragma unroll
    for (int k = nF / 3 + 1; k <= (2 * nF) / 3; ++k) a[k] = pin_fma(cur.t, cur.c[k], a[k + 1]);
    e = pin_nfma(nxt.d, y, 1.0);
#pragma unroll
    for (int k = (2 * nF) / 3 + 1; k <= nF; ++k) a[k] = pin_fma(cur.t, cur.c[k], a[k + 1]);
    y = pin_fma(y, e, y);
    a[0] = a0n;
    a[w] = a[w + 1];     // the spare slot only shifts
    a[w + 1] = 0.0;
    r = rn;
    row += piv ? 32 : 0;
    const int re = j + 2 + w;
    const int ent = (lane == (re & 31) && re < n) ? 1 : 0;
    row = ent ? re : row;
    const bool below = row > j + 1 && row <= j + 1 + w && row < n;
    nxt.t = below ? -a0n * y : 0.0;
    // fetch row re one position to the right: a[1 .. NW] <- band[re][0 .. w]  (used one step from now)
    const unsigned baddr = (unsigned)__cvta_generic_to_shared(band + (ent ? re : 0) * NW);
#pragma unroll
    for (int p = 0; p < NW / 2; ++p) {
        double q0 = a[2 * p + 1], q1 = a[2 * p + 2];
        lds_v2_if(q0, q1, baddr + 16u * p, ent);
        a[2 * p + 1] = q0;
        a[2 * p + 2] = q1;
    }
    lds_if(r, (unsigned)__cvta_generic_to_shared(rhs + (ent ? re : 0)), ent);
}

template <int NW>
__device__ inline int band_ldlt_solve_warp(double* band, double* rhs, const int n, double* colbuf /* 256 doubles, 16-byte aligned */,
                                           double* pinv /* 3 * ((n + 1) / 2) doubles */)
{
    constexpr int w = NW - 1;
    const int lane = threadIdx.x & 31;
#ifdef MCCBA_BAND_DBG
    if (lane == 0) g_band_ts[0] = clock64();
#endif
    double a[NW + 1];   // a[k] = A[row][j + k] at pivot j, plus one spare slot for the row waiting to become active
    double r = 0.0;
    int row = lane;     // the row this lane holds (or will hold next)
    // rows 0 .. w enter before pivot 0 (window starts at column 0); row w + 1 waits one position to the right
    {
        const bool in = row <= w && row < n;
        const bool nx = row == w + 1 && row < n;
#pragma unroll
        for (int k = 0; k < NW + 1; ++k) {
            double x = 0.0;
            if (in && k <= row) x = band[row * NW + (w - row) + k];
            if (nx && k >= 1) x = band[row * NW + k - 1];
            a[k] = x;
        }
        r = (in || nx) ? rhs[row] : 0.0;
        if (!in && !nx && row <= w + 1) row += 32;   // n too small: nothing to hold
    }
    for (int q = lane; q < 256; q += 32) colbuf[q] = 0.0;
    __syncwarp();
    int bad = 0;
    BandCol<NW> c0, c1;
    {   // prologue: publish and fetch column 0
        colbuf[lane] = a[0]; colbuf[lane + 32] = a[0];
        colbuf[64 + lane] = r;
        __syncwarp();
        band_fetch<NW, false>(colbuf, 0, c0);
        const bool below = row > 0 && row <= w && row < n;
        c0.t = below ? -a[0] * pivot_rcp(c0.d) : 0.0;
    }
#pragma unroll 1
    for (int j = 0; j < n; j += 2) {
        band_step<NW, true>(band, rhs, colbuf, lane, j, n, a, r, row, c0, c1, bad);
        if (j + 1 < n) band_step<NW, false>(band, rhs, colbuf, lane, j + 1, n, a, r, row, c1, c0, bad);
    }
    __syncwarp();
#ifdef MCCBA_BAND_DBG
    if (lane == 0) g_band_ts[1] = clock64();
#endif
    // Backward sweep, two rows per step.  With inv_q = 1 / d_q and acc_q = rhs_q - sum_{i > pair} u_iq x_i, the pair
    // (j, j+1) is a 2 x 2 back-substitution   x_j+1 = inv_j+1 acc_j+1,  x_j = inv_j acc_j - (u_j+1,j inv_j inv_j+1) acc_j+1,
    // so one round of shuffles serves two rows and the dependent chain is halved.  The row count is padded to even with
    // an identity row (band and rhs have room for it).
    const int ne = n + (n & 1);
    if (ne != n && lane == 0) {
        for (int k = 0; k < NW; ++k) band[n * NW + k] = k == w ? 1.0 : 0.0;
        rhs[n] = 0.0;
    }
    __syncwarp();
    for (int q = lane; q < n; q += 32) band[q * NW + w] = 1.0 / band[q * NW + w];   // 1 / d_q, off the chain
    __syncwarp();
    for (int q = lane; 2 * q < ne; q += 32) {
        const double i0 = band[(2 * q) * NW + w], i1 = band[(2 * q + 1) * NW + w];
        pinv[3 * q] = i0;
        pinv[3 * q + 1] = -band[(2 * q + 1) * NW + w - 1] * i0 * i1;
        pinv[3 * q + 2] = i1;
    }
    __syncwarp();
    const int last = ne - 1;
    int rj = last - ((last - lane) & 31);
    double acc = rj >= 0 ? rhs[rj] : 0.0;
    int rn = rj - 32;
    double nacc = rn >= 0 ? rhs[rn] : 0.0;
    auto fetch_u = [&](int j, int rrow, double& uu0, double& uu1) {
        const int o0 = rrow - (j - w), o1 = rrow - (j + 1 - w);
        const int l0 = (j >= 0 && rrow >= 0 && rrow < j && o0 >= 0) ? 1 : 0;
        const int l1 = (j >= 0 && rrow >= 0 && rrow < j && o1 >= 0) ? 1 : 0;
        uu0 = 0.0; uu1 = 0.0;
        lds_if(uu0, (unsigned)__cvta_generic_to_shared(band + (l0 ? j * NW + o0 : 0)), l0);
        lds_if(uu1, (unsigned)__cvta_generic_to_shared(band + (l1 ? (j + 1) * NW + o1 : 0)), l1);
    };
    double u0 = 0.0, u1 = 0.0, p00 = 0.0, p01 = 0.0, p11 = 0.0;
    fetch_u(ne - 2, rj, u0, u1);
    if (ne >= 2) { p00 = pinv[3 * ((ne - 2) >> 1)]; p01 = pinv[3 * ((ne - 2) >> 1) + 1]; p11 = pinv[3 * ((ne - 2) >> 1) + 2]; }
#pragma unroll 1
    for (int j = ne - 2; j >= 0; j -= 2) {
        const double A0 = __shfl_sync(0xffffffffu, acc, j & 31);
        const double A1 = __shfl_sync(0xffffffffu, acc, (j + 1) & 31);
        // bookkeeping for the next pair in the shadow of the shuffles
        const bool own0 = rj == j, own1 = rj == j + 1, own = own0 || own1;
        const double acc_own = nacc;
        const int rj_new = own ? rn : rj;
        const int rn_new = rn - (own ? 32 : 0);
        const int ok = (own && rn_new >= 0) ? 1 : 0;
        lds_if(nacc, (unsigned)__cvta_generic_to_shared(rhs + (ok ? rn_new : 0)), ok);
        double nu0, nu1;
        fetch_u(j - 2, rj_new, nu0, nu1);
        double q00 = 0.0, q01 = 0.0, q11 = 0.0;
        if (j >= 2) { q00 = pinv[3 * ((j - 2) >> 1)]; q01 = pinv[3 * ((j - 2) >> 1) + 1]; q11 = pinv[3 * ((j - 2) >> 1) + 2]; }
        // consume
        const double x0 = fma(p00, A0, p01 * A1), x1 = p11 * A1;
        if (own0) rhs[j] = x0;
        if (own1) rhs[j + 1] = x1;
        acc = own ? acc_own : fma(-u1, x1, fma(-u0, x0, acc));
        rj = rj_new; rn = rn_new; u0 = nu0; u1 = nu1; p00 = q00; p01 = q01; p11 = q11;
    }
    __syncwarp();
#ifdef MCCBA_BAND_DBG
    if (lane == 0) g_band_ts[2] = clock64();
#endif
    return bad;
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
