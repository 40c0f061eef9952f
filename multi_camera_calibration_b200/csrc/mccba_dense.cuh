// mccba_dense.cuh -- dense SPD solve of the reduced camera system S dc = g (n_s = 6(nC-1); 378 for 64 cameras).
//
// Replaces the reference's Eigen ConjugateGradient on the full P x P system (src/multicalib.cpp:565-592): after
// the Schur complement only the camera unknowns are left, and a direct factorisation gives the exact solution the CG
// iterates converge to.
//
// Tiled Cholesky (tile 32) on the augmented matrix A = [S; g^T] ((n+1) x n, row-major, ld = n, exactly the exchange buffer
// layout), so y = L^-1 g falls out of the factorisation as the last row: ONE launch, one CTA per lower tile, all
// co-resident, hand-overs through self-validating words (chol_dag_tile below).  This is the solver of DENSE camera graphs;
// block-banded ones (rig cameras that overlap with their neighbours only) go through the block cyclic reduction of
// mccba_bcr.cuh.  Only the lower triangle of S is read.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mccba {

constexpr int kCT = 32;        // tile
constexpr int kCLD = kCT + 2;  // padded shared-memory row stride (doubles): 16-byte aligned rows, 2 wavefronts per fragment load

__host__ __device__ inline int chol_col_tiles(int n) { return (n + kCT - 1) / kCT; }
__host__ __device__ inline int chol_row_tiles(int n) { return (n + 1 + kCT - 1) / kCT; }

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


}  // namespace mccba
