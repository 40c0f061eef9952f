// mccba_bcr.cuh -- block cyclic reduction of the block-banded reduced camera system (replaces the Eigen conjugate
// gradient of src/multicalib.cpp:565-592 on the Schur-reduced system; included by mccba_kernels.cuh).
//
// Rig cameras overlap with their neighbours only, so under the camera numbering the reduced system S (6 x 6 blocks, one
// per camera pair) is block-banded with block bandwidth m.  Grouping m consecutive cameras into one super-block of
// B = 6 m rows makes S block-TRIDIAGONAL:  D_0 C_1^T | C_1 D_1 C_2^T | ...  (Nb = ceil(n / B) super-blocks).
// A serial pivot chain (n dependent pivots) is the wrong algorithm for that; cyclic reduction eliminates every other
// super-block at once -- log2(Nb) levels of INDEPENDENT B x B eliminations:
//   level with stride s:  for every i = s (mod 2s), neighbours a = i - s, b = i + s:
//     phase 1 (one warp per eliminated block, lane = column of [D_i | C_a | C_b | r_i]): Gauss-Jordan without pivoting
//             (S is SPD; the pivots are those of LDL^T) turns it into [I | F_a | F_b | f],  F = D_i^-1 C;
//     phase 2 (one warp per surviving block j = 0 (mod 2s)):  D_j -= C^T F from both eliminated neighbours,
//             r_j -= C^T f, and the new coupling to j - 2s is  -C_b^T F_a.
//   Back-substitution runs the levels in reverse:  x_i = f - F_a x_a - F_b x_b  (plain matrix-vector products).
// Everything lives in shared memory of ONE CTA; two __syncthreads per level, no atomics, fixed summation order
// (bit-stable run to run and identical on every rank).  At n = 378, m = 1 (config #5): 63 blocks of 6, 6 levels.
#pragma once
#include <cuda_runtime.h>

namespace mccba {

__host__ __device__ inline int bcr_blocks(int n, int B) { return (n + B - 1) / B; }
// shared memory (doubles): Dg[Nb][B][B] | Lo[Nb][B][B] | Tmp[ceil(Nb/2)][B][B] | rhs[Nb * B] | pivot scratch[32 warps][2][B]
__host__ __device__ inline size_t bcr_smem_bytes(int n, int B)
{
    const size_t nb = (size_t)bcr_blocks(n, B);
    return sizeof(double) * ((2 * nb + (nb + 1) / 2) * (size_t)B * B + nb * (size_t)B + 64 * (size_t)B + 8);
}
template <int B>
struct BcrCfg {
    // one warp per block of the widest level where the registers allow it (the elimination keeps ceil((3B+1)/32) columns of
    // B rows per lane): B = 6 -> 32 warps, 63 blocks at config #5 are eliminated / updated in one round per level
    static constexpr int kThreads = B <= 6 ? 1024 : (B <= 12 ? 512 : 256);
};

__device__ __forceinline__ double bcr_rcp(double d)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    return fma(y, e, y);
}

// phase 1: eliminate super-block i (neighbours a, b; -1 = none).  F_a -> Dg[i], F_b -> Tmp[slot], f -> rhs[i].
template <int B>
__device__ __forceinline__ int bcr_eliminate(double* Dg, const double* Lo, double* Tmp, double* rhs, int i, int a, int b,
                                             int slot, int lane, double* piv /* this warp's 2 x B scratch */)
{
    constexpr bool live = true;
    constexpr int NCOL = 3 * B + 1, CPL = (NCOL + 31) / 32;
    double col[CPL][B];
    // Branch-free staging: every lane turns its column index into (pointer, stride, valid) once and issues B predicated
    // loads back to back.  (Branching per element serialised four divergent paths per row: 1700 cycles for 19 columns.)
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        const int j = lane + 32 * c;
        const double* src = Dg + (size_t)i * B * B + j;       // j < B: column j of D_i
        int stride = B;
        bool valid = j < B;
        if (j >= B && j < 2 * B) { src = Lo + (size_t)i * B * B + (j - B); valid = a >= 0; }
        if (j >= 2 * B && j < 3 * B) { src = Lo + ((size_t)(b >= 0 ? b : 0) * B + (j - 2 * B)) * B; stride = 1; valid = b >= 0; }   // C_b = Lo[b]^T
        if (j == 3 * B) { src = rhs + (size_t)i * B; stride = 1; valid = true; }
#pragma unroll
        for (int r = 0; r < B; ++r) col[c][r] = valid ? src[r * stride] : 0.0;
    }
    int bad = 0;
#ifdef MCCBA_BCR_TS
    if (slot == 777 && lane == 0) g_bcr_ts[40] = clock64();
#endif
#pragma unroll
    for (int k = 0; k < B; ++k) {
        // the pivot column (column k lives in lane k, B < 32) goes to all lanes through the warp's scratch row: B broadcast
        // loads instead of 2 B shuffles (with 16 warps eliminating at once the shuffle unit was the bottleneck);
        // double-buffered, so one __syncwarp per pivot is enough
        double pk[B];
        double* pv = piv + (k & 1) * B;
        if (lane == k) {
#pragma unroll
            for (int r = 0; r < B; ++r) pv[r] = col[0][r];
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < B; ++r) pk[r] = pv[r];
        const double d = pk[k];
        if (!(d > 0.0) || !isfinite(d)) bad = 1;
        const double inv = bcr_rcp(d);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const double t = col[c][k] * inv;
#pragma unroll
            for (int r = 0; r < B; ++r)
                if (r != k) col[c][r] = fma(-pk[r], t, col[c][r]);
            col[c][k] = t;
        }
#ifdef MCCBA_BCR_TS
        if (slot == 777 && lane == 0) g_bcr_ts[41 + k] = clock64();
#endif
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        const int j = lane + 32 * c;
        double* dst = Dg + (size_t)i * B * B + (j - B);                                      // F_a
        bool valid = live && j >= B && j < 2 * B;
        int stride = B;
        if (j >= 2 * B && j < 3 * B) { dst = Tmp + (size_t)(slot == 777 ? 0 : slot) * B * B + (j - 2 * B); valid = live; }   // F_b (777: timing hook of scripts/ubench/bcr_ts.cu)
        if (j == 3 * B) { dst = rhs + (size_t)i * B; stride = 1; valid = live; }            // f
#pragma unroll
        for (int r = 0; r < B; ++r)
            if (valid) dst[r * stride] = col[c][r];
    }
    return bad;
}

// phase 2: surviving super-block j takes the Schur updates of its eliminated neighbours p = j - s and q = j + s
// (-1 = none); sp, sq = their Tmp slots.  Afterwards F_b(q) moves from Tmp into Lo[q] (nobody else reads Lo[q]).
template <int B>
__device__ __forceinline__ void bcr_update(double* Dg, double* Lo, const double* Tmp, double* rhs, int j, int p, int q, int sp,
                                           int sq, int lane)
{
    constexpr int NE = (B * B + 31) / 32;
    double nl[NE], nd[NE];
    const double* __restrict__ Lj = Lo + (size_t)j * B * B;
    const double* __restrict__ Lq = Lo + (size_t)(q >= 0 ? q : 0) * B * B;
    const double* __restrict__ Fbp = Tmp + (size_t)(sp >= 0 ? sp : 0) * B * B;
    const double* __restrict__ Fbq = Tmp + (size_t)(sq >= 0 ? sq : 0) * B * B;
    const double* __restrict__ Fap = Dg + (size_t)(p >= 0 ? p : 0) * B * B;
    const double* __restrict__ Faq = Dg + (size_t)(q >= 0 ? q : 0) * B * B;
    // all loads and sums first (no store in between: the loads of both element slots are in flight together) ...
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int idx = lane + 32 * e;
        nl[e] = 0.0; nd[e] = 0.0;
        if (idx < B * B) {
            const int u = idx / B, v = idx % B;
            double acc = 0.0, acc2 = 0.0, accl = 0.0;
            if (p >= 0) {
#pragma unroll
                for (int t = 0; t < B; ++t) {
                    const double l = Lj[u * B + t];
                    acc = fma(l, Fbp[t * B + v], acc);
                    accl = fma(l, Fap[t * B + v], accl);
                }
            }
            if (q >= 0) {
#pragma unroll
                for (int t = 0; t < B; ++t) acc2 = fma(Lq[t * B + u], Faq[t * B + v], acc2);
            }
            nd[e] = acc + acc2;
            nl[e] = -accl;
        }
    }
    double racc = 0.0;
    if (lane < B) {
        const int u = lane;
        double acc = 0.0, acc2 = 0.0;
        if (p >= 0) {
#pragma unroll
            for (int t = 0; t < B; ++t) acc = fma(Lj[u * B + t], rhs[(size_t)p * B + t], acc);
        }
        if (q >= 0) {
#pragma unroll
            for (int t = 0; t < B; ++t) acc2 = fma(Lq[t * B + u], rhs[(size_t)q * B + t], acc2);
        }
        racc = acc + acc2;
    }
    double fb[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int idx = lane + 32 * e;
        fb[e] = (q >= 0 && idx < B * B) ? Fbq[idx] : 0.0;
    }
    __syncwarp();   // ... every lane is done with Lo[j] and Lo[q]; now the stores
    if (lane < B) rhs[(size_t)j * B + lane] -= racc;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int idx = lane + 32 * e;
        if (idx < B * B) {
            Dg[(size_t)j * B * B + idx] -= nd[e];
            if (p >= 0) Lo[(size_t)j * B * B + idx] = nl[e];     // coupling to j - 2s
            if (q >= 0) Lo[(size_t)q * B * B + idx] = fb[e];     // F_b(q), for the back-substitution
        }
    }
}

// The whole solve; every thread of the CTA calls it (blockDim.x a multiple of 32).  On entry Dg / Lo / rhs hold the
// system, on return rhs holds the solution.  Returns non-zero (to every thread) if a pivot was not positive.
template <int B>
__device__ inline int bcr_solve_cta(double* Dg, double* Lo, double* Tmp, double* rhs, int Nb)
{
    __shared__ int s_bad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double* piv = rhs + (size_t)Nb * B + (size_t)warp * 2 * B;   // behind rhs (bcr_smem_bytes reserves 32 warps x 2 x B)
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
#ifdef MCCBA_BCR_TS
    int ts_i = 3;
#define BCR_STAMP() do { if (threadIdx.x == 0 && ts_i < 63) g_bcr_ts[ts_i] = clock64(); ++ts_i; } while (0)
#else
#define BCR_STAMP() do { } while (0)
#endif
    int s = 1;
    for (; s < Nb; s <<= 1) {
        const int n_el = (Nb - s + 2 * s - 1) / (2 * s);            // eliminated: s, 3s, 5s, ... < Nb
        for (int e = warp; e < n_el; e += nwarp) {
            const int i = s + 2 * s * e;
            const int b = i + s < Nb ? i + s : -1;
            if (bcr_eliminate<B>(Dg, Lo, Tmp, rhs, i, i - s, b, e, lane, piv) && lane == 0) s_bad = 1;
        }
        __syncthreads();
        BCR_STAMP();
        const int n_sv = (Nb + 2 * s - 1) / (2 * s);               // surviving: 0, 2s, 4s, ... < Nb
        for (int e = warp; e < n_sv; e += nwarp) {
            const int j = 2 * s * e;
            const int p = j >= s ? j - s : -1, q = j + s < Nb ? j + s : -1;
            if (p < 0 && q < 0) continue;
            bcr_update<B>(Dg, Lo, Tmp, rhs, j, p, q, p >= 0 ? (p - s) / (2 * s) : -1, q >= 0 ? (q - s) / (2 * s) : -1, lane);
        }
        __syncthreads();
        BCR_STAMP();
    }
    if (warp == 0) {
        if (bcr_eliminate<B>(Dg, Lo, Tmp, rhs, 0, -1, -1, 0, lane, piv) && lane == 0) s_bad = 1;
    }
    __syncthreads();
    BCR_STAMP();
    for (s >>= 1; s >= 1; s >>= 1) {
        const int n_el = (Nb - s + 2 * s - 1) / (2 * s);
        for (int w = threadIdx.x; w < n_el * B; w += blockDim.x) {
            const int e = w / B, u = w - e * B;
            const int i = s + 2 * s * e, a = i - s, b = i + s < Nb ? i + s : -1;
            double acc = rhs[(size_t)i * B + u];
            const double* Fa = Dg + ((size_t)i * B + u) * B;
#pragma unroll
            for (int t = 0; t < B; ++t) acc = fma(-Fa[t], rhs[(size_t)a * B + t], acc);
            if (b >= 0) {
                const double* Fb = Lo + ((size_t)i * B + u) * B;
#pragma unroll
                for (int t = 0; t < B; ++t) acc = fma(-Fb[t], rhs[(size_t)b * B + t], acc);
            }
            rhs[(size_t)i * B + u] = acc;
        }
        __syncthreads();
        BCR_STAMP();
    }
    return s_bad;
}

// Stage the system from global memory.  packed != 0: A = [band (n x NW, band[r][c - r + NW - 1], lower triangle) | g]
// (what reduce_records writes and the exchange moves), NW = B + 6;  packed == 0: A = [S (n x n) | g] dense (test hook).
// Rows beyond n pad the last super-block with the identity.
template <int B>
__device__ inline void bcr_stage(const double* __restrict__ A, int n, int packed, double* Dg, double* Lo, double* rhs, int Nb,
                                 int skip = 0 /* leading threads of the CTA that do not take part */)
{
    const int tid = (int)threadIdx.x - skip, nthr = (int)blockDim.x - skip;
    constexpr int NW = B + 6, w = NW - 1;
    auto at = [&](int r, int c) -> double {      // r >= c, both < n
        if (packed) return (r - c <= w) ? A[(size_t)r * NW + (c - r + w)] : 0.0;
        return A[(size_t)r * n + c];
    };
    const int total = Nb * B * B;
    // kU elements per thread and trip, all their loads issued before the first store: the staging is a handful of L2 round
    // trips, and one element per trip serialised them (7 300 cycles for 2 268 elements on 992 threads)
    constexpr int kU = 4;
    for (int base = tid; base < total; base += nthr * kU) {
        double v[kU], l[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int idx = base + u * nthr;
            v[u] = 0.0; l[u] = 0.0;
            if (idx < total) {
                const int I = idx / (B * B), pq = idx - I * B * B, p = pq / B, q = pq - p * B;
                const int r = I * B + p, c = I * B + q;
                if (r >= n || c >= n) v[u] = (p == q) ? 1.0 : 0.0;
                else v[u] = r >= c ? at(r, c) : at(c, r);
                if (I > 0 && r < n) l[u] = at(r, (I - 1) * B + q);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int idx = base + u * nthr;
            if (idx < total) { Dg[idx] = v[u]; Lo[idx] = l[u]; }
        }
    }
    const size_t goff = packed ? (size_t)n * NW : (size_t)n * n;
    for (int idx = tid; idx < Nb * B; idx += nthr) rhs[idx] = idx < n ? A[goff + idx] : 0.0;
}

}  // namespace mccba
