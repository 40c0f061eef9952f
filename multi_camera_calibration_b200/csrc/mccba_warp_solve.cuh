// mccba_warp_solve.cuh -- Gaussian elimination with partial pivoting of a small dense system by ONE warp (lane = row).
// Used for the bordered (m + 1) x (m + 1) systems of the omnidir::calibrate loop (m <= 10) and of the omnidir stereo loop
// (m <= 26): the reference inverts (J^T J + eps 11^T) with cv::invert; here the rank-one term is a bordered unknown and the
// system is indefinite, hence pivoting.  Elimination order, pivot rule (first row of maximal modulus) and the order of
// every sum are those of the textbook serial algorithm; the rows just live in different lanes.  Rows are exchanged and the
// pivot row is broadcast through shared memory (one store / load round per step instead of N double shuffles).
#pragma once
#include <cuda_runtime.h>

namespace mccba {

// B: row `lane` of the N x N matrix (lanes >= N hold nothing), rhs: its right-hand side.  Rows and columns from Q on must be
// padded with the identity (B[c] = lane == c, rhs = 0), so that all N steps run unconditionally and every shuffle sits in
// straight-line code.  sm: 2 (N + 1) doubles of shared memory owned by the warp.  x: the solution, in every lane.
// Returns true (in every lane) if a pivot among the first Q rows is zero or not finite.
template <int N>
__device__ __forceinline__ bool warp_gauss_solve(double (&B)[N], double& rhs, int lane, int Q, double* sm, double (&x)[N])
{
    static_assert(N <= 32, "one row per lane");
    double* rowA = sm;              // the pivot row of the step (row piv before the exchange)
    double* rowB = sm + (N + 1);    // row k before the exchange
    bool fail = false;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const bool live = lane >= k && lane < N;
        fail = fail || (live && lane < Q && !isfinite(B[k]));
        double v = live ? fabs(B[k]) : -1.0;
        if (!(v == v)) v = -1.0;
        int idx = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        fail = fail || (k < Q && !(v > 0.0));
        const int piv = idx;
        __syncwarp();               // the previous step's reads of the two rows are over
        if (lane == piv) {
#pragma unroll
            for (int j = 0; j < N; ++j) rowA[j] = B[j];
            rowA[N] = rhs;
        }
        if (lane == k) {
#pragma unroll
            for (int j = 0; j < N; ++j) rowB[j] = B[j];
            rowB[N] = rhs;
        }
        __syncwarp();
        if (lane == k) {            // rows k and piv change places (nothing happens when piv == k)
#pragma unroll
            for (int j = 0; j < N; ++j) B[j] = rowA[j];
            rhs = rowA[N];
        } else if (lane == piv) {
#pragma unroll
            for (int j = 0; j < N; ++j) B[j] = rowB[j];
            rhs = rowB[N];
        }
        const double fct = B[k] / rowA[k];
        const bool below = lane > k && lane < N;
#pragma unroll
        for (int j = k; j < N; ++j) {
            const double pj = rowA[j];
            if (below) B[j] -= fct * pj;
        }
        if (below) rhs -= fct * rowA[N];
    }
    fail = __any_sync(0xffffffffu, fail);
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        double s = rhs;
#pragma unroll
        for (int j = i + 1; j < N; ++j) s -= B[j] * x[j];
        s = s / B[i];
        x[i] = __shfl_sync(0xffffffffu, s, i);
    }
    return fail;
}

}  // namespace mccba
