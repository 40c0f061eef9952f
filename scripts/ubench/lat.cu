// Latency microbenchmarks for the pivot chain of the tiled Cholesky (development aid, not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../multi_camera_calibration_b200/csrc/mccba_dense.cuh"
#include "tile_variants.cuh"
using namespace mccba;

__global__ void lat_kernel(double* out, long long* cyc, double seed)
{
    __shared__ __align__(16) double buf[64];
    const int lane = threadIdx.x;
    double x = seed + lane * 1e-3;
    long long t0, t1;
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = fma(x, 1.0000001, 1e-9);
    t1 = clock64();
    if (lane == 0) cyc[0] = (t1 - t0) / 64;
    // 2. dependent DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = x * 1.0000001;
    t1 = clock64();
    if (lane == 0) cyc[1] = (t1 - t0) / 64;
    // 3. fast_rcp chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) x = pivot_rcp(x) + 1.5;
    t1 = clock64();
    if (lane == 0) cyc[2] = (t1 - t0) / 16;
    // 4. fast_rsqrt chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) x = fast_rsqrt(x) + 1.5;
    t1 = clock64();
    if (lane == 0) cyc[3] = (t1 - t0) / 16;
    // 5. library 1/x and rsqrt
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) x = 1.0 / x + 1.5;
    t1 = clock64();
    if (lane == 0) cyc[4] = (t1 - t0) / 16;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) x = rsqrt(x) + 1.5;
    t1 = clock64();
    if (lane == 0) cyc[5] = (t1 - t0) / 16;
    // 6. double shuffle chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64();
    if (lane == 0) cyc[6] = (t1 - t0) / 32;
    // 7. STS -> syncwarp -> LDS round trip
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        buf[lane] = x;
        __syncwarp();
        x = buf[(lane + 1) & 31];
        __syncwarp();
    }
    t1 = clock64();
    if (lane == 0) cyc[7] = (t1 - t0) / 32;
    // 8. sqrt + division (library)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) x = sqrt(x) + 1.5;
    t1 = clock64();
    if (lane == 0) cyc[8] = (t1 - t0) / 16;
    // 9. raw MUFU rcp approx chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
    t1 = clock64();
    if (lane == 0) cyc[9] = (t1 - t0) / 16;
    // 10. float conversions
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) x = (double)((float)x) ;
    t1 = clock64();
    if (lane == 0) cyc[10] = (t1 - t0) / 16;
    out[lane] = x;
}

// POTRF variants on a 32x32 SPD tile: cold (first call) and warm (second call in the same kernel)
__global__ void potrf_kernel(const double* A, double* L, long long* cyc, int variant, int wtest)
{
    __shared__ __align__(16) double C[kCT][kCLD];
    __shared__ double s_rinv[kCT];
    __shared__ __align__(16) double s_col[8 * kCT];
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    for (int rep = 0; rep < 3; ++rep) {
        for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) C[idx >> 5][idx & 31] = A[idx];
        if (tid == 0) s_bad = 0;
        __syncthreads();
        const long long t0 = clock64();
        if (variant == 0) tile_potrf(&C[0][0], wtest, s_rinv, &s_bad);
        else if (variant == 1) { if (tid < 32) tile_potrf_warp(&C[0][0], wtest, &s_bad, s_col); __syncthreads(); tile_potrf_scale(&C[0][0], wtest, s_rinv); }
        else tile_potrf_blocked(&C[0][0], wtest, s_rinv, &s_bad, s_col);
        __syncthreads();
        const long long t1 = clock64();
        if (tid == 0) cyc[rep] = t1 - t0;
    }
    for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) L[idx] = ((idx & 31) <= (idx >> 5)) ? C[idx >> 5][idx & 31] : 0.0;
}

// TRSM variants: X L^T = P for a 32 x 32 tile P
__global__ void trsm_kernel(const double* Lg, const double* Pg, double* X, long long* cyc, int variant)
{
    __shared__ __align__(16) double C[kCT][kCLD], B1[kCT + 8][kCLD];
    __shared__ double s_rinv[kCT], s_winv[kCT * 8];
    const int tid = threadIdx.x;
    for (int rep = 0; rep < 3; ++rep) {
        for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) { C[idx >> 5][idx & 31] = Pg[idx]; B1[idx >> 5][idx & 31] = Lg[idx]; }
        for (int idx = tid; idx < 8 * kCLD; idx += blockDim.x) (&B1[kCT][0])[idx] = 0.0;
        if (tid < 32) s_rinv[tid] = 1.0 / Lg[tid * 32 + tid];
        __syncthreads();
        const long long t0 = clock64();
        if (variant == 0) { if (tid < kCT) tile_trsm_row(&C[tid][0], &B1[0][0], s_rinv, kCT); }
        else tile_trsm_dmma(&C[0][0], kCT, &B1[0][0], s_rinv, kCT, s_winv);
        __syncthreads();
        const long long t1 = clock64();
        if (tid == 0) cyc[rep] = t1 - t0;
    }
    for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) X[idx] = C[idx >> 5][idx & 31];
}

// the in-DAG tile update C -= A1 * B1^T from shared memory
__global__ void gemm_kernel(const double* Ag, double* X, long long* cyc)
{
    __shared__ __align__(16) double C[kCT][kCLD], A1[kCT + 8][kCLD], B1[kCT + 8][kCLD];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) { C[idx >> 5][idx & 31] = 0; A1[idx >> 5][idx & 31] = Ag[idx]; B1[idx >> 5][idx & 31] = Ag[idx] + 1; }
    for (int rep = 0; rep < 3; ++rep) {
        __syncthreads();
        const long long t0 = clock64();
        tile_gemm_sub(&C[0][0], &A1[0][0], &B1[0][0]);
        __syncthreads();
        const long long t1 = clock64();
        if (tid == 0) cyc[rep] = t1 - t0;
    }
    for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) X[idx] = C[idx >> 5][idx & 31];
}

int main()
{
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 64 * 8); cudaMalloc(&d_cyc, 16 * 8);
    lat_kernel<<<1, 32>>>(d_out, d_cyc, 1.25);
    long long c[16];
    cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const char* nm[] = {"dfma", "dmul", "fast_rcp+add", "fast_rsqrt+add", "1/x+add", "rsqrt()+add", "shfl64", "sts-sync-lds-sync", "sqrt()+add", "mufu.rcp64h", "f2f round trip"};
    for (int i = 0; i < 11; ++i) printf("%-20s %lld cycles\n", nm[i], c[i]);
    // SPD tile
    static double hA[32 * 32], hL[32 * 32], hP[32 * 32], hX[32 * 32];
    for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) hA[i * 32 + j] = (i == j ? 40.0 : 0.0) + 1.0 / (1 + i + j);
    for (int i = 0; i < 1024; ++i) hP[i] = sin(0.37 * i) + 0.1;
    double *dA, *dL, *dP, *dX; cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dL, sizeof(hL)); cudaMalloc(&dP, sizeof(hP)); cudaMalloc(&dX, sizeof(hX));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    cudaMemcpy(dP, hP, sizeof(hP), cudaMemcpyHostToDevice);
    for (int wt : {32, 26, 5})
    for (int v = 0; v < 3; ++v) {
        potrf_kernel<<<1, 256>>>(dA, dL, d_cyc, v, wt);
        cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
        cudaMemcpy(hL, dL, sizeof(hL), cudaMemcpyDeviceToHost);
        double err = 0, rowkeep = 0;
        for (int i = 0; i < wt; ++i) for (int j = 0; j <= i; ++j) {
            double s = 0; for (int k = 0; k <= j; ++k) s += hL[i * 32 + k] * hL[j * 32 + k];
            err = fmax(err, fabs(s - hA[i * 32 + j]));
        }
        if (wt < 32) for (int j = 0; j < wt; ++j) rowkeep = fmax(rowkeep, fabs(hL[wt * 32 + j] - hA[wt * 32 + j]));   // the row below must survive
        printf("potrf variant %d w %d: cold %lld, warm %lld %lld cycles, |LL^T-A| %.2e, row w changed by %.1e\n", v, wt, c[0], c[1], c[2], err, rowkeep);
    }
    for (int v = 0; v < 2; ++v) {
        trsm_kernel<<<1, 256>>>(dL, dP, dX, d_cyc, v);
        cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
        cudaMemcpy(hX, dX, sizeof(hX), cudaMemcpyDeviceToHost);
        double err = 0;
        for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) {
            double s = 0; for (int k = 0; k <= j; ++k) s += hX[i * 32 + k] * hL[j * 32 + k];
            err = fmax(err, fabs(s - hP[i * 32 + j]));
        }
        printf("trsm variant %d: cold %lld, warm %lld %lld cycles, |XL^T-P| %.2e\n", v, c[0], c[1], c[2], err);
    }
    gemm_kernel<<<1, 256>>>(dA, dX, d_cyc);
    cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("tile gemm: cold %lld, warm %lld %lld cycles\n", c[0], c[1], c[2]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
