// DMMA (mma.sync m8n8k4 f64) latency and a 32x32x32 tile update built on it (development aid)
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
constexpr int LD = 34;
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// C -= A * B^T, all 32 x 32, stride LD, 256 threads (8 warps x two 8x8 output tiles)
__device__ __forceinline__ void tile_gemm_sub(double* C, const double* A, const double* B)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int tr = warp >> 1, tc0 = (warp & 1) * 2;
    const double* ap = A + (8 * tr + g) * LD + t;
    const double* bp0 = B + (8 * tc0 + g) * LD + t;
    const double* bp1 = bp0 + 8 * LD;
    double c00 = 0, c01 = 0, c10 = 0, c11 = 0, d00 = 0, d01 = 0, d10 = 0, d11 = 0;
#pragma unroll
    for (int ks = 0; ks < 8; ks += 2) {
        const double a0 = ap[4 * ks], a1 = ap[4 * ks + 4];
        dmma(c00, c01, a0, bp0[4 * ks]);
        dmma(c10, c11, a0, bp1[4 * ks]);
        dmma(d00, d01, a1, bp0[4 * ks + 4]);
        dmma(d10, d11, a1, bp1[4 * ks + 4]);
    }
    double* cp = C + (8 * tr + g) * LD + 8 * tc0 + 2 * t;
    cp[0] -= c00 + d00; cp[1] -= c01 + d01; cp[8] -= c10 + d10; cp[9] -= c11 + d11;
}
__global__ void k(const double* Ag, double* X, long long* cyc)
{
    __shared__ __align__(16) double C[32 * LD], A1[32 * LD], B1[32 * LD];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < 1024; idx += blockDim.x) { const int r = idx >> 5, c = idx & 31; C[r * LD + c] = 0; A1[r * LD + c] = Ag[idx]; B1[r * LD + c] = Ag[idx] + 1; }
    __syncthreads();
    // DMMA dependent-chain latency
    {
        double c0 = 0, c1 = 0, a = A1[tid & 31], b = B1[tid & 31];
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < 32; ++i) dmma(c0, c1, a, b);
        const long long t1 = clock64();
        if (tid == 0) cyc[4] = (t1 - t0) / 32;
        if (c0 + c1 == 12345.678) X[0] = c0;
    }
    for (int rep = 0; rep < 3; ++rep) {
        __syncthreads();
        const long long t0 = clock64();
        tile_gemm_sub(C, A1, B1);
        __syncthreads();
        const long long t1 = clock64();
        if (tid == 0) cyc[rep] = t1 - t0;
    }
    for (int idx = tid; idx < 1024; idx += blockDim.x) X[idx] = C[(idx >> 5) * LD + (idx & 31)];
}
int main()
{
    static double hA[1024], hX[1024];
    for (int i = 0; i < 1024; ++i) hA[i] = sin(0.37 * i) + 0.1;
    double *dA, *dX; long long* dc; long long c[8];
    cudaMalloc(&dA, 8192); cudaMalloc(&dX, 8192); cudaMalloc(&dc, 64);
    cudaMemcpy(dA, hA, 8192, cudaMemcpyHostToDevice);
    k<<<1, 256>>>(dA, dX, dc);
    cudaMemcpy(c, dc, 64, cudaMemcpyDeviceToHost); cudaMemcpy(hX, dX, 8192, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) {
        double s = 0; for (int m = 0; m < 32; ++m) s += hA[i * 32 + m] * (hA[j * 32 + m] + 1);
        err = fmax(err, fabs(hX[i * 32 + j] + 3 * s));
    }
    printf("dmma chain %lld cycles/op; tile gemm %lld %lld %lld cycles; max err %.2e\n", c[4], c[0], c[1], c[2], err);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
