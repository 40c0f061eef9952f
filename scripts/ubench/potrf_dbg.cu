// phase timing of the blocked POTRF (development aid)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../multi_camera_calibration_b200/csrc/mccba_dense.cuh"
using namespace mccba;
__global__ void k(const double* A, long long* cyc, int w)
{
    __shared__ __align__(16) double Cs[kCT][kCLD];
    __shared__ __align__(16) double T[256];
    __shared__ double rinv[32];
    double* C = &Cs[0][0];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    for (int idx = tid; idx < 1024; idx += blockDim.x) Cs[idx >> 5][idx & 31] = A[idx];
    __syncthreads();
    long long ts[24]; int nts = 0;
    ts[nts++] = clock64();
    for (int p0 = 0; p0 < w; p0 += 8) {
        if (warp == 0) {
            double a[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = lane < w ? C[lane * kCLD + p0 + k] : (p0 + k == lane ? 1.0 : 0.0);
            if (p0 == 0) ts[nts++] = clock64();
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = p0 + jj;
                if (j < w) {
                    const double d = __shfl_sync(0xffffffffu, a[jj], j);
                    double ck[8];
#pragma unroll
                    for (int k = jj + 1; k < 8; ++k) ck[k] = __shfl_sync(0xffffffffu, a[jj], p0 + k);
                    const double tt = a[jj] * pivot_rcp(d);
                    T[lane * 8 + jj] = lane > j ? tt : 0.0;
#pragma unroll
                    for (int k = jj + 1; k < 8; ++k) a[k] = fma(-tt, ck[k], a[k]);
                } else T[lane * 8 + jj] = 0.0;
                if (p0 == 0 && jj < 3) ts[nts++] = clock64();
            }
            if (lane < w) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (p0 + k < w && lane >= p0 + k) C[lane * kCLD + p0 + k] = a[k];
            }
        }
        if (warp == 0) ts[nts++] = clock64();
        __syncthreads();
        if (warp == 0) ts[nts++] = clock64();
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
        if (warp == 0) ts[nts++] = clock64();
        __syncthreads();
        if (warp == 0) ts[nts++] = clock64();
    }
    if (tid < kCT) rinv[tid] = tid < w ? rsqrt(C[tid * kCLD + tid]) : 0.0;
    __syncthreads();
    if (warp == 0) ts[nts++] = clock64();
    {
        const int r = tid >> 5, c = tid & 31;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (r + 8 * q < w && c <= r + 8 * q) C[(r + 8 * q) * kCLD + c] *= rinv[c];
    }
    __syncthreads();
    if (warp == 0) ts[nts++] = clock64();
    if (tid == 0) { for (int i = 0; i < nts; ++i) cyc[i] = ts[i] - ts[0]; cyc[nts] = -1; }
}
int main()
{
    static double hA[1024];
    for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) hA[i * 32 + j] = (i == j ? 40.0 : 0.0) + 1.0 / (1 + i + j);
    double* dA; long long* dc; long long c[32];
    cudaMalloc(&dA, 8192); cudaMalloc(&dc, 256);
    cudaMemcpy(dA, hA, 8192, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 256>>>(dA, dc, 32);
        cudaMemcpy(c, dc, 256, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 32 && c[i] >= 0; ++i) printf(" %lld", c[i]);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
