// phase timing of the banded LDL^T warp solver (development aid)
#define MCCBA_BAND_DBG 1
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../multi_camera_calibration_b200/csrc/mccba_dense.cuh"
using namespace mccba;
template <int NW>
__global__ void k(const double* A, int n, double* x, long long* out)
{
    extern __shared__ __align__(16) unsigned char sm[];
    double* colbuf = reinterpret_cast<double*>(sm);
    double* rhs = colbuf + 256;
    double* band = rhs + ((n + 1) & ~1) + 2;
    double* pinv = band + (size_t)(n + 1) * NW;
    constexpr int w = NW - 1;
    for (int idx = threadIdx.x; idx < n * NW; idx += blockDim.x) {
        const int r = idx / NW, kk = idx - r * NW, c = r - w + kk;
        band[idx] = c >= 0 ? A[(long long)r * n + c] : 0.0;
    }
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) rhs[idx] = A[(long long)n * n + idx];
    __syncthreads();
    if (threadIdx.x < 32) band_ldlt_solve_warp<NW>(band, rhs, n, colbuf, pinv);
    __syncthreads();
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) x[idx] = rhs[idx];
    if (threadIdx.x == 0) { out[0] = g_band_ts[1] - g_band_ts[0]; out[1] = g_band_ts[2] - g_band_ts[1]; }
}
int main()
{
    const int n = 378, NW = 12, w = 11;
    std::vector<double> A((size_t)(n + 1) * n, 0.0);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (abs(i - j) <= w) A[(size_t)i * n + j] = (i == j ? 40.0 : 0.0) + 1.0 / (1 + abs(i - j));
    for (int i = 0; i < n; ++i) A[(size_t)n * n + i] = sin(0.3 * i);
    double *dA, *dx; long long* dc; long long c[2];
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dx, n * 8); cudaMalloc(&dc, 16);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    const size_t smem = 8 * ((size_t)(n + 1) * NW + (n + 2) + 3 * ((n + 1) / 2) + 2 + 256 + 8);
    cudaFuncSetAttribute(k<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
        k<NW><<<1, 256, smem>>>(dA, n, dx, dc);
        cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
        printf("forward %lld cycles (%.0f / pivot), backward %lld cycles (%.0f / step)\n", c[0], (double)c[0] / n, c[1], (double)c[1] / n);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
