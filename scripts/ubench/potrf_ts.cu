// per-pivot time stamps of the warp-level POTRF (development aid)
#define MCCBA_POTRF_DBG 1
#include <cstdio>
#include <cuda_runtime.h>
#include "../../multi_camera_calibration_b200/csrc/mccba_dense.cuh"
using namespace mccba;
__global__ void k(const double* A, long long* out, int w)
{
    __shared__ __align__(16) double C[kCT][kCLD];
    __shared__ __align__(16) double s_col[4 * kCT];
    __shared__ double s_rinv[kCT];
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < 1024; idx += blockDim.x) C[idx >> 5][idx & 31] = A[idx];
    __syncthreads();
    const long long t0 = clock64();
    if (tid < 32) tile_potrf_warp(&C[0][0], w, &s_bad, s_col);
    __syncthreads();
    tile_potrf_scale(&C[0][0], w, s_rinv);
    if (tid == 0) g_potrf_ts[33] = clock64();
    if (tid == 0) { for (int i = 0; i < 34; ++i) out[i] = g_potrf_ts[i] - t0; }
}
int main()
{
    static double hA[1024];
    for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) hA[i * 32 + j] = (i == j ? 40.0 : 0.0) + 1.0 / (1 + i + j);
    double* dA; long long* dc; long long c[40];
    cudaMalloc(&dA, 8192); cudaMalloc(&dc, 320);
    cudaMemcpy(dA, hA, 8192, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 256>>>(dA, dc, 32);
        cudaMemcpy(c, dc, 320, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 34; ++i) printf(" %lld", c[i]);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
