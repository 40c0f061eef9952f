// timeline of the block cyclic reduction (mccba_bcr.cuh) on a synthetic block-tridiagonal SPD system (development aid)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define MCCBA_BCR_TS 1
__device__ long long g_bcr_ts[64];
#include "../../multi_camera_calibration_b200/csrc/mccba_bcr.cuh"
using namespace mccba;

template <int B>
__global__ void __launch_bounds__(BcrCfg<B>::kThreads) k(const double* A, int n, double* xout)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int Nb = bcr_blocks(n, B);
    double* Dg = reinterpret_cast<double*>(smem);
    double* Lo = Dg + (size_t)Nb * B * B;
    double* Tmp = Lo + (size_t)Nb * B * B;
    double* rhs = Tmp + (size_t)((Nb + 1) / 2) * B * B;
    if (threadIdx.x == 0) g_bcr_ts[0] = clock64();
    bcr_stage<B>(A, n, 0, Dg, Lo, rhs, Nb);
    __syncthreads();
    int fail = 0;
    if (threadIdx.x < 32) {   // one elimination alone, with per-pivot stamps (slot 777 switches them on) ...
        if (threadIdx.x == 0) g_bcr_ts[39] = clock64();
        fail += bcr_eliminate<B>(Dg, Lo, Tmp, rhs, 1, 0, 2, 777, threadIdx.x, rhs + (size_t)Nb * B);
        if (threadIdx.x == 0) g_bcr_ts[50] = clock64();
    }
    __syncthreads();
    bcr_stage<B>(A, n, 0, Dg, Lo, rhs, Nb);   // ... then the system is staged again for the real solve
    __syncthreads();
    if (threadIdx.x == 0) g_bcr_ts[1] = clock64();
    fail += bcr_solve_cta<B>(Dg, Lo, Tmp, rhs, Nb);
    if (threadIdx.x == 0) g_bcr_ts[2] = clock64();
    for (int i = threadIdx.x; i < n; i += blockDim.x) xout[i] = rhs[i] + fail;
}

int main()
{
    const int nc = 63, n = 6 * nc, B = 6;
    std::vector<double> A((size_t)(n + 1) * n, 0.0);
    srand(1);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j)
            if (i / 6 - j / 6 <= 1) { double v = (rand() % 1000) / 1000.0 - 0.5; A[(size_t)i * n + j] = v; A[(size_t)j * n + i] = v; }
    for (int i = 0; i < n; ++i) A[(size_t)i * n + i] = 20.0;
    for (int i = 0; i < n; ++i) A[(size_t)n * n + i] = (rand() % 1000) / 1000.0;
    double *dA, *dx;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dx, n * 8);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    const size_t smem = bcr_smem_bytes(n, B);
    cudaFuncSetAttribute(k<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<B><<<1, BcrCfg<B>::kThreads, smem>>>(dA, n, dx);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long ts[64];
        cudaMemcpyFromSymbol(ts, g_bcr_ts, sizeof(ts));
        printf("kernel %.2f us; stage %lld cyc, solve %lld cyc | levels:", ms * 1e3, ts[1] - ts[0], ts[2] - ts[1]);
        for (int i = 3; i < 40 && ts[i]; ++i) printf(" %lld", ts[i] - ts[i - 1 < 3 ? 1 : i - 1]);
        printf("\n   single elimination: load %lld | pivots", ts[40] - ts[39]);
        for (int i = 41; i < 47; ++i) printf(" %lld", ts[i] - ts[i - 1]);
        printf(" | store %lld\n", ts[50] - ts[46]);
    }
    std::vector<double> x(n);
    cudaMemcpy(x.data(), dx, n * 8, cudaMemcpyDeviceToHost);
    double res = 0;
    for (int i = 0; i < n; ++i) { double s = -A[(size_t)n * n + i]; for (int j = 0; j < n; ++j) s += A[(size_t)i * n + j] * x[j]; res = fmax(res, fabs(s)); }
    printf("residual %.3e (%s)\n", res, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
