// phase timing of the blocked TRSM (development aid)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../multi_camera_calibration_b200/csrc/mccba_dense.cuh"
using namespace mccba;
template <int kVar> __global__ void k(const double* Lg, const double* Pg, double* X, long long* cyc)
{
    __shared__ double C[kCT][kCLD], B1[kCT + 8][kCLD];
    __shared__ double s_rinv[kCT];
    __shared__ __align__(16) double winv[kCT * 8], sx[32 * 8];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) { C[idx >> 5][idx & 31] = Pg[idx]; B1[idx >> 5][idx & 31] = Lg[idx]; }
    for (int idx = tid; idx < 8 * kCLD; idx += blockDim.x) (&B1[kCT][0])[idx] = 0.0;
    if (tid < 32) s_rinv[tid] = 1.0 / Lg[tid * 32 + tid];
    __syncthreads();
    double* P = &C[0][0]; const double* L = &B1[0][0]; const double* rinv = s_rinv; const int w = 32, h = 32;
    long long ts[8]; int nts = 0;
    ts[nts++] = clock64();
    const int q = tid & 7, r = tid >> 3, lane = tid & 31;
    if (tid < 32) {
        const int b0 = (tid >> 3) * 8;
        double wv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double acc = (i == q) ? 1.0 : 0.0;
#pragma unroll
            for (int m = 0; m < i; ++m) acc = fma(-L[(b0 + i) * kCLD + b0 + m], wv[m], acc);
            wv[i] = (i >= q && b0 + i < w) ? acc * rinv[b0 + i] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) winv[(b0 + i) * 8 + q] = wv[i];
    }
    ts[nts++] = clock64();
    __syncthreads();
    ts[nts++] = clock64();
    const bool live = r < h;
    double* prow = P + (live ? r : 0) * kCLD;
    for (int cb = 0; cb < w; cb += 8) {
        const double* lrow = L + (cb + q) * kCLD;
        double s0 = prow[cb + q], s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int m = 0; m < cb; m += 8) {
            double x8[8], l8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { x8[u] = prow[m + u]; l8[u] = lrow[m + u]; }
            s0 = fma(-x8[0], l8[0], s0); s1 = fma(-x8[1], l8[1], s1); s2 = fma(-x8[2], l8[2], s2); s3 = fma(-x8[3], l8[3], s3);
            s0 = fma(-x8[4], l8[4], s0); s1 = fma(-x8[5], l8[5], s1); s2 = fma(-x8[6], l8[6], s2); s3 = fma(-x8[7], l8[7], s3);
        }
        const double s = (s0 + s1) + (s2 + s3);
        const double* wrow = winv + (cb + q) * 8;
        double x0 = 0.0, x1 = 0.0;
        if (kVar == 0) {
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                x0 = fma(__shfl_sync(0xffffffffu, s, (lane & ~7) + u), wrow[u], x0);
                x1 = fma(__shfl_sync(0xffffffffu, s, (lane & ~7) + u + 1), wrow[u + 1], x1);
            }
        } else {
            sx[r * 8 + q] = s;
            __syncwarp();
            const double* sr = sx + r * 8;
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                x0 = fma(sr[u], wrow[u], x0);
                x1 = fma(sr[u + 1], wrow[u + 1], x1);
            }
        }
        if (live && cb + q < w) prow[cb + q] = x0 + x1;
        __syncwarp();
        ts[nts++] = clock64();
    }
    __syncthreads();
    ts[nts++] = clock64();
    if (tid == 0 || tid == blockDim.x - 1) for (int i = 0; i < nts; ++i) cyc[(tid ? 8 : 0) + i] = ts[i] - ts[0];
    for (int idx = tid; idx < kCT * kCT; idx += blockDim.x) X[idx] = C[idx >> 5][idx & 31];
}
int main()
{
    static double hL[1024], hP[1024];
    for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) hL[i * 32 + j] = j < i ? 0.1 / (1 + i + j) : (i == j ? 3.0 : 0.0);
    for (int i = 0; i < 1024; ++i) hP[i] = sin(0.37 * i) + 0.1;
    double *dL, *dP, *dX; long long* dc; long long c[16];
    cudaMalloc(&dL, 8192); cudaMalloc(&dP, 8192); cudaMalloc(&dX, 8192); cudaMalloc(&dc, 128);
    cudaMemcpy(dL, hL, 8192, cudaMemcpyHostToDevice); cudaMemcpy(dP, hP, 8192, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 4; ++rep) {
        if (rep == 0) k<0><<<1, 256>>>(dL, dP, dX, dc);
        if (rep == 1) k<1><<<1, 256>>>(dL, dP, dX, dc);
        if (rep == 2) k<0><<<1, 32>>>(dL, dP, dX, dc);
        if (rep == 3) k<1><<<1, 32>>>(dL, dP, dX, dc);
        cudaMemcpy(c, dc, 128, cudaMemcpyDeviceToHost);
        for (int t = 0; t < 2; ++t) { printf("thread %d:", t ? 255 : 0); for (int i = 0; i < 8; ++i) printf(" %lld", c[8 * t + i]); printf("\n"); }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
