// single-warp issue-rate microbenchmarks (development aid)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void thr_kernel(double* out, long long* cyc, double seed, int off)
{
    __shared__ __align__(16) double buf[128];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 128; i += blockDim.x) buf[i] = seed + i;
    __syncthreads();
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = seed + lane * 1e-3 + i;
    long long t0, t1;
    t0 = clock64();
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], 1.0000001, 1e-9);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);   // 128 independent-ish DFMA
    // LDS.64 broadcast, dynamic address
    double s = 0;
    const double* p = buf + off;
    t0 = clock64();
    double v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = p[i];
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0);   // 32 LDS + 32 dependent DADD
    // LDS + DFMA interleaved as in the pivot update
    t0 = clock64();
    double a[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = x[i & 7] + i;
    t1 = clock64();
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 31; ++i) a[i] = fma(-s, p[i + 1], a[i + 1]);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0);   // 31 LDS + 31 DFMA
    double acc = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += a[i];
    // LDS.128
    const double2* p2 = reinterpret_cast<const double2*>(buf) + off;
    t0 = clock64();
    double2 w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = p2[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(-s, w[i].x, a[i]);
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i + 16] = fma(-s, w[i].y, a[i + 16]);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0);   // 16 LDS.128 + 32 DFMA
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += a[i];
    out[threadIdx.x] = acc + s + x[0] + x[1] + x[2] + x[3] + x[4] + x[5] + x[6] + x[7];
}
int main()
{
    double* d_out; long long* d_cyc; long long c[8];
    cudaMalloc(&d_out, 1024 * 8); cudaMalloc(&d_cyc, 64);
    for (int nt : {32, 128, 256}) {
        thr_kernel<<<1, nt>>>(d_out, d_cyc, 1.25, 3);
        cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
        printf("threads %d: 128 indep DFMA %lld | 32 LDS64 + 32 dep DADD %lld | 31 LDS64+31 DFMA %lld | 16 LDS128 + 32 DFMA %lld\n", nt, c[0], c[1], c[2], c[3]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
