// FFMA2 / FFMA / DFMA throughput as a function of operand reuse (development aid; profiles/r2_pipe_peaks.txt).
//   chain  d = fma(d, A, B)          A, B the same two registers for every chain (operand-reuse cache hits)
//   acc    acc[i][j] += x[i] * y[j]  the accumulate pattern of the residual kernel: 64 accumulators, 8 + 8 operands
//   3reg   d[i] = fma(a[i], b[i], d[i]) with 3 distinct registers per instruction and no reuse between neighbours
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mk(float x, float y) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }

template <int kOp>
__global__ void __launch_bounds__(256) k(double* out, int outer, float seed)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (kOp == 0 || kOp == 1) {          // packed: acc pattern (0), 3reg pattern (1)
        u64 acc[32], x[8], y[8];
        for (int i = 0; i < 32; ++i) acc[i] = mk(seed + i, seed - i);
        for (int i = 0; i < 8; ++i) { x[i] = mk(1.0f + 1e-7f * (i + tid), 1.0f - 1e-7f * i); y[i] = mk(1e-9f * (i + 1), 2e-9f * i + seed * 1e-12f); }
        for (int o = 0; o < outer; ++o) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (kOp == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i * 4 + j] = fma2(x[i], y[(j + r) & 7], acc[i * 4 + j]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = fma2(x[(i * 3 + r) & 7], y[(i * 5 + 2 * r + 1) & 7], acc[i]);
                }
            }
        }
        double s = 0;
        for (int i = 0; i < 32; ++i) s += (double)(acc[i] & 0xffff);
        out[tid] = s;
    } else if (kOp == 2 || kOp == 3) {   // scalar FFMA: acc pattern (2), 3reg (3)
        float acc[32], x[8], y[8];
        for (int i = 0; i < 32; ++i) acc[i] = seed + i;
        for (int i = 0; i < 8; ++i) { x[i] = 1.0f + 1e-7f * (i + tid); y[i] = 1e-9f * (i + 1) + seed * 1e-12f; }
        for (int o = 0; o < outer; ++o) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (kOp == 2) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i * 4 + j] = fmaf(x[i], y[(j + r) & 7], acc[i * 4 + j]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = fmaf(x[(i * 3 + r) & 7], y[(i * 5 + 2 * r + 1) & 7], acc[i]);
                }
            }
        }
        double s = 0;
        for (int i = 0; i < 32; ++i) s += acc[i];
        out[tid] = s;
    } else {                             // DFMA: acc pattern (4), 3reg (5)
        double acc[32], x[8], y[8];
        for (int i = 0; i < 32; ++i) acc[i] = seed + i;
        for (int i = 0; i < 8; ++i) { x[i] = 1.0 + 1e-7 * (i + tid); y[i] = 1e-9 * (i + 1) + seed * 1e-12; }
        for (int o = 0; o < outer; ++o) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (kOp == 4) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i * 4 + j] = fma(x[i], y[(j + r) & 7], acc[i * 4 + j]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = fma(x[(i * 3 + r) & 7], y[(i * 5 + 2 * r + 1) & 7], acc[i]);
                }
            }
        }
        double s = 0;
        for (int i = 0; i < 32; ++i) s += acc[i];
        out[tid] = s;
    }
}
template <int kOp>
static void run(const char* name, int sms, double* out, int warps_per_sm)
{
    const int grid = sms * warps_per_sm / 8, outer = 400;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<kOp><<<grid, 256>>>(out, 10, 1.0f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0); k<kOp><<<grid, 256>>>(out, outer, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double li = (double)grid * 256 * outer * 8 * 32;
    printf("%-12s %2d warps/SM  %7.3f ms  %6.1f lane-instr/clk/SM @1.965GHz\n", name, warps_per_sm, best, li / (best * 1e-3) / sms / 1.965e9);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 64 * 32);
    for (int w : {8, 16, 32}) {
        run<0>("FFMA2 acc", p.multiProcessorCount, out, w); run<1>("FFMA2 3reg", p.multiProcessorCount, out, w);
        run<2>("FFMA acc", p.multiProcessorCount, out, w); run<3>("FFMA 3reg", p.multiProcessorCount, out, w);
        run<4>("DFMA acc", p.multiProcessorCount, out, w); run<5>("DFMA 3reg", p.multiProcessorCount, out, w);
    }
    return 0;
}
