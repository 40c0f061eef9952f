// Full-chip arithmetic pipe peaks on the box (development aid; numbers recorded in profiles/r2_pipe_peaks.txt and used
// as the FP64 / FP32 roofline denominators next to MEASURED_PEAKS.json's HBM figure).
//   DFMA   fma.rn.f64            FFMA   fma.rn.f32            FFMA2  fma.rn.f32x2 (two FMAs per lane-instruction)
//   F2F    cvt.f64.f32 + cvt.rn.f32.f64 round trip            MIX    FFMA2 and DFMA interleaved 4:1
// Every thread runs kChains independent dependency chains so the pipe, not the latency, is the limit.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kChains = 8;
constexpr int kInner = 64;

template <int kOp>
__global__ void __launch_bounds__(256) peak_kernel(double* out, int outer, double seed)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double d[kChains];
    float f[kChains];
    unsigned long long p[kChains];
    for (int i = 0; i < kChains; ++i) {
        d[i] = seed + 1e-3 * i + 1e-6 * tid;
        f[i] = (float)d[i];
        const float2 v = make_float2(f[i], f[i] + 0.5f);
        p[i] = *reinterpret_cast<const unsigned long long*>(&v);
    }
    const double da = 1.0000001, db = 1e-9;
    const float fa = 1.0000001f, fb = 1e-9f;
    const float2 pa2 = make_float2(fa, fa), pb2 = make_float2(fb, fb);
    const unsigned long long pa = *reinterpret_cast<const unsigned long long*>(&pa2), pb = *reinterpret_cast<const unsigned long long*>(&pb2);
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int r = 0; r < kInner; ++r) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
                if (kOp == 0) d[i] = fma(d[i], da, db);
                if (kOp == 1) f[i] = fmaf(f[i], fa, fb);
                if (kOp == 2) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(pa), "l"(pb));
                if (kOp == 3) { d[i] = (double)f[i] + db; f[i] = (float)d[i]; }
                if (kOp == 4) {
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(pa), "l"(pb));
                    if ((i & 3) == 0) d[i] = fma(d[i], da, db);
                }
            }
        }
    }
    double s = 0;
    for (int i = 0; i < kChains; ++i) {
        const float2 v = *reinterpret_cast<const float2*>(&p[i]);
        s += d[i] + f[i] + v.x + v.y;
    }
    out[tid] = s;
}

template <int kOp>
static double run(const char* name, double ops_per_iter, int sms, double* out)
{
    const int grid = sms * 8, outer = 200;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    peak_kernel<kOp><<<grid, 256>>>(out, 10, 1.0);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        peak_kernel<kOp><<<grid, 256>>>(out, outer, 1.0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double lane_instr = (double)grid * 256 * outer * kInner * kChains;
    const double rate = lane_instr / (best * 1e-3);
    printf("%-6s %8.3f ms  %8.2f G lane-instr/s  %7.1f lane-instr/clk/SM @1.965GHz  %8.2f T%s/s\n", name, best, rate * 1e-9,
           rate / sms / 1.965e9, rate * ops_per_iter * 1e-12, kOp == 3 ? "cvt-pairs" : "FLOP");
    return rate;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("device %s, %d SMs, clock %d MHz (max)\n", prop.name, prop.multiProcessorCount, prop.clockRate / 1000);
    double* out;
    cudaMalloc(&out, sizeof(double) * prop.multiProcessorCount * 8 * 256);
    run<0>("DFMA", 2, prop.multiProcessorCount, out);
    run<1>("FFMA", 2, prop.multiProcessorCount, out);
    run<2>("FFMA2", 4, prop.multiProcessorCount, out);
    run<3>("F2F", 1, prop.multiProcessorCount, out);
    run<4>("MIX", 4.5, prop.multiProcessorCount, out);
    cudaFree(out);
    return 0;
}
