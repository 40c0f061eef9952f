// dependent-chain attribution for the Cholesky pivot (development aid)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double seed(double a) { double d; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(d) : "d"(a)); return d; }
__device__ __forceinline__ double vf(double a, double b, double c) { double d; asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c)); return d; }
template <int kVar>
__global__ void k(double* out, long long* cyc, double s0)
{
    __shared__ __align__(16) double buf[128];
    const int lane = threadIdx.x;
    buf[lane] = 1.0; buf[lane + 32] = 1.0; buf[lane + 64] = 1.0; buf[lane + 96] = 1.0;
    __syncwarp();
    double x = s0 + lane * 1e-3;
    const long long t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
        double d = x;
        if (kVar & 1) {   // shared-memory broadcast round trip
            double* cb = buf + (j & 1) * 64;
            cb[lane] = x;
            __syncwarp();
            double v0, v1;
            const unsigned addr = (unsigned)__cvta_generic_to_shared(cb + (j & ~1));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(addr) : "memory");
            d = (j & 1) ? v1 : v0;
        }
        double y = d;
        if (kVar & 2) {   // reciprocal: seed + two Newton steps
            y = seed(d);
            double r = vf(-d, y, 1.0);
            y = vf(y, r, y);
            r = vf(-d, y, 1.0);
            y = vf(y, r, y);
        }
        if (kVar & 4) {   // multiply + update fma
            const double t = x * y;
            x = vf(-t, 0.25, x + 1.0);
        } else x = y + 1.5;
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[0] = (t1 - t0) / 32;
    out[lane] = x;
}
int main()
{
    double* d_out; long long* d_cyc; long long c;
    cudaMalloc(&d_out, 256); cudaMalloc(&d_cyc, 8);
#define RUN(V, NAME) k<V><<<1, 32>>>(d_out, d_cyc, 1.25); cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost); printf("%-40s %lld cycles/iter\n", NAME, c);
    RUN(0, "add only");
    RUN(1, "smem round trip + add");
    RUN(2, "rcp (seed + 2 Newton) + add");
    RUN(4, "mul + fma + add");
    RUN(3, "smem + rcp + add");
    RUN(6, "rcp + mul + fma");
    RUN(7, "smem + rcp + mul + fma (full chain)");
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
