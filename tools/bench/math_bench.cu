// math_bench.cu -- device microbenchmark of the glibc-exact sin / cos / tan / atan (csrc/ali_glibcmath.cuh)
// against CUDA's libm: latency of a dependent chain in one warp (uniform and divergent arguments) and
// throughput with a full device.  Dev tool:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o math_bench math_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../ali_fmm_and_ray_tracing_b200/csrc/ali_glxmath.cuh"

template <int F> __device__ __forceinline__ double call(double x)
{
    if (F == 0) return ali_glibc_sin(x);
    if (F == 1) return ali_glibc_cos(x);
    if (F == 2) return ali_glibc_tan(x);
    if (F == 3) return ali_glibc_atan(x);
    if (F == 4) return ali_glibc_sin(x) + ali_glibc_cos(x);
    if (F == 5) { double s, c; ali_gx_sincos(x, ali_gl_sincostab, s, c); return s + c; }
    if (F == 6) return ali_gx_atan(x, ali_gl_atan_cij);
    if (F == 10) return sin(x);
    if (F == 11) return cos(x);
    if (F == 12) return tan(x);
    if (F == 13) return atan(x);
    if (F == 14) { double s, c; sincos(x, &s, &c); return s + c; }
    if (F == 20) return 1.0 / x;
    if (F == 21) return sqrt(x);
    return x;
}

template <int F> __global__ void chain(double x0, double spread, int iters, long long *cyc, double *out)
{
    double x = x0 + spread * (threadIdx.x & 31);
    const double base = x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) x = base + 1e-9 * call<F>(x);   // dependent chain, argument stays near base
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <int F> void run(const char *name)
{
    long long *cyc; double *out;
    cudaMalloc(&cyc, 8); cudaMalloc(&out, 8 * 148 * 8 * 256);
    const int iters = 2000;
    const struct { const char *what; double x0, spread; } cases[] = {
        {"uniform x=0.5", 0.5, 0.0}, {"uniform x=1.7", 1.7, 0.0}, {"uniform x=2.9", 2.9, 0.0},
        {"divergent 0.05..3.1", 0.05, 0.0984}, {"divergent 0.01..30 (atan-like)", 0.01, 0.97}};
    for (auto &c : cases) {
        long long h = 0;
        chain<F><<<1, 32>>>(c.x0, c.spread, iters, cyc, out);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        chain<F><<<148 * 8, 256>>>(c.x0, c.spread, iters, cyc, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double calls = 148.0 * 8 * 256 * iters;
        printf("%-22s %-32s latency %7.0f cycles/call (1 warp) | throughput %7.2f G calls/s (full device)\n", name, c.what,
               (double)h / iters, calls / (ms * 1e-3) / 1e9);
    }
    cudaFree(cyc); cudaFree(out);
}

int main()
{
    run<0>("glibc sin"); run<10>("cuda sin");
    run<1>("glibc cos"); run<11>("cuda cos");
    run<4>("glibc sin+cos"); run<5>("branch-light sin+cos"); run<14>("cuda sincos");
    run<2>("glibc tan"); run<12>("cuda tan");
    run<3>("glibc atan"); run<6>("branch-light atan"); run<13>("cuda atan");
    run<20>("1/x"); run<21>("sqrt");
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
