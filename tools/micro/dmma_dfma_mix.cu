// dmma_dfma_mix.cu -- do mma.sync.m8n8k4.f64 (DMMA) and plain DFMA share one FP64 datapath on sm_100a?
// Register-only kernels: (1) 8 warps/SM of DMMA alone, (2) 8 warps/SM of DFMA alone, (3) both sets in the same CTA.
// If (3) takes max((1), (2)) the pipes are independent (a DMMA kernel could borrow the FP64 FMA pipe for extra
// flops); if it takes (1) + (2) they are one datapath.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_dfma_mix dmma_dfma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: DMMA warps only, 1: DFMA warps only, 2: warps 0..7 DMMA + warps 8..15 DFMA, 3: every warp both
__global__ void __launch_bounds__(512) k(double *out, int iters, double seed)
{
    const int warp = threadIdx.x >> 5;
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    double c[8][2], f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { c[u][0] = c[u][1] = 0.0; f[u] = u; }
    const bool do_mma = MODE == 0 || MODE == 3 || (MODE == 2 && warp < 8);
    const bool do_fma = MODE == 1 || MODE == 3 || (MODE == 2 && warp >= 8);
    if (MODE == 0 && warp >= 8) return;
    if (MODE == 1 && warp < 8) return;
    for (int it = 0; it < iters; ++it) {
        if (do_mma) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a), "d"(b));
        }
        if (do_fma) {
#pragma unroll
            for (int r = 0; r < 8; ++r)      // 64 DFMA per lane = 8 DMMA-equivalents of flops per warp (8*8*4*2 = 512 = 32 lanes * 16)
#pragma unroll
                for (int u = 0; u < 8; ++u) f[u] = fma(a, b, f[u]);
        }
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1] + f[u];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_it(F fn)
{
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    fn();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(s); fn(); cudaEventRecord(e); cudaEventSynchronize(e);
        float ms; cudaEventElapsedTime(&ms, s, e);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount, iters = 8192;
    double *out; cudaMalloc(&out, 64);
    printf("device %s, %d SMs; per iteration a DMMA warp issues 8 m8n8k4 (4096 flop), a DFMA warp 64 FMA per lane (4096 flop)\n", pr.name, sms);
    const double fl_mma = 8.0 * 512 * iters * 8 * sms, fl_fma = 64.0 * 2 * 32 * iters * 8 * sms;
    float t0 = time_it([&]() { k<0><<<sms, 512>>>(out, iters, 1.0); });
    float t1 = time_it([&]() { k<1><<<sms, 512>>>(out, iters, 1.0); });
    float t2 = time_it([&]() { k<2><<<sms, 512>>>(out, iters, 1.0); });
    float t3 = time_it([&]() { k<3><<<sms, 512>>>(out, iters, 1.0); });
    printf("DMMA alone  (8 warps/SM)            %8.3f ms  %6.2f TFLOP/s\n", t0, fl_mma / t0 / 1e9);
    printf("DFMA alone  (8 warps/SM)            %8.3f ms  %6.2f TFLOP/s\n", t1, fl_fma / t1 / 1e9);
    printf("8 DMMA warps + 8 DFMA warps         %8.3f ms  %6.2f TFLOP/s   (sum of the two alone: %.3f ms, max: %.3f ms)\n", t2,
           (fl_mma + fl_fma) / t2 / 1e9, t0 + t1, t0 > t1 ? t0 : t1);
    printf("16 warps, each DMMA + DFMA          %8.3f ms  %6.2f TFLOP/s\n", t3, 2 * (fl_mma + fl_fma) / t3 / 1e9);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
