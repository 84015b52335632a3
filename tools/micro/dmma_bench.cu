// dmma_bench.cu -- register-only throughput of the FP64 mma.sync shapes on sm_100a (no memory traffic):
// which DMMA shape / warps-per-SM / independent-accumulator count reaches the FP64 tensor peak.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int ILP>
__global__ void __launch_bounds__(1024) k(double *out, int iters, double seed)
{
    double a[8], b[4], c[ILP][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = seed * 0.5 + i;
#pragma unroll
    for (int u = 0; u < ILP; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[u][i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            if (SHAPE == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a[0]), "d"(b[0]));
            else if (SHAPE == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                             : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (SHAPE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                             : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < ILP; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += c[u][i];
    if (s == 12345.678) out[0] = s;
}

// plain DFMA for comparison
template <int ILP>
__global__ void __launch_bounds__(1024) kf(double *out, int iters, double seed)
{
    double c[ILP], a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
#pragma unroll
    for (int u = 0; u < ILP; ++u) c[u] = u;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < ILP; ++u) c[u] = fma(a, b, c[u]);
    double s = 0;
#pragma unroll
    for (int u = 0; u < ILP; ++u) s += c[u];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_it(F f)
{
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(s);
        f();
        cudaEventRecord(e);
        cudaEventSynchronize(e);
        float ms; cudaEventElapsedTime(&ms, s, e);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    double *out; cudaMalloc(&out, 64);
    const int iters = 4096;
    const double flops_per[4] = {2.0 * 8 * 8 * 4, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16};
    const char *names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
    printf("device %s, %d SMs\n", pr.name, sms);
    for (int warps = 4; warps <= 32; warps *= 2) {
#define RUN(S, I)                                                                                        \
    {                                                                                                    \
        float ms = time_it([&]() { k<S, I><<<sms, warps * 32>>>(out, iters, 1.0); });                    \
        double tf = flops_per[S] * I * (double)iters * warps * sms / (ms * 1e-3) / 1e12;                 \
        printf("%-9s warps/SM=%2d ilp=%d  %8.3f ms  %7.2f TFLOP/s\n", names[S], warps, I, ms, tf);      \
    }
        RUN(0, 1) RUN(0, 4) RUN(0, 8)
        RUN(1, 1) RUN(1, 4) RUN(1, 8)
        RUN(2, 1) RUN(2, 4) RUN(2, 8)
        RUN(3, 1) RUN(3, 4) RUN(3, 8)
        {
            float ms = time_it([&]() { kf<8><<<sms, warps * 32>>>(out, iters, 1.0); });
            printf("DFMA      warps/SM=%2d ilp=8  %8.3f ms  %7.2f TFLOP/s\n", warps, ms,
                   2.0 * 8 * iters * warps * 32.0 * sms / (ms * 1e-3) / 1e12);
        }
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(err));
    return 0;
}
