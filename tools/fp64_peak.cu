// Microbenchmark: FP64 FMA / FP32 FMA peak and DMMA (mma.sync m8n8k4 f64) on the box, alone and mixed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void fma_loop(T *out, int iters, T a, T b) {
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_loop(double *out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    for (int it = 0; it < iters; ++it) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

// Do DFMA (FP64 FMA pipe) and DMMA (mma.sync m8n8k4 f64) run on separate pipes, i.e. does a mix of the
// two exceed either peak?  mode 0: every warp interleaves 8 DFMA chains with 4 DMMA chains;
// mode 1: even warps DFMA only, odd warps DMMA only.  Flops are counted per instruction issued.
__global__ void mixed_loop(double *out, int iters, int mode, double fa, double fb) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    const bool do_fma = mode == 0 || ((threadIdx.x >> 5) & 1) == 0;
    const bool do_mma = mode == 0 || ((threadIdx.x >> 5) & 1) == 1;
    for (int it = 0; it < iters; ++it) {
        if (do_fma) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = acc[i] * fa + fb;
        }
        if (do_mma) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
        }
    }
    double s = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    void *out; cudaMalloc(&out, 8 * 148 * 8 * 1024);
    const int blocks = sms * 4, threads = 512, iters = 20000;
    {
        float ms = time_ms([&] { fma_loop<double, 8><<<blocks, threads>>>((double *)out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 8 * iters * (double)blocks * threads;
        printf(", \"fp64_fma_tflops\": %.2f", fl / ms / 1e9);
    }
    {
        float ms = time_ms([&] { fma_loop<float, 8><<<blocks, threads>>>((float *)out, iters, 1.0000001f, 1e-9f); });
        double fl = 2.0 * 8 * iters * (double)blocks * threads;
        printf(", \"fp32_fma_tflops\": %.2f", fl / ms / 1e9);
    }
    {
        float ms = time_ms([&] { dmma_loop<<<blocks, threads>>>((double *)out, iters / 4); });
        double fl = 2.0 * 8 * 8 * 4 * 4.0 * (iters / 4) * (double)blocks * (threads / 32);
        printf(", \"fp64_dmma_tflops\": %.2f", fl / ms / 1e9);
    }
    for (int mode = 0; mode < 2; ++mode) {
        const int it = iters / 4;
        float ms = time_ms([&] { mixed_loop<<<blocks, threads>>>((double *)out, it, mode, 1.0000001, 1e-9); });
        const double warps = (double)blocks * (threads / 32);
        const double share = mode == 0 ? 1.0 : 0.5;                       // fraction of warps doing each kind
        double fl_fma = 2.0 * 8 * it * warps * 32 * share;
        double fl_mma = 2.0 * 8 * 8 * 4 * 4.0 * it * warps * share;
        printf(", \"mixed_mode%d\": {\"ms\": %.3f, \"fma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"sum_tflops\": %.2f}", mode, ms,
               fl_fma / ms / 1e9, fl_mma / ms / 1e9, (fl_fma + fl_mma) / ms / 1e9);
    }
    printf("}\n");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
