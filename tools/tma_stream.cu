// tma_stream.cu — data-movement skeleton of the persistent pass kernel, as a microbenchmark.
//
// One persistent CTA per SM: a producer warp streams tiles HBM -> shared memory with
// cp.async.bulk (TMA, one bulk copy per contiguous row of the tile) into a ring of NBUF
// buffers guarded by full/empty mbarriers; G consumer groups each take every G-th tile,
// pull it into registers (16 amplitudes per thread), optionally burn `delay` cycles and do
// `xch` shared-memory round trips in place (the inter-round exchange of the real kernel),
// then store registers -> HBM with coalesced STG.  Reports achieved GB/s (read + write).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_build/tma_stream tools/tma_stream.cu
// Run  : tma_stream <n> <tile_log2> <row_log2> <nbuf> <groups> <delay_cycles> <xch_rounds> [ctas_per_sm]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Params {
    double2 *src, *dst;
    int n, t, a;          // state bits, tile bits, contiguous low bits (row = 2^a amps)
    int nbuf, groups, delay, xch;
    unsigned long long n_tiles;
};

// tile id -> base amplitude index: tile bits = low a bits + top (t-a) bits; the tile id
// fills the middle bits.
__device__ __forceinline__ unsigned long long tile_base(const Params &P, unsigned long long tile) {
    return tile << P.a;
}
__device__ __forceinline__ unsigned long long row_offset(const Params &P, unsigned r) {
    return (unsigned long long)r << (P.n - (P.t - P.a));
}

template <int MAXT, int T>
__global__ void __launch_bounds__(MAXT, 1) k_stream(Params P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int A = P.a;
    constexpr uint32_t tile_amps = 1u << T, tile_bytes = tile_amps * 16u;
    double2 *bufs = reinterpret_cast<double2 *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)P.nbuf * tile_bytes);
    uint64_t *empty = full + P.nbuf;
    constexpr int gthreads = tile_amps / 16;             // threads per consumer group
    const int n_cons = gthreads * P.groups;
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int b = 0; b < P.nbuf; ++b) { mbar_init(&full[b], 1); mbar_init(&empty[b], gthreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const unsigned long long first = blockIdx.x, stride = gridDim.x;
    if (tid >= n_cons) {
        // ---------------- producer warp ----------------
        const int lane = tid - n_cons;
        if (lane < 32) {
            const unsigned rows = 1u << (T - A);
            const uint32_t row_bytes = 16u << A;
            unsigned long long s = 0;
            for (unsigned long long tile = first; tile < P.n_tiles; tile += stride, ++s) {
                const int b = (int)(s % P.nbuf);
                const uint32_t use = (uint32_t)(s / P.nbuf);
                if (use > 0) mbar_wait(&empty[b], (use - 1) & 1);
                if (lane == 0) mbar_expect_tx(&full[b], tile_bytes);
                __syncwarp();
                const double2 *g = P.src + tile_base(P, tile);
                double2 *d = bufs + (size_t)b * tile_amps;
                for (unsigned r = lane; r < rows; r += 32)
                    bulk_g2s(d + ((size_t)r << A), g + row_offset(P, r), row_bytes, &full[b]);
            }
        }
        return;
    }
    // ---------------- consumer groups ----------------
    const int grp = tid / gthreads, gt = tid % gthreads;
    unsigned long long s = grp;
    for (unsigned long long tile = first + (unsigned long long)grp * stride; tile < P.n_tiles;
         tile += stride * P.groups, s += P.groups) {
        const int b = (int)(s % P.nbuf);
        const uint32_t use = (uint32_t)(s / P.nbuf);
        double2 *buf = bufs + (size_t)b * tile_amps;
        mbar_wait(&full[b], use & 1);
        double2 v[16];
        // registers <- tile: thread gt owns amplitudes gt + j * gthreads (register bits = top 4)
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = buf[gt + j * gthreads];
        for (int x = 0; x < P.xch; ++x) {
            asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(gthreads));
            const int rot = (x * 37 + 5) % gthreads;
#pragma unroll
            for (int j = 0; j < 16; ++j) buf[((gt + rot) % gthreads) + j * gthreads] = v[j];
            asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(gthreads));
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = buf[gt + j * gthreads];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&empty[b]);
        if (P.delay) {
            long long t0 = clock64();
            while (clock64() - t0 < P.delay) { }
        }
        // registers -> HBM (same addressing as the load: identity)
        double2 *o = P.dst + tile_base(P, tile);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const unsigned x = gt + j * gthreads;         // tile index
            const unsigned r = x >> A, c = x & ((1u << A) - 1);
            o[row_offset(P, r) + c] = v[j];
        }
    }
}

int main(int argc, char **argv) {
    Params P{};
    P.n = argc > 1 ? atoi(argv[1]) : 28;
    P.t = argc > 2 ? atoi(argv[2]) : 11;
    P.a = argc > 3 ? atoi(argv[3]) : 5;
    P.nbuf = argc > 4 ? atoi(argv[4]) : 6;
    P.groups = argc > 5 ? atoi(argv[5]) : 3;
    P.delay = argc > 6 ? atoi(argv[6]) : 0;
    P.xch = argc > 7 ? atoi(argv[7]) : 0;
    const int ctas_per_sm = argc > 8 ? atoi(argv[8]) : 1;
    const size_t amps = (size_t)1 << P.n;
    P.n_tiles = amps >> P.t;
    cudaMalloc(&P.src, amps * 16);
    cudaMalloc(&P.dst, amps * 16);
    cudaMemset(P.src, 1, amps * 16);
    const int threads = (1 << (P.t - 4)) * P.groups + 32;
    const size_t smem = ((size_t)P.nbuf << P.t) * 16 + 2 * P.nbuf * 8 + 64;
    auto kern = P.t == 11 ? k_stream<544, 11> : (P.t == 12 ? k_stream<800, 12> : k_stream<544, 10>);
    if (threads > (P.t == 12 ? 800 : 544)) { printf("{\"error\": \"too many threads\"}\n"); return 1; }
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev_sms = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = dev_sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, threads, smem>>>(P);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(err)); return 1; }
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        kern<<<grid, threads, smem>>>(P);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // verify identity copy on a sample
    unsigned char hs[64], hd[64];
    cudaMemcpy(hs, (char *)P.src + amps * 8, 64, cudaMemcpyDeviceToHost);
    cudaMemcpy(hd, (char *)P.dst + amps * 8, 64, cudaMemcpyDeviceToHost);
    int ok = 1;
    for (int i = 0; i < 64; ++i) ok &= hs[i] == hd[i];
    printf("{\"n\": %d, \"t\": %d, \"a\": %d, \"nbuf\": %d, \"groups\": %d, \"delay\": %d, \"xch\": %d, \"ctas_per_sm\": %d, "
           "\"threads\": %d, \"smem\": %zu, \"ms\": %.3f, \"gbs\": %.0f, \"ok\": %d}\n",
           P.n, P.t, P.a, P.nbuf, P.groups, P.delay, P.xch, ctas_per_sm, threads, smem, best,
           2.0 * amps * 16 / best / 1e6, ok);
    return 0;
}
