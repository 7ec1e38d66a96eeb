// xchg_bench.cu — microbenchmark of csrc/xchg.cuh on 2 GPUs of one box (one process, peer access):
// NVLink GB/s per direction of the in-place TMA exchange kernel as a function of the number of SMs it
// gets and of the run size, alone and beside a kernel that streams HBM on the remaining SMs (the
// situation inside a pipelined stage transition).  Verifies the exchanged data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/xchg_bench tools/xchg_bench.cu && tools/_build/xchg_bench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../quantum_simulations_b200/csrc/xchg.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_fill(ulonglong2 *p, size_t n, unsigned long long tag) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_ulonglong2(tag | i, ~i);
}
// after the swap of the top local bit: element i of rank r with top bit b holds (rank b, index with top bit = r)
__global__ void k_check(const ulonglong2 *p, size_t n, int rank, int top, unsigned long long *bad) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)((i >> top) & 1);
        const size_t src = (i & ~((size_t)1 << top)) | ((size_t)rank << top);
        const unsigned long long want = ((unsigned long long)b << 56) | src;
        if (p[i].x != want || p[i].y != ~src) atomicAdd(bad, 1ull);
    }
}
// a stand-in for the pass kernel: persistent CTAs streaming read+write over HBM
__global__ void __launch_bounds__(512, 1) k_stream(ulonglong2 *p, size_t n, int reps) {
    extern __shared__ unsigned char hog[];
    (void)hog;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            ulonglong2 v = p[i]; v.x ^= 1; p[i] = v;
        }
}

int main(int argc, char **argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("{\"error\": \"needs 2 GPUs\"}\n"); return 0; }
    const int n_local = argc > 1 ? atoi(argv[1]) : 28;           // 2^28 x 16 B = 4 GiB per GPU
    const size_t n = (size_t)1 << n_local;
    char *buf[2]; ulonglong2 *side[2]; unsigned long long *bad[2];
    cudaStream_t sx[2], sc[2];
    cudaEvent_t e0[2], e1[2], c0[2], c1[2];
    for (int r = 0; r < 2; ++r) {
        CK(cudaSetDevice(r));
        CK(cudaDeviceEnablePeerAccess(1 - r, 0));
        CK(cudaMalloc(&buf[r], n * 16 + qsvx::kTailBytes));
        CK(cudaMalloc(&side[r], n * 16));
        CK(cudaMalloc(&bad[r], 8));
        CK(cudaMemset(buf[r] + n * 16, 0, qsvx::kTailBytes));
        CK(cudaStreamCreateWithFlags(&sx[r], cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&sc[r], cudaStreamNonBlocking));
        CK(cudaEventCreate(&e0[r])); CK(cudaEventCreate(&e1[r])); CK(cudaEventCreate(&c0[r])); CK(cudaEventCreate(&c1[r]));
        CK(cudaFuncSetAttribute(qsvx::k_xchg_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsvx::kXchgMaxSmem));
        CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    unsigned long long seq = 0;
    const int top = n_local - 1;
    auto run = [&](int sms, int run_log2, int mode /*0 tma 1 ldst*/, bool beside, bool verify, unsigned nw = 4, unsigned nr = 10, unsigned nl = 4, unsigned sl2 = 12) {
        ++seq;
        for (int r = 0; r < 2; ++r) {
            CK(cudaSetDevice(r));
            if (verify) { k_fill<<<592, 512, 0, sx[r]>>>((ulonglong2 *)buf[r], n, (unsigned long long)r << 56); CK(cudaMemsetAsync(bad[r], 0, 8, sx[r])); }
            CK(cudaStreamSynchronize(sx[r]));
        }
        for (int r = 0; r < 2; ++r) {
            CK(cudaSetDevice(r));
            qsvx::XchgArgs A = {};
            A.mine = buf[r];
            for (int d = 0; d < 2; ++d) { A.peer[d] = buf[d]; A.peer_flags[d] = (unsigned long long *)(buf[d] + n * 16); A.rank_of[d] = d; }
            A.my_flags = (unsigned long long *)(buf[r] + n * 16);
            A.my_rank = r; A.n_peers = 2; A.me = r; A.n_special = 1; A.pos[0] = top; A.swap_pos[0] = top; A.chunk_val = 0;
            A.elem_log2 = 4; A.run_log2 = run_log2; A.stage_log2 = 14;
            A.half_elems = n >> 2;                                   // pair = 2^(n_local-1) elements, half of it
            A.units_per_half = (A.half_elems * 16) >> A.stage_log2;
            A.seq = seq; A.timeout_ns = 5000000000ull; A.n_warps = nw; A.n_remote = nr; A.n_local = nl; A.stage_log2 = sl2; A.units_per_half = (A.half_elems * 16) >> A.stage_log2; if (A.run_log2 > A.stage_log2) A.run_log2 = A.stage_log2;
            if (beside) { CK(cudaEventRecord(c0[r], sc[r])); k_stream<<<148 - sms, 512, 200 * 1024, sc[r]>>>(side[r], n, 2); CK(cudaEventRecord(c1[r], sc[r])); }
            CK(cudaEventRecord(e0[r], sx[r]));
            if (mode == 0) qsvx::k_xchg_tma<<<sms, qsvx::kXchgThreads, qsvx::xchg_smem_bytes(sl2, nw, nr, nl), sx[r]>>>(A);
            else qsvx::k_xchg_ldst<4><<<sms, 1024, 0, sx[r]>>>(A);
            CK(cudaGetLastError());
            CK(cudaEventRecord(e1[r], sx[r]));
        }
        float ms = 0, cms = 0;
        unsigned long long nbad = 0;
        for (int r = 0; r < 2; ++r) {
            CK(cudaSetDevice(r));
            CK(cudaStreamSynchronize(sx[r])); CK(cudaStreamSynchronize(sc[r]));
            float t; CK(cudaEventElapsedTime(&t, e0[r], e1[r])); ms = t > ms ? t : ms;
            if (beside) { CK(cudaEventElapsedTime(&t, c0[r], c1[r])); cms = t > cms ? t : cms; }
            if (verify) {
                k_check<<<592, 512, 0, sx[r]>>>((const ulonglong2 *)buf[r], n, r, top, bad[r]);
                unsigned long long b; CK(cudaMemcpyAsync(&b, bad[r], 8, cudaMemcpyDeviceToHost, sx[r])); CK(cudaStreamSynchronize(sx[r])); nbad += b;
            }
        }
        const double sent = (double)(n / 2) * 16;                    // bytes leaving each GPU
        printf("{\"kernel\": \"%s\", \"sms\": %d, \"run_bytes\": %d, \"beside_stream\": %d, \"ms\": %.3f, \"gbs_per_direction\": %.1f, "
               "\"stream_ms\": %.3f, \"stream_gbs\": %.1f, \"verified\": %d, \"bad\": %llu, \"warps_remote_local\": [%u, %u, %u], \"stage_bytes\": %u}\n",
               mode == 0 ? "tma" : "ldst", sms, 1 << run_log2, (int)beside, ms, sent / ms / 1e6, cms,
               beside ? 2.0 * 2 * n * 16 / cms / 1e6 : 0.0, (int)verify, nbad, nw, nr, nl, 1u << sl2);
        fflush(stdout);
    };
    run(16, 14, 0, false, true);                                     // correctness first
    run(16, 9, 0, false, true);                                      // 512-byte runs: 32 bulk copies per half
    run(16, 12, 1, false, true);
    for (int sms : {6, 8, 12, 16}) {
        run(sms, 14, 0, false, false, 4, 10, 4, 12);
        run(sms, 14, 0, false, false, 4, 11, 3, 12);
        run(sms, 14, 0, false, false, 4, 7, 7, 12);
        run(sms, 14, 0, false, false, 2, 10, 4, 13);
        run(sms, 14, 0, false, false, 2, 21, 7, 12);
        run(sms, 14, 0, false, false, 1, 10, 4, 14);
    }
    for (int sms : {8, 12, 16}) { run(sms, 14, 0, true, false, 4, 10, 4, 12); run(sms, 14, 0, true, false, 2, 10, 4, 13); }
    for (int sms : {16, 32}) run(sms, 12, 1, false, false);
    return 0;
}
