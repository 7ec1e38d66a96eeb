// tma_tensor.cu — can the TENSOR-MAP form of TMA (cp.async.bulk.tensor, one instruction per tile) stage the pass
// kernel's tiles as fast as cp.async (LDGSTS, 16 B per lane)?   
//
// A pass tile is 2^t amplitudes (t = 11: 32 KB of complex128) addressed by `a` contiguous low index bits
// (2^a * 16-byte rows) plus t-a FREE bit positions anywhere in the index.  As a tensor map over doubles:
//     dim 0            the row: 2 * 2^a doubles, contiguous                                    box = whole row
//     one dim per RUN  of consecutive free bits [p, p+len): extent 2^len, stride 16 B << p     box = 2^len
//     index dim(s)     the bits that select WHICH tile: box = 1, the coordinate carries the tile index
// cuTensorMap has at most 5 dimensions.  When the regions of index bits between the runs would need more, ALL
// index bits share one dimension of stride 16 B << (lowest index bit) whose coordinate is the tile's base index
// shifted down (the dimensions then overlap in address space, which a load / store does not mind); and a tile
// with more than 3 runs is moved by 2^(bits of the excess runs) instructions, the excess bits riding in that
// same coordinate.  So: 1 instruction per tile for <= 3 runs, 2 / 4 / ... beyond.
//
// The kernel is the data-movement skeleton of k_pass_jit: one persistent CTA per SM, a ring of NBUF tile buffers
// with full / empty mbarriers, tiles copied src -> dst through shared memory.  Four variants per tile shape:
//     load  0 = tensor TMA (1 thread)      1 = cp.async 16 B per lane (128 threads; what the pass kernel does)
//     store 0 = tensor TMA (1 thread)      1 = LDS.128 + STG.128 from 128 threads (what the pass kernel does)
// dst is compared with src over the whole array after the first launch of every variant.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_build/tma_tensor tools/tma_tensor.cu
// Run  : tma_tensor <n> <a> <nbuf> <free bits, comma separated> [<free bits> ...]
//        e.g. tma_tensor 30 3 6 3,4,5,6,7,8,9,10 22,23,24,25,26,27,28,29 11,19,20,21,22,23,27,29
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda.h>
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
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, const int c[5], uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap *map, const int c[5], const void *src) {
    asm volatile(
        "cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
        ::"l"(map), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(smem_u32(src))
        : "memory");
}
__device__ __forceinline__ void cp_async_16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kT = 11;                       // tile bits
constexpr int kGroup = 128;                  // threads of the load group and of the store group
constexpr uint32_t kTileBytes = (1u << kT) * 16u;

struct Params {
    double2 *src, *dst;
    unsigned long long n_tiles;
    unsigned long long index_bits;           // amplitude-index bits that select the tile (the tile id is spread over them)
    int n, a, nbuf, load_mode, store_mode;
    int blk, pair;                           // CTA owns 2^blk consecutive tiles at a time; cp.async loads two tiles at once
    int ops;                                 // TMA instructions per tile
    unsigned long long split_bits;           // the bits the ops of one tile differ in (excess runs), 0 if ops == 1
    unsigned long long dim_mask[5];          // coordinate k = (base & dim_mask[k]) >> dim_shift[k]
    int dim_shift[5];
    int tile_bit[kT];                        // amplitude-index bit of tile position i
};

__device__ __forceinline__ unsigned long long deposit(unsigned long long v, unsigned long long mask) {
    unsigned long long out = 0;
    for (int b = 0; mask; ++b, mask >>= 1)
        if (mask & 1) { out |= (v & 1ull) << b; v >>= 1; }
    return out;
}

// s-th tile of this CTA: blocks of 2^blk consecutive tiles are dealt to the CTAs round-robin (blk = 0: tile s * grid + cta)
__device__ __forceinline__ unsigned long long seq_tile(unsigned long long s, int blk) {
    return ((((s >> blk) * gridDim.x) + blockIdx.x) << blk) + (s & ((1ull << blk) - 1));
}

// 2^PM tiles at once: lanes 0-7 copy a row of tile s, lanes 8-15 the same row of tile s + 1, ... (the next 2^ib
// amplitudes when the CTA owns consecutive tiles), so one warp instruction covers 2^PM x 128 contiguous bytes
template <int PM>
__device__ __forceinline__ void load_multi(const Params &P, double2 *bufs, uint64_t *full, uint64_t *empty, int gt) {
    constexpr int M = 1 << PM, TB = 4 - PM, KB = 4 + PM;
    const int sel = (gt >> 3) & (M - 1);
    unsigned long long po = 0;
    unsigned xlo = gt & 7;
    for (int k = 0; k < 3; ++k) po |= (unsigned long long)((gt >> k) & 1) << P.tile_bit[k];
    for (int k = 0; k < TB; ++k) {
        po |= (unsigned long long)((gt >> (3 + PM + k)) & 1) << P.tile_bit[3 + k];
        xlo |= ((gt >> (3 + PM + k)) & 1) << (3 + k);
    }
    for (unsigned long long s = 0;; s += M) {
        unsigned long long t[M];
#pragma unroll
        for (int j = 0; j < M; ++j) t[j] = seq_tile(s + j, P.blk);
        if (t[0] >= P.n_tiles) break;
#pragma unroll
        for (int j = 0; j < M; ++j) {
            const uint32_t use = (uint32_t)((s + j) / P.nbuf);
            if (t[j] < P.n_tiles && use > 0) mbar_wait(&empty[(s + j) % P.nbuf], (use - 1) & 1);
        }
        unsigned long long mine = t[0];
#pragma unroll
        for (int j = 1; j < M; ++j) if (sel == j) mine = t[j];
        if (mine < P.n_tiles) {
            const double2 *g = P.src + deposit(mine, P.index_bits) + po;
            double2 *d = bufs + (size_t)((s + sel) % P.nbuf) * (1u << kT);
#pragma unroll
            for (int i = 0; i < (1 << KB); ++i) {
                unsigned long long o = 0;
#pragma unroll
                for (int k = 0; k < KB; ++k) if ((i >> k) & 1) o |= 1ull << P.tile_bit[3 + TB + k];
                cp_async_16(d + (i << (3 + TB)) + xlo, g + o);
            }
        }
#pragma unroll
        for (int j = 0; j < M; ++j) if (t[j] < P.n_tiles) cp_async_arrive(&full[(s + j) % P.nbuf]);
    }
}

__global__ void __launch_bounds__(2 * kGroup, 1)
k_copy(const __grid_constant__ CUtensorMap map_src, const __grid_constant__ CUtensorMap map_dst, Params P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    double2 *bufs = reinterpret_cast<double2 *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)P.nbuf * kTileBytes);
    uint64_t *empty = full + P.nbuf;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < P.nbuf; ++b) {
            mbar_init(&full[b], P.load_mode == 0 ? 1 : kGroup);
            mbar_init(&empty[b], P.store_mode == 0 ? 1 : kGroup);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool loader = tid >= kGroup;
    const int gt = tid & (kGroup - 1);
    // tile-local element x = i * 128 + gt  ->  offset inside the state (amplitudes)
    unsigned long long off_t = 0, off_i[16];
    for (int k = 0; k < 7; ++k) off_t |= (unsigned long long)((gt >> k) & 1) << P.tile_bit[k];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        off_i[i] = 0;
        for (int k = 0; k < 4; ++k) off_i[i] |= (unsigned long long)((i >> k) & 1) << P.tile_bit[7 + k];
    }

    if (loader) {
        if (P.load_mode == 0 && gt != 0) return;
        if (P.load_mode == 1 && P.pair == 1) { load_multi<1>(P, bufs, full, empty, gt); return; }
        if (P.load_mode == 1 && P.pair == 2) { load_multi<2>(P, bufs, full, empty, gt); return; }
        for (unsigned long long s = 0;; ++s) {
            const unsigned long long tile = seq_tile(s, P.blk);
            if (tile >= P.n_tiles) break;
            const int b = (int)(s % P.nbuf);
            const uint32_t use = (uint32_t)(s / P.nbuf);
            if (use > 0) mbar_wait(&empty[b], (use - 1) & 1);
            const unsigned long long base = deposit(tile, P.index_bits);
            double2 *d = bufs + (size_t)b * (1u << kT);
            if (P.load_mode == 0) {
                mbar_expect_tx(&full[b], kTileBytes);
                for (int o = 0; o < P.ops; ++o) {
                    const unsigned long long bo = base | deposit((unsigned long long)o, P.split_bits);
                    int c[5];
#pragma unroll
                    for (int k = 0; k < 5; ++k) c[k] = (int)((bo & P.dim_mask[k]) >> P.dim_shift[k]);
                    tma_load_5d(reinterpret_cast<unsigned char *>(d) + (size_t)o * (kTileBytes / P.ops), &map_src, c, &full[b]);
                }
            } else {
                const double2 *g = P.src + base + off_t;
#pragma unroll
                for (int i = 0; i < 16; ++i) cp_async_16(d + i * kGroup + gt, g + off_i[i]);
                cp_async_arrive(&full[b]);
            }
        }
        return;
    }
    // ---- store group ----
    if (P.store_mode == 0 && gt != 0) return;
    int prev_b = -1;
    for (unsigned long long s = 0;; ++s) {
        const unsigned long long tile = seq_tile(s, P.blk);
        if (tile >= P.n_tiles) break;
        const int b = (int)(s % P.nbuf);
        const uint32_t use = (uint32_t)(s / P.nbuf);
        mbar_wait(&full[b], use & 1);
        const unsigned long long base = deposit(tile, P.index_bits);
        double2 *d = bufs + (size_t)b * (1u << kT);
        if (P.store_mode == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int o = 0; o < P.ops; ++o) {
                const unsigned long long bo = base | deposit((unsigned long long)o, P.split_bits);
                int c[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) c[k] = (int)((bo & P.dim_mask[k]) >> P.dim_shift[k]);
                tma_store_5d(&map_dst, c, reinterpret_cast<unsigned char *>(d) + (size_t)o * (kTileBytes / P.ops));
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // two store groups in flight: the previous tile's buffer is free once its group has READ shared memory
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            if (prev_b >= 0) mbar_arrive(&empty[prev_b]);
            prev_b = b;
        } else {
            double2 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = d[i * kGroup + gt];
            mbar_arrive(&empty[b]);
            double2 *o = P.dst + base + off_t;
#pragma unroll
            for (int i = 0; i < 16; ++i) o[off_i[i]] = v[i];
        }
    }
    if (P.store_mode == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void k_fill(double2 *p, unsigned long long n) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        p[i] = make_double2((double)i, -(double)(i ^ 0x5555ull));
}
__global__ void k_compare(const double2 *a, const double2 *b, unsigned long long n, unsigned long long *bad) {
    unsigned long long local = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        local += (a[i].x != b[i].x) || (a[i].y != b[i].y);
    if (local) atomicAdd(bad, local);
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Dim { unsigned long long extent, stride_bytes, mask; int shift; unsigned box; };

// The tensor map of one tile shape.  Returns the number of TMA instructions per tile (0 = cannot be expressed).
static int build_shape(int n, int a, const std::vector<int> &free_bits, Params &P, std::vector<Dim> &dims, bool &aliased) {
    unsigned long long freem = 0;
    for (int b : free_bits) freem |= 1ull << b;
    const unsigned long long all = (n >= 64 ? ~0ull : (1ull << n) - 1);
    const unsigned long long lowm = (1ull << a) - 1;
    P.index_bits = all & ~freem & ~lowm;
    // regions of consecutive bits of one kind above the row
    struct Region { int lo, len; bool is_free; };
    std::vector<Region> regs;
    for (int b = a; b < n;) {
        const bool f = (freem >> b) & 1;
        int e = b;
        while (e < n && (((freem >> e) & 1) != 0) == f) ++e;
        regs.push_back({b, e - b, f});
        b = e;
    }
    dims.clear();
    unsigned row_doubles = 2u << a;
    size_t r0 = 0;
    if (!regs.empty() && regs[0].is_free && (row_doubles << regs[0].len) <= 256) {     // free bits right above the row extend it
        row_doubles <<= regs[0].len;
        r0 = 1;
    }
    dims.push_back({row_doubles, 8, 0, 0, row_doubles});
    int n_runs = 0;
    for (size_t r = r0; r < regs.size(); ++r) n_runs += regs[r].is_free;
    aliased = (regs.size() - r0) > 4;
    P.ops = 1;
    P.split_bits = 0;
    if (!aliased) {
        for (size_t r = r0; r < regs.size(); ++r) {
            const Region &g = regs[r];
            const unsigned long long m = ((1ull << g.len) - 1) << g.lo;
            dims.push_back({1ull << g.len, 16ull << g.lo, g.is_free ? 0 : m, g.lo, g.is_free ? (1u << g.len) : 1u});
        }
    } else {
        std::vector<Region> runs;
        for (size_t r = r0; r < regs.size(); ++r) if (regs[r].is_free) runs.push_back(regs[r]);
        while (runs.size() > 3) {                        // the top runs beyond 3 are walked by separate instructions
            const Region g = runs.back();
            runs.pop_back();
            P.split_bits |= ((1ull << g.len) - 1) << g.lo;
            P.ops <<= g.len;
        }
        int ib = 0;
        while (!((P.index_bits >> ib) & 1)) ++ib;
        const unsigned long long idxm = (P.index_bits | P.split_bits) & ~((1ull << ib) - 1);
        dims.push_back({1ull << (n - ib), 16ull << ib, idxm, ib, 1u});
        for (const Region &g : runs) dims.push_back({1ull << g.len, 16ull << g.lo, 0, g.lo, 1u << g.len});
    }
    std::sort(dims.begin() + 1, dims.end(), [](const Dim &x, const Dim &y) { return x.stride_bytes < y.stride_bytes; });
    if (dims.size() > 5) return 0;
    while (dims.size() < 5) dims.push_back({1, dims.back().stride_bytes * dims.back().extent, 0, 0, 1});
    for (int k = 0; k < 5; ++k) { P.dim_mask[k] = dims[k].mask; P.dim_shift[k] = dims[k].shift; }
    // tile position -> amplitude bit: the row bits, then the free bits ascending (the order TMA writes the box)
    int k = 0;
    for (int b = 0; b < a; ++b) P.tile_bit[k++] = b;
    std::vector<int> fb = free_bits;
    std::sort(fb.begin(), fb.end());
    for (int b : fb) P.tile_bit[k++] = b;
    return k == kT ? P.ops : 0;
}

static std::vector<int> env_list(const char *name) {
    std::vector<int> out;
    const char *e = getenv(name);
    if (e) {
        char tmp[128];
        strncpy(tmp, e, 127);
        tmp[127] = 0;
        char *save = nullptr;
        for (char *tok = strtok_r(tmp, ",", &save); tok; tok = strtok_r(nullptr, ",", &save)) out.push_back(atoi(tok));
    }
    if (out.empty()) out.push_back(0);
    return out;
}

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: tma_tensor n a nbuf bits[,bits...] ...\n"); return 2; }
    const int n = atoi(argv[1]), a = atoi(argv[2]), nbuf = atoi(argv[3]);
    const size_t amps = (size_t)1 << n;
    if (getenv("TMA_DRY")) {                              // host-side check of the shapes, no device needed
        for (int s = 4; s < argc; ++s) {
            std::vector<int> fb;
            char tmp[256];
            strncpy(tmp, argv[s], 255);
            tmp[255] = 0;
            for (char *tok = strtok(tmp, ","); tok; tok = strtok(nullptr, ",")) fb.push_back(atoi(tok));
            Params P{};
            std::vector<Dim> dims;
            bool aliased = false;
            const int ops = (int)fb.size() == kT - a ? build_shape(n, a, fb, P, dims, aliased) : 0;
            printf("%s: ops %d aliased %d split %llx index %llx |", argv[s], ops, (int)aliased, P.split_bits, P.index_bits);
            for (size_t k = 0; k < dims.size(); ++k)
                printf(" [ext %llu str %llu box %u mask %llx sh %d]", dims[k].extent, dims[k].stride_bytes, dims[k].box, dims[k].mask, dims[k].shift);
            printf("\n");
        }
        return 0;
    }
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        printf("{\"error\": \"cuTensorMapEncodeTiled not available\"}\n");
        return 1;
    }
    double2 *src, *dst;
    unsigned long long *bad;
    if (cudaMalloc(&src, amps * 16) != cudaSuccess || cudaMalloc(&dst, amps * 16) != cudaSuccess) {
        printf("{\"error\": \"cudaMalloc\"}\n");
        return 1;
    }
    cudaMalloc(&bad, 8);
    k_fill<<<148 * 8, 256>>>(src, amps);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = (size_t)nbuf * kTileBytes + 2 * nbuf * 8 + 64;
    cudaFuncSetAttribute(k_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);

    for (int s = 4; s < argc; ++s) {
        std::vector<int> fb;
        for (char *tok = strtok(argv[s], ","); tok; tok = strtok(nullptr, ",")) fb.push_back(atoi(tok));
        Params P{};
        P.src = src; P.dst = dst; P.n = n; P.a = a; P.nbuf = nbuf;
        P.n_tiles = amps >> kT;
        std::vector<Dim> dims;
        bool aliased = false;
        const int ops = (int)fb.size() == kT - a ? build_shape(n, a, fb, P, dims, aliased) : 0;
        char shape[128] = "";
        for (size_t i = 0; i < fb.size(); ++i) sprintf(shape + strlen(shape), "%s%d", i ? "," : "", fb[i]);
        if (!ops) { printf("{\"free_bits\": [%s], \"error\": \"shape not expressible\"}\n", shape); continue; }
        CUtensorMap maps[2];
        CUresult res = CUDA_SUCCESS;
        cuuint64_t gdim[5], gstr[4];
        cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
        for (int k = 0; k < 5; ++k) { gdim[k] = dims[k].extent; box[k] = dims[k].box; if (k) gstr[k - 1] = dims[k].stride_bytes; }
        for (int m = 0; m < 2 && res == CUDA_SUCCESS; ++m)
            res = encode(&maps[m], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, m ? (void *)dst : (void *)src, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        char dimtxt[256] = "";
        for (int k = 0; k < 5; ++k)
            sprintf(dimtxt + strlen(dimtxt), "%s[%llu, %llu, %u]", k ? ", " : "", (unsigned long long)dims[k].extent,
                    (unsigned long long)dims[k].stride_bytes, dims[k].box);
        // variants: TMA_ONLY=cp keeps only cp.async + STG (what the pass kernel does); TMA_BLK="0,2,4": tile-to-CTA
        // mappings (blocks of 2^blk consecutive tiles per CTA); TMA_PAIR="0,1": cp.async loads two tiles at once
        const bool only_cp = getenv("TMA_ONLY") && !strcmp(getenv("TMA_ONLY"), "cp");
        std::vector<int> blks = env_list("TMA_BLK"), pairs = env_list("TMA_PAIR");
        for (int lm = only_cp ? 1 : 0; lm < 2; ++lm)
            for (int sm = only_cp ? 1 : 0; sm < 2; ++sm)
                for (int blk : blks)
                    for (int pair : pairs) {
                if (pair && lm != 1) continue;
                if ((lm == 0 || sm == 0) && res != CUDA_SUCCESS) {
                    printf("{\"free_bits\": [%s], \"load\": %d, \"store\": %d, \"error\": \"cuTensorMapEncodeTiled = %d\", "
                           "\"dims_extent_stride_box\": [%s]}\n", shape, lm, sm, (int)res, dimtxt);
                    continue;
                }
                P.load_mode = lm; P.store_mode = sm; P.blk = blk; P.pair = pair;
                cudaMemsetAsync(dst, 0, amps * 16);
                cudaMemsetAsync(bad, 0, 8);
                k_copy<<<sms, 2 * kGroup, smem>>>(maps[0], maps[1], P);
                k_compare<<<148 * 8, 256>>>(src, dst, amps, bad);
                unsigned long long hbad = ~0ull;
                cudaError_t err = cudaMemcpy(&hbad, bad, 8, cudaMemcpyDeviceToHost);
                if (err != cudaSuccess) { printf("{\"free_bits\": [%s], \"error\": \"%s\"}\n", shape, cudaGetErrorString(err)); return 1; }
                float best = 1e9f, sum = 0;
                const int reps = 4;
                for (int rep = 0; rep < reps; ++rep) {
                    cudaEventRecord(e0);
                    k_copy<<<sms, 2 * kGroup, smem>>>(maps[0], maps[1], P);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    best = ms < best ? ms : best;
                    sum += ms;
                }
                printf("{\"n\": %d, \"a\": %d, \"nbuf\": %d, \"free_bits\": [%s], \"load\": \"%s\", \"store\": \"%s\", "
                       "\"blk\": %d, \"pair\": %d, \"tma_ops_per_tile\": %d, \"aliased_index_dim\": %d, \"dims_extent_stride_box\": [%s], "
                       "\"ms_best\": %.3f, \"ms_mean\": %.3f, \"gbs\": %.0f, \"mismatches\": %llu}\n",
                       n, a, nbuf, shape, lm ? "cp.async" : "tma_tensor", sm ? "stg" : "tma_tensor", blk, pair, ops, (int)aliased,
                       only_cp ? "" : dimtxt, best, sum / reps, 2.0 * amps * 16 / (sum / reps) / 1e6, hbad);
                fflush(stdout);
            }
    }
    return 0;
}
