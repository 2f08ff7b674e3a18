// Probe (profiles/r04d_stream_rate.txt was taken with a 4 GB buffer: its strided bulk-copy cases beyond 4 GB were
// skipped and the run ended at the first two-CTAs-per-SM tensor-map case; the buffer below covers them all):
// how fast can one producer thread per SM stream HBM into shared memory with cp.async.bulk / TMA boxes,
// as a function of the contiguous chunk size, the ring depth and the stage size?  (No compute: the consumer
// releases a stage as soon as it is full.)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t par) {
    const long long t0 = clock64();
    while (!mbar_try(bar, par)) if (clock64() - t0 > (1ll << 31)) __trap();
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// mode 0: linear bulk copies.  A stage = nchunk chunks of `chunk` bytes; chunk j of a stage comes from "row" j of the
//         CTA's region (rows `rstride` bytes apart), at offset stage_index * chunk inside the row: the wgrad / fc pattern
//         (several rows advance in lock step, each contiguous along its own row).  rstride == chunk: one contiguous stream.
// mode 1: one 2-D TMA box per `boxes` per stage: box = {64 bf16, 128 rows}, row stride rstride; boxes adjacent along the row.
struct P { const uint8_t* src; long long cta_stride; int mode, nst, nchunk, chunk, nstages_total, boxes, nprod; long long rstride; };

__global__ void __launch_bounds__(160) k_stream(const __grid_constant__ CUtensorMap tm, P p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t s_full = sb, s_empty = sb + 128, s_data = sb + 1024;
    const uint32_t stage_bytes = p.mode == 0 ? (uint32_t)p.nchunk * p.chunk : (uint32_t)p.boxes * 16384u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) { mbar_init(s_full + 8 * i, 1); mbar_init(s_empty + 8 * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint8_t* base = p.src + (long long)blockIdx.x * p.cta_stride;
    const int pw = threadIdx.x >> 5;                       // producers: lane 0 of warps 0 .. nprod-1; consumer: warp 4
    if ((threadIdx.x & 31) == 0 && pw < p.nprod) {
        uint32_t st = 0, ph = 0;
        for (int k = 0; k < p.nstages_total; ++k) {
            mbar_wait(s_empty + 8 * st, ph ^ 1);
            const uint32_t bar = s_full + 8 * st, dst = s_data + st * stage_bytes;
            if (pw == 0) mbar_expect(bar, stage_bytes);
            if (p.mode == 0) {
                for (int j = pw; j < p.nchunk; j += p.nprod)
                    bulk_g2s(dst + j * p.chunk, base + (long long)j * p.rstride + (long long)k * p.chunk, p.chunk, bar);
            } else {
                for (int j = pw; j < p.boxes; j += p.nprod)
                    tma_2d(dst + j * 16384, &tm, (k * p.boxes + j) * 64, blockIdx.x * 128, bar);
            }
            if (++st == (uint32_t)p.nst) { st = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 128) {
        uint32_t st = 0, ph = 0;
        for (int k = 0; k < p.nstages_total; ++k) {
            mbar_wait(s_full + 8 * st, ph);
            mbar_arrive(s_empty + 8 * st);
            if (++st == (uint32_t)p.nst) { st = 0; ph ^= 1; }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const size_t BYTES = 8ull << 30;            // the TMA-box cases walk 148 x 2 x 128 rows of 165,376 B = 6.3 GB
    uint8_t* buf; CK(cudaMalloc(&buf, BYTES)); CK(cudaMemset(buf, 1, BYTES));
    uint8_t* flush; CK(cudaMalloc(&flush, 512ull << 20));
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)f;
    CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CUtensorMap tm0; memset(&tm0, 0, sizeof(tm0));
    auto run = [&](const char* name, const CUtensorMap& tm, P p, int grid, size_t smem, double bytes) {
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
            CK(cudaEventRecord(e0));
            k_stream<<<grid, 160, smem>>>(tm, p);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it > 0 && ms < best) best = ms;
        }
        printf("%-86s %8.1f us  %7.0f GB/s\n", name, best * 1e3, bytes / (best * 1e-3) / 1e9);
        fflush(stdout);
    };
    char name[256];
    // ---- mode 0: linear bulk copies, ~80 MB per launch
    for (int per_sm = 1; per_sm <= 2; ++per_sm)
    for (int nst : {2, 4}) {
        for (int chunk : {128, 512, 1024, 2048, 4096, 16384}) {
          for (int nprod : {1, 2, 4}) {
            for (int contiguous = 0; contiguous < 2; ++contiguous) {
                if (nprod > 1 && (chunk > 2048 || contiguous)) continue;
                int stage_bytes = (per_sm == 1 ? 48 : 24) * 1024;
                if (nst == 2) stage_bytes *= 2;
                if (stage_bytes < chunk) continue;
                P p; memset(&p, 0, sizeof(p));
                p.src = buf; p.mode = 0; p.nst = nst; p.chunk = chunk; p.nchunk = stage_bytes / chunk; p.nprod = nprod;
                const int grid = 148 * per_sm;
                const double total = 1200e6;
                p.nstages_total = (int)(total / grid / stage_bytes);
                // strided: row j is a separate 1 MB region (nchunk rows); contiguous: rows back to back inside one stage-sized block
                p.rstride = contiguous ? chunk : (1 << 20);
                if (contiguous) { p.rstride = (long long)p.nstages_total * chunk; }      // each row a contiguous stream, rows adjacent
                p.cta_stride = contiguous ? (long long)p.nchunk * p.rstride : (long long)p.nchunk * (1 << 20);
                if ((double)p.cta_stride * grid > (double)BYTES) { continue; }
                const double bytes = (double)grid * p.nstages_total * stage_bytes;
                snprintf(name, sizeof(name), "bulk  CTAs/SM %d  producers %d  stages %d x %3d KB  chunk %5d B  rows %s", per_sm, nprod, nst, stage_bytes / 1024, chunk,
                         contiguous ? "adjacent (row = nstages*chunk)" : "1 MB apart");
                run(name, tm0, p, grid, 1024 + (size_t)nst * stage_bytes, bytes);
            }
          }
        }
    }
    // ---- mode 1: TMA boxes {64 bf16 = 128 B, 128 rows}; rows 165,376 B apart (the fc forward's A operand)
    {
        const long long rstride = 165376;
        const int rows = 148 * 2 * 128;   // 37,888 rows x 165 KB = 6.3 GB
        const cuuint64_t gd[2] = {(cuuint64_t)(rstride / 2), (cuuint64_t)rows};
        const cuuint64_t gs[1] = {(cuuint64_t)rstride};
        const cuuint32_t bx[2] = {64, 128};
        const cuuint32_t es[2] = {1, 1};
        for (int sw = 0; sw < 2; ++sw) {
            CUtensorMap tm;
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            for (int per_sm = 1; per_sm <= 2; ++per_sm)
            for (int boxes : {1, 2, 4})
            for (int nst : {2, 4, 6, 12})
            for (int nprod : {1, 2}) {
                const int stage_bytes = boxes * 16384;
                if (nprod > boxes) continue;
                if ((size_t)nst * stage_bytes * per_sm > 200 * 1024) continue;
                P p; memset(&p, 0, sizeof(p));
                p.src = buf; p.mode = 1; p.nst = nst; p.boxes = boxes; p.nprod = nprod;
                const int grid = 148 * per_sm;
                p.nstages_total = (int)(1200e6 / grid / stage_bytes);
                const double bytes = (double)grid * p.nstages_total * stage_bytes;
                snprintf(name, sizeof(name), "TMA box 128 rows x 128 B (rows 165 KB apart) swizzle %d  CTAs/SM %d  producers %d  stages %d x %d boxes", sw, per_sm, nprod, nst, boxes);
                run(name, tm, p, grid, 1024 + (size_t)nst * stage_bytes, bytes);
            }
        }
    }
    return 0;
}
