// Micro-benchmark (not product code): tcgen05.ld throughput and its interference with
// tcgen05.mma on one SM.  mode bit 0: MMA warp issues `iters` M128 N32 K16 MMAs;
// mode bit 1: `nld_warps` warps each do `lds` tcgen05.ld.32x32b.x32 (4 KB per warp per ld).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I curla_b200/csrc -o scratch/tmem_rate profiles/tools/tmem_rate.cu
#include "tc.cuh"
#include <cstdio>
using namespace curla;

__global__ void __launch_bounds__(32 * 18, 1) k_rate(int mode, int iters, int lds, int nld_warps, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_bar = s_base, s_tptr = s_base + 8, s_a = s_base + 1024, s_b = s_a + 64 * 1024;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (128 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + 1024)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(s_bar, 1); fence_mbar_init(); }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tptr), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 8);
    if (warp == 16) {
        if ((mode & 1) && elect_one()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t a_hi = make_desc(0, 8192, 128), b_hi = make_desc(0, 512, 128);
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint64_t ad = a_hi | (uint64_t)(((s_a >> 4) + (uint32_t)(i & 63)) & 0x3FFFu);
                const uint64_t bd = b_hi | (uint64_t)((s_b >> 4) & 0x3FFFu);
                umma_bf16_rt(tmem + (uint32_t)((i / 18) & 7) * 32u, ad, bd, idesc, 1u);
            }
            umma_commit(s_bar);
            mbar_wait(s_bar, 0);
            out[0] = clock64() - t0;
        }
        __syncwarp();
    } else if (warp < nld_warps && (mode & 2)) {
        uint32_t r[32];
        uint32_t acc = 0;
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)((warp >> 2) & 3) * 32u;
        const long long t0 = clock64();
        for (int i = 0; i < lds; ++i) {
            tmem_ld32(taddr, r);
#pragma unroll
            for (int k = 0; k < 32; ++k) acc ^= r[k];
        }
        const long long t1 = clock64();
        if (lane == 0) out[1 + warp] = t1 - t0;
        if (acc == 0x12345678u) out[40] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int main() {
    long long* d; cudaMalloc(&d, 64 * 8);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4096;
    for (int nw : {4, 8, 16})
        for (int mode : {1, 2, 3}) {
            // lds chosen so that the ld loop lasts about as long as the MMA loop
            const int lds = 4096 * 2 / (nw / 4);
            cudaMemset(d, 0, 64 * 8);
            k_rate<<<1, 32 * 18, 200 * 1024>>>(mode, iters, lds, nw, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[64]; cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < nw; ++i) mx = h[1 + i] > mx ? h[1 + i] : mx;
            printf("ld warps %2d mode %d: MMA %7.1f clk/MMA | ld: %d x 4KB per warp, %8lld clk -> %6.1f B/clk/SM  (%s)\n", nw, mode,
                   (double)h[0] / iters, lds, mx, mx ? (double)lds * 4096.0 * nw / mx : 0.0, cudaGetErrorString(e));
        }
    return 0;
}
