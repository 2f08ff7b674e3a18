// Micro-benchmark (not product code): clocks per tcgen05.mma (SS mode, bf16, M=128, K=16,
// no-swizzle K-major operands) as a function of N, issued back to back by one thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I curla_b200/csrc -o /tmp/mma_rate profiles/tools/mma_rate.cu
#include "tc.cuh"
#include <cstdio>
using namespace curla;

__global__ void __launch_bounds__(128, 1) k_rate(int N, int iters, int a_step16, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_bar = s_base, s_tptr = s_base + 8, s_a = s_base + 1024, s_b = s_a + 64 * 1024;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem + 1024)[i] = make_uint4(0, 0, 0, 0);
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
    if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
            // A: planes of 4 KB rows*16B: lbo = 8192 (plane), sbo = 128; B: lbo = N*16, sbo = 128
            const uint64_t a_hi = make_desc(0, 8192, 128), b_hi = make_desc(0, (uint32_t)N * 16u, 128);
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint64_t ad = a_hi | (uint64_t)(((s_a >> 4) + (uint32_t)((i & 7) * a_step16)) & 0x3FFFu);
                const uint64_t bd = b_hi | (uint64_t)((s_b >> 4) & 0x3FFFu);
                umma_bf16_rt(tmem, ad, bd, idesc, 1u);
            }
            umma_commit(s_bar);
            mbar_wait(s_bar, 0);
            const long long t1 = clock64();
            out[blockIdx.x] = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int main() {
    long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4096;
    for (int grid : {1, 148})
        for (int step : {0, 1, 8})
            for (int N : {16, 32, 64, 96, 128, 192, 256}) {
                k_rate<<<grid, 128, 200 * 1024>>>(N, iters, step, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
                long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("grid %3d a_step %d rows  N %3d : %7.1f clk/MMA  (%s)\n", grid, step, N, (double)mx / iters, cudaGetErrorString(e));
            }
    return 0;
}
