// Micro-benchmark (not product code): clocks per tcgen05.mma (SS mode, bf16, M=128, K=16) with MN-MAJOR operands in the
// no-swizzle layout the conv weight-gradient kernel uses (8 channels = 16 B contiguous, K rows 16 B apart, M / N chunks a
// constant stride apart: conv_wgrad_tc.cu), as a function of N and of the chunk strides, issued back to back by one thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I curla_b200/csrc -o scratch/probe/mma_rate_mn profiles/tools/mma_rate_mn.cu
#include "tc.cuh"
#include <cstdio>
using namespace curla;

// mode 0: A and B MN-major (wgrad); mode 1: A and B K-major no swizzle (conv forward) for reference
__global__ void __launch_bounds__(128, 1) k_rate(int mode, int N, int iters, uint32_t a_sbo, uint32_t b_sbo, int kstep16, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_bar = s_base, s_tptr = s_base + 8, s_a = s_base + 1024, s_b = s_a + 96 * 1024;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (200 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem + 1024)[i] = make_uint4(0, 0, 0, 0);
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
            uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
            uint64_t a_hi, b_hi;
            if (mode == 0) {
                idesc |= (1u << 15) | (1u << 16);
                a_hi = make_desc(0, 128, a_sbo); b_hi = make_desc(0, 128, b_sbo);
            } else {
                a_hi = make_desc(0, 8192, 128); b_hi = make_desc(0, (uint32_t)N * 16u, 128);
            }
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint64_t ad = a_hi | (uint64_t)(((s_a >> 4) + (uint32_t)((i & 3) * kstep16)) & 0x3FFFu);
                const uint64_t bd = b_hi | (uint64_t)(((s_b >> 4) + (uint32_t)((i & 3) * kstep16)) & 0x3FFFu);
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
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    const int iters = 4096;
    auto run = [&](int mode, int N, uint32_t asbo, uint32_t bsbo, int kstep) {
        k_rate<<<148, 128, 201 * 1024>>>(mode, N, iters, asbo, bsbo, kstep, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%s  N %3d  A chunk stride %5u B  B chunk stride %5u B  K window step %3d rows : %7.1f clk/MMA  (%s)\n",
               mode == 0 ? "MN-major" : "K-major ", N, asbo, bsbo, kstep, (double)mx / iters, cudaGetErrorString(e));
    };
    for (int N : {32, 64, 96}) run(1, N, 0, 0, 0);
    for (int kstep : {0, 16})
        for (int N : {32, 64, 96, 128})
            for (uint32_t asbo : {1088u, 1024u, 1152u, 1280u, 2048u, 4096u})
                for (uint32_t bsbo : {1280u, 1024u, 1152u, 1088u}) {
                    if (asbo * 16 > 90 * 1024 || bsbo * (N / 8) > 90 * 1024) continue;
                    if (asbo != 1088u && bsbo != 1280u && asbo != bsbo) continue;
                    run(0, N, asbo, bsbo, kstep);
                }
    return 0;
}
