// Probe (B200): semantics and cost of tcgen05.shift, tcgen05.cp.128x256b and tcgen05.mma with the A operand in TMEM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../curla_b200/csrc probe_shift.cu -o probe_shift
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "tc.cuh"
using namespace curla;

namespace curla {
void set_last_error(const char*, ...) {}
int check_launch(const char*) { return 0; }
void set_launch_tag(const char*) {}
bool pdl_enabled() { return false; }
int sm_count() { return 148; }
int conv_grid_cap() { return 148; }
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// out layout (uint32 words): test t at out + t * 128 * 64: [lane][64 words]
__global__ void __launch_bounds__(128, 1) k_probe(uint32_t* out, long long* clk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_bar = s_base, s_tptr = s_base + 16;
    const uint32_t s_a = s_base + 1024;               // A slab: 2 planes x 160 rows x 16 B (plane stride PS)
    const uint32_t PS = 160 * 16;
    const uint32_t s_b = s_a + 2 * PS;                // B: [2 k chunks][32 n][8 k] bf16 (K-major, no swizzle) = 1024 B
    if (tid == 0) { mbar_init(s_bar, 1); fence_mbar_init(); }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tptr), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // A[row r][k] = bf16(r + k/16.0 ... ) : small integers exactly representable: value = (r % 64) + 64 * (k & 1)  (k = 0..15)
    bf16* A = reinterpret_cast<bf16*>(smem + 1024);
    for (int i = tid; i < 2 * 160 * 8; i += 128) {
        const int c = i / (160 * 8), rem = i - c * 160 * 8, r = rem >> 3, j = rem & 7, k = c * 8 + j;
        A[i] = __float2bfloat16((float)((r % 64) + (k == 3 ? 64 : 0)));
    }
    // B[n][k] = 1 if k == n % 16 (n < 32) -> D[r][n] = A[r][n % 16]
    bf16* Bm = reinterpret_cast<bf16*>(smem + 1024 + 2 * PS);
    for (int i = tid; i < 2 * 32 * 8; i += 128) {
        const int kc = i / 256, rem = i - kc * 256, n = rem >> 3, j = rem & 7, k = kc * 8 + j;
        Bm[i] = __float2bfloat16(k == (n % 16) ? 1.f : 0.f);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *reinterpret_cast<volatile uint32_t*>(smem + 16);
    const uint32_t lane_base = ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;
    auto sync_mma = [&]() {      // thread 0 commits, everyone waits
        if (tid == 0) umma_commit(s_bar);
        mbar_wait(s_bar, phase);
        phase ^= 1;
        tc_fence_after();
    };
    auto dump = [&](int test, uint32_t col0, int ncols8) {
        for (int c8 = 0; c8 < ncols8; ++c8) {
            uint32_t r[8];
            tmem_ld8(tb + lane_base + col0 + c8 * 8, r);
            for (int i = 0; i < 8; ++i) out[(size_t)test * 128 * 64 + (size_t)tid * 64 + c8 * 8 + i] = r[i];
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    };
    // ---------- test 0/1: fill columns [0,16) with lane*100 + col, shift.down at column 0 (lane 0), dump before / after
    {
        for (int c8 = 0; c8 < 2; ++c8) {
            uint32_t r[8];
            for (int i = 0; i < 8; ++i) r[i] = (uint32_t)(tid * 100 + c8 * 8 + i);
            tmem_st8(tb + lane_base + c8 * 8, r);
        }
        tc_fence_before(); __syncthreads(); tc_fence_after();
        dump(0, 0, 2);
        if (tid == 0) asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tb) : "memory");
        sync_mma();
        dump(1, 0, 2);
        // test 2: a second shift at column 8, lane 32
        if (tid == 0) asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tb + (32u << 16) + 8u) : "memory");
        sync_mma();
        dump(2, 0, 2);
    }
    // ---------- test 3: tcgen05.cp 128x256b of A rows [0,128) (descriptor: LBO = PS, SBO = 128) into columns [64, 72)
    {
        const uint64_t ad = make_desc(s_a, PS, 128);
        if (tid == 0) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tb + 64u), "l"(ad) : "memory");
        sync_mma();
        dump(3, 64, 1);
    }
    // ---------- test 4: D = A(smem) x B  (SS mode, columns [128,160)); test 5: D = A(tmem cols 64..71) x B (columns [160,192))
    {
        const uint64_t ad = make_desc(s_a, PS, 128), bd = make_desc(s_b, 512, 128);
        if (tid == 0) umma_bf16_rt(tb + 128u, ad, bd, kIdesc, 0u);
        sync_mma();
        dump(4, 128, 4);
        if (tid == 0) umma_ts(tb + 160u, tb + 64u, bd, kIdesc, 0u);
        sync_mma();
        dump(5, 160, 4);
        // test 6: shift the TMEM copy of A down by one row, multiply again (columns [192,224))
        if (tid == 0) asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tb + 64u) : "memory");
        sync_mma();
        if (tid == 0) umma_ts(tb + 192u, tb + 64u, bd, kIdesc, 0u);
        sync_mma();
        dump(6, 192, 4);
    }
    // ---------- timing: back-to-back accumulating MMAs (SS vs TS), shifts, copies; one commit at the end each
    {
        const uint64_t ad = make_desc(s_a, PS, 128), bd = make_desc(s_b, 512, 128);
        const int N = 256;
        for (int variant = 0; variant < 5; ++variant) {
            __syncthreads();
            const long long t0 = clock64();
            if (tid == 0) {
                for (int i = 0; i < N; ++i) {
                    if (variant == 0) umma_bf16_rt(tb + 256u, ad, bd, kIdesc, 1u);
                    else if (variant == 1) umma_ts(tb + 256u, tb + 64u, bd, kIdesc, 1u);
                    else if (variant == 2) asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tb + 64u) : "memory");
                    else if (variant == 3) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tb + 64u), "l"(ad) : "memory");
                    else { asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tb + 64u) : "memory"); umma_ts(tb + 256u, tb + 64u, bd, kIdesc, 1u); }
                }
            }
            sync_mma();
            if (tid == 0) clk[variant] = (clock64() - t0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
    }
}

int main() {
    uint32_t* d_out; long long* d_clk;
    const size_t n = 7 * 128 * 64;
    cudaMalloc(&d_out, n * 4); cudaMemset(d_out, 0xFF, n * 4);
    cudaMalloc(&d_clk, 8 * 8); cudaMemset(d_clk, 0, 64);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    k_probe<<<1, 128, 16384>>>(d_out, d_clk);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    std::vector<uint32_t> h(n);
    cudaMemcpy(h.data(), d_out, n * 4, cudaMemcpyDeviceToHost);
    long long clk[8];
    cudaMemcpy(clk, d_clk, 64, cudaMemcpyDeviceToHost);
    const int lanes[] = {0, 1, 2, 3, 30, 31, 32, 33, 34, 63, 64, 65, 95, 96, 97, 126, 127};
    auto show_u = [&](int t, const char* what, int ncols) {
        printf("== test %d: %s\n", t, what);
        for (int l : lanes) {
            printf("  lane %3d:", l);
            for (int c = 0; c < ncols; ++c) printf(" %6u", h[(size_t)t * 128 * 64 + (size_t)l * 64 + c]);
            printf("\n");
        }
    };
    auto show_bf = [&](int t, const char* what, int nwords) {
        printf("== test %d: %s (bf16 pairs)\n", t, what);
        for (int l : lanes) {
            printf("  lane %3d:", l);
            for (int c = 0; c < nwords; ++c) {
                uint32_t w = h[(size_t)t * 128 * 64 + (size_t)l * 64 + c];
                uint32_t lo = w << 16, hi = w & 0xFFFF0000u; float a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4);
                printf(" (%g,%g)", a, b);
            }
            printf("\n");
        }
    };
    auto show_f = [&](int t, const char* what, int ncols) {
        printf("== test %d: %s (fp32)\n", t, what);
        for (int l : lanes) {
            printf("  lane %3d:", l);
            for (int c = 0; c < ncols; ++c) { float f; memcpy(&f, &h[(size_t)t * 128 * 64 + (size_t)l * 64 + c], 4); printf(" %g", f); }
            printf("\n");
        }
    };
    show_u(0, "columns 0..15 = lane*100 + col, before any shift", 16);
    show_u(1, "after tcgen05.shift.down [lane 0, col 0]", 16);
    show_u(2, "after a second shift [lane 32, col 8]", 16);
    show_bf(3, "tcgen05.cp.128x256b of A rows 0..127 -> 8 columns: A[r][k] = r%64 (+64 at k=3)", 8);
    show_f(4, "SS MMA: D[r][n] = A[r][n%16], n = 0..31", 32);
    show_f(5, "TS MMA (A from the TMEM copy): same D expected", 32);
    show_f(6, "TS MMA after shifting the TMEM copy of A down once", 32);
    printf("clk per op (256 back to back + commit): SS mma N32 %.1f | TS mma N32 %.1f | shift %.1f | cp 128x256b %.1f | shift+TS mma %.1f\n",
           clk[0] / 256.0, clk[1] / 256.0, clk[2] / 256.0, clk[3] / 256.0, clk[4] / 256.0);
    return 0;
}
