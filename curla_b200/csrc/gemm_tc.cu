// K7/K9 on the 5th-generation tensor cores: the bf16 GEMM of gemm.cu for problems with at least
// 128 rows -- the MLP trunks (encoder.py:98 -> curl_sac.py:70-74,129-133: 512 x 1024 x 1024) and the
// encoder fc forward / dgrad / wgrad (512 x 64 x 67,456 and transposes) -- issued with tcgen05.mma,
// accumulators in TMEM.  Same argument block, same four operand layouts, same epilogues
// (alpha, bias, ReLU, ReLU mask, bf16 / fp32, split-K partials, batched, segmented operands).
//
// One CTA (256 threads) = one 128 x 64 output tile, K in steps of 64, four stages:
//   * all threads copy the stage's operand tiles with 16-byte cp.async straight into the
//     tensor core's canonical no-swizzle layouts (any row stride / segmented index in global
//     memory; zero fill past the edges):
//        K-major  operand (stored [rows][K]):  chunk (row r, 8-element K chunk kc) at kc*LBO + r*16
//                 -> core matrix = 8 rows x 16 B, SBO = 128 B, LBO = rows*16 + 16
//        MN-major operand (stored [K][cols]):  chunk (k, 8-column chunk c)        at c*SBO + k*16
//                 -> LBO = 128 B (8 K rows), SBO = 64*16 + 16
//     (the +16 makes the eight chunks a quarter-warp writes land in eight different bank groups);
//   * cp.async.wait_group + fence.proxy.async + one __syncthreads per K step hand the stage to the
//     async proxy; thread 0 then issues the step's four M=128 N=64 K=16 MMAs and commits them to
//     the stage's mbarrier, which the copy of step k+3 into the same slot waits for;
//   * after the last commit each warp reads 32 TMEM lanes x 32 columns (one output row per thread)
//     and applies the epilogue with 16-byte stores.
// An SS-mode MMA of this shape streams (128 + 64) rows x 32 B of operands from shared memory:
// 48 clk per instruction (profiles/r01c_microbench.txt), 3,072 clk for K = 1024.
#include "common.cuh"
#include "gemm.cuh"
#include "tc.cuh"

#include <stdlib.h>

namespace curla {

namespace {

constexpr int TM = 128, TN = 64, TK = 64, TST = 4;
constexpr uint32_t kLboA = TM * 16 + 16;      // K-major A: stride between 8-element K chunks
constexpr uint32_t kSboA = TK * 16 + 16;      // MN-major A: stride between 8-row chunks
constexpr uint32_t kLboB = TN * 16 + 16;      // K-major B
constexpr uint32_t kSboB = TK * 16 + 16;      // MN-major B
constexpr uint32_t kABytes = 16 * kSboA;      // 16,640 >= 8 * kLboA = 16,512
constexpr uint32_t kBBytes = 8 * kLboB;       // 8,320 (both layouts)
constexpr uint32_t kStageBytesTc = kABytes + kBBytes;      // 24,960 (multiple of 128)
constexpr uint32_t kHdrTc = 128;              // freed[4] @0, done @32, tmem ptr @40
constexpr uint32_t kSmemTc = kHdrTc + TST * kStageBytesTc;

constexpr uint32_t idesc_tc(bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
           ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// timing experiments: clocks of thread 32 of CTA (0,0,0): [0] wait for the stage's copies, [1] fence + barrier,
// [2] wait for the slot's MMAs, [3] issuing the next copies, [4] K loop total, [5] K steps, [6] kernel total
__device__ long long g_gtc_dbg[8];
constexpr int kTcGemmThreads = 256;          // 8 warps: all copy operands; the epilogue splits the 64 columns between warps 0-3 and 4-7

template <bool A_KMAJOR, bool B_KMAJOR, bool SEG>
__global__ void __launch_bounds__(kTcGemmThreads)
k_gemm_tc(GemmArgs p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_freed = s_base, s_done = s_base + 32, s_tptr = s_base + 40;
    const uint32_t s_stage0 = s_base + kHdrTc;
    if (tid == 0) {
        for (int i = 0; i < TST; ++i) mbar_init(s_freed + 8 * i, 1);
        mbar_init(s_done, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tptr), "r"((uint32_t)TN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int bz = p.batch > 1 ? blockIdx.z : 0, sz = p.batch > 1 ? 0 : blockIdx.z;
    const bf16* __restrict__ Ap = p.A + bz * p.bsA;
    const bf16* __restrict__ Bp = p.B + bz * p.bsB;
    const int kbeg = sz * p.k_per_split;
    int kend = kbeg + p.k_per_split;
    if (kend > p.K) kend = p.K;
    const int nk = (kend - kbeg + TK - 1) / TK;

    // Per-thread copy plan, fixed for the whole K loop: 4 chunks of A and 2 of B per stage.
    // Chunk c = tid + 256 i.  K-major operand: row = c / 8, K chunk = c % 8; MN-major operand:
    // column chunk = c % (cols / 8), K row = c / (cols / 8).  Only the K coordinate moves.
    const bf16* a_src[4]; uint32_t a_dst[4]; int a_k[4]; bool a_ok[4];
    const bf16* b_src[2]; uint32_t b_dst[2]; int b_k[2]; bool b_ok[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = tid + i * kTcGemmThreads;
        if (A_KMAJOR) {
            const int r = c >> 3, kc = c & 7;
            a_ok[i] = m0 + r < p.M; a_k[i] = kc * 8;
            a_src[i] = Ap + (long long)(m0 + r) * p.lda;                 // + (segmented) K index
            a_dst[i] = kc * kLboA + r * 16;
        } else {
            const int mc = c & 15, kk = c >> 4;
            a_ok[i] = m0 + mc * 8 < p.M; a_k[i] = kk;
            a_src[i] = Ap + seg_off(m0 + mc * 8, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 1));   // + K * lda
            a_dst[i] = mc * kSboA + kk * 16;
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = tid + i * kTcGemmThreads;
        if (B_KMAJOR) {
            const int r = c >> 3, kc = c & 7;
            b_ok[i] = n0 + r < p.N; b_k[i] = kc * 8;
            b_src[i] = Bp + (long long)(n0 + r) * p.ldb;
            b_dst[i] = kABytes + kc * kLboB + r * 16;
        } else {
            const int nc = c & 7, kk = c >> 3;
            b_ok[i] = n0 + nc * 8 < p.N; b_k[i] = kk;
            b_src[i] = Bp + seg_off(n0 + nc * 8, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 2));
            b_dst[i] = kABytes + nc * kSboB + kk * 16;
        }
    }
    auto load_stage = [&](int kt, int st) {
        const int k0 = kbeg + kt * TK;
        const uint32_t sS = s_stage0 + st * kStageBytesTc;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + a_k[i];
            const bool ok = a_ok[i] && k < kend;
            const bf16* src = A_KMAJOR ? a_src[i] + seg_off(k, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 1))
                                       : a_src[i] + (long long)k * p.lda;
            cp_async16(sS + a_dst[i], ok ? (const void*)src : (const void*)p.A, ok ? 16 : 0);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int k = k0 + b_k[i];
            const bool ok = b_ok[i] && k < kend;
            const bf16* src = B_KMAJOR ? b_src[i] + seg_off(k, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 2))
                                       : b_src[i] + (long long)k * p.ldb;
            cp_async16(sS + b_dst[i], ok ? (const void*)src : (const void*)p.B, ok ? 16 : 0);
        }
    };

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + 40);
    pdl_grid_sync();              // everything above is independent of earlier kernels

#pragma unroll
    for (int s = 0; s < TST - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    // descriptor templates: K-major (lbo = K-chunk stride, sbo = 128); MN-major (lbo = 128 = eight K rows, sbo = chunk stride)
    const uint64_t a_hi = A_KMAJOR ? make_desc(0, kLboA, 128) : make_desc(0, 128, kSboA);
    const uint64_t b_hi = B_KMAJOR ? make_desc(0, kLboB, 128) : make_desc(0, 128, kSboB);
    constexpr uint32_t kId = idesc_tc(!A_KMAJOR, !B_KMAJOR);

    const bool dbg = tid == 32 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0;
    long long d0 = 0, d1 = 0, d2 = 0, d3 = 0;
    const long long dl0 = clock64();
    for (int kt = 0; kt < nk; ++kt) {
        const long long c0 = dbg ? clock64() : 0;
        cp_async_wait<TST - 2>();
        const long long c1 = dbg ? clock64() : 0;
        fence_proxy_async();          // this thread's cp.async writes -> visible to the tensor core
        __syncthreads();              // ... and everybody else's
        const long long c2 = dbg ? clock64() : 0;
        const int st = kt % TST;
        if (tid == 0) {
            tc_fence_after();
            const uint32_t sA = s_stage0 + st * kStageBytesTc, sB = sA + kABytes;
#pragma unroll
            for (int ks = 0; ks < TK / 16; ++ks) {
                const uint32_t ao = A_KMAJOR ? (uint32_t)(2 * ks) * kLboA : (uint32_t)ks * 256u;
                const uint32_t bo = B_KMAJOR ? (uint32_t)(2 * ks) * kLboB : (uint32_t)ks * 256u;
                const uint64_t ad = a_hi | (uint64_t)(((sA + ao) >> 4) & 0x3FFFu);
                const uint64_t bd = b_hi | (uint64_t)(((sB + bo) >> 4) & 0x3FFFu);
                umma_bf16_rt(tmem_base, ad, bd, kId, (uint32_t)((kt | ks) != 0));
            }
            umma_commit(s_freed + 8 * st);        // slot reusable once these MMAs have read it
        }
        // stage kt + 3 goes into the slot of stage kt - 1: wait until its MMAs are done
        const int nxt = kt + TST - 1;
        long long c3 = c2;
        if (nxt < nk) {
            if (kt >= 1) mbar_wait(s_freed + 8 * ((kt - 1) % TST), (uint32_t)(((kt - 1) / TST) & 1));
            c3 = dbg ? clock64() : 0;
            load_stage(nxt, nxt % TST);
        }
        cp_async_commit();
        if (dbg) { d0 += c1 - c0; d1 += c2 - c1; d2 += c3 - c2; d3 += clock64() - c3; }
    }
    if (dbg) { g_gtc_dbg[0] = d0; g_gtc_dbg[1] = d1; g_gtc_dbg[2] = d2; g_gtc_dbg[3] = d3; g_gtc_dbg[4] = clock64() - dl0; g_gtc_dbg[5] = nk; }
    cp_async_wait<0>();
    if (tid == 0) umma_commit(s_done);
    mbar_wait(s_done, 0);
    tc_fence_after();

    // ---- epilogue: thread = output row m0 + 32 (warp % 4) + lane, 32 columns (warps 0-3: 0..31, warps 4-7: 32..63)
    const int m = m0 + (warp & 3) * 32 + lane;
    const int h = warp >> 2;
    float* Cf = (float*)p.C + (long long)sz * p.split_stride + bz * p.bsC;
    bf16* Cb = (bf16*)p.C + bz * p.bsC;
    const float* __restrict__ biasp = p.bias ? p.bias + bz * p.bsBias : nullptr;
    const bf16* __restrict__ maskp = p.mask ? p.mask + bz * p.bsMask : nullptr;
    const bool f4 = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cf) & 15) == 0);
    {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(h * 32), r);
        if (m < p.M) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + h * 32 + j * 8;
            if (n >= p.n_store) continue;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                v[e] = __uint_as_float(r[j * 8 + e]) * p.alpha;
                if (biasp && n + e < p.n_store) v[e] += biasp[n + e];
                if (p.relu) v[e] = fmaxf(v[e], 0.f);
            }
            const long long nc = seg_off(n, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 4));
            const bool whole = n + 8 <= p.n_store;
            if (maskp) {
                if (whole) {
                    const uint4 mk = *reinterpret_cast<const uint4*>(maskp + (long long)m * p.ldmask + nc);
                    const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 mv = unpack_bf16x2(mw[e]);
                        v[2 * e] = mv.x > 0.f ? v[2 * e] : 0.f;
                        v[2 * e + 1] = mv.y > 0.f ? v[2 * e + 1] : 0.f;
                    }
                } else {
                    for (int e = 0; e < 8 && n + e < p.n_store; ++e)
                        v[e] = __bfloat162float(maskp[(long long)m * p.ldmask + nc + e]) > 0.f ? v[e] : 0.f;
                }
            }
            const long long o = (long long)m * p.ldc + nc;
            if (p.out_bf16) {
                if (whole) {
                    *reinterpret_cast<uint4*>(Cb + o) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                                    pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                } else {
                    for (int e = 0; e < 8 && n + e < p.n_store; ++e) Cb[o + e] = __float2bfloat16(v[e]);
                }
            } else if (whole && f4) {
                *reinterpret_cast<float4*>(Cf + o) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(Cf + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
                for (int e = 0; e < 8 && n + e < p.n_store; ++e) Cf[o + e] = v[e];
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (dbg) g_gtc_dbg[6] = clock64() - dl0;
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TN) : "memory");
    }
}

template <bool AK, bool BK_, bool SG>
int launch_one(const GemmArgs& p, dim3 grid, cudaStream_t stream) {
    auto kern = k_gemm_tc<AK, BK_, SG>;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTc);
        if (e != cudaSuccess) { set_last_error("gemm_tc: cudaFuncSetAttribute(smem=%u): %s", kSmemTc, cudaGetErrorString(e)); return -1; }
        attr = true;
    }
    // short K loops never refill a slot: only the stages they fill are allocated (fc dgrad, K = 64:
    // one 25 KB stage, eight CTAs per SM instead of two)
    const int nk = cdiv(p.k_per_split < p.K ? p.k_per_split : p.K, TK);
    const uint32_t smem = kHdrTc + (uint32_t)(nk < TST ? nk : TST) * kStageBytesTc;
    launch_k(kern, grid, dim3(kTcGemmThreads), smem, stream, p);
    return 1;
}

}  // namespace

int gemm_tc_try_launch(const GemmArgs& p, int layout, int splits, cudaStream_t stream) {
    // Off unless CURLA_GEMM_TC=1 (read per call: the tests compare both kernels).  Fed by 16-byte
    // cp.async this kernel is bound by the LSU path (~21 B/clk per SM of operand copies, see
    // tests/bench_gemm.py / profiles/r02a_gemm_tcgen05_vs_mma_sync.txt), not by the tensor core, and
    // is slower than gemm.cu's bulk-copy ring on every shape of the update; it needs tensor-map
    // TMA loads (128-byte swizzled tiles) before it can replace it.
    const char* env = getenv("CURLA_GEMM_TC");
    if (!(env && env[0] == '1')) return 0;
    // at least one full tile of rows, or a long streaming problem with few rows (fc wgrad: 50 x 67,456 x B)
    if (p.M < TM && !(p.N >= 4096 && p.K >= 128)) return 0;
    // 16-byte epilogue accesses: bf16 rows (and mask rows) must be 16-byte addressable
    const uintptr_t cptr = reinterpret_cast<uintptr_t>(p.C);
    if (p.out_bf16 && !(p.ldc % 8 == 0 && p.bsC % 8 == 0 && (cptr & 15) == 0)) return 0;
    if (p.mask && !(p.ldmask % 8 == 0 && p.bsMask % 8 == 0 && (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0)) return 0;
    if ((p.seg_mask & 4) && p.seg_len % 8 != 0) return 0;
    dim3 grid(cdiv(p.N, TN), cdiv(p.M, TM), p.batch > 1 ? p.batch : splits);
    if (grid.y > 65535 || grid.z > 65535) return 0;
    switch ((layout & 3) | (p.seg_mask ? 4 : 0)) {
        case 3: return launch_one<true, true, false>(p, grid, stream);
        case 1: return launch_one<true, false, false>(p, grid, stream);
        case 2: return launch_one<false, true, false>(p, grid, stream);
        case 0: return launch_one<false, false, false>(p, grid, stream);
        case 7: return launch_one<true, true, true>(p, grid, stream);
        case 5: return launch_one<true, false, true>(p, grid, stream);
        case 6: return launch_one<false, true, true>(p, grid, stream);
        default: return launch_one<false, false, true>(p, grid, stream);
    }
}

}  // namespace curla

extern "C" int curla_gemm_tc_debug_read(long long* out8) {
    cudaError_t e = cudaMemcpyFromSymbol(out8, curla::g_gtc_dbg, sizeof(long long) * 8);
    CURLA_CHECK(e == cudaSuccess, "gemm_tc_debug_read: %s", cudaGetErrorString(e));
    return 0;
}
