// K7/K9 on the 5th-generation tensor cores: the bf16 GEMM of gemm.cu for problems with at least
// 128 rows -- the MLP trunks (curl_sac.py:70-74,129-133: 512 x 1024 x 1024) and the encoder fc
// forward / dgrad / wgrad (encoder.py:98: 512 x 64 x 67,456 and transposes) -- issued with
// tcgen05.mma, accumulators in TMEM, operands brought in by tensor-map TMA loads.  Same argument
// block, same four operand layouts and the same epilogues as gemm.cu (alpha, bias, ReLU, ReLU mask,
// bf16 / fp32, split-K partials, batched, segmented operands), so gemm.cu stays the fallback for
// small or unaligned problems.
//
// Why TMA: fed by 16-byte cp.async, both this kernel and gemm.cu top out near 20 B/clk per SM of
// operand copies (LSU path; tests/bench_gemm.py, profiles/r02a_gemm_tcgen05_vs_mma_sync.txt) --
// the tensor core idles.  One cp.async.bulk.tensor per operand tile and stage does not go
// through the LSU.
//
// One CTA = one 128 x 64 output tile, K in steps of 64, up to four 24 KB stages:
//   warp 0 (one lane)  producer: waits for the stage's `empty` mbarrier, arms `full` with the
//                      stage's byte count and issues the tile loads.  Tiles land in the 128-byte
//                      swizzled canonical layouts (cute/atom/mma_traits_sm100.hpp):
//                        K-major  operand (stored [rows][K]):  box {64 k, rows}  -> [row][128 B], SBO = 1024
//                        MN-major operand (stored [K][cols]):  box {64 cols, 64 k} -> [k][128 B], SBO = 1024,
//                                                              LBO = 8192 between 64-column blocks
//                      Out-of-range rows / K tails are zero-filled by the TMA unit.  Segmented
//                      operands (channel-plane activations) are 3-D maps
//                      {seg_len, segments, rows}: a box never leaves its segment.
//   warp 1 (one lane)  MMA issuer: four M=128 N=64 K=16 MMAs per stage, tcgen05.commit -> `empty`.
//   warps 2..5         epilogue: one output row per thread (TMEM lane), 64 columns, 16-byte stores.
#include "common.cuh"
#include "gemm.cuh"
#include "tc.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <unordered_map>

namespace curla {

namespace {

constexpr int TM = 128, TN = 64, TK = 64, kMaxSt = 8, kDefSt = 4;
constexpr uint32_t kABytes = TM * TK * 2, kBBytes = TN * TK * 2, kStageBytesTc = kABytes + kBBytes;   // 16 K + 8 K
constexpr uint32_t kHdrTc = 1024;             // full[8] @0, empty[8] @64, done @128, tmem ptr @136; keeps the stages 1024-byte aligned
constexpr int kTcGemmThreads = 192;

struct TmaPlan {
    int a_segk;            // A's K index is segmented: map {seg_len, segment, rows}
    int b_segn;            // B's N index is segmented: map {seg_len, segment, K}; one N tile never leaves its segment
    int a_segm;            // A's M index is segmented (MN-major A): map {seg_len, segment, K}; one M tile never leaves its segment
    int c_trans;           // store element (m, n) at C[n * ldc + m] (the problem was launched with its operands swapped)
    int per_seg;           // K steps per segment (a_segk) or N / M tiles per segment (b_segn / a_segm)
    int total_steps;       // a_segk: segments * per_seg
    int nst;               // stages allocated
    int stamps;            // timing experiments: per-CTA %globaltimer stamps (g_gtc_stamps)
    int ksub;              // 64-wide K steps per pipeline stage (1 or 2): one full / empty barrier round trip and one commit per
                           // stage cost the MMA thread ~400 clk against 192 clk of MMAs per 64-wide step (the fc forward with
                           // its A loads removed still took 19 of its 28 us: gpurun r05b), so long K loops take two steps per stage
};

constexpr uint32_t idesc_tc(bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
           ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
// shared-memory matrix descriptor, 128-byte swizzle (layout type 2), version 1
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// timing experiments: clocks of the MMA thread of CTA (0,0,0): [0] waiting for `full`, [1] issuing, [4] K loop, [5] K steps, [6] kernel
__device__ long long g_gtc_dbg[8];
// timing experiments (CURLA_GEMM_STAMPS=1): %globaltimer stamps (ns) of every CTA of the last k_gemm_tc launch --
// [0] kernel entry, [1] after the dependency wait, [2] first stage full (MMA warp), [3] K loop issued, [4] accumulator
// complete (epilogue warp 2), [5] epilogue stores issued, [6] epilogue: tile in shared memory, [7] SM id
__device__ long long g_gtc_stamps[1024][8];
__device__ long long g_gtc_epi_clk[1024][2];        // SM clocks of epilogue phase 1 / phase 2 (warp 2)
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <bool A_KMAJOR, bool B_KMAJOR, bool SEG>
__global__ void __launch_bounds__(kTcGemmThreads)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs p, TmaPlan pl) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_full = s_base, s_empty = s_base + 64, s_done = s_base + 128, s_tptr = s_base + 136;
    const uint32_t s_stage0 = s_base + kHdrTc;
    const long long t_start = clock64();
    const int cta_lin = (int)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
    const bool stamp = pl.stamps && cta_lin < 1024;
    if (stamp && tid == 0) {
        g_gtc_stamps[cta_lin][0] = gtimer();
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_gtc_stamps[cta_lin][7] = smid;
    }
    if (tid == 0) {
        for (int i = 0; i < kMaxSt; ++i) { mbar_init(s_full + 8 * i, 1); mbar_init(s_empty + 8 * i, 1); }
        mbar_init(s_done, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tptr), "r"((uint32_t)TN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + 136);
    if (tid == 0) {               // the tensor maps are kernel parameters: fetch them while waiting for the previous kernel
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    pdl_grid_sync();              // everything above is independent of earlier kernels
    if (stamp && tid == 0) g_gtc_stamps[cta_lin][1] = gtimer();

    // ---- tile and K-step ranges
    int m0 = blockIdx.y * TM, mrows = TM, am_c0 = 0, am_c1 = 0;
    if (SEG && pl.a_segm) {
        const int sg = blockIdx.y / pl.per_seg, j = blockIdx.y - sg * pl.per_seg;
        m0 = sg * p.seg_len + j * TM;
        mrows = p.seg_len - j * TM < TM ? p.seg_len - j * TM : TM;
        am_c0 = j * TM; am_c1 = sg;
    }
    const int bz = p.batch > 1 ? blockIdx.z : 0, sz = p.batch > 1 ? 0 : blockIdx.z;
    int n0, ncols = TN, bn_c0 = 0, bn_c1 = 0;
    if (SEG && pl.b_segn) {
        const int sg = blockIdx.x / pl.per_seg, j = blockIdx.x - sg * pl.per_seg;
        n0 = sg * p.seg_len + j * TN;
        ncols = p.seg_len - j * TN < TN ? p.seg_len - j * TN : TN;
        bn_c0 = j * TN; bn_c1 = sg;
    } else {
        n0 = blockIdx.x * TN;
    }
    int s_beg, s_end;
    if (SEG && pl.a_segk) {
        const int zs = p.batch > 1 ? 1 : (int)gridDim.z;
        const int sps = (pl.total_steps + zs - 1) / zs;
        s_beg = sz * sps;
        s_end = s_beg + sps < pl.total_steps ? s_beg + sps : pl.total_steps;
    } else {
        const int kbeg = sz * p.k_per_split;
        int kend = kbeg + p.k_per_split;
        if (kend > p.K) kend = p.K;
        s_beg = kbeg / TK;
        s_end = (kend + TK - 1) / TK;
    }
    const int nk = s_end > s_beg ? s_end - s_beg : 0;
    const int nst = pl.nst, ksub = pl.ksub;
    const int nstg = (nk + ksub - 1) / ksub;                 // pipeline stages this CTA runs through
    const uint32_t stage_bytes = (uint32_t)ksub * kStageBytesTc;

    if (warp == 0) {
        // ================= producer (the whole warp runs the loop, one elected lane issues: a loop confined to
        // lane 0 makes the compiler wrap every TMA / MMA instruction in an elect loop over the active lanes)
        {
            const int za = p.bsA ? bz : 0, zb = p.bsB ? bz : 0;
            uint32_t stage = 0, phase = 0;
#pragma unroll 1
            for (int kt = 0; kt < nstg; ++kt) {
                mbar_wait(s_empty + 8 * stage, phase ^ 1);
                const uint32_t bar = s_full + 8 * stage;
                const int nsub = nk - kt * ksub < ksub ? nk - kt * ksub : ksub;
                if (elect_one()) {
                mbar_expect_tx(bar, (uint32_t)nsub * kStageBytesTc);
                for (int sub = 0; sub < nsub; ++sub) {
                const int step = s_beg + kt * ksub + sub;
                const uint32_t dA = s_stage0 + stage * stage_bytes + (uint32_t)sub * kStageBytesTc, dB = dA + kABytes;
                int kb = step * TK;                           // B's K coordinate (logical K)
                if (A_KMAJOR) {
                    if (SEG && pl.a_segk) {
                        const int sg = step / pl.per_seg, j = step - sg * pl.per_seg;
                        tma_load_3d(dA, &tmA, j * TK, sg, m0, bar);
                        kb = sg * p.seg_len + j * TK;
                    } else {
                        tma_load_3d(dA, &tmA, step * TK, m0, za, bar);
                    }
                } else if (SEG && pl.a_segm) {
                    tma_load_3d(dA, &tmA, am_c0, am_c1, step * TK, bar);
                    tma_load_3d(dA + kABytes / 2, &tmA, am_c0 + 64, am_c1, step * TK, bar);
                } else {
                    tma_load_3d(dA, &tmA, m0, step * TK, za, bar);
                    tma_load_3d(dA + kABytes / 2, &tmA, m0 + 64, step * TK, za, bar);
                }
                if (B_KMAJOR) tma_load_3d(dB, &tmB, kb, n0, zb, bar);
                else if (SEG && pl.b_segn) tma_load_3d(dB, &tmB, bn_c0, bn_c1, step * TK, bar);
                else tma_load_3d(dB, &tmB, n0, step * TK, zb, bar);
                }
                }
                __syncwarp();
                if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer
        {
            constexpr uint32_t kId = idesc_tc(!A_KMAJOR, !B_KMAJOR);
            const bool dbg = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0;
            long long d0 = 0, d1 = 0;
            const long long dl0 = clock64();
            uint32_t stage = 0, phase = 0;
#pragma unroll 1
            for (int kt = 0; kt < nstg; ++kt) {
                const long long c0 = dbg ? clock64() : 0;
                mbar_wait(s_full + 8 * stage, phase);
                const long long c1 = dbg ? clock64() : 0;
                if (stamp && kt == 0 && lane == 0) g_gtc_stamps[cta_lin][2] = gtimer();
                tc_fence_after();
                const int nsub = nk - kt * ksub < ksub ? nk - kt * ksub : ksub;
                if (elect_one()) {
                    for (int sub = 0; sub < nsub; ++sub) {
                        const uint32_t sA = s_stage0 + stage * stage_bytes + (uint32_t)sub * kStageBytesTc, sB = sA + kABytes;
#pragma unroll
                        for (int ks = 0; ks < TK / 16; ++ks) {
                            // K-major: 16 k = 32 bytes further along the swizzled 128-byte rows; MN-major: 16 k rows = 2048 bytes
                            const uint64_t ad = A_KMAJOR ? desc_sw128(sA + ks * 32, 16, 1024) : desc_sw128(sA + ks * 2048, kABytes / 2, 1024);
                            const uint64_t bd = B_KMAJOR ? desc_sw128(sB + ks * 32, 16, 1024) : desc_sw128(sB + ks * 2048, kBBytes, 1024);
                            umma_bf16_rt(tmem_base, ad, bd, kId, (uint32_t)((kt | sub | ks) != 0));
                        }
                    }
                    umma_commit(s_empty + 8 * stage);        // slot free once these MMAs have read it
                }
                __syncwarp();
                if (dbg) { d0 += c1 - c0; d1 += clock64() - c1; }
                if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(s_done);
            __syncwarp();
            if (stamp && lane == 0) g_gtc_stamps[cta_lin][3] = gtimer();
            if (dbg) { g_gtc_dbg[0] = d0; g_gtc_dbg[1] = d1; g_gtc_dbg[4] = clock64() - dl0; g_gtc_dbg[5] = nk; }
        }
    } else {
        // ================= epilogue: thread = output row m0 + 32 (warp % 4) + lane, 64 columns in two TMEM loads
        const int q = warp & 3;
        const int mr = q * 32 + lane;                          // row inside the tile
        const int m = m0 + mr;
        float* Cf = (float*)p.C + (long long)sz * p.split_stride + bz * p.bsC;
        bf16* Cb = (bf16*)p.C + bz * p.bsC;
        const float* __restrict__ biasp = p.bias ? p.bias + bz * p.bsBias : nullptr;
        const bf16* __restrict__ maskp = p.mask ? p.mask + bz * p.bsMask : nullptr;
        const bool f4 = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cf) & 15) == 0);
        const bool bias_al = (reinterpret_cast<uintptr_t>(biasp) & 15) == 0;
        // (Code size matters here: the epilogue's first execution fetches its instructions cold, behind the operand stream
        // of every SM.  With its row loops fully unrolled -- 50 KB of code per instantiation -- the sixteen shared-memory
        // reads + stores of the fp32 path took 13,000 clk; as plain loops, 2,900.  Per-CTA %globaltimer stamps,
        // tests/manual/fc_fwd_roles.py, gpurun r05e-r05n: 28.7 -> 23.5 us for the fc forward alone, 1.908 -> 1.886 ms per
        // update.  Unrolling by four, or a dry run of the epilogue at kernel start to warm the instruction cache, measured no
        // further gain.)
        mbar_wait(s_done, 0);
        tc_fence_after();
        if (stamp && tid == 64) g_gtc_stamps[cta_lin][4] = gtimer();
        const long long ec0 = clock64();
        if (pl.c_trans) {
            // operands were swapped: this thread's row is a COLUMN of the caller's fp32 matrix, so a
            // warp writes 32 consecutive floats per column (no bias / mask on this path)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                __syncwarp();                                     // lanes re-converge before the warp-collective TMEM load
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32), r);
                if (m >= p.M || mr >= mrows) continue;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int n = n0 + h * 32 + c;
                    if (n < p.n_store) Cf[(long long)n * p.ldc + m] = nk > 0 ? __uint_as_float(r[c]) * p.alpha : 0.f;
                }
            }
        } else {
            // Phase 1: one row per lane (the TMEM layout): alpha, bias, ReLU in fp32 -> this warp's patch of
            // shared memory (the operand stages are free: every MMA has completed), rows 272 bytes apart.
            // Phase 2: read back as 16-byte chunks with 8 (bf16 out) or 16 (fp32 out) lanes per row, so
            // that mask loads and stores are whole 128 / 256-byte row pieces per instruction (one row
            // per lane = 32 scattered 16-byte pieces per instruction: measured 2-3x slower, see the N-loop kernel).
            const uint32_t patch = s_stage0 + (uint32_t)q * (32 * 272);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32), r);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = n0 + h * 32 + j * 4;
                    float v[4], bv[4] = {0.f, 0.f, 0.f, 0.f};
                    if (biasp) {
                        if (bias_al && n + 4 <= p.n_store) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(biasp + n));
                            bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) if (n + e < p.n_store) bv[e] = biasp[n + e];
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        v[e] = (nk > 0 ? __uint_as_float(r[j * 4 + e]) : 0.f) * p.alpha + bv[e];
                        if (p.relu) v[e] = fmaxf(v[e], 0.f);
                    }
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(patch + lane * 272 + (h * 32 + j * 4) * 4), "f"(v[0]), "f"(v[1]),
                                 "f"(v[2]), "f"(v[3]) : "memory");
                }
            }
            __syncwarp();
            const long long ec1 = clock64();
            if (stamp && tid == 64) { g_gtc_stamps[cta_lin][6] = gtimer(); g_gtc_epi_clk[cta_lin][0] = ec1 - ec0; g_gtc_epi_clk[cta_lin][1] = ec1; }
            if (p.out_bf16) {
                const int cr = lane >> 3, cc = lane & 7;
                const int c = cc * 8, n = n0 + c;
                int lim = p.n_store - n;
                if (ncols - c < lim) lim = ncols - c;
                const long long nc = seg_off(n, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 4));
#pragma unroll 1
                for (int t = 0; t < 8; ++t) {
                    const int row = t * 4 + cr, mm = m0 + q * 32 + row;
                    float4 a, b;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(patch + row * 272 + c * 4));
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(patch + row * 272 + c * 4 + 16));
                    if (mm >= p.M || q * 32 + row >= mrows || lim <= 0) continue;
                    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    const long long o = (long long)mm * p.ldc + nc;
                    if (lim >= 8) {
                        if (maskp) {
                            const uint4 mk = *reinterpret_cast<const uint4*>(maskp + (long long)mm * p.ldmask + nc);
                            const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 mv = unpack_bf16x2(mw[e]);
                                v[2 * e] = mv.x > 0.f ? v[2 * e] : 0.f;
                                v[2 * e + 1] = mv.y > 0.f ? v[2 * e + 1] : 0.f;
                            }
                        }
                        *reinterpret_cast<uint4*>(Cb + o) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                                        pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    } else {
                        for (int e = 0; e < lim; ++e) {
                            float x = v[e];
                            if (maskp) x = __bfloat162float(maskp[(long long)mm * p.ldmask + nc + e]) > 0.f ? x : 0.f;
                            Cb[o + e] = __float2bfloat16(x);
                        }
                    }
                }
            } else {
                const int cr = lane >> 4, cc = lane & 15;
                const int c = cc * 4, n = n0 + c;
                int lim = p.n_store - n;
                if (ncols - c < lim) lim = ncols - c;
                const long long nc = seg_off(n, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 4));
#pragma unroll 1
                for (int t = 0; t < 16; ++t) {
                    const int row = t * 2 + cr, mm = m0 + q * 32 + row;
                    float4 a;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(patch + row * 272 + c * 4));
                    if (mm >= p.M || q * 32 + row >= mrows || lim <= 0) continue;
                    float v[4] = {a.x, a.y, a.z, a.w};
                    if (maskp)
                        for (int e = 0; e < 4 && e < lim; ++e)
                            v[e] = __bfloat162float(maskp[(long long)mm * p.ldmask + nc + e]) > 0.f ? v[e] : 0.f;
                    const long long o = (long long)mm * p.ldc + nc;
                    if (lim >= 4 && f4) *reinterpret_cast<float4*>(Cf + o) = make_float4(v[0], v[1], v[2], v[3]);
                    else for (int e = 0; e < 4 && e < lim; ++e) Cf[o + e] = v[e];
                }
            }
        }
    }
    if (stamp && tid == 64) { g_gtc_stamps[cta_lin][5] = gtimer(); g_gtc_epi_clk[cta_lin][1] = clock64() - g_gtc_epi_clk[cta_lin][1]; }
    tc_fence_before();
    __syncthreads();
    if (tid == 32 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) g_gtc_dbg[6] = clock64() - t_start;
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TN) : "memory");
    }
}

// ---------------------------------------------------------------- one K step, many N tiles per CTA
// fc dgrad (encoder.py:98 backward): dact[B][67,456] = (act > 0) * dfc[B][64] . Wfc[64][67,456] -- K = 64 is a
// single step, the work is the epilogue (16 KB of mask in, 16 KB of bf16 out per 128 x 64 tile).  One CTA
// per tile pays its fixed cost 4,216 times; here a CTA keeps its A tile (128 rows of dfc) resident and
// walks over a contiguous range of N tiles: producer warp (B tiles through a 4-deep ring), MMA warp
// (4 MMAs per tile into one of four TMEM accumulators), four epilogue groups of four warps taking
// tiles round robin.  An epilogue warp owns 32 rows x 64 columns: TMEM -> registers (one row per lane)
// -> bf16 -> its private shared-memory patch -> read back as 16-byte chunks with eight lanes per
// row, so that mask loads and stores are 4 rows x 128 contiguous bytes per instruction (one row
// per lane = 32 scattered 16-byte pieces per instruction measured 85-100 us instead of ~50).
constexpr int kNlThreads = 18 * 32;      // producer, MMA, 4 groups x 4 epilogue warps
constexpr int kNlAcc = 4;
constexpr uint32_t kNlPatch = 32 * 144;  // one warp's 32 rows x (128 + 16) bytes
constexpr uint32_t kNlSmem = kHdrTc + kABytes + 4 * kBBytes + 16 * kNlPatch;

template <bool SEG>
__global__ void __launch_bounds__(kNlThreads, 1)
k_gemm_tc_nloop(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs p, int tiles_n) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s_base = smem_u32(smem);
    // header: b_full[4] @0, b_empty[4] @32, tfull[4] @64, tempty[4] @96, a_full @128, tmem ptr @136
    const uint32_t s_bfull = s_base, s_bempty = s_base + 32, s_tfull = s_base + 64, s_tempty = s_base + 96;
    const uint32_t s_afull = s_base + 128, s_tptr = s_base + 136;
    const uint32_t sA = s_base + kHdrTc, sB0 = sA + kABytes, sP0 = sB0 + 4 * kBBytes;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(s_bfull + 8 * i, 1); mbar_init(s_bempty + 8 * i, 1); }
        for (int i = 0; i < kNlAcc; ++i) { mbar_init(s_tfull + 8 * i, 1); mbar_init(s_tempty + 8 * i, 4); }
        mbar_init(s_afull, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tptr), "r"((uint32_t)(kNlAcc * TN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + 136);
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    pdl_grid_sync();

    const int m0 = blockIdx.y * TM;
    // a CTA walks over a CONTIGUOUS range of N tiles: every row's 128-byte pieces are then touched in address order
    const int per = (tiles_n + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t_first = (int)blockIdx.x * per;
    const int my_tiles = tiles_n - t_first < per ? (tiles_n - t_first > 0 ? tiles_n - t_first : 0) : per;

    if (warp == 0) {
        // (whole-warp loops with one elected lane issuing, as in k_gemm_tc)
        if (elect_one()) {
            mbar_expect_tx(s_afull, kABytes);
            tma_load_3d(sA, &tmA, 0, m0, 0, s_afull);
        }
        __syncwarp();
        uint32_t stage = 0, phase = 0;
#pragma unroll 1
        for (int i = 0; i < my_tiles; ++i) {
            mbar_wait(s_bempty + 8 * stage, phase ^ 1);
            if (elect_one()) {
                mbar_expect_tx(s_bfull + 8 * stage, kBBytes);
                tma_load_3d(sB0 + stage * kBBytes, &tmB, (t_first + i) * TN, 0, 0, s_bfull + 8 * stage);
            }
            __syncwarp();
            if (++stage == 4) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t kId = idesc_tc(false, true);          // A K-major, B MN-major
        mbar_wait(s_afull, 0);
        uint32_t stage = 0, phase = 0;
#pragma unroll 1
        for (int i = 0; i < my_tiles; ++i) {
            const uint32_t acc = (uint32_t)(i % kNlAcc), acc_phase = (uint32_t)((i / kNlAcc) & 1);
            mbar_wait(s_bfull + 8 * stage, phase);
            mbar_wait(s_tempty + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t sB = sB0 + stage * kBBytes;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < TK / 16; ++ks)
                    umma_bf16_rt(tmem_base + acc * TN, desc_sw128(sA + ks * 32, 16, 1024), desc_sw128(sB + ks * 2048, kBBytes, 1024), kId,
                                 (uint32_t)(ks != 0));
                umma_commit(s_bempty + 8 * stage);
                umma_commit(s_tfull + 8 * acc);
            }
            __syncwarp();
            if (++stage == 4) { stage = 0; phase ^= 1; }
        }
    } else {
        // ================= epilogue group g (tiles i with i % 4 == g); warp = TMEM lane quarter q = rows 32q .. 32q+31
        const int g = (warp - 2) >> 2, q = warp & 3;
        const uint32_t patch = sP0 + (uint32_t)(warp - 2) * kNlPatch;
        bf16* __restrict__ Cb = (bf16*)p.C;
        const bf16* __restrict__ maskp = p.mask;
        // read-back / global access pattern: pass t (0..7): row 4t + lane / 8 of the warp's 32, 16-byte chunk lane % 8
        const int cr = lane >> 3, cc = lane & 7;
        uint4 mk[8];
        auto load_mask = [&](int i) {
            const int n = (t_first + i) * TN + cc * 8;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                mk[t] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);      // "keep" when there is no mask
                const int m = m0 + q * 32 + t * 4 + cr;
                if (maskp && m < p.M && n + 8 <= p.n_store)
                    mk[t] = *reinterpret_cast<const uint4*>(maskp + (long long)m * p.ldmask +
                                                            seg_off(n, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 4)));
            }
        };
        if (g < my_tiles) load_mask(g);
#pragma unroll 1
        for (int i = g; i < my_tiles; i += kNlAcc) {
            const uint32_t acc = (uint32_t)(i % kNlAcc), acc_phase = (uint32_t)((i / kNlAcc) & 1);
            mbar_wait(s_tfull + 8 * acc, acc_phase);
            __syncwarp();                                         // lanes re-converge before the warp-collective TMEM loads
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * TN + (uint32_t)(h * 32), r);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        w[e] = pack_bf16x2(__uint_as_float(r[j * 8 + 2 * e]) * p.alpha, __uint_as_float(r[j * 8 + 2 * e + 1]) * p.alpha);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(patch + lane * 144 + (h * 4 + j) * 16), "r"(w[0]), "r"(w[1]),
                                 "r"(w[2]), "r"(w[3]) : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_tempty + 8 * acc);
            const int n = (t_first + i) * TN + cc * 8;
            const long long nc = seg_off(n, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 4));
            // (n, hence the ragged-tail case, does not depend on the row: the per-element tail path lives in its own plain
            // loop instead of being replicated in each of the eight unrolled row steps)
            if (n + 8 <= p.n_store) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int row = t * 4 + cr, m = m0 + q * 32 + row;
                    uint4 v;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(patch + row * 144 + cc * 16));
                    if (m >= p.M) continue;
                    auto keep = [](uint32_t val, uint32_t mw) {
                        const float2 mv = unpack_bf16x2(mw);
                        return (mv.x > 0.f ? val & 0x0000FFFFu : 0u) | (mv.y > 0.f ? val & 0xFFFF0000u : 0u);
                    };
                    v.x = keep(v.x, mk[t].x); v.y = keep(v.y, mk[t].y); v.z = keep(v.z, mk[t].z); v.w = keep(v.w, mk[t].w);
                    *reinterpret_cast<uint4*>(Cb + (long long)m * p.ldc + nc) = v;
                }
            } else if (n < p.n_store) {
#pragma unroll 1
                for (int t = 0; t < 8; ++t) {
                    const int row = t * 4 + cr, m = m0 + q * 32 + row;
                    uint4 v;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(patch + row * 144 + cc * 16));
                    if (m >= p.M) continue;
                    const long long o = (long long)m * p.ldc + nc;
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                    for (int e = 0; e < p.n_store - n; ++e) {
                        uint16_t x = (uint16_t)(w[e >> 1] >> ((e & 1) * 16));
                        if (maskp && !(__bfloat162float(maskp[(long long)m * p.ldmask + nc + e]) > 0.f)) x = 0;
                        reinterpret_cast<uint16_t*>(Cb + o)[e] = x;
                    }
                }
            }
            __syncwarp();                                         // the patch is rewritten by the next tile
            if (i + kNlAcc < my_tiles) load_mask(i + kNlAcc);     // lands while the other three groups work
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(kNlAcc * TN)) : "memory");
    }
}

// ---------------------------------------------------------------- tensor maps (host)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
        cudaGetLastError();
    }
    return fn;
}

struct TmKey {
    const void* ptr; unsigned long long d[3], s[2]; unsigned b[3]; unsigned pad;
    bool operator==(const TmKey& o) const { return memcmp(this, &o, sizeof(TmKey)) == 0; }
};
struct TmKeyHash {
    size_t operator()(const TmKey& k) const {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
        size_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(TmKey); ++i) h = (h ^ p[i]) * 1099511628211ull;
        return h;
    }
};
// 3-D bf16 map, 128-byte swizzle, zero fill; cached (the update's operands live at fixed addresses)
const CUtensorMap* get_map(const void* ptr, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                           unsigned long long s1_bytes, unsigned long long s2_bytes, unsigned b0, unsigned b1, unsigned b2) {
    static thread_local std::unordered_map<TmKey, CUtensorMap, TmKeyHash> cache;
    TmKey k;
    memset(&k, 0, sizeof(k));
    k.ptr = ptr; k.d[0] = d0; k.d[1] = d1; k.d[2] = d2; k.s[0] = s1_bytes; k.s[1] = s2_bytes; k.b[0] = b0; k.b[1] = b1; k.b[2] = b2;
    auto it = cache.find(k);
    if (it != cache.end()) return &it->second;
    CUtensorMap tm;
    const cuuint64_t gd[3] = {d0, d1, d2};
    const cuuint64_t gs[2] = {s1_bytes, s2_bytes};
    const cuuint32_t bx[3] = {b0, b1, b2};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gd, gs, bx, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) dims %llu x %llu x %llu strides %llu, %llu box %u x %u x %u",
                       (int)r, d0, d1, d2, s1_bytes, s2_bytes, b0, b1, b2);
        return nullptr;
    }
    // A launch looks up two maps and keeps both pointers: nothing cached is ever freed (unordered_map
    // nodes do not move on rehash).  Once the cache is full -- a process walking over thousands of
    // distinct operand addresses -- further maps are built uncached in a small ring of slots, more than
    // one launch uses.
    if (cache.size() >= 4096) {
        static thread_local CUtensorMap overflow[8];
        static thread_local unsigned next = 0;
        CUtensorMap* slot = &overflow[next++ % 8];
        *slot = tm;
        return slot;
    }
    return &cache.emplace(k, tm).first->second;
}

// cudaFuncSetAttribute is per device: one flag per (kernel instantiation, device)
bool attr_needed(bool (&done)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
}

template <bool AK, bool BK_, bool SG>
int launch_one(const CUtensorMap* ta, const CUtensorMap* tb, const GemmArgs& p, const TmaPlan& pl, dim3 grid, cudaStream_t stream) {
    auto kern = k_gemm_tc<AK, BK_, SG>;
    static bool attr[64] = {};
    if (attr_needed(attr)) {
        const int mx = (int)(kHdrTc + kMaxSt * kStageBytesTc);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (e != cudaSuccess) { set_last_error("gemm_tc: cudaFuncSetAttribute(smem=%d): %s", mx, cudaGetErrorString(e)); return -1; }
    }
    uint32_t smem = kHdrTc + (uint32_t)(pl.nst * pl.ksub) * kStageBytesTc;
    if (smem < kHdrTc + 4 * 32 * 272) smem = kHdrTc + 4 * 32 * 272;        // the epilogue's four 32-row fp32 patches reuse the stages
    launch_k(kern, grid, dim3(kTcGemmThreads), smem, stream, *ta, *tb, p, pl);
    return 1;
}

}  // namespace

static int tc_swapped_wgrad(const GemmArgs& p, cudaStream_t stream);
static int tc_nloop(const GemmArgs& p, cudaStream_t stream);

int gemm_tc_try_launch(const GemmArgs& p, int layout, int splits, cudaStream_t stream) {
    // CURLA_GEMM_TC (read per call: the tests compare both kernels): 0 = never, 2 = whenever the
    // shape is supported, default = where it measured faster than gemm.cu (tests/bench_gemm.py,
    // profiles/r02d_gemm_tma_vs_mma_sync.txt).  One CTA per output tile pays ~3 us of fixed cost
    // (TMEM allocation, descriptor fetch, first load), so short K loops (fc dgrad: one step) and
    // tall-skinny problems made of thousands of small tiles (fc wgrad) stay with gemm.cu until
    // this kernel loops over tiles.
    const char* env = getenv("CURLA_GEMM_TC");
    if (env && env[0] == '0') return 0;
    const bool force = env && env[0] == '2';
    // fc wgrad: dW[feat <= 64][N] = A^T . B with B's N index segmented, fp32.  Launched with the operands
    // swapped (M' = N in tiles of 128, N' = feat), so the big operand streams once through full tiles.
    if (layout == 0 && p.seg_mask == 2 && p.M <= TN && p.N >= 4096 && p.K >= 128 && !p.out_bf16 && !p.bias && !p.mask &&
        !p.relu && p.batch <= 1 && splits == 1 && p.n_store == p.N) {
        const int r = tc_swapped_wgrad(p, stream);
        if (r) return r;
    }
    // fc dgrad: one K step, thousands of N tiles, bf16 out (+ ReLU mask): the N-loop kernel
    if (layout == 1 && p.K <= TK && p.N >= 4096 && p.M >= TM && p.out_bf16 && !p.bias && !p.relu && p.batch <= 1 && splits == 1 &&
        (p.seg_mask == 0 || p.seg_mask == 4)) {
        const int r = tc_nloop(p, stream);
        if (r) return r;
    }
    if (p.M < TM && !(p.N >= 4096 && p.K >= 128)) return 0;      // at least one full tile of rows, or fc wgrad (forced only)
    if (!force) {
        const int steps = cdiv(p.k_per_split < p.K ? p.k_per_split : p.K, TK);
        if (p.M < TM || !(steps >= 12 || (steps >= 8 && p.N >= 128))) return 0;
    }
    const bool ak = layout & 1, bk = layout & 2;
    // supported segment patterns: none; A's K (fc fwd, both K-major); B's N (fc wgrad, both MN-major);
    // C / mask columns (fc dgrad: epilogue only)
    if (p.seg_mask && !((p.seg_mask == 1 && ak && bk) || (p.seg_mask == 2 && !ak && !bk) || p.seg_mask == 4)) return 0;
    if (p.seg_mask && (p.seg_len % 8 || p.seg_stride % 8 || p.batch > 1)) return 0;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(p.A) || !al16(p.B) || p.lda % 8 || p.ldb % 8 || p.bsA % 8 || p.bsB % 8 || p.K % 8) return 0;
    // 16-byte epilogue accesses: bf16 rows (and mask rows) must be 16-byte addressable
    if (p.out_bf16 && !(p.ldc % 8 == 0 && p.bsC % 8 == 0 && al16(p.C))) return 0;
    if (p.mask && !(p.ldmask % 8 == 0 && p.bsMask % 8 == 0 && al16(p.mask))) return 0;
    if (!encode_fn()) return 0;

    const int nb = p.batch > 1 ? p.batch : 1;
    auto r8 = [](long long x) { return (unsigned long long)((x + 7) / 8 * 8); };
    TmaPlan pl;
    memset(&pl, 0, sizeof(pl));
    const CUtensorMap *ta, *tb;
    int nsteps = cdiv(p.k_per_split < p.K ? p.k_per_split : p.K, TK);
    unsigned gx = (unsigned)cdiv(p.N, TN);
    // an operand shared between the batched problems (stride 0) gets a 1-deep batch dimension
    auto bdim = [&](long long bs) { return (unsigned long long)(bs ? nb : 1); };
    auto bstr = [&](long long bs, unsigned long long rows, long long ld) { return (unsigned long long)(bs ? bs : (long long)rows * ld) * 2ull; };
    if (p.seg_mask == 1) {
        const int nseg = cdiv(p.K, p.seg_len);
        pl.a_segk = 1; pl.per_seg = cdiv(p.seg_len, TK); pl.total_steps = nseg * pl.per_seg;
        ta = get_map(p.A, (unsigned long long)p.seg_len, (unsigned long long)nseg, (unsigned long long)p.M,
                     (unsigned long long)p.seg_stride * 2, (unsigned long long)p.lda * 2, TK, 1, TM);
        nsteps = cdiv(pl.total_steps, splits);
    } else if (ak) {
        ta = get_map(p.A, (unsigned long long)p.K, (unsigned long long)p.M, bdim(p.bsA), (unsigned long long)p.lda * 2,
                     bstr(p.bsA, (unsigned long long)p.M, p.lda), TK, TM, 1);
    } else {
        ta = get_map(p.A, r8(p.M), (unsigned long long)p.K, bdim(p.bsA), (unsigned long long)p.lda * 2,
                     bstr(p.bsA, (unsigned long long)p.K, p.lda), 64, TK, 1);
    }
    if (p.seg_mask == 2) {
        const int nseg = cdiv(p.N, p.seg_len);
        pl.b_segn = 1; pl.per_seg = cdiv(p.seg_len, TN);
        gx = (unsigned)(nseg * pl.per_seg);
        tb = get_map(p.B, (unsigned long long)p.seg_len, (unsigned long long)nseg, (unsigned long long)p.K,
                     (unsigned long long)p.seg_stride * 2, (unsigned long long)p.ldb * 2, TN, 1, TK);
    } else if (bk) {
        tb = get_map(p.B, (unsigned long long)p.K, (unsigned long long)p.N, bdim(p.bsB), (unsigned long long)p.ldb * 2,
                     bstr(p.bsB, (unsigned long long)p.N, p.ldb), TK, TN, 1);
    } else {
        tb = get_map(p.B, r8(p.N), (unsigned long long)p.K, bdim(p.bsB), (unsigned long long)p.ldb * 2,
                     bstr(p.bsB, (unsigned long long)p.K, p.ldb), TN, TK, 1);
    }
    if (!ta || !tb) return -1;
    // ring depth: 4 stages (two CTAs per SM); the split-K fc forward streams its operands once from HBM and
    // takes CURLA_FC_STAGES (<= 8: one CTA per SM with twice the bytes in flight per CTA)
    int st_max = kDefSt;
    { const char* e = getenv("CURLA_GEMM_STAMPS"); pl.stamps = (e && e[0] == '1') ? 1 : 0; }
    pl.ksub = 1;
    if (pl.a_segk) {
        // the split-K fc forward: CURLA_FC_KSUB=2 takes two 64-wide steps per stage (four 48 KB stages, one CTA per SM; wants
        // CURLA_FC_SPLITS=37) -- measured no faster than one (gpurun r05c): the MMA thread's per-stage cost is not what paces it
        pl.ksub = 1;
        { const char* e = getenv("CURLA_FC_KSUB"); if (e && (e[0] == '1' || e[0] == '2')) pl.ksub = e[0] - '0'; }
        { const char* e = getenv("CURLA_FC_STAGES"); if (e && e[0] >= '2' && e[0] <= '8') st_max = e[0] - '0'; }
        if (st_max * pl.ksub > kMaxSt) st_max = kMaxSt / pl.ksub;
    }
    const int nstages = cdiv(nsteps, pl.ksub);
    pl.nst = nstages < st_max ? (nstages < 1 ? 1 : nstages) : st_max;
    dim3 grid(gx, cdiv(p.M, TM), p.batch > 1 ? p.batch : splits);
    if (grid.y > 65535 || grid.z > 65535) return 0;
    switch ((layout & 3) | (p.seg_mask ? 4 : 0)) {
        case 3: return launch_one<true, true, false>(ta, tb, p, pl, grid, stream);
        case 1: return launch_one<true, false, false>(ta, tb, p, pl, grid, stream);
        case 2: return launch_one<false, true, false>(ta, tb, p, pl, grid, stream);
        case 0: return launch_one<false, false, false>(ta, tb, p, pl, grid, stream);
        case 7: return launch_one<true, true, true>(ta, tb, p, pl, grid, stream);
        case 5: return launch_one<true, false, true>(ta, tb, p, pl, grid, stream);
        case 6: return launch_one<false, true, true>(ta, tb, p, pl, grid, stream);
        default: return launch_one<false, false, true>(ta, tb, p, pl, grid, stream);
    }
}

static int tc_nloop(const GemmArgs& p, cudaStream_t stream) {
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(p.A) || !al16(p.B) || !al16(p.C) || p.lda % 8 || p.ldb % 8 || p.ldc % 8 || p.K % 8 || !encode_fn()) return 0;
    if (p.mask && (!al16(p.mask) || p.ldmask % 8)) return 0;
    if (p.seg_mask && (p.seg_len % 8 || p.seg_stride % 8)) return 0;
    const CUtensorMap* ta = get_map(p.A, (unsigned long long)p.K, (unsigned long long)p.M, 1ull, (unsigned long long)p.lda * 2,
                                    (unsigned long long)p.M * p.lda * 2, TK, TM, 1);
    const CUtensorMap* tb = get_map(p.B, (unsigned long long)((p.N + 7) / 8 * 8), (unsigned long long)p.K, 1ull,
                                    (unsigned long long)p.ldb * 2, (unsigned long long)p.K * p.ldb * 2, TN, TK, 1);
    if (!ta || !tb) return -1;
    const int tiles_n = cdiv(p.N, TN), mt = cdiv(p.M, TM);
    int strips = cdiv(sm_count(), mt);                     // one CTA per SM
    if (strips > tiles_n) strips = tiles_n;
    if (mt > 65535) return 0;
    static bool attr2[2][64] = {};
    const int sg = p.seg_mask ? 1 : 0;
    if (attr_needed(attr2[sg])) {
        cudaError_t e = sg ? cudaFuncSetAttribute(k_gemm_tc_nloop<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNlSmem)
                           : cudaFuncSetAttribute(k_gemm_tc_nloop<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNlSmem);
        if (e != cudaSuccess) { set_last_error("gemm_tc: cudaFuncSetAttribute(smem=%u): %s", kNlSmem, cudaGetErrorString(e)); return -1; }
    }
    if (sg) launch_k(k_gemm_tc_nloop<true>, dim3(strips, mt), dim3(kNlThreads), kNlSmem, stream, *ta, *tb, p, tiles_n);
    else launch_k(k_gemm_tc_nloop<false>, dim3(strips, mt), dim3(kNlThreads), kNlSmem, stream, *ta, *tb, p, tiles_n);
    return 1;
}

static int tc_swapped_wgrad(const GemmArgs& p, cudaStream_t stream) {
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(p.A) || !al16(p.B) || p.lda % 8 || p.ldb % 8 || p.K % 8 || p.seg_len % 8 || p.seg_stride % 8 || !encode_fn()) return 0;
    GemmArgs q = p;
    q.A = p.B; q.lda = p.ldb; q.B = p.A; q.ldb = p.lda;
    q.M = p.N; q.N = p.M; q.n_store = p.M;                  // rows = the caller's columns, columns = the caller's rows
    q.k_per_split = (p.K + TK - 1) / TK * TK;
    TmaPlan pl;
    memset(&pl, 0, sizeof(pl));
    const int nseg = cdiv(p.N, p.seg_len);
    pl.a_segm = 1; pl.c_trans = 1; pl.per_seg = cdiv(p.seg_len, TM);
    const CUtensorMap* ta = get_map(q.A, (unsigned long long)p.seg_len, (unsigned long long)nseg, (unsigned long long)p.K,
                                    (unsigned long long)p.seg_stride * 2, (unsigned long long)q.lda * 2, 64, 1, TK);
    const CUtensorMap* tb = get_map(q.B, (unsigned long long)((p.M + 7) / 8 * 8), (unsigned long long)p.K, 1ull,
                                    (unsigned long long)q.ldb * 2, (unsigned long long)p.K * q.ldb * 2, TN, TK, 1);
    if (!ta || !tb) return -1;
    const int nsteps = cdiv(p.K, TK);
    pl.ksub = 1;
    pl.nst = nsteps < kDefSt ? nsteps : kDefSt;
    dim3 grid(1, nseg * pl.per_seg, 1);
    if (grid.y > 65535) return 0;
    return launch_one<false, false, true>(ta, tb, q, pl, grid, stream);
}

}  // namespace curla

extern "C" int curla_gemm_tc_stamps_read(long long* out, int nctas) {
    CURLA_CHECK(nctas >= 1 && nctas <= 1024, "gemm_tc_stamps_read: 1..1024 CTAs");
    cudaError_t e = cudaMemcpyFromSymbol(out, curla::g_gtc_stamps, sizeof(long long) * 8 * nctas);
    CURLA_CHECK(e == cudaSuccess, "gemm_tc_stamps_read: %s", cudaGetErrorString(e));
    // (experiments: the epilogue's SM clock counts replace slots [2] / [3] when CURLA_GEMM_STAMPS_CLK is set)
    if (getenv("CURLA_GEMM_STAMPS_CLK")) {
        static long long tmp[1024][2];
        e = cudaMemcpyFromSymbol(tmp, curla::g_gtc_epi_clk, sizeof(tmp));
        CURLA_CHECK(e == cudaSuccess, "gemm_tc_stamps_read: %s", cudaGetErrorString(e));
        for (int i = 0; i < nctas; ++i) { out[i * 8 + 2] = tmp[i][0]; out[i * 8 + 3] = tmp[i][1]; }
    }
    return 0;
}

extern "C" int curla_gemm_tc_debug_read(long long* out8) {
    cudaError_t e = cudaMemcpyFromSymbol(out8, curla::g_gtc_dbg, sizeof(long long) * 8);
    CURLA_CHECK(e == cudaSuccess, "gemm_tc_debug_read: %s", cudaGetErrorString(e));
    return 0;
}
