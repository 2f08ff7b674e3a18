// K7/K9: one bf16 tensor-core GEMM for every dense layer on the path:
//   encoder fc fwd (split-K) / dgrad / wgrad      encoder.py:98
//   actor / Q-function trunk layers fwd / bwd      curl_sac.py:70-74,129-133
//
// C[M,N] = epilogue( A[M,K] . B[K,N] ) with either operand stored K-major or MN-major
// (so forward, dgrad and wgrad all read the SAME bf16 buffers, no transposes in HBM):
//   A_KMAJOR : A stored [M][lda] (k contiguous)     else stored [K][lda] (m contiguous)
//   B_KMAJOR : B stored [N][ldb] (k contiguous)     else stored [K][ldb] (n contiguous)
// "Row" extents (the slow index of a stored matrix) may be ragged; contiguous extents
// must be multiples of 8 elements (16-byte cp.async chunks, zero padded buffers).
// 64x64x32 CTA tile, 4 warps, 4-stage cp.async pipeline, mma.sync m16n8k16 bf16->fp32.
// Measured alternatives (tests/bench_gemm.py, profiles/r02a_gemm_tcgen05_vs_mma_sync.txt): every shape of
// the update moves ~4-5 TB/s of operands through 16-byte cp.async (16-19 B/clk per SM).  A
// tcgen05 kernel fed the same way (gemm_tc.cu) is 1.3-2x slower, and this kernel with one
// 128-byte cp.async.bulk per operand row instead of cp.async is 1.5-2x slower still; the way
// forward is whole-tile tensor-map TMA loads.
#include "common.cuh"
#include "gemm.cuh"
#include "../../include/curla_b200.h"

namespace curla {

constexpr int BM = 64, BN = 64, BK = 32, STAGES = 4;

__device__ __forceinline__ uint32_t off64(int row, int chunk) {      // 64-byte rows
    return (uint32_t)row * 64u + (uint32_t)((chunk ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ uint32_t off128(int row, int chunk) {     // 128-byte rows
    return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}

template <bool A_KMAJOR, bool B_KMAJOR, bool SEG>
__global__ void __launch_bounds__(128)
k_gemm(GemmArgs p) {
    pdl_grid_sync();
    __shared__ __align__(128) uint8_t smem[STAGES * (BM * BK * 2 + BN * BK * 2)];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int bz = p.batch > 1 ? blockIdx.z : 0, sz = p.batch > 1 ? 0 : blockIdx.z;
    const bf16* __restrict__ Ap = p.A + bz * p.bsA;
    const bf16* __restrict__ Bp = p.B + bz * p.bsB;
    const int kbeg = sz * p.k_per_split;
    int kend = kbeg + p.k_per_split;
    if (kend > p.K) kend = p.K;
    const int nk = (kend - kbeg + BK - 1) / BK;
    const uint32_t sA0 = smem_u32(smem), sB0 = sA0 + STAGES * BM * BK * 2;

    auto load_stage = [&](int kt, int st) {
        const int k0 = kbeg + kt * BK;
        const uint32_t sA = sA0 + st * (BM * BK * 2), sB = sB0 + st * (BN * BK * 2);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = tid + i * 128;
            if (A_KMAJOR) {
                const int r = c >> 2, ch = c & 3;
                const int m = m0 + r, k = k0 + ch * 8;
                const bool ok = (m < p.M) && (k < kend);
                cp_async16(sA + off64(r, ch), ok ? (const void*)(Ap + (long long)m * p.lda + seg_off(k, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 1))) : (const void*)p.A, ok ? 16 : 0);
            } else {
                const int r = c >> 3, ch = c & 7;
                const int k = k0 + r, m = m0 + ch * 8;
                const bool ok = (k < kend) && (m < p.M);
                cp_async16(sA + off128(r, ch), ok ? (const void*)(Ap + (long long)k * p.lda + seg_off(m, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 1))) : (const void*)p.A, ok ? 16 : 0);
            }
            if (B_KMAJOR) {
                const int r = c >> 2, ch = c & 3;
                const int n = n0 + r, k = k0 + ch * 8;
                const bool ok = (n < p.N) && (k < kend);
                cp_async16(sB + off64(r, ch), ok ? (const void*)(Bp + (long long)n * p.ldb + seg_off(k, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 2))) : (const void*)p.B, ok ? 16 : 0);
            } else {
                const int r = c >> 3, ch = c & 7;
                const int k = k0 + r, n = n0 + ch * 8;
                const bool ok = (k < kend) && (n < p.N);
                cp_async16(sB + off128(r, ch), ok ? (const void*)(Bp + (long long)k * p.ldb + seg_off(n, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 2))) : (const void*)p.B, ok ? 16 : 0);
            }
        }
    };

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    // The epilogue's ReLU-mask words are fetched now, behind the operand loads already in
    // flight, so their DRAM latency is not serialised after the MMAs (fc dgrad: K = 64 only).
    const bf16* __restrict__ maskp = p.mask ? p.mask + bz * p.bsMask : nullptr;
    uint32_t mkw[16];          // staged epilogue: chunk i -> words 4i..4i+3; direct epilogue: [(mt*2+h)*4 + nt]
    if (maskp && p.vec_c) {
        // staged epilogue: thread owns 16-byte chunks c = tid + 128*i of the 64 x 64 tile (row c>>3, chunk c&7)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + i * 128, m = m0 + (c >> 3), n = n0 + (c & 7) * 8;
            uint4 t = make_uint4(0u, 0u, 0u, 0u);
            if (m < p.M && n < p.n_store)
                t = __ldg(reinterpret_cast<const uint4*>(
                    maskp + (long long)m * p.ldmask + seg_off(n, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 4))));
            mkw[4 * i] = t.x; mkw[4 * i + 1] = t.y; mkw[4 * i + 2] = t.z; mkw[4 * i + 3] = t.w;
        }
    } else if (maskp) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = m0 + wm * 32 + mt * 16 + h * 8 + (lane >> 2);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int n = n0 + wn * 32 + nt * 8 + (lane & 3) * 2;
                    mkw[(mt * 2 + h) * 4 + nt] = 0u;
                    if (m < p.M && n < p.n_store)
                        mkw[(mt * 2 + h) * 4 + nt] = __ldg(reinterpret_cast<const uint32_t*>(
                            maskp + (long long)m * p.ldmask + seg_off(n, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 4))));
                }
            }
    }

    const int l7 = lane & 7, j1 = (lane >> 3) & 1, j2 = lane >> 4;
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = kt + STAGES - 1;
        if (nxt < nk) load_stage(nxt, nxt % STAGES);
        cp_async_commit();
        const int st = kt % STAGES;
        const uint32_t sA = sA0 + st * (BM * BK * 2), sB = sB0 + st * (BN * BK * 2);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t a[2][4], b[4][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                if (A_KMAJOR)
                    ldmatrix_x4(sA + off64(wm * 32 + mt * 16 + j1 * 8 + l7, ks * 2 + j2),
                                a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
                else
                    ldmatrix_x4_trans(sA + off128(ks * 16 + j2 * 8 + l7, (wm * 32 + mt * 16) / 8 + j1),
                                      a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
            }
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                if (B_KMAJOR)
                    ldmatrix_x4(sB + off64(wn * 32 + np * 16 + j2 * 8 + l7, ks * 2 + j1),
                                b[np * 2][0], b[np * 2][1], b[np * 2 + 1][0], b[np * 2 + 1][1]);
                else
                    ldmatrix_x4_trans(sB + off128(ks * 16 + j1 * 8 + l7, (wn * 32 + np * 16) / 8 + j2),
                                      b[np * 2][0], b[np * 2][1], b[np * 2 + 1][0], b[np * 2 + 1][1]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_bf16(acc[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b[nt][0], b[nt][1]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue
    const int gq = lane >> 2, q = lane & 3;
    float* Cf = (float*)p.C + (long long)sz * p.split_stride + bz * p.bsC;
    bf16* Cb = (bf16*)p.C + bz * p.bsC;
    const float* __restrict__ biasp = p.bias ? p.bias + bz * p.bsBias : nullptr;
    if (p.vec_c) {
        // bf16 output through shared memory: the tile is written back as 16-byte chunks, 128
        // contiguous bytes per row (and the mask is read the same way) instead of 4-byte pieces
        constexpr int CP_ = BN + 8;                       // padded row: conflict-free 16-byte reads
        bf16* sC = reinterpret_cast<bf16*>(smem);
        __syncthreads();                                  // every warp is done with the operand stages
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = wm * 32 + mt * 16 + h * 8 + gq;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int cl = wn * 32 + nt * 8 + q * 2, n = n0 + cl;
                    float v0 = acc[mt][nt][h * 2] * p.alpha, v1 = acc[mt][nt][h * 2 + 1] * p.alpha;
                    if (biasp && n < p.n_store) { v0 += biasp[n]; v1 += (n + 1 < p.n_store) ? biasp[n + 1] : 0.f; }
                    if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                    *reinterpret_cast<uint32_t*>(sC + r * CP_ + cl) = pack_bf16x2(v0, v1);
                }
            }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + i * 128, r = c >> 3, m = m0 + r, n = n0 + (c & 7) * 8;
            if (m >= p.M || n >= p.n_store) continue;
            uint4 v = *reinterpret_cast<const uint4*>(sC + r * CP_ + (c & 7) * 8);
            if (maskp) {
                auto keep = [](uint32_t val, uint32_t mw) {
                    const float2 mv = unpack_bf16x2(mw);
                    return (mv.x > 0.f ? val & 0x0000FFFFu : 0u) | (mv.y > 0.f ? val & 0xFFFF0000u : 0u);
                };
                v.x = keep(v.x, mkw[4 * i]); v.y = keep(v.y, mkw[4 * i + 1]); v.z = keep(v.z, mkw[4 * i + 2]); v.w = keep(v.w, mkw[4 * i + 3]);
            }
            bf16* dst = Cb + (long long)m * p.ldc + seg_off(n, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 4));
            if (n + 8 <= p.n_store) {
                *reinterpret_cast<uint4*>(dst) = v;
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t w = k < 2 ? v.x : (k < 4 ? v.y : (k < 6 ? v.z : v.w));
                    if (n + k < p.n_store) reinterpret_cast<uint16_t*>(dst)[k] = (uint16_t)(w >> ((k & 1) * 16));
                }
            }
        }
        return;
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + wm * 32 + mt * 16 + h * 8 + gq;
            if (m >= p.M) continue;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int n = n0 + wn * 32 + nt * 8 + q * 2;
                if (n >= p.n_store) continue;
                float v0 = acc[mt][nt][h * 2] * p.alpha, v1 = acc[mt][nt][h * 2 + 1] * p.alpha;
                if (biasp) { v0 += biasp[n]; v1 += (n + 1 < p.n_store) ? biasp[n + 1] : 0.f; }
                if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                const long long nc = seg_off(n, p.seg_len, p.seg_stride, p.seg_inv, SEG && (p.seg_mask & 4));
                if (maskp) {
                    const float2 mv = unpack_bf16x2(mkw[(mt * 2 + h) * 4 + nt]);
                    v0 = mv.x > 0.f ? v0 : 0.f;
                    v1 = mv.y > 0.f ? v1 : 0.f;
                }
                const long long o = (long long)m * p.ldc + nc;
                if (p.out_bf16) {
                    *reinterpret_cast<uint32_t*>(Cb + o) = pack_bf16x2(v0, v1);
                } else if (n + 1 < p.n_store) {
                    *reinterpret_cast<float2*>(Cf + o) = make_float2(v0, v1);
                } else {
                    Cf[o] = v0;
                }
            }
        }
}

}  // namespace curla

using namespace curla;

// Generic entry.  layout bits: 1 = A stored K-major ([M][lda]), 2 = B stored K-major ([N][ldb]).
// splits > 1 writes fp32 partials C + z*split_stride (caller reduces); requires out fp32.
extern "C" int curla_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* C,
                               long long ldc, int M, int N, int K, int layout, int n_store,
                               int out_bf16, const float* bias, int relu, const void* mask,
                               long long ldmask, int splits, long long split_stride, float alpha,
                               cudaStream_t stream) {
    return curla_gemm_bf16_seg(A, lda, B, ldb, C, ldc, M, N, K, layout, n_store, out_bf16, bias, relu, mask,
                               ldmask, splits, split_stride, alpha, 0, 0, 0, stream);
}

static int gemm_launch(GemmArgs& p, int layout, int splits, cudaStream_t stream);

// `batch` independent GEMMs of one shape in one launch (element strides between problems; a
// stride of 0 shares that operand): the critic's Q1 || Q2 trunks (curl_sac.py:158-169).
extern "C" int curla_gemm_bf16_batched(const void* A, long long lda, const void* B, long long ldb,
                                       void* C, long long ldc, int M, int N, int K, int layout,
                                       int n_store, int out_bf16, const float* bias, int relu,
                                       const void* mask, long long ldmask, float alpha, int batch,
                                       long long bsA, long long bsB, long long bsC, long long bsBias,
                                       long long bsMask, cudaStream_t stream) {
    CURLA_CHECK(M > 0 && N > 0 && K > 0 && batch >= 1, "gemm: empty problem");
    CURLA_CHECK(lda % 8 == 0 && ldb % 8 == 0 && bsA % 8 == 0 && bsB % 8 == 0, "gemm: lda/ldb/batch strides must be multiples of 8");
    CURLA_CHECK((ldc % 2) == 0 && (bsC % 2) == 0 && (bsMask % 2) == 0, "gemm: ldc / C batch stride must be even");
    GemmArgs p;
    p.A = (const bf16*)A; p.lda = lda; p.B = (const bf16*)B; p.ldb = ldb; p.C = C; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.n_store = n_store > 0 ? n_store : N; p.out_bf16 = out_bf16;
    p.bias = bias; p.relu = relu; p.mask = (const bf16*)mask; p.ldmask = ldmask;
    p.split_stride = 0; p.alpha = alpha;
    p.seg_len = 0; p.seg_stride = 0; p.seg_mask = 0;
    p.batch = batch; p.bsA = bsA; p.bsB = bsB; p.bsC = bsC; p.bsBias = bsBias; p.bsMask = bsMask;
    return gemm_launch(p, layout, 1, stream);
}

// Same GEMM with one operand's contiguous index split into equal segments (see GemmArgs).
extern "C" int curla_gemm_bf16_seg(const void* A, long long lda, const void* B, long long ldb, void* C,
                                   long long ldc, int M, int N, int K, int layout, int n_store,
                                   int out_bf16, const float* bias, int relu, const void* mask,
                                   long long ldmask, int splits, long long split_stride, float alpha,
                                   int seg_len, long long seg_stride, int seg_mask, cudaStream_t stream) {
    CURLA_CHECK(M > 0 && N > 0 && K > 0, "gemm: empty problem");
    CURLA_CHECK(seg_mask == 0 || (seg_len > 0 && seg_len % 8 == 0 && seg_stride % 8 == 0),
                "gemm: segment length/stride must be positive multiples of 8");
    CURLA_CHECK(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8");
    CURLA_CHECK((ldc % 2) == 0, "gemm: ldc must be even");
    CURLA_CHECK(splits >= 1 && (splits == 1 || !out_bf16), "gemm: split-K needs fp32 output");
    GemmArgs p;
    p.A = (const bf16*)A; p.lda = lda; p.B = (const bf16*)B; p.ldb = ldb; p.C = C; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.n_store = n_store > 0 ? n_store : N; p.out_bf16 = out_bf16;
    p.bias = bias; p.relu = relu; p.mask = (const bf16*)mask; p.ldmask = ldmask;
    p.split_stride = split_stride; p.alpha = alpha;
    p.seg_len = seg_len; p.seg_stride = seg_stride; p.seg_mask = seg_mask;
    p.batch = 1; p.bsA = p.bsB = p.bsC = p.bsBias = p.bsMask = 0;
    return gemm_launch(p, layout, splits, stream);
}

static int gemm_launch(GemmArgs& p, int layout, int splits, cudaStream_t stream) {
    // K range of a split: a multiple of 64 (the tcgen05 kernel's K step; also fine for BK = 32),
    // so that the number of partial slices does not depend on which kernel runs
    const int kt = cdiv(p.K, 64);
    p.k_per_split = cdiv(kt, splits) * 64;
    const int zs = cdiv(p.K, p.k_per_split);
    dim3 grid(cdiv(p.N, BN), cdiv(p.M, BM), p.batch > 1 ? p.batch : zs);
    p.seg_inv = p.seg_mask ? 1.0f / (float)p.seg_len : 0.f;
    {
        const int r = gemm_tc_try_launch(p, layout, zs, stream);
        if (r < 0) return -1;
        if (r > 0) return check_launch("gemm_bf16");
    }
    p.vec_c = p.out_bf16 && p.ldc % 8 == 0 && p.bsC % 8 == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 &&
              (!p.mask || (p.ldmask % 8 == 0 && p.bsMask % 8 == 0 && (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0));
    switch ((layout & 3) | (p.seg_mask ? 4 : 0)) {
        case 3: launch_k(k_gemm<true, true, false>, dim3(grid), dim3(128), 0, stream, p); break;
        case 1: launch_k(k_gemm<true, false, false>, dim3(grid), dim3(128), 0, stream, p); break;
        case 2: launch_k(k_gemm<false, true, false>, dim3(grid), dim3(128), 0, stream, p); break;
        case 0: launch_k(k_gemm<false, false, false>, dim3(grid), dim3(128), 0, stream, p); break;
        case 7: launch_k(k_gemm<true, true, true>, dim3(grid), dim3(128), 0, stream, p); break;
        case 5: launch_k(k_gemm<true, false, true>, dim3(grid), dim3(128), 0, stream, p); break;
        case 6: launch_k(k_gemm<false, true, true>, dim3(grid), dim3(128), 0, stream, p); break;
        default: launch_k(k_gemm<false, false, true>, dim3(grid), dim3(128), 0, stream, p); break;
    }
    return check_launch("gemm_bf16");
}

// number of K splits curla_gemm_bf16 will actually launch for (K, splits)
extern "C" int curla_gemm_effective_splits(int K, int splits) {
    const int kt = cdiv(K, 64);
    const int kps = cdiv(kt, splits) * 64;
    return cdiv(K, kps);
}
