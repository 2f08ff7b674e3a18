// K12: CURL bilinear logits + softmax cross-entropy, forward and backward.
//
// Replaces CURL.compute_logits (curl_sac.py:211-222) and nn.CrossEntropyLoss with
// labels = arange(B) (curl_sac.py:411-413) plus their autograd.
//   U      = z_pos . W^T                 (= (W z_pos^T)^T)            [Bg, feat]
//   logits = z_a . U^T  - rowmax                                       [B, Bg]
//   loss   = mean_i( logsumexp_j(logits_i) - logits_i[label_i] )
//   dlogit = (softmax - onehot) * grad_scale
//   dz_a   = dlogit . U ;  T = dlogit^T . z_a ;  dW = T^T . z_pos
// Data parallel: z_pos holds the ALL-GATHERED keys of every rank (Bg rows) and the local
// rows' labels are label0 + i, so each shard sees the reference's full-batch negatives
// and its exact per-row max.  fp32 throughout (CUDA cores; the contraction is 0.03% of
// the update's FLOPs at the default batch).
#include "common.cuh"

#include <stdlib.h>

namespace curla {

// C[m][n] = sum_k A(m,k) * B(k,n), arbitrary element strides.  32x32 tile, 16x16 threads.
__global__ void __launch_bounds__(256)
k_sgemm_strided(const float* __restrict__ A, long long sam, long long sak,
                const float* __restrict__ Bm, long long sbk, long long sbn,
                float* __restrict__ C, long long ldc, float* __restrict__ Ct, long long ldct, int M, int N, int K) {
    pdl_grid_sync();
    __shared__ float sA[32][33], sB[32][33];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int i = threadIdx.x; i < 1024; i += 256) {
            // choose the faster-varying index to follow the unit stride where possible
            int r, c;
            if (sak == 1) { c = i & 31; r = i >> 5; } else { r = i & 31; c = i >> 5; }   // A tile [m][k]
            const int m = m0 + r, k = k0 + c;
            sA[r][c] = (m < M && k < K) ? A[m * sam + k * sak] : 0.f;
            int rk, cn;
            if (sbn == 1) { cn = i & 31; rk = i >> 5; } else { rk = i & 31; cn = i >> 5; }   // B tile [k][n]
            const int kk = k0 + rk, n = n0 + cn;
            sB[rk][cn] = (kk < K && n < N) ? Bm[kk * sbk + n * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float a0 = sA[ty][k], a1 = sA[ty + 16][k];
            const float b0 = sB[k][tx], b1 = sB[k][tx + 16];
            acc[0][0] += a0 * b0; acc[0][1] += a0 * b1;
            acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = m0 + ty + i * 16, n = n0 + tx + j * 16;
            if (m < M && n < N) {
                C[m * ldc + n] = acc[i][j];
                if (Ct) Ct[n * ldct + m] = acc[i][j];
            }
        }
}

// Fused rows kernel: one CTA owns R local rows i.  Their logits row block [R][Bg] lives in
// shared memory from the contraction to the gradient -- nothing of size B x Bg ever touches
// HBM:   logits = z_a . U^T  ->  softmax / per-row loss  ->  dl = (p - onehot) * grad_scale
//        dz_a[i] = dl[i] . U ,  V[i] = dl[i] . z_pos            (both [R][64])
// U and z_pos ([Bg][64] fp32, <= 1 MB each at Bg = 4096) are streamed from L2 by every CTA.
// dW = z_a^T . V is finished by k_curl_dw (the reference's dW = (dl^T z_a)^T z_pos
// re-associated so that no [Bg][feat] intermediate and no strided pass over dl is needed).
template <int R>
__global__ void __launch_bounds__(256)
k_curl_rows(const float* __restrict__ z_a, const float* __restrict__ U, const float* __restrict__ Ut,
            const float* __restrict__ z_pos,
            int B, int Bg, int label0, float grad_scale, float* __restrict__ row_loss,
            float* __restrict__ dz_a, float* __restrict__ V, float* __restrict__ logits_copy) {
    pdl_grid_sync();
    extern __shared__ __align__(16) float s_dyn[];
    const int BgP = (Bg + 3) & ~3;                // row stride of the logits block (keeps s_za 16-byte aligned)
    float* s_log = s_dyn;                         // [R][BgP]
    float* s_za = s_dyn + (size_t)R * BgP;        // [R][64]
    float* s_red = s_za + R * 64;                 // [4][2][R][64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.x * R;
    for (int t = tid; t < R * 64; t += 256) {
        const int r = t >> 6, i = i0 + r;
        s_za[t] = i < B ? z_a[(long long)i * 64 + (t & 63)] : 0.f;
    }
    __syncthreads();
    // ---- logits[r][j] = z_a[i0+r] . U[j].  Ut[k][j] makes consecutive lanes read consecutive j;
    // a thread owns 4 consecutive j (one 16-byte load per k) and one half of the k range, so each
    // shared-memory broadcast of z_a feeds four FMAs.
    {
        const int kh = tid >> 7;                                // k in [32*kh, 32*kh + 32)
        const bool vec = (Bg & 3) == 0;
        for (int pass = 0; pass < 2; ++pass) {
            // pass 0: half 0 stores its partial sums; pass 1: half 1 adds its own (fixed order => deterministic)
            if (kh == pass)
                for (int j4 = (tid & 127) * 4; j4 < Bg; j4 += 512) {
                    float acc[R][4];
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
#pragma unroll 8
                    for (int k = kh * 32; k < kh * 32 + 32; ++k) {
                        float4 u;
                        const float* up = Ut + (long long)k * Bg + j4;
                        if (vec) u = *reinterpret_cast<const float4*>(up);
                        else { u.x = up[0]; u.y = j4 + 1 < Bg ? up[1] : 0.f; u.z = j4 + 2 < Bg ? up[2] : 0.f; u.w = j4 + 3 < Bg ? up[3] : 0.f; }
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const float a = s_za[r * 64 + k];
                            acc[r][0] += a * u.x; acc[r][1] += a * u.y; acc[r][2] += a * u.z; acc[r][3] += a * u.w;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (j4 + e < Bg) {
                                float* d = s_log + (size_t)r * BgP + j4 + e;
                                *d = pass ? *d + acc[r][e] : acc[r][e];
                            }
                }
            __syncthreads();
        }
    }
    __syncthreads();
    if (logits_copy)
        for (int t = tid; t < R * Bg; t += 256) {
            const int r = t / Bg, j = t - r * Bg;
            if (i0 + r < B) logits_copy[(long long)(i0 + r) * Bg + j] = s_log[(size_t)r * BgP + j];
        }
    // ---- softmax statistics: one warp per row (rows r = warp, warp + 8, ...)
    for (int r = warp; r < R; r += 8) {
        float* row = s_log + (size_t)r * BgP;
        const int i = i0 + r;
        float mx = -INFINITY;
        for (int j = lane; j < Bg; j += 32) mx = fmaxf(mx, row[j]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int j = lane; j < Bg; j += 32) se += expf(row[j] - mx);
        se = warp_sum(se);
        const float inv = 1.f / se;
        const int lab = label0 + i;
        if (lane == 0 && i < B) row_loss[i] = logf(se) - (row[lab] - mx);
        __syncwarp();
        const float gsc = i < B ? grad_scale : 0.f;
        for (int j = lane; j < Bg; j += 32) row[j] = (expf(row[j] - mx) * inv - (j == lab ? 1.f : 0.f)) * gsc;
    }
    __syncthreads();
    // ---- dz_a[r][a] = sum_j dl[r][j] U[j][a] ; V[r][a] = sum_j dl[r][j] z_pos[j][a]
    const int a = tid & 63, q = tid >> 6;
    float dz[R], vv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { dz[r] = 0.f; vv[r] = 0.f; }
#pragma unroll 8
    for (int j = q; j < Bg; j += 4) {
        const float u = U[(long long)j * 64 + a], zp = z_pos[(long long)j * 64 + a];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float d = s_log[(size_t)r * BgP + j];
            dz[r] += d * u;
            vv[r] += d * zp;
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s_red[((q * 2 + 0) * R + r) * 64 + a] = dz[r];
        s_red[((q * 2 + 1) * R + r) * 64 + a] = vv[r];
    }
    __syncthreads();
    for (int t = tid; t < 2 * R * 64; t += 256) {
        const int w = t / (R * 64), rem = t - w * (R * 64);      // rem = r*64 + a
        const float s = s_red[((0 * 2 + w) * R) * 64 + rem] + s_red[((1 * 2 + w) * R) * 64 + rem] +
                        s_red[((2 * 2 + w) * R) * 64 + rem] + s_red[((3 * 2 + w) * R) * 64 + rem];
        const int i = i0 + (rem >> 6);
        if (i < B) (w ? V : dz_a)[(long long)i * 64 + (rem & 63)] = s;
    }
}

// dW[a][b] = sum_i z_a[i][a] * V[i][b] (one CTA per a; 64 columns x 4 row groups, fixed-order
// tree => deterministic); CTA 0 also writes the mean of the per-row losses -- read from row_loss, or (pstat != nullptr:
// the tensor-core path with a single column block, where no reduce kernel runs) formed here from the per-tile softmax
// statistics:  loss_i = log(sum_t s_t exp(m_t - M)) - (l_i,label - M).
__global__ void __launch_bounds__(256)
k_curl_dw(const float* __restrict__ z_a, const float* __restrict__ V, const float* __restrict__ row_loss,
          const float* __restrict__ pstat, const float* __restrict__ lab_logit, int Bp, int NT,
          int B, int feat, float* __restrict__ dW, float* __restrict__ loss_out) {
    pdl_grid_sync();
    __shared__ float s_red[4][64];
    __shared__ float s_l[8];
    const int a = blockIdx.x, b = threadIdx.x & 63, q = threadIdx.x >> 6;
    float s = 0.f;
#pragma unroll 8
    for (int i = q; i < B; i += 4) s += z_a[(long long)i * 64 + a] * V[(long long)i * 64 + b];
    s_red[q][b] = s;
    __syncthreads();
    if (q == 0 && b < feat) dW[a * feat + b] = s_red[0][b] + s_red[1][b] + s_red[2][b] + s_red[3][b];
    if (blockIdx.x == 0) {
        float l = 0.f;
        for (int i = threadIdx.x; i < B; i += 256) {
            if (pstat) {
                float M = -INFINITY, S = 0.f;
                for (int st = 0; st < NT; ++st) M = fmaxf(M, pstat[((long long)st * Bp + i) * 2]);
                for (int st = 0; st < NT; ++st) S += pstat[((long long)st * Bp + i) * 2 + 1] * expf(pstat[((long long)st * Bp + i) * 2] - M);
                l += logf(S) - (lab_logit[i] - M);
            } else {
                l += row_loss[i];
            }
        }
        l = warp_sum(l);
        if ((threadIdx.x & 31) == 0) s_l[threadIdx.x >> 5] = l;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < 8; ++w) t += s_l[w];
            *loss_out = t / B;
        }
    }
}

// ------------------------------------------------------------------ tensor-core path
// The B x Bg bilinear contraction and its two gradient contractions on the tensor cores
// (mma.sync m16n8k16, bf16 operands, fp32 accumulate).  The reference computes them in fp32
// (curl_sac.py:211-222) and logits of magnitude ~30 sit in an exp(), so every fp32 operand x is split
// into THREE bf16 pieces  hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid)  (24 mantissa bits: the fp32 value
// exactly, up to the last rounding) and a product is  hi.hi + hi.mid + mid.hi + mid.mid + hi.lo + lo.hi  (six MMAs; the
// dropped mid.lo / lo.mid / lo.lo terms are 2^-24 relative): fp32-equivalent results, like the fp32 FFMA kernel
// (CURLA_CURL_TC=0) it replaces.
//
// Work unit = one warp: 16 local rows x 64 key columns.  A CTA is 8 warps = the same 16 rows x 512
// columns; grid (B/16 row blocks, Bg/512 column blocks), so a global batch of 4096 keys spreads over
// 8x more CTAs instead of one CTA streaming all of U per row block.  Softmax needs whole rows, hence
// two passes over the same logits (recomputed bit-identically, they are 0.03 % of the update's FLOPs):
//   k_curl_tc<0>: logits tile -> per (64-column tile, row) partial (max, sum exp) -> pstat, label logit
//   k_curl_tc<1>: logits tile again, global (max, sum) from pstat, dl = (softmax - onehot) * grad_scale
//                 kept in registers -- the m16n8 accumulator fragment of two adjacent column tiles IS the
//                 m16k16 A fragment of the next MMA -- then dz_a += dl . U, V += dl . z_pos; the 8 warps
//                 are summed in shared memory in a fixed order and written as per-column-block partials
//   k_curl_reduce: sums the column-block partials (fixed order: deterministic), per-row loss
// Operands are prepared once per update by k_curl_prep_u (U = z_pos . W^T included) as bf16 hi / mid / lo arrays in the
// two orientations the B fragments want: Uk[j][k] (logits) and UT[a][j], ZT[a][j] (gradients).
__device__ __forceinline__ void split_bf16(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    const float2 hf = __bfloat1622float2(h);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = pack_bf16x2(x - hf.x, y - hf.y);
}
__device__ __forceinline__ void split3_bf16(float x, float y, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    const float2 hf = __bfloat1622float2(h);
    const float rx = x - hf.x, ry = y - hf.y;                 // exact
    const __nv_bfloat162 m = __floats2bfloat162_rn(rx, ry);
    const float2 mf = __bfloat1622float2(m);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    mid = *reinterpret_cast<const uint32_t*>(&m);
    lo = pack_bf16x2(rx - mf.x, ry - mf.y);
}
__device__ __forceinline__ void split3_store(float x, bf16* hi, bf16* mid, bf16* lo, long long o) {
    const bf16 h = __float2bfloat16_rn(x);
    const float r = x - __bfloat162float(h);
    const bf16 m = __float2bfloat16_rn(r);
    hi[o] = h; mid[o] = m; lo[o] = __float2bfloat16_rn(r - __bfloat162float(m));
}
// acc += a . b for 3-piece operands (a: A fragments [piece][4], b: B fragment pairs [piece][2])
__device__ __forceinline__ void mma6(float* acc, const uint32_t (&ah)[4], const uint32_t (&am)[4], const uint32_t (&al)[4],
                                     uint32_t bh0, uint32_t bh1, uint32_t bm0, uint32_t bm1, uint32_t bl0, uint32_t bl1) {
    mma_bf16(acc, al[0], al[1], al[2], al[3], bh0, bh1);      // smallest terms first
    mma_bf16(acc, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
    mma_bf16(acc, am[0], am[1], am[2], am[3], bm0, bm1);
    mma_bf16(acc, am[0], am[1], am[2], am[3], bh0, bh1);
    mma_bf16(acc, ah[0], ah[1], ah[2], ah[3], bm0, bm1);
    mma_bf16(acc, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
}
__device__ __forceinline__ uint32_t ldg_u32(const bf16* p) { return __ldg(reinterpret_cast<const unsigned int*>(p)); }

// Uk_*: bf16 [BgP][64]; UT_*, ZT_*: bf16 [64][BgP]; rows j >= Bg are zero.
// U = z_pos . W^T and the nine bf16 operand arrays in ONE launch (was: a strided fp32 GEMM + k_curl_prep, 28 us of
// two latency-bound launches for 2 MFLOP).  A CTA owns 16 keys: W (zero padded to 64 x 64) and the 16 key rows sit in
// shared memory, a thread computes 4 of the 16 x 64 outputs (k = 0..63 in order, one accumulator: the same sums as
// k_sgemm_strided), then the CTA writes its rows of Uk (4-byte pairs) and its 16 columns of UT / ZT.
__global__ void __launch_bounds__(256)
k_curl_prep_u(const float* __restrict__ z_pos, const float* __restrict__ W, int Bg, int BgP, int feat,
              bf16* __restrict__ Uk_hi, bf16* __restrict__ Uk_mid, bf16* __restrict__ Uk_lo,
              bf16* __restrict__ UT_hi, bf16* __restrict__ UT_mid, bf16* __restrict__ UT_lo,
              bf16* __restrict__ ZT_hi, bf16* __restrict__ ZT_mid, bf16* __restrict__ ZT_lo) {
    pdl_grid_sync();
    __shared__ float sW[64][65];
    __shared__ float sz[16][65], su[16][65];
    const int tid = threadIdx.x, j0 = blockIdx.x * 16;
    for (int t = tid; t < 4096; t += 256) {
        const int a = t >> 6, b = t & 63;
        sW[a][b] = (a < feat && b < feat) ? W[a * feat + b] : 0.f;
    }
    for (int t = tid; t < 1024; t += 256) {
        const int jj = t >> 6, k = t & 63, j = j0 + jj;
        sz[jj][k] = j < Bg ? z_pos[(long long)j * 64 + k] : 0.f;
    }
    __syncthreads();
    {
        const int jj = tid >> 4, a4 = (tid & 15) * 4;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int b = 0; b < 64; ++b) {
            const float z = sz[jj][b];
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] += z * sW[a4 + e][b];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) su[jj][a4 + e] = acc[e];
    }
    __syncthreads();
    for (int t = tid; t < 512; t += 256) {
        {   // Uk[j][k], k pairs
            const int jj = t >> 5, k = (t & 31) * 2;
            uint32_t h, m, l;
            split3_bf16(su[jj][k], su[jj][k + 1], h, m, l);
            const long long o = (long long)(j0 + jj) * 64 + k;
            *reinterpret_cast<uint32_t*>(Uk_hi + o) = h;
            *reinterpret_cast<uint32_t*>(Uk_mid + o) = m;
            *reinterpret_cast<uint32_t*>(Uk_lo + o) = l;
        }
        {   // UT[a][j], ZT[a][j], j pairs
            const int a = t >> 3, jj = (t & 7) * 2;
            const long long o = (long long)a * BgP + j0 + jj;
            uint32_t h, m, l;
            split3_bf16(su[jj][a], su[jj + 1][a], h, m, l);
            *reinterpret_cast<uint32_t*>(UT_hi + o) = h;
            *reinterpret_cast<uint32_t*>(UT_mid + o) = m;
            *reinterpret_cast<uint32_t*>(UT_lo + o) = l;
            split3_bf16(sz[jj][a], sz[jj + 1][a], h, m, l);
            *reinterpret_cast<uint32_t*>(ZT_hi + o) = h;
            *reinterpret_cast<uint32_t*>(ZT_mid + o) = m;
            *reinterpret_cast<uint32_t*>(ZT_lo + o) = l;
        }
    }
}

struct CurlTc {
    const float* z_a;                          // [B][64] fp32
    const bf16 *Uk_hi, *Uk_mid, *Uk_lo, *UT_hi, *UT_mid, *UT_lo, *ZT_hi, *ZT_mid, *ZT_lo;
    float* pstat;                              // [NT][Bp][2]  (max, sum exp) per 64-column tile
    float* lab_logit;                          // [Bp]
    float* pdz; float* pV;                     // [NCB][Bp][64] column-block partials (NCB > 1), else dz_a / V directly
    float* logits_copy;                        // optional [B][Bg]
    int B, Bg, BgP, Bp, NT, label0;
    float grad_scale;
};

// NTW = 8-column tiles per warp (8: 64 key columns, 4: 32 -- twice the warps for the same work: small global batches)
template <int PASS, int NTW>
__global__ void __launch_bounds__(256)
k_curl_tc(const CurlTc p) {
    constexpr int CW = NTW * 8;                                // key columns per warp
    pdl_grid_sync();
    extern __shared__ __align__(16) float s_red[];            // PASS 1: [8 warps][2][16][64]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * 16;
    const int jw = (blockIdx.y * 8 + warp) * CW;              // this warp's key columns
    const bool active = jw < p.BgP;
    const int rowA = r0 + g, rowB = r0 + g + 8;
    float lg[NTW][4];
    if (active) {
        // ---- A = z_a rows (hi / lo), all of K = 64 in registers
        uint32_t a_hi[4][4], a_mid[4][4], a_lo[4][4];
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int row = (h & 1) ? rowB : rowA, k = kt * 16 + 2 * t + (h >> 1) * 8;
                float2 v = make_float2(0.f, 0.f);
                if (row < p.B) v = *reinterpret_cast<const float2*>(p.z_a + (long long)row * 64 + k);
                split3_bf16(v.x, v.y, a_hi[kt][h], a_mid[kt][h], a_lo[kt][h]);
            }
        }
        // ---- logits[16][CW] = z_a . U^T : NTW column tiles of 8
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt) {
            lg[nt][0] = lg[nt][1] = lg[nt][2] = lg[nt][3] = 0.f;
            const long long rowoff = (long long)(jw + nt * 8 + g) * 64 + 2 * t;
#pragma unroll
            for (int kt = 0; kt < 4; ++kt) {
                const uint32_t bh0 = ldg_u32(p.Uk_hi + rowoff + kt * 16), bh1 = ldg_u32(p.Uk_hi + rowoff + kt * 16 + 8);
                const uint32_t bm0 = ldg_u32(p.Uk_mid + rowoff + kt * 16), bm1 = ldg_u32(p.Uk_mid + rowoff + kt * 16 + 8);
                const uint32_t bl0 = ldg_u32(p.Uk_lo + rowoff + kt * 16), bl1 = ldg_u32(p.Uk_lo + rowoff + kt * 16 + 8);
                mma6(lg[nt], a_hi[kt], a_mid[kt], a_lo[kt], bh0, bh1, bm0, bm1, bl0, bl1);
            }
        }
    }
    const int labA = p.label0 + rowA, labB = p.label0 + rowB;
    if (PASS == 0) {
        if (!active) return;
        float mA = -INFINITY, mB = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = jw + nt * 8 + 2 * t + e;
                if (j < p.Bg) {
                    mA = fmaxf(mA, lg[nt][e]); mB = fmaxf(mB, lg[nt][2 + e]);
                    if (j == labA && rowA < p.B) p.lab_logit[rowA] = lg[nt][e];
                    if (j == labB && rowB < p.B) p.lab_logit[rowB] = lg[nt][2 + e];
                    if (p.logits_copy) {
                        if (rowA < p.B) p.logits_copy[(long long)rowA * p.Bg + j] = lg[nt][e];
                        if (rowB < p.B) p.logits_copy[(long long)rowB * p.Bg + j] = lg[nt][2 + e];
                    }
                }
            }
        mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1)); mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
        mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1)); mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
        float sA = 0.f, sB = 0.f;
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = jw + nt * 8 + 2 * t + e;
                if (j < p.Bg) { sA += expf(lg[nt][e] - mA); sB += expf(lg[nt][2 + e] - mB); }
            }
        sA += __shfl_xor_sync(0xffffffffu, sA, 1); sA += __shfl_xor_sync(0xffffffffu, sA, 2);
        sB += __shfl_xor_sync(0xffffffffu, sB, 1); sB += __shfl_xor_sync(0xffffffffu, sB, 2);
        if (t == 0) {
            float* ps = p.pstat + ((long long)(jw / CW) * p.Bp) * 2;
            ps[rowA * 2] = mA; ps[rowA * 2 + 1] = sA;
            ps[rowB * 2] = mB; ps[rowB * 2 + 1] = sB;
        }
    } else {
        float dz[8][4], vv[8][4];
#pragma unroll
        for (int na = 0; na < 8; ++na)
#pragma unroll
            for (int e = 0; e < 4; ++e) { dz[na][e] = 0.f; vv[na][e] = 0.f; }
        if (active) {
            // ---- global row statistics from the per-tile partials (fixed order)
            float MA = -INFINITY, MB = -INFINITY;
            for (int st = 0; st < p.NT; ++st) {
                MA = fmaxf(MA, p.pstat[((long long)st * p.Bp + rowA) * 2]);
                MB = fmaxf(MB, p.pstat[((long long)st * p.Bp + rowB) * 2]);
            }
            float SA = 0.f, SB = 0.f;
            for (int st = 0; st < p.NT; ++st) {
                const float* qa = p.pstat + ((long long)st * p.Bp + rowA) * 2;
                const float* qb = p.pstat + ((long long)st * p.Bp + rowB) * 2;
                SA += qa[1] * expf(qa[0] - MA);
                SB += qb[1] * expf(qb[0] - MB);
            }
            const float iA = 1.f / SA, iB = 1.f / SB;
            const float gA = rowA < p.B ? p.grad_scale : 0.f, gB = rowB < p.B ? p.grad_scale : 0.f;
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = jw + nt * 8 + 2 * t + e;
                    const bool ok = j < p.Bg;
                    lg[nt][e] = ok ? (expf(lg[nt][e] - MA) * iA - (j == labA ? 1.f : 0.f)) * gA : 0.f;
                    lg[nt][2 + e] = ok ? (expf(lg[nt][2 + e] - MB) * iB - (j == labB ? 1.f : 0.f)) * gB : 0.f;
                }
            // ---- dz_a[16][64] += dl . U ; V[16][64] += dl . z_pos   (k = j: 4 steps of 16 columns)
#pragma unroll
            for (int m = 0; m < NTW / 2; ++m) {
                uint32_t dh[4], dm[4], dlo[4];
                split3_bf16(lg[2 * m][0], lg[2 * m][1], dh[0], dm[0], dlo[0]);
                split3_bf16(lg[2 * m][2], lg[2 * m][3], dh[1], dm[1], dlo[1]);
                split3_bf16(lg[2 * m + 1][0], lg[2 * m + 1][1], dh[2], dm[2], dlo[2]);
                split3_bf16(lg[2 * m + 1][2], lg[2 * m + 1][3], dh[3], dm[3], dlo[3]);
#pragma unroll
                for (int na = 0; na < 8; ++na) {
                    const long long off = (long long)(na * 8 + g) * p.BgP + jw + m * 16 + 2 * t;
                    {
                        const uint32_t uh0 = ldg_u32(p.UT_hi + off), uh1 = ldg_u32(p.UT_hi + off + 8);
                        const uint32_t um0 = ldg_u32(p.UT_mid + off), um1 = ldg_u32(p.UT_mid + off + 8);
                        const uint32_t ul0 = ldg_u32(p.UT_lo + off), ul1 = ldg_u32(p.UT_lo + off + 8);
                        mma6(dz[na], dh, dm, dlo, uh0, uh1, um0, um1, ul0, ul1);
                    }
                    {
                        const uint32_t zh0 = ldg_u32(p.ZT_hi + off), zh1 = ldg_u32(p.ZT_hi + off + 8);
                        const uint32_t zm0 = ldg_u32(p.ZT_mid + off), zm1 = ldg_u32(p.ZT_mid + off + 8);
                        const uint32_t zl0 = ldg_u32(p.ZT_lo + off), zl1 = ldg_u32(p.ZT_lo + off + 8);
                        mma6(vv[na], dh, dm, dlo, zh0, zh1, zm0, zm1, zl0, zl1);
                    }
                }
            }
        }
        // ---- the CTA's 8 warps (8 x 64 columns) summed in a fixed order
        float* mine = s_red + (size_t)warp * 2048;
#pragma unroll
        for (int na = 0; na < 8; ++na) {
            const int a = na * 8 + 2 * t;
            *reinterpret_cast<float2*>(mine + g * 64 + a) = make_float2(dz[na][0], dz[na][1]);
            *reinterpret_cast<float2*>(mine + (g + 8) * 64 + a) = make_float2(dz[na][2], dz[na][3]);
            *reinterpret_cast<float2*>(mine + 1024 + g * 64 + a) = make_float2(vv[na][0], vv[na][1]);
            *reinterpret_cast<float2*>(mine + 1024 + (g + 8) * 64 + a) = make_float2(vv[na][2], vv[na][3]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2048; i += 256) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) sum += s_red[(size_t)w * 2048 + i];
            const int which = i >> 10, rem = i & 1023, row = r0 + (rem >> 6);
            float* dst = which ? p.pV : p.pdz;
            if (gridDim.y > 1) dst[((long long)blockIdx.y * p.Bp + row) * 64 + (rem & 63)] = sum;
            else if (row < p.B) dst[(long long)row * 64 + (rem & 63)] = sum;
        }
    }
}

// row_loss[i] = log(sum_j exp(l_ij - M_i)) - (l_i,label - M_i); with more than one column block also
// dz_a / V = sum over the column-block partials (fixed order).
__global__ void __launch_bounds__(256)
k_curl_reduce(const float* __restrict__ pstat, const float* __restrict__ lab_logit, const float* __restrict__ pdz,
              const float* __restrict__ pV, int B, int Bp, int NT, int NCB, float* __restrict__ row_loss,
              float* __restrict__ dz_a, float* __restrict__ V) {
    pdl_grid_sync();
    const int i = blockIdx.x * 4 + (threadIdx.x >> 6), a = threadIdx.x & 63;
    if (i >= B) return;
    if (NCB > 1) {
        float sd = 0.f, sv = 0.f;
        for (int cb = 0; cb < NCB; ++cb) {
            sd += pdz[((long long)cb * Bp + i) * 64 + a];
            sv += pV[((long long)cb * Bp + i) * 64 + a];
        }
        dz_a[(long long)i * 64 + a] = sd;
        V[(long long)i * 64 + a] = sv;
    }
    if (a == 0) {
        float M = -INFINITY, S = 0.f;
        for (int st = 0; st < NT; ++st) M = fmaxf(M, pstat[((long long)st * Bp + i) * 2]);
        for (int st = 0; st < NT; ++st) S += pstat[((long long)st * Bp + i) * 2 + 1] * expf(pstat[((long long)st * Bp + i) * 2] - M);
        row_loss[i] = logf(S) - (lab_logit[i] - M);
    }
}

template <int R>
static int launch_curl_rows(const float* z_a, const float* U, const float* Ut, const float* z_pos, int B, int Bg, int label0,
                            float grad_scale, float* row_loss, float* dz_a, float* V, float* logits_copy,
                            cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)R * ((Bg + 3) & ~3) + R * 64 + 8 * R * 64);
    auto kern = k_curl_rows<R>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_error("curl: %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return -1; }
    }
    launch_k(kern, dim3(cdiv(B, R)), dim3(256), smem, st, z_a, U, Ut, z_pos, B, Bg, label0, grad_scale, row_loss, dz_a, V, logits_copy);
    return check_launch("curl_rows");
}

static int sgemm(const float* A, long long sam, long long sak, const float* Bm, long long sbk,
                 long long sbn, float* C, long long ldc, float* Ct, long long ldct, int M, int N, int K, cudaStream_t st) {
    dim3 grid(cdiv(N, 32), cdiv(M, 32));
    launch_k(k_sgemm_strided, dim3(grid), dim3(256), 0, st, A, sam, sak, Bm, sbk, sbn, C, ldc, Ct, ldct, M, N, K);
    return check_launch("curl_sgemm");
}

}  // namespace curla

using namespace curla;

namespace {
// ntw = 8-column tiles per warp (k_curl_tc): 4 below a global batch of 2048 (twice the warps, half the serial MMA chain per
// warp), 8 above (CURLA_CURL_NTW=4|8 overrides).  NT = column tiles of the softmax statistics, NCB = column blocks (CTAs
// along the keys).  The workspace is sized for the finer split.
struct CurlPlan { long long BgP, Bp, NT, NCB, off_bf16, off_pstat, off_lab, off_pdz, off_pV, total; int ntw; };
int curl_ntw(int Bg) {
    const char* e = getenv("CURLA_CURL_NTW");
    if (e && (e[0] == '4' || e[0] == '8')) return e[0] - '0';
    return Bg < 2048 ? 4 : 8;
}
CurlPlan curl_plan(int B, int Bg, int ntw) {
    CurlPlan c;
    c.ntw = ntw;
    const long long cw = ntw * 8;
    c.BgP = (Bg + 63) / 64 * 64; c.Bp = (B + 15) / 16 * 16;
    c.NT = c.BgP / cw; c.NCB = (c.NT + 7) / 8;
    const long long NT4 = c.BgP / 32, NCB4 = (NT4 + 7) / 8;       // sizes under the finer split
    long long o = 2LL * Bg * 64 + (long long)B * 64 + B;       // U, Ut, V, row_loss (the CUDA-core path's layout)
    o = (o + 3) / 4 * 4;
    c.off_bf16 = o; o += 9 * c.BgP * 64 / 2 + 2;               // nine bf16 [BgP x 64] arrays
    c.off_pstat = o; o += NT4 * c.Bp * 2;
    c.off_lab = o; o += c.Bp;
    c.off_pdz = o; o += NCB4 > 1 ? NCB4 * c.Bp * 64 : 0;
    c.off_pV = o; o += NCB4 > 1 ? NCB4 * c.Bp * 64 : 0;
    c.total = o;
    return c;
}
template <int NTW>
int launch_curl_tc(const CurlTc& t, const CurlPlan& cp, cudaStream_t stream) {
    const dim3 grid((unsigned)(cp.Bp / 16), (unsigned)cp.NCB);
    launch_k(k_curl_tc<0, NTW>, grid, dim3(256), 0, stream, t);
    if (check_launch("curl_rows")) return -1;
    const size_t smem = sizeof(float) * 8 * 2048;
    static bool attr[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr[dev]) {
        cudaError_t e2 = cudaFuncSetAttribute(k_curl_tc<1, NTW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        CURLA_CHECK(e2 == cudaSuccess, "curl: %zu B of shared memory: %s", smem, cudaGetErrorString(e2));
        attr[dev] = true;
    }
    launch_k(k_curl_tc<1, NTW>, grid, dim3(256), smem, stream, t);
    return check_launch("curl_rows");
}
}  // namespace

extern "C" long long curla_curl_workspace_floats(int B, int Bg) { return curl_plan(B, Bg, 4).total; }

// z_a [B][64] (local rows), z_pos [Bg][64] (all ranks' keys), W [feat][feat].
// Outputs: loss_out (mean over the LOCAL rows), dz_a [B][64], dW [feat][feat] (local
// contribution), logits_copy [B][Bg] (optional: the raw logits before the row-max shift).
extern "C" int curla_curl_fwd_bwd(const float* z_a, const float* z_pos, const float* W, int B,
                                  int Bg, int feat, int label0, float grad_scale,
                                  float* workspace, float* loss_out, float* dz_a, float* dW,
                                  float* logits_copy, cudaStream_t stream) {
    CURLA_CHECK(feat <= 64 && B >= 1 && Bg >= B && label0 >= 0 && label0 + B <= Bg, "curl: bad shape");
    float* U = workspace;
    float* Ut = U + (long long)Bg * 64;
    float* V = Ut + (long long)Bg * 64;
    float* row_loss = V + (long long)B * 64;
    {   // tensor-core path (default); CURLA_CURL_TC=0 keeps the CUDA-core rows kernel (A/B, tests)
        const char* e = getenv("CURLA_CURL_TC");
        if (!(e && e[0] == '0')) {
            const CurlPlan cp = curl_plan(B, Bg, curl_ntw(Bg));
            bf16* hb = reinterpret_cast<bf16*>(workspace + cp.off_bf16);
            const long long n = cp.BgP * 64;
            CurlTc t;
            t.z_a = z_a;
            t.Uk_hi = hb; t.Uk_mid = hb + n; t.Uk_lo = hb + 2 * n; t.UT_hi = hb + 3 * n; t.UT_mid = hb + 4 * n; t.UT_lo = hb + 5 * n;
            t.ZT_hi = hb + 6 * n; t.ZT_mid = hb + 7 * n; t.ZT_lo = hb + 8 * n;
            t.pstat = workspace + cp.off_pstat; t.lab_logit = workspace + cp.off_lab;
            t.pdz = cp.NCB > 1 ? workspace + cp.off_pdz : dz_a;
            t.pV = cp.NCB > 1 ? workspace + cp.off_pV : V;
            t.logits_copy = logits_copy;
            t.B = B; t.Bg = Bg; t.BgP = (int)cp.BgP; t.Bp = (int)cp.Bp; t.NT = (int)cp.NT; t.label0 = label0; t.grad_scale = grad_scale;
            // U = z_pos . W^T and the bf16 operand arrays (one launch)
            launch_k(k_curl_prep_u, dim3((unsigned)(cp.BgP / 16)), dim3(256), 0, stream, z_pos, W, Bg, (int)cp.BgP, feat,
                     hb, hb + n, hb + 2 * n, hb + 3 * n, hb + 4 * n, hb + 5 * n, hb + 6 * n, hb + 7 * n, hb + 8 * n);
            if (check_launch("curl_prep")) return -1;
            if ((cp.ntw == 4 ? launch_curl_tc<4>(t, cp, stream) : launch_curl_tc<8>(t, cp, stream))) return -1;
            if (cp.NCB > 1) {
                launch_k(k_curl_reduce, dim3(cdiv(B, 4)), dim3(256), 0, stream, (const float*)t.pstat, (const float*)t.lab_logit,
                         (const float*)t.pdz, (const float*)t.pV, B, (int)cp.Bp, (int)cp.NT, (int)cp.NCB, row_loss, dz_a, V);
                if (check_launch("curl_reduce")) return -1;
                launch_k(k_curl_dw, dim3(feat), dim3(256), 0, stream, z_a, (const float*)V, (const float*)row_loss, (const float*)nullptr,
                         (const float*)nullptr, 0, 0, B, feat, dW, loss_out);
            } else {
                // one column block: dz_a / V are final, the per-row losses are formed in k_curl_dw from the tile statistics
                launch_k(k_curl_dw, dim3(feat), dim3(256), 0, stream, z_a, (const float*)V, (const float*)nullptr, (const float*)t.pstat,
                         (const float*)t.lab_logit, (int)cp.Bp, (int)cp.NT, B, feat, dW, loss_out);
            }
            return check_launch("curl_dw");
        }
    }
    // CUDA-core path: U[j][a] = sum_b z_pos[j][b] * W[a][b]   (columns >= feat stay zero: zero-initialised workspace),
    // written in both orientations: U for the gradient pass, Ut for the coalesced logits pass
    if (sgemm(z_pos, 64, 1, W, 1, feat, U, 64, Ut, Bg, Bg, feat, feat, stream)) return -1;
    // rows per CTA: every CTA streams all of U and z_pos (Bg x 512 B) from L2, so the re-read
    // traffic is (B / R) x that -- at a global batch of 4096 it is what bounds the kernel, and more
    // rows per CTA beat more CTAs; at small Bg the kernel is latency-bound and wants >= ~1 CTA/SM.
    const size_t cap = 200 * 1024 / sizeof(float);
    int R = 8;
    while (R > 1 && ((size_t)R * (Bg + 3) + 9 * R * 64 > cap || (Bg < 2048 && cdiv(B, R) < sm_count() / 2))) R >>= 1;
    CURLA_CHECK((size_t)R * (Bg + 3) + 9 * R * 64 <= cap, "curl: global batch %d does not fit shared memory", Bg);
    int rc;
    switch (R) {
        case 8: rc = launch_curl_rows<8>(z_a, U, Ut, z_pos, B, Bg, label0, grad_scale, row_loss, dz_a, V, logits_copy, stream); break;
        case 4: rc = launch_curl_rows<4>(z_a, U, Ut, z_pos, B, Bg, label0, grad_scale, row_loss, dz_a, V, logits_copy, stream); break;
        case 2: rc = launch_curl_rows<2>(z_a, U, Ut, z_pos, B, Bg, label0, grad_scale, row_loss, dz_a, V, logits_copy, stream); break;
        default: rc = launch_curl_rows<1>(z_a, U, Ut, z_pos, B, Bg, label0, grad_scale, row_loss, dz_a, V, logits_copy, stream); break;
    }
    if (rc) return -1;
    launch_k(k_curl_dw, dim3(feat), dim3(256), 0, stream, z_a, (const float*)V, (const float*)row_loss, (const float*)nullptr,
             (const float*)nullptr, 0, 0, B, feat, dW, loss_out);
    return check_launch("curl_dw");
}
