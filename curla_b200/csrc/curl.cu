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
// tree => deterministic); CTA 0 also writes the mean of the per-row losses.
__global__ void __launch_bounds__(256)
k_curl_dw(const float* __restrict__ z_a, const float* __restrict__ V, const float* __restrict__ row_loss,
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
        for (int i = threadIdx.x; i < B; i += 256) l += row_loss[i];
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

extern "C" long long curla_curl_workspace_floats(int B, int Bg) {
    // U [Bg][64] + Ut [64][Bg] + V [B][64] + row_loss [B]
    return 2LL * Bg * 64 + (long long)B * 64 + B;
}

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
    // U[j][a] = sum_b z_pos[j][b] * W[a][b]   (columns >= feat stay zero: zero-initialised workspace),
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
    launch_k(k_curl_dw, dim3(feat), dim3(256), 0, stream, z_a, V, row_loss, B, feat, dW, loss_out);
    return check_launch("curl_dw");
}
