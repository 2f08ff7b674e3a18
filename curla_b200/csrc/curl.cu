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
                float* __restrict__ C, long long ldc, int M, int N, int K) {
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
            if (m < M && n < N) C[m * ldc + n] = acc[i][j];
        }
}

// one CTA per local row: softmax stats, per-row loss, logits -> dlogits in place.
__global__ void __launch_bounds__(256)
k_curl_softmax(float* __restrict__ logits, int B, int Bg, int label0, float grad_scale,
               float* __restrict__ row_loss) {
    __shared__ float s_red[8];
    __shared__ float s_b;
    const int i = blockIdx.x;
    float* row = logits + (long long)i * Bg;
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < Bg; j += blockDim.x) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = s_red[0];
        for (int w = 1; w < (blockDim.x >> 5); ++w) m = fmaxf(m, s_red[w]);
        s_b = m;
    }
    __syncthreads();
    mx = s_b;
    float se = 0.f;
    for (int j = threadIdx.x; j < Bg; j += blockDim.x) se += expf(row[j] - mx);
    se = warp_sum(se);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = se;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) s += s_red[w];
        s_b = s;
        row_loss[i] = logf(s) - (row[label0 + i] - mx);
    }
    __syncthreads();
    const float inv = 1.f / s_b;
    for (int j = threadIdx.x; j < Bg; j += blockDim.x) {
        const float p = expf(row[j] - mx) * inv;
        row[j] = (p - (j == label0 + i ? 1.f : 0.f)) * grad_scale;
    }
}

__global__ void __launch_bounds__(256)
k_mean_to(const float* __restrict__ v, int n, float* __restrict__ out) {
    __shared__ float s_red[8];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += s_red[w];
        *out = t / n;
    }
}

static int sgemm(const float* A, long long sam, long long sak, const float* Bm, long long sbk,
                 long long sbn, float* C, long long ldc, int M, int N, int K, cudaStream_t st) {
    dim3 grid(cdiv(N, 32), cdiv(M, 32));
    k_sgemm_strided<<<grid, 256, 0, st>>>(A, sam, sak, Bm, sbk, sbn, C, ldc, M, N, K);
    return check_launch("curl_sgemm");
}

}  // namespace curla

using namespace curla;

extern "C" long long curla_curl_workspace_floats(int B, int Bg) {
    // U [Bg][64] + T [Bg][64] + logits [B][Bg] + row_loss [B]
    return 2LL * Bg * 64 + (long long)B * Bg + B;
}

// z_a [B][64] (local rows), z_pos [Bg][64] (all ranks' keys), W [feat][feat].
// Outputs: loss_out (mean over the LOCAL rows), dz_a [B][64] (padded cols untouched ->
// caller passes a zeroed buffer once), dW [feat][feat] (local contribution).
extern "C" int curla_curl_fwd_bwd(const float* z_a, const float* z_pos, const float* W, int B,
                                  int Bg, int feat, int label0, float grad_scale,
                                  float* workspace, float* loss_out, float* dz_a, float* dW,
                                  float* logits_copy, cudaStream_t stream) {
    float* U = workspace;
    float* T = U + (long long)Bg * 64;
    float* logits = T + (long long)Bg * 64;
    float* row_loss = logits + (long long)B * Bg;
    // U[j][a] = sum_b z_pos[j][b] * W[a][b]
    if (sgemm(z_pos, 64, 1, W, 1, feat, U, 64, Bg, feat, feat, stream)) return -1;
    // logits[i][j] = sum_a z_a[i][a] * U[j][a]
    if (sgemm(z_a, 64, 1, U, 1, 64, logits, Bg, B, Bg, feat, stream)) return -1;
    if (logits_copy)
        cudaMemcpyAsync(logits_copy, logits, sizeof(float) * (size_t)B * Bg, cudaMemcpyDeviceToDevice, stream);
    k_curl_softmax<<<B, 256, 0, stream>>>(logits, B, Bg, label0, grad_scale, row_loss);
    if (check_launch("curl_softmax")) return -1;
    k_mean_to<<<1, 256, 0, stream>>>(row_loss, B, loss_out);
    if (check_launch("curl_loss")) return -1;
    // dz_a[i][a] = sum_j dl[i][j] * U[j][a]
    if (sgemm(logits, Bg, 1, U, 64, 1, dz_a, 64, B, feat, Bg, stream)) return -1;
    // T[j][a] = sum_i dl[i][j] * z_a[i][a]
    if (sgemm(logits, 1, Bg, z_a, 64, 1, T, 64, Bg, feat, B, stream)) return -1;
    // dW[a][b] = sum_j T[j][a] * z_pos[j][b]
    return sgemm(T, 1, 64, z_pos, 64, 1, dW, feat, feat, feat, Bg, stream);
}
