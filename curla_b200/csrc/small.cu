// K8 LayerNorm, K10 tanh-Gaussian policy head, K11 SAC losses, and the skinny MLP
// head layers: warp-shuffle / single-CTA kernels around the tensor-core GEMMs.
//
//   encoder.py:98-110      fc bias + LayerNorm(eps 1e-5, affine) [+ tanh]
//   curl_sac.py:20-35      gaussian_logprob, squash
//   curl_sac.py:79-110     Actor.forward tail (chunk, tanh-rescaled log_std, reparam sample)
//   curl_sac.py:349-404    critic / actor / alpha losses (and their gradients)
#include "common.cuh"
#include "../../include/curla_b200.h"
#include <math.h>

namespace curla {

constexpr int FP = 64;   // padded feature width of every [B][FP] fp32 / bf16 row buffer

// ------------------------------------------------------------------ LayerNorm fwd
// x = sum_s partial[s][b][:] + bias ; z = LN(x)*gamma + beta ; optional tanh.
// Saves x (pre-norm) for the backward.  One CTA per row: 4 split groups x 64 columns sum the
// split-K partials of the fc GEMM (fixed order => deterministic), then warp 0 normalises
// (lane owns cols lane, lane+32).
__global__ void __launch_bounds__(256)
k_ln_fwd(const float* __restrict__ partial, int nsplit, long long split_stride,
         const float* __restrict__ bias, const float* __restrict__ gamma,
         const float* __restrict__ beta, int B, int feat, int apply_tanh,
         float* __restrict__ x_out, float* __restrict__ z_out,
         const float* __restrict__ act, int A, bf16* __restrict__ X_out) {
    pdl_grid_sync();
    __shared__ float s_x[4][FP];
    const int row = blockIdx.x;
    {
        const int c = threadIdx.x & 63, sg = threadIdx.x >> 6;
        float s = 0.f;
        if (c < feat)
            for (int k = sg; k < nsplit; k += 4) s += partial[k * split_stride + (long long)row * FP + c];
        s_x[sg][c] = s;
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    float x[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int c = lane + e * 32;
        x[e] = c < feat ? ((s_x[0][c] + s_x[1][c]) + (s_x[2][c] + s_x[3][c])) + bias[c] : 0.f;
    }
    const float mean = warp_sum(x[0] + x[1]) / feat;
    float d0 = lane < feat ? x[0] - mean : 0.f, d1 = lane + 32 < feat ? x[1] - mean : 0.f;
    const float var = warp_sum(d0 * d0 + d1 * d1) / feat;
    const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int c = lane + e * 32;
        float z = 0.f;
        if (c < feat) {
            z = (e ? d1 : d0) * rstd * gamma[c] + beta[c];
            if (apply_tanh) z = tanhf(z);
        }
        x_out[(long long)row * FP + c] = x[e];
        z_out[(long long)row * FP + c] = z;
        if (X_out) {   // the MLP input row [z | action | 0] (torch.cat([obs, action], dim=1), curl_sac.py:137)
            if (act && c >= feat && c < feat + A) z = act[row * A + (c - feat)];
            X_out[(long long)row * FP + c] = __float2bfloat16(z);
        }
    }
}

// ------------------------------------------------------------------ LayerNorm bwd
// dz = dz_a (+ dz_b); writes dx (fp32 + bf16, padded cols zero) and dz*xhat for the
// parameter reduction.
__global__ void __launch_bounds__(256)
k_ln_bwd(const float* __restrict__ dz_a, const float* __restrict__ dz_b,
         const float* __restrict__ x_in, const float* __restrict__ gamma, int B, int feat,
         float* __restrict__ dx_f32, bf16* __restrict__ dx_bf16, float* __restrict__ dzsum,
         float* __restrict__ dzx) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B) return;
    float x[2], dz[2], g[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int c = lane + e * 32;
        const long long o = (long long)row * FP + c;
        const bool ok = c < feat;
        x[e] = ok ? x_in[o] : 0.f;
        dz[e] = ok ? dz_a[o] + (dz_b ? dz_b[o] : 0.f) : 0.f;
        g[e] = ok ? gamma[c] : 0.f;
    }
    const float mean = warp_sum(x[0] + x[1]) / feat;
    const float d0 = lane < feat ? x[0] - mean : 0.f, d1 = lane + 32 < feat ? x[1] - mean : 0.f;
    const float var = warp_sum(d0 * d0 + d1 * d1) / feat;
    const float rstd = rsqrtf(var + 1e-5f);
    const float xh0 = d0 * rstd, xh1 = d1 * rstd;
    const float dh0 = dz[0] * g[0], dh1 = dz[1] * g[1];
    const float m1 = warp_sum(dh0 + dh1) / feat;
    const float m2 = warp_sum(dh0 * xh0 + dh1 * xh1) / feat;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int c = lane + e * 32;
        const long long o = (long long)row * FP + c;
        float dx = 0.f;
        if (c < feat) dx = rstd * ((e ? dh1 : dh0) - m1 - (e ? xh1 : xh0) * m2);
        dx_f32[o] = dx;
        dx_bf16[o] = __float2bfloat16(dx);
        dzsum[o] = dz[e];
        dzx[o] = dz[e] * (e ? xh1 : xh0);
    }
}

// column sums over B rows of up to three [B][FP] fp32 buffers -> three [feat] vectors.
// Single CTA (deterministic): 64 columns x 16 row groups.
__global__ void __launch_bounds__(1024)
k_colsum3(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
          int B, int feat, float* __restrict__ oa, float* __restrict__ ob, float* __restrict__ oc) {
    pdl_grid_sync();
    __shared__ float s[3][16][FP];
    const int col = threadIdx.x & 63, rg = threadIdx.x >> 6;
    float sa = 0.f, sb = 0.f, sc = 0.f;
#pragma unroll 8
    for (int r = rg; r < B; r += 16) {
        const long long o = (long long)r * FP + col;
        sa += a[o]; sb += b[o]; sc += c[o];
    }
    s[0][rg][col] = sa; s[1][rg][col] = sb; s[2][rg][col] = sc;
    __syncthreads();
    if (threadIdx.x < 3 * FP) {
        const int w = threadIdx.x / FP, cc = threadIdx.x % FP;
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) t += s[w][r][cc];
        if (cc < feat) (w == 0 ? oa : (w == 1 ? ob : oc))[cc] = t;
    }
}

// X[b][0:feat] = z, X[b][feat:feat+A] = act (optional), rest 0   (bf16 MLP input rows)
__global__ void k_pack_x(const float* __restrict__ z, const float* __restrict__ act, int B,
                         int feat, int A, bf16* __restrict__ X) {
    pdl_grid_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * FP) return;
    const int b = i / FP, c = i % FP;
    float v = 0.f;
    if (c < feat) v = z[i];
    else if (act && c < feat + A) v = act[b * A + (c - feat)];
    X[i] = __float2bfloat16(v);
}

// ------------------------------------------------------------------ skinny head layer
// out[b][o] = sum_j H[b][j] * W[o][j] + bias[o],  No <= 4.  One warp per row.
__global__ void __launch_bounds__(256)
k_head_fwd(const bf16* __restrict__ H, int ldh, const float* __restrict__ W,
           const float* __restrict__ bias, int B, int hid, int No, float* __restrict__ out,
           long long bsH, long long bsW, long long bsOut) {
    pdl_grid_sync();
    H += blockIdx.y * bsH; W += blockIdx.y * bsW; bias += blockIdx.y * bsW; out += blockIdx.y * bsOut;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bf16* hrow = H + (long long)row * ldh;
    if ((hid & 7) == 0 && (ldh & 7) == 0 && ((reinterpret_cast<uintptr_t>(H) | reinterpret_cast<uintptr_t>(W)) & 15) == 0) {
        // 8 hidden units (one 16-byte load of H, two of each W row) per lane per step; the steps of
        // a row are independent, so the compiler keeps all of their loads in flight
#pragma unroll 4
        for (int j = lane * 8; j < hid; j += 256) {
            const uint4 hv = *reinterpret_cast<const uint4*>(hrow + j);
            const float2 h0 = unpack_bf16x2(hv.x), h1 = unpack_bf16x2(hv.y), h2 = unpack_bf16x2(hv.z), h3 = unpack_bf16x2(hv.w);
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (o < No) {
                    const float4 w0 = *reinterpret_cast<const float4*>(W + o * hid + j);
                    const float4 w1 = *reinterpret_cast<const float4*>(W + o * hid + j + 4);
                    acc[o] += h0.x * w0.x + h0.y * w0.y + h1.x * w0.z + h1.y * w0.w +
                              h2.x * w1.x + h2.y * w1.y + h3.x * w1.z + h3.y * w1.w;
                }
        }
    } else {
        for (int j = lane * 2; j < hid; j += 64) {
            const float2 h = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(hrow + j));
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (o < No) acc[o] += h.x * W[o * hid + j] + h.y * W[o * hid + j + 1];
        }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
        if (o < No) {
            const float s = warp_sum(acc[o]);
            if (lane == 0) out[row * No + o] = s + bias[o];
        }
}

// dH[b][j] = H[b][j] > 0 ? sum_o dOut[b][o] * W[o][j] : 0
__global__ void __launch_bounds__(256)
k_head_bwd(const float* __restrict__ dOut, const float* __restrict__ W,
           const bf16* __restrict__ H, int B, int hid, int No, bf16* __restrict__ dH,
           long long bsDOut, long long bsW, long long bsH, long long bsDH) {
    pdl_grid_sync();
    dOut += blockIdx.y * bsDOut; W += blockIdx.y * bsW; H += blockIdx.y * bsH; dH += blockIdx.y * bsDH;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * hid) return;
    const int b = (int)(i / hid), j = (int)(i % hid);
    float s = 0.f;
    for (int o = 0; o < No; ++o) s += dOut[b * No + o] * W[o * hid + j];
    dH[i] = __float2bfloat16(__bfloat162float(H[i]) > 0.f ? s : 0.f);
}

// dW[o][j] = sum_b dOut[b][o]*H[b][j];  db[o] = sum_b dOut[b][o]
// block = 64 columns (32 bf16x2 lanes) x 16 row groups, fixed-order shared-memory tree
// (deterministic); grid = hid / 64.
__global__ void __launch_bounds__(512)
k_head_wgrad(const float* __restrict__ dOut, const bf16* __restrict__ H, int B, int hid, int No,
             float* __restrict__ dW, float* __restrict__ db, long long bsDOut, long long bsH, long long bsW) {
    pdl_grid_sync();
    dOut += blockIdx.y * bsDOut; H += blockIdx.y * bsH; dW += blockIdx.y * bsW; db += blockIdx.y * bsW;
    __shared__ float red[16][4][64];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int j = (blockIdx.x * 32 + tx) * 2;
    float acc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    if (j < hid) {
#pragma unroll 8
        for (int b = ty; b < B; b += 16) {
            const float2 h = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(H + (long long)b * hid + j));
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (o < No) {
                    const float d = dOut[b * No + o];
                    acc[o][0] += d * h.x; acc[o][1] += d * h.y;
                }
        }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) { red[ty][o][tx * 2] = acc[o][0]; red[ty][o][tx * 2 + 1] = acc[o][1]; }
    __syncthreads();
    const int t = ty * 32 + tx;                      // 0..511: (o, column) pairs of this block
    if (t < 4 * 64) {
        const int o = t >> 6, c = t & 63;
        if (o < No && blockIdx.x * 64 + c < hid) {
            float sum = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) sum += red[r][o][c];
            dW[o * hid + blockIdx.x * 64 + c] = sum;
        }
    }
    if (blockIdx.x == 0 && ty == 0 && tx < No) {
        float sum = 0.f;
        for (int b = 0; b < B; ++b) sum += dOut[b * No + tx];
        db[tx] = sum;
    }
}

// db[j] = sum_b dH[b][j]  (bias grads of the hidden layers); same block shape as above
__global__ void __launch_bounds__(512)
k_colsum_bf16(const bf16* __restrict__ dH, int B, int hid, float* __restrict__ db, long long bsDH,
              long long bsDb) {
    pdl_grid_sync();
    dH += blockIdx.y * bsDH; db += blockIdx.y * bsDb;
    __shared__ float red[16][64];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int j = (blockIdx.x * 32 + tx) * 2;
    float s0 = 0.f, s1 = 0.f;
    if (j < hid) {
#pragma unroll 8
        for (int b = ty; b < B; b += 16) {
            const float2 h = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dH + (long long)b * hid + j));
            s0 += h.x; s1 += h.y;
        }
    }
    red[ty][tx * 2] = s0; red[ty][tx * 2 + 1] = s1;
    __syncthreads();
    const int t = ty * 32 + tx;
    if (t < 64 && blockIdx.x * 64 + t < hid) {
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) sum += red[r][t];
        db[blockIdx.x * 64 + t] = sum;
    }
}

// ------------------------------------------------------------------ policy head
// Philox4x32-10 for the built-in noise source (torch.randn_like replacement,
// curl_sac.py:97); tests inject the noise instead.
// t[b][0:A]=mu_raw, t[b][A:2A]=log_std_raw.  A <= 4.
// Outputs: mu=tanh(mu_raw), pi=tanh(mu_raw + n*exp(ls)), log_pi, ls (rescaled log_std),
// noise_out (the noise actually used; kept for the backward).
__global__ void k_policy_fwd(const float* __restrict__ t, const float* __restrict__ noise_in,
                             unsigned long long seed, unsigned long long offset, int B, int A,
                             float ls_min, float ls_max, int compute_pi, int compute_log_pi,
                             float* __restrict__ mu_out, float* __restrict__ pi_out,
                             float* __restrict__ log_pi_out, float* __restrict__ ls_out,
                             float* __restrict__ noise_out, int row0,
                             const unsigned long long* __restrict__ offset_dev) {
    pdl_grid_sync();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (offset_dev) offset += *offset_dev;          // replayed CUDA graph: the per-update counter lives in device memory
    float n[4] = {0.f, 0.f, 0.f, 0.f};
    if (compute_pi) {
        if (noise_in) {
            for (int k = 0; k < A; ++k) n[k] = noise_in[b * A + k];
        } else {
            // counter = GLOBAL row: a data-parallel shard draws exactly the rows of the global batch's noise
            const uint4 r = philox4x32(make_uint4((uint32_t)(b + row0), (uint32_t)offset, (uint32_t)(offset >> 32), 0u),
                                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            const float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
            n[0] = g0.x; n[1] = g0.y; n[2] = g1.x; n[3] = g1.y;
        }
    }
    float lp = 0.f, corr = 0.f;
    for (int k = 0; k < A; ++k) {
        const float mu = t[b * 2 * A + k];
        float ls = tanhf(t[b * 2 * A + A + k]);
        ls = ls_min + 0.5f * (ls_max - ls_min) * (ls + 1.f);
        if (ls_out) ls_out[b * A + k] = ls;
        if (mu_out) mu_out[b * A + k] = tanhf(mu);
        if (compute_pi) {
            const float pi = tanhf(mu + n[k] * expf(ls));
            if (pi_out) pi_out[b * A + k] = pi;
            if (noise_out) noise_out[b * A + k] = n[k];
            lp += -0.5f * n[k] * n[k] - ls;
            corr += logf(fmaxf(1.f - pi * pi, 0.f) + 1e-6f);
        }
    }
    if (compute_pi && compute_log_pi && log_pi_out)
        log_pi_out[b] = lp - 0.5f * 1.8378770664093453f * A - corr;   // log(2*pi)
}

// d(trunk_out) from dL/dpi (= dx1[:, feat:feat+A] + dx2[...]) and dL/dlog_pi = glogpi.
__global__ void k_policy_bwd(const float* __restrict__ dx1, const float* __restrict__ dx2, int feat,
                             const float* __restrict__ glogpi_ptr, const float* __restrict__ t,
                             const float* __restrict__ noise, const float* __restrict__ pi_in,
                             const float* __restrict__ ls_in, int B, int A, float ls_min,
                             float ls_max, float* __restrict__ dt) {
    pdl_grid_sync();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float glp = *glogpi_ptr;
    for (int k = 0; k < A; ++k) {
        const float pi = pi_in[b * A + k], ls = ls_in[b * A + k], n = noise[b * A + k];
        float gpi = dx1[(long long)b * FP + feat + k] + (dx2 ? dx2[(long long)b * FP + feat + k] : 0.f);
        const float om = 1.f - pi * pi;
        // log_pi -= log(relu(1-pi^2)+1e-6)
        if (om > 0.f) gpi += glp * (2.f * pi / (om + 1e-6f));
        const float du = gpi * om;                       // through pi = tanh(u)
        const float dls = du * n * expf(ls) - glp;       // u = mu + n*exp(ls); log_pi has -ls
        const float th = tanhf(t[b * 2 * A + A + k]);
        dt[b * 2 * A + k] = du;
        dt[b * 2 * A + A + k] = dls * 0.5f * (ls_max - ls_min) * (1.f - th * th);
    }
}

// ------------------------------------------------------------------ SAC losses (single CTA)
__device__ __forceinline__ float block_sum(float v, float* s_red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += s_red[i];
    return t;
}

// target_Q = r + not_done*discount*(min(tq1,tq2) - alpha*log_pi')      curl_sac.py:353-355
// loss = mse(q1,t) + mse(q2,t); dq = 2(q-t)/B                           curl_sac.py:359
// metrics[0] = mean(reward), metrics[1] = critic loss
__global__ void __launch_bounds__(256)
k_critic_loss(const float* __restrict__ tq1, const float* __restrict__ tq2,
              const float* __restrict__ logpi_next, const float* __restrict__ reward,
              const float* __restrict__ not_done, const double* __restrict__ log_alpha,
              float discount, const float* __restrict__ q1, const float* __restrict__ q2, int B,
              float grad_scale, float* __restrict__ target_q, float* __restrict__ dq1,
              float* __restrict__ dq2, float* __restrict__ metrics) {
    pdl_grid_sync();
    __shared__ float s_red[8];
    const float alpha = (float)exp(*log_alpha);
    float l = 0.f, rs = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float tv = fminf(tq1[b], tq2[b]) - alpha * logpi_next[b];
        const float tq = reward[b] + not_done[b] * discount * tv;
        target_q[b] = tq;
        const float e1 = q1[b] - tq, e2 = q2[b] - tq;
        l += e1 * e1 + e2 * e2;
        rs += reward[b];
        dq1[b] = 2.f * e1 * grad_scale;
        dq2[b] = 2.f * e2 * grad_scale;
    }
    l = block_sum(l, s_red);
    rs = block_sum(rs, s_red);
    if (threadIdx.x == 0) { metrics[0] = rs / B; metrics[1] = l / B; }
}

// actor_loss = mean(alpha*log_pi - min(q1,q2));  alpha_loss = mean(alpha*(-log_pi - te))
// dq = -grad_scale on the min branch; glogpi = alpha*grad_scale; d(alpha_loss)/d(log_alpha)
// entropy = 0.5*A*(1+log 2pi) + sum log_std                         curl_sac.py:373-404
__global__ void __launch_bounds__(256)
k_actor_loss(const float* __restrict__ log_pi, const float* __restrict__ q1,
             const float* __restrict__ q2, const float* __restrict__ ls, int B, int A,
             const double* __restrict__ log_alpha, float target_entropy, float grad_scale,
             float* __restrict__ dq1, float* __restrict__ dq2, float* __restrict__ glogpi,
             double* __restrict__ g_log_alpha, float* __restrict__ metrics) {
    pdl_grid_sync();
    __shared__ float s_red[8];
    const double alpha_d = exp(*log_alpha);
    const float alpha = (float)alpha_d;
    float la = 0.f, ent = 0.f, lal = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const bool first = q1[b] <= q2[b];
        la += alpha * log_pi[b] - (first ? q1[b] : q2[b]);
        dq1[b] = first ? -grad_scale : 0.f;
        dq2[b] = first ? 0.f : -grad_scale;
        float e = 0.f;
        for (int k = 0; k < A; ++k) e += ls[b * A + k];
        ent += 0.5f * A * (1.f + 1.8378770664093453f) + e;
        lal += -log_pi[b] - target_entropy;
    }
    la = block_sum(la, s_red);
    ent = block_sum(ent, s_red);
    lal = block_sum(lal, s_red);
    if (threadIdx.x == 0) {
        metrics[2] = la / B; metrics[3] = ent / B;
        metrics[4] = alpha * (lal / B); metrics[5] = alpha; metrics[7] = target_entropy;
        *glogpi = alpha * grad_scale;
        // local-shard contribution; grad_scale = 1/(global batch)
        *g_log_alpha = alpha_d * (double)(lal * grad_scale);
    }
}

// out = a + b over n floats (sums the two Q heads' input gradients)
__global__ void k_add2(const float* __restrict__ a, const float* __restrict__ b, long long n,
                       float* __restrict__ out) {
    pdl_grid_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}

}  // namespace curla

using namespace curla;

extern "C" int curla_ln_fwd(const float* partial, int nsplit, long long split_stride,
                            const float* bias, const float* gamma, const float* beta, int B,
                            int feat, int apply_tanh, float* x_out, float* z_out,
                            cudaStream_t stream) {
    return curla_ln_fwd_x(partial, nsplit, split_stride, bias, gamma, beta, B, feat, apply_tanh, x_out, z_out,
                          nullptr, 0, nullptr, stream);
}

// Same, and also emits the bf16 input row of the MLP that consumes z: X[b] = [z | act[b] | 0].
extern "C" int curla_ln_fwd_x(const float* partial, int nsplit, long long split_stride,
                              const float* bias, const float* gamma, const float* beta, int B,
                              int feat, int apply_tanh, float* x_out, float* z_out,
                              const float* act, int A, void* X_out, cudaStream_t stream) {
    CURLA_CHECK(feat <= FP && feat + A <= FP, "ln_fwd: feature_dim (+ action_dim) > 64 unsupported");
    launch_k(k_ln_fwd, dim3(B), dim3(256), 0, stream, partial, nsplit, split_stride, bias, gamma, beta, B, feat, apply_tanh,
                                    x_out, z_out, act, A, (bf16*)X_out);
    return check_launch("ln_fwd");
}

// scratch: 2*B*FP floats (dz sum, dz*xhat)
extern "C" int curla_ln_bwd(const float* dz_a, const float* dz_b, const float* x_in,
                            const float* gamma, int B, int feat, float* dx_f32, void* dx_bf16,
                            float* scratch, float* dgamma, float* dbeta, float* dbias_fc,
                            cudaStream_t stream) {
    float* dzsum = scratch;
    float* dzx = scratch + (long long)B * FP;
    launch_k(k_ln_bwd, dim3(cdiv(B, 8)), dim3(256), 0, stream, dz_a, dz_b, x_in, gamma, B, feat, dx_f32,
                                             (bf16*)dx_bf16, dzsum, dzx);
    if (check_launch("ln_bwd")) return -1;
    launch_k(k_colsum3, dim3(1), dim3(1024), 0, stream, dzx, dzsum, dx_f32, B, feat, dgamma, dbeta, dbias_fc);
    return check_launch("ln_bwd_params");
}

extern "C" int curla_pack_x(const float* z, const float* act, int B, int feat, int A, void* X,
                            cudaStream_t stream) {
    launch_k(k_pack_x, dim3(cdiv((long long)B * FP, 256)), dim3(256), 0, stream, z, act, B, feat, A, (bf16*)X);
    return check_launch("pack_x");
}

// The *_batched forms run `nb` independent heads of one shape in one launch (Q1 || Q2);
// strides in elements: bsP applies to every parameter / parameter-gradient pointer (W and
// bias live at the same relative offsets in both heads).
extern "C" int curla_head_fwd_batched(const void* H, int ldh, const float* W, const float* bias, int B,
                                      int hid, int No, float* out, int nb, long long bsH, long long bsP,
                                      long long bsOut, cudaStream_t stream) {
    CURLA_CHECK(No <= 4 && hid % 2 == 0, "head_fwd: No<=4, even hidden");
    launch_k(k_head_fwd, dim3(cdiv(B, 8), nb), dim3(256), 0, stream, (const bf16*)H, ldh, W, bias, B, hid, No, out, bsH, bsP, bsOut);
    return check_launch("head_fwd");
}
extern "C" int curla_head_fwd(const void* H, int ldh, const float* W, const float* bias, int B,
                              int hid, int No, float* out, cudaStream_t stream) {
    return curla_head_fwd_batched(H, ldh, W, bias, B, hid, No, out, 1, 0, 0, 0, stream);
}

extern "C" int curla_head_bwd_batched(const float* dOut, const float* W, const void* H, int B, int hid,
                                      int No, void* dH, int nb, long long bsDOut, long long bsP,
                                      long long bsH, long long bsDH, cudaStream_t stream) {
    launch_k(k_head_bwd, dim3(cdiv((long long)B * hid, 256), nb), dim3(256), 0, stream, dOut, W, (const bf16*)H, B, hid, No,
                                                                         (bf16*)dH, bsDOut, bsP, bsH, bsDH);
    return check_launch("head_bwd");
}
extern "C" int curla_head_bwd(const float* dOut, const float* W, const void* H, int B, int hid,
                              int No, void* dH, cudaStream_t stream) {
    return curla_head_bwd_batched(dOut, W, H, B, hid, No, dH, 1, 0, 0, 0, 0, stream);
}

extern "C" int curla_head_wgrad_batched(const float* dOut, const void* H, int B, int hid, int No,
                                        float* dW, float* db, int nb, long long bsDOut, long long bsH,
                                        long long bsP, cudaStream_t stream) {
    launch_k(k_head_wgrad, dim3(cdiv(hid, 64), nb), dim3(32, 16), 0, stream, dOut, (const bf16*)H, B, hid, No, dW, db,
                                                                      bsDOut, bsH, bsP);
    return check_launch("head_wgrad");
}
extern "C" int curla_head_wgrad(const float* dOut, const void* H, int B, int hid, int No,
                                float* dW, float* db, cudaStream_t stream) {
    return curla_head_wgrad_batched(dOut, H, B, hid, No, dW, db, 1, 0, 0, 0, stream);
}

extern "C" int curla_colsum_bf16_batched(const void* dH, int B, int hid, float* db, int nb, long long bsDH,
                                         long long bsP, cudaStream_t stream) {
    launch_k(k_colsum_bf16, dim3(cdiv(hid, 64), nb), dim3(32, 16), 0, stream, (const bf16*)dH, B, hid, db, bsDH, bsP);
    return check_launch("colsum_bf16");
}
extern "C" int curla_colsum_bf16(const void* dH, int B, int hid, float* db, cudaStream_t stream) {
    return curla_colsum_bf16_batched(dH, B, hid, db, 1, 0, 0, stream);
}

extern "C" int curla_policy_fwd(const float* t, const float* noise_in, unsigned long long seed,
                                unsigned long long offset, int B, int A, float ls_min,
                                float ls_max, int compute_pi, int compute_log_pi, float* mu,
                                float* pi, float* log_pi, float* ls, float* noise_out,
                                cudaStream_t stream) {
    return curla_policy_fwd_rows(t, noise_in, seed, offset, 0, B, A, ls_min, ls_max, compute_pi, compute_log_pi, mu, pi,
                                 log_pi, ls, noise_out, stream);
}

// Same for rows [row0, row0 + B) of a larger (global) batch: the built-in Philox noise of local
// row b is the noise global row row0 + b would get on a single GPU.
// The update's logged scalars, published to the host while the rest of the update is still running: 15 floats +
// a sequence word into mapped pinned host memory (payload, system-scope fence, then the word the host polls).
__global__ void k_publish_metrics(const float* __restrict__ metrics, float* __restrict__ mailbox, unsigned seq_host,
                                  const unsigned long long* __restrict__ seq_dev) {
    pdl_grid_sync();
    const int t = threadIdx.x;
    if (t < 15) {
        reinterpret_cast<volatile float*>(mailbox)[t] = metrics[t];
        __threadfence_system();
    }
    __syncthreads();
    if (t == 0) {
        const unsigned seq = seq_dev ? (unsigned)((*seq_dev >> 1) + 1ull) : seq_host;     // dev_state holds 2 * update count
        reinterpret_cast<volatile unsigned*>(mailbox)[15] = seq;
        __threadfence_system();
    }
}
extern "C" int curla_publish_metrics(const float* metrics, float* mailbox_host_mapped, unsigned seq,
                                     const unsigned long long* seq_dev, cudaStream_t stream) {
    launch_k(k_publish_metrics, dim3(1), dim3(32), 0, stream, metrics, mailbox_host_mapped, seq, seq_dev);
    return check_launch("publish_metrics");
}

extern "C" int curla_policy_fwd_rows(const float* t, const float* noise_in, unsigned long long seed,
                                     unsigned long long offset, int row0, int B, int A, float ls_min,
                                     float ls_max, int compute_pi, int compute_log_pi, float* mu,
                                     float* pi, float* log_pi, float* ls, float* noise_out,
                                     cudaStream_t stream) {
    return curla_policy_fwd_rows_dyn(t, noise_in, seed, offset, nullptr, row0, B, A, ls_min, ls_max, compute_pi, compute_log_pi,
                                     mu, pi, log_pi, ls, noise_out, stream);
}

// Same with the Philox offset = offset + *offset_dev (offset_dev: device pointer or NULL): lets a captured
// CUDA graph of the update draw fresh policy noise on every replay.
extern "C" int curla_policy_fwd_rows_dyn(const float* t, const float* noise_in, unsigned long long seed,
                                         unsigned long long offset, const unsigned long long* offset_dev, int row0,
                                         int B, int A, float ls_min, float ls_max, int compute_pi,
                                         int compute_log_pi, float* mu, float* pi, float* log_pi, float* ls,
                                         float* noise_out, cudaStream_t stream) {
    CURLA_CHECK(A <= 4, "policy_fwd: action dim > 4 unsupported");
    launch_k(k_policy_fwd, dim3(cdiv(B, 128)), dim3(128), 0, stream, t, noise_in, seed, offset, B, A, ls_min, ls_max,
                                                   compute_pi, compute_log_pi, mu, pi, log_pi, ls,
                                                   noise_out, row0, offset_dev);
    return check_launch("policy_fwd");
}

extern "C" int curla_policy_bwd(const float* dx1, const float* dx2, int feat, const float* glogpi,
                                const float* t, const float* noise, const float* pi,
                                const float* ls, int B, int A, float ls_min, float ls_max,
                                float* dt, cudaStream_t stream) {
    launch_k(k_policy_bwd, dim3(cdiv(B, 128)), dim3(128), 0, stream, dx1, dx2, feat, glogpi, t, noise, pi, ls, B, A,
                                                   ls_min, ls_max, dt);
    return check_launch("policy_bwd");
}

extern "C" int curla_critic_loss(const float* tq1, const float* tq2, const float* logpi_next,
                                 const float* reward, const float* not_done,
                                 const double* log_alpha, float discount, const float* q1,
                                 const float* q2, int B, float grad_scale, float* target_q,
                                 float* dq1, float* dq2, float* metrics, cudaStream_t stream) {
    launch_k(k_critic_loss, dim3(1), dim3(256), 0, stream, tq1, tq2, logpi_next, reward, not_done, log_alpha, discount,
                                         q1, q2, B, grad_scale, target_q, dq1, dq2, metrics);
    return check_launch("critic_loss");
}

extern "C" int curla_actor_loss(const float* log_pi, const float* q1, const float* q2,
                                const float* ls, int B, int A, const double* log_alpha,
                                float target_entropy, float grad_scale, float* dq1, float* dq2,
                                float* glogpi, double* g_log_alpha, float* metrics,
                                cudaStream_t stream) {
    launch_k(k_actor_loss, dim3(1), dim3(256), 0, stream, log_pi, q1, q2, ls, B, A, log_alpha, target_entropy,
                                        grad_scale, dq1, dq2, glogpi, g_log_alpha, metrics);
    return check_launch("actor_loss");
}

extern "C" int curla_add2(const float* a, const float* b, long long n, float* out,
                          cudaStream_t stream) {
    launch_k(k_add2, dim3(cdiv(n, 256)), dim3(256), 0, stream, a, b, n, out);
    return check_launch("add2");
}
