// K13 fused Adam over flat parameter arenas, K14 EMA (soft target update) and the
// fp32 -> bf16 "shadow" packers that lay weights out for the tensor-core kernels.
//
// Replaces torch.optim.Adam.step (curl_sac.py:299-313,368,392,404,419-420) and
// utils.soft_update_params (utils.py:37-41; call sites curl_sac.py:443-445).
// All tensors of one optimizer are contiguous in the arena, so one launch covers what
// the reference does with ~4 tiny kernels per tensor.
#include "common.cuh"
#include <math.h>

namespace curla {

struct AdamHyper {
    double lr, b1, b2, eps;
};

// torch._single_tensor_adam maths:
//   m.lerp_(g, 1-b1); v = v*b2 + (1-b2)*g*g;
//   denom = sqrt(v)/sqrt(1-b2^t) + eps;  p += -(lr/(1-b1^t)) * (m/denom)
// Elements at index >= double_from receive the parameter step TWICE (sequentially, in
// fp32) -- this reproduces encoder_optimizer.step(); cpc_optimizer.step() acting on the
// critic encoder with identical optimizer states (curl_sac.py:419-420).
__global__ void __launch_bounds__(256)
k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
       float* __restrict__ v, long long n, long long double_from, AdamHyper h, int t_host,
       const int* __restrict__ t_dev) {
    pdl_grid_sync();
    __shared__ float s_c[2];
    if (threadIdx.x == 0) {
        const int t = t_dev ? *t_dev : t_host;
        const double bc1 = 1.0 - pow(h.b1, (double)t);
        const double bc2 = 1.0 - pow(h.b2, (double)t);
        s_c[0] = (float)(-(h.lr / bc1));
        s_c[1] = (float)sqrt(bc2);
    }
    __syncthreads();
    const float neg_step = s_c[0], bc2_sqrt = s_c[1];
    const float w1 = (float)(1.0 - h.b1), b2 = (float)h.b2, w2 = (float)(1.0 - h.b2);
    const float eps = (float)h.eps;
    auto step1 = [&](float& pi, float gi, float& mi, float& vi, bool twice) {
        mi = __fadd_rn(mi, __fmul_rn(w1, __fsub_rn(gi, mi)));
        vi = __fadd_rn(__fmul_rn(vi, b2), __fmul_rn(w2, __fmul_rn(gi, gi)));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), eps);
        const float u = __fmul_rn(neg_step, __fdiv_rn(mi, denom));
        pi = __fadd_rn(pi, u);
        if (twice) pi = __fadd_rn(pi, u);
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    const long long n4 = vec ? n >> 2 : 0;
    for (long long i = t0; i < n4; i += stride) {           // 16-byte accesses: 28 B/parameter at HBM speed
        float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        const long long e = i << 2;
        step1(pv.x, gv.x, mv.x, vv.x, e >= double_from);
        step1(pv.y, gv.y, mv.y, vv.y, e + 1 >= double_from);
        step1(pv.z, gv.z, mv.z, vv.z, e + 2 >= double_from);
        step1(pv.w, gv.w, mv.w, vv.w, e + 3 >= double_from);
        reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = (n4 << 2) + t0; i < n; i += stride) {
        float pi = p[i], mi = m[i], vi = v[i];
        step1(pi, g[i], mi, vi, i >= double_from);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

// log_alpha is a float64 0-dim tensor (curl_sac.py:292); state[0]=m, state[1]=v.
__global__ void k_adam_f64_scalar(double* p, const double* g, double* state, AdamHyper h,
                                  int t_host, const int* t_dev) {
    pdl_grid_sync();
    if (threadIdx.x || blockIdx.x) return;
    const int t = t_dev ? *t_dev : t_host;
    double m = state[0], v = state[1];
    const double gi = *g;
    m = m + (1.0 - h.b1) * (gi - m);
    v = v * h.b2 + (1.0 - h.b2) * gi * gi;
    const double bc1 = 1.0 - pow(h.b1, (double)t), bc2 = 1.0 - pow(h.b2, (double)t);
    const double denom = sqrt(v) / sqrt(bc2) + h.eps;
    *p = *p + (-(h.lr / bc1)) * (m / denom);
    state[0] = m; state[1] = v;
}

// target = tau*p + (1-tau)*target  (two roundings of the products, one of the sum, as
// the reference's three tensor ops); [0,split) uses tau_a, [split,n) tau_b.
__global__ void __launch_bounds__(256)
k_ema(float* __restrict__ tgt, const float* __restrict__ p, long long n, long long split,
      float tau_a, float omt_a, float tau_b, float omt_b) {
    pdl_grid_sync();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    auto one = [&](float t, float q, long long i) {
        const float tau = i < split ? tau_a : tau_b, omt = i < split ? omt_a : omt_b;
        return __fadd_rn(__fmul_rn(tau, q), __fmul_rn(omt, t));
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(tgt) | reinterpret_cast<uintptr_t>(p)) & 15) == 0;
    const long long n4 = vec ? n >> 2 : 0;
    for (long long i = t0; i < n4; i += stride) {
        float4 t = reinterpret_cast<float4*>(tgt)[i];
        const float4 q = reinterpret_cast<const float4*>(p)[i];
        const long long e = i << 2;
        t.x = one(t.x, q.x, e); t.y = one(t.y, q.y, e + 1); t.z = one(t.z, q.z, e + 2); t.w = one(t.w, q.w, e + 3);
        reinterpret_cast<float4*>(tgt)[i] = t;
    }
    for (long long i = (n4 << 2) + t0; i < n; i += stride) tgt[i] = one(tgt[i], p[i], i);
}

// ---------------------------------------------------------------- shadow packers
enum { PACK_ROWS = 0, PACK_CONV = 1, PACK_CONV1_S2D = 2, PACK_CONV96 = 3 };
struct PackSeg {
    long long src_off;   // floats into the fp32 arena
    long long dst_off;   // bf16 elements into the shadow arena
    int kind, rows, cols, rows_pad, cols_pad;   // conv: rows=F(out), cols=Cin, cols_pad=CP
};
#define CURLA_MAX_PACK 32
struct PackTable {
    int n;
    PackSeg seg[CURLA_MAX_PACK];
    int blk0[CURLA_MAX_PACK + 1];     // first CTA of each segment (CTAs are dealt out in proportion to the work)
};

__global__ void __launch_bounds__(256)
k_pack(const float* __restrict__ src_arena, bf16* __restrict__ dst_arena, PackTable tbl) {
    pdl_grid_sync();
    int si = 0;
    while (si + 1 < tbl.n && (int)blockIdx.x >= tbl.blk0[si + 1]) ++si;
    const PackSeg s = tbl.seg[si];
    const float* src = src_arena + s.src_off;
    bf16* dst = dst_arena + s.dst_off;
    const long long stride = (long long)(tbl.blk0[si + 1] - tbl.blk0[si]) * blockDim.x;
    long long i = (long long)((int)blockIdx.x - tbl.blk0[si]) * blockDim.x + threadIdx.x;
    if (s.kind == PACK_ROWS && (s.cols_pad & 7) == 0 && (s.dst_off & 7) == 0) {
        // 8 outputs (one 16-byte store) per thread; the big rows are the three fc weights (50 x Kfc -> 64 x Kfc)
        const int cpv = s.cols_pad >> 3;
        const long long n8 = (long long)s.rows_pad * cpv;
        const bool src4 = (s.cols & 3) == 0 && (s.src_off & 3) == 0;
        for (; i < n8; i += stride) {
            const int r = (int)(i / cpv), c = (int)(i - (long long)r * cpv) << 3;
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = 0.f;
            if (r < s.rows) {
                const float* sp = src + (long long)r * s.cols + c;
                if (src4 && c + 8 <= s.cols) {
                    const float4 a = *reinterpret_cast<const float4*>(sp), b = *reinterpret_cast<const float4*>(sp + 4);
                    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (c + k < s.cols) f[k] = sp[k];
                }
            }
            *reinterpret_cast<uint4*>(dst + (i << 3)) =
                make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
    } else if (s.kind == PACK_ROWS) {
        const long long n = (long long)s.rows_pad * s.cols_pad;
        for (; i < n; i += stride) {
            const int r = (int)(i / s.cols_pad), c = (int)(i % s.cols_pad);
            dst[i] = __float2bfloat16((r < s.rows && c < s.cols) ? src[(long long)r * s.cols + c] : 0.f);
        }
    } else if (s.kind == PACK_CONV) {
        // OIHW [F][Cin][3][3] -> [tap=ky*3+kx][F][cols_pad]
        const long long n = 9LL * s.rows * s.cols_pad;
        for (; i < n; i += stride) {
            const int ci = (int)(i % s.cols_pad);
            const int co = (int)((i / s.cols_pad) % s.rows);
            const int t = (int)(i / ((long long)s.cols_pad * s.rows));
            dst[i] = __float2bfloat16(ci < s.cols ? src[((long long)co * s.cols + ci) * 9 + t] : 0.f);
        }
    } else if (s.kind == PACK_CONV96) {
        // OIHW [32][32][3][3] -> the N = 96 operand image of conv_tc.cu: [dy][k chunk][n = dx*32 + co][8 k]
        const long long n = 9LL * 32 * 32;
        for (; i < n; i += stride) {
            const int k8 = (int)(i & 7), r = (int)(i >> 3);
            const int col = r % 96, dk = r / 96, kc = dk & 3, dy = dk >> 2, dx = col >> 5, co = col & 31;
            dst[i] = __float2bfloat16(src[((long long)co * 32 + kc * 8 + k8) * 9 + dy * 3 + dx]);
        }
    } else {
        // stride-2 3x3 conv1 as a 2x2 conv on the space-to-depth input:
        // [tap=by*2+bx][F][cp], cp = c*4 + sy*2 + sx  <->  (ky,kx) = (2by+sy, 2bx+sx)
        const long long n = 4LL * s.rows * s.cols_pad;
        for (; i < n; i += stride) {
            const int cp = (int)(i % s.cols_pad);
            const int co = (int)((i / s.cols_pad) % s.rows);
            const int t = (int)(i / ((long long)s.cols_pad * s.rows));
            const int c = cp >> 2, ky = 2 * (t >> 1) + ((cp >> 1) & 1), kx = 2 * (t & 1) + (cp & 1);
            float val = 0.f;
            if (c < s.cols && ky < 3 && kx < 3) val = src[((long long)co * s.cols + c) * 9 + ky * 3 + kx];
            dst[i] = __float2bfloat16(val);
        }
    }
}

}  // namespace curla

using namespace curla;

static int ew_grid(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int curla_adam_f32(float* p, const float* g, float* m, float* v, long long n,
                              long long double_from, double lr, double beta1, double beta2,
                              double eps, int t_host, const int* t_dev, cudaStream_t stream) {
    if (n <= 0) return 0;
    AdamHyper h{lr, beta1, beta2, eps};
    launch_k(k_adam, dim3(ew_grid(n)), dim3(256), 0, stream, p, g, m, v, n, double_from, h, t_host, t_dev);
    return check_launch("adam_f32");
}

extern "C" int curla_adam_f64_scalar(double* p, const double* g, double* state, double lr,
                                     double beta1, double beta2, double eps, int t_host,
                                     const int* t_dev, cudaStream_t stream) {
    AdamHyper h{lr, beta1, beta2, eps};
    launch_k(k_adam_f64_scalar, dim3(1), dim3(32), 0, stream, p, g, state, h, t_host, t_dev);
    return check_launch("adam_f64_scalar");
}

// Per-update scalars of a replayed CUDA graph: the four Adam step counters and the Philox offset of the policy
// noise live in 24 bytes of device memory.  They arrive as kernel PARAMETERS of this one-thread launch (copied at
// launch time), so the host may run any number of updates ahead of the device without a staging buffer to protect.
__global__ void k_set_state(int* __restrict__ state, int t0, int t1, int t2, int t3, unsigned long long off2) {
    pdl_grid_sync();
    if (threadIdx.x || blockIdx.x) return;
    state[0] = t0; state[1] = t1; state[2] = t2; state[3] = t3;
    *reinterpret_cast<unsigned long long*>(state + 4) = off2;
}
extern "C" int curla_set_dev_state(int* state, int t_critic, int t_actor, int t_alpha, int t_cpc,
                                   unsigned long long philox_offset, cudaStream_t stream) {
    launch_k(k_set_state, dim3(1), dim3(32), 0, stream, state, t_critic, t_actor, t_alpha, t_cpc, philox_offset);
    return check_launch("set_dev_state");
}

extern "C" int curla_ema_f32(float* target, const float* p, long long n, long long split,
                             double tau_a, double tau_b, cudaStream_t stream) {
    if (n <= 0) return 0;
    launch_k(k_ema, dim3(ew_grid(n)), dim3(256), 0, stream, target, p, n, split, (float)tau_a, (float)(1.0 - tau_a),
                                          (float)tau_b, (float)(1.0 - tau_b));
    return check_launch("ema_f32");
}

// segs: n x 7 int64 rows {src_off, dst_off, kind, rows, cols, rows_pad, cols_pad}
extern "C" int curla_pack_shadows(const float* src_arena, void* dst_arena, const long long* segs,
                                  int n, cudaStream_t stream) {
    for (int base = 0; base < n; base += CURLA_MAX_PACK) {
        PackTable tbl;
        tbl.n = (n - base < CURLA_MAX_PACK) ? n - base : CURLA_MAX_PACK;
        int nblk = 0;
        for (int i = 0; i < tbl.n; ++i) {
            const long long* r = segs + (long long)(base + i) * 7;
            PackSeg& s = tbl.seg[i];
            s.src_off = r[0]; s.dst_off = r[1]; s.kind = (int)r[2]; s.rows = (int)r[3];
            s.cols = (int)r[4]; s.rows_pad = (int)r[5]; s.cols_pad = (int)r[6];
            long long cnt = s.kind == PACK_ROWS ? (long long)s.rows_pad * s.cols_pad
                          : (s.kind == PACK_CONV || s.kind == PACK_CONV96 ? 9LL : 4LL) * s.rows * s.cols_pad;
            // the vectorised row path moves 8 elements per thread; a CTA gets ~2 items per thread
            const long long items = (s.kind == PACK_ROWS && (s.cols_pad & 7) == 0 && (s.dst_off & 7) == 0) ? cnt / 8 : cnt;
            long long b = (items + 511) / 512;
            if (b < 1) b = 1;
            if (b > 2048) b = 2048;
            tbl.blk0[i] = nblk;
            nblk += (int)b;
        }
        tbl.blk0[tbl.n] = nblk;
        launch_k(k_pack, dim3((unsigned)nblk), dim3(256), 0, stream, src_arena, (bf16*)dst_arena, tbl);
        if (check_launch("pack_shadows")) return -1;
    }
    return 0;
}
