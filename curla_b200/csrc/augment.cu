// K2/K3: the two device-side augmentations of the else-branch of ReplayBuffer.sample_cpc
// (utils.py:168-182), on float NCHW batches (B, 3*frame_stack, H, W) in [0, 255].
//
// Replaces
//   augmentations.py:106-136  ColorJiggle.training_augmentation
//        (x/255 -> kornia ColorJiggle(brightness 0, contrast .2, saturation .5, hue .5,
//         p .85, per-image parameters) -> x*255), one image = one RGB frame of the stack
//   augmentations.py:172-205  NoisyCover.training_augmentation
//        (index_fill_ of the cover rows per R/G/B plane, + N(0, 10^2), clamp [0,255])
//
// kornia is an unpinned pip dependency of the reference and is not available here, so the
// colour transform follows kornia's documented behaviour (PARITY UNPINNED, DESIGN.md
// section 1): contrast = clamp(x*f, 0, 1);
// saturation = HSV s *= f (clamp); hue = HSV h += f*2pi (mod 2pi); one op order per call;
// images not selected by the Bernoulli(p) mask are returned unchanged.  Per-image parameters
// are either injected (tests) or drawn from Philox(seed, image).
#include "common.cuh"

namespace curla {

struct Rgb { float r, g, b; };

__device__ __forceinline__ void rgb_to_hsv(Rgb c, float& h, float& s, float& v) {
    const float maxc = fmaxf(c.r, fmaxf(c.g, c.b)), minc = fminf(c.r, fminf(c.g, c.b));
    const float delta = maxc - minc;
    v = maxc;
    s = delta / (maxc + 1e-8f);
    const float d = delta == 0.f ? 1.f : delta;
    const float rc = maxc - c.r, gc = maxc - c.g, bc = maxc - c.b;
    float hh = (maxc == c.r) ? (bc - gc) : ((maxc == c.g) ? (2.f * d + rc - bc) : (4.f * d + gc - rc));
    hh = hh / d / 6.f;
    hh = hh - floorf(hh);                       // python-style % 1.0
    h = hh * 6.283185307179586f;
}
__device__ __forceinline__ Rgb hsv_to_rgb(float h, float s, float v) {
    const float h6 = h * (6.f / 6.283185307179586f);
    float hi = floorf(h6);
    hi = hi - 6.f * floorf(hi / 6.f);           // % 6
    float m6 = h6 - 6.f * floorf(h6 / 6.f);
    const float f = m6 - hi;
    const float p = v * (1.f - s), q = v * (1.f - f * s), t = v * (1.f - (1.f - f) * s);
    const int i = (int)hi;
    Rgb o;
    o.r = i == 0 ? v : (i == 1 ? q : (i == 2 ? p : (i == 3 ? p : (i == 4 ? t : v))));
    o.g = i == 0 ? t : (i == 1 ? v : (i == 2 ? v : (i == 3 ? q : (i == 4 ? p : p))));
    o.b = i == 0 ? p : (i == 1 ? p : (i == 2 ? t : (i == 3 ? v : (i == 4 ? v : q))));
    return o;
}

// params (optional, injected): [4][n_images] = contrast, saturation, hue, apply(0/1)
// order: 4 op codes packed 4 bits each, first op in the low nibble (0 brightness (no-op),
// 1 contrast, 2 saturation, 3 hue)
__global__ void __launch_bounds__(256)
k_color_jiggle(float* __restrict__ x, int n_images, int HW, const float* __restrict__ params,
               unsigned long long seed, unsigned long long offset, float c_rng, float s_rng, float h_rng,
               float p_apply, int order, float* __restrict__ params_out) {
    pdl_grid_sync();
    const int img = blockIdx.y;
    float cf, sf, hf;
    bool apply;
    if (params) {
        cf = params[img]; sf = params[n_images + img]; hf = params[2 * n_images + img];
        apply = params[3 * n_images + img] != 0.f;
    } else {
        const uint4 r = philox4x32(make_uint4((uint32_t)img, (uint32_t)offset, (uint32_t)(offset >> 32), 0x434Au),
                                   make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        cf = 1.f + c_rng * (2.f * u01(r.x) - 1.f);        // U[1-c, 1+c]
        sf = 1.f + s_rng * (2.f * u01(r.y) - 1.f);        // U[1-s, 1+s]
        hf = h_rng * (2.f * u01(r.z) - 1.f);              // U[-h, h]
        apply = u01(r.w) < p_apply;
    }
    if (params_out && blockIdx.x == 0 && threadIdx.x == 0) {
        params_out[img] = cf; params_out[n_images + img] = sf; params_out[2 * n_images + img] = hf;
        params_out[3 * n_images + img] = apply ? 1.f : 0.f;
    }
    if (!apply) return;                               // untouched image: x/255*255 is x up to 1 ulp
    float* base = x + (long long)img * 3 * HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        Rgb c;
        c.r = base[i] * (1.f / 255.f); c.g = base[HW + i] * (1.f / 255.f); c.b = base[2 * HW + i] * (1.f / 255.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int op = (order >> (4 * k)) & 15;
            if (op == 1) {
                c.r = fminf(fmaxf(c.r * cf, 0.f), 1.f); c.g = fminf(fmaxf(c.g * cf, 0.f), 1.f);
                c.b = fminf(fmaxf(c.b * cf, 0.f), 1.f);
            } else if (op == 2) {
                float h, s, v;
                rgb_to_hsv(c, h, s, v);
                c = hsv_to_rgb(h, fminf(fmaxf(s * sf, 0.f), 1.f), v);
            } else if (op == 3) {
                float h, s, v;
                rgb_to_hsv(c, h, s, v);
                float hh = fmodf(h + hf * 6.283185307179586f, 6.283185307179586f);
                if (hh < 0.f) hh += 6.283185307179586f;
                c = hsv_to_rgb(hh, s, v);
            }
        }
        base[i] = c.r * 255.f; base[HW + i] = c.g * 255.f; base[2 * HW + i] = c.b * 255.f;
    }
}

// x[img][ch][y][x]: rows y < top or y >= H - bottom of plane ch <- cover[ch]; + noise; clamp.
// noise_in (optional, injected, already scaled by std) has x's shape; otherwise
// std * N(0,1) from Philox(seed, element/4).
__global__ void __launch_bounds__(256)
k_noisy_cover(float* __restrict__ x, long long n4, int H, int W, int top, int bottom, float c0, float c1,
              float c2, float stdv, const float* __restrict__ noise_in, unsigned long long seed,
              unsigned long long offset) {
    pdl_grid_sync();
    const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= n4) return;
    const long long e = i4 * 4;                     // W % 4 == 0: the 4 elements share (img, ch, y)
    const int y = (int)((e / W) % H);
    const int ch = (int)((e / ((long long)W * H)) % 3);
    float4 v = *reinterpret_cast<const float4*>(x + e);
    if (y < top || y >= H - bottom) {
        const float c = ch == 0 ? c0 : (ch == 1 ? c1 : c2);
        v = make_float4(c, c, c, c);
    }
    float4 n;
    if (noise_in) {
        n = *reinterpret_cast<const float4*>(noise_in + e);
    } else {
        const uint4 r = philox4x32(make_uint4((uint32_t)i4, (uint32_t)(i4 >> 32), (uint32_t)offset, 0x4E43u),
                                   make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
        n = make_float4(g0.x * stdv, g0.y * stdv, g1.x * stdv, g1.y * stdv);
    }
    v.x = fminf(fmaxf(v.x + n.x, 0.f), 255.f); v.y = fminf(fmaxf(v.y + n.y, 0.f), 255.f);
    v.z = fminf(fmaxf(v.z + n.z, 0.f), 255.f); v.w = fminf(fmaxf(v.w + n.w, 0.f), 255.f);
    *reinterpret_cast<float4*>(x + e) = v;
}

}  // namespace curla

using namespace curla;

extern "C" int curla_color_jiggle(float* x, int n_images, int H, int W, const float* params,
                                  unsigned long long seed, unsigned long long offset, float contrast,
                                  float saturation, float hue, float p, int order, float* params_out,
                                  cudaStream_t stream) {
    CURLA_CHECK(n_images > 0 && H > 0 && W > 0, "color_jiggle: bad shape");
    const int HW = H * W;
    dim3 grid(cdiv(HW, 256 * 4) > 0 ? cdiv(HW, 256 * 4) : 1, n_images);
    launch_k(k_color_jiggle, dim3(grid), dim3(256), 0, stream, x, n_images, HW, params, seed, offset, contrast, saturation, hue, p,
                                            order, params_out);
    return check_launch("color_jiggle");
}

extern "C" int curla_noisy_cover(float* x, int n_images, int H, int W, int top, int bottom,
                                 const float* cover3_host, float stdv, const float* noise_in,
                                 unsigned long long seed, unsigned long long offset, cudaStream_t stream) {
    CURLA_CHECK(n_images > 0 && H > 0 && W > 0 && W % 4 == 0, "noisy_cover: W must be a multiple of 4");
    CURLA_CHECK(((uintptr_t)x & 15) == 0, "noisy_cover: x must be 16-byte aligned");
    const long long n4 = (long long)n_images * 3 * H * W / 4;
    launch_k(k_noisy_cover, dim3(cdiv(n4, 256)), dim3(256), 0, stream, x, n4, H, W, top, bottom, cover3_host[0], cover3_host[1],
                                                    cover3_host[2], stdv, noise_in, seed, offset);
    return check_launch("noisy_cover");
}
