// K4/K5/K6: the 3x3 convolution stack of CNNEncoder as "shifted GEMMs".
//
// Replaces cuDNN conv fwd / dgrad / wgrad + ReLU launched by
//   encoder.py:77-90   CNNEncoder.forward_conv  (obs/255, conv s2 + 3x conv s1, ReLU)
//   curl_sac.py:367,417  loss.backward() through those layers
//
// Data layout: activations are NHWC bf16 with a FIXED row pitch
// (pitch = ceil(W/2), the space-to-depth width of the encoder input) and a fixed
// per-sample stride S = ceil(H/2)*pitch positions for every layer.  With that layout a
// 3x3 valid convolution is   out[P] = sum_t in[P + dy_t*pitch + dx_t] * W_t   over the
// FLATTENED position index P: each tap is the same matrix shifted by a constant number
// of rows, so one shared-memory slab of TM + 2*pitch + 2 rows feeds all nine taps (input
// read once).  Positions whose (y, x) fall outside the layer's valid output are written
// as exact zeros so they contribute nothing downstream (fc weights are zero there, dY is
// zero there).  conv1 (stride 2, Cin=9) runs on the space-to-depth input as a 2x2 conv
// with 4 taps x 48 channels (36 real).  dgrad is the same kernel with negated shifts and
// the transposed weight operand; wgrad reduces over positions with one warp per tap.
//
// This file holds wgrad (mma.sync m16n8k16 bf16 -> fp32, split over positions with a
// deterministic second stage); forward and dgrad are the tcgen05 kernels of conv_tc.cu.
#include "common.cuh"

namespace curla {

constexpr int TM = 128;   // positions per tile (4 warps x 32 rows)

// shared-memory row layouts: conflict-free for ldmatrix (8 rows x 16 B)
template <int CP> struct RowLayout;
template <> struct RowLayout<32> {          // 64-byte rows, 64B XOR swizzle
    static constexpr int kRowBytes = 64;
    __device__ static __forceinline__ uint32_t off(int row, int chunk) {
        return (uint32_t)row * 64u + (uint32_t)((chunk ^ ((row >> 1) & 3)) << 4);
    }
};
template <> struct RowLayout<48> {          // 96 data bytes in 112-byte rows
    static constexpr int kRowBytes = 112;
    __device__ static __forceinline__ uint32_t off(int row, int chunk) {
        return (uint32_t)row * 112u + (uint32_t)(chunk << 4);
    }
};

struct ConvGeom {
    int pitch;        // row pitch in positions
    int S;            // positions per sample (sample stride)
    int Hv, Wv;       // valid OUTPUT rows/cols of this launch (epilogue mask)
    int tiles_per_sample;
    int total_tiles;
    int slab_rows;    // TM + span
    int min_off;      // most negative tap shift (0 for fwd, -(2p+2) for dgrad)
};

struct TapOffsets { int off[9]; };

// global activations are channel planes [CP/8][S][8] per sample; the
// shared-memory rows keep all CP channels of a position together for ldmatrix.
template <int CP>
__device__ __forceinline__ void load_rows_async(uint32_t smem_base, const bf16* gsrc, long long plane,
                                                int nrows, int tid, int nthreads) {
    constexpr int CH = CP / 8;
#pragma unroll
    for (int c = 0; c < CH; ++c)
        for (int row = tid; row < nrows; row += nthreads)
            cp_async16(smem_base + RowLayout<CP>::off(row, c), gsrc + c * plane + (long long)row * 8, 16);
}

// ------------------------------------------------------------------ wgrad
// partial[cta][t][ci][co] = sum over this CTA's position chunks of in[P+off_t][ci]*dy[P][co]
// plus partial bias grad [32].  One warp per tap.
template <int CP, int NTAPS>
__global__ void __launch_bounds__(NTAPS * 32)
k_conv_wgrad(const bf16* __restrict__ in, long long in_sstride,
             const bf16* __restrict__ dy, long long dy_sstride,
             float* __restrict__ partial, ConvGeom g, TapOffsets taps) {
    extern __shared__ __align__(128) uint8_t smem[];
    using L = RowLayout<CP>;
    using LY = RowLayout<32>;
    constexpr int MT = CP / 16;
    constexpr int NTH = NTAPS * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t slab_bytes = ((uint32_t)g.slab_rows * L::kRowBytes + 127u) & ~127u;
    const uint32_t dy_bytes = TM * 64;
    const uint32_t stage_bytes = slab_bytes + dy_bytes;
    const uint32_t s_base = smem_u32(smem);

    auto prefetch = [&](int t, int buf) {
        const int b = t / g.tiles_per_sample;
        const int p0 = (t - b * g.tiles_per_sample) * TM;
        const uint32_t s = s_base + buf * stage_bytes;
        const long long plane = (long long)g.S * 8;
        load_rows_async<CP>(s, in + (long long)b * in_sstride + (long long)p0 * 8, plane, g.slab_rows, tid, NTH);
        load_rows_async<32>(s + slab_bytes, dy + (long long)b * dy_sstride + (long long)p0 * 8, plane, TM, tid, NTH);
    };
    int tile = blockIdx.x;
    if (tile < g.total_tiles) prefetch(tile, 0);
    cp_async_commit();

    float acc[MT][4][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
    float2 bsum = make_float2(0.f, 0.f);

    // A (trans): rows = k (positions), chunk = m chunk
    const int a_row = (lane >> 4) * 8 + (lane & 7), a_chk = (lane >> 3) & 1;
    // B (trans): rows = k, chunk = n chunk
    const int b_row = ((lane >> 3) & 1) * 8 + (lane & 7), b_chk = lane >> 4;
    const int toff = taps.off[warp];

    int buf = 0;
    for (; tile < g.total_tiles; tile += gridDim.x, buf ^= 1) {
        const int next = tile + gridDim.x;
        if (next < g.total_tiles) prefetch(next, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint32_t s_slab = s_base + buf * stage_bytes;
        const uint32_t s_dy = s_slab + slab_bytes;
#pragma unroll 2
        for (int ks = 0; ks < TM / 16; ++ks) {
            uint32_t a[MT][4], bfr[4][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                ldmatrix_x4_trans(s_slab + L::off(ks * 16 + toff + a_row, mt * 2 + a_chk),
                                  a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
#pragma unroll
            for (int np = 0; np < 2; ++np)
                ldmatrix_x4_trans(s_dy + LY::off(ks * 16 + b_row, np * 2 + b_chk),
                                  bfr[np * 2][0], bfr[np * 2][1], bfr[np * 2 + 1][0], bfr[np * 2 + 1][1]);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_bf16(acc[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], bfr[nt][0], bfr[nt][1]);
        }
        if (tid < 128) {   // bias grad: column sums of the dy tile
            const int cp2 = tid & 15, rg = tid >> 4;
#pragma unroll 4
            for (int r = rg; r < TM; r += 8) {
                const uint32_t o = LY::off(r, cp2 >> 2) + (cp2 & 3) * 4;
                const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(smem + (s_dy - s_base) + o));
                bsum.x += v.x; bsum.y += v.y;
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();

    // ---- write partials: [t][ci][co]
    float* ws = partial + (long long)blockIdx.x * (NTAPS * CP * 32 + 32);
    const int gq = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ci = mt * 16 + h * 8 + gq, co = nt * 8 + q * 2;
                float2 v = make_float2(acc[mt][nt][h * 2], acc[mt][nt][h * 2 + 1]);
                *reinterpret_cast<float2*>(ws + ((long long)warp * CP + ci) * 32 + co) = v;
            }
    __shared__ float s_b[8][32];
    if (tid < 128) { s_b[tid >> 4][(tid & 15) * 2] = bsum.x; s_b[tid >> 4][(tid & 15) * 2 + 1] = bsum.y; }
    __syncthreads();
    if (tid < 32) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += s_b[r][tid];
        ws[NTAPS * CP * 32 + tid] = s;
    }
}

// Deterministic second stage: sum CTA partials, write OIHW fp32 grads (+ bias grad).
//  s2d=0: dW[co][ci][t]      = scale * sum ws[cta][t][ci][co]           (Cin = 32)
//  s2d=1: dW[co][c][ky][kx]  = scale * sum ws[cta][by*2+bx][c*4+sy*2+sx][co]
__global__ void __launch_bounds__(256)
k_conv_wgrad_reduce_legacy(const float* __restrict__ partial, int nparts, int ntaps,
                    int CP, int Cin, int s2d, float scale,
                    float* __restrict__ dW, float* __restrict__ db) {
    // block = 32 consecutive elements x 8 partial-slices; fixed summation tree => deterministic
    __shared__ float red[8][33];
    const int per = ntaps * CP * 32 + 32;
    const int i = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (i < per)
        for (int c = threadIdx.y; c < nparts; c += 8) s += partial[(long long)c * per + i];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y != 0 || i >= per) return;
#pragma unroll
    for (int r = 1; r < 8; ++r) s += red[r][threadIdx.x];
    if (i >= ntaps * CP * 32) { db[i - ntaps * CP * 32] = s; return; }
    const int co = i & 31, ci = (i >> 5) % CP, t = (i >> 5) / CP;
    if (!s2d) {
        if (ci < Cin) dW[((long long)co * Cin + ci) * 9 + t] = s * scale;
    } else {
        const int c = ci >> 2, ky = 2 * (t >> 1) + ((ci >> 1) & 1), kx = 2 * (t & 1) + (ci & 1);
        if (c < Cin && ky < 3 && kx < 3) dW[((long long)co * Cin + c) * 9 + ky * 3 + kx] = s * scale;
    }
}

template <typename K>
static int set_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { set_last_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e)); return -1; }
    }
    return 0;
}

static ConvGeom make_geom(int B, int pitch, int S, int Hv, int Wv, int span, int min_off) {
    ConvGeom g;
    g.pitch = pitch; g.S = S; g.Hv = Hv; g.Wv = Wv;
    g.tiles_per_sample = cdiv((long long)Hv * pitch, TM);
    g.total_tiles = B * g.tiles_per_sample;
    g.slab_rows = TM + span;
    g.min_off = min_off;
    return g;
}

}  // namespace curla

using namespace curla;


// dW (OIHW fp32, Cin real channels), db[32].  Hv/Wv = valid dims of dY (this layer's OUTPUT).
int curla::legacy_conv_wgrad(const void* in, long long in_sstride, const void* dy,
                             long long dy_sstride, float* workspace, float* dW, float* db,
                             float scale, int B, int pitch, int S, int Hv, int Wv, int Cin,
                             int first_layer, cudaStream_t stream) {
    TapOffsets taps;
    const int grid_cap = sm_count() * 2;
    int grid;
    if (first_layer) {
        for (int t = 0; t < 4; ++t) taps.off[t] = (t >> 1) * pitch + (t & 1);
        ConvGeom g = make_geom(B, pitch, S, Hv, Wv, pitch + 1, 0);
        size_t smem = 2 * ((((size_t)g.slab_rows * 112 + 127) & ~127) + TM * 64);
        auto kern = k_conv_wgrad<48, 4>;
        if (set_smem(kern, smem)) return -1;
        grid = g.total_tiles < grid_cap ? g.total_tiles : grid_cap;
        kern<<<grid, 128, smem, stream>>>((const bf16*)in, in_sstride, (const bf16*)dy, dy_sstride,
                                          workspace, g, taps);
        if (check_launch("conv_wgrad")) return -1;
        const int per = 4 * 48 * 32 + 32;
        k_conv_wgrad_reduce_legacy<<<cdiv(per, 32), dim3(32, 8), 0, stream>>>(workspace, grid, 4, 48, Cin, 1, scale, dW, db);
    } else {
        for (int t = 0; t < 9; ++t) taps.off[t] = (t / 3) * pitch + (t % 3);
        ConvGeom g = make_geom(B, pitch, S, Hv, Wv, 2 * pitch + 2, 0);
        size_t smem = 2 * ((((size_t)g.slab_rows * 64 + 127) & ~127) + TM * 64);
        auto kern = k_conv_wgrad<32, 9>;
        if (set_smem(kern, smem)) return -1;
        grid = g.total_tiles < grid_cap ? g.total_tiles : grid_cap;
        kern<<<grid, 288, smem, stream>>>((const bf16*)in, in_sstride, (const bf16*)dy, dy_sstride,
                                          workspace, g, taps);
        if (check_launch("conv_wgrad")) return -1;
        const int per = 9 * 32 * 32 + 32;
        k_conv_wgrad_reduce_legacy<<<cdiv(per, 32), dim3(32, 8), 0, stream>>>(workspace, grid, 9, 32, Cin, 0, scale, dW, db);
    }
    return check_launch("conv_wgrad_reduce");
}
