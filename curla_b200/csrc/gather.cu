// K1: replay gather + random-crop kernels.
//
// Replaces, on device and without any host gather / PCIe upload:
//   utils.py:147-166   ReplayBuffer.sample_cpc  (self.obses[idxs] fancy-index gather,
//                                                torch.as_tensor(...).float())
//   augmentations.py:47-75  RandomCrop.training_augmentation (view_as_windows crop)
//   encoder.py:78      obs / 255.   (folded into the conv1 epilogue, see conv.cu)
//
// Two output formats:
//   * NCHW float32  -- what sample_cpc() hands to user code (public API parity)
//   * "s2d" bf16    -- what the conv stack consumes: space-to-depth by 2 so that the
//                      stride-2 3x3 conv1 becomes a stride-1 2x2 "shifted GEMM" with the
//                      same flattened-position indexing as conv2..4, stored as channel
//                      planes [CP/8][S][8] per sample.
//                      uint8 values are exact in bf16.
#include "common.cuh"
#include "tc.cuh"
#include "../../include/curla_b200.h"

#include <stdlib.h>
#include <string.h>

namespace curla {

// ------------------------------------------------------------------ NCHW f32
// one warp per output row (b, c, y); lanes stride over x.
__global__ void __launch_bounds__(256)
k_gather_crop_f32(const uint8_t* __restrict__ frames, int C, int Hf, int Wf,
                  const int64_t* __restrict__ idxs, const int64_t* __restrict__ h1,
                  const int64_t* __restrict__ w1, int B, int H, int W,
                  float* __restrict__ out) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)B * C * H;
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long stride = (long long)gridDim.x * (blockDim.x >> 5);
    for (; row < rows; row += stride) {
        const int y = (int)(row % H);
        const long long bc = row / H;
        const int c = (int)(bc % C);
        const int b = (int)(bc / C);
        const long long fi = idxs ? idxs[b] : b;
        const int oy = h1 ? (int)h1[b] : 0;
        const int ox = w1 ? (int)w1[b] : 0;
        const uint8_t* src = frames + ((fi * C + c) * Hf + (oy + y)) * (long long)Wf + ox;
        float* dst = out + row * W;
        for (int x = lane; x < W; x += 32) dst[x] = (float)__ldg(src + x);
    }
}

// ------------------------------------------------------------------ s2d bf16 (channel planes)
// One CTA per (sample b, output channel-chunk plane j).  Plane j holds s2d channels 8j..8j+7 =
// input channels 2j and 2j+1 (x 2x2 sub-positions), so the CTA reads the crop rows of exactly
// two stored channel planes -- each ONE contiguous run of H*Wf elements starting at row oy
// (whole stored rows: a crop window at an arbitrary byte offset touches every 32-byte sector of a
// row anyway) -- stages them in shared memory with 16-byte cp.async copies, and writes the whole
// output plane [Hs*Ws][8] bf16 as one contiguous 16-bytes-per-thread stream.  Long contiguous
// reads and writes keep DRAM pages open; the per-block-row version of this kernel moved the same
// bytes in 160-byte pieces and ran at half the bandwidth.
//   out[b][j][yb*Ws + xb][e]  with s2d channel ch = j*8 + e = c*4 + sy*2 + sx
//                             = frame[idx[b]][c][oy + 2yb+sy][ox + 2xb+sx]
// zero where 2yb+sy >= H, 2xb+sx >= W or c >= C.  Planes with 8j >= 4C are all zero and are
// never written (the caller's buffer is zero-initialised): grid.x covers only the real planes.
template <typename SrcT>
__global__ void __launch_bounds__(256)
k_gather_s2d(const SrcT* __restrict__ frames, int C, int Hf, int Wf,
             const int64_t* __restrict__ idxs, const int64_t* __restrict__ h1,
             const int64_t* __restrict__ w1, int H, int W, int Hs, int Ws, int CP,
             long long out_sample_stride, bf16* __restrict__ out) {
    pdl_grid_sync();
    extern __shared__ __align__(16) uint8_t s_raw[];
    const int b = blockIdx.y, j = blockIdx.x;
    const long long fi = idxs ? idxs[b] : b;
    const int oy = h1 ? (int)h1[b] : 0;
    const int ox = w1 ? (int)w1[b] : 0;
    constexpr int V = 16 / (int)sizeof(SrcT);      // source elements per 16-byte vector
    const int plane_elems = H * Wf;                // staged run of one input channel
    const int ppitch = (plane_elems + V - 1) / V * V + V;
    SrcT* s_pl = reinterpret_cast<SrcT*>(s_raw);   // [2][ppitch]
    const int c0 = 2 * j;
    const bool vec_ok = (Wf % V == 0) && ((reinterpret_cast<uintptr_t>(frames) & 15) == 0);
    for (int lp = 0; lp < 2; ++lp) {
        const int c = c0 + lp;
        if (c >= C) break;
        const SrcT* src = frames + ((fi * C + c) * Hf + oy) * (long long)Wf;
        if (vec_ok) {
            const uint32_t s0 = smem_u32(s_pl + lp * ppitch);
            for (int i = threadIdx.x; i < plane_elems / V; i += blockDim.x)
                cp_async16(s0 + (uint32_t)(i * 16), src + (long long)i * V, 16);
        } else {
            for (int i = threadIdx.x; i < plane_elems; i += blockDim.x) s_pl[lp * ppitch + i] = src[i];
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const long long S = (long long)Hs * Ws;
    bf16* oplane = out + b * out_sample_stride + (long long)j * S * 8;
    for (int p = threadIdx.x; p < Hs * Ws; p += blockDim.x) {
        const int yb = p / Ws, xb = p - yb * Ws;
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {          // word h: input channel c0 + (h >> 1), sub-row sy = h & 1, sx = 0, 1
            const int lp = h >> 1, sy = h & 1, y = 2 * yb + sy;
            float v0 = 0.f, v1 = 0.f;
            if (c0 + lp < C && y < H) {
                const SrcT* rp = s_pl + lp * ppitch + y * Wf + ox + 2 * xb;
                if (2 * xb < W) v0 = (float)rp[0];
                if (2 * xb + 1 < W) v1 = (float)rp[1];
            }
            w[h] = pack_bf16x2(v0, v1);
        }
        *reinterpret_cast<uint4*>(oplane + (long long)p * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ------------------------------------------------------------------ s2d bf16 from uint8 frames, bulk-staged
// The replay path of the update (uint8 frames, 16-byte aligned rows).  Same CTA = (sample, output
// plane) decomposition and output as k_gather_s2d<uint8_t>, with three differences:
//  * staging: ONE elected thread issues one cp.async.bulk (the TMA unit's linear copy) per stored
//    channel -- H*Wf contiguous bytes each -- completing on an mbarrier; no thread spends issue slots
//    on 16-byte cp.async and the CTA's other warps go straight to the barrier;
//  * conversion: a thread produces FOUR consecutive output positions (64 bytes) from four unaligned
//    8-byte runs, each read as three aligned 32-bit words + two funnel shifts, and turns bytes into
//    bf16 with the 2^23 magic-number add (PRMT + FADD per byte, full-rate pipes) instead of one
//    LDS.U8 + I2F (quarter-rate conversion pipe) per byte: the old kernel was issue-bound at 70 % SM
//    throughput and 33 % of DRAM peak (profiles/r02n_gemm_wgrad_gather_ncu_full.txt);
//  * up to three streams (obs / pos / next_obs: utils.py:151-158) and the gather of the batch's
//    action / reward / not_done rows (utils.py:163-165) share ONE launch (blockIdx.z = stream).
struct GatherSegs {
    const uint8_t* frames[3];
    const int64_t* h1[3];
    const int64_t* w1[3];
    bf16* out[3];
    // optional row gather, done by the CTAs (plane 0, stream 0): out_x[b] = x[idxs[b]]
    const float* actions; const float* rewards; const float* not_dones;
    float* out_act; float* out_rew; float* out_nd;
    int A;
};
template <typename T>
__device__ __forceinline__ T pick3(T const (&a)[3], int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); }

__global__ void __launch_bounds__(256)
k_gather_s2d_u8(const __grid_constant__ GatherSegs sg, const int64_t* __restrict__ idxs, int C, int Hf, int Wf,
                int H, int W, int Hs, int Ws, long long out_sample_stride) {
    extern __shared__ __align__(128) uint8_t s_raw[];      // [0,16) mbarrier | [2][ppitch] staged channels
    const int tid = threadIdx.x;
    const int j = blockIdx.x, b = blockIdx.y, s = blockIdx.z;
    const uint32_t bar = smem_u32(s_raw);
    const int plane_bytes = H * Wf;                         // multiple of 16 (launch precondition)
    const int ppitch = plane_bytes + 16;                    // slack: the last run's aligned words end past the plane
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    pdl_grid_sync();
    const int64_t* h1 = pick3(sg.h1, s);
    const int64_t* w1 = pick3(sg.w1, s);
    const long long fi = idxs ? idxs[b] : b;
    const int oy = h1 ? (int)h1[b] : 0;
    const int ox = w1 ? (int)w1[b] : 0;
    const int c0 = 2 * j;
    const int nch = C - c0 < 2 ? C - c0 : 2;
    if (tid == 0) {
        const uint8_t* frames = pick3(sg.frames, s);
        mbar_expect_tx(bar, (uint32_t)(nch * plane_bytes));
        for (int lp = 0; lp < nch; ++lp)
            bulk_g2s(smem_u32(s_raw + 16 + lp * ppitch), frames + ((fi * C + c0 + lp) * Hf + oy) * (long long)Wf,
                     (uint32_t)plane_bytes, bar);
    }
    if (j == 0 && s == 0 && sg.actions && tid < sg.A + 2) {
        if (tid < sg.A) sg.out_act[b * sg.A + tid] = sg.actions[fi * sg.A + tid];
        else if (tid == sg.A) sg.out_rew[b] = sg.rewards[fi];
        else sg.out_nd[b] = sg.not_dones[fi];
    }
    mbar_wait(bar, 0);
    const uint8_t* pl = s_raw + 16;
    const int QW = (Ws + 3) >> 2;                           // groups of four positions per output row
    const int nq = Hs * QW;
    bf16* oplane = pick3(sg.out, s) + b * out_sample_stride + (long long)j * Hs * Ws * 8;
    for (int q = tid; q < nq; q += 256) {
        const int yb = q / QW, xb0 = (q - yb * QW) * 4;
        const int nv = W - 2 * xb0;                         // valid bytes of the 8-byte runs of this group
        uint32_t w[4][4];                                   // [position][word h]: word h = channel c0 + (h >> 1), sub-row h & 1
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int lp = h >> 1, y = 2 * yb + (h & 1);
            uint32_t lo = 0u, hi = 0u;
            if (lp < nch && y < H && nv > 0) {
                const int o = lp * ppitch + y * Wf + ox + 2 * xb0;
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(pl + (o & ~3));
                const uint32_t w0 = wp[0], w1_ = wp[1], w2 = wp[2];
                const int sh = (o & 3) * 8;
                lo = __funnelshift_r(w0, w1_, sh);
                hi = __funnelshift_r(w1_, w2, sh);
                if (nv < 8) {                               // the crop's right edge: bytes at x >= W are zero
                    if (nv <= 4) { hi = 0u; if (nv < 4) lo &= (1u << (8 * nv)) - 1u; }
                    else hi &= (1u << (8 * (nv - 4))) - 1u;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t src = i < 2 ? lo : hi;
                const int k = (i & 1) * 2;
                // 0x4B0000bb = 2^23 + bb as a float; minus 2^23 = (float)bb exactly; exact in bf16 too
                const float v0 = __uint_as_float(__byte_perm(src, 0x4B000000u, 0x7650 + k)) - 8388608.f;
                const float v1 = __uint_as_float(__byte_perm(src, 0x4B000000u, 0x7651 + k)) - 8388608.f;
                w[i][h] = pack_bf16x2(v0, v1);
            }
        }
        bf16* op = oplane + ((long long)yb * Ws + xb0) * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (xb0 + i < Ws) *reinterpret_cast<uint4*>(op + i * 8) = make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
    }
}

// ------------------------------------------------------------------ small row gather
// out[b][k] = src[idx[b]][k]   (actions / rewards / not_dones: utils.py:163-165)
__global__ void k_gather_rows_f32(const float* __restrict__ src, const int64_t* __restrict__ idxs,
                                  int B, int K, float* __restrict__ out) {
    pdl_grid_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * K) {
        const int b = i / K, k = i % K;
        out[i] = src[idxs[b] * K + k];
    }
}

// ReplayBuffer.add (utils.py:120-128): one staged transition vec = [action[na] | reward | not_done]
// written to row `row` of the three fp32 ring arrays.
__global__ void k_scatter_transition(const float* __restrict__ vec, int na, long long row,
                                     float* __restrict__ actions, float* __restrict__ rewards,
                                     float* __restrict__ not_dones) {
    pdl_grid_sync();
    const int t = threadIdx.x;
    if (t < na) actions[row * na + t] = vec[t];
    else if (t == na) rewards[row] = vec[na];
    else if (t == na + 1) not_dones[row] = vec[na + 1];
}

}  // namespace curla

using namespace curla;

extern "C" int curla_gather_crop_f32(const uint8_t* frames, int C, int Hf, int Wf,
                                     const int64_t* idxs, const int64_t* h1, const int64_t* w1,
                                     int B, int H, int W, float* out, cudaStream_t stream) {
    CURLA_CHECK(B > 0 && H > 0 && W > 0 && H <= Hf && W <= Wf, "gather_crop_f32: bad shape");
    const long long rows = (long long)B * C * H;
    const int wpb = 8;
    int grid = (int)((rows + wpb - 1) / wpb);
    const int cap = sm_count() * 32;
    if (grid > cap) grid = cap;
    launch_k(k_gather_crop_f32, dim3(grid), dim3(wpb * 32), 0, stream, frames, C, Hf, Wf, idxs, h1, w1, B, H, W, out);
    return check_launch("gather_crop_f32");
}

template <typename SrcT>
static int launch_s2d(const SrcT* frames, int C, int Hf, int Wf, const int64_t* idxs,
                      const int64_t* h1, const int64_t* w1, int B, int H, int W, int CP,
                      long long out_sample_stride, bf16* out, cudaStream_t stream) {
    CURLA_CHECK(B > 0 && H <= Hf && W <= Wf && CP % 8 == 0 && CP >= 4 * C, "gather_s2d: bad shape");
    const int Hs = (H + 1) / 2, Ws = (W + 1) / 2;
    CURLA_CHECK(out_sample_stride >= (long long)Hs * Ws * CP, "gather_s2d: sample stride too small");
    const int real_planes = (4 * C + 7) / 8;        // planes beyond these are all zero: left untouched
    dim3 grid(real_planes, B);
    const int V = 16 / (int)sizeof(SrcT);
    const size_t smem = (size_t)2 * (((size_t)H * Wf + V - 1) / V * V + V) * sizeof(SrcT);
    CURLA_CHECK(smem <= 220 * 1024, "gather_s2d: a %dx%d window of %zu-byte elements does not fit shared memory", H, Wf, sizeof(SrcT));
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_gather_s2d<SrcT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        CURLA_CHECK(e == cudaSuccess, "gather_s2d: %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    }
    launch_k(k_gather_s2d<SrcT>, dim3(grid), dim3(256), smem, stream, frames, C, Hf, Wf, idxs, h1, w1, H, W, Hs, Ws,
                                                    CP, out_sample_stride, out);
    return check_launch("gather_s2d");
}

// Up to three streams of the same geometry (and the batch's action / reward / not_done rows) in one
// launch of the bulk-staged kernel; falls back to one k_gather_s2d<uint8_t> launch per stream (+ three
// row gathers) when a frame pointer or the stored row length is not 16-byte aligned.
extern "C" int curla_gather_crop_s2d_multi(const curla_gather_seg* segs, int nseg, const int64_t* idxs,
                                           int C, int Hf, int Wf, int B, int H, int W, int CP,
                                           long long out_sample_stride, const curla_gather_rows* rows,
                                           cudaStream_t stream) {
    CURLA_CHECK(nseg >= 1 && nseg <= 3, "gather_s2d_multi: 1..3 streams (got %d)", nseg);
    CURLA_CHECK(B > 0 && H <= Hf && W <= Wf && CP % 8 == 0 && CP >= 4 * C, "gather_s2d: bad shape");
    const int Hs = (H + 1) / 2, Ws = (W + 1) / 2;
    CURLA_CHECK(out_sample_stride >= (long long)Hs * Ws * CP, "gather_s2d: sample stride too small");
    if (rows) CURLA_CHECK(rows->action_dim >= 1 && rows->action_dim <= 62 && idxs, "gather_s2d_multi: bad row gather");
    bool bulk = (Wf % 16 == 0) && B <= 65535;
    for (int i = 0; i < nseg; ++i) bulk = bulk && segs[i].frames && segs[i].out && ((reinterpret_cast<uintptr_t>(segs[i].frames) & 15) == 0);
    const size_t smem = 16 + 2 * ((size_t)H * Wf + 16);
    if (smem > 220 * 1024) bulk = false;
    { const char* e = getenv("CURLA_GATHER_BULK"); if (e && e[0] == '0') bulk = false; }
    if (!bulk) {
        for (int i = 0; i < nseg; ++i)
            if (launch_s2d<uint8_t>((const uint8_t*)segs[i].frames, C, Hf, Wf, idxs, segs[i].h1, segs[i].w1, B, H, W, CP,
                                    out_sample_stride, (bf16*)segs[i].out, stream)) return -1;
        if (rows) {
            if (curla_gather_rows_f32(rows->actions, idxs, B, rows->action_dim, rows->out_actions, stream)) return -1;
            if (curla_gather_rows_f32(rows->rewards, idxs, B, 1, rows->out_rewards, stream)) return -1;
            if (curla_gather_rows_f32(rows->not_dones, idxs, B, 1, rows->out_not_dones, stream)) return -1;
        }
        return 0;
    }
    GatherSegs sg;
    memset(&sg, 0, sizeof(sg));
    for (int i = 0; i < nseg; ++i) {
        sg.frames[i] = (const uint8_t*)segs[i].frames; sg.h1[i] = segs[i].h1; sg.w1[i] = segs[i].w1; sg.out[i] = (bf16*)segs[i].out;
    }
    if (rows) {
        sg.actions = rows->actions; sg.rewards = rows->rewards; sg.not_dones = rows->not_dones;
        sg.out_act = rows->out_actions; sg.out_rew = rows->out_rewards; sg.out_nd = rows->out_not_dones; sg.A = rows->action_dim;
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_gather_s2d_u8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        CURLA_CHECK(e == cudaSuccess, "gather_s2d: %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    }
    const int real_planes = (4 * C + 7) / 8;        // planes beyond these are all zero: left untouched
    launch_k(k_gather_s2d_u8, dim3(real_planes, B, nseg), dim3(256), smem, stream, sg, idxs, C, Hf, Wf, H, W, Hs, Ws,
             out_sample_stride);
    return check_launch("gather_s2d");
}

extern "C" int curla_gather_crop_s2d(const uint8_t* frames, int C, int Hf, int Wf,
                                     const int64_t* idxs, const int64_t* h1, const int64_t* w1,
                                     int B, int H, int W, int CP, long long out_sample_stride,
                                     void* out, cudaStream_t stream) {
    const curla_gather_seg seg = {frames, h1, w1, out};
    return curla_gather_crop_s2d_multi(&seg, 1, idxs, C, Hf, Wf, B, H, W, CP, out_sample_stride, nullptr, stream);
}

extern "C" int curla_f32_to_s2d(const float* obs, int C, int H, int W, int B, int CP,
                                long long out_sample_stride, void* out, cudaStream_t stream) {
    return launch_s2d<float>(obs, C, H, W, nullptr, nullptr, nullptr, B, H, W, CP,
                             out_sample_stride, (bf16*)out, stream);
}

extern "C" int curla_gather_rows_f32(const float* src, const int64_t* idxs, int B, int K,
                                     float* out, cudaStream_t stream) {
    launch_k(k_gather_rows_f32, dim3(cdiv((long long)B * K, 256)), dim3(256), 0, stream, src, idxs, B, K, out);
    return check_launch("gather_rows_f32");
}

extern "C" int curla_scatter_transition(const float* vec, int na, long long row, float* actions,
                                        float* rewards, float* not_dones, cudaStream_t stream) {
    CURLA_CHECK(na >= 1 && na <= 62 && row >= 0, "scatter_transition: bad shape");
    launch_k(k_scatter_transition, dim3(1), dim3(64), 0, stream, vec, na, row, actions, rewards, not_dones);
    return check_launch("scatter_transition");
}

// ReplayBuffer.add in one call (utils.py:120-128): two async H2D copies of the pinned frame
// stacks into ring row `row`, one of the staged [action | reward | not_done] vector, and the
// scatter kernel.  All pointers except the *_pinned ones are device pointers.
extern "C" int curla_replay_add(const uint8_t* obs_pinned, const uint8_t* next_pinned, long long frame_bytes,
                                const float* vec_pinned, float* vec_dev, int na, long long row,
                                uint8_t* obses, uint8_t* next_obses, float* actions, float* rewards,
                                float* not_dones, cudaStream_t stream) {
    CURLA_CHECK(na >= 1 && na <= 62 && row >= 0 && frame_bytes > 0, "replay_add: bad shape");
    cudaError_t e = cudaMemcpyAsync(obses + row * frame_bytes, obs_pinned, (size_t)frame_bytes, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(next_obses + row * frame_bytes, next_pinned, (size_t)frame_bytes, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(vec_dev, vec_pinned, sizeof(float) * (na + 2), cudaMemcpyHostToDevice, stream);
    CURLA_CHECK(e == cudaSuccess, "replay_add: %s", cudaGetErrorString(e));
    return curla_scatter_transition(vec_dev, na, row, actions, rewards, not_dones, stream);
}
