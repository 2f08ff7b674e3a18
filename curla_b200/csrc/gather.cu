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
#include "../../include/curla_b200.h"

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

extern "C" int curla_gather_crop_s2d(const uint8_t* frames, int C, int Hf, int Wf,
                                     const int64_t* idxs, const int64_t* h1, const int64_t* w1,
                                     int B, int H, int W, int CP, long long out_sample_stride,
                                     void* out, cudaStream_t stream) {
    return launch_s2d<uint8_t>(frames, C, Hf, Wf, idxs, h1, w1, B, H, W, CP, out_sample_stride,
                               (bf16*)out, stream);
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
