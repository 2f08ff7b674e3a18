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
// One CTA per (sample b, s2d block-row yb).  The 2*C source rows are staged in shared memory
// with 16-byte loads of WHOLE stored rows (a crop window at an arbitrary byte offset touches
// every 32-byte sector of the row anyway), then every thread emits one 16-byte group of 8
// s2d channels; consecutive threads write consecutive positions of one channel plane, so the
// stores of a warp are 512 contiguous bytes.
//   out[b][j][yb*Ws + xb][e]  with s2d channel ch = j*8 + e = c*4 + sy*2 + sx
//                             = frame[idx[b]][c][oy + 2yb+sy][ox + 2xb+sx]
// zero where 2yb+sy >= H, 2xb+sx >= W or ch >= 4C.  Plane stride = S positions (S*8 elements).
template <typename SrcT>
__global__ void __launch_bounds__(128)
k_gather_s2d(const SrcT* __restrict__ frames, int C, int Hf, int Wf,
             const int64_t* __restrict__ idxs, const int64_t* __restrict__ h1,
             const int64_t* __restrict__ w1, int H, int W, int Hs, int Ws, int CP,
             long long out_sample_stride, bf16* __restrict__ out) {
    pdl_grid_sync();
    extern __shared__ __align__(16) uint8_t s_raw[];
    const int b = blockIdx.y, yb = blockIdx.x;
    const long long fi = idxs ? idxs[b] : b;
    const int oy = h1 ? (int)h1[b] : 0;
    const int ox = w1 ? (int)w1[b] : 0;
    const int nrows = 2 * C;                       // r = c*2 + sy
    constexpr int V = 16 / (int)sizeof(SrcT);      // source elements per 16-byte vector
    const int rowv = (Wf + V - 1) / V;             // vectors per staged row
    const int rpitch = rowv * V + V;               // elements; +V keeps rows apart in the banks
    SrcT* s_rows = reinterpret_cast<SrcT*>(s_raw);
    const bool vec_ok = (Wf % V == 0) && ((reinterpret_cast<uintptr_t>(frames) & 15) == 0);
    if (vec_ok) {
        for (int i = threadIdx.x; i < nrows * rowv; i += blockDim.x) {
            const int r = i / rowv, v = i - r * rowv;
            const int c = r >> 1, y = 2 * yb + (r & 1);
            uint4 val = make_uint4(0u, 0u, 0u, 0u);
            if (y < H)
                val = *reinterpret_cast<const uint4*>(frames + ((fi * C + c) * Hf + (oy + y)) * (long long)Wf + v * V);
            *reinterpret_cast<uint4*>(s_rows + r * rpitch + v * V) = val;
        }
    } else {
        for (int i = threadIdx.x; i < nrows * Wf; i += blockDim.x) {
            const int r = i / Wf, x = i - r * Wf;
            const int c = r >> 1, y = 2 * yb + (r & 1);
            s_rows[r * rpitch + x] = (y < H) ? frames[((fi * C + c) * Hf + (oy + y)) * (long long)Wf + x] : (SrcT)0;
        }
    }
    __syncthreads();
    const int chunks = CP / 8;
    const long long S = (long long)Hs * Ws;
    bf16* obase = out + b * out_sample_stride + (long long)yb * Ws * 8;
    for (int i = threadIdx.x; i < Ws * chunks; i += blockDim.x) {
        const int j = i / Ws, xb = i - j * Ws;
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {          // two s2d channels per 32-bit word: sx = 0, 1
            const int ch = j * 8 + h * 2;       // = c*4 + sy*2 (+ sx)
            const int c = ch >> 2, sy = (ch >> 1) & 1;
            float v0 = 0.f, v1 = 0.f;
            if (c < C) {
                const SrcT* rp = s_rows + (c * 2 + sy) * rpitch + ox + 2 * xb;
                if (2 * xb < W) v0 = (float)rp[0];
                if (2 * xb + 1 < W) v1 = (float)rp[1];
            }
            w[h] = pack_bf16x2(v0, v1);
        }
        *reinterpret_cast<uint4*>(obase + (long long)j * S * 8 + (long long)xb * 8) = make_uint4(w[0], w[1], w[2], w[3]);
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
    dim3 grid(Hs, B);
    const int V = 16 / (int)sizeof(SrcT);
    size_t smem = (size_t)2 * C * ((Wf + V - 1) / V * V + V) * sizeof(SrcT);
    launch_k(k_gather_s2d<SrcT>, dim3(grid), dim3(128), smem, stream, frames, C, Hf, Wf, idxs, h1, w1, H, W, Hs, Ws,
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
