// Argument block shared by the two bf16 GEMM kernels (gemm.cu: mma.sync, any shape;
// gemm_tc.cu: tcgen05 + TMEM, M >= 128) and the segmented-index helper.
#pragma once
#include "common.cuh"

namespace curla {

struct GemmArgs {
    const bf16* A; long long lda;
    const bf16* B; long long ldb;
    void* C; long long ldc;
    int M, N, K;
    int n_store;              // columns [0, n_store) are written
    int out_bf16;             // 1: bf16 output, 0: fp32
    const float* bias;        // per column, optional
    int relu;
    const bf16* mask; long long ldmask;   // optional: out = mask[m][n] > 0 ? v : 0
    int k_per_split;          // multiple of BK; blockIdx.z selects the K range
    long long split_stride;   // elements between split outputs (fp32 partials)
    float alpha;
    // Segmented contiguous index (the channel-plane activation layout of the conv stack,
    // DESIGN.md section 3): index i of the chosen operand lives at (i / seg_len) * seg_stride +
    // i % seg_len.  seg_mask: 1 = A's contiguous index, 2 = B's, 4 = C's and the mask's column.
    int seg_len; long long seg_stride; int seg_mask; float seg_inv;
    // Batched problems (Q1 || Q2 of the critic: same shapes, different weights): blockIdx.z is
    // the batch index (split-K and batching are mutually exclusive); element strides.
    int batch; long long bsA, bsB, bsC, bsBias, bsMask;
    int vec_c;                // host: bf16 output (and mask) rows are 16-byte addressable -> staged epilogue
};

// i / seg_len through a float reciprocal: exact here because i is a multiple of 2 (columns)
// or 8 (chunks) below 2^22 with only a handful of segments, so (i + 0.5) / seg_len is never
// within float rounding of an integer.
__device__ __forceinline__ long long seg_off(int i, int seg_len, long long seg_stride, float inv_len, bool on) {
    if (!on) return i;
    const int s = __float2int_rd(((float)i + 0.5f) * inv_len);
    return (long long)s * seg_stride + (i - s * seg_len);
}

// tcgen05 path (gemm_tc.cu): returns 1 if it launched the problem, 0 if the shape is left to the
// mma.sync kernel, -1 on error.  grid_z = batch count or K splits, as in gemm.cu.
int gemm_tc_try_launch(const GemmArgs& p, int layout, int splits, cudaStream_t stream);

}  // namespace curla
