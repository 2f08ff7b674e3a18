// K6 on the 5th-generation tensor cores: conv weight gradient.
//
// Replaces cuDNN wgrad launched by loss.backward() (curl_sac.py:367,417) through
// encoder.py:81-87:   dW[ky][kx][ci][co] = sum_P in[P + ky*pitch + kx][ci] * dy[P][co],
//                     db[co] = sum_P dy[P][co]            (P over all positions of the batch)
//
// GEMM shape.  The reduction runs over positions, so positions are the K dimension and both
// operands are "MN-major" (8 channels contiguous, K rows 16 B apart): exactly the channel-plane
// layout the activations already have.  tcgen05.mma needs M >= 64 while a
// tap only has 32 input channels, and the M chunks (8 channels each) of one descriptor must be
// a CONSTANT byte stride apart.  Trick: shared memory holds the input one IMAGE ROW per slot,
//     slot r = [plane 0 | plane 1 | ... | plane CH-1 | ones-plane]   each `pitch` rows x 16 B,
// slots of consecutive image rows back to back.  With chunk stride = one plane, chunk index
// i = g*(CH+1) + c  addresses plane c of image row y+g: ONE M=128 descriptor covers all the
// vertical taps g of all input-channel chunks c (plus a ones-plane per row whose D rows are
// the bias gradient for free), with no data replication.
//
// The horizontal taps ride in the N dimension.  An SS-mode tcgen05.mma streams its whole A tile
// (128 rows x 32 B) from shared memory whatever N is, and at N = 32 that read -- not the tensor
// pipe -- sets the pace (measured: ~64 clk per M128 N32 K16 instruction, 4x the N/256 floor).
// So dy is staged NDX times, copy dx shifted DOWN by dx rows (cp.async.bulk with a +16*dx byte
// destination offset; the second and third copies are L2 hits), and ONE MMA per K window
// computes all horizontal taps at once with N = 32*NDX:
//   D[(g, c, e)][(dx, co)] += sum_{k'} A[(g,c,e)][k'] * dy[k' - dx][co]
// K windows are 16 positions inside one image row; every staged dy plane has zero rows in front
// of (k' < dx) and behind (k' >= pitch + dx) the copied row, so out-of-row taps multiply zeros.
//
// What paces it (round 2, gpurun r04a-r04i; the experiments are in git history at 3b521f2): a 3x3 layer needs 280 clk
// of MMAs per image row (5 K windows x 56 clk) and runs at ~700.  Neither the bytes pulled from L2 (taps on the A side,
// CURLA_WG_COPIES=1: 19 -> 10 KB per row, slower: 15 N = 32 MMAs per row), nor the pipeline shape (a ring of image rows
// with a CTA walking a contiguous range of rows so that halo rows are fetched once and HBM never idles: 0.51 vs 0.37 ms)
// nor the number of requests (tensor-map boxes) moves it; per-role cycle counters of the ring variant show the MMA
// thread itself blocked ~140 clk per N = 96 MMA (profiles/tools/mma_rate_mn.cu: 56 clk with nothing else running): the
// bulk copies landing in shared memory (19 KB per row, in 1 KB pieces at 16-byte offsets) and the operand reads (35 KB
// per row) share the shared-memory port.  What did help: more threads issuing the copies (one thread issues a
// cp.async.bulk every 50-110 clk whatever its size, profiles/tools/stream_rate.cu) -- the four epilogue warps, idle
// until the end, issue the dy copies: 0.405 -> 0.368 ms per update.
//
// One persistent CTA per SM: producer warp (cp.async.bulk, one copy per plane per image row),
// MMA warp (R rows x kwin windows, one N = 32*NDX MMA each per stage), accumulators stay in TMEM for
// the CTA's whole share of the batch; 4 epilogue warps write one fp32 partial per CTA and a
// deterministic second-stage kernel sums the partials into OIHW (run-to-run reproducible).
#include "common.cuh"
#include "tc.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <vector>

namespace curla {

struct WgGeom {
    int pitch, S, Hv, Wv, B;
    int R;                 // dy image rows per stage
    int runs_per_sample, total_runs;
    int kwin;              // 16-position K windows per image row = ceil((Wv + NDX - 1) / 16)
    int dyr;               // rows of a staged dy plane (>= pitch + NDX - 1, >= kwin*16), zero outside the copy
    int nst;               // pipeline stages (2..4): more, shorter stages keep more of the fill latency covered
    int tmap;              // 1: operands staged by tensor-map TMA boxes (one per image row / per dy copy, all channel
                           //    planes at once); 0: one linear cp.async.bulk per plane per image row
    int pss;               // shared-memory stride of one plane of one image row (bytes): pitch*16, rounded up to 128 with tmap
    int dy1;               // 1: dy comes from L2 ONCE; the shifted copies of the N = 32*NDX trick are made inside shared memory
                           //    by the four otherwise idle epilogue warps (L2 -> SM traffic of a 3x3 layer: 96 -> 52 KB per run)
    int nh;                // helper producers (0..4): epilogue warps that issue the dy copies while the producer warp issues the
                           //    input rows -- one thread issues a cp.async.bulk every ~50-110 clk whatever its size
                           //    (profiles/tools/stream_rate.cu), and a stage is 88 copies of 1 KB
    int nc;                // staged copies of dy (1..NDX).  nc < NDX: the remaining horizontal taps come from the A side --
                           //    a K window of A starts s = nc, 2 nc, .. rows later (the input rows are contiguous in K, so
                           //    any 16-byte start is a legal descriptor), one more MMA of N = 32 * (taps covered) each
};

// Tensor-map staging.  The producer of the first version issued one 1088-byte linear bulk copy per channel plane
// per image row: 88 copies per 5-row stage, and the fill of a stage -- not the MMAs (1400 clk) nor the bytes (L2
// delivered half its cap) -- set the kernel's pace at ~3800 clk per stage (the TMA unit spends a fixed cost per
// request).  A 4-D tensor map over the activation buffer, (x in 8-byte units, y, plane, sample), lets ONE request
// bring all planes of an image row (A: 4-6 planes, ~4.5 KB) or one shifted copy of a dy row (4 planes, 5 KB): 22
// requests per stage.  The shift and the zero rows around a dy copy come from the box itself: it starts at
// x = -dx and is dyr positions long, and the TMA unit zero-fills what lies outside [0, pitch).
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

constexpr int kWgThreads = 6 * 32;      // 4 epilogue warps + MMA + producer
// D=f32, A=B=bf16, A and B MN-major (bits 15, 16), M=128, N = n
constexpr uint32_t idesc_mn(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// CP: input channels (32, or 48 for the space-to-depth conv1 input); GR x NDX taps (3x3 or 2x2)
template <int CP, int GR, int NDX>
__global__ void __launch_bounds__(kWgThreads, 1)
k_conv_wgrad_tc(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmDy,
                const bf16* __restrict__ in, long long in_sstride, const bf16* __restrict__ dy,
                long long dy_sstride, float* __restrict__ partial, WgGeom g) {
    constexpr int CH = CP / 8, CPL = CH + 1;          // planes per slot incl. the ones-plane
    static_assert(GR * CPL <= 16, "M = 128 holds 16 chunks");
    constexpr uint32_t TMEM_COLS = 128;
    constexpr int NTAPS = GR * NDX;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s_base = smem_u32(smem);
    // header: full[4] @0, empty[4] @32, done @64, tmem ptr @72, ready[4] @96
    const uint32_t s_full = s_base, s_empty = s_base + 32, s_done = s_base + 64, s_tptr = s_base + 72, s_ready = s_base + 96;
    const int nst = g.nst;
    const uint32_t PS = (uint32_t)g.pss;                               // one plane of one image row (shared memory)
    const uint32_t PG = (uint32_t)g.pitch * 16u;                       // ... in global memory
    const uint32_t slot_bytes = CPL * PS;
    const int slots = g.R + GR - 1;
    const uint32_t a_bytes = (uint32_t)slots * slot_bytes;
    const uint32_t DYB = (uint32_t)g.dyr * 16u;                         // one staged dy plane
    const int NC = g.nc;                                                // staged dy copies per image row
    const uint32_t dy_bytes = (uint32_t)(g.R * 4 * NC) * DYB;
    const uint32_t stage_bytes = (a_bytes + dy_bytes + 127u) & ~127u;
    const uint32_t s_stage0 = s_base + 128;
    const long long plane = (long long)g.S * 8;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(s_full + 8 * i, 1); mbar_init(s_empty + 8 * i, 1); mbar_init(s_ready + 8 * i, 4); }
        mbar_init(s_done, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tptr),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // every byte the tensor core may touch must be finite (junk rows are multiplied by exact
    // zeros of dy, and 0 * NaN would poison a useful accumulator row): clear both stages once
    for (uint32_t i = tid; i < ((uint32_t)nst * stage_bytes) / 16; i += kWgThreads)
        reinterpret_cast<uint4*>(smem + 128)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // constant parts of both stages: ones-planes (bf16 1.0); the tail of every dy plane stays zero
    for (int st = 0; st < nst; ++st) {
        uint8_t* sb = smem + 128 + (size_t)st * stage_bytes;
        const int prow = (int)(PS / 16);
        const int ones16 = slots * prow;                                 // 16-byte rows of ones
        for (int i = tid; i < ones16; i += kWgThreads) {
            const int r = i / prow, row = i - r * prow;
            *reinterpret_cast<uint4*>(sb + (size_t)r * slot_bytes + (size_t)CH * PS + (size_t)row * 16) =
                make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        }
    }
    pdl_grid_sync();     // the prologue above touched no global memory
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + 72);
    const int my_runs = g.total_runs > (int)blockIdx.x ? (g.total_runs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 5) {
        // ================= producer
        uint32_t stage = 0, phase = 0;
        for (int run = blockIdx.x; run < g.total_runs; run += gridDim.x) {
            const int b = run / g.runs_per_sample;
            const int y0 = (run - b * g.runs_per_sample) * g.R;
            const int rv = (g.Hv - y0) < g.R ? (g.Hv - y0) : g.R;          // valid dy rows of this run
            const uint32_t sa = s_stage0 + stage * stage_bytes, sd = sa + a_bytes;
            mbar_wait(s_empty + 8 * stage, phase ^ 1);
            if (elect_one()) {
                const uint32_t bar = s_full + 8 * stage;
                const int arows = rv + GR - 1;
                const int ndx = g.dy1 ? 1 : NC;              // copies of dy that come from global memory
                if (g.tmap) {
                    mbar_expect_tx(bar, (uint32_t)arows * CH * PS + (uint32_t)(rv * ndx) * 4u * DYB);
                    for (int r = 0; r < arows; ++r)
                        tma_load_4d(sa + (uint32_t)r * slot_bytes, &tmIn, 0, y0 + r, 0, b, bar);
                    for (int r = 0; r < rv; ++r)
#pragma unroll
                        for (int dx = 0; dx < NDX; ++dx)
                            if (dx < ndx) tma_load_4d(sd + (uint32_t)((r * NC + dx) * 4) * DYB, &tmDy, -2 * dx, y0 + r, 0, b, bar);
                } else {
                mbar_expect_tx(bar, (uint32_t)(arows * CH + rv * 4 * ndx) * PG);
                const bf16* src_a = in + (long long)b * in_sstride + (long long)y0 * g.pitch * 8;
                for (int r = 0; r < arows; ++r)
#pragma unroll
                    for (int c = 0; c < CH; ++c)
                        bulk_g2s(sa + (uint32_t)r * slot_bytes + (uint32_t)c * PS,
                                 src_a + c * plane + (long long)r * g.pitch * 8, PG, bar);
                const bf16* src_d = dy + (long long)b * dy_sstride + (long long)y0 * g.pitch * 8;
                if (g.nh == 0)
                for (int r = 0; r < rv; ++r)
#pragma unroll
                    for (int dx = 0; dx < NDX; ++dx)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (dx < ndx)
                                bulk_g2s(sd + (uint32_t)((r * NC + dx) * 4 + c) * DYB + (uint32_t)dx * 16u,
                                         src_d + c * plane + (long long)r * g.pitch * 8, PG, bar);
                }
            }
            __syncwarp();
            if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 4) {
        // ================= MMA issuer
        uint32_t stage = 0, phase = 0, accum = 0;
        const uint64_t a_hi = make_desc(0, 128, PS), b_hi = make_desc(0, 128, DYB);   // (lbo = K groups, sbo = chunks)
        // MMA groups of one K window: ng = ceil(NDX / NC) instructions, group j covers min(NC, NDX - j NC) taps
        const int ng = (NDX + NC - 1) / NC;
        const uint32_t id0 = idesc_mn(32u * (uint32_t)(NDX < NC ? NDX : NC));
        const uint32_t id1 = idesc_mn(32u * (uint32_t)(NDX - NC < NC ? (NDX - NC > 0 ? NDX - NC : 1) : NC));
        const uint32_t id2 = idesc_mn(32u * (uint32_t)(NDX - 2 * NC > 0 ? NDX - 2 * NC : 1));
        for (int run = blockIdx.x; run < g.total_runs; run += gridDim.x) {
            const int b = run / g.runs_per_sample;
            const int y0 = (run - b * g.runs_per_sample) * g.R;
            const int rv = (g.Hv - y0) < g.R ? (g.Hv - y0) : g.R;
            mbar_wait((g.dy1 ? s_ready : s_full) + 8 * stage, phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sa16 = (s_stage0 + stage * stage_bytes) >> 4, sd16 = sa16 + (a_bytes >> 4);
                for (int r = 0; r < rv; ++r) {
                    uint32_t a16 = sa16 + (uint32_t)r * (slot_bytes >> 4);
                    uint32_t b16 = sd16 + (uint32_t)(r * 4 * NC) * (DYB >> 4);
                    for (int kw = 0; kw < g.kwin; ++kw, a16 += 16u, b16 += 16u) {
                        const uint64_t bd = b_hi | (uint64_t)(b16 & 0x3FFFu);
                        // group j: taps dx = j NC .. : the A window starts j NC rows later, against dy copies 0 .. cnt-1
                        umma_bf16_rt(tmem_base, a_hi | (uint64_t)(a16 & 0x3FFFu), bd, id0, accum);
                        if (ng > 1) umma_bf16_rt(tmem_base + (uint32_t)(32 * NC), a_hi | (uint64_t)((a16 + (uint32_t)NC) & 0x3FFFu), bd, id1, accum);
                        if (ng > 2) umma_bf16_rt(tmem_base + (uint32_t)(64 * NC), a_hi | (uint64_t)((a16 + 2u * (uint32_t)NC) & 0x3FFFu), bd, id2, accum);
                        accum = 1;
                    }
                }
                umma_commit(s_empty + 8 * stage);
            }
            accum = 1;
            __syncwarp();
            if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(s_done);
        __syncwarp();
    } else {
        // ================= helper producers: this warp issues the dy copies of rows r = warp, warp + nh, ..
        // (the producer warp has armed the stage's barrier with the byte count of ALL copies; a copy that
        // completes before the expect_tx only drives the pending count negative for a moment)
        if (warp < g.nh) {
            uint32_t stage = 0, phase = 0;
            for (int run = blockIdx.x; run < g.total_runs; run += gridDim.x) {
                const int b = run / g.runs_per_sample;
                const int y0 = (run - b * g.runs_per_sample) * g.R;
                const int rv = (g.Hv - y0) < g.R ? (g.Hv - y0) : g.R;
                const uint32_t sd = s_stage0 + stage * stage_bytes + a_bytes;
                mbar_wait(s_empty + 8 * stage, phase ^ 1);
                if (elect_one()) {
                    const uint32_t bar = s_full + 8 * stage;
                    const bf16* src_d = dy + (long long)b * dy_sstride + (long long)y0 * g.pitch * 8;
                    for (int r = warp; r < rv; r += g.nh)
#pragma unroll
                        for (int dx = 0; dx < NDX; ++dx)
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                if (dx < NC)
                                    bulk_g2s(sd + (uint32_t)((r * NC + dx) * 4 + c) * DYB + (uint32_t)dx * 16u,
                                             src_d + c * plane + (long long)r * g.pitch * 8, PG, bar);
                }
                __syncwarp();
                if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1; }
            }
        }
        // ================= while the runs stream through: the shifted dy copies, made inside shared memory
        if (g.dy1) {
            uint32_t stage = 0, phase = 0;
            for (int run = blockIdx.x; run < g.total_runs; run += gridDim.x) {
                const int b = run / g.runs_per_sample;
                const int y0 = (run - b * g.runs_per_sample) * g.R;
                const int rv = (g.Hv - y0) < g.R ? (g.Hv - y0) : g.R;
                const uint32_t sd = s_stage0 + stage * stage_bytes + a_bytes;
                mbar_wait(s_full + 8 * stage, phase);
                const int per = 4 * g.pitch;                                  // 16-byte rows of one dy image row (4 planes)
                for (int i = tid; i < rv * per; i += 4 * 32) {
                    const int r = i / per, rem = i - r * per, c = rem / g.pitch, k = rem - c * g.pitch;
                    const uint32_t src = sd + (uint32_t)((r * NC) * 4 + c) * DYB + (uint32_t)k * 16u;
                    uint32_t v0, v1, v2, v3;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(src) : "memory");
#pragma unroll
                    for (int dx = 1; dx < NDX; ++dx)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(src + (uint32_t)(dx * 4) * DYB + (uint32_t)dx * 16u),
                                     "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
                }
                fence_proxy_async();            // generic-proxy writes -> visible to the tensor core's operand reads
                __syncwarp();
                if (lane == 0) mbar_arrive(s_ready + 8 * stage);
                if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1; }
            }
        }
        // ================= epilogue (once): TMEM -> this CTA's fp32 partial [tap][ci][co] + bias[co]
        float* ws = partial + (long long)blockIdx.x * (NTAPS * CP * 32 + 32);
        const int m = warp * 32 + lane;                     // D row = chunk*8 + e
        const int chunk = m >> 3, e = m & 7;
        const int gg = chunk / CPL, c = chunk - gg * CPL;
        uint32_t r[32];
        if (my_runs > 0) {
            mbar_wait(s_done, 0);
            tc_fence_after();
        }
#pragma unroll
        for (int dx = 0; dx < NDX; ++dx) {
            if (my_runs > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(dx * 32), r);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = 0u;
            }
            if (gg < GR && c < CH) {
                float* dst = ws + ((long long)(gg * NDX + dx) * CP + c * 8 + e) * 32;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<uint4*>(dst + i * 4) = make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
            } else if (gg == 0 && c == CH && e == 0 && dx == 0) {
                float* dst = ws + (long long)NTAPS * CP * 32;          // ones-plane row: sum_P dy[P][co]
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<uint4*>(dst + i * 4) = make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                     : "memory");
    }
}

// Deterministic second stage: sum CTA partials, write OIHW fp32 grads (+ bias grad).
//  s2d=0: dW[co][ci][t]      = scale * sum ws[cta][t][ci][co]           (Cin = 32)
//  s2d=1: dW[co][c][ky][kx]  = scale * sum ws[cta][by*2+bx][c*4+sy*2+sx][co]
__global__ void __launch_bounds__(256)
k_conv_wgrad_reduce(const float* __restrict__ partial, int nparts, int ntaps,
                    int CP, int Cin, int s2d, float scale,
                    float* __restrict__ dW, float* __restrict__ db) {
    pdl_grid_sync();
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

// Same reduction for up to 4 layers in one launch (blockIdx.y = layer): the engine defers the
// second stage of every layer of a backward pass to a single kernel after the last dgrad.
struct WgReduceJob { const float* partial; int nparts, ntaps, CP, Cin, s2d; float scale; float* dW; float* db; };
struct WgReduceJobs { WgReduceJob j[4]; };
__global__ void __launch_bounds__(256)
k_conv_wgrad_reduce_multi(WgReduceJobs jobs) {
    pdl_grid_sync();
    __shared__ float red[8][33];
    const WgReduceJob& J = jobs.j[blockIdx.y];
    const int per = J.ntaps * J.CP * 32 + 32;
    const int i = blockIdx.x * 32 + threadIdx.x;
    if (blockIdx.x * 32 >= per) return;
    float s = 0.f;
    if (i < per)
        for (int c = threadIdx.y; c < J.nparts; c += 8) s += J.partial[(long long)c * per + i];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y != 0 || i >= per) return;
#pragma unroll
    for (int r = 1; r < 8; ++r) s += red[r][threadIdx.x];
    if (i >= J.ntaps * J.CP * 32) { J.db[i - J.ntaps * J.CP * 32] = s; return; }
    const int co = i & 31, ci = (i >> 5) % J.CP, t = (i >> 5) / J.CP;
    if (!J.s2d) {
        if (ci < J.Cin) J.dW[((long long)co * J.Cin + ci) * 9 + t] = s * J.scale;
    } else {
        const int c = ci >> 2, ky = 2 * (t >> 1) + ((ci >> 1) & 1), kx = 2 * (t & 1) + (ci & 1);
        if (c < J.Cin && ky < 3 && kx < 3) J.dW[((long long)co * J.Cin + c) * 9 + ky * 3 + kx] = s * J.scale;
    }
}

// ---- tensor maps (host): (x in 8-byte units, image row, channel plane, sample) over a channel-plane buffer
typedef CUresult (*WgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WgEncodeFn wg_encode_fn() {
    static WgEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (WgEncodeFn)f;
        cudaGetLastError();
    }
    return fn;
}
// false: no map (the caller falls back to linear bulk copies)
static bool wg_map(CUtensorMap* out, const void* ptr, int pitch, int rows, int planes, int B, long long plane_elems,
                   long long sstride_elems, int box_x8) {
    struct Key { const void* p; long long a, b; int c[5]; };
    static thread_local std::vector<std::pair<Key, CUtensorMap>> cache;
    Key k;
    memset(&k, 0, sizeof(k));
    k.p = ptr; k.a = plane_elems; k.b = sstride_elems; k.c[0] = pitch; k.c[1] = rows; k.c[2] = planes; k.c[3] = B; k.c[4] = box_x8;
    for (auto& e : cache) if (!memcmp(&e.first, &k, sizeof(k))) { *out = e.second; return true; }
    WgEncodeFn enc = wg_encode_fn();
    if (!enc || box_x8 > 256 || (reinterpret_cast<uintptr_t>(ptr) & 15)) return false;
    const cuuint64_t gd[4] = {(cuuint64_t)2 * pitch, (cuuint64_t)rows, (cuuint64_t)planes, (cuuint64_t)B};
    const cuuint64_t gs[3] = {(cuuint64_t)pitch * 16, (cuuint64_t)plane_elems * 2, (cuuint64_t)sstride_elems * 2};
    const cuuint32_t bx[4] = {(cuuint32_t)box_x8, 1, (cuuint32_t)planes, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMap tm;
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(ptr), gd, gs, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    if (cache.size() >= 64) cache.clear();            // maps are handed out by value: nothing dangles
    cache.emplace_back(k, tm);
    *out = tm;
    return true;
}

template <int CP, int GR, int NDX>
static int launch_wgrad(const void* in, long long in_sstride, const void* dy, long long dy_sstride,
                        float* workspace, int B, int pitch, int S, int Hv, int Wv, int* grid_out,
                        cudaStream_t stream) {
    constexpr int CPL = CP / 8 + 1;
    WgGeom g;
    g.pitch = pitch; g.S = S; g.Hv = Hv; g.Wv = Wv; g.B = B;
    // CURLA_WG_COPIES: staged dy copies (horizontal taps carried in N); the remaining taps are A-side shifts
    g.nc = NDX;
    { const char* e = getenv("CURLA_WG_COPIES"); if (e && e[0] >= '1' && e[0] <= '0' + NDX) g.nc = e[0] - '0'; }
    const int NC = g.nc;
    g.kwin = cdiv(Wv + NC - 1, 16);
    g.dyr = g.kwin * 16 > pitch + NC - 1 ? g.kwin * 16 : pitch + NC - 1;
    g.dyr = (g.dyr + 7) / 8 * 8;
    // operand staging: tensor-map boxes (CURLA_WG_TMAP=0: one linear bulk copy per plane per image row)
    CUtensorMap tmIn, tmDy;
    memset(&tmIn, 0, sizeof(tmIn));
    memset(&tmDy, 0, sizeof(tmDy));
    // CURLA_WG_TMAP=1 (experiment, off by default: measured no faster, 0.380 vs 0.382 ms per update -- the number of
    // requests per stage was not what paces the fill)
    g.tmap = 0;
    { const char* e = getenv("CURLA_WG_TMAP"); if (e && e[0] == '1') g.tmap = 1; }
    g.dy1 = 0;
    { const char* e = getenv("CURLA_WG_DY1"); if (e && e[0] == '1' && NC == NDX) g.dy1 = 1; }
    // CURLA_WG_PRODUCERS = 1 + helper warps issuing the dy copies (linear staging only)
    // (default: the producer warp + four helpers -- measured 0.405 -> 0.368 ms per update, gpurun r04e)
    g.nh = 4;
    { const char* e = getenv("CURLA_WG_PRODUCERS"); if (e && e[0] >= '1' && e[0] <= '5') g.nh = e[0] - '1'; }
    if (g.tmap || g.dy1) g.nh = 0;
    g.pss = pitch * 16;
    if (g.tmap) {
        const int pss = (pitch * 16 + 127) / 128 * 128;
        const bool ok = S % pitch == 0 &&
                        wg_map(&tmIn, in, pitch, S / pitch, CP / 8, B, (long long)S * 8, in_sstride, pss / 8) &&
                        wg_map(&tmDy, dy, pitch, S / pitch, 4, B, (long long)S * 8, dy_sstride, g.dyr * 2);
        if (ok) g.pss = pss; else g.tmap = 0;
    }
    const size_t PS = (size_t)g.pss, DYB = (size_t)g.dyr * 16;
    const size_t budget = 225 * 1024 - 256;
    // CURLA_WG_STAGES (timing experiments): ring depth 2..4; R = the most dy rows per stage that fit
    int nst = 2;
    { const char* e = getenv("CURLA_WG_STAGES"); if (e && e[0] >= '2' && e[0] <= '4') nst = e[0] - '0'; }
    int R = 0;
    for (int r = 1; r <= 16; ++r) {
        const size_t stage = (((size_t)(r + GR - 1) * CPL * PS + (size_t)r * 4 * NC * DYB) + 127) & ~(size_t)127;
        // the junk chunks of the last row (up to chunk 15) must stay inside the stage: reads reach
        // slot r-1 + ceil(16 / CPL) planes, covered by the dy region behind the A slots
        if ((size_t)nst * stage <= budget) R = r;
    }
    if (R < 1) { set_last_error("conv_wgrad: pitch %d does not fit shared memory", pitch); return -1; }
    if (R > Hv) R = Hv;
    { const char* e = getenv("CURLA_WG_ROWS"); if (e && atoi(e) >= 1 && atoi(e) <= R) R = atoi(e); }     // timing experiments
    g.R = R;
    g.nst = nst;
    g.runs_per_sample = cdiv(Hv, R);
    g.total_runs = B * g.runs_per_sample;
    const size_t stage = (((size_t)(R + GR - 1) * CPL * PS + (size_t)R * 4 * NC * DYB) + 127) & ~(size_t)127;
    const size_t smem = 128 + (size_t)nst * stage;
    auto kern = k_conv_wgrad_tc<CP, GR, NDX>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_last_error("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e)); return -1; }
    const int cap = conv_grid_cap();
    const int grid = g.total_runs < cap ? g.total_runs : cap;
    launch_k(kern, dim3(grid), dim3(kWgThreads), smem, stream, tmIn, tmDy, (const bf16*)in, in_sstride, (const bf16*)dy, dy_sstride,
             workspace, g);
    *grid_out = grid;
    return 0;
}

}  // namespace curla

using namespace curla;

extern "C" long long curla_conv_wgrad_workspace_floats(int first_layer) {
    const long long per = first_layer ? (4 * 48 * 32 + 32) : (9 * 32 * 32 + 32);
    return per * (long long)(sm_count() * 2);
}

// First stage only: per-CTA partials into `workspace` (curla_conv_wgrad_workspace_floats);
// *nparts_out = number of partial slices written.  Finish with curla_conv_wgrad_reduce_multi.
extern "C" int curla_conv_wgrad_partial(const void* in, long long in_sstride, const void* dy,
                                        long long dy_sstride, float* workspace, int B, int pitch,
                                        int S, int Hv, int Wv, int first_layer, int* nparts_out,
                                        cudaStream_t stream) {
    if (first_layer) {
        if (launch_wgrad<48, 2, 2>(in, in_sstride, dy, dy_sstride, workspace, B, pitch, S, Hv, Wv, nparts_out, stream)) return -1;
    } else {
        if (launch_wgrad<32, 3, 3>(in, in_sstride, dy, dy_sstride, workspace, B, pitch, S, Hv, Wv, nparts_out, stream)) return -1;
    }
    return check_launch("conv_wgrad");
}

// Deterministic second stage of n <= 4 layers in one launch.  Per layer l: workspace[l],
// nparts[l], first_layer[l] (selects the 2x2 space-to-depth tap mapping), Cin[l], scale[l],
// dW[l] (OIHW fp32), db[l] (32 floats).
extern "C" int curla_conv_wgrad_reduce_multi(int n, float* const* workspace, const int* nparts,
                                             const int* first_layer, const int* Cin,
                                             const float* scale, float* const* dW, float* const* db,
                                             cudaStream_t stream) {
    CURLA_CHECK(n >= 1 && n <= 4, "conv_wgrad_reduce_multi: 1..4 layers");
    WgReduceJobs jobs;
    int per_max = 0;
    for (int l = 0; l < n; ++l) {
        WgReduceJob& J = jobs.j[l];
        J.partial = workspace[l]; J.nparts = nparts[l]; J.ntaps = first_layer[l] ? 4 : 9; J.CP = first_layer[l] ? 48 : 32;
        J.Cin = Cin[l]; J.s2d = first_layer[l] ? 1 : 0; J.scale = scale[l]; J.dW = dW[l]; J.db = db[l];
        const int per = J.ntaps * J.CP * 32 + 32;
        per_max = per > per_max ? per : per_max;
    }
    launch_k(k_conv_wgrad_reduce_multi, dim3(cdiv(per_max, 32), n), dim3(32, 8), 0, stream, jobs);
    return check_launch("conv_wgrad_reduce");
}

// dW (OIHW fp32, Cin real channels), db[32].  Hv/Wv = valid dims of dY (this layer's OUTPUT).
extern "C" int curla_conv_wgrad(const void* in, long long in_sstride, const void* dy,
                                long long dy_sstride, float* workspace, float* dW, float* db,
                                float scale, int B, int pitch, int S, int Hv, int Wv, int Cin,
                                int first_layer, cudaStream_t stream) {
    int grid = 0;
    if (first_layer) {
        if (launch_wgrad<48, 2, 2>(in, in_sstride, dy, dy_sstride, workspace, B, pitch, S, Hv, Wv, &grid, stream)) return -1;
        if (check_launch("conv_wgrad")) return -1;
        const int per = 4 * 48 * 32 + 32;
        launch_k(k_conv_wgrad_reduce, dim3(cdiv(per, 32)), dim3(32, 8), 0, stream, workspace, grid, 4, 48, Cin, 1, scale, dW, db);
    } else {
        if (launch_wgrad<32, 3, 3>(in, in_sstride, dy, dy_sstride, workspace, B, pitch, S, Hv, Wv, &grid, stream)) return -1;
        if (check_launch("conv_wgrad")) return -1;
        const int per = 9 * 32 * 32 + 32;
        launch_k(k_conv_wgrad_reduce, dim3(cdiv(per, 32)), dim3(32, 8), 0, stream, workspace, grid, 9, 32, Cin, 0, scale, dW, db);
    }
    return check_launch("conv_wgrad_reduce");
}
