// K4/K5 on the 5th-generation tensor cores: conv forward and dgrad as shifted GEMMs issued
// with tcgen05.mma, accumulators in TMEM.
//
// Replaces cuDNN conv fwd / dgrad (+ ReLU, + obs/255) launched by
//   encoder.py:77-90     CNNEncoder.forward_conv
//   curl_sac.py:367,417  loss.backward() through the conv layers
//
// Data layout: channel-plane bf16 activations with a fixed row
// pitch, so a 3x3 valid conv over the flattened position index P is
//     out[P][0:32] = sum_t in[P + off_t][0:CP] . W_t          (off_t = dy*pitch + dx)
// i.e. nine GEMMs whose A operand is the SAME matrix shifted by off_t rows.
//
// Shared-memory operand layout -- the point of this file.  tcgen05.mma reads A and B through
// 64-bit matrix descriptors.  In the K-major / no-swizzle canonical layout a core matrix is
// 8 rows x 16 bytes stored contiguously (rows 16 B apart), 8-row groups are SBO bytes apart
// and the two 16-byte K chunks of one K=16 step are LBO bytes apart.  The slab is stored as
// CHANNEL-CHUNK PLANES  slab[chunk c][row r][8 channels]  (16 B per row per plane), so that
// SBO = 128 makes ALL rows of a plane 16 B apart and LBO = plane size.  A tap shift of off_t
// rows is then just  start_address += 16 * off_t : nine taps = nine descriptors into one
// slab, no im2col, no data movement, any shift (only 16-byte alignment is required when no
// swizzle is used).  Weights use the same layout [tap][chunk][32 n][8 k].
//
// One persistent CTA per SM over tiles of 256 positions, warp-specialised: a producer thread
// hands out tiles (atomic counter) and fills a ring of slabs with one cp.async.bulk per channel
// plane; two issuer warps alternate tiles, each issuing 2*NTAPS*(CP/16) MMAs of M=128, N=32,
// K=16, bf16 -> fp32 into one of four TMEM accumulators; tcgen05.commit arrives on mbarriers;
// two groups of eight epilogue warps pull their 32 TMEM lanes with tcgen05.ld -- one output
// position per thread, 32 channels -- and apply scale+bias+ReLU or the ReLU mask of dgrad; a
// warp stores 32 consecutive positions x 16 B = 512 contiguous bytes per channel plane.  Up to
// three independent passes (own input / output / weight set) share one launch.  See the kernel.
#include "common.cuh"
#include "tc.cuh"
#include "../../include/curla_b200.h"

#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <utility>

namespace curla {

struct TcGeom {
    int pitch, S, Hv, Wv;
    int tiles_per_sample, total_tiles;
    int slab_rows;       // rows actually loaded (TM + span)
    int plane_rows;      // slab_rows rounded up to 8 (plane stride = plane_rows * 16 B)
    int min_off;         // most negative tap shift
    int stages;          // slab ring depth (host: as many as fit in shared memory, <= 8)
    int plane_bytes;     // shared-memory stride between channel planes of a slab
    int zero_planes;         // trailing channel planes of the input that are all zero (conv-1: 4C = 36 real channels of 48):
                             // never loaded, their shared-memory planes are zeroed once
    int dynamic, ctr_slot;   // tiles handed out through an atomic counter (g_tc_ctr pair ctr_slot) instead of round robin
    float inv_tps, inv_pitch;   // 1/tiles_per_sample, 1/pitch: index divisions through a float reciprocal (exact for
                         // these ranges: (i + 0.5) / d is never within float rounding of an integer)
    int debug;           // CURLA_TC_DEBUG bitmask (timing experiments only): 1 no loads, 2 no MMA, 4 no stores,
                         // 8 tap shifts rounded to 8 rows (WRONG results: aligned core matrices), 16 plane skew +64 B,
                         // 64 role cycle counters, 128 single MMA issuer warp
};
struct TcTaps { int off[9]; };
// Up to three independent passes ("segments") of the same layer in one launch: tiles
// [tile_end[s-1], tile_end[s]) belong to segment s, which has its own input/output buffers and
// one of (at most) two weight sets staged in shared memory.  Unused tile_end entries hold the
// total tile count.
struct TcSegs {
    const bf16* in[3];
    bf16* out[3];
    const bf16* wts[2];
    const float* bias[2];
    int tile_end[3];
    int wsel[3];
    int nw;
};
struct TileLoc { int seg, b, t; };
__device__ __forceinline__ int tc_div(int i, float inv_d) { return __float2int_rd(((float)i + 0.5f) * inv_d); }
__device__ __forceinline__ TileLoc tc_locate(const TcSegs& s, int tiles_per_sample, float inv_tps, int tile) {
    int seg = 0, start = 0;
    if (tile >= s.tile_end[0]) { seg = 1; start = s.tile_end[0]; }
    if (tile >= s.tile_end[1]) { seg = 2; start = s.tile_end[1]; }
    const int local = tile - start;
    const int b = tc_div(local, inv_tps);
    TileLoc l;
    l.seg = seg; l.b = b; l.t = local - b * tiles_per_sample;
    return l;
}
template <typename T>
__device__ __forceinline__ T tc_pick(T const (&a)[3], int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); }

// CURLA_TC_DEBUG & 64: per-CTA cycle counters of the MMA thread and the producer (timing
// experiments only): [0] MMA wait tempty, [1] MMA wait full, [2] MMA issue, [3] producer wait
// empty, [4] kernel total, [5] tiles, [6] epilogue warp 0 wait tfull, [7] epilogue warp 0 busy
//   [8] globaltimer (ns) at kernel entry, [9] at MMA loop start, [10] at MMA loop end, [11] at CTA exit, [12] %smid,
//   [13] issuer warp 1: loop end (globaltimer), [14] issuer warp 1: issue+sync clk,
//   [15] producer: waiting for a tile id (counter atomic), [16] producer: publishing + issuing the bulk copies, [17] producer loop total
__device__ long long g_tc_dbg[160][20];
// Dynamic tile scheduler: pairs {next tile, finished CTAs}; a launch takes the next pair of the
// pool (host side, round robin) and its last CTA zeroes the pair again.
constexpr int kCtrSlots = 256;
__device__ int g_tc_ctr[2 * kCtrSlots];
__device__ __forceinline__ int ld_volatile_s32(const void* p) { return *reinterpret_cast<const volatile int*>(p); }
// Ring entry of local tile j: (j mod 1024) << 20 | (tile + 1); tile = -1 ends the stream.  ONE 32-bit
// store publishes id and readiness together, so the producer needs no fence between "id" and
// "count" stores -- a MEMBAR there also waits for its in-flight counter atomics (measured: the
// producer then became the bottleneck).  Slots start as 0xFFFFFFFF (tag 4095: matches no j).
// atomicAdd(p, 1) -- and the same atom.add / atom.inc written as inline PTX, even with a
// clock-derived addend -- is compiled into a warp-aggregated sequence (vote, leader atomic, SHFL
// of the result) whose SHFL consumes the result at once: the issuing thread waits out every
// round trip.  An addend ptxas cannot prove warp-uniform (a volatile shared-memory load from a
// per-lane address; the words all hold 1) keeps it a plain ATOMG whose result register is only
// waited for where it is first read.
__device__ __forceinline__ int atom_add_nonuniform(int* p, int addend) {
    int old;
    asm volatile("atom.relaxed.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(addend) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t ring_entry(int j, int tile) { return ((uint32_t)(j & 0x3FF) << 20) | (uint32_t)(tile + 1); }
__device__ __forceinline__ bool ring_ready(uint32_t e, int j) { return (e >> 20) == (uint32_t)(j & 0x3FF); }
__device__ __forceinline__ int ring_tile(uint32_t e) { return (int)(e & 0xFFFFFu) - 1; }
__device__ __forceinline__ long long tc_gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ---------------------------------------------------------------- the kernel
// Warp-specialised, persistent, one CTA per SM:
//   warps 0..15  epilogue   (two groups of 8 alternating tiles; TMEM -> registers -> bias/ReLU or ReLU-mask -> plane stores: a
//                            warp writes 32 consecutive positions x 16 B = 512 contiguous bytes)
//   warps 16,17  MMA issuers (one elected lane each, alternating tiles; 2 x NTAPS x CP/16 tcgen05.mma per tile)
//   warp  18     producer   (one elected lane: CP/8 bulk copies cp.async.bulk global -> shared
//                            per tile, one per channel plane -- the global layout IS the
//                            shared-memory operand layout -- completing on the stage's mbarrier)
// Tiles: the producer decides which tile is the CTA's local tile j (the first two round robin,
// the rest from a global atomic counter: SMs drain tiles at visibly different rates -- the even
// SM of a TPC ~14 % slower in an MMA-only run, profiles/r01l_conv_sm_imbalance.txt -- and a static
// split waits for the slowest) and publishes it in a shared-memory ring before filling the slab;
// two -1 entries end the stream (one per issuer warp / epilogue group).
// Pipelines: full/empty per slab stage (producer <-> MMA), tfull/tempty per TMEM accumulator
// stage (MMA <-> epilogue; 4 stages x 2 sub-tiles x 32 columns = 256 TMEM columns), so loads
// of tile i+k, the MMAs of tile i+1 and the epilogue of tile i all run concurrently.
//  DGRAD=false: B = W_t[n][k]            -> out = relu(acc*scale + bias)
//  DGRAD=true : B = W_t^T (n=ci, k=co)   -> out = (X > 0) ? acc : 0
constexpr int kTcSub = 2;                 // 256 positions per tile
constexpr int kEpiWarps = 4 * kTcSub;     // 8 warps drain one tile (2 sub-tiles x 4 lane quarters)
constexpr int kEpiGroups = 2;             // two such groups alternate tiles: a warp's per-tile latency
                                          // (wait, TMEM load, math, stores) is hidden behind the other group
constexpr int kEpiAll = kEpiWarps * kEpiGroups;      // 16
constexpr int kMmaWarps = 2;                         // two issuer warps alternate tiles (see the kernel)
constexpr int kTcThreads = (kEpiAll + kMmaWarps + 1) * 32;       // 608
constexpr int kMaxStages = 8;
constexpr int kAccStages = 4;            // 256 TMEM columns (8 stages = all 512 columns measured no faster)
constexpr int kSmemHdr = 768;             // barriers + tmem ptr + bias + tile ring + 32 ones
constexpr int kRing = 16;                 // published tile ids: local tile j -> slot j % 16 (deeper than slab + accumulator stages)

template <int CP, int NTAPS, bool DGRAD>
__global__ void __launch_bounds__(kTcThreads, 1)
k_conv_tc(const __grid_constant__ TcSegs sg, long long in_sstride,    // weights: [NTAPS][32][CP] per set
          float scale, const bf16* __restrict__ relu_src, long long out_sstride, TcGeom g, TcTaps taps) {
    constexpr int CH = CP / 8, KS = CP / 16, TM = kTcSub * 128;
    constexpr uint32_t W_BYTES = NTAPS * CH * 512;
    constexpr uint32_t TMEM_COLS = kAccStages * kTcSub * 32;   // 256
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s_base = smem_u32(smem);
    if ((g.debug & 64) && tid == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_tc_dbg[blockIdx.x][8] = tc_gtime();
        g_tc_dbg[blockIdx.x][12] = smid;
    }
    // header: full[8] @0, empty[8] @64, tfull[8] @128, tempty[8] @192, bias set 0 @256, set 1 @384,
    //         tile ring[16] @512, tmem ptr @576, ones[32] @640 (see atom_add_nonuniform)
    const uint32_t s_full = s_base, s_empty = s_base + 64, s_tfull = s_base + 128, s_tempty = s_base + 192;
    const uint32_t s_tptr = s_base + 576;
    const uint32_t s_w = s_base + kSmemHdr;
    const uint32_t PS = (uint32_t)g.plane_bytes;
    const uint32_t slab_bytes = CH * PS;
    const uint32_t s_slab0 = s_w + (uint32_t)sg.nw * W_BYTES;
    const int stages = g.stages;
    const long long plane = (long long)g.S * 8;      // elements between channel planes (in and out)

    // ---- one-time setup: barriers, TMEM, weights, bias
    if (tid == 0) {
        for (int i = 0; i < kRing; ++i) reinterpret_cast<volatile uint32_t*>(smem + 512)[i] = 0xFFFFFFFFu;
        for (int i = 0; i < 32; ++i) reinterpret_cast<volatile int*>(smem + 640)[i] = 1;
        for (int i = 0; i < stages; ++i) {
            mbar_init(s_full + 8 * i, 1);
            mbar_init(s_empty + 8 * i, 1);
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(s_tfull + 8 * i, 1);
            mbar_init(s_tempty + 8 * i, kEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tptr),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();              // barriers initialised, TMEM address published
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + 576);
    // everything above is independent of earlier kernels; from here on we read their outputs
    pdl_grid_sync();
    // The producer warp goes straight to its loop (its first bulk copies overlap the weight
    // staging below); the other 17 warps stage the weights and meet on named barrier 1.
    constexpr int kStageThreads = kTcThreads - 32;
    if (warp != kEpiAll + kMmaWarps) {
        for (int ws = 0; ws < sg.nw; ++ws) {
            const bf16* __restrict__ wts = ws ? sg.wts[1] : sg.wts[0];
            if (!DGRAD) {
                for (int i = tid; i < NTAPS * 32 * CH; i += kStageThreads) {
                    const int t = i / (32 * CH), rem = i - t * (32 * CH), n = rem / CH, kc = rem - n * CH;
                    cp_async16(s_w + ws * W_BYTES + (uint32_t)(((t * CH + kc) * 32 + n) * 16), wts + ((t * 32 + n) * CP + kc * 8), 16);
                }
                if (tid < 32) reinterpret_cast<float*>(smem + 256 + ws * 128)[tid] = (ws ? sg.bias[1] : sg.bias[0])[tid];
            } else {
                // B[n=ci][k=co] = W_t[co][ci]: transpose while staging (once per persistent CTA)
                for (int i = tid; i < NTAPS * 32 * CP; i += kStageThreads) {
                    const int t = i / (32 * CP), rem = i - t * (32 * CP), co = rem / CP, ci = rem - co * CP;
                    reinterpret_cast<bf16*>(smem + kSmemHdr + ws * W_BYTES)[((t * CH + (co >> 3)) * 32 + ci) * 8 + (co & 7)] = wts[i];
                }
            }
        }
        if (g.zero_planes > 0) {
            const int per = g.zero_planes * (int)(PS / 16);            // 16-byte words per stage
            for (int i = tid; i < stages * per; i += kStageThreads) {
                const int st = i / per, w = i - st * per;
                asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(s_slab0 + st * slab_bytes + (CH - g.zero_planes) * PS + w * 16), "r"(0) : "memory");
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();      // generic-proxy writes of the weights (and zero planes) -> visible to the tensor core
        asm volatile("bar.sync 1, %0;" ::"n"(kStageThreads) : "memory");
    }

    if (warp < kEpiAll) {
        // ================= epilogue: one output position per thread per tile
        const int egroup = warp / kEpiWarps, ew = warp % kEpiWarps;
        const int sub = ew >> 2, quarter = ew & 3;
        const int row_in_tile = sub * 128 + quarter * 32 + lane;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sub * 32);
        uint32_t acc = 0, acc_phase = 0;
        // bias in registers: shared memory is saturated by the tensor core's operand reads
        float bz[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) bz[i] = 0.f;
        int bz_set = -1;                           // weight set whose bias bz[] holds (reloaded when a CTA crosses into another set)
        // dgrad: the ReLU mask (X at the same position) is fetched one tile ahead so that its
        // DRAM latency never sits between an accumulator becoming ready and being drained
        uint4 xn[4];
        int xn_tile = -1;
        auto load_mask = [&](int tile_, uint4 (&dst)[4]) {          // dgrad launches have one segment
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[c] = make_uint4(0u, 0u, 0u, 0u);
            const int b_ = tc_div(tile_, g.inv_tps);
            const int p_ = (tile_ - b_ * g.tiles_per_sample) * TM + row_in_tile;
            const int y_ = tc_div(p_, g.inv_pitch), x_ = p_ - y_ * g.pitch;
            if (y_ < g.Hv && x_ < g.Wv && p_ < g.S) {
                const long long o_ = (long long)b_ * out_sstride + (long long)p_ * 8;
#pragma unroll
                for (int c = 0; c < 4; ++c) dst[c] = *reinterpret_cast<const uint4*>(relu_src + o_ + c * plane);
            }
        };
        long long e_wait = 0;
        const bool edbg = (g.debug & 64) != 0;
        const long long e_start = clock64();
        acc = egroup;                              // local tile j uses accumulator stage j % kAccStages
        for (int j = egroup;; j += kEpiGroups) {
            uint32_t ent = (uint32_t)ld_volatile_s32(smem + 512 + 4 * (j & (kRing - 1)));
            if (!ring_ready(ent, j)) {                          // not published yet (rare: the producer runs stages ahead)
                const long long t0 = clock64();
                while (!ring_ready(ent = (uint32_t)ld_volatile_s32(smem + 512 + 4 * (j & (kRing - 1))), j))
                    if (clock64() - t0 > (1ll << 31)) __trap();
            }
            const int tile = ring_tile(ent);
            if (tile < 0) break;
            const TileLoc tl = tc_locate(sg, g.tiles_per_sample, g.inv_tps, tile);
            const int p = tl.t * TM + row_in_tile;
            const int y = tc_div(p, g.inv_pitch), x = p - y * g.pitch;
            const bool valid = (y < g.Hv) && (x < g.Wv);
            bf16* __restrict__ out = tc_pick(sg.out, tl.seg);
            const long long o = (long long)tl.b * out_sstride + (long long)p * 8;   // + c*plane
            if (!DGRAD) {
                const int ws = tc_pick(sg.wsel, tl.seg);
                if (ws != bz_set) {
                    bz_set = ws;
#pragma unroll
                    for (int i = 0; i < 32; ++i) bz[i] = reinterpret_cast<const float*>(smem + 256 + ws * 128)[i];
                }
            }
            uint4 xm[4];
            if (DGRAD) {
                if (xn_tile == tile) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) xm[c] = xn[c];
                } else {
                    load_mask(tile, xm);
                }
                // this group's next tile, if the producer has published it already (it normally
                // has: the producer runs a ring of slabs ahead): its mask is fetched now
                xn_tile = -1;
                const uint32_t e2 = (uint32_t)ld_volatile_s32(smem + 512 + 4 * ((j + kEpiGroups) & (kRing - 1)));
                if (ring_ready(e2, j + kEpiGroups)) {
                    xn_tile = ring_tile(e2);
                    if (xn_tile >= 0) load_mask(xn_tile, xn);
                }
            }
            const long long e0 = edbg ? clock64() : 0;
            mbar_wait(s_tfull + 8 * acc, acc_phase);
            if (edbg) e_wait += clock64() - e0;
            tc_fence_after();
            uint32_t r[32];
            tmem_ld32(taddr0 + acc * (kTcSub * 32), r);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_tempty + 8 * acc);   // this warp's 32 lanes are drained
            if (p < g.S && !(g.debug & 4)) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float v0 = __uint_as_float(r[c * 8 + j * 2]), v1 = __uint_as_float(r[c * 8 + j * 2 + 1]);
                        if (!DGRAD) {
                            v0 = valid ? fmaxf(fmaf(v0, scale, bz[c * 8 + j * 2]), 0.f) : 0.f;
                            v1 = valid ? fmaxf(fmaf(v1, scale, bz[c * 8 + j * 2 + 1]), 0.f) : 0.f;
                        } else {
                            const uint32_t m = j == 0 ? xm[c].x : (j == 1 ? xm[c].y : (j == 2 ? xm[c].z : xm[c].w));
                            const float2 xv = unpack_bf16x2(m);
                            v0 = xv.x > 0.f ? v0 : 0.f;
                            v1 = xv.y > 0.f ? v1 : 0.f;
                        }
                        w[j] = pack_bf16x2(v0, v1);
                    }
                    *reinterpret_cast<uint4*>(out + o + c * plane) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            acc += kEpiGroups;
            if (acc >= kAccStages) { acc -= kAccStages; acc_phase ^= 1; }
        }
        if ((g.debug & 64) && tid == 0) { g_tc_dbg[blockIdx.x][6] = e_wait; g_tc_dbg[blockIdx.x][7] = clock64() - e_start - e_wait; }
    } else if (warp < kEpiAll + kMmaWarps) {
        // ================= MMA issuers: warp mw takes the CTA's local tiles mw, mw + 2, ...
        // Measured with one issuer (profiles/r01k_conv_roles.txt): a tile's 36 MMAs execute in
        // ~1440 clk, but the issuing thread blocks while the MMA queue is full and then spends
        // ~690 clk per tile on the commits, the two barrier waits and the tile bookkeeping before
        // it can issue again -- the tensor pipe idled a third of the time.  With two issuers the
        // other warp's MMAs are already queued while this one does its bookkeeping.
        const int mw = warp - kEpiAll;
        const int nmw = (g.debug & 128) ? 1 : kMmaWarps;
        const bool dbg = (g.debug & 64) != 0;
        if (mw < nmw) {
        uint32_t stage = (uint32_t)mw, phase = 0, acc = (uint32_t)mw, acc_phase = 0;
        // descriptors differ only in the 14-bit start-address field: precompute the field
        // offsets (16-byte units) once, add the stage base per tile
        const uint64_t a_hi = make_desc(0, PS, 128), b_hi = make_desc(0, 512, 128);
        uint32_t a_off[NTAPS * KS];
#pragma unroll
        for (int t = 0; t < NTAPS; ++t)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
                a_off[t * KS + ks] = ((uint32_t)(2 * ks) * PS +
                                      (uint32_t)((g.debug & 8) ? ((taps.off[t] - g.min_off) & ~7) : (taps.off[t] - g.min_off)) * 16u) >> 4;
        long long c_te = 0, c_fu = 0, c_is = 0, c_n = 0;
        const long long c_start = clock64();
        if (dbg && lane == 0 && mw == 0) g_tc_dbg[blockIdx.x][9] = tc_gtime();
        const int valid_pos = g.Hv * g.pitch;
        for (int j = mw;; j += nmw) {
            const long long c1 = dbg ? clock64() : 0;
            mbar_wait(s_full + 8 * stage, phase);                  // slab filled, tile id published
            const long long c2 = dbg ? clock64() : 0;
            const int tile = ring_tile((uint32_t)ld_volatile_s32(smem + 512 + 4 * (j & (kRing - 1))));
            if (tile < 0) break;
            const TileLoc tl = tc_locate(sg, g.tiles_per_sample, g.inv_tps, tile);
            // sub-tiles that start at or beyond the last valid output row hold no valid position:
            // their MMAs are skipped (the epilogue writes exact zeros there whatever TMEM holds)
            const int nsub = (tl.t * TM + 128 >= valid_pos) ? 1 : kTcSub;
            const uint32_t w16 = (s_w + (uint32_t)tc_pick(sg.wsel, tl.seg) * W_BYTES) >> 4;
            const uint32_t slab16 = (s_slab0 + stage * slab_bytes) >> 4;
            mbar_wait(s_tempty + 8 * acc, acc_phase ^ 1);
            const long long c3 = dbg ? clock64() : 0;
            tc_fence_after();
            if (elect_one()) {
                if (!(g.debug & 2))
#pragma unroll
                for (int s = 0; s < kTcSub; ++s) {
                    if (s >= nsub) break;
                    const uint32_t d = tmem_base + acc * (kTcSub * 32) + (uint32_t)(s * 32);
#pragma unroll
                    for (int t = 0; t < NTAPS; ++t) {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint64_t ad = a_hi | (uint64_t)((slab16 + (uint32_t)(s * 128) + a_off[t * KS + ks]) & 0x3FFFu);
                            const uint64_t bd = b_hi | (uint64_t)((w16 + (uint32_t)((t * CH + 2 * ks) * 32)) & 0x3FFFu);
                            if (t | ks) umma_bf16<1>(d, ad, bd); else umma_bf16<0>(d, ad, bd);
                        }
                    }
                }
                umma_commit(s_empty + 8 * stage);     // slab free once these MMAs have read it
                umma_commit(s_tfull + 8 * acc);       // accumulators complete
            }
            __syncwarp();
            if (dbg) { c_te += c3 - c2; c_fu += c2 - c1; ++c_n; c_is += clock64() - c3; }
            stage += (uint32_t)nmw;
            if (stage >= (uint32_t)stages) { stage -= (uint32_t)stages; phase ^= 1; }
            acc += (uint32_t)nmw;
            if (acc >= kAccStages) { acc -= kAccStages; acc_phase ^= 1; }
        }
        if (dbg && lane == 0) {
            long long* d = g_tc_dbg[blockIdx.x];
            if (mw == 0) { d[0] = c_te; d[1] = c_fu; d[2] = c_is; d[4] = clock64() - c_start; d[5] = c_n; d[10] = tc_gtime(); }
            else { d[13] = tc_gtime(); d[14] = c_is; }
        }
        }
    } else {
        const int one = ld_volatile_s32(smem + 640 + 4 * lane);       // 1, but not provably warp-uniform
        if (lane == 0) {
        // ================= producer (one thread): slab[c][0:rows][16 B] <- plane c rows [p0+min_off, +rows)
        uint32_t stage = 0, phase = 0;
        long long c_pw = 0, c_q = 0, c_cp = 0;
        const long long c_p0 = clock64();
        const uint32_t bytes = (uint32_t)g.slab_rows * 16u;
        int* const ctr = g_tc_ctr + 2 * g.ctr_slot;
        const int G = (int)gridDim.x;
        // Tile queue of four FIXED registers (the loop is unrolled by four): slot u is consumed,
        // then immediately re-armed with an atomic whose value is first read four tiles later, so
        // up to four counter round trips (~1 us each under load) are in flight and none is waited
        // for.  (A shifting queue q0 <- q1 <- q2 stalls on the register move of the youngest entry.)
        int q[4];
        // q[] holds tile - 2G (the raw counter value): adding the offset where the atomic is issued
        // would consume its result -- and wait for it -- on the spot
        q[0] = (int)blockIdx.x - 2 * G; q[1] = q[0] + G;
        q[2] = g.dynamic ? atom_add_nonuniform(ctr, one) : q[1] + G;
        q[3] = g.dynamic ? atom_add_nonuniform(ctr, one) : q[2] + G;
        int sentinels = 0, j = 0;
        while (sentinels < 2) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (sentinels >= 2) break;
                const long long cq = (g.debug & 64) ? clock64() : 0;
                const int tile = q[u] + 2 * G;
                const bool real = tile < g.total_tiles;
                const long long c0 = (g.debug & 64) ? clock64() + (real ? 0 : 1) : 0;      // (depends on the id: read after it arrived)
                if (g.debug & 64) c_q += c0 - cq;
                mbar_wait(s_empty + 8 * stage, phase ^ 1);
                const long long c1 = (g.debug & 64) ? clock64() : 0;
                if (g.debug & 64) c_pw += c1 - c0;
                *reinterpret_cast<volatile uint32_t*>(smem + 512 + 4 * (j & (kRing - 1))) = ring_entry(j, real ? tile : -1);
                const uint32_t bar = s_full + 8 * stage;
                if (!real || (g.debug & 1)) {
                    mbar_arrive(bar);
                } else {
                    const TileLoc tl = tc_locate(sg, g.tiles_per_sample, g.inv_tps, tile);
                    const int p0 = tl.t * TM;
                    const bf16* src = tc_pick(sg.in, tl.seg) + (long long)tl.b * in_sstride + (long long)(p0 + g.min_off) * 8;
                    const uint32_t dst = s_slab0 + stage * slab_bytes;
                    const int chl = CH - g.zero_planes;
                    mbar_expect_tx(bar, bytes * chl);
#pragma unroll
                    for (int c = 0; c < CH; ++c)
                        if (c < chl) bulk_g2s(dst + c * PS, src + c * plane, bytes, bar);
                }
                if (g.debug & 64) c_cp += clock64() - c1;
                ++j;
                if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
                if (!real) ++sentinels;                       // ids only grow: everything after is past the end too
                else q[u] = g.dynamic ? atom_add_nonuniform(ctr, one) : tile + 2 * G;
            }
        }
        if (g.dynamic) {
            // every tile this CTA will ever ask for has been asked for: the last CTA to get here
            // re-arms the counter pair for the launch that takes it next
            __threadfence();
            if (atomicAdd(ctr + 1, 1) == G - 1) { ctr[0] = 0; ctr[1] = 0; __threadfence(); }
        }
        if (g.debug & 64) {
            long long* d = g_tc_dbg[blockIdx.x];
            d[3] = c_pw; d[15] = c_q; d[16] = c_cp; d[17] = clock64() - c_p0;
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                     : "memory");
        if ((g.debug & 64) && lane == 0) g_tc_dbg[blockIdx.x][11] = tc_gtime();
    }
}

// ---------------------------------------------------------------- N = 96: the three horizontal taps in ONE MMA
// What bounds k_conv_tc on layers 2..4 is the A-operand stream: every tcgen05.mma re-reads its 128 x 32 B
// activation tile from shared memory whatever N is, so nine N = 32 taps x two K steps move 90 KB of operands
// per 128 positions (40 clk per MMA, 16 of them tensor work: profiles/r01c_microbench.txt).  Here the weights of
// the three horizontal taps sit side by side in N,
//     D[p][dx*32 + n] = sum_dy sum_k A[p + dy*pitch][k] . W(dy,dx)[n][k]          (6 MMAs of N = 96, 42 KB)
// and the horizontal shift moves from the operand address into the epilogue:
//     out[p][n] = D[p][n] + D[p+1][32 + n] + D[p+2][64 + n]      (dgrad: p-1, p-2 and transposed weights)
// Positions are TMEM lanes, so the shift is a warp shuffle of the accumulator registers; the two lanes at a
// warp's edge take their neighbours' rows from a 384-byte per-warp exchange slot in shared memory (one named
// barrier per tile among the 8 warps of an epilogue group), and tiles overlap by two rows (254 outputs per
// 256-row tile) so that no tile needs another tile's accumulators.
// Pipeline roles are those of k_conv_tc; 2 accumulator stages x 2 sub-tiles x 96 columns = 384 TMEM columns, so
// (issuer warp w, stage w, epilogue group w) form two independent lanes that alternate tiles.
constexpr int kT96Out = 254;
constexpr int kAcc96 = 2;
constexpr uint32_t kIdesc96 = (1u << 4) | (1u << 7) | (1u << 10) | ((96u >> 3) << 17) | ((128u >> 4) << 24);
constexpr int kXch = 96;                                     // floats per warp slot
constexpr int kSmemHdr96 = 768 + 2 * kEpiAll * kXch * 4;     // header + exchange slots [tile parity][warp][96]
constexpr uint32_t kW96Bytes = 3 * 4 * 96 * 16;              // one weight set: [dy][k chunk][96 n][8 k]

template <bool DGRAD>
__global__ void __launch_bounds__(kTcThreads, 1)
k_conv_tc96(const __grid_constant__ TcSegs sg, long long in_sstride, float scale, const bf16* __restrict__ relu_src,
            long long out_sstride, TcGeom g) {
    constexpr int CH = 4, KS = 2, TM = kTcSub * 128;
    constexpr uint32_t TMEM_COLS = 512;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_full = s_base, s_empty = s_base + 64, s_tfull = s_base + 128, s_tempty = s_base + 192;
    const uint32_t s_tptr = s_base + 576;
    const uint32_t s_w = s_base + kSmemHdr96;
    const uint32_t PS = (uint32_t)g.plane_bytes;
    const uint32_t slab_bytes = CH * PS;
    const uint32_t s_slab0 = s_w + (uint32_t)sg.nw * kW96Bytes;
    const int stages = g.stages;
    const long long plane = (long long)g.S * 8;
    const int tile_shift = DGRAD ? -2 : 0;                   // position of a tile's row 0: t * 254 + tile_shift

    if (tid == 0) {
        for (int i = 0; i < kRing; ++i) reinterpret_cast<volatile uint32_t*>(smem + 512)[i] = 0xFFFFFFFFu;
        for (int i = 0; i < 32; ++i) reinterpret_cast<volatile int*>(smem + 640)[i] = 1;
        for (int i = 0; i < stages; ++i) {
            mbar_init(s_full + 8 * i, 1);
            mbar_init(s_empty + 8 * i, 1);
        }
        for (int i = 0; i < kAcc96; ++i) {
            mbar_init(s_tfull + 8 * i, 1);
            mbar_init(s_tempty + 8 * i, kEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tptr), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + 576);
    pdl_grid_sync();
    constexpr int kStageThreads = kTcThreads - 32;
    if (warp != kEpiAll + kMmaWarps) {
        // weights: B[n = dx*32 + (co | ci)][k] per dy, K-major, [dy][k chunk][96 n][8 k]
        for (int ws = 0; ws < sg.nw; ++ws) {
            const bf16* __restrict__ wts = ws ? sg.wts[1] : sg.wts[0];            // global: [tap = dy*3 + dx][co][ci]
            if (!DGRAD) {
                for (int i = tid; i < 9 * 32 * CH; i += kStageThreads) {
                    const int t = i / (32 * CH), rem = i - t * (32 * CH), n = rem / CH, kc = rem - n * CH;
                    const int dy = t / 3, dx = t - dy * 3;
                    cp_async16(s_w + ws * kW96Bytes + (uint32_t)((((dy * CH + kc) * 96) + dx * 32 + n) * 16), wts + ((t * 32 + n) * 32 + kc * 8), 16);
                }
                if (tid < 32) reinterpret_cast<float*>(smem + 256 + ws * 128)[tid] = (ws ? sg.bias[1] : sg.bias[0])[tid];
            } else {
                for (int i = tid; i < 9 * 32 * 32; i += kStageThreads) {
                    const int t = i / 1024, rem = i - t * 1024, co = rem >> 5, ci = rem & 31;
                    const int dy = t / 3, dx = t - dy * 3;
                    reinterpret_cast<bf16*>(smem + kSmemHdr96 + ws * kW96Bytes)[(((dy * CH + (co >> 3)) * 96) + dx * 32 + ci) * 8 + (co & 7)] = wts[i];
                }
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(kStageThreads) : "memory");
    }

    if (warp < kEpiAll) {
        // ================= epilogue
        const int egroup = warp / kEpiWarps, ew = warp % kEpiWarps;
        const int sub = ew >> 2, quarter = ew & 3;
        const int row_in_tile = sub * 128 + quarter * 32 + lane;
        const bool row_out = DGRAD ? (row_in_tile >= 2) : (row_in_tile < kT96Out);      // rows this tile stores
        const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sub * 96);
        float* const xch = reinterpret_cast<float*>(smem + 768);
        uint32_t acc_phase = 0;
        const uint32_t acc = (uint32_t)egroup;
        auto load_mask = [&](int tile_, uint4 (&dst)[4]) {          // dgrad launches have one segment
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[c] = make_uint4(0u, 0u, 0u, 0u);
            const int b_ = tc_div(tile_, g.inv_tps);
            const int p_ = (tile_ - b_ * g.tiles_per_sample) * kT96Out + tile_shift + row_in_tile;
            if (p_ >= 0 && p_ < g.S && row_out) {
                const int y_ = tc_div(p_, g.inv_pitch), x_ = p_ - y_ * g.pitch;
                if (y_ < g.Hv && x_ < g.Wv) {
                    const long long o_ = (long long)b_ * out_sstride + (long long)p_ * 8;
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[c] = *reinterpret_cast<const uint4*>(relu_src + o_ + c * plane);
                }
            }
        };
        int par = 0;
        const bool edbg = (g.debug & 64) != 0;             // CURLA_TC_DEBUG=64: cycle counters of epilogue warp 0 and issuer 0
        long long e_wait = 0, e_tm = 0, e_bar = 0, e_rest = 0;
        for (int j = egroup;; j += kEpiGroups, par ^= 1) {
            uint32_t ent = (uint32_t)ld_volatile_s32(smem + 512 + 4 * (j & (kRing - 1)));
            if (!ring_ready(ent, j)) {
                const long long t0 = clock64();
                while (!ring_ready(ent = (uint32_t)ld_volatile_s32(smem + 512 + 4 * (j & (kRing - 1))), j))
                    if (clock64() - t0 > (1ll << 31)) __trap();
            }
            const int tile = ring_tile(ent);
            if (tile < 0) break;
            const TileLoc tl = tc_locate(sg, g.tiles_per_sample, g.inv_tps, tile);
            const int p = tl.t * kT96Out + tile_shift + row_in_tile;
            const bool inside = p >= 0 && p < g.S && row_out;
            int y = 0, x = 0;
            if (inside) { y = tc_div(p, g.inv_pitch); x = p - y * g.pitch; }
            const bool valid = inside && (y < g.Hv) && (x < g.Wv);
            bf16* __restrict__ out = tc_pick(sg.out, tl.seg);
            const long long o = (long long)tl.b * out_sstride + (long long)p * 8;   // + c*plane
            const float* bias = reinterpret_cast<const float*>(smem + 256 + (DGRAD ? 0 : tc_pick(sg.wsel, tl.seg)) * 128);
            // dgrad: the ReLU mask (X at the same position) is requested before the wait for the accumulator:
            // its latency overlaps this tile's MMAs (the tile id is published a ring of slabs ahead)
            uint4 xm[4];
            if (DGRAD) load_mask(tile, xm);
            float* const xw = xch + (size_t)(par * kEpiAll + warp) * kXch;
            const long long e0 = edbg ? clock64() : 0;
            mbar_wait(s_tfull + 8 * acc, acc_phase);
            const long long e1 = edbg ? clock64() : 0;
            tc_fence_after();
            const uint32_t taddr = taddr0 + acc * (kTcSub * 96);
            float av[32];
            uint32_t r[32];
            // ---- D1 (dx = 1): row p +- 1
            tmem_ld32(taddr + 32, r);
            if (lane == (DGRAD ? 31 : 0)) {
#pragma unroll
                for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(xw + i * 4) = make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float v = DGRAD ? __shfl_up_sync(0xffffffffu, __uint_as_float(r[i]), 1) : __shfl_down_sync(0xffffffffu, __uint_as_float(r[i]), 1);
                av[i] = (lane == (DGRAD ? 0 : 31)) ? 0.f : v;
            }
            // ---- D2 (dx = 2): row p +- 2
            tmem_ld32(taddr + 64, r);
            if (lane == (DGRAD ? 31 : 0)) {
#pragma unroll
                for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(xw + 32 + i * 4) = make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
            }
            if (lane == (DGRAD ? 30 : 1)) {
#pragma unroll
                for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(xw + 64 + i * 4) = make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float v = DGRAD ? __shfl_up_sync(0xffffffffu, __uint_as_float(r[i]), 2) : __shfl_down_sync(0xffffffffu, __uint_as_float(r[i]), 2);
                av[i] += (DGRAD ? lane < 2 : lane >= 30) ? 0.f : v;
            }
            // ---- D0
            tmem_ld32(taddr, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) av[i] += __uint_as_float(r[i]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_tempty + 8 * acc);      // accumulator stage drained by this warp
            const long long e2 = edbg ? clock64() : 0;
            // ---- the warp's edge rows: neighbours' accumulators through the exchange slots
            asm volatile("bar.sync %0, %1;" ::"r"(2 + egroup), "n"(kEpiWarps * 32) : "memory");
            const long long e3 = edbg ? clock64() : 0;
            if (!DGRAD) {
                if (ew + 1 < kEpiWarps && lane >= 30) {
                    const float* xo = xch + (size_t)(par * kEpiAll + warp + 1) * kXch;
                    if (lane == 31) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) av[i] += xo[i] + xo[64 + i];
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) av[i] += xo[32 + i];
                    }
                }
            } else {
                if (ew > 0 && lane < 2) {
                    const float* xo = xch + (size_t)(par * kEpiAll + warp - 1) * kXch;
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) av[i] += xo[i] + xo[64 + i];
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) av[i] += xo[32 + i];
                    }
                }
            }
            if (inside) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float v0 = av[c * 8 + q * 2], v1 = av[c * 8 + q * 2 + 1];
                        if (!DGRAD) {
                            v0 = valid ? fmaxf(fmaf(v0, scale, bias[c * 8 + q * 2]), 0.f) : 0.f;
                            v1 = valid ? fmaxf(fmaf(v1, scale, bias[c * 8 + q * 2 + 1]), 0.f) : 0.f;
                        } else {
                            const uint32_t m = q == 0 ? xm[c].x : (q == 1 ? xm[c].y : (q == 2 ? xm[c].z : xm[c].w));
                            const float2 xv = unpack_bf16x2(m);
                            v0 = xv.x > 0.f ? v0 : 0.f;
                            v1 = xv.y > 0.f ? v1 : 0.f;
                        }
                        w[q] = pack_bf16x2(v0, v1);
                    }
                    *reinterpret_cast<uint4*>(out + o + c * plane) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            acc_phase ^= 1;
            if (edbg) { const long long e4 = clock64(); e_wait += e1 - e0; e_tm += e2 - e1; e_bar += e3 - e2; e_rest += e4 - e3; }
        }
        if (edbg && tid == 0) {
            long long* d = g_tc_dbg[blockIdx.x];
            d[6] = e_wait; d[7] = e_tm + e_bar + e_rest; d[15] = e_tm; d[16] = e_bar; d[17] = e_rest;
        }
    } else if (warp < kEpiAll + kMmaWarps) {
        // ================= MMA issuers: warp mw takes the CTA's local tiles mw, mw + 2, ... into accumulator stage mw
        const int mw = warp - kEpiAll;
        uint32_t stage = (uint32_t)mw, phase = 0, acc_phase = 0;
        const uint32_t acc = (uint32_t)mw;
        const uint64_t a_hi = make_desc(0, PS, 128), b_hi = make_desc(0, 96 * 16, 128);
        uint32_t a_off[3 * KS];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
                a_off[dy * KS + ks] = ((uint32_t)(2 * ks) * PS + (uint32_t)((DGRAD ? 2 - dy : dy) * g.pitch) * 16u) >> 4;
        const int valid_pos = g.Hv * g.pitch;
        const bool dbg = (g.debug & 64) != 0;
        long long c_te = 0, c_fu = 0, c_is = 0, c_n = 0;
        const long long c_start = clock64();
        for (int j = mw;; j += kMmaWarps) {
            const long long c1 = dbg ? clock64() : 0;
            mbar_wait(s_full + 8 * stage, phase);
            const long long c2 = dbg ? clock64() : 0;
            const int tile = ring_tile((uint32_t)ld_volatile_s32(smem + 512 + 4 * (j & (kRing - 1))));
            if (tile < 0) break;
            const TileLoc tl = tc_locate(sg, g.tiles_per_sample, g.inv_tps, tile);
            // the second sub-tile holds no valid position when it starts at or beyond the last valid row
            const int nsub = (tl.t * kT96Out + tile_shift + 128 >= valid_pos) ? 1 : kTcSub;
            const uint32_t w16 = (s_w + (uint32_t)tc_pick(sg.wsel, tl.seg) * kW96Bytes) >> 4;
            const uint32_t slab16 = (s_slab0 + stage * slab_bytes) >> 4;
            mbar_wait(s_tempty + 8 * acc, acc_phase ^ 1);
            const long long c3 = dbg ? clock64() : 0;
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int s2 = 0; s2 < kTcSub; ++s2) {
                    if (s2 >= nsub) break;
                    const uint32_t d = tmem_base + acc * (kTcSub * 96) + (uint32_t)(s2 * 96);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint64_t ad = a_hi | (uint64_t)((slab16 + (uint32_t)(s2 * 128) + a_off[dy * KS + ks]) & 0x3FFFu);
                            const uint64_t bd = b_hi | (uint64_t)((w16 + (uint32_t)((dy * CH + 2 * ks) * 96)) & 0x3FFFu);
                            umma_bf16_rt(d, ad, bd, kIdesc96, (dy | ks) ? 1u : 0u);
                        }
                    }
                }
                umma_commit(s_empty + 8 * stage);
                umma_commit(s_tfull + 8 * acc);
            }
            __syncwarp();
            if (dbg) { c_fu += c2 - c1; c_te += c3 - c2; c_is += clock64() - c3; ++c_n; }
            stage += (uint32_t)kMmaWarps;
            if (stage >= (uint32_t)stages) { stage -= (uint32_t)stages; phase ^= 1; }
            acc_phase ^= 1;
        }
        if (dbg && lane == 0 && mw == 0) {
            long long* d = g_tc_dbg[blockIdx.x];
            d[0] = c_te; d[1] = c_fu; d[2] = c_is; d[4] = clock64() - c_start; d[5] = c_n;
        }
    } else {
        const int one = ld_volatile_s32(smem + 640 + 4 * lane);
        if (lane == 0) {
            // ================= producer: identical to k_conv_tc's, tiles 254 positions apart
            uint32_t stage = 0, phase = 0;
            const uint32_t bytes = (uint32_t)g.slab_rows * 16u;
            int* const ctr = g_tc_ctr + 2 * g.ctr_slot;
            const int G = (int)gridDim.x;
            int q[4];
            q[0] = (int)blockIdx.x - 2 * G; q[1] = q[0] + G;
            q[2] = g.dynamic ? atom_add_nonuniform(ctr, one) : q[1] + G;
            q[3] = g.dynamic ? atom_add_nonuniform(ctr, one) : q[2] + G;
            int sentinels = 0, j = 0;
            while (sentinels < 2) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (sentinels >= 2) break;
                    const int tile = q[u] + 2 * G;
                    const bool real = tile < g.total_tiles;
                    mbar_wait(s_empty + 8 * stage, phase ^ 1);
                    *reinterpret_cast<volatile uint32_t*>(smem + 512 + 4 * (j & (kRing - 1))) = ring_entry(j, real ? tile : -1);
                    const uint32_t bar = s_full + 8 * stage;
                    if (!real) {
                        mbar_arrive(bar);
                    } else {
                        const TileLoc tl = tc_locate(sg, g.tiles_per_sample, g.inv_tps, tile);
                        const int p0 = tl.t * kT96Out + tile_shift;
                        const bf16* src = tc_pick(sg.in, tl.seg) + (long long)tl.b * in_sstride + (long long)(p0 + g.min_off) * 8;
                        const uint32_t dst = s_slab0 + stage * slab_bytes;
                        mbar_expect_tx(bar, bytes * CH);
#pragma unroll
                        for (int c = 0; c < CH; ++c) bulk_g2s(dst + c * PS, src + c * plane, bytes, bar);
                    }
                    ++j;
                    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
                    if (!real) ++sentinels;
                    else q[u] = g.dynamic ? atom_add_nonuniform(ctr, one) : tile + 2 * G;
                }
            }
            if (g.dynamic) {
                __threadfence();
                if (atomicAdd(ctr + 1, 1) == G - 1) { ctr[0] = 0; ctr[1] = 0; __threadfence(); }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- N = 64 + 32: two horizontal taps side by side, the third by K accumulation
// What bounds k_conv_tc<32,9> is the A-operand stream (40 clk per N = 32 MMA, 16 of them tensor work); the N = 96
// kernel below cut the stream by 2.1x but its epilogue -- three TMEM loads, 64 shuffles, two exchanged edge rows per
// warp and a group barrier per tile, ~650 instructions per warp per tile on two accumulator stages -- is
// issue-bound at ~2900 clk per tile (measured: gpurun r03d, profiles/r03_conv_n96_roles.txt) against 1600 for the
// N = 32 kernel.  The middle ground: per (dy, K step)
//     MMA 1 (N = 64): D[p][0:32]  += A[p + dy*pitch]     . W(dy,0)      D[p][32:64] += A[p + dy*pitch] . W(dy,1)
//     MMA 2 (N = 32): D[p][0:32]  += A[p + dy*pitch + 2] . W(dy,2)
// 12 MMAs of 48 + 40 clk per 128 positions = 528 clk (720 for nine N = 32 taps), and ONE shifted add in the epilogue,
//     out[p] = D[p][0:32] + D[p + 1][32:64]            (dgrad: transposed weights, A shifts 2 / 2 / 0 rows, D[p - 1])
// i.e. two TMEM loads, 32 shuffles and one exchanged edge row per warp; 64 accumulator columns per sub-tile leave
// room for the four accumulator stages of k_conv_tc (512 TMEM columns), so MMAs and epilogue stay decoupled.
// Tiles overlap by one row (255 outputs per 256-row tile).  Pipeline roles, tile scheduler and slab ring are
// k_conv_tc's.
constexpr int kT64Out = 255;
constexpr uint32_t kIdesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr int kXch64 = 32;                                    // floats per warp slot: one accumulator row of D[.][32:64]
constexpr int kSmemHdr64 = 768 + 2 * kEpiAll * kXch64 * 4;    // header + exchange slots [tile parity of the group][warp][32]
constexpr uint32_t kW64Dy = 6144;                             // per dy: B1 [4 k chunks][64 n][8 k] (4096 B) | B2 [4][32][8] (2048 B)
constexpr uint32_t kW64Bytes = 3 * kW64Dy;

template <bool DGRAD>
__global__ void __launch_bounds__(kTcThreads, 1)
k_conv_tc64(const __grid_constant__ TcSegs sg, long long in_sstride, float scale, const bf16* __restrict__ relu_src,
            long long out_sstride, TcGeom g) {
    constexpr int CH = 4, KS = 2, TM = kTcSub * 128;
    constexpr uint32_t TMEM_COLS = kAccStages * kTcSub * 64;      // 512
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_full = s_base, s_empty = s_base + 64, s_tfull = s_base + 128, s_tempty = s_base + 192;
    const uint32_t s_tptr = s_base + 576;
    const uint32_t s_w = s_base + kSmemHdr64;
    const uint32_t PS = (uint32_t)g.plane_bytes;
    const uint32_t slab_bytes = CH * PS;
    const uint32_t s_slab0 = s_w + (uint32_t)sg.nw * kW64Bytes;
    const int stages = g.stages;
    const long long plane = (long long)g.S * 8;
    const int tile_shift = DGRAD ? -1 : 0;                    // position of a tile's row 0: t * 255 + tile_shift

    if (tid == 0) {
        for (int i = 0; i < kRing; ++i) reinterpret_cast<volatile uint32_t*>(smem + 512)[i] = 0xFFFFFFFFu;
        for (int i = 0; i < 32; ++i) reinterpret_cast<volatile int*>(smem + 640)[i] = 1;
        for (int i = 0; i < stages; ++i) {
            mbar_init(s_full + 8 * i, 1);
            mbar_init(s_empty + 8 * i, 1);
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(s_tfull + 8 * i, 1);
            mbar_init(s_tempty + 8 * i, kEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tptr), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + 576);
    pdl_grid_sync();
    constexpr int kStageThreads = kTcThreads - 32;
    if (warp != kEpiAll + kMmaWarps) {
        // weights per dy: B1[n = dx*32 + (co | ci)][k] for dx = 0, 1 and B2[n][k] for dx = 2, K-major [k chunk][n][8 k]
        for (int ws = 0; ws < sg.nw; ++ws) {
            const bf16* __restrict__ wts = ws ? sg.wts[1] : sg.wts[0];            // global: [tap = dy*3 + dx][co][ci]
            if (!DGRAD) {
                for (int i = tid; i < 9 * 32 * CH; i += kStageThreads) {
                    const int t = i / (32 * CH), rem = i - t * (32 * CH), n = rem / CH, kc = rem - n * CH;
                    const int dy = t / 3, dx = t - dy * 3;
                    const uint32_t off = dx < 2 ? (uint32_t)((kc * 64 + dx * 32 + n) * 16) : 4096u + (uint32_t)((kc * 32 + n) * 16);
                    cp_async16(s_w + ws * kW64Bytes + dy * kW64Dy + off, wts + ((t * 32 + n) * 32 + kc * 8), 16);
                }
                if (tid < 32) reinterpret_cast<float*>(smem + 256 + ws * 128)[tid] = (ws ? sg.bias[1] : sg.bias[0])[tid];
            } else {
                for (int i = tid; i < 9 * 32 * 32; i += kStageThreads) {
                    const int t = i / 1024, rem = i - t * 1024, co = rem >> 5, ci = rem & 31;
                    const int dy = t / 3, dx = t - dy * 3;
                    const int kc = co >> 3, k8 = co & 7;
                    const uint32_t off = dx < 2 ? (uint32_t)(((kc * 64 + dx * 32 + ci) * 8 + k8) * 2)
                                                : 4096u + (uint32_t)(((kc * 32 + ci) * 8 + k8) * 2);
                    *reinterpret_cast<bf16*>(smem + kSmemHdr64 + ws * kW64Bytes + dy * kW64Dy + off) = wts[i];
                }
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(kStageThreads) : "memory");
    }

    if (warp < kEpiAll) {
        // ================= epilogue
        const int egroup = warp / kEpiWarps, ew = warp % kEpiWarps;
        const int sub = ew >> 2, quarter = ew & 3;
        const int row_in_tile = sub * 128 + quarter * 32 + lane;
        const bool row_out = DGRAD ? (row_in_tile >= 1) : (row_in_tile < kT64Out);      // rows this tile stores
        const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sub * 64);
        float* const xch = reinterpret_cast<float*>(smem + 768);
        uint32_t acc = (uint32_t)egroup, acc_phase = 0;
        auto load_mask = [&](int tile_, uint4 (&dst)[4]) {          // dgrad launches have one segment
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[c] = make_uint4(0u, 0u, 0u, 0u);
            const int b_ = tc_div(tile_, g.inv_tps);
            const int p_ = (tile_ - b_ * g.tiles_per_sample) * kT64Out + tile_shift + row_in_tile;
            if (p_ >= 0 && p_ < g.S && row_out) {
                const int y_ = tc_div(p_, g.inv_pitch), x_ = p_ - y_ * g.pitch;
                if (y_ < g.Hv && x_ < g.Wv) {
                    const long long o_ = (long long)b_ * out_sstride + (long long)p_ * 8;
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[c] = *reinterpret_cast<const uint4*>(relu_src + o_ + c * plane);
                }
            }
        };
        int par = 0;
        for (int j = egroup;; j += kEpiGroups, par ^= 1) {
            uint32_t ent = (uint32_t)ld_volatile_s32(smem + 512 + 4 * (j & (kRing - 1)));
            if (!ring_ready(ent, j)) {
                const long long t0 = clock64();
                while (!ring_ready(ent = (uint32_t)ld_volatile_s32(smem + 512 + 4 * (j & (kRing - 1))), j))
                    if (clock64() - t0 > (1ll << 31)) __trap();
            }
            const int tile = ring_tile(ent);
            if (tile < 0) break;
            const TileLoc tl = tc_locate(sg, g.tiles_per_sample, g.inv_tps, tile);
            const int p = tl.t * kT64Out + tile_shift + row_in_tile;
            const bool inside = p >= 0 && p < g.S && row_out;
            int y = 0, x = 0;
            if (inside) { y = tc_div(p, g.inv_pitch); x = p - y * g.pitch; }
            const bool valid = inside && (y < g.Hv) && (x < g.Wv);
            bf16* __restrict__ out = tc_pick(sg.out, tl.seg);
            const long long o = (long long)tl.b * out_sstride + (long long)p * 8;   // + c*plane
            const float4* bias4 = reinterpret_cast<const float4*>(smem + 256 + (DGRAD ? 0 : tc_pick(sg.wsel, tl.seg)) * 128);
            uint4 xm[4];
            if (DGRAD) load_mask(tile, xm);             // requested before the accumulator wait: its latency overlaps this tile's MMAs
            float* const xw = xch + (size_t)(par * kEpiAll + warp) * kXch64;
            mbar_wait(s_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = taddr0 + acc * (kTcSub * 64);
            float av[32];
            uint32_t r[32];
            // ---- D[.][32:64]: the row next to this one (forward: p + 1, dgrad: p - 1)
            tmem_ld32(taddr + 32, r);
            if (lane == (DGRAD ? 31 : 0)) {             // the neighbouring warp's edge lane needs this row
#pragma unroll
                for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(xw + i * 4) = make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i)
                av[i] = DGRAD ? __shfl_up_sync(0xffffffffu, __uint_as_float(r[i]), 1) : __shfl_down_sync(0xffffffffu, __uint_as_float(r[i]), 1);
            // ---- D[.][0:32]
            tmem_ld32(taddr, r);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_tempty + 8 * acc);      // accumulator stage drained by this warp
            const bool edge = lane == (DGRAD ? 0 : 31);          // its shuffled value is its own row: replaced below
            asm volatile("bar.sync %0, %1;" ::"r"(2 + egroup), "n"(kEpiWarps * 32) : "memory");
            if (edge) {
                const bool has = DGRAD ? (ew > 0) : (ew + 1 < kEpiWarps);       // the tile's outermost row is not stored
                const float4* xo = reinterpret_cast<const float4*>(xch + (size_t)(par * kEpiAll + warp + (DGRAD ? -1 : 1)) * kXch64);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v = has ? xo[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    av[i * 4] = v.x; av[i * 4 + 1] = v.y; av[i * 4 + 2] = v.z; av[i * 4 + 3] = v.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) av[i] += __uint_as_float(r[i]);
            if (inside) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t w[4];
                    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                    if (!DGRAD) { b0 = bias4[c * 2]; b1 = bias4[c * 2 + 1]; }
                    const float bq[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float v0 = av[c * 8 + q * 2], v1 = av[c * 8 + q * 2 + 1];
                        if (!DGRAD) {
                            v0 = valid ? fmaxf(fmaf(v0, scale, bq[q * 2]), 0.f) : 0.f;
                            v1 = valid ? fmaxf(fmaf(v1, scale, bq[q * 2 + 1]), 0.f) : 0.f;
                        } else {
                            const uint32_t m = q == 0 ? xm[c].x : (q == 1 ? xm[c].y : (q == 2 ? xm[c].z : xm[c].w));
                            const float2 xv = unpack_bf16x2(m);
                            v0 = xv.x > 0.f ? v0 : 0.f;
                            v1 = xv.y > 0.f ? v1 : 0.f;
                        }
                        w[q] = pack_bf16x2(v0, v1);
                    }
                    *reinterpret_cast<uint4*>(out + o + c * plane) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            acc += kEpiGroups;
            if (acc >= kAccStages) { acc -= kAccStages; acc_phase ^= 1; }
        }
    } else if (warp < kEpiAll + kMmaWarps) {
        // ================= MMA issuers: warp mw takes the CTA's local tiles mw, mw + 2, ...
        const int mw = warp - kEpiAll;
        uint32_t stage = (uint32_t)mw, phase = 0, acc = (uint32_t)mw, acc_phase = 0;
        const uint64_t a_hi = make_desc(0, PS, 128), b1_hi = make_desc(0, 64 * 16, 128), b2_hi = make_desc(0, 32 * 16, 128);
        // A row shifts (relative to the slab start): MMA 1 / MMA 2 per dy
        uint32_t a1_off[3 * KS], a2_off[3 * KS];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int base = (DGRAD ? 2 - dy : dy) * g.pitch;
                a1_off[dy * KS + ks] = ((uint32_t)(2 * ks) * PS + (uint32_t)(base + (DGRAD ? 2 : 0)) * 16u) >> 4;
                a2_off[dy * KS + ks] = ((uint32_t)(2 * ks) * PS + (uint32_t)(base + (DGRAD ? 0 : 2)) * 16u) >> 4;
            }
        const int valid_pos = g.Hv * g.pitch;
        for (int j = mw;; j += kMmaWarps) {
            mbar_wait(s_full + 8 * stage, phase);
            const int tile = ring_tile((uint32_t)ld_volatile_s32(smem + 512 + 4 * (j & (kRing - 1))));
            if (tile < 0) break;
            const TileLoc tl = tc_locate(sg, g.tiles_per_sample, g.inv_tps, tile);
            const int nsub = (tl.t * kT64Out + tile_shift + 128 >= valid_pos) ? 1 : kTcSub;
            const uint32_t w16 = (s_w + (uint32_t)tc_pick(sg.wsel, tl.seg) * kW64Bytes) >> 4;
            const uint32_t slab16 = (s_slab0 + stage * slab_bytes) >> 4;
            mbar_wait(s_tempty + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int s2 = 0; s2 < kTcSub; ++s2) {
                    if (s2 >= nsub) break;
                    const uint32_t d = tmem_base + acc * (kTcSub * 64) + (uint32_t)(s2 * 64);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint32_t wdy = w16 + (uint32_t)(dy * (kW64Dy >> 4));
                            const uint64_t ad1 = a_hi | (uint64_t)((slab16 + (uint32_t)(s2 * 128) + a1_off[dy * KS + ks]) & 0x3FFFu);
                            const uint64_t bd1 = b1_hi | (uint64_t)((wdy + (uint32_t)(2 * ks * 64)) & 0x3FFFu);
                            umma_bf16_rt(d, ad1, bd1, kIdesc64, (dy | ks) ? 1u : 0u);
                            const uint64_t ad2 = a_hi | (uint64_t)((slab16 + (uint32_t)(s2 * 128) + a2_off[dy * KS + ks]) & 0x3FFFu);
                            const uint64_t bd2 = b2_hi | (uint64_t)((wdy + 256u + (uint32_t)(2 * ks * 32)) & 0x3FFFu);
                            umma_bf16_rt(d, ad2, bd2, kIdesc, 1u);
                        }
                    }
                }
                umma_commit(s_empty + 8 * stage);
                umma_commit(s_tfull + 8 * acc);
            }
            __syncwarp();
            stage += (uint32_t)kMmaWarps;
            if (stage >= (uint32_t)stages) { stage -= (uint32_t)stages; phase ^= 1; }
            acc += (uint32_t)kMmaWarps;
            if (acc >= kAccStages) { acc -= kAccStages; acc_phase ^= 1; }
        }
    } else {
        const int one = ld_volatile_s32(smem + 640 + 4 * lane);
        if (lane == 0) {
            // ================= producer: k_conv_tc's, tiles 255 positions apart
            uint32_t stage = 0, phase = 0;
            const uint32_t bytes = (uint32_t)g.slab_rows * 16u;
            int* const ctr = g_tc_ctr + 2 * g.ctr_slot;
            const int G = (int)gridDim.x;
            int q[4];
            q[0] = (int)blockIdx.x - 2 * G; q[1] = q[0] + G;
            q[2] = g.dynamic ? atom_add_nonuniform(ctr, one) : q[1] + G;
            q[3] = g.dynamic ? atom_add_nonuniform(ctr, one) : q[2] + G;
            int sentinels = 0, j = 0;
            while (sentinels < 2) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (sentinels >= 2) break;
                    const int tile = q[u] + 2 * G;
                    const bool real = tile < g.total_tiles;
                    mbar_wait(s_empty + 8 * stage, phase ^ 1);
                    *reinterpret_cast<volatile uint32_t*>(smem + 512 + 4 * (j & (kRing - 1))) = ring_entry(j, real ? tile : -1);
                    const uint32_t bar = s_full + 8 * stage;
                    if (!real) {
                        mbar_arrive(bar);
                    } else {
                        const TileLoc tl = tc_locate(sg, g.tiles_per_sample, g.inv_tps, tile);
                        const int p0 = tl.t * kT64Out + tile_shift;
                        const bf16* src = tc_pick(sg.in, tl.seg) + (long long)tl.b * in_sstride + (long long)(p0 + g.min_off) * 8;
                        const uint32_t dst = s_slab0 + stage * slab_bytes;
                        mbar_expect_tx(bar, bytes * CH);
#pragma unroll
                        for (int c = 0; c < CH; ++c) bulk_g2s(dst + c * PS, src + c * plane, bytes, bar);
                    }
                    ++j;
                    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
                    if (!real) ++sentinels;
                    else q[u] = g.dynamic ? atom_add_nonuniform(ctr, one) : tile + 2 * G;
                }
            }
            if (g.dynamic) {
                __threadfence();
                if (atomicAdd(ctr + 1, 1) == G - 1) { ctr[0] = 0; ctr[1] = 0; __threadfence(); }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- layers 2..4 fused: one sample resident in shared memory
// k_conv_tc96 still writes every layer's activations to HBM and the next layer's launch reads them back: per
// encoder pass 428 MB of the 616 MB the conv stack moves.  Here a CTA keeps ONE sample's activation map in
// shared memory (crop geometry: 2552 positions x 64 B = 163 KB) and runs conv-2, conv-3 and conv-4 over it IN
// PLACE: a 3x3 valid conv only looks "forward" in the flattened position index (out[p] reads in[p .. p + 2*pitch
// + 2]), tiles are processed in increasing position order, so tile t's output can overwrite positions
// [254 t, 254 t + 254) as soon as the MMAs of tiles t-1 and t have read them.  HBM traffic per pass: act0 in, act3
// out (+ act1 / act2 out for the passes whose backward needs them).  The weights of all three layers of one
// weight set stay resident next to the sample (55 KB, pre-packed by curla_pack_shadows in the N = 96 operand
// layout, one bulk copy).
//
// Pipeline (warp roles as in k_conv_tc96).  The buffer is a ring of T slots of 254 positions:
//   full[t]   producer -> issuers   slot t holds the sample's conv-1 output (4 bulk copies, one per channel plane)
//   wr[t]     epilogue -> issuers   slot t holds this layer's output (8 warp arrivals, after fence.proxy.async)
//   empty[t]  issuers  -> producer  conv-4's MMAs of tile t are complete: slot t may receive the NEXT sample --
//                                   so the next sample streams in behind the last layer, slot by slot
//   tfull / tempty[2]               accumulator stages, (issuer warp w, stage w, epilogue group w) as in k_conv_tc96
// Tile (layer l, slot t) reads slots t and t+1: its issuer waits for full / wr of both.  Nothing relies on the
// order in which the two issuer warps' MMAs execute: before an epilogue group overwrites slot t it waits until
// the OTHER group has drained the accumulators of the previous tile (a monotonic counter in shared memory), whose
// MMAs are the only other readers of slot t; the producer takes empty[0], empty[1], ... in order.
struct StackSegs {
    const bf16* in[3];           // conv-1 output of each pass
    bf16* out[3][3];             // [pass][layer]: conv-2 / conv-3 output (NULL: not kept), conv-4 output
    const bf16* w96[2];          // packed weights of the (at most) two weight sets: [3 layers][3 dy][4 k chunks][96 n][8 k]
    const float* bias[2][3];
    int item_end[3];             // items (samples) [item_end[s-1], item_end[s]) belong to pass s
    int wsel[3];
};
struct StackGeom {
    int pitch, S, rows, T;       // buffer rows (positions), slots
    int Hv[3], Wv[3];
    int total_items;
    float inv_pitch;
};
constexpr int kStackHdr = 768;           // barriers @0 (full[16] @0, empty[16] @128, wr[16] @256, tfull[2] @384, tempty[2] @400, wfull @416),
                                         // tmem ptr @424, drained[2] @432
constexpr int kStackXch = 2 * kEpiAll * kXch * 4;
constexpr uint32_t kStackWBytes = 3 * kW96Bytes;
constexpr int kStackMaxT = 16;

__global__ void __launch_bounds__(kTcThreads, 1)
k_conv_stack96(const __grid_constant__ StackSegs sg, long long sstride, StackGeom g) {
    constexpr int KS = 2;
    constexpr uint32_t TMEM_COLS = 512;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_full = s_base, s_empty = s_base + 128, s_wr = s_base + 256, s_tfull = s_base + 384, s_tempty = s_base + 400;
    const uint32_t s_wfull = s_base + 416, s_tptr = s_base + 424;
    volatile int* const drained = reinterpret_cast<volatile int*>(smem + 432);
    const uint32_t s_w = s_base + kStackHdr + kStackXch;
    const uint32_t s_buf = s_w + kStackWBytes;
    const uint32_t PS = (uint32_t)g.rows * 16u;
    const long long plane = (long long)g.S * 8;
    const int T = g.T;
    // this CTA's contiguous range of samples
    const int i0 = (int)(((long long)blockIdx.x * g.total_items) / gridDim.x);
    const int i1 = (int)(((long long)(blockIdx.x + 1) * g.total_items) / gridDim.x);

    if (tid == 0) {
        for (int i = 0; i < kStackMaxT; ++i) {
            mbar_init(s_full + 8 * i, 1);
            mbar_init(s_empty + 8 * i, 1);
            mbar_init(s_wr + 8 * i, kEpiWarps);
        }
        for (int i = 0; i < kAcc96; ++i) {
            mbar_init(s_tfull + 8 * i, 1);
            mbar_init(s_tempty + 8 * i, kEpiWarps);
        }
        mbar_init(s_wfull, 1);
        drained[0] = 0; drained[1] = 0;
        fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tptr), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + 424);
    pdl_grid_sync();

    auto seg_of = [&](int item) { return item >= sg.item_end[1] ? 2 : (item >= sg.item_end[0] ? 1 : 0); };
    auto seg_start = [&](int s) { return s == 0 ? 0 : (s == 1 ? sg.item_end[0] : sg.item_end[1]); };

    if (warp < kEpiAll) {
        // ================= epilogue: group eg drains the tiles with t & 1 == eg
        const int eg = warp / kEpiWarps, ew = warp % kEpiWarps;
        const int sub = ew >> 2, quarter = ew & 3;
        const int row_in_tile = sub * 128 + quarter * 32 + lane;
        const bool row_out = row_in_tile < kT96Out;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sub * 96) + (uint32_t)eg * (kTcSub * 96);
        float* const xch = reinterpret_cast<float*>(smem + kStackHdr);
        uint32_t acc_phase = 0;
        int par = 0;
        int j = eg;                                    // running tile index of this CTA: ((k*3 + l) * T + t); T is even
        for (int item = i0; item < i1; ++item) {
            const int seg = seg_of(item), b = item - seg_start(seg);
            const int ws = tc_pick(sg.wsel, seg);
            for (int l = 0; l < 3; ++l) {
                bf16* __restrict__ outg = seg == 0 ? sg.out[0][l] : (seg == 1 ? sg.out[1][l] : sg.out[2][l]);
                const float* __restrict__ bias = ws ? sg.bias[1][l] : sg.bias[0][l];
                const int Hv = g.Hv[l], Wv = g.Wv[l];
                for (int t = eg; t < T; t += kEpiGroups, j += kEpiGroups, par ^= 1) {
                    const int p = t * kT96Out + row_in_tile;
                    const bool inside = p < g.rows && row_out;
                    int y = 0, x = 0;
                    if (inside) { y = tc_div(p, g.inv_pitch); x = p - y * g.pitch; }
                    const bool valid = inside && (y < Hv) && (x < Wv);
                    float* const xw = xch + (size_t)(par * kEpiAll + warp) * kXch;
                    mbar_wait(s_tfull + 8 * eg, acc_phase);
                    tc_fence_after();
                    float av[32];
                    uint32_t r[32];
                    // ---- D1 (dx = 1): row p + 1
                    tmem_ld32(taddr + 32, r);
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(xw + i * 4) = make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float v = __shfl_down_sync(0xffffffffu, __uint_as_float(r[i]), 1);
                        av[i] = (lane == 31) ? 0.f : v;
                    }
                    // ---- D2 (dx = 2): row p + 2
                    tmem_ld32(taddr + 64, r);
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(xw + 32 + i * 4) = make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
                    }
                    if (lane == 1) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(xw + 64 + i * 4) = make_uint4(r[i * 4], r[i * 4 + 1], r[i * 4 + 2], r[i * 4 + 3]);
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float v = __shfl_down_sync(0xffffffffu, __uint_as_float(r[i]), 2);
                        av[i] += (lane >= 30) ? 0.f : v;
                    }
                    // ---- D0
                    tmem_ld32(taddr, r);
#pragma unroll
                    for (int i = 0; i < 32; ++i) av[i] += __uint_as_float(r[i]);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_tempty + 8 * eg);
                    asm volatile("bar.sync %0, %1;" ::"r"(2 + eg), "n"(kEpiWarps * 32) : "memory");
                    // every warp of the group has its accumulators in registers: this tile's MMAs are complete
                    if (ew == 0 && lane == 0) drained[eg] = (j >> 1) + 1;
                    if (ew + 1 < kEpiWarps && lane >= 30) {
                        const float* xo = xch + (size_t)(par * kEpiAll + warp + 1) * kXch;
                        if (lane == 31) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) av[i] += xo[i] + xo[64 + i];
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) av[i] += xo[32 + i];
                        }
                    }
                    uint4 ov[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t w[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 bq = __ldg(reinterpret_cast<const float2*>(bias) + c * 4 + q);
                            const float v0 = valid ? fmaxf(av[c * 8 + q * 2] + bq.x, 0.f) : 0.f;
                            const float v1 = valid ? fmaxf(av[c * 8 + q * 2 + 1] + bq.y, 0.f) : 0.f;
                            w[q] = pack_bf16x2(v0, v1);
                        }
                        ov[c] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    if (inside && outg && p < g.S) {
                        bf16* o = outg + (long long)b * sstride + (long long)p * 8;
#pragma unroll
                        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(o + c * plane) = ov[c];
                    }
                    if (l < 2) {
                        // in place: slot t is also read by the previous tile's MMAs (the other group's tile)
                        const int need = (j + 1) >> 1;
                        if (drained[eg ^ 1] < need) {
                            const long long t0 = clock64();
                            while (drained[eg ^ 1] < need)
                                if (clock64() - t0 > (1ll << 31)) __trap();
                        }
                        if (inside) {
                            const uint32_t so = s_buf + (uint32_t)p * 16u;
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(so + (uint32_t)c * PS), "r"(ov[c].x), "r"(ov[c].y),
                                             "r"(ov[c].z), "r"(ov[c].w) : "memory");
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(s_wr + 8 * t);
                    }
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp < kEpiAll + kMmaWarps) {
        // ================= MMA issuers: warp mw issues the tiles with t & 1 == mw into accumulator stage mw
        const int mw = warp - kEpiAll;
        uint32_t acc_phase = 0, wphase = 0;
        const uint64_t a_hi = make_desc(0, PS, 128), b_hi = make_desc(0, 96 * 16, 128);
        uint32_t a_off[3 * KS];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
                a_off[dy * KS + ks] = ((uint32_t)(2 * ks) * PS + (uint32_t)(dy * g.pitch) * 16u) >> 4;
        int cur_w = -1;
        for (int item = i0, k = 0; item < i1; ++item, ++k) {
            const int ws = tc_pick(sg.wsel, seg_of(item));
            if (ws != cur_w) {
                mbar_wait(s_wfull, wphase);
                wphase ^= 1;
                cur_w = ws;
            }
            for (int l = 0; l < 3; ++l) {
                const int valid_pos = g.Hv[l] * g.pitch;
                const uint32_t w16 = (s_w + (uint32_t)l * kW96Bytes) >> 4;
                for (int t = mw; t < T; t += kMmaWarps) {
                    const int t1 = t + 1 < T ? t + 1 : t;
                    if (l == 0) {
                        mbar_wait(s_full + 8 * t, (uint32_t)(k & 1));
                        mbar_wait(s_full + 8 * t1, (uint32_t)(k & 1));
                    } else {
                        mbar_wait(s_wr + 8 * t, (uint32_t)(l - 1));
                        mbar_wait(s_wr + 8 * t1, (uint32_t)(l - 1));
                    }
                    const int p0 = t * kT96Out;
                    const int nsub = p0 >= valid_pos ? 0 : (p0 + 128 >= valid_pos ? 1 : kTcSub);
                    const uint32_t slab16 = (s_buf + (uint32_t)p0 * 16u) >> 4;
                    mbar_wait(s_tempty + 8 * mw, acc_phase ^ 1);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int s2 = 0; s2 < kTcSub; ++s2) {
                            if (s2 >= nsub) break;
                            const uint32_t d = tmem_base + (uint32_t)mw * (kTcSub * 96) + (uint32_t)(s2 * 96);
#pragma unroll
                            for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                                for (int ks = 0; ks < KS; ++ks) {
                                    const uint64_t ad = a_hi | (uint64_t)((slab16 + (uint32_t)(s2 * 128) + a_off[dy * KS + ks]) & 0x3FFFu);
                                    const uint64_t bd = b_hi | (uint64_t)((w16 + (uint32_t)((dy * 4 + 2 * ks) * 96)) & 0x3FFFu);
                                    umma_bf16_rt(d, ad, bd, kIdesc96, (dy | ks) ? 1u : 0u);
                                }
                            }
                        }
                        if (l == 2) umma_commit(s_empty + 8 * t);      // the sample's last reader of slot t (with tile t-1, committed before)
                        umma_commit(s_tfull + 8 * mw);
                    }
                    __syncwarp();
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        if (lane == 0) {
            // ================= producer: weights (one bulk copy per weight-set change) and the sample, slot by slot
            int cur_w = -1;
            for (int item = i0, k = 0; item < i1; ++item, ++k) {
                const int seg = seg_of(item), b = item - seg_start(seg);
                const int ws = tc_pick(sg.wsel, seg);
                if (ws != cur_w) {
                    // every MMA that reads the old weights is complete once the last slot of the previous sample is released
                    // (each issuer's commits complete in its own order: its last tile's commit covers all of its MMAs)
                    if (k > 0) {
                        mbar_wait(s_empty + 8 * (T - 2), (uint32_t)((k - 1) & 1));
                        mbar_wait(s_empty + 8 * (T - 1), (uint32_t)((k - 1) & 1));
                    }
                    mbar_expect_tx(s_wfull, kStackWBytes);
                    bulk_g2s(s_w, ws ? sg.w96[1] : sg.w96[0], kStackWBytes, s_wfull);
                    cur_w = ws;
                }
                const bf16* src = tc_pick(sg.in, seg) + (long long)b * sstride;
                for (int t = 0; t < T; ++t) {
                    if (k > 0) mbar_wait(s_empty + 8 * t, (uint32_t)((k - 1) & 1));
                    const int p0 = t * kT96Out;
                    const int pend = (t + 1 < T && (t + 1) * kT96Out < g.rows) ? (t + 1) * kT96Out : g.rows;
                    const int np = pend - p0;
                    const uint32_t bar = s_full + 8 * t;
                    if (np <= 0) {                             // T is rounded up to even: a last slot past the buffer stays empty
                        mbar_arrive(bar);
                        continue;
                    }
                    mbar_expect_tx(bar, (uint32_t)np * 64u);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        bulk_g2s(s_buf + (uint32_t)c * PS + (uint32_t)p0 * 16u, src + c * plane + (long long)p0 * 8, (uint32_t)np * 16u, bar);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

static TcGeom make_tc_geom(int B, int pitch, int S, int Hv, int Wv, int span, int min_off) {
    TcGeom g;
    g.pitch = pitch; g.S = S; g.Hv = Hv; g.Wv = Wv;
    g.tiles_per_sample = cdiv((long long)Hv * pitch, kTcSub * 128);
    g.total_tiles = B * g.tiles_per_sample;
    g.slab_rows = kTcSub * 128 + span;
    g.plane_rows = (g.slab_rows + 7) / 8 * 8;
    g.min_off = min_off;
    g.stages = 2;
    g.dynamic = 0; g.ctr_slot = 0; g.zero_planes = 0;
    g.inv_tps = 1.0f / (float)g.tiles_per_sample;
    g.inv_pitch = 1.0f / (float)pitch;
    g.plane_bytes = g.plane_rows * 16;
    g.debug = 0;
    return g;
}

// Counter pairs are handed out per (device, stream): launches on one stream are serialised (the next
// kernel asks for tiles only after griddepcontrol.wait, i.e. after the previous one has re-armed its
// pair), so a stream cycles through its own group of four pairs and two engines on different streams
// or devices of one process never share one.  g_tc_ctr itself exists once per device.
static int next_ctr_slot(cudaStream_t stream) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, std::pair<int, unsigned>> groups;     // -> (group, sequence)
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = groups.find({dev, stream});
    if (it == groups.end()) {
        int used = 0;
        for (auto& kv : groups) if (kv.first.first == dev) ++used;
        it = groups.emplace(std::make_pair(dev, stream), std::make_pair(used % (kCtrSlots / 4), 0u)).first;
    }
    return it->second.first * 4 + (int)(it->second.second++ & 3u);
}

template <typename K>
static int tc_set_smem(K kern, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_last_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e)); return -1; }
    return 0;
}

template <int CP, int NTAPS, bool DGRAD>
static int launch_tc(const TcSegs& sg, long long in_sstride, float scale, const void* relu_src, long long out_sstride,
                     TcGeom g, const TcTaps& taps, cudaStream_t stream) {
    constexpr int CH = CP / 8;
    const size_t fixed = kSmemHdr + (size_t)sg.nw * NTAPS * CH * 512;
    { const char* e = getenv("CURLA_TC_DEBUG"); g.debug = e ? atoi(e) : 0; }
    g.plane_bytes = g.plane_rows * 16 + ((g.debug & 16) ? 64 : 0);
    const size_t slab = (size_t)CH * g.plane_bytes;
    // shared-memory budget of the slab ring (CURLA_TC_SMEM_KB, timing experiments: 64..225)
    size_t budget = 190 * 1024;
    { const char* e = getenv("CURLA_TC_SMEM_KB"); if (e && atoi(e) >= 64 && atoi(e) <= 225) budget = (size_t)atoi(e) * 1024; }
    int stages = (int)((budget - fixed) / slab);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) { set_last_error("conv_tc: pitch %d needs %zu B per slab stage", g.pitch, slab); return -1; }
    // Even ring depth: issuer warp 0 then only ever touches the even stages' barriers and issuer 1
    // the odd ones.  With an odd depth the two warps alternate on every full[] barrier, and a
    // parity wait for phase n+1 passes spuriously while phase n is still incomplete (one warp a
    // ring ahead of the other's bulk copies) -- observed as a trapped launch on the B200.
    stages &= ~1;
    g.stages = stages;
    { const char* e = getenv("CURLA_TC_STAGES"); if (e && atoi(e) >= 2 && atoi(e) <= stages) g.stages = atoi(e) & ~1; }
    const size_t smem = fixed + stages * slab;
    auto kern = k_conv_tc<CP, NTAPS, DGRAD>;
    if (tc_set_smem(kern, smem)) return -1;
    const int cap = conv_grid_cap();
    const int grid = g.total_tiles < cap ? g.total_tiles : cap;
    {   // dynamic tile hand-out once every CTA has more than its two round-robin tiles
        const char* e = getenv("CURLA_TC_STATIC");
        g.dynamic = (g.total_tiles > 2 * grid && !(e && e[0] == '1')) ? 1 : 0;
        g.ctr_slot = g.dynamic ? next_ctr_slot(stream) : 0;
    }
    launch_k(kern, dim3(grid), dim3(kTcThreads), smem, stream, sg, in_sstride, scale, (const bf16*)relu_src, out_sstride, g, taps);
    return 0;
}

// N = 96 variant: layers 2..4 forward and dgrad, CURLA_CONV_N96=1.  Off by default: measured on the B200 at the
// benchmarked batch it is SLOWER than the N = 32 kernel (conv forward 1.00 vs 0.71 ms per update, dgrad 0.42 vs
// 0.29: gpurun r03c) -- the MMAs of a tile shrink from 1440 to 672 clk, but every tile now pays three serialised
// TMEM loads, 64 shuffles and a named barrier among eight epilogue warps on only two accumulator stages.
static bool use_n96(int pitch, int Wv, bool dgrad) {
    const char* e = getenv("CURLA_CONV_N96");
    if (!(e && e[0] == '1')) return false;
    // forward: outputs whose horizontal taps would wrap into the next image row must be invalid columns
    return dgrad || Wv <= pitch - 2;
}

// N = 64 + 32 variant (k_conv_tc64): layers 2..4 forward and dgrad, CURLA_CONV_N64=1
static bool use_n64() {
    const char* e = getenv("CURLA_CONV_N64");
    return e && e[0] == '1';
}

template <bool DGRAD>
static int launch_tc64(const TcSegs& sg, long long in_sstride, float scale, const void* relu_src, long long out_sstride,
                       TcGeom g, cudaStream_t stream) {
    const size_t fixed = kSmemHdr64 + (size_t)sg.nw * kW64Bytes;
    g.debug = 0;
    g.plane_bytes = g.plane_rows * 16;
    const size_t slab = (size_t)4 * g.plane_bytes;
    size_t budget = 200 * 1024;
    { const char* e = getenv("CURLA_TC_SMEM_KB"); if (e && atoi(e) >= 64 && atoi(e) <= 225) budget = (size_t)atoi(e) * 1024; }
    int stages = (int)((budget - fixed) / slab);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) { set_last_error("conv_tc64: pitch %d needs %zu B per slab stage", g.pitch, slab); return -1; }
    stages &= ~1;                                   // even ring depth: see launch_tc
    g.stages = stages;
    const size_t smem = fixed + stages * slab;
    auto kern = k_conv_tc64<DGRAD>;
    if (tc_set_smem(kern, smem)) return -1;
    const int cap = conv_grid_cap();
    const int grid = g.total_tiles < cap ? g.total_tiles : cap;
    {
        const char* e = getenv("CURLA_TC_STATIC");
        g.dynamic = (g.total_tiles > 2 * grid && !(e && e[0] == '1')) ? 1 : 0;
        g.ctr_slot = g.dynamic ? next_ctr_slot(stream) : 0;
    }
    launch_k(kern, dim3(grid), dim3(kTcThreads), smem, stream, sg, in_sstride, scale, (const bf16*)relu_src, out_sstride, g);
    return 0;
}

static TcGeom make_tc64_geom(int pitch, int S, int Hv, int Wv, bool dgrad) {
    TcGeom g = make_tc_geom(1, pitch, S, Hv, Wv, 2 * pitch + 2, dgrad ? -(2 * pitch + 2) : 0);
    g.tiles_per_sample = cdiv((long long)Hv * pitch, kT64Out);
    g.total_tiles = g.tiles_per_sample;
    g.inv_tps = 1.0f / (float)g.tiles_per_sample;
    return g;
}

template <bool DGRAD>
static int launch_tc96(const TcSegs& sg, long long in_sstride, float scale, const void* relu_src, long long out_sstride,
                       TcGeom g, cudaStream_t stream) {
    const size_t fixed = kSmemHdr96 + (size_t)sg.nw * kW96Bytes;
    { const char* e = getenv("CURLA_TC_DEBUG"); g.debug = e ? atoi(e) : 0; }
    g.plane_bytes = g.plane_rows * 16;
    const size_t slab = (size_t)4 * g.plane_bytes;
    size_t budget = 200 * 1024;
    { const char* e = getenv("CURLA_TC_SMEM_KB"); if (e && atoi(e) >= 64 && atoi(e) <= 225) budget = (size_t)atoi(e) * 1024; }
    int stages = (int)((budget - fixed) / slab);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) { set_last_error("conv_tc96: pitch %d needs %zu B per slab stage", g.pitch, slab); return -1; }
    stages &= ~1;                                   // even ring depth: see launch_tc
    g.stages = stages;
    const size_t smem = fixed + stages * slab;
    auto kern = k_conv_tc96<DGRAD>;
    if (tc_set_smem(kern, smem)) return -1;
    const int cap = conv_grid_cap();
    const int grid = g.total_tiles < cap ? g.total_tiles : cap;
    {
        const char* e = getenv("CURLA_TC_STATIC");
        g.dynamic = (g.total_tiles > 2 * grid && !(e && e[0] == '1')) ? 1 : 0;
        g.ctr_slot = g.dynamic ? next_ctr_slot(stream) : 0;
    }
    launch_k(kern, dim3(grid), dim3(kTcThreads), smem, stream, sg, in_sstride, scale, (const bf16*)relu_src, out_sstride, g);
    return 0;
}

static TcGeom make_tc96_geom(int pitch, int S, int Hv, int Wv, bool dgrad) {
    TcGeom g = make_tc_geom(1, pitch, S, Hv, Wv, 2 * pitch, dgrad ? -2 * pitch : 0);
    g.tiles_per_sample = cdiv((long long)Hv * pitch, kT96Out);
    g.total_tiles = g.tiles_per_sample;
    g.inv_tps = 1.0f / (float)g.tiles_per_sample;
    return g;
}

// fills the segment table; total tiles -> g.total_tiles
static int make_segs(const curla_conv_seg* segs, int nseg, TcGeom& g, TcSegs& sg) {
    CURLA_CHECK(nseg >= 1 && nseg <= 3, "conv: 1..3 segments per launch (got %d)", nseg);
    memset(&sg, 0, sizeof(sg));
    int tiles = 0;
    for (int s = 0; s < nseg; ++s) {
        CURLA_CHECK(segs[s].B >= 1 && segs[s].in && segs[s].out && segs[s].wts, "conv: bad segment %d", s);
        sg.in[s] = (const bf16*)segs[s].in;
        sg.out[s] = (bf16*)segs[s].out;
        int w = -1;
        for (int k = 0; k < sg.nw; ++k) if (sg.wts[k] == (const bf16*)segs[s].wts && sg.bias[k] == segs[s].bias) w = k;
        if (w < 0) {
            CURLA_CHECK(sg.nw < 2, "conv: at most two distinct weight sets per launch");
            w = sg.nw++;
            sg.wts[w] = (const bf16*)segs[s].wts;
            sg.bias[w] = segs[s].bias;
        }
        sg.wsel[s] = w;
        tiles += segs[s].B * g.tiles_per_sample;
        sg.tile_end[s] = tiles;
    }
    CURLA_CHECK(tiles < (1 << 20) - 1, "conv: %d tiles in one launch (limit 2^20 - 2)", tiles);
    for (int s = nseg; s < 3; ++s) sg.tile_end[s] = tiles;
    g.total_tiles = tiles;
    return 0;
}


// Geometry of the fused launch; false when a sample (plus weights) does not fit one SM's shared memory or the
// N = 96 formulation does not apply (90 x 160 inputs: 225 KB per sample -- those run layer by layer).
static bool make_stack_geom(int pitch, int S, const int* Hv, const int* Wv, StackGeom& g, size_t& smem) {
    for (int l = 0; l < 3; ++l)
        if (Wv[l] > pitch - 2 || Hv[l] < 1 || Wv[l] < 1) return false;
    if (Hv[1] > Hv[0] || Hv[2] > Hv[1]) return false;
    const int valid0 = Hv[0] * pitch;
    int T = cdiv(valid0, kT96Out);
    T = (T + 1) & ~1;                                  // even: (issuer w, accumulator w, epilogue group w) keep their tile parity across layers
    if (T < 2 || T > kStackMaxT) return false;
    // rows the MMAs of the first fused layer read: its last non-empty sub-tile + 127 rows + two image rows below
    const int tl = (valid0 - 1) / kT96Out;
    const int last_sub = tl * kT96Out + ((tl * kT96Out + 128 < valid0) ? 128 : 0);
    int rows = last_sub + 128 + 2 * pitch;
    rows = (rows + 7) / 8 * 8;
    if (rows > S + 256) return false;                  // reads past a plane's S positions stay inside the buffers' padding (curla_conv_pad_rows)
    g.pitch = pitch; g.S = S; g.rows = rows; g.T = T;
    for (int l = 0; l < 3; ++l) { g.Hv[l] = Hv[l]; g.Wv[l] = Wv[l]; }
    g.inv_pitch = 1.0f / (float)pitch;
    g.total_items = 0;
    smem = (size_t)kStackHdr + kStackXch + kStackWBytes + (size_t)rows * 64;
    return smem <= 232448;
}
}  // namespace curla

using namespace curla;

// timing experiments: copies the CURLA_TC_DEBUG&64 counters of `n` CTAs (20 int64 each) to the host
extern "C" int curla_conv_debug_read(long long* out, int n) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_tc_dbg, sizeof(long long) * 20 * (n < 160 ? n : 160));
    CURLA_CHECK(e == cudaSuccess, "conv_debug_read: %s", cudaGetErrorString(e));
    return 0;
}

// Rows of zero padding the caller keeps in front of sample 0 and behind sample B-1 of every
// activation buffer handed to the conv kernels (slabs read past the tile on both sides).
extern "C" int curla_conv_pad_rows(int pitch) { return 128 * (kTcSub + 1) + 2 * pitch + 2; }

// layer 1: in = s2d bf16 [B][S][48]; layers 2..4: in = bf16 [B][S][32].  out bf16 [B][S][32].
// Hv/Wv = valid output dims of this layer.
extern "C" int curla_conv_fwd_multi(const curla_conv_seg* segs, int nseg, long long in_sstride, float scale,
                                    long long out_sstride, int pitch, int S, int Hv, int Wv, int first_layer,
                                    cudaStream_t stream) {
    TcTaps taps;
    TcSegs sg;
    if (first_layer) {
        for (int t = 0; t < 4; ++t) taps.off[t] = (t >> 1) * pitch + (t & 1);
        for (int t = 4; t < 9; ++t) taps.off[t] = 0;
        TcGeom g = make_tc_geom(1, pitch, S, Hv, Wv, pitch + 1, 0);
        if (first_layer > 1) {                    // number of real s2d channels: whole planes above it are zero
            CURLA_CHECK(first_layer <= 48, "conv_fwd: first_layer = %d real channels (max 48)", first_layer);
            g.zero_planes = (48 - first_layer) / 8;
        }
        if (make_segs(segs, nseg, g, sg)) return -1;
        if (launch_tc<48, 4, false>(sg, in_sstride, scale, nullptr, out_sstride, g, taps, stream)) return -1;
    } else if (use_n64()) {
        TcGeom g = make_tc64_geom(pitch, S, Hv, Wv, false);
        if (make_segs(segs, nseg, g, sg)) return -1;
        if (launch_tc64<false>(sg, in_sstride, scale, nullptr, out_sstride, g, stream)) return -1;
    } else if (use_n96(pitch, Wv, false)) {
        TcGeom g = make_tc96_geom(pitch, S, Hv, Wv, false);
        if (make_segs(segs, nseg, g, sg)) return -1;
        if (launch_tc96<false>(sg, in_sstride, scale, nullptr, out_sstride, g, stream)) return -1;
    } else {
        for (int t = 0; t < 9; ++t) taps.off[t] = (t / 3) * pitch + (t % 3);
        TcGeom g = make_tc_geom(1, pitch, S, Hv, Wv, 2 * pitch + 2, 0);
        if (make_segs(segs, nseg, g, sg)) return -1;
        if (launch_tc<32, 9, false>(sg, in_sstride, scale, nullptr, out_sstride, g, taps, stream)) return -1;
    }
    return check_launch("conv_fwd");
}

extern "C" int curla_conv_fwd(const void* in, long long in_sstride, const void* wts,
                              const float* bias, float scale, void* out, long long out_sstride,
                              int B, int pitch, int S, int Hv, int Wv, int first_layer,
                              cudaStream_t stream) {
    const curla_conv_seg seg = {in, wts, bias, out, B};
    return curla_conv_fwd_multi(&seg, 1, in_sstride, scale, out_sstride, pitch, S, Hv, Wv, first_layer, stream);
}

// dX[P] = relu'(X[P]) * sum_t dY[P - off_t] . Wt^T ; Hv/Wv = valid dims of X (this layer's INPUT).
extern "C" int curla_conv_dgrad(const void* dy, long long dy_sstride, const void* wts,
                                const void* x, void* dx, long long dx_sstride, int B, int pitch,
                                int S, int Hv, int Wv, cudaStream_t stream) {
    TcTaps taps;
    TcSegs sg;
    const curla_conv_seg seg = {dy, wts, nullptr, dx, B};
    if (use_n64()) {
        TcGeom g = make_tc64_geom(pitch, S, Hv, Wv, true);
        if (make_segs(&seg, 1, g, sg)) return -1;
        if (launch_tc64<true>(sg, dy_sstride, 1.f, x, dx_sstride, g, stream)) return -1;
        return check_launch("conv_dgrad");
    }
    if (use_n96(pitch, Wv, true)) {
        TcGeom g = make_tc96_geom(pitch, S, Hv, Wv, true);
        if (make_segs(&seg, 1, g, sg)) return -1;
        if (launch_tc96<true>(sg, dy_sstride, 1.f, x, dx_sstride, g, stream)) return -1;
        return check_launch("conv_dgrad");
    }
    for (int t = 0; t < 9; ++t) taps.off[t] = -((t / 3) * pitch + (t % 3));
    TcGeom g = make_tc_geom(1, pitch, S, Hv, Wv, 2 * pitch + 2, -(2 * pitch + 2));
    if (make_segs(&seg, 1, g, sg)) return -1;
    if (launch_tc<32, 9, true>(sg, dy_sstride, 1.f, x, dx_sstride, g, taps, stream)) return -1;
    return check_launch("conv_dgrad");
}

// ---- conv-2..4 fused (k_conv_stack96).  Hv / Wv: valid output dims of the three layers.
extern "C" int curla_conv_stack_fits(int pitch, int S, const int* Hv, const int* Wv) {
    const char* e = getenv("CURLA_CONV_FUSED");
    if (!(e && e[0] == '1')) return 0;                 // opt-in (experiment): built on the N = 96 epilogue, see use_n96
    StackGeom g;
    size_t smem = 0;
    return make_stack_geom(pitch, S, Hv, Wv, g, smem) ? 1 : 0;
}

extern "C" int curla_conv_stack_fwd(const curla_conv_stack_seg* segs, int nseg, long long sstride, int pitch, int S,
                                    const int* Hv, const int* Wv, cudaStream_t stream) {
    CURLA_CHECK(nseg >= 1 && nseg <= 3, "conv_stack: 1..3 passes per launch (got %d)", nseg);
    StackGeom g;
    size_t smem = 0;
    CURLA_CHECK(make_stack_geom(pitch, S, Hv, Wv, g, smem), "conv_stack: geometry (pitch %d, %d x %d) does not fit shared memory", pitch, Hv[0], Wv[0]);
    StackSegs sg;
    memset(&sg, 0, sizeof(sg));
    int items = 0, nw = 0;
    for (int s = 0; s < nseg; ++s) {
        CURLA_CHECK(segs[s].B >= 1 && segs[s].in && segs[s].w96 && segs[s].out[2], "conv_stack: bad pass %d", s);
        CURLA_CHECK((reinterpret_cast<uintptr_t>(segs[s].w96) & 15) == 0, "conv_stack: packed weights must be 16-byte aligned");
        sg.in[s] = (const bf16*)segs[s].in;
        for (int l = 0; l < 3; ++l) sg.out[s][l] = (bf16*)segs[s].out[l];
        int w = -1;
        for (int k = 0; k < nw; ++k) if (sg.w96[k] == (const bf16*)segs[s].w96) w = k;
        if (w < 0) {
            CURLA_CHECK(nw < 2, "conv_stack: at most two distinct weight sets per launch");
            w = nw++;
            sg.w96[w] = (const bf16*)segs[s].w96;
            for (int l = 0; l < 3; ++l) sg.bias[w][l] = segs[s].bias[l];
        }
        sg.wsel[s] = w;
        items += segs[s].B;
        sg.item_end[s] = items;
    }
    for (int s = nseg; s < 3; ++s) sg.item_end[s] = items;
    g.total_items = items;
    if (tc_set_smem(k_conv_stack96, smem)) return -1;
    const int cap = conv_grid_cap();
    const int grid = items < cap ? items : cap;
    launch_k(k_conv_stack96, dim3(grid), dim3(kTcThreads), smem, stream, sg, sstride, g);
    return check_launch("conv_stack");
}
