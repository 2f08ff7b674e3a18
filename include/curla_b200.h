/* curla_b200 -- C ABI of the B200-native CurlSacAgent.update hot path.
 *
 * The reference (paulvantieghem/curla) is pure Python/PyTorch and has no FFI; the
 * boundary it exposes for this path is the Python API of curl_sac.py / encoder.py /
 * utils.py / augmentations.py.  This header is the C ABI the Python host side
 * (the curla_b200 Python package, via ctypes) binds; every entry point cites the reference
 * code it replaces (paths relative to the reference repo root).  INTEGRATION.md shows the
 * ctypes stubs a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; device pointers unless stated; the caller owns
 *    every buffer (PyTorch allocations); nothing returned by the library is freed by
 *    the caller except handles via their *_destroy.
 *  - every call is asynchronous on `stream` and never synchronises.
 *  - return 0 on success, negative on error; curla_last_error() gives the message.
 *  - bf16 buffers are passed as void*.
 */
#ifndef CURLA_B200_H
#define CURLA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* curla_stream_t; /* == cudaStream_t */

const char* curla_last_error(void);
int curla_version(void);

/* ---- K1: replay gather + crop  (utils.py:147-166, augmentations.py:47-75) -------------
 * frames: uint8 [capacity][C][Hf][Wf]; idxs/h1/w1: int64 [B] (h1,w1 may be NULL = 0,
 * idxs may be NULL = arange).  out_f32: [B][C][H][W].                                   */
int curla_gather_crop_f32(const uint8_t* frames, int C, int Hf, int Wf, const int64_t* idxs,
                          const int64_t* h1, const int64_t* w1, int B, int H, int W,
                          float* out, curla_stream_t stream);
/* same gather, written as the bf16 space-to-depth channel planes the conv stack consumes:
 * out[b][ch/8][yb*Ws+xb][ch%8] with ch = c*4+sy*2+sx, CP channels (>= 4C, multiple of 8).
 * Planes that hold only padding channels (8j >= 4C) are NOT written: the caller provides a
 * zero-initialised buffer once (the engine's arenas are).                                */
int curla_gather_crop_s2d(const uint8_t* frames, int C, int Hf, int Wf, const int64_t* idxs,
                          const int64_t* h1, const int64_t* w1, int B, int H, int W, int CP,
                          long long out_sample_stride, void* out, curla_stream_t stream);
/* float NCHW observations (already augmented) -> s2d rows (encoder.py:77 input). */
int curla_f32_to_s2d(const float* obs, int C, int H, int W, int B, int CP,
                     long long out_sample_stride, void* out, curla_stream_t stream);
/* actions / rewards / not_dones rows (utils.py:163-165) */
/* ReplayBuffer.add (utils.py:120-128): vec = [action[na] | reward | not_done] -> ring row */
int curla_scatter_transition(const float* vec, int na, long long row, float* actions,
                             float* rewards, float* not_dones, curla_stream_t stream);
/* the whole add(): pinned host frames + staged vec -> ring row (H2D copies + scatter) */
int curla_replay_add(const uint8_t* obs_pinned, const uint8_t* next_pinned, long long frame_bytes,
                     const float* vec_pinned, float* vec_dev, int na, long long row, uint8_t* obses,
                     uint8_t* next_obses, float* actions, float* rewards, float* not_dones,
                     curla_stream_t stream);
/* up to three streams (obs / pos / next_obs of one sampled index block: utils.py:151-158) of the
 * same geometry in ONE launch, frames staged with cp.async.bulk; `rows` (optional) also gathers the
 * batch's action / reward / not_done rows (utils.py:163-165).  Bitwise identical to nseg
 * curla_gather_crop_s2d calls + three curla_gather_rows_f32 calls.                              */
typedef struct curla_gather_seg {
    const void* frames;              /* uint8 [capacity][C][Hf][Wf] */
    const int64_t* h1;               /* crop offsets of this stream (NULL = 0) */
    const int64_t* w1;
    void* out;                       /* bf16 s2d planes */
} curla_gather_seg;
typedef struct curla_gather_rows {
    const float* actions; const float* rewards; const float* not_dones;
    float* out_actions; float* out_rewards; float* out_not_dones;
    int action_dim;
} curla_gather_rows;
int curla_gather_crop_s2d_multi(const curla_gather_seg* segs, int nseg, const int64_t* idxs, int C,
                                int Hf, int Wf, int B, int H, int W, int CP,
                                long long out_sample_stride, const curla_gather_rows* rows,
                                curla_stream_t stream);
int curla_gather_rows_f32(const float* src, const int64_t* idxs, int B, int K, float* out,
                          curla_stream_t stream);

/* ---- K2/K3: device-side augmentations of the non-crop sample_cpc branch (utils.py:168-182)
 * x: float [n_images][3][H][W] in [0,255], modified in place; one image = one RGB frame of a
 * frame stack.  color_jiggle (augmentations.py:106-136): params NULL => per-image contrast
 * U[1-c,1+c], saturation U[1-s,1+s], hue U[-h,h], apply ~ Bernoulli(p) from Philox(seed,offset);
 * else params = float [4][n_images] {contrast, saturation, hue, apply}.  order: op codes
 * (1 contrast, 2 saturation, 3 hue, 0 brightness = no-op) 4 bits each, first op lowest.
 * noisy_cover (augmentations.py:172-205): cover3 = HOST float[3]; noise_in NULL => std*N(0,1). */
int curla_color_jiggle(float* x, int n_images, int H, int W, const float* params,
                       unsigned long long seed, unsigned long long offset, float contrast,
                       float saturation, float hue, float p, int order, float* params_out,
                       curla_stream_t stream);
int curla_noisy_cover(float* x, int n_images, int H, int W, int top, int bottom,
                      const float* cover3, float stdv, const float* noise_in,
                      unsigned long long seed, unsigned long long offset, curla_stream_t stream);

/* ---- K4-K6: conv stack as shifted GEMMs  (encoder.py:77-90 + autograd) ------------------
 * fwd/dgrad: tcgen05.mma + TMEM (conv_tc.cu); wgrad: split-K over positions (conv.cu).
 * Every activation buffer needs curla_conv_pad_rows(pitch) zero rows before sample 0 and
 * after sample B-1.                                                                       */
int curla_conv_pad_rows(int pitch);
int curla_conv_fwd(const void* in, long long in_sstride, const void* wts, const float* bias,
                   float scale, void* out, long long out_sstride, int B, int pitch, int S,
                   int Hv, int Wv, int first_layer, curla_stream_t stream);
int curla_conv_dgrad(const void* dy, long long dy_sstride, const void* wts, const void* x,
                     void* dx, long long dx_sstride, int B, int pitch, int S, int Hv, int Wv,
                     curla_stream_t stream);
/* One launch for up to three independent encoder passes of the same layer (F1
 * conv_theta(next_obs), F2 conv_target(next_obs), F3 conv_theta(obs) have no dependency on each
 * other, nor have the CURL anchor and key passes): segment s maps in[s] -> out[s] with its own
 * weights/bias (at most two DISTINCT weight pointers per call) and batch.  Same geometry,
 * strides and scale for all segments.  Bitwise identical to nseg curla_conv_fwd calls.
 * first_layer: 0 = layers 2..4; 1 = the space-to-depth first layer; > 1 = first layer whose input
 * has that many real channels (4*C): channel planes above it hold zeros and are not read.   */
typedef struct curla_conv_seg {
    const void* in;
    const void* wts;
    const float* bias;
    void* out;
    int B;
} curla_conv_seg;
int curla_conv_fwd_multi(const curla_conv_seg* segs, int nseg, long long in_sstride, float scale,
                         long long out_sstride, int pitch, int S, int Hv, int Wv,
                         int first_layer, curla_stream_t stream);
/* conv-2 .. conv-4 (encoder.py:83-87, the three 32 -> 32 channel 3x3 layers + ReLU) in ONE launch: a CTA keeps
 * one sample's activation map in shared memory and runs the three layers over it in place, so the intermediate
 * activations only reach HBM where the caller asks for them (out[0], out[1] non-NULL: passes whose backward reads
 * them).  in = conv-1 output, out[2] = conv-4 output, all [B][4 planes][S][8] bf16 with sample stride `sstride`
 * elements; w96 = the three layers' weights packed by curla_pack_shadows kind 3 ([layer][dy][k chunk][96][8]).
 * Up to three passes (two distinct weight sets) per launch.  curla_conv_stack_fits: 1 when the geometry fits
 * one SM's shared memory (the 76 x 135 crop does, 90 x 160 does not) and CURLA_CONV_FUSED=1 (opt-in experiment). */
typedef struct curla_conv_stack_seg {
    const void* in;
    const void* w96;
    const float* bias[3];
    void* out[3];
    int B;
} curla_conv_stack_seg;
int curla_conv_stack_fits(int pitch, int S, const int* Hv, const int* Wv);
int curla_conv_stack_fwd(const curla_conv_stack_seg* segs, int nseg, long long sstride, int pitch, int S,
                         const int* Hv, const int* Wv, curla_stream_t stream);
/* timing experiments (CURLA_TC_DEBUG=64): per-CTA cycle counters of the conv pipeline roles */
int curla_conv_debug_read(long long* out, int n);
long long curla_conv_wgrad_workspace_floats(int first_layer);
int curla_conv_wgrad(const void* in, long long in_sstride, const void* dy, long long dy_sstride,
                     float* workspace, float* dW, float* db, float scale, int B, int pitch,
                     int S, int Hv, int Wv, int Cin, int first_layer, curla_stream_t stream);

/* two-stage form: per-CTA partials for one layer, then ONE deterministic reduction launch for
 * up to four layers (the engine defers every layer's second stage to the end of a backward) */
int curla_conv_wgrad_partial(const void* in, long long in_sstride, const void* dy,
                             long long dy_sstride, float* workspace, int B, int pitch, int S,
                             int Hv, int Wv, int first_layer, int* nparts_out,
                             curla_stream_t stream);
int curla_conv_wgrad_reduce_multi(int n, float* const* workspace, const int* nparts,
                                  const int* first_layer, const int* Cin, const float* scale,
                                  float* const* dW, float* const* db, curla_stream_t stream);

/* ---- K7/K9: bf16 tensor-core GEMM  (encoder.py:98; curl_sac.py:70-74,129-133) ----------- */
int curla_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* C,
                    long long ldc, int M, int N, int K, int layout, int n_store, int out_bf16,
                    const float* bias, int relu, const void* mask, long long ldmask, int splits,
                    long long split_stride, float alpha, curla_stream_t stream);
/* timing experiments: clock counters of one thread of the tcgen05 GEMM (gemm_tc.cu) */
int curla_gemm_tc_debug_read(long long* out8);
/* timing experiments (CURLA_GEMM_STAMPS=1): eight %globaltimer stamps (ns) per CTA of the last tcgen05 GEMM launch */
int curla_gemm_tc_stamps_read(long long* out, int nctas);
/* same, with the contiguous index of A (seg_mask&1), B (&2) or C+mask columns (&4) split into
 * segments of seg_len elements seg_stride apart (channel-plane activations) */
int curla_gemm_bf16_seg(const void* A, long long lda, const void* B, long long ldb, void* C,
                        long long ldc, int M, int N, int K, int layout, int n_store, int out_bf16,
                        const float* bias, int relu, const void* mask, long long ldmask, int splits,
                        long long split_stride, float alpha, int seg_len, long long seg_stride,
                        int seg_mask, curla_stream_t stream);
int curla_gemm_effective_splits(int K, int splits);
/* `batch` independent GEMMs of one shape in one launch (the critic's Q1 || Q2 trunks,
 * curl_sac.py:158-169); bs* = element strides between problems, 0 shares the operand */
int curla_gemm_bf16_batched(const void* A, long long lda, const void* B, long long ldb, void* C,
                            long long ldc, int M, int N, int K, int layout, int n_store,
                            int out_bf16, const float* bias, int relu, const void* mask,
                            long long ldmask, float alpha, int batch, long long bsA, long long bsB,
                            long long bsC, long long bsBias, long long bsMask,
                            curla_stream_t stream);

/* ---- K8 LayerNorm, K10 policy head, K11 losses, MLP heads  (encoder.py:98-110;
 *      curl_sac.py:20-35,79-110,349-404) ------------------------------------------------ */
int curla_ln_fwd(const float* partial, int nsplit, long long split_stride, const float* bias,
                 const float* gamma, const float* beta, int B, int feat, int apply_tanh,
                 float* x_out, float* z_out, curla_stream_t stream);
/* same + the bf16 input row of the MLP that consumes z: X[b] = [z | act[b] | 0]
 * (torch.cat([obs, action], dim=1), curl_sac.py:137) */
int curla_ln_fwd_x(const float* partial, int nsplit, long long split_stride, const float* bias,
                   const float* gamma, const float* beta, int B, int feat, int apply_tanh,
                   float* x_out, float* z_out, const float* act, int A, void* X_out,
                   curla_stream_t stream);
int curla_ln_bwd(const float* dz_a, const float* dz_b, const float* x_in, const float* gamma,
                 int B, int feat, float* dx_f32, void* dx_bf16, float* scratch, float* dgamma,
                 float* dbeta, float* dbias_fc, curla_stream_t stream);
int curla_pack_x(const float* z, const float* act, int B, int feat, int A, void* X,
                 curla_stream_t stream);
int curla_head_fwd(const void* H, int ldh, const float* W, const float* bias, int B, int hid,
                   int No, float* out, curla_stream_t stream);
int curla_head_bwd(const float* dOut, const float* W, const void* H, int B, int hid, int No,
                   void* dH, curla_stream_t stream);
int curla_head_wgrad(const float* dOut, const void* H, int B, int hid, int No, float* dW,
                     float* db, curla_stream_t stream);
int curla_colsum_bf16(const void* dH, int B, int hid, float* db, curla_stream_t stream);
/* nb heads of one shape per launch (Q1 || Q2); bsP = stride of every parameter(-gradient) pointer */
int curla_head_fwd_batched(const void* H, int ldh, const float* W, const float* bias, int B, int hid,
                           int No, float* out, int nb, long long bsH, long long bsP, long long bsOut,
                           curla_stream_t stream);
int curla_head_bwd_batched(const float* dOut, const float* W, const void* H, int B, int hid, int No,
                           void* dH, int nb, long long bsDOut, long long bsP, long long bsH,
                           long long bsDH, curla_stream_t stream);
int curla_head_wgrad_batched(const float* dOut, const void* H, int B, int hid, int No, float* dW,
                             float* db, int nb, long long bsDOut, long long bsH, long long bsP,
                             curla_stream_t stream);
int curla_colsum_bf16_batched(const void* dH, int B, int hid, float* db, int nb, long long bsDH,
                              long long bsP, curla_stream_t stream);
int curla_policy_fwd(const float* t, const float* noise_in, unsigned long long seed,
                     unsigned long long offset, int B, int A, float ls_min, float ls_max,
                     int compute_pi, int compute_log_pi, float* mu, float* pi, float* log_pi,
                     float* ls, float* noise_out, curla_stream_t stream);
int curla_policy_fwd_rows(const float* t, const float* noise_in, unsigned long long seed,
                          unsigned long long offset, int row0, int B, int A, float ls_min,
                          float ls_max, int compute_pi, int compute_log_pi, float* mu, float* pi,
                          float* log_pi, float* ls, float* noise_out, curla_stream_t stream);
/* Philox offset = offset + *offset_dev (device pointer, may be NULL): the update's captured CUDA graph
 * keeps its per-update counter in device memory */
int curla_policy_fwd_rows_dyn(const float* t, const float* noise_in, unsigned long long seed,
                              unsigned long long offset, const unsigned long long* offset_dev,
                              int row0, int B, int A, float ls_min, float ls_max, int compute_pi,
                              int compute_log_pi, float* mu, float* pi, float* log_pi, float* ls,
                              float* noise_out, curla_stream_t stream);
/* metrics[0..14] and a sequence word (index 15, uint32: seq, or (*seq_dev / 2) + 1 when seq_dev != NULL) written
 * into mapped pinned host memory: the host polls word 15 and reads an update's logged scalars while the update's
 * tail (CURL backward, optimizer steps) is still running */
int curla_publish_metrics(const float* metrics, float* mailbox_host_mapped, unsigned seq,
                          const unsigned long long* seq_dev, curla_stream_t stream);
int curla_policy_bwd(const float* dx1, const float* dx2, int feat, const float* glogpi,
                     const float* t, const float* noise, const float* pi, const float* ls, int B,
                     int A, float ls_min, float ls_max, float* dt, curla_stream_t stream);
int curla_critic_loss(const float* tq1, const float* tq2, const float* logpi_next,
                      const float* reward, const float* not_done, const double* log_alpha,
                      float discount, const float* q1, const float* q2, int B, float grad_scale,
                      float* target_q, float* dq1, float* dq2, float* metrics,
                      curla_stream_t stream);
int curla_actor_loss(const float* log_pi, const float* q1, const float* q2, const float* ls, int B,
                     int A, const double* log_alpha, float target_entropy, float grad_scale,
                     float* dq1, float* dq2, float* glogpi, double* g_log_alpha, float* metrics,
                     curla_stream_t stream);
int curla_add2(const float* a, const float* b, long long n, float* out, curla_stream_t stream);

/* ---- K12: CURL bilinear logits + cross-entropy fwd/bwd  (curl_sac.py:211-222,406-417) -- */
long long curla_curl_workspace_floats(int B, int Bg);
int curla_curl_fwd_bwd(const float* z_a, const float* z_pos, const float* W, int B, int Bg,
                       int feat, int label0, float grad_scale, float* workspace, float* loss_out,
                       float* dz_a, float* dW, float* logits_copy, curla_stream_t stream);

/* ---- K13 Adam, K14 EMA, shadow packing  (curl_sac.py:299-313,368,392,404,419-420;
 *      utils.py:37-41) ------------------------------------------------------------------- */
int curla_adam_f32(float* p, const float* g, float* m, float* v, long long n,
                   long long double_from, double lr, double beta1, double beta2, double eps,
                   int t_host, const int* t_dev, curla_stream_t stream);
int curla_adam_f64_scalar(double* p, const double* g, double* state, double lr, double beta1,
                          double beta2, double eps, int t_host, const int* t_dev,
                          curla_stream_t stream);
/* device-resident per-update scalars of a replayed update graph: state = int[4] Adam step counters
 * {critic, actor, log_alpha, encoder+cpc} followed (byte 16) by the uint64 Philox offset of the policy noise */
int curla_set_dev_state(int* state, int t_critic, int t_actor, int t_alpha, int t_cpc,
                        unsigned long long philox_offset, curla_stream_t stream);
int curla_ema_f32(float* target, const float* p, long long n, long long split, double tau_a,
                  double tau_b, curla_stream_t stream);
int curla_pack_shadows(const float* src_arena, void* dst_arena, const long long* segs, int n,
                       curla_stream_t stream);

/* ---- the agent engine: whole CurlSacAgent.update (curl_sac.py:426-451) ------------------ */
typedef struct curla_agent curla_agent;

typedef struct curla_agent_config {
    int C, H, W;              /* encoder input (post-augmentation) shape, obs_shape      */
    int Hf, Wf;               /* stored frame size in the replay buffer                  */
    int feature_dim, hidden_dim, action_dim, num_filters, num_layers;
    int batch;                /* local batch B                                            */
    int global_batch;         /* B * world                                                */
    int rank, world;
    int detach_encoder, pixel_sac;
    int actor_update_freq, critic_target_update_freq, cpc_update_freq;
    double discount, critic_tau, encoder_tau;
    double actor_lr, actor_beta, critic_lr, critic_beta, alpha_lr, alpha_beta, encoder_lr;
    double log_std_min, log_std_max, target_entropy;
} curla_agent_config;

/* arenas the caller allocates (zero-initialised) and binds */
enum {
    CURLA_ARENA_PARAMS = 0,   /* fp32 master parameters (W | critic | actor-own | target)  */
    CURLA_ARENA_SHADOW = 1,   /* bf16 kernel-layout copies                                 */
    CURLA_ARENA_GRADS = 2,    /* fp32 [g_critic | g_actor | g_cpc]                         */
    CURLA_ARENA_ADAM = 3,     /* fp32 m,v for the three fp32 optimizer states              */
    CURLA_ARENA_WORK = 4,     /* activations / scratch (bytes)                             */
    CURLA_ARENA_COUNT = 5
};

curla_agent* curla_agent_create(const curla_agent_config* cfg);
void curla_agent_destroy(curla_agent* a);
/* bytes needed for arena `which` */
long long curla_agent_arena_bytes(const curla_agent* a, int which);
int curla_agent_bind(curla_agent* a, void* const* arenas);
/* tensor table: i in [0, curla_agent_num_tensors): name, arena, byte offset, dims (<=4) */
int curla_agent_num_tensors(const curla_agent* a);
int curla_agent_tensor_info(const curla_agent* a, int i, char* name, int name_cap, int* arena,
                            long long* byte_offset, int* ndim, long long* dims, int* dtype);
/* rebuild every bf16 shadow from the fp32 masters (after load_state_dict / init) */
int curla_agent_refresh_shadows(curla_agent* a, curla_stream_t stream);

typedef struct curla_update_args {
    /* replay storage (device): utils.py:92-96 */
    const uint8_t* obses; const uint8_t* next_obses;
    const float* actions; const float* rewards; const float* not_dones;
    /* sampled indices of the LOCAL shard (device int64[B]); h1/w1 NULL unless random_crop */
    const int64_t* idxs;
    const int64_t* h1_obs; const int64_t* w1_obs;
    const int64_t* h1_next; const int64_t* w1_next;
    const int64_t* h1_pos; const int64_t* w1_pos;
    /* optional pre-augmented float observations [B][C][H][W] (colour-jiggle / noisy-cover);
       when set they replace the uint8 gather for that stream */
    const float* obs_f32; const float* next_f32; const float* pos_f32;
    /* injected policy noise [B][A] (NULL => Philox with seed/offset) */
    const float* noise_next; const float* noise_cur;
    unsigned long long seed, offset;
    int step;                 /* global env step (frequencies use it: curl_sac.py:439-449) */
    int only_cpc;
    int pos_is_obs;           /* identity branch: pos is a value-equal clone of obs      */
    int phases;               /* 0 = the whole update; else a mask of CURLA_PHASE_* so that a
                                 caller (the parity tests) can run curl_sac.py:429 sample,
                                 :349-371 critic, :373-404 actor+alpha, :442-445 EMA and
                                 :406-423 CPC as separate calls on the same staged minibatch */
} curla_update_args;
enum { CURLA_PHASE_SAMPLE = 1, CURLA_PHASE_CRITIC = 2, CURLA_PHASE_ACTOR = 4, CURLA_PHASE_EMA = 8,
       CURLA_PHASE_CPC = 16, CURLA_PHASE_ALL = 31 };

/* One whole update.  metrics: device float[16]
 * {batch_reward, critic_loss, actor_loss, entropy, alpha_loss, alpha, curl_loss, target_entropy} */
int curla_agent_update(curla_agent* a, const curla_update_args* args, curla_stream_t stream);
/* 1: every encoder pass stores its conv-2 / conv-3 activations (the --log_param_hist_imgs taps read them,
 * encoder.py:118-130); 0 (default): only the passes a backward follows (the fused conv kernel keeps the rest on chip) */
int curla_agent_set_keep_acts(curla_agent* a, int on);
/* mailbox: 64 bytes of mapped pinned host memory (or NULL: off).  When set, every update publishes its logged
 * scalars there right after the last kernel that writes one (curla_publish_metrics; sequence word = the update's
 * args->offset + 1), so a caller that logs every step does not wait for the whole update */
int curla_agent_set_mailbox(curla_agent* a, float* mailbox_host_mapped);
/* number of kernel launches issued by the last curla_agent_update */
int curla_agent_last_launches(const curla_agent* a);
/* per-optimizer Adam step counters {critic, actor, log_alpha, encoder+cpc} (the reference
 * keeps them in torch.optim.Adam.state[...]['step']); settable so a run can be resumed or
 * teacher-forced from another implementation's optimizer state */
int curla_agent_set_opt_steps(curla_agent* a, int t_critic, int t_actor, int t_alpha, int t_cpc);
int curla_agent_get_opt_steps(const curla_agent* a, int* out4);
/* CUDA-event profiler (measurement only).  on = 1: an event after every launch, everything serialised on the engine
 * stream: per-launch device times.  on = 2: side streams stay on, an event on the main stream at each phase boundary of
 * the update (joins included): where the main stream's time goes.  Profiled updates are launched eagerly. */
int curla_profile_enable(int on);
int curla_profile_read(char* buf, int cap);

/* inference entry points (sample_action / select_action / eval / latent extraction:
 * curl_sac.py:330-347, plot_tsne/latent_data.py:63-104).  obs_s2d rows as produced by
 * curla_gather_crop_s2d / curla_f32_to_s2d; net: 0 actor, 1 critic, 2 target.           */
int curla_agent_encode(curla_agent* a, int net, const void* obs_s2d, int B, int apply_tanh,
                       float* z_out /* [B][64] */, curla_stream_t stream);
int curla_agent_actor_head(curla_agent* a, const float* z, int B, const float* noise,
                           unsigned long long seed, unsigned long long offset, int compute_pi,
                           int compute_log_pi, float* mu, float* pi, float* log_pi,
                           float* log_std, curla_stream_t stream);
int curla_agent_q_heads(curla_agent* a, int net, const float* z, const float* action, int B,
                        float* q1, float* q2, curla_stream_t stream);

/* data-parallel plumbing: NCCL communicator shared by all collectives of the update */
int curla_nccl_unique_id(void* out128);
int curla_agent_init_comm(curla_agent* a, const void* id128);
/* hand `from`'s communicator to `to` (an engine re-created with another batch size on the same
 * rank); not a collective.  curla_agent_destroy destroys a communicator the agent still owns. */
int curla_agent_take_comm(curla_agent* to, curla_agent* from);

#ifdef __cplusplus
}
#endif
#endif /* CURLA_B200_H */
