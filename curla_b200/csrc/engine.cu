// The agent engine: memory plan + launch sequence of one whole CurlSacAgent.update
// (curl_sac.py:426-451) on one GPU, data-parallel over NCCL when world > 1.
//
// Host-side C++ only (no kernels here).  The caller (curla_b200/curl_sac.py) owns five
// zero-initialised arenas; this file decides where every tensor lives in them, exposes
// that table (curla_agent_tensor_info) and issues the kernels of gather.cu / conv_tc.cu /
// conv_wgrad_tc.cu / gemm_tc.cu / gemm.cu / small.cu / curl.cu / optim.cu in dependency order:
// the big kernels on the caller's stream, the chains of small kernels that feed nothing downstream
// (pass tails, Q-head / fc weight gradients, the actor's backward + optimizer step) on an
// engine-owned side stream, the collectives and the optimizer steps they feed on a communication
// stream (world > 1) -- event fork / join only, no host synchronisation; the whole sequence is
// captured once per update variant and replayed as a CUDA graph.
//
// Distinct-work schedule (numerically identical to the reference's 7 encoder forwards,
// SURVEY.md 3.4): F1 conv_theta(next), F2 conv_target(next), F3 conv_theta(obs)+bwd,
// F4 conv_theta'(obs) shared by the actor step, the critic-on-pi evaluation and the CURL
// anchor, F7 conv_target'(pos).
#include "common.cuh"
#include "../../include/curla_b200.h"

#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>      // header-only: ranges are no-ops unless a profiler is attached
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <map>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace curla {

// ---------------------------------------------------------------- error plumbing
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static thread_local long long g_launches = 0;
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("CURLA_NO_PDL"); on = (e && e[0] == '1') ? 0 : 1; }
    return on == 1;
}

// Optional CUDA-event profiler: one event after every launch on the engine's stream; the
// interval between consecutive events is that launch's device time (single stream, so
// launches are serialised).  Used by bench.py for the live per-kernel roofline numbers.
struct Profiler {
    bool on = false;
    bool phases = false;       // phase mode: no per-launch marks, side streams stay on, explicit marks at phase boundaries
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
    void mark(const char* what) {
        cudaEvent_t e = get();
        cudaEventRecord(e, stream);
        marks.push_back({what, e});
    }
};
// per host thread: an agent is driven by one thread (include/curla_b200.h: thread-compatible), so two
// engines driven by two threads profile independently
static thread_local Profiler g_prof;
void profile_mark(const char* what) { if (g_prof.on && !g_prof.phases) g_prof.mark(what); }
// phase mode (curla_profile_enable(2)): where the MAIN stream's time goes with every side stream running -- an event
// on the main stream at each phase boundary, waits for joins included
static void phase_mark(const char* what, cudaStream_t st) { if (g_prof.on && g_prof.phases && st == g_prof.stream) g_prof.mark(what); }

static thread_local const char* g_tag = nullptr;
void set_launch_tag(const char* tag) { g_tag = tag; }

int check_launch(const char* what) {
    ++g_launches;
    if (g_tag) what = g_tag;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("%s: %s", what, cudaGetErrorString(e));
        return -1;
    }
    if (g_prof.on && !g_prof.phases) g_prof.mark(what);
    return 0;
}
int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// The conv kernels are persistent with one CTA per SM and nearly all of its shared memory and registers: an NCCL
// kernel launched beside them finds no SM to run on until a conv kernel ends.  With world > 1 the engine therefore
// keeps a few SMs out of the conv grids (the all-reduce of a gradient slice then really runs during the conv
// backward instead of between its launches).
static int g_reserved_sms = -1;
int conv_grid_cap() {
    if (g_reserved_sms < 0) { const char* e = getenv("CURLA_RESERVE_SMS"); g_reserved_sms = e ? atoi(e) : 0; }
    int r = g_reserved_sms;
    if (r < 0) r = 0;
    if (r > sm_count() / 2) r = sm_count() / 2;
    return sm_count() - r;
}

enum { DT_F32 = 0, DT_BF16 = 1, DT_F64 = 2, DT_I32 = 3, DT_U8 = 4 };
static const int kDtSize[] = {4, 2, 8, 4, 1};

struct TensorInfo {
    std::string name;
    int arena;
    long long byte_off;
    int ndim;
    long long dims[4];
    int dtype;
};

struct EncP { long long conv_w[4], conv_b[4], fc_w, fc_b, ln_w, ln_b; };      // float offsets
struct MlpP { long long w0, b0, w1, b1, w2, b2; int in_real, out; };
struct EncS { long long conv[4], fc, conv96; };                                // bf16 offsets (conv96: conv-2..4 packed for the fused kernel)
struct MlpS { long long w0, w1; };
struct MlpBuf { bf16 *X, *H1, *H2; float* out; };
struct TailBuf { float *fc_out, *z; };

}  // namespace curla

using namespace curla;

struct curla_agent {
    curla_agent_config cfg;
    // geometry
    int Hs, pitch, S, Ho[4], Wo[4], Kfc, CP1, PAD, fc_splits;
    // parameter layout (float offsets into ARENA_PARAMS)
    long long off_W, n_W, off_critic, n_critic, n_enc, off_actor, n_actor, off_target, n_params;
    EncP enc_critic, enc_target, enc_actor;   // enc_actor.conv_* alias the critic's (tied)
    MlpP q_critic[2], q_target[2], trunk_actor;
    // shadows
    EncS s_critic, s_target; long long s_actor_fc;
    MlpS sq_critic[2], sq_target[2], s_trunk;
    long long n_shadow;
    std::vector<long long> pack_critic, pack_critic_enc, pack_actor, pack_target;   // 7 per seg
    // grads / adam
    long long g_critic, g_actor, g_cpc, n_cpc, n_grads;
    long long a_m1, a_v1, a_m2, a_v2, a_m3, a_v3, n_adam;
    std::vector<TensorInfo> tensors;
    long long arena_bytes[CURLA_ARENA_COUNT];
    // bound arenas
    float* P; bf16* Sh; float* G; float* Ad; uint8_t* Wk;
    bool bound;
    // workspace pointers
    bf16 *s2d_obs, *s2d_next, *s2d_pos, *actA[4], *actB[4], *actC[4], *dact[4];
    long long act_sstride, s2d_sstride;
    float *fc_partial, *fc_partial2, *wgrad_ws, *curl_ws, *ln_scratch, *ln_scratch_b;
    // side stream: the latency-bound tails (fc + LayerNorm + MLP heads) of one encoder pass run
    // there while the main stream already runs the next pass's conv stack
    cudaStream_t side; cudaEvent_t ev[11]; int side_state;   // 0 = not created, 1 = ready, -1 = disabled
    // communication stream (world > 1): gradient all-reduce + Adam of a bucket slice run there while the
    // main stream is still in the conv backward (critic, CURL) or already in the next phase (actor)
    cudaStream_t comm_st; cudaEvent_t cev[7]; int comm_state;
    long long wgrad_ws_stride;
    TailBuf t_p1, t_p2, t_p3, t_p4, t_p5, t_p7;
    MlpBuf m_p1, m_p2q[2], m_p3q[2], m_p4, m_p5q[2];
    float *t_out1, *t_out4, *a_next, *logpi_next, *mu_scratch, *pi4, *logpi4, *ls4, *noise4, *ls1;
    float *tq[2], *q3[2], *q5[2], *target_q, *dq[2], *dt4;
    bf16 *dH2, *dH1;
    float *dX[2], *dXa, *dz_curl, *dfc_f32, *dfc_f32_b; bf16 *dfc_bf16, *dfc_bf16_b;   // _b: the actor's backward beside the CURL phase
    float *z_pos_all, *act_b, *rew_b, *nd_b, *metrics, *metrics_avg, *glogpi;
    double *log_alpha, *g_log_alpha, *alpha_state;
    // optimizer step counters (host)
    int t_critic, t_actor, t_alpha, t_cpc;
    long long last_launches;
    // whole-update CUDA graphs: one per update variant (step parity, only_cpc, argument pointers); the per-update
    // scalars a replay needs (Adam step counters, Philox offset) live in dev_state (see curla_set_dev_state)
    int keep_acts;                           // 1: every encoder pass stores conv-2 / conv-3 activations (logging taps read them)
    float* mailbox;                          // mapped pinned host memory for the logged scalars (curla_agent_set_mailbox) or NULL
    int* dev_state;
    cudaStream_t cap_st;                     // private capture stream (the caller's may be the legacy default stream,
                                             // which cannot be captured); replays are launched into the caller's stream
    struct GraphEntry { cudaGraphExec_t exec; long long launches; int seen; };
    std::map<std::string, GraphEntry> graphs;
    int graph_mode;                          // -1 = not decided yet, 0 = off (CURLA_GRAPH=0), 1 = on
    // NCCL (dlopen'ed)
    void* nccl_lib; void* comm;
};

namespace {

// NVTX range per phase of the update (nsys / ncu --nvtx show "curla/critic" ... around its launches)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ---------------------------------------------------------------- layout helpers
struct Builder {
    curla_agent* a;
    long long cur[CURLA_ARENA_COUNT] = {0, 0, 0, 0, 0};
    long long add(int arena, const char* name, int dtype, int ndim, const long long* dims,
                  long long align_elems, long long front_pad_elems = 0, long long back_pad_elems = 0) {
        long long n = 1;
        for (int i = 0; i < ndim; ++i) n *= dims[i];
        const int es = kDtSize[dtype];
        long long off_bytes = cur[arena];
        const long long al = align_elems * es;
        off_bytes = (off_bytes + al - 1) / al * al;
        off_bytes += front_pad_elems * es;
        TensorInfo t;
        t.name = name; t.arena = arena; t.byte_off = off_bytes; t.ndim = ndim; t.dtype = dtype;
        for (int i = 0; i < 4; ++i) t.dims[i] = i < ndim ? dims[i] : 1;
        a->tensors.push_back(t);
        cur[arena] = off_bytes + (n + back_pad_elems) * es;
        return off_bytes / es;
    }
    long long p(const std::string& name, std::vector<long long> dims) {   // fp32 parameter
        return add(CURLA_ARENA_PARAMS, name.c_str(), DT_F32, (int)dims.size(), dims.data(), 4);
    }
    long long s(const std::string& name, std::vector<long long> dims) {   // bf16 shadow
        return add(CURLA_ARENA_SHADOW, ("shadow." + name).c_str(), DT_BF16, (int)dims.size(), dims.data(), 64);
    }
    template <typename T>
    T* w(const std::string& name, int dtype, std::vector<long long> dims, long long fp = 0, long long bp = 0) {
        long long off = add(CURLA_ARENA_WORK, name.c_str(), dtype, (int)dims.size(), dims.data(),
                            256 / kDtSize[dtype], fp, bp);
        return reinterpret_cast<T*>(off * kDtSize[dtype]);   // offset now, rebased at bind
    }
};

void layout_encoder(Builder& b, const std::string& pre, EncP& e, const curla_agent* a, const EncP* tie) {
    const auto& c = a->cfg;
    for (int i = 0; i < 4; ++i) {
        if (tie) { e.conv_w[i] = tie->conv_w[i]; e.conv_b[i] = tie->conv_b[i]; continue; }
        const long long cin = i == 0 ? c.C : c.num_filters;
        e.conv_w[i] = b.p(pre + "convs." + std::to_string(i) + ".weight", {c.num_filters, cin, 3, 3});
        e.conv_b[i] = b.p(pre + "convs." + std::to_string(i) + ".bias", {c.num_filters});
    }
    // canonical fc layout is the kernel layout: [feat][F/8 planes][Ho4*pitch][8] (zero in pad columns)
    e.fc_w = b.p(pre + "fc.weight_canon", {c.feature_dim, c.num_filters / 8, a->Ho[3] * a->pitch, 8});
    e.fc_b = b.p(pre + "fc.bias", {c.feature_dim});
    e.ln_w = b.p(pre + "ln.weight", {c.feature_dim});
    e.ln_b = b.p(pre + "ln.bias", {c.feature_dim});
}
void layout_mlp(Builder& b, const std::string& pre, MlpP& m, int in_real, int hid, int out) {
    m.in_real = in_real; m.out = out;
    m.w0 = b.p(pre + "0.weight", {hid, in_real}); m.b0 = b.p(pre + "0.bias", {hid});
    m.w1 = b.p(pre + "2.weight", {hid, hid});     m.b1 = b.p(pre + "2.bias", {hid});
    m.w2 = b.p(pre + "4.weight", {out, hid});     m.b2 = b.p(pre + "4.bias", {out});
}
void seg(std::vector<long long>& v, long long src, long long dst, int kind, int rows, int cols,
         int rows_pad, int cols_pad) {
    const long long r[7] = {src, dst, kind, rows, cols, rows_pad, cols_pad};
    v.insert(v.end(), r, r + 7);
}
void shadow_encoder(Builder& b, const std::string& pre, const EncP& e, EncS& s, curla_agent* a,
                    std::vector<long long>* packs[], int npacks, bool convs) {
    const auto& c = a->cfg;
    if (convs) {
        s.conv[0] = b.s(pre + "convs.0", {4, c.num_filters, a->CP1});
        for (int k = 0; k < npacks; ++k) seg(*packs[k], e.conv_w[0], s.conv[0], 2, c.num_filters, c.C, c.num_filters, a->CP1);
        for (int i = 1; i < 4; ++i) {
            s.conv[i] = b.s(pre + "convs." + std::to_string(i), {9, c.num_filters, c.num_filters});
            for (int k = 0; k < npacks; ++k) seg(*packs[k], e.conv_w[i], s.conv[i], 1, c.num_filters, c.num_filters, c.num_filters, c.num_filters);
        }
        // conv-2..4 once more in the N = 96 operand image of the fused kernel (curla_conv_stack_fwd): one 55 KB bulk copy
        s.conv96 = b.s(pre + "convs96", {3, 12, 96, 8});
        for (int i = 1; i < 4; ++i)
            for (int k = 0; k < npacks; ++k) seg(*packs[k], e.conv_w[i], s.conv96 + (long long)(i - 1) * 9216, 3, 32, 32, 32, 32);
    }
    s.fc = b.s(pre + "fc", {64, a->Kfc});
    for (int k = 0; k < npacks; ++k) seg(*packs[k], e.fc_w, s.fc, 0, c.feature_dim, a->Kfc, 64, a->Kfc);
}
void shadow_mlp(Builder& b, const std::string& pre, const MlpP& m, MlpS& s, int hid,
                std::vector<long long>& pack) {
    s.w0 = b.s(pre + "0", {hid, 64});
    seg(pack, m.w0, s.w0, 0, hid, m.in_real, hid, 64);
    s.w1 = b.s(pre + "2", {hid, hid});
    seg(pack, m.w1, s.w1, 0, hid, hid, hid, hid);
}

}  // namespace

extern "C" const char* curla_last_error(void) { return g_err; }
extern "C" int curla_version(void) { return 100; }

extern "C" curla_agent* curla_agent_create(const curla_agent_config* cfg) {
    const auto& c = *cfg;
    if (c.num_filters != 32 || c.num_layers != 4) { set_last_error("agent: only num_filters=32, num_layers=4 are built"); return nullptr; }
    if (c.feature_dim > 64 || c.feature_dim + c.action_dim > 64 || c.action_dim > 4 || (c.feature_dim & 1) || ((c.feature_dim + c.action_dim) & 1)) {
        set_last_error("agent: need even feature_dim, feature_dim+action_dim <= 64, action_dim <= 4"); return nullptr; }
    if (c.hidden_dim % 8 || c.hidden_dim < 8) { set_last_error("agent: hidden_dim must be a multiple of 8"); return nullptr; }
    if (4 * c.C > 48) { set_last_error("agent: at most 12 input channels"); return nullptr; }
    if (c.H < 11 || c.W < 11 || c.batch < 1 || c.global_batch != c.batch * c.world) { set_last_error("agent: bad shape/batch"); return nullptr; }
    curla_agent* a = new curla_agent();
    a->cfg = c;
    a->Hs = (c.H + 1) / 2; a->pitch = (c.W + 1) / 2; a->S = a->Hs * a->pitch;
    a->Ho[0] = (c.H - 3) / 2 + 1; a->Wo[0] = (c.W - 3) / 2 + 1;
    for (int i = 1; i < 4; ++i) { a->Ho[i] = a->Ho[i - 1] - 2; a->Wo[i] = a->Wo[i - 1] - 2; }
    a->Kfc = a->Ho[3] * a->pitch * 32;
    a->CP1 = 48;
    a->PAD = curla_conv_pad_rows(a->pitch);
    const int B = c.batch, Bg = c.global_batch, hid = c.hidden_dim, feat = c.feature_dim, A = c.action_dim;
    Builder b; b.a = a;

    // ---- parameters: [W][critic: enc Q1 Q2][actor-own: fc ln trunk][target: enc Q1 Q2]
    a->off_W = b.p("CURL.W", {feat, feat});
    a->off_critic = (b.cur[0] + 15) / 16 * 4;
    a->n_W = a->off_critic - a->off_W;            // incl. alignment tail (zero grads there)
    layout_encoder(b, "critic.encoder.", a->enc_critic, a, nullptr);
    a->n_enc = (b.cur[0] + 15) / 16 * 4 - a->off_critic;
    layout_mlp(b, "critic.Q1.trunk.", a->q_critic[0], feat + A, hid, 1);
    layout_mlp(b, "critic.Q2.trunk.", a->q_critic[1], feat + A, hid, 1);
    a->off_actor = (b.cur[0] + 15) / 16 * 4;
    a->n_critic = a->off_actor - a->off_critic;
    {   // actor-own: fc, ln (convs tied to the critic's: curl_sac.py:290)
        EncP& e = a->enc_actor;
        for (int i = 0; i < 4; ++i) { e.conv_w[i] = a->enc_critic.conv_w[i]; e.conv_b[i] = a->enc_critic.conv_b[i]; }
        e.fc_w = b.p("actor.encoder.fc.weight_canon", {feat, 4, a->Ho[3] * a->pitch, 8});
        e.fc_b = b.p("actor.encoder.fc.bias", {feat});
        e.ln_w = b.p("actor.encoder.ln.weight", {feat});
        e.ln_b = b.p("actor.encoder.ln.bias", {feat});
    }
    layout_mlp(b, "actor.trunk.", a->trunk_actor, feat, hid, 2 * A);
    a->off_target = (b.cur[0] + 15) / 16 * 4;
    a->n_actor = a->off_target - a->off_actor;
    layout_encoder(b, "target.encoder.", a->enc_target, a, nullptr);
    layout_mlp(b, "target.Q1.trunk.", a->q_target[0], feat + A, hid, 1);
    layout_mlp(b, "target.Q2.trunk.", a->q_target[1], feat + A, hid, 1);
    a->n_params = (b.cur[0] + 15) / 16 * 4;
    if (a->n_params - a->off_target != a->n_critic) { set_last_error("agent: internal layout mismatch"); delete a; return nullptr; }
    a->arena_bytes[CURLA_ARENA_PARAMS] = a->n_params * 4;

    // ---- shadows + pack tables
    {
        std::vector<long long>* pc[] = {&a->pack_critic, &a->pack_critic_enc};
        shadow_encoder(b, "critic.encoder.", a->enc_critic, a->s_critic, a, pc, 2, true);
        shadow_mlp(b, "critic.Q1.", a->q_critic[0], a->sq_critic[0], hid, a->pack_critic);
        shadow_mlp(b, "critic.Q2.", a->q_critic[1], a->sq_critic[1], hid, a->pack_critic);
        std::vector<long long>* pa[] = {&a->pack_actor};
        EncS tmp;
        shadow_encoder(b, "actor.encoder.", a->enc_actor, tmp, a, pa, 1, false);
        a->s_actor_fc = tmp.fc;
        shadow_mlp(b, "actor.trunk.", a->trunk_actor, a->s_trunk, hid, a->pack_actor);
        std::vector<long long>* pt[] = {&a->pack_target};
        shadow_encoder(b, "target.encoder.", a->enc_target, a->s_target, a, pt, 1, true);
        shadow_mlp(b, "target.Q1.", a->q_target[0], a->sq_target[0], hid, a->pack_target);
        shadow_mlp(b, "target.Q2.", a->q_target[1], a->sq_target[1], hid, a->pack_target);
        a->n_shadow = b.cur[1] / 2;
        a->arena_bytes[CURLA_ARENA_SHADOW] = (b.cur[1] + 255) / 256 * 256;
    }
    // ---- grads: g_critic mirrors the critic segment, g_actor the actor-own segment,
    //      g_cpc mirrors [W | critic.encoder]
    a->g_critic = 0; a->g_actor = a->n_critic; a->g_cpc = a->n_critic + a->n_actor;
    a->n_cpc = a->n_W + a->n_enc;
    a->n_grads = a->g_cpc + a->n_cpc;
    a->arena_bytes[CURLA_ARENA_GRADS] = a->n_grads * 4;
    a->a_m1 = 0; a->a_v1 = a->n_critic;
    a->a_m2 = 2 * a->n_critic; a->a_v2 = a->a_m2 + a->n_actor;
    a->a_m3 = a->a_v2 + a->n_actor; a->a_v3 = a->a_m3 + a->n_cpc;
    a->n_adam = a->a_v3 + a->n_cpc;
    a->arena_bytes[CURLA_ARENA_ADAM] = a->n_adam * 4;
    {   // expose the grad / adam sub-arenas as named tensors
        const long long d1[1] = {a->n_critic}, d2[1] = {a->n_actor}, d3[1] = {a->n_cpc};
        auto named = [&](int arena, const char* nm, long long off, const long long* d) {
            TensorInfo t; t.name = nm; t.arena = arena; t.byte_off = off * 4; t.ndim = 1; t.dtype = DT_F32;
            t.dims[0] = d[0]; t.dims[1] = t.dims[2] = t.dims[3] = 1; a->tensors.push_back(t);
        };
        named(CURLA_ARENA_GRADS, "grad.critic", a->g_critic, d1);
        named(CURLA_ARENA_GRADS, "grad.actor", a->g_actor, d2);
        named(CURLA_ARENA_GRADS, "grad.cpc", a->g_cpc, d3);
        named(CURLA_ARENA_ADAM, "adam.critic.m", a->a_m1, d1); named(CURLA_ARENA_ADAM, "adam.critic.v", a->a_v1, d1);
        named(CURLA_ARENA_ADAM, "adam.actor.m", a->a_m2, d2);  named(CURLA_ARENA_ADAM, "adam.actor.v", a->a_v2, d2);
        named(CURLA_ARENA_ADAM, "adam.cpc.m", a->a_m3, d3);    named(CURLA_ARENA_ADAM, "adam.cpc.v", a->a_v3, d3);
    }

    // ---- workspace
    const long long R = (long long)B * a->S, PADR = a->PAD;
    a->act_sstride = (long long)a->S * 32; a->s2d_sstride = (long long)a->S * a->CP1;
    a->s2d_obs = b.w<bf16>("s2d.obs", DT_BF16, {R, a->CP1}, PADR * a->CP1, PADR * a->CP1);
    a->s2d_next = b.w<bf16>("s2d.next", DT_BF16, {R, a->CP1}, PADR * a->CP1, PADR * a->CP1);
    a->s2d_pos = b.w<bf16>("s2d.pos", DT_BF16, {R, a->CP1}, PADR * a->CP1, PADR * a->CP1);
    for (int i = 0; i < 4; ++i) {
        a->actA[i] = b.w<bf16>("actA." + std::to_string(i), DT_BF16, {R, 32}, PADR * 32, PADR * 32);
        a->actB[i] = b.w<bf16>("actB." + std::to_string(i), DT_BF16, {R, 32}, PADR * 32, PADR * 32);
        a->actC[i] = b.w<bf16>("actC." + std::to_string(i), DT_BF16, {R, 32}, PADR * 32, PADR * 32);
        a->dact[i] = b.w<bf16>("dact." + std::to_string(i), DT_BF16, {R, 32}, PADR * 32, PADR * 32);
    }
    {   // split-K for the fc forward: ~4 CTAs of 64x64 tiles per SM (bytes in flight, not FLOPs, bound it)
        const int mt = cdiv(B, 64);
        int sp = cdiv(4 * sm_count(), mt);
        const int ktiles = cdiv(a->Kfc, 32);
        if (sp > ktiles / 4) sp = ktiles / 4 > 0 ? ktiles / 4 : 1;
        { const char* e = getenv("CURLA_FC_SPLITS"); if (e && atoi(e) >= 1 && atoi(e) <= sp) sp = atoi(e); }     // timing experiments
        a->fc_splits = curla_gemm_effective_splits(a->Kfc, sp);
    }
    a->fc_partial = b.w<float>("fc_partial", DT_F32, {a->fc_splits, B, 64});
    a->fc_partial2 = b.w<float>("fc_partial2", DT_F32, {a->fc_splits, B, 64});
    {
        long long w1 = curla_conv_wgrad_workspace_floats(1), w2 = curla_conv_wgrad_workspace_floats(0);
        a->wgrad_ws = b.w<float>("wgrad_ws", DT_F32, {4, w1 > w2 ? w1 : w2});   // one slice per conv layer
        a->wgrad_ws_stride = w1 > w2 ? w1 : w2;
    }
    a->curl_ws = b.w<float>("curl_ws", DT_F32, {curla_curl_workspace_floats(B, Bg)});
    a->ln_scratch = b.w<float>("ln_scratch", DT_F32, {2, B, 64});
    a->ln_scratch_b = b.w<float>("ln_scratch_b", DT_F32, {2, B, 64});
    auto tail = [&](const char* nm, TailBuf& t) {
        t.fc_out = b.w<float>(std::string(nm) + ".fc_out", DT_F32, {B, 64});
        t.z = b.w<float>(std::string(nm) + ".z", DT_F32, {B, 64});
    };
    auto mlp = [&](const char* nm, MlpBuf& m, int out) {
        m.X = b.w<bf16>(std::string(nm) + ".X", DT_BF16, {B, 64});
        m.H1 = b.w<bf16>(std::string(nm) + ".H1", DT_BF16, {B, hid});
        m.H2 = b.w<bf16>(std::string(nm) + ".H2", DT_BF16, {B, hid});
        m.out = b.w<float>(std::string(nm) + ".out", DT_F32, {B, out});
    };
    tail("p1", a->t_p1); tail("p2", a->t_p2); tail("p3", a->t_p3); tail("p4", a->t_p4);
    tail("p5", a->t_p5); tail("p7", a->t_p7);
    mlp("p1.trunk", a->m_p1, 2 * A);
    mlp("p2.q1", a->m_p2q[0], 1); mlp("p2.q2", a->m_p2q[1], 1);
    mlp("p3.q1", a->m_p3q[0], 1); mlp("p3.q2", a->m_p3q[1], 1);
    mlp("p4.trunk", a->m_p4, 2 * A);
    mlp("p5.q1", a->m_p5q[0], 1); mlp("p5.q2", a->m_p5q[1], 1);
    a->t_out1 = a->m_p1.out; a->t_out4 = a->m_p4.out;
    a->tq[0] = a->m_p2q[0].out; a->tq[1] = a->m_p2q[1].out;
    a->q3[0] = a->m_p3q[0].out; a->q3[1] = a->m_p3q[1].out;
    a->q5[0] = a->m_p5q[0].out; a->q5[1] = a->m_p5q[1].out;
    a->a_next = b.w<float>("next_action", DT_F32, {B, A});
    a->logpi_next = b.w<float>("next_log_pi", DT_F32, {B});
    a->ls1 = b.w<float>("next_log_std", DT_F32, {B, A});
    a->mu_scratch = b.w<float>("mu", DT_F32, {B, A});
    a->pi4 = b.w<float>("pi", DT_F32, {B, A});
    a->logpi4 = b.w<float>("log_pi", DT_F32, {B});
    a->ls4 = b.w<float>("log_std", DT_F32, {B, A});
    a->noise4 = b.w<float>("noise_used", DT_F32, {B, A});
    a->target_q = b.w<float>("target_q", DT_F32, {B});
    a->dq[0] = b.w<float>("dq1", DT_F32, {B}); a->dq[1] = b.w<float>("dq2", DT_F32, {B});
    a->dt4 = b.w<float>("d_trunk_out", DT_F32, {B, 2 * A});
    a->dH2 = b.w<bf16>("dH2", DT_BF16, {2, B, hid}); a->dH1 = b.w<bf16>("dH1", DT_BF16, {2, B, hid});   // Q1 || Q2
    a->dX[0] = b.w<float>("dX1", DT_F32, {B, 64}); a->dX[1] = b.w<float>("dX2", DT_F32, {B, 64});
    a->dXa = b.w<float>("dXa", DT_F32, {B, 64});
    a->dz_curl = b.w<float>("dz_curl", DT_F32, {B, 64});
    a->dfc_f32 = b.w<float>("dfc_f32", DT_F32, {B, 64});
    a->dfc_bf16 = b.w<bf16>("dfc_bf16", DT_BF16, {B, 64});
    a->dfc_f32_b = b.w<float>("dfc_f32_b", DT_F32, {B, 64});
    a->dfc_bf16_b = b.w<bf16>("dfc_bf16_b", DT_BF16, {B, 64});
    a->z_pos_all = b.w<float>("z_pos_all", DT_F32, {Bg, 64});
    a->act_b = b.w<float>("batch.action", DT_F32, {B, A});
    a->rew_b = b.w<float>("batch.reward", DT_F32, {B});
    a->nd_b = b.w<float>("batch.not_done", DT_F32, {B});
    a->metrics = b.w<float>("metrics", DT_F32, {16});
    a->metrics_avg = b.w<float>("metrics_avg", DT_F32, {16});
    a->glogpi = b.w<float>("glogpi", DT_F32, {4});
    a->log_alpha = b.w<double>("log_alpha", DT_F64, {1});
    a->g_log_alpha = b.w<double>("grad.log_alpha", DT_F64, {1});
    a->alpha_state = b.w<double>("adam.log_alpha", DT_F64, {2});
    a->dev_state = b.w<int>("dev_state", DT_I32, {8});
    a->arena_bytes[CURLA_ARENA_WORK] = (b.cur[4] + 255) / 256 * 256;
    {   // the batched Q1 || Q2 launches rely on one constant stride per arena between the two heads
        bool okp = true;
        for (const MlpP* q : {a->q_critic, a->q_target}) {
            const long long d = q[1].w0 - q[0].w0;
            okp = okp && q[1].b0 - q[0].b0 == d && q[1].w1 - q[0].w1 == d && q[1].b1 - q[0].b1 == d &&
                  q[1].w2 - q[0].w2 == d && q[1].b2 - q[0].b2 == d && d % 4 == 0;
        }
        for (const MlpS* q : {a->sq_critic, a->sq_target}) okp = okp && q[1].w1 - q[0].w1 == q[1].w0 - q[0].w0;
        for (const MlpBuf* m : {a->m_p2q, a->m_p3q, a->m_p5q}) okp = okp && m[1].H2 - m[0].H2 == m[1].H1 - m[0].H1;
        if (!okp) { set_last_error("agent: internal layout mismatch (Q1/Q2 strides)"); delete a; return nullptr; }
    }
    a->bound = false;
    a->t_critic = a->t_actor = a->t_alpha = a->t_cpc = 0;
    a->last_launches = 0;
    a->graph_mode = -1;
    a->keep_acts = 0;
    a->mailbox = nullptr;
    a->cap_st = nullptr;
    a->nccl_lib = nullptr; a->comm = nullptr;
    a->side = nullptr; a->side_state = 0;
    a->comm_st = nullptr; a->comm_state = 0;
    return a;
}

static void destroy_comm(curla_agent* a);
extern "C" void curla_agent_destroy(curla_agent* a) {
    if (!a) return;
    for (auto& kv : a->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (a->cap_st) cudaStreamDestroy(a->cap_st);
    if (a->side_state == 1) {
        for (auto& e : a->ev) cudaEventDestroy(e);
        cudaStreamDestroy(a->side);
    }
    if (a->comm_state == 1) {
        for (auto& e : a->cev) cudaEventDestroy(e);
        cudaStreamDestroy(a->comm_st);
    }
    destroy_comm(a);
    delete a;
}
extern "C" long long curla_agent_arena_bytes(const curla_agent* a, int which) {
    return (which >= 0 && which < CURLA_ARENA_COUNT) ? a->arena_bytes[which] : -1;
}
extern "C" int curla_agent_num_tensors(const curla_agent* a) { return (int)a->tensors.size(); }
extern "C" int curla_agent_tensor_info(const curla_agent* a, int i, char* name, int name_cap,
                                       int* arena, long long* byte_offset, int* ndim,
                                       long long* dims, int* dtype) {
    if (i < 0 || i >= (int)a->tensors.size()) { set_last_error("tensor_info: index"); return -1; }
    const TensorInfo& t = a->tensors[i];
    snprintf(name, name_cap, "%s", t.name.c_str());
    *arena = t.arena; *byte_offset = t.byte_off; *ndim = t.ndim; *dtype = t.dtype;
    for (int k = 0; k < 4; ++k) dims[k] = t.dims[k];
    return 0;
}

extern "C" int curla_agent_bind(curla_agent* a, void* const* arenas) {
    for (int i = 0; i < CURLA_ARENA_COUNT; ++i)
        CURLA_CHECK(arenas[i] != nullptr && ((uintptr_t)arenas[i] % 256) == 0, "bind: arena %d null or not 256-byte aligned", i);
    CURLA_CHECK(!a->bound, "bind: already bound");
    a->P = (float*)arenas[0]; a->Sh = (bf16*)arenas[1]; a->G = (float*)arenas[2];
    a->Ad = (float*)arenas[3]; a->Wk = (uint8_t*)arenas[4];
    const uintptr_t base = (uintptr_t)a->Wk;
    auto rb = [&](auto*& p) { p = reinterpret_cast<std::remove_reference_t<decltype(p)>>(base + (uintptr_t)p); };
    rb(a->s2d_obs); rb(a->s2d_next); rb(a->s2d_pos);
    for (int i = 0; i < 4; ++i) { rb(a->actA[i]); rb(a->actB[i]); rb(a->actC[i]); rb(a->dact[i]); }
    rb(a->fc_partial); rb(a->fc_partial2); rb(a->wgrad_ws); rb(a->curl_ws); rb(a->ln_scratch); rb(a->ln_scratch_b);
    TailBuf* tb[] = {&a->t_p1, &a->t_p2, &a->t_p3, &a->t_p4, &a->t_p5, &a->t_p7};
    for (auto t : tb) { rb(t->fc_out); rb(t->z); }
    MlpBuf* mb[] = {&a->m_p1, &a->m_p2q[0], &a->m_p2q[1], &a->m_p3q[0], &a->m_p3q[1], &a->m_p4, &a->m_p5q[0], &a->m_p5q[1]};
    for (auto m : mb) { rb(m->X); rb(m->H1); rb(m->H2); rb(m->out); }
    rb(a->t_out1); rb(a->t_out4); rb(a->tq[0]); rb(a->tq[1]); rb(a->q3[0]); rb(a->q3[1]); rb(a->q5[0]); rb(a->q5[1]);
    rb(a->a_next); rb(a->logpi_next); rb(a->ls1); rb(a->mu_scratch); rb(a->pi4); rb(a->logpi4); rb(a->ls4); rb(a->noise4);
    rb(a->target_q); rb(a->dq[0]); rb(a->dq[1]); rb(a->dt4); rb(a->dH2); rb(a->dH1);
    rb(a->dX[0]); rb(a->dX[1]); rb(a->dXa); rb(a->dz_curl); rb(a->dfc_f32); rb(a->dfc_bf16); rb(a->dfc_f32_b); rb(a->dfc_bf16_b);
    rb(a->z_pos_all); rb(a->act_b); rb(a->rew_b); rb(a->nd_b); rb(a->metrics); rb(a->metrics_avg); rb(a->glogpi);
    rb(a->log_alpha); rb(a->g_log_alpha); rb(a->alpha_state); rb(a->dev_state);
    a->bound = true;
    return 0;
}

// ================================================================== launch helpers
namespace {

// CURLA_MERGE: how independent encoder passes share conv launches (results are bit-identical
// in every mode).  0 = one launch per pass and layer; 1 = F1+F2 (both read next_obs) share
// launches, F3 runs beside their tails, CURL anchor+key share launches; 2 (default) = F1+F2+F3
// share launches as well.
int merge_mode() {
    static int m = -1;
    if (m < 0) { const char* e = getenv("CURLA_MERGE"); m = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2; }
    return m;
}

struct Run {
    curla_agent* a;
    cudaStream_t st;
    int rc = 0;
    bool ok() const { return rc == 0; }
    void chk(int r) { if (r && !rc) rc = r; }
    float* P(long long off) const { return a->P + off; }
    bf16* Sh(long long off) const { return a->Sh + off; }

    void pack(const std::vector<long long>& segs) {
        if (!ok()) return;
        chk(curla_pack_shadows(a->P, a->Sh, segs.data(), (int)(segs.size() / 7), st));
    }
    // conv stack forward: s2d input -> acts[0..3].  Up to three independent passes (own input,
    // weights and output buffers) run as ONE launch per layer (curla_conv_fwd_multi): 4 launches
    // instead of 4 per pass, and no pipeline fill/drain between passes.  CURLA_MERGE=0 launches
    // each pass separately (same results bit for bit).
    // keep: the pass's backward (or a logging tap) reads its conv-2 / conv-3 activations; the others only need conv-4's
    struct Pass { const bf16* s2d; const EncP* e; const EncS* s; bf16* const* acts; bool keep; };
    void conv_stack_multi(const Pass* ps, int np, int B = 0) {
        if (!ok() || np <= 0) return;
        if (!merge_mode() && np > 1) {
            for (int k = 0; k < np; ++k) conv_stack_multi(ps + k, 1, B);
            return;
        }
        if (B <= 0) B = a->cfg.batch;
        // conv-2..4 as ONE launch that keeps each sample in shared memory (the 76 x 135 crop fits, 90 x 160 does not)
        const int Hv3[3] = {a->Ho[1], a->Ho[2], a->Ho[3]}, Wv3[3] = {a->Wo[1], a->Wo[2], a->Wo[3]};
        const bool fused = curla_conv_stack_fits(a->pitch, a->S, Hv3, Wv3) != 0;
        for (int i = 0; i < (fused ? 1 : 4) && ok(); ++i) {
            curla_conv_seg sg[3];
            for (int k = 0; k < np; ++k) {
                sg[k].in = i == 0 ? ps[k].s2d : ps[k].acts[i - 1];
                sg[k].wts = Sh(ps[k].s->conv[i]);
                sg[k].bias = P(ps[k].e->conv_b[i]);
                sg[k].out = ps[k].acts[i];
                sg[k].B = B;
            }
            chk(curla_conv_fwd_multi(sg, np, i == 0 ? a->s2d_sstride : a->act_sstride, i == 0 ? 1.0f / 255.0f : 1.0f,
                                     a->act_sstride, a->pitch, a->S, a->Ho[i], a->Wo[i], i == 0 ? 4 * a->cfg.C : 0, st));
        }
        if (fused && ok()) {
            curla_conv_stack_seg sg[3];
            for (int k = 0; k < np; ++k) {
                const bool keep = ps[k].keep || a->keep_acts;
                sg[k].in = ps[k].acts[0];
                sg[k].w96 = Sh(ps[k].s->conv96);
                for (int l = 0; l < 3; ++l) sg[k].bias[l] = P(ps[k].e->conv_b[l + 1]);
                sg[k].out[0] = keep ? ps[k].acts[1] : nullptr;
                sg[k].out[1] = keep ? ps[k].acts[2] : nullptr;
                sg[k].out[2] = ps[k].acts[3];
                sg[k].B = B;
            }
            chk(curla_conv_stack_fwd(sg, np, a->act_sstride, a->pitch, a->S, Hv3, Wv3, st));
        }
    }
    void conv_stack(const bf16* s2d, const EncP& e, const EncS& s, bf16* const acts[4], int B = 0) {
        const Pass p = {s2d, &e, &s, acts, false};
        conv_stack_multi(&p, 1, B);
    }
    // fc (split-K) + bias + LayerNorm
    void tail_fc(const bf16* act4, long long fc_shadow, int B, float* partial) {
        if (!ok()) return;
        set_launch_tag("gemm_fc_fwd");
        chk(curla_gemm_bf16_seg(act4, a->act_sstride, Sh(fc_shadow), a->Kfc, partial, 64, B, 64, a->Kfc,
                                3, 64, 0, nullptr, 0, nullptr, 0, a->fc_splits, (long long)B * 64, 1.f,
                                a->Kfc / 4, (long long)a->S * 8, 1, st));
        set_launch_tag(nullptr);
    }
    void tail_ln(const float* partial, const EncP& e, TailBuf& t, int B, int apply_tanh, const float* act, bf16* X_out) {
        if (!ok()) return;
        chk(curla_ln_fwd_x(partial, a->fc_splits, (long long)B * 64, P(e.fc_b), P(e.ln_w), P(e.ln_b), B,
                           a->cfg.feature_dim, apply_tanh, t.fc_out, t.z, act, act ? a->cfg.action_dim : 0, X_out, st));
    }
    void tail(const bf16* act4, long long fc_shadow, const EncP& e, TailBuf& t, int B, int apply_tanh = 0,
              const float* act = nullptr, bf16* X_out = nullptr, float* partial = nullptr) {
        if (!partial) partial = a->fc_partial;
        tail_fc(act4, fc_shadow, B, partial);
        tail_ln(partial, e, t, B, apply_tanh, act, X_out);
    }
    // nb MLPs of one shape in one launch each (nb = 2: the critic's Q1 || Q2 on the shared input
    // rows X; curl_sac.py:158-169).  m/s/buf point at nb consecutive descriptors; the strides
    // between the two heads' parameters, shadows and buffers are constant by construction.
    void mlp_fwd_n(const bf16* X, const MlpP* m, const MlpS* s, MlpBuf* buf, int nb, int B) {
        if (!ok()) return;
        const int hid = a->cfg.hidden_dim;
        const long long sP = nb > 1 ? m[1].w0 - m[0].w0 : 0, sS = nb > 1 ? s[1].w0 - s[0].w0 : 0;
        const long long sH = nb > 1 ? buf[1].H1 - buf[0].H1 : 0, sO = nb > 1 ? buf[1].out - buf[0].out : 0;
        chk(curla_gemm_bf16_batched(X, 64, Sh(s[0].w0), 64, buf[0].H1, hid, B, hid, 64, 3, hid, 1, P(m[0].b0), 1,
                                    nullptr, 0, 1.f, nb, 0, sS, sH, sP, 0, st));
        if (!ok()) return;
        chk(curla_gemm_bf16_batched(buf[0].H1, hid, Sh(s[0].w1), hid, buf[0].H2, hid, B, hid, hid, 3, hid, 1, P(m[0].b1), 1,
                                    nullptr, 0, 1.f, nb, sH, sS, sH, sP, 0, st));
        if (!ok()) return;
        chk(curla_head_fwd_batched(buf[0].H2, hid, P(m[0].w2), P(m[0].b2), B, hid, m[0].out, buf[0].out, nb, sH, sP, sO, st));
    }
    // single MLP from fp32 z (+ action): the inference entry points
    void mlp_fwd(const float* z, const float* act, const MlpP& m, const MlpS& s, MlpBuf& buf, int B) {
        if (!ok()) return;
        chk(curla_pack_x(z, act, B, a->cfg.feature_dim, a->cfg.action_dim, buf.X, st));
        mlp_fwd_n(buf.X, &m, &s, &buf, 1, B);
    }
    // backward of nb 3-layer MLPs.  gbase = gradient buffer mirroring the parameter segment that
    // starts at float offset pbase, or nullptr for "input gradient only".  dOut / dX: head 0's
    // pointer + stride to head 1.
    // wr != nullptr: the weight / bias gradients (nothing in the backward chain reads them) are issued on wr's stream,
    // forked behind the kernels that produce dH2 (event e1) and dH1 (event e2); the caller joins wr before the optimizer
    void mlp_bwd_n(const float* dOut, long long sDOut, const bf16* X, const MlpP* m, const MlpS* s, const MlpBuf* buf,
                   int nb, float* gbase, long long pbase, float* dX, long long sDX, Run* wr = nullptr, cudaEvent_t e1 = nullptr,
                   cudaEvent_t e2 = nullptr) {
        if (!ok()) return;
        const int B = a->cfg.batch, hid = a->cfg.hidden_dim, No = m[0].out, in_real = m[0].in_real;
        const long long sP = nb > 1 ? m[1].w0 - m[0].w0 : 0, sS = nb > 1 ? s[1].w0 - s[0].w0 : 0;
        const long long sH = nb > 1 ? buf[1].H1 - buf[0].H1 : 0, sD = (long long)B * hid;
        auto g = [&](long long poff) { return gbase + (poff - pbase); };
        Run& w = (wr && gbase) ? *wr : *this;
        const bool forked = &w != this;
        chk(curla_head_bwd_batched(dOut, P(m[0].w2), buf[0].H2, B, hid, No, a->dH2, nb, sDOut, sP, sH, sD, st));
        if (forked && ok()) { cudaEventRecord(e1, st); cudaStreamWaitEvent(w.st, e1, 0); w.rc = rc; }
        if (gbase && ok()) w.chk(curla_head_wgrad_batched(dOut, buf[0].H2, B, hid, No, g(m[0].w2), g(m[0].b2), nb, sDOut, sH, sP, w.st));
        if (gbase && ok()) {
            w.chk(curla_gemm_bf16_batched(a->dH2, hid, buf[0].H1, hid, g(m[0].w1), hid, hid, hid, B, 0, hid, 0, nullptr, 0,
                                          nullptr, 0, 1.f, nb, sD, sH, sP, 0, 0, w.st));
            if (w.ok()) w.chk(curla_colsum_bf16_batched(a->dH2, B, hid, g(m[0].b1), nb, sD, sP, w.st));
        }
        if (ok()) chk(curla_gemm_bf16_batched(a->dH2, hid, Sh(s[0].w1), hid, a->dH1, hid, B, hid, hid, 1, hid, 1, nullptr, 0,
                                              buf[0].H1, hid, 1.f, nb, sD, sS, sD, 0, sH, st));
        if (forked && ok()) { cudaEventRecord(e2, st); cudaStreamWaitEvent(w.st, e2, 0); }
        if (gbase && ok()) {
            w.chk(curla_gemm_bf16_batched(a->dH1, hid, X, 64, g(m[0].w0), in_real, hid, 64, B, 0, in_real, 0, nullptr, 0,
                                          nullptr, 0, 1.f, nb, sD, 0, sP, 0, 0, w.st));
            if (w.ok()) w.chk(curla_colsum_bf16_batched(a->dH1, B, hid, g(m[0].b0), nb, sD, sP, w.st));
        }
        if (dX && ok()) chk(curla_gemm_bf16_batched(a->dH1, hid, Sh(s[0].w0), 64, dX, 64, B, 64, hid, 1, 64, 0, nullptr, 0,
                                                    nullptr, 0, 1.f, nb, sD, sS, sDX, 0, 0, st));
        if (forked) chk(w.rc);
    }
    // LayerNorm + fc backward (+ conv stack backward when conv==true)
    void enc_bwd(const float* dz_a, const float* dz_b, const TailBuf& t, const EncP& e, long long fc_shadow,
                 const EncS* convs, bf16* const acts[4], const bf16* s2d, float* gbase, long long pbase, bool conv,
                 const std::function<void()>& after_fc_wgrad = nullptr, bool scratch_b = false, Run* wr = nullptr,
                 cudaEvent_t e3 = nullptr) {
        if (!ok()) return;
        const int B = a->cfg.batch, feat = a->cfg.feature_dim;
        auto g = [&](long long poff) { return gbase + (poff - pbase); };
        // (scratch_b: a second set of the small buffers, for a backward that runs beside another one)
        float* const dfc_f32 = scratch_b ? a->dfc_f32_b : a->dfc_f32;
        bf16* const dfc_bf16 = scratch_b ? a->dfc_bf16_b : a->dfc_bf16;
        float* const ln_scratch = scratch_b ? a->ln_scratch_b : a->ln_scratch;
        chk(curla_ln_bwd(dz_a, dz_b, t.fc_out, P(e.ln_w), B, feat, dfc_f32, dfc_bf16, ln_scratch,
                         g(e.ln_w), g(e.ln_b), g(e.fc_b), st));
        // dWfc[feat][Kfc] = dfc^T . act4
        // (wr: on wr's stream, forked behind the LayerNorm backward -- the conv backward does not read it)
        Run& w = wr ? *wr : *this;
        if (wr && ok()) { cudaEventRecord(e3, st); cudaStreamWaitEvent(w.st, e3, 0); w.rc = rc; }
        set_launch_tag("gemm_fc_wgrad");
        if (ok()) w.chk(curla_gemm_bf16_seg(dfc_bf16, 64, acts[3], a->act_sstride, g(e.fc_w), a->Kfc, feat, a->Kfc, B, 0,
                                            a->Kfc, 0, nullptr, 0, nullptr, 0, 1, 0, 1.f,
                                            a->Kfc / 4, (long long)a->S * 8, 2, w.st));
        set_launch_tag(nullptr);
        if (wr) chk(w.rc);
        // everything of this bucket except the conv-layer gradients is final here (data parallel: its
        // all-reduce + Adam start now, beside the conv backward below)
        if (after_fc_wgrad && ok()) after_fc_wgrad();
        if (!conv) return;
        // d(act4) = relu'(act4) * dfc . Wfc
        set_launch_tag("gemm_fc_dgrad");
        if (ok()) chk(curla_gemm_bf16_seg(dfc_bf16, 64, Sh(fc_shadow), a->Kfc, a->dact[3], a->act_sstride, B, a->Kfc, 64, 1,
                                          a->Kfc, 1, nullptr, 0, acts[3], a->act_sstride, 1, 0, 1.f,
                                          a->Kfc / 4, (long long)a->S * 8, 4, st));
        set_launch_tag(nullptr);
        // wgrad first stages interleaved with the dgrad chain; all four deterministic second
        // stages run as ONE launch at the end (nothing before the optimizer reads dW)
        int nparts[4] = {0, 0, 0, 0};
        auto ws = [&](int l) { return a->wgrad_ws + (long long)l * a->wgrad_ws_stride; };
        for (int i = 3; i >= 1 && ok(); --i) {
            chk(curla_conv_wgrad_partial(acts[i - 1], a->act_sstride, a->dact[i], a->act_sstride, ws(i), B, a->pitch, a->S,
                                         a->Ho[i], a->Wo[i], 0, &nparts[i], st));
            if (ok()) chk(curla_conv_dgrad(a->dact[i], a->act_sstride, Sh(convs->conv[i]), acts[i - 1], a->dact[i - 1],
                                           a->act_sstride, B, a->pitch, a->S, a->Ho[i - 1], a->Wo[i - 1], st));
        }
        if (ok()) chk(curla_conv_wgrad_partial(s2d, a->s2d_sstride, a->dact[0], a->act_sstride, ws(0), B, a->pitch, a->S,
                                               a->Ho[0], a->Wo[0], 1, &nparts[0], st));
        if (ok()) {
            float* wsp[4] = {ws(0), ws(1), ws(2), ws(3)};
            const int first[4] = {1, 0, 0, 0}, cin[4] = {a->cfg.C, 32, 32, 32};
            const float sc[4] = {1.0f / 255.0f, 1.f, 1.f, 1.f};
            float* dWp[4]; float* dbp[4];
            for (int l = 0; l < 4; ++l) { dWp[l] = g(e.conv_w[l]); dbp[l] = g(e.conv_b[l]); }
            chk(curla_conv_wgrad_reduce_multi(4, wsp, nparts, first, cin, sc, dWp, dbp, st));
        }
    }
};

// ---- NCCL via dlopen (no link-time dependency; CPU-only hosts can still load the .so)
typedef int (*nccl_allreduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_allgather_t)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*nccl_getid_t)(void*);
struct NcclApi {
    void* lib = nullptr;
    nccl_allreduce_t all_reduce = nullptr;
    nccl_allgather_t all_gather = nullptr;
    nccl_getid_t get_id = nullptr;
    void* init_rank = nullptr;
    int (*comm_destroy)(void*) = nullptr;
    const char* (*err_str)(int) = nullptr;
};
NcclApi g_nccl;
int load_nccl() {
    if (g_nccl.lib) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    CURLA_CHECK(g_nccl.lib, "nccl: cannot dlopen libnccl.so.2 (%s)", dlerror());
    g_nccl.all_reduce = (nccl_allreduce_t)dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.all_gather = (nccl_allgather_t)dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.get_id = (nccl_getid_t)dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.init_rank = dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.comm_destroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.err_str = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    CURLA_CHECK(g_nccl.all_reduce && g_nccl.all_gather && g_nccl.get_id && g_nccl.init_rank, "nccl: missing symbols");
    return 0;
}
enum { NCCL_F32 = 7, NCCL_F64 = 8, NCCL_SUM = 0, NCCL_AVG = 4 };
int all_reduce(curla_agent* a, void* buf, size_t n, int dt, cudaStream_t st) {
    if (a->cfg.world == 1) return 0;
    CURLA_CHECK(a->comm, "update: world>1 but no communicator (curla_agent_init_comm)");
    const int r = g_nccl.all_reduce(buf, buf, n, dt, NCCL_SUM, a->comm, st);
    CURLA_CHECK(r == 0, "ncclAllReduce: %s", g_nccl.err_str ? g_nccl.err_str(r) : "error");
    profile_mark("nccl_all_reduce");
    return 0;
}

}  // namespace

static void destroy_comm(curla_agent* a) {
    if (a->comm && g_nccl.comm_destroy) g_nccl.comm_destroy(a->comm);
    a->comm = nullptr;
}

// Hands the NCCL communicator of `from` (an engine being replaced: same rank / world, another batch
// size) to `to`.  Not a collective: a rank may re-create its engine on its own (curl_sac.py
// _grow_for_inference) without the other ranks taking part.
extern "C" int curla_agent_take_comm(curla_agent* to, curla_agent* from) {
    CURLA_CHECK(to && from && to != from, "take_comm: bad handles");
    CURLA_CHECK(to->cfg.world == from->cfg.world && to->cfg.rank == from->cfg.rank, "take_comm: rank/world differ");
    destroy_comm(to);
    to->comm = from->comm;
    from->comm = nullptr;
    return 0;
}

extern "C" int curla_nccl_unique_id(void* out128) {
    if (load_nccl()) return -1;
    const int r = g_nccl.get_id(out128);
    CURLA_CHECK(r == 0, "ncclGetUniqueId failed (%d)", r);
    return 0;
}
extern "C" int curla_agent_init_comm(curla_agent* a, const void* id128) {
    if (a->cfg.world == 1) return 0;
    if (load_nccl()) return -1;
    destroy_comm(a);
    struct Id { char b[128]; } id;
    memcpy(&id, id128, 128);
    typedef int (*init_t)(void**, int, Id, int);
    const int r = ((init_t)g_nccl.init_rank)(&a->comm, a->cfg.world, id, a->cfg.rank);
    CURLA_CHECK(r == 0, "ncclCommInitRank: %s", g_nccl.err_str ? g_nccl.err_str(r) : "error");
    return 0;
}

extern "C" int curla_agent_refresh_shadows(curla_agent* a, cudaStream_t stream) {
    CURLA_CHECK(a->bound, "agent not bound");
    Run r{a, stream};
    r.pack(a->pack_critic); r.pack(a->pack_actor); r.pack(a->pack_target);
    return r.rc;
}

extern "C" int curla_agent_last_launches(const curla_agent* a) { return (int)a->last_launches; }

// 1: every encoder pass of the update stores its conv-2 / conv-3 activations (--log_param_hist_imgs reads the
// target encoder's, encoder.py:118-130); 0 (default): only the passes a backward follows.  Captured update graphs
// are dropped: the launch arguments change.
extern "C" int curla_agent_set_mailbox(curla_agent* a, float* mailbox_host_mapped) {
    if (mailbox_host_mapped != a->mailbox) {         // a launch argument of the captured graphs
        for (auto& kv : a->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        a->graphs.clear();
    }
    a->mailbox = mailbox_host_mapped;
    return 0;
}

extern "C" int curla_agent_set_keep_acts(curla_agent* a, int on) {
    if ((on != 0) != (a->keep_acts != 0)) {
        for (auto& kv : a->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        a->graphs.clear();
    }
    a->keep_acts = on ? 1 : 0;
    return 0;
}

extern "C" int curla_agent_set_opt_steps(curla_agent* a, int t_critic, int t_actor, int t_alpha, int t_cpc) {
    a->t_critic = t_critic; a->t_actor = t_actor; a->t_alpha = t_alpha; a->t_cpc = t_cpc;
    return 0;
}
extern "C" int curla_agent_get_opt_steps(const curla_agent* a, int* out4) {
    out4[0] = a->t_critic; out4[1] = a->t_actor; out4[2] = a->t_alpha; out4[3] = a->t_cpc;
    return 0;
}

extern "C" int curla_profile_enable(int on) {
    g_prof.on = on != 0;
    g_prof.phases = on == 2;
    return 0;
}
// Synchronises the recorded events and writes "name count total_ms\n" lines (per launch
// site name) into buf; clears the recording.  Returns bytes written or -1.
extern "C" int curla_profile_read(char* buf, int cap) {
    std::map<std::string, std::pair<long long, double>> acc;
    for (size_t i = 1; i < g_prof.marks.size(); ++i) {
        if (!strcmp(g_prof.marks[i].first, "__begin__")) continue;
        cudaError_t e = cudaEventSynchronize(g_prof.marks[i].second);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, g_prof.marks[i - 1].second, g_prof.marks[i].second);
        CURLA_CHECK(e == cudaSuccess, "profile_read: %s", cudaGetErrorString(e));
        auto& slot = acc[g_prof.marks[i].first];
        slot.first += 1; slot.second += ms;
    }
    g_prof.marks.clear();
    g_prof.used = 0;
    int n = 0;
    for (auto& kv : acc) {
        const int w = snprintf(buf + n, cap - n, "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        if (w < 0 || n + w >= cap) break;
        n += w;
    }
    return n;
}

// Side stream (created lazily: needs the CUDA context of the caller's device).  CURLA_STREAMS=1
// keeps everything on the caller's stream; so does the CUDA-event profiler (per-kernel times
// are only meaningful when launches are serialised).
static cudaStream_t side_stream(curla_agent* a, cudaStream_t st) {
    if (a->side_state == 0) {
        const char* e = getenv("CURLA_STREAMS");
        a->side_state = -1;
        if (!(e && e[0] == '1') && cudaStreamCreateWithFlags(&a->side, cudaStreamNonBlocking) == cudaSuccess) {
            bool okev = true;
            for (auto& ev : a->ev) okev = okev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
            if (okev) a->side_state = 1;
        }
    }
    return (a->side_state == 1 && (!g_prof.on || g_prof.phases)) ? a->side : st;
}

// Communication stream (created lazily, world > 1 only).  CURLA_COMM_OVERLAP=0 keeps the collectives
// and the optimizer steps on the caller's stream, in line (so does the CUDA-event profiler).
static cudaStream_t comm_stream(curla_agent* a, cudaStream_t st) {
    if (a->cfg.world == 1) return st;
    if (a->comm_state == 0) {
        const char* e = getenv("CURLA_COMM_OVERLAP");
        a->comm_state = -1;
        if (!(e && e[0] == '0') && cudaStreamCreateWithFlags(&a->comm_st, cudaStreamNonBlocking) == cudaSuccess) {
            bool okev = true;
            for (auto& ev : a->cev) okev = okev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
            if (okev) a->comm_state = 1;
        }
    }
    return (a->comm_state == 1 && (!g_prof.on || g_prof.phases)) ? a->comm_st : st;
}

// ================================================================== the update
// The launch sequence of one update.  dev != nullptr (graph capture): the Adam step counters and the Philox offset
// are read from a->dev_state by the kernels instead of being baked into their parameters.
static int update_body(curla_agent* a, const curla_update_args* u, cudaStream_t st, const int* dev) {
    const auto& c = a->cfg;
    const int* const td_critic = dev ? dev + 0 : nullptr;
    const int* const td_actor = dev ? dev + 1 : nullptr;
    const int* const td_alpha = dev ? dev + 2 : nullptr;
    const int* const td_cpc = dev ? dev + 3 : nullptr;
    const unsigned long long* const off_dev = dev ? reinterpret_cast<const unsigned long long*>(dev + 4) : nullptr;
    const unsigned long long off_next = dev ? 0ull : u->offset * 2, off_cur = dev ? 1ull : u->offset * 2 + 1;
    const int B = c.batch, A = c.action_dim, feat = c.feature_dim;
    const float gs = 1.0f / (float)c.global_batch;
    const long long launches0 = g_launches;
    if (g_prof.on) { g_prof.stream = st; g_prof.mark("__begin__"); }
    Run r{a, st};
    const cudaStream_t cs = comm_stream(a, st);
    const bool overlap = cs != st;                 // world > 1: bucket slices are reduced + stepped on cs
    bool actor_pending = false;
    // [split, n) of a bucket on the communication stream: all-reduce, then Adam on the reduced slice
    auto reduce_step = [&](int ev_fork, float* pbase, float* gbase, float* m, float* v, long long from, long long to,
                           long long double_from, double lr, double beta, int t, const int* t_dev) {
        if (!r.ok()) return;
        cudaEventRecord(a->cev[ev_fork], st);
        cudaStreamWaitEvent(cs, a->cev[ev_fork], 0);
        r.chk(all_reduce(a, gbase + from, (size_t)(to - from), NCCL_F32, cs));
        const long long df = double_from > from ? double_from - from : 0;
        if (r.ok()) r.chk(curla_adam_f32(pbase + from, gbase + from, m + from, v + from, to - from, df, lr, beta, 0.999, 1e-8, t,
                                         t_dev, cs));
    };
    const int ph = u->phases ? u->phases : CURLA_PHASE_ALL;
    const bool do_sac = !u->only_cpc;
    const bool do_critic = do_sac && (ph & CURLA_PHASE_CRITIC);
    const bool do_actor = do_sac && (u->step % c.actor_update_freq == 0) && (ph & CURLA_PHASE_ACTOR);
    const bool do_ema = do_sac && (u->step % c.critic_target_update_freq == 0) && (ph & CURLA_PHASE_EMA);
    const bool do_cpc = !c.pixel_sac && (u->step % c.cpc_update_freq == 0) && (ph & CURLA_PHASE_CPC);
    // the logged scalars go to the host right after the last kernel of this update that writes one
    const int last_writer = do_cpc ? CURLA_PHASE_CPC : (do_actor ? CURLA_PHASE_ACTOR : (do_critic ? CURLA_PHASE_CRITIC : 0));
    // (data parallel: the scalars are rank means of local-batch means -- one 64-byte ncclAvg in front of the publish,
    // on the communication stream when there is one, so the collective's latency stays off the update's critical path)
    auto publish = [&](int phase, Run& rr) {
        if (!(a->mailbox && phase == last_writer && ph == CURLA_PHASE_ALL && rr.ok())) return;
        if (c.world == 1) {
            rr.chk(curla_publish_metrics(a->metrics, a->mailbox, (unsigned)(u->offset + 1), off_dev, rr.st));
            return;
        }
        if (!a->comm) { rr.chk(-1); set_last_error("update: world>1 but no communicator"); return; }
        cudaStream_t ps = rr.st;
        if (overlap) { cudaEventRecord(a->cev[6], rr.st); cudaStreamWaitEvent(cs, a->cev[6], 0); ps = cs; actor_pending = true; }
        const int rc_ = g_nccl.all_reduce(a->metrics, a->metrics_avg, 16, NCCL_F32, NCCL_AVG, a->comm, ps);
        if (rc_ != 0) { rr.chk(-1); set_last_error("ncclAllReduce(metrics) failed (%d)", rc_); return; }
        profile_mark("nccl_all_reduce");
        rr.chk(curla_publish_metrics(a->metrics_avg, a->mailbox, (unsigned)(u->offset + 1), off_dev, ps));
    };

    // ---- sample: gather (+crop) straight into the conv stack's input layout
    auto stage = [&](const float* f32, const uint8_t* frames, const int64_t* h1, const int64_t* w1, bf16* dst) {
        if (!r.ok()) return;
        if (f32) r.chk(curla_f32_to_s2d(f32, c.C, c.H, c.W, B, a->CP1, a->s2d_sstride, dst, st));
        else r.chk(curla_gather_crop_s2d(frames, c.C, c.Hf, c.Wf, u->idxs, h1, w1, B, c.H, c.W, a->CP1, a->s2d_sstride, dst, st));
    };
    NvtxRange nv_update("curla/update");
    nvtxRangePushA("curla/sample");
    const bool do_sample = (ph & CURLA_PHASE_SAMPLE) != 0;
    const bool want_cpc = !c.pixel_sac && (u->step % c.cpc_update_freq == 0);
    const bf16* s2d_pos = u->pos_is_obs ? a->s2d_obs : a->s2d_pos;
    const bool need_pos = want_cpc && !u->pos_is_obs;
    if (do_sample && !u->obs_f32 && !(need_pos && u->pos_f32) && !(do_sac && u->next_f32)) {
        // the replay path: obs / pos / next_obs windows and the batch's action / reward / not_done rows
        // in ONE bulk-staged launch (utils.py:151-166)
        curla_gather_seg gs[3];
        int ng = 0;
        gs[ng++] = {u->obses, u->h1_obs, u->w1_obs, a->s2d_obs};
        if (need_pos) gs[ng++] = {u->obses, u->h1_pos, u->w1_pos, a->s2d_pos};
        if (do_sac) gs[ng++] = {u->next_obses, u->h1_next, u->w1_next, a->s2d_next};
        const curla_gather_rows rows = {u->actions, u->rewards, u->not_dones, a->act_b, a->rew_b, a->nd_b, A};
        r.chk(curla_gather_crop_s2d_multi(gs, ng, u->idxs, c.C, c.Hf, c.Wf, B, c.H, c.W, a->CP1, a->s2d_sstride,
                                          do_sac ? &rows : nullptr, st));
    } else if (do_sample) {
        stage(u->obs_f32, u->obses, u->h1_obs, u->w1_obs, a->s2d_obs);
        if (need_pos) stage(u->pos_f32, u->obses, u->h1_pos, u->w1_pos, a->s2d_pos);
        if (do_sac) {
            stage(u->next_f32, u->next_obses, u->h1_next, u->w1_next, a->s2d_next);
            if (r.ok()) r.chk(curla_gather_rows_f32(u->actions, u->idxs, B, A, a->act_b, st));
            if (r.ok()) r.chk(curla_gather_rows_f32(u->rewards, u->idxs, B, 1, a->rew_b, st));
            if (r.ok()) r.chk(curla_gather_rows_f32(u->not_dones, u->idxs, B, 1, a->nd_b, st));
        }
    }

    nvtxRangePop();
    bool have_p5 = false;
    phase_mark("P01_gather", st);
    if (do_critic) {
        NvtxRange nv("curla/critic");
        // ---------------- update_critic (curl_sac.py:349-371)
        // The tail of each pass (fc split-K GEMM, LayerNorm, MLP heads: latency-bound kernels that
        // fill a fraction of the GPU) runs on the side stream while the main stream already runs
        // the next pass's conv stack; fork/join with events, no host synchronisation.
        const cudaStream_t ss = side_stream(a, st);
        const bool forked = ss != st;
        Run r2{a, ss};
        auto fork = [&](int k) { if (forked) { cudaEventRecord(a->ev[k], st); cudaStreamWaitEvent(ss, a->ev[k], 0); } };
        const int mm = merge_mode();
        // F1: actor(next_obs) -> a', log_pi';  F2: critic_target(next_obs, a');  F3: critic(obs, action)
        bf16* const* act2 = (forked || mm) ? a->actC : a->actB;      // F1's tail may still be reading actB[3]
        const Run::Pass p1 = {a->s2d_next, &a->enc_critic, &a->s_critic, a->actB, false};
        const Run::Pass p2 = {a->s2d_next, &a->enc_target, &a->s_target, act2, false};
        const Run::Pass p3 = {a->s2d_obs, &a->enc_critic, &a->s_critic, a->actA, true};
        auto tail1 = [&]() {
            r2.tail(a->actB[3], a->s_actor_fc, a->enc_actor, a->t_p1, B, 0, nullptr, a->m_p1.X, a->fc_partial2);
            r2.mlp_fwd_n(a->m_p1.X, &a->trunk_actor, &a->s_trunk, &a->m_p1, 1, B);
            if (r2.ok()) r2.chk(curla_policy_fwd_rows_dyn(a->t_out1, u->noise_next, u->seed, off_next, off_dev, c.rank * B, B, A,
                                                     (float)c.log_std_min, (float)c.log_std_max, 1, 1, a->mu_scratch, a->a_next,
                                                     a->logpi_next, a->ls1, nullptr, ss));
        };
        auto tail2 = [&]() {
            r2.tail(act2[3], a->s_target.fc, a->enc_target, a->t_p2, B, 0, a->a_next, a->m_p2q[0].X, a->fc_partial2);
            r2.mlp_fwd_n(a->m_p2q[0].X, a->q_target, a->sq_target, a->m_p2q, 2, B);
        };
        if (mm == 0) {
            r.conv_stack_multi(&p1, 1);
            fork(0);
            r2.rc = r.rc;
            tail1();
            r.conv_stack_multi(&p2, 1);
            fork(1);
            tail2();
            r.conv_stack_multi(&p3, 1);
        } else if (mm == 1) {
            const Run::Pass ps[2] = {p1, p2};
            r.conv_stack_multi(ps, 2);
            fork(0);
            r2.rc = r.rc;
            tail1();
            tail2();
            r.conv_stack_multi(&p3, 1);
        } else {
            const Run::Pass ps[3] = {p1, p2, p3};
            r.conv_stack_multi(ps, 3);
            phase_mark("P02_conv_forward_F1+F2+F3", st);
            fork(0);
            r2.rc = r.rc;
            tail1();
            tail2();
        }
        r.tail(a->actA[3], a->s_critic.fc, a->enc_critic, a->t_p3, B, 0, a->act_b, a->m_p3q[0].X);
        r.mlp_fwd_n(a->m_p3q[0].X, a->q_critic, a->sq_critic, a->m_p3q, 2, B);
        if (forked) { cudaEventRecord(a->ev[2], ss); cudaStreamWaitEvent(st, a->ev[2], 0); }
        r.chk(r2.rc);
        if (r.ok()) r.chk(curla_critic_loss(a->tq[0], a->tq[1], a->logpi_next, a->rew_b, a->nd_b, a->log_alpha, (float)c.discount,
                                            a->q3[0], a->q3[1], B, gs, a->target_q, a->dq[0], a->dq[1], a->metrics, st));
        publish(CURLA_PHASE_CRITIC, r);
        // backward
        phase_mark("P03_tails_F1_F3_+_critic_loss", st);
        float* gC = a->G + a->g_critic;
        // The weight gradients of the Q heads and of the encoder fc (six small kernels + one HBM-bound GEMM that nothing
        // in the backward chain reads) go to the side stream, idle since the tails joined: the chain
        // head -> trunk dgrad -> dX -> LayerNorm -> fc dgrad -> conv backward no longer waits for them.  CURLA_WGRAD_SIDE=0: in line.
        static const bool wgrad_side = [] { const char* e = getenv("CURLA_WGRAD_SIDE"); return !(e && e[0] == '0'); }();
        Run rw{a, ss};
        Run* const wr = (forked && wgrad_side) ? &rw : nullptr;
        r.mlp_bwd_n(a->dq[0], a->dq[1] - a->dq[0], a->m_p3q[0].X, a->q_critic, a->sq_critic, a->m_p3q, 2, gC, a->off_critic,
                    a->dX[0], a->dX[1] - a->dX[0], wr, a->ev[6], a->ev[7]);
        phase_mark("P04_Q_heads_backward_chain", st);
        const int tc_ = ++a->t_critic;
        // critic bucket = [conv w,b x4 | fc_w fc_b ln_w ln_b | Q1 | Q2]: everything from fc_w on (99 % of the
        // bytes) is final after the fc weight gradient, i.e. BEFORE the conv backward (0.7 ms at the
        // default batch) starts: its all-reduce and Adam overlap that
        const long long split1 = a->enc_critic.fc_w - a->off_critic;
        auto early1 = [&]() {
            if (wr) cudaEventRecord(a->ev[8], ss);                     // the side stream's weight gradients are complete
            if (overlap) {
                if (wr) cudaStreamWaitEvent(cs, a->ev[8], 0);
                reduce_step(0, a->P + a->off_critic, gC, a->Ad + a->a_m1, a->Ad + a->a_v1, split1, a->n_critic, a->n_critic,
                            c.critic_lr, c.critic_beta, tc_, td_critic);
            }
        };
        r.enc_bwd(a->dX[0], a->dX[1], a->t_p3, a->enc_critic, a->s_critic.fc, &a->s_critic, a->actA, a->s2d_obs, gC,
                  a->off_critic, !c.detach_encoder, early1, false, wr, a->ev[9]);
        phase_mark("P05_critic_LayerNorm_+_fc_dgrad_+_conv_backward", st);
        if (wr) cudaStreamWaitEvent(st, a->ev[8], 0);                  // join (long complete: the conv backward ran meanwhile)
        if (overlap) {
            reduce_step(1, a->P + a->off_critic, gC, a->Ad + a->a_m1, a->Ad + a->a_v1, 0, split1, a->n_critic,
                        c.critic_lr, c.critic_beta, tc_, td_critic);
            cudaEventRecord(a->cev[2], cs);
            cudaStreamWaitEvent(st, a->cev[2], 0);
        } else {
            if (r.ok()) r.chk(all_reduce(a, gC, (size_t)a->n_critic, NCCL_F32, st));
            if (r.ok()) r.chk(curla_adam_f32(a->P + a->off_critic, gC, a->Ad + a->a_m1, a->Ad + a->a_v1, a->n_critic, a->n_critic,
                                             c.critic_lr, c.critic_beta, 0.999, 1e-8, tc_, td_critic, st));
        }
        r.pack(a->pack_critic);
    }
    const int mm = merge_mode();
    phase_mark("P06_critic_join_+_Adam_+_shadow_pack", st);
    auto ema = [&]() {
        // soft_update_params x3 (curl_sac.py:442-445): encoder tau on [0,n_enc), critic tau on Q1,Q2
        if (r.ok()) r.chk(curla_ema_f32(a->P + a->off_target, a->P + a->off_critic, a->n_critic, a->n_enc, c.encoder_tau,
                                        c.critic_tau, st));
        r.pack(a->pack_target);
    };
    // The reference runs the EMA after the actor step (curl_sac.py:439-445); it reads the critic
    // and writes the target only, neither of which the actor step touches, so running it first
    // is the same computation -- and lets the post-EMA key pass F7 share conv launches with F4.
    const bool ema_first = mm != 0;
    if (do_sac && do_ema && ema_first) { NvtxRange nv("curla/ema"); ema(); }
    phase_mark("P07_EMA_+_target_shadow_pack", st);
    const cudaStream_t ss7 = do_cpc ? side_stream(a, st) : st;
    const bool forked7 = ss7 != st;
    const Run::Pass p_anchor = {a->s2d_obs, &a->enc_critic, &a->s_critic, a->actA, true};    // F4 == F5 == F6 conv part
    const Run::Pass p_key = {s2d_pos, &a->enc_target, &a->s_target, a->actB, false};          // F7
    bool key_done = false, join7 = false, f5_early = false;
    // tail of one pass on the side stream (fc_partial2 is the side stream's split-K buffer)
    // all-gather of the CURL keys (the one real exchange step of the update): issued on the stream that
    // produced them, as early as they exist -- before the actor bucket's all-reduce is queued on the
    // communicator, which runs its collectives in issue order
    bool keys_gathered = false;
    auto gather_keys = [&](cudaStream_t s_) {
        if (c.world == 1 || !r.ok()) return;
        if (!a->comm) { r.chk(-1); set_last_error("update: world>1 but no communicator"); return; }
        const int rc_ = g_nccl.all_gather(a->t_p7.z, a->z_pos_all, (size_t)B * 64, NCCL_F32, a->comm, s_);
        if (rc_ != 0) { r.chk(-1); set_last_error("ncclAllGather failed (%d)", rc_); return; }
        profile_mark("nccl_all_gather");
        keys_gathered = true;
    };
    auto keys_gathered_or_local = [&]() { return c.world == 1 || keys_gathered; };
    auto side_tail = [&](const bf16* act4, long long fc_shadow, const EncP& e, TailBuf& t, bool keys) {
        Run r2{a, ss7};
        r2.rc = r.rc;
        if (forked7) { cudaEventRecord(a->ev[3], st); cudaStreamWaitEvent(ss7, a->ev[3], 0); }
        r2.tail(act4, fc_shadow, e, t, B, 0, nullptr, nullptr, a->fc_partial2);
        r.chk(r2.rc);
        if (keys) gather_keys(ss7);
        if (forked7) { cudaEventRecord(a->ev[2], ss7); join7 = true; }
    };
    // CURL contraction (logits, cross-entropy, dz_anchor, dW): needs the anchor latents (t_p5) and the gathered keys.
    // On a step that also runs the actor update it is issued on the side stream right after the actor loss and runs
    // BESIDE the actor's backward + Adam (both are chains of small latency-bound kernels: the contraction fills 32 of
    // the 148 SMs at a global batch of 512); CURLA_CURL_OVERLAP=0 keeps it on the main stream after the actor step.
    bool curl_done = false;
    auto curl_contract = [&](Run& rr) {
        const float* zpos = c.world > 1 ? a->z_pos_all : a->t_p7.z;
        if (rr.ok()) rr.chk(curla_curl_fwd_bwd(a->t_p5.z, zpos, a->P + a->off_W, B, c.global_batch, feat, c.rank * B, gs, a->curl_ws,
                                               a->metrics + 6, a->dz_curl, a->G + a->g_cpc, nullptr, rr.st));
        publish(CURLA_PHASE_CPC, rr);
        curl_done = true;
    };
    static const bool curl_overlap = [] { const char* e = getenv("CURLA_CURL_OVERLAP"); return !(e && e[0] == '0'); }();
    static const bool actor_on_side = [] { const char* e = getenv("CURLA_ACTOR_SIDE"); return !(e && e[0] == '0'); }();
    bool actor_side_pending = false;
    if (do_sac) {
        if (do_actor) {
            NvtxRange nv("curla/actor_alpha");
            // ---------------- update_actor_and_alpha (curl_sac.py:373-404)
            // F4: conv_theta'(obs) shared by actor(obs), critic(obs, pi) and the CURL anchor
            if (mm && do_cpc) {
                const Run::Pass ps[2] = {p_anchor, p_key};
                r.conv_stack_multi(ps, 2);
                phase_mark("P08_conv_forward_anchor_+_key", st);
                side_tail(a->actB[3], a->s_target.fc, a->enc_target, a->t_p7, true);   // keys: beside the actor step
                key_done = true;
                // F5 = critic.encoder(obs) for Q(obs, pi): its fc GEMM depends on the conv stack only (pi enters at the
                // LayerNorm's concat), so it runs on the side stream behind the key tail while the main stream is in the
                // actor's fc -> LayerNorm -> trunk -> policy chain (CURLA_F5_EARLY=0: in line)
                static const bool f5_on = [] { const char* e = getenv("CURLA_F5_EARLY"); return !(e && e[0] == '0'); }();
                if (f5_on && forked7 && r.ok()) {
                    Run r5{a, ss7};
                    r5.rc = r.rc;
                    r5.tail_fc(a->actA[3], a->s_critic.fc, B, a->fc_partial2);
                    r.chk(r5.rc);
                    cudaEventRecord(a->ev[10], ss7);
                    f5_early = true;
                }
            } else {
                r.conv_stack_multi(&p_anchor, 1);
            }
            r.tail(a->actA[3], a->s_actor_fc, a->enc_actor, a->t_p4, B, 0, nullptr, a->m_p4.X);
            r.mlp_fwd_n(a->m_p4.X, &a->trunk_actor, &a->s_trunk, &a->m_p4, 1, B);
            if (r.ok()) r.chk(curla_policy_fwd_rows_dyn(a->t_out4, u->noise_cur, u->seed, off_cur, off_dev, c.rank * B, B, A,
                                                   (float)c.log_std_min, (float)c.log_std_max, 1, 1, a->mu_scratch, a->pi4, a->logpi4,
                                                   a->ls4, a->noise4, st));
            if (f5_early) {
                // (its fc GEMM ran on the side stream behind the key tail; the LayerNorm needs pi, hence it is here)
                cudaStreamWaitEvent(st, a->ev[10], 0);
                r.tail_ln(a->fc_partial2, a->enc_critic, a->t_p5, B, 0, a->pi4, a->m_p5q[0].X);
            } else {
                r.tail(a->actA[3], a->s_critic.fc, a->enc_critic, a->t_p5, B, 0, a->pi4, a->m_p5q[0].X);
            }
            have_p5 = true;
            r.mlp_fwd_n(a->m_p5q[0].X, a->q_critic, a->sq_critic, a->m_p5q, 2, B);
            if (r.ok()) r.chk(curla_actor_loss(a->logpi4, a->q5[0], a->q5[1], a->ls4, B, A, a->log_alpha, (float)c.target_entropy, gs,
                                               a->dq[0], a->dq[1], a->glogpi, a->g_log_alpha, a->metrics, st));
            publish(CURLA_PHASE_ACTOR, r);
            phase_mark("P09_actor_forward_tails_+_loss", st);
            // Two ways to use the side stream from here on (it already holds the key tail + all-gather):
            //   actor_side (default): the actor's whole backward + optimizer step runs there, beside the CURL contraction
            //     and the contrastive backward on the main stream -- nothing later in THIS update reads the actor's own
            //     parameters, its gradients or log_alpha (the CURL phase uses the critic / target encoders);
            //   CURLA_ACTOR_SIDE=0: the CURL contraction runs there beside the actor's backward (CURLA_CURL_OVERLAP=0: neither).
            const bool can_fork = key_done && forked7 && r.ok();
            const bool actor_side = can_fork && actor_on_side;
            if (can_fork && !actor_side && curl_overlap && keys_gathered_or_local()) {
                cudaEventRecord(a->ev[4], st);
                cudaStreamWaitEvent(ss7, a->ev[4], 0);
                Run rc{a, ss7};
                rc.rc = r.rc;
                curl_contract(rc);
                r.chk(rc.rc);
                cudaEventRecord(a->ev[2], ss7);          // re-recorded: the join below now also covers the contraction
            }
            const cudaStream_t bs = actor_side ? ss7 : st;            // the stream of the actor's backward
            if (actor_side) { cudaEventRecord(a->ev[4], st); cudaStreamWaitEvent(ss7, a->ev[4], 0); }
            Run rb{a, bs};
            rb.rc = r.rc;
            rb.mlp_bwd_n(a->dq[0], a->dq[1] - a->dq[0], a->m_p5q[0].X, a->q_critic, a->sq_critic, a->m_p5q, 2, nullptr, 0,
                         a->dX[0], a->dX[1] - a->dX[0]);
            if (rb.ok()) rb.chk(curla_policy_bwd(a->dX[0], a->dX[1], feat, a->glogpi, a->t_out4, a->noise4, a->pi4, a->ls4, B, A,
                                                 (float)c.log_std_min, (float)c.log_std_max, a->dt4, bs));
            float* gA = a->G + a->g_actor;
            rb.mlp_bwd_n(a->dt4, 0, a->m_p4.X, &a->trunk_actor, &a->s_trunk, &a->m_p4, 1, gA, a->off_actor, a->dXa, 0);
            rb.enc_bwd(a->dXa, nullptr, a->t_p4, a->enc_actor, a->s_actor_fc, nullptr, a->actA, nullptr, gA, a->off_actor, false, nullptr,
                       actor_side);
            // with world > 1 the reduce -> Adam -> shadow pack chain runs on the communication stream and is joined at
            // the end of the update
            const cudaStream_t as = overlap ? cs : bs;
            Run ra{a, as};
            ra.rc = rb.rc;
            if (overlap) { cudaEventRecord(a->cev[3], bs); cudaStreamWaitEvent(cs, a->cev[3], 0); actor_pending = true; }
            if (ra.ok()) ra.chk(all_reduce(a, gA, (size_t)a->n_actor, NCCL_F32, as));
            if (ra.ok()) ra.chk(all_reduce(a, a->g_log_alpha, 1, NCCL_F64, as));
            if (ra.ok()) ra.chk(curla_adam_f32(a->P + a->off_actor, gA, a->Ad + a->a_m2, a->Ad + a->a_v2, a->n_actor, a->n_actor,
                                               c.actor_lr, c.actor_beta, 0.999, 1e-8, ++a->t_actor, td_actor, as));
            ra.pack(a->pack_actor);
            if (ra.ok()) ra.chk(curla_adam_f64_scalar(a->log_alpha, a->g_log_alpha, a->alpha_state, c.alpha_lr, c.alpha_beta, 0.999,
                                                      1e-8, ++a->t_alpha, td_alpha, as));
            if (actor_side && !overlap) { cudaEventRecord(a->ev[5], ss7); actor_side_pending = true; }
            r.chk(ra.rc);
        }
        if (do_ema && !ema_first) ema();
    }

    if (do_cpc) {
        NvtxRange nv("curla/cpc");
        // ---------------- update_cpc (curl_sac.py:406-423)
        if (!have_p5) {
            // F6: anchor through the current critic encoder (its tail on the side stream);
            // F7: keys through the (post-EMA) target encoder, no grad
            if (mm) {
                const Run::Pass ps[2] = {p_anchor, p_key};
                r.conv_stack_multi(ps, 2);
                phase_mark("P08_conv_forward_anchor_+_key", st);
            } else {
                r.conv_stack_multi(&p_anchor, 1);
            }
            side_tail(a->actA[3], a->s_critic.fc, a->enc_critic, a->t_p5, false);
        }
        if (!key_done) {
            if (!(mm && !have_p5)) r.conv_stack_multi(&p_key, 1);
            r.tail(a->actB[3], a->s_target.fc, a->enc_target, a->t_p7, B);
        }
        phase_mark("P10_cpc_main_stream_tails", st);
        if (join7) cudaStreamWaitEvent(st, a->ev[2], 0);
        float* gK = a->G + a->g_cpc;
        if (!curl_done) {
            if (!keys_gathered) gather_keys(st);
            curl_contract(r);
        }
        phase_mark("P11_CURL_contraction_+_wait_for_keys", st);
        // g_cpc mirrors [W | critic.encoder]: encoder grads start at n_W
        // encoder_optimizer.step(); cpc_optimizer.step(): encoder twice, W once (double_from = n_W).
        // Bucket = [W | conv w,b x4 | fc_w fc_b ln_w ln_b]: the fc/ln tail is final after the fc weight gradient
        const int tk_ = ++a->t_cpc;
        const long long split3 = a->n_W + (a->enc_critic.fc_w - a->off_critic);
        // (the fc weight gradient off the backward chain, as in the critic phase: on the side stream -- behind the actor's
        // backward when that runs there -- joined before the optimizer step)
        static const bool wgrad_side3 = [] { const char* e = getenv("CURLA_WGRAD_SIDE"); return !(e && e[0] == '0'); }();
        Run rw3{a, ss7};
        Run* const wr3 = (forked7 && wgrad_side3) ? &rw3 : nullptr;
        auto early3 = [&]() {
            if (wr3) cudaEventRecord(a->ev[8], ss7);
            if (overlap) {
                if (wr3) cudaStreamWaitEvent(cs, a->ev[8], 0);
                reduce_step(4, a->P + a->off_W, gK, a->Ad + a->a_m3, a->Ad + a->a_v3, split3, a->n_cpc, a->n_W,
                            c.encoder_lr, 0.9, tk_, td_cpc);
            }
        };
        r.enc_bwd(a->dz_curl, nullptr, a->t_p5, a->enc_critic, a->s_critic.fc, &a->s_critic, a->actA, a->s2d_obs,
                  gK + a->n_W, a->off_critic, true, early3, false, wr3, a->ev[9]);
        if (wr3) cudaStreamWaitEvent(st, a->ev[8], 0);
        phase_mark("P12_cpc_LayerNorm_+_fc_dgrad_+_conv_backward", st);
        if (overlap) {
            reduce_step(5, a->P + a->off_W, gK, a->Ad + a->a_m3, a->Ad + a->a_v3, 0, split3, a->n_W, c.encoder_lr, 0.9, tk_, td_cpc);
            cudaEventRecord(a->cev[2], cs);
            cudaStreamWaitEvent(st, a->cev[2], 0);
            actor_pending = false;                 // cs is in order: the actor chain issued before is complete as well
        } else {
            if (r.ok()) r.chk(all_reduce(a, gK, (size_t)a->n_cpc, NCCL_F32, st));
            if (r.ok()) r.chk(curla_adam_f32(a->P + a->off_W, gK, a->Ad + a->a_m3, a->Ad + a->a_v3, a->n_cpc, a->n_W, c.encoder_lr, 0.9,
                                             0.999, 1e-8, tk_, td_cpc, st));
        }
        r.pack(a->pack_critic_enc);
        phase_mark("P13_cpc_Adam_+_shadow_pack", st);
    }
    if (actor_pending) {                           // join: everything of this update is complete when `st` is
        cudaEventRecord(a->cev[2], cs);
        cudaStreamWaitEvent(st, a->cev[2], 0);
    }
    if (actor_side_pending) cudaStreamWaitEvent(st, a->ev[5], 0);
    phase_mark("P14_join_side_streams", st);
    a->last_launches = g_launches - launches0;
    return r.rc;
}

// Whole-update CUDA graph.  An update variant is everything that shapes the launch sequence: which phases the step
// runs (actor / EMA / CPC frequencies, only_cpc), every argument pointer (the replay arrays, the index staging slot)
// and the noise seed.  A variant runs eagerly the first time it is seen (kernel attributes, streams, NCCL
// connections are set up outside any capture), is captured on its second appearance and replayed from then on:
// one cudaGraphLaunch (+ the one-thread curla_set_dev_state launch that carries this update's Adam step counters
// and Philox offset) instead of ~90 kernel launches and a dozen event operations issued by the host.
// CURLA_GRAPH=0 keeps every update eager.  Not captured: teacher-forced calls (phases mask, injected noise),
// pre-augmented float inputs (their buffers change per call), profiled runs, callers that are capturing themselves.
static bool graph_on(curla_agent* a) {
    if (a->graph_mode < 0) { const char* e = getenv("CURLA_GRAPH"); a->graph_mode = (e && e[0] == '0') ? 0 : 1; }
    return a->graph_mode == 1;
}

extern "C" int curla_agent_update(curla_agent* a, const curla_update_args* u, cudaStream_t st) {
    CURLA_CHECK(a->bound, "agent not bound");
    const auto& c = a->cfg;
    bool eligible = graph_on(a) && !g_prof.on && (u->phases == 0 || u->phases == CURLA_PHASE_ALL) && !u->noise_next &&
                    !u->noise_cur && !u->obs_f32 && !u->next_f32 && !u->pos_f32;
    if (eligible) {
        cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cst) != cudaSuccess || cst != cudaStreamCaptureStatusNone) eligible = false;
    }
    if (!eligible) return update_body(a, u, st, nullptr);

    const bool do_sac = !u->only_cpc;
    const bool do_actor = do_sac && (u->step % c.actor_update_freq == 0);
    const bool do_ema = do_sac && (u->step % c.critic_target_update_freq == 0);
    const bool do_cpc = !c.pixel_sac && (u->step % c.cpc_update_freq == 0);
    curla_update_args k = *u;
    k.step = (do_actor ? 1 : 0) | (do_ema ? 2 : 0) | (do_cpc ? 4 : 0);
    k.offset = 0;
    k.phases = 0;
    const std::string key(reinterpret_cast<const char*>(&k), sizeof(k));
    auto& ent = a->graphs[key];                      // value-initialised on first sight: {nullptr, 0, 0}
    if (ent.seen++ == 0) return update_body(a, u, st, nullptr);
    const int t0[4] = {a->t_critic, a->t_actor, a->t_alpha, a->t_cpc};
    if (!ent.exec) {
        if (a->graphs.size() > 64) {                 // a caller cycling through pointers: stop collecting graphs
            a->graph_mode = 0;
            return update_body(a, u, st, nullptr);
        }
        cudaGraph_t graph = nullptr;
        if (!a->cap_st && cudaStreamCreateWithFlags(&a->cap_st, cudaStreamNonBlocking) != cudaSuccess) {
            a->cap_st = nullptr; a->graph_mode = 0; cudaGetLastError();
            return update_body(a, u, st, nullptr);
        }
        cudaError_t e = cudaStreamBeginCapture(a->cap_st, cudaStreamCaptureModeRelaxed);
        CURLA_CHECK(e == cudaSuccess, "update: cudaStreamBeginCapture: %s", cudaGetErrorString(e));
        const int rc = update_body(a, u, a->cap_st, a->dev_state);
        e = cudaStreamEndCapture(a->cap_st, &graph);
        a->t_critic = t0[0]; a->t_actor = t0[1]; a->t_alpha = t0[2]; a->t_cpc = t0[3];     // nothing has run yet
        if (rc) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
        CURLA_CHECK(e == cudaSuccess && graph, "update: cudaStreamEndCapture: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&ent.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { ent.exec = nullptr; set_last_error("update: cudaGraphInstantiate: %s", cudaGetErrorString(e)); return -1; }
        ent.launches = a->last_launches;
    }
    // this update's scalars, then the replay
    const int t1[4] = {t0[0] + (do_sac ? 1 : 0), t0[1] + (do_actor ? 1 : 0), t0[2] + (do_actor ? 1 : 0), t0[3] + (do_cpc ? 1 : 0)};
    if (curla_set_dev_state(a->dev_state, t1[0], t1[1], t1[2], t1[3], u->offset * 2, st)) return -1;
    const cudaError_t e = cudaGraphLaunch(ent.exec, st);
    CURLA_CHECK(e == cudaSuccess, "update: cudaGraphLaunch: %s", cudaGetErrorString(e));
    a->t_critic = t1[0]; a->t_actor = t1[1]; a->t_alpha = t1[2]; a->t_cpc = t1[3];
    a->last_launches = ent.launches + 1;
    return 0;
}

// ================================================================== inference entry points
extern "C" int curla_agent_encode(curla_agent* a, int net, const void* obs_s2d, int B,
                                  int apply_tanh, float* z_out, cudaStream_t st) {
    CURLA_CHECK(a->bound, "agent not bound");
    CURLA_CHECK(B >= 1 && B <= a->cfg.batch, "encode: B must be in [1, batch]");
    Run r{a, st};
    const EncP& e = net == 0 ? a->enc_actor : (net == 1 ? a->enc_critic : a->enc_target);
    const EncS& s = net == 2 ? a->s_target : a->s_critic;
    const long long fc = net == 0 ? a->s_actor_fc : (net == 1 ? a->s_critic.fc : a->s_target.fc);
    r.conv_stack((const bf16*)obs_s2d, e, s, a->actB, B);      // the first B samples only
    TailBuf t{a->t_p1.fc_out, z_out};
    r.tail(a->actB[3], fc, e, t, B, apply_tanh);
    return r.rc;
}

extern "C" int curla_agent_actor_head(curla_agent* a, const float* z, int B, const float* noise,
                                      unsigned long long seed, unsigned long long offset,
                                      int compute_pi, int compute_log_pi, float* mu, float* pi,
                                      float* log_pi, float* log_std, cudaStream_t st) {
    CURLA_CHECK(a->bound, "agent not bound");
    CURLA_CHECK(B >= 1 && B <= a->cfg.batch, "actor_head: B must be in [1, batch]");
    Run r{a, st};
    r.mlp_fwd(z, nullptr, a->trunk_actor, a->s_trunk, a->m_p1, B);
    if (r.ok()) r.chk(curla_policy_fwd(a->t_out1, noise, seed, offset, B, a->cfg.action_dim, (float)a->cfg.log_std_min,
                                       (float)a->cfg.log_std_max, compute_pi, compute_log_pi, mu, pi, log_pi, log_std, nullptr, st));
    return r.rc;
}

extern "C" int curla_agent_q_heads(curla_agent* a, int net, const float* z, const float* action,
                                   int B, float* q1, float* q2, cudaStream_t st) {
    CURLA_CHECK(a->bound, "agent not bound");
    CURLA_CHECK(B >= 1 && B <= a->cfg.batch, "q_heads: B must be in [1, batch]");
    Run r{a, st};
    const MlpP* m = net == 2 ? a->q_target : a->q_critic;
    const MlpS* s = net == 2 ? a->sq_target : a->sq_critic;
    float* outs[2] = {q1, q2};
    for (int k = 0; k < 2; ++k) {
        MlpBuf buf = a->m_p2q[k];
        buf.out = outs[k];
        r.mlp_fwd(z, action, m[k], s[k], buf, B);
    }
    return r.rc;
}
