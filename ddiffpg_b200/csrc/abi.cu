// extern "C" entry points of libddiffpg_b200.so (declared in include/ddiffpg_b200.h).
#include <string.h>
#include "actor_layout.cuh"
#include "q_layout.cuh"

namespace ddp {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// implemented in the kernel translation units
int pack_actor_fp32(const ActorLayout& L, const float* const p[12], float* out, bool fma_operands, cudaStream_t st);
int pack_actor_tc(const ActorLayout& L, const float* const p[12], void* packed, bool sampler, bool train, cudaStream_t st);
int actor_sample_fma(const ActorLayout& L, const float* pk, const float* state, const float* noise, float* out,
                     long B, const ExplNoise& expl, cudaStream_t st);
int actor_sample_tc(const ActorLayout& L, const void* packed, const float* state, const float* noise, float* out,
                    long B, const ExplNoise& expl, void* ws, size_t ws_bytes, cudaStream_t st);
size_t actor_sample_tc_workspace(const ActorLayout& L, long B);
size_t actor_train_workspace(const ActorLayout& L, long B);
int actor_train_fma(const ActorLayout& L, const float* pk, const float* const p[12], const float* state,
                    const float* action, const float* noise, const int64_t* t, float inv_count, float* loss_out,
                    float* grads, long B, void* ws, size_t ws_bytes, cudaStream_t st);
int pack_actor_train_tc(const ActorLayout& L, const float* const p[12], void* packed, cudaStream_t st);
size_t actor_train_tc_workspace(const ActorLayout& L, long B);
int actor_train_tc(const ActorLayout& L, const void* packed, const float* const p[12], const float* state,
                   const float* action, const float* noise, const int64_t* t, float inv_count, float* loss_out,
                   float* grads, long B, void* ws, size_t ws_bytes, cudaStream_t st, const cudaEvent_t* group_ev);
int clip_adamw(float* p, float* g, float* m, float* v, size_t n, int step, float lr, float b1, float b2, float eps,
               float wd, float max_norm, float* norm_out, float* scratch, cudaStream_t st);
int clip_adamw_dev(float* p, float* g, float* m, float* v, size_t n, int* step_dev, float lr, float b1, float b2,
                   float eps, float wd, float max_norm, float* norm_out, float* scratch, cudaStream_t st);
int pack_q_fp32(const QLayout& L, const float* const p[], float* out, cudaStream_t st, bool bias_only);
int q_forward_fma(const QLayout& L, const float* pk, const int64_t* seg_off, const float* obs, const float* act,
                  float* qmin, float* p1, float* p2, float* dq_da, long B, cudaStream_t st);
size_t q_ascent_workspace(const QLayout& L, long B, int iters);
int q_ascent_fma(const QLayout& L, const float* pk, const int64_t* seg_off, const int64_t* seg_cnt, const float* obs,
                 float* action, int iters, float lr, float b1, float b2, float eps, float max_norm, float lim,
                 float* mean_abs, float* gnorm_out, long B, void* ws, size_t ws_bytes, cudaStream_t st,
                 ddp_gsq_reduce_fn reduce, void* reduce_user);

int pack_q_tc(const QLayout& L, const float* const p[], void* packed, cudaStream_t st);
size_t q_tc_workspace(const QLayout& L, long B, int iters);
int q_forward_tc(const QLayout& L, const void* packed, const int64_t* seg_off, const float* obs, const float* act,
                 float* qmin, float* p1, float* p2, float* dq_da, long B, void* ws, size_t ws_bytes, cudaStream_t st);
int q_ascent_tc(const QLayout& L, const void* packed, const int64_t* seg_off, const int64_t* seg_cnt, const float* obs,
                float* action, int iters, float lr, float b1, float b2, float eps, float max_norm, float lim,
                float* mean_abs, float* gnorm_out, long B, void* ws, size_t ws_bytes, cudaStream_t st,
                ddp_gsq_reduce_fn reduce, void* reduce_user);

size_t q_critic_train_workspace(const QLayout& L, long B);
size_t q_grad_count(const QLayout& L);
int q_critic_train_fma(const QLayout& L, const float* pk, const float* pk_target, const float* obs, const float* act,
                       const float* next_obs, const float* next_act, const float* reward, const float* done, float gamma,
                       float* loss_out, float* grads, long B, void* ws, size_t ws_bytes, cudaStream_t st);

size_t q_critic_train_tc_workspace(const QLayout& L, long B);
int q_critic_train_tc(const QLayout& L, const void* packed, const void* packed_target, const float* obs, const float* act,
                      const float* next_obs, const float* next_act, const float* reward, const float* done, float gamma,
                      float* loss_out, float* grads, long B, void* ws, size_t ws_bytes, cudaStream_t st);

size_t rnd_grad_count(const QLayout& L);
size_t rnd_train_workspace(const QLayout& L, long B);
int rnd_novelty_fma(const QLayout& L, const float* pk, const float* x, float* novelty, float* pred, float* target,
                    long B, cudaStream_t st);
int rnd_train_fma(const QLayout& L, const float* pk, const float* x, float* loss_out, float* grads, float* novelty,
                  long B, void* ws, size_t ws_bytes, cudaStream_t st);

size_t rnd_tc_workspace(const QLayout& L, long B);
int rnd_novelty_tc(const QLayout& L, const void* packed, const float* x, float* novelty, float* pred, float* target,
                   long B, void* ws, size_t ws_bytes, cudaStream_t st);
int rnd_train_tc(const QLayout& L, const void* packed, const float* x, float* loss_out, float* grads, float* novelty,
                 long B, void* ws, size_t ws_bytes, cudaStream_t st);

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace ddp

using namespace ddp;

extern "C" {

int ddp_abi_version(void) { return DDP_ABI_VERSION; }
const char* ddp_last_error(void) { return g_err; }

size_t ddp_actor_packed_bytes(const ddp_actor_shape* s, int precision) {
    if (check_actor_shape(s) != DDP_OK) return 0;
    return make_actor_layout(*s, precision).total_bytes;
}

int ddp_actor_pack_parts(const ddp_actor_shape* s, const float* const params[12], void* packed, int precision,
                         int parts, void* stream) {
    int rc = check_actor_shape(s);
    if (rc != DDP_OK) return rc;
    if (!params || !packed) DDP_FAIL(DDP_ERR_ARG, "ddp_actor_pack: NULL argument");
    for (int i = 0; i < 12; ++i)
        if (!params[i]) DDP_FAIL(DDP_ERR_ARG, "ddp_actor_pack: params[%d] is NULL", i);
    if (!aligned16(packed)) DDP_FAIL(DDP_ERR_ARG, "ddp_actor_pack: packed buffer must be 16-byte aligned");
    if (precision != DDP_FP32 && precision != DDP_BF16) DDP_FAIL(DDP_ERR_ARG, "unknown precision %d", precision);
    if (!(parts & (DDP_PACK_SAMPLE | DDP_PACK_TRAIN))) DDP_FAIL(DDP_ERR_ARG, "ddp_actor_pack_parts: empty part mask %d", parts);
    ActorLayout L = make_actor_layout(*s, precision);
    cudaStream_t st = (cudaStream_t)stream;
    // biases, scheduler constants and the time table serve every consumer; the transposed fp32 weights only the
    // warp-FMA kernels
    rc = pack_actor_fp32(L, params, (float*)packed, precision == DDP_FP32, st);
    if (rc != DDP_OK) return rc;
    if (precision == DDP_BF16) {
        rc = pack_actor_tc(L, params, packed, (parts & DDP_PACK_SAMPLE) != 0, (parts & DDP_PACK_TRAIN) != 0, st);
        if (rc != DDP_OK) return rc;
        if (parts & DDP_PACK_TRAIN) return pack_actor_train_tc(L, params, packed, st);
    }
    return DDP_OK;
}

int ddp_actor_pack(const ddp_actor_shape* s, const float* const params[12], void* packed, int precision,
                   void* stream) {
    return ddp_actor_pack_parts(s, params, packed, precision, DDP_PACK_SAMPLE | DDP_PACK_TRAIN, stream);
}

size_t ddp_actor_sample_workspace_bytes(const ddp_actor_shape* s, long B, int precision) {
    if (check_actor_shape(s) != DDP_OK) return 0;
    if (precision != DDP_BF16) return 0;
    return actor_sample_tc_workspace(make_actor_layout(*s, precision), B);
}

int ddp_actor_sample_noisy(const ddp_actor_shape* s, const void* packed, const float* state, const float* noise,
                           const float* expl_noise, float std_min, float std_max, float noise_bound,
                           float* action_out, long B, int precision, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_actor_shape(s);
    if (rc != DDP_OK) return rc;
    if (B < 0) DDP_FAIL(DDP_ERR_SHAPE, "ddp_actor_sample: negative batch");
    if (B == 0) return DDP_OK;
    if (!packed || !state || !noise || !action_out) DDP_FAIL(DDP_ERR_ARG, "ddp_actor_sample: NULL argument");
    if (expl_noise && (std_min < 0.f || std_max < 0.f)) DDP_FAIL(DDP_ERR_ARG, "ddp_actor_sample_noisy: negative std");
    ActorLayout L = make_actor_layout(*s, precision);
    cudaStream_t st = (cudaStream_t)stream;
    ExplNoise e{expl_noise, std_min, std_max, noise_bound};
    if (precision == DDP_FP32) return actor_sample_fma(L, (const float*)packed, state, noise, action_out, B, e, st);
    if (precision == DDP_BF16) return actor_sample_tc(L, packed, state, noise, action_out, B, e, ws, ws_bytes, st);
    DDP_FAIL(DDP_ERR_ARG, "unknown precision %d", precision);
}

int ddp_actor_sample(const ddp_actor_shape* s, const void* packed, const float* state, const float* noise,
                     float* action_out, long B, int precision, void* ws, size_t ws_bytes, void* stream) {
    return ddp_actor_sample_noisy(s, packed, state, noise, nullptr, 0.f, 0.f, 0.f, action_out, B, precision, ws,
                                  ws_bytes, stream);
}

size_t ddp_actor_grad_count(const ddp_actor_shape* s) {
    if (check_actor_shape(s) != DDP_OK) return 0;
    return actor_grad_offsets(*s).off[12];
}

size_t ddp_actor_train_workspace_bytes(const ddp_actor_shape* s, long B, int precision) {
    if (check_actor_shape(s) != DDP_OK || B <= 0) return 0;
    if (precision == DDP_BF16) return actor_train_tc_workspace(make_actor_layout(*s, precision), B);
    return actor_train_workspace(make_actor_layout(*s, precision), B);
}

int ddp_actor_loss_fwd_bwd_ev(const ddp_actor_shape* s, const void* packed, const float* const params[12],
                              const float* state, const float* action, const float* noise, const int64_t* t,
                              float inv_count, float* loss_out, float* grads_flat, long B, int precision, void* ws,
                              size_t ws_bytes, void* stream, void* const* group_events) {
    int rc = check_actor_shape(s);
    if (rc != DDP_OK) return rc;
    if (B <= 0) DDP_FAIL(DDP_ERR_SHAPE, "ddp_actor_loss_fwd_bwd: batch must be positive");
    if (!packed || !params || !state || !action || !noise || !t || !loss_out || !grads_flat || !ws)
        DDP_FAIL(DDP_ERR_ARG, "ddp_actor_loss_fwd_bwd: NULL argument");
    if (precision != DDP_FP32 && precision != DDP_BF16) DDP_FAIL(DDP_ERR_ARG, "unknown precision %d", precision);
    ActorLayout L = make_actor_layout(*s, precision);
    if (!aligned16(ws) || !aligned16(grads_flat)) DDP_FAIL(DDP_ERR_ARG, "workspace/grads must be 16-byte aligned");
    cudaEvent_t ev[DDP_ACTOR_GRAD_GROUPS] = {};
    if (group_events)
        for (int g = 0; g < DDP_ACTOR_GRAD_GROUPS; ++g) ev[g] = (cudaEvent_t)group_events[g];
    if (precision == DDP_BF16)
        return actor_train_tc(L, packed, params, state, action, noise, t, inv_count, loss_out, grads_flat, B, ws,
                              ws_bytes, (cudaStream_t)stream, ev);
    if (ws_bytes < actor_train_workspace(L, B)) DDP_FAIL(DDP_ERR_ARG, "ddp_actor_loss_fwd_bwd: workspace too small");
    rc = actor_train_fma(L, (const float*)packed, params, state, action, noise, t, inv_count, loss_out, grads_flat,
                         B, ws, ws_bytes, (cudaStream_t)stream);
    // the fp32 path finishes all groups together: every event marks the end of the call
    for (int g = 0; rc == DDP_OK && g < DDP_ACTOR_GRAD_GROUPS; ++g)
        if (ev[g]) DDP_CUDA_CHECK(cudaEventRecord(ev[g], (cudaStream_t)stream));
    return rc;
}

int ddp_actor_loss_fwd_bwd(const ddp_actor_shape* s, const void* packed, const float* const params[12],
                           const float* state, const float* action, const float* noise, const int64_t* t,
                           float inv_count, float* loss_out, float* grads_flat, long B, int precision, void* ws,
                           size_t ws_bytes, void* stream) {
    return ddp_actor_loss_fwd_bwd_ev(s, packed, params, state, action, noise, t, inv_count, loss_out, grads_flat, B,
                                     precision, ws, ws_bytes, stream, nullptr);
}

int ddp_clip_adamw_step(float* params_flat, float* grads_flat, float* exp_avg, float* exp_avg_sq, size_t n, int step,
                        float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                        float* norm_out, float* scratch, void* stream) {
    if (!params_flat || !grads_flat || !exp_avg || !exp_avg_sq || !norm_out || !scratch)
        DDP_FAIL(DDP_ERR_ARG, "ddp_clip_adamw_step: NULL argument");
    if (step < 1) DDP_FAIL(DDP_ERR_ARG, "ddp_clip_adamw_step: step is 1-based");
    if (n == 0) return DDP_OK;
    return clip_adamw(params_flat, grads_flat, exp_avg, exp_avg_sq, n, step, lr, beta1, beta2, eps, weight_decay,
                      max_norm, norm_out, scratch, (cudaStream_t)stream);
}

int ddp_clip_adamw_step_dev(float* params_flat, float* grads_flat, float* exp_avg, float* exp_avg_sq, size_t n,
                            int* step_counter, float lr, float beta1, float beta2, float eps, float weight_decay,
                            float max_norm, float* norm_out, float* scratch, void* stream) {
    if (!params_flat || !grads_flat || !exp_avg || !exp_avg_sq || !norm_out || !scratch || !step_counter)
        DDP_FAIL(DDP_ERR_ARG, "ddp_clip_adamw_step_dev: NULL argument");
    if (n == 0) return DDP_OK;
    return clip_adamw_dev(params_flat, grads_flat, exp_avg, exp_avg_sq, n, step_counter, lr, beta1, beta2, eps,
                          weight_decay, max_norm, norm_out, scratch, (cudaStream_t)stream);
}

size_t ddp_q_packed_bytes(const ddp_q_shape* s, int precision) {
    if (check_q_shape(s) != DDP_OK) return 0;
    return make_q_layout(*s, precision).total_bytes;
}

int ddp_q_pack(const ddp_q_shape* s, const float* const params[], void* packed, int precision, void* stream) {
    int rc = check_q_shape(s);
    if (rc != DDP_OK) return rc;
    if (!params || !packed) DDP_FAIL(DDP_ERR_ARG, "ddp_q_pack: NULL argument");
    for (int i = 0; i < 16 * s->n_modes; ++i)
        if (!params[i]) DDP_FAIL(DDP_ERR_ARG, "ddp_q_pack: params[%d] is NULL", i);
    if (precision != DDP_FP32 && precision != DDP_BF16) DDP_FAIL(DDP_ERR_ARG, "unknown precision %d", precision);
    QLayout L = make_q_layout(*s, precision);
    rc = pack_q_fp32(L, params, (float*)packed, (cudaStream_t)stream, precision == DDP_BF16);
    if (rc != DDP_OK || precision == DDP_FP32) return rc;
    return pack_q_tc(L, params, packed, (cudaStream_t)stream);
}

static int check_segments(const ddp_q_shape* s, const int64_t* seg_off, long B) {
    if (!seg_off) DDP_FAIL(DDP_ERR_ARG, "seg_off is NULL");
    if (seg_off[0] != 0 || seg_off[s->n_modes] != B) DDP_FAIL(DDP_ERR_SHAPE, "seg_off must start at 0 and end at B");
    for (int m = 0; m < s->n_modes; ++m)
        if (seg_off[m + 1] < seg_off[m]) DDP_FAIL(DDP_ERR_SHAPE, "seg_off must be non-decreasing");
    return DDP_OK;
}

size_t ddp_q_forward_workspace_bytes(const ddp_q_shape* s, long B, int precision) {
    if (check_q_shape(s) != DDP_OK || B <= 0 || precision != DDP_BF16) return 0;
    return q_tc_workspace(make_q_layout(*s, precision), B, 0);
}

int ddp_q_forward(const ddp_q_shape* s, const void* packed, const int64_t* seg_off, const float* obs,
                  const float* act, float* q_min_out, float* p1_out, float* p2_out, float* dq_da_out, long B,
                  int precision, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_q_shape(s);
    if (rc != DDP_OK) return rc;
    if (B < 0) DDP_FAIL(DDP_ERR_SHAPE, "ddp_q_forward: negative batch");
    if (B == 0) return DDP_OK;
    if (!packed || !obs || !act) DDP_FAIL(DDP_ERR_ARG, "ddp_q_forward: NULL argument");
    if (precision != DDP_FP32 && precision != DDP_BF16) DDP_FAIL(DDP_ERR_ARG, "unknown precision %d", precision);
    rc = check_segments(s, seg_off, B);
    if (rc != DDP_OK) return rc;
    if (precision == DDP_BF16)
        return q_forward_tc(make_q_layout(*s, precision), packed, seg_off, obs, act, q_min_out, p1_out, p2_out,
                            dq_da_out, B, ws, ws_bytes, (cudaStream_t)stream);
    return q_forward_fma(make_q_layout(*s, precision), (const float*)packed, seg_off, obs, act, q_min_out, p1_out,
                         p2_out, dq_da_out, B, (cudaStream_t)stream);
}

size_t ddp_q_ascent_workspace_bytes(const ddp_q_shape* s, long B, int iters) {
    if (check_q_shape(s) != DDP_OK || B <= 0 || iters <= 0) return 0;
    const size_t a = q_ascent_workspace(make_q_layout(*s, DDP_FP32), B, iters);
    const size_t b = q_tc_workspace(make_q_layout(*s, DDP_BF16), B, iters);
    return a > b ? a : b;
}

int ddp_q_action_ascent_sharded(const ddp_q_shape* s, const void* packed, const int64_t* seg_off,
                                const int64_t* seg_mean_count, const float* obs, float* action_inout, int iters, float lr,
                                float beta1, float beta2, float eps, float max_norm, float lim, float* mean_abs_out,
                                float* gnorm_out, long B, int precision, void* ws, size_t ws_bytes, void* stream,
                                ddp_gsq_reduce_fn reduce, void* reduce_user) {
    int rc = check_q_shape(s);
    if (rc != DDP_OK) return rc;
    if (B <= 0 || iters <= 0) DDP_FAIL(DDP_ERR_SHAPE, "ddp_q_action_ascent: B and iters must be positive");
    if (!packed || !obs || !action_inout || !mean_abs_out || !ws || !seg_mean_count)
        DDP_FAIL(DDP_ERR_ARG, "ddp_q_action_ascent: NULL argument");
    if (precision != DDP_FP32 && precision != DDP_BF16) DDP_FAIL(DDP_ERR_ARG, "unknown precision %d", precision);
    rc = check_segments(s, seg_off, B);
    if (rc != DDP_OK) return rc;
    QLayout L = make_q_layout(*s, precision);
    if (precision == DDP_BF16)
        return q_ascent_tc(L, packed, seg_off, seg_mean_count, obs, action_inout, iters, lr, beta1, beta2, eps, max_norm,
                           lim, mean_abs_out, gnorm_out, B, ws, ws_bytes, (cudaStream_t)stream, reduce, reduce_user);
    if (ws_bytes < q_ascent_workspace(L, B, iters)) DDP_FAIL(DDP_ERR_ARG, "ddp_q_action_ascent: workspace too small");
    return q_ascent_fma(L, (const float*)packed, seg_off, seg_mean_count, obs, action_inout, iters, lr, beta1, beta2,
                        eps, max_norm, lim, mean_abs_out, gnorm_out, B, ws, ws_bytes, (cudaStream_t)stream, reduce,
                        reduce_user);
}

int ddp_q_action_ascent(const ddp_q_shape* s, const void* packed, const int64_t* seg_off,
                        const int64_t* seg_mean_count, const float* obs, float* action_inout, int iters, float lr,
                        float beta1, float beta2, float eps, float max_norm, float lim, float* mean_abs_out,
                        float* gnorm_out, long B, int precision, void* ws, size_t ws_bytes, void* stream) {
    return ddp_q_action_ascent_sharded(s, packed, seg_off, seg_mean_count, obs, action_inout, iters, lr, beta1, beta2, eps,
                                       max_norm, lim, mean_abs_out, gnorm_out, B, precision, ws, ws_bytes, stream, nullptr,
                                       nullptr);
}

size_t ddp_q_grad_count(const ddp_q_shape* s) {
    if (check_q_shape(s) != DDP_OK) return 0;
    return q_grad_count(make_q_layout(*s, DDP_FP32));
}

size_t ddp_q_critic_train_workspace_bytes(const ddp_q_shape* s, long B, int precision) {
    if (check_q_shape(s) != DDP_OK || B <= 0 || (precision != DDP_FP32 && precision != DDP_BF16)) return 0;
    if (precision == DDP_BF16) return q_critic_train_tc_workspace(make_q_layout(*s, precision), B);
    return q_critic_train_workspace(make_q_layout(*s, precision), B);
}

int ddp_q_critic_loss_fwd_bwd(const ddp_q_shape* s, const void* packed, const void* packed_target, const float* obs,
                              const float* action, const float* next_obs, const float* next_action,
                              const float* reward, const float* done, float gamma_n, float* loss_out,
                              float* grads_flat, long B, int precision, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_q_shape(s);
    if (rc != DDP_OK) return rc;
    if (s->n_modes != 1) DDP_FAIL(DDP_ERR_SHAPE, "ddp_q_critic_loss_fwd_bwd: one critic per call (n_modes must be 1)");
    if (B <= 0) DDP_FAIL(DDP_ERR_SHAPE, "ddp_q_critic_loss_fwd_bwd: batch must be positive");
    if (!packed || !packed_target || !obs || !action || !next_obs || !next_action || !reward || !done || !loss_out ||
        !grads_flat || !ws)
        DDP_FAIL(DDP_ERR_ARG, "ddp_q_critic_loss_fwd_bwd: NULL argument");
    if (precision != DDP_FP32 && precision != DDP_BF16) DDP_FAIL(DDP_ERR_ARG, "critic update: unknown precision %d", precision);
    if (!aligned16(ws) || !aligned16(grads_flat)) DDP_FAIL(DDP_ERR_ARG, "workspace/grads must be 16-byte aligned");
    if (precision == DDP_BF16)      // both packs must have been made with DDP_BF16 (they carry the 16-bit operand section)
        return q_critic_train_tc(make_q_layout(*s, precision), packed, packed_target, obs, action, next_obs, next_action,
                                 reward, done, gamma_n, loss_out, grads_flat, B, ws, ws_bytes, (cudaStream_t)stream);
    return q_critic_train_fma(make_q_layout(*s, precision), (const float*)packed, (const float*)packed_target, obs,
                              action, next_obs, next_action, reward, done, gamma_n, loss_out, grads_flat, B, ws,
                              ws_bytes, (cudaStream_t)stream);
}

// RNDModel as a two-net pack: O = D, no action columns, atoms = F, net 0 = predictor, net 1 = target
static int rnd_layout(const ddp_rnd_shape* s, QLayout* L, int precision = DDP_FP32) {
    if (!s) DDP_FAIL(DDP_ERR_ARG, "RND shape is NULL");
    if (precision != DDP_FP32 && precision != DDP_BF16) DDP_FAIL(DDP_ERR_ARG, "unknown precision %d", precision);
    if (s->D <= 0 || s->D > 512 || s->F < 4 || s->F > 256 || s->F % 4)
        DDP_FAIL(DDP_ERR_SHAPE, "RND shape: need 0<D<=512 and F a multiple of 4 in [4,256] (got D=%d F=%d)", s->D, s->F);
    if (s->hid1 <= 0 || s->hid2 <= 0 || s->hid3 <= 0 || s->hid1 % 4 || s->hid2 % 4 || s->hid3 % 4 || s->hid1 > 1024 ||
        s->hid2 > 1024 || s->hid3 > 1024)
        DDP_FAIL(DDP_ERR_SHAPE, "RND shape: hidden widths must be multiples of 4 in (0,1024]");
    if (precision == DDP_BF16 && (s->D > 256 || s->hid1 % 64 || s->hid2 % 64 || s->hid3 % 64))
        DDP_FAIL(DDP_ERR_UNSUPPORTED, "RND tensor path: need D<=256 and hidden widths multiples of 64");
    ddp_q_shape q{};
    q.O = s->D; q.A = 0; q.atoms = s->F; q.v_min = 0.f; q.v_max = 1.f; q.n_modes = 1;
    q.hid1 = s->hid1; q.hid2 = s->hid2; q.hid3 = s->hid3;
    *L = make_q_layout(q, precision);
    return DDP_OK;
}

size_t ddp_rnd_packed_bytes_p(const ddp_rnd_shape* s, int precision) {
    QLayout L;
    return rnd_layout(s, &L, precision) == DDP_OK ? L.total_bytes : 0;
}
size_t ddp_rnd_packed_bytes(const ddp_rnd_shape* s) { return ddp_rnd_packed_bytes_p(s, DDP_FP32); }

int ddp_rnd_pack_p(const ddp_rnd_shape* s, const float* const params[16], void* packed, int precision, void* stream) {
    QLayout L;
    int rc = rnd_layout(s, &L, precision);
    if (rc != DDP_OK) return rc;
    if (!params || !packed) DDP_FAIL(DDP_ERR_ARG, "ddp_rnd_pack: NULL argument");
    for (int i = 0; i < 16; ++i)
        if (!params[i]) DDP_FAIL(DDP_ERR_ARG, "ddp_rnd_pack: params[%d] is NULL", i);
    rc = pack_q_fp32(L, params, (float*)packed, (cudaStream_t)stream, precision == DDP_BF16);
    if (rc != DDP_OK || precision == DDP_FP32) return rc;
    return pack_q_tc(L, params, packed, (cudaStream_t)stream);
}
int ddp_rnd_pack(const ddp_rnd_shape* s, const float* const params[16], void* packed, void* stream) {
    return ddp_rnd_pack_p(s, params, packed, DDP_FP32, stream);
}

size_t ddp_rnd_workspace_bytes_p(const ddp_rnd_shape* s, long B, int precision) {
    QLayout L;
    if (rnd_layout(s, &L, precision) != DDP_OK || B <= 0) return 0;
    return precision == DDP_BF16 ? rnd_tc_workspace(L, B) : rnd_train_workspace(L, B);
}

int ddp_rnd_novelty_p(const ddp_rnd_shape* s, const void* packed, const float* x, float* novelty_out, float* pred_out,
                      float* target_out, long B, int precision, void* ws, size_t ws_bytes, void* stream) {
    QLayout L;
    int rc = rnd_layout(s, &L, precision);
    if (rc != DDP_OK) return rc;
    if (B <= 0) DDP_FAIL(DDP_ERR_SHAPE, "ddp_rnd_novelty: batch must be positive");
    if (!packed || !x) DDP_FAIL(DDP_ERR_ARG, "ddp_rnd_novelty: NULL argument");
    if (precision == DDP_BF16) {
        if (!ws || !aligned16(ws)) DDP_FAIL(DDP_ERR_ARG, "ddp_rnd_novelty: the tensor path needs a 16-byte aligned workspace");
        return rnd_novelty_tc(L, packed, x, novelty_out, pred_out, target_out, B, ws, ws_bytes, (cudaStream_t)stream);
    }
    return rnd_novelty_fma(L, (const float*)packed, x, novelty_out, pred_out, target_out, B, (cudaStream_t)stream);
}
int ddp_rnd_novelty(const ddp_rnd_shape* s, const void* packed, const float* x, float* novelty_out, float* pred_out,
                    float* target_out, long B, void* stream) {
    return ddp_rnd_novelty_p(s, packed, x, novelty_out, pred_out, target_out, B, DDP_FP32, nullptr, 0, stream);
}

size_t ddp_rnd_grad_count(const ddp_rnd_shape* s) {
    QLayout L;
    return rnd_layout(s, &L) == DDP_OK ? rnd_grad_count(L) : 0;
}

size_t ddp_rnd_train_workspace_bytes(const ddp_rnd_shape* s, long B) { return ddp_rnd_workspace_bytes_p(s, B, DDP_FP32); }

int ddp_rnd_loss_fwd_bwd_p(const ddp_rnd_shape* s, const void* packed, const float* x, float* loss_out, float* grads_flat,
                           float* novelty_out, long B, int precision, void* ws, size_t ws_bytes, void* stream) {
    QLayout L;
    int rc = rnd_layout(s, &L, precision);
    if (rc != DDP_OK) return rc;
    if (B <= 0) DDP_FAIL(DDP_ERR_SHAPE, "ddp_rnd_loss_fwd_bwd: batch must be positive");
    if (!packed || !x || !loss_out || !grads_flat || !ws) DDP_FAIL(DDP_ERR_ARG, "ddp_rnd_loss_fwd_bwd: NULL argument");
    if (!aligned16(ws) || !aligned16(grads_flat)) DDP_FAIL(DDP_ERR_ARG, "workspace/grads must be 16-byte aligned");
    if (precision == DDP_BF16)
        return rnd_train_tc(L, packed, x, loss_out, grads_flat, novelty_out, B, ws, ws_bytes, (cudaStream_t)stream);
    return rnd_train_fma(L, (const float*)packed, x, loss_out, grads_flat, novelty_out, B, ws, ws_bytes,
                         (cudaStream_t)stream);
}
int ddp_rnd_loss_fwd_bwd(const ddp_rnd_shape* s, const void* packed, const float* x, float* loss_out, float* grads_flat,
                         float* novelty_out, long B, void* ws, size_t ws_bytes, void* stream) {
    return ddp_rnd_loss_fwd_bwd_p(s, packed, x, loss_out, grads_flat, novelty_out, B, DDP_FP32, ws, ws_bytes, stream);
}

}  // extern "C"
