/*
 * ddiffpg_b200 -- C ABI of the B200-native DDiffPG hot path (sm_100a).
 *
 * The reference (sayantanauddy/ddiffpg) is pure Python/PyTorch and has no FFI; the seam it offers is
 * the Python object protocol of two nn.Modules and one agent method.  Each
 * entry point below names the reference interface it replaces (paths relative to the reference repo).
 * The Python host layer (ddiffpg_b200/models.py, ddiffpg_b200/algo.py) binds these with ctypes and
 * mirrors the reference call surface one to one; INTEGRATION.md shows the stub a maintainer adds.
 *
 * Conventions
 *  - every data pointer is a DEVICE pointer on the current device, fp32 row-major contiguous unless
 *    stated; `*_shape` structs and `params[]` pointer arrays are HOST memory;
 *  - the caller owns all memory (inputs, outputs, packed weights, workspace); functions only enqueue
 *    work on `stream` (a cudaStream_t passed as void*), never allocate and never synchronise;
 *  - return 0 on success, a negative ddp_status otherwise; ddp_last_error() gives the thread-local
 *    message of the last failure; no C++ types or exceptions cross the boundary;
 *  - `precision`: DDP_FP32 = warp-level FMA path (parity <= 1e-4 relative against the reference),
 *    DDP_BF16 = tcgen05/TMEM tensor-core path (bf16 operands, fp32 accumulate, parity <= 1e-2).
 */
#ifndef DDIFFPG_B200_H
#define DDIFFPG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDP_ABI_VERSION 1

enum ddp_status {
    DDP_OK = 0,
    DDP_ERR_SHAPE = -1,        /* unsupported or inconsistent shape                     */
    DDP_ERR_ARG = -2,          /* null pointer, misaligned buffer, workspace too small  */
    DDP_ERR_UNSUPPORTED = -3,  /* e.g. precision not available for this shape           */
    DDP_ERR_CUDA = -4          /* a CUDA runtime call failed (message in last_error)    */
};

enum ddp_precision { DDP_FP32 = 0, DDP_BF16 = 1 };

/* Diffusion actor: DiffusionPolicy(state_dim=S, action_dim=A, diffusion_iter=T) with
 * DiffusionNet(dim=D) and trunk widths h1,h2,h3 (ddiffpg/models/diffusion_mlp.py:24-58,148-173).
 * Reference values: S=34, A=8, T=5, D=256, h1=1024, h2=512, h3=256. */
typedef struct { int S, A, T, D, h1, h2, h3; } ddp_actor_shape;

/* DistributionalDoubleQ(state_dim=O, act_dim=A, v_min, v_max, num_atoms) with MLPNet hidden layers
 * hid1,hid2,hid3 (ddiffpg/models/mlp.py:23-35,131-141); n_modes critics, one per behaviour mode
 * (ddiffpg/utils/Q_scheduler.py:16-45).  Reference: O=29, A=8, atoms=51, v in [0,5], 512/256/128. */
typedef struct { int O, A, atoms; float v_min, v_max; int n_modes; int hid1, hid2, hid3; } ddp_q_shape;

int ddp_abi_version(void);
const char* ddp_last_error(void);

/* ----------------------------------------------------------------------------------------------
 * Actor weights.  params[12] are the tensors of DiffusionPolicy.state_dict() in its own order
 * (net.time_mlp.{1,3}.{weight,bias}, net.mlp.{0,2,4,6}.{weight,bias}).  Packing (a) transposes the
 * trunk weights for coalesced streaming, (b) folds SinusoidalPosEmb + time_mlp + the 256 time
 * columns of net.mlp.0 + its bias into a [T, h1] table (diffusion_mlp.py:14-21,38-43,68-70: these
 * depend on t only), (c) evaluates the DDPMScheduler constants (diffusers, call sites
 * diffusion_mlp.py:167-173,243-247,309-310) and (d) for DDP_BF16 adds bf16 tensor-core tiles.
 * Re-pack whenever a parameter changes. */
size_t ddp_actor_packed_bytes(const ddp_actor_shape* shape, int precision);
int ddp_actor_pack(const ddp_actor_shape* shape, const float* const params[12], void* packed,
                   int precision, void* stream);
/* The same, restricted to what one consumer reads (a training loop re-packs after every optimizer step and never
 * samples in between): parts = DDP_PACK_SAMPLE (ddp_actor_sample*), DDP_PACK_TRAIN (ddp_actor_loss_fwd_bwd) or both.
 * A buffer packed for a precision serves calls of that precision only. */
enum ddp_pack_parts { DDP_PACK_SAMPLE = 1, DDP_PACK_TRAIN = 2 };
int ddp_actor_pack_parts(const ddp_actor_shape* shape, const float* const params[12], void* packed,
                         int precision, int parts, void* stream);

/* Replaces DiffusionPolicy.forward / get_actions(sample=True, add_noise=False)
 * (ddiffpg/models/diffusion_mlp.py:184-185,219-251): the whole T-step reverse chain in one launch.
 * state [B,S]; noise [T,B,A] with noise[0] = x_T (the draw at :222) and noise[j] = the Gaussian the
 * scheduler adds at t = T-j (j >= 1); action_out [B,A] in [-1,1].
 * DDP_BF16 needs a 16-byte aligned device workspace of ddp_actor_sample_workspace_bytes() (per-CTA scratch for the
 * state part of the first layer, 256 KB per SM at width 1024; queried with the target device current); DDP_FP32
 * needs none (0 bytes, ws may be NULL).  Like every workspace of this API it belongs to one call at a time: launches
 * that may overlap (different streams) need their own. */
size_t ddp_actor_sample_workspace_bytes(const ddp_actor_shape* shape, long B, int precision);
int ddp_actor_sample(const ddp_actor_shape* shape, const void* packed, const float* state,
                     const float* noise, float* action_out, long B, int precision,
                     void* ws, size_t ws_bytes, void* stream);

/* ddp_actor_sample fused with the noise its callers add right after it: AgentDDiffPG.get_actions
 * (add_mixed_normal_noise / add_normal_noise, ddiffpg/algo/ddiffpg.py:88-99) and get_tgt_policy_actions
 * (:102-109), i.e. ddiffpg/utils/noise.py:19-41:
 *   out = clamp(a + clamp(std_r * z, -noise_bound, +noise_bound), -1, 1), std_r = linspace(std_min, std_max, B)[r]
 * expl_noise [B,A] is pre-drawn standard normal noise (NULL = plain ddp_actor_sample); std_min == std_max is
 * the 'fixed' type; noise_bound <= 0 disables the inner clamp. */
int ddp_actor_sample_noisy(const ddp_actor_shape* shape, const void* packed, const float* state,
                           const float* noise, const float* expl_noise, float std_min, float std_max,
                           float noise_bound, float* action_out, long B, int precision,
                           void* ws, size_t ws_bytes, void* stream);

/* Replaces DiffusionPolicy.get_loss + the backward of optimizer_update
 * (ddiffpg/models/diffusion_mlp.py:294-321, ddiffpg/algo/ac_base.py:83-85).
 * state [B,S], action [B,A], noise [B,A], t [B] int64 in [0,T).  inv_count = 1/(B_global*A).
 * loss_out[0]      += sum((eps_hat-noise)^2) * inv_count    (zero it first; partial sums add up)
 * grads_flat[...]   = d loss / d params, flat in state_dict order (overwritten, not accumulated).
 * params[12] are the live fp32 parameters (the backward streams them in their own layout). */
size_t ddp_actor_grad_count(const ddp_actor_shape* shape);
size_t ddp_actor_train_workspace_bytes(const ddp_actor_shape* shape, long B, int precision);
int ddp_actor_loss_fwd_bwd(const ddp_actor_shape* shape, const void* packed,
                           const float* const params[12], const float* state, const float* action,
                           const float* noise, const int64_t* t, float inv_count, float* loss_out,
                           float* grads_flat, long B, int precision, void* ws, size_t ws_bytes,
                           void* stream);
/* The same call with completion events for the data-parallel exchange step (the sum all-reduce of the
 * flat gradient is the only collective of the path).  The backward produces the gradient in four groups, last layer
 * first; group_events[g] (a cudaEvent_t, or NULL to skip) is recorded on `stream` as soon as group g is final, so the
 * caller can start reducing it on another stream while the rest of the backward still runs:
 *   g = 0: net.mlp.6.{weight,bias} (and loss_out)   1: net.mlp.4.*   2: net.mlp.2.*
 *   g = 3: net.time_mlp.* and net.mlp.0.* (recorded at the end of the call).
 * group_events == NULL is ddp_actor_loss_fwd_bwd. */
enum { DDP_ACTOR_GRAD_GROUPS = 4 };
int ddp_actor_loss_fwd_bwd_ev(const ddp_actor_shape* shape, const void* packed,
                              const float* const params[12], const float* state, const float* action,
                              const float* noise, const int64_t* t, float inv_count, float* loss_out,
                              float* grads_flat, long B, int precision, void* ws, size_t ws_bytes,
                              void* stream, void* const* group_events);

/* Tail of ActorCriticBase.optimizer_update (ddiffpg/algo/ac_base.py:86-91) for a flat parameter
 * vector: g_norm = ||g||_2, g *= min(1, max_norm/(g_norm+1e-6)), AdamW step (torch.optim.AdamW
 * semantics, decoupled weight decay, bias-corrected).  norm_out[0] receives the pre-clip norm.
 * step is the 1-based step count after this update.  scratch: DDP_ADAMW_SCRATCH_FLOATS floats (per-block partial
 * sums of ||g||^2, added up in a fixed order: the norm, hence the clip coefficient and the parameters, are
 * bit-identical on data-parallel replicas that hold the same reduced gradient; no atomics). */
enum { DDP_ADAMW_SCRATCH_FLOATS = 640 };
int ddp_clip_adamw_step(float* params_flat, float* grads_flat, float* exp_avg, float* exp_avg_sq,
                        size_t n, int step, float lr, float beta1, float beta2, float eps,
                        float weight_decay, float max_norm, float* norm_out, float* scratch,
                        void* stream);

/* The same step with the 1-based step count kept in DEVICE memory: *step_counter is incremented by the call and the
 * bias corrections are derived from it on the device, so the launch sequence carries no per-step host value and
 * a whole training step can be captured once into a CUDA graph.  scratch: DDP_ADAMW_SCRATCH_FLOATS floats. */
int ddp_clip_adamw_step_dev(float* params_flat, float* grads_flat, float* exp_avg, float* exp_avg_sq, size_t n,
                            int* step_counter, float lr, float beta1, float beta2, float eps, float weight_decay,
                            float max_norm, float* norm_out, float* scratch, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Critics.  params[16*n_modes]: per mode the tensors of DistributionalDoubleQ.state_dict()
 * (net_q1.net.{0,2,4,6}.{weight,bias}, net_q2.net....). */
size_t ddp_q_packed_bytes(const ddp_q_shape* shape, int precision);
int ddp_q_pack(const ddp_q_shape* shape, const float* const params[], void* packed, int precision,
               void* stream);

/* Replaces DistributionalDoubleQ.get_q1_q2 / get_q_min (ddiffpg/models/mlp.py:143-151).
 * Rows are sorted by mode; seg_off[n_modes+1] (HOST) gives the row range of each mode's critic.
 * q_min_out [B]; p1_out/p2_out [B,atoms] may be NULL; dq_da_out [B,A] (d q_min / d action, the
 * autograd result of get_q_min w.r.t. action) may be NULL.  ws: ddp_q_forward_workspace_bytes (0 for
 * DDP_FP32, then ws may be NULL). */
size_t ddp_q_forward_workspace_bytes(const ddp_q_shape* shape, long B, int precision);
int ddp_q_forward(const ddp_q_shape* shape, const void* packed, const int64_t* seg_off,
                  const float* obs, const float* act, float* q_min_out, float* p1_out, float* p2_out,
                  float* dq_da_out, long B, int precision, void* ws, size_t ws_bytes, void* stream);

/* Replaces AgentDDiffPG.update_target_action (ddiffpg/algo/ddiffpg.py:358-373, identical copy
 * ddiffpg/algo/dipo.py:246-261) including its optimizer_update tail (ac_base.py:83-92): pre-clamp,
 * then `iters` x { g = grad_a(-mean_seg min(Q1,Q2)); global-L2 clip over the segment to max_norm;
 * Adam(lr, betas, eps, fresh state); clamp(+-lim) }.  Each mode segment is one reference call: its
 * mean uses seg_mean_count[m] rows (pass the segment length for reference semantics, or the global
 * mode batch when sharded) and its own clip norm.  action_inout [B,A] is updated in place;
 * mean_abs_out [n_modes] receives mean|a| per segment; gnorm_out [n_modes*iters] the pre-clip norms
 * (may be NULL).  ws: ddp_q_ascent_workspace_bytes. */
size_t ddp_q_ascent_workspace_bytes(const ddp_q_shape* shape, long B, int iters);   /* covers both precisions */
int ddp_q_action_ascent(const ddp_q_shape* shape, const void* packed, const int64_t* seg_off,
                        const int64_t* seg_mean_count, const float* obs, float* action_inout,
                        int iters, float lr, float beta1, float beta2, float eps, float max_norm,
                        float lim, float* mean_abs_out, float* gnorm_out, long B, int precision,
                        void* ws, size_t ws_bytes, void* stream);

/* The same loop on a ROW-SHARDED batch with global-batch semantics (SURVEY.md 8e, H2 semantics (ii); the reference
 * itself is single-process): every rank holds a slice of each mode segment, passes the GLOBAL mode batch as
 * seg_mean_count[m], and `reduce` is the one exchange step of an iteration -- called between the gradient pass and the
 * Adam step with the device vector gsq_dev[n_modes] = sum of g^2 over THIS rank's rows of each mode; it must replace it by
 * the sum over all ranks, in stream order on `stream` (ncclAllReduce, torch.distributed.all_reduce, ...), and return 0.
 * The clip coefficient of ac_base.py:87-88 then comes from the norm over the whole mode batch, so the sharded ranks take
 * exactly the steps one process would take on the gathered batch.  gnorm_out receives the global norms; mean_abs_out
 * stays the mean over this rank's rows.  reduce == NULL is ddp_q_action_ascent (shard-local norm). */
typedef int (*ddp_gsq_reduce_fn)(float* gsq_dev, int n_modes, void* stream, void* user);
int ddp_q_action_ascent_sharded(const ddp_q_shape* shape, const void* packed, const int64_t* seg_off,
                                const int64_t* seg_mean_count, const float* obs, float* action_inout,
                                int iters, float lr, float beta1, float beta2, float eps, float max_norm,
                                float lim, float* mean_abs_out, float* gnorm_out, long B, int precision,
                                void* ws, size_t ws_bytes, void* stream, ddp_gsq_reduce_fn reduce, void* reduce_user);

/* ----------------------------------------------------------------------------------------------
 * Critic update.  Replaces the loss/backward half of AgentDDiffPG.update_critic
 * (ddiffpg/algo/ddiffpg.py:322-351) for ONE critic (shape->n_modes must be 1):
 *   target  = min(projection(Q1_tgt), projection(Q2_tgt))   critic_target.get_q1_q2(next_obs, next_action) and the
 *             C51 projection of reward + (1-done)*gamma_n*z (ddiffpg/utils/distl_util.py:4-20), no gradient;
 *   loss    = BCE(current_Q1, target) + BCE(current_Q2, target)   (F.binary_cross_entropy, mean over B*atoms);
 *   grads   = d loss / d params of `packed`, flat in DistributionalDoubleQ.state_dict() order (overwritten).
 * reward, done: [B] fp32.  loss_out[0] += loss (zero it first).  The optimizer step (clip_grad_norm_ + AdamW,
 * ac_base.py:86-91) stays with the caller: ddp_clip_adamw_step on a flat vector, or torch's own.
 * precision: DDP_FP32 (FMA tile kernel + fp32 dW GEMMs, 1e-4) or DDP_BF16 (target heads by the fused tcgen05 chain,
 * forward / dX / dW by the tcgen05 row and dW GEMMs, softmax + BCE + projection in fp32; 1e-2) -- `packed` and
 * `packed_target` must both have been packed with the same precision. */
size_t ddp_q_grad_count(const ddp_q_shape* shape);
size_t ddp_q_critic_train_workspace_bytes(const ddp_q_shape* shape, long B, int precision);
int ddp_q_critic_loss_fwd_bwd(const ddp_q_shape* shape, const void* packed, const void* packed_target,
                              const float* obs, const float* action, const float* next_obs,
                              const float* next_action, const float* reward, const float* done, float gamma_n,
                              float* loss_out, float* grads_flat, long B, int precision, void* ws, size_t ws_bytes,
                              void* stream);

/* ----------------------------------------------------------------------------------------------
 * RND / NovelD intrinsic reward.  RNDModel(state_dim=D) (ddiffpg/models/mlp.py:233-267):
 * predictor and target, each Linear(D,hid1)-ELU-Linear(hid1,hid2)-ELU-Linear(hid2,hid3)-ELU-Linear(hid3,F).
 * Reference: D = 69 (29 + 2*2*10 positional encoding), 512/256/128, F = 128.  params[16] = RNDModel.state_dict()
 * order (predictor.{0,2,4,6}.{weight,bias}, target....). */
typedef struct { int D, F, hid1, hid2, hid3; } ddp_rnd_shape;
size_t ddp_rnd_packed_bytes(const ddp_rnd_shape* shape);
int ddp_rnd_pack(const ddp_rnd_shape* shape, const float* const params[16], void* packed, void* stream);

/* Replaces RNDModel.forward and IntrinsicM.get_novelty (ddiffpg/models/mlp.py:262-266, ddiffpg/utils/intrinsic.py:62-65):
 * novelty_out[r] = ||predictor(x_r) - target(x_r)||_2; pred_out / target_out [B,F] receive the two feature vectors.
 * Any of the three outputs may be NULL.  x [B,D] is the (already position-encoded) observation. */
int ddp_rnd_novelty(const ddp_rnd_shape* shape, const void* packed, const float* x, float* novelty_out, float* pred_out,
                    float* target_out, long B, void* stream);

/* Replaces the loss/backward half of IntrinsicM.update (ddiffpg/utils/intrinsic.py:67-75):
 * loss_out[0] += mse_loss(predictor(x), target(x)) (zero it first); grads_flat = d loss / d predictor parameters,
 * flat in state_dict order (the target net is frozen), overwritten.  novelty_out may be NULL. */
size_t ddp_rnd_grad_count(const ddp_rnd_shape* shape);
size_t ddp_rnd_train_workspace_bytes(const ddp_rnd_shape* shape, long B);
int ddp_rnd_loss_fwd_bwd(const ddp_rnd_shape* shape, const void* packed, const float* x, float* loss_out,
                         float* grads_flat, float* novelty_out, long B, void* ws, size_t ws_bytes, void* stream);

/* The same three calls with a precision argument (the forms above are DDP_FP32).  DDP_BF16: tcgen05 row / dW GEMMs
 * with fp32 accumulation, novelty / mse / gradients of the loss in fp32 (1e-2); for the update-batch sizes the reference
 * calls IntrinsicM.compute_reward / update with (ddiffpg/algo/ddiffpg.py:225,296-299: 4 096 rows, NovelD 8 192).
 * Needs D <= 256 and hidden widths multiples of 64.  `packed` must come from ddp_rnd_pack_p with the same precision;
 * ws: ddp_rnd_workspace_bytes_p (novelty needs it on the tensor path only; may be NULL with DDP_FP32). */
size_t ddp_rnd_packed_bytes_p(const ddp_rnd_shape* shape, int precision);
int ddp_rnd_pack_p(const ddp_rnd_shape* shape, const float* const params[16], void* packed, int precision, void* stream);
size_t ddp_rnd_workspace_bytes_p(const ddp_rnd_shape* shape, long B, int precision);
int ddp_rnd_novelty_p(const ddp_rnd_shape* shape, const void* packed, const float* x, float* novelty_out,
                      float* pred_out, float* target_out, long B, int precision, void* ws, size_t ws_bytes, void* stream);
int ddp_rnd_loss_fwd_bwd_p(const ddp_rnd_shape* shape, const void* packed, const float* x, float* loss_out,
                           float* grads_flat, float* novelty_out, long B, int precision, void* ws, size_t ws_bytes,
                           void* stream);

/* ----------------------------------------------------------------------------------------------
 * Batch assembly / scatter-back around the hot path, all mode groups in one launch.
 * Replay storage as in DiffusionReplayBuffer (ddiffpg/replay/simple_replay.py:98-200): buf_obs / buf_next_obs [N,O],
 * buf_action [N,A], buf_target_action [n_groups,N,A], buf_reward [N], buf_done [N] (bool bytes).
 * Output row r reads replay row indices[r] for mode group[r] (NULL = group 0):
 *   the six gathers of sample_batch (:150-163; done as float), plus add_embedding (ddiffpg/utils/torch_util.py:17-43)
 *   applied to state and next state: [obs | embeddings[group[r]]] with the embedding zeroed where zero_state[r] /
 *   zero_next[r] is non-zero (the reference draws those rows with np.random.choice; the caller passes the draw).
 * Every output pointer may be NULL (that output is skipped); embeddings == NULL writes zero embeddings.
 * Rows whose index is outside [0,N) produce zeros. */
typedef struct { int O, A, E, n_groups; } ddp_batch_shape;
int ddp_replay_gather(const ddp_batch_shape* shape, const float* buf_obs, const float* buf_action,
                      const float* buf_target_action, const float* buf_reward, const float* buf_next_obs,
                      const uint8_t* buf_done, long N, const int64_t* indices, const int32_t* group,
                      const float* embeddings, const uint8_t* zero_state, const uint8_t* zero_next, float* obs_out,
                      float* action_out, float* target_action_out, float* reward_out, float* next_obs_out,
                      float* done_out, float* state_emb_out, float* next_state_emb_out, long n, void* stream);

/* Replaces DiffusionReplayBuffer.update_target_action (simple_replay.py:198-200) for all groups at once:
 * buf_target_action[group[r], indices[r], :] = new_action[r, :].  Duplicate (group, index) pairs keep one of the
 * candidates, as torch's indexed assignment does. */
int ddp_replay_scatter_target(const ddp_batch_shape* shape, float* buf_target_action, long N, const float* new_action,
                              const int64_t* indices, const int32_t* group, long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DDIFFPG_B200_H */
