// N2: batch assembly and scatter-back around the hot path, one launch for all mode groups.
//
// Reference semantics (paths relative to the reference repo):
//   DiffusionReplayBuffer.sample_batch (the six gathers)        ddiffpg/replay/simple_replay.py:150-163
//   add_embedding (state | embedding, zeroed for a row subset)  ddiffpg/utils/torch_util.py:17-43
//   DiffusionReplayBuffer.update_target_action (scatter)        ddiffpg/replay/simple_replay.py:198-200
// The random choices (row indices, which rows lose their embedding) stay with the caller and arrive as index /
// mask arrays, so that the reference's generators can be replayed bit for bit.
#include "common.cuh"

namespace ddp {
namespace {

struct GatherArgs {
    const float *buf_obs, *buf_action, *buf_target, *buf_reward, *buf_next_obs;
    const uint8_t* buf_done;
    long N;
    const int64_t* indices;
    const int32_t* group;
    const float* emb;
    const uint8_t *zero_state, *zero_next;
    float *obs, *action, *target, *reward, *next_obs, *done, *state_emb, *next_emb;
    long n;
    int O, A, E, n_groups;
};

// one thread per output element of a virtual row [obs | action | target | reward | next_obs | done | state_emb | next_emb]
__global__ void replay_gather_kernel(GatherArgs a) {
    const int W = 4 * a.O + 2 * a.A + 2 * a.E + 2;
    const long total = a.n * W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / W;
        int c = (int)(i - r * W);
        const long src = a.indices[r];
        const bool ok = src >= 0 && src < a.N;
        const int g = a.group ? a.group[r] : 0;
        if (c < a.O) { if (a.obs) a.obs[r * a.O + c] = ok ? a.buf_obs[src * a.O + c] : 0.f; continue; }
        c -= a.O;
        if (c < a.A) { if (a.action) a.action[r * a.A + c] = ok ? a.buf_action[src * a.A + c] : 0.f; continue; }
        c -= a.A;
        if (c < a.A) {
            if (a.target) a.target[r * a.A + c] = (ok && g >= 0 && g < a.n_groups) ? a.buf_target[((long)g * a.N + src) * a.A + c] : 0.f;
            continue;
        }
        c -= a.A;
        if (c < 1) { if (a.reward) a.reward[r] = ok ? a.buf_reward[src] : 0.f; continue; }
        c -= 1;
        if (c < a.O) { if (a.next_obs) a.next_obs[r * a.O + c] = ok ? a.buf_next_obs[src * a.O + c] : 0.f; continue; }
        c -= a.O;
        if (c < 1) { if (a.done) a.done[r] = (ok && a.buf_done[src]) ? 1.f : 0.f; continue; }
        c -= 1;
        const int WE = a.O + a.E;
        const bool second = c >= WE;
        if (second) c -= WE;
        float* dst = second ? a.next_emb : a.state_emb;
        if (!dst) continue;
        float v;
        if (c < a.O) v = ok ? (second ? a.buf_next_obs : a.buf_obs)[src * a.O + c] : 0.f;
        else {
            const uint8_t* z = second ? a.zero_next : a.zero_state;
            const bool zero = (z && z[r]) || !a.emb || g < 0 || g >= a.n_groups;
            v = zero ? 0.f : a.emb[(long)g * a.E + (c - a.O)];
        }
        dst[r * WE + c] = v;
    }
}

__global__ void replay_scatter_kernel(float* __restrict__ buf_target, long N, int A, int n_groups,
                                      const float* __restrict__ new_action, const int64_t* __restrict__ indices,
                                      const int32_t* __restrict__ group, long n) {
    const long total = n * A;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / A;
        const int c = (int)(i - r * A);
        const long dst = indices[r];
        const int g = group ? group[r] : 0;
        if (dst < 0 || dst >= N || g < 0 || g >= n_groups) continue;
        buf_target[((long)g * N + dst) * A + c] = new_action[i];
    }
}

}  // namespace
}  // namespace ddp

extern "C" {

int ddp_replay_gather(const ddp_batch_shape* s, const float* buf_obs, const float* buf_action,
                      const float* buf_target_action, const float* buf_reward, const float* buf_next_obs,
                      const uint8_t* buf_done, long N, const int64_t* indices, const int32_t* group,
                      const float* embeddings, const uint8_t* zero_state, const uint8_t* zero_next, float* obs_out,
                      float* action_out, float* target_action_out, float* reward_out, float* next_obs_out,
                      float* done_out, float* state_emb_out, float* next_state_emb_out, long n, void* stream) {
    using namespace ddp;
    if (!s) DDP_FAIL(DDP_ERR_ARG, "batch shape is NULL");
    if (s->O <= 0 || s->A <= 0 || s->E < 0 || s->n_groups < 1)
        DDP_FAIL(DDP_ERR_SHAPE, "batch shape: need O>0, A>0, E>=0, n_groups>=1 (got O=%d A=%d E=%d groups=%d)", s->O, s->A, s->E, s->n_groups);
    if (n < 0 || N <= 0) DDP_FAIL(DDP_ERR_SHAPE, "ddp_replay_gather: need n >= 0 and a non-empty buffer");
    if (n == 0) return DDP_OK;
    if (!indices) DDP_FAIL(DDP_ERR_ARG, "ddp_replay_gather: indices is NULL");
    if ((obs_out || state_emb_out) && !buf_obs) DDP_FAIL(DDP_ERR_ARG, "ddp_replay_gather: buf_obs is NULL");
    if ((next_obs_out || next_state_emb_out) && !buf_next_obs) DDP_FAIL(DDP_ERR_ARG, "ddp_replay_gather: buf_next_obs is NULL");
    if ((action_out && !buf_action) || (target_action_out && !buf_target_action) || (reward_out && !buf_reward) ||
        (done_out && !buf_done))
        DDP_FAIL(DDP_ERR_ARG, "ddp_replay_gather: an output is requested whose source buffer is NULL");
    GatherArgs a;
    a.buf_obs = buf_obs; a.buf_action = buf_action; a.buf_target = buf_target_action; a.buf_reward = buf_reward;
    a.buf_next_obs = buf_next_obs; a.buf_done = buf_done; a.N = N; a.indices = indices; a.group = group;
    a.emb = embeddings; a.zero_state = zero_state; a.zero_next = zero_next;
    a.obs = obs_out; a.action = action_out; a.target = target_action_out; a.reward = reward_out;
    a.next_obs = next_obs_out; a.done = done_out; a.state_emb = state_emb_out; a.next_emb = next_state_emb_out;
    a.n = n; a.O = s->O; a.A = s->A; a.E = s->E; a.n_groups = s->n_groups;
    const long total = n * (4L * s->O + 2 * s->A + 2 * s->E + 2);
    long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    replay_gather_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    DDP_LAUNCH_CHECK("replay_gather_kernel");
    return DDP_OK;
}

int ddp_replay_scatter_target(const ddp_batch_shape* s, float* buf_target_action, long N, const float* new_action,
                              const int64_t* indices, const int32_t* group, long n, void* stream) {
    using namespace ddp;
    if (!s) DDP_FAIL(DDP_ERR_ARG, "batch shape is NULL");
    if (s->A <= 0 || s->n_groups < 1) DDP_FAIL(DDP_ERR_SHAPE, "batch shape: need A>0, n_groups>=1");
    if (n < 0 || N <= 0) DDP_FAIL(DDP_ERR_SHAPE, "ddp_replay_scatter_target: need n >= 0 and a non-empty buffer");
    if (n == 0) return DDP_OK;
    if (!buf_target_action || !new_action || !indices) DDP_FAIL(DDP_ERR_ARG, "ddp_replay_scatter_target: NULL argument");
    long blocks = (n * s->A + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    replay_scatter_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(buf_target_action, N, s->A, s->n_groups,
                                                                           new_action, indices, group, n);
    DDP_LAUNCH_CHECK("replay_scatter_kernel");
    return DDP_OK;
}

}  // extern "C"
