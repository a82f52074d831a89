"""Batch assembly / scatter-back around the hot path.

``DiffusionReplayBuffer`` mirrors ``ddiffpg.replay.simple_replay.DiffusionReplayBuffer`` (:98-200) -- same storage,
same methods -- with ``sample_batch`` / ``update_target_action`` running as one gather / scatter launch of
``libddiffpg_b200.so``; ``add_embedding`` mirrors ``ddiffpg.utils.torch_util.add_embedding`` (:17-43).
``sample_groups`` / ``scatter_groups`` are the fused form ``DiffusionGoalBuffer.sample_batch`` +
``AgentDDiffPG.update_net`` need: every mode group (rows sorted by mode, the layout of ``q_action_ascent_segments``)
including the embedded states in one launch.  Random draws stay in torch / numpy as in the reference and can be
injected (``indices=``, ``zero_indices=``) so that results are reproducible against the reference's generators.
"""
from copy import deepcopy

import numpy as np
import torch

from ._lib import BatchShape, check, lib, ptr, stream_ptr


def _cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device (no CPU fallback)")
    return t


def _u8(mask, n, dev):
    if mask is None:
        return None
    m = torch.as_tensor(mask, device=dev)
    if m.dtype != torch.bool and m.dtype != torch.uint8:      # an index list
        z = torch.zeros(n, dtype=torch.uint8, device=dev)
        z[m.long()] = 1
        return z
    return m.to(torch.uint8).contiguous()


def add_embedding(state, embedding, p=0.5, modes=[], zero_indices=None):
    """[state | embedding], with the embedding zeroed on ``int(N*p)`` rows drawn without replacement
    (``zero_indices`` injects the draw).  The ``modes`` variant of the reference is not used by DDiffPG's update."""
    if len(modes) != 0:
        raise NotImplementedError("add_embedding(modes=...) is not on the accelerated path")
    _cuda(state, "state")
    n, O = state.shape
    E = embedding.shape[0]
    if zero_indices is None:
        s = int(n * p)
        zero_indices = torch.as_tensor(np.random.choice(n, size=s, replace=False)) if s else None
    st = state.detach().to(torch.float32).contiguous()
    emb = embedding.detach().to(device=st.device, dtype=torch.float32).contiguous()
    out = torch.empty(n, O + E, device=st.device)
    if n:
        idx = torch.arange(n, device=st.device)
        zero = _u8(zero_indices, n, st.device) if zero_indices is not None and len(zero_indices) else None
        shape = BatchShape(O, 1, E, 1)
        with torch.cuda.device(st.device):
            check(lib().ddp_replay_gather(shape, ptr(st), None, None, None, None, None, n, ptr(idx), None, ptr(emb),
                                          ptr(zero), None, None, None, None, None, None, None, ptr(out), None, n,
                                          stream_ptr()), "ddp_replay_gather")
    return out


class DiffusionReplayBuffer:
    def __init__(self, capacity, obs_dim, action_dim, device="cuda"):
        self.obs_dim = (obs_dim,) if isinstance(obs_dim, int) else obs_dim
        self.action_dim = action_dim
        self.device = device
        self.cur_capacity = 0
        self.capacity = int(capacity)
        self.last_sample = None
        self.buf_obs = self.buf_action = self.buf_next_obs = self.buf_reward = None
        self.buf_done = self.buf_id = self.buf_target_action = None

    @torch.no_grad()
    def add_to_buffer(self, trajectory, traj_id):
        obs, actions, target_actions, rewards, next_obs, dones = trajectory
        obs = obs.reshape(-1, *self.obs_dim)
        actions = actions.reshape(-1, self.action_dim)
        target_actions = target_actions.reshape(1, -1, self.action_dim)
        rewards = rewards.reshape(-1, 1)
        next_obs = next_obs.reshape(-1, *self.obs_dim)
        dones = dones.reshape(-1, 1).bool()
        traj_id = torch.ones_like(rewards) * traj_id
        if self.buf_obs is None:
            self.buf_obs, self.buf_action, self.buf_next_obs = obs, actions, next_obs
            self.buf_reward, self.buf_done, self.buf_id = rewards, dones, traj_id
            self.buf_target_action = target_actions
        else:
            target_actions = target_actions.repeat(self.buf_target_action.shape[0], 1, 1)
            self.buf_obs = torch.cat([self.buf_obs, obs])
            self.buf_action = torch.cat([self.buf_action, actions])
            self.buf_next_obs = torch.cat([self.buf_next_obs, next_obs])
            self.buf_reward = torch.cat([self.buf_reward, rewards])
            self.buf_done = torch.cat([self.buf_done, dones])
            self.buf_id = torch.cat([self.buf_id, traj_id])
            self.buf_target_action = torch.cat([self.buf_target_action, target_actions], dim=1)
        self.cur_capacity = self.buf_obs.shape[0]

    def available_indices(self, cluster_idx):
        dev = self.buf_id.device
        return torch.where(torch.isin(self.buf_id, torch.tensor(cluster_idx, device=dev)))[0]

    def _gather(self, indices, group, embeddings=None, zero_state=None, zero_next=None, want_embedded=False):
        _cuda(self.buf_obs, "the replay storage")
        dev = self.buf_obs.device
        n = indices.shape[0]
        O, A = self.buf_obs.shape[1], self.action_dim
        K, N = self.buf_target_action.shape[0], self.buf_obs.shape[0]
        E = embeddings.shape[1] if embeddings is not None else 0
        f = lambda *s: torch.empty(*s, device=dev)
        out = dict(obs=f(n, O), action=f(n, A), target=f(n, A), reward=f(n, 1), next_obs=f(n, O), done=f(n, 1))
        se = f(n, O + E) if want_embedded else None
        ne = f(n, O + E) if want_embedded else None
        if n:
            c = lambda t: t.contiguous()
            done_u8 = c(self.buf_done).view(torch.uint8)
            emb = c(embeddings.detach().to(device=dev, dtype=torch.float32)) if embeddings is not None else None
            shape = BatchShape(O, A, E, K)
            with torch.cuda.device(dev):
                check(lib().ddp_replay_gather(shape, ptr(c(self.buf_obs)), ptr(c(self.buf_action)),
                                              ptr(c(self.buf_target_action)), ptr(c(self.buf_reward)),
                                              ptr(c(self.buf_next_obs)), ptr(done_u8), N, ptr(c(indices.long())),
                                              ptr(group), ptr(emb), ptr(_u8(zero_state, n, dev)),
                                              ptr(_u8(zero_next, n, dev)), ptr(out["obs"]), ptr(out["action"]),
                                              ptr(out["target"]), ptr(out["reward"]), ptr(out["next_obs"]),
                                              ptr(out["done"]), ptr(se), ptr(ne), n, stream_ptr()), "ddp_replay_gather")
        return out, se, ne

    @torch.no_grad()
    def sample_batch(self, batch_size, cluster_idx, target_idx, device="cuda", indices=None):
        """simple_replay.py:150-163.  ``indices`` (positions into the available rows) replaces the randint draw."""
        available_idx = self.available_indices(cluster_idx)
        if indices is None:
            indices = torch.randint(available_idx.shape[0], size=(batch_size,), device=available_idx.device)
        indices = available_idx[indices.to(available_idx.device)]
        group = torch.full((indices.shape[0],), int(target_idx), dtype=torch.int32, device=indices.device)
        o, _, _ = self._gather(indices, group)
        return (o["obs"], o["action"], o["target"], o["reward"], o["next_obs"], o["done"]), indices

    @torch.no_grad()
    def sample_groups(self, group_indices, embeddings=None, zero_state=None, zero_next=None):
        """All mode groups in one launch.  ``group_indices[g]``: replay rows (absolute) drawn for group g.  Returns the
        six batch tensors with rows sorted by group, ``seg_off``, the flat absolute indices, the int32 group ids and
        (with ``embeddings`` [K,E]) the embedded states / next states of ``add_embedding``."""
        dev = self.buf_obs.device
        idx = torch.cat([torch.as_tensor(i, device=dev).long() for i in group_indices])
        sizes = [len(i) for i in group_indices]
        group = torch.repeat_interleave(torch.arange(len(sizes), dtype=torch.int32, device=dev),
                                        torch.tensor(sizes, device=dev))
        seg_off = [0] + list(np.cumsum(sizes))
        o, se, ne = self._gather(idx, group.contiguous(), embeddings, zero_state, zero_next,
                                 want_embedded=embeddings is not None)
        return o, [int(v) for v in seg_off], idx, group, se, ne

    @torch.no_grad()
    def update_target_action(self, new_action, indices, i):
        """simple_replay.py:198-200: buf_target_action[i, indices] = new_action."""
        group = torch.full((indices.shape[0],), int(i), dtype=torch.int32, device=indices.device)
        self.scatter_groups(new_action, indices, group)

    @torch.no_grad()
    def scatter_groups(self, new_action, indices, group):
        _cuda(self.buf_target_action, "the replay storage")
        if not self.buf_target_action.is_contiguous():
            self.buf_target_action = self.buf_target_action.contiguous()
        K, N, A = self.buf_target_action.shape
        n = indices.shape[0]
        if n == 0:
            return
        na = new_action.detach().to(device=self.buf_target_action.device, dtype=torch.float32).contiguous()
        shape = BatchShape(self.buf_obs.shape[1], A, 0, K)
        with torch.cuda.device(self.buf_target_action.device):
            check(lib().ddp_replay_scatter_target(shape, ptr(self.buf_target_action), N, ptr(na),
                                                  ptr(indices.long().contiguous()), ptr(group.contiguous()), n,
                                                  stream_ptr()), "ddp_replay_scatter_target")

    def remove(self, target_idx, device="cuda"):
        dev = self.buf_id.device
        remove_idx = torch.where(torch.isin(self.buf_id, torch.tensor(target_idx, device=dev)))[0]
        keep_idx = torch.ones(self.buf_obs.shape[0], dtype=bool, device=dev)
        keep_idx[remove_idx] = False
        self.buf_obs, self.buf_action = self.buf_obs[keep_idx], self.buf_action[keep_idx]
        self.buf_next_obs, self.buf_reward = self.buf_next_obs[keep_idx], self.buf_reward[keep_idx]
        self.buf_done, self.buf_id = self.buf_done[keep_idx], self.buf_id[keep_idx]
        self.buf_target_action = self.buf_target_action[:, keep_idx]
        self.cur_capacity = self.buf_obs.shape[0]

    def get_buffer_size(self, cluster_idx):
        if self.buf_id is None:
            return 0
        return self.available_indices(cluster_idx).shape[0]

    def update_target_action_dim(self, indices):
        if len(indices) == 0:
            return
        new_target_action = [deepcopy(self.buf_target_action[0])]
        assert max(indices) < self.buf_target_action.shape[0]
        for idx in indices:
            new_target_action.append(deepcopy(self.buf_action if idx == -1 else self.buf_target_action[idx]))
        self.buf_target_action = torch.stack(new_target_action)
