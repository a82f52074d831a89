"""Batch assembly / scatter-back around the hot path.

``ReplayKernels`` is a mixin for ``ddiffpg.replay.simple_replay.DiffusionReplayBuffer`` (:98-200): it overrides the two
methods that move batch data -- ``sample_batch`` (:150-163) and ``update_target_action`` (:198-200) -- with one gather /
one scatter launch of ``libddiffpg_b200.so`` and adds their fused forms ``sample_groups`` / ``scatter_groups`` (every
mode group of ``DiffusionGoalBuffer.sample_batch`` + ``AgentDDiffPG.update_net``, rows sorted by mode -- the layout of
``q_action_ascent_segments`` -- including the embedded states, in one launch).  Storage and bookkeeping
(``add_to_buffer``, ``remove``, ``get_buffer_size``, ``update_target_action_dim``) stay the reference's own code:
``DiffusionReplayBuffer`` below is ``ReplayKernels`` in front of the reference class when ``ddiffpg`` is importable, and
in front of a bare attribute holder otherwise (tests load the ``buf_*`` tensors from fixtures).
``GoalBufferKernels`` does the same for ``ddiffpg.replay.diffusion_replay.DiffusionGoalBuffer.sample_batch`` /
``add_temp_data`` (:250-332).
``add_embedding`` mirrors ``ddiffpg.utils.torch_util.add_embedding`` (:17-43), both branches.  Random draws stay in
torch / numpy as in the reference and can be injected (``indices=``, ``zero_indices=``) for reproducible comparisons.
"""
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
    """utils/torch_util.py:17-43: ``[state | embedding]`` where, of the first ``s = int(N*p)`` choices,
    * ``modes`` empty: ``s`` rows drawn without replacement get a zero embedding (``zero_indices`` injects the draw);
    * ``modes`` given (exploration with mode embeddings, ddiffpg.py:163-168): rows ``0 .. s-1`` get the embeddings of
      ``modes`` in equal consecutive blocks (the first block takes the remainder), all other rows ``embedding``.
    One gather launch either way."""
    _cuda(state, "state")
    n, O = state.shape
    E = embedding.shape[0]
    s = int(n * p)
    st = state.detach().to(torch.float32).contiguous()
    dev = st.device
    emb = embedding.detach().to(device=dev, dtype=torch.float32).reshape(1, E)
    zero = group = None
    if len(modes) != 0:
        per = [s // len(modes)] * len(modes)
        per[0] += s % len(modes)
        emb = torch.cat([emb] + [m.detach().to(device=dev, dtype=torch.float32).reshape(1, E) for m in modes])
        group = torch.zeros(n, dtype=torch.int32, device=dev)
        group[:s] = torch.repeat_interleave(torch.arange(1, len(modes) + 1, dtype=torch.int32, device=dev),
                                            torch.tensor(per, device=dev))
    else:
        if zero_indices is None and s:
            zero_indices = torch.as_tensor(np.random.choice(n, size=s, replace=False))
        if zero_indices is not None and len(zero_indices):
            zero = _u8(zero_indices, n, dev)
    emb = emb.contiguous()
    out = torch.empty(n, O + E, device=dev)
    if n:
        idx = torch.arange(n, device=dev)
        shape = BatchShape(O, 1, E, emb.shape[0])
        with torch.cuda.device(dev):
            check(lib().ddp_replay_gather(shape, ptr(st), None, None, None, None, None, n, ptr(idx), ptr(group), ptr(emb),
                                          ptr(zero), None, None, None, None, None, None, None, ptr(out), None, n,
                                          stream_ptr()), "ddp_replay_gather")
    return out


class ReplayKernels:
    """Mixin: the data-moving methods of ``DiffusionReplayBuffer`` on the device (see the module docstring)."""

    check_indices = True      # raise IndexError on out-of-range rows like the reference's indexing (costs one host sync)

    def available_indices(self, cluster_idx):
        """Rows whose trajectory id is in ``cluster_idx`` (the first line of simple_replay.py:151), ascending like
        ``torch.where(torch.isin(...))``: one table look-up per stored row (trajectory ids are small non-negative
        integers, diffusion_replay.py:84-118) instead of isin's sort / broadcast compare."""
        dev = self.buf_id.device
        # the answer only changes when the storage does, and every change of the reference's storage (add_to_buffer,
        # remove) re-allocates buf_id: cache per id set on (data_ptr, rows)
        key = (tuple(int(i) for i in cluster_idx), self.buf_id.data_ptr(), self.buf_id.shape[0])
        cache = self.__dict__.setdefault("_avail_cache", {})
        hit = cache.get(key)
        if hit is not None:
            return hit
        if len(cache) >= 64:
            cache.clear()
        out = self._available_indices(cluster_idx, dev)
        cache[key] = out
        return out

    def _available_indices(self, cluster_idx, dev):
        ids = torch.as_tensor(cluster_idx, device=dev).long().reshape(-1)
        if ids.numel() == 0:
            return torch.empty(0, dtype=torch.int64, device=dev)
        rows = self.buf_id.reshape(-1)
        if rows.is_floating_point() and not bool((rows == rows.round()).all()):
            return torch.where(torch.isin(self.buf_id, ids.to(self.buf_id.dtype)))[0]       # not ids at all: as written
        rows = rows.long()
        r_lo, r_hi = torch.aminmax(rows)
        i_lo, i_hi = torch.aminmax(ids)
        lo, hi = torch.stack([torch.minimum(r_lo, i_lo), torch.maximum(r_hi, i_hi)]).tolist()        # one host sync
        lo, hi = int(lo), int(hi)
        if hi - lo > 4 * rows.numel() + 65536:                                              # sparse id space: as written
            return torch.where(torch.isin(self.buf_id, ids.to(self.buf_id.dtype)))[0]
        lut = torch.zeros(hi - lo + 1, dtype=torch.bool, device=dev)
        lut[ids - lo] = True
        return torch.where(lut[rows - lo])[0]

    def get_buffer_size(self, cluster_idx):
        """simple_replay.py:183-186."""
        return 0 if self.buf_id is None else int(self.available_indices(cluster_idx).shape[0])

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
            # every converted operand is bound to a local: a temporary would be freed (and its block re-used by the next
            # conversion) before the launch reads it
            idx = indices.to(device=dev, dtype=torch.int64).contiguous()
            if self.check_indices:
                lo, hi = torch.aminmax(idx)
                if int(lo) < 0 or int(hi) >= N:
                    raise IndexError(f"replay index out of range: [{int(lo)}, {int(hi)}] for {N} stored rows")
            grp = None if group is None else group.to(device=dev, dtype=torch.int32).contiguous()
            if grp is not None and self.check_indices and embeddings is None:
                glo, ghi = torch.aminmax(grp)
                if int(glo) < 0 or int(ghi) >= K:
                    raise IndexError(f"target-action slot out of range: [{int(glo)}, {int(ghi)}] for {K} slots")
            obs, act, tgt = self.buf_obs.contiguous(), self.buf_action.contiguous(), self.buf_target_action.contiguous()
            rew, nobs = self.buf_reward.contiguous(), self.buf_next_obs.contiguous()
            done_u8 = self.buf_done.contiguous().view(torch.uint8)
            emb = embeddings.detach().to(device=dev, dtype=torch.float32).contiguous() if embeddings is not None else None
            zs, zn = _u8(zero_state, n, dev), _u8(zero_next, n, dev)
            shape = BatchShape(O, A, E, K)
            with torch.cuda.device(dev):
                check(lib().ddp_replay_gather(shape, ptr(obs), ptr(act), ptr(tgt), ptr(rew), ptr(nobs), ptr(done_u8), N,
                                              ptr(idx), ptr(grp), ptr(emb), ptr(zs), ptr(zn), ptr(out["obs"]),
                                              ptr(out["action"]), ptr(out["target"]), ptr(out["reward"]),
                                              ptr(out["next_obs"]), ptr(out["done"]), ptr(se), ptr(ne), n, stream_ptr()),
                      "ddp_replay_gather")
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
        (with ``embeddings`` [K,E]) the embedded states / next states of ``add_embedding``; ``zero_state`` /
        ``zero_next`` are the two independent zeroing draws of the reference (masks or index lists over the flat rows)."""
        dev = self.buf_obs.device
        idx = torch.cat([torch.as_tensor(i, device=dev).long() for i in group_indices])
        sizes = [len(i) for i in group_indices]
        group = torch.repeat_interleave(torch.arange(len(sizes), dtype=torch.int32, device=dev),
                                        torch.tensor(sizes, device=dev)).contiguous()
        seg_off = [0] + list(np.cumsum(sizes))
        o, se, ne = self._gather(idx, group, embeddings, zero_state, zero_next, want_embedded=embeddings is not None)
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
        dev = self.buf_target_action.device
        na = new_action.detach().to(device=dev, dtype=torch.float32).contiguous()
        idx = indices.to(device=dev, dtype=torch.int64).contiguous()
        grp = group.to(device=dev, dtype=torch.int32).contiguous()
        if self.check_indices:
            lo, hi = torch.aminmax(idx)
            glo, ghi = torch.aminmax(grp)
            if int(lo) < 0 or int(hi) >= N or int(glo) < 0 or int(ghi) >= K:
                raise IndexError(f"scatter out of range: rows [{int(lo)}, {int(hi)}] of {N}, slots [{int(glo)}, {int(ghi)}] of {K}")
        shape = BatchShape(self.buf_obs.shape[1], A, 0, K)
        with torch.cuda.device(dev):
            check(lib().ddp_replay_scatter_target(shape, ptr(self.buf_target_action), N, ptr(na), ptr(idx), ptr(grp), n,
                                                  stream_ptr()), "ddp_replay_scatter_target")


class GoalBufferKernels:
    """Mixin for ``ddiffpg.replay.diffusion_replay.DiffusionGoalBuffer``: ``sample_batch`` (:250-283) and
    ``add_temp_data`` (:285-332) with the replay rows of ALL mode groups gathered by one launch (``sample_groups``); the
    group split, the temp-buffer share of group 0 and the order of the random draws are the reference's.  Clustering,
    trajectory bookkeeping and the Q scheduler stay the reference's own code."""

    def _group_plan(self, batch_size):
        groups = [self.success_id + list(self.unsuccess_id)]
        for i in range(len(self.clusters)):
            groups.append(self.clusters[i] + self.unsuccess_clusters[i])
        sizes = [batch_size // len(groups)] * len(groups)
        sizes[0] += batch_size % len(groups)
        assert len(self.Qs) == len(groups) and len(self.Qs) == len(self.embeddings)
        if self.replay_buffer.buf_target_action is not None:
            assert len(self.Qs) == self.replay_buffer.buf_target_action.shape[0]
        return groups, sizes

    def _draw(self, batch_size, cluster_idx, if_add_temp):
        """The index draws of one ``add_temp_data`` call, in the reference's order (replay rows, then temp rows)."""
        temp_size = self.temp_state.shape[0]
        b_temp = 0
        if if_add_temp:
            buffer_size = self.replay_buffer.get_buffer_size(cluster_idx)
            b_temp = int((temp_size / (temp_size + buffer_size)) * batch_size)
        b_sample = batch_size - b_temp
        rows = temp_rows = None
        if b_sample != 0:
            avail = self.replay_buffer.available_indices(cluster_idx)
            rows = avail[torch.randint(avail.shape[0], size=(b_sample,), device=avail.device)]
        if b_temp != 0:
            temp_rows = torch.randint(temp_size, size=(b_temp,), device=self.device)
        return rows, temp_rows

    def _with_temp(self, parts, temp_rows, device):
        if temp_rows is not None:
            t = temp_rows.to(self.temp_state.device)
            temp = (self.temp_state[t], self.temp_action[t], self.temp_action[t], self.temp_reward[t],
                    self.temp_next_state[t], self.temp_done[t].float())
            dev = parts[0].device if parts is not None else torch.device(device)
            temp = tuple(x.to(device=dev, dtype=torch.float32) for x in temp)
            parts = temp if parts is None else tuple(torch.cat([a, b]) for a, b in zip(parts, temp))
        return tuple(x.to(device) for x in parts)

    @torch.no_grad()
    def sample_batch(self, batch_size, device=None):
        device = self.device if device is None else device
        groups, sizes = self._group_plan(batch_size)
        draws = [self._draw(sizes[i], groups[i], i == 0) for i in range(len(groups))]
        live = [i for i, (rows, _) in enumerate(draws) if rows is not None]
        data_list, parts = [], {}
        if live:
            rb = self.replay_buffer
            idx = torch.cat([draws[i][0] for i in live])
            slot = torch.repeat_interleave(torch.tensor(live, dtype=torch.int32, device=idx.device),
                                           torch.tensor([draws[i][0].shape[0] for i in live], device=idx.device))
            o, _, _ = rb._gather(idx, slot)
            off = 0
            for i in live:
                n = draws[i][0].shape[0]
                parts[i] = tuple(o[k][off:off + n] for k in ("obs", "action", "target", "reward", "next_obs", "done"))
                off += n
        for i in range(len(groups)):
            rows, temp_rows = draws[i]
            data_list.append({"Q": self.Qs[i], "batch": self._with_temp(parts.get(i), temp_rows, device),
                              "indices": rows, "embedding": self.embeddings[i]})
        return data_list

    def add_temp_data(self, batch_size, cluster_idx, target_idx, if_add_temp=True, device=None):
        device = self.device if device is None else device
        rows, temp_rows = self._draw(batch_size, cluster_idx, if_add_temp)
        parts = None
        if rows is not None:
            slot = torch.full((rows.shape[0],), int(target_idx), dtype=torch.int32, device=rows.device)
            o, _, _ = self.replay_buffer._gather(rows, slot)
            parts = tuple(o[k] for k in ("obs", "action", "target", "reward", "next_obs", "done"))
        return self._with_temp(parts, temp_rows, device), rows


def accelerate_goal_buffer(base):
    """``GoalBufferKernels`` in front of ``base`` (the reference's ``DiffusionGoalBuffer``); the inner replay buffer the
    reference constructs (diffusion_replay.py:45-48) is switched to the accelerated class in place."""
    def __init__(self, *args, **kwargs):
        base.__init__(self, *args, **kwargs)
        rb = self.replay_buffer
        if not isinstance(rb, ReplayKernels):
            rb.__class__ = accelerate_replay_buffer(type(rb))
    return type(base.__name__, (GoalBufferKernels, base), {"__init__": __init__, "__doc__": GoalBufferKernels.__doc__})


class _ReplayStorage:
    """The attributes ``DiffusionReplayBuffer.__init__`` (simple_replay.py:99-116) sets, nothing else: the base of
    ``DiffusionReplayBuffer`` where the reference package is not importable."""

    def __init__(self, capacity, obs_dim, action_dim, device="cuda"):
        self.obs_dim = (obs_dim,) if isinstance(obs_dim, int) else obs_dim
        self.action_dim, self.device, self.capacity, self.cur_capacity, self.last_sample = action_dim, device, int(capacity), 0, None
        self.buf_obs = self.buf_action = self.buf_next_obs = self.buf_reward = None
        self.buf_done = self.buf_id = self.buf_target_action = None


def accelerate_replay_buffer(base):
    """``ReplayKernels`` in front of ``base`` (the reference's ``DiffusionReplayBuffer`` or a subclass of it)."""
    return type(base.__name__, (ReplayKernels, base), {"__doc__": ReplayKernels.__doc__})


try:        # the reference's own storage / bookkeeping when it is installed next to this package
    from ddiffpg.replay.simple_replay import DiffusionReplayBuffer as _ReferenceReplayBuffer
except Exception:        # not installed (tests, benchmarks): attribute holder only
    _ReferenceReplayBuffer = _ReplayStorage
DiffusionReplayBuffer = accelerate_replay_buffer(_ReferenceReplayBuffer)
try:        # needs the reference's clustering dependencies (dtaidistance, scipy) as well
    from ddiffpg.replay.diffusion_replay import DiffusionGoalBuffer as _ReferenceGoalBuffer
    DiffusionGoalBuffer = accelerate_goal_buffer(_ReferenceGoalBuffer)
except Exception:
    DiffusionGoalBuffer = None
