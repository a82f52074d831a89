"""Host-side mirror of the agent-level hot-path methods of ``ddiffpg/algo/ddiffpg.py`` and
``ddiffpg/algo/ac_base.py`` (same names, argument meaning and return values).

``update_target_action`` replaces ``AgentDDiffPG.update_target_action`` (ddiffpg.py:358-373);
``optimizer_update`` is ``ActorCriticBase.optimizer_update`` (ac_base.py:83-92) verbatim in behaviour;
``update_actor`` is ``AgentDDiffPG.update_actor`` (ddiffpg.py:353-356).  ``HotPathMixin`` bundles them so an
agent class can inherit the accelerated versions without touching the rest of its code.
"""
from copy import deepcopy

import torch
from torch.nn.utils import clip_grad_norm_

from . import _lib
from . import dist as ddist
from ._lib import check, lib, ptr, stream_ptr
from .models import _PackCache, pack_critics

_ws_cache = {}


def _workspace(name, nbytes, dev):
    buf = _ws_cache.get((name, str(dev)))
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
        _ws_cache[(name, str(dev))] = buf
    return buf


def q_action_ascent_segments(critics, obs, action, seg_off, iters=20, lr=0.03, eps=1e-5, max_norm=1.0,
                             betas=(0.9, 0.999), lim=1 - 1e-5, mean_counts=None, cache=None, return_norms=False,
                             precision=None, process_group=None, gsq_reduce=None):
    """Run the reference's action-ascent loop for several mode segments in one launch sequence.

    ``obs`` [B, O] / ``action`` [B, A] hold the rows of all modes, sorted by mode; ``seg_off`` (len
    n_modes+1) delimits them; ``critics[m]`` is the critic of mode m.  ``action`` is updated IN PLACE like
    the reference (ddiffpg.py:361-369).  Each segment is one reference call: its own 1/B_m in ``-Q.mean()``
    (or ``mean_counts[m]`` when the segment is a shard of a larger batch) and its own clip norm.

    Row-sharded batches: by default every rank runs the reference on its own rows (shard-local semantics,
    no collective).  With ``process_group`` (and ``mean_counts`` = the global rows per mode) the per-mode sum of g^2 is
    all-reduced between the gradient pass and the Adam step of every iteration (K floats, ``ddp_q_action_ascent_sharded``),
    so the clip norm is the one of the whole mode batch and the ranks take exactly the steps a single process would take
    on the gathered batch; the returned mean|a| and norms are then global as well.  ``gsq_reduce(t)`` replaces the
    all-reduce by any in-place reduction of the [n_modes] device tensor (tests, other transports).
    Returns (mean|a| per segment [n_modes] tensor, optional pre-clip norms [n_modes, iters])."""
    if max_norm is None:
        max_norm = float("inf")
    if not action.is_cuda or action.dtype != torch.float32 or not action.is_contiguous():
        raise RuntimeError("action must be a contiguous fp32 CUDA tensor (it is updated in place)")
    cache = cache if cache is not None else critics[0]._cache if len(critics) == 1 else _PackCache()
    packed, shape, prec = pack_critics(list(critics), cache, precision)
    dev = packed.device
    obs = obs.detach().to(device=dev, dtype=torch.float32).contiguous()
    B = action.shape[0]
    n_modes = len(critics)
    seg_off = [int(v) for v in seg_off]
    if len(seg_off) != n_modes + 1:
        raise ValueError("seg_off must have n_modes + 1 entries")
    counts = [seg_off[m + 1] - seg_off[m] for m in range(n_modes)] if mean_counts is None else list(mean_counts)
    mean_abs = torch.zeros(n_modes, device=dev)
    norms = torch.zeros(n_modes, iters, device=dev) if return_norms else None
    if process_group is not None and gsq_reduce is None and ddist.world_info(process_group)[1] > 1:
        if mean_counts is None:
            raise ValueError("process_group needs mean_counts (the global rows per mode) for the 1/B factor")
        gsq_reduce = lambda t: torch.distributed.all_reduce(t, group=process_group)          # noqa: E731
    if B and gsq_reduce is None:
        with torch.cuda.device(dev):
            ws_bytes = lib().ddp_q_ascent_workspace_bytes(shape, B, iters)
            ws = _workspace("q_ascent", ws_bytes, dev)
            check(lib().ddp_q_action_ascent(shape, ptr(packed), _lib.i64_array(seg_off), _lib.i64_array(counts),
                                            ptr(obs), ptr(action), iters, lr, betas[0], betas[1], eps, max_norm,
                                            lim, ptr(mean_abs), ptr(norms), B, prec, ptr(ws), ws_bytes,
                                            stream_ptr()), "ddp_q_action_ascent")
    elif B:
        with torch.cuda.device(dev):
            ws_bytes = lib().ddp_q_ascent_workspace_bytes(shape, B, iters)
            ws = _workspace("q_ascent", ws_bytes, dev)
            failure = []

            def _reduce(gsq_dev, n, _stream, _user):
                # the vector lives inside the workspace: hand it to torch as a view (the call runs on torch's current
                # stream, so the reduction is ordered between the gradient pass and the Adam step)
                try:
                    off = int(gsq_dev) - ws.data_ptr()
                    gsq_reduce(ws[off:off + 4 * n].view(torch.float32))
                    return 0
                except BaseException as exc:        # never unwind through the C frames
                    failure.append(exc)
                    return 1

            cb = _lib.GSQ_REDUCE_FN(_reduce)
            rc = lib().ddp_q_action_ascent_sharded(shape, ptr(packed), _lib.i64_array(seg_off), _lib.i64_array(counts),
                                                   ptr(obs), ptr(action), iters, lr, betas[0], betas[1], eps, max_norm,
                                                   lim, ptr(mean_abs), ptr(norms), B, prec, ptr(ws), ws_bytes,
                                                   stream_ptr(), cb, None)
            if failure:
                raise failure[0]
            check(rc, "ddp_q_action_ascent_sharded")
    elif gsq_reduce is not None:
        # a rank without rows still takes part in the exchange step of every iteration
        for _ in range(iters):
            gsq_reduce(torch.zeros(n_modes, device=dev))
    if process_group is not None and ddist.world_info(process_group)[1] > 1:
        local = torch.tensor([seg_off[m + 1] - seg_off[m] for m in range(n_modes)], dtype=torch.float32, device=dev)
        tot = torch.stack([mean_abs * local, local])
        torch.distributed.all_reduce(tot, group=process_group)
        mean_abs = tot[0] / tot[1].clamp_min(1.0)
    return (mean_abs, norms) if return_norms else mean_abs


def update_target_action(obs, action, critic, action_lr=0.03, update_times=20, max_grad_norm=1.0):
    """``AgentDDiffPG.update_target_action(obs, action, critic)`` (ddiffpg.py:358-373).

    ``action`` is clamped and updated in place; returns ``(mean |action| as a Python float, deep copy of the
    new actions)`` like the reference.  The critic's ``requires_grad`` flags are left as the reference leaves
    them (True on exit)."""
    critic.requires_grad_(False)
    mean_abs = q_action_ascent_segments([critic], obs, action, [0, action.shape[0]], iters=update_times,
                                        lr=action_lr, eps=1e-5, max_norm=max_grad_norm)
    update = deepcopy(action.detach())
    critic.requires_grad_(True)
    return mean_abs[0].item(), update


def get_actions(actor, obs, sample=True, noise_type="mixed", std_min=0.05, std_max=0.8, std=None, noise=None,
                expl_noise=None):
    """``AgentDDiffPG.get_actions`` (ddiffpg.py:82-100) after the optional obs normalisation: ``actor(obs)``
    followed by ``add_mixed_normal_noise`` (type 'mixed') or ``add_normal_noise`` (type 'fixed', ``std``) with
    ``out_bounds=[-1, 1]`` -- one fused launch."""
    if not sample:
        return actor.get_actions(obs, noise=noise)
    if noise_type == "mixed":
        return actor.get_actions(obs, noise=noise, expl_std=(std_min, std_max), expl_noise=expl_noise)
    if noise_type == "fixed":
        return actor.get_actions(obs, noise=noise, expl_std=(std, std), expl_noise=expl_noise)
    raise NotImplementedError(noise_type)


def get_tgt_policy_actions(actor_target, obs, sample=True, tgt_pol_std=0.8, tgt_pol_noise_bound=0.2, noise=None,
                           expl_noise=None):
    """``AgentDDiffPG.get_tgt_policy_actions`` (ddiffpg.py:102-110): target actor + clipped Gaussian smoothing."""
    if not sample:
        return actor_target.get_actions(obs, noise=noise)
    return actor_target.get_actions(obs, noise=noise, expl_std=(tgt_pol_std, tgt_pol_std), expl_noise=expl_noise,
                                    noise_bound=tgt_pol_noise_bound)


def optimizer_update(optimizer, objective, max_grad_norm=1.0):
    """``ActorCriticBase.optimizer_update`` (ac_base.py:83-92): zero_grad, backward, clip, step."""
    optimizer.zero_grad(set_to_none=True)
    objective.backward()
    if max_grad_norm is not None:
        grad_norm = clip_grad_norm_(parameters=optimizer.param_groups[0]["params"], max_norm=max_grad_norm)
    else:
        grad_norm = None
    optimizer.step()
    return grad_norm


def update_actor(actor, actor_optimizer, obs, target_action, max_grad_norm=1.0, noise=None, timesteps=None):
    """``AgentDDiffPG.update_actor`` (ddiffpg.py:353-356): returns (loss, pre-clip grad norm) as floats."""
    actor_loss = actor.get_loss(obs, target_action, noise=noise, timesteps=timesteps)
    grad_norm = optimizer_update(actor_optimizer, actor_loss, max_grad_norm)
    return actor_loss.item(), grad_norm.item()


def _critic_pack_cache(c, precision):
    """Pack cache of `c` for `precision`: the update's pack is kept apart from an inference pack of another precision."""
    if getattr(c, "precision", "fp32") == precision:
        return c._cache
    name = "_cache_" + precision
    if not hasattr(c, name):
        setattr(c, name, _PackCache())
    return getattr(c, name)


def _critic_loss_into(critic, critic_target, obs, action, next_obs, next_actions, reward, done, gamma_n, precision, loss,
                      grads, ws_holder):
    """The C-ABI call behind ``critic_loss_and_grads``: ``loss`` (1 float, zeroed here) and ``grads`` (flat) are written
    in place; ``ws_holder`` is a dict that keeps the kernel workspace alive (a captured graph holds pointers into it)."""
    packed, shape, prec = pack_critics([critic], _critic_pack_cache(critic, precision), precision)
    # the target critic is written through param.data by soft_update (ddiffpg.py:266): always re-packed (see _PackCache)
    packed_t, _, _ = pack_critics([critic_target], _critic_pack_cache(critic_target, precision), precision, force=True)
    dev = packed.device
    f = lambda x: x.detach().to(device=dev, dtype=torch.float32).contiguous()
    obs, action, next_obs, next_actions = f(obs), f(action), f(next_obs), f(next_actions)
    reward, done = f(reward).reshape(-1), f(done).reshape(-1)
    B = obs.shape[0]
    if reward.numel() != B or done.numel() != B:
        raise ValueError("reward and done must have one entry per row")
    if grads.numel() != lib().ddp_q_grad_count(shape):
        raise ValueError("gradient buffer does not match the critic")
    loss.zero_()
    with torch.cuda.device(dev):
        ws_bytes = lib().ddp_q_critic_train_workspace_bytes(shape, B, prec)
        if ws_holder.get("ws") is None or ws_holder["ws"].numel() < ws_bytes or ws_holder["ws"].device != dev:
            ws_holder["ws"] = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=dev)
        ws = ws_holder["ws"]
        check(lib().ddp_q_critic_loss_fwd_bwd(shape, ptr(packed), ptr(packed_t), ptr(obs), ptr(action), ptr(next_obs),
                                              ptr(next_actions), ptr(reward), ptr(done), float(gamma_n), ptr(loss),
                                              ptr(grads), B, prec, ptr(ws), ws_bytes, stream_ptr()),
              "ddp_q_critic_loss_fwd_bwd")


_CRITIC_WS = {}


def critic_loss_and_grads(critic, critic_target, obs, action, next_obs, next_actions, reward, done, gamma_n,
                          precision=None):
    """Loss and flat gradient (state_dict order) of the reference's critic objective (ddiffpg.py:325-349) for one
    critic: target = min of the two C51-projected target heads, loss = BCE(Q1, target) + BCE(Q2, target).
    ``precision``: "fp32" (FMA kernels, 1e-4 parity) or "bf16" (tcgen05 GEMMs, 1e-2 parity); default
    ``critic.train_precision``."""
    precision = precision or getattr(critic, "train_precision", "fp32")
    dev = next(critic.parameters()).device
    n = sum(p.numel() for p in critic.parameters())
    grads = torch.empty(n, device=dev, dtype=torch.float32)
    loss = torch.zeros(1, device=dev, dtype=torch.float32)
    _critic_loss_into(critic, critic_target, obs, action, next_obs, next_actions, reward, done, gamma_n, precision, loss,
                      grads, _CRITIC_WS.setdefault(str(dev), {}))
    return loss[0], grads


def update_critic(critic, critic_target, critic_optimizer, obs, action, reward, next_obs, next_actions, done,
                  gamma_n=0.99, max_grad_norm=1.0, process_group=None, sync=True):
    """``AgentDDiffPG.update_critic`` (ddiffpg.py:322-351) with ``next_actions`` already sampled
    (``get_tgt_policy_actions``): fused target/loss/backward, then the reference's own clip + optimizer step on
    the ``.grad`` fields.  Returns ``(critic, loss float, pre-clip grad norm float)``.

    ``process_group``: data-parallel replicas, every rank holding an equal share of one batch -- the flat gradient and
    the loss are averaged over the ranks (one all-reduce, the critic's counterpart of H3's exchange step) before the
    identical clip + step on every replica.  ``sync=False`` returns the loss and norm as 0-dim device tensors instead
    of floats (no host synchronisation inside the call)."""
    loss, grads = critic_loss_and_grads(critic, critic_target, obs, action, next_obs, next_actions, reward, done,
                                        gamma_n)
    _, world = ddist.world_info(process_group) if process_group is not None else (0, 1)
    if world > 1:
        flat = torch.cat([grads, loss.reshape(1)])            # the loss rides in the gradient buffer: one collective
        ddist.allreduce_sum_(flat, group=process_group)
        flat.div_(world)
        grads, loss = flat[:-1], flat[-1]
    critic_optimizer.zero_grad(set_to_none=True)
    off = 0
    for p in critic.parameters():
        p.grad = grads[off:off + p.numel()].view(p.shape)
        off += p.numel()
    if max_grad_norm is not None:
        grad_norm = clip_grad_norm_(parameters=critic_optimizer.param_groups[0]["params"], max_norm=max_grad_norm)
    else:
        grad_norm = None
    critic_optimizer.step()
    if not sync:
        return critic, loss, grad_norm
    return critic, loss.item(), (grad_norm.item() if grad_norm is not None else None)


@torch.no_grad()
def soft_update(target_net, current_net, tau):
    """``ddiffpg/utils/torch_util.py:9-12`` plus the cache invalidation ``.data`` writes cannot signal.  The reference's
    per-parameter ``tar.data.copy_(cur.data * tau + tar.data * (1.0 - tau))`` (three launches per tensor, 48 per critic) is
    issued as three multi-tensor launches over all parameters -- the same two products and one sum per element, hence the
    same bits."""
    tars = [p.data for p in target_net.parameters()]
    curs = [p.data for p in current_net.parameters()]
    if tars:
        scaled = torch._foreach_mul(curs, tau)
        torch._foreach_mul_(tars, 1.0 - tau)
        torch._foreach_add_(tars, scaled)
    if hasattr(target_net, "mark_dirty"):
        target_net.mark_dirty()


class FusedActorTrainer:
    """Whole ``update_actor`` on the device: loss + 12 gradients (one C-ABI call), the data-parallel gradient
    all-reduce, then clip + AdamW on a flat parameter vector (``ddp_clip_adamw_step``).

    The module's parameters are re-pointed at slices of one flat fp32 buffer (state_dict keys, shapes and
    values unchanged), so the flat gradient of ``ddp_actor_loss_fwd_bwd`` lines up with it.  Hyper-parameters default
    to the reference's ``torch.optim.AdamW(actor.parameters(), actor_lr)`` (ac_base.py:52) and ``max_grad_norm`` 1.0.

    Data parallel (``torch.distributed`` initialised, or ``process_group`` given; ``process_group=False`` switches
    the collective off): the backward hands the gradient over in four groups, last layer first
    (``ddp_actor_loss_fwd_bwd_ev``); each group is summed over the ranks on a side stream as soon as it is final, so
    only the reduction of the last group (time_mlp + net.mlp.0) is not hidden behind the rest of the backward.  The
    loss rides in the same buffer (one float behind the gradient): there is no separate collective for it."""

    GROUPS = 4

    def __init__(self, actor, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=1.0,
                 process_group=None, precision=None, graph=False, buckets=None):
        self.actor = actor
        # collectives per step: 4 = every gradient group as soon as it is final, 2 = groups 0-2 together (behind the
        # layer-1 weight gradient) + the last group, 1 = one all-reduce behind the backward; None: by world size (see
        # _bucket_plan)
        if buckets not in (None, 1, 2, 4):
            raise ValueError("buckets must be None, 1, 2 or 4")
        self.buckets = buckets
        # graph=True: the whole step (re-pack, forward/backward, all-reduce, clip + AdamW) is captured once per batch
        # shape into a CUDA graph and replayed; the Adam step count lives on the device
        self.use_graph = graph
        self._graphs = {}
        self.precision = precision          # None: actor.train_precision ("fp32" | "bf16")
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = float("inf") if max_grad_norm is None else max_grad_norm
        self.group = process_group
        params = actor._params()
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        off, self._offsets = 0, [0]
        for p in params:
            self.flat[off:off + p.numel()].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + p.numel()].view(p.shape)
            off += p.numel()
            self._offsets.append(off)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = 0
        self._norm = torch.zeros(1, device=dev)
        self._scratch = torch.zeros(640, device=dev)          # DDP_ADAMW_SCRATCH_FLOATS
        self._step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self._side = None                   # stream of the overlapped all-reduce, events of the gradient groups
        self._events = None
        actor.mark_dirty()

    # ------------------------------------------------------------------ plumbing
    def world_size(self):
        if self.group is False:
            return 1
        if self.group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return torch.distributed.get_world_size(self.group)
        return 1

    def launches_per_step(self):
        """Own kernels of one step on the tensor path: re-pack 6 (small jobs, the three T-row layers of the time branch,
        16-bit operands, training operands), prep 1, fused forward 1, backward row GEMMs 3, dW GEMMs 4, time branch 6
        (transpose, three T-row weight gradients, two T-row products), norm / scalars / clip + AdamW 3."""
        return 24

    def close(self):
        """Release the captured CUDA graphs (and with them the collectives they hold).  Call before
        ``torch.distributed.destroy_process_group()``: tearing NCCL down under a live graph that captured an
        all-reduce blocks."""
        dev = self.flat.device
        torch.cuda.synchronize(dev)
        self._graphs.clear()
        torch.cuda.synchronize(dev)

    def _check_flat(self):
        off = 0
        for p in self.actor._params():
            if p.data_ptr() != self.flat.data_ptr() + 4 * off:
                raise RuntimeError("actor parameters were re-allocated since FusedActorTrainer was built "
                                   "(e.g. .to() or load into new storage); build a new trainer")
            off += p.numel()

    def _group_slices(self, n):
        """Element ranges of the four gradient groups in the flat [n + 1] buffer (state_dict order: time_mlp.1/.3,
        mlp.0, mlp.2, mlp.4, mlp.6, then the loss): group 0 = mlp.6 + loss, 1 = mlp.4, 2 = mlp.2, 3 = the rest."""
        o = self._offsets
        return [(o[10], n + 1), (o[8], o[10]), (o[6], o[8]), (0, o[6])]

    def _bucket_plan(self, n):
        """[(event index to wait for, lo, hi)] of the all-reduces of one step.  The flat buffer is in state_dict order,
        so groups 0-2 (mlp.6 + loss, mlp.4, mlp.2) are one contiguous range and so is the whole buffer."""
        sl = self._group_slices(n)
        buckets = self.buckets
        if buckets is None:
            # measured at 8 ranks, 131 072 rows per rank (profiles/r02/bench_train_n8_b{4,2,1}.json): 4 buckets 1.219 ms,
            # 2 buckets 1.269 ms, 1 bucket 1.226 ms per step -- per-group reduction stays the default at every world size
            buckets = 4
        if buckets == 4:
            return [(g, lo, hi) for g, (lo, hi) in enumerate(sl)]
        if buckets == 2:
            return [(2, sl[2][0], sl[0][1]), (3, sl[3][0], sl[3][1])]
        return [(3, 0, n + 1)]

    def step(self, state, action, noise=None, timesteps=None, global_batch=None):
        """One training step; returns (loss, pre-clip grad norm) as 0-dim device tensors (no host sync).
        ``global_batch``: rows over all data-parallel ranks (defaults to local rows x world size)."""
        self._check_flat()
        actor = self.actor
        dev = self.flat.device
        B = action.shape[0]
        if global_batch is None:
            global_batch = B * self.world_size()
        if noise is None:
            noise = torch.randn(action.shape, device=dev, dtype=torch.float32)
        if timesteps is None:
            timesteps = torch.randint(0, actor.diffusion_iter, (B,), device=dev)
        if self.use_graph:
            return self._step_graph(state, action, noise, timesteps, global_batch)
        bufs = self._buffers(B, None)
        loss, norm = self._step_body(state, action, noise, timesteps, global_batch, bufs)
        return loss.clone(), norm.clone()

    def _buffers(self, B, owner):
        """Gradient (+ loss) buffer and kernel workspace of a step.  A captured graph keeps raw pointers into them, so
        every graph owns its own pair (``owner`` = its entry); eager steps share one that only ever grows."""
        actor = self.actor
        dev = self.flat.device
        prec = _lib.PRECISIONS[self.precision or actor.train_precision]
        n = self.flat.numel()
        with torch.cuda.device(dev):
            ws_bytes = lib().ddp_actor_train_workspace_bytes(actor._shape(), B, prec)
        holder = owner if owner is not None else self.__dict__.setdefault("_eager", {})
        if holder.get("ws") is None or holder["ws"].numel() < ws_bytes:
            holder["ws"] = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=dev)
        if holder.get("gbuf") is None:
            holder["gbuf"] = torch.zeros(n + 4, device=dev, dtype=torch.float32)      # [gradient | loss | pad]
        return holder

    def _step_body(self, state, action, noise, timesteps, global_batch, bufs):
        actor = self.actor
        dev = self.flat.device
        n = self.flat.numel()
        precision = self.precision or actor.train_precision
        packed, shape, prec = actor._packed(precision, need=2)
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        state, action, noise = f32(state), f32(action), f32(noise)
        timesteps = timesteps.detach().to(device=dev, dtype=torch.int64).contiguous()
        B = action.shape[0]
        gbuf, ws = bufs["gbuf"], bufs["ws"]
        grads, loss = gbuf[:n], gbuf[n:n + 1]
        loss.zero_()
        world = self.world_size()
        params = actor._params()
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            ev_arr = None
            if world > 1:
                if self._side is None:
                    self._side = torch.cuda.Stream()
                    self._events = [torch.cuda.Event() for _ in range(self.GROUPS)]
                    for ev in self._events:          # torch creates the CUDA event lazily: force the handles into being
                        ev.record(main)
                ev_arr = (_lib.c_void_p * self.GROUPS)(*[ev.cuda_event for ev in self._events])
            check(lib().ddp_actor_loss_fwd_bwd_ev(shape, ptr(packed), _lib.ptr_array([p.detach() for p in params]),
                                                  ptr(state), ptr(action), ptr(noise), ptr(timesteps),
                                                  1.0 / (global_batch * actor.action_dim), ptr(loss), ptr(grads), B,
                                                  prec, ptr(ws), ws.numel(), stream_ptr(), ev_arr),
                  "ddp_actor_loss_fwd_bwd_ev")
            if world > 1:
                # the one exchange step of the path: each gradient group is summed over NVLink on the side stream as soon
                # as the backward has finished it; the main stream joins before the clip
                side = self._side
                group = None if self.group in (None, False) else self.group
                with torch.cuda.stream(side):
                    for g, lo, hi in self._bucket_plan(n):
                        side.wait_event(self._events[g])
                        torch.distributed.all_reduce(gbuf[lo:hi], group=group)
                main.wait_stream(side)
            self.step_count += 1
            check(lib().ddp_clip_adamw_step_dev(ptr(self.flat), ptr(grads), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                                n, ptr(self._step_dev), self.lr, self.betas[0],
                                                self.betas[1], self.eps, self.weight_decay, self.max_grad_norm,
                                                ptr(self._norm), ptr(self._scratch), stream_ptr()),
                  "ddp_clip_adamw_step_dev")
        actor.mark_dirty()
        return loss[0], self._norm[0]

    def _step_graph(self, state, action, noise, timesteps, global_batch):
        dev = self.flat.device
        key = (tuple(state.shape), tuple(action.shape), int(global_batch), self.precision or self.actor.train_precision)
        ent = self._graphs.get(key)
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32)
        if ent is None:
            # first step of this shape runs eagerly (sizes the workspaces and the weight pack); the static inputs
            # the graph will read are created here
            ent = {"state": f32(state).clone(), "action": f32(action).clone(), "noise": f32(noise).clone(),
                   "ts": timesteps.detach().to(device=dev, dtype=torch.int64).clone(), "graph": None}
            self._graphs[key] = ent
            self._buffers(action.shape[0], ent)
            loss, norm = self._step_body(ent["state"], ent["action"], ent["noise"], ent["ts"], global_batch, ent)
            return loss.clone(), norm.clone()
        ent["state"].copy_(state); ent["action"].copy_(action); ent["noise"].copy_(noise); ent["ts"].copy_(timesteps)
        if ent["graph"] is None:
            self.actor.mark_dirty()                  # the capture must contain the weight re-pack
            g = torch.cuda.CUDAGraph()
            count = self.step_count
            with torch.cuda.graph(g):
                ent["loss"], ent["norm"] = self._step_body(ent["state"], ent["action"], ent["noise"], ent["ts"],
                                                           global_batch, ent)
            self.step_count = count                  # capturing records, it does not execute
            ent["graph"] = g
        ent["graph"].replay()
        self.step_count += 1
        self.actor.mark_dirty()
        return ent["loss"].clone(), ent["norm"].clone()


class _FlatTrainer:
    """Shared machinery of the fused trainers of the callers' networks (critic, RND predictor): the parameters re-pointed
    at slices of one flat fp32 buffer (state_dict keys, shapes and values unchanged), a loss / flat-gradient call that
    writes into a persistent ``[gradient | loss]`` buffer, the data-parallel average (one all-reduce, the loss in the same
    buffer), clip + AdamW on the flat vector (``ddp_clip_adamw_step_dev``: the step count lives on the device), and the
    whole update replayed as one CUDA graph per batch shape (``graph=True``).  Subclasses provide ``_loss_into`` and
    ``_dirty``."""

    def __init__(self, params, lr, betas, eps, weight_decay, max_grad_norm, process_group, graph):
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = float("inf") if max_grad_norm is None else max_grad_norm
        self.group, self.use_graph, self._graphs = process_group, graph, {}
        params = list(params)
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        off = 0
        for p in params:
            self.flat[off:off + p.numel()].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self._norm = torch.zeros(1, device=dev)
        self._scratch = torch.zeros(640, device=dev)          # DDP_ADAMW_SCRATCH_FLOATS
        self._step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.step_count = 0
        self._dirty()

    def world_size(self):
        if self.group is False:
            return 1
        if self.group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return torch.distributed.get_world_size(self.group)
        return 1

    def close(self):
        """Release the captured graphs (they hold the collective) before ``destroy_process_group()``."""
        torch.cuda.synchronize(self.flat.device)
        self._graphs.clear()
        torch.cuda.synchronize(self.flat.device)

    def _body(self, batch, extra, holder):
        n = self.flat.numel()
        dev = self.flat.device
        if holder.get("gbuf") is None:
            holder["gbuf"] = torch.zeros(n + 4, device=dev, dtype=torch.float32)        # [gradient | loss | pad]
        gbuf = holder["gbuf"]
        self._dirty()                     # the flat vector is written by raw pointer: the weight pack is redone every step
        self._loss_into(batch, extra, gbuf[n:n + 1], gbuf[:n], holder)
        world = self.world_size()
        if world > 1:
            torch.distributed.all_reduce(gbuf[:n + 1], group=None if self.group in (None, False) else self.group)
            gbuf[:n + 1].mul_(1.0 / world)
        self.step_count += 1
        with torch.cuda.device(dev):
            check(lib().ddp_clip_adamw_step_dev(ptr(self.flat), ptr(gbuf), ptr(self.exp_avg), ptr(self.exp_avg_sq), n,
                                                ptr(self._step_dev), self.lr, self.betas[0], self.betas[1], self.eps,
                                                self.weight_decay, self.max_grad_norm, ptr(self._norm),
                                                ptr(self._scratch), stream_ptr()), "ddp_clip_adamw_step_dev")
        self._dirty()
        return gbuf[n], self._norm[0]

    def _run(self, batch, extra):
        """(loss, pre-clip grad norm) as 0-dim device tensors; ``extra``: hashable call constants (part of the graph key)."""
        if not self.use_graph:
            loss, norm = self._body(batch, extra, self.__dict__.setdefault("_eager", {}))
            return loss.clone(), norm.clone()
        dev = self.flat.device
        key = (tuple(tuple(t.shape) for t in batch), extra)
        ent = self._graphs.get(key)
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32)
        if ent is None:       # first step of this shape runs eagerly (sizes workspaces); the graph's static inputs are made here
            ent = {"batch": tuple(f32(t).clone() for t in batch), "graph": None}
            self._graphs[key] = ent
            loss, norm = self._body(ent["batch"], extra, ent)
            return loss.clone(), norm.clone()
        for dst, src in zip(ent["batch"], batch):
            dst.copy_(src)
        if ent["graph"] is None:
            g = torch.cuda.CUDAGraph()
            count = self.step_count
            with torch.cuda.graph(g):
                ent["loss"], ent["norm"] = self._body(ent["batch"], extra, ent)
            self.step_count = count
            ent["graph"] = g
        ent["graph"].replay()
        self.step_count += 1
        self._dirty()
        return ent["loss"].clone(), ent["norm"].clone()


class FusedCriticTrainer(_FlatTrainer):
    """Whole ``update_critic`` for one critic on the device: fused target / loss / backward (one C-ABI call), the
    data-parallel gradient average, then clip + AdamW on a flat parameter vector instead of ``clip_grad_norm_`` +
    ``torch.optim.AdamW.step`` (ddiffpg.py:322-351, ac_base.py:83-92; hyper-parameters default to the reference's
    ``AdamW(critic.parameters(), critic_lr)`` and ``max_grad_norm`` 1.0).  ``graph=True`` captures the whole update once per
    batch shape into a CUDA graph -- the launch-bound regime of the reference's 4 096-row updates."""

    def __init__(self, critic, critic_target, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=1.0,
                 process_group=None, precision=None, graph=False):
        self.critic, self.target = critic, critic_target
        self.precision = precision or getattr(critic, "train_precision", "fp32")
        super().__init__(critic.parameters(), lr, betas, eps, weight_decay, max_grad_norm, process_group, graph)

    def _dirty(self):
        self.critic.mark_dirty()

    def _loss_into(self, batch, gamma_n, loss, grads, holder):
        obs, action, reward, next_obs, next_actions, done = batch
        _critic_loss_into(self.critic, self.target, obs, action, next_obs, next_actions, reward, done, gamma_n,
                          self.precision, loss, grads, holder)

    def step(self, obs, action, reward, next_obs, next_actions, done, gamma_n=0.99):
        """One update (argument order of ``update_critic``); returns (loss, pre-clip grad norm) as 0-dim device tensors."""
        return self._run((obs, action, reward, next_obs, next_actions, done), float(gamma_n))


class FusedRNDTrainer(_FlatTrainer):
    """``IntrinsicM.update`` (utils/intrinsic.py:67-83) on the device: mse loss + predictor backward (one C-ABI call),
    clip at 1.0 + AdamW(lr 1e-4) on the flat predictor vector, optionally one CUDA graph per batch shape."""

    def __init__(self, rnd_model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=1.0,
                 process_group=False, graph=False):
        self.model = rnd_model
        super().__init__(rnd_model.predictor.parameters(), lr, betas, eps, weight_decay, max_grad_norm, process_group, graph)

    def _dirty(self):
        self.model.mark_dirty()

    def _loss_into(self, batch, extra, loss, grads, holder):
        self.model.loss_and_grads_into(batch[0], loss, grads)

    def step(self, x):
        """One predictor update on the (already encoded) states ``x``; returns (loss, pre-clip grad norm)."""
        return self._run((x,), None)


class HotPathMixin:
    """Mix into an agent class (before ``ActorCriticBase``) to route its hot-path methods through the
    kernels while keeping the reference's signatures; reads the same cfg keys
    (cfg.diffusion.action_lr / update_times, cfg.algo.max_grad_norm)."""

    # ``noise`` / ``expl_noise`` (keyword-only, not in the reference) inject the Gaussian draws for reproducible runs
    def get_actions(self, obs, sample=True, *, noise=None, expl_noise=None):
        if self.cfg.algo.obs_norm:
            obs = self.obs_rms.normalize(obs)
        n = self.cfg.algo.noise
        return get_actions(self.actor, obs, sample=sample, noise_type=n.type, std_min=n.get("std_min", 0.0),
                           std_max=n.get("std_max", 0.0), std=self.get_noise_std() if n.type == "fixed" else None,
                           noise=noise, expl_noise=expl_noise)

    def get_tgt_policy_actions(self, obs, sample=True, *, noise=None, expl_noise=None):
        n = self.cfg.algo.noise
        return get_tgt_policy_actions(self.actor_target, obs, sample=sample, tgt_pol_std=n.tgt_pol_std,
                                      tgt_pol_noise_bound=n.tgt_pol_noise_bound, noise=noise, expl_noise=expl_noise)

    def update_critic(self, critic, critic_target, critic_optimizer, obs, action, reward, next_obs,
                      embedded_next_obs, done, *, noise=None, expl_noise=None):
        next_actions = self.get_tgt_policy_actions(embedded_next_obs, noise=noise, expl_noise=expl_noise)
        return update_critic(critic, critic_target, critic_optimizer, obs, action, reward, next_obs, next_actions,
                             done, gamma_n=self.cfg.algo.gamma ** self.cfg.algo.nstep,
                             max_grad_norm=self.cfg.algo.max_grad_norm)

    def update_target_action(self, obs, action, critic):
        return update_target_action(obs, action, critic, action_lr=self.cfg.diffusion.action_lr,
                                    update_times=self.cfg.diffusion.update_times,
                                    max_grad_norm=self.cfg.algo.max_grad_norm)

    def optimizer_update(self, optimizer, objective):
        return optimizer_update(optimizer, objective, self.cfg.algo.max_grad_norm)

    def update_actor(self, obs, target_action):
        return update_actor(self.actor, self.actor_optimizer, obs, target_action, self.cfg.algo.max_grad_norm)
