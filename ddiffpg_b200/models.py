"""Host-side mirror of the reference model layer for the hot path.

Same class names, constructor arguments, method names, ``state_dict()`` keys and error behaviour as
``ddiffpg/models/diffusion_mlp.py`` (``DiffusionPolicy``) and ``ddiffpg/models/mlp.py``
(``DistributionalDoubleQ``), so ``ddiffpg/algo/ac_base.py:29-31`` and ``ddiffpg/utils/Q_scheduler.py:16-23``
can construct them unchanged.  Parameters stay fp32 ``nn.Parameter``s (master weights: AdamW, deepcopy and
soft_update mutate them); every forward pass goes through the C ABI (``include/ddiffpg_b200.h``).
There is no eager or CPU fallback for the accelerated calls.
"""
import math
from collections.abc import Sequence

import torch
import torch.nn as nn

from . import _lib
from ._lib import ActorShape, QShape, check, lib, ptr, ptr_array, stream_ptr


class SinusoidalPosEmb(nn.Module):
    """Placeholder that keeps the ``net.time_mlp.{1,3}`` key numbering of the reference
    (diffusion_mlp.py:9-21,38-43); the embedding itself is evaluated inside ``ddp_actor_pack``."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):  # pragma: no cover - never on the accelerated path
        raise RuntimeError("SinusoidalPosEmb is folded into the packed time table; call DiffusionPolicy")


class DiffusionNet(nn.Module):
    """Parameter container with the reference layout (diffusion_mlp.py:24-58)."""

    def __init__(self, transition_dim, cond_dim, dim=256, hidden=(1024, 512, 256)):
        super().__init__()
        self.time_dim = dim
        self.transition_dim = transition_dim
        self.action_dim = transition_dim - cond_dim
        h1, h2, h3 = hidden
        act = nn.Mish()
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, dim * 4), act, nn.Linear(dim * 4, dim))
        self.mlp = nn.Sequential(nn.Linear(dim + transition_dim, h1), act, nn.Linear(h1, h2), act,
                                 nn.Linear(h2, h3), act, nn.Linear(h3, self.action_dim))

    def forward(self, x, time, cond):  # pragma: no cover
        raise RuntimeError("DiffusionNet has no stand-alone forward on the accelerated path; "
                           "use DiffusionPolicy.forward / get_loss")


class _PackCache:
    """Packed weights, rebuilt when any parameter's (data_ptr, _version) changes.

    In-place updates through ``param.data`` (the reference's ``soft_update``, utils/torch_util.py:9-12) do
    not bump ``_version``.  The one network the reference updates that way, the target critic, is therefore re-packed
    on every use by ``critic_loss_and_grads`` (its only reader on the hot path); for any other ``.data`` write call
    ``mark_dirty()`` (``ddiffpg_b200.algo.soft_update`` does)."""

    def __init__(self):
        self.key = None
        self.buf = None
        self.dirty = True

    def stale(self, params, extra):
        key = (extra,) + tuple((p.data_ptr(), p._version) for p in params)
        if self.dirty or key != self.key:
            self.key = key
            return True
        return False


class _ActorLossFn(torch.autograd.Function):
    """loss = mse(eps_hat, noise); the C ABI returns the loss and all 12 gradients in one pass, so
    ``objective.backward()`` in ``optimizer_update`` (ac_base.py:83-85) only scales and scatters them."""

    @staticmethod
    def forward(ctx, policy, state, action, noise, timesteps, *params):
        loss, grads = policy._loss_and_grads(state, action, noise, timesteps)
        ctx.grads = grads
        ctx.shapes = [p.shape for p in params]
        return loss

    @staticmethod
    def backward(ctx, gout):
        flat = ctx.grads * gout
        outs, off = [], 0
        for shp in ctx.shapes:
            n = math.prod(shp)
            outs.append(flat[off:off + n].view(shp))
            off += n
        return (None, None, None, None, None, *outs)


def host_batch_ranges(B, chunks, sms):
    """Row ranges of a host-resident batch for the copy / compute pipeline of ``get_actions_host``.  Below ~8k rows per
    range the copies are microseconds and extra launches cost more than they hide: one range.  Up to three waves of the
    sampler (``sms`` tiles of 128 rows) ranges are whole waves, at most ``chunks`` of them, with the partial wave FIRST:
    the first launch then waits for the smallest upload, and no range ends in a second partial wave.  Beyond: one wave,
    then the rest."""
    chunks = max(1, min(chunks, B // 8192))
    if chunks == 1:
        return [(0, B)] if B > 0 else []
    wave = sms * 128
    if B >= 3 * wave:
        # large batches: one wave first (its upload is the only exposed one), everything else in a second launch -- the
        # sampler deals tile-steps out evenly once a launch holds two waves or more, so nothing is lost to a partial wave
        return [(0, wave), (wave, B)]
    n_waves, rem = divmod(B, wave)
    if n_waves == 0:
        return [(0, B)]
    ranges = [(0, rem)] if rem else []
    groups = max(1, min(n_waves, chunks - len(ranges)))
    lo = rem
    for g in range(groups):
        w = n_waves // groups + (1 if g < n_waves % groups else 0)
        ranges.append((lo, lo + w * wave))
        lo += w * wave
    return ranges


class DiffusionPolicy(nn.Module):
    """Drop-in for ``ddiffpg.models.diffusion_mlp.DiffusionPolicy`` (:148-321).

    Extra keyword arguments (all optional, defaults reproduce the reference): ``hidden`` trunk widths,
    ``precision`` ("fp32" FMA path, parity 1e-4; "bf16" tcgen05 path, parity 1e-2)."""

    def __init__(self, state_dim, action_dim, diffusion_iter, num_mode=0, tau1=0.4, tau2=0.9, noise_min=0.0,
                 noise_max=0.25, noise_type="mixed", psi=1.0, energy=False, device="cuda",
                 hidden=(1024, 512, 256), precision="fp32"):
        super().__init__()
        if isinstance(state_dim, Sequence):
            state_dim = state_dim[0]
        if energy:
            raise NotImplementedError("energy=True (EBMDiffusionModel) is dead code in the reference and is not "
                                      "part of the accelerated path")
        self.state_dim = state_dim
        self.action_dim = action_dim
        self.diffusion_iter = diffusion_iter
        self.device = device
        self.num_mode = num_mode
        self.hidden = tuple(hidden)
        self.precision = precision
        self.train_precision = "fp32"       # "bf16": tcgen05 GEMM path for get_loss / FusedActorTrainer
        self.net = DiffusionNet(transition_dim=state_dim + action_dim + num_mode, cond_dim=state_dim + num_mode,
                                hidden=self.hidden)
        # noise-related attributes kept for interface parity (diffusion_mlp.py:176-182)
        self.tau1, self.tau2, self.noise_min, self.noise_max = tau1, tau2, noise_min, noise_max
        self.noise_type, self.psi, self.rescale = noise_type, psi, True
        self._cache = {}
        self._ws = {}

    # ------------------------------------------------------------------ plumbing
    def _shape(self):
        h1, h2, h3 = self.hidden
        return ActorShape(self.state_dim + self.num_mode, self.action_dim, self.diffusion_iter,
                          self.net.time_dim, h1, h2, h3)

    def _params(self):
        return [p for _, p in self.net.named_parameters()]      # state_dict order

    def mark_dirty(self):
        for c in self._cache.values():
            c.dirty = True

    def _packed(self, precision, need=3):
        """Packed weights for ``precision``; ``need`` = 1 sampler, 2 training, 3 both (parts packed lazily)."""
        prec = _lib.PRECISIONS[precision]
        params = self._params()
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("ddiffpg_b200.DiffusionPolicy runs on CUDA only (no CPU fallback); "
                               "move the module with .to('cuda')")
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("parameters must be contiguous fp32 tensors")
        cache = self._cache.setdefault(precision, _PackCache())
        shape = self._shape()
        if cache.stale(params, (precision, self.diffusion_iter, str(dev))):
            cache.parts = 0
        missing = need & ~getattr(cache, "parts", 0)
        if missing:
            nbytes = lib().ddp_actor_packed_bytes(shape, prec)
            if nbytes == 0:
                check(-1, "ddp_actor_packed_bytes")
            if cache.buf is None or cache.buf.numel() != nbytes or cache.buf.device != dev:
                cache.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                missing = need
                cache.parts = 0
            with torch.cuda.device(dev):
                check(lib().ddp_actor_pack_parts(shape, ptr_array([p.detach() for p in params]), ptr(cache.buf), prec,
                                                 missing, stream_ptr()), "ddp_actor_pack_parts")
            cache.parts = getattr(cache, "parts", 0) | missing
            cache.dirty = False
        return cache.buf, shape, prec

    def _workspace(self, name, nbytes, dev):
        buf = self._ws.get(name)
        if buf is None or buf.numel() < nbytes or buf.device != dev:
            buf = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
            self._ws[name] = buf
        return buf

    # ------------------------------------------------------------------ reference call surface
    def forward(self, x, sample=True, add_noise=False):
        return self.get_actions(x, sample=sample, add_noise=add_noise)

    def get_actions(self, state, sample=True, add_noise=False, noise=None, precision=None, expl_std=None,
                    expl_noise=None, noise_bound=0.0):
        """diffusion_mlp.py:219-251.  ``noise`` ([T, B, A]; [0] = x_T, [j] = step noise at t = T-j) may be
        injected for reproducibility; otherwise it is drawn on the device.

        ``expl_std=(std_min, std_max)`` fuses the noise the agent adds right after the call
        (``add_mixed_normal_noise`` / ``add_normal_noise``, utils/noise.py:19-41) into the sampler epilogue:
        ``clamp(a + clamp(linspace(std_min, std_max, B)[r] * z, +-noise_bound), -1, 1)`` with ``z`` =
        ``expl_noise`` ([B, A] standard normal; drawn on the device when omitted)."""
        if not sample:
            raise NotImplementedError("sample=False (autograd through the chain) is unused by the DDiffPG agent "
                                      "and not provided by the fused sampler")
        if add_noise:
            raise NotImplementedError("add_noise=True (state-noise annealing) is never enabled by the reference "
                                      "agents and is not part of the fused sampler")
        precision = precision or self.precision
        state = state.detach()
        if state.dim() != 2 or state.shape[1] != self.state_dim + self.num_mode:
            raise ValueError(f"state must be [B, {self.state_dim + self.num_mode}], got {tuple(state.shape)}")
        B, T, A = state.shape[0], self.diffusion_iter, self.action_dim
        packed, shape, prec = self._packed(precision, need=1)
        dev = packed.device
        state = state.to(device=dev, dtype=torch.float32).contiguous()
        if noise is None:
            noise = torch.randn((T, B, A), device=dev, dtype=torch.float32)
        else:
            if tuple(noise.shape) != (T, B, A):
                raise ValueError(f"noise must be [T={T}, B={B}, A={A}], got {tuple(noise.shape)}")
            noise = noise.to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty((B, A), device=dev, dtype=torch.float32)
        if B == 0:
            return out
        std_min = std_max = 0.0
        if expl_std is not None:
            std_min, std_max = (expl_std, expl_std) if isinstance(expl_std, (int, float)) else expl_std
            if expl_noise is None:
                expl_noise = torch.randn((B, A), device=dev, dtype=torch.float32)
            elif tuple(expl_noise.shape) != (B, A):
                raise ValueError(f"expl_noise must be [B={B}, A={A}], got {tuple(expl_noise.shape)}")
            expl_noise = expl_noise.to(device=dev, dtype=torch.float32).contiguous()
        else:
            expl_noise = None
        with torch.cuda.device(dev):
            ws_bytes = lib().ddp_actor_sample_workspace_bytes(shape, B, prec)
            ws = self._workspace("sample", ws_bytes, dev) if ws_bytes else None
            check(lib().ddp_actor_sample_noisy(shape, ptr(packed), ptr(state), ptr(noise), ptr(expl_noise),
                                               float(std_min), float(std_max), float(noise_bound or 0.0), ptr(out),
                                               B, prec, ptr(ws), ws_bytes, stream_ptr()), "ddp_actor_sample_noisy")
        return out

    def get_actions_host(self, state_host, out_host=None, chunks=4, precision=None):
        """``actor(obs)`` for a batch that lives in (pinned) HOST memory, the situation of the reference's env
        wrappers (wrappers/d4rl_wrapper.py:21-45 copy observations up and actions down around every call).
        The batch is cut into at most ``chunks`` row ranges (``host_batch_ranges``: whole sampler waves, the partial
        wave first); the H2D copy of range i+1 and the D2H copy of range i-1
        overlap the sampler launch of range i on two side streams, so PCIe time hides behind the kernel.
        Returns ``out_host`` ([B, A] fp32, pinned; valid once the call returns)."""
        precision = precision or self.precision
        B, A, T = state_host.shape[0], self.action_dim, self.diffusion_iter
        packed, shape, prec = self._packed(precision, need=1)
        dev = packed.device
        if out_host is None:
            out_host = torch.empty((B, A), dtype=torch.float32).pin_memory()
        if B == 0:
            return out_host
        ranges = host_batch_ranges(B, chunks, torch.cuda.get_device_properties(dev).multi_processor_count)
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            if "h2d" not in self._ws:
                self._ws["h2d"], self._ws["d2h"] = torch.cuda.Stream(), torch.cuda.Stream()
            h2d, d2h = self._ws["h2d"], self._ws["d2h"]
            h2d.wait_stream(main)
            stage = self._ws.get("host_stage")
            if stage is None or stage[0] != (B, state_host.shape[1], T, str(dev)):
                # staging buffers persist across calls of the same shape (no allocator traffic in the steady state)
                stage = ((B, state_host.shape[1], T, str(dev)),
                         torch.empty((B, state_host.shape[1]), device=dev, dtype=torch.float32),
                         torch.empty((B, A), device=dev, dtype=torch.float32),
                         [torch.empty((T, hi - lo, A), device=dev, dtype=torch.float32) for lo, hi in ranges])
                self._ws["host_stage"] = stage
            _, st_dev, out_dev, noise_bufs = stage
            main.wait_stream(d2h)              # the previous call's downloads have left out_dev
            # every upload is queued first (they run back to back on the copy engine), and all noise is drawn while the
            # first one is in flight: nothing but the first range's upload sits in front of the first launch
            uploaded = []
            with torch.cuda.stream(h2d):
                for lo, hi in ranges:
                    st_dev[lo:hi].copy_(state_host[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(h2d)
                    uploaded.append(ev)
            for buf in noise_bufs:
                buf.normal_()
            for ri, (lo, hi) in enumerate(ranges):
                main.wait_event(uploaded[ri])
                n = hi - lo
                noise = noise_bufs[ri]
                ws_bytes = lib().ddp_actor_sample_workspace_bytes(shape, n, prec)
                ws = self._workspace("sample", ws_bytes, dev) if ws_bytes else None
                check(lib().ddp_actor_sample(shape, ptr(packed), ptr(st_dev[lo:hi]), ptr(noise), ptr(out_dev[lo:hi]), n,
                                             prec, ptr(ws), ws_bytes, stream_ptr()), "ddp_actor_sample")
                ev2 = torch.cuda.Event()
                ev2.record(main)
                with torch.cuda.stream(d2h):
                    d2h.wait_event(ev2)
                    out_host[lo:hi].copy_(out_dev[lo:hi], non_blocking=True)
            main.wait_stream(d2h)
            d2h.synchronize()
        return out_host

    def _loss_and_grads(self, state, action, noise, timesteps, inv_count=None, precision=None):
        packed, shape, prec = self._packed(precision or self.train_precision, need=2)
        dev = packed.device
        B = action.shape[0]
        state = state.detach().to(device=dev, dtype=torch.float32).contiguous()
        action = action.detach().to(device=dev, dtype=torch.float32).contiguous()
        noise = noise.detach().to(device=dev, dtype=torch.float32).contiguous()
        timesteps = timesteps.detach().to(device=dev, dtype=torch.int64).contiguous()
        n_grad = lib().ddp_actor_grad_count(shape)
        grads = torch.empty(n_grad, device=dev, dtype=torch.float32)
        loss = torch.zeros((), device=dev, dtype=torch.float32)
        if inv_count is None:
            inv_count = 1.0 / (B * self.action_dim)
        with torch.cuda.device(dev):
            ws_bytes = lib().ddp_actor_train_workspace_bytes(shape, B, prec)
            ws = self._workspace("train", ws_bytes, dev)
            params = self._params()
            check(lib().ddp_actor_loss_fwd_bwd(shape, ptr(packed), ptr_array([p.detach() for p in params]),
                                               ptr(state), ptr(action), ptr(noise), ptr(timesteps), inv_count,
                                               ptr(loss), ptr(grads), B, prec, ptr(ws), ws_bytes, stream_ptr()),
                  "ddp_actor_loss_fwd_bwd")
        return loss, grads

    def get_loss(self, state, action, noise=None, timesteps=None):
        """diffusion_mlp.py:294-321.  Returns a 0-dim tensor; ``.backward()`` fills ``.grad`` of the 12
        parameters exactly like the reference's autograd graph would."""
        B = action.shape[0]
        if B == 0:
            raise ValueError("get_loss needs a non-empty batch (mse_loss of an empty batch is NaN in the reference)")
        dev = self._params()[0].device
        if noise is None:
            noise = torch.randn(action.shape, device=dev, dtype=torch.float32)
        if timesteps is None:
            timesteps = torch.randint(0, self.diffusion_iter, (B,), device=dev).long()
        return _ActorLossFn.apply(self, state, action, noise, timesteps, *self._params())


# ------------------------------------------------------------------------------------------- critics
def _create_simple_mlp(in_dim, out_dim, hidden_layers):
    dims = [in_dim, *hidden_layers, out_dim]
    layers = []
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        layers.append(nn.Linear(a, b))
        if i < len(dims) - 2:
            layers.append(nn.ELU())
    return nn.Sequential(*layers)


class MLPNet(nn.Module):
    """Parameter container matching ``ddiffpg/models/mlp.py:23-35`` (keys ``net.{0,2,4,6}``)."""

    def __init__(self, in_dim, out_dim, hidden_layers=None):
        super().__init__()
        if isinstance(in_dim, Sequence):
            in_dim = in_dim[0]
        if hidden_layers is None:
            hidden_layers = [512, 256, 128]
        self.net = _create_simple_mlp(in_dim, out_dim, hidden_layers)

    def forward(self, x):
        return self.net(x)


def pack_critics(critics, cache, precision=None, force=False):
    """Pack one or more DistributionalDoubleQ modules (one per behaviour mode) into one buffer.
    ``force``: re-pack even if no parameter's ``(data_ptr, _version)`` changed -- for networks that are written through
    ``param.data`` (the reference's ``soft_update`` of the target critics, utils/torch_util.py:9-12), which leaves both
    untouched."""
    first = critics[0]
    precision = precision or getattr(first, "precision", "fp32")
    params = [p for c in critics for _, p in c.named_parameters()]
    dev = params[0].device
    if dev.type != "cuda":
        raise RuntimeError("ddiffpg_b200.DistributionalDoubleQ runs on CUDA only (no CPU fallback)")
    for p in params:
        if p.dtype != torch.float32 or not p.is_contiguous():
            raise RuntimeError("parameters must be contiguous fp32 tensors")
    shape = QShape(first.state_dim, first.act_dim, first.num_atoms, float(first.v_min), float(first.v_max),
                   len(critics), *first.hidden_layers)
    prec = _lib.PRECISIONS[precision]
    if cache.stale(params, (precision, len(critics), str(dev))) or force:
        nbytes = lib().ddp_q_packed_bytes(shape, prec)
        if nbytes == 0:
            check(-1, "ddp_q_packed_bytes")
        if cache.buf is None or cache.buf.numel() != nbytes or cache.buf.device != dev:
            cache.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().ddp_q_pack(shape, ptr_array([p.detach() for p in params]), ptr(cache.buf), prec,
                                   stream_ptr()), "ddp_q_pack")
        cache.dirty = False
    return cache.buf, shape, prec


class _QMinFn(torch.autograd.Function):
    """q_min with its gradient w.r.t. the action from the same kernel (critic weights are constants)."""

    @staticmethod
    def forward(ctx, critic, obs, action):
        need = action.requires_grad
        q, _, _, dq = critic._forward_raw(obs, action, want_probs=False, want_grad=need)
        ctx.dq = dq
        return q

    @staticmethod
    def backward(ctx, gout):
        if ctx.dq is None:
            return None, None, None
        return None, None, ctx.dq * gout.unsqueeze(1)


class DistributionalDoubleQ(nn.Module):
    """Drop-in for ``ddiffpg.models.mlp.DistributionalDoubleQ`` (:131-155)."""

    def __init__(self, state_dim, act_dim, v_min=-10, v_max=10, num_atoms=51, device="cuda", hidden_layers=None,
                 precision="fp32"):
        super().__init__()
        self.precision = precision      # "fp32" FMA path | "bf16" tcgen05 GEMM path (large batches)
        self.train_precision = "fp32"   # precision of update_critic's loss/backward ("bf16": tcgen05 GEMMs)
        if isinstance(state_dim, Sequence):
            state_dim = state_dim[0]
        self.device = device
        self.state_dim, self.act_dim, self.num_atoms = state_dim, act_dim, num_atoms
        self.hidden_layers = list(hidden_layers) if hidden_layers is not None else [512, 256, 128]
        self.net_q1 = MLPNet(in_dim=state_dim + act_dim, out_dim=num_atoms, hidden_layers=self.hidden_layers)
        self.net_q2 = MLPNet(in_dim=state_dim + act_dim, out_dim=num_atoms, hidden_layers=self.hidden_layers)
        self.v_min = v_min
        self.v_max = v_max
        self.z_atoms = torch.linspace(v_min, v_max, num_atoms, device=device if torch.cuda.is_available() else "cpu")
        self._cache = _PackCache()

    def mark_dirty(self):
        self._cache.dirty = True
        for name in ("_cache_fp32", "_cache_bf16"):
            if hasattr(self, name):
                getattr(self, name).dirty = True

    def _forward_raw(self, obs, action, want_probs, want_grad):
        packed, shape, prec = pack_critics([self], self._cache)
        dev = packed.device
        obs = obs.detach().to(device=dev, dtype=torch.float32).contiguous()
        action = action.detach().to(device=dev, dtype=torch.float32).contiguous()
        B = obs.shape[0]
        if obs.shape != (B, self.state_dim) or action.shape != (B, self.act_dim):
            raise ValueError(f"expected obs [B,{self.state_dim}] and action [B,{self.act_dim}], got "
                             f"{tuple(obs.shape)} and {tuple(action.shape)}")
        q = torch.empty(B, device=dev)
        p1 = torch.empty(B, self.num_atoms, device=dev) if want_probs else None
        p2 = torch.empty(B, self.num_atoms, device=dev) if want_probs else None
        dq = torch.empty(B, self.act_dim, device=dev) if want_grad else None
        if B:
            with torch.cuda.device(dev):
                ws_bytes = lib().ddp_q_forward_workspace_bytes(shape, B, prec)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
                check(lib().ddp_q_forward(shape, ptr(packed), _lib.i64_array([0, B]), ptr(obs), ptr(action), ptr(q),
                                          ptr(p1), ptr(p2), ptr(dq), B, prec, ptr(ws), ws_bytes, stream_ptr()),
                      "ddp_q_forward")
        return q, p1, p2, dq

    def _params_need_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def get_q_min(self, state, action):
        """mlp.py:143-147; differentiable w.r.t. ``action`` (the Q-ascent use, ddiffpg.py:365)."""
        if self._params_need_grad():
            # gradient w.r.t. the critic weights through q_min is never requested by the reference
            raise NotImplementedError("get_q_min with trainable critic weights: freeze the critic "
                                      "(critic.requires_grad_(False)) as update_target_action does")
        return _QMinFn.apply(self, state, action)

    def get_q1_q2(self, state, action):
        """mlp.py:149-151, forward values through the kernel (no autograd graph: the probabilities are constants).
        The reference's one caller that differentiates them w.r.t. the critic weights is the critic update
        (ddiffpg.py:348-349); on this path that is ``update_critic`` -> ``ddp_q_critic_loss_fwd_bwd``, which returns
        the loss and all gradients from one fused pass.  Asking for an autograd graph here raises instead of silently
        running eager torch."""
        if self._params_need_grad():
            raise NotImplementedError(
                "get_q1_q2 under autograd with trainable critic weights has no kernel: use ddiffpg_b200.update_critic / "
                "critic_loss_and_grads (fused BCE loss + backward), or call it under torch.no_grad() / with the critic "
                "frozen for forward values")
        _, p1, p2, _ = self._forward_raw(state, action, want_probs=True, want_grad=False)
        return p1, p2

    def get_q1(self, state, action):
        return self.get_q1_q2(state, action)[0]
