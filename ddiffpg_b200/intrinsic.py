"""RND / NovelD intrinsic reward on the accelerated path.

``RNDModel`` mirrors ``ddiffpg.models.mlp.RNDModel`` (:233-267) as the parameter container the kernels read;
``IntrinsicKernels`` is a mixin for ``ddiffpg.utils.intrinsic.IntrinsicM`` (:8-94) that overrides only the methods
that run the networks: the two MLPs, the novelty norm, the mse loss and the predictor's backward go through
``libddiffpg_b200.so`` (``RNDModel.precision``: "fp32" FMA tile kernel, or "bf16" tcgen05 GEMMs for the update-batch
sizes), the optimizer step is torch's own AdamW as in the reference, and everything else
(reward shaping, positional encoding, running statistics) remains the reference's code.
"""
from collections.abc import Sequence

import numpy as np
import torch
import torch.nn as nn
from torch.nn.utils import clip_grad_norm_

from . import _lib
from ._lib import RndShape, check, lib, ptr, ptr_array, stream_ptr
from .models import _PackCache


class RNDModel(nn.Module):
    def __init__(self, state_dim, hidden=(512, 256, 128), feature_dim=128, precision="fp32"):
        super().__init__()
        self.precision = precision      # "fp32" | "bf16" (tensor path: D <= 256, hidden widths multiples of 64)
        if isinstance(state_dim, Sequence):
            state_dim = state_dim[0]
        self.state_dim, self.hidden, self.feature_dim = int(state_dim), tuple(hidden), int(feature_dim)
        h1, h2, h3 = self.hidden

        def mlp():
            return nn.Sequential(nn.Linear(self.state_dim, h1), nn.ELU(), nn.Linear(h1, h2), nn.ELU(),
                                 nn.Linear(h2, h3), nn.ELU(), nn.Linear(h3, self.feature_dim))
        self.predictor = mlp()
        self.target = mlp()
        for p in self.modules():
            if isinstance(p, nn.Linear):
                nn.init.orthogonal_(p.weight, np.sqrt(2))
                p.bias.data.zero_()
        for param in self.target.parameters():
            param.requires_grad = False
        self._cache = _PackCache()
        self._ws = None

    def mark_dirty(self):
        self._cache.dirty = True

    def _packed(self):
        params = [p for _, p in self.named_parameters()]
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("ddiffpg_b200.RNDModel runs on CUDA only (no CPU fallback)")
        shape = RndShape(self.state_dim, self.feature_dim, *self.hidden)
        prec = _lib.PRECISIONS[self.precision]
        if self._cache.stale(params, ("rnd", self.precision, str(dev))):
            nbytes = lib().ddp_rnd_packed_bytes_p(shape, prec)
            if nbytes == 0:
                check(-1, "ddp_rnd_packed_bytes_p")
            if self._cache.buf is None or self._cache.buf.numel() != nbytes or self._cache.buf.device != dev:
                self._cache.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                check(lib().ddp_rnd_pack_p(shape, ptr_array([p.detach() for p in params]), ptr(self._cache.buf), prec,
                                           stream_ptr()), "ddp_rnd_pack_p")
            self._cache.dirty = False
        return self._cache.buf, shape, prec

    def _workspace(self, shape, B, prec, dev):
        ws_bytes = lib().ddp_rnd_workspace_bytes_p(shape, B, prec)
        if self._ws is None or self._ws.numel() < ws_bytes or self._ws.device != dev:
            self._ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=dev)
        return self._ws, ws_bytes

    def _x(self, state, dev):
        x = state.detach().to(device=dev, dtype=torch.float32).contiguous()
        if x.dim() != 2 or x.shape[1] != self.state_dim:
            raise ValueError(f"expected state [B,{self.state_dim}], got {tuple(x.shape)}")
        return x

    @torch.no_grad()
    def features(self, state, want_features=True):
        """(novelty [B], predict_feature [B,F], target_feature [B,F]) from one launch."""
        packed, shape, prec = self._packed()
        x = self._x(state, packed.device)
        B = x.shape[0]
        nov = torch.empty(B, device=x.device)
        pf = torch.empty(B, self.feature_dim, device=x.device) if want_features else None
        tf = torch.empty(B, self.feature_dim, device=x.device) if want_features else None
        if B:
            with torch.cuda.device(x.device):
                ws, ws_bytes = self._workspace(shape, B, prec, x.device)
                check(lib().ddp_rnd_novelty_p(shape, ptr(packed), ptr(x), ptr(nov), ptr(pf), ptr(tf), B, prec, ptr(ws),
                                              ws_bytes, stream_ptr()), "ddp_rnd_novelty_p")
        return nov, pf, tf

    def forward(self, state):
        """mlp.py:262-266 (forward values only: the predictor's gradient comes from ``loss_and_grads``)."""
        _, pf, tf = self.features(state)
        return pf, tf

    def novelty(self, state):
        return self.features(state, want_features=False)[0]

    def loss_and_grads_into(self, state, loss, grads):
        """mse_loss(predictor(x), target(x)) into ``loss`` (1 float, zeroed here) and its flat gradient w.r.t. the predictor
        parameters into ``grads``, both written in place (persistent buffers of a fused trainer / captured graph)."""
        packed, shape, prec = self._packed()
        x = self._x(state, packed.device)
        B = x.shape[0]
        if grads.numel() != lib().ddp_rnd_grad_count(shape):
            raise ValueError("gradient buffer does not match the predictor")
        loss.zero_()
        with torch.cuda.device(x.device):
            ws, ws_bytes = self._workspace(shape, B, prec, x.device)
            check(lib().ddp_rnd_loss_fwd_bwd_p(shape, ptr(packed), ptr(x), ptr(loss), ptr(grads), None, B, prec, ptr(ws),
                                               ws_bytes, stream_ptr()), "ddp_rnd_loss_fwd_bwd_p")

    def loss_and_grads(self, state):
        """mse_loss(predictor(x), target(x)) and its flat gradient w.r.t. the predictor parameters."""
        dev = next(self.parameters()).device
        n = sum(p.numel() for p in self.predictor.parameters())
        grads = torch.empty(n, device=dev)
        loss = torch.zeros(1, device=dev)
        self.loss_and_grads_into(state, loss, grads)
        return loss[0], grads


class IntrinsicKernels:
    """Mixin for ``ddiffpg.utils.intrinsic.IntrinsicM`` (:8-94): the three methods that run the two networks --
    ``get_novelty`` (:62-65), ``update`` (:67-75) and its ``optimizer_update`` (:77-83) -- go through the kernels;
    construction, ``compute_reward``, ``encode_obs`` and the running statistics stay the reference's own code.
    ``accelerate_intrinsic(IntrinsicM)`` builds the class; it swaps ``rnd_model`` for the kernel-backed ``RNDModel``
    (same weights) and re-creates the AdamW over its predictor."""

    def _adopt_rnd_model(self):
        old = self.rnd_model
        if isinstance(old, RNDModel):
            return
        first = old.predictor[0]
        new = RNDModel(first.in_features, hidden=tuple(l.out_features for l in list(old.predictor)[0:5:2]),
                       feature_dim=old.predictor[-1].out_features)
        new.load_state_dict(old.state_dict())
        self.rnd_model = new.to(first.weight.device)
        self.rnd_optimizer = torch.optim.AdamW(self.rnd_model.parameters(), 1e-4)

    def get_novelty(self, obs):
        return self.rnd_model.novelty(obs)

    def enable_fused_update(self, graph=True):
        """Route ``update`` through ``FusedRNDTrainer`` (flat-vector clip + AdamW on the device, one CUDA graph per batch
        shape) instead of ``rnd_optimizer``; same hyper-parameters (AdamW 1e-4, clip 1.0)."""
        from .algo import FusedRNDTrainer
        self.rnd_trainer = FusedRNDTrainer(self.rnd_model, lr=1e-4, graph=graph)
        return self.rnd_trainer

    def update(self, obs):
        if self.pos_enc:
            obs = self.encode_obs(obs)
        if getattr(self, "rnd_trainer", None) is not None:
            dynamic_loss, dynamic_grad_norm = self.rnd_trainer.step(obs)
            self.update_step += 1
            return dynamic_loss.item(), dynamic_grad_norm.item()
        dynamic_loss, grads = self.rnd_model.loss_and_grads(obs)
        dynamic_grad_norm = self.optimizer_update(self.rnd_optimizer, (dynamic_loss, grads))
        self.update_step += 1
        return dynamic_loss.item(), dynamic_grad_norm.item()

    def optimizer_update(self, optimizer, objective):
        """The backward is already done by the kernel: scatter the flat gradient, clip at 1.0, step."""
        _, grads = objective
        optimizer.zero_grad(set_to_none=True)
        off = 0
        for p in self.rnd_model.predictor.parameters():
            p.grad = grads[off:off + p.numel()].view(p.shape)
            off += p.numel()
        grad_norm = clip_grad_norm_(parameters=optimizer.param_groups[0]["params"], max_norm=1.0)
        optimizer.step()
        return grad_norm


def accelerate_intrinsic(base):
    """``IntrinsicKernels`` in front of ``base`` (the reference's ``IntrinsicM``)."""
    def __init__(self, *args, **kwargs):
        base.__init__(self, *args, **kwargs)
        self._adopt_rnd_model()
    return type(base.__name__, (IntrinsicKernels, base), {"__init__": __init__, "__doc__": IntrinsicKernels.__doc__})


try:        # needs the reference package with its own dependencies (gym, ...) importable
    from ddiffpg.utils.intrinsic import IntrinsicM as _ReferenceIntrinsicM
    IntrinsicM = accelerate_intrinsic(_ReferenceIntrinsicM)
except Exception:
    IntrinsicM = None       # build it with accelerate_intrinsic(<the reference class>) where the reference is installed
