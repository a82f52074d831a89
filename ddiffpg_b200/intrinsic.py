"""RND / NovelD intrinsic reward on the accelerated path.

``RNDModel`` mirrors ``ddiffpg.models.mlp.RNDModel`` (:233-267), ``IntrinsicM`` mirrors
``ddiffpg.utils.intrinsic.IntrinsicM`` (:8-94) call for call; the two MLPs, the novelty norm, the mse loss and the
predictor's backward run through ``libddiffpg_b200.so`` (fp32 path), the optimizer step is torch's own AdamW as in
the reference.
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
    def __init__(self, state_dim, hidden=(512, 256, 128), feature_dim=128):
        super().__init__()
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
        if self._cache.stale(params, ("rnd", str(dev))):
            nbytes = lib().ddp_rnd_packed_bytes(shape)
            if nbytes == 0:
                check(-1, "ddp_rnd_packed_bytes")
            if self._cache.buf is None or self._cache.buf.numel() != nbytes or self._cache.buf.device != dev:
                self._cache.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                check(lib().ddp_rnd_pack(shape, ptr_array([p.detach() for p in params]), ptr(self._cache.buf),
                                         stream_ptr()), "ddp_rnd_pack")
            self._cache.dirty = False
        return self._cache.buf, shape

    def _x(self, state, dev):
        x = state.detach().to(device=dev, dtype=torch.float32).contiguous()
        if x.dim() != 2 or x.shape[1] != self.state_dim:
            raise ValueError(f"expected state [B,{self.state_dim}], got {tuple(x.shape)}")
        return x

    @torch.no_grad()
    def features(self, state, want_features=True):
        """(novelty [B], predict_feature [B,F], target_feature [B,F]) from one launch."""
        packed, shape = self._packed()
        x = self._x(state, packed.device)
        B = x.shape[0]
        nov = torch.empty(B, device=x.device)
        pf = torch.empty(B, self.feature_dim, device=x.device) if want_features else None
        tf = torch.empty(B, self.feature_dim, device=x.device) if want_features else None
        if B:
            with torch.cuda.device(x.device):
                check(lib().ddp_rnd_novelty(shape, ptr(packed), ptr(x), ptr(nov), ptr(pf), ptr(tf), B, stream_ptr()),
                      "ddp_rnd_novelty")
        return nov, pf, tf

    def forward(self, state):
        """mlp.py:262-266 (forward values only: the predictor's gradient comes from ``loss_and_grads``)."""
        _, pf, tf = self.features(state)
        return pf, tf

    def novelty(self, state):
        return self.features(state, want_features=False)[0]

    def loss_and_grads(self, state):
        """mse_loss(predictor(x), target(x)) and its flat gradient w.r.t. the predictor parameters."""
        packed, shape = self._packed()
        x = self._x(state, packed.device)
        B = x.shape[0]
        n = lib().ddp_rnd_grad_count(shape)
        grads = torch.empty(n, device=x.device)
        loss = torch.zeros((), device=x.device)
        with torch.cuda.device(x.device):
            ws_bytes = lib().ddp_rnd_train_workspace_bytes(shape, B)
            if self._ws is None or self._ws.numel() < ws_bytes or self._ws.device != x.device:
                self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            check(lib().ddp_rnd_loss_fwd_bwd(shape, ptr(packed), ptr(x), ptr(loss), ptr(grads), None, B, ptr(self._ws),
                                             ws_bytes, stream_ptr()), "ddp_rnd_loss_fwd_bwd")
        return loss, grads


class RunningMeanStd:
    """ddiffpg/utils/torch_util.py:99-146 (parallel-variance update), tensors on ``device``."""

    def __init__(self, epsilon=1e-4, shape=(), device="cuda"):
        self.device = device
        self.mean = torch.zeros(shape, device=device)
        self.var = torch.ones(shape, device=device)
        self.epsilon = epsilon
        self.count = epsilon

    def update(self, x):
        self.update_from_moments(x.mean(dim=0), x.var(dim=0), x.shape[0])

    def normalize(self, x):
        return (x - self.mean) / torch.sqrt(self.var + self.epsilon)

    def unnormalize(self, x):
        return x * torch.sqrt(self.var + self.epsilon) + self.mean

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_2 = self.var * self.count + batch_var * batch_count + delta ** 2 * self.count * batch_count / tot_count
        self.mean, self.var, self.count = new_mean, m_2 / tot_count, tot_count


def get_embedder(multires, input_dims=2):
    """NeRF positional encoding as configured by utils/intrinsic.py:122-171: [x, sin(2^k x), cos(2^k x)], k < multires."""
    freq_bands = 2. ** torch.linspace(0., multires - 1, steps=multires)

    def embed(x):
        outs = [x]
        for freq in freq_bands:
            outs.append(torch.sin(x * freq))
            outs.append(torch.cos(x * freq))
        return torch.cat(outs, -1)
    return embed, input_dims * (1 + 2 * multires)


class IntrinsicM:
    def __init__(self, obs_dim, type="noveld", env_name=None, normalize=True, pos_enc=True, L=10, warm_up=1000,
                 device="cuda"):
        self.obs_dim = obs_dim
        self.type = type
        self.env_name = env_name
        self.normalize = normalize
        self.device = device
        self.pos_enc = pos_enc
        self.update_step = 0
        self.warm_up = warm_up
        self.L = L
        if self.pos_enc:
            dims = 2 if "antmaze" in self.env_name else 3
            self.embedder, _ = get_embedder(self.L, input_dims=2)
            self.rnd_model = RNDModel(self.obs_dim[0] + dims * 2 * L).to(self.device)
        else:
            self.rnd_model = RNDModel(self.obs_dim).to(self.device)
        self.rnd_optimizer = torch.optim.AdamW(self.rnd_model.parameters(), 1e-4)
        self.rnd_rms = RunningMeanStd(shape=(1), device=self.device)

    def compute_reward(self, obs, next_obs=None):
        if self.pos_enc:
            obs = self.encode_obs(obs)
            if next_obs is not None:
                next_obs = self.encode_obs(next_obs)
        if self.type == "rnd":
            novelty_obs = self.get_novelty(obs)
            if self.normalize and self.update_step > self.warm_up:
                self.rnd_rms.update(novelty_obs)
                novelty_obs = self.rnd_rms.normalize(novelty_obs)
            return novelty_obs.unsqueeze(1)
        elif self.type == "noveld":
            assert next_obs is not None
            # both batches in one launch; the running statistics see them in the reference's order
            n = obs.shape[0]
            nov = self.get_novelty(torch.cat([obs, next_obs]))
            novelty_obs, novelty_nextobs = nov[:n], nov[n:]
            if self.normalize and self.update_step > self.warm_up:
                self.rnd_rms.update(novelty_obs)
                self.rnd_rms.update(novelty_nextobs)
                novelty_obs = self.rnd_rms.normalize(novelty_obs)
                novelty_nextobs = self.rnd_rms.normalize(novelty_nextobs)
            intrinsic = novelty_nextobs - 0.5 * novelty_obs
            return 0.01 * torch.max(intrinsic, torch.zeros(intrinsic.shape, device=intrinsic.device)).unsqueeze(1)
        else:
            raise NotImplementedError

    def get_novelty(self, obs):
        return self.rnd_model.novelty(obs)

    def update(self, obs):
        if self.pos_enc:
            obs = self.encode_obs(obs)
        dynamic_loss, grads = self.rnd_model.loss_and_grads(obs)
        dynamic_grad_norm = self.optimizer_update(self.rnd_optimizer, (dynamic_loss, grads))
        self.update_step += 1
        return dynamic_loss.item(), dynamic_grad_norm.item()

    def optimizer_update(self, optimizer, objective):
        """intrinsic.py:77-83 with the backward already done by the kernel: scatter, clip, step."""
        _, grads = objective
        optimizer.zero_grad(set_to_none=True)
        off = 0
        for p in self.rnd_model.predictor.parameters():
            p.grad = grads[off:off + p.numel()].view(p.shape)
            off += p.numel()
        grad_norm = clip_grad_norm_(parameters=optimizer.param_groups[0]["params"], max_norm=1.0)
        optimizer.step()
        return grad_norm

    def encode_obs(self, obs):
        k = 2 if "antmaze" in self.env_name else 3       # ant 2-d position / end-effector 3-d position
        return torch.cat([self.embedder(obs[:, :k]), obs[:, k:]], dim=1)
