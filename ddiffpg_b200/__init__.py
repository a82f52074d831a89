"""ddiffpg_b200 -- B200-native (sm_100a) hot path of DDiffPG behind the reference's call surface.

``DiffusionPolicy`` / ``DistributionalDoubleQ`` mirror ``ddiffpg/models``; ``update_target_action``,
``optimizer_update`` and ``update_actor`` mirror ``ddiffpg/algo``.  Everything computes through
``libddiffpg_b200.so`` (C ABI: ``include/ddiffpg_b200.h``); importing works without a GPU, calling does not.
"""
from .models import DiffusionPolicy, DistributionalDoubleQ, DiffusionNet, MLPNet  # noqa: F401
from .algo import (FusedActorTrainer, FusedCriticTrainer, FusedRNDTrainer, HotPathMixin, critic_loss_and_grads, get_actions,  # noqa: F401
                   get_tgt_policy_actions, update_critic,
                   optimizer_update, q_action_ascent_segments, soft_update, update_actor, update_target_action)
from .intrinsic import IntrinsicKernels, IntrinsicM, RNDModel, accelerate_intrinsic  # noqa: F401
from .replay import (DiffusionGoalBuffer, DiffusionReplayBuffer, GoalBufferKernels, ReplayKernels,  # noqa: F401
                     accelerate_goal_buffer, accelerate_replay_buffer, add_embedding)

__all__ = ["IntrinsicKernels", "IntrinsicM", "RNDModel", "accelerate_intrinsic", "DiffusionReplayBuffer", "ReplayKernels", "DiffusionGoalBuffer",
           "GoalBufferKernels", "accelerate_goal_buffer",
           "accelerate_replay_buffer", "add_embedding", "DiffusionPolicy", "DistributionalDoubleQ", "DiffusionNet", "MLPNet", "FusedActorTrainer", "FusedCriticTrainer", "FusedRNDTrainer",
           "HotPathMixin", "critic_loss_and_grads", "update_critic", "get_actions", "get_tgt_policy_actions", "optimizer_update", "q_action_ascent_segments", "soft_update", "update_actor",
           "update_target_action"]
