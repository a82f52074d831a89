"""ctypes binding of libddiffpg_b200.so (the C ABI in include/ddiffpg_b200.h).

There is no CPU or eager fallback: if the library is missing, or a call fails, this raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_long, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DDP_LIB_PATH") or os.path.join(HERE, "libddiffpg_b200.so")   # override: A/B builds

DDP_FP32, DDP_BF16 = 0, 1
PRECISIONS = {"fp32": DDP_FP32, "bf16": DDP_BF16}


class ActorShape(Structure):
    _fields_ = [(n, c_int) for n in ("S", "A", "T", "D", "h1", "h2", "h3")]


class QShape(Structure):
    _fields_ = [("O", c_int), ("A", c_int), ("atoms", c_int), ("v_min", c_float), ("v_max", c_float),
                ("n_modes", c_int), ("hid1", c_int), ("hid2", c_int), ("hid3", c_int)]


class BatchShape(Structure):
    _fields_ = [("O", c_int), ("A", c_int), ("E", c_int), ("n_groups", c_int)]


class RndShape(Structure):
    _fields_ = [("D", c_int), ("F", c_int), ("hid1", c_int), ("hid2", c_int), ("hid3", c_int)]


# name -> (restype, argtypes); every symbol declared in include/ddiffpg_b200.h
PROTOTYPES = {
    "ddp_abi_version": (c_int, []),
    "ddp_last_error": (c_char_p, []),
    "ddp_actor_packed_bytes": (c_size_t, [POINTER(ActorShape), c_int]),
    "ddp_actor_pack": (c_int, [POINTER(ActorShape), POINTER(c_void_p), c_void_p, c_int, c_void_p]),
    "ddp_actor_pack_parts": (c_int, [POINTER(ActorShape), POINTER(c_void_p), c_void_p, c_int, c_int, c_void_p]),
    "ddp_actor_sample_workspace_bytes": (c_size_t, [POINTER(ActorShape), c_long, c_int]),
    "ddp_actor_sample": (c_int, [POINTER(ActorShape), c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int,
                                 c_void_p, c_size_t, c_void_p]),
    "ddp_actor_sample_noisy": (c_int, [POINTER(ActorShape), c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                       c_float, c_void_p, c_long, c_int, c_void_p, c_size_t, c_void_p]),
    "ddp_actor_grad_count": (c_size_t, [POINTER(ActorShape)]),
    "ddp_actor_train_workspace_bytes": (c_size_t, [POINTER(ActorShape), c_long, c_int]),
    "ddp_actor_loss_fwd_bwd": (c_int, [POINTER(ActorShape), c_void_p, POINTER(c_void_p), c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_long, c_int, c_void_p,
                                       c_size_t, c_void_p]),
    "ddp_actor_loss_fwd_bwd_ev": (c_int, [POINTER(ActorShape), c_void_p, POINTER(c_void_p), c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_long, c_int, c_void_p,
                                          c_size_t, c_void_p, POINTER(c_void_p)]),
    "ddp_clip_adamw_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_float, c_float,
                                    c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "ddp_clip_adamw_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_float, c_float,
                                        c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "ddp_q_packed_bytes": (c_size_t, [POINTER(QShape), c_int]),
    "ddp_q_pack": (c_int, [POINTER(QShape), POINTER(c_void_p), c_void_p, c_int, c_void_p]),
    "ddp_q_forward_workspace_bytes": (c_size_t, [POINTER(QShape), c_long, c_int]),
    "ddp_q_forward": (c_int, [POINTER(QShape), c_void_p, POINTER(c_int64), c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_long, c_int, c_void_p, c_size_t, c_void_p]),
    "ddp_q_ascent_workspace_bytes": (c_size_t, [POINTER(QShape), c_long, c_int]),
    "ddp_q_action_ascent": (c_int, [POINTER(QShape), c_void_p, POINTER(c_int64), POINTER(c_int64), c_void_p,
                                    c_void_p, c_int, c_float, c_float, c_float, c_float, c_float, c_float, c_void_p,
                                    c_void_p, c_long, c_int, c_void_p, c_size_t, c_void_p]),
    "ddp_q_action_ascent_sharded": (c_int, [POINTER(QShape), c_void_p, POINTER(c_int64), POINTER(c_int64), c_void_p,
                                            c_void_p, c_int, c_float, c_float, c_float, c_float, c_float, c_float,
                                            c_void_p, c_void_p, c_long, c_int, c_void_p, c_size_t, c_void_p, c_void_p,
                                            c_void_p]),
    "ddp_q_grad_count": (c_size_t, [POINTER(QShape)]),
    "ddp_q_critic_train_workspace_bytes": (c_size_t, [POINTER(QShape), c_long, c_int]),
    "ddp_q_critic_loss_fwd_bwd": (c_int, [POINTER(QShape), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_long, c_int, c_void_p,
                                          c_size_t, c_void_p]),
    "ddp_rnd_packed_bytes": (c_size_t, [POINTER(RndShape)]),
    "ddp_rnd_pack": (c_int, [POINTER(RndShape), POINTER(c_void_p), c_void_p, c_void_p]),
    "ddp_rnd_novelty": (c_int, [POINTER(RndShape), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_void_p]),
    "ddp_rnd_grad_count": (c_size_t, [POINTER(RndShape)]),
    "ddp_rnd_train_workspace_bytes": (c_size_t, [POINTER(RndShape), c_long]),
    "ddp_rnd_loss_fwd_bwd": (c_int, [POINTER(RndShape), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long,
                                     c_void_p, c_size_t, c_void_p]),
    "ddp_rnd_packed_bytes_p": (c_size_t, [POINTER(RndShape), c_int]),
    "ddp_rnd_pack_p": (c_int, [POINTER(RndShape), POINTER(c_void_p), c_void_p, c_int, c_void_p]),
    "ddp_rnd_workspace_bytes_p": (c_size_t, [POINTER(RndShape), c_long, c_int]),
    "ddp_rnd_novelty_p": (c_int, [POINTER(RndShape), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int,
                                  c_void_p, c_size_t, c_void_p]),
    "ddp_rnd_loss_fwd_bwd_p": (c_int, [POINTER(RndShape), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int,
                                       c_void_p, c_size_t, c_void_p]),
    "ddp_replay_gather": (c_int, [POINTER(BatchShape)] + [c_void_p] * 6 + [c_long] + [c_void_p] * 13 + [c_long, c_void_p]),
    "ddp_replay_scatter_target": (c_int, [POINTER(BatchShape), c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_long,
                                          c_void_p]),
}

# ddp_gsq_reduce_fn (include/ddiffpg_b200.h): int (*)(float* gsq_dev, int n_modes, void* stream, void* user)
GSQ_REDUCE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_int, c_void_p, c_void_p)

_lib = None


class DdpError(RuntimeError):
    pass


def lib():
    """Load the shared library once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m ddiffpg_b200.build` "
                "(or __graft_entry__.build()); ddiffpg_b200 has no CPU/eager fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)      # AttributeError if the ABI and the header disagree
            fn.restype, fn.argtypes = res, args
        if handle.ddp_abi_version() != 1:
            raise ImportError("libddiffpg_b200.so: ABI version mismatch")
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().ddp_last_error()
        raise DdpError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def i64_array(values):
    arr = (c_int64 * len(values))()
    for i, v in enumerate(values):
        arr[i] = int(v)
    return arr


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
