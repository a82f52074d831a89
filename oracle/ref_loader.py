"""Import the reference's own hot-path modules by file path (TEST INFRASTRUCTURE).

Only usable where ``/root/reference`` is mounted (the build container); the GPU box does not have
it, so nothing on the ``-m gpu`` / smoke / bench paths may call this.  ``import ddiffpg.models``
fails here (gym / wandb / omegaconf missing), hence by-path loading with the
restated scheduler placed at ``diffusers.schedulers.scheduling_ddpm``.
"""
import importlib.util
import os
import sys
import types

from . import ddpm

REF_ROOT = os.environ.get("DDIFFPG_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "ddiffpg/models/diffusion_mlp.py"))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns a namespace with the reference's DiffusionPolicy, DistributionalDoubleQ and the
    in-tree DDPM (``Diffusion`` + ``cosine_beta_schedule``)."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if "diffusers.schedulers.scheduling_ddpm" not in sys.modules:
        for name in ("diffusers", "diffusers.schedulers"):
            sys.modules.setdefault(name, types.ModuleType(name))
        shim = types.ModuleType("diffusers.schedulers.scheduling_ddpm")
        shim.DDPMScheduler = ddpm.DDPMSchedulerRestated
        sys.modules["diffusers.schedulers.scheduling_ddpm"] = shim
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)          # for ddiffpg.utils.torch_util (imports cleanly)
    dm = _load("_ref_diffusion_mlp", "ddiffpg/models/diffusion_mlp.py")
    mlp = _load("_ref_mlp", "ddiffpg/models/mlp.py")
    # baseline_models does `from ddiffpg.models.baseline_helpers import ...`; ddiffpg.models's
    # own __init__ pulls gym, so register an empty shim package first.
    import ddiffpg  # noqa: F401  (empty __init__)
    if "ddiffpg.models" not in sys.modules:
        pkg = types.ModuleType("ddiffpg.models")
        pkg.__path__ = [os.path.join(REF_ROOT, "ddiffpg/models")]
        sys.modules["ddiffpg.models"] = pkg
    helpers = _load("ddiffpg.models.baseline_helpers", "ddiffpg/models/baseline_helpers.py")
    bm = _load("ddiffpg.models.baseline_models", "ddiffpg/models/baseline_models.py")
    return types.SimpleNamespace(DiffusionPolicy=dm.DiffusionPolicy, DiffusionNet=dm.DiffusionNet,
                                 SinusoidalPosEmb=dm.SinusoidalPosEmb,
                                 DistributionalDoubleQ=mlp.DistributionalDoubleQ, MLPNet=mlp.MLPNet,
                                 Diffusion=bm.Diffusion,
                                 cosine_beta_schedule=helpers.cosine_beta_schedule)
