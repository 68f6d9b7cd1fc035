"""pcgrl-b200: B200-native batched PCGRL environment (see README.md / DESIGN.md).

Public surface (mirrors gym_pcgrl): ``PcgrlEnv`` (single env, classic gym API), ``BatchedPcgrlEnv``
(N lock-step envs, CUDA tensors), ``PROBLEMS`` / ``REPRESENTATIONS`` registries and ``make(id)`` for
the ``"{problem}-{representation}-v0"`` ids (gym_pcgrl/__init__.py:6-12).
"""
from .async_env import AsyncGroupedEnv
from .envs.pcgrl_env import BatchedPcgrlEnv, HostRolloutIO, HostStepIO, PcgrlEnv
from .envs.plugin_env import PluginBatchedEnv, is_native
from .envs.probs import PROBLEMS
from .envs.reps import REPRESENTATIONS

__version__ = "0.1.0"

# same id scheme as the reference's gym registration
REGISTRY = {"%s-%s-v0" % (p, r): {"prob": p, "rep": r} for p in PROBLEMS for r in REPRESENTATIONS}


def register(env_id, prob, rep):
    """Register a "{problem}-{representation}-v0" style id whose problem / representation may be USER-DEFINED classes
    (the reference: add the class to PROBLEMS / REPRESENTATIONS and gym-register the id, gym_pcgrl/__init__.py:6-12).
    Names are looked up in PROBLEMS / REPRESENTATIONS; classes are added to them under their ``name`` attribute."""
    for obj, reg in ((prob, PROBLEMS), (rep, REPRESENTATIONS)):
        if isinstance(obj, type):
            key = getattr(obj, "name", None) or obj.__name__.lower()
            reg.setdefault(key, obj)
    p = prob if isinstance(prob, str) else (getattr(prob, "name", None) or prob.__name__.lower())
    r = rep if isinstance(rep, str) else (getattr(rep, "name", None) or rep.__name__.lower())
    REGISTRY[env_id] = {"prob": p, "rep": r}


def make(env_id, num_envs=None, **kwargs):
    """``make("binary-narrow-v0")`` -> PcgrlEnv;  ``make(id, num_envs=4096)`` -> BatchedPcgrlEnv.  Ids whose problem or
    representation is not one of the built-in classes come back as a ``PluginBatchedEnv`` (torch path, envs/plugin_env.py)."""
    spec = REGISTRY[env_id]
    if not is_native(PROBLEMS[spec["prob"]](), REPRESENTATIONS[spec["rep"]]()):
        return PluginBatchedEnv(spec["prob"], spec["rep"], num_envs=num_envs or 1, **kwargs)
    if num_envs is None:
        return PcgrlEnv(spec["prob"], spec["rep"], **kwargs)
    return BatchedPcgrlEnv(spec["prob"], spec["rep"], num_envs=num_envs, **kwargs)


__all__ = ["PcgrlEnv", "BatchedPcgrlEnv", "AsyncGroupedEnv", "PluginBatchedEnv", "register", "HostStepIO", "HostRolloutIO", "PROBLEMS", "REPRESENTATIONS", "REGISTRY", "make"]
