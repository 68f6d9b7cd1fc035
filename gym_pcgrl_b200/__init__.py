"""pcgrl-b200: B200-native batched PCGRL environment (see README.md / DESIGN.md).

Public surface (mirrors gym_pcgrl): ``PcgrlEnv`` (single env, classic gym API), ``BatchedPcgrlEnv``
(N lock-step envs, CUDA tensors), ``PROBLEMS`` / ``REPRESENTATIONS`` registries and ``make(id)`` for
the ``"{problem}-{representation}-v0"`` ids (gym_pcgrl/__init__.py:6-12).
"""
from .async_env import AsyncGroupedEnv
from .envs.pcgrl_env import BatchedPcgrlEnv, HostRolloutIO, HostStepIO, PcgrlEnv
from .envs.probs import PROBLEMS
from .envs.reps import REPRESENTATIONS

__version__ = "0.1.0"

# same id scheme as the reference's gym registration
REGISTRY = {"%s-%s-v0" % (p, r): {"prob": p, "rep": r} for p in PROBLEMS for r in REPRESENTATIONS}


def make(env_id, num_envs=None, **kwargs):
    """``make("binary-narrow-v0")`` -> PcgrlEnv;  ``make(id, num_envs=4096)`` -> BatchedPcgrlEnv."""
    spec = REGISTRY[env_id]
    if num_envs is None:
        return PcgrlEnv(spec["prob"], spec["rep"], **kwargs)
    return BatchedPcgrlEnv(spec["prob"], spec["rep"], num_envs=num_envs, **kwargs)


__all__ = ["PcgrlEnv", "BatchedPcgrlEnv", "AsyncGroupedEnv", "HostStepIO", "HostRolloutIO", "PROBLEMS", "REPRESENTATIONS", "REGISTRY", "make"]
