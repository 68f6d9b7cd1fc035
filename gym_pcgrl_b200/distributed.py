"""Multi-GPU plumbing: one process per GPU, env-index sharding, zero data-path collectives.

Environments never interact (the reference replicates whole envs over SubprocVecEnv workers, utils.py:64-70), so
rank g simply owns the global env indices [g*n, (g+1)*n) and steps them with no communication.  Seeds / MT19937
states are functions of the GLOBAL index (``BatchedPcgrlEnv(env_offset=...)``), hence env i follows the same
trajectory for any world size.  The only collective offered is the optional all-gather that presents one
contiguous [world*n] reward/done (or observation) tensor to a single learner.
"""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def env_offset(rank, envs_per_rank):
    """Global index of the first env owned by `rank`."""
    return int(rank) * int(envs_per_rank)


def init_process_group(backend=None, device=None):
    """torch.distributed init for `torchrun` launches (NCCL on GPUs, gloo on CPU); no-op for world size 1."""
    import torch.distributed as dist
    rank, world, _ = rank_world()
    if world == 1 or dist.is_initialized():
        return rank, world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        import torch
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kwargs = {"device_id": device} if (backend == "nccl" and device is not None) else {}
    dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world


def make_sharded_env(prob, rep, envs_per_rank, device, seed=0, **kwargs):
    """BatchedPcgrlEnv owning this rank's shard of a global batch of world*envs_per_rank envs."""
    from .envs.pcgrl_env import BatchedPcgrlEnv
    rank, world, _ = rank_world()
    return BatchedPcgrlEnv(prob, rep, num_envs=envs_per_rank, device=device, seed=seed,
                           env_offset=env_offset(rank, envs_per_rank), **kwargs)


def all_gather_outputs(*tensors):
    """One all_gather_into_tensor per tensor: [n, ...] on every rank -> [world*n, ...] (rank-major == global env
    order).  Returns the inputs unchanged for world size 1."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensors if len(tensors) > 1 else tensors[0]
    world = dist.get_world_size()
    outs = []
    for t in tensors:
        t = t.contiguous()
        view = t.view(torch.uint8) if t.dtype == torch.bool else t
        out = torch.empty((world * view.shape[0],) + tuple(view.shape[1:]), dtype=view.dtype, device=view.device)
        dist.all_gather_into_tensor(out, view)
        outs.append(out.view(torch.bool) if t.dtype == torch.bool else out)
    return tuple(outs) if len(outs) > 1 else outs[0]
