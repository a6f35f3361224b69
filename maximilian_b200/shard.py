"""Voice sharding across the GPUs of one box: the only multi-GPU structure the path has.

Voices (and STFT channels) are independent (no reference function couples two voices), so rank r
of `world` owns the contiguous voice range shard_range(V, r, world) with all of its state; nothing moves
between GPUs except the stereo mix bus: one sum all-reduce of mix[n_frames][2] fp64 (16 KiB at 1024 frames)
per block, over NCCL/NVLink on the device (or gloo on the CPU for the host-logic tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [lo, hi) of `total` items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mix(mix, group=None):
    """In-place sum of the per-rank mix bus. `mix`: torch tensor [n_frames][channels] (cuda for nccl, cpu for gloo).
    Summation order over ranks is fixed by the backend's algorithm for a given world size: results are
    run-to-run reproducible, and differ from a sequential CPU sum only by fp64 reassociation (<= 1e-12 rel)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mix, op=dist.ReduceOp.SUM, group=group)
    return mix


def max_over_ranks(value, device):
    """Device-side max of a python float over all ranks (timing: a multi-GPU step is as slow as its slowest rank)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
