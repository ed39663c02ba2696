"""Clip-parallel multi-GPU driver: clips are independent end to end, weights are replicated, rank r owns
clips r, r+W, r+2W, ...; the only exchange of the path is ONE all-gather of the last-position logits (or final hidden
state) per batch, over NCCL/NVLink on GPUs (gloo in the CPU tests).  The reference runs inference on a single GPU
(`mv_bench.py:241`), so this layer is new; it adds no arithmetic."""
import torch
import torch.distributed as dist


def shard_clips(n_clips, rank, world):
    """indices of the clips rank `rank` processes (round-robin, so any prefix of the batch is balanced)"""
    return list(range(rank, n_clips, world))


def owner_of(clip, world):
    return clip % world


def gather_clip_outputs(local, n_clips, group=None):
    """local: [n_local, ...] outputs of this rank's clips (in shard_clips order) -> [n_clips, ...] on every rank.
    Ranks may own different numbers of clips (n_clips % world != 0): shorter shards are padded for the collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    per = (n_clips + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.view((world, per) + tuple(local.shape[1:]))
    # un-interleave: clip c lives at [c % world, c // world]
    idx_r = torch.tensor([c % world for c in range(n_clips)], device=local.device)
    idx_i = torch.tensor([c // world for c in range(n_clips)], device=local.device)
    return out[idx_r, idx_i]
