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


# ---------------------------------------------------------------------------------------------------------------
# Long-video (global-local, video_input == 'residual') path with the FRAMES of one clip sharded over the ranks
# (BASELINE config 4: 64 frames, 4 GPUs, 16 frames each).  Frames are independent through ViT + Q-Former + llama_proj
# (everything before st_llm.py:463), so each rank encodes its contiguous block of frames and ONE all-gather of the
# per-frame tokens ([T,32,4096] fp16 = 16 MB at T=64) gives every rank what `aggregate_video` needs (the R selected
# local frames and the mean over all T frames).  The merge and the LLM prefill then run replicated (or on rank 0).
# ---------------------------------------------------------------------------------------------------------------
def frame_shard(T, rank, world):
    """contiguous block of frame indices owned by `rank` (first T % world ranks get one extra frame)"""
    base, extra = divmod(T, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def encode_frames_sharded(encode_fn, frames_local, T, group=None):
    """encode_fn(frames [t,3,H,W]) -> tokens [t,Q,D]; frames_local = this rank's block (frame_shard order).
    Returns the tokens of ALL T frames, in frame order, on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    tok = encode_fn(frames_local)
    if world == 1:
        return tok
    rank = dist.get_rank(group)
    per = (T + world - 1) // world
    pad = torch.zeros((per,) + tuple(tok.shape[1:]), dtype=tok.dtype, device=tok.device)
    pad[: tok.shape[0]] = tok
    out = torch.empty((world * per,) + tuple(tok.shape[1:]), dtype=tok.dtype, device=tok.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.view((world, per) + tuple(tok.shape[1:]))
    pieces = [out[r, : len(frame_shard(T, r, world))] for r in range(world)]
    return torch.cat(pieces, dim=0)


@torch.no_grad()
def long_video_prefill(model, frames_local, T, prompt_ids, group=None):
    """frames_local [t_loc,3,H,W] -> logits [1,S,V] of [global-local video tokens ‖ prompt] (replicated on all ranks)."""
    sm = model.get_model().stllm_model
    tokens = encode_frames_sharded(lambda f: sm.encode_img(f)[0], frames_local, T, group)       # [T,Q,D]
    vid = sm.aggregate_video(tokens.unsqueeze(0))                                                  # [1,1,R*Q,D]
    prompt = sm._embed(prompt_ids)                                                                 # [1,P,D]
    seq = torch.cat([vid[:, 0], prompt], dim=1)
    return model(samples=None, inputs_embeds=seq).logits
