"""world_size-2 gloo test of the clip-parallel host logic (the N>1 path of bench.py / SURVEY 8e) on CPU."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_clips, ret):
    sys.path.insert(0, ROOT)
    from stllm_b200.clip_parallel import gather_clip_outputs, shard_clips
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_clips(n_clips, rank, world)
    # "logits" of clip c: a vector filled with c (so ordering mistakes are visible)
    local = torch.stack([torch.full((7,), float(c)) for c in mine]) if mine else torch.zeros(0, 7)
    full = gather_clip_outputs(local, n_clips)
    ok = full.shape == (n_clips, 7) and all(bool((full[c] == c).all()) for c in range(n_clips))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def _run(n_clips, port):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_clips, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def test_clip_parallel_gather_even():
    _run(4, 29611)


def test_clip_parallel_gather_ragged():
    _run(5, 29612)


def test_shard_assignment():
    sys.path.insert(0, ROOT)
    from stllm_b200.clip_parallel import owner_of, shard_clips
    assert shard_clips(8, 3, 8) == [3] and shard_clips(5, 1, 2) == [1, 3] and shard_clips(1, 1, 2) == []
    assert [owner_of(c, 4) for c in range(6)] == [0, 1, 2, 3, 0, 1]


def _worker_frames(rank, world, port, T, ret):
    sys.path.insert(0, ROOT)
    from stllm_b200.clip_parallel import encode_frames_sharded, frame_shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = torch.arange(T, dtype=torch.float32).view(T, 1, 1, 1).expand(T, 3, 2, 2).contiguous()
    mine = frame_shard(T, rank, world)
    enc = lambda f: f[:, 0, 0, 0].view(-1, 1, 1).expand(-1, 4, 5).contiguous() * 10.0     # fake per-frame encoder
    tok = encode_frames_sharded(enc, frames[mine], T)
    ret[rank] = bool(tok.shape == (T, 4, 5) and torch.equal(tok[:, 0, 0], torch.arange(T, dtype=torch.float32) * 10))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharded_encode_gathers_in_frame_order():
    for T, port in ((8, 29613), (7, 29614)):
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_worker_frames, args=(2, port, T, ret), nprocs=2, join=True)
        assert all(ret[r] for r in range(2)), (T, dict(ret))


def test_frame_shard_blocks():
    sys.path.insert(0, ROOT)
    from stllm_b200.clip_parallel import frame_shard
    assert frame_shard(64, 1, 4) == list(range(16, 32))
    assert sum((frame_shard(7, r, 3) for r in range(3)), []) == list(range(7))
    # config 4: residual index [8,24,40,56] -> frame 8+16r lives on rank r
    assert all((8 + 16 * r) in frame_shard(64, r, 4) for r in range(4))
