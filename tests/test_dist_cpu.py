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
