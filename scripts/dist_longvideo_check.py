"""torchrun check (N GPUs): BASELINE config 4 shape -- T frames of one clip sharded over the ranks, one all-gather of the
frame tokens, global-local merge (R=4), prefill -- must equal the single-GPU result bit for bit; prints timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from stllm_b200.clip_parallel import frame_shard, long_video_prefill
from stllm_b200.pipeline import build_model

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
T = int(os.environ.get("T_FRAMES", "64"))
layers = int(os.environ.get("LLAMA_LAYERS", "32"))
model = build_model(llama_layers=layers, vit_depth=39, qformer_layers=12, video_input="residual", residual_size=4,
                    device=dev, seed=0)       # same seed on every rank -> identical replicated weights
g = torch.Generator().manual_seed(1234)
frames = torch.randn(T, 3, 224, 224, generator=g).half().to(dev)
ids = torch.randint(3, 32000, (1, 64), generator=g).to(dev)
mine = frame_shard(T, rank, world)


def timed(fn, n=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, t.item()

sharded, ms_sh = timed(lambda: long_video_prefill(model, frames[mine], T, ids))
# single-GPU reference on rank 0: run the same code with the whole clip and no process group
if rank == 0:
    sm = model.get_model().stllm_model
    tok = sm.encode_img(frames)[0]
    vid = sm.aggregate_video(tok.unsqueeze(0))
    seq = torch.cat([vid[:, 0], sm._embed(ids)], dim=1)
    ref = model(samples=None, inputs_embeds=seq).logits
    same = torch.equal(ref, sharded)
    print(f"config4-like: T={T} world={world} S={seq.shape[1]} sharded {ms_sh:.2f} ms/clip = {T/ms_sh*1e3:.0f} frames/s; "
          f"bit-identical to single GPU: {same}; residual_index={sm.residual_index.tolist()}", flush=True)
    assert same
dist.barrier()
dist.destroy_process_group()
