"""torchrun probe: where does time go when a CUDA-graph step is followed by an NCCL all-gather without host sync?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from stllm_b200.pipeline import VideoPrefillPipeline, build_model

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
model = build_model(llama_layers=8, vit_depth=10, qformer_layers=4, device=dev)
pipe = VideoPrefillPipeline(model, 1, 16, 64)
pipe.capture()
buf = torch.empty(world, 32000, dtype=torch.float16, device=dev)
side = torch.cuda.Stream()


def timed(name, fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize(); t1 = time.time()
    if rank == 0:
        print(f"{name:45s} device {e0.elapsed_time(e1)/n:7.2f} ms/step   host {1e3*(t1-t0)/n:7.2f} ms/step", flush=True)


def graph_only():
    pipe.run()

def graph_ag():
    pipe.run(); dist.all_gather_into_tensor(buf, pipe.last_logits)

def graph_ag_sync():
    pipe.run(); dist.all_gather_into_tensor(buf, pipe.last_logits); torch.cuda.current_stream().synchronize()

def ag_only():
    dist.all_gather_into_tensor(buf, pipe.last_logits)

def graph_ag_async():
    pipe.run()
    h = dist.all_gather_into_tensor(buf, pipe.last_logits, async_op=True)
    return h

timed("graph only", graph_only)
timed("all_gather only", ag_only)
timed("graph + all_gather (no host sync)", graph_ag)
timed("graph + all_gather + stream sync", graph_ag_sync)
timed("graph + all_gather async_op", graph_ag_async)
os.environ["X"] = "1"
dist.destroy_process_group()
