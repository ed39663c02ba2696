"""Two eager decode steps (no CUDA graph) of the 32-layer Vicuna-7B-size decoder inside a cudaProfilerStart/Stop range,
for `ncu --profile-from-start off --metrics gpu__time_duration.sum` (per-kernel launch list of a decode step).
Usage: python scripts/decode_profile.py [--beams N] [--prefix S] [--gen G]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stllm_b200.decode import DecodeEngine
from stllm_b200.pipeline import build_model


def arg(name, dflt):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else dflt


beams, S, gen = arg("--beams", 5), arg("--prefix", 576), arg("--gen", 40)
model = build_model(llama_layers=32, vit_depth=1, qformer_layers=1, device=torch.device("cuda", 0), seed=0)
g = torch.Generator(device="cuda").manual_seed(5)
emb = (torch.randn(1, S, 4096, device="cuda", generator=g) * 0.05).half()
out = model(samples=None, inputs_embeds=emb, use_cache=True, cache_max_len=S + gen + 8)
eng = DecodeEngine(model, out.past_key_values, S, beams, gen + 8, use_graph=False)
tok = torch.randint(3, 32000, (beams,), device="cuda", generator=g)
for _ in range(gen):          # fill some suffix so the per-beam part of the attention has work
    eng.step(tok)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    eng.step(tok)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled 2 decode steps, launches/step", eng.launches_per_step)
