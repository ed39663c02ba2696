"""Open item of profiles/r01_decode_summary.md section 3: is the first 5-beam engine of a process slow only for its first
steps?  Times consecutive windows of 10 steps of ONE cold engine (CUDA events), with and without the per-step ancestry
re-ordering, and prints one line per window.
Usage: python scripts/decode_transient.py [--beams 5] [--windows 30] [--no-reorder]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stllm_b200.decode import DecodeEngine
from stllm_b200.pipeline import build_model


def arg(name, dflt):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else dflt


beams, windows, S = arg("--beams", 5), arg("--windows", 30), 576
reorder = "--no-reorder" not in sys.argv
model = build_model(llama_layers=32, vit_depth=1, qformer_layers=1, device=torch.device("cuda", 0), seed=0)
g = torch.Generator(device="cuda").manual_seed(5)
emb = (torch.randn(1, S, 4096, device="cuda", generator=g) * 0.05).half()
out = model(samples=None, inputs_embeds=emb, use_cache=True, cache_max_len=S + 10 * windows + 8)
eng = DecodeEngine(model, out.past_key_values, S, beams, 10 * windows + 8)
tok = torch.randint(3, 32000, (beams,), device="cuda", generator=g)
idx = torch.arange(beams, device="cuda").flip(0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(windows + 1)]
torch.cuda.synchronize()
ev[0].record()
for w in range(windows):
    for _ in range(10):
        if reorder and beams > 1:
            eng.reorder(idx)
        eng.step(tok)
    ev[w + 1].record()
torch.cuda.synchronize()
ms = [round(ev[w].elapsed_time(ev[w + 1]) / 10, 3) for w in range(windows)]
print("TRANSIENT " + json.dumps(dict(beams=beams, reorder=reorder, ms_per_token_by_window_of_10_steps=ms)))
