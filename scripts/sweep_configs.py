"""BASELINE.json configs 3 and 5 on ONE B200 (the multi-GPU legs of those configs are clip-parallel replicas of this):
  * config-5-like sweep: T in {8,16,32} x clips in {1,4,16}, video_input='all' (S = 32 T + 64), full 39/12/32-layer model,
    graph-captured VideoPrefillPipeline, frames/s device-resident;
  * config 3: 8 clips x 16 frames, dynamic mask (rate fixed 0.5) + MVM head: STLLMForCausalLM.forward(samples) -- the
    training-branch hot path a1-a22 (encode_img, mask, prompt wrap, masked S~320 and unmasked S~576 LLM passes with all
    hidden states, MVM + CE losses), eager (host-side prompt assembly as in the reference).
Synthetic frames, random-init weights.  Usage: python scripts/sweep_configs.py [--quick]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stllm_b200.pipeline import VideoPrefillPipeline, build_model
from tests.fake_tokenizer import FakeTokenizer

quick = "--quick" in sys.argv
dev = torch.device("cuda", 0)
res = []


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


model = build_model(device=dev, seed=0)
for T in ((16,) if quick else (8, 16, 32)):
    for B in ((1, 4) if quick else (1, 4, 16)):
        pipe = VideoPrefillPipeline(model, B, T, 64, 224, use_graph=True)
        pipe.frames.normal_()
        pipe.prompt_ids.random_(3, 32000)
        pipe.capture()
        ms = timed(pipe.run, 5 if B * T >= 256 else 10)
        r = dict(config="5-like", frames=T, clips=B, seq_len=32 * T + 64, ms_per_step=ms, frames_per_s=B * T / ms * 1e3)
        print(json.dumps(r), flush=True)
        res.append(r)
        del pipe
        torch.cuda.empty_cache()
del model
torch.cuda.empty_cache()

# ---- config 3: dynamic mask + MVM head, 8 clips ----
ltok = FakeTokenizer(32000, pad_token_id=0, bos_token_id=1, eos_token_id=2)
model = build_model(device=dev, seed=0, use_mask=True, mvm_decode=True, llama_tokenizer=ltok)
sm = model.get_model().stllm_model
sm.fixed_mask_rate = 0.5                      # config 3 fixes the dynamic mask ratio at 0.5
B, T = 8, 16
g = torch.Generator().manual_seed(1234)
samples = {"image": torch.randn(B, T, 3, 224, 224, generator=g).half().to(dev),
           "instruction_input": ["<ImageHere>###Human: what is the person doing in this video ? ###Assistant: "] * B,
           "answer": ["the person is cooking dinner in the kitchen"] * B}


def step3():
    np.random.seed(1237)
    out = model(samples=samples)
    return out.loss


loss = step3()
ms = timed(step3, 3 if quick else 5)
r = dict(config="3", clips=B, frames=T, ms_per_step=ms, frames_per_s=B * T / ms * 1e3, loss=float(loss),
         masked_len=int(sm.mask_img_len), img_len=int(sm.img_len), note="eager; two LLM passes + MVM + CE")
print(json.dumps(r), flush=True)
res.append(r)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/sweep_configs.json", "w") as f:
    json.dump(res, f, indent=1)
