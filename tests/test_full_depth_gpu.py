"""GPU: FULL-DEPTH parity of the configuration bench.py times (BASELINE config 2: 16 frames, EVA-ViT-g 39 blocks, Q-Former 12
layers, Vicuna-7B dims 32 layers, S = 576, logits for all positions) -- VERDICT r01 "full depth is never parity-checked".

Reference: the oracle's arithmetic (oracle/stllm_oracle.py: the reference's functions restated as torch ops) evaluated in
fp32 ON THE GPU (torch.backends.cuda.matmul.allow_tf32 = False: true fp32 GEMMs) on the same random-init weights and inputs;
an fp32 CPU run of the same code would need 32 GB of fp32 weights on the host and minutes of CPU time.  north_star's one-ulp
tolerance cannot hold through 83 fp16 layers (nor for the reference's own fp16 GPU execution), so the bar is stated as an
ERROR BUDGET next to what torch-fp16-eager (= what the reference executes on a GPU) achieves on the same box:
   rel. Frobenius error of the logits vs fp32:  ours <= 1.5 x torch-fp16-eager + 1e-3,  and ours < 3e-2 absolute budget."""
import pytest
import torch

from oracle import stllm_oracle as O

pytestmark = pytest.mark.gpu


def rel_fro(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm()).item()


def test_config2_full_depth_logits_vs_fp32_arithmetic_and_fp16_eager():
    from stllm_b200.pipeline import VideoPrefillPipeline, build_model
    assert torch.backends.cuda.matmul.allow_tf32 is False
    dev = torch.device("cuda", 0)
    model = build_model(device=dev, seed=0)                               # 39 / 12 / 32 layers, Vicuna-7B dims
    pipe = VideoPrefillPipeline(model, 1, 16, 64, 224, use_graph=False, frames_dtype=torch.float32)
    g = torch.Generator().manual_seed(1234)
    frames = torch.randn(1, 16, 3, 224, 224, generator=g)
    ids = torch.randint(3, 32000, (1, 64), generator=torch.Generator().manual_seed(1235))
    pipe.frames.copy_(frames)
    pipe.prompt_ids.copy_(ids)
    ours = pipe.run().float()
    torch.cuda.synchronize()
    assert ours.shape == (1, 576, 32000) and torch.isfinite(ours).all()
    cfg = dict(vit_depth=39, vit_heads=16, qf_layers=12, llama_layers=32, llama_heads=32)
    sd = model.state_dict()
    with torch.no_grad():
        # the product rounds the fp32 frames to fp16 in the patch-embed loader, as the reference's autocast conv does:
        # feed both references the same fp16-rounded frames
        fr16 = frames.half().to(dev)
        w16 = O.split_state_dict(sd, dtype=torch.float16, device=dev)          # everything fp16, as demo.py:46
        eager = O.video_prefill(w16, fr16, ids.to(dev), cfg).float()
        del w16
        w32 = O.split_state_dict(sd, dtype=torch.float32, device=dev)
        ref = O.video_prefill(w32, fr16.float(), ids.to(dev), cfg)
        del w32
    e_ours, e_eager = rel_fro(ours, ref), rel_fro(eager, ref)
    amax = (ours - ref).abs().max().item()
    top1 = (ours.argmax(-1) == ref.argmax(-1)).float().mean().item()
    print(f"FULL DEPTH config 2: rel.Fro vs fp32  ours {e_ours:.3e}  torch-fp16-eager {e_eager:.3e}  max|err| {amax:.3e} "
          f"(|ref|max {ref.abs().max().item():.2f})  top-1 agreement with fp32 {top1:.4f}")
    assert e_ours <= 1.5 * e_eager + 1e-3, (e_ours, e_eager)
    assert e_ours < 3e-2, e_ours
