"""Pins oracle/stllm_oracle.py against fixtures produced by the REFERENCE ITSELF (tests/golden/make_golden.py).
fp32 on CPU both sides -> tight tolerances (only op-order differences)."""
import numpy as np
import torch

from oracle import stllm_oracle as O
from tests.util import assert_close, load_golden


def test_vit_matches_reference():
    sd, g = load_golden("vit_small.npz")
    frames = torch.from_numpy(g["frames"])
    pe = O.vit_patch_embed(sd, frames, 14)
    assert_close(pe, g["patch_embed"], 1e-5, 1e-5, "patch_embed")
    out = O.vit_forward(sd, frames, int(g["depth"]), int(g["heads"]))
    assert_close(out, g["out"], 2e-5, 1e-5, "vit out")


def test_qformer_matches_reference():
    sd, g = load_golden("qformer_small.npz")
    out = O.qformer_forward(sd, torch.from_numpy(g["query_tokens"]), torch.from_numpy(g["image_embeds"]),
                            int(g["layers"]), int(g["heads"]))
    assert_close(out, g["out"], 2e-5, 1e-5, "qformer out")


def test_qformer_text_matches_reference():
    sd, g = load_golden("qformer_text_small.npz")
    out = O.qformer_forward(sd, torch.from_numpy(g["query_tokens"]), torch.from_numpy(g["image_embeds"]),
                            int(g["layers"]), int(g["heads"]), text_ids=torch.from_numpy(g["text_ids"]),
                            text_mask=torch.from_numpy(g["text_mask"]))
    assert_close(out, g["out"], 2e-5, 1e-5, "qformer text out")


def test_llama_matches_reference():
    sd, g = load_golden("llama_small.npz")
    msd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
    emb = torch.from_numpy(g["inputs_embeds"])
    att = torch.from_numpy(g["attention_mask"])
    hid, hs = O.llama_forward(msd, emb, att, int(g["layers"]), int(g["heads"]), all_hidden=True)
    # padded query rows are don't-care in the reference too, but they are deterministic: compare everything
    assert_close(hs[1], g["layer1_hidden"], 2e-5, 1e-5, "layer-1 hidden")
    assert_close(hid, g["last_hidden"], 2e-5, 1e-5, "final hidden")
    logits = torch.nn.functional.linear(hid, sd["lm_head.weight"])
    assert_close(logits, g["logits"], 5e-5, 1e-5, "logits")


def test_mask_and_residual_index_bit_exact():
    _, g = load_golden("mask_kat.npz")
    for seed, (L, B) in {0: (8, 2), 1237: (512, 8), 7: (128, 3)}.items():
        np.random.seed(seed)
        rate = O.draw_mask_rate()
        m = O.random_masking_generator(L, rate, B)
        assert rate == float(g[f"seed{seed}_rate"])
        assert np.array_equal(m.numpy(), g[f"seed{seed}_mask"])
        np.random.seed(seed)
        m2 = O.random_masking_generator(L, 0.5, B)
        assert np.array_equal(m2.numpy(), g[f"seed{seed}_mask_fixed_half"])
        assert int(m2.sum()) == B * int(0.5 * L)
    for (R, T) in [(4, 64), (16, 64), (4, 16), (2, 4), (3, 10), (5, 7)]:
        assert np.array_equal(O.get_residual_index(R, T), g[f"residual_R{R}_T{T}"])
    # known answers quoted in SURVEY.md 8c
    assert O.get_residual_index(4, 64).tolist() == [8, 24, 40, 56]
    assert O.get_residual_index(4, 16).tolist() == [2, 6, 10, 14]
    assert O.get_residual_index(16, 64).tolist() == [2 + 4 * i for i in range(16)]


def test_btadapter_matches_reference():
    """eva_btadapter.py: oracle vs the reference's own EVAVisionTransformer_BTAdapter"""
    sd, g = load_golden("btadapter_small.npz")
    frames = torch.from_numpy(g["frames"])
    out = O.bt_adapter_forward(sd, frames, int(g["depth"]), int(g["heads"]), bt_depth=int(g["bt_depth"]))
    assert_close(out, g["out"], 3e-5, 1e-5, "bt-adapter out [B,T]")
    out1 = O.bt_adapter_forward(sd, frames[0], int(g["depth"]), int(g["heads"]), bt_depth=int(g["bt_depth"]))
    assert_close(out1, g["out_clip0"], 3e-5, 1e-5, "bt-adapter out, 4-D single clip")


def test_encode_img_without_qformer_matches_reference():
    """SURVEY a15': has_qformer=False (st_llm.py:300-301,369-373), fixture from the reference's own encode_img"""
    sd, g = load_golden("noqformer_small.npz")
    w = {"vit": {k[len("visual_encoder."):]: v for k, v in sd.items() if k.startswith("visual_encoder.")},
         "ln_vision": {k[len("ln_vision."):]: v for k, v in sd.items() if k.startswith("ln_vision.")},
         "llama_proj": {k[len("llama_proj."):]: v for k, v in sd.items() if k.startswith("llama_proj.")}}
    cfg = dict(vit_depth=int(g["depth"]), vit_heads=int(g["heads"]), has_qformer=False)
    clips = torch.from_numpy(g["clips"])
    out5 = O.encode_img(w, clips, cfg)
    out4 = O.encode_img(w, clips[0], cfg)
    assert tuple(out5.shape) == (2, 3, 4, 64) == tuple(g["out5"].shape) and tuple(out4.shape) == (3, 4, 64)
    assert tuple(g["atts5_shape"]) == (2, 3, 4) and tuple(g["atts4_shape"]) == (3, 4) and g["use_image"].tolist() == [0, 1]
    assert_close(out5, g["out5"], atol=2e-5, rtol=1e-4, what="encode_img has_qformer=False (5-D)")
    assert_close(out4, g["out4"], atol=2e-5, rtol=1e-4, what="encode_img has_qformer=False (4-D)")
