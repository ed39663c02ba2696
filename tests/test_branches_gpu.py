"""GPU: the non-default branches of the path (SURVEY 8a a15'), through the C ABI, against fixtures written by the REFERENCE's
own methods:
  * has_qformer=False (st_llm.py:300-301,369-373): no Q-Former, cls dropped, 4 neighbouring patch tokens folded into one row,
    llama_proj 4C -> D -- `noqformer_small.npz` from the reference's encode_img (5-D clip batch and 4-D frame stack);
  * pre_encoding=True (st_llm.py:452-455): the input already holds Q-Former features, only llama_proj + 'mean' aggregation +
    sequence assembly run -- `pre_encoding_small.npz` from the reference's STLLMModel.forward (MiniGPT-4 layout)."""
from functools import partial

import numpy as np
import pytest
import torch
import torch.nn as nn

from tests.fake_tokenizer import FakeTokenizer
from tests.util import assert_close, load_golden

pytestmark = pytest.mark.gpu


def test_encode_img_without_qformer_vs_reference_fixture():
    from stllm_b200 import st_llm as S
    from stllm_b200.eva_vit import VisionTransformer
    sd, g = load_golden("noqformer_small.npz")
    sm = S.STLLMModel(has_qformer=False, vit_depth=1, llama_hidden=64)
    assert not hasattr(sm, "Qformer")
    # the fixture's encoder: the reference's VisionTransformer class at width 176 (2 heads of 88), 56 x 56 images, depth 2
    sm.visual_encoder = VisionTransformer(img_size=56, patch_size=14, use_mean_pooling=False, embed_dim=176, depth=int(g["depth"]),
                                          num_heads=int(g["heads"]), mlp_ratio=4.3637, qkv_bias=True,
                                          norm_layer=partial(nn.LayerNorm, eps=1e-6))
    sm.ln_vision = S.LayerNorm(176)
    sm.llama_proj = nn.Linear(176 * 4, 64)
    res = sm.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and not [k for k in res.missing_keys if "embed_tokens" not in k], res
    sm = sm.cuda().half()
    clips = torch.from_numpy(g["clips"]).cuda()
    out5, atts5, use5 = sm.encode_img(clips)
    out4, atts4, use4 = sm.encode_img(clips[0])
    assert tuple(atts5.shape) == tuple(g["atts5_shape"]) and tuple(atts4.shape) == tuple(g["atts4_shape"])
    assert [int(use5), int(use4)] == g["use_image"].tolist() and atts5.dtype == torch.long and bool((atts5 == 1).all())
    # 2 ViT blocks + LN + one GEMM in fp16 against the reference's fp32 CPU result: |out| <= 3
    assert_close(out5, g["out5"], atol=6e-3, rtol=4e-3, what="encode_img has_qformer=False (5-D) vs reference")
    assert_close(out4, g["out4"], atol=6e-3, rtol=4e-3, what="encode_img has_qformer=False (4-D) vs reference")


def test_pre_encoding_forward_vs_reference_fixture():
    from stllm_b200 import st_llm as S
    sd, g = load_golden("pre_encoding_small.npz")
    V, D = sd["embed_tokens.weight"].shape
    sm = S.STLLMModel.__new__(S.STLLMModel)
    nn.Module.__init__(sm)
    sm.pre_encoding, sm.video_input, sm.use_mask, sm.mvm_decode = True, "mean", False, False
    sm.qformer_text_input, sm.max_txt_len, sm.end_sym, sm.has_qformer = False, 8, " ###", True
    sm.llama_tokenizer = FakeTokenizer(V, pad_token_id=0, bos_token_id=1, eos_token_id=2)
    sm.llama_proj = nn.Linear(768, D)
    sm.llama_proj.load_state_dict({"weight": sd["llama_proj.weight"], "bias": sd["llama_proj.bias"]})
    emb = nn.Embedding(V, D)
    emb.weight.data.copy_(sd["embed_tokens.weight"])
    sm = sm.cuda().half()
    sm.embed_tokens = emb.cuda().half()
    samples = {"image": torch.from_numpy(g["feats"]).cuda(), "instruction_input": [str(p) for p in g["prompts"]],
               "answer": [str(a) for a in g["answers"]]}
    inputs_embeds, att, emb_u, att_u, targets = sm(samples)
    assert emb_u is None and att_u is None
    assert np.array_equal(att.cpu().numpy(), g["attention_mask"]) and np.array_equal(targets.cpu().numpy(), g["targets"])
    # token rows are exact copies of fp16-exact embeddings; the 32 video rows are llama_proj (K = 768) + mean over 3 frames
    ref = torch.from_numpy(g["inputs_embeds"])
    is_video = np.zeros(ref.shape[1], dtype=bool)
    is_video[8:8 + 32] = True                       # BOS + 7 prefix tokens, then the 32 'mean' video tokens
    got = inputs_embeds.float().cpu()
    assert torch.equal(got[:, ~is_video], ref[:, ~is_video]), "token-embedding rows must be bit-exact copies"
    assert_close(got[:, is_video], ref[:, is_video], atol=2e-3, rtol=2e-3, what="llama_proj + mean aggregation vs reference")
