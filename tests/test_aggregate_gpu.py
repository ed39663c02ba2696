"""GPU: video aggregation modes (st_llm.py:463-478 / conversation.py:281-293), encode_img input conventions
(5-D training input vs 4-D inference input, st_llm.py:324-375) and the fixed-rate dynamic mask of BASELINE config 3."""
import numpy as np
import pytest
import torch

from oracle import stllm_oracle as O
from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_model():
    from stllm_b200.pipeline import build_model
    m = build_model(llama_layers=1, vit_depth=1, qformer_layers=2, vocab_size=128, llama_inter=256,
                    video_input="residual", residual_size=2, device=torch.device("cuda", 0), seed=3)
    return m


@pytest.mark.parametrize("mode", ["all", "mean", "residual"])
def test_aggregate_modes_vs_oracle(small_model, mode):
    sm = small_model.get_model().stllm_model
    g = torch.Generator().manual_seed(7)
    emb = (torch.randn(2, 4, 32, 4096, generator=g) * 0.5).half()
    w = O.split_state_dict(small_model.state_dict())
    ref = O.video_aggregate(emb.float(), mode, 2, w["residual"])
    sm.video_input = mode
    try:
        out = sm.aggregate_video(emb.cuda())
    finally:
        sm.video_input = "residual"
    assert tuple(out.shape) == tuple(ref.shape)
    assert_close(out, ref, atol=4e-3, rtol=4e-3, what=f"aggregate {mode}")
    if mode == "residual":
        assert sm.residual_index.tolist() == O.get_residual_index(2, 4).tolist() == [1, 3]


def test_encode_img_4d_inference_and_5d_training_inputs(small_model):
    sm = small_model.get_model().stllm_model
    g = torch.Generator().manual_seed(8)
    frames = torch.randn(3, 3, 224, 224, generator=g)
    w = O.split_state_dict(small_model.state_dict())
    cfg = dict(vit_depth=1, vit_heads=16, qf_layers=2)
    ref = O.encode_img(w, frames, cfg)                                   # [T,Q,D] (infer=True branch)
    out, atts, use_image = sm.encode_img(frames.cuda())
    assert tuple(out.shape) == (3, 32, 4096) and tuple(atts.shape) == (3, 32) and atts.dtype == torch.long
    assert use_image is True                                              # 4-D input: st_llm.py:326
    assert_close(out, ref, atol=3e-2, rtol=2e-2, what="encode_img 4-D")
    out5, atts5, use_image5 = sm.encode_img(frames[None].cuda())         # [1,T,3,H,W]
    assert tuple(out5.shape) == (1, 3, 32, 4096) and use_image5 is False
    assert torch.equal(out5[0], out)                                      # same frames, same kernels: bit-identical


def test_fixed_half_mask_matches_reference_generator(small_model):
    """config 3 fixes the mask rate at 0.5: n = int(0.5 L) masked tokens per row, same numpy stream as the reference"""
    from stllm_b200.st_llm import RandomMaskingGenerator
    np.random.seed(1237)
    _ = np.random.normal(0.5, 0.1)
    m = RandomMaskingGenerator(512, 0.5, 8, "cpu")
    np.random.seed(1237)
    _ = O.draw_mask_rate()
    ref = O.random_masking_generator(512, 0.5, 8)
    assert torch.equal(m, ref) and int(m.sum()) == 8 * 256
