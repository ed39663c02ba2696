"""Frame preprocessing: the numpy oracle against fixtures produced by the reference's own transform
classes (bit-exact, CPU), the host window tables, and -- on the GPU -- the CUDA kernels against both, bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as PO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_frames.npz")
CASES = ("land", "port", "up")


def test_oracle_bit_exact_vs_reference_transform():
    g = np.load(GOLDEN)
    for name in CASES:
        got = PO.preprocess_frames([g[f"{name}_frame"]])
        ref = g[f"{name}_out"]
        assert got.dtype == np.float32 and got.shape == ref.shape == (1, 3, 224, 224)
        assert np.array_equal(got, ref), name                      # fp32 bit-equal: integer resampler + 3 IEEE ops


def test_size_rules_known_answers():
    # torchvision Resize(224): shorter edge -> 224, longer = int(224 * long / short); centre crop offsets
    assert PO.resized_size(360, 640, 224) == (224, 398)
    assert PO.resized_size(480, 360, 224) == (298, 224)
    assert PO.resized_size(224, 224, 224) == (224, 224)
    assert PO.crop_offsets(224, 398, 224) == (0, 87)
    assert PO.crop_offsets(298, 224, 224) == (37, 0)
    b, k = PO.precompute_coeffs(448, 224)
    assert k.shape == (224, 9) and int(k[100].sum()) in range((1 << 22) - 4, (1 << 22) + 5)   # weights sum to ~1.0
    assert tuple(b[0]) == (0, 5) and tuple(b[223]) == (443, 5)                                 # clipped at the borders


def test_host_window_tables_match_oracle():
    from stllm_b200 import video_preprocess as VP
    for a, b in [(640, 398), (360, 224), (100, 224), (1280, 398), (250, 224), (300, 268)]:
        b1, k1 = VP.resample_windows(a, b)
        b2, k2 = PO.precompute_coeffs(a, b)
        assert np.array_equal(b1, b2) and np.array_equal(k1, k2)
    b, k = VP.resample_windows(224, 224)
    assert np.array_equal(b[:, 0], np.arange(224)) and (k == 1 << 22).all()


@pytest.mark.gpu
def test_cuda_preprocess_bit_exact_vs_reference_fixture_and_oracle():
    from stllm_b200.video_preprocess import FramePreprocessor
    pre = FramePreprocessor(device="cuda")
    g = np.load(GOLDEN)
    for name in CASES:
        out = pre(g[f"{name}_frame"][None])
        assert out.dtype == torch.float16 and tuple(out.shape) == (1, 3, 224, 224)
        ref = torch.from_numpy(g[f"{name}_out"]).half()             # the reference casts to fp16 at the patch-embed conv
        assert torch.equal(out.cpu(), ref), name
    rng = np.random.default_rng(7)
    for (h, w, n) in [(360, 640, 16), (480, 360, 3), (224, 224, 2), (720, 1280, 4), (100, 160, 2), (333, 223, 1)]:
        frames = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
        out = pre(torch.from_numpy(frames))
        ref = torch.from_numpy(PO.preprocess_frames(list(frames))).half()
        assert torch.equal(out.cpu(), ref), (h, w)
    pil_like = [frames[0], frames[0]]                               # list input (PIL images / arrays of one size)
    assert torch.equal(pre(pil_like)[0], pre(pil_like)[1])


@pytest.mark.gpu
def test_preprocessed_frames_feed_encode_img():
    """upload_video (conversation.py:274-279): transform -> view(bt//3, 3, w, h) -> encode_img on a 4-D tensor"""
    from stllm_b200.pipeline import build_model
    from stllm_b200.video_preprocess import FramePreprocessor
    model = build_model(llama_layers=1, vit_depth=1, qformer_layers=2, vocab_size=64, llama_inter=256)
    sm = model.get_model().stllm_model
    frames = np.random.default_rng(3).integers(0, 256, (4, 240, 426, 3), dtype=np.uint8)
    x = FramePreprocessor(device="cuda")(frames)
    emb, atts, use_image = sm.encode_img(x)
    assert tuple(emb.shape) == (4, 32, 4096) and use_image is True and torch.isfinite(emb).all()
