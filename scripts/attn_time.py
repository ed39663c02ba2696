"""CUDA-event timing of the ViT / Llama prefill attention shapes (tcgen05 kernels), with the pieces of the wide-tile ViT
kernel isolated: full shape (257 x 257), no tail rows (256 x 257), no leftover key (256 x 256)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stllm_b200 import ops

dev = "cuda"


def t_attn(B, H, Sq, Skv, D, causal=False, pre=True, iters=30):
    q = torch.randn(B, Sq, H, D, device=dev, dtype=torch.float16)
    k = torch.randn(B, Skv, H, D, device=dev, dtype=torch.float16)
    v = torch.randn(B, Skv, H, D, device=dev, dtype=torch.float16)
    out = torch.empty(B, Sq, H, D, device=dev, dtype=torch.float16)
    fn = lambda: ops.attention(q, k, v, 1.0 if pre else D ** -0.5, causal=causal, q_prescale=D ** -0.5 if pre else 0.0, out=out)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    # device time: the launches are captured into a CUDA graph (as in the product's step) -- eager back-to-back calls are
    # HOST-bound at ~19 us per call (ctypes + three cuTensorMapEncodeTiled + launch)
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
    torch.cuda.current_stream().wait_stream(st)
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    fl = 4.0 * B * H * Sq * Skv * D * (0.5 if causal else 1.0)
    print(f"TIME attn B={B} H={H} Sq={Sq} Skv={Skv} D={D} causal={int(causal)}: {best:.1f} us ({fl / best / 1e6:.0f} TF)", flush=True)


from stllm_b200 import _lib  # noqa: E402
for tail in (0, 1, 2):
    _lib.load().stllm_set_attention_variant(0, tail)
    print(f"--- ViT shapes, trailing-row mode {tail} (0: third query tile, 1: tail CTAs in the grid, 2: separate tail launch)")
    for shape in ((16, 16, 257, 257, 88), (128, 16, 257, 257, 88)):
        t_attn(*shape)
_lib.load().stllm_set_attention_variant(-1, -1)
for shape in ((16, 16, 256, 257, 88), (16, 16, 256, 256, 88), (16, 16, 128, 256, 88)):
    t_attn(*shape)
t_attn(1, 32, 576, 576, 128, causal=True, pre=False)
t_attn(8, 32, 576, 576, 128, causal=True, pre=False)
