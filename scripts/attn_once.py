"""one warm + a few profiled launches of the ViT-shape attention (for ncu --set full)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stllm_b200 import ops
B, H, S, D = 16, 16, 257, 88
qkv = torch.randn(B, S, 3, H, D, device="cuda", dtype=torch.float16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
out = torch.empty(B, S, H, D, device="cuda", dtype=torch.float16)
for _ in range(4):
    ops.attention(q, k, v, 1.0, q_prescale=D ** -0.5, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    ops.attention(q, k, v, 1.0, q_prescale=D ** -0.5, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
