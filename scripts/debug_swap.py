"""debug aid: one GEMM case per process (a CUDA fault poisons the context)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stllm_b200 import ops
M, N, K, tn = [int(x) for x in sys.argv[1:5]]
epi = int(sys.argv[5]) if len(sys.argv) > 5 else 0
res = int(sys.argv[6]) if len(sys.argv) > 6 else 0
torch.manual_seed(0)
a = torch.randn(M, K, device="cuda", dtype=torch.float16) * 0.5
w = torch.randn(N, K, device="cuda", dtype=torch.float16) * 0.05
n_out = N // 2 if epi == 2 else N
r = torch.randn(M, n_out, device="cuda", dtype=torch.float16) if res else None
ref = a.float() @ w.float().t()
if epi == 2:
    ref = torch.nn.functional.silu(ref[:, 0::2]) * ref[:, 1::2]
if res:
    ref = ref + r.float()
out = ops.gemm(a, w, residual=r, epilogue=epi, tile_n=tn)
torch.cuda.synchronize()
err = (out.float() - ref).abs()
bad = (err > 2e-2 + 1e-2 * ref.abs())
print(f"M={M} N={N} K={K} tn={tn} epi={epi} res={res}: max_err={err.max().item():.3e} bad={bad.sum().item()}")
if bad.any():
    rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
    print("  bad rows", rows.numel(), rows[:10].tolist(), "bad cols", cols.numel(), cols[:10].tolist())
