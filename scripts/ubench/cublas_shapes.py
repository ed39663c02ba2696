"""What does cuBLAS (torch F.linear, fp16) launch for the hot GEMM shapes of config 2?  Run under
`ncu --set full --profile-from-start off` to read its kernel names (tile shape, 1SM/2SM), grid, cluster, tensor-pipe
utilisation -- the comparison bar VERDICT r01 asks for.  Without ncu: prints CUDA-event times (L2-warm, back to back)."""
import sys
import torch
import torch.nn.functional as F

SHAPES = [("vit_qkv", 4112, 4224, 1408), ("vit_proj", 4112, 1408, 1408), ("vit_fc1", 4112, 6144, 1408),
          ("vit_fc2", 4112, 1408, 6144), ("llama_qkv", 576, 12288, 4096), ("llama_o", 576, 4096, 4096),
          ("llama_gateup", 576, 22016, 4096), ("llama_down", 576, 4096, 11008), ("lm_head", 576, 32000, 4096)]
dev = "cuda"
torch.manual_seed(0)
ops = []
for name, M, N, K in SHAPES:
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    w = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.02
    ops.append((name, M, N, K, a, w))
for name, M, N, K, a, w in ops:
    for _ in range(5):
        F.linear(a, w)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, M, N, K, a, w in ops:
    F.linear(a, w)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
if "--time" in sys.argv:
    for name, M, N, K, a, w in ops:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            F.linear(a, w)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"cublas {name} M={M} N={N} K={K}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TF")
