"""Where does each GEMM role wait?  Prints, per shape and tile engine, the average cycles per CTA that the TMA producer,
the MMA issuer and the epilogue spend blocked on their mbarriers, as a fraction of the kernel's duration."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stllm_b200 import _lib, ops

dev = "cuda"
lib = _lib.load()


def counters():
    buf = (ctypes.c_ulonglong * 8)()
    _lib.check(lib.stllm_gemm_debug_counters(ctypes.cast(buf, ctypes.c_void_p)), "counters")
    return list(buf)


def run(M, N, K, epi, tn, bias=False, residual=False, iters=10):
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    w = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.02
    n_out = N // 2 if epi == ops.EPI_SWIGLU else N
    out = torch.empty(M, n_out, device=dev, dtype=torch.float16)
    b = torch.randn(N, device=dev, dtype=torch.float16) if bias else None
    r = torch.randn(M, n_out, device=dev, dtype=torch.float16) if residual else None
    for _ in range(3):
        ops.gemm(a, w, bias=b, residual=r, epilogue=epi, out=out, tile_n=tn)
    counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, w, bias=b, residual=r, epilogue=epi, out=out, tile_n=tn)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    c = counters()
    launches = max(c[5], 1)
    ctas = 148
    kcyc = c[4] / launches
    f = lambda x: x / launches / ctas / max(kcyc, 1)
    print(f"M={M} N={N} K={K} epi={epi} tn={tn}: {us:6.1f}us {2.0*M*N*K/us/1e6:6.0f}TF  mma-warp cycles/launch={kcyc:8.0f}  "
          f"wait/CTA as fraction of kernel: producer(empty)={f(c[0]):.2f} mma(full)={f(c[1]):.2f} "
          f"mma(tmem_empty)={f(c[2]):.2f} epi(tmem_full)={f(c[3]):.2f}", flush=True)


for tn in (513, 257):
    for epi, bias in ((ops.EPI_NONE, False), (ops.EPI_NONE, True), (ops.EPI_GELU, True), (ops.EPI_RELU, True)):
        run(4112, 6144, 1408, epi, tn, bias=bias)
    run(4112, 1408, 6144, ops.EPI_NONE, tn, bias=True, residual=True)
    run(4112, 1408, 6144, ops.EPI_NONE, tn)
    run(576, 22016, 4096, ops.EPI_SWIGLU, tn)
    run(576, 22016, 4096, ops.EPI_NONE, tn)
