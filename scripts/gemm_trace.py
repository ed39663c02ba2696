"""Timeline of the CTA-pair GEMM: where do the microseconds of one launch go?  (debug aid, B200)
For each shape: 6 back-to-back launches with tracing on; prints, for the last two launches, the phases relative to
the first CTA's entry (median / max over CTAs, in us) and the gap to the previous launch's last exit."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stllm_b200 import _lib, ops

lib = _lib.load()
dev = "cuda"


def get_trace(on=1):
    buf = np.zeros((4, 160, 12), dtype=np.uint64)
    nxt = ctypes.c_int(0)
    _lib.check(lib.stllm_gemm_debug_trace(on, buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nxt)), "trace")
    return buf, nxt.value


def run(name, M, N, K, epi=0, tn=0, bias=False, residual=False, between=None):
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    w = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.02
    n_out = N // 2 if epi == ops.EPI_SWIGLU else N
    out = torch.empty(M, n_out, device=dev, dtype=torch.float16)
    b = torch.randn(N, device=dev, dtype=torch.float16) if bias else None
    r = torch.randn(M, n_out, device=dev, dtype=torch.float16) if residual else None
    for _ in range(30):
        ops.gemm(a, w, bias=b, residual=r, epilogue=epi, out=out, tile_n=tn)
    get_trace(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        ops.gemm(a, w, bias=b, residual=r, epilogue=epi, out=out, tile_n=tn)
    e1.record()
    tr, nxt = get_trace(0)
    us = e0.elapsed_time(e1) / 8 * 1e3
    order = [(nxt + i) % 4 for i in range(4)]      # oldest .. newest
    plan = ops.gemm_plan(M, N, K, epi) if tn == 0 else ("forced", tn)
    print(f"== {name}: M={M} N={N} K={K} epi={epi} tn={tn} plan={plan}  {us:.1f} us/launch  {2.0*M*N*K/us/1e6:.0f} TF")
    prev_exit = None
    for s in order[-2:]:
        t = tr[s].astype(np.int64)
        act = t[:, 0] > 0
        if not act.any():
            continue
        t = t[act]
        t0 = t[:, 0].min()
        rel = lambda c: (t[:, c][t[:, c] > 0] - t0) / 1e3
        def mm(c):
            x = rel(c)
            return f"{np.median(x):6.2f}/{x.max():6.2f}" if x.size else "   -  /   -  "
        gap = f"{(t0 - prev_exit)/1e3:6.2f}" if prev_exit is not None else "   -  "
        dur = (t[:, 6].max() - t0) / 1e3
        lead = t[:, 8] > 0
        cyc = np.median(t[lead, 8]); ns = np.median((t[lead, 4] - t[lead, 3]))
        kbs = int(t[:, 7].max()) * ((K + 63) // 64)
        print(f"  slot {s}: mma phase {cyc:.0f} cyc / {ns/1e3:.2f} us = {cyc/ns:.3f} GHz; ~{cyc/kbs:.0f} cyc per k-block (max-tiles CTA); "
              f"wait-operands {np.median(t[lead, 9])/cyc*100:.1f}% wait-acc {np.median(t[lead, 10])/cyc*100:.1f}%")
        print(f"  slot {s}: ctas={act.sum():3d} gap_prev={gap} entry={mm(0)} setup={mm(1)} pdl={mm(2)} mma0={mm(3)} "
              f"mmaN={mm(4)} epiN={mm(5)} exit={mm(6)} total={dur:6.2f} tiles/cta max={int(t[:,7].max())}")
        prev_exit = t[:, 6].max()


run("1024x1408 tn 1280", 4112, 1408, 1408, bias=True, residual=True, tn=1024 + 256 + 1)
run("proj bn128", 4112, 1408, 1408, bias=True, residual=True, tn=1024 + 128 + 1)
run("vit qkv", 4112, 4224, 1408, bias=True)
run("vit qkv 513", 4112, 4224, 1408, bias=True, tn=513)
run("vit proj", 4112, 1408, 1408, bias=True, residual=True)
run("vit proj 513", 4112, 1408, 1408, bias=True, residual=True, tn=513)
run("vit fc1", 4112, 6144, 1408, epi=ops.EPI_GELU, bias=True)
run("vit fc1 513", 4112, 6144, 1408, epi=ops.EPI_GELU, bias=True, tn=513)
run("vit fc1 nogelu 513", 4112, 6144, 1408, bias=True, tn=513)
run("vit fc2", 4112, 1408, 6144, bias=True, residual=True)
run("vit fc2 513", 4112, 1408, 6144, bias=True, residual=True, tn=513)
run("l qkv", 576, 12288, 4096)
run("l qkv 513", 576, 12288, 4096, tn=513)
run("l o", 576, 4096, 4096, residual=True)
run("l o 513", 576, 4096, 4096, residual=True, tn=513)
run("l gateup", 576, 22016, 4096, epi=ops.EPI_SWIGLU)
run("l gateup 513", 576, 22016, 4096, epi=ops.EPI_SWIGLU, tn=513)
run("l down", 576, 4096, 11008, residual=True)
run("l down 513", 576, 4096, 11008, residual=True, tn=513)
run("8192", 8192, 8192, 8192, tn=513)
