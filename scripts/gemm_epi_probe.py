"""What does each fused-epilogue feature cost?  ViT q|k|v / fc1 / proj and Llama shapes, one feature at a time, graph-timed
back to back (same weights: L2-warm, so this isolates the kernel's own behaviour from HBM streaming), with the kernel's
wait-cycle counters (stllm_gemm_debug_counters).  Usage: python scripts/gemm_epi_probe.py [--once name]  (--once: one launch
of one variant, for ncu)"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from stllm_b200 import _lib, ops  # noqa: E402

dev = "cuda"


def counters():
    buf = (ctypes.c_ulonglong * 8)()
    _lib.load().stllm_gemm_debug_counters(buf)
    return list(buf)


def make(M, N, K, epi=0, bias=False, res=False, norm=0, stats=0, tile=0):
    g = torch.Generator(device=dev).manual_seed(1)
    a = torch.randn(M, K, device=dev, generator=g).half()
    w = (torch.randn(N, K, device=dev, generator=g) * 0.02).half()
    n_out = N // 2 if epi == ops.EPI_SWIGLU else N
    c = torch.empty(M, n_out, device=dev, dtype=torch.float16)
    b = torch.randn(N, device=dev, generator=g).half() if bias else None
    r = torch.randn(M, n_out, device=dev, generator=g).half() if res else None
    kw = {}
    if norm:
        rs = ops.row_stats_buffer(1, M, K, norm, dev)[0]
        rs.fill_(0.0)
        rs[:, :norm] = torch.tensor([0.1, float(K)][2 - norm:], device=dev)      # plausible sums: mean ~0, E[x^2] ~1
        kw.update(norm=norm, rs_in=rs, norm_eps=1e-5)
        if norm == ops.NORM_LN:
            kw["col_sb"] = torch.randn(N, 2, device=dev, generator=g).float().contiguous()
    if stats:
        kw.update(rs_out=ops.row_stats_buffer(1, M, n_out, stats, dev)[0], rs_out_mode=stats)
    return lambda: ops.gemm(a, w, bias=b, residual=r, epilogue=epi, out=c, tile_n=tile, **kw)


def timed(fn, n=40):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    c = counters()
    return e0.elapsed_time(e1) / n * 1e3, c


VARIANTS = {
    # name: (M, N, K, kwargs)
    "vqkv_plain": (4112, 4224, 1408, {}),
    "vqkv_bias": (4112, 4224, 1408, dict(bias=True)),
    "vqkv_ln": (4112, 4224, 1408, dict(norm=2)),
    "vqkv_rms": (4112, 4224, 1408, dict(norm=1)),
    "vfc1_plain": (4112, 6144, 1408, {}),
    "vfc1_gelu": (4112, 6144, 1408, dict(epi=1, bias=True)),
    "vfc1_ln": (4112, 6144, 1408, dict(norm=2)),
    "vfc1_ln_gelu": (4112, 6144, 1408, dict(epi=1, norm=2)),
    "vproj_plain": (4112, 1408, 1408, {}),
    "vproj_res": (4112, 1408, 1408, dict(bias=True, res=True)),
    "vproj_res_stats": (4112, 1408, 1408, dict(bias=True, res=True, stats=2)),
    "vfc2_plain": (4112, 1408, 6144, {}),
    "vfc2_res_stats": (4112, 1408, 6144, dict(bias=True, res=True, stats=2)),
    "lqkv_plain": (576, 12288, 4096, {}),
    "lqkv_rms": (576, 12288, 4096, dict(norm=1)),
    "lqkv_pair256": (576, 12288, 4096, dict(tile=512 + 1)),
    "lqkv_pair192": (576, 12288, 4096, dict(tile=1024 + 192 + 1)),
    "lo_plain": (576, 4096, 4096, {}),
    "lo_res_stats": (576, 4096, 4096, dict(res=True, stats=1)),
    "ldown_plain": (576, 4096, 11008, {}),
    "ldown_res_stats": (576, 4096, 11008, dict(res=True, stats=1)),
    "lgu_plain": (576, 22016, 4096, dict(epi=2)),
    "lgu_rms": (576, 22016, 4096, dict(epi=2, norm=1)),
    "lgu_rms_swap192": (576, 22016, 4096, dict(epi=2, norm=1, tile=2048 + 192 + 1)),
    "lgu_rms_pair256": (576, 22016, 4096, dict(epi=2, norm=1, tile=512 + 1)),
}


def main():
    if "--once" in sys.argv:
        name = sys.argv[sys.argv.index("--once") + 1]
        M, N, K, kw = VARIANTS[name]
        fn = make(M, N, K, **kw)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        return
    print(f"{'variant':18s} {'us':>8} | per launch and MMA-issuing CTA: cycles, wait_full, wait_acc | TMA wait_empty, epi wait_full (warp 0)")
    for name, (M, N, K, kw) in VARIANTS.items():
        t, c = timed(make(M, N, K, **kw))
        nl = max(c[5], 1)
        plan = ops.gemm_plan(M, N, K, kw.get("epi", 0))
        ncta = 148 if plan[0] == 0 and not kw.get("tile") else 74
        print(f"{name:18s} {t:8.2f} | {c[4] / nl:9.0f} {c[1] / nl / ncta:8.0f} {c[2] / nl / ncta:8.0f} | {c[0] / nl / 148:8.0f} "
              f"{c[3] / nl / 148:8.0f}   plan(no norm) {plan[:2]}", flush=True)


if __name__ == "__main__":
    main()
