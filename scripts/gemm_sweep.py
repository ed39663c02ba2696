"""Sweep the tile configurations of the tcgen05 GEMM over the hot shapes of config 2 (and bs 8) next to cuBLAS, to
(re-)fit the planner in csrc/gemm_tcgen05.cu.  Interleaved rounds, CUDA-event timing, L2-warm back-to-back launches
(as inside a step).  Output: one line per shape with every variant's median, sorted; JSON at the end."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stllm_b200 import ops

dev = "cuda"
SHAPES = [("vit_qkv", 4112, 4224, 1408, 0, True, False), ("vit_proj", 4112, 1408, 1408, 0, True, True),
          ("vit_fc1", 4112, 6144, 1408, 1, True, False), ("vit_fc2", 4112, 1408, 6144, 0, True, True),
          ("qf_crosskv", 4112, 9216, 1408, 0, True, False),
          ("llama_qkv", 576, 12288, 4096, 0, False, False), ("llama_o", 576, 4096, 4096, 0, False, True),
          ("llama_gateup", 576, 22016, 4096, 2, False, False), ("llama_down", 576, 4096, 11008, 0, False, True),
          ("lm_head", 576, 32000, 4096, 0, False, False),
          ("llama_qkv_bs8", 4608, 12288, 4096, 0, False, False), ("llama_o_bs8", 4608, 4096, 4096, 0, False, True)]
SMALL = [("qf_qkv", 512, 2304, 768, 0, True, False), ("qf_selfout", 512, 768, 768, 0, True, True),
         ("qf_ffn_in", 512, 3072, 768, 1, True, False), ("qf_ffn_out", 512, 768, 3072, 0, True, True),
         ("llama_proj", 512, 4096, 768, 0, True, False), ("patch_embed", 4096, 1408, 640, 0, True, False)]
if "--quick" in sys.argv:
    SHAPES = SHAPES[:10]
if "--small" in sys.argv:
    SHAPES = SMALL


def variants(M, N, K, epi):
    v = [0, 65, 129, 257]
    v += [1024 + bn + 1 for bn in (96, 112, 128, 144, 160, 176, 192, 208, 224, 240, 256)]
    if M <= 1024:
        v += [2048 + bn + 1 for bn in (32, 48, 64, 80, 96, 112, 128, 144, 160, 176, 192, 208, 224, 240, 256) if bn <= ((M + 15) // 16) * 16]
    if os.environ.get("SWEEP_SK"):
        v = [0] + [1024 + bn for bn in (176, 192, 208, 224, 240, 256)] + [256]
    return v


def main():
    out = {}
    for name, M, N, K, epi, bias, res in SHAPES:
        a = torch.randn(M, K, device=dev, dtype=torch.float16)
        w = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.02
        n_out = N // 2 if epi == ops.EPI_SWIGLU else N
        c = torch.empty(M, n_out, device=dev, dtype=torch.float16)
        b = torch.randn(N, device=dev, dtype=torch.float16) if bias else None
        r = torch.randn(M, n_out, device=dev, dtype=torch.float16) if res else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def run(fn, iters=10):
            fn()
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters * 1e3
        vs = variants(M, N, K, epi)
        t = {v: [] for v in vs}
        t["cublas"] = []
        for _ in range(3):
            for v in vs:
                try:
                    t[v].append(run(lambda: ops.gemm(a, w, bias=b, residual=r, epilogue=epi, out=c, tile_n=v)))
                except Exception as e:   # a configuration the kernel rejects
                    t[v].append(float("nan"))
            t["cublas"].append(run(lambda: torch.matmul(a, w.t())))
        med = {k: sorted(x)[len(x) // 2] for k, x in t.items()}
        fl = 2.0 * M * N * K
        order = sorted((k for k in med if med[k] == med[k]), key=lambda k: med[k])
        plan = ops.gemm_plan(M, N, K, epi)
        print(f"SWEEP {name} M={M} N={N} K={K} epi={epi} plan={plan} auto={med[0]:.1f}us cublas={med['cublas']:.1f}us "
              f"({fl/med['cublas']/1e6:.0f}TF) | " + "  ".join(f"{k}:{med[k]:.1f}" for k in order[:8]), flush=True)
        out[name] = {str(k): med[k] for k in med}
    print("SWEEP_JSON " + json.dumps(out))


if __name__ == "__main__":
    main()
