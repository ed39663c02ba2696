"""Sustained A/B of the step's GEMM launches, group by group: ours (all fused epilogue work) against cuBLAS (`torch.mm`, plain
GEMM) on the SAME weight tensors, each group = every launch of one (M, N, K, epilogue, norm) class in step order (so each launch
streams its own layer's weights from HBM, as in the real step), replayed back to back from one CUDA graph for >= 60 ms.
Answers: which shape classes lose against the library once the part sits at its power cap (burst timings of single shapes do
not show it).  Usage: python scripts/gemm_list_ab.py [--plain]   (--plain: ours without bias/act/residual/norm as a third arm)
"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

T_FRAMES, PROMPT_LEN, IMG, VOCAB = 16, 64, 224, 32000


class Sampler:
    """SM clock / board power while a timed loop runs (NVML, 10 ms polling)"""
    def __init__(self):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)

    def run(self, stop, out):
        import time
        while not stop.is_set():
            try:
                out.append((self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM),
                            self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3))
            except Exception:
                pass
            time.sleep(0.01)


SAMPLER = None
LAST = {}


def timed_graph(fn, min_ms=60.0):
    import threading
    global SAMPLER
    if SAMPLER is None:
        try:
            SAMPLER = Sampler()
        except Exception as e:
            print("no NVML:", e)
            SAMPLER = False
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        g.replay()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    one = max(e0.elapsed_time(e1), 1e-3)
    n = max(3, int(min_ms / one) + 1)
    stop, smp = threading.Event(), []
    th = None
    if SAMPLER:
        th = threading.Thread(target=SAMPLER.run, args=(stop, smp))
        th.start()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    if th is not None:
        stop.set()
        th.join()
    smp = smp[len(smp) // 3:]                               # drop the ramp
    LAST["mhz"] = sorted(x[0] for x in smp)[len(smp) // 2] if smp else 0
    LAST["watt"] = sorted(x[1] for x in smp)[len(smp) // 2] if smp else 0
    return e0.elapsed_time(e1) / n


def main():
    from stllm_b200 import ops
    from stllm_b200.pipeline import VideoPrefillPipeline, build_model
    plain = "--plain" in sys.argv
    dev = torch.device("cuda", 0)
    model = build_model(llama_layers=32, vit_depth=39, qformer_layers=12, vocab_size=VOCAB, device=dev, seed=0)
    pipe = VideoPrefillPipeline(model, 1, T_FRAMES, PROMPT_LEN, IMG, use_graph=False, frames_dtype=torch.float32)
    pipe.frames.copy_(torch.randn(1, T_FRAMES, 3, IMG, IMG))
    pipe.prompt_ids.copy_(torch.randint(3, VOCAB, (1, PROMPT_LEN)))
    pipe.warmup(1)
    ops.gemm_recorder = []
    pipe._step()
    torch.cuda.synchronize()
    rec, ops.gemm_recorder = ops.gemm_recorder, None

    groups = collections.OrderedDict()
    for r in rec:
        key = (int(r["M"]), int(r["N"]), int(r["K"]), str(r["epilogue"]), int(r["norm"] or 0), r["bias"] is not None,
               r["residual"] is not None)
        groups.setdefault(key, []).append(r)

    scratch, cb_a, cb_o = {}, {}, {}

    def ours(rs, strip=False):
        def fn():
            for r in rs:
                k = (tuple(r["out"].shape), r["out"].dtype)
                if k not in scratch:
                    scratch[k] = torch.empty_like(r["out"])
                if strip:
                    ep = r["epilogue"] if r["epilogue"] == ops.EPI_SWIGLU else ops.EPI_NONE
                    ops.gemm(r["a"], r["w"], epilogue=ep, out=scratch[k], tile_n=r["tile_n"])
                else:
                    ops.gemm(r["a"], r["w"], bias=r["bias"], residual=r["residual"], epilogue=r["epilogue"], out=scratch[k],
                             tile_n=r["tile_n"], norm=r["norm"], rs_in=r["rs_in"], col_sb=r["col_sb"], norm_eps=r["norm_eps"])
        return fn

    def cublas(rs):
        def fn():
            for r in rs:
                M, N = int(r["M"]), int(r["N"])
                K = r["w"].numel() // N
                if (M, K) not in cb_a:
                    cb_a[(M, K)] = (torch.randn(M, K, device=dev) * 0.05).half()
                if (M, N) not in cb_o:
                    cb_o[(M, N)] = torch.empty(M, N, dtype=torch.float16, device=dev)
                torch.mm(cb_a[(M, K)], r["w"].detach().reshape(N, K).t(), out=cb_o[(M, N)])
        return fn

    import ctypes
    from stllm_b200 import _lib
    lib = _lib.load()

    def counters():
        buf = (ctypes.c_ulonglong * 8)()
        lib.stllm_gemm_debug_counters(buf)
        return list(buf)

    rows = []
    print(f"{'M':>6} {'N':>6} {'K':>6} {'epilogue':>10} nrm b r {'n':>4} | {'ours us':>8} {'cublas us':>9} {'ratio':>6}"
          + (f" {'plain us':>8}" if plain else "") + f" | {'ours ms/step':>12} {'cublas':>7}")
    tot_o = tot_c = 0.0
    for key, rs in groups.items():
        M, N, K, epi, nrm, hb, hr = key
        reps = max(1, 40 // len(rs))                      # at least ~40 launches per replay
        rs_rep = rs * reps
        MS = float(os.environ.get("AB_MIN_MS", "250"))
        counters()
        t_o = timed_graph(ours(rs_rep), MS) / len(rs_rep) * 1e3
        c_o, s_o = counters(), dict(LAST)
        t_c = timed_graph(cublas(rs_rep), MS) / len(rs_rep) * 1e3
        s_c = dict(LAST)
        t_p = timed_graph(ours(rs_rep, strip=True), MS) / len(rs_rep) * 1e3 if plain else None
        c_p, s_p = counters(), dict(LAST)
        nl = max(c_o[5], 1)
        # counters: [1] MMA warps waiting for operands, [2] MMA warps waiting for a free accumulator (epilogue-bound),
        # [4]/[5] cycles of CTA 0's MMA role per launch
        extra = (f" | clk {s_o['mhz']:4d}/{s_c['mhz']:4d}/{s_p['mhz']:4d} MHz  {s_o['watt']:4.0f}/{s_c['watt']:4.0f}/{s_p['watt']:4.0f} W"
                 f" | cta0 cyc/launch {c_o[4] / nl:8.0f} (plain {c_p[4] / max(c_p[5], 1):8.0f})"
                 f" wait_full {c_o[1] / nl / 74:7.0f} wait_acc {c_o[2] / nl / 74:7.0f} (plain {c_p[1] / max(c_p[5], 1) / 74:7.0f} {c_p[2] / max(c_p[5], 1) / 74:7.0f})")
        tot_o += t_o * len(rs) / 1e3
        tot_c += t_c * len(rs) / 1e3
        rows.append({"M": M, "N": N, "K": K, "epilogue": epi, "norm": nrm, "bias": hb, "residual": hr, "launches": len(rs),
                     "ours_us": t_o, "cublas_us": t_c, "plain_us": t_p})
        print(f"{M:6d} {N:6d} {K:6d} {epi:>10} {nrm:3d} {int(hb)} {int(hr)} {len(rs):4d} | {t_o:8.2f} {t_c:9.2f} {t_o / t_c:6.2f}"
              + (f" {t_p:8.2f}" if plain else "") + f" | {t_o * len(rs) / 1e3:12.3f} {t_c * len(rs) / 1e3:7.3f}" + extra, flush=True)
    print(f"sum over groups: ours {tot_o:.3f} ms   cuBLAS {tot_c:.3f} ms")
    t_all_o = timed_graph(ours(rec), 200.0)
    t_all_c = timed_graph(cublas(rec), 200.0)
    print(f"whole list in step order: ours {t_all_o:.3f} ms   cuBLAS {t_all_c:.3f} ms   ratio {t_all_o / t_all_c:.3f}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"groups": rows, "whole_list_ms": {"ours": t_all_o, "cublas": t_all_c}}, open("gpurun_out/r02_gemm_list_ab.json", "w"),
              indent=1)


if __name__ == "__main__":
    main()
