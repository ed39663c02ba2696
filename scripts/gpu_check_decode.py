"""Decode-step kernels (csrc/decode.cu) and the decode engine on a B200:
  * every kernel against a plain torch fp32 statement of the same op (all cases run, no stop at the first failure),
  * the engine's logits against the general-kernel route (model.forward with a KVCache) step by step, with beam
    re-orderings,
  * timings: each weight-streaming launch in GB/s against the HBM peak, and whole decode steps (tokens/s) at Vicuna-7B
    size for 1 and 5 beams, engine vs general route.
Usage: python scripts/gpu_check_decode.py [--no-time] [--layers N]"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stllm_b200 import ops

dev = "cuda"
fails = []


def report(name, got, ref, atol, rtol):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).sum().item()
    finite = torch.isfinite(got).all().item()
    ok = bad == 0 and finite
    print(f"{'OK  ' if ok else 'FAIL'} {name:70s} max_abs={err.max().item():.3e} ref_max={ref.abs().max().item():.3e} "
          f"bad={bad}/{err.numel()} finite={finite}", flush=True)
    if not ok:
        fails.append(name)
        if err.dim() == 2:
            rows = (err > tol).any(dim=1).nonzero().flatten()
            cols = (err > tol).any(dim=0).nonzero().flatten()
            print(f"     bad rows: n={rows.numel()} first={rows[:8].tolist()}  bad cols: n={cols.numel()} "
                  f"first={cols[:8].tolist()} last={cols[-4:].tolist()}")
            print("     got[0,:8]", got[0, :8].tolist())
            print("     ref[0,:8]", ref[0, :8].tolist())
    return ok


def rms_ref(x, w, eps):
    """w * fp16(x * rsqrt(mean(x^2) + eps))  (modeling_llama_mem.py:61-78)"""
    xf = x.float()
    h = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).half()
    return (w * h) if w is not None else h


def linear_case(R, N, K, norm_mode=0, epi=ops.EPI_NONE, residual=False, alias=False, x_pad=0):
    x_full = torch.randn(R, K + x_pad, device=dev, dtype=torch.float16)
    x = x_full[:, :K]
    w = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.03
    gamma = (1.0 + 0.1 * torch.randn(K, device=dev)).half()
    eps = 1e-6
    n_out = N // 2 if epi == ops.EPI_SWIGLU else N
    res = torch.randn(R, n_out, device=dev, dtype=torch.float16) if residual else None
    if norm_mode == 0:
        ref = x.float() @ w.float().t()
        w_used = w
    elif norm_mode == 1:   # folded: W' = fp16(gamma (.) W), y = rstd * (x W'^T)
        w_used = (w.float() * gamma.float()[None, :]).half()
        rstd = torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + eps)
        ref = rstd * (x.float() @ w_used.float().t())
    else:
        w_used = w
        ref = rms_ref(x, gamma, eps).float() @ w.float().t()
    if epi == ops.EPI_SWIGLU:
        ref = torch.nn.functional.silu(ref[:, 0::2]) * ref[:, 1::2]
    if residual:
        ref = ref + res.float()
    out = res if (alias and residual) else torch.full((R, n_out), float("nan"), device=dev, dtype=torch.float16)
    ops.decode_linear(x, w_used, out, norm_mode=norm_mode, norm_weight=gamma if norm_mode == 2 else None, norm_eps=eps,
                      residual=res, epilogue=epi)
    torch.cuda.synchronize()
    name = f"decode_linear R={R} N={N} K={K} norm={norm_mode} epi={epi} res={int(residual)} alias={int(alias)} pad={x_pad}"
    # fp32 accumulation, one fp16 rounding of the result: north_star's rtol 1e-3 / atol 1e-4 for the plain and folded forms;
    # the explicit RMSNorm form (norm_mode 2) rounds the normalised activations to fp16 first, like the reference: +1 ulp
    atol, rtol = (1e-4, 1e-3) if norm_mode != 2 else (1e-3, 2e-3)
    if epi == ops.EPI_SWIGLU:
        atol = max(atol, 2e-4 * ref.abs().max().item())
    return report(name, out, ref, atol=atol, rtol=rtol)


def attn_setup(R, H, Sp, n_gen, l_max, mask=False, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    D = 128
    HD = H * D
    q = torch.randn(R, 3 * HD, device=dev, dtype=torch.float16, generator=g)
    # prefix lives inside a longer [max_len, HD] slab like the KVCache
    kp_slab = torch.randn(Sp + 7, HD, device=dev, dtype=torch.float16, generator=g)
    vp_slab = torch.randn(Sp + 7, HD, device=dev, dtype=torch.float16, generator=g)
    ks = torch.randn(R, l_max, HD, device=dev, dtype=torch.float16, generator=g)
    vs = torch.randn(R, l_max, HD, device=dev, dtype=torch.float16, generator=g)
    src = torch.randint(0, R, (R, l_max), device=dev, generator=g).to(torch.int32)
    pm = None
    if mask:
        pm = (torch.rand(Sp, device=dev, generator=g) > 0.3).to(torch.uint8)
        pm[0] = 1
    return q, kp_slab[:Sp], vp_slab[:Sp], ks, vs, src, pm


def attn_ref(q, kp, vp, ks, vs, src, pm, R, H, n_keys_suffix, scale):
    D = 128
    out = torch.zeros(R, H * D, device=dev)
    for r in range(R):
        rows = src[r, :n_keys_suffix].long()
        k_suf = ks[rows, torch.arange(n_keys_suffix, device=dev)]   # [n, HD]
        v_suf = vs[rows, torch.arange(n_keys_suffix, device=dev)]
        K = torch.cat([kp, k_suf], 0).float().view(-1, H, D)
        V = torch.cat([vp, v_suf], 0).float().view(-1, H, D)
        qq = q[r, :H * D].float().view(H, D)
        s = torch.einsum("hd,shd->hs", qq, K) * scale
        if pm is not None:
            m = torch.cat([pm.bool(), torch.ones(n_keys_suffix, dtype=torch.bool, device=dev)])
            s = s.masked_fill(~m[None, :], float("-inf"))
        p = s.softmax(-1)
        out[r] = torch.einsum("hs,shd->hd", p, V).reshape(-1)
    return out


def attn_case(R, H, Sp, n_gen, l_max, n_split, mask=False):
    q, kp, vp, ks, vs, src, pm = attn_setup(R, H, Sp, n_gen, l_max, mask)
    ng = torch.tensor([n_gen], device=dev, dtype=torch.int32)
    out = torch.full((R, H * 128), float("nan"), device=dev, dtype=torch.float16)
    ws = ops.decode_attention_workspace(H, n_split, dev)
    scale = 1.0 / math.sqrt(128)
    ops.decode_attention(q, kp, vp, ks, vs, src, ng, out, ws, H, 128, scale, n_split=n_split, prefix_mask=pm)
    torch.cuda.synchronize()
    ref = attn_ref(q, kp, vp, ks, vs, src, pm, R, H, n_gen + 1, scale)
    return report(f"decode_attention R={R} H={H} Sp={Sp} n_gen={n_gen} l_max={l_max} splits={n_split} mask={int(mask)}",
                  out, ref, atol=3e-3, rtol=3e-3)


def attn_fused_case(R, H, Sp, n_gen, l_max, n_split, mask=False, repeats=2, prefetch=False):
    """rope + append + attention + merge in one launch, against the unfused pieces' references; launched `repeats` times on
    the same workspace (the ticket counters must reset themselves)"""
    HD = H * 128
    q, kp, vp, ks, vs, src, pm = attn_setup(R, H, Sp, n_gen, l_max, mask, seed=n_gen + 1)
    qkv = q   # [R, 3*HD] un-roped q | k | v of the new token
    cos_t, sin_t = ops.rope_tables(128, 2048)
    ws = ops.decode_attention_fused_workspace(H, n_split, dev)
    scale = 1.0 / math.sqrt(128)
    ng = torch.tensor([n_gen], device=dev, dtype=torch.int32)
    # reference: prefill RoPE kernel at the same position, append, then the attention reference
    roped = qkv.clone()
    ops.rope_(roped, torch.full((R,), Sp + n_gen, device=dev, dtype=torch.int32), cos_t, sin_t, 2 * H, 128)
    ks_ref, vs_ref, src_ref = ks.clone(), vs.clone(), src.clone()
    ks_ref[:, n_gen] = roped[:, HD:2 * HD]
    vs_ref[:, n_gen] = qkv[:, 2 * HD:]
    src_ref[:, n_gen] = torch.arange(R, device=dev, dtype=torch.int32)
    ref = attn_ref(roped, kp, vp, ks_ref, vs_ref, src_ref, pm, R, H, n_gen + 1, scale)
    ok = True
    qkv_before = qkv.clone()
    for it in range(repeats):
        out = torch.full((R, HD), float("nan"), device=dev, dtype=torch.float16)
        pf = torch.randn(1031, 1024, device=dev).half() if prefetch else None   # ~2 MB, not a multiple of the grid
        ops.decode_attention_fused(qkv, cos_t, sin_t, kp, vp, ks, vs, src, ng, out, ws, H, 128, scale, n_split=n_split,
                                   prefix_mask=pm, prefetch=pf)
        torch.cuda.synchronize()
        ok &= report(f"decode_attention_fused R={R} H={H} Sp={Sp} n_gen={n_gen} l_max={l_max} splits={n_split} "
                     f"mask={int(mask)} run {it}", out, ref, atol=3e-3, rtol=3e-3)
    ok &= report("  fused: k appended (bit-exact with rope_kernel)", ks.reshape(R, -1), ks_ref.reshape(R, -1), 0, 0)
    ok &= report("  fused: v appended", vs.reshape(R, -1), vs_ref.reshape(R, -1), 0, 0)
    ok &= report("  fused: ancestry", src.float(), src_ref.float(), 0, 0)
    ok &= report("  fused: qkv untouched", qkv, qkv_before, 0, 0)
    return ok


def rope_append_case(R, H, prefix_len, n_gen, l_max):
    D = 128
    HD = H * D
    qkv = torch.randn(R, 3 * HD, device=dev, dtype=torch.float16)
    orig = qkv.clone()
    cos_t, sin_t = ops.rope_tables(D, 2048)
    ks = torch.zeros(R, l_max, HD, device=dev, dtype=torch.float16)
    vs = torch.zeros(R, l_max, HD, device=dev, dtype=torch.float16)
    src = torch.full((R, l_max), -1, device=dev, dtype=torch.int32)
    ng = torch.tensor([n_gen], device=dev, dtype=torch.int32)
    ops.decode_rope_append(qkv, cos_t, sin_t, prefix_len, ng, ks, vs, src, H, D)
    torch.cuda.synchronize()
    # reference = the prefill RoPE kernel at the same position (bit-exact: same arithmetic) + plain copies
    ref = orig.clone()
    pos = torch.full((R,), prefix_len + n_gen, device=dev, dtype=torch.int32)
    ops.rope_(ref, pos, cos_t, sin_t, 2 * H, D)
    ok = report(f"decode_rope_append q R={R} H={H} pos={prefix_len + n_gen}", qkv[:, :HD], ref[:, :HD], 0, 0)
    ok &= report("decode_rope_append k -> cache", ks[:, n_gen], ref[:, HD:2 * HD], 0, 0)
    ok &= report("decode_rope_append v -> cache", vs[:, n_gen], orig[:, 2 * HD:], 0, 0)
    exp_src = torch.full((R, l_max), -1, device=dev, dtype=torch.int32)
    exp_src[:, n_gen] = torch.arange(R, device=dev, dtype=torch.int32)
    ok &= report("decode_rope_append ancestry", src.float(), exp_src.float(), 0, 0)
    untouched = ks.clone()
    untouched[:, n_gen] = 0
    ok &= report("decode_rope_append leaves other slots", untouched.reshape(R, -1), torch.zeros_like(untouched).reshape(R, -1), 0, 0)
    return ok


def engine_vs_general(llama_layers, hidden, inter, heads, vocab, S, beams, steps, use_graph=True, mask=False):
    """same tokens through the decode engine and through model.forward + KVCache (general kernels); beams re-ordered
    with a fixed permutation pattern every step"""
    from stllm_b200.decode import DecodeEngine
    from stllm_b200.pipeline import build_model
    model = build_model(llama_layers=llama_layers, vit_depth=1, qformer_layers=1, vocab_size=vocab, llama_hidden=hidden,
                        llama_inter=inter, llama_heads=heads, device=torch.device("cuda", 0), seed=3)
    for lyr in model.get_model().layers:   # non-trivial norm weights so the folding is exercised
        lyr.input_layernorm.weight.data.uniform_(0.5, 1.5)
        lyr.post_attention_layernorm.weight.data.uniform_(0.5, 1.5)
    model.get_model().norm.weight.data.uniform_(0.5, 1.5)
    model.get_model().repack()
    g = torch.Generator(device=dev).manual_seed(11)
    emb = (torch.randn(1, S, hidden, device=dev, generator=g) * 0.05).half()
    am = None
    if mask:
        am = torch.ones(1, S, dtype=torch.long, device=dev)
        am[0, S // 3] = 0
        am[0, S // 2] = 0
    out = model(samples=None, inputs_embeds=emb, attention_mask=am, use_cache=True, cache_max_len=S + steps + 1)
    cache_e = out.past_key_values
    eng = DecodeEngine(model, cache_e, S, beams, steps, prefix_mask=am, use_graph=use_graph)
    out2 = model(samples=None, inputs_embeds=emb, attention_mask=am, use_cache=True, cache_max_len=S + steps + 1)
    cache_g = out2.past_key_values
    cache_g.expand_batch(beams)
    ok = True
    worst = 0.0
    for t in range(steps):
        tok = torch.randint(3, vocab, (beams,), device=dev, generator=g)
        if t > 0 and beams > 1:
            idx = torch.tensor([(b * 2 + t) % beams for b in range(beams)], device=dev)
            eng.reorder(idx)
            cache_g.reorder(idx)
        le = eng.step(tok).float().clone()
        lg = model(samples=None, input_ids=tok[:, None], past_key_values=cache_g, use_cache=True).logits[:, -1].float()
        rel = ((le - lg).norm() / lg.norm()).item()
        worst = max(worst, rel)
        ok &= report(f"engine vs general L={llama_layers} C={hidden} beams={beams} graph={int(use_graph)} "
                     f"mask={int(mask)} step {t} (rel_fro {rel:.2e})", le, lg, atol=3e-2, rtol=2e-2)
    del model
    torch.cuda.empty_cache()
    return ok, worst


def time_fn(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def time_linear(R, N, K, norm_mode=0, epi=ops.EPI_NONE, residual=False, copies=24):
    """weights rotate over `copies` buffers (> L2) so every launch streams from HBM"""
    ws = [torch.randn(N, K, device=dev, dtype=torch.float16) * 0.03 for _ in range(copies)]
    x = torch.randn(R, K, device=dev, dtype=torch.float16)
    n_out = N // 2 if epi == ops.EPI_SWIGLU else N
    out = torch.zeros(R, n_out, device=dev, dtype=torch.float16)
    res = out if residual else None
    gamma = torch.ones(K, device=dev, dtype=torch.float16)
    i = [0]

    def fn():
        ops.decode_linear(x, ws[i[0] % copies], out, norm_mode=norm_mode, norm_weight=gamma if norm_mode == 2 else None,
                          norm_eps=1e-6, residual=res, epilogue=epi)
        i[0] += 1
    us = time_fn(fn, iters=3 * copies, warm=copies)
    gb = N * K * 2 / 1e9

    def fn2():
        ops.gemm(x, ws[i[0] % copies], out=None)
        i[0] += 1
    us2 = time_fn(fn2, iters=3 * copies, warm=copies) if epi == ops.EPI_NONE else float("nan")
    print(f"TIME decode_linear R={R} N={N} K={K} norm={norm_mode} epi={epi}: {us:.1f} us = {gb / us * 1e6:.0f} GB/s"
          f"   (tcgen05 GEMM on the same shape: {us2:.1f} us)", flush=True)
    return dict(R=R, N=N, K=K, us=us, gbs=gb / us * 1e6, gemm_us=us2)


def time_decode(layers, beams, S=576, steps=48, general=True):
    from stllm_b200.decode import DecodeEngine
    from stllm_b200.pipeline import build_model
    model = build_model(llama_layers=layers, vit_depth=1, qformer_layers=1, device=torch.device("cuda", 0), seed=0)
    g = torch.Generator(device=dev).manual_seed(5)
    emb = (torch.randn(1, S, 4096, device=dev, generator=g) * 0.05).half()
    res = dict(layers=layers, beams=beams, prefix=S, steps=steps)
    out = model(samples=None, inputs_embeds=emb, use_cache=True, cache_max_len=S + steps + 8)
    eng = DecodeEngine(model, out.past_key_values, S, beams, steps + 8)
    tok = torch.randint(3, 32000, (beams,), device=dev, generator=g)
    idx = torch.arange(beams, device=dev).flip(0)
    for _ in range(4):
        eng.step(tok)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        if beams > 1:
            eng.reorder(idx)
        eng.step(tok)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    cfg = model.config
    wbytes = layers * (4 * 4096 * 4096 + 3 * 4096 * cfg.intermediate_size) * 2 + cfg.vocab_size * 4096 * 2
    res.update(engine_ms_per_token=ms, engine_tokens_per_s=1e3 / ms, weight_gb=wbytes / 1e9,
               engine_weight_gbs=wbytes / 1e9 / (ms * 1e-3), launches_per_step=eng.launches_per_step)
    if general:
        out = model(samples=None, inputs_embeds=emb, use_cache=True, cache_max_len=S + steps + 8)
        cache = out.past_key_values
        cache.expand_batch(beams)
        for _ in range(3):
            model(samples=None, input_ids=tok[:, None], past_key_values=cache, use_cache=True)
        torch.cuda.synchronize()
        n2 = min(steps, 16)
        t0 = time.time()
        for _ in range(n2):
            if beams > 1:
                cache.reorder(idx)
            model(samples=None, input_ids=tok[:, None], past_key_values=cache, use_cache=True)
        torch.cuda.synchronize()
        res["general_ms_per_token"] = (time.time() - t0) / n2 * 1e3
    print("TIME decode " + json.dumps(res), flush=True)
    del model, eng
    torch.cuda.empty_cache()
    return res


AB_VARIANTS = (("default", {}), ("unfused_attention", {"STLLM_DECODE_FUSED_ATTN": "0"}),
               ("no_weight_prefetch", {"STLLM_DECODE_PREFETCH": "0"}), ("prefetch_next_l2", {"STLLM_DECODE_PREFETCH_NEXT": "1"}),
               ("xdirect", {"STLLM_DECODE_XDIRECT": "1"}), ("splits_3", {"STLLM_DECODE_SPLITS": "3"}))


def time_decode_ab(layers, beams, S=576, steps=32, rounds=3):
    """A/B of the engine's switches inside ONE process on ONE model: every variant captures its own graph; timing rounds
    alternate between the variants (boxes and clocks differ from run to run, so only same-process numbers compare)"""
    from stllm_b200.decode import DecodeEngine
    from stllm_b200.pipeline import build_model
    model = build_model(llama_layers=layers, vit_depth=1, qformer_layers=1, device=torch.device("cuda", 0), seed=0)
    g = torch.Generator(device=dev).manual_seed(5)
    emb = (torch.randn(1, S, 4096, device=dev, generator=g) * 0.05).half()
    lmax = 8 + steps * rounds + 8
    out = model(samples=None, inputs_embeds=emb, use_cache=True, cache_max_len=S + lmax)
    tok = torch.randint(3, 32000, (beams,), device=dev, generator=g)
    idx = torch.arange(beams, device=dev).flip(0)
    engines = []
    for name, env in AB_VARIANTS:
        os.environ.update(env)
        engines.append((name, DecodeEngine(model, out.past_key_values, S, beams, lmax)))
        for k in env:
            os.environ.pop(k)
    res = {name: [] for name, _ in engines}
    for name, eng in engines:
        for _ in range(8):
            eng.step(tok)
    for _ in range(rounds):
        for name, eng in engines:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                if beams > 1:
                    eng.reorder(idx)
                eng.step(tok)
            e1.record()
            torch.cuda.synchronize()
            res[name].append(round(e0.elapsed_time(e1) / steps, 4))
    cfg = model.config
    wbytes = layers * (4 * 4096 * 4096 + 3 * 4096 * cfg.intermediate_size) * 2 + cfg.vocab_size * 4096 * 2
    summary = dict(layers=layers, beams=beams, prefix=S, weight_gb=wbytes / 1e9, ms_per_token=res,
                   best_gbs={k: round(wbytes / 1e9 / (min(v) * 1e-3)) for k, v in res.items()})
    print("TIME decode_ab " + json.dumps(summary), flush=True)
    del engines, model
    torch.cuda.empty_cache()
    return summary


def time_decode_skip(layers, beams, S=576, steps=32, rounds=2):
    """in-graph cost of each kernel kind: the step timed with one kind left out (debug_skip), same process, alternating"""
    from stllm_b200.decode import DecodeEngine
    from stllm_b200.pipeline import build_model
    model = build_model(llama_layers=layers, vit_depth=1, qformer_layers=1, device=torch.device("cuda", 0), seed=0)
    g = torch.Generator(device=dev).manual_seed(5)
    emb = (torch.randn(1, S, 4096, device=dev, generator=g) * 0.05).half()
    lmax = 8 + steps * rounds + 8
    out = model(samples=None, inputs_embeds=emb, use_cache=True, cache_max_len=S + lmax)
    tok = torch.randint(3, 32000, (beams,), device=dev, generator=g)
    idx = torch.arange(beams, device=dev).flip(0)
    kinds = ("", "qkv", "attn", "o", "gu", "down")
    engines = [(k or "full", DecodeEngine(model, out.past_key_values, S, beams, lmax, debug_skip=(k,) if k else ()))
               for k in kinds]
    res = {name: [] for name, _ in engines}
    for name, eng in engines:
        for _ in range(8):
            eng.step(tok)
    for _ in range(rounds):
        for name, eng in engines:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                if beams > 1:
                    eng.reorder(idx)
                eng.step(tok)
            e1.record()
            torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / steps)
    full = min(res["full"])
    per_launch = {k: round((full - min(v)) / layers * 1e3, 2) for k, v in res.items() if k != "full"}
    summary = dict(layers=layers, beams=beams, full_ms=round(full, 4), in_graph_us_per_launch=per_launch)
    print("TIME decode_skip " + json.dumps(summary), flush=True)
    del engines, model
    torch.cuda.empty_cache()
    return summary


def time_generate(layers, beams, S=576, new_tokens=64):
    """the public call: llm.generate(inputs_embeds=[1,S,4096], ...) end to end (prefill + graph capture + decode steps +
    host-side beam bookkeeping), decode engine vs the general-kernel route"""
    from stllm_b200.pipeline import build_model
    model = build_model(llama_layers=layers, vit_depth=1, qformer_layers=1, device=torch.device("cuda", 0), seed=0)
    g = torch.Generator(device=dev).manual_seed(5)
    emb = (torch.randn(1, S, 4096, device=dev, generator=g) * 0.05).half()
    res = dict(layers=layers, beams=beams, prefix=S, new_tokens=new_tokens)
    for route in ("engine", "general"):
        os.environ["STLLM_DECODE_ENGINE"] = "1" if route == "engine" else "0"
        n = new_tokens if route == "engine" else min(new_tokens, 24)
        model.generate(inputs_embeds=emb, max_new_tokens=4, num_beams=beams, eos_token_id=-1)   # warm-up
        torch.cuda.synchronize()
        t0 = time.time()
        out = model.generate(inputs_embeds=emb, max_new_tokens=n, num_beams=beams, eos_token_id=-1)
        torch.cuda.synchronize()
        dt = time.time() - t0
        res[route + "_s"] = dt
        res[route + "_tokens"] = int(out.shape[1])
        res[route + "_tokens_per_s_incl_prefill"] = out.shape[1] / dt
    os.environ.pop("STLLM_DECODE_ENGINE", None)
    print("TIME generate " + json.dumps(res), flush=True)
    del model
    torch.cuda.empty_cache()
    return res


CASES = []     # (name, callable -> bool); tests/test_decode_gpu.py parametrises over them


def _c(fn, *a, **kw):
    name = f"{len(CASES):03d}:{fn.__name__} " + " ".join(str(x) for x in a) + "".join(f" {k}={v}" for k, v in kw.items())

    def run():
        torch.manual_seed(len(name))
        return bool(fn(*a, **kw))
    CASES.append((name, run))


def build_cases():
    if CASES:
        return CASES
    # ---- decode_linear: tiny first ----
    _c(linear_case, 1, 16, 32)
    _c(linear_case, 1, 16, 64)
    _c(linear_case, 3, 48, 96)                       # K*2 % 128 == 64: no smem padding
    _c(linear_case, 8, 256, 256)
    _c(linear_case, 5, 200, 512)                     # N not a multiple of 16
    _c(linear_case, 5, 100, 512, residual=True)
    _c(linear_case, 2, 768, 256, norm_mode=1)
    _c(linear_case, 2, 768, 256, norm_mode=2)
    _c(linear_case, 2, 1024, 256, norm_mode=1, epi=ops.EPI_SWIGLU)
    _c(linear_case, 2, 256, 512, residual=True, alias=True)
    _c(linear_case, 4, 128, 256, norm_mode=2, x_pad=8)
    # Vicuna-7B shapes
    _c(linear_case, 5, 12288, 4096, norm_mode=1)
    _c(linear_case, 5, 12288, 4096, norm_mode=2)
    _c(linear_case, 1, 4096, 4096, residual=True, alias=True)
    _c(linear_case, 5, 22016, 4096, norm_mode=1, epi=ops.EPI_SWIGLU)
    _c(linear_case, 8, 22016, 4096, norm_mode=2, epi=ops.EPI_SWIGLU)
    _c(linear_case, 5, 4096, 11008, residual=True, alias=True)
    _c(linear_case, 8, 4096, 11008, residual=True)
    _c(linear_case, 5, 32000, 4096, norm_mode=2)
    _c(linear_case, 1, 32001, 4096, norm_mode=2)     # vocab with the added [PAD] (st_llm.py:306-310)
    # ---- rope + cache append ----
    _c(rope_append_case, 5, 32, 576, 0, 16)
    _c(rope_append_case, 1, 2, 12, 7, 8)
    _c(rope_append_case, 8, 32, 100, 15, 16)
    # ---- attention ----
    _c(attn_case, 1, 2, 12, 0, 8, 1)
    _c(attn_case, 1, 2, 12, 3, 8, 4)
    _c(attn_case, 3, 2, 5, 6, 8, 4)                   # fewer prefix keys than splits*... (empty slices)
    _c(attn_case, 5, 32, 576, 0, 64, 4)
    _c(attn_case, 5, 32, 576, 37, 64, 4)
    _c(attn_case, 5, 32, 577, 63, 64, 3, mask=True)
    _c(attn_case, 8, 32, 1500, 299, 300, 4)
    _c(attn_case, 2, 4, 0, 5, 8, 2)                   # no prefix at all
    _c(attn_case, 5, 32, 576, 20, 64, 1)
    # ---- fused rope + append + attention + merge ----
    _c(attn_fused_case, 1, 2, 12, 0, 8, 1)
    _c(attn_fused_case, 1, 2, 12, 3, 8, 4)
    _c(attn_fused_case, 3, 2, 5, 6, 8, 4)
    _c(attn_fused_case, 5, 32, 576, 0, 64, 5)
    _c(attn_fused_case, 5, 32, 576, 37, 64, 5, repeats=3)
    _c(attn_fused_case, 5, 32, 576, 38, 64, 4, prefetch=True)
    _c(attn_fused_case, 5, 32, 577, 63, 64, 3, mask=True)      # > 128 prompt keys per split: second batch of loads
    _c(attn_fused_case, 8, 32, 1500, 299, 300, 12)
    _c(attn_fused_case, 8, 32, 1500, 299, 300, 4)
    _c(attn_fused_case, 2, 4, 0, 5, 8, 2)
    _c(attn_fused_case, 1, 32, 576, 150, 256, 5)               # greedy: all 16 warps on one beam's generated keys
    _c(attn_fused_case, 5, 32, 576, 20, 64, 1)
    # ---- engine vs the general kernels (model.forward + KVCache) ----
    _c(engine_vs_general, 2, 256, 512, 2, 128, 12, 1, 5, use_graph=False)
    _c(engine_vs_general, 2, 256, 512, 2, 128, 12, 3, 6, use_graph=True, mask=True)
    _c(engine_vs_general, 2, 4096, 11008, 32, 32000, 64, 5, 6, use_graph=True)
    return CASES


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    torch.manual_seed(0)
    t0 = time.time()
    if "--time-only" in sys.argv:
        return timings()
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    for name, fn in build_cases():
        if only and not any(o in name for o in only):
            continue
        fn()
    print(f"decode checks done in {time.time() - t0:.1f}s; failures: {len(fails)}", flush=True)
    for f in fails:
        print("  FAILED:", f)
    if "--no-time" not in sys.argv:
        timings()
    return 1 if fails else 0


def timings():
    if True:
        rows = []
        for R in (1, 5):
            rows.append(time_linear(R, 12288, 4096, norm_mode=1))
            rows.append(time_linear(R, 4096, 4096, residual=True))
            rows.append(time_linear(R, 22016, 4096, norm_mode=1, epi=ops.EPI_SWIGLU))
            rows.append(time_linear(R, 4096, 11008, residual=True))
            rows.append(time_linear(R, 32000, 4096, norm_mode=2, copies=12))
        layers = 32
        if "--layers" in sys.argv:
            layers = int(sys.argv[sys.argv.index("--layers") + 1])
        dec = [time_decode(layers, 1), time_decode(layers, 5)]
        if "--ab" in sys.argv:
            dec += [time_decode_ab(layers, 1), time_decode_ab(layers, 5)]
        if "--generate" in sys.argv:
            dec += [time_generate(layers, 1), time_generate(layers, 5)]
        os.makedirs("gpurun_out", exist_ok=True)
        tag = os.environ.get("DECODE_TIMING_TAG", "")
        with open(f"gpurun_out/decode_timing{tag}.json", "w") as f:
            json.dump(dict(linear=rows, decode=dec), f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
