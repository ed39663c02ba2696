"""Decode engine: the cached single-token steps of `generate`.

The reference decodes through HF `GenerationMixin` (conversation.py:231-243, mvbench_conversation.py:114-125,
max_new_tokens 100-300, num_beams 1 or 5): every step re-runs LlamaForCausalLM.forward on `num_beams` rows with a KV
cache that holds `num_beams` copies of the prompt and is index_selected as a whole after every beam re-ordering
(modeling_llama_mem.py:205-217,715-752).  Here a step is

    embedding gather -> 32 x [ (RMSNorm) q|k|v -> {RoPE + cache append + attention over (shared prompt KV || per-beam
    suffix) + split merge: one launch} -> o_proj + residual -> (RMSNorm) gate|up SwiGLU -> down + residual ] -> (RMSNorm) lm_head

on the HBM-bound kernels of csrc/decode.cu (weights streamed once per token for all beams), captured ONCE into a CUDA graph
and replayed per token; the position / suffix length lives in device memory.  The prompt's K/V (written by the prefill
kernels into the KVCache) is read in place, one copy for all beams; generated K/V sits in per-layer suffix buffers addressed
through an ancestry table, so a beam re-ordering permutes `beams x steps` ints.
"""
import os

import torch

from . import _lib, ops


def supported(llm, n_rows):
    """shapes the decode kernels take: head dim 128, K multiples of 32, activations of the widest layer fit in smem"""
    cfg = llm.config
    C, F, H = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads
    if C % H or C // H != 128 or C % 32 or F % 32 or not 1 <= n_rows <= ops.DECODE_MAX_ROWS:
        return False
    return n_rows * (max(C, F) + 32) * 2 + 17 * 1024 <= 220 * 1024


class DecodeEngine:
    def __init__(self, llm, cache, prefix_len, n_rows, max_steps, prefix_mask=None, use_graph=None, n_split=None,
                 debug_skip=()):
        """llm: STLLMForCausalLM; cache: the KVCache the prefill filled (batch 1); n_rows: beams (1 = greedy/sampling);
        max_steps: decode steps that may follow."""
        assert supported(llm, n_rows)
        model = llm.get_model()
        if not model._packed_valid():
            model._pack()
        self.P = model._packed
        cfg = llm.config
        self.cfg = cfg
        self.R, self.Sp, self.Lmax = n_rows, int(prefix_len), max(int(max_steps), 1)
        self.C, self.H = cfg.hidden_size, cfg.num_attention_heads
        self.d = self.C // self.H
        self.eps = cfg.rms_norm_eps
        dev = llm.lm_head.weight.device
        self.dev = dev
        assert self.Sp + self.Lmax <= cfg.max_position_embeddings
        ew = model.embed_tokens.weight
        self.embed_w = ew if ew.dtype == torch.float16 else ew.half()
        lw = llm.lm_head.weight
        self.lm_w = (lw if lw.dtype == torch.float16 else lw.half()).contiguous()
        self.cos_t, self.sin_t = model._rope_tables(dev)
        nL = len(self.P["layers"])
        C, R, F = self.C, self.R, cfg.intermediate_size
        f16 = dict(dtype=torch.float16, device=dev)
        self.tok = torch.zeros(R, dtype=torch.int64, device=dev)
        self.x = torch.zeros(R, C, **f16)
        self.qkv = torch.zeros(R, 3 * C, **f16)
        self.a = torch.zeros(R, C, **f16)
        self.g = torch.zeros(R, F, **f16)
        self.logits = torch.zeros(R, cfg.vocab_size, **f16)
        self.n_gen = torch.zeros(1, dtype=torch.int32, device=dev)
        self.src = torch.zeros(R, self.Lmax, dtype=torch.int32, device=dev)
        self.ks = torch.empty(nL, R, self.Lmax, C, **f16)   # every slot is written before it is read
        self.vs = torch.empty(nL, R, self.Lmax, C, **f16)
        # the prompt's K/V stay where the prefill wrote them: [Sp, H*d] rows of batch element 0
        self.kp = [cache.k[li][0, :self.Sp].reshape(self.Sp, C) for li in range(nL)]
        self.vp = [cache.v[li][0, :self.Sp].reshape(self.Sp, C) for li in range(nL)]
        self.prefix_mask = None
        if prefix_mask is not None and not bool(prefix_mask.all()):
            self.prefix_mask = prefix_mask.reshape(-1)[:self.Sp].to(torch.uint8).contiguous()
        # key splits: <= 128 prompt keys per CTA (all of a CTA's prompt K/V is then requested in one batch of loads), but
        # never more CTAs than SMs -- the attention CTAs are one per SM, a second wave costs more than longer slices
        # (measured: 4 splits 2.95 ms/token, 5 splits 3.09, 8 splits 3.13 at 32 heads, 576 prompt keys)
        if n_split is None:
            n_split = int(os.environ.get("STLLM_DECODE_SPLITS", "0"))
            if n_split <= 0:
                sms = int(_lib.load().stllm_device_sm_count())
                n_split = min((self.Sp + 127) // 128, max(1, sms // self.H))
        self.n_split = min(max(1, n_split), 16)
        self.fused_attn = os.environ.get("STLLM_DECODE_FUSED_ATTN", "1") != "0"
        # experiment, OFF: let the attention launch (few bytes, ~12 us) pull the o_proj weights (33.5 MB << 126 MB L2) and the
        # head of gate|up into L2.  Same-process A/B: 2.671 (on) vs 2.664 ms/token (off) at 1 beam, 3.33 vs 3.28 at 5.
        self.prefetch_next = os.environ.get("STLLM_DECODE_PREFETCH_NEXT", "0") != "0"
        # ... and the head of the gate|up weights (rows are consumed in this order by the first wave of the GEMV)
        gu_mb = float(os.environ.get("STLLM_DECODE_PREFETCH_GU_MB", "48")) if self.prefetch_next else 0.0
        self.prefetch_gu_rows = min(2 * F, int(gu_mb * 1e6 / (2 * C))) // 16 * 16
        if self.fused_attn:
            self.ws = ops.decode_attention_fused_workspace(self.H, self.n_split, dev)
        else:
            self.ws = ops.decode_attention_workspace(self.H, self.n_split, dev)
        self.steps_done = 0
        self.graph = None
        self.debug_skip = frozenset(debug_skip)   # timing studies only: kernel kinds left out of the step
        if use_graph is None:
            use_graph = os.environ.get("STLLM_DECODE_GRAPH", "1") != "0"
        self.launches_per_step = None
        if use_graph:
            self._capture()
        if self.R > 1:
            # first use of the re-ordering path loads torch's index_select / copy kernels (CUDA lazy module loading, ~40 ms
            # once per process): do it here, not inside the first beam step.  This one-off cost -- landing in the first 10
            # steps of the first 5-beam engine of a process -- was round 1's "bimodal" 5-beam timing (3.1 vs 3.9-4.4 ms/token
            # over 48 steps; profiles/r02_decode_transient.md).
            self.reorder(torch.arange(self.R, device=dev))

    # one token for every row: self.tok -> self.logits; device-side position counter advanced at the end
    def _step_impl(self):
        P, R, C = self.P, self.R, self.C
        fuse = P["fuse"]
        x = ops.gather_rows(self.embed_w, self.tok, out=self.x)
        scale = 1.0 / (self.d ** 0.5)
        for li, L in enumerate(P["layers"]):
            skip = self.debug_skip
            if "qkv" in skip:
                pass
            elif fuse and li > 0:
                ops.decode_linear(x, L["w_qkv"], self.qkv, norm_mode=ops.NORM_RMS, norm_eps=self.eps)
            else:
                ops.decode_linear(x, L["w_qkv"], self.qkv, norm_mode=2, norm_weight=L["ln1"], norm_eps=self.eps)
            if "attn" in skip:
                pass
            elif self.fused_attn:
                ops.decode_attention_fused(self.qkv, self.cos_t, self.sin_t, self.kp[li], self.vp[li], self.ks[li],
                                           self.vs[li], self.src, self.n_gen, self.a, self.ws, self.H, self.d, scale,
                                           n_split=self.n_split, prefix_mask=self.prefix_mask,
                                           prefetch=L["w_o"] if self.prefetch_next else None,
                                           prefetch2=L["w_gu"][:self.prefetch_gu_rows] if self.prefetch_gu_rows else None)
            else:
                ops.decode_rope_append(self.qkv, self.cos_t, self.sin_t, self.Sp, self.n_gen, self.ks[li],
                                       self.vs[li], self.src, self.H, self.d)
                ops.decode_attention(self.qkv, self.kp[li], self.vp[li], self.ks[li], self.vs[li], self.src,
                                     self.n_gen, self.a, self.ws, self.H, self.d, scale, n_split=self.n_split,
                                     prefix_mask=self.prefix_mask)
            if "o" not in skip:
                ops.decode_linear(self.a, L["w_o"], x, residual=x)
            if "gu" in skip:
                pass
            elif fuse:
                ops.decode_linear(x, L["w_gu"], self.g, norm_mode=ops.NORM_RMS, norm_eps=self.eps,
                                  epilogue=ops.EPI_SWIGLU)
            else:
                ops.decode_linear(x, L["w_gu"], self.g, norm_mode=2, norm_weight=L["ln2"], norm_eps=self.eps,
                                  epilogue=ops.EPI_SWIGLU)
            if "down" not in skip:
                ops.decode_linear(self.g, L["w_down"], x, residual=x)
        ops.decode_linear(x, self.lm_w, self.logits, norm_mode=2, norm_weight=P["norm"], norm_eps=self.eps)
        self.n_gen.add_(1)

    def _capture(self):
        # one eager step first: kernel attributes are configured outside the capture; its writes (suffix slot 0) are
        # overwritten by the first real step
        c0 = ops.launch_count
        self._step_impl()
        self.launches_per_step = ops.launch_count - c0
        self.n_gen.zero_()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._step_impl()
        torch.cuda.current_stream().wait_stream(s)
        self.n_gen.zero_()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            self._step_impl()
        self.n_gen.zero_()
        torch.cuda.synchronize()
        self.graph = g

    def step(self, tokens):
        """tokens: int64 [R] (device).  Returns the fp16 logits [R, V] of the next position (a static buffer: consume it
        before the next step)."""
        assert self.steps_done < self.Lmax, "decode engine: more steps than the suffix cache was sized for"
        self.tok.copy_(tokens.reshape(-1))
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_impl()
        self.steps_done += 1
        return self.logits

    def reorder(self, beam_idx):
        """beam b continues the history of beam beam_idx[b]: permute the ancestry table (no K/V is moved)"""
        self.src.copy_(self.src.index_select(0, beam_idx.to(torch.int64)))
