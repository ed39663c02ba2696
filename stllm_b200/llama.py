"""Vicuna / Llama decoder (prefill + cached decode) on the sm_100a kernels.

The reference uses HuggingFace transformers==4.28.0 `LlamaModel` / `LlamaForCausalLM` (st_llm.py:20,45,95); its
in-tree statement of the same math is stllm/models/modeling_llama_mem.py.  This file keeps the HF module tree and
state-dict keys (model.embed_tokens, model.layers.N.{self_attn.{q,k,v,o}_proj, mlp.{gate,up,down}_proj,
input_layernorm, post_attention_layernorm}, model.norm, lm_head) and runs:
  RMSNorm -> fused q|k|v GEMM -> RoPE in place -> causal flash attention (fp32 softmax, key-padding mask)
  -> o_proj GEMM (+residual) -> RMSNorm -> gate|up GEMM with SwiGLU epilogue (interleaved rows) -> down GEMM
  (+residual); final RMSNorm; lm_head GEMM.                     modeling_llama_mem.py:61-144,172-316,464-594
"""
import os

import torch
import torch.nn as nn

from . import ops
from .packing import PackedWeightsMixin


class LlamaConfig:
    model_type = "llama"

    def __init__(self, vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                 num_attention_heads=32, max_position_embeddings=2048, rms_norm_eps=1e-6, initializer_range=0.02,
                 pad_token_id=0, bos_token_id=1, eos_token_id=2, rope_theta=10000.0, **unused):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.max_position_embeddings = max_position_embeddings
        self.rms_norm_eps = rms_norm_eps
        self.initializer_range = initializer_range
        self.pad_token_id = pad_token_id
        self.bos_token_id = bos_token_id
        self.eos_token_id = eos_token_id
        self.rope_theta = rope_theta
        self.use_return_dict = True

    def get(self, k, default=None):
        return getattr(self, k, default)


class LlamaRMSNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps


class LlamaAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        h = config.hidden_size
        self.q_proj = nn.Linear(h, h, bias=False)
        self.k_proj = nn.Linear(h, h, bias=False)
        self.v_proj = nn.Linear(h, h, bias=False)
        self.o_proj = nn.Linear(h, h, bias=False)


class LlamaMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.gate_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.down_proj = nn.Linear(config.intermediate_size, config.hidden_size, bias=False)
        self.up_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)


class LlamaDecoderLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self_attn = LlamaAttention(config)
        self.mlp = LlamaMLP(config)
        self.input_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)


class BaseModelOutputWithPast:
    def __init__(self, last_hidden_state, past_key_values=None, hidden_states=None, attentions=None):
        self.last_hidden_state = last_hidden_state
        self.past_key_values = past_key_values
        self.hidden_states = hidden_states
        self.attentions = attentions

    def __getitem__(self, i):
        return (self.last_hidden_state, self.past_key_values, self.hidden_states, self.attentions)[i]


class CausalLMOutputWithPast:
    def __init__(self, loss=None, logits=None, past_key_values=None, hidden_states=None, attentions=None):
        self.loss = loss
        self.logits = logits
        self.past_key_values = past_key_values
        self.hidden_states = hidden_states
        self.attentions = attentions


class KVCache:
    """contiguous per-layer K/V cache [B, max_len, H, D] (fp16), filled by prefill and appended by decode steps"""

    def __init__(self, n_layers, batch, max_len, n_heads, head_dim, device):
        self.k = [torch.empty(batch, max_len, n_heads, head_dim, dtype=torch.float16, device=device)
                  for _ in range(n_layers)]
        self.v = [torch.empty(batch, max_len, n_heads, head_dim, dtype=torch.float16, device=device)
                  for _ in range(n_layers)]
        self.length = 0
        self.max_len = max_len
        self.key_mask = torch.ones(batch, max_len, dtype=torch.uint8, device=device)

    def expand_batch(self, n):
        """share one prefilled prompt across n beams"""
        for i in range(len(self.k)):
            self.k[i] = self.k[i].expand(n, -1, -1, -1).contiguous()
            self.v[i] = self.v[i].expand(n, -1, -1, -1).contiguous()
        self.key_mask = self.key_mask.expand(n, -1).contiguous()

    def reorder(self, beam_idx):
        for i in range(len(self.k)):
            self.k[i] = self.k[i].index_select(0, beam_idx)
            self.v[i] = self.v[i].index_select(0, beam_idx)
        self.key_mask = self.key_mask.index_select(0, beam_idx)


class LlamaModel(PackedWeightsMixin, nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.padding_idx = config.pad_token_id
        self.vocab_size = config.vocab_size
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size)
        self.layers = nn.ModuleList([LlamaDecoderLayer(config) for _ in range(config.num_hidden_layers)])
        self.norm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self._packed = None
        self._rope = None
        std = config.initializer_range
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                m.weight.data.normal_(mean=0.0, std=std)

    # ---- packing ----
    def _apply(self, fn, *a, **kw):
        self._packed = None
        self._rope = None
        return super()._apply(fn, *a, **kw)

    def _pack(self):
        dev = self.norm.weight.device
        f16 = lambda t: t.detach().to(device=dev, dtype=torch.float16).contiguous()
        # RMSNorm folded into the consuming GEMM (ops.gemm norm=NORM_RMS): the packed q|k|v and gate|up copies carry
        # the norm weight, W' = fp16(w_norm (.) W); the row scale rsqrt(mean(x^2) + eps) is applied in the epilogue from
        # the sum of squares the producing o_proj / down_proj GEMM accumulated.  Layer 0's input norm has no producing
        # GEMM (inputs_embeds) and the final norm's output is returned to the caller: both stay standalone kernels.
        fuse = os.environ.get("STLLM_FUSE_NORM", "1") != "0"

        def fold(w, g):
            return (w.float() * g.float()[None, :]).to(torch.float16).contiguous()
        layers = []
        for li, lyr in enumerate(self.layers):
            a, m = lyr.self_attn, lyr.mlp
            w_qkv = torch.cat([f16(a.q_proj.weight), f16(a.k_proj.weight), f16(a.v_proj.weight)], 0)
            # interleave rows: 2j = gate_j, 2j+1 = up_j  (SwiGLU epilogue pairs adjacent accumulator columns)
            w_gu = torch.stack([f16(m.gate_proj.weight), f16(m.up_proj.weight)], dim=1).reshape(
                2 * m.gate_proj.weight.shape[0], m.gate_proj.weight.shape[1]).contiguous()
            ln1, ln2 = f16(lyr.input_layernorm.weight), f16(lyr.post_attention_layernorm.weight)
            if fuse:
                if li > 0:
                    w_qkv = fold(w_qkv, ln1)
                w_gu = fold(w_gu, ln2)
            layers.append(dict(w_qkv=w_qkv, w_o=f16(a.o_proj.weight), w_gu=w_gu, w_down=f16(m.down_proj.weight),
                               ln1=ln1, ln2=ln2))
        self._packed = dict(layers=layers, norm=f16(self.norm.weight), fuse=fuse)
        self._mark_packed()

    def _rope_tables(self, device):
        if self._rope is None or self._rope[0].device != device:
            d = self.config.hidden_size // self.config.num_attention_heads
            self._rope = ops.rope_tables(d, self.config.max_position_embeddings, self.config.rope_theta, device)
        return self._rope

    def get_input_embeddings(self):
        return self.embed_tokens

    def _embed_ids(self, input_ids):
        w = self.embed_tokens.weight
        return ops.gather_rows(w if w.dtype == torch.float16 else w.half(), input_ids.reshape(-1)).view(
            tuple(input_ids.shape) + (w.shape[1],))

    # ---- forward ----
    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, use_cache=False,
                output_hidden_states=False, return_dict=True, past_key_values=None, cache_max_len=None, **unused):
        """LlamaModel.forward (modeling_llama_mem.py:464-594).  attention_mask [B, S_total] 1 = token, 0 = pad
        (right padding as st_llm.py builds it).  past_key_values: a KVCache for decode steps."""
        if inputs_embeds is None:
            assert input_ids is not None, "You have to specify either input_ids or inputs_embeds"
            w = self.embed_tokens.weight
            inputs_embeds = ops.gather_rows(w if w.dtype == torch.float16 else w.half(), input_ids.reshape(-1)).view(
                input_ids.shape + (w.shape[1],))
        if not self._packed_valid():
            self._pack()
        P = self._packed
        cfg = self.config
        B, S, C = inputs_embeds.shape
        H = cfg.num_attention_heads
        d = C // H
        dev = inputs_embeds.device
        cos_t, sin_t = self._rope_tables(dev)
        cache = past_key_values
        past = cache.length if cache is not None else 0
        assert past + S <= cfg.max_position_embeddings, "sequence longer than max_position_embeddings"
        pos = (torch.arange(S, device=dev, dtype=torch.int32) + past).repeat(B)
        x = inputs_embeds.to(torch.float16).reshape(B * S, C).clone()
        if use_cache and cache is None:
            cache = KVCache(cfg.num_hidden_layers, B, cache_max_len or cfg.max_position_embeddings, H, d, dev)
        key_mask = None
        if attention_mask is not None:
            key_mask = attention_mask.to(torch.uint8).contiguous()
            if cache is not None:
                cache.key_mask[:, past:past + S] = key_mask[:, -S:]
                key_mask = cache.key_mask[:, :past + S].contiguous()
        elif cache is not None and past > 0:
            key_mask = cache.key_mask[:, :past + S].contiguous()
        scale = 1.0 / (d ** 0.5)
        hs = [x.view(B, S, C).clone()] if output_hidden_states else None
        h = torch.empty_like(x)
        eps = cfg.rms_norm_eps
        fuse = P["fuse"]
        nL = len(P["layers"])
        # per-row sum of squares of the residual stream after o_proj (slot 2*li) and after down_proj (slot 2*li + 1)
        stats = None
        if fuse:
            stats = ops.cached_row_stats(self, 2 * nL, B * S, C, 1, dev)   # prefill / decode: one buffer per row count
        for li, L in enumerate(P["layers"]):
            if fuse and li > 0:
                qkv = ops.gemm(x, L["w_qkv"], norm=ops.NORM_RMS, rs_in=stats[2 * li - 1], norm_eps=eps)
            else:
                ops.rmsnorm(x, L["ln1"], eps, out=h)
                qkv = ops.gemm(h, L["w_qkv"])
            ops.rope_(qkv, pos, cos_t, sin_t, 2 * H, d)
            qkv5 = qkv.view(B, S, 3, H, d)
            q, k, v = qkv5[:, :, 0], qkv5[:, :, 1], qkv5[:, :, 2]
            if cache is not None:
                cache.k[li][:, past:past + S] = k
                cache.v[li][:, past:past + S] = v
                k, v = cache.k[li][:, :past + S], cache.v[li][:, :past + S]
            a = ops.attention(q, k, v, scale=scale, causal=True, key_mask=key_mask, mask_value=-65504.0)
            if fuse:
                ops.gemm(a.view(B * S, C), L["w_o"], residual=x, out=x, rs_out=stats[2 * li], rs_out_mode=1)
                g = ops.gemm(x, L["w_gu"], epilogue=ops.EPI_SWIGLU, norm=ops.NORM_RMS, rs_in=stats[2 * li], norm_eps=eps)
                last = li + 1 == nL
                ops.gemm(g, L["w_down"], residual=x, out=x, rs_out=None if last else stats[2 * li + 1],
                         rs_out_mode=0 if last else 1)
            else:
                ops.gemm(a.view(B * S, C), L["w_o"], residual=x, out=x)
                ops.rmsnorm(x, L["ln2"], eps, out=h)
                g = ops.gemm(h, L["w_gu"], epilogue=ops.EPI_SWIGLU)
                ops.gemm(g, L["w_down"], residual=x, out=x)
            if output_hidden_states and li + 1 < len(P["layers"]):
                hs.append(x.view(B, S, C).clone())
        out = ops.rmsnorm(x, P["norm"], eps).view(B, S, C)
        if output_hidden_states:
            hs.append(out)  # HF: the last hidden state is post-norm
        if cache is not None:
            cache.length = past + S
        return BaseModelOutputWithPast(out, cache if use_cache or past_key_values is not None else None,
                                       tuple(hs) if hs is not None else None)
