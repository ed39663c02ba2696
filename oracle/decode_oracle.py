"""CPU oracle of the decode step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (only tests/ may import it).

The reference decodes through HuggingFace `generate` (call sites conversation.py:231-243, mvbench_conversation.py:114-125):
every step runs LlamaForCausalLM.forward on one new token per beam with `past_key_values` that hold one full copy of the
prompt's K/V PER BEAM, and `_reorder_cache` index_selects every cache tensor after each beam re-ordering
(modeling_llama_mem.py:205-217 cache append, :715-752 prepare_inputs / reorder).

The B200 decode engine (stllm_b200/decode.py, csrc/decode.cu) computes the same function with another data structure:
ONE copy of the prompt's K/V shared by all beams, the generated tokens' K/V in per-beam suffix buffers that are never moved,
an ancestry table src[beam][j] (physical row that holds token j of the beam's history), keys split into `n_split` slices
whose partial softmax results (max, sum, unnormalised output) are merged afterwards.  This file restates THAT algorithm in
plain fp32 torch so that its equivalence with the reference semantics -- a full recompute of every beam's actual token
history by oracle/stllm_oracle.py:llama_forward, which is pinned against the reference's own modules -- can be checked on
the CPU (tests/test_decode_oracle.py), duplicated ancestors included.
"""
import math

import torch
import torch.nn.functional as F

from . import stllm_oracle as O


def _rope_at(x, cos, sin):
    """x [..., d] at ONE position: cos/sin [d]  (apply_rotary_pos_emb, modeling_llama_mem.py:120-127)"""
    return x * cos + O._rotate_half(x) * sin


class DecodeOracle:
    def __init__(self, sd, lm_head_w, prompt_embeds, n_beams, max_steps, num_layers, num_heads, eps=1e-6, n_split=4,
                 prompt_mask=None):
        """sd: LlamaModel state dict with keys relative to `model.`; prompt_embeds [1, S, C]; prompt_mask [S] (1 = token)."""
        self.sd, self.lm_w = sd, lm_head_w
        self.L, self.H, self.eps, self.NS = num_layers, num_heads, eps, n_split
        self.R, self.Lmax = n_beams, max_steps
        _, S, C = prompt_embeds.shape
        self.S, self.C, self.d = S, C, C // num_heads
        self.cos, self.sin = O.llama_rope_tables(self.d, S + max_steps)
        self.prompt_mask = prompt_mask
        # ---- prefill: the prompt's roped K and V of every layer, ONE copy ----
        neg = torch.finfo(prompt_embeds.dtype).min
        add_mask = torch.full((S, S), neg).triu(1)[None, None]
        if prompt_mask is not None:
            pad = (1.0 - prompt_mask[None, None, None, :].to(prompt_embeds.dtype)).expand(1, 1, S, S)
            add_mask = add_mask + pad.masked_fill(pad.bool(), neg)
        cos, sin = self.cos[None, None, :S], self.sin[None, None, :S]
        x = prompt_embeds
        self.kp, self.vp = [], []
        for i in range(num_layers):
            p = f"layers.{i}."
            h = O.llama_rmsnorm(x, sd[p + "input_layernorm.weight"], eps)
            k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(1, S, num_heads, self.d).transpose(1, 2)
            self.kp.append((k * cos + O._rotate_half(k) * sin)[0].transpose(0, 1).contiguous())     # [S, H, d]
            self.vp.append(F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(S, num_heads, self.d))
            x = O.llama_layer(sd, i, x, add_mask, cos, sin, num_heads, eps)
        self.prefill_logits = F.linear(O.llama_rmsnorm(x, sd["norm.weight"], eps), lm_head_w)[0, -1]
        # ---- suffix caches (never moved) + ancestry table ----
        self.ks = [torch.zeros(n_beams, max_steps, num_heads, self.d) for _ in range(num_layers)]
        self.vs = [torch.zeros(n_beams, max_steps, num_heads, self.d) for _ in range(num_layers)]
        self.src = torch.full((n_beams, max_steps), -1, dtype=torch.long)
        self.n_gen = 0

    def reorder(self, beam_idx):
        """beam b continues the history of beam beam_idx[b]: only the ancestry table is permuted"""
        self.src = self.src[beam_idx.long()].clone()

    def _attention(self, li, q):
        """q [R, H, d] (roped) -> [R, H*d]; keys = shared prompt ++ generated tokens 0..n_gen of each beam through src;
        split over keys like the kernel (slice s of the prompt and slice s of the suffix per split), merged afterwards"""
        R, H, d, NS, S = self.R, self.H, self.d, self.NS, self.S
        n_suf = self.n_gen + 1
        out = torch.zeros(R, H, d)
        scale = 1.0 / math.sqrt(d)
        for r in range(R):
            rows = self.src[r, :n_suf]
            assert (rows >= 0).all(), "ancestry table entry used before it was written"
            k_suf = self.ks[li][rows, torch.arange(n_suf)]            # [n_suf, H, d]
            v_suf = self.vs[li][rows, torch.arange(n_suf)]
            parts = []
            for s in range(NS):
                pk0, pk1 = S * s // NS, S * (s + 1) // NS
                sk0, sk1 = n_suf * s // NS, n_suf * (s + 1) // NS
                K = torch.cat([self.kp[li][pk0:pk1], k_suf[sk0:sk1]], 0)
                V = torch.cat([self.vp[li][pk0:pk1], v_suf[sk0:sk1]], 0)
                sc = torch.einsum("hd,khd->hk", q[r], K) * scale
                if self.prompt_mask is not None and pk1 > pk0:
                    keep = torch.cat([self.prompt_mask[pk0:pk1].bool(), torch.ones(sk1 - sk0, dtype=torch.bool)])
                    sc = sc.masked_fill(~keep[None, :], float("-inf"))
                if K.shape[0] == 0:
                    parts.append((torch.full((H,), float("-inf")), torch.zeros(H), torch.zeros(H, d)))
                    continue
                m = sc.max(dim=1).values
                e = torch.where(torch.isinf(m)[:, None], torch.zeros_like(sc), torch.exp(sc - m[:, None]))
                parts.append((m, e.sum(1), torch.einsum("hk,khd->hd", e, V)))
            M = torch.stack([p[0] for p in parts]).max(dim=0).values                      # [H]
            w = [torch.where(torch.isinf(p[0]), torch.zeros_like(M), torch.exp(p[0] - M)) for p in parts]
            Lsum = sum(p[1] * wi for p, wi in zip(parts, w))
            o = sum(p[2] * wi[:, None] for p, wi in zip(parts, w))
            out[r] = o / Lsum[:, None]
        return out.reshape(R, H * d)

    def step(self, tokens, embed_w):
        """one new token per beam (after an optional reorder): returns next-position logits [R, V]"""
        assert self.n_gen < self.Lmax
        sd, H, d, eps = self.sd, self.H, self.d, self.eps
        pos = self.S + self.n_gen
        cos, sin = self.cos[pos], self.sin[pos]
        x = embed_w[tokens.long()]                                                          # [R, C]
        for i in range(self.L):
            p = f"layers.{i}."
            h = O.llama_rmsnorm(x, sd[p + "input_layernorm.weight"], eps)
            q = _rope_at(F.linear(h, sd[p + "self_attn.q_proj.weight"]).view(self.R, H, d), cos, sin)
            k = _rope_at(F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(self.R, H, d), cos, sin)
            v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(self.R, H, d)
            self.ks[i][:, self.n_gen] = k                       # the new token of beam r lives in physical row r
            self.vs[i][:, self.n_gen] = v
            self.src[:, self.n_gen] = torch.arange(self.R)
            a = self._attention(i, q)
            x = x + F.linear(a, sd[p + "self_attn.o_proj.weight"])
            h = O.llama_rmsnorm(x, sd[p + "post_attention_layernorm.weight"], eps)
            g = F.silu(F.linear(h, sd[p + "mlp.gate_proj.weight"])) * F.linear(h, sd[p + "mlp.up_proj.weight"])
            x = x + F.linear(g, sd[p + "mlp.down_proj.weight"])
        self.n_gen += 1
        return F.linear(O.llama_rmsnorm(x, sd["norm.weight"], eps), self.lm_w)
