"""BLIP-2 Q-Former (BERT-base + cross-attention every 2nd layer) on the sm_100a kernels.

Mirrors the module tree / state-dict keys of the reference's stllm/models/Qformer.py (BertLMHeadModel.bert.
{embeddings, encoder.layer.N.{attention, crossattention, intermediate, output, intermediate_query, output_query}})
so `Qformer.bert.*` checkpoints load unchanged; nn.Modules only store parameters.  forward():
  embeddings  : [LN(query_tokens)] ‖ [LN(word + position)]                              Qformer.py:78-108
  per layer   : fused q|k|v GEMM -> attention (scale 1/8, additive -10000 key mask) -> dense GEMM(+bias,+residual)
                -> post-LN;  on even layers cross-attention of the 32 queries to the 257 image tokens, whose K|V
                projections for ALL cross layers come from ONE GEMM over ln_vision(ViT) ([N*257,1408] x
                [n_cross*2*768,1408]^T) so the image tokens are read from HBM once;  query FFN (GELU-erf epilogue)
                + post-LN; text FFN for the text tokens when qformer_text_input.              Qformer.py:402-484
"""
import torch
import torch.nn as nn

from . import ops
from .packing import PackedWeightsMixin


class BertConfig:
    """the fields of transformers.BertConfig the Q-Former uses (defaults == bert-base-uncased, blip2.py:48-53)"""

    def __init__(self, vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, max_position_embeddings=512, layer_norm_eps=1e-12, initializer_range=0.02,
                 pad_token_id=0, encoder_width=1408, add_cross_attention=True, cross_attention_freq=2,
                 query_length=32, **unused):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.max_position_embeddings = max_position_embeddings
        self.layer_norm_eps = layer_norm_eps
        self.initializer_range = initializer_range
        self.pad_token_id = pad_token_id
        self.encoder_width = encoder_width
        self.add_cross_attention = add_cross_attention
        self.cross_attention_freq = cross_attention_freq
        self.query_length = query_length
        self.hidden_act = "gelu"

    def get(self, k, default=None):
        return getattr(self, k, default)


class BertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))


class BertSelfAttention(nn.Module):
    def __init__(self, config, is_cross_attention):
        super().__init__()
        kv_in = config.encoder_width if is_cross_attention else config.hidden_size
        self.query = nn.Linear(config.hidden_size, config.hidden_size)
        self.key = nn.Linear(kv_in, config.hidden_size)
        self.value = nn.Linear(kv_in, config.hidden_size)


class BertSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertAttention(nn.Module):
    def __init__(self, config, is_cross_attention=False):
        super().__init__()
        self.self = BertSelfAttention(config, is_cross_attention)
        self.output = BertSelfOutput(config)


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class BertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertLayer(nn.Module):
    def __init__(self, config, layer_num):
        super().__init__()
        self.layer_num = layer_num
        self.attention = BertAttention(config)
        self.has_cross_attention = bool(config.add_cross_attention and layer_num % config.cross_attention_freq == 0)
        if self.has_cross_attention:
            self.crossattention = BertAttention(config, is_cross_attention=True)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)
        self.intermediate_query = BertIntermediate(config)
        self.output_query = BertOutput(config)


class BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(config, i) for i in range(config.num_hidden_layers)])


class BertModel(PackedWeightsMixin, nn.Module):
    """Qformer.py:677-965 (add_pooling_layer=False as BertLMHeadModel builds it, :977)"""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config)
        self.apply(self._init_weights)
        self._packed = None

    def _init_weights(self, module):
        """Qformer.py:664-674"""
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def _apply(self, fn, *a, **kw):
        self._packed = None
        return super()._apply(fn, *a, **kw)

    def _pack(self):
        dev = self.embeddings.LayerNorm.weight.device
        f16 = lambda t: t.detach().to(device=dev, dtype=torch.float16).contiguous()
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()

        def ln(m):
            return f32(m.weight), f32(m.bias), m.eps

        emb = self.embeddings
        P = dict(emb_ln=ln(emb.LayerNorm))
        if emb.word_embeddings is not None:
            P["word"] = f16(emb.word_embeddings.weight)
            P["pos"] = f16(emb.position_embeddings.weight)
        layers = []
        cross_k, cross_v, cross_kb, cross_vb = [], [], [], []
        for lyr in self.encoder.layer:
            sa = lyr.attention
            d = dict(
                w_qkv=torch.cat([f16(sa.self.query.weight), f16(sa.self.key.weight), f16(sa.self.value.weight)], 0),
                b_qkv=torch.cat([f16(sa.self.query.bias), f16(sa.self.key.bias), f16(sa.self.value.bias)], 0),
                w_o=f16(sa.output.dense.weight), b_o=f16(sa.output.dense.bias), ln_o=ln(sa.output.LayerNorm),
                w_iq=f16(lyr.intermediate_query.dense.weight), b_iq=f16(lyr.intermediate_query.dense.bias),
                w_oq=f16(lyr.output_query.dense.weight), b_oq=f16(lyr.output_query.dense.bias),
                ln_oq=ln(lyr.output_query.LayerNorm), cross=None)
            if lyr.intermediate is not None:
                d.update(w_it=f16(lyr.intermediate.dense.weight), b_it=f16(lyr.intermediate.dense.bias),
                         w_ot=f16(lyr.output.dense.weight), b_ot=f16(lyr.output.dense.bias),
                         ln_ot=ln(lyr.output.LayerNorm))
            if lyr.has_cross_attention:
                ca = lyr.crossattention
                d["cross"] = dict(idx=len(cross_k), w_q=f16(ca.self.query.weight), b_q=f16(ca.self.query.bias),
                                  w_o=f16(ca.output.dense.weight), b_o=f16(ca.output.dense.bias),
                                  ln_o=ln(ca.output.LayerNorm))
                cross_k.append(f16(ca.self.key.weight)); cross_kb.append(f16(ca.self.key.bias))
                cross_v.append(f16(ca.self.value.weight)); cross_vb.append(f16(ca.self.value.bias))
            layers.append(d)
        P["layers"] = layers
        P["n_cross"] = len(cross_k)
        if cross_k:
            # rows: [layer c: K (hidden) | V (hidden)] for every cross layer -> one GEMM over the image tokens
            P["w_kv_all"] = torch.cat([torch.cat([k, v], 0) for k, v in zip(cross_k, cross_v)], 0).contiguous()
            P["b_kv_all"] = torch.cat([torch.cat([k, v], 0) for k, v in zip(cross_kb, cross_vb)], 0).contiguous()
        self._packed = P
        self._mark_packed()

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, query_embeds=None, encoder_hidden_states=None,
                encoder_attention_mask=None, return_dict=True, return_layer_outputs=False):
        """Same call shape as st_llm.py:353-367.  query_embeds [N,Q,C] (expanded query tokens), encoder_hidden_states
        [N,P,W]; optional input_ids [N,Lt] with attention_mask [N,Q+Lt].  encoder_attention_mask must be all ones
        (it is in every reference call site: st_llm.py:333)."""
        assert query_embeds is not None, "You have to specify query_embeds when input_ids is None"
        if not self._packed_valid():
            self._pack()
        P = self._packed
        cfg = self.config
        C, Hh = cfg.hidden_size, cfg.num_attention_heads
        d = C // Hh
        N, Q, _ = query_embeds.shape
        scale = 1.0 / (d ** 0.5)
        g, b, eps = P["emb_ln"]
        # ---- embeddings ----
        if input_ids is not None:
            Lt = input_ids.shape[1]
            S = Q + Lt
            txt = ops.gather_rows(P["word"], input_ids.reshape(-1), addend=P["pos"][:Lt].contiguous())
            pre = torch.empty(N, S, C, dtype=torch.float16, device=txt.device)
            pre[:, :Q] = query_embeds.to(torch.float16)
            pre[:, Q:] = txt.view(N, Lt, C)
            hidden = ops.layernorm(pre.view(N * S, C), g, b, eps)
            key_mask = attention_mask.to(torch.uint8).contiguous() if attention_mask is not None else None
        else:
            S = Q
            hidden = ops.layernorm(query_embeds.to(torch.float16).contiguous().view(N * Q, C), g, b, eps)
            key_mask = None
        # ---- K|V of the image tokens for all cross-attention layers in one GEMM ----
        enc = encoder_hidden_states
        Pn, Wd = enc.shape[1], enc.shape[2]
        kv_all = None
        if P["n_cross"]:
            kv_all = ops.gemm(enc.reshape(N * Pn, Wd), P["w_kv_all"], bias=P["b_kv_all"])
            kv_all = kv_all.view(N, Pn, P["n_cross"], 2, Hh, d)
        layer_outs = []
        for L in P["layers"]:
            qkv = ops.gemm(hidden, L["w_qkv"], bias=L["b_qkv"]).view(N, S, 3, Hh, d)
            ctx = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale=scale, key_mask=key_mask,
                                mask_value=-10000.0)
            att = ops.gemm(ctx.view(N * S, C), L["w_o"], bias=L["b_o"], residual=hidden)
            att = ops.layernorm(att, *L["ln_o"])
            if S > Q:
                att3 = att.view(N, S, C)
                q_out = att3[:, :Q].contiguous().view(N * Q, C)
                t_out = att3[:, Q:].contiguous().view(N * (S - Q), C)
            else:
                q_out, t_out = att, None
            if L["cross"] is not None:
                X = L["cross"]
                qc = ops.gemm(q_out, X["w_q"], bias=X["b_q"]).view(N, Q, Hh, d)
                kc = kv_all[:, :, X["idx"], 0]
                vc = kv_all[:, :, X["idx"], 1]
                cctx = ops.attention(qc, kc, vc, scale=scale)
                q_new = ops.gemm(cctx.view(N * Q, C), X["w_o"], bias=X["b_o"], residual=q_out)
                q_out = ops.layernorm(q_new, *X["ln_o"])
            inter = ops.gemm(q_out, L["w_iq"], bias=L["b_iq"], epilogue=ops.EPI_GELU)
            oq = ops.gemm(inter, L["w_oq"], bias=L["b_oq"], residual=q_out)
            oq = ops.layernorm(oq, *L["ln_oq"])
            if t_out is not None:
                it = ops.gemm(t_out, L["w_it"], bias=L["b_it"], epilogue=ops.EPI_GELU)
                ot = ops.gemm(it, L["w_ot"], bias=L["b_ot"], residual=t_out)
                ot = ops.layernorm(ot, *L["ln_ot"])
                hidden3 = torch.empty(N, S, C, dtype=torch.float16, device=oq.device)
                hidden3[:, :Q] = oq.view(N, Q, C)
                hidden3[:, Q:] = ot.view(N, S - Q, C)
                hidden = hidden3.view(N * S, C)
            else:
                hidden = oq
            if return_layer_outputs:
                layer_outs.append(hidden.view(N, S, C).clone())
        out = hidden.view(N, S, C)
        res = _Output(out)
        if return_layer_outputs:
            res.layer_outputs = layer_outs
        return res if return_dict else (out,)


class _Output:
    def __init__(self, last_hidden_state):
        self.last_hidden_state = last_hidden_state


class BertLMHeadModel(nn.Module):
    """Qformer.py:968- : only `.bert` is on the ST-LLM path (`.cls` is dropped at st_llm.py:288)"""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.bert = BertModel(config)
        self.cls = None

    def resize_token_embeddings(self, new_num_tokens):
        """st_llm.py:285 -> HF resize: keep old rows, init new rows ~ N(0, initializer_range)"""
        old = self.bert.embeddings.word_embeddings
        if old is None or new_num_tokens == old.num_embeddings:
            return old
        new = nn.Embedding(new_num_tokens, old.embedding_dim, padding_idx=old.padding_idx).to(
            device=old.weight.device, dtype=old.weight.dtype)
        new.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        n = min(old.num_embeddings, new_num_tokens)
        new.weight.data[:n] = old.weight.data[:n]
        self.bert.embeddings.word_embeddings = new
        self.config.vocab_size = new_num_tokens
        self.bert.repack()
        return new
