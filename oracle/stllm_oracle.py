"""CPU oracle for the ST-LLM video-token hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch, fp32, CPU restatement of the reference's algorithm for the path SURVEY.md section 8 names.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package (stllm_b200/) never does.

Every function cites the reference file:line it follows.  All functions are purely functional over a
*state dict with the reference's parameter names* (e.g. "blocks.0.attn.qkv.weight"), so the same weights
can be loaded into (a) the real reference modules (tests/golden/make_golden.py, run where /root/reference
exists), (b) this oracle and (c) the CUDA path.

Pinning: the reference has no tests or golden vectors for this path ("parity unpinned" by the
reference itself).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: the fixtures in
tests/golden/*.npz were produced by importing the reference's own modules (tests/golden/make_golden.py, via
oracle/ref_shim.py) on seeded inputs; tests/test_oracle_golden.py checks this file against them.

Third-party arithmetic not in /root/reference: HuggingFace transformers==4.28.0 (requirement.txt:42)
`transformers.models.llama.modeling_llama`.  Its algorithm is restated in llama_* below following the
reference's in-tree twin stllm/models/modeling_llama_mem.py (same math, attention call replaced by the
eager masked-softmax attention of HF 4.28), and is pinned through the reference's own call sites
(st_llm.py:56-146) executed with the installed transformers in make_golden.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# EVA-ViT-g  (stllm/models/eva_vit.py)
# --------------------------------------------------------------------------------------------------


def vit_patch_embed(sd, x, patch):
    """PatchEmbed.forward, eva_vit.py:198-204: Conv2d(k=stride=patch) -> flatten(2).transpose(1,2)"""
    y = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch)
    return y.flatten(2).transpose(1, 2)


def vit_attention(sd, pre, x, num_heads):
    """Attention.forward, eva_vit.py:118-148 (no rel-pos bias: use_rel_pos_bias=False in create_eva_vit_g)"""
    B, N, C = x.shape
    qkv_bias = None
    if (pre + "q_bias") in sd:
        qb, vb = sd[pre + "q_bias"], sd[pre + "v_bias"]
        qkv_bias = torch.cat((qb, torch.zeros_like(vb), vb))                      # :121-122
    qkv = F.linear(x, sd[pre + "qkv.weight"], qkv_bias)                           # :124
    qkv = qkv.reshape(B, N, 3, num_heads, -1).permute(2, 0, 3, 1, 4)              # :125
    q, k, v = qkv[0], qkv[1], qkv[2]
    head_dim = q.shape[-1]
    q = q * (head_dim ** -0.5)                                                    # :128
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)                              # :129,142
    x = (attn @ v).transpose(1, 2).reshape(B, N, -1)                              # :145
    return F.linear(x, sd[pre + "proj.weight"], sd[pre + "proj.bias"])            # :146


def vit_block(sd, i, x, num_heads, eps=1e-6):
    """Block.forward, eva_vit.py:173-180 with gamma_1 is None (init_values=None, :415-428); LN eps 1e-6 (:426)"""
    p = f"blocks.{i}."
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    x = x + vit_attention(sd, p + "attn.", h, num_heads)
    h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])             # Mlp.forward :54-61
    h = F.gelu(h)                                                                 # nn.GELU (erf)
    h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + h


def vit_forward(sd, frames, depth, num_heads, patch=14, collect=None):
    """VisionTransformer.forward_features, eva_vit.py:324-340: no final norm, all tokens returned"""
    x = vit_patch_embed(sd, frames, patch)
    cls = sd["cls_token"].expand(x.shape[0], -1, -1)                              # :328
    x = torch.cat((cls, x), dim=1)                                                # :329
    x = x + sd["pos_embed"]                                                       # :331
    if collect is not None:
        collect["vit_embed"] = x
    for i in range(depth):
        x = vit_block(sd, i, x, num_heads)
        if collect is not None:
            collect[f"vit_block{i}"] = x
    return x


# --------------------------------------------------------------------------------------------------
# BT-Adapter encoder  (stllm/models/eva_btadapter.py) -- the vit_model="eva_btadapter_g" encoder of the shipped
# QA / MVBench configs: EVA-ViT-g plus a temporal+spatial branch on the last `bt_depth` blocks
# --------------------------------------------------------------------------------------------------


def _bt_temporal(sd, j, x, T, num_heads, eps=1e-6):
    """BTAdapter_Temp.forward, eva_btadapter.py:291-308: attention over the T frames of each patch position,
    temporal_fc, residual; the cls row passes through"""
    p = f"BTAdapter_T.{j}."
    residual = x[:, 1:, :]
    init_cls = x[:, :1, :]
    b, pt, m = residual.shape
    pn = pt // T
    q = residual.reshape(b * pn, T, m)                                            # :300 rows are (p t)-ordered
    h = F.layer_norm(q, (m,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    h = vit_attention(sd, p + "attn.", h, num_heads)                              # :302 (drop_path: eval identity)
    h = F.linear(h, sd[p + "temporal_fc.weight"], sd[p + "temporal_fc.bias"])     # :304
    return torch.cat((init_cls, h.reshape(b, pn * T, m) + residual), 1)           # :306-307


def _bt_spatial(sd, j, x, T, num_heads, eps=1e-5):
    """BTAdapter_Spatial.forward, eva_btadapter.py:258-279: a ViT Block applied per frame on [cls ; patches of that
    frame]; the cls output is averaged over frames.  LN eps is 1e-5 here: the class is built with Block's DEFAULT
    norm_layer=nn.LayerNorm (eva_btadapter.py:255, eva_vit.py:152), unlike the ViT's own blocks (eps 1e-6)"""
    p = f"BTAdapter_S.{j}."
    residual = x
    init_cls = x[:, :1, :]
    qs = x[:, 1:, :]
    b, pt, m = qs.shape
    pn = pt // T
    cls = init_cls.unsqueeze(1).repeat(1, T, 1, 1).reshape(b * T, 1, m)           # :267
    qs = qs.reshape(b, pn, T, m).permute(0, 2, 1, 3).reshape(b * T, pn, m)        # 'b (p t) m -> (b t) p m' :268
    h = torch.cat((cls, qs), 1)
    h = F.layer_norm(h, (m,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    h = vit_attention(sd, p + "attn.", h, num_heads)                              # :271
    cls = h[:, :1, :].reshape(b, T, 1, m).mean(1)                                 # :273
    rs = h[:, 1:, :].reshape(b, T, pn, m).permute(0, 2, 1, 3).reshape(b, pn * T, m)   # '(b t) p m -> b (p t) m' :274
    x = residual + torch.cat((cls, rs), 1)                                        # :276
    h = F.layer_norm(x, (m,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    return x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])      # :277


def _bt_init_input(sd, x4):
    """EVAVisionTransformer_BTAdapter.init_input, eva_btadapter.py:203-226 (mask=None: mask_rate 0 / eval)"""
    b, t, l1, d = x4.shape
    cls_x = x4[:, :, 0].mean(dim=1).unsqueeze(1)                                  # :204
    x = x4[:, :, 1:, :].reshape(b * t, l1 - 1, d)                                 # :205-207
    cls_branch = sd["BTAdapter_cls"].expand(x.shape[0], 1, -1)                    # :209
    x = torch.cat((cls_branch, x), dim=1) + sd["pos_embed"]                       # :210-211
    cls_branch = x[:b, 0, :].unsqueeze(1)                                         # :214
    x = x[:, 1:, :].reshape(b, t, l1 - 1, d).permute(0, 2, 1, 3)                  # '(b t) l d -> (b l) t d' :215
    x = x + sd["BTAdapter_position.weight"][:t][None, None]                       # :216-218
    x = x.reshape(b, (l1 - 1) * t, d)                                             # '(b l) t d -> b (l t) d' :220
    cls = (cls_x + cls_branch) / 2                                                # :223
    return torch.cat((cls, x), dim=1)


def bt_adapter_forward(sd, frames, depth, num_heads, bt_depth=3, patch=14, collect=None):
    """EVAVisionTransformer_BTAdapter.forward / forward_features, eva_btadapter.py:147-181,228-250.
    frames [B,T,3,H,W] (or [T,3,H,W] = one clip, :236-238) -> [(B T), 1+P, D]"""
    if frames.dim() == 4:
        frames = frames[None]
    elif frames.shape[1] == 3:
        frames = frames.permute(0, 2, 1, 3, 4)                                    # [B,C,T,H,W] input, :231-233 (also
    B, T = frames.shape[:2]                                                       # taken by a [B,T=3,C,H,W] clip!)
    x = vit_patch_embed(sd, frames.reshape((-1,) + frames.shape[2:]), patch)
    x = torch.cat((sd["cls_token"].expand(x.shape[0], -1, -1), x), dim=1) + sd["pos_embed"]   # :152-156
    branch = None
    for i in range(depth):
        x = vit_block(sd, i, x, num_heads)
        if i >= depth - bt_depth:                                                 # :166-168
            j = i + bt_depth - depth
            x4 = x.reshape(B, T, x.shape[1], x.shape[2])                          # forward_branch :184
            if branch is not None:
                cls_branch = x4[:, :, 0].mean(dim=1).unsqueeze(1)                 # :186-187
                xp = x4[:, :, 1:].permute(0, 2, 1, 3).reshape(B, -1, x.shape[2])  # 'b t l d -> b (l t) d' :188
                z = torch.cat((cls_branch, xp), dim=1) + branch                   # :192-193
            else:
                z = x4
            if j == 0:
                z = _bt_init_input(sd, z)                                         # :195-196
            z = _bt_temporal(sd, j, z, T, num_heads)                              # :202
            branch = _bt_spatial(sd, j, z, T, num_heads)
            if collect is not None:
                collect[f"bt_branch{j}"] = branch
    pn = x.shape[1] - 1                                                           # :178-181
    b_cls, b_patch = branch[:, 0], branch[:, 1:]
    b_patch = b_patch.reshape(B, pn, T, -1).permute(0, 2, 1, 3).reshape(B * T, pn, -1)   # 'b (p t) m -> (b t) p m'
    b_cls = b_cls.repeat(1, T).view(B * T, -1).unsqueeze(1)
    return (x + torch.cat((b_cls, b_patch), dim=1)) / 2


def ln_vision(sd, x, eps=1e-5):
    """blip2.LayerNorm (fp32 LayerNorm, nn default eps), blip2.py:73,103-109; call site st_llm.py:331"""
    ret = F.layer_norm(x.float(), (x.shape[-1],), sd["weight"].float(), sd["bias"].float(), eps)
    return ret.to(x.dtype)                       # blip2.py:105-109: computed in fp32, cast back to the input dtype


# --------------------------------------------------------------------------------------------------
# Q-Former  (stllm/models/Qformer.py)
# --------------------------------------------------------------------------------------------------


def _bert_attention(sd, pre, hidden, kv_source, add_mask, num_heads, ln_eps):
    """BertAttention = BertSelfAttention (Qformer.py:169-275) + BertSelfOutput (:278-289)"""
    def heads(t):
        b, s, _ = t.shape
        return t.view(b, s, num_heads, -1).permute(0, 2, 1, 3)                    # transpose_for_scores :160-166
    k = heads(F.linear(kv_source, sd[pre + "self.key.weight"], sd[pre + "self.key.bias"]))
    v = heads(F.linear(kv_source, sd[pre + "self.value.weight"], sd[pre + "self.value.bias"]))
    q = heads(F.linear(hidden, sd[pre + "self.query.weight"], sd[pre + "self.query.bias"]))
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.shape[-1])        # :208,244
    if add_mask is not None:
        scores = scores + add_mask                                                # :247
    probs = torch.softmax(scores, dim=-1)                                         # :250 (dropout: eval)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
    ctx = ctx.view(ctx.shape[0], ctx.shape[1], -1)                                # :263-265
    out = F.linear(ctx, sd[pre + "output.dense.weight"], sd[pre + "output.dense.bias"])
    return F.layer_norm(out + hidden, (hidden.shape[-1],), sd[pre + "output.LayerNorm.weight"],
                        sd[pre + "output.LayerNorm.bias"], ln_eps)                # :286-288


def _bert_ffn(sd, pre_i, pre_o, x, ln_eps):
    """BertIntermediate (Qformer.py:349-361) + BertOutput (:364-375)"""
    h = F.gelu(F.linear(x, sd[pre_i + "dense.weight"], sd[pre_i + "dense.bias"]))
    h = F.linear(h, sd[pre_o + "dense.weight"], sd[pre_o + "dense.bias"])
    return F.layer_norm(h + x, (x.shape[-1],), sd[pre_o + "LayerNorm.weight"], sd[pre_o + "LayerNorm.bias"], ln_eps)


def qformer_forward(sd, query_tokens, image_embeds, num_layers, num_heads=12, cross_freq=2, ln_eps=1e-12,
                    text_ids=None, text_mask=None, collect=None):
    """BertModel.forward (Qformer.py:804-965) as called from st_llm.py:353-367.

    sd keys are relative to `Qformer.bert.` (e.g. "encoder.layer.0.attention.self.query.weight").
    query_tokens [1,Q,768]; image_embeds [N,P,1408]; optional text_ids [N,Lt] int64 + text_mask [N,Lt].
    Returns last_hidden_state [N, Q(+Lt), 768]."""
    N = image_embeds.shape[0]
    Q = query_tokens.shape[1]
    q_emb = query_tokens.expand(N, -1, -1)
    # BertEmbeddings.forward, Qformer.py:78-108
    if text_ids is not None:
        Lt = text_ids.shape[1]
        emb = sd["embeddings.word_embeddings.weight"][text_ids]
        emb = emb + sd["embeddings.position_embeddings.weight"][torch.arange(Lt)][None]
        emb = torch.cat((q_emb, emb), dim=1)
    else:
        emb = q_emb
    C = emb.shape[-1]
    hidden = F.layer_norm(emb, (C,), sd["embeddings.LayerNorm.weight"], sd["embeddings.LayerNorm.bias"], ln_eps)
    # get_extended_attention_mask (non-decoder), Qformer.py:784-802: (1 - mask) * -10000
    if text_ids is not None:
        att = torch.cat([torch.ones(N, Q), text_mask.float()], dim=1)             # st_llm.py:352-353
        self_mask = ((1.0 - att) * -10000.0)[:, None, None, :]
    else:
        self_mask = None                                                          # all ones -> zeros
    # encoder_attention_mask = ones -> invert_attention_mask -> zeros: omitted
    for i in range(num_layers):
        p = f"encoder.layer.{i}."
        # BertLayer.forward, Qformer.py:402-474
        att_out = _bert_attention(sd, p + "attention.", hidden, hidden, self_mask, num_heads, ln_eps)
        q_out = att_out[:, :Q]
        if i % cross_freq == 0:                                                   # :386-395
            q_out = _bert_attention(sd, p + "crossattention.", q_out, image_embeds, None, num_heads, ln_eps)
        out = _bert_ffn(sd, p + "intermediate_query.", p + "output_query.", q_out, ln_eps)   # :449-455
        if att_out.shape[1] > Q:                                                  # :456-463
            out_t = _bert_ffn(sd, p + "intermediate.", p + "output.", att_out[:, Q:], ln_eps)
            out = torch.cat([out, out_t], dim=1)
        hidden = out
        if collect is not None:
            collect[f"qformer_layer{i}"] = hidden
    return hidden


# --------------------------------------------------------------------------------------------------
# Llama  (transformers 4.28 modeling_llama; in-tree twin stllm/models/modeling_llama_mem.py)
# --------------------------------------------------------------------------------------------------


def llama_rmsnorm(x, w, eps):
    """LlamaRMSNorm.forward, modeling_llama_mem.py:70-78"""
    var = x.float().pow(2).mean(-1, keepdim=True)
    h = x * torch.rsqrt(var + eps)
    if w.dtype in (torch.float16, torch.bfloat16):       # :75-76 "convert into half-precision if necessary"
        h = h.to(w.dtype)
    return w * h


def llama_rope_tables(head_dim, seq_len, base=10000.0):
    """LlamaRotaryEmbedding, modeling_llama_mem.py:81-110"""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(seq_len, dtype=torch.float32)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def _rotate_half(x):
    """modeling_llama_mem.py:113-117"""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def llama_layer(sd, i, x, add_mask, cos, sin, num_heads, eps):
    """LlamaDecoderLayer.forward, modeling_llama_mem.py:264-316 with the HF-4.28 eager attention
    (q k^T / sqrt(d) + mask, clamp at dtype min, fp32 softmax) instead of the flash-attn call (:219-244)."""
    p = f"layers.{i}."
    B, S, C = x.shape
    d = C // num_heads
    h = llama_rmsnorm(x, sd[p + "input_layernorm.weight"], eps)
    q = F.linear(h, sd[p + "self_attn.q_proj.weight"]).view(B, S, num_heads, d).transpose(1, 2)
    k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(B, S, num_heads, d).transpose(1, 2)
    v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(B, S, num_heads, d).transpose(1, 2)
    q = q * cos + _rotate_half(q) * sin                                           # apply_rotary_pos_emb :120-127
    k = k * cos + _rotate_half(k) * sin
    w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(d)
    w = w + add_mask
    w = torch.max(w, torch.tensor(torch.finfo(w.dtype).min))
    w = torch.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    a = torch.matmul(w, v).transpose(1, 2).reshape(B, S, C)
    x = x + F.linear(a, sd[p + "self_attn.o_proj.weight"])
    h = llama_rmsnorm(x, sd[p + "post_attention_layernorm.weight"], eps)
    g = F.silu(F.linear(h, sd[p + "mlp.gate_proj.weight"])) * F.linear(h, sd[p + "mlp.up_proj.weight"])  # :143
    return x + F.linear(g, sd[p + "mlp.down_proj.weight"])


def llama_forward(sd, inputs_embeds, attention_mask, num_layers, num_heads, eps=1e-6, all_hidden=False, collect=None):
    """LlamaModel.forward, modeling_llama_mem.py:464-594.  sd keys relative to `model.` ("layers.0...", "norm.weight").
    attention_mask [B,S] (1 = token, 0 = pad) or None.  Returns final-norm hidden [B,S,C] (and the list of all
    hidden states like output_hidden_states=True: embeddings, each layer, with the LAST entry being post-norm)."""
    B, S, C = inputs_embeds.shape
    d = C // num_heads
    dev, dt = inputs_embeds.device, inputs_embeds.dtype     # fp32 CPU in the tests; fp16 CUDA for the "reference
    cos, sin = llama_rope_tables(d, S)                       # arithmetic in fp16 on the GPU" error-budget test
    cos, sin = cos[None, None].to(dev, dt), sin[None, None].to(dev, dt)   # HF casts the cached tables to x.dtype
    neg = torch.finfo(dt).min
    # _prepare_decoder_attention_mask (HF 4.28): causal (dtype-min above diagonal) + expanded padding mask
    causal = torch.full((S, S), neg, device=dev, dtype=dt).triu(1)[None, None]
    add_mask = causal
    if attention_mask is not None:
        pad = (1.0 - attention_mask[:, None, None, :].to(dev, dt)).expand(B, 1, S, S)
        add_mask = causal + pad.masked_fill(pad.bool(), neg)
    x = inputs_embeds
    hs = [x]
    for i in range(num_layers):
        x = llama_layer(sd, i, x, add_mask, cos, sin, num_heads, eps)
        hs.append(x)
        if collect is not None:
            collect[f"llama_layer{i}"] = x
    x = llama_rmsnorm(x, sd["norm.weight"], eps)
    hs[-1] = x  # HF adds the post-norm state as the last hidden state
    return (x, hs) if all_hidden else x


# --------------------------------------------------------------------------------------------------
# ST-LLM glue  (stllm/models/st_llm.py, stllm/models/utils.py)
# --------------------------------------------------------------------------------------------------


def get_residual_index(sample_segments, total_segments):
    """STLLMModel.get_residual_index, st_llm.py:434-445"""
    seg_size = float(total_segments) / sample_segments
    return np.array([int((seg_size / 2) + np.round(seg_size * idx)) for idx in range(sample_segments)])


def random_masking_generator(num_patches, mask_ratio, batch):
    """RandomMaskingGenerator, utils.py:4-16 -- consumes numpy's GLOBAL RNG exactly like the reference
    (one np.random.shuffle per batch row). Returns bool [batch, num_patches], True = masked."""
    num_mask = int(mask_ratio * num_patches)
    rows = []
    for _ in range(batch):
        mask = np.hstack([np.zeros(num_patches - num_mask), np.ones(num_mask)])
        np.random.shuffle(mask)
        rows.append(mask)
    return torch.tensor(np.array(rows)).to(torch.bool)


def draw_mask_rate():
    """st_llm.py:484-485: rate ~ N(0.5, 0.1) clipped to [0.1, 0.7] (global numpy RNG)"""
    return float(np.clip(np.random.normal(0.5, 0.1), 0.1, 0.7))


def encode_img(w, image, cfg, text_ids=None, text_mask=None, collect=None):
    """STLLMModel.encode_img, st_llm.py:321-377.
    w: dict of sub-state-dicts {"vit","ln_vision","qformer","query_tokens","llama_proj"};
    image [B,T,3,H,W] or [T,3,H,W].  Returns inputs_llama [B,T,Q,4096] or [T,Q,4096]
    (cfg["has_qformer"] = False, :369-373: [B,T,P/4,4096] from the patch tokens themselves)."""
    infer = image.dim() == 4
    T = image.shape[1]
    if cfg.get("vit_model", "eva_clip_g") == "eva_btadapter_g":
        # :327 -- only eva_clip_g gets the flattened frames; the BT-Adapter encoder takes the clip tensor as is
        x = bt_adapter_forward(w["vit"], image, cfg["vit_depth"], cfg["vit_heads"], cfg.get("bt_depth", 3),
                               cfg.get("patch", 14), collect)
    else:
        if not infer:
            image = image.reshape((-1,) + tuple(image.shape[2:]))                # 'B T C H W -> (B T) C H W' :328
        x = vit_forward(w["vit"], image, cfg["vit_depth"], cfg["vit_heads"], cfg.get("patch", 14), collect)
    x = ln_vision(w["ln_vision"], x)                                             # :331
    if collect is not None:
        collect["ln_vision"] = x
    if not cfg.get("has_qformer", True):
        x = x[:, 1:, :]                                                           # :370 drop the cls token
        bs, pn, hs = x.shape
        x = x.reshape(bs, int(pn / 4), int(hs * 4))                               # :371-372 4 neighbouring tokens -> 1
        out = F.linear(x, w["llama_proj"]["weight"], w["llama_proj"]["bias"])     # :373 (llama_proj: 4 C -> D, :301)
        if not infer:
            out = out.view(-1, T, out.shape[1], out.shape[-1])                    # :374-375
        return out
    if text_ids is not None and text_ids.shape[0] != x.shape[0]:
        text_ids = text_ids.repeat_interleave(T, dim=0)                          # :340-343 (text repeated per frame)
        text_mask = text_mask.repeat_interleave(T, dim=0)
    h = qformer_forward(w["qformer"], w["query_tokens"], x, cfg["qf_layers"], cfg.get("qf_heads", 12),
                        text_ids=text_ids, text_mask=text_mask, collect=collect)
    Q = w["query_tokens"].shape[1]
    out = F.linear(h[:, :Q], w["llama_proj"]["weight"], w["llama_proj"]["bias"])  # :368
    if not infer:
        out = out.view(-1, T, Q, out.shape[-1])                                   # :374-375
    return out


def video_aggregate(img_embeds, video_input, residual_size=4, res_w=None):
    """st_llm.py:463-478: 'all' | 'mean' | 'residual'.  img_embeds [B,T,Q,D] -> [B,1,L,D]."""
    B, T = img_embeds.shape[:2]
    D = img_embeds.shape[-1]
    if video_input == "all":
        return img_embeds.reshape(B, 1, -1, D)
    if video_input == "mean":
        return img_embeds.mean(dim=1, keepdim=True)
    if video_input == "residual":
        idx = torch.from_numpy(get_residual_index(residual_size, T))
        g = img_embeds.mean(dim=1, keepdim=True)
        local = img_embeds[:, idx]
        g = g.expand(-1, residual_size, -1, -1)
        g = F.linear(F.relu(F.linear(g, res_w["down_proj.weight"], res_w["down_proj.bias"])),
                     res_w["up_proj.weight"], res_w["up_proj.bias"])
        return (local + g).reshape(B, 1, -1, D)
    raise ValueError(video_input)


def mvm_loss(sd_dec, mask_hidden, unmask_hidden, mask, img_start, img_len, mask_img_len):
    """STLLMLlamaModel.forward MVM branch, st_llm.py:71-91 + Linear_Decoder :35-43.
    mask_hidden / unmask_hidden: final hidden states [B,S,D] of the masked / unmasked passes; mask bool [B,1,L]."""
    B, _, D = mask_hidden.shape
    m = mask_hidden[:, img_start:img_start + mask_img_len]
    m = F.layer_norm(F.linear(m, sd_dec["head.weight"], sd_dec["head.bias"]), (D,), sd_dec["norm.weight"],
                     sd_dec["norm.bias"], 1e-5)
    u = unmask_hidden[:, img_start:img_start + img_len]
    u = u[~(mask.squeeze(1))].reshape(B, -1, D)
    m = m / m.norm(dim=-1, keepdim=True)
    u = u / u.norm(dim=-1, keepdim=True)
    return (2 - 2 * (m * u).sum(dim=-1)).mean()


def causal_lm_loss(logits, labels, vocab):
    """STLLMForCausalLM.forward, st_llm.py:124-135: shift, flatten, CrossEntropyLoss (ignore -100)"""
    sl = logits[..., :-1, :].contiguous().view(-1, vocab)
    tl = labels[..., 1:].contiguous().view(-1)
    return F.cross_entropy(sl, tl)


def assemble_sequence(video_embeds, prompt_embeds, answer_embeds, answer_mask, pad_embed, answer_ids, pad_id,
                      before_embeds=None, bos_embed=None):
    """prompt_wrap (st_llm.py:379-407) + concat_emb_input_output (:409-432) + BOS prepend (:519-530) + targets
    (:532-542).  Two layouts:
      * qformer_text_input / 'simple' (before_embeds=None, bos_embed=None): '<ImageHere>' comes first =>
        [video ; prompt ; answer ; pad], label offset 0;
      * MiniGPT-4 (before_embeds = list of [P0_i,D] embeddings of the text before '<ImageHere>', bos_embed [D]):
        [BOS ; before ; video ; prompt ; answer ; pad], label offset 1 (st_llm.py:537-541).
    video_embeds [B,1,L,D]; prompt_embeds: list of [P_i,D]; answer_embeds [B,A,D]; answer_mask [B,A].
    Returns inputs_embeds [B,S,D], attention_mask [B,S] int, targets [B,S] int64."""
    B = video_embeds.shape[0]
    D = video_embeds.shape[-1]
    wrapped = []
    for i in range(B):
        parts = [video_embeds[i, 0], prompt_embeds[i]]
        if before_embeds is not None and before_embeds[i].shape[0] > 0:       # st_llm.py:392,396-397
            parts = [before_embeds[i]] + parts
        wrapped.append(torch.cat(parts, dim=0))
    lens = [wv.shape[0] for wv in wrapped]
    Lm = max(lens)
    A = answer_embeds.shape[1]
    off = 0 if bos_embed is None else 1
    embs = pad_embed.expand(B, off + Lm + A, D).clone()
    atts = torch.zeros(B, off + Lm + A, dtype=torch.int)
    targets = torch.full((B, off + Lm + A), -100, dtype=torch.long)
    part = answer_ids.masked_fill(answer_ids == pad_id, -100)
    for i in range(B):
        n = lens[i]
        if off:
            embs[i, 0] = bos_embed
            atts[i, 0] = 1
        embs[i, off:off + n] = wrapped[i]
        embs[i, off + n:off + n + A] = answer_embeds[i]
        # remaining positions keep pad_embed (the padded tail of the wrapped prompt, st_llm.py:420)
        atts[i, off:off + n] = 1
        atts[i, off + n:off + n + A] = answer_mask[i]
        targets[i, off + n:off + n + A] = part[i]
    return embs, atts, targets


# --------------------------------------------------------------------------------------------------
# whole-path composition on a full state dict with the reference's key names
# --------------------------------------------------------------------------------------------------


def split_state_dict(sd, dtype=torch.float32, device="cpu"):
    """STLLMForCausalLM.state_dict() (reference key names, st_llm.py / SURVEY section 5) -> oracle sub-dicts, fp32 CPU by
    default.  dtype / device select where the SAME arithmetic is evaluated (tests and bench.py's gpu_eager_fp16 leg run it in
    fp16 or fp32 on the GPU as the library bar / full-depth reference)."""
    f = {k: (v.detach().to(device=device, dtype=dtype) if v.is_floating_point() else v.detach().to(device=device))
         for k, v in sd.items()}
    pre = "model.stllm_model."

    def sub(prefix):
        return {k[len(prefix):]: v for k, v in f.items() if k.startswith(prefix)}
    w = {"vit": sub(pre + "visual_encoder."), "ln_vision": sub(pre + "ln_vision."),
         "qformer": sub(pre + "Qformer.bert."), "query_tokens": f.get(pre + "query_tokens"),
         "llama_proj": sub(pre + "llama_proj."), "residual": {k[len(pre):]: v for k, v in f.items()
                                                               if k.startswith(pre + "down_proj.") or k.startswith(pre + "up_proj.")},
         "mvm_decoder": sub(pre + "mvm_decoder."),
         "llama": {k[len("model."):]: v for k, v in f.items()
                   if k.startswith("model.") and not k.startswith(pre)},
         "lm_head": f.get("lm_head.weight")}
    return w


def video_prefill(w, frames, prompt_ids, cfg, video_input="all", residual_size=4, collect=None):
    """The inference path of bench.py / smoke(): frames [B,T,3,H,W] -> encode_img -> aggregate -> [video ‖ prompt]
    -> LlamaModel -> lm_head logits [B,S,V]   (conversation.py:274-293,322-340 + st_llm.py:118-122)."""
    emb = encode_img(w, frames, cfg, collect=collect)
    vid = video_aggregate(emb, video_input, residual_size, w["residual"])
    prompt = w["llama"]["embed_tokens.weight"][prompt_ids]
    seq = torch.cat([vid[:, 0], prompt], dim=1)
    if collect is not None:
        collect["inputs_embeds"] = seq
    hid = llama_forward(w["llama"], seq, None, cfg["llama_layers"], cfg["llama_heads"], collect=collect)
    return F.linear(hid, w["lm_head"])
