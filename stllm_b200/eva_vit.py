"""EVA-CLIP ViT-g/14 image encoder on the sm_100a kernels.

Mirrors the interface of the reference's stllm/models/eva_vit.py (class/ctor/parameter names and state-dict
keys: patch_embed.proj.*, cls_token, pos_embed, blocks.N.{norm1,attn.{qkv,q_bias,v_bias,proj},norm2,
mlp.{fc1,fc2}}) so checkpoints and callers are interchangeable.  nn.Module is used for parameter storage only;
forward() runs entirely in libstllm_b200.so:
  patch-embed  : im2col + tcgen05 GEMM (K 588 -> 640 zero-padded) + [cls;x]+pos     eva_vit.py:198-204,328-332
  per block    : LN -> qkv GEMM(+bias) -> fused attention (q pre-scaled in fp16) -> proj GEMM(+bias,+residual)
                 -> LN -> fc1 GEMM(+bias, GELU-erf epilogue) -> fc2 GEMM(+bias,+residual)   eva_vit.py:118-180
All frames of all clips are flattened to one token matrix [(B*T)*257, 1408], so each launch covers every frame.
"""
import math
import os
from functools import partial

import torch
import torch.nn as nn

from . import ops
from .packing import PackedWeightsMixin


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, std=std)


class PatchEmbed(nn.Module):
    """eva_vit.py:183-204"""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0])
        self.patch_shape = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.img_size = img_size
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class Attention(nn.Module):
    """parameter container of eva_vit.py:66-116 (qkv without bias + learned q_bias / v_bias)"""

    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.v_bias = nn.Parameter(torch.zeros(dim))
        else:
            self.q_bias = None
            self.v_bias = None
        self.proj = nn.Linear(dim, dim)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.gamma_1, self.gamma_2 = None, None


def pack_block(blk, dev, D, fold_norm1=False, fold_norm2=False):
    """device-side weights of one Block.  fold_norm*: the LayerNorm in front of qkv / fc1 is folded into that GEMM
    (ops.gemm norm=NORM_LN): W' = fp16(gamma (.) W), col_sb[n] = (sum_k W'[n, k], bias[n] + sum_k beta[k] W[n, k]),
    so that  LN(x) W^T + b = rstd (x W'^T - mu s) + b'  with the row statistics accumulated by the producing GEMM."""
    f16 = lambda t: t.detach().to(device=dev, dtype=torch.float16).contiguous()
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    a = blk.attn
    if a.q_bias is not None:
        qkv_bias = torch.cat((f16(a.q_bias), torch.zeros(D, dtype=torch.float16, device=dev), f16(a.v_bias)))
    else:
        qkv_bias = None

    def fold(w, bias, gamma, beta):
        wf = w.float()
        w2 = (wf * gamma[None, :]).to(torch.float16).contiguous()
        b2 = wf @ beta
        if bias is not None:
            b2 = b2 + bias.float()
        return w2, torch.stack((w2.float().sum(1), b2), dim=1).contiguous()
    bw = dict(g1=f32(blk.norm1.weight), b1=f32(blk.norm1.bias), g2=f32(blk.norm2.weight), b2=f32(blk.norm2.bias),
              w_qkv=f16(a.qkv.weight), qkv_bias=qkv_bias, w_proj=f16(a.proj.weight), b_proj=f16(a.proj.bias),
              w_fc1=f16(blk.mlp.fc1.weight), b_fc1=f16(blk.mlp.fc1.bias),
              w_fc2=f16(blk.mlp.fc2.weight), b_fc2=f16(blk.mlp.fc2.bias), qkv_sb=None, fc1_sb=None)
    if fold_norm1:
        bw["w_qkv"], bw["qkv_sb"] = fold(bw["w_qkv"], qkv_bias, bw["g1"], bw["b1"])
    if fold_norm2:
        bw["w_fc1"], bw["fc1_sb"] = fold(bw["w_fc1"], bw["b_fc1"], bw["g2"], bw["b2"])
    return bw


def run_block(bw, xt, h, B, n_tok, Hh, d, eps, stats_in=None, stats_mid=None, stats_out=None):
    """Block.forward (eva_vit.py:173-180) in place on the token matrix xt [B * n_tok, D]; h = scratch of that shape.
    stats_in: (sum, sumsq) rows of xt from the producing GEMM (needs bw['qkv_sb']), else norm1 runs standalone;
    stats_mid: zeroed buffer for the rows after proj (needs bw['fc1_sb']); stats_out: zeroed buffer for the rows
    after fc2 (consumed by the next block's folded norm1)."""
    D = Hh * d
    if stats_in is not None and bw["qkv_sb"] is not None:
        qkv = ops.gemm(xt, bw["w_qkv"], norm=ops.NORM_LN, rs_in=stats_in, col_sb=bw["qkv_sb"], norm_eps=eps)
    else:
        assert bw["qkv_sb"] is None, "block packed with a folded norm1 needs the row statistics of its input"
        ops.layernorm(xt, bw["g1"], bw["b1"], eps, out=h)
        qkv = ops.gemm(h, bw["w_qkv"], bias=bw["qkv_bias"])
    qkv = qkv.view(B, n_tok, 3, Hh, d)
    a = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale=1.0, q_prescale=d ** -0.5)
    if bw["fc1_sb"] is not None:
        ops.gemm(a.view(B * n_tok, D), bw["w_proj"], bias=bw["b_proj"], residual=xt, out=xt, rs_out=stats_mid,
                 rs_out_mode=2)
        m = ops.gemm(xt, bw["w_fc1"], epilogue=ops.EPI_GELU, norm=ops.NORM_LN, rs_in=stats_mid, col_sb=bw["fc1_sb"],
                     norm_eps=eps)
    else:
        ops.gemm(a.view(B * n_tok, D), bw["w_proj"], bias=bw["b_proj"], residual=xt, out=xt)
        ops.layernorm(xt, bw["g2"], bw["b2"], eps, out=h)
        m = ops.gemm(h, bw["w_fc1"], bias=bw["b_fc1"], epilogue=ops.EPI_GELU)
    ops.gemm(m, bw["w_fc2"], bias=bw["b_fc2"], residual=xt, out=xt, rs_out=stats_out,
             rs_out_mode=2 if stats_out is not None else 0)


class VisionTransformer(PackedWeightsMixin, nn.Module):
    """eva_vit.py:244-349.  Options the ST-LLM path never enables (rel-pos bias, layer-scale, drop-path > 0 in
    eval, checkpointing) are rejected loudly rather than silently ignored."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., norm_layer=nn.LayerNorm, init_values=None, use_abs_pos_emb=True,
                 use_rel_pos_bias=False, use_shared_rel_pos_bias=False, use_mean_pooling=True, init_scale=0.001,
                 use_checkpoint=False):
        super().__init__()
        if use_rel_pos_bias or use_shared_rel_pos_bias or init_values or qk_scale is not None:
            raise NotImplementedError("rel-pos bias / layer-scale / qk_scale are not part of the ST-LLM hot path")
        if not use_abs_pos_emb:
            raise NotImplementedError("use_abs_pos_emb=False is not part of the ST-LLM hot path")
        self.image_size = img_size
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.depth = depth
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.ln_eps = self.blocks[0].norm1.eps if depth > 0 else 1e-6
        # init as the reference does (eva_vit.py:288-315)
        _trunc_normal_(self.pos_embed, std=.02)
        _trunc_normal_(self.cls_token, std=.02)
        self.apply(self._init_weights)
        self.fix_init_weight()
        self._packed = None

    def fix_init_weight(self):
        for layer_id, layer in enumerate(self.blocks):
            layer.attn.proj.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))
            layer.mlp.fc2.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ---- packed device-side weights -------------------------------------------------------------
    def _apply(self, fn, *a, **kw):
        self._packed = None  # .to()/.half()/.cuda() invalidate the packed copies
        return super()._apply(fn, *a, **kw)

    def _pack(self):
        dev = self.cls_token.device
        f16 = lambda t: t.detach().to(device=dev, dtype=torch.float16).contiguous()
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        D = self.embed_dim
        pw = self.patch_embed.proj.weight
        kreal = pw[0].numel()
        kpad = (kreal + 63) // 64 * 64
        w_patch = torch.zeros(D, kpad, dtype=torch.float16, device=dev)
        w_patch[:, :kreal] = f16(pw).reshape(D, kreal)
        fuse = os.environ.get("STLLM_FUSE_NORM", "1") != "0"
        blocks = [pack_block(blk, dev, D, fold_norm1=fuse and i > 0, fold_norm2=fuse) for i, blk in enumerate(self.blocks)]
        self._packed = dict(w_patch=w_patch, kpad=kpad, kreal=kreal, b_patch=f16(self.patch_embed.proj.bias),
                            cls=f16(self.cls_token).reshape(-1), pos=f16(self.pos_embed).reshape(-1, D), blocks=blocks,
                            fuse=fuse)
        self._mark_packed()

    # ---- forward --------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_features(self, x, return_block_outputs=False):
        B, C, H, W = x.shape
        assert H == self.patch_embed.img_size[0] and W == self.patch_embed.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.patch_embed.img_size[0]}*{self.patch_embed.img_size[1]})."
        if not self._packed_valid():
            self._pack()
        P = self._packed
        if x.dtype not in (torch.float16, torch.float32):
            x = x.float()
        D, Hh = self.embed_dim, self.num_heads
        d = D // Hh
        n_tok = self.patch_embed.num_patches + 1
        cols = ops.im2col_patch(x.contiguous(), self.patch_embed.patch_size[0], P["kpad"])
        patches = ops.gemm(cols, P["w_patch"], bias=P["b_patch"], k_real=P["kreal"])
        xt = ops.vit_assemble(patches, P["cls"], P["pos"], B, self.patch_embed.num_patches).view(B * n_tok, D)
        outs = []
        h = torch.empty_like(xt)
        nb = len(P["blocks"])
        # folded LayerNorms: (sum, sum of squares) per token row after proj (slot 2i) and after fc2 (slot 2i + 1)
        stats = None
        if P["fuse"]:
            stats = ops.cached_row_stats(self, 2 * nb, B * n_tok, D, 2, xt.device)
        for i, bw in enumerate(P["blocks"]):
            run_block(bw, xt, h, B, n_tok, Hh, d, self.ln_eps,
                      stats_in=stats[2 * i - 1] if (stats is not None and i > 0) else None,
                      stats_mid=stats[2 * i] if stats is not None else None,
                      stats_out=stats[2 * i + 1] if (stats is not None and i + 1 < nb) else None)
            if return_block_outputs:
                outs.append(xt.view(B, n_tok, D).clone())
        out = xt.view(B, n_tok, D)
        return (out, outs) if return_block_outputs else out

    def forward(self, x):
        return self.forward_features(x)

    def get_num_layer(self, var_name=""):
        return len(self.blocks)


def convert_weights_to_fp16(model: nn.Module):
    """eva_vit.py:397-412: Conv/Linear weights and biases to fp16 (LayerNorm parameters stay fp32)"""
    def _c(l):
        if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            l.weight.data = l.weight.data.half()
            if l.bias is not None:
                l.bias.data = l.bias.data.half()
    model.apply(_c)


def create_eva_vit_g(img_size=224, drop_path_rate=0.4, use_checkpoint=False, precision="fp16", checkpoint_path=None,
                     depth=39):
    """eva_vit.py:415-443.  The reference downloads eva_vit_g.pth; offline, weights are random-initialised unless
    `checkpoint_path` points at a local copy of that file (same keys)."""
    model = VisionTransformer(
        img_size=img_size, patch_size=14, use_mean_pooling=False, embed_dim=1408, depth=depth, num_heads=1408 // 88,
        mlp_ratio=4.3637, qkv_bias=True, drop_path_rate=drop_path_rate, norm_layer=partial(nn.LayerNorm, eps=1e-6),
        use_checkpoint=use_checkpoint)
    if checkpoint_path is not None:
        state_dict = torch.load(checkpoint_path, map_location="cpu")
        interpolate_pos_embed(model, state_dict)
        model.load_state_dict(state_dict, strict=False)
    if precision == "fp16":
        convert_weights_to_fp16(model)
    return model


def interpolate_pos_embed(model, checkpoint_model):
    """eva_vit.py:373-394: bicubic resize of the grid part of pos_embed when the image size differs"""
    if "pos_embed" not in checkpoint_model:
        return
    pos_embed_checkpoint = checkpoint_model["pos_embed"].float()
    embedding_size = pos_embed_checkpoint.shape[-1]
    num_patches = model.patch_embed.num_patches
    num_extra_tokens = model.pos_embed.shape[-2] - num_patches
    orig_size = int((pos_embed_checkpoint.shape[-2] - num_extra_tokens) ** 0.5)
    new_size = int(num_patches ** 0.5)
    if orig_size != new_size:
        extra_tokens = pos_embed_checkpoint[:, :num_extra_tokens]
        pos_tokens = pos_embed_checkpoint[:, num_extra_tokens:]
        pos_tokens = pos_tokens.reshape(-1, orig_size, orig_size, embedding_size).permute(0, 3, 1, 2)
        pos_tokens = torch.nn.functional.interpolate(pos_tokens, size=(new_size, new_size), mode="bicubic",
                                                     align_corners=False)
        pos_tokens = pos_tokens.permute(0, 2, 3, 1).flatten(1, 2)
        checkpoint_model["pos_embed"] = torch.cat((extra_tokens, pos_tokens), dim=1)
