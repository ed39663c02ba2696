"""EVA-ViT-g with the BT-Adapter temporal branch (vit_model="eva_btadapter_g") on the sm_100a kernels.

Mirrors the reference's stllm/models/eva_btadapter.py: class / attribute / state-dict names
(patch_embed.*, cls_token, pos_embed, blocks.N.*, BTAdapter_cls, BTAdapter_S.N.{norm1,attn.*,norm2,mlp.*},
BTAdapter_T.N.{norm1,attn.*,temporal_fc}, BTAdapter_position.weight), the `forward(x)` input conventions
(eva_btadapter.py:228-250) and `self.T`.  It is the encoder of the shipped QA / MVBench configs
(stllm/configs/models/*_btadapter.yaml:5); SURVEY.md 8f rank 2.

Data layout (B200-first): the ViT stream stays the flattened token matrix [(B*T)*257, 1408]; the adapter branch is ONE
matrix [B + B*256*T, 1408] -- B clip-level cls rows followed by the patch rows in the reference's 'b (p t) m' order, so
  * the temporal attention of a patch position (eva_btadapter.py:300-302) reads T contiguous rows: the attention kernel
    runs with batch = B*256, seq = T straight on the q|k|v GEMM output (no gather),
  * LayerNorms and the MLP GEMMs run over all rows (cls + patches) in one launch,
  * only the spatial attention needs the frame-major order: one row gather in, one transposing residual add out
    (ops.bt_vit_to_branch), both moving whole 2816-byte rows.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .packing import PackedWeightsMixin
from .eva_vit import Attention, Block, create_eva_vit_g, convert_weights_to_fp16, pack_block, run_block


class BTAdapter_Spatial(Block):
    """eva_btadapter.py:253-279: a ViT Block (same parameters) applied per frame on [cls ; patches of the frame]"""

    def __init__(self, d_model, n_head, drop_num=0.1):
        super().__init__(dim=d_model, num_heads=n_head, mlp_ratio=4.3637, qkv_bias=True,
                         norm_layer=lambda d: nn.LayerNorm(d, eps=1e-5))   # Block's default norm_layer: nn.LayerNorm


class BTAdapter_Temp(nn.Module):
    """eva_btadapter.py:281-308"""

    def __init__(self, d_model, n_head, drop_num=0.1):
        super().__init__()
        self.attn = Attention(d_model, num_heads=n_head, qkv_bias=True)
        self.norm1 = nn.LayerNorm(d_model, eps=1e-6)
        self.temporal_fc = nn.Linear(d_model, d_model)
        nn.init.constant_(self.temporal_fc.weight, 0)      # constant_init(..., val=0, bias=0) :297
        nn.init.constant_(self.temporal_fc.bias, 0)


class EVAVisionTransformer_BTAdapter(PackedWeightsMixin, nn.Module):
    """eva_btadapter.py:46-250.  `clip` lets tests build the adapter around a shallower ViT-g-width encoder; the
    reference always wraps create_eva_vit_g(224, 0, False, "fp16") (:52-54)."""

    def __init__(self, depth=4, mask_rate=0, clip=None):
        super().__init__()
        if clip is None:
            clip = create_eva_vit_g(224, 0, False, "fp16")
        if mask_rate > 0:
            raise NotImplementedError("tube masking (training-time only, eva_btadapter.py:243-246) is outside the hot path")
        self.image_size = clip.image_size
        self.num_classes = clip.num_classes
        self.num_features = self.embed_dim = clip.embed_dim
        self.num_heads = clip.num_heads
        self.patch_embed = clip.patch_embed
        self.num_patches = self.patch_embed.num_patches
        self.cls_token = clip.cls_token
        self.pos_embed = clip.pos_embed
        self.blocks = clip.blocks
        self.ln_eps = clip.ln_eps
        self._clip = [clip]          # not a submodule: patch-embed packing / helpers only
        self.depth = depth
        self.mask_rate = mask_rate
        dpr = np.linspace(0, 0.1, depth)
        self.BTAdapter_cls = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
        self.BTAdapter_S = nn.ModuleList([BTAdapter_Spatial(self.embed_dim, self.num_heads, drop_num=dpr[i])
                                          for i in range(depth)])
        self.BTAdapter_T = nn.ModuleList([BTAdapter_Temp(self.embed_dim, self.num_heads, drop_num=dpr[i])
                                          for i in range(depth)])
        self.BTAdapter_position = nn.Embedding(64, self.embed_dim)
        self.init_weights()
        self.T = 1
        self._packed = None
        self._idx = {}

    def init_weights(self):
        """eva_btadapter.py:89-99: the spatial adapters start as copies of the last `depth` ViT blocks"""
        total_depth = len(self.blocks)
        self.num_layers = total_depth
        load_start = total_depth - self.depth
        spatial = {}
        for k, v in self.blocks.state_dict().items():
            n = int(k.split(".")[0])
            if n >= load_start:
                spatial[k.replace(str(n), str(n - load_start), 1)] = v.clone()
        self.BTAdapter_S.load_state_dict(spatial)

    def get_num_layers(self):
        return len(self.blocks)

    # ---- packed device-side weights -------------------------------------------------------------
    def _apply(self, fn, *a, **kw):
        self._packed = None
        return super()._apply(fn, *a, **kw)

    def _pack(self):
        dev = self.cls_token.device
        D = self.embed_dim
        f16 = lambda t: t.detach().to(device=dev, dtype=torch.float16).contiguous()
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        clip = self._clip[0]
        clip._packed = None
        clip._pack()                                   # patch-embed weights, cls, pos, ViT blocks (with folded norms)
        P = dict(clip._packed)
        P["bt_cls"] = f16(self.BTAdapter_cls).reshape(-1)
        P["bt_time"] = f16(self.BTAdapter_position.weight)
        P["S"] = [pack_block(blk, dev, D) for blk in self.BTAdapter_S]
        P["S_eps"] = [blk.norm1.eps for blk in self.BTAdapter_S]
        P["T"] = []
        for blk in self.BTAdapter_T:
            a = blk.attn
            qkv_bias = torch.cat((f16(a.q_bias), torch.zeros(D, dtype=torch.float16, device=dev), f16(a.v_bias)))
            P["T"].append(dict(g1=f32(blk.norm1.weight), b1=f32(blk.norm1.bias), eps=blk.norm1.eps,
                               w_qkv=f16(a.qkv.weight), qkv_bias=qkv_bias, w_proj=f16(a.proj.weight),
                               b_proj=f16(a.proj.bias), w_fc=f16(blk.temporal_fc.weight),
                               b_fc=f16(blk.temporal_fc.bias)))
        self._packed = P
        self._mark_packed()

    def _spatial_index(self, B, T, P_, dev):
        """row gather branch layout -> frame-major [(b t), 1 + p]: every frame gets the clip's cls row first
        ('b (p t) m -> (b t) p m' + the repeated cls token, eva_btadapter.py:267-270)"""
        key = (B, T, P_, str(dev))
        if key not in self._idx:
            b = torch.arange(B).view(B, 1, 1)
            t = torch.arange(T).view(1, T, 1)
            l = torch.arange(P_).view(1, 1, P_)
            patch = B + (b * P_ + l) * T + t                                  # [B, T, P]
            cls = b.expand(B, T, 1)
            self._idx[key] = torch.cat((cls, patch), dim=2).reshape(-1).to(device=dev, dtype=torch.int64)
        return self._idx[key]

    # ---- forward --------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_features(self, x, B, T, return_branches=False):
        clip = self._clip[0]
        n_img, C, H, W = x.shape
        assert H == self.patch_embed.img_size[0] and W == self.patch_embed.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.patch_embed.img_size[0]}*{self.patch_embed.img_size[1]})."
        assert T <= self.BTAdapter_position.num_embeddings, "BTAdapter_position holds 64 frame positions"
        if not self._packed_valid():
            self._pack()
        P = self._packed
        if x.dtype not in (torch.float16, torch.float32):
            x = x.float()
        D, Hh = self.embed_dim, self.num_heads
        d = D // Hh
        NP = self.num_patches
        n_tok = NP + 1
        cols = ops.im2col_patch(x.contiguous(), self.patch_embed.patch_size[0], P["kpad"])
        patches = ops.gemm(cols, P["w_patch"], bias=P["b_patch"], k_real=P["kreal"])
        xt = ops.vit_assemble(patches, P["cls"], P["pos"], n_img, NP).view(n_img * n_tok, D)
        h = torch.empty_like(xt)
        nb = len(P["blocks"])
        stats = None
        if P["fuse"]:
            stats = ops.cached_row_stats(self, 2 * nb, n_img * n_tok, D, 2, xt.device)
        z = None
        zh = None
        branches = []
        idx_sp = self._spatial_index(B, T, NP, xt.device)
        for i, bw in enumerate(P["blocks"]):
            run_block(bw, xt, h, n_img, n_tok, Hh, d, self.ln_eps,
                      stats_in=stats[2 * i - 1] if (stats is not None and i > 0) else None,
                      stats_mid=stats[2 * i] if stats is not None else None,
                      stats_out=stats[2 * i + 1] if (stats is not None and i + 1 < nb) else None)
            if i >= nb - self.depth:
                j = i + self.depth - nb
                # forward_branch (:183-201): bring the ViT stream into the branch layout (+ previous branch / init)
                if z is None:
                    z = ops.bt_vit_to_branch(xt, B, T, NP, pos=P["pos"], time_emb=P["bt_time"], bt_cls=P["bt_cls"])
                    zh = torch.empty_like(z)
                else:
                    ops.bt_vit_to_branch(xt, B, T, NP, z_prev=z, out=z)
                self._temporal(P["T"][j], z, zh, B, T, NP, Hh, d)
                self._spatial(P["S"][j], P["S_eps"][j], z, zh, h, idx_sp, B, T, NP, Hh, d)
                if return_branches:
                    branches.append(torch.cat((z[:B].view(B, 1, D), z[B:].view(B, NP * T, D)), dim=1).clone())
        out = ops.bt_branch_to_vit_avg(xt, z, B, T, NP).view(n_img, n_tok, D)
        return (out, branches) if return_branches else out

    @staticmethod
    def _temporal(tw, z, zh, B, T, NP, Hh, d):
        """BTAdapter_Temp.forward (:291-308) on the patch rows of z (the B cls rows pass through)"""
        D = Hh * d
        zp, hp = z[B:], zh[B:]
        ops.layernorm(zp, tw["g1"], tw["b1"], tw["eps"], out=hp)
        qkv = ops.gemm(hp, tw["w_qkv"], bias=tw["qkv_bias"]).view(B * NP, T, 3, Hh, d)
        a = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale=1.0, q_prescale=d ** -0.5)
        y = ops.gemm(a.view(B * NP * T, D), tw["w_proj"], bias=tw["b_proj"])
        ops.gemm(y, tw["w_fc"], bias=tw["b_fc"], residual=zp, out=zp)

    @staticmethod
    def _spatial(bw, eps, z, zh, h, idx_sp, B, T, NP, Hh, d):
        """BTAdapter_Spatial.forward (:258-279)"""
        D = Hh * d
        n_tok = NP + 1
        ops.layernorm(z, bw["g1"], bw["b1"], eps, out=zh)          # LayerNorm is per row: norm first, then reorder
        s = ops.gather_rows(zh, idx_sp, out=h)                      # [(b t), 1 + p, D] frame-major, cls row repeated
        qkv = ops.gemm(s, bw["w_qkv"], bias=bw["qkv_bias"]).view(B * T, n_tok, 3, Hh, d)
        a = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale=1.0, q_prescale=d ** -0.5)
        y = ops.gemm(a.view(B * T * n_tok, D), bw["w_proj"], bias=bw["b_proj"])
        ops.bt_vit_to_branch(y, B, T, NP, z_prev=z, out=z)          # residual + [mean_t cls ; patches back in (p t) order]
        ops.layernorm(z, bw["g2"], bw["b2"], eps, out=zh)
        m = ops.gemm(zh, bw["w_fc1"], bias=bw["b_fc1"], epilogue=ops.EPI_GELU)
        ops.gemm(m, bw["w_fc2"], bias=bw["b_fc2"], residual=z, out=z)

    def forward(self, x, return_all_features=False):
        """input conventions of eva_btadapter.py:228-241: [B,T,3,H,W], [B,3,T,H,W], [T,3,H,W] (one clip) or images"""
        if x.ndim == 5:
            if x.shape[1] == 3:
                x = x.permute(0, 2, 1, 3, 4)
            B, T = x.shape[:2]
            x = x.reshape((-1,) + tuple(x.shape[2:]))
        elif x.ndim == 4:
            T, B = x.shape[0], 1
        else:
            raise ValueError("expected a 4-D or 5-D frame tensor")
        self.T = T
        return self.forward_features(x, B, T)


def create_eva_btadapter(precision="fp16", clip=None, depth=3):
    """eva_btadapter.py:310-315 (depth=3)"""
    model = EVAVisionTransformer_BTAdapter(depth=depth, clip=clip)
    if precision == "fp16":
        convert_weights_to_fp16(model)
    return model
