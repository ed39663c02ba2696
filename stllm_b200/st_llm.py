"""ST-LLM model API on the sm_100a kernels: STLLMModel / STLLMLlamaModel / STLLMForCausalLM.

Drop-in for the reference's stllm/models/st_llm.py: same class names, attribute graph
(`model.model.stllm_model.{encode_img, video_input, residual_size, up_proj, down_proj, non_linear_func,
embed_tokens, llama_tokenizer, qformer_text_input, mask, img_len, mask_img_len, mvm_decoder}`, `model.get_model()`),
state-dict keys (`model.stllm_model.visual_encoder.*`, `.ln_vision.*`, `.Qformer.bert.*`, `.query_tokens`,
`.llama_proj.*`, `.up_proj/.down_proj`, `.mvm_decoder.*`, `model.layers.*`, `lm_head.*`), registry name
("st_llm_hf") and error behaviour (asserts of st_llm.py:336, eva_vit.py:201).  nn.Modules hold parameters only;
all arithmetic of the path runs in libstllm_b200.so through stllm_b200.ops (no torch math, no CPU fallback).
Host-side integer logic that must be bit-exact with the reference (dynamic mask RNG order, residual frame index)
stays in numpy exactly as the reference does it (st_llm.py:434-445,482-493; utils.py:4-16).
"""
import math
import os
import re

import numpy as np
import torch
import torch.nn as nn

from . import ops
from . import prompt_assembly as PA
from .eva_vit import create_eva_vit_g, VisionTransformer
from .llama import (BaseModelOutputWithPast, CausalLMOutputWithPast, KVCache, LlamaConfig, LlamaModel)
from .qformer import BertConfig, BertLMHeadModel
from .registry import registry


class StllmConfig(LlamaConfig):
    model_type = "st_llm_hf"


def RandomMaskingGenerator(num_patches, mask_ratio, batch, device="cuda"):
    """utils.py:4-16 -- identical consumption of numpy's global RNG (one shuffle per row); bool [batch, L]"""
    num_mask = int(mask_ratio * num_patches)
    mask_list = []
    for _ in range(batch):
        mask = np.hstack([np.zeros(num_patches - num_mask), np.ones(num_mask)])
        np.random.shuffle(mask)
        mask_list.append(mask)
    return torch.from_numpy(np.array(mask_list)).to(torch.bool).to(device)


class LayerNorm(nn.LayerNorm):
    """blip2.py:103-109 (`ln_vision`): fp32 statistics, output cast back to the input dtype"""

    @torch.no_grad()
    def forward(self, x):
        return ops.layernorm(x if x.dtype == torch.float16 else x.half(), self.weight.float(), self.bias.float(),
                             self.eps)


class Linear_Decoder(nn.Module):
    """st_llm.py:35-43"""

    def __init__(self, output_dim=4096, embed_dim=4096):
        super().__init__()
        self.head = nn.Linear(embed_dim, output_dim)
        self.norm = nn.LayerNorm(output_dim)

    @torch.no_grad()
    def forward(self, x):
        y = ops.gemm(x, self.head.weight.half(), bias=self.head.bias.half())
        return ops.layernorm(y, self.norm.weight.float(), self.norm.bias.float(), self.norm.eps)


def _linear(mod, x, epilogue=ops.EPI_NONE):
    w = mod.weight if mod.weight.dtype == torch.float16 else mod.weight.half()
    b = None if mod.bias is None else (mod.bias if mod.bias.dtype == torch.float16 else mod.bias.half())
    return ops.gemm(x, w, bias=b, epilogue=epilogue)


class STLLMModel(nn.Module):
    """st_llm.py:205-606 (Blip2Base subclass in the reference)"""

    def __init__(self, vit_model="eva_clip_g", q_former_model=None, img_size=224, pre_encoding=False, use_mask=False,
                 mvm_decode=False, video_input=None, residual_size=4, qformer_text_input=False, drop_path_rate=0,
                 use_grad_checkpoint=False, vit_precision="fp16", freeze_vit=True, has_qformer=True,
                 freeze_qformer=True, num_query_token=32, llama_model="", max_txt_len=32, end_sym="\n",
                 # offline / test hooks (the reference downloads these):
                 tokenizer=None, llama_tokenizer=None, vit_depth=39, qformer_config=None, vit_checkpoint=None,
                 llama_hidden=4096):
        super().__init__()
        assert vit_model in ["eva_clip_g", "eva_btadapter_g"], \
            "vit model must be eva_clip_g for current version of MiniGPT-4"          # blip2.py:65
        self.tokenizer = tokenizer
        self.pre_encoding = pre_encoding
        self.video_input = video_input
        self.use_mask = use_mask
        self.mvm_decode = mvm_decode
        self.qformer_text_input = qformer_text_input
        self.residual_size = residual_size
        D = llama_hidden
        if self.video_input == "residual":
            self.down_proj = nn.Linear(D, 1024)
            self.non_linear_func = nn.ReLU()
            self.up_proj = nn.Linear(1024, D)
            nn.init.kaiming_uniform_(self.down_proj.weight, a=math.sqrt(5))
            nn.init.zeros_(self.up_proj.weight)
            nn.init.zeros_(self.down_proj.bias)
            nn.init.zeros_(self.up_proj.bias)
        if self.mvm_decode:
            self.mvm_decoder = Linear_Decoder(D, D)
        self.vit_model = vit_model
        self.visual_encoder = create_eva_vit_g(img_size, drop_path_rate, use_grad_checkpoint, vit_precision,
                                               checkpoint_path=vit_checkpoint, depth=vit_depth)
        if vit_model == "eva_btadapter_g":                                          # blip2.py:66-71
            from .eva_btadapter import create_eva_btadapter
            self.visual_encoder = create_eva_btadapter(vit_precision, clip=self.visual_encoder)
        self.ln_vision = LayerNorm(self.visual_encoder.num_features)
        self.has_qformer = has_qformer
        self.llama_tokenizer = llama_tokenizer
        if has_qformer:
            qcfg = qformer_config or BertConfig()
            qcfg.encoder_width = self.visual_encoder.num_features
            qcfg.query_length = num_query_token
            self.Qformer = BertLMHeadModel(qcfg)
            self.query_tokens = nn.Parameter(torch.zeros(1, num_query_token, qcfg.hidden_size))
            self.query_tokens.data.normal_(mean=0.0, std=qcfg.initializer_range)
            if not qformer_text_input:                                   # st_llm.py:277-283
                self.Qformer.bert.embeddings.word_embeddings = None
                self.Qformer.bert.embeddings.position_embeddings = None
                for layer in self.Qformer.bert.encoder.layer:
                    layer.output = None
                    layer.intermediate = None
            elif tokenizer is not None:
                self.Qformer.resize_token_embeddings(len(tokenizer))    # st_llm.py:285
            self.Qformer.cls = None
            img_f_dim = qcfg.hidden_size                                 # st_llm.py:298
        else:
            img_f_dim = self.visual_encoder.num_features * 4             # st_llm.py:300-301: no Q-Former, no query tokens
        self.llama_proj = nn.Linear(img_f_dim, D)
        self.max_txt_len = max_txt_len
        self.end_sym = end_sym
        self.embed_tokens = None  # set by STLLMLlamaModel.initialize_vision_modules (st_llm.py:54)
        for p in self.parameters():
            p.requires_grad = False

    # keep embed_tokens out of this module's parameters/state_dict (it belongs to the LLM, as in the reference
    # where the attribute is assigned after construction and shares the LLM's tensor)
    def __setattr__(self, name, value):
        if name == "embed_tokens":
            object.__setattr__(self, name, value)
        else:
            super().__setattr__(name, value)

    @property
    def device(self):
        return self.llama_proj.weight.device

    def _embed(self, ids):
        w = self.embed_tokens.weight
        w = w if w.dtype == torch.float16 else w.half()
        return ops.gather_rows(w, ids.reshape(-1).to(w.device)).view(tuple(ids.shape) + (w.shape[1],))

    @torch.no_grad()
    def encode_img(self, image, text=None):
        """st_llm.py:321-377"""
        device = image.device
        T = image.shape[1]
        infer = True if len(image.shape) == 4 else False
        use_image = True if T == 1 or (len(image.shape) == 4) else False
        if (not use_image or len(image.shape) == 5) and (self.vit_model == "eva_clip_g"):
            image = image.reshape((-1,) + tuple(image.shape[2:]))   # 'B T C H W -> (B T) C H W'
        image_embeds = self.visual_encoder(image)
        image_embeds = self.ln_vision(image_embeds)
        if not self.has_qformer:                                     # st_llm.py:369-373 (SURVEY a15')
            image_embeds = image_embeds[:, 1:, :]                    # drop the cls token
            bs, pn, hs = image_embeds.shape
            image_embeds = image_embeds.reshape(bs, int(pn / 4), int(hs * 4)).contiguous()   # 4 neighbouring tokens -> 1
            inputs_llama = _linear(self.llama_proj, image_embeds)    # GEMM 4*C -> D
            if not infer:
                inputs_llama = inputs_llama.view(-1, T, inputs_llama.shape[1], inputs_llama.shape[-1])
            atts_llama = torch.ones(inputs_llama.size()[:-1], dtype=torch.long, device=device)
            return inputs_llama, atts_llama, use_image
        query_tokens = self.query_tokens.expand(image_embeds.shape[0], -1, -1)
        if self.qformer_text_input:
            assert text
            if isinstance(text, str):
                text = [text] * query_tokens.size(0)
            elif len(text) != query_tokens.size(0):
                text_ = []
                for t in text:
                    text_ += [t] * T
                text = text_
            text_Qformer = self.tokenizer(text, padding="longest", truncation=True, max_length=self.max_txt_len,
                                          return_tensors="pt").to(device)
            query_atts = torch.ones(query_tokens.size()[:-1], dtype=torch.long, device=device)
            Qformer_atts = torch.cat([query_atts, text_Qformer.attention_mask], dim=1)
            query_output = self.Qformer.bert(text_Qformer.input_ids, attention_mask=Qformer_atts,
                                             query_embeds=query_tokens, encoder_hidden_states=image_embeds,
                                             return_dict=True)
        else:
            query_output = self.Qformer.bert(query_embeds=query_tokens, encoder_hidden_states=image_embeds,
                                             return_dict=True)
        Q = query_tokens.size(1)
        last = query_output.last_hidden_state
        q_hidden = last if last.shape[1] == Q else last[:, :Q].contiguous()
        inputs_llama = _linear(self.llama_proj, q_hidden)
        if not infer:
            inputs_llama = inputs_llama.view(-1, T, Q, inputs_llama.shape[-1])
        atts_llama = torch.ones(inputs_llama.size()[:-1], dtype=torch.long, device=device)
        return inputs_llama, atts_llama, use_image

    # the north_star's name for the same function (VideoChat2 naming)
    encode_videoQformer_visual = encode_img

    # ---- sequence assembly: host-built row plan + one two-source gather per LLM pass (prompt_assembly.py) ----------
    def _prompt_token_ids(self, prompts, batch):
        """token ids before / after `<ImageHere>` per sample (st_llm.py:385-389); every distinct prompt is tokenised once"""
        if isinstance(prompts, str):
            prompts = [prompts] * batch
        cache, before, after = {}, [], []
        for prompt in list(prompts)[:batch]:
            if prompt not in cache:
                p_before, p_after = PA.split_prompt(prompt)
                tb = self.llama_tokenizer(p_before, return_tensors="pt", add_special_tokens=False).input_ids
                ta = self.llama_tokenizer(p_after, return_tensors="pt",
                                          add_special_tokens=self.qformer_text_input).input_ids
                cache[prompt] = (tb.reshape(-1).cpu().numpy().astype(np.int64),
                                 ta.reshape(-1).cpu().numpy().astype(np.int64))
            before.append(cache[prompt][0])
            after.append(cache[prompt][1])
        return before, after

    def _gather_plan(self, rows, video2d):
        """rows: int64 [B, S] row codes (device) -> [B, S, D] embeddings in ONE launch"""
        w = self.embed_tokens.weight
        w = w if w.dtype == torch.float16 else w.half()
        out = ops.gather_rows2(w, video2d, rows.reshape(-1))
        return out.view(rows.shape[0], rows.shape[1], w.shape[1])

    @torch.no_grad()
    def prompt_wrap(self, img_embeds, atts_img, prompts):
        """st_llm.py:379-407 for callers that hold embeddings: [before | img | after] per sample, right-padded with
        embed(pad).  Not used by forward(), which plans the whole sequence (answer, BOS, padding) at once."""
        if not prompts:
            return img_embeds, atts_img
        B = img_embeds.shape[0]
        vid = img_embeds.reshape(B, -1, img_embeds.shape[-1])
        before, after = self._prompt_token_ids(prompts, B)
        plan = PA.plan_sequences(PA.video_row_codes(B, vid.shape[1]), before, after, None, None,
                                 self.llama_tokenizer.pad_token_id)
        dev = img_embeds.device
        embs = self._gather_plan(torch.from_numpy(plan.rows).to(dev), vid.reshape(-1, vid.shape[-1]).contiguous())
        return embs, torch.from_numpy(plan.attention_mask).to(device=dev, dtype=torch.int)

    @torch.no_grad()
    def concat_emb_input_output(self, input_embs, input_atts, output_embs, output_atts):
        """st_llm.py:409-432 for callers that hold embeddings: splice the answer block in after each sample's valid
        inputs -- one two-source gather over (input rows, answer rows)."""
        B, W, D = input_embs.shape
        La = output_embs.shape[1]
        atts_host = input_atts.cpu().numpy()
        rows = torch.from_numpy(PA.splice_rows(atts_host, La)).to(input_embs.device)
        embs = ops.gather_rows2(input_embs.reshape(B * W, D).contiguous(), output_embs.reshape(B * La, D).contiguous(),
                                rows.reshape(-1)).view(B, W + La, D)
        n = atts_host.sum(axis=1)
        out_atts = output_atts.cpu().numpy().astype(np.int64)
        pos = np.arange(W + La)[None, :]
        atts = np.where(pos < n[:, None], 1, 0).astype(np.int64)
        in_ans = (pos >= n[:, None]) & (pos < n[:, None] + La)
        atts[in_ans] = out_atts.reshape(-1)
        return embs, torch.from_numpy(atts).to(input_embs.device), [int(v) for v in n]

    def get_residual_index(self, sample_segments, total_segments, devices):
        """st_llm.py:434-445"""
        if hasattr(self, "residual_index"):
            return self.residual_index
        seg_size = float(total_segments) / sample_segments
        frame_indices = np.array([int((seg_size / 2) + np.round(seg_size * idx)) for idx in range(sample_segments)])
        frame_indices = torch.from_numpy(frame_indices).to(devices)
        self.register_buffer("residual_index", frame_indices)
        return frame_indices

    @torch.no_grad()
    def aggregate_video(self, img_embeds):
        """st_llm.py:463-478 (inference twin conversation.py:281-293): [B,T,Q,D] -> [B,1,L,D]"""
        T = img_embeds.size(1)
        if self.video_input == "all":
            return img_embeds.reshape(img_embeds.size(0), 1, -1, img_embeds.size(-1)).contiguous()
        if self.video_input == "mean":
            return ops.mean_frames(img_embeds.contiguous()).unsqueeze(1)
        if self.video_input == "residual":
            residual_index = self.get_residual_index(self.residual_size, T, img_embeds.device)
            g = ops.mean_frames(img_embeds.contiguous())                      # [B,Q,D]
            g = _linear(self.up_proj, _linear(self.down_proj, g, ops.EPI_RELU))
            out = ops.residual_merge(img_embeds.contiguous(), g, residual_index)   # [B,R,Q,D]
            return out.view(img_embeds.size(0), 1, -1, img_embeds.size(-1))
        return img_embeds

    @torch.no_grad()
    def forward(self, samples):
        """st_llm.py:447-546: samples -> (inputs_embeds, attention_mask, unmask_inputs_embeds, unmask_attention_mask,
        targets).  Encoder and aggregation run in the kernels; everything after them is ONE host-built row plan per LLM
        pass (prompt_assembly.plan_sequences) + one gather launch, instead of the reference's per-sample cat/pad loops."""
        image = samples["image"]
        prompts = samples["instruction_input"] if "instruction_input" in samples else None
        use_image = False
        if self.pre_encoding:                                            # st_llm.py:452-455
            video = _linear(self.llama_proj, image.half())
        else:
            q_text = [it.split("Human: ")[1].split(" ###")[0] for it in prompts] if self.qformer_text_input else None
            video, _, use_image = self.encode_img(image, q_text)
        if not use_image:
            video = self.aggregate_video(video)
        if video.dim() != 4:
            raise ValueError("video tokens must be [B, 1, L, D] here (set video_input to all | mean | residual)")
        B, _, L, D = video.size()
        dev = video.device
        keep = None
        if not use_image and self.use_mask:                              # st_llm.py:482-493 (numpy RNG order preserved)
            self.img_len = L
            rate = np.random.normal(0.5, 0.1)
            mask_rate = float(np.clip(rate, 0.1, 0.7)) if getattr(self, "fixed_mask_rate", None) is None \
                else float(self.fixed_mask_rate)
            mask_cpu = RandomMaskingGenerator(L, mask_rate, B, "cpu").unsqueeze(1)
            self.mask = mask_cpu.to(dev)
            keep = ~mask_cpu.reshape(B, L).numpy()
            self._keep_host = keep
            self.mask_img_len = int(keep[0].sum())
        if not prompts:
            raise ValueError("samples['instruction_input'] with an <ImageHere> placeholder is required")
        tok = self.llama_tokenizer
        before, after = self._prompt_token_ids(prompts, B)
        tok.padding_side = "right"                                       # st_llm.py:497
        text = [t + (tok.eos_token if self.qformer_text_input else self.end_sym) for t in samples["answer"]]
        ans = tok(text, return_tensors="pt", padding="longest", truncation=True, max_length=self.max_txt_len,
                  add_special_tokens=False)
        a_ids, a_att = ans.input_ids.cpu().numpy(), ans.attention_mask.cpu().numpy()
        bos = None if self.qformer_text_input else tok.bos_token_id       # st_llm.py:519-530
        plans = [PA.plan_sequences(PA.video_row_codes(B, L, keep), before, after, a_ids, a_att, tok.pad_token_id, bos)]
        if keep is not None:
            plans.append(PA.plan_sequences(PA.video_row_codes(B, L), before, after, a_ids, a_att, tok.pad_token_id, bos))
        # one host -> device copy for all integer planes of all passes
        flat = np.concatenate([np.concatenate([p.rows.reshape(-1), p.attention_mask.reshape(-1)]) for p in plans]
                              + [plans[0].targets.reshape(-1)])
        packed = torch.from_numpy(flat).to(dev)
        video2d = video.reshape(B * L, D)
        video2d = video2d if video2d.is_contiguous() else video2d.contiguous()
        outs, o = [], 0
        for p in plans:
            n = p.rows.size
            rows = packed[o:o + n].view(p.rows.shape)
            att = packed[o + n:o + 2 * n].view(p.rows.shape)
            o += 2 * n
            outs.append((self._gather_plan(rows, video2d), att))
        targets = packed[o:].view(plans[0].targets.shape)
        inputs_embeds, attention_mask = outs[0]
        unmask_inputs_embeds, unmask_attention_mask = outs[1] if keep is not None else (None, None)
        return inputs_embeds, attention_mask, unmask_inputs_embeds, unmask_attention_mask, targets

    @classmethod
    def from_config(cls, cfg, **overrides):
        """st_llm.py:549-606: same keys / defaults"""
        kw = dict(
            vit_model=cfg.get("vit_model", "eva_clip_g"),
            q_former_model=cfg.get("q_former_model", None),
            img_size=cfg.get("image_size", 224),
            num_query_token=cfg.get("num_query_token", 32),
            llama_model=cfg.get("llama_model", ""),
            drop_path_rate=cfg.get("drop_path_rate", 0),
            use_grad_checkpoint=cfg.get("use_grad_checkpoint", False),
            vit_precision=cfg.get("vit_precision", "fp16"),
            freeze_vit=cfg.get("freeze_vit", True),
            has_qformer=cfg.get("has_qformer", True),
            freeze_qformer=cfg.get("freeze_qformer", True),
            max_txt_len=cfg.get("max_txt_len", 32),
            end_sym=cfg.get("end_sym", "\n"),
            pre_encoding=cfg.get("pre_encoding", False),
            use_mask=cfg.get("use_mask", False),
            mvm_decode=cfg.get("mvm_decode", False),
            video_input=cfg.get("video_input", None),
            residual_size=cfg.get("residual_size", 4),
            qformer_text_input=cfg.get("qformer_text_input", False),
        )
        for k in ("tokenizer", "llama_tokenizer", "vit_depth", "qformer_config", "vit_checkpoint"):
            if cfg.get(k, None) is not None:
                kw[k] = cfg.get(k)
        kw.update(overrides)
        return cls(**kw)


class STLLMLlamaModel(LlamaModel):
    """st_llm.py:45-92"""
    config_class = StllmConfig

    def initialize_vision_modules(self, cfg, **overrides):
        self.stllm_model = STLLMModel.from_config(cfg, llama_hidden=self.config.hidden_size, **overrides)
        if cfg.get("qformer_text_input", False) and self.stllm_model.llama_tokenizer is not None:
            self.resize_token_embeddings(len(self.stllm_model.llama_tokenizer))
        self.stllm_model.embed_tokens = self.embed_tokens

    def resize_token_embeddings(self, n):
        old = self.embed_tokens
        if n == old.num_embeddings:
            return old
        new = nn.Embedding(n, old.embedding_dim).to(device=old.weight.device, dtype=old.weight.dtype)
        new.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        m = min(n, old.num_embeddings)
        new.weight.data[:m] = old.weight.data[:m]
        self.embed_tokens = new
        self.config.vocab_size = n
        self.vocab_size = n
        if hasattr(self, "stllm_model"):
            self.stllm_model.embed_tokens = new
        return new

    @torch.no_grad()
    def forward(self, samples=None, inputs_embeds=None, **kwargs):
        if samples is None:
            return super().forward(inputs_embeds=inputs_embeds, **kwargs)
        inputs_embeds, attention_mask, unmask_inputs_embeds, unmask_attention_mask, labels = self.stllm_model(samples)
        if unmask_inputs_embeds is None:
            outputs = super().forward(input_ids=None, attention_mask=attention_mask, inputs_embeds=inputs_embeds,
                                      use_cache=False, output_hidden_states=False, return_dict=True)
            return outputs, None, labels
        sm = self.stllm_model
        rows = self.mvm_select_rows(sm._keep_host, unmask_inputs_embeds.shape[1]).to(inputs_embeds.device)
        outputs, loss_mvm = self.mvm_passes(inputs_embeds, attention_mask, unmask_inputs_embeds, unmask_attention_mask,
                                            rows, sm.mask_img_len, output_hidden_states=True)
        return outputs, loss_mvm, labels

    def mvm_select_rows(self, keep_host, unmask_seq_len):
        """row numbers, in the flattened [B * S_u] unmasked hidden states, of the VISIBLE video positions:
        `unmask_output[:, img_start:img_start + img_len][~mask]` (st_llm.py:85-86) as a host-built index (bit-exact)"""
        img_start = 0 if self.stllm_model.qformer_text_input else 8                   # st_llm.py:71
        B, L = keep_host.shape
        b, l = np.nonzero(keep_host)                                                  # row-major: the order of [~mask]
        return torch.from_numpy((b * unmask_seq_len + img_start + l).astype(np.int64))

    @torch.no_grad()
    def mvm_passes(self, inputs_embeds, attention_mask, unmask_inputs_embeds, unmask_attention_mask, rows, mask_img_len,
                   output_hidden_states=False):
        """st_llm.py:62-91, device work only (no host synchronisation: capturable in a CUDA graph): masked LLM pass ->
        MVM decoder on the visible video positions; unmasked LLM pass -> same positions gathered; cosine loss."""
        sm = self.stllm_model
        img_start = 0 if sm.qformer_text_input else 8                                  # st_llm.py:71
        outputs = super().forward(input_ids=None, attention_mask=attention_mask, inputs_embeds=inputs_embeds,
                                  use_cache=False, output_hidden_states=output_hidden_states, return_dict=True)
        mask_output = outputs.last_hidden_state                                       # == hidden_states[-1] (post-norm)
        B, _, D = mask_output.size()
        mask_img_output = mask_output[:, img_start:img_start + mask_img_len].contiguous()
        if hasattr(sm, "mvm_decoder"):
            mask_img_output = sm.mvm_decoder(mask_img_output)
        unmask_outputs = super().forward(inputs_embeds=unmask_inputs_embeds, attention_mask=unmask_attention_mask,
                                         return_dict=True, use_cache=False, output_hidden_states=output_hidden_states)
        unmask_output = unmask_outputs.last_hidden_state
        S_u = unmask_output.shape[1]
        unmask_img_output = ops.gather_rows(unmask_output.reshape(B * S_u, D), rows)
        loss_mvm = ops.cosine_loss(mask_img_output.reshape(-1, D), unmask_img_output)  # st_llm.py:89-91
        return outputs, loss_mvm


class STLLMForCausalLM(nn.Module):
    """st_llm.py:94-203"""
    config_class = StllmConfig
    PRETRAINED_MODEL_CONFIG_DICT = {
        "instructblip_vicuna0": "configs/models/instructblip_vicuna0.yaml",
        "instructblip_vicuna0_btadapter": "configs/models/instructblip_vicuna0_btadapter.yaml",
        "minigpt4_vicuna0": "configs/models/minigpt4_vicuna0.yaml",
        "minigpt4_vicuna0_btadapter": "configs/models/minigpt4_vicuna0_btadapter.yaml",
    }

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.model = STLLMLlamaModel(config)
        self.vocab_size = config.vocab_size
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.lm_head.weight.data.normal_(mean=0.0, std=config.initializer_range)

    def get_model(self):
        return self.model

    @property
    def device(self):
        return self.lm_head.weight.device

    def _lm_head(self, hidden):
        w = self.lm_head.weight
        return ops.gemm(hidden, w if w.dtype == torch.float16 else w.half())

    @torch.no_grad()
    def forward(self, samples=None, inputs_embeds=None, input_ids=None, attention_mask=None, labels=None,
                past_key_values=None, use_cache=False, output_hidden_states=False, return_dict=True,
                cache_max_len=None, **kwargs):
        if samples is None:
            # == HF LlamaForCausalLM.forward (used by generate's prefill/decode steps)
            out = self.model(samples=None, inputs_embeds=inputs_embeds, input_ids=input_ids,
                             attention_mask=attention_mask, past_key_values=past_key_values, use_cache=use_cache,
                             output_hidden_states=output_hidden_states, cache_max_len=cache_max_len)
            logits = self._lm_head(out.last_hidden_state)
            loss = None
            if labels is not None:
                loss = self._ce(logits, labels)
            return CausalLMOutputWithPast(loss, logits, out.past_key_values, out.hidden_states, None)
        outputs, loss_pretrain, labels = self.model(samples)
        hidden_states = outputs[0]
        logits = self._lm_head(hidden_states)
        loss = None
        if labels is not None:
            loss = self._ce(logits, labels)
        if loss_pretrain is not None:
            loss = loss + loss_pretrain
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=outputs.past_key_values,
                                      hidden_states=outputs.hidden_states, attentions=outputs.attentions)

    def _ce(self, logits, labels):
        """st_llm.py:124-135: logits[..., :-1, :] vs labels[..., 1:], ignore_index -100 -- without copying the logits"""
        B, S, V = logits.shape
        shifted = torch.full((B, S), -100, dtype=torch.long, device=logits.device)
        shifted[:, :-1] = labels[:, 1:]
        return ops.cross_entropy(logits.view(B * S, V), shifted.view(-1))

    # ---- generation (greedy / beam / sampling) over the kernel forward ------------------------------------
    @torch.no_grad()
    def generate(self, inputs_embeds=None, input_ids=None, max_new_tokens=20, stopping_criteria=None, num_beams=1,
                 do_sample=False, min_length=0, top_p=1.0, repetition_penalty=1.0, length_penalty=1.0,
                 temperature=1.0, eos_token_id=None, pad_token_id=None, attention_mask=None, hf_semantics=None, **unused):
        """HF GenerationMixin.generate for the options the reference passes (conversation.py:231-243); hf_semantics:
        "4.28" (pinned by the reference, default) | "5" (installed release, fixture-pinned) -- see generation.py"""
        from .generation import generate as _generate
        return _generate(self, inputs_embeds=inputs_embeds, input_ids=input_ids, max_new_tokens=max_new_tokens,
                         stopping_criteria=stopping_criteria, num_beams=num_beams, do_sample=do_sample,
                         min_length=min_length, top_p=top_p, repetition_penalty=repetition_penalty,
                         length_penalty=length_penalty, temperature=temperature, eos_token_id=eos_token_id,
                         pad_token_id=pad_token_id, attention_mask=attention_mask, hf_semantics=hf_semantics)

    # ---- construction -------------------------------------------------------------------------------------
    @classmethod
    def get_state_dict(cls, path, prefix="pytorch_model"):
        """st_llm.py:149-158: merge sharded `pytorch_model-N-of-M.bin`"""
        pattern = re.compile(rf"{prefix}-(\d+)-of-(\d+).bin")
        sd = {}
        for fn in sorted(f for f in os.listdir(path) if pattern.match(f)):
            sd.update(torch.load(os.path.join(path, fn), map_location="cpu"))
        return sd

    @classmethod
    def load_hf_weights(cls, path):
        """Every layout `LlamaForCausalLM.from_pretrained` (st_llm.py:164) reads from a local directory: sharded
        `pytorch_model-N-of-M.bin` (+ index), a single `pytorch_model.bin`, `model.safetensors` or sharded
        `model-N-of-M.safetensors`.  Raises when the directory holds none of them -- a model that silently keeps its random
        initialisation is worse than an error."""
        files = sorted(os.listdir(path))
        sd = cls.get_state_dict(path)
        if not sd and "pytorch_model.bin" in files:
            sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
        st_files = [f for f in files if f.endswith(".safetensors")]
        if not sd and st_files:
            try:
                from safetensors.torch import load_file
            except ImportError as e:
                raise RuntimeError(f"{path} holds safetensors weights but the safetensors package is missing") from e
            for fn in st_files:
                sd.update(load_file(os.path.join(path, fn), device="cpu"))
        if not sd:
            raise FileNotFoundError(f"no LLM weights found in {path!r} (looked for pytorch_model-N-of-M.bin, pytorch_model.bin, "
                                    f"*.safetensors)")
        return sd

    def repack(self):
        """drop every packed (fused / norm-folded) fp16 weight copy; they are rebuilt at the next forward.  Loading a state
        dict or editing parameters in place invalidates them automatically (stllm_b200/packing.py); this is the explicit form."""
        from .packing import PackedWeightsMixin
        for m in self.modules():
            if isinstance(m, PackedWeightsMixin):
                m.repack()

    @classmethod
    def from_config(cls, cfg, **overrides):
        """st_llm.py:160-203.  `llama_model` = HF checkpoint dir (config.json + pytorch_model-*.bin) when available;
        offline, `llama_config` (dict of LlamaConfig fields) builds a random-init LLM of that architecture."""
        llama_model = cfg.get("llama_model", "")
        lcfg = dict(cfg.get("llama_config", None) or {})
        sd = None
        if llama_model and os.path.isdir(llama_model):
            import json
            with open(os.path.join(llama_model, "config.json")) as f:
                hf = json.load(f)
            for k in ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                      "max_position_embeddings", "rms_norm_eps", "pad_token_id", "bos_token_id", "eos_token_id"):
                if k in hf and hf[k] is not None:
                    lcfg.setdefault(k, hf[k])
            sd = cls.load_hf_weights(llama_model)                         # raises when the directory holds no weights
        elif llama_model and not lcfg:
            raise FileNotFoundError(f"llama_model={llama_model!r} is not a directory and no llama_config was given: pass the "
                                    f"HF checkpoint directory (st_llm.py:164) or llama_config=dict(...) for a random-init LLM")
        if cfg.get("lora_r", 0) > 0:
            raise NotImplementedError("LoRA training is outside the hot path (SURVEY section 2: peft_model.py out of scope)")
        # the reference loads the LLM in fp16 (st_llm.py:164 torch_dtype=torch.float16): build it in fp16 so that a 7B model is
        # 13.5 GB on the host, not 27, and no per-forward .half() of lm_head / embed_tokens is needed
        old_dtype = torch.get_default_dtype()
        torch.set_default_dtype(torch.float16)
        try:
            model = cls(StllmConfig(**lcfg))
        finally:
            torch.set_default_dtype(old_dtype)
        if sd is not None:
            res = model.load_state_dict(sd, strict=False)
            missing = [k for k in res.missing_keys if "stllm_model" not in k and "rotary_emb.inv_freq" not in k]
            if missing or res.unexpected_keys:
                import logging
                logging.warning("LLM weights from %s: %d missing keys %s, %d unexpected keys %s", llama_model, len(missing),
                                missing[:8], len(res.unexpected_keys), list(res.unexpected_keys)[:8])
            if not (set(sd) & set(model.state_dict())):
                raise RuntimeError(f"{llama_model!r}: none of the {len(sd)} tensors in the checkpoint matches a model parameter "
                                   f"(first keys {sorted(sd)[:3]}): wrong directory or key layout")
            model.llm_load_report = dict(missing=missing, unexpected=list(res.unexpected_keys))
        model.get_model().initialize_vision_modules(cfg, **overrides)
        if cfg.get("qformer_text_input", False) and model.get_model().stllm_model.llama_tokenizer is not None:
            n = len(model.get_model().stllm_model.llama_tokenizer)
            if n != model.lm_head.out_features:                       # resize_token_embeddings also resizes lm_head in HF
                old = model.lm_head
                new = nn.Linear(old.in_features, n, bias=False)
                new.weight.data.normal_(mean=0.0, std=model.config.initializer_range)
                new.weight.data[:min(n, old.out_features)] = old.weight.data[:min(n, old.out_features)]
                model.lm_head = new
                model.vocab_size = n
        ckpt_path = cfg.get("ckpt", "")                                   # st_llm.py:188-201
        if ckpt_path:
            if os.path.isdir(ckpt_path):
                ckpt = cls.get_state_dict(ckpt_path)                      # sharded pytorch_model-N-of-M.bin
            else:
                ckpt = torch.load(ckpt_path, map_location="cpu")          # a single MiniGPT-4 / BLIP-2 style file
            if "model" in ckpt:
                ckpt = ckpt["model"]
            for key in list(ckpt.keys()):                                 # :197-199 (the released files say llm_proj)
                if "llm_proj" in key:
                    ckpt[key.replace("llm_proj", "llama_proj")] = ckpt.pop(key)
            res = model.load_state_dict(ckpt, strict=False)
            if res.unexpected_keys:
                import logging
                logging.warning("ST-LLM checkpoint %s: %d unexpected keys %s", ckpt_path, len(res.unexpected_keys),
                                list(res.unexpected_keys)[:8])
            model.ckpt_load_report = dict(missing=list(res.missing_keys), unexpected=list(res.unexpected_keys))
        return model

    @classmethod
    def default_config_path(cls, model_type):
        assert model_type in cls.PRETRAINED_MODEL_CONFIG_DICT, "Unknown model type {}".format(model_type)
        return cls.PRETRAINED_MODEL_CONFIG_DICT[model_type]

    @classmethod
    def from_pretrained(cls, model_type, **overrides):
        """base_model.py:56-71: build from the model type's default configuration (stllm_b200/models_api.py holds the four
        YAML defaults as dicts); `overrides` are merged over its `model` section (checkpoint paths, injected tokenizers,
        `llama_config` / reduced depths for an offline random-init model)"""
        from .models_api import default_config
        cfg = default_config(cls, model_type)["model"]
        cfg.update(overrides)
        return cls.from_config(cfg)

    def load_checkpoint(self, url_or_filename):
        """base_model.py:29-52: load a finetuned checkpoint file ('model' wrapper accepted), strict=False; returns the
        missing / unexpected key report"""
        import logging
        from .models_api import _checkpoint_state_dict
        msg = self.load_state_dict(_checkpoint_state_dict(url_or_filename), strict=False)
        logging.info("Missing keys {}".format(msg.missing_keys))
        logging.info("load checkpoint from %s" % url_or_filename)
        return msg


registry.register_model("st_llm_hf")(STLLMForCausalLM)
