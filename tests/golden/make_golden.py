"""Generate the golden fixtures in tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
It imports the UNMODIFIED reference modules through oracle/ref_shim.py, builds small random-init instances of
the reference's own classes (eva_vit.VisionTransformer, Qformer.BertLMHeadModel, st_llm.STLLMForCausalLM /
STLLMModel methods, utils.RandomMaskingGenerator), runs them on seeded inputs in fp32 on CPU and stores
weights (rounded to fp16 so the files stay small; the fp16 values are what both sides load), inputs and
outputs.  tests/test_oracle_golden.py checks oracle/stllm_oracle.py against these files; the GPU tests check
the CUDA path against the oracle.  The fixtures are the pin for "parity" (the reference has no tests).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shim  # noqa: E402
from tests import fake_tokenizer  # noqa: E402


def _round_fp16_(module_or_sd):
    sd = module_or_sd.state_dict() if hasattr(module_or_sd, "state_dict") else module_or_sd
    for k, v in sd.items():
        if v.is_floating_point():
            v.copy_(v.half().float())
    return sd


def _save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path)/1e6:.2f} MB)")


def _sd_arrays(prefix, sd):
    return {f"{prefix}{k}": v.half().numpy() if v.is_floating_point() else v.numpy() for k, v in sd.items()}


@torch.no_grad()
def golden_vit(ref):
    torch.manual_seed(0)
    vit = ref_shim.build_vit(ref, img_size=56, embed_dim=176, depth=2, num_heads=2)
    # the reference initialises biases / q_bias / v_bias to zero; give them values so the fixture pins them
    for n, p in vit.named_parameters():
        if n.endswith("bias") or n.endswith("q_bias") or n.endswith("v_bias"):
            p.data.normal_(0, 0.05)
        if "norm" in n and n.endswith("weight"):
            p.data.add_(torch.randn_like(p) * 0.1)
    sd = _round_fp16_(vit)
    frames = torch.randn(3, 3, 56, 56, generator=torch.Generator().manual_seed(1234))
    out = vit(frames)
    x0 = vit.patch_embed(frames)
    _save("vit_small.npz", frames=frames.numpy(), out=out.numpy(), patch_embed=x0.numpy(), depth=2, heads=2,
          **_sd_arrays("w.", sd))


@torch.no_grad()
def golden_qformer(ref, text):
    torch.manual_seed(1)
    qf, qtok = ref_shim.build_qformer(ref, hidden=128, heads=2, inter=256, layers=3, enc_width=176, n_query=8,
                                      vocab=97, max_pos=64, text=text)
    for n, p in qf.named_parameters():
        if n.endswith("bias"):
            p.data.normal_(0, 0.05)
        elif p.dim() >= 2:
            p.data.normal_(0, 0.08)   # larger than BERT's 0.02 so attention is not near-uniform
    sd = _round_fp16_(qf.bert)
    qtok.data.copy_(qtok.data.half().float())
    g = torch.Generator().manual_seed(1236)
    img = torch.randn(5, 17, 176, generator=g)
    kw = {}
    extra = {}
    if text:
        ids = torch.randint(3, 97, (5, 6), generator=g)
        tmask = torch.ones(5, 6, dtype=torch.long)
        tmask[1, 4:] = 0
        tmask[3, 2:] = 0
        att = torch.cat([torch.ones(5, 8, dtype=torch.long), tmask], dim=1)
        out = qf.bert(ids, attention_mask=att, query_embeds=qtok.expand(5, -1, -1), encoder_hidden_states=img,
                      encoder_attention_mask=torch.ones(5, 17, dtype=torch.long), return_dict=True)
        extra = dict(text_ids=ids.numpy(), text_mask=tmask.numpy())
    else:
        out = qf.bert(query_embeds=qtok.expand(5, -1, -1), encoder_hidden_states=img,
                      encoder_attention_mask=torch.ones(5, 17, dtype=torch.long), return_dict=True)
    sd = {k: v for k, v in sd.items() if k != "embeddings.position_ids"}
    _save("qformer_text_small.npz" if text else "qformer_small.npz", image_embeds=img.numpy(),
          query_tokens=qtok.data.numpy(), out=out.last_hidden_state.numpy(), layers=3, heads=2, **extra,
          **_sd_arrays("w.", sd))


@torch.no_grad()
def golden_llama(ref):
    torch.manual_seed(2)
    st = ref.st_llm
    cfg = st.StllmConfig(hidden_size=256, num_attention_heads=2, num_key_value_heads=2, intermediate_size=512,
                         num_hidden_layers=2, vocab_size=128, max_position_embeddings=2048, rms_norm_eps=1e-6,
                         attn_implementation="eager")
    llm = st.STLLMForCausalLM(cfg).eval()
    for n, p in llm.named_parameters():
        if p.dim() >= 2:
            p.data.normal_(0, 0.06)
        else:
            p.data.add_(torch.randn_like(p) * 0.1)
    sd = _round_fp16_(llm)
    g = torch.Generator().manual_seed(1235)
    emb = torch.randn(3, 21, 256, generator=g)
    att = torch.ones(3, 21, dtype=torch.long)
    att[1, 15:] = 0   # right padding as produced by prompt_wrap / concat_emb_input_output
    att[2, 9:] = 0
    out = llm(samples=None, inputs_embeds=emb, attention_mask=att, output_hidden_states=True, return_dict=True,
              use_cache=False)
    _save("llama_small.npz", inputs_embeds=emb.numpy(), attention_mask=att.numpy(), logits=out.logits.numpy(),
          last_hidden=out.hidden_states[-1].numpy(), layer1_hidden=out.hidden_states[1].numpy(), layers=2, heads=2,
          **_sd_arrays("w.", sd))


@torch.no_grad()
def golden_btadapter(ref):
    """The reference's EVAVisionTransformer_BTAdapter (eva_btadapter.py) at a small width.  Its constructor hard-codes
    create_eva_vit_g(224, 0, False, "fp16") (a 1 B-parameter ViT-g, :52-54); for the fixture that ONE factory call is
    pointed at a small VisionTransformer of the same reference class -- every line of the BT-Adapter code (init_weights,
    forward_features, forward_branch, init_input, BTAdapter_Spatial / BTAdapter_Temp) runs unmodified."""
    import sys
    bt = sys.modules.get("stllm.models.eva_btadapter")
    assert bt is not None, "eva_btadapter is imported by blip2.py"
    torch.manual_seed(3)
    small = lambda *a, **k: ref_shim.build_vit(ref, img_size=56, embed_dim=176, depth=4, num_heads=2)
    orig = bt.create_eva_vit_g
    bt.create_eva_vit_g = small
    try:
        model = bt.EVAVisionTransformer_BTAdapter(depth=3).eval()
    finally:
        bt.create_eva_vit_g = orig
    for n, p in model.named_parameters():
        if n.endswith("bias") or n.endswith("q_bias") or n.endswith("v_bias"):
            p.data.normal_(0, 0.05)
        elif "norm" in n and n.endswith("weight"):
            p.data.add_(torch.randn_like(p) * 0.1)
        elif "temporal_fc.weight" in n:
            p.data.normal_(0, 0.05)            # zero-init in the reference (:297) would hide the temporal branch
        elif n in ("BTAdapter_cls",):
            p.data.normal_(0, 0.05)
    sd = _round_fp16_(model)
    frames = torch.randn(2, 4, 3, 56, 56, generator=torch.Generator().manual_seed(1234))   # B=2 clips, T=4 frames
    out = model(frames)
    out1 = model(frames[0])                                                                  # 4-D: one clip
    _save("btadapter_small.npz", frames=frames.numpy(), out=out.numpy(), out_clip0=out1.numpy(), depth=4, heads=2,
          bt_depth=3, **_sd_arrays("w.", sd))


def golden_preprocess():
    """Chat.transform (conversation.py:192-200) built from the reference's OWN transform classes
    (stllm/test/video_transforms.py, loaded from /root/reference) on seeded synthetic RGB frames: down-scaling
    landscape, down-scaling portrait and an up-scaling case.  Third-party arithmetic underneath: Pillow's resampler and
    torchvision's Resize / CenterCrop size rules (versions installed here: see oracle/preprocess_oracle.py)."""
    import importlib.util
    import torchvision.transforms as T
    from PIL import Image
    from torchvision.transforms import InterpolationMode
    spec = importlib.util.spec_from_file_location("ref_video_transforms",
                                                  os.path.join(ref_shim.REF, "stllm/test/video_transforms.py"))
    vt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vt)
    mean, std = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
    tr = T.Compose([vt.GroupScale(int(224), interpolation=InterpolationMode.BICUBIC), vt.GroupCenterCrop(224),
                    vt.Stack(), vt.ToTorchFormatTensor(), vt.GroupNormalize(mean, std)])
    rng = np.random.default_rng(1234)
    out = {}
    for name, (h, w) in {"land": (270, 480), "port": (300, 250), "up": (120, 200)}.items():
        yy, xx = np.mgrid[0:h, 0:w]
        base = np.stack([(np.sin(xx / 7.0) * 0.5 + 0.5) * 255, (np.cos(yy / 5.0 + xx / 11.0) * 0.5 + 0.5) * 255,
                         ((xx * 3 + yy * 5) % 256)], -1)
        frame = np.clip(base + rng.normal(0, 20, base.shape), 0, 255).astype(np.uint8)
        video = tr([Image.fromarray(frame)])
        bt, hh, ww = video.size()
        out[f"{name}_frame"] = frame
        out[f"{name}_out"] = video.view(bt // 3, 3, hh, ww).numpy()        # conversation.py:277
    _save("preprocess_frames.npz", **out)


def golden_mask(ref):
    """RandomMaskingGenerator (utils.py:4-16) + the rate draw (st_llm.py:484-485) under fixed numpy seeds"""
    out = {}
    for seed, (L, B) in {0: (8, 2), 1237: (512, 8), 7: (128, 3)}.items():
        np.random.seed(seed)
        rate = np.random.normal(0.5, 0.1)
        mask_rate = float(np.clip(rate, 0.1, 0.7))
        m = ref.utils.RandomMaskingGenerator(L, mask_rate, B, "cpu")
        out[f"seed{seed}_rate"] = np.float64(mask_rate)
        out[f"seed{seed}_mask"] = m.numpy()
        np.random.seed(seed)
        m2 = ref.utils.RandomMaskingGenerator(L, 0.5, B, "cpu")
        out[f"seed{seed}_mask_fixed_half"] = m2.numpy()
    # get_residual_index (st_llm.py:434-445) through the reference's own method
    obj = type("O", (), {"register_buffer": lambda self, n, v: None})()
    for (R, T) in [(4, 64), (16, 64), (4, 16), (2, 4), (3, 10), (5, 7)]:
        idx = ref.st_llm.STLLMModel.get_residual_index(obj, R, T, "cpu")
        out[f"residual_R{R}_T{T}"] = idx.numpy()
    _save("mask_kat.npz", **out)


def flow_weights(seed=0):
    """Deterministic weights for the end-to-end flow fixture (too large to store: 4096-wide LLM layer).
    Regenerated from the seed by the tests; only outputs are stored."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=0.02):
        return (torch.randn(*shape, generator=g) * std).half().float()
    return g, rn


@torch.no_grad()
def golden_flow(ref, text=True):
    """Whole training-branch hot path a1-a22 through the reference's own methods:
    STLLMForCausalLM.forward(samples) -> STLLMLlamaModel.forward -> STLLMModel.forward (encode_img, residual
    aggregation, dynamic mask, prompt_wrap, concat, targets) -> LLM masked+unmasked passes -> MVM + CE loss.
    text=True: the InstructBLIP layout (Q-Former text input, video first, img_start 0) -> stllm_flow.npz;
    text=False: the MiniGPT-4 layout (BOS + 7-token prefix, img_start 8, st_llm.py:71,519-530) -> stllm_flow_minigpt4.npz.
    The integer planes of STLLMModel.forward (attention masks, targets) are stored too: they pin the sequence assembly."""
    import json
    from tests.flow_fixture import build_reference_flow, flow_samples, weights_checksum
    llm, M, sd = build_reference_flow(ref, text=text)
    samples = flow_samples(text=text)
    np.random.seed(1237)
    out = llm(samples=samples)
    mask, img_len, mask_img_len = M.mask.numpy().copy(), M.img_len, M.mask_img_len
    np.random.seed(1237)
    emb, att, emb_u, att_u, targets = M(samples)          # the reference's own STLLMModel.forward, same RNG state
    assert np.array_equal(M.mask.numpy(), mask)
    _save("stllm_flow.npz" if text else "stllm_flow_minigpt4.npz", loss=out.loss.numpy(),
          logits_head=out.logits[:, :, :64].numpy(), logits_shape=np.array(out.logits.shape), mask=mask, img_len=img_len,
          mask_img_len=mask_img_len, residual_index=M.residual_index.numpy(),
          weights_checksum=np.float64(weights_checksum(sd)), attention_mask=att.numpy(),
          unmask_attention_mask=att_u.numpy(), targets=targets.numpy(), inputs_embeds_head=emb[:, :, :16].numpy(),
          unmask_inputs_embeds_head=emb_u[:, :, :16].numpy())
    if text:
        # the reference's own state-dict keys (the checkpoint contract of SURVEY section 5)
        keys = {k: list(v.shape) for k, v in llm.state_dict().items()}
        with open(os.path.join(HERE, "stllm_flow_state_dict_keys.json"), "w") as f:
            json.dump(keys, f, indent=0, sort_keys=True)


@torch.no_grad()
def golden_assembly(ref):
    """The host half of the reference's STLLMModel.forward (st_llm.py:447-546: mask draw + compaction, prompt_wrap,
    concat_emb_input_output, BOS, targets) run UNMODIFIED on a stub object whose encode_img returns seeded video tokens
    (D = 8): ragged prompts and answers, both prompt layouts, with and without the dynamic mask -> assembly_kat.npz.
    Pins stllm_b200/prompt_assembly.py bit-exactly (everything here is copies and integer logic)."""
    import torch.nn as nn
    st = ref.st_llm
    B, T, Q, D, V = 3, 2, 4, 8, 64
    g = torch.Generator().manual_seed(77)
    table = torch.randn(V, D, generator=g)
    video = torch.randn(B, T, Q, D, generator=g)
    answers = ["yes", "a man is slicing bread on a table in the kitchen", "two dogs"]
    cases = {
        "text": dict(text=True, end_sym="\n", prompts=[
            "<ImageHere>###Human: what happens ? ###Assistant: ",
            "<ImageHere>###Human: describe the video in detail please ###Assistant: ",
            "<ImageHere>###Human: why ###Assistant: "]),
        "minigpt4": dict(text=False, end_sym=" ###", prompts=[
            "### Human : < Video > clip <ImageHere> </Video> what happens ? ###Assistant: ",
            "### Human : < Video > clip <ImageHere> </Video> describe the video in detail please ###Assistant: ",
            "###Human: <Video> <ImageHere> </Video> why ###Assistant: "]),        # a shorter prefix too (ragged before)
        "same_prompt": dict(text=True, end_sym="\n", prompts="<ImageHere>###Human: what is shown ? ###Assistant: "),
    }
    out = dict(table=table.numpy(), video=video.numpy(), answers=np.array(answers))
    for name, c in cases.items():
        for use_mask in (False, True):
            M = st.STLLMModel.__new__(st.STLLMModel)
            nn.Module.__init__(M)
            M.pre_encoding, M.video_input, M.use_mask, M.mvm_decode = False, "all", use_mask, False
            M.qformer_text_input, M.max_txt_len, M.end_sym = c["text"], 8, c["end_sym"]
            M.llama_tokenizer = fake_tokenizer.FakeTokenizer(V, pad_token_id=0, bos_token_id=1, eos_token_id=2)
            M.embed_tokens = nn.Embedding(V, D)
            M.embed_tokens.weight.data.copy_(table)
            M.encode_img = lambda image, text: (video.clone(), torch.ones(B, T, Q, dtype=torch.long), False)
            prompts = c["prompts"] if isinstance(c["prompts"], list) else [c["prompts"]] * B
            samples = {"image": torch.zeros(B, T, 3, 2, 2), "instruction_input": prompts, "answer": answers}
            np.random.seed(4321)
            emb, att, emb_u, att_u, targets = st.STLLMModel.forward(M, samples)
            key = f"{name}.mask{int(use_mask)}."
            out[key + "prompts"] = np.array(prompts)
            out[key + "inputs_embeds"] = emb.numpy()
            out[key + "attention_mask"] = att.numpy()
            out[key + "targets"] = targets.numpy()
            if use_mask:
                out[key + "mask"] = M.mask.numpy()
                out[key + "unmask_inputs_embeds"] = emb_u.numpy()
                out[key + "unmask_attention_mask"] = att_u.numpy()
    _save("assembly_kat.npz", **out)


@torch.no_grad()
def golden_noqformer(ref):
    """encode_img with has_qformer=False (st_llm.py:300-301,369-373: drop cls, fold 4 neighbouring patch tokens into one,
    llama_proj 4*C -> D) through the reference's own unbound STLLMModel.encode_img on a hand-assembled object, for a
    5-D clip batch and a 4-D frame stack."""
    import contextlib
    import torch.nn as nn
    st = ref.st_llm
    torch.manual_seed(0)
    M = st.STLLMModel.__new__(st.STLLMModel)
    nn.Module.__init__(M)
    M.vit_model, M.has_qformer, M.qformer_text_input = "eva_clip_g", False, False
    M.visual_encoder = ref_shim.build_vit(ref, img_size=56, embed_dim=176, depth=2, num_heads=2)
    M.ln_vision = ref.blip2.LayerNorm(176)
    M.llama_proj = nn.Linear(176 * 4, 64)                       # img_f_dim = num_features * 4 (st_llm.py:301)
    M.maybe_autocast = lambda dtype=None: contextlib.nullcontext()
    for n, p in M.named_parameters():
        if n.endswith("bias") or n.endswith("q_bias") or n.endswith("v_bias"):
            p.data.normal_(0, 0.05)
        if ("norm" in n or n.startswith("ln_vision")) and n.endswith("weight"):
            p.data.add_(torch.randn_like(p) * 0.1)
    sd = _round_fp16_(M)
    M = M.eval()
    g = torch.Generator().manual_seed(1234)
    clips = torch.randn(2, 3, 3, 56, 56, generator=g)
    out5, atts5, use5 = st.STLLMModel.encode_img(M, clips)
    out4, atts4, use4 = st.STLLMModel.encode_img(M, clips[0])
    _save("noqformer_small.npz", clips=clips.numpy(), out5=out5.numpy(), out4=out4.numpy(), atts5_shape=np.array(atts5.shape),
          atts4_shape=np.array(atts4.shape), use_image=np.array([int(use5), int(use4)]), depth=2, heads=2,
          **_sd_arrays("w.", sd))


@torch.no_grad()
def golden_pre_encoding(ref):
    """`pre_encoding=True` (st_llm.py:452-455): samples["image"] already holds Q-Former features [B,T,32,768]; only
    llama_proj runs, then 'mean' aggregation and the sequence assembly -- the reference's own STLLMModel.forward on a stub
    object (MiniGPT-4 layout: BOS + prefix) -> pre_encoding_small.npz"""
    import torch.nn as nn
    st = ref.st_llm
    B, T, Q, C, D, V = 2, 3, 32, 768, 128, 80
    g = torch.Generator().manual_seed(31)
    M = st.STLLMModel.__new__(st.STLLMModel)
    nn.Module.__init__(M)
    M.pre_encoding, M.video_input, M.use_mask, M.mvm_decode = True, "mean", False, False
    M.qformer_text_input, M.max_txt_len, M.end_sym = False, 8, " ###"
    M.llama_tokenizer = fake_tokenizer.FakeTokenizer(V, pad_token_id=0, bos_token_id=1, eos_token_id=2)
    M.embed_tokens = nn.Embedding(V, D)
    M.llama_proj = nn.Linear(C, D)
    M.embed_tokens.weight.data.copy_((torch.randn(V, D, generator=g) * 0.5).half().float())
    M.llama_proj.weight.data.copy_((torch.randn(D, C, generator=g) * 0.05).half().float())
    M.llama_proj.bias.data.copy_((torch.randn(D, generator=g) * 0.1).half().float())
    feats = (torch.randn(B, T, Q, C, generator=g)).half().float()
    samples = {"image": feats,
               "instruction_input": ["### Human : < Video > clip <ImageHere> </Video> what happens ? ###Assistant: ",
                                     "### Human : < Video > clip <ImageHere> </Video> describe it ###Assistant: "],
               "answer": ["a cat sleeps", "two people talk on a bench"]}
    emb, att, emb_u, att_u, targets = st.STLLMModel.forward(M, samples)
    assert emb_u is None
    _save("pre_encoding_small.npz", feats=feats.numpy(), inputs_embeds=emb.numpy(), attention_mask=att.numpy(),
          targets=targets.numpy(), prompts=np.array(samples["instruction_input"]), answers=np.array(samples["answer"]),
          **{"w.embed_tokens.weight": M.embed_tokens.weight.numpy(), "w.llama_proj.weight": M.llama_proj.weight.numpy(),
             "w.llama_proj.bias": M.llama_proj.bias.numpy()})


GEN_CASES = {
    # name: generate kwargs (the harness settings of SURVEY 3A / conversation.py:231-243: do_sample False, min_length 1, ...).
    # min_length counts the prompt when generate() gets inputs_embeds only (installed transformers: `_prepare_generated_length`
    # subtracts the prompt length), so values below prompt_len = 19 are no-ops -- like the reference's min_length = 1.
    "greedy": dict(num_beams=1, max_new_tokens=24, min_length=1, repetition_penalty=1.0, length_penalty=1.0),
    "greedy_rp15_min6": dict(num_beams=1, max_new_tokens=24, min_length=19 + 6, repetition_penalty=1.5, length_penalty=1.0),
    "beam5": dict(num_beams=5, max_new_tokens=24, min_length=1, repetition_penalty=1.0, length_penalty=1.0),
    "beam5_rp15": dict(num_beams=5, max_new_tokens=24, min_length=1, repetition_penalty=1.5, length_penalty=1.0),
    "beam5_min8_lp2": dict(num_beams=5, max_new_tokens=20, min_length=19 + 8, repetition_penalty=1.0, length_penalty=2.0),
    "beam3_lp05_short": dict(num_beams=3, max_new_tokens=6, min_length=1, repetition_penalty=1.2, length_penalty=0.5),
    "beam5_noeos_maxlen": dict(num_beams=5, max_new_tokens=10, min_length=19 + 11, repetition_penalty=1.0, length_penalty=1.0),
    "beam4_rp13_lp15": dict(num_beams=4, max_new_tokens=16, min_length=19 + 3, repetition_penalty=1.3, length_penalty=1.5),
}
GEN_MODEL = dict(hidden=256, heads=2, layers=2, inter=512, vocab=96, prompt_len=19)


def build_reference_generate_model(ref, seed=11):
    """small reference STLLMForCausalLM (head dim 128: the decode engine's shape) with a peaky next-token distribution"""
    st = ref.st_llm
    torch.manual_seed(seed)
    m = GEN_MODEL
    cfg = st.StllmConfig(hidden_size=m["hidden"], num_attention_heads=m["heads"], num_key_value_heads=m["heads"],
                         intermediate_size=m["inter"], num_hidden_layers=m["layers"], vocab_size=m["vocab"],
                         max_position_embeddings=256, rms_norm_eps=1e-6, attn_implementation="eager", pad_token_id=0,
                         bos_token_id=1, eos_token_id=2)
    llm = st.STLLMForCausalLM(cfg).eval()
    g = torch.Generator().manual_seed(seed)
    for n, p in llm.named_parameters():
        if p.dim() >= 2:
            std = 0.05
            if "embed_tokens" in n:
                std = 1.0
            if "lm_head" in n:
                std = 0.35                      # peaky distributions: decisions are robust to fp16 rounding
            p.data.copy_(torch.randn(p.shape, generator=g) * std)
        else:
            p.data.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
    llm.lm_head.weight.data[2] *= 2.5           # EOS (id 2) wins often enough that finished hypotheses compete
    _round_fp16_(llm)
    return llm


@torch.no_grad()
def golden_generate(ref):
    """Token ids of the reference's OWN `STLLMForCausalLM.generate(inputs_embeds=...)` (inherited HF GenerationMixin; the
    transformers release installed here executes it -- recorded in the file) for greedy and beam search with the harness
    settings; 3 prompts per case.  Also the per-step decision margins, so that a fixture whose decisions hinge on fp16
    rounding is rejected when it is made."""
    import transformers
    llm = build_reference_generate_model(ref)
    m = GEN_MODEL
    g = torch.Generator().manual_seed(5)
    prompts = (torch.randn(3, m["prompt_len"], m["hidden"], generator=g)).half().float()
    out = dict(prompts=prompts.numpy(), transformers_version=np.array(transformers.__version__),
               **_sd_arrays("w.", llm.state_dict()))
    for name, kw in GEN_CASES.items():
        for pi in range(prompts.shape[0]):
            ids = llm.generate(inputs_embeds=prompts[pi:pi + 1], do_sample=False, top_p=0.9, temperature=1.0, **kw)
            assert ids.shape[0] == 1
            out[f"{name}.p{pi}"] = ids[0].numpy()
            print(name, pi, ids[0].tolist())
    _save("generate_small.npz", **out)


if __name__ == "__main__":
    ref = ref_shim.load_reference()
    which = sys.argv[1:] or ["vit", "qformer", "qformer_text", "llama", "mask", "flow", "flow_minigpt4", "assembly",
                             "btadapter", "preprocess", "noqformer", "generate", "pre_encoding"]
    if "vit" in which:
        golden_vit(ref)
    if "qformer" in which:
        golden_qformer(ref, text=False)
    if "qformer_text" in which:
        golden_qformer(ref, text=True)
    if "llama" in which:
        golden_llama(ref)
    if "mask" in which:
        golden_mask(ref)
    if "flow" in which:
        golden_flow(ref)
    if "flow_minigpt4" in which:
        golden_flow(ref, text=False)
    if "assembly" in which:
        golden_assembly(ref)
    if "btadapter" in which:
        golden_btadapter(ref)
    if "preprocess" in which:
        golden_preprocess()
    if "noqformer" in which:
        golden_noqformer(ref)
    if "generate" in which:
        golden_generate(ref)
    if "pre_encoding" in which:
        golden_pre_encoding(ref)
