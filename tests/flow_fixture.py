"""Shared set-up of the END-TO-END flow fixture (training-branch hot path a1-a22 of SURVEY section 8):
B=2 clips x T=4 frames -> ViT-g width (2 blocks) -> Q-Former (2 layers, WITH text input) -> llama_proj ->
global-local 'residual' aggregation (R=2) -> dynamic mask -> prompt_wrap/concat/targets -> masked + unmasked LLM
passes (1 layer, hidden 4096 as st_llm.py hard-codes) -> MVM loss + CE loss.

The weights (~170 M parameters) are too large to commit; they are regenerated from a seed with a fixed
torch.Generator recipe (`flow_state_dict`) by the golden generator, the oracle test and the GPU test alike; the
fixture stores a checksum of them so RNG drift is detected instead of silently mis-compared.
"""
import numpy as np
import torch
import torch.nn as nn

from tests.fake_tokenizer import FakeTokenizer

FLOW = dict(B=2, T=4, vit_depth=2, qf_layers=2, llama_layers=1, llama_inter=256, vocab=128, bert_vocab=97,
            residual_size=2, max_txt_len=16, n_query=32)


def flow_tokenizers():
    return FakeTokenizer(FLOW["bert_vocab"]), FakeTokenizer(FLOW["vocab"], pad_token_id=0, bos_token_id=1, eos_token_id=2)


# MiniGPT-4 layout (qformer_text_input=False): BOS + the text before <ImageHere> + video.  STLLMLlamaModel.forward
# hard-codes img_start = 8 for it (st_llm.py:71: BOS + the 7 Llama tokens of "###Human: <Video>"); the whitespace FakeTokenizer
# needs 7 words to put the video at position 8 as well.
MINIGPT4_PREFIX = "### Human : < Video > clip "


def flow_samples(text=True):
    g = torch.Generator().manual_seed(1234)
    image = torch.randn(FLOW["B"], FLOW["T"], 3, 224, 224, generator=g)
    if text:
        instr = ["<ImageHere>###Human: what is the person doing in this video ? ###Assistant: ",
                 "<ImageHere>###Human: describe the scene ###Assistant: "]
    else:
        instr = [MINIGPT4_PREFIX + "<ImageHere> </Video> what is the person doing in this video ? ###Assistant: ",
                 MINIGPT4_PREFIX + "<ImageHere> </Video> describe the scene ###Assistant: "]
    return {
        "image": image,
        "instruction_input": instr,
        "answer": ["the person is cooking", "a dog runs on the beach happily"],
    }


def flow_state_dict(shapes, seed=0):
    """shapes: ordered {name: shape} of floating-point tensors (reference key names) -> seeded fp16-exact values"""
    import zlib
    sd = {}
    for name in sorted(shapes):
        shape = shapes[name]
        g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(name.encode()))   # value depends on the name only
        low = name.lower()
        if len(shape) >= 2:
            std = 0.02
            if "embed_tokens" in low or "word_embeddings" in low or "position_embeddings" in low:
                std = 0.5          # token embeddings on the scale of the projected video tokens
            if "llama_proj.weight" in low:
                std = 0.05
            t = torch.randn(shape, generator=g) * std
        elif low.endswith("weight") and ("norm" in low):
            t = 1.0 + torch.randn(shape, generator=g) * 0.1
        else:
            t = torch.randn(shape, generator=g) * 0.02
        sd[name] = t.half().float()
    return sd


def weights_checksum(sd):
    return float(sum(v.double().abs().sum().item() for v in sd.values()))


def float_shapes(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items() if v.is_floating_point()}


# ------------------------------------------------------------------------------------------ reference side
def build_reference_flow(ref, text=True):
    """hand-assembled reference STLLMForCausalLM + STLLMModel, weights from flow_state_dict.
    text=False: the MiniGPT-4 variant (no Q-Former text input, BOS + 7-token prefix, end_sym '###')"""
    from oracle import ref_shim
    st = ref.st_llm
    torch.manual_seed(0)
    cfg = st.StllmConfig(hidden_size=4096, num_attention_heads=32, num_key_value_heads=32,
                         intermediate_size=FLOW["llama_inter"], num_hidden_layers=FLOW["llama_layers"],
                         vocab_size=FLOW["vocab"], max_position_embeddings=2048, rms_norm_eps=1e-6,
                         attn_implementation="eager", pad_token_id=0, bos_token_id=1, eos_token_id=2)
    llm = st.STLLMForCausalLM(cfg).eval()
    M = st.STLLMModel.__new__(st.STLLMModel)
    nn.Module.__init__(M)
    M.pre_encoding, M.video_input, M.use_mask, M.mvm_decode = False, "residual", True, True
    M.qformer_text_input, M.residual_size, M.vit_model, M.has_qformer = text, FLOW["residual_size"], "eva_clip_g", True
    M.max_txt_len, M.end_sym = FLOW["max_txt_len"], "\n" if text else " ###"
    M.visual_encoder = ref_shim.build_vit(ref, 224, 1408, FLOW["vit_depth"], 16)
    M.ln_vision = ref.blip2.LayerNorm(1408)
    M.Qformer, M.query_tokens = ref_shim.build_qformer(ref, 768, 12, 3072, FLOW["qf_layers"], 1408, FLOW["n_query"],
                                                       vocab=FLOW["bert_vocab"], max_pos=512, text=text)
    M.llama_proj = nn.Linear(768, 4096)
    M.down_proj, M.non_linear_func, M.up_proj = nn.Linear(4096, 1024), nn.ReLU(), nn.Linear(1024, 4096)
    M.mvm_decoder = st.Linear_Decoder()
    M.tokenizer, M.llama_tokenizer = flow_tokenizers()
    llm.model.stllm_model = M
    M.embed_tokens = llm.model.embed_tokens
    shapes = {k: v for k, v in float_shapes(llm).items() if ".stllm_model.embed_tokens." not in k}
    sd = flow_state_dict(shapes)
    res = llm.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(("position_ids" in k) or (".stllm_model.embed_tokens." in k) or ("inv_freq" in k)
               for k in res.missing_keys), res.missing_keys
    return llm.eval(), M, sd


# ------------------------------------------------------------------------------------------ oracle side
def oracle_flow(sd, samples, text=True):
    """the same flow composed from oracle/stllm_oracle.py functions; consumes np.random like the reference"""
    from oracle import stllm_oracle as O
    w = O.split_state_dict(sd)
    btok, ltok = flow_tokenizers()
    instr, answers, image = samples["instruction_input"], samples["answer"], samples["image"]
    B = image.shape[0]
    cfg = dict(vit_depth=FLOW["vit_depth"], vit_heads=16, qf_layers=FLOW["qf_layers"])
    if text:
        qtext = [it.split("Human: ")[1].split(" ###")[0] for it in instr]                  # st_llm.py:458
        enc = btok(qtext, padding="longest", truncation=True, max_length=FLOW["max_txt_len"], return_tensors="pt")
        emb = O.encode_img(w, image, cfg, text_ids=enc.input_ids, text_mask=enc.attention_mask)  # [B,T,Q,D]
    else:
        emb = O.encode_img(w, image, cfg)
    vid = O.video_aggregate(emb, "residual", FLOW["residual_size"], w["residual"])           # [B,1,L,D]
    L = vid.shape[2]
    rate = O.draw_mask_rate()
    mask = O.random_masking_generator(L, rate, B).unsqueeze(1)                               # [B,1,L]
    D = vid.shape[-1]
    masked_vid = vid[~mask].reshape(B, 1, -1, D)
    E = w["llama"]["embed_tokens.weight"]
    prompts = [E[ltok(it.split("<ImageHere>")[1], return_tensors="pt", add_special_tokens=text).input_ids[0]]
               for it in instr]                                                             # st_llm.py:388-389
    ltok.padding_side = "right"
    ans = ltok([t + (ltok.eos_token if text else " ###") for t in answers], return_tensors="pt", padding="longest",
               truncation=True, max_length=FLOW["max_txt_len"], add_special_tokens=False)   # st_llm.py:498-508
    pad_embed = E[ltok.pad_token_id]
    args = (prompts, E[ans.input_ids], ans.attention_mask, pad_embed, ans.input_ids, ltok.pad_token_id)
    kw = {}
    if not text:
        kw = dict(before_embeds=[E[ltok(it.split("<ImageHere>")[0], return_tensors="pt",
                                        add_special_tokens=False).input_ids[0]] for it in instr],
                  bos_embed=E[ltok.bos_token_id])
    emb_m, att_m, targets = O.assemble_sequence(masked_vid, *args, **kw)
    emb_u, att_u, _ = O.assemble_sequence(vid, *args, **kw)
    hid_m = O.llama_forward(w["llama"], emb_m, att_m, FLOW["llama_layers"], 32)
    hid_u = O.llama_forward(w["llama"], emb_u, att_u, FLOW["llama_layers"], 32)
    img_start = 0 if text else 8                                                            # st_llm.py:71
    loss_mvm = O.mvm_loss(w["mvm_decoder"], hid_m, hid_u, mask, img_start, L, masked_vid.shape[2])
    logits = torch.nn.functional.linear(hid_m, w["lm_head"])
    loss = O.causal_lm_loss(logits, targets, logits.shape[-1]) + loss_mvm
    return dict(loss=loss, logits=logits, mask=mask, img_len=L, mask_img_len=masked_vid.shape[2], targets=targets,
                attention_mask=att_m, loss_mvm=loss_mvm, inputs_embeds=emb_m, unmask_inputs_embeds=emb_u,
                unmask_attention_mask=att_u)


# ------------------------------------------------------------------------------------------ product side
def build_product_flow(device, text=True):
    from stllm_b200.pipeline import build_model
    btok, ltok = flow_tokenizers()
    model = build_model(llama_layers=FLOW["llama_layers"], vit_depth=FLOW["vit_depth"], qformer_layers=FLOW["qf_layers"],
                        vocab_size=FLOW["vocab"], llama_inter=FLOW["llama_inter"], video_input="residual",
                        residual_size=FLOW["residual_size"], use_mask=True, mvm_decode=True, qformer_text_input=text,
                        tokenizer=btok, llama_tokenizer=ltok, device=device)
    model.get_model().stllm_model.max_txt_len = FLOW["max_txt_len"]
    if not text:
        model.get_model().stllm_model.end_sym = " ###"
    sd = flow_state_dict(float_shapes(model))
    model.load_state_dict(sd, strict=False)
    return model, sd
