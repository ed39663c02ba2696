"""CPU tests of the host side: C-ABI export contract, state-dict/attribute contract with the reference, registry,
bit-exact integer paths (mask, residual index), error behaviour without a GPU."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from tests.util import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    from stllm_b200 import _lib
    header = open(os.path.join(ROOT, "include", "stllm_b200.h")).read()
    declared = set(re.findall(r"\b(stllm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/stllm_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES) | {"stllm_last_error"}
    assert _lib.load().stllm_abi_version() == 1


def test_no_cpu_fallback():
    from stllm_b200 import _lib, ops
    a = torch.zeros(8, 64, dtype=torch.float16)
    with pytest.raises(_lib.StllmNativeError):
        ops.gemm(a, a)
    with pytest.raises(_lib.StllmNativeError):
        ops.layernorm(a, torch.ones(64), torch.zeros(64), 1e-5)


def test_bad_arguments_return_error_codes_not_crashes():
    from stllm_b200 import _lib
    lib = _lib.load()
    rc = lib.stllm_gemm_f16(None, 64, None, 64, None, 64, 8, 8, 64, None, None, 0, 0, 0, 0, None)
    assert rc == -1 and b"bad argument" in lib.stllm_last_error()
    rc = lib.stllm_layernorm_f16(None, 8, None, 8, None, None, 1, 8, 1e-5, None)
    assert rc == -1


def test_decode_entry_points_validate_before_touching_the_device():
    # argument validation of the decode-step C ABI happens before any CUDA call: exercised here without a GPU, with fake
    # (never dereferenced) 16-byte-aligned addresses
    import ctypes
    from stllm_b200 import _lib, ops
    lib = _lib.load()
    P = ctypes.c_void_p(0x1000)
    names = ["x", "ldx", "W", "ldw", "y", "ldy", "R", "N", "K", "norm_mode", "norm_w", "eps", "res", "ldr", "epi", "st"]
    ok_args = [P, 4096, P, 4096, P, 4096, 5, 4096, 4096, 0, None, 1e-6, None, 0, 0, None]

    def lin(**kw):
        a = list(ok_args)
        for k, v in kw.items():
            a[names.index(k)] = v
        return lib.stllm_decode_linear_f16(*a)
    assert lin(x=None) == -1 and b"bad argument" in lib.stllm_last_error()
    assert lin(R=9) == -1 and lin(R=0) == -1                       # at most 8 rows (beams) per step
    assert lin(norm_mode=2) == -1                                   # explicit RMSNorm needs the norm weight
    assert lin(epi=1) == -1 and lin(epi=2, N=4095) == -1            # only NONE / SWIGLU; SwiGLU pairs rows
    assert lin(ldx=4090) == -1 and lin(x=ctypes.c_void_p(0x1008)) == -1
    assert lin(K=4104, ldx=4104, ldw=4104) == -3 and b"multiple of 32" in lib.stllm_last_error()   # unsupported, not bad
    assert lin(R=8, K=16384, ldx=16384, ldw=16384) == -3            # 8 x 16384 activations exceed shared memory
    with pytest.raises(_lib.StllmNativeError):
        ops.decode_linear(torch.zeros(1, 64, dtype=torch.float16), torch.zeros(16, 64, dtype=torch.float16),
                          torch.zeros(1, 16, dtype=torch.float16))  # CPU tensors: no fallback
    assert lib.stllm_decode_attention_workspace_bytes(32, 4) == 32 * 4 * 8 * 132 * 4
    assert lib.stllm_decode_attention_fused_workspace_bytes(32, 4) == 32 * 4 * 8 * 132 * 4 + 128
    I = ctypes.c_void_p(0x2000)
    # q, ldq, kp, vp, p_stride, mask, Sp, ks, vs, beam stride, token stride, l_max, src, src_ld, n_gen, out, ldo, ws,
    # ws bytes, R, H, D, splits, scale, stream
    att = [P, 12288, P, P, 4096, None, 576, P, P, 4096 * 64, 4096, 64, I, 64, I, P, 4096, P, 1 << 30, 5, 32, 128, 4,
           0.088, None]
    for pos, val, want in ((21, 64, -3),      # head dim 64: unsupported
                           (19, 9, -1),       # 9 beams
                           (18, 1024, -1),    # workspace too small
                           (13, 32, -1),      # ancestry table narrower than the suffix cache
                           (0, None, -1)):    # no q
        bad = list(att)
        bad[pos] = val
        assert lib.stllm_decode_attention_f16(*bad) == want, (pos, val)


def test_encode_img_without_qformer_host_logic(monkeypatch):
    # SURVEY a15' (st_llm.py:300-301,369-373): the product's has_qformer=False branch -- construction (no Q-Former, llama_proj
    # takes 4 C features) and the token folding of encode_img, checked on the CPU against the reference's own output with the
    # encoder / LayerNorm / GEMM calls replaced by the oracle's fp32 functions (the kernels themselves are covered elsewhere)
    import torch.nn as nn
    from oracle import stllm_oracle as O
    from stllm_b200 import st_llm as S
    from tests.util import assert_close, load_golden
    sm = S.STLLMModel(has_qformer=False, vit_depth=1, llama_hidden=64)
    assert not hasattr(sm, "Qformer") and not hasattr(sm, "query_tokens")
    assert sm.llama_proj.in_features == 4 * sm.visual_encoder.num_features == 5632 and sm.llama_proj.out_features == 64
    assert not any(k.startswith(("Qformer", "query_tokens")) for k in sm.state_dict())
    sd, g = load_golden("noqformer_small.npz")
    vit_sd = {k[len("visual_encoder."):]: v for k, v in sd.items() if k.startswith("visual_encoder.")}
    ln_sd = {k[len("ln_vision."):]: v for k, v in sd.items() if k.startswith("ln_vision.")}

    class Enc(nn.Module):
        num_features = 176

        def forward(self, img):
            return O.vit_forward(vit_sd, img, int(g["depth"]), int(g["heads"]))

    class LN(nn.Module):
        def forward(self, x):
            return O.ln_vision(ln_sd, x)
    sm.visual_encoder, sm.ln_vision = Enc(), LN()
    sm.llama_proj = nn.Linear(176 * 4, 64)
    sm.llama_proj.load_state_dict({"weight": sd["llama_proj.weight"], "bias": sd["llama_proj.bias"]})
    monkeypatch.setattr(S, "_linear", lambda mod, x, epilogue=0: torch.nn.functional.linear(x, mod.weight, mod.bias))
    clips = torch.from_numpy(g["clips"])
    out5, atts5, use5 = sm.encode_img(clips)
    out4, atts4, use4 = sm.encode_img(clips[0])
    assert tuple(atts5.shape) == tuple(g["atts5_shape"]) and tuple(atts4.shape) == tuple(g["atts4_shape"])
    assert [int(use5), int(use4)] == g["use_image"].tolist() and atts5.dtype == torch.long
    assert_close(out5, g["out5"], atol=2e-5, rtol=1e-4, what="product encode_img has_qformer=False (5-D)")
    assert_close(out4, g["out4"], atol=2e-5, rtol=1e-4, what="product encode_img has_qformer=False (4-D)")


def test_mask_and_residual_index_bit_exact_with_reference():
    from stllm_b200.st_llm import RandomMaskingGenerator, STLLMModel
    g = np.load(os.path.join(GOLDEN, "mask_kat.npz"))
    for seed, (L, B) in {0: (8, 2), 1237: (512, 8), 7: (128, 3)}.items():
        np.random.seed(seed)
        rate = float(np.clip(np.random.normal(0.5, 0.1), 0.1, 0.7))
        m = RandomMaskingGenerator(L, rate, B, "cpu")
        assert m.dtype == torch.bool and np.array_equal(m.numpy(), g[f"seed{seed}_mask"])
    obj = type("O", (), {"register_buffer": lambda self, n, v: setattr(self, n, v)})()
    for (R, T) in [(4, 64), (16, 64), (4, 16), (2, 4), (3, 10), (5, 7)]:
        o = type("O", (), {"register_buffer": lambda self, n, v: None})()
        idx = STLLMModel.get_residual_index(o, R, T, "cpu")
        assert np.array_equal(idx.numpy(), g[f"residual_R{R}_T{T}"])


def _small_model():
    from stllm_b200.qformer import BertConfig
    from stllm_b200.st_llm import STLLMForCausalLM
    from tests.flow_fixture import FLOW, flow_tokenizers
    btok, ltok = flow_tokenizers()
    cfg = dict(llama_config=dict(vocab_size=FLOW["vocab"], hidden_size=4096, intermediate_size=FLOW["llama_inter"],
                                 num_hidden_layers=1, num_attention_heads=32),
               video_input="residual", residual_size=2, vit_depth=2, use_mask=True, mvm_decode=True,
               qformer_text_input=True, qformer_config=BertConfig(num_hidden_layers=2, vocab_size=FLOW["bert_vocab"]),
               tokenizer=btok, llama_tokenizer=ltok)
    return STLLMForCausalLM.from_config(cfg)


def test_state_dict_and_attribute_contract_with_reference():
    """every parameter key/shape of the reference's STLLMForCausalLM (fixture written from the reference's own
    state_dict()) exists in ours; and the attribute graph the callers walk exists."""
    from stllm_b200.registry import registry
    model = _small_model()
    with open(os.path.join(GOLDEN, "stllm_flow_state_dict_keys.json")) as f:
        ref_keys = json.load(f)
    ours = {k: list(v.shape) for k, v in model.state_dict().items()}
    for k, shape in ref_keys.items():
        if ".stllm_model.embed_tokens." in k:
            continue  # alias of model.embed_tokens (same tensor) in the reference
        if k.endswith("residual_index"):
            continue  # buffer registered lazily by get_residual_index on the first forward (st_llm.py:444), here too
        assert k in ours, f"reference key {k} missing"
        assert ours[k] == shape, (k, ours[k], shape)
    sm = model.model.stllm_model
    for attr in ("encode_img", "video_input", "residual_size", "up_proj", "down_proj", "non_linear_func",
                 "embed_tokens", "llama_tokenizer", "qformer_text_input", "mvm_decoder", "prompt_wrap",
                 "concat_emb_input_output", "get_residual_index", "Qformer", "query_tokens", "ln_vision",
                 "visual_encoder", "llama_proj"):
        assert hasattr(sm, attr), attr
    assert sm.embed_tokens is model.model.embed_tokens
    assert model.get_model() is model.model
    assert registry.get_model_class("st_llm_hf") is type(model)
    assert callable(model.generate)


def test_encode_img_asserts_like_reference():
    model = _small_model()
    sm = model.model.stllm_model
    with pytest.raises(AssertionError):   # eva_vit.py:201-202 image-size assert (raised before any kernel launch)
        sm.visual_encoder(torch.zeros(1, 3, 64, 64))


def test_from_config_loads_hf_shards_and_stllm_checkpoint(tmp_path):
    """st_llm.py:149-203: `llama_model` = HF directory (config.json + pytorch_model-N-of-M.bin shards), `ckpt` = a
    directory of shards or a single file, optionally wrapped in {'model': ...}, with the llm_proj -> llama_proj rename"""
    import json as _json
    from stllm_b200.qformer import BertConfig
    from stllm_b200.st_llm import STLLMForCausalLM
    from tests.flow_fixture import FLOW, flow_tokenizers
    btok, ltok = flow_tokenizers()
    hf_dir = tmp_path / "vicuna"
    hf_dir.mkdir()
    hf_cfg = dict(vocab_size=FLOW["vocab"], hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                  num_attention_heads=2, max_position_embeddings=2048, rms_norm_eps=1e-6)
    (hf_dir / "config.json").write_text(_json.dumps(hf_cfg))
    g = torch.Generator().manual_seed(5)
    shard1 = {"model.layers.0.self_attn.q_proj.weight": torch.randn(256, 256, generator=g),
              "model.embed_tokens.weight": torch.randn(FLOW["vocab"], 256, generator=g)}
    shard2 = {"model.layers.1.mlp.down_proj.weight": torch.randn(256, 512, generator=g),
              "lm_head.weight": torch.randn(FLOW["vocab"], 256, generator=g)}
    torch.save(shard1, hf_dir / "pytorch_model-00001-of-00002.bin")
    torch.save(shard2, hf_dir / "pytorch_model-00002-of-00002.bin")
    ck = {"model": {"model.stllm_model.llm_proj.weight": torch.randn(256, 768, generator=g),
                    "model.stllm_model.llm_proj.bias": torch.randn(256, generator=g),
                    "model.stllm_model.query_tokens": torch.randn(1, 32, 768, generator=g)}}
    ck_file = tmp_path / "stllm_ckpt.pth"
    torch.save(ck, ck_file)
    cfg = dict(llama_model=str(hf_dir), ckpt=str(ck_file), vit_depth=1, video_input="all", llama_hidden=256,
               qformer_config=BertConfig(num_hidden_layers=2, vocab_size=FLOW["bert_vocab"]), tokenizer=btok,
               llama_tokenizer=ltok)
    model = STLLMForCausalLM.from_config(cfg)
    assert model.config.hidden_size == 256 and len(model.model.layers) == 2
    sd = model.state_dict()
    for k, v in {**shard1, **shard2}.items():            # the LLM is built in fp16, like the reference (st_llm.py:164)
        assert sd[k].dtype == torch.float16 and torch.equal(sd[k].float(), v.half().float()), k
    assert model.llm_load_report["unexpected"] == [] and len(model.llm_load_report["missing"]) > 0
    assert torch.equal(sd["model.stllm_model.llama_proj.weight"].float(), ck["model"]["model.stllm_model.llm_proj.weight"])
    assert torch.equal(sd["model.stllm_model.query_tokens"].float(), ck["model"]["model.stllm_model.query_tokens"])
    # the same checkpoint as a directory of shards (the layout the released ST-LLM weights use)
    ck_dir = tmp_path / "stllm_ckpt_dir"
    ck_dir.mkdir()
    torch.save(ck["model"], ck_dir / "pytorch_model-00001-of-00001.bin")
    # a single pytorch_model.bin and safetensors are accepted too; an empty directory is an error, not a random-init model
    single = tmp_path / "vicuna_single"
    single.mkdir()
    (single / "config.json").write_text(_json.dumps(hf_cfg))
    torch.save({**shard1, **shard2}, single / "pytorch_model.bin")
    m_single = STLLMForCausalLM.from_config(dict(cfg, llama_model=str(single)))
    assert torch.equal(m_single.state_dict()["lm_head.weight"].float(), shard2["lm_head.weight"].half().float())
    try:
        from safetensors.torch import save_file
        st_dir = tmp_path / "vicuna_st"
        st_dir.mkdir()
        (st_dir / "config.json").write_text(_json.dumps(hf_cfg))
        save_file({k: v.contiguous() for k, v in {**shard1, **shard2}.items()}, str(st_dir / "model.safetensors"))
        m_st = STLLMForCausalLM.from_config(dict(cfg, llama_model=str(st_dir)))
        assert torch.equal(m_st.state_dict()["lm_head.weight"].float(), shard2["lm_head.weight"].half().float())
    except ImportError:
        pass
    empty = tmp_path / "vicuna_empty"
    empty.mkdir()
    (empty / "config.json").write_text(_json.dumps(hf_cfg))
    with pytest.raises(FileNotFoundError, match="no LLM weights"):
        STLLMForCausalLM.from_config(dict(cfg, llama_model=str(empty)))
    with pytest.raises(FileNotFoundError):
        STLLMForCausalLM.from_config(dict(cfg, llama_model=str(tmp_path / "does_not_exist"), llama_config=None))
    model2 = STLLMForCausalLM.from_config(dict(cfg, ckpt=str(ck_dir)))
    assert torch.equal(model2.state_dict()["model.stllm_model.llama_proj.bias"].float(),
                       ck["model"]["model.stllm_model.llm_proj.bias"])


def test_packed_weights_are_invalidated_by_parent_load_state_dict_and_in_place_edits():
    """nn.Module.load_state_dict on a PARENT never calls a child's load_state_dict; the packed fp16 copies must
    still be rebuilt (stllm_b200/packing.py: _load_from_state_dict hook + (address, version) signature of every parameter)"""
    model = _small_model()
    llm, vit = model.model, model.model.stllm_model.visual_encoder
    qf = model.model.stllm_model.Qformer.bert
    for m in (llm, vit, qf):
        m._packed = {"stale": True}
        m._mark_packed()
        assert m._packed_valid()
    model.load_state_dict(model.state_dict(), strict=False)          # load on the TOP-LEVEL module
    assert not llm._packed_valid() and not vit._packed_valid() and not qf._packed_valid()
    for m in (llm, vit, qf):
        m._packed = {"stale": True}
        m._mark_packed()
    with torch.no_grad():                                             # in-place edits bump the version counter
        llm.layers[0].mlp.down_proj.weight.mul_(1.0)                  # (edits through `.data` bypass it: call repack())
        vit.blocks[0].attn.qkv.weight.add_(0.0)
    assert not llm._packed_valid() and not vit._packed_valid() and qf._packed_valid()
    model.repack()
    assert qf._packed is None
    llm._packed = {"stale": True}
    llm._mark_packed()
    model.half()                                                      # dtype change -> new storage
    assert not llm._packed_valid()
