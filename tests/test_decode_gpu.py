"""GPU: decode-step kernels (csrc/decode.cu) against torch fp32 statements of the same ops, and the decode engine
(CUDA-graph step, shared prompt KV, ancestry table) against the general-kernel cache route, through the C ABI."""
import pytest

pytestmark = pytest.mark.gpu


def pytest_generate_tests(metafunc):
    if "decode_case" in metafunc.fixturenames:
        import torch
        if torch.cuda.is_available():
            from scripts import gpu_check_decode as K
            names = [n for n, _ in K.build_cases()]
        else:
            names = ["needs-gpu"]
        metafunc.parametrize("decode_case", names)


def test_decode_kernel_or_engine_case(decode_case):
    """one pytest item per case of scripts/gpu_check_decode.py: decode_linear (all norm / epilogue forms, Vicuna shapes),
    rope + cache append (bit-exact), attention over shared prompt K/V + per-beam suffix through the ancestry table (unfused
    and fused), engine vs the general-kernel route"""
    from scripts import gpu_check_decode as K
    fn = dict(K.build_cases())[decode_case]
    K.fails.clear()
    assert fn(), f"{decode_case}: {K.fails}"


def test_engine_matches_decode_oracle_with_duplicated_ancestors():
    """DecodeEngine (fp16, CUDA graph) against oracle/decode_oracle.py (fp32 CPU restatement of the same algorithm, itself
    pinned against a full recompute of every beam's history): beam re-orderings WITH repeats, which the engine-vs-general
    check of scripts/gpu_check_decode.py (permutations only) does not cover."""
    import torch
    from oracle.decode_oracle import DecodeOracle
    from stllm_b200.decode import DecodeEngine
    from stllm_b200.st_llm import STLLMForCausalLM, StllmConfig
    from tests.util import assert_close, load_golden
    sd, g = load_golden("llama_small.npz")
    cfg = StllmConfig(hidden_size=256, num_attention_heads=2, intermediate_size=512, num_hidden_layers=2,
                      vocab_size=128, max_position_embeddings=2048, rms_norm_eps=1e-6, eos_token_id=2)
    llm = STLLMForCausalLM(cfg)
    llm.load_state_dict(sd, strict=False)
    llm = llm.half().cuda()
    msd = {k[len("model."):]: v.float() for k, v in sd.items() if k.startswith("model.")}
    E, lm_w = sd["model.embed_tokens.weight"].float(), sd["lm_head.weight"].float()
    S, beams, steps = 12, 3, 6
    prompt = torch.from_numpy(g["inputs_embeds"]).float()[:1, :S]
    out = llm(samples=None, inputs_embeds=prompt.cuda().half(), use_cache=True, cache_max_len=S + steps + 1)
    eng = DecodeEngine(llm, out.past_key_values, S, beams, steps)
    orc = DecodeOracle(msd, lm_w, prompt, beams, steps, num_layers=2, num_heads=2, n_split=eng.n_split)
    gen = torch.Generator().manual_seed(4)
    for t in range(steps):
        tok = torch.randint(3, 128, (beams,), generator=gen)
        if t > 0:
            idx = torch.randint(0, beams, (beams,), generator=gen)
            eng.reorder(idx.cuda())
            orc.reorder(idx)
        got = eng.step(tok.cuda()).float().cpu()
        ref = orc.step(tok, E)
        assert_close(got, ref, atol=5e-2, rtol=2e-2, what=f"engine vs decode oracle, step {t}")
