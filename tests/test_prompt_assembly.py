"""Sequence assembly (SURVEY 8a rows a17, a18, a21'): stllm_b200/prompt_assembly.py + the host half of the product's
STLLMModel.forward against `assembly_kat.npz` -- written by the reference's OWN STLLMModel.forward (st_llm.py:447-546:
mask draw + compaction, prompt_wrap, concat_emb_input_output, BOS, targets) on a stub whose encoder returns seeded video
tokens (tests/golden/make_golden.py golden_assembly).  Everything here is copies and integer logic: bit-exact."""
import os

import numpy as np
import pytest
import torch

from tests.fake_tokenizer import FakeTokenizer
from tests.util import GOLDEN

CASES = [(name, m) for name in ("text", "minigpt4", "same_prompt") for m in (0, 1)]


def _kat():
    return np.load(os.path.join(GOLDEN, "assembly_kat.npz"))


def _gather(rows, table, video2d):
    """numpy model of ops.gather_rows2: code >= 0 -> table row, code < 0 -> video row -1-code"""
    out = np.empty(rows.shape + (table.shape[1],), dtype=table.dtype)
    pos = rows >= 0
    out[pos] = table[rows[pos]]
    out[~pos] = video2d[-1 - rows[~pos]]
    return out


@pytest.mark.parametrize("name,use_mask", CASES)
def test_plan_sequences_matches_reference_forward(name, use_mask):
    from stllm_b200 import prompt_assembly as PA
    g = _kat()
    key = f"{name}.mask{use_mask}."
    table, video = g["table"], g["video"]
    B, T, Q, D = video.shape
    L = T * Q
    text = name != "minigpt4"
    tok = FakeTokenizer(table.shape[0], pad_token_id=0, bos_token_id=1, eos_token_id=2)
    prompts = [str(p) for p in g[key + "prompts"]]
    before, after = [], []
    for p in prompts:
        b, a = PA.split_prompt(p)
        before.append(tok(b, return_tensors="pt", add_special_tokens=False).input_ids.numpy().reshape(-1))
        after.append(tok(a, return_tensors="pt", add_special_tokens=text).input_ids.numpy().reshape(-1))
    end = tok.eos_token if text else " ###"
    ans = tok([str(t) + end for t in g["answers"]], return_tensors="pt", padding="longest", truncation=True, max_length=8,
              add_special_tokens=False)
    keep = ~g[key + "mask"].reshape(B, L) if use_mask else None
    bos = None if text else tok.bos_token_id
    plan = PA.plan_sequences(PA.video_row_codes(B, L, keep), before, after, ans.input_ids.numpy(),
                             ans.attention_mask.numpy(), tok.pad_token_id, bos)
    video2d = video.reshape(B * L, D)
    assert np.array_equal(plan.attention_mask, g[key + "attention_mask"])
    assert np.array_equal(plan.targets, g[key + "targets"])
    assert np.array_equal(_gather(plan.rows, table, video2d), g[key + "inputs_embeds"])       # bit-exact: pure copies
    if use_mask:
        plan_u = PA.plan_sequences(PA.video_row_codes(B, L), before, after, ans.input_ids.numpy(),
                                   ans.attention_mask.numpy(), tok.pad_token_id, bos)
        assert np.array_equal(plan_u.attention_mask, g[key + "unmask_attention_mask"])
        assert np.array_equal(_gather(plan_u.rows, table, video2d), g[key + "unmask_inputs_embeds"])
        if not text:   # MiniGPT-4 layout: BOS + 7 prefix tokens put the video at position 8 = the hard-coded img_start
            n_before = len(before[0])
            assert 1 + n_before == 8 and np.all(plan_u.rows[0, 8:8 + L] < 0) and plan_u.rows[0, 7] >= 0


def test_split_prompt_needs_exactly_one_placeholder():
    from stllm_b200 import prompt_assembly as PA
    assert PA.split_prompt("a <ImageHere> b") == ("a ", " b")
    with pytest.raises(ValueError):
        PA.split_prompt("no placeholder")
    with pytest.raises(ValueError):
        PA.split_prompt("<ImageHere> x <ImageHere>")


def test_splice_rows_is_concat_emb_input_output():
    # st_llm.py:409-432 on explicit tensors
    from stllm_b200 import prompt_assembly as PA
    rng = np.random.default_rng(0)
    B, W, La, D = 3, 7, 4, 5
    inp, ans = rng.standard_normal((B, W, D)), rng.standard_normal((B, La, D))
    att = np.array([[1] * 7, [1] * 4 + [0] * 3, [1] * 1 + [0] * 6])
    rows = PA.splice_rows(att, La)
    got = _gather(rows, inp.reshape(B * W, D), ans.reshape(B * La, D))
    for i in range(B):
        n = int(att[i].sum())
        want = np.concatenate([inp[i, :n], ans[i], inp[i, n:]])
        assert np.array_equal(got[i], want)


@pytest.mark.parametrize("name,use_mask", CASES)
def test_product_forward_host_half_matches_reference(name, use_mask, monkeypatch):
    """the product's STLLMModel.forward with the encoder stubbed and the gather kernel replaced by its numpy model"""
    from stllm_b200 import ops
    from stllm_b200 import st_llm as S
    g = _kat()
    key = f"{name}.mask{use_mask}."
    table, video = torch.from_numpy(g["table"]), torch.from_numpy(g["video"])
    B, T, Q, D = video.shape
    text = name != "minigpt4"
    sm = S.STLLMModel.__new__(S.STLLMModel)
    torch.nn.Module.__init__(sm)
    sm.pre_encoding, sm.video_input, sm.use_mask, sm.mvm_decode = False, "all", bool(use_mask), False
    sm.qformer_text_input, sm.max_txt_len, sm.end_sym = text, 8, "\n" if text else " ###"
    sm.llama_tokenizer = FakeTokenizer(table.shape[0], pad_token_id=0, bos_token_id=1, eos_token_id=2)
    emb = torch.nn.Embedding(table.shape[0], D)
    emb.weight.data.copy_(table)
    sm.embed_tokens = emb
    sm.encode_img = lambda image, text=None: (video.clone(), torch.ones(B, T, Q, dtype=torch.long), False)
    monkeypatch.setattr(S.STLLMModel, "_gather_plan", lambda self, rows, v2d: torch.from_numpy(
        _gather(rows.numpy(), table.numpy(), v2d.numpy())))
    samples = {"image": torch.zeros(B, T, 3, 2, 2), "instruction_input": [str(p) for p in g[key + "prompts"]],
               "answer": [str(a) for a in g["answers"]]}
    np.random.seed(4321)
    emb_m, att_m, emb_u, att_u, targets = sm.forward(samples)
    assert np.array_equal(emb_m.numpy(), g[key + "inputs_embeds"])
    assert np.array_equal(att_m.numpy(), g[key + "attention_mask"])
    assert np.array_equal(targets.numpy(), g[key + "targets"])
    if use_mask:
        assert np.array_equal(sm.mask.numpy(), g[key + "mask"])
        assert np.array_equal(emb_u.numpy(), g[key + "unmask_inputs_embeds"])
        assert np.array_equal(att_u.numpy(), g[key + "unmask_attention_mask"])
        assert sm.img_len == T * Q and sm.mask_img_len == int((~g[key + "mask"][0, 0]).sum())
    else:
        assert emb_u is None and att_u is None
