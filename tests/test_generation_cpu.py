"""CPU: host logic of generate() (stllm_b200/generation.py: greedy / sampling filters / beam search bookkeeping, the part of
SURVEY 8f rank 1 that is not a kernel) on a deterministic fake language model -- no kernels are involved: the fake's
`forward` is a table lookup on the token history, its cache implements expand_batch / reorder like KVCache.
With at least as many beams as there are partial sequences the search is exhaustive and must return the globally best
sequence; greedy must follow the step-wise argmax; EOS / min_length / repetition-penalty rules are checked as properties."""
import itertools
import math

import pytest
import torch

from stllm_b200 import generation as G


class _Cfg:
    def __init__(self, vocab, eos, max_pos=64):
        self.vocab_size = vocab
        self.eos_token_id = eos
        self.pad_token_id = 0
        self.max_position_embeddings = max_pos


class _Out:
    def __init__(self, logits, cache):
        self.logits = logits
        self.past_key_values = cache


class _HistCache:
    """per-row token history; the same interface generate() uses on KVCache"""

    def __init__(self, rows):
        self.hist = [[] for _ in range(rows)]

    def expand_batch(self, n):
        self.hist = [list(self.hist[0]) for _ in range(n)]

    def reorder(self, idx):
        self.hist = [list(self.hist[i]) for i in idx.tolist()]


class FakeLM:
    """next-token logits = table[hash(last two tokens)] (history dependent, so beams differ)"""

    def __init__(self, vocab=4, eos=-1, seed=0, scale=2.0):
        self.config = _Cfg(vocab, eos)
        g = torch.Generator().manual_seed(seed)
        self.M = 97
        self.table = torch.randn(self.M, vocab, generator=g) * scale
        self.calls = 0

    def get_model(self):
        return self

    def _logits_for(self, hist):
        a = hist[-1] if len(hist) > 0 else 7
        b = hist[-2] if len(hist) > 1 else 3
        return self.table[(a * 31 + b * 17 + len(hist)) % self.M]

    def forward(self, inputs_embeds=None, input_ids=None, attention_mask=None, past_key_values=None, use_cache=False,
                cache_max_len=None):
        self.calls += 1
        if past_key_values is None:
            B, S, _ = inputs_embeds.shape
            cache = _HistCache(B)
            logits = torch.stack([self._logits_for(h) for h in cache.hist])[:, None, :].expand(B, S, -1)
            return _Out(logits, cache)
        cache = past_key_values
        toks = input_ids.reshape(-1).tolist()
        assert len(toks) == len(cache.hist)
        for h, t in zip(cache.hist, toks):
            h.append(int(t))
        return _Out(torch.stack([self._logits_for(h) for h in cache.hist])[:, None, :], cache)

    def seq_logprob(self, seq):
        lp, hist = 0.0, []
        for t in seq:
            lp += float(torch.log_softmax(self._logits_for(hist).float(), -1)[t])
            hist.append(t)
        return lp


def _emb(B=1, S=3):
    return torch.zeros(B, S, 8)


def test_greedy_follows_stepwise_argmax_and_batches_rows():
    lm = FakeLM(vocab=6, eos=-1, seed=1)
    out = G.generate(lm, inputs_embeds=_emb(B=2), max_new_tokens=7, num_beams=1)
    assert out.shape == (2, 7)
    for row in out.tolist():
        hist = []
        for t in row:
            assert t == int(lm._logits_for(hist).argmax())
            hist.append(t)
    assert lm.calls == 7   # one prefill + 6 cached steps: the last token needs no forward


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_exhaustive_beam_search_finds_the_global_optimum(seed):
    V, n = 4, 3
    lm = FakeLM(vocab=V, eos=-1, seed=seed)
    best = max(itertools.product(range(V), repeat=n), key=lm.seq_logprob)
    out = G.generate(lm, inputs_embeds=_emb(), max_new_tokens=n, num_beams=V ** n, length_penalty=1.0)
    assert out.shape == (1, n)
    assert tuple(out[0].tolist()) == best
    # a narrow beam can only be worse or equal, and never worse than greedy
    narrow = G.generate(lm, inputs_embeds=_emb(), max_new_tokens=n, num_beams=2)
    greedy = G.generate(lm, inputs_embeds=_emb(), max_new_tokens=n, num_beams=1)
    lp = lambda t: lm.seq_logprob(t[0].tolist())
    assert lp(out) + 1e-6 >= lp(narrow) >= lp(greedy) - 1e-6


def test_beam_search_beats_greedy_on_a_garden_path():
    """token 1 looks best first but leads to a flat continuation; token 2 leads to a certain one"""
    lm = FakeLM(vocab=4, eos=-1)
    flat = torch.zeros(4)
    sharp = torch.tensor([-20.0, -20.0, -20.0, 5.0])

    def logits_for(hist):
        if len(hist) == 0:
            return torch.tensor([-9.0, 1.0, 0.8, -9.0])
        return flat if hist[0] == 1 else sharp
    lm._logits_for = logits_for
    greedy = G.generate(lm, inputs_embeds=_emb(), max_new_tokens=3, num_beams=1)
    beam = G.generate(lm, inputs_embeds=_emb(), max_new_tokens=3, num_beams=2)
    assert greedy[0, 0].item() == 1 and beam[0].tolist() == [2, 3, 3]


def test_eos_min_length_and_padding_rules():
    eos = 2
    lm = FakeLM(vocab=5, eos=eos, seed=5)
    always_eos = torch.tensor([0.0, 0.0, 9.0, 0.0, 0.0])
    lm._logits_for = lambda hist: always_eos
    # min_length keeps EOS out of the first min_length - 1 positions (HF MinLengthLogitsProcessor: cur_len < min_length)
    out = G.generate(lm, inputs_embeds=_emb(B=2), max_new_tokens=8, num_beams=1, min_length=4)
    assert out.shape[1] == 4 and (out[:, :3] != eos).all() and (out[:, 3] == eos).all()
    beam = G.generate(lm, inputs_embeds=_emb(), max_new_tokens=8, num_beams=3, min_length=4)
    toks = beam[0].tolist()
    assert toks[-1] == eos and eos not in toks[:-1] and len(toks) >= 4
    # a finished row is padded while the other row keeps generating
    def logits_for(hist):
        return always_eos if (len(hist) >= 1 and hist[0] == 1) else torch.tensor([0.0, 5.0, -9.0, 4.0, 0.0]) \
            if len(hist) == 0 else torch.tensor([0.0, 0.0, -9.0, 5.0, 0.0])
    lm._logits_for = logits_for
    out = G.generate(lm, inputs_embeds=_emb(B=1), max_new_tokens=5, num_beams=1)
    assert out[0].tolist() == [1, eos]


def test_repetition_penalty_and_top_p_filter():
    logits = torch.tensor([[2.0, -1.0, 0.5, 0.0]])
    gen = torch.tensor([[0, 1]])
    out = G._process_scores(logits.clone(), gen, min_length=1, eos_ids=[], repetition_penalty=2.0)
    assert torch.allclose(out, torch.tensor([[1.0, -2.0, 0.5, 0.0]]))   # positive / 2, negative * 2 (HF rule)
    out = G._process_scores(logits.clone(), gen, min_length=3, eos_ids=[2], repetition_penalty=1.0)
    assert math.isinf(float(out[0, 2])) and torch.allclose(out[0, [0, 1, 3]], torch.tensor([2.0, -1.0, 0.0]))
    out = G._process_scores(logits.clone(), gen, min_length=2, eos_ids=[2], repetition_penalty=1.0)
    assert torch.isfinite(out).all()                                     # history as long as min_length: EOS allowed
    p = torch.tensor([[0.5, 0.3, 0.15, 0.05]]).log()
    kept = torch.isfinite(G._top_p_filter(p, 0.8))
    assert kept.tolist() == [[True, True, False, False]]                 # smallest set with mass >= 0.8
    assert torch.isfinite(G._top_p_filter(p, 1.0)).all()


def test_min_length_counts_the_prompt_in_hf5_and_the_dummy_bos_in_hf428():
    assert G._effective_min_length(1, 576, "5") == 0          # the harnesses' min_length=1 is a no-op on either release
    assert G._effective_min_length(580, 576, "5") == 4
    assert G._effective_min_length(1, 576, "4.28") == 1       # [dummy BOS] already has length 1: EOS allowed at once
    lm = FakeLM(vocab=5, eos=2, seed=5)
    lm._logits_for = lambda hist: torch.tensor([0.0, 0.0, 9.0, 0.0, 0.0])
    for sem, want_len in (("4.28", 4), ("5", 7)):             # prompt S=3: "5" bans EOS for min_length - 3 = ... tokens
        out = G.generate(lm, inputs_embeds=_emb(), max_new_tokens=12, num_beams=1, min_length=4 if sem == "4.28" else 9,
                         hf_semantics=sem)
        assert out.shape[1] == want_len and out[0, -1].item() == 2 and (out[0, :-1] != 2).all(), (sem, out)


def test_beam_sample_is_rejected_loudly():
    with pytest.raises(NotImplementedError):
        G.generate(FakeLM(), inputs_embeds=_emb(), max_new_tokens=3, num_beams=2, do_sample=True)


def test_sampling_is_seeded_and_respects_top_p():
    lm = FakeLM(vocab=6, eos=-1, seed=9, scale=3.0)
    torch.manual_seed(123)
    a = G.generate(lm, inputs_embeds=_emb(), max_new_tokens=6, num_beams=1, do_sample=True, top_p=0.9, temperature=0.7)
    torch.manual_seed(123)
    b = G.generate(lm, inputs_embeds=_emb(), max_new_tokens=6, num_beams=1, do_sample=True, top_p=0.9, temperature=0.7)
    assert a.tolist() == b.tolist() and a.shape == (1, 6)
    # top_p -> 0 keeps only the most probable token: sampling degenerates to greedy
    torch.manual_seed(7)
    c = G.generate(lm, inputs_embeds=_emb(), max_new_tokens=6, num_beams=1, do_sample=True, top_p=1e-6)
    g = G.generate(lm, inputs_embeds=_emb(), max_new_tokens=6, num_beams=1)
    assert c.tolist() == g.tolist()


def test_decode_engine_shape_gate():
    from stllm_b200 import decode

    class L:
        pass

    def llm(C, F, H):
        o = L()
        o.config = L()
        o.config.hidden_size, o.config.intermediate_size, o.config.num_attention_heads = C, F, H
        return o
    assert decode.supported(llm(4096, 11008, 32), 1) and decode.supported(llm(4096, 11008, 32), 8)
    assert decode.supported(llm(256, 512, 2), 5)
    assert not decode.supported(llm(4096, 11008, 32), 9)        # more rows than an mma n-tile of beams
    assert not decode.supported(llm(4096, 11008, 64), 1)        # head dim 64
    assert not decode.supported(llm(4096, 11000, 32), 1)        # K not a multiple of 32
    assert not decode.supported(llm(8192, 28672, 64), 8)        # 8 x 28672 activations do not fit in shared memory
