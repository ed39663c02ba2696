"""CPU: the decode engine's ALGORITHM (one shared copy of the prompt K/V, per-beam suffix caches that are never moved, an
ancestry table permuted on beam re-ordering, key splits merged after a partial softmax; oracle/decode_oracle.py) computes the
same logits as the reference semantics -- a full recompute of every beam's actual token history
(oracle/stllm_oracle.py:llama_forward, pinned against the reference's own modules in test_oracle_golden.py).
Re-orderings with DUPLICATED ancestors (two beams continuing the same parent, as beam search produces) are included."""
import pytest
import torch

from oracle import stllm_oracle as O
from oracle.decode_oracle import DecodeOracle
from tests.util import load_golden


def _setup():
    sd, g = load_golden("llama_small.npz")
    msd = {k[len("model."):]: v.float() for k, v in sd.items() if k.startswith("model.")}
    return msd, sd["lm_head.weight"].float(), sd["model.embed_tokens.weight"].float(), torch.from_numpy(g["inputs_embeds"]).float()


@pytest.mark.parametrize("beams,n_split,masked", [(1, 1, False), (3, 4, False), (5, 3, True), (4, 7, False)])
def test_shared_prefix_ancestry_table_equals_full_recompute(beams, n_split, masked):
    msd, lm_w, E, emb = _setup()
    S, steps = 12, 7
    prompt = emb[:1, :S]
    mask = None
    if masked:
        mask = torch.ones(S)
        mask[3] = 0
        mask[8] = 0
    orc = DecodeOracle(msd, lm_w, prompt, beams, steps, num_layers=2, num_heads=2, n_split=n_split, prompt_mask=mask)
    gen = torch.Generator().manual_seed(beams * 10 + n_split)
    hist = [[] for _ in range(beams)]
    # the prefill's last-position logits are the reference's too
    ref0 = torch.nn.functional.linear(O.llama_forward(msd, prompt, None if mask is None else mask[None], 2, 2)[:, -1], lm_w)
    assert torch.allclose(orc.prefill_logits, ref0[0], atol=2e-4, rtol=1e-4)
    for t in range(steps):
        tok = torch.randint(3, 128, (beams,), generator=gen)
        if t > 0:
            idx = torch.randint(0, beams, (beams,), generator=gen)      # with repeats: beams sharing an ancestor
            orc.reorder(idx)
            hist = [list(hist[i]) for i in idx.tolist()]
        logits = orc.step(tok, E)
        for b in range(beams):
            hist[b].append(int(tok[b]))
            seq = torch.cat([prompt, E[torch.tensor(hist[b])][None]], dim=1)
            am = None
            if mask is not None:
                am = torch.cat([mask, torch.ones(len(hist[b]))])[None]
            ref = torch.nn.functional.linear(O.llama_forward(msd, seq, am, 2, 2)[:, -1], lm_w)[0]
            err = (logits[b] - ref).abs().max().item()
            assert err < 3e-4, f"step {t} beam {b}: max abs err {err:.2e} (fp32: only the summation order differs)"


def test_ancestry_table_is_the_only_thing_reordered():
    msd, lm_w, E, emb = _setup()
    orc = DecodeOracle(msd, lm_w, emb[:1, :6], 3, 4, num_layers=2, num_heads=2, n_split=2)
    orc.step(torch.tensor([5, 6, 7]), E)
    ks_before = [k.clone() for k in orc.ks]
    orc.reorder(torch.tensor([2, 2, 0]))
    assert orc.src[:, 0].tolist() == [2, 2, 0]
    assert all(torch.equal(a, b) for a, b in zip(ks_before, orc.ks))      # no K/V moved
    orc.step(torch.tensor([8, 9, 10]), E)
    assert orc.src[:, 1].tolist() == [0, 1, 2]                             # new tokens live in their own rows
