"""Token generation over the kernel forward: greedy / sampling / beam search with a KV cache.

The reference inherits HF `GenerationMixin.generate` (call sites conversation.py:231-243,
mvbench_conversation.py:114-125: inputs_embeds=[1,S,4096], max_new_tokens, stopping_criteria, num_beams,
do_sample, min_length, top_p, repetition_penalty, length_penalty, temperature).  This is a self-contained
implementation of those options.  The prompt is prefetched ONCE and its KV cache is shared by all beams (HF 4.28
prefills `num_beams` identical rows).  Returns generated token ids [B, n] (no dummy BOS column;
the reference callers strip a leading 0/1 token only if present, conversation.py:245-248).
Sampling/beam bookkeeping is host-side torch on tiny [beams, vocab] tensors; every model forward runs in the
sm_100a kernels.

Decode steps: for batch 1 (every reference call site) the steps run on `decode.DecodeEngine` -- the HBM-bound kernels of
csrc/decode.cu replayed as one CUDA graph per token, prompt K/V shared by the beams, ancestry table instead of cache
re-ordering.  Other shapes (batch > 1, head dim != 128) take the general kernels through model.forward with a KVCache.
STLLM_DECODE_ENGINE=0 forces the general route.
"""
import os

import torch

from . import decode as _decode


def _process_logits(logits, generated, cur_len, min_length, eos_ids, repetition_penalty, temperature):
    logits = logits.to(torch.float32, copy=True)   # never write into the caller's (possibly expanded) logits
    if repetition_penalty != 1.0 and generated.shape[1] > 0:
        score = torch.gather(logits, 1, generated)
        score = torch.where(score < 0, score * repetition_penalty, score / repetition_penalty)
        logits = logits.scatter(1, generated, score)
    if cur_len < min_length:
        for e in eos_ids:
            logits[:, e] = -float("inf")
    if temperature != 1.0:
        logits = logits / temperature
    return logits


def _top_p_filter(logits, top_p):
    if top_p >= 1.0:
        return logits
    sorted_logits, sorted_idx = torch.sort(logits, descending=False)
    cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    remove = cum <= (1 - top_p)
    remove[..., -1:] = False
    remove = remove.scatter(1, sorted_idx, remove)
    return logits.masked_fill(remove, -float("inf"))


def _stopped(stopping_criteria, ids, scores):
    if not stopping_criteria:
        return False
    for crit in stopping_criteria:
        r = crit(ids, scores)
        if bool(r.all() if isinstance(r, torch.Tensor) else r):
            return True
    return False


@torch.no_grad()
def generate(model, inputs_embeds=None, input_ids=None, max_new_tokens=20, stopping_criteria=None, num_beams=1,
             do_sample=False, min_length=1, top_p=1.0, repetition_penalty=1.0, length_penalty=1.0, temperature=1.0,
             eos_token_id=None, pad_token_id=None, attention_mask=None):
    cfg = model.config
    eos = cfg.eos_token_id if eos_token_id is None else eos_token_id
    eos_ids = [] if eos is None else (list(eos) if isinstance(eos, (list, tuple)) else [eos])
    pad = pad_token_id if pad_token_id is not None else (cfg.pad_token_id if cfg.pad_token_id is not None else 0)
    if inputs_embeds is None:
        inputs_embeds = model.get_model()._embed_ids(input_ids)
    B, S, _ = inputs_embeds.shape
    dev = inputs_embeds.device
    max_len = min(cfg.max_position_embeddings, S + max_new_tokens)
    if max_len <= S:
        return torch.zeros(B, 0, dtype=torch.long, device=dev)
    out = model.forward(inputs_embeds=inputs_embeds, attention_mask=attention_mask, use_cache=True,
                        cache_max_len=max_len)
    cache = out.past_key_values
    logits = out.logits[:, -1]
    n_steps = max_len - S
    engine = None
    n_rows = max(1, num_beams)
    if B == 1 and n_steps > 1 and inputs_embeds.is_cuda and os.environ.get("STLLM_DECODE_ENGINE", "1") != "0" \
            and _decode.supported(model, n_rows):
        engine = _decode.DecodeEngine(model, cache, S, n_rows, n_steps - 1, prefix_mask=attention_mask)
    if num_beams <= 1:
        generated = torch.zeros(B, 0, dtype=torch.long, device=dev)
        finished = torch.zeros(B, dtype=torch.bool, device=dev)
        for step in range(n_steps):
            lg = _process_logits(logits, generated, generated.shape[1] + 1, min_length, eos_ids, repetition_penalty,
                                 temperature if do_sample else 1.0)
            if do_sample:
                probs = _top_p_filter(lg, top_p).softmax(dim=-1)
                nxt = torch.multinomial(probs, 1).squeeze(1)
            else:
                nxt = lg.argmax(dim=-1)
            nxt = torch.where(finished, torch.full_like(nxt, pad), nxt)
            generated = torch.cat([generated, nxt[:, None]], dim=1)
            for e in eos_ids:
                finished |= nxt == e
            if bool(finished.all()) or _stopped(stopping_criteria, generated, lg) or step + 1 == n_steps:
                break
            if engine is not None:
                logits = engine.step(nxt)
            else:
                out = model.forward(input_ids=nxt[:, None], past_key_values=cache, use_cache=True)
                logits = out.logits[:, -1]
        return generated
    # ---------------- beam search (batch 1, as every reference call site uses) ----------------
    assert B == 1, "beam search is implemented for batch size 1 (the reference's harnesses generate one clip at a time)"
    nb = num_beams
    if engine is None:
        cache.expand_batch(nb)
    logits = logits.expand(nb, -1)
    beam_scores = torch.zeros(nb, device=dev)
    beam_scores[1:] = -1e9
    seqs = torch.zeros(nb, 0, dtype=torch.long, device=dev)
    hyps = []  # (normalised score, tokens)
    worst = 1e9
    done = False
    for step in range(n_steps):
        cur_len = seqs.shape[1] + 1
        lg = _process_logits(logits, seqs, cur_len, min_length, eos_ids, repetition_penalty, 1.0)
        logp = lg.log_softmax(dim=-1) + beam_scores[:, None]
        V = logp.shape[1]
        top_s, top_i = logp.view(-1).topk(2 * nb)
        # one device->host transfer (one synchronisation) for the 2*nb candidates; indices are exact in float64
        cand = torch.stack([top_s.double(), top_i.double()]).cpu()
        cand_s, cand_i = cand[0].tolist(), [int(v) for v in cand[1].tolist()]
        nb_idx, nb_tok, nb_score = [], [], []
        for rank, (s, i) in enumerate(zip(cand_s, cand_i)):
            b, t = i // V, i % V
            if t in eos_ids:
                if rank >= nb:
                    continue
                L = max(seqs.shape[1], 1)
                sc = s / (L ** length_penalty)
                hyps.append((sc, seqs[b].clone()))
                hyps.sort(key=lambda h: -h[0])
                hyps = hyps[:nb]
                worst = hyps[-1][0] if len(hyps) >= nb else -1e9
            else:
                nb_idx.append(b); nb_tok.append(t); nb_score.append(s)
            if len(nb_idx) == nb:
                break
        if len(hyps) >= nb:
            best_possible = max(cand_s) / (cur_len ** length_penalty)
            if worst >= best_possible:
                done = True
        # one host->device transfer for (source beam, token, score) of the surviving beams
        pk = torch.tensor([nb_idx, nb_tok, nb_score], dtype=torch.float64).to(dev, non_blocking=True)
        idx, tok, beam_scores = pk[0].long(), pk[1].long(), pk[2].float()
        seqs = torch.cat([seqs[idx], tok[:, None]], dim=1)
        if done or _stopped(stopping_criteria, seqs[:1], lg) or step + 1 == n_steps:
            break
        if engine is not None:
            engine.reorder(idx)
            logits = engine.step(tok)
        else:
            cache.reorder(idx)
            out = model.forward(input_ids=tok[:, None], past_key_values=cache, use_cache=True)
            logits = out.logits[:, -1]
    if not done:
        for b in range(nb):
            sc = beam_scores[b].item() / (max(seqs.shape[1], 1) ** length_penalty)
            hyps.append((sc, seqs[b].clone()))
    hyps.sort(key=lambda h: -h[0])
    best = hyps[0][1]
    if len(eos_ids) and best.numel() < n_steps:
        best = torch.cat([best, torch.tensor([eos_ids[0]], device=dev)])
    return best[None]
