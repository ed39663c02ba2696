"""Token generation over the kernel forward: greedy / sampling / beam search with a KV cache.

The reference inherits HF `GenerationMixin.generate` (call sites conversation.py:231-243,
mvbench_conversation.py:114-125: inputs_embeds=[1,S,4096], max_new_tokens, stopping_criteria, num_beams,
do_sample, min_length, top_p, repetition_penalty, length_penalty, temperature).  This is a self-contained
implementation of those options.  The prompt is prefilled ONCE and its KV cache is shared by all beams (HF prefills
`num_beams` identical rows).  Returns generated token ids [B, n] (no dummy BOS column; the reference
callers strip a leading 0/1 token only if present, conversation.py:245-248).

Token selection follows HF exactly -- it is integer work, pinned bit-exactly:
  * beam search scores are log-softmax FIRST, then the repetition penalty / min_length processors on the log-probs, then
    the running beam score (HF `beam_search`; a penalised log-prob is not renormalised);
  * `hf_semantics` selects the release whose bookkeeping is reproduced: "4.28" (what the reference pins, requirement.txt:42;
    default) or "5" (the release installed in the build container, which is what runs when the reference's own generate is
    executed there: tests/golden/generate_small.npz holds ITS token ids and tests/test_generate_*.py require identical ids
    from this module in "5" mode).  The differences are confined to `_BeamBookV4` / `_BeamBookV5` and
    `_effective_min_length`; oracle/generate_oracle.py states them.
The per-step device work is one [beams, vocab] log-softmax + top-k; the candidates (2 x beams numbers) cross to the host
in one copy, the surviving (beam, token, score) triples go back in one copy.

Decode steps: for batch 1 (every reference call site) the steps run on `decode.DecodeEngine` -- the HBM-bound kernels of
csrc/decode.cu replayed as one CUDA graph per token, prompt K/V shared by the beams, ancestry table instead of cache
re-ordering.  Other shapes (batch > 1, head dim != 128) take the general kernels through model.forward with a KVCache.
STLLM_DECODE_ENGINE=0 forces the general route.
"""
import os

import torch

from . import decode as _decode


def _default_semantics():
    return os.environ.get("STLLM_HF_SEMANTICS", "4.28")


def _effective_min_length(min_length, prompt_len, semantics):
    """"5": generate(inputs_embeds=...) subtracts the prompt length from min_length (`_prepare_generated_length`), so the
    harnesses' min_length=1 never suppresses EOS; "4.28": the processor sees [dummy BOS ; generated] and min_length as given"""
    if semantics == "4.28":
        return min_length or 0
    return max((min_length or 0) - prompt_len, 0)


def _process_scores(scores, history, min_length, eos_ids, repetition_penalty):
    """RepetitionPenaltyLogitsProcessor + MinLengthLogitsProcessor on `scores` [rows, V] (raw logits for greedy / sampling,
    log-probabilities for beam search); history [rows, n] int64 = the ids the processors see."""
    if repetition_penalty != 1.0 and history.shape[1] > 0:
        score = torch.gather(scores, 1, history)
        score = torch.where(score < 0, score * repetition_penalty, score / repetition_penalty)
        scores = scores.scatter(1, history, score)
    if min_length > 0 and history.shape[1] < min_length:
        for e in eos_ids:
            scores[:, e] = -float("inf")
    return scores


def _top_p_filter(logits, top_p):
    if top_p >= 1.0:
        return logits
    sorted_logits, sorted_idx = torch.sort(logits, descending=False)
    cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    remove = cum <= (1 - top_p)
    remove[..., -1:] = False
    remove = remove.scatter(1, sorted_idx, remove)
    return logits.masked_fill(remove, -float("inf"))


def _stopped(stopping_criteria, ids, scores=None):
    if not stopping_criteria:
        return False
    for crit in stopping_criteria:
        r = crit(ids, scores)
        if bool(r.all() if isinstance(r, torch.Tensor) else r):
            return True
    return False


# ----------------------------------------------------------------------------------------------------------------------
# beam bookkeeping (host, tiny): candidates of one step in, surviving beams out
# ----------------------------------------------------------------------------------------------------------------------
class _BeamBookV4:
    """HF 4.28 `BeamSearchScorer.process` / `finalize` + `BeamHypotheses` (early_stopping=False, one batch item).
    Sequences here exclude the dummy BOS column HF prepends for inputs_embeds-only generation; `dummy` = 1 adds it back
    wherever a LENGTH enters a score."""

    def __init__(self, nb, max_new_tokens, eos_ids, length_penalty, custom_stop):
        self.nb, self.max_new, self.eos, self.lp, self.custom = nb, max_new_tokens, tuple(eos_ids), length_penalty, custom_stop
        self.running = [[] for _ in range(nb)]
        self.scores = [0.0] + [-1e9] * (nb - 1)
        self.hyps, self.worst, self.done, self.dummy = [], 1e9, False, 1

    def _add(self, tokens, sum_logprobs):
        score = sum_logprobs / ((len(tokens) + self.dummy) ** self.lp)
        if len(self.hyps) < self.nb or score > self.worst:
            self.hyps.append((score, list(tokens)))
            if len(self.hyps) > self.nb:
                order = sorted((s, i) for i, (s, _) in enumerate(self.hyps))
                del self.hyps[order[0][1]]
                self.worst = order[1][0]
            else:
                self.worst = min(score, self.worst)

    def step(self, cands):
        """cands: [(score, beam, token)] sorted by score, best first.  Returns (next beams [(beam, token, score)], finished)"""
        nxt = []
        for rank, (s, b, t) in enumerate(cands):
            if t in self.eos:
                if rank >= self.nb:
                    continue
                self._add(self.running[b], s)
            else:
                nxt.append((b, t, s))
            if len(nxt) == self.nb:
                break
        if len(nxt) < self.nb:
            raise RuntimeError("beam search: fewer than num_beams non-EOS candidates survived")
        cur_len = len(self.running[0]) + self.dummy + 1          # the length the candidate scores were computed on
        if len(self.hyps) >= self.nb and self.worst >= max(c[0] for c in cands[:2 * self.nb]) / (cur_len ** self.lp):
            self.done = True
        self.running = [self.running[b] + [t] for (b, t, _) in nxt]
        self.scores = [s for (_, _, s) in nxt]
        stop = self.done or len(self.running[0]) >= self.max_new or self.custom(self.running[0])
        return nxt, stop

    def finalize(self):
        if not self.done:
            for b in range(self.nb):
                self._add(self.running[b], self.scores[b])
        best = sorted(self.hyps, key=lambda x: x[0])[-1][1]
        if len(best) < self.max_new and self.eos:
            best = best + [self.eos[0]]
        return best


class _BeamBookV5:
    """HF 5.x `_beam_search` (vectorised there): stopping criteria -- EOS, max length, custom -- FINISH individual candidates;
    only the top num_beams candidates of a step may finish; the search ends when no running beam can still beat the worst
    finished one (`_check_early_stop_heuristic`, early_stopping=False) or every candidate hit a criterion."""

    def __init__(self, nb, max_new_tokens, eos_ids, length_penalty, custom_stop):
        self.nb, self.max_new, self.eos, self.lp, self.custom = nb, max_new_tokens, tuple(eos_ids), length_penalty, custom_stop
        self.running = [[] for _ in range(nb)]
        self.scores = [0.0] + [-1e9] * (nb - 1)
        self.fin = [(-1.0e9, [], False) for _ in range(nb)]      # (score, tokens, finished), best first
        self.unsat, self.cur_len = True, 0

    def step(self, cands):
        f32 = lambda x: float(torch.tensor(x, dtype=torch.float32))
        seqs = [self.running[b] + [t] for (_, b, t) in cands]
        row0 = self.custom(seqs[0])                               # old-style criteria inspect row 0: one verdict for all rows
        hits = [(q[-1] in self.eos) or len(q) >= self.max_new or row0 for q in seqs]
        run = [f32(f32(s) + (-1.0e9 if h else 0.0)) for (s, _, _), h in zip(cands, hits)]
        order = sorted(range(len(cands)), key=lambda i: (-run[i], i))[:self.nb]
        nxt = [(cands[i][1], cands[i][2], run[i]) for i in order]
        merged = list(self.fin)
        for i, ((s, _, _), q) in enumerate(zip(cands, seqs)):
            just = hits[i] and i < self.nb
            sc = f32(f32(s) / ((self.cur_len + 1) ** self.lp))
            sc = f32(sc + (0.0 if self.unsat else -1.0e9))
            sc = f32(sc + (0.0 if just else -1.0e9))
            merged.append((sc, q, just))
        keep = sorted(range(len(merged)), key=lambda i: (-merged[i][0], i))[:self.nb]
        self.fin = [merged[i] for i in keep]
        self.running = [seqs[i] for i in order]
        self.scores = [run[i] for i in order]
        self.cur_len += 1
        best_possible = f32(self.scores[0] / (self.cur_len ** self.lp))
        low = min(f[0] for f in self.fin)
        self.unsat = self.unsat and any(best_possible > (low if f[2] else -1.0e9) for f in self.fin)
        return nxt, (not self.unsat) or all(hits)

    def finalize(self):
        return list(self.fin[0][1])


@torch.no_grad()
def generate(model, inputs_embeds=None, input_ids=None, max_new_tokens=20, stopping_criteria=None, num_beams=1,
             do_sample=False, min_length=0, top_p=1.0, repetition_penalty=1.0, length_penalty=1.0, temperature=1.0,
             eos_token_id=None, pad_token_id=None, attention_mask=None, hf_semantics=None):
    cfg = model.config
    sem = hf_semantics or _default_semantics()
    if sem not in ("4.28", "5"):
        raise ValueError(f"hf_semantics must be '4.28' or '5', got {sem!r}")
    if do_sample and num_beams > 1:
        raise NotImplementedError("beam_sample (do_sample=True with num_beams > 1) is not implemented; no reference call "
                                  "site uses it (conversation.py:231-243 passes do_sample=False whenever num_beams=5)")
    eos = cfg.eos_token_id if eos_token_id is None else eos_token_id
    eos_ids = [] if eos is None else (list(eos) if isinstance(eos, (list, tuple)) else [eos])
    pad = pad_token_id if pad_token_id is not None else (cfg.pad_token_id if cfg.pad_token_id is not None else 0)
    bos = cfg.bos_token_id if getattr(cfg, "bos_token_id", None) is not None else 1
    embeds_only = inputs_embeds is not None
    if inputs_embeds is None:
        inputs_embeds = model.get_model()._embed_ids(input_ids)
    B, S, _ = inputs_embeds.shape
    dev = inputs_embeds.device
    max_len = min(cfg.max_position_embeddings, S + max_new_tokens)
    if max_len <= S:
        return torch.zeros(B, 0, dtype=torch.long, device=dev)
    min_len = _effective_min_length(min_length, S if embeds_only else 0, sem)
    # ids the logits processors see in front of the generated ones: HF 4.28 seeds inputs_embeds-only generation with a
    # dummy BOS column, otherwise the prompt ids (when the caller passed ids rather than embeddings)
    if embeds_only:
        prefix_ids = torch.full((B, 1), bos, dtype=torch.long, device=dev) if sem == "4.28" else \
            torch.zeros(B, 0, dtype=torch.long, device=dev)
    else:
        prefix_ids = input_ids.to(torch.long)
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
            hist = torch.cat([prefix_ids, generated], dim=1)
            lg = _process_scores(logits.to(torch.float32, copy=True), hist, min_len, eos_ids, repetition_penalty)
            if do_sample:
                if temperature != 1.0:
                    lg = lg / temperature
                probs = _top_p_filter(lg, top_p).softmax(dim=-1)
                nxt = torch.multinomial(probs, 1).squeeze(1)
            else:
                nxt = lg.argmax(dim=-1)
            nxt = torch.where(finished, torch.full_like(nxt, pad), nxt)
            generated = torch.cat([generated, nxt[:, None]], dim=1)
            for e in eos_ids:
                finished |= nxt == e
            seen = torch.cat([prefix_ids, generated], dim=1)   # what HF hands to the stopping criteria
            if bool(finished.all()) or _stopped(stopping_criteria, seen, lg) or step + 1 == n_steps:
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
    K = max(2, 1 + len(eos_ids)) * nb
    pre = prefix_ids[0].tolist()

    def custom_stop(tokens):
        if not stopping_criteria:
            return False
        return _stopped(stopping_criteria, torch.tensor([pre + list(tokens)], dtype=torch.long, device=dev))
    book = (_BeamBookV4 if sem == "4.28" else _BeamBookV5)(nb, n_steps, eos_ids, length_penalty, custom_stop)
    beam_scores = torch.tensor(book.scores, dtype=torch.float32, device=dev)
    seqs = torch.zeros(nb, 0, dtype=torch.long, device=dev)
    prefix_rows = prefix_ids.expand(nb, -1)
    for step in range(n_steps):
        logp = logits.float().log_softmax(dim=-1)                       # HF: log-softmax first ...
        hist = torch.cat([prefix_rows, seqs], dim=1)
        logp = _process_scores(logp, hist, min_len, eos_ids, repetition_penalty)    # ... then the processors ...
        logp = logp + beam_scores[:, None]                              # ... then the running beam scores
        V = logp.shape[1]
        top_s, top_i = logp.view(-1).topk(K)
        # one device->host transfer (one synchronisation) for the K candidates; indices are exact in float64
        cand = torch.stack([top_s.double(), top_i.double()]).cpu()
        cands = [(s, int(i) // V, int(i) % V) for s, i in zip(cand[0].tolist(), cand[1].tolist())]
        nxt, stop = book.step(cands)
        # one host->device transfer for (source beam, token, score) of the surviving beams
        pk = torch.tensor([[b for b, _, _ in nxt], [t for _, t, _ in nxt], [s for _, _, s in nxt]],
                          dtype=torch.float64).to(dev, non_blocking=True)
        idx, tok, beam_scores = pk[0].long(), pk[1].long(), pk[2].float()
        seqs = torch.cat([seqs[idx], tok[:, None]], dim=1)
        if stop or step + 1 == n_steps:
            break
        if engine is not None:
            engine.reorder(idx)
            logits = engine.step(tok)
        else:
            cache.reorder(idx)
            out = model.forward(input_ids=tok[:, None], past_key_values=cache, use_cache=True)
            logits = out.logits[:, -1]
    best = book.finalize()
    return torch.tensor([best], dtype=torch.long, device=dev)
