"""TEST INFRASTRUCTURE (oracle): CPU restatement of the token-selection algorithm behind the reference's
`STLLMForCausalLM.generate(inputs_embeds=...)` -- HuggingFace `GenerationMixin` (third-party, NOT under /root/reference:
`transformers==4.28.0`, requirement.txt:42; call sites conversation.py:231-243, mvbench_conversation.py:114-125:
num_beams 1 | 5, do_sample False in every harness, min_length 1, repetition_penalty 1.0 | 1.5, length_penalty 1,
stopping_criteria = StoppingCriteriaSub on '###').

Two restatements, selected by `semantics`:

  "5"    -- the transformers release INSTALLED in the build container (5.5.0), which is what executes when the
            reference's own generate runs here through oracle/ref_shim.py.  Restated from its published source
            (generation/utils.py `_sample`, `_beam_search`, `_get_top_k_continuations`, `_update_finished_beams`,
            `_check_early_stop_heuristic`; generation/logits_process.py).  PINNED: tests/golden/generate_small.npz holds token
            ids written by the reference's generate; tests/test_generate_oracle.py requires identical ids (and fuzzes this
            file against the live library where /root/reference exists).
  "4.28" -- the release the reference pins.  Its source is not available offline; restated from the published algorithm
            (generation/utils.py `greedy_search` / `beam_search`, generation/beam_search.py `BeamSearchScorer.process` /
            `finalize`, `BeamHypotheses.add` / `is_done`).  PARITY UNPINNED for this mode: no fixture can be produced
            here.  The known differences from "5" are exactly the branches on `self.v4` below:
              * inputs_embeds-only generation seeds input_ids with a dummy BOS column: it counts for min_length, is seen by
                the repetition penalty, and is part of every hypothesis length (and of the returned ids);
              * 2 * num_beams candidates per step; a hypothesis ending in EOS is scored s / len(input_ids)^lp;
              * early stop: worst kept score >= best candidate score / (len(input_ids) + 1)^lp;
              * a custom stopping criterion (or max_length) ends the WHOLE search; the running beams are then added as
                hypotheses with s / len(input_ids)^lp, and EOS is appended to a returned hypothesis shorter than max_length.

`logits_fn(histories) -> [n, V] float tensor` gives the next-token logits for each token history (list of ints); the
tests back it with oracle.stllm_oracle.llama_forward (full recompute) or with a table.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
import math

import torch


def _apply_processors(scores, histories, min_length, eos_ids, repetition_penalty):
    """RepetitionPenaltyLogitsProcessor then MinLengthLogitsProcessor (the order `_get_logits_processor` builds them in),
    on whatever `scores` are (raw logits for greedy / sampling, log-probabilities for beam search)."""
    scores = scores.clone()
    for r, hist in enumerate(histories):
        if repetition_penalty != 1.0 and len(hist) > 0:
            ids = torch.tensor(sorted(set(hist)), dtype=torch.long)
            s = scores[r, ids]
            scores[r, ids] = torch.where(s < 0, s * repetition_penalty, s / repetition_penalty)
        if min_length is not None and min_length > 0 and eos_ids and len(hist) < min_length:
            scores[r, torch.tensor(list(eos_ids), dtype=torch.long)] = -math.inf
    return scores


def _custom_stop(stopping, seq):
    """the reference's StoppingCriteriaSub protocol: callable(input_ids [1, n] tensor, scores) -> bool"""
    if not stopping:
        return False
    ids = torch.tensor([seq], dtype=torch.long)
    for crit in stopping:
        r = crit(ids, None)
        if bool(r.all() if isinstance(r, torch.Tensor) else r):
            return True
    return False


def effective_min_length(min_length, prompt_len, semantics):
    """"5": generate() called with inputs_embeds only subtracts the prompt length from max_length AND min_length
    (`_prepare_generated_length`), so the reference's min_length=1 never suppresses EOS; "4.28": MinLengthLogitsProcessor sees
    [dummy BOS ; generated] and the configured min_length."""
    if semantics == "4.28":
        return min_length
    return max((min_length or 0) - prompt_len, 0)


def greedy(logits_fn, max_new_tokens, eos_ids=(), bos_id=1, min_length=0, repetition_penalty=1.0, stopping=None,
           semantics="5", prompt_len=0):
    """batch 1.  Returns the generated ids (without the 4.28 dummy BOS)."""
    v4 = semantics == "4.28"
    min_length = effective_min_length(min_length, prompt_len, semantics)
    seq = []
    for _ in range(max_new_tokens):
        hist_proc = ([bos_id] + seq) if v4 else seq
        logits = logits_fn([seq]).float()
        scores = _apply_processors(logits, [hist_proc], min_length, eos_ids, repetition_penalty)
        tok = int(scores[0].argmax())
        seq = seq + [tok]
        if tok in eos_ids or _custom_stop(stopping, ([bos_id] + seq) if v4 else seq):
            break
    return seq


class _Hyps:
    """generation/beam_search.py BeamHypotheses (4.28): keeps the num_beams best (score, tokens)"""

    def __init__(self, num_beams, length_penalty):
        self.nb, self.lp, self.beams, self.worst = num_beams, length_penalty, [], 1e9

    def add(self, tokens, sum_logprobs):
        score = sum_logprobs / (len(tokens) ** self.lp)
        if len(self.beams) < self.nb or score > self.worst:
            self.beams.append((score, list(tokens)))
            if len(self.beams) > self.nb:
                order = sorted((s, i) for i, (s, _) in enumerate(self.beams))
                del self.beams[order[0][1]]
                self.worst = order[1][0]
            else:
                self.worst = min(score, self.worst)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self.beams) < self.nb:
            return False
        return self.worst >= best_sum_logprobs / (cur_len ** self.lp)      # early_stopping=False (the default)


def _beam_v4(logits_fn, nb, max_new_tokens, eos_ids, bos_id, min_length, rp, lp, stopping):
    max_length = 1 + max_new_tokens                      # dummy BOS + new tokens
    ids = [[bos_id] for _ in range(nb)]
    beam_scores = [0.0] + [-1e9] * (nb - 1)
    hyps = _Hyps(nb, lp)
    done = False
    while True:
        logits = logits_fn([s[1:] for s in ids]).float()
        scores = torch.log_softmax(logits, dim=-1)
        scores = _apply_processors(scores, ids, min_length, eos_ids, rp)
        scores = scores + torch.tensor(beam_scores)[:, None]
        V = scores.shape[1]
        top_s, top_i = scores.reshape(-1).topk(2 * nb)
        nxt = []
        for rank, (s, i) in enumerate(zip(top_s.tolist(), top_i.tolist())):
            b, t = i // V, i % V
            if t in eos_ids:
                if rank >= nb:
                    continue
                hyps.add(ids[b], s)
            else:
                nxt.append((s, t, b))
            if len(nxt) == nb:
                break
        done = done or hyps.is_done(max(top_s.tolist()), len(ids[0]) + 1)
        ids = [ids[b] + [t] for (_, t, b) in nxt]
        beam_scores = [s for (s, _, _) in nxt]
        if done or len(ids[0]) >= max_length or _custom_stop(stopping, ids[0]):
            break
    if not done:
        for b in range(nb):
            hyps.add(ids[b], beam_scores[b])
    best = sorted(hyps.beams, key=lambda x: x[0])[-1][1]
    if len(best) < max_length and eos_ids:
        best = best + [list(eos_ids)[0]]
    return best[1:]                                      # without the dummy BOS


def _beam_v5(logits_fn, nb, max_new_tokens, eos_ids, min_length, rp, lp, stopping):
    max_length = max_new_tokens
    K = max(2, 1 + len(eos_ids)) * nb
    running = [[] for _ in range(nb)]
    running_scores = [0.0] + [-1e9] * (nb - 1)
    fin_seqs, fin_scores, fin_flags = [[] for _ in range(nb)], [-1e9] * nb, [False] * nb
    unsat = True
    cur_len = 0
    while True:
        logits = logits_fn(running).float()
        logp = torch.log_softmax(logits, dim=-1)
        logp = _apply_processors(logp, running, min_length, eos_ids, rp)
        logp = logp + torch.tensor(running_scores, dtype=torch.float32)[:, None]
        V = logp.shape[1]
        top_s, top_i = logp.reshape(-1).topk(K)
        cand_s = top_s.tolist()
        cand_seq = [running[i // V] + [i % V] for i in top_i.tolist()]
        row0_custom = _custom_stop(stopping, cand_seq[0])   # old-style criteria look at row 0 and return one bool for all
        hits = [(s[-1] in eos_ids) or len(s) >= max_length or row0_custom for s in cand_seq]
        # e. running beams of the next iteration: best num_beams candidates that did not finish
        run_s = torch.tensor([s + (-1.0e9 if h else 0.0) for s, h in zip(cand_s, hits)], dtype=torch.float32)
        sel = run_s.topk(nb)[1].tolist()
        running = [cand_seq[i] for i in sel]
        running_scores = [float(run_s[i]) for i in sel]
        # f. finished beams: only the top num_beams candidates may finish
        m_scores, m_seqs, m_flags = list(fin_scores), list(fin_seqs), list(fin_flags)
        for i in range(K):
            just = hits[i] and i < nb
            s = torch.tensor(cand_s[i], dtype=torch.float32) / ((cur_len + 1) ** lp)
            s = float(s) + (0.0 if unsat else -1.0e9) + (0.0 if just else -1.0e9)
            m_scores.append(s)
            m_seqs.append(cand_seq[i])
            m_flags.append(just)
        order = torch.tensor(m_scores, dtype=torch.float32).topk(nb)[1].tolist()
        fin_scores = [m_scores[i] for i in order]
        fin_seqs = [m_seqs[i] for i in order]
        fin_flags = [m_flags[i] for i in order]
        cur_len += 1
        best_possible = running_scores[0] / (cur_len ** lp)
        worst = [min(fin_scores) if f else -1.0e9 for f in fin_flags]
        unsat = unsat and any(best_possible > w for w in worst)
        if not unsat or all(hits):
            break
    return fin_seqs[0]


def beam_search(logits_fn, num_beams, max_new_tokens, eos_ids=(), bos_id=1, min_length=0, repetition_penalty=1.0,
                length_penalty=1.0, stopping=None, semantics="5", prompt_len=0):
    """batch 1, do_sample False, early_stopping False.  Returns the generated ids of the best hypothesis."""
    min_length = effective_min_length(min_length, prompt_len, semantics)
    if semantics == "4.28":
        return _beam_v4(logits_fn, num_beams, max_new_tokens, tuple(eos_ids), bos_id, min_length, repetition_penalty,
                        length_penalty, stopping)
    return _beam_v5(logits_fn, num_beams, max_new_tokens, tuple(eos_ids), min_length, repetition_penalty, length_penalty,
                    stopping)


def llama_logits_fn(sd, prompt_embeds, num_layers, num_heads):
    """next-token logits by FULL RECOMPUTE of [prompt ; embed(history)] with oracle.stllm_oracle.llama_forward (fp32): the
    reference semantics a KV-cached implementation must reproduce.  sd: reference key names ('model.*', 'lm_head.weight')."""
    from oracle import stllm_oracle as O
    w = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
    E, head = w["embed_tokens.weight"], sd["lm_head.weight"]

    def fn(histories):
        out = []
        for h in histories:
            x = prompt_embeds if len(h) == 0 else torch.cat([prompt_embeds, E[torch.tensor(h, dtype=torch.long)][None]], 1)
            att = torch.ones(1, x.shape[1], dtype=torch.long)
            hid = O.llama_forward(w, x, att, num_layers, num_heads)
            out.append(torch.nn.functional.linear(hid[0, -1], head))
        return torch.stack(out)
    return fn
