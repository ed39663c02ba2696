"""Sequence assembly for `STLLMModel.forward(samples)`: the LLM input of a batch is described by ONE integer row-index
matrix built on the host, and materialised by ONE gather launch per LLM pass.

What the reference does (st_llm.py:379-432, 496-546): per sample it tokenises the text before / after `<ImageHere>`,
embeds both, concatenates [before | video | after], right-pads the batch with embed(pad), splices the (padded) answer
embeddings in after each sample's valid part, optionally prepends BOS, and builds attention mask and labels -- a
chain of small torch.cat / slice-assign kernels per sample.  Here the same layout is computed as index arithmetic:

    row code  >= 0 : row `code` of the token-embedding table
    row code  <  0 : row `-1 - code` of the flattened video-token matrix [B * L, D]

so a batch costs: tokeniser calls (one per distinct prompt), a few vectorised numpy ops on [B, S] int64 arrays, one
host->device copy of the packed plan and one `ops.gather_rows2` launch.  The masked pass of the MVM branch indexes the
kept video rows directly (no separate compaction of the video tokens).
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np


@dataclass
class SequencePlan:
    rows: np.ndarray            # int64 [B, S] row codes (see module docstring)
    attention_mask: np.ndarray  # int64 [B, S]
    targets: np.ndarray         # int64 [B, S], -100 outside the answer tokens
    input_lens: np.ndarray      # int64 [B] valid length of [before | video | after] per sample (st_llm.py:413-415)
    wrapped_len: int            # width of the padded prompt block (`max(emb_lens)`, st_llm.py:398-400)


def split_prompt(prompt: str):
    """`p_before, p_after = each_prompt.split('<ImageHere>')` (st_llm.py:386) -- exactly one placeholder"""
    parts = prompt.split("<ImageHere>")
    if len(parts) != 2:
        raise ValueError("too many values to unpack" if len(parts) > 2 else "not enough values to unpack")
    return parts[0], parts[1]


def video_row_codes(batch: int, length: int, keep: Optional[np.ndarray] = None) -> List[np.ndarray]:
    """row codes of each sample's video tokens in the flattened [batch * length, D] matrix; `keep` (bool [batch, length])
    selects the visible tokens of the masked pass in row-major order, the order of `img_embeds[~mask]` (st_llm.py:491)"""
    base = np.arange(batch, dtype=np.int64)[:, None] * length + np.arange(length, dtype=np.int64)[None, :]
    codes = -1 - base
    if keep is None:
        return [codes[i] for i in range(batch)]
    keep = np.asarray(keep, dtype=bool).reshape(batch, length)
    return [codes[i][keep[i]] for i in range(batch)]


def plan_sequences(video_codes: Sequence[np.ndarray], before_ids: Sequence[np.ndarray], after_ids: Sequence[np.ndarray],
                   answer_ids: Optional[np.ndarray], answer_mask: Optional[np.ndarray], pad_id: int,
                   bos_id: Optional[int] = None) -> SequencePlan:
    """Layout per sample i (n_i = len(before_i) + len(video_i) + len(after_i), W = max n_i, La = answer width):

        [BOS]? | before_i | video_i | after_i | answer_i (La, incl. its own padding) | pad x (W - n_i)

    attention mask: 1 on BOS and the n_i input positions, the tokeniser's mask on the answer, 0 on the padding;
    targets: the answer ids (pad -> -100) at [off + n_i, off + n_i + La), -100 elsewhere (st_llm.py:532-542)."""
    B = len(video_codes)
    assert len(before_ids) == B and len(after_ids) == B
    pieces = [np.concatenate([np.asarray(before_ids[i], dtype=np.int64).reshape(-1),
                              np.asarray(video_codes[i], dtype=np.int64).reshape(-1),
                              np.asarray(after_ids[i], dtype=np.int64).reshape(-1)]) for i in range(B)]
    n = np.array([p.size for p in pieces], dtype=np.int64)
    W = int(n.max()) if B else 0
    La = 0 if answer_ids is None else int(answer_ids.shape[1])
    off = 0 if bos_id is None else 1
    S = off + W + La
    pos = np.arange(S, dtype=np.int64)[None, :]
    start = (off + n)[:, None]
    in_input = (pos >= off) & (pos < start)
    in_answer = (pos >= start) & (pos < start + La)
    rows = np.full((B, S), int(pad_id), dtype=np.int64)
    att = np.zeros((B, S), dtype=np.int64)
    targets = np.full((B, S), -100, dtype=np.int64)
    rows[in_input] = np.concatenate(pieces) if B else np.zeros(0, dtype=np.int64)   # boolean-mask order is row-major
    att[in_input] = 1
    if La:
        a_ids = np.asarray(answer_ids, dtype=np.int64)
        rows[in_answer] = a_ids.reshape(-1)
        att[in_answer] = np.asarray(answer_mask, dtype=np.int64).reshape(-1)
        targets[in_answer] = np.where(a_ids == pad_id, -100, a_ids).reshape(-1)
    if bos_id is not None:
        rows[:, 0] = int(bos_id)
        att[:, 0] = 1
    return SequencePlan(rows=rows, attention_mask=att, targets=targets, input_lens=n, wrapped_len=W)


def splice_rows(input_mask: np.ndarray, answer_len: int) -> np.ndarray:
    """Row codes of `concat_emb_input_output` (st_llm.py:409-432) for already-embedded inputs: code >= 0 -> row of the
    flattened input block [B * W], code < 0 -> row `-1 - code` of the flattened answer block [B * La]."""
    input_mask = np.asarray(input_mask)
    B, W = input_mask.shape
    n = input_mask.sum(axis=1).astype(np.int64)
    S = W + answer_len
    pos = np.arange(S, dtype=np.int64)[None, :]
    before = pos < n[:, None]
    in_answer = (pos >= n[:, None]) & (pos < n[:, None] + answer_len)
    src_in = np.where(before, pos, pos - answer_len) + np.arange(B, dtype=np.int64)[:, None] * W
    src_ans = -1 - ((pos - n[:, None]) + np.arange(B, dtype=np.int64)[:, None] * answer_len)
    return np.where(in_answer, src_ans, src_in).astype(np.int64)
