"""Deterministic offline stand-in for the HF tokenizers the reference constructs from downloaded files
(LlamaTokenizer st_llm.py:305, BertTokenizer blip2.py:32).  Words -> ids by a stable hash; implements exactly the
call protocol STLLMModel uses (st_llm.py:345-351, 386-389, 501-508)."""
import zlib

import torch


class _Enc:
    def __init__(self, input_ids, attention_mask):
        self.input_ids = input_ids
        self.attention_mask = attention_mask

    def to(self, device):
        return _Enc(self.input_ids.to(device), self.attention_mask.to(device))


class FakeTokenizer:
    def __init__(self, vocab_size, pad_token_id=0, bos_token_id=1, eos_token_id=2, eos_token="</s>"):
        self.vocab_size = vocab_size
        self.pad_token_id = pad_token_id
        self.bos_token_id = bos_token_id
        self.eos_token_id = eos_token_id
        self.eos_token = eos_token
        self.padding_side = "right"

    def __len__(self):
        return self.vocab_size

    def _encode(self, text, add_special_tokens):
        ids = []
        if add_special_tokens:
            ids.append(self.bos_token_id)
        for w in text.replace(self.eos_token, " " + self.eos_token + " ").split():
            if w == self.eos_token:
                ids.append(self.eos_token_id)
            else:
                ids.append(3 + zlib.crc32(w.encode()) % (self.vocab_size - 3))
        return ids

    def decode(self, ids, add_special_tokens=False, skip_special_tokens=False):
        """ids -> "w<id>" words (no inverse of the hash exists; callers only split / strip the text)"""
        return " ".join(f"w{int(i)}" for i in (ids.tolist() if hasattr(ids, "tolist") else ids))

    def __call__(self, text, return_tensors="pt", padding=False, truncation=False, max_length=None,
                 add_special_tokens=True):
        single = isinstance(text, str)
        texts = [text] if single else list(text)
        seqs = [self._encode(t, add_special_tokens) for t in texts]
        if truncation and max_length is not None:
            seqs = [s[:max_length] for s in seqs]
        L = max((len(s) for s in seqs), default=0)
        ids = torch.full((len(seqs), L), self.pad_token_id, dtype=torch.long)
        att = torch.zeros((len(seqs), L), dtype=torch.long)
        for i, s in enumerate(seqs):
            if self.padding_side == "right":
                ids[i, :len(s)] = torch.tensor(s, dtype=torch.long)
                att[i, :len(s)] = 1
            else:
                ids[i, L - len(s):] = torch.tensor(s, dtype=torch.long)
                att[i, L - len(s):] = 1
        return _Enc(ids, att)
