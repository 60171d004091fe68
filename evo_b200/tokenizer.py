"""Byte-level tokenizer (the reference's evo/tokenizer.py contract): id = byte value,
eod/eos = 0, pad = 1, vocabulary 512; detokenisation clamps ids into [32, vocab]."""
from __future__ import annotations

from typing import List, Sequence, Union

import numpy as np
import torch


class CharLevelTokenizer:
    name = "CharLevelTokenizer"
    eod_id = 0
    eos_id = 0
    pad_id = 1

    def __init__(self, vocab_size: int = 512):
        self._vocab_size = int(vocab_size)

    vocab_size = property(lambda self: self._vocab_size)
    eod = property(lambda self: self.eod_id)
    eos = property(lambda self: self.eod_id)

    # text -> ids
    def tokenize(self, text: str) -> List[int]:
        return np.frombuffer(text.encode(), dtype=np.uint8).tolist()

    def tokenize_array(self, text: str) -> np.ndarray:
        """uint8 view of the encoded text (no Python-int list), for bulk batching."""
        return np.frombuffer(text.encode(), dtype=np.uint8)

    def tokenize_batch(self, text_batch: Union[Sequence[str], str]):
        if isinstance(text_batch, str):
            return self.tokenize(text_batch)
        return [self.tokenize(t) for t in text_batch]

    # ids -> text
    def clamp(self, n: int) -> int:
        return max(32, min(int(n), self._vocab_size))

    def decode_token(self, token: int) -> str:
        return chr(self.clamp(token))

    def detokenize(self, token_ids) -> str:
        return "".join(self.decode_token(t) for t in token_ids)

    def detokenize_batch(self, token_ids):
        if isinstance(token_ids, torch.Tensor):
            token_ids = token_ids.tolist()
        if isinstance(token_ids, list) and token_ids and isinstance(token_ids[0], (list, tuple)):
            return [self.detokenize(row) for row in token_ids]
        if isinstance(token_ids, list) and not token_ids:
            return []
        return self.detokenize(token_ids)
