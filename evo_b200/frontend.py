"""Batch front-end: sequences on disk -> padded id matrices on the GPU, in length buckets.

What it replaces in the reference:
  * scripts/score.py:40-55       Bio.SeqIO FASTA parse + fixed-size batches in file order (one pad-to-longest per batch)
  * evo/scoring.py:9-33          prepare_batch: Python-list tokenisation, one int64 tensor and one H2D copy PER SEQUENCE
  * semantic_design.py:82-100    read_prompts: group prompts of identical length into batches of <= batch_size
Here: a FASTA / CSV reader with no dependencies, `length_buckets` (identical-length groups like read_prompts, or a
sorted token-budget packing that bounds the padding waste), and `device_batch`: the sequences' bytes travel to the GPU
ONCE as uint8 (1 byte per nucleotide, pinned staging) and evo_tokenize_pad writes the (B, width) id matrix there."""
from __future__ import annotations

import csv
import ctypes as C
from typing import Dict, Iterable, Iterator, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .tokenizer import CharLevelTokenizer


# ----------------------------------------------------------------------------- readers
def read_fasta(path: str) -> Tuple[List[str], List[str]]:
    """(names, sequences) of a FASTA file; sequence lines are joined, whitespace dropped (what str(record.seq) of
    Bio.SeqIO gives, scripts/score.py:40)."""
    names, seqs, cur = [], [], None
    with open(path, "r", encoding="utf-8-sig") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if cur is not None:
                    seqs.append("".join(cur))
                names.append(line[1:].split()[0] if len(line) > 1 else "")
                cur = []
            elif cur is not None:
                cur.append(line)
            else:
                raise ValueError(f"{path}: sequence data before the first '>' header")
    if cur is not None:
        seqs.append("".join(cur))
    return names, seqs


def read_prompts_csv(path: str) -> List[str]:
    """First column of a CSV with a header row (semantic_design.py:72-75)."""
    with open(path, encoding="utf-8-sig", newline="") as f:
        reader = csv.reader(f)
        next(reader)
        return [row[0] for row in reader]


# ----------------------------------------------------------------------------- bucketing
def length_buckets(seqs: Sequence[str], batch_size: int = 32, mode: str = "exact", max_tokens: int = 0,
                   max_waste: float = 0.1) -> List[List[int]]:
    """Index batches.
    mode "exact":  sequences of identical length share a batch of <= batch_size (read_prompts' rule, in first-seen order
                   of the lengths; a length seen once is its own batch) -- zero padding, what batched generation needs.
    mode "sorted": sort by length, then cut a batch when it would exceed batch_size sequences, `max_tokens` padded tokens
                   (0 = no budget) or a padding waste above `max_waste` -- what scoring wants: bounded padding FLOPs and
                   a bounded activation footprint whatever the length mix."""
    if batch_size < 1:
        raise ValueError("batch_size must be >= 1")
    if mode == "exact":
        groups: Dict[int, List[int]] = {}
        for i, s in enumerate(seqs):
            groups.setdefault(len(s), []).append(i)
        out = []
        for idx in groups.values():
            out.extend(idx[k:k + batch_size] for k in range(0, len(idx), batch_size))
        return out
    if mode != "sorted":
        raise ValueError(f"unknown bucketing mode {mode!r}")
    order = sorted(range(len(seqs)), key=lambda i: (len(seqs[i]), i))
    out, cur, cur_tokens = [], [], 0
    for i in order:
        n = len(seqs[i])                                      # ascending, so n is the width of the batch if i joins it
        width_tokens = n * (len(cur) + 1)
        waste = 1.0 - (cur_tokens + n) / max(width_tokens, 1)
        if cur and (len(cur) >= batch_size or (max_tokens and width_tokens > max_tokens) or waste > max_waste):
            out.append(cur)
            cur, cur_tokens = [], 0
        cur.append(i)
        cur_tokens += n
    if cur:
        out.append(cur)
    return out


# ----------------------------------------------------------------------------- host bytes -> device ids
def device_batch(seqs: Sequence[str], tokenizer: CharLevelTokenizer, prepend_bos: bool = True, device: str = "cuda:0",
                 dtype: torch.dtype = torch.long) -> Tuple[torch.Tensor, List[int]]:
    """(B, bos + max_len) ids on `device`, right-padded with pad_id, BOS = eod_id -- prepare_batch's result -- from ONE
    pinned uint8 transfer (the bytes) plus B+1 offsets; the id matrix is written by evo_tokenize_pad on the GPU."""
    lengths = [len(s) for s in seqs]
    raw = [s.encode() for s in seqs]
    blens = [len(r) for r in raw]
    if blens != lengths:
        raise ValueError("non-ASCII characters: the byte-level tokenizer would not give one token per character")
    total = sum(blens)
    stage = torch.empty(max(total, 1) + 8 * (len(seqs) + 1), dtype=torch.uint8).pin_memory()
    view = stage.numpy()
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(blens, out=offs[1:])
    obytes = 8 * (len(seqs) + 1)
    view[:obytes] = offs.view(np.uint8)                       # offsets first (8-byte aligned), bytes behind them
    pos = obytes
    for r in raw:
        view[pos:pos + len(r)] = np.frombuffer(r, dtype=np.uint8)
        pos += len(r)
    dev = torch.device(device)
    on_dev = stage.to(dev, non_blocking=True)
    width = (max(lengths) if lengths else 0) + int(prepend_bos)
    ids = torch.empty(len(seqs), width, dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().evo_tokenize_pad(C.c_void_p(on_dev.data_ptr() + obytes), C.c_void_p(on_dev.data_ptr()), _lib.ptr(ids), int(dtype == torch.long),
                                               len(seqs), width, int(prepend_bos), tokenizer.eod_id, tokenizer.pad_id,
                                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "evo_tokenize_pad")
    return ids, lengths


def score_many(seqs: Sequence[str], model, tokenizer: CharLevelTokenizer, batch_size: int = 32, max_tokens: int = 1 << 17,
               reduce_method: str = "mean", device: str = "cuda:0") -> List[float]:
    """score_sequences over an arbitrary list in length buckets; scores come back in input order (the loop of
    scripts/score.py:45-55 with bounded padding)."""
    from .scoring import score_sequences
    scores: List[float] = [0.0] * len(seqs)
    for idx in length_buckets(seqs, batch_size=batch_size, mode="sorted", max_tokens=max_tokens):
        got = score_sequences([seqs[i] for i in idx], model, tokenizer, reduce_method=reduce_method, device=device)
        for i, g in zip(idx, got):
            scores[i] = float(g)
    return scores
