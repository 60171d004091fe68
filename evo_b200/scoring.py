"""Sequence scoring on top of `model(input_ids)`; same entry points and semantics as the
reference's evo/scoring.py (prepare_batch :9-33, logits_to_logprobs :36-59,
score_sequences :62-96, positional_entropies :99-131).

Differences, all inside the contract: the sequences' bytes move with ONE pinned host->device
copy of 1 byte per nucleotide and the id matrix is built on the GPU (the reference does one int64
copy per sequence, :22-30); score_sequences / positional_entropies read log-likelihood and entropy
from the fused scoring head (model.score_tokens: unembed GEMM + log-softmax + gather in one kernel,
the (B, L, 512) logits are never written; fp32 statistics instead of the reference's bf16
log_softmax, quirk Q4 in SURVEY.md 8c)."""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np
import torch

from . import _lib
from .tokenizer import CharLevelTokenizer


def prepare_batch(seqs: List[str], tokenizer: CharLevelTokenizer, prepend_bos: bool = True,
                  device: str = "cuda:0") -> Tuple[torch.Tensor, List[int]]:
    """(B, bos + max_len) int64 ids, right-padded with pad_id, BOS = eod_id.
    On a CUDA device the bytes go over once as uint8 and the id matrix is written there (frontend.device_batch)."""
    if torch.device(device).type == "cuda" and torch.cuda.is_available():
        from .frontend import device_batch
        return device_batch(seqs, tokenizer, prepend_bos=prepend_bos, device=device)
    lengths = [len(s) for s in seqs]
    width = max(lengths) + int(prepend_bos)
    host = torch.full((len(seqs), width), tokenizer.pad_id, dtype=torch.long)
    if torch.cuda.is_available():
        host = host.pin_memory()
    arr = host.numpy()
    first = int(prepend_bos)
    if prepend_bos:
        arr[:, 0] = tokenizer.eod_id
    for row, s in enumerate(seqs):
        ids = np.frombuffer(s.encode(), dtype=np.uint8)
        arr[row, first:first + ids.shape[0]] = ids
    return host.to(device, non_blocking=True), lengths


def logits_to_logprobs(logits: torch.Tensor, input_ids: torch.Tensor, trim_bos: bool = True) -> torch.Tensor:
    """log p(token_t | prefix) gathered at the observed tokens; (B, L) or (B, L-1) when trim_bos.
    Output dtype follows `logits` like the reference (log_softmax in the logits' dtype)."""
    lp = torch.log_softmax(logits, dim=-1)
    if trim_bos:
        lp, input_ids = lp[:, :-1], input_ids[:, 1:]
    if lp.shape[1] != input_ids.shape[1]:
        raise AssertionError("logits and input_ids lengths differ")
    return torch.gather(lp, 2, input_ids.unsqueeze(-1).long()).squeeze(-1)


def _fused_logprobs(logits: torch.Tensor, input_ids: torch.Tensor) -> torch.Tensor:
    """evo_logprobs: out[b, t] = log_softmax(logits[b, t])[ids[b, t+1]] (0 at the last position)."""
    B, L, V = logits.shape
    targets = torch.full((B, L), -1, dtype=torch.long, device=logits.device)
    targets[:, :-1] = input_ids[:, 1:]
    out = torch.empty(B, L, dtype=torch.float32, device=logits.device)
    _lib.check(_lib.lib().evo_logprobs(_lib.ptr(logits), _lib.ptr(targets), _lib.ptr(out), B * L, V,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "evo_logprobs")
    return out[:, :-1]


def score_sequences(seqs: List[str], model, tokenizer: CharLevelTokenizer, reduce_method: str = "mean",
                    device: str = "cuda:0") -> List[float]:
    if reduce_method not in ("mean", "sum"):
        raise ValueError(f"Invalid reduce_method {reduce_method}")
    input_ids, lengths = prepare_batch(seqs, tokenizer, device=device, prepend_bos=True)
    if hasattr(model, "score_tokens") and input_ids.is_cuda:
        # fused head: unembed + log-softmax + gather in the GEMM epilogue, logits never written
        logprobs = model.score_tokens(input_ids, want_logprobs=True)[0][:, :-1]
    else:
        with torch.inference_mode():
            logits, _ = model(input_ids)
        if logits.is_cuda and logits.dtype == torch.bfloat16 and logits.is_contiguous():
            logprobs = _fused_logprobs(logits, input_ids)
        else:
            logprobs = logits_to_logprobs(logits, input_ids, trim_bos=True).float()
    logprobs = logprobs.cpu().numpy()
    reduce = np.mean if reduce_method == "mean" else np.sum
    return [reduce(logprobs[i][:n]) for i, n in enumerate(lengths)]


def positional_entropies(seqs: List[str], model, tokenizer: CharLevelTokenizer, device: str = "cuda:0") -> List[np.ndarray]:
    input_ids, lengths = prepare_batch(seqs, tokenizer, device=device, prepend_bos=True)
    if hasattr(model, "score_tokens") and input_ids.is_cuda:
        ent = model.score_tokens(input_ids, want_logprobs=False, want_entropy=True)[1][:, :-1].cpu().numpy()
    else:
        with torch.inference_mode():
            logits, _ = model(input_ids)
        lp = torch.log_softmax(logits.float(), dim=-1)[:, :-1]
        ent = -(lp.exp() * lp).sum(dim=-1).cpu().numpy()
    out = [ent[i][:n] for i, n in enumerate(lengths)]
    if any(len(s) != len(e) for s, e in zip(seqs, out)):
        raise AssertionError("entropy length mismatch")
    return out
