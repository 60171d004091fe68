"""Autoregressive generation over the stateful `model(x, inference_params_dict=...)` call;
entry points and semantics of the reference's evo/generation.py (Generator.generate
:38-204, generate :207-297): the first `force_prompt_threshold` prompt tokens are prefilled
in one parallel forward, the remaining prompt tokens are teacher-forced one step at a time,
then `num_tokens` tokens are sampled.  Returned scores keep the reference's alignment
(quirk Q3).  Unlike the reference, a prompt of any length can be prefilled in
one pass (pass force_prompt_threshold >= prompt length): the Hyena prefill state comes out
of the scan kernel, there is no (B, D, 8, 2L) FFT temporary."""
from __future__ import annotations

import sys
from typing import List, Tuple

import numpy as np
import torch

from .scoring import logits_to_logprobs, prepare_batch
from .stripedhyena.sample import sample
from .tokenizer import CharLevelTokenizer


class Generator:
    def __init__(self, model, tokenizer: CharLevelTokenizer, top_k: int = 50, top_p: float = 0.7, temperature: float = 1.0):
        self.model = model
        self.tokenizer = tokenizer
        self.top_k = top_k
        self.top_p = top_p
        self.temperature = temperature
        self.untils = ["\n\n"]

    def generate(self, device: str, input_string: str = None, input_ids: torch.Tensor = None, num_tokens: int = 32,
                 cached_generation: bool = True, force_prompt_threshold: int = 128, print_generation: bool = True,
                 verbose: bool = False, skip_special_tokens: bool = False, stop_at_eos: bool = True,
                 max_seqlen: int = None, inference_params_dict: dict = None) -> Tuple[torch.Tensor, torch.Tensor, dict]:
        tok = self.tokenizer
        eos_ids = torch.tensor([tok.eos], dtype=torch.long, device=device)
        if input_ids is None:
            prompt = torch.tensor(tok.tokenize(input_string), dtype=torch.long, device=device).unsqueeze(0)
        else:
            prompt = input_ids
        x = prompt if max_seqlen is None else prompt[:, -max_seqlen:]
        num_tokens = int(num_tokens)
        batch, prompt_len = x.shape

        n_forced = max(0, prompt_len - force_prompt_threshold)
        forced = x[:, force_prompt_threshold:] if n_forced else None
        if n_forced:
            x = x[:, :force_prompt_threshold]

        out_ids = torch.empty(batch, num_tokens, dtype=torch.long, device=x.device)
        out_logits = torch.empty(batch, num_tokens, tok.vocab_size, dtype=torch.float, device=x.device)

        prefilled = False
        if inference_params_dict is not None:      # resume from a caller-held state (evo/generation.py:105-114)
            cached_generation, prefilled = True, True
            mha, hy = inference_params_dict["mha"], inference_params_dict["hyena"]
            for store in (mha.key_value_memory_dict, hy.fir_state_dict, hy.state_dict):
                for k in list(store):
                    store[k] = store[k].to(x.device)
        elif cached_generation:
            inference_params_dict = self.model.initialize_inference_params()
            inference_params_dict["mha"].max_batch_size = batch
            inference_params_dict["hyena"].max_batch_size = batch

        if verbose:
            print(f"Memory after tokenization: {torch.cuda.memory_allocated(device=x.device) / 1e9} GB")
            print("Starting generation...")
            print("Prompt: " + input_string if input_string is not None else f"Prompt ids: {input_ids} {input_ids.shape}")

        total = n_forced + num_tokens
        step = -1
        for step in range(total):
            stepping = prefilled or (cached_generation and step > 0)
            if stepping:
                x = x[:, -1:]
                mha, hy = inference_params_dict["mha"], inference_params_dict["hyena"]
                if mha.seqlen_offset == 0:
                    # the reference jumps to the FULL prompt length here even when only
                    # `force_prompt_threshold` tokens were prefilled (quirk Q1); kept as is
                    mha.seqlen_offset = hy.seqlen_offset = prompt.shape[-1]
                else:
                    mha.seqlen_offset += 1
                    hy.seqlen_offset += 1

            with torch.inference_mode():
                logits, inference_params_dict = self.model(x, inference_params_dict=inference_params_dict)
            last = logits[:, -1]

            if step < n_forced:
                nxt = forced[:, step]
            else:
                nxt = sample(last, top_k=self.top_k, top_p=self.top_p, temperature=self.temperature)

            if stop_at_eos and num_tokens >= 2 and bool((out_ids[0, -2:] == eos_ids).all()):
                print("Stopping generation at EOS")   # the reference only reports it (quirk Q2)
            if print_generation and verbose and batch == 1:
                print(f"{tok.detokenize([nxt.item()])}", end=" ")

            slot = step - n_forced
            if slot >= 0:
                out_logits[:, slot] = last
                out_ids[:, slot] = nxt

            x = nxt[:, None] if stepping else torch.cat([x, nxt[:, None]], dim=-1)

        if verbose:
            text = tok.detokenize_batch(out_ids[:, : step + 1])
            for until in self.untils:
                if until in text:
                    text = text.split(until)[0]
                    break
            print(f"\nInput: {input_string}, Output: {text}")
            print(f"Memory after generation: {torch.cuda.memory_allocated(device=x.device) / 1e9} GB")

        return out_ids[:, : step + 1], out_logits[:, : step + 1], inference_params_dict


def generate(prompt_seqs: List[str], model, tokenizer: CharLevelTokenizer, n_tokens: int = 100, temperature: float = 0.0,
             top_k: int = 1, top_p: float = 1.0, batched: bool = True, prepend_bos: bool = False,
             cached_generation: bool = False, force_prompt_threshold: int = 128, verbose: int = 1,
             device: str = "cuda:0", **kwargs) -> Tuple[List[str], List[float]]:
    """Generate from a list of prompts; equal-length prompts are batched."""
    model.eval()
    g = Generator(model, tokenizer, top_k=top_k, top_p=top_p, temperature=temperature)
    same_len = all(len(s) == len(prompt_seqs[0]) for s in prompt_seqs)
    if batched and same_len:
        groups = [prompt_seqs]
    else:
        if verbose:
            if not same_len:
                sys.stderr.write("Note: Prompts are of different lengths.\n")
            sys.stderr.write("Note: Will not do batched generation.\n")
        groups = [[s] for s in prompt_seqs]

    seqs_out: List[str] = []
    scores_out: List[float] = []
    for grp in groups:
        ids = prepare_batch(grp, tokenizer, prepend_bos=prepend_bos, device=device)[0]
        out_ids, logits, _ = g.generate(input_ids=ids, num_tokens=n_tokens, cached_generation=cached_generation,
                                        force_prompt_threshold=force_prompt_threshold, device=device,
                                        print_generation=(verbose > 1), verbose=(verbose > 1), stop_at_eos=False)
        if verbose > 1:
            print("input_ids.shape", ids.shape)
            print("output_ids.shape", out_ids.shape)
            print("logits.shape", logits.shape)
        texts = tokenizer.detokenize_batch(out_ids)
        if len(texts) != ids.shape[0]:
            raise AssertionError("batch size mismatch after detokenisation")
        seqs_out += list(texts)
        lp = logits_to_logprobs(logits, out_ids).float().cpu().numpy()   # alignment as in the reference (Q3)
        scores_out += [float(np.mean(lp[i])) for i in range(ids.shape[0])]

    if verbose:
        for seq, score, prompt in zip(seqs_out, scores_out, prompt_seqs):
            print(f'Prompt: "{prompt}",\tOutput: "{seq}",\tScore: {score}')
    return seqs_out, scores_out
