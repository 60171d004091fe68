"""
Usage: python -m scripts.generate \
           --model-name evo-1-131k-base \
           --prompt ACGT \
           --n-samples 10 \
           --n-tokens 100 \
           --temperature 1. \
           --top-k 4 \
           --device cuda:0

Generates sequences from a prompt with the reference's sampling arguments (scripts/generate.py of evo-design/evo), on the
evo_b200 engine.  With --cached-generation (the default) the token loop runs on the GPU (one CUDA graph per token, device
sampler); --seed makes a sampled run reproducible.  The boolean flags keep the reference's `type=bool` parsing (any non-empty
value is True).
"""
import argparse

import torch

from evo_b200 import Evo, generate


def build_parser():
    parser = argparse.ArgumentParser(description='Generate sequences with an Evo model on the evo_b200 engine.')
    parser.add_argument('--model-name', type=str, default='evo-1-131k-base', help='Evo model name')
    parser.add_argument('--prompt', type=str, default='ACGT', help='Prompt for generation')
    parser.add_argument('--n-samples', type=int, default=3, help='Number of sequences to sample at once')
    parser.add_argument('--n-tokens', type=int, default=100, help='Number of tokens to generate')
    parser.add_argument('--temperature', type=float, default=1.0, help='Temperature during sampling')
    parser.add_argument('--top-k', type=int, default=4, help='Top K during sampling')
    parser.add_argument('--top-p', type=float, default=1., help='Top P during sampling')
    parser.add_argument('--cached-generation', type=bool, default=True, help='Use KV caching during generation')
    parser.add_argument('--batched', type=bool, default=True, help='Use batched generation')
    parser.add_argument('--prepend-bos', type=bool, default=False, help='Prepend BOS token')
    parser.add_argument('--device', type=str, default='cuda:0', help='Device for generation')
    parser.add_argument('--verbose', type=int, default=1, help='Verbosity level')
    parser.add_argument('--seed', type=int, default=None, help='torch.manual_seed before sampling')
    parser.add_argument('--random-init', action='store_true', help='Skip the checkpoint download (random weights; smoke runs offline)')
    parser.add_argument('--model-dir', type=str, default=None, help='Directory of an already downloaded HF snapshot')
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    evo_model = Evo(args.model_name, device=args.device, random_init=args.random_init, model_dir=args.model_dir)
    model, tokenizer = evo_model.model, evo_model.tokenizer
    model.eval()
    if args.seed is not None:
        torch.manual_seed(args.seed)
    print('Generated sequences:')
    return generate([args.prompt] * args.n_samples, model, tokenizer, n_tokens=args.n_tokens, temperature=args.temperature,
                    top_k=args.top_k, top_p=args.top_p, cached_generation=args.cached_generation, batched=args.batched,
                    prepend_bos=args.prepend_bos, device=args.device, verbose=args.verbose)


if __name__ == '__main__':
    main()
