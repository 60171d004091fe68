"""
Usage: python -m scripts.example_inference [--model-name evo-1-131k-base] [--random-init]

Logits for one sequence and for a padded batch -- the two calls of the reference's scripts/example_inference.py
(`model(input_ids)` with int32 ids, and `prepare_batch(..., prepend_bos=False)`), on the evo_b200 engine.
"""
import argparse

import torch

from evo_b200 import Evo
from evo_b200.scoring import prepare_batch


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--model-name', type=str, default='evo-1-131k-base')
    parser.add_argument('--device', type=str, default='cuda:0')
    parser.add_argument('--random-init', action='store_true')
    args = parser.parse_args(argv)
    evo_model = Evo(args.model_name, device=args.device, random_init=args.random_init)
    model, tokenizer = evo_model.model, evo_model.tokenizer
    model.eval()

    input_ids = torch.tensor(tokenizer.tokenize('ACGT'), dtype=torch.int).to(args.device).unsqueeze(0)
    logits, _ = model(input_ids)                      # (batch, length, vocab)
    print('Logits: ', logits)
    print('Shape (batch, length, vocab): ', logits.shape)

    input_ids, seq_lengths = prepare_batch(['ACGT', 'A', 'AAAAACCCCCGGGGGTTTTT'], tokenizer, prepend_bos=False, device=args.device)
    logits, _ = model(input_ids)
    print('Batch logits: ', logits)
    print('Batch shape (batch, length, vocab): ', logits.shape)
    return logits


if __name__ == '__main__':
    main()
