"""Multi-GPU check of evo_b200.parallel.sequence_parallel_forward (run under torchrun):
every rank computes the unsharded forward (small model) and its shard of the sequence-
parallel forward; the shards must match the unsharded logits.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/harness/seqpar_check.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from evo_b200.parallel import sequence_parallel_forward
from evo_b200.stripedhyena import StripedHyena, dotdict
from oracle import stripedhyena_oracle as O


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = f"cuda:{local}"
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device(dev))
    out = {}
    for name, cfg, B, L in (("tiny4", O.tiny_config(num_layers=4, attn_layer_idxs=(1, 3), hidden_size=128 * max(2, world), num_heads=max(2, world)), 2, 1024 * world),
                            ("wide2_131k_rope", dict(O.evo_config("evo-1-131k-base"), num_layers=2, attn_layer_idxs=[1], hyena_layer_idxs=[0]), 1, 2048 * world)):
        sd = O.random_state_dict(cfg, seed=5)
        m = StripedHyena(dotdict(cfg))
        m.load_state_dict(sd, strict=True)
        m.to_bfloat16_except_poles_residues()
        m = m.to(dev)
        g = torch.Generator().manual_seed(1)
        ids = (torch.randint(0, 4, (B, L), generator=g) * 3 + 65).to(dev)
        full, _ = m(ids)
        Lr = L // world
        ref = full[:, rank * Lr:(rank + 1) * Lr]
        for transport in ("nccl", "peer"):
            try:
                mine = sequence_parallel_forward(m, ids[:, rank * Lr:(rank + 1) * Lr].contiguous(), rank, world, transport=transport)
                mine2 = sequence_parallel_forward(m, ids[:, rank * Lr:(rank + 1) * Lr].contiguous(), rank, world, transport=transport)   # ring reuse
                d = torch.maximum((mine.float() - ref.float()).abs(), (mine2.float() - ref.float()).abs())
                stats = torch.tensor([d.max().item(), d.mean().item(), ref.float().abs().max().item(),
                                      (mine.argmax(-1) == ref.argmax(-1)).float().mean().item()], device=dev, dtype=torch.float64)
            except Exception as ex:  # noqa
                print(f"[rank {rank}] transport {transport} failed: {ex!r}", flush=True)
                stats = torch.tensor([float("nan")] * 4, device=dev, dtype=torch.float64)
            allst = [torch.empty_like(stats) for _ in range(world)]
            dist.all_gather(allst, stats)
            out[f"{name}/{transport}"] = [[float(v) for v in s.tolist()] for s in allst]
    if rank == 0:
        print(json.dumps({"seqpar_check": out, "world": world, "cols": ["max_abs", "mean_abs", "ref_max", "argmax_agree"]}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
