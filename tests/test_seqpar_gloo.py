"""N > 1 path on CPU: two gloo ranks run the sequence-sharded Hyena algebra of
evo_b200/parallel.py (halo exchange + zero-start end states -> all-gather -> fold with
p^(shard length) -> output scan from the folded state) with the oracle's time-domain operator
and must reproduce the unsharded result; plus the head<->sequence all-to-all (Ulysses) form
of causal attention."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import stripedhyena_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D, H, B, L = 128, 1, 2, 96
        cfg = O.tiny_config(num_layers=1, attn_layer_idxs=(), hidden_size=D, num_heads=H)
        sd = O.random_state_dict(cfg, seed=4)
        pre = "blocks.0.filter."
        w, bb, Dk = sd[pre + "short_filter_weight"].double(), sd[pre + "short_filter_bias"].double(), sd[pre + "D"].double()
        poles, res = sd[pre + "poles"], sd[pre + "residues"]
        torch.manual_seed(0)                                  # same data on every rank
        z = torch.randn(B, L, 3 * D, dtype=torch.float64)
        y_full, st_full = O.hyena_operator_time_domain(z, w, bb, Dk, poles, res, H, D)
        Lr = L // world
        zl = z[:, rank * Lr:(rank + 1) * Lr].contiguous()
        # halo: every rank publishes its last two rows
        tails = [torch.empty(B, 2, 3 * D, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(tails, zl[:, -2:].contiguous())
        halo = tails[rank - 1] if rank > 0 else None
        # zero-start end state of the local shard
        _, end = O.hyena_operator_time_domain(zl, w, bb, Dk, poles, res, H, D, halo=halo)
        ends = [torch.empty_like(torch.view_as_real(end)) for _ in range(world)]
        dist.all_gather(ends, torch.view_as_real(end).contiguous())
        p = torch.view_as_complex(poles.double())[..., 0]
        s_in = torch.zeros_like(end)
        for q_ in range(rank):                                 # S_in = sum_{q<r} p^{(r-1-q) Lr} E_q
            s_in = (p ** Lr)[None] * s_in + torch.view_as_complex(ends[q_])
        y_loc, st_loc = O.hyena_operator_time_domain(zl, w, bb, Dk, poles, res, H, D, halo=halo, state_in=s_in)
        err = (y_loc - y_full[:, rank * Lr:(rank + 1) * Lr]).abs().max().item()
        st_err = (st_loc - st_full).abs().max().item() if rank == world - 1 else 0.0

        # attention: head <-> sequence re-shard (Ulysses), as evo_b200/parallel.py does it
        Hh, dh = 2 * world, 16
        qkv = torch.randn(B, L, 3, Hh, dh, dtype=torch.float64)
        ref = O.causal_attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2])
        Hl = Hh // world
        loc = qkv[:, rank * Lr:(rank + 1) * Lr]                                        # (B, Lr, 3, H, d)
        send = loc.reshape(B, Lr, 3, world, Hl, dh).permute(3, 0, 1, 2, 4, 5).contiguous()
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        full = recv.permute(1, 0, 2, 3, 4, 5).reshape(B, L, 3, Hl, dh)                 # my heads, whole sequence
        o = O.causal_attention(full[:, :, 0], full[:, :, 1], full[:, :, 2])            # (B, L, Hl, d)
        send2 = o.reshape(B, world, Lr, Hl * dh).permute(1, 0, 2, 3).contiguous()
        recv2 = torch.empty_like(send2)
        dist.all_to_all_single(recv2, send2)
        mine = recv2.permute(1, 2, 0, 3).reshape(B, Lr, Hh, dh)
        a_err = (mine - ref[:, rank * Lr:(rank + 1) * Lr]).abs().max().item()
        q.put((rank, err, st_err, a_err))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sequence_sharded_algebra_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, err, st_err, a_err in results:
        assert err < 1e-9, (rank, err)
        assert st_err < 1e-9 and a_err < 1e-12


def test_time_domain_operator_equals_fft_operator():
    D, H = 128, 1
    cfg = O.tiny_config(num_layers=1, attn_layer_idxs=(), hidden_size=D, num_heads=H)
    sd = O.random_state_dict(cfg, seed=4)
    m = O.OracleStripedHyena(cfg, sd, torch.float64)
    z = torch.randn(2, 50, 3 * D, dtype=torch.float64)
    ip = m.initialize_inference_params()["hyena"]
    y_fft = m.hyena_operator(0, z, ip)
    pre = "blocks.0.filter."
    y_td, st = O.hyena_operator_time_domain(z, sd[pre + "short_filter_weight"].double(), sd[pre + "short_filter_bias"].double(),
                                            sd[pre + "D"].double(), sd[pre + "poles"], sd[pre + "residues"], H, D)
    assert (y_fft - y_td).abs().max() < 1e-9
    assert (ip.state_dict[0] - st).abs().max() < 1e-9
