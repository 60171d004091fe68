"""Bridge to the REAL `stripedhyena` package, for the day it is importable next to this repo.

TEST INFRASTRUCTURE ONLY (same rule as stripedhyena_oracle.py: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this).  stripedhyena==0.2.2 (/root/reference/requirements.txt:1) holds the arithmetic of the reference path and could be
obtained neither in the build container nor on the GPU boxes (profiles/r02_pin_attempt_call1.log), so NOTHING in this file has run
against the real package: it is written from the package's published interface -- `stripedhyena.model.StripedHyena(config)`,
`.load_state_dict(sd, strict=True)`, `.to_bfloat16_except_poles_residues()`, `model(ids, inference_params_dict=...)`,
`.initialize_inference_params()`, `stripedhyena.utils.dotdict` (the call sites evo/models.py:141-150, evo/scoring.py:81,
evo/generation.py:117,152 pin exactly these) -- and exercised here only against a stand-in package (tests/test_oracle.py).
What it is for:
  * `available()`      is the package importable?
  * `build(...)`       the real model on CPU: flash kernels off (use_flash_attn / use_flash_rmsnorm / use_flash_depthwise /
                       use_flashfft False -- the torch branches), and the one statement of flash_attn's rotary that is a Triton
                       kernel replaced by flash_attn's own torch statement of it (SURVEY.md 8c names both patches);
  * `compare(...)`     SURVEY.md A.9's verify-first checklist: same weights into the real model and into the oracle, logits of the
                       stateless forward, of a prefill and of single-token steps, and the recurrent states, side by side;
                       tests/test_oracle.py::test_oracle_against_the_real_stripedhyena_package asserts on it whenever the
                       package is present (and is skipped, saying so, when it is not) -- that test passing is what would remove
                       the "parity unpinned" label;
  * bench.py's CPU legs time the real model instead of the oracle port when `build` succeeds (cpu_baseline.kind "reference")."""
from __future__ import annotations

import contextlib
import importlib
from typing import Dict, Optional

import torch

from . import stripedhyena_oracle as O


def available() -> Optional[str]:
    """Version string (or "unknown") of an importable stripedhyena, else None."""
    try:
        pkg = importlib.import_module("stripedhyena")
        importlib.import_module("stripedhyena.model")
    except Exception:
        return None
    return str(getattr(pkg, "__version__", "unknown"))


@contextlib.contextmanager
def cpu_rotary():
    """flash_attn.layers.rotary.apply_rotary_emb_qkv_ (a Triton kernel) -> flash_attn's own apply_rotary_emb_torch on q and k,
    for the duration of the block.  Also patched wherever stripedhyena re-imported the name."""
    try:
        import flash_attn.layers.rotary as R
    except Exception:
        yield
        return

    def qkv_rotary_torch(qkv, cos, sin, cos_k=None, sin_k=None, interleaved=False, seqlen_offsets=0, num_heads_q=None):
        L = qkv.shape[1]
        off = int(seqlen_offsets)
        c, s = cos[off:off + L], sin[off:off + L]
        ck, sk = (c, s) if cos_k is None else (cos_k[off:off + L], sin_k[off:off + L])
        q = R.apply_rotary_emb_torch(qkv[:, :, 0], c, s, interleaved)
        k = R.apply_rotary_emb_torch(qkv[:, :, 1], ck, sk, interleaved)
        return torch.stack([q, k, qkv[:, :, 2]], dim=2)

    holders = [R]
    for name in ("stripedhyena.positional_embeddings", "stripedhyena.layers", "stripedhyena.model"):
        try:
            mod = importlib.import_module(name)
            if hasattr(mod, "apply_rotary_emb_qkv_"):
                holders.append(mod)
        except Exception:
            pass
    saved = [(h, h.apply_rotary_emb_qkv_) for h in holders]
    for h in holders:
        h.apply_rotary_emb_qkv_ = qkv_rotary_torch
    try:
        yield
    finally:
        for h, fn in saved:
            h.apply_rotary_emb_qkv_ = fn


_CPU_SWITCHES = {"use_flash_attn": False, "use_flash_rmsnorm": False, "use_flash_depthwise": False, "use_flashfft": False}


def build(cfg: dict, state_dict: Optional[Dict[str, torch.Tensor]], dtype=torch.bfloat16, share_blocks: bool = False):
    """The real StripedHyena on CPU with `state_dict` (or its own random init when None).
    dtype bfloat16 = the reference's own policy (to_bfloat16_except_poles_residues, evo/models.py:148); float32 = parameters
    left as constructed (fp32) with the checkpoint values copied in.
    share_blocks: every block of a kind aliases the first one's parameters (the 7B shape in ~1 GB for the bench's CPU leg)."""
    from stripedhyena.model import StripedHyena
    from stripedhyena.utils import dotdict
    conf = dotdict({**cfg, **_CPU_SWITCHES})
    if share_blocks:
        first = {}
        with torch.device("meta"):
            model = StripedHyena(conf)
        model = model.to_empty(device="cpu")
        gen = torch.Generator().manual_seed(0)
        with torch.no_grad():
            for i, blk in enumerate(model.blocks):
                kind = type(blk).__name__
                if kind not in first:
                    first[kind] = blk
                    continue
                src = dict(first[kind].named_parameters())
                for name, p in blk.named_parameters():
                    p.data = src[name].data
            seen = set()
            for name, p in model.named_parameters():
                if p.data_ptr() in seen:
                    continue
                seen.add(p.data_ptr())
                if name.endswith("poles"):          # inside the unit disc, like the oracle's random_state_dict
                    mag = 0.5 + 0.45 * torch.rand(p.shape[:-1], generator=gen)
                    ang = (torch.rand(p.shape[:-1], generator=gen) * 2 - 1) * 3.141592653589793
                    p.copy_(torch.stack([mag * torch.cos(ang), mag * torch.sin(ang)], dim=-1))
                elif name.endswith("scale"):
                    p.fill_(1.0)
                else:
                    p.normal_(0.0, 0.02, generator=gen)
            for name, b in model.named_buffers():       # to_empty left the buffers uninitialised too
                if name.endswith("inv_freq"):
                    dim = 2 * b.numel()
                    b.copy_(1.0 / (float(cfg.get("rotary_emb_base") or 10000) ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim)))
    else:
        model = StripedHyena(conf)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    if dtype == torch.bfloat16:
        model.to_bfloat16_except_poles_residues()
    model.eval()
    return model


def _run(model, ids, ipd=None):
    with torch.inference_mode(), cpu_rotary():
        out = model(ids) if ipd is None else model(ids, inference_params_dict=ipd)
    return out[0] if isinstance(out, tuple) else out


def compare(num_layers: int = 3, attn_layer_idxs=(1,), seed: int = 7, L: int = 48, steps: int = 3, extra: Optional[dict] = None) -> dict:
    """Real package vs oracle on one tiny model (hidden 256, 2 heads, head_dim 128 like the 7B).  Returns, per arithmetic mode,
    the largest |difference| of the logits (stateless / prefill / each step) and of the Hyena states, next to the logits' scale."""
    cfg = O.tiny_config(num_layers=num_layers, attn_layer_idxs=attn_layer_idxs, hidden_size=256, num_heads=2, **(extra or {}))
    cfg["max_seqlen"] = 128
    sd = O.random_state_dict(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, 4, (2, L + steps), generator=g) * 3 + 65
    ids[:, 0] = 0
    report = {"keys_real_minus_oracle": None, "keys_oracle_minus_real": None}
    for mode, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        real = build(cfg, sd, dtype)
        if report["keys_real_minus_oracle"] is None:
            rk, ok = set(real.state_dict()), set(O.state_dict_spec(cfg))
            report["keys_real_minus_oracle"], report["keys_oracle_minus_real"] = sorted(rk - ok), sorted(ok - rk)
        oracle = O.OracleStripedHyena(cfg, sd, dtype)
        r, o = _run(real, ids[:, :L]).float(), oracle(ids[:, :L])[0].float()
        out = {"scale": o.abs().max().item(), "stateless": (r - o).abs().max().item()}
        dr, do = real.initialize_inference_params(), oracle.initialize_inference_params()
        for d in (dr, do):
            d["mha"].max_batch_size = 2
            if hasattr(d["hyena"], "max_batch_size"):
                d["hyena"].max_batch_size = 2
        r, o = _run(real, ids[:, :L], dr).float(), oracle(ids[:, :L], do)[0].float()
        out["prefill"] = (r - o).abs().max().item()
        out["state_keys_equal"] = sorted(dr["hyena"].state_dict) == sorted(do["hyena"].state_dict) and sorted(dr["mha"].key_value_memory_dict) == sorted(do["mha"].key_value_memory_dict)
        out["state"] = max((torch.view_as_real(dr["hyena"].state_dict[k].to(torch.complex64)) - torch.view_as_real(do["hyena"].state_dict[k].to(torch.complex64))).abs().max().item()
                           for k in do["hyena"].state_dict)
        out["fir_state"] = max((dr["hyena"].fir_state_dict[k].float() - do["hyena"].fir_state_dict[k].float()).abs().max().item() for k in do["hyena"].fir_state_dict)
        worst = 0.0
        for t in range(steps):
            for d in (dr, do):
                d["mha"].seqlen_offset = d["hyena"].seqlen_offset = L + t
            r, o = _run(real, ids[:, L + t:L + t + 1], dr).float(), oracle(ids[:, L + t:L + t + 1], do)[0].float()
            worst = max(worst, (r - o).abs().max().item())
        out["steps"] = worst
        report[mode] = out
    return report
