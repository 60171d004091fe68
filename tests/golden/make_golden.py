"""Generates the committed fixtures under tests/golden/ (run in the BUILD container only).

1. attention_flash_attn.npz -- outputs of flash_attn's OWN PyTorch code paths, imported from
   the installed package (v2.8.3): RotaryEmbedding._update_cos_sin_cache (layers/rotary.py:
   382-416), apply_rotary_emb_torch (layers/rotary.py:23-35) and SelfAttention (modules/
   mha.py:230-279).  These pin the oracle's rotary/attention restatement to the real thing.
2. tiny_model_oracle.npz -- ids + fp64 logits + prefill states of the oracle on a 3-layer
   model (Hyena, attention, Hyena).  stripedhyena itself is not importable here (SURVEY.md
   section 0.1), so this fixture guards the oracle against regressions; it does not pin it.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def attention_fixture():
    from flash_attn.layers.rotary import RotaryEmbedding, apply_rotary_emb_torch
    from flash_attn.modules.mha import SelfAttention

    torch.manual_seed(1234)
    B, L, H, d = 2, 48, 2, 128
    qkv = torch.randn(B, L, 3, H, d, dtype=torch.float32)
    out = {"qkv": qkv.numpy()}
    for name, scaling in (("s1", 1.0), ("s16", 16.0)):
        rot = RotaryEmbedding(d, base=10000.0, interleaved=False, device="cpu")
        if scaling == 1.0:
            rot._update_cos_sin_cache(L, device="cpu", dtype=torch.float32)
            cos, sin = rot._cos_cached, rot._sin_cached
        else:
            # stripedhyena's LinearlyScaledRotaryEmbedding: same code with t /= scaling_factor
            t = torch.arange(L, dtype=torch.float32) / scaling
            freqs = torch.outer(t, rot.inv_freq.to(torch.float32))
            cos, sin = torch.cos(freqs), torch.sin(freqs)
        q = apply_rotary_emb_torch(qkv[:, :, 0], cos, sin, interleaved=False)
        k = apply_rotary_emb_torch(qkv[:, :, 1], cos, sin, interleaved=False)
        rq = torch.stack([q, k, qkv[:, :, 2]], dim=2)
        ctx = SelfAttention(causal=True)(rq)
        out[f"cos_{name}"] = cos.numpy(); out[f"sin_{name}"] = sin.numpy()
        out[f"q_{name}"] = q.numpy(); out[f"k_{name}"] = k.numpy(); out[f"ctx_{name}"] = ctx.numpy()
    np.savez_compressed(os.path.join(HERE, "attention_flash_attn.npz"), **out)


def tiny_model_fixture():
    from oracle.stripedhyena_oracle import OracleStripedHyena, random_state_dict, tiny_config
    cfg = tiny_config(num_layers=3, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    sd = random_state_dict(cfg, seed=7)
    rng = np.random.default_rng(7)
    ids = torch.from_numpy(rng.choice(np.array([65, 67, 71, 84]), size=(2, 77))).long()
    ids[:, 0] = 0
    m = OracleStripedHyena(cfg, sd, torch.float64)
    logits, _ = m(ids)
    d = m.initialize_inference_params()
    d["mha"].max_batch_size = 2; d["mha"].max_seqlen = 128
    m(ids[:, :40], d)
    np.savez_compressed(os.path.join(HERE, "tiny_model_oracle.npz"), ids=ids.numpy(), logits=logits.numpy().astype(np.float32),
                        state0_re=d["hyena"].state_dict[0].real.numpy().astype(np.float32),
                        state0_im=d["hyena"].state_dict[0].imag.numpy().astype(np.float32),
                        fir0=d["hyena"].fir_state_dict[0].to(torch.float32).numpy())


if __name__ == "__main__":
    attention_fixture()
    tiny_model_fixture()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
