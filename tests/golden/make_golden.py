"""Generates the committed fixtures under tests/golden/ (run in the BUILD container only).

1. attention_flash_attn.npz -- outputs of flash_attn's OWN PyTorch code paths, imported from
   the installed package (v2.8.3): RotaryEmbedding._update_cos_sin_cache (layers/rotary.py:
   382-416), apply_rotary_emb_torch (layers/rotary.py:23-35) and SelfAttention (modules/
   mha.py:230-279).  These pin the oracle's rotary/attention restatement to the real thing.
2. tiny_model_oracle.npz -- ids + fp64 logits + prefill states of the oracle on a 3-layer
   model (Hyena, attention, Hyena).  stripedhyena itself is not importable here (SURVEY.md
   section 0.1), so this fixture guards the oracle against regressions; it does not pin it.

3. kvcache_sampling_flash_attn.npz -- flash_attn's _update_kv_cache, CrossAttention(causal=True) and utils.generation.sample
   run on CPU (see kvcache_and_sampling_fixture).
4. mha_flash_attn.npz -- flash_attn's whole MHA.forward on CPU (stateless, prefill + steps) with the Triton rotary call replaced
   by flash_attn's own torch statement of it (see mha_block_fixture).

    python tests/golden/make_golden.py [attention] [tiny_model] [kvcache_sampling] [mha]
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


def kvcache_and_sampling_fixture():
    """3. kvcache_sampling_flash_attn.npz -- more of flash_attn's OWN torch code, run on CPU:
       * _update_kv_cache (modules/mha.py:338-367) with flash_attn.utils.generation.InferenceParams: a prefill of 5 tokens, then
         two single-token appends at seqlen_offset 5 and 6 into a (max_batch 3, max_seqlen 16) cache, batch 2;
       * CrossAttention(causal=True) (mha.py:280-335) on (q of the new tokens, kv = cache[:, :offset + Lq]): the decode / continued
         prefill attention MHA.forward takes when use_flash_attn is False (mha.py:502-540, 689-700) -- bottom-right aligned mask;
       * utils/generation.sample (generation.py:70-98), the code stripedhyena/sample.py copies: picks for seeded draws over
         top-k / top-p / temperature settings on fp32 logits."""
    from flash_attn.modules.mha import CrossAttention, _update_kv_cache
    from flash_attn.utils.generation import InferenceParams, sample

    torch.manual_seed(4321)
    B, H, d = 2, 2, 128
    out = {}
    ip = InferenceParams(max_seqlen=16, max_batch_size=3)
    attn = CrossAttention(causal=True)
    steps = [(0, 5), (5, 1), (6, 1), (7, 3)]            # (seqlen_offset, new tokens): prefill, two decode steps, a continued prefill
    for n, (off, L) in enumerate(steps):
        q = torch.randn(B, L, H, d, dtype=torch.float32)
        kv = torch.randn(B, L, 2, H, d, dtype=torch.float32)
        ip.seqlen_offset = off
        fresh = 7 not in ip.key_value_memory_dict
        seen = _update_kv_cache(kv, ip, 7)
        if fresh:       # torch.empty in the reference: give the never-written part a value so the fixture is reproducible
            cache = ip.key_value_memory_dict[7]
            cache[B:] = 0
            cache[:, L:] = 0
        assert tuple(seen.shape) == (B, off + L, 2, H, d)
        ctx = attn(q, seen)
        out[f"q_{n}"], out[f"kv_{n}"], out[f"ctx_{n}"] = q.numpy(), kv.numpy(), ctx.numpy()
    out["cache_final"] = ip.key_value_memory_dict[7].numpy().copy()
    out["steps"] = np.array(steps)
    # sampling
    g = torch.Generator().manual_seed(99)
    logits = torch.randn(6, 512, generator=g) * 2.5
    out["sample_logits"] = logits.numpy()
    cases = [(1, 0.0, 1.0), (4, 0.0, 1.0), (4, 1.0, 0.7), (50, 0.7, 1.0), (50, 0.7, 0.5), (0, 0.9, 1.0), (0, 0.0, 1.3), (600, 0.5, 1.0)]
    out["sample_cases"] = np.array(cases, dtype=np.float64)
    for n, (k, p, t) in enumerate(cases):
        picks = []
        for seed in range(8):
            torch.manual_seed(1000 + seed)
            picks.append(sample(logits.clone(), top_k=k, top_p=p, temperature=t).numpy())
        out[f"sample_picks_{n}"] = np.stack(picks)
    np.savez_compressed(os.path.join(HERE, "kvcache_sampling_flash_attn.npz"), **out)


def mha_block_fixture():
    """4. mha_flash_attn.npz -- flash_attn's MHA.forward itself (modules/mha.py:573-704) on CPU, fp32, configured as StripedHyena's
    AttentionBlock configures it (causal, qkv and out-proj bias, rotary over the full head_dim 128, use_flash_attn=False -- the
    torch SelfAttention / CrossAttention branch): stateless over 8 tokens; then with flash_attn's InferenceParams a 6-token
    prefill and two single-token steps (rotary at seqlen_offset with the table built to max_seqlen, _update_kv_cache,
    cache-form attention).  ONE patch, the one SURVEY.md 8c names: RotaryEmbedding.forward ends in a Triton kernel that cannot run
    on CPU, so `apply_rotary_emb_qkv_` is replaced by flash_attn's own torch statement of it (apply_rotary_emb_torch on q and k at
    cos/sin[offset : offset + L]); everything else -- Wqkv, the `(three h d)` split, the cos/sin cache, the cache write, the masks,
    out_proj -- is flash_attn's code as installed."""
    import flash_attn.layers.rotary as R
    from flash_attn.modules.mha import MHA
    from flash_attn.utils.generation import InferenceParams

    def qkv_rotary_torch(qkv, cos, sin, cos_k=None, sin_k=None, interleaved=False, seqlen_offsets=0, num_heads_q=None):
        assert cos_k is None and sin_k is None and isinstance(seqlen_offsets, int) and num_heads_q is None
        L = qkv.shape[1]
        c, s_ = cos[seqlen_offsets:seqlen_offsets + L], sin[seqlen_offsets:seqlen_offsets + L]
        q = R.apply_rotary_emb_torch(qkv[:, :, 0], c, s_, interleaved)
        k = R.apply_rotary_emb_torch(qkv[:, :, 1], c, s_, interleaved)
        return torch.stack([q, k, qkv[:, :, 2]], dim=2)

    real = R.apply_rotary_emb_qkv_
    R.apply_rotary_emb_qkv_ = qkv_rotary_torch
    try:
        torch.manual_seed(777)
        D, H, B = 256, 2, 2
        mha = MHA(D, H, qkv_proj_bias=True, out_proj_bias=True, causal=True, layer_idx=1, rotary_emb_dim=D // H, use_flash_attn=False)
        with torch.no_grad():
            mha.Wqkv.bias.normal_(0, 0.1)
            mha.out_proj.bias.normal_(0, 0.1)
        x = torch.randn(B, 8, D)
        out = {"x": x.numpy(), "Wqkv_w": mha.Wqkv.weight.detach().numpy(), "Wqkv_b": mha.Wqkv.bias.detach().numpy(),
               "out_w": mha.out_proj.weight.detach().numpy(), "out_b": mha.out_proj.bias.detach().numpy(), "inv_freq": mha.rotary_emb.inv_freq.numpy()}
        with torch.no_grad():
            out["y_stateless"] = mha(x).numpy()
            ip = InferenceParams(max_seqlen=32, max_batch_size=B)
            out["y_prefill"] = mha(x[:, :6], inference_params=ip).numpy()
            ip.key_value_memory_dict[1][:, 6:] = 0           # torch.empty in flash_attn: pin the never-written rows
            ip.seqlen_offset = 6
            out["y_step6"] = mha(x[:, 6:7], inference_params=ip).numpy()
            ip.seqlen_offset = 7
            out["y_step7"] = mha(x[:, 7:8], inference_params=ip).numpy()
            out["cache"] = ip.key_value_memory_dict[1].numpy().copy()
    finally:
        R.apply_rotary_emb_qkv_ = real
    np.savez_compressed(os.path.join(HERE, "mha_flash_attn.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["attention", "tiny_model", "kvcache_sampling", "mha"]
    if "mha" in which:
        mha_block_fixture()
    if "attention" in which:
        attention_fixture()
    if "tiny_model" in which:
        tiny_model_fixture()
    if "kvcache_sampling" in which:
        kvcache_and_sampling_fixture()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
