"""CPU oracle for the StripedHyena (Evo-1 / Evo-1.5 7B) forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``evo_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and there only as the checker / the CPU arm.

PARITY UNPINNED (stated as the task requires): the arithmetic of the reference
path lives in the third-party package ``stripedhyena==0.2.2`` (pinned at
/root/reference/requirements.txt:1 and environment.yml:10), which is neither
vendored in /root/reference nor installed in the build container -- nor obtainable on the
GPU box (round 2: no package index, no wheel in /opt/wheelhouse, no HF cache; log in
profiles/r02_pin_attempt_call1.log) -- and the reference repository has no tests, golden
vectors or fixtures for this path.  This file therefore RESTATES the published
algorithm of stripedhyena 0.2.2 (model.py / engine.py / layers.py / cache.py /
sample.py / positional_embeddings.py of github.com/togethercomputer/stripedhyena
at the 0.2.x tag) in plain PyTorch, anchored on the reference's call sites:

  evo/models.py:141-150   StripedHyena(cfg); load_state_dict(strict=True);
                          to_bfloat16_except_poles_residues(); .to(device)
  evo/scoring.py:80-81    logits, _ = model(input_ids)
  evo/generation.py:117   model.initialize_inference_params()
  evo/generation.py:152   logits, d = model(x, inference_params_dict=d)
  evo/generation.py:162   stripedhyena.sample.sample(...)

What IS run from the reference itself: the host code around the model call.  tests/golden/
make_reference_host_golden.py imports /root/reference/evo/{tokenizer,scoring,generation,models}.py,
scripts/{score,generate}.py and semantic_design/semantic_design.py unmodified (a stand-in satisfies
their `import stripedhyena`) and drives them with THIS model; tests/golden/reference_host.{json,npz}
hold what they returned.  That pins evo_b200's host layer, not this file's arithmetic.

The attention sub-path is restated from flash_attn (installed in the build
container, v2.8.3): flash_attn/modules/mha.py:573-704 (MHA.forward),
mha.py:230-279 (SelfAttention), layers/rotary.py:23-35 and :382-416.  That part
IS pinned: tests/golden/make_golden.py imports flash_attn's own torch code paths
and the fixtures under tests/golden/ hold its outputs (tests/test_oracle.py): rotary
tables and application, SelfAttention, _update_kv_cache, CrossAttention in cache form,
utils.generation.sample, and MHA.forward as a whole (stateless, prefill, steps).

Two arithmetic modes:
  dtype=torch.bfloat16  "faithful": every op runs in the dtype the reference runs
                        it in (bf16 parameters/activations, fp32 FFT / filter /
                        softmax accumulation, fp32 poles & residues), so the
                        rounding points match the reference's.
  dtype=torch.float32 / float64  "truth": same graph, wide arithmetic.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# Config (keys: /root/reference/evo/configs/evo-1-8k-base_inference.yml:1-38,
#         evo-1-131k-base_inference.yml:39-40)
# --------------------------------------------------------------------------

_EVO_7B = dict(
    vocab_size=512, hidden_size=4096, num_filters=4096, max_sequence_len=8192,
    attn_layer_idxs=[8, 16, 24],
    hyena_layer_idxs=[i for i in range(32) if i not in (8, 16, 24)],
    num_layers=32, short_filter_length=3, num_attention_heads=32,
    short_filter_bias=True, eps=1.0e-6, state_size=8, inner_size_multiple_of=16,
    proj_groups=1, hyena_filter_groups=1, model_parallel_size=1,
    tie_embeddings=True, mha_out_proj_bias=True, qkv_proj_bias=True,
    final_norm=True, column_split=True, prefill_style="fft",
    mlp_activation="gelu",
)


def evo_config(name: str = "evo-1-8k-base") -> dict:
    cfg = dict(_EVO_7B)
    if name == "evo-1-131k-base":
        cfg["use_interpolated_rotary_pos_emb"] = True
        cfg["rotary_emb_scaling_factor"] = 16
    return cfg


def tiny_config(num_layers=2, attn_layer_idxs=(1,), hidden_size=256, num_heads=2,
                vocab_size=512, **extra) -> dict:
    """Small StripedHyena with the 7B's structure (head_dim stays 128)."""
    cfg = dict(_EVO_7B)
    cfg.update(
        vocab_size=vocab_size, hidden_size=hidden_size, num_filters=hidden_size,
        num_layers=num_layers, num_attention_heads=num_heads,
        attn_layer_idxs=list(attn_layer_idxs),
        hyena_layer_idxs=[i for i in range(num_layers) if i not in attn_layer_idxs],
    )
    cfg.update(extra)
    return cfg


def mlp_inner_size(cfg) -> int:
    """stripedhyena layers.ParallelGatedMLP.__init__ (0.2.2)."""
    mult = cfg.get("inner_size_multiple_of", 64) * cfg.get("model_parallel_size", 1)
    inner = int(2 * cfg["hidden_size"] * 4 / 3)
    inner = mult * ((inner + mult - 1) // mult)
    if cfg.get("inner_mlp_size") is not None:
        inner = cfg["inner_mlp_size"]
    return inner


# --------------------------------------------------------------------------
# State dict (key names = HF checkpoint after the 'backbone.' strip,
# /root/reference/evo/models.py:124-137)
# --------------------------------------------------------------------------

def state_dict_spec(cfg) -> Dict[str, tuple]:
    D, V, S = cfg["hidden_size"], cfg["vocab_size"], cfg["state_size"]
    inner = mlp_inner_size(cfg)
    k = cfg["short_filter_length"]
    spec = {"embedding_layer.weight": (V, D), "unembed.weight": (V, D), "norm.scale": (D,)}
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}."
        spec[p + "pre_norm.scale"] = (D,)
        spec[p + "post_norm.scale"] = (D,)
        spec[p + "mlp.l1.weight"] = (inner, D)
        spec[p + "mlp.l2.weight"] = (inner, D)
        spec[p + "mlp.l3.weight"] = (D, inner)
        if i in cfg["attn_layer_idxs"]:
            spec[p + "inner_mha_cls.Wqkv.weight"] = (3 * D, D)
            spec[p + "inner_mha_cls.Wqkv.bias"] = (3 * D,)
            spec[p + "inner_mha_cls.out_proj.weight"] = (D, D)
            spec[p + "inner_mha_cls.out_proj.bias"] = (D,)
            spec[p + "inner_mha_cls.rotary_emb.inv_freq"] = (D // cfg["num_attention_heads"] // 2,)
        else:
            spec[p + "projections.weight"] = (3 * D, D)
            spec[p + "projections.bias"] = (3 * D,)
            spec[p + "out_filter_dense.weight"] = (D, D)
            spec[p + "out_filter_dense.bias"] = (D,)
            spec[p + "filter.short_filter_weight"] = (3 * D, 1, k)
            spec[p + "filter.short_filter_bias"] = (3 * D,)
            spec[p + "filter.D"] = (D,)
            spec[p + "filter.poles"] = (D // cfg["hyena_filter_groups"], S, 1, 2)
            spec[p + "filter.residues"] = spec[p + "filter.poles"]
    return spec


def random_state_dict(cfg, seed: int = 0, dtype=torch.bfloat16, share_blocks: bool = False) -> Dict[str, torch.Tensor]:
    """Random-init weights with sane activation scales.  Poles are drawn inside the
    unit disc (|p| in [0.5, 0.999], random phase) as SURVEY.md section 8d prescribes.
    Everything is `dtype` except poles/residues (fp32), mirroring
    to_bfloat16_except_poles_residues (evo/models.py:148).  share_blocks=True makes
    every block of a kind alias block 0's tensors (CPU-baseline timing with a
    small host-RAM footprint; the arithmetic per block is unchanged)."""
    g = torch.Generator().manual_seed(seed)
    D = cfg["hidden_size"]
    hd = D // cfg["num_attention_heads"]
    sd: Dict[str, torch.Tensor] = {}
    first_of_kind: Dict[str, str] = {}

    def rn(shape, std):
        return (torch.randn(shape, generator=g) * std)

    for name, shape in state_dict_spec(cfg).items():
        if share_blocks and name.startswith("blocks."):
            idx = int(name.split(".")[1])
            kind = "a" if idx in cfg["attn_layer_idxs"] else "h"
            suffix = name.split(".", 2)[2]
            key = kind + suffix
            if key in first_of_kind:
                sd[name] = sd[first_of_kind[key]]
                continue
            first_of_kind[key] = name
        if name == "unembed.weight":
            continue
        if name.endswith("poles"):
            mag = 0.5 + 0.499 * torch.rand(shape[:2], generator=g)
            ph = (torch.rand(shape[:2], generator=g) * 2 - 1) * math.pi
            t = torch.stack([mag * torch.cos(ph), mag * torch.sin(ph)], -1)[:, :, None, :]
            sd[name] = t.float().contiguous()
        elif name.endswith("residues"):
            sd[name] = rn(shape, 0.35).float()
        elif name.endswith("inv_freq"):
            sd[name] = (1.0 / (10000 ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))).to(dtype)
        elif name.endswith("scale"):
            sd[name] = (1.0 + rn(shape, 0.1)).to(dtype)
        elif name.endswith("short_filter_weight"):
            sd[name] = rn(shape, 0.5).to(dtype)
        elif name.endswith("filter.D"):
            sd[name] = rn(shape, 0.5).to(dtype)
        elif name.endswith("bias"):
            sd[name] = rn(shape, 0.1).to(dtype)
        elif name == "embedding_layer.weight":
            sd[name] = rn(shape, 2.5 / math.sqrt(D)).to(dtype)  # logits std ~2.5 through the tied unembed
        else:  # Linear weights (out, in)
            sd[name] = rn(shape, 1.0 / math.sqrt(shape[1])).to(dtype)
    sd["unembed.weight"] = sd["embedding_layer.weight"]  # tied (evo/models.py:136-137)
    return sd


# --------------------------------------------------------------------------
# Inference-state holders (stripedhyena cache.py; attribute names pinned by
# /root/reference/evo/generation.py:109-119,140-148)
# --------------------------------------------------------------------------

@dataclass
class InferenceParams:
    max_seqlen: int
    max_batch_size: int
    seqlen_offset: int = 0
    batch_size_offset: int = 0
    key_value_memory_dict: dict = field(default_factory=dict)
    lengths_per_sample: Optional[torch.Tensor] = None


@dataclass
class RecurrentInferenceParams:
    fir_filter_length: int = 3
    state_dim: int = 16
    seqlen_offset: int = 0
    fir_state_dict: dict = field(default_factory=dict)
    state_dict: dict = field(default_factory=dict)


# --------------------------------------------------------------------------
# Layers
# --------------------------------------------------------------------------

def rms_norm(x, scale, eps):
    """stripedhyena layers.RMSNorm.forward, non-flash branch: eps is added to the
    RMS (outside the root), all in the input dtype."""
    d = x.shape[-1]
    y = x / (x.norm(2, dim=-1, keepdim=True) * d ** (-1.0 / 2) + eps)
    return scale * y


def gated_mlp(x, w1, w2, w3):
    """ParallelGatedMLP.forward: l3(gelu(l1 x) * l2 x), exact-erf GELU, no biases."""
    return F.linear(F.gelu(F.linear(x, w1)) * F.linear(x, w2), w3)


def column_split(z, num_heads, head_dim):
    """engine.parallel_iir / utils.column_split: channels viewed as
    (heads, 3*head_dim); per head the first head_dim -> x2, next -> x1, last -> v.
    z: (B, 3D, L) or (B, 3D)."""
    if z.dim() == 3:
        zz = z.reshape(z.shape[0], num_heads, 3 * head_dim, z.shape[2])
        x2, x1, v = zz[:, :, :head_dim], zz[:, :, head_dim:2 * head_dim], zz[:, :, 2 * head_dim:]
        return tuple(t.reshape(t.shape[0], -1, t.shape[-1]) for t in (x2, x1, v))
    zz = z.reshape(z.shape[0], num_heads, 3 * head_dim)
    x2, x1, v = zz[:, :, :head_dim], zz[:, :, head_dim:2 * head_dim], zz[:, :, 2 * head_dim:]
    return tuple(t.reshape(t.shape[0], -1) for t in (x2, x1, v))


def fir_parallel(u, weight, bias):
    """engine.parallel_fir (F.conv1d branch).  u: (B, L, 3D).  Returns z_pre
    (B, 3D, L) and fir_state = last (k-1) inputs, (B, 3D, k-1).  Bias is added
    separately after the conv, as the reference does."""
    L = u.shape[1]
    k = weight.shape[-1]
    ut = u.permute(0, 2, 1)
    z = F.conv1d(ut, weight, bias=None, stride=1, padding=k - 1, groups=ut.shape[1])[..., :L]
    z = z + bias[None, :, None]
    return z, ut[..., -(k - 1):]


def hyena_filter(poles, residues, L):
    """ParallelHyenaFilter.compute_filter: h[c,t] = Re sum_s R[c,s] exp(t log p[c,s]), fp32."""
    t = torch.arange(L, device=poles.device)[None, None]
    wide = torch.float64 if poles.dtype == torch.float64 else torch.float32  # reference: fp32
    r = torch.view_as_complex(residues.to(wide))
    lp = torch.view_as_complex(poles.to(wide)).log()
    h = (r * (lp * t).exp()).real.sum(1)[None]
    return h  # (1, D, L) fp32


def iir_parallel(z_pre, h, Dskip, poles, num_heads, head_dim, want_state: bool):
    """engine.parallel_iir, long_fir_threshold=None, use_flashfft=False.
    z_pre: (B, 3D, L).  Returns y (B, L, D) and, when want_state, the modal state
    at the last position (prefill_via_modal_fft), complex64 (B, D, S)."""
    L = z_pre.shape[-1]
    fft_size = 2 * L
    x2, x1, v = column_split(z_pre, num_heads, head_dim)
    x1v = x1 * v
    wide = torch.float64 if z_pre.dtype == torch.float64 else torch.float32  # reference: fp32
    H = torch.fft.rfft(h.to(wide), n=fft_size) / fft_size
    X_s = torch.fft.fft(x1v.to(wide), n=fft_size)
    X = X_s[..., : H.shape[-1]]
    y = torch.fft.irfft(X * H, n=fft_size, norm="forward")[..., :L]
    y = y.to(dtype=x1v.dtype)
    y = (y + x1v * Dskip.unsqueeze(-1)) * x2
    state = None
    if want_state:
        t = torch.arange(L, device=poles.device)[None, None]
        p = torch.view_as_complex(poles.to(wide))
        state_s = p ** t
        state_S = torch.fft.fft(state_s, n=fft_size)[None]
        st = torch.fft.ifft(X_s[..., None, :] * state_S, n=fft_size)
        state = st[..., L - 1].to(torch.complex128 if wide == torch.float64 else torch.complex64)
    return y.permute(0, 2, 1), state


def fir_step(u, fir_state, weight, bias):
    """engine.step_fir.  u: (B, 3D); fir_state (B, 3D, k-1) holds u[t-2], u[t-1]."""
    h0, h = weight[..., 0, -1], weight[..., 0, :-1]
    y = h0[None] * u + torch.sum(fir_state * h[None], dim=-1) + bias
    fir_state = torch.roll(fir_state, -1, dims=2)
    fir_state[..., -1] = u
    return y, fir_state


def iir_step(x2, x1, v, Dskip, residues, poles, iir_state):
    """engine.step_iir: state <- p*state + x1v; y = x2 * (Re sum R*state + D*x1v)."""
    x1v = x1 * v
    wide = torch.float64 if poles.dtype == torch.float64 else torch.float32  # reference: fp32
    r = torch.view_as_complex(residues.to(wide))[..., 0][None]
    p = torch.view_as_complex(poles.to(wide))[..., 0][None]
    iir_state = p * iir_state + x1v[..., None]
    res_state = torch.sum(r * iir_state, dim=-1).real
    y = x2 * (res_state + Dskip * x1v)
    return y, iir_state


def rotary_tables(seqlen, head_dim, base=10000.0, scaling_factor=1.0, dtype=torch.bfloat16):
    """flash_attn layers/rotary.py:382-416 with pos_idx_in_fp32=True; stripedhyena's
    LinearlyScaledRotaryEmbedding divides the positions by the scaling factor."""
    t = torch.arange(seqlen, dtype=torch.float32)
    t = t / scaling_factor
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    freqs = torch.outer(t, inv_freq)
    return torch.cos(freqs).to(dtype), torch.sin(freqs).to(dtype)


def apply_rotary(x, cos, sin):
    """Non-interleaved (NeoX) rotary over the full head_dim; x (B, L, H, d), cos/sin
    (L, d/2).  Arithmetic as flash_attn's kernel (ops/triton/rotary.py): operands
    upcast to fp32, one rounding on store."""
    d2 = cos.shape[-1]
    xf = x.to(torch.float32) if x.dtype in (torch.bfloat16, torch.float16) else x
    c = cos.to(xf.dtype)[None, :, None, :]
    s = sin.to(xf.dtype)[None, :, None, :]
    x0, x1 = xf[..., :d2], xf[..., d2:]
    return torch.cat([x0 * c - x1 * s, x0 * s + x1 * c], dim=-1).to(x.dtype)


def causal_attention(q, k, v, q_offset: int = 0):
    """softmax(q k^T / sqrt(d)) v with a causal mask, FlashAttention arithmetic:
    fp32 scores / softmax statistics, probabilities rounded to the value dtype
    before the PV product, normalisation applied in fp32 at the end.
    q (B, Lq, H, d); k, v (B, Lk, H, d); query i sits at absolute position q_offset+i."""
    B, Lq, H, d = q.shape
    Lk = k.shape[1]
    wide = torch.float32 if q.dtype in (torch.bfloat16, torch.float16) else q.dtype
    scale = 1.0 / math.sqrt(d)
    out = torch.empty_like(q)
    qi = torch.arange(Lq)[:, None] + q_offset
    kj = torch.arange(Lk)[None, :]
    mask = kj > qi
    blk = max(1, min(Lq, (1 << 24) // max(Lk, 1)))
    for b in range(B):
        for hh in range(H):
            kf = k[b, :, hh].to(wide)
            vf = v[b, :, hh].to(wide)
            for r0 in range(0, Lq, blk):
                r1 = min(Lq, r0 + blk)
                s = (q[b, r0:r1, hh].to(wide) @ kf.T) * scale
                s = s.masked_fill(mask[r0:r1], float("-inf"))
                m = s.max(dim=-1, keepdim=True).values
                p = torch.exp(s - m)
                denom = p.sum(dim=-1, keepdim=True)
                o = (p.to(v.dtype).to(wide) @ vf) / denom
                out[b, r0:r1, hh] = o.to(q.dtype)
    return out


# --------------------------------------------------------------------------
# Model
# --------------------------------------------------------------------------

class OracleStripedHyena:
    """Restatement of stripedhyena.model.StripedHyena for inference."""

    def __init__(self, cfg: dict, state_dict: Dict[str, torch.Tensor], dtype=torch.bfloat16):
        self.cfg = cfg
        self.dtype = dtype
        missing = set(state_dict_spec(cfg)) - set(state_dict)
        extra = set(state_dict) - set(state_dict_spec(cfg))
        if missing or extra:  # load_state_dict(strict=True), evo/models.py:147
            raise RuntimeError(f"state_dict mismatch: missing={sorted(missing)[:4]} unexpected={sorted(extra)[:4]}")
        self.sd = {}
        for k, v in state_dict.items():
            if "poles" in k or "residues" in k:
                self.sd[k] = v.to(torch.float32 if dtype != torch.float64 else torch.float64)
            else:
                self.sd[k] = v.to(dtype)
        self.D = cfg["hidden_size"]
        self.H = cfg["num_attention_heads"]
        self.hd = self.D // self.H
        self.rope_scale = float(cfg.get("rotary_emb_scaling_factor", 1.0)) if cfg.get("use_interpolated_rotary_pos_emb", False) else 1.0
        self.taps = {}  # optional per-block activations for tests

    # -- API mirrored from stripedhyena.model.StripedHyena ------------------
    def initialize_inference_params(self):
        return {
            "mha": InferenceParams(max_seqlen=self.cfg.get("max_seqlen", 8192),
                                   max_batch_size=self.cfg.get("max_batch_size", 1), seqlen_offset=0),
            "hyena": RecurrentInferenceParams(fir_filter_length=self.cfg["short_filter_length"],
                                              state_dim=self.cfg["state_size"], seqlen_offset=0),
        }

    def __call__(self, ids, inference_params_dict=None, tap: bool = False):
        return self.forward(ids, inference_params_dict, tap)

    def forward(self, ids, inference_params_dict=None, tap: bool = False):
        x = F.embedding(ids.long(), self.sd["embedding_layer.weight"])
        for i in range(self.cfg["num_layers"]):
            if i in self.cfg["attn_layer_idxs"]:
                ip = inference_params_dict["mha"] if inference_params_dict is not None else None
                x = self.attention_block(i, x, ip)
            else:
                ip = inference_params_dict["hyena"] if inference_params_dict is not None else None
                x = self.hyena_block(i, x, ip)
            if tap:
                self.taps[f"block{i}"] = x.clone()
        if self.cfg.get("final_norm", True):
            x = rms_norm(x, self.sd["norm.scale"], self.cfg["eps"])
        logits = x @ self.sd["unembed.weight"].T
        return logits, inference_params_dict

    # -- blocks --------------------------------------------------------------
    def _mlp_res(self, p, u):
        sd = self.sd
        xn = rms_norm(u, sd[p + "post_norm.scale"], self.cfg["eps"])
        return gated_mlp(xn, sd[p + "mlp.l1.weight"], sd[p + "mlp.l2.weight"], sd[p + "mlp.l3.weight"]) + u

    def hyena_block(self, i, u, ip=None):
        """ParallelGatedConvBlock.forward."""
        sd, p = self.sd, f"blocks.{i}."
        z = F.linear(rms_norm(u, sd[p + "pre_norm.scale"], self.cfg["eps"]), sd[p + "projections.weight"], sd[p + "projections.bias"])
        y = self.hyena_operator(i, z, ip)
        z_in = F.linear(y, sd[p + "out_filter_dense.weight"], sd[p + "out_filter_dense.bias"]) + u
        return self._mlp_res(p, z_in)

    def hyena_operator(self, i, z, ip=None):
        """ParallelHyenaFilter.forward: sequential path iff this layer already has
        a fir_state, else the parallel (FFT) path, which populates the states when
        inference params are given."""
        sd, p = self.sd, f"blocks.{i}.filter."
        w, b, Dk = sd[p + "short_filter_weight"], sd[p + "short_filter_bias"], sd[p + "D"]
        poles, residues = sd[p + "poles"], sd[p + "residues"]
        if ip is not None and i in ip.fir_state_dict:
            u = z[:, -1]
            z_pre, fir_state = fir_step(u, ip.fir_state_dict[i], w, b)
            x2, x1, v = column_split(z_pre, self.H, self.hd)
            y, st = iir_step(x2, x1, v, Dk, residues, poles, ip.state_dict[i])
            ip.fir_state_dict[i] = fir_state
            ip.state_dict[i] = st
            return y.to(z.dtype)[:, None]
        L = z.shape[1]
        z_pre, fir_state = fir_parallel(z, w, b)
        h = hyena_filter(poles, residues, L)
        y, state = iir_parallel(z_pre, h, Dk, poles, self.H, self.hd, want_state=ip is not None)
        if ip is not None:
            ip.fir_state_dict[i] = fir_state.clone()
            ip.state_dict[i] = state
        return y

    def attention_block(self, i, u, ip=None):
        """AttentionBlock.forward -> flash_attn MHA.forward (mha.py:573-704)."""
        sd, p = self.sd, f"blocks.{i}."
        B, L, _ = u.shape
        xn = rms_norm(u, sd[p + "pre_norm.scale"], self.cfg["eps"])
        qkv = F.linear(xn, sd[p + "inner_mha_cls.Wqkv.weight"], sd[p + "inner_mha_cls.Wqkv.bias"])
        qkv = qkv.reshape(B, L, 3, self.H, self.hd)
        off = ip.seqlen_offset if ip is not None else 0
        cos, sin = rotary_tables(off + L, self.hd, scaling_factor=self.rope_scale, dtype=self.dtype)
        q = apply_rotary(qkv[:, :, 0], cos[off:], sin[off:])
        k = apply_rotary(qkv[:, :, 1], cos[off:], sin[off:])
        v = qkv[:, :, 2]
        if ip is not None:
            if i not in ip.key_value_memory_dict:  # mha.py:344-353
                ip.key_value_memory_dict[i] = torch.zeros(ip.max_batch_size, ip.max_seqlen, 2, self.H, self.hd, dtype=u.dtype)
            cache = ip.key_value_memory_dict[i]
            assert off + L <= cache.shape[1]  # mha.py:367
            cache[:B, off:off + L, 0] = k
            cache[:B, off:off + L, 1] = v
            k, v = cache[:B, :off + L, 0], cache[:B, :off + L, 1]
        ctx = causal_attention(q, k, v, q_offset=off).reshape(B, L, self.D)
        a = F.linear(ctx, sd[p + "inner_mha_cls.out_proj.weight"], sd[p + "inner_mha_cls.out_proj.bias"]) + u
        return self._mlp_res(p, a)


# --------------------------------------------------------------------------
# Independent time-domain definitions (used by tests to cross-check the FFT path)
# --------------------------------------------------------------------------

def long_conv_direct(x1v, poles, residues):
    """y[c,t] = sum_{tau<=t} h[c,t-tau] x1v[c,tau] by the modal recurrence in
    float64/complex128.  x1v: (B, D, L) any float dtype -> (B, D, L) float64,
    final state (B, D, S) complex128."""
    p = torch.view_as_complex(poles.to(torch.float64))[..., 0]
    r = torch.view_as_complex(residues.to(torch.float64))[..., 0]
    B, Dm, L = x1v.shape
    st = torch.zeros(B, Dm, p.shape[-1], dtype=torch.complex128)
    out = torch.empty(B, Dm, L, dtype=torch.float64)
    xd = x1v.to(torch.float64)
    for t in range(L):
        st = p[None] * st + xd[:, :, t, None]
        out[:, :, t] = (r[None] * st).sum(-1).real
    return out, st


# --------------------------------------------------------------------------
# Sampling (stripedhyena sample.py; call site evo/generation.py:162-167)
# --------------------------------------------------------------------------

def sample(logits, top_k=1, top_p=0.0, temperature=1.0):
    logits = logits.squeeze(1) if logits.dim() == 3 else logits
    if top_k == 1:
        return logits.argmax(dim=-1)
    if top_k > 0:
        top_k = min(top_k, logits.size(-1))
        lt, idx = torch.topk(logits, top_k, dim=-1)
        if temperature != 1.0:
            lt = lt / temperature
        _top_p_filter(lt, top_p)
        pick = torch.multinomial(torch.softmax(lt, dim=-1), num_samples=1).squeeze(-1)
        return idx[torch.arange(idx.shape[0]), pick]
    lt = logits / temperature if temperature != 1.0 else logits.clone()
    _top_p_filter(lt, top_p)
    return torch.multinomial(torch.softmax(lt, dim=-1), num_samples=1).squeeze(-1)


def _top_p_filter(logits, top_p):
    if top_p <= 0.0 or top_p >= 1.0:
        return
    sl, si = torch.sort(logits, descending=False)
    cp = sl.softmax(dim=-1).cumsum(dim=-1)
    rem = cp <= (1 - top_p)
    logits.masked_fill_(rem.scatter(1, si, rem), float("-inf"))


# --------------------------------------------------------------------------
# Time-domain form of the Hyena operator with explicit history (tests: continued prefill,
# sequence sharding).  Wide arithmetic only; the algebra, not the rounding, is what it checks.
# --------------------------------------------------------------------------

def hyena_operator_time_domain(z, fir_w, fir_b, Dskip, poles, residues, num_heads, head_dim, halo=None, state_in=None):
    """z (B, L, 3D) float64.  halo (B, 2, 3D): the two z rows before row 0 (zeros if None);
    state_in (B, D, S) complex128: modal state before row 0.  Returns y (B, L, D) float64 and
    the state after the last row."""
    B, L, C3 = z.shape
    zz = torch.cat([halo if halo is not None else torch.zeros(B, 2, C3, dtype=z.dtype), z], dim=1)
    w = fir_w.to(z.dtype)[:, 0]                                                     # (3D, 3)
    zp = zz[:, 0:L] * w[:, 0] + zz[:, 1:L + 1] * w[:, 1] + zz[:, 2:L + 2] * w[:, 2] + fir_b.to(z.dtype)
    x2, x1, v = column_split(zp.permute(0, 2, 1), num_heads, head_dim)              # (B, D, L)
    x1v = x1 * v
    p = torch.view_as_complex(poles.to(torch.float64))[..., 0]
    r = torch.view_as_complex(residues.to(torch.float64))[..., 0]
    st = state_in.clone() if state_in is not None else torch.zeros(B, p.shape[0], p.shape[1], dtype=torch.complex128)
    conv = torch.empty_like(x1v)
    for t in range(L):
        st = p[None] * st + x1v[:, :, t, None]
        conv[:, :, t] = (r[None] * st).sum(-1).real
    y = (conv + x1v * Dskip.to(z.dtype)[None, :, None]) * x2
    return y.permute(0, 2, 1), st
