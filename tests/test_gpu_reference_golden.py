"""The CUDA path against numbers produced by the REFERENCE'S OWN scoring code (tests/golden/reference_host.json / .npz, made
by tests/golden/make_reference_host_golden.py: /root/reference/evo/scoring.py run unmodified over the oracle model).

`score_sequences` / `positional_entropies` of evo_b200 on the GPU -- device tokenise + pad, the sm_100a forward, the fused
unembed + log-softmax + gather / entropy head -- must land on what evo/scoring.py:62-131 returned for the same sequences and
the same weights in exact (fp64) arithmetic, within the bf16 noise of this model (stated per assertion; the reference's own
bf16 pipeline, also in the fixture, sits 0.009-0.02 nats from the fp64 numbers on these sequences)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

from oracle import stripedhyena_oracle as O          # noqa: E402  (tests may use the oracle)
from evo_b200 import _lib                             # noqa: E402
from evo_b200.stripedhyena import StripedHyena, dotdict  # noqa: E402

DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    _lib.lib()        # raises if the extension is missing: no fallback


@pytest.fixture(scope="module")
def ref(golden_dir):
    with open(os.path.join(golden_dir, "reference_host.json")) as f:
        doc = json.load(f)
    return doc, np.load(os.path.join(golden_dir, "reference_host.npz"))


def _model():
    cfg = O.tiny_config(num_layers=3, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)     # the fixture's model
    cfg["max_seqlen"] = 128
    m = StripedHyena(dotdict(cfg))
    m.load_state_dict(O.random_state_dict(cfg, seed=7), strict=True)
    m.to_bfloat16_except_poles_residues()
    return m.to(DEV)


def test_gpu_scores_land_on_the_reference_scoring_code_output(ref):
    import evo_b200
    doc, _ = ref
    sc = doc["scoring"]
    m, tok = _model(), evo_b200.CharLevelTokenizer(512)
    seqs = sc["seqs"]
    mean = evo_b200.score_sequences(seqs, m, tok, reduce_method="mean", device=DEV)
    total = evo_b200.score_sequences(seqs, m, tok, reduce_method="sum", device=DEV)
    for k, s in enumerate(seqs):
        # per-position bf16 noise of this tiny random model is ~3.5e-2 nats (tests/test_gpu_parity.py::test_model_logits_vs_oracle);
        # a mean over 6..30 positions stays within 8e-2 of exact arithmetic, and within 0.1 of the reference's bf16 pipeline
        assert abs(float(mean[k]) - sc["score_mean_fp64"][k]) < 8e-2, (k, float(mean[k]), sc["score_mean_fp64"][k])
        assert abs(float(mean[k]) - sc["score_mean_bf16"][k]) < 0.1, (k, float(mean[k]), sc["score_mean_bf16"][k])
        assert abs(float(total[k]) - sc["score_sum_fp64"][k]) < 8e-2 * len(s), (k, float(total[k]), sc["score_sum_fp64"][k])
        assert abs(float(total[k]) - float(mean[k]) * len(s)) < 1e-3 * len(s)
    with pytest.raises(ValueError) as ex:
        evo_b200.score_sequences(seqs, m, tok, reduce_method="median", device=DEV)
    assert str(ex.value) == sc["bad_reduce"]


def test_gpu_entropies_land_on_the_reference_scoring_code_output(ref):
    import evo_b200
    doc, arr = ref
    seqs = doc["scoring"]["seqs"]
    ent = evo_b200.positional_entropies(seqs, _model(), evo_b200.CharLevelTokenizer(512), device=DEV)
    assert [len(e) for e in ent] == [len(s) for s in seqs]
    diffs = np.concatenate([np.abs(np.asarray(e, dtype=np.float64) - arr[f"entropy_fp64_{k}"]) for k, e in enumerate(ent)])
    # entropies here span 0.5 .. 4.5 nats; the reference's bf16 pipeline (restated on CPU) is 0.021 mean / 0.082 max from exact arithmetic
    assert diffs.mean() < 0.1 and diffs.max() < 0.4, (diffs.mean(), diffs.max())


def test_gpu_kv_append_replays_flash_attn_update_kv_cache(golden_dir):
    """evo_kv_append against flash_attn's own _update_kv_cache (modules/mha.py:338-367, run on CPU for
    tests/golden/kvcache_sampling_flash_attn.npz): a 5-token prefill, two single-token appends and a 3-token continuation into a
    (3, 16) cache at batch 2 must leave the same cache (a copy: bit-exact on the bf16-rounded inputs)."""
    import ctypes as C
    g = np.load(os.path.join(golden_dir, "kvcache_sampling_flash_attn.npz"))
    lib = _lib.lib()
    H, d = g["q_0"].shape[2], g["q_0"].shape[3]
    cache = torch.zeros(3, 16, 2, H, d, dtype=torch.bfloat16, device=DEV)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for n, (off, L) in enumerate(g["steps"]):
        off, L = int(off), int(L)
        kv = torch.from_numpy(g[f"kv_{n}"]).bfloat16()                     # (B, L, 2, H, d)
        qkv = torch.cat([torch.from_numpy(g[f"q_{n}"]).bfloat16()[:, :, None], kv], dim=2).contiguous().to(DEV)     # (B, L, 3, H, d)
        _lib.check(lib.evo_kv_append(_lib.ptr(qkv), _lib.ptr(cache), qkv.shape[0], L, H, d, off, 16, stream), "evo_kv_append")
    torch.cuda.synchronize()
    assert torch.equal(cache.cpu(), torch.from_numpy(g["cache_final"]).bfloat16())


def test_gpu_attention_layer_through_the_c_abi_lands_on_flash_attn_mha_forward(golden_dir):
    """The product's attention layer -- evo_rope_tables, evo_gemm with the bias + rotary epilogue (Wqkv), evo_attn_fwd_ws (the
    default variant), evo_kv_append, evo_gemm with the bias epilogue (out_proj) -- against flash_attn's MHA.forward run on CPU in
    fp32 (tests/golden/mha_flash_attn.npz, made by make_golden.py: mha_block_fixture): the stateless forward over 8 tokens, and
    token 6 as a single-token step over a cache holding the 6-token prefill.  Tolerance: the bf16 pipeline restated on CPU is
    3.8e-3 max / 6.5e-4 mean from these fp32 numbers (outputs up to 1.06); the GPU gets 0.02 / 0.004."""
    import ctypes as C
    import math
    sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
    import gpu_bringup as G
    from evo_b200.stripedhyena import model as M_
    G._imports()
    g = np.load(os.path.join(golden_dir, "mha_flash_attn.npz"))
    lib = _lib.lib()
    stream = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    B, L, D, H, hd = 2, 8, 256, 2, 128
    dev = lambda name: torch.from_numpy(g[name]).bfloat16().to(DEV).contiguous()
    x, wqkv, bqkv, wo, bo = dev("x"), dev("Wqkv_w"), dev("Wqkv_b"), dev("out_w"), dev("out_b")
    inv_freq = torch.from_numpy(g["inv_freq"]).float().to(DEV).contiguous()
    cos = torch.empty(32, hd // 2, dtype=torch.bfloat16, device=DEV)
    sin = torch.empty_like(cos)
    _lib.check(lib.evo_rope_tables(_lib.ptr(cos), _lib.ptr(sin), _lib.ptr(inv_freq), 0, 32, hd // 2, 1.0, stream()), "evo_rope_tables")

    def wqkv_rope(rows, n_seq, n_tok, off):
        out = torch.full((n_seq * n_tok, 3 * D), float("nan"), dtype=torch.bfloat16, device=DEV)
        p = _lib.GemmParams(A=rows.data_ptr(), lda=D, W=wqkv.data_ptr(), C=out.data_ptr(), ldc=3 * D, bias=bqkv.data_ptr(), residual=None, ldr=3 * D,
                            M=n_seq * n_tok, N=3 * D, K=D, epilogue=_lib.EPI_BIAS_ROPE, variant=0,
                            rope_cos=cos.data_ptr() + off * (hd // 2) * 2, rope_sin=sin.data_ptr() + off * (hd // 2) * 2, rope_L=n_tok, rope_cols=2 * D)
        _lib.check(lib.evo_gemm(C.byref(p), stream()), "evo_gemm(rope)")
        return out

    def close(got, want):
        err = (got.float().cpu() - torch.from_numpy(want).reshape(got.shape)).abs()
        assert torch.isfinite(got.float()).all() and err.max().item() < 0.02 and err.mean().item() < 0.004, (err.max().item(), err.mean().item())

    # stateless: MHA.forward(x)
    qkv = wqkv_rope(x.view(B * L, D), B, L, 0)
    ctx = G._attn(qkv, B, L, H, M_.ATTN_VARIANT)
    y = G._gemm(ctx.view(B * L, D), wo, B * L, D, D, _lib.EPI_BIAS, 0, bias=bo)
    close(y.view(B, L, D), g["y_stateless"])
    # prefill of 6 tokens into a (2, 32) cache, then token 6 as one step at seqlen_offset 6 (mha.py:344-367, 502-540)
    cache = torch.zeros(B, 32, 2, H, hd, dtype=torch.bfloat16, device=DEV)
    pre = wqkv_rope(x[:, :6].contiguous().view(B * 6, D), B, 6, 0)
    _lib.check(lib.evo_kv_append(_lib.ptr(pre), _lib.ptr(cache), B, 6, H, hd, 0, 32, stream()), "evo_kv_append")
    y_pre = G._gemm(G._attn(pre, B, 6, H, M_.ATTN_VARIANT, cache=cache, off=0).view(B * 6, D), wo, B * 6, D, D, _lib.EPI_BIAS, 0, bias=bo)
    close(y_pre.view(B, 6, D), g["y_prefill"])
    step = wqkv_rope(x[:, 6:7].contiguous().view(B, D), B, 1, 6)
    _lib.check(lib.evo_kv_append(_lib.ptr(step), _lib.ptr(cache), B, 1, H, hd, 6, 32, stream()), "evo_kv_append")
    y6 = G._gemm(G._attn(step, B, 1, H, M_.ATTN_VARIANT, cache=cache, off=6).view(B, D), wo, B, D, D, _lib.EPI_BIAS, 0, bias=bo)
    close(y6.view(B, 1, D), g["y_step6"])
    want_cache = torch.from_numpy(g["cache"])[:, :7]
    err = (cache[:, :7].float().cpu() - want_cache).abs()
    assert err.max().item() < 0.03 * max(1.0, want_cache.abs().max().item())        # rotary'd k / projected v, bf16 vs fp32
