"""Host-side logic that needs no GPU: tokenizer, batching, scoring math, sampling, config,
the StripedHyena parameter tree / strict loading, and the generation loop's state protocol
(driven with a CPU stand-in model)."""
import json
import os

import numpy as np
import pytest
import torch
import yaml

import evo_b200
from evo_b200 import CharLevelTokenizer, logits_to_logprobs, prepare_batch
from evo_b200.configs import MODEL_NAMES, get_config
from evo_b200.generation import Generator
from evo_b200.models import load_checkpoint
from evo_b200.stripedhyena import StripedHyena, dotdict, sample
from oracle import stripedhyena_oracle as O


def test_tokenizer_roundtrip():
    tok = CharLevelTokenizer(512)
    assert tok.tokenize("ACGT") == [65, 67, 71, 84]
    assert (tok.eod_id, tok.eos_id, tok.pad_id, tok.vocab_size) == (0, 0, 1, 512)
    assert tok.detokenize([65, 67, 71, 84]) == "ACGT"
    assert tok.detokenize([0, 1, 31]) == "   "          # clamped to >= 32 (evo/tokenizer.py:22-23)
    assert tok.detokenize_batch(torch.tensor([[65, 67], [71, 84]])) == ["AC", "GT"]
    assert tok.tokenize_batch(["A", "CG"]) == [[65], [67, 71]]


def test_prepare_batch_pads_and_prepends_bos():
    tok = CharLevelTokenizer(512)
    ids, lens = prepare_batch(["ACGT", "AC"], tok, prepend_bos=True, device="cpu")
    assert lens == [4, 2] and ids.dtype == torch.long
    assert ids.tolist() == [[0, 65, 67, 71, 84], [0, 65, 67, 1, 1]]
    ids, _ = prepare_batch(["ACGT"], tok, prepend_bos=False, device="cpu")
    assert ids.tolist() == [[65, 67, 71, 84]]


def test_logits_to_logprobs_alignment():
    torch.manual_seed(0)
    logits = torch.randn(2, 5, 8)
    ids = torch.randint(0, 8, (2, 5))
    lp = logits_to_logprobs(logits, ids, trim_bos=True)
    ref = torch.log_softmax(logits, -1)
    assert lp.shape == (2, 4)
    for b in range(2):
        for t in range(4):
            assert lp[b, t] == ref[b, t, ids[b, t + 1]]
    assert logits_to_logprobs(logits, ids, trim_bos=False).shape == (2, 5)


def test_sample_matches_oracle_semantics():
    lg = torch.randn(6, 512)
    assert torch.equal(sample(lg, top_k=1), lg.argmax(-1))
    assert torch.equal(sample(lg[:, None, :], top_k=1), lg.argmax(-1))
    torch.manual_seed(3); a = sample(lg.clone(), top_k=8, top_p=0.8, temperature=0.9)
    torch.manual_seed(3); b = O.sample(lg.clone(), top_k=8, top_p=0.8, temperature=0.9)
    assert torch.equal(a, b)
    torch.manual_seed(4); a = sample(lg.clone(), top_k=0, top_p=0.5, temperature=1.3)
    torch.manual_seed(4); b = O.sample(lg.clone(), top_k=0, top_p=0.5, temperature=1.3)
    assert torch.equal(a, b)


def test_configs_match_reference_values():
    assert MODEL_NAMES == ["evo-1.5-8k-base", "evo-1-8k-base", "evo-1-131k-base", "evo-1-8k-crispr", "evo-1-8k-transposon"]
    c = get_config("evo-1-8k-base")
    assert (c["hidden_size"], c["num_layers"], c["num_attention_heads"], c["state_size"], c["vocab_size"]) == (4096, 32, 32, 8, 512)
    assert c["attn_layer_idxs"] == [8, 16, 24] and len(c["hyena_layer_idxs"]) == 29
    assert c["eps"] == 1e-6 and c["short_filter_length"] == 3 and c["inner_size_multiple_of"] == 16
    assert "rotary_emb_scaling_factor" not in c
    c2 = get_config("evo-1-131k-base")
    assert c2["use_interpolated_rotary_pos_emb"] is True and c2["rotary_emb_scaling_factor"] == 16
    with pytest.raises(ValueError):
        get_config("evo-2")
    with pytest.raises(ValueError):
        evo_b200.Evo("not-a-model")


def test_dotdict():
    d = dotdict({"a": 1}, Loader="x")
    assert d.a == 1 and d.Loader == "x" and d.missing is None
    d.b = 2
    assert d["b"] == 2


def test_parameter_tree_matches_checkpoint_keys_and_strict_loading():
    cfg = O.tiny_config(num_layers=3, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    m = StripedHyena(dotdict(cfg))
    spec = O.state_dict_spec(cfg)
    sd = m.state_dict()
    assert set(sd) == set(spec)
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    good = O.random_state_dict(cfg)
    m.load_state_dict(good, strict=True)
    bad = dict(good); bad.pop("blocks.0.filter.poles")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)
    m.to_bfloat16_except_poles_residues()
    for k, p in m.named_parameters():
        assert p.dtype == (torch.float32 if ("poles" in k or "residues" in k) else torch.bfloat16), k
    assert m.unembed is m.embedding_layer              # tied (evo/models.py:136-137)
    ip = m.initialize_inference_params()
    assert ip["mha"].max_seqlen == 8192 and ip["mha"].seqlen_offset == 0 and ip["hyena"].state_dim == 8
    assert ip["hyena"].fir_state_dict == {} and ip["mha"].key_value_memory_dict == {}


def test_no_cpu_path():
    cfg = O.tiny_config(num_layers=1, attn_layer_idxs=(), hidden_size=256, num_heads=2)
    m = StripedHyena(dotdict(cfg)).to_bfloat16_except_poles_residues()
    with pytest.raises(evo_b200._lib.EvoError, match="CUDA"):
        m(torch.zeros(1, 4, dtype=torch.long))


class _OracleAsModel:
    """CPU stand-in with the boundary protocol, to drive the host generation loop."""

    def __init__(self, cfg, sd):
        self.m = O.OracleStripedHyena(cfg, sd, torch.float32)
        self.calls = []

    def eval(self):
        return self

    def initialize_inference_params(self):
        d = self.m.initialize_inference_params()
        d["mha"].max_seqlen = 256
        return d

    def __call__(self, x, inference_params_dict=None):
        self.calls.append((tuple(x.shape), None if inference_params_dict is None else inference_params_dict["mha"].seqlen_offset))
        return self.m(x, inference_params_dict)


@pytest.mark.parametrize("cached", [True, False])
def test_generation_loop_protocol_and_greedy_parity(cached):
    cfg = O.tiny_config(num_layers=2, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    sd = O.random_state_dict(cfg, seed=2)
    model = _OracleAsModel(cfg, sd)
    tok = CharLevelTokenizer(512)
    g = Generator(model, tok, top_k=1)
    prompt = torch.tensor([[65, 67, 71, 84, 65, 65]])
    out_ids, out_logits, d = g.generate(device="cpu", input_ids=prompt, num_tokens=5, cached_generation=cached,
                                        force_prompt_threshold=128, print_generation=False, stop_at_eos=False)
    assert out_ids.shape == (1, 5) and out_logits.shape == (1, 5, 512)
    # greedy continuation must equal repeated full forwards of the oracle
    seq = prompt.clone()
    for t in range(5):
        lg, _ = model.m(seq)
        nxt = lg[:, -1].argmax(-1)
        assert nxt.item() == out_ids[0, t].item()
        seq = torch.cat([seq, nxt[:, None]], 1)
    if cached:
        assert model.calls[0] == ((1, 6), 0)
        assert model.calls[1] == ((1, 1), 6) and model.calls[2] == ((1, 1), 7)
        assert set(d["hyena"].fir_state_dict) == {0} and set(d["mha"].key_value_memory_dict) == {1}
    else:
        assert d is None and model.calls[1][0] == (1, 7)


def test_generation_prompt_forcing_bookkeeping():
    """threshold < prompt length: prefill `thr` tokens, teacher-force the rest one by one;
    the reference then sets seqlen_offset to the FULL prompt length (quirk Q1)."""
    cfg = O.tiny_config(num_layers=2, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    model = _OracleAsModel(cfg, O.random_state_dict(cfg, seed=2))
    g = Generator(model, CharLevelTokenizer(512), top_k=1)
    prompt = torch.tensor([[65, 67, 71, 84, 65, 65, 67, 67]])
    out_ids, _, d = g.generate(device="cpu", input_ids=prompt, num_tokens=3, cached_generation=True,
                               force_prompt_threshold=3, print_generation=False, stop_at_eos=False)
    assert out_ids.shape == (1, 3)
    assert model.calls[0] == ((1, 3), 0)
    assert model.calls[1] == ((1, 1), 8)          # jump to len(prompt), not 3
    assert len(model.calls) == 5 + 3


def test_load_checkpoint_from_local_safetensors(tmp_path):
    """Checkpoint ingest (evo/models.py:96-150 contract): sharded safetensors with the HF 'backbone.' prefix and NO
    'unembed.weight' (tied embeddings) -> strict load -> bf16 except poles/residues; config from a YAML with the
    reference's keys."""
    import json
    import yaml
    from safetensors.torch import save_file
    from evo_b200.models import load_checkpoint

    cfg = O.tiny_config(num_layers=3, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    sd = O.random_state_dict(cfg, seed=9)
    sd.pop("unembed.weight")
    names = sorted(sd)
    shards = {"model-00001-of-00002.safetensors": names[: len(names) // 2], "model-00002-of-00002.safetensors": names[len(names) // 2:]}
    weight_map = {}
    for fname, keys in shards.items():
        save_file({"backbone." + k: sd[k].contiguous() for k in keys}, str(tmp_path / fname))
        weight_map.update({"backbone." + k: fname for k in keys})
    (tmp_path / "model.safetensors.index.json").write_text(json.dumps({"weight_map": weight_map}))
    cfg_path = tmp_path / "tiny.yml"
    cfg_path.write_text(yaml.safe_dump(cfg))
    m = load_checkpoint("evo-1-8k-base", config_path=str(cfg_path), model_dir=str(tmp_path))
    got = m.state_dict()
    assert set(got) == set(O.state_dict_spec(cfg))
    for k in names:
        want = sd[k] if ("poles" in k or "residues" in k) else sd[k].to(torch.bfloat16)
        assert torch.equal(got[k].cpu(), want), k
    assert torch.equal(got["unembed.weight"], got["embedding_layer.weight"])
    assert got["blocks.0.filter.poles"].dtype == torch.float32 and got["blocks.0.mlp.l1.weight"].dtype == torch.bfloat16
    # a checkpoint with a missing tensor must fail loudly (strict=True)
    bad = tmp_path / "bad"
    bad.mkdir()
    save_file({"backbone." + k: sd[k].contiguous() for k in names[1:]}, str(bad / "model.safetensors"))
    with pytest.raises(RuntimeError):
        load_checkpoint("evo-1-8k-base", config_path=str(cfg_path), model_dir=str(bad))
    with pytest.raises(FileNotFoundError):
        load_checkpoint("evo-1-8k-base", config_path=str(cfg_path), model_dir=str(tmp_path / "nope"))


def test_rmsnorm_reciprocal_multiply_is_exact_after_bf16_rounding():
    """csrc/elementwise.cu computes bf16(x / n) as bf16(x * rcp(n)).  Exhaustive over every pair of bf16 significands
    (and a spread of exponents): the two agree bit for bit, also when the reciprocal is one fp32 ulp off."""
    import torch

    def all_bf16(e_lo, e_hi):
        sig = torch.arange(128, dtype=torch.float32) / 128 + 1.0
        ex = torch.arange(e_lo, e_hi, dtype=torch.float32)
        return (sig[None, :] * torch.pow(torch.tensor(2.0), ex)[:, None]).reshape(-1)

    x, n = all_bf16(-4, 4), all_bf16(-20, 6)
    q = (x[:, None] / n[None, :]).bfloat16()
    r = 1.0 / n
    assert torch.equal((x[:, None] * r[None, :]).bfloat16(), q)
    assert torch.equal((x[:, None] * torch.nextafter(r, torch.tensor(float("inf")))[None, :]).bfloat16(), q)


def test_prepare_batch_empty_list_raises_like_the_reference():
    """evo/scoring.py:21 takes max() of the lengths: an empty batch is a ValueError there, and here."""
    tok = evo_b200.CharLevelTokenizer(512)
    with pytest.raises(ValueError):
        evo_b200.prepare_batch([], tok, device="cpu")
    ids, lengths = evo_b200.prepare_batch(["", "AC"], tok, device="cpu")      # an empty sequence is padded, length 0
    assert ids.tolist() == [[0, 1, 1], [0, 65, 67]] and lengths == [0, 2]


def test_streamk_partition_invariants():
    """Python mirror of the index math in csrc/gemm_smallm.cu (range_start / cta_of / slot choice): for many
    (tiles, k-blocks, grid) the ranges tile [0, total) without gaps, every partial segment gets a slot nobody else uses,
    and the contributor window [c_first, c_last] of a tile is exactly the set of CTAs with a segment in it."""
    def range_start(c, total, parts):
        return (c * total) // parts

    def cta_of(it, total, parts):
        c = (it * parts) // total
        while c + 1 < parts and range_start(c + 1, total, parts) <= it:
            c += 1
        while c > 0 and range_start(c, total, parts) > it:
            c -= 1
        return c

    rng = np.random.default_rng(0)
    cases = [(96, 64, 148), (32, 64, 148), (86, 64, 148), (32, 172, 148), (4, 64, 32), (1, 1, 1), (2, 1, 2), (3, 7, 2)]
    cases += [(int(rng.integers(1, 200)), int(rng.integers(1, 200)), 148) for _ in range(40)]
    for n_tiles, KB, sms in cases:
        total = n_tiles * KB
        parts = max(1, min(sms, total // 8))
        assert total * sms < 2 ** 31
        segs = {}                                            # tile -> list of (cta, k0, k1, slot)
        covered = 0
        for c in range(parts):
            b, e = range_start(c, total, parts), range_start(c + 1, total, parts)
            assert b == covered and e > b
            covered = e
            it, first_tile = b, b // KB
            while it < e:
                tile, k0 = it // KB, it % KB
                k1 = min(KB, k0 + (e - it))
                if not (k0 == 0 and k1 == KB):
                    segs.setdefault(tile, []).append((c, k0, k1, 2 * c + (0 if tile == first_tile else 1)))
                it += k1 - k0
        assert covered == total
        slots = [s for v in segs.values() for (_, _, _, s) in v]
        assert len(slots) == len(set(slots)) and all(s < 2 * parts for s in slots)
        for tile, v in segs.items():
            c_first, c_last = cta_of(tile * KB, total, parts), cta_of((tile + 1) * KB - 1, total, parts)
            assert [c for (c, _, _, _) in v] == list(range(c_first, c_last + 1))
            assert sum(k1 - k0 for (_, k0, k1, _) in v) == KB
            for (c, _, _, s) in v:                           # the finisher recomputes every contributor's slot the same way
                assert s == 2 * c + (0 if range_start(c, total, parts) // KB == tile else 1)


def test_generate_function_batches_equal_lengths_and_scores_like_the_reference(capsys):
    """evo/generation.py:207-297: equal-length prompts run as one batch, ragged ones one by one (with notes on stderr);
    the returned score is mean(logits_to_logprobs(new_logits, new_ids)) with the reference's trim_bos alignment (Q3)."""
    cfg = O.tiny_config(num_layers=2, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    sd = O.random_state_dict(cfg, seed=2)
    tok = CharLevelTokenizer(512)
    model = _OracleAsModel(cfg, sd)
    texts, scores = evo_b200.generate(["ACGT", "TTGA"], model, tok, n_tokens=4, top_k=1, cached_generation=True, verbose=0, device="cpu")
    assert len(texts) == 2 and all(len(t) == 4 for t in texts) and len(scores) == 2
    assert model.calls[0][0] == (2, 4)                              # one batched prefill
    # the same prompts one at a time give the same greedy continuations and the same scores
    model1 = _OracleAsModel(cfg, sd)
    texts1, scores1 = evo_b200.generate(["ACGT", "TTGA"], model1, tok, n_tokens=4, top_k=1, cached_generation=True, verbose=1,
                                        batched=False, device="cpu")
    err = capsys.readouterr().err
    assert "Will not do batched generation" in err and "different lengths" not in err
    assert texts1 == texts and np.allclose(scores1, scores, atol=1e-4)
    assert model1.calls[0][0] == (1, 4)
    # score definition: recompute from a direct Generator call
    g = Generator(_OracleAsModel(cfg, sd), tok, top_k=1)
    ids, _ = prepare_batch(["ACGT"], tok, prepend_bos=False, device="cpu")
    new_ids, new_logits, _ = g.generate(device="cpu", input_ids=ids, num_tokens=4, cached_generation=True, print_generation=False, stop_at_eos=False)
    want = logits_to_logprobs(new_logits, new_ids).float().mean().item()
    assert abs(want - scores[0]) < 1e-4
    # ragged prompts: falls back to one at a time and says why
    evo_b200.generate(["ACGT", "TT"], _OracleAsModel(cfg, sd), tok, n_tokens=2, top_k=1, cached_generation=False, verbose=1, device="cpu")
    assert "different lengths" in capsys.readouterr().err


def test_config_path_resolution_matches_the_reference(tmp_path):
    """evo/models.py:141 loads config_path relative to the package, or fails; a typo must not silently run the default geometry."""
    from evo_b200.models import _resolve_config
    from evo_b200.configs import get_config
    assert _resolve_config("evo-1-8k-base", None) == get_config("evo-1-8k-base")
    assert _resolve_config("evo-1-131k-base", "configs/evo-1-131k-base_inference.yml")["rotary_emb_scaling_factor"] == 16
    assert "rotary_emb_scaling_factor" not in _resolve_config("evo-1-8k-base", "configs/evo-1-8k-base_inference.yml")
    p = tmp_path / "custom.yml"
    p.write_text("hidden_size: 64\n")
    assert _resolve_config("evo-1-8k-base", str(p)) == {"hidden_size": 64}
    with pytest.raises(FileNotFoundError):
        _resolve_config("evo-1-8k-base", str(tmp_path / "typo.yml"))
    with pytest.raises(FileNotFoundError):
        _resolve_config("evo-1-8k-base", "elsewhere/evo-1-8k-base_inference.yml")


def _write_sharded_checkpoint(tmp_path, cfg, seed=9, drop=("unembed.weight",), fp32_keys=()):
    from safetensors.torch import save_file
    sd = O.random_state_dict(cfg, seed=seed)
    for k in drop:
        sd.pop(k)
    names = sorted(sd)
    cut = len(names) // 2
    shards = {"model-00001-of-00002.safetensors": names[:cut], "model-00002-of-00002.safetensors": names[cut:]}
    weight_map = {}
    for fname, keys in shards.items():
        save_file({"backbone." + k: (sd[k].float() if k in fp32_keys else sd[k]).contiguous() for k in keys}, str(tmp_path / fname))
        weight_map.update({"backbone." + k: fname for k in keys})
    (tmp_path / "model.safetensors.index.json").write_text(json.dumps({"weight_map": weight_map}))
    cfg_path = tmp_path / "tiny.yml"
    cfg_path.write_text(yaml.safe_dump(cfg))
    return sd, names, str(cfg_path)


def test_streaming_ingest_equals_host_state_dict_path(tmp_path):
    """evo_b200/ingest.py (mmap -> staging -> in-place cast / pack) vs the reference's sequence (load_file -> strict
    load_state_dict -> bf16): identical state dicts, including fp32-on-disk tensors cast on arrival, the tied unembed, the
    MLP weights packed into w12 / w3 with exact zero padding, and strict-mode failures."""
    from evo_b200.ingest import iter_safetensors, shard_files
    cfg = O.tiny_config(num_layers=3, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    fp32_on_disk = ("blocks.0.mlp.l1.weight", "blocks.2.projections.weight", "norm.scale")
    sd, names, cfg_path = _write_sharded_checkpoint(tmp_path, cfg, fp32_keys=fp32_on_disk)
    a = load_checkpoint("evo-1-8k-base", config_path=cfg_path, model_dir=str(tmp_path), streaming=True)
    b = load_checkpoint("evo-1-8k-base", config_path=cfg_path, model_dir=str(tmp_path), streaming=False)
    sa, sb = a.state_dict(), b.state_dict()
    assert set(sa) == set(sb) == set(O.state_dict_spec(cfg))
    for k in sa:
        assert sa[k].dtype == sb[k].dtype and torch.equal(sa[k], sb[k]), k
    assert a.ingest_stats["tensors"] == len(sa)                       # every key incl. the tied unembed was accounted for
    mlp = a.blocks[0].mlp
    assert torch.equal(mlp.w12, b.blocks[0].mlp.w12) and torch.equal(mlp.w3, b.blocks[0].mlp.w3)
    assert mlp.w3[:, mlp.inner:].abs().sum() == 0
    g = mlp.w12.view(mlp.ipad // 128, 2, 128, -1)
    assert g[:, 0].reshape(mlp.ipad, -1)[mlp.inner:].abs().sum() == 0 and g[:, 1].reshape(mlp.ipad, -1)[mlp.inner:].abs().sum() == 0
    # the raw reader sees what safetensors wrote
    seen = {n: (dt, shape) for f in shard_files(str(tmp_path)) for n, dt, shape, raw in iter_safetensors(f)}
    assert set(seen) == {"backbone." + k for k in names}
    assert seen["backbone.norm.scale"][0] == torch.float32 and seen["backbone.blocks.1.inner_mha_cls.Wqkv.weight"][0] == torch.bfloat16
    # strict mode: a missing tensor, an unexpected tensor and a wrong shape all fail loudly
    from safetensors.torch import save_file
    for tag, edit in (("missing", lambda d: d.pop("blocks.0.mlp.l2.weight")), ("extra", lambda d: d.__setitem__("blocks.0.bogus", torch.zeros(3))),
                      ("shape", lambda d: d.__setitem__("blocks.0.filter.D", torch.zeros(7, dtype=torch.bfloat16)))):
        bad = tmp_path / tag
        bad.mkdir()
        d = {k: sd[k] for k in names}
        edit(d)
        save_file({"backbone." + k: v.contiguous() for k, v in d.items()}, str(bad / "model.safetensors"))
        with pytest.raises(RuntimeError, match="Missing key|Unexpected key|size mismatch"):
            load_checkpoint("evo-1-8k-base", config_path=cfg_path, model_dir=str(bad))


def test_untied_model_without_unembed_in_the_checkpoint_is_filled_from_the_embedding(tmp_path):
    """evo/models.py:133-137 copies embedding_layer.weight into a missing unembed.weight whatever tie_embeddings says; both ingest
    paths must do the same (the Evo configs are tied, so this is the edge the reference's unconditional copy creates)."""
    from safetensors.torch import save_file
    cfg = O.tiny_config(num_layers=2, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    cfg["tie_embeddings"] = False
    sd = O.random_state_dict(cfg, seed=3)
    sd.pop("unembed.weight")
    save_file({"backbone." + k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "model.safetensors"))
    (tmp_path / "c.yml").write_text(yaml.safe_dump(cfg))
    a = load_checkpoint("evo-1-8k-base", config_path=str(tmp_path / "c.yml"), model_dir=str(tmp_path), streaming=True)
    b = load_checkpoint("evo-1-8k-base", config_path=str(tmp_path / "c.yml"), model_dir=str(tmp_path), streaming=False)
    assert a.unembed is not a.embedding_layer
    sa, sb = a.state_dict(), b.state_dict()
    assert set(sa) == set(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    assert torch.equal(sa["unembed.weight"], sa["embedding_layer.weight"])


def test_checkpoint_without_inv_freq_keys_loads_with_the_analytic_buffer(tmp_path):
    """flash_attn registers rotary_emb.inv_freq as a non-persistent buffer; that stripedhyena re-registers it persistently (so that
    the HF checkpoints carry it) is recalled, not verified.  Both ingest paths therefore accept a checkpoint with
    or without those keys: present -> the checkpoint's values (see test_streaming_ingest...), absent -> the analytic table."""
    from safetensors.torch import save_file
    cfg = O.tiny_config(num_layers=2, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    sd = O.random_state_dict(cfg, seed=3)
    sd.pop("unembed.weight")
    dropped = [k for k in sd if k.endswith("rotary_emb.inv_freq")]
    assert dropped == ["blocks.1.inner_mha_cls.rotary_emb.inv_freq"]
    for k in dropped:
        sd.pop(k)
    save_file({"backbone." + k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "model.safetensors"))
    (tmp_path / "c.yml").write_text(yaml.safe_dump(cfg))
    for streaming in (True, False):
        m = load_checkpoint("evo-1-8k-base", config_path=str(tmp_path / "c.yml"), model_dir=str(tmp_path), streaming=streaming)
        rot = m.blocks[1].inner_mha_cls.rotary_emb
        want = 1.0 / (10000 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))
        assert rot.inv_freq.dtype == torch.float32 and torch.equal(rot.inv_freq, want)
        assert "blocks.1.inner_mha_cls.rotary_emb.inv_freq" in m.state_dict()          # still exported under the reference's key
    # any OTHER missing key is still a strict-mode error
    sd.pop("blocks.1.inner_mha_cls.out_proj.bias")
    save_file({"backbone." + k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "model.safetensors"))
    for streaming in (True, False):
        with pytest.raises(RuntimeError, match="Missing key"):
            load_checkpoint("evo-1-8k-base", config_path=str(tmp_path / "c.yml"), model_dir=str(tmp_path), streaming=streaming)


def test_mlp_parameters_live_only_in_the_packed_layouts():
    """No second copy of the MLP weights: the module owns w12 / w3, the reference's key names exist at the state-dict boundary only."""
    cfg = O.tiny_config(num_layers=1, attn_layer_idxs=(), hidden_size=256, num_heads=2)
    sd = O.random_state_dict(cfg, seed=1)
    m = StripedHyena(dotdict(cfg))
    m.load_state_dict(sd, strict=True)
    own = dict(m.blocks[0].mlp.named_parameters())
    assert set(own) == {"w12", "w3"}
    out = m.state_dict()
    for k in ("l1", "l2", "l3"):
        assert torch.equal(out[f"blocks.0.mlp.{k}.weight"], sd[f"blocks.0.mlp.{k}.weight"].float())
    assert "blocks.0.mlp.w12" not in out
    bad = dict(sd)
    bad.pop("blocks.0.mlp.l3.weight")
    with pytest.raises(RuntimeError, match="mlp.l3.weight"):
        StripedHyena(dotdict(cfg)).load_state_dict(bad, strict=True)


def test_frontend_fasta_reader_and_length_buckets(tmp_path):
    from evo_b200.frontend import length_buckets, read_fasta, read_prompts_csv
    fa = tmp_path / "x.fasta"
    fa.write_text(">s1 first\nACGT\nAC\n\n>s2\nTTTT\n>s3\nG\n")
    names, seqs = read_fasta(str(fa))
    assert names == ["s1", "s2", "s3"] and seqs == ["ACGTAC", "TTTT", "G"]
    csvf = tmp_path / "p.csv"
    csvf.write_text("prompt,other\nACGT,1\nGG,2\n")
    assert read_prompts_csv(str(csvf)) == ["ACGT", "GG"]
    seqs = ["A" * n for n in (5, 9, 5, 5, 100, 9, 5, 101, 7)]
    exact = length_buckets(seqs, batch_size=3, mode="exact")
    assert exact == [[0, 2, 3], [6], [1, 5], [4], [7], [8]]            # read_prompts' grouping (semantic_design.py:82-100)
    assert all(len({len(seqs[i]) for i in b}) == 1 for b in exact)
    srt = length_buckets(seqs, batch_size=4, mode="sorted", max_tokens=250, max_waste=0.25)
    assert sorted(i for b in srt for i in b) == list(range(len(seqs)))
    for b in srt:
        width = max(len(seqs[i]) for i in b)
        assert len(b) <= 4 and (len(b) == 1 or width * len(b) <= 250)
        assert 1.0 - sum(len(seqs[i]) for i in b) / (width * len(b)) <= 0.25 + 1e-9
    with pytest.raises(ValueError):
        length_buckets(seqs, mode="nope")


@pytest.mark.timeout(240)
def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours): exactly one JSON line on stdout carrying the
    contract's keys, a cpu_baseline that describes the run, and an e2e that repeats the line's own value with no copies."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=220, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "nt/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and abs(cb["value"] - d["value"]) < 1e-6
    assert d["e2e"] == {"value": d["value"], "unit": "nt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_cpu_leg_switches_to_the_reference_package_when_it_is_importable(monkeypatch):
    """bench.py's CPU legs (cpu_baseline, --impl reference) time the reference's own package instead of the oracle port whenever
    `import stripedhyena` works (kind "reference"); absent -- everywhere so far -- they time the port.  Exercised with a stand-in
    package and the geometry shrunk, so that the switch itself is tested, not the package."""
    import importlib.util
    from tests.test_oracle import _install_stand_in_package
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    tiny = O.tiny_config(num_layers=3, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    monkeypatch.setattr(O, "evo_config", lambda name="evo-1-8k-base": dict(tiny))
    base, _, run = bench.cpu_baseline("evo-1-8k-base", target_seconds=0.3, threads=2)
    assert base["kind"] == "port" and base["sample"].startswith("oracle (stripedhyena 0.2.2 restatement)")
    _install_stand_in_package(monkeypatch)
    base, _, run = bench.cpu_baseline("evo-1-8k-base", target_seconds=0.3, threads=2)
    assert base["kind"] == "reference" and base["sample"].startswith("stripedhyena stand-in") and base["value"] > 0
    assert run(32) > 0


def test_bench_helpers_assemble_the_contract_fields():
    """bench.py's pure helpers: synthetic inputs are reproducible per seed, the roofline records carry the contract's keys and
    consistent arithmetic (achieved = work / time, frac = achieved / peak), and the workloads are the BASELINE.json configs."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    a1, a2, a3 = b.synthetic_seqs(2, 50, seed=1), b.synthetic_seqs(2, 50, seed=1), b.synthetic_seqs(2, 50, seed=2)
    assert a1 == a2 and a1 != a3 and set("".join(a1)) <= set("ACGT") and [len(s) for s in a1] == [50, 50]
    assert b.WORKLOADS["8k"]["batch"] * b.WORKLOADS["8k"]["nt"] == 65536 and b.WORKLOADS["131k"]["nt"] == 131072 and b.WORKLOADS["gen"]["batch"] == 16
    peaks = {"hbm_gbs": 6000.0, "bf16_tflops": 1600.0, "bf16_tflops_sustained": 1400.0, "source": "measured"}
    by = {"gemm": [7.0e14, 500.0, 129.0], "hyena": [2.0e9 * 29, 29.0, 29.0], "attn": [1.4e13, 20.0, 3.0], "rmsnorm": [1.0, 10.0, 65.0]}
    r = b.rooflines(by, 600.0, "8k", peaks)
    g, h, at = r["roofline"], r["roofline_hyena"], r["roofline_attn"]
    for rec, bound, unit in ((g, "tensor", "TFLOP/s"), (h, "hbm", "GB/s"), (at, "tensor", "TFLOP/s")):
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in rec
        assert rec["bound"] == bound and rec["unit"] == unit and abs(rec["frac"] - rec["achieved"] / rec["peak"]) < 1e-12
    assert abs(g["achieved"] - 7.0e14 / 0.5 / 1e12) < 1e-6 and abs(h["achieved"] - 2.0e9 * 29 / 0.029 / 1e9) < 1e-6
    assert abs(g["share_of_step"] - 500.0 / 600.0) < 1e-12 and g["launches_per_step"] == 129.0


def test_cli_scripts_keep_the_reference_arguments():
    """scripts/score.py and scripts/generate.py: the reference's argument names and defaults (scripts/score.py:24-30,
    scripts/generate.py:24-37 of evo-design/evo), plus the offline switches; the example FASTA parses to three sequences."""
    import scripts.generate as g
    import scripts.score as s
    from evo_b200.frontend import read_fasta
    a = vars(s.build_parser().parse_args(["--input-fasta", "in.fa", "--output-tsv", "out.tsv"]))
    assert {k: a[k] for k in ("input_fasta", "output_tsv", "model_name", "batch_size", "device")} == \
        {"input_fasta": "in.fa", "output_tsv": "out.tsv", "model_name": "evo-1-131k-base", "batch_size": 32, "device": "cuda:0"}
    b = vars(g.build_parser().parse_args([]))
    want = {"model_name": "evo-1-131k-base", "prompt": "ACGT", "n_samples": 3, "n_tokens": 100, "temperature": 1.0, "top_k": 4, "top_p": 1.0,
            "cached_generation": True, "batched": True, "prepend_bos": False, "device": "cuda:0", "verbose": 1}
    assert {k: b[k] for k in want} == want
    with pytest.raises(SystemExit):
        s.build_parser().parse_args([])                       # both paths are required, as in the reference
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names, seqs = read_fasta(os.path.join(root, "examples", "example_seqs.fasta"))
    assert names == ["seq0", "seq1", "seq2"] and [len(x) for x in seqs] == [4, 11, 32] and set("".join(seqs)) <= set("ACGT")


def test_die_classifier_on_synthetic_latencies(tmp_path):
    """csrc/die_classify.h (the host half of the SM -> die calibration): recovers a 70 / 78 split from latencies with per-SM and
    per-line offsets, clock drift and noise; refuses to answer without a two-group structure or when a CTA pair straddles TPCs."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include "die_classify.h"
#include <cstdio>
#include <cstdlib>
#include <random>
int main(int argc, char** argv) {
  const int sms = 148, P = 24;
  const double gap = atof(argv[1]), noise = atof(argv[2]);
  const int bad_pair = atoi(argv[3]);
  std::mt19937 rng(atoi(argv[4]));
  std::normal_distribution<double> nd(0.0, 1.0);
  std::vector<int> truth(sms);
  for (int t = 0; t < 74; ++t) { int d = (t < 18 || (t >= 37 && t < 54)) ? 0 : 1; truth[2 * t] = truth[2 * t + 1] = d; }
  std::vector<int> home(P); for (int p = 0; p < P; ++p) home[p] = rng() & 1;
  std::vector<float> lat(sms * P);
  for (int s = 0; s < sms; ++s) {
    const double off = 8 * nd(rng), scale = 1.0 + 0.05 * s / sms;
    for (int p = 0; p < P; ++p) lat[s * P + p] = (float)(scale * (234 + off + 3 * std::sin(p) + (truth[s] != home[p] ? gap : 0.0)) + noise * nd(rng));
  }
  std::vector<unsigned> where(sms);
  for (int b = 0; b < sms; ++b) where[b] = (unsigned)((b + 142) % sms) | ((b & 1) << 16);
  if (bad_pair) std::swap(where[3], where[5]);
  std::vector<int> die;
  const char* why = evo::classify_dies(lat, sms, P, where, die);
  if (why) { printf("REJECT %s\n", why); return 0; }
  int wrong = 0, n0 = 0; for (int s = 0; s < sms; ++s) { wrong += die[s] != (truth[s] ^ truth[0]); n0 += die[s] == 0; }
  printf("OK wrong=%d die0=%d\n", wrong, n0);
}
''')
    exe = tmp_path / "t"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "evo_b200", "csrc"), str(src), "-o", str(exe)], check=True)
    run = lambda *a: subprocess.run([str(exe), *map(str, a)], capture_output=True, text=True, check=True).stdout.strip()
    for seed in (1, 2, 3):
        assert run(28, 2, 0, seed) == "OK wrong=0 die0=70"
    assert run(12, 3, 0, 4) == "OK wrong=0 die0=70"
    assert run(0, 2, 0, 1).startswith("REJECT no dominant")
    assert run(28, 25, 0, 1).startswith("REJECT")
    assert run(28, 2, 1, 1).startswith("REJECT a CTA pair")
