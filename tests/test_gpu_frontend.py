"""GPU side of the batch front-end and of the streaming checkpoint ingest: integer / byte work, so
everything here is BIT-EXACT against the host implementations that mirror the reference (evo/scoring.py:9-33
prepare_batch; evo/models.py:96-150 load sequence)."""
import json
import os
import sys

import numpy as np
import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

from oracle import stripedhyena_oracle as O          # noqa: E402
from evo_b200 import _lib, CharLevelTokenizer, prepare_batch  # noqa: E402
from evo_b200.frontend import device_batch, score_many        # noqa: E402
from evo_b200.models import load_checkpoint                   # noqa: E402

DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    _lib.lib()


def reference_prepare_batch(seqs, tok, prepend_bos):
    """evo/scoring.py:9-33, verbatim semantics, on the host."""
    lengths = [len(s) for s in seqs]
    width = max(lengths)
    rows = [([tok.eod_id] * int(prepend_bos)) + [int(t) for t in tok.tokenize(s)] + [tok.pad_id] * (width - len(s)) for s in seqs]
    return torch.tensor(rows, dtype=torch.long), lengths


@pytest.mark.parametrize("prepend_bos", [True, False])
@pytest.mark.parametrize("dtype", [torch.long, torch.int32])
def test_device_tokenise_pad_equals_reference_prepare_batch(prepend_bos, dtype):
    tok = CharLevelTokenizer(512)
    rng = np.random.default_rng(0)
    seqs = ["".join(rng.choice(list("ACGTN|~ "), size=n)) for n in (1, 17, 4096, 300, 1, 8192, 33)]
    want, wl = reference_prepare_batch(seqs, tok, prepend_bos)
    got, gl = device_batch(seqs, tok, prepend_bos=prepend_bos, device=DEV, dtype=dtype)
    assert gl == wl and got.dtype == dtype and got.device.type == "cuda"
    assert torch.equal(got.cpu().long(), want)
    ids, lens = prepare_batch(seqs, tok, prepend_bos=prepend_bos, device=DEV)      # the public entry point takes the same path
    assert torch.equal(ids.cpu(), want) and lens == wl
    one, _ = device_batch(["G"], tok, prepend_bos=prepend_bos, device=DEV)
    assert one.tolist() == ([[0, 71]] if prepend_bos else [[71]])
    with pytest.raises(ValueError):
        device_batch(["ACé"], tok, device=DEV)


def test_streaming_ingest_to_the_gpu_and_bucketed_scoring(tmp_path):
    from safetensors.torch import save_file
    cfg = O.tiny_config(num_layers=3, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    sd = O.random_state_dict(cfg, seed=9)
    sd.pop("unembed.weight")
    names = sorted(sd)
    files = {"model-00001-of-00002.safetensors": names[::2], "model-00002-of-00002.safetensors": names[1::2]}
    wm = {}
    for f, keys in files.items():
        save_file({"backbone." + k: (sd[k].float() if "mlp.l2" in k else sd[k]).contiguous() for k in keys}, str(tmp_path / f))
        wm.update({"backbone." + k: f for k in keys})
    (tmp_path / "model.safetensors.index.json").write_text(json.dumps({"weight_map": wm}))
    cfgp = tmp_path / "tiny.yml"
    cfgp.write_text(yaml.safe_dump(cfg))
    a = load_checkpoint("evo-1-8k-base", config_path=str(cfgp), model_dir=str(tmp_path), device=DEV, streaming=True)
    b = load_checkpoint("evo-1-8k-base", config_path=str(cfgp), model_dir=str(tmp_path), device=DEV, streaming=False)
    sa, sb = a.state_dict(), b.state_dict()
    assert set(sa) == set(sb)
    for k in sa:
        assert sa[k].device.type == "cuda" and sa[k].dtype == sb[k].dtype and torch.equal(sa[k], sb[k]), k
    assert torch.equal(a.blocks[0].mlp.w12, b.blocks[0].mlp.w12) and torch.equal(a.blocks[2].mlp.w3, b.blocks[2].mlp.w3)
    ids = (torch.randint(0, 4, (2, 70)) * 3 + 65).to(DEV)
    assert torch.equal(a(ids)[0], b(ids)[0])
    # bucketed scoring returns the same numbers as one-by-one scoring, in input order
    import evo_b200
    tok = CharLevelTokenizer(512)
    rng = np.random.default_rng(1)
    seqs = ["".join(rng.choice(list("ACGT"), size=n)) for n in (40, 7, 40, 129, 8, 41, 128)]
    many = score_many(seqs, a, tok, batch_size=3, max_tokens=400, device=DEV)
    single = [float(evo_b200.score_sequences([s], a, tok, device=DEV)[0]) for s in seqs]
    assert np.allclose(many, single, atol=2e-2)                   # padding changes nothing causal; batch shapes change bf16 GEMM tiling only
    assert len(many) == len(seqs)


def test_public_scoring_entry_points_take_the_device_front_end_and_the_fused_head(monkeypatch):
    """Routing, not numbers: score_sequences / positional_entropies must tokenise on the GPU (frontend.device_batch ->
    evo_tokenize_pad) and read their statistics from the fused head (model.score_tokens -> evo_unembed_score), never from
    materialised logits; the library's launch counter sees exactly the extra launches of that route."""
    import evo_b200
    from evo_b200 import frontend
    from evo_b200.stripedhyena import StripedHyena, dotdict
    cfg = O.tiny_config(num_layers=2, attn_layer_idxs=(1,), hidden_size=256, num_heads=2)
    m = StripedHyena(dotdict(cfg))
    m.load_state_dict(O.random_state_dict(cfg, seed=2), strict=True)
    m.to_bfloat16_except_poles_residues()
    m = m.to(DEV)
    calls = {"batch": 0, "score": 0, "forward": 0}
    real_batch, real_score, real_forward = frontend.device_batch, m.score_tokens, m.forward
    monkeypatch.setattr(frontend, "device_batch", lambda *a, **k: (calls.__setitem__("batch", calls["batch"] + 1), real_batch(*a, **k))[1])
    m.score_tokens = lambda *a, **k: (calls.__setitem__("score", calls["score"] + 1), real_score(*a, **k))[1]
    m.forward = lambda *a, **k: (calls.__setitem__("forward", calls["forward"] + 1), real_forward(*a, **k))[1]
    tok = CharLevelTokenizer(512)
    lib = _lib.lib()
    evo_b200.score_sequences(["ACGTACGTAC", "TTGACA"], m, tok, device=DEV)          # warm: the rope tables are built once
    lib.evo_reset_launch_count()
    s = evo_b200.score_sequences(["ACGTACGTAC", "TTGACA"], m, tok, device=DEV)
    n_score = lib.evo_launch_count()
    e = evo_b200.positional_entropies(["ACGTACGTAC", "TTGACA"], m, tok, device=DEV)
    assert calls == {"batch": 3, "score": 3, "forward": 0}
    assert len(s) == 2 and [len(x) for x in e] == [10, 6]
    # tokenise + embed + 2 blocks + final norm + fused head (GEMM + finish): the plain forward would end in one GEMM and a logprobs kernel
    lib.evo_reset_launch_count()
    ids, _ = prepare_batch(["ACGTACGTAC", "TTGACA"], tok, device=DEV)
    n_tok = lib.evo_launch_count()
    real_forward(ids)
    n_fwd = lib.evo_launch_count() - n_tok
    assert n_tok == 1 and n_score == n_tok + n_fwd + 1            # fused head = the forward's last GEMM + one finishing kernel


def test_cli_scripts_run_end_to_end_on_random_weights(tmp_path, capsys):
    """python -m scripts.score / scripts.generate / scripts.example_inference with --random-init (no network on the GPU box): the
    7B architecture is built on the device, the example FASTA is scored through the bucketed front-end, a short sample is generated."""
    import scripts.example_inference as ex
    import scripts.generate as gen
    import scripts.score as sc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tsv = tmp_path / "scores.tsv"
    scores = sc.main(["--input-fasta", os.path.join(root, "examples", "example_seqs.fasta"), "--output-tsv", str(tsv), "--model-name", "evo-1-8k-base",
                      "--device", DEV, "--random-init"])
    rows = tsv.read_text().strip().split("\n")
    assert rows[0] == "seqs\tscores" and len(rows) == 4 and len(scores) == 3 and all(np.isfinite(scores))
    assert [len(r.split("\t")[0]) for r in rows[1:]] == [4, 11, 32]
    torch.cuda.empty_cache()
    out, sc_ = gen.main(["--model-name", "evo-1-8k-base", "--prompt", "ACGTAC", "--n-samples", "2", "--n-tokens", "8", "--device", DEV, "--random-init", "--seed", "3", "--verbose", "0"])
    assert len(out) == 2 and all(len(o) == 8 for o in out) and len(sc_) == 2
    torch.cuda.empty_cache()
    logits = ex.main(["--model-name", "evo-1-8k-base", "--device", DEV, "--random-init"])
    assert tuple(logits.shape) == (3, 20, 512)
