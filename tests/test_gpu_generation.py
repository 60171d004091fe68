"""Device-side sampler and on-device generation loop against the host implementations that
mirror stripedhyena.sample.sample and the reference's per-token loop (evo/generation.py:131-189).

Bars: greedy picks and everything the loop records (tokens, logits, final positions) BIT-EXACT against the
per-token host loop; stochastic picks are a different RNG stream by construction (the reference uses torch's global
generator), so they are held to the same kept set and to the same distribution (total-variation bound over 40 000
draws) as oracle.sample's filter pipeline."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

from oracle import stripedhyena_oracle as O          # noqa: E402
from evo_b200 import _lib                             # noqa: E402
from evo_b200.stripedhyena import StripedHyena, dotdict  # noqa: E402

DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    _lib.lib()


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dev_sample(logits, top_k, top_p, temperature, seed=1, step=0):
    out = torch.empty(logits.shape[0], dtype=torch.long, device=DEV)
    lg = logits.to(DEV).contiguous()
    _lib.check(_lib.lib().evo_sample(_lib.ptr(lg), _lib.ptr(out), lg.shape[0], lg.shape[1], top_k, top_p, temperature, seed, step, stream()), "evo_sample")
    return out.cpu()


def test_sampler_greedy_is_first_argmax():
    torch.manual_seed(0)
    lg = (torch.randn(67, 512) * 4).bfloat16()
    lg[3, 100] = lg[3, 400] = lg[3].max() + 1          # a tie: the first maximum wins (torch.argmax on CPU)
    lg[5] = 0
    got = dev_sample(lg, 1, 0.0, 1.0)
    assert torch.equal(got, lg.float().argmax(-1))
    assert got[3].item() == 100 and got[5].item() == 0
    small = (torch.randn(9, 40) * 2).bfloat16()         # V not a multiple of 32
    assert torch.equal(dev_sample(small, 1, 0.0, 1.0), small.float().argmax(-1))


def host_pipeline_probs(row, top_k, top_p, temperature):
    """Probability vector over the vocabulary that oracle.sample's filter pipeline hands to torch.multinomial."""
    lg = row.clone()[None]
    V = lg.shape[-1]
    if top_k > 0:
        k = min(top_k, V)
        kept, idx = torch.topk(lg, k, dim=-1)
        if temperature != 1.0:
            kept = kept / temperature
        O._top_p_filter(kept, top_p)
        p = torch.zeros(V)
        p[idx[0]] = torch.softmax(kept.float(), -1)[0]
        return p
    lt = lg / temperature if temperature != 1.0 else lg.clone()
    O._top_p_filter(lt, top_p)
    return torch.softmax(lt.float(), -1)[0]


@pytest.mark.parametrize("top_k,top_p,temperature", [(4, 1.0, 1.0), (50, 0.73, 1.0), (8, 0.9, 1.25), (0, 0.5, 1.0), (0, 1.0, 1.3), (512, 0.0, 1.0)])
def test_sampler_distribution_matches_host_pipeline(top_k, top_p, temperature):
    torch.manual_seed(3)
    row = torch.randn(512).bfloat16()
    row[[65, 67, 71, 84]] += torch.tensor([4.0, 3.5, 3.0, 4.5]).bfloat16()     # an ACGT-like head over a flat tail, like real Evo logits
    want = host_pipeline_probs(row, top_k, top_p, temperature).double()
    N = 40000
    draws = dev_sample(row[None].expand(N, -1), top_k, top_p, temperature, seed=12345)
    got = torch.bincount(draws, minlength=512).double() / N
    support = want > 0
    # the kept set: nothing outside the host's support, except entries sitting on the top-p boundary (bf16 cumsum vs fp32 suffix sums)
    outside = got[~support].sum().item()
    assert outside <= 2e-3, outside
    tv = 0.5 * (got - want).abs().sum().item()
    noise = 0.5 * 0.8 * torch.sqrt(want * (1 - want) / N).sum().item()      # E|binomial deviation| summed over the support
    assert tv <= 2.0 * noise + 0.01, (tv, noise)
    assert int(support.sum()) >= 2                       # the case really samples
    # a second seed gives different draws, the same seed the same draws
    assert torch.equal(draws, dev_sample(row[None].expand(N, -1), top_k, top_p, temperature, seed=12345))
    assert not torch.equal(draws, dev_sample(row[None].expand(N, -1), top_k, top_p, temperature, seed=54321))


def _tiny(layers=4, attn=(1, 3), seed=7):
    cfg = O.tiny_config(num_layers=layers, attn_layer_idxs=attn, hidden_size=256, num_heads=2)
    sd = O.random_state_dict(cfg, seed=seed)
    m = StripedHyena(dotdict(cfg))
    m.load_state_dict(sd, strict=True)
    m.to_bfloat16_except_poles_residues()
    return cfg, sd, m.to(DEV)


@pytest.mark.parametrize("threshold", [128, 5])
def test_device_loop_equals_per_token_host_loop_greedy(threshold):
    """Generator.generate with the loop on the GPU vs the reference-shaped per-token loop: same tokens, same logits, same
    final positions -- with a full prefill (threshold >= prompt) and with a teacher-forced prompt tail (threshold 5)."""
    import evo_b200
    from evo_b200.generation import Generator
    cfg, sd, m = _tiny()
    tok = evo_b200.CharLevelTokenizer(512)
    ids = torch.tensor([tok.tokenize("ACGTTGCAACGTAC"), tok.tokenize("TTGACCAAGGTCAT")], dtype=torch.long, device=DEV)
    outs = []
    for loop in (True, False):
        g = Generator(m, tok, top_k=1, top_p=1.0, temperature=1.0)
        g.device_loop = loop
        m._decode = m._loop = None
        toks, lgs, st = g.generate(device=DEV, input_ids=ids, num_tokens=24, cached_generation=True, force_prompt_threshold=threshold,
                                   print_generation=False, verbose=False, stop_at_eos=False)
        outs.append((toks.cpu(), lgs.cpu(), st["mha"].seqlen_offset, st["hyena"].seqlen_offset))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    assert outs[0][2:] == outs[1][2:]
    assert outs[0][0].shape == (2, 24) and outs[0][1].shape == (2, 24, 512)


def test_device_loop_resume_and_reuse_of_the_captured_graph():
    """A second generate() on the same state object continues where the first stopped (evo/generation.py:105-114) and reuses the
    captured step; the continuation equals one long generation."""
    import evo_b200
    from evo_b200.generation import Generator
    cfg, sd, m = _tiny()
    tok = evo_b200.CharLevelTokenizer(512)
    ids = torch.tensor([tok.tokenize("ACGTTGCAACGTACGG")], dtype=torch.long, device=DEV)
    g = Generator(m, tok, top_k=1)
    m._decode = m._loop = None
    long_t, long_l, _ = g.generate(device=DEV, input_ids=ids, num_tokens=20, cached_generation=True, force_prompt_threshold=64, print_generation=False, stop_at_eos=False)
    m._loop = None
    a_t, a_l, st = g.generate(device=DEV, input_ids=ids, num_tokens=8, cached_generation=True, force_prompt_threshold=64, print_generation=False, stop_at_eos=False)
    graph = m._loop["graph"]
    # resume: the caller feeds the last generated token; the reference then steps from seqlen_offset + 1
    b_t, b_l, st = g.generate(device=DEV, input_ids=a_t[:, -1:], num_tokens=12, cached_generation=True, print_generation=False, stop_at_eos=False,
                              inference_params_dict=st)
    assert m._loop["graph"] is graph                       # same state tensors -> same captured step
    assert torch.equal(torch.cat([a_t, b_t], 1), long_t)
    assert torch.equal(torch.cat([a_l, b_l], 1), long_l)


def test_device_loop_sampling_is_reproducible_under_manual_seed():
    import evo_b200
    cfg, sd, m = _tiny(layers=3, attn=(1,))
    tok = evo_b200.CharLevelTokenizer(512)
    prompts = ["ACGTACGT", "TTGACCAA"]
    kw = dict(n_tokens=16, top_k=4, top_p=0.95, temperature=1.0, cached_generation=True, verbose=0, device=DEV)
    torch.manual_seed(11)
    a, sa = evo_b200.generate(prompts, m, tok, **kw)
    torch.manual_seed(11)
    b, sb = evo_b200.generate(prompts, m, tok, **kw)
    torch.manual_seed(12)
    c, _ = evo_b200.generate(prompts, m, tok, **kw)
    assert a == b and sa == sb
    assert a != c
    assert all(len(x) == 16 for x in a)
