"""Token sampling with the semantics of stripedhyena.sample.sample (call site
evo/generation.py:162-167): greedy when top_k == 1, else top-k -> temperature -> top-p
-> multinomial.  512-way logits per row: host-side torch ops, not a hot path.
Every op runs in the logits' own dtype, like the code it mirrors (stripedhyena/sample.py is flash_attn's
utils/generation.py:70-98): the same logits and the same torch seed give the same picks -- checked against flash_attn's function
and, through the reference's own generation loop, against fixtures made by running it (tests/golden/)."""
import torch


def _mask_top_p(scores: torch.Tensor, top_p: float) -> None:
    """Drop the low-probability tail whose cumulative mass is <= 1 - top_p (in place)."""
    if top_p <= 0.0 or top_p >= 1.0:
        return
    asc, order = torch.sort(scores, descending=False)
    tail = asc.softmax(dim=-1).cumsum(dim=-1) <= (1.0 - top_p)
    scores.masked_fill_(tail.scatter(1, order, tail), float("-inf"))


def sample(logits: torch.Tensor, top_k: int = 1, top_p: float = 0.0, temperature: float = 1.0) -> torch.Tensor:
    if logits.dim() == 3:
        logits = logits.squeeze(1)
    if top_k == 1:
        return logits.argmax(dim=-1)
    if top_p > 0.0 and top_p > 1.0:
        raise ValueError("top-p should be in (0, 1]")
    rows = torch.arange(logits.shape[0], device=logits.device)
    if top_k > 0:
        k = min(top_k, logits.size(-1))
        kept, kept_idx = torch.topk(logits, k, dim=-1)
        if temperature != 1.0:
            kept = kept / temperature
        _mask_top_p(kept, top_p)
        choice = torch.multinomial(torch.softmax(kept, dim=-1), num_samples=1).squeeze(-1)
        return kept_idx[rows, choice]
    scaled = logits / temperature if temperature != 1.0 else logits.clone()
    _mask_top_p(scaled, top_p)
    return torch.multinomial(torch.softmax(scaled, dim=-1), num_samples=1).squeeze(-1)
