"""evo_b200: B200-native StripedHyena forward engine behind the evo-design/evo surface
(reference evo/__init__.py:3-6 exports Evo, generate, score_sequences, positional_entropies)."""
__version__ = "0.1.0"

from .models import Evo, load_checkpoint
from .generation import generate, Generator
from .scoring import score_sequences, positional_entropies, prepare_batch, logits_to_logprobs
from .tokenizer import CharLevelTokenizer

__all__ = ["Evo", "load_checkpoint", "generate", "Generator", "score_sequences", "positional_entropies",
           "prepare_batch", "logits_to_logprobs", "CharLevelTokenizer"]
