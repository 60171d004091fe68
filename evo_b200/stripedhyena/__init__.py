"""Host-side mirror of the `stripedhyena` package surface that evo-design/evo imports
(evo/models.py:8-9, evo/scoring.py:5, evo/generation.py:6-7): StripedHyena, dotdict,
sample, InferenceParams / RecurrentInferenceParams.  Every FLOP runs in libevo_b200.so."""
from .utils import dotdict
from .cache import InferenceParams, RecurrentInferenceParams
from .sample import sample
from .model import StripedHyena

__all__ = ["StripedHyena", "dotdict", "sample", "InferenceParams", "RecurrentInferenceParams"]
