"""StripedHyena hyper-parameters of the Evo checkpoints.  Same keys and values as the
reference's evo/configs/evo-1-8k-base_inference.yml (lines 1-38) and
evo-1-131k-base_inference.yml (lines 39-40 add the rotary interpolation); kept as Python
dicts so the package has no data files to locate.  A YAML path with the same keys is
accepted by load_checkpoint(config_path=...)."""
from __future__ import annotations

import copy

_ATTN = [8, 16, 24]

EVO_1_8K = {
    # geometry
    "vocab_size": 512, "hidden_size": 4096, "num_filters": 4096, "num_layers": 32,
    "num_attention_heads": 32, "max_sequence_len": 8192,
    "attn_layer_idxs": list(_ATTN),
    "hyena_layer_idxs": [i for i in range(32) if i not in _ATTN],
    # hyena operator
    "short_filter_length": 3, "short_filter_bias": True, "state_size": 8,
    "hyena_filter_groups": 1, "column_split": True, "split_k0": True, "prefill_style": "fft",
    # attention
    "proj_groups": 1, "smeared_gqa": False, "qkv_proj_bias": True, "mha_out_proj_bias": True,
    # mlp / norms / embeddings
    "inner_size_multiple_of": 16, "inner_mlp_size": None, "mlp_activation": "gelu",
    "mlp_init_method": "torch.nn.init.zeros_", "mlp_output_init_method": "torch.nn.init.zeros_",
    "eps": 1.0e-6, "final_norm": True, "tie_embeddings": True, "make_vocab_size_divisible_by": 8,
    # engine switches of the reference (read, and required to have these values)
    "use_flash_attn": True, "use_flash_rmsnorm": False, "use_flash_depthwise": False, "use_flashfft": False,
    "inference_mode": True, "log_intermediate_values": False, "rng_fork": False,
    "model_parallel_size": 1, "pile_parallel_size": 1, "tokenizer_type": "CharLevelTokenizer",
}

EVO_1_131K = dict(copy.deepcopy(EVO_1_8K), use_interpolated_rotary_pos_emb=True, rotary_emb_scaling_factor=16)

# model name -> (config, HuggingFace repo, revision); evo/models.py:13-19,65-71,92
MODELS = {
    "evo-1.5-8k-base": (EVO_1_8K, "evo-design/evo-1.5-8k-base", "main"),
    "evo-1-8k-base": (EVO_1_8K, "togethercomputer/evo-1-8k-base", "1.1_fix"),
    "evo-1-131k-base": (EVO_1_131K, "togethercomputer/evo-1-131k-base", "1.1_fix"),
    "evo-1-8k-crispr": (EVO_1_8K, "LongSafari/evo-1-8k-crispr", "main"),
    "evo-1-8k-transposon": (EVO_1_8K, "LongSafari/evo-1-8k-transposon", "main"),
}
MODEL_NAMES = list(MODELS)


def get_config(model_name: str) -> dict:
    if model_name not in MODELS:
        raise ValueError(f"Invalid model name {model_name}. Should be one of: {', '.join(MODEL_NAMES)}.")
    return copy.deepcopy(MODELS[model_name][0])
