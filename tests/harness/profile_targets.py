"""Small single-kernel drivers for ncu captures (development tool).

    ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 2 -o gpurun_out/gemm python tests/harness/profile_targets.py gemm1
"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_bringup as G

G._imports()
import torch  # noqa: E402

what = sys.argv[1]
dev = "cuda:0"
if what in ("gemm1", "gemm2"):
    variant = 1 if what == "gemm1" else 0
    M, N, K = 16384, 12288, 4096
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) / 64).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16()
    for _ in range(3):
        G._gemm(a, w, M, N, K, 1, variant, bias=bias)
    torch.cuda.synchronize()
elif what == "gemm_gate":
    M, N, K = 16384, 22016, 4096
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) / 64).bfloat16()
    for _ in range(3):
        G._gemm(a, w, M, N, K, 4, 1, ldc=N // 2)
    torch.cuda.synchronize()
elif what == "hyena":
    from oracle import stripedhyena_oracle as O
    D, H, B, L = 4096, 32, 8, 8193
    cfg = O.tiny_config(num_layers=1, attn_layer_idxs=(), hidden_size=D, num_heads=H)
    sd = O.random_state_dict(cfg, seed=1)
    p = "blocks.0.filter."
    f = {"w": sd[p + "short_filter_weight"].to(dev), "b": sd[p + "short_filter_bias"].to(dev), "D": sd[p + "D"].to(dev), "p": sd[p + "poles"].to(dev), "r": sd[p + "residues"].to(dev)}
    z = torch.randn(B, L, 3 * D, device=dev).bfloat16()
    for _ in range(3):
        G._hyena_call(z, f, B, L, D, H, want_state=False)
elif what == "attn2":
    B, L, H = 2, 8193, 32
    qkv = torch.randn(B, L, 3, H, 128, device=dev).bfloat16()
    for _ in range(3):
        G._attn(qkv, B, L, H, 2)
elif what == "attn":
    B, L, H = 2, 8193, 32
    qkv = torch.randn(B, L, 3, H, 128, device=dev).bfloat16()
    for _ in range(3):
        G._attn(qkv, B, L, H, 1)
print("done", what)
