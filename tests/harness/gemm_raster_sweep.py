"""Sweep the GEMM rasterisation knobs (experiments): each setting runs in its own process because the
library reads EVO_B200_GEMM_GROUP / EVO_B200_GEMM_RASTER_N once.   python tests/harness/gemm_raster_sweep.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import gpu_bringup as G
    G._imports()
    import torch
    dev = "cuda:0"
    out = {}
    for (M, N, K, epi) in ((65544, 12288, 4096, 1), (65544, 4096, 4096, 2), (65544, 22016, 4096, 4), (65544, 4096, 11008, 3)):
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) / 64).bfloat16()
        bias = torch.randn(N, device=dev).bfloat16()
        nout = N // 2 if epi == 4 else N
        resid = torch.randn(M, nout, device=dev).bfloat16()
        for variant in (1, 0):
            ms = G.timeit(lambda: G._gemm(a, w, M, N, K, epi, variant, bias=bias, resid=resid, ldc=nout), iters=12, warm=3)
            out[f"{N}x{K}/e{epi}/v{variant}"] = round(2.0 * M * N * K / ms / 1e9, 1)
    print(json.dumps(out))
else:
    settings = [{}, {"EVO_B200_GEMM_RASTER_N": "0", "EVO_B200_GEMM_GROUP": "16"}, {"EVO_B200_GEMM_RASTER_N": "0", "EVO_B200_GEMM_GROUP": "32"},
                {"EVO_B200_GEMM_RASTER_N": "0", "EVO_B200_GEMM_GROUP": "64"}, {"EVO_B200_GEMM_RASTER_N": "1", "EVO_B200_GEMM_GROUP": "8"},
                {"EVO_B200_GEMM_RASTER_N": "1", "EVO_B200_GEMM_GROUP": "16"}, {"EVO_B200_GEMM_RASTER_N": "1", "EVO_B200_GEMM_GROUP": "32"}]
    for s in settings:
        env = dict(os.environ, **s)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=300)
        print(json.dumps({"setting": s or "heuristic", "tflops": json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-300:]}), flush=True)
