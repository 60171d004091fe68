"""Rasterisation sweep of the big-tile GEMM (development tool): DRAM bytes and duration per (shape, grouping) under ncu.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:gemm_tcgen05 --csv --log-file gpurun_out/raster.csv python tools/gemm_raster_sweep.py --ncu
    python tools/gemm_raster_sweep.py --table gpurun_out/raster.csv        # joins the launch list with the sweep order

Without ncu the same script times every configuration with CUDA events (isolated, L2 flushed in between)."""
import ctypes as C
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

M = 65544
SHAPES = [  # name, N, K, epilogue id, algorithmic bytes (A + W + C [+ residual])
    ("l3", 4096, 11008, 3), ("gate", 22016, 4096, 4), ("proj", 12288, 4096, 1), ("out", 4096, 4096, 2)]
CONFIGS = {
    "l3": [None, (1, 2), (1, 4), (1, 8), (1, 16), (0, 2), (0, 4), (0, 9), (0, 19), (0, 37)],
    "gate": [None, (1, 4), (1, 9), (1, 43), (1, 86), (0, 4), (0, 9), (0, 19), (0, 37), (0, 74)],
    "proj": [None, (1, 4), (1, 9), (1, 24), (1, 48), (0, 4), (0, 9), (0, 19), (0, 37)],
    "out": [None, (1, 4), (1, 16), (0, 9), (0, 37)],
}
SKEWS = [0, 128, 512, 2048, 8192]       # --skew: default grouping, producer start skew in cycles per tile slot
if "--skew" in sys.argv:
    SHAPES = SHAPES[:3]
    CONFIGS = {name: [("skew", c) for c in SKEWS] for name, _, _, _ in SHAPES}
if "--die" in sys.argv:                  # --die: default grouping, die-oblivious vs die-aware tile walk
    CONFIGS = {name: [("die", 0), ("die", 1), ("die", 0), ("die", 1)] for name, _, _, _ in SHAPES}


def label(cfg):
    if cfg is None:
        return "default"
    if cfg[0] == "skew":
        return "skew=%d cycles/slot" % cfg[1]
    if cfg[0] == "die":
        return "die-aware walk" if cfg[1] else "die-oblivious walk"
    return "raster_n=%d group=%d" % cfg


def order():
    return [(name, cfg) for name, _, _, _ in SHAPES for cfg in CONFIGS[name]]


def algorithmic(name):
    _, N, K, epi = next(s for s in SHAPES if s[0] == name)
    n_out = N // 2 if epi == 4 else N
    return 2.0 * (M * K + N * K + M * n_out + (M * n_out if epi in (2, 3) else 0))


def table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    while rows[0][0] != "ID":
        rows.pop(0)
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        if not r[ix["ID"]].isdigit():
            continue
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[r[ix["Metric Unit"]]]
        per.setdefault(int(r[ix["ID"]]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * scale
    ids = sorted(per)
    assert len(ids) == len(order()), (len(ids), len(order()))
    for i, (name, cfg) in zip(ids, order()):
        m = per[i]
        b = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        print(f"{name:5s} {label(cfg):22s} {b / 1e9:7.2f} GB = {b / algorithmic(name):5.2f}x algorithmic   {m['gpu__time_duration.sum']:9.1f} us")


def main():
    import torch
    from evo_b200 import _lib
    lib = _lib.lib()
    dev = "cuda:0"
    under_ncu = "--ncu" in sys.argv
    torch.manual_seed(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for name, N, K, epi in SHAPES:
        n_out = N // 2 if epi == 4 else N
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) / 64).bfloat16()
        bias = torch.zeros(N, dtype=torch.bfloat16, device=dev)
        resid = torch.zeros(M, n_out, dtype=torch.bfloat16, device=dev)
        out = torch.empty(M, n_out, dtype=torch.bfloat16, device=dev)
        p = _lib.GemmParams(A=a.data_ptr(), lda=K, W=w.data_ptr(), C=out.data_ptr(), ldc=n_out, bias=bias.data_ptr() if epi in (1, 2) else None,
                            residual=resid.data_ptr() if epi in (2, 3) else None, ldr=n_out, M=M, N=N, K=K, epilogue=epi, variant=0)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for cfg in CONFIGS[name]:
            for k in ("EVO_B200_GEMM_RASTER_N", "EVO_B200_GEMM_GROUP", "EVO_B200_GEMM_SKEW"):
                os.environ.pop(k, None)
            if cfg is not None and cfg[0] == "skew":
                os.environ["EVO_B200_GEMM_SKEW"] = str(cfg[1])
            elif cfg is not None and cfg[0] == "die":
                os.environ["EVO_B200_GEMM_DIE_RASTER"] = str(cfg[1])
            elif cfg is not None:
                os.environ["EVO_B200_GEMM_RASTER_N"], os.environ["EVO_B200_GEMM_GROUP"] = str(cfg[0]), str(cfg[1])
            if under_ncu:
                _lib.check(lib.evo_gemm(C.byref(p), st), "evo_gemm")
                continue
            ts = []
            for _ in range(4):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.evo_gemm(C.byref(p), st), "evo_gemm")
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts[1:])[1]
            print(f"{name:5s} {label(cfg):22s} {t * 1e3:9.1f} us  {2.0 * M * N * K / t / 1e9:7.1f} TFLOP/s", flush=True)
        del a, w, resid, out
    torch.cuda.synchronize()


if __name__ == "__main__":
    if "--table" in sys.argv:
        table(sys.argv[sys.argv.index("--table") + 1])
    else:
        main()
