# ncu evidence for round 2 (1 GPU).  Numbers printed under ncu are never bench values.
set -x
ncu --set full --clock-control none --import-source on -k regex:hyena_scan_ms -s 2 -c 1 -o gpurun_out/r02_ncu_hyena_ms python tests/harness/profile_targets.py hyena > gpurun_out/ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 260 --csv --log-file gpurun_out/r02_launches_bench8k.csv python bench.py --steps 1 --warmup 1 --no-sub --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none -k regex:"gemm_tcgen05|hyena_scan|rmsnorm" -s 238 -c 9 -o gpurun_out/r02_ncu_block python bench.py --steps 1 --warmup 1 --no-sub --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none -k regex:"attn_pp" -s 3 -c 1 -o gpurun_out/r02_ncu_attn python bench.py --steps 1 --warmup 1 --no-sub --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log
