# die-aware rasterisation of the big-tile GEMM: (1) ground truth: does the second die go to DRAM on its own? (2) bit-identity test,
# (3) DRAM bytes per launch with the die-oblivious / die-aware walk under ncu, (4) in-step interleaved A/B
timeout 200 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,gpu__time_duration.sum -k regex:touch --csv \
  --log-file gpurun_out/r02_die_probe.csv tools/microbench/die_probe > gpurun_out/r02_die_probe.log 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "die_aware" 2>&1 | tail -3
export EVO_B200_GEMM_DIE_DUMP=gpurun_out/r02_die_map.txt
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv \
  --log-file gpurun_out/r02_die_sweep.csv python tools/gemm_raster_sweep.py --ncu --die > gpurun_out/r02_die_ncu.log 2>&1
python tools/gemm_raster_sweep.py --die --table gpurun_out/r02_die_sweep.csv | tee gpurun_out/r02_die_sweep.txt
head -1 gpurun_out/r02_die_map.txt
unset EVO_B200_GEMM_DIE_DUMP
for r in 0 1 0 1; do EVO_B200_GEMM_DIE_RASTER=$r timeout 200 python bench.py --steps 12 --warmup 3 --no-sub --no-gen --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/r02_bench_8k_die$r.json; python -c "import json; d=json.load(open('gpurun_out/r02_bench_8k_die$r.json')); print('die_raster=$r', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" | tee -a gpurun_out/r02_die_sweep.txt; done
