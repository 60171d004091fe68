# big-tile GEMM: de-phased producers (tile slot t starts t * skew cycles late) -- DRAM bytes per launch under ncu, then in-step A/B
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv \
  --log-file gpurun_out/r02_skew_sweep.csv python tools/gemm_raster_sweep.py --ncu --skew > gpurun_out/r02_skew_ncu.log 2>&1
python tools/gemm_raster_sweep.py --skew --table gpurun_out/r02_skew_sweep.csv | tee gpurun_out/r02_skew_sweep.txt
for r in 0 512 2048 0 512; do EVO_B200_GEMM_SKEW=$r python bench.py --steps 12 --warmup 3 --no-sub --no-gen --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/r02_bench_8k_skew$r.json; python -c "import json; d=json.load(open('gpurun_out/r02_bench_8k_skew$r.json')); print('skew=$r', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" | tee -a gpurun_out/r02_skew_sweep.txt; done
