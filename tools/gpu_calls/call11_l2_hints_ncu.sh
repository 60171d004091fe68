# L2 eviction hints on the GEMM's TMA loads: parity, A/B inside the bench, DRAM traffic under ncu; plus the launch list of one bench step.
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm" 2>&1 | tail -3
for f in 0 1 0 1; do EVO_B200_GEMM_L2_HINTS=$f python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r02_bench_hints$f.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r02_bench_hints$f.json')); print('l2_hints=$f', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved'],1))"; done
K='regex:gemm_tcgen05|hyena_scan|attn_pp|rmsnorm|embed_kernel|score_finish|logprobs|rotary|tokenize'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 231 -c 231 --csv --log-file gpurun_out/r02_launches_bench8k.csv python bench.py --steps 1 --warmup 1 --no-sub --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none -k 'regex:gemm_tcgen05|hyena_scan|rmsnorm' -s 238 -c 7 -o gpurun_out/r02_ncu_block_hints python bench.py --steps 1 --warmup 1 --no-sub --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
EVO_B200_GEMM_L2_HINTS=0 ncu --set full --clock-control none -k 'regex:gemm_tcgen05|hyena_scan|rmsnorm' -s 238 -c 7 -o gpurun_out/r02_ncu_block_nohints python bench.py --steps 1 --warmup 1 --no-sub --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
