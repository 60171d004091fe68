# final checks of the round: the whole GPU suite, the default bench line (with every sub-record), the reference arm, and one last A/B
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_8k_call15.json 2> gpurun_out/r02_bench_8k_call15.err; tail -c 400 gpurun_out/r02_bench_8k_call15.json; tail -2 gpurun_out/r02_bench_8k_call15.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm_call15.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_reference_arm_call15.json
for f in 0 2 0 2; do EVO_B200_GEMM_L2_HINTS=$f python bench.py --steps 4 --warmup 2 --no-sub --no-cpu-baseline > gpurun_out/r02_bench_hints_m$f.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r02_bench_hints_m$f.json')); print('l2_hints=$f', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved'],1))"; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
