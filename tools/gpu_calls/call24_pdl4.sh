# decode: level 4 = level 3 + evo_hyena_step launched dependent, its parameter/state loads hoisted above the wait; parity first
python -m pytest tests/test_gpu_parity.py tests/test_gpu_generation.py -m gpu -q -x -k "decode or device_loop" 2>&1 | tail -3
for f in 3 4 3 4; do EVO_B200_DECODE_PDL=$f python bench.py --workload gen --gen-tokens 64 --steps 128 > gpurun_out/r02_bench_gen_pdl$f.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r02_bench_gen_pdl$f.json')); print('pdl=$f', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"; done
