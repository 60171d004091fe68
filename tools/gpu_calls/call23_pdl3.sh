# decode: programmatic dependent launch for the row-norm kernels too (level 3) vs GEMMs only (level 2); parity first
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode" 2>&1 | tail -3
for f in 2 3 2 3; do EVO_B200_DECODE_PDL=$f python bench.py --workload gen --gen-tokens 64 --steps 128 > gpurun_out/r02_bench_gen_pdl$f.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r02_bench_gen_pdl$f.json')); print('pdl=$f', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"; done
