# 4 GPUs: the default bench line (replicas + sp131k with the default PeerUlysses re-shard), the one N the round had not touched yet
export NCCL_DEBUG_FILE=/dev/stderr
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29507 bench.py --gpus 4 --steps 5 --warmup 3 2>gpurun_out/bench4.err | grep '^{' > gpurun_out/r02_bench_8k_4gpu_call17.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_8k_4gpu_call17.json'))
s = d['sp131k']
print('8k replicas', round(d['value']), [round(x, 1) for x in d['per_rank_ms_per_step']])
print('sp131k', round(s['value']), round(s['ms_per_step'], 1), s.get('attention_reshard'))
print('comm', {k: round(v, 2) for k, v in s['comm_ms'].items()})
print('kern', {k: round(v, 1) for k, v in s['kernel_ms'].items()}, 'unattr', round(s['unattributed_ms'], 1))
print('check', [(round(r['max_abs'], 4), round(r['mean_abs'], 5), round(r['argmax_agree'], 4)) for r in s['sp_check']['per_rank']])
PY
tail -2 gpurun_out/bench4.err | cut -c1-300
