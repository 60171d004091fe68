# 8-GPU: (1) correctness of both sequence-parallel transports incl. PeerUlysses, (2) same-box A/B of the attention re-shard, (3) the default bench line
export NCCL_DEBUG_FILE=/dev/stderr
EVO_B200_PEER_ULYSSES=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/harness/seqpar_check.py 2>gpurun_out/seqpar8.err | grep seqpar_check > gpurun_out/r02_seqpar_check_8gpu_call13.json
cut -c1-2500 gpurun_out/r02_seqpar_check_8gpu_call13.json; tail -2 gpurun_out/seqpar8.err | cut -c1-300
for f in 0 1 0 1; do
EVO_B200_PEER_ULYSSES=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2950$f bench.py --gpus 8 --workload 131k --steps 5 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r02_sp8_ab_$f.json
python -c "import json; d=json.load(open('gpurun_out/r02_sp8_ab_$f.json')); print('peer_ulysses=$f', round(d['value']), round(d['ms_per_step'],1), [round(x,1) for x in d['per_rank_ms_per_step']])"
done
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29507 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/bench8.err | grep '^{' > gpurun_out/r02_bench_8k_8gpu_call13.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_8k_8gpu_call13.json'))
s = d['sp131k']
print('8k replicas', round(d['value']), [round(x, 1) for x in d['per_rank_ms_per_step']])
print('sp131k', round(s['value']), round(s['ms_per_step'], 1), s.get('attention_reshard'))
print('comm', {k: round(v, 2) for k, v in s['comm_ms'].items()})
print('kern', {k: round(v, 1) for k, v in s['kernel_ms'].items()}, 'unattr', round(s['unattributed_ms'], 1))
print('check', [(round(r['max_abs'], 4), round(r['mean_abs'], 5), round(r['argmax_agree'], 4)) for r in s['sp_check']['per_rank']])
PY
tail -2 gpurun_out/bench8.err | cut -c1-300
