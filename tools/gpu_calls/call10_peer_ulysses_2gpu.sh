# 2-GPU validation of the NCCL-free attention re-shard (PeerUlysses): correctness first, then the bench line.
export EVO_B200_PEER_ULYSSES=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/harness/seqpar_check.py 2>gpurun_out/seqpar2.err | grep seqpar_check > gpurun_out/r02_seqpar_check_2gpu_peer_ulysses_call10.json
cut -c1-1200 gpurun_out/r02_seqpar_check_2gpu_peer_ulysses_call10.json; tail -3 gpurun_out/seqpar2.err
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_8k_2gpu_peer_ulysses_call10.json 2> gpurun_out/bench2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_8k_2gpu_peer_ulysses_call10.json'))
s = d['sp131k']
print('sp131k', round(s['value']), round(s['ms_per_step'], 1), s.get('attention_reshard'), 'comm', {k: round(v, 2) for k, v in s['comm_ms'].items()}, 'kern', {k: round(v, 1) for k, v in s['kernel_ms'].items()}, 'unattr', round(s['unattributed_ms'], 1))
print('check', s['sp_check']['per_rank'])
PY
tail -3 gpurun_out/bench2.err
