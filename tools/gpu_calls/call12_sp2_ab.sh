# same-box A/B of the attention re-shard on 2 GPUs: NCCL Ulysses (0) vs peer stores fused into the epilogues (1)
for f in 0 1 0 1; do
EVO_B200_PEER_ULYSSES=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2950$f bench.py --gpus 2 --workload 131k --steps 4 --warmup 2 2>/dev/null | grep '^{' > gpurun_out/r02_sp2_ab_$f.json
python -c "import json; d=json.load(open('gpurun_out/r02_sp2_ab_$f.json')); print('peer_ulysses=$f', round(d['value']), round(d['ms_per_step'],1), d['per_rank_ms_per_step'])"
done
