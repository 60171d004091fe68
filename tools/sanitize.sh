#!/bin/bash
# compute-sanitizer passes over the small-shape GPU tests (development tool, run under gpurun).
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards; synccheck: barrier misuse.
set -u
OUT=gpurun_out
mkdir -p $OUT
SEL='rmsnorm or embed or rope or kv_append or hyena_operator or hyena_step or continuation or (gemm_all and 300-512) or (attention_vs_oracle and 300) or decode_attention or stateful'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "$SEL" > $OUT/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/sanitizer_$tool.log | tr '\n' ' ')"
done
