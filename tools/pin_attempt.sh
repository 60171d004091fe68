#!/bin/bash
# Round 2, VERDICT item 3: try to obtain stripedhyena==0.2.2 (the package holding the reference's arithmetic,
# /root/reference/requirements.txt:1) on the GPU box so the oracle can be pinned.  Logs everything to gpurun_out/.
out=gpurun_out/pin_attempt.log
mkdir -p gpurun_out
{
  echo "== date"; date -u
  echo "== python -c import stripedhyena"; python -c 'import stripedhyena, sys; print(stripedhyena.__file__)' 2>&1 | tail -2
  echo "== pip download stripedhyena==0.2.2 --no-deps (index)"; timeout 60 python -m pip download stripedhyena==0.2.2 --no-deps -d /tmp/sh 2>&1 | tail -8
  echo "== pip install from /opt/wheelhouse"; timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse --no-deps --target /tmp/sh_t stripedhyena==0.2.2 2>&1 | tail -5
  echo "== wheelhouse listing (hyena|evo|flash)"; ls /opt/wheelhouse 2>/dev/null | grep -i -E 'hyena|evo|flash' ; echo "(end)"
  echo "== find hyena / evo checkpoints / HF cache"; find / -xdev \( -iname '*hyena*' -o -iname '*evo-1*' -o -iname 'models--togethercomputer*' \) -not -path '/proc/*' -not -path "$PWD/*" -not -path '/root/repo/*' 2>/dev/null | head -20; echo "(end)"
  echo "== HF cache dirs"; ls -d ~/.cache/huggingface /root/.cache/huggingface/hub/* 2>&1 | head
  echo "== network probe"; timeout 10 python - <<'PY'
import socket
try:
    socket.create_connection(("pypi.org", 443), timeout=5); print("pypi reachable")
except Exception as e:
    print("pypi unreachable:", repr(e))
PY
  echo "== host"; nproc; grep -m1 'model name' /proc/cpuinfo; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
} > $out 2>&1
cat $out
