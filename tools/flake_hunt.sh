#!/bin/bash
# Repeats the streaming-head parity tests to look for run-to-run differences; prints the first failure in full.
O=gpurun_out; mkdir -p $O
fails=0
for i in $(seq 1 ${1:-12}); do
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or head_against_oracle or bf16_head or non_cubic" > $O/flake_$i.log 2>&1
  rc=$?
  echo "run $i rc=$rc $(tail -1 $O/flake_$i.log)"
  if [ $rc -ne 0 ]; then fails=$((fails+1)); [ $fails -eq 1 ] && grep -v Warn $O/flake_$i.log | grep -E "^E |Error|assert|FAILED|^tests" | head -30; fi
done
echo "failures: $fails"
