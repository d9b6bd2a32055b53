#!/bin/bash
# compute-sanitizer passes over the small-size GPU parity tests: memcheck, racecheck, synccheck,
# initcheck.  Run on the GPU box:  bash tools/sanitize.sh [outdir]   -> <outdir>/sanitizer_<tool>.log + a summary line each.
# Full-size tests are left out (a sanitized 4.5 GB stream takes minutes and checks the same code as the 32^3 cases).
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SEL='not full_size and not largest_sweep and not incumbent and not model'
FILES="tests/test_gpu_parity.py tests/test_gpu_convhead.py tests/test_gpu_skeleton.py tests/test_gpu_eval.py"
for tool in memcheck racecheck synccheck initcheck; do
  log="$OUT/sanitizer_${tool}.log"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
      python -m pytest $FILES -m gpu -q -x -k "$SEL" -p no:cacheprovider > "$log" 2>&1
  rc=$?
  echo "[sanitize] $tool rc=$rc :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$log" | tail -1) :: $(grep -E 'passed|failed' "$log" | tail -1)"
done
