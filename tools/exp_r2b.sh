#!/bin/bash
# Experiments of the second half of round 2: finaliser count of the streaming forward at short units, L2 hand-over of the
# end of the volume from the forward to the backward.  Prints one line per setting; raw lines under gpurun_out/exp/.
set -u
O=gpurun_out/exp; mkdir -p $O
FAST="--no-cpu --no-e2e --no-configs --no-sustained"
echo "== parity (small sizes) =="
timeout 600 python -m pytest tests -x -q -m gpu -k "parity or guards or golden" 2>&1 | tail -3
echo "== K1 alone, finaliser count =="
for nf in 2 3; do
  for spec in "32 4096 bf16" "32 4096 f32" "32 1024 bf16" "64 1024 bf16" "64 256 f32"; do
    set -- $spec
    echo -n "nf=$nf "; XSUP_K1_FINALISERS=$nf python tools/k1_probe.py --res $1 --batch $2 --dtype $3 --iters 10 | tail -1
  done
done
echo "== whole step, L2 keep =="
for keep in 0 32 64 96; do
  for spec in "c4 64" "c3 128" "c2 256"; do
    set -- $spec
    XSUP_L2_KEEP_MB=$keep python bench.py --config $1 --batch $2 --steps 50 --warmup 5 $FAST > $O/keep${keep}_$1_b$2.json 2>/dev/null
    python - <<PY
import json
d=json.load(open("$O/keep${keep}_$1_b$2.json"))
print("keep=$keep $1 B=$2 eager %.4f ms frac8=%.3f | graph %.4f ms | K3 %.4f ms K1 %.4f ms gate %s" % (d["ms_per_step"], d["roofline"]["whole_step"]["frac_of_8TBs"], d["cuda_graph_replay"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["fwd_kernel"]["ms_per_launch"], d["parity_gate"]["ok"]))
PY
  done
done
