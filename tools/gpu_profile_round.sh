#!/bin/bash
# One GPU-box pass that produces the round's profiling evidence under gpurun_out/ (copied into profiles/ afterwards):
# ncu launch list of the default bench command, `ncu --set full` of K1/K3 for c2/c3/c4, small-batch (8-GPU shard size) lines.
set -u
O=gpurun_out; mkdir -p $O
FAST="--no-cpu --no-e2e --no-configs --no-sustained"
python bench.py --steps 3 --warmup 3 $FAST > $O/p_plain.json 2> $O/p_plain.err || { echo "plain bench failed"; tail -5 $O/p_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_raw.csv \
    python bench.py --steps 3 --warmup 3 $FAST > $O/p_ncu.log 2>&1
echo "launch list rc=$?"
for cfg in c2 c3 c4; do
  B=256; [ $cfg = c3 ] && B=128; [ $cfg = c4 ] && B=64
  [ $cfg = c2 ] && B=256
  ncu --set full --clock-control none --import-source on -k regex:'integral_(fwd|bwd)_kernel' -s 14 -c 2 -f -o $O/r2_full_${cfg}_b${B} \
      python bench.py --config $cfg --batch $B --steps 2 --warmup 3 $FAST > $O/p_full_$cfg.log 2>&1
  echo "full $cfg rc=$?"
done
# shard sizes of BASELINE configs[2]/[3] on 8 GPUs, on one GPU (local scope: no exchange)
for spec in "c3 128" "c3 256" "c3 512" "c4 64" "c4 128" "c2 256"; do
  set -- $spec
  python bench.py --config $1 --batch $2 --steps 50 --warmup 5 $FAST > $O/small_$1_b$2.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("$O/small_$1_b$2.json"))
print("$1 B=$2 eager %.4f ms frac8=%.3f | graph %.4f ms | K3 %.4f ms K1 %.4f ms loss %s" % (d["ms_per_step"], d["roofline"]["whole_step"]["frac_of_8TBs"], d["cuda_graph_replay"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["fwd_kernel"]["ms_per_launch"], d["roofline"]["loss_kernels"]))
PY
done
