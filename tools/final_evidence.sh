#!/bin/bash
# Round-end evidence on one GPU: what the driver runs (tools/validate_head.sh), the configs[4] sweeps, the profiling pass
# (tools/gpu_profile_round.sh) and one `ncu --set full` capture of the streaming forward at 64 KB bf16 units.
set -u
O=gpurun_out; mkdir -p $O
bash tools/validate_head.sh
python tools/sweep.py --dtype f32 --out $O/r2_sweep_f32.csv > $O/sw_f32.log 2>&1; echo "sweep f32 rc=$?"
python tools/sweep.py --dtype bf16 --out $O/r2_sweep_bf16.csv > $O/sw_bf16.log 2>&1; echo "sweep bf16 rc=$?"
python tools/sweep.py --dtype f32 --graph --quick --out $O/r2_sweep_f32_graph.csv > $O/sw_f32g.log 2>&1; echo "sweep f32 graph rc=$?"
bash tools/gpu_profile_round.sh
python tools/stream_probe.py --res 32 --batch 4096 --dtype bf16 --iters 3 > $O/k1_bf16_32_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'integral_fwd_kernel' -s 2 -c 1 -f -o $O/r2_k1_bf16_32_nf3 \
    python tools/stream_probe.py --res 32 --batch 4096 --dtype bf16 --iters 1 > $O/k1_bf16_32_ncu.log 2>&1; echo "ncu k1 bf16 32 rc=$?"
python tools/k2_probe.py > $O/r2_k2_probe.txt 2>&1; tail -8 $O/r2_k2_probe.txt
