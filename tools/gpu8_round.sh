#!/bin/bash
# One 8-GPU box pass: topology, the default bench line at N=8 (parity gate, configs[2]/[3], conv-fused, e2e), the configs[4] sweep at 8 GPUs.
O=gpurun_out; mkdir -p $O
{ nvidia-smi topo -m; echo; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; echo; free -g | head -2; } > $O/r2_topology_n8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29541 bench.py --gpus 8 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err; echo "bench n8 rc=$?"
$TR --master-port 29542 tools/sweep.py --quick --scope global --out $O/r2_sweep_f32_n8.csv > $O/r2_sweep_n8.log 2>&1; echo "sweep f32 n8 rc=$?"
$TR --master-port 29543 tools/sweep.py --quick --scope global --graph --out $O/r2_sweep_f32_n8_graph.csv > $O/r2_sweep_n8g.log 2>&1; echo "sweep f32 graph n8 rc=$?"
tail -c 1500 $O/r2_bench_n8.json; tail -3 $O/r2_bench_n8.err; tail -5 $O/r2_sweep_f32_n8.csv
