#!/bin/bash
# One 2-GPU box pass: the multi-rank GPU tests (skipped on a 1-GPU box) and the default bench line at N=2 (parity gate in global scope).
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu -k "multirank or multi_gpu or two_gpu or 2gpu" > $O/r2_gputest_n2.log 2>&1; echo "pytest n2 rc=$? $(tail -1 $O/r2_gputest_n2.log)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29551 bench.py --gpus 2 > $O/r2_bench_n2.json 2> $O/r2_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n2.json"))
print("N=2 value %.0f ms %.4f gate %s" % (d["value"], d["ms_per_step"], json.dumps(d["parity_gate"])[:600]))
print("e2e", d["e2e"]["value"], "c3", d["configs"]["c3"]["eager"], "c4", d["configs"]["c4"]["eager"])
PY
tail -3 $O/r2_bench_n2.err
