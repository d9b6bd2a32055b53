#!/bin/bash
# What the driver runs at round end, in the same order, on one GPU box: GPU tests, smoke, the reference arm, the default bench.
set -u
O=gpurun_out; mkdir -p $O
S=$(date +%s)
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2_gputest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/r2_gputest.log) [$(( $(date +%s) - S )) s]"
S=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/r2_smoke.log) [$(( $(date +%s) - S )) s]"
S=$(date +%s)
timeout 600 python bench.py --impl reference > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; echo "reference arm rc=$? [$(( $(date +%s) - S )) s]"; cut -c1-300 $O/r2_bench_reference.json
S=$(date +%s)
timeout 900 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err; echo "bench rc=$? [$(( $(date +%s) - S )) s]"; cut -c1-400 $O/r2_bench.json
