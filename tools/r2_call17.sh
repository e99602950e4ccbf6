#!/bin/bash
# Round 2, GPU call 17 (1 GPU): GPU suite + default bench on the final tree (e2e warm-up, two-level beta^2 sum).
set -u
O=gpurun_out/r2c17
mkdir -p $O
echo "== GPU suite" | tee -a $O/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -x -rxXs 2>&1 | tail -6 | tee -a $O/summary.txt
echo "== smoke" | tee -a $O/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $O/summary.txt
echo "== bench (default flags)" | tee -a $O/summary.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench.json | python tools/pj.py final | tee -a $O/summary.txt
echo "== reference arm" | tee -a $O/summary.txt
timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_ref.json | cut -c1-600 | tee -a $O/summary.txt
