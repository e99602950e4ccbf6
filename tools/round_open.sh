#!/bin/bash
# First GPU call of a round, everything that DESIGN.md section 8 leaves open in ONE box acquisition (run under gpurun from
# the repo root: `gpurun --timeout 1500 -- 'bash tools/round_open.sh'`).  Every step is bounded by its own timeout and
# writes into gpurun_out/; nothing printed under ncu is a bench value.
set -u
O=gpurun_out/round_open
mkdir -p $O
step() { echo "== $1" | tee -a $O/summary.txt; }

step "full GPU suite (xfail-marked tests report XPASS/XFAIL)"
timeout 600 python -m pytest tests -q -m gpu -rxX 2>&1 | tail -25 | tee -a $O/summary.txt

step "pipe mix microbenchmark (DESIGN section 5, lever 1)"
if [ -x tools/hybrid_micro ]; then timeout 60 tools/hybrid_micro 2>&1 | tee $O/hybrid_micro.txt | tail -30 >> $O/summary.txt; else echo "tools/hybrid_micro not built (nvcc line in tools/README.md)" | tee -a $O/summary.txt; fi

step "predict throughput (UKB-shaped slice, 1 and 64 marker blocks)"
timeout 300 python tools/predict_bench.py --blocks 1 2>&1 | tail -1 | tee -a $O/summary.txt
timeout 300 python tools/predict_bench.py --blocks 64 2>&1 | tail -1 | tee -a $O/summary.txt

step "bench, 1 GPU, default steps"
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err; tail -1 $O/bench.json | python tools/pj.py 2>/dev/null | tee -a $O/summary.txt

step "launch list of the same command (shares only)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
echo "rc=$?" | tee -a $O/summary.txt
