#!/bin/bash
# Round 2, GPU call 25 (1 GPU): partial sums written by the last pass itself -- parity subset on the product build, A/B on one box.
set -u
O=gpurun_out/r2c25
mkdir -p $O
echo "== parity subset (product build)" | tee -a $O/summary.txt
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_predict.py -q -m gpu -x -k "dot or production or replay_golden or predict" 2>&1 | tail -3 | tee -a $O/summary.txt
run() { # tag, lib
  echo "== bench $1" | tee -a $O/summary.txt
  env GMRM_B200_LIB=$2 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-setup-probes > $O/bench_$1.json 2> $O/bench_$1.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python tools/pj.py $1 | cut -c1-420 | tee -a $O/summary.txt
}
P=$PWD/gmrm_b200
run pd0 $P/variants/lib_pd0.so
run pd1 $P/variants/lib_pd1.so
run pd0b $P/variants/lib_pd0.so
run pd1b $P/variants/lib_pd1.so
