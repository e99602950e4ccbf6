#!/bin/bash
# Round 2, GPU call 22 (1 GPU): A/B on one box of the table build fused behind the first batch loads (UKB size).
set -u
O=gpurun_out/r2c22
mkdir -p $O
run() { # tag, lib
  echo "== bench $1" | tee -a $O/summary.txt
  env GMRM_B200_LIB=$2 timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-setup-probes > $O/bench_$1.json 2> $O/bench_$1.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python tools/pj.py $1 | cut -c1-420 | tee -a $O/summary.txt
}
P=$PWD/gmrm_b200
run bf0 $P/variants/lib_bf0.so
run bf1 $P/variants/lib_bf1.so
run bf0b $P/variants/lib_bf0.so
run bf1b $P/variants/lib_bf1.so
