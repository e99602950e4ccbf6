#!/bin/bash
# Round 2, GPU call 20 (1 GPU): A/B on one box of the L2 evict-first policy on the genotype stream (UKB size).
set -u
O=gpurun_out/r2c20
mkdir -p $O
run() { # tag, lib
  echo "== bench $1" | tee -a $O/summary.txt
  env GMRM_B200_LIB=$2 timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-setup-probes > $O/bench_$1.json 2> $O/bench_$1.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python tools/pj.py $1 | cut -c1-420 | tee -a $O/summary.txt
}
P=$PWD/gmrm_b200
run noef $P/variants/lib_noef.so
run ef $P/variants/lib_ef.so
run noef2 $P/variants/lib_noef.so
run ef2 $P/variants/lib_ef.so
