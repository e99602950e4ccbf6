#!/bin/bash
# Round 2, GPU call 13 (1 GPU): full GPU suite on the new stream loop, then A/B of library variants on one box
# (same marker slice, N = 458,000): base = stream loop of commit c51b70e, new = product, w12 / w14 = 12 / 14 warps.
set -u
O=gpurun_out/r2c13
mkdir -p $O
echo "== GPU suite" | tee -a $O/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -x -rxXs 2>&1 | tail -8 | tee -a $O/summary.txt
run() { # tag, lib
  echo "== bench $1" | tee -a $O/summary.txt
  env GMRM_B200_LIB=$2 timeout 600 python bench.py --steps 4 --warmup 3 --markers 262144 --no-cpu-baseline --no-setup-probes > $O/bench_$1.json 2> $O/bench_$1.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python tools/pj.py $1 | tee -a $O/summary.txt
}
P=$PWD/gmrm_b200
run base $P/variants/lib_base.so
run new $P/libgmrm_b200.so
run w12 $P/variants/lib_w12.so
run w14 $P/variants/lib_w14.so
run base2 $P/variants/lib_base.so
run new2 $P/libgmrm_b200.so
