#!/bin/bash
# Round 2, GPU call 15 (1 GPU): the hybrid stream (direct rows) -- GPU suite, A/B on one box, UKB-size bench with the in-kernel profile.
set -u
O=gpurun_out/r2c15
mkdir -p $O
echo "== GPU suite (hybrid plan on for one trait)" | tee -a $O/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -x -rxXs 2>&1 | tail -8 | tee -a $O/summary.txt
run() { # tag, env, args
  echo "== bench $1" | tee -a $O/summary.txt
  env $2 timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-setup-probes $3 > $O/bench_$1.json 2> $O/bench_$1.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python tools/pj.py $1 | tee -a $O/summary.txt
  grep "step prof" $O/bench_$1.err | tail -1 | cut -c1-900 | tee -a $O/summary.txt
}
P=$PWD/gmrm_b200
run p16 "GMRM_B200_LIB=$P/variants/lib_p16.so" "--markers 262144"
run hyb12 "X=1" "--markers 262144"
run pure12 "GMRM_HYBRID=0" "--markers 262144"
run ukb_hyb "GMRM_STEP_PROF=4" ""
