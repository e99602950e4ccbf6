#!/bin/bash
# Round 2, GPU call 4b: chunked large-V steps on hardware: parity, V sweep at UKB shape, UKB-shape chains at V = 2048 / 8192 / 16384.
set -u
O=gpurun_out/r2c4b
mkdir -p $O
echo "== GPU parity (production streams incl. V*T > 2048)" | tee -a $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "production or dot" 2>&1 | tail -5 | tee -a $O/summary.txt
for V in 2048 4096 8192 16384; do
  echo "== bench V=$V" | tee -a $O/summary.txt
  timeout 600 python bench.py --vranks-per-gpu $V --no-cpu-baseline --no-setup-probes --steps 6 --warmup 3 > $O/bench_V$V.json 2> $O/bench_V$V.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_V$V.json | python tools/pj.py V$V | tee -a $O/summary.txt
  tail -2 $O/bench_V$V.err | tee -a $O/summary.txt
done
echo "== UKB shape (N=458000, M=1000000): chains at V = 2048, 8192, 16384" | tee -a $O/summary.txt
timeout 1500 python tools/chain_probe.py --workload ukb --vranks 2048,8192,16384 --iterations 400 --burn 200 --out $O/ukb > $O/ukb.log 2>&1
tail -4 $O/ukb.log | cut -c1-900 | tee -a $O/summary.txt
