#!/bin/bash
# Round 2, GPU call 4: global partial sums on hardware, V sweep at UKB shape, staleness vs V at N = 458,000, cost per published marker.
set -u
O=gpurun_out/r2c4
mkdir -p $O
echo "== GPU parity (production streams incl. V*T > 2048, staged outputs, CLI)" | tee -a $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cli_host.py -q -m gpu -x -k "production or staged or cli or dot" 2>&1 | tail -5 | tee -a $O/summary.txt
for V in 1024 2048 4096 8192; do
  echo "== bench V=$V" | tee -a $O/summary.txt
  timeout 600 python bench.py --vranks-per-gpu $V --no-cpu-baseline --no-setup-probes --steps 6 --warmup 3 > $O/bench_V$V.json 2> $O/bench_V$V.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_V$V.json | python tools/pj.py V$V | tee -a $O/summary.txt
  tail -2 $O/bench_V$V.err | tee -a $O/summary.txt
done
echo "== ukbn (N=458000, M=50000): staleness vs V" | tee -a $O/summary.txt
timeout 1500 python tools/chain_probe.py --workload ukbn --vranks 128,1024,2048,4096,8192,16384 --iterations 1500 --burn 300 --out $O/ukbn > $O/ukbn.log 2>&1
tail -7 $O/ukbn.log | cut -c1-800 | tee -a $O/summary.txt
