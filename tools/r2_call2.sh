#!/bin/bash
# Round 2, GPU call 2: segmented sampler on hardware (full GPU suite + bench), publish rate over a longer UKB chain,
# posterior validity on C2 at several numbers of virtual ranks.
set -u
O=gpurun_out/r2c2
mkdir -p $O
echo "== GPU suite" | tee -a $O/summary.txt
timeout 900 python -m pytest tests -q -m gpu -rxXs -x 2>&1 | tail -12 | tee -a $O/summary.txt
echo "== bench" | tee -a $O/summary.txt
timeout 900 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench.json | python tools/pj.py segsampler | tee -a $O/summary.txt
echo "== long UKB chain" | tee -a $O/summary.txt
timeout 600 python tools/chain_probe.py --workload ukb --vranks 2048 --iterations 80 --burn 40 --trace-every 5 --out $O/ukb > $O/ukb.log 2>&1
tail -3 $O/ukb.log | cut -c1-600 | tee -a $O/summary.txt
echo "== C2 posterior probes" | tee -a $O/summary.txt
timeout 1500 python tools/chain_probe.py --workload c2 --vranks 64,512,2048,8192,16384 --iterations 2000 --burn 500 --out $O/c2 > $O/c2.log 2>&1
tail -6 $O/c2.log | cut -c1-900 | tee -a $O/summary.txt
