#!/bin/bash
# Round 2, GPU call 3: pair-table update phase + PDL + async iteration on hardware; PDL A/B; C1 replay test; more V points.
set -u
O=gpurun_out/r2c3
mkdir -p $O
echo "== GPU suite" | tee -a $O/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -rxXs -x 2>&1 | tail -12 | tee -a $O/summary.txt
echo "== bench (PDL on)" | tee -a $O/summary.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench.json | python tools/pj.py pdl_on | tee -a $O/summary.txt
tail -3 $O/bench.err | tee -a $O/summary.txt
echo "== bench (PDL off)" | tee -a $O/summary.txt
GMRM_PDL=0 timeout 900 python bench.py --no-cpu-baseline --no-setup-probes > $O/bench_nopdl.json 2> $O/bench_nopdl.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_nopdl.json | python tools/pj.py pdl_off | tee -a $O/summary.txt
echo "== C2 posterior probes, more V" | tee -a $O/summary.txt
timeout 1500 python tools/chain_probe.py --workload c2 --vranks 64,256,800,1024 --iterations 2000 --burn 500 --seed 5 --out $O/c2 > $O/c2.log 2>&1
tail -5 $O/c2.log | cut -c1-700 | tee -a $O/summary.txt
