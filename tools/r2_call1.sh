#!/bin/bash
# Round 2, GPU call 1: full GPU suite on the current tree, pipe-mix microbenchmark, bench with the SURVEY 8d phenotype.
set -u
O=gpurun_out/r2c1
mkdir -p $O
step() { echo "== $1" | tee -a $O/summary.txt; }
step "box"
(nvidia-smi -L; nproc; free -g | head -2; df -h /tmp / | tail -3) 2>&1 | tee -a $O/summary.txt
step "full GPU suite"
timeout 900 python -m pytest tests -q -m gpu -rxXs -x 2>&1 | tail -30 | tee -a $O/summary.txt
step "hybrid_micro"
timeout 120 tools/hybrid_micro 2>&1 | tee $O/hybrid_micro.txt | tail -40 >> $O/summary.txt
step "bench 1 GPU (8d phenotype, default flags)"
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench.json | python tools/pj.py new8d | tee -a $O/summary.txt
tail -5 $O/bench.err >> $O/summary.txt
step "bench 1 GPU, 200 causal markers (round-1 workload), in-kernel cycle counters from iteration 5"
GMRM_STEP_PROF=5 timeout 600 python bench.py --causal-frac 0.0002 --no-setup-probes --no-cpu-baseline --steps 3 --warmup 3 > $O/bench_r1wl.json 2> $O/bench_r1wl.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_r1wl.json | python tools/pj.py r1workload | tee -a $O/summary.txt
grep "step prof" $O/bench_r1wl.err | tail -2 >> $O/summary.txt
step "bench 1 GPU 8d phenotype with in-kernel cycle counters, 6+3 iterations"
GMRM_STEP_PROF=8 timeout 600 python bench.py --no-setup-probes --no-cpu-baseline --steps 3 --warmup 6 > $O/bench_prof.json 2> $O/bench_prof.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_prof.json | python tools/pj.py 8d_prof | tee -a $O/summary.txt
grep "step prof" $O/bench_prof.err | tail -2 >> $O/summary.txt
