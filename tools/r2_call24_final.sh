#!/bin/bash
# Round 2, final GPU call (1 GPU): GPU suite, smoke, default bench, launch list and ncu --set full captures of the final tree.
set -u
O=gpurun_out/r2c24
mkdir -p $O
echo "== GPU suite" | tee -a $O/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -x -rxXs 2>&1 | tail -6 | tee -a $O/summary.txt
echo "== smoke" | tee -a $O/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $O/summary.txt
echo "== bench (default flags)" | tee -a $O/summary.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench.json | python tools/pj.py final | tee -a $O/summary.txt
K='regex:step_kernel|sample_kernel|eps_|beta_sq|global_draw|steptab|group_consts|mu_draw|init_sigmae'
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-setup-probes"
echo "== launch list" | tee -a $O/summary.txt
timeout 600 $CMD > $O/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 2400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "rc=$?" | tee -a $O/summary.txt
echo "== ncu --set full: step_kernel, sample_kernel" | tee -a $O/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 700 -c 1 -o $O/prof_step $CMD > $O/ncu_step.log 2>&1
echo "rc=$?" | tee -a $O/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sample_kernel -s 700 -c 1 -o $O/prof_sample $CMD > $O/ncu_sample.log 2>&1
echo "rc=$?" | tee -a $O/summary.txt
