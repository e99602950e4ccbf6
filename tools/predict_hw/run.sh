#!/bin/bash
# runs the CLI's --predict mode for 1 and 3 marker blocks on the prepared dataset; outputs under gpurun_out/
D=gpurun_in/pred
for R in 1 3; do
  O=gpurun_out/pred$R
  mkdir -p $O && cp $D/bet/*.bet $O/
  timeout 8 gmrm_b200/gmrm_b200_cli --bed-file $D/syn.bed --dim-file $D/syn.dim --phen-files $D/syn_t0.phen,$D/syn_t1.phen \
      --out-dir $O --predict --bim-file $D/p.bim --ref-bim-file $D/pref.bim --vranks $R > $O/log.txt 2>&1
  echo "R=$R rc=$?"
done
