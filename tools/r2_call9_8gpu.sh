#!/bin/bash
# Round 2, GPU call 9 (8 GPUs): exchange variants at the UKB shape, 4 GPUs, C5 sync-rate sweep, C4.
set -u
O=gpurun_out/r2c9
mkdir -p $O
nvidia-smi -L | wc -l | tee -a $O/summary.txt
run() { # tag, ngpu, env, extra args
  echo "== $1" | tee -a $O/summary.txt
  env $3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --steps 4 --warmup 2 $4 > $O/bench_$1.json 2> $O/bench_$1.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python tools/pj.py $1 | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('parity_check', d.get('parity_check'))" | tee -a $O/summary.txt
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" $O/bench_$1.err | tail -2 | tee -a $O/summary.txt
}
run ukb8_rowshard 8 "GMRM_ROWSHARD=1" ""
run ukb8_allrows 8 "GMRM_ROWSHARD=0" "--no-parity-check"
run ukb8_delta 8 "GMRM_EXCHANGE=delta" "--no-parity-check"
run ukb4_rowshard 4 "GMRM_ROWSHARD=1" "--no-parity-check"
run c5_sync1 8 "X=1" "--workload c5 --sync-rate 1 --no-parity-check"
run c5_sync2 8 "X=1" "--workload c5 --sync-rate 2 --no-parity-check"
run c5_sync4 8 "X=1" "--workload c5 --sync-rate 4 --no-parity-check"
run c5_sync8 8 "X=1" "--workload c5 --sync-rate 8 --no-parity-check"
run c4_8 8 "X=1" "--workload c4 --no-parity-check"
