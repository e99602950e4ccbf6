#!/bin/bash
# Round 2, GPU call 14 (8 GPUs): the fused increment exchange at 8 GPUs -- UKB shape (with the parity check), C5 at sync rate 1, C4.
set -u
O=gpurun_out/r2c14
mkdir -p $O
nvidia-smi -L | wc -l | tee -a $O/summary.txt
run() { # tag, ngpu, env, extra args
  echo "== $1" | tee -a $O/summary.txt
  env $3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --steps 4 --warmup 2 $4 > $O/bench_$1.json 2> $O/bench_$1.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python tools/pj.py $1 | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('parity_check', d.get('parity_check'))" | tee -a $O/summary.txt
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" $O/bench_$1.err | tail -2 | tee -a $O/summary.txt
}
run ukb8_xdelta 8 "X=1" ""
run c5_sync1 8 "X=1" "--workload c5 --sync-rate 1 --no-parity-check"
run c4_8 8 "X=1" "--workload c4 --no-parity-check"
