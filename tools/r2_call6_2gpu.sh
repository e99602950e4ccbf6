#!/bin/bash
# Round 2, GPU call 6 (2 GPUs): sampler with one system fence per CTA; list exchange vs residual-delta all-reduce at sync rate 1.
set -u
O=gpurun_out/r2c6
mkdir -p $O
run() { # tag, env, extra args
  echo "== bench 2 GPUs: $1" | tee -a $O/summary.txt
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 $3 > $O/bench_$1.json 2> $O/bench_$1.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python tools/pj.py $1 | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('parity_check', d.get('parity_check'))" | tee -a $O/summary.txt
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" $O/bench_$1.err | tail -3 | tee -a $O/summary.txt
}
run lists "GMRM_EXCHANGE=lists" ""
run delta "GMRM_EXCHANGE=delta" ""
run delta_sync4 "X=1" "--sync-rate 4"
echo "== 2-GPU chain tests" | tee -a $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "chain" 2>&1 | tail -3 | tee -a $O/summary.txt
