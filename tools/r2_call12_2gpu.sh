#!/bin/bash
# Round 2, GPU call 12 / 18 (2 GPUs): the data-driven increment exchange (no flags, no fences) -- parity tests, bench.
set -u
O=gpurun_out/r2c18
mkdir -p $O
echo "== 2-GPU tests" | tee -a $O/summary.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_cli_host.py -q -m gpu -rxXs -k "two_gpu or two_gpus" 2>&1 | tail -15 | tee -a $O/summary.txt
run() { # tag, env, extra args
  echo "== bench 2 GPUs: $1" | tee -a $O/summary.txt
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 $3 > $O/bench_$1.json 2> $O/bench_$1.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python tools/pj.py $1 | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('parity_check', d.get('parity_check'))" | tee -a $O/summary.txt
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" $O/bench_$1.err | tail -3 | tee -a $O/summary.txt
}
run xdelta "X=1" ""
