#!/bin/bash
set -u
O=gpurun_out/r2c8
mkdir -p $O
echo "== 2-GPU tests + 1-GPU parity subset" | tee -a $O/summary.txt
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_cli_host.py tests/test_gpu_parity.py -q -m gpu -rxXs -k "two_gpu or two_gpus or production or replay_golden or update" 2>&1 | grep -E "passed|failed|FAILED|Error" | tail -8 | tee -a $O/summary.txt
run() { # tag, env, extra args
  echo "== bench 2 GPUs: $1" | tee -a $O/summary.txt
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 $3 > $O/bench_$1.json 2> $O/bench_$1.err; echo "rc=$?" | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python tools/pj.py $1 | tee -a $O/summary.txt
  tail -1 $O/bench_$1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('parity_check', d.get('parity_check'))" | tee -a $O/summary.txt
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" $O/bench_$1.err | tail -2 | tee -a $O/summary.txt
}
run rowshard "GMRM_ROWSHARD=1" ""
run allrows "GMRM_ROWSHARD=0" "--no-parity-check"
echo "== bench 1 GPU" | tee -a $O/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-setup-probes > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_1gpu.json | python tools/pj.py 1gpu | tee -a $O/summary.txt
