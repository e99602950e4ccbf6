#!/bin/bash
set -u
O=gpurun_out/r2c7b
mkdir -p $O
echo "== 2-GPU tests" | tee -a $O/summary.txt
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_cli_host.py -q -m gpu -rxXs -k "two_gpu or two_gpus" 2>&1 | grep -E "passed|failed|FAILED|PASSED|Error" | tail -12 | tee -a $O/summary.txt
echo "== bench 2 GPUs: rowshard" | tee -a $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_rowshard.json 2> $O/bench_rowshard.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_rowshard.json | python tools/pj.py rowshard | tee -a $O/summary.txt
tail -1 $O/bench_rowshard.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('parity_check', d.get('parity_check'))" | tee -a $O/summary.txt
