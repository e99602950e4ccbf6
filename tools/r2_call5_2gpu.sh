#!/bin/bash
# Round 2, GPU call 5 (2 GPUs): multi-GPU parity on the final tree (chain at sync rate 1 and 3, predict, CLI --gpus 2), bench at 2 GPUs.
set -u
O=gpurun_out/r2c5
mkdir -p $O
nvidia-smi -L | tee -a $O/summary.txt
echo "== 2-GPU tests" | tee -a $O/summary.txt
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_cli_host.py -q -m gpu -rxXs -k "two_gpu or two_gpus" 2>&1 | tail -12 | tee -a $O/summary.txt
echo "== bench 2 GPUs" | tee -a $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_2gpu.json | python tools/pj.py 2gpu | tee -a $O/summary.txt
tail -1 $O/bench_2gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('parity_check', d.get('parity_check'))" | tee -a $O/summary.txt
tail -3 $O/bench_2gpu.err | tee -a $O/summary.txt
echo "== bench 1 GPU (same box)" | tee -a $O/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-setup-probes > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_1gpu.json | python tools/pj.py 1gpu | tee -a $O/summary.txt
