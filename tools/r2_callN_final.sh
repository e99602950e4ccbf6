#!/bin/bash
# Round 2, final multi-GPU line of the final tree: bash tools/r2_callN_final.sh N  (bench.py --gpus N with its parity check)
set -u
N=$1
O=gpurun_out/r2final_$N
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench.json | python tools/pj.py gpus$N | tee -a $O/summary.txt
tail -1 $O/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('parity_check', d.get('parity_check')); print('e2e', d['e2e']['value'])" | tee -a $O/summary.txt
