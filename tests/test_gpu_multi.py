"""GPU, >= 2 devices: marker-sharded chain over NCCL against the oracle (tests/mgpu_check.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sync_rate,exchange", [(1, "xdelta"), (1, "lists"), (1, "delta"), (3, "delta")])
def test_two_gpu_chain_matches_oracle(sync_rate, exchange):
    """sync rate 1 with the fused increment exchange inside the step kernel (the default: own list applied as increments, reduced
    and broadcast over NVLink peer memory; one chain, bit-identical replicas), with the published lists over peer memory (same
    guarantee) and with NCCL-all-reduced residual deltas (GMRM_EXCHANGE=delta: one replica per GPU, merged inside the next step
    kernel); sync rate 3: deltas."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29611 + sync_rate + {"delta": 7, "xdelta": 11}.get(exchange, 0)), os.path.join(ROOT, "tests", "mgpu_check.py"), str(sync_rate)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, GMRM_EXCHANGE=exchange))
    assert p.returncode == 0 and "MGPU_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def test_two_gpu_predict_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "mgpu_check.py"), "1", "predict"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "MGPU_PREDICT_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
