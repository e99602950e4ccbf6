"""Writes gpurun_in/pred/: a small dataset, a .bim pair and .bet histories for a hardware run of the CLI's --predict
mode (tools/predict_hw/run.sh on the GPU box, tools/predict_hw/check.py here against the oracle afterwards)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gmrm_b200 import synth            # noqa: E402
import test_cli_host as T              # noqa: E402

tmp = os.path.join(ROOT, "gpurun_in", "pred")
os.makedirs(tmp, exist_ok=True)
d = synth.write_dataset(tmp, N=1003, M=300, n_traits=2, n_groups=2, na_rate=0.01, missing_rate=0.005, seed=3)
d["tmp"] = tmp
bim, ref, hists, keep = T.write_predict_inputs(d, os.path.join(tmp, "bet"), 300)
np.save(os.path.join(tmp, "hists.npy"), np.stack(hists)); np.save(os.path.join(tmp, "keep.npy"), keep)
print("written", tmp)
