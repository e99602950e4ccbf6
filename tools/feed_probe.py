"""Probe: step-kernel phase cycles (GMRM_STEP_PROF) when the V columns are distinct (HBM) vs all the same (L2)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmrm_b200 import api
N, M, V = 458000, 8192, 2048
e = api.Engine(N=N, Mt=M, vranks=1)
e.generate_bed(seed=1)
e.finalize_bed()
rng = np.random.default_rng(0)
y = rng.normal(size=N); y -= y.mean(); y /= y.std()
mask4 = np.full((N + 3) // 4, 0xF, dtype=np.uint8)
e.set_phenotype(0, y, mask4, N)
e.set_groups(np.zeros(M, dtype=np.int32), np.array([[0.0, 1e-4, 1e-3, 1e-2]]))
e.compute_marker_stats()
print("distinct", flush=True)
for _ in range(3):
    e.dot_products(rng.permutation(M)[:V].astype(np.int32))
print("same", flush=True)
for _ in range(3):
    e.dot_products(np.zeros(V, dtype=np.int32))
e.close()
