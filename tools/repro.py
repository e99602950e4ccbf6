import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmrm_b200 import api
N, M = int(sys.argv[1]), int(sys.argv[2]); V = int(sys.argv[3]); nsm = int(sys.argv[4])
e = api.Engine(N=N, Mt=M, vranks=1, nsm=nsm)
print('created', flush=True)
e.generate_bed(seed=1, missing_rate=0.01)
print('generated', flush=True)
e.finalize_bed()
rng = np.random.default_rng(0)
y = rng.normal(size=N); y -= y.mean(); y /= y.std()
mask4 = np.full((N + 3) // 4, 0xF, dtype=np.uint8)
if N % 4: mask4[-1] = (1 << (N % 4)) - 1
e.set_phenotype(0, y, mask4, N)
e.set_groups(np.zeros(M, dtype=np.int32), np.array([[0.0, 1e-4, 1e-3, 1e-2]]))
print('phen/groups set', flush=True)
e.compute_marker_stats()
print('stats ok', flush=True)
t = time.time(); r = e.dot_products(np.arange(V, dtype=np.int32) % M); print("dot ok", r[:3, 0], time.time() - t, flush=True)
e.apply_update(0, 3, 0.01); print("update ok", flush=True)
r = e.dot_products(np.arange(V, dtype=np.int32) % M); print("dot2 ok", r[:3, 0], flush=True)
e.close()
