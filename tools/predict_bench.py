"""Times gmrm_predict (the --predict sums) on a UKB-shaped slice with genotypes generated on the device -- prepared for round 2,
not yet run.  Prints markers/s and the HBM rate the three passes over the genotypes amount to (N/4 bytes per marker and pass).

    python tools/predict_bench.py [--markers 131072] [--blocks 1] [--individuals 458000] [--reps 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmrm_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--individuals", type=int, default=458000)
    ap.add_argument("--markers", type=int, default=131072)
    ap.add_argument("--blocks", type=int, default=1, help="marker blocks = ranks of the reference (vranks)")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    N, M, R = a.individuals, a.markers, a.blocks
    rng = np.random.default_rng(1)
    e = api.Engine(N=N, Mt=M, T=1, vranks=R)
    e.generate_bed(seed=1, missing_rate=0.005)
    e.finalize_bed()
    y = rng.normal(size=N)
    y = (y - y.mean()) / y.std()
    mask4 = np.full((N + 3) // 4, 0x0F, dtype=np.uint8)
    if N % 4:
        mask4[-1] = (1 << (N % 4)) - 1
    e.set_phenotype(0, y, mask4, N)
    e.compute_marker_stats()
    beta = rng.normal(0, 0.01, size=M) * (rng.random(M) < 0.05)
    e.predict(0, y, beta)                                    # warm-up (allocations, first launches)
    ts = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        out = e.predict(0, y, beta)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    passes = 3
    print(json.dumps({"what": "gmrm_predict", "N": N, "markers": M, "blocks": R, "seconds": t, "markers_per_s": M / t,
                      "hbm_gbs_3_passes": passes * M * ((N + 3) // 4) / t / 1e9, "pval_min": float(np.nanmin(out["pval"])),
                      "note": "host wall clock around the C-ABI call: uploads of y / beta and read-back of 4 x M doubles included"}))
    e.close()


if __name__ == "__main__":
    main()
