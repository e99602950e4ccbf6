"""Long-chain probe (posterior validity at production numbers of virtual ranks, publish rates): runs the Gibbs chain on a
simulated data set (genotypes generated on the device, phenotype y = scale(X) b + e as in bench.py / data_sim.R) for one or
more numbers of virtual ranks V and prints, per V, posterior means with batch-means standard errors of sigmaG, sigmaE, h2, pi,
the number of markers in the model, and the agreement of the posterior-mean effects with the simulated ones.

    python tools/chain_probe.py --workload c2 --vranks 64,2048,16384 --iterations 2000 --burn 500 --out gpurun_out/post
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (phenotype recipe shared with the bench)
from gmrm_b200 import api  # noqa: E402


def batch_se(x, nb=20):
    x = np.asarray(x, dtype=np.float64)
    n = (len(x) // nb) * nb
    if n < nb * 2:
        return float("nan")
    m = x[:n].reshape(nb, -1).mean(axis=1)
    return float(m.std(ddof=1) / np.sqrt(nb))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--markers", type=int, default=0)
    ap.add_argument("--vranks", default="64,2048")
    ap.add_argument("--iterations", type=int, default=2000)
    ap.add_argument("--burn", type=int, default=500)
    ap.add_argument("--causal-frac", type=float, default=0.25)
    ap.add_argument("--seed", type=int, default=171014)
    ap.add_argument("--out", default="gpurun_out/chain_probe")
    ap.add_argument("--trace-every", type=int, default=0, help="print every n-th iteration")
    a = ap.parse_args()
    w = bench.WORKLOADS[a.workload]
    N, M, T, G = w["N"], a.markers or w["M"], w["T"], w["G"]
    K = len(bench.MIXTURES)
    os.makedirs(a.out, exist_ok=True)
    results = []
    y = beta_true = None
    for V in [int(x) for x in a.vranks.split(",")]:
        V = min(V, M)
        e = api.Engine(N=N, Mt=M, T=T, G=G, K=K, vranks=V, seed=a.seed)
        e.generate_bed(seed=1)
        e.finalize_bed()
        if y is None:
            y, ncausal = bench.phenotype_from_engine(e, N, M, T, a.causal_frac, 0.5, seed=a.seed)
            beta_true, _ = bench.causal_effects(M, T, a.causal_frac, 0.5, a.seed)
        for t in range(T):
            c, mask4, nonas = bench.standardise(y[t], np.zeros(N, dtype=bool))
            e.set_phenotype(t, c, mask4, nonas)
        e.set_groups(np.zeros(M, dtype=np.int32), np.stack([np.array(bench.MIXTURES)] * G))
        e.compute_marker_stats()
        e.init_chain(None)
        tr = {k: [] for k in ("sigmag", "sigmae", "h2", "m0", "published", "ms")}
        pis = []
        bsum = np.zeros(M); nz = np.zeros(M); nkeep = 0
        t0 = time.time()
        for it in range(1, a.iterations + 1):
            e.run_iteration(it)
            st = e.state(); tm = e.timing()
            sg, se = float(st["sigmag"][0].sum()), float(st["sigmae"][0])
            tr["sigmag"].append(sg); tr["sigmae"].append(se); tr["h2"].append(sg / (sg + se)); tr["m0"].append(int(st["m0"][0].sum()))
            tr["published"].append(int(tm["published"])); tr["ms"].append(tm["iteration_ms"])
            pis.append(st["pi"][0, 0].copy())
            if it > a.burn and (it - a.burn) % 5 == 0:          # thinned read-back of the effects
                b = e.betas(0)
                bsum += b; nz += b != 0; nkeep += 1
            if a.trace_every and it % a.trace_every == 0:
                print(f"V={V} it {it}: {tm['iteration_ms']:.2f} ms published {tm['published']} sigmaE {se:.4f} sigmaG {sg:.4f} m0 {tr['m0'][-1]}", flush=True)
        wall = time.time() - t0
        e.close()
        post = slice(a.burn, None)
        bmean = bsum / max(nkeep, 1)
        pip = nz / max(nkeep, 1)
        bt = beta_true[0]
        causal = bt != 0
        slope = float((bmean * bt).sum() / (bt * bt).sum())
        r = {"V": V, "V_over_M": V / M, "N": N, "M": M, "iterations": a.iterations, "burn": a.burn, "wall_s": wall,
             "ms_per_iteration": float(np.mean(tr["ms"][post])),
             "published_per_iter_first10": tr["published"][:10], "published_per_iter_post": float(np.mean(tr["published"][post])),
             "published_frac_post": float(np.mean(tr["published"][post])) / M}
        for k in ("sigmag", "sigmae", "h2", "m0"):
            r[k] = {"mean": float(np.mean(tr[k][post])), "se": batch_se(tr[k][post])}
        P = np.array(pis)[post]
        r["pi"] = {"mean": P.mean(axis=0).tolist(), "se": [batch_se(P[:, k]) for k in range(P.shape[1])]}
        r["beta"] = {"cor_true": float(np.corrcoef(bmean, bt)[0, 1]), "slope_on_true": slope,
                     "mean_pip_causal": float(pip[causal].mean()), "mean_pip_null": float(pip[~causal].mean()),
                     "kept_draws": nkeep}
        np.save(os.path.join(a.out, f"bmean_{a.workload}_V{V}.npy"), bmean.astype(np.float32))
        json.dump(tr, open(os.path.join(a.out, f"trace_{a.workload}_V{V}.json"), "w"))
        results.append(r)
        print(json.dumps(r), flush=True)
    # agreement of the posterior-mean effects between the runs (first V is the reference run)
    if len(results) > 1:
        ref = np.load(os.path.join(a.out, f"bmean_{a.workload}_V{results[0]['V']}.npy"))
        for r in results[1:]:
            b = np.load(os.path.join(a.out, f"bmean_{a.workload}_V{r['V']}.npy"))
            r["beta"]["cor_with_first_run"] = float(np.corrcoef(ref, b)[0, 1])
            r["beta"]["slope_on_first_run"] = float((b * ref).sum() / (ref * ref).sum())
    json.dump(results, open(os.path.join(a.out, f"summary_{a.workload}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
