"""GPU: long-chain posterior means do not depend on the number of virtual ranks within Monte Carlo error, as long as the
markers in flight per step stay a small fraction of the individuals (north star: "long-chain posterior means agreeing within
Monte Carlo error on simulated data").

A run with R virtual ranks IS the reference under `mpirun -n R` (the production-stream trajectory tests pin that, component
for component, against the oracle at the same R).  What is left to show is that the production numbers of virtual ranks give
the same posterior as a small R: two chains on data simulated as example/data_sim.R does (y = scale(X) b + e, 25 % causal
markers, h2 = 0.5), R = 8 and R = 48 (0.8 % of N), compared through batch-means standard errors.  DESIGN.md holds the table
of the full-size probes (tools/chain_probe.py: N = 20,000 and N = 458,000, R up to 16,384)."""
import numpy as np
import pytest

from gmrm_b200 import synth

pytestmark = pytest.mark.gpu


def batch_se(x, nb=20):
    x = np.asarray(x, dtype=np.float64)
    n = (len(x) // nb) * nb
    return float(x[:n].reshape(nb, -1).mean(axis=1).std(ddof=1) / np.sqrt(nb))


def run_chain(api, N, M, bed, eps0, mask4, nonas, V, iters, burn, seed):
    e = api.Engine(N=N, Mt=M, T=1, G=1, K=4, vranks=V, seed=seed)
    e.upload_bed(bed)
    e.finalize_bed()
    e.set_phenotype(0, eps0, mask4, nonas)
    e.set_groups(np.zeros(M, dtype=np.int32), np.array([[0.0, 1e-4, 1e-3, 1e-2]]))
    e.compute_marker_stats()
    e.init_chain(None)
    tr = {"h2": [], "sigmae": [], "m0": []}
    for it in range(1, iters + 1):
        e.run_iteration(it)
        st = e.state()
        sg, se = float(st["sigmag"][0].sum()), float(st["sigmae"][0])
        tr["h2"].append(sg / (sg + se)); tr["sigmae"].append(se); tr["m0"].append(float(st["m0"][0].sum()))
    e.close()
    return {k: (float(np.mean(v[burn:])), batch_se(v[burn:])) for k, v in tr.items()}


def test_posterior_means_do_not_depend_on_the_virtual_ranks():
    from gmrm_b200 import api
    N, M, iters, burn = 6000, 3000, 1400, 400
    d = synth.make_genotypes(N, M, seed=11)
    y, _ = synth.make_phenotypes(d, 1, h2=0.5, causal_frac=0.25, seed=171014)
    bed = synth.pack_bed(d)
    c = y[0] - y[0].mean()
    c *= np.sqrt((N - 1) / (c ** 2).sum())                      # Phenotype::read_file's scaling (phenotype.cpp:647-667)
    mask4 = np.full((N + 3) // 4, 0x0F, dtype=np.uint8)
    a = run_chain(api, N, M, bed, c, mask4, N, 8, iters, burn, seed=3)
    b = run_chain(api, N, M, bed, c, mask4, N, 48, iters, burn, seed=3)
    for k in ("h2", "sigmae", "m0"):
        se = np.hypot(a[k][1], b[k][1])
        assert abs(a[k][0] - b[k][0]) <= 4.0 * se, (k, a[k], b[k])
    # and both recover the simulated heritability (0.5) to within what N = 6,000 allows
    assert abs(a["h2"][0] - 0.5) < 0.08 and abs(b["h2"][0] - 0.5) < 0.08
