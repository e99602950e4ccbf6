"""GPU, BASELINE.json's full N (458,000 individuals; a slice of the markers): the parity checks that do not need a
full-size oracle run -- bit-exact transcode round trip, direct oracle parity on a sample of columns, and the
size-independent properties of the path (linearity of the dot products in the residuals, and the exact effect
of one published update on every dot product)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, M = 458_000, 4096


@pytest.fixture(scope="module")
def api():
    from gmrm_b200 import api as A
    A.lib()
    return A


def full_mask(n):
    m = np.full((n + 3) // 4, 0xF, dtype=np.uint8)
    if n % 4:
        m[-1] = (1 << (n % 4)) - 1
    return m


def standardise(y):
    y = y - y.mean()
    return y * np.sqrt((y.size - 1) / (y ** 2).sum())


def test_round_trip_random_bytes_full_width(api):
    rng = np.random.default_rng(11)
    mbytes = (N + 3) // 4
    bed = rng.integers(0, 256, size=(64, mbytes), dtype=np.uint8)      # every code incl. missing, in every position
    e = api.Engine(N=N, Mt=64)
    e.upload_bed(bed)
    assert np.array_equal(e.download_bed(), bed)
    e.close()


def test_dot_parity_linearity_and_update_identity_full_width(api, oracle):
    rng = np.random.default_rng(12)
    e = api.Engine(N=N, Mt=M, vranks=1)
    e.generate_bed(seed=7, missing_rate=0.002)
    e.finalize_bed()
    e.set_groups(np.zeros(M, dtype=np.int32), np.array([[0.0, 1e-4, 1e-3, 1e-2]]))
    mask4 = full_mask(N)
    ids = np.arange(M, dtype=np.int32)
    a, b = standardise(rng.normal(size=N)), standardise(rng.normal(size=N))
    dots = {}
    for name, eps in (("a", a), ("b", b), ("ab", a + b)):
        e.set_phenotype(0, eps, mask4, N)
        e.compute_marker_stats()
        dots[name] = e.dot_products(ids)[:, 0]
    scale = np.abs(dots["ab"]).max()
    # linearity in the residuals (the table build and the 148 x 3 partial sums are exact up to fp64 rounding)
    assert np.abs(dots["a"] + dots["b"] - dots["ab"]).max() <= 1e-11 * scale
    # direct parity with the oracle (Bayes::dot_product restated) on a sample of columns, incl. missing genotypes
    sample = rng.choice(M, 24, replace=False)
    bed = e.download_bed()
    mave, msig = e.marker_stats(0)
    eps = np.zeros(4 * ((N + 3) // 4)); eps[:N] = a + b
    for j in sample:
        mo, so = oracle.marker_stats(bed[j:j + 1], N, mask4, N)
        # msig: the GPU forms it from exact integer counts, the reference / oracle from a 458k-term floating sum
        # (phenotype.cpp:544-548) whose own rounding is ~3e-12 at this N
        assert abs(mave[j] - mo[0]) <= 1e-13 and abs(msig[j] - so[0]) <= 1e-10 * so[0]
        want = oracle.dot(bed[j], eps, mo[0], so[0])
        assert abs(dots["ab"][j] - want) <= 1e-11 * scale
    # one published update of marker j:  eps += x_j * (dbeta*msig_j)  with x_j centred, 0 at missing; so every dot
    # product moves by  msig_k * dbeta*msig_j * sum_i x~_ki x~_ji, which the oracle's update gives for the sample
    j, db = int(sample[0]), 0.03
    e.apply_update(0, j, db)
    after = e.dot_products(ids)[:, 0]
    mo, so = oracle.marker_stats(bed[j:j + 1], N, mask4, N)
    oracle.update_eps(eps, mask4, bed[j], db, mo[0], so[0])
    np.testing.assert_allclose(e.epsilon(0), eps[:N], rtol=0, atol=1e-12)
    for k in sample[:8]:
        mk, sk = oracle.marker_stats(bed[k:k + 1], N, mask4, N)
        assert abs(after[k] - oracle.dot(bed[k], eps, mk[0], sk[0])) <= 1e-11 * scale
    e.close()
