"""CPU: the production random streams of gmrm_b200/csrc/gmrm_rng.h against their TARGET distributions.

The GPU kernels and the oracle share this header (so that production-stream trajectories can be compared bit for bit);
a wrong transform in it would therefore pass every parity test.  These tests compile the header for the host and check
each transform on its own: Kolmogorov-Smirnov distances and moments of u01 / box_muller / draw_gamma (the shapes the chain
uses: 0.5*(V0+m0) for sigmaG, cass+1 for the Dirichlet, 0.5*(V0+N) for sigmaE; reference src/distributions.hpp:5-61),
Beta(1,1) for the sigmaG start (bayes.cpp:326-331), Dirichlet moments through the gammas (phenotype.cpp:227-237), and
perm_at as a permutation (bijection for every n, positions uniform; replaces the shuffle of phenotype.cpp:314-323)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy import stats

from conftest import ROOT

SRC = r'''
#include "gmrm_rng.h"
using namespace gmrm;
extern "C" {
void rng_u01(uint32_t seed, int n, double* out) { for (int i = 0; i < n; i++) out[i] = draw_uniform(seed, STREAM_SAMPLER_U, 3, (uint32_t)i, 0); }
void rng_normal(uint32_t seed, int n, double* out) { for (int i = 0; i < n; i++) out[i] = draw_normal(seed, STREAM_SAMPLER_N, 7, (uint32_t)i, 1); }
void rng_gamma(double shape, uint32_t seed, int n, double* out) { for (int i = 0; i < n; i++) out[i] = draw_gamma(shape, seed, STREAM_PI, (uint32_t)(i >> 16), (uint32_t)(i & 0xffff), 0); }
void rng_perm(uint32_t n, uint32_t seed, uint32_t it, uint32_t rank, uint32_t* out) { for (uint32_t s = 0; s < n; s++) out[s] = perm_at(s, n, seed, it, rank); }
void rng_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) { U4 r = philox4x32(k0, k1, c0, c1, c2, c3); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w; }
}
'''


@pytest.fixture(scope="module")
def rng(tmp_path_factory):
    d = tmp_path_factory.mktemp("rng")
    cpp = d / "rng_host.cpp"
    cpp.write_text(SRC)
    so = d / "librng_host.so"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "gmrm_b200", "csrc"),
                    str(cpp), "-o", str(so)], check=True)
    L = C.CDLL(str(so))
    return L


def draw(f, n, *args):
    out = np.empty(n)
    f(*args, C.c_int(n), C.c_void_p(out.ctypes.data))
    return out


def test_philox_known_answers(rng):
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors: counter, key -> output)."""
    out = (C.c_uint32 * 4)()
    rng.rng_philox(0, 0, 0, 0, 0, 0, out)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    rng.rng_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    rng.rng_philox(0xa4093822, 0x299f31d0, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, out)
    assert [hex(x) for x in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_uniform(rng):
    n = 400_000
    u = draw(rng.rng_u01, n, C.c_uint32(11))
    assert u.min() > 0.0 and u.max() < 1.0
    assert stats.kstest(u, "uniform").statistic < 1.63 / np.sqrt(n)        # 1 % critical value
    assert abs(u.mean() - 0.5) < 4 * np.sqrt(1 / 12 / n)
    # serial independence of consecutive counters
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 4 / np.sqrt(n)


def test_normal(rng):
    n = 400_000
    z = draw(rng.rng_normal, n, C.c_uint32(5))
    assert stats.kstest(z, "norm").statistic < 1.63 / np.sqrt(n)
    assert abs(z.mean()) < 4 / np.sqrt(n)
    assert abs(z.var() - 1.0) < 4 * np.sqrt(2 / n)
    assert abs(stats.kurtosis(z)) < 4 * np.sqrt(24 / n)
    assert (np.abs(z) > 4.5).sum() <= 12                                   # tails exist and are not inflated (expected 2.7)


@pytest.mark.parametrize("shape", [0.3, 1.0, 2.5, 5.0e4, 2.29e5])
def test_gamma(rng, shape):
    """draw_gamma(shape) ~ Gamma(shape, 1): shapes below one (boosted), the Dirichlet's small counts, and the half-degrees
    of freedom of sigmaG (0.5*(1e-4 + m0), m0 ~ 1e5) and sigmaE (0.5*(1e-4 + 458,000))."""
    n = 200_000
    x = draw(rng.rng_gamma, n, C.c_double(shape), C.c_uint32(23))
    assert x.min() > 0.0
    assert stats.kstest(x, "gamma", args=(shape,)).statistic < 1.63 / np.sqrt(n)
    assert abs(x.mean() - shape) < 4 * np.sqrt(shape / n)
    assert abs(x.var() - shape) < 5 * shape * np.sqrt(2 / n + 6 / (shape * n))


def test_beta11_and_inverse_scaled_chisq(rng):
    """sigmaG start ~ Beta(1,1) == U(0,1) (bayes.cpp:326-331): draw_uniform is used; inv_scaled_chisq(a, b) =
    1/rgamma(a/2, 2/(a b)) (distributions.hpp:24-30) has mean a b / (a - 2)."""
    n = 200_000
    a, b = 12.0, 0.7
    unit = draw(rng.rng_gamma, n, C.c_double(0.5 * a), C.c_uint32(3))
    x = 1.0 / (unit * (1.0 / (0.5 * a * b)))                               # sampler.h: inv_scaled_chisq_from_unit
    assert abs(x.mean() - a * b / (a - 2.0)) < 5 * x.std() / np.sqrt(n)
    assert stats.kstest(0.5 * a * b / x, "gamma", args=(0.5 * a,)).statistic < 1.63 / np.sqrt(n)


def test_dirichlet_through_gammas(rng):
    """pi ~ Dirichlet(cass + 1) as normalised gammas (phenotype.cpp:227-237): component means and variances."""
    alpha = np.array([7001.0, 211.0, 38.0, 3.0])
    n = 50_000
    g = np.stack([draw(rng.rng_gamma, n, C.c_double(a), C.c_uint32(100 + k)) for k, a in enumerate(alpha)], axis=1)
    p = g / g.sum(axis=1, keepdims=True)
    a0 = alpha.sum()
    mean, var = alpha / a0, alpha * (a0 - alpha) / (a0 * a0 * (a0 + 1))
    assert np.all(np.abs(p.mean(axis=0) - mean) < 5 * np.sqrt(var / n))
    assert np.all(np.abs(p.var(axis=0) / var - 1.0) < 0.05)


@pytest.mark.parametrize("n", list(range(1, 70)) + [255, 256, 257, 488, 489, 1023, 4096, 4097, 125_000])
def test_perm_is_a_bijection(rng, n):
    out = np.empty(n, dtype=np.uint32)
    rng.rng_perm(n, 171014, 3, 5, C.c_void_p(out.ctypes.data))
    assert np.array_equal(np.sort(out), np.arange(n, dtype=np.uint32))


def test_perm_positions_are_uniform_and_streams_differ(rng):
    """Over many (iteration, rank) streams every position of a 61-element permutation takes every value equally often
    (chi-square per position), and two streams are not the same permutation."""
    n, reps = 61, 6000
    counts = np.zeros((n, n))
    out = np.empty(n, dtype=np.uint32)
    seen = set()
    for r in range(reps):
        rng.rng_perm(n, 99, r // 64 + 1, r % 64, C.c_void_p(out.ctypes.data))
        counts[np.arange(n), out] += 1
        seen.add(out.tobytes())
    assert len(seen) == reps
    chi2 = ((counts - reps / n) ** 2 / (reps / n)).sum(axis=1)             # per position, n - 1 degrees of freedom
    assert chi2.max() < stats.chi2.ppf(1 - 1e-4 / n, n - 1)
    # first-order structure: value at position s+1 given position s is not tied to it
    pair = np.zeros((n, n))
    for r in range(2000):
        rng.rng_perm(n, 7, r + 1, 0, C.c_void_p(out.ctypes.data))
        pair[out[:-1], out[1:]] += 1
    off = ~np.eye(n, dtype=bool)
    expect = 2000 * (n - 1) / (n * (n - 1))                                # every ordered pair of distinct values equally often
    chi2p = ((pair[off] - expect) ** 2 / expect).sum()
    assert not pair[~off].any() and chi2p < stats.chi2.ppf(1 - 1e-4, n * (n - 1) - 1)
