"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the
committed golden fixtures.  Tolerances:
  * .bed transcode round trip, decode tables, NA masks, component indicators: bit-exact
  * marker statistics, dot products, residual updates: 1e-12 relative to the vector's scale
    (the reference itself is built -Ofast, so its own low-order digits move with thread count)
  * replayed beta / sigma / pi trajectories: 1e-9 relative (north-star asks 1e-6)
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from gmrm_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from gmrm_b200 import api as A
    A.lib()
    return A


def make_case(oracle, tmp, *, N, M, T=1, G=1, na_rate=0.0, missing_rate=0.0, seed=1):
    d = synth.write_dataset(str(tmp), N=N, M=M, n_traits=T, n_groups=G, na_rate=na_rate, missing_rate=missing_rate, seed=seed)
    p = d["paths"]
    inp = oracle.load_inputs(p["bed"], p["dim"], p["phen"], p["gri"], p["grm"])
    inp["paths"] = p
    return inp


def engine_for(api, inp, *, vranks=1, nsm=0, seed=0, shuffle=True):
    G, K = inp["cva"].shape
    T = inp["eps0"].shape[0]
    e = api.Engine(N=inp["N"], Mt=inp["Mt"], T=T, G=G, K=K, vranks=vranks, seed=seed, nsm=nsm, shuffle=shuffle)
    e.upload_bed(inp["bed"])
    e.finalize_bed()
    for t in range(T):
        e.set_phenotype(t, inp["eps0"][t], inp["mask4"][t], int(inp["nonas"][t]))
    e.set_groups(inp["group_index"], inp["cva"])
    e.compute_marker_stats()
    return e


# ------------------------------------------------------------------ layout / decode: bit-exact
@pytest.mark.parametrize("N,nsm", [(203, 1), (512, 1), (1024, 1), (1300, 1), (4 * 128 * 3 + 1, 1), (2500, 1), (3000, 1),
                                   (128 * 28 * 2 - 5, 2), (128 * 32 * 2, 2), (20000, 0), (7, 1)])
def test_bed_round_trip_bit_exact(api, N, nsm):
    rng = np.random.default_rng(N)
    M = 37
    mbytes = (N + 3) // 4
    bed = rng.integers(0, 256, size=(M, mbytes), dtype=np.uint8)       # every code incl. missing, arbitrary pad bits
    e = api.Engine(N=N, Mt=M, nsm=nsm)
    e.upload_bed(bed)
    assert np.array_equal(e.download_bed(), bed)
    e.close()


def test_chunked_upload_from_pinned_and_pageable_memory(api, monkeypatch):
    """gmrm_upload_bed with several staging chunks (double-buffered H2D + transcode), from a pinned (gmrm_host_alloc)
    and from a pageable source, in one call and in two calls: bit-exact round trip, missing lists included."""
    monkeypatch.setenv("GMRM_STAGE_MB", "1")
    rng = np.random.default_rng(3)
    N, M = 20000, 1000                                   # 5,000 B per column -> 209 markers per 1 MB chunk
    mbytes = (N + 3) // 4
    bed = rng.integers(0, 256, size=(M, mbytes), dtype=np.uint8)
    pinned = api.host_array(bed.shape)
    pinned[...] = bed
    for src in (pinned, bed):
        e = api.Engine(N=N, Mt=M)
        e.upload_bed(src)
        assert np.array_equal(e.download_bed(), bed)
        e.close()
    e = api.Engine(N=N, Mt=M)
    e.upload_bed(pinned[:418])                           # two whole chunks first, the rest in a second call
    e.upload_bed(pinned[418:], marker_begin=418)
    assert np.array_equal(e.download_bed(), bed)
    e.close()


def test_decode_matches_reference_tables(api, oracle):
    # all 256 byte values in every byte position class, against dotp_lut_a/b as shipped by the reference
    raw = np.fromfile(os.path.join(GOLDEN, "lut_ref.bin"))
    lut_a, lut_b, lut_na = raw[:1024], raw[1024:2048], raw[2048:]
    N = 7000                              # 28 rows over two CTAs
    mbytes = N // 4
    M = 256
    bed = np.empty((M, mbytes), dtype=np.uint8)
    for j in range(M):
        bed[j] = (np.arange(mbytes) * 7 + j) % 256
    e = api.Engine(N=N, Mt=M, nsm=2)
    e.upload_bed(bed)
    for j in (0, 1, 77, 255):
        a, b = e.decode_marker(j)
        assert np.array_equal(a, lut_a.reshape(256, 4)[bed[j]].reshape(-1)[:N])
        assert np.array_equal(b, lut_b.reshape(256, 4)[bed[j]].reshape(-1)[:N])
    # NA mask nibbles against na_lut
    rng = np.random.default_rng(5)
    mask4 = rng.integers(0, 16, size=mbytes, dtype=np.uint8)
    nonas = int(sum(bin(int(x)).count("1") for x in mask4))
    e.finalize_bed()
    e.set_phenotype(0, np.zeros(N), mask4, nonas)
    assert np.array_equal(e.decode_namask(0), lut_na.reshape(16, 4)[mask4].reshape(-1)[:N])
    e.close()


def test_generated_bed_is_valid_plink(api):
    e = api.Engine(N=1001, Mt=64, nsm=1)
    e.generate_bed(seed=3, missing_rate=0.01)
    bed = e.download_bed()
    codes = np.stack([(bed >> (2 * k)) & 3 for k in range(4)], axis=-1).reshape(64, -1)
    assert (codes[:, 1001:] == 0).all()                     # PLINK pads with 00
    miss = (codes[:, :1001] == 1).mean()
    assert 0.003 < miss < 0.03
    dos = np.select([codes[:, :1001] == 0, codes[:, :1001] == 2, codes[:, :1001] == 3], [2, 1, 0], 0)
    maf = dos.mean(axis=1) / 2
    assert maf.min() > 0.01 and maf.max() < 0.6
    e.close()


# ------------------------------------------------------------------ statistics, dot, axpy
@pytest.mark.parametrize("N,M,T,nsm,na,miss", [(203, 120, 2, 1, 0.05, 0.02), (3000, 64, 1, 1, 0.0, 0.0),
                                               (7001, 96, 3, 2, 0.01, 0.005), (20000, 200, 1, 0, 0.0, 0.01),
                                               (128 * 32 - 3, 40, 4, 1, 0.02, 0.0)])
def test_marker_stats_dot_and_update(api, oracle, tmp_path, N, M, T, nsm, na, miss):
    inp = make_case(oracle, tmp_path, N=N, M=M, T=T, na_rate=na, missing_rate=miss, seed=N % 97)
    e = engine_for(api, inp, nsm=nsm)
    rng = np.random.default_rng(1)
    eps = inp["eps0"].copy()
    for t in range(T):
        mave_o, msig_o = oracle.marker_stats(inp["bed"], N, inp["mask4"][t], int(inp["nonas"][t]))
        mave, msig = e.marker_stats(t)
        np.testing.assert_allclose(mave, mave_o, rtol=1e-13)
        np.testing.assert_allclose(msig, msig_o, rtol=1e-12)
    ids = np.arange(M, dtype=np.int32)
    got = e.dot_products(ids)
    for t in range(T):
        mave_o, msig_o = oracle.marker_stats(inp["bed"], N, inp["mask4"][t], int(inp["nonas"][t]))
        want = np.array([oracle.dot(inp["bed"][j], eps[t], mave_o[j], msig_o[j]) for j in range(M)])
        scale = np.abs(want).max()
        assert np.abs(got[:, t] - want).max() <= 1e-12 * scale
    # a few published updates, then dots again (exercises the update phase of the step kernel and the residual sums)
    for n in range(6):
        t = n % T
        j = int(rng.integers(0, M))
        db = float(rng.normal(0, 0.05))
        mave_o, msig_o = oracle.marker_stats(inp["bed"][j:j + 1], N, inp["mask4"][t], int(inp["nonas"][t]))
        oracle.update_eps(eps[t], inp["mask4"][t], inp["bed"][j], db, mave_o[0], msig_o[0])
        e.apply_update(t, j, db)
    for t in range(T):
        # the GPU applies the update with ITS msig (count-based, 1e-13 from the oracle's summed one)
        np.testing.assert_allclose(e.epsilon(t), eps[t][:N], rtol=0, atol=1e-12)
    got = e.dot_products(ids[::3])
    for t in range(T):
        mave_o, msig_o = oracle.marker_stats(inp["bed"], N, inp["mask4"][t], int(inp["nonas"][t]))
        want = np.array([oracle.dot(inp["bed"][j], eps[t], mave_o[j], msig_o[j]) for j in ids[::3]])
        assert np.abs(got[:, t] - want).max() <= 1e-12 * np.abs(want).max()
    e.close()


@pytest.mark.parametrize("N,nsm,T,M", [(1024 * 5 - 1, 1, 1, 203), (1024 * 13, 1, 1, 40), (256 * 7 + 3, 2, 2, 77), (20000, 0, 1, 203),
                                        (777, 1, 3, 33), (256 * 4, 1, 4, 50), (256 * 3 - 2, 1, 5, 19), (3000, 1, 7, 21),
                                        (256 * 11 + 9, 3, 2, 130), (100000, 0, 1, 48)])
def test_step_kernel_shapes(api, oracle, tmp_path, N, nsm, T, M):
    """The table-lookup step kernel against the oracle's dot products over the shapes that change its code path:
    1..5 rows per pass, several passes per CTA, CTAs without rows, 1..5 traits per launch and trait chunks (T=7),
    marker counts that are not a multiple of the 16-marker batch."""
    inp = make_case(oracle, tmp_path, N=N, M=M, T=T, na_rate=0.01 if T > 1 else 0.0, missing_rate=0.01, seed=N % 31 + T)
    e = engine_for(api, inp, nsm=nsm)
    got = e.dot_products(np.arange(M, dtype=np.int32))
    for t in range(T):
        mave_o, msig_o = oracle.marker_stats(inp["bed"], N, inp["mask4"][t], int(inp["nonas"][t]))
        want = np.array([oracle.dot(inp["bed"][j], inp["eps0"][t], mave_o[j], msig_o[j]) for j in range(M)])
        assert np.abs(got[:, t] - want).max() <= 1e-12 * np.abs(want).max()
    e.close()


# ------------------------------------------------------------------ replay trajectories
def replay_dict(res, i):
    return {"perm": res["perm"][i], "u": res["u"][i], "z": res["z"][i], "mu_draw": res["mu_draw"][i],
            "sigg_unit": res["sigg_unit"][i], "pi_unit": res["pi_unit"][i], "sige_unit": res["sige_unit"][i]}


def run_replay(api, inp, res, R, iters, nsm=0):
    e = engine_for(api, inp, vranks=R, nsm=nsm)
    e.init_chain(res["sigmag_init"])
    hist = []
    T = inp["eps0"].shape[0]
    for i in range(iters):
        e.run_iteration(i + 1, replay_dict(res, i))
        st = e.state()
        st["betas"] = np.stack([e.betas(t) for t in range(T)])
        st["comp"] = np.stack([e.components(t) for t in range(T)])
        hist.append(st)
    eps = np.stack([e.epsilon(t) for t in range(T)])
    e.close()
    return hist, eps


def check_traj(hist, res, rtol):
    for i, st in enumerate(hist):
        assert np.array_equal(st["comp"], res["comp"][i]), f"components differ in iteration {i + 1}"
        np.testing.assert_allclose(st["betas"], res["betas"][i], rtol=rtol, atol=1e-13)
        np.testing.assert_allclose(st["sigmag"], res["sigmag"][i], rtol=rtol)
        np.testing.assert_allclose(st["sigmae"], res["sigmae"][i], rtol=rtol)
        np.testing.assert_allclose(st["pi"], res["pi"][i], rtol=rtol)
        np.testing.assert_allclose(st["mu"], res["mu"][i], rtol=rtol, atol=1e-15)
        assert np.array_equal(st["m0"], res["m0"][i])


@pytest.mark.parametrize("R", [1, 3])
def test_replay_golden_reference_variates(api, oracle, g1, R):
    """The reference's OWN variates (committed logs) through the GPU path: trajectories of the first 5
    iterations must match the reference's .bet/.cpn/.csv (2 traits with NAs, missing genotypes, 2 groups)."""
    res = oracle.gibbs(g1["bed"], g1["eps0"], g1["mask4"], g1["nonas"], g1["group_index"], g1["cva"], N=g1["N"],
                       R=R, iterations=5, rng_mode=0, replay_dir=os.path.join(g1["dir"], f"log{R}"))
    hist, eps = run_replay(api, g1, res, R, 5, nsm=1)
    check_traj(hist, res, 1e-9)
    for t in range(2):
        _, bet = oracle.read_bet(os.path.join(g1["dir"], f"out{R}", f"syn_t{t}.bet"))
        _, cpn = oracle.read_cpn(os.path.join(g1["dir"], f"out{R}", f"syn_t{t}.cpn"))
        for i in range(5):
            assert np.array_equal(hist[i]["comp"][t], cpn[i])
            np.testing.assert_allclose(hist[i]["betas"][t], bet[i], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(eps, res["eps_final"][:, : g1["N"]], rtol=0, atol=1e-11)


@pytest.mark.parametrize("N,M,T,G,R,nsm", [(5000, 800, 1, 1, 1, 0), (3001, 640, 2, 3, 16, 2), (20000, 2000, 1, 1, 64, 0)])
def test_replay_live_reference(api, oracle, tmp_path, N, M, T, G, R, nsm):
    """Same, with variates logged by the reference binary run on this host (oracle/_ref travels with the repo)."""
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/gmrm_ref not present")
    inp = make_case(oracle, tmp_path, N=N, M=M, T=T, G=G, na_rate=0.01 if T > 1 else 0.0, missing_rate=0.002, seed=N % 89)
    p = inp["paths"]
    log = str(tmp_path / "log")
    oracle.run_reference(str(tmp_path), p["bed"], p["dim"], p["phen"], p["gri"], p["grm"], str(tmp_path / "out"),
                         iterations=4, seed=9, nranks=R, log_dir=log, timeout=1200)
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R,
                       iterations=4, rng_mode=0, replay_dir=log)
    hist, _ = run_replay(api, inp, res, R, 4, nsm=nsm)
    check_traj(hist, res, 1e-8)


def test_replay_c2_full_size_first_10_iterations(api, oracle, tmp_path):
    """BASELINE.json config C2 at full size (N=20,000, M=50,000, 1 trait, 1 group): the reference binary runs 10
    iterations with 64 ranks on this host, its logged variates are replayed through the GPU path with 64 virtual
    ranks, and the beta / sigma / pi trajectories must agree to 1e-6 relative (the north-star bound; observed ~1e-11)."""
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/gmrm_ref not present")
    N, M, R, iters = 20000, 50000, 64, 10
    inp = make_case(oracle, tmp_path, N=N, M=M, T=1, G=1, missing_rate=0.001, seed=21)
    p = inp["paths"]
    log = str(tmp_path / "log")
    oracle.run_reference(str(tmp_path), p["bed"], p["dim"], p["phen"], p["gri"], p["grm"], str(tmp_path / "out"),
                         iterations=iters, seed=171014, nranks=R, log_dir=log, timeout=2400)
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R,
                       iterations=iters, rng_mode=0, replay_dir=log)
    hist, _ = run_replay(api, inp, res, R, iters)
    check_traj(hist, res, 1e-6)
    # and against the reference's own output files
    stem = os.path.splitext(os.path.basename(p["phen"][0]))[0]
    _, bet = oracle.read_bet(str(tmp_path / "out" / (stem + ".bet")))
    _, cpn = oracle.read_cpn(str(tmp_path / "out" / (stem + ".cpn")))
    for i in range(iters):
        assert np.array_equal(hist[i]["comp"][0], cpn[i])
        np.testing.assert_allclose(hist[i]["betas"][0], bet[i], rtol=1e-6, atol=1e-12)


# ------------------------------------------------------------------ production (Philox) streams
# the last two cases sample more than 2,048 (marker, trait) pairs per step: partial sums in global memory, several list segments
@pytest.mark.parametrize("N,M,T,G,R,nsm", [(2000, 500, 1, 1, 1, 1), (4100, 900, 2, 2, 32, 2), (20000, 3000, 1, 1, 128, 0),
                                           (9000, 9000, 1, 1, 3000, 0), (3001, 2600, 2, 2, 1290, 3)])
def test_production_streams_match_oracle(api, oracle, tmp_path, N, M, T, G, R, nsm):
    """Counter-based device RNG: the oracle follows the same Philox streams on the CPU, so whole
    trajectories (permutation, component draws, betas, variances) are comparable for any seed."""
    inp = make_case(oracle, tmp_path, N=N, M=M, T=T, G=G, na_rate=0.01, missing_rate=0.003, seed=N % 83)
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R,
                       iterations=4, rng_mode=1, seed=4242)
    e = engine_for(api, inp, vranks=R, nsm=nsm, seed=4242)
    e.init_chain(None)
    hist = []
    for i in range(4):
        e.run_iteration(i + 1)
        st = e.state()
        st["betas"] = np.stack([e.betas(t) for t in range(T)])
        st["comp"] = np.stack([e.components(t) for t in range(T)])
        hist.append(st)
    e.close()
    check_traj(hist, res, 1e-8)


def test_staged_outputs_equal_direct_reads(api, oracle, tmp_path):
    """gmrm_stage_outputs / gmrm_fetch_outputs (asynchronous read-back overlapped with the next iteration) hand out
    exactly what gmrm_get_betas / gmrm_get_components returned for the staged iteration."""
    inp = make_case(oracle, tmp_path, N=3001, M=640, T=2, G=2, na_rate=0.01, missing_rate=0.005, seed=5)
    e = engine_for(api, inp, vranks=16, seed=11)
    e.init_chain(None)
    prev = None
    for i in range(4):
        e.run_iteration(i + 1)
        if prev is not None:                              # staged after iteration i, fetched after iteration i+1 ran
            for t in range(2):
                b, c = e.fetch_outputs(t)
                assert np.array_equal(b, prev[t][0]) and np.array_equal(c, prev[t][1])
        prev = [(e.betas(t), e.components(t)) for t in range(2)]
        e.stage_outputs()
    for t in range(2):
        b, c = e.fetch_outputs(t)
        assert np.array_equal(b, prev[t][0]) and np.array_equal(c, prev[t][1])
    e.close()


def test_posterior_recovers_simulated_effects(api, oracle, tmp_path):
    """Long-ish chain on simulated data (data_sim.R recipe): posterior means of the GPU chain and of the
    oracle chain (different seeds) agree within Monte Carlo error, and h2 lands near the simulated 0.5."""
    N, M = 4000, 1000
    inp = make_case(oracle, tmp_path, N=N, M=M, seed=2)
    iters, burn = 160, 60
    e = engine_for(api, inp, vranks=8, seed=1)
    e.init_chain(None)
    h2, bsum = [], np.zeros(M)
    for i in range(iters):
        e.run_iteration(i + 1)
        if i >= burn:
            st = e.state()
            h2.append(st["sigmag"].sum() / (st["sigmag"].sum() + st["sigmae"][0]))
            bsum += e.betas(0)
    e.close()
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=8,
                       iterations=iters, rng_mode=1, seed=2)
    h2_o = res["sigmag"][burn:, 0].sum(axis=1) / (res["sigmag"][burn:, 0].sum(axis=1) + res["sigmae"][burn:, 0])
    assert abs(np.mean(h2) - h2_o.mean()) < 0.05
    assert 0.3 < np.mean(h2) < 0.7
    bm, bo = bsum / (iters - burn), res["betas"][burn:, 0].mean(axis=0)
    assert np.corrcoef(bm, bo)[0, 1] > 0.9
