"""CPU: the oracle against the reference binary run live (only where oracle/_ref was built, i.e.
where /root/reference exists, or where the prebuilt binary travelled to)."""
import os

import numpy as np
import pytest

from gmrm_b200 import synth


@pytest.mark.parametrize("R,T,G,N,M", [(1, 1, 1, 64, 90), (2, 3, 3, 250, 301), (4, 1, 2, 1001, 257)])
def test_live_reference_replay(oracle, tmp_path, R, T, G, N, M):
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/gmrm_ref not built")
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=T, n_groups=G, na_rate=0.03, missing_rate=0.02, seed=11 + R)
    p = d["paths"]
    out = str(tmp_path / "out")
    log = str(tmp_path / "log")
    oracle.run_reference(str(tmp_path), p["bed"], p["dim"], p["phen"], p["gri"], p["grm"], out, iterations=4,
                         seed=5, nranks=R, log_dir=log)
    inp = oracle.load_inputs(p["bed"], p["dim"], p["phen"], p["gri"], p["grm"])
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R,
                       iterations=4, rng_mode=0, replay_dir=log)
    assert res["max_log_relerr"] < 1e-10
    for t in range(T):
        stem = os.path.splitext(os.path.basename(p["phen"][t]))[0]
        _, bet = oracle.read_bet(os.path.join(out, stem + ".bet"))
        _, cpn = oracle.read_cpn(os.path.join(out, stem + ".cpn"))
        assert np.array_equal(cpn, res["comp"][:, t])
        np.testing.assert_allclose(res["betas"][:, t], bet, rtol=1e-11, atol=1e-14)
        csv = oracle.read_csv(os.path.join(out, stem + ".csv"))
        for i, row in enumerate(csv):
            assert abs(res["sigmae"][i, t] - row["sigmae"]) < 1e-12


def test_live_reference_replay_with_an_empty_group(oracle, tmp_path):
    """A group without markers (and, through tiny mixture variances, groups that lose all their markers) takes the
    reference through its dead-group branches: sigmaG = 0 for good once m0 == 0 or sum(cass) == 0 (bayes.cpp:396-400,
    608-611), no sigmaG / pi draw for it.  The restatement must follow the same variate stream."""
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/gmrm_ref not built")
    R, T, G, N, M = 2, 2, 3, 250, 301
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=T, n_groups=G, na_rate=0.03, missing_rate=0.02, seed=21)
    p = d["paths"]
    rows = [l.split() for l in open(p["gri"])]
    with open(p["gri"], "w") as f:                       # group 2 keeps its mixture row but loses every marker
        f.writelines(f"{a} {0 if b == '2' else b}\n" for a, b in rows)
    out, log = str(tmp_path / "out"), str(tmp_path / "log")
    oracle.run_reference(str(tmp_path), p["bed"], p["dim"], p["phen"], p["gri"], p["grm"], out, iterations=6, seed=5, nranks=R, log_dir=log)
    inp = oracle.load_inputs(p["bed"], p["dim"], p["phen"], p["gri"], p["grm"])
    assert not (inp["group_index"] == 2).any()
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R,
                       iterations=6, rng_mode=0, replay_dir=log)
    assert res["max_log_relerr"] < 1e-10
    for t in range(T):
        stem = os.path.splitext(os.path.basename(p["phen"][t]))[0]
        _, bet = oracle.read_bet(os.path.join(out, stem + ".bet"))
        _, cpn = oracle.read_cpn(os.path.join(out, stem + ".cpn"))
        assert np.array_equal(cpn, res["comp"][:, t])
        np.testing.assert_allclose(res["betas"][:, t], bet, rtol=1e-11, atol=1e-14)
        csv = oracle.read_csv(os.path.join(out, stem + ".csv"))
        for i, row in enumerate(csv):
            assert row["sigmag"][2] == 0.0 and res["sigmag"][i, t][2] == 0.0       # the empty group stays dead
            np.testing.assert_allclose(res["sigmag"][i, t], row["sigmag"], atol=1e-12)
            assert abs(res["sigmae"][i, t] - row["sigmae"]) < 1e-12
            np.testing.assert_allclose(res["pi"][i, t], row["pi"], atol=1e-12)


def test_live_reference_without_shuffle_and_with_truncated_markers(oracle, tmp_path):
    """--shuffle-markers 0 (markers visited in file order, phenotype.cpp:308-323 not called) and --trunc-markers n
    (only the first n markers of the .bed / .gri are used, dimensions.hpp:12-14)."""
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/gmrm_ref not built")
    R, T, G, N, M, keep = 3, 1, 2, 203, 180, 125
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=T, n_groups=G, na_rate=0.02, missing_rate=0.01, seed=8)
    p = d["paths"]
    out, log = str(tmp_path / "out"), str(tmp_path / "log")
    oracle.run_reference(str(tmp_path), p["bed"], p["dim"], p["phen"], p["gri"], p["grm"], out, iterations=4, seed=3, nranks=R,
                         shuffle=0, log_dir=log, extra=("--trunc-markers", str(keep)))
    inp = oracle.load_inputs(p["bed"], p["dim"], p["phen"], p["gri"], p["grm"])
    res = oracle.gibbs(inp["bed"][:keep], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"][:keep], inp["cva"], N=N, R=R,
                       iterations=4, shuffle=False, rng_mode=0, replay_dir=log)
    assert res["max_log_relerr"] < 1e-10
    stem = os.path.splitext(os.path.basename(p["phen"][0]))[0]
    its, bet = oracle.read_bet(os.path.join(out, stem + ".bet"))
    _, cpn = oracle.read_cpn(os.path.join(out, stem + ".cpn"))
    assert bet.shape == (4, keep)                                  # the history holds the truncated marker count
    assert np.array_equal(cpn, res["comp"][:, 0])
    np.testing.assert_allclose(res["betas"][:, 0], bet, rtol=1e-11, atol=1e-14)
    # without shuffling every rank walks its block in order: step s of rank r is local marker s
    blocks = [oracle.block_of_markers(keep, R, r) for r in range(R)]
    for r, (S, Mr, Mm) in enumerate(blocks):
        assert np.array_equal(res["perm"][0][r][:Mr], np.arange(Mr))
