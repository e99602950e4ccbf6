"""CPU: the oracle (oracle/oracle.cpp) against the committed outputs of the reference itself.

tests/golden/g1 holds inputs, the reference's logged random variates and the reference's own
.bet/.cpn/.csv for 1 and 3 ranks (tests/golden/make_golden.py).  Replaying the variates through
the restatement must reproduce those files, and every (mean, sd)/(shape, scale) the reference
passed to its distributions must equal what the restatement computes at the same point."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_decode_tables_match_reference_headers(oracle):
    raw = np.fromfile(os.path.join(GOLDEN, "lut_ref.bin"))
    a, b, na = oracle.decode_tables()
    assert np.array_equal(raw[:1024], a)       # src/dotp_lut.hpp dotp_lut_a
    assert np.array_equal(raw[1024:2048], b)   # dotp_lut_b
    assert np.array_equal(raw[2048:], na)      # src/na_lut.hpp na_lut


@pytest.mark.parametrize("R", [1, 3])
def test_replay_reproduces_reference_outputs(oracle, g1, R):
    res = oracle.gibbs(g1["bed"], g1["eps0"], g1["mask4"], g1["nonas"], g1["group_index"], g1["cva"], N=g1["N"],
                       R=R, iterations=5, rng_mode=0, replay_dir=os.path.join(g1["dir"], f"log{R}"))
    assert res["n_log_checked"] > 1500
    assert res["max_log_relerr"] < 1e-10        # per-draw parameters as the reference computed them
    for t in range(2):
        its, bet = oracle.read_bet(os.path.join(g1["dir"], f"out{R}", f"syn_t{t}.bet"))
        _, cpn = oracle.read_cpn(os.path.join(g1["dir"], f"out{R}", f"syn_t{t}.cpn"))
        csv = oracle.read_csv(os.path.join(g1["dir"], f"out{R}", f"syn_t{t}.csv"))
        assert list(its) == [1, 2, 3, 4, 5]
        assert np.array_equal(cpn, res["comp"][:, t])                       # integer work: bit-exact
        np.testing.assert_allclose(res["betas"][:, t], bet, rtol=1e-11, atol=1e-14)
        for i, row in enumerate(csv):                                       # .csv prints 15 decimals
            np.testing.assert_allclose(res["sigmag"][i, t], row["sigmag"], atol=2e-15 + 1e-12 * abs(row["sigmag"]).max())
            assert abs(res["sigmae"][i, t] - row["sigmae"]) < 1e-12
            np.testing.assert_allclose(res["pi"][i, t], row["pi"], atol=1e-12)
            assert int(res["m0"][i, t].sum()) == row["m0_sum"]


def test_read_phen_edge_cases(oracle, tmp_path):
    # N % 4 != 0, an NA in the last byte, leading/trailing blanks, tabs (phenotype.cpp:587-673)
    p = tmp_path / "x.phen"
    p.write_text("1 1 0.5\n2 2 NA\n3\t3\t-1.25\n4 4 2.0  \n5 5 3.5\n6 6 NA\n")
    eps, mask4, nonas, nas = oracle.read_phen(str(p), 6)
    assert (nonas, nas) == (4, 2)
    assert list(mask4) == [0b1101, 0b0001]
    y = np.array([0.5, -1.25, 2.0, 3.5])
    c = y - y.mean()
    c *= np.sqrt(3.0 / (c ** 2).sum())
    np.testing.assert_allclose(eps[[0, 2, 3, 4]], c, rtol=1e-14)
    assert eps[1] == 0.0 and eps[5] == 0.0 and eps[6] == 0.0 and eps[7] == 0.0


def test_block_of_markers(oracle):
    # bayes.cpp:903-925: first Mt % R ranks get one more
    assert oracle.block_of_markers(10, 3, 0) == (0, 4, 4)
    assert oracle.block_of_markers(10, 3, 1) == (4, 3, 4)
    assert oracle.block_of_markers(10, 3, 2) == (7, 3, 4)
    assert oracle.block_of_markers(9, 3, 2) == (6, 3, 3)


def test_production_streams_are_rank_count_invariant_at_R1(oracle, g1):
    # Philox streams are keyed by marker, not by rank: two runs with the same seed agree bit for bit
    kw = dict(N=g1["N"], R=1, iterations=3, rng_mode=1, seed=7)
    a = oracle.gibbs(g1["bed"], g1["eps0"], g1["mask4"], g1["nonas"], g1["group_index"], g1["cva"], **kw)
    b = oracle.gibbs(g1["bed"], g1["eps0"], g1["mask4"], g1["nonas"], g1["group_index"], g1["cva"], **kw)
    assert np.array_equal(a["betas"], b["betas"])
    assert np.isfinite(a["sigmae"]).all() and (a["sigmae"] > 0).all()
    # permutation really is a permutation
    for r in range(1):
        assert sorted(a["perm"][0, r].tolist()) == list(range(g1["Mt"]))
