"""CPU: the oracle's restatement of Bayes::predict (src/bayes.cpp:14-284) against the reference's own .mlma files --
the committed ones (tests/golden/g1/out{1,3}, written by tests/golden/make_golden.py from the reference's .bet
histories under 1 and 3 ranks) and, where oracle/_ref/gmrm_ref exists, a live run on a fresh dataset.

The .mlma prints 15 decimals; the restatement sums in the reference's order, so the bar is 1e-13 absolute."""
import os

import numpy as np
import pytest

from gmrm_b200 import synth

TOL = 1e-13


def keep_flags(bim, ref_bim):
    """bayes.cpp:286-316: ids of the .bim in row order; a marker is kept if its id occurs in the reference .bim."""
    ids = [l.split()[1] for l in open(bim) if l.strip()]
    ref = {l.split()[1]: i for i, l in enumerate(open(ref_bim)) if l.strip()}
    return ids, ref, np.array([i in ref for i in ids], dtype=np.uint8)


def check_against_mlma(oracle, inp, t, R, bet_path, mlma_path, bim, ref_bim):
    ids, ref, keep = keep_flags(bim, ref_bim)
    _, hist = oracle.read_bet(bet_path)
    mave, msig = oracle.marker_stats(inp["bed"], inp["N"], inp["mask4"][t], int(inp["nonas"][t]))
    res = oracle.predict(inp["bed"], inp["mask4"][t], int(inp["nonas"][t]), inp["eps0"][t], mave, msig, hist,
                         N=inp["N"], R=R, keep=keep)
    rows = oracle.read_mlma(mlma_path)
    assert len(rows) == int(keep.sum())                          # unmatched ids are dropped (bayes.cpp:224-229)
    assert os.path.getsize(mlma_path) == 123 * len(rows)         # fixed-width lines (LLEN - 1, bayes.cpp:218)
    idx = np.array([r[1] for r in rows])
    assert np.array_equal(idx, np.flatnonzero(keep))             # marker order, ranks concatenated (bayes.cpp:240-252)
    assert [r[0] for r in rows] == [ids[i] for i in idx]
    assert [r[2] for r in rows] == [ref[ids[i]] for i in idx]    # index in the reference .bim
    for c, name in ((3, "beta"), (4, "tdist"), (5, "se"), (6, "pval")):
        got = res[name][idx]
        assert np.all(np.isfinite(got))
        np.testing.assert_allclose(got, np.array([r[c] for r in rows]), rtol=0, atol=TOL, err_msg=name)
    assert np.all(np.isnan(res["beta"][keep == 0]))
    return res


@pytest.mark.parametrize("R", [1, 3])
@pytest.mark.parametrize("t", [0, 1])
def test_predict_matches_committed_reference_mlma(oracle, g1, R, t):
    d = g1["dir"]
    res = check_against_mlma(oracle, g1, t, R, os.path.join(d, f"out{R}", f"syn_t{t}.bet"),
                             os.path.join(d, f"out{R}", f"syn_t{t}.mlma"), os.path.join(d, "syn.bim"), os.path.join(d, "ref.bim"))
    assert res["sigma"].shape == (R,)


def test_one_rank_leaves_the_phenotype_untouched(oracle, g1):
    # bayes.cpp:146-147 removes only the OTHER ranks' genetic values: with one rank y_k == y, whatever the betas
    _, hist = oracle.read_bet(os.path.join(g1["dir"], "out1", "syn_t0.bet"))
    mave, msig = oracle.marker_stats(g1["bed"], g1["N"], g1["mask4"][0], int(g1["nonas"][0]))
    a = oracle.predict(g1["bed"], g1["mask4"][0], int(g1["nonas"][0]), g1["eps0"][0], mave, msig, hist, N=g1["N"], R=1)
    b = oracle.predict(g1["bed"], g1["mask4"][0], int(g1["nonas"][0]), g1["eps0"][0], mave, msig, 3.0 * hist, N=g1["N"], R=1)
    assert np.array_equal(a["beta"], b["beta"]) and np.array_equal(a["pval"], b["pval"])
    np.testing.assert_allclose(b["g"], 3.0 * a["g"], rtol=1e-12, atol=1e-15)
    y = g1["eps0"][0][: g1["N"]]
    assert abs(a["sigma"][0] - float(np.dot(y, y)) / int(g1["nonas"][0])) < 1e-12


@pytest.mark.parametrize("R,N,M", [(1, 64, 90), (2, 250, 301), (5, 1001, 257)])
def test_predict_matches_live_reference(oracle, tmp_path, R, N, M):
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/gmrm_ref not built")
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=2, n_groups=1, na_rate=0.03, missing_rate=0.02, seed=40 + R)
    p = d["paths"]
    out = str(tmp_path / "out")
    oracle.run_reference(str(tmp_path), p["bed"], p["dim"], p["phen"], p["gri"], p["grm"], out, iterations=3, seed=9, nranks=R)
    bim, ref = str(tmp_path / "syn.bim"), str(tmp_path / "ref.bim")
    with open(bim, "w") as f:
        f.writelines(f"1 rs{i} 0 {i} A G\n" for i in range(M))
    with open(ref, "w") as f:                                   # reversed order, every 50th id replaced
        f.writelines(f"1 {'rs' if i % 50 else 'gone'}{i} 0 {i} A G\n" for i in reversed(range(M)))
    oracle.run_reference(str(tmp_path), p["bed"], p["dim"], p["phen"], p["gri"], p["grm"], out, iterations=3, seed=9, nranks=R,
                         extra=("--predict", "--bim-file", bim, "--ref-bim-file", ref))
    inp = oracle.load_inputs(p["bed"], p["dim"], p["phen"], p["gri"], p["grm"])
    for t in range(2):
        stem = os.path.splitext(os.path.basename(p["phen"][t]))[0]
        check_against_mlma(oracle, inp, t, R, os.path.join(out, stem + ".bet"), os.path.join(out, stem + ".mlma"), bim, ref)
