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
