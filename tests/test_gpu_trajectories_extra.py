"""GPU: trajectories on inputs off the main road (no marker shuffling; a group without markers).  The oracle is pinned on
both branches against the live reference (tests/test_oracle_vs_reference.py).  Green on B200s since round 1's driver run."""
import numpy as np
import pytest

from test_gpu_parity import api, check_traj, engine_for, make_case  # noqa: F401  (api is a fixture)

pytestmark = pytest.mark.gpu


def test_production_streams_without_shuffle(api, oracle, tmp_path):
    """--shuffle-markers 0: every virtual rank walks its block of markers in file order (the oracle is pinned on this
    mode against the live reference, tests/test_oracle_vs_reference.py)."""
    N, M, T, R = 2500, 700, 1, 8
    inp = make_case(oracle, tmp_path, N=N, M=M, T=T, G=1, na_rate=0.01, missing_rate=0.003, seed=17)
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R,
                       iterations=3, shuffle=False, rng_mode=1, seed=7)
    e = engine_for(api, inp, vranks=R, seed=7, shuffle=False)
    e.init_chain(None)
    hist = []
    for i in range(3):
        e.run_iteration(i + 1)
        st = e.state()
        st["betas"] = np.stack([e.betas(t) for t in range(T)])
        st["comp"] = np.stack([e.components(t) for t in range(T)])
        hist.append(st)
    e.close()
    check_traj(hist, res, 1e-8)


def test_production_streams_with_an_empty_group(api, oracle, tmp_path):
    """A group without markers: sigmaG = 0 for good, no sigmaG / pi draws for it (bayes.cpp:396-400, 597-611).  The oracle
    is pinned on this branch against the live reference (tests/test_oracle_vs_reference.py)."""
    N, M, T, G, R = 3001, 640, 2, 3, 16
    inp = make_case(oracle, tmp_path, N=N, M=M, T=T, G=G, na_rate=0.01, missing_rate=0.003, seed=31)
    inp["group_index"] = np.where(inp["group_index"] == 2, 0, inp["group_index"]).astype(np.int32)
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R,
                       iterations=4, rng_mode=1, seed=99)
    e = engine_for(api, inp, vranks=R, seed=99)
    e.init_chain(None)
    hist = []
    for i in range(4):
        e.run_iteration(i + 1)
        st = e.state()
        assert st["sigmag"][:, 2].max() == 0.0
        st["betas"] = np.stack([e.betas(t) for t in range(T)])
        st["comp"] = np.stack([e.components(t) for t in range(T)])
        hist.append(st)
    e.close()
    check_traj(hist, res, 1e-8)
