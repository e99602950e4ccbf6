import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def g1(oracle):
    """The committed golden case (tests/golden/make_golden.py)."""
    d = os.path.join(GOLDEN, "g1")
    phens = [os.path.join(d, "syn_t0.phen"), os.path.join(d, "syn_t1.phen")]
    inp = oracle.load_inputs(os.path.join(d, "syn.bed"), os.path.join(d, "syn.dim"), phens,
                             os.path.join(d, "syn.gri"), os.path.join(d, "syn.grm"))
    inp["dir"] = d
    inp["phens"] = phens
    return inp
