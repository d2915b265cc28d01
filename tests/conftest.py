import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fixture_input():
    d = np.load(os.path.join(GOLDEN, "fixture_input.npz"))
    return {"aln": d["aln"], "pos": d["pos"], "names": [str(x) for x in d["names"]]}


@pytest.fixture(scope="session")
def fixture_expected():
    return dict(np.load(os.path.join(GOLDEN, "fixture_expected.npz")))


@pytest.fixture(scope="session")
def fixture_snp(fixture_expected):
    import ldw_oracle as O
    e = fixture_expected
    return O.snp_dat_from_codes(e["codes"], e["relaxed_POS"], 50000)
