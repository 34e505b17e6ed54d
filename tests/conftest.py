import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    return O.Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("reference build (oracle/_ref) not available")
    return O.Ref()
