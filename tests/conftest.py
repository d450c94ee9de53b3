import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def al_profiles():
    return load_golden("al_profiles.json")


@pytest.fixture(scope="session")
def al_fitch():
    return load_golden("al_fitch.json")


@pytest.fixture(scope="session")
def al_greedy():
    return load_golden("al_greedy.json")


@pytest.fixture(scope="session")
def random_fitch():
    return load_golden("random_fitch.json")
