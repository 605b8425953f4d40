import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz")) as f:
        return json.load(f)


def bad_states_text(block):
    """The reference's bad-states fixtures, recorded verbatim in tests/golden/ by make_golden.py."""
    return load_golden("bad_states")[{0: "up_left", 2: "down_right"}[block]]


@pytest.fixture(scope="session")
def golden():
    return load_golden
