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


def q_close(got, want):
    """float32 table entry against the reference's float64 value (north star: 1e-5 relative).  Purely relative for
    |value| >= 1.  The few entries below 1 in magnitude (16 of 3 230 in the recorded 2-car run, down to 0.017) are
    differences of terms of magnitude >= 10 (reward -10 / +10 010, discount 0.15): float32 leaves them an ABSOLUTE
    error of ~1e-6 whatever their size (2 of them are 3e-5 off relatively), so they are held to 2e-6 absolute."""
    got, want = float(got), float(want)
    return abs(got - want) <= (1e-5 * abs(want) if abs(want) >= 1.0 else 2e-6)
