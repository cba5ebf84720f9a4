import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENES = os.path.join(ROOT, "scenes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Run the opt-in configurations (`replicated` levels: PTD_DN_REPL_LEVEL=3) after everything else: the round-end driver runs
    `pytest -m gpu -x`, and a failure in an opt-in mode must not hide the results of the default configuration."""
    late = [it for it in items if "replicated" in it.nodeid or "experimental" in it.nodeid]
    if late:
        ids = {id(it) for it in late}
        items[:] = [it for it in items if id(it) not in ids] + late
    # The "experimental" cases cover code paths written after round 1's GPU budget was spent: never run on a GPU yet,
    # all opt-in, and bench.py re-verifies each of them at run time before using it.  They are part of the suite only on request
    # (PTD_OPTIN_TESTS=1, as tools/gpu_round2_first.sh sets it), so that `pytest -m gpu` reports the validated configuration.
    if os.environ.get("PTD_OPTIN_TESTS") != "1":
        skip = pytest.mark.skip(reason="opt-in code path not yet validated on a GPU: set PTD_OPTIN_TESTS=1 to run it")
        for it in items:
            if "experimental" in it.nodeid or "replicated" in it.nodeid:     # (replicated levels: validated bit-exact with 8 processes, profiles/r02k_*;
                it.add_marker(skip)                                          #  the single-GPU lock-step form of that opt-in has not run since it became opt-in)


@pytest.fixture(scope="session")
def root():
    return ROOT
