import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_terminal_summary(terminalreporter):
    """Raw parity statistics of every check_against call of the run (tests/cases.py), next to the pass/fail line."""
    try:
        from cases import PARITY_LOG
    except ImportError:
        return
    if PARITY_LOG:
        terminalreporter.section("parity statistics (error / max|reference|)")
        for line in PARITY_LOG:
            terminalreporter.write_line(line)
