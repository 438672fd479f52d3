import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLD, "reference_results.json")) as fh:
        res = json.load(fh)["cases"]
    labels = np.load(os.path.join(GOLD, "reference_labels.npz"))
    inputs = np.load(os.path.join(GOLD, "example_inputs.npz"))
    return dict(cases=res, labels=labels, inputs=inputs)


_REPORT = []


@pytest.fixture
def report_line():
    """Lines a test wants in the terminal summary (they survive `-q` and land in the driver's log)."""
    return _REPORT.append


def pytest_terminal_summary(terminalreporter):
    if _REPORT:
        terminalreporter.write_sep("=", "parity / measurement summary")
        for line in _REPORT:
            terminalreporter.write_line(line)
