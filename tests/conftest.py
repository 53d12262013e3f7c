"""pytest configuration: markers, repo root on sys.path, golden-fixture loaders."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def garr():
    return dict(np.load(os.path.join(GOLDEN, "golden_arrays.npz")))


def unnan(d):
    """golden.json stores NaN as the string 'nan' (strict JSON)."""
    return {k: (float("nan") if v == "nan" else v) for k, v in d.items()}
