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


def cases(n: int):
    """Seeds of a randomised test: range(n), or range(n * B200MC_SOAK) for a soak run on the GPU box
    (`B200MC_SOAK=20 python -m pytest tests -m gpu -k random`; profiles/r02_soak.txt)."""
    return range(n * max(1, int(os.environ.get("B200MC_SOAK", "1"))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # The CUDA library is a build artefact (git-ignored).  A fresh checkout has none: build it once (nvcc cross-compiles
    # sm_100a without a GPU).  The tests never fall back to anything else when it is missing.
    lib = os.path.join(ROOT, "monte_carlo_option_simulator_b200", "libb200mc.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "monte_carlo_option_simulator_b200", "csrc"), "-j8", "-s"],
                           check=False)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def garr():
    return dict(np.load(os.path.join(GOLDEN, "golden_arrays.npz")))


@pytest.fixture(scope="session")
def risk_golden():
    """StressTestEngine / HedgingBacktest outputs of the reference (tests/golden/make_risk_callers_golden.py)."""
    with open(os.path.join(GOLDEN, "risk_callers_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def iv_golden():
    """Implied-volatility surface written by the reference's extract_iv_surface (tests/golden/make_iv_golden.py)."""
    return dict(np.load(os.path.join(GOLDEN, "iv_golden.npz")))


def assert_tree_close(got, want, rel=1e-9, abs_=1e-9, path=""):
    """Nested dicts / lists of floats, NaN == NaN; keys starting with '_' in `got` are ignored."""
    import math
    if isinstance(want, dict):
        assert set(k for k in got if not k.startswith("_")) == set(want), path
        for k in want:
            assert_tree_close(got[k], want[k], rel, abs_, f"{path}.{k}")
    elif isinstance(want, list):
        assert len(got) == len(want), path
        for i, (g, w) in enumerate(zip(got, want)):
            assert_tree_close(g, w, rel, abs_, f"{path}[{i}]")
    elif isinstance(want, float) and math.isnan(want):
        assert math.isnan(got), path
    else:
        assert got == pytest.approx(want, rel=rel, abs=abs_), path


def unnan(d):
    """golden.json stores NaN as the string 'nan' (strict JSON)."""
    return {k: (float("nan") if v == "nan" else v) for k, v in d.items()}
