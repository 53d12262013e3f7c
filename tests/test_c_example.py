"""The C ABI used from plain C (examples/price_c.c, compiled with gcc against include/b200mc.h): the header is C, the
entry points take plain pointers and sizes, and without a device the program fails loudly instead of falling back."""
import math
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "monte_carlo_option_simulator_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cexample") / "price_c")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-std=c99", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "price_c.c"), "-o", out, "-L" + PKG, "-lb200mc", "-Wl,-rpath," + PKG, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


def test_c_program_compiles_and_fails_loudly_without_a_gpu(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; this test covers the GPU-less box")
    res = subprocess.run([exe, "1000"], capture_output=True, text=True)
    assert res.returncode == 2 and "no CPU fallback" in res.stderr and res.stdout == ""


@pytest.mark.gpu
def test_c_program_equals_the_python_binding(exe):
    from monte_carlo_option_simulator_b200 import SVJParams, _lib
    n = 200_000
    res = subprocess.run([exe, str(n)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lines = res.stdout.strip().splitlines()
    h = _lib.Handle(0)
    try:
        p = SVJParams()
        ks = [21000.0, 22500.0, 24000.0]
        sums = h.price_european(p, 22500.0, 0.25, 63, n, 42, ks, True, _lib.ANTITHETIC)
        prices = []
        for line, K, row in zip(lines, ks, sums):
            m = re.match(r"K=(\d+) price=([\d.]+) std_error=([\d.]+)", line)
            want = math.exp(-p.r * 0.25) * 0.5 * (row[1] + row[2]) / n
            assert float(m.group(1)) == K and float(m.group(2)) == pytest.approx(want, abs=1e-6)
            prices.append(want)
        iv = [float(x) for x in lines[3].split("=")[1].split()]
        np.testing.assert_allclose(iv, h.implied_vol(prices, 22500.0, ks, 0.25, p.r, p.q), atol=1e-6)
        S = h.simulate_terminal(p, 22500.0, 0.25, 63, n, 42, _lib.FP64, np.float64)[0]
        pnl = math.exp(-p.r * 0.25) * np.maximum(S - 22500.0, 0.0) - prices[1]
        want = h.risk_metrics(pnl, 0.99)
        got = {k: float(v) for k, v in re.findall(r"(\w+)=([-\d.]+)", lines[4])}
        assert got["var99"] == pytest.approx(want[0], abs=1e-5) and got["cvar99"] == pytest.approx(want[1], abs=1e-5)
        assert got["mean"] == pytest.approx(want[6], abs=1e-5)
    finally:
        h.close()
