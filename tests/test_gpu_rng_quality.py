"""GPU: statistical evidence for the in-register generator (Philox4x32-10 + one Box-Muller pair per 32-bit word,
csrc/philox.cuh) where the driver can see it.  The pair shares 14 bits between radius and angle, so the JOINT law of
consecutive normals is what has to be shown: chi-square on a 64 x 64 grid of equiprobable cells over 1e9 draws (same-word
pairs and adjacent-word pairs), Kolmogorov distance of the marginal, and a 10-step far-out-of-the-money price against
Black-Scholes at 1e9 paths (10 = the reference's minimum number of steps, engine/monte_carlo.py:287)."""
import math

import numpy as np
import pytest

from monte_carlo_option_simulator_b200 import SVJParams, _lib, bs_price

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    h = _lib.Handle(0)
    yield h
    h.close()


def _chi2_z(counts):
    n = counts.sum()
    e = n / counts.size
    chi2 = float(((counts - e) ** 2).sum() / e)
    dof = counts.size - 1
    return (chi2 - dof) / math.sqrt(2 * dof), n


@pytest.mark.parametrize("lag", [0, 1])
@pytest.mark.parametrize("seed", [42, 0xDEADBEEFCAFE])
def test_joint_law_of_consecutive_normals_chi_square(H, lag, seed):
    # 4e6 paths x 32 blocks x 4 pairs = 5.1e8 pairs = 1.0e9 normals (lag 0); 3.8e8 pairs for lag 1
    c = H.normal_hist2d(seed, 4_000_000, 32, lag).astype(np.float64)
    z, n = _chi2_z(c)
    assert n == 4_000_000 * 32 * (4 if lag == 0 else 3)
    assert abs(z) < 4.5, (z, n)                     # chi-square with 4095 degrees of freedom, normal approximation
    # marginals (64 equiprobable cells each): chi-square and Kolmogorov distance D sqrt(N) (P(> 1.95) = 0.001)
    for m in (c.sum(axis=1), c.sum(axis=0)):
        zm, _ = _chi2_z(m)
        assert abs(zm) < 4.5, zm
        d = np.abs(np.cumsum(m) / n - np.arange(1, 65) / 64.0).max()
        assert d * math.sqrt(n) < 1.95, d * math.sqrt(n)
    # independence beyond the marginals: the interaction chi-square of the 64 x 64 table (3969 degrees of freedom)
    e = np.outer(c.sum(axis=1), c.sum(axis=0)) / n
    zi = (float(((c - e) ** 2 / e).sum()) - 63 * 63) / math.sqrt(2 * 63 * 63)
    assert abs(zi) < 4.5, zi


def test_ten_step_far_otm_price_against_black_scholes(H):
    """1e9 paths x 10 steps, strikes 3 and 4 standard deviations out of the money: the tail of a SHORT sum of draws."""
    p = SVJParams.gbm(0.30, r=0.065, q=0.0)
    S0, T, steps, n = 2500.0, 0.04, 10, 1_000_000_000
    sd = 0.30 * math.sqrt(T)
    ks = [S0 * math.exp(3 * sd), S0 * math.exp(4 * sd), S0 * math.exp(-3 * sd)]
    disc = math.exp(-p.r * T)
    rows = H.price_european(p, S0, T, steps, n, 2024, ks, True, 0, None)
    for K, row in zip(ks, rows):
        mean, m2 = row[1] / n, row[3] / n
        price, se = disc * mean, disc * math.sqrt(max(m2 - mean * mean, 0.0) / n)
        bs = bs_price(S0, K, T, p.r, p.q, 0.30, True)
        assert abs(price - bs) < 4 * se, (K, price, bs, se)
        assert se / bs < 0.02                                   # the comparison has teeth: 2 % or better
    rows = H.price_european(p, S0, T, steps, n, 2025, [ks[2]], False, 0, None)           # far OTM put
    mean, m2 = rows[0][1] / n, rows[0][3] / n
    bs = bs_price(S0, ks[2], T, p.r, p.q, 0.30, False)
    assert abs(disc * mean - bs) < 4 * disc * math.sqrt(max(m2 - mean * mean, 0.0) / n)
