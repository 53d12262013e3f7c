"""GPU parity tests: every CUDA entry point of libb200mc (through its C ABI / ctypes) against the CPU oracle, the
golden fixtures written by the reference, and size-independent properties at BASELINE sizes.

Tolerances (BASELINE.json north_star):
  deterministic mode (identical draws)   <= 1e-6 relative in fp64, <= 1e-4 in fp32   (we assert far tighter in fp64)
  production mode                        within 3 standard errors of the reference, within 1e-3 (relative) of
                                         closed-form Black-Scholes in the GBM limit
"""
import math

import numpy as np
import pytest

from conftest import unnan, cases
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RTOL64 = 1e-11      # fp64 kernels vs oracle on identical inputs (north_star bound: 1e-6)
RTOL32 = 1e-4       # fp32 path state vs the fp64 oracle on identical draws (north_star bound)


@pytest.fixture(scope="module")
def H():
    from monte_carlo_option_simulator_b200 import _lib
    h = _lib.Handle(0)
    yield h
    h.close()


@pytest.fixture(scope="module")
def L():
    from monte_carlo_option_simulator_b200 import _lib
    return _lib


def P(golden, name):
    return O.Params(**golden["params"][name])


PNAMES = ["svj_default", "gbm_cfg1", "heston", "jumpy"]


# ---------------------------------------------------------------------------------------------- device + RNG
def test_device_is_b200(H):
    info = H.device_info()
    assert info["cc"] // 10 == 10 and info["sm_count"] >= 100


def test_philox_words_bit_exact(H, L):
    for seed, off, stream in [(42, 0, 0), (0x1234567890ABCDEF, (1 << 32) - 3, 1), (2 ** 64 - 1, 2 ** 40 + 5, 2)]:
        got = H.dump_philox(seed, 37, 9, stream, path_offset=off)
        np.testing.assert_array_equal(got, O.philox_block_words(seed, off, 37, 9, stream))


def test_philox_kat_on_device(H, golden):
    # ctr = (path_lo, path_hi, block, stream), key = seed: reproduce the Random123 vectors through the dump
    kat = golden["cases"]["philox_kat"][0]
    assert all(int(x, 16) == 0 for x in kat["ctr"] + kat["key"])
    got = H.dump_philox(0, 1, 1, 0)[0, 0]
    np.testing.assert_array_equal(got, np.array([int(x, 16) for x in kat["out"]], dtype=np.uint32))


def test_normals_are_standard(H, L):
    z = H.dump_normals(7, 20000, 64, L.STREAM_GBM, L.Z1)
    assert abs(z.mean()) < 4 / math.sqrt(z.size) and abs(z.var() - 1) < 5 * math.sqrt(2 / z.size)
    assert abs((z ** 3).mean()) < 5 * math.sqrt(15 / z.size) and abs((z ** 4).mean() - 3) < 5 * math.sqrt(96 / z.size)
    z1 = H.dump_normals(7, 20000, 33, L.STREAM_SVJ, L.Z1)
    z2 = H.dump_normals(7, 20000, 33, L.STREAM_SVJ, L.Z2)
    u = H.dump_normals(7, 20000, 33, L.STREAM_SVJ, L.ZJUMP_U, jump_prob=0.25)     # jump times -> per-step uniforms
    zj = H.dump_normals(7, 20000, 33, L.STREAM_SVJ, L.ZJUMP_SIZE, jump_prob=0.25)
    assert abs(np.corrcoef(z1.ravel(), z2.ravel())[0, 1]) < 5 / math.sqrt(z1.size)
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 5 / math.sqrt(12 * u.size)
    zf = zj[u < 0.25]                    # jump sizes exist where the jump fires (U < jump_prob) and are N(0,1) there
    assert abs(zf.size / u.size - 0.25) < 0.01 and np.all(zj[u >= 0.25] == 0)
    assert abs(zf.mean()) < 5 / math.sqrt(zf.size) and abs(zf.var() - 1) < 5 * math.sqrt(2 / zf.size)
    assert abs(np.corrcoef(zf, z1[u < 0.25])[0, 1]) < 5 / math.sqrt(zf.size)


# ---------------------------------------------------------------------------------------------- a1: given normals
@pytest.mark.parametrize("name", PNAMES)
def test_given_normals_golden(H, golden, garr, name):
    p = P(golden, name)
    S, v, paths = H.simulate_given_normals(p, 22500.0, 0.25, garr["k_small_Z1"], garr["k_small_Z2"], garr["k_small_Zj"],
                                           garr["k_small_Zjs"], 40, True)
    np.testing.assert_allclose(S, garr[f"k_small_{name}_S"], rtol=RTOL64)
    np.testing.assert_allclose(v, garr[f"k_small_{name}_v"], rtol=RTOL64, atol=1e-17)
    np.testing.assert_allclose(paths, garr[f"k_small_{name}_paths"], rtol=RTOL64)
    assert np.all(paths[:, 0] == 22500.0)


def test_given_normals_pcg64_4096(H, golden, garr):
    c = golden["cases"]["k_4096"]
    p = P(golden, c["params"])
    Z1, Z2, Zj, Zjs = O.draw_pcg64(c["seed"], c["n"], c["steps"])
    S, v, none = H.simulate_given_normals(p, c["spot"], c["T"], Z1, Z2, Zj, Zjs, c["steps"])
    assert none is None
    np.testing.assert_allclose(S, garr["k_4096_S"], rtol=RTOL64)
    np.testing.assert_allclose(v, garr["k_4096_v"], rtol=RTOL64, atol=1e-17)
    Sa = H.simulate_given_normals(p, c["spot"], c["T"], -Z1, -Z2, Zj, -Zjs, c["steps"])[0]
    np.testing.assert_allclose(Sa, garr["k_4096_S_anti"], rtol=RTOL64)
    Su = H.simulate_given_normals(p.replace(v0=p.v0 + 0.01), c["spot"], c["T"], Z1, Z2, Zj, Zjs, c["steps"])[0]
    np.testing.assert_allclose(Su, garr["k_4096_S_v0up"], rtol=RTOL64)


@pytest.mark.parametrize("n,steps", [(1, 1), (31, 7), (33, 16), (100, 17), (257, 50)])
def test_given_normals_ragged(H, golden, n, steps):
    p = P(golden, "svj_default")
    g = np.random.default_rng(n * 1000 + steps)
    Z1, Z2, Zjs = (g.standard_normal((n, steps)) for _ in range(3))
    Zj = g.random((n, steps))
    S, v, paths = H.simulate_given_normals(p, 100.0, 0.5, Z1, Z2, Zj, Zjs, steps, True)
    So, vo, po = O.simulate_svj(100.0, p.v0, p.r, p.q, 0.5, p.kappa, p.theta, p.xi, p.rho, p.lambda_j, p.mu_j, p.sigma_j,
                                Z1, Z2, Zj, Zjs, steps, True)
    np.testing.assert_allclose(S, So, rtol=RTOL64)
    np.testing.assert_allclose(v, vo, rtol=RTOL64, atol=1e-17)
    np.testing.assert_allclose(paths, po, rtol=RTOL64)


@pytest.mark.parametrize("mode", ["gbm", "heston", "svj"])
def test_given_normals_tile_walks_agree_bitwise(H, mode, monkeypatch):
    """The kernel walks a full tile either step by step or in three sweeps (recurrence + exponents, the tile's
    exponentials side by side, running product), with 8- or 16-step tiles: the same separately rounded operations per
    path, so all four variants must agree bit for bit -- and with the oracle to rounding -- on ragged shapes too."""
    from monte_carlo_option_simulator_b200 import SVJParams
    p = {"gbm": SVJParams.gbm(0.3), "heston": SVJParams(lambda_j=0.0), "svj": SVJParams()}[mode]
    for n, steps in ((1, 1), (33, 7), (64, 8), (257, 41), (1000, 250)):
        g = np.random.default_rng(n + steps)
        Z1, Z2, Zjs = (g.standard_normal((n, steps)) for _ in range(3))
        Zj = g.random((n, steps))
        out = []
        for ilp in ("0", "1"):
            for tile in ("8", "16"):
                monkeypatch.setenv("B200MC_GN_ILP", ilp)
                monkeypatch.setenv("B200MC_GN_TILE", tile)
                out.append(H.simulate_given_normals(p, 2500.0, 1.0, Z1, Z2, Zj, Zjs, steps, True))
        monkeypatch.delenv("B200MC_GN_ILP")
        monkeypatch.delenv("B200MC_GN_TILE")
        for o in out[1:]:
            for a, b in zip(o, out[0]):
                np.testing.assert_array_equal(a, b)
        So, vo, po = O.simulate_svj(2500.0, p.v0, p.r, p.q, 1.0, p.kappa, p.theta, p.xi, p.rho, p.lambda_j, p.mu_j, p.sigma_j,
                                    Z1, Z2, Zj, Zjs, steps, True)
        np.testing.assert_allclose(out[0][0], So, rtol=RTOL64)
        np.testing.assert_allclose(out[0][2], po, rtol=RTOL64)


def test_given_normals_empty_and_errors(H, L, golden):
    p = P(golden, "svj_default")
    e = np.zeros((0, 5))
    S, v, _ = H.simulate_given_normals(p, 100.0, 0.5, e, e, e, e, 5)
    assert S.shape == (0,) and v.shape == (0,)
    with pytest.raises(L.B200MCError):
        H.simulate_given_normals(p, 100.0, 0.5, np.zeros((4, 5)), np.zeros((3, 5)), np.zeros((4, 5)), np.zeros((4, 5)), 5)
    with pytest.raises(L.B200MCError):
        H.price_european(p, 100.0, 0.5, 0, 10, 1, [100.0])
    with pytest.raises(L.B200MCError):
        H.price_european(p, 100.0, -1.0, 10, 10, 1, [100.0])


def test_drop_in_kernel_function(golden, garr):
    from monte_carlo_option_simulator_b200 import _simulate_svj_paths_numba as f
    p = P(golden, "jumpy")
    S, v, paths = f(22500.0, p.v0, p.r, p.q, 0.25, p.kappa, p.theta, p.xi, p.rho, p.lambda_j, p.mu_j, p.sigma_j,
                    garr["k_small_Z1"], garr["k_small_Z2"], garr["k_small_Zj"], garr["k_small_Zjs"], 40)
    assert paths.shape == (0, 0)
    np.testing.assert_allclose(S, garr["k_small_jumpy_S"], rtol=RTOL64)


# ---------------------------------------------------------------------------------------------- fused: deterministic mode
def _mode_params(golden, mode):
    if mode == "gbm":
        return P(golden, "gbm_cfg1").replace(kappa=0.0), 0
    if mode == "detvar":
        return P(golden, "gbm_cfg1").replace(kappa=3.0, theta=0.05), 0
    if mode == "heston":
        return P(golden, "heston"), 1
    return P(golden, "svj_default"), 2


def _draws(H, L, seed, n, steps, stream, off=0, p=None, T=None):
    """The four arrays the reference kernel consumes, exactly as the fused kernels draw them.  jump_prob (needed for
    the jump sizes of the SVJ stream) is lambda_j * (T / steps), computed in the library's own order."""
    jp = p.lambda_j * (T / steps) if p is not None else 0.0
    return [H.dump_normals(seed, n, steps, stream, w, path_offset=off, jump_prob=jp)
            for w in (L.Z1, L.Z2, L.ZJUMP_U, L.ZJUMP_SIZE)]


@pytest.mark.parametrize("mode", ["gbm", "detvar", "heston", "svj"])
@pytest.mark.parametrize("steps", [250, 63, 10, 7])
def test_fused_terminal_identical_draws(H, L, golden, mode, steps):
    """north_star deterministic mode: the reference recurrence (oracle) fed the IDENTICAL draws."""
    p, stream = _mode_params(golden, mode)
    n, seed, off, T, S0 = 1500, 99, 12345, 0.8, 2500.0
    Z1, Z2, Zj, Zjs = _draws(H, L, seed, n, steps, stream, off, p, T)
    So, vo, _ = O._sim(p, S0, T, Z1, Z2, Zj, Zjs, steps)
    Sao = O._sim(p, S0, T, -Z1, -Z2, Zj, -Zjs, steps)[0]
    S, A, V = H.simulate_terminal(p, S0, T, steps, n, seed, L.ANTITHETIC | L.FP64, np.float64, off, True, True)
    np.testing.assert_allclose(S, So, rtol=1e-10)
    np.testing.assert_allclose(A, Sao, rtol=1e-10)
    np.testing.assert_allclose(V, vo, rtol=1e-9, atol=1e-15)
    S32, A32, V32 = H.simulate_terminal(p, S0, T, steps, n, seed, L.ANTITHETIC, np.float32, off, True, True)
    assert S32.dtype == np.float32
    np.testing.assert_allclose(S32, So, rtol=RTOL32)
    np.testing.assert_allclose(A32, Sao, rtol=RTOL32)
    np.testing.assert_allclose(V32, vo, rtol=2e-3, atol=2e-6)   # variance near the zero boundary amplifies fp32 rounding


@pytest.mark.parametrize("case", cases(24))
def test_fused_random_configurations(H, L, case):
    """Randomised deterministic-mode parity: random SVJ parameters (every mode), shapes, offsets (also across the
    2^32 path-index carry), flags; terminal values, variance, sums and the path matrix against the oracle fed the
    identical draws (fp64 state)."""
    g = np.random.default_rng(1000 + case)
    xi = 0.0 if g.random() < 0.4 else float(g.uniform(0.05, 1.2))
    lam = 0.0 if g.random() < 0.5 else float(g.uniform(0.2, 8.0))
    kappa = 0.0 if g.random() < 0.25 else float(g.uniform(0.2, 6.0))
    v0 = float(g.uniform(0.005, 0.3))
    theta = v0 if g.random() < 0.3 else float(g.uniform(0.005, 0.3))
    p = O.Params(kappa=kappa, theta=theta, xi=xi, rho=float(g.uniform(-0.95, 0.95)), v0=v0, lambda_j=lam,
                 mu_j=float(g.uniform(-0.2, 0.1)), sigma_j=float(g.uniform(0.01, 0.3)), r=float(g.uniform(0.0, 0.1)),
                 q=float(g.uniform(0.0, 0.05)))
    n, steps = int(g.integers(1, 700)), int(g.integers(1, 300))
    T = float(g.uniform(0.02, 3.0))
    S0 = float(g.uniform(1.0, 30000.0))
    seed = int(g.integers(0, 2 ** 63))
    off = int(g.choice([0, 17, 2 ** 32 - n // 2 - 1, 2 ** 40 + 3]))
    anti = bool(g.integers(0, 2))
    fl = L.FP64 | (L.ANTITHETIC if anti else 0)
    stream = L.select_stream(p, T, steps, fl)
    Z1, Z2, Zj, Zjs = _draws(H, L, seed, n, steps, stream, off, p, T)
    So, vo, po = O._sim(p, S0, T, Z1, Z2, Zj, Zjs, steps, record=True)
    S, A, V = H.simulate_terminal(p, S0, T, steps, n, seed, fl, np.float64, off, anti, True)
    np.testing.assert_allclose(S, So, rtol=2e-9)
    np.testing.assert_allclose(V, vo, rtol=1e-8, atol=1e-14)
    if anti:
        np.testing.assert_allclose(A, O._sim(p, S0, T, -Z1, -Z2, Zj, -Zjs, steps)[0], rtol=2e-9)
    np.testing.assert_allclose(H.generate_paths(p, S0, T, steps, n, seed, L.FP64, np.float64, off), po, rtol=2e-9)
    ks = sorted(float(S0 * x) for x in g.uniform(0.6, 1.4, size=int(g.choice([1, 2, 7]))))
    is_call = bool(g.integers(0, 2))
    rows = H.price_european(p, S0, T, steps, n, seed, ks, is_call, fl, None, path_offset=off)
    for K, row in zip(ks, rows):
        pay = np.maximum(S - K, 0.0) if is_call else np.maximum(K - S, 0.0)
        assert row[0] == n and row[1] == pytest.approx(pay.sum(), rel=1e-11, abs=1e-9)
        assert row[3] == pytest.approx((pay * pay).sum(), rel=1e-11, abs=1e-9)
    # fp32 state on the same draws: inside the north star's 1e-4 unless the variance sits on the zero boundary
    S32 = H.simulate_terminal(p, S0, T, steps, n, seed, 0, np.float32, off)[0]
    assert np.median(np.abs(S32 / So - 1)) < 2e-6


@pytest.mark.parametrize("name", ["gbm", "detvar", "heston", "svj", "jumpy"])
def test_fused_modes_against_the_reference_itself(H, L, name):
    """The north star's deterministic mode, literally: the REFERENCE kernel (run in the dev container, fixtures in
    tests/golden/fused_golden.npz) was fed the draws this library dumps; the fused kernels must (a) still produce those
    draws bit for bit and (b) reproduce the reference's outputs: <= 1e-6 relative with the fp64 state (observed ~1e-14),
    <= 1e-4 with the fp32 state."""
    import json
    import os
    from conftest import GOLDEN
    from monte_carlo_option_simulator_b200 import SVJParams
    c = json.load(open(os.path.join(GOLDEN, "fused_golden_cases.json")))[name]
    d = dict(np.load(os.path.join(GOLDEN, "fused_golden.npz")))
    p = SVJParams(**c["params"])
    n, steps, T, S0, seed, off = c["n"], c["steps"], c["T"], c["S0"], c["seed"], c["path_offset"]
    stream = L.select_stream(p, T, steps)
    assert stream == int(d[f"{name}_stream"][0])
    for i, w in enumerate(("Z1", "Z2", "Zj", "Zjs")):
        got = H.dump_normals(seed, n, steps, stream, i, path_offset=off, jump_prob=p.lambda_j * (T / steps))
        np.testing.assert_array_equal(got, d[f"{name}_{w}"])
    S, A, V = H.simulate_terminal(p, S0, T, steps, n, seed, L.FP64 | L.ANTITHETIC, np.float64, off, True, True)
    np.testing.assert_allclose(S, d[f"{name}_ref_S"], rtol=1e-6)
    np.testing.assert_allclose(S, d[f"{name}_ref_S"], rtol=1e-10)            # what we actually achieve
    np.testing.assert_allclose(A, d[f"{name}_ref_S_anti"], rtol=1e-10)
    np.testing.assert_allclose(V, d[f"{name}_ref_v"], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(H.generate_paths(p, S0, T, steps, n, seed, L.FP64, np.float64, off), d[f"{name}_ref_paths"],
                               rtol=1e-10)
    S32, A32, _ = H.simulate_terminal(p, S0, T, steps, n, seed, L.ANTITHETIC, np.float32, off, True, False)
    np.testing.assert_allclose(S32, d[f"{name}_ref_S"], rtol=1e-4)
    np.testing.assert_allclose(A32, d[f"{name}_ref_S_anti"], rtol=1e-4)
    np.testing.assert_allclose(H.generate_paths(p, S0, T, steps, n, seed, 0, np.float32, off), d[f"{name}_ref_paths"], rtol=1e-4)
    bumps = L.Bumps(0.01, p.v0 + 0.01, max(p.v0 - 0.01, 0.001), p.r + 1e-4, max(p.r - 1e-4, 0))
    row = H.price_european(p, S0, T, steps, n, seed, [S0], True, L.FP64 | L.GREEKS, bumps, path_offset=off)[0]
    assert row[L.SUMS_FIELDS.index("sum_v0_up")] == pytest.approx(np.maximum(d[f"{name}_ref_S_v0up"] - S0, 0).sum(), rel=1e-9)
    assert row[L.SUMS_FIELDS.index("sum_a")] == pytest.approx(np.maximum(d[f"{name}_ref_S"] - S0, 0).sum(), rel=1e-9)


def test_fused_jumps_fire_like_reference(H, L, golden):
    """A jump-heavy parameter set: the integer jump test in the kernel must equal the reference's float compare."""
    p = P(golden, "jumpy")
    n, steps, seed = 4000, 40, 5
    Z1, Z2, Zj, Zjs = _draws(H, L, seed, n, steps, 2, 0, p, 0.5)
    fired = Zj < p.lambda_j * (0.5 / steps)
    assert fired.sum() > 100
    # jump sizes are standard normal where the jump fires, and are never consulted elsewhere
    zs = Zjs[fired]
    assert abs(zs.mean()) < 4 / math.sqrt(zs.size) and abs(zs.var() - 1) < 5 * math.sqrt(2 / zs.size)
    assert np.all(Zjs[~fired] == 0.0)
    So = O._sim(p, 100.0, 0.5, Z1, Z2, Zj, Zjs, steps)[0]
    S = H.simulate_terminal(p, 100.0, 0.5, steps, n, seed, L.FP64, np.float64)[0]
    np.testing.assert_allclose(S, So, rtol=1e-10)


def test_very_long_deterministic_variance_run_uses_the_heston_stream(H, L, golden):
    """xi = 0 with a moving variance and more steps than the weight table holds: the library falls back to the general
    stochastic-variance kernel (Heston stream); b200mc_select_stream tells the caller which draws to dump."""
    p, _ = _mode_params(golden, "detvar")
    n, steps, T = 64, 4100, 4.0
    stream = L.select_stream(p, T, steps)
    assert stream == L.STREAM_HESTON and L.select_stream(p, T, 4096) == L.STREAM_GBM
    Z = _draws(H, L, 6, n, steps, stream, 0, p, T)
    want = O._sim(p, 100.0, T, *Z, steps)[0]
    got = H.simulate_terminal(p, 100.0, T, steps, n, 6, L.FP64, np.float64)[0]
    np.testing.assert_allclose(got, want, rtol=1e-9)


def test_force_svj_on_gbm_params(H, L, golden):
    p, _ = _mode_params(golden, "gbm")
    n, steps, seed = 512, 30, 3
    Z1, Z2, Zj, Zjs = _draws(H, L, seed, n, steps, 2, 0, p, 1.0)
    So = O._sim(p, 2500.0, 1.0, Z1, Z2, Zj, Zjs, steps)[0]
    S = H.simulate_terminal(p, 2500.0, 1.0, steps, n, seed, L.FP64 | L.FORCE_SVJ, np.float64)[0]
    np.testing.assert_allclose(S, So, rtol=1e-10)


# ---------------------------------------------------------------------------------------------- fused: sums
def _np_sums(L, S, A, K, is_call, S0, extra=None):
    pay = (lambda s: np.maximum(s - K, 0.0)) if is_call else (lambda s: np.maximum(K - s, 0.0))
    a = pay(S)
    b = pay(A) if A is not None else np.zeros_like(a)
    s_avg = 0.5 * (S + A) if A is not None else S
    payc = 0.5 * (a + b) if A is not None else a
    d = dict(n=len(S), sum_a=a.sum(), sum_b=b.sum(), sum_aa=(a * a).sum(), sum_bb=(b * b).sum(), sum_ab=(a * b).sum(),
             sum_s=s_avg.sum(), sum_ss=(s_avg ** 2).sum(), sum_ps=(payc * s_avg).sum())
    return d


@pytest.mark.parametrize("mode", ["gbm", "detvar", "heston", "svj"])
@pytest.mark.parametrize("anti", [False, True])
@pytest.mark.parametrize("is_call", [True, False])
def test_fused_sums_match_terminal_values(H, L, golden, mode, anti, is_call):
    """The on-chip reduction equals NumPy sums over the terminal values of the same paths (fp64 state), for a
    ragged path count, a path offset and several strike counts (1, 3, 21, 64)."""
    p, _ = _mode_params(golden, mode)
    n, steps, seed, off, S0, T = 3001, 50, 11, 777, 2500.0, 1.0
    fl = L.FP64 | (L.ANTITHETIC if anti else 0)
    S, A, _ = H.simulate_terminal(p, S0, T, steps, n, seed, fl, np.float64, off, anti, False)
    for ks in ([2500.0], [2000.0, 2500.0, 3100.0], list(np.linspace(0.7, 1.3, 21) * S0), list(np.linspace(0.7, 1.3, 64) * S0)):
        rows = H.price_european(p, S0, T, steps, n, seed, ks, is_call, fl, None, path_offset=off)
        assert rows.shape == (len(ks), L.NSUMS)
        for K, row in zip(ks, rows):
            want = _np_sums(L, S, A if anti else None, K, is_call, S0)
            for k, w in want.items():
                assert row[L.SUMS_FIELDS.index(k)] == pytest.approx(w, rel=1e-11, abs=1e-6), (k, K)


@pytest.mark.parametrize("mode", ["gbm", "detvar", "heston", "svj"])
@pytest.mark.parametrize("is_call", [True, False])
def test_fused_greek_sums_vs_oracle(H, L, golden, mode, is_call):
    """Bump accumulators of the fused launch == re-simulating with the oracle on the identical draws (the
    reference's CRN construction, greeks.py:65-80,:124-147)."""
    p, stream = _mode_params(golden, mode)
    n, steps, seed, S0, T, K, b = 2000, 40, 21, 2500.0, 0.5, 2450.0, 0.01
    Z = _draws(H, L, seed, n, steps, stream, 0, p, T)
    bumps = L.Bumps(b, p.v0 + 0.01, max(p.v0 - 0.01, 0.001), p.r + 1e-4, max(p.r - 1e-4, 0))
    row = H.price_european(p, S0, T, steps, n, seed, [K], is_call, L.FP64 | L.GREEKS, bumps)[0]
    col = {k: row[i] for i, k in enumerate(L.SUMS_FIELDS)}
    pay = (lambda s: np.maximum(s - K, 0.0)) if is_call else (lambda s: np.maximum(K - s, 0.0))
    S = O._sim(p, S0, T, *Z, steps)[0]
    itm = (S > K) if is_call else (S < K)
    want = {
        "sum_a": pay(S).sum(),
        "sum_pw_delta": (itm * S / S0).sum(),
        "sum_spot_up": pay(O._sim(p, S0 * (1 + b), T, *Z, steps)[0]).sum(),
        "sum_spot_dn": pay(O._sim(p, S0 * (1 - b), T, *Z, steps)[0]).sum(),
        "sum_v0_up": pay(O._sim(p, S0, T, *Z, steps, v0=bumps.v0_up)[0]).sum(),
        "sum_v0_dn": pay(O._sim(p, S0, T, *Z, steps, v0=bumps.v0_dn)[0]).sum(),
        "sum_r_up": pay(O._sim(p.replace(r=bumps.r_up), S0, T, *Z, steps)[0]).sum(),
        "sum_r_dn": pay(O._sim(p.replace(r=bumps.r_dn), S0, T, *Z, steps)[0]).sum(),
    }
    for k, w in want.items():
        assert col[k] == pytest.approx(w, rel=1e-9), k


@pytest.mark.parametrize("mode", ["gbm", "svj"])
def test_multi_strike_greek_sums_equal_single_strike_launches(H, L, golden, mode):
    """The strike-major phase B with bump accumulators (multi-strike + GREEKS) against one launch per strike."""
    p, _ = _mode_params(golden, mode)
    ks = [2300.0, 2500.0, 2750.0]
    bumps = L.Bumps(0.01, p.v0 + 0.01, max(p.v0 - 0.01, 0.001), p.r + 1e-4, max(p.r - 1e-4, 0))
    for fl, tol in ((L.GREEKS | L.FP64, 1e-12), (L.GREEKS | L.FP64 | L.ANTITHETIC, 1e-12), (L.GREEKS, 2e-6)):
        many = H.price_european(p, 2500.0, 0.5, 40, 5000, 3, ks, False, fl, bumps)
        for K, row in zip(ks, many):
            one = H.price_european(p, 2500.0, 0.5, 40, 5000, 3, [K], False, fl, bumps)[0]
            np.testing.assert_allclose(row, one, rtol=tol, atol=1e-9)


@pytest.mark.parametrize("mode", ["gbm", "detvar", "heston", "svj"])
def test_edge_shapes(H, L, golden, mode):
    """One path, one step, 256 strikes, very many steps, tiny maturities: shapes the reference's callers can produce
    (steps = max(int(252 T), 10), verify.py uses T = 0.04)."""
    p, stream = _mode_params(golden, mode)
    for n, steps, T in [(1, 1, 0.004), (3, 10, 0.04), (257, 9, 0.3), (40, 1200, 3.0)]:
        Z = _draws(H, L, 17, n, steps, stream, 5, p, T)
        want = O._sim(p, 100.0, T, *Z, steps)[0]
        got = H.simulate_terminal(p, 100.0, T, steps, n, 17, L.FP64, np.float64, 5)[0]
        np.testing.assert_allclose(got, want, rtol=1e-9)
        ks = list(np.linspace(60.0, 140.0, 256))
        rows = H.price_european(p, 100.0, T, steps, n, 17, ks, True, L.FP64, None, path_offset=5)
        np.testing.assert_allclose(rows[:, 1], [np.maximum(want - K, 0).sum() for K in ks], rtol=1e-9, atol=1e-9)
        paths = H.generate_paths(p, 100.0, T, steps, n, 17, L.FP64, np.float64, 5)
        np.testing.assert_allclose(paths[:, -1], want, rtol=1e-9)
    with pytest.raises(L.B200MCError):
        H.price_european(p, 100.0, 1.0, 10, 10, 1, list(np.linspace(60.0, 140.0, 257)))


def test_strike_count_sequence_regression(H, L, golden):
    """Regression (found by the randomised test): large, small, large strike counts on the same kernel -- the cached
    launch configuration must not leave the dynamic shared-memory cap at the smaller size."""
    p, _ = _mode_params(golden, "svj")
    for nk in (64, 2, 64, 7, 256, 3, 256):
        ks = list(np.linspace(80.0, 120.0, nk))
        rows = H.price_european(p, 100.0, 0.5, 16, 700, 1, ks, True, L.FP64)
        assert rows.shape == (nk, L.NSUMS) and np.all(rows[:, 0] == 700)


@pytest.mark.parametrize("case", cases(12))
def test_greek_sums_random_configurations(H, L, case):
    """Randomised parameters / bumps / strike counts: every bump accumulator of the fused launch against oracle
    re-simulation on the identical draws (the reference's CRN construction), antithetic on or off."""
    g = np.random.default_rng(2000 + case)
    xi = 0.0 if g.random() < 0.4 else float(g.uniform(0.05, 1.0))
    lam = 0.0 if g.random() < 0.5 else float(g.uniform(0.2, 6.0))
    kappa = 0.0 if g.random() < 0.3 else float(g.uniform(0.2, 5.0))
    v0 = float(g.uniform(0.02, 0.25))
    p = O.Params(kappa=kappa, theta=v0 if g.random() < 0.3 else float(g.uniform(0.02, 0.25)), xi=xi,
                 rho=float(g.uniform(-0.9, 0.9)), v0=v0, lambda_j=lam, mu_j=float(g.uniform(-0.15, 0.05)),
                 sigma_j=float(g.uniform(0.02, 0.25)), r=float(g.uniform(0.0, 0.08)), q=float(g.uniform(0.0, 0.04)))
    n, steps, T, S0 = int(g.integers(50, 900)), int(g.integers(3, 120)), float(g.uniform(0.05, 2.0)), float(g.uniform(50, 5000))
    b = float(g.choice([0.01, 0.02, 0.005]))
    bumps = L.Bumps(b, p.v0 + 0.01, max(p.v0 - 0.01, 0.001), p.r + 1e-4, max(p.r - 1e-4, 0))
    anti = bool(g.integers(0, 2))
    fl = L.FP64 | L.GREEKS | (L.ANTITHETIC if anti else 0)
    is_call = bool(g.integers(0, 2))
    ks = sorted(float(S0 * x) for x in g.uniform(0.7, 1.3, size=int(g.choice([1, 3]))))
    seed = int(g.integers(0, 2 ** 40))
    stream = L.select_stream(p, T, steps, fl, bumps)
    Z = _draws(H, L, seed, n, steps, stream, 0, p, T)
    rows = H.price_european(p, S0, T, steps, n, seed, ks, is_call, fl, bumps)
    S = O._sim(p, S0, T, *Z, steps)[0]
    for K, row in zip(ks, rows):
        col = {k: row[i] for i, k in enumerate(L.SUMS_FIELDS)}
        pay = (lambda s: np.maximum(s - K, 0.0)) if is_call else (lambda s: np.maximum(K - s, 0.0))
        itm = (S > K) if is_call else (S < K)
        want = {"sum_a": pay(S).sum(), "sum_pw_delta": (itm * S / S0).sum(),
                "sum_spot_up": pay(O._sim(p, S0 * (1 + b), T, *Z, steps)[0]).sum(),
                "sum_spot_dn": pay(O._sim(p, S0 * (1 - b), T, *Z, steps)[0]).sum(),
                "sum_v0_up": pay(O._sim(p, S0, T, *Z, steps, v0=bumps.v0_up)[0]).sum(),
                "sum_v0_dn": pay(O._sim(p, S0, T, *Z, steps, v0=bumps.v0_dn)[0]).sum(),
                "sum_r_up": pay(O._sim(p.replace(r=bumps.r_up), S0, T, *Z, steps)[0]).sum(),
                "sum_r_dn": pay(O._sim(p.replace(r=bumps.r_dn), S0, T, *Z, steps)[0]).sum()}
        if anti:
            want["sum_b"] = pay(O._sim(p, S0, T, -Z[0], -Z[1], Z[2], -Z[3], steps)[0]).sum()
        for k, w in want.items():
            assert col[k] == pytest.approx(w, rel=2e-9, abs=1e-7), (k, K)


def test_fused_fp32_sums_close_to_fp64(H, L, golden):
    p, _ = _mode_params(golden, "gbm")
    ks = list(np.linspace(0.8, 1.2, 5) * 2500.0)
    r64 = H.price_european(p, 2500.0, 1.0, 250, 200_000, 42, ks, True, L.FP64 | L.ANTITHETIC)
    r32 = H.price_european(p, 2500.0, 1.0, 250, 200_000, 42, ks, True, L.ANTITHETIC)
    np.testing.assert_allclose(r32[:, :9], r64[:, :9], rtol=2e-5)


def test_fused_launch_is_reproducible_and_offsets_add(H, L, golden):
    p, _ = _mode_params(golden, "svj")
    for ks in ([95.0, 105.0], [100.0]):
        a = H.price_european(p, 100.0, 0.5, 20, 10_000, 9, ks, True, L.ANTITHETIC)
        b = H.price_european(p, 100.0, 0.5, 20, 10_000, 9, ks, True, L.ANTITHETIC)
        np.testing.assert_array_equal(a, b)                     # same geometry => bitwise identical
        for fl, tol in ((L.ANTITHETIC | L.FP64, 1e-12), (L.ANTITHETIC, 2e-6)):
            whole = H.price_european(p, 100.0, 0.5, 20, 10_000, 9, ks, True, fl)
            lo = H.price_european(p, 100.0, 0.5, 20, 3_333, 9, ks, True, fl)
            hi = H.price_european(p, 100.0, 0.5, 20, 6_667, 9, ks, True, fl, path_offset=3_333)
            # disjoint path ranges add up (what multi-GPU sharding relies on): exactly in fp64; in fp32 the
            # multi-strike kernel folds per-batch fp32 partials, so the sum depends on the batching at the 1e-7 level
            np.testing.assert_allclose(lo + hi, whole, rtol=tol)


# ---------------------------------------------------------------------------------------------- production mode
def test_production_gbm_vs_black_scholes_and_reference(golden):
    """BASELINE config 1 market, 2e7 paths x 250 steps, plain MC (no antithetic, no CV): |MC - BS| <= 1e-3 * BS
    (SE ~ 575 / sqrt(n) = 0.13 => 1e-3 relative is ~2.9 sigma; we also assert the 3-sigma band itself) and within
    3 combined standard errors of the reference's own 50k-path value (golden)."""
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    p = SVJParams(**golden["params"]["gbm_cfg1"])
    ref = [c for c in golden["cases"]["price"] if c["n"] == 50_000 and c["is_call"] and not c["sobol"]
           and not c["anti"] and not c["cv"]][0]["result"]
    bs = golden["cases"]["bs"]["cfg1_call"]
    for prec in ("fp32", "fp64"):
        eng = MonteCarloEngine(p, 20_000_000, 250, seed=42, use_sobol=False, use_antithetic=False,
                               use_control_variate=False, rng="philox", precision=prec)
        r = eng.price(2500.0, 2500.0, 1.0, True)
        assert r["num_steps"] == 250 and r["num_paths_used"] == 20_000_000
        assert abs(r["price"] - bs) <= 3 * r["std_error"]
        assert abs(r["price"] - bs) <= 1e-3 * bs
        assert abs(r["price"] - ref["price"]) <= 3 * math.hypot(r["std_error"], ref["std_error"])
        assert r["std_error"] == pytest.approx(ref["std_error"] * math.sqrt(50_000 / 20_000_000), rel=0.02)
        # the genuine control variate (new key) is far tighter and still unbiased
        assert abs(r["price_cv_spot"] - bs) <= 4 * r["std_error_cv_spot"]
        assert r["std_error_cv_spot"] < 0.5 * r["std_error"]


def test_production_put_call_parity_and_antithetic(golden):
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    p = SVJParams(**golden["params"]["gbm_cfg1"])
    eng = MonteCarloEngine(p, 4_000_000, 250, seed=7, use_sobol=False, use_antithetic=True, use_control_variate=False,
                           rng="philox")
    c = eng.price(2500.0, 2400.0, 1.0, True)
    q = eng.price(2500.0, 2400.0, 1.0, False)
    fwd = 2500.0 * math.exp(-p.q) - 2400.0 * math.exp(-p.r)
    assert abs((c["price"] - q["price"]) - fwd) <= 4 * math.hypot(c["std_error"], q["std_error"])
    assert abs(c["price"] - O.bs_price(2500.0, 2400.0, 1.0, p.r, p.q, 0.3, True)) <= 4 * c["std_error"]


@pytest.mark.parametrize("case", [0, 1, 2])
def test_production_svj_within_3se_of_reference(golden, case):
    """Default SVJ parameters (stochastic vol + jumps): fused Philox engine vs the reference's golden prices at
    4096 paths (pseudo-random, no CV), using many more paths on our side."""
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    cases = [c for c in golden["cases"]["price"] if c["params"] == "svj_default" and not c["sobol"] and not c["cv"]]
    c = cases[case % len(cases)]
    p = SVJParams(**golden["params"][c["params"]])
    eng = MonteCarloEngine(p, 2_000_000, c["num_steps"], seed=1, use_sobol=False, use_antithetic=c["anti"],
                           use_control_variate=False, rng="philox")
    r = eng.price(c["spot"], c["strike"], c["T"], c["is_call"])
    assert r["num_steps"] == c["result"]["num_steps"]
    assert abs(r["price"] - c["result"]["price"]) <= 3 * math.hypot(r["std_error"], c["result"]["std_error"])


@pytest.mark.parametrize("pname", ["svj_default", "heston", "jumpy"])
def test_production_sv_modes_against_a_large_oracle_sample(golden, pname):
    """Production mode for the stochastic-variance / jump kernels: 4e6 Philox paths on the GPU against 2e5 PCG64 paths
    of the oracle (the reference's own estimator, antithetic, no CV) -- a 7x tighter band than the 4096-path goldens."""
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    p = SVJParams(**golden["params"][pname])
    spot, T = 22500.0, 0.25
    ours = MonteCarloEngine(p, 4_000_000, 252, seed=5, use_sobol=False, use_antithetic=True, use_control_variate=False,
                            rng="philox")
    ref = O.MonteCarloOracle(O.Params(**golden["params"][pname]), 200_000, 252, 42, False, True, False)
    for K, is_call in ((21500.0, False), (22500.0, True), (23500.0, True)):
        a = ours.price(spot, K, T, is_call)
        b = ref.price(spot, K, T, is_call)
        assert abs(a["price"] - b["price"]) <= 3.5 * math.hypot(a["std_error"], b["std_error"]), (pname, K, a, b)
        assert a["std_error"] == pytest.approx(b["std_error"] * math.sqrt(200_000 / 4_000_000), rel=0.05)


def test_pseudo_cv_formula_reproduced(golden):
    """Quirk 2: with antithetic off the reference's 'control variate' returns exactly bs_ref with SE 0; with it
    on, price = bs_ref + D*mean((b - a)/2).  Same dictionary keys as the reference."""
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    p = SVJParams(**golden["params"]["gbm_cfg1"])
    want = [c for c in golden["cases"]["price"] if c["n"] == 50_000 and c["is_call"] and not c["sobol"] and c["cv"]]
    for c in want:
        eng = MonteCarloEngine(p, 50_000, 250, seed=42, use_sobol=False, use_antithetic=c["anti"],
                               use_control_variate=True, rng="philox")
        r = eng.price(2500.0, 2500.0, 1.0, True)
        assert set(c["result"]) <= set(r)
        assert r["bs_ref"] == pytest.approx(c["result"]["bs_ref"], rel=1e-13)
        if not c["anti"]:
            assert r["price"] == pytest.approx(r["bs_ref"], rel=1e-12) and r["std_error"] < 1e-9
        else:
            assert abs(r["price"] - c["result"]["price"]) <= 3 * math.hypot(r["std_error"], c["result"]["std_error"])
            assert r["std_error"] == pytest.approx(c["result"]["std_error"], rel=0.05)


# ---------------------------------------------------------------------------------------------- rng="reference" (bit parity)
def test_reference_rng_matches_golden_prices(golden):
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    cases = [c for c in golden["cases"]["price"] if c["n"] <= 5000]
    assert len(cases) >= 24
    for c in cases:
        eng = MonteCarloEngine(SVJParams(**golden["params"][c["params"]]), c["n"], c["num_steps"], c["seed"],
                               c["sobol"], c["anti"], c["cv"], rng="reference")
        got = eng.price(c["spot"], c["strike"], c["T"], c["is_call"])
        assert set(got) == set(c["result"])
        for k, w in c["result"].items():
            assert got[k] == pytest.approx(w, rel=1e-9, abs=1e-8), (k, c)


def test_reference_rng_price_batch_and_sample_paths(golden, garr):
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    for c in golden["cases"]["price_batch"]:
        eng = MonteCarloEngine(SVJParams(**golden["params"][c["params"]]), c["n"], c["num_steps"], c["seed"], False,
                               c["anti"], c["cv"], rng="reference")
        got = eng.price_batch(c["spot"], np.array(c["strikes"]), c["T"], c["is_call"])
        for g, w in zip(got, c["result"]):
            assert set(g) == set(w)
            for k in w:
                assert g[k] == pytest.approx(w[k], rel=1e-9, abs=1e-8)
    eng = MonteCarloEngine(SVJParams(**golden["params"]["svj_default"]), 1000, seed=42, rng="reference")
    got = eng.get_sample_paths(22500.0, 0.1, 8)
    np.testing.assert_allclose(got, garr["sample_paths_svj"], rtol=RTOL64)


def test_philox_price_batch_keys_and_consistency(golden):
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    c = golden["cases"]["price_batch"][0]
    p = SVJParams(**golden["params"][c["params"]])
    eng = MonteCarloEngine(p, 500_000, c["num_steps"], c["seed"], False, c["anti"], c["cv"], rng="philox")
    got = eng.price_batch(c["spot"], np.array(c["strikes"]), c["T"], c["is_call"])
    assert len(got) == len(c["result"])
    for g, w in zip(got, c["result"]):
        assert set(g) == set(w) and g["strike"] == w["strike"]
        # raw std errors: 4096 paths in the golden, 500k here
        assert abs(g["price"] - w["price"]) <= 4 * math.hypot(g["std_error"], w["std_error"]) + 1e-9
        single = eng.price(c["spot"], w["strike"], c["T"], c["is_call"])
        assert single["price"] == pytest.approx(g["price"], rel=1e-9)      # same paths whatever the strike count


def test_price_grid_equals_price_batch_per_expiry(golden):
    """BASELINE config 3 in miniature: 64 strikes x 4 expiries, (n_mat, n_k) layout, same numbers as price_batch."""
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    p = SVJParams(**golden["params"]["gbm_cfg1"])
    ks = np.linspace(0.7, 1.3, 64) * 2500.0
    Ts = [0.125, 0.25, 1.0, 2.0]
    eng = MonteCarloEngine(p, 200_000, 250, 42, use_sobol=False, use_antithetic=True, use_control_variate=False, rng="philox")
    g = eng.price_grid(2500.0, ks, Ts, True)
    assert g["prices"].shape == (4, 64) and list(g["num_steps"]) == [31, 62, 250, 500]
    for j, T in enumerate(Ts):
        rows = eng.price_batch(2500.0, ks, T, True)
        np.testing.assert_allclose(g["prices"][j], [r["price"] for r in rows], rtol=1e-12)
        np.testing.assert_allclose(g["std_errors"][j], [r["std_error"] for r in rows], rtol=1e-12)
        bs = np.array([O.bs_price(2500.0, K, T, p.r, p.q, 0.3, True) for K in ks])
        assert np.all(np.abs(g["prices"][j] - bs) <= 4.5 * g["std_errors"][j] + 1e-9)
    assert np.all(np.diff(g["prices"], axis=1) <= 1e-9)          # calls decrease in strike (same paths)
    # benchmark reading: every cell has its own paths (disjoint counter ranges); cells dealt to ranks give the same grid
    from monte_carlo_option_simulator_b200.dist import Comm
    ind = eng.price_grid(2500.0, ks[::8], Ts[:2], True, independent_cells=True)
    assert ind["prices"].shape == (2, 8)
    one = MonteCarloEngine(p, 200_000, 250, 42, use_sobol=False, use_antithetic=True, use_control_variate=False, rng="philox")
    cell = 1 * 8 + 3
    row = one.handle.price_european(p, 2500.0, Ts[1], 62, 200_000, 42, [ks[24]], True, one._flags(), None, path_offset=cell * 200_000)[0]
    assert ind["prices"][1, 3] == pytest.approx(math.exp(-p.r * Ts[1]) * 0.5 * (row[1] + row[2]) / row[0], rel=1e-12)
    # three emulated ranks (threads, one handle each): cells are dealt round-robin, one all-reduce gathers the sums
    import threading
    from monte_carlo_option_simulator_b200 import _lib
    shared = {"buf": [None] * 3, "bar": threading.Barrier(3)}
    res = [None] * 3

    def run(r):
        hh = _lib.Handle(0)
        try:
            e3 = MonteCarloEngine(p, 200_000, 250, 42, use_sobol=False, use_antithetic=True, use_control_variate=False,
                                  rng="philox", handle=hh, comm=_ThreadComm(r, 3, shared))
            res[r] = e3.price_grid(2500.0, ks[::8], Ts[:2], True, independent_cells=True)
        finally:
            hh.close()

    th = [threading.Thread(target=run, args=(r,)) for r in range(3)]
    [x.start() for x in th]
    [x.join() for x in th]
    for r in range(3):
        np.testing.assert_allclose(res[r]["prices"], ind["prices"], rtol=1e-12)


# ---------------------------------------------------------------------------------------------- Greeks
def test_greeks_engine_reference_rng_matches_golden(golden):
    from monte_carlo_option_simulator_b200 import GreeksEngine, SVJParams
    c = [g for g in golden["cases"]["greeks"] if g["n"] <= 4096][0]
    g = GreeksEngine(SVJParams(**golden["params"][c["params"]]), c["n"], c["num_steps"], c["seed"], rng="reference")
    args = (c["spot"], c["strike"], c["T"], c["is_call"])
    for name in ("delta", "vega", "gamma"):
        got = getattr(g, name)(*args)
        assert set(got) == set(c[name])
        for k, w in c[name].items():
            assert got[k] == pytest.approx(w, rel=1e-7, abs=1e-9), (name, k)


def test_greeks_engine_philox_vs_black_scholes(golden):
    """cfg2: European call/put + all Greeks from ONE fused launch (GBM, kappa = 0), against closed forms."""
    from monte_carlo_option_simulator_b200 import GreeksEngine, SVJParams
    from scipy.stats import norm
    S0 = K = 2500.0
    T, r, sig = 1.0, 0.065, 0.3
    p = SVJParams.gbm(sig, r=r, q=0.0)
    d1 = (math.log(S0 / K) + (r + 0.5 * sig * sig) * T) / (sig * math.sqrt(T))
    d2 = d1 - sig * math.sqrt(T)
    bs_gamma = norm.pdf(d1) / (S0 * sig * math.sqrt(T))
    bs_vega = S0 * norm.pdf(d1) * math.sqrt(T)
    for is_call in (True, False):
        g = GreeksEngine(p, 10_000_000, 250, seed=42, rng="philox")
        d = g.delta(S0, K, T, is_call)
        bs_d = norm.cdf(d1) if is_call else norm.cdf(d1) - 1
        assert d["pathwise"] == pytest.approx(bs_d, abs=2e-3) and d["finite_diff"] == pytest.approx(bs_d, abs=2e-3)
        assert d["diff_pct"] < 0.5
        assert g.gamma(S0, K, T, is_call)["gamma"] == pytest.approx(bs_gamma, rel=0.03)
        v = g.vega(S0, K, T, is_call)
        assert v["vega_per_vol_point"] == pytest.approx(bs_vega, rel=0.01)
        assert v["pathwise_vega_sigma"] == pytest.approx(bs_vega, rel=0.01)
        bs_rho = K * T * math.exp(-r * T) * (norm.cdf(d2) if is_call else -norm.cdf(-d2))
        assert g.rho(S0, K, T, is_call)["rho_crn"] == pytest.approx(bs_rho, rel=0.01)


# ---------------------------------------------------------------------------------------------- path store
@pytest.mark.parametrize("mode", ["gbm", "detvar", "heston", "svj"])
@pytest.mark.parametrize("n,steps", [(70, 250), (33, 31), (5, 50), (64, 96)])
def test_generate_paths_identical_draws(H, L, golden, mode, n, steps):
    p, stream = _mode_params(golden, mode)
    seed, off, S0, T = 4, 1000, 2500.0, 1.0
    Z = _draws(H, L, seed, n, steps, stream, off, p, T)
    want = O._sim(p, S0, T, *Z, steps, record=True)[2]
    got = H.generate_paths(p, S0, T, steps, n, seed, L.FP64, np.float64, off)
    assert got.shape == (n, steps + 1) and np.all(got[:, 0] == S0)
    np.testing.assert_allclose(got, want, rtol=1e-10)
    got32 = H.generate_paths(p, S0, T, steps, n, seed, 0, np.float32, off)
    assert got32.dtype == np.float32
    np.testing.assert_allclose(got32, want, rtol=RTOL32)
    padded = H.generate_paths(p, S0, T, steps, n, seed, L.FP64, np.float64, off, ld=steps + 4)
    np.testing.assert_allclose(padded, got, rtol=1e-13)     # row-tiled kernel vs TMA-tiled kernel: summation order


@pytest.mark.parametrize("steps", [1000, 1024, 1500])
def test_generate_paths_long_paths(H, L, golden, steps):
    """Many steps: 32 warps per CTA in the TMA-tiled kernel (fp32), shared-memory limit fallback (fp64), row-tiled
    kernel beyond 1024 steps."""
    p, stream = _mode_params(golden, "gbm")
    n, seed, T = 70, 8, 2.0
    Z = _draws(H, L, seed, n, steps, stream, 0, p, T)
    want = O._sim(p, 2500.0, T, *Z, steps, record=True)[2]
    np.testing.assert_allclose(H.generate_paths(p, 2500.0, T, steps, n, seed, L.FP64, np.float64), want, rtol=1e-10)
    np.testing.assert_allclose(H.generate_paths(p, 2500.0, T, steps, n, seed, 0, np.float32), want, rtol=RTOL32)


def test_generate_paths_terminal_column_equals_terminal_mode(H, L, golden):
    p, _ = _mode_params(golden, "svj")
    paths = H.generate_paths(p, 100.0, 0.5, 77, 1000, 3, L.FP64, np.float64)
    S = H.simulate_terminal(p, 100.0, 0.5, 77, 1000, 3, L.FP64, np.float64)[0]
    np.testing.assert_allclose(paths[:, -1], S, rtol=1e-13)


def test_sharded_path_store_rows_equal_single_run(H, L, golden):
    """SURVEY 8e, path-storing mode: every rank keeps its own shard, rows are those of the single-GPU run."""
    from monte_carlo_option_simulator_b200.dist import Comm, sharded_generate_paths
    p, _ = _mode_params(golden, "gbm")
    whole = H.generate_paths(p, 2500.0, 1.0, 250, 1000, 9, 0, np.float32)
    for world in (2, 3):
        for r in range(world):
            c = Comm()
            c.rank, c.world = r, world
            lo, hi, part = sharded_generate_paths(H, c, p, 2500.0, 1.0, 250, 1000, 9, 0, np.float32)
            np.testing.assert_array_equal(part, whole[lo:hi])


def test_sample_paths_philox_shape(golden):
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    eng = MonteCarloEngine(SVJParams(**golden["params"]["svj_default"]), 1000, seed=42, rng="philox")
    sp = eng.get_sample_paths(22500.0, 0.1, 8)
    assert sp.shape == (8, 51) and sp.dtype == np.float64 and np.all(sp[:, 0] == 22500.0) and np.all(sp > 0)


# ---------------------------------------------------------------------------------------------- a10: risk metrics
def test_risk_metrics_golden(H, golden, garr):
    from monte_carlo_option_simulator_b200 import compute_risk_metrics
    for c in golden["cases"]["risk"]:
        got = compute_risk_metrics(garr[f"risk_{c['name']}"], c["confidence"], handle=H)
        want = unnan(c["result"])
        assert set(got) == set(want)
        for k, w in want.items():
            if np.isnan(w):
                assert np.isnan(got[k]), (c["name"], k)
            else:
                assert got[k] == pytest.approx(w, rel=1e-11, abs=1e-13), (c["name"], k)


def test_risk_metrics_large_and_fp32(H):
    g = np.random.default_rng(0)
    x = g.standard_t(4, size=4_000_000) * 0.01
    got = H.risk_metrics(x, 0.99)
    want = O.risk_metrics(x, 0.99)
    for k, v in zip(("var", "cvar", "skewness", "kurtosis", "excess_kurtosis", "tail_index", "mean", "std"), got):
        assert v == pytest.approx(want[k], rel=1e-9, abs=1e-12), k
    x32 = x.astype(np.float32)
    got32 = H.risk_metrics(x32, 0.95)
    want32 = O.risk_metrics(x32.astype(np.float64), 0.95)
    assert got32[0] == pytest.approx(want32["var"], rel=1e-12) and got32[1] == pytest.approx(want32["cvar"], rel=1e-9)


@pytest.mark.parametrize("case", cases(16))
def test_risk_metrics_random(H, case):
    """Randomised: sizes from 1 to 3e5, heavy tails, heavy ties, few or no losses, float32 input, several confidence
    levels -- device radix select against the oracle's sort (engine/risk.py:117-173 conventions)."""
    g = np.random.default_rng(500 + case)
    n = int(g.choice([1, 2, 19, 21, 22, 97, 1000, 4001, 50_000, 300_001]))
    kind = int(g.integers(0, 5))
    if kind == 0:
        x = g.standard_normal(n) * 0.02
    elif kind == 1:
        x = g.standard_t(3, size=n) * 0.01
    elif kind == 2:
        x = np.round(g.standard_normal(n), 1)                  # heavy ties, exact zeros
    elif kind == 3:
        x = np.abs(g.standard_normal(n)) + (g.random(n) < 15 / max(n, 15)) * -5.0     # ~15 losses at most
    else:
        x = -np.abs(g.standard_t(4, size=n))                    # every return is a loss
    conf = float(g.choice([0.9, 0.95, 0.99, 0.999]))
    if bool(g.integers(0, 2)):
        x = x.astype(np.float32)
    got = H.risk_metrics(x, conf)
    want = O.risk_metrics(x.astype(np.float64), conf)
    for k, v in zip(("var", "cvar", "skewness", "kurtosis", "excess_kurtosis", "tail_index", "mean", "std"), got):
        w = want[k]
        if math.isnan(w):
            assert math.isnan(v), (k, n, kind)
        else:
            assert v == pytest.approx(w, rel=1e-9, abs=1e-12), (k, n, kind)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_risk_metrics_device_vectors_at_any_alignment(H, dtype):
    """The passes read 16 bytes per load: a device vector that starts off a 16-byte boundary (a view into a path matrix, a
    shard) or is shorter than one vector takes the peeled head / tail loops -- same metrics as the oracle either way."""
    g = np.random.default_rng(77)
    host = (g.standard_t(4, size=200_003) * 0.01).astype(dtype)
    item = np.dtype(dtype).itemsize
    buf = H.malloc(host.nbytes)
    try:
        H.h2d(buf, host)
        for off, n in ((0, 200_003), (1, 200_000), (2, 199_999), (3, 65_537), (1, 1), (1, 2), (3, 5), (0, 3), (1, 31), (2, 4097)):
            got = H.risk_metrics(buf + off * item, 0.99, n=n, dtype=dtype)
            want = O.risk_metrics(host[off:off + n].astype(np.float64), 0.99)
            for k, v in zip(("var", "cvar", "skewness", "kurtosis", "excess_kurtosis", "tail_index", "mean", "std"), got):
                if math.isnan(want[k]):
                    assert math.isnan(v), (off, n, k)
                else:
                    assert v == pytest.approx(want[k], rel=1e-9, abs=1e-12), (off, n, k)
    finally:
        H.free(buf)


def test_risk_metrics_edge_cases(H, L):
    from monte_carlo_option_simulator_b200 import compute_risk_metrics
    with pytest.raises(IndexError):
        compute_risk_metrics(np.array([]))
    one = compute_risk_metrics(np.array([-0.5]), handle=H)
    assert one["var"] == 0.5 and one["cvar"] == 0.5 and math.isnan(one["tail_index"]) and one["std"] == 0.0
    same = compute_risk_metrics(np.full(1000, -1.0), handle=H)
    want = O.risk_metrics(np.full(1000, -1.0))
    assert same["var"] == want["var"] and same["cvar"] == pytest.approx(want["cvar"]) and math.isnan(same["tail_index"])


def test_cfg4_terminal_pnl_var(H, L, golden):
    """BASELINE config 4 reduced: terminal P&L of a short-dated option, device select vs oracle sort."""
    p, _ = _mode_params(golden, "gbm")
    S = H.simulate_terminal(p, 2500.0, 1.0, 250, 300_000, 42, 0, np.float64)[0]
    pnl = math.exp(-p.r) * np.maximum(S - 2500.0, 0.0) - 374.07
    got = H.risk_metrics(pnl, 0.99)
    want = O.risk_metrics(pnl, 0.99)
    for k, v in zip(("var", "cvar", "skewness", "kurtosis", "excess_kurtosis", "tail_index", "mean", "std"), got):
        if math.isnan(want[k]):
            assert math.isnan(v)
        else:
            assert v == pytest.approx(want[k], rel=1e-9, abs=1e-9), k


class _ThreadComm:
    """In-process communicator for emulating ranks with threads (one Handle per thread on the same GPU)."""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self.s = rank, world, shared

    def allreduce_sum(self, a):
        import numpy as _np
        self.s["buf"][self.rank] = _np.array(a, dtype=_np.float64, copy=True)
        self.s["bar"].wait()
        out = sum(self.s["buf"])
        self.s["bar"].wait()
        return out


@pytest.mark.parametrize("world", [1, 2, 3])
def test_sharded_risk_metrics_equal_the_global_ones(L, garr, world):
    """VaR/CVaR/Hill over a P&L vector sharded across ranks; only histograms and a few sums are exchanged.
    Ranks are emulated as threads with one handle each; shards are ragged (one of them may be empty)."""
    import threading
    from monte_carlo_option_simulator_b200.risk import compute_risk_metrics_sharded
    g = np.random.default_rng(5)
    cases = [(garr["risk_student_t3"], 0.99), (garr["risk_ties"], 0.95), (g.standard_t(4, size=300_001) * 0.01, 0.99),
             (garr["risk_tiny"], 0.99)]
    for x, conf in cases:
        want = O.risk_metrics(x, conf)
        cuts = [0] + sorted(g.integers(0, x.size + 1, size=world - 1).tolist()) + [x.size]
        shared = {"buf": [None] * world, "bar": threading.Barrier(world)}
        res = [None] * world

        def run(r):
            h = L.Handle(0)
            try:
                res[r] = compute_risk_metrics_sharded(x[cuts[r]:cuts[r + 1]], conf, comm=_ThreadComm(r, world, shared), handle=h)
            finally:
                h.close()

        th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        [t_.start() for t_ in th]
        [t_.join() for t_ in th]
        for r in range(world):
            assert res[r] is not None
            for k, w in want.items():
                if math.isnan(w):
                    assert math.isnan(res[r][k]), k
                else:
                    assert res[r][k] == pytest.approx(w, rel=1e-10, abs=1e-13), (k, world)


def test_cfg4_pipeline_stays_on_device(H, L, golden):
    """paths / terminal spots -> option P&L -> VaR/CVaR, all device resident, against the host pipeline + oracle sort."""
    from monte_carlo_option_simulator_b200.risk import terminal_pnl_metrics
    from monte_carlo_option_simulator_b200 import SVJParams
    p = SVJParams(**golden["params"]["gbm_cfg1"])
    got = terminal_pnl_metrics(p, 2500.0, 2500.0, 1.0, 200_000, 250, seed=42, handle=H)
    S = H.simulate_terminal(p, 2500.0, 1.0, 250, 200_000, 42, L.FP64, np.float64)[0]
    pay = math.exp(-p.r) * np.maximum(S - 2500.0, 0.0)
    assert got["premium"] == pytest.approx(pay.mean(), rel=1e-12)
    want = O.risk_metrics(pay - pay.mean(), 0.99)
    for k, w in want.items():
        if math.isnan(w):          # all the largest losses are ties (-premium): the Hill sum is 0 -> NaN, as in the reference
            assert math.isnan(got[k]), k
        else:
            assert got[k] == pytest.approx(w, rel=1e-9, abs=1e-9), k
    # the last column of a device-resident path matrix works as the spot vector (stride = ld)
    n, steps = 5000, 50
    mat = H.malloc(n * (steps + 1) * 8)
    pnl = H.malloc(n * 8)
    try:
        H.generate_paths(p, 2500.0, 1.0, steps, n, 3, L.FP64, np.float64, out_dev=mat)
        H.option_pnl(mat + steps * 8, n, 2400.0, False, 0.9, 10.0, pnl, stride=steps + 1)
        host = np.empty(n)
        H.d2h(host, pnl)
    finally:
        H.free(mat)
        H.free(pnl)
    paths = H.generate_paths(p, 2500.0, 1.0, steps, n, 3, L.FP64, np.float64)
    np.testing.assert_allclose(host, 0.9 * np.maximum(2400.0 - paths[:, -1], 0.0) - 10.0, rtol=1e-13, atol=1e-12)


# ---------------------------------------------------------------------------------------------- patched reference callers
def test_full_size_linearity_property(H, L, golden):
    """BASELINE config 2 size (1e7 x 250): sums over two halves of the path range add up to the whole-range sums
    (the property multi-GPU sharding relies on), and sum_s/n hits the forward within 4 standard errors."""
    p, _ = _mode_params(golden, "gbm")
    n = 10_000_000
    whole = H.price_european(p, 2500.0, 1.0, 250, n, 42, [2500.0], True, 0)[0]
    a = H.price_european(p, 2500.0, 1.0, 250, n // 2, 42, [2500.0], True, 0)[0]
    b = H.price_european(p, 2500.0, 1.0, 250, n - n // 2, 42, [2500.0], True, 0, path_offset=n // 2)[0]
    np.testing.assert_allclose(a + b, whole, rtol=1e-11)
    mean_s = whole[L.SUMS_FIELDS.index("sum_s")] / n
    sd_s = math.sqrt(whole[L.SUMS_FIELDS.index("sum_ss")] / n - mean_s ** 2)
    assert abs(mean_s - 2500.0 * math.exp(p.r - p.q)) <= 4 * sd_s / math.sqrt(n)


def test_cfg2_full_size_greeks_linearity_fp32_and_fp64(H, L, golden):
    """BASELINE config 2 at its full size (1e7 paths x 250 steps, call + every bump accumulator), fp32 and fp64 state:
    three unequal shards of the path range add up to the whole-range sums (all 17 of them), the two precisions agree on
    the same draws within the north star's fp32 tolerance, and the pathwise delta sits within 4 standard errors of
    Black-Scholes' N(d1)."""
    from monte_carlo_option_simulator_b200 import bs_delta
    p, _ = _mode_params(golden, "gbm")
    n = 10_000_000
    bumps = L.Bumps(0.01, p.v0 + 0.01, max(p.v0 - 0.01, 0.001), p.r + 1e-4, max(p.r - 1e-4, 0))
    cuts = [0, 1_234_567, 7_000_001, n]
    rows = {}
    for fl, name, tol in ((L.GREEKS, "f32", 2e-7), (L.GREEKS | L.FP64, "f64", 1e-11)):
        whole = H.price_european(p, 2500.0, 1.0, 250, n, 42, [2500.0], True, fl, bumps)[0]
        parts = sum(H.price_european(p, 2500.0, 1.0, 250, hi - lo, 42, [2500.0], True, fl, bumps, path_offset=lo)[0]
                    for lo, hi in zip(cuts[:-1], cuts[1:]))
        np.testing.assert_allclose(parts, whole, rtol=tol)        # fp32: per-thread partials regroup with the launch shape
        rows[name] = whole
    np.testing.assert_allclose(rows["f32"], rows["f64"], rtol=1e-4)
    i_pw, i_n = L.SUMS_FIELDS.index("sum_pw_delta"), L.SUMS_FIELDS.index("n")
    delta = math.exp(-p.r) * rows["f64"][i_pw] / rows["f64"][i_n]
    assert abs(delta - bs_delta(2500.0, 2500.0, 1.0, p.r, p.q, math.sqrt(p.v0), True)) < 4 * 0.6 / math.sqrt(n) + 1e-4


def test_cfg5_full_size_eight_way_split_adds_up(H, L, golden):
    """BASELINE config 5's smallest point (1e8 total paths x 250 steps) the way eight GPUs take it: the eight contiguous
    shards of dist.shard_range, launched one after the other here, add up to the single whole-range launch."""
    from monte_carlo_option_simulator_b200.dist import shard_range
    p, _ = _mode_params(golden, "gbm")
    n = 100_000_000
    whole = H.price_european(p, 2500.0, 1.0, 250, n, 7, [2500.0], True, 0)[0]
    parts = 0
    covered = 0
    for r in range(8):
        lo, hi = shard_range(n, r, 8)
        assert lo == covered
        covered = hi
        parts = parts + H.price_european(p, 2500.0, 1.0, 250, hi - lo, 7, [2500.0], True, 0, path_offset=lo)[0]
    assert covered == n and parts[L.SUMS_FIELDS.index("n")] == n
    np.testing.assert_allclose(parts, whole, rtol=2e-7)
    price = math.exp(-p.r) * whole[L.SUMS_FIELDS.index("sum_a")] / n
    from monte_carlo_option_simulator_b200 import bs_price
    assert abs(price - bs_price(2500.0, 2500.0, 1.0, p.r, p.q, math.sqrt(p.v0), True)) < 4 * 570.0 / math.sqrt(n)


def test_cfg4_full_size_store_properties(H, L, golden):
    """BASELINE config 4 at its full size (4M paths x 251 columns, fp32, reference layout), checked on the device through
    properties that do not need a 4 GB oracle: column 0 is the spot, the last column is the terminal mode's value for the
    same path (two different kernels: additive log state vs multiplicative carry) within the fp32 tolerance, interior
    columns are martingales up to the drift, a shard stored on its own equals the same rows of the full store bit for
    bit, and paths -> option P&L -> VaR / CVaR on the device equals a sort of the same P&L vector."""
    import torch
    p, _ = _mode_params(golden, "gbm")
    n, steps, S0 = 4_000_000, 250, 2500.0
    mat = torch.empty((n, steps + 1), dtype=torch.float32, device="cuda")
    H.generate_paths(p, S0, 1.0, steps, n, 42, 0, np.float32, out_dev=mat.data_ptr())
    term = torch.empty(n, dtype=torch.float32, device="cuda")
    H.simulate_terminal(p, S0, 1.0, steps, n, 42, 0, np.float32, dev_ptrs=(term.data_ptr(), 0, 0))
    H.synchronize()
    assert bool((mat[:, 0] == S0).all()) and bool((mat > 0).all())
    rel = (mat[:, -1].double() / term.double() - 1).abs()
    assert float(rel.max()) < 1e-4 and float(rel.median()) < 2e-6
    for c in (1, 50, 125, 250):
        col = mat[:, c].double()
        fwd = S0 * math.exp((p.r - p.q) * c / steps)
        assert abs(float(col.mean()) - fwd) < 4 * float(col.std()) / math.sqrt(n) + 1e-3
    lo, hi = 1_234_567, 1_300_000
    part = torch.empty((hi - lo, steps + 1), dtype=torch.float32, device="cuda")
    H.generate_paths(p, S0, 1.0, steps, hi - lo, 42, 0, np.float32, path_offset=lo, out_dev=part.data_ptr())
    H.synchronize()
    assert torch.equal(part, mat[lo:hi])
    pnl = torch.empty(n, dtype=torch.float64, device="cuda")
    disc, premium = math.exp(-p.r), 374.0712289657911
    H.option_pnl(mat.data_ptr() + steps * 4, n, 2500.0, True, disc, premium, pnl.data_ptr(), dtype_in=np.float32, stride=steps + 1)
    got = H.risk_metrics(pnl.data_ptr(), 0.99, n=n, dtype=np.float64)
    want_pnl = disc * torch.clamp(mat[:, -1].double() - 2500.0, min=0.0) - premium
    assert bool(torch.allclose(want_pnl, pnl, rtol=1e-15, atol=1e-12))      # (the kernel may contract multiply and subtract)
    srt = torch.sort(pnl).values
    cut = int(n * (1 - 0.99))
    assert got[0] == pytest.approx(-float(srt[cut]), rel=1e-12)
    assert got[1] == pytest.approx(-float(srt[:cut].mean()), rel=1e-9)
    assert got[6] == pytest.approx(float(pnl.mean()), rel=1e-9, abs=1e-9)


def test_cfg3_full_size_grid_properties(golden):
    """BASELINE config 3 at its full size (64 strikes x 16 expiries, 1M paths per expiry shared across the strikes):
    on shared paths call - put equals D (mean S_T - K) path set by path set, so the two grids must satisfy put-call
    parity against the SAME run's forward to rounding; calls fall and are convex in the strike, both are within 4
    standard errors of Black-Scholes, and the independent-cells reading agrees with the shared-paths one statistically."""
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams, bs_price
    p = SVJParams(**golden["params"]["gbm_cfg1"])
    strikes = np.linspace(0.7 * 2500.0, 1.3 * 2500.0, 64)
    mats = np.linspace(1 / 12, 2.0, 16)
    eng = MonteCarloEngine(p, 1_000_000, 250, seed=11, use_sobol=False, use_antithetic=False, use_control_variate=False, rng="philox")
    calls = eng.price_grid(2500.0, strikes, mats, True)
    puts = eng.price_grid(2500.0, strikes, mats, False)
    assert calls["prices"].shape == (16, 64)
    for j, T in enumerate(mats):
        D = math.exp(-p.r * T)
        # the forward of THIS path set, from the deepest in-the-money call and put of the row (parity at one strike) ...
        fwd = (calls["prices"][j, 0] - puts["prices"][j, 0]) / D + strikes[0]
        # ... must then hold at every other strike of the row, to fp32-sum rounding
        np.testing.assert_allclose(calls["prices"][j] - puts["prices"][j], D * (fwd - strikes), rtol=0, atol=2e-3)
        assert abs(fwd - 2500.0 * math.exp((p.r - p.q) * T)) < 4 * 2500.0 * math.sqrt(p.v0 * T) * 1.2 / 1e3
        c = calls["prices"][j]
        assert np.all(np.diff(c) < 0) and np.all(np.diff(c, 2) > -1e-3)                  # monotone, convex (same paths)
        bs = np.array([bs_price(2500.0, K, T, p.r, p.q, math.sqrt(p.v0), True) for K in strikes])
        assert np.all(np.abs(c - bs) < 4 * calls["std_errors"][j] + 1e-3 * bs + 1e-6)
    ind = MonteCarloEngine(p, 200_000, 250, seed=11, use_sobol=False, use_antithetic=False, use_control_variate=False,
                           rng="philox").price_grid(2500.0, strikes[::8], mats[::4], True, independent_cells=True)
    sub = calls["prices"][::4, ::8]
    assert np.all(np.abs(ind["prices"] - sub) < 5 * np.hypot(ind["std_errors"], calls["std_errors"][::4, ::8]) + 1e-6)
