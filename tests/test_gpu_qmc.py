"""GPU tests of the working quasi-Monte Carlo front end: device Sobol points bitwise SciPy's, the
correct Brownian bridge, the sums against the oracle recurrence on the same draws, and the point of it all -- an error
far below plain Monte Carlo's at equal path counts."""
import numpy as np
import pytest

from conftest import cases

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from monte_carlo_option_simulator_b200 import _lib
    return _lib


@pytest.fixture(scope="module")
def H(L):
    h = L.Handle(0)
    yield h
    h.close()


GBM = O.Params(kappa=0.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.0)
HES = O.Params(kappa=3.0, theta=0.04, xi=0.5, rho=-0.7, v0=0.04, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.012)
SVJ = O.Params(kappa=3.0, theta=0.04, xi=0.5, rho=-0.7, v0=0.04, lambda_j=2.0, mu_j=-0.05, sigma_j=0.10, r=0.065, q=0.012)


@pytest.mark.parametrize("steps,n,off", [(1, 5, 0), (7, 100, 3), (63, 257, 0), (250, 64, 1000), (64, 1000, 12345)])
def test_device_draws_equal_scipy_plus_textbook_bridge(H, L, steps, n, off):
    tables = L.sobol_tables(4 * steps, 11)
    Z1, Z2, Zj, Zjs = O.qmc_draws(11, n, steps, 4, path_offset=off)
    for which, want in ((L.Z1, Z1), (L.Z2, Z2), (L.ZJUMP_U, Zj), (L.ZJUMP_SIZE, Zjs)):
        got = H.qmc_normals(tables, n, steps, which, path_offset=off)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    # the uniforms ARE SciPy's points (bitwise, away from the clip)
    import warnings
    from scipy.stats.qmc import Sobol
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        pts = Sobol(d=4 * steps, scramble=True, seed=11).random(off + n)[off:, 3 * steps:]
    np.testing.assert_array_equal(H.qmc_normals(tables, n, steps, L.ZJUMP_U, path_offset=off), np.clip(pts, 1e-10, 1 - 1e-10))


def test_bridge_is_not_degenerate(H, L):
    """W_T = sum of the step normals has variance n_steps (the reference's bridge gives exactly 0), and it
    is carried by dimension 0 alone."""
    steps, n = 63, 4096
    Z1 = H.qmc_normals(L.sobol_tables(steps, 42), n, steps, L.Z1)
    wT = Z1.sum(axis=1)
    assert wT.var() == pytest.approx(steps, rel=5e-3) and abs(wT.mean()) < 0.02
    assert np.abs(Z1.var(axis=0) - 1).max() < 0.05


@pytest.mark.parametrize("p,name", [(GBM, "gbm"), (HES, "heston"), (SVJ, "svj")])
@pytest.mark.parametrize("anti", [False, True])
def test_sums_vs_oracle_on_the_same_draws(H, L, p, name, anti):
    steps, n, off, T, S0 = 40, 777, 5, 0.25, 22500.0
    nb = 4 if p.lambda_j > 0 else (2 if p.xi != 0 else 1)
    tables = L.sobol_tables(nb * steps, 7)
    ks = [21000.0, 22500.0, 24000.0]
    got = H.price_european_qmc(p, S0, T, steps, n, tables, ks, False, L.ANTITHETIC if anti else 0, path_offset=off)
    Z1, Z2, Zj, Zjs = O.qmc_draws(7, n, steps, nb, path_offset=off)
    S = O._sim(p, S0, T, Z1, Z2, Zj, Zjs, steps)[0]
    A = O._sim(p, S0, T, -Z1, -Z2, Zj, -Zjs, steps)[0] if anti else None
    for K, row in zip(ks, got):
        a = np.maximum(K - S, 0.0)
        b = np.maximum(K - A, 0.0) if anti else np.zeros_like(a)
        s_avg = 0.5 * (S + A) if anti else S
        want = [n, a.sum(), b.sum(), (a * a).sum(), (b * b).sum(), (a * b).sum(), s_avg.sum(), (s_avg ** 2).sum(),
                ((0.5 * (a + b) if anti else a) * s_avg).sum()]
        np.testing.assert_allclose(row[:9], want, rtol=1e-9, atol=1e-6)
        assert not row[9:].any()


def test_engine_sobol_mode_beats_plain_monte_carlo(H, L):
    """GBM call, 16384 paths x 64 steps: the QMC error against Black-Scholes is far below the Monte Carlo standard error
    at the same path count; chunked and sharded ranges add up; errors are reported."""
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    p = SVJParams.gbm(0.3, r=0.065)
    bs = O.bs_price(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, True)
    errs = []
    for seed in (1, 2, 3):
        q = MonteCarloEngine(p, 16384, 64, seed, use_antithetic=False, use_control_variate=False, rng="sobol", handle=H)
        errs.append(abs(q.price(2500.0, 2500.0, 1.0)["price"] - bs))
    mc = MonteCarloEngine(p, 16384, 64, 1, use_antithetic=False, use_control_variate=False, handle=H).price(2500.0, 2500.0, 1.0)
    assert max(errs) < 0.1 * mc["std_error"], (errs, mc["std_error"])
    # price_batch / price_grid run through the same front end
    q = MonteCarloEngine(p, 8192, 64, 5, rng="sobol", handle=H)
    rows = q.price_batch(2500.0, [2300.0, 2500.0, 2700.0], 0.5)
    for r_ in rows:
        assert r_["price"] == pytest.approx(O.bs_price(2500.0, r_["strike"], 0.5, 0.065, 0.0, 0.3, True), rel=2e-3)
    # path ranges add (the multi-GPU sharding rule)
    tables = L.sobol_tables(64, 5)
    whole = H.price_european_qmc(p, 2500.0, 1.0, 64, 5000, tables, [2500.0])
    a = H.price_european_qmc(p, 2500.0, 1.0, 64, 2000, tables, [2500.0])
    b = H.price_european_qmc(p, 2500.0, 1.0, 64, 3000, tables, [2500.0], path_offset=2000)
    np.testing.assert_allclose(a + b, whole, rtol=1e-12)
    with pytest.raises(L.B200MCError):
        H.price_european_qmc(p, 2500.0, 1.0, 64, 100, L.sobol_tables(32, 5), [2500.0])        # too few dimensions
    with pytest.raises(L.B200MCError, match="Greek"):
        H.price_european_qmc(p, 2500.0, 1.0, 64, 100, tables, [2500.0], flags=L.GREEKS)
    with pytest.raises(L.B200MCError):
        H.price_european_qmc(p, 2500.0, 1.0, 64, 100, tables, [2500.0], path_offset=2 ** 30)  # beyond the sequence


def test_heston_qmc_within_mc_error_of_plain_mc(H, L):
    from monte_carlo_option_simulator_b200 import MonteCarloEngine
    mc = MonteCarloEngine(HES, 4_000_000, 63, 3, use_control_variate=False, handle=H).price(22500.0, 22500.0, 0.25)
    q = MonteCarloEngine(HES, 65536, 63, 3, use_control_variate=False, rng="sobol", handle=H).price(22500.0, 22500.0, 0.25)
    assert abs(q["price"] - mc["price"]) < 4 * mc["std_error"] + 0.002 * mc["price"]


@pytest.mark.parametrize("pname", ["svj", "heston", "gbm"])
def test_reference_sobol_front_end_on_the_device(H, L, pname, monkeypatch):
    """rng="reference", use_sobol=True -- the reference's DEFAULT configuration: the device-side front end (SciPy's points,
    norm.ppf, the reference's own bridge table, host PCG64 jump uniforms) gives the results of the host front end
    (the reference's NumPy/SciPy code path, pinned to the reference's golden prices elsewhere)."""
    from monte_carlo_option_simulator_b200 import MonteCarloEngine
    p = {"svj": SVJ, "heston": HES, "gbm": GBM}[pname]
    for n, steps_T in ((3000, 0.25), (257, 0.04)):
        kw = dict(num_paths=n, num_steps=252, seed=13, use_sobol=True, rng="reference", handle=H)
        monkeypatch.setenv("B200MC_REFERENCE_SOBOL", "host")
        want = MonteCarloEngine(p, **kw).price(22500.0, 22000.0, steps_T, False)
        want_b = MonteCarloEngine(p, **kw).price_batch(22500.0, [21000.0, 23000.0], steps_T, True)
        monkeypatch.setenv("B200MC_REFERENCE_SOBOL", "device")
        got = MonteCarloEngine(p, **kw).price(22500.0, 22000.0, steps_T, False)
        got_b = MonteCarloEngine(p, **kw).price_batch(22500.0, [21000.0, 23000.0], steps_T, True)
        assert set(got) == set(want)
        for k in want:
            assert got[k] == pytest.approx(want[k], rel=1e-9, abs=1e-9), (pname, k)
        for g_, w_ in zip(got_b, want_b):
            assert g_ == pytest.approx(w_, rel=1e-9, abs=1e-9)
    # the degenerate bridge as a table: the step normals sum to 0 on every path (W_T == 0)
    from monte_carlo_option_simulator_b200.monte_carlo import reference_bridge_nodes
    with pytest.raises(L.B200MCError, match="bridge table"):
        bad = reference_bridge_nodes(10).copy()
        bad[3]["t"] = bad[2]["t"]
        H.qmc_terminal(GBM, 100.0, 1.0, 10, 8, L.sobol_tables(30, 1), bad)


def test_numpy_pcg64_uniforms_on_the_device_bitwise(H, L):
    """default_rng(seed).random(...) -- the reference's jump uniforms (monte_carlo.py:308) -- reproduced on the device."""
    for seed, first, n in ((43, 0, 1), (43, 0, 1000), (0, 5, 64), (2 ** 63 + 11, 123_457, 10_001), (7, 2 ** 33 + 3, 500)):
        want_rng = np.random.default_rng(seed)
        if first < 10 ** 7:
            want = want_rng.random(first + n)[first:]
        else:
            want_rng.bit_generator.advance(first)
            want = want_rng.random(n)
        np.testing.assert_array_equal(H.pcg64_random(seed, n, first), want)
    with pytest.raises(L.B200MCError):
        H.pcg64_random(1, 0)


def test_chunked_runs_equal_offset_runs(H, L):
    """Long paths force the path range through several chunks (2^27 / n_steps paths each for the terminal entry point, 2^28 /
    n_steps for the sums): the result must equal that of separate calls on sub-ranges (which are single chunks)."""
    steps, n = 2000, 150_000                       # chunks of 67_108 paths (terminal) / 134_217 (sums)
    tables = L.sobol_tables(steps, 3)
    p = GBM
    S, A = H.qmc_terminal(p, 100.0, 1.0, steps, n, tables, None, None, L.ANTITHETIC)
    parts = [H.qmc_terminal(p, 100.0, 1.0, steps, hi - lo, tables, None, None, L.ANTITHETIC, path_offset=lo)
             for lo, hi in ((0, 60_000), (60_000, 120_000), (120_000, n))]
    np.testing.assert_array_equal(S, np.concatenate([q[0] for q in parts]))
    np.testing.assert_array_equal(A, np.concatenate([q[1] for q in parts]))
    whole = H.price_european_qmc(p, 100.0, 1.0, steps, n, tables, [100.0], True, L.ANTITHETIC)
    pay_a, pay_b = np.maximum(S - 100.0, 0.0), np.maximum(A - 100.0, 0.0)
    np.testing.assert_allclose(whole[0, :6], [n, pay_a.sum(), pay_b.sum(), (pay_a ** 2).sum(), (pay_b ** 2).sum(),
                                              (pay_a * pay_b).sum()], rtol=1e-12)


def _sobol_model(tables, n, off):
    """Points off .. off + n - 1 of the scrambled sequence from its direction numbers: shift ^ XOR_{b in gray(i)} sv[:, b]
    (what the device evaluates; equal to SciPy's Sobol.random, checked below where SciPy is fast enough to ask)."""
    sv, shift, bits = tables
    out = np.empty((n, sv.shape[0]))
    for r, i in enumerate(range(off, off + n)):
        gray, x, b = i ^ (i >> 1), shift.copy(), 0
        while gray:
            if gray & 1:
                x ^= sv[:, b]
            gray >>= 1
            b += 1
        out[r] = x.astype(np.float64) * 2.0 ** -bits
    return out


@pytest.mark.parametrize("case", cases(12))
def test_device_draws_random_shapes(H, L, case):
    """Random step counts (the bridge over non-powers of two, tiles of 32 / 16 / 8 / fewer paths), path counts and offsets
    deep into the sequence: bridged normals, plain normals and uniforms against the textbook bridge over the points of the
    sequence (SciPy's, where SciPy can fast-forward in reasonable time; its direction numbers otherwise)."""
    import warnings
    from scipy.stats import norm
    from scipy.stats.qmc import Sobol
    g = np.random.default_rng(4000 + case)
    steps = int(g.choice([1, 2, 3, 5, 31, 33, 100, 251, 400, 777, 1500, 3000])) if case < 6 else int(g.integers(1, 500))
    n = int(g.integers(1, 120))
    off = int(g.choice([0, 1, 1023, 2 ** 20 + 17, 2 ** 29 - n - 1]))
    seed = int(g.integers(0, 2 ** 31))
    nb = 1 if steps > 800 else 4
    tables = L.sobol_tables(nb * steps, seed)
    u = _sobol_model(tables, n, off)
    if off <= 1023:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)
            np.testing.assert_array_equal(u, Sobol(d=nb * steps, scramble=True, seed=seed).random(off + n)[off:])
    u = np.clip(u, 1e-10, 1 - 1e-10)
    np.testing.assert_allclose(H.qmc_normals(tables, n, steps, L.Z1, path_offset=off), O.qmc_bridge(norm.ppf(u[:, :steps])),
                               rtol=1e-11, atol=1e-11)
    if nb == 4:
        np.testing.assert_allclose(H.qmc_normals(tables, n, steps, L.ZJUMP_SIZE, path_offset=off), norm.ppf(u[:, 2 * steps:3 * steps]),
                                   rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(H.qmc_normals(tables, n, steps, L.ZJUMP_U, path_offset=off), u[:, 3 * steps:])
