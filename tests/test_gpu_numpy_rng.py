"""GPU: NumPy's pseudo-random front end (PCG64 + Ziggurat standard_normal, PCG64 random) generated on the device must equal
NumPy's own arrays BIT FOR BIT -- it is what makes rng="reference" (use_sobol=False) and GreeksEngine's shared randoms
(engine/greeks.py:33-41, engine/monte_carlo.py:301-308, :458-462) both exact and fast."""
import numpy as np
import pytest

from monte_carlo_option_simulator_b200 import GreeksEngine, MonteCarloEngine, SVJParams, _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    h = _lib.Handle(0)
    yield h
    h.close()


@pytest.mark.parametrize("seed,n", [(42, 1), (42, 127), (42, 128), (7, 129), (7, 10_000), (123456789, 300_001), (0, 3_000_000)])
def test_standard_normal_bitwise(H, seed, n):
    got, used = H.numpy_fill(seed, n)
    g = np.random.default_rng(seed)
    want = g.standard_normal(n)
    np.testing.assert_array_equal(got.view(np.uint64), want.view(np.uint64))
    # the generator outputs consumed: what NumPy's own generator has advanced by (next raw output must agree)
    nxt = np.random.PCG64(seed).random_raw(used + 1)[used]
    assert g.bit_generator.random_raw(1)[0] == nxt
    assert n <= used <= 1.05 * n + 64


def test_reference_size_stream_and_chained_calls(H):
    """The reference's shape: three (n, steps) normal arrays from ONE generator = one stream; seeds {42, 7}."""
    n, steps = 50_000, 250
    for seed in (42, 7):
        g = np.random.default_rng(seed)
        want = np.concatenate([g.standard_normal((n, steps)).ravel() for _ in range(3)])
        got, used = H.numpy_fill(seed, 3 * n * steps)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
        # a second call continues where the first stopped (first_raw = outputs consumed)
        more, _ = H.numpy_fill(seed, 1000, first_raw=used)
        assert np.array_equal(more, g.standard_normal(1000))
        uni, _ = H.numpy_fill(seed, 1000, _lib.NUMPY_RANDOM, first_raw=used)
        raw = np.random.PCG64(seed).random_raw(used + 1000)[used:]
        assert np.array_equal(uni, (raw >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0))


def test_tail_draws_are_exercised_and_exact(H):
    """|z| > 3.654 comes only from the log1p tail loop: enough draws to see hundreds of them, all bit-exact."""
    n = 4_000_000
    got, _ = H.numpy_fill(2024, n)
    want = np.random.default_rng(2024).standard_normal(n)
    tail = np.abs(want) > 3.6541528853610088
    assert tail.sum() > 500
    assert np.array_equal(got[tail].view(np.uint64), want[tail].view(np.uint64))
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_reference_draws_object_equals_numpy(H):
    n, steps, seed = 3000, 37, 11
    d = _lib.ReferenceDraws(H, seed, n, steps)
    Z1, Z2, Zj, Zjs = d.host_arrays()
    g = np.random.default_rng(seed)
    for got in (Z1, Z2, Zjs):
        assert np.array_equal(got, g.standard_normal((n, steps)))
    assert np.array_equal(Zj, np.random.default_rng(seed + 1).random((n, steps)))
    d.close()
    # get_sample_paths order: ONE generator, normals then uniforms
    d = _lib.ReferenceDraws(H, seed + 999, 50, 63, uniform_seed=None)
    Z1, Z2, Zj, Zjs = d.host_arrays()
    g = np.random.default_rng(seed + 999)
    for got in (Z1, Z2, Zjs):
        assert np.array_equal(got, g.standard_normal((50, 63)))
    assert np.array_equal(Zj, g.random((50, 63)))
    d.close()


def test_reference_mode_pcg64_prices_equal_the_host_front_end(H, monkeypatch):
    """rng="reference", use_sobol=False: the device front end gives the same dicts as NumPy draws on the host, to the
    last bit of the terminal spots (same recurrence kernel, same doubles), including the antithetic twin."""
    p = SVJParams()
    kw = dict(num_paths=20_000, num_steps=252, seed=42, use_sobol=False, use_antithetic=True, use_control_variate=True,
              rng="reference", handle=H)
    dev = MonteCarloEngine(p, **kw).price(22500.0, 22500.0, 0.25, True)
    devb = MonteCarloEngine(p, **kw).price_batch(22500.0, [21000.0, 22500.0, 24000.0], 0.25, True)
    devs = MonteCarloEngine(p, **kw).get_sample_paths(22500.0, 0.25, 50)
    devg = GreeksEngine(p, 20_000, 252, 42, rng="reference", handle=H)
    dg = (devg.delta(22500.0, 22500.0, 0.25), devg.vega(22500.0, 22500.0, 0.25), devg.gamma(22500.0, 22500.0, 0.25))
    monkeypatch.setenv("B200MC_REFERENCE_PCG64", "host")
    host = MonteCarloEngine(p, **kw).price(22500.0, 22500.0, 0.25, True)
    hostb = MonteCarloEngine(p, **kw).price_batch(22500.0, [21000.0, 22500.0, 24000.0], 0.25, True)
    hosts = MonteCarloEngine(p, **kw).get_sample_paths(22500.0, 0.25, 50)
    hostg = GreeksEngine(p, 20_000, 252, 42, rng="reference", handle=H)
    hg = (hostg.delta(22500.0, 22500.0, 0.25), hostg.vega(22500.0, 22500.0, 0.25), hostg.gamma(22500.0, 22500.0, 0.25))
    assert dev == host and devb == hostb and dg == hg
    assert np.array_equal(devs, hosts)


def test_reference_mode_matches_the_reference_golden(H, golden):
    """SURVEY 8(c) SVJ pricer smoke (reference run in the dev container): price 1053.1318703389352 etc. at n = 4096."""
    r = MonteCarloEngine(SVJParams(), 4096, 252, 42, use_sobol=False, use_antithetic=True, use_control_variate=True,
                         rng="reference", handle=H).price(22500.0, 22500.0, 0.25, True)
    assert r["price"] == pytest.approx(1053.1318703389352, rel=1e-10)
    assert r["std_error"] == pytest.approx(19.404241035651566, rel=1e-9)
    assert r["raw_mc_price"] == pytest.approx(1143.3802931930884, rel=1e-10)
