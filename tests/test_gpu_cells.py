"""GPU parity of the (cell x path) grid launch, b200mc_price_cells: every cell against the single-problem entry point
b200mc_price_european (which the other GPU tests pin to the oracle and to the reference's own outputs) and, for
fp64 cells, directly against the oracle fed the identical draws."""
import math

import numpy as np
import pytest

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


def _random_cell(g, mode=None):
    mode = mode or g.choice(["gbm", "detvar", "heston", "svj"])
    v0 = float(g.uniform(0.01, 0.3))
    kw = dict(kappa=0.0, theta=v0, xi=0.0, rho=float(g.uniform(-0.9, 0.9)), v0=v0, lambda_j=0.0,
              mu_j=float(g.uniform(-0.15, 0.05)), sigma_j=float(g.uniform(0.02, 0.25)), r=float(g.uniform(0.0, 0.1)),
              q=float(g.uniform(0.0, 0.04)))
    if mode != "gbm":
        kw.update(kappa=float(g.uniform(0.3, 5.0)), theta=float(g.uniform(0.01, 0.3)))
    if mode in ("heston", "svj"):
        kw.update(xi=float(g.uniform(0.1, 1.0)))
    if mode == "svj":
        kw.update(lambda_j=float(g.uniform(0.3, 6.0)))
    return dict(params=O.Params(**kw), S0=float(g.uniform(10.0, 30000.0)), T=float(g.uniform(0.05, 2.0)),
                n_steps=int(g.integers(1, 130)), n_paths=int(g.integers(1, 3000)), seed=int(g.integers(0, 2 ** 63)),
                path_offset=int(g.choice([0, 5, 2 ** 32 - 100, 2 ** 41])), is_call=bool(g.integers(0, 2)))


@pytest.mark.parametrize("n_strikes", [1, 3, 21])
@pytest.mark.parametrize("fp64", [False, True])
@pytest.mark.parametrize("anti", [False, True])
def test_cells_equal_single_problem_launches(H, L, n_strikes, fp64, anti):
    """37 random cells of all four modes with different sizes in one call == 37 b200mc_price_european calls."""
    g = np.random.default_rng(7 + n_strikes + 2 * fp64 + anti)
    cells = [_random_cell(g) for _ in range(37)]
    ks = np.array([sorted(c["S0"] * g.uniform(0.6, 1.4, size=n_strikes)) for c in cells])
    fl = (L.FP64 if fp64 else 0) | (L.ANTITHETIC if anti else 0)
    got = H.price_cells(cells, ks, fl)
    assert got.shape == (37, n_strikes, L.NSUMS)
    for c, k, rows in zip(cells, ks, got):
        want = H.price_european(c["params"], c["S0"], c["T"], c["n_steps"], c["n_paths"], c["seed"], k, c["is_call"], fl,
                                None, path_offset=c["path_offset"])
        # same draws, same per-path arithmetic, same per-batch fp32 partials: only the order of the fp64 adds differs
        np.testing.assert_allclose(rows[:, :9], want[:, :9], rtol=1e-12, atol=1e-9)
        assert not rows[:, 9:].any()


@pytest.mark.parametrize("mode", ["gbm", "detvar", "heston", "svj"])
def test_cells_against_the_oracle(H, L, mode):
    """fp64 cells vs the oracle recurrence on the cells' own draws (dumped from the device)."""
    g = np.random.default_rng(50 + len(mode))
    cells = [_random_cell(g, mode) for _ in range(5)]
    for c in cells:
        c["n_paths"] = int(g.integers(200, 900))
    ks = np.array([[c["S0"] * 0.95] for c in cells])
    got = H.price_cells(cells, ks, L.FP64 | L.ANTITHETIC)
    for c, k, rows in zip(cells, ks, got):
        p, steps, n = c["params"], c["n_steps"], c["n_paths"]
        stream = L.select_stream(p, c["T"], steps, L.FP64)
        Z = [H.dump_normals(c["seed"], n, steps, stream, w, path_offset=c["path_offset"],
                            jump_prob=p.lambda_j * c["T"] / steps) for w in (L.Z1, L.Z2, L.ZJUMP_U, L.ZJUMP_SIZE)]
        S = O._sim(p, c["S0"], c["T"], Z[0], Z[1], Z[2], Z[3], steps)[0]
        A = O._sim(p, c["S0"], c["T"], -Z[0], -Z[1], Z[2], -Z[3], steps)[0]
        pay = (lambda s: np.maximum(s - k[0], 0.0)) if c["is_call"] else (lambda s: np.maximum(k[0] - s, 0.0))
        a, b = pay(S), pay(A)
        want = [n, a.sum(), b.sum(), (a * a).sum(), (b * b).sum(), (a * b).sum(), (0.5 * (S + A)).sum()]
        np.testing.assert_allclose(rows[0, :7], want, rtol=2e-9, atol=1e-9)


def test_cells_many_small_and_one_large(H, L):
    """1000 cells of 2000 paths (the hedging-backtest shape) and a single 3M-path cell go through the same code."""
    p = O.Params(kappa=0.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.0)
    cells = [dict(params=p, S0=2500.0, T=0.25, n_steps=63, n_paths=2000, seed=42 + i) for i in range(1000)]
    got = H.price_cells(cells, np.full(1000, 2500.0), L.ANTITHETIC)
    for i in (0, 1, 499, 999):
        want = H.price_european(p, 2500.0, 0.25, 63, 2000, 42 + i, [2500.0], True, L.ANTITHETIC)
        np.testing.assert_allclose(got[i, 0, :9], want[0, :9], rtol=1e-12)
    price = math.exp(-p.r * 0.25) * 0.5 * (got[:, 0, 1] + got[:, 0, 2]).sum() / (1000 * 2000)
    assert price == pytest.approx(O.bs_price(2500.0, 2500.0, 0.25, p.r, p.q, 0.3, True), rel=3e-3)
    big = [dict(params=p, S0=2500.0, T=1.0, n_steps=250, n_paths=3_000_000, seed=9)]
    np.testing.assert_allclose(H.price_cells(big, [2500.0], 0)[0, 0, :9],
                               H.price_european(p, 2500.0, 1.0, 250, 3_000_000, 9, [2500.0], True, 0)[0, :9], rtol=1e-11)


def test_cells_errors(H, L):
    p = O.Params(kappa=0.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.0)
    ok = dict(params=p, S0=100.0, T=1.0, n_steps=10, n_paths=10, seed=1)
    with pytest.raises(L.B200MCError):
        H.price_cells([], np.zeros((0, 1)))
    with pytest.raises(L.B200MCError, match="Greek"):
        H.price_cells([ok], [100.0], L.GREEKS)
    for bad in (dict(ok, n_paths=0), dict(ok, n_steps=0), dict(ok, T=-1.0)):
        with pytest.raises(L.B200MCError):
            H.price_cells([ok, bad], [100.0, 100.0])
    with pytest.raises(L.B200MCError):
        H.price_cells([ok], np.zeros((1, 257)))
    # the handle still works after the rejected calls
    assert H.price_cells([ok], [100.0])[0, 0, 0] == 10


def test_cells_edge_shapes(H, L):
    """256 strikes in one cell, a single path, a single step, a 4096-step deterministic-variance cell (the longest weight
    table), and the structured-array front end with broadcasting."""
    p = O.Params(kappa=0.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.0)
    ks = np.linspace(0.5, 1.5, 256) * 100.0
    got = H.price_cells([dict(params=p, S0=100.0, T=1.0, n_steps=12, n_paths=1000, seed=5)], ks[None, :], L.ANTITHETIC)
    want = H.price_european(p, 100.0, 1.0, 12, 1000, 5, ks, True, L.ANTITHETIC)
    np.testing.assert_allclose(got[0, :, :9], want[:, :9], rtol=1e-12, atol=1e-9)
    for n_paths, n_steps in ((1, 1), (1, 300), (300, 1), (257, 9)):
        c = dict(params=p, S0=100.0, T=0.5, n_steps=n_steps, n_paths=n_paths, seed=9, path_offset=7, is_call=False)
        np.testing.assert_allclose(H.price_cells([c], [110.0], L.FP64)[0, 0, :9],
                                   H.price_european(p, 100.0, 0.5, n_steps, n_paths, 9, [110.0], False, L.FP64, None, path_offset=7)[0, :9],
                                   rtol=1e-12, atol=1e-12)
    dv = O.Params(kappa=2.0, theta=0.05, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.03, q=0.0)
    cells = [dict(params=dv, S0=100.0, T=2.0, n_steps=4096, n_paths=600, seed=1),
             dict(params=dv, S0=100.0, T=1.0, n_steps=17, n_paths=600, seed=2)]          # two table lengths in one group
    got = H.price_cells(cells, [100.0, 95.0], 0)
    for c, K, row in zip(cells, (100.0, 95.0), got[:, 0]):
        np.testing.assert_allclose(row[:9], H.price_european(dv, 100.0, c["T"], c["n_steps"], 600, c["seed"], [K], True, 0)[0, :9],
                                   rtol=1e-12)
    arr = L.make_cells(p, [100.0, 101.0, 102.0], 1.0, 12, 500, 5, 0, [True, False, True])
    assert arr.shape == (3,) and arr["is_call"].tolist() == [1, 0, 1] and arr["S0"].tolist() == [100.0, 101.0, 102.0]
    got = H.price_cells(arr, [100.0, 100.0, 100.0])
    np.testing.assert_allclose(got[1, 0, :9], H.price_european(p, 101.0, 1.0, 12, 500, 5, [100.0], False, 0)[0, :9], rtol=1e-12)


def test_population_and_seed_batches_equal_the_scalar_calls(H, L):
    """MonteCarloEngine.price_population (a DE generation in one launch) == price_batch per candidate, and
    prices_for_seeds (the hedging premiums) == price() per seed."""
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    ks = np.array([21500.0, 22500.0, 23500.0, 24000.0])
    cols = dict(kappa=np.array([3.0, 1.0, 5.0]), theta=np.array([0.04, 0.02, 0.09]), xi=np.array([0.5, 0.6, 0.2]),
                rho=np.array([-0.7, -0.3, -0.1]), v0=np.array([0.04, 0.05, 0.01]), lambda_j=0.0, mu_j=0.0, sigma_j=0.01,
                r=0.065, q=0.012)
    for is_call in (True, False):
        eng = MonteCarloEngine(SVJParams(), num_paths=20_000, num_steps=50, handle=H)
        got = eng.price_population(cols, 22500.0, ks, 0.08, is_call)
        assert got.shape == (3, 4)
        for s_ in range(3):
            p = SVJParams(kappa=cols["kappa"][s_], theta=cols["theta"][s_], xi=cols["xi"][s_], rho=cols["rho"][s_],
                          v0=cols["v0"][s_], lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.012)
            want = [r_["price"] for r_ in MonteCarloEngine(p, num_paths=20_000, num_steps=50, handle=H).price_batch(22500.0, ks, 0.08, is_call)]
            np.testing.assert_allclose(got[s_], want, rtol=1e-10, atol=1e-9)
    eng = MonteCarloEngine(SVJParams(), num_paths=5_000, handle=H)
    seeds = [42, 43, 1000, 2 ** 40]
    got = eng.prices_for_seeds(22500.0, 22000.0, 0.25, False, seeds)
    want = [MonteCarloEngine(SVJParams(), num_paths=5_000, seed=s_, handle=H).price(22500.0, 22000.0, 0.25, False)["price"] for s_ in seeds]
    np.testing.assert_allclose(got, want, rtol=1e-12)
