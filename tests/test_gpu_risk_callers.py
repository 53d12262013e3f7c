"""GPU parity of the callers of engine/risk.py: StressTestEngine and HedgingBacktest on the
CUDA path against (a) the outputs of the reference itself (tests/golden/risk_callers_golden.json, reproduced with
rng="reference": the reference's own host draws, recurrence and walk on the GPU) and (b) the oracle on the device's
draws (rng="philox")."""
import math

import numpy as np
import pytest

from conftest import assert_tree_close
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


# ---------------------------------------------------------------------------------------------- the hedging walk
@pytest.mark.parametrize("case", range(8))
def test_hedge_walk_given_normals_vs_oracle(H, L, case):
    g = np.random.default_rng(300 + case)
    p = O.Params(kappa=1.0, theta=0.04, xi=0.3, rho=-0.5, v0=float(g.uniform(0.005, 0.4)), lambda_j=0.5, mu_j=-0.05,
                 sigma_j=0.1, r=float(g.uniform(0.0, 0.1)), q=float(g.uniform(0.0, 0.05)))
    n, days = int(g.integers(1, 700)), int(g.integers(1, 90))
    S0 = float(g.uniform(10.0, 30000.0))
    K = S0 * float(g.uniform(0.8, 1.2))
    T = float(g.uniform(0.01, 1.0))
    call = bool(g.integers(0, 2))
    txn, slip = float(g.uniform(0.0, 10.0)), float(g.uniform(0.0, 5.0))
    prem = g.uniform(0.0, 0.1 * S0, size=n)
    Z = g.standard_normal((n, days))
    pnl, cost = H.hedge_walk(p, S0, K, T, call, days, n, txn + slip, prem, Z)
    want_pnl, want_cost = O.hedge_walk(p, S0, K, T, call, days, txn, slip, prem, Z)
    # cash flows are O(S0); the P&L is their small residual: tolerance relative to the flows
    np.testing.assert_allclose(pnl, want_pnl, rtol=1e-10, atol=1e-10 * S0)
    np.testing.assert_allclose(cost, want_cost, rtol=1e-11, atol=1e-13 * S0)


def test_hedge_walk_philox_draws_and_sharding(H, L):
    """Z = None: the walk consumes exactly the draws b200mc_dump_normals(STREAM_HEDGE) exports, keyed by the GLOBAL
    scenario index -- so a scenario range can be split across ranks."""
    p = O.Params(kappa=3.0, theta=0.04, xi=0.5, rho=-0.7, v0=0.04, lambda_j=1.0, mu_j=-0.05, sigma_j=0.1, r=0.065, q=0.012)
    n, days, seed = 1000, 63, 42
    fair = O.bs_price(22500.0, 22500.0, 0.25, p.r, p.q, 0.2, True)
    prem = np.full(n, fair)
    pnl, cost = H.hedge_walk(p, 22500.0, 22500.0, 0.25, True, days, n, 7.0, prem, None, seed=seed)
    Z = H.dump_normals(seed, n, days, L.STREAM_HEDGE, L.Z1)
    assert abs(Z.mean()) < 4 / math.sqrt(n * days) and abs(Z.std() - 1) < 0.01
    pnl2, cost2 = H.hedge_walk(p, 22500.0, 22500.0, 0.25, True, days, n, 7.0, prem, Z)
    np.testing.assert_array_equal(pnl, pnl2)
    np.testing.assert_array_equal(cost, cost2)
    want = O.hedge_walk(p, 22500.0, 22500.0, 0.25, True, days, 5.0, 2.0, prem, Z)
    np.testing.assert_allclose(pnl, want[0], rtol=1e-10, atol=1e-6)
    a = H.hedge_walk(p, 22500.0, 22500.0, 0.25, True, days, 400, 7.0, prem[:400], None, seed=seed)
    b = H.hedge_walk(p, 22500.0, 22500.0, 0.25, True, days, 600, 7.0, prem[400:], None, seed=seed, scenario_offset=400)
    np.testing.assert_array_equal(np.concatenate([a[0], b[0]]), pnl)
    # a delta-hedged short option sold at its Black-Scholes value: the P&L is small against the premium (the reference's
    # walk pays no interest on the cash account, risk.py:289, so it is not centred at minus the costs)
    assert abs(pnl.mean()) < 0.3 * fair and pnl.std() < 0.5 * fair


def test_hedge_walk_edge_shapes(H, L):
    """One day, one scenario, zero volatility, zero costs, no premiums."""
    p = O.Params(kappa=3.0, theta=0.04, xi=0.5, rho=-0.7, v0=0.04, lambda_j=1.0, mu_j=-0.05, sigma_j=0.1, r=0.065, q=0.012)
    for n, days in ((1, 1), (1, 40), (33, 1)):
        Z = np.random.default_rng(n + days).standard_normal((n, days))
        pnl, cost = H.hedge_walk(p, 100.0, 105.0, 0.1, False, days, n, 0.0, None, Z)
        want = O.hedge_walk(p, 100.0, 105.0, 0.1, False, days, 0.0, 0.0, np.zeros(n), Z)
        np.testing.assert_allclose(pnl, want[0], rtol=1e-10, atol=1e-9)
        assert not cost.any()
    flat = O.Params(kappa=0.0, theta=0.0, xi=0.0, rho=0.0, v0=0.0, lambda_j=0.0, mu_j=0.0, sigma_j=0.0, r=0.05, q=0.0)
    pnl, _ = H.hedge_walk(flat, 100.0, 90.0, 1.0, True, 10, 3, 7.0, [12.0, 12.0, 12.0], None, seed=1)
    assert np.isnan(pnl).all() or np.isfinite(pnl).all()           # sigma = 0: d1 = +-inf or nan, as in the reference's formula


def test_hedge_walk_errors(H, L):
    p = O.Params(kappa=3.0, theta=0.04, xi=0.5, rho=-0.7, v0=0.04, lambda_j=1.0, mu_j=-0.05, sigma_j=0.1, r=0.065, q=0.012)
    for kw in (dict(n_days=0), dict(n_scenarios=0), dict(T=0.0), dict(T=float("nan"))):
        a = dict(params=p, S0=100.0, strike=100.0, T=0.25, is_call=True, n_days=5, n_scenarios=4, cost_bps=7.0)
        a.update(kw)
        with pytest.raises(L.B200MCError):
            H.hedge_walk(**a)
    with pytest.raises(L.B200MCError):
        H.hedge_walk(p, 100.0, 100.0, 0.25, True, 5, 4, 7.0, premiums=np.zeros(3))
    with pytest.raises(L.B200MCError):
        H.hedge_walk(p, 100.0, 100.0, 0.25, True, 5, 4, 7.0, Z=np.zeros((4, 6)))


# ---------------------------------------------------------------------------------------------- against the reference itself
@pytest.mark.parametrize("idx", [0, 1])
def test_hedging_backtest_reproduces_the_reference(H, risk_golden, idx):
    from monte_carlo_option_simulator_b200.risk import HedgingBacktest
    c = risk_golden["hedge"][idx]
    p = O.Params(**risk_golden["params"][c["params"]])
    got = HedgingBacktest(p, seed=c["seed"], rng="reference", handle=H).run_backtest(
        c["spot"], c["strike"], c["T"], c["is_call"], c["num_days"], c["txn_cost_bps"], c["slippage_bps"],
        c["num_scenarios"], c["num_mc_paths"])
    assert_tree_close(got, c["result"], rel=1e-9, abs_=1e-6)


@pytest.mark.parametrize("idx", [0, 1])
def test_stress_report_reproduces_the_reference(H, risk_golden, idx):
    from monte_carlo_option_simulator_b200.risk import StressTestEngine
    c = risk_golden["stress"][idx]
    p = O.Params(**risk_golden["params"][c["params"]])
    st = StressTestEngine(p, num_paths=c["num_paths"], seed=c["seed"], rng="reference", handle=H)
    assert_tree_close(st.full_stress_report(c["spot"], c["strike"], c["T"], c["is_call"]), c["report"], rel=1e-9, abs_=1e-7)


# ---------------------------------------------------------------------------------------------- device draws
def test_stress_report_is_one_launch_and_equals_price_calls(H, L, risk_golden):
    from monte_carlo_option_simulator_b200 import MonteCarloEngine
    from monte_carlo_option_simulator_b200.risk import StressTestEngine, SPOT_SHOCKS
    p = O.Params(**risk_golden["params"]["svj_default"])
    st = StressTestEngine(p, num_paths=200_000, seed=42, handle=H)
    before = H.launches
    rep = st.full_stress_report(22500.0, 22500.0, 0.25, True)
    assert H.launches - before == 2          # all 11 cells are SVJ-mode: one cell kernel + its fold
    eng = MonteCarloEngine(p, num_paths=200_000, seed=42, handle=H)
    base = eng.price(22500.0, 22500.0, 0.25, True)["price"]
    assert rep["jump_scenario"]["base_price"] == pytest.approx(base, rel=1e-12)
    for sh, row in zip(SPOT_SHOCKS, rep["spot_shocks"]):
        pr = eng.price(22500.0 * (1 + sh), 22500.0, 0.25, True)["price"]
        assert row["price"] == pytest.approx(pr, rel=1e-12) and row["pnl"] == pytest.approx(pr - base, rel=1e-9, abs=1e-9)
    for row in rep["vol_shocks"]:
        sp = O.vol_shocked_params(p, row["vol_shock"] / 100)
        pr = MonteCarloEngine(sp, num_paths=200_000, seed=42, handle=H).price(22500.0, 22500.0, 0.25, True)["price"]
        assert row["price"] == pytest.approx(pr, rel=1e-12) and row["v0"] == sp.v0
    # monotone in spot for a call, and vol up > base > vol down
    prices = [r["price"] for r in rep["spot_shocks"]]
    assert prices == sorted(prices)
    assert rep["vol_shocks"][0]["price"] < base < rep["vol_shocks"][1]["price"]


def test_hedging_backtest_device_draws_vs_oracle_walk(H, L, risk_golden):
    """rng="philox": premiums from one (cell x path) launch, walk on Philox draws; rebuilt here from single price() calls
    and the oracle walk on the dumped draws."""
    from monte_carlo_option_simulator_b200 import MonteCarloEngine
    from monte_carlo_option_simulator_b200.risk import HedgingBacktest
    p = O.Params(**risk_golden["params"]["gbm_cfg1"])
    nsc, days = 200, 21
    got = HedgingBacktest(p, seed=11, handle=H).run_backtest(2500.0, 2500.0, 21 / 252, True, num_scenarios=nsc,
                                                             num_mc_paths=20_000)
    prem = np.array([MonteCarloEngine(p, 20_000, seed=11 + s, handle=H).price(2500.0, 2500.0, 21 / 252, True)["price"]
                     for s in (0, 1, nsc - 1)])
    Z = H.dump_normals(11, nsc, days, L.STREAM_HEDGE, L.Z1)
    # the engine's default flags include the pseudo control variate (quirk 2): every premium equals bs_ref + noise/2
    bs = O.bs_price(2500.0, 2500.0, 21 / 252, p.r, p.q, 0.3, True)
    assert np.all(np.abs(prem - bs) < 0.05 * bs)
    engine = MonteCarloEngine(p, 20_000, seed=11, handle=H)
    allprem = [r["price"] for r in engine.price_many(2500.0, 2500.0, 21 / 252, True, seeds=[11 + s for s in range(nsc)])]
    np.testing.assert_allclose([allprem[0], allprem[1], allprem[-1]], prem, rtol=1e-12)
    pnl, cost = O.hedge_walk(p, 2500.0, 2500.0, 21 / 252, True, days, 5.0, 2.0, allprem, Z)
    assert got["mean_pnl"] == pytest.approx(pnl.mean(), rel=1e-9, abs=1e-9)
    assert got["std_pnl"] == pytest.approx(pnl.std(), rel=1e-9)
    assert got["total_txn_cost_avg"] == pytest.approx(cost[-1], rel=1e-10)
    want = O.risk_metrics(pnl, 0.99)
    for k, w in want.items():
        if math.isnan(w):
            assert math.isnan(got["risk_metrics"][k])
        else:
            assert got["risk_metrics"][k] == pytest.approx(w, rel=1e-8, abs=1e-8), k
    for q in (1, 50, 99):
        assert got["pnl_percentiles"][f"{q}%"] == pytest.approx(np.percentile(pnl, q), rel=1e-9, abs=1e-9)
