"""CPU: the host-side mirror (dict assembly from the fused sums, steps rule, Sobol/bridge front end, sharding,
patch_reference) against the oracle and the golden fixtures.  The CUDA library is replaced here by a NumPy stand-in
that produces b200mc_sums from the ORACLE's terminal spots, so that the algebra from sums to result dictionaries
(including the reference's pseudo control variate, quirk 2) is checked against MonteCarloOracle / GreeksOracle."""
import math
import os
import sys
import types

import numpy as np
import pytest

from oracle import oracle as O
from monte_carlo_option_simulator_b200 import _lib, monte_carlo as MC
from monte_carlo_option_simulator_b200 import GreeksEngine, MonteCarloEngine, SVJParams
from monte_carlo_option_simulator_b200.dist import Comm, shard_range, sharded_sums


def P(golden, name):
    return O.Params(**golden["params"][name])


class OracleBackedHandle:
    """Stands in for _lib.Handle: same price_european signature, sums computed with NumPy from oracle paths drawn
    with PCG64 (so results can be compared with the oracle's own reductions on the same draws)."""

    def __init__(self, n_global=None):
        self.calls = 0
        self.n_global = n_global       # total paths of the job: PCG64 streams depend on it, Philox counters do not

    def price_european(self, params, S0, T, n_steps, n_paths, seed, strikes, is_call=True, flags=0, bumps=None,
                       path_offset=0, out_dev=None):
        self.calls += 1
        p = O.as_params(params)
        Z1, Z2, Zj, Zjs = O.draw_pcg64(seed, self.n_global or (path_offset + n_paths), n_steps)
        Z1, Z2, Zj, Zjs = (z[path_offset:path_offset + n_paths] for z in (Z1, Z2, Zj, Zjs))
        S = O._sim(p, S0, T, Z1, Z2, Zj, Zjs, n_steps)[0]
        A = O._sim(p, S0, T, -Z1, -Z2, Zj, -Zjs, n_steps)[0] if flags & _lib.ANTITHETIC else None
        rows = []
        for K in np.atleast_1d(strikes):
            pay = (lambda s: np.maximum(s - K, 0.0)) if is_call else (lambda s: np.maximum(K - s, 0.0))
            a = pay(S)
            b = pay(A) if A is not None else np.zeros_like(a)
            s_avg = 0.5 * (S + A) if A is not None else S
            pc = 0.5 * (a + b) if A is not None else a
            row = [n_paths, a.sum(), b.sum(), (a * a).sum(), (b * b).sum(), (a * b).sum(), s_avg.sum(),
                   (s_avg ** 2).sum(), (pc * s_avg).sum()] + [0.0] * 8
            if flags & _lib.GREEKS:
                itm = (S > K) if is_call else (S < K)
                sb = bumps.spot_bump
                row[9] = (itm * S / S0).sum()
                row[10] = pay(S * (1 + sb)).sum()
                row[11] = pay(S * (1 - sb)).sum()
                row[12] = pay(O._sim(p, S0, T, Z1, Z2, Zj, Zjs, n_steps, v0=bumps.v0_up)[0]).sum()
                row[13] = pay(O._sim(p, S0, T, Z1, Z2, Zj, Zjs, n_steps, v0=bumps.v0_dn)[0]).sum()
                row[14] = pay(S * math.exp((bumps.r_up - p.r) * T)).sum()
                row[15] = pay(S * math.exp((bumps.r_dn - p.r) * T)).sum()
            rows.append(row)
        return np.array(rows, dtype=np.float64)

    def price_cells(self, cells, strikes, flags=0, out_dev=None):
        """Every cell through price_european above (cells: the structured array of _lib.make_cells)."""
        ks = np.asarray(strikes, dtype=np.float64).reshape(len(cells), -1)
        out = []
        for c, k in zip(cells, ks):
            p = O.Params(**{f: float(c[f]) for f in _lib.PARAM_FIELDS})
            out.append(self.price_european(p, float(c["S0"]), float(c["T"]), int(c["n_steps"]), int(c["n_paths"]),
                                           int(c["seed"]), k, bool(c["is_call"]), flags, None, int(c["path_offset"])))
        return np.array(out)

    def hedge_walk(self, params, S0, strike, T, is_call, n_days, n_scenarios, cost_bps, premiums=None, Z=None, seed=0,
                   scenario_offset=0):
        if Z is None:       # stand-in for the device's Philox draws
            Z = np.random.default_rng(seed + 12345).standard_normal((scenario_offset + n_scenarios, n_days))[scenario_offset:]
        return O.hedge_walk(params, S0, strike, T, is_call, n_days, cost_bps, 0.0,
                            premiums if premiums is not None else np.zeros(n_scenarios), Z)

    def risk_metrics(self, pnl, confidence=0.99, n=None, dtype=None):
        r = O.risk_metrics(np.asarray(pnl, dtype=np.float64), confidence)
        return np.array([r[k] for k in ("var", "cvar", "skewness", "kurtosis", "excess_kurtosis", "tail_index", "mean", "std")])


def test_bs_closed_forms_match_golden(golden):
    b = golden["cases"]["bs"]
    assert MC.bs_price(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, True) == pytest.approx(b["cfg1_call"], rel=1e-13)
    assert MC.bs_price(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, False) == pytest.approx(b["cfg1_put"], rel=1e-13)
    assert MC.bs_delta(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, True) == pytest.approx(b["cfg1_delta_call"], rel=1e-13)
    assert MC.bs_delta(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, False) == pytest.approx(b["cfg1_delta_put"], rel=1e-13)
    assert MC.bs_price(110.0, 100.0, 0.0, 0.05, 0.0, 0.2, True) == b["expired_call"]
    assert MC.bs_delta(90.0, 100.0, 0.0, 0.05, 0.0, 0.2, False) == b["expired_put_delta"]
    assert MC.bs_price(22500.0, 22500.0, 0.04, 0.065, 0.012, 0.2, True) == pytest.approx(b["verify_py"], rel=1e-13)


def test_steps_rule():
    assert MC.steps_for(252, 0.04) == 10 and MC.steps_for(250, 1.0) == 250 and MC.steps_for(252, 0.25) == 63
    assert MC.steps_for(252, 0.1, floor=50) == 50 and MC.steps_for(250, 2.0) == 500


def test_reference_front_end_sobol_and_bridge(golden, garr):
    np.testing.assert_array_equal(MC.generate_sobol_normals(100, 12, seed=3), garr["sobol_100x12_seed3"])
    for k, want in golden["cases"]["bb_order"].items():
        assert MC._bb_ordering(int(k)) == want
    np.testing.assert_allclose(MC.brownian_bridge_reorder(garr["bb_in"], 31), garr["bb_out"], rtol=1e-13, atol=1e-16)
    Z = MC._reference_draws(42, 64, 20, use_sobol=False)
    Zo = O.draw_pcg64(42, 64, 20)
    for a, b in zip(Z, Zo):
        np.testing.assert_array_equal(a, b)
    Z = MC._reference_draws(5, 64, 12, use_sobol=True)
    Zo = O.draw_sobol(5, 64, 12)
    for a, b in zip(Z, Zo):
        np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 10, 63, 64, 250])
def test_reference_bridge_as_a_node_table(n):
    """The table handed to the device (b200mc_bridge_node) reproduces the reference's brownian_bridge_reorder bit for bit,
    degeneracy included: the first placement is the endpoint with sd == 0."""
    g = np.random.default_rng(n)
    z = g.standard_normal((6, n))
    nodes = MC.reference_bridge_nodes(n)
    W = np.zeros((6, n + 1))
    for nd in nodes:
        W[:, nd["t"]] = W[:, nd["l"]] + ((W[:, nd["r"]] - W[:, nd["l"]]) * nd["a"]) / nd["b"] + nd["sd"] * z[:, nd["dim"]]
    np.testing.assert_array_equal(np.diff(W, axis=1), MC.brownian_bridge_reorder(z, n))
    np.testing.assert_array_equal(np.diff(W, axis=1), O.bb_reorder(z, n))
    assert nodes[0]["t"] == n and nodes[0]["sd"] == 0.0 and sorted(nodes["t"].tolist()) == list(range(1, n + 1))
    assert nodes.dtype.itemsize == 40


@pytest.mark.parametrize("anti", [False, True])
@pytest.mark.parametrize("cv", [False, True])
@pytest.mark.parametrize("is_call", [True, False])
@pytest.mark.parametrize("pname", ["svj_default", "gbm_cfg1"])
def test_price_dict_from_sums_equals_oracle(golden, anti, cv, is_call, pname):
    p = P(golden, pname)
    n, steps, seed = 3000, 252, 11
    spot, K, T = 22500.0, 22000.0, 0.25
    eng = MonteCarloEngine(p, n, steps, seed, use_sobol=False, use_antithetic=anti, use_control_variate=cv,
                           rng="philox", handle=OracleBackedHandle())
    got = eng.price(spot, K, T, is_call)
    want = O.MonteCarloOracle(p, n, steps, seed, False, anti, cv).price(spot, K, T, is_call)
    assert set(want) <= set(got)
    for k, w in want.items():
        assert got[k] == pytest.approx(w, rel=1e-9, abs=2e-7), k        # abs: the degenerate SE == 0 case (quirk 2)
    assert {"price_cv_spot", "std_error_cv_spot"} <= set(got)


def test_price_batch_from_sums_equals_oracle(golden):
    p = P(golden, "svj_default")
    ks = np.array([21000.0, 22500.0, 24000.0])
    for anti, cv in [(True, True), (False, True), (True, False)]:
        eng = MonteCarloEngine(p, 2048, 252, 7, use_sobol=False, use_antithetic=anti, use_control_variate=cv,
                               rng="philox", handle=OracleBackedHandle())
        got = eng.price_batch(22500.0, ks, 0.25, True)
        want = O.MonteCarloOracle(p, 2048, 252, 7, False, anti, cv).price_batch(22500.0, ks, 0.25, True)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert set(g) == set(w)
            for k in w:
                assert g[k] == pytest.approx(w[k], rel=1e-9, abs=1e-9)


@pytest.mark.parametrize("is_call", [True, False])
def test_greeks_from_sums_equal_oracle(golden, is_call):
    p = P(golden, "svj_default")
    n, steps, seed = 2048, 252, 42
    h = OracleBackedHandle()
    g = GreeksEngine(p, n, steps, seed, rng="philox", handle=h)
    o = O.GreeksOracle(p, n, steps, seed)
    args = (22500.0, 22500.0, 0.25, is_call)
    for name in ("delta", "vega", "gamma"):
        got, want = getattr(g, name)(*args), getattr(o, name)(*args)
        for k, w in want.items():
            assert got[k] == pytest.approx(w, rel=1e-8, abs=1e-10), (name, k)
    assert h.calls == 1          # ONE fused launch served delta, vega and gamma (the reference runs 9 simulations)


def test_int_spot_is_cast_to_float(golden):
    """Documented divergence from the reference (quirk 3): an int spot must not truncate the paths."""
    p = P(golden, "gbm_cfg1")
    eng = MonteCarloEngine(p, 500, 50, 3, use_sobol=False, rng="philox", handle=OracleBackedHandle())
    a, b = eng.price(2500, 2500, 1.0), eng.price(2500.0, 2500.0, 1.0)
    assert a["price"] == b["price"]


def test_shard_range_partitions_the_paths():
    for n, w in [(10, 3), (7, 8), (10_000_000, 8), (1, 2), (0, 4)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_sharded_sums_single_rank_is_identity(golden):
    p = P(golden, "gbm_cfg1")
    h = OracleBackedHandle()
    whole = h.price_european(p, 100.0, 0.5, 20, 600, 9, [95.0, 105.0], True, _lib.ANTITHETIC)
    got = sharded_sums(h, Comm(), p, 100.0, 0.5, 20, 600, 9, [95.0, 105.0], True, _lib.ANTITHETIC)
    np.testing.assert_array_equal(whole, got)


def test_patch_reference_rebinds_import_by_name_sites():
    from monte_carlo_option_simulator_b200 import patch_reference
    pkg = types.ModuleType("fakeengine")
    pkg.__path__ = []
    mods = {}
    for name, attrs in {"monte_carlo": ["_simulate_svj_paths_numba", "MonteCarloEngine", "bs_price", "bs_delta"],
                        "greeks": ["_simulate_svj_paths_numba", "MonteCarloEngine", "GreeksEngine"],
                        "risk": ["MonteCarloEngine", "compute_risk_metrics", "StressTestEngine", "HedgingBacktest",
                                 "LiquidityStress"],
                        "calibration": ["MonteCarloEngine"],
                        "app": ["MonteCarloEngine", "GreeksEngine", "StressTestEngine", "HedgingBacktest"]}.items():
        m = types.ModuleType(f"fakeengine.{name}")
        for a in attrs:
            setattr(m, a, object())
        mods[name] = m
        sys.modules[f"fakeengine.{name}"] = m
    sys.modules["fakeengine"] = pkg
    try:
        from monte_carlo_option_simulator_b200 import risk as R
        with pytest.raises(ValueError):
            patch_reference("fakeengine", rng="mt19937")
        old = os.environ.get("B200MC_RNG")
        try:
            patch_reference("fakeengine", batch_scenarios=False, rng="reference")
            assert MonteCarloEngine(SVJParams()).rng == "reference"
        finally:
            if old is None:
                os.environ.pop("B200MC_RNG", None)
            else:
                os.environ["B200MC_RNG"] = old
        done = patch_reference("fakeengine", batch_scenarios=False)
        assert not isinstance(mods["app"].StressTestEngine, type)       # caller classes: left alone on request
        assert not isinstance(mods["risk"].HedgingBacktest, type)
        assert len(done) == 12          # (no calibration objectives in this stub module)
        done = patch_reference("fakeengine")
        assert mods["risk"].StressTestEngine is R.StressTestEngine and mods["app"].HedgingBacktest is R.HedgingBacktest
        assert mods["risk"].LiquidityStress is R.LiquidityStress
        assert len(done) == 17
        assert mods["monte_carlo"].MonteCarloEngine is MonteCarloEngine
        assert mods["greeks"]._simulate_svj_paths_numba is MC._simulate_svj_paths_numba
        assert mods["greeks"].GreeksEngine is GreeksEngine
        assert mods["risk"].MonteCarloEngine is MonteCarloEngine
        assert mods["calibration"].MonteCarloEngine is MonteCarloEngine
        assert mods["app"].GreeksEngine is GreeksEngine
    finally:
        for k in list(sys.modules):
            if k.startswith("fakeengine"):
                del sys.modules[k]


def test_engine_rejects_unknown_modes():
    with pytest.raises(ValueError):
        MonteCarloEngine(SVJParams(), rng="mt19937")
    with pytest.raises(ValueError):
        MonteCarloEngine(SVJParams(), precision="fp16")


class NumpyRiskHandle:
    """NumPy stand-in for the three multi-rank tail-metric primitives of libb200mc (csrc/risk.cu)."""

    def risk_begin(self, pnl, n=None, dtype=None):
        self.x = np.asarray(pnl, dtype=np.float64).ravel()
        bits = self.x.view(np.uint64)
        neg = (bits >> np.uint64(63)).astype(bool)
        self.keys = np.where(neg, ~bits, bits | np.uint64(1 << 63))
        return np.array([self.x.sum(), float((self.x < 0).sum())])

    def risk_hist(self, radix_pass, nsel, prefix):
        hist = np.zeros((2, 256), dtype=np.uint64)
        shift = np.uint64(8 * radix_pass)
        digit = ((self.keys >> shift) & np.uint64(255)).astype(np.int64)
        for s in range(nsel):
            if radix_pass == 7:
                sel = np.ones(self.keys.size, dtype=bool)
            else:
                hi = np.uint64(8 * radix_pass + 8)
                sel = (self.keys >> hi) == (np.uint64(prefix[s]) >> hi)
            hist[s] = np.bincount(digit[sel], minlength=256).astype(np.uint64)
        return hist

    def risk_finish(self, mean, nsel, thr):
        d = self.x - mean
        lt = self.x < thr[0]
        tail = np.log(self.x[self.x < thr[1]] / thr[1]).sum() if nsel > 1 else 0.0
        return np.array([(d ** 2).sum(), (d ** 3).sum(), (d ** 4).sum(), float(lt.sum()), self.x[lt].sum(), tail])


def test_sharded_risk_host_logic_single_rank_equals_oracle(golden, garr):
    """The host side of the distributed radix select (risk.compute_risk_metrics_sharded) against the oracle."""
    from conftest import unnan
    from monte_carlo_option_simulator_b200.risk import compute_risk_metrics_sharded, _key_to_value
    assert _key_to_value(0x8000000000000000 | np.array([1.5]).view(np.uint64)[0].item()) == 1.5
    assert _key_to_value((~np.array([-2.25]).view(np.uint64)[0]).item() & 0xFFFFFFFFFFFFFFFF) == -2.25
    for c in golden["cases"]["risk"]:
        got = compute_risk_metrics_sharded(garr[f"risk_{c['name']}"], c["confidence"], comm=Comm(), handle=NumpyRiskHandle())
        want = unnan(c["result"])
        assert set(got) == set(want)
        for k, w in want.items():
            if np.isnan(w):
                assert np.isnan(got[k]), (c["name"], k)
            else:
                assert got[k] == pytest.approx(w, rel=1e-11, abs=1e-13), (c["name"], k)


def test_batched_calibration_objectives_equal_the_per_strike_loop(golden, monkeypatch):
    """the patched objectives (one price_batch launch per candidate) return what the reference's
    per-strike loop of price() calls returns (engine/calibration.py:53-135)."""
    from monte_carlo_option_simulator_b200 import patch_reference
    from monte_carlo_option_simulator_b200.models import SVJParams as OurParams
    handle = OracleBackedHandle()
    monkeypatch.setattr(_lib, "default_handle", lambda device=None: handle)
    monkeypatch.setenv("B200MC_RNG", "philox")
    REG = {"xi": 0.01, "rho": 0.005, "lambda_j": 0.01}

    def check_feller(kappa, theta, xi):
        return 2.0 * kappa * theta > xi * xi

    def ref_heston(x, spot, strikes, T, mkt, w, r, q, is_call, num_paths, num_steps):          # calibration.py:53-95
        kappa, theta, xi, rho, v0 = x
        pen = 0.0 if check_feller(kappa, theta, xi) else 10.0 * (xi ** 2 - 2 * kappa * theta) ** 2
        eng = MonteCarloEngine(OurParams(kappa=kappa, theta=theta, xi=xi, rho=rho, v0=v0, lambda_j=0.0, mu_j=0.0,
                                         sigma_j=0.01, r=r, q=q), num_paths=num_paths, num_steps=num_steps)
        tot = sum(w[i] * (eng.price(spot, K, T, is_call=is_call)["price"] - mkt[i]) ** 2 for i, K in enumerate(strikes))
        return tot + REG["xi"] * xi ** 2 + REG["rho"] * rho ** 2 + pen

    calib = types.ModuleType("fakecal.calibration")
    calib.MonteCarloEngine, calib.REGULARIZATION, calib.check_feller, calib.SVJParams = object(), REG, check_feller, OurParams
    calib._heston_objective = calib._svj_objective = None
    pkg = types.ModuleType("fakecal")
    pkg.__path__ = []
    sys.modules["fakecal"], sys.modules["fakecal.calibration"] = pkg, calib
    try:
        done = patch_reference("fakecal")
        assert "fakecal.calibration._heston_objective" in done and "fakecal.calibration._svj_objective" in done
        ks, mkt, w = np.array([21500.0, 22500.0, 23500.0]), np.array([1500.0, 900.0, 480.0]), np.array([0.3, 0.4, 0.3])
        for x in ([3.0, 0.04, 0.5, -0.7, 0.04], [1.0, 0.02, 0.6, -0.3, 0.05]):       # Feller satisfied / violated
            got = calib._heston_objective(np.array(x), 22500.0, ks, 0.25, mkt, w, 0.065, 0.012, True, num_paths=600, num_steps=80)
            want = ref_heston(x, 22500.0, ks, 0.25, mkt, w, 0.065, 0.012, True, 600, 80)
            assert got == pytest.approx(want, rel=1e-10)
        calls = handle.calls
        calib._svj_objective(np.array([1.0, -0.05, 0.1]), np.array([3.0, 0.04, 0.5, -0.7, 0.04]), 22500.0, ks, 0.25, mkt, w,
                             0.065, 0.012, True, num_paths=600, num_steps=80)
        assert handle.calls == calls + 1                   # ONE launch for the three strikes
        # population forms (8f-2, second half): column s of x == the scalar objective at x[:, s]
        X = np.array([[3.0, 0.04, 0.5, -0.7, 0.04], [1.0, 0.02, 0.6, -0.3, 0.05], [5.0, 0.09, 0.2, -0.1, 0.01]]).T
        got = calib._heston_objective.population(X, 22500.0, ks, 0.25, mkt, w, 0.065, 0.012, True, 600, 80)
        want = [calib._heston_objective(X[:, s], 22500.0, ks, 0.25, mkt, w, 0.065, 0.012, True, 600, 80) for s in range(3)]
        np.testing.assert_allclose(got, want, rtol=1e-10)
        XJ = np.array([[1.0, -0.05, 0.1], [4.0, 0.02, 0.3]]).T
        hp = np.array([3.0, 0.04, 0.5, -0.7, 0.04])
        got = calib._svj_objective.population(XJ, hp, 22500.0, ks, 0.25, mkt, w, 0.065, 0.012, False, 600, 80)
        want = [calib._svj_objective(XJ[:, s], hp, 22500.0, ks, 0.25, mkt, w, 0.065, 0.012, False, 600, 80) for s in range(2)]
        np.testing.assert_allclose(got, want, rtol=1e-10)
        # the DE wrapper: objectives with a population form run vectorised / deferred, anything else goes through
        seen = {}

        def fake_de(func, bounds, args=(), **kw):
            seen.update(kw, func=func)
            return "result"
        calib.differential_evolution = fake_de
        done = patch_reference("fakecal", batch_population=True)
        assert "fakecal.calibration.differential_evolution" in done
        assert calib.differential_evolution(calib._heston_objective, [(0, 1)] * 5, args=(1,), workers=1, seed=42) == "result"
        assert seen["vectorized"] is True and seen["updating"] == "deferred" and "workers" not in seen and seen["seed"] == 42
        assert seen["func"] is calib._heston_objective.population
        seen.clear()
        calib.differential_evolution(len, [(0, 1)], workers=1)
        assert seen["func"] is len and "vectorized" not in seen and seen["workers"] == 1
    finally:
        for k in list(sys.modules):
            if k.startswith("fakecal"):
                del sys.modules[k]


# ---------------------------------------------------------------------------------------------- a11: risk.py callers
def test_price_many_equals_a_loop_of_price(golden):
    """price_many builds the cells (params, spot, T -> steps, seed per problem) and the reference's dicts from the sums:
    same dicts as separate engines' price() calls."""
    p, p2 = P(golden, "svj_default"), P(golden, "gbm_cfg1")
    h = OracleBackedHandle()
    eng = MonteCarloEngine(p, 700, 252, 5, use_sobol=False, handle=h, rng="philox")
    spots, ks, Ts = [22500.0, 21000.0, 22500.0], [22500.0, 22000.0, 23000.0], [0.25, 0.1, 0.02]
    got = eng.price_many(spots, ks, Ts, [True, False, True], params=[p, p2, p], seeds=[5, 6, 7])
    assert h.calls == 3
    for i, g in enumerate(got):
        e = MonteCarloEngine([p, p2, p][i], 700, 252, [5, 6, 7][i], use_sobol=False, handle=OracleBackedHandle(), rng="philox")
        want = e.price(spots[i], ks[i], Ts[i], [True, False, True][i])
        assert g == pytest.approx(want, rel=1e-12, abs=1e-12)
    # scalars broadcast, and a mismatch in lengths is an error
    assert len(eng.price_many(22500.0, [22000.0, 23000.0], 0.25)) == 2
    with pytest.raises(ValueError):
        eng.price_many([1.0, 2.0], [1.0, 2.0, 3.0], 0.25)


@pytest.mark.parametrize("idx", [0, 1])
def test_stress_engine_keys_and_formulas(risk_golden, idx):
    """StressTestEngine on the stand-in handle: same structure and keys as the reference's report, every number equal to
    the reference's formulas applied to this engine's own price() results."""
    from monte_carlo_option_simulator_b200.risk import StressTestEngine, SPOT_SHOCKS, VOL_SHOCKS
    from conftest import assert_tree_close
    c = risk_golden["stress"][idx]
    p = O.Params(**risk_golden["params"][c["params"]])
    h = OracleBackedHandle()
    st = StressTestEngine(p, num_paths=400, seed=c["seed"], rng="philox", handle=h)
    rep = st.full_stress_report(c["spot"], c["strike"], c["T"], c["is_call"])
    assert h.calls == 11                                  # 1 base + 6 spot + 2 vol + 2 gap cells, one price_cells call

    def price(params, spot):
        return MonteCarloEngine(params, 400, seed=c["seed"], rng="philox", handle=OracleBackedHandle()).price(
            spot, c["strike"], c["T"], c["is_call"])["price"]
    base = price(p, c["spot"])
    want = {"spot_shocks": [], "vol_shocks": [], "jump_scenario": None}
    for sh in SPOT_SHOCKS:
        pr = price(p, c["spot"] * (1 + sh))
        want["spot_shocks"].append({"shock_pct": sh * 100, "spot": c["spot"] * (1 + sh), "price": pr, "pnl": pr - base,
                                    "pnl_pct": (pr - base) / max(base, 1e-6) * 100})
    for sh in VOL_SHOCKS:
        sp = O.vol_shocked_params(p, sh)
        pr = price(sp, c["spot"])
        want["vol_shocks"].append({"vol_shock": sh * 100, "v0": sp.v0, "price": pr, "pnl": pr - base})
    dn, up = price(p, c["spot"] * 0.96), price(p, c["spot"] * 1.04)
    want["jump_scenario"] = {"base_price": base, "gap_down_price": dn, "gap_down_pnl": dn - base, "gap_up_price": up,
                             "gap_up_pnl": up - base, "gap_size_pct": 4.0}
    assert_tree_close(rep, want, rel=1e-12, abs_=1e-10)
    assert_tree_close(st.spot_shock_ladder(c["spot"], c["strike"], c["T"], c["is_call"]), want["spot_shocks"], 1e-12, 1e-10)
    assert_tree_close(st.vol_shock_ladder(c["spot"], c["strike"], c["T"], c["is_call"]), want["vol_shocks"], 1e-12, 1e-10)
    assert_tree_close(st.jump_scenario(c["spot"], c["strike"], c["T"], c["is_call"]), want["jump_scenario"], 1e-12, 1e-10)
    # same keys as the reference's own report
    assert set(rep) == set(c["report"]) and set(rep["spot_shocks"][0]) == set(c["report"]["spot_shocks"][0])
    assert set(rep["vol_shocks"][0]) == set(c["report"]["vol_shocks"][0])
    assert set(rep["jump_scenario"]) == set(c["report"]["jump_scenario"])


def test_liquidity_stress_transforms():
    from monte_carlo_option_simulator_b200.risk import LiquidityStress
    p = O.Params(kappa=3.0, theta=0.04, xi=0.5, rho=-0.7, v0=0.04, lambda_j=1.0, mu_j=-0.05, sigma_j=0.1, r=0.065, q=0.012)
    assert LiquidityStress.bid_ask_widening(2.0) == {"base_spread": 2.0, "stressed_spread": 6.0, "slippage_increase": 4.0}
    g = LiquidityStress.vol_gap_no_spot_move(p, 0.05)
    assert g.v0 == pytest.approx(0.04 + 2 * 0.2 * 0.05 + 0.0025) and g.theta == p.theta and g.rho == p.rho
    c = LiquidityStress.expiry_vol_crush(p, 0.30)
    assert c.v0 == pytest.approx(0.028) and c.theta == pytest.approx(0.034) and c.kappa == p.kappa
    assert LiquidityStress.expiry_vol_crush(p, 1.0).v0 == 0.001


def test_hedging_backtest_host_logic(risk_golden):
    """HedgingBacktest on the stand-in handle == the oracle's walk on the same premiums and normals; keys and the
    'last scenario' quirk of total_txn_cost_avg as in the reference."""
    from monte_carlo_option_simulator_b200.risk import HedgingBacktest
    c = risk_golden["hedge"][1]
    p = O.Params(**risk_golden["params"][c["params"]])
    h = OracleBackedHandle()
    bt = HedgingBacktest(p, seed=3, rng="philox", handle=h)
    got = bt.run_backtest(c["spot"], c["strike"], c["T"], c["is_call"], num_days=c["num_days"], txn_cost_bps=3.0,
                          slippage_bps=1.0, num_scenarios=15, num_mc_paths=300)
    prem = [MonteCarloEngine(p, 300, seed=3 + s, rng="philox", handle=OracleBackedHandle()).price(
        c["spot"], c["strike"], c["T"], c["is_call"])["price"] for s in range(15)]
    Z = np.random.default_rng(3 + 12345).standard_normal((15, c["num_days"]))
    pnl, cost = O.hedge_walk(p, c["spot"], c["strike"], c["T"], c["is_call"], c["num_days"], 3.0, 1.0, prem, Z)
    assert set(got) == set(c["result"])
    assert got["mean_pnl"] == pytest.approx(pnl.mean(), rel=1e-10) and got["std_pnl"] == pytest.approx(pnl.std(), rel=1e-10)
    assert got["total_txn_cost_avg"] == pytest.approx(cost[-1], rel=1e-12)
    assert got["pnl_percentiles"]["5%"] == pytest.approx(np.percentile(pnl, 5), rel=1e-10)
    assert got["risk_metrics"]["var"] == pytest.approx(O.risk_metrics(pnl)["var"], rel=1e-10)
    # default day count: max(int(T * 252), 1)
    got = bt.run_backtest(c["spot"], c["strike"], 0.003, True, num_scenarios=4, num_mc_paths=50)
    assert got["num_scenarios"] == 4


# ---------------------------------------------------------------------------------------------- 8(f)-3: surface host logic
class OracleIvHandle:
    """Stands in for _lib.Handle.implied_vol: the oracle's brentq per element, NaN for None."""

    def implied_vol(self, prices, S, strikes, maturities, r, q, is_call=True, lo=0.001, hi=5.0):
        pr, ks, ts, cl = np.broadcast_arrays(np.asarray(prices, dtype=float), np.asarray(strikes, dtype=float),
                                             np.asarray(maturities, dtype=float), np.asarray(is_call))
        out = np.full(pr.shape, np.nan)
        for idx in np.ndindex(pr.shape):
            v = O.implied_vol(pr[idx], S, ks[idx], ts[idx], r, q, bool(cl[idx]), lo, hi)
            if v is not None:
                out[idx] = v
        return out


@pytest.mark.parametrize("with_spreads", [True, False])
def test_extract_iv_surface_host_logic(iv_golden, with_spreads):
    """Broadcasting of the chain, the bid-ask filter, NaN / valid_mask conventions of surface.extract_iv_surface against the
    surface written by the reference (the inversions themselves come from the stand-in above)."""
    from monte_carlo_option_simulator_b200.surface import extract_iv_surface, implied_vol
    g = iv_golden
    tag = "" if with_spreads else "_ns"
    s_ = extract_iv_surface(float(g["spot"]), float(g["r"]), float(g["q"]), g["strikes"], g["maturities"], g["calls"], g["puts"],
                            g["spreads"] if with_spreads else None, handle=OracleIvHandle())
    np.testing.assert_array_equal(s_["valid_mask"], g["valid" + tag])
    np.testing.assert_allclose(s_["iv_call"], g["iv_call" + tag], atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(s_["iv_put"], g["iv_put" + tag], atol=1e-12, equal_nan=True)
    assert implied_vol(1e9, 100.0, 100.0, 1.0, 0.05, 0.0, handle=OracleIvHandle()) is None
    assert implied_vol(10.45, 100.0, 100.0, 1.0, 0.05, 0.0, handle=OracleIvHandle()) == pytest.approx(0.2, abs=1e-3)


def test_surface_closed_forms_match_the_oracle():
    from monte_carlo_option_simulator_b200 import surface as SF
    for S, K, T, r, q, sig in [(100.0, 110.0, 0.5, 0.03, 0.01, 0.25), (22500.0, 21000.0, 0.08, 0.065, 0.012, 0.18),
                               (100.0, 100.0, 1e-11, 0.05, 0.0, 0.2), (100.0, 90.0, 1.0, 0.05, 0.0, 1e-11)]:
        assert SF.bs_call_price(S, K, T, r, q, sig) == pytest.approx(O.bs_call_price(S, K, T, r, q, sig), rel=1e-13, abs=1e-13)
        assert SF.bs_put_price(S, K, T, r, q, sig) == pytest.approx(O.bs_put_price(S, K, T, r, q, sig), rel=1e-13, abs=1e-13)
    assert SF.bs_vega(100.0, 100.0, 1e-11, 0.05, 0.0, 0.2) == 0.0
    assert SF.bs_vega(100.0, 100.0, 1.0, 0.05, 0.0, 0.2) == pytest.approx(37.524, abs=1e-3)


def test_expired_option_returns_the_intrinsic_value_like_the_reference():
    """T == 0: the reference walks 10 steps of dt = 0 and returns the intrinsic value with zero standard error
    (engine/monte_carlo.py:287, bs_price :31-34); the mirror answers without a launch (no GPU needed)."""
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
    for rng in ("philox", "reference"):
        e = MonteCarloEngine(SVJParams(), 1000, 252, 42, rng=rng)
        r = e.price(105.0, 100.0, 0.0, True)
        assert r == {"price": 5.0, "std_error": 0.0, "num_paths_used": 1000, "num_steps": 10, "bs_cv_adjustment": 0.0,
                     "bs_ref": 5.0, "raw_mc_price": 5.0}
        assert e.price(105.0, 100.0, 0.0, False)["price"] == 0.0
        b = e.price_batch(105.0, [100.0, 110.0], 0.0, True)
        assert [x["price"] for x in b] == [5.0, 0.0] and b[0]["strike"] == 100.0 and b[0]["bs_ref"] == 5.0
    e = MonteCarloEngine(SVJParams(), 1000, 252, 42, use_control_variate=False, rng="philox")
    assert set(e.price(95.0, 100.0, 0.0, False)) == {"price", "std_error", "num_paths_used", "num_steps"}


def test_parameter_helpers_equal_the_reference():
    """jump_compensation / feller_satisfied / to_array / from_array / validate of the parameter container against
    the values the reference's dataclass gave (tests/golden/make_params_golden.py; engine/models.py:46-84)."""
    import json
    from monte_carlo_option_simulator_b200.models import SVJParams as P
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "params_golden.json"), encoding="utf-8"))
    assert len(gold) >= 10
    for name, g in gold.items():
        p = P(**g["kwargs"])
        assert p.jump_compensation == g["jump_compensation"], name
        assert p.feller_satisfied is g["feller_satisfied"], name
        arr = p.to_array()
        assert arr.dtype == np.float64 and arr.tolist() == g["to_array"], name
        assert p.validate() == g["validate"], name
        back = P.from_array(arr, r=0.02, q=0.005)
        assert {f: getattr(back, f) for f in g["from_array_fields"]} == g["from_array_fields"], name
    assert P.from_array(P().to_array()) == P()          # defaults of r, q are the reference's market constants


def test_price_european_marshalling_without_a_device():
    """Handle.price_european hands the library a ctypes strike array (lists / tuples) or the ndarray's address, plain
    Python scalars, and returns the landing buffer as a writable [n_strikes, NSUMS] array -- checked against a recording
    stand-in for the C entry point (no device, no compute)."""
    import ctypes as C

    class RecordingLib:
        def __init__(self):
            self.calls = []

        def b200mc_price_european(self, h, sp, S0, T, n_steps, n_paths, seed, path_offset, ks, nk, is_call, flags, bumps, out):
            strikes = [ks[i] for i in range(nk)] if not isinstance(ks, int) else \
                list((C.c_double * nk).from_address(ks))
            self.calls.append(dict(S0=S0, T=T, n_steps=n_steps, n_paths=n_paths, seed=seed, path_offset=path_offset,
                                   strikes=strikes, is_call=is_call, flags=flags, v0=sp._obj.v0, bumps=bumps))
            for i in range(nk):
                for j in range(_lib.NSUMS):
                    out[i * _lib.NSUMS + j] = 1000.0 * strikes[i] + j
            return 0

        def b200mc_destroy(self, h):
            return 0

    h = _lib.Handle.__new__(_lib.Handle)
    h.h, h.lib = C.c_void_p(1), RecordingLib()
    p = SVJParams(v0=0.0625)
    for strikes in ([1.0, 2.0, 3.0], (1.0, np.float64(2.0), 3), np.array([1.0, 2.0, 3.0]), np.array([1, 2, 3], dtype=np.int32)):
        out = h.price_european(p, np.float32(22500.0), 1, np.int64(50), np.int32(1000), 2 ** 64 + 5, strikes, np.bool_(True),
                               _lib.ANTITHETIC, None, path_offset=np.int64(7))
        c = h.lib.calls[-1]
        assert c["strikes"] == [1.0, 2.0, 3.0] and c["v0"] == 0.0625 and c["seed"] == 5 and c["path_offset"] == 7
        assert (c["S0"], c["T"], c["n_steps"], c["n_paths"], c["is_call"], c["flags"]) == (22500.0, 1.0, 50, 1000, 1, 1)
        assert all(type(c[k]) is t for k, t in (("S0", float), ("T", float), ("n_steps", int), ("n_paths", int)))
        assert out.shape == (3, _lib.NSUMS) and out.dtype == np.float64 and out.flags.writeable
        assert out[2, 4] == 3004.0 and out[0, 0] == 1000.0
    one = h.price_european(p, 22500.0, 1.0, 50, 1000, 42, 2500.0, False)          # a scalar strike
    assert one.shape == (1, _lib.NSUMS) and h.lib.calls[-1]["is_call"] == 0

    class Loose:                                                                  # fields float() understands only
        pass
    q = Loose()
    for f in _lib.PARAM_FIELDS:
        setattr(q, f, "0.25")
    assert _lib.to_params(q).kappa == 0.25 and _lib.to_params(SVJParams()).rho == -0.7
    h.h = None


def test_price_batch_reference_price_is_bs_price_bit_for_bit():
    """price_batch hoists the strike-independent terms of the Black-Scholes reference out of its strike loop; the result
    must be bs_price() to the last bit (same operations, same order), for calls and puts, lists and arrays."""
    p = SVJParams(kappa=2.0, theta=0.05, xi=0.4, rho=-0.5, v0=0.0625, lambda_j=0.3, mu_j=-0.02, sigma_j=0.07, r=0.031, q=0.017)
    e = MonteCarloEngine(p, 400, 40, 7, use_sobol=False, rng="philox", handle=OracleBackedHandle())
    ks = [70.0 + 3.7 * i for i in range(17)]
    for is_call in (True, False):
        for strikes in (ks, np.array(ks), tuple(ks)):
            rows = e.price_batch(100.0, strikes, 0.37, is_call)
            assert [r["strike"] for r in rows] == ks
            for r, K in zip(rows, ks):
                assert r["bs_ref"] == MC.bs_price(100.0, K, 0.37, p.r, p.q, math.sqrt(p.v0), is_call)
                one = e.price(100.0, K, 0.37, is_call)
                assert r["price"] == pytest.approx(one["price"], rel=1e-12) and r["bs_ref"] == one["bs_ref"]
