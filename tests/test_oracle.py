"""CPU: pin the oracle (oracle/) against fixtures produced by running the reference itself."""
import ctypes
import hashlib
import math
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_tree_close, unnan
from oracle import oracle as O

PNAMES = ["svj_default", "gbm_cfg1", "heston", "jumpy"]
RTOL = 1e-12   # oracle vs reference on identical float64 inputs (libm vs numba exp: ~1 ulp per step)


def P(golden, name):
    return O.Params(**golden["params"][name])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", PNAMES)
@pytest.mark.parametrize("impl", ["c", "numpy"])
def test_kernel_stored_inputs(golden, garr, name, impl):
    p = P(golden, name)
    f = O.simulate_svj if impl == "c" else O.simulate_svj_numpy
    S, v, paths = f(22500.0, p.v0, p.r, p.q, 0.25, p.kappa, p.theta, p.xi, p.rho, p.lambda_j, p.mu_j, p.sigma_j,
                    garr["k_small_Z1"], garr["k_small_Z2"], garr["k_small_Zj"], garr["k_small_Zjs"], 40, True)
    np.testing.assert_allclose(S, garr[f"k_small_{name}_S"], rtol=RTOL)
    np.testing.assert_allclose(v, garr[f"k_small_{name}_v"], rtol=RTOL, atol=1e-18)
    np.testing.assert_allclose(paths, garr[f"k_small_{name}_paths"], rtol=RTOL)
    assert paths.shape == (96, 41) and np.all(paths[:, 0] == 22500.0)


def test_kernel_pcg64_inputs(golden, garr):
    c = golden["cases"]["k_4096"]
    p = P(golden, c["params"])
    Z1, Z2, Zj, Zjs = O.draw_pcg64(c["seed"], c["n"], c["steps"])
    assert sha(Z1) == c["Z_sha256"]["Z1"] and sha(Z2) == c["Z_sha256"]["Z2"]
    assert sha(Zj) == c["Z_sha256"]["Zj"] and sha(Zjs) == c["Z_sha256"]["Zjs"]
    np.testing.assert_allclose(Z1[0, :3], c["Z1_head"], rtol=0, atol=0)
    S, v, empty = O._sim(p, c["spot"], c["T"], Z1, Z2, Zj, Zjs, c["steps"])
    assert empty.shape == (0, 0)
    np.testing.assert_allclose(S, garr["k_4096_S"], rtol=RTOL)
    np.testing.assert_allclose(v, garr["k_4096_v"], rtol=RTOL, atol=1e-18)
    # SURVEY.md 8c probe values
    assert abs(S.mean() - 22787.2608410784) < 1e-6 and abs(S.std() - 2512.2146334655) < 1e-6
    np.testing.assert_allclose(O._sim(p, c["spot"], c["T"], -Z1, -Z2, Zj, -Zjs, c["steps"])[0],
                               garr["k_4096_S_anti"], rtol=RTOL)
    np.testing.assert_allclose(O._sim(p, c["spot"], c["T"], Z1, Z2, Zj, Zjs, c["steps"], v0=p.v0 + 0.01)[0],
                               garr["k_4096_S_v0up"], rtol=RTOL)


def test_gbm_limit_identity(golden, garr):
    """xi=0, lambda=0, theta=v0: the kernel equals the closed-form log-sum."""
    p = P(golden, "gbm_cfg1")
    Z1 = np.random.default_rng(7).standard_normal((512, 250))
    S = O._sim(p, 2500.0, 1.0, Z1, np.zeros_like(Z1), np.ones_like(Z1), np.zeros_like(Z1), 250)[0]
    np.testing.assert_allclose(S, garr["k_gbm_S"], rtol=RTOL)
    dt = 1.0 / 250
    closed = 2500.0 * np.exp((p.r - p.q - 0.5 * p.v0) * 1.0 + np.sqrt(p.v0 * dt) * Z1.sum(axis=1))
    np.testing.assert_allclose(S, closed, rtol=1e-12)
    assert golden["cases"]["k_gbm"]["max_abs_v_minus_v0"] < 1e-15


def test_bs_closed_form(golden):
    b = golden["cases"]["bs"]
    assert O.bs_price(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, True) == pytest.approx(b["cfg1_call"], rel=1e-14)
    assert O.bs_price(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, False) == pytest.approx(b["cfg1_put"], rel=1e-14)
    assert O.bs_delta(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, True) == pytest.approx(b["cfg1_delta_call"], rel=1e-14)
    assert O.bs_delta(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, False) == pytest.approx(b["cfg1_delta_put"], rel=1e-14)
    assert O.bs_price(110.0, 100.0, 0.0, 0.05, 0.0, 0.2, True) == b["expired_call"] == 10.0
    assert O.bs_delta(90.0, 100.0, 0.0, 0.05, 0.0, 0.2, False) == b["expired_put_delta"] == -1.0
    assert O.bs_price(22500.0, 22500.0, 0.04, 0.065, 0.012, 0.2, True) == pytest.approx(b["verify_py"], rel=1e-14)
    assert b["cfg1_call"] == pytest.approx(374.0712289657911, rel=1e-14)


def _price_cases(golden, big):
    return [c for c in golden["cases"]["price"] if (c["n"] > 5000) == big]


def _check_price(golden, c):
    eng = O.MonteCarloOracle(P(golden, c["params"]), c["n"], c["num_steps"], c["seed"],
                             c["sobol"], c["anti"], c["cv"])
    got = eng.price(c["spot"], c["strike"], c["T"], c["is_call"])
    want = c["result"]
    assert set(got) == set(want)
    for k, w in want.items():
        # std_error of the degenerate anti-off pseudo-CV is exactly 0 in the reference (quirk 2)
        assert got[k] == pytest.approx(w, rel=1e-10, abs=1e-9), (k, c)


def test_price_small(golden):
    cases = _price_cases(golden, big=False)
    assert len(cases) >= 24
    for c in cases:
        _check_price(golden, c)


def test_price_cfg1_50k(golden):
    """BASELINE config 1 at the reference's own size (50k x 250): a few seconds of NumPy RNG."""
    cases = [c for c in _price_cases(golden, big=True) if c["is_call"]]
    for c in cases:
        _check_price(golden, c)
    plain = [c for c in cases if not c["anti"] and not c["cv"]][0]["result"]
    assert plain["price"] == pytest.approx(373.80196, abs=1e-5)
    assert plain["std_error"] == pytest.approx(2.57214, abs=1e-5)


def test_price_batch(golden):
    for c in golden["cases"]["price_batch"]:
        eng = O.MonteCarloOracle(P(golden, c["params"]), c["n"], c["num_steps"], c["seed"], False, c["anti"], c["cv"])
        got = eng.price_batch(c["spot"], np.array(c["strikes"]), c["T"], c["is_call"])
        assert len(got) == len(c["result"])
        for g, w in zip(got, c["result"]):
            assert set(g) == set(w)
            for k in w:
                assert g[k] == pytest.approx(w[k], rel=1e-10, abs=1e-9)


def test_sample_paths(golden, garr):
    got = O.MonteCarloOracle(P(golden, "svj_default"), 1000, seed=42).get_sample_paths(22500.0, 0.1, 8)
    assert got.shape == garr["sample_paths_svj"].shape == (8, 51)
    np.testing.assert_allclose(got, garr["sample_paths_svj"], rtol=RTOL)
    got = O.MonteCarloOracle(P(golden, "gbm_cfg1"), 10, 250, seed=1).get_sample_paths(2500.0, 1.0, 5)
    np.testing.assert_allclose(got, garr["sample_paths_gbm_1y"], rtol=RTOL)


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_greeks_small(golden, idx):
    c = [g for g in golden["cases"]["greeks"] if g["n"] <= 4096][idx]
    g = O.GreeksOracle(P(golden, c["params"]), c["n"], c["num_steps"], c["seed"])
    args = (c["spot"], c["strike"], c["T"], c["is_call"])
    for name in ("delta", "vega", "gamma", "theta", "rho"):
        got = getattr(g, name)(*args)
        assert set(got) == set(c[name])
        for k, w in c[name].items():
            # theta/rho are finite differences of near-equal prices / 1e-4: allow the amplified ulp noise
            tol = 1e-6 if name in ("theta", "rho") else 1e-9
            assert got[k] == pytest.approx(w, rel=tol, abs=1e-9), (name, k)


def test_greeks_cfg1_50k_delta_gamma(golden):
    c = [g for g in golden["cases"]["greeks"] if g["n"] == 50_000 and g["is_call"]][0]
    g = O.GreeksOracle(P(golden, c["params"]), c["n"], c["num_steps"], c["seed"])
    d = g.delta(c["spot"], c["strike"], c["T"], True)
    assert d["pathwise"] == pytest.approx(c["delta"]["pathwise"], rel=1e-10)
    assert d["pathwise"] == pytest.approx(0.640732, abs=1e-6)


def test_risk_metrics(golden, garr):
    for c in golden["cases"]["risk"]:
        got = O.risk_metrics(garr[f"risk_{c['name']}"], c["confidence"])
        want = unnan(c["result"])
        assert set(got) == set(want)
        for k, w in want.items():
            if np.isnan(w):
                assert np.isnan(got[k]), (c["name"], k)
            else:
                assert got[k] == pytest.approx(w, rel=1e-12, abs=1e-15), (c["name"], k)


def test_bb_order_and_reorder(golden, garr):
    for k, want in golden["cases"]["bb_order"].items():
        assert O.bb_order(int(k)) == want
    out = O.bb_reorder(garr["bb_in"], 31)
    np.testing.assert_allclose(out, garr["bb_out"], rtol=1e-13, atol=1e-16)
    # quirk 1: the bridge pins W_T to 0
    assert np.max(np.abs(out.sum(axis=1))) < 1e-12


def test_sobol_front_end(garr):
    np.testing.assert_array_equal(O.sobol_normals(100, 12, seed=3), garr["sobol_100x12_seed3"])


def test_philox_kat(golden):
    lib = O._load()
    for kat in golden["cases"]["philox_kat"]:
        ctr = np.array([int(x, 16) for x in kat["ctr"]], dtype=np.uint32)
        key = np.array([int(x, 16) for x in kat["key"]], dtype=np.uint32)
        want = np.array([int(x, 16) for x in kat["out"]], dtype=np.uint32)
        np.testing.assert_array_equal(O.philox4x32_10(ctr, key), want)
        out = np.zeros(4, dtype=np.uint32)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        lib.oracle_philox4x32_10(ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p), out.ctypes.data_as(u32p))
        np.testing.assert_array_equal(out, want)


def test_philox_block_words_layout():
    seed, off = 0x1234567890ABCDEF, (1 << 32) - 2       # crosses the path_lo carry
    w = O.philox_block_words(seed, off, 5, 3, 1)
    ctr = np.zeros((5, 3, 4), dtype=np.uint32)
    for i in range(5):
        for b in range(3):
            p = off + i
            ctr[i, b] = (p & 0xFFFFFFFF, p >> 32, b, 1)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32), (5, 3, 2))
    np.testing.assert_array_equal(w, O.philox4x32_10(ctr, key))


def test_oracle_equals_reference_on_the_device_draws():
    """tests/golden/fused_golden.npz: draws dumped from the GPU library, outputs computed by the REFERENCE kernel on them
    (make_fused_golden.py).  Pins the oracle to the reference on exactly the inputs the fused-mode parity tests use."""
    import json
    import os
    from conftest import GOLDEN
    cases = json.load(open(os.path.join(GOLDEN, "fused_golden_cases.json")))
    d = dict(np.load(os.path.join(GOLDEN, "fused_golden.npz")))
    assert set(cases) == {"gbm", "detvar", "heston", "svj", "jumpy"}
    for name, c in cases.items():
        p = O.Params(**c["params"])
        Z = [d[f"{name}_{w}"] for w in ("Z1", "Z2", "Zj", "Zjs")]
        S, v, paths = O._sim(p, c["S0"], c["T"], *Z, c["steps"], record=True)
        np.testing.assert_allclose(S, d[f"{name}_ref_S"], rtol=RTOL)
        np.testing.assert_allclose(v, d[f"{name}_ref_v"], rtol=RTOL, atol=1e-18)
        np.testing.assert_allclose(paths, d[f"{name}_ref_paths"], rtol=RTOL)
        np.testing.assert_allclose(O._sim(p, c["S0"], c["T"], -Z[0], -Z[1], Z[2], -Z[3], c["steps"])[0],
                                   d[f"{name}_ref_S_anti"], rtol=RTOL)
        if name == "jumpy":
            assert (Z[2] < p.lambda_j * (c["T"] / c["steps"])).sum() > 500          # the jumps really fire


# ---------------------------------------------------------------------------------------------- a11: risk.py callers
@pytest.mark.parametrize("idx", [0, 1])
def test_stress_report_against_the_reference(risk_golden, idx):
    c = risk_golden["stress"][idx]
    o = O.StressOracle(O.Params(**risk_golden["params"][c["params"]]), num_paths=c["num_paths"], seed=c["seed"])
    assert_tree_close(o.full_stress_report(c["spot"], c["strike"], c["T"], c["is_call"]), c["report"], rel=1e-9, abs_=1e-8)


@pytest.mark.parametrize("idx", [0, 1])
def test_hedging_backtest_against_the_reference(risk_golden, idx):
    c = risk_golden["hedge"][idx]
    o = O.HedgingOracle(O.Params(**risk_golden["params"][c["params"]]), seed=c["seed"])
    got = o.run_backtest(c["spot"], c["strike"], c["T"], c["is_call"], c["num_days"], c["txn_cost_bps"], c["slippage_bps"],
                         c["num_scenarios"], c["num_mc_paths"])
    assert_tree_close(got, c["result"], rel=1e-9, abs_=1e-7)


# ---------------------------------------------------------------------------------------------- 8(f)-3: implied vol
def test_iv_surface_against_the_reference(iv_golden):
    g = iv_golden
    for spreads, tag in ((g["spreads"], ""), (None, "_ns")):
        s = O.extract_iv_surface(float(g["spot"]), float(g["r"]), float(g["q"]), g["strikes"], g["maturities"], g["calls"],
                                 g["puts"], spreads)
        np.testing.assert_array_equal(s["valid_mask"], g["valid" + tag])
        np.testing.assert_allclose(s["iv_call"], g["iv_call" + tag], rtol=0, atol=1e-12, equal_nan=True)
        np.testing.assert_allclose(s["iv_put"], g["iv_put" + tag], rtol=0, atol=1e-12, equal_nan=True)
    for price, K, T, call, lo, hi, want in g["scalar"]:
        got = O.implied_vol(price, float(g["spot"]), K, T, float(g["r"]), float(g["q"]), bool(call), lo, hi)
        assert (got is None and np.isnan(want)) or got == pytest.approx(want, abs=1e-12)


# ---------------------------------------------------------------------------------------------- 8(f)-4: working QMC front end
@pytest.mark.parametrize("n", [1, 2, 7, 16, 63, 250])
def test_qmc_bridge_is_an_orthogonal_map(n):
    """Independent N(0,1) draws in bridge order -> independent N(0,1) step normals: the linear map is orthogonal, and
    its first column (dimension 0) carries the whole of W_T.  (The reference's own bridge fails this: bb_reorder above.)"""
    B = O.qmc_bridge(np.eye(n))                    # row k = image of unit draw k
    np.testing.assert_allclose(B @ B.T, np.eye(n), atol=1e-12)
    np.testing.assert_allclose(B.sum(axis=1), [math.sqrt(n)] + [0.0] * (n - 1), atol=1e-12)
    assert sorted(t for t, *_ in O.qmc_bridge_nodes(n)) == list(range(1, n + 1))


def test_qmc_draws_blocks():
    Z1, Z2, Zj, Zjs = O.qmc_draws(3, 64, 10, 4)
    assert Z1.shape == Z2.shape == Zj.shape == Zjs.shape == (64, 10)
    assert 0 < Zj.min() and Zj.max() < 1
    Z1b, Z2b, Zjb, Zjsb = O.qmc_draws(3, 64, 10, 1)
    assert not Z2b.any() and (Zjb == 1).all() and not Zjsb.any()


# ---------------------------------------------------------------------------------------------- NumPy's RNG front end
def test_ziggurat_tables_in_the_library_are_numpys():
    """csrc/np_ziggurat_tables.inc (compiled into libb200mc.so) against numpy's own static library, bit for bit."""
    import shutil
    from monte_carlo_option_simulator_b200 import _lib
    ki, wi, fi = _lib.numpy_ziggurat_tables()
    assert ki[0] == 0x000EF33D8025EF6A and ki[1] == 0 and wi[0] == 8.68362706080130616677e-16 and fi[0] == 1.0
    assert np.all(np.diff(fi) < 0) and fi[255] == pytest.approx(np.exp(-0.5 * 3.6541528853610088 ** 2), rel=1e-14)
    if shutil.which("ar"):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import extract_ziggurat_tables as X
        k2, w2, f2 = X.extract()
        assert np.array_equal(ki, k2) and np.array_equal(wi.view(np.uint64), w2.view(np.uint64))
        assert np.array_equal(fi.view(np.uint64), f2.view(np.uint64))


@pytest.mark.parametrize("seed,n", [(42, 200_000), (7, 50_000), (2 ** 63 + 5, 1000)])
def test_np_standard_normal_restatement_equals_numpy(seed, n):
    """The oracle's PCG64 + Ziggurat (checker of csrc/np_normal.cu) against NumPy itself, bit for bit."""
    from monte_carlo_option_simulator_b200 import _lib
    got, used = O.np_standard_normal(seed, n, _lib.numpy_ziggurat_tables())
    g = np.random.default_rng(seed)
    want = g.standard_normal(n)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert g.bit_generator.random_raw(1)[0] == np.random.PCG64(seed).random_raw(used + 1)[used]


def test_glibc_fma_log1p_restatement_is_the_hosts_log1p():
    """The tail draws of NumPy's Ziggurat go through the host's log1p; the device carries glibc's FMA build of it.  On a
    host with FMA (every x86-64 server CPU of the last decade) the restatement must equal libm bit for bit."""
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        flags = ""
    if " fma" not in flags:
        pytest.skip("host CPU without FMA: glibc selects its non-FMA log1p here")
    assert O.log1p_mismatches(2_000_000) == 0
    assert O._load().oracle_glibc_log1p_fma(-0.5) == math.log1p(-0.5)
