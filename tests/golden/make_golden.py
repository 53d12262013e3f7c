#!/usr/bin/env python
"""
tests/golden/make_golden.py -- generate the golden fixtures by RUNNING THE REFERENCE ITSELF.

Run in the dev container only (the reference tree does not exist on the GPU box):

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden.py [/root/reference]

Writes tests/golden/golden.json (scalars / result dicts) and tests/golden/golden_arrays.npz (small arrays).
Nothing from the reference's sources is copied: only the OUTPUTS of its public functions, on inputs
stated here, are stored.  The reference ships no golden vectors of its own, so these
pin the oracle (tests/test_oracle.py) and, through the oracle and directly, the CUDA path (tests/test_gpu_*.py).

All spots are passed as float (an int spot makes the numba kernel integer-typed).
"""
import hashlib
import json
import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
from engine.models import SVJParams  # noqa: E402
from engine import monte_carlo as ref_mc  # noqa: E402
from engine import greeks as ref_gk  # noqa: E402
from engine import risk as ref_risk  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
J = {"versions": {"numpy": np.__version__}, "cases": {}}
A = {}


def fl(d):
    """dict of numpy scalars -> plain floats (NaN kept as the string 'nan' for strict JSON)."""
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out[k] = fl(v)
        elif isinstance(v, (list, tuple)):
            out[k] = [fl(x) if isinstance(x, dict) else float(x) for x in v]
        else:
            v = float(v)
            out[k] = "nan" if np.isnan(v) else v
    return out


def pdict(p):
    return {k: float(getattr(p, k)) for k in
            ("kappa", "theta", "xi", "rho", "v0", "lambda_j", "mu_j", "sigma_j", "r", "q")}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


SVJ_DEFAULT = SVJParams()
GBM_CFG1 = SVJParams(kappa=3.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09,
                     lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.0)
HESTON = SVJParams(kappa=5.0, theta=0.04, xi=0.3, rho=-0.7, v0=0.04, lambda_j=0.0, mu_j=0.0, sigma_j=0.01)
JUMPY = SVJParams(kappa=2.0, theta=0.06, xi=0.8, rho=-0.5, v0=0.05, lambda_j=6.0, mu_j=-0.08, sigma_j=0.2,
                  r=0.03, q=0.01)
PSETS = {"svj_default": SVJ_DEFAULT, "gbm_cfg1": GBM_CFG1, "heston": HESTON, "jumpy": JUMPY}
J["params"] = {k: pdict(v) for k, v in PSETS.items()}


def kernel(p, spot, T, Z1, Z2, Zj, Zjs, steps, record=False, v0=None):
    return ref_mc._simulate_svj_paths_numba(
        float(spot), p.v0 if v0 is None else v0, p.r, p.q, T, p.kappa, p.theta, p.xi, p.rho,
        p.lambda_j, p.mu_j, p.sigma_j, Z1, Z2, Zj, Zjs, steps, record)


# ---- (1) kernel-level, inputs STORED (independent of numpy's generator) ------------------------------
rng = np.random.default_rng(20261018)
n, steps = 96, 40
Z = {k: rng.standard_normal((n, steps)) for k in ("Z1", "Z2", "Zjs")}
Z["Zj"] = rng.random((n, steps))
for k, v in Z.items():
    A[f"k_small_{k}"] = v
for name, p in PSETS.items():
    S, v, paths = kernel(p, 22500.0, 0.25, Z["Z1"], Z["Z2"], Z["Zj"], Z["Zjs"], steps, True)
    A[f"k_small_{name}_S"] = S
    A[f"k_small_{name}_v"] = v
    A[f"k_small_{name}_paths"] = paths

# ---- (2) kernel-level, inputs REGENERATED from the PCG64 seed (SURVEY.md section 8c probe) -----------
ge = ref_gk.GreeksEngine(SVJ_DEFAULT, 4096, 252, 42)
Z1, Z2, Zj, Zjs = ge._generate_shared_randoms(63)
S, v, _ = kernel(SVJ_DEFAULT, 22500.0, 0.25, Z1, Z2, Zj, Zjs, 63)
A["k_4096_S"] = S
A["k_4096_v"] = v
J["cases"]["k_4096"] = {
    "params": "svj_default", "n": 4096, "steps": 63, "seed": 42, "spot": 22500.0, "T": 0.25,
    "Z1_head": Z1[0, :3].tolist(), "Zj_head": Zj[0, :3].tolist(),
    "Z_sha256": {"Z1": sha(Z1), "Z2": sha(Z2), "Zj": sha(Zj), "Zjs": sha(Zjs)},
    "mean_S": float(S.mean()), "std_S": float(S.std()), "mean_v": float(v.mean()),
}
# antithetic twin and a v0-bumped run on the same draws (what price() :319-324 and vega :136-147 do)
A["k_4096_S_anti"] = kernel(SVJ_DEFAULT, 22500.0, 0.25, -Z1, -Z2, Zj, -Zjs, 63)[0]
A["k_4096_S_v0up"] = kernel(SVJ_DEFAULT, 22500.0, 0.25, Z1, Z2, Zj, Zjs, 63, v0=SVJ_DEFAULT.v0 + 0.01)[0]

# ---- (3) GBM-limit identity --------------------------------------------------------------------------
Z1g = np.random.default_rng(7).standard_normal((512, 250))
zeros = np.zeros_like(Z1g)
ones = np.ones_like(Z1g)
Sg, vg, _ = kernel(GBM_CFG1, 2500.0, 1.0, Z1g, zeros, ones, zeros, 250)
A["k_gbm_S"] = Sg
J["cases"]["k_gbm"] = {"seed": 7, "n": 512, "steps": 250, "spot": 2500.0, "T": 1.0,
                       "max_abs_v_minus_v0": float(np.max(np.abs(vg - GBM_CFG1.v0)))}

# ---- (4) pricer-level dicts --------------------------------------------------------------------------
pricer_cases = []
for pname, spot, K, T, n, ns, seed in [
    ("gbm_cfg1", 2500.0, 2500.0, 1.0, 50_000, 250, 42),          # BASELINE config 1
    ("svj_default", 22500.0, 22500.0, 0.25, 4096, 252, 42),      # SURVEY 8c SVJ smoke
    ("heston", 22500.0, 23000.0, 0.04, 4096, 100, 7),            # verify.py-like short expiry -> 10 steps
    ("jumpy", 100.0, 95.0, 0.5, 4096, 252, 3),
]:
    for sob, anti, cv in [(False, False, False), (False, True, False), (False, True, True), (False, False, True)]:
        for is_call in (True, False):
            eng = ref_mc.MonteCarloEngine(PSETS[pname], num_paths=n, num_steps=ns, seed=seed,
                                          use_sobol=sob, use_antithetic=anti, use_control_variate=cv)
            res = eng.price(spot, K, T, is_call)
            pricer_cases.append({"params": pname, "spot": spot, "strike": K, "T": T, "n": n, "num_steps": ns,
                                 "seed": seed, "sobol": sob, "anti": anti, "cv": cv, "is_call": is_call,
                                 "result": fl(res)})
# Sobol (host, degenerate bridge - quirk 1): small n so the fixture generation stays quick
for pname, spot, K, T, n, ns, seed, anti, cv in [
    ("gbm_cfg1", 2500.0, 2500.0, 1.0, 1024, 50, 42, True, False),
    ("gbm_cfg1", 2500.0, 2500.0, 1.0, 1024, 50, 42, True, True),
    ("svj_default", 22500.0, 22500.0, 0.25, 1000, 100, 5, True, True),
    ("svj_default", 22500.0, 22500.0, 0.25, 1000, 100, 5, False, False),
]:
    eng = ref_mc.MonteCarloEngine(PSETS[pname], num_paths=n, num_steps=ns, seed=seed,
                                  use_sobol=True, use_antithetic=anti, use_control_variate=cv)
    pricer_cases.append({"params": pname, "spot": spot, "strike": K, "T": T, "n": n, "num_steps": ns,
                         "seed": seed, "sobol": True, "anti": anti, "cv": cv, "is_call": True,
                         "result": fl(eng.price(spot, K, T, True))})
J["cases"]["price"] = pricer_cases
J["cases"]["bs"] = {
    "cfg1_call": float(ref_mc.bs_price(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, True)),
    "cfg1_put": float(ref_mc.bs_price(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, False)),
    "cfg1_delta_call": float(ref_mc.bs_delta(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, True)),
    "cfg1_delta_put": float(ref_mc.bs_delta(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, False)),
    "expired_call": float(ref_mc.bs_price(110.0, 100.0, 0.0, 0.05, 0.0, 0.2, True)),
    "expired_put_delta": float(ref_mc.bs_delta(90.0, 100.0, 0.0, 0.05, 0.0, 0.2, False)),
    "verify_py": float(ref_mc.bs_price(22500.0, 22500.0, 0.04, 0.065, 0.012, 0.2, True)),
}

# ---- (5) price_batch -----------------------------------------------------------------------------------
batch_cases = []
for pname, spot, strikes, T, n, ns, seed, anti, cv, is_call in [
    ("svj_default", 22500.0, [21000.0, 22500.0, 24000.0], 0.25, 4096, 252, 42, True, True, True),
    ("svj_default", 22500.0, [21000.0, 22500.0, 24000.0], 0.25, 4096, 252, 42, False, False, False),
    ("gbm_cfg1", 2500.0, list(np.linspace(0.7, 1.3, 21) * 2500.0), 0.5, 4096, 250, 11, True, True, False),
]:
    eng = ref_mc.MonteCarloEngine(PSETS[pname], num_paths=n, num_steps=ns, seed=seed,
                                  use_sobol=False, use_antithetic=anti, use_control_variate=cv)
    out = eng.price_batch(spot, np.array(strikes), T, is_call)
    batch_cases.append({"params": pname, "spot": spot, "strikes": [float(k) for k in strikes], "T": T, "n": n,
                        "num_steps": ns, "seed": seed, "anti": anti, "cv": cv, "is_call": is_call,
                        "result": [fl(r) for r in out]})
J["cases"]["price_batch"] = batch_cases

# ---- (6) sample paths ------------------------------------------------------------------------------------
eng = ref_mc.MonteCarloEngine(SVJ_DEFAULT, num_paths=1000, seed=42)
A["sample_paths_svj"] = eng.get_sample_paths(22500.0, 0.1, 8)       # steps = max(int(25.2), 50) = 50
A["sample_paths_gbm_1y"] = ref_mc.MonteCarloEngine(GBM_CFG1, num_paths=10, num_steps=250, seed=1) \
    .get_sample_paths(2500.0, 1.0, 5)

# ---- (7) Greeks ------------------------------------------------------------------------------------------
greek_cases = []
for pname, spot, K, T, n, ns, seed in [
    ("svj_default", 22500.0, 22500.0, 0.25, 4096, 252, 42),
    ("gbm_cfg1", 2500.0, 2500.0, 1.0, 4096, 250, 42),
    ("gbm_cfg1", 2500.0, 2500.0, 1.0, 50_000, 250, 42),         # SURVEY 8c values (delta/vega/gamma only)
]:
    for is_call in (True, False):
        g = ref_gk.GreeksEngine(PSETS[pname], num_paths=n, num_steps=ns, seed=seed)
        row = {"params": pname, "spot": spot, "strike": K, "T": T, "n": n, "num_steps": ns, "seed": seed,
               "is_call": is_call,
               "delta": fl(g.delta(spot, K, T, is_call)),
               "vega": fl(g.vega(spot, K, T, is_call)),
               "gamma": fl(g.gamma(spot, K, T, is_call))}
        if n <= 4096:          # theta / rho run the default Sobol engine (quirk 4): keep them small
            row["theta"] = fl(g.theta(spot, K, T, is_call))
            row["rho"] = fl(g.rho(spot, K, T, is_call))
        greek_cases.append(row)
J["cases"]["greeks"] = greek_cases

# ---- (8) risk metrics ------------------------------------------------------------------------------------
risk_cases = []
r0 = np.random.default_rng(42).standard_normal(10000) * 0.02 - 0.001      # verify.py:84-85
g8 = np.random.default_rng(8)
risk_inputs = {
    "verify_py": (r0, 0.99),
    "student_t3": (g8.standard_t(3, 50_000) * 0.01, 0.99),
    "conf95": (g8.standard_normal(4001) * 3.0 + 0.5, 0.95),
    "few_losses": (np.abs(g8.standard_normal(300)) - 0.02, 0.99),         # <= 20 losses -> tail_index NaN
    "tiny": (g8.standard_normal(25), 0.99),                               # cutoff == 0 branch
    "ties": (np.round(g8.standard_normal(5000), 1), 0.975),
    "option_pnl": (np.exp(-0.065) * np.maximum(A["k_gbm_S"] - 2500.0, 0.0) - 374.0712289657911, 0.99),
}
for name, (arr, conf) in risk_inputs.items():
    A[f"risk_{name}"] = arr
    risk_cases.append({"name": name, "confidence": conf, "result": fl(ref_risk.compute_risk_metrics(arr, conf))})
J["cases"]["risk"] = risk_cases

# ---- (9) Sobol / Brownian-bridge front end ---------------------------------------------------------------
J["cases"]["bb_order"] = {str(k): [int(x) for x in ref_mc._bb_ordering(k)] for k in (1, 2, 3, 7, 10, 16, 31, 50, 63)}
zz = np.random.default_rng(99).standard_normal((6, 31))
A["bb_in"] = zz
A["bb_out"] = ref_mc.brownian_bridge_reorder(zz, 31)
sob = ref_mc.generate_sobol_normals(100, 12, seed=3)
A["sobol_100x12_seed3"] = sob

# ---- Philox4x32-10 known answers (Random123 kat_vectors) --------------------------------------------------
J["cases"]["philox_kat"] = [
    {"ctr": ["00000000"] * 4, "key": ["00000000"] * 2, "out": ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]},
    {"ctr": ["ffffffff"] * 4, "key": ["ffffffff"] * 2, "out": ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]},
    {"ctr": ["243f6a88", "85a308d3", "13198a2e", "03707344"], "key": ["a4093822", "299f31d0"],
     "out": ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]},
]

with open(os.path.join(HERE, "golden.json"), "w") as f:
    json.dump(J, f, indent=1, sort_keys=True)
np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **A)
print("wrote", len(J["cases"]), "case groups,", len(A), "arrays,",
      sum(a.nbytes for a in A.values()) // 1024, "KiB raw")
