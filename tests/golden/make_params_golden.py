"""Golden values of the helpers on the reference's parameter dataclass (engine/models.py:46-84), written by the
reference itself:

    python tests/golden/make_params_golden.py      ->  tests/golden/params_golden.json

Parameter sets: the defaults, verify.py's set, the GBM special case, and sets built to trip every validate() branch
(Feller violated, |rho| above 0.999, v0 and theta above MAX_VARIANCE) singly and together.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from engine.models import SVJParams   # noqa: E402  (the reference itself)

SETS = {
    "defaults": {},
    "verify_py": dict(kappa=5.0, theta=0.04, xi=0.3, rho=-0.7, v0=0.04, lambda_j=0.5, mu_j=-0.03, sigma_j=0.08),
    "gbm": dict(kappa=0.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.0, r=0.065, q=0.0),
    "feller_edge": dict(kappa=2.0, theta=0.04, xi=0.4),                   # 2 kappa theta == xi^2: violated (strict >)
    "feller_violated": dict(kappa=1.0, theta=0.02, xi=0.9),
    "rho_limit": dict(rho=-0.9995),
    "rho_at_limit": dict(rho=0.999),
    "v0_large": dict(v0=10.5),
    "theta_large": dict(theta=12.0, kappa=0.1, xi=3.0),
    "everything": dict(kappa=0.5, theta=11.0, xi=5.0, rho=0.99999, v0=10.0001, lambda_j=3.0, mu_j=0.2, sigma_j=0.5,
                       r=0.01, q=0.03),
}

out = {}
for name, kw in SETS.items():
    p = SVJParams(**kw)
    arr = p.to_array()
    back = SVJParams.from_array(arr * 1.0, r=0.02, q=0.005)
    out[name] = {"kwargs": kw, "jump_compensation": float(p.jump_compensation), "feller_satisfied": bool(p.feller_satisfied),
                 "to_array": arr.tolist(), "validate": p.validate(),
                 "from_array_fields": {f: float(getattr(back, f)) for f in
                                       ("kappa", "theta", "xi", "rho", "v0", "lambda_j", "mu_j", "sigma_j", "r", "q")}}
with open(os.path.join(HERE, "params_golden.json"), "w", encoding="utf-8") as f:
    json.dump(out, f, indent=1, ensure_ascii=False, sort_keys=True)
print("wrote params_golden.json:", {k: len(v["validate"]) for k, v in out.items()})
