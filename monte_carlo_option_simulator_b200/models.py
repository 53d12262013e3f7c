"""Parameter container with the field set of the reference's SVJParams (engine/models.py:31-44).

The engines are duck-typed on these ten attributes, so the reference's own dataclass works too; this one exists so
the package can be used without the reference on the path."""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import List

import numpy as np

RISK_FREE_RATE = 0.065     # engine/config.py
DIVIDEND_YIELD = 0.012
MAX_VARIANCE = 10.0        # engine/config.py:76
_ARRAY_FIELDS = ("kappa", "theta", "xi", "rho", "v0", "lambda_j", "mu_j", "sigma_j")   # order of to_array / from_array


@dataclass
class SVJParams:
    kappa: float = 3.0
    theta: float = 0.04
    xi: float = 0.5
    rho: float = -0.7
    v0: float = 0.04
    lambda_j: float = 1.0
    mu_j: float = -0.05
    sigma_j: float = 0.10
    r: float = RISK_FREE_RATE
    q: float = DIVIDEND_YIELD

    def replace(self, **kw) -> "SVJParams":
        return replace(self, **kw)

    # The helpers of the reference's dataclass (engine/models.py:46-84) that its callers use on a parameter set:
    # verify.py:14 and engine/guards.py:68,93 read the two properties, engine/calibration.py:242-266 calls
    # validate() / to_array() on the optimum.  Same names, same values, so either class can be handed around.
    @property
    def jump_compensation(self) -> float:
        """Drift compensator of the jumps, E[exp(J)] - 1 with J ~ N(mu_j, sigma_j^2) -- the `k` of the step drift
        (r - q - lambda_j k - v/2) dt (engine/models.py:46-49, engine/monte_carlo.py:209)."""
        return float(np.exp(self.mu_j + 0.5 * self.sigma_j ** 2) - 1.0)   # NumPy's exp, as there: same last bit

    @property
    def feller_satisfied(self) -> bool:
        """2 kappa theta > xi^2 (engine/config.py:141-143)."""
        return 2.0 * self.kappa * self.theta > self.xi * self.xi

    def to_array(self) -> np.ndarray:
        """The eight model parameters in the optimiser's order (engine/models.py:55-60); r and q are market data."""
        return np.array([getattr(self, f) for f in _ARRAY_FIELDS], dtype=float)

    @classmethod
    def from_array(cls, arr, r: float = RISK_FREE_RATE, q: float = DIVIDEND_YIELD) -> "SVJParams":
        """Inverse of to_array (engine/models.py:62-70)."""
        return cls(**{f: arr[i] for i, f in enumerate(_ARRAY_FIELDS)}, r=r, q=q)

    def validate(self) -> List[str]:
        """Warnings in the reference's wording (engine/models.py:72-84): Feller, |rho| > 0.999, v0 or theta above
        MAX_VARIANCE.  An empty list means the set is usable."""
        out = []
        if not self.feller_satisfied:
            out.append(f"Feller violated: 2κθ={2 * self.kappa * self.theta:.4f} ≤ ξ²={self.xi ** 2:.4f}")
        if abs(self.rho) > 0.999:
            out.append(f"|ρ|={abs(self.rho):.4f} exceeds 0.999")
        for name, value in (("v0", self.v0), ("θ", self.theta)):
            if value > MAX_VARIANCE:
                out.append(f"{name}={value:.4f} exceeds MAX_VARIANCE={MAX_VARIANCE}")
        return out

    @classmethod
    def gbm(cls, sigma: float, r: float = RISK_FREE_RATE, q: float = 0.0) -> "SVJParams":
        """Black-Scholes dynamics as the SVJ special case xi = 0, lambda_j = 0, kappa = 0 (variance frozen at
        sigma^2)."""
        return cls(kappa=0.0, theta=sigma * sigma, xi=0.0, rho=0.0, v0=sigma * sigma, lambda_j=0.0, mu_j=0.0,
                   sigma_j=0.0, r=r, q=q)


def copy_with(p, **kw):
    """A copy of any SVJParams-like object (reference's or ours) with some fields replaced."""
    fields = ("kappa", "theta", "xi", "rho", "v0", "lambda_j", "mu_j", "sigma_j", "r", "q")
    d = {f: getattr(p, f) for f in fields}
    d.update(kw)
    try:
        return type(p)(**d)
    except TypeError:
        return SVJParams(**d)
