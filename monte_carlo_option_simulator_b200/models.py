"""Parameter container with the field set of the reference's SVJParams (engine/models.py:31-44).

The engines are duck-typed on these ten attributes, so the reference's own dataclass works too; this one exists so
the package can be used without the reference on the path."""
from __future__ import annotations

from dataclasses import dataclass, replace

RISK_FREE_RATE = 0.065     # engine/config.py
DIVIDEND_YIELD = 0.012


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
