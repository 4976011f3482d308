"""Host-side samplers with the reference's ``lp.*`` distribution API (reference distributions.py:17-135).

``lp.poisson(lam=3)`` etc. return a callable ``dist(size) -> numpy array``; they are used only by the
one-off initialisers (timers, re-seeded infection timers), never inside the per-tick device path.
The reference fills normal / gamma samples with numba ``prange`` loops over numba's own RNG
(distributions.py:138-151); here they come from numpy's global stream, which is what
``lp.seed`` / ``pars.seed`` seeds.
"""

from __future__ import annotations

import numpy as np

__all__ = ["Distribution", "constant", "exponential", "gamma", "lognormal", "normal", "poisson", "uniform"]

_SUPPORTED = ("constant", "exponential", "gamma", "lognormal", "normal", "poisson", "uniform")


class Distribution:
    def __init__(self, dist_type: str, **pars):
        if dist_type not in _SUPPORTED:
            raise ValueError(f"Unsupported distribution: {dist_type}. Supported: {set(_SUPPORTED)}")
        self.dist_type = dist_type
        self.pars = pars

    def sample(self, size=1):
        kind, p = self.dist_type, self.pars
        if kind == "constant":
            return np.full(size, p.get("value", 1))
        if kind == "exponential":
            return np.random.exponential(p.get("scale", 1.0), size)
        if kind == "gamma":
            return np.random.gamma(p.get("shape", 2.0), p.get("scale", 1.0), size)
        if kind == "lognormal":
            m, s = p.get("mean", 1.0), p.get("sigma", 0.5)
            if m <= 0:
                return np.zeros(size)
            mu = np.log(m**2 / np.sqrt(s**2 + m**2))  # parameters of the underlying normal
            sg = np.sqrt(np.log(s**2 / m**2 + 1))
            return np.random.lognormal(mean=mu, sigma=sg, size=size)
        if kind == "normal":
            return np.random.normal(p.get("mean", 0.0), p.get("std", 1.0), size)
        if kind == "poisson":
            return np.random.poisson(p.get("lam", 5), size)
        return np.random.randint(p.get("min", 2), p.get("max", 10), size)  # "uniform": integers in [min, max)

    __call__ = sample

    def __repr__(self):
        return f"Distribution(type={self.dist_type}, pars={self.pars})"


def constant(value):
    return Distribution("constant", value=value)


def exponential(scale):
    return Distribution("exponential", scale=scale)


def gamma(shape, scale):
    return Distribution("gamma", shape=shape, scale=scale)


def lognormal(mean, sigma):
    return Distribution("lognormal", mean=mean, sigma=sigma)


def normal(mean, std):
    return Distribution("normal", mean=mean, std=std)


def poisson(lam):
    return Distribution("poisson", lam=lam)


def uniform(min, max):  # noqa: A002 - the reference's keyword names
    return Distribution("uniform", min=min, max=max)
