"""CPU: host-side logic of the device initialisers (popinit / netbuild): the lp.* distribution -> struct lpk_dist mapping
(parameter defaults of reference distributions.py:36-108), the copula parameters of reference model.py:838-847, and that the
product path refuses to run without a CUDA device instead of falling back."""

import numpy as np
import pytest

import laser_polio_b200 as lp
from laser_polio_b200 import _lpk, netbuild, popinit


def test_dist_struct_mapping_and_defaults():
    k = _lpk.DIST_KINDS
    d = popinit.dist_struct(lp.poisson(lam=3))
    assert (d.kind, d.a) == (k["poisson"], 3.0)
    d = popinit.dist_struct(lp.gamma(shape=4.51, scale=5.32))
    assert (d.kind, d.a, d.b) == (k["gamma"], 4.51, 5.32)
    d = popinit.dist_struct(lp.normal(mean=2, std=0.5))
    assert (d.kind, d.a, d.b) == (k["normal"], 2.0, 0.5)
    d = popinit.dist_struct(lp.uniform(min=2, max=9))
    assert (d.kind, d.a, d.b) == (k["uniform"], 2.0, 9.0)
    d = popinit.dist_struct(lp.constant(value=7))
    assert (d.kind, d.a) == (k["constant"], 7.0)
    d = popinit.dist_struct(lp.exponential(scale=2.5))
    assert (d.kind, d.a) == (k["exponential"], 2.5)
    # lp.lognormal is parameterised by the mean / sigma of the variable itself (reference distributions.py:73-87)
    d = popinit.dist_struct(lp.lognormal(mean=12.5, sigma=3.5))
    assert d.kind == k["lognormal"]
    assert np.isclose(np.exp(d.a + d.b**2 / 2), 12.5) and np.isclose((np.exp(d.b**2) - 1) * np.exp(2 * d.a + d.b**2), 3.5**2)
    d = popinit.dist_struct(lp.lognormal(mean=0, sigma=1))  # the reference returns zeros for a non-positive mean
    assert (d.kind, d.a) == (k["constant"], 0.0)
    with pytest.raises(TypeError):
        popinit.dist_struct(3.0)


def test_heterogeneity_parameters_follow_the_reference():
    pars = lp.PropertySet({"risk_mult_var": 4.0, "r0": 14.0, "corr_risk_inf": 0.8, "dur_inf": lp.gamma(shape=4.51, scale=5.32)})
    mu, sg, scale, rho, mean = popinit.heterogeneity_parameters(pars, mean_dur_inf=24.0)
    # model.py:838-847: lognormal with mean 1 and variance risk_mult_var; gamma(1) with mean r0 / E[dur_inf]; rho = 2 sin(pi c / 6)
    assert np.isclose(np.exp(mu + sg**2 / 2), 1.0) and np.isclose((np.exp(sg**2) - 1) * np.exp(2 * mu + sg**2), 4.0)
    assert np.isclose(scale, 14.0 / 24.0) and np.isclose(mean, 14.0 / 24.0) and np.isclose(rho, 2 * np.sin(np.pi * 0.8 / 6))
    np.random.seed(1)
    _, _, scale2, _, _ = popinit.heterogeneity_parameters(pars)  # the reference's own estimate from 1000 draws (model.py:842)
    assert abs(scale2 / (14.0 / (4.51 * 5.32)) - 1) < 0.1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("only meaningful without a CUDA device")
    pars = lp.PropertySet({"seed": 1, "risk_mult_var": 4.0, "r0": 14.0, "corr_risk_inf": 0.8, "individual_heterogeneity": True,
                           "dur_inf": lp.gamma(shape=4.51, scale=5.32)})
    a, b = torch.zeros(8), torch.zeros(8)
    with pytest.raises(_lpk.LpkError):
        popinit.populate_heterogeneous_values(0, 8, a, b, pars, mean_dur_inf=24.0)
    with pytest.raises(Exception):
        netbuild.distance_matrix([0.0, 1.0], [0.0, 1.0], device="cuda")
