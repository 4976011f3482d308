"""CPU: the host network construction (core.py: distance / gravity / radiation / row_normalizer, the default path of
Transmission_ABM without pars.device_init) against the oracle's definition-level restatement (oracle/oracle.py net_*), i.e.
the same checker the CUDA kernels of csrc/lpk_net.cu are held to in tests/test_gpu_net.py."""

import numpy as np
import pytest

from laser_polio_b200 import core
from oracle import oracle as orc


def nodes(n, seed):
    rng = np.random.default_rng(seed)
    lat, lon = rng.uniform(4, 14, n), rng.uniform(3, 15, n)
    lat[n // 2], lon[n // 2] = lat[0], lon[0]  # two coincident nodes
    return lat, lon, np.round(np.exp(rng.normal(11, 1, n)))


@pytest.mark.parametrize("n", [2, 9, 60])
def test_core_network_vs_definitions(n):
    lat, lon, pops = nodes(n, n)
    d = core.distance(lat[:, None], lon[:, None], lat[None, :], lon[None, :])
    d[~np.eye(n, dtype=bool) & (d == 0)] = 1.0
    np.fill_diagonal(d, 0.0)
    do = orc.net_haversine(lat, lon)
    np.testing.assert_allclose(d, do, rtol=1e-12)
    g = core.gravity(pops, d, 2.0, 1.0, 0.8, 1.7) / pops.sum() ** 1.7
    np.testing.assert_allclose(g, orc.net_gravity(pops, do, 2.0, 1.0, 0.8, 1.7, norm=pops.sum() ** 1.7), rtol=1e-12, atol=1e-300)
    for home in (False, True):
        r = core.radiation(pops, d, 0.4, include_home=home)
        np.testing.assert_allclose(r, orc.net_radiation(pops, do, 0.4, include_home=home), rtol=1e-11, atol=1e-300)
        assert np.all(np.diag(r) == 0)
    for cap in (0.01, 0.3, 5.0):
        got = core.row_normalizer(r, cap)
        np.testing.assert_allclose(got, orc.net_row_normalize(r, cap), rtol=1e-12, atol=1e-300)
        assert (got.sum(1) <= cap * (1 + 1e-12)).all()
        under = r.sum(1) <= cap
        assert np.array_equal(got[under], r[under])  # rows under the cap are left alone (debug_row_normalizer.py:30-33)


def test_radiation_ties_share_a_radius():
    # three destinations at the same distance from node 0: each sees the other two inside its radius
    pops = np.array([100.0, 10.0, 20.0, 30.0])
    d = np.array([[0, 5, 5, 5], [5, 0, 7, 7], [5, 7, 0, 7], [5, 7, 7, 0.0]])
    r = core.radiation(pops, d, 1.0, include_home=False)
    for j, pj in ((1, 10.0), (2, 20.0), (3, 30.0)):
        s = 60.0 - pj
        assert np.isclose(r[0, j], 100.0 * pj / ((100.0 + s) * (100.0 + pj + s)))
    np.testing.assert_allclose(r, orc.net_radiation(pops, d, 1.0), rtol=1e-13)
