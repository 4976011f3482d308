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


def test_row_normalizer_on_the_references_worked_example():
    """scripts/sandbox/debug_row_normalizer.py:24-33 of the reference spells the operation out by hand -- rows whose sum exceeds
    the cap are multiplied by cap / rowsum, the others are left alone -- on this very matrix."""
    network = np.array([[0, 6, 2], [10, 0, 13], [15, 10, 0]], dtype=float)
    cap = 0.3
    manual = network.copy()
    rowsums = manual.sum(axis=1)
    big = rowsums > cap
    manual[big] = manual[big] * cap / rowsums[big, np.newaxis]
    got = core.row_normalizer(network, cap)
    np.testing.assert_allclose(got, manual, rtol=1e-15)
    np.testing.assert_allclose(got.sum(1), [0.3, 0.3, 0.3])
    small = network / 150.0  # every row under the cap: untouched
    assert np.array_equal(core.row_normalizer(small, cap), small)


def test_gravity_on_the_references_worked_example():
    """debug_row_normalizer.py:6-15 / debug_negative_network.py:74 give the formula in a comment: k * (pop^a * pop^b) / dist^c."""
    pops = np.array([99510, 595855, 263884]) / 1e3
    dist = np.array([[0, 4, 66], [4, 0, 827], [66, 827, 0]], dtype=float)
    g = core.gravity(pops, dist, 100, 1, 1, 2.0)
    for i in range(3):
        for j in range(3):
            want = 0.0 if i == j else 100 * pops[i] * pops[j] / dist[i, j] ** 2
            assert np.isclose(g[i, j], want, rtol=1e-14)


def test_radiation_telescopes_to_the_published_normalisation():
    """Simini et al.'s radiation model, the published algorithm laser-core's ``radiation`` implements (the package is not in
    the checkout and the reference's tests pin it only qualitatively, tests/test_migration.py:13-72): with distinct distances,
    p_i p_j / ((p_i + s_ij)(p_i + p_j + s_ij)) telescopes over destinations in order of distance, so every row sums to
    k (1 - p_i / P) exactly -- a closed form that fixes which population counts as 'within the radius' (neither i nor j)."""
    rng = np.random.default_rng(8)
    n = 40
    pops = np.round(np.exp(rng.normal(10, 1.2, n)))
    xy = rng.uniform(0, 500, (n, 2))
    d = np.sqrt(((xy[:, None] - xy[None]) ** 2).sum(-1))
    r = core.radiation(pops, d, 0.7, include_home=False)
    np.testing.assert_allclose(r.sum(axis=1), 0.7 * (1.0 - pops / pops.sum()), rtol=1e-12)
    # and it is monotone: a nearer destination of equal size always receives more
    i = 0
    order = np.argsort(d[i])
    same = core.radiation(np.full(n, 1000.0), d, 1.0, include_home=False)[i, order[1:]]
    assert np.all(np.diff(same) < 0)
