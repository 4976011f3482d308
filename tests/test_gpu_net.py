"""GPU: network construction kernels (csrc/lpk_net.cu through netbuild) against the oracle's definition-level restatement.
float64 throughout; sin / cos / asin / pow differ in the last bits between the CUDA and the host math libraries and row sums
are reduced in a different order: tolerance rtol 1e-12 (atol 1e-300: entries span 30 orders of magnitude)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def env(oracle):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import laser_polio_b200 as lp
    from laser_polio_b200 import core, netbuild

    return lp, netbuild, core, oracle


def nodes(n, seed):
    rng = np.random.default_rng(seed)
    lat, lon = rng.uniform(4, 14, n), rng.uniform(3, 15, n)  # Nigeria's bounding box
    lat[n // 2], lon[n // 2] = lat[0], lon[0]                # two coincident nodes
    pops = np.round(np.exp(rng.normal(11, 1, n)))
    return lat, lon, pops


@pytest.mark.parametrize("n", [2, 37, 300])
def test_network_vs_oracle(env, n):
    lp, nb, core, orc = env
    lat, lon, pops = nodes(n, n)
    d = nb.distance_matrix(lat, lon)
    do = orc.net_haversine(lat, lon)
    np.testing.assert_allclose(d.cpu().numpy(), do, rtol=1e-12)
    assert do[0, n // 2] == 1.0 and np.all(np.diag(d.cpu().numpy()) == 0)
    g = nb.gravity(pops, d, 3.0, 1.0, 0.7, 1.8, norm=float(pops.sum() ** 1.8))
    np.testing.assert_allclose(g.cpu().numpy(), orc.net_gravity(pops, do, 3.0, 1.0, 0.7, 1.8, norm=pops.sum() ** 1.8), rtol=1e-12, atol=1e-300)
    for home in (False, True):
        r = nb.radiation(pops, d, 0.3, include_home=home)
        np.testing.assert_allclose(r.cpu().numpy(), orc.net_radiation(pops, d.cpu().numpy(), 0.3, include_home=home), rtol=1e-11, atol=1e-300)
    for cap in (0.01, 0.2, 10.0):
        got = nb.row_normalizer(r, cap).cpu().numpy()
        np.testing.assert_allclose(got, orc.net_row_normalize(r.cpu().numpy(), cap), rtol=1e-12, atol=1e-300)
        assert (got.sum(1) <= cap * (1 + 1e-12)).all()
    # and against the product's own host path (core.py, what runs without pars.device_init)
    np.testing.assert_allclose(r.cpu().numpy(), core.radiation(pops, d.cpu().numpy(), 0.3, include_home=True), rtol=1e-11, atol=1e-300)
    host = core.distance(lat[:, None], lon[:, None], lat[None, :], lon[None, :])
    host[~np.eye(n, dtype=bool) & (host == 0)] = 1.0
    np.fill_diagonal(host, 0.0)
    np.testing.assert_allclose(d.cpu().numpy(), host, rtol=1e-12)


def test_build_network_matches_host_path(env):
    lp, nb, core, orc = env
    n = 774
    lat, lon, pops = nodes(n, 5)
    lookup = {i: {"lat": float(lat[i]), "lon": float(lon[i])} for i in range(n)}
    for method, extra in (("gravity", {"gravity_k": 0.5, "gravity_k_exponent": -2.0, "gravity_a": 1.0, "gravity_b": 1.0, "gravity_c": 1.5}),
                          ("radiation", {"radiation_k_log10": -0.3})):
        pars = lp.PropertySet({"distances": None, "node_lookup": lookup, "migration_method": method, "max_migr_frac": 0.1, **extra})
        net = nb.build_network(pars, pops).cpu().numpy()
        dist = core.distance(lat[:, None], lon[:, None], lat[None, :], lon[None, :])
        dist[~np.eye(n, dtype=bool) & (dist == 0)] = 1
        if method == "gravity":
            host = core.gravity(pops, dist, 0.5 * 10 ** -2.0, 1.0, 1.0, 1.5) / pops.sum() ** 1.5
        else:
            host = core.radiation(pops, dist, 10 ** -0.3, include_home=False)
        np.testing.assert_allclose(net, core.row_normalizer(host, 0.1), rtol=1e-10, atol=1e-300)
    with pytest.raises(ValueError):
        nb.build_network(lp.PropertySet({"distances": None, "node_lookup": lookup, "migration_method": "teleport", "max_migr_frac": 0.1}), pops)
