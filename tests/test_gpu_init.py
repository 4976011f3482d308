"""GPU: the population-initialisation kernels (csrc/lpk_init.cu, through popinit / the C ABI) against the CPU oracle.

Integer outputs that involve no transcendental function (age bin, age, lifespan, missed flags) must be bit-exact.  Outputs
that pass through log / cos / exp / erfc / lgamma -- whose last bit differs between the CUDA and glibc math libraries -- are
held to: float32 columns rtol 2e-6; integer timers identical except where a double straddles an integer boundary
(allowance: 5 agents per million, each off by the sampler's granularity).
"""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

N = 1_000_003
SEED = 20261017


@pytest.fixture(scope="module")
def env(oracle):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import laser_polio_b200 as lp
    from laser_polio_b200 import _lpk, popinit

    return lp, _lpk, popinit, oracle


def pars_of(lp, **over):
    p = dict(seed=SEED, risk_mult_var=4.0, r0=14.0, corr_risk_inf=0.8, individual_heterogeneity=True, dur_exp=lp.poisson(lam=3),
             dur_inf=lp.gamma(shape=4.51, scale=5.32), t_to_paralysis=lp.lognormal(mean=12.5, sigma=3.5), missed_frac=0.1)
    p.update(over)
    return lp.PropertySet(p)


def test_heterogeneity_vs_oracle(env):
    lp, _lpk, popinit, orc = env
    pars = pars_of(lp)
    mean_dur = 4.51 * 5.32
    r, f = (torch.zeros(N, dtype=torch.float32, device="cuda") for _ in range(2))
    popinit.populate_heterogeneous_values(0, N, r, f, pars, mean_dur_inf=mean_dur)
    mu, sg, scale, rho, mean = popinit.heterogeneity_parameters(pars, mean_dur)
    ro, fo = np.zeros(N, np.float32), np.zeros(N, np.float32)
    orc.init_heterogeneity(0, N, ro, fo, mu, sg, scale, rho, True, mean, SEED)
    np.testing.assert_allclose(r.cpu().numpy(), ro, rtol=2e-6)
    np.testing.assert_allclose(f.cpu().numpy(), fo, rtol=2e-6, atol=1e-12)
    # slices with id_base reproduce the whole (pure function of seed, agent, stage)
    r2, f2 = (torch.zeros(N - 1000, dtype=torch.float32, device="cuda") for _ in range(2))
    popinit.populate_heterogeneous_values(0, N - 1000, r2, f2, pars, id_base=1000, mean_dur_inf=mean_dur)
    assert torch.equal(r2, r[1000:]) and torch.equal(f2, f[1000:])
    # heterogeneity off (model.py:864-866)
    popinit.populate_heterogeneous_values(0, 100, r, f, pars_of(lp, individual_heterogeneity=False), mean_dur_inf=mean_dur)
    assert (r[:100] == 1).all() and torch.allclose(f[:100], torch.tensor(float(14.0 / mean_dur)))


@pytest.mark.parametrize("dists", [
    ("poisson", "gamma", "lognormal"), ("normal", "exponential", "uniform"), ("constant", "poisson45", "gamma_small")])
def test_timers_vs_oracle(env, dists):
    lp, _lpk, popinit, orc = env
    table = {"poisson": (lp.poisson(lam=3), orc.dist("poisson", 3)), "gamma": (lp.gamma(shape=4.51, scale=5.32), orc.dist("gamma", 4.51, 5.32)),
             "lognormal": (lp.lognormal(mean=12.5, sigma=3.5), orc.dist("lognormal", *orc.lognormal_mu_sigma(12.5, 3.5))),
             "normal": (lp.normal(mean=4, std=2), orc.dist("normal", 4, 2)), "exponential": (lp.exponential(scale=20), orc.dist("exponential", 20)),
             "uniform": (lp.uniform(min=2, max=30), orc.dist("uniform", 2, 30)), "constant": (lp.constant(value=3), orc.dist("constant", 3)),
             "poisson45": (lp.poisson(lam=45), orc.dist("poisson", 45)), "gamma_small": (lp.gamma(shape=0.6, scale=20), orc.dist("gamma", 0.6, 20))}
    (de, oe), (di, oi), (dp, op) = (table[k] for k in dists)
    pars = pars_of(lp, dur_exp=de, dur_inf=di, t_to_paralysis=dp)
    et, it, pt = (torch.zeros(N, dtype=torch.int8, device="cuda") for _ in range(3))
    popinit.init_timers(0, N, et, it, pt, pars)
    eo, io, po = (np.zeros(N, np.int8) for _ in range(3))
    orc.init_timers(0, N, eo, io, po, oe, oi, op, SEED)
    for dev, ref in ((et, eo), (it, io), (pt, po)):
        bad = int((dev.cpu().numpy() != ref).sum())
        assert bad <= 5, bad
    assert int(et.min()) >= 0 and int(it.min()) >= 0 and bool((pt <= it).all())


def test_demography_and_missed_vs_oracle(env):
    lp, _lpk, popinit, orc = env
    pyr = np.array([[5 * k, 5 * k + 4, int(1.7e7 * np.exp(-0.16 * k)), int(1.6e7 * np.exp(-0.16 * k))] for k in range(20)] + [[100, 100, 300, 500]])
    from laser_polio_b200 import utils
    cum = utils.create_cumulative_deaths(2_000_000, max_age_years=100)
    dob, dod = (torch.zeros(N, dtype=torch.int32, device="cuda") for _ in range(2))
    ri = torch.zeros(N, dtype=torch.int16, device="cuda")
    popinit.init_demography(0, N, dob, dod, ri, pyr, cum, SEED)
    lo = np.maximum(pyr[:, 0] * 365, 1).astype(np.int32)
    hi = ((pyr[:, 1] + 1) * 365).astype(np.int32)
    cdf = np.cumsum((pyr[:, 2] + pyr[:, 3]).astype(np.float64))
    do, dd, ro = np.zeros(N, np.int32), np.zeros(N, np.int32), np.zeros(N, np.int16)
    orc.init_demography(0, N, do, dd, ro, cdf, lo, hi, np.insert(np.asarray(cum, np.int64), 0, 0), 100, SEED)
    assert np.array_equal(dob.cpu().numpy(), do) and np.array_equal(dod.cpu().numpy(), dd) and np.array_equal(ri.cpu().numpy(), ro)
    # ages only (no VitalDynamics): NULL date_of_death / ri_timer
    dob2 = torch.zeros(N, dtype=torch.int32, device="cuda")
    popinit.init_demography(0, N, dob2, None, None, pyr, None, SEED)
    assert torch.equal(dob2, dob)
    for k in (0, 1, N // 10, N):
        m = torch.zeros(N, dtype=torch.uint8, device="cuda")
        popinit.init_missed(N, k, m, SEED)
        mo = np.zeros(N, np.uint8)
        orc.init_missed(N, k, mo, SEED)
        assert int(m.sum()) == k and np.array_equal(m.cpu().numpy(), mo)


def test_bad_arguments_raise(env):
    lp, _lpk, popinit, orc = env
    t = torch.zeros(8, dtype=torch.int8, device="cuda")
    with pytest.raises(TypeError):
        popinit.init_timers(0, 8, t.float(), t, t, pars_of(lp))
    bad = _lpk.Dist(99, 0.0, 0.0)
    ok = popinit.dist_struct(lp.poisson(lam=3))
    with pytest.raises(ValueError):
        _lpk.check(_lpk.lib().lpk_init_timers(C.c_int64(0), C.c_int64(8), _lpk.ptr(t), _lpk.ptr(t), _lpk.ptr(t), C.byref(bad), C.byref(ok),
                                              C.byref(ok), C.c_uint64(1), C.c_uint64(0), _lpk.stream_handle()), "lpk_init_timers")
    with pytest.raises(TypeError):
        popinit.dist_struct("poisson")


def test_sim_with_device_init(env, tmp_path):
    """SEIR_ABM(pars.device_init=True): every per-agent draw comes from the kernels; same seed -> same table and results
    (the reference's reproducibility contract, tests/test_prng_seeding.py:21-34), different seed -> different table; the
    population identities of tests/test_vital_dynamics.py:44-58 hold."""
    lp = env[0]
    path = tmp_path / "pyramid.csv"
    rows = ["Age,M,F"] + [f"{5 * k}-{5 * k + 4},{int(1.7e7 * np.exp(-0.16 * k))},{int(1.6e7 * np.exp(-0.16 * k))}" for k in range(20)] + ["100+,300,500"]
    path.write_text("\n".join(rows) + "\n")

    def build(seed):
        pars = lp.PropertySet({"start_date": lp.date("2020-01-01"), "dur": 40, "init_pop": np.array([30_000, 20_000, 10_000]),
                               "cbr": np.array([35.0, 30.0, 25.0]), "r0_scalars": np.ones(3), "age_pyramid_path": str(path),
                               "init_immun": 0.3, "init_prev": 0.01, "r0": 14, "missed_frac": 0.1, "vx_prob_ri": 0.5, "vx_prob_ipv": 0.5,
                               "stop_if_no_cases": False, "verbose": 0, "seed": seed, "device_init": True,
                               "distances": np.array([[0, 50, 90], [50, 0, 60], [90, 60, 0.0]]), "migration_method": "gravity",
                               "gravity_k": 0.5, "gravity_a": 1, "gravity_b": 1, "gravity_c": 2.0, "max_migr_frac": 0.1})
        sim = lp.SEIR_ABM(pars)
        sim.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.Transmission_ABM]
        return sim

    a, b, c = build(11), build(11), build(12)
    n = a.people.count
    for col in ("acq_risk_multiplier", "daily_infectivity", "exposure_timer", "infection_timer", "paralysis_timer", "date_of_birth",
                "date_of_death", "ri_timer", "chronically_missed"):
        assert np.array_equal(getattr(a.people, col), getattr(b.people, col)), col
        assert not np.array_equal(getattr(a.people, col), getattr(c.people, col)), col
    assert int(a.people.chronically_missed.sum()) == int(0.1 * n)
    assert abs(float(a.people.acq_risk_multiplier.mean()) - 1.0) < 0.05 and (a.people.date_of_birth[:n] < 0).all()
    assert (a.people.date_of_death[:n] >= 1).all()
    a.run()
    b.run()
    for name in ("S", "E", "I", "R", "births", "deaths", "new_exposed", "ri_vaccinated"):
        assert np.array_equal(getattr(a.results, name), getattr(b.results, name)), name
    r = a.results
    assert r.new_exposed.sum() > 0 and r.births.sum() > 0
    assert np.array_equal(r.pop[1:], r.pop[:-1] + r.births[1:] - r.deaths[1:])
