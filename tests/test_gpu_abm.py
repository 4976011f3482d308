"""The reference's own integration tests for the per-tick path, re-pointed at this implementation.

Each test builds a tiny real sim through the component API exactly as the reference's tests do
(``lp.SEIR_ABM(PropertySet{...})`` + ``sim.components = [...]`` + ``sim.run()``), mutates
``sim.people.<col>`` as host numpy before the run and reads columns / ``sim.results`` back afterwards.
The reference test each one mirrors is cited (paths relative to the reference checkout).
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def lp():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import laser_polio_b200 as lp

    return lp


@pytest.fixture(scope="module")
def pyramid(tmp_path_factory):
    """Synthetic 5-year-bin age pyramid in the reference's CSV format (Age,M,F; last row '100+')."""
    path = tmp_path_factory.mktemp("data") / "pyramid.csv"
    rows = ["Age,M,F"]
    for k in range(20):
        n = int(17_000_000 * np.exp(-0.16 * k))
        rows.append(f"{5 * k}-{5 * k + 4},{n},{int(n * 0.97)}")
    rows.append("100+,300,500")
    path.write_text("\n".join(rows) + "\n")
    return str(path)


def base_pars(lp, pyramid, **over):
    p = {
        "start_date": lp.date("2020-01-01"), "dur": 30, "init_pop": np.array([1000, 500]), "cbr": np.array([30, 25]),
        "r0_scalars": np.array([0.5, 2.0]), "age_pyramid_path": pyramid, "init_immun": 0.0, "init_prev": 0.0,
        "stop_if_no_cases": False, "verbose": 0, "seed": 7,
    }
    p.update(over)
    return lp.PropertySet(p)


# ---------------------------------------------------------------- tests/test_diseasestate_abm.py:64-122
def test_progression_without_transmission(lp, pyramid):
    sims = []
    for dur in (1, 2, 3):
        sim = lp.SEIR_ABM(base_pars(lp, pyramid, dur=dur, dur_exp=lp.constant(value=1), dur_inf=lp.constant(value=1)))
        sim.components = [lp.DiseaseState_ABM]
        sims.append(sim)
    n = 1500
    assert np.all(sims[0].people.exposure_timer[:n] == 1) and np.all(sims[0].people.infection_timer[:n] == 1)
    for sim, expect in zip(sims, (1, 2, 3)):  # all E after 1 day, all I after 2, all R after 3
        sim.people.disease_state[:n] = 1
        sim.run()
        for state in (0, 1, 2, 3):
            assert np.sum(sim.people.disease_state == state) == (n if state == expect else 0), (expect, state)


# ---------------------------------------------------------------- tests/test_diseasestate_abm.py:258-351
def test_disease_timers_with_trans_explicit(lp, pyramid):
    """Two agents in two nodes; the lone susceptible sits in the OTHER node, so its exposure goes through the importation
    branch with an expected count of ~1: in the reference a Poisson-like draw that is >= 1 with probability ~0.63, which its
    seed = 124 happens to realise on day 1.  Here the same event has the same probability (the node's scale is
    solved for E[min(K, 1)] = 1 - exp(-E)), so the test walks seeds from 124 until the exposure lands on day 1 -- the timer
    arithmetic after that is seed-free -- and also checks that it does not land there for every seed."""
    dur_exp, dur_inf = 2, 3
    day1 = []
    for seed in range(124, 140):
        pars = base_pars(
            lp, pyramid, start_date=lp.date("2018-01-01"), dur=30, init_pop=np.array([1, 1]), cbr=np.array([0, 0]),
            r0_scalars=np.array([1.0, 1.0]), init_prev=1, dur_exp=lp.constant(value=dur_exp), dur_inf=lp.constant(value=dur_inf),
            t_to_paralysis=lp.constant(value=10), p_paralysis=1 / 2000, r0=999, distances=np.array([[0, 1], [1, 0]]),
            individual_heterogeneity=False, seed=seed,
        )
        sim = lp.SEIR_ABM(pars)
        sim.components = [lp.DiseaseState_ABM, lp.Transmission_ABM, lp.VitalDynamics_ABM]
        state = sim.people.disease_state[: sim.people.count]
        assert (np.sum(state == 0), np.sum(state == 1), np.sum(state == 2), np.sum(state == 3)) == (1, 0, 1, 0)
        assert np.all(sim.people.exposure_timer == dur_exp) and np.all(sim.people.infection_timer == dur_inf)
        assert np.all(sim.people.paralysis_timer <= sim.people.infection_timer)
        sim.run()
        day1.append(bool(sim.results.E[1].sum() == 1))
        if not day1[-1]:
            continue
        n_s, n_e, n_i, n_r = (np.sum(getattr(sim.results, k), axis=1) for k in "SEIR")
        n_npp = np.sum(sim.results.new_potentially_paralyzed, axis=1)
        zeros = np.zeros(sim.pars.dur + 1, int)
        s_exp = zeros.copy(); s_exp[0] = 1  # noqa: E702
        e_exp = zeros.copy(); e_exp[1 : 2 + dur_exp] = 1  # noqa: E702
        i_exp = zeros.copy(); i_exp[0 : dur_inf + 1] += 1; i_exp[2 + dur_exp : 2 + dur_exp + dur_inf] += 1  # noqa: E702
        r_exp = zeros.copy(); r_exp[1 + dur_inf :] += 1; r_exp[2 + dur_exp + dur_inf :] += 1  # noqa: E702
        p_exp = zeros.copy(); p_exp[1 + dur_inf] += 1; p_exp[2 + dur_exp + dur_inf] += 1  # noqa: E702
        assert np.all(n_s == s_exp) and np.all(n_e == e_exp) and np.all(n_i == i_exp) and np.all(n_r == r_exp)
        assert np.all(n_npp == p_exp)
    assert 4 <= sum(day1) <= 15, day1  # ~0.63 per seed: neither never nor always


# ---------------------------------------------------------------- tests/test_diseasestate_abm.py:497-541
def test_paralysis_progression_manual(lp, pyramid):
    pars = base_pars(lp, pyramid, dur=3, init_pop=np.array([4, 4]), cbr=np.array([0]), r0_scalars=np.array([0.0]),
                     dur_exp=lp.constant(value=1), dur_inf=lp.constant(value=1), p_paralysis=1.0)
    sim = lp.SEIR_ABM(pars)
    sim.components = [lp.DiseaseState_ABM, lp.Transmission_ABM]
    sim.people.disease_state[:] = np.array([0, 0, 1, 1, 2, 2, 3, 3])
    sim.people.paralysis_timer[:] = 1
    protected = np.array([1, 3, 5, 7])
    unprotected = np.setdiff1d(np.arange(sim.people.count), protected)
    sim.people.ipv_protected[:] = 0
    sim.people.ipv_protected[protected] = 1
    sim.run()
    assert np.sum(sim.people.potentially_paralyzed[protected] <= 0) == 4
    assert np.sum(sim.people.potentially_paralyzed[unprotected] > 0) == 2
    assert np.sum(sim.people.paralyzed[protected] <= 0) == 4
    assert np.sum(sim.people.paralyzed[unprotected] > 0) == 2
    assert np.sum(sim.results.potentially_paralyzed[-1]) == 2 and np.sum(sim.results.paralyzed[-1]) == 2


# ---------------------------------------------------------------- tests/test_transmission.py:63-113
def tx_sim(lp, pyramid, dur=1, r0=14, r0_scalars=None, init_prev=0.01, seed=None, init_immun=0.8):
    pars = base_pars(
        lp, pyramid, dur=dur, init_pop=np.array([10000, 10000]),
        r0_scalars=np.array([0.5, 2.0], dtype=np.float32) if r0_scalars is None else r0_scalars, init_immun=init_immun,
        init_prev=init_prev, r0=r0, seasonal_amplitude=0.0, distances=np.array([[0, 1], [1, 0]]), migration_method="gravity",
        gravity_k=0.5, max_migr_frac=0.01, seed=seed if seed is not None else 0,
    )
    sim = lp.SEIR_ABM(pars)
    sim.components = [lp.DiseaseState_ABM, lp.Transmission_ABM, lp.VitalDynamics_ABM]
    return sim


def test_trans_default_mean_matches_force_of_infection(lp, pyramid):
    exposures, sim = [], None
    for rep in range(10):
        sim = tx_sim(lp, pyramid, seed=100 + rep)
        sim.run()
        exposures.append(sim.results.E[1:].sum())
    assert np.all(np.array(exposures) > 0)
    D = np.mean(sim.pars["dur_inf"](1000))
    S, E, I, R = (getattr(sim.results, k)[0] for k in "SEIR")  # noqa: E741
    N = S + E + I + R
    p_inf = 1 - np.exp(-(sim.pars["r0"] / D) * np.array(sim.pars["r0_scalars"]) * I / N)
    exp_E = np.sum(S * p_inf)
    stderr = np.sqrt(S * p_inf * (1 - p_inf)).sum()
    assert abs(np.mean(exposures) - exp_E) < 2 * stderr, (np.mean(exposures), exp_E, stderr)


def test_zero_trans(lp, pyramid):
    for kw in ({"r0": 0}, {"r0_scalars": np.array([0.0, 0.0])}, {"init_prev": 0.0}):
        sim = tx_sim(lp, pyramid, **kw)
        sim.run()
        assert sim.results.E[1:].sum() == 0, kw


def test_linear_transmission_scaling(lp, pyramid):
    def mean_E(**kw):
        out = []
        for seed in range(10):
            sim = tx_sim(lp, pyramid, seed=seed, **kw)
            sim.run()
            out.append(sim.results.E[1:].sum())
        return np.mean(out)

    base = mean_E()
    assert np.isclose(mean_E(r0=28), 2 * base, rtol=0.2)
    assert np.isclose(mean_E(r0_scalars=np.array([1.0, 4.0])), 2 * base, rtol=0.2)
    assert np.isclose(mean_E(init_prev=0.02), 2 * base, rtol=0.2)


# ---------------------------------------------------------------- tests/test_strain_transmission.py:58-82, 283-290, 534
def test_strain_result_shapes_and_totals(lp, pyramid):
    sim = tx_sim(lp, pyramid, dur=20, init_prev=0.02)
    sim.run()
    nt, nn, ns = sim.pars.dur + 1, 2, 3
    for k in ("E_by_strain", "I_by_strain", "new_exposed_by_strain"):
        assert getattr(sim.results, k).shape == (nt, nn, ns) and getattr(sim.results, k).dtype == np.int32
    assert np.array_equal(sim.results.E, sim.results.E_by_strain.sum(axis=2))
    assert np.array_equal(sim.results.I, sim.results.I_by_strain.sum(axis=2))
    assert np.array_equal(sim.results.new_exposed, sim.results.new_exposed_by_strain.sum(axis=2))
    assert sim.results.new_exposed.sum() > 0
    assert sim.results.E_by_strain[:, :, 1:].sum() == 0  # VDPV2 only without vaccination
    # bookkeeping identity: no births here would be S[t] = S[t-1] - new_exposed[t] - deaths of S; with VD on, alive adds up
    alive = np.sum(sim.people.disease_state[: sim.people.count] >= 0)
    assert alive == sim.pars.init_pop.sum() + sim.results.births.sum() - sim.results.deaths.sum()
    assert np.array_equal((sim.results.S + sim.results.E + sim.results.I + sim.results.R).sum(axis=1)[-1:], [alive])


# ---------------------------------------------------------------- tests/test_interventions.py
def vx_sim(lp, pyramid, dur=30, init_pop=None, vx_prob_ri=0.5, vx_prob_ipv=0.75, cbr=None, r0=14, new_pars=None, seed=123):
    pars = base_pars(
        lp, pyramid, start_date=lp.date("2019-01-01"), dur=dur, init_pop=np.array([50000, 50000]) if init_pop is None else init_pop,
        cbr=np.array([30, 25]) if cbr is None else cbr, r0=r0, dur_exp=lp.constant(value=2), dur_inf=lp.constant(value=1),
        vx_prob_ri=vx_prob_ri, vx_prob_ipv=vx_prob_ipv, seed=seed, strain_r0_scalars={0: 1.0, 1: 0.0, 2: 0.0},
        r0_scalars=np.array([0.8, 1.2]),
    )
    pars += new_pars if new_pars is not None else {}
    sim = lp.SEIR_ABM(pars)
    sim.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.SIA_ABM, lp.Transmission_ABM]
    return sim


def test_ri_manually_seeded(lp, pyramid):
    n_vx, dur = 1000, 28
    sim = vx_sim(lp, pyramid, dur=dur, vx_prob_ri=1.0, vx_prob_ipv=1.0)
    sim.people.ri_timer[:n_vx] = np.random.randint(0, dur, n_vx)
    sim.run()
    assert sim.results.ri_vaccinated.sum() >= n_vx and sim.results.ipv_vaccinated.sum() >= n_vx


def test_ri_zero(lp, pyramid):
    sim = vx_sim(lp, pyramid, dur=365, cbr=np.array([0, 0]), vx_prob_ri=1.0, vx_prob_ipv=1.0)
    sim.run()
    assert sim.results.ri_vaccinated[:112].sum() > 0
    assert sim.results.ri_vaccinated[98 + 14 :].sum() == 0 and sim.results.ipv_vaccinated[98 + 14 :].sum() == 0
    sim = vx_sim(lp, pyramid, dur=120, cbr=np.array([300, 250]), vx_prob_ri=0.0, vx_prob_ipv=0.0)
    sim.run()
    assert sim.results.ri_vaccinated.sum() == 0 and sim.results.ipv_vaccinated.sum() == 0


def test_ri_no_effect_on_non_susceptibles(lp, pyramid):
    sim = vx_sim(lp, pyramid, init_pop=np.array([10, 10]), r0=0, vx_prob_ri=1.0)
    sim.people.ri_timer[:20] = 0
    sim.people.disease_state[:5] = 1
    sim.people.disease_state[5:10] = 2
    sim.people.disease_state[10:15] = 3
    sim.run()
    assert np.sum(sim.results.ri_vaccinated) == 20
    assert np.sum(sim.results.new_exposed) == 5


def test_sia_schedule(lp, pyramid):
    sia = {"sia_schedule": [{"date": "2019-01-10", "nodes": [0], "age_range": (0, 5 * 365), "vaccinetype": "nOPV2"}],
           "vx_prob_sia": [0.6, 0.8]}
    sim = vx_sim(lp, pyramid, vx_prob_ri=0, new_pars=sia)
    sim.run()
    day10 = np.sum(sim.results.sia_vaccinated[9, :])
    assert day10 > 0 and np.sum(sim.results.sia_vaccinated) == day10 and sim.results.sia_vaccinated[9, 1] == 0
    assert np.isclose(np.sum(sim.results.sia_protected), day10 * 0.56, atol=100)
    alive0 = (sim.people.node_id == 0) & (sim.people.disease_state >= 0)
    age = sim.t - sim.people.date_of_birth[alive0]
    assert np.isclose(day10, np.sum(age < 5 * 365) * 0.6, atol=500)
    assert np.sum(sim.results.new_exposed) == np.sum(sim.results.sia_protected) > 0
    exposed = sim.people.disease_state == 1
    assert np.all(sim.t - sim.people.date_of_birth[exposed] <= 5 * 365 + 22)
    assert sim.results.sia_new_exposed_by_strain[9, 0, 2] == sim.results.sia_protected[9, 0]


def test_chronically_missed(lp, pyramid):
    sia = {"sia_schedule": [{"date": "2019-01-10", "nodes": [0, 1], "age_range": (0, 5 * 365), "vaccinetype": "perfect"}],
           "vx_prob_sia": [1.0, 1.0], "missed_frac": 0.2}
    sim = vx_sim(lp, pyramid, vx_prob_ri=0, new_pars=sia)
    sim.run()
    missed = sim.people.chronically_missed[: sim.people.count].astype(bool)
    assert np.isclose(missed.sum(), sim.pars.init_pop.sum() * 0.2, atol=100)
    eir = np.isin(sim.people.disease_state[: sim.people.count], [1, 2, 3])
    assert not np.any(eir & missed)
    age = sim.t - sim.people.date_of_birth[: sim.people.count]
    assert np.isclose(eir.sum(), np.sum(age < 5 * 365) * 0.8, atol=500)


# ---------------------------------------------------------------- tests/test_vital_dynamics.py
def vd_sim(lp, pyramid, step_size=1, cbr=None):
    pars = base_pars(lp, pyramid, init_pop=np.array([10000, 5000]), step_size_VitalDynamics_ABM=step_size,
                     cbr=np.array([30, 25]) if cbr is None else cbr)
    sim = lp.SEIR_ABM(pars)
    sim.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.Transmission_ABM]
    return sim


def _fold_day0(a):
    a = a.copy()
    a[1] += a[0]
    return a[1:]


def test_births_generated(lp, pyramid):
    sim = vd_sim(lp, pyramid)
    n0 = sim.people.count
    sim.run()
    assert sim.people.count > n0
    assert n0 + sim.results.births.sum() - sim.results.deaths.sum() == np.sum(sim.people.disease_state > -1)
    dobs = sim.people.date_of_birth[: sim.people.count]
    expected = _fold_day0(np.bincount(dobs[dobs >= 0], minlength=sim.pars.dur + 1))
    assert np.array_equal(expected, _fold_day0(np.sum(sim.results.births, axis=1)))
    assert np.array_equal(sim.results.pop[-1], sim.results.pop[0] + sim.results.births.sum(0) - sim.results.deaths.sum(0))


@pytest.mark.parametrize("step_size", [1, 7])
def test_deaths_occur(lp, pyramid, step_size):
    sim = vd_sim(lp, pyramid, step_size=step_size)
    sim.people.date_of_death[:5] = 1
    sim.run()
    assert np.all(sim.people.date_of_death[:5] == 1) and np.all(sim.people.disease_state[:5] == -1)
    dods = sim.people.date_of_death[: sim.people.count]
    expected = np.bincount(dods[dods >= 0])[: sim.pars.dur + 1]
    observed = np.sum(sim.results.deaths, axis=1)
    if step_size == 1:
        assert np.array_equal(_fold_day0(expected), _fold_day0(observed))
    else:  # deaths are collected on the step days: compare per bin of `step_size` days
        edges = np.arange(0, sim.pars.dur + 1, step_size)
        assert np.array_equal(np.add.reduceat(_fold_day0(expected)[: edges[-1]], edges[:-1]),
                              np.add.reduceat(_fold_day0(observed)[: edges[-1]], edges[:-1]))


def test_no_births_with_zero_cbr(lp, pyramid):
    sim = vd_sim(lp, pyramid, cbr=np.array([0, 0]))
    n0 = sim.people.count
    sim.run()
    assert sim.people.count == n0 and sim.results.births.sum() == 0


# ---------------------------------------------------------------- early stop + seed schedule (model.py:759-795)
def test_seed_schedule_and_early_stop(lp, pyramid):
    pars = base_pars(lp, pyramid, dur=60, init_pop=np.array([2000, 2000]), stop_if_no_cases=True, r0=0,
                     seed_schedule=[{"timestep": 5, "node_id": 1, "prevalence": 50}, {"timestep": 9, "node_id": 0, "prevalence": 0.01}],
                     dur_inf=lp.constant(value=4))
    sim = lp.SEIR_ABM(pars)
    sim.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.Transmission_ABM]
    sim.run()
    assert sim.results.I[4].sum() == 0 and sim.results.I[5, 1] == 50 and sim.results.I[5, 0] == 0
    assert sim.results.I[9, 0] == 20
    assert sim.should_stop and sim.t < sim.nt  # infections burn out (r0 = 0) -> stops early
    assert sim.results.S[sim.t :].sum() == 0  # rows after the stop stay zero


def test_same_seed_same_results(lp, pyramid):
    """tests/test_prng_seeding.py:21-54 (on synthetic inputs): identical seeds give identical result arrays."""
    outs = []
    for seed in (11, 11, 12):
        sim = vx_sim(lp, pyramid, dur=40, init_pop=np.array([20000, 10000]), seed=seed,
                     new_pars={"sia_schedule": [{"date": "2019-01-20", "nodes": [0, 1], "age_range": (0, 5 * 365), "vaccinetype": "nOPV2"}],
                               "vx_prob_sia": [0.5, 0.5]})
        sim.people.disease_state[:200] = 2
        sim.run()
        outs.append({k: getattr(sim.results, k).copy() for k in ("births", "deaths", "paralyzed", "ri_vaccinated", "sia_protected",
                                                                "sia_vaccinated", "S", "E", "I", "R", "new_exposed")})
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k
    assert any(not np.array_equal(outs[0][k], outs[2][k]) for k in outs[0])


# ---------------------------------------------------------------- tests/test_init_pop.py:126-185 (snapshot -> init_from_file)
def test_snapshot_reload_runs_like_a_fresh_sim(lp, pyramid, tmp_path):
    """save_snapshot of a freshly built sim, load_snapshot + the components' init_from_file (reference run_sim.py:388-409):
    the live agents come back identical, and the reloaded sim's epidemic tracks the fresh one (the newborn slots are re-drawn
    at load, so the two are not bitwise equal once cohorts are born -- the reference's test compares totals too)."""
    def pars():
        return base_pars(lp, pyramid, dur=60, init_pop=np.array([20_000, 12_000, 8_000]), cbr=np.array([35.0, 30.0, 25.0]),
                         r0_scalars=np.ones(3), init_prev=0.01, r0=14, seed=5, vx_prob_ri=0.3, vx_prob_ipv=0.3,
                         distances=np.array([[0, 40, 70], [40, 0, 50], [70, 50, 0.0]]), migration_method="gravity", gravity_k=0.5,
                         gravity_a=1, gravity_b=1, gravity_c=2.0, max_migr_frac=0.1)

    np.random.seed(5)
    fresh = lp.SEIR_ABM(pars())
    fresh.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.Transmission_ABM]
    path = tmp_path / "init_pop.h5"
    fresh.people.save_snapshot(path, fresh.results.R[:], fresh.pars)

    p = pars()
    people, R, loaded_pars = lp.LaserFrame.load_snapshot(path, n_ppl=p["init_pop"], cbr=p["cbr"], nt=p["dur"] + 10)
    assert loaded_pars["r0"] == 14 and people.count == fresh.people.count
    sim = lp.SEIR_ABM.init_from_file(people, p)
    ds, vd = lp.DiseaseState_ABM.init_from_file(sim), lp.VitalDynamics_ABM.init_from_file(sim)
    ri, tx = lp.RI_ABM.init_from_file(sim), lp.Transmission_ABM.init_from_file(sim)
    sim.results.R = R
    sim._components = [type(vd), type(ds), type(ri), type(tx)]
    sim.instances = [vd, ds, ri, tx]
    n = fresh.people.count
    for name, col in fresh.people.columns().items():
        assert np.array_equal(col[:n], getattr(sim.people, name)[:n]), name
    assert (sim.people.disease_state[n:] == -1).all() and sim.people.acq_risk_multiplier[n:].min() > 0
    fresh.run()
    sim.run()
    a, b = fresh.results, sim.results
    assert np.array_equal(a.S[0], b.S[0]) and np.array_equal(a.I[0], b.I[0])
    tot_a, tot_b = a.new_exposed.sum(), b.new_exposed.sum()
    assert tot_a > 500 and abs(tot_a - tot_b) < 0.1 * tot_a
    assert abs(int(a.births.sum()) - int(b.births.sum())) < 0.2 * a.births.sum() + 10


# ---------------------------------------------------------------- tests_scientific/outbreak_size.py:60-80 (analytic anchor)
@pytest.mark.parametrize("r0", [1.5, 2.5, 5.0])
def test_final_size_matches_kermack_mckendrick(lp, pyramid, r0):
    """Single node, everybody susceptible, no heterogeneity, no exposed period, no births / deaths / vaccination / season:
    the total number of infections must solve the final-size relation z = S0 (1 - exp(-R0 (z + I0))) the reference's
    scientific test compares against.  (The fused engine runs every day.)"""
    from scipy.optimize import brentq

    n, i0 = 300_000, 60
    sim = lp.SEIR_ABM(base_pars(lp, pyramid, dur=1200, init_pop=np.array([n]), cbr=np.array([0.0]), r0_scalars=np.array([1.0]), r0=r0,
                                init_prev=i0 / n, init_immun=0.0, seasonal_amplitude=0.0, dur_exp=lp.constant(value=0),
                                individual_heterogeneity=False, vx_prob_ri=None, vx_prob_sia=None, seed=int(10 * r0),
                                distances=np.zeros((1, 1))))
    sim.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.Transmission_ABM]
    # the reproduction number the model implements: infectivity per day x days infectious.  The infectious period is the
    # gamma draw truncated to whole int8 days (model.py:575-579: half a day short of its mean) and the daily infectivity is
    # r0 over a 1000-draw sample mean (model.py:845), so R differs from pars.r0 by ~2 %
    r_eff = float(sim.people.daily_infectivity[0]) * float(sim.people.infection_timer[: sim.people.count].mean())
    assert abs(r_eff / r0 - 1) < 0.06
    sim.run()
    total = sim.results.new_exposed.sum() / n
    expected = brentq(lambda z: z - (1 - np.exp(-r_eff * (z + i0 / n))), 1e-6, 1.0)
    assert abs(total - expected) < 0.02, (r0, r_eff, total, expected)
    assert sim.results.I[-1].sum() == 0  # the outbreak is over inside the window


# ---------------------------------------------------------------- tests/test_migration.py:13-72 (radiation spread is monotone in k)
def test_radiation_spread_increases_with_k(lp, pyramid):
    n_nodes = 14
    rs = np.random.RandomState(2)
    xy = rs.uniform(0, 200, (n_nodes, 2))
    d = np.sqrt(((xy[:, None] - xy[None]) ** 2).sum(-1))

    def nodes_reached(k_log10, seed):
        sim = lp.SEIR_ABM(base_pars(lp, pyramid, dur=30, init_pop=np.full(n_nodes, 8000), cbr=np.zeros(n_nodes), r0_scalars=np.ones(n_nodes),
                                    init_prev=[0.01] + [0.0] * (n_nodes - 1), r0=14, distances=d, migration_method="radiation",
                                    radiation_k_log10=k_log10, max_migr_frac=1.0, vx_prob_ri=None, vx_prob_sia=None, seed=seed))
        # VitalDynamics_ABM maintains results.pop, the denominator of the force of infection (without it the reference divides
        # by max(0, 1), model.py:1344-1347, and the home node's outbreak is 8000 times stronger)
        sim.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.Transmission_ABM]
        sim.run()
        return np.count_nonzero(sim.results.I.sum(axis=0) > 0)

    zero, low, high = (np.mean([nodes_reached(k, 1000 + i) for i in range(4)]) for k in (-9, -2.5, -1))
    assert zero == 1.0, zero          # no migration: the infection stays in its home node
    assert low > zero and high >= low + 2, (zero, low, high)


# ---------------------------------------------------------------- LaserFrame.add raising when a cohort does not fit (laser-core)
def test_cohort_that_does_not_fit_raises_after_results_are_back(lp, pyramid):
    """The reference's LaserFrame.add raises at the tick whose cohort exceeds the capacity.  Births are created on the
    device here, so the overflow is a sticky device flag: no later cohort is created either, the host sees the flag one call
    late, and run() raises ValueError naming the tick -- after the columns and results were copied back."""
    sim = lp.SEIR_ABM(base_pars(lp, pyramid, dur=40, init_pop=np.array([30_000, 20_000]), cbr=np.array([30.0, 25.0]), init_prev=0.01,
                                distances=np.array([[0, 50], [50, 0.0]])))
    sim.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.Transmission_ABM]
    vd = next(i for i in sim.instances if type(i).__name__ == "VitalDynamics_ABM")
    room = sim.people.capacity - sim.people.count
    vd.birth_rate[:] = 0.4 * room / (7 * 50_000)  # each cohort takes ~40 % of the room: the third one (t = 21) no longer fits
    with pytest.raises(ValueError, match="exceeds capacity .* at tick 21"):
        sim.run()
    assert sim.dev is None and sim.results.S[30].sum() > 0 and sim.results.births[14].sum() > 0
    assert sim.results.births[21].sum() == 0 and sim.results.births[28].sum() == 0  # sticky: nobody is created afterwards
    assert sim.people.count <= sim.people.capacity


def test_network_is_read_only_while_resident(lp, pyramid):
    """tx.network is re-read every tick by the reference (model.py:1335); the device holds a copy, so an in-place edit after
    to_device() raises instead of being silently ignored, and replacing the array is picked up."""
    sim = lp.SEIR_ABM(base_pars(lp, pyramid, dur=10, init_pop=np.array([5_000, 5_000]), init_prev=[0.02, 0.0], r0=14,
                                distances=np.array([[0, 50], [50, 0.0]])))
    sim.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.Transmission_ABM]
    tx = next(i for i in sim.instances if type(i).__name__ == "Transmission_ABM")
    sim.run_ticks(3)
    with pytest.raises(ValueError):
        tx.network[0, 1] = 0.5
    tx.network = np.zeros((2, 2))  # cut the nodes off from now on
    sim.run_ticks(20)
    sim.to_host()
    tx.network[0, 1] = 0.0  # writable again once the population is back on the host
    assert sim.t == sim.nt
