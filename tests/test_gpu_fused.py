"""The fused tick (lpk_tick_pass + lpk_tick_node, engine.FusedEngine) must reproduce the component-by-component
path bit for bit: every results array and every agent column, on schedules that mix fused days with days the
engine hands back to the components (vital dynamics, SIA campaigns, seed_schedule injections, RI)."""

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
    path = tmp_path_factory.mktemp("data") / "pyramid.csv"
    rows = ["Age,M,F"] + [f"{5 * k}-{5 * k + 4},{int(1.7e7 * np.exp(-0.16 * k))},{int(1.6e7 * np.exp(-0.16 * k))}" for k in range(20)]
    rows.append("100+,300,500")
    path.write_text("\n".join(rows) + "\n")
    return str(path)


def make(lp, pyramid, fused, components, **over):
    n_nodes = over.pop("n_nodes", 5)
    pop_lo, pop_hi = over.pop("pop_range", (3000, 30000))
    rs = np.random.RandomState(3)
    init_pop = rs.randint(pop_lo, pop_hi, n_nodes)
    d = rs.uniform(5, 300, (n_nodes, n_nodes))
    d = (d + d.T) / 2
    np.fill_diagonal(d, 0)
    p = {
        "start_date": lp.date("2019-01-01"), "dur": 45, "init_pop": init_pop, "cbr": np.full(n_nodes, 35.0),
        "r0_scalars": rs.uniform(0.5, 1.5, n_nodes), "age_pyramid_path": pyramid, "init_immun": 0.3,
        "init_prev": [0.01] + [0.0] * (n_nodes - 1), "r0": 14, "distances": d, "stop_if_no_cases": False, "verbose": 0, "seed": 99,
        "vx_prob_ri": rs.uniform(0.3, 0.9, n_nodes), "vx_prob_ipv": rs.uniform(0.3, 0.9, n_nodes), "missed_frac": 0.1,
        "p_paralysis": 0.3, "node_seeding_zero_inflation": 0.2, "node_seeding_dispersion": 2,
        "sia_schedule": [
            {"date": "2019-01-10", "nodes": [0, 2], "age_range": (0, 5 * 365), "vaccinetype": "nOPV2"},
            {"date": "2019-01-10", "nodes": [1, 2, 3], "age_range": (0, 10 * 365), "vaccinetype": "mOPV2"},
            {"date": "2019-01-29", "nodes": list(range(n_nodes)), "age_range": (0, 5 * 365), "vaccinetype": "mOPV2"},
        ],
        "vx_prob_sia": rs.uniform(0.4, 0.9, n_nodes).tolist(),
        "seed_schedule": [{"timestep": 12, "node_id": 3, "prevalence": 40}, {"timestep": 28, "node_id": 1, "prevalence": 0.002}],
    }
    p.update(over)
    sim = lp.SEIR_ABM(lp.PropertySet(p))
    sim.components = components
    sim.fused = fused
    return sim


def run_pair(lp, pyramid, components, **over):
    out = []
    for fused in (False, True):
        sim = make(lp, pyramid, fused, components, **over)
        if hasattr(sim.people, "ri_timer"):
            k = len(sim.people.ri_timer[: sim.people.count : 3])
            sim.people.ri_timer[: sim.people.count : 3] = np.random.RandomState(1).randint(-10, 40, k)
        from laser_polio_b200 import kernels as K

        K.STATS.reset()
        sim.run()
        out.append((sim, dict(K.STATS.calls)))
    return out


def assert_identical(a, b):
    for name, arr in a.results.__dict__.items():
        if isinstance(arr, np.ndarray):
            assert np.array_equal(arr, getattr(b.results, name)), f"results.{name}"
    assert a.people.count == b.people.count
    for name, col in a.people.columns().items():
        assert np.array_equal(col, getattr(b.people, name)), f"people.{name}"


def test_fused_equals_components_full_feature_set(lp, pyramid):
    comps = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.SIA_ABM, lp.Transmission_ABM]
    (ref, calls_ref), (fus, calls_fus) = run_pair(lp, pyramid, comps)
    assert calls_fus.get("tick_pass", 0) >= 30 and "tick_pass" not in calls_ref  # the fused path really ran
    assert_identical(ref, fus)
    r = fus.results
    assert r.new_exposed.sum() > 500 and r.ri_vaccinated.sum() > 0 and r.sia_protected.sum() > 0 and r.new_potentially_paralyzed.sum() > 0
    assert r.births.sum() > 0 and r.deaths.sum() > 0
    assert np.array_equal(r.E, r.E_by_strain.sum(axis=2)) and np.array_equal(r.I, r.I_by_strain.sum(axis=2))
    # incremental paralysis census == cumulative new minus the dead (cross-check against the agent table)
    alive = fus.people.disease_state[: fus.people.count] >= 0
    assert r.potentially_paralyzed[-1].sum() == np.sum((fus.people.potentially_paralyzed[: fus.people.count] == 1) & alive)
    assert r.paralyzed[-1].sum() == np.sum((fus.people.paralyzed[: fus.people.count] == 1) & alive)


def test_fused_equals_components_large_nodes(lp, pyramid):
    """Nodes of 120-260 K agents: most 32 K-agent chunks lie inside one node and take the pass's single-node loop (one
    Philox block per 8 agents, high-half pre-test); node boundaries, newborn cohorts and RI ticks take the general rows."""
    comps = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.SIA_ABM, lp.Transmission_ABM]
    (ref, _), (fus, calls) = run_pair(lp, pyramid, comps, n_nodes=4, pop_range=(120_000, 260_000), dur=40, r0=6,
                                      init_prev=[0.004, 0.0, 0.001, 0.0])
    assert calls.get("tick_pass", 0) >= 30
    assert_identical(ref, fus)
    r = fus.results
    assert r.new_exposed.sum() > 20_000 and r.deaths.sum() > 0 and r.births.sum() > 0 and r.ri_vaccinated.sum() > 0
    assert (r.R[-1] > r.R[0]).all() and np.array_equal(r.E, r.E_by_strain.sum(axis=2))


SIA_DAYS = [  # single-event days: fused inside the pass; tick 7 = vital dynamics, tick 14 = vital dynamics + RI, tick 19 plain
    {"date": "2019-01-08", "nodes": [0, 2], "age_range": (0, 5 * 365), "vaccinetype": "nOPV2"},
    {"date": "2019-01-15", "nodes": [1, 2, 3], "age_range": (200, 10 * 365), "vaccinetype": "mOPV2"},
    {"date": "2019-01-20", "nodes": [0, 1, 2, 3], "age_range": (0, 15 * 365), "vaccinetype": "nOPV2"},
    {"date": "2019-01-24", "nodes": [3], "age_range": (0, 5 * 365), "vaccinetype": "mOPV2"},   # two events on one day:
    {"date": "2019-01-24", "nodes": [0, 3], "age_range": (0, 5 * 365), "vaccinetype": "nOPV2"},  # handed to the components
]


def test_fused_sia_days_small_nodes(lp, pyramid):
    """Campaign days run inside the pass (LPK_F_SIA) on nodes of a few thousand agents: general rows and mixed quads."""
    comps = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.SIA_ABM, lp.Transmission_ABM]
    (ref, _), (fus, calls) = run_pair(lp, pyramid, comps, sia_schedule=[dict(e) for e in SIA_DAYS], dur=35)
    assert calls.get("tick_pass", 0) >= 30 and calls.get("fast_sia", 0) == 2  # only the two-event day left the pass
    assert_identical(ref, fus)
    r = fus.results
    assert (r.sia_protected[[7, 14, 19]].sum(axis=1) > 0).all() and r.sia_vaccinated.sum() > r.sia_protected.sum()


def test_fused_sia_days_large_nodes(lp, pyramid):
    """The same on nodes of 120-260 K agents: the streaming loop's SIA rows (TMA-staged date_of_birth, ring entries)."""
    comps = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.SIA_ABM, lp.Transmission_ABM]
    (ref, _), (fus, calls) = run_pair(lp, pyramid, comps, n_nodes=4, pop_range=(120_000, 260_000), dur=30, r0=6,
                                      init_prev=[0.004, 0.0, 0.001, 0.0], sia_schedule=[dict(e) for e in SIA_DAYS], seed_schedule=None)
    assert calls.get("tick_pass", 0) >= 28
    assert_identical(ref, fus)
    r = fus.results
    assert r.sia_protected.sum() > 50_000 and r.ri_vaccinated.sum() > 0 and r.deaths.sum() > 0


def test_fused_equals_components_saturating_force_of_infection(lp, pyramid):
    """r0 = 999 (the reference's own 'everybody gets exposed' regime, tests/test_diseasestate_abm.py:258): tau is huge or
    'everybody', so the high-half pre-test passes for every susceptible and the exact path decides."""
    comps = [lp.DiseaseState_ABM, lp.Transmission_ABM]
    (ref, _), (fus, _) = run_pair(lp, pyramid, comps, n_nodes=3, pop_range=(70_000, 90_000), dur=12, r0=999, sia_schedule=None,
                                  seed_schedule=None, vx_prob_ri=None, init_immun=0.0, init_prev=[0.01, 0.01, 0.01])
    assert_identical(ref, fus)
    assert fus.results.S[-1].sum() < 0.005 * fus.results.S[0].sum()  # all but a few agents of negligible risk


def test_fused_equals_components_transmission_only_many_nodes(lp, pyramid):
    # no vital dynamics: every tick after 0 is fused; pop[t] stays 0 so the rate denominator is max(0, 1) like the reference
    comps = [lp.DiseaseState_ABM, lp.Transmission_ABM]
    (ref, _), (fus, calls) = run_pair(lp, pyramid, comps, n_nodes=40, dur=30, sia_schedule=None, seed_schedule=None, vx_prob_ri=None,
                                      r0=0.002)
    assert calls.get("tick_pass", 0) == 30
    assert_identical(ref, fus)
    assert fus.results.new_exposed.sum() > 100


def test_fused_equals_components_ri_without_vd_or_sia(lp, pyramid):
    comps = [lp.DiseaseState_ABM, lp.RI_ABM, lp.Transmission_ABM, lp.VitalDynamics_ABM]
    (ref, _), (fus, calls) = run_pair(lp, pyramid, comps, sia_schedule=None, seed_schedule=None, step_size_VitalDynamics_ABM=30,
                                      cbr=np.zeros(5))
    assert calls.get("tick_pass", 0) >= 40
    assert_identical(ref, fus)
    assert fus.results.ri_vaccinated.sum() > 0


def test_fused_early_stop_rule(lp, pyramid):
    """pars.stop_if_no_cases (the reference's default, model.py:789-795): the fused engine stops on the same tick as the
    component path -- the tick after the last exposed / infectious agent is gone -- without a device sync per tick."""
    comps = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.Transmission_ABM]
    (ref, _), (fus, calls) = run_pair(lp, pyramid, comps, n_nodes=3, dur=200, r0=0.0001, sia_schedule=None, seed_schedule=None,
                                      vx_prob_ri=None, init_prev=[0.003, 0.0, 0.001], stop_if_no_cases=True)
    assert calls.get("tick_pass", 0) >= 20
    assert ref.should_stop and fus.should_stop and ref.t == fus.t and 20 < fus.t < fus.nt
    assert_identical(ref, fus)
    assert fus.results.E[fus.t - 2].sum() + fus.results.I[fus.t - 2].sum() == 0  # the census that triggered the stop
    # a pending seed_schedule event keeps the run alive past the extinction
    (ref2, _), (fus2, _) = run_pair(lp, pyramid, comps, n_nodes=3, dur=200, r0=0.0001, sia_schedule=None, vx_prob_ri=None,
                                    seed_schedule=[{"timestep": 190, "node_id": 1, "prevalence": 5}], init_prev=[0.003, 0.0, 0.001],
                                    stop_if_no_cases=True)
    assert ref2.t == fus2.t and fus2.t > 190
    assert_identical(ref2, fus2)


def test_step_tick_resume_after_to_host(lp, pyramid):
    """to_host() mid-run drains the pipeline; resuming gives the same answer as one uninterrupted run."""
    comps = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.SIA_ABM, lp.Transmission_ABM]
    whole = make(lp, pyramid, True, comps)
    whole.run()
    parts = make(lp, pyramid, True, comps)
    for t in range(0, 20):
        parts.step_tick(t)
    parts.to_host()
    snapshot = parts.results.S[19].copy()
    for t in range(20, parts.nt):
        parts.step_tick(t)
    parts.to_host()
    assert np.array_equal(snapshot, whole.results.S[19])
    assert_identical(whole, parts)


def test_graph_launched_spans_equal_stream_launched(lp, pyramid):
    """lpk_run_days with lpk_run.graph: every span of fused days captured into one CUDA graph -- same results."""
    comps = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.SIA_ABM, lp.Transmission_ABM]
    plain = make(lp, pyramid, True, comps, sia_schedule=[dict(e) for e in SIA_DAYS], dur=35)
    plain.run()
    graph = make(lp, pyramid, True, comps, sia_schedule=[dict(e) for e in SIA_DAYS], dur=35)
    graph.cuda_graph = True
    graph.run()
    assert_identical(plain, graph)
    assert graph.results.sia_protected.sum() > 0 and graph.results.births.sum() > 0


def test_run_ticks_in_pieces_equals_one_run(lp, pyramid):
    """run_ticks(n) leaves the population resident: spans of 1, 3 and the rest give the same answer as run()."""
    comps = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.SIA_ABM, lp.Transmission_ABM]
    whole = make(lp, pyramid, True, comps)
    whole.run()
    parts = make(lp, pyramid, True, comps)
    for n in (1, 1, 3, 9, 1, 100):
        parts.run_ticks(n)
    parts.to_host()
    assert parts.t == whole.t
    assert_identical(whole, parts)
