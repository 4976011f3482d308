"""TEST INFRASTRUCTURE: the fused engine (SEIR_ABM.step_tick -> lpk_vd_births / lpk_tick_pass / lpk_tick_node) held to the
CPU oracle's tick loop (oracle/tick_loop.py) end to end -- not to the CUDA component kernels.  Used by
tests/test_gpu_engine_vs_oracle.py and __graft_entry__.smoke(); nothing in the product imports it.

The run is a synthetic full-feature workload (laser_polio_b200.synth.synth_sim: vital dynamics every 7 ticks, RI every
14, campaign days, gravity network, three strains).  Every results array and every agent column must be bit-identical;
the node-level exposure scale tau of each tick is taken from the device (the per-agent trial depends on it bit for bit)
and checked on its own against the float64 restatement of the node math at 2e-6.
"""

from __future__ import annotations

import numpy as np

from . import oracle as orc
from . import tick_loop

RESULT_ROWS = ("S", "E", "I", "R", "new_exposed", "births", "deaths", "pop", "new_potentially_paralyzed", "new_paralyzed",
               "potentially_paralyzed", "paralyzed", "ri_vaccinated", "ri_protected", "ipv_vaccinated", "sia_vaccinated", "sia_protected",
               "E_by_strain", "I_by_strain", "new_exposed_by_strain", "ri_new_exposed_by_strain", "sia_new_exposed_by_strain")


def run_and_compare(n_agents, n_nodes, dur, seed, cbr=37.0, node_math_ticks=(1, 2, 7, 14), pars_over=None, pop_over=None, compact_every=0):
    """Runs ``dur`` ticks of the fused engine and of the oracle on the same table; raises AssertionError on the first
    difference.  Returns a summary dict (counts that show the run was not trivial)."""
    from laser_polio_b200 import kernels as K
    from laser_polio_b200 import synth, utils

    pars_over = dict(pars_over or {})
    if compact_every:
        pars_over["compact_every"] = int(compact_every)
    sim, n = synth.synth_sim(n_agents, n_nodes, dur, seed, cbr=cbr, pars_over=pars_over, pop_over=pop_over)
    people, pars = sim.people, sim.pars
    cols = {name: col.copy() for name, col in people.columns().items()}
    count0, cap, ns = people.count, people.capacity, len(pars.strain_ids)
    by = {type(i).__name__: i for i in sim.instances}
    vd, tx, ri, sia = by["VitalDynamics_ABM"], by["Transmission_ABM"], by["RI_ABM"], by["SIA_ABM"]
    R0_rows = sim.results.R.copy()  # pre-seeded non-agent immunes (zero for a table of agents only)

    # ---- the device run, recording the node-level outputs of every tick
    K.STATS.reset()
    taus, cdfs = {}, {}
    sim.to_device()
    for t in range(sim.nt):
        sim.step_tick(t)
        if t >= 1:
            eng = sim._engine
            assert eng, "the fused engine did not engage"
            if eng.pending:  # a fused tick: its node kernel left tau / cdf for the exposure the next pass applies
                taus[t], cdfs[t] = eng.q.cpu().numpy().copy(), eng.cdf.cpu().numpy().copy()
            else:  # a day handed to the components
                taus[t], cdfs[t] = sim.dev.node_out[0].cpu().numpy().copy(), sim.dev.node_out[1].cpu().numpy().copy()
    calls = dict(K.STATS.calls)
    sim.to_host()
    assert calls.get("tick_pass", 0) >= sim.nt - 1 - 4, f"fused passes: {calls}"

    # ---- the oracle run
    events = {}
    for t, evs in sia._by_tick.items():
        out = []
        for e in evs:
            targeted = np.zeros(n_nodes, np.uint8)
            targeted[np.asarray(e["nodes"], dtype=np.int64)] = 1
            vtype = e["vaccinetype"]
            out.append((targeted, np.asarray(pars.vx_prob_sia, np.float32), float(pars.vx_efficacy[vtype]), int(e["age_range"][0]),
                        int(e["age_range"][1]), 2 if "nOPV" in vtype else 1))
        events[t] = out
    r, count, tallies = tick_loop.run(
        cols, count0, cap, n_nodes, ns, sim.nt, seed=int(pars.seed), id_base=sim.id_base,
        strain_r0_scalars=list(pars.strain_r0_scalars.values())[:ns], p_paralysis=float(np.float32(pars.p_paralysis)),
        tau_of_tick=taus, cdf_of_tick=cdfs, vd_step=int(vd.step_size), birth_rate=vd.birth_rate,
        cum_deaths=np.asarray(vd.death_estimator._cd, np.int64), pop0=np.asarray(pars.init_pop, np.int32),
        ri_step=int(ri.step_size), vx_prob_ri=np.asarray(pars.vx_prob_ri, np.float64), vx_prob_ipv=np.asarray(pars.vx_prob_ipv, np.float64),
        ri_strain=1, sia_events=events, compact_every=int(compact_every))

    # ---- bit-exact: head count, every results array, every agent column
    assert count == people.count, f"count {people.count} vs oracle {count}"
    for name in RESULT_ROWS:
        want = r[name] + (R0_rows if name == "R" else 0)
        got = getattr(sim.results, name)
        if not np.array_equal(got, want):
            bad = np.argwhere(got != want)[0]
            raise AssertionError(f"results.{name} differs first at {tuple(bad)}: device {got[tuple(bad)]} oracle {want[tuple(bad)]}")
    for name, col in people.columns().items():
        if not np.array_equal(col[:count], cols[name][:count]):
            i = int(np.flatnonzero(col[:count] != cols[name][:count])[0])
            raise AssertionError(f"people.{name} differs first at agent {i}: device {col[i]} oracle {cols[name][i]}")

    # ---- the node math of a few ticks: device tau vs the float64 restatement on the oracle's own tallies
    W = np.asarray(tx.network, np.float64)
    for t in node_math_ticks:
        if t >= sim.nt:
            continue
        bfx, efx, hist = tallies[t]
        sim.t = t
        season = float(utils.get_seasonality(sim))
        q_o, cdf_o, _, _ = orc.tx_node_math_device(bfx, efx, hist, W, season, np.asarray(tx.r0_scalars, np.float64), r["pop"][t],
                                                   float(pars.node_seeding_zero_inflation), float(pars.node_seeding_dispersion), int(pars.seed), t)
        np.testing.assert_allclose(taus[t], q_o, rtol=2e-6, atol=1e-30, err_msg=f"tau of tick {t}")
        np.testing.assert_allclose(cdfs[t], cdf_o, rtol=1e-6, atol=1e-12, err_msg=f"strain cdf of tick {t}")
    sim.t = sim.nt
    res = sim.results
    return {
        "agents": n, "nodes": n_nodes, "ticks": sim.nt, "final_count": int(count), "cohort_share": float(count - count0) / float(count),
        "new_exposed": int(res.new_exposed.sum()), "deaths": int(res.deaths.sum()), "births": int(res.births.sum()),
        "ri_vaccinated": int(res.ri_vaccinated.sum()), "sia_protected": int(res.sia_protected.sum()),
        "new_potentially_paralyzed": int(res.new_potentially_paralyzed.sum()), "calls": calls,
        "compactions": int(getattr(sim, "_compactions", 0)),
    }
