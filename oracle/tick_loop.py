"""TEST INFRASTRUCTURE: the reference's tick loop (SEIR_ABM.run, model.py:246-289) on the CPU oracle, stage by stage on
canonical numpy columns -- the checker the fused engine is held to end to end (tests/test_gpu_engine_vs_oracle.py,
__graft_entry__.smoke()).  Only tests/, smoke() and bench.py's CPU legs may import anything under oracle/.

Per tick t >= 1, in ``default_run_order`` (pars.py:99):
    VitalDynamics (t % vd_step == 0): get_deaths (model.py:1767-1781), births (1711-1734, device Philox scheme), pop row
    DiseaseState:  disease_state_step (344-454)
    RI (t % ri_step == 0): fast_ri (1805-1855)
    SIA (campaign days): fast_sia (1994-2060), one call per event
    Transmission: tx_step_prep (932-1007) -> node math -> exposure (1010-1149 replacement) ; log: count_SEIRP (869-929)
The node-level exposure scale tau / strain cdf of every tick are INPUTS (the device's own, recorded by the caller): the
per-agent trial depends on tau bit for bit, and the float64 node math is checked separately against
``oracle.tx_node_math_device`` at its own tolerance.
"""

from __future__ import annotations

import numpy as np

from . import oracle as orc

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)


def philox_np(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint32 arrays (vectorised restatement of lp_oracle.c's block function)."""
    c0, c1, c2, c3 = (np.asarray(x, np.uint32).copy() for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    for _ in range(10):
        p0 = c0.astype(np.uint64) * _M0
        p1 = c2.astype(np.uint64) * _M1
        n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
        n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
        c1, c3 = p1.astype(np.uint32), p0.astype(np.uint32)
        c0, c2 = n0, n2
        k0 = np.uint32((int(k0) + 0x9E3779B9) & 0xFFFFFFFF)
        k1 = np.uint32((int(k1) + 0xBB67AE85) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _u53(hi, lo):
    return ((hi.astype(np.uint64) << np.uint64(32) | lo.astype(np.uint64)) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def births(pop_prev, birth_rate, step_size, cum_deaths, count, capacity, seed, tick, id_base=0, max_year=100):
    """Vectorised form of ``oracle.vd_births_device`` (same draws, same results; that one is the scalar restatement the
    small GPU test uses).  Returns (births[nodes], node_id[total], date_of_death[total], new_count, status)."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    n = len(pop_prev)
    expected = float(step_size) * np.asarray(birth_rate, np.float64) * np.asarray(pop_prev, np.float64)
    whole = expected.astype(np.int64)
    x = philox_np(np.arange(n, dtype=np.uint32), 0, tick, orc.STAGE_BIRTH, k0, k1)
    b = np.maximum(whole + (_u53(x[0], x[1]) < expected - whole), 0).astype(np.int32)
    total = int(b.sum())
    if count + total > capacity:
        return np.zeros(n, np.int32), np.zeros(0, np.int16), np.zeros(0, np.int32), count, 1
    cd = np.asarray(cum_deaths, np.int64)
    tot = max(int(cd[max_year + 1]), 1)
    g = np.arange(count, count + total, dtype=np.uint64) + np.uint64(id_base)
    x = philox_np((g & np.uint64(0xFFFFFFFF)).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32), tick, orc.STAGE_LIFESPAN, k0, k1)
    draw = 1 + np.floor(_u53(x[0], x[1]) * tot).astype(np.int64)
    yod = np.clip(np.searchsorted(cd[: max_year + 2], draw, side="left") - 1, 0, max_year)
    u2 = _u53(x[2], x[3])
    doy = np.where(yod == 0, 1 + np.floor(u2 * 364.0), np.floor(u2 * 365.0)).astype(np.int64)
    dod = (tick + yod * 365 + doy).astype(np.int32)
    return b, np.repeat(np.arange(n, dtype=np.int16), b), dod, count + total, 0


def run(cols, count, capacity, n_nodes, ns, ticks, *, seed, id_base=0, strain_r0_scalars, p_paralysis, tau_of_tick=None, cdf_of_tick=None,
        vd_step=0, birth_rate=None, cum_deaths=None, pop0=None, ri_step=0, vx_prob_ri=None, vx_prob_ipv=None, ri_strain=1,
        sia_events=None, node_lo=0, node_hi=None, node_math=None, compact_every=0):
    """Ticks 0 .. ticks-1 on the canonical columns ``cols`` (mutated in place).  Returns (results dict of [ticks, nodes(, ns)]
    int32 rows, final count, per-tick tallies {t: (beta_fx, exposure_fx, risk_hist)} for the node-math check).

    sia_events: {tick: [(targeted uint8[nodes], vx_prob float32[nodes], vx_eff, min_age, max_age, strain), ...]}
    compact_every: the device's table compaction (laser_polio_b200.device.DeviceState.compact) before every tick t > 1 with
                t % compact_every == 0: live agents stably sorted by node, then the unborn slots, then the newly dead, which
                leave the swept range.  The returned columns are back in the reference's order and ``count`` is the
                reference's (every agent ever created).
    node_math:  instead of tau_of_tick / cdf_of_tick (pass None for both), the inputs of the node step -- {"network", "r0_scalars",
                "season": callable(t) or float, "zero_inflation", "dispersion"} -- and tau / cdf come from
                ``oracle.tx_node_math_device`` on the loop's own tallies (a free-running oracle simulation)."""
    i32 = lambda *s: np.zeros(s, np.int32)  # noqa: E731
    names2 = ("S", "E", "I", "R", "new_exposed", "births", "deaths", "pop", "new_potentially_paralyzed", "new_paralyzed",
              "potentially_paralyzed", "paralyzed", "ri_vaccinated", "ri_protected", "ipv_vaccinated", "sia_vaccinated", "sia_protected")
    names3 = ("E_by_strain", "I_by_strain", "new_exposed_by_strain", "ri_new_exposed_by_strain", "sia_new_exposed_by_strain")
    r = {k: i32(ticks, n_nodes) for k in names2}
    r.update({k: i32(ticks, n_nodes, ns) for k in names3})
    if pop0 is not None:
        r["pop"][0] = pop0
    srs = np.asarray(strain_r0_scalars, np.float64)
    sia_events = sia_events or {}
    tallies = {}
    c = cols

    def census(t):
        S, E, I, R, Ebs, Ibs, PP, Pz = orc.count_SEIRP(c["node_id"], c["disease_state"], c["strain"], c["potentially_paralyzed"],  # noqa: E741
                                                         c["paralyzed"], n_nodes, ns, count)
        r["S"][t], r["E"][t], r["I"][t], r["E_by_strain"][t], r["I_by_strain"][t] = S, E, I, Ebs, Ibs
        r["R"][t] += R
        r["potentially_paralyzed"][t], r["paralyzed"][t] = PP, Pz

    cap_eff, graveyard = capacity, 0
    orig = np.arange(capacity, dtype=np.int64)
    census(0)
    for t in range(1, ticks):
        if compact_every and t > 1 and t % compact_every == 0:
            key = np.full(cap_eff, n_nodes, np.int64)
            key[:count] = np.where(c["disease_state"][:count] >= 0, c["node_id"][:count], n_nodes + 1)
            live = int((key[:count] < n_nodes).sum())
            perm = np.argsort(key, kind="stable")
            for name in c:
                c[name][:cap_eff] = c[name][:cap_eff][perm]
            orig[:cap_eff] = orig[:cap_eff][perm]
            dead = count - live
            cap_eff, graveyard, count = cap_eff - dead, graveyard + dead, live
        if vd_step:
            if t % vd_step == 0:
                dying = i32(n_nodes)
                orc.get_deaths(n_nodes, count, c["disease_state"], c["node_id"], c["date_of_death"], t, dying)
                rate = np.asarray(birth_rate, np.float64).copy()
                if node_hi is not None:  # a node shard creates the cohorts of its own nodes only
                    rate[:node_lo] = 0.0
                    rate[node_hi:] = 0.0
                b, nid, dod, new_count, status = births(r["pop"][t - 1], rate, vd_step, cum_deaths, count, cap_eff, seed, t, id_base)
                assert status == 0, "oracle births: capacity"
                c["node_id"][count:new_count], c["date_of_death"][count:new_count] = nid, dod
                c["date_of_birth"][count:new_count], c["disease_state"][count:new_count] = t, 0
                count = new_count
                r["births"][t], r["deaths"][t] = b, dying
                r["pop"][t] = r["pop"][t - 1] + b - dying
            else:
                r["pop"][t] = r["pop"][t - 1]
        orc.disease_state_step(c["node_id"], n_nodes, c["disease_state"], c["strain"], count, c["exposure_timer"], c["infection_timer"],
                               c["potentially_paralyzed"], c["paralyzed"], c["ipv_protected"], c["paralysis_timer"], p_paralysis,
                               r["new_potentially_paralyzed"][t], r["new_paralyzed"][t], seed=seed, tick=t, id_base=id_base)
        if ri_step and t % ri_step == 0:
            a, p_, v = i32(n_nodes), i32(n_nodes), i32(n_nodes)
            orc.fast_ri(ri_step, c["node_id"], c["disease_state"], c["strain"], c["ipv_protected"], c["ri_timer"], t, vx_prob_ri, vx_prob_ipv,
                        count, a, p_, v, c["chronically_missed"], ri_strain, seed=seed, tick=t, id_base=id_base)
            r["ri_vaccinated"][t], r["ri_protected"][t], r["ipv_vaccinated"][t] = a, p_, v
            r["new_exposed"][t] += p_
            r["new_exposed_by_strain"][t, :, ri_strain] += p_
            r["ri_new_exposed_by_strain"][t, :, ri_strain] = p_
        for k, (targeted, vx_prob, eff, lo, hi, strain) in enumerate(sia_events.get(t, [])):
            a, p_ = i32(n_nodes), i32(n_nodes)
            orc.fast_sia(c["node_id"], c["disease_state"], c["strain"], c["date_of_birth"], t, vx_prob, eff, count, targeted, lo, hi, a, p_,
                         c["chronically_missed"], strain, seed=seed, tick=t, event_idx=k, id_base=id_base)
            r["sia_vaccinated"][t], r["sia_protected"][t] = a, p_
            r["new_exposed"][t] += p_
            r["new_exposed_by_strain"][t, :, strain] += p_
            r["sia_new_exposed_by_strain"][t, :, strain] += p_
        _, _, _, bfx, efx = orc.tx_step_prep(n_nodes, count, ns, c["strain"], srs, c["disease_state"], c["node_id"], c["daily_infectivity"],
                                             c["acq_risk_multiplier"], mode="fx")
        tallies[t] = (bfx, efx, orc.tx_step_prep.last_hist.copy())
        if node_math is not None:
            season = node_math["season"]
            tau, cdf, _, _ = orc.tx_node_math_device(bfx, efx, tallies[t][2], node_math["network"], season(t) if callable(season) else season,
                                                     node_math["r0_scalars"], r["pop"][t], node_math["zero_inflation"],
                                                     node_math["dispersion"], seed, t)
        else:
            tau, cdf = tau_of_tick[t], cdf_of_tick[t]
        new = orc.tx_infect_bernoulli(n_nodes, count, ns, c["node_id"], c["strain"], c["disease_state"], c["acq_risk_multiplier"],
                                      np.ascontiguousarray(tau, np.float32), np.ascontiguousarray(cdf, np.float64),
                                      seed=seed, tick=t, id_base=id_base)
        r["new_exposed"][t] += new.sum(axis=1, dtype=np.int32)
        r["new_exposed_by_strain"][t] += new
        census(t)
    if graveyard:  # back to the reference's order
        for name in c:
            out = np.empty_like(c[name])
            out[orig] = c[name]
            c[name][:] = out
        count += graveyard
    return r, count, tallies
