"""Generate the golden vectors under tests/golden/ from the REFERENCE's own numba kernels.

Runs only in the build container (needs /root/reference; the kernels are
AST-loaded read-only by oracle/ref_loader.py, RNG call sites replaced by
injected per-agent uniform arrays).  The fixtures are small .npz files holding
the seeded inputs AND the reference outputs, so the tests that consume them
(tests/test_oracle_golden.py on CPU, tests/test_gpu_parity.py on the B200) never
need the reference checkout.

    python tests/golden/make_golden.py

Deterministic: the same command reproduces the same bytes (inputs come from
numpy Generators with fixed seeds; the reference kernels are integer state
machines once the uniforms are injected; the float32 tallies depend on the numba
thread count, which is pinned to 4 here).
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))

import numba as nb  # noqa: E402

import laser_polio_b200.synth as synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = Path(__file__).resolve().parent
N_AGENTS, CAPACITY, N_NODES, N_STRAINS = 20_000, 20_480, 7, 3
AGENT_COLS = list(synth.COLUMNS)


def population(seed, **kw):
    p = synth.synth_population(N_AGENTS, N_NODES, seed=seed, capacity=CAPACITY, f_exposed=0.08, f_infected=0.12,
                               f_recovered=0.1, f_dead=0.03, **kw)
    return p


def save(name, **arrays):
    np.savez_compressed(OUT / f"{name}.npz", **arrays)
    print(f"wrote {name}.npz  ({(OUT / f'{name}.npz').stat().st_size / 1024:.0f} KiB)")


def main():
    nb.set_num_threads(4)
    ref = ref_loader.load(inject_uniforms=True)
    rng = np.random.default_rng(20261017)

    # ---- D1: disease_state_step, 6 consecutive ticks, two paralysis probabilities ------------
    for tag, p_par in (("ds_p03", 0.3), ("ds_p2000", 1 / 2000)):
        p = population(11)
        inputs = {f"in_{k}": p[k].copy() for k in AGENT_COLS}
        ticks = 6
        u = rng.random((ticks, CAPACITY))
        new_pot = np.zeros((ticks, N_NODES), np.int32)
        new_par = np.zeros((ticks, N_NODES), np.int32)
        for t in range(ticks):
            ref["disease_state_step"](
                p["node_id"], N_NODES, p["disease_state"], p["strain"], p["count"], p["exposure_timer"],
                p["infection_timer"], p["potentially_paralyzed"], p["paralyzed"], p["ipv_protected"],
                p["paralysis_timer"], nb.float32(p_par), new_pot[t], new_par[t], u_inj=u[t],
            )
        save(tag, count=p["count"], n_nodes=N_NODES, p_paralysis=np.float32(p_par), u=u, new_potential=new_pot,
             new_paralyzed=new_par, **inputs, **{f"out_{k}": p[k] for k in AGENT_COLS})

    # ---- V1: get_deaths ---------------------------------------------------------------------
    p = population(12)
    p["date_of_death"][: p["count"]] = rng.integers(-5, 40, p["count"]).astype(np.int32)
    inputs = {f"in_{k}": p[k].copy() for k in AGENT_COLS}
    t = 21
    tl = np.zeros((nb.get_num_threads(), N_NODES), np.int32)
    dying = np.zeros(N_NODES, np.int32)
    ref["get_deaths"](np.int32(N_NODES), np.int32(p["count"]), p["disease_state"], p["node_id"], p["date_of_death"],
                      np.int32(t), tl, dying)
    save("deaths", count=p["count"], n_nodes=N_NODES, t=t, num_dying=dying, **inputs,
         out_disease_state=p["disease_state"])

    # ---- R1: fast_ri on the first RI tick (closed window) and a later one (half-open) ---------
    for tag, sim_t in (("ri_t14", 14), ("ri_t28", 28)):
        p = population(13)
        p["ri_timer"][: p["count"]] = rng.integers(-20, 40, p["count"]).astype(np.int16)
        inputs = {f"in_{k}": p[k].copy() for k in AGENT_COLS}
        pr = rng.uniform(0.2, 0.9, N_NODES)
        pi = rng.uniform(0.2, 0.9, N_NODES)
        u1, u2 = rng.random(CAPACITY), rng.random(CAPACITY)
        nt = nb.get_num_threads()
        c1, c2, c3 = (np.zeros((nt, N_NODES), np.int32) for _ in range(3))
        ref["fast_ri"](np.int32(14), p["node_id"], p["disease_state"], p["strain"], p["ipv_protected"], p["ri_timer"],
                       np.int32(sim_t), pr, pi, np.int32(p["count"]), c1, c2, c3, p["chronically_missed"], np.int8(1),
                       u1, u2)
        save(tag, count=p["count"], n_nodes=N_NODES, sim_t=sim_t, step_size=14, vx_prob_ri=pr, vx_prob_ipv=pi, u1=u1,
             u2=u2, ri_counts=c1.sum(0), ri_protected=c2.sum(0), ipv_counts=c3.sum(0), vaccine_strain=1, **inputs,
             **{f"out_{k}": p[k] for k in ("disease_state", "strain", "ipv_protected", "ri_timer")})

    # ---- S1: fast_sia ---------------------------------------------------------------------------
    p = population(14, max_age_days=8 * 365)
    inputs = {f"in_{k}": p[k].copy() for k in AGENT_COLS}
    vx = rng.uniform(0.3, 0.95, N_NODES).astype(np.float32)
    targeted = np.array([1, 0, 1, 1, 0, 1, 1], np.uint8)
    u = rng.random(CAPACITY)
    nt = nb.get_num_threads()
    lv, lp_ = np.zeros((nt, N_NODES), np.int32), np.zeros((nt, N_NODES), np.int32)
    sim_t, vx_eff, amin, amax = 100, 0.7 * 0.8, 0, 5 * 365
    ref["fast_sia"](p["node_id"], p["disease_state"], p["strain"], p["date_of_birth"], sim_t, vx, vx_eff, p["count"],
                    targeted, amin, amax, lv, lp_, p["chronically_missed"], np.int8(2), u)
    save("sia", count=p["count"], n_nodes=N_NODES, sim_t=sim_t, vx_prob=vx, vx_eff=vx_eff, nodes_to_vaccinate=targeted,
         min_age=amin, max_age=amax, u=u, vaccinated=lv.sum(0), protected=lp_.sum(0), vaccine_strain=2, **inputs,
         **{f"out_{k}": p[k] for k in ("disease_state", "strain")})

    # ---- T1 + C1: tallies and census on one population ------------------------------------------
    p = population(15)
    srs = np.array([1.0, 0.25, 0.125])
    n = p["count"]
    beta, expo, sus = ref["tx_step_prep"](N_NODES, n, N_STRAINS, p["strain"][:n], srs, p["disease_state"][:n],
                                          p["node_id"][:n], p["daily_infectivity"][:n], p["acq_risk_multiplier"][:n])
    S, E, I, R, Ebs, Ibs, PP, P = ref["count_SEIRP"](p["node_id"], p["disease_state"], p["strain"],  # noqa: E741
                                                      p["potentially_paralyzed"], p["paralyzed"], np.int32(N_NODES),
                                                      np.int32(N_STRAINS), np.int32(n))
    save("tally_census", count=n, n_nodes=N_NODES, n_strains=N_STRAINS, strain_r0_scalars=srs, beta=beta,
         exposure=expo, sus=sus, S=S, E=E, I=I, R=R, E_by_strain=Ebs, I_by_strain=Ibs, POTP=PP, P=P,
         **{f"in_{k}": p[k] for k in AGENT_COLS})

    # ---- T3: tx_infect_nb statistics (distributional pin: it consumes its own numba RNG) ----------
    # 400 repetitions on one small population: per-agent selection frequency and per-strain split.
    p = population(16)
    n = p["count"]
    state0 = p["disease_state"][:n].copy()
    sus_by_node = np.bincount(p["node_id"][:n][state0 == 0], minlength=N_NODES).astype(np.int64)
    prob = np.tile(np.array([[0.02, 0.01, 0.005]]), (N_NODES, 1))
    want = np.zeros((N_NODES, N_STRAINS), np.int32)
    want[:, 0] = [60, 0, 40, 25, 80, 10, 5]
    want[:, 1] = [20, 0, 10, 5, 10, 0, 5]
    reps = 400
    hits = np.zeros(n, np.int32)
    by_strain = np.zeros((N_NODES, N_STRAINS), np.int64)
    si = np.zeros(CAPACITY, np.int32)
    sp = np.zeros(CAPACITY, np.float32)
    for _ in range(reps):
        st = state0.copy()
        strain = p["strain"][:n].copy()
        nn = ref["tx_infect_nb"](N_NODES, n, N_STRAINS, sus_by_node, p["node_id"][:n], strain, st, si, sp,
                                 p["acq_risk_multiplier"][:n], prob, want)
        hits += (st == 1) & (state0 == 0)
        by_strain += nn
    save("tx_infect_stats", count=n, n_nodes=N_NODES, n_strains=N_STRAINS, reps=reps, prob=prob, want=want,
         sus_by_node=sus_by_node, hits=hits, new_by_strain=by_strain, **{f"in_{k}": p[k] for k in AGENT_COLS})


def ensemble():
    """Gate 3 fixture: 64-seed ensemble of per-node daily incidence from the REFERENCE's own transmission path --
    its numba tx_step_prep / tx_infect_nb / disease_state_step (native numba RNG, no injection) with the node-level
    block of Transmission_ABM.step (a method body, not extractable) restated by oracle.tx_foi + tx_draw_counts_ref
    on numpy's global stream, exactly the calls model.py:1332-1407 makes.  Two configurations: Poisson-like
    importation (defaults) and a calibrated-style zero-inflated, over-dispersed one."""
    from oracle import oracle as orc

    nb.set_num_threads(4)
    ref = ref_loader.load(inject_uniforms=False)

    @nb.njit(parallel=True)
    def seed_threads(s):
        for t in nb.prange(nb.get_num_threads()):
            np.random.seed(s + 7919 * nb.get_thread_id())

    n_agents, n_nodes, n_strains, ticks, seeds = 40_000, 5, 3, 30, 64
    p0 = synth.synth_population(n_agents, n_nodes, seed=77, f_exposed=0.0, f_infected=0.0, f_recovered=0.2, f_dead=0.0)
    n = p0["count"]
    # infections seeded in nodes 0 and 1 only: nodes 2-4 are reached through the network (importation branch)
    rs = np.random.default_rng(5)
    in01 = np.where((p0["node_id"][:n] <= 1) & (p0["disease_state"][:n] == 0))[0]
    p0["disease_state"][rs.choice(in01, 150, replace=False)] = 2
    srs = np.array([1.0, 0.25, 0.125])
    W = np.full((n_nodes, n_nodes), 0.01)
    np.fill_diagonal(W, 0.0)
    r0s = np.array([1.0, 0.8, 1.2, 1.0, 0.9])
    pop = np.bincount(p0["node_id"][:n], minlength=n_nodes).astype(np.int32)
    out = {}
    for tag, zi, disp in (("poisson", 0.0, 1000), ("zinb", 0.5, 1)):
        inc = np.zeros((seeds, ticks, n_nodes), np.int32)
        for s in range(seeds):
            np.random.seed(1000 + s)
            seed_threads(5000 + 31 * s)
            p = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in p0.items()}
            si, sp = np.zeros(len(p["disease_state"]), np.int32), np.zeros(len(p["disease_state"]), np.float32)
            for t in range(1, ticks + 1):
                a, b = np.zeros(n_nodes, np.int32), np.zeros(n_nodes, np.int32)
                ref["disease_state_step"](p["node_id"], n_nodes, p["disease_state"], p["strain"], n, p["exposure_timer"],
                                          p["infection_timer"], p["potentially_paralyzed"], p["paralyzed"], p["ipv_protected"],
                                          p["paralysis_timer"], nb.float32(1 / 2000), a, b)
                beta, expo, sus = ref["tx_step_prep"](n_nodes, n, n_strains, p["strain"][:n], srs, p["disease_state"][:n],
                                                      p["node_id"][:n], p["daily_infectivity"][:n], p["acq_risk_multiplier"][:n])
                beta_pre, prob = orc.tx_foi(beta, W, 1.0, r0s, pop)
                want, _ = orc.tx_draw_counts_ref(beta_pre, prob, expo, zi, disp, rs=np.random)
                new = ref["tx_infect_nb"](n_nodes, n, n_strains, sus, p["node_id"][:n], p["strain"][:n], p["disease_state"][:n],
                                          si, sp, p["acq_risk_multiplier"][:n], prob, want)
                inc[s, t - 1] = new.sum(axis=1)
        out[f"incidence_{tag}"] = inc
        print(tag, "mean cumulative incidence per node", inc.sum(1).mean(0))
    save("ensemble_ref", count=n, n_nodes=n_nodes, n_strains=n_strains, ticks=ticks, strain_r0_scalars=srs, network=W, r0_scalars=r0s,
         pop=pop, **out, **{f"in_{k}": p0[k] for k in AGENT_COLS})


from make_golden_defs import FULL, full_table  # noqa: E402  (shared with tests/test_ensemble_full.py)


def ensemble_full():
    """Gate 3 on a second shape with every stage of the tick on: 60 nodes, vital dynamics every 7 ticks (the reference's
    get_deaths; its host-side births block, model.py:1711-1734, restated on numpy's stream with core.KaplanMeierEstimator),
    disease_state_step, fast_ri every 14, one fast_sia campaign, transmission (tx_step_prep / node block / tx_infect_nb).
    Stored per seed and node: daily incidence (all new exposures of the tick: RI + SIA + transmission, i.e.
    results.new_exposed), new_potentially_paralyzed, new_paralyzed."""
    from laser_polio_b200 import core, utils
    from oracle import oracle as orc

    nb.set_num_threads(4)
    ref = ref_loader.load(inject_uniforms=False)

    @nb.njit(parallel=True)
    def seed_threads(s):
        for t in nb.prange(nb.get_num_threads()):
            np.random.seed(s + 7919 * nb.get_thread_id())

    c = FULL
    p0, node = full_table()
    n0, nn, ns, ticks, seeds = p0["count"], c["n_nodes"], c["n_strains"], c["ticks"], c["seeds"]
    srs = np.array([1.0, 0.25, 0.125])
    km = core.KaplanMeierEstimator(utils.create_cumulative_deaths(int(node["pop0"].sum()), max_age_years=100))
    birth_rate = np.full(nn, c["cbr"] / (365 * 1000))
    targeted = np.zeros(nn, np.uint8)
    targeted[: c["sia_nodes"]] = 1
    nt = nb.get_num_threads()
    inc = np.zeros((seeds, ticks, nn), np.int32)
    potp = np.zeros((seeds, ticks, nn), np.int32)
    par = np.zeros((seeds, ticks, nn), np.int32)
    for s in range(seeds):
        np.random.seed(3000 + s)
        seed_threads(7000 + 31 * s)
        p = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in p0.items()}
        cap = len(p["disease_state"])
        si, sp = np.zeros(cap, np.int32), np.zeros(cap, np.float32)
        n = n0
        pop = node["pop0"].copy()
        for t in range(1, ticks + 1):
            if t % c["vd_step"] == 0:
                tl, dying = np.zeros((nt, nn), np.int32), np.zeros(nn, np.int32)
                ref["get_deaths"](np.int32(nn), np.int32(n), p["disease_state"], p["node_id"], p["date_of_death"], np.int32(t), tl, dying)
                expected = c["vd_step"] * birth_rate * pop  # model.py:1712-1716
                whole = expected.astype(np.int32)
                births = whole + np.random.binomial(1, expected - whole)
                total = int(births.sum())
                if total > 0:
                    lo, hi = n, n + total
                    p["date_of_birth"][lo:hi] = t
                    p["date_of_death"][lo:hi] = t + km.predict_age_at_death(np.zeros(total, np.int32), max_year=100)
                    p["disease_state"][lo:hi] = 0
                    p["node_id"][lo:hi] = np.repeat(np.arange(nn, dtype=np.int16), births)
                    n = hi
                pop = pop + births - dying
            a, b = np.zeros(nn, np.int32), np.zeros(nn, np.int32)
            ref["disease_state_step"](p["node_id"], nn, p["disease_state"], p["strain"], n, p["exposure_timer"], p["infection_timer"],
                                      p["potentially_paralyzed"], p["paralyzed"], p["ipv_protected"], p["paralysis_timer"],
                                      nb.float32(c["p_paralysis"]), a, b)
            potp[s, t - 1], par[s, t - 1] = a, b
            if t % c["ri_step"] == 0:
                l1, l2, l3 = (np.zeros((nt, nn), np.int32) for _ in range(3))
                ref["fast_ri"](c["ri_step"], p["node_id"], p["disease_state"], p["strain"], p["ipv_protected"], p["ri_timer"], t,
                               node["vx_prob_ri"], node["vx_prob_ipv"], n, l1, l2, l3, p["chronically_missed"], np.int8(1))
                inc[s, t - 1] += l2.sum(axis=0)
            if t == c["sia_tick"]:
                l1, l2 = np.zeros((nt, nn), np.int32), np.zeros((nt, nn), np.int32)
                ref["fast_sia"](p["node_id"], p["disease_state"], p["strain"], p["date_of_birth"], t, node["vx_prob_sia"], c["sia_eff"], n,
                                targeted, c["sia_age"][0], c["sia_age"][1], l1, l2, p["chronically_missed"], np.int8(2))
                inc[s, t - 1] += l2.sum(axis=0)
            beta, expo, sus = ref["tx_step_prep"](nn, n, ns, p["strain"][:n], srs, p["disease_state"][:n], p["node_id"][:n],
                                                  p["daily_infectivity"][:n], p["acq_risk_multiplier"][:n])
            beta_pre, prob = orc.tx_foi(beta, node["network"], 1.0, node["r0_scalars"], pop)
            want, _ = orc.tx_draw_counts_ref(beta_pre, prob, expo, c["zi"], c["disp"], rs=np.random)
            new = ref["tx_infect_nb"](nn, n, ns, sus, p["node_id"][:n], p["strain"][:n], p["disease_state"][:n], si, sp,
                                      p["acq_risk_multiplier"][:n], prob, want)
            inc[s, t - 1] += new.sum(axis=1)
        if s % 16 == 0:
            print("seed", s, "agents", n, "incidence", inc[s].sum(), "potentially paralysed", potp[s].sum(), "paralysed", par[s].sum())
    print("mean cumulative incidence of the network", inc.sum((1, 2)).mean(), "paralysed", par.sum((1, 2)).mean())
    save("ensemble_full_ref", incidence=inc.astype(np.int16), new_potentially_paralyzed=potp.astype(np.int16), new_paralyzed=par.astype(np.int16))


if __name__ == "__main__":
    if "--ensemble" in sys.argv:
        ensemble()
        raise SystemExit(0)
    if "--ensemble-full" in sys.argv:
        ensemble_full()
        raise SystemExit(0)
    if not ref_loader.available():
        raise SystemExit("needs the reference checkout at /root/reference")
    main()
    ensemble()
    ensemble_full()
