"""The agenda representation of the fused pass (csrc/lpk_hot.cuh: one byte per agent, deadline timers, event handler)
against the oracle's canonical tick loop, on the CPU.

tests/hot_model.cu compiles the very functions the device kernel runs (LPK_HD) for the host and restates only the sweep's
control flow; here every tick of a synthetic population runs twice -- oracle stage by stage on canonical columns
(reference order: deaths, disease state, RI, SIA, tally, exposure, census) and the host model as ONE pipelined pass --
and after every tick the carried tallies, the per-node counts and (after settling the deadlines) every agent column must
be identical.  Covers the deadline algebra incl. int8 wrap-around, 6-bit agenda days with check-ins, the risk code's
superset property, deaths by pair_min_dod, RI / SIA takes and the born-today rule.
"""

import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hot_model.cu"
OUT = ROOT / "tests" / "_build" / "libhot_model.so"
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def hm():
    if not Path(NVCC).exists():
        pytest.skip("nvcc not available")
    deps = [SRC] + sorted((ROOT / "laser-polio_b200" / "csrc").glob("*.cuh")) + [ROOT / "include" / "lpk.h"]
    if not OUT.exists() or any(p.stat().st_mtime > OUT.stat().st_mtime for p in deps):
        OUT.parent.mkdir(exist_ok=True)
        subprocess.run([NVCC, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", str(OUT), str(SRC)], check=True)
    lib = C.CDLL(str(OUT))
    lib.hm_risk_e0.argtypes = [C.c_float]
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data


class Table:
    """Agent columns + carried tallies + the ctypes structs of include/lpk.h pointing at them (host memory)."""

    def __init__(self, hm, cols, n, cap, nodes, ns, seed, orc):
        from laser_polio_b200 import _lpk

        self.hm, self.c, self.n, self.cap, self.nodes, self.ns, self.seed = hm, cols, n, cap, nodes, ns, seed
        padded = (cap + 2047) // 2048 * 2048
        self.hot = np.zeros(padded, np.uint8)
        self.pair_min = np.zeros(padded // 256, np.int32)
        self.pair_ri = np.zeros(padded // 256, np.int32)
        self.ri_kcol = np.zeros(padded, np.uint8)
        self.rec = np.zeros(cap, np.uint64)
        self.ri_k = 0
        tiles = (cap + 511) // 512
        nid = cols["node_id"]
        self.tile_node = np.full(tiles, -1, np.int32)
        for k in range(tiles):
            blk = nid[k * 512:(k + 1) * 512]
            if len(blk) == 512 and blk[0] >= 0 and np.all(blk == blk[0]):
                self.tile_node[k] = blk[0]
        P = _lpk.People()
        for name in ("disease_state", "strain", "exposure_timer", "infection_timer", "paralysis_timer", "potentially_paralyzed",
                     "paralyzed", "ipv_protected", "chronically_missed", "node_id", "ri_timer", "acq_risk_multiplier",
                     "daily_infectivity", "date_of_birth", "date_of_death"):
            setattr(P, name, _ptr(cols[name]))
        P.tile_node, P.capacity, P.hot, P.pair_min_dod = _ptr(self.tile_node), cap, _ptr(self.hot), _ptr(self.pair_min)
        P.risk_e0 = hm.hm_risk_e0(C.c_float(float(cols["acq_risk_multiplier"].max())))
        P.pair_ri_max, P.ri_k, P.rec = _ptr(self.pair_ri), _ptr(self.ri_kcol), _ptr(self.rec)
        self.P = P
        i32 = lambda *s: np.zeros(s, np.int32)  # noqa: E731
        self.E_cur, self.I_cur, self.R_cur = i32(nodes, ns), i32(nodes, ns), i32(nodes)
        self.beta, self.expo, self.sus = np.zeros((nodes, ns), np.int64), np.zeros(nodes, np.int64), np.zeros(nodes, np.int64)
        self.hist = i32(nodes, 192)
        self.counts = np.array([n, n], np.int64)

    def rebase(self, orc, srs, t_next, ri_step=14):
        c, n = self.c, self.n
        _, _, sus, bfx, efx = orc.tx_step_prep(self.nodes, n, self.ns, c["strain"], srs, c["disease_state"], c["node_id"],
                                               c["daily_infectivity"], c["acq_risk_multiplier"], mode="fx")
        self.beta[:], self.expo[:], self.sus[:], self.hist[:] = bfx, efx, sus, orc.tx_step_prep.last_hist
        S, E, I, R, Ebs, Ibs, PP, Pz = orc.count_SEIRP(c["node_id"], c["disease_state"], c["strain"], c["potentially_paralyzed"],
                                                         c["paralyzed"], self.nodes, self.ns, n)
        self.E_cur[:], self.I_cur[:], self.R_cur[:] = Ebs, Ibs, R
        assert self.hm.hm_build(C.byref(self.P), C.c_int64(self.cap), C.c_int32(t_next), C.c_int32(ri_step)) == 0
        self.ri_k = 0


def run_case(hm, orc, n=60_000, nodes=7, ticks=40, seed=11, p_paralysis=0.3, big_first=True, weird_timers=False, tau_boost=1.0,
             sia_ticks=(9, 23), vd_step=7, ri_step=14, r0=3.0, f_exposed=0.04, f_infected=0.04, settle_every=13):
    from laser_polio_b200 import _lpk, synth

    ns = 3
    cap = n + 1500
    srs = np.array([1.0, 0.25, 0.125])
    can = synth.synth_population(n, nodes, seed=seed, capacity=cap, f_exposed=f_exposed, f_infected=f_infected, f_dead=0.02, r0=r0)
    can = {k: v for k, v in can.items() if isinstance(v, np.ndarray) and k != "node_sizes"}
    rs = np.random.RandomState(seed)
    if big_first:  # make node ids contiguous but with a first node large enough for uniform 512-agent tiles
        can["node_id"][:n] = np.sort(can["node_id"][:n])
    if weird_timers:  # negative / extreme timers: int8 wrap-around must be reproduced
        for name in ("exposure_timer", "infection_timer", "paralysis_timer"):
            k = rs.choice(n, n // 20, replace=False)
            can[name][k] = rs.randint(-128, 128, len(k)).astype(np.int8)
    can["date_of_death"][:n] = rs.randint(-5, 400, n)  # plenty of deaths on the vital-dynamics ticks
    can["ri_timer"][:n] = rs.randint(-20, 60, n)
    can["date_of_birth"][:n] = -rs.randint(1, 8 * 365, n)
    mod = {k: v.copy() for k, v in can.items()}
    T = Table(hm, mod, n, cap, nodes, ns, seed, orc)
    T.rebase(orc, srs, 1, ri_step)

    W = rs.random_sample((nodes, nodes)) * (0.1 / nodes)
    np.fill_diagonal(W, 0.0)
    r0s = rs.uniform(0.5, 1.5, nodes) * tau_boost
    pop = np.bincount(can["node_id"][:n], minlength=nodes).astype(np.int32)
    pr, pi = rs.uniform(0.3, 0.8, nodes), rs.uniform(0.3, 0.8, nodes)
    vx = rs.uniform(0.4, 0.9, nodes).astype(np.float32)
    targeted = (rs.random_sample(nodes) < 0.7).astype(np.uint8)
    targeted[0] = 1

    q_prev = np.zeros(nodes, np.float32)
    cdf_prev = np.zeros((nodes, ns), np.float64)
    pending = False
    i32 = lambda *s: np.zeros(s, np.int32)  # noqa: E731
    stats = np.zeros(4, np.int64)
    total_hits = 0
    for t in range(1, ticks + 1):
        is_vd, is_ri, is_sia = (t % vd_step == 0), (t % ri_step == 0), (t in sia_ticks)
        # ---------------- oracle, canonical, stage by stage
        dying_o = i32(nodes)
        if is_vd:
            orc.get_deaths(nodes, n, can["disease_state"], can["node_id"], can["date_of_death"], t, dying_o)
        pot_o, par_o = i32(nodes), i32(nodes)
        orc.disease_state_step(can["node_id"], nodes, can["disease_state"], can["strain"], n, can["exposure_timer"],
                               can["infection_timer"], can["potentially_paralyzed"], can["paralyzed"], can["ipv_protected"],
                               can["paralysis_timer"], p_paralysis, pot_o, par_o, seed=seed, tick=t)
        ri_o = [i32(nodes) for _ in range(3)]
        if is_ri:
            orc.fast_ri(ri_step, can["node_id"], can["disease_state"], can["strain"], can["ipv_protected"], can["ri_timer"], t, pr, pi, n,
                        *ri_o, can["chronically_missed"], 1, seed=seed, tick=t)
        sia_o = [i32(nodes), i32(nodes)]
        if is_sia:
            orc.fast_sia(can["node_id"], can["disease_state"], can["strain"], can["date_of_birth"], t, vx, 0.7, n, targeted, 0, 5 * 365,
                         *sia_o, can["chronically_missed"], 2, seed=seed, tick=t)
        _, _, sus_o, bfx_o, efx_o = orc.tx_step_prep(nodes, n, ns, can["strain"], srs, can["disease_state"], can["node_id"],
                                                     can["daily_infectivity"], can["acq_risk_multiplier"], mode="fx")
        hist_o = orc.tx_step_prep.last_hist.copy()
        census_o = orc.count_SEIRP(can["node_id"], can["disease_state"], can["strain"], can["potentially_paralyzed"], can["paralyzed"],
                                   nodes, ns, n)

        # ---------------- host model: one pass = pending exposure of t-1, then the stages of t
        A = _lpk.TickArgs()
        A.flags = _lpk.F_STAGES | (_lpk.F_PENDING if pending else 0) | (_lpk.F_DEATHS if is_vd else 0) | (_lpk.F_RI if is_ri else 0) | \
            (_lpk.F_SIA if is_sia else 0)
        A.tick, A.n_nodes, A.n_strains, A.seed, A.id_base = t, nodes, ns, seed, 0
        A.counts = _ptr(T.counts)
        A.q_prev, A.cdf_prev = _ptr(q_prev), _ptr(cdf_prev)
        ne_prev, nes_prev, tx_hits, tx_hits_s = i32(nodes), i32(nodes, ns), i32(nodes), i32(nodes, ns)
        A.new_exposed_prev, A.new_exposed_by_strain_prev, A.tx_hits, A.tx_hits_by_strain = map(_ptr, (ne_prev, nes_prev, tx_hits, tx_hits_s))
        A.p_paralysis = float(np.float32(p_paralysis))
        pot_m, par_m, d_m, dpp_m, dpar_m = i32(nodes), i32(nodes), i32(nodes), i32(nodes), i32(nodes)
        A.new_potential, A.new_paralyzed, A.deaths, A.dead_pp, A.dead_par = map(_ptr, (pot_m, par_m, d_m, dpp_m, dpar_m))
        ri_m = [i32(nodes) for _ in range(3)]
        ne_m, nes_m, rines_m = i32(nodes), i32(nodes, ns), i32(nodes, ns)
        A.ri_step, A.ri_strain, A.vx_prob_ri, A.vx_prob_ipv = ri_step, 1, _ptr(pr), _ptr(pi)
        A.ri_lazy_k = T.ri_k
        A.ri_vaccinated, A.ri_protected, A.ipv_vaccinated = map(_ptr, ri_m)
        A.new_exposed, A.new_exposed_by_strain, A.ri_new_exposed_by_strain = map(_ptr, (ne_m, nes_m, rines_m))
        sia_m = [i32(nodes), i32(nodes)]
        sianes_m = i32(nodes, ns)
        A.sia_targeted, A.vx_prob_sia, A.sia_vx_eff = _ptr(targeted), _ptr(vx), 0.7
        A.sia_min_age, A.sia_max_age, A.sia_strain, A.sia_event_idx = 0, 5 * 365, 2, 0
        A.sia_vaccinated, A.sia_protected, A.sia_new_exposed_by_strain = _ptr(sia_m[0]), _ptr(sia_m[1]), _ptr(sianes_m)
        for s in range(ns):
            A.strain_r0_scalars[s] = float(srs[s])
        A.beta_fx, A.E_cur, A.I_cur, A.exposure_fx, A.sus, A.risk_hist, A.R_cur = map(
            _ptr, (T.beta, T.E_cur, T.I_cur, T.expo, T.sus, T.hist, T.R_cur))
        rc = hm.hm_pass(C.byref(T.P), C.byref(A), C.c_int64(n), _ptr(stats))
        assert rc == 0, f"hm_pass rc={rc}"
        T.ri_k += 1 if is_ri else 0

        # the pass found tick t-1's exposures: they must be what the oracle's tx_infect made on the canonical table
        if pending:
            assert np.array_equal(nes_prev, new_o_prev), f"tick {t}: exposures of t-1"
            assert np.array_equal(ne_prev, new_o_prev.sum(axis=1))
            total_hits += int(ne_prev.sum())
        # stages of tick t
        assert np.array_equal(d_m, dying_o), f"tick {t}: deaths"
        assert np.array_equal(pot_m, pot_o) and np.array_equal(par_m, par_o), f"tick {t}: paralysis"
        for a, b in zip(ri_m, ri_o):
            assert np.array_equal(a, b), f"tick {t}: RI"
        for a, b in zip(sia_m, sia_o):
            assert np.array_equal(a, b), f"tick {t}: SIA"
        # carried tallies == from-scratch tallies of the canonical table after tick t's stages
        assert np.array_equal(T.sus, sus_o), f"tick {t}: sus"
        assert np.array_equal(T.expo, efx_o), f"tick {t}: exposure_fx"
        assert np.array_equal(T.hist, hist_o), f"tick {t}: risk_hist"
        assert np.array_equal(T.beta, bfx_o), f"tick {t}: beta_fx"
        assert np.array_equal(T.E_cur, census_o[4]) and np.array_equal(T.I_cur, census_o[5]) and np.array_equal(T.R_cur, census_o[3]), \
            f"tick {t}: E / I / R census"
        # the records' state byte follows the canonical disease_state tick by tick
        assert np.array_equal((T.rec[:n] & 0xFF).astype(np.uint8).view(np.int8), can["disease_state"][:n]), f"tick {t}: disease_state"

        # ---------------- node math (same for both) and the oracle's exposure of tick t
        q, cdf, _, _ = orc.tx_node_math_device(bfx_o, efx_o, hist_o, W, 1.05, r0s, pop, 0.2, 2.0, seed, t)
        if t % 11 == 5:
            q[0] = np.float32(3.0e38)  # "everybody" in node 0 once in a while
        new_o_prev = orc.tx_infect_bernoulli(nodes, n, ns, can["node_id"], can["strain"], can["disease_state"],
                                             can["acq_risk_multiplier"], q, cdf, seed=seed, tick=t)
        q_prev[:], cdf_prev[:] = q, cdf
        pending = True

        if t % settle_every == 0 or t == ticks:  # leave the fused representation: settle, apply the pending exposure canonically, compare all
            assert hm.hm_settle(C.byref(T.P), C.c_int64(cap), C.c_int32(t + 1), C.c_int32(T.ri_k), C.c_int32(ri_step)) == 0
            orc.tx_infect_bernoulli(nodes, n, ns, mod["node_id"], mod["strain"], mod["disease_state"], mod["acq_risk_multiplier"], q, cdf,
                                    seed=seed, tick=t)
            for name in can:
                assert np.array_equal(mod[name], can[name]), f"tick {t}: column {name}"
            T.rebase(orc, srs, t + 1, ri_step)
            pending = False
    return stats, total_hits


def test_agenda_pass_equals_canonical_loop(hm, oracle):
    stats, hits = run_case(hm, oracle)
    assert hits > 300 and stats[1] >= hits  # candidates are a superset of the hits
    assert stats[1] < 3 * hits + 2000  # ... and a tight one (the risk code costs < 25 % extra candidates + saturated nodes)


def test_agenda_pass_wraparound_timers_and_mixed_nodes(hm, oracle):
    # small nodes (every tile mixed -> general path), timers anywhere in int8
    stats, hits = run_case(hm, oracle, n=20_000, nodes=40, ticks=70, seed=5, big_first=False, weird_timers=True, sia_ticks=(4, 28, 29))
    assert hits > 50


def test_agenda_pass_long_infections_check_in(hm, oracle):
    # 140 ticks: infections longer than the 63-day look-ahead need check-in events; p_paralysis = 1 exercises the gate
    stats, hits = run_case(hm, oracle, n=30_000, nodes=3, ticks=140, seed=7, p_paralysis=1.0, r0=1.5, sia_ticks=(50,), settle_every=75)
    assert hits > 100


def test_agenda_pass_high_force_of_infection(hm, oracle):
    stats, hits = run_case(hm, oracle, n=30_000, nodes=4, ticks=25, seed=9, tau_boost=400.0, sia_ticks=())
    assert hits > 5000


def test_risk_code_is_an_upper_bound(hm):
    lib = C.CDLL(str(OUT))
    # exercised through the model: every risk's code bound is >= the risk, < 1.25x + one step above it
    rs = np.random.RandomState(0)
    risks = np.exp(rs.normal(-0.8, 1.27, 200_000)).astype(np.float32)
    e0 = lib.hm_risk_e0(C.c_float(float(risks.max())))
    b = risks.view(np.uint32)
    code = ((b >> 23).astype(np.int64) - 127 + e0) * 4 + ((b >> 21) & 3) + ((b & 0x1FFFFF) != 0)
    code = np.clip(code, 0, 63)
    ub = np.ldexp(1.0 + 0.25 * (code & 3), (code >> 2) - e0)
    assert code.max() <= 63 and np.all(ub >= risks)
    tight = ub[code > 0] / risks[code > 0]
    assert tight.max() <= 1.25 + 1e-6 and tight.mean() < 1.13
