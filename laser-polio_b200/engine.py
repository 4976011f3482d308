"""Fused-tick engine: what ``SEIR_ABM.run()`` drives when the component list is the stock one.

The reference's loop body (model.py:252-263) calls every component's ``step()`` and then ``log(t)``; here the same
work for a whole day is ONE streaming pass over the agent table (``lpk_tick_pass``) plus node-level kernels
(``lpk_vd_births`` on vital-dynamics ticks, ``lpk_tick_node``), software-pipelined by one tick: the pass for tick t
first applies tick t-1's exposure trial and takes tick t-1's census, then runs tick t's deaths / disease-state / RI
stages and tick t's infectivity tally.  Results are identical to calling the components one by one (every draw is
keyed on (seed, agent, tick, stage)); ``tests/test_gpu_fused.py`` asserts that bit for bit.

The host never waits for the device inside a fused day: births are created in HBM and the live-slot count stays
there.  Days that need something the pass does not fuse -- an SIA campaign or a ``seed_schedule`` injection -- are
run through the components after draining the pending exposure + census, so any schedule is legal.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lpk
from . import kernels as K
from ._lpk import F_DEATHS, F_RI, F_SIA, Day, People, Rows, Run, check, dp, stream_handle


def eligible(sim) -> bool:
    """Fused path preconditions: stock components incl. disease state + transmission, no per-tick host decisions."""
    from . import abm

    names = [type(i).__name__ for i in sim.instances]
    if not all(type(i) is getattr(abm, n, None) for i, n in zip(sim.instances, names)):
        return False  # subclassed / foreign components: call their step() as written
    if "DiseaseState_ABM" not in names or "Transmission_ABM" not in names:
        return False
    return getattr(sim, "fused", True) and sim.verbose < 3


class FusedEngine:
    def __init__(self, sim):
        self.sim = sim
        self.dev = dev = sim.dev
        self.by_name = {type(i).__name__: i for i in sim.instances}
        n, ns, d = dev.n_nodes, dev.n_strains, dev.device
        c = dev.cols
        self.pending = False
        cap = sim.people.capacity
        n_tiles = (cap + _lpk.TILE_AGENTS - 1) // _lpk.TILE_AGENTS
        self.tile_node = dev.tile_node = torch.empty(n_tiles, dtype=torch.int32, device=d)
        self.rebuild_tiles(0)
        dev.set_count(sim.people.count)
        i32 = lambda *s: torch.zeros(s, dtype=torch.int32, device=d)  # noqa: E731
        i64 = lambda *s: torch.zeros(s, dtype=torch.int64, device=d)  # noqa: E731
        # tallies carried from tick to tick and corrected by the pass where an agent changes class (include/lpk.h,
        # lpk_tick_args): infectivity per node and strain, susceptibles / their risk sum / risk histogram per node,
        # exposed / infectious per node and strain, recovered per node
        self.beta, self.beta_sum = i64(n, ns), i64(n, ns)  # beta_sum: the all-reduced copy of a sharded run
        self.expo, self.sus, self.hist = i64(n), i64(n), i32(n, _lpk.RISK_BINS)
        self.E_cur, self.I_cur, self.R_cur = i32(n, ns), i32(n, ns), i32(n)
        # exposures of the pending tick found by a pass; census snapshots (lpk_node_args)
        self.tx_hits, self.tx_hits_s, self.S_snap, self.R_snap = i32(n), i32(n, ns), i32(n), i32(n)
        self.E_snap, self.I_snap = i32(n, ns), i32(n, ns)
        self._census_scratch = [i32(n) for _ in range(6)]
        self.deaths, self.dead_pp, self.dead_par = i32(n), i32(n), i32(n)
        self.cur_potp, self.cur_p = i32(n), i32(n)
        self.q = torch.zeros(n, dtype=torch.float32, device=d)
        self.cdf = torch.zeros((n, ns), dtype=torch.float64, device=d)
        self.prob = torch.zeros((n, ns), dtype=torch.float64, device=d)
        self.expected = torch.zeros(n, dtype=torch.float64, device=d)
        self.rowsum = torch.zeros(2 * n, dtype=torch.float64, device=d)
        self._rowsum_of = None  # the network tensor whose row sums self.rowsum[:n] holds
        # scheduling hint for the pass: the node-contiguous initial population ends here, appended cohorts follow
        init_total = int(np.sum(np.asarray(sim.pars.init_pop))) if "init_pop" in sim.pars else 0
        self.uniform_agents = init_total if 0 < init_total <= dev.count0 else int(dev.count0)
        self.dummy_row = i32(n * max(ns, 1))  # sink for rows of components that are absent
        # early stop (pars.stop_if_no_cases): "somebody is still exposed or infectious after tick t", one flag per tick,
        # mirrored into pinned host memory with an event so that the host never waits for more than one tick
        self.stop_rule = bool(sim.pars["stop_if_no_cases"])
        # capacity overflow at birth (lpk_births_args.status) mirrored into pinned memory behind an event: the host looks at it
        # one call late instead of synchronising, and raises like LaserFrame.add does in the reference
        self.status_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self.status_evt = None
        if self.stop_rule:
            self.cases_dev = i32(sim.nt + 1)
            self.cases_host = torch.zeros(sim.nt + 1, dtype=torch.int32).pin_memory()
            self.cases_evt = {}
        if sim.t > 0:  # resuming mid-run: re-base the incremental paralysis census on the last logged row
            self.cur_potp.copy_(dev.res["potentially_paralyzed"][sim.t - 1])
            self.cur_p.copy_(dev.res["paralyzed"][sim.t - 1])
        # agenda bytes (csrc/lpk_hot.cuh): the one byte per agent the pass streams, the earliest death date of every
        # 256-slot pair, the pass's work counter; the exponent bias of the 6-bit risk code comes from the largest risk
        padded = int(_lpk.lib().lpk_hot_padded(cap))
        self.hot = torch.empty(padded, dtype=torch.uint8, device=d)
        self.rec = torch.empty(cap, dtype=torch.int64, device=d)  # per-agent event records (lpk_people.rec)
        self.pair_min_dod = torch.empty(padded // 256, dtype=torch.int32, device=d) if "date_of_death" in c else None
        self.pair_ri_max = torch.empty(padded // 256, dtype=torch.int32, device=d) if "ri_timer" in c else None
        self.ri_k = torch.zeros(padded, dtype=torch.uint8, device=d) if "ri_timer" in c else None
        self.ri_lazy_k = 0  # RI ticks whose ri_timer subtraction is still owed (lpk_tick_args.ri_lazy_k)
        ri = self.by_name.get("RI_ABM")
        self.ri_step = int(ri.step_size) if ri is not None else 0
        self.work_counter = torch.zeros(64, dtype=torch.int32, device=d)
        risk = c["acq_risk_multiplier"]
        rmax = float(torch.where(torch.isfinite(risk), risk, torch.zeros_like(risk)).max().item()) if risk.numel() else 1.0
        if not bool(torch.isfinite(risk).all().item()):
            raise ValueError("acq_risk_multiplier must be finite")
        self.hot_valid = False  # True while exposure / infection / paralysis timers of E / I agents are deadlines (lpk_hot.cuh)
        P = People()
        for name in ("disease_state", "strain", "exposure_timer", "infection_timer", "paralysis_timer", "potentially_paralyzed",
                     "paralyzed", "ipv_protected", "chronically_missed", "node_id", "ri_timer", "acq_risk_multiplier",
                     "daily_infectivity", "date_of_birth", "date_of_death"):
            setattr(P, name, dp(c.get(name)))
        P.tile_node = dp(self.tile_node)
        P.capacity = cap
        P.hot, P.pair_min_dod, P.pair_ri_max, P.ri_k = dp(self.hot), dp(self.pair_min_dod), dp(self.pair_ri_max), dp(self.ri_k)
        P.rec = dp(self.rec)
        P.risk_e0 = int(_lpk.lib().lpk_hot_risk_e0(C.c_float(rmax)))
        self.P = P
        self.use_graph = bool(getattr(sim, "cuda_graph", False))
        if int(getattr(sim.pars, "compact_every", 0) or 0) > 0:
            dev.warm_compaction()
        self.xchg = self._open_exchange()
        self._template()
        if sim.t > 0:  # resuming mid-run; a fresh run builds tallies and agenda after tick 0 (after_component_tick)
            self.rebase_tallies(sim.t)

    def _open_exchange(self):
        """Peer-memory tally exchange of a node-sharded run (include/lpk.h, lpk_xchg): every rank's receive buffer is
        mapped into every other rank through CUDA IPC.  Used when the shard's process group runs on NCCL (one process per
        GPU); other groups (gloo in the tests, several ranks on one GPU) keep torch.distributed's all-reduce."""
        import os

        sim = self.sim
        if sim.shard is None or sim.shard.world <= 1 or os.environ.get("LPK_XCHG", "1") == "0":
            return None
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_backend(sim.shard.group) != "nccl":
            return None
        lib = _lpk.lib()
        x, handle = C.c_void_p(), (C.c_char * _lpk.XCHG_HANDLE_BYTES)()
        check(lib.lpk_xchg_create(C.c_int32(sim.shard.rank), C.c_int32(sim.shard.world), C.c_int64(self.dev.n_nodes * self.dev.n_strains),
                                  C.byref(x), handle), "lpk_xchg_create")
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(self.dev.device)
        every = [torch.empty_like(mine) for _ in range(sim.shard.world)]
        dist.all_gather(every, mine, group=sim.shard.group)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in every)
        check(lib.lpk_xchg_connect(x, blob), "lpk_xchg_connect")
        dist.barrier(group=sim.shard.group)
        return x

    # ------------------------------------------------------------------ helpers
    def rebuild_tiles(self, first_agent: int):
        first_tile = first_agent // _lpk.TILE_AGENTS
        check(_lpk.lib().lpk_build_tile_nodes(_lpk.ptr(self.dev.cols["node_id"]), C.c_int64(first_tile),
                                              C.c_int64(self.sim.people.capacity), _lpk.ptr(self.tile_node), stream_handle()),
              "lpk_build_tile_nodes")

    def settle(self, t_next):
        """Deadline timers of the exposed / infectious agents -> the countdown values tick ``t_next`` would test: the
        table is canonical again (what the per-function kernels and the host read)."""
        if self.hot_valid:
            check(_lpk.lib().lpk_hot_settle(C.byref(self.P), C.c_int64(self.sim.people.capacity), C.c_int32(t_next),
                                            C.c_int32(self.ri_lazy_k), C.c_int32(self.ri_step), stream_handle()), "lpk_hot_settle")
            self.hot_valid = False
            self.ri_lazy_k = 0

    def rebase_tallies(self, t_next):
        """From-scratch tallies and agenda bytes of the table as it stands before tick ``t_next`` (engine start, and after
        any tick that ran through the components, whose kernels do not maintain the carried values)."""
        sim, dev, c = self.sim, self.dev, self.dev.cols
        assert not self.hot_valid
        K.tx_step_prep(dev.n_nodes, sim.people.count, dev.n_strains, c["strain"], list(sim.pars.strain_r0_scalars.values()),
                       c["disease_state"], c["node_id"], c["daily_infectivity"], c["acq_risk_multiplier"],
                       out=(self.beta, self.expo, self.sus, self.hist))
        S, E, I, R, POTP, Pz = self._census_scratch
        K.count_SEIRP(c["node_id"], c["disease_state"], c["strain"], c["potentially_paralyzed"], c["paralyzed"], dev.n_nodes,
                      dev.n_strains, sim.people.count, out=(S, E, I, self.R_cur, self.E_cur, self.I_cur, POTP, Pz))
        self.tx_hits.zero_()
        self.tx_hits_s.zero_()
        check(_lpk.lib().lpk_hot_build(C.byref(self.P), C.c_int64(sim.people.capacity), C.c_int32(t_next), C.c_int32(max(self.ri_step, 1)), _lpk.ptr(dev.status),
                                       stream_handle()), "lpk_hot_build")
        self.hot_valid = True

    def _row(self, name, t):
        r = self.dev.res.get(name)
        return self.dummy_row if r is None else r[t]

    def sia_events(self, t):
        sia = self.by_name.get("SIA_ABM")
        if sia is None or self.sim.pars.vx_prob_sia is None:
            return []
        return sia._by_tick.get(t) or []

    def needs_components(self, t) -> bool:
        """Days the pass does not fuse: seed_schedule injections, several campaign events on one day (the reference
        overwrites sia_vaccinated / sia_protected per event, model.py:2142-2145), a campaign with an empty age window."""
        events = self.sia_events(t)
        if len(events) > 1 or (events and not (int(events[0]["age_range"][0]) <= int(events[0]["age_range"][1]))):
            return True
        ds = self.by_name["DiseaseState_ABM"]
        return t in ds.seed_schedule

    def early_stop_rule(self, t):
        """DiseaseState_ABM.step's early-stop test for tick t (reference model.py:789-795): nobody exposed or infectious
        in tick t-1's census and no seed_schedule event left -> tick t is the last one.  After a fused tick t-1 the
        census of t-1 is still in flight, but it has E or I agents exactly when some node had any after tick t-1's stages
        (with nobody infectious there is no force of infection, so tick t-1's transmission adds nobody): the flag
        lpk_tick_node(t-1) left, copied to pinned memory behind an event.  The host therefore runs at most one tick ahead
        of the device instead of synchronising every tick."""
        sim = self.sim
        ds = self.by_name["DiseaseState_ABM"]
        if any(ts > t for ts in ds.seed_schedule):
            return
        evt = self.cases_evt.get(t - 1)
        if evt is not None:
            evt.synchronize()
            active = int(self.cases_host[t - 1]) > 0
        else:  # tick t-1 ran through the components (or was tick 0): its census rows are complete on the device
            ei = self.dev.res["E"][t - 1].sum() + self.dev.res["I"][t - 1].sum()
            if sim.shard is not None and sim.shard.world > 1:
                import torch.distributed as dist

                dist.all_reduce(ei, group=sim.shard.group)
            active = int(ei.item()) > 0
        if not active:
            sim.should_stop = True

    # ------------------------------------------------------------------ pipeline
    def drain(self):
        """Back to the canonical table: settle the deadline timers, then apply the pending exposure trial and census of
        the last fused tick with the per-function kernels."""
        self.settle(self.sim.t)
        if not self.pending:
            return
        sim, dev = self.sim, self.dev
        t = sim.t - 1  # the tick the pending work belongs to
        c, r = dev.cols, dev.res
        n, ns = dev.n_nodes, dev.n_strains
        count = dev.sync_count(check=False)  # births of fused vital-dynamics ticks happened on the device (an overflow is reported by
        # the next run_days call or by download(), not from inside the drain)
        K.tx_infect(n, count, ns, c["node_id"], c["strain"], c["disease_state"], c["acq_risk_multiplier"], self.q, self.cdf,
                    rng=K.make_rng(sim.pars.seed, t, id_base=sim.id_base), out=dev.n_new)
        r["new_exposed"][t] += dev.n_new.sum(dim=1, dtype=torch.int32)
        r["new_exposed_by_strain"][t] += dev.n_new
        self.by_name["Transmission_ABM"].log(t)
        self.cur_potp.copy_(r["potentially_paralyzed"][t])
        self.cur_p.copy_(r["paralyzed"][t])
        self.pending = False

    def after_component_tick(self, t):
        """An unfused tick left nothing pending; re-base the incremental paralysis census on its full census."""
        r = self.dev.res
        self.cur_potp.copy_(r["potentially_paralyzed"][t])
        self.cur_p.copy_(r["paralyzed"][t])
        self.dev.set_count(self.sim.people.count)
        self.rebase_tallies(t + 1)

    def verify(self) -> dict:
        """Self-check used by bench.py and the tests: settle the table and compare every carried tally with the
        per-function kernels run from scratch on it (tx_step_prep, count_SEIRP), plus the head-count identity.  The
        table is left canonical with the last tick's exposure still pending (to_host() / the next tick finish it)."""
        sim, dev, c = self.sim, self.dev, self.dev.cols
        self.settle(sim.t)
        count = dev.sync_count()
        n, ns = dev.n_nodes, dev.n_strains
        beta, expo, sus, hist = K.tx_step_prep(n, count, ns, c["strain"], list(sim.pars.strain_r0_scalars.values()), c["disease_state"],
                                               c["node_id"], c["daily_infectivity"], c["acq_risk_multiplier"])
        S, E, I, R, Ebs, Ibs, PP, Pz = K.count_SEIRP(c["node_id"], c["disease_state"], c["strain"], c["potentially_paralyzed"],  # noqa: E741
                                                     c["paralyzed"], n, ns, count)
        state = c["disease_state"][:count]
        dead = int((state < 0).sum().item())
        out = {
            "beta_fx": bool(torch.equal(beta, self.beta)), "exposure_fx": bool(torch.equal(expo, self.expo)),
            "sus": bool(torch.equal(sus, self.sus)), "risk_hist": bool(torch.equal(hist, self.hist)),
            "E_by_strain": bool(torch.equal(Ebs, self.E_cur)), "I_by_strain": bool(torch.equal(Ibs, self.I_cur)),
            "R": bool(torch.equal(R, self.R_cur)), "S_equals_sus": bool(torch.equal(S.to(torch.int64), self.sus)),
            "head_count": int(S.sum().item() + E.sum().item() + I.sum().item() + R.sum().item()) + dead == count,
        }
        out["ok"] = all(out.values())
        return out

    # ------------------------------------------------------------------ the run template (include/lpk.h, lpk_run)
    def _template(self):
        """Everything a fused day needs that does not change from tick to tick, laid out once as liblpk's ``lpk_run``:
        lpk_run_days derives each day's births / pass / node arguments from it, so a day costs the host its launches only."""
        sim, dev, pars = self.sim, self.dev, self.sim.pars
        n, ns = dev.n_nodes, dev.n_strains
        R = Run()
        R.people = self.P
        A = R.tick
        A.n_nodes, A.n_strains = n, ns
        A.seed, A.id_base = int(pars.seed) & 0xFFFFFFFFFFFFFFFF, sim.id_base
        A.counts = dp(dev.counts)
        A.q_prev, A.cdf_prev = dp(self.q), dp(self.cdf)
        A.tx_hits, A.tx_hits_by_strain, A.R_cur = dp(self.tx_hits), dp(self.tx_hits_s), dp(self.R_cur)
        A.E_cur, A.I_cur = dp(self.E_cur), dp(self.I_cur)
        A.p_paralysis = float(np.float32(pars.p_paralysis))
        A.deaths, A.dead_pp, A.dead_par = dp(self.deaths), dp(self.dead_pp), dp(self.dead_par)
        A.ri_step = self.ri_step
        ri = self.by_name.get("RI_ABM")
        self.ri_active = ri is not None and pars["vx_prob_ri"] is not None
        if self.ri_active:
            p_ri, p_ipv = ri._probs(dev)
            A.ri_strain = 2 if "nOPV" in getattr(pars, "ri_vaccine_type", "tOPV") else 1
            A.vx_prob_ri, A.vx_prob_ipv = dp(p_ri), dp(p_ipv)
        for k, v in enumerate(list(pars.strain_r0_scalars.values())[:ns]):
            A.strain_r0_scalars[k] = float(v)
        A.beta_fx, A.exposure_fx, A.sus, A.risk_hist = dp(self.beta), dp(self.expo), dp(self.sus), dp(self.hist)
        A.uniform_agents = self.uniform_agents
        N = R.node
        N.n_nodes, N.n_strains, N.seed = n, ns, A.seed
        if sim.shard is not None:  # the node kernels touch this rank's nodes only: 1 / world of the network per tick
            N.node_lo, N.node_hi = int(sim.shard.node_lo), int(sim.shard.node_hi)
        N.beta_fx, N.exposure_fx, N.risk_hist = dp(self.beta), dp(self.expo), dp(self.hist)
        N.zero_inflation, N.dispersion = float(pars.node_seeding_zero_inflation), float(pars.node_seeding_dispersion)
        N.q, N.strain_cdf, N.prob, N.expected, N.rowsum_ws = dp(self.q), dp(self.cdf), dp(self.prob), dp(self.expected), dp(self.rowsum)
        N.deaths, N.dead_pp, N.dead_par = dp(self.deaths), dp(self.dead_pp), dp(self.dead_par)
        N.cur_potp, N.cur_p = dp(self.cur_potp), dp(self.cur_p)
        N.E_cur, N.I_cur, N.E_snap, N.I_snap = dp(self.E_cur), dp(self.I_cur), dp(self.E_snap), dp(self.I_snap)
        N.tx_hits_by_strain = dp(self.tx_hits_s)
        N.sus, N.R_cur, N.tx_hits, N.S_snap, N.R_snap = dp(self.sus), dp(self.R_cur), dp(self.tx_hits), dp(self.S_snap), dp(self.R_snap)
        N.counts = dp(dev.counts)
        if n > 1024:  # the network transfer is summed per chunk of 1024 source rows: scratch owned by this table
            self.matvec_ws = torch.zeros(((n + 1023) // 1024) * _lpk.MAX_STRAINS * n, dtype=torch.float64, device=dev.device)
            N.matvec_ws = dp(self.matvec_ws)
        vd = self.by_name.get("VitalDynamics_ABM")
        if vd is not None and pars.cbr is not None:
            R.births = vd.births_args(dev, 1, self.tile_node, tallies=(self.sus, self.expo, self.hist),
                                      hot=(self.hot, self.pair_min_dod, int(self.P.risk_e0), self.pair_ri_max, 0, self.ri_step, self.ri_k, self.rec))
        elif vd is not None:  # the component is there but cannot create anybody: its rows are still maintained; a
            R.births.capacity = -1  # vital-dynamics tick raises (ValueError, like the component's step)
        for name in Rows.NAMES:
            setattr(R.rows, name, dp(dev.res.get(name)))
        R.rows.sink = dp(self.dummy_row)
        R.zero_pop = dp(dev.zero_pop)
        R.any_cases = dp(self.cases_dev) if self.stop_rule else None
        R.work_counters = dp(self.work_counter)
        R.xchg = self.xchg
        R.graph = 1 if (self.use_graph and not self.stop_rule) else 0
        old = getattr(self, "R", None)
        if old is not None:  # rebuilt after a compaction: the exchange's sequence number and the row sums carry over
            R.seq, R.rowsums_valid = old.seq, old.rowsums_valid
        self.R = R

    def _days(self, t0, n_days):
        """``lpk_day`` array for ticks t0 .. t0 + n_days - 1 (flags, seasonality, the day's single campaign event) and the
        per-span refresh of what the reference re-reads from host attributes (the network, r0_scalars)."""
        from . import utils

        sim, dev, pars, R = self.sim, self.dev, self.sim.pars, self.R
        tx = self.by_name["Transmission_ABM"]
        vd, ri = self.by_name.get("VitalDynamics_ABM"), self.by_name.get("RI_ABM")
        net = dev.network_tensor(tx.network)
        if self._rowsum_of is not net:  # a new matrix: its row sums are recomputed by the first node kernel that sees it
            R.rowsums_valid = 0
            self._rowsum_of = net
        R.node.network, R.node.r0_scalars = dp(net), dp(tx._r0_scalars_dev(dev))
        days = (Day * n_days)()
        n_vd = 0
        t_saved = sim.t
        for k in range(n_days):
            t, d = t0 + k, days[k]
            flags = 0
            if vd is not None and t % vd.step_size == 0:
                if pars.cbr is None:
                    raise ValueError("VitalDynamics_ABM needs pars.cbr")
                flags |= F_DEATHS
                n_vd += 1
            if self.ri_active and t % ri.step_size == 0:
                flags |= F_RI
            events = self.sia_events(t)
            if events:  # exactly one (needs_components): the campaign runs inside the pass, after RI, like SIA_ABM.step
                flags |= F_SIA
                targeted, vx_prob, vx_eff, lo, hi, vstrain = self.by_name["SIA_ABM"].event_args(dev, events[0])
                R.tick.vx_prob_sia = dp(vx_prob)
                d.sia_targeted, d.sia_vx_eff = dp(targeted), float(vx_eff)
                d.sia_min_age, d.sia_max_age, d.sia_strain = int(lo), int(hi), int(vstrain)
            sim.t = t
            d.tick, d.flags, d.beta_seasonality = t, flags, float(utils.get_seasonality(sim))
        sim.t = t_saved
        return days, n_vd

    def run_days(self, t0, n_days):
        """Ticks t0 .. t0 + n_days - 1 as fused days, launched from C back to back (lpk_run_days); none of them may need the
        components (needs_components).  Does not advance sim.t."""
        sim, R = self.sim, self.R
        if self.ri_lazy_k >= self.RI_DEBT_MAX:  # pay the lazy RI countdown's debt before the eligibility byte runs out of range
            self.drain()
            self.rebase_tallies(t0)
        if self.status_evt is not None:  # the previous call's births: did a cohort not fit?
            self.status_evt.synchronize()
            self.status_evt = None
            self.dev.check_status(self.status_host.tolist())
        if not self.hot_valid:  # the table was settled behind the engine's back (verify()): finish and re-base
            self.drain()
            self.rebase_tallies(t0)
        days, n_vd = self._days(t0, n_days)
        R.pending, R.ri_lazy_k = int(self.pending), int(self.ri_lazy_k)
        launches = C.c_int64(0)
        lib = _lpk.lib()
        st = K.STATS
        if sim.shard is not None and sim.shard.world > 1 and self.xchg is None:
            # no peer-memory exchange (gloo / CPU-side process groups): the tally is summed by torch.distributed between
            # the two halves of every day
            from . import sharding

            for k in range(n_days):
                st.record("tick_pass", lambda k=k: check(lib.lpk_run_day_pass(C.byref(R), C.byref(days[k]), C.byref(launches), stream_handle()),
                                                        "lpk_run_day_pass"), 0)
                self.beta_sum.copy_(self.beta)  # the carried tally stays local; the node math sees the sum over ranks
                sharding.allreduce_tally(self.beta_sum, sim.shard)
                st.record("tick_node", lambda k=k: check(lib.lpk_run_day_node(C.byref(R), C.byref(days[k]), _lpk.ptr(self.beta_sum),
                                                                              C.byref(launches), stream_handle()), "lpk_run_day_node"), 0)
        else:
            ms = (C.c_float * (3 * n_days))() if st.timing else None
            check(lib.lpk_run_days(C.byref(R), days, C.c_int32(n_days), ms, C.byref(launches), stream_handle()), "lpk_run_days")
            for name, cnt in (("tick_pass", n_days), ("tick_node", n_days), ("vd_births", n_vd)):
                if cnt:
                    st.calls[name] = st.calls.get(name, 0) + cnt
            if ms is not None:
                for k in range(n_days):
                    if days[k].flags & F_DEATHS:
                        st.times.setdefault("vd_births", []).append(ms[3 * k])
                    st.times.setdefault("tick_pass", []).append(ms[3 * k + 1])
                    st.times.setdefault("tick_node", []).append(ms[3 * k + 2])
        st.launches += int(launches.value)
        self.pending, self.ri_lazy_k = bool(R.pending), int(R.ri_lazy_k)
        if n_vd:
            self.status_host.copy_(self.dev.status, non_blocking=True)
            self.status_evt = torch.cuda.Event()
            self.status_evt.record()
        if self.stop_rule:
            t = t0 + n_days - 1
            flag = self.cases_dev[t:t + 1]
            if sim.shard is not None and sim.shard.world > 1:
                import torch.distributed as dist

                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=sim.shard.group)
            self.cases_host[t:t + 1].copy_(flag, non_blocking=True)
            evt = torch.cuda.Event()
            evt.record()
            self.cases_evt = {t: evt}

    MAX_SPAN = 256   # days per lpk_run_days call: bounds what one call adds to the lazy RI debt (and the size of a captured graph)
    RI_DEBT_MAX = 200  # RI ticks owed (ri_lazy_k) at which the table is settled and re-based: lpk_people.ri_k is one byte (< 254)

    def fused_span(self, t0, t_end) -> int:
        """How many consecutive ticks from t0 (below t_end) can run as one lpk_run_days call."""
        if self.stop_rule:  # the host decides tick by tick (one tick behind the device)
            return 0 if self.needs_components(t0) else 1
        n = 0
        while (t0 + n < t_end and n < self.MAX_SPAN and not self.needs_components(t0 + n)
               and not (n > 0 and self.compaction_due(t0 + n))):
            n += 1
        return n

    # ------------------------------------------------------------------ compaction (pars.compact_every)
    def compaction_due(self, t) -> bool:
        k = int(getattr(self.sim.pars, "compact_every", 0) or 0)
        return k > 0 and t > 1 and t % k == 0

    def compact(self, t):
        """Before tick t: finish what is pipelined, compact the canonical table (device.DeviceState.compact), re-derive tiles,
        tallies and agenda from it.  Every live agent is node-contiguous again (uniform tiles) and the dead leave the sweep."""
        self.drain()
        live, dead = self.dev.compact()
        self.rebuild_tiles(0)
        self.uniform_agents = live
        self._template()
        self.rebase_tallies(t)
        self.sim._compactions = getattr(self.sim, "_compactions", 0) + 1
        return live, dead

    def close(self):
        """Release the peer-memory exchange (collective: every rank of the shard group calls it)."""
        _lpk.lib().lpk_run_release()
        if self.xchg is not None:
            import torch.distributed as dist

            lib = _lpk.lib()
            lib.lpk_xchg_disconnect(self.xchg)
            dist.barrier(group=self.sim.shard.group)
            lib.lpk_xchg_destroy(self.xchg)
            self.xchg = None
            self.R.xchg = None

    def tick(self, t):
        """Tick t >= 1 (tick 0 only logs, through the components)."""
        sim = self.sim
        if self.needs_components(t):
            self.drain()
            old = self.dev.sync_count()
            for component in sim.instances:
                with sim.perf_stats.start(component.__class__.__name__ + ".step()"):
                    component.step()
            sim.log_results(t)
            self.after_component_tick(t)
        else:
            if self.stop_rule:
                self.early_stop_rule(t)
            with sim.perf_stats.start("FusedTick.step()"):
                self.run_days(t, 1)
        sim.t += 1

    def advance(self, t_end):
        """Ticks sim.t .. t_end - 1 (or up to an early stop): fused days in spans launched from C, the days in between
        through the components.  What SEIR_ABM.run() and run_ticks() drive."""
        sim = self.sim
        while sim.t < t_end and not (sim.t > 1 and sim.should_stop):
            if self.compaction_due(sim.t):
                self.compact(sim.t)
            n = self.fused_span(sim.t, t_end)
            if n <= 1 or self.stop_rule:
                self.tick(sim.t)
                continue
            with sim.perf_stats.start("FusedTick.step()"):
                self.run_days(sim.t, n)
            sim.t += n
