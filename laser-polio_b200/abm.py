"""SEIR_ABM and its components: the reference's component API over the device-resident agent table.

Drop-in surface (reference ``src/laser_polio/model.py``):

    sim = lp.SEIR_ABM(pars)                     # model.py:101-175
    sim.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.SIA_ABM, lp.Transmission_ABM]
    sim.run()                                   # model.py:246-282
    sim.results.S, sim.people.disease_state ... # host numpy, same shapes / dtypes / indexing

Construction (population, timers, immunity, seeding, network) is one-off host numpy work, as in
the reference.  ``run()`` copies the agent columns to HBM once, every ``component.step()`` /
``component.log(t)`` then launches the sm_100a kernels of liblpk on the current CUDA stream, and the
columns and results come back in one bulk copy when the run ends.  There is no host compute
fallback for the per-tick path: without a CUDA device or without liblpk.so, ``run()`` raises.

Host-side on purpose (rare and tiny): ``seed_schedule`` injections (a handful of ticks
per run).  Births are drawn and appended on the device (include/lpk.h V2), so ``sim.people.count`` is refreshed
from the device at the synchronisation points (``to_host()``, and every vital-dynamics tick in component mode).
"""

from __future__ import annotations

import logging
import numbers
from collections import defaultdict
from copy import deepcopy
from datetime import datetime

import numpy as np

from . import core, netbuild, pars as _pars, popinit, utils
from .core import LaserFrame, PropertySet

logger = logging.getLogger("laser-polio-b200")

__all__ = ["SEIR_ABM", "DiseaseState_ABM", "Transmission_ABM", "VitalDynamics_ABM", "RI_ABM", "SIA_ABM",
           "populate_heterogeneous_values"]


# value an agent column holds in a slot nobody has been born into yet (reference model.py:144-162, 1550, 1603, 1891)
_UNBORN_DEFAULTS = {"disease_state": -1, "potentially_paralyzed": -1, "paralyzed": 0, "ipv_protected": 0, "strain": 0,
                    "chronically_missed": 0, "node_id": -1, "date_of_birth": -1, "date_of_death": 0, "ri_timer": -1}


def _say(colour: str, msg: str) -> None:
    code = {"cyan": 36, "red": 31, "green": 32, "yellow": 33}[colour]
    print(f"\033[{code}m{msg}\033[0m")


def _device_init(pars) -> bool:
    """``pars.device_init`` (an extension key, default False): draw the per-agent columns in HBM (popinit / lpk_init.cu)
    instead of with the host's numpy stream.  Same distributions, Philox streams keyed on (seed, agent, stage)."""
    return bool(pars["device_init"]) if "device_init" in pars else False


def _verbosity(pars) -> int:
    return pars["verbose"] if "verbose" in pars else 1


# =========================================================================== SEIR_ABM
class SEIR_ABM:
    """Agent-based SEIR polio model; disease_state codes -1 dead/unborn, 0 S, 1 E, 2 I, 3 R (reference model.py:51-341)."""

    def _common_init(self, pars, verbose):
        self.perf_stats = utils.TimingStats()
        self.pars = deepcopy(_pars.default_pars)
        if pars is not None:
            unknown = set(pars.to_dict()) - set(self.pars.to_dict())
            if unknown:  # reference model.py:64-69: warn and drop
                _say("red", f"Warning: ignoring unexpected parameters: {sorted(unknown)}")
            self.pars <<= {k: v for k, v in pars.items() if k in self.pars}
        pars = self.pars
        self.verbose = _verbosity(pars)
        if pars.seed is None:
            now = datetime.now()  # noqa: DTZ005
            pars.seed = now.microsecond ^ int(now.timestamp())
            if self.verbose >= 1:
                _say("green", f"No seed provided. Using random seed of {pars.seed}.")
        core.seed(pars.seed)
        self.t = 0
        self.nt = pars.dur + 1  # tick 0 records the initial conditions, then pars.dur steps
        self.datevec = utils.daterange(pars["start_date"], days=self.nt)
        self.should_stop = False
        self.dev = None  # DeviceState while the agent table is resident in HBM
        self._engine = None  # engine.FusedEngine while ticks are being fused
        self.fused = True  # set False to force component-by-component ticks
        self.id_base = 0  # global id of local agent 0 when this table is one node-shard of a larger population
        self.shard = None  # sharding.Shard when the population is split by node over several GPUs
        if self.verbose >= 1:
            _say("cyan", "Initializing simulation...")

    def __init__(self, pars: PropertySet = None, verbose=1):
        self._common_init(pars, verbose)
        pars = self.pars
        pars.init_pop = np.atleast_1d(pars.init_pop).astype(int)
        if pars.init_sus_by_age is not None:  # agents are the susceptibles only (reference model.py:115-121)
            init_sus = pars.init_sus_by_age.groupby("node_id")["n_susceptible"].sum().astype(int)
            pars.init_sus = np.atleast_1d(init_sus)
            total = int(pars.init_sus.sum())
        else:
            total = int(np.sum(pars.init_pop))
        # capacity = live agents + expected births with the reference's adaptive safety margin (model.py:124-137)
        expected_births = 0
        if pars.cbr is not None and len(pars.cbr) >= 1:
            cbr = pars.cbr[0] if len(pars.cbr) == 1 else np.mean(pars.cbr)
            expected_births = float(core.calc_capacity(pars.init_pop.sum(), pars.dur + 100, cbr)) - pars.init_pop.sum()
        fudge = 1 + 4 / np.sqrt(expected_births) if expected_births > 0 else 1
        capacity = int(fudge * (total + expected_births))
        self.people = LaserFrame(capacity=capacity, initial_count=total)
        people = self.people
        people.add_scalar_property("disease_state", dtype=np.int8, default=-1)
        people.disease_state[: people.count] = 0
        people.add_scalar_property("potentially_paralyzed", dtype=np.int8, default=-1)
        people.add_scalar_property("paralyzed", dtype=np.int8, default=0)
        people.add_scalar_property("ipv_protected", dtype=np.int8, default=0)
        self.results = LaserFrame(capacity=1)
        people.add_scalar_property("strain", dtype=np.int8, default=0)
        people.add_scalar_property("chronically_missed", dtype=np.uint8, default=0)
        n_missed = int(pars.missed_frac * people.count)
        if _device_init(pars):
            popinit.init_frame_device(people, pars, {"missed"})
        else:
            people.chronically_missed[np.random.choice(people.count, size=n_missed, replace=False)] = 1
        people.add_scalar_property("node_id", dtype=np.int16, default=-1)
        self.nodes = np.arange(len(pars.init_pop))
        by_node = pars.init_sus if pars.init_sus_by_age is not None else pars.init_pop
        if np.any(by_node <= 0) or np.any(np.isnan(by_node)):
            raise ValueError("pop_by_node must be positive & non-nan")
        people.node_id[: int(np.sum(by_node))] = np.repeat(np.arange(len(by_node)), by_node)  # node-contiguous
        self._components = []
        self.instances = []

    @classmethod
    def init_from_file(cls, people: LaserFrame, pars: PropertySet = None):
        """Build around a pre-made agent table (reference model.py:177-195)."""
        sim = cls.__new__(cls)
        sim._common_init(pars, verbose=2)
        sim.people = people
        sim.nodes = np.unique(people.node_id[: people.count])
        sim.results = LaserFrame(capacity=1)
        sim._components = []
        sim.instances = []
        return sim

    # ------------------------------------------------------------------ components
    @property
    def components(self) -> list:
        return self._components

    @components.setter
    def components(self, components: list) -> None:
        """Keep the classes named in ``default_run_order``, in that order, and instantiate them (model.py:209-244)."""
        by_name = {c.__name__: c for c in components}
        self._components = [by_name[n] for n in _pars.default_run_order if n in by_name]
        self.instances = []
        for c in self._components:
            with self.perf_stats.start(c.__name__ + ".__init__()"):
                self.instances.append(c(self))
        if self.verbose >= 2:
            print(f"Initialized components: {self.instances}")

    # ------------------------------------------------------------------ node sharding (one process per GPU)
    def shard_to(self, rank: int, world: int, group=None):
        """Keep only this rank's contiguous block of nodes: call on every rank after the components are
        set and before ``run()``, on identically constructed sims (same pars / seed).  Node ids stay global; per-node
        results are filled for owned nodes only.  Returns the :class:`sharding.Shard`."""
        from . import sharding

        if self.dev is not None:
            raise RuntimeError("shard_to() must be called before the population is moved to the device")
        people, n_nodes = self.people, len(self.nodes)
        count = people.count
        nid = people.node_id[:count]
        if np.any(np.diff(nid.astype(np.int32)) < 0):
            raise ValueError("shard_to() needs a node-contiguous agent table (true for a freshly constructed sim)")
        sizes = np.bincount(nid, minlength=n_nodes)
        blocks = sharding.plan_node_blocks(sizes, world)
        starts = np.concatenate([[0], np.cumsum(sizes)])
        spare = people.capacity - count
        caps = []
        for lo, hi in blocks:  # spare capacity (room for births) in proportion to the block's agents
            n_r = int(starts[hi] - starts[lo])
            caps.append(n_r + int(np.ceil(spare * n_r / max(count, 1))) + 16)
        lo, hi = blocks[rank]
        a0, a1 = int(starts[lo]), int(starts[hi])
        n_r, cap_r = a1 - a0, caps[rank]
        unborn0 = count + sum(c - int(starts[b[1]] - starts[b[0]]) for c, b in zip(caps[:rank], blocks[:rank]))
        local = LaserFrame(capacity=cap_r, initial_count=n_r)
        for name, col in people.columns().items():
            local.add_scalar_property(name, dtype=col.dtype, default=0)
            new = getattr(local, name)
            new[:n_r] = col[a0:a1]
            tail = col[unborn0 : unborn0 + (cap_r - n_r)]  # pre-drawn per-slot values of not-yet-born agents
            new[n_r : n_r + len(tail)] = tail
            if len(tail) < cap_r - n_r:  # more room than the original table had: state-like columns get their unborn default,
                # pre-drawn per-slot columns (timers, risk, infectivity: iid draws) are recycled from the live slots
                new[n_r + len(tail):] = (_UNBORN_DEFAULTS[name] if name in _UNBORN_DEFAULTS
                                         else np.resize(col[:count], cap_r - n_r - len(tail)))
        self.people = local
        for inst in self.instances:
            inst.people = local
        self.id_base = sharding.id_bases(sizes, blocks, capacity_per_block=caps)[rank]
        self.shard = sharding.Shard(rank=rank, world=world, node_lo=lo, node_hi=hi, group=group)
        return self.shard

    # ------------------------------------------------------------------ device residency
    def to_device(self, device=None):
        """H2D of every agent column and results array (idempotent while resident)."""
        if self.dev is None:
            from .device import DeviceState

            self.dev = DeviceState(self, device)
            self._engine = None
        return self.dev

    def to_host(self):
        """Finish any pipelined work, D2H the agent columns and results into the host arrays (in place), drop the device copy."""
        if self.dev is not None:
            if self._engine:  # None: never started, False: components only
                self._engine.drain()
                self._engine.close()
            self._engine = None
            dev, self.dev = self.dev, None
            try:
                dev.download()  # raises (after copying everything back) if a newborn cohort did not fit the frame
            finally:
                self.io_bytes = (dev.h2d_bytes, dev.d2h_bytes)

    # ------------------------------------------------------------------ tick loop
    def _component_tick(self, tick: int) -> None:
        """The reference's loop body verbatim (model.py:252-263): every step() in run order, then every log()."""
        if tick > 0:
            for component in self.instances:
                with self.perf_stats.start(component.__class__.__name__ + ".step()"):
                    component.step()
        self.log_results(tick)
        self.t += 1

    def step_tick(self, tick: int) -> None:
        """Advance one tick on the device.  With the stock component list the day runs as the fused pass
        (engine.FusedEngine: same results, one sweep over the agent table); otherwise component by component."""
        self.to_device()
        if getattr(self, "_engine", None) is None:
            from . import engine

            self._engine = engine.FusedEngine(self) if engine.eligible(self) else False
        if self._engine is False:
            self._component_tick(tick)
        elif tick == 0:
            self._component_tick(tick)
            self._engine.after_component_tick(0)
        else:
            if self._engine.compaction_due(tick):
                self._engine.compact(tick)
            self._engine.tick(tick)

    def run_ticks(self, n: int) -> None:
        """Advance up to ``n`` ticks on the device (fewer on an early stop), leaving the population resident.  With the
        stock component list consecutive fused days are launched from C in one call (engine.FusedEngine.advance)."""
        t_end = min(self.t + int(n), self.nt)
        while self.t < t_end:
            if getattr(self, "_engine", None) and self.t >= 1:
                self._engine.advance(t_end)
            else:
                self.step_tick(self.t)
            if self.t > 1 and self.should_stop:
                break

    def run(self):
        if self.verbose >= 1:
            _say("cyan", "Initialization complete. Running simulation...")
        self.to_device()
        try:
            self.run_ticks(self.nt - self.t)
            if self.should_stop and self.verbose >= 1:
                _say("yellow", f"[SEIR_ABM] Early stopping at t={self.t}: no E/I and no future seed_schedule events. "
                               "This stops all components (e.g., no births, deaths, or vaccination)")
        finally:
            self.to_host()
        if self.verbose >= 1:
            _say("cyan", "Simulation complete.")
        self.perf_stats.log(logger)

    def log_results(self, t):
        for component in self.instances:
            with self.perf_stats.start(component.__class__.__name__ + ".log()"):
                component.log(t)

    def plot(self, save=False, results_path=None):
        raise NotImplementedError("plotting is outside the per-tick hot path; use the reference's plots on sim.results")

    def rng(self, tick=None):
        from . import kernels

        return kernels.make_rng(self.pars.seed, self.t if tick is None else tick, id_base=self.id_base)


def _need_dev(sim):
    if sim.dev is None:
        sim.to_device()
    return sim.dev


# =========================================================================== DiseaseState_ABM
class DiseaseState_ABM:
    """E->I->R timers, paralysis gate, scheduled seeding, early stop (reference model.py:505-811)."""

    def _common_init(self, sim):
        self.sim, self.people, self.pars, self.nodes, self.results = sim, sim.people, sim.pars, sim.nodes, sim.results
        self.verbose = _verbosity(self.pars)
        self.seed_schedule = defaultdict(list)  # tick -> [(node_id, prevalence or count)]
        for entry in self.pars.seed_schedule or []:
            if "date" in entry and "dot_name" in entry:
                t = (utils.date(entry["date"]) - self.pars.start_date).days
                nid = next((n for n, info in self.pars.node_lookup.items() if info["dot_name"] == entry["dot_name"]), None)
                if nid is not None:
                    self.seed_schedule[t].append((nid, entry["prevalence"]))
            elif "timestep" in entry and "node_id" in entry:
                self.seed_schedule[entry["timestep"]].append((entry["node_id"], entry["prevalence"]))

    def _init_results(self):
        nt, nn, ns = self.sim.nt, len(self.nodes), len(self.pars.strain_ids)
        for name in ("S", "E", "I", "R", "potentially_paralyzed", "paralyzed", "new_potentially_paralyzed", "new_paralyzed", "pop"):
            self.results.add_array_property(name, shape=(nt, nn), dtype=np.int32)
        for name in ("E_by_strain", "I_by_strain"):
            self.results.add_array_property(name, shape=(nt, nn, ns), dtype=np.int32)
        self.results.pop[0] = self.pars.init_pop

    @classmethod
    def init_from_file(cls, sim):
        self = cls.__new__(cls)
        self._common_init(sim)
        self._init_results()
        people, pars = self.people, self.pars
        lo, cap = getattr(people, "_unborn_from", None), people.capacity
        if lo is not None and lo < cap:
            # a table loaded from a snapshot holds the live prefix only: draw the columns of the slots future newborns will
            # take, as the reference does at this point (model.py:506-524, including its -1 defaults for the three flags)
            if _device_init(pars):  # (the device kernels apply the constructor's clipped paralysis-timer rule here too)
                popinit.init_frame_device(people, pars, {"heterogeneity", "timers"}, start=lo)
            else:
                populate_heterogeneous_values(lo, cap, people.acq_risk_multiplier, people.daily_infectivity, pars)
                people.exposure_timer[lo:cap] = pars.dur_exp(cap - lo)
                people.infection_timer[lo:cap] = pars.dur_inf(cap - lo)
                people.paralysis_timer[lo:cap] = pars.t_to_paralysis(cap - lo)
            people.potentially_paralyzed[lo:cap] = -1
            people.paralyzed[lo:cap] = -1
            people.ipv_protected[lo:cap] = -1
            people.disease_state[lo:cap] = -1
            people.node_id[lo:cap] = -1
            if hasattr(people, "ri_timer"):
                people.ri_timer[lo:cap] = -1
            if hasattr(people, "date_of_birth"):
                people.date_of_birth[lo:cap] = -1
            people._unborn_from = None
        return self

    def __init__(self, sim):
        self._common_init(sim)
        self._init_results()
        people, pars = self.people, self.pars
        cap = people.capacity
        # timers for every slot, born or not (reference model.py:571-587): int8, truncation before the clip
        for name in ("exposure_timer", "infection_timer", "paralysis_timer"):
            people.add_scalar_property(name, dtype=np.int8, default=0)
        if _device_init(pars):
            popinit.init_frame_device(people, pars, {"timers"})
        else:
            people.exposure_timer[:] = pars.dur_exp(cap)
            people.infection_timer[:] = pars.dur_inf(cap)
            people.exposure_timer[:] = np.clip(people.exposure_timer, 0, 127)
            people.infection_timer[:] = np.clip(people.infection_timer, 0, 127)
            remaining = pars.t_to_paralysis(cap) - people.exposure_timer  # onset measured from exposure
            people.paralysis_timer[:] = np.clip(remaining, 0, np.minimum(people.infection_timer, 127)).astype(np.int8)
        self._init_immunity()
        self._seed_initial_infections()

    def _init_immunity(self):
        sim, people, pars = self.sim, self.people, self.pars
        if pars.init_sus_by_age is None:
            imm = pars.init_immun
            if isinstance(imm, float):
                fracs = np.full(len(pars.init_pop), imm, dtype=np.float32)
            elif isinstance(imm, list):
                fracs = np.asarray(imm, dtype=np.float32)
            elif isinstance(imm, np.ndarray):
                fracs = imm
                assert fracs.shape == pars.init_pop.shape, "init_immun must match init_pop shape"
            else:
                raise ValueError(f"Unsupported init_immun type: {type(imm)}")
            starts = np.concatenate([[0], np.cumsum(pars.init_pop)])
            for nid, (frac, npop) in enumerate(zip(fracs, pars.init_pop)):
                assert 0.0 <= frac <= 1.0, f"Invalid immunity fraction: {frac} for node {nid}"
                k = int(frac * npop)
                if k > 0:
                    members = np.arange(starts[nid], starts[nid + 1])  # initial agents are node-contiguous
                    people.disease_state[np.random.choice(members, size=k, replace=False)] = 3
        else:
            # susceptibles-only table: immunes are carried as counts in results.R (reference model.py:614-667)
            people.disease_state[:] = 0
            sim.results.R[:, :] += pars.init_pop - pars.init_sus
            people.ipv_protected[: people.count] = 0
            table = pars.init_sus_by_age
            ipv = table[table["n_ipv_protected"] > 0]
            for node in ipv["node_id"].unique():
                idx = np.where(people.node_id[: people.count] == node)[0]
                if len(idx) == 0:
                    continue
                age_yr = -people.date_of_birth[idx] / 365.0
                for _, row in ipv[ipv["node_id"] == node].iterrows():
                    pool = idx[(age_yr >= row["age_min_yr"]) & (age_yr < row["age_max_yr"])]
                    k = min(int(row["n_ipv_protected"]), len(pool))
                    if k > 0:
                        people.ipv_protected[np.random.choice(pool, size=k, replace=False)] = 1
            if hasattr(self.results, "deaths"):
                self.results.R[:, :] -= np.cumsum(self.results.deaths, axis=0)

    def _seed_initial_infections(self):
        people, pars = self.people, self.pars
        prev = pars.init_prev
        total = int(sum(pars.init_pop))
        if isinstance(prev, float):
            chosen = np.random.choice(total, size=int(total * prev), replace=False)
        elif isinstance(prev, int):
            chosen = np.random.choice(total, size=min(prev, total), replace=False)
        elif isinstance(prev, (list, np.ndarray)):
            if len(prev) != len(pars.init_pop):
                raise ValueError(f"Length mismatch: init_prev has {len(prev)} entries, expected {len(pars.init_pop)} nodes.")
            nid = people.node_id[: people.count]
            alive = people.disease_state[: people.count] >= 0
            picks = []
            for node, v in enumerate(prev):
                if not isinstance(v, numbers.Real):
                    raise ValueError(f"Unsupported value in init_prev list at node {node}: {v}")
                k = int(pars.init_pop[node] * v) if 0 < v < 1 else min(int(v), pars.init_pop[node])
                pool = np.where((nid == node) & alive)[0]
                picks.extend(np.random.choice(pool, size=min(k, len(pool)), replace=False))
            chosen = np.asarray(picks, dtype=np.int64)
        else:
            raise ValueError(f"Unsupported init_prev type: {type(prev)}")
        people.disease_state[chosen] = 2

    # ------------------------------------------------------------------ per tick
    def step(self):
        from . import kernels as K

        sim = self.sim
        dev = _need_dev(sim)
        t, c = sim.t, dev.cols
        K.disease_state_step(
            c["node_id"], dev.n_nodes, c["disease_state"], c["strain"], self.people.count, c["exposure_timer"],
            c["infection_timer"], c["potentially_paralyzed"], c["paralyzed"], c["ipv_protected"], c["paralysis_timer"],
            np.float32(self.pars.p_paralysis), dev.res["new_potentially_paralyzed"][t], dev.res["new_paralyzed"][t],
            rng=sim.rng(),
        )
        if t in self.seed_schedule:
            self._apply_seed_schedule(t, dev)
        if self.pars["stop_if_no_cases"]:
            # reference model.py:789-795: E/I of the previous tick and pending seeds decide; costs one device sync per tick
            ei = dev.res["E"][t - 1].sum() + dev.res["I"][t - 1].sum()
            if sim.shard is not None and sim.shard.world > 1:
                import torch.distributed as dist

                dist.all_reduce(ei, group=sim.shard.group)
            active = int(ei.item()) > 0
            if not (active or any(ts > t for ts in self.seed_schedule)):
                sim.should_stop = True

    def _apply_seed_schedule(self, t, dev):
        """Scheduled importations (reference model.py:759-779): host picks, device state is patched in place."""
        import torch

        count = dev.sync_count()  # cohorts born on the device are part of the candidate pool
        state = dev.cols["disease_state"][:count].cpu().numpy()
        node_id = dev.cols["node_id"][:count].cpu().numpy()
        for node, value in self.seed_schedule[t]:
            if self.sim.shard is not None and not self.sim.shard.owns(node):
                continue  # another rank owns this node
            pool = np.where((node_id == node) & (state >= 0))[0]
            if isinstance(value, float):
                k = int(len(pool) * value)
            elif isinstance(value, int):
                k = min(value, len(pool))
            else:
                raise ValueError(f"Unsupported seed value type: {type(value)}")
            if k <= 0:
                continue
            chosen = np.random.choice(pool, size=k, replace=False)
            idx = torch.from_numpy(chosen).to(dev.device)
            dev.cols["disease_state"][idx] = 2
            state[chosen] = 2
            timers = dev.cols["infection_timer"][idx]
            spent = idx[timers <= 0]  # previously infected: needs a fresh infectious period
            if spent.numel() > 0:
                fresh = np.asarray(self.pars.dur_inf(int(spent.numel())))
                dev.cols["infection_timer"][spent] = torch.from_numpy(fresh).to(dev.device).to(torch.int8)
            if self.verbose >= 1:
                print(f"[DiseaseState_ABM] t={t}: Seeded {k} infections in node {node}")

    def log(self, t):
        pass

    def plot(self, save=False, results_path=None):
        pass


# =========================================================================== heterogeneity (init-time, host)
def populate_heterogeneous_values(start, end, acq_risk_out, infectivity_out, pars):
    """Correlated lognormal acquisition risk / exponential infectivity via a Gaussian copula (reference model.py:816-866)."""
    from scipy import stats

    var = pars.risk_mult_var
    mu_ln = np.log(1.0 / np.sqrt(var + 1.0))
    sigma_ln = np.sqrt(np.log(var + 1.0))
    mean_inf = pars.r0 / np.mean(pars.dur_inf(1000))
    scale = max(mean_inf, 1e-10)
    rho = 2.0 * np.sin(np.pi * pars.corr_risk_inf / 6)
    chol = np.linalg.cholesky(np.array([[1, rho], [rho, 1]]))
    for lo in range(start, end, 1_000_000):
        hi = min(lo + 1_000_000, end)
        z = np.random.normal(size=(hi - lo, 2)) @ chol.T
        if pars.individual_heterogeneity:
            acq_risk_out[lo:hi] = np.exp(mu_ln + sigma_ln * z[:, 0])
            infectivity_out[lo:hi] = stats.gamma.ppf(stats.norm.cdf(z[:, 1]), a=1, scale=scale)
        else:
            acq_risk_out[lo:hi] = 1.0
            infectivity_out[lo:hi] = mean_inf


# =========================================================================== Transmission_ABM
class Transmission_ABM:
    """Per-node infectivity tally, network transfer, exposure, census (reference model.py:1152-1490)."""

    def _wire(self, sim):
        self.sim, self.people, self.pars, self.results = sim, sim.people, sim.pars, sim.results
        self.nodes = np.arange(len(sim.pars.init_pop))
        self.verbose = _verbosity(self.pars)

    def __init__(self, sim):
        self._wire(sim)
        self.r0_scalars = np.array(self.pars.r0_scalars)
        cap = self.people.capacity
        self.people.add_scalar_property("acq_risk_multiplier", dtype=np.float32, default=1.0)
        self.people.add_scalar_property("daily_infectivity", dtype=np.float32, default=1.0)
        if _device_init(self.pars):
            popinit.init_frame_device(self.people, self.pars, {"heterogeneity"})
        else:
            populate_heterogeneous_values(0, cap, self.people.acq_risk_multiplier, self.people.daily_infectivity, self.pars)
        self._init_common()

    @classmethod
    def init_from_file(cls, sim):
        self = cls.__new__(cls)
        self._wire(sim)
        self.r0_scalars = np.array(self.pars.r0_scalars)
        if "old_r0" in self.pars and self.pars.r0 != self.pars.old_r0:  # reference model.py:1182-1185
            self.people.daily_infectivity *= self.pars.r0 / self.pars.old_r0
        self._init_common()
        return self

    def _init_common(self):
        pars, n = self.pars, len(self.sim.nodes)
        init_pops = np.asarray(pars.init_pop)
        if _device_init(pars) and n <= 8192:  # distances, gravity / radiation, row normalisation on the GPU (netbuild.py)
            self.network = netbuild.build_network(pars, init_pops).cpu().numpy()
            self._init_results_rows()
            return
        if pars.distances is not None:
            dist = np.asarray(pars.distances)
        else:  # Haversine all-pairs from node_lookup (reference model.py:1224-1240), vectorised
            ids = sorted(pars.node_lookup.keys())
            lat = np.array([pars.node_lookup[i]["lat"] for i in ids])
            lon = np.array([pars.node_lookup[i]["lon"] for i in ids])
            dist = core.distance(lat[:, None], lon[:, None], lat[None, :], lon[None, :])
            off = ~np.eye(n, dtype=bool)
            dist[off & (dist == 0)] = 1  # coincident nodes: epsilon of 1 km
        method = pars.migration_method.lower()
        if method == "gravity":
            k = pars.gravity_k * 10 ** pars.gravity_k_exponent
            net = core.gravity(init_pops, dist, k, pars.gravity_a, pars.gravity_b, pars.gravity_c)
            net /= np.power(init_pops.sum(), pars.gravity_c)
        elif method == "radiation":
            net = core.radiation(init_pops, dist, 10**pars.radiation_k_log10, include_home=False)
        else:
            raise ValueError(f"Unknown migration method: {pars.migration_method}")
        self.network = core.row_normalizer(net, pars.max_migr_frac)
        self._init_results_rows()

    def _init_results_rows(self):
        pars = self.pars
        nt, ns = self.sim.nt, len(pars.strain_ids)
        self.results.add_array_property("new_exposed", shape=(nt, len(self.nodes)), dtype=np.int32)
        self.results.add_array_property("new_exposed_by_strain", shape=(nt, len(self.nodes), ns), dtype=np.int32)
        self.step_stats = utils.TimingStats()

    def step(self):
        from . import kernels as K

        sim, pars = self.sim, self.pars
        dev = _need_dev(sim)
        t, c, count = sim.t, dev.cols, self.people.count
        n, ns = dev.n_nodes, dev.n_strains
        srs = list(pars.strain_r0_scalars.values())  # positional, like the reference (model.py:1289)
        season = float(utils.get_seasonality(sim))
        K.tx_step_prep(n, count, ns, c["strain"], srs, c["disease_state"], c["node_id"], c["daily_infectivity"],
                       c["acq_risk_multiplier"], out=dev.tally)
        beta_fx, exposure_fx, _, risk_hist = dev.tally
        if sim.shard is not None:  # the one per-tick exchange: every node's infectivity feeds the network transfer
            from . import sharding

            sharding.allreduce_tally(beta_fx, sim.shard)
        r0s = self._r0_scalars_dev(dev)
        # results.pop[t] as it stands (all zeros when VitalDynamics_ABM is not a component -> divide by max(0, 1))
        pop = dev.pop_row(t)
        q, cdf, _, _ = K.tx_node_math(beta_fx, exposure_fx, risk_hist, dev.network_tensor(self.network), season, r0s, pop,
                                      float(pars.node_seeding_zero_inflation), float(pars.node_seeding_dispersion),
                                      rng=sim.rng(), out=dev.node_out)
        K.tx_infect(n, count, ns, c["node_id"], c["strain"], c["disease_state"], c["acq_risk_multiplier"], q, cdf,
                    rng=sim.rng(), out=dev.n_new)
        dev.res["new_exposed"][t] += dev.n_new.sum(dim=1, dtype=dev.n_new.dtype)  # += : RI / SIA exposures share the row
        dev.res["new_exposed_by_strain"][t] += dev.n_new

    def _r0_scalars_dev(self, dev):
        import torch

        src = self.r0_scalars
        if getattr(self, "_r0_src", None) is not src or getattr(self, "_r0_dev_owner", None) is not dev:
            # the reference broadcasts r0_scalars[:, None] against [nodes, strains] (model.py:1341): length 1 is legal
            arr = np.broadcast_to(np.asarray(src, dtype=np.float64), (dev.n_nodes,)).copy()
            self._r0_dev = torch.from_numpy(arr).to(dev.device)
            self._r0_src, self._r0_dev_owner = src, dev
        return self._r0_dev

    def log(self, t):
        """Census of the post-transmission state into row t (reference model.py:1461-1483)."""
        from . import kernels as K

        dev = _need_dev(self.sim)
        c, r = dev.cols, dev.res
        S, E, I, R, Ebs, Ibs, PP, P = K.count_SEIRP(  # noqa: E741
            c["node_id"], c["disease_state"], c["strain"], c["potentially_paralyzed"], c["paralyzed"], dev.n_nodes,
            dev.n_strains, self.people.count, out=dev.census)
        r["S"][t] = S
        r["E"][t] = E
        r["I"][t] = I
        r["E_by_strain"][t] = Ebs
        r["I_by_strain"][t] = Ibs
        r["R"][t] += R  # on top of the pre-seeded non-agent immunes
        r["potentially_paralyzed"][t] = PP
        r["paralyzed"][t] = P

    def plot(self, save=False, results_path=""):
        pass


# =========================================================================== VitalDynamics_ABM
class VitalDynamics_ABM:
    """Deaths by date_of_death and births by crude birth rate every ``step_size`` ticks (reference model.py:1521-1781)."""

    def _wire(self, sim):
        self.sim, self.people, self.nodes, self.results, self.pars = sim, sim.people, sim.nodes, sim.results, sim.pars
        self.step_size = self.pars.step_size_VitalDynamics_ABM
        self.verbose = _verbosity(self.pars)

    def __init__(self, sim):
        self._wire(sim)
        self._init_ages()
        self._init_deaths()
        self._init_birth_rates()

    @classmethod
    def init_from_file(cls, sim):
        self = cls.__new__(cls)
        self._wire(sim)
        for name in ("births", "deaths"):
            if name not in self.results.__dict__:
                self.results.add_array_property(name, shape=(sim.nt, len(self.nodes)), dtype=np.int32)
        self._init_birth_rates()
        self.death_estimator = core.KaplanMeierEstimator(utils.create_cumulative_deaths(np.sum(self.pars.init_pop), max_age_years=100))
        return self

    def _init_ages(self):
        people, pars = self.people, self.pars
        people.add_scalar_property("date_of_birth", dtype=np.int32, default=-1)
        if pars.init_sus_by_age is not None:
            for node in self.nodes:
                pyr = pars.init_sus_by_age[pars.init_sus_by_age["node_id"] == node].reset_index(drop=True)
                bins = core.AliasedDistribution(pyr["n_susceptible"]).sample(pars.init_sus[node])
                lo = (pyr["age_min_yr"] * 365).astype(int).to_numpy()
                hi = (pyr["age_max_yr"] * 365).astype(int).to_numpy()
                ages = np.random.randint(lo[bins], hi[bins]).astype(np.int32)
                ages[ages <= 0] = 1
                people.date_of_birth[np.where(people.node_id[: people.count] == node)[0]] = -ages
        elif _device_init(pars):
            self._pyramid = core.load_pyramid_csv(pars.age_pyramid_path)
            popinit.init_frame_device(people, pars, {"demography"}, pyramid=self._pyramid)
            self.sim._device_demography = True  # date_of_death / ri_timer continue the same per-agent streams
        else:
            pyr = core.load_pyramid_csv(pars.age_pyramid_path)
            bins = core.AliasedDistribution(pyr[:, 2] + pyr[:, 3]).sample(people.count)
            lo = np.maximum(pyr[:, 0] * 365, 1)  # nobody born on day 0
            hi = (pyr[:, 1] + 1) * 365
            ages = np.random.randint(lo[bins], hi[bins]).astype(np.int32)
            ages[ages == 0] = 1
            people.date_of_birth[: people.count] = -ages

    def _init_deaths(self):
        people, pars = self.people, self.pars
        if pars.cbr is None:
            return
        nt, nn = self.sim.nt, len(self.nodes)
        self.results.add_array_property("births", shape=(nt, nn), dtype=np.int32)
        self.results.add_array_property("deaths", shape=(nt, nn), dtype=np.int32)
        people.add_scalar_property("date_of_death", dtype=np.int32, default=0)
        self.death_estimator = core.KaplanMeierEstimator(utils.create_cumulative_deaths(np.sum(pars.init_pop), max_age_years=100))
        ages = -people.date_of_birth[: people.count]
        if getattr(self.sim, "_device_demography", False):
            self._cum_deaths = utils.create_cumulative_deaths(np.sum(pars.init_pop), max_age_years=100)
            popinit.init_frame_device(people, pars, {"demography"}, pyramid=self._pyramid, cum_deaths=self._cum_deaths)
            self.sim._device_cum_deaths = self._cum_deaths
            lifespans = people.date_of_death[: people.count] + ages
        else:
            lifespans = self.death_estimator.predict_age_at_death(ages, max_year=100)
            people.date_of_death[: people.count] = lifespans - ages
        nid = people.node_id[: people.count]
        cnt = np.bincount(nid, minlength=nn)
        life = np.zeros(nn)
        np.divide(np.bincount(nid, weights=lifespans / 365, minlength=nn), cnt, out=life, where=cnt > 0)
        pars.life_expectancies = life
        if pars.init_sus_by_age is not None:
            # pre-modelled mortality of the non-agent immunes (reference model.py:1651-1669)
            df = pars.init_sus_by_age
            df["avg_age_yr"] = (df["age_min_yr"] + df["age_max_yr"]) / 2
            df["life_expectancy_yr"] = df["node_id"].map(lambda n: pars.life_expectancies[n])
            df["remaining_life_yr"] = np.maximum(df["life_expectancy_yr"] - df["avg_age_yr"], 1.0)
            df["daily_mortality_rate"] = 1 / (df["remaining_life_yr"] * 365)
            df["expected_deaths_per_day"] = df["n_immune"] * df["daily_mortality_rate"]
            per_node = df.groupby("node_id")["expected_deaths_per_day"].sum().to_numpy()
            self.results.deaths += np.random.poisson(np.outer(np.ones(nt), per_node)).astype(np.int32)
            self.results.deaths[0, :] = 0

    def _init_birth_rates(self):
        cbr = self.pars.cbr
        self.birth_rate = np.zeros(len(self.nodes))
        if cbr is not None:
            self.birth_rate[:] = (cbr[0] if (isinstance(cbr, (float, int)) or len(cbr) == 1) else np.array(cbr)) / (365 * 1000)

    def births_args(self, dev, t, tile_node=None, tallies=None, hot=None):
        """Argument block of lpk_vd_births for tick t (device-side births, include/lpk.h V2)."""
        import torch

        from . import _lpk

        if getattr(self, "_dev_owner", None) is not dev:
            cd = np.ascontiguousarray(self.death_estimator._cd, dtype=np.int64)
            self._cd_dev = torch.from_numpy(cd).to(dev.device)
            rate = np.ascontiguousarray(self.birth_rate, dtype=np.float64).copy()
            if self.sim.shard is not None:
                rate[~self.sim.shard.owned_mask(len(rate))] = 0.0  # other ranks create the cohorts of their own nodes
            self._rate_dev = torch.from_numpy(rate).to(dev.device)
            self._dev_owner = dev
        c, r, sim = dev.cols, dev.res, self.sim
        a = _lpk.BirthsArgs()
        a.tick, a.n_nodes = t, dev.n_nodes
        a.seed, a.id_base = int(sim.pars.seed) & 0xFFFFFFFFFFFFFFFF, sim.id_base
        a.step_size = float(self.step_size)
        a.birth_rate, a.pop_prev, a.births_row = self._rate_dev.data_ptr(), r["pop"][t - 1].data_ptr(), r["births"][t].data_ptr()
        a.counts, a.capacity = dev.counts.data_ptr(), dev.cap_eff  # (the graveyard of a compacted table is not available for births)
        a.cum_deaths, a.max_year = self._cd_dev.data_ptr(), min(100, len(self.death_estimator._cd) - 2)
        a.ri_newborn_timer = 182 if (getattr(self.pars, "ri_newborn_timer", False) and "ri_timer" in c) else -1
        a.node_offsets_ws, a.cohort_ws, a.status = dev.node_offsets_ws.data_ptr(), dev.cohort_ws.data_ptr(), dev.status.data_ptr()
        a.disease_state, a.node_id = c["disease_state"].data_ptr(), c["node_id"].data_ptr()
        a.date_of_birth, a.date_of_death = c["date_of_birth"].data_ptr(), c["date_of_death"].data_ptr()
        a.ri_timer = c["ri_timer"].data_ptr() if "ri_timer" in c else None
        a.tile_node = tile_node.data_ptr() if tile_node is not None else None
        if tallies is not None:  # fused path: the cohort joins the carried susceptible-side tallies
            sus, expo, hist = tallies
            a.acq_risk_multiplier = c["acq_risk_multiplier"].data_ptr()
            a.sus, a.exposure_fx, a.risk_hist = sus.data_ptr(), expo.data_ptr(), hist.data_ptr()
        if hot is not None:  # fused path: newborns get their agenda byte (csrc/lpk_hot.cuh)
            hot_bytes, pair_min_dod, e0, pair_ri_max, ri_k, ri_step, ri_kcol, rec = hot
            a.rec = rec.data_ptr()
            for name in ("strain", "exposure_timer", "infection_timer", "paralysis_timer", "potentially_paralyzed", "paralyzed",
                         "ipv_protected"):
                setattr(a, name, c[name].data_ptr())
            a.ri_k = ri_kcol.data_ptr() if ri_kcol is not None else None
            a.hot = hot_bytes.data_ptr()
            a.pair_min_dod = pair_min_dod.data_ptr() if pair_min_dod is not None else None
            a.risk_e0 = e0
            a.pair_ri_max = pair_ri_max.data_ptr() if pair_ri_max is not None else None
            a.ri_lazy_k, a.ri_step = ri_k, ri_step
        self._keep = (a,)
        return a

    def step(self):
        import ctypes as C

        from . import _lpk
        from . import kernels as K

        sim, people = self.sim, self.people
        dev = _need_dev(sim)
        t, r, c = sim.t, dev.res, dev.cols
        if t % self.step_size != 0:
            r["pop"][t] = r["pop"][t - 1]  # no births or deaths this tick (reference model.py:1691-1695)
            return
        if self.pars.cbr is None:
            raise ValueError("VitalDynamics_ABM needs pars.cbr")
        dying = dev.scratch_i32[0]
        K.get_deaths(dev.n_nodes, people.count, c["disease_state"], c["node_id"], c["date_of_death"], t, dying)
        # births (reference model.py:1712-1734) are drawn and appended on the device: no host round trip per step
        dev.set_count(people.count)
        args = self.births_args(dev, t, dev.tile_node)
        K.STATS.record("vd_births", lambda: _lpk.check(_lpk.lib().lpk_vd_births(C.byref(args), _lpk.stream_handle()), "lpk_vd_births"), 3)
        r["deaths"][t] = dying  # "=": overwrites the pre-modelled deaths of non-agent immunes (model.py:1749)
        r["pop"][t] = r["pop"][t - 1] + r["births"][t] - r["deaths"][t]
        dev.sync_count()  # component-by-component mode launches with a host-side count

    def log(self, t):
        pass

    def plot(self, save=False, results_path=None):
        pass


# =========================================================================== RI_ABM
class RI_ABM:
    """Routine immunisation every ``step_size_RI_ABM`` ticks (reference model.py:1858-1991)."""

    def _wire(self, sim):
        self.sim, self.people, self.nodes, self.pars, self.results = sim, sim.people, sim.nodes, sim.pars, sim.results
        self.step_size = sim.pars.step_size_RI_ABM
        self.verbose = _verbosity(self.pars)
        nt, nn, ns = sim.nt, len(sim.nodes), len(self.pars.strain_ids)
        for name in ("ri_vaccinated", "ri_protected", "ipv_vaccinated"):
            self.results.add_array_property(name, shape=(nt, nn), dtype=np.int32)
        self.results.add_array_property("ri_new_exposed_by_strain", shape=(nt, nn, ns), dtype=np.int32)
        self._prob_cache = None

    def __init__(self, sim):
        self._wire(sim)
        people = self.people
        people.add_scalar_property("ri_timer", dtype=np.int16, default=-1)
        dob = people.date_of_birth[: people.count]
        if getattr(sim, "_device_demography", False):
            vd = next((i for i in getattr(sim, "instances", []) if isinstance(i, VitalDynamics_ABM)), None)
            pyramid = getattr(vd, "_pyramid", None) if vd is not None else None
            if pyramid is None:
                pyramid = core.load_pyramid_csv(sim.pars.age_pyramid_path)
            popinit.init_frame_device(people, sim.pars, {"demography"}, pyramid=pyramid,
                                      cum_deaths=getattr(sim, "_device_cum_deaths", None))
        else:
            people.ri_timer[: people.count] = (dob + np.random.uniform(42, 98, people.count)).astype(np.int32)

    @classmethod
    def init_from_file(cls, sim):
        self = cls.__new__(cls)
        self._wire(sim)
        return self

    def _probs(self, dev):
        import torch

        pars, n = self.pars, len(self.sim.nodes)
        key = (id(pars["vx_prob_ri"]), id(pars["vx_prob_ipv"]), id(dev))
        if self._prob_cache is None or self._prob_cache[0] != key:
            ri = pars["vx_prob_ri"]
            ipv = pars["vx_prob_ipv"] if pars["vx_prob_ipv"] is not None else np.zeros(n)
            ri = np.full(n, ri, dtype=np.float64) if np.isscalar(ri) else np.ascontiguousarray(ri, dtype=np.float64)
            ipv = np.full(n, ipv, dtype=np.float64) if np.isscalar(ipv) else np.ascontiguousarray(ipv, dtype=np.float64)
            self._prob_cache = (key, torch.from_numpy(ri).to(dev.device), torch.from_numpy(ipv).to(dev.device))
        return self._prob_cache[1], self._prob_cache[2]

    def step(self):
        from . import kernels as K

        pars, sim = self.pars, self.sim
        if pars["vx_prob_ri"] is None:
            return
        if sim.t % self.step_size != 0:
            return
        dev = _need_dev(sim)
        vtype = getattr(pars, "ri_vaccine_type", "tOPV")
        vstrain = 2 if "nOPV" in vtype else 1
        p_ri, p_ipv = self._probs(dev)
        c, r, t = dev.cols, dev.res, sim.t
        n_ri, n_prot, n_ipv = dev.scratch_i32[1], dev.scratch_i32[2], dev.scratch_i32[3]
        K.fast_ri(self.step_size, c["node_id"], c["disease_state"], c["strain"], c["ipv_protected"], c["ri_timer"], t, p_ri,
                  p_ipv, self.people.count, n_ri, n_prot, n_ipv, c["chronically_missed"], vstrain, rng=sim.rng())
        r["ri_vaccinated"][t] = n_ri
        r["ri_protected"][t] = n_prot
        r["new_exposed"][t] += n_prot
        r["new_exposed_by_strain"][t, :, vstrain] += n_prot
        r["ri_new_exposed_by_strain"][t, :, vstrain] = n_prot
        r["ipv_vaccinated"][t] = n_ipv

    def log(self, t):
        pass

    def plot(self, save=False, results_path=None):
        pass


# =========================================================================== SIA_ABM
class SIA_ABM:
    """Scheduled vaccination campaigns (reference model.py:2063-2162)."""

    def __init__(self, sim):
        self.sim, self.people, self.nodes, self.pars, self.results = sim, sim.people, sim.nodes, sim.pars, sim.results
        self.verbose = _verbosity(self.pars)
        nt, nn, ns = sim.nt, len(self.nodes), len(self.pars.strain_ids)
        self.results.add_array_property("sia_vaccinated", shape=(nt, nn), dtype=np.int32)
        self.results.add_array_property("sia_protected", shape=(nt, nn), dtype=np.int32)
        self.results.add_array_property("sia_new_exposed_by_strain", shape=(nt, nn, ns), dtype=np.int32)
        schedule = self.pars["sia_schedule"] if "sia_schedule" in self.pars and self.pars["sia_schedule"] is not None else []
        self.sia_schedule = schedule
        self._by_tick = defaultdict(list)  # the reference scans the whole list every tick (model.py:2103-2104)
        for event in schedule:
            event["date"] = utils.date(event["date"])
            k = (event["date"] - sim.datevec[0]).days
            if 0 <= k < sim.nt:
                self._by_tick[k].append(event)
        self._vx_cache = None

    @classmethod
    def init_from_file(cls, sim):
        return cls(sim)

    def event_args(self, dev, event):
        """Device-side arguments of one campaign event (model.py:2106-2140): targeted-node mask (uploaded once per
        event), per-node vaccination probability, efficacy, age window in days, vaccine strain."""
        import torch

        if self._vx_cache is None or self._vx_cache[0] is not dev:
            vx = np.ascontiguousarray(np.array(self.pars["vx_prob_sia"], dtype=np.float32))
            self._vx_cache = (dev, torch.from_numpy(vx).to(dev.device), {})
        _, vx_prob, masks = self._vx_cache
        targeted = masks.get(id(event))
        if targeted is None:
            host = np.zeros(len(self.sim.nodes), np.uint8)
            host[event["nodes"]] = 1
            targeted = masks[id(event)] = torch.from_numpy(host).to(dev.device)
        vtype = event["vaccinetype"]
        lo, hi = event["age_range"]
        return targeted, vx_prob, self.pars["vx_efficacy"][vtype], lo, hi, (2 if "nOPV" in vtype else 1)

    def step(self):
        from . import kernels as K

        sim, pars = self.sim, self.pars
        t = sim.t
        events = self._by_tick.get(t)
        if not events or pars.vx_prob_sia is None:
            return
        dev = _need_dev(sim)
        c, r = dev.cols, dev.res
        vacc, prot = dev.scratch_i32[1], dev.scratch_i32[2]
        for k, event in enumerate(events):
            targeted, vx_prob, vx_eff, lo, hi, vstrain = self.event_args(dev, event)
            K.fast_sia(c["node_id"], c["disease_state"], c["strain"], c["date_of_birth"], t, vx_prob, float(vx_eff),
                       self.people.count, targeted, int(lo), int(hi), vacc, prot,
                       c["chronically_missed"], vstrain, event_idx=k, rng=sim.rng())
            r["sia_vaccinated"][t] = vacc  # overwritten per event, like the reference (model.py:2142-2145)
            r["sia_protected"][t] = prot
            r["new_exposed"][t] += prot
            r["new_exposed_by_strain"][t, :, vstrain] += prot
            r["sia_new_exposed_by_strain"][t, :, vstrain] += prot

    def log(self, t):
        pass

    def plot(self, save=False, results_path=None):
        pass
