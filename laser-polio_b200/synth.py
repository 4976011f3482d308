"""Synthetic populations of the shapes BASELINE.json names.

The real population inputs of the reference are large blobs that are not
shipped, so parity tests and benchmarks run on seeded synthetic agent tables
whose column set, dtypes and distributions follow the reference's initialisers:

* column set / dtypes: reference model.py:144-162, 571-573, 1201-1206, 1550, 1603, 1891
* acq_risk_multiplier ~ LogN(mean 1, var 4), daily_infectivity ~ Exp(r0 / E[dur_inf]) via a
  Gaussian copula with rho = 2 sin(pi * 0.8 / 6)  (model.py:839-863, pars.py:38-43)
* exposure_timer ~ Poisson(3), infection_timer ~ Gamma(4.51, 5.32) truncated to int8 [0, 127],
  paralysis_timer = clip(LogN(12.5, 3.5) - etimer, 0, itimer)  (model.py:575-587)
* node sizes ~ lognormal(sigma = 1), agents stored node-contiguous (model.py:171-172)

``synth_population`` builds numpy columns on the host (oracle side, small sizes);
``synth_population_device`` builds the same distributions directly in HBM with
torch for the full-size configurations.
"""

from __future__ import annotations

import math

import numpy as np

COLUMNS = {
    "disease_state": np.int8,
    "potentially_paralyzed": np.int8,
    "paralyzed": np.int8,
    "ipv_protected": np.int8,
    "strain": np.int8,
    "chronically_missed": np.uint8,
    "node_id": np.int16,
    "exposure_timer": np.int8,
    "infection_timer": np.int8,
    "paralysis_timer": np.int8,
    "acq_risk_multiplier": np.float32,
    "daily_infectivity": np.float32,
    "date_of_birth": np.int32,
    "date_of_death": np.int32,
    "ri_timer": np.int16,
}
COLUMN_DEFAULTS = {
    "disease_state": -1,
    "potentially_paralyzed": -1,
    "paralyzed": 0,
    "ipv_protected": 0,
    "strain": 0,
    "chronically_missed": 0,
    "node_id": -1,
    "exposure_timer": 0,
    "infection_timer": 0,
    "paralysis_timer": 0,
    "acq_risk_multiplier": 1.0,
    "daily_infectivity": 1.0,
    "date_of_birth": -1,
    "date_of_death": 0,
    "ri_timer": -1,
}
BYTES_PER_AGENT = sum(np.dtype(d).itemsize for d in COLUMNS.values())  # 29

# named shapes
SHAPES = {
    "zamfara": {"n_nodes": 14, "n_agents": 5_000_000, "ticks": 365},
    "nigeria": {"n_nodes": 774, "n_agents": 220_000_000, "ticks": 2555},
    "west_africa": {"n_nodes": 1921, "n_agents": 430_000_000, "ticks": 2655},
    "africa": {"n_nodes": 5672, "n_agents": 1_300_000_000, "ticks": 1095},
}


def node_sizes(n_agents: int, n_nodes: int, rng: np.random.Generator) -> np.ndarray:
    """Heavy-tailed node populations (lognormal sigma=1) that sum exactly to n_agents, each >= 1."""
    w = rng.lognormal(0.0, 1.0, n_nodes)
    sizes = np.maximum(1, np.floor(w / w.sum() * (n_agents - n_nodes)).astype(np.int64) + 1)
    sizes[np.argmax(sizes)] += n_agents - sizes.sum()
    assert sizes.sum() == n_agents and sizes.min() >= 1
    return sizes


def synth_population(
    n_agents: int,
    n_nodes: int,
    seed: int = 0,
    capacity: int | None = None,
    f_exposed: float = 0.01,
    f_infected: float = 0.01,
    f_recovered: float = 0.05,
    f_dead: float = 0.0,
    r0: float = 14.0,
    n_strains: int = 3,
    missed_frac: float = 0.1,
    ipv_frac: float = 0.3,
    max_age_days: int = 15 * 365,
    sorted_nodes: bool = True,
) -> dict:
    """Host (numpy) synthetic agent table; returns {"count", "capacity", "n_nodes", column: array...}."""
    rng = np.random.default_rng(seed)
    capacity = int(capacity or n_agents)
    n = int(n_agents)
    cols = {k: np.full(capacity, COLUMN_DEFAULTS[k], dtype=d) for k, d in COLUMNS.items()}

    sizes = node_sizes(n, n_nodes, rng)
    nid = np.repeat(np.arange(n_nodes, dtype=np.int16), sizes)
    if not sorted_nodes:
        rng.shuffle(nid)
    cols["node_id"][:n] = nid

    u = rng.random(n)
    state = np.zeros(n, np.int8)
    edges = np.cumsum([f_dead, f_exposed, f_infected, f_recovered])
    state[u < edges[0]] = -1
    state[(u >= edges[0]) & (u < edges[1])] = 1
    state[(u >= edges[1]) & (u < edges[2])] = 2
    state[(u >= edges[2]) & (u < edges[3])] = 3
    cols["disease_state"][:n] = state
    ei = (state == 1) | (state == 2)
    cols["strain"][:n] = np.where(ei, rng.integers(0, n_strains, n), 0).astype(np.int8)

    # whole-capacity per-slot draws, like the reference (timers/risk/infectivity exist for unborn slots too)
    et = np.clip(rng.poisson(3.0, capacity), 0, 127).astype(np.int8)
    it = np.clip(rng.gamma(4.51, 5.32, capacity).astype(np.int8), 0, 127).astype(np.int8)
    mu = math.log(12.5**2 / math.sqrt(3.5**2 + 12.5**2))
    sg = math.sqrt(math.log(3.5**2 / 12.5**2 + 1))
    raw = rng.lognormal(mu, sg, capacity) - et
    cols["exposure_timer"][:] = et
    cols["infection_timer"][:] = it
    cols["paralysis_timer"][:] = np.clip(raw, 0, np.minimum(it, 127)).astype(np.int8)
    # agents already in E/I are part-way through their timers
    prog = rng.random(n)
    cols["exposure_timer"][:n] = np.where(state == 1, (et[:n] * prog).astype(np.int8), et[:n])
    cols["infection_timer"][:n] = np.where(state == 2, (it[:n] * prog).astype(np.int8), it[:n])

    rho = 2.0 * math.sin(math.pi * 0.8 / 6.0)
    z1 = rng.standard_normal(capacity)
    z2 = rho * z1 + math.sqrt(1 - rho * rho) * rng.standard_normal(capacity)
    mu_ln = math.log(1.0 / math.sqrt(4.0 + 1.0))
    sg_ln = math.sqrt(math.log(4.0 + 1.0))
    cols["acq_risk_multiplier"][:] = np.exp(mu_ln + sg_ln * z1).astype(np.float32)
    mean_inf = r0 / (4.51 * 5.32)
    from scipy.special import ndtr

    cols["daily_infectivity"][:] = (-mean_inf * np.log1p(-np.clip(ndtr(z2), 0.0, 1 - 1e-16))).astype(np.float32)

    age = rng.integers(1, max_age_days, n)
    cols["date_of_birth"][:n] = -age
    # remaining life: geometric-ish so that ~2%/yr die; a few percent die inside a 7-year window
    cols["date_of_death"][:n] = rng.exponential(50 * 365.0, n).astype(np.int32) + 1
    cols["ri_timer"][:n] = (cols["date_of_birth"][:n] + rng.uniform(42, 98, n)).astype(np.int32).astype(np.int16)
    cols["chronically_missed"][:n] = (rng.random(n) < missed_frac).astype(np.uint8)
    cols["ipv_protected"][:n] = (rng.random(n) < ipv_frac).astype(np.int8)

    out = {"count": n, "capacity": capacity, "n_nodes": int(n_nodes), "n_strains": int(n_strains), "node_sizes": sizes}
    out.update(cols)
    return out


def synth_population_device(
    n_agents: int,
    n_nodes: int,
    seed: int = 0,
    capacity: int | None = None,
    device="cuda",
    f_exposed: float = 0.01,
    f_infected: float = 0.01,
    f_recovered: float = 0.05,
    f_dead: float = 0.0,
    r0: float = 14.0,
    n_strains: int = 3,
    missed_frac: float = 0.1,
    ipv_frac: float = 0.3,
    max_age_days: int = 15 * 365,
    sizes=None,
    node_offset: int = 0,
) -> dict:
    """Same distributions as :func:`synth_population`, generated directly in HBM with torch (seeded).

    ``sizes`` (optional): explicit agents per node for the ``n_nodes`` nodes of this table (a node shard of a larger
    population); node ids are ``node_offset + 0 .. n_nodes - 1``.


    Used for the full-size configurations (2.2e8 .. 1.3e9 agents) where a host-side build would
    dominate the run.  Values differ from the numpy generator (different RNG); shapes, dtypes and
    distributions are the same.  Returns {"count", "capacity", "n_nodes", "n_strains", "node_sizes", column: tensor}.
    """
    import torch

    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    capacity = int(capacity or n_agents)
    n = int(n_agents)
    sizes = node_sizes(n, n_nodes, np.random.default_rng(seed)) if sizes is None else np.asarray(sizes, dtype=np.int64)
    assert len(sizes) == n_nodes and int(sizes.sum()) == n
    tdt = {np.int8: torch.int8, np.uint8: torch.uint8, np.int16: torch.int16, np.int32: torch.int32, np.float32: torch.float32}
    cols = {k: torch.full((capacity,), COLUMN_DEFAULTS[k], dtype=tdt[d], device=dev) for k, d in COLUMNS.items()}

    cols["node_id"][:n] = torch.repeat_interleave(
        torch.arange(node_offset, node_offset + n_nodes, dtype=torch.int16, device=dev), torch.from_numpy(sizes).to(dev), output_size=n
    )
    u = torch.rand(n, generator=g, device=dev)
    e = np.cumsum([f_dead, f_exposed, f_infected, f_recovered])
    state = torch.zeros(n, dtype=torch.int8, device=dev)
    state[u < e[0]] = -1
    state[(u >= e[0]) & (u < e[1])] = 1
    state[(u >= e[1]) & (u < e[2])] = 2
    state[(u >= e[2]) & (u < e[3])] = 3
    cols["disease_state"][:n] = state
    ei = (state == 1) | (state == 2)
    cols["strain"][:n] = torch.where(ei, torch.randint(0, n_strains, (n,), generator=g, device=dev), 0).to(torch.int8)
    del u

    et = torch.poisson(torch.full((capacity,), 3.0, device=dev), generator=g).clamp_(0, 127).to(torch.int8)
    it = (torch._standard_gamma(torch.full((capacity,), 4.51, device=dev), generator=g) * 5.32).clamp_(0, 127).to(torch.int8)
    mu = math.log(12.5**2 / math.sqrt(3.5**2 + 12.5**2))
    sg = math.sqrt(math.log(3.5**2 / 12.5**2 + 1))
    raw = torch.empty(capacity, device=dev).log_normal_(mu, sg, generator=g) - et
    cols["paralysis_timer"][:] = torch.minimum(raw.clamp_(min=0), it.to(torch.float32)).to(torch.int8)
    prog = torch.rand(n, generator=g, device=dev)
    cols["exposure_timer"][:] = et
    cols["infection_timer"][:] = it
    cols["exposure_timer"][:n] = torch.where(state == 1, (et[:n] * prog).to(torch.int8), et[:n])
    cols["infection_timer"][:n] = torch.where(state == 2, (it[:n] * prog).to(torch.int8), it[:n])
    del raw, prog, et, it

    rho = 2.0 * math.sin(math.pi * 0.8 / 6.0)
    z1 = torch.randn(capacity, generator=g, device=dev)
    z2 = rho * z1 + math.sqrt(1 - rho * rho) * torch.randn(capacity, generator=g, device=dev)
    cols["acq_risk_multiplier"][:] = torch.exp(math.log(1.0 / math.sqrt(5.0)) + math.sqrt(math.log(5.0)) * z1)
    mean_inf = r0 / (4.51 * 5.32)
    cols["daily_infectivity"][:] = -mean_inf * torch.log1p(-torch.special.ndtr(z2.double()).clamp_(0.0, 1 - 1e-16)).float()
    del z1, z2

    age = torch.randint(1, max_age_days, (n,), generator=g, device=dev, dtype=torch.int32)
    cols["date_of_birth"][:n] = -age
    cols["date_of_death"][:n] = torch.empty(n, device=dev).exponential_(1.0 / (50 * 365.0), generator=g).to(torch.int32) + 1
    cols["ri_timer"][:n] = (-age + (42 + 56 * torch.rand(n, generator=g, device=dev)).to(torch.int32)).to(torch.int16)
    cols["chronically_missed"][:n] = (torch.rand(n, generator=g, device=dev) < missed_frac).to(torch.uint8)
    cols["ipv_protected"][:n] = (torch.rand(n, generator=g, device=dev) < ipv_frac).to(torch.int8)
    del age

    out = {"count": n, "capacity": capacity, "n_nodes": int(n_nodes), "n_strains": int(n_strains), "node_sizes": sizes}
    out.update(cols)
    return out


# ------------------------------------------------------------------------------------------ a runnable sim on a synthetic table
WORKLOAD_R0 = 1.1  # sets daily_infectivity in the synthetic table; see workload_pars


def campaign_schedule(start, n_nodes, years, rng):
    """~8 campaigns per year, under-5s, 30-100 % of nodes, mOPV2 then nOPV2 after year 3 (Nigeria)."""
    import datetime as dt

    events = []
    for y in range(years + 1):
        for k in range(8):
            day = y * 365 + 20 + k * 44
            frac = rng.uniform(0.3, 1.0)
            nodes = np.sort(rng.choice(n_nodes, size=max(1, int(frac * n_nodes)), replace=False)).tolist()
            events.append({"date": start + dt.timedelta(days=day), "nodes": nodes, "age_range": (0, 5 * 365),
                           "vaccinetype": "mOPV2" if y < 3 else "nOPV2"})
    return events


def campaign_day(t: int) -> bool:
    """Is tick t a campaign day of :func:`campaign_schedule`?"""
    d = t % 365
    return d >= 20 and (d - 20) % 44 == 0 and (d - 20) // 44 < 8


def workload_pars(sizes, dur, seed, rng, cbr=37.0, **over):
    """PropertySet of the full-feature workload on nodes of the given sizes: vital dynamics every 7 ticks, RI every 14,
    ~8 campaigns a year, gravity network with seasonality, three strains."""
    import datetime as dt

    from .core import PropertySet

    n = len(sizes)
    xy = rng.uniform(0, 1000.0, (n, 2))  # synthetic node coordinates, km
    dist = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
    dist[dist == 0] = 1.0
    np.fill_diagonal(dist, 0.0)
    start = dt.date(2017, 1, 1)
    p = {
        "seed": seed, "start_date": start, "dur": dur, "init_pop": np.asarray(sizes), "cbr": np.full(n, float(cbr)),
        # r0 chosen so that R_eff ~ 1 with 93 % susceptible agents: prevalence stays near the canonical mix of SURVEY 8(d)
        # (f_S 0.93, f_E = f_I 0.01) for the whole timed window instead of exploding (the real Nigeria runs divide the force
        # of infection by a population that is mostly non-agent immunes, model.py:1344-1347)
        "r0": WORKLOAD_R0, "r0_scalars": rng.uniform(0.8, 1.2, n), "seasonal_amplitude": 0.1, "seasonal_peak_doy": 159,
        "distances": dist, "migration_method": "gravity", "gravity_k": 0.5, "gravity_k_exponent": -1.0, "gravity_c": 1.5,
        "max_migr_frac": 0.1, "vx_prob_ri": rng.uniform(0.3, 0.8, n), "vx_prob_ipv": rng.uniform(0.3, 0.8, n),
        "vx_prob_sia": rng.uniform(0.4, 0.9, n).tolist(), "sia_schedule": campaign_schedule(start, n, dur // 365 + 1, rng),
        "stop_if_no_cases": False, "verbose": 0, "node_seeding_zero_inflation": 0.0, "node_seeding_dispersion": 1000,
    }
    p.update(over)
    return PropertySet(p)


def synth_sim(n_agents, n_nodes, dur, seed, device="cuda", rank=0, world=1, mode="shard", cbr=37.0, pars_over=None, pop_over=None):
    """A SEIR_ABM with the five stock components on a synthetic table generated in HBM and mirrored into host columns (the
    reference-facing LaserFrame), wrapped by SEIR_ABM.init_from_file + Component.init_from_file -- the reference's route for
    a pre-built table (run_sim.py:398-409).  Returns (sim, agents on this rank).

    n_agents / n_nodes are the totals over all ranks.  Every rank builds the node block it owns: in 'shard' mode contiguous
    blocks of ONE heavy-tailed size vector balanced by agents (sharding.plan_node_blocks), in 'weak' mode block r is the
    r-th copy of a per-GPU shape.  Node ids are global, the network spans all nodes."""
    import torch

    from . import abm, core, sharding

    if mode == "weak":
        per_nodes, per_agents = n_nodes // world, n_agents // world
        sizes_all = np.concatenate([node_sizes(per_agents, per_nodes, np.random.default_rng(seed + r)) for r in range(world)])
        blocks = [(r * per_nodes, (r + 1) * per_nodes) for r in range(world)]
    else:
        sizes_all = node_sizes(n_agents, n_nodes, np.random.default_rng(seed))
        blocks = sharding.plan_node_blocks(sizes_all, world) if world > 1 else [(0, n_nodes)]
    births_room = 1.0 + float(cbr) / 1000.0 * (dur + 100) / 365.0 * 1.15
    caps = [int(int(sizes_all[lo:hi].sum()) * births_room) + 4096 for lo, hi in blocks]
    lo, hi = blocks[rank]
    n_r, capacity = int(sizes_all[lo:hi].sum()), caps[rank]
    pop = synth_population_device(n_r, hi - lo, seed=seed + rank, capacity=capacity, device=device, r0=WORKLOAD_R0,
                                  sizes=sizes_all[lo:hi], node_offset=lo, **(pop_over or {}))
    people = core.LaserFrame(capacity=capacity, initial_count=n_r)
    for name, dtype in COLUMNS.items():
        people.add_scalar_property(name, dtype=dtype, default=COLUMN_DEFAULTS[name])
        torch.from_numpy(getattr(people, name)).copy_(pop[name])
    del pop
    torch.cuda.empty_cache()
    pars = workload_pars(sizes_all, dur, seed, np.random.default_rng(seed + 1000), cbr=cbr, **(pars_over or {}))
    sim = abm.SEIR_ABM.init_from_file(people, pars)
    sim.verbose = 0
    sim.nodes = np.arange(n_nodes)
    sim._components = [abm.VitalDynamics_ABM, abm.DiseaseState_ABM, abm.RI_ABM, abm.SIA_ABM, abm.Transmission_ABM]
    sim.instances = [c.init_from_file(sim) for c in sim._components]
    if world > 1:
        sim.shard = sharding.Shard(rank=rank, world=world, node_lo=lo, node_hi=hi)
        sim.id_base = sharding.id_bases(sizes_all, blocks, capacity_per_block=caps)[rank]
    return sim, n_r
