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
) -> dict:
    """Same distributions as :func:`synth_population`, generated directly in HBM with torch (seeded).

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
    sizes = node_sizes(n, n_nodes, np.random.default_rng(seed))
    tdt = {np.int8: torch.int8, np.uint8: torch.uint8, np.int16: torch.int16, np.int32: torch.int32, np.float32: torch.float32}
    cols = {k: torch.full((capacity,), COLUMN_DEFAULTS[k], dtype=tdt[d], device=dev) for k, d in COLUMNS.items()}

    cols["node_id"][:n] = torch.repeat_interleave(
        torch.arange(n_nodes, dtype=torch.int16, device=dev), torch.from_numpy(sizes).to(dev), output_size=n
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
