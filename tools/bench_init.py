#!/usr/bin/env python
"""Population initialisation and network construction: device kernels (popinit / netbuild) timed with CUDA events at the
Nigeria shape, next to the host numpy path the reference uses (abm.populate_heterogeneous_values etc.) and the C oracle on
a bounded sample.  usage: tools/bench_init.py [agents] [nodes]   -> one JSON line"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import laser_polio_b200 as lp  # noqa: E402
from laser_polio_b200 import abm, core, netbuild, popinit, utils  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 220_000_000
nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 774
pars = lp.PropertySet(dict(seed=1, risk_mult_var=4.0, r0=14.0, corr_risk_inf=0.8, individual_heterogeneity=True, dur_exp=lp.poisson(lam=3),
                           dur_inf=lp.gamma(shape=4.51, scale=5.32), t_to_paralysis=lp.lognormal(mean=12.5, sigma=3.5), missed_frac=0.1))
pyr = np.array([[5 * k, 5 * k + 4, int(1.7e7 * np.exp(-0.16 * k)), int(1.6e7 * np.exp(-0.16 * k))] for k in range(20)] + [[100, 100, 300, 500]])
cum = utils.create_cumulative_deaths(n, max_age_years=100)
dev = "cuda"
r, f = (torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2))
et, it, pt = (torch.empty(n, dtype=torch.int8, device=dev) for _ in range(3))
dob, dod = (torch.empty(n, dtype=torch.int32, device=dev) for _ in range(2))
ri = torch.empty(n, dtype=torch.int16, device=dev)
ms = torch.empty(n, dtype=torch.uint8, device=dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


out = {"agents": n, "nodes": nodes, "device_ms": {}, "bytes_written_per_agent": {"heterogeneity": 8, "timers": 3, "demography": 10, "missed": 1}}
out["device_ms"]["heterogeneity"] = timed(lambda: popinit.populate_heterogeneous_values(0, n, r, f, pars, mean_dur_inf=24.0))
out["device_ms"]["timers"] = timed(lambda: popinit.init_timers(0, n, et, it, pt, pars))
out["device_ms"]["demography"] = timed(lambda: popinit.init_demography(0, n, dob, dod, ri, pyr, cum, 1))
out["device_ms"]["missed"] = timed(lambda: popinit.init_missed(n, n // 10, ms, 1))
rng = np.random.default_rng(0)
lat, lon, pops = rng.uniform(4, 14, nodes), rng.uniform(3, 15, nodes), np.round(np.exp(rng.normal(11, 1, nodes)))
lookup = {i: {"lat": float(lat[i]), "lon": float(lon[i])} for i in range(nodes)}
npars = lp.PropertySet({"distances": None, "node_lookup": lookup, "migration_method": "radiation", "radiation_k_log10": -0.3, "max_migr_frac": 0.1})
out["device_ms"]["network_radiation"] = timed(lambda: netbuild.build_network(npars, pops))
out["device_total_ms"] = sum(out["device_ms"].values())

# host path of the product without device_init == the reference's numpy expressions, on a bounded sample
m = min(n, 2_000_000)
np.random.seed(1)
t0 = time.perf_counter()
hr, hf = np.zeros(m, np.float32), np.zeros(m, np.float32)
abm.populate_heterogeneous_values(0, m, hr, hf, pars)
t_het = time.perf_counter() - t0
t0 = time.perf_counter()
e = np.clip(pars.dur_exp(m).astype(np.int8), 0, 127)
i_ = np.clip(pars.dur_inf(m).astype(np.int8), 0, 127)
_ = np.clip(pars.t_to_paralysis(m) - e, 0, np.minimum(i_, 127)).astype(np.int8)
t_tim = time.perf_counter() - t0
t0 = time.perf_counter()
bins = core.AliasedDistribution(pyr[:, 2] + pyr[:, 3]).sample(m)
lo, hi = np.maximum(pyr[:, 0] * 365, 1), (pyr[:, 1] + 1) * 365
ages = np.random.randint(lo[bins], hi[bins]).astype(np.int32)
_ = core.KaplanMeierEstimator(cum).predict_age_at_death(ages, max_year=100)
_ = (-ages + np.random.uniform(42, 98, m)).astype(np.int32)
t_dem = time.perf_counter() - t0
t0 = time.perf_counter()
_ = np.random.choice(m, size=m // 10, replace=False)
t_mis = time.perf_counter() - t0
out["host_numpy_sample"] = {"agents": m, "seconds": {"heterogeneity": t_het, "timers": t_tim, "demography": t_dem, "missed": t_mis},
                            "agents_per_s": m / (t_het + t_tim + t_dem + t_mis)}
out["device_agents_per_s"] = n / (sum(out["device_ms"][k] for k in ("heterogeneity", "timers", "demography", "missed")) / 1e3)
out["speedup_vs_host_numpy"] = out["device_agents_per_s"] / out["host_numpy_sample"]["agents_per_s"]
out["hbm_write_GBps"] = {k: out["bytes_written_per_agent"][k] * n / (out["device_ms"][k] / 1e3) / 1e9 for k in out["bytes_written_per_agent"]}
print(json.dumps(out))
