#!/usr/bin/env python
"""bench.py -- agent-days/sec of the per-tick agent update on the shapes BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--config auto|zamfara|nigeria|west_africa|africa] [--scaling auto|shard|weak] [--agents A] [--nodes M]

A "step" is one simulated day over the whole population: every component's step() in the reference's run order
(deaths / births every 7th tick, disease state, RI every 14th, SIA on campaign days, transmission) followed by the
census, exactly what ``SEIR_ABM.run()`` does per tick.

Workload (``config.workload`` names it on every line):
  --config auto (default)  the population BASELINE.json's metric is quoted on -- Nigeria, 774 nodes, 2.2e8 agents -- at every GPU
                           count: ONE population sharded by node over the N ranks (contiguous node blocks balanced by agents),
                           i.e. strong scaling
  --config <name>          zamfara | nigeria | west_africa | africa: that population on any N
  --scaling weak           N copies of the shape's per-GPU load instead (every rank holds the config's agents and nodes,
                           network over all N x nodes): the round-1 behaviour, kept as a second line
The only per-tick exchange is the nodes x strains infectivity tally.

* value    whole-job agent-days/s with the agent table already resident in HBM (CUDA events, max over ranks)
* e2e      the same K ticks through ``SEIR_ABM`` from HOST (pinned) columns: H2D of every agent column and results array,
           the ticks, and D2H of everything the device can have changed, inside the timed region
* roofline the dominant kernel's algorithmic bytes / its mean CUDA-event time inside the timed region; ``traffic`` = DRAM
           bytes per launch measured by ncu in this round (profiles/r2_traffic.json names the capture), per day class
* verified after the timed region (outside it) the carried tallies are recomputed from scratch by the per-function
           kernels and the head-count / census identities are checked on the table the number was measured on
* named_shape  (N > 1, --config auto) the same K days on the shape BASELINE.json's configs name for this GPU count -- West Africa
           (1921 nodes, 4.3e8 agents) on 2 and 4 GPUs, Africa (5672 nodes, 1.3e9 agents) on 8 -- device-resident, verified
* cpu_baseline  the CPU oracle (C + OpenMP restatement of the reference's numba kernels, "port") timed on this box's host
           cores on a bounded sample of the same workload

--impl reference: the CPU leg alone, with every host thread, K ticks per run.
"""

from __future__ import annotations

import os
import sys

if "--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"] or "--impl=reference" in sys.argv:
    # the CPU leg uses every host core: torchrun pins OMP_NUM_THREADS=1 per rank, and OpenMP reads it when the first
    # library that links it is loaded -- so it has to be corrected before ANY other import
    if os.environ.get("OMP_NUM_THREADS") == "1" or "OMP_NUM_THREADS" not in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import argparse  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402
from pathlib import Path  # noqa: E402

import numpy as np  # noqa: E402

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

HBM_FALLBACK_GBS = 6650.0
ALGO_BYTES_PER_AGENT_TICK = 14.0  # 6 + 8 f_S + 2 f_E + 11 f_I at f_S -> 1, reference column dtypes

# BASELINE.json configs: nodes, agents, default shape per GPU count
SHAPES = {
    "zamfara": {"nodes": 14, "agents": 5_000_000},
    "nigeria": {"nodes": 774, "agents": 220_000_000},
    "west_africa": {"nodes": 1921, "agents": 430_000_000},
    "africa": {"nodes": 5672, "agents": 1_300_000_000},
}
# the shape BASELINE.json's configs name for each GPU count (reported next to the headline as "named_shape" at N > 1)
NAMED_SHAPE = {1: "nigeria", 2: "west_africa", 4: "west_africa", 8: "africa"}

# algorithmic bytes per agent per launch of each kernel, reference column dtypes, each needed column touched once
# (f_S = 0.93, f_E = f_I = 0.01 synthetic mix; derivations in DESIGN.md section 4)
KERNEL_BYTES = {
    "tick_pass": ALGO_BYTES_PER_AGENT_TICK,                # fused day: SURVEY 8(d) daily figure (pass A of t + pass B of t-1)
    "tick_node": 0.0,                                      # node-level epilogue + node math: no per-agent traffic
    "vd_births": 0.0,
    "tx_step_prep": 1 + 2 + 4 * 0.93 + 5 * 0.01,          # state, node_id, risk (S), infectivity + strain (I)
    "tx_infect": 1 + 2 + 4 * 0.93,                         # state, node_id, risk (S)
    "count_SEIRP": 1 + 2 + 1 + 1 + 0.02,                   # state, node_id, potentially_paralyzed, paralyzed, strain (E/I)
    "disease_state_step": 1 + 2 * 0.01 + 8 * 0.01,         # state; etimer r/w (E); itimer r/w, strain, ptimer r/w, pp, ipv (I)
    "get_deaths": 1 + 4,                                   # state, date_of_death
    "fast_ri": 1 + 1 + 2 + 2,                              # state, missed, ri_timer r/w
    "fast_sia": 1 + 1 + 4 + 2 * 0.3,                       # state, missed, dob, node_id (eligible quads)
}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """DRAM bytes per agent of one tick_pass launch per day class, from this round's ncu --set full capture (the file names it)."""
    p = ROOT / "profiles" / "r2_traffic.json"
    return json.loads(p.read_text()) if p.exists() else None


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled every 5 ms through NVML while the timed region runs (nvidia-smi's own loop is
    too coarse for a region of a few hundred ms; it is the fallback when NVML cannot be loaded)."""

    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.max_mhz, self._stop_evt, self.source = index, [], set(), None, threading.Event(), "nvml"

    def run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._stop_evt.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception:  # noqa: BLE001 - clocks are diagnostics
            self.source = "nvidia-smi"
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                    "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                while not self._stop_evt.is_set():
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                    if len(out) >= 6:
                        self.sm.append(float(out[0]))
                        self.max_mhz = float(out[1])
                        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:6]):
                            if v.strip().lower().startswith("active"):
                                self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=15)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": self.source}


# ------------------------------------------------------------------------------------------ workload
def plan(args, world):
    """(shape name, total nodes, total agents, mode) of this run; mode 'shard' = one population over the ranks,
    'weak' = every rank holds the shape's full per-GPU load."""
    name = "nigeria" if args.config == "auto" else args.config  # BASELINE.json metric: Nigeria 774 at 1 / 2 / 4 / 8 B200
    nodes = args.nodes or SHAPES[name]["nodes"]
    agents = args.agents or SHAPES[name]["agents"]
    mode = "shard" if args.scaling in ("auto", "shard", "strong") else "weak"
    if mode == "weak":
        return name, nodes * world, agents * world, mode
    return name, nodes, agents, mode


# ------------------------------------------------------------------------------------------ CPU leg (oracle "port")
def cpu_tick_loop(n_agents, n_nodes, ticks, warm, seed=5):
    """The reference's per-tick call sequence on the CPU oracle; returns (agent_days_per_s, threads, per-stage seconds)."""
    from laser_polio_b200 import synth

    from oracle import oracle as orc

    p = synth.synth_population(n_agents, n_nodes, seed=seed)
    n, ns = n_agents, 3
    srs = np.array([1.0, 0.25, 0.125])
    rng = np.random.default_rng(seed)
    W = rng.random((n_nodes, n_nodes)) * (0.1 / n_nodes)
    np.fill_diagonal(W, 0.0)
    r0s = rng.uniform(0.5, 1.5, n_nodes)
    pr, pi = rng.uniform(0.3, 0.8, n_nodes), rng.uniform(0.3, 0.8, n_nodes)
    vx = rng.uniform(0.4, 0.9, n_nodes).astype(np.float32)
    targeted = (rng.random(n_nodes) < 0.6).astype(np.uint8)
    pop = np.bincount(p["node_id"][:n], minlength=n_nodes).astype(np.int32)
    si, sp = np.zeros(n, np.int32), np.zeros(n, np.float32)
    rs = np.random.RandomState(seed)
    stage = {}

    def timed(name, fn):
        t0 = time.perf_counter()
        out = fn()
        stage[name] = stage.get(name, 0.0) + (time.perf_counter() - t0 if measuring else 0.0)
        return out

    t_total, measuring = 0.0, False
    first = 14 - warm  # so that the timed window contains a vital-dynamics tick and an RI tick (t = 14)
    for k in range(warm + ticks):
        t = first + k
        measuring = k >= warm
        t0 = time.perf_counter()
        if t % 7 == 0:
            dying = np.zeros(n_nodes, np.int32)
            timed("get_deaths", lambda: orc.get_deaths(n_nodes, n, p["disease_state"], p["node_id"], p["date_of_death"], t, dying))
        pot, par = np.zeros(n_nodes, np.int32), np.zeros(n_nodes, np.int32)
        timed("disease_state_step", lambda: orc.disease_state_step(
            p["node_id"], n_nodes, p["disease_state"], p["strain"], n, p["exposure_timer"], p["infection_timer"],
            p["potentially_paralyzed"], p["paralyzed"], p["ipv_protected"], p["paralysis_timer"], 1 / 2000, pot, par, seed=seed, tick=t))
        if t % 14 == 0:
            c = [np.zeros(n_nodes, np.int32) for _ in range(3)]
            timed("fast_ri", lambda: orc.fast_ri(14, p["node_id"], p["disease_state"], p["strain"], p["ipv_protected"], p["ri_timer"],
                                                 t, pr, pi, n, c[0], c[1], c[2], p["chronically_missed"], 1, seed=seed, tick=t))
        if t % 44 == 20 % 44 or k == warm + 1:  # ~8 campaigns / year; make sure one lands in a short window
            v, pr_ = np.zeros(n_nodes, np.int32), np.zeros(n_nodes, np.int32)
            timed("fast_sia", lambda: orc.fast_sia(p["node_id"], p["disease_state"], p["strain"], p["date_of_birth"], t, vx, 0.56, n,
                                                   targeted, 0, 5 * 365, v, pr_, p["chronically_missed"], 2, seed=seed, tick=t))
        beta, expo, sus, _, _ = timed("tx_step_prep", lambda: orc.tx_step_prep(
            n_nodes, n, ns, p["strain"], srs, p["disease_state"], p["node_id"], p["daily_infectivity"], p["acq_risk_multiplier"], mode="f32"))
        beta_pre, prob = timed("node_math", lambda: orc.tx_foi(beta.astype(np.float32), W, 1.05, r0s, pop))
        want, _ = timed("node_math", lambda: orc.tx_draw_counts_ref(beta_pre, prob, expo, 0.0, 1000, rs=rs))
        timed("tx_infect", lambda: orc.tx_infect_ref(n_nodes, n, ns, sus, p["node_id"], p["strain"], p["disease_state"], si, sp,
                                                      p["acq_risk_multiplier"], prob, want, seed=seed, tick=t))
        timed("count_SEIRP", lambda: orc.count_SEIRP(p["node_id"], p["disease_state"], p["strain"], p["potentially_paralyzed"],
                                                      p["paralyzed"], n_nodes, ns, n))
        if measuring:
            t_total += time.perf_counter() - t0
    return n * ticks / t_total, orc.num_threads(), stage, t_total


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    name, nodes, agents, mode = plan(args, world)
    n = min(agents, args.cpu_agents)
    cpu_nodes = min(nodes, max(14, n // 20_000))  # keep the sample's agents-per-node near the workload's
    value, threads, stage, secs = cpu_tick_loop(n, cpu_nodes, args.steps, max(1, min(args.warmup, 2)))
    sample = f"{n} agents x {cpu_nodes} nodes (bounded sample of the {agents}-agent {name} workload), {args.steps} ticks"
    line = {
        "impl": "reference", "metric": "agent-days/sec", "value": value, "unit": "agent-days/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int8 state machine + f32/f64 tallies", "data": "synthetic",
        "config": {"workload": f"{name} shape: {agents} agents, {nodes} nodes", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "agent-days/s", "cores": threads, "kind": "port", "sample": sample,
                         "stage_seconds": {k: round(v, 4) for k, v in stage.items()}},
        "e2e": {"value": value, "unit": "agent-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU leg
def day_class(t):
    """The day classes of the bench schedule (VD every 7, RI every 14, campaigns on day 20 + 44 k of every year)."""
    from laser_polio_b200 import synth

    return ("sia" if synth.campaign_day(t) else "") + ("vd" if t % 7 == 0 else "") + ("ri" if t % 14 == 0 else "") or "plain"


def measure_named_shape(name, K_, W_, local, rank, world, barrier):
    """K device-resident days of the named shape sharded over the ranks (collective: every rank calls it); rank 0 gets the
    summary dict, timed like the headline (CUDA events around the K days, max over ranks)."""
    import torch
    import torch.distributed as dist

    from laser_polio_b200 import kernels as K
    from laser_polio_b200 import synth

    n_nodes, n_agents = SHAPES[name]["nodes"], SHAPES[name]["agents"]
    sim, n_local = synth.synth_sim(n_agents, n_nodes, K_ + W_ + 40, seed=20261018, device=f"cuda:{local}", rank=rank, world=world, mode="shard")
    sim.to_device()
    sim.run_ticks(W_)
    K.STATS.reset()
    K.STATS.timing = True
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sim.run_ticks(K_)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    kstats = K.STATS.summary()
    K.STATS.timing = False
    ok = torch.tensor([1 if sim._engine.verify().get("ok") else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    sim.to_host()
    ms = float(ms.item())
    return {"workload": f"{name} shape: {n_agents} agents, {n_nodes} nodes, one population node-sharded x{world}", "value": n_agents * K_ / (ms / 1e3),
            "unit": "agent-days/s", "ms_per_step": ms / K_, "steps": K_, "agents_rank0": n_local, "verified": bool(ok.item()),
            "kernel_mean_ms_rank0": {k: round(m, 4) for k, (c, m) in kstats.items()}}


def run_b200(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if rank == 0:
        entry.build_product()
    if world > 1:
        dist.barrier()
    from laser_polio_b200 import kernels as K

    name, n_nodes, n_agents, mode = plan(args, world)
    K_, W_ = args.steps, max(args.warmup, 3)
    e2e_reps = 3 if K_ <= 400 else 1
    dur = (1 + e2e_reps) * (K_ + W_) + 40
    from laser_polio_b200 import synth

    sim, n_local = synth.synth_sim(n_agents, n_nodes, dur, seed=20261017, device=f"cuda:{local}", rank=rank, world=world, mode=mode,
                                   pars_over={"compact_every": args.compact_every} if args.compact_every else None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput
    sim.cuda_graph = bool(args.graph)
    sim.to_device()
    sim.run_ticks(W_)
    K.STATS.reset()
    K.STATS.timing = not args.graph  # per-kernel CUDA events (liblpk records them around every launch) unless graph-launched
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_first = sim.t
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sim.run_ticks(K_)  # K days: fused days are launched from C back to back (lpk_run_days), no Python per day
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    kstats = K.STATS.summary()
    if not kstats:  # graph-launched spans carry no per-kernel events: the whole step stands in for its dominant kernel
        kstats = {"tick_pass": (K_, ms / K_)}
    pass_ms = list(K.STATS.times.get("tick_pass", [])) or [a.elapsed_time(b) for a, b in K.STATS.events.get("tick_pass", [])]
    launches = K.STATS.launches
    K.STATS.timing = False
    # ---- self-check of the table the number was measured on (outside the timed region)
    eng = sim._engine
    checks = eng.verify() if eng else {"ok": False, "engine": "components only"}
    sim.to_host()
    r = sim.results
    t_last = sim.t - 1
    census = {
        "E_equals_sum_by_strain": bool(np.array_equal(r.E[:sim.t], r.E_by_strain[:sim.t].sum(axis=2))),
        "I_equals_sum_by_strain": bool(np.array_equal(r.I[:sim.t], r.I_by_strain[:sim.t].sum(axis=2))),
    }
    own = slice(sim.shard.node_lo, sim.shard.node_hi) if sim.shard is not None else slice(None)  # a rank maintains the rows of its nodes
    census["pop_bookkeeping"] = bool(np.array_equal(r.pop[t_last, own], (r.pop[0] + r.births[: sim.t].sum(axis=0) - r.deaths[: sim.t].sum(axis=0))[own]))
    st = sim.people.disease_state[: sim.people.count]
    census["census_row_equals_table"] = bool(
        int((st == 0).sum()) == int(r.S[t_last, own].sum()) and int((st == 1).sum()) == int(r.E[t_last, own].sum())
        and int((st == 2).sum()) == int(r.I[t_last, own].sum()))
    verified = torch.tensor([1 if (checks.get("ok") and all(census.values())) else 0], device="cuda")
    if world > 1:
        dist.all_reduce(verified, op=dist.ReduceOp.MIN)
    verified = bool(verified.item())

    # ---- end to end through the component API from host columns (H2D + ticks + D2H inside the timed region)
    # repeated, median reported: one pass is mostly PCIe traffic from a shared host, and single shots came out bimodal
    # on this pool; every repetition is listed in the JSON line
    e2e_runs = []
    for _ in range(e2e_reps):
        barrier()
        t0 = time.perf_counter()
        sim.to_device()
        sim.run_ticks(K_)
        sim.to_host()
        barrier()
        rep_s = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(rep_s, op=dist.ReduceOp.MAX)
        e2e_runs.append(float(rep_s.item()))
    e2e_s = float(np.median(e2e_runs))
    h2d, d2h = sim.io_bytes
    live = torch.tensor([n_local], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(live)
    agents_total = int(live.item())

    # ---- the shape BASELINE.json's configs name for this GPU count, as a second measurement in the same line (N > 1 only:
    # the headline at every N is the Nigeria population sharded N ways; this one keeps the per-GPU load near one GPU's)
    named = None
    if world > 1 and args.config == "auto" and mode == "shard" and not args.no_named_shape:
        del sim
        import gc

        gc.collect()
        torch.cuda.empty_cache()
        named = measure_named_shape(NAMED_SHAPE.get(world, "nigeria"), K_, W_, local, rank, world, barrier)

    if rank != 0:
        return
    value = agents_total * K_ / (ms / 1e3)
    peak, peak_src = measured_peak()
    top = max(kstats, key=lambda k: kstats[k][0] * kstats[k][1])
    calls, mean_ms = kstats[top]
    algo = KERNEL_BYTES[top] * n_local
    achieved = algo / (mean_ms / 1e3) / 1e9
    kernel_share = {k: round(c * m / ms, 4) for k, (c, m) in kstats.items()}
    by_class = {}
    for k, v in enumerate(pass_ms):
        by_class.setdefault(day_class(t_first + k), []).append(v)
    traffic = measured_traffic()
    per_class = {}
    for cls, v in sorted(by_class.items()):
        m = float(np.mean(v))
        per_class[cls] = {"launches": len(v), "mean_ms": round(m, 4), "frac_of_14B_roofline": round(ALGO_BYTES_PER_AGENT_TICK * n_local / (m / 1e3) / 1e9 / peak, 4)}
        if traffic and cls in traffic.get("bytes_per_agent", {}):
            moved = traffic["bytes_per_agent"][cls] * n_local
            per_class[cls]["traffic"] = moved
            per_class[cls]["moved_frac_of_peak"] = round(moved / (m / 1e3) / 1e9 / peak, 4)
    plain_traffic = traffic["bytes_per_agent"].get("plain") if traffic else None
    cpu_base = None
    if world == 1:  # the CPU leg is reported at N = 1 only (rank 0's host cores)
        cpu_n = min(n_agents, args.cpu_agents)
        cpu_nodes = min(n_nodes, max(14, cpu_n // 20_000))
        cpu_value, threads, stage, cpu_secs = cpu_tick_loop(cpu_n, cpu_nodes, args.cpu_ticks, 1)
        cpu_base = {"value": cpu_value, "unit": "agent-days/s", "cores": threads, "kind": "port",
                    "sample": f"{cpu_n} agents x {cpu_nodes} nodes, {args.cpu_ticks} ticks ({cpu_secs:.1f} s)",
                    "stage_seconds": {k: round(v, 4) for k, v in stage.items()}}
    line = {
        "metric": "agent-days/sec", "value": value, "unit": "agent-days/s", "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": ms / K_, "higher_is_better": True, "scaling": "weak" if mode == "weak" else "strong",
        "vs_baseline": None, "dtype": "u8 agenda bytes + int8 state machine + int64 fixed-point tallies + f64 node math", "data": "synthetic",
        "config": {"workload": f"{name} shape (BASELINE.json configs): {agents_total} agents, {n_nodes} nodes over {world} GPU(s), 3 strains, "
                               "VD every 7, RI every 14, ~8 SIA/yr, daily transmission + census",
                   "shape": name, "agents_total": agents_total, "agents_rank0": n_local, "nodes": n_nodes,
                   "l2_policy": "per-tick working set (>= 0.2 GB of agenda bytes + event records) far larger than the 126 MB L2",
                   "parallelism": (f"{'one population node-sharded' if mode == 'shard' else 'per-GPU replicas of the shape, node-sharded'} x{world}, "
                                   "one exchange of the nodes x strains tally per tick" if world > 1 else "single GPU")},
        "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (plain_traffic * n_local) if (plain_traffic and top == "tick_pass") else None,
                     "traffic_source": traffic.get("source") if traffic else None,
                     "note": "achieved = 14 B/agent-tick (SURVEY 8d, reference dtypes) x agents / mean launch time, as the contract defines it; the pass "
                             "itself moves ~3 B/agent (agenda byte + event records), so frac > what the DRAM counters show: see per_day_class.moved_frac_of_peak",
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": algo, "mean_ms": mean_ms, "launches": calls,
                     "per_day_class": per_class, "kernel_share_of_step": kernel_share},
        "verified": verified, "verified_checks": {**checks, **census},
        "compact_every": args.compact_every, "compactions": int(getattr(sim, "_compactions", 0)) if world == 1 or args.config != "auto" else None,
        "named_shape": named,
        "cpu_baseline": cpu_base,
        "e2e": {"value": agents_total * K_ / e2e_s, "unit": "agent-days/s", "h2d_bytes_per_step": h2d / K_, "d2h_bytes_per_step": d2h / K_,
                "seconds": e2e_s, "seconds_each": [round(x, 4) for x in e2e_runs],
                "note": f"SEIR_ABM.to_device() + K step_tick() + to_host() from pinned host columns, median of {e2e_reps} repetition(s); "
                        "byte counts are rank 0's"},
        "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=140)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="auto", choices=["auto"] + list(SHAPES))
    ap.add_argument("--scaling", default="auto", choices=["auto", "shard", "strong", "weak"])
    ap.add_argument("--agents", type=int, default=0, help="total agents (default: the shape's)")
    ap.add_argument("--nodes", type=int, default=0, help="total nodes (default: the shape's)")
    ap.add_argument("--graph", type=int, default=0, help="1: lpk_run_days captures each span of days into a CUDA graph (no per-kernel timing)")
    ap.add_argument("--compact-every", type=int, default=0, help="pars.compact_every: compact the device table every that many ticks (0 = never)")
    ap.add_argument("--no-named-shape", action="store_true", help="N > 1: skip the second measurement on the shape named for this GPU count")
    ap.add_argument("--cpu-agents", type=int, default=20_000_000)
    ap.add_argument("--cpu-ticks", type=int, default=60)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
