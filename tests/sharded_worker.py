"""Worker processes for the world_size = 2 sharding tests (CPU gloo test and single-GPU gloo test)."""

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pyramid_file(path):
    rows = ["Age,M,F"] + [f"{5 * k}-{5 * k + 4},{int(1.7e7 * np.exp(-0.16 * k))},{int(1.6e7 * np.exp(-0.16 * k))}" for k in range(20)]
    rows.append("100+,300,500")
    Path(path).write_text("\n".join(rows) + "\n")
    return str(path)


def make_sim(lp, pyramid, dur=40):
    n_nodes = 9
    rs = np.random.RandomState(11)
    d = rs.uniform(5, 300, (n_nodes, n_nodes))
    d = (d + d.T) / 2
    np.fill_diagonal(d, 0)
    pars = lp.PropertySet({
        "start_date": lp.date("2019-01-01"), "dur": dur, "init_pop": rs.randint(2000, 12000, n_nodes), "cbr": np.zeros(n_nodes),
        "r0_scalars": rs.uniform(0.5, 1.5, n_nodes), "age_pyramid_path": pyramid, "init_immun": 0.3,
        "init_prev": [0.02] + [0.0] * (n_nodes - 1), "r0": 14, "distances": d, "stop_if_no_cases": False, "verbose": 0, "seed": 5,
        "vx_prob_ri": rs.uniform(0.3, 0.9, n_nodes), "vx_prob_ipv": rs.uniform(0.3, 0.9, n_nodes), "missed_frac": 0.1, "p_paralysis": 0.3,
        "node_seeding_zero_inflation": 0.3, "node_seeding_dispersion": 2, "max_migr_frac": 0.3,
        "sia_schedule": [{"date": "2019-01-12", "nodes": [0, 2, 5, 8], "age_range": (0, 5 * 365), "vaccinetype": "nOPV2"}],
        "vx_prob_sia": rs.uniform(0.4, 0.9, n_nodes).tolist(),
        "seed_schedule": [{"timestep": 15, "node_id": 7, "prevalence": 30}],
    })
    sim = lp.SEIR_ABM(pars)
    sim.components = [lp.VitalDynamics_ABM, lp.DiseaseState_ABM, lp.RI_ABM, lp.SIA_ABM, lp.Transmission_ABM]
    sim.people.ri_timer[: sim.people.count : 3] = rs.randint(-10, 40, len(sim.people.ri_timer[: sim.people.count : 3]))
    return sim


RESULT_KEYS = ("S", "E", "I", "R", "E_by_strain", "I_by_strain", "new_exposed", "new_exposed_by_strain", "potentially_paralyzed",
               "paralyzed", "new_potentially_paralyzed", "new_paralyzed", "deaths", "ri_vaccinated", "ri_protected", "ipv_vaccinated",
               "sia_vaccinated", "sia_protected", "sia_new_exposed_by_strain")


def gpu_rank(rank, world, port, workdir, fused, backend="gloo"):
    """One rank of a sharded run.  gloo: both ranks share cuda:0 and the tally all-reduce goes through torch.distributed;
    nccl: one GPU per rank, the tally travels through the peer-memory exchange of liblpk (lpk_xchg) on fused days."""
    import torch
    import torch.distributed as dist

    import laser_polio_b200 as lp

    torch.cuda.set_device(rank if backend == "nccl" else 0)
    kw = {"device_id": torch.device(f"cuda:{rank}")} if backend == "nccl" else {}
    dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, **kw)
    np.random.seed(0)
    sim = make_sim(lp, os.path.join(workdir, "pyramid.csv"))
    sim.fused = fused
    shard = sim.shard_to(rank, world)
    sim.to_device()
    sim.step_tick(0)
    sim.step_tick(1)  # creates the engine: is the tally exchange the one this backend should get?
    if fused:
        assert (sim._engine.xchg is not None) == (backend == "nccl"), "peer-memory exchange"
    sim.run()
    out = {k: getattr(sim.results, k) for k in RESULT_KEYS}
    out["node_lo"], out["node_hi"], out["id_base"], out["count"] = shard.node_lo, shard.node_hi, sim.id_base, sim.people.count
    out["disease_state"] = sim.people.disease_state[: sim.people.count]
    out["strain"] = sim.people.strain[: sim.people.count]
    np.savez(os.path.join(workdir, f"rank{rank}_{int(fused)}.npz"), **out)
    dist.destroy_process_group()


def cpu_rank(rank, world, port, workdir):
    """CPU-only: the sharded fixed-point tally summed over gloo equals the tally of the whole population."""
    import torch
    import torch.distributed as dist

    import laser_polio_b200.synth as synth
    from laser_polio_b200 import sharding
    from oracle import oracle as orc

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n_nodes, n = 23, 60_000
    p = synth.synth_population(n, n_nodes, seed=2, f_infected=0.1)
    blocks = sharding.plan_node_blocks(p["node_sizes"], world)
    lo, hi = blocks[rank]
    starts = np.concatenate([[0], np.cumsum(p["node_sizes"])])
    a0, a1 = int(starts[lo]), int(starts[hi])
    srs = np.array([1.0, 0.25, 0.125])
    _, _, _, bfx, _ = orc.tx_step_prep(n_nodes, a1 - a0, 3, p["strain"][a0:a1].copy(), srs, p["disease_state"][a0:a1].copy(),
                                       p["node_id"][a0:a1].copy(), p["daily_infectivity"][a0:a1].copy(),
                                       p["acq_risk_multiplier"][a0:a1].copy(), mode="fx")
    assert np.all(bfx[:lo] == 0) and np.all(bfx[hi:] == 0)  # a rank only ever touches the rows of its own nodes
    t = torch.from_numpy(bfx.copy())
    sharding.allreduce_tally(t, sharding.Shard(rank, world, lo, hi))
    _, _, _, whole, _ = orc.tx_step_prep(n_nodes, n, 3, p["strain"], srs, p["disease_state"], p["node_id"], p["daily_infectivity"],
                                         p["acq_risk_multiplier"], mode="fx")
    ok = bool(np.array_equal(t.numpy(), whole))
    Path(workdir, f"cpu_rank{rank}.txt").write_text("ok" if ok else "MISMATCH")
    dist.destroy_process_group()
