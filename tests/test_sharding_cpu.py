"""Node sharding, host side (no GPU): block planning, per-shard agent tables, and the world_size = 2 gloo all-reduce of
the fixed-point infectivity tally."""

import socket
from pathlib import Path

import numpy as np
import pytest


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_plan_node_blocks_balances_agents():
    from laser_polio_b200 import sharding

    rng = np.random.default_rng(0)
    sizes = rng.lognormal(0, 1, 774) * 1000 + 1
    for world in (1, 2, 4, 8):
        blocks = sharding.plan_node_blocks(sizes, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == 774 and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        per = np.array([sizes[lo:hi].sum() for lo, hi in blocks])
        assert per.min() > 0 and per.max() / per.mean() < 1.15
        bases = sharding.id_bases(sizes.astype(int), blocks)
        assert all(b % 4 == 0 for b in bases) and all(x < y for x, y in zip(bases, bases[1:]))
    assert sharding.plan_node_blocks([5, 5, 5], 3) == [(0, 1), (1, 2), (2, 3)]
    assert sharding.plan_node_blocks([1000, 1, 1, 1], 4) == [(0, 1), (1, 2), (2, 3), (3, 4)]
    with pytest.raises(ValueError):
        sharding.plan_node_blocks([1, 2], 3)


def test_shard_to_partitions_the_agent_table(tmp_path):
    import laser_polio_b200 as lp
    from sharded_worker import make_sim, pyramid_file

    pyr = pyramid_file(tmp_path / "pyramid.csv")
    np.random.seed(0)
    whole = make_sim(lp, pyr)
    n_nodes, count = len(whole.nodes), whole.people.count
    seen, id_ranges = 0, []
    for rank in range(3):
        np.random.seed(0)
        sim = make_sim(lp, pyr)
        shard = sim.shard_to(rank, 3)
        p = sim.people
        own = (whole.people.node_id[:count] >= shard.node_lo) & (whole.people.node_id[:count] < shard.node_hi)
        assert p.count == own.sum() and p.capacity > p.count
        for name, col in whole.people.columns().items():
            assert np.array_equal(getattr(p, name)[: p.count], col[:count][own]), name
        assert np.all(p.disease_state[p.count:] == -1) and np.all(p.node_id[p.count:] == -1)
        assert all(inst.people is p for inst in sim.instances)
        assert sim.id_base % 256 == 0
        id_ranges.append((sim.id_base, sim.id_base + p.capacity))
        seen += p.count
        assert np.array_equal(sim.results.pop[0], whole.results.pop[0])  # per-node arrays keep the global length
    assert seen == count
    assert all(a[1] <= b[0] for a, b in zip(id_ranges, id_ranges[1:]))  # disjoint Philox id ranges


def test_world2_gloo_allreduce_of_tally(tmp_path):
    import torch.multiprocessing as mp
    from sharded_worker import cpu_rank

    port = free_port()
    mp.spawn(cpu_rank, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        assert Path(tmp_path, f"cpu_rank{rank}.txt").read_text() == "ok"
