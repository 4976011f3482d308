"""Node-sharded run == single-table run, bit for bit.

Two ranks share cuda:0 (the tally all-reduce goes over gloo, which is what a one-GPU test box allows; on the 8-GPU box
the same code path runs over NCCL).  The reference is ONE process holding one table laid out as
[shard 0's table | shard 1's table] so that every agent has the same global id as in the sharded run."""

import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("fused", [True, False])
def test_two_shards_equal_one_table(tmp_path, fused):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    two_shards_vs_one_table(tmp_path, fused, "gloo")


def test_two_shards_on_two_gpus_peer_memory_exchange(tmp_path):
    """One rank per GPU over NCCL: on fused days the tally goes through liblpk's peer-memory exchange (CUDA IPC + NVLink
    stores + system-scope flags) instead of an all-reduce call.  Needs two GPUs (gpurun --gpus 2); skipped on one."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    two_shards_vs_one_table(tmp_path, True, "nccl")


def two_shards_vs_one_table(tmp_path, fused, backend):
    import torch.multiprocessing as mp

    import laser_polio_b200 as lp
    from sharded_worker import RESULT_KEYS, gpu_rank, make_sim, pyramid_file

    pyr = pyramid_file(tmp_path / "pyramid.csv")
    mp.spawn(gpu_rank, args=(2, free_port(), str(tmp_path), fused, backend), nprocs=2, join=True)
    ranks = [dict(np.load(tmp_path / f"rank{r}_{int(fused)}.npz")) for r in range(2)]

    # the same population as one table with the shards' global ids
    tables = []
    for r in range(2):
        np.random.seed(0)
        s = make_sim(lp, pyr)
        s.shard_to(r, 2)
        tables.append((s.people, s.id_base))
    np.random.seed(0)
    whole = make_sim(lp, pyr)
    (p0, b0), (p1, b1) = tables
    assert b0 == 0 and b1 >= p0.capacity and b1 % 4 == 0
    frame = lp.LaserFrame(capacity=b1 + p1.capacity, initial_count=b1 + p1.count)
    from laser_polio_b200.abm import _UNBORN_DEFAULTS

    for name, col in p0.columns().items():
        frame.add_scalar_property(name, dtype=col.dtype, default=_UNBORN_DEFAULTS.get(name, 0))
        new = getattr(frame, name)
        new[: p0.capacity] = col
        new[b1 : b1 + p1.capacity] = getattr(p1, name)
    whole.people = frame
    for inst in whole.instances:
        inst.people = frame
    whole.fused = fused
    whole.run()

    n_nodes = len(whole.nodes)
    owned0 = np.zeros(n_nodes, bool)
    owned0[int(ranks[0]["node_lo"]) : int(ranks[0]["node_hi"])] = True
    for key in RESULT_KEYS:
        ref = getattr(whole.results, key)
        mask = owned0.reshape((1, n_nodes) + (1,) * (ref.ndim - 2))
        combined = np.where(mask, ranks[0][key], ranks[1][key])
        assert np.array_equal(ref, combined), key
    assert whole.results.new_exposed.sum() > 200 and whole.results.sia_protected.sum() > 0 and whole.results.deaths.sum() > 0
    assert whole.results.I[15, 7] >= 30  # the seed_schedule injection landed on the rank that owns node 7
    n0, n1 = int(ranks[0]["count"]), int(ranks[1]["count"])
    assert np.array_equal(whole.people.disease_state[:n0], ranks[0]["disease_state"])
    assert np.array_equal(whole.people.disease_state[b1 : b1 + n1], ranks[1]["disease_state"])
    assert np.array_equal(whole.people.strain[b1 : b1 + n1], ranks[1]["strain"])
