"""Node sharding across the GPUs of one box.

Agents never change node (node_id is written at construction and at birth only; "migration" moves infectivity
through the network matrix, reference model.py:1328-1338), so every per-agent kernel and every per-node output is
node-local.  Each rank owns a contiguous block of node ids and the agents living there; node ids stay GLOBAL inside
every rank's table, so all per-node arrays have the global length and a rank simply never touches rows it does not own.

The only per-tick exchange is the sum of the [nodes, strains] infectivity tally before the network transfer
(``beta += W^T beta - beta * rowsum(W)`` needs every node's infectivity): one all-reduce of an int64 fixed-point
array -- 774 x 3 x 8 B = 18.6 KB at Nigeria scale, 136 KB for the continental config -- which is exact and
order-independent, so a sharded run is bit-identical to the single-GPU run of the same population
(tests/test_gpu_sharded.py) as long as agent ids are global: ``sim.id_base`` carries the global id of the
shard's first agent into every Philox counter.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Shard:
    rank: int
    world: int
    node_lo: int  # first owned node
    node_hi: int  # one past the last owned node
    group: object = None  # torch.distributed process group (None = default)

    def owns(self, node: int) -> bool:
        return self.node_lo <= node < self.node_hi

    def owned_mask(self, n_nodes: int) -> np.ndarray:
        m = np.zeros(n_nodes, dtype=bool)
        m[self.node_lo : self.node_hi] = True
        return m


def plan_node_blocks(node_agents, world: int):
    """Contiguous node blocks with (nearly) equal agent counts: cut where the running sum crosses k/world of the total.

    Returns [(lo, hi)] * world; every block is non-empty as long as there are at least `world` nodes."""
    sizes = np.asarray(node_agents, dtype=np.int64)
    n = len(sizes)
    if world < 1 or n < world:
        raise ValueError(f"cannot split {n} nodes over {world} ranks")
    cum = np.cumsum(sizes)
    total = cum[-1]
    cuts = [0]
    for k in range(1, world):
        target = total * k / world
        c = int(np.searchsorted(cum, target, side="left")) + 1  # first boundary whose prefix reaches the target
        c = max(c, cuts[-1] + 1)          # at least one node per block
        c = min(c, n - (world - k))       # leave a node for every later block
        cuts.append(c)
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def id_bases(node_agents, blocks, capacity_per_block=None):
    """Global id of each shard's first agent: the running sum of shard sizes rounded up to a multiple of 256
    (exposure draws are made per aligned group of 256 agents: one Philox block per lane and pair of 128-agent rows)."""
    sizes = np.asarray(node_agents, dtype=np.int64)
    bases, run = [], 0
    for k, (lo, hi) in enumerate(blocks):
        bases.append(run)
        span = int(sizes[lo:hi].sum()) if capacity_per_block is None else int(capacity_per_block[k])
        run += (span + 255) // 256 * 256
    return bases


def allreduce_tally(beta_fx, shard: Shard | None):
    """Sum the fixed-point infectivity tally over ranks, in place (NCCL on GPUs; gloo in the CPU tests)."""
    if shard is None or shard.world == 1:
        return beta_fx
    import torch.distributed as dist

    dist.all_reduce(beta_fx, op=dist.ReduceOp.SUM, group=shard.group)
    return beta_fx
