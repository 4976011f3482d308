"""The vectorised pieces of oracle/tick_loop.py against the scalar restatements they replace."""

import numpy as np


def test_vectorised_philox_matches_the_block_function(oracle):
    from oracle import tick_loop as tl

    rs = np.random.RandomState(0)
    ctr = rs.randint(0, 2**32, (50, 4), dtype=np.uint64).astype(np.uint32)
    k0, k1 = 0xDEADBEEF, 0x12345678
    got = np.stack(tl.philox_np(ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3], k0, k1), axis=1)
    for i in range(len(ctr)):
        assert np.array_equal(got[i], oracle.philox4x32_10([int(v) for v in ctr[i]], [k0, k1]))


def test_vectorised_births_match_the_scalar_restatement(oracle):
    from laser_polio_b200 import utils
    from oracle import tick_loop as tl

    rs = np.random.default_rng(4)
    n_nodes, count, cap, tick, seed, id_base = 23, 5_003, 40_000, 21, 0xABCDEF0123, 4096
    pop_prev = rs.integers(2_000, 400_000, n_nodes).astype(np.int32)
    pop_prev[5] = 0
    rate = rs.uniform(20, 45, n_nodes) / 365000.0
    cd = np.insert(utils.create_cumulative_deaths(int(pop_prev.sum()), 100).astype(np.int64), 0, 0)
    a = oracle.vd_births_device(pop_prev, rate, 7, cd, count, cap, seed, tick, id_base=id_base)
    b = tl.births(pop_prev, rate, 7, cd, count, cap, seed, tick, id_base=id_base)
    assert a[3] == b[3] and a[4] == b[4] == 0 and a[0].sum() > 300
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
    assert tl.births(pop_prev, rate, 7, cd, count, count + int(a[0].sum()) - 1, seed, tick, id_base=id_base)[4] == 1
