"""Parity of the CUDA kernels (through the C ABI, liblpk.so) against
(a) the golden vectors produced by the REFERENCE's numba kernels with injected uniforms, and
(b) the CPU oracle in Philox mode on larger seeded populations, including ragged / unsorted / empty inputs.

Bar: bit-exact for every integer column and per-node count; float tallies are exact fixed point
(bit-exact against the oracle's fixed-point flavour, 1e-6 relative against its float64 sum -- gate 2);
node-level float64 math to 1e-6 relative (+1e-15 absolute for probabilities below 1e-9, where
1 - exp(-x) itself carries ~1e-16 absolute rounding in both implementations).
"""

import numpy as np
import pytest
from conftest import golden_inputs, load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from laser_polio_b200 import kernels

    return kernels


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def to_dev(p):
    return {k: dev(v) for k, v in p.items() if isinstance(v, np.ndarray) and v.ndim == 1 and k != "node_sizes"}


def host(t):
    return t.cpu().numpy()


KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_known_answers_on_device(K, oracle):
    ctr = np.array([c for c, _, _ in KAT], np.uint32)
    key = np.array([k for _, k, _ in KAT], np.uint32)
    out = host(K.philox_selftest(dev(ctr.view(np.int32)), dev(key.view(np.int32))))
    assert np.array_equal(out.view(np.uint32), np.array([w for _, _, w in KAT], np.uint32))
    rng = np.random.default_rng(1)
    ctr = rng.integers(0, 2**32, (4096, 4), dtype=np.uint32)
    key = rng.integers(0, 2**32, (4096, 2), dtype=np.uint32)
    out = host(K.philox_selftest(dev(ctr.view(np.int32)), dev(key.view(np.int32)))).view(np.uint32)
    for i in range(0, 4096, 97):
        assert np.array_equal(out[i], oracle.philox4x32_10(ctr[i], key[i]))


# ----------------------------------------------------------------- (a) against the reference's golden vectors
def test_disease_state_step_vs_reference_golden(K):
    for name in ("ds_p03", "ds_p2000"):
        g = load_golden(name)
        d = to_dev(golden_inputs(g))
        n_nodes, count = int(g["n_nodes"]), int(g["count"])
        for t in range(g["u"].shape[0]):
            pot = torch.zeros(n_nodes, dtype=torch.int32, device="cuda")
            par = torch.zeros(n_nodes, dtype=torch.int32, device="cuda")
            u = dev(g["u"][t])
            K.disease_state_step(d["node_id"], n_nodes, d["disease_state"], d["strain"], count, d["exposure_timer"],
                                 d["infection_timer"], d["potentially_paralyzed"], d["paralyzed"], d["ipv_protected"],
                                 d["paralysis_timer"], float(g["p_paralysis"]), pot, par, rng=K.make_rng(u1=u))
            assert np.array_equal(host(pot), g["new_potential"][t]), (name, t)
            assert np.array_equal(host(par), g["new_paralyzed"][t]), (name, t)
        for k in ("disease_state", "exposure_timer", "infection_timer", "paralysis_timer", "potentially_paralyzed",
                  "paralyzed"):
            assert np.array_equal(host(d[k]), g[f"out_{k}"]), (name, k)


def test_get_deaths_vs_reference_golden(K):
    g = load_golden("deaths")
    d = to_dev(golden_inputs(g))
    dying = torch.full((int(g["n_nodes"]),), 7, dtype=torch.int32, device="cuda")  # must be overwritten
    K.get_deaths(int(g["n_nodes"]), int(g["count"]), d["disease_state"], d["node_id"], d["date_of_death"], int(g["t"]), dying)
    assert np.array_equal(host(dying), g["num_dying"])
    assert np.array_equal(host(d["disease_state"]), g["out_disease_state"])


def test_fast_ri_vs_reference_golden(K):
    for name in ("ri_t14", "ri_t28"):
        g = load_golden(name)
        d = to_dev(golden_inputs(g))
        n_nodes = int(g["n_nodes"])
        c = [torch.full((n_nodes,), 3, dtype=torch.int32, device="cuda") for _ in range(3)]
        K.fast_ri(int(g["step_size"]), d["node_id"], d["disease_state"], d["strain"], d["ipv_protected"], d["ri_timer"],
                  int(g["sim_t"]), dev(g["vx_prob_ri"]), dev(g["vx_prob_ipv"]), int(g["count"]), c[0], c[1], c[2],
                  d["chronically_missed"], int(g["vaccine_strain"]), rng=K.make_rng(u1=dev(g["u1"]), u2=dev(g["u2"])))
        assert np.array_equal(host(c[0]), g["ri_counts"]), name
        assert np.array_equal(host(c[1]), g["ri_protected"]), name
        assert np.array_equal(host(c[2]), g["ipv_counts"]), name
        for k in ("disease_state", "strain", "ipv_protected", "ri_timer"):
            assert np.array_equal(host(d[k]), g[f"out_{k}"]), (name, k)


def test_fast_sia_vs_reference_golden(K):
    g = load_golden("sia")
    d = to_dev(golden_inputs(g))
    n_nodes = int(g["n_nodes"])
    v = torch.full((n_nodes,), 9, dtype=torch.int32, device="cuda")
    pr = torch.full((n_nodes,), 9, dtype=torch.int32, device="cuda")
    K.fast_sia(d["node_id"], d["disease_state"], d["strain"], d["date_of_birth"], int(g["sim_t"]), dev(g["vx_prob"]),
               float(g["vx_eff"]), int(g["count"]), dev(g["nodes_to_vaccinate"]), int(g["min_age"]), int(g["max_age"]), v,
               pr, d["chronically_missed"], int(g["vaccine_strain"]), rng=K.make_rng(u1=dev(g["u"])))
    assert np.array_equal(host(v), g["vaccinated"])
    assert np.array_equal(host(pr), g["protected"])
    for k in ("disease_state", "strain"):
        assert np.array_equal(host(d[k]), g[f"out_{k}"]), k


def test_tally_and_census_vs_reference_golden(K, oracle):
    g = load_golden("tally_census")
    p = golden_inputs(g)
    d = to_dev(p)
    n, n_nodes, n_strains = int(g["count"]), int(g["n_nodes"]), int(g["n_strains"])
    beta_fx, expo_fx, sus, hist = K.tx_step_prep(n_nodes, n, n_strains, d["strain"], g["strain_r0_scalars"], d["disease_state"],
                                                 d["node_id"], d["daily_infectivity"], d["acq_risk_multiplier"])
    assert np.array_equal(host(hist).sum(axis=1), g["sus"])  # one histogram entry per susceptible
    assert np.array_equal(host(sus), g["sus"])
    # gate 2: node tallies within 1e-6 relative (reference float32 sums are themselves only ~5e-6 accurate -> 2e-5)
    np.testing.assert_allclose(host(beta_fx) / 2.0**30, g["beta"], rtol=2e-5)
    np.testing.assert_allclose(host(expo_fx) / 2.0**30, g["exposure"], rtol=2e-5)
    b64, e64, _, bi, ei = oracle.tx_step_prep(n_nodes, n, n_strains, p["strain"], g["strain_r0_scalars"], p["disease_state"],
                                              p["node_id"], p["daily_infectivity"], p["acq_risk_multiplier"], mode="fx")
    assert np.array_equal(host(beta_fx), bi) and np.array_equal(host(expo_fx), ei)  # fixed point: bit-exact
    assert np.array_equal(host(hist), oracle.tx_step_prep.last_hist)
    t64 = oracle.tx_step_prep(n_nodes, n, n_strains, p["strain"], g["strain_r0_scalars"], p["disease_state"], p["node_id"],
                              p["daily_infectivity"], p["acq_risk_multiplier"], mode="f64")
    np.testing.assert_allclose(host(beta_fx) / 2.0**30, t64[0], rtol=1e-6)
    np.testing.assert_allclose(host(expo_fx) / 2.0**30, t64[1], rtol=1e-6)

    out = K.count_SEIRP(d["node_id"], d["disease_state"], d["strain"], d["potentially_paralyzed"], d["paralyzed"], n_nodes,
                        n_strains, n)
    for got, key in zip(out, ("S", "E", "I", "R", "E_by_strain", "I_by_strain", "POTP", "P")):
        assert np.array_equal(host(got), g[key]), key


# ----------------------------------------------------------------- (b) against the oracle, Philox mode, larger inputs
CASES = [  # (n_agents, capacity, n_nodes, sorted)
    (1_000_003, 1_000_448, 37, True),   # ragged: not a multiple of 4 or 512
    (300_001, 300_032, 774, False),     # node ids shuffled (births appended out of order in the reference)
    (513, 1024, 3, True),
    (3, 16, 2, True),
]


def population(n, cap, nodes, srt, seed):
    import laser_polio_b200.synth as synth

    return synth.synth_population(n, nodes, seed=seed, capacity=cap, f_exposed=0.05, f_infected=0.06, f_recovered=0.1,
                                  f_dead=0.04, sorted_nodes=srt)


@pytest.mark.parametrize("n,cap,nodes,srt", CASES)
def test_all_stages_vs_oracle_philox(K, oracle, n, cap, nodes, srt):
    p = population(n, cap, nodes, srt, seed=n % 1000)
    p["date_of_death"][:n] = np.random.default_rng(5).integers(-3, 60, n).astype(np.int32)
    p["ri_timer"][:n] = np.random.default_rng(6).integers(-20, 40, n).astype(np.int16)
    d = to_dev(p)
    seed = 0xC0FFEE1234567
    ns = 3
    srs = np.array([1.0, 0.3, 0.125])
    for tick in (7, 14, 28):
        # V1
        dying_o = np.zeros(nodes, np.int32)
        oracle.get_deaths(nodes, n, p["disease_state"], p["node_id"], p["date_of_death"], tick, dying_o)
        dying = torch.empty(nodes, dtype=torch.int32, device="cuda")
        K.get_deaths(nodes, n, d["disease_state"], d["node_id"], d["date_of_death"], tick, dying)
        assert np.array_equal(host(dying), dying_o)
        # D1
        pot_o, par_o = np.zeros(nodes, np.int32), np.zeros(nodes, np.int32)
        oracle.disease_state_step(p["node_id"], nodes, p["disease_state"], p["strain"], n, p["exposure_timer"],
                                  p["infection_timer"], p["potentially_paralyzed"], p["paralyzed"], p["ipv_protected"],
                                  p["paralysis_timer"], 0.25, pot_o, par_o, seed=seed, tick=tick)
        pot = torch.zeros(nodes, dtype=torch.int32, device="cuda")
        par = torch.zeros(nodes, dtype=torch.int32, device="cuda")
        K.disease_state_step(d["node_id"], nodes, d["disease_state"], d["strain"], n, d["exposure_timer"],
                             d["infection_timer"], d["potentially_paralyzed"], d["paralyzed"], d["ipv_protected"],
                             d["paralysis_timer"], 0.25, pot, par, rng=K.make_rng(seed, tick))
        assert np.array_equal(host(pot), pot_o) and np.array_equal(host(par), par_o)
        # R1
        pr = np.linspace(0.2, 0.9, nodes)
        pi = np.linspace(0.8, 0.1, nodes)
        co = [np.zeros(nodes, np.int32) for _ in range(3)]
        oracle.fast_ri(14, p["node_id"], p["disease_state"], p["strain"], p["ipv_protected"], p["ri_timer"], tick, pr, pi, n,
                       co[0], co[1], co[2], p["chronically_missed"], 1, seed=seed, tick=tick)
        c = [torch.empty(nodes, dtype=torch.int32, device="cuda") for _ in range(3)]
        K.fast_ri(14, d["node_id"], d["disease_state"], d["strain"], d["ipv_protected"], d["ri_timer"], tick, dev(pr), dev(pi),
                  n, c[0], c[1], c[2], d["chronically_missed"], 1, rng=K.make_rng(seed, tick))
        for a, b in zip(c, co):
            assert np.array_equal(host(a), b)
        # S1, two campaigns the same day
        for ev, (vs, eff) in enumerate(((2, 0.56), (1, 0.7))):
            vx = np.linspace(0.3, 0.95, nodes).astype(np.float32)
            tg = (np.arange(nodes) % 3 != ev).astype(np.uint8)
            vo, po = np.zeros(nodes, np.int32), np.zeros(nodes, np.int32)
            oracle.fast_sia(p["node_id"], p["disease_state"], p["strain"], p["date_of_birth"], tick, vx, eff, n, tg, 0, 5 * 365,
                            vo, po, p["chronically_missed"], vs, seed=seed, tick=tick, event_idx=ev)
            v = torch.empty(nodes, dtype=torch.int32, device="cuda")
            pr_ = torch.empty(nodes, dtype=torch.int32, device="cuda")
            K.fast_sia(d["node_id"], d["disease_state"], d["strain"], d["date_of_birth"], tick, dev(vx), eff, n, dev(tg), 0,
                       5 * 365, v, pr_, d["chronically_missed"], vs, event_idx=ev, rng=K.make_rng(seed, tick))
            assert np.array_equal(host(v), vo) and np.array_equal(host(pr_), po)
        # T1
        _, _, sus_o, bfx_o, efx_o = oracle.tx_step_prep(nodes, n, ns, p["strain"], srs, p["disease_state"], p["node_id"],
                                                        p["daily_infectivity"], p["acq_risk_multiplier"], mode="fx")
        hist_o = oracle.tx_step_prep.last_hist
        bfx, efx, sus, hist = K.tx_step_prep(nodes, n, ns, d["strain"], srs, d["disease_state"], d["node_id"],
                                             d["daily_infectivity"], d["acq_risk_multiplier"])
        assert np.array_equal(host(bfx), bfx_o) and np.array_equal(host(efx), efx_o) and np.array_equal(host(sus), sus_o)
        assert np.array_equal(host(hist), hist_o)
        # T2 (float64 node math: tolerance) then T3 with the DEVICE's q / cdf on both sides (bit-exact)
        rs = np.random.default_rng(tick)
        W = rs.random((nodes, nodes)) * (0.1 / nodes)
        np.fill_diagonal(W, 0.0)
        r0s = rs.uniform(0.5, 2.0, nodes)
        pop = np.maximum(np.bincount(p["node_id"][:n][p["disease_state"][:n] >= 0], minlength=nodes), 0).astype(np.int32)
        q, cdf, prob, expd = K.tx_node_math(bfx, efx, hist, dev(W), 1.07, dev(r0s), dev(pop), 0.3, 2.0, rng=K.make_rng(seed, tick))
        q_o, cdf_o, prob_o, exp_o = oracle.tx_node_math_device(bfx_o, efx_o, hist_o, W, 1.07, r0s, pop, 0.3, 2.0, seed, tick)
        np.testing.assert_allclose(host(prob), prob_o, rtol=1e-6, atol=1e-15)
        np.testing.assert_allclose(host(cdf), cdf_o, rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(host(expd), exp_o, rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(host(q), q_o, rtol=2e-6, atol=1e-30)  # tau: float64 Newton solve rounded to float32
        q_h, cdf_h = host(q), host(cdf)
        new_o = oracle.tx_infect_bernoulli(nodes, n, ns, p["node_id"], p["strain"], p["disease_state"],
                                           p["acq_risk_multiplier"], q_h, cdf_h, seed=seed, tick=tick)
        new = K.tx_infect(nodes, n, ns, d["node_id"], d["strain"], d["disease_state"], d["acq_risk_multiplier"], q, cdf,
                          rng=K.make_rng(seed, tick))
        assert np.array_equal(host(new), new_o)
        # C1
        out_o = oracle.count_SEIRP(p["node_id"], p["disease_state"], p["strain"], p["potentially_paralyzed"], p["paralyzed"],
                                   nodes, ns, n)
        out = K.count_SEIRP(d["node_id"], d["disease_state"], d["strain"], d["potentially_paralyzed"], d["paralyzed"], nodes,
                            ns, n)
        for a, b in zip(out, out_o):
            assert np.array_equal(host(a), b)
        # every mutated column, every tick
        for k in ("disease_state", "strain", "exposure_timer", "infection_timer", "paralysis_timer", "potentially_paralyzed",
                  "paralyzed", "ipv_protected", "ri_timer"):
            assert np.array_equal(host(d[k]), p[k]), (tick, k)
    if n > 1000:
        assert (p["disease_state"][:n] == 1).sum() > 0 and new_o.sum() > 0


def test_node_math_many_nodes(K, oracle):
    """More than 1024 nodes: the network transfer is summed per chunk of 1024 source rows by a 2-D grid and the chunks are
    added in index order (lpk_kernels.cu, k_node_matvec_partial) -- same tolerances as the single-block form, and the same
    bits from run to run."""
    nodes, ns, seed, tick = 2600, 3, 11, 9
    rs = np.random.default_rng(3)
    bfx = (rs.random((nodes, ns)) * (rs.random((nodes, ns)) < 0.7) * 5.0 * 2**30).astype(np.int64)
    hist = rs.integers(0, 40, (nodes, 192)).astype(np.int32)
    hist[rs.random(nodes) < 0.05] = 0
    efx = (hist.sum(1) * 1.1 * 2**30).astype(np.int64)
    W = rs.random((nodes, nodes)) * (0.1 / nodes)
    np.fill_diagonal(W, 0.0)
    r0s = rs.uniform(0.5, 2.0, nodes)
    pop = rs.integers(5_000, 50_000, nodes).astype(np.int32)
    args = (dev(bfx), dev(efx), dev(hist), dev(W), 1.07, dev(r0s), dev(pop), 0.3, 2.0)
    q, cdf, prob, expd = K.tx_node_math(*args, rng=K.make_rng(seed, tick))
    q_o, cdf_o, prob_o, exp_o = oracle.tx_node_math_device(bfx, efx, hist, W, 1.07, r0s, pop, 0.3, 2.0, seed, tick)
    np.testing.assert_allclose(host(prob), prob_o, rtol=1e-6, atol=1e-15)
    np.testing.assert_allclose(host(cdf), cdf_o, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(host(expd), exp_o, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(host(q), q_o, rtol=2e-6, atol=1e-30)
    q1, cdf1 = host(q).copy(), host(cdf).copy()
    q2, cdf2, _, _ = K.tx_node_math(*args, rng=K.make_rng(seed, tick))
    assert np.array_equal(q1, host(q2)) and np.array_equal(cdf1, host(cdf2))


def test_empty_population_and_bad_arguments(K):
    z8 = torch.zeros(16, dtype=torch.int8, device="cuda")
    z16 = torch.zeros(16, dtype=torch.int16, device="cuda")
    z32 = torch.zeros(16, dtype=torch.int32, device="cuda")
    out = torch.full((4,), 5, dtype=torch.int32, device="cuda")
    K.get_deaths(4, 0, z8, z16, z32, 3, out)
    assert host(out).sum() == 0  # overwritten even when there is nobody
    with pytest.raises(ValueError):
        K.get_deaths(0, 16, z8, z16, z32, 3, out)
    with pytest.raises(ValueError):
        K.get_deaths(4, -1, z8, z16, z32, 3, out)
    f32 = torch.zeros(16, dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError):  # more strains than the ABI supports
        K.tx_step_prep(4, 16, 9, z8, np.ones(9), z8, z16, f32, f32)
    with pytest.raises(ValueError):  # misaligned view
        K.get_deaths(4, 8, z8[1:], z16, z32, 3, out)
    from laser_polio_b200 import _lpk

    with pytest.raises(_lpk.LpkError):  # host tensors are refused, there is no CPU fallback
        K.get_deaths(4, 8, z8.cpu(), z16, z32, 3, out)


def test_device_births_vs_oracle(K, oracle):
    """lpk_vd_births (reference model.py:1711-1734, Philox draws): births per node, node-major cohort, lifespans,
    slot counter, tile table, and the capacity-overflow flag -- bit-exact against the numpy restatement."""
    import ctypes as C

    from laser_polio_b200 import _lpk, utils

    n_nodes, count, cap, tick, seed, id_base = 37, 5_003, 20_000, 21, 0xABCDEF0123, 4096
    rs = np.random.default_rng(4)
    pop_prev = rs.integers(2_000, 400_000, n_nodes).astype(np.int32)
    pop_prev[5] = 0
    rate = rs.uniform(20, 45, n_nodes) / 365000.0
    cd = np.insert(utils.create_cumulative_deaths(int(pop_prev.sum()), 100).astype(np.int64), 0, 0)
    want = oracle.vd_births_device(pop_prev, rate, 7, cd, count, cap, seed, tick, id_base=id_base)
    births_o, node_o, dod_o, new_count_o, status_o = want
    assert status_o == 0 and births_o.sum() > 500 and births_o[5] == 0

    def run(capacity):
        d = {"state": torch.full((cap,), -1, dtype=torch.int8, device="cuda"), "node": torch.full((cap,), -1, dtype=torch.int16, device="cuda"),
             "dob": torch.full((cap,), -1, dtype=torch.int32, device="cuda"), "dod": torch.zeros(cap, dtype=torch.int32, device="cuda"),
             "births": torch.full((n_nodes,), 9, dtype=torch.int32, device="cuda"),
             "counts": torch.tensor([count, count], dtype=torch.int64, device="cuda"), "status": torch.zeros(1, dtype=torch.int32, device="cuda"),
             "off": torch.zeros(n_nodes + 1, dtype=torch.int32, device="cuda"), "coh": torch.zeros(2, dtype=torch.int64, device="cuda"),
             "tiles": torch.full(((cap + 511) // 512,), -7, dtype=torch.int32, device="cuda"),
             "rate": dev(rate), "pop": dev(pop_prev), "cd": dev(cd)}
        d["node"][:count] = 3
        a = _lpk.BirthsArgs()
        a.tick, a.n_nodes, a.seed, a.id_base, a.step_size = tick, n_nodes, seed, id_base, 7.0
        a.birth_rate, a.pop_prev, a.births_row = d["rate"].data_ptr(), d["pop"].data_ptr(), d["births"].data_ptr()
        a.counts, a.capacity, a.cum_deaths, a.max_year, a.ri_newborn_timer = d["counts"].data_ptr(), capacity, d["cd"].data_ptr(), 100, -1
        a.node_offsets_ws, a.cohort_ws, a.status = d["off"].data_ptr(), d["coh"].data_ptr(), d["status"].data_ptr()
        a.disease_state, a.node_id, a.date_of_birth, a.date_of_death = (d[k].data_ptr() for k in ("state", "node", "dob", "dod"))
        a.ri_timer, a.tile_node = None, d["tiles"].data_ptr()
        _lpk.check(_lpk.lib().lpk_vd_births(C.byref(a), _lpk.stream_handle()), "lpk_vd_births")
        torch.cuda.synchronize()
        return d

    d = run(cap)
    total = int(births_o.sum())
    assert np.array_equal(host(d["births"]), births_o)
    assert host(d["counts"]).tolist() == [count, new_count_o] and int(d["status"].item()) == 0
    assert np.array_equal(host(d["node"])[count:count + total], node_o)
    assert np.array_equal(host(d["dod"])[count:count + total], dod_o)
    assert np.all(host(d["dob"])[count:count + total] == tick) and np.all(host(d["state"])[count:count + total] == 0)
    assert np.all(host(d["state"])[count + total:] == -1) and np.all(host(d["node"])[count + total:] == -1)
    assert np.all(dod_o > tick)
    tiles = host(d["tiles"])
    node_all = host(d["node"])
    for t_ in range(count // 512, (count + total - 1) // 512 + 1):
        seg = node_all[t_ * 512:(t_ + 1) * 512]
        assert tiles[t_] == (seg[0] if len(seg) == 512 and np.all(seg == seg[0]) and seg[0] >= 0 else -1)
    assert np.all(tiles[: count // 512] == -7)  # untouched tiles keep their value
    d = run(count + total - 1)  # one slot short: nobody is born, the flag is raised
    assert int(d["status"].item()) == 1 and host(d["counts"]).tolist() == [count, count] and host(d["births"]).sum() == 0
