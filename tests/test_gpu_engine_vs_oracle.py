"""The fused engine -- the thing bench.py times -- against the CPU oracle's tick loop directly (not against the CUDA
component kernels): every results array and every agent column bit for bit, the node-level tau at 2e-6, on a full-feature
schedule (vital dynamics, RI, campaign days).  The large case is the Nigeria node count at 2e7 agents with a birth rate
high enough that appended cohorts make up more than a tenth of the table by the end."""

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def check(oracle):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import engine_check

    return engine_check


def test_engine_vs_oracle_small_nodes(check):
    out = check.run_and_compare(300_000, 23, 50, seed=11, cbr=60.0)
    assert out["new_exposed"] > 500 and out["deaths"] > 0 and out["births"] > 0 and out["ri_vaccinated"] > 0 and out["sia_protected"] > 0


def test_engine_vs_oracle_774_nodes_20M_agents_60_days(check):
    # cbr 800 / 1000 / year: 60 days of births add > 10 % of the table as appended cohorts (mixed-node pairs)
    out = check.run_and_compare(20_000_000, 774, 60, seed=20261018, cbr=800.0, node_math_ticks=(1, 7, 14, 21, 43, 59))
    assert out["cohort_share"] >= 0.10, out
    assert out["new_exposed"] > 100_000 and out["deaths"] > 1000 and out["ri_vaccinated"] > 1000 and out["sia_protected"] > 100_000
    assert out["new_potentially_paralyzed"] > 0


def test_engine_vs_oracle_with_compaction(check):
    """pars.compact_every: every 10 ticks the live agents are stably re-sorted by node and the dead leave the swept range (the
    oracle's tick loop applies the same table operation); columns come back in the reference's order, bit for bit, tombstones
    and cohorts included.  cbr 400 and the synthetic death dates make both births and deaths plentiful."""
    out = check.run_and_compare(400_000, 23, 55, seed=12, cbr=400.0, compact_every=10, node_math_ticks=(1, 10, 11, 30, 54))
    assert out["compactions"] == 5 and out["deaths"] > 200 and out["births"] > 10_000 and out["new_exposed"] > 500
    assert out["sia_protected"] > 0 and out["ri_vaccinated"] > 0


def test_engine_vs_oracle_ten_years(check):
    """3700 days on a small table: the agenda's 63-day check-ins, int8 deadline wrap-around, more than 254 RI ticks (the lazy RI
    debt is paid and the table re-based on the way), ~80 campaigns, hundreds of cohorts -- still bit-identical to the oracle."""
    out = check.run_and_compare(60_000, 5, 3700, seed=13, cbr=37.0, node_math_ticks=(1, 1000, 3600))
    assert out["ticks"] == 3701 and out["deaths"] > 1000 and out["births"] > 10_000 and out["sia_protected"] > 10_000
    assert out["calls"]["tick_pass"] >= 3690
