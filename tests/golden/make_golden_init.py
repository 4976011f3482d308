"""Golden samples for the population initialisers, produced by the REFERENCE's own code.

Runs only in the build container (needs /root/reference).  ``populate_heterogeneous_values`` (model.py:816-866) is sliced
out of the reference's model.py with ``ast`` and executed unmodified under numpy / scipy with a seeded global stream; the
timer block of ``DiseaseState_ABM.__init__`` (model.py:575-587) is sliced as source lines and executed on a stand-in
``sim.people`` with the reference's own ``lp.poisson / lp.gamma / lp.lognormal`` samplers (distributions.py, sliced the same
way).  The draws come from numpy's Mersenne stream, which the Philox-keyed device samplers cannot replay, so the fixture is
a SAMPLE: tests compare distributions (two-sample KS, moments, rank correlation), not values.

    python tests/golden/make_golden_init.py   ->  tests/golden/init_ref.npz  (~0.5 MB)
"""

from __future__ import annotations

import ast
import logging
import sys
import types
from pathlib import Path

import numpy as np
from scipy import stats

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/src/laser_polio")
OUT = Path(__file__).resolve().parent / "init_ref.npz"
N = 40_000


def slice_def(path: Path, name: str) -> str:
    src = path.read_text()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name == name:
            first = min([node.lineno] + [d.lineno for d in node.decorator_list])
            return "\n".join(src.splitlines()[first - 1:node.end_lineno]) + "\n"
    raise KeyError(name)


def reference_distributions():
    """The reference's Distribution class and lp.* constructors (distributions.py), executed as they are."""
    import numba as nb

    src = (REF / "distributions.py").read_text().replace("cache=True", "cache=False")  # the checkout is read-only
    ns = {"np": np, "nb": nb, "__name__": "ref_distributions"}
    exec(compile(src, "<reference:distributions.py>", "exec"), ns)  # noqa: S102
    return types.SimpleNamespace(**{k: ns[k] for k in ("poisson", "gamma", "lognormal", "normal", "constant")})


def main():
    lp = reference_distributions()
    np.random.seed(20261017)
    pars = types.SimpleNamespace(risk_mult_var=4.0, r0=14.0, dur_inf=lp.gamma(shape=4.51, scale=5.32), corr_risk_inf=0.8,
                                 individual_heterogeneity=True, dur_exp=lp.poisson(lam=3),
                                 t_to_paralysis=lp.lognormal(mean=12.5, sigma=3.5))
    ns = {"np": np, "stats": stats, "logger": logging.getLogger("ref")}
    exec(compile(slice_def(REF / "model.py", "populate_heterogeneous_values"), "<reference:populate_heterogeneous_values>", "exec"), ns)  # noqa: S102
    risk, inf = np.zeros(N, np.float32), np.zeros(N, np.float32)
    ns["populate_heterogeneous_values"](0, N, risk, inf, pars)

    # timer block, model.py:575-587, verbatim source lines on a stand-in frame
    lines = (REF / "model.py").read_text().splitlines()
    block = "\n".join(line[8:] for line in lines[574:587])  # body of __init__, de-indented
    assert "sim.people.exposure_timer[:] = self.pars.dur_exp(sim.people.capacity)" in block and "paralysis_timer" in block
    people = types.SimpleNamespace(capacity=N, exposure_timer=np.zeros(N, np.int8), infection_timer=np.zeros(N, np.int8),
                                   paralysis_timer=np.zeros(N, np.int8))
    env = {"np": np, "sim": types.SimpleNamespace(people=people), "self": types.SimpleNamespace(pars=pars)}
    exec(compile(block, "<reference:model.py:575-587>", "exec"), env)  # noqa: S102
    np.savez_compressed(OUT, acq_risk_multiplier=risk, daily_infectivity=inf, exposure_timer=people.exposure_timer,
                        infection_timer=people.infection_timer, paralysis_timer=people.paralysis_timer,
                        r0=np.float64(pars.r0), risk_mult_var=np.float64(4.0), corr_risk_inf=np.float64(0.8))
    print("wrote", OUT, "risk mean", risk.mean(), "inf mean", inf.mean(), "spearman", stats.spearmanr(risk, inf)[0],
          "timers", people.exposure_timer.mean(), people.infection_timer.mean(), people.paralysis_timer.mean())


if __name__ == "__main__":
    main()
