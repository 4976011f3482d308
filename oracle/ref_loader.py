"""TEST INFRASTRUCTURE ONLY -- loads the reference's own numba kernels, read-only.

The reference package cannot be imported in the build container (its top-level
imports need laser_core / sciris / matplotlib, none installed), but every
hot-path kernel is a free function over plain numpy arrays.  This module slices
those FunctionDefs out of ``/root/reference/src/laser_polio/model.py`` with
``ast`` at *run time*, compiles them under the local numba and hands them back.
Nothing from the reference is written into this repository.

Only ``tests/golden/make_golden.py`` and the optional live-pinning tests call
this; it needs ``/root/reference`` and therefore never runs on the GPU box.

Injected-uniform mode (north-star gate 1): each ``np.random.random()`` /
``np.random.rand()`` call site inside a kernel body is replaced textually by a
read from an extra per-agent array argument, so the reference's integer state
machine can be driven by the very same uniforms as the oracle and the CUDA
kernels.  Call sites (reference model.py): disease_state_step_kernel:441,
fast_ri:1845 and :1852, fast_sia:2049.
"""

from __future__ import annotations

import ast
import re
from pathlib import Path

REF_ROOT = Path("/root/reference")
REF_MODEL = REF_ROOT / "src" / "laser_polio" / "model.py"

HOT_FUNCS = (
    "disease_state_step",
    "disease_state_step_kernel",
    "count_SEIRP",
    "count_SEIRP_kernel",
    "tx_step_prep",
    "tx_step_prep_kernel",
    "tx_infect_nb",
    "get_deaths",
    "fast_ri",
    "fast_sia",
)


def available() -> bool:
    return REF_MODEL.exists()


def _slice_function(src: str, tree: ast.Module, name: str) -> str:
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            first = min([node.lineno] + [d.lineno for d in node.decorator_list])
            lines = src.splitlines()[first - 1 : node.end_lineno]
            return "\n".join(lines) + "\n"
    raise KeyError(name)


def _strip_signature(text: str) -> str:
    """Drop an explicit numba signature tuple from an ``@nb.njit(( ... ), ...)`` decorator."""
    m = re.search(r"@nb\.njit\(\s*\(", text)
    if not m:
        return text
    # find the matching close paren of the signature tuple
    i = m.end() - 1
    depth = 0
    for j in range(i, len(text)):
        if text[j] == "(":
            depth += 1
        elif text[j] == ")":
            depth -= 1
            if depth == 0:
                break
    rest = text[j + 1 :].lstrip()
    if rest.startswith(","):
        rest = rest[1:]
    return text[: m.end() - 1] + rest


def load(inject_uniforms: bool = False) -> dict:
    """Return ``{name: callable}`` for the reference hot-path functions.

    With ``inject_uniforms=True`` the three RNG-consuming kernels take extra
    trailing array arguments instead of calling numba's thread-local RNG:

    * ``disease_state_step_kernel(..., local_new_paralyzed, u_inj)``
    * ``fast_ri(..., ri_vaccine_strain, u1_inj, u2_inj)``
    * ``fast_sia(..., sia_vaccine_strain, u_inj)``

    and ``disease_state_step`` forwards a ``u_inj`` keyword to its kernel.
    """
    import numba as nb
    import numpy as np

    src = REF_MODEL.read_text()
    tree = ast.parse(src)
    ns: dict = {"nb": nb, "np": np}
    for name in HOT_FUNCS:
        text = _slice_function(src, tree, name).replace("cache=True", "cache=False")
        if inject_uniforms:
            if name == "disease_state_step_kernel":
                assert text.count("np.random.random()") == 1
                text = text.replace("np.random.random()", "u_inj[i]")
                text = text.replace("local_new_paralyzed,\n):", "local_new_paralyzed,\n    u_inj,\n):")
            elif name == "disease_state_step":
                text = text.replace("new_paralyzed,\n):", "new_paralyzed,\n    u_inj=None,\n):")
                text = text.replace("local_new_paralyzed,\n    )", "local_new_paralyzed,\n        u_inj,\n    )")
            elif name == "fast_ri":
                assert text.count("np.random.rand()") == 2
                text = _strip_signature(text)
                text = text.replace("np.random.rand()", "u1_inj[i]", 1).replace("np.random.rand()", "u2_inj[i]", 1)
                text = text.replace("ri_vaccine_strain,\n):", "ri_vaccine_strain,\n    u1_inj,\n    u2_inj,\n):")
            elif name == "fast_sia":
                assert text.count("np.random.rand()") == 1
                text = text.replace("np.random.rand()", "u_inj[i]")
                text = text.replace("sia_vaccine_strain,\n):", "sia_vaccine_strain,\n    u_inj,\n):")
        exec(compile(text, f"<reference:{name}>", "exec"), ns)  # noqa: S102
    return {k: ns[k] for k in HOT_FUNCS}
