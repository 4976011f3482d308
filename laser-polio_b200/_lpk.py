"""ctypes binding of liblpk.so (the C ABI declared in include/lpk.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There
is NO fallback: if the shared object is missing or a call fails, this module raises.
Device memory, streams and process groups come from PyTorch; this module only
passes raw device pointers and the current stream handle across the C boundary.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
SO_PATH = _HERE / "liblpk.so"

LPK_OK, LPK_ERR_ARG, LPK_ERR_CUDA = 0, -1, -2
MAX_STRAINS = 4
RISK_BINS = 192
FX_SCALE = float(2**30)

EXPORTS = (
    "lpk_last_error", "lpk_version", "lpk_philox_selftest", "lpk_get_deaths", "lpk_disease_state_step", "lpk_fast_ri",
    "lpk_fast_sia", "lpk_tx_step_prep", "lpk_tx_node_math", "lpk_tx_infect", "lpk_count_seirp", "lpk_build_tile_nodes",
    "lpk_tick_pass", "lpk_tick_node", "lpk_vd_births", "lpk_run_days", "lpk_run_day_pass", "lpk_run_day_node", "lpk_run_release",
    "lpk_xchg_create", "lpk_xchg_connect", "lpk_xchg_disconnect", "lpk_xchg_destroy", "lpk_hot_build", "lpk_hot_settle", "lpk_hot_padded", "lpk_hot_risk_e0",
    "lpk_init_heterogeneity", "lpk_init_timers", "lpk_init_demography", "lpk_init_missed",
    "lpk_net_haversine", "lpk_net_gravity", "lpk_net_radiation", "lpk_net_row_normalize",
)


class LpkError(RuntimeError):
    pass


class Rng(C.Structure):
    """struct lpk_rng (include/lpk.h)."""

    _fields_ = [
        ("seed", C.c_uint64), ("tick", C.c_uint32), ("_pad", C.c_uint32),
        ("u1", C.c_void_p), ("u2", C.c_void_p), ("x", C.c_void_p), ("id_base", C.c_uint64),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load liblpk.so; raise loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not SO_PATH.exists():
            raise LpkError(
                f"{SO_PATH} is missing: the CUDA extension has not been built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback."
            )
        _lib = C.CDLL(str(SO_PATH))
        _lib.lpk_last_error.restype = C.c_char_p
        _lib.lpk_hot_padded.restype = C.c_int64
        _lib.lpk_hot_padded.argtypes = [C.c_int64]
        _lib.lpk_hot_risk_e0.restype = C.c_int32
        _lib.lpk_hot_risk_e0.argtypes = [C.c_float]
        for name in EXPORTS:
            if not hasattr(_lib, name):
                raise LpkError(f"liblpk.so does not export {name}")
    return _lib


def check(status: int, where: str) -> None:
    if status == LPK_OK:
        return
    msg = lib().lpk_last_error().decode()
    if status == LPK_ERR_ARG:
        raise ValueError(f"{where}: {msg}")
    raise LpkError(f"{where}: {msg}")


def ptr(t):
    """Raw device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise LpkError("expected a CUDA tensor; the device path has no host fallback")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def stream_handle():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_rng(seed=0, tick=0, u1=None, u2=None, x=None, id_base=0) -> Rng:
    r = Rng()
    r.seed, r.tick, r._pad = int(seed) & 0xFFFFFFFFFFFFFFFF, int(tick) & 0xFFFFFFFF, 0
    r.u1 = u1.data_ptr() if u1 is not None else None
    r.u2 = u2.data_ptr() if u2 is not None else None
    r.x = x.data_ptr() if x is not None else None
    r.id_base = int(id_base)
    if r.id_base % 256:
        raise ValueError("id_base must be a multiple of 256 (exposure draws are shared by aligned groups of 256 agents)")
    r._keep = (u1, u2, x)  # the struct only holds raw pointers: keep the tensors alive with it
    return r


# ----------------------------------------------------------------------------- fused-tick structs (include/lpk.h)
_VP = C.c_void_p


class People(C.Structure):
    """struct lpk_people"""

    _fields_ = [(n, _VP) for n in (
        "disease_state", "strain", "exposure_timer", "infection_timer", "paralysis_timer", "potentially_paralyzed",
        "paralyzed", "ipv_protected", "chronically_missed", "node_id", "ri_timer", "acq_risk_multiplier",
        "daily_infectivity", "date_of_birth", "date_of_death", "tile_node")] + [
            ("capacity", C.c_int64), ("hot", _VP), ("pair_min_dod", _VP), ("risk_e0", C.c_int32), ("pair_ri_max", _VP), ("rec", _VP), ("ri_k", _VP)]


class TickArgs(C.Structure):
    """struct lpk_tick_args"""

    _fields_ = [
        ("flags", C.c_uint32), ("tick", C.c_int32), ("n_nodes", C.c_int32), ("n_strains", C.c_int32),
        ("seed", C.c_uint64), ("id_base", C.c_uint64), ("counts", _VP),
        ("q_prev", _VP), ("cdf_prev", _VP), ("new_exposed_prev", _VP), ("new_exposed_by_strain_prev", _VP),
        ("tx_hits", _VP), ("tx_hits_by_strain", _VP),
        ("p_paralysis", C.c_float), ("new_potential", _VP), ("new_paralyzed", _VP),
        ("deaths", _VP), ("dead_pp", _VP), ("dead_par", _VP),
        ("ri_step", C.c_int32), ("ri_strain", C.c_int32), ("vx_prob_ri", _VP), ("vx_prob_ipv", _VP),
        ("ri_vaccinated", _VP), ("ri_protected", _VP), ("ipv_vaccinated", _VP),
        ("new_exposed", _VP), ("new_exposed_by_strain", _VP), ("ri_new_exposed_by_strain", _VP),
        ("sia_targeted", _VP), ("vx_prob_sia", _VP), ("sia_vx_eff", C.c_double),
        ("sia_min_age", C.c_int32), ("sia_max_age", C.c_int32), ("sia_strain", C.c_int32), ("sia_event_idx", C.c_uint32),
        ("sia_vaccinated", _VP), ("sia_protected", _VP), ("sia_new_exposed_by_strain", _VP),
        ("strain_r0_scalars", C.c_double * MAX_STRAINS),
        ("beta_fx", _VP), ("E_cur", _VP), ("I_cur", _VP), ("exposure_fx", _VP), ("sus", _VP), ("risk_hist", _VP), ("R_cur", _VP),
        ("uniform_agents", C.c_int64), ("ri_lazy_k", C.c_int32), ("work_counter", _VP), ("work_counter_next", _VP),
    ]


class NodeArgs(C.Structure):
    """struct lpk_node_args"""

    _fields_ = [
        ("flags", C.c_uint32), ("tick", C.c_int32), ("n_nodes", C.c_int32), ("n_strains", C.c_int32), ("seed", C.c_uint64),
        ("beta_fx", _VP), ("exposure_fx", _VP), ("risk_hist", _VP), ("network", _VP), ("r0_scalars", _VP),
        ("beta_seasonality", C.c_double), ("zero_inflation", C.c_double), ("dispersion", C.c_double),
        ("q", _VP), ("strain_cdf", _VP), ("prob", _VP), ("expected", _VP), ("rowsum_ws", _VP),
        ("pop_prev", _VP), ("pop", _VP), ("births_row", _VP), ("deaths_row", _VP),
        ("deaths", _VP), ("dead_pp", _VP), ("dead_par", _VP),
        ("cur_potp", _VP), ("cur_p", _VP), ("new_potential", _VP), ("new_paralyzed", _VP), ("potp_row", _VP), ("p_row", _VP),
        ("E_by_strain_prev", _VP), ("I_by_strain_prev", _VP), ("E_prev", _VP), ("I_prev", _VP),
        ("E_cur", _VP), ("I_cur", _VP), ("E_snap", _VP), ("I_snap", _VP), ("tx_hits_by_strain", _VP), ("any_cases", _VP),
        ("sus", _VP), ("R_cur", _VP), ("tx_hits", _VP), ("S_snap", _VP), ("R_snap", _VP), ("S_prev", _VP), ("R_prev", _VP),
        ("counts", _VP), ("node_lo", C.c_int32), ("node_hi", C.c_int32),
        ("xchg_flags", _VP), ("xchg_world", C.c_int32), ("xchg_seq", C.c_uint32), ("matvec_ws", _VP),
    ]


class BirthsArgs(C.Structure):
    """struct lpk_births_args"""

    _fields_ = [
        ("tick", C.c_int32), ("n_nodes", C.c_int32), ("seed", C.c_uint64), ("id_base", C.c_uint64), ("step_size", C.c_double),
        ("birth_rate", _VP), ("pop_prev", _VP), ("births_row", _VP), ("counts", _VP), ("capacity", C.c_int64),
        ("cum_deaths", _VP), ("max_year", C.c_int32), ("ri_newborn_timer", C.c_int32), ("node_offsets_ws", _VP),
        ("cohort_ws", _VP), ("status", _VP), ("disease_state", _VP), ("node_id", _VP), ("date_of_birth", _VP),
        ("date_of_death", _VP), ("ri_timer", _VP), ("tile_node", _VP),
        ("acq_risk_multiplier", _VP), ("sus", _VP), ("exposure_fx", _VP), ("risk_hist", _VP),
        ("hot", _VP), ("pair_min_dod", _VP), ("risk_e0", C.c_int32), ("rec", _VP), ("strain", _VP), ("exposure_timer", _VP),
        ("infection_timer", _VP), ("paralysis_timer", _VP), ("potentially_paralyzed", _VP), ("paralyzed", _VP), ("ipv_protected", _VP),
        ("ri_k", _VP), ("pair_ri_max", _VP), ("ri_lazy_k", C.c_int32),
        ("ri_step", C.c_int32),
    ]


class Rows(C.Structure):
    """struct lpk_rows"""

    NAMES = ("S", "E", "I", "R", "pop", "births", "deaths", "new_exposed", "new_potentially_paralyzed", "new_paralyzed",
             "potentially_paralyzed", "paralyzed", "ri_vaccinated", "ri_protected", "ipv_vaccinated", "sia_vaccinated", "sia_protected",
             "E_by_strain", "I_by_strain", "new_exposed_by_strain", "ri_new_exposed_by_strain", "sia_new_exposed_by_strain")
    _fields_ = [(n, _VP) for n in NAMES] + [("sink", _VP)]


class Day(C.Structure):
    """struct lpk_day"""

    _fields_ = [("tick", C.c_int32), ("flags", C.c_uint32), ("beta_seasonality", C.c_double), ("sia_targeted", _VP),
                ("sia_vx_eff", C.c_double), ("sia_min_age", C.c_int32), ("sia_max_age", C.c_int32), ("sia_strain", C.c_int32),
                ("_pad", C.c_int32)]


class Run(C.Structure):
    """struct lpk_run"""

    _fields_ = [("people", People), ("tick", TickArgs), ("node", NodeArgs), ("births", BirthsArgs), ("rows", Rows),
                ("zero_pop", _VP), ("any_cases", _VP), ("work_counters", _VP), ("xchg", _VP),
                ("pending", C.c_int32), ("ri_lazy_k", C.c_int32), ("rowsums_valid", C.c_int32), ("graph", C.c_int32),
                ("seq", C.c_uint32), ("_pad", C.c_int32)]


XCHG_HANDLE_BYTES = 64


class Dist(C.Structure):
    """struct lpk_dist"""

    _fields_ = [("kind", C.c_int32), ("a", C.c_double), ("b", C.c_double)]


class DemogArgs(C.Structure):
    """struct lpk_demog_args"""

    _fields_ = [
        ("start", C.c_int64), ("end", C.c_int64), ("date_of_birth", _VP), ("date_of_death", _VP), ("ri_timer", _VP),
        ("bin_cdf", _VP), ("bin_lo", _VP), ("bin_hi", _VP), ("n_bins", C.c_int32), ("cum_deaths", _VP), ("max_year", C.c_int32),
        ("seed", C.c_uint64), ("id_base", C.c_uint64),
    ]


DIST_KINDS = {"constant": 0, "exponential": 1, "gamma": 2, "lognormal": 3, "normal": 4, "poisson": 5, "uniform": 6}
MISSED_WS_WORDS = 65536 + 4

F_PENDING, F_STAGES, F_DEATHS, F_RI, F_SIA, F_ROWSUMS = 1, 2, 4, 8, 16, 32
TILE_AGENTS = 512


def dp(t):
    """data_ptr of a tensor (None -> NULL) for struct fields."""
    return None if t is None else t.data_ptr()
