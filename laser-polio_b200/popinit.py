"""Population initialisation in HBM: the reference's one-off column draws as device kernels.

The reference fills the agent table on the host at construction -- ``populate_heterogeneous_values`` (model.py:816-866,
flagged slow by its own FIXME), the timer block of ``DiseaseState_ABM.__init__`` (model.py:571-587), ages / lifespans /
routine-immunisation dates (model.py:1578-1596, 1605-1611, 1893-1894) and the chronically missed (model.py:154-159).
The functions here keep those names and argument meanings but write torch CUDA tensors through ``liblpk.so``
(``csrc/lpk_init.cu``): every value is a pure function of (seed, agent id, stage), so slices can be drawn in any order, on
any number of GPUs, with the same result.  ``init_frame_device`` fills the pinned host columns of a ``LaserFrame`` the same
way (kernels + one D2H per column) -- what ``SEIR_ABM`` does when ``pars.device_init`` is set.

No CPU fallback: without the extension or a CUDA device these raise (``_lpk.lib``).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lpk
from .distributions import Distribution

__all__ = ["dist_struct", "populate_heterogeneous_values", "init_timers", "init_demography", "init_missed", "init_frame_device",
           "heterogeneity_parameters"]


def dist_struct(d) -> _lpk.Dist:
    """``lp.poisson(lam=3)`` etc. -> ``struct lpk_dist`` (include/lpk.h); parameter defaults as reference distributions.py:36-108."""
    if not isinstance(d, Distribution):
        raise TypeError(f"expected an lp.* Distribution, got {type(d).__name__}")
    k, p = d.dist_type, d.pars
    if k == "constant":
        a, b = p.get("value", 1), 0.0
    elif k == "exponential":
        a, b = p.get("scale", 1.0), 0.0
    elif k == "gamma":
        a, b = p.get("shape", 2.0), p.get("scale", 1.0)
    elif k == "lognormal":
        m, s = p.get("mean", 1.0), p.get("sigma", 0.5)
        if m <= 0:
            return _lpk.Dist(_lpk.DIST_KINDS["constant"], 0.0, 0.0)  # the reference returns zeros for a non-positive mean
        a, b = np.log(m**2 / np.sqrt(s**2 + m**2)), np.sqrt(np.log(s**2 / m**2 + 1))
    elif k == "normal":
        a, b = p.get("mean", 0.0), p.get("std", 1.0)
    elif k == "poisson":
        a, b = p.get("lam", 5), 0.0
    elif k == "uniform":
        a, b = p.get("min", 2), p.get("max", 10)
    else:
        raise ValueError(f"Unsupported distribution: {k}")
    return _lpk.Dist(_lpk.DIST_KINDS[k], float(a), float(b))


def heterogeneity_parameters(pars, mean_dur_inf=None):
    """mu_ln, sigma_ln, scale_gamma, rho, mean_gamma of reference model.py:838-847.  ``mean_dur_inf``: the mean infectious
    period; the reference estimates it from 1000 draws of ``pars.dur_inf`` (model.py:842), which is the default here too."""
    var_ln = float(pars.risk_mult_var)
    mu_ln = float(np.log(1.0 / np.sqrt(var_ln + 1.0)))
    sigma_ln = float(np.sqrt(np.log(var_ln + 1.0)))
    if mean_dur_inf is None:
        mean_dur_inf = float(np.mean(pars.dur_inf(1000)))
    mean_gamma = float(pars.r0) / mean_dur_inf
    rho = float(2.0 * np.sin(np.pi * float(pars.corr_risk_inf) / 6.0))
    return mu_ln, sigma_ln, max(mean_gamma, 1e-10), rho, mean_gamma


def populate_heterogeneous_values(start, end, acq_risk_out, infectivity_out, pars, seed=None, id_base=0, mean_dur_inf=None):
    """Device form of reference ``populate_heterogeneous_values(start, end, acq_risk_out, infectivity_out, pars)``:
    ``acq_risk_out`` / ``infectivity_out`` are float32 CUDA tensors, filled in place on ``[start, end)``."""
    mu_ln, sigma_ln, scale_gamma, rho, mean_gamma = heterogeneity_parameters(pars, mean_dur_inf)
    seed = int(pars.seed if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
    if acq_risk_out.dtype != torch.float32 or infectivity_out.dtype != torch.float32:
        raise TypeError("acq_risk_out / infectivity_out must be float32")
    _lpk.check(_lpk.lib().lpk_init_heterogeneity(
        C.c_int64(start), C.c_int64(end), _lpk.ptr(acq_risk_out), _lpk.ptr(infectivity_out), C.c_double(mu_ln), C.c_double(sigma_ln),
        C.c_double(scale_gamma), C.c_double(rho), C.c_int32(int(bool(pars.individual_heterogeneity))), C.c_double(mean_gamma),
        C.c_uint64(seed), C.c_uint64(id_base), _lpk.stream_handle()), "lpk_init_heterogeneity")


def init_timers(start, end, exposure_timer, infection_timer, paralysis_timer, pars, seed=None, id_base=0):
    """Timer block of reference ``DiseaseState_ABM.__init__`` (model.py:571-587) on int8 CUDA tensors."""
    seed = int(pars.seed if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
    de, di, dp = dist_struct(pars.dur_exp), dist_struct(pars.dur_inf), dist_struct(pars.t_to_paralysis)
    for t in (exposure_timer, infection_timer, paralysis_timer):
        if t.dtype != torch.int8:
            raise TypeError("timer columns must be int8")
    _lpk.check(_lpk.lib().lpk_init_timers(
        C.c_int64(start), C.c_int64(end), _lpk.ptr(exposure_timer), _lpk.ptr(infection_timer), _lpk.ptr(paralysis_timer),
        C.byref(de), C.byref(di), C.byref(dp), C.c_uint64(seed), C.c_uint64(id_base), _lpk.stream_handle()), "lpk_init_timers")


def init_demography(start, end, date_of_birth, date_of_death, ri_timer, pyramid, cum_deaths, seed, id_base=0, max_year=100):
    """Ages from the pyramid, lifespans, routine-immunisation dates (reference model.py:1578-1596, 1605-1611, 1893-1894).

    ``pyramid``: rows ``[min_age_yr, max_age_yr, M, F]`` (``load_pyramid_csv``); ``cum_deaths``: the table handed to
    ``KaplanMeierEstimator`` (``create_cumulative_deaths``), or None with ``date_of_death`` None; ``ri_timer`` may be None."""
    dev = date_of_birth.device
    pyr = np.asarray(pyramid)
    lo = np.maximum(pyr[:, 0].astype(np.int64) * 365, 1).astype(np.int32)  # nobody is born on day 0 (model.py:1586)
    hi = ((pyr[:, 1].astype(np.int64) + 1) * 365).astype(np.int32)
    cdf = np.cumsum((pyr[:, 2] + pyr[:, 3]).astype(np.float64))
    a = _lpk.DemogArgs()
    keep = [torch.from_numpy(cdf).to(dev), torch.from_numpy(lo).to(dev), torch.from_numpy(hi).to(dev)]
    a.start, a.end = int(start), int(end)
    a.date_of_birth, a.date_of_death, a.ri_timer = _lpk.dp(date_of_birth), _lpk.dp(date_of_death), _lpk.dp(ri_timer)
    a.bin_cdf, a.bin_lo, a.bin_hi, a.n_bins = keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), len(cdf)
    if date_of_death is not None:
        cd = np.insert(np.asarray(cum_deaths, dtype=np.int64), 0, 0)  # cd[y] = deaths before age y (core.KaplanMeierEstimator)
        if len(cd) < max_year + 2:
            raise ValueError("cum_deaths must cover ages 0..max_year")
        keep.append(torch.from_numpy(cd).to(dev))
        a.cum_deaths = keep[3].data_ptr()
    a.max_year = int(max_year)
    a.seed, a.id_base = int(seed) & 0xFFFFFFFFFFFFFFFF, int(id_base)
    _lpk.check(_lpk.lib().lpk_init_demography(C.byref(a), _lpk.stream_handle()), "lpk_init_demography")
    torch.cuda.current_stream().synchronize()  # the small tables above must outlive the kernel


def init_missed(n, n_missed, chronically_missed, seed, id_base=0):
    """Exactly ``n_missed`` of the first ``n`` agents flagged, uniformly without replacement (reference model.py:154-159)."""
    if chronically_missed.dtype != torch.uint8:
        raise TypeError("chronically_missed must be uint8")
    ws = torch.zeros(_lpk.MISSED_WS_WORDS + 2, dtype=torch.int32, device=chronically_missed.device)
    _lpk.check(_lpk.lib().lpk_init_missed(C.c_int64(n), C.c_int64(n_missed), _lpk.ptr(chronically_missed),
                                           C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), C.c_uint64(id_base), _lpk.ptr(ws),
                                           _lpk.stream_handle()), "lpk_init_missed")
    torch.cuda.current_stream().synchronize()


def init_frame_device(people, pars, columns, device=None, pyramid=None, cum_deaths=None, chunk=1 << 26, start=0):
    """Fill the host columns of ``people`` (a LaserFrame) named in ``columns`` by drawing them on the GPU, ``chunk`` agents
    at a time (kernels + one D2H per column and chunk): what the components do when ``pars.device_init`` is set.

    ``columns`` is any subset of {"heterogeneity", "timers", "demography", "missed"}; ``start`` restricts "heterogeneity" and
    "timers" to the slots [start, capacity) (the unborn tail of a table loaded from a snapshot)."""
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    cap, count, seed = int(people.capacity), int(people.count), int(pars.seed)
    out = lambda name, lo, hi: torch.from_numpy(getattr(people, name)[lo:hi])  # noqa: E731
    if "heterogeneity" in columns:
        mean_dur = float(np.mean(pars.dur_inf(1000)))
        for lo in range(int(start), cap, chunk):
            hi = min(cap, lo + chunk)
            r, f = (torch.empty(hi - lo, dtype=torch.float32, device=device) for _ in range(2))
            populate_heterogeneous_values(0, hi - lo, r, f, pars, seed=seed, id_base=lo, mean_dur_inf=mean_dur)
            out("acq_risk_multiplier", lo, hi).copy_(r)
            out("daily_infectivity", lo, hi).copy_(f)
    if "timers" in columns:
        for lo in range(int(start), cap, chunk):
            hi = min(cap, lo + chunk)
            e, i, p = (torch.empty(hi - lo, dtype=torch.int8, device=device) for _ in range(3))
            init_timers(0, hi - lo, e, i, p, pars, seed=seed, id_base=lo)
            out("exposure_timer", lo, hi).copy_(e)
            out("infection_timer", lo, hi).copy_(i)
            out("paralysis_timer", lo, hi).copy_(p)
    if "demography" in columns:
        has_dod, has_ri = hasattr(people, "date_of_death") and cum_deaths is not None, hasattr(people, "ri_timer")
        for lo in range(0, count, chunk):
            hi = min(count, lo + chunk)
            dob = torch.empty(hi - lo, dtype=torch.int32, device=device)
            dod = torch.empty(hi - lo, dtype=torch.int32, device=device) if has_dod else None
            ri = torch.empty(hi - lo, dtype=torch.int16, device=device) if has_ri else None
            init_demography(0, hi - lo, dob, dod, ri, pyramid, cum_deaths, seed, id_base=lo)
            out("date_of_birth", lo, hi).copy_(dob)
            if has_dod:
                out("date_of_death", lo, hi).copy_(dod)
            if has_ri:
                out("ri_timer", lo, hi).copy_(ri)
    if "missed" in columns:
        m = torch.zeros(count, dtype=torch.uint8, device=device)
        init_missed(count, int(float(pars.missed_frac) * count), m, seed)
        out("chronically_missed", 0, count).copy_(m)
