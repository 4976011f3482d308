"""Device versions of the reference's hot-path free functions.

Same names, same argument order and meaning as ``laser_polio.model`` (reference
``src/laser_polio/model.py``): the arrays are torch CUDA tensors (the agent
structure-of-arrays lives in HBM) instead of numpy arrays, the reference's
``[threads, nodes]`` scratch arguments become per-node output tensors, and the
uniform source is explicit (``rng``: Philox key/tick, or injected per-agent
uniforms for parity runs).  Each function is a thin ctypes call into liblpk.so on
the current CUDA stream; nothing here computes on the host.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lpk
from ._lpk import check, make_rng, ptr, stream_handle

__all__ = [
    "get_deaths", "disease_state_step", "fast_ri", "fast_sia", "tx_step_prep", "tx_node_math", "tx_infect",
    "count_SEIRP", "make_rng", "philox_selftest",
]


def _rng_ref(rng):
    return C.byref(rng) if rng is not None else None


def philox_selftest(ctr: torch.Tensor, key: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(ctr)
    check(_lpk.lib().lpk_philox_selftest(ptr(ctr), ptr(key), ptr(out), C.c_int64(ctr.shape[0]), stream_handle()),
          "lpk_philox_selftest")
    return out


def get_deaths(num_nodes, num_people, disease_state, node_id, date_of_death, t, num_dying):
    """reference model.py:1772; ``num_dying`` int32[num_nodes] is overwritten."""
    check(_lpk.lib().lpk_get_deaths(C.c_int32(num_nodes), C.c_int64(num_people), ptr(disease_state), ptr(node_id),
                                    ptr(date_of_death), C.c_int32(t), ptr(num_dying), stream_handle()), "lpk_get_deaths")


def disease_state_step(node_id, n_nodes, disease_state, strain, active_count, exposure_timer, infection_timer,
                       potentially_paralyzed, paralyzed, ipv_protected, paralysis_timer, p_paralysis, new_potential,
                       new_paralyzed, rng=None):
    """reference model.py:344-359; ``new_potential`` / ``new_paralyzed`` int32[n_nodes] are added to."""
    check(_lpk.lib().lpk_disease_state_step(
        ptr(node_id), C.c_int32(n_nodes), ptr(disease_state), ptr(strain), C.c_int64(active_count), ptr(exposure_timer),
        ptr(infection_timer), ptr(potentially_paralyzed), ptr(paralyzed), ptr(ipv_protected), ptr(paralysis_timer),
        C.c_float(np.float32(p_paralysis)), ptr(new_potential), ptr(new_paralyzed), _rng_ref(rng), stream_handle(),
    ), "lpk_disease_state_step")


def fast_ri(step_size, node_id, disease_state, strain, ipv_protected, ri_timer, sim_t, vx_prob_ri, vx_prob_ipv,
            num_people, ri_counts, ri_protected, ipv_counts, chronically_missed, ri_vaccine_strain, rng=None):
    """reference model.py:1805-1821; the three count tensors are int32[n_nodes], overwritten."""
    check(_lpk.lib().lpk_fast_ri(
        C.c_int64(step_size), ptr(node_id), ptr(disease_state), ptr(strain), ptr(ipv_protected), ptr(ri_timer),
        C.c_int64(sim_t), ptr(vx_prob_ri), ptr(vx_prob_ipv), C.c_int64(num_people), C.c_int32(vx_prob_ri.shape[0]),
        ptr(ri_counts), ptr(ri_protected), ptr(ipv_counts), ptr(chronically_missed), C.c_int8(int(ri_vaccine_strain)),
        _rng_ref(rng), stream_handle(),
    ), "lpk_fast_ri")


def fast_sia(node_ids, disease_states, strain, dobs, sim_t, vx_prob, vx_eff, count, nodes_to_vaccinate, min_age,
             max_age, vaccinated, protected, chronically_missed, sia_vaccine_strain, event_idx=0, rng=None):
    """reference model.py:1995-2011; ``vaccinated`` / ``protected`` int32[n_nodes] are overwritten."""
    check(_lpk.lib().lpk_fast_sia(
        ptr(node_ids), ptr(disease_states), ptr(strain), ptr(dobs), C.c_int64(sim_t), ptr(vx_prob), C.c_double(vx_eff),
        C.c_int64(count), ptr(nodes_to_vaccinate), C.c_int64(min_age), C.c_int64(max_age), C.c_int32(vx_prob.shape[0]),
        ptr(vaccinated), ptr(protected), ptr(chronically_missed), C.c_int8(int(sia_vaccine_strain)),
        C.c_uint32(event_idx), _rng_ref(rng), stream_handle(),
    ), "lpk_fast_sia")


def tx_step_prep(num_nodes, num_people, n_strains, strains, strain_r0_scalars, disease_states, node_ids,
                 daily_infectivity, risks, out=None):
    """reference model.py:932-942.  Returns (beta_fx int64[nodes, strains], exposure_fx int64[nodes], sus int64[nodes],
    risk_hist int32[nodes, RISK_BINS]); the float tallies are exact 2^30 fixed point (divide by ``FX_SCALE``)."""
    dev = disease_states.device
    if out is None:
        out = (torch.empty((num_nodes, n_strains), dtype=torch.int64, device=dev),
               torch.empty(num_nodes, dtype=torch.int64, device=dev), torch.empty(num_nodes, dtype=torch.int64, device=dev),
               torch.empty((num_nodes, _lpk.RISK_BINS), dtype=torch.int32, device=dev))
    beta_fx, exposure_fx, sus, risk_hist = out
    srs = (C.c_double * n_strains)(*[float(v) for v in strain_r0_scalars])
    check(_lpk.lib().lpk_tx_step_prep(
        C.c_int32(num_nodes), C.c_int64(num_people), C.c_int32(n_strains), ptr(strains), srs, ptr(disease_states),
        ptr(node_ids), ptr(daily_infectivity), ptr(risks), ptr(beta_fx), ptr(exposure_fx), ptr(sus), ptr(risk_hist),
        stream_handle(),
    ), "lpk_tx_step_prep")
    return beta_fx, exposure_fx, sus, risk_hist


def tx_node_math(beta_fx, exposure_fx, risk_hist, network, beta_seasonality, r0_scalars, alive_counts, zero_inflation,
                 dispersion, rng=None, out=None):
    """Node-level block of Transmission_ABM.step (reference model.py:1332-1351, 1362-1407) on the device.
    Returns (tau float32[nodes], strain_cdf float64[nodes, strains], prob float64[nodes, strains], expected float64[nodes])."""
    n, ns = beta_fx.shape
    dev = beta_fx.device
    if out is None:
        out = (torch.empty(n, dtype=torch.float32, device=dev), torch.empty((n, ns), dtype=torch.float64, device=dev),
               torch.empty((n, ns), dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.float64, device=dev),
               torch.empty(2 * n, dtype=torch.float64, device=dev))
    tau, cdf, prob, expected, ws = out
    check(_lpk.lib().lpk_tx_node_math(
        C.c_int32(n), C.c_int32(ns), ptr(beta_fx), ptr(exposure_fx), ptr(risk_hist), ptr(network), C.c_double(beta_seasonality),
        ptr(r0_scalars), ptr(alive_counts), C.c_double(zero_inflation), C.c_double(dispersion), ptr(tau), ptr(cdf),
        ptr(prob), ptr(expected), ptr(ws), _rng_ref(rng), stream_handle(),
    ), "lpk_tx_node_math")
    return tau, cdf, prob, expected


def tx_infect(num_nodes, num_people, num_strains, node_ids, strain, disease_state, risks, tau, strain_cdf, rng=None,
              out=None):
    """reference model.py:1011-1024 (per-agent Bernoulli scheme, see include/lpk.h T3); returns n_new int32[nodes, strains]."""
    n_new = out if out is not None else torch.empty((num_nodes, num_strains), dtype=torch.int32, device=disease_state.device)
    check(_lpk.lib().lpk_tx_infect(
        C.c_int32(num_nodes), C.c_int64(num_people), C.c_int32(num_strains), ptr(node_ids), ptr(strain),
        ptr(disease_state), ptr(risks), ptr(tau), ptr(strain_cdf), ptr(n_new), _rng_ref(rng), stream_handle(),
    ), "lpk_tx_infect")
    return n_new


def count_SEIRP(node_id, disease_state, strain, potentially_paralyzed, paralyzed, n_nodes, n_strains, n_people, out=None):
    """reference model.py:869; returns the same 8-tuple (S, E, I, R, E_by_strain, I_by_strain, POTP, P) as int32 tensors."""
    dev = disease_state.device
    if out is None:
        mk = lambda *shape: torch.empty(shape, dtype=torch.int32, device=dev)  # noqa: E731
        out = (mk(n_nodes), mk(n_nodes), mk(n_nodes), mk(n_nodes), mk(n_nodes, n_strains), mk(n_nodes, n_strains),
               mk(n_nodes), mk(n_nodes))
    S, E, I, R, Ebs, Ibs, PP, P = out  # noqa: E741
    check(_lpk.lib().lpk_count_seirp(
        ptr(node_id), ptr(disease_state), ptr(strain), ptr(potentially_paralyzed), ptr(paralyzed), C.c_int32(n_nodes),
        C.c_int32(n_strains), C.c_int64(n_people), ptr(S), ptr(E), ptr(I), ptr(R), ptr(Ebs), ptr(Ibs), ptr(PP), ptr(P),
        stream_handle(),
    ), "lpk_count_seirp")
    return out


# ----------------------------------------------------------------------------- launch accounting / per-kernel timing
class LaunchStats:
    """Counts liblpk kernel launches and, when ``timing`` is on, brackets every call with CUDA events on the
    launching stream so bench.py can report each kernel's average device time inside the timed region."""

    KERNELS_PER_CALL = {"get_deaths": 1, "disease_state_step": 1, "fast_ri": 1, "fast_sia": 1, "tx_step_prep": 1,
                        "tx_node_math": 2, "tx_infect": 1, "count_SEIRP": 2}

    def __init__(self):
        self.reset()
        self.timing = False

    def reset(self):
        self.launches = 0
        self.calls = {}
        self.events = {}
        self.times = {}  # name -> [ms]: launches timed by liblpk itself (lpk_run_days brackets them with its own events)

    def record(self, name, fn, n_kernels):
        """Run ``fn`` (a liblpk launch), counting it and, when timing is on, bracketing it with CUDA events."""
        self.launches += n_kernels  # 0: the call counts its own launches
        self.calls[name] = self.calls.get(name, 0) + 1
        if not self.timing:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.events.setdefault(name, []).append((e0, e1))
        return out

    def summary(self):
        """{name: (calls, mean_ms)}; synchronises the device."""
        torch.cuda.synchronize()
        out = {k: (len(v), sum(a.elapsed_time(b) for a, b in v) / max(len(v), 1)) for k, v in self.events.items()}
        out.update({k: (len(v), sum(v) / max(len(v), 1)) for k, v in self.times.items()})
        return out


STATS = LaunchStats()


def _instrument(fn):
    import functools

    name = fn.__name__
    per_call = LaunchStats.KERNELS_PER_CALL[name]

    @functools.wraps(fn)
    def wrapper(*a, **kw):
        STATS.launches += per_call
        STATS.calls[name] = STATS.calls.get(name, 0) + 1
        if not STATS.timing:
            return fn(*a, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **kw)
        e1.record()
        STATS.events.setdefault(name, []).append((e0, e1))
        return out

    return wrapper


for _name in LaunchStats.KERNELS_PER_CALL:
    globals()[_name] = _instrument(globals()[_name])
