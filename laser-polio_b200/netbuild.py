"""The infection-migration network built in HBM.

Device form of reference ``Transmission_ABM._initialize_common`` (model.py:1216-1258): all-pairs Haversine distances (a Python
double loop in the reference), gravity or radiation model, row normalisation -- four kernels of ``csrc/lpk_net.cu`` over a
float64 ``[nodes, nodes]`` CUDA tensor.  Function names and argument meanings follow laser-core's ``migration`` module
(``distance`` / ``gravity`` / ``radiation`` / ``row_normalizer``) as the reference calls them.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lpk

__all__ = ["distance_matrix", "gravity", "radiation", "row_normalizer", "build_network"]


def _f64(x, device):
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(device) if not isinstance(x, torch.Tensor) else x


def distance_matrix(lats, lons, device="cuda", epsilon=1.0):
    """All-pairs Haversine distances in km; coincident nodes get ``epsilon`` (reference model.py:1224-1240)."""
    lat, lon = _f64(lats, device), _f64(lons, device)
    n = lat.numel()
    out = torch.empty((n, n), dtype=torch.float64, device=lat.device)
    _lpk.check(_lpk.lib().lpk_net_haversine(_lpk.ptr(lat), _lpk.ptr(lon), C.c_int32(n), C.c_double(epsilon), _lpk.ptr(out),
                                             _lpk.stream_handle()), "lpk_net_haversine")
    return out


def gravity(pops, distances, k, a, b, c, norm=1.0):
    """``k * p_i^a * p_j^b / d_ij^c / norm``, zero diagonal (laser-core ``gravity``; the reference divides by ``sum(p)^c``)."""
    d = _f64(distances, "cuda")
    p = _f64(pops, d.device)
    n = p.numel()
    out = torch.empty((n, n), dtype=torch.float64, device=d.device)
    _lpk.check(_lpk.lib().lpk_net_gravity(_lpk.ptr(p), _lpk.ptr(d), C.c_int32(n), C.c_double(k), C.c_double(a), C.c_double(b),
                                           C.c_double(c), C.c_double(norm), _lpk.ptr(out), _lpk.stream_handle()), "lpk_net_gravity")
    return out


def radiation(pops, distances, k, include_home=False):
    """Radiation model (laser-core ``radiation``), at most 8192 nodes."""
    d = _f64(distances, "cuda")
    p = _f64(pops, d.device)
    n = p.numel()
    out = torch.empty((n, n), dtype=torch.float64, device=d.device)
    _lpk.check(_lpk.lib().lpk_net_radiation(_lpk.ptr(p), _lpk.ptr(d), C.c_int32(n), C.c_double(k), C.c_int32(int(bool(include_home))),
                                             _lpk.ptr(out), _lpk.stream_handle()), "lpk_net_radiation")
    return out


def row_normalizer(network, max_rowsum):
    """Rows whose sum exceeds ``max_rowsum`` are rescaled to it (laser-core ``row_normalizer``); returns a new tensor."""
    net = _f64(network, "cuda").clone()
    _lpk.check(_lpk.lib().lpk_net_row_normalize(_lpk.ptr(net), C.c_int32(net.shape[0]), C.c_double(max_rowsum), _lpk.stream_handle()),
               "lpk_net_row_normalize")
    return net


def build_network(pars, init_pops, device="cuda"):
    """The whole of reference model.py:1219-1258 on the device; returns the float64 ``[nodes, nodes]`` CUDA tensor."""
    pops = np.asarray(init_pops, dtype=np.float64)
    if pars.distances is not None:
        dist = _f64(pars.distances, device)
    else:
        ids = sorted(pars.node_lookup.keys())
        dist = distance_matrix([pars.node_lookup[i]["lat"] for i in ids], [pars.node_lookup[i]["lon"] for i in ids], device)
    method = pars.migration_method.lower()
    if method == "gravity":
        k = pars.gravity_k * 10 ** pars.gravity_k_exponent
        net = gravity(pops, dist, k, pars.gravity_a, pars.gravity_b, pars.gravity_c, norm=float(np.power(pops.sum(), pars.gravity_c)))
    elif method == "radiation":
        net = radiation(pops, dist, 10 ** pars.radiation_k_log10, include_home=False)
    else:
        raise ValueError(f"Unknown migration method: {pars.migration_method}")
    return row_normalizer(net, pars.max_migr_frac)
