"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (``lp_oracle.c``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(``laser-polio_b200/``) never does.

The per-agent stages live in C (``lp_oracle.c``); the node-level transmission
math (reference ``model.py:1327-1407``) is restated here in numpy because it is
~40 lines of dense array arithmetic whose dtype flow (float32 tallies promoted
by float64 scalars) numpy reproduces exactly.

Arguments mirror the reference's free functions (same names, same order, arrays
mutated in place) with the uniform source appended: ``u_inj`` arrays for
gate-1 runs, else Philox keyed on ``(seed, tick)``.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liblp_oracle.so"

FX_SCALE = float(2**30)
RISK_BINS = 192

STAGE_PARALYSIS, STAGE_RI, STAGE_SIA, STAGE_EXPOSE, STAGE_STRAIN, STAGE_NODE, STAGE_BIRTH, STAGE_LIFESPAN = range(8)
STAGE_NODE_VAR = 9


def build(force: bool = False) -> Path:
    newest = max((_HERE / f).stat().st_mtime for f in ("lp_oracle.c", "lp_oracle_init.c"))
    if force or not _SO.exists() or _SO.stat().st_mtime < newest:
        subprocess.run(["make", "-C", str(_HERE)] + (["-B"] if force else []), check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        # sleeping (not spinning) idle OpenMP workers: on shared / oversubscribed hosts spinning workers steal the
        # cores the working threads need (measured here: 4x slower with the default active policy)
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        _lib = C.CDLL(str(build()))
        _lib.orc_uniform53.restype = C.c_double
        _lib.orc_uniform53.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a, dtype=None):
    """Pointer to a C-contiguous numpy array (dtype-checked), or NULL for None."""
    if a is None:
        return None
    if dtype is not None and a.dtype != np.dtype(dtype):
        raise TypeError(f"expected {np.dtype(dtype)}, got {a.dtype}")
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def uniform53(seed: int, idx: int, tick: int, stage: int, pair: int = 0) -> float:
    return float(lib().orc_uniform53(seed, idx, tick, stage, pair))


# --------------------------------------------------------------------- stages
def get_deaths(num_nodes, num_people, disease_state, node_id, date_of_death, t, num_dying):
    """reference model.py:1767-1781"""
    lib().orc_get_deaths(
        C.c_int32(num_nodes), C.c_int64(num_people), _p(disease_state, np.int8), _p(node_id, np.int16),
        _p(date_of_death, np.int32), C.c_int32(t), _p(num_dying, np.int32),
    )


def disease_state_step(
    node_id, n_nodes, disease_state, strain, active_count, exposure_timer, infection_timer, potentially_paralyzed,
    paralyzed, ipv_protected, paralysis_timer, p_paralysis, new_potential, new_paralyzed, u_inj=None, seed=0, tick=0,
    id_base=0,
):
    """reference model.py:344-454"""
    lib().orc_disease_state_step(
        _p(node_id, np.int16), C.c_int32(n_nodes), _p(disease_state, np.int8), _p(strain, np.int8),
        C.c_int64(active_count), _p(exposure_timer, np.int8), _p(infection_timer, np.int8),
        _p(potentially_paralyzed, np.int8), _p(paralyzed, np.int8), _p(ipv_protected, np.int8),
        _p(paralysis_timer, np.int8), C.c_float(np.float32(p_paralysis)), _p(new_potential, np.int32),
        _p(new_paralyzed, np.int32), _p(u_inj, np.float64), C.c_uint64(seed), C.c_uint32(tick), C.c_uint64(id_base),
    )


def fast_ri(
    step_size, node_id, disease_state, strain, ipv_protected, ri_timer, sim_t, vx_prob_ri, vx_prob_ipv, num_people,
    ri_counts, ri_protected, ipv_counts, chronically_missed, ri_vaccine_strain, u1_inj=None, u2_inj=None, seed=0, tick=0,
    id_base=0,
):
    """reference model.py:1805-1855; the three count outputs are per-node int32 (already thread-reduced)."""
    n_nodes = len(vx_prob_ri)
    lib().orc_fast_ri(
        C.c_int64(step_size), _p(node_id, np.int16), _p(disease_state, np.int8), _p(strain, np.int8),
        _p(ipv_protected, np.int8), _p(ri_timer, np.int16), C.c_int64(sim_t), _p(vx_prob_ri, np.float64),
        _p(vx_prob_ipv, np.float64), C.c_int64(num_people), C.c_int32(n_nodes), _p(ri_counts, np.int32),
        _p(ri_protected, np.int32), _p(ipv_counts, np.int32), _p(chronically_missed, np.uint8),
        C.c_int8(int(ri_vaccine_strain)), _p(u1_inj, np.float64), _p(u2_inj, np.float64), C.c_uint64(seed),
        C.c_uint32(tick), C.c_uint64(id_base),
    )


def fast_sia(
    node_ids, disease_states, strain, dobs, sim_t, vx_prob, vx_eff, count, nodes_to_vaccinate, min_age, max_age,
    vaccinated, protected, chronically_missed, sia_vaccine_strain, u_inj=None, seed=0, tick=0, event_idx=0, id_base=0,
):
    """reference model.py:1995-2060; vaccinated/protected are per-node int32 (already thread-reduced)."""
    n_nodes = len(vx_prob)
    lib().orc_fast_sia(
        _p(node_ids, np.int16), _p(disease_states, np.int8), _p(strain, np.int8), _p(dobs, np.int32),
        C.c_int64(sim_t), _p(vx_prob, np.float32), C.c_double(vx_eff), C.c_int64(count),
        _p(nodes_to_vaccinate, np.uint8), C.c_int64(min_age), C.c_int64(max_age), C.c_int32(n_nodes),
        _p(vaccinated, np.int32), _p(protected, np.int32), _p(chronically_missed, np.uint8),
        C.c_int8(int(sia_vaccine_strain)), _p(u_inj, np.float64), C.c_uint64(seed), C.c_uint32(tick),
        C.c_uint32(event_idx), C.c_uint64(id_base),
    )


def tx_step_prep(num_nodes, num_people, n_strains, strains, strain_r0_scalars, disease_states, node_ids,
                 daily_infectivity, risks, mode="fx"):
    """reference model.py:932-1007.  mode: 'f32' (reference-like), 'f64' (truth), 'fx' (device fixed point).

    Returns (beta[nodes,strains] f64, exposure[nodes] f64, sus[nodes] i64, beta_fx i64, exposure_fx i64); the histogram of
    the susceptibles' risks (int32[nodes, RISK_BINS], what the device's node step solves tau on) is kept in
    ``tx_step_prep.last_hist``.
    """
    m = {"f32": 0, "f64": 1, "fx": 2}[mode]
    beta = np.zeros((num_nodes, n_strains), np.float64)
    expo = np.zeros(num_nodes, np.float64)
    sus = np.zeros(num_nodes, np.int64)
    beta_fx = np.zeros((num_nodes, n_strains), np.int64)
    expo_fx = np.zeros(num_nodes, np.int64)
    hist = np.zeros((num_nodes, RISK_BINS), np.int32)
    srs = np.ascontiguousarray(strain_r0_scalars, dtype=np.float64)
    lib().orc_tx_step_prep(
        C.c_int32(num_nodes), C.c_int64(num_people), C.c_int32(n_strains), _p(strains, np.int8), _p(srs),
        _p(disease_states, np.int8), _p(node_ids, np.int16), _p(daily_infectivity, np.float32),
        _p(risks, np.float32), C.c_int(m), _p(beta), _p(expo), _p(sus), _p(beta_fx), _p(expo_fx), _p(hist),
    )
    tx_step_prep.last_hist = hist
    return beta, expo, sus, beta_fx, expo_fx


def count_SEIRP(node_id, disease_state, strain, potentially_paralyzed, paralyzed, n_nodes, n_strains, n_people):
    """reference model.py:869-929; same 8-tuple as the reference."""
    S = np.zeros(n_nodes, np.int32); E = np.zeros(n_nodes, np.int32); I = np.zeros(n_nodes, np.int32)  # noqa: E702
    R = np.zeros(n_nodes, np.int32); PP = np.zeros(n_nodes, np.int32); P = np.zeros(n_nodes, np.int32)  # noqa: E702
    Ebs = np.zeros((n_nodes, n_strains), np.int32); Ibs = np.zeros((n_nodes, n_strains), np.int32)  # noqa: E702
    lib().orc_count_seirp(
        _p(node_id, np.int16), _p(disease_state, np.int8), _p(strain, np.int8), _p(potentially_paralyzed, np.int8),
        _p(paralyzed, np.int8), C.c_int32(n_nodes), C.c_int32(n_strains), C.c_int64(n_people), _p(S), _p(E), _p(I),
        _p(R), _p(Ebs), _p(Ibs), _p(PP), _p(P),
    )
    return S, E, I, R, Ebs, Ibs, PP, P


def tx_infect_ref(num_nodes, num_people, num_strains, sus_by_node, node_ids, strain, disease_state, sus_indices,
                  sus_probs, risks, prob_exp_by_node_strain, n_exposures_to_create_by_node_strain, seed=0, tick=0):
    """reference model.py:1010-1149 (distributional restatement, own RNG stream)."""
    n_new = np.zeros((num_nodes, num_strains), np.int32)
    sus64 = np.ascontiguousarray(sus_by_node, dtype=np.int64)
    pe = np.ascontiguousarray(prob_exp_by_node_strain, dtype=np.float64)
    nc = np.ascontiguousarray(n_exposures_to_create_by_node_strain, dtype=np.int32)
    lib().orc_tx_infect_ref(
        C.c_int32(num_nodes), C.c_int64(num_people), C.c_int32(num_strains), _p(sus64), _p(node_ids, np.int16),
        _p(strain, np.int8), _p(disease_state, np.int8), _p(sus_indices, np.int32), _p(sus_probs, np.float32),
        _p(risks, np.float32), _p(pe), _p(nc), _p(n_new), C.c_uint64(seed), C.c_uint32(tick),
    )
    return n_new


def tx_infect_bernoulli(num_nodes, num_people, num_strains, node_ids, strain, disease_state, risks, q, strain_cdf,
                        x_inj=None, u_strain_inj=None, seed=0, tick=0, id_base=0):
    """Device exposure scheme: susceptible i of node n is exposed w.p. 1 - exp(-risk_i * q[n]), q = tau (see lp_oracle.c)."""
    n_new = np.zeros((num_nodes, num_strains), np.int32)
    lib().orc_tx_infect_bernoulli(
        C.c_int32(num_nodes), C.c_int64(num_people), C.c_int32(num_strains), _p(node_ids, np.int16),
        _p(strain, np.int8), _p(disease_state, np.int8), _p(risks, np.float32), _p(q, np.float32),
        _p(strain_cdf, np.float64), _p(n_new), _p(x_inj, np.uint32), _p(u_strain_inj, np.float64),
        C.c_uint64(seed), C.c_uint32(tick), C.c_uint64(id_base),
    )
    return n_new


# ------------------------------------------------------- node-level math (T2)
def tx_foi(beta_by_node_strain, network, beta_seasonality, r0_scalars, alive_counts):
    """reference model.py:1332-1351: network transfer, seasonality x r0_scalars, rate -> probability.

    ``beta_by_node_strain`` keeps whatever dtype the caller passes (float32 in the
    reference, so the in-place ``+=`` of the transfer rounds to float32 there).
    Returns (beta_pre copy, prob_exp_by_node_strain float64).
    """
    beta = np.array(beta_by_node_strain, copy=True)
    beta_pre = beta.copy()
    for s in range(beta.shape[1]):
        transfer = (beta[:, s] * network.T).T
        beta[:, s] += transfer.sum(axis=0) - transfer.sum(axis=1)
    beta = beta * beta_seasonality * np.asarray(r0_scalars)[:, np.newaxis]
    rate = beta / np.maximum(np.asarray(alive_counts)[:, np.newaxis], 1)
    prob = np.maximum(1 - np.exp(-rate), 0)
    return beta_pre, prob


def tx_draw_counts_ref(beta_pre, prob, exposure_by_node, zero_inflation, dispersion, rs=np.random):
    """reference model.py:1362-1407: Poisson / zero-inflated NB count per node + multinomial by strain.

    ``rs`` is ``numpy.random`` (global legacy stream, like the reference) or a RandomState.
    """
    n_nodes, n_strains = prob.shape
    total_prob = prob.sum(axis=1)
    expected = exposure_by_node * total_prob
    out = np.zeros_like(prob, dtype=np.int32)
    for n in range(n_nodes):
        e = expected[n]
        if e < 0:
            e = 0
        if e == 0:
            k = 0
        elif np.sum(beta_pre[n]) == 0:
            if zero_inflation >= 1.0:
                k = 0
            else:
                mean = e / (1 - zero_inflation)
                r = max(1, int(np.round(dispersion)))
                p = r / (r + mean)
                k = 0 if rs.rand() < zero_inflation else rs.negative_binomial(r, p)
        else:
            k = rs.poisson(e)
        if k > 0:
            sp = prob[n] / total_prob[n] if total_prob[n] > 0 else np.zeros(n_strains)
            out[n] = rs.multinomial(k, sp)
    return out, expected


def seasonality(doy: int, days_in_year: int, amplitude: float, peak_doy: float) -> float:
    """reference utils.py:616-625"""
    return 1 + amplitude * np.cos(2 * np.pi * (doy - peak_doy) / days_in_year)


def risk_bin_weights():
    b = np.arange(RISK_BINS)
    return np.ldexp(1.0 + ((b & 7) + 0.5) / 8.0, (b >> 3) - 12)


def poisson_min_moments(e, S):
    """(E[min(K, S)], P(K < S), Var[min(K, S)], P(K = S - 1)) for K ~ Poisson(e), S a positive integer: the device's
    windowed sums (lpk_kernels.cu, poisson_min_moments) in float64."""
    sd = np.sqrt(e)
    if S > e + 12.0 * sd + 12.0:
        return e, 1.0, e, 0.0
    if S < e - 12.0 * sd - 12.0:
        return float(S), 0.0, 0.0, 0.0
    from scipy.special import gammaln

    k0 = max(int(np.floor(e - 12.0 * sd - 12.0)), 0)
    k = np.arange(k0, int(S), dtype=np.float64)
    pmf = np.exp(k * np.log(e) - e - gammaln(k + 1.0))
    d = S - k
    b0, b1, b2 = pmf.sum(), (d * pmf).sum(), (d * d * pmf).sum()
    return S - b1, min(b0, 1.0), max(b2 - b1 * b1, 0.0), float(pmf[-1])


def _node_u_var(seed, node, k, tick, pair):
    x = philox4x32_10([node, k, tick, STAGE_NODE_VAR], [seed & 0xFFFFFFFF, seed >> 32])
    hi, lo = (int(x[2]), int(x[3])) if pair else (int(x[0]), int(x[1]))
    return float((((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0))


def unit_gamma(seed, node, tick, shape):
    """Gamma(shape, 1 / shape) on the node's NODE_VAR stream; mirrors the device function draw for draw."""
    a = shape + 1.0 if shape < 1.0 else shape
    d = a - 1.0 / 3.0
    c = 1.0 / np.sqrt(9.0 * d)
    k = 1
    while True:
        u1, u2 = _node_u_var(seed, node, k, tick, 0), _node_u_var(seed, node, k, tick, 1)
        x = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
        k += 1
        v = 1.0 + c * x
        if v <= 0:
            continue
        v = v * v * v
        u = _node_u_var(seed, node, k, tick, 0)
        k += 1
        if np.log(1.0 - u) < 0.5 * x * x + d - d * v + d * np.log(v):
            g = d * v
            break
    if shape < 1.0:
        g *= (1.0 - _node_u_var(seed, node, 0, tick, 0)) ** (1.0 / shape)
    return g / shape


def _tau_lower_bound(S, Wsum, T):
    """the equal-weights solution: a lower bound of the root by Jensen"""
    return -np.log1p(-T / S) * S / Wsum


def _newton_tau(h, w, T, t, tol):
    """tau with sum_b h[b] (1 - exp(-w[b] tau)) = T by Newton from t at or left of the root (the function is concave
    increasing, so the iteration is monotone)."""
    for _ in range(100):
        em = np.expm1(-w * t)
        F, dF = -(h * em).sum(), (h * w * (em + 1.0)).sum()
        step = (T - F) / dF
        if not step > 0.0:
            break
        t += step
        if step <= tol * t:
            break
    return t


def solve_tau(hist_row, E, expo=None, seed=0, node=0, tick=0):
    """The device's node scale (lpk_kernels.cu, solve_tau_warp).  Bin centres are rescaled to the exact risk sum ``expo``.
    tau0 solves sum_b hist[b] (1 - exp(-w_b tau)) = T0 = E[min(K, S)], K ~ Poisson(E); the independent trials at tau0 have
    variance V_own = sum_b hist[b] p_b (1 - p_b); what is missing against Var[min(K, S)] is supplied by a unit-mean gamma
    multiplier g2 on E with CV^2 = (Var[min(K, S)] - V_own) / (P(K < S) E)^2 (delta method), at most 1 / (S_eff - 1) (its value
    far from saturation, where the delta method is exact), faded by P(K < S)^2; then tau solves for T = E[min(Poisson(E g2), S)]."""
    h = np.asarray(hist_row, np.float64)
    w = risk_bin_weights()
    S, Wsum = h.sum(), (h * w).sum()
    if not (E > 0.0) or S <= 0 or not (Wsum > 0.0):
        return 0.0
    c = expo / Wsum if (expo is not None and expo > 0.0) else 1.0
    w = w * c
    Wsum, W2 = (h * w).sum(), (h * w * w).sum()
    Seff = Wsum * Wsum / W2
    T, pless, V, plast = poisson_min_moments(E, S)
    if T >= S * (1.0 - 1e-12):
        return 3.0e38
    if not (T > 0.0):
        return 0.0
    t = _newton_tau(h, w, T, _tau_lower_bound(S, Wsum, T), 1e-7)
    redo = False
    slope = pless * E
    if Seff > 1.0 + 1e-9 and slope > 0.0:
        pb = -np.expm1(-w * t)
        # the top-up fades with P(K < S)^2 as the node saturates: there the response T(E) is strongly concave, the delta
        # method no longer holds and the mean matters more than the variance; what concavity remains is compensated to second
        # order (T'' = -P(K = S - 1)) so that the mean over the multiplier stays T(E)
        v = min((V - (h * pb * (1.0 - pb)).sum()) / (slope * slope), 1.0 / (Seff - 1.0)) * pless * pless
        if v > 1e-9:
            E2 = E * (1.0 + 0.5 * plast * E * v / pless) * unit_gamma(seed, node, tick, 1.0 / v)
            T = poisson_min_moments(E2, S)[0] if E2 > 0.0 else 0.0
            if T >= S * (1.0 - 1e-12):
                return 3.0e38
            if not (T > 0.0):
                return 0.0
            redo = True
    t = _newton_tau(h, w, T, _tau_lower_bound(S, Wsum, T) if redo else t, 1e-11)
    return min(t, 3.0e38)


def tx_node_math_device(beta_fx, exposure_fx, risk_hist, network, beta_seasonality, r0_scalars, alive_counts, zero_inflation,
                        dispersion, seed, tick):
    """float64 restatement of the DEVICE node step (lpk_tx_node_math): the reference's formulae up to the probability
    and the expected exposures per node (model.py:1332-1351, 1362-1363); then, instead of an integer count draw, the
    node's exposure scale tau[n] with  sum_{i in S_n} (1 - exp(-w_i tau[n])) = T_n  on the risk histogram, T_n the mean of
    the reference's count min(K, S_n), K ~ Poisson(expected[n] * g_n * g2_n); g_n = 1 when the node has local infectivity
    (sum_s beta_pre > 0), else 0 w.p. zi, else Gamma(r, 1/r) / (1 - zi) (model.py:1381-1393's ZINB as a zero-inflated
    gamma-Poisson mixture), r = max(1, round(dispersion)); g2_n the unit-mean gamma that tops the variance of independent
    trials up to the reference's (include/lpk.h, T2; solve_tau).

    Returns (tau float32[nodes], strain_cdf float64[nodes, strains], prob float64[nodes, strains], expected[nodes]).
    """
    beta_pre = np.asarray(beta_fx, np.float64) / FX_SCALE
    exposure = np.asarray(exposure_fx, np.float64) / FX_SCALE
    W = np.asarray(network, np.float64)
    beta = beta_pre + W.T @ beta_pre - beta_pre * W.sum(axis=1)[:, None]
    beta = beta * float(beta_seasonality) * np.asarray(r0_scalars, np.float64)[:, None]
    rate = beta / np.maximum(np.asarray(alive_counts, np.float64)[:, None], 1.0)
    prob = np.maximum(1.0 - np.exp(-rate), 0.0)
    P = np.zeros(prob.shape[0])
    cdf = np.zeros_like(prob)
    for s in range(prob.shape[1]):  # sequential sum, same association order as the device
        P = P + prob[:, s]
    run = np.zeros(prob.shape[0])
    with np.errstate(divide="ignore", invalid="ignore"):
        for s in range(prob.shape[1]):
            run = run + np.where(P > 0, prob[:, s] / P, 0.0)
            cdf[:, s] = run
    g = np.ones(prob.shape[0])
    local = np.zeros(prob.shape[0])
    for s in range(prob.shape[1]):
        local = local + beta_pre[:, s]
    r = max(1, int(np.round(dispersion)))
    for n in np.nonzero((local == 0) & (P > 0))[0]:
        g[n] = node_importation_multiplier(int(seed), int(n), int(tick), float(zero_inflation), r)
    expected = exposure * P
    tau = np.array([solve_tau(risk_hist[n], expected[n] * g[n], exposure[n], int(seed), n, int(tick)) for n in range(prob.shape[0])],
                   np.float64).astype(np.float32)
    return tau, cdf, prob, expected


def _node_u(seed, node, k, tick, pair):
    x = philox4x32_10([node, k, tick, STAGE_NODE], [seed & 0xFFFFFFFF, seed >> 32])
    hi, lo = (int(x[2]), int(x[3])) if pair else (int(x[0]), int(x[1]))
    return float((((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0))


def node_importation_multiplier(seed, node, tick, zi, r):
    """g_n for an importation-only node; mirrors lpk's device function draw for draw."""
    if zi >= 1.0:
        return 0.0
    if _node_u(seed, node, 0, tick, 0) < zi:
        return 0.0
    # Marsaglia-Tsang gamma(shape=r >= 1, scale=1); normals by Box-Muller on counter k = 1, 2, ...
    d = r - 1.0 / 3.0
    c = 1.0 / np.sqrt(9.0 * d)
    k = 1
    while True:
        u1 = _node_u(seed, node, k, tick, 0)
        u2 = _node_u(seed, node, k, tick, 1)
        x = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
        k += 1
        v = 1.0 + c * x
        if v <= 0:
            continue
        v = v * v * v
        u = _node_u(seed, node, k, tick, 0)
        k += 1
        if np.log(1.0 - u) < 0.5 * x * x + d - d * v + d * np.log(v):
            return (d * v) / r / (1.0 - zi)


# ------------------------------------------------------------------- births (device scheme, include/lpk.h V2)
def _u53_pair(x, pair):
    hi, lo = (int(x[2]), int(x[3])) if pair else (int(x[0]), int(x[1]))
    return float((((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0))


def vd_births_device(pop_prev, birth_rate, step_size, cum_deaths, count, capacity, seed, tick, id_base=0, max_year=100):
    """Restatement of lpk_vd_births (reference model.py:1711-1734 with Philox draws instead of the host numpy stream).

    Returns (births[nodes] int32, node_id[total] int16, date_of_death[total] int32, new_count, status); the cohort
    occupies slots [count, count + total) node-major, date_of_birth = tick, disease_state = 0.
    """
    key = [seed & 0xFFFFFFFF, seed >> 32]
    n = len(pop_prev)
    births = np.zeros(n, np.int32)
    for node in range(n):
        expected = float(step_size) * float(birth_rate[node]) * float(pop_prev[node])
        whole = int(expected)
        x = philox4x32_10([node, 0, tick, STAGE_BIRTH], key)
        births[node] = max(whole + (1 if _u53_pair(x, 0) < expected - whole else 0), 0)
    total = int(births.sum())
    if count + total > capacity:
        return np.zeros(n, np.int32), np.zeros(0, np.int16), np.zeros(0, np.int32), count, 1
    cd = np.asarray(cum_deaths, np.int64)
    tot_deaths = max(int(cd[max_year + 1]), 1)
    node_id = np.repeat(np.arange(n, dtype=np.int16), births)
    dod = np.zeros(total, np.int32)
    for k in range(total):
        g = count + k + id_base
        x = philox4x32_10([g & 0xFFFFFFFF, g >> 32, tick, STAGE_LIFESPAN], key)
        draw = 1 + int(np.floor(_u53_pair(x, 0) * tot_deaths))
        yod = int(np.searchsorted(cd[: max_year + 2], draw, side="left")) - 1
        yod = min(max(yod, 0), max_year)
        u2 = _u53_pair(x, 1)
        doy = 1 + int(np.floor(u2 * 364.0)) if yod == 0 else int(np.floor(u2 * 365.0))
        dod[k] = tick + yod * 365 + doy
    return births, node_id, dod, count + total, 0


# ------------------------------------------------------------------------------------------ population initialisers
# (lp_oracle_init.c; checker for laser-polio_b200/csrc/lpk_init.cu)
DIST_KINDS = {"constant": 0, "exponential": 1, "gamma": 2, "lognormal": 3, "normal": 4, "poisson": 5, "uniform": 6}


class Dist(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_double), ("b", C.c_double)]


def dist(kind: str, a: float = 0.0, b: float = 0.0) -> Dist:
    """kind / a / b as in include/lpk.h (lognormal takes mu, sigma of the underlying normal)."""
    return Dist(DIST_KINDS[kind], float(a), float(b))


def lognormal_mu_sigma(mean: float, sigma: float):
    """lp.lognormal(mean, sigma) -> parameters of the underlying normal (reference distributions.py:73-87)."""
    return float(np.log(mean**2 / np.sqrt(sigma**2 + mean**2))), float(np.sqrt(np.log(sigma**2 / mean**2 + 1)))


def init_draw(n: int, d: Dist, seed: int, stage: int = 17) -> np.ndarray:
    out = np.zeros(n, np.float64)
    lib().orc_init_draw(C.c_int64(n), C.byref(d), C.c_uint64(seed), C.c_uint32(stage), _p(out))
    return out


def init_heterogeneity(start, end, acq_risk_out, infectivity_out, mu_ln, sigma_ln, scale_gamma, rho, heterogeneity, mean_gamma,
                       seed, id_base=0):
    lib().orc_init_heterogeneity(C.c_int64(start), C.c_int64(end), _p(acq_risk_out, np.float32), _p(infectivity_out, np.float32),
                                 C.c_double(mu_ln), C.c_double(sigma_ln), C.c_double(scale_gamma), C.c_double(rho),
                                 C.c_int32(int(bool(heterogeneity))), C.c_double(mean_gamma), C.c_uint64(seed), C.c_uint64(id_base))


def init_timers(start, end, exposure_timer, infection_timer, paralysis_timer, dur_exp: Dist, dur_inf: Dist, t_to_paralysis: Dist,
                seed, id_base=0):
    lib().orc_init_timers(C.c_int64(start), C.c_int64(end), _p(exposure_timer, np.int8), _p(infection_timer, np.int8),
                          _p(paralysis_timer, np.int8), C.byref(dur_exp), C.byref(dur_inf), C.byref(t_to_paralysis),
                          C.c_uint64(seed), C.c_uint64(id_base))


def init_demography(start, end, date_of_birth, date_of_death, ri_timer, bin_cdf, bin_lo, bin_hi, cum_deaths, max_year, seed, id_base=0):
    lib().orc_init_demography(C.c_int64(start), C.c_int64(end), _p(date_of_birth, np.int32), _p(date_of_death, np.int32),
                              _p(ri_timer, np.int16), _p(bin_cdf, np.float64), _p(bin_lo, np.int32), _p(bin_hi, np.int32),
                              C.c_int32(len(bin_cdf)), _p(cum_deaths, np.int64), C.c_int32(max_year), C.c_uint64(seed),
                              C.c_uint64(id_base))


def init_missed(n, n_missed, chronically_missed, seed, id_base=0):
    lib().orc_init_missed(C.c_int64(n), C.c_int64(n_missed), _p(chronically_missed, np.uint8), C.c_uint64(seed), C.c_uint64(id_base))


# ------------------------------------------------------------------------------------------ network construction
# (checker for laser-polio_b200/csrc/lpk_net.cu; reference model.py:1216-1258.  laser-core ~=0.6 is not in the checkout:
# gravity follows the in-tree evidence scripts/sandbox/debug_negative_network.py:74, row_normalizer debug_row_normalizer.py:30-33,
# radiation the published model -- parity unpinned for radiation and distance.)
def net_haversine(lat, lon, epsilon=1.0):
    """The reference's double loop (model.py:1231-1240) with laser-core's Haversine ``distance`` (km, R = 6371)."""
    lat, lon = np.radians(np.asarray(lat, np.float64)), np.radians(np.asarray(lon, np.float64))
    n = len(lat)
    out = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            a = np.sin((lat[j] - lat[i]) / 2) ** 2 + np.cos(lat[i]) * np.cos(lat[j]) * np.sin((lon[j] - lon[i]) / 2) ** 2
            d = 6371.0 * 2 * np.arcsin(np.sqrt(a))
            if d == 0:
                d = epsilon
            out[i, j] = out[j, i] = d
    return out


def net_gravity(pops, dist, k, a, b, c, norm=1.0):
    pops, dist = np.asarray(pops, np.float64), np.asarray(dist, np.float64)
    n = len(pops)
    out = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                out[i, j] = k * pops[i] ** a * pops[j] ** b / dist[i, j] ** c / norm
    return out


def net_radiation(pops, dist, k, include_home=False):
    """From the definition: s_ij = population at distance <= d_ij from i, without j, without i unless include_home."""
    pops, dist = np.asarray(pops, np.float64), np.asarray(dist, np.float64)
    n = len(pops)
    out = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            inside = dist[i] <= dist[i, j]
            s = pops[inside].sum() - pops[j] - (0.0 if include_home else pops[i])
            s = max(s, 0.0)
            out[i, j] = k * pops[i] * pops[j] / ((pops[i] + s) * (pops[i] + pops[j] + s))
    return out


def net_row_normalize(net, max_rowsum):
    net = np.array(net, np.float64, copy=True)
    for i in range(len(net)):
        rs = net[i].sum()
        if rs > max_rowsum:
            net[i] *= max_rowsum / rs
    return net
