// lpk_node.cuh -- the node-level epilogue of a fused tick (bookkeeping of one node), shared by k_tick_epilogue (lpk_tick.cu)
// and k_tx_node_math (lpk_kernels.cu), which runs it inline from the second tick of a run on.
#pragma once
#include "lpk_common.cuh"

// one thread per node n: population row, paralysis census, S / E / I / R rows of tick t-1 from the carried counts, snapshots
__device__ __forceinline__ void epilogue_node(const lpk_node_args &a, int n, int n_lo) {
    if (n == n_lo && a.counts) a.counts[0] = a.counts[1];
    const int ns = a.n_strains;
    int d = 0, dpp = 0, dpar = 0;
    if (a.deaths) {
        d = a.deaths[n]; dpp = a.dead_pp[n]; dpar = a.dead_par[n];
        a.deaths[n] = 0; a.dead_pp[n] = 0; a.dead_par[n] = 0;
    }
    if (a.pop) {
        const int births = a.births_row ? a.births_row[n] : 0;
        if (a.flags & LPK_F_DEATHS) {
            a.deaths_row[n] = d;  // "=": overwrites pre-modelled deaths of the non-agent immunes (model.py:1749)
            a.pop[n] = a.pop_prev[n] + births - d;
        } else {
            a.pop[n] = a.pop_prev[n];
        }
    }
    if (a.cur_potp) {
        const int potp = a.cur_potp[n] + a.new_potential[n] - dpp;
        const int par = a.cur_p[n] + a.new_paralyzed[n] - dpar;
        a.cur_potp[n] = potp; a.cur_p[n] = par;
        a.potp_row[n] = potp; a.p_row[n] = par;
    }
    if (a.S_snap) {
        if (a.flags & LPK_F_PENDING) {
            a.S_prev[n] = a.S_snap[n] - a.tx_hits[n];  // "=" (model.py:1476)
            a.R_prev[n] += a.R_snap[n];                // "+=" on top of the pre-seeded immunes (model.py:1481)
        }
        a.tx_hits[n] = 0;
        a.S_snap[n] = (int32_t)a.sus[n];
        a.R_snap[n] = a.R_cur[n];
    }
    // exposed / infectious census of tick t-1 from the carried counts: the snapshot taken when tick t-1's stages ended,
    // plus tick t-1's exposures (found by this pass); "=" like Transmission_ABM.log (model.py:1477-1480)
    if (a.flags & LPK_F_PENDING) {
        int e = 0, i = 0;
        for (int s = 0; s < ns; ++s) {
            const int64_t c = (int64_t)n * ns + s;
            const int es = a.E_snap[c] + a.tx_hits_by_strain[c], is = a.I_snap[c];
            a.E_by_strain_prev[c] = es; a.I_by_strain_prev[c] = is;
            e += es; i += is;
        }
        a.E_prev[n] = e; a.I_prev[n] = i;
    }
    int cases = 0;
    for (int s = 0; s < ns; ++s) {
        const int64_t c = (int64_t)n * ns + s;
        a.tx_hits_by_strain[c] = 0;
        const int e = a.E_cur[c], i = a.I_cur[c];
        a.E_snap[c] = e;
        a.I_snap[c] = i;
        cases |= e | i;
    }
    if (cases && a.any_cases) *a.any_cases = 1;
}
