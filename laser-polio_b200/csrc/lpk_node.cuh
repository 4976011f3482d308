// lpk_node.cuh -- the node-level epilogue of a fused tick (bookkeeping of one node), shared by k_tick_epilogue (lpk_tick.cu)
// and k_tx_node_math (lpk_kernels.cu), which runs it inline from the second tick of a run on.
#pragma once
#include "lpk_common.cuh"

// one thread per node n: population row, paralysis census, S / E / I / R rows of tick t-1 from the carried counts, snapshots.
// Every input is loaded BEFORE the first store: the pointers may alias as far as the compiler knows, so loads interleaved
// with stores became a chain of ~15 dependent L2 round trips on the critical path of the node step.
__device__ __forceinline__ void epilogue_node(const lpk_node_args &a, int n, int n_lo) {
    const int ns = a.n_strains;
    const bool pending = (a.flags & LPK_F_PENDING) != 0, vd_day = (a.flags & LPK_F_DEATHS) != 0;
    int d = 0, dpp = 0, dpar = 0, births = 0, pop_prev = 0, potp = 0, par = 0, npot = 0, npar = 0;
    int s_snap = 0, hits = 0, r_snap = 0, r_prev = 0, r_cur = 0;
    long long sus = 0, cnt1 = 0;
    int e_snap[LPK_MAX_STRAINS], i_snap[LPK_MAX_STRAINS], hits_s[LPK_MAX_STRAINS], e_cur[LPK_MAX_STRAINS], i_cur[LPK_MAX_STRAINS];
    if (n == n_lo && a.counts) cnt1 = a.counts[1];
    if (a.deaths) { d = a.deaths[n]; dpp = a.dead_pp[n]; dpar = a.dead_par[n]; }
    if (a.pop) { births = a.births_row ? a.births_row[n] : 0; pop_prev = a.pop_prev[n]; }
    if (a.cur_potp) { potp = a.cur_potp[n]; par = a.cur_p[n]; npot = a.new_potential[n]; npar = a.new_paralyzed[n]; }
    if (a.S_snap) {
        s_snap = a.S_snap[n]; hits = a.tx_hits[n]; r_snap = a.R_snap[n]; sus = a.sus[n]; r_cur = a.R_cur[n];
        if (pending) r_prev = a.R_prev[n];
    }
#pragma unroll
    for (int s = 0; s < LPK_MAX_STRAINS; ++s) {
        e_snap[s] = i_snap[s] = hits_s[s] = e_cur[s] = i_cur[s] = 0;
        if (s < ns) {
            const int64_t c = (int64_t)n * ns + s;
            e_snap[s] = a.E_snap[c]; i_snap[s] = a.I_snap[c]; hits_s[s] = a.tx_hits_by_strain[c];
            e_cur[s] = a.E_cur[c]; i_cur[s] = a.I_cur[c];
        }
    }
    // ---- stores
    if (n == n_lo && a.counts) a.counts[0] = cnt1;
    if (a.deaths) { a.deaths[n] = 0; a.dead_pp[n] = 0; a.dead_par[n] = 0; }
    if (a.pop) {
        if (vd_day) {
            a.deaths_row[n] = d;  // "=": overwrites pre-modelled deaths of the non-agent immunes (model.py:1749)
            a.pop[n] = pop_prev + births - d;
        } else {
            a.pop[n] = pop_prev;
        }
    }
    if (a.cur_potp) {
        potp += npot - dpp;
        par += npar - dpar;
        a.cur_potp[n] = potp; a.cur_p[n] = par;
        a.potp_row[n] = potp; a.p_row[n] = par;
    }
    if (a.S_snap) {
        if (pending) {
            a.S_prev[n] = s_snap - hits;   // "=" (model.py:1476)
            a.R_prev[n] = r_prev + r_snap;  // "+=" on top of the pre-seeded immunes (model.py:1481)
        }
        a.tx_hits[n] = 0;
        a.S_snap[n] = (int32_t)sus;
        a.R_snap[n] = r_cur;
    }
    // exposed / infectious census of tick t-1 from the carried counts: the snapshot taken when tick t-1's stages ended,
    // plus tick t-1's exposures (found by this pass); "=" like Transmission_ABM.log (model.py:1477-1480)
    int e_tot = 0, i_tot = 0, cases = 0;
#pragma unroll
    for (int s = 0; s < LPK_MAX_STRAINS; ++s) {
        if (s < ns) {
            const int64_t c = (int64_t)n * ns + s;
            if (pending) {
                const int es = e_snap[s] + hits_s[s];
                a.E_by_strain_prev[c] = es; a.I_by_strain_prev[c] = i_snap[s];
                e_tot += es; i_tot += i_snap[s];
            }
            a.tx_hits_by_strain[c] = 0;
            a.E_snap[c] = e_cur[s];
            a.I_snap[c] = i_cur[s];
            cases |= e_cur[s] | i_cur[s];
        }
    }
    if (pending) { a.E_prev[n] = e_tot; a.I_prev[n] = i_tot; }
    if (cases && a.any_cases) *a.any_cases = 1;
}
