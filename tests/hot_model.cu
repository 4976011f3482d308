// tests/hot_model.cu -- TEST INFRASTRUCTURE: the per-agent logic of the fused pass (laser-polio_b200/csrc/lpk_hot.cuh:
// agenda bytes, risk code and pre-test, deadline timers, hot_event) compiled for the HOST and driven by a sequential
// emulation of the sweep, so that tests/test_hot_model.py can hold it to the oracle's canonical tick loop without a GPU.
// The device kernel (csrc/lpk_tick.cu) runs the very same LPK_HD functions; what this file restates is only the sweep's
// control flow (which agents are sent to hot_event with which flags) and the application of the node-level deltas.
//
// Build (tests/test_hot_model.py does it): nvcc -O2 -std=c++17 -Xcompiler -fPIC -shared -o tests/_build/libhot_model.so tests/hot_model.cu
#include <climits>
#include <cstdint>
#include <cstring>

#include "../laser-polio_b200/csrc/lpk_hot.cuh"

static void apply_delta(const lpk_tick_args &A, const HotDelta &d) {
    const int nd = d.nd, ns = A.n_strains;
    const int64_t c = (int64_t)nd * ns + d.st;
    if (d.hbin >= 0) A.risk_hist[(int64_t)nd * LPK_RISK_BINS + d.hbin] -= 1;
    if (d.gate) { A.new_potential[nd] += 1; if (d.gate & 2) A.new_paralyzed[nd] += 1; }
    if (d.died) { A.deaths[nd] += 1; if (d.died & 2) A.dead_pp[nd] += 1; if (d.died & 4) A.dead_par[nd] += 1; }
    if (d.hit) {
        A.new_exposed_by_strain_prev[c] += 1; A.tx_hits_by_strain[c] += 1;
        A.new_exposed_prev[nd] += 1; A.tx_hits[nd] += 1;
        A.sus[nd] -= 1;
    }
    A.E_cur[c] += d.dE; A.I_cur[c] += d.dI; A.R_cur[nd] += d.dR;
    A.beta_fx[c] += d.dbeta;
    A.exposure_fx[nd] -= d.efx;
    if (d.died & 8) A.sus[nd] -= 1;
    if (d.vx & 1) A.ri_vaccinated[nd] += 1;
    if (d.vx & 4) A.ipv_vaccinated[nd] += 1;
    if (d.vx & 2) {
        const int64_t cr = (int64_t)nd * ns + A.ri_strain;
        A.E_cur[cr] += 1; A.ri_protected[nd] += 1; A.new_exposed[nd] += 1; A.new_exposed_by_strain[cr] += 1;
        A.ri_new_exposed_by_strain[cr] += 1; A.sus[nd] -= 1;
    }
    if (d.vx & 8) A.sia_vaccinated[nd] += 1;
    if (d.vx & 16) {
        const int64_t cs = (int64_t)nd * ns + A.sia_strain;
        A.E_cur[cs] += 1; A.sia_protected[nd] += 1; A.new_exposed[nd] += 1; A.new_exposed_by_strain[cs] += 1;
        A.sia_new_exposed_by_strain[cs] += 1; A.sus[nd] -= 1;
    }
}

extern "C" int hm_build(const lpk_people *people, int64_t n_slots, int32_t t_next, int32_t ri_step) {
    const lpk_people &P = *people;
    const int64_t padded = (P.capacity + 2047) / 2048 * 2048;
    bool over = false;
    for (int64_t i = 0; i < padded; ++i) P.hot[i] = (i < n_slots) ? hot_build_agent(P, i, t_next, P.risk_e0, &over) : (uint8_t)HOT_DEAD;
    if (P.pair_min_dod && P.date_of_death)
        for (int64_t gp = 0; gp < padded / 256; ++gp) {
            int m = INT_MAX;
            for (int k = 0; k < 256; ++k) {
                const int64_t i = gp * 256 + k;
                if (i < n_slots && P.disease_state[i] >= 0 && P.date_of_death[i] < m) m = P.date_of_death[i];
            }
            P.pair_min_dod[gp] = m;
        }
    if (P.pair_ri_max && P.ri_timer)
        for (int64_t gp = 0; gp < padded / 256; ++gp) {
            int m = INT_MIN;
            for (int k = 0; k < 256; ++k) {
                const int64_t i = gp * 256 + k;
                uint8_t rk = 0;
                if (i < n_slots && P.disease_state[i] >= 0 && P.chronically_missed[i] != 1) {
                    if (P.ri_timer[i] > m) m = P.ri_timer[i];
                    rk = ri_tick_index(P.ri_timer[i], 0, ri_step, t_next - 1);
                }
                P.ri_k[i] = rk;
            }
            P.pair_ri_max[gp] = m;
        }
    return over ? 2 : 0;
}
extern "C" int hm_settle(const lpk_people *people, int64_t n_slots, int32_t t_next, int32_t ri_k, int32_t ri_step) {
    for (int64_t i = 0; i < n_slots; ++i) hot_settle_agent(*people, i, t_next, ri_k, ri_step);
    return 0;
}
extern "C" int hm_risk_e0(float rmax) { return hot_risk_e0(rmax); }

// One pass of tick A->tick over agents [0, n).  stats[0] events, [1] candidates, [2] hits, [3] pairs whose date_of_death was read
extern "C" int hm_pass(const lpk_people *people, const lpk_tick_args *args, int64_t n, int64_t *stats) {
    const lpk_people &P = *people;
    const lpk_tick_args &A = *args;
    const bool pending = (A.flags & LPK_F_PENDING) != 0, deaths = (A.flags & LPK_F_DEATHS) != 0;
    const bool ri = (A.flags & LPK_F_RI) != 0, sia = (A.flags & LPK_F_SIA) != 0;
    const int tick = A.tick, e0 = P.risk_e0;
    const uint32_t today = hot_today(tick);
    const float tau_all = ldexpf(1.0f, e0);
    const int64_t total_pairs = (n + 255) >> 8;
    for (int64_t gp = 0; gp < total_pairs; ++gp) {
        const int tn = P.tile_node ? P.tile_node[gp >> 1] : -1;
        if (tn < 0) {  // general pair: one agent at a time, its own node
            for (int64_t i = gp * 256; i < gp * 256 + 256 && i < n; ++i) {
                const uint32_t hb = P.hot[i];
                if (hb == HOT_DEAD) continue;
                const int nd = P.node_id[i];
                uint32_t fl = 0;
                if (pending && hot_is_S(hb)) {
                    const float tau = A.q_prev[nd];
                    if (tau > 0.f) {
                        const uint64_t id = (uint64_t)i + A.id_base;
                        const uint64_t c = expose_ctr(id);
                        uint32_t x[4];
                        philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(tick - 1), LPK_STAGE_EXPOSE, (uint32_t)A.seed,
                                      (uint32_t)(A.seed >> 32), x);
                        const float U = 8388608.0f + (float)half_word(x, expose_hw(id));
                        if (U < fmaf(risk_code_ub((int)(hb & 63u), e0), tau * 65536.0f, 8388609.0f)) fl |= EV_CAND;
                    }
                }
                if ((hb | 0x40u) == (today & 0xFFu)) fl |= EV_FIRE;
                bool dying = false;
                if (deaths && P.date_of_death[i] <= tick) { fl |= EV_DEATH; dying = true; }
                if ((ri || sia) && !dying && (!sia || P.chronically_missed[i] != 1)) {
                    if (ri && P.ri_k[i] == (uint8_t)(A.ri_lazy_k + 1)) fl |= EV_RI;
                    if (ri && P.disease_state[i] >= 0 && P.chronically_missed[i] != 1 &&
                        (P.ri_k[i] == (uint8_t)(A.ri_lazy_k + 1)) != ri_eligible(P.ri_timer[i], A.ri_lazy_k, A.ri_step, tick)) return -11;
                    if (sia && (uint32_t)(tick - P.date_of_birth[i] - A.sia_min_age) <= (uint32_t)(A.sia_max_age - A.sia_min_age) &&
                        A.sia_targeted[nd] != 0)
                        fl |= EV_SIA;
                }
                if (fl) {
                    stats[0]++;
                    if (fl & EV_CAND) stats[1]++;
                    const HotDelta d = hot_event(P, A, i, nd, fl, hot_preload(P, i, fl));
                    stats[2] += d.hit;
                    apply_delta(A, d);
                }
            }
            continue;
        }
        // node-uniform pair: the sweep's byte-lane predicates, lane by lane
        const float tau = pending ? A.q_prev[tn] : 0.f;
        const int mode = !(tau > 0.f) ? 0 : (tau >= tau_all ? 2 : 1);
        const float tauS = hot_tau_scale(tau, e0);
        const bool camp = sia && A.sia_targeted[tn] != 0;
        bool flagged = false;
        if (deaths && P.pair_min_dod[gp] <= tick) { flagged = true; stats[3]++; }
        int left = INT_MAX;
        uint32_t F[2][32];
        for (int lane = 0; lane < 32; ++lane) {
            const int64_t bA = gp * 256 + lane * 4, bB = bA + 128;
            uint32_t hA, hB;
            memcpy(&hA, P.hot + bA, 4); memcpy(&hB, P.hot + bB, 4);
            uint32_t cA = 0, cB = 0;
            if (mode == 1) {
                const uint64_t c = (((uint64_t)gp + (A.id_base >> 8)) << 5) + (uint64_t)lane;
                uint32_t x[4];
                philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(tick - 1), LPK_STAGE_EXPOSE, (uint32_t)A.seed,
                              (uint32_t)(A.seed >> 32), x);
                cA = hot_pretest(hA, x[0], x[1], tauS);
                cB = hot_pretest(hB, x[2], x[3], tauS);
            } else if (mode == 2) {
                cA = cB = 0x01010101u;
            }
            const uint32_t vA = hot_due_word(hA, today), vB = hot_due_word(hB, today);
            const uint32_t fireA = zero_bytes(vA), fireB = zero_bytes(vB);
            if ((any_zero_byte(vA) != 0) != (fireA != 0) || (any_zero_byte(vB) != 0) != (fireB != 0)) return -10;  // SWAR self-check
            uint32_t dmA = 0, dmB = 0, eA = 0, eB = 0, sA = 0, sB = 0;
            const uint32_t aA0 = hot_mask_alive(hA), aB0 = hot_mask_alive(hB);
            if (flagged) {
                for (int k = 0; k < 4; ++k) {
                    if (((aA0 >> (8 * k)) & 1u) && P.date_of_death[bA + k] <= tick) dmA |= 1u << (8 * k);
                    if (((aB0 >> (8 * k)) & 1u) && P.date_of_death[bB + k] <= tick) dmB |= 1u << (8 * k);
                }
                for (int k = 0; k < 4; ++k) {
                    if (((aA0 & ~dmA) >> (8 * k)) & 1u) left = P.date_of_death[bA + k] < left ? P.date_of_death[bA + k] : left;
                    if (((aB0 & ~dmB) >> (8 * k)) & 1u) left = P.date_of_death[bB + k] < left ? P.date_of_death[bB + k] : left;
                }
            }
            const bool ri_pair = ri && P.pair_ri_max[gp] >= A.ri_lazy_k * A.ri_step;
            if (ri_pair || camp) {
                const uint32_t aA = aA0 & ~dmA, aB = aB0 & ~dmB;
                for (int r = 0; r < 2; ++r)
                    for (int k = 0; k < 4; ++k) {
                        const int64_t i = (r ? bB : bA) + k;
                        const bool dying = (((r ? dmB : dmA) >> (8 * k)) & 1u) != 0;
                        if (ri_pair && !dying && P.ri_k[i] == (uint8_t)(A.ri_lazy_k + 1)) (r ? eB : eA) |= 1u << (8 * k);
                        if (!((((r ? aB : aA)) >> (8 * k)) & 1u) || P.chronically_missed[i] == 1) continue;
                        if (ri_pair && (P.ri_k[i] == (uint8_t)(A.ri_lazy_k + 1)) != ri_eligible(P.ri_timer[i], A.ri_lazy_k, A.ri_step, tick)) return -12;
                        if (camp && (uint32_t)(tick - P.date_of_birth[i] - A.sia_min_age) <= (uint32_t)(A.sia_max_age - A.sia_min_age))
                            (r ? sB : sA) |= 1u << (8 * k);
                    }
            }
            F[0][lane] = (cA & hot_mask_S(hA)) | (fireA << 1) | (dmA << 2) | (eA << 3) | (sA << 4);
            F[1][lane] = (cB & hot_mask_S(hB)) | (fireB << 1) | (dmB << 2) | (eB << 3) | (sB << 4);
        }
        if (flagged) P.pair_min_dod[gp] = left;
        for (int lane = 0; lane < 32; ++lane)
            for (int r = 0; r < 2; ++r)
                for (int k = 0; k < 4; ++k) {
                    const uint32_t f = (F[r][lane] >> (8 * k)) & 0x1Fu;
                    if (!f) continue;
                    const int64_t i = gp * 256 + r * 128 + lane * 4 + k;
                    const uint32_t fl = f << 16;
                    stats[0]++;
                    if (fl & EV_CAND) stats[1]++;
                    const HotDelta d = hot_event(P, A, i, tn, fl, hot_preload(P, i, fl));
                    stats[2] += d.hit;
                    apply_delta(A, d);
                }
    }
    return 0;
}
