// lpk_hot.cuh -- the "agenda" representation of the fused pass and the per-agent event logic on it.
//
// What an agent needs from the daily sweep depends on its class only:
//   susceptible   one exposure trial (its risk against the node's tau)
//   exposed       nothing until the day it turns infectious
//   infectious    nothing until the day its paralysis gate opens (paralytic strain) or it recovers
//   recovered / dead / unborn   nothing
// so the sweep reads ONE byte per agent, the agenda byte `hot[i]`:
//   bits 7:6  class   00 exposed   01 infectious   10 inactive (payload 0 recovered, 1 dead / unborn)   11 susceptible
//   bits 5:0  payload susceptible: a 6-bit upper bound of acq_risk_multiplier (4 steps per octave, see risk_code)
//                     exposed / infectious: the day of the agent's next event, modulo 64
// and everything else is event driven: an agent whose payload equals today (or that passes the pre-test of the exposure
// trial) is handed to hot_event, which works on the full-width columns of the reference.
//
// Deadline timers.  The reference tests a countdown timer and then decrements it on every step the agent spends in the
// state that owns the timer (model.py:419-452), with int8 wrap-around; the value tested on tick t is timer0 - (t - t0).
// While an agent is exposed / infectious the owning column therefore holds the DEADLINE  d = (timer + t) mod 256  (t = the
// tick whose test would see `timer`), the value tested on any tick t is (int8)(d - t), and nothing is written on the days in
// between.  Columns are converted back at the state's exit, at death, and by hot_settle_agent whenever the table leaves the
// fused path (to_host(), a day run through the component kernels):
//   exposure_timer    deadline form while disease_state == 1
//   infection_timer   deadline form while disease_state == 2
//   paralysis_timer   deadline form while disease_state == 2 and strain == 0 (the only strain whose timer runs, model.py:447)
// All of it is exact modular arithmetic, so the columns the host reads back equal the reference's bit for bit
// (tests/test_hot_model.py on the CPU, tests/test_gpu_fused.py on the device).
#pragma once
#include "lpk_common.cuh"

#define HOT_E 0x00u
#define HOT_I 0x40u
#define HOT_R 0x80u
#define HOT_DEAD 0x81u
#define HOT_S 0xC0u
#define HOT_LOOKAHEAD 63  // an event further away is reached through check-in events every 63 days

// ring / event flags (bits 16+ of the second entry word; bits 0-15 carry the node)
#define EV_CAND (1u << 16)   // susceptible that passed the pre-test of tick t-1's exposure trial
#define EV_FIRE (1u << 17)   // exposed / infectious agent whose agenda day is today
#define EV_DEATH (1u << 18)  // date_of_death <= t on a vital-dynamics tick
#define EV_RI (1u << 19)     // routine-immunisation eligible today
#define EV_SIA (1u << 20)    // inside today's campaign (node, age window, not chronically missed)

// ---------------------------------------------------------------- risk code
// 6-bit code c = e * 4 + m of the smallest value  ub(c) = 2^(e - e0) * (1 + m / 4)  that is >= risk.  e0 is a per-table
// constant chosen so that the largest risk in the table fits (hot_risk_e0).  Non-positive / NaN risks never hit
// (p_expose) and take code 0.  *over is set when the risk exceeds ub(63).
LPK_HD int risk_code(float rk, int e0, bool *over) {
    if (!(rk > 0.f)) return 0;
    const uint32_t b = lpk_f2u(rk);
    const int c = ((int)(b >> 23) - 127 + e0) * 4 + (int)((b >> 21) & 3u) + ((b & 0x1FFFFFu) ? 1 : 0);
    if (c > 63) { if (over) *over = true; return 63; }
    return c < 0 ? 0 : c;
}
LPK_HD float risk_code_ub(int c, int e0) { return ldexpf(1.0f + 0.25f * (float)(c & 3), (c >> 2) - e0); }
// e0 for a table whose largest finite risk is rmax: ub(63) = 1.75 * 2^(15 - e0) >= rmax
LPK_HD int hot_risk_e0(float rmax) {
    if (!(rmax > 0.f)) return 0;
    int ex;
    const float fr = frexpf(rmax, &ex);  // rmax = fr * 2^ex, fr in [0.5, 1)
    const int need = (fr > 0.875f) ? ex : ex - 1;  // smallest k with 1.75 * 2^k >= rmax
    int e0 = 15 - need;
    return e0 < -60 ? -60 : (e0 > 60 ? 60 : e0);
}
// In the sweep the code is decoded without arithmetic: the agenda byte placed at bits 21-28 of a float is
// 2^(48 + e - 127) * (1 + m / 4) for a susceptible (class bits 11) and at most 2^(33 - 127) for every other class, so
//     U < fma(decoded, tau * 2^(95 - e0), 2^23 + 1),   U = 2^23 + h16
// is the 16-bit pre-test of lpk_tick.cu with the agent's risk replaced by its upper bound.
LPK_HD float hot_tau_scale(float tau, int e0) { return ldexpf(tau, 95 - e0); }

LPK_HD uint8_t hot_due(int tick, int days) {
    const int d = days < 0 ? 0 : (days > HOT_LOOKAHEAD ? HOT_LOOKAHEAD : days);
    return (uint8_t)((tick + d) & 63);
}

// ---- byte-lane helpers (masks carry bit 0 of each byte) ------------------------------------------------------------
LPK_HD uint32_t zero_bytes(uint32_t v) {  // exact, per byte: v == 0
    return (~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) >> 7) & 0x01010101u;
}
LPK_HD uint32_t hot_mask_S(uint32_t h) { return ((h & (h << 1)) >> 7) & 0x01010101u; }
LPK_HD bool hot_is_S(uint32_t byte) { return (byte >> 6) == 3u; }
LPK_HD uint32_t hot_mask_alive(uint32_t h) { return zero_bytes(h ^ (HOT_DEAD * 0x01010101u)) ^ 0x01010101u; }
// agents of the quad whose agenda day is today: class 0x (exposed / infectious) and payload == tick mod 64; the word
// has a zero byte exactly for those.  today = hot_today(tick)
LPK_HD uint32_t hot_today(int tick) { return (0x40u | ((uint32_t)tick & 63u)) * 0x01010101u; }
LPK_HD uint32_t hot_due_word(uint32_t h, uint32_t today) { return (h | 0x40404040u) ^ today; }
LPK_HD uint32_t any_zero_byte(uint32_t v) { return (v - 0x01010101u) & ~v & 0x80808080u; }  // != 0 iff some byte is 0

// Pre-test of the exposure trial for the four agents of agenda word h (lpk_hot.cuh, risk code): bit 0 of byte k set when
// agent k's 16-bit high half can still be a hit.  xa / xb: the quad's two words of the EXPOSE block (high halves of the
// draws: agent 0 = low half of xa, 1 = high half of xa, 2 = low half of xb, 3 = high half of xb).  Six instructions per
// agent; classes other than susceptible decode to < 2^-16 of the smallest susceptible bound and pass with probability
// ~2^-16 (the caller masks with hot_mask_S when anything passed).
LPK_HD uint32_t hot_pretest(uint32_t h, uint32_t xa, uint32_t xb, float tauS) {
    const uint32_t KM = 0x1FFFFFFFu, k23 = 0x4B000000u;
    const float c = 8388609.0f;
    const float d0 = lpk_u2f((h << 21) & KM), d1 = lpk_u2f((h << 13) & KM), d2 = lpk_u2f((h << 5) & KM), d3 = lpk_u2f(h >> 3);
#ifdef __CUDA_ARCH__
    const float u0 = __uint_as_float(__byte_perm(xa, k23, 0x7610)), u1 = __uint_as_float(__byte_perm(xa, k23, 0x7632));
    const float u2 = __uint_as_float(__byte_perm(xb, k23, 0x7610)), u3 = __uint_as_float(__byte_perm(xb, k23, 0x7632));
#else
    const float u0 = lpk_u2f(k23 | (xa & 0xFFFFu)), u1 = lpk_u2f(k23 | (xa >> 16));
    const float u2 = lpk_u2f(k23 | (xb & 0xFFFFu)), u3 = lpk_u2f(k23 | (xb >> 16));
#endif
    return ((u0 < fmaf(d0, tauS, c)) ? 1u : 0u) | ((u1 < fmaf(d1, tauS, c)) ? 0x100u : 0u) |
           ((u2 < fmaf(d2, tauS, c)) ? 0x10000u : 0u) | ((u3 < fmaf(d3, tauS, c)) ? 0x1000000u : 0u);
}


// ---------------------------------------------------------------- the event record
// The eight byte columns an event can touch (disease_state, strain, exposure_timer, infection_timer, paralysis_timer,
// potentially_paralyzed, paralyzed, ipv_protected) are gathered, while the table is on the fused path, in ONE 8-byte record
// per agent, lpk_people.rec[i]: an event costs one 8-byte load and one 8-byte store instead of up to twelve scattered
// sector accesses in eight arrays (profiles/r2_fused_v31_*: after a campaign the handler, not the sweep, set the pace).
// The reference-dtype columns stay allocated and canonical but are not touched by the pass; lpk_hot_build fills the
// records from them and lpk_hot_settle writes them back (deadlines converted to countdown values).
struct HotRec {
    int8_t state, strain, et, it, pt, pq, par, ipv;  // et / it / pt: deadlines while the owning state lasts (see top)
};
union HotRecU {
    unsigned long long u;
    HotRec r;
};

// ---------------------------------------------------------------- canonical <-> agenda
// Record and agenda byte of agent i as the table stands before tick t_next runs (timers of exposed / infectious agents
// become deadlines in the record).  Returns the byte; *over: risk beyond the code range.
LPK_HD uint8_t hot_build_agent(const lpk_people &P, int64_t i, int t_next, int e0, bool *over) {
    HotRecU x;
    x.r.state = P.disease_state[i]; x.r.strain = P.strain[i];
    x.r.et = P.exposure_timer[i]; x.r.it = P.infection_timer[i]; x.r.pt = P.paralysis_timer[i];
    x.r.pq = P.potentially_paralyzed[i]; x.r.par = P.paralyzed[i]; x.r.ipv = P.ipv_protected[i];
    const int8_t s = x.r.state;
    const uint8_t t8 = (uint8_t)t_next;
    uint8_t h = s == 3 ? (uint8_t)HOT_R : (uint8_t)HOT_DEAD;
    if (s == 0) {
        h = (uint8_t)(HOT_S | risk_code(P.acq_risk_multiplier[i], e0, over));
    } else if (s == 1) {
        h = (uint8_t)(HOT_E | hot_due(t_next, x.r.et));
        x.r.et = (int8_t)(uint8_t)((uint8_t)x.r.et + t8);
    } else if (s == 2) {
        int next = x.r.it < 0 ? 0 : x.r.it;
        x.r.it = (int8_t)(uint8_t)((uint8_t)x.r.it + t8);
        if (x.r.strain == 0) {
            if (x.r.pq == -1) { const int g = x.r.pt < 0 ? 0 : x.r.pt; if (g < next) next = g; }
            x.r.pt = (int8_t)(uint8_t)((uint8_t)x.r.pt + t8);
        }
        h = (uint8_t)(HOT_I | hot_due(t_next, next));
    }
    P.rec[i] = x.u;
    return h;
}
// Lazy RI countdown (include/lpk.h, lpk_tick_args.ri_lazy_k): the agent's timer after k owed subtractions, and whether the
// RI tick `tick` (the k_after-th subtraction) finds it eligible -- reference model.py:1833-1843 on the wrapped int16 value.
LPK_HD int16_t ri_owed(int16_t stored, int k, int step) { return (int16_t)(uint16_t)((uint16_t)stored - (uint16_t)(k * step)); }
LPK_HD bool ri_eligible(int16_t stored, int k_before, int step, int tick) {
    const int timer = (int)ri_owed(stored, k_before, step) - step;
    return (tick == step) ? (timer <= 0 && timer >= -step) : (tick > step && timer <= 0 && timer > -step);
}
// Which RI tick after a rebase (1 = the next one, ...) finds an alive, not chronically missed agent eligible, given its
// timer `stored` with `k_done` subtractions owed and `t_last` = the tick before the next one to run; 0 = none of the next
// 254.  The timer after j more subtractions is stored - (k_done + j) * step; it lies in (-step, 0] for exactly one j
// (timer > 0), and on the very first RI tick of a run (tick == step) the window also holds -step (model.py:1836-1843).
LPK_HD uint8_t ri_tick_index(int16_t stored, int k_done, int step, int t_last) {
    const int cur = (int)ri_owed(stored, k_done, step);
    const int next_ri_tick = (t_last / step + 1) * step;  // first RI tick > t_last
    int j = 0;
    if (cur >= 1) j = (cur + step - 1) / step;
    else if (cur == 0 && next_ri_tick == step) j = 1;
    return (j >= 1 && j <= 254) ? (uint8_t)j : (uint8_t)0;
}
// The inverse: the record back into the reference's columns, deadlines -> the values tick t_next would test; the RI
// countdown's debt is paid.
LPK_HD void hot_settle_agent(const lpk_people &P, int64_t i, int t_next, int ri_k, int ri_step) {
    HotRecU x;
    x.u = P.rec[i];
    const int8_t s = x.r.state;
    const uint8_t t8 = (uint8_t)t_next;
    if (ri_k && P.ri_timer && s >= 0 && P.chronically_missed[i] != 1) P.ri_timer[i] = ri_owed(P.ri_timer[i], ri_k, ri_step);
    if (s == 1) {
        x.r.et = (int8_t)(uint8_t)((uint8_t)x.r.et - t8);
    } else if (s == 2) {
        x.r.it = (int8_t)(uint8_t)((uint8_t)x.r.it - t8);
        if (x.r.strain == 0) x.r.pt = (int8_t)(uint8_t)((uint8_t)x.r.pt - t8);
    }
    P.disease_state[i] = x.r.state; P.strain[i] = x.r.strain;
    P.exposure_timer[i] = x.r.et; P.infection_timer[i] = x.r.it; P.paralysis_timer[i] = x.r.pt;
    P.potentially_paralyzed[i] = x.r.pq; P.paralyzed[i] = x.r.par; P.ipv_protected[i] = x.r.ipv;
}

// ---------------------------------------------------------------- the event
// What hot_event needs from the agent's columns, loaded one ring batch ahead of its use on the device.
struct HotPre {
    unsigned long long rec;
    float rk, inf;
};
LPK_HD HotPre hot_preload(const lpk_people &P, int64_t i, uint32_t fl) {
    HotPre r;
    r.rec = P.rec[i];
    r.inf = (fl & (EV_FIRE | EV_CAND | EV_DEATH)) ? P.daily_infectivity[i] : 0.f;
    r.rk = (fl & (EV_CAND | EV_RI | EV_SIA | EV_DEATH)) ? P.acq_risk_multiplier[i] : 0.f;
    return r;
}
// Node-level consequences of one event, applied by the caller (per-warp shared-memory accumulators on the device,
// plain arrays in the host model).
struct HotDelta {
    int nd;
    int8_t st;      // strain of the hit / the E, I and infectivity changes
    int8_t dE, dI, dR;
    uint8_t hit;    // exposure hit of tick t-1
    uint8_t vx;     // 1 RI vaccinated, 2 RI protected, 4 IPV vaccinated, 8 SIA vaccinated, 16 SIA protected
    uint8_t gate;   // 1 newly potentially paralysed, 2 newly paralysed
    uint8_t died;   // 1 died, 2 was potentially paralysed, 4 was paralysed, 8 died susceptible
    long long dbeta;  // change of the infectivity tally (2^30 fixed point)
    long long efx;    // risk (fixed point) of an agent that left the susceptible class; hbin its histogram bin, else -1
    int hbin;
};

// the exact exposure trial of tick t-1 for one susceptible (both Philox blocks; the sweep only pre-tests)
LPK_HD bool hot_exact_trial(const lpk_tick_args &A, int64_t i, int nd, float rk) {
    const float tau = A.q_prev[nd];
    if (!(tau > 0.f)) return false;
    const uint64_t id = (uint64_t)i + A.id_base;
    const uint64_t c = expose_ctr(id);
    const int hw = expose_hw(id);
    uint32_t h[4], l[4];
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(A.tick - 1), LPK_STAGE_EXPOSE, (uint32_t)A.seed, (uint32_t)(A.seed >> 32), h);
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(A.tick - 1), LPK_STAGE_EXPOSE_LO, (uint32_t)A.seed, (uint32_t)(A.seed >> 32), l);
    const uint32_t X = (half_word(h, hw) << 16) | half_word(l, hw);
#ifdef __CUDA_ARCH__
    const float x = __fmul_rn(rk, tau);
#else
    const float x = rk * tau;
#endif
    return expose_test(p_expose(x), X);
}
LPK_HD int8_t hot_pick_strain(const lpk_tick_args &A, int64_t i, int nd) {  // model.py:1127-1141
    uint32_t y[4];
    philox_agent(A.seed, (uint64_t)i + A.id_base, (uint32_t)(A.tick - 1), LPK_STAGE_STRAIN, y);
    const double u = u53(y[0], y[1]);
    const int ns = A.n_strains;
    for (int k = 0; k < ns; ++k)
        if (u < A.cdf_prev[(int64_t)nd * ns + k]) return (int8_t)k;
    return 0;
}

// One agent with something to do in the pass of tick t, in the reference's order: pending exposure trial of t-1
// (model.py:1010-1149 replacement) -> death (1767-1781) -> disease-state step (419-452) -> RI draws (1825-1854) -> campaign
// draws (2030-2059).  `fl` says why the sweep sent it; every condition is re-checked against the agent's record.
LPK_HD HotDelta hot_event(const lpk_people &P, const lpk_tick_args &A, int64_t i, int nd, uint32_t fl, const HotPre &pre) {
    HotDelta d;
    d.nd = nd; d.st = 0; d.dE = d.dI = d.dR = 0; d.hit = d.vx = d.gate = d.died = 0; d.dbeta = 0; d.efx = 0; d.hbin = -1;
    const int tick = A.tick;
    const uint8_t t8 = (uint8_t)tick;
    HotRecU x;
    x.u = pre.rec;
    HotRec &r = x.r;
    bool hot_set = false;
    uint8_t hot_new = 0;

    // ---- 1. exposure trial of tick t-1 (agents born today were not there)
    if ((fl & EV_CAND) && r.state == 0 && (A.flags & LPK_F_PENDING)) {
        const bool born_today = (A.flags & LPK_F_DEATHS) && P.date_of_birth && P.date_of_birth[i] == tick;
        if (!born_today && hot_exact_trial(A, i, nd, pre.rk)) {
            r.state = 1;
            r.strain = hot_pick_strain(A, i, nd);
            d.hit = 1; d.dE = 1;
            d.efx = risk_fx(pre.rk); d.hbin = risk_bin(pre.rk);
            r.et = (int8_t)(uint8_t)((uint8_t)r.et + t8);  // deadline: the first step that tests the exposure timer is today's
        }
    }
    // ---- 2. death: the record goes back to countdown values as of today's (skipped) step
    if ((fl & EV_DEATH) && r.state >= 0 && P.date_of_death[i] <= tick) {
        d.died = 1;
        if (r.state == 0) { d.died |= 8; d.efx = risk_fx(pre.rk); d.hbin = risk_bin(pre.rk); }
        if (r.state == 1) {
            d.dE -= 1;
            r.et = (int8_t)(uint8_t)((uint8_t)r.et - t8);
        } else if (r.state == 2) {
            d.dI -= 1;
            d.dbeta -= to_fx((double)pre.inf * A.strain_r0_scalars[r.strain]);
            r.it = (int8_t)(uint8_t)((uint8_t)r.it - t8);
            if (r.strain == 0) r.pt = (int8_t)(uint8_t)((uint8_t)r.pt - t8);
        } else if (r.state == 3) {
            d.dR -= 1;
        }
        d.st = r.strain;
        if (r.pq == 1) d.died |= 2;
        if (r.par == 1) d.died |= 4;
        if (A.ri_lazy_k && P.ri_timer && P.chronically_missed[i] != 1)  // the dead stop counting down: pay the debt so far
            P.ri_timer[i] = ri_owed(P.ri_timer[i], A.ri_lazy_k, A.ri_step);
        r.state = -1;
        P.rec[i] = x.u;
        P.hot[i] = HOT_DEAD;
        return d;
    }
    // ---- 3. disease-state step of tick t
    bool turned = false;
    if (r.state == 1) {
        const int8_t etv = (int8_t)(uint8_t)((uint8_t)r.et - t8);  // the value today's step tests
        if (etv <= 0) {
            r.et = (int8_t)(etv - 1);
            r.state = 2; turned = true;
        } else {
            hot_new = (uint8_t)(HOT_E | hot_due(tick, etv)); hot_set = true;
        }
    }
    if (r.state == 2) {
        if (turned) r.it = (int8_t)(uint8_t)((uint8_t)r.it + t8);
        const int8_t itv = (int8_t)(uint8_t)((uint8_t)r.it - t8);
        const long long fxv = to_fx((double)pre.inf * A.strain_r0_scalars[r.strain]);
        if (turned) { d.dE -= 1; d.dI += 1; d.dbeta += fxv; }
        const bool recover = itv <= 0;
        const bool wild = r.strain == 0;
        int8_t ptv = 0;
        if (wild) {  // paralysis part, model.py:432-452
            if (turned) r.pt = (int8_t)(uint8_t)((uint8_t)r.pt + t8);
            ptv = (int8_t)(uint8_t)((uint8_t)r.pt - t8);
            if (ptv <= 0 && r.pq == -1) {
                if (r.ipv == 0) {
                    r.pq = 1;
                    d.gate = 1;
                    uint32_t y[4];
                    philox_agent(A.seed, (uint64_t)i + A.id_base, (uint32_t)tick, LPK_STAGE_PARALYSIS, y);
                    if (u53(y[0], y[1]) < (double)A.p_paralysis) { r.par = 1; d.gate = 3; }
                } else {
                    r.pq = 0;
                }
            }
        }
        if (recover) {
            r.it = (int8_t)(itv - 1);
            if (wild) r.pt = (int8_t)(ptv - 1);
            d.dI -= 1; d.dbeta -= fxv; d.dR += 1;
            r.state = 3;
            hot_new = HOT_R; hot_set = true;
        } else {
            int next = itv;
            if (wild && r.pq == -1 && ptv < next) next = ptv;  // the gate is still closed: ptv > 0
            hot_new = (uint8_t)(HOT_I | hot_due(tick, next)); hot_set = true;
        }
    }
    d.st = r.strain;
    // ---- 4. vaccine draws, after the agent's own disease-state step (the reference's run order); the dead take none
    int8_t vstrain = -1;
    if ((fl & EV_RI) && (A.flags & LPK_F_RI) && r.state >= 0) {
        uint32_t y[4];
        philox_agent(A.seed, (uint64_t)i + A.id_base, (uint32_t)tick, LPK_STAGE_RI, y);
        if (u53(y[0], y[1]) < A.vx_prob_ri[nd]) {
            d.vx |= 1;
            if (r.state == 0) { r.state = 1; vstrain = (int8_t)A.ri_strain; d.vx |= 2; }
        }
        if (u53(y[2], y[3]) < A.vx_prob_ipv[nd]) { d.vx |= 4; r.ipv = 1; }
    }
    if ((fl & EV_SIA) && (A.flags & LPK_F_SIA) && r.state >= 0) {
        uint32_t y[4];
        philox_agent(A.seed, (uint64_t)i + A.id_base, (uint32_t)tick, LPK_STAGE_SIA | (A.sia_event_idx << 8), y);
        const double u = u53(y[0], y[1]), pv = (double)A.vx_prob_sia[nd];
        if (u < pv) {
            d.vx |= 8;
            if (r.state == 0 && u < pv * A.sia_vx_eff) { r.state = 1; vstrain = (int8_t)A.sia_strain; d.vx |= 16; }
        }
    }
    if (d.vx & 18) {  // left S through a vaccine: exposed from tomorrow's step on
        r.strain = vstrain;
        d.st = vstrain;
        d.efx = risk_fx(pre.rk); d.hbin = risk_bin(pre.rk);
        const int8_t et0 = r.et;
        r.et = (int8_t)(uint8_t)((uint8_t)et0 + t8 + 1u);
        hot_new = (uint8_t)(HOT_E | hot_due(tick, 1 + (et0 < 0 ? 0 : et0))); hot_set = true;
    }
    if (x.u != pre.rec) P.rec[i] = x.u;
    if (hot_set) P.hot[i] = hot_new;
    return d;
}
