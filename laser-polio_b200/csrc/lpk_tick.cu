// lpk_tick.cu -- the fused tick: ONE streaming pass over the agent table per simulated day.
//
// Pass for tick t, per agent, in the reference's order (include/lpk.h, "Fused tick"):
//   pending tick t-1:  exposure trial (tx_infect)  ->  census (count_SEIRP)
//   tick t:            deaths (get_deaths) -> disease state (disease_state_step) -> RI (fast_ri) -> tally (tx_step_prep)
// Every draw is Philox(seed; agent, tick, stage), so the fused pass reproduces the per-function kernels bit for bit
// (tests/test_gpu_fused.py).
//
// The pass is instruction-issue bound, not bandwidth bound (profiles/r1_fused_v8_*: 534 warp-instructions per 128 agents
// at 20 % of DRAM peak), so the design minimises instructions per agent:
//   * per-node integer tallies (susceptibles, recovered, risk sum, risk histogram) are CARRIED from tick to tick and only
//     corrected where an agent changes class, so the streaming loop counts nothing;
//   * chunks of 32 K agents that lie in one node (nearly all of them) run a loop in which a lane owns 8 agents per
//     iteration (its quads in an even / odd row pair) served by ONE Philox block: the 16-bit high halves reject > 99.9 %
//     of the trials with 3 instructions per agent, the exact 32-bit test runs out of line for the rest;
//   * the 2 % of agents that are exposed or infectious are pushed to a per-warp shared-memory ring and handled 32 at a
//     time with all lanes busy (they sit in ~90 % of the 128-agent rows, so handling them in place made every warp walk
//     the long disease-state path with one or two live lanes);
//   * everything rare (a hit, a death, an RI-eligible agent, a node boundary) lives in __noinline__ functions.
#include <cstdlib>

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched from the driver at run time, liblpk does not link libcuda)

#include "lpk_host.cuh"
#include "lpk_stages.cuh"

struct PassParams {
    CUtensorMap tmap;    // the six byte columns of the disease state as ONE 2-D tensor [column, agent] (use_tmap)
    lpk_people P;
    lpk_tick_args A;
    uint32_t *unit_ctr;  // work counter of this launch (zeroed on the stream before the kernel)
    uint32_t debug;      // timing experiments only (LPK_PASS_DEBUG): 1 = drop ring batches
    uint32_t use_tmap;   // the byte columns lie at one constant stride (device.DeviceState's arena): one tensor copy per pair
};

#define QCAP 512         // ring entries per warp: 31 left over + the 256 agents of one iteration fit
#define LPK_UNIT_LOG 3   // a work unit = 8 consecutive pairs of 128-agent rows (2048 agents), claimed by one warp at a time
#define LPK_UNIT_PAIRS (1 << LPK_UNIT_LOG)

// ------------------------------------------------------------------ rare paths (out of line, direct atomics)
__device__ __forceinline__ DevRng stage_rng(const PassParams &pp) {
    DevRng rng;
    rng.seed = pp.A.seed; rng.tick = (uint32_t)pp.A.tick; rng.u1 = nullptr; rng.u2 = nullptr; rng.x = nullptr;
    rng.id_base = pp.A.id_base;
    return rng;
}

// The susceptible-side tallies (count, sum of risks, risk histogram per node) are carried from tick to tick and only
// CORRECTED when an agent leaves the susceptible state (exposure hit, RI exposure, death) or is born; they are exact
// integers, so the running values equal a from-scratch tally bit for bit (tests/test_gpu_fused.py).
__device__ __noinline__ void leave_S(const PassParams &pp, int64_t i, int nd) {
    const lpk_tick_args &A = pp.A;
    const float rk = pp.P.acq_risk_multiplier[i];
    atomicAdd(reinterpret_cast<unsigned long long *>(&A.sus[nd]), (unsigned long long)(-1ll));
    red_add(&A.exposure_fx[nd], -__float2ll_rn(rk * 1073741824.0f));
    atomicAdd(&A.risk_hist[(int64_t)nd * LPK_RISK_BINS + risk_bin(rk)], -1);
}

// bookkeeping of an exposure hit of tick t-1 (the agent's risk already in a register): categorical strain pick
// (model.py:1127-1141), rows t-1, susceptible-side tallies; returns the strain
__device__ __noinline__ int8_t expose_bookkeeping_rk(const PassParams &pp, int64_t i, int nd, float rk) {
    const lpk_tick_args &A = pp.A;
    uint32_t y[4];
    philox_agent(A.seed, (uint64_t)i + A.id_base, (uint32_t)(A.tick - 1), LPK_STAGE_STRAIN, y);
    const double r = u53(y[0], y[1]);
    const int ns = A.n_strains;
    int assigned = 0;
    for (int s = 0; s < ns; ++s)
        if (r < A.cdf_prev[(int64_t)nd * ns + s]) { assigned = s; break; }
    pp.P.strain[i] = (int8_t)assigned;
    atomicAdd(&A.new_exposed_prev[nd], 1);
    atomicAdd(&A.new_exposed_by_strain_prev[(int64_t)nd * ns + assigned], 1);
    atomicAdd(&A.tx_hits[nd], 1);
    atomicAdd(reinterpret_cast<unsigned long long *>(&A.sus[nd]), (unsigned long long)(-1ll));
    red_add(&A.exposure_fx[nd], -__float2ll_rn(rk * 1073741824.0f));
    atomicAdd(&A.risk_hist[(int64_t)nd * LPK_RISK_BINS + risk_bin(rk)], -1);
    return (int8_t)assigned;
}
// The exposed / infectious census by strain and the infectivity tally are CARRIED like the susceptible-side tallies:
// E_cur / I_cur / beta_fx change only when an agent changes class, so the pass does nothing for an agent that merely
// counts a timer down.  Slow-path form (direct atomics): the agent moves from class `from` to class `to` (-1 dead, 0 S,
// 1 E, 2 I, 3 R); its strain must already be final.
__device__ __noinline__ void carried_move(const PassParams &pp, int64_t i, int nd, int8_t from, int8_t to) {
    const lpk_tick_args &A = pp.A;
    if (from == to) return;
    const bool ei = from == 1 || from == 2 || to == 1 || to == 2;
    if (ei) {
        const int st = pp.P.strain[i];
        const int64_t c = (int64_t)nd * A.n_strains + st;
        if (from == 1) atomicAdd(&A.E_cur[c], -1);
        if (to == 1) atomicAdd(&A.E_cur[c], 1);
        if (from == 2 || to == 2) {
            const long long fx = to_fx((double)pp.P.daily_infectivity[i] * A.strain_r0_scalars[st]);
            atomicAdd(&A.I_cur[c], to == 2 ? 1 : -1);
            red_add(&A.beta_fx[c], to == 2 ? fx : -fx);
        }
    }
    if (to == 3) atomicAdd(&A.R_cur[nd], 1);
    if (from == 3) atomicAdd(&A.R_cur[nd], -1);
}
__device__ __noinline__ void expose_agent(const PassParams &pp, int64_t i, int nd) {
    const int st = expose_bookkeeping_rk(pp, i, nd, pp.P.acq_risk_multiplier[i]);
    const int64_t c = (int64_t)nd * pp.A.n_strains + st;
    atomicAdd(&pp.A.tx_hits_by_strain[c], 1);
    atomicAdd(&pp.A.E_cur[c], 1);
}

__device__ __noinline__ void kill_agent(const PassParams &pp, int64_t i, int nd, int8_t state_before) {
    if (state_before == 0) leave_S(pp, i, nd);
    else carried_move(pp, i, nd, state_before, -1);
    atomicAdd(&pp.A.deaths[nd], 1);
    if (pp.P.potentially_paralyzed[i] == 1) atomicAdd(&pp.A.dead_pp[nd], 1);
    if (pp.P.paralyzed[i] == 1) atomicAdd(&pp.A.dead_par[nd], 1);
}

__device__ __noinline__ int8_t ds_agent_ol(const PassParams &pp, int64_t i, int8_t s, int nd) {
    const lpk_people &P = pp.P;
    const int8_t s2 = ds_agent(i, s, P.node_id, P.strain, P.exposure_timer, P.infection_timer, P.potentially_paralyzed, P.paralyzed,
                               P.ipv_protected, P.paralysis_timer, (double)pp.A.p_paralysis, pp.A.new_potential, pp.A.new_paralyzed,
                               stage_rng(pp));
    carried_move(pp, i, nd, s, s2);
    return s2;
}

// routine immunisation for one quad (reference model.py:1825-1854); returns the new state word
__device__ __noinline__ uint32_t ri_quad(const PassParams &pp, int64_t base, int valid, uint32_t w) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const int step = A.ri_step;
    const int64_t t = A.tick;
    const bool first = (t == step), later = (t > step);
    const uint32_t m = load_b4(reinterpret_cast<const int8_t *>(P.chronically_missed), base, valid, 1);
    int tm[4];
    load_s4(P.ri_timer, base, valid, tm);
    bool touched = false;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const int8_t s = byte_of(w, k);
        if (s < 0 || byte_of(m, k) == 1) continue;
        const int timer = tm[k] - step;
        tm[k] = timer;
        touched = true;
        const bool eligible = first ? (timer <= 0 && timer >= -step) : (later && timer <= 0 && timer > -step);
        if (!eligible) continue;
        const int64_t i = base + k;
        const int nd = P.node_id[i];
        uint32_t x[4];
        philox_agent(A.seed, (uint64_t)i + A.id_base, (uint32_t)A.tick, LPK_STAGE_RI, x);
        const double u1 = u53(x[0], x[1]), u2 = u53(x[2], x[3]);
        if (u1 < A.vx_prob_ri[nd]) {
            atomicAdd(&A.ri_vaccinated[nd], 1);
            if (s == 0) {
                w = set_byte(w, k, 1);
                P.strain[i] = (int8_t)A.ri_strain;
                leave_S(pp, i, nd);
                const int64_t c = (int64_t)nd * A.n_strains + A.ri_strain;
                atomicAdd(&A.E_cur[c], 1);
                atomicAdd(&A.ri_protected[nd], 1);
                atomicAdd(&A.new_exposed[nd], 1);
                atomicAdd(&A.new_exposed_by_strain[c], 1);
                atomicAdd(&A.ri_new_exposed_by_strain[c], 1);
            }
        }
        if (u2 < A.vx_prob_ipv[nd]) { atomicAdd(&A.ipv_vaccinated[nd], 1); P.ipv_protected[i] = 1; }
    }
    if (touched) {
        if (valid == 4) *reinterpret_cast<short4 *>(P.ri_timer + base) = make_short4((short)tm[0], (short)tm[1], (short)tm[2], (short)tm[3]);
        else for (int k = 0; k < valid; ++k) P.ri_timer[base + k] = (int16_t)tm[k];
    }
    return w;
}

// one campaign event for one quad (reference model.py:2030-2059), direct atomics; returns the new state word
__device__ __forceinline__ uint32_t sia_age_mask(const int4 &d, int tick, int lo, uint32_t span) {
    return ((uint32_t)(tick - d.x - lo) <= span ? 1u : 0u) | ((uint32_t)(tick - d.y - lo) <= span ? 0x100u : 0u) |
           ((uint32_t)(tick - d.z - lo) <= span ? 0x10000u : 0u) | ((uint32_t)(tick - d.w - lo) <= span ? 0x1000000u : 0u);
}
__device__ __noinline__ uint32_t sia_quad(const PassParams &pp, int64_t base, int valid, uint32_t w) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const uint32_t span = (uint32_t)(A.sia_max_age - A.sia_min_age);
#pragma unroll 1
    for (int k = 0; k < valid; ++k) {
        const int8_t s = byte_of(w, k);
        const int64_t i = base + k;
        if (s < 0 || P.chronically_missed[i] == 1) continue;
        if ((uint32_t)(A.tick - P.date_of_birth[i] - A.sia_min_age) > span) continue;
        const int nd = P.node_id[i];
        if (A.sia_targeted[nd] == 0) continue;
        uint32_t x[4];
        philox_agent(A.seed, (uint64_t)i + A.id_base, (uint32_t)A.tick, LPK_STAGE_SIA | (A.sia_event_idx << 8), x);
        const double r = u53(x[0], x[1]), pv = (double)A.vx_prob_sia[nd];
        if (r < pv) {
            atomicAdd(&A.sia_vaccinated[nd], 1);
            if (s == 0 && r < pv * A.sia_vx_eff) {
                w = set_byte(w, k, 1);
                P.strain[i] = (int8_t)A.sia_strain;
                leave_S(pp, i, nd);
                const int64_t c = (int64_t)nd * A.n_strains + A.sia_strain;
                atomicAdd(&A.E_cur[c], 1);
                atomicAdd(&A.sia_protected[nd], 1);
                atomicAdd(&A.new_exposed[nd], 1);
                atomicAdd(&A.new_exposed_by_strain[c], 1);
                atomicAdd(&A.sia_new_exposed_by_strain[c], 1);
            }
        }
    }
    return w;
}

// Generic quad: mixed node ids, the table's tail, or agents born after tick t-1's transmission.  One agent at a
// time with direct atomics; reached for a few quads per node boundary, so its cost is irrelevant.
__device__ __noinline__ uint32_t slow_quad(const PassParams &pp, int64_t b, int valid, uint32_t w, bool deaths, bool ri,
                                           int64_t count_prev) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const bool pending = (A.flags & LPK_F_PENDING) != 0;
    uint32_t nw = w;
    uint32_t x[4] = {0u, 0u, 0u, 0u};
    if (pending) expose_words_quad(A.seed, (uint64_t)b + A.id_base, (uint32_t)(A.tick - 1), x);
#pragma unroll 1
    for (int k = 0; k < valid; ++k) {
        int8_t s = byte_of(nw, k);
        if (s < 0) continue;
        const int64_t i = b + k;
        const int nd = P.node_id[i];
        if (pending && i < count_prev) {
            if (s == 0) {
                const float tau = A.q_prev[nd];
                if (tau > 0.f && expose_test(p_expose(__fmul_rn(P.acq_risk_multiplier[i], tau)), x[k])) { expose_agent(pp, i, nd); s = 1; }
            }
        }
        if (deaths && P.date_of_death[i] <= A.tick) { kill_agent(pp, i, nd, s); s = -1; }
        if (s == 1 || s == 2) s = ds_agent_ol(pp, i, s, nd);
        nw = set_byte(nw, k, s);
    }
    if (ri) nw = ri_quad(pp, b, valid, nw);
    if (A.flags & LPK_F_SIA) nw = sia_quad(pp, b, valid, nw);
    return nw;
}

// ------------------------------------------------------------------ exposure trial of a quad
// High halves of the quad's four draws: x[2 * par], x[2 * par + 1] of the pair's EXPOSE block (par = row parity).
// Pre-test, 3 instructions per agent: U = 2^23 + h16 as a float (one PRMT), T = fma(risk, tau * 2^16, 2^23 + 1); a hit needs
// X < floor(p * 2^32) with p <= risk * tau, hence h16 < risk * tau * 2^16, hence U < T (the + 1 covers both roundings).
__device__ __forceinline__ bool pretest_quad(uint32_t xa, uint32_t xb, const float4 &rk, float tau16) {
    const uint32_t k23 = 0x4B000000u;
    const float c = 8388609.0f;
    return (__uint_as_float(__byte_perm(xa, k23, 0x7610)) < fmaf(rk.x, tau16, c)) |
           (__uint_as_float(__byte_perm(xa, k23, 0x7632)) < fmaf(rk.y, tau16, c)) |
           (__uint_as_float(__byte_perm(xb, k23, 0x7610)) < fmaf(rk.z, tau16, c)) |
           (__uint_as_float(__byte_perm(xb, k23, 0x7632)) < fmaf(rk.w, tau16, c));
}
// The exact trial for the susceptibles of the quad (state word w): generates the low halves; returns the hit mask.
__device__ __noinline__ uint32_t exact_quad(const PassParams &pp, uint32_t c0, uint32_t c1, int par, uint32_t xa, uint32_t xb, uint32_t w,
                                            float4 rk, float tau) {
    uint32_t l[4];
    philox4x32_10(c0, c1, (uint32_t)(pp.A.tick - 1), LPK_STAGE_EXPOSE_LO, (uint32_t)pp.A.seed, (uint32_t)(pp.A.seed >> 32), l);
    const uint32_t la = par ? l[2] : l[0], lb = par ? l[3] : l[1];
    const uint32_t X[4] = {(xa << 16) | (la & 0xFFFFu), (xa & 0xFFFF0000u) | (la >> 16), (xb << 16) | (lb & 0xFFFFu),
                           (xb & 0xFFFF0000u) | (lb >> 16)};
    const float r[4] = {rk.x, rk.y, rk.z, rk.w};
    const uint32_t mS = mask_S(w);
    uint32_t hits = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (((mS >> (8 * k)) & 1u) && expose_test(p_expose(__fmul_rn(r[k], tau)), X[k])) hits |= 1u << (8 * k);
    return hits;
}

// per-agent form of the pre-test: bit 0 of byte k set when agent k of the quad passes it
__device__ __forceinline__ uint32_t pretest_mask(uint32_t xa, uint32_t xb, const float4 &rk, float tau16) {
    const uint32_t k23 = 0x4B000000u;
    const float c = 8388609.0f;
    return ((__uint_as_float(__byte_perm(xa, k23, 0x7610)) < fmaf(rk.x, tau16, c)) ? 1u : 0u) |
           ((__uint_as_float(__byte_perm(xa, k23, 0x7632)) < fmaf(rk.y, tau16, c)) ? 0x100u : 0u) |
           ((__uint_as_float(__byte_perm(xb, k23, 0x7610)) < fmaf(rk.z, tau16, c)) ? 0x10000u : 0u) |
           ((__uint_as_float(__byte_perm(xb, k23, 0x7632)) < fmaf(rk.w, tau16, c)) ? 0x1000000u : 0u);
}
// The exact exposure trial of tick t-1 for ONE susceptible agent (both Philox blocks regenerated): the streaming loop only
// pre-tests and sends the few candidates to the ring, where 32 of them are decided at a time with all lanes busy
// (profiles/r1_fused_v17_late_*: run in place it was 1-2 live lanes in 40 % of the iterations).
__device__ __noinline__ bool exact_agent(const PassParams &pp, int64_t i, int nd, float rk) {
    const lpk_tick_args &A = pp.A;
    const float tau = __ldg(&A.q_prev[nd]);
    if (!(tau > 0.f)) return false;
    const uint64_t id = (uint64_t)i + A.id_base;
    const uint64_t c = expose_ctr(id);
    const int hw = expose_hw(id);
    uint32_t h[4], l[4];
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(A.tick - 1), LPK_STAGE_EXPOSE, (uint32_t)A.seed, (uint32_t)(A.seed >> 32), h);
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(A.tick - 1), LPK_STAGE_EXPOSE_LO, (uint32_t)A.seed, (uint32_t)(A.seed >> 32), l);
    const uint32_t hs = (hw & 2) ? ((hw & 4) ? h[3] : h[1]) : ((hw & 4) ? h[2] : h[0]);
    const uint32_t ls = (hw & 2) ? ((hw & 4) ? l[3] : l[1]) : ((hw & 4) ? l[2] : l[0]);
    const uint32_t sh = 16u * (uint32_t)(hw & 1);
    const uint32_t X = (((hs >> sh) & 0xFFFFu) << 16) | ((ls >> sh) & 0xFFFFu);
    return expose_test(p_expose(__fmul_rn(rk, tau)), X);
}

// ---- active-agent queue -------------------------------------------------------------------------------------
// E / I agents (and fresh exposure hits) are appended to the warp's shared-memory ring and, whenever 32 have
// accumulated, processed one per lane with all lanes busy.  An entry carries everything the handler needs; the handler
// owns the agent's state byte from then on (the owning lane already stored the quad's word; both stores come from the
// same warp, ordered by the warp-wide reduction in q_commit).
// entry = {agent index (tables hold < 2^32 slots), node | state << 16 | hit << 20}
struct ActiveRegs {
    uint2 e;
    int8_t ipvv;
    float inf, rk;
};
struct WarpAcc;
struct WarpQueue {
    uint2 *q;
    uint32_t *tail;  // shared, monotonic
    uint32_t head;   // warp-uniform
    int count;       // warp-uniform
    bool loaded;     // warp-uniform: `pend` holds a batch whose loads are in flight
    ActiveRegs pend;
    WarpAcc *acc;
};

// census (rows t-1) -> disease state (tick t) -> infectivity tally (tick t) for one active agent, in two phases a ring
// batch apart: active_load issues every load the agent can need (one round trip, nothing dependent), active_process runs
// the state machine on those registers when the NEXT batch is loaded, so the scattered-load latency is spent streaming
// (profiles/r1_fused_v10_postsia_*: 35 % of the stall samples sat in the handler waiting for its own loads).  Between the
// two phases nobody else touches the agent: it is pushed once per pass, and the death / RI paths never push what they
// handle themselves.
// Ring entry: x = agent index, y = node | F << 16 | strain << 24 | exposure candidate << 26 with the flag byte
//   F = state before tick t's disease-state step (bits 0-1) | E -> I today << 2 | I -> R today << 3 | exposure hit of t-1 << 4
//       | RI-eligible << 5 | SIA-eligible << 6 | paralysis gate fires today << 7
// The streaming loop has already counted the timers down and written the new state (ds_quad): the handler only does what
// needs scattered columns -- class-change bookkeeping, the paralysis gate, the strain pick of a hit, the vaccine draws.
#define EF_TE (1u << 18)
#define EF_TI (1u << 19)
#define EF_HIT (1u << 20)
#define EF_RI (1u << 21)
#define EF_SIA (1u << 22)
#define EF_GATE (1u << 23)
#define EF_CAND (1u << 26)  // susceptible that passed the pre-test of tick t-1's exposure trial: the handler decides
__device__ __forceinline__ ActiveRegs active_load(const PassParams &pp, uint2 e) {
    const lpk_people &P = pp.P;
    const int64_t i = (int64_t)e.x;
    ActiveRegs r;
    r.e = e;
    r.ipvv = (e.y & EF_GATE) ? P.ipv_protected[i] : (int8_t)0;
    r.inf = (e.y & (EF_TE | EF_TI)) ? P.daily_infectivity[i] : 0.f;
    // a hit left S; a susceptible that is here for RI / SIA may leave it
    r.rk = ((e.y & EF_HIT) || ((e.y >> 16) & 3u) == 0u) ? P.acq_risk_multiplier[i] : 0.f;
    return r;
}
// Per-warp accumulators of the handler's node-level counts (shared memory; lanes add with shared-memory atomics, lane 0
// flushes).  The agents of a warp come from one node for many batches and all SMs work in few nodes at a time, so
// per-agent global atomics land on a handful of addresses from every SM at once and serialise in L2 (diag_v11: a tighter
// work window made the post-SIA pass 2x slower).  Counts are collected here and flushed with one global atomic per
// counter when the warp's node changes.
struct WarpAcc {
    int node;
    int E[LPK_MAX_STRAINS], I[LPK_MAX_STRAINS];  // changes of the carried exposed / infectious counts
    int H[LPK_MAX_STRAINS];                      // exposure hits of t-1 per strain (already included in E)
    int R;                                       // recoveries of tick t
    int riV, riP, ipvV, siaV, siaP;              // RI / SIA of tick t: vaccinated, protected (S -> E)
    long long beta[LPK_MAX_STRAINS];             // change of the carried infectivity tally (fixed point)
    long long expo;                              // risk (fixed point) of the agents that left S
};
__device__ __forceinline__ void acc_clear(WarpAcc *a) {
#pragma unroll
    for (int s = 0; s < LPK_MAX_STRAINS; ++s) { a->E[s] = 0; a->I[s] = 0; a->H[s] = 0; a->beta[s] = 0; }
    a->R = 0;
    a->riV = a->riP = a->ipvV = a->siaV = a->siaP = 0;
    a->expo = 0;
}
// lane 0 only
__device__ __noinline__ void acc_flush(const PassParams &pp, WarpAcc *a) {
    const lpk_tick_args &A = pp.A;
    const int nd = a->node, ns = A.n_strains;
    if (nd < 0) return;
    int hits = 0;
    if (a->riP) a->E[A.ri_strain] += a->riP;
    if (a->siaP) a->E[A.sia_strain] += a->siaP;
    for (int s = 0; s < ns; ++s) {
        const int64_t c = (int64_t)nd * ns + s;
        red_add(&A.E_cur[c], a->E[s]);
        red_add(&A.I_cur[c], a->I[s]);
        red_add(&A.new_exposed_by_strain_prev[c], a->H[s]);
        red_add(&A.tx_hits_by_strain[c], a->H[s]);
        red_add(&A.beta_fx[c], a->beta[s]);
        hits += a->H[s];
    }
    if (hits) {
        atomicAdd(&A.new_exposed_prev[nd], hits);
        atomicAdd(&A.tx_hits[nd], hits);
    }
    if (a->riV) atomicAdd(&A.ri_vaccinated[nd], a->riV);
    if (a->ipvV) atomicAdd(&A.ipv_vaccinated[nd], a->ipvV);
    if (a->riP) {
        const int64_t c = (int64_t)nd * ns + A.ri_strain;
        atomicAdd(&A.ri_protected[nd], a->riP);
        atomicAdd(&A.new_exposed[nd], a->riP);
        atomicAdd(&A.new_exposed_by_strain[c], a->riP);
        atomicAdd(&A.ri_new_exposed_by_strain[c], a->riP);
    }
    if (a->siaV) atomicAdd(&A.sia_vaccinated[nd], a->siaV);
    if (a->siaP) {
        const int64_t c = (int64_t)nd * ns + A.sia_strain;
        atomicAdd(&A.sia_protected[nd], a->siaP);
        atomicAdd(&A.new_exposed[nd], a->siaP);
        atomicAdd(&A.new_exposed_by_strain[c], a->siaP);
        atomicAdd(&A.sia_new_exposed_by_strain[c], a->siaP);
    }
    red_add(&A.sus[nd], -(long long)(hits + a->riP + a->siaP));
    red_add(&A.exposure_fx[nd], -a->expo);
    red_add(&A.R_cur[nd], a->R);
    acc_clear(a);
}
__device__ __forceinline__ void acc_add(long long *p, long long v) { atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v); }

// the rare draws of the handler, out of line so that the common path stays compact
__device__ __noinline__ int8_t pick_strain(const PassParams &pp, int64_t i, int nd) {  // model.py:1127-1141
    const lpk_tick_args &A = pp.A;
    uint32_t y[4];
    philox_agent(A.seed, (uint64_t)i + A.id_base, (uint32_t)(A.tick - 1), LPK_STAGE_STRAIN, y);
    const double u = u53(y[0], y[1]);
    const int ns = A.n_strains;
    for (int k = 0; k < ns; ++k)
        if (u < A.cdf_prev[(int64_t)nd * ns + k]) return (int8_t)k;
    return 0;
}
// routine immunisation (model.py:1825-1854) and the campaign (model.py:2030-2059) for one agent in state s (after this
// tick's disease-state step); returns s | vx << 8, vx bit 0 RI vaccinated, 1 RI protected, 2 IPV vaccinated, 3 SIA
// vaccinated, 4 SIA protected
__device__ __noinline__ uint32_t vaccine_draws(const PassParams &pp, int64_t i, int nd, int8_t s, uint32_t ey) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    uint32_t vx = 0u;
    if ((ey >> 21) & 1u) {
        uint32_t x[4];
        philox_agent(A.seed, (uint64_t)i + A.id_base, (uint32_t)A.tick, LPK_STAGE_RI, x);
        if (u53(x[0], x[1]) < A.vx_prob_ri[nd]) {
            vx |= 1u;
            if (s == 0) { s = 1; P.strain[i] = (int8_t)A.ri_strain; vx |= 2u; }
        }
        if (u53(x[2], x[3]) < A.vx_prob_ipv[nd]) { vx |= 4u; P.ipv_protected[i] = 1; }
    }
    if ((ey >> 22) & 1u) {
        uint32_t x[4];
        philox_agent(A.seed, (uint64_t)i + A.id_base, (uint32_t)A.tick, LPK_STAGE_SIA | (A.sia_event_idx << 8), x);
        const double u = u53(x[0], x[1]), pv = (double)A.vx_prob_sia[nd];
        if (u < pv) {
            vx |= 8u;
            if (s == 0 && u < pv * A.sia_vx_eff) { s = 1; P.strain[i] = (int8_t)A.sia_strain; vx |= 16u; }
        }
    }
    return (uint32_t)(uint8_t)s | (vx << 8);
}

// paralysis of one agent of the paralytic strain in the infected block (out of line: rare).  gate_only: the streaming loop
// found the gate open (timer run out, potentially_paralyzed still -1) and has counted the timer down itself; else the whole
// step on the agent's columns (an agent exposed yesterday that is infectious today: its strain was not known there).
__device__ __noinline__ void paralysis_agent(const PassParams &pp, int64_t i, int nd, int8_t ipvv, bool gate_only) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    int8_t pq = -1, par = 0;
    int flags = 0;
    if (gate_only) {
        paralysis_gate(i, ipvv, pq, par, (double)A.p_paralysis, stage_rng(pp), flags);
        P.potentially_paralyzed[i] = pq;
    } else {
        int8_t pt = P.paralysis_timer[i];
        const int8_t pq0 = pq = P.potentially_paralyzed[i];
        paralysis_step(i, P.ipv_protected[i], pt, pq, par, (double)A.p_paralysis, stage_rng(pp), flags);
        P.paralysis_timer[i] = pt;
        if (pq != pq0) P.potentially_paralyzed[i] = pq;
    }
    if (flags) {
        atomicAdd(&A.new_potential[nd], 1);
        if (flags & 2) { P.paralyzed[i] = 1; atomicAdd(&A.new_paralyzed[nd], 1); }
    }
}

// warp-collective: every lane calls it; `valid` lanes carry an agent
__device__ __noinline__ void active_process(const PassParams &pp, ActiveRegs r, bool valid, WarpAcc *acc, int lane) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const int64_t i = (int64_t)r.e.x;
    uint32_t ey = valid ? r.e.y : 0u;
    const int nd = (int)(int16_t)(ey & 0xFFFFu);
    if (ey & EF_CAND) {  // a susceptible that passed the pre-test of tick t-1's exposure trial
        if (exact_agent(pp, i, nd, r.rk)) {
            // exposed yesterday: the streaming loop saw a susceptible, so today's disease-state step (model.py:419-431) runs here
            const int8_t et = P.exposure_timer[i];
            uint32_t f = EF_HIT | (1u << 16);
            if (et <= 0) {
                const int8_t it = P.infection_timer[i];
                f |= EF_TE | (it <= 0 ? EF_TI : 0u);
                P.infection_timer[i] = (int8_t)(it - 1);
                r.inf = P.daily_infectivity[i];
            }
            P.exposure_timer[i] = (int8_t)(et - 1);
            ey |= f;
            P.disease_state[i] = (int8_t)(1 + ((f >> 18) & 1u) + ((f >> 19) & 1u));
        }
    }
    const int8_t s0 = (int8_t)((ey >> 16) & 3u);
    const int8_t sd = (int8_t)(s0 + ((ey >> 18) & 1u) + ((ey >> 19) & 1u));  // state after this tick's disease-state step
    const bool hit = (ey & EF_HIT) != 0u;
    int8_t st = (int8_t)((ey >> 24) & 3u), s = sd;
    long long efx = 0;
    uint32_t vx = 0u;
    if (valid) {
        if (hit) {  // exposure hit of tick t-1
            st = pick_strain(pp, i, nd);
            P.strain[i] = st;
            efx = __float2ll_rn(r.rk * 1073741824.0f);
            atomicAdd(&A.risk_hist[(int64_t)nd * LPK_RISK_BINS + risk_bin(r.rk)], -1);
            if (st == 0 && (ey & EF_TE)) paralysis_agent(pp, i, nd, 0, false);  // infectious on the day after exposure
        }
        if (ey & EF_GATE) paralysis_agent(pp, i, nd, r.ipvv, true);
        if (ey & (EF_RI | EF_SIA)) {  // after the disease-state step (the reference's run order; it used the ipv_protected loaded before)
            const uint32_t o = vaccine_draws(pp, i, nd, s, ey);
            s = (int8_t)(o & 0xFFu);
            vx = o >> 8;
            if (vx & 18u) {  // left S through a vaccine
                efx = __float2ll_rn(r.rk * 1073741824.0f);
                atomicAdd(&A.risk_hist[(int64_t)nd * LPK_RISK_BINS + risk_bin(r.rk)], -1);
                P.disease_state[i] = s;
            }
        }
    }
    // node-level counts, one group of same-node lanes at a time (one group except at a node boundary)
    uint32_t todo = __ballot_sync(LPK_FULL, valid);
    while (todo) {
        const int nd0 = __shfl_sync(LPK_FULL, nd, __ffs(todo) - 1);
        const bool mine = valid && nd == nd0;
        if (lane == 0 && acc->node != nd0) { acc_flush(pp, acc); acc->node = nd0; }
        __syncwarp();
        if (mine) {
            if (hit) { atomicAdd(&acc->H[st], 1); atomicAdd(&acc->E[st], 1); }
            if (sd != s0) {  // s0 is E or I here
                const long long fx = to_fx((double)r.inf * A.strain_r0_scalars[st]);
                if (s0 == 1) atomicAdd(&acc->E[st], -1);
                else { atomicAdd(&acc->I[st], -1); acc_add(&acc->beta[st], -fx); }
                if (sd == 2) { atomicAdd(&acc->I[st], 1); acc_add(&acc->beta[st], fx); }
                else atomicAdd(&acc->R, 1);
            }
            if (efx) acc_add(&acc->expo, efx);
            if (vx) {
                if (vx & 1u) atomicAdd(&acc->riV, 1);
                if (vx & 2u) atomicAdd(&acc->riP, 1);
                if (vx & 4u) atomicAdd(&acc->ipvV, 1);
                if (vx & 8u) atomicAdd(&acc->siaV, 1);
                if (vx & 16u) atomicAdd(&acc->siaP, 1);
            }
        }
        __syncwarp();
        todo &= ~__ballot_sync(LPK_FULL, mine);
    }
}

// Disease-state step of a node-uniform quad on byte lanes (reference model.py:419-431): every exposed agent's timer
// counts down and those at <= 0 turn infectious; every agent in the infected block (infectious before, or just turned)
// counts its timer down and those at <= 0 recover.  nw0: state word after tick t-1's hits and tick t's deaths; et / it /
// sw / pt / pq: the quad's exposure_timer, infection_timer, strain, paralysis_timer and potentially_paralyzed words.  Timers
// are written back here.
struct DsQuad {
    uint32_t nw, f, m;  // new state word; flag bytes of the agents that need the handler; their mask (bit 0 per byte)
};
// byte lanes by hand (the __v*4 intrinsics are emulated with ~10 instructions each): masks carry bit 0 of each byte
__device__ __forceinline__ uint32_t bytes_le0(uint32_t x) {  // per byte, signed: byte <= 0  (zero, or bit 7 set)
    const uint32_t nz = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;  // bit 7 set iff the byte is not zero
    return ((~nz | x) >> 7) & 0x01010101u;
}
__device__ __forceinline__ uint32_t bytes_dec(uint32_t x, uint32_t m) {  // per byte: x - m (m is 0 / 1), wrapping like int8
    const uint32_t t = (x | 0x80808080u) - m;  // bit 7 forced: no borrow leaves a byte
    return (t & 0x7F7F7F7Fu) | ((x ^ ~t) & 0x80808080u);
}
__device__ __forceinline__ DsQuad ds_quad(const lpk_people &P, int64_t b, uint32_t nw0, uint32_t hits, uint32_t et, uint32_t it, uint32_t sw,
                                          uint32_t pt, uint32_t pq) {
    const uint32_t K1 = 0x01010101u;
    const uint32_t mE = nw0 & ~(nw0 >> 1) & K1, mI = (nw0 >> 1) & ~nw0 & K1;  // state bytes are 0, 1, 2, 3 or 0xFF
    const uint32_t tE = mE & bytes_le0(et);
    const uint32_t mJ = mI | tE;
    const uint32_t tI = mJ & bytes_le0(it);
    if (mE) *reinterpret_cast<uint32_t *>(P.exposure_timer + b) = bytes_dec(et, mE);
    if (mJ) *reinterpret_cast<uint32_t *>(P.infection_timer + b) = bytes_dec(it, mJ);
    // the paralytic strain (0 of 0..3) in the infected block: its paralysis timer counts down here too, and the agents whose
    // gate opens today (timer run out, potentially_paralyzed still -1 = 0xFF) go to the handler.  A hit's strain is picked by
    // the handler, which then runs the whole step.
    const uint32_t wild = mJ & ~(sw | (sw >> 1)) & ~hits;
    uint32_t gate = 0u;
    if (wild) {
        gate = wild & bytes_le0(pt) & (pq >> 7);
        *reinterpret_cast<uint32_t *>(P.paralysis_timer + b) = bytes_dec(pt, wild);
    }
    DsQuad o;
    o.nw = nw0 + tE + tI;
    o.f = (nw0 & 0x03030303u) | (tE << 2) | (tI << 3) | (hits << 4) | (gate << 7);
    o.m = tE | tI | gate | hits;
    return o;
}
// The same step without a branch, for the streaming loop: the two quads a lane owns run through it back to back, so the
// compiler interleaves two independent dependency chains (the pass is bound by fixed-latency dependencies, not by issue
// slots: profiles/r1_fused_v19_220M_*: stall_wait 2.2 cycles per instruction at 4 warps per scheduler).  No hits here.
__device__ __forceinline__ DsQuad ds_quad_flat(const lpk_people &P, uint32_t b, uint32_t nw0, uint32_t et, uint32_t it, uint32_t sw,
                                               uint32_t pt, uint32_t pq) {
    const uint32_t K1 = 0x01010101u, K80 = 0x80808080u, K7F = 0x7F7F7F7Fu;
    const uint32_t mE = nw0 & ~(nw0 >> 1) & K1, mI = (nw0 >> 1) & ~nw0 & K1;
    // count down and test in one go: t = (x | 0x80) - m never borrows across bytes; where m = 1, bit 7 of t is clear iff the
    // low 7 bits of x were zero, so "x <= 0" (zero, or bit 7 set) is (~t | x) >> 7; the decremented byte keeps t's low 7 bits
    // and takes bit 7 from x ^ ~t
    const uint32_t te = (et | K80) - mE;
    const uint32_t tE = mE & ((~te | et) >> 7);
    const uint32_t mJ = mI | tE;
    const uint32_t ti = (it | K80) - mJ;
    const uint32_t tI = mJ & ((~ti | it) >> 7);
    const uint32_t wild = mJ & ~(sw | (sw >> 1));
    const uint32_t tp = (pt | K80) - wild;
    const uint32_t gate = wild & ((~tp | pt) >> 7) & (pq >> 7);
    if (mE) *reinterpret_cast<uint32_t *>(P.exposure_timer + b) = (te & K7F) | ((et ^ ~te) & K80);
    if (mJ) *reinterpret_cast<uint32_t *>(P.infection_timer + b) = (ti & K7F) | ((it ^ ~ti) & K80);
    if (wild) *reinterpret_cast<uint32_t *>(P.paralysis_timer + b) = (tp & K7F) | ((pt ^ ~tp) & K80);
    DsQuad o;
    o.nw = nw0 + tE + tI;
    o.f = nw0 | (tE << 2) | (tI << 3) | (gate << 7);  // flag bytes are read for the agents of o.m only (state byte 1 or 2)
    o.m = tE | tI | gate;
    return o;
}
// append the agents of mask m (bit 0 of byte k = agent idx0 + k); f = their flag bytes, g = their strain (bits 0-1) and
// exposure-candidate (bit 2) bytes; returns how many
__device__ __forceinline__ int q_push(uint2 *q, uint32_t *tail, uint32_t idx0, int nd, uint32_t f, uint32_t g, uint32_t m) {
    const int cnt = __popc(m);
    while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1u;
        const uint32_t pos = atomicAdd(tail, 1u) & (QCAP - 1);
        q[pos] = make_uint2(idx0 + (uint32_t)(bit >> 3), ((uint32_t)nd & 0xFFFFu) | (((f >> bit) & 0xFFu) << 16) | (((g >> bit) & 7u) << 24));
    }
    return cnt;
}
// the same for the two quads a lane owns in a row pair (B = A + 128 agents): one loop for both
__device__ __forceinline__ int q_push_pair(uint2 *q, uint32_t *tail, uint32_t idxA, int nd, uint32_t fA, uint32_t gA, uint32_t mA,
                                           uint32_t fB, uint32_t gB, uint32_t mB) {
    uint32_t m = mA | (mB << 4);
    const int cnt = __popc(m);
    while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1u;
        const bool rowB = (bit & 4) != 0;
        const uint32_t f = rowB ? fB : fA, g = rowB ? gB : gA;
        const uint32_t pos = atomicAdd(tail, 1u) & (QCAP - 1);
        q[pos] = make_uint2(idxA + (uint32_t)(bit >> 3) + (rowB ? 128u : 0u),
                            ((uint32_t)nd & 0xFFFFu) | (((f >> (bit & 24)) & 0xFFu) << 16) | (((g >> (bit & 24)) & 7u) << 24));
    }
    return cnt;
}
// every lane pushed `mine` entries: drain the ring 32 at a time (warp-uniform control flow)
__device__ __forceinline__ void q_commit(const PassParams &pp, WarpQueue &Q, int mine, int lane) {
    Q.count += __reduce_add_sync(LPK_FULL, mine);
    while (Q.count >= 32) {
        __syncwarp();
        if (pp.debug & 1u) { Q.head += 32; Q.count -= 32; continue; }
        // process the batch loaded a commit ago FIRST: a call boundary waits for every load in flight, so the new batch's
        // loads are issued after it and land while the warp streams the next pair (profiles/r1_fused_v12_*: with the
        // loads issued before the call, 8 % of all stall samples sat on the call instruction)
        if (Q.loaded) active_process(pp, Q.pend, true, Q.acc, lane);
        Q.pend = active_load(pp, Q.q[(Q.head + lane) & (QCAP - 1)]);
        Q.loaded = true;
        Q.head += 32;
        Q.count -= 32;
    }
}

// out-of-line part of a death in a node-uniform quad: the agents in mask dm die on tick t (after tick t-1's pending
// exposure); returns {new state word, remaining hits | remaining exposure candidates << 1}
__device__ __noinline__ uint2 death_quad(const PassParams &pp, int64_t b, int nd, uint32_t nw, uint32_t hits, uint32_t cand, uint32_t dm) {
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const uint32_t bit = 1u << (8 * k);
        if (!(dm & bit)) continue;
        int8_t s = byte_of(nw, k);
        if (cand & bit) {  // the exposure trial of t-1 comes before the death of t: decide it now
            cand &= ~bit;
            if (exact_agent(pp, b + k, nd, pp.P.acq_risk_multiplier[b + k])) { hits |= bit; s = 1; }
        }
        if (hits & bit) { expose_agent(pp, b + k, nd); hits &= ~bit; }
        kill_agent(pp, b + k, nd, s);
        nw = set_byte(nw, k, -1);
    }
    return make_uint2(nw, hits | (cand << 1));
}
__device__ __forceinline__ uint32_t death_mask(const int4 &d, int tick, uint32_t w) {
    return ((d.x <= tick ? 1u : 0u) | (d.y <= tick ? 0x100u : 0u) | (d.z <= tick ? 0x10000u : 0u) | (d.w <= tick ? 0x1000000u : 0u)) &
           mask_alive(w);
}

// ---- routine immunisation in a node-uniform quad (reference model.py:1825-1854) -----------------------------------
// Every alive, not chronically missed agent's ri_timer goes down by the step (four int16 lanes at a time); an agent is
// eligible when the new timer lies in (-step, 0] ([-step, 0] on the first RI tick).  Eligible agents are the few in
// the age window; they go to the ring and take their two draws in the handler, after their disease-state step.
__device__ __forceinline__ uint32_t ri_timers_quad(const PassParams &pp, int64_t b, uint32_t w, uint32_t missed, uint2 tm) {
    const int step = pp.A.ri_step;
    const uint32_t ok8 = mask_alive(w) & ~missed;  // missed bytes are 0 / 1
    if (!ok8) return 0u;
    const uint2 tn = make_uint2(__vsub2(tm.x, __byte_perm(ok8, 0u, 0x4140) * (uint32_t)step),
                                __vsub2(tm.y, __byte_perm(ok8, 0u, 0x4342) * (uint32_t)step));
    *reinterpret_cast<uint2 *>(pp.P.ri_timer + b) = tn;
    const int lo = (pp.A.tick == step) ? -step : 1 - step;  // eligible: lo <= timer <= 0
    const uint32_t lo2 = ((uint32_t)lo & 0xFFFFu) * 0x10001u, span2 = ((uint32_t)(-lo) & 0xFFFFu) * 0x10001u;
    const uint32_t ex = __vcmpleu2(__vsub2(tn.x, lo2), span2), ey = __vcmpleu2(__vsub2(tn.y, lo2), span2);
    return __byte_perm(ex, ey, 0x6420) & ok8;
}
// ------------------------------------------------------------------ general pair (out of line): 256 agents that are not
// all in one node or were not all present at tick t-1 (node boundaries, newborn cohorts, the table's tail).  One row at a
// time, no software pipeline.  Returns how many agents this lane appended to the ring.
template <bool kDeaths, bool kRI, bool kSIA>
__device__ __noinline__ int general_pair(const PassParams &pp, uint2 *q, uint32_t *q_tail, int64_t gp, int64_t n, int64_t count_prev,
                                         int lane) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const bool pending = (A.flags & LPK_F_PENDING) != 0;
    const int tick = A.tick;
    int mine = 0;
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        const int64_t b = gp * 256 + r * 128 + lane * 4;
        const int valid = quad_valid(b, n);
        if (!valid) continue;
        const uint32_t w = load_b4(P.disease_state, b, valid);
        if ((w & 0x80808080u) == 0x80808080u) continue;  // nobody alive
        uint32_t nw = w, hits = 0u, cand = 0u, elig = 0u, camp = 0u, fl = 0u, sw = 0u;
        int nd = -1;
        bool fast = (valid == 4) && (!pending || b + 4 <= count_prev);
        if (fast) {
            const uint2 ids = *reinterpret_cast<const uint2 *>(P.node_id + b);
            nd = (int)(int16_t)(ids.x & 0xFFFFu);
            fast = ids.x == ids.y && (ids.x >> 16) == (ids.x & 0xFFFFu) && nd >= 0;
        }
        if (!fast) {
            nw = slow_quad(pp, b, valid, w, kDeaths, kRI, count_prev);
        } else {
            if (pending && mask_S(w)) {  // exposure trial of tick t-1
                const float tau = __ldg(&A.q_prev[nd]);
                if (tau > 0.f) {
                    const uint64_t id0 = (uint64_t)b + A.id_base;
                    const uint64_t c = expose_ctr(id0);
                    const int par = (int)((id0 >> 7) & 1u);
                    uint32_t x[4];
                    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(tick - 1), LPK_STAGE_EXPOSE, (uint32_t)A.seed,
                                  (uint32_t)(A.seed >> 32), x);
                    const float4 rk = __ldg(reinterpret_cast<const float4 *>(P.acq_risk_multiplier + b));
                    // the 16-bit pre-test first, as in the streaming loop: newborn cohorts are all susceptible and never leave
                    // this path (their 512-agent tiles span several nodes), and the exact trial is ~200 instructions
                    const uint32_t xa = par ? x[2] : x[0], xb = par ? x[3] : x[1];
                    if (pretest_quad(xa, xb, rk, tau * 65536.0f)) hits = exact_quad(pp, (uint32_t)c, (uint32_t)(c >> 32), par, xa, xb, w, rk, tau);
                    nw |= hits;  // S (0) -> E (1)
                }
            }
            if (kDeaths) {
                const int4 dd = __ldg(reinterpret_cast<const int4 *>(P.date_of_death + b));
                const uint32_t dm = death_mask(dd, tick, nw);
                if (dm) { const uint2 o = death_quad(pp, b, nd, nw, hits, 0u, dm); nw = o.x; hits = o.y; }
            }
            if (mask_EI(nw)) {  // disease-state step of tick t on byte lanes
                sw = *reinterpret_cast<const uint32_t *>(P.strain + b);
                const DsQuad d = ds_quad(P, b, nw, hits, *reinterpret_cast<const uint32_t *>(P.exposure_timer + b),
                                         *reinterpret_cast<const uint32_t *>(P.infection_timer + b), sw,
                                         *reinterpret_cast<const uint32_t *>(P.paralysis_timer + b),
                                         *reinterpret_cast<const uint32_t *>(P.potentially_paralyzed + b));
                nw = d.nw; fl = d.f; cand = d.m;
            } else {
                fl = (nw & 0x03030303u) | (hits << 4);  // no E / I in the quad: no hit either
            }
            uint32_t missed = 0u;
            if (kRI || kSIA) missed = *reinterpret_cast<const uint32_t *>(P.chronically_missed + b);
            if (kRI) elig = ri_timers_quad(pp, b, nw, missed, *reinterpret_cast<const uint2 *>(P.ri_timer + b));
            if (kSIA && A.sia_targeted[nd])
                camp = sia_age_mask(__ldg(reinterpret_cast<const int4 *>(P.date_of_birth + b)), tick, A.sia_min_age,
                                    (uint32_t)(A.sia_max_age - A.sia_min_age)) & mask_alive(nw) & ~missed;
        }
        if (nw != w) store_b4(P.disease_state, b, valid, nw);
        mine += q_push(q, q_tail, (uint32_t)b, nd, fl | (elig << 5) | (camp << 6), sw, cand | elig | camp);
    }
    return mine;
}

// ------------------------------------------------------------------ the pass
// Unit of work: a PAIR of 128-agent rows (256 consecutive agents); a lane owns its quad in the even row (A) and in the
// odd row (B).  Warps claim units of LPK_UNIT_PAIRS consecutive pairs from a global counter (see the kernel body).  A
// warp's pairs form one sequence s = 0, 1, ... served by the warp's PRIVATE ring of kStages
// shared-memory slots: one elected lane asks the TMA engine for the pair's columns (cp.async.bulk: 256 B of state, 1 KB
// of risk, + date_of_death / chronically_missed / ri_timer on vital-dynamics / RI ticks) kStages iterations ahead, the
// bytes land on the slot's mbarrier, and the warp reads its quads from shared memory.  The copies cost no registers and
// no per-lane load instructions, so the depth of the memory pipeline is set by shared memory (40 KB per block), not by
// occupancy.  The pair's node (tile table) and the node's exposure scale tau are looked up at issue time: with no force
// of infection on the node neither risk nor random numbers are touched.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 26)) __trap();  // a lost copy would otherwise hang the device; this turns it into an error
}
__device__ __forceinline__ void tma_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// [6 columns, 256 agents] of the byte-column tensor starting at agent a0 -> 1536 contiguous bytes
__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap *map, uint32_t a0, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(a0), "r"(0), "r"(bar) : "memory");
}
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0u;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <bool kDeaths, bool kRI, bool kSIA, int kWarps, int kOcc>
struct PassSmem {
    // a stage: state 256 | exposure_timer 256 | infection_timer 256 | strain 256 | paralysis_timer 256 | potentially_paralyzed 256
    //          (= the [6, 256] box of the tensor copy) | risk 1024 | [date_of_death 1024] | [chronically_missed 256] | [ri_timer 512]
    //          | [date_of_birth 1024]
    static constexpr int kOffEt = 256, kOffIt = 512, kOffSt = 768, kOffPt = 1024, kOffPq = 1280, kOffRisk = 1536, kOffDod = 2560;
    static constexpr int kOffMissed = kOffDod + (kDeaths ? 1024 : 0), kOffTimer = kOffMissed + 256;
    static constexpr int kOffDob = kOffMissed + ((kRI || kSIA) ? 256 : 0) + (kRI ? 512 : 0);
    static constexpr int kStageBytes = kOffDob + (kSIA ? 1024 : 0);
    static constexpr int kFit = (227 * 1024 / kOcc - 1024 - kWarps * QCAP * 8 - 2048) / (kWarps * kStageBytes);  // kOcc blocks per SM
    static constexpr int kStages = kFit >= 4 ? 4 : (kFit < 1 ? 1 : kFit);
    static constexpr int kOffQueue = 0;
    static constexpr int kOffSlots = kOffQueue + kWarps * QCAP * 8;
    static constexpr int kOffBars = kOffSlots + kWarps * kStages * kStageBytes;
    static constexpr int kOffMeta = (kOffBars + kWarps * kStages * 8 + 15) & ~15;
    static constexpr int kOffTail = kOffMeta + kWarps * kStages * 16;
    static constexpr int kOffAcc = (kOffTail + kWarps * 4 + 15) & ~15;
    static constexpr int kBytes = kOffAcc + kWarps * (int)sizeof(WarpAcc) + 32;
    static_assert(kStages + 1 <= LPK_UNIT_PAIRS, "the producer may not run further ahead than one work unit");
    static_assert(kOcc * (kBytes + 1024) <= 228 * 1024, "kOcc blocks per SM");
};

template <bool kDeaths, bool kRI, bool kSIA, int kWarps, int kOcc>
__global__ void __launch_bounds__(kWarps * 32, kOcc) k_tick_pass(const __grid_constant__ PassParams pp) {
    typedef PassSmem<kDeaths, kRI, kSIA, kWarps, kOcc> L;
    constexpr int NST = L::kStages;
    extern __shared__ __align__(128) unsigned char smem[];
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t count_prev = A.counts[0], n = A.counts[1];
    const bool pending = (A.flags & LPK_F_PENDING) != 0;
    const int tick = A.tick;
    const uint32_t k0 = (uint32_t)A.seed, k1 = (uint32_t)(A.seed >> 32);
    const uint32_t total_pairs = (uint32_t)((n + 255) >> 8);
    const uint32_t full_pairs = P.tile_node ? (uint32_t)(count_prev >> 8) : 0u;  // pairs whose 256 agents all existed at tick t-1
    const uint32_t n_units = (total_pairs + LPK_UNIT_PAIRS - 1) >> LPK_UNIT_LOG;

    unsigned char *slots = smem + L::kOffSlots + warp * NST * L::kStageBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::kOffBars) + warp * NST;
    int4 *meta = reinterpret_cast<int4 *>(smem + L::kOffMeta) + warp * NST;
    WarpQueue Q;
    Q.q = reinterpret_cast<uint2 *>(smem + L::kOffQueue) + warp * QCAP;
    Q.tail = reinterpret_cast<uint32_t *>(smem + L::kOffTail) + warp;
    Q.head = 0u;
    Q.count = 0;
    Q.loaded = false;
    Q.acc = reinterpret_cast<WarpAcc *>(smem + L::kOffAcc) + warp;
    if (lane == 0) {
        Q.acc->node = -1;
        acc_clear(Q.acc);
        *Q.tail = 0u;
        for (int k = 0; k < NST; ++k) mbar_init(&bars[k], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // Work distribution: a warp claims RUNS of consecutive units (LPK_UNIT_PAIRS pairs each) from a global counter, so a
    // warp that meets regions dense in E / I agents (an SIA wave hits whole nodes) simply claims less
    // (profiles/r1_fused_v10_*: with static round-robin chunks the average SM was busy 58-66 % of the kernel's duration).
    // Guided self-scheduling: a run is 1 / (3 x warps in the grid) of the units still unclaimed (at most 64, at least 1) --
    // long runs while there is plenty of work, so that a warp stays in one node and its per-node accumulators are flushed
    // rarely, single units at the end for balance.
    // A warp's pairs form one sequence s = 0, 1, ...; the unit of sequence position s sits in register ua / ub (parity of
    // s >> LPK_UNIT_LOG); the producer runs at most kStages + 1 <= LPK_UNIT_PAIRS positions ahead of the consumer, so two
    // registers suffice.  The next run is claimed when the last unit of the current one is taken, a unit's worth of time
    // before it is needed, from a counter value read another unit earlier: no claim latency is ever waited for.
    const uint32_t kNoUnit = 0xFFFFFFFFu;
    const uint32_t guide = 3u * gridDim.x * kWarps;
    uint32_t ua = kNoUnit, ub = kNoUnit;
    uint32_t run_next = 0u, run_end = 0u;  // warp-uniform: the units of the current run not yet taken
    bool exhausted = false;                // warp-uniform: a claim came back beyond the last unit
    uint32_t claim_first = 0u, claim_cnt = 0u, seen = 0u;  // lane 0: the prefetched claim, the counter as last read
    const uint32_t stream_units = A.uniform_agents > 0 ? (uint32_t)(A.uniform_agents >> (8 + LPK_UNIT_LOG)) : n_units;
    const uint32_t tail_units = n_units - (stream_units < n_units ? stream_units : n_units);  // warp-uniform
    auto run_length = [&](uint32_t ctr) -> uint32_t {
        if (ctr < tail_units) return 1u;
        const uint32_t left = ctr < n_units ? n_units - ctr : 0u;
        const uint32_t r = left / guide;
        return r < 1u ? 1u : (r > 64u ? 64u : r);
    };
    if (lane == 0) {
        claim_cnt = run_length(0u);
        claim_first = atomicAdd(pp.unit_ctr, claim_cnt);
    }
    uint32_t gp_looked = 0u;  // pair index of the position node_of looked up last
    auto pair_of = [&](int s) -> uint32_t {
        const uint32_t u = ((s >> LPK_UNIT_LOG) & 1) ? ub : ua;
        return u == kNoUnit ? kNoUnit : (u << LPK_UNIT_LOG) + (uint32_t)(s & (LPK_UNIT_PAIRS - 1));
    };
    // node of the pair at position s (called once per s, in order): >= 0 all 256 agents in that node and present at tick
    // t-1; -1 general handling; -2 no more work for this warp; -3 no such pair (past the end of the table)
    auto node_of = [&](int s) -> int {
        if ((s & (LPK_UNIT_PAIRS - 1)) == 0) {
            uint32_t u = kNoUnit;
            if (!exhausted) {
                if (run_next >= run_end) {  // take the prefetched claim
                    const uint32_t first = __shfl_sync(LPK_FULL, claim_first, 0), cnt = __shfl_sync(LPK_FULL, claim_cnt, 0);
                    if (first >= n_units) exhausted = true;
                    else { run_next = first; run_end = first + cnt < n_units ? first + cnt : n_units; }
                }
                if (!exhausted) {
                    // claim index -> unit: the units behind the node-contiguous initial population (appended newborn cohorts:
                    // their tiles span several nodes, so they run through the general path, ~50 us of a warp's time per unit)
                    // are handed out FIRST, one per claim; claimed last they formed a tail of that length on every day
                    // after the first births (plain day 0.54 -> 0.59 ms).  The streaming part follows in ascending order.
                    const uint32_t c = run_next++;
                    u = c < tail_units ? n_units - 1u - c : c - tail_units;
                    if (run_next >= run_end && lane == 0) {  // the run's last unit: claim the next run now
                        claim_cnt = run_length(seen);
                        claim_first = atomicAdd(pp.unit_ctr, claim_cnt);
                    }
                    if (lane == 0) seen = *reinterpret_cast<volatile uint32_t *>(pp.unit_ctr);
                }
            }
            if ((s >> LPK_UNIT_LOG) & 1) ub = u; else ua = u;
        }
        const uint32_t gp = pair_of(s);
        gp_looked = gp;
        if (gp == kNoUnit) return -2;
        if (gp >= total_pairs) return -3;  // beyond the table inside the last unit: nothing to do, but the warp goes on
        return gp < full_pairs ? __ldg(&P.tile_node[gp >> 1]) : -1;
    };
    int tc_node = -2;  // one-entry cache of tau (and the campaign's target flag) per node: a warp stays in one node for long
    float tc_tau = 0.f;
    bool tc_sia = false;
    const int sia_lo = kSIA ? A.sia_min_age : 0;
    const uint32_t sia_span = kSIA ? (uint32_t)(A.sia_max_age - A.sia_min_age) : 0u;
    int tn_next = node_of(0);  // node (and pair index) of the next pair to be requested, looked up one request ahead
    uint32_t gp_next = gp_looked;
    // request pair s into slot (warp-uniform; the elected lane talks to the TMA engine)
    auto produce = [&](int s, int slot) {
        const int tn = tn_next;
        const uint32_t gp_req = gp_next;
        tn_next = node_of(s + 1);
        gp_next = gp_looked;
        float tau = 0.f;
        if (tn >= 0) {
            if (tn != tc_node) {
                tc_node = tn;
                tc_tau = pending ? __ldg(&A.q_prev[tn]) : 0.f;
                if (kSIA) tc_sia = __ldg(&A.sia_targeted[tn]) != 0;
            }
            tau = tc_tau;
        }
        const bool camp = kSIA && tn >= 0 && tc_sia;
        // What the copy engine needs goes through ONE warp reduction: its result lives in a uniform register, so the bulk
        // copies below are issued straight from the uniform datapath by the elected lane.  With the operands in ordinary
        // registers the compiler wraps every copy in an ELECT / R2UR x 4 / branch loop: 13 instructions per copy, 7-11
        // copies per pair -- a third of the plain day's instructions (SASS of v19).
        const uint32_t u = __reduce_max_sync(LPK_FULL, (gp_req & 0xFFFFFFu) | (tau > 0.f ? 1u << 24 : 0u) | (tn >= 0 ? 1u << 25 : 0u) |
                                                           (camp ? 1u << 26 : 0u) | ((uint32_t)slot << 28));
        if (elect_one()) {
            meta[slot] = make_int4(tn, __float_as_int(tau) | (camp ? (int)0x80000000u : 0), (int)gp_req, 0);  // tau >= 0: the sign bit is free
            const uint32_t uslot = u >> 28;
            const uint32_t bar = smem_u32(bars) + uslot * 8u;
            if (u & (1u << 25)) {
                const int64_t a0 = (int64_t)(u & 0xFFFFFFu) * 256;
                const uint32_t dst = smem_u32(slots) + uslot * (uint32_t)L::kStageBytes;
                const bool risk = (u & (1u << 24)) != 0u, ucamp = kSIA && (u & (1u << 26)) != 0u;
                fence_proxy_async_smem();  // the warp's reads of this slot (previous use) precede the engine's writes
                const bool missed = kRI || ucamp;
                mbar_arrive_expect_tx(bar, 1536u + (risk ? 1024u : 0u) + (kDeaths ? 1024u : 0u) + (missed ? 256u : 0u) + (kRI ? 512u : 0u) +
                                               (ucamp ? 1024u : 0u));
                if (pp.use_tmap) {
                    tma_load_box(dst, &pp.tmap, (u & 0xFFFFFFu) << 8, bar);
                } else {
                    tma_load(dst, P.disease_state + a0, 256u, bar);
                    tma_load(dst + L::kOffEt, P.exposure_timer + a0, 256u, bar);
                    tma_load(dst + L::kOffIt, P.infection_timer + a0, 256u, bar);
                    tma_load(dst + L::kOffSt, P.strain + a0, 256u, bar);
                    tma_load(dst + L::kOffPt, P.paralysis_timer + a0, 256u, bar);
                    tma_load(dst + L::kOffPq, P.potentially_paralyzed + a0, 256u, bar);
                }
                if (risk) tma_load(dst + L::kOffRisk, P.acq_risk_multiplier + a0, 1024u, bar);
                if (kDeaths) tma_load(dst + L::kOffDod, P.date_of_death + a0, 1024u, bar);
                if (missed) tma_load(dst + L::kOffMissed, P.chronically_missed + a0, 256u, bar);
                if (kRI) tma_load(dst + L::kOffTimer, P.ri_timer + a0, 512u, bar);
                if (ucamp) tma_load(dst + L::kOffDob, P.date_of_birth + a0, 1024u, bar);
            } else {
                mbar_arrive(bar);
            }
        }
    };

    // process pair s from slot, then re-arm the slot with pair s + NST
    auto consume = [&](int s, int slot, uint32_t parity) -> bool {
        mbar_wait(&bars[slot], parity);
        const int4 mt = meta[slot];
        const int tn = mt.x;
        const float tau = __int_as_float(mt.y & 0x7FFFFFFF);
        const bool camp = kSIA && mt.y < 0;
        const unsigned char *src = slots + slot * L::kStageBytes;
        uint32_t wA = 0u, wB = 0u;
        if (tn >= 0) {
            wA = *reinterpret_cast<const uint32_t *>(src + lane * 4);
            wB = *reinterpret_cast<const uint32_t *>(src + 128 + lane * 4);
        }
        if (tn < 0) {  // nothing was copied for this position
            __syncwarp();
            produce(s + NST, slot);
            if (tn == -2) return false;
            if (tn == -1) q_commit(pp, Q, general_pair<kDeaths, kRI, kSIA>(pp, Q.q, Q.tail, (int64_t)(uint32_t)mt.z, n, count_prev, lane), lane);
            return true;
        }
        const int64_t gp = (int64_t)(uint32_t)mt.z;
        const int nd = tn;
        const int64_t bA = gp * 256 + lane * 4, bB = bA + 128;
        uint32_t nwA = wA, nwB = wB, xcA = 0u, xcB = 0u;
        uint32_t fA = 0u, fB = 0u, gA = 0u, gB = 0u, cA = 0u, cB = 0u;
        // disease-state step of tick t on byte lanes, both rows back to back and without a branch, in the same basic block
        // as the Philox rounds of the exposure trial: the pass is bound by fixed-latency dependencies (stall_wait 2.2 cycles
        // per instruction at 4 warps per scheduler), and these are three independent chains.  Only class changes, the
        // paralytic strain's gate and vaccine-eligible agents go to the ring (their draws follow their own step).
        auto ds_both = [&]() {
            gA = *reinterpret_cast<const uint32_t *>(src + L::kOffSt + lane * 4);
            gB = *reinterpret_cast<const uint32_t *>(src + L::kOffSt + 128 + lane * 4);
            const DsQuad dA = ds_quad_flat(P, (uint32_t)bA, nwA, *reinterpret_cast<const uint32_t *>(src + L::kOffEt + lane * 4),
                                           *reinterpret_cast<const uint32_t *>(src + L::kOffIt + lane * 4), gA,
                                           *reinterpret_cast<const uint32_t *>(src + L::kOffPt + lane * 4),
                                           *reinterpret_cast<const uint32_t *>(src + L::kOffPq + lane * 4));
            const DsQuad dB = ds_quad_flat(P, (uint32_t)bB, nwB, *reinterpret_cast<const uint32_t *>(src + L::kOffEt + 128 + lane * 4),
                                           *reinterpret_cast<const uint32_t *>(src + L::kOffIt + 128 + lane * 4), gB,
                                           *reinterpret_cast<const uint32_t *>(src + L::kOffPt + 128 + lane * 4),
                                           *reinterpret_cast<const uint32_t *>(src + L::kOffPq + 128 + lane * 4));
            nwA = dA.nw; fA = dA.f; cA = dA.m;
            nwB = dB.nw; fB = dB.f; cB = dB.m;
        };
        if (tau > 0.f) {  // exposure trial of tick t-1: pre-test here, candidates decided by the ring handler
            const float4 rA = *reinterpret_cast<const float4 *>(src + L::kOffRisk + lane * 16);
            const float4 rB = *reinterpret_cast<const float4 *>(src + L::kOffRisk + 512 + lane * 16);
            const uint64_t c = (((uint64_t)gp + (A.id_base >> 8)) << 5) + (uint64_t)lane;
            uint32_t x[4];
            philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(tick - 1), LPK_STAGE_EXPOSE, k0, k1, x);
            const uint32_t sA0 = mask_S(nwA), sB0 = mask_S(nwB);  // susceptible at the end of tick t-1
            if (!kDeaths) ds_both();
            const float tau16 = tau * 65536.0f;
            if (pretest_quad(x[0], x[1], rA, tau16) | pretest_quad(x[2], x[3], rB, tau16)) {
                xcA = pretest_mask(x[0], x[1], rA, tau16) & sA0;
                xcB = pretest_mask(x[2], x[3], rB, tau16) & sB0;
            }
        } else if (!kDeaths) {
            ds_both();
        }
        if (kDeaths) {  // tick t
            const int4 dA = *reinterpret_cast<const int4 *>(src + L::kOffDod + lane * 16);
            const int4 dB = *reinterpret_cast<const int4 *>(src + L::kOffDod + 512 + lane * 16);
            const uint32_t dmA = death_mask(dA, tick, nwA), dmB = death_mask(dB, tick, nwB);
            if (dmA) { const uint2 o = death_quad(pp, bA, nd, nwA, 0u, xcA, dmA); nwA = o.x; xcA = (o.y >> 1) & 0x01010101u; }
            if (dmB) { const uint2 o = death_quad(pp, bB, nd, nwB, 0u, xcB, dmB); nwB = o.x; xcB = (o.y >> 1) & 0x01010101u; }
            ds_both();
        }
        uint32_t eA = 0u, eB = 0u, sA = 0u, sB = 0u;
        if (kRI || camp) {
            const uint32_t mA = *reinterpret_cast<const uint32_t *>(src + L::kOffMissed + lane * 4);
            const uint32_t mB = *reinterpret_cast<const uint32_t *>(src + L::kOffMissed + 128 + lane * 4);
            if (kRI) {
                eA = ri_timers_quad(pp, bA, nwA, mA, *reinterpret_cast<const uint2 *>(src + L::kOffTimer + lane * 8));
                eB = ri_timers_quad(pp, bB, nwB, mB, *reinterpret_cast<const uint2 *>(src + L::kOffTimer + 256 + lane * 8));
            }
            if (camp) {
                sA = sia_age_mask(*reinterpret_cast<const int4 *>(src + L::kOffDob + lane * 16), tick, sia_lo, sia_span) & mask_alive(nwA) & ~mA;
                sB = sia_age_mask(*reinterpret_cast<const int4 *>(src + L::kOffDob + 512 + lane * 16), tick, sia_lo, sia_span) & mask_alive(nwB) & ~mB;
            }
        }
        __syncwarp();
        produce(s + NST, slot);  // every read of the slot is done: re-arm it
        if (nwA != wA) *reinterpret_cast<uint32_t *>(P.disease_state + bA) = nwA;
        if (nwB != wB) *reinterpret_cast<uint32_t *>(P.disease_state + bB) = nwB;
        q_commit(pp, Q, q_push_pair(Q.q, Q.tail, (uint32_t)bA, nd, fA | (eA << 5) | (sA << 6), (gA & 0x03030303u) | (xcA << 2),
                                    cA | eA | sA | xcA, fB | (eB << 5) | (sB << 6), (gB & 0x03030303u) | (xcB << 2), cB | eB | sB | xcB), lane);
        return true;
    };

    // one copy of the loop body (runtime slot index): the unrolled variant was 100 KB of code and stalled on instruction
    // fetch (profiles/r1_fused_v10_*: no_instruction 8.4 per issue).  Once a position has no work none after it has, and
    // nothing was requested from the TMA engine for those, so the warp can leave at the first one.
    {
#pragma unroll 1
        for (int k = 0; k < NST; ++k) produce(k, k);
        uint32_t parity = 0u;
        int slot = 0;
#pragma unroll 1
        for (int s = 0;; ++s) {
            if (!consume(s, slot, parity)) break;
            if (++slot == NST) { slot = 0; parity ^= 1u; }
        }
    }
    __syncwarp();
    if (Q.loaded) active_process(pp, Q.pend, true, Q.acc, lane);
    {
        const bool valid = lane < Q.count;
        ActiveRegs last = {};
        if (valid) last = active_load(pp, Q.q[(Q.head + lane) & (QCAP - 1)]);
        active_process(pp, last, valid, Q.acc, lane);
    }
    if (lane == 0) acc_flush(pp, Q.acc);
}

template <bool kDeaths, bool kRI, bool kSIA, int kWarps, int kOcc>
static int launch_pass(const PassParams &pp, cudaStream_t st) {
    typedef PassSmem<kDeaths, kRI, kSIA, kWarps, kOcc> L;
    static bool configured = false;
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(k_tick_pass<kDeaths, kRI, kSIA, kWarps, kOcc>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes),
                 "tick_pass smem");
        configured = true;
    }
    const int grid = lpk_sm_count() * kOcc;
    CUDA_TRY(cudaMemsetAsync(pp.unit_ctr, 0, sizeof(uint32_t), st), "tick_pass work counter");
    k_tick_pass<kDeaths, kRI, kSIA, kWarps, kOcc><<<grid, kWarps * 32, L::kBytes, st>>>(pp);
    return LPK_OK;
}
// one 4-byte work counter per device, allocated on first use (the pass is launched on one stream per device)
static uint32_t *pass_unit_counter() {
    static uint32_t *ctr[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!ctr[dev] && cudaMalloc(&ctr[dev], 256) != cudaSuccess) ctr[dev] = nullptr;
    return ctr[dev];
}
// The byte columns of the disease state as one 2-D uint8 tensor [6 columns, capacity agents]: possible when the caller
// allocated them at one constant stride, in this order (device.DeviceState does); one tensor copy per pair then replaces six
// bulk copies.  Encoded once per (base, stride, capacity); any failure just leaves the per-column copies in charge.
typedef CUresult (*lpk_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool pass_tensor_map(const lpk_people &P, CUtensorMap *out) {
    static lpk_encode_tiled_fn encode = nullptr;
    static bool looked = false;
    static struct { const void *base; int64_t stride, capacity; CUtensorMap map; bool ok; } cache = {nullptr, 0, 0, {}, false};
    const char *env = getenv("LPK_PASS_TMAP");
    if (env && env[0] == '0') return false;
    const int8_t *cols[6] = {P.disease_state, P.exposure_timer, P.infection_timer, P.strain, P.paralysis_timer, P.potentially_paralyzed};
    const int64_t stride = cols[1] - cols[0];
    if (stride < P.capacity || (stride & 15) || P.capacity >= (1ll << 31)) return false;
    for (int k = 1; k < 6; ++k)
        if (cols[k] - cols[k - 1] != stride) return false;
    if (cache.base == cols[0] && cache.stride == stride && cache.capacity == P.capacity) {
        if (cache.ok) *out = cache.map;
        return cache.ok;
    }
    if (!looked) {
        looked = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<lpk_encode_tiled_fn>(fn);
        else
            (void)cudaGetLastError();
    }
    cache.base = cols[0]; cache.stride = stride; cache.capacity = P.capacity; cache.ok = false;
    if (!encode) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)P.capacity, 6}, strides[1] = {(cuuint64_t)stride};
    const cuuint32_t box[2] = {256, 6}, estr[2] = {1, 1};
    cache.ok = encode(&cache.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t *>(cols[0]), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    if (cache.ok) *out = cache.map;
    return cache.ok;
}

// Block shape: 2 blocks of 8 warps per SM for every variant.  3 x 6 warps (96 registers) and 2 x 10 warps were measured
// (-3 % / +1 %): the pass is bound by the ALU pipe, not by latency hiding.
extern "C" int lpk_tick_pass(const lpk_people *people, const lpk_tick_args *args, void *stream) {
    REQUIRE(people && args, "tick_pass null struct");
    const lpk_people &P = *people;
    const lpk_tick_args &A = *args;
    REQUIRE(A.n_nodes > 0 && A.n_strains >= 1 && A.n_strains <= LPK_MAX_STRAINS, "tick_pass sizes");
    REQUIRE(P.capacity > 0 && P.capacity < (1ll << 32) && A.counts, "tick_pass counts (tables hold < 2^32 slots)");
    REQUIRE(P.disease_state && P.strain && P.exposure_timer && P.infection_timer && P.paralysis_timer && P.potentially_paralyzed &&
                P.paralyzed && P.ipv_protected && P.node_id && P.acq_risk_multiplier && P.daily_infectivity, "tick_pass agent columns");
    REQUIRE(ALIGNED(P.disease_state, 4) && ALIGNED(P.node_id, 8) && ALIGNED(P.acq_risk_multiplier, 16), "tick_pass alignment");
    REQUIRE((A.flags & LPK_F_STAGES) != 0, "tick_pass always runs the stages of its tick (LPK_F_STAGES)");
    REQUIRE((A.id_base & 255) == 0, "tick_pass id_base must be a multiple of 256");
    REQUIRE(A.new_potential && A.new_paralyzed && A.beta_fx && A.exposure_fx && A.sus && A.risk_hist && A.R_cur && A.tx_hits,
            "tick_pass stage outputs");
    REQUIRE(A.E_cur && A.I_cur && A.tx_hits_by_strain && A.new_exposed_prev && A.new_exposed_by_strain_prev, "tick_pass census");
    if (A.flags & LPK_F_PENDING) REQUIRE(A.q_prev && A.cdf_prev, "tick_pass pending exposure inputs");
    const bool deaths = (A.flags & LPK_F_DEATHS) != 0, ri = (A.flags & LPK_F_RI) != 0, sia = (A.flags & LPK_F_SIA) != 0;
    if (sia) REQUIRE(P.date_of_birth && ALIGNED(P.date_of_birth, 16) && P.chronically_missed && ALIGNED(P.chronically_missed, 16) &&
                         A.sia_targeted && A.vx_prob_sia && A.sia_vaccinated && A.sia_protected && A.sia_new_exposed_by_strain &&
                         A.new_exposed && A.new_exposed_by_strain && A.sia_max_age >= A.sia_min_age && A.sia_strain >= 0 &&
                         A.sia_strain < A.n_strains, "tick_pass SIA");
    if (deaths) REQUIRE(P.date_of_death && ALIGNED(P.date_of_death, 16) && A.deaths && A.dead_pp && A.dead_par, "tick_pass deaths");
    if (ri) REQUIRE(P.ri_timer && ALIGNED(P.ri_timer, 8) && P.chronically_missed && ALIGNED(P.chronically_missed, 4) && A.vx_prob_ri &&
                        A.vx_prob_ipv && A.ri_vaccinated && A.ri_protected && A.ipv_vaccinated && A.new_exposed &&
                        A.new_exposed_by_strain && A.ri_new_exposed_by_strain && A.ri_step > 0 && A.ri_strain >= 0 &&
                        A.ri_strain < A.n_strains, "tick_pass RI");
    PassParams pp;
    pp.P = P;
    pp.A = A;
    pp.unit_ctr = pass_unit_counter();
    pp.use_tmap = pass_tensor_map(P, &pp.tmap) ? 1u : 0u;
    {
        static int dbg = -1;
        if (dbg < 0) { const char *e = getenv("LPK_PASS_DEBUG"); dbg = e ? atoi(e) : 0; }
        pp.debug = (uint32_t)dbg;
    }
    REQUIRE(pp.unit_ctr, "tick_pass work counter allocation");
    REQUIRE(ALIGNED(P.disease_state, 16) && ALIGNED(P.exposure_timer, 16) && ALIGNED(P.infection_timer, 16) && ALIGNED(P.strain, 16) &&
                ALIGNED(P.paralysis_timer, 16) && ALIGNED(P.potentially_paralyzed, 16) &&
                (!ri || (ALIGNED(P.chronically_missed, 16) && ALIGNED(P.ri_timer, 16))),
            "tick_pass alignment (bulk copies need 16-byte aligned columns)");
    cudaStream_t st = as_stream(stream);
    int rc;
    if (sia) {
        if (deaths && ri) rc = launch_pass<true, true, true, 8, 2>(pp, st);
        else if (deaths) rc = launch_pass<true, false, true, 8, 2>(pp, st);
        else if (ri) rc = launch_pass<false, true, true, 8, 2>(pp, st);
        else rc = launch_pass<false, false, true, 8, 2>(pp, st);
    } else if (deaths && ri) rc = launch_pass<true, true, false, 8, 2>(pp, st);
    else if (deaths) rc = launch_pass<true, false, false, 8, 2>(pp, st);
    else if (ri) rc = launch_pass<false, true, false, 8, 2>(pp, st);
    else rc = launch_pass<false, false, false, 8, 2>(pp, st);
    if (rc != LPK_OK) return rc;
    CUDA_TRY(cudaGetLastError(), "lpk_tick_pass");
    return LPK_OK;
}

// ------------------------------------------------------------------ tile -> node table
__global__ void k_build_tile_nodes(const int16_t *__restrict__ node_id, int64_t first_tile, int64_t n_tiles, int64_t n_slots,
                                   int32_t *__restrict__ tile_node) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = first_tile + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const int64_t lo = tile * LPK_TILE;
    const int first = node_id[lo];
    bool same = true;
    for (int k = lane; k < LPK_TILE; k += 32) {
        const int64_t i = lo + k;
        same &= (i < n_slots) && (node_id[i] == first);
    }
    same = __all_sync(LPK_FULL, same);
    if (lane == 0) tile_node[tile] = (same && first >= 0) ? first : -1;
}
extern "C" int lpk_build_tile_nodes(const int16_t *node_id, int64_t first_tile, int64_t n_slots, int32_t *tile_node, void *stream) {
    REQUIRE(node_id && tile_node && first_tile >= 0 && n_slots >= 0, "build_tile_nodes");
    const int64_t n_tiles = (n_slots + LPK_TILE - 1) / LPK_TILE;
    if (first_tile >= n_tiles) return LPK_OK;
    const int64_t todo = n_tiles - first_tile;
    k_build_tile_nodes<<<(unsigned)((todo + 7) / 8), 256, 0, as_stream(stream)>>>(node_id, first_tile, n_tiles, n_slots, tile_node);
    CUDA_TRY(cudaGetLastError(), "lpk_build_tile_nodes");
    return LPK_OK;
}

// ------------------------------------------------------------------ node-level epilogue of tick t
// one warp per node: the row sum of the network (coalesced) by all lanes, the node's bookkeeping by lane 0
__global__ void __launch_bounds__(256) k_tick_epilogue(const __grid_constant__ lpk_node_args a) {
    const int n_lo = a.node_hi > 0 ? a.node_lo : 0, n_hi = a.node_hi > 0 ? a.node_hi : a.n_nodes;
    const int n = n_lo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= n_hi) return;
    if (!(a.flags & LPK_F_ROWSUMS)) {  // once per network: at 6192 nodes re-reading the rows every tick cost 60 us of a 1 ms day
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;  // four independent chains: the loads of a row overlap
        const double *row = a.network + (int64_t)n * a.n_nodes;
        int j = lane;
        for (; j + 96 < a.n_nodes; j += 128) { s0 += row[j]; s1 += row[j + 32]; s2 += row[j + 64]; s3 += row[j + 96]; }
        for (; j < a.n_nodes; j += 32) s0 += row[j];
        double sum = (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(LPK_FULL, sum, o);
        if (lane == 0) a.rowsum_ws[n] = sum;
    }
    if (lane != 0) return;
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

int lpk_launch_node_math(int32_t num_nodes, int32_t n_strains, const int64_t *beta_fx, const int64_t *exposure_fx,
                         const int32_t *risk_hist, const double *network, double beta_seasonality, const double *r0_scalars,
                         const int32_t *alive_counts, double zero_inflation, double dispersion, float *tau, double *strain_cdf,
                         double *prob, double *expected, double *ws, uint64_t seed, uint32_t tick, cudaStream_t st, bool rowsums_done,
                         int32_t node_lo, int32_t node_hi);

extern "C" int lpk_tick_node(const lpk_node_args *args, void *stream) {
    REQUIRE(args, "tick_node null struct");
    const lpk_node_args &a = *args;
    REQUIRE(a.n_nodes > 0 && a.n_strains >= 1 && a.n_strains <= LPK_MAX_STRAINS, "tick_node sizes");
    REQUIRE(a.beta_fx && a.exposure_fx && a.risk_hist && a.network && a.r0_scalars && a.q && a.strain_cdf && a.prob && a.expected &&
                a.rowsum_ws, "tick_node node-math pointers");
    REQUIRE(!a.pop || (a.pop_prev && (!(a.flags & LPK_F_DEATHS) || (a.deaths_row && a.deaths))), "tick_node population rows");
    REQUIRE(!a.cur_potp || (a.cur_p && a.new_potential && a.new_paralyzed && a.potp_row && a.p_row), "tick_node paralysis rows");
    REQUIRE(!a.deaths || (a.dead_pp && a.dead_par), "tick_node death scratch");
    REQUIRE(!a.S_snap || (a.R_snap && a.sus && a.R_cur && a.tx_hits && (!(a.flags & LPK_F_PENDING) || (a.S_prev && a.R_prev))),
            "tick_node carried census");
    REQUIRE(a.pop || a.pop_prev, "tick_node needs a population row for the rate denominator");
    REQUIRE(a.E_cur && a.I_cur && a.E_snap && a.I_snap && a.tx_hits_by_strain &&
                (!(a.flags & LPK_F_PENDING) || (a.E_by_strain_prev && a.I_by_strain_prev && a.E_prev && a.I_prev)), "tick_node E / I census");
    cudaStream_t st = as_stream(stream);
    REQUIRE(a.node_hi == 0 || (a.node_lo >= 0 && a.node_lo < a.node_hi && a.node_hi <= a.n_nodes), "tick_node node shard");
    const int owned = a.node_hi > 0 ? a.node_hi - a.node_lo : a.n_nodes;
    k_tick_epilogue<<<(owned + 7) / 8, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError(), "lpk_tick_node epilogue");
    return lpk_launch_node_math(a.n_nodes, a.n_strains, a.beta_fx, a.exposure_fx, a.risk_hist, a.network, a.beta_seasonality,
                                a.r0_scalars, a.pop ? a.pop : a.pop_prev, a.zero_inflation, a.dispersion, a.q, a.strain_cdf, a.prob,
                                a.expected, a.rowsum_ws, a.seed, (uint32_t)a.tick, st, true, a.node_hi > 0 ? a.node_lo : 0,
                                a.node_hi > 0 ? a.node_hi : a.n_nodes);
}
