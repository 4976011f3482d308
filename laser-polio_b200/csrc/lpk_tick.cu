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
#include "lpk_host.cuh"
#include "lpk_stages.cuh"

struct PassParams {
    lpk_people P;
    lpk_tick_args A;
};

#define QCAP 512             // ring entries per warp: 31 left over + the 256 agents of one iteration fit
#define LPK_CHUNK_ROWS 256   // rows of 128 agents per chunk (32 K agents), dealt round-robin to the blocks
#define LPK_CHUNK_AGENTS (LPK_CHUNK_ROWS * 128)
#ifndef LPK_PASS_BLOCKS_PER_SM
#define LPK_PASS_BLOCKS_PER_SM 2
#endif

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

// bookkeeping of an exposure hit of tick t-1: categorical strain pick (model.py:1127-1141), rows t-1; returns the strain
__device__ __forceinline__ int8_t expose_bookkeeping(const PassParams &pp, int64_t i, int nd) {
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
    leave_S(pp, i, nd);
    return (int8_t)assigned;
}
__device__ __noinline__ void expose_agent(const PassParams &pp, int64_t i, int nd) { expose_bookkeeping(pp, i, nd); }

// census of one E or I agent (rows t-1)
__device__ __noinline__ void census_ei(const PassParams &pp, int64_t i, int nd, int8_t s) {
    const int64_t c = (int64_t)nd * pp.A.n_strains + pp.P.strain[i];
    atomicAdd(s == 1 ? &pp.A.E_by_strain_prev[c] : &pp.A.I_by_strain_prev[c], 1);
}

__device__ __noinline__ void kill_agent(const PassParams &pp, int64_t i, int nd, int8_t state_before) {
    if (state_before == 0) leave_S(pp, i, nd);
    if (state_before == 3) atomicAdd(&pp.A.R_cur[nd], -1);
    atomicAdd(&pp.A.deaths[nd], 1);
    if (pp.P.potentially_paralyzed[i] == 1) atomicAdd(&pp.A.dead_pp[nd], 1);
    if (pp.P.paralyzed[i] == 1) atomicAdd(&pp.A.dead_par[nd], 1);
}

__device__ __noinline__ int8_t ds_agent_ol(const PassParams &pp, int64_t i, int8_t s, int nd) {
    const lpk_people &P = pp.P;
    const int8_t s2 = ds_agent(i, s, P.node_id, P.strain, P.exposure_timer, P.infection_timer, P.potentially_paralyzed, P.paralyzed,
                               P.ipv_protected, P.paralysis_timer, (double)pp.A.p_paralysis, pp.A.new_potential, pp.A.new_paralyzed,
                               stage_rng(pp));
    if (s2 == 3) atomicAdd(&pp.A.R_cur[nd], 1);  // s was E or I: a recovery
    return s2;
}

// infectivity tally of one infectious agent (tick t)
__device__ __noinline__ void tally_infectious(const PassParams &pp, int64_t i, int nd) {
    const int s = pp.P.strain[i];
    red_add(&pp.A.beta_fx[(int64_t)nd * pp.A.n_strains + s], to_fx((double)pp.P.daily_infectivity[i] * pp.A.strain_r0_scalars[s]));
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
            if (s == 1 || s == 2) census_ei(pp, i, nd, s);
        }
        if (deaths && P.date_of_death[i] <= A.tick) { kill_agent(pp, i, nd, s); s = -1; }
        if (s == 1 || s == 2) s = ds_agent_ol(pp, i, s, nd);
        nw = set_byte(nw, k, s);
    }
    if (ri) nw = ri_quad(pp, b, valid, nw);
#pragma unroll 1
    for (int k = 0; k < valid; ++k)
        if (byte_of(nw, k) == 2) tally_infectious(pp, b + k, P.node_id[b + k]);
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

// ---- active-agent queue -------------------------------------------------------------------------------------
// E / I agents (and fresh exposure hits) are appended to the warp's shared-memory ring and, whenever 32 have
// accumulated, processed one per lane with all lanes busy.  An entry carries everything the handler needs; the handler
// owns the agent's state byte from then on (the owning lane already stored the quad's word; both stores come from the
// same warp, ordered by the warp-wide reduction in q_commit).
// entry = {agent index (tables hold < 2^32 slots), node | state << 16 | hit << 20}
struct WarpQueue {
    uint2 *q;
    uint32_t *tail;  // shared, monotonic
    uint32_t head;   // warp-uniform
    int count;       // warp-uniform
};

// census (rows t-1) -> disease state (tick t) -> infectivity tally (tick t) for one active agent.
// Every column the agent can need is loaded up front, in one round trip, and the state machine runs on registers
// (profiles/r1_fused_v6_postsia_*: a third of the stall samples sat on the serial etimer -> itimer -> strain -> ptimer
// -> infectivity chain of dependent scattered loads).
__device__ __noinline__ void active_agent(const PassParams &pp, uint2 e) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const int64_t i = (int64_t)e.x;
    const int nd = (int)(int16_t)(e.y & 0xFFFFu);
    const int8_t s0 = (int8_t)((e.y >> 16) & 0xFu);
    const bool hit = (e.y >> 20) & 1u;
    const int ns = A.n_strains;
    int8_t st = hit ? (int8_t)0 : P.strain[i];
    int8_t et = (s0 == 1) ? P.exposure_timer[i] : (int8_t)1;
    int8_t it = P.infection_timer[i], pt = P.paralysis_timer[i], pq = P.potentially_paralyzed[i];
    const int8_t ipvv = P.ipv_protected[i];
    const float inf = P.daily_infectivity[i];
    if (hit) st = expose_bookkeeping(pp, i, nd);  // exposure hit of tick t-1
    if (A.flags & LPK_F_PENDING) {
        const int64_t c = (int64_t)nd * ns + st;
        atomicAdd(s0 == 1 ? &A.E_by_strain_prev[c] : &A.I_by_strain_prev[c], 1);
    }
    int8_t s = s0;
    if (s == 1) {  // model.py:419-422
        if (et <= 0) s = 2;
        P.exposure_timer[i] = (int8_t)(et - 1);
    }
    if (s == 2) {
        const int8_t pq0 = pq;
        int8_t par = 0;
        int flags;
        s = ds_infected(i, st, ipvv, it, pt, pq, par, (double)A.p_paralysis, stage_rng(pp), flags);
        P.infection_timer[i] = it;
        if (st == 0) {
            P.paralysis_timer[i] = pt;
            if (pq != pq0) P.potentially_paralyzed[i] = pq;
            if (flags) {
                atomicAdd(&A.new_potential[nd], 1);
                if (flags & 2) { P.paralyzed[i] = 1; atomicAdd(&A.new_paralyzed[nd], 1); }
            }
        }
    }
    if (s != s0) P.disease_state[i] = s;
    if (s == 2) red_add(&A.beta_fx[(int64_t)nd * ns + st], to_fx((double)inf * A.strain_r0_scalars[st]));
    if (s == 3) atomicAdd(&A.R_cur[nd], 1);
}

// append the agents of mask m (bit 0 of byte k = agent idx0 + k) of state word nw; returns how many
__device__ __forceinline__ int q_push(WarpQueue &Q, uint32_t idx0, int nd, uint32_t nw, uint32_t hits, uint32_t m) {
    const uint32_t comb = nw | (hits << 4);
    const int cnt = __popc(m);
    while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1u;
        const uint32_t pos = atomicAdd(Q.tail, 1u) & (QCAP - 1);
        Q.q[pos] = make_uint2(idx0 + (uint32_t)(bit >> 3), ((uint32_t)nd & 0xFFFFu) | (((comb >> bit) & 0xFFu) << 16));
    }
    return cnt;
}
// every lane pushed `mine` entries: drain the ring 32 at a time (warp-uniform control flow)
__device__ __forceinline__ void q_commit(const PassParams &pp, WarpQueue &Q, int mine, int lane) {
    Q.count += __reduce_add_sync(LPK_FULL, mine);
    while (Q.count >= 32) {
        __syncwarp();
        active_agent(pp, Q.q[(Q.head + lane) & (QCAP - 1)]);
        Q.head += 32;
        Q.count -= 32;
    }
}

// out-of-line part of a death in a fast chunk: the quad's agents in mask dm die on tick t (after tick t-1's pending
// exposure + census); returns {new state word, remaining hits}
__device__ __noinline__ uint2 death_quad(const PassParams &pp, int64_t b, int nd, uint32_t nw, uint32_t hits, uint32_t dm) {
    const bool pending = (pp.A.flags & LPK_F_PENDING) != 0;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        if (!((dm >> (8 * k)) & 1u)) continue;
        const int8_t s = byte_of(nw, k);
        if ((hits >> (8 * k)) & 1u) { expose_agent(pp, b + k, nd); hits &= ~(1u << (8 * k)); }
        if (pending && (s == 1 || s == 2)) census_ei(pp, b + k, nd, s);
        kill_agent(pp, b + k, nd, s);
        nw = set_byte(nw, k, -1);
    }
    return make_uint2(nw, hits);
}
__device__ __forceinline__ uint32_t death_mask(const int4 &d, int tick, uint32_t w) {
    return ((d.x <= tick ? 1u : 0u) | (d.y <= tick ? 0x100u : 0u) | (d.z <= tick ? 0x10000u : 0u) | (d.w <= tick ? 0x1000000u : 0u)) &
           mask_alive(w);
}

// ------------------------------------------------------------------ fast chunk: 32 K agents of ONE node, all of them
// present at tick t-1.  A warp takes every 8th row pair; a lane owns its quad in the even row (A) and in the odd row (B).
struct PairData {
    uint32_t wA, wB;
    float4 rA, rB;
    int4 dA, dB;
};
template <bool kDeaths>
__device__ __forceinline__ void fast_chunk(const PassParams &pp, WarpQueue &Q, int64_t base, int nd, int lane, int warp) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const bool pending = (A.flags & LPK_F_PENDING) != 0;
    const float tau = pending ? __ldg(&A.q_prev[nd]) : 0.f;
    const bool expose = tau > 0.f;  // no force of infection on the node: neither risk nor random numbers are needed
    const float tau16 = tau * 65536.0f;
    const int tick = A.tick;
    const uint32_t k0 = (uint32_t)A.seed, k1 = (uint32_t)(A.seed >> 32);
    const int64_t lane_base = base + lane * 4;
    const uint64_t ctr_base = ((((uint64_t)base + A.id_base) >> 8) << 5) + (uint64_t)lane;
    const int8_t *sp = P.disease_state + lane_base;
    const float *rp = P.acq_risk_multiplier + lane_base;
    const int32_t *dp = kDeaths ? P.date_of_death + lane_base : nullptr;
    auto load = [&](int p, PairData &d) {
        const int o = p * 256;
        d.wA = *reinterpret_cast<const uint32_t *>(sp + o);
        d.wB = *reinterpret_cast<const uint32_t *>(sp + o + 128);
        if (expose) {
            d.rA = __ldg(reinterpret_cast<const float4 *>(rp + o));
            d.rB = __ldg(reinterpret_cast<const float4 *>(rp + o + 128));
        }
        if (kDeaths) {
            d.dA = __ldg(reinterpret_cast<const int4 *>(dp + o));
            d.dB = __ldg(reinterpret_cast<const int4 *>(dp + o + 128));
        }
    };
    PairData cur, nxt;
    load(warp, cur);
#pragma unroll 2
    for (int p = warp; p < LPK_CHUNK_ROWS / 2; p += LPK_WARPS) {
        if (p + LPK_WARPS < LPK_CHUNK_ROWS / 2) load(p + LPK_WARPS, nxt);
        const int64_t bA = lane_base + p * 256, bB = bA + 128;
        uint32_t nwA = cur.wA, nwB = cur.wB, hA = 0u, hB = 0u;
        if (expose) {  // exposure trial of tick t-1
            const uint64_t c = ctr_base + (uint64_t)(p * 32);
            uint32_t x[4];
            philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(tick - 1), LPK_STAGE_EXPOSE, k0, k1, x);
            if (pretest_quad(x[0], x[1], cur.rA, tau16) | pretest_quad(x[2], x[3], cur.rB, tau16)) {
                hA = exact_quad(pp, (uint32_t)c, (uint32_t)(c >> 32), 0, x[0], x[1], nwA, cur.rA, tau);
                hB = exact_quad(pp, (uint32_t)c, (uint32_t)(c >> 32), 1, x[2], x[3], nwB, cur.rB, tau);
                nwA |= hA;  // S (0) -> E (1)
                nwB |= hB;
            }
        }
        if (kDeaths) {  // tick t
            const uint32_t dmA = death_mask(cur.dA, tick, nwA), dmB = death_mask(cur.dB, tick, nwB);
            if (dmA) { const uint2 r = death_quad(pp, bA, nd, nwA, hA, dmA); nwA = r.x; hA = r.y; }
            if (dmB) { const uint2 r = death_quad(pp, bB, nd, nwB, hB, dmB); nwB = r.x; hB = r.y; }
        }
        if (nwA != cur.wA) *reinterpret_cast<uint32_t *>(P.disease_state + bA) = nwA;
        if (nwB != cur.wB) *reinterpret_cast<uint32_t *>(P.disease_state + bB) = nwB;
        // exposed / infectious agents (fresh hits included): census of t-1, disease state and tally of t in the handler
        int mine = q_push(Q, (uint32_t)bA, nd, nwA, hA, mask_EI(nwA));
        mine += q_push(Q, (uint32_t)bB, nd, nwB, hB, mask_EI(nwB));
        q_commit(pp, Q, mine, lane);
        cur = nxt;
    }
}

// ------------------------------------------------------------------ general rows: node boundaries, newborn cohorts, the
// table's tail, RI ticks.  A warp walks rows of 128 agents (one quad per lane).  Software pipeline: the state word of row
// r+2, and the risk / node / date_of_death words of row r+1 (predicated on its state, which arrived an iteration ago),
// are in flight while row r is processed.
struct RowData {
    uint32_t w;      // 4 state bytes
    float4 rk;       // acq_risk_multiplier of the quad (if it has a susceptible)
    uint2 nd;        // 4 node ids (only when the tile is not node-uniform)
    int4 dd;         // date_of_death (vital-dynamics ticks only)
    int tn;          // tile's node or -1
};
struct TauCache {
    int node;
    float tau;
};

template <bool kDeaths>
__device__ __forceinline__ void issue_row_loads(const lpk_people &P, const float *tau_prev, int64_t row, int lane, int64_t n,
                                                uint32_t w, int tn, TauCache &tc, RowData &d) {
    const int64_t b = (row * 32 + lane) * 4;
    d.w = w;
    d.tn = tn;
    const bool full = b + 4 <= n;
    const bool alive = (w & 0x80808080u) != 0x80808080u;
    d.rk = make_float4(0.f, 0.f, 0.f, 0.f);
    d.nd = make_uint2(0u, 0u);
    if (full && alive) {
        // risk is only needed for the exposure trial: skipped when nothing is pending or the tile's node has no force of infection
        bool live = tau_prev != nullptr;
        if (live && d.tn >= 0) {
            if (d.tn != tc.node) { tc.node = d.tn; tc.tau = __ldg(&tau_prev[d.tn]); }  // warp-uniform branch
            live = tc.tau > 0.f;
        }
        if (live && mask_S(w)) d.rk = __ldg(reinterpret_cast<const float4 *>(P.acq_risk_multiplier + b));
        if (d.tn < 0) d.nd = *reinterpret_cast<const uint2 *>(P.node_id + b);
        if (kDeaths) d.dd = __ldg(reinterpret_cast<const int4 *>(P.date_of_death + b));
    }
}
__device__ __forceinline__ int load_tile_node(const lpk_people &P, int64_t row) {
    return P.tile_node ? __ldg(&P.tile_node[row >> 2]) : -1;
}
__device__ __forceinline__ uint32_t load_state_row(const lpk_people &P, int64_t row, int lane, int64_t n) {
    const int64_t b = (row * 32 + lane) * 4;
    const int v = quad_valid(b, n);
    return v ? load_b4(P.disease_state, b, v) : 0xFFFFFFFFu;
}

template <bool kDeaths, bool kRI>
__device__ __forceinline__ void general_rows(const PassParams &pp, WarpQueue &Q, int64_t lo, int64_t hi, int64_t n, int64_t count_prev,
                                             TauCache &tc_load, TauCache &tc_use, int lane, int warp) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const bool pending = (A.flags & LPK_F_PENDING) != 0;
    const float *tau_prev = pending ? A.q_prev : nullptr;
    const int tick = A.tick;
    int64_t row = lo + warp;
    RowData cur, nxt;
    uint32_t w2 = 0xFFFFFFFFu;  // state word and tile node two rows ahead
    int tn2 = -1;
    cur.w = 0xFFFFFFFFu;
    if (row < hi) {
        issue_row_loads<kDeaths>(P, tau_prev, row, lane, n, load_state_row(P, row, lane, n), load_tile_node(P, row), tc_load, cur);
        if (row + LPK_WARPS < hi) { w2 = load_state_row(P, row + LPK_WARPS, lane, n); tn2 = load_tile_node(P, row + LPK_WARPS); }
    }
#pragma unroll 1
    for (; row < hi; row += LPK_WARPS) {
        // ---- keep the pipeline full
        const int64_t r1 = row + LPK_WARPS, r2 = row + 2 * LPK_WARPS;
        nxt.w = 0xFFFFFFFFu;
        if (r1 < hi) issue_row_loads<kDeaths>(P, tau_prev, r1, lane, n, w2, tn2, tc_load, nxt);
        w2 = (r2 < hi) ? load_state_row(P, r2, lane, n) : 0xFFFFFFFFu;
        tn2 = (r2 < hi) ? load_tile_node(P, r2) : -1;

        // ---- row `row`
        const uint32_t w = cur.w;
        const int64_t b = (row * 32 + lane) * 4;
        uint32_t cand = 0u, hits = 0u, nw = w;  // cand: agents that go to the active queue
        int nd = cur.tn;
        if ((w & 0x80808080u) != 0x80808080u) {  // somebody alive in the quad
            const int valid = quad_valid(b, n);
            bool fast = (valid == 4) && (!pending || b + 4 <= count_prev);
            if (nd < 0) {
                nd = (int)(int16_t)(cur.nd.x & 0xFFFFu);
                fast = fast && cur.nd.x == cur.nd.y && (cur.nd.x >> 16) == (cur.nd.x & 0xFFFFu) && nd >= 0;
            }
            if (!fast) {
                nw = slow_quad(pp, b, valid, w, kDeaths, kRI, count_prev);
            } else {
                if (pending && mask_S(w)) {  // exposure trial of tick t-1
                    if (nd != tc_use.node) { tc_use.node = nd; tc_use.tau = __ldg(&A.q_prev[nd]); }
                    const float tau = tc_use.tau;
                    if (tau > 0.f) {
                        const uint64_t id0 = (uint64_t)b + A.id_base;
                        const uint64_t c = expose_ctr(id0);
                        const int par = (int)((id0 >> 7) & 1u);
                        uint32_t x[4];
                        philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(tick - 1), LPK_STAGE_EXPOSE, (uint32_t)A.seed,
                                      (uint32_t)(A.seed >> 32), x);
                        const uint32_t xa = par ? x[2] : x[0], xb = par ? x[3] : x[1];
                        if (pretest_quad(xa, xb, cur.rk, tau * 65536.0f)) {
                            hits = exact_quad(pp, (uint32_t)c, (uint32_t)(c >> 32), par, xa, xb, w, cur.rk, tau);
                            nw |= hits;  // S (0) -> E (1)
                        }
                    }
                }
                // ---- tick t
                if (kDeaths) {
                    const uint32_t dm = death_mask(cur.dd, tick, nw);
                    if (dm) { const uint2 r = death_quad(pp, b, nd, nw, hits, dm); nw = r.x; hits = r.y; }
                }
                cand = mask_EI(nw);
                if (kRI) {
                    // RI may set ipv_protected, which the disease-state step of the SAME tick must not see yet
                    // (reference order: DiseaseState_ABM before RI_ABM), so nothing is deferred on RI ticks.
#pragma unroll 1
                    for (int k = 0; k < 4; ++k) {
                        if (!((cand >> (8 * k)) & 1u)) continue;
                        if ((hits >> (8 * k)) & 1u) expose_agent(pp, b + k, nd);
                        int8_t sk = byte_of(nw, k);
                        if (pending) census_ei(pp, b + k, nd, sk);
                        sk = ds_agent_ol(pp, b + k, sk, nd);
                        nw = set_byte(nw, k, sk);
                        if (sk == 2) tally_infectious(pp, b + k, nd);
                    }
                    cand = 0u;
                    nw = ri_quad(pp, b, 4, nw);
                }
            }
            if (nw != w) store_b4(P.disease_state, b, valid, nw);
        }
        q_commit(pp, Q, q_push(Q, (uint32_t)b, nd, nw, hits, cand), lane);
        cur = nxt;
    }
}

template <bool kDeaths, bool kRI>
__global__ void __launch_bounds__(LPK_BLOCK, LPK_PASS_BLOCKS_PER_SM) k_tick_pass(const __grid_constant__ PassParams pp) {
    __shared__ uint2 queue[LPK_WARPS][QCAP];
    __shared__ uint32_t q_tail[LPK_WARPS];
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t count_prev = A.counts[0], n = A.counts[1];
    const int64_t rows = (n + 127) >> 7;
    const int64_t n_chunks = (rows + LPK_CHUNK_ROWS - 1) / LPK_CHUNK_ROWS;
    WarpQueue Q;
    Q.q = queue[warp];
    Q.tail = &q_tail[warp];
    Q.head = 0u;
    Q.count = 0;
    if (lane == 0) q_tail[warp] = 0u;
    __syncwarp();
    TauCache tc_load = {-2, 0.f}, tc_use = {-2, 0.f};

    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int64_t base = chunk * LPK_CHUNK_AGENTS;
        // one node, every slot occupied since before tick t-1's transmission?  (64 tiles of 512 agents)
        int cn = -1;
        if (!kRI && P.tile_node && base + LPK_CHUNK_AGENTS <= count_prev) {
            const int t0 = __ldg(&P.tile_node[chunk * 64 + lane]), t1 = __ldg(&P.tile_node[chunk * 64 + 32 + lane]);
            const int first = __shfl_sync(LPK_FULL, t0, 0);
            if (__all_sync(LPK_FULL, t0 == first && t1 == first)) cn = first;
        }
        if (cn >= 0) {
            fast_chunk<kDeaths>(pp, Q, base, cn, lane, warp);
        } else {
            const int64_t lo = chunk * LPK_CHUNK_ROWS;
            const int64_t hi = (lo + LPK_CHUNK_ROWS < rows) ? lo + LPK_CHUNK_ROWS : rows;
            general_rows<kDeaths, kRI>(pp, Q, lo, hi, n, count_prev, tc_load, tc_use, lane, warp);
        }
    }
    __syncwarp();
    if (lane < Q.count) active_agent(pp, Q.q[(Q.head + lane) & (QCAP - 1)]);
}

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
    REQUIRE(A.E_by_strain_prev && A.I_by_strain_prev && A.new_exposed_prev && A.new_exposed_by_strain_prev, "tick_pass census rows");
    if (A.flags & LPK_F_PENDING) REQUIRE(A.q_prev && A.cdf_prev, "tick_pass pending exposure inputs");
    const bool deaths = (A.flags & LPK_F_DEATHS) != 0, ri = (A.flags & LPK_F_RI) != 0;
    if (deaths) REQUIRE(P.date_of_death && ALIGNED(P.date_of_death, 16) && A.deaths && A.dead_pp && A.dead_par, "tick_pass deaths");
    if (ri) REQUIRE(P.ri_timer && ALIGNED(P.ri_timer, 8) && P.chronically_missed && ALIGNED(P.chronically_missed, 4) && A.vx_prob_ri &&
                        A.vx_prob_ipv && A.ri_vaccinated && A.ri_protected && A.ipv_vaccinated && A.new_exposed &&
                        A.new_exposed_by_strain && A.ri_new_exposed_by_strain && A.ri_step > 0, "tick_pass RI");
    PassParams pp;
    pp.P = P;
    pp.A = A;
    const int grid = lpk_agent_grid(P.capacity, LPK_PASS_BLOCKS_PER_SM);
    cudaStream_t st = as_stream(stream);
    if (deaths && ri) k_tick_pass<true, true><<<grid, LPK_BLOCK, 0, st>>>(pp);
    else if (deaths) k_tick_pass<true, false><<<grid, LPK_BLOCK, 0, st>>>(pp);
    else if (ri) k_tick_pass<false, true><<<grid, LPK_BLOCK, 0, st>>>(pp);
    else k_tick_pass<false, false><<<grid, LPK_BLOCK, 0, st>>>(pp);
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
__global__ void k_tick_epilogue(lpk_node_args a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0 && a.counts) a.counts[0] = a.counts[1];
    if (n >= a.n_nodes) return;
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
    if ((a.flags & LPK_F_PENDING) && a.E_prev) {
        int e = 0, i = 0;
        for (int s = 0; s < ns; ++s) { e += a.E_by_strain_prev[(int64_t)n * ns + s]; i += a.I_by_strain_prev[(int64_t)n * ns + s]; }
        a.E_prev[n] = e; a.I_prev[n] = i;
    }
    if (a.next_beta_fx)
        for (int s = 0; s < ns; ++s) a.next_beta_fx[(int64_t)n * ns + s] = 0;
}

int lpk_launch_node_math(int32_t num_nodes, int32_t n_strains, const int64_t *beta_fx, const int64_t *exposure_fx,
                         const int32_t *risk_hist, const double *network, double beta_seasonality, const double *r0_scalars,
                         const int32_t *alive_counts, double zero_inflation, double dispersion, float *tau, double *strain_cdf,
                         double *prob, double *expected, double *ws, uint64_t seed, uint32_t tick, cudaStream_t st);

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
    cudaStream_t st = as_stream(stream);
    k_tick_epilogue<<<(a.n_nodes + 127) / 128, 128, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError(), "lpk_tick_node epilogue");
    return lpk_launch_node_math(a.n_nodes, a.n_strains, a.beta_fx, a.exposure_fx, a.risk_hist, a.network, a.beta_seasonality,
                                a.r0_scalars, a.pop ? a.pop : a.pop_prev, a.zero_inflation, a.dispersion, a.q, a.strain_cdf, a.prob,
                                a.expected, a.rowsum_ws, a.seed, (uint32_t)a.tick, st);
}
