// lpk_tick.cu -- the fused tick: ONE sweep over the agent table per simulated day, reading one byte per agent.
//
// Pass for tick t, per agent, in the reference's order (include/lpk.h, "Fused tick"):
//   pending tick t-1:  exposure trial (tx_infect)  ->  census (count_SEIRP)
//   tick t:            deaths (get_deaths) -> disease state (disease_state_step) -> RI (fast_ri) -> SIA (fast_sia) -> tally
// Every draw is Philox(seed; agent, tick, stage), so the fused pass reproduces the per-function kernels bit for bit
// (tests/test_gpu_fused.py).
//
// Round 1's pass streamed 12 B per agent (state, five timer / flag byte columns, risk) and spent a third of its
// instructions counting three timers down on byte lanes (profiles/r1_fused_v25_*: 389 M warp-instructions, 630 us at
// 2.2e8 agents, issue bound).  This one keeps an AGENDA BYTE per agent (lpk_hot.cuh): class + either a 6-bit upper bound of
// the agent's risk (susceptibles) or the day of its next event (exposed / infectious; timers are deadlines, nothing counts
// down).  The sweep reads that byte only:
//   * susceptibles: one Philox block per 8 agents, U = 2^23 + h16 against fma(decoded bound, tau', 2^23 + 1) -- the
//     16-bit pre-test of round 1 with the risk replaced by its bound, so the candidates are a superset of the hits;
//   * exposed / infectious: one byte-equality test of the payload against today (SWAR, 7 instructions per 8 agents);
//   * recovered / dead / unborn slots: nothing.
// Candidates and agents whose day has come (~1 % of the table per day) go to the warp's shared-memory ring and are
// handled 32 at a time by hot_event on the full-width columns (exact trial, strain pick, state change, paralysis gate,
// vaccine draws), with the node-level counts collected in per-warp accumulators.  Tallies (susceptibles, risk sums,
// risk histogram, E / I / R census, infectivity) are CARRIED from tick to tick and only corrected by events.
// Vital-dynamics days read date_of_death only for pairs whose earliest death date has come (pair_min_dod); RI days add
// chronically_missed + ri_timer, campaign days chronically_missed + date_of_birth in targeted nodes.
#include <climits>
#include <cstdlib>

#include "lpk_host.cuh"
#include "lpk_hot.cuh"
#include "lpk_node.cuh"

struct PassParams {
    lpk_people P;
    lpk_tick_args A;
    uint32_t *unit_ctr;  // work counter of this launch (zero when the kernel starts)
    uint32_t *unit_ctr_next;  // optional: the counter of the NEXT launch, zeroed by this one (lpk_tick_args.work_counter_next)
    uint32_t unit_log;   // log2 of the pairs per work unit of this launch (1 .. LPK_UNIT_LOG): LPK_PASS_UNIT_LOG experiments
    uint32_t debug;      // timing experiments only (LPK_PASS_DEBUG): 1 = drop ring batches
    uint32_t rk[20];     // Philox round keys of the seed (key schedule done once on the host: the sweep's block reads them
                         // straight from the constant bank instead of spending 20 additions per 8 agents)
};

// Philox4x32-10 with the key schedule taken from the kernel parameters
__device__ __forceinline__ void philox_sweep(const PassParams &pp, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ pp.rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ pp.rk[2 * r + 1];
        c3 = lo0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define QCAP 1024        // ring entries per warp: 31 left over + the 512 agents of one tile fit
#define LPK_UNIT_LOG 3   // a work unit = 8 consecutive pairs of 128-agent rows (2048 agents = 2 KB of agenda bytes)
#define LPK_UNIT_PAIRS (1 << LPK_UNIT_LOG)
#define LPK_UNIT_AGENTS (256 << LPK_UNIT_LOG)

// ------------------------------------------------------------------ per-warp accumulators of the node-level counts
// The agents of a warp come from one node for many ring batches and all SMs work in few nodes at a time, so per-agent
// global atomics land on a handful of addresses from every SM at once and serialise in L2 (round 1, diag_v11).  Counts
// are collected here (shared memory; lanes add with shared-memory atomics) and flushed by lane 0 with one global atomic
// per counter when the warp's node changes.
struct WarpAcc {
    int node;
    int E[LPK_MAX_STRAINS], I[LPK_MAX_STRAINS];  // changes of the carried exposed / infectious counts
    int H[LPK_MAX_STRAINS];                      // exposure hits of t-1 per strain (already included in E)
    int R;                                       // change of the carried recovered count
    int riV, riP, ipvV, siaV, siaP;              // RI / SIA of tick t: vaccinated, protected (S -> E)
    int Sdead;                                   // susceptibles that died
    long long beta[LPK_MAX_STRAINS];             // change of the carried infectivity tally (fixed point)
    long long expo;                              // risk (fixed point) of the agents that left S
};
__device__ __forceinline__ void acc_clear(WarpAcc *a) {
#pragma unroll
    for (int s = 0; s < LPK_MAX_STRAINS; ++s) { a->E[s] = 0; a->I[s] = 0; a->H[s] = 0; a->beta[s] = 0; }
    a->R = 0;
    a->riV = a->riP = a->ipvV = a->siaV = a->siaP = a->Sdead = 0;
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
    red_add(&A.sus[nd], -(long long)(hits + a->riP + a->siaP + a->Sdead));
    red_add(&A.exposure_fx[nd], -a->expo);
    red_add(&A.R_cur[nd], a->R);
    acc_clear(a);
}
__device__ __forceinline__ void acc_add(long long *p, long long v) { atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v); }

// ---- event ring ----------------------------------------------------------------------------------------------
// Agents with something to do are appended to the warp's shared-memory ring and, whenever 32 have accumulated, handled
// one per lane with all lanes busy (they sit in most 256-agent pairs, so handling them in place would make every warp
// walk the long event path with one or two live lanes).  entry = {agent index (tables hold < 2^32 slots),
// node | EV_* flags}.  The handler runs in two phases a ring batch apart: active_load issues the scattered loads,
// active_process runs hot_event on those registers when the NEXT batch is loaded, so the load latency is spent sweeping.
struct WarpQueue {
    uint2 *q;
    uint32_t *tail;  // shared, monotonic
    uint32_t head;   // warp-uniform
    int count;       // warp-uniform
    bool loaded;     // warp-uniform: `pend` holds a batch whose loads are in flight
    uint2 pend_e;
    HotPre pend;
    WarpAcc *acc;
};

// warp-collective: every lane calls it; `valid` lanes carry an agent
__device__ __noinline__ void active_process(const PassParams &pp, uint2 e, HotPre pre, bool valid, WarpAcc *acc, int lane) {
    const lpk_tick_args &A = pp.A;
    HotDelta d;
    d.nd = -1; d.st = 0; d.dE = d.dI = d.dR = 0; d.hit = d.vx = d.gate = d.died = 0; d.dbeta = 0; d.efx = 0; d.hbin = -1;
    if (valid) {
        d = hot_event(pp.P, A, (int64_t)e.x, (int)(int16_t)(e.y & 0xFFFFu), e.y, pre);
        // rare bookkeeping: direct atomics
        if (d.hbin >= 0) atomicAdd(&A.risk_hist[(int64_t)d.nd * LPK_RISK_BINS + d.hbin], -1);
        if (d.gate) {
            atomicAdd(&A.new_potential[d.nd], 1);
            if (d.gate & 2) atomicAdd(&A.new_paralyzed[d.nd], 1);
        }
        if (d.died) {
            atomicAdd(&A.deaths[d.nd], 1);
            if (d.died & 2) atomicAdd(&A.dead_pp[d.nd], 1);
            if (d.died & 4) atomicAdd(&A.dead_par[d.nd], 1);
        }
    }
    const bool any = valid && (d.hit | d.dE | d.dI | d.dR | d.vx | (d.died & 8) | (d.efx != 0) | (d.dbeta != 0));
    // node-level counts, one group of same-node lanes at a time (one group except at a node boundary)
    uint32_t todo = __ballot_sync(LPK_FULL, any);
    while (todo) {
        const int nd0 = __shfl_sync(LPK_FULL, d.nd, __ffs(todo) - 1);
        const bool mine = any && d.nd == nd0;
        if (lane == 0 && acc->node != nd0) { acc_flush(pp, acc); acc->node = nd0; }
        __syncwarp();
        if (mine) {
            if (d.hit) atomicAdd(&acc->H[d.st], 1);
            if (d.dE) atomicAdd(&acc->E[d.st], (int)d.dE);
            if (d.dI) atomicAdd(&acc->I[d.st], (int)d.dI);
            if (d.dR) atomicAdd(&acc->R, (int)d.dR);
            if (d.dbeta) acc_add(&acc->beta[d.st], d.dbeta);
            if (d.efx) acc_add(&acc->expo, d.efx);
            if (d.died & 8) atomicAdd(&acc->Sdead, 1);
            if (d.vx) {
                if (d.vx & 1u) atomicAdd(&acc->riV, 1);
                if (d.vx & 2u) atomicAdd(&acc->riP, 1);
                if (d.vx & 4u) atomicAdd(&acc->ipvV, 1);
                if (d.vx & 8u) atomicAdd(&acc->siaV, 1);
                if (d.vx & 16u) atomicAdd(&acc->siaP, 1);
            }
        }
        __syncwarp();
        todo &= ~__ballot_sync(LPK_FULL, mine);
    }
}

// every lane pushed `mine` entries: drain the ring 32 at a time (warp-uniform control flow)
__device__ __forceinline__ void q_commit(const PassParams &pp, WarpQueue &Q, int mine, int lane) {
    Q.count += __reduce_add_sync(LPK_FULL, mine);
    while (Q.count >= 32) {
        __syncwarp();
        if (pp.debug & 1u) { Q.head += 32; Q.count -= 32; continue; }
        // process the batch loaded a commit ago FIRST: a call boundary waits for every load in flight, so the new batch's
        // loads are issued after it and land while the warp sweeps the next pairs
        if (Q.loaded) active_process(pp, Q.pend_e, Q.pend, true, Q.acc, lane);
        Q.pend_e = Q.q[(Q.head + lane) & (QCAP - 1)];
        Q.pend = hot_preload(pp.P, (int64_t)Q.pend_e.x, Q.pend_e.y);
        Q.loaded = true;
        Q.head += 32;
        Q.count -= 32;
    }
}

__device__ __forceinline__ uint32_t death_mask(const int4 &d, int tick) {
    return (d.x <= tick ? 1u : 0u) | (d.y <= tick ? 0x100u : 0u) | (d.z <= tick ? 0x10000u : 0u) | (d.w <= tick ? 0x1000000u : 0u);
}
__device__ __forceinline__ int min_dod_left(const int4 &d, uint32_t keep) {  // earliest date of death among the agents in mask `keep`
    int m = INT_MAX;
    if (keep & 1u) m = min(m, d.x);
    if (keep & 0x100u) m = min(m, d.y);
    if (keep & 0x10000u) m = min(m, d.z);
    if (keep & 0x1000000u) m = min(m, d.w);
    return m;
}
__device__ __forceinline__ uint32_t sia_age_mask(const int4 &d, int tick, int lo, uint32_t span) {
    return ((uint32_t)(tick - d.x - lo) <= span ? 1u : 0u) | ((uint32_t)(tick - d.y - lo) <= span ? 0x100u : 0u) |
           ((uint32_t)(tick - d.z - lo) <= span ? 0x10000u : 0u) | ((uint32_t)(tick - d.w - lo) <= span ? 0x1000000u : 0u);
}
// ---- routine immunisation in a quad (reference model.py:1825-1854) -----------------------------------------------
// The reference subtracts the step from every alive, not chronically missed agent's ri_timer on every RI tick; an agent is
// eligible when the new timer lies in (-step, 0] ([-step, 0] on the first RI tick).  Here the countdown is lazy
// (lpk_tick_args.ri_lazy_k): nothing is written.  Eligible agents are the few in the age window; they go to the ring and
// take their two draws in the handler, after their disease-state step.
// The sweep reads the agent's precomputed eligibility tick (lpk_people.ri_k: one byte, built by lpk_hot_build / births)
// and compares it with the number of today's RI tick on byte lanes.
// append the agents of the two quads a lane owns in a row pair (B = A + 128 agents); F = per-agent flag bytes
// (bit 0 candidate, 1 agenda day, 2 death, 3 RI, 4 SIA); returns how many
__device__ __forceinline__ int q_push_pair(uint2 *q, uint32_t *tail, uint32_t idxA, int nd, uint32_t FA, uint32_t FB) {
    uint32_t m = (zero_bytes(FA) ^ 0x01010101u) | ((zero_bytes(FB) ^ 0x01010101u) << 4);
    const int cnt = __popc(m);
    while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1u;
        const bool rowB = (bit & 4) != 0;
        const uint32_t f = ((rowB ? FB : FA) >> (bit & 24)) & 0x1Fu;
        const uint32_t pos = atomicAdd(tail, 1u) & (QCAP - 1);
        q[pos] = make_uint2(idxA + (uint32_t)(bit >> 3) + (rowB ? 128u : 0u), ((uint32_t)nd & 0xFFFFu) | (f << 16));
    }
    return cnt;
}

// the same for a pair whose agents live in several nodes: nA / nB = the quads' four int16 node ids
__device__ __forceinline__ int q_push_pair_nodes(uint2 *q, uint32_t *tail, uint32_t idxA, uint2 nA, uint2 nB, uint32_t FA, uint32_t FB) {
    uint32_t m = (zero_bytes(FA) ^ 0x01010101u) | ((zero_bytes(FB) ^ 0x01010101u) << 4);
    const int cnt = __popc(m);
    while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1u;
        const bool rowB = (bit & 4) != 0;
        const int k = bit >> 3;
        const uint32_t f = ((rowB ? FB : FA) >> (bit & 24)) & 0x1Fu;
        const uint2 nn = rowB ? nB : nA;
        const uint32_t nd = (((k & 2) ? nn.y : nn.x) >> (16 * (k & 1))) & 0xFFFFu;
        const uint32_t pos = atomicAdd(tail, 1u) & (QCAP - 1);
        q[pos] = make_uint2(idxA + (uint32_t)k + (rowB ? 128u : 0u), nd | (f << 16));
    }
    return cnt;
}
// pre-test with a scale per agent (agents of different nodes in one quad)
__device__ __forceinline__ uint32_t hot_pretest_each(uint32_t h, uint32_t xa, uint32_t xb, float t0, float t1, float t2, float t3) {
    const uint32_t KM = 0x1FFFFFFFu, k23 = 0x4B000000u;
    const float c = 8388609.0f;
    const float d0 = __uint_as_float((h << 21) & KM), d1 = __uint_as_float((h << 13) & KM), d2 = __uint_as_float((h << 5) & KM),
                d3 = __uint_as_float(h >> 3);
    const float u0 = __uint_as_float(__byte_perm(xa, k23, 0x7610)), u1 = __uint_as_float(__byte_perm(xa, k23, 0x7632));
    const float u2 = __uint_as_float(__byte_perm(xb, k23, 0x7610)), u3 = __uint_as_float(__byte_perm(xb, k23, 0x7632));
    return ((u0 < fmaf(d0, t0, c)) ? 1u : 0u) | ((u1 < fmaf(d1, t1, c)) ? 0x100u : 0u) | ((u2 < fmaf(d2, t2, c)) ? 0x10000u : 0u) |
           ((u3 < fmaf(d3, t3, c)) ? 0x1000000u : 0u);
}

// ------------------------------------------------------------------ mixed pair: 256 live slots in SEVERAL nodes (appended
// newborn cohorts are node-major runs of a few hundred agents; node boundaries of the initial population).  Same byte-lane
// sweep as a node-uniform pair, with the node id read per agent (2 B) and the exposure scale gathered per agent (the ids
// come in runs, so the gathers hit one or two cache lines per warp).  tau is clamped where every susceptible passes
// anyway, which also covers tau = "everybody" without a special case.  Round 1 sent these pairs through the one-agent-
// at-a-time path: ~50 us of a warp's time per 2048 agents, 26 % of the table after seven years of births.
template <bool kDeaths, bool kRI, bool kSIA>
__device__ __noinline__ int mixed_pair(const PassParams &pp, uint2 *q, uint32_t *q_tail, uint32_t gp, uint32_t hA, uint32_t hB, int md, int rm,
                                       uint64_t ctr, int lane) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const int tick = A.tick, e0 = P.risk_e0;
    const int64_t bA = (int64_t)gp * 256 + lane * 4, bB = bA + 128;
    const uint2 nA = __ldg(reinterpret_cast<const uint2 *>(P.node_id + bA)), nB = __ldg(reinterpret_cast<const uint2 *>(P.node_id + bB));
    uint32_t cA = 0u, cB = 0u;
    if (A.flags & LPK_F_PENDING) {
        const float clampv = ldexpf(1.0f, e0 + 5), scale = ldexpf(1.0f, 95 - e0);
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint2 nn = (k & 4) ? nB : nA;
            const int nd = (int)(int16_t)((((k & 2) ? nn.y : nn.x) >> (16 * (k & 1))) & 0xFFFFu);
            const float tau = nd >= 0 ? __ldg(&A.q_prev[nd]) : 0.f;
            t[k] = fminf(fmaxf(tau, 0.f), clampv) * scale;
        }
        uint32_t x[4];
        philox_sweep(pp, (uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)(tick - 1), LPK_STAGE_EXPOSE, x);
        cA = hot_pretest_each(hA, x[0], x[1], t[0], t[1], t[2], t[3]);
        cB = hot_pretest_each(hB, x[2], x[3], t[4], t[5], t[6], t[7]);
    }
    const uint32_t today = hot_today(tick);
    const uint32_t vA = hot_due_word(hA, today), vB = hot_due_word(hB, today);
    uint32_t dmA = 0u, dmB = 0u, eA = 0u, eB = 0u, sA = 0u, sB = 0u;
    if (kDeaths && md <= tick) {
        const int4 dA = __ldg(reinterpret_cast<const int4 *>(P.date_of_death + bA));
        const int4 dB = __ldg(reinterpret_cast<const int4 *>(P.date_of_death + bB));
        const uint32_t aA = hot_mask_alive(hA), aB = hot_mask_alive(hB);
        dmA = death_mask(dA, tick) & aA;
        dmB = death_mask(dB, tick) & aB;
        const int left = __reduce_min_sync(LPK_FULL, min(min_dod_left(dA, aA & ~dmA), min_dod_left(dB, aB & ~dmB)));
        if (lane == 0) P.pair_min_dod[gp] = left;
    }
    if (kRI && rm >= A.ri_lazy_k * A.ri_step) {
        const uint32_t ri_today = ((uint32_t)(A.ri_lazy_k + 1) & 0xFFu) * 0x01010101u;
        eA = zero_bytes(*reinterpret_cast<const uint32_t *>(P.ri_k + bA) ^ ri_today) & ~dmA;
        eB = zero_bytes(*reinterpret_cast<const uint32_t *>(P.ri_k + bB) ^ ri_today) & ~dmB;
    }
    if (kSIA) {
        const uint32_t span = (uint32_t)(A.sia_max_age - A.sia_min_age);
        uint32_t tgA = 0u, tgB = 0u;  // agents whose node is targeted by today's campaign
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int na = (int)(int16_t)((((k & 2) ? nA.y : nA.x) >> (16 * (k & 1))) & 0xFFFFu);
            const int nb = (int)(int16_t)((((k & 2) ? nB.y : nB.x) >> (16 * (k & 1))) & 0xFFFFu);
            if (na >= 0 && __ldg(&A.sia_targeted[na]) != 0) tgA |= 1u << (8 * k);
            if (nb >= 0 && __ldg(&A.sia_targeted[nb]) != 0) tgB |= 1u << (8 * k);
        }
        if (tgA | tgB) {
            const uint32_t mA = *reinterpret_cast<const uint32_t *>(P.chronically_missed + bA);
            const uint32_t mB = *reinterpret_cast<const uint32_t *>(P.chronically_missed + bB);
            sA = sia_age_mask(__ldg(reinterpret_cast<const int4 *>(P.date_of_birth + bA)), tick, A.sia_min_age, span) & hot_mask_alive(hA) & ~dmA & ~mA & tgA;
            sB = sia_age_mask(__ldg(reinterpret_cast<const int4 *>(P.date_of_birth + bB)), tick, A.sia_min_age, span) & hot_mask_alive(hB) & ~dmB & ~mB & tgB;
        }
    }
    const uint32_t FA = (cA & hot_mask_S(hA)) | (zero_bytes(vA) << 1) | (dmA << 2) | (eA << 3) | (sA << 4);
    const uint32_t FB = (cB & hot_mask_S(hB)) | (zero_bytes(vB) << 1) | (dmB << 2) | (eB << 3) | (sB << 4);
    return (FA | FB) ? q_push_pair_nodes(q, q_tail, (uint32_t)bA, nA, nB, FA, FB) : 0;
}

// ------------------------------------------------------------------ general pair (out of line): 256 agents that are not
// all in one node (node boundaries, appended cohorts, the table's tail).  One agent at a time; every agent with
// something to do goes to the ring with its own node.  h: the lane's agenda words of the two rows.
template <bool kDeaths, bool kRI, bool kSIA>
__device__ __noinline__ int general_pair(const PassParams &pp, uint2 *q, uint32_t *q_tail, int64_t gp, int64_t n, uint32_t hA, uint32_t hB,
                                         int lane) {
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const bool pending = (A.flags & LPK_F_PENDING) != 0;
    const int tick = A.tick, e0 = P.risk_e0;
    const uint32_t today = hot_today(tick) & 0xFFu;
    int mine = 0;
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        const int64_t b = gp * 256 + r * 128 + lane * 4;
        const int valid = quad_valid(b, n);
        const uint32_t h = r ? hB : hA;
        if (!valid || h == HOT_DEAD * 0x01010101u) continue;
        uint32_t x[4];
        bool have_x = false;
#pragma unroll 1
        for (int k = 0; k < valid; ++k) {
            const uint32_t hb = (h >> (8 * k)) & 0xFFu;
            if (hb == HOT_DEAD) continue;
            const int64_t i = b + k;
            const int nd = P.node_id[i];
            uint32_t fl = 0u;
            if (pending && hot_is_S(hb)) {  // susceptible: pre-test of tick t-1's exposure trial
                const float tau = __ldg(&A.q_prev[nd]);
                if (tau > 0.f) {
                    if (!have_x) {
                        const uint64_t c = expose_ctr((uint64_t)b + A.id_base);
                        philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(tick - 1), LPK_STAGE_EXPOSE, (uint32_t)A.seed,
                                      (uint32_t)(A.seed >> 32), x);
                        have_x = true;
                    }
                    const float U = 8388608.0f + (float)half_word(x, expose_hw((uint64_t)i + A.id_base));
                    if (U < fmaf(risk_code_ub((int)(hb & 63u), e0), tau * 65536.0f, 8388609.0f)) fl |= EV_CAND;
                }
            }
            if ((hb | 0x40u) == today) fl |= EV_FIRE;
            bool dying = false;
            if (kDeaths && P.date_of_death[i] <= tick) { fl |= EV_DEATH; dying = true; }
            if ((kRI || kSIA) && !dying && (!kSIA || P.chronically_missed[i] != 1)) {
                if (kRI && P.ri_k[i] == (uint8_t)(A.ri_lazy_k + 1)) fl |= EV_RI;
                if (kSIA && (uint32_t)(tick - P.date_of_birth[i] - A.sia_min_age) <= (uint32_t)(A.sia_max_age - A.sia_min_age) &&
                    A.sia_targeted[nd] != 0)
                    fl |= EV_SIA;
            }
            if (fl) {
                const uint32_t pos = atomicAdd(q_tail, 1u) & (QCAP - 1);
                q[pos] = make_uint2((uint32_t)i, ((uint32_t)nd & 0xFFFFu) | fl);
                ++mine;
            }
        }
    }
    return mine;
}

// ------------------------------------------------------------------ the pass
// Unit of work: 8 consecutive PAIRS of 128-agent rows (2048 agents); a lane owns its quad in the even row (A) and in the
// odd row (B) of every pair -- also the unit of the exposure RNG (one Philox block = 8 x 16-bit high halves).  Warps claim
// runs of units from a global counter (guided self-scheduling) and fetch a unit's 2 KB of agenda bytes with ONE bulk copy
// (cp.async.bulk on an mbarrier, issued by an elected lane one unit ahead) into a private double-buffered stage.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
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
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0u;
}
// ask the copy engine to bring a block of global memory into L2 (no destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// two blocks at once (counters ca, cb; same tick / stage): independent dependency chains in one basic block
__device__ __forceinline__ void philox_sweep2(const PassParams &pp, uint64_t ca, uint64_t cb, uint32_t c2, uint32_t c3, uint32_t x[4],
                                              uint32_t y[4]) {
    uint32_t a0 = (uint32_t)ca, a1 = (uint32_t)(ca >> 32), a2 = c2, a3 = c3;
    uint32_t b0 = (uint32_t)cb, b1 = (uint32_t)(cb >> 32), b2 = c2, b3 = c3;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t ah0 = __umulhi(0xD2511F53u, a0), al0 = 0xD2511F53u * a0, ah1 = __umulhi(0xCD9E8D57u, a2), al1 = 0xCD9E8D57u * a2;
        const uint32_t bh0 = __umulhi(0xD2511F53u, b0), bl0 = 0xD2511F53u * b0, bh1 = __umulhi(0xCD9E8D57u, b2), bl1 = 0xCD9E8D57u * b2;
        a0 = ah1 ^ a1 ^ pp.rk[2 * r]; a1 = al1; a2 = ah0 ^ a3 ^ pp.rk[2 * r + 1]; a3 = al0;
        b0 = bh1 ^ b1 ^ pp.rk[2 * r]; b1 = bl1; b2 = bh0 ^ b3 ^ pp.rk[2 * r + 1]; b3 = bl0;
    }
    x[0] = a0; x[1] = a1; x[2] = a2; x[3] = a3;
    y[0] = b0; y[1] = b1; y[2] = b2; y[3] = b3;
}

template <int kWarps>
struct PassSmem {
    static constexpr int kOffQueue = 0;
    static constexpr int kOffStage = kOffQueue + kWarps * QCAP * 8;          // 2 x 2 KB per warp
    static constexpr int kOffBars = kOffStage + kWarps * 2 * LPK_UNIT_AGENTS;  // 2 mbarriers per warp
    static constexpr int kOffTail = kOffBars + kWarps * 2 * 8;
    static constexpr int kOffAcc = (kOffTail + kWarps * 4 + 15) & ~15;
    static constexpr int kBytes = kOffAcc + kWarps * (int)sizeof(WarpAcc) + 32;
};

template <bool kDeaths, bool kRI, bool kSIA, int kWarps, int kOcc>
__global__ void __launch_bounds__(kWarps * 32, kOcc) k_tick_pass(const __grid_constant__ PassParams pp) {
    typedef PassSmem<kWarps> L;
    extern __shared__ __align__(128) unsigned char smem[];
    const lpk_people &P = pp.P;
    const lpk_tick_args &A = pp.A;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (pp.unit_ctr_next && blockIdx.x == 0 && threadIdx.x == 0) *pp.unit_ctr_next = 0u;  // nobody claims from it during this launch
    const int64_t n = A.counts[1];
    const bool pending = (A.flags & LPK_F_PENDING) != 0;
    const int tick = A.tick, e0 = P.risk_e0;
    const uint64_t ctr_base = ((A.id_base >> 8) << 5) + (uint64_t)lane;  // Philox counter of pair 0 for this lane
    const uint32_t total_pairs = (uint32_t)((n + 255) >> 8);
    const uint32_t full_pairs = (uint32_t)(n >> 8);  // pairs whose 256 slots are all in use
    const uint32_t upl = pp.unit_log, unit_pairs = 1u << upl, unit_bytes = 256u << upl;
    const uint32_t n_units = (total_pairs + unit_pairs - 1u) >> upl;
    const uint32_t today = hot_today(tick);
    const float tau_all = ldexpf(1.0f, e0);  // from here on every susceptible passes the pre-test anyway (bound of code 0 x tau >= 1)

    const uint32_t *stage = reinterpret_cast<const uint32_t *>(smem + L::kOffStage + warp * 2 * LPK_UNIT_AGENTS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::kOffBars) + warp * 2;
    WarpQueue Q;
    Q.q = reinterpret_cast<uint2 *>(smem + L::kOffQueue) + warp * QCAP;
    Q.tail = reinterpret_cast<uint32_t *>(smem + L::kOffTail) + warp;
    Q.head = 0u;
    Q.count = 0;
    Q.loaded = false;
    Q.pend_e = make_uint2(0u, 0u);
    Q.pend.rec = 0ull; Q.pend.rk = 0.f; Q.pend.inf = 0.f;
    Q.acc = reinterpret_cast<WarpAcc *>(smem + L::kOffAcc) + warp;
    if (lane == 0) {
        Q.acc->node = -1;
        acc_clear(Q.acc);
        *Q.tail = 0u;
        mbar_init(&bars[0], 1u);
        mbar_init(&bars[1], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // Work distribution: guided self-scheduling -- a run is 1 / (3 x warps in the grid) of the units still unclaimed (at
    // most 64, at least 1): long runs while there is plenty of work, so that a warp stays in one node and its per-node
    // accumulators are flushed rarely, single units at the end for balance.  The units behind the node-contiguous initial
    // population (appended cohorts: mixed-node pairs, the slow general path) are handed out FIRST, one per claim.
    const uint32_t kNoUnit = 0xFFFFFFFFu;
    const uint32_t guide = 3u * gridDim.x * kWarps;
    const uint32_t stream_units = A.uniform_agents > 0 ? (uint32_t)(A.uniform_agents >> (8 + upl)) : n_units;
    const uint32_t tail_units = n_units - (stream_units < n_units ? stream_units : n_units);
    uint32_t run_next = 0u, run_end = 0u;  // warp-uniform: claim indices of the current run not yet taken
    auto next_unit = [&]() -> uint32_t {   // warp-uniform result
        if (run_next >= run_end) {
            uint32_t first = 0u, cnt = 0u;
            if (lane == 0) {
                const uint32_t seen = *reinterpret_cast<volatile uint32_t *>(pp.unit_ctr);
                if (seen < n_units) {
                    const uint32_t left = n_units - seen;
                    cnt = seen < tail_units ? 1u : left / guide;
                    cnt = cnt < 1u ? 1u : (cnt > 64u ? 64u : cnt);
                    first = atomicAdd(pp.unit_ctr, cnt);
                } else {
                    first = n_units;
                }
            }
            first = __shfl_sync(LPK_FULL, first, 0);
            cnt = __shfl_sync(LPK_FULL, cnt, 0);
            if (first >= n_units) return kNoUnit;
            run_next = first;
            run_end = first + cnt < n_units ? first + cnt : n_units;
        }
        const uint32_t c = run_next++;
        return c < tail_units ? n_units - 1u - c : c - tail_units;
    };
    auto fetch = [&](uint32_t u, int buf) {  // one lane asks the copy engine for the unit's agenda bytes
        if (elect_one()) {
            const uint32_t bar = smem_u32(bars) + (uint32_t)buf * 8u;
            fence_proxy_async_smem();  // the warp's reads of this buffer (previous use) precede the engine's writes
            mbar_arrive_expect_tx(bar, unit_bytes);
            tma_load(smem_u32(stage) + (uint32_t)buf * LPK_UNIT_AGENTS, P.hot + (int64_t)u * unit_bytes, unit_bytes, bar);
            // the columns the special days read for (nearly) every pair of the unit: into L2 a unit ahead, so that the loads in
            // the tile loop cost an L2 hit instead of a DRAM round trip (profiles/r2: the RI day was latency bound, 1.1 ms)
            const int64_t a0 = (int64_t)u * unit_bytes;
            if ((kRI || kSIA) && a0 + unit_bytes <= P.capacity) {
                if (kRI) tma_prefetch_l2(P.ri_k + a0, unit_bytes);
                if (kSIA) { tma_prefetch_l2(P.chronically_missed + a0, unit_bytes); tma_prefetch_l2(P.date_of_birth + a0, 4u * unit_bytes); }
            }
        }
    };

    int tc_node = -2;  // one-entry cache of the node's exposure scale (and the campaign's target flag): a warp stays in one node for long
    int tc_mode = 0;   // 0 no force of infection, 1 pre-test, 2 every susceptible is a candidate
    float tc_tauS = 0.f;
    bool tc_sia = false;
    const int ri_debt = kRI ? A.ri_lazy_k * A.ri_step : 0;  // an agent whose stored timer is below this can never be eligible again
    const uint32_t ri_today = ((uint32_t)(A.ri_lazy_k + 1) & 0xFFu) * 0x01010101u;  // today's RI tick as lpk_people.ri_k counts them
    const int sia_lo = kSIA ? A.sia_min_age : 0;
    const uint32_t sia_span = kSIA ? (uint32_t)(A.sia_max_age - A.sia_min_age) : 0u;

    // per-unit metadata, loaded a unit ahead like the agenda bytes: the unit's 4 tile nodes (lanes 0-3), on vital-dynamics
    // ticks its 8 earliest death dates and on RI ticks its 8 largest RI timers (lanes 0-7)
    int tnv = -1, mdv = INT_MAX, rmv = INT_MIN;
    auto fetch_meta = [&](uint32_t u) {
        const uint32_t g0 = u << upl;
        tnv = -1; mdv = INT_MAX; rmv = INT_MIN;
        if (P.tile_node && (uint32_t)lane < (unit_pairs >> 1) && g0 + 2u * (uint32_t)lane < total_pairs) tnv = __ldg(&P.tile_node[(g0 >> 1) + lane]);
        if (kDeaths && (uint32_t)lane < unit_pairs && g0 + (uint32_t)lane < total_pairs) mdv = P.pair_min_dod[g0 + lane];
        if (kRI && (uint32_t)lane < unit_pairs && g0 + (uint32_t)lane < total_pairs) rmv = P.pair_ri_max[g0 + lane];
    };
    uint32_t u_cur = next_unit();
    if (u_cur != kNoUnit) { fetch(u_cur, 0); fetch_meta(u_cur); }
    uint32_t par0 = 0u, par1 = 0u;  // phase parity of the two buffers
    int buf = 0;
#pragma unroll 1
    while (u_cur != kNoUnit) {
        const uint32_t u_nxt = next_unit();
        const uint32_t gp0 = u_cur << upl;
        const int tnv_cur = tnv, mdv_cur = mdv, rmv_cur = rmv;
        if (u_nxt != kNoUnit) { fetch(u_nxt, buf ^ 1); fetch_meta(u_nxt); }
        mbar_wait(&bars[buf], buf ? par1 : par0);
        if (buf) par1 ^= 1u; else par0 ^= 1u;
        const uint32_t *src = stage + buf * (LPK_UNIT_AGENTS / 4);
#pragma unroll 1
        for (int tp = 0; tp < (int)(unit_pairs >> 1); ++tp) {  // one TILE (two pairs, 16 agents per lane) per iteration: the two
            const uint32_t gp = gp0 + 2u * (uint32_t)tp;    // Philox blocks are independent chains and interleave
            if (gp >= total_pairs) break;
            const int tn = __shfl_sync(LPK_FULL, tnv_cur, tp);
            const uint32_t hA0 = src[tp * 128 + lane], hB0 = src[tp * 128 + 32 + lane];
            const uint32_t hA1 = src[tp * 128 + 64 + lane], hB1 = src[tp * 128 + 96 + lane];
            if (tn < 0) {  // several nodes in the tile: the mixed sweep for pairs of 256 live slots, one agent at a time at the tail
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {
                    if (gp + (uint32_t)j >= total_pairs) break;
                    const uint32_t hA = j ? hA1 : hA0, hB = j ? hB1 : hB0;
                    int mine;
                    if (gp + (uint32_t)j < full_pairs)
                        mine = mixed_pair<kDeaths, kRI, kSIA>(pp, Q.q, Q.tail, gp + (uint32_t)j, hA, hB, __shfl_sync(LPK_FULL, mdv_cur, 2 * tp + j),
                                                              __shfl_sync(LPK_FULL, rmv_cur, 2 * tp + j), ctr_base + ((uint64_t)(gp + (uint32_t)j) << 5), lane);
                    else
                        mine = general_pair<kDeaths, kRI, kSIA>(pp, Q.q, Q.tail, (int64_t)gp + j, n, hA, hB, lane);
                    q_commit(pp, Q, mine, lane);
                }
                continue;
            }
            if (tn != tc_node) {
                tc_node = tn;
                const float tau = pending ? __ldg(&A.q_prev[tn]) : 0.f;
                tc_mode = !(tau > 0.f) ? 0 : (tau >= tau_all ? 2 : 1);
                tc_tauS = hot_tau_scale(tau, e0);
                if (kSIA) tc_sia = __ldg(&A.sia_targeted[tn]) != 0;
            }
            const int64_t bA0 = (int64_t)gp * 256 + lane * 4;  // quads: A0 = bA0, B0 = +128, A1 = +256, B1 = +384
            // ---- exposure trial of tick t-1: pre-test on the risk bound; candidates are decided by the ring handler
            uint32_t cA0 = 0u, cB0 = 0u, cA1 = 0u, cB1 = 0u;
            if (tc_mode == 1) {
                const uint64_t c = ctr_base + ((uint64_t)gp << 5);
                uint32_t x[4], y[4];
                philox_sweep2(pp, c, c + 32u, (uint32_t)(tick - 1), LPK_STAGE_EXPOSE, x, y);
                cA0 = hot_pretest(hA0, x[0], x[1], tc_tauS);
                cB0 = hot_pretest(hB0, x[2], x[3], tc_tauS);
                cA1 = hot_pretest(hA1, y[0], y[1], tc_tauS);
                cB1 = hot_pretest(hB1, y[2], y[3], tc_tauS);
            } else if (tc_mode == 2) {
                cA0 = cB0 = cA1 = cB1 = 0x01010101u;
            }
            // ---- agenda: exposed / infectious agents whose day is today
            const uint32_t vA0 = hot_due_word(hA0, today), vB0 = hot_due_word(hB0, today);
            const uint32_t vA1 = hot_due_word(hA1, today), vB1 = hot_due_word(hB1, today);
            uint32_t ev = cA0 | cB0 | cA1 | cB1 | any_zero_byte(vA0) | any_zero_byte(vB0) | any_zero_byte(vA1) | any_zero_byte(vB1);
            // ---- tick t: deaths, RI timers, campaign window (extra columns only where they can matter)
            uint32_t xA0 = 0u, xB0 = 0u, xA1 = 0u, xB1 = 0u;  // per-agent flag bits 2 (death), 3 (RI), 4 (SIA)
            if (kDeaths || kRI || kSIA) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t hA = j ? hA1 : hA0, hB = j ? hB1 : hB0;
                    const int64_t bA = bA0 + 256 * j, bB = bA + 128;
                    uint32_t dmA = 0u, dmB = 0u, eA = 0u, eB = 0u, sA = 0u, sB = 0u;
                    if (kDeaths) {
                        const int md = __shfl_sync(LPK_FULL, mdv_cur, 2 * tp + j);
                        if (md <= tick) {  // warp-uniform: somebody in this pair can die today
                            const int4 dA = __ldg(reinterpret_cast<const int4 *>(P.date_of_death + bA));
                            const int4 dB = __ldg(reinterpret_cast<const int4 *>(P.date_of_death + bB));
                            const uint32_t aA = hot_mask_alive(hA), aB = hot_mask_alive(hB);
                            dmA = death_mask(dA, tick) & aA;
                            dmB = death_mask(dB, tick) & aB;
                            const int left = __reduce_min_sync(LPK_FULL, min(min_dod_left(dA, aA & ~dmA), min_dod_left(dB, aB & ~dmB)));
                            if (lane == 0) P.pair_min_dod[gp + j] = left;
                        }
                    }
                    // RI: only pairs in which somebody's timer has not run out for good (stored >= debt) can hold an eligible agent
                    if (kRI && __shfl_sync(LPK_FULL, rmv_cur, 2 * tp + j) >= ri_debt) {  // (the dead are turned away by the handler)
                        eA = zero_bytes(*reinterpret_cast<const uint32_t *>(P.ri_k + bA) ^ ri_today) & ~dmA;
                        eB = zero_bytes(*reinterpret_cast<const uint32_t *>(P.ri_k + bB) ^ ri_today) & ~dmB;
                    }
                    if (kSIA && tc_sia) {
                        const uint32_t mA = *reinterpret_cast<const uint32_t *>(P.chronically_missed + bA);
                        const uint32_t mB = *reinterpret_cast<const uint32_t *>(P.chronically_missed + bB);
                        const uint32_t aA = hot_mask_alive(hA) & ~dmA, aB = hot_mask_alive(hB) & ~dmB;
                        {
                            sA = sia_age_mask(__ldg(reinterpret_cast<const int4 *>(P.date_of_birth + bA)), tick, sia_lo, sia_span) & aA & ~mA;
                            sB = sia_age_mask(__ldg(reinterpret_cast<const int4 *>(P.date_of_birth + bB)), tick, sia_lo, sia_span) & aB & ~mB;
                        }
                    }
                    const uint32_t fA = (dmA << 2) | (eA << 3) | (sA << 4), fB = (dmB << 2) | (eB << 3) | (sB << 4);
                    if (j) { xA1 = fA; xB1 = fB; } else { xA0 = fA; xB0 = fB; }
                }
                ev |= xA0 | xB0 | xA1 | xB1;
            }
            int mine = 0;
            if (ev) {  // rare per lane (a few per cent), common per warp: keep it short
                const uint32_t FA0 = (cA0 & hot_mask_S(hA0)) | (zero_bytes(vA0) << 1) | xA0, FB0 = (cB0 & hot_mask_S(hB0)) | (zero_bytes(vB0) << 1) | xB0;
                const uint32_t FA1 = (cA1 & hot_mask_S(hA1)) | (zero_bytes(vA1) << 1) | xA1, FB1 = (cB1 & hot_mask_S(hB1)) | (zero_bytes(vB1) << 1) | xB1;
                if (FA0 | FB0) mine = q_push_pair(Q.q, Q.tail, (uint32_t)bA0, tn, FA0, FB0);
                if (FA1 | FB1) mine += q_push_pair(Q.q, Q.tail, (uint32_t)bA0 + 256u, tn, FA1, FB1);
            }
            q_commit(pp, Q, mine, lane);
        }
        u_cur = u_nxt;
        buf ^= 1;
    }
    __syncwarp();
    if (Q.loaded) active_process(pp, Q.pend_e, Q.pend, true, Q.acc, lane);
    {
        const bool valid = lane < Q.count;
        uint2 e = make_uint2(0u, 0u);
        HotPre last;
        last.rec = 0ull; last.rk = 0.f; last.inf = 0.f;
        if (valid) { e = Q.q[(Q.head + lane) & (QCAP - 1)]; last = hot_preload(P, (int64_t)e.x, e.y); }
        active_process(pp, e, last, valid, Q.acc, lane);
    }
    if (lane == 0) acc_flush(pp, Q.acc);
}

template <bool kDeaths, bool kRI, bool kSIA, int kWarps, int kOcc>
static int launch_pass(const PassParams &pp, cudaStream_t st) {
    typedef PassSmem<kWarps> L;
    static bool configured = false;
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(k_tick_pass<kDeaths, kRI, kSIA, kWarps, kOcc>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes),
                 "tick_pass smem");
        configured = true;
    }
    const int grid = lpk_sm_count() * kOcc;
    if (!pp.unit_ctr_next) CUDA_TRY(cudaMemsetAsync(pp.unit_ctr, 0, sizeof(uint32_t), st), "tick_pass work counter");
    k_tick_pass<kDeaths, kRI, kSIA, kWarps, kOcc><<<grid, kWarps * 32, L::kBytes, st>>>(pp);
    return LPK_OK;
}

extern "C" int lpk_tick_pass(const lpk_people *people, const lpk_tick_args *args, void *stream) {
    REQUIRE(people && args, "tick_pass null struct");
    const lpk_people &P = *people;
    const lpk_tick_args &A = *args;
    REQUIRE(A.n_nodes > 0 && A.n_strains >= 1 && A.n_strains <= LPK_MAX_STRAINS, "tick_pass sizes");
    REQUIRE(P.capacity > 0 && P.capacity < (1ll << 32) && A.counts, "tick_pass counts (tables hold < 2^32 slots)");
    REQUIRE(P.disease_state && P.strain && P.exposure_timer && P.infection_timer && P.paralysis_timer && P.potentially_paralyzed &&
                P.paralyzed && P.ipv_protected && P.node_id && P.acq_risk_multiplier && P.daily_infectivity, "tick_pass agent columns");
    REQUIRE(P.rec && ALIGNED(P.rec, 8), "tick_pass event records (lpk_hot_build fills them)");
    REQUIRE(P.hot && ALIGNED(P.hot, 16), "tick_pass agenda bytes (lpk_hot_build fills them; 16-byte aligned, padded to 2048 agents)");
    REQUIRE((A.flags & LPK_F_STAGES) != 0, "tick_pass always runs the stages of its tick (LPK_F_STAGES)");
    REQUIRE((A.id_base & 255) == 0, "tick_pass id_base must be a multiple of 256");
    REQUIRE(A.work_counter, "tick_pass work counter (caller-owned device uint32)");
    REQUIRE(A.new_potential && A.new_paralyzed && A.beta_fx && A.exposure_fx && A.sus && A.risk_hist && A.R_cur && A.tx_hits,
            "tick_pass stage outputs");
    REQUIRE(A.E_cur && A.I_cur && A.tx_hits_by_strain && A.new_exposed_prev && A.new_exposed_by_strain_prev, "tick_pass census");
    if (A.flags & LPK_F_PENDING) REQUIRE(A.q_prev && A.cdf_prev, "tick_pass pending exposure inputs");
    const bool deaths = (A.flags & LPK_F_DEATHS) != 0, ri = (A.flags & LPK_F_RI) != 0, sia = (A.flags & LPK_F_SIA) != 0;
    if (sia) REQUIRE(P.date_of_birth && ALIGNED(P.date_of_birth, 16) && P.chronically_missed && ALIGNED(P.chronically_missed, 4) &&
                         A.sia_targeted && A.vx_prob_sia && A.sia_vaccinated && A.sia_protected && A.sia_new_exposed_by_strain &&
                         A.new_exposed && A.new_exposed_by_strain && A.sia_max_age >= A.sia_min_age && A.sia_strain >= 0 &&
                         A.sia_strain < A.n_strains, "tick_pass SIA");
    if (deaths) REQUIRE(P.date_of_death && ALIGNED(P.date_of_death, 16) && P.pair_min_dod && A.deaths && A.dead_pp && A.dead_par,
                        "tick_pass deaths");
    if (A.ri_lazy_k) REQUIRE(A.ri_lazy_k > 0 && A.ri_step > 0 && P.ri_timer && P.chronically_missed, "tick_pass lazy RI countdown");
    if (ri) REQUIRE(P.pair_ri_max && P.ri_k && ALIGNED(P.ri_k, 16) && A.ri_lazy_k < 254 && P.ri_timer && P.chronically_missed && A.vx_prob_ri &&
                        A.vx_prob_ipv && A.ri_vaccinated && A.ri_protected && A.ipv_vaccinated && A.new_exposed &&
                        A.new_exposed_by_strain && A.ri_new_exposed_by_strain && A.ri_step > 0 && A.ri_strain >= 0 &&
                        A.ri_strain < A.n_strains, "tick_pass RI");
    PassParams pp;
    pp.P = P;
    pp.A = A;
    pp.unit_ctr = A.work_counter;
    pp.unit_ctr_next = A.work_counter_next;
    REQUIRE(A.work_counter_next != A.work_counter, "tick_pass work counters must differ");
    for (int r = 0; r < 10; ++r) {
        pp.rk[2 * r] = (uint32_t)A.seed + (uint32_t)r * 0x9E3779B9u;
        pp.rk[2 * r + 1] = (uint32_t)(A.seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    {
        static int dbg = -1;
        if (dbg < 0) { const char *e = getenv("LPK_PASS_DEBUG"); dbg = e ? atoi(e) : 0; }
        pp.debug = (uint32_t)dbg;
    }
    {   // pairs per work unit: 8; LPK_PASS_UNIT_LOG=1|2 for experiments (smaller units did not help small tables: DESIGN.md section 4)
        static int forced = -2;
        if (forced == -2) { const char *e = getenv("LPK_PASS_UNIT_LOG"); forced = e ? atoi(e) : -1; }
        pp.unit_log = (uint32_t)((forced >= 1 && forced <= LPK_UNIT_LOG) ? forced : LPK_UNIT_LOG);
    }
    cudaStream_t st = as_stream(stream);
    int rc;
    // block shape: warps per block x blocks per SM (LPK_PASS_SHAPE: experiments)
    static int shape = -1;
    if (shape < 0) { const char *e = getenv("LPK_PASS_SHAPE"); shape = e ? atoi(e) : 1; }
#define LPK_DISPATCH(W, O)                                                                          \
    do {                                                                                            \
        if (sia) {                                                                                  \
            if (deaths && ri) rc = launch_pass<true, true, true, W, O>(pp, st);                     \
            else if (deaths) rc = launch_pass<true, false, true, W, O>(pp, st);                     \
            else if (ri) rc = launch_pass<false, true, true, W, O>(pp, st);                         \
            else rc = launch_pass<false, false, true, W, O>(pp, st);                                \
        } else if (deaths && ri) rc = launch_pass<true, true, false, W, O>(pp, st);                 \
        else if (deaths) rc = launch_pass<true, false, false, W, O>(pp, st);                        \
        else if (ri) rc = launch_pass<false, true, false, W, O>(pp, st);                            \
        else rc = launch_pass<false, false, false, W, O>(pp, st);                                   \
    } while (0)
    if (shape == 0) LPK_DISPATCH(8, 3);
    else if (shape == 2) LPK_DISPATCH(9, 2);
    else LPK_DISPATCH(8, 2);
#undef LPK_DISPATCH
    if (rc != LPK_OK) return rc;
    CUDA_TRY(cudaGetLastError(), "lpk_tick_pass");
    return LPK_OK;
}

// ------------------------------------------------------------------ canonical columns <-> agenda bytes
__global__ void __launch_bounds__(256) k_hot_build(lpk_people P, int64_t n_slots, int64_t padded, int t_next, int32_t *status) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += (int64_t)gridDim.x * blockDim.x) {
        uint8_t h = HOT_DEAD;
        if (i < n_slots) {
            bool over = false;
            h = hot_build_agent(P, i, t_next, P.risk_e0, &over);
            if (over && status) *status = 2;
        }
        P.hot[i] = h;
    }
}
// earliest date of death among the alive agents of every pair: one warp per pair
__global__ void __launch_bounds__(256) k_pair_min_dod(lpk_people P, int64_t n_slots, int64_t n_pairs) {
    const int lane = threadIdx.x & 31;
    for (int64_t gp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); gp < n_pairs; gp += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        int m = INT_MAX;
        for (int k = lane; k < 256; k += 32) {
            const int64_t i = gp * 256 + k;
            if (i < n_slots && P.disease_state[i] >= 0) m = min(m, P.date_of_death[i]);
        }
        m = __reduce_min_sync(LPK_FULL, m);
        if (lane == 0) P.pair_min_dod[gp] = m;
    }
}
// largest stored ri_timer among the alive, not chronically missed agents of every pair
__global__ void __launch_bounds__(256) k_pair_ri_max(lpk_people P, int64_t n_slots, int64_t n_pairs, int t_next, int ri_step) {
    const int lane = threadIdx.x & 31;
    for (int64_t gp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); gp < n_pairs; gp += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        int m = INT_MIN;
        for (int k = lane; k < 256; k += 32) {
            const int64_t i = gp * 256 + k;
            uint8_t rk = 0;
            if (i < n_slots && P.disease_state[i] >= 0 && P.chronically_missed[i] != 1) {
                m = max(m, (int)P.ri_timer[i]);
                rk = ri_tick_index(P.ri_timer[i], 0, ri_step, t_next - 1);
            }
            P.ri_k[i] = rk;
        }
        m = __reduce_max_sync(LPK_FULL, m);
        if (lane == 0) P.pair_ri_max[gp] = m;
    }
}
__global__ void __launch_bounds__(256) k_hot_settle(lpk_people P, int64_t n_slots, int t_next, int ri_k, int ri_step) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (int64_t)gridDim.x * blockDim.x)
        hot_settle_agent(P, i, t_next, ri_k, ri_step);
}
static inline int64_t hot_padded(int64_t capacity) { return (capacity + LPK_UNIT_AGENTS - 1) / LPK_UNIT_AGENTS * LPK_UNIT_AGENTS; }

extern "C" int lpk_hot_build(const lpk_people *people, int64_t n_slots, int32_t tick_next, int32_t ri_step, int32_t *status, void *stream) {
    REQUIRE(people, "hot_build null struct");
    const lpk_people &P = *people;
    REQUIRE(P.hot && P.rec && P.disease_state && P.strain && P.exposure_timer && P.infection_timer && P.paralysis_timer &&
                P.potentially_paralyzed && P.paralyzed && P.ipv_protected && P.acq_risk_multiplier, "hot_build columns");
    REQUIRE(n_slots >= 0 && n_slots <= P.capacity, "hot_build n_slots");
    cudaStream_t st = as_stream(stream);
    const int64_t padded = hot_padded(P.capacity);
    k_hot_build<<<lpk_sm_count() * 8, 256, 0, st>>>(P, n_slots, padded, tick_next, status);
    CUDA_TRY(cudaGetLastError(), "lpk_hot_build");
    if (P.pair_min_dod && P.date_of_death) {
        k_pair_min_dod<<<lpk_sm_count() * 8, 256, 0, st>>>(P, n_slots, padded / 256);
        CUDA_TRY(cudaGetLastError(), "lpk_hot_build pair_min_dod");
    }
    if (P.pair_ri_max && P.ri_timer) {
        REQUIRE(P.ri_k && P.chronically_missed && ri_step > 0 && tick_next >= 1, "hot_build RI companions");
        k_pair_ri_max<<<lpk_sm_count() * 8, 256, 0, st>>>(P, n_slots, padded / 256, tick_next, ri_step);
        CUDA_TRY(cudaGetLastError(), "lpk_hot_build pair_ri_max");
    }
    return LPK_OK;
}
extern "C" int lpk_hot_settle(const lpk_people *people, int64_t n_slots, int32_t tick_next, int32_t ri_lazy_k, int32_t ri_step, void *stream) {
    REQUIRE(people, "hot_settle null struct");
    const lpk_people &P = *people;
    REQUIRE(P.rec && P.disease_state && P.strain && P.exposure_timer && P.infection_timer && P.paralysis_timer &&
                P.potentially_paralyzed && P.paralyzed && P.ipv_protected, "hot_settle columns");
    REQUIRE(n_slots >= 0 && n_slots <= P.capacity, "hot_settle n_slots");
    if (n_slots == 0) return LPK_OK;
    REQUIRE(ri_lazy_k == 0 || (P.ri_timer && P.chronically_missed && ri_step > 0), "hot_settle RI countdown");
    k_hot_settle<<<lpk_sm_count() * 8, 256, 0, as_stream(stream)>>>(P, n_slots, tick_next, ri_lazy_k, ri_step);
    CUDA_TRY(cudaGetLastError(), "lpk_hot_settle");
    return LPK_OK;
}
extern "C" int64_t lpk_hot_padded(int64_t capacity) { return hot_padded(capacity); }
extern "C" int32_t lpk_hot_risk_e0(float max_risk) { return hot_risk_e0(max_risk); }

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
// one warp per node: the row sum of the network (coalesced) by all lanes, the node's bookkeeping by lane 0.  Launched on
// the first tick of a network only (row sums); afterwards the bookkeeping runs inside k_tx_node_math (lpk_node.cuh).
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
    epilogue_node(a, n, n_lo);
}

int lpk_launch_node_math(int32_t num_nodes, int32_t n_strains, const int64_t *beta_fx, const int64_t *exposure_fx,
                         const int32_t *risk_hist, const double *network, double beta_seasonality, const double *r0_scalars,
                         const int32_t *alive_counts, double zero_inflation, double dispersion, float *tau, double *strain_cdf,
                         double *prob, double *expected, double *ws, uint64_t seed, uint32_t tick, cudaStream_t st, bool rowsums_done,
                         int32_t node_lo, int32_t node_hi, const uint32_t *xchg_flags, int32_t xchg_world, uint32_t xchg_seq,
                         const lpk_node_args *inline_epilogue, double *matvec_ws);

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
    const bool inline_ep = (a.flags & LPK_F_ROWSUMS) != 0;  // nothing but bookkeeping left: k_tx_node_math does it itself
    if (!inline_ep) {
        k_tick_epilogue<<<(owned + 7) / 8, 256, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError(), "lpk_tick_node epilogue");
    }
    return lpk_launch_node_math(a.n_nodes, a.n_strains, a.beta_fx, a.exposure_fx, a.risk_hist, a.network, a.beta_seasonality,
                                a.r0_scalars, a.pop ? a.pop : a.pop_prev, a.zero_inflation, a.dispersion, a.q, a.strain_cdf, a.prob,
                                a.expected, a.rowsum_ws, a.seed, (uint32_t)a.tick, st, true, a.node_hi > 0 ? a.node_lo : 0,
                                a.node_hi > 0 ? a.node_hi : a.n_nodes, a.xchg_flags, a.xchg_world, a.xchg_seq, inline_ep ? &a : nullptr,
                                a.matvec_ws);
}
