// lpk_stages.cuh -- per-agent stage logic shared by the component kernels (lpk_kernels.cu) and the fused
// tick pass (lpk_tick.cu), so both paths execute the very same state machine.
#pragma once
#include "lpk_common.cuh"

struct DevRng {
    uint64_t seed;
    uint32_t tick;
    const double *u1;
    const double *u2;
    const uint32_t *x;
    uint64_t id_base;
};
static inline DevRng dev_rng(const lpk_rng *r) {
    DevRng d;
    d.seed = r ? r->seed : 0; d.tick = r ? r->tick : 0;
    d.u1 = r ? r->u1 : nullptr; d.u2 = r ? r->u2 : nullptr; d.x = r ? r->x : nullptr;
    d.id_base = r ? r->id_base : 0;
    return d;
}


// Paralysis part of the infected block (reference model.py:432-452), paralytic strain 0 only: the gate fires once per
// agent (when the paralysis timer has run out and potentially_paralyzed is still -1) with one uniform; then the
// paralysis timer counts down.  flags bit 0 = newly potentially paralysed, bit 1 = newly paralysed.
__device__ __forceinline__ void paralysis_gate(int64_t i, int8_t ipvv, int8_t &pq, int8_t &par, double p_paralysis, const DevRng &rng,
                                               int &flags) {
    if (ipvv == 0) {
        pq = 1;
        flags |= 1;
        double u;
        if (rng.u1) u = rng.u1[i];
        else { uint32_t x[4]; philox_agent(rng.seed, (uint64_t)i + rng.id_base, rng.tick, LPK_STAGE_PARALYSIS, x); u = u53(x[0], x[1]); }
        if (u < p_paralysis) { par = 1; flags |= 2; }
    } else {
        pq = 0;
    }
}
__device__ __forceinline__ void paralysis_step(int64_t i, int8_t ipvv, int8_t &pt, int8_t &pq, int8_t &par, double p_paralysis,
                                               const DevRng &rng, int &flags) {
    if (pt <= 0 && pq == -1) paralysis_gate(i, ipvv, pq, par, p_paralysis, rng, flags);
    pt = (int8_t)(pt - 1);
}

// Infected block of the state machine on register copies (reference model.py:425-452): recovery timer, then the
// paralysis part for strain 0.  Returns the new state (2 or 3).
__device__ __forceinline__ int8_t ds_infected(int64_t i, int8_t st, int8_t ipvv, int8_t &it, int8_t &pt, int8_t &pq, int8_t &par,
                                              double p_paralysis, const DevRng &rng, int &flags) {
    int8_t s = 2;
    flags = 0;
    if (it <= 0) s = 3;
    it = (int8_t)(it - 1);
    if (st == 0) paralysis_step(i, ipvv, pt, pq, par, p_paralysis, rng, flags);
    return s;
}

// One agent of the state machine (reference model.py:419-452) on the column arrays; returns the new state.
__device__ __forceinline__ int8_t ds_agent(int64_t i, int8_t s, const int16_t *node_id, const int8_t *strain,
                                           int8_t *etimer, int8_t *itimer, int8_t *pot_par, int8_t *paralyzed,
                                           const int8_t *ipv, int8_t *ptimer, double p_paralysis, int32_t *new_pot,
                                           int32_t *new_par, const DevRng &rng) {
    if (s == 1) {
        const int8_t e = etimer[i];
        if (e <= 0) s = 2;
        etimer[i] = (int8_t)(e - 1);
    }
    if (s == 2) {
        const int8_t st = strain[i];
        int8_t it = itimer[i], pt = 0, pq = 0, par = 0, ipvv = 0;
        if (st == 0) { pt = ptimer[i]; pq = pot_par[i]; ipvv = ipv[i]; }
        const int8_t pq0 = pq;
        int flags;
        s = ds_infected(i, st, ipvv, it, pt, pq, par, p_paralysis, rng, flags);
        itimer[i] = it;
        if (st == 0) {
            ptimer[i] = pt;
            if (pq != pq0) pot_par[i] = pq;
            if (flags) {
                const int nd = node_id[i];
                atomicAdd(&new_pot[nd], 1);
                if (flags & 2) { paralyzed[i] = 1; atomicAdd(&new_par[nd], 1); }
            }
        }
    }
    return s;
}


__device__ __forceinline__ bool expose_hit(float p, uint32_t x) {
    if (!(p > 0.f)) return false;
    if (p >= 1.f) return true;
    return x < (uint32_t)__float2uint_rz(p * 4294967296.0f);
}

