/*
 * lp_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement (plain C + OpenMP) of the per-tick agent update of
 * laser-polio v0.2.31, used as the parity checker for the CUDA kernels and as
 * the "port" CPU baseline in bench.py.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference checkout, model.py = src/laser_polio/model.py).
 *
 * Pinning: tests/golden/make_golden.py runs the reference's own numba kernels
 * (AST-loaded read-only from the reference checkout, with the RNG call sites
 * replaced by injected per-agent uniform arrays) and stores seeded inputs and
 * outputs under tests/golden/; tests/test_oracle_golden.py checks this file
 * against them bit-for-bit (integers) / to 2e-5 (the reference's own float32
 * tallies).  tx_infect_ref (weighted sampling without replacement) consumes
 * data-dependent amounts of randomness and is pinned distributionally.
 *
 * Uniform sources.  Each RNG-consuming stage takes `u_inj` pointers; when they
 * are non-NULL the per-agent uniforms are read from them (gate 1: "identical
 * injected uniform draws"), otherwise they come from Philox4x32-10 keyed on
 * (seed, agent, tick, stage) exactly as the CUDA kernels do, so the oracle is
 * bit-comparable with the device path at any size.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#else
static int omp_get_max_threads(void) { return 1; }
static int omp_get_thread_num(void) { return 0; }
#endif

/* ------------------------------------------------------------------ Philox */
/* Philox4x32-10 (Salmon et al., SC'11), independent restatement; checked against
 * the Random123 known-answer vectors in tests/test_philox_kat.py. */
#define ORC_M0 0xD2511F53u
#define ORC_M1 0xCD9E8D57u
#define ORC_W0 0x9E3779B9u
#define ORC_W1 0xBB67AE85u

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)ORC_M0 * c0;
        uint64_t p1 = (uint64_t)ORC_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += ORC_W0; k1 += ORC_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum {
    ORC_STAGE_PARALYSIS = 0,
    ORC_STAGE_RI = 1,
    ORC_STAGE_SIA = 2,
    ORC_STAGE_EXPOSE = 3,
    ORC_STAGE_STRAIN = 4,
    ORC_STAGE_NODE = 5,
    ORC_STAGE_BIRTH = 6,
    ORC_STAGE_LIFESPAN = 7,
    ORC_STAGE_EXPOSE_LO = 8
};

static inline void agent_block(uint64_t seed, uint64_t idx, uint32_t tick, uint32_t stage, uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), tick, stage};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(ctr, key, out);
}

static inline double u53(uint32_t hi, uint32_t lo) {
    uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}

/* exported so tests can compare the device-side uniform derivation */
double orc_uniform53(uint64_t seed, uint64_t idx, uint32_t tick, uint32_t stage, int pair) {
    uint32_t x[4];
    agent_block(seed, idx, tick, stage, x);
    return pair ? u53(x[2], x[3]) : u53(x[0], x[1]);
}

#define ORC_FX_SCALE 1073741824.0 /* 2^30 fixed-point scale of the float tallies */
#define ORC_RISK_BINS 192         /* risk histogram: 8 bins per octave over [2^-12, 2^12) */

/* bin of a susceptible's acq_risk_multiplier: (exponent, top three mantissa bits) relative to 2^-12, clamped */
static inline int risk_bin(float w) {
    if (!(w > 0.f)) return 0;
    uint32_t bits;
    memcpy(&bits, &w, 4);
    int b = (int)(bits >> 20) - ((127 - 12) << 3);
    return b < 0 ? 0 : (b >= ORC_RISK_BINS ? ORC_RISK_BINS - 1 : b);
}

/* 1 - exp(-x) for x >= 0 through fmaf / floorf / exact power-of-two scaling only (bit-identical on the device) */
static inline float p_expose(float x) {
    if (!(x > 0.f)) return 0.f;
    if (x < 0.0625f) {
        float t = fmaf(-x, 0.008333333767950535f, 0.0416666679084301f);
        t = fmaf(-x, t, 0.1666666716337204f);
        t = fmaf(-x, t, 0.5f);
        t = fmaf(-x, t, 1.0f);
        return x * t;
    }
    if (x >= 17.f) return 1.f;
    const float y = x * 1.4426950216293335f;
    const float n = floorf(y);
    const float f = y - n;
    float r = 0.00010938752529909834f;
    r = fmaf(r, f, -0.0012757162330672145f);
    r = fmaf(r, f, 0.009580058045685291f);
    r = fmaf(r, f, -0.05549103394150734f);
    r = fmaf(r, f, 0.24022436141967773f);
    r = fmaf(r, f, -0.6931470632553101f);
    r = fmaf(r, f, 1.0f);
    uint32_t sb = (uint32_t)(127 - (int)n) << 23;
    float scale;
    memcpy(&scale, &sb, 4);
    return 1.f - r * scale;
}

float orc_p_expose(float x) { return p_expose(x); }
int orc_risk_bin(float w) { return risk_bin(w); }

static int32_t *tl_alloc(int n_threads, int64_t n) {
    return (int32_t *)calloc((size_t)n_threads * (size_t)n, sizeof(int32_t));
}
static void tl_reduce_add(const int32_t *tl, int n_threads, int64_t n, int32_t *out, int accumulate) {
    for (int64_t j = 0; j < n; ++j) {
        int64_t s = 0;
        for (int t = 0; t < n_threads; ++t) s += tl[(int64_t)t * n + j];
        out[j] = accumulate ? out[j] + (int32_t)s : (int32_t)s;
    }
}

/* ---------------------------------------------------------------- V1 deaths */
/* model.py:1767-1781 get_deaths: alive and date_of_death <= t -> state = -1, count per node. */
void orc_get_deaths(int32_t n_nodes, int64_t n_people, int8_t *state, const int16_t *node_id,
                    const int32_t *dod, int32_t t, int32_t *num_dying) {
    int nt = omp_get_max_threads();
    int32_t *tl = tl_alloc(nt, n_nodes);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_people; ++i) {
        if (state[i] >= 0 && dod[i] <= t) {
            state[i] = -1;
            tl[(int64_t)omp_get_thread_num() * n_nodes + node_id[i]] += 1;
        }
    }
    tl_reduce_add(tl, nt, n_nodes, num_dying, 0);
    free(tl);
}

/* ------------------------------------------------------- D1 disease state */
/* model.py:344-454 disease_state_step(+_kernel).  E block then I block (an E that
 * converts this tick falls through into the I block, model.py:419-425); paralysis
 * gate only for strain 0, only once (potentially_paralyzed leaves -1), one uniform
 * per gated agent compared with p_paralysis rounded to float32 (model.py:734, 441).
 * Outputs are ADDED to new_potential/new_paralyzed (model.py:389-390). */
void orc_disease_state_step(const int16_t *node_id, int32_t n_nodes, int8_t *state, const int8_t *strain,
                            int64_t count, int8_t *etimer, int8_t *itimer, int8_t *pot_par, int8_t *paralyzed,
                            const int8_t *ipv, int8_t *ptimer, float p_paralysis, int32_t *new_potential,
                            int32_t *new_paralyzed, const double *u_inj, uint64_t seed, uint32_t tick, uint64_t id_base) {
    int nt = omp_get_max_threads();
    int32_t *tl_pot = tl_alloc(nt, n_nodes), *tl_par = tl_alloc(nt, n_nodes);
    const double p = (double)p_paralysis;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < count; ++i) {
        int tid = omp_get_thread_num();
        if (state[i] == 1) {
            if (etimer[i] <= 0) state[i] = 2;
            etimer[i] = (int8_t)(etimer[i] - 1);
        }
        if (state[i] == 2) {
            if (itimer[i] <= 0) state[i] = 3;
            itimer[i] = (int8_t)(itimer[i] - 1);
            if (strain[i] == 0) {
                if (ptimer[i] <= 0 && pot_par[i] == -1) {
                    if (ipv[i] == 0) {
                        pot_par[i] = 1;
                        tl_pot[(int64_t)tid * n_nodes + node_id[i]] += 1;
                        double u;
                        if (u_inj) u = u_inj[i];
                        else { uint32_t x[4]; agent_block(seed, (uint64_t)i + id_base, tick, ORC_STAGE_PARALYSIS, x); u = u53(x[0], x[1]); }
                        if (u < p) {
                            paralyzed[i] = 1;
                            tl_par[(int64_t)tid * n_nodes + node_id[i]] += 1;
                        }
                    } else {
                        pot_par[i] = 0;
                    }
                }
                ptimer[i] = (int8_t)(ptimer[i] - 1);
            }
        }
    }
    tl_reduce_add(tl_pot, nt, n_nodes, new_potential, 1);
    tl_reduce_add(tl_par, nt, n_nodes, new_paralyzed, 1);
    free(tl_pot); free(tl_par);
}

/* ---------------------------------------------------------------- R1 RI */
/* model.py:1805-1855 fast_ri.  Alive, not chronically missed: timer -= step stored;
 * eligible window closed on the low side only on the first RI tick (1839-1842);
 * two uniforms per eligible agent, OPV first then IPV (1845, 1852). */
void orc_fast_ri(int64_t step_size, const int16_t *node_id, int8_t *state, int8_t *strain, int8_t *ipv,
                 int16_t *ri_timer, int64_t sim_t, const double *prob_ri, const double *prob_ipv,
                 int64_t num_people, int32_t n_nodes, int32_t *ri_counts, int32_t *ri_protected,
                 int32_t *ipv_counts, const uint8_t *missed, int8_t vaccine_strain, const double *u1_inj,
                 const double *u2_inj, uint64_t seed, uint32_t tick, uint64_t id_base) {
    int nt = omp_get_max_threads();
    int32_t *tl_ri = tl_alloc(nt, n_nodes), *tl_pr = tl_alloc(nt, n_nodes), *tl_ipv = tl_alloc(nt, n_nodes);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < num_people; ++i) {
        int8_t s = state[i];
        if (s < 0) continue;
        if (missed[i] == 1) continue;
        int32_t node = node_id[i];
        int64_t timer = (int64_t)ri_timer[i] - step_size; /* eligibility on the unwrapped value (numba int64) */
        ri_timer[i] = (int16_t)timer;
        int eligible = 0;
        if (sim_t == step_size) eligible = (timer <= 0 && timer >= -step_size);
        else if (sim_t > step_size) eligible = (timer <= 0 && timer > -step_size);
        if (!eligible) continue;
        double u1, u2;
        if (u1_inj) { u1 = u1_inj[i]; u2 = u2_inj[i]; }
        else { uint32_t x[4]; agent_block(seed, (uint64_t)i + id_base, tick, ORC_STAGE_RI, x); u1 = u53(x[0], x[1]); u2 = u53(x[2], x[3]); }
        int64_t row = (int64_t)omp_get_thread_num() * n_nodes + node;
        if (u1 < prob_ri[node]) {
            tl_ri[row] += 1;
            if (s == 0) { state[i] = 1; strain[i] = vaccine_strain; tl_pr[row] += 1; }
        }
        if (u2 < prob_ipv[node]) { tl_ipv[row] += 1; ipv[i] = 1; }
    }
    tl_reduce_add(tl_ri, nt, n_nodes, ri_counts, 0);
    tl_reduce_add(tl_pr, nt, n_nodes, ri_protected, 0);
    tl_reduce_add(tl_ipv, nt, n_nodes, ipv_counts, 0);
    free(tl_ri); free(tl_pr); free(tl_ipv);
}

/* ---------------------------------------------------------------- S1 SIA */
/* model.py:1995-2060 fast_sia.  One uniform r reused for both thresholds
 * (2049-2056); vx_prob is float32 (2109) and the take threshold is the product
 * prob_vx * vx_eff evaluated in double as numba does for f32*f64. */
void orc_fast_sia(const int16_t *node_id, int8_t *state, int8_t *strain, const int32_t *dob, int64_t sim_t,
                  const float *vx_prob, double vx_eff, int64_t count, const uint8_t *nodes_to_vaccinate,
                  int64_t min_age, int64_t max_age, int32_t n_nodes, int32_t *vaccinated, int32_t *protected_,
                  const uint8_t *missed, int8_t vaccine_strain, const double *u_inj, uint64_t seed,
                  uint32_t tick, uint32_t event_idx, uint64_t id_base) {
    int nt = omp_get_max_threads();
    int32_t *tl_v = tl_alloc(nt, n_nodes), *tl_p = tl_alloc(nt, n_nodes);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < count; ++i) {
        if (state[i] < 0) continue;
        if (missed[i] == 1) continue;
        int64_t age = sim_t - dob[i];
        if (!(min_age <= age && age <= max_age)) continue;
        int32_t node = node_id[i];
        if (nodes_to_vaccinate[node] == 0) continue;
        double r;
        if (u_inj) r = u_inj[i];
        else { uint32_t x[4]; agent_block(seed, (uint64_t)i + id_base, tick, ORC_STAGE_SIA | (event_idx << 8), x); r = u53(x[0], x[1]); }
        double pv = (double)vx_prob[node];
        if (r < pv) {
            int64_t row = (int64_t)omp_get_thread_num() * n_nodes + node;
            tl_v[row] += 1;
            if (state[i] == 0 && r < pv * vx_eff) { state[i] = 1; strain[i] = vaccine_strain; tl_p[row] += 1; }
        }
    }
    tl_reduce_add(tl_v, nt, n_nodes, vaccinated, 0);
    tl_reduce_add(tl_p, nt, n_nodes, protected_, 0);
    free(tl_v); free(tl_p);
}

/* ------------------------------------------------------------- T1 tallies */
/* model.py:932-1007 tx_step_prep(+_kernel).  Three accumulation flavours:
 *   mode 0: float32 thread-local accumulators like the reference (order dependent);
 *   mode 1: float64 accumulators (the "truth" gate 2 is quoted against);
 *   mode 2: exact 2^30 fixed-point int64 (what the device does; order independent).
 * beta_out/exposure_out are doubles in all modes; *_fx are filled in mode 2 only. */
void orc_tx_step_prep(int32_t n_nodes, int64_t n_people, int32_t n_strains, const int8_t *strain,
                      const double *strain_r0_scalars, const int8_t *state, const int16_t *node_id,
                      const float *infectivity, const float *risk, int mode, double *beta_out,
                      double *exposure_out, int64_t *sus_out, int64_t *beta_fx, int64_t *exposure_fx, int32_t *risk_hist) {
    int nt = omp_get_max_threads();
    int64_t nb = (int64_t)n_nodes * n_strains;
    float *f_beta = NULL, *f_exp = NULL;
    double *d_beta = NULL, *d_exp = NULL;
    int64_t *x_beta = NULL, *x_exp = NULL;
    if (mode == 0) { f_beta = calloc((size_t)nt * nb, sizeof(float)); f_exp = calloc((size_t)nt * n_nodes, sizeof(float)); }
    if (mode == 1) { d_beta = calloc((size_t)nt * nb, sizeof(double)); d_exp = calloc((size_t)nt * n_nodes, sizeof(double)); }
    if (mode == 2) { x_beta = calloc((size_t)nt * nb, sizeof(int64_t)); x_exp = calloc((size_t)nt * n_nodes, sizeof(int64_t)); }
    int32_t *tl_sus = tl_alloc(nt, n_nodes);
    int32_t *tl_hist = risk_hist ? tl_alloc(nt, (int64_t)n_nodes * ORC_RISK_BINS) : NULL;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_people; ++i) {
        int tid = omp_get_thread_num();
        int8_t s = state[i];
        int32_t nid = node_id[i];
        if (s == 0) {
            tl_sus[(int64_t)tid * n_nodes + nid] += 1;
            if (tl_hist) tl_hist[((int64_t)tid * n_nodes + nid) * ORC_RISK_BINS + risk_bin(risk[i])] += 1;
            if (mode == 0) f_exp[(int64_t)tid * n_nodes + nid] += risk[i];
            else if (mode == 1) d_exp[(int64_t)tid * n_nodes + nid] += (double)risk[i];
            else x_exp[(int64_t)tid * n_nodes + nid] += llrint((double)risk[i] * ORC_FX_SCALE);
        } else if (s == 2) {
            int32_t st = strain[i];
            double v = (double)infectivity[i] * strain_r0_scalars[st];
            int64_t k = (int64_t)tid * nb + (int64_t)nid * n_strains + st;
            if (mode == 0) f_beta[k] = (float)((double)f_beta[k] + v);
            else if (mode == 1) d_beta[k] += v;
            else x_beta[k] += llrint(v * ORC_FX_SCALE);
        }
    }
    for (int64_t j = 0; j < nb; ++j) {
        if (mode == 0) { float s = 0.f; for (int t = 0; t < nt; ++t) s += f_beta[(int64_t)t * nb + j]; beta_out[j] = s; }
        else if (mode == 1) { double s = 0; for (int t = 0; t < nt; ++t) s += d_beta[(int64_t)t * nb + j]; beta_out[j] = s; }
        else { int64_t s = 0; for (int t = 0; t < nt; ++t) s += x_beta[(int64_t)t * nb + j]; beta_fx[j] = s; beta_out[j] = (double)s / ORC_FX_SCALE; }
    }
    for (int64_t j = 0; j < n_nodes; ++j) {
        if (mode == 0) { float s = 0.f; for (int t = 0; t < nt; ++t) s += f_exp[(int64_t)t * n_nodes + j]; exposure_out[j] = s; }
        else if (mode == 1) { double s = 0; for (int t = 0; t < nt; ++t) s += d_exp[(int64_t)t * n_nodes + j]; exposure_out[j] = s; }
        else { int64_t s = 0; for (int t = 0; t < nt; ++t) s += x_exp[(int64_t)t * n_nodes + j]; exposure_fx[j] = s; exposure_out[j] = (double)s / ORC_FX_SCALE; }
        int64_t c = 0;
        for (int t = 0; t < nt; ++t) c += tl_sus[(int64_t)t * n_nodes + j];
        sus_out[j] = c;
    }
    if (tl_hist) { tl_reduce_add(tl_hist, nt, (int64_t)n_nodes * ORC_RISK_BINS, risk_hist, 0); free(tl_hist); }
    free(f_beta); free(f_exp); free(d_beta); free(d_exp); free(x_beta); free(x_exp); free(tl_sus);
}

/* ------------------------------------------------------------- C1 census */
/* model.py:869-929 count_SEIRP(+_kernel): alive agents only; S,R per node; E,I per
 * node x strain; potentially_paralyzed==1 and paralyzed==1 per node. */
void orc_count_seirp(const int16_t *node_id, const int8_t *state, const int8_t *strain, const int8_t *pot_par,
                     const int8_t *paralyzed, int32_t n_nodes, int32_t n_strains, int64_t n_people, int32_t *S,
                     int32_t *E, int32_t *I, int32_t *R, int32_t *Ebs, int32_t *Ibs, int32_t *POTP, int32_t *P) {
    int nt = omp_get_max_threads();
    int64_t nb = (int64_t)n_nodes * n_strains;
    int32_t *tS = tl_alloc(nt, n_nodes), *tR = tl_alloc(nt, n_nodes), *tPP = tl_alloc(nt, n_nodes),
            *tP = tl_alloc(nt, n_nodes), *tE = tl_alloc(nt, nb), *tI = tl_alloc(nt, nb);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_people; ++i) {
        int8_t ds = state[i];
        if (ds < 0) continue;
        int tid = omp_get_thread_num();
        int32_t nd = node_id[i];
        int32_t st = strain[i];
        if (ds == 0) tS[(int64_t)tid * n_nodes + nd] += 1;
        else if (ds == 1) tE[(int64_t)tid * nb + (int64_t)nd * n_strains + st] += 1;
        else if (ds == 2) tI[(int64_t)tid * nb + (int64_t)nd * n_strains + st] += 1;
        else if (ds == 3) tR[(int64_t)tid * n_nodes + nd] += 1;
        if (pot_par[i] == 1) tPP[(int64_t)tid * n_nodes + nd] += 1;
        if (paralyzed[i] == 1) tP[(int64_t)tid * n_nodes + nd] += 1;
    }
    tl_reduce_add(tS, nt, n_nodes, S, 0);
    tl_reduce_add(tR, nt, n_nodes, R, 0);
    tl_reduce_add(tPP, nt, n_nodes, POTP, 0);
    tl_reduce_add(tP, nt, n_nodes, P, 0);
    tl_reduce_add(tE, nt, nb, Ebs, 0);
    tl_reduce_add(tI, nt, nb, Ibs, 0);
    for (int32_t n = 0; n < n_nodes; ++n) {
        int32_t e = 0, ii = 0;
        for (int32_t s = 0; s < n_strains; ++s) { e += Ebs[(int64_t)n * n_strains + s]; ii += Ibs[(int64_t)n * n_strains + s]; }
        E[n] = e; I[n] = ii;
    }
    free(tS); free(tR); free(tPP); free(tP); free(tE); free(tI);
}

/* ----------------------------------------- T3 (reference sampling scheme) */
/* model.py:1010-1149 tx_infect_nb: serial bucket pass of susceptibles of nodes with
 * requests (1050-1061), then per node successive weighted sampling without
 * replacement by cumsum / searchsorted(right) / unique with retries (1096-1122) and a
 * categorical strain draw per pick (1127-1147).  The reference draws from numba's
 * per-thread Mersenne streams in data-dependent amounts, so this restatement is
 * distributional: it uses xoshiro256** seeded per (seed, node, tick). */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
typedef struct { uint64_t s[4]; } xo256;
static inline uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline void xo_seed(xo256 *g, uint64_t a) { for (int i = 0; i < 4; ++i) g->s[i] = splitmix64(&a); }
static inline uint64_t xo_next(xo256 *g) {
    uint64_t *s = g->s, r = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return r;
}
static inline double xo_u(xo256 *g) { return (double)(xo_next(g) >> 11) * (1.0 / 9007199254740992.0); }

static int cmp_i32(const void *a, const void *b) { int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return (x > y) - (x < y); }

void orc_tx_infect_ref(int32_t n_nodes, int64_t n_people, int32_t n_strains, const int64_t *sus_by_node,
                       const int16_t *node_id, int8_t *strain, int8_t *state, int32_t *sus_indices,
                       float *sus_probs, const float *risk, const double *prob_exp /*[nodes,strains]*/,
                       const int32_t *n_to_create /*[nodes,strains]*/, int32_t *n_new /*[nodes,strains]*/,
                       uint64_t seed, uint32_t tick) {
    int64_t *offsets = calloc((size_t)n_nodes, sizeof(int64_t));
    int64_t *next = calloc((size_t)n_nodes, sizeof(int64_t));
    int64_t *total = calloc((size_t)n_nodes, sizeof(int64_t));
    for (int32_t n = 1; n < n_nodes; ++n) offsets[n] = offsets[n - 1] + sus_by_node[n - 1];
    for (int32_t n = 0; n < n_nodes; ++n) {
        next[n] = offsets[n];
        for (int32_t s = 0; s < n_strains; ++s) total[n] += n_to_create[(int64_t)n * n_strains + s];
    }
    for (int64_t i = 0; i < n_people; ++i) { /* serial in the reference too */
        int32_t nid = node_id[i];
        if (total[nid] > 0 && state[i] == 0) {
            int64_t idx = next[nid];
            sus_indices[idx] = (int32_t)i;
            sus_probs[idx] = risk[i];
            next[nid] = idx + 1;
        }
    }
    memset(n_new, 0, sizeof(int32_t) * (size_t)n_nodes * n_strains);
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t node = 0; node < n_nodes; ++node) {
        int64_t need = total[node];
        int64_t sus_count = sus_by_node[node];
        if (need <= 0 || sus_count == 0) continue;
        const double *pe = prob_exp + (int64_t)node * n_strains;
        double total_foi = 0;
        for (int32_t s = 0; s < n_strains; ++s) total_foi += pe[s];
        if (total_foi <= 0) continue;
        const int32_t *idx = sus_indices + offsets[node];
        const float *base = sus_probs + offsets[node];
        int64_t size = need < sus_count ? need : sus_count;
        float *p = malloc(sizeof(float) * (size_t)sus_count);
        float *cdf = malloc(sizeof(float) * (size_t)sus_count);
        int32_t *sel = malloc(sizeof(int32_t) * (size_t)size);
        int32_t *probe = malloc(sizeof(int32_t) * (size_t)size);
        for (int64_t k = 0; k < sus_count; ++k) p[k] = (float)((double)base[k] * total_foi);
        xo256 g;
        xo_seed(&g, seed ^ ((uint64_t)node << 32) ^ ((uint64_t)tick * 0x9E3779B97F4A7C15ull));
        int64_t n_uniq = 0;
        while (n_uniq < size) {
            int64_t m = size - n_uniq;
            float acc = 0.f;
            for (int64_t k = 0; k < sus_count; ++k) { acc += p[k]; cdf[k] = acc; }
            if (cdf[sus_count - 1] <= 0) break;
            double top = (double)cdf[sus_count - 1];
            for (int64_t j = 0; j < m; ++j) {
                double x = xo_u(&g) * top;
                int64_t lo = 0, hi = sus_count; /* searchsorted(side="right") */
                while (lo < hi) { int64_t mid = (lo + hi) >> 1; if ((double)cdf[mid] <= x) lo = mid + 1; else hi = mid; }
                probe[j] = (int32_t)lo;
            }
            qsort(probe, (size_t)m, sizeof(int32_t), cmp_i32);
            int64_t nu = 0;
            for (int64_t j = 0; j < m; ++j) if (j == 0 || probe[j] != probe[j - 1]) probe[nu++] = probe[j];
            int64_t take = nu < m ? nu : m;
            for (int64_t j = 0; j < take; ++j) sel[n_uniq + j] = probe[j];
            n_uniq += take;
            if (n_uniq < size) for (int64_t j = 0; j < nu; ++j) p[probe[j]] = 0.f;
        }
        for (int64_t k = 0; k < n_uniq; ++k) {
            int32_t person = idx[sel[k]];
            if (state[person] != 0) continue;
            double r = xo_u(&g), cum = 0;
            int32_t assigned = 0;
            for (int32_t s = 0; s < n_strains; ++s) { cum += pe[s] / total_foi; if (r < cum) { assigned = s; break; } }
            state[person] = 1;
            strain[person] = (int8_t)assigned;
            n_new[(int64_t)node * n_strains + assigned] += 1;
        }
        free(p); free(cdf); free(sel); free(probe);
    }
    free(offsets); free(next); free(total);
}

/* ------------------------------- T3 (device scheme: per-agent Bernoulli) */
/* SURVEY Appendix F option F1 + importation gate, the scheme the north star
 * prescribes for the device: susceptible agent i of node n is exposed iff
 * x_i < thr(risk_i * q[n]) with x_i = expose_word(seed, i, tick) (below) and
 * q[n] = tau[n] (oracle.py: tx_node_math); strain by the cumulative
 * categorical of model.py:1127-1141 on a second block (seed, i, tick, STRAIN).
 * Marginal exposure probability w_i * P_n and node mean exposure[n] * P_n equal the
 * reference's (model.py:1362-1363, 1087). */
/* The 32-bit exposure word of agent id (= index + id_base): X = h16 << 16 | l16, the half-words
 * hw = ((id >> 7) & 1) * 4 + (id & 3) of the two Philox blocks (seed; ctr, tick, EXPOSE / EXPOSE_LO),
 * ctr = (id >> 8) * 32 + ((id >> 2) & 31): one block serves the 8 agents a device lane owns in a pair of
 * 128-agent rows, and the device generates the low block only when the high half cannot decide the trial
 * (include/lpk.h, T3). */
static inline uint32_t expose_word(uint64_t seed, uint64_t id, uint32_t tick) {
    uint64_t ctr = ((id >> 8) << 5) | ((id >> 2) & 31u);
    int hw = (int)(((id >> 7) & 1u) * 4u + (id & 3u));
    uint32_t h[4], l[4];
    agent_block(seed, ctr, tick, ORC_STAGE_EXPOSE, h);
    agent_block(seed, ctr, tick, ORC_STAGE_EXPOSE_LO, l);
    uint32_t h16 = (h[hw >> 1] >> (16 * (hw & 1))) & 0xFFFFu, l16 = (l[hw >> 1] >> (16 * (hw & 1))) & 0xFFFFu;
    return (h16 << 16) | l16;
}

static inline uint32_t expose_threshold(float p, int *always) {
    *always = 0;
    if (!(p > 0.f)) return 0u;
    if (p >= 1.f) { *always = 1; return 0xFFFFFFFFu; }
    return (uint32_t)(p * 4294967296.0f); /* truncation; p < 1 so the product < 2^32 */
}

void orc_tx_infect_bernoulli(int32_t n_nodes, int64_t n_people, int32_t n_strains, const int16_t *node_id,
                             int8_t *strain, int8_t *state, const float *risk, const float *q /*[nodes]*/,
                             const double *strain_cdf /*[nodes,strains] cumulative*/, int32_t *n_new,
                             const uint32_t *x_inj, const double *u_strain_inj, uint64_t seed, uint32_t tick,
                             uint64_t id_base) {
    int nt = omp_get_max_threads();
    int64_t nb = (int64_t)n_nodes * n_strains;
    int32_t *tl = tl_alloc(nt, nb);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_people; ++i) {
        if (state[i] != 0) continue;
        int32_t nid = node_id[i];
        float qn = q[nid];
        if (!(qn > 0.f)) continue;
        float p = p_expose(risk[i] * qn); /* qn = tau[node] */
        int always;
        uint32_t thr = expose_threshold(p, &always);
        uint32_t x;
        if (x_inj) x = x_inj[i];
        else x = expose_word(seed, (uint64_t)i + id_base, tick);
        if (!(always || x < thr)) continue;
        double r;
        if (u_strain_inj) r = u_strain_inj[i];
        else { uint32_t b[4]; agent_block(seed, (uint64_t)i + id_base, tick, ORC_STAGE_STRAIN, b); r = u53(b[0], b[1]); }
        int32_t assigned = 0;
        for (int32_t s = 0; s < n_strains; ++s) if (r < strain_cdf[(int64_t)nid * n_strains + s]) { assigned = s; break; }
        state[i] = 1;
        strain[i] = (int8_t)assigned;
        tl[(int64_t)omp_get_thread_num() * nb + (int64_t)nid * n_strains + assigned] += 1;
    }
    tl_reduce_add(tl, nt, nb, n_new, 0);
    free(tl);
}

int orc_num_threads(void) { return omp_get_max_threads(); }
