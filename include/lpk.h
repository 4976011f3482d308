/*
 * lpk.h -- C ABI of liblpk.so, the B200 (sm_100a) kernels for laser-polio's
 * per-tick agent update.
 *
 * Drop-in boundary: one entry point per free function of the reference's hot
 * path.  Each declaration cites the reference
 * signature it replaces (paths relative to the reference checkout,
 * model.py = src/laser_polio/model.py).  Array arguments keep the reference's
 * order, meaning and dtypes; they are DEVICE pointers into the structure-of-
 * arrays agent table (owned by the caller -- PyTorch buffers in the shipped
 * host code; the library borrows and never frees).  Per-node outputs are device
 * pointers too.  `stream` is a cudaStream_t passed as void*.  No torch types
 * cross this boundary.
 *
 * Every call is asynchronous on `stream` and returns a status:
 *   0                      success (kernel enqueued)
 *   LPK_ERR_ARG  (-1)      null pointer / negative size / unsupported n_strains
 *   LPK_ERR_CUDA (-2)      a CUDA runtime error; text via lpk_last_error()
 *
 * Uniform source (`lpk_rng`).  The reference draws from numba's per-thread
 * Mersenne streams (np.random.random() at model.py:441, np.random.rand() at
 * :1845, :1852, :2049), which no parallel device schedule can reproduce.  Here
 * every draw is Philox4x32-10 keyed on (seed, agent index, tick, stage), so a
 * result depends neither on grid shape nor on GPU count.  For parity runs the
 * optional u1/u2/x arrays inject the same per-agent uniforms the reference
 * (RNG call sites textually replaced) and the oracle consume.
 */
#ifndef LPK_H
#define LPK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPK_OK 0
#define LPK_ERR_ARG (-1)
#define LPK_ERR_CUDA (-2)

#define LPK_MAX_STRAINS 4
#define LPK_FX_SCALE 1073741824.0 /* 2^30: fixed-point scale of the float tallies */
#define LPK_RISK_BINS 192          /* risk histogram: 8 bins per octave over [2^-12, 2^12) */

/* Philox stage ids (counter word 3) */
#define LPK_STAGE_PARALYSIS 0u
#define LPK_STAGE_RI 1u
#define LPK_STAGE_SIA 2u /* | event index << 8 */
#define LPK_STAGE_EXPOSE 3u
#define LPK_STAGE_STRAIN 4u
#define LPK_STAGE_NODE 5u
#define LPK_STAGE_BIRTH 6u
#define LPK_STAGE_LIFESPAN 7u
#define LPK_STAGE_EXPOSE_LO 8u /* low half of the exposure word, generated only when the high half cannot decide */
#define LPK_STAGE_NODE_VAR 9u  /* node-level variance multiplier of the exposure count (lpk_tx_node_math) */

typedef struct lpk_rng {
    uint64_t seed;      /* Philox key */
    uint32_t tick;      /* Philox counter word 2 */
    uint32_t _pad;
    const double *u1;   /* optional injected per-agent uniforms (device), else NULL */
    const double *u2;   /* second injected stream (fast_ri IPV draw; strain pick in tx_infect) */
    const uint32_t *x;  /* optional injected per-agent 32-bit words for the exposure trial */
    uint64_t id_base;   /* added to the agent index in every Philox counter: the global id of local agent 0 when the
                           table is one node-shard of a larger population (multiple of 256); 0 on a single GPU */
} lpk_rng;

const char *lpk_last_error(void);
int lpk_version(void);
/* Philox4x32-10 on the device for `n` (ctr,key) pairs -- known-answer testing only. */
int lpk_philox_selftest(const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4, int64_t n, void *stream);

/* V1  replaces get_deaths(num_nodes, num_people, disease_state, node_id, date_of_death, t, tl_dying, num_dying)
 *     model.py:1767-1781.  num_dying[n_nodes] is OVERWRITTEN (reference: num_dying[:] = tl_dying.sum(axis=0)).
 *     The reference's thread-local scratch (tl_dying) has no device counterpart. */
int lpk_get_deaths(int32_t num_nodes, int64_t num_people, int8_t *disease_state, const int16_t *node_id,
                   const int32_t *date_of_death, int32_t t, int32_t *num_dying, void *stream);

/* D1  replaces disease_state_step(node_id, n_nodes, disease_state, strain, active_count, exposure_timer,
 *     infection_timer, potentially_paralyzed, paralyzed, ipv_protected, paralysis_timer, p_paralysis,
 *     new_potential, new_paralyzed)  model.py:344-454.  new_potential/new_paralyzed[n_nodes] are ADDED to
 *     (reference: new_potential[:] += ..., model.py:389-390). */
int lpk_disease_state_step(const int16_t *node_id, int32_t n_nodes, int8_t *disease_state, const int8_t *strain,
                           int64_t active_count, int8_t *exposure_timer, int8_t *infection_timer,
                           int8_t *potentially_paralyzed, int8_t *paralyzed, const int8_t *ipv_protected,
                           int8_t *paralysis_timer, float p_paralysis, int32_t *new_potential,
                           int32_t *new_paralyzed, const lpk_rng *rng, void *stream);

/* R1  replaces fast_ri(step_size, node_id, disease_state, strain, ipv_protected, ri_timer, sim_t, vx_prob_ri,
 *     vx_prob_ipv, num_people, local_ri_counts, local_ri_protected, local_ipv_counts, chronically_missed,
 *     ri_vaccine_strain)  model.py:1805-1855.  The three [threads, nodes] scratch arrays of the reference
 *     become per-node outputs ri_counts/ri_protected/ipv_counts[n_nodes], OVERWRITTEN
 *     (reference: results.ri_vaccinated[t] = local.sum(axis=0), model.py:1970-1983). */
int lpk_fast_ri(int64_t step_size, const int16_t *node_id, int8_t *disease_state, int8_t *strain,
                int8_t *ipv_protected, int16_t *ri_timer, int64_t sim_t, const double *vx_prob_ri,
                const double *vx_prob_ipv, int64_t num_people, int32_t n_nodes, int32_t *ri_counts,
                int32_t *ri_protected, int32_t *ipv_counts, const uint8_t *chronically_missed,
                int8_t ri_vaccine_strain, const lpk_rng *rng, void *stream);

/* S1  replaces fast_sia(node_ids, disease_states, strain, dobs, sim_t, vx_prob, vx_eff, count, nodes_to_vaccinate,
 *     min_age, max_age, local_vaccinated, local_protected, chronically_missed, sia_vaccine_strain)
 *     model.py:1995-2060.  vaccinated/protected_[n_nodes] OVERWRITTEN.  event_idx distinguishes several
 *     campaigns on one day (each sees the previous one's state changes, model.py:2103). */
int lpk_fast_sia(const int16_t *node_ids, int8_t *disease_states, int8_t *strain, const int32_t *dobs,
                 int64_t sim_t, const float *vx_prob, double vx_eff, int64_t count,
                 const uint8_t *nodes_to_vaccinate, int64_t min_age, int64_t max_age, int32_t n_nodes,
                 int32_t *vaccinated, int32_t *protected_, const uint8_t *chronically_missed,
                 int8_t sia_vaccine_strain, uint32_t event_idx, const lpk_rng *rng, void *stream);

/* T1  replaces tx_step_prep(num_nodes, num_people, n_strains, strains, strain_r0_scalars, disease_states,
 *     node_ids, daily_infectivity, risks)  model.py:932-1007.  Outputs OVERWRITTEN:
 *       beta_fx[num_nodes * n_strains]  sum over infectious of infectivity * strain_r0_scalars[strain]
 *       exposure_fx[num_nodes]          sum over susceptibles of acq_risk_multiplier
 *       sus[num_nodes]                  susceptible count
 *       risk_hist[num_nodes * LPK_RISK_BINS]  histogram of the susceptibles' acq_risk_multiplier (log-spaced bins);
 *                                       it replaces the reference's per-node sus_probs scratch (model.py:1055-1061):
 *                                       lpk_tx_node_math solves the node's exposure scale on it
 *     The float sums are exact 2^30 fixed point in int64 (value = fx / LPK_FX_SCALE): order independent,
 *     bitwise reproducible across launch shapes and GPU counts, within 1e-9 of the float64 sum (the
 *     reference's per-thread float32 accumulation is itself ~5e-6 off it).  strain_r0_scalars: HOST double[n_strains]. */
int lpk_tx_step_prep(int32_t num_nodes, int64_t num_people, int32_t n_strains, const int8_t *strains,
                     const double *h_strain_r0_scalars, const int8_t *disease_states, const int16_t *node_ids,
                     const float *daily_infectivity, const float *risks, int64_t *beta_fx, int64_t *exposure_fx,
                     int64_t *sus, int32_t *risk_hist, void *stream);

/* T2  replaces the node-level block of Transmission_ABM.step, model.py:1332-1351 and 1362-1407:
 *     network transfer beta += W^T beta - beta * rowsum(W), x seasonality x r0_scalars, / max(pop, 1),
 *     p = max(1 - exp(-rate), 0), expected[n] = exposure[n] * sum_s p[n,s] (model.py:1363).
 *     The reference then draws an integer count per node on the host (Poisson; zero-inflated NB for nodes without
 *     local infectivity) and tx_infect_nb picks that many susceptibles by successive weighted sampling, which selects
 *     agent i with probability 1 - exp(-w_i tau), tau fixed by the count.  Here the node's scale is solved for the
 *     EXPECTED count instead:   sum_{i in S_n} (1 - exp(-w_i tau[n])) = T_n   (on risk_hist, bin centres rescaled so that
 *     their sum is the exact risk sum exposure[n]), with
 *       E1 = expected[n] * g_n,  g_n = 1 with local infectivity, else 0 w.p. zero_inflation, else
 *            Gamma(r, 1/r) / (1 - zero_inflation), r = max(1, round(dispersion)) (the ZINB as a zero-inflated
 *            gamma-Poisson mixture; draws Philox(seed; node, k, tick, NODE));
 *       E2 = E1 * g2, g2 a unit-mean gamma that supplies the variance independent trials lack against the reference's
 *            count min(K, S), K ~ Poisson(E1): CV^2 = (Var[min(K, S)] - T (1 - T / S_eff)) / (P(K < S) E1)^2,
 *            S_eff = (sum w)^2 / sum w^2 (draws Philox(seed; node, k, tick, NODE_VAR); 1 / S_eff when E1 << S);
 *       T_n = E[min(Poisson(E2), S_n)]  -- the reference's mean count, = E2 unless the node is close to saturation.
 *     Independent per-agent trials (lpk_tx_infect) then have the reference's per-agent marginals, node mean and node
 *     variance.  tau = 3e38 means "everybody" (T_n = S_n; the reference takes min(count, susceptibles)).
 *       network        double[num_nodes * num_nodes] row-major, W[i,j] = fraction moving i -> j
 *       r0_scalars     double[num_nodes];  alive_counts int32[num_nodes] (results.pop[t], model.py:1344)
 *     outputs: tau float[num_nodes], strain_cdf double[num_nodes * n_strains] (cumulative p[n,s] / P_n),
 *              prob double[num_nodes * n_strains], expected double[num_nodes].
 *     ws: caller-owned scratch, double[2 * num_nodes] (row sums of W, recomputed every call because the reference
 *     re-reads tx.network each tick, model.py:1335; and the gated targets). */
int lpk_tx_node_math(int32_t num_nodes, int32_t n_strains, const int64_t *beta_fx, const int64_t *exposure_fx,
                     const int32_t *risk_hist, const double *network, double beta_seasonality, const double *r0_scalars,
                     const int32_t *alive_counts, double zero_inflation, double dispersion, float *tau,
                     double *strain_cdf, double *prob, double *expected, double *ws, const lpk_rng *rng, void *stream);

/* T3  replaces tx_infect_nb(num_nodes, num_people, num_strains, sus_by_node, node_ids, strain, disease_state,
 *     sus_indices_storage, sus_probs_storage, risks, prob_exp_by_node_strain, n_exposures_to_create_by_node_strain)
 *     model.py:1010-1149.  Per-agent Bernoulli: susceptible i of node n is exposed iff
 *     x_i < floor(p_i * 2^32), p_i = 1 - exp(-risk_i * tau[n]) (evaluated by an exactly specified fmaf polynomial),
 *     x_i = h16 << 16 | l16, the half-words hw = ((i >> 7) & 1) * 4 + (i & 3) of the two blocks
 *     Philox(seed; (i >> 8) * 32 + ((i >> 2) & 31), tick, EXPOSE / EXPOSE_LO) (i = agent index + id_base: one block
 *     serves the 8 agents a lane owns in a pair of 128-agent rows; the fused pass generates the low block only when
 *     the high half cannot decide); strain by the cumulative categorical of model.py:1127-1141.
 *     No bucket pass, no scratch columns.  n_new[num_nodes * num_strains] OVERWRITTEN. */
int lpk_tx_infect(int32_t num_nodes, int64_t num_people, int32_t num_strains, const int16_t *node_ids,
                  int8_t *strain, int8_t *disease_state, const float *risks, const float *tau,
                  const double *strain_cdf, int32_t *n_new, const lpk_rng *rng, void *stream);

/* C1  replaces count_SEIRP(node_id, disease_state, strain, potentially_paralyzed, paralyzed, n_nodes, n_strains,
 *     n_people)  model.py:869-929.  All eight outputs OVERWRITTEN (the reference returns fresh arrays):
 *     S, E, I, R, POTP, P int32[n_nodes]; E_by_strain, I_by_strain int32[n_nodes * n_strains]. */
int lpk_count_seirp(const int16_t *node_id, const int8_t *disease_state, const int8_t *strain,
                    const int8_t *potentially_paralyzed, const int8_t *paralyzed, int32_t n_nodes,
                    int32_t n_strains, int64_t n_people, int32_t *S, int32_t *E, int32_t *I, int32_t *R,
                    int32_t *E_by_strain, int32_t *I_by_strain, int32_t *POTP, int32_t *P, void *stream);

/* =====================================================================================================
 * Fused tick: the fast path behind SEIR_ABM.run() (reference model.py:246-289).
 *
 * One streaming pass over the agent table per simulated day instead of one per component.  Because every
 * draw is keyed on (seed, agent, tick, stage), fusing changes no result: the pass for tick t executes, per
 * agent and in the reference's order,
 *     [pending from tick t-1]  tx_infect (model.py:1010-1149 replacement) and the census (model.py:869-929)
 *     [tick t]                 get_deaths (1767-1781), disease_state_step (344-454), fast_ri (1805-1855),
 *                              fast_sia (1994-2060; one campaign event per tick), tx_step_prep tally (932-1007)
 * and lpk_tick_node then does the node-level work of tick t (model.py:1332-1351 + population / paralysis
 * bookkeeping), producing q / strain_cdf that the NEXT pass applies.  seed_schedule days, days with several
 * campaign events and the final tick are run through the per-function entry points above (the host drains the
 * pending exposure first).
 *
 * All result rows are device pointers to row t (or t-1) of the [nt, nodes(, strains)] int32 arrays; counts are
 * accumulated with atomics, so "=" rows must be zero beforehand (they are: fresh result arrays) and "+=" rows
 * (R on top of pre-seeded immunes, new_exposed shared with RI/SIA) keep the reference's semantics for free.
 * ===================================================================================================== */
#define LPK_TILE_AGENTS 512

typedef struct lpk_people {
    int8_t *disease_state, *strain, *exposure_timer, *infection_timer, *paralysis_timer;
    int8_t *potentially_paralyzed, *paralyzed, *ipv_protected;
    const uint8_t *chronically_missed;
    const int16_t *node_id;
    int16_t *ri_timer;               /* NULL when RI_ABM is not a component */
    const float *acq_risk_multiplier, *daily_infectivity;
    const int32_t *date_of_birth;    /* read on SIA days only (LPK_F_SIA); may be NULL otherwise */
    const int32_t *date_of_death;    /* NULL when VitalDynamics_ABM is not a component */
    const int32_t *tile_node;        /* [ceil(capacity / 512)]: node id shared by every agent slot of the tile, or -1;
                                        NULL = always read node_id (lpk_build_tile_nodes fills it) */
    int64_t capacity;
    /* device-only companions of the table, owned by the caller, filled by lpk_hot_build (see "Agenda bytes" below) */
    uint8_t *hot;                    /* [lpk_hot_padded(capacity)] one agenda byte per slot, 16-byte aligned */
    int32_t *pair_min_dod;           /* [lpk_hot_padded(capacity) / 256] earliest date_of_death among the alive agents of
                                        each 256-slot pair (INT32_MAX: nobody); NULL when there is no date_of_death */
    int32_t risk_e0;                 /* exponent bias of the 6-bit risk code: lpk_hot_risk_e0(largest finite risk in the table) */
    int32_t *pair_ri_max;            /* [lpk_hot_padded(capacity) / 256] largest stored ri_timer among the alive, not chronically
                                        missed agents of each pair (INT32_MIN: nobody); NULL when there is no ri_timer */
    uint64_t *rec;                   /* [capacity] per-agent event record while the table is on the fused path: the bytes
                                        {disease_state, strain, exposure_timer, infection_timer, paralysis_timer,
                                        potentially_paralyzed, paralyzed, ipv_protected}, timers as deadlines (see below).
                                        The pass reads and writes THIS instead of the eight columns; lpk_hot_build fills
                                        it from them, lpk_hot_settle writes it back */
    uint8_t *ri_k;                   /* [lpk_hot_padded(capacity)] which RI tick after lpk_hot_build (1, 2, ... 254) finds the agent
                                        eligible, 0 = none of them (expired, chronically missed, dead); NULL when there is no
                                        ri_timer.  The pass reads this byte on RI ticks instead of ri_timer + chronically_missed */
} lpk_people;

#define LPK_F_PENDING 1u /* apply tick-1's exposure (q_prev / cdf_prev) and take tick-1's census */
#define LPK_F_STAGES 2u  /* run tick t's own stages and tally */
#define LPK_F_DEATHS 4u  /* tick t is a vital-dynamics tick: mark deaths (needs date_of_death) */
#define LPK_F_RI 8u      /* tick t is a routine-immunisation tick (needs ri_timer) */
#define LPK_F_SIA 16u    /* tick t carries ONE campaign event (needs date_of_birth and the sia_* fields) */
#define LPK_F_ROWSUMS 32u /* lpk_tick_node only: rowsum_ws[0 .. nodes) already holds the row sums of `network` (the caller
                             sets it from the second tick on while the network is unchanged; saves re-reading the matrix) */

typedef struct lpk_tick_args {
    uint32_t flags;
    int32_t tick; /* t */
    int32_t n_nodes, n_strains;
    uint64_t seed, id_base;
    const int64_t *counts; /* device int64[2] = {agents alive-or-dead in the table when tick t-1 ended, agents now} */
    /* ---- pending exposure + census of tick t-1 (LPK_F_PENDING) */
    const float *q_prev;    /* [nodes]  tau of tick t-1, from lpk_tick_node / lpk_tx_node_math */
    const double *cdf_prev; /* [nodes, strains] */
    int32_t *new_exposed_prev, *new_exposed_by_strain_prev; /* rows t-1, += (the S / E / I / R rows of t-1 are written by
                                                               lpk_tick_node from the carried counts below) */
    int32_t *tx_hits;           /* [nodes] scratch, += : exposures of tick t-1 found by this pass; consumed by lpk_tick_node */
    int32_t *tx_hits_by_strain; /* [nodes, strains] scratch, += : the same per strain */
    /* ---- stages of tick t (LPK_F_STAGES) */
    float p_paralysis;
    int32_t *new_potential, *new_paralyzed; /* rows t, += */
    int32_t *deaths, *dead_pp, *dead_par;   /* [nodes] scratch, += : deaths, and how many of the dying had
                                               potentially_paralyzed == 1 / paralyzed == 1 (census bookkeeping) */
    int32_t ri_step;
    int32_t ri_strain;
    const double *vx_prob_ri, *vx_prob_ipv;                          /* [nodes] */
    int32_t *ri_vaccinated, *ri_protected, *ipv_vaccinated;          /* rows t */
    int32_t *new_exposed, *new_exposed_by_strain, *ri_new_exposed_by_strain; /* rows t (new_exposed* shared with SIA) */
    /* ---- one SIA campaign event on tick t (LPK_F_SIA), same meaning as lpk_fast_sia's arguments; an agent's draw is
     *      Philox(seed; agent, tick, SIA | event_idx << 8) exactly as there.  Rows t, += (zero beforehand). */
    const uint8_t *sia_targeted;   /* [nodes] nodes_to_vaccinate */
    const float *vx_prob_sia;      /* [nodes] */
    double sia_vx_eff;
    int32_t sia_min_age, sia_max_age, sia_strain;
    uint32_t sia_event_idx;
    int32_t *sia_vaccinated, *sia_protected, *sia_new_exposed_by_strain;
    double strain_r0_scalars[LPK_MAX_STRAINS];
    int64_t *beta_fx;      /* [nodes, strains] infectivity tally, CARRIED: sum over the infectious agents of
                              round(infectivity * strain_r0_scalar * 2^30); the pass adds an agent's term when it turns
                              infectious and subtracts it when it recovers or dies (exact integers: equals
                              lpk_tx_step_prep's from-scratch tally bit for bit) */
    int32_t *E_cur, *I_cur; /* [nodes, strains] exposed / infectious agents per strain, carried the same way; initialised
                               from lpk_count_seirp's E_by_strain / I_by_strain */
    int64_t *exposure_fx, *sus; /* susceptible-side tallies, CARRIED from tick to tick: the pass only corrects them when */
    int32_t *risk_hist;    /* an agent leaves the susceptible state (hit, RI exposure, death); lpk_vd_births adds cohorts.
                              Exact integers, so they equal lpk_tx_step_prep's from-scratch values; the caller initialises
                              them with lpk_tx_step_prep and re-initialises after any tick run outside the pass */
    int32_t *R_cur;        /* [nodes] recovered agents per node, carried the same way (+1 on recovery, -1 when a recovered
                              agent dies); initialised / re-initialised from lpk_count_seirp's R */
    /* scheduling hint (results do not depend on it): slots [0, uniform_agents) hold the node-contiguous initial population,
     * slots beyond it appended newborn cohorts; the pass hands the cohort region out first.  0 = unknown */
    int64_t uniform_agents;
    /* Lazy RI countdown.  fast_ri subtracts ri_step from every alive, not chronically missed agent's ri_timer on every RI
     * tick (model.py:1825-1833) although only agents a few weeks old can ever become eligible.  The pass does not write
     * the column: ri_lazy_k = RI ticks run since the last lpk_hot_build whose subtraction is still owed (excluding tick t
     * itself); an agent's current timer is stored - ri_lazy_k * ri_step, the column is read only in pairs where
     * pair_ri_max says somebody can still become eligible, an agent that dies gets its debt settled on the spot, newborns
     * enter with the debt added (lpk_births_args.ri_lazy_k), and lpk_hot_settle pays it for everybody else.  ri_step must
     * be set on every tick while ri_lazy_k != 0. */
    int32_t ri_lazy_k;
    uint32_t *work_counter; /* caller-owned device uint32 (one per table / stream): the pass zeroes it on `stream` and claims
                               work units from it, so two tables on one device never share scheduling state */
    uint32_t *work_counter_next; /* optional (lpk_run_days): when set, *work_counter is taken to be zero already (no memset
                               node on the stream) and this launch zeroes *work_counter_next for the launch after it */
} lpk_tick_args;

/* tile_node[k] for tiles first_tile .. last tile covering [0, n_slots): node id if node_id is constant over the
 * tile's slots (unborn slots carry -1 and make a tile mixed), else -1. */
int lpk_build_tile_nodes(const int16_t *node_id, int64_t first_tile, int64_t n_slots, int32_t *tile_node, void *stream);

int lpk_tick_pass(const lpk_people *people, const lpk_tick_args *args, void *stream);

/* ---- Agenda bytes: what the pass streams (csrc/lpk_hot.cuh) ----------------------------------------------------------
 * lpk_tick_pass reads ONE byte per agent per day, people->hot[i]:
 *   bits 7:6 class: 00 exposed, 01 infectious, 10 inactive (payload 0 recovered, 1 dead / unborn), 11 susceptible
 *   bits 5:0 susceptible: 6-bit upper bound of acq_risk_multiplier (4 steps per octave; 2^(e - risk_e0) * (1 + m / 4));
 *            exposed / infectious: day of the agent's next event (E -> I, paralysis gate, I -> R) modulo 64
 * and touches full-width data only for agents with an event: the 8-byte record people->rec[i] (the eight byte columns of
 * the disease state side by side) plus risk / infectivity / dates where the event needs them.  While an agent is exposed /
 * infectious the countdown bytes of its record hold DEADLINES, (timer + tick) mod 256 -- exposure_timer while
 * disease_state == 1, infection_timer while == 2, paralysis_timer while == 2 and strain == 0 -- so nothing is decremented
 * on the days in between (the reference decrements daily, model.py:419-452; the tested values are identical, int8
 * wrap-around included).  The reference-dtype columns are not touched by the pass and are stale until lpk_hot_settle.
 *   lpk_hot_build   canonical columns (as every per-function entry point above reads them) -> records + agenda bytes +
 *                   pair_min_dod / pair_ri_max / ri_k, for the table as it stands BEFORE tick `tick_next`.  *status = 2 if a risk exceeds the
 *                   range of the code (risk_e0 too large); slots >= n_slots and the padding read as dead.
 *   lpk_hot_settle  the inverse: records -> the eight columns (deadline -> the value tick `tick_next` would test; ri_timer -=
 *                   ri_lazy_k * ri_step for the alive, not chronically missed agents, see lpk_tick_args.ri_lazy_k): call
 *                   before handing the table to the per-function entry points or to the host. */
int lpk_hot_build(const lpk_people *people, int64_t n_slots, int32_t tick_next, int32_t ri_step, int32_t *status, void *stream);
int lpk_hot_settle(const lpk_people *people, int64_t n_slots, int32_t tick_next, int32_t ri_lazy_k, int32_t ri_step, void *stream);
int64_t lpk_hot_padded(int64_t capacity); /* capacity rounded up to the pass's work unit (2048 agents) */
int32_t lpk_hot_risk_e0(float max_risk);

typedef struct lpk_node_args {
    uint32_t flags; /* LPK_F_PENDING: finish rows t-1 (E, I totals); LPK_F_DEATHS: tick t is a vital-dynamics tick */
    int32_t tick, n_nodes, n_strains;
    uint64_t seed;
    /* transmission node math of tick t (same meaning as lpk_tx_node_math) */
    const int64_t *beta_fx, *exposure_fx;
    const int32_t *risk_hist;
    const double *network, *r0_scalars;
    double beta_seasonality, zero_inflation, dispersion;
    float *q; /* tau */
    double *strain_cdf, *prob, *expected, *rowsum_ws; /* rowsum_ws: double[2 * nodes] */
    /* population bookkeeping: pop[t] = pop[t-1] + births[t] - deaths (model.py:1751-1755); NULL pop rows = no VD */
    const int32_t *pop_prev;
    int32_t *pop, *births_row, *deaths_row;
    int32_t *deaths, *dead_pp, *dead_par; /* consumed and zeroed */
    /* paralysis census kept incrementally: cur += new - dead, row t = cur (model.py:1482-1483 equivalent) */
    int32_t *cur_potp, *cur_p;
    const int32_t *new_potential, *new_paralyzed; /* rows t */
    int32_t *potp_row, *p_row;                     /* rows t */
    /* exposed / infectious census of t-1 from the carried counts (model.py:1477-1480 equivalent): with LPK_F_PENDING,
     *   E_by_strain_prev = E_snap + tx_hits_by_strain, I_by_strain_prev = I_snap ("="), E_prev / I_prev their sums;
     * then tx_hits_by_strain is zeroed and the snapshots are retaken for tick t: E_snap = E_cur, I_snap = I_cur. */
    int32_t *E_by_strain_prev, *I_by_strain_prev; /* rows t-1 */
    int32_t *E_prev, *I_prev;                     /* rows t-1 */
    const int32_t *E_cur, *I_cur;
    int32_t *E_snap, *I_snap, *tx_hits_by_strain;
    int32_t *any_cases; /* optional: one int32, set to 1 when some node still has an exposed or infectious agent after
                           tick t's stages (the early-stop test of the next tick, model.py:789-795, reads it) */
    /* S and R census from the carried per-node counts (model.py:1476-1481 equivalent): with LPK_F_PENDING,
     *   S_prev[n] = S_snap[n] - tx_hits[n]   (susceptibles when tick t-1's stages ended, minus tick t-1's exposures)
     *   R_prev[n] += R_snap[n]               ("+=": on top of the pre-seeded non-agent immunes, model.py:1481)
     * then tx_hits is zeroed and the snapshots are retaken for tick t: S_snap = sus, R_snap = R_cur. */
    const int64_t *sus;
    const int32_t *R_cur;
    int32_t *tx_hits, *S_snap, *R_snap;
    int32_t *S_prev, *R_prev; /* rows t-1 */
    int64_t *counts; /* counts[0] = counts[1] once tick t is complete */
    /* node shard: only nodes [node_lo, node_hi) are this rank's -- their rows are written, their columns of
     * the network are read (1 / world of the matrix per tick); node_hi == 0 means every node */
    int32_t node_lo, node_hi;
    /* tally exchange of a sharded run (lpk_xchg, set by lpk_run_days): beta_fx is this rank's receive buffer, and the node
     * kernels wait until xchg_flags[r] has reached xchg_seq for every rank r < xchg_world before reading it; NULL = no wait */
    const uint32_t *xchg_flags;
    int32_t xchg_world;
    uint32_t xchg_seq;
    /* optional caller-owned scratch for n_nodes > 1024 (the transfer is then summed per chunk of 1024 source rows):
     * double[ceil(n_nodes / 1024) * LPK_MAX_STRAINS * n_nodes]; NULL = a library-owned buffer per device (not safe for two
     * tables on one device from different streams) */
    double *matvec_ws;
} lpk_node_args;

int lpk_tick_node(const lpk_node_args *args, void *stream);

/* V2  replaces the births block of VitalDynamics_ABM.step (model.py:1711-1734): per node
 *     births = floor(e) + Bernoulli(e - floor(e)), e = step_size * birth_rate * pop[t-1]; the cohort is appended
 *     node-major at [count, count + sum(births)) with date_of_birth = t, date_of_death = t + lifespan(age 0) drawn by
 *     inverse CDF on cum_deaths (laser-core KaplanMeierEstimator.predict_age_at_death semantics: year by
 *     searchsorted-left on the cumulative table, uniform day within the year), disease_state = 0.  Draws are
 *     Philox(seed; node, tick, BIRTH) / Philox(seed; agent, tick, LIFESPAN) instead of the host numpy stream.
 *     counts[1] += sum(births); if that would exceed capacity nobody is born, status[0] |= 1 and status[1] = tick (the
 *     device analogue of LaserFrame.add raising; the host mirrors status behind an event and raises one call later). */
typedef struct lpk_births_args {
    int32_t tick, n_nodes;
    uint64_t seed, id_base;
    double step_size;            /* days covered by one vital-dynamics step (pars.step_size_VitalDynamics_ABM) */
    const double *birth_rate;    /* [nodes] births per capita per day = cbr / (365 * 1000) */
    const int32_t *pop_prev;     /* [nodes] results.pop[t-1] */
    int32_t *births_row;         /* [nodes] results.births[t], overwritten */
    int64_t *counts;             /* device int64[2], see lpk_tick_args.counts */
    int64_t capacity;
    const int64_t *cum_deaths;   /* [max_year + 2], cum_deaths[0] = 0 */
    int32_t max_year;
    int32_t ri_newborn_timer;    /* < 0: leave ri_timer as pre-set (reference behaviour); else value for newborns */
    int32_t *node_offsets_ws;    /* scratch int32[nodes + 1] */
    int64_t *cohort_ws;          /* scratch int64[2] = {first slot, size} of the cohort just created */
    int32_t *status;             /* device int32[2]: flag bits (1 = a cohort did not fit: sticky, no later cohort is created
                                    either; 2 = a risk beyond the agenda code), and the tick of the first overflow */
    int8_t *disease_state;
    int16_t *node_id;
    int32_t *date_of_birth, *date_of_death;
    int16_t *ri_timer;           /* may be NULL */
    int32_t *tile_node;          /* may be NULL */
    /* optional (all or none): carried susceptible-side tallies to which the cohort is added, see lpk_tick_args */
    const float *acq_risk_multiplier;
    int64_t *sus, *exposure_fx;
    int32_t *risk_hist;
    /* optional: agenda bytes of the table (lpk_people.hot / pair_min_dod / risk_e0); newborns enter as susceptibles */
    uint8_t *hot;
    int32_t *pair_min_dod;
    int32_t risk_e0;
    uint64_t *rec;               /* the newborn's event record (lpk_people.rec), from the pre-drawn columns of its slot */
    const int8_t *strain, *exposure_timer, *infection_timer, *paralysis_timer, *potentially_paralyzed, *paralyzed, *ipv_protected;
    uint8_t *ri_k;               /* with ri_timer: the newborn's eligibility tick (lpk_people.ri_k) */
    int32_t *pair_ri_max;        /* ... its timer enters the pair's maximum ... */
    int32_t ri_lazy_k, ri_step;  /* ... stored with the lazy countdown's debt added (lpk_tick_args.ri_lazy_k) */
} lpk_births_args;

int lpk_vd_births(const lpk_births_args *args, void *stream);

/* ---- A run of fused days driven from C (SURVEY.md 7 "hard parts": host glue; 8e: the tick is launched without the host) ----
 * The reference's loop body is Python per tick (model.py:252-263).  At 2.2e8 agents a fused day is ~0.3 ms of device time
 * and at 8 GPUs ~0.05 ms, less than the Python that assembles its arguments, so the loop itself lives here:
 * lpk_run_days() derives every day's lpk_births_args / lpk_tick_args / lpk_node_args from ONE template (row pointers are
 * base + tick * row length) and launches vital dynamics -> pass -> [tally exchange] -> node kernels for n_days consecutive
 * days back to back on `stream`, with no host work in between other than the launches (~4 per plain day).  With
 * run->graph != 0 the days are captured into a CUDA graph and launched as one.  Results are those of calling
 * lpk_vd_births / lpk_tick_pass / lpk_tick_node day by day (tests/test_gpu_fused.py runs both). */
typedef struct lpk_rows { /* device results arrays of the run: [nt, nodes] or [nt, nodes, strains] int32; NULL = absent */
    int32_t *S, *E, *I, *R, *pop, *births, *deaths, *new_exposed, *new_potentially_paralyzed, *new_paralyzed;
    int32_t *potentially_paralyzed, *paralyzed, *ri_vaccinated, *ri_protected, *ipv_vaccinated, *sia_vaccinated, *sia_protected;
    int32_t *E_by_strain, *I_by_strain, *new_exposed_by_strain, *ri_new_exposed_by_strain, *sia_new_exposed_by_strain;
    int32_t *sink; /* [nodes * strains] scratch row that stands in for the rows of absent arrays */
} lpk_rows;

typedef struct lpk_day { /* what changes from day to day */
    int32_t tick;
    uint32_t flags;          /* LPK_F_DEATHS | LPK_F_RI | LPK_F_SIA of this day (PENDING / STAGES / ROWSUMS are the driver's) */
    double beta_seasonality; /* get_seasonality(sim) of the day (utils.py:616-625) */
    const uint8_t *sia_targeted; /* the day's single campaign event (LPK_F_SIA): lpk_tick_args.sia_* */
    double sia_vx_eff;
    int32_t sia_min_age, sia_max_age, sia_strain, _pad;
} lpk_day;

/* Tally exchange of a node-sharded run: every rank's rows [node_lo, node_hi) of the nodes x strains infectivity
 * tally, written straight into every peer's HBM over NVLink by a kernel between the pass and the node kernels (peer
 * memory mapped through CUDA IPC; one process per GPU), followed by a system-scope flag; the node kernels wait on the
 * flags of all ranks before they read the gathered tally.  No NCCL call, no host synchronisation per tick.  Rows are
 * disjoint between ranks, so gathering them IS the all-reduce (sum) of the reference design, and it is exact. */
typedef struct lpk_xchg lpk_xchg;
#define LPK_XCHG_HANDLE_BYTES 64
/* allocate this rank's receive buffers (2 x n_elems int64 + flags) and export their IPC handle */
int lpk_xchg_create(int32_t rank, int32_t world, int64_t n_elems, lpk_xchg **out, void *handle_out);
/* handles: world x LPK_XCHG_HANDLE_BYTES, every rank's handle in rank order (exchanged by the caller, e.g. all_gather) */
int lpk_xchg_connect(lpk_xchg *x, const void *handles);
/* teardown in two phases: every rank disconnects (unmaps its peers), the caller runs a barrier, every rank destroys */
int lpk_xchg_disconnect(lpk_xchg *x);
int lpk_xchg_destroy(lpk_xchg *x);

typedef struct lpk_run {
    lpk_people people;
    lpk_tick_args tick;     /* template: constants, carried tallies, scratch; per-day fields are filled by the driver */
    lpk_node_args node;     /* template likewise (network, r0_scalars, q / cdf outputs, snapshots, node shard) */
    lpk_births_args births; /* template; births.capacity == 0: VitalDynamics_ABM is not a component */
    lpk_rows rows;
    const int32_t *zero_pop; /* [nodes] zeros: the population row when nobody maintains results.pop */
    int32_t *any_cases;      /* optional device int32[nt]: lpk_node_args.any_cases of tick t is any_cases + t */
    uint32_t *work_counters; /* device uint32[2], both zero before the first day: the passes alternate between them */
    lpk_xchg *xchg;          /* NULL: single table */
    int32_t pending;         /* in / out: the previous tick's exposure trial + census are still to be applied */
    int32_t ri_lazy_k;       /* in / out: lpk_tick_args.ri_lazy_k */
    int32_t rowsums_valid;   /* in / out: node.rowsum_ws holds the row sums of node.network */
    int32_t graph;           /* != 0: capture the days into a CUDA graph and launch that */
    uint32_t seq;            /* in / out: exchange sequence number (flags are compared against it) */
    int32_t _pad;
} lpk_run;

/* ms: optional HOST float[n_days][3]; when given, every day's births kernels, pass and node kernels (exchange included) are
 * bracketed by CUDA events on `stream`, the call synchronises at the end and writes the three elapsed times per day
 * (measurement runs only; forces graph off).  launches: optional, += the number of kernels launched. */
int lpk_run_days(lpk_run *run, const lpk_day *days, int32_t n_days, float *ms, int64_t *launches, void *stream);
/* the two halves of one day, for callers that do the tally exchange themselves between them (torch.distributed) */
int lpk_run_day_pass(lpk_run *run, const lpk_day *day, int64_t *launches, void *stream);
int lpk_run_day_node(lpk_run *run, const lpk_day *day, const int64_t *beta_all, int64_t *launches, void *stream);
/* waits for and frees the executable graphs of earlier lpk_run_days calls of this thread (they are otherwise freed lazily) */
int lpk_run_release(void);

/* ---- Population initialisation in HBM -------------------------------------------------------
 * The reference draws every per-agent column on the host at construction.  Each entry below replaces one of those
 * blocks with a kernel whose output is a pure function of (seed, agent id = slot + id_base, stage): Philox4x32-10,
 * counter = (id, block, stage), 53-bit uniforms taken two words at a time.  Slices [start, end) can be drawn in any
 * order and on any number of GPUs with the same result.  oracle/lp_oracle_init.c restates every sampler. */
#define LPK_STAGE_INIT_HET 16u
#define LPK_STAGE_INIT_EXP 17u
#define LPK_STAGE_INIT_INF 18u
#define LPK_STAGE_INIT_PAR 19u
#define LPK_STAGE_INIT_AGE 20u
#define LPK_STAGE_INIT_LIFE 21u
#define LPK_STAGE_INIT_RI 22u
#define LPK_STAGE_INIT_MISSED 23u

/* the lp.* distributions of reference distributions.py:17-135 (host objects ``Distribution(dist_type, **pars)``) */
#define LPK_DIST_CONSTANT 0    /* a = value */
#define LPK_DIST_EXPONENTIAL 1 /* a = scale */
#define LPK_DIST_GAMMA 2       /* a = shape, b = scale */
#define LPK_DIST_LOGNORMAL 3   /* a, b = mu, sigma of the underlying normal (the host converts lp.lognormal's mean / sigma) */
#define LPK_DIST_NORMAL 4      /* a = mean, b = std */
#define LPK_DIST_POISSON 5     /* a = lam */
#define LPK_DIST_UNIFORM 6     /* integers in [a, b)  (np.random.randint) */
typedef struct lpk_dist {
    int32_t kind;
    double a, b;
} lpk_dist;

/* populate_heterogeneous_values(start, end, acq_risk_out, infectivity_out, pars), reference model.py:816-866:
 * (z0, z1) standard normal, zc = rho * z0 + sqrt(1 - rho^2) * z1;  acq_risk = exp(mu_ln + sigma_ln * z0);
 * infectivity = gamma.ppf(norm.cdf(zc), a = 1, scale = scale_gamma) = -scale_gamma * log(erfc(zc / sqrt 2) / 2).
 * heterogeneity == 0: acq_risk = 1, infectivity = mean_gamma (pars.individual_heterogeneity False). */
int lpk_init_heterogeneity(int64_t start, int64_t end, float *acq_risk_out, float *infectivity_out, double mu_ln, double sigma_ln,
                           double scale_gamma, double rho, int32_t heterogeneity, double mean_gamma, uint64_t seed, uint64_t id_base,
                           void *stream);

/* DiseaseState_ABM.__init__ timers, reference model.py:571-587: exposure_timer = clip(int8(dur_exp), 0, 127),
 * infection_timer likewise, paralysis_timer = int8(clip(t_to_paralysis - exposure_timer, 0, min(infection_timer, 127))). */
int lpk_init_timers(int64_t start, int64_t end, int8_t *exposure_timer, int8_t *infection_timer, int8_t *paralysis_timer,
                    const lpk_dist *dur_exp, const lpk_dist *dur_inf, const lpk_dist *t_to_paralysis, uint64_t seed, uint64_t id_base,
                    void *stream);

/* VitalDynamics_ABM._initialize_ages_and_births / _initialize_deaths and RI_ABM._initialize_people_fields, reference
 * model.py:1578-1596, 1605-1611, 1893-1894, fused: age bin ~ pyramid counts (inverse CDF), age = randint(lo, hi) days
 * (0 -> 1), date_of_birth = -age; date_of_death = KaplanMeier age at death given the age (laser-core, restated in
 * core.py) - age; ri_timer = int16(int32(date_of_birth + U(42, 98))).  date_of_death / ri_timer may be NULL. */
typedef struct lpk_demog_args {
    int64_t start, end;
    int32_t *date_of_birth, *date_of_death;
    int16_t *ri_timer;
    const double *bin_cdf;             /* device [n_bins] cumulative pyramid counts */
    const int32_t *bin_lo, *bin_hi;    /* device [n_bins] age range of the bin in days: [lo, hi) */
    int32_t n_bins;
    const int64_t *cum_deaths;         /* device [max_year + 2]: cum_deaths[y] = deaths before age y (leading 0) */
    int32_t max_year;
    uint64_t seed, id_base;
} lpk_demog_args;
int lpk_init_demography(const lpk_demog_args *args, void *stream);

/* chronically_missed, reference model.py:154-159: exactly n_missed of the n agents, uniformly without replacement
 * (the n_missed smallest 64-bit Philox keys; found by a 4-pass radix select that never stores the keys).
 * ws: device scratch of LPK_MISSED_WS_WORDS uint32, 8-byte aligned. */
#define LPK_MISSED_WS_WORDS (65536 + 4)
int lpk_init_missed(int64_t n, int64_t n_missed, uint8_t *chronically_missed, uint64_t seed, uint64_t id_base, uint32_t *ws, void *stream);

/* ---- The infection-migration network in HBM ---------------------------------------------------
 * Transmission_ABM._initialize_common, reference model.py:1216-1258; laser-core's gravity / radiation / row_normalizer /
 * distance (~=0.6, not in the checkout) as restated in laser-polio_b200/core.py.  All matrices device float64 [n, n]. */
/* all-pairs Haversine (km, R = 6371); d == 0 off the diagonal -> epsilon (model.py:1232-1238) */
int lpk_net_haversine(const double *lat_deg, const double *lon_deg, int32_t n, double epsilon, double *dist, void *stream);
/* net[i, j] = k * p_i^a * p_j^b * d_ij^-c / norm, zero diagonal; norm = (sum p)^c (model.py:1243-1251) */
int lpk_net_gravity(const double *pops, const double *dist, int32_t n, double k, double a, double b, double c, double norm, double *net,
                    void *stream);
/* T_ij = k p_i p_j / ((p_i + s_ij)(p_i + p_j + s_ij)), s_ij = population within d_ij of i without j (and without i unless
 * include_home); n <= 8192 (model.py:1252-1254) */
int lpk_net_radiation(const double *pops, const double *dist, int32_t n, double k, int32_t include_home, double *net, void *stream);
/* rows whose sum exceeds max_rowsum are rescaled to it, in place (model.py:1258) */
int lpk_net_row_normalize(double *net, int32_t n, double max_rowsum, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LPK_H */
