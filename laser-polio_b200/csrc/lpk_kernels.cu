// lpk_kernels.cu -- sm_100a kernels + C-ABI launchers, one per reference hot-path function
// (include/lpk.h cites the reference signature each replaces).  All kernels are HBM-bound
// streaming scans over the agent structure-of-arrays; see lpk_common.cuh for the quad tiling.
#include <cstdio>
#include <cstring>

#include <cstdlib>

#include "lpk_host.cuh"
#include "lpk_node.cuh"
#include "lpk_stages.cuh"

// ------------------------------------------------------------------ error plumbing
thread_local char lpk_g_err[256] = "";
extern "C" const char *lpk_last_error(void) { return lpk_g_err; }
extern "C" int lpk_version(void) { return 2; }

int lpk_sm_count() {
    static int count = 0;
    if (count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || count <= 0)
            count = 148;
    }
    return count;
}
#define agent_grid lpk_agent_grid

// ------------------------------------------------------------------ Philox self test
__global__ void k_philox_selftest(const uint32_t *ctr, const uint32_t *key, uint32_t *out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o[4];
    philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1], o);
    for (int k = 0; k < 4; ++k) out[4 * i + k] = o[k];
}
extern "C" int lpk_philox_selftest(const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4, int64_t n, void *stream) {
    REQUIRE(ctr4 && key2 && out4 && n >= 0, "philox_selftest");
    if (n == 0) return LPK_OK;
    k_philox_selftest<<<(unsigned)((n + 127) / 128), 128, 0, as_stream(stream)>>>(ctr4, key2, out4, n);
    CUDA_TRY(cudaGetLastError(), "lpk_philox_selftest");
    return LPK_OK;
}

// ------------------------------------------------------------------ V1 get_deaths
__global__ void __launch_bounds__(LPK_BLOCK) k_get_deaths(int64_t n, int8_t *__restrict__ state,
                                                           const int16_t *__restrict__ node_id,
                                                           const int32_t *__restrict__ dod, int32_t t,
                                                           int32_t *__restrict__ num_dying) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TileRange tr = block_tiles(n);
    for (int64_t tile = tr.lo + warp; tile < tr.hi; tile += LPK_WARPS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t base = quad_base(tile, j, lane);
            const int valid = quad_valid(base, n);
            if (valid == 0) continue;
            const uint32_t w = load_b4(state, base, valid);
            if ((w & 0x80808080u) == 0x80808080u) continue;  // whole quad dead / unborn
            int d[4];
            load_i4(dod, base, valid, d);
            uint32_t nw = w;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (byte_of(w, k) >= 0 && d[k] <= t) {
                    nw = set_byte(nw, k, -1);
                    atomicAdd(&num_dying[node_id[base + k]], 1);
                }
            }
            if (nw != w) store_b4(state, base, valid, nw);
        }
    }
}
extern "C" int lpk_get_deaths(int32_t num_nodes, int64_t num_people, int8_t *disease_state, const int16_t *node_id,
                              const int32_t *date_of_death, int32_t t, int32_t *num_dying, void *stream) {
    REQUIRE(num_nodes > 0 && num_people >= 0, "get_deaths sizes");
    REQUIRE(disease_state && node_id && date_of_death && num_dying, "get_deaths null pointer");
    REQUIRE(ALIGNED(disease_state, 4) && ALIGNED(date_of_death, 16), "get_deaths alignment");
    CUDA_TRY(cudaMemsetAsync(num_dying, 0, sizeof(int32_t) * num_nodes, as_stream(stream)), "get_deaths memset");
    if (num_people == 0) return LPK_OK;
    k_get_deaths<<<agent_grid(num_people, 8), LPK_BLOCK, 0, as_stream(stream)>>>(num_people, disease_state, node_id,
                                                                                 date_of_death, t, num_dying);
    CUDA_TRY(cudaGetLastError(), "lpk_get_deaths");
    return LPK_OK;
}

// ------------------------------------------------------------------ D1 disease_state_step
__global__ void __launch_bounds__(LPK_BLOCK) k_disease_state(int64_t n, const int16_t *__restrict__ node_id,
                                                              int8_t *__restrict__ state, const int8_t *__restrict__ strain,
                                                              int8_t *etimer, int8_t *itimer, int8_t *pot_par,
                                                              int8_t *paralyzed, const int8_t *__restrict__ ipv,
                                                              int8_t *ptimer, double p_paralysis, int32_t *new_pot,
                                                              int32_t *new_par, DevRng rng) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TileRange tr = block_tiles(n);
    for (int64_t tile = tr.lo + warp; tile < tr.hi; tile += LPK_WARPS) {
        uint32_t w[4];
        int64_t base[4];
        int valid[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // all four state words in flight before any is used
            base[j] = quad_base(tile, j, lane);
            valid[j] = quad_valid(base[j], n);
            w[j] = valid[j] ? load_b4(state, base[j], valid[j]) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!(any_byte_eq(w[j], 1u) || any_byte_eq(w[j], 2u))) continue;  // no E / I in this quad
            uint32_t nw = w[j];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int8_t s = byte_of(w[j], k);
                if (s == 1 || s == 2) {
                    const int8_t ns = ds_agent(base[j] + k, s, node_id, strain, etimer, itimer, pot_par, paralyzed, ipv,
                                               ptimer, p_paralysis, new_pot, new_par, rng);
                    nw = set_byte(nw, k, ns);
                }
            }
            if (nw != w[j]) store_b4(state, base[j], valid[j], nw);
        }
    }
}
extern "C" int lpk_disease_state_step(const int16_t *node_id, int32_t n_nodes, int8_t *disease_state,
                                      const int8_t *strain, int64_t active_count, int8_t *exposure_timer,
                                      int8_t *infection_timer, int8_t *potentially_paralyzed, int8_t *paralyzed,
                                      const int8_t *ipv_protected, int8_t *paralysis_timer, float p_paralysis,
                                      int32_t *new_potential, int32_t *new_paralyzed, const lpk_rng *rng, void *stream) {
    REQUIRE(n_nodes > 0 && active_count >= 0, "disease_state_step sizes");
    REQUIRE(node_id && disease_state && strain && exposure_timer && infection_timer && potentially_paralyzed &&
                paralyzed && ipv_protected && paralysis_timer && new_potential && new_paralyzed,
            "disease_state_step null pointer");
    REQUIRE(ALIGNED(disease_state, 4), "disease_state_step alignment");
    if (active_count == 0) return LPK_OK;
    k_disease_state<<<agent_grid(active_count, 8), LPK_BLOCK, 0, as_stream(stream)>>>(
        active_count, node_id, disease_state, strain, exposure_timer, infection_timer, potentially_paralyzed, paralyzed,
        ipv_protected, paralysis_timer, (double)p_paralysis, new_potential, new_paralyzed, dev_rng(rng));
    CUDA_TRY(cudaGetLastError(), "lpk_disease_state_step");
    return LPK_OK;
}

// ------------------------------------------------------------------ R1 fast_ri
__global__ void __launch_bounds__(LPK_BLOCK) k_fast_ri(int64_t n, int step, const int16_t *__restrict__ node_id,
                                                        int8_t *__restrict__ state, int8_t *strain, int8_t *ipv,
                                                        int16_t *__restrict__ ri_timer, int64_t sim_t,
                                                        const double *__restrict__ prob_ri,
                                                        const double *__restrict__ prob_ipv, int32_t *ri_counts,
                                                        int32_t *ri_protected, int32_t *ipv_counts,
                                                        const uint8_t *__restrict__ missed, int8_t vaccine_strain,
                                                        DevRng rng) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TileRange tr = block_tiles(n);
    const bool first = (sim_t == step), later = (sim_t > step);
    for (int64_t tile = tr.lo + warp; tile < tr.hi; tile += LPK_WARPS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t base = quad_base(tile, j, lane);
            const int valid = quad_valid(base, n);
            if (valid == 0) continue;
            const uint32_t w = load_b4(state, base, valid);
            if ((w & 0x80808080u) == 0x80808080u) continue;
            const uint32_t m = load_b4(reinterpret_cast<const int8_t *>(missed), base, valid, 1);
            int tm[4];
            load_s4(ri_timer, base, valid, tm);
            uint32_t nw = w;
            bool touched = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int8_t s = byte_of(w, k);
                if (s < 0 || byte_of(m, k) == 1) continue;
                const int timer = tm[k] - step;  // eligibility uses the unwrapped value, the store wraps to int16
                tm[k] = timer;
                touched = true;
                const bool eligible = first ? (timer <= 0 && timer >= -step) : (later && timer <= 0 && timer > -step);
                if (!eligible) continue;
                const int64_t i = base + k;
                const int nd = node_id[i];
                double u1, u2;
                if (rng.u1) { u1 = rng.u1[i]; u2 = rng.u2[i]; }
                else { uint32_t x[4]; philox_agent(rng.seed, (uint64_t)i + rng.id_base, rng.tick, LPK_STAGE_RI, x); u1 = u53(x[0], x[1]); u2 = u53(x[2], x[3]); }
                if (u1 < prob_ri[nd]) {
                    atomicAdd(&ri_counts[nd], 1);
                    if (s == 0) { nw = set_byte(nw, k, 1); strain[i] = vaccine_strain; atomicAdd(&ri_protected[nd], 1); }
                }
                if (u2 < prob_ipv[nd]) { atomicAdd(&ipv_counts[nd], 1); ipv[i] = 1; }
            }
            if (touched) {
                if (valid == 4) *reinterpret_cast<short4 *>(ri_timer + base) = make_short4((short)tm[0], (short)tm[1], (short)tm[2], (short)tm[3]);
                else for (int k = 0; k < valid; ++k) ri_timer[base + k] = (int16_t)tm[k];
            }
            if (nw != w) store_b4(state, base, valid, nw);
        }
    }
}
extern "C" int lpk_fast_ri(int64_t step_size, const int16_t *node_id, int8_t *disease_state, int8_t *strain,
                           int8_t *ipv_protected, int16_t *ri_timer, int64_t sim_t, const double *vx_prob_ri,
                           const double *vx_prob_ipv, int64_t num_people, int32_t n_nodes, int32_t *ri_counts,
                           int32_t *ri_protected, int32_t *ipv_counts, const uint8_t *chronically_missed,
                           int8_t ri_vaccine_strain, const lpk_rng *rng, void *stream) {
    REQUIRE(n_nodes > 0 && num_people >= 0 && step_size > 0 && step_size < 32768, "fast_ri sizes");
    REQUIRE(node_id && disease_state && strain && ipv_protected && ri_timer && vx_prob_ri && vx_prob_ipv && ri_counts &&
                ri_protected && ipv_counts && chronically_missed, "fast_ri null pointer");
    REQUIRE(!rng || !rng->u1 || rng->u2, "fast_ri needs both injected streams");
    REQUIRE(ALIGNED(disease_state, 4) && ALIGNED(chronically_missed, 4) && ALIGNED(ri_timer, 8), "fast_ri alignment");
    cudaStream_t st = as_stream(stream);
    CUDA_TRY(cudaMemsetAsync(ri_counts, 0, sizeof(int32_t) * n_nodes, st), "fast_ri memset");
    CUDA_TRY(cudaMemsetAsync(ri_protected, 0, sizeof(int32_t) * n_nodes, st), "fast_ri memset");
    CUDA_TRY(cudaMemsetAsync(ipv_counts, 0, sizeof(int32_t) * n_nodes, st), "fast_ri memset");
    if (num_people == 0) return LPK_OK;
    k_fast_ri<<<agent_grid(num_people, 8), LPK_BLOCK, 0, st>>>(num_people, (int)step_size, node_id, disease_state, strain,
                                                               ipv_protected, ri_timer, sim_t, vx_prob_ri, vx_prob_ipv,
                                                               ri_counts, ri_protected, ipv_counts, chronically_missed,
                                                               ri_vaccine_strain, dev_rng(rng));
    CUDA_TRY(cudaGetLastError(), "lpk_fast_ri");
    return LPK_OK;
}

// ------------------------------------------------------------------ S1 fast_sia
__global__ void __launch_bounds__(LPK_BLOCK) k_fast_sia(int64_t n, const int16_t *__restrict__ node_id,
                                                         int8_t *__restrict__ state, int8_t *strain,
                                                         const int32_t *__restrict__ dob, int64_t sim_t,
                                                         const float *__restrict__ vx_prob, double vx_eff,
                                                         const uint8_t *__restrict__ targeted, int64_t min_age,
                                                         int64_t max_age, int32_t *vaccinated, int32_t *protected_,
                                                         const uint8_t *__restrict__ missed, int8_t vaccine_strain,
                                                         uint32_t stage, DevRng rng) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TileRange tr = block_tiles(n);
    NodeAcc<2, 0> acc;
    acc.init();
    auto flush = [&](int nd, const int *ci, const long long *) {
        red_add(&vaccinated[nd], ci[0]);
        red_add(&protected_[nd], ci[1]);
    };
    for (int64_t tile = tr.lo + warp; tile < tr.hi; tile += LPK_WARPS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t base = quad_base(tile, j, lane);
            const int valid = quad_valid(base, n);
            if (valid == 0) continue;
            const uint32_t w = load_b4(state, base, valid);
            if ((w & 0x80808080u) == 0x80808080u) continue;
            const uint32_t m = load_b4(reinterpret_cast<const int8_t *>(missed), base, valid, 1);
            int d[4];
            load_i4(dob, base, valid, d);
            bool elig[4];
            bool any = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t age = sim_t - d[k];
                elig[k] = byte_of(w, k) >= 0 && byte_of(m, k) != 1 && min_age <= age && age <= max_age;
                any |= elig[k];
            }
            if (!any) continue;
            int nd[4];
            load_s4(node_id, base, valid, nd);
            uint32_t nw = w;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!elig[k] || targeted[nd[k]] == 0) continue;
                const int64_t i = base + k;
                double r;
                if (rng.u1) r = rng.u1[i];
                else { uint32_t x[4]; philox_agent(rng.seed, (uint64_t)i + rng.id_base, rng.tick, stage, x); r = u53(x[0], x[1]); }
                const double pv = (double)vx_prob[nd[k]];
                if (r < pv) {
                    acc.select(nd[k], flush);
                    acc.ci[0] += 1;
                    if (byte_of(w, k) == 0 && r < pv * vx_eff) {
                        nw = set_byte(nw, k, 1);
                        strain[i] = vaccine_strain;
                        acc.ci[1] += 1;
                    }
                }
            }
            if (nw != w) store_b4(state, base, valid, nw);
        }
    }
    acc.finish_warp(flush);
}
extern "C" int lpk_fast_sia(const int16_t *node_ids, int8_t *disease_states, int8_t *strain, const int32_t *dobs,
                            int64_t sim_t, const float *vx_prob, double vx_eff, int64_t count,
                            const uint8_t *nodes_to_vaccinate, int64_t min_age, int64_t max_age, int32_t n_nodes,
                            int32_t *vaccinated, int32_t *protected_, const uint8_t *chronically_missed,
                            int8_t sia_vaccine_strain, uint32_t event_idx, const lpk_rng *rng, void *stream) {
    REQUIRE(n_nodes > 0 && count >= 0, "fast_sia sizes");
    REQUIRE(node_ids && disease_states && strain && dobs && vx_prob && nodes_to_vaccinate && vaccinated && protected_ &&
                chronically_missed, "fast_sia null pointer");
    REQUIRE(ALIGNED(disease_states, 4) && ALIGNED(chronically_missed, 4) && ALIGNED(dobs, 16) && ALIGNED(node_ids, 8),
            "fast_sia alignment");
    cudaStream_t st = as_stream(stream);
    CUDA_TRY(cudaMemsetAsync(vaccinated, 0, sizeof(int32_t) * n_nodes, st), "fast_sia memset");
    CUDA_TRY(cudaMemsetAsync(protected_, 0, sizeof(int32_t) * n_nodes, st), "fast_sia memset");
    if (count == 0) return LPK_OK;
    k_fast_sia<<<agent_grid(count, 8), LPK_BLOCK, 0, st>>>(count, node_ids, disease_states, strain, dobs, sim_t, vx_prob,
                                                           vx_eff, nodes_to_vaccinate, min_age, max_age, vaccinated,
                                                           protected_, chronically_missed, sia_vaccine_strain,
                                                           LPK_STAGE_SIA | (event_idx << 8), dev_rng(rng));
    CUDA_TRY(cudaGetLastError(), "lpk_fast_sia");
    return LPK_OK;
}

// ------------------------------------------------------------------ T1 tx_step_prep (tally)
struct StrainScalars { double v[LPK_MAX_STRAINS]; };

__global__ void __launch_bounds__(LPK_BLOCK) k_tx_step_prep(int64_t n, int n_strains, const int8_t *__restrict__ strain,
                                                             StrainScalars srs, const int8_t *__restrict__ state,
                                                             const int16_t *__restrict__ node_id,
                                                             const float *__restrict__ infectivity,
                                                             const float *__restrict__ risk, int64_t *beta_fx,
                                                             int64_t *exposure_fx, int64_t *sus, int32_t *risk_hist) {
    __shared__ int s_hist[LPK_WARPS][LPK_RISK_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TileRange tr = block_tiles(n);
    NodeAcc<1, 1 + LPK_MAX_STRAINS> acc;
    acc.init();
    WarpHist wh;
    wh.init(s_hist[warp], lane);
    auto flush = [&](int nd, const int *ci, const long long *cl) {
        if (ci[0]) atomicAdd(reinterpret_cast<unsigned long long *>(&sus[nd]), (unsigned long long)ci[0]);
        red_add(&exposure_fx[nd], cl[0]);
#pragma unroll
        for (int s = 0; s < LPK_MAX_STRAINS; ++s)
            if (s < n_strains) red_add(&beta_fx[(int64_t)nd * n_strains + s], cl[1 + s]);
    };
    for (int64_t tile = tr.lo + warp; tile < tr.hi; tile += LPK_WARPS) {
        uint32_t w[4];
        int64_t base[4];
        int valid[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            base[j] = quad_base(tile, j, lane);
            valid[j] = quad_valid(base[j], n);
            w[j] = valid[j] ? load_b4(state, base[j], valid[j]) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool anyS = any_byte_eq(w[j], 0u), anyI = any_byte_eq(w[j], 2u);
            int nd[4] = {-1, -1, -1, -1};
            float rk[4] = {0.f, 0.f, 0.f, 0.f};
            if (anyS || anyI) load_s4(node_id, base[j], valid[j], nd);
            if (anyS) load_f4(risk, base[j], valid[j], rk);
            // the histogram of the susceptibles' risks goes through the warp's shared-memory bins when the whole warp is
            // on one node (the common case); quads with mixed nodes fall back to global atomics
            int mine = -1;  // -1: no susceptible in this quad, -2: its susceptibles sit in different nodes
            if (anyS) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (byte_of(w[j], k) == 0) mine = (mine == -1 || mine == nd[k]) ? nd[k] : -2;
            }
            const int ref = __reduce_max_sync(LPK_FULL, mine);
            const bool uni = ref >= 0 && __all_sync(LPK_FULL, mine == ref || mine == -1);
            if (uni) wh.select(ref, risk_hist, lane);
            if (!(anyS || anyI)) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int8_t s = byte_of(w[j], k);
                if (s == 0) {
                    acc.select(nd[k], flush);
                    acc.ci[0] += 1;
                    acc.cl[0] += to_fx((double)rk[k]);
                    if (uni) atomicAdd(&wh.h[risk_bin(rk[k])], 1);
                    else atomicAdd(&risk_hist[(int64_t)nd[k] * LPK_RISK_BINS + risk_bin(rk[k])], 1);
                } else if (s == 2) {
                    acc.select(nd[k], flush);
                    const int stn = strain[base[j] + k];
                    const double v = (double)infectivity[base[j] + k] * srs.v[stn];
                    const long long fx = to_fx(v);
#pragma unroll
                    for (int q = 0; q < LPK_MAX_STRAINS; ++q) acc.cl[1 + q] += (q == stn) ? fx : 0ll;
                }
            }
        }
    }
    wh.flush(risk_hist, lane);
    acc.finish_warp(flush);
}
extern "C" int lpk_tx_step_prep(int32_t num_nodes, int64_t num_people, int32_t n_strains, const int8_t *strains,
                                const double *h_strain_r0_scalars, const int8_t *disease_states,
                                const int16_t *node_ids, const float *daily_infectivity, const float *risks,
                                int64_t *beta_fx, int64_t *exposure_fx, int64_t *sus, int32_t *risk_hist, void *stream) {
    REQUIRE(num_nodes > 0 && num_people >= 0, "tx_step_prep sizes");
    REQUIRE(n_strains >= 1 && n_strains <= LPK_MAX_STRAINS, "tx_step_prep n_strains");
    REQUIRE(strains && h_strain_r0_scalars && disease_states && node_ids && daily_infectivity && risks && beta_fx &&
                exposure_fx && sus && risk_hist, "tx_step_prep null pointer");
    REQUIRE(ALIGNED(disease_states, 4) && ALIGNED(node_ids, 8) && ALIGNED(risks, 16), "tx_step_prep alignment");
    cudaStream_t st = as_stream(stream);
    CUDA_TRY(cudaMemsetAsync(beta_fx, 0, sizeof(int64_t) * num_nodes * n_strains, st), "tx_step_prep memset");
    CUDA_TRY(cudaMemsetAsync(exposure_fx, 0, sizeof(int64_t) * num_nodes, st), "tx_step_prep memset");
    CUDA_TRY(cudaMemsetAsync(sus, 0, sizeof(int64_t) * num_nodes, st), "tx_step_prep memset");
    CUDA_TRY(cudaMemsetAsync(risk_hist, 0, sizeof(int32_t) * num_nodes * LPK_RISK_BINS, st), "tx_step_prep memset");
    if (num_people == 0) return LPK_OK;
    StrainScalars srs;
    for (int s = 0; s < LPK_MAX_STRAINS; ++s) srs.v[s] = s < n_strains ? h_strain_r0_scalars[s] : 0.0;
    k_tx_step_prep<<<agent_grid(num_people, 8), LPK_BLOCK, 0, st>>>(num_people, n_strains, strains, srs, disease_states,
                                                                    node_ids, daily_infectivity, risks, beta_fx,
                                                                    exposure_fx, sus, risk_hist);
    CUDA_TRY(cudaGetLastError(), "lpk_tx_step_prep");
    return LPK_OK;
}

// ------------------------------------------------------------------ C1 count_SEIRP (census)
__global__ void __launch_bounds__(LPK_BLOCK) k_count_seirp(int64_t n, int n_strains, const int16_t *__restrict__ node_id,
                                                            const int8_t *__restrict__ state,
                                                            const int8_t *__restrict__ strain,
                                                            const int8_t *__restrict__ pot_par,
                                                            const int8_t *__restrict__ paralyzed, int32_t *S, int32_t *R,
                                                            int32_t *Ebs, int32_t *Ibs, int32_t *POTP, int32_t *P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TileRange tr = block_tiles(n);
    // ci: 0 S, 1 R, 2 POTP, 3 P, 4.. E by strain, 4+MAX.. I by strain
    NodeAcc<4 + 2 * LPK_MAX_STRAINS, 0> acc;
    acc.init();
    auto flush = [&](int nd, const int *ci, const long long *) {
        red_add(&S[nd], ci[0]); red_add(&R[nd], ci[1]); red_add(&POTP[nd], ci[2]); red_add(&P[nd], ci[3]);
#pragma unroll
        for (int s = 0; s < LPK_MAX_STRAINS; ++s)
            if (s < n_strains) {
                red_add(&Ebs[(int64_t)nd * n_strains + s], ci[4 + s]);
                red_add(&Ibs[(int64_t)nd * n_strains + s], ci[4 + LPK_MAX_STRAINS + s]);
            }
    };
    for (int64_t tile = tr.lo + warp; tile < tr.hi; tile += LPK_WARPS) {
        uint32_t w[4], pp[4], pz[4];
        int64_t base[4];
        int valid[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            base[j] = quad_base(tile, j, lane);
            valid[j] = quad_valid(base[j], n);
            w[j] = valid[j] ? load_b4(state, base[j], valid[j]) : 0xFFFFFFFFu;
            pp[j] = valid[j] ? load_b4(pot_par, base[j], valid[j], 0) : 0u;
            pz[j] = valid[j] ? load_b4(paralyzed, base[j], valid[j], 0) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if ((w[j] & 0x80808080u) == 0x80808080u) continue;
            int nd[4];
            load_s4(node_id, base[j], valid[j], nd);
            const bool anyEI = any_byte_eq(w[j], 1u) || any_byte_eq(w[j], 2u);
            const uint32_t sw = anyEI ? load_b4(strain, base[j], valid[j], 0) : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int8_t s = byte_of(w[j], k);
                if (s < 0) continue;
                acc.select(nd[k], flush);
                acc.ci[0] += (s == 0);
                acc.ci[1] += (s == 3);
                acc.ci[2] += (byte_of(pp[j], k) == 1);
                acc.ci[3] += (byte_of(pz[j], k) == 1);
                if (s == 1 || s == 2) {
                    const int stn = byte_of(sw, k);
#pragma unroll
                    for (int q = 0; q < LPK_MAX_STRAINS; ++q) {
                        acc.ci[4 + q] += (s == 1 && q == stn);
                        acc.ci[4 + LPK_MAX_STRAINS + q] += (s == 2 && q == stn);
                    }
                }
            }
        }
    }
    acc.finish_warp(flush);
}
__global__ void k_sum_strains(int n_nodes, int n_strains, const int32_t *__restrict__ Ebs, const int32_t *__restrict__ Ibs,
                              int32_t *E, int32_t *I) {
    const int nd = blockIdx.x * blockDim.x + threadIdx.x;
    if (nd >= n_nodes) return;
    int e = 0, i = 0;
    for (int s = 0; s < n_strains; ++s) { e += Ebs[(int64_t)nd * n_strains + s]; i += Ibs[(int64_t)nd * n_strains + s]; }
    E[nd] = e; I[nd] = i;
}
extern "C" int lpk_count_seirp(const int16_t *node_id, const int8_t *disease_state, const int8_t *strain,
                               const int8_t *potentially_paralyzed, const int8_t *paralyzed, int32_t n_nodes,
                               int32_t n_strains, int64_t n_people, int32_t *S, int32_t *E, int32_t *I, int32_t *R,
                               int32_t *E_by_strain, int32_t *I_by_strain, int32_t *POTP, int32_t *P, void *stream) {
    REQUIRE(n_nodes > 0 && n_people >= 0, "count_seirp sizes");
    REQUIRE(n_strains >= 1 && n_strains <= LPK_MAX_STRAINS, "count_seirp n_strains");
    REQUIRE(node_id && disease_state && strain && potentially_paralyzed && paralyzed && S && E && I && R && E_by_strain &&
                I_by_strain && POTP && P, "count_seirp null pointer");
    REQUIRE(ALIGNED(disease_state, 4) && ALIGNED(strain, 4) && ALIGNED(potentially_paralyzed, 4) && ALIGNED(paralyzed, 4) &&
                ALIGNED(node_id, 8), "count_seirp alignment");
    cudaStream_t st = as_stream(stream);
    const size_t nb = sizeof(int32_t) * n_nodes;
    CUDA_TRY(cudaMemsetAsync(S, 0, nb, st), "count_seirp memset");
    CUDA_TRY(cudaMemsetAsync(R, 0, nb, st), "count_seirp memset");
    CUDA_TRY(cudaMemsetAsync(POTP, 0, nb, st), "count_seirp memset");
    CUDA_TRY(cudaMemsetAsync(P, 0, nb, st), "count_seirp memset");
    CUDA_TRY(cudaMemsetAsync(E_by_strain, 0, nb * n_strains, st), "count_seirp memset");
    CUDA_TRY(cudaMemsetAsync(I_by_strain, 0, nb * n_strains, st), "count_seirp memset");
    if (n_people > 0) {
        k_count_seirp<<<agent_grid(n_people, 8), LPK_BLOCK, 0, st>>>(n_people, n_strains, node_id, disease_state, strain,
                                                                     potentially_paralyzed, paralyzed, S, R, E_by_strain,
                                                                     I_by_strain, POTP, P);
        CUDA_TRY(cudaGetLastError(), "lpk_count_seirp");
    }
    k_sum_strains<<<(n_nodes + 127) / 128, 128, 0, st>>>(n_nodes, n_strains, E_by_strain, I_by_strain, E, I);
    CUDA_TRY(cudaGetLastError(), "lpk_count_seirp sum");
    return LPK_OK;
}

// ------------------------------------------------------------------ T3 tx_infect (per-agent Bernoulli)
__global__ void __launch_bounds__(LPK_BLOCK) k_tx_infect(int64_t n, int n_strains, const int16_t *__restrict__ node_id,
                                                          int8_t *strain, int8_t *__restrict__ state,
                                                          const float *__restrict__ risk, const float *__restrict__ q,
                                                          const double *__restrict__ strain_cdf, int32_t *n_new,
                                                          DevRng rng) {  // q = tau[nodes]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TileRange tr = block_tiles(n);
    NodeAcc<LPK_MAX_STRAINS, 0> acc;
    acc.init();
    auto flush = [&](int nd, const int *ci, const long long *) {
#pragma unroll
        for (int s = 0; s < LPK_MAX_STRAINS; ++s)
            if (s < n_strains) red_add(&n_new[(int64_t)nd * n_strains + s], ci[s]);
    };
    for (int64_t tile = tr.lo + warp; tile < tr.hi; tile += LPK_WARPS) {
        uint32_t w[4];
        int64_t base[4];
        int valid[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            base[j] = quad_base(tile, j, lane);
            valid[j] = quad_valid(base[j], n);
            w[j] = valid[j] ? load_b4(state, base[j], valid[j]) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!any_byte_eq(w[j], 0u)) continue;  // no susceptible in the quad
            int nd[4];
            load_s4(node_id, base[j], valid[j], nd);
            float qn[4];
            bool live = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                qn[k] = (byte_of(w[j], k) == 0) ? __ldg(&q[nd[k]]) : 0.f;
                live |= qn[k] > 0.f;
            }
            if (!live) continue;  // no force of infection on this quad's nodes: risk is never read
            float rk[4];
            load_f4(risk, base[j], valid[j], rk);
            uint32_t x[4];
            if (rng.x) { for (int k = 0; k < 4; ++k) x[k] = (k < valid[j]) ? rng.x[base[j] + k] : 0u; }
            else expose_words_quad(rng.seed, (uint64_t)base[j] + rng.id_base, rng.tick, x);
            uint32_t nw = w[j];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!(qn[k] > 0.f)) continue;
                if (!expose_test(p_expose(__fmul_rn(rk[k], qn[k])), x[k])) continue;
                const int64_t i = base[j] + k;
                double r;
                if (rng.u2) r = rng.u2[i];
                else { uint32_t y[4]; philox_agent(rng.seed, (uint64_t)i + rng.id_base, rng.tick, LPK_STAGE_STRAIN, y); r = u53(y[0], y[1]); }
                int assigned = 0;
                for (int s = 0; s < n_strains; ++s)
                    if (r < strain_cdf[(int64_t)nd[k] * n_strains + s]) { assigned = s; break; }
                nw = set_byte(nw, k, 1);
                strain[i] = (int8_t)assigned;
                acc.select(nd[k], flush);
#pragma unroll
                for (int s = 0; s < LPK_MAX_STRAINS; ++s) acc.ci[s] += (s == assigned);
            }
            if (nw != w[j]) store_b4(state, base[j], valid[j], nw);
        }
    }
    acc.finish_warp(flush);
}
extern "C" int lpk_tx_infect(int32_t num_nodes, int64_t num_people, int32_t num_strains, const int16_t *node_ids,
                             int8_t *strain, int8_t *disease_state, const float *risks, const float *q,
                             const double *strain_cdf, int32_t *n_new, const lpk_rng *rng, void *stream) {
    REQUIRE(num_nodes > 0 && num_people >= 0, "tx_infect sizes");
    REQUIRE(num_strains >= 1 && num_strains <= LPK_MAX_STRAINS, "tx_infect n_strains");
    REQUIRE(node_ids && strain && disease_state && risks && q && strain_cdf && n_new, "tx_infect null pointer");
    REQUIRE(ALIGNED(disease_state, 4) && ALIGNED(node_ids, 8) && ALIGNED(risks, 16), "tx_infect alignment");
    REQUIRE(!rng || (rng->id_base & 255) == 0, "tx_infect id_base must be a multiple of 256 (exposure draws are shared by aligned groups of 256 agents)");
    cudaStream_t st = as_stream(stream);
    CUDA_TRY(cudaMemsetAsync(n_new, 0, sizeof(int32_t) * num_nodes * num_strains, st), "tx_infect memset");
    if (num_people == 0) return LPK_OK;
    k_tx_infect<<<agent_grid(num_people, 8), LPK_BLOCK, 0, st>>>(num_people, num_strains, node_ids, strain, disease_state,
                                                                 risks, q, strain_cdf, n_new, dev_rng(rng));
    CUDA_TRY(cudaGetLastError(), "lpk_tx_infect");
    return LPK_OK;
}

// ------------------------------------------------------------------ T2 node-level math
// rowsum[i] = sum_j W[i, j]: one warp per row, coalesced.
__global__ void k_row_sums(int n, const double *__restrict__ W, double *__restrict__ rowsum) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    double s = 0.0;
    for (int j = threadIdx.x & 31; j < n; j += 32) s += W[(int64_t)row * n + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(LPK_FULL, s, o);
    if ((threadIdx.x & 31) == 0) rowsum[row] = s;
}

__device__ double node_uniform(uint64_t seed, uint32_t node, uint32_t k, uint32_t tick, int pair) {
    uint32_t x[4];
    philox4x32_10(node, k, tick, LPK_STAGE_NODE, (uint32_t)seed, (uint32_t)(seed >> 32), x);
    return pair ? u53(x[2], x[3]) : u53(x[0], x[1]);
}
// zero-inflated gamma multiplier of an importation-only node (see lpk.h T2); Marsaglia-Tsang, r >= 1
__device__ double importation_multiplier(uint64_t seed, uint32_t node, uint32_t tick, double zi, double r) {
    if (zi >= 1.0) return 0.0;
    if (node_uniform(seed, node, 0, tick, 0) < zi) return 0.0;
    const double d = r - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    uint32_t k = 1;
    for (;;) {
        const double u1 = node_uniform(seed, node, k, tick, 0), u2 = node_uniform(seed, node, k, tick, 1);
        const double x = sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * 3.14159265358979323846 * u2);
        ++k;
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = node_uniform(seed, node, k, tick, 0);
        ++k;
        if (log(1.0 - u) < 0.5 * x * x + d - d * v + d * log(v)) return (d * v) / r / (1.0 - zi);
    }
}

// ---- how many exposures a node should realise: matching the reference's count law ---------------------------------
// Reference (model.py:1368-1407, 1096-1122): K ~ Poisson(E) (ZINB for importation-only nodes = Poisson mixed over a
// zero-inflated gamma, drawn by importation_multiplier), and exactly min(K, S) distinct susceptibles are picked.  The device
// runs independent per-agent trials with probabilities p_i = 1 - exp(-w_i tau), whose count is Poisson-binomial: mean
// sum p_i, variance V_own = sum p_i (1 - p_i).  tau is therefore solved for  T = E[min(K, S)]  (not for E: a node with one
// susceptible and E = 0.5 is exposed with probability 1 - exp(-0.5), as in the reference, and T -> S smoothly as E passes
// S), and the variance the independent trials lack against Var[min(K, S)] is supplied by one more unit-mean gamma
// multiplier g2 on E (delta method: Var[T(g2 E)] ~ (dT/dE)^2 E^2 CV^2, dT/dE = P(K < S)):
//     CV^2 = min((Var[min(K, S)] - V_own) / (P(K < S) E)^2, 1 / (S_eff - 1)) P(K < S)^2,   S_eff = (sum w)^2 / sum w^2.
// Far from saturation this is exactly 1 / S_eff-ish and restores the Poisson variance; as the node saturates the response
// T(E) becomes strongly concave, so the top-up fades with P(K < S)^2 and the remaining concavity is compensated to second
// order (T'' = -P(K = S - 1)) to keep the mean at T(E).  tests/test_count_law.py holds mean and variance to the reference's.
struct PoisMin {
    double T, Pless, V, Plast;  // E[min(K, S)], P(K < S), Var[min(K, S)], P(K = S - 1)
};
// warp-collective; S integer-valued >= 1, e > 0.  Sums over the 24-sigma window below S in 32 chunks (recurrence per lane).
__device__ PoisMin poisson_min_moments(double e, double S, int lane) {
    PoisMin m;
    const double sd = sqrt(e);
    m.Plast = 0.0;
    if (S > e + 12.0 * sd + 12.0) { m.T = e; m.Pless = 1.0; m.V = e; return m; }
    if (S < e - 12.0 * sd - 12.0) { m.T = S; m.Pless = 0.0; m.V = 0.0; return m; }
    const double lo = floor(e - 12.0 * sd - 12.0);
    const long long k0 = lo > 0.0 ? (long long)lo : 0ll, k1 = (long long)S;  // k in [k0, k1)
    const long long n = k1 - k0, chunk = (n + 31) / 32;
    const long long ka = k0 + lane * chunk, kb = (ka + chunk < k1) ? ka + chunk : k1;
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;  // sums of pmf, (S - k) pmf, (S - k)^2 pmf; pmf(S - 1)
    if (ka < kb) {
        double p = exp((double)ka * log(e) - e - lgamma((double)ka + 1.0));
        for (long long k = ka; k < kb; ++k) {
            const double d = S - (double)k;
            b0 += p; b1 += d * p; b2 += d * d * p;
            if (k == k1 - 1) b3 = p;
            p *= e / (double)(k + 1);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        b0 += __shfl_xor_sync(LPK_FULL, b0, o); b1 += __shfl_xor_sync(LPK_FULL, b1, o); b2 += __shfl_xor_sync(LPK_FULL, b2, o);
        b3 += __shfl_xor_sync(LPK_FULL, b3, o);
    }
    b0 = __shfl_sync(LPK_FULL, b0, 0); b1 = __shfl_sync(LPK_FULL, b1, 0); b2 = __shfl_sync(LPK_FULL, b2, 0);
    m.Plast = __shfl_sync(LPK_FULL, b3, 0);
    m.T = S - b1;                       // E[min] = S - E[(S - K)+]
    m.Pless = b0 < 1.0 ? b0 : 1.0;
    m.V = fmax(b2 - b1 * b1, 0.0);      // Var[min] = Var[(S - K)+]
    return m;
}
__device__ double node_uniform_var(uint64_t seed, uint32_t node, uint32_t k, uint32_t tick, int pair) {
    uint32_t x[4];
    philox4x32_10(node, k, tick, LPK_STAGE_NODE_VAR, (uint32_t)seed, (uint32_t)(seed >> 32), x);
    return pair ? u53(x[2], x[3]) : u53(x[0], x[1]);
}
// Gamma(shape, scale = 1 / shape): unit mean, CV^2 = 1 / shape.  Marsaglia-Tsang; shape < 1 through Gamma(shape + 1) U^(1/shape).
__device__ double unit_gamma(uint64_t seed, uint32_t node, uint32_t tick, double shape) {
    const double a = shape < 1.0 ? shape + 1.0 : shape;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    uint32_t k = 1;
    double g;
    for (;;) {
        const double u1 = node_uniform_var(seed, node, k, tick, 0), u2 = node_uniform_var(seed, node, k, tick, 1);
        const double x = sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * 3.14159265358979323846 * u2);
        ++k;
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = node_uniform_var(seed, node, k, tick, 0);
        ++k;
        if (log(1.0 - u) < 0.5 * x * x + d - d * v + d * log(v)) { g = d * v; break; }
    }
    if (shape < 1.0) g *= pow(1.0 - node_uniform_var(seed, node, 0, tick, 0), 1.0 / shape);
    return g / shape;
}

// tau[j]: sum over the node's susceptibles of (1 - exp(-risk * tau)) = T, on the risk histogram.
// Successive weighted sampling without replacement of K agents (the reference, model.py:1096-1122) selects agent i with
// probability 1 - exp(-w_i tau), tau fixed by the count; solving for the EXPECTED count gives independent per-agent
// trials with the same marginals and the same node mean.  One warp per node, 6 bins per lane; Newton from the left of a
// concave increasing function converges monotonically.  The bins' representative weights (bin centres) are rescaled so that
// their sum equals the exact risk sum of the node's susceptibles (expo): a table whose risks are all equal (no individual
// heterogeneity) sits on a bin edge, and the centre would be 6 % off.  E: expected exposures x importation multiplier.
#define TAU_BINS (LPK_RISK_BINS / 32)
// t: a point at or left of the root, e.g. tau_lower_bound
__device__ __forceinline__ double tau_lower_bound(double S, double Wsum, double T) {
    return -log1p(-T / S) * S / Wsum;  // the equal-weights solution: a lower bound of the root (Jensen), tight for small T
}
__device__ __forceinline__ double newton_tau(const double (&h)[TAU_BINS], const double (&w)[TAU_BINS], double T, double t, double tol) {
    for (int it = 0; it < 100; ++it) {
        double F = 0.0, dF = 0.0;
#pragma unroll
        for (int i = 0; i < TAU_BINS; ++i) {
            const double em = expm1(-w[i] * t);
            F -= h[i] * em;
            dF += h[i] * w[i] * (em + 1.0);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { F += __shfl_xor_sync(LPK_FULL, F, o); dF += __shfl_xor_sync(LPK_FULL, dF, o); }
        F = __shfl_sync(LPK_FULL, F, 0);
        dF = __shfl_sync(LPK_FULL, dF, 0);
        const double step = (T - F) / dF;
        if (!(step > 0.0)) break;
        t += step;
        if (step <= tol * t) break;
    }
    return t;
}
__device__ float solve_tau_warp(int j, const int32_t *__restrict__ hist, double E, double expo, uint64_t seed, uint32_t tick, int lane) {
    double h[TAU_BINS], w[TAU_BINS];
    double S = 0.0, Wsum = 0.0;
#pragma unroll
    for (int i = 0; i < TAU_BINS; ++i) {
        const int b = lane + 32 * i;
        h[i] = (double)hist[(int64_t)j * LPK_RISK_BINS + b];
        w[i] = risk_bin_weight(b);
        S += h[i];
        Wsum += h[i] * w[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { S += __shfl_xor_sync(LPK_FULL, S, o); Wsum += __shfl_xor_sync(LPK_FULL, Wsum, o); }
    S = __shfl_sync(LPK_FULL, S, 0);
    Wsum = __shfl_sync(LPK_FULL, Wsum, 0);
    if (!(E > 0.0) || !(S > 0.0) || !(Wsum > 0.0)) return 0.f;
    const double c = expo > 0.0 ? expo / Wsum : 1.0;  // first moment of the histogram := the exact risk sum
    double W1 = 0.0, W2 = 0.0;
#pragma unroll
    for (int i = 0; i < TAU_BINS; ++i) { w[i] *= c; W1 += h[i] * w[i]; W2 += h[i] * w[i] * w[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { W1 += __shfl_xor_sync(LPK_FULL, W1, o); W2 += __shfl_xor_sync(LPK_FULL, W2, o); }
    Wsum = __shfl_sync(LPK_FULL, W1, 0);
    W2 = __shfl_sync(LPK_FULL, W2, 0);
    const double Seff = Wsum * Wsum / W2;
    const PoisMin m = poisson_min_moments(E, S, lane);
    double T = m.T;
    if (T >= S * (1.0 - 1e-12)) return 3.0e38f;  // the reference's size = min(K, S) with K >= S: everybody
    if (!(T > 0.0)) return 0.f;
    double t = newton_tau(h, w, T, tau_lower_bound(S, Wsum, T), 1e-7);  // enough for V_own; refined below if it is final
    bool redo = false;
    const double slope = m.Pless * E;
    if (Seff > 1.0 + 1e-9 && slope > 0.0) {
        double vown = 0.0;
#pragma unroll
        for (int i = 0; i < TAU_BINS; ++i) { const double pb = -expm1(-w[i] * t); vown += h[i] * pb * (1.0 - pb); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vown += __shfl_xor_sync(LPK_FULL, vown, o);
        vown = __shfl_sync(LPK_FULL, vown, 0);
        const double v = fmin((m.V - vown) / (slope * slope), 1.0 / (Seff - 1.0)) * m.Pless * m.Pless;
        if (v > 1e-9) {
            double g2 = 0.0;
            if (lane == 0) g2 = unit_gamma(seed, (uint32_t)j, tick, 1.0 / v);
            g2 = __shfl_sync(LPK_FULL, g2, 0);
            const double E2 = E * (1.0 + 0.5 * m.Plast * E * v / m.Pless) * g2;
            T = E2 > 0.0 ? poisson_min_moments(E2, S, lane).T : 0.0;
            if (T >= S * (1.0 - 1e-12)) return 3.0e38f;
            if (!(T > 0.0)) return 0.f;
            redo = true;
        }
    }
    t = newton_tau(h, w, T, redo ? tau_lower_bound(S, Wsum, T) : t, 1e-11);
    return (float)(t > 3.0e38 ? 3.0e38 : t);
}

// Sharded runs (lpk_xchg): the gathered tally in beta_fx is complete once every rank's flag has reached `seq`.  One thread
// per block acquires the flags at system scope (the peers' stores arrive over NVLink), the block barrier publishes that to
// the other threads.  A peer that never arrives turns into a trap after ~10 s instead of a hung device.
__device__ __forceinline__ void xchg_wait(const uint32_t *flags, int world, uint32_t seq) {
    if (flags) {
        if (threadIdx.x == 0) {
            for (int r = 0; r < world; ++r) {
                uint32_t v, spins = 0;
                for (;;) {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
                    if ((int32_t)(v - seq) >= 0) break;
                    if (++spins > (1u << 24)) __trap();
                    __nanosleep(64);
                }
            }
        }
        __syncthreads();
    }
}

// One block = 32 destination nodes x 32 source slices (1024 threads).  The infectivity tally is staged through shared
// memory 1024 source rows at a time (the previous version re-read it from global in a dependent loop and took 41 us at 774
// nodes: profiles/r1_v18_launches.csv); the block's 32 warps then solve tau for its 32 nodes.
#define NM_ROWS 1024
// Many nodes (continental config; 8 x 774 in the weak-scaling bench): one block per 32 destination nodes walking every source
// row leaves most SMs idle (25 blocks for 38 MB of network at 6192 nodes, 50 us).  The transfer is then summed per chunk of
// 1024 source rows by a 2-D grid (destinations x chunks) and k_tx_node_math adds the chunks in index order, so the result
// does not depend on which block finishes first.
__global__ void __launch_bounds__(1024) k_node_matvec_partial(int n, int n_strains, const int64_t *__restrict__ beta_fx,
                                                               const double *__restrict__ W, int j_lo, int j_hi,
                                                               double *__restrict__ partial, const uint32_t *xflags, int xworld,
                                                               uint32_t xseq) {
    __shared__ double sbeta[LPK_MAX_STRAINS][NM_ROWS];
    __shared__ unsigned char snz[NM_ROWS];
    xchg_wait(xflags, xworld, xseq);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = j_lo + blockIdx.x * 32 + tx, base = blockIdx.y * NM_ROWS;
    const int rows = min(NM_ROWS, n - base);
    if ((int)threadIdx.x < rows) {
        bool nz = false;
#pragma unroll
        for (int s = 0; s < LPK_MAX_STRAINS; ++s) {
            const long long b = (s < n_strains) ? __ldcg(&beta_fx[(int64_t)(base + threadIdx.x) * n_strains + s]) : 0;
            nz |= (b != 0);
            sbeta[s][threadIdx.x] = (double)b / LPK_FX_SCALE;
        }
        snz[threadIdx.x] = nz ? 1 : 0;
    }
    __syncthreads();
    double in[LPK_MAX_STRAINS] = {0.0, 0.0, 0.0, 0.0};
    if (j < j_hi) {
#pragma unroll 4
        for (int i = ty; i < rows; i += 32) {
            if (!snz[i]) continue;
            const double w = W[(int64_t)(base + i) * n + j];
#pragma unroll
            for (int s = 0; s < LPK_MAX_STRAINS; ++s) in[s] += sbeta[s][i] * w;
        }
    }
    __syncthreads();
    double *part = &sbeta[0][0];
#pragma unroll
    for (int s = 0; s < LPK_MAX_STRAINS; ++s) part[(ty * LPK_MAX_STRAINS + s) * 32 + tx] = in[s];
    __syncthreads();
    if (ty == 0 && j < j_hi) {
#pragma unroll
        for (int s = 0; s < LPK_MAX_STRAINS; ++s) {
            double acc = 0.0;
            for (int y = 0; y < 32; ++y) acc += part[(y * LPK_MAX_STRAINS + s) * 32 + tx];
            partial[((int64_t)blockIdx.y * LPK_MAX_STRAINS + s) * n + j] = acc;
        }
    }
}
// One block = NM_NPB destination nodes x 32 source slices.  With 32 nodes per block (round 1) the 774 nodes of Nigeria made 25
// blocks on 25 SMs, each running 32 double-precision Newton solves; eight nodes per block spread the solves over 97 SMs (node
// step of a day at 2.75e7 agents: 41 -> 32 us).  The solve itself is one dependent chain of ~5000 instructions per warp
// (profiles/r2_nodemath_*): latency, not throughput.
#define NM_NPB 8
#define NM_THREADS (NM_NPB * 32)
__global__ void __launch_bounds__(NM_THREADS) k_tx_node_math(int n, int n_strains, const int64_t *__restrict__ beta_fx,
                                                              const int64_t *__restrict__ exposure_fx,
                                                              const double *__restrict__ W, const double *__restrict__ rowsum,
                                                              double season, const double *__restrict__ r0_scalars,
                                                              const int32_t *__restrict__ alive, double zi, double disp_r,
                                                              double *target, double *strain_cdf, double *prob, double *expected,
                                                              const int32_t *__restrict__ hist, float *__restrict__ tau, uint64_t seed,
                                                              uint32_t tick, int j_lo, int j_hi, const double *__restrict__ partial,
                                                              int n_chunks, const uint32_t *xflags, int xworld, uint32_t xseq,
                                                              const __grid_constant__ lpk_node_args ep, int run_ep) {
    __shared__ double sbeta[LPK_MAX_STRAINS][NM_ROWS];  // 32 KB; its head is reused as part[32 slices][strains][NM_NPB nodes]
    __shared__ unsigned char snz[NM_ROWS];
    __shared__ double stgt[NM_NPB];
    xchg_wait(xflags, xworld, xseq);
    const int tx = threadIdx.x % NM_NPB, ty = threadIdx.x / NM_NPB;  // node of the block, slice of the source rows (0 .. 31)
    const int j = j_lo + blockIdx.x * NM_NPB + tx;  // destination nodes [j_lo, j_hi): all of them, or this rank's shard
    double in[LPK_MAX_STRAINS] = {0.0, 0.0, 0.0, 0.0};
    if (partial) {  // the transfer was summed per chunk of source rows by k_node_matvec_partial: add the chunks in order
        if (ty == 0 && j < j_hi) {
            for (int c = 0; c < n_chunks; ++c)
#pragma unroll
                for (int s = 0; s < LPK_MAX_STRAINS; ++s) in[s] += partial[((int64_t)c * LPK_MAX_STRAINS + s) * n + j];
        }
    } else
    for (int base = 0; base < n; base += NM_ROWS) {
        const int rows = min(NM_ROWS, n - base);
        __syncthreads();
        for (int r = threadIdx.x; r < rows; r += NM_THREADS) {
            bool nz = false;
#pragma unroll
            for (int s = 0; s < LPK_MAX_STRAINS; ++s) {
                const long long b = (s < n_strains) ? __ldcg(&beta_fx[(int64_t)(base + r) * n_strains + s]) : 0;
                nz |= (b != 0);
                sbeta[s][r] = (double)b / LPK_FX_SCALE;
            }
            snz[r] = nz ? 1 : 0;
        }
        __syncthreads();
        if (j < j_hi) {
#pragma unroll 4
            for (int i = ty; i < rows; i += 32) {
                if (!snz[i]) continue;  // rows without infectivity contribute nothing
                const double w = W[(int64_t)(base + i) * n + j];
#pragma unroll
                for (int s = 0; s < LPK_MAX_STRAINS; ++s) in[s] += sbeta[s][i] * w;
            }
        }
    }
    __syncthreads();
    double *part = &sbeta[0][0];  // [ty][s][tx]
#pragma unroll
    for (int s = 0; s < LPK_MAX_STRAINS; ++s) part[(ty * LPK_MAX_STRAINS + s) * NM_NPB + tx] = in[s];
    __syncthreads();
    if (ty == 0) {
        double tgt = 0.0;
        if (j < j_hi) {
            if (run_ep) epilogue_node(ep, j, j_lo);  // the tick's bookkeeping of node j (writes results.pop[t][j], read just below)
            double P = 0.0, local = 0.0, p[LPK_MAX_STRAINS];
            const double popn = fmax((double)alive[j], 1.0);
            for (int s = 0; s < n_strains; ++s) {
                double inc = 0.0;
                for (int y = 0; y < 32; ++y) inc += part[(y * LPK_MAX_STRAINS + s) * NM_NPB + tx];
                const double pre = (double)__ldcg(&beta_fx[(int64_t)j * n_strains + s]) / LPK_FX_SCALE;
                local += pre;
                double b = pre + inc - pre * rowsum[j];
                b = b * season * r0_scalars[j];
                const double rate = b / popn;
                p[s] = fmax(1.0 - exp(-rate), 0.0);
                prob[(int64_t)j * n_strains + s] = p[s];
                P += p[s];
            }
            double run = 0.0;
            for (int s = 0; s < n_strains; ++s) {
                run += (P > 0.0) ? p[s] / P : 0.0;
                strain_cdf[(int64_t)j * n_strains + s] = run;
            }
            const double e = ((double)exposure_fx[j] / LPK_FX_SCALE) * P;  // model.py:1363
            expected[j] = e;
            double g = 1.0;
            if (local == 0.0 && P > 0.0) g = importation_multiplier(seed, (uint32_t)j, tick, zi, disp_r);
            tgt = e * g;  // expected number of exposures the node's susceptibles must realise this tick
            target[j] = tgt;
        }
        stgt[tx] = tgt;
    }
    __syncthreads();
    const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jn = j_lo + blockIdx.x * NM_NPB + wq;  // warp wq solves node jn
    if (jn < j_hi) {
        const float t = solve_tau_warp(jn, hist, stgt[wq], (double)exposure_fx[jn] / LPK_FX_SCALE, seed, tick, lane);
        if (lane == 0) tau[jn] = t;
    }
}

int lpk_launch_node_math(int32_t num_nodes, int32_t n_strains, const int64_t *beta_fx, const int64_t *exposure_fx,
                         const int32_t *risk_hist, const double *network, double beta_seasonality, const double *r0_scalars,
                         const int32_t *alive_counts, double zero_inflation, double dispersion, float *tau, double *strain_cdf,
                         double *prob, double *expected, double *ws, uint64_t seed, uint32_t tick, cudaStream_t st, bool rowsums_done,
                         int32_t node_lo, int32_t node_hi, const uint32_t *xchg_flags, int32_t xchg_world, uint32_t xchg_seq,
                         const lpk_node_args *inline_epilogue, double *matvec_ws) {
    double *rowsum = ws, *target = ws + num_nodes;
    if (!rowsums_done) {
        k_row_sums<<<(num_nodes + 7) / 8, 256, 0, st>>>(num_nodes, network, rowsum);
        CUDA_TRY(cudaGetLastError(), "node_math rowsums");
    }
    double r = nearbyint(dispersion);
    if (r < 1.0) r = 1.0;
    const double *partial = nullptr;
    const int n_chunks = (num_nodes + NM_ROWS - 1) / NM_ROWS;
    static int split = -1;  // LPK_NODE_SPLIT=0: experiments only (single-block form for any size)
    if (split < 0) { const char *e = getenv("LPK_NODE_SPLIT"); split = (e && e[0] == '0') ? 0 : 1; }
    if (n_chunks > 1 && split && matvec_ws) {  // caller-owned scratch (lpk_node_args.matvec_ws)
        k_node_matvec_partial<<<dim3((node_hi - node_lo + 31) / 32, n_chunks), 1024, 0, st>>>(num_nodes, n_strains, beta_fx, network, node_lo,
                                                                                             node_hi, matvec_ws, xchg_flags, xchg_world, xchg_seq);
        CUDA_TRY(cudaGetLastError(), "node_math matvec");
        partial = matvec_ws;
    } else if (n_chunks > 1 && split) {  // library-owned scratch, one per device, grown on demand (chunks x strains x nodes doubles: 1.4 MB at 6192 nodes)
        static double *scratch[64] = {nullptr};
        static size_t scratch_bytes[64] = {0};
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev), "node_math device");
        REQUIRE(dev >= 0 && dev < 64, "node_math device index");
        const size_t need = (size_t)n_chunks * LPK_MAX_STRAINS * num_nodes * sizeof(double);
        if (scratch_bytes[dev] < need) {
            CUDA_TRY(cudaStreamSynchronize(st), "node_math scratch");
            if (scratch[dev]) cudaFree(scratch[dev]);
            scratch[dev] = nullptr; scratch_bytes[dev] = 0;
            CUDA_TRY(cudaMalloc(&scratch[dev], need), "node_math scratch");
            scratch_bytes[dev] = need;
        }
        k_node_matvec_partial<<<dim3((node_hi - node_lo + 31) / 32, n_chunks), 1024, 0, st>>>(num_nodes, n_strains, beta_fx, network, node_lo,
                                                                                             node_hi, scratch[dev], xchg_flags, xchg_world, xchg_seq);
        CUDA_TRY(cudaGetLastError(), "node_math matvec");
        partial = scratch[dev];
    }
    k_tx_node_math<<<(node_hi - node_lo + NM_NPB - 1) / NM_NPB, NM_THREADS, 0, st>>>(num_nodes, n_strains, beta_fx, exposure_fx, network, rowsum,
                                                                   beta_seasonality, r0_scalars, alive_counts, zero_inflation, r, target,
                                                                   strain_cdf, prob, expected, risk_hist, tau, seed, tick, node_lo, node_hi,
                                                                   partial, n_chunks, xchg_flags, xchg_world, xchg_seq,
                                                                   inline_epilogue ? *inline_epilogue : lpk_node_args{}, inline_epilogue ? 1 : 0);
    CUDA_TRY(cudaGetLastError(), "node_math");
    return LPK_OK;
}

extern "C" int lpk_tx_node_math(int32_t num_nodes, int32_t n_strains, const int64_t *beta_fx, const int64_t *exposure_fx,
                                const int32_t *risk_hist, const double *network, double beta_seasonality,
                                const double *r0_scalars, const int32_t *alive_counts, double zero_inflation,
                                double dispersion, float *tau, double *strain_cdf, double *prob, double *expected,
                                double *ws, const lpk_rng *rng, void *stream) {
    REQUIRE(num_nodes > 0, "tx_node_math sizes");
    REQUIRE(n_strains >= 1 && n_strains <= LPK_MAX_STRAINS, "tx_node_math n_strains");
    REQUIRE(beta_fx && exposure_fx && risk_hist && network && r0_scalars && alive_counts && tau && strain_cdf && prob &&
                expected && ws, "tx_node_math null pointer");
    return lpk_launch_node_math(num_nodes, n_strains, beta_fx, exposure_fx, risk_hist, network, beta_seasonality, r0_scalars,
                                alive_counts, zero_inflation, dispersion, tau, strain_cdf, prob, expected, ws,
                                rng ? rng->seed : 0, rng ? rng->tick : 0, as_stream(stream), false, 0, num_nodes, nullptr, 0, 0u, nullptr, nullptr);
}
