// lpk_vital.cu -- births on the device (reference VitalDynamics_ABM.step, model.py:1711-1734).
//
// The reference draws per-node birth counts and newborn lifespans on the host (numpy), appends the cohort with
// LaserFrame.add and writes four column slices.  At 2.2e8 agents that is ~10 ms of host work plus a device sync on
// every vital-dynamics tick -- more than the whole fused day costs -- so here the cohort is created in HBM:
//   k_births_count   per node: expected = step * rate * pop[t-1]; births = floor + Bernoulli(frac) (Philox(node, tick));
//                    exclusive scan -> cohort offsets; count += total (or a status flag when capacity would overflow,
//                    the device analogue of LaserFrame.add raising)
//   k_births_fill    per newborn: slot = old_count + k, node by binary search of the offsets (node-major cohort, like
//                    np.repeat(arange(nodes), births)), date_of_birth = t, date_of_death = t + lifespan (inverse-CDF draw
//                    on the cumulative-deaths table, Philox(agent, tick)), disease_state = 0
//   k_births_tiles   re-derive the tile -> node table for the tiles the cohort touched
// Newborns keep whatever was pre-drawn in their slot for every other column, exactly like the reference (timers, risk,
// infectivity are drawn for the whole capacity at construction; ri_timer stays at its default).
#include "lpk_host.cuh"
#include "lpk_hot.cuh"
#include "lpk_stages.cuh"

__global__ void __launch_bounds__(1024) k_births_count(lpk_births_args a) {
    __shared__ int s_scan[1024];
    __shared__ int s_carry;
    const int tid = threadIdx.x;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < a.n_nodes; base += 1024) {
        const int n = base + tid;
        int b = 0;
        if (n < a.n_nodes) {
            const double expected = a.step_size * a.birth_rate[n] * (double)a.pop_prev[n];
            const int whole = (int)expected;  // truncation, like astype(np.int32)
            const double frac = expected - (double)whole;
            uint32_t x[4];
            philox4x32_10((uint32_t)n, 0u, (uint32_t)a.tick, LPK_STAGE_BIRTH, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), x);
            b = whole + ((u53(x[0], x[1]) < frac) ? 1 : 0);
            if (b < 0) b = 0;
        }
        // inclusive scan of this chunk
        s_scan[tid] = b;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const int v = (tid >= off) ? s_scan[tid - off] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int carry = s_carry;
        if (n < a.n_nodes) {
            a.node_offsets_ws[n] = carry + s_scan[tid] - b;  // exclusive
            a.births_row[n] = b;
        }
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_scan[1023];
        __syncthreads();
    }
    if (tid == 0) {
        int total = s_carry;
        const long long old_count = a.counts[1];
        if (old_count + total > a.capacity || (*a.status & 1)) {  // LaserFrame.add would raise: flag it (sticky: no later
            if (!(*a.status & 1)) a.status[1] = a.tick;                    // cohort is appended either), remember the tick,
            *a.status |= 1;                                                // create nobody
            total = 0;
            for (int n = 0; n < a.n_nodes; ++n) { a.births_row[n] = 0; a.node_offsets_ws[n] = 0; }
        }
        a.node_offsets_ws[a.n_nodes] = total;
        a.cohort_ws[0] = old_count;
        a.cohort_ws[1] = total;
        a.counts[1] = old_count + total;
    }
}

__device__ __forceinline__ int newborn_lifespan(const lpk_births_args &a, uint64_t agent) {
    uint32_t x[4];
    philox_agent(a.seed, agent + a.id_base, (uint32_t)a.tick, LPK_STAGE_LIFESPAN, x);
    const double u1 = u53(x[0], x[1]), u2 = u53(x[2], x[3]);
    const long long total = a.cum_deaths[a.max_year + 1];
    const long long draw = 1 + (long long)floor(u1 * (double)(total > 1 ? total : 1));
    int lo = 0, hi = a.max_year + 2;  // first j with cum_deaths[j] >= draw  (searchsorted side="left")
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a.cum_deaths[mid] < draw) lo = mid + 1; else hi = mid;
    }
    int yod = lo - 1;
    yod = yod < 0 ? 0 : (yod > a.max_year ? a.max_year : yod);
    const int doy = (yod == 0) ? 1 + (int)floor(u2 * 364.0) : (int)floor(u2 * 365.0);
    return yod * 365 + doy;
}

__global__ void k_births_fill(lpk_births_args a) {
    const long long old_count = a.cohort_ws[0], total = a.cohort_ws[1];
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = a.n_nodes;  // last node with offset <= k
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (a.node_offsets_ws[mid] <= k) lo = mid; else hi = mid;
        }
        // skip empty nodes that share the same offset: take the LAST node whose offset <= k and whose count > 0
        while (lo + 1 < a.n_nodes && a.node_offsets_ws[lo + 1] <= k) ++lo;
        const long long slot = old_count + k;
        a.node_id[slot] = (int16_t)lo;
        a.date_of_birth[slot] = a.tick;
        a.date_of_death[slot] = a.tick + newborn_lifespan(a, (uint64_t)slot);
        a.disease_state[slot] = 0;
        if (a.ri_timer && a.ri_newborn_timer >= 0) a.ri_timer[slot] = (int16_t)a.ri_newborn_timer;
        if (a.ri_timer && a.ri_lazy_k) {  // lazy RI countdown (lpk_tick_args.ri_lazy_k): the newborn does not owe the earlier ticks
            a.ri_timer[slot] = (int16_t)(uint16_t)((uint16_t)a.ri_timer[slot] + (uint16_t)(a.ri_lazy_k * a.ri_step));
        }
        if (a.ri_timer && a.pair_ri_max) atomicMax(&a.pair_ri_max[slot >> 8], (int)a.ri_timer[slot]);
        if (a.ri_timer && a.ri_k) {  // today's RI tick, if any, comes after the births: the newborn takes part in it
            const uint8_t j = ri_tick_index(a.ri_timer[slot], a.ri_lazy_k, a.ri_step, a.tick - 1);
            a.ri_k[slot] = (j && a.ri_lazy_k + (int)j <= 254) ? (uint8_t)(a.ri_lazy_k + j) : (uint8_t)0;
        }
        if (a.hot) {  // fused path: the newborn's event record, agenda byte (a susceptible) and its pair's earliest death date
            HotRecU x;
            x.r.state = 0; x.r.strain = a.strain[slot]; x.r.et = a.exposure_timer[slot]; x.r.it = a.infection_timer[slot];
            x.r.pt = a.paralysis_timer[slot]; x.r.pq = a.potentially_paralyzed[slot]; x.r.par = a.paralyzed[slot];
            x.r.ipv = a.ipv_protected[slot];
            a.rec[slot] = x.u;
            bool over = false;
            a.hot[slot] = (uint8_t)(HOT_S | risk_code(a.acq_risk_multiplier[slot], a.risk_e0, &over));
            if (over) *a.status = 2;
            if (a.pair_min_dod) atomicMin(&a.pair_min_dod[slot >> 8], a.date_of_death[slot]);
        }
        if (a.sus) {  // fused path: the susceptible-side tallies are carried incrementally (lpk_tick.cu)
            const float rk = a.acq_risk_multiplier[slot];
            atomicAdd(reinterpret_cast<unsigned long long *>(&a.sus[lo]), 1ull);
            red_add(&a.exposure_fx[lo], __float2ll_rn(rk * 1073741824.0f));
            atomicAdd(&a.risk_hist[(int64_t)lo * LPK_RISK_BINS + risk_bin(rk)], 1);
        }
    }
}

__global__ void k_births_tiles(lpk_births_args a) {
    const long long old_count = a.cohort_ws[0], total = a.cohort_ws[1];
    if (total == 0 || !a.tile_node) return;
    const int lane = threadIdx.x & 31;
    const long long first = old_count / LPK_TILE, last = (old_count + total - 1) / LPK_TILE;
    for (long long tile = first + (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile <= last;
         tile += (long long)gridDim.x * (blockDim.x >> 5)) {
        const long long lo = tile * LPK_TILE;
        const int firstn = a.node_id[lo];
        bool same = true;
        for (int k = lane; k < LPK_TILE; k += 32) {
            const long long i = lo + k;
            same &= (i < a.capacity) && (a.node_id[i] == firstn);
        }
        same = __all_sync(LPK_FULL, same);
        if (lane == 0) a.tile_node[tile] = (same && firstn >= 0) ? firstn : -1;
    }
}

extern "C" int lpk_vd_births(const lpk_births_args *args, void *stream) {
    REQUIRE(args, "vd_births null struct");
    const lpk_births_args &a = *args;
    REQUIRE(a.n_nodes > 0 && a.capacity > 0 && a.max_year > 0, "vd_births sizes");
    REQUIRE(a.birth_rate && a.pop_prev && a.births_row && a.counts && a.cum_deaths && a.node_offsets_ws && a.cohort_ws && a.status,
            "vd_births node-level pointers");
    REQUIRE(a.disease_state && a.node_id && a.date_of_birth && a.date_of_death, "vd_births agent columns");
    REQUIRE(!a.hot || (a.rec && a.acq_risk_multiplier && a.strain && a.exposure_timer && a.infection_timer && a.paralysis_timer &&
                       a.potentially_paralyzed && a.paralyzed && a.ipv_protected), "vd_births fused-path companions");
    cudaStream_t st = as_stream(stream);
    k_births_count<<<1, 1024, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError(), "lpk_vd_births count");
    k_births_fill<<<lpk_sm_count() * 2, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError(), "lpk_vd_births fill");
    if (a.tile_node) {
        k_births_tiles<<<lpk_sm_count(), 256, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError(), "lpk_vd_births tiles");
    }
    return LPK_OK;
}
