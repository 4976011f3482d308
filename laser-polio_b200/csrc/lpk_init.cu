// lpk_init.cu -- the agent table is drawn in HBM.
//
// The reference fills every per-agent column on the host at construction: correlated risk / infectivity through a Gaussian
// copula in 1 M-agent numpy batches with scipy.stats.gamma.ppf (model.py:816-866, its own FIXME "known to be slow"), the three
// disease timers for the whole capacity (model.py:571-587), ages from the pyramid, lifespans from the Kaplan-Meier table,
// routine-immunisation dates (model.py:1578-1596, 1605-1611, 1893-1894) and the chronically missed (model.py:154-159).  At
// 2.2e8 agents that is minutes of host time in front of a simulation that takes two seconds.  Here each column is a pure
// function of (seed, agent id, stage): one Philox4x32-10 stream per agent and stage, so a table drawn on 1 or 8 GPUs, in one
// piece or slice by slice, is the same table, and the CPU checker under tests/ restates it draw for draw.
//
// Samplers (each consumes 53-bit uniforms from the agent's stream, block after block):
//   normal        Box-Muller, cosine branch (the copula uses both branches of one pair)
//   exponential   -scale * log(u),  u in (0, 1]
//   gamma         Marsaglia-Tsang squeeze-free form (shape < 1 through the u^(1/shape) boost)
//   poisson       inversion by sequential search for lam < 30, Hoermann's PTRS otherwise (the algorithm numpy uses)
//   lognormal     exp(mu + sigma * normal);  uniform  min + floor(u * (max - min))  (np.random.randint)
// Float -> int8 follows numpy's assignment cast (truncate toward zero, keep the low byte) BEFORE the clip to [0, 127], as the
// reference does.
#include "lpk_host.cuh"

namespace {

struct Stream {  // the 53-bit uniforms of Philox(seed; id, block, stage), block = 0, 1, ...
    uint64_t seed, id;
    uint32_t stage, blk;
    uint32_t x[4];
    int pos;
    __device__ __forceinline__ Stream(uint64_t seed_, uint64_t id_, uint32_t stage_) : seed(seed_), id(id_), stage(stage_), blk(0u), pos(4) {}
    __device__ __forceinline__ void refill() {
        philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), blk++, stage, (uint32_t)seed, (uint32_t)(seed >> 32), x);
        pos = 0;
    }
    __device__ __forceinline__ uint64_t bits53() {
        if (pos >= 4) refill();
        const uint64_t v = (((uint64_t)x[pos] << 32) | x[pos + 1]) >> 11;
        pos += 2;
        return v;
    }
    __device__ __forceinline__ double u() { return (double)bits53() * (1.0 / 9007199254740992.0); }          // [0, 1)
    __device__ __forceinline__ double uo() { return (double)(bits53() + 1ull) * (1.0 / 9007199254740992.0); }  // (0, 1]
};

__device__ __forceinline__ double draw_normal(Stream &s) {
    const double r = sqrt(__dmul_rn(-2.0, log(s.uo())));
    return __dmul_rn(r, cos(__dmul_rn(6.283185307179586, s.u())));
}
__device__ double draw_gamma(Stream &s, double shape, double scale) {
    double boost = 1.0;
    if (shape < 1.0) {
        boost = pow(s.uo(), 1.0 / shape);
        shape += 1.0;
    }
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(__dmul_rn(9.0, d));
    for (int it = 0; it < 64; ++it) {
        const double z = draw_normal(s);
        const double w = __dadd_rn(1.0, __dmul_rn(c, z));
        const double uu = s.uo();
        if (w <= 0.0) continue;
        const double v = __dmul_rn(__dmul_rn(w, w), w);
        const double lhs = log(uu);
        const double rhs = __dadd_rn(__dadd_rn(__dmul_rn(0.5, __dmul_rn(z, z)), d), __dadd_rn(__dmul_rn(-d, v), __dmul_rn(d, log(v))));
        if (lhs < rhs) return __dmul_rn(__dmul_rn(__dmul_rn(d, v), scale), boost);
    }
    return __dmul_rn(__dmul_rn(d, scale), boost);  // 64 rejections in a row: probability < 1e-80
}
__device__ double draw_poisson(Stream &s, double lam) {
    if (!(lam > 0.0)) return 0.0;
    if (lam < 30.0) {
        const double u = s.u();
        double p = exp(-lam), cum = p;
        int k = 0;
        while (u >= cum && k < 1000) {
            ++k;
            p = __ddiv_rn(__dmul_rn(p, lam), (double)k);
            cum = __dadd_rn(cum, p);
        }
        return (double)k;
    }
    const double slam = sqrt(lam), loglam = log(lam), b = __dadd_rn(0.931, __dmul_rn(2.53, slam)), a = __dadd_rn(-0.059, __dmul_rn(0.02483, b));
    const double invalpha = __dadd_rn(1.1239, __ddiv_rn(1.1328, b - 3.4)), vr = __dadd_rn(0.9277, -__ddiv_rn(3.6224, b - 2.0));
    for (int it = 0; it < 256; ++it) {
        const double U = s.u() - 0.5, V = s.uo();
        const double us = 0.5 - fabs(U);
        const double k = floor(__dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(__ddiv_rn(__dmul_rn(2.0, a), us), b), U), lam), 0.43));
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        const double lhs = __dadd_rn(__dadd_rn(log(V), log(invalpha)), -log(__dadd_rn(__ddiv_rn(a, __dmul_rn(us, us)), b)));
        const double rhs = __dadd_rn(__dadd_rn(-lam, __dmul_rn(k, loglam)), -lgamma(k + 1.0));
        if (lhs <= rhs) return k;
    }
    return floor(lam);
}
__device__ double draw_dist(Stream &s, const lpk_dist &d) {
    switch (d.kind) {
        case LPK_DIST_CONSTANT: return d.a;
        case LPK_DIST_EXPONENTIAL: return __dmul_rn(-d.a, log(s.uo()));
        case LPK_DIST_GAMMA: return draw_gamma(s, d.a, d.b);
        case LPK_DIST_LOGNORMAL: return exp(__dadd_rn(d.a, __dmul_rn(d.b, draw_normal(s))));
        case LPK_DIST_NORMAL: return __dadd_rn(d.a, __dmul_rn(d.b, draw_normal(s)));
        case LPK_DIST_POISSON: return draw_poisson(s, d.a);
        default: return __dadd_rn(d.a, floor(__dmul_rn(s.u(), d.b - d.a)));  // LPK_DIST_UNIFORM: integers in [a, b)
    }
}
// numpy's float64 -> int8 assignment: truncate toward zero, keep the low byte
__device__ __forceinline__ int wrap8(double v) {
    const double t = trunc(v);
    const long long q = (t >= 9.2e18 || t <= -9.2e18 || t != t) ? 0ll : (long long)t;
    return (int)(int8_t)(uint8_t)(q & 0xFF);
}
__device__ __forceinline__ int clip_int(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- acq_risk_multiplier / daily_infectivity (reference populate_heterogeneous_values, model.py:816-866)
__global__ void k_init_heterogeneity(int64_t start, int64_t end, float *__restrict__ risk, float *__restrict__ inf, double mu_ln,
                                     double sigma_ln, double scale_gamma, double rho, double rho_c, int heterogeneity, double mean_gamma,
                                     uint64_t seed, uint64_t id_base) {
    for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (int64_t)gridDim.x * blockDim.x) {
        if (!heterogeneity) {
            risk[i] = 1.0f;
            inf[i] = (float)mean_gamma;
            continue;
        }
        Stream s(seed, (uint64_t)i + id_base, LPK_STAGE_INIT_HET);
        const double r = sqrt(__dmul_rn(-2.0, log(s.uo()))), th = __dmul_rn(6.283185307179586, s.u());
        const double z0 = __dmul_rn(r, cos(th)), z1 = __dmul_rn(r, sin(th));
        const double zc = __dadd_rn(__dmul_rn(rho, z0), __dmul_rn(rho_c, z1));  // row 2 of the Cholesky factor
        risk[i] = (float)exp(__dadd_rn(mu_ln, __dmul_rn(sigma_ln, z0)));
        // gamma.ppf(norm.cdf(z), a = 1, scale) = -scale * log(1 - Phi(z)) = -scale * log(erfc(z / sqrt 2) / 2)
        inf[i] = (float)__dmul_rn(-scale_gamma, log(__dmul_rn(0.5, erfc(__dmul_rn(zc, 0.7071067811865476)))));
    }
}

// ---- exposure / infection / paralysis timers for the whole capacity (reference DiseaseState_ABM.__init__, model.py:571-587)
__global__ void k_init_timers(int64_t start, int64_t end, int8_t *__restrict__ et, int8_t *__restrict__ it, int8_t *__restrict__ pt,
                              lpk_dist dexp, lpk_dist dinf, lpk_dist dpar, uint64_t seed, uint64_t id_base) {
    for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t id = (uint64_t)i + id_base;
        Stream se(seed, id, LPK_STAGE_INIT_EXP), si(seed, id, LPK_STAGE_INIT_INF), sp(seed, id, LPK_STAGE_INIT_PAR);
        const int e = clip_int(wrap8(draw_dist(se, dexp)), 0, 127);
        const int f = clip_int(wrap8(draw_dist(si, dinf)), 0, 127);
        const double raw = __dadd_rn(draw_dist(sp, dpar), -(double)e);  // onset is measured from exposure
        const double hi = (double)f, clipped = fmin(fmax(raw, 0.0), hi);   // np.clip(raw, 0, min(infection_timer, 127))
        et[i] = (int8_t)e;
        it[i] = (int8_t)f;
        pt[i] = (int8_t)wrap8(clipped);
    }
}

// ---- ages, lifespans, routine-immunisation dates (reference model.py:1578-1596, 1605-1611, 1893-1894; laser-core's
//      KaplanMeierEstimator.predict_age_at_death as restated in core.py)
__global__ void k_init_demography(lpk_demog_args a) {
    extern __shared__ double s_cdf[];
    for (int k = threadIdx.x; k < a.n_bins; k += blockDim.x) s_cdf[k] = a.bin_cdf[k];
    __syncthreads();
    const double total_w = s_cdf[a.n_bins - 1];
    for (int64_t i = a.start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.end; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t id = (uint64_t)i + a.id_base;
        Stream sa(a.seed, id, LPK_STAGE_INIT_AGE);
        const double ub = __dmul_rn(sa.u(), total_w);
        int lo = 0, hi = a.n_bins;  // searchsorted(cdf, ub, side="right")
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_cdf[mid] <= ub) lo = mid + 1; else hi = mid;
        }
        const int bin = lo < a.n_bins - 1 ? lo : a.n_bins - 1;
        const int blo = a.bin_lo[bin], bhi = a.bin_hi[bin];
        int age = blo + (int)floor(__dmul_rn(sa.u(), (double)(bhi - blo)));  // np.random.randint(lo, hi)
        if (age <= 0) age = 1;
        a.date_of_birth[i] = -age;
        if (a.date_of_death) {
            Stream sl(a.seed, id, LPK_STAGE_INIT_LIFE);
            const double u1 = sl.u(), u2 = sl.u();
            int ay = age / 365;
            if (ay > a.max_year) ay = a.max_year;
            const long long total = a.cum_deaths[a.max_year + 1], already = a.cum_deaths[ay];
            const long long left = total - already > 1 ? total - already : 1;
            const long long draw = already + 1 + (long long)floor(__dmul_rn(u1, (double)left));
            int l2 = 0, h2 = a.max_year + 2;  // searchsorted(cd, draw, side="left")
            while (l2 < h2) {
                const int mid = (l2 + h2) >> 1;
                if (a.cum_deaths[mid] < draw) l2 = mid + 1; else h2 = mid;
            }
            int yod = l2 - 1;
            yod = yod < ay ? ay : (yod > a.max_year ? a.max_year : yod);
            const int rest = age % 365;
            const int doy = (yod == ay) ? rest + 1 + (int)floor(__dmul_rn(u2, (double)(364 - rest > 1 ? 364 - rest : 1)))
                                        : (int)floor(__dmul_rn(u2, 365.0));
            a.date_of_death[i] = yod * 365 + doy - age;
        }
        if (a.ri_timer) {
            Stream sr(a.seed, id, LPK_STAGE_INIT_RI);
            const double due = __dadd_rn((double)(-age), __dadd_rn(42.0, __dmul_rn(56.0, sr.u())));  // dob + U(42, 98)
            a.ri_timer[i] = (int16_t)(uint16_t)((long long)trunc(due) & 0xFFFF);                      // astype(int32) into an int16 column
        }
    }
}

// ---- chronically missed: EXACTLY n_missed agents, uniformly without replacement (np.random.choice, model.py:154-159).
// Agent i carries the 64-bit key Philox(seed; i, MISSED); the missed are the n_missed smallest keys.  The keys are never
// stored: a 4-pass radix select (16 bits per pass) regenerates them, counts digits, and narrows the threshold's prefix.
__device__ __forceinline__ uint64_t missed_key(uint64_t seed, uint64_t id) {
    uint32_t x[4];
    philox_agent(seed, id, 0u, LPK_STAGE_INIT_MISSED, x);
    return ((uint64_t)x[0] << 32) | x[1];
}
// ws: uint32 hist[65536]; then uint64 {prefix, remaining}
__global__ void k_missed_hist(int64_t n, uint64_t seed, uint64_t id_base, int pass, uint32_t *__restrict__ ws) {
    const uint64_t *state = reinterpret_cast<const uint64_t *>(ws + 65536);
    const uint64_t prefix = state[0];
    const int shift = 48 - 16 * pass;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = missed_key(seed, (uint64_t)i + id_base);
        if (pass == 0 || (key >> (shift + 16)) == prefix) atomicAdd(&ws[(key >> shift) & 0xFFFFu], 1u);
    }
}
__global__ void __launch_bounds__(1024) k_missed_pick(uint32_t *__restrict__ ws) {
    __shared__ unsigned long long s_part[1024];
    uint64_t *state = reinterpret_cast<uint64_t *>(ws + 65536);
    const unsigned long long want = state[1];  // the threshold is the want-th smallest key with this prefix (1-based)
    const int tid = threadIdx.x;
    unsigned long long sum = 0;
    for (int k = 0; k < 64; ++k) sum += ws[tid * 64 + k];
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        unsigned long long cum = 0;
        int c = 0;
        while (c < 1023 && cum + s_part[c] < want) cum += s_part[c++];
        int b = c * 64;
        while (b < c * 64 + 63 && cum + ws[b] < want) cum += ws[b++];
        state[0] = (state[0] << 16) | (uint64_t)b;
        state[1] = want - cum;
    }
    __syncthreads();
    for (int k = tid; k < 65536; k += 1024) ws[k] = 0u;
}
__global__ void k_missed_begin(uint32_t *__restrict__ ws, unsigned long long want) {
    uint64_t *state = reinterpret_cast<uint64_t *>(ws + 65536);
    state[0] = 0ull;
    state[1] = want;
}
__global__ void k_missed_mark(int64_t n, uint64_t seed, uint64_t id_base, const uint32_t *__restrict__ ws, uint8_t *__restrict__ missed,
                              int none) {
    const uint64_t threshold = reinterpret_cast<const uint64_t *>(ws + 65536)[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        missed[i] = (!none && missed_key(seed, (uint64_t)i + id_base) <= threshold) ? 1 : 0;
}

int init_grid() { return lpk_sm_count() * 8; }

}  // namespace

extern "C" int lpk_init_heterogeneity(int64_t start, int64_t end, float *acq_risk, float *infectivity, double mu_ln, double sigma_ln,
                                      double scale_gamma, double rho, int32_t heterogeneity, double mean_gamma, uint64_t seed,
                                      uint64_t id_base, void *stream) {
    REQUIRE(acq_risk && infectivity && start >= 0 && end >= start, "init_heterogeneity arrays");
    REQUIRE(rho >= -1.0 && rho <= 1.0, "init_heterogeneity rho");
    if (end == start) return LPK_OK;
    k_init_heterogeneity<<<init_grid(), 256, 0, as_stream(stream)>>>(start, end, acq_risk, infectivity, mu_ln, sigma_ln, scale_gamma, rho,
                                                                     sqrt(1.0 - rho * rho), heterogeneity, mean_gamma, seed, id_base);
    CUDA_TRY(cudaGetLastError(), "lpk_init_heterogeneity");
    return LPK_OK;
}

static bool dist_ok(const lpk_dist *d) { return d && d->kind >= LPK_DIST_CONSTANT && d->kind <= LPK_DIST_UNIFORM; }

extern "C" int lpk_init_timers(int64_t start, int64_t end, int8_t *exposure_timer, int8_t *infection_timer, int8_t *paralysis_timer,
                               const lpk_dist *dur_exp, const lpk_dist *dur_inf, const lpk_dist *t_to_paralysis, uint64_t seed,
                               uint64_t id_base, void *stream) {
    REQUIRE(exposure_timer && infection_timer && paralysis_timer && start >= 0 && end >= start, "init_timers arrays");
    REQUIRE(dist_ok(dur_exp) && dist_ok(dur_inf) && dist_ok(t_to_paralysis), "init_timers distributions");
    if (end == start) return LPK_OK;
    k_init_timers<<<init_grid(), 256, 0, as_stream(stream)>>>(start, end, exposure_timer, infection_timer, paralysis_timer, *dur_exp, *dur_inf,
                                                              *t_to_paralysis, seed, id_base);
    CUDA_TRY(cudaGetLastError(), "lpk_init_timers");
    return LPK_OK;
}

extern "C" int lpk_init_demography(const lpk_demog_args *args, void *stream) {
    REQUIRE(args, "init_demography null struct");
    const lpk_demog_args &a = *args;
    REQUIRE(a.date_of_birth && a.bin_cdf && a.bin_lo && a.bin_hi && a.n_bins > 0 && a.n_bins <= 4096 && a.start >= 0 && a.end >= a.start,
            "init_demography ages");
    REQUIRE(!a.date_of_death || (a.cum_deaths && a.max_year > 0), "init_demography lifespans");
    if (a.end == a.start) return LPK_OK;
    k_init_demography<<<init_grid(), 256, a.n_bins * sizeof(double), as_stream(stream)>>>(a);
    CUDA_TRY(cudaGetLastError(), "lpk_init_demography");
    return LPK_OK;
}

extern "C" int lpk_init_missed(int64_t n, int64_t n_missed, uint8_t *chronically_missed, uint64_t seed, uint64_t id_base, uint32_t *ws,
                               void *stream) {
    REQUIRE(chronically_missed && ws && n >= 0 && n_missed >= 0 && n_missed <= n, "init_missed");
    REQUIRE(ALIGNED(ws, 8), "init_missed workspace alignment");
    if (n == 0) return LPK_OK;
    cudaStream_t st = as_stream(stream);
    CUDA_TRY(cudaMemsetAsync(ws, 0, LPK_MISSED_WS_WORDS * sizeof(uint32_t), st), "lpk_init_missed workspace");
    if (n_missed > 0) {
        k_missed_begin<<<1, 1, 0, st>>>(ws, (unsigned long long)n_missed);
        for (int pass = 0; pass < 4; ++pass) {
            k_missed_hist<<<init_grid(), 256, 0, st>>>(n, seed, id_base, pass, ws);
            k_missed_pick<<<1, 1024, 0, st>>>(ws);
        }
    }
    k_missed_mark<<<init_grid(), 256, 0, st>>>(n, seed, id_base, ws, chronically_missed, n_missed == 0 ? 1 : 0);
    CUDA_TRY(cudaGetLastError(), "lpk_init_missed mark");
    return LPK_OK;
}
