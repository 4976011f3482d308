/*
 * lp_oracle_init.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement of the population initialisers of laser-polio, the checker for
 * laser-polio_b200/csrc/lpk_init.cu.  What is restated and where the reference does it:
 *   orc_init_heterogeneity  populate_heterogeneous_values, model.py:816-866 (Gaussian copula: lognormal risk, gamma(1)
 *                           infectivity through gamma.ppf(norm.cdf(z)))
 *   orc_init_timers         DiseaseState_ABM.__init__, model.py:571-587 (dur_exp / dur_inf / t_to_paralysis, int8 casts, clips)
 *   orc_init_demography     VitalDynamics_ABM._initialize_ages_and_births, model.py:1578-1596; _initialize_deaths, 1605-1611
 *                           (laser-core KaplanMeierEstimator.predict_age_at_death, ~=0.6, not in the checkout: restated from
 *                           its published behaviour, see laser-polio_b200/core.py -- parity unpinned for that part);
 *                           RI_ABM._initialize_people_fields, model.py:1893-1894
 *   orc_init_missed         SEIR_ABM.__init__, model.py:154-159 (np.random.choice(n, int(missed_frac * n), replace=False))
 *
 * The reference draws from numpy's global Mersenne stream, which no device can replay; the distributions are what is
 * pinned: tests/test_oracle_init.py holds these samplers to the reference's own numpy / scipy expressions (KS and moment
 * tests, and a golden sample of populate_heterogeneous_values produced by the reference function itself,
 * tests/golden/make_golden_init.py).  The CUDA kernels are then held to this file draw for draw: every value is a pure
 * function of Philox4x32-10(seed; agent id, block, stage).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

enum { ST_HET = 16, ST_EXP = 17, ST_INF = 18, ST_PAR = 19, ST_AGE = 20, ST_LIFE = 21, ST_RI = 22, ST_MISSED = 23 };
enum { D_CONSTANT = 0, D_EXPONENTIAL, D_GAMMA, D_LOGNORMAL, D_NORMAL, D_POISSON, D_UNIFORM };

typedef struct { int32_t kind; double a, b; } orc_dist;

typedef struct {
    uint64_t seed, id;
    uint32_t stage, blk, x[4];
    int pos;
} stream_t;

static void st_open(stream_t *s, uint64_t seed, uint64_t id, uint32_t stage) { s->seed = seed; s->id = id; s->stage = stage; s->blk = 0; s->pos = 4; }
static uint64_t st_bits53(stream_t *s) {
    if (s->pos >= 4) {
        const uint32_t ctr[4] = {(uint32_t)s->id, (uint32_t)(s->id >> 32), s->blk++, s->stage};
        const uint32_t key[2] = {(uint32_t)s->seed, (uint32_t)(s->seed >> 32)};
        orc_philox4x32_10(ctr, key, s->x);
        s->pos = 0;
    }
    const uint64_t v = (((uint64_t)s->x[s->pos] << 32) | s->x[s->pos + 1]) >> 11;
    s->pos += 2;
    return v;
}
static double st_u(stream_t *s) { return (double)st_bits53(s) * (1.0 / 9007199254740992.0); }           /* [0, 1) */
static double st_uo(stream_t *s) { return (double)(st_bits53(s) + 1ull) * (1.0 / 9007199254740992.0); } /* (0, 1] */

static double draw_normal(stream_t *s) { /* Box-Muller, cosine branch */
    const double r = sqrt(-2.0 * log(st_uo(s)));
    return r * cos(6.283185307179586 * st_u(s));
}
static double draw_gamma(stream_t *s, double shape, double scale) { /* Marsaglia & Tsang 2000 */
    double boost = 1.0;
    if (shape < 1.0) { boost = pow(st_uo(s), 1.0 / shape); shape += 1.0; }
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 64; ++it) {
        const double z = draw_normal(s);
        const double w = 1.0 + c * z;
        const double uu = st_uo(s);
        if (w <= 0.0) continue;
        const double v = (w * w) * w;
        const double lhs = log(uu);
        const double rhs = ((0.5 * (z * z)) + d) + ((-d * v) + (d * log(v)));
        if (lhs < rhs) return ((d * v) * scale) * boost;
    }
    return (d * scale) * boost;
}
static double draw_poisson(stream_t *s, double lam) {
    if (!(lam > 0.0)) return 0.0;
    if (lam < 30.0) { /* inversion by sequential search */
        const double u = st_u(s);
        double p = exp(-lam), cum = p;
        int k = 0;
        while (u >= cum && k < 1000) { ++k; p = (p * lam) / (double)k; cum = cum + p; }
        return (double)k;
    }
    /* PTRS, Hoermann 1993 (numpy's algorithm for lam >= 10) */
    const double slam = sqrt(lam), loglam = log(lam), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 + -(3.6224 / (b - 2.0));
    for (int it = 0; it < 256; ++it) {
        const double U = st_u(s) - 0.5, V = st_uo(s);
        const double us = 0.5 - fabs(U);
        const double k = floor(((((2.0 * a) / us) + b) * U + lam) + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        const double lhs = (log(V) + log(invalpha)) + -log((a / (us * us)) + b);
        const double rhs = (-lam + (k * loglam)) + -lgamma(k + 1.0);
        if (lhs <= rhs) return k;
    }
    return floor(lam);
}
static double draw_dist(stream_t *s, const orc_dist *d) {
    switch (d->kind) {
        case D_CONSTANT: return d->a;
        case D_EXPONENTIAL: return -d->a * log(st_uo(s));
        case D_GAMMA: return draw_gamma(s, d->a, d->b);
        case D_LOGNORMAL: return exp(d->a + (d->b * draw_normal(s)));
        case D_NORMAL: return d->a + (d->b * draw_normal(s));
        case D_POISSON: return draw_poisson(s, d->a);
        default: return d->a + floor(st_u(s) * (d->b - d->a));
    }
}
/* one sample of a distribution per agent (distribution tests) */
void orc_init_draw(int64_t n, const orc_dist *d, uint64_t seed, uint32_t stage, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        stream_t s;
        st_open(&s, seed, (uint64_t)i, stage);
        out[i] = draw_dist(&s, d);
    }
}

/* numpy's float64 -> int8 assignment cast: truncate toward zero, keep the low byte */
static int wrap8(double v) {
    const double t = trunc(v);
    const long long q = (t >= 9.2e18 || t <= -9.2e18 || t != t) ? 0ll : (long long)t;
    return (int)(int8_t)(uint8_t)(q & 0xFF);
}
static int clip_int(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void orc_init_heterogeneity(int64_t start, int64_t end, float *risk, float *inf, double mu_ln, double sigma_ln, double scale_gamma,
                            double rho, int32_t heterogeneity, double mean_gamma, uint64_t seed, uint64_t id_base) {
    const double rho_c = sqrt(1.0 - rho * rho); /* np.linalg.cholesky([[1, rho], [rho, 1]]) = [[1, 0], [rho, rho_c]], model.py:848-849 */
#pragma omp parallel for schedule(static)
    for (int64_t i = start; i < end; ++i) {
        if (!heterogeneity) { risk[i] = 1.0f; inf[i] = (float)mean_gamma; continue; } /* model.py:864-866 */
        stream_t s;
        st_open(&s, seed, (uint64_t)i + id_base, ST_HET);
        const double r = sqrt(-2.0 * log(st_uo(&s))), th = 6.283185307179586 * st_u(&s);
        const double z0 = r * cos(th), z1 = r * sin(th);
        const double zc = (rho * z0) + (rho_c * z1);                  /* z @ L.T, model.py:858 */
        risk[i] = (float)exp(mu_ln + (sigma_ln * z0));                /* model.py:861 */
        inf[i] = (float)(-scale_gamma * log(0.5 * erfc(zc * 0.7071067811865476))); /* gamma.ppf(norm.cdf(zc), a=1, scale), model.py:862 */
    }
}

void orc_init_timers(int64_t start, int64_t end, int8_t *et, int8_t *it, int8_t *pt, const orc_dist *dexp, const orc_dist *dinf,
                     const orc_dist *dpar, uint64_t seed, uint64_t id_base) {
#pragma omp parallel for schedule(static)
    for (int64_t i = start; i < end; ++i) {
        const uint64_t id = (uint64_t)i + id_base;
        stream_t se, si, sp;
        st_open(&se, seed, id, ST_EXP); st_open(&si, seed, id, ST_INF); st_open(&sp, seed, id, ST_PAR);
        const int e = clip_int(wrap8(draw_dist(&se, dexp)), 0, 127); /* model.py:575, 578 */
        const int f = clip_int(wrap8(draw_dist(&si, dinf)), 0, 127); /* model.py:576, 579 */
        const double raw = draw_dist(&sp, dpar) + -(double)e;        /* model.py:583-584 */
        const double clipped = fmin(fmax(raw, 0.0), (double)f);      /* model.py:585-586 */
        et[i] = (int8_t)e; it[i] = (int8_t)f; pt[i] = (int8_t)wrap8(clipped);
    }
}

void orc_init_demography(int64_t start, int64_t end, int32_t *dob, int32_t *dod, int16_t *ri_timer, const double *bin_cdf,
                         const int32_t *bin_lo, const int32_t *bin_hi, int32_t n_bins, const int64_t *cum_deaths, int32_t max_year,
                         uint64_t seed, uint64_t id_base) {
    const double total_w = bin_cdf[n_bins - 1];
#pragma omp parallel for schedule(static)
    for (int64_t i = start; i < end; ++i) {
        const uint64_t id = (uint64_t)i + id_base;
        stream_t sa;
        st_open(&sa, seed, id, ST_AGE);
        const double ub = st_u(&sa) * total_w;
        int lo = 0, hi = n_bins; /* searchsorted(cdf, ub, side="right") */
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (bin_cdf[mid] <= ub) lo = mid + 1; else hi = mid; }
        const int bin = lo < n_bins - 1 ? lo : n_bins - 1;
        int age = bin_lo[bin] + (int)floor(st_u(&sa) * (double)(bin_hi[bin] - bin_lo[bin])); /* model.py:1592 */
        if (age <= 0) age = 1;                                                             /* model.py:1594 */
        dob[i] = -age;
        if (dod) {
            stream_t sl;
            st_open(&sl, seed, id, ST_LIFE);
            const double u1 = st_u(&sl), u2 = st_u(&sl);
            int ay = age / 365;
            if (ay > max_year) ay = max_year;
            const long long total = cum_deaths[max_year + 1], already = cum_deaths[ay];
            const long long left = total - already > 1 ? total - already : 1;
            const long long draw = already + 1 + (long long)floor(u1 * (double)left);
            int l2 = 0, h2 = max_year + 2;
            while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (cum_deaths[mid] < draw) l2 = mid + 1; else h2 = mid; }
            int yod = l2 - 1;
            yod = yod < ay ? ay : (yod > max_year ? max_year : yod);
            const int rest = age % 365;
            const int doy = (yod == ay) ? rest + 1 + (int)floor(u2 * (double)(364 - rest > 1 ? 364 - rest : 1)) : (int)floor(u2 * 365.0);
            dod[i] = yod * 365 + doy - age; /* model.py:1607-1609 */
        }
        if (ri_timer) {
            stream_t sr;
            st_open(&sr, seed, id, ST_RI);
            const double due = (double)(-age) + (42.0 + (56.0 * st_u(&sr))); /* model.py:1893-1894 */
            ri_timer[i] = (int16_t)(uint16_t)((long long)trunc(due) & 0xFFFF);
        }
    }
}

static int cmp_u64(const void *a, const void *b) { const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return (x > y) - (x < y); }
void orc_init_missed(int64_t n, int64_t n_missed, uint8_t *missed, uint64_t seed, uint64_t id_base) {
    if (n <= 0) return;
    uint64_t *keys = (uint64_t *)malloc((size_t)n * sizeof(uint64_t)), *sorted = (uint64_t *)malloc((size_t)n * sizeof(uint64_t));
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t id = (uint64_t)i + id_base;
        const uint32_t ctr[4] = {(uint32_t)id, (uint32_t)(id >> 32), 0u, ST_MISSED};
        const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t x[4];
        orc_philox4x32_10(ctr, key, x);
        keys[i] = ((uint64_t)x[0] << 32) | x[1];
    }
    memcpy(sorted, keys, (size_t)n * sizeof(uint64_t));
    qsort(sorted, (size_t)n, sizeof(uint64_t), cmp_u64);
    for (int64_t i = 0; i < n; ++i) missed[i] = (n_missed > 0 && keys[i] <= sorted[n_missed - 1]) ? 1 : 0;
    free(keys); free(sorted);
}
