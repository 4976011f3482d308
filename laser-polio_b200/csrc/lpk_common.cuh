// lpk_common.cuh -- device helpers shared by the sm_100a kernels of liblpk.
//
// Work decomposition ("quad tiling").  A warp processes the agent table in tiles
// of 512 consecutive agents.  Inside a tile, lane L owns the four *quads*
// q = j*32 + L (j = 0..3), a quad being 4 consecutive agents.  With that mapping
// every column is read with one fully coalesced instruction per j:
//   1-byte columns (state, strain, timers, flags)  -> 32-bit  loads, 128 B per warp
//   2-byte columns (node_id, ri_timer)             -> 64-bit  loads, 256 B per warp
//   4-byte columns (risk, infectivity, dob, dod)   -> 128-bit loads, 512 B per warp
// and a quad is also the unit of the exposure RNG (one Philox4x32 block = four
// 32-bit words = one word per agent of the quad).
//
// Blocks own contiguous ranges of tiles (warps interleaved inside the range) so a
// thread's per-node partial sums stay in registers across tiles and are flushed
// with warp-aggregated reductions only when the node changes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lpk.h"

#define LPK_TILE 512            // agents per warp tile
#define LPK_BLOCK 256           // threads per block
#define LPK_WARPS (LPK_BLOCK / 32)
#define LPK_FULL 0xFFFFFFFFu

// Functions marked LPK_HD also compile for the host: tests/hot_model builds the per-agent logic of the fused pass
// (lpk_hot.cuh) as a plain CPU program and checks it on the CPU without a GPU.
#define LPK_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define LPK_MULHI(a, b) __umulhi((a), (b))
#else
#include <cmath>
#include <cstring>
#define LPK_MULHI(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
#endif
LPK_HD uint32_t lpk_f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
LPK_HD float lpk_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// ---------------------------------------------------------------- Philox4x32-10
LPK_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = LPK_MULHI(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = LPK_MULHI(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

LPK_HD void philox_agent(uint64_t seed, uint64_t idx, uint32_t tick, uint32_t stage,
                                             uint32_t out[4]) {
    philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), tick, stage, (uint32_t)seed, (uint32_t)(seed >> 32), out);
}

LPK_HD double u53(uint32_t hi, uint32_t lo) {
    const unsigned long long v = (((unsigned long long)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------- exposure draw
// Agent `id` (= slot + id_base) owns one 32-bit word X = h16 << 16 | l16 per tick for its exposure trial.  The halves are
// the same half-word `hw` of two Philox blocks (stages EXPOSE and EXPOSE_LO) shared by 8 agents -- the quads one lane owns
// in an even / odd pair of 128-agent rows:
//     ctr = (id >> 8) * 32 + ((id >> 2) & 31)        hw = ((id >> 7) & 1) * 4 + (id & 3)
// so ONE Philox block serves 8 agents in the streaming pass: the high half alone decides > 99.9 % of the trials
// (h16 >= risk * tau * 2^16 cannot be a hit) and the low-half block is generated only for the rest.  The trial itself is
// unchanged: hit iff X < floor(p * 2^32).
LPK_HD uint64_t expose_ctr(uint64_t id) { return ((id >> 8) << 5) | ((id >> 2) & 31u); }
LPK_HD int expose_hw(uint64_t id) { return (int)(((id >> 7) & 1u) * 4u + (id & 3u)); }
LPK_HD uint32_t half_word(const uint32_t x[4], int hw) { return (x[hw >> 1] >> (16 * (hw & 1))) & 0xFFFFu; }
// the four words of the aligned quad starting at agent id0 (id0 % 4 == 0)
LPK_HD void expose_words_quad(uint64_t seed, uint64_t id0, uint32_t tick, uint32_t out[4]) {
    const uint64_t c = expose_ctr(id0);
    const int hw0 = expose_hw(id0);
    uint32_t h[4], l[4];
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), tick, LPK_STAGE_EXPOSE, (uint32_t)seed, (uint32_t)(seed >> 32), h);
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), tick, LPK_STAGE_EXPOSE_LO, (uint32_t)seed, (uint32_t)(seed >> 32), l);
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = (half_word(h, hw0 + k) << 16) | half_word(l, hw0 + k);
}

// ---------------------------------------------------------------- state-word masks
// The four state bytes of a quad are 0xFF (dead / unborn), 0 (S), 1 (E), 2 (I), 3 (R); each mask has bit 0 of byte k set
// when agent k is in the class.
__device__ __forceinline__ uint32_t mask_S(uint32_t w) { return ~(w | (w >> 1) | (w >> 7)) & 0x01010101u; }
__device__ __forceinline__ uint32_t mask_EI(uint32_t w) { return (w ^ (w >> 1)) & 0x01010101u; }
__device__ __forceinline__ uint32_t mask_alive(uint32_t w) { return ~(w >> 7) & 0x01010101u; }

// ---------------------------------------------------------------- byte-lane helpers
__device__ __forceinline__ int8_t byte_of(uint32_t w, int k) { return (int8_t)((w >> (8 * k)) & 0xFFu); }
__device__ __forceinline__ uint32_t set_byte(uint32_t w, int k, int8_t v) {
    return (w & ~(0xFFu << (8 * k))) | (((uint32_t)(uint8_t)v) << (8 * k));
}
// any byte of w equal to the byte value b?
__device__ __forceinline__ bool any_byte_eq(uint32_t w, uint32_t b) { return __vcmpeq4(w, b * 0x01010101u) != 0u; }
// number of bytes of w equal to b
__device__ __forceinline__ int count_byte_eq(uint32_t w, uint32_t b) { return __popc(__vcmpeq4(w, b * 0x01010101u)) >> 3; }

// ---------------------------------------------------------------- quad loads (valid = agents of the quad in range)
// Out-of-range agents read as "dead" (state -1) so every stage skips them.
__device__ __forceinline__ uint32_t load_b4(const int8_t *col, int64_t i, int valid, int8_t fill = -1) {
    if (valid == 4) return *reinterpret_cast<const uint32_t *>(col + i);
    uint32_t w = ((uint32_t)(uint8_t)fill) * 0x01010101u;
    for (int k = 0; k < valid; ++k) w = set_byte(w, k, col[i + k]);
    return w;
}
__device__ __forceinline__ void store_b4(int8_t *col, int64_t i, int valid, uint32_t w) {
    if (valid == 4) { *reinterpret_cast<uint32_t *>(col + i) = w; return; }
    for (int k = 0; k < valid; ++k) col[i + k] = byte_of(w, k);
}
__device__ __forceinline__ void load_s4(const int16_t *col, int64_t i, int valid, int v[4]) {
    if (valid == 4) {
        const short4 s = *reinterpret_cast<const short4 *>(col + i);
        v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w;
        return;
    }
    for (int k = 0; k < 4; ++k) v[k] = (k < valid) ? (int)col[i + k] : -1;
}
__device__ __forceinline__ void load_f4(const float *col, int64_t i, int valid, float v[4]) {
    if (valid == 4) {
        const float4 f = __ldg(reinterpret_cast<const float4 *>(col + i));
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        return;
    }
    for (int k = 0; k < 4; ++k) v[k] = (k < valid) ? col[i + k] : 0.f;
}
__device__ __forceinline__ void load_i4(const int32_t *col, int64_t i, int valid, int v[4]) {
    if (valid == 4) {
        const int4 f = __ldg(reinterpret_cast<const int4 *>(col + i));
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        return;
    }
    for (int k = 0; k < 4; ++k) v[k] = (k < valid) ? col[i + k] : 0;
}

// ---------------------------------------------------------------- tile iteration
struct TileRange {
    int64_t lo, hi;  // this block's tiles [lo, hi)
};
__device__ __forceinline__ TileRange block_tiles(int64_t n_agents) {
    const int64_t tiles = (n_agents + LPK_TILE - 1) / LPK_TILE;
    TileRange r;
    r.lo = tiles * (int64_t)blockIdx.x / gridDim.x;
    r.hi = tiles * (int64_t)(blockIdx.x + 1) / gridDim.x;
    return r;
}
// first agent of quad j of this lane in tile t, and how many of its 4 agents are < n
__device__ __forceinline__ int64_t quad_base(int64_t tile, int j, int lane) {
    return tile * LPK_TILE + (int64_t)(j * 32 + lane) * 4;
}
__device__ __forceinline__ int quad_valid(int64_t base, int64_t n) {
    const int64_t r = n - base;
    return r >= 4 ? 4 : (r <= 0 ? 0 : (int)r);
}

// ---------------------------------------------------------------- per-node register accumulators
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LPK_FULL, v, o);
    return v;
}

template <int NI, int NL>
struct NodeAcc {
    int node;
    int ci[NI > 0 ? NI : 1];
    long long cl[NL > 0 ? NL : 1];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < NI; ++k) ci[k] = 0;
#pragma unroll
        for (int k = 0; k < NL; ++k) cl[k] = 0;
    }
    __device__ __forceinline__ void init() { node = -1; clear(); }
    // make `nd` the current node, flushing the previous node's partials (thread-level atomics)
    template <class Flush>
    __device__ __forceinline__ void select(int nd, Flush &flush) {
        if (nd != node) {
            if (node >= 0) flush(node, ci, cl);
            clear();
            node = nd;
        }
    }
    // end of the warp's range: one aggregated flush per warp when every lane ended on the same node
    template <class Flush>
    __device__ __forceinline__ void finish_warp(Flush &flush) {
        const int n0 = __reduce_max_sync(LPK_FULL, node);
        const bool uniform = __all_sync(LPK_FULL, node == n0 || node < 0);
        if (uniform) {
            if (n0 < 0) return;
            int ri[NI > 0 ? NI : 1];
            long long rl[NL > 0 ? NL : 1];
#pragma unroll
            for (int k = 0; k < NI; ++k) ri[k] = __reduce_add_sync(LPK_FULL, ci[k]);
#pragma unroll
            for (int k = 0; k < NL; ++k) rl[k] = warp_sum_ll(cl[k]);
            if ((threadIdx.x & 31) == 0) flush(n0, ri, rl);
        } else if (node >= 0) {
            flush(node, ci, cl);
        }
        init();
    }
};

__device__ __forceinline__ void red_add(int32_t *p, int v) { if (v) atomicAdd(p, v); }
__device__ __forceinline__ void red_add(int64_t *p, long long v) {
    if (v) atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
}

LPK_HD long long to_fx(double v) {
#ifdef __CUDA_ARCH__
    return __double2ll_rn(v * LPK_FX_SCALE);
#else
    return llrint(v * LPK_FX_SCALE);
#endif
}
// a float weight (risk) in the 2^30 fixed point of the carried tallies
LPK_HD long long risk_fx(float rk) {
#ifdef __CUDA_ARCH__
    return __float2ll_rn(rk * 1073741824.0f);
#else
    return llrintf(rk * 1073741824.0f);
#endif
}


// ---------------------------------------------------------------- exposure probability and the risk histogram
// Susceptible i of node n is exposed with probability p_i = 1 - exp(-risk_i * tau_n); tau_n solves
// sum_i p_i = expected exposures of the node (include/lpk.h, T2/T3).  p_expose is specified with fmaf / floorf /
// exact power-of-two scaling only, so a CPU restatement reproduces it bit for bit.
LPK_HD float p_expose(float x) {
    if (!(x > 0.f)) return 0.f;
    if (x < 0.0625f) {  // x - x^2/2 + x^3/6 - x^4/24 + x^5/120, relative error < 2e-9
        float t = fmaf(-x, 0.008333333767950535f, 0.0416666679084301f);
        t = fmaf(-x, t, 0.1666666716337204f);
        t = fmaf(-x, t, 0.5f);
        t = fmaf(-x, t, 1.0f);
        return x * t;
    }
    if (x >= 17.f) return 1.f;
    const float y = x * 1.4426950216293335f;  // log2(e)
    const float n = floorf(y);
    const float f = y - n;  // exact
    float r = 0.00010938752529909834f;         // 2^-f on [0, 1), |error| < 8e-8
    r = fmaf(r, f, -0.0012757162330672145f);
    r = fmaf(r, f, 0.009580058045685291f);
    r = fmaf(r, f, -0.05549103394150734f);
    r = fmaf(r, f, 0.24022436141967773f);
    r = fmaf(r, f, -0.6931470632553101f);
    r = fmaf(r, f, 1.0f);
    const float scale = lpk_u2f((uint32_t)(127 - (int)n) << 23);  // 2^-n, n in [0, 24]
    return 1.f - r * scale;
}
// x < floor(p * 2^32) with p in [0, 1]; p == 1 always hits
LPK_HD bool expose_test(float p, uint32_t x) {
#ifdef __CUDA_ARCH__
    return (unsigned long long)x < __float2ull_rz(p * 4294967296.0f);
#else
    return (unsigned long long)x < (unsigned long long)(p * 4294967296.0f);
#endif
}

// risk histogram: 8 bins per octave over [2^-12, 2^12), clamped; non-positive / NaN weights fall in bin 0
LPK_HD int risk_bin(float w) {
    if (!(w > 0.f)) return 0;
    const int b = (int)(lpk_f2u(w) >> 20) - ((127 - 12) << 3);
    return b < 0 ? 0 : (b >= LPK_RISK_BINS ? LPK_RISK_BINS - 1 : b);
}
__device__ __forceinline__ double risk_bin_weight(int b) {  // representative weight: middle of the bin
    return ldexp(1.0 + ((double)(b & 7) + 0.5) / 8.0, (b >> 3) - 12);
}

// per-warp shared-memory histogram of the current node's susceptibles, flushed to global on node change
struct WarpHist {
    int *h;
    int node;  // warp-uniform
    __device__ __forceinline__ void init(int *smem, int lane) {
        h = smem; node = -1;
        for (int b = lane; b < LPK_RISK_BINS; b += 32) h[b] = 0;
        __syncwarp();
    }
    __device__ __forceinline__ void flush(int32_t *g, int lane) {
        __syncwarp();
        if (node >= 0) {
            for (int b = lane; b < LPK_RISK_BINS; b += 32) {
                const int v = h[b];
                if (v) { atomicAdd(&g[(int64_t)node * LPK_RISK_BINS + b], v); h[b] = 0; }
            }
        }
        __syncwarp();
    }
    // make `nd` (warp-uniform) the current node
    __device__ __forceinline__ void select(int nd, int32_t *g, int lane) {
        if (nd != node) { flush(g, lane); node = nd; }
    }
};
