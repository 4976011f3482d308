// lpk_host.cuh -- host-side plumbing shared by the translation units of liblpk (errors, grid sizing).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/lpk.h"
#include "lpk_common.cuh"

extern thread_local char lpk_g_err[256];

static inline int lpk_set_cuda_err(cudaError_t e, const char *where) {
    snprintf(lpk_g_err, sizeof(lpk_g_err), "%s: %s", where, cudaGetErrorString(e));
    return LPK_ERR_CUDA;
}
static inline int lpk_set_arg_err(const char *what) {
    snprintf(lpk_g_err, sizeof(lpk_g_err), "bad argument: %s", what);
    return LPK_ERR_ARG;
}

#define REQUIRE(cond, what) do { if (!(cond)) return lpk_set_arg_err(what); } while (0)
#define ALIGNED(p, a) ((reinterpret_cast<uintptr_t>(p) & ((a) - 1)) == 0)
#define CUDA_TRY(expr, where) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return lpk_set_cuda_err(e_, where); } while (0)

int lpk_sm_count();
// persistent grid: a multiple of the SM count, never more blocks than there are groups of LPK_WARPS tiles
static inline int lpk_agent_grid(int64_t n, int blocks_per_sm) {
    const int64_t tiles = (n + LPK_TILE - 1) / LPK_TILE;
    const int64_t want = (tiles + LPK_WARPS - 1) / LPK_WARPS;
    const int64_t cap = (int64_t)lpk_sm_count() * blocks_per_sm;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}
static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
