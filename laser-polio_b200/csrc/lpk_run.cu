// lpk_run.cu -- the tick loop of the fused path, driven from C (include/lpk.h, "A run of fused days driven from C").
//
// Reference: SEIR_ABM.run's loop body (model.py:252-263) is Python per tick; the round-1 engine assembled three argument
// structs per day in Python (~0.2 ms of host time per day against ~0.3 ms of device time at 2.2e8 agents and ~0.05 ms at
// 8 GPUs).  Here one template is kept per run and every day's arguments are derived from it in a few dozen pointer
// additions, so the host side of a day is its kernel launches and nothing else; with lpk_run.graph the launches of a
// whole span of days are captured into one CUDA graph.
//
// Node-sharded runs: the one per-tick exchange (the reference design's all-reduce of the nodes x strains
// infectivity tally) is done by k_xchg_push: every rank stores its own rows straight into every peer's HBM over NVLink
// (buffers mapped through CUDA IPC), then releases a system-scope flag; the node kernels of the consumer acquire the flags
// of all ranks before reading the gathered tally (lpk_kernels.cu, xchg_wait).  Double-buffered by sequence parity: a rank
// can only be one exchange ahead of its slowest peer, because its next push comes after its own wait on that peer's flag.
#include <vector>

#include "lpk_host.cuh"

#define LPK_XCHG_MAX_RANKS 64
#define LPK_XCHG_FLAG_STRIDE 64  // uint32 flags per parity

struct lpk_xchg {
    int rank, world;
    int64_t n_elems;
    size_t bytes;
    char *base;                             // this rank's receive area: int64 [2][n_elems], then uint32 flags [2][64]
    char *peer[LPK_XCHG_MAX_RANKS];         // every rank's receive area as mapped here (peer[rank] == base)
    char **d_peer;                          // the same table in device memory
    bool connected;
};

static __host__ __device__ inline size_t xchg_flag_offset(int64_t n_elems) { return (size_t)2 * (size_t)n_elems * sizeof(int64_t); }

extern "C" int lpk_xchg_create(int32_t rank, int32_t world, int64_t n_elems, lpk_xchg **out, void *handle_out) {
    REQUIRE(out && handle_out && world >= 1 && world <= LPK_XCHG_MAX_RANKS && rank >= 0 && rank < world && n_elems > 0, "xchg_create");
    static_assert(sizeof(cudaIpcMemHandle_t) == LPK_XCHG_HANDLE_BYTES, "IPC handle size");
    lpk_xchg *x = new lpk_xchg();
    x->rank = rank; x->world = world; x->n_elems = n_elems; x->connected = false; x->d_peer = nullptr;
    for (int r = 0; r < LPK_XCHG_MAX_RANKS; ++r) x->peer[r] = nullptr;
    x->bytes = xchg_flag_offset(n_elems) + 2 * LPK_XCHG_FLAG_STRIDE * sizeof(uint32_t);
    cudaError_t e = cudaMalloc(&x->base, x->bytes);  // plain cudaMalloc: the allocation must be exportable through CUDA IPC
    if (e != cudaSuccess) { delete x; return lpk_set_cuda_err(e, "xchg_create alloc"); }
    e = cudaMemset(x->base, 0, x->bytes);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle_out), x->base);
    if (e != cudaSuccess) { cudaFree(x->base); delete x; return lpk_set_cuda_err(e, "xchg_create handle"); }
    *out = x;
    return LPK_OK;
}

extern "C" int lpk_xchg_connect(lpk_xchg *x, const void *handles) {
    REQUIRE(x && handles && !x->connected, "xchg_connect");
    const cudaIpcMemHandle_t *h = reinterpret_cast<const cudaIpcMemHandle_t *>(handles);
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) { x->peer[r] = x->base; continue; }
        void *p = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&p, h[r], cudaIpcMemLazyEnablePeerAccess), "xchg_connect open (peer access between the GPUs?)");
        x->peer[r] = static_cast<char *>(p);
    }
    CUDA_TRY(cudaMalloc(&x->d_peer, sizeof(char *) * LPK_XCHG_MAX_RANKS), "xchg_connect table");
    CUDA_TRY(cudaMemcpy(x->d_peer, x->peer, sizeof(char *) * LPK_XCHG_MAX_RANKS, cudaMemcpyHostToDevice), "xchg_connect table");
    x->connected = true;
    return LPK_OK;
}

// Teardown is two-phase because an exporter must not free memory its peers still have mapped: every rank disconnects
// (unmaps the peers' areas), the caller runs a barrier, then every rank destroys (frees its own area).
extern "C" int lpk_xchg_disconnect(lpk_xchg *x) {
    if (!x) return LPK_OK;
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank && x->peer[r]) { cudaIpcCloseMemHandle(x->peer[r]); x->peer[r] = nullptr; }
    x->connected = false;
    return LPK_OK;
}
extern "C" int lpk_xchg_destroy(lpk_xchg *x) {
    if (!x) return LPK_OK;
    lpk_xchg_disconnect(x);
    if (x->d_peer) cudaFree(x->d_peer);
    cudaFree(x->base);
    delete x;
    return LPK_OK;
}

// block p: this rank's rows -> peer p's receive buffer of this parity, then the flag that says so
__global__ void __launch_bounds__(256) k_xchg_push(const int64_t *__restrict__ beta, int64_t lo, int64_t hi, char *const *__restrict__ peers,
                                                    int64_t n_elems, int parity, int rank, uint32_t seq) {
    char *area = peers[blockIdx.x];
    int64_t *dst = reinterpret_cast<int64_t *>(area) + (int64_t)parity * n_elems;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = __ldcg(beta + i);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        uint32_t *flag = reinterpret_cast<uint32_t *>(area + xchg_flag_offset(n_elems)) + parity * LPK_XCHG_FLAG_STRIDE + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
    }
}

// ------------------------------------------------------------------ one day from the template
static inline int32_t *row2(int32_t *base, int t, const lpk_run &R) { return base ? base + (int64_t)t * R.tick.n_nodes : R.rows.sink; }
static inline int32_t *row3(int32_t *base, int t, const lpk_run &R) {
    return base ? base + (int64_t)t * R.tick.n_nodes * R.tick.n_strains : R.rows.sink;
}

static int check_run(const lpk_run *run, const lpk_day *days, int n_days) {
    REQUIRE(run && days && n_days >= 0, "run_days null");
    const lpk_run &R = *run;
    REQUIRE(R.tick.n_nodes > 0 && R.tick.n_strains >= 1 && R.tick.n_strains <= LPK_MAX_STRAINS, "run_days sizes");
    REQUIRE(R.rows.sink && R.rows.S && R.rows.E && R.rows.I && R.rows.R && R.rows.E_by_strain && R.rows.I_by_strain &&
                R.rows.new_exposed && R.rows.new_exposed_by_strain && R.rows.new_potentially_paralyzed && R.rows.new_paralyzed,
            "run_days results arrays of DiseaseState_ABM / Transmission_ABM");
    REQUIRE(R.work_counters && R.zero_pop, "run_days scratch");
    REQUIRE(R.births.capacity == 0 || (R.rows.pop && R.rows.births && R.rows.deaths), "run_days vital-dynamics rows");
    for (int d = 0; d < n_days; ++d) {
        REQUIRE(days[d].tick >= 1, "run_days tick (tick 0 only logs)");
        REQUIRE(d == 0 || days[d].tick == days[d - 1].tick + 1, "run_days consecutive ticks");
        REQUIRE((days[d].flags & ~(LPK_F_DEATHS | LPK_F_RI | LPK_F_SIA)) == 0, "run_days day flags");
        REQUIRE(!(days[d].flags & LPK_F_DEATHS) || R.births.capacity != 0, "run_days: a vital-dynamics day needs the births template");
    }
    return LPK_OK;
}

static void fill_pass(const lpk_run &R, const lpk_day &D, lpk_births_args &B, lpk_tick_args &A, bool first_of_call) {
    const int t = D.tick, tp = t > 0 ? t - 1 : 0;
    A = R.tick;
    A.tick = t;
    A.flags = LPK_F_STAGES | (R.pending ? LPK_F_PENDING : 0u) | D.flags;
    A.new_exposed_prev = row2(R.rows.new_exposed, tp, R);
    A.new_exposed_by_strain_prev = row3(R.rows.new_exposed_by_strain, tp, R);
    A.new_potential = row2(R.rows.new_potentially_paralyzed, t, R);
    A.new_paralyzed = row2(R.rows.new_paralyzed, t, R);
    A.ri_lazy_k = R.ri_lazy_k;
    if (D.flags & (LPK_F_RI | LPK_F_SIA)) {
        A.new_exposed = row2(R.rows.new_exposed, t, R);
        A.new_exposed_by_strain = row3(R.rows.new_exposed_by_strain, t, R);
    }
    if (D.flags & LPK_F_RI) {
        A.ri_vaccinated = row2(R.rows.ri_vaccinated, t, R);
        A.ri_protected = row2(R.rows.ri_protected, t, R);
        A.ipv_vaccinated = row2(R.rows.ipv_vaccinated, t, R);
        A.ri_new_exposed_by_strain = row3(R.rows.ri_new_exposed_by_strain, t, R);
    }
    if (D.flags & LPK_F_SIA) {
        A.sia_targeted = D.sia_targeted;
        A.sia_vx_eff = D.sia_vx_eff;
        A.sia_min_age = D.sia_min_age; A.sia_max_age = D.sia_max_age; A.sia_strain = D.sia_strain; A.sia_event_idx = 0u;
        A.sia_vaccinated = row2(R.rows.sia_vaccinated, t, R);
        A.sia_protected = row2(R.rows.sia_protected, t, R);
        A.sia_new_exposed_by_strain = row3(R.rows.sia_new_exposed_by_strain, t, R);
    }
    // the passes alternate between the two work counters and each zeroes the other one for its successor; the first day
    // of a call cannot know who ran last, so day_pass puts a memset of its own counter in front of it
    const uint32_t par = (uint32_t)t & 1u;
    A.work_counter = R.work_counters + par;
    A.work_counter_next = R.work_counters + (par ^ 1u);
    (void)first_of_call;
    if (D.flags & LPK_F_DEATHS) {
        B = R.births;
        B.tick = t;
        B.pop_prev = row2(R.rows.pop, t - 1, R);
        B.births_row = row2(R.rows.births, t, R);
        B.ri_lazy_k = R.ri_lazy_k;
    }
}

static void fill_node(const lpk_run &R, const lpk_day &D, const lpk_tick_args &A, lpk_node_args &N) {
    const int t = D.tick, tp = t > 0 ? t - 1 : 0;
    N = R.node;
    N.flags = A.flags | (R.rowsums_valid ? LPK_F_ROWSUMS : 0u);
    N.tick = t;
    N.beta_seasonality = D.beta_seasonality;
    if (R.births.capacity != 0) {  // pop[t] = pop[t-1] (+ births[t] - deaths on vital-dynamics ticks), model.py:1694, 1751-1755
        N.pop_prev = row2(R.rows.pop, t - 1, R);
        N.pop = row2(R.rows.pop, t, R);
        if (D.flags & LPK_F_DEATHS) { N.births_row = row2(R.rows.births, t, R); N.deaths_row = row2(R.rows.deaths, t, R); }
        else { N.births_row = nullptr; N.deaths_row = nullptr; }
    } else {  // nobody maintains results.pop: rows after 0 stay as they are, the rate is divided by max(pop[t], 1)
        N.pop_prev = R.rows.pop ? R.rows.pop + (int64_t)t * R.tick.n_nodes : R.zero_pop;
        N.pop = nullptr; N.births_row = nullptr; N.deaths_row = nullptr;
    }
    N.new_potential = A.new_potential;
    N.new_paralyzed = A.new_paralyzed;
    N.potp_row = row2(R.rows.potentially_paralyzed, t, R);
    N.p_row = row2(R.rows.paralyzed, t, R);
    N.E_by_strain_prev = row3(R.rows.E_by_strain, tp, R);
    N.I_by_strain_prev = row3(R.rows.I_by_strain, tp, R);
    N.E_prev = row2(R.rows.E, tp, R);
    N.I_prev = row2(R.rows.I, tp, R);
    N.S_prev = row2(R.rows.S, tp, R);
    N.R_prev = row2(R.rows.R, tp, R);
    N.any_cases = R.any_cases ? R.any_cases + t : nullptr;
}

// ev: optional events {before births, before the pass, after the pass}
static int day_pass(lpk_run &R, const lpk_day &D, lpk_tick_args &A, bool first_of_call, int64_t *launches, cudaStream_t st,
                    cudaEvent_t *ev = nullptr) {
    lpk_births_args B;
    fill_pass(R, D, B, A, first_of_call);
    if (ev) CUDA_TRY(cudaEventRecord(ev[0], st), "run_days event");
    if (D.flags & LPK_F_DEATHS) {  // births first: the cohort takes part in this tick's tally (the reference runs VitalDynamics first)
        const int rc = lpk_vd_births(&B, st);
        if (rc != LPK_OK) return rc;
        if (launches) *launches += B.tile_node ? 3 : 2;
    }
    if (ev) CUDA_TRY(cudaEventRecord(ev[1], st), "run_days event");
    if (first_of_call) CUDA_TRY(cudaMemsetAsync(A.work_counter, 0, sizeof(uint32_t), st), "run_days work counter");
    const int rc = lpk_tick_pass(&R.people, &A, st);
    if (rc != LPK_OK) return rc;
    if (ev) CUDA_TRY(cudaEventRecord(ev[2], st), "run_days event");
    if (launches) *launches += 1;
    return LPK_OK;
}

static int day_node(lpk_run &R, const lpk_day &D, const lpk_tick_args &A, const int64_t *beta_all, int64_t *launches, cudaStream_t st) {
    lpk_node_args N;
    fill_node(R, D, A, N);
    if (beta_all) N.beta_fx = beta_all;
    const int rc = lpk_tick_node(&N, st);
    if (rc != LPK_OK) return rc;
    // k_tx_node_math (+ k_tick_epilogue on a network's first tick, + k_node_matvec_partial above 1024 nodes)
    if (launches) *launches += 1 + (R.rowsums_valid ? 0 : 1) + (N.n_nodes > 1024 ? 1 : 0);
    R.rowsums_valid = 1;
    R.pending = 1;
    if (D.flags & LPK_F_RI) R.ri_lazy_k += 1;
    return LPK_OK;
}

extern "C" int lpk_run_day_pass(lpk_run *run, const lpk_day *day, int64_t *launches, void *stream) {
    const int rc = check_run(run, day, 1);
    if (rc != LPK_OK) return rc;
    lpk_tick_args A;
    return day_pass(*run, *day, A, true, launches, as_stream(stream));
}

extern "C" int lpk_run_day_node(lpk_run *run, const lpk_day *day, const int64_t *beta_all, int64_t *launches, void *stream) {
    const int rc = check_run(run, day, 1);
    if (rc != LPK_OK) return rc;
    lpk_births_args B;
    lpk_tick_args A;
    fill_pass(*run, *day, B, A, true);  // the node kernels share the day's flags and rows with the pass
    return day_node(*run, *day, A, beta_all, launches, as_stream(stream));
}

static int run_span(lpk_run &R, const lpk_day *days, int n_days, cudaEvent_t *ev, int64_t *launches, cudaStream_t st) {
    lpk_xchg *x = R.xchg;
    for (int d = 0; d < n_days; ++d) {
        const lpk_day &D = days[d];
        lpk_tick_args A;
        int rc = day_pass(R, D, A, d == 0, launches, st, ev ? ev + 4 * d : nullptr);
        if (rc != LPK_OK) return rc;
        const int64_t *beta_all = nullptr;
        if (x) {
            R.seq += 1u;
            const int parity = (int)(R.seq & 1u);
            const int64_t ns = R.tick.n_strains;
            const int lo = R.node.node_hi > 0 ? R.node.node_lo : 0, hi = R.node.node_hi > 0 ? R.node.node_hi : R.tick.n_nodes;
            k_xchg_push<<<x->world, 256, 0, st>>>(R.tick.beta_fx, lo * ns, hi * ns, x->d_peer, x->n_elems, parity, x->rank, R.seq);
            CUDA_TRY(cudaGetLastError(), "run_days tally push");
            if (launches) *launches += 1;
            beta_all = reinterpret_cast<const int64_t *>(x->base) + (int64_t)parity * x->n_elems;
            R.node.xchg_flags = reinterpret_cast<const uint32_t *>(x->base + xchg_flag_offset(x->n_elems)) + parity * LPK_XCHG_FLAG_STRIDE;
            R.node.xchg_world = x->world;
            R.node.xchg_seq = R.seq;
        }
        rc = day_node(R, D, A, beta_all, launches, st);
        if (rc != LPK_OK) return rc;
        if (ev) CUDA_TRY(cudaEventRecord(ev[4 * d + 3], st), "run_days event");
    }
    return LPK_OK;
}

// executable graphs that may still be running: destroyed when their completion event has fired
struct InFlight {
    cudaGraphExec_t exec = nullptr;
    cudaEvent_t done = nullptr;
};
static thread_local std::vector<InFlight> g_in_flight;
static void reap_graphs(bool wait) {
    for (size_t i = 0; i < g_in_flight.size();) {
        InFlight &g = g_in_flight[i];
        if (wait) cudaEventSynchronize(g.done);
        if (wait || cudaEventQuery(g.done) == cudaSuccess) {
            cudaGraphExecDestroy(g.exec);
            cudaEventDestroy(g.done);
            g_in_flight.erase(g_in_flight.begin() + (long)i);
        } else {
            ++i;
        }
    }
}

extern "C" int lpk_run_days(lpk_run *run, const lpk_day *days, int32_t n_days, float *ms, int64_t *launches, void *stream) {
    int rc = check_run(run, days, n_days);
    if (rc != LPK_OK) return rc;
    if (n_days == 0) return LPK_OK;
    lpk_run &R = *run;
    REQUIRE(!R.xchg || (R.xchg->connected && R.xchg->n_elems == (int64_t)R.tick.n_nodes * R.tick.n_strains), "run_days tally exchange");
    cudaStream_t st = as_stream(stream);
    if (ms) {
        std::vector<cudaEvent_t> ev(4 * (size_t)n_days);
        for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e), "run_days event");
        rc = run_span(R, days, n_days, ev.data(), launches, st);
        if (rc == LPK_OK) {
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc = lpk_set_cuda_err(e, "run_days sync");
            for (int d = 0; d < n_days && rc == LPK_OK; ++d)
                for (int k = 0; k < 3; ++k) cudaEventElapsedTime(&ms[3 * d + k], ev[4 * d + k], ev[4 * d + k + 1]);
        }
        for (auto &e : ev) cudaEventDestroy(e);
        return rc;
    }
    if (!R.graph) return run_span(R, days, n_days, nullptr, launches, st);
    // The span as one graph launch: the kernels and the work-counter memset become graph nodes, the host pays one launch.
    // The very first day of a run goes out directly (one-off allocations and function attributes are not capturable).
    if (!R.rowsums_valid) {
        rc = run_span(R, days, 1, nullptr, launches, st);
        if (rc != LPK_OK || n_days == 1) return rc;
        ++days; --n_days;
    }
    reap_graphs(false);
    cudaGraph_t graph = nullptr;
    InFlight g;
    // captured on a private stream (the caller's may be the legacy default stream, which cannot capture), launched on the caller's
    static thread_local cudaStream_t cap = nullptr;
    if (!cap) CUDA_TRY(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking), "run_days capture stream");
    CUDA_TRY(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal), "run_days capture");
    rc = run_span(R, days, n_days, nullptr, launches, cap);
    cudaError_t e = cudaStreamEndCapture(cap, &graph);
    if (rc != LPK_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return lpk_set_cuda_err(e, "run_days end capture");
    e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return lpk_set_cuda_err(e, "run_days instantiate");
    e = cudaGraphLaunch(g.exec, st);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(g.done, st);
    if (e != cudaSuccess) { cudaGraphExecDestroy(g.exec); return lpk_set_cuda_err(e, "run_days graph launch"); }
    g_in_flight.push_back(g);  // destroyed once the launch has completed (next call / lpk_run_release)
    return LPK_OK;
}

extern "C" int lpk_run_release(void) {
    reap_graphs(true);
    return LPK_OK;
}
