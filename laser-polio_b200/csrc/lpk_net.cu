// lpk_net.cu -- the infection-migration network is built in HBM.
//
// Reference Transmission_ABM._initialize_common (model.py:1216-1258): an all-pairs Haversine distance matrix filled by a Python
// double loop (16 M distance() calls at 5672 nodes; "This network calc is a little slow too..."), then laser-core's
// gravity / radiation model and row_normalizer.  All of it is O(nodes^2) float64 arithmetic with no dependence between
// entries, so here it is four small kernels over the [nodes, nodes] matrix; the matrix never visits the host unless asked for.
//   k_net_haversine   d_ij = 2 R asin(sqrt(sin^2(dlat / 2) + cos(lat_i) cos(lat_j) sin^2(dlon / 2))), R = 6371 km;
//                     coincident nodes (d == 0, i != j) get epsilon = 1 km (model.py:1232-1238); zero diagonal
//   k_net_gravity     k p_i^a p_j^b d_ij^-c / (sum p)^c, zero diagonal (model.py:1243-1251)
//   k_net_radiation   one block per origin i: destinations sorted by distance in shared memory (bitonic sort of (d, j) keys,
//                     ties in distance share one radius), s_ij = population strictly closer than the radius ring around i
//                     (excluding i unless include_home), T_ij = k p_i p_j / ((p_i + s_ij)(p_i + p_j + s_ij)) (model.py:1252-1254)
//   k_net_row_normalize  rows whose sum exceeds max_rowsum are rescaled to it (model.py:1258)
#include "lpk_host.cuh"

namespace {

__global__ void k_net_haversine(const double *__restrict__ lat, const double *__restrict__ lon, int n, double epsilon,
                                double *__restrict__ dist) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    const double rad = 0.017453292519943295;  // pi / 180
    const double la1 = __dmul_rn(lat[i], rad), la2 = __dmul_rn(lat[j], rad), lo1 = __dmul_rn(lon[i], rad), lo2 = __dmul_rn(lon[j], rad);
    const double s1 = sin(__dmul_rn(la2 - la1, 0.5)), s2 = sin(__dmul_rn(lo2 - lo1, 0.5));
    const double a = __dadd_rn(__dmul_rn(s1, s1), __dmul_rn(__dmul_rn(cos(la1), cos(la2)), __dmul_rn(s2, s2)));
    double d = __dmul_rn(12742.0, asin(sqrt(a)));
    if (i == j) d = 0.0;
    else if (d == 0.0) d = epsilon;
    dist[(int64_t)i * n + j] = d;
}

__global__ void k_net_gravity(const double *__restrict__ pops, const double *__restrict__ dist, int n, double k, double a, double b, double c,
                              double norm, double *__restrict__ net) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    double v = 0.0;
    if (i != j) {
        v = __dmul_rn(__dmul_rn(__dmul_rn(k, pow(pops[i], a)), pow(pops[j], b)), pow(dist[(int64_t)i * n + j], -c));
        v = __ddiv_rn(v, norm);
    }
    net[(int64_t)i * n + j] = v;
}

// one block per origin; shared: key[m] (distance bits << 32 would lose precision, so two arrays), idx[m], cum[m]; m = pow2 >= n
__global__ void __launch_bounds__(1024) k_net_radiation(const double *__restrict__ pops, const double *__restrict__ dist, int n, int m, double k,
                                                        int include_home, double *__restrict__ net) {
    extern __shared__ unsigned char raw[];
    double *sd = reinterpret_cast<double *>(raw);         // [m] distances (sorted in place)
    double *sc = sd + m;                                  // [m] populations in sorted order -> inclusive cumulative sums
    int *si = reinterpret_cast<int *>(sc + m);            // [m] destination index
    const int i = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    for (int j = tid; j < m; j += nt) {
        sd[j] = j < n ? dist[(int64_t)i * n + j] : __longlong_as_double(0x7FF0000000000000ll);  // +inf pads
        si[j] = j;
    }
    __syncthreads();
    // bitonic sort by (distance, index): the index breaks ties the way a stable argsort does
    for (int size = 2; size <= m; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (m >> 1); t += nt) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const double dl = sd[lo], dh = sd[hi];
                const int il = si[lo], ih = si[hi];
                const bool gt = dl > dh || (dl == dh && il > ih);
                if (gt == up) { sd[lo] = dh; sd[hi] = dl; si[lo] = ih; si[hi] = il; }
            }
            __syncthreads();
        }
    }
    for (int j = tid; j < m; j += nt) sc[j] = (j < n) ? pops[si[j]] : 0.0;
    __syncthreads();
    // inclusive scan in sorted order, serial per chunk + chunk offsets (n <= 8192: 8 elements per thread at most)
    __shared__ double s_part[1024];
    const int per = (m + nt - 1) / nt, b0 = tid * per;
    double run = 0.0;
    for (int q = 0; q < per && b0 + q < m; ++q) { run = __dadd_rn(run, sc[b0 + q]); sc[b0 + q] = run; }
    s_part[tid] = run;
    __syncthreads();
    if (tid == 0) { double acc = 0.0; for (int t = 0; t < nt; ++t) { const double v = s_part[t]; s_part[t] = acc; acc = __dadd_rn(acc, v); } }
    __syncthreads();
    const double off = s_part[tid];
    for (int q = 0; q < per && b0 + q < m; ++q) sc[b0 + q] = __dadd_rn(sc[b0 + q], off);
    __syncthreads();
    const double pi = pops[i];
    for (int j = tid; j < n; j += nt) {
        // last position sharing this distance (ties share the radius): binary search for the first distance > sd[j]
        const double dj = sd[j];
        int lo = j + 1, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sd[mid] <= dj) lo = mid + 1; else hi = mid; }
        const int dest = si[j];
        const double pj = pops[dest];
        double within = __dadd_rn(sc[lo - 1], -pj);          // population within the radius, excluding j itself
        if (!include_home) within = __dadd_rn(within, -pi);
        const double s = within > 0.0 ? within : 0.0;
        const double val = __ddiv_rn(__dmul_rn(__dmul_rn(k, pi), pj), __dmul_rn(__dadd_rn(pi, s), __dadd_rn(__dadd_rn(pi, pj), s)));
        net[(int64_t)i * n + dest] = dest == i ? 0.0 : val;
    }
}

__global__ void __launch_bounds__(256) k_net_row_normalize(double *__restrict__ net, int n, double max_rowsum) {
    __shared__ double s_sum[256];
    const int i = blockIdx.x, tid = threadIdx.x;
    double acc = 0.0;
    for (int j = tid; j < n; j += 256) acc += net[(int64_t)i * n + j];
    s_sum[tid] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (tid < o) s_sum[tid] += s_sum[tid + o]; __syncthreads(); }
    const double rs = s_sum[0];
    if (!(rs > max_rowsum)) return;
    const double f = __ddiv_rn(max_rowsum, rs);
    for (int j = tid; j < n; j += 256) net[(int64_t)i * n + j] = __dmul_rn(net[(int64_t)i * n + j], f);
}

}  // namespace

extern "C" int lpk_net_haversine(const double *lat, const double *lon, int32_t n, double epsilon, double *dist, void *stream) {
    REQUIRE(lat && lon && dist && n > 0, "net_haversine");
    k_net_haversine<<<dim3((n + 255) / 256, n), 256, 0, as_stream(stream)>>>(lat, lon, n, epsilon, dist);
    CUDA_TRY(cudaGetLastError(), "lpk_net_haversine");
    return LPK_OK;
}
extern "C" int lpk_net_gravity(const double *pops, const double *dist, int32_t n, double k, double a, double b, double c, double norm,
                               double *net, void *stream) {
    REQUIRE(pops && dist && net && n > 0 && norm != 0.0, "net_gravity");
    k_net_gravity<<<dim3((n + 255) / 256, n), 256, 0, as_stream(stream)>>>(pops, dist, n, k, a, b, c, norm, net);
    CUDA_TRY(cudaGetLastError(), "lpk_net_gravity");
    return LPK_OK;
}
extern "C" int lpk_net_radiation(const double *pops, const double *dist, int32_t n, double k, int32_t include_home, double *net, void *stream) {
    REQUIRE(pops && dist && net && n > 0 && n <= 8192, "net_radiation (at most 8192 nodes: one origin's row is sorted in shared memory)");
    int m = 1;
    while (m < n) m <<= 1;
    if (m < 2) m = 2;
    const size_t smem = (size_t)m * (2 * sizeof(double) + sizeof(int));
    static bool configured = false;
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(k_net_radiation, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 20), "net_radiation smem");
        configured = true;
    }
    k_net_radiation<<<n, 1024, smem, as_stream(stream)>>>(pops, dist, n, m, k, include_home, net);
    CUDA_TRY(cudaGetLastError(), "lpk_net_radiation");
    return LPK_OK;
}
extern "C" int lpk_net_row_normalize(double *net, int32_t n, double max_rowsum, void *stream) {
    REQUIRE(net && n > 0, "net_row_normalize");
    k_net_row_normalize<<<n, 256, 0, as_stream(stream)>>>(net, n, max_rowsum);
    CUDA_TRY(cudaGetLastError(), "lpk_net_row_normalize");
    return LPK_OK;
}
