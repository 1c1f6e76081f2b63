// K4: Cluster-GCN sub-graph extraction (partition -> node remap -> induced-edge compaction).
//
// Replaces PyG ClusterData._partition / _permute_data / __getitem__ and the reference's
// `cluster.n_id[cluster.edge_index]` remap back to GLOBAL ids
// (/root/reference/data/dataset_handler.py:273-282).  Net effect per part p:
//     {(r,c) in edges : cluster[r] == cluster[c] == p}   ordered by (inv[r], inv[c]),
// inv = inverse permutation of the STABLE sort of `cluster` (so inside a part nodes keep ascending
// id).  All integer work, bit-exact given `cluster`:
//   1. stable radix sort of (cluster, node)        -> node_perm, inv, first node rank of each part
//   2. per edge: key = inv[r]*N + inv[c] if both ends share a part, else the sentinel N*N
//   3. one keys-only 64-bit radix sort (the payload is decodable from the key)
//   4. decode key -> (node_perm[key / N], node_perm[key % N]); part_ptr by binary search.
// HBM-bound streaming/sort passes over the edge list; no host round trip, no Python loop over parts.
#include "common.cuh"
#include <cub/cub.cuh>

namespace lgcn {

__global__ void cluster_convert_kernel(const int64_t *__restrict__ cluster, int64_t N, int64_t P,
                                       int *__restrict__ c32, int *__restrict__ ids, int *bad) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int64_t c = cluster[n];
    if (c < 0 || c >= P) atomicAdd(bad, 1);
    c32[n] = (int)(c < 0 ? 0 : (c >= P ? P - 1 : c));
    ids[n] = (int)n;
}

__global__ void invert_kernel(const int *__restrict__ perm, int64_t N, int *__restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) inv[perm[i]] = (int)i;
}

__global__ void part_start_kernel(const int *__restrict__ sorted_cluster, int64_t N, int64_t P,
                                  long long *__restrict__ start) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > P) return;
    int64_t lo = 0, hi = N;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_cluster[mid] < (int)p) lo = mid + 1; else hi = mid;
    }
    start[p] = lo;
}

__global__ void edge_key_kernel(const int64_t *__restrict__ ei, int64_t E, int64_t N, const int *__restrict__ c32,
                                const int *__restrict__ inv, unsigned long long *__restrict__ keys, int *bad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t r = ei[e], c = ei[E + e];
    const unsigned long long sentinel = (unsigned long long)N * (unsigned long long)N;
    if (r < 0 || r >= N || c < 0 || c >= N) { atomicAdd(bad, 1); keys[e] = sentinel; return; }
    keys[e] = c32[r] == c32[c] ? (unsigned long long)inv[r] * (unsigned long long)N + (unsigned long long)inv[c]
                               : sentinel;
}

__global__ void decode_kernel(const unsigned long long *__restrict__ sorted, int64_t E, int64_t N,
                              const int *__restrict__ perm, int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    const unsigned long long k = sorted[i], n = (unsigned long long)N;
    if (k >= n * n) return;                     // dropped inter-cluster edge
    out[i] = perm[k / n];
    out[E + i] = perm[k % n];
}

__global__ void part_ptr_kernel(const unsigned long long *__restrict__ sorted, int64_t E, int64_t N, int64_t P,
                                const long long *__restrict__ start, int64_t *__restrict__ part_ptr) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > P) return;
    const unsigned long long bound = (unsigned long long)start[p] * (unsigned long long)N;
    int64_t lo = 0, hi = E;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted[mid] < bound) lo = mid + 1; else hi = mid;
    }
    part_ptr[p] = lo;
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static int bits64(unsigned long long n) { int b = 1; while (b < 64 && (1ull << b) <= n) ++b; return b; }

struct ClusterWs {
    int *c32, *ids, *c_sorted, *perm, *inv, *bad;
    long long *start;
    unsigned long long *keys, *keys_sorted;
    void *cub_temp;
    size_t cub_bytes, total;
};

static ClusterWs cluster_carve(void *base, int64_t N, int64_t E, int64_t P) {
    ClusterWs w{};
    char *p = (char *)base;
    auto take = [&](size_t bytes) { char *q = p; p += al256(bytes); return q; };
    const size_t nb = sizeof(int) * (size_t)(N > 0 ? N : 1), eb = sizeof(unsigned long long) * (size_t)(E > 0 ? E : 1);
    w.c32 = (int *)take(nb); w.ids = (int *)take(nb); w.c_sorted = (int *)take(nb);
    w.perm = (int *)take(nb); w.inv = (int *)take(nb); w.bad = (int *)take(256);
    w.start = (long long *)take(sizeof(long long) * (size_t)(P + 1));
    w.keys = (unsigned long long *)take(eb); w.keys_sorted = (unsigned long long *)take(eb);
    size_t a = 0, b = 0;
    int *k = nullptr; unsigned long long *k64 = nullptr;
    cub::DeviceRadixSort::SortPairs(nullptr, a, k, k, k, k, (int)N, 0, 32);
    cub::DeviceRadixSort::SortKeys(nullptr, b, k64, k64, (int)E, 0, 64);
    w.cub_bytes = al256(a > b ? a : b) + 256;
    w.cub_temp = take(w.cub_bytes);
    w.total = (size_t)(p - (char *)base);
    return w;
}


// ---- to_undirected (PyG 2.4.0 utils/undirected.py + coalesce; /root/reference/data/dataset_handler.py:141) ----
// Both directions of every edge, sorted by (row, col), duplicates dropped:
//   1. keys[i] = r*N + c, keys[E + i] = c*N + r       2. one keys-only 64-bit radix sort
//   3. head flags (key != predecessor) scanned to output slots      4. decode the kept keys.
// Integer work, bit-exact; the count of distinct edges is data-dependent, so the decode kernel reads it from the
// scan's last cell and writes rows at [0,count) and columns at [count, 2*count) of the flat output.
__global__ void undirected_key_kernel(const int64_t *__restrict__ ei, int64_t E, int64_t N,
                                      unsigned long long *__restrict__ keys, int *bad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t r = ei[e], c = ei[E + e];
    if (r < 0 || r >= N || c < 0 || c >= N) { atomicAdd(bad, 1); r = c = 0; }
    keys[e] = (unsigned long long)r * (unsigned long long)N + (unsigned long long)c;
    keys[E + e] = (unsigned long long)c * (unsigned long long)N + (unsigned long long)r;
}

struct HeadFlag {
    const unsigned long long *k;
    int64_t n;
    __host__ __device__ int operator()(int i) const { return i < n && (i == 0 || k[i] != k[i - 1]) ? 1 : 0; }
};

__global__ void undirected_decode_kernel(const unsigned long long *__restrict__ sorted, const int *__restrict__ slot,
                                         int64_t M, int64_t N, int64_t *__restrict__ out, int64_t *count_dev) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t count = slot[M];
    if (i == 0) *count_dev = count;
    if (slot[i + 1] == slot[i]) return;                  // duplicate of its predecessor
    const unsigned long long k = sorted[i], n = (unsigned long long)N;
    out[slot[i]] = (int64_t)(k / n);
    out[count + slot[i]] = (int64_t)(k % n);
}

struct UndirWs {
    unsigned long long *keys, *keys_sorted;
    int *slot, *bad;
    int64_t *count;
    void *cub_temp;
    size_t cub_bytes, total;
};

static UndirWs undir_carve(void *base, int64_t E) {
    UndirWs w{};
    char *p = (char *)base;
    auto take = [&](size_t bytes) { char *q = p; p += al256(bytes); return q; };
    const int64_t M = 2 * E;
    const size_t kb = sizeof(unsigned long long) * (size_t)(M > 0 ? M : 1);
    w.keys = (unsigned long long *)take(kb); w.keys_sorted = (unsigned long long *)take(kb);
    w.slot = (int *)take(sizeof(int) * (size_t)(M + 1));
    w.bad = (int *)take(256); w.count = (int64_t *)take(256);
    size_t a = 0, b = 0;
    unsigned long long *k64 = nullptr; int *k = nullptr;
    cub::DeviceRadixSort::SortKeys(nullptr, a, k64, k64, (int)M, 0, 64);
    cub::CountingInputIterator<int> idx(0);
    cub::TransformInputIterator<int, HeadFlag, cub::CountingInputIterator<int>> flags(idx, HeadFlag{nullptr, 0});
    cub::DeviceScan::ExclusiveSum(nullptr, b, flags, k, (int)(M + 1));
    w.cub_bytes = al256(a > b ? a : b) + 256;
    w.cub_temp = take(w.cub_bytes);
    w.total = (size_t)(p - (char *)base);
    return w;
}

}  // namespace lgcn

extern "C" size_t lgcn_cluster_extract_workspace_bytes(int64_t N, int64_t E, int64_t P) {
    if (N <= 0 || E < 0 || P <= 0 || N >= INT32_MAX || E >= INT32_MAX) return 0;
    return lgcn::cluster_carve(nullptr, N, E, P).total;
}

extern "C" int lgcn_cluster_extract(const int64_t *edge_index, int64_t E, int64_t N, const int64_t *cluster,
                                    int64_t P, int64_t *out_edges, int64_t *part_ptr, void *workspace,
                                    size_t workspace_bytes, void *stream) {
    using namespace lgcn;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(cluster && part_ptr && workspace && (E == 0 || (edge_index && out_edges)), LGCN_E_INVALID,
                 "cluster_extract: null argument");
    LGCN_REQUIRE(N > 0 && P > 0 && E >= 0, LGCN_E_INVALID, "cluster_extract: bad sizes");
    LGCN_REQUIRE(N < INT32_MAX && E < INT32_MAX, LGCN_E_RANGE, "cluster_extract: N/E exceed int32");
    ClusterWs w = cluster_carve(workspace, N, E, P);
    LGCN_REQUIRE(workspace_bytes >= w.total, LGCN_E_WORKSPACE, "cluster_extract: workspace %zu < %zu",
                 workspace_bytes, w.total);
    const int T = 256;
    LGCN_CUDA(cudaMemsetAsync(w.bad, 0, sizeof(int), st));
    cluster_convert_kernel<<<cdiv(N, T), T, 0, st>>>(cluster, N, P, w.c32, w.ids, w.bad);
    LGCN_LAUNCH_CHECK();
    size_t tb = w.cub_bytes;
    LGCN_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.c32, w.c_sorted, w.ids, w.perm, (int)N, 0,
                                              bits64((unsigned long long)P), st));
    invert_kernel<<<cdiv(N, T), T, 0, st>>>(w.perm, N, w.inv);
    LGCN_LAUNCH_CHECK();
    part_start_kernel<<<cdiv(P + 1, T), T, 0, st>>>(w.c_sorted, N, P, w.start);
    LGCN_LAUNCH_CHECK();
    if (E > 0) {
        edge_key_kernel<<<cdiv(E, T), T, 0, st>>>(edge_index, E, N, w.c32, w.inv, w.keys, w.bad);
        LGCN_LAUNCH_CHECK();
        tb = w.cub_bytes;
        LGCN_CUDA(cub::DeviceRadixSort::SortKeys(w.cub_temp, tb, w.keys, w.keys_sorted, (int)E, 0,
                                                 bits64((unsigned long long)N * (unsigned long long)N), st));
        decode_kernel<<<cdiv(E, T), T, 0, st>>>(w.keys_sorted, E, N, w.perm, out_edges);
        LGCN_LAUNCH_CHECK();
    }
    part_ptr_kernel<<<cdiv(P + 1, T), T, 0, st>>>(w.keys_sorted, E, N, P, w.start, part_ptr);
    LGCN_LAUNCH_CHECK();
    int bad = 0;
    LGCN_CUDA(cudaMemcpyAsync(&bad, w.bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    LGCN_CUDA(cudaStreamSynchronize(st));
    LGCN_REQUIRE(bad == 0, LGCN_E_INVALID, "cluster_extract: %d ids outside [0,N) or parts outside [0,P)", bad);
    return LGCN_OK;
}

extern "C" size_t lgcn_to_undirected_workspace_bytes(int64_t E) {
    if (E < 0 || 2 * E + 1 >= INT32_MAX) return 0;
    return lgcn::undir_carve(nullptr, E).total;
}

extern "C" int lgcn_to_undirected(const int64_t *edge_index, int64_t E, int64_t N, int64_t *out_edges,
                                  int64_t *count_out, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace lgcn;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(count_out && workspace && (E == 0 || (edge_index && out_edges)), LGCN_E_INVALID,
                 "to_undirected: null argument");
    LGCN_REQUIRE(N > 0 && E >= 0, LGCN_E_INVALID, "to_undirected: bad sizes");
    LGCN_REQUIRE(N < INT32_MAX && 2 * E + 1 < INT32_MAX, LGCN_E_RANGE, "to_undirected: N / 2E exceed int32");
    *count_out = 0;
    if (E == 0) return LGCN_OK;
    UndirWs w = undir_carve(workspace, E);
    LGCN_REQUIRE(workspace_bytes >= w.total, LGCN_E_WORKSPACE, "to_undirected: workspace %zu < %zu", workspace_bytes,
                 w.total);
    const int T = 256;
    const int64_t M = 2 * E;
    LGCN_CUDA(cudaMemsetAsync(w.bad, 0, sizeof(int), st));
    undirected_key_kernel<<<cdiv(E, T), T, 0, st>>>(edge_index, E, N, w.keys, w.bad);
    LGCN_LAUNCH_CHECK();
    size_t tb = w.cub_bytes;
    LGCN_CUDA(cub::DeviceRadixSort::SortKeys(w.cub_temp, tb, w.keys, w.keys_sorted, (int)M, 0,
                                             bits64((unsigned long long)N * (unsigned long long)N), st));
    cub::CountingInputIterator<int> idx(0);
    cub::TransformInputIterator<int, HeadFlag, cub::CountingInputIterator<int>> flags(idx, HeadFlag{w.keys_sorted, M});
    tb = w.cub_bytes;
    LGCN_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, flags, w.slot, (int)(M + 1), st));
    undirected_decode_kernel<<<cdiv(M, T), T, 0, st>>>(w.keys_sorted, w.slot, M, N, out_edges, w.count);
    LGCN_LAUNCH_CHECK();
    int bad = 0;
    int64_t count = 0;
    LGCN_CUDA(cudaMemcpyAsync(&bad, w.bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    LGCN_CUDA(cudaMemcpyAsync(&count, w.count, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    LGCN_CUDA(cudaStreamSynchronize(st));
    LGCN_REQUIRE(bad == 0, LGCN_E_INVALID, "to_undirected: %d node ids outside [0,N)", bad);
    *count_out = count;
    return LGCN_OK;
}
