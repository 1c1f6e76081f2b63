// K0: edge list -> CSR by target + CSR by source + in-degree normalisation + warp task lists.
//
// Replaces what PyG's gcn_norm recomputes in EVERY layer of every forward
// (/root/reference/models/light_gcn.py:33 -> gcn_conv.py::gcn_norm: deg = scatter_add(ones, col),
// dis = deg^-1/2 with inf -> 0) and the implicit COO traversal order of scatter_add_.  Built once
// per edge list and cached by the host side.
//
// Integer outputs are bit-exact and deterministic: both CSRs keep the ORIGINAL edge order inside a
// row (stable LSD radix sort on the 32-bit node id, CUB), degrees are pointer differences, the
// triplet id of an edge is its rank among edges with source < U (utils/helpers.py:98-99).
#include "common.cuh"
#include <cub/cub.cuh>

namespace lgcn {

enum { META_P = 0, META_BAD = 1, META_IN_TASKS, META_OUT_TASKS, META_IN_USER_TASKS, META_OUT_USER_TASKS,
       META_IN_SLOTS, META_OUT_SLOTS, META_ACTIVE, META_UNSORTED, META_COUNT = 16 };

struct IsUser {
    int num_users;
    __host__ __device__ int operator()(int r) const { return r < num_users ? 1 : 0; }
};

__global__ void convert_kernel(const int64_t *__restrict__ ei, int64_t E, int64_t N, int64_t U,
                               int *__restrict__ row32, int *__restrict__ col32, int *__restrict__ eid,
                               long long *meta) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t r = ei[e], c = ei[E + e];
    const bool bad = r < 0 || r >= N || c < 0 || c >= N || ((r < U) == (c < U));
    if (bad) atomicAdd((unsigned long long *)(meta + META_BAD), 1ull);
    row32[e] = bad ? 0 : (int)r;
    col32[e] = bad ? 0 : (int)c;
    eid[e] = (int)e;
}

__global__ void count_triplets_kernel(const int *row32, const int *trip, int64_t E, int U, long long *meta) {
    meta[META_P] = E > 0 ? (long long)trip[E - 1] + (row32[E - 1] < U ? 1 : 0) : 0;
}

// nbr[s] = other endpoint of the s-th edge in sorted order; trip[s] = triplet id or -1
__global__ void fill_csr_kernel(const int *__restrict__ eid_sorted, const int *__restrict__ other32,
                                const int *__restrict__ row32, const int *__restrict__ trip_id, int64_t E,
                                int U, int *__restrict__ nbr, int *__restrict__ trip) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= E) return;
    const int e = eid_sorted[s];
    nbr[s] = other32[e];
    trip[s] = row32[e] < U ? trip_id[e] : -1;
}

// counts positions where two neighbouring entries of the same CSR row have descending neighbour ids
__global__ void unsorted_kernel(const int *__restrict__ keys_sorted, const int *__restrict__ nbr, int64_t E, long long *meta) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s + 1 < E && keys_sorted[s] == keys_sorted[s + 1] && nbr[s] > nbr[s + 1])
        atomicAdd((unsigned long long *)(meta + META_UNSORTED), 1ull);
}

// ptr[n] = first position whose sorted key is >= n (n = 0..N)
__global__ void ptr_kernel(const int *__restrict__ keys_sorted, int64_t E, int64_t N, int *__restrict__ ptr) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n > N) return;
    int64_t lo = 0, hi = E;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys_sorted[mid] < (int)n) lo = mid + 1; else hi = mid;
    }
    ptr[n] = (int)lo;
}

__global__ void node_kernel(const int *__restrict__ in_ptr, const int *__restrict__ out_ptr, int64_t N,
                            int split, float *__restrict__ dis, uint8_t *__restrict__ active,
                            int *__restrict__ cnt_in, int *__restrict__ slot_in, int *__restrict__ cnt_out,
                            int *__restrict__ slot_out, int *__restrict__ act32) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n > N) return;
    if (n == N) { cnt_in[n] = slot_in[n] = cnt_out[n] = slot_out[n] = act32[n] = 0; return; }
    const int din = in_ptr[n + 1] - in_ptr[n], dout = out_ptr[n + 1] - out_ptr[n];
    // deg.pow(-0.5) with inf -> 0 (gcn_norm); deg is an exact small integer in fp32
    dis[n] = din > 0 ? 1.0f / sqrtf((float)din) : 0.f;
    const int a = (din > 0 || dout > 0) ? 1 : 0;
    active[n] = (uint8_t)a;
    act32[n] = a;
    const int pin = (din + split - 1) / split, pout = (dout + split - 1) / split;
    cnt_in[n] = a ? max(1, pin) : 0;
    cnt_out[n] = a ? max(1, pout) : 0;
    slot_in[n] = pin > 1 ? pin : 0;
    slot_out[n] = pout > 1 ? pout : 0;
}

__global__ void task_kernel(const int *__restrict__ ptr, const int *__restrict__ in_ptr,
                            const int *__restrict__ out_ptr, const int *__restrict__ task_off,
                            const int *__restrict__ slot_off, int64_t N, int split, lgcn_task *__restrict__ tasks) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int cnt = task_off[n + 1] - task_off[n];
    if (cnt == 0) return;
    const int b = ptr[n], e = ptr[n + 1];
    const int din = in_ptr[n + 1] - in_ptr[n], dout = out_ptr[n + 1] - out_ptr[n];
    lgcn_task *t = tasks + task_off[n];
    if (cnt == 1) {
        t[0] = lgcn_task{(int)n, b, e, -1, 0, 1, din, dout};
        return;
    }
    const int s0 = slot_off[n];
    for (int i = 0; i < cnt; ++i) {
        const int tb = b + i * split;
        t[i] = lgcn_task{(int)n, tb, min(e, tb + split), s0 + i, i, cnt, din, dout};
    }
}

// active_list[rank of n among active nodes] = n   (act_off = exclusive scan of the 0/1 flags)
__global__ void active_list_kernel(const uint8_t *__restrict__ active, const int *__restrict__ act_off, int64_t N,
                                   int *__restrict__ list) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N && active[n]) list[act_off[n]] = (int)n;
}

__global__ void meta_kernel(const int *task_in, const int *task_out, const int *slot_in, const int *slot_out,
                            const int *act_off, int64_t N, int U, long long *meta) {
    meta[META_IN_TASKS] = task_in[N];
    meta[META_OUT_TASKS] = task_out[N];
    meta[META_IN_USER_TASKS] = task_in[U];
    meta[META_OUT_USER_TASKS] = task_out[U];
    meta[META_IN_SLOTS] = slot_in[N];
    meta[META_OUT_SLOTS] = slot_out[N];
    meta[META_ACTIVE] = act_off[N];
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int bits_for(int64_t n) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n) ++b;
    return b;
}

static size_t cub_temp_bytes(int64_t N, int64_t E) {
    size_t a = 0, b = 0, c = 0;
    int *k = nullptr;
    cub::DeviceRadixSort::SortPairs(nullptr, a, k, k, k, k, (int)E, 0, bits_for(N));
    cub::DeviceScan::ExclusiveSum(nullptr, b, k, k, (int)(N + 1));
    cub::TransformInputIterator<int, IsUser, const int *> it(k, IsUser{0});
    cub::DeviceScan::ExclusiveSum(nullptr, c, it, k, (int)E);
    return align256(a > b ? (a > c ? a : c) : (b > c ? b : c)) + 256;
}

struct BuildWs {
    int *row32, *col32, *eid, *trip_id, *keys_sorted, *eid_sorted;
    int *cnt_in, *slot_in, *cnt_out, *slot_out, *act32;      // [N+1] each, scanned in place
    long long *meta;
    void *cub_temp;
    size_t cub_bytes, total;
};

static BuildWs carve(void *base, int64_t N, int64_t E) {
    BuildWs w{};
    char *p = (char *)base;
    auto take = [&](size_t bytes) { char *q = p; p += align256(bytes); return q; };
    const size_t eb = sizeof(int) * (size_t)(E > 0 ? E : 1), nb = sizeof(int) * (size_t)(N + 1);
    w.row32 = (int *)take(eb); w.col32 = (int *)take(eb); w.eid = (int *)take(eb);
    w.trip_id = (int *)take(eb); w.keys_sorted = (int *)take(eb); w.eid_sorted = (int *)take(eb);
    w.cnt_in = (int *)take(nb); w.slot_in = (int *)take(nb); w.cnt_out = (int *)take(nb);
    w.slot_out = (int *)take(nb); w.act32 = (int *)take(nb);
    w.meta = (long long *)take(sizeof(long long) * META_COUNT);
    w.cub_bytes = cub_temp_bytes(N, E);
    w.cub_temp = take(w.cub_bytes);
    w.total = (size_t)(p - (char *)base);
    return w;
}

}  // namespace lgcn

extern "C" int lgcn_graph_sizes_query(int64_t N, int64_t E, lgcn_graph_sizes *out) {
    using namespace lgcn;
    LGCN_REQUIRE(out && N >= 0 && E >= 0, LGCN_E_INVALID, "graph_sizes_query: bad argument");
    LGCN_REQUIRE(N < (int64_t)INT32_MAX - 1 && E < (int64_t)INT32_MAX - 64, LGCN_E_RANGE,
                 "graph: N=%lld / E=%lld exceed the int32 internal range (shard the edge list)",
                 (long long)N, (long long)E);
    const size_t e1 = (size_t)(E > 0 ? E : 1);
    const size_t max_slots = 2 * e1 / (E < LGCN_SMALL_GRAPH ? LGCN_ROW_SPLIT_SMALL : LGCN_ROW_SPLIT) + 2;
    out->ptr_bytes = sizeof(int32_t) * (size_t)(N + 1);
    out->nbr_bytes = sizeof(int32_t) * e1;
    out->dis_bytes = sizeof(float) * (size_t)(N > 0 ? N : 1);
    out->active_bytes = (size_t)(N > 0 ? N : 1);
    out->task_bytes = sizeof(lgcn_task) * ((size_t)N + max_slots);
    out->partial_bytes = sizeof(float) * PARTIAL_STRIDE * max_slots;
    out->counter_bytes = sizeof(int32_t) * max_slots;
    out->workspace_bytes = carve(nullptr, N, E).total;
    out->active_list_bytes = sizeof(int32_t) * (size_t)(N > 0 ? N : 1);
    return LGCN_OK;
}

extern "C" int lgcn_graph_build(const int64_t *edge_index, int64_t E, int64_t N, int64_t U, lgcn_graph *g,
                                void *workspace, size_t workspace_bytes, void *stream) {
    using namespace lgcn;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(g && workspace && (E == 0 || edge_index), LGCN_E_INVALID, "graph_build: null argument");
    LGCN_REQUIRE(N > 0 && U >= 0 && U <= N && E >= 0, LGCN_E_INVALID, "graph_build: bad sizes N=%lld U=%lld E=%lld",
                 (long long)N, (long long)U, (long long)E);
    lgcn_graph_sizes sz;
    int rc = lgcn_graph_sizes_query(N, E, &sz);
    if (rc) return rc;
    LGCN_REQUIRE(workspace_bytes >= sz.workspace_bytes, LGCN_E_WORKSPACE, "graph_build: workspace %zu < %zu",
                 workspace_bytes, sz.workspace_bytes);
    LGCN_REQUIRE(g->in_ptr && g->in_nbr && g->in_trip && g->out_ptr && g->out_nbr && g->out_trip && g->dis &&
                 g->active && g->in_tasks && g->out_tasks && g->partials && g->slot_counters && g->active_list,
                 LGCN_E_INVALID, "graph_build: graph arrays not allocated");
    BuildWs w = carve(workspace, N, E);
    const int T = 256;
    const int gE = cdiv(E > 0 ? E : 1, T), gN = cdiv(N + 1, T);
    const int bits = bits_for(N);
    const int split = E < LGCN_SMALL_GRAPH ? LGCN_ROW_SPLIT_SMALL : LGCN_ROW_SPLIT;
    int *in_ptr = (int *)g->in_ptr, *out_ptr = (int *)g->out_ptr;

    LGCN_CUDA(cudaMemsetAsync(w.meta, 0, sizeof(long long) * META_COUNT, st));
    LGCN_CUDA(cudaMemsetAsync(g->slot_counters, 0, sz.counter_bytes, st));
    if (E > 0) {
        convert_kernel<<<gE, T, 0, st>>>(edge_index, E, N, U, w.row32, w.col32, w.eid, w.meta);
        LGCN_LAUNCH_CHECK();
        cub::TransformInputIterator<int, IsUser, const int *> flags(w.row32, IsUser{(int)U});
        size_t tb = w.cub_bytes;
        LGCN_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, flags, w.trip_id, (int)E, st));
        count_triplets_kernel<<<1, 1, 0, st>>>(w.row32, w.trip_id, E, (int)U, w.meta);
        LGCN_LAUNCH_CHECK();
        // CSR by target
        tb = w.cub_bytes;
        LGCN_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.col32, w.keys_sorted, w.eid, w.eid_sorted,
                                                  (int)E, 0, bits, st));
        fill_csr_kernel<<<gE, T, 0, st>>>(w.eid_sorted, w.row32, w.row32, w.trip_id, E, (int)U,
                                         (int *)g->in_nbr, (int *)g->in_trip);
        LGCN_LAUNCH_CHECK();
        ptr_kernel<<<gN, T, 0, st>>>(w.keys_sorted, E, N, in_ptr);
        LGCN_LAUNCH_CHECK();
        unsorted_kernel<<<gE, T, 0, st>>>(w.keys_sorted, (const int *)g->in_nbr, E, w.meta);
        LGCN_LAUNCH_CHECK();
        // CSR by source
        tb = w.cub_bytes;
        LGCN_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.row32, w.keys_sorted, w.eid, w.eid_sorted,
                                                  (int)E, 0, bits, st));
        fill_csr_kernel<<<gE, T, 0, st>>>(w.eid_sorted, w.col32, w.row32, w.trip_id, E, (int)U,
                                         (int *)g->out_nbr, (int *)g->out_trip);
        LGCN_LAUNCH_CHECK();
        ptr_kernel<<<gN, T, 0, st>>>(w.keys_sorted, E, N, out_ptr);
        LGCN_LAUNCH_CHECK();
    } else {
        LGCN_CUDA(cudaMemsetAsync(in_ptr, 0, sz.ptr_bytes, st));
        LGCN_CUDA(cudaMemsetAsync(out_ptr, 0, sz.ptr_bytes, st));
    }
    node_kernel<<<gN, T, 0, st>>>(in_ptr, out_ptr, N, split, (float *)g->dis, (uint8_t *)g->active, w.cnt_in,
                                  w.slot_in, w.cnt_out, w.slot_out, w.act32);
    LGCN_LAUNCH_CHECK();
    int *scans[5] = {w.cnt_in, w.slot_in, w.cnt_out, w.slot_out, w.act32};
    for (int i = 0; i < 5; ++i) {
        size_t tb = w.cub_bytes;
        LGCN_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, scans[i], scans[i], (int)(N + 1), st));
    }
    task_kernel<<<gN, T, 0, st>>>(in_ptr, in_ptr, out_ptr, w.cnt_in, w.slot_in, N, split, (lgcn_task *)g->in_tasks);
    LGCN_LAUNCH_CHECK();
    task_kernel<<<gN, T, 0, st>>>(out_ptr, in_ptr, out_ptr, w.cnt_out, w.slot_out, N, split, (lgcn_task *)g->out_tasks);
    LGCN_LAUNCH_CHECK();
    active_list_kernel<<<gN, T, 0, st>>>(g->active, w.act32, N, (int *)g->active_list);
    LGCN_LAUNCH_CHECK();
    meta_kernel<<<1, 1, 0, st>>>(w.cnt_in, w.cnt_out, w.slot_in, w.slot_out, w.act32, N, (int)U, w.meta);
    LGCN_LAUNCH_CHECK();
    long long meta[META_COUNT];
    LGCN_CUDA(cudaMemcpyAsync(meta, w.meta, sizeof(meta), cudaMemcpyDeviceToHost, st));
    LGCN_CUDA(cudaStreamSynchronize(st));
    LGCN_REQUIRE(meta[META_BAD] == 0, LGCN_E_INVALID,
                 "graph_build: %lld edges have an id outside [0,%lld) or do not join a user (<%lld) and a movie",
                 meta[META_BAD], (long long)N, (long long)U);
    g->num_nodes = (int32_t)N;
    g->num_users = (int32_t)U;
    g->num_edges = E;
    g->num_triplets = meta[META_P];
    g->n_in_tasks = (int32_t)meta[META_IN_TASKS];
    g->n_out_tasks = (int32_t)meta[META_OUT_TASKS];
    g->n_in_user_tasks = (int32_t)meta[META_IN_USER_TASKS];
    g->n_out_user_tasks = (int32_t)meta[META_OUT_USER_TASKS];
    g->n_in_slots = (int32_t)meta[META_IN_SLOTS];
    g->n_out_slots = (int32_t)meta[META_OUT_SLOTS];
    g->num_active = (int32_t)meta[META_ACTIVE];
    g->row_split = split;
    g->in_src_sorted = meta[META_UNSORTED] == 0 ? 1 : 0;
    g->reserved0 = 0;
    return LGCN_OK;
}
