// K0b: MANY edge lists -> many lgcn_graphs in ONE pass (a fixed number of launches, one sync).
//
// Why: a Cluster-GCN epoch hands the hot path ~100 small edge lists (data/dataset_handler.py:277-285,
// utils/train_test.py:86-88).  Built one by one (graph_build.cu) each costs ~25 launches, a stream
// sync and O(N) scans for a few thousand edges, so an epoch that uploads its batches from the host
// spent > 80 % of its time there.  Here all B lists are sorted together on a composite key
// (batch, node), every per-node array is produced for the B x N grid by one kernel, and the per-batch
// sizes come back in a single copy.
//
// Each graph's arrays are slices of one arena and are BIT-IDENTICAL to what lgcn_graph_build gives for
// that edge list alone (tests/test_gpu_batched_build.py): original edge order inside a CSR row
// (stable radix sort), the same row-split rule per batch, the same task order.  The graphs of one
// build SHARE partials / slot_counters / sched, so they must be used one after another on one stream
// (training does exactly that).
#include "common.cuh"
#include <cub/cub.cuh>
#include <limits.h>
#include <stdlib.h>

namespace lgcn {
namespace gbb {

enum { M_P = 0, M_IN_TASKS, M_OUT_TASKS, M_IN_USER_TASKS, M_OUT_USER_TASKS, M_IN_SLOTS, M_OUT_SLOTS, M_ACTIVE,
       M_IN_TASK_BASE, M_OUT_TASK_BASE, M_ACTIVE_BASE, M_COUNT = 16 };

typedef unsigned long long key_t;

struct IsUserFlag {
    int num_users;
    __host__ __device__ int operator()(int r) const { return r < num_users ? 1 : 0; }
};

__device__ __forceinline__ int batch_of(const long long *__restrict__ eoff, int B, long long e) {
    int lo = 0, hi = B;                                  // last b with eoff[b] <= e
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(eoff + mid) <= e) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int split_of(long long edges) {
    return edges < LGCN_SMALL_GRAPH ? LGCN_ROW_SPLIT_SMALL : LGCN_ROW_SPLIT;
}

// batch b's block of `edges` is its contiguous [2,E_b] tensor: E_b sources, then E_b targets
__global__ void convert_kernel(const int64_t *__restrict__ edges, const long long *__restrict__ eoff, int B,
                               long long E, long long N, long long U, int *__restrict__ row32, int *__restrict__ col32,
                               int *__restrict__ eid, int *__restrict__ bid, unsigned long long *bad) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e == E) row32[E] = INT_MAX;                      // sentinel: the triplet scan runs over E + 1 flags
    if (e >= E) return;
    const int b = batch_of(eoff, B, e);
    const long long e0 = __ldg(eoff + b), eb = __ldg(eoff + b + 1) - e0;
    const long long r = edges[2 * e0 + (e - e0)], c = edges[2 * e0 + eb + (e - e0)];
    const bool wrong = r < 0 || r >= N || c < 0 || c >= N || ((r < U) == (c < U));
    if (wrong) atomicAdd(bad, 1ull);
    row32[e] = wrong ? 0 : (int)r;
    col32[e] = wrong ? 0 : (int)c;
    eid[e] = (int)e;
    bid[e] = b;
}

__global__ void keys_kernel(const int *__restrict__ node32, const int *__restrict__ bid, long long E, int nbits,
                            key_t *__restrict__ keys) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) keys[e] = ((key_t)bid[e] << nbits) | (key_t)(unsigned)node32[e];
}

// sorted position s (global = batch offset + position inside the batch's CSR)
__global__ void fill_csr_kernel(const int *__restrict__ eid_sorted, const int *__restrict__ other32,
                                const int *__restrict__ row32, const int *__restrict__ trip_scan,
                                const int *__restrict__ bid, const long long *__restrict__ eoff, long long E, int U,
                                int *__restrict__ nbr, int *__restrict__ trip) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= E) return;
    const int e = eid_sorted[s];
    nbr[s] = other32[e];
    trip[s] = row32[e] < U ? trip_scan[e] - trip_scan[__ldg(eoff + bid[e])] : -1;
}

// ptr[b][n] = first position (relative to the batch) whose sorted key is >= (b, n), for all B x (N+1) cells: degree
// histogram (one atomic per edge) + ONE exclusive scan over the whole grid.  The spare cell [b][N] of every row holds
// -E_b, so the running sum is back to zero at the start of the next row and the scan yields batch-relative positions
// directly.  (A bisection per cell took 0.3 ms per direction at 100 x 221,589 cells: 13 dependent loads each.)
__global__ void degree_kernel(const key_t *__restrict__ keys_sorted, const long long *__restrict__ eoff, long long E,
                              long long N, int B, int nbits, int *__restrict__ ptr) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < B) ptr[e * (N + 1) + N] = -(int)(__ldg(eoff + e + 1) - __ldg(eoff + e));
    if (e >= E) return;
    const key_t k = keys_sorted[e];
    const long long b = (long long)(k >> nbits), n = (long long)(k & (((key_t)1 << nbits) - 1));
    atomicAdd(ptr + b * (N + 1) + n, 1);
}

__global__ void node_kernel(const int *__restrict__ in_ptr, const int *__restrict__ out_ptr,
                            const long long *__restrict__ eoff, long long N, int B, float *__restrict__ dis,
                            uint8_t *__restrict__ active, int *__restrict__ cnt_in, int *__restrict__ slot_in,
                            int *__restrict__ cnt_out, int *__restrict__ slot_out, int *__restrict__ act32) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n > N) return;
    const int b = blockIdx.y;
    const long long idx = (long long)b * (N + 1) + n;
    if (n == N) {
        cnt_in[idx] = slot_in[idx] = cnt_out[idx] = slot_out[idx] = act32[idx] = 0;
        if (b == B - 1) cnt_in[idx + 1] = slot_in[idx + 1] = cnt_out[idx + 1] = slot_out[idx + 1] = act32[idx + 1] = 0;
        return;
    }
    const int split = split_of(__ldg(eoff + b + 1) - __ldg(eoff + b));
    const int din = in_ptr[idx + 1] - in_ptr[idx], dout = out_ptr[idx + 1] - out_ptr[idx];
    dis[(long long)b * N + n] = din > 0 ? 1.0f / sqrtf((float)din) : 0.f;      // deg^-1/2, inf -> 0 (gcn_norm)
    const int a = (din > 0 || dout > 0) ? 1 : 0;
    active[(long long)b * N + n] = (uint8_t)a;
    act32[idx] = a;
    const int pin = (din + split - 1) / split, pout = (dout + split - 1) / split;
    cnt_in[idx] = a ? max(1, pin) : 0;
    cnt_out[idx] = a ? max(1, pout) : 0;
    slot_in[idx] = pin > 1 ? pin : 0;
    slot_out[idx] = pout > 1 ? pout : 0;
}

__global__ void task_kernel(const int *__restrict__ ptr, const int *__restrict__ in_ptr,
                            const int *__restrict__ out_ptr, const int *__restrict__ task_scan,
                            const int *__restrict__ slot_scan, const long long *__restrict__ eoff, long long N,
                            lgcn_task *__restrict__ tasks) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int b = blockIdx.y;
    const long long idx = (long long)b * (N + 1) + n;
    const int off = task_scan[idx], cnt = task_scan[idx + 1] - off;
    if (cnt == 0) return;
    const int split = split_of(__ldg(eoff + b + 1) - __ldg(eoff + b));
    const int pb = ptr[idx], pe = ptr[idx + 1];
    const int din = in_ptr[idx + 1] - in_ptr[idx], dout = out_ptr[idx + 1] - out_ptr[idx];
    lgcn_task *t = tasks + off;
    if (cnt == 1) {
        t[0] = lgcn_task{(int)n, pb, pe, -1, 0, 1, din, dout};
        return;
    }
    const int s0 = slot_scan[idx] - slot_scan[(long long)b * (N + 1)];
    for (int i = 0; i < cnt; ++i) {
        const int tb = pb + i * split;
        t[i] = lgcn_task{(int)n, tb, min(pe, tb + split), s0 + i, i, cnt, din, dout};
    }
}

__global__ void active_list_kernel(const uint8_t *__restrict__ active, const int *__restrict__ act_scan, long long N,
                                   int *__restrict__ list) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int b = blockIdx.y;
    if (active[(long long)b * N + n]) list[act_scan[(long long)b * (N + 1) + n]] = (int)n;
}

__global__ void meta_kernel(const int *__restrict__ trip_scan, const long long *__restrict__ eoff,
                            const int *__restrict__ cin, const int *__restrict__ cout, const int *__restrict__ sin,
                            const int *__restrict__ sout, const int *__restrict__ act, long long N, int U, int B,
                            long long *__restrict__ meta) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long lo = (long long)b * (N + 1), hi = lo + N + 1;
    long long *m = meta + (long long)b * M_COUNT;
    m[M_P] = trip_scan[eoff[b + 1]] - trip_scan[eoff[b]];
    m[M_IN_TASKS] = cin[hi] - cin[lo];
    m[M_OUT_TASKS] = cout[hi] - cout[lo];
    m[M_IN_USER_TASKS] = cin[lo + U] - cin[lo];
    m[M_OUT_USER_TASKS] = cout[lo + U] - cout[lo];
    m[M_IN_SLOTS] = sin[hi] - sin[lo];
    m[M_OUT_SLOTS] = sout[hi] - sout[lo];
    m[M_ACTIVE] = act[hi] - act[lo];
    m[M_IN_TASK_BASE] = cin[lo];
    m[M_OUT_TASK_BASE] = cout[lo];
    m[M_ACTIVE_BASE] = act[lo];
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int bits_for(long long n) {
    int b = 1;
    while (b < 62 && ((long long)1 << b) < n) ++b;
    return b;
}

struct Caps {                 // exact upper bounds from the per-batch edge counts (host)
    size_t tasks, active, slots;
    long long E;
};

static Caps caps_of(const int64_t *eoff, int64_t B, int64_t N) {
    Caps c{0, 0, 2, eoff[B] - eoff[0]};
    for (int64_t b = 0; b < B; ++b) {
        const long long eb = eoff[b + 1] - eoff[b];
        const long long act = 2 * eb < N ? 2 * eb : N;              // an edge touches two nodes
        const long long split = eb < LGCN_SMALL_GRAPH ? LGCN_ROW_SPLIT_SMALL : LGCN_ROW_SPLIT;
        c.active += (size_t)act;
        c.tasks += (size_t)(act + eb / split + 1);                  // sum over active rows of max(1, ceil(d/split))
        const size_t slots = (size_t)(2 * eb / split + 2);          // rows with > split edges: ceil(d/split) <= 2d/split
        if (slots > c.slots) c.slots = slots;
    }
    return c;
}

struct Arena {
    int *in_ptr, *out_ptr, *in_nbr, *in_trip, *out_nbr, *out_trip, *active_list, *slot_counters, *sched;
    float *dis, *partials;
    uint8_t *active;
    lgcn_task *in_tasks, *out_tasks;
    size_t total;
};

static Arena carve_arena(void *base, int64_t N, int64_t B, const Caps &c) {
    Arena a{};
    char *p = (char *)base;
    auto take = [&](size_t bytes) { char *q = p; p += align256(bytes ? bytes : 1); return q; };
    const size_t eb = sizeof(int) * (size_t)(c.E > 0 ? c.E : 1), pb = sizeof(int) * (size_t)B * (size_t)(N + 1);
    a.in_ptr = (int *)take(pb); a.out_ptr = (int *)take(pb);
    a.in_nbr = (int *)take(eb); a.in_trip = (int *)take(eb); a.out_nbr = (int *)take(eb); a.out_trip = (int *)take(eb);
    a.dis = (float *)take(sizeof(float) * (size_t)B * (size_t)N);
    a.active = (uint8_t *)take((size_t)B * (size_t)N);
    a.in_tasks = (lgcn_task *)take(sizeof(lgcn_task) * c.tasks);
    a.out_tasks = (lgcn_task *)take(sizeof(lgcn_task) * c.tasks);
    a.active_list = (int *)take(sizeof(int) * c.active);
    a.partials = (float *)take(sizeof(float) * PARTIAL_STRIDE * c.slots);
    a.slot_counters = (int *)take(sizeof(int) * c.slots);
    a.sched = (int *)take(sizeof(int) * 2);
    a.total = (size_t)(p - (char *)base);
    return a;
}

struct Work {
    int *row32, *col32, *eid, *bid, *trip_scan, *eid_sorted;
    key_t *keys, *keys_sorted;
    int *scan[5];             // cnt_in, slot_in, cnt_out, slot_out, act32: [B*(N+1)+1], scanned in place
    long long *eoff, *meta;
    unsigned long long *bad;
    void *cub_temp;
    size_t cub_bytes, total;
};

static size_t cub_temp_bytes(long long cells, long long E, int end_bit) {
    size_t a = 0, b = 0, c = 0;
    key_t *k = nullptr;
    int *v = nullptr;
    cub::DeviceRadixSort::SortPairs(nullptr, a, k, k, v, v, (int)E, 0, end_bit);
    cub::DeviceScan::ExclusiveSum(nullptr, b, v, v, (int)cells);
    cub::TransformInputIterator<int, IsUserFlag, const int *> it(v, IsUserFlag{0});
    cub::DeviceScan::ExclusiveSum(nullptr, c, it, v, (int)(E + 1));
    return align256(a > b ? (a > c ? a : c) : (b > c ? b : c)) + 256;
}

static Work carve_work(void *base, int64_t N, int64_t B, long long E, int end_bit) {
    Work w{};
    char *p = (char *)base;
    auto take = [&](size_t bytes) { char *q = p; p += align256(bytes ? bytes : 1); return q; };
    const size_t e1 = (size_t)(E + 1), cells = (size_t)B * (size_t)(N + 1) + 1;
    w.row32 = (int *)take(4 * e1); w.col32 = (int *)take(4 * e1); w.eid = (int *)take(4 * e1);
    w.bid = (int *)take(4 * e1); w.trip_scan = (int *)take(4 * e1); w.eid_sorted = (int *)take(4 * e1);
    w.keys = (key_t *)take(8 * e1); w.keys_sorted = (key_t *)take(8 * e1);
    for (int i = 0; i < 5; ++i) w.scan[i] = (int *)take(4 * cells);
    w.eoff = (long long *)take(8 * (size_t)(B + 1));
    w.meta = (long long *)take(8 * (size_t)B * M_COUNT);
    w.bad = (unsigned long long *)take(8);
    w.cub_bytes = cub_temp_bytes((long long)cells, E, end_bit);
    w.cub_temp = take(w.cub_bytes);
    w.total = (size_t)(p - (char *)base);
    return w;
}

static int check_shape(const int64_t *eoff, int64_t B, int64_t N) {
    LGCN_REQUIRE(eoff && B >= 1 && B <= 65535 && N > 0, LGCN_E_INVALID, "graph_build_batched: bad B=%lld / N=%lld",
                 (long long)B, (long long)N);
    for (int64_t b = 0; b < B; ++b)
        LGCN_REQUIRE(eoff[b + 1] >= eoff[b], LGCN_E_INVALID, "graph_build_batched: edge_off not ascending at %lld", (long long)b);
    LGCN_REQUIRE(eoff[0] == 0, LGCN_E_INVALID, "graph_build_batched: edge_off[0] must be 0");
    LGCN_REQUIRE(eoff[B] < (int64_t)INT32_MAX - 64 && (long long)B * (N + 1) + 1 < (long long)INT32_MAX - 64 &&
                 N < (int64_t)INT32_MAX - 1, LGCN_E_RANGE,
                 "graph_build_batched: %lld edges / %lld x %lld cells exceed the int32 internal range (build in chunks)",
                 (long long)eoff[B], (long long)B, (long long)N);
    return LGCN_OK;
}

}  // namespace gbb
}  // namespace lgcn

extern "C" int lgcn_graph_batched_sizes(int64_t N, int64_t B, const int64_t *edge_off, lgcn_batched_sizes *out) {
    using namespace lgcn;
    using namespace lgcn::gbb;
    LGCN_REQUIRE(out, LGCN_E_INVALID, "graph_batched_sizes: null argument");
    int rc = check_shape(edge_off, B, N);
    if (rc) return rc;
    const Caps c = caps_of(edge_off, B, N);
    out->arena_bytes = carve_arena(nullptr, N, B, c).total;
    out->workspace_bytes = carve_work(nullptr, N, B, c.E, bits_for(N) + bits_for(B)).total;
    return LGCN_OK;
}

extern "C" int lgcn_graph_build_batched(const int64_t *edges, const int64_t *edge_off, int64_t B, int64_t N, int64_t U,
                                        lgcn_graph *graphs, void *arena, size_t arena_bytes, void *workspace,
                                        size_t workspace_bytes, void *stream) {
    using namespace lgcn;
    using namespace lgcn::gbb;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(graphs && arena && workspace, LGCN_E_INVALID, "graph_build_batched: null argument");
    LGCN_REQUIRE(U >= 0 && U <= N, LGCN_E_INVALID, "graph_build_batched: bad num_users %lld", (long long)U);
    int rc = check_shape(edge_off, B, N);
    if (rc) return rc;
    const Caps caps = caps_of(edge_off, B, N);
    const long long E = caps.E;
    LGCN_REQUIRE(E == 0 || edges, LGCN_E_INVALID, "graph_build_batched: null edge buffer");
    const int nbits = bits_for(N), end_bit = nbits + bits_for(B);
    const Arena a = carve_arena(arena, N, B, caps);
    const Work w = carve_work(workspace, N, B, E, end_bit);
    LGCN_REQUIRE(arena_bytes >= a.total, LGCN_E_WORKSPACE, "graph_build_batched: arena %zu < %zu", arena_bytes, a.total);
    LGCN_REQUIRE(workspace_bytes >= w.total, LGCN_E_WORKSPACE, "graph_build_batched: workspace %zu < %zu", workspace_bytes,
                 w.total);
    const int T = 256;
    const int gE = cdiv(E + 1, T);
    const dim3 gN(cdiv(N + 1, T), (unsigned)B);
    const long long cells = (long long)B * (N + 1) + 1;

    LGCN_CUDA(cudaMemcpyAsync(w.eoff, edge_off, 8 * (size_t)(B + 1), cudaMemcpyHostToDevice, st));
    LGCN_CUDA(cudaMemsetAsync(w.bad, 0, 8, st));
    LGCN_CUDA(cudaMemsetAsync(a.slot_counters, 0, sizeof(int) * caps.slots, st));
    LGCN_CUDA(cudaMemsetAsync(a.sched, 0, sizeof(int) * 2, st));
    convert_kernel<<<gE, T, 0, st>>>(edges, w.eoff, (int)B, E, N, U, w.row32, w.col32, w.eid, w.bid, w.bad);
    LGCN_LAUNCH_CHECK();
    {
        cub::TransformInputIterator<int, IsUserFlag, const int *> flags(w.row32, IsUserFlag{(int)U});
        size_t tb = w.cub_bytes;
        LGCN_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, flags, w.trip_scan, (int)(E + 1), st));
    }
    if (E > 0) {
        for (int dir = 0; dir < 2; ++dir) {              // 0: CSR by target, 1: CSR by source
            const int *key_node = dir == 0 ? w.col32 : w.row32, *other = dir == 0 ? w.row32 : w.col32;
            keys_kernel<<<gE, T, 0, st>>>(key_node, w.bid, E, nbits, w.keys);
            LGCN_LAUNCH_CHECK();
            size_t tb = w.cub_bytes;
            LGCN_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.keys, w.keys_sorted, w.eid, w.eid_sorted, (int)E,
                                                      0, end_bit, st));
            fill_csr_kernel<<<gE, T, 0, st>>>(w.eid_sorted, other, w.row32, w.trip_scan, w.bid, w.eoff, E, (int)U,
                                             dir == 0 ? a.in_nbr : a.out_nbr, dir == 0 ? a.in_trip : a.out_trip);
            LGCN_LAUNCH_CHECK();
            int *ptr = dir == 0 ? a.in_ptr : a.out_ptr;
            const long long pcells = (long long)B * (N + 1);
            LGCN_CUDA(cudaMemsetAsync(ptr, 0, sizeof(int) * (size_t)pcells, st));
            degree_kernel<<<cdiv((E > B ? E : B), T), T, 0, st>>>(w.keys_sorted, w.eoff, E, N, (int)B, nbits, ptr);
            LGCN_LAUNCH_CHECK();
            tb = w.cub_bytes;
            LGCN_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, ptr, ptr, (int)pcells, st));
        }
    } else {
        LGCN_CUDA(cudaMemsetAsync(a.in_ptr, 0, sizeof(int) * (size_t)B * (size_t)(N + 1), st));
        LGCN_CUDA(cudaMemsetAsync(a.out_ptr, 0, sizeof(int) * (size_t)B * (size_t)(N + 1), st));
    }
    node_kernel<<<gN, T, 0, st>>>(a.in_ptr, a.out_ptr, w.eoff, N, (int)B, a.dis, a.active, w.scan[0], w.scan[1],
                                  w.scan[2], w.scan[3], w.scan[4]);
    LGCN_LAUNCH_CHECK();
    for (int i = 0; i < 5; ++i) {
        size_t tb = w.cub_bytes;
        LGCN_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, w.scan[i], w.scan[i], (int)cells, st));
    }
    task_kernel<<<gN, T, 0, st>>>(a.in_ptr, a.in_ptr, a.out_ptr, w.scan[0], w.scan[1], w.eoff, N, a.in_tasks);
    LGCN_LAUNCH_CHECK();
    task_kernel<<<gN, T, 0, st>>>(a.out_ptr, a.in_ptr, a.out_ptr, w.scan[2], w.scan[3], w.eoff, N, a.out_tasks);
    LGCN_LAUNCH_CHECK();
    active_list_kernel<<<gN, T, 0, st>>>(a.active, w.scan[4], N, a.active_list);
    LGCN_LAUNCH_CHECK();
    meta_kernel<<<cdiv(B, 128), 128, 0, st>>>(w.trip_scan, w.eoff, w.scan[0], w.scan[2], w.scan[1], w.scan[3], w.scan[4],
                                             N, (int)U, (int)B, w.meta);
    LGCN_LAUNCH_CHECK();

    // one copy + one sync for all B graphs
    static thread_local long long *host_meta = nullptr;
    static thread_local size_t host_meta_cap = 0;
    const size_t need = (size_t)B * M_COUNT + 1;
    if (host_meta_cap < need) {
        free(host_meta);
        host_meta = (long long *)malloc(8 * need);
        host_meta_cap = host_meta ? need : 0;
        LGCN_REQUIRE(host_meta, LGCN_E_INVALID, "graph_build_batched: out of host memory");
    }
    LGCN_CUDA(cudaMemcpyAsync(host_meta, w.meta, 8 * (size_t)B * M_COUNT, cudaMemcpyDeviceToHost, st));
    LGCN_CUDA(cudaMemcpyAsync(host_meta + (size_t)B * M_COUNT, w.bad, 8, cudaMemcpyDeviceToHost, st));
    LGCN_CUDA(cudaStreamSynchronize(st));
    const long long nbad = host_meta[(size_t)B * M_COUNT];
    LGCN_REQUIRE(nbad == 0, LGCN_E_INVALID,
                 "graph_build_batched: %lld edges have an id outside [0,%lld) or do not join a user (<%lld) and a movie",
                 nbad, (long long)N, (long long)U);
    for (int64_t b = 0; b < B; ++b) {
        const long long *m = host_meta + (size_t)b * M_COUNT;
        const long long e0 = edge_off[b], eb = edge_off[b + 1] - e0;
        lgcn_graph *g = graphs + b;
        g->num_nodes = (int32_t)N;
        g->num_users = (int32_t)U;
        g->num_edges = eb;
        g->num_triplets = m[M_P];
        g->in_ptr = a.in_ptr + (size_t)b * (size_t)(N + 1);
        g->out_ptr = a.out_ptr + (size_t)b * (size_t)(N + 1);
        g->in_nbr = a.in_nbr + e0; g->in_trip = a.in_trip + e0;
        g->out_nbr = a.out_nbr + e0; g->out_trip = a.out_trip + e0;
        g->dis = a.dis + (size_t)b * (size_t)N;
        g->active = a.active + (size_t)b * (size_t)N;
        g->in_tasks = a.in_tasks + m[M_IN_TASK_BASE];
        g->out_tasks = a.out_tasks + m[M_OUT_TASK_BASE];
        g->n_in_tasks = (int32_t)m[M_IN_TASKS];
        g->n_out_tasks = (int32_t)m[M_OUT_TASKS];
        g->n_in_user_tasks = (int32_t)m[M_IN_USER_TASKS];
        g->n_out_user_tasks = (int32_t)m[M_OUT_USER_TASKS];
        g->n_in_slots = (int32_t)m[M_IN_SLOTS];
        g->n_out_slots = (int32_t)m[M_OUT_SLOTS];
        g->partials = a.partials;
        g->slot_counters = a.slot_counters;
        g->num_active = (int32_t)m[M_ACTIVE];
        g->row_split = (int32_t)(eb < LGCN_SMALL_GRAPH ? LGCN_ROW_SPLIT_SMALL : LGCN_ROW_SPLIT);
        g->active_list = a.active_list + m[M_ACTIVE_BASE];
        g->sched = a.sched;
        g->in_src_sorted = 0;            // not needed by the single-GPU steps these graphs serve
        g->reserved0 = 0;
    }
    return LGCN_OK;
}

// `batch.to(device)` for all lists of an epoch in one call: list b (its contiguous [2,E_b] int64 tensor in HOST
// memory, pinned or pageable) goes to dst + 2 * edge_off[b] -- the layout lgcn_graph_build_batched reads.  One
// cudaMemcpyAsync per list (asynchronous for pinned sources, staged by the driver for pageable ones).
extern "C" int lgcn_upload_lists(const int64_t *const *host_lists, const int64_t *edge_off, int64_t B, int64_t *dst,
                                 void *stream) {
    using namespace lgcn;
    LGCN_REQUIRE(host_lists && edge_off && dst && B >= 1, LGCN_E_INVALID, "upload_lists: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t b = 0; b < B; ++b) {
        const int64_t e = edge_off[b + 1] - edge_off[b];
        LGCN_REQUIRE(e >= 0 && (e == 0 || host_lists[b]), LGCN_E_INVALID, "upload_lists: bad list %lld", (long long)b);
        if (e == 0) continue;
        LGCN_CUDA(cudaMemcpyAsync(dst + 2 * edge_off[b], host_lists[b], sizeof(int64_t) * 2 * (size_t)e, cudaMemcpyHostToDevice, st));
    }
    return LGCN_OK;
}
