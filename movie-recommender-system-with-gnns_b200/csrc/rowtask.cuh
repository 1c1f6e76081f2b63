// Warp-per-row-task driver shared by the SpMM (forward / backward) and BPR kernels.
//
// One warp owns one lgcn_task = a run of at most LGCN_ROW_SPLIT edges of one CSR row.  The warp
// is split in two half-warps; each half-warp gathers one neighbour row at a time with 16
// 128-bit loads (16 lanes x float4 = the 256 B embedding row), UNROLL rows in flight per
// half-warp, so a warp keeps 2*UNROLL independent 256 B gathers outstanding.  Lane l (mod 16)
// accumulates columns 4l..4l+3; the two halves are combined with one xor-16 shuffle.
//
// Rows longer than LGCN_ROW_SPLIT are cut into several tasks.  Each writes its partial sum to a
// slot; the task that arrives last (atomic counter per row) re-reads all partials IN SLOT ORDER
// and runs the row epilogue, so the result is deterministic and no float atomics are used.
//
// An Op provides:
//   accumulate(row, begin, end, lane, float4& acc, float& sc, float& ex0, float& ex1)
//                                                               edge loop (acc: this lane's 4
//                                                               columns; sc: per-lane scalar)
//   epilogue(row, lane, acc, sc, float& ex0, float& ex1)        acc/sc are full-row sums,
//                                                               identical in both half-warps
//   static constexpr bool kExtras; double *extra0, *extra1      optional per-CTA reduced sums
#pragma once
#include "common.cuh"

namespace lgcn {

constexpr int UNROLL = LGCN_UNROLL;   // neighbour rows in flight per half-warp (SpMM)

// Iterate the edges [begin,end) of a CSR row.  `fetch(lane_edge_index)` is evaluated once per
// edge by the lane that owns it within a 32-edge chunk and returns a small POD that is then
// broadcast to the half-warp that processes the edge; `body(u, item)` loads, `apply(u, item)`
// consumes.  Kept as a macro-free template so every user gets the same load batching.
template <class Item, int kUnroll = UNROLL, class Fetch, class Load, class Apply>
__device__ __forceinline__ void for_each_edge(int begin, int end, int lane, Fetch fetch, Load load,
                                              Apply apply) {
    const int half = lane >> 4;
    // software pipeline: the (coalesced) index/metadata load of chunk c+1 is issued before the row
    // gathers of chunk c, so its latency hides behind them instead of heading the next iteration
    Item mine = fetch(begin + lane < end ? begin + lane : -1);
    for (int base = begin; base < end; base += 32) {
        const int n = min(32, end - base);
        const int nb = base + 32 + lane;
        Item next = fetch(nb < end ? nb : -1);
#pragma unroll
        for (int j = 0; j < 32; j += 2 * kUnroll) {
            if (j >= n) break;                       // warp-uniform
            Item it[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                it[u] = mine.shfl(j + 2 * u + half);
                load(u, it[u]);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) apply(u, it[u]);
        }
        mine = next;
    }
}

// Persistent, dynamically scheduled: a launch has (at most) as many CTAs as fit on the chip; every
// warp pulls task indices from a device counter (the fetch of the NEXT index is issued before the
// current task is processed, so its latency is hidden).  Removes both the tail of a CTA waiting for its
// longest task and the per-task CTA launch cost (measured: 1-warp CTAs were 22 % faster than 8-warp
// CTAs with static assignment).  sched[0] = next task, sched[1] = finished CTAs; the last CTA to
// finish resets both, so the pair is zero again when the next launch starts.
// Minimum resident CTAs per SM the kernel is compiled for (caps its registers); an Op may specialise it.
template <class Op>
struct MinBlocks { static constexpr int value = 0; };   // 0 = no hint (ptxas picks the register count itself)

template <class Op>
__global__ void __launch_bounds__(CTA_THREADS, MinBlocks<Op>::value)
rowtask_kernel(Op op, const lgcn_task *__restrict__ tasks, int task_begin, int task_end,
               float *__restrict__ partials, int *__restrict__ counters, int *__restrict__ sched) {
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int total = task_end - task_begin;
    float ex0 = 0.f, ex1 = 0.f;
    int t = 0;
    if (lane == 0) t = atomicAdd(sched, 1);
    t = __shfl_sync(FULL, t, 0);
    while (t < total) {
        int nxt = 0;
        if (lane == 0) nxt = atomicAdd(sched, 1);
        const int tix = task_begin + t;
        const int4 ta = __ldg(reinterpret_cast<const int4 *>(tasks + tix));
        const int4 tb = __ldg(reinterpret_cast<const int4 *>(tasks + tix) + 1);
        const int row = ta.x, begin = ta.y, end = ta.z, slot = ta.w, part = tb.x, nparts = tb.y;
        float4 acc = f4zero();
        float sc = 0.f;
        op.accumulate(row, begin, end, lane, acc, sc, ex0, ex1);
        f4add(acc, f4shfl_xor16(acc));
        sc = warp_sum(sc);
        bool run_epilogue = slot < 0;
        if (slot >= 0) {
            float *p = partials + (size_t)slot * PARTIAL_STRIDE;
            if (lane < 16) reinterpret_cast<float4 *>(p)[lane] = acc;
            if (lane == 16) p[D] = sc;
            __threadfence();
            const int first = slot - part;
            int old = 0;
            if (lane == 0) old = atomicAdd(counters + first, 1);
            old = __shfl_sync(FULL, old, 0);
            if (old == nparts - 1) {                 // last arriver reduces in slot order
                __threadfence();
                // in double: a hub row has hundreds of partials (a 250 k-edge item: 500), and their sequential fp32 sum
                // was the largest error of the whole step (1e-5 of the row's magnitude at the 10x graph)
                double ax = 0.0, ay = 0.0, az = 0.0, aw = 0.0, as = 0.0;
                for (int i = 0; i < nparts; ++i) {
                    const float *q = partials + (size_t)(first + i) * PARTIAL_STRIDE;
                    const float4 v = __ldcg(reinterpret_cast<const float4 *>(q) + (lane & 15));
                    ax += v.x; ay += v.y; az += v.z; aw += v.w;
                    as += __ldcg(q + D);
                }
                acc = make_float4((float)ax, (float)ay, (float)az, (float)aw);
                sc = (float)as;
                if (lane == 0) counters[first] = 0;  // ready for the next launch
                run_epilogue = true;
            }
        }
        if (run_epilogue) op.epilogue(row, lane, acc, sc, ex0, ex1);
        t = __shfl_sync(FULL, nxt, 0);
    }
    __shared__ float s_ex[WARPS_PER_CTA][2];
    if constexpr (Op::kExtras) {
        if (lane == 0) { s_ex[wid][0] = ex0; s_ex[wid][1] = ex1; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if constexpr (Op::kExtras) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int w = 0; w < WARPS_PER_CTA; ++w) { a += s_ex[w][0]; b += s_ex[w][1]; }
            if (op.extra0 && a != 0.0) atomicAdd(op.extra0, a);
            if (op.extra1 && b != 0.0) atomicAdd(op.extra1, b);
        }
        if (atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }
    }
}

template <class Op>
static inline int resident_ctas() {
    static int cached = 0;                      // per kernel instantiation
    if (cached == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rowtask_kernel<Op>, CTA_THREADS, 0);
        cached = (sms > 0 ? sms : 148) * (per_sm > 0 ? per_sm : 1);
    }
    return cached;
}

template <class Op>
static inline cudaError_t launch_rowtasks(const Op &op, const lgcn_task *tasks, int task_begin,
                                          int task_end, float *partials, int *counters, int *sched,
                                          cudaStream_t stream) {
    const int n = task_end - task_begin;
    if (n <= 0) return cudaSuccess;
    if (!sched || !tasks) return cudaErrorInvalidValue;
    const int want = cdiv(n, WARPS_PER_CTA), cap = resident_ctas<Op>();
    rowtask_kernel<Op><<<want < cap ? want : cap, CTA_THREADS, 0, stream>>>(op, tasks, task_begin, task_end,
                                                                            partials, counters, sched);
    return cudaGetLastError();
}

}  // namespace lgcn
