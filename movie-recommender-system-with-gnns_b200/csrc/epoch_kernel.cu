// A whole RUN of sparse (touched-rows) training steps in ONE persistent cooperative launch.
//
// Replaces, for a sequence of Cluster-GCN batches, the loop of /root/reference/utils/train_test.py:
// 86-101 (`for batch in train_loader:` zero_grad, forward, bpr_loss, backward, clip_grad_norm_,
// Adam.step).  Arithmetic per step is that of lgcn_train_step_sparse (sparse_step.cu): only the rows
// a batch touches are visited, the zero-gradient Adam updates of the others are replayed exactly when
// they are next touched.
//
// Why one kernel: a median ML-25M cluster batch has 2.1 k active rows, 6.8 k edges and 3.4 k triplets; as 16
// launches per step the epoch is bound by launch/drain latency, ~185 us per step measured.  Here three CTAs per SM
// stay resident for the whole run and phases are separated by device-wide barriers.  What a phase costs is not
// memory (one L2 round trip of a 256-byte row is 310 cycles) but the chain of dependent instructions every warp
// of an SM runs at the same moment, so the design keeps that chain short:
//   * per-node metadata rides in the task descriptors (deg_in / deg_out): no O(N) array is touched;
//   * each warp keeps a shared-memory image of its tasks (WarpCache: descriptors, neighbour ids, per-edge
//     weights dis[nbr], triplet / negative ids), fetched once per step while the previous step's last barrier
//     is waited for; phases read indices by shared-memory broadcast (no shuffle, no descriptor -> index ->
//     row chain of dependent loads);
//   * layer 1 reads the weight tables with the per-edge weight (no pre-scaled table, no pre-scale phase).
//
// Two CTA roles, taken per SM (so the two instruction streams do not share schedulers or instruction cache).
// The replay of pending zero-gradient Adam steps is a long SEQUENTIAL chain per row (a user row waits ~50
// steps, each a sqrt and two divisions per element) -- the same arithmetic dense Adam performs, but on the
// critical path if done when the row is needed.  So:
//   HELPER CTAs work one step AHEAD: during step b they prepare step b+1 -- stamp its active rows, collect
//     its distinct inactive negatives (the run's negatives are sampled up front), bring all those rows up to
//     date (rows step b touches itself are skipped: phase J leaves them up to date), and prefetch step b+2's
//     arrays into L2.
//   MAIN CTAs run the phases of step b, separated by barriers among themselves:
//     B  forward layers 1..K (barrier after each; the last forms the layer mean and 1/||.||)
//     E  BPR: the CTA's half-warps share the triplets of its resident user tasks (user-row sums in shared
//        memory, gradient of BOTH items by vector atomics), the owner runs the row epilogue
//     G  backward layers 1..K (barrier after each); layer 1 gathers dL/dfinal with the weight dis[target]
//        and also forms the gradient rows of the inactive negatives
//     J  clip + Adam step on the touched rows, restore the all-zero invariants, loss; then fill the
//        WarpCache for step b+1 between the arrive and the wait of the end-of-step barrier
//   Everybody meets at the end of the step.  Step 0 is prepared by all CTAs before the loop.
// Stamp / list arrays are double-buffered by step parity so that preparing b+1 never disturbs b.
//
// Memory rules inside the kernel: everything another SM may have written earlier in the launch is read
// with ld.global.cg (L2, the coherence point); __ldg only for data that is immutable for the whole
// launch (task lists, index arrays, negatives, dis, the bias-correction table).
//
// Diagnostics: -DEP_TRACE builds the per-CTA / per-warp barrier trace (tools/epoch_trace.py); LGCN_EPOCH_PROF=1
// makes the shipped kernel write one globaltimer stamp per phase (tools/epoch_breakdown.py).
#include "adam.cuh"
#include "rowtask.cuh"
#include <stdlib.h>
#include <vector>

namespace lgcn {
namespace ep {

#ifndef EP_GATHER_UNROLL
#define EP_GATHER_UNROLL 8
#endif
#ifndef EP_BPR_A_UNROLL
#define EP_BPR_A_UNROLL 2
#endif
#ifndef EP_CTAS_PER_SM
#define EP_CTAS_PER_SM 3
#endif
constexpr int EP_WARPS = 8;        // EP_CTAS_PER_SM CTAs per SM: one of them a HELPER, the others MAIN
constexpr int EP_THREADS = EP_WARPS * 32;

struct StepDesc {
    const lgcn_task *in_tasks, *out_tasks;
    const int32_t *in_nbr, *in_trip, *out_nbr, *out_trip;
    const float *dis;            // [N] in-degree^-1/2 of THIS batch graph (weight of a layer-1 in-edge)
    float *partials;
    int32_t *slot_counters;
    const int64_t *neg;
    float *loss_out;
    long long P;
    int n_in_tasks, n_out_tasks, n_in_user_tasks, n_out_user_tasks;
    long long num_edges;
};

struct EpochArgs {
    const StepDesc *steps;
    int num_steps, K;
    float *user_w, *item_w;
    int num_users, num_items;
    float4 *m, *v;
    int32_t *row_step;
    int64_t *step;
    AdamHyper h;
    float *final_emb, *rnorm, *G, *grad, *work;
    int32_t *neg_count;
    int32_t *act_stamp;  // [2][N]  step in which the node is active, by step parity
    int32_t *neg_flag;   // [2][I]  step in which the item is an inactive negative
    int32_t *neg_list;   // [2][I]  the distinct inactive negatives
    float *scratch;
    double *accum;       // [2][4], by step parity
    int32_t *counts;     // [2] length of neg_list, by step parity
    unsigned *bar;       // monotonic arrival counters: [0] all CTAs, [32] main CTAs, [64] helper CTAs
    int num_helpers;     // how many CTAs prepare the next step (0: single-step run, everybody is MAIN)
    int main_sms;        // SMs [0, main_sms) host the MAIN CTAs, the others the helpers
    float bpr_coeff;
    long long *prof;     // optional [num_steps][16] globaltimer stamps at the phase boundaries (diagnostics)
};

__device__ __forceinline__ float4 ldcg4(const float4 *p) { return __ldcg(p); }

__device__ __forceinline__ void stamp(long long *prof, int b, int &slot, int gtid) {
    if (prof && gtid == 0) {
        long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (slot < 16) prof[(size_t)b * 16 + slot] = now;
    }
    ++slot;
}

#ifdef EP_TRACE
// Diagnostic build (tools/epoch_trace.py): every CTA records when it arrives at and when it leaves each barrier.
__device__ long long *g_trace;            // [num_steps][16][gridDim.x][2]
__device__ long long *g_trace_warp;       // [num_steps][16][gridDim.x][8]: when each WARP reached the barrier
__device__ int *g_trace_smid;             // [2][gridDim.x]: SM, role << 16 | index within the role
__device__ long long *g_trace_fine;       // [num_steps][16][main warps (<= 4096)][8]: clock64 at points inside a phase
__device__ __forceinline__ long long gtimer() {
    long long now;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
    return now;
}
#define EP_TRACE_AT(b, slot, which)                                                                     \
    do {                                                                                                \
        if (g_trace && threadIdx.x == 0 && (slot) < 16)                                                 \
            g_trace[(((size_t)(b) * 16 + (slot)) * gridDim.x + blockIdx.x) * 2 + (which)] = gtimer();   \
    } while (0)
#define EP_TRACE_WARP(b, slot)                                                                           \
    do {                                                                                                \
        if (g_trace_warp && (threadIdx.x & 31) == 0 && (slot) < 16)                                     \
            g_trace_warp[(((size_t)(b) * 16 + (slot)) * gridDim.x + blockIdx.x) * 8 + (threadIdx.x >> 5)] = gtimer(); \
    } while (0)
#define EP_FINE(b, slot, gw, point)                                                                     \
    do {                                                                                                \
        if (g_trace_fine && (threadIdx.x & 31) == 0 && (slot) < 16 && (gw) < 4096)                      \
            g_trace_fine[(((size_t)(b) * 16 + (slot)) * 4096 + (gw)) * 8 + (point)] = clock64();         \
    } while (0)
#else
#define EP_TRACE_AT(b, slot, which) do { } while (0)
#define EP_TRACE_WARP(b, slot) do { } while (0)
#define EP_FINE(b, slot, gw, point) do { } while (0)
#endif

// Arrive (release) on a monotonic counter, wait (acquire) until `count` more CTAs have arrived than at
// the previous barrier on it.  bar.sync before/after extends the ordering to the whole CTA.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned &target, unsigned count, int tb = 0, int tslot = 16) {
    EP_TRACE_WARP(tb, tslot);
    __syncthreads();
    target += count;
    EP_TRACE_AT(tb, tslot, 0);
    if (threadIdx.x == 0) {
        unsigned cur;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(bar), "r"(1u) : "memory");
        do {          // (relaxed polls + one acquire fence were measured: 0.25-0.5 us SLOWER per barrier)
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(bar) : "memory");
        } while ((int)(cur - target) < 0);
    }
    EP_TRACE_AT(tb, tslot, 1);
    __syncthreads();
}

// The same barrier in two halves: work placed between arrive and wait overlaps the wait.  Only work that
// reads launch-immutable data and writes CTA-private state may go there.
__device__ __forceinline__ void barrier_arrive(unsigned *bar, int tb = 0, int tslot = 16) {
    EP_TRACE_WARP(tb, tslot);
    __syncthreads();
    EP_TRACE_AT(tb, tslot, 0);
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(bar), "r"(1u) : "memory");
}
__device__ __forceinline__ void barrier_wait(unsigned *bar, unsigned &target, unsigned count, int tb = 0, int tslot = 16) {
    target += count;
    if (threadIdx.x == 0) {
        unsigned cur;
        do {          // (relaxed polls + one acquire fence were measured: 0.25-0.5 us SLOWER per barrier)
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(bar) : "memory");
        } while ((int)(cur - target) < 0);
    }
    EP_TRACE_AT(tb, tslot, 1);
    __syncthreads();
}

// Per-CTA sum of two per-warp scalars, then ONE double atomic per CTA and target: with one atomic per warp the
// ~4 k same-address atomics of a phase serialised at the L2 and delayed every CTA's arrival at the barrier.
// Contains a __syncthreads(); the caller's grid barrier orders the reuse of `red`.
__device__ __forceinline__ void cta_add2(float (*red)[2], int lane, float v0, float v1, double *d0, double *d1) {
    if (lane == 0) { red[threadIdx.x >> 5][0] = v0; red[threadIdx.x >> 5][1] = v1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < EP_THREADS / 32; ++i) { s0 += (double)red[i][0]; s1 += (double)red[i][1]; }
        if (d0 && s0 != 0.0) atomicAdd(d0, s0);
        if (d1 && s1 != 0.0) atomicAdd(d1, s1);
    }
}

// Per-warp shared-memory image of the step's work.  Main warp gw owns tasks gw, gw + nw, ... of both task
// lists in EVERY phase of a step, so the descriptors and index runs of its first EP_CT tasks per list are
// fetched once per step (during the wait of the previous step's last barrier) instead of heading every
// phase with a descriptor -> index -> row chain of dependent L2 round trips.  Everything cached is
// immutable for the launch.  Tasks beyond EP_CT per warp, or longer than EP_CE edges, use the global path.
#ifndef EP_CACHE_TASKS
#define EP_CACHE_TASKS 2
#endif
#ifndef EP_CACHE_EDGES
#define EP_CACHE_EDGES 64
#endif
constexpr int EP_CT = EP_CACHE_TASKS;       // resident tasks per warp and direction
constexpr int EP_CE = EP_CACHE_EDGES;       // edges per cache slot (a multiple of 32)
constexpr int EP_SCRATCH = EP_CT;           // slot index of the staging area for everything that is not resident
constexpr int DIR_IN = 0, DIR_OUT = 1;
struct __align__(16) WarpCache {
    int4 ta[2][EP_CT], tc[2][EP_CT];          // [direction][slot]: the lgcn_task
    int32_t nbr[2][EP_CT + 1][EP_CE];         // in: source ids; out: target ids
    int32_t trip[2][EP_CT + 1][EP_CE];        // in (item rows) / out (user rows): triplet number of the edge
    int32_t ng[EP_CT + 1][EP_CE];             // out, user rows: num_users + negative item of the triplet
    float wgt[2][EP_CT + 1][EP_CE];           // dis[neighbour]: forward layer 1 reads e0 itself, backward layer 1 dL/dfinal
};

// What a phase needs staged besides the neighbour ids when a chunk is not resident.
enum { ST_NONE = 0, ST_WGT = 1, ST_TRIP = 2, ST_NG = 4 };

struct Tab {                        // mutable e0 = (user_w, item_w)
    float *user, *item;
    int num_users;
    __device__ __forceinline__ float4 *row4(int r) const {
        float *p = r < num_users ? user + (size_t)r * D : item + (size_t)(r - num_users) * D;
        return reinterpret_cast<float4 *>(p);
    }
};

__device__ __forceinline__ float dis_of(int deg_in) { return deg_in > 0 ? 1.0f / sqrtf((float)deg_in) : 0.f; }

// Edges [0,n) of a chunk whose index runs sit in shared memory.  Half-warp h takes edges h, h+2, ...; kUnroll
// edges per half-warp are in flight.  `fetch(e)` (e = -1 past the end) is evaluated by all 16 lanes of the half
// (shared-memory broadcasts and same-address global loads), so no shuffle is needed to hand an edge to its lanes --
// the shuffle-broadcast traversal of rowtask.cuh cost ~1400 cycles for a four-edge row here, against 310 for the
// one L2 round trip in it.  The loop is rolled: the kernel holds every phase's code at once.
template <class Item, int kUnroll, class Fetch, class Load, class Apply>
__device__ __forceinline__ void for_each_edge_smem(int n, int half, Fetch fetch, Load load, Apply apply) {
#pragma unroll 1
    for (int o = 0; o < n; o += 2 * kUnroll) {
        Item it[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int e = o + 2 * u + half;
            it[u] = fetch(e < n ? e : -1);
            load(u, it[u]);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) apply(u, it[u]);
    }
}

struct Nbr {
    int nbr;
    float w;
    float4 v;
};

// Stage edges [e0, e0+n) (n <= EP_CE) of a task into the warp's scratch slot.
struct StepDesc;
template <int kNeed>
__device__ __forceinline__ void stage_chunk(WarpCache &wc, int dir, const int32_t *__restrict__ nbr,
                                            const int32_t *__restrict__ trp, const float *__restrict__ dis,
                                            const int64_t *__restrict__ neg, int e0, int n, int lane, int U) {
    int nb[EP_CE / 32], tr[EP_CE / 32];
#pragma unroll
    for (int h = 0; h < EP_CE / 32; ++h) {
        const int e = lane + 32 * h;
        nb[h] = e < n ? __ldg(nbr + e0 + e) : -1;
        tr[h] = (kNeed & (ST_TRIP | ST_NG)) && e < n ? __ldg(trp + e0 + e) : -1;
    }
    float wg[EP_CE / 32];
    int ngv[EP_CE / 32];
#pragma unroll
    for (int h = 0; h < EP_CE / 32; ++h) {
        wg[h] = (kNeed & ST_WGT) && nb[h] >= 0 ? __ldg(dis + nb[h]) : 0.f;
        ngv[h] = (kNeed & ST_NG) && tr[h] >= 0 ? U + (int)__ldg(neg + tr[h]) : 0;
    }
    __syncwarp();                                     // the slot may still be read by lanes of the previous chunk
#pragma unroll
    for (int h = 0; h < EP_CE / 32; ++h) {
        const int e = lane + 32 * h;
        wc.nbr[dir][EP_SCRATCH][e] = nb[h];
        if (kNeed & (ST_TRIP | ST_NG)) wc.trip[dir][EP_SCRATCH][e] = tr[h];
        if (kNeed & ST_WGT) wc.wgt[dir][EP_SCRATCH][e] = wg[h];
        if (kNeed & ST_NG) wc.ng[EP_SCRATCH][e] = ngv[h];
    }
    __syncwarp();
}

// Walk the tasks [tb,te) of one list that belong to this warp (t = gw, gw + nw, ...: the same ownership in
// every phase, which is what makes the WarpCache valid); same split-row protocol as rowtask_kernel.
// accumulate(row, n, s, acc, sc): edges [0,n) of cache slot s.  A task that is not resident (beyond EP_CT per
// warp, or longer than EP_CE edges) goes through the scratch slot in chunks of EP_CE edges, staged with what
// the phase needs (kNeed); the traversal itself is the same.  kSc: the phase uses the per-row scalar.
template <int kNeed, bool kSc, class Acc, class Epi>
__device__ __forceinline__ void run_tasks(WarpCache &wc, int dir, const StepDesc &d, int tb, int te, int gw, int nw, int lane,
                                          int U, Acc accumulate, Epi epilogue, int fb = 0, int fslot = 16) {
    const lgcn_task *__restrict__ tasks = dir == DIR_IN ? d.in_tasks : d.out_tasks;
    const int32_t *__restrict__ nbr = dir == DIR_IN ? d.in_nbr : d.out_nbr;
    const int32_t *__restrict__ trp = dir == DIR_IN ? d.in_trip : d.out_trip;
    float *const partials = d.partials;
    int *const counters = d.slot_counters;
    int ci = 0;
    EP_FINE(fb, fslot, gw, 0);
    for (int t = gw; t < te; t += nw, ++ci) {
        if (t < tb) continue;
        int4 ta, tc;
        if (ci < EP_CT) {
            ta = wc.ta[dir][ci];
            tc = wc.tc[dir][ci];
        } else {
            ta = __ldg(reinterpret_cast<const int4 *>(tasks + t));
            tc = __ldg(reinterpret_cast<const int4 *>(tasks + t) + 1);
        }
        const int row = ta.x, begin = ta.y, end = ta.z, slot = ta.w, part = tc.x, nparts = tc.y;
        const bool resident = ci < EP_CT && end - begin <= EP_CE;
        float4 acc = f4zero();
        float sc = 0.f;
        if (ci == 0) EP_FINE(fb, fslot, gw, 1);
        int e0 = begin;
        do {
            const int n = resident ? end - begin : min(EP_CE, end - e0);
            if (!resident) stage_chunk<kNeed>(wc, dir, nbr, trp, d.dis, d.neg, e0, n, lane, U);
            accumulate(row, n, resident ? ci : EP_SCRATCH, acc, sc);
            e0 += EP_CE;
        } while (!resident && e0 < end);
        if (ci == 0) EP_FINE(fb, fslot, gw, 2);
        f4add(acc, f4shfl_xor16(acc));
        if (kSc) sc = warp_sum(sc);
        if (ci == 0) EP_FINE(fb, fslot, gw, 3);
        bool run = slot < 0;
        if (slot >= 0) {
            float *p = partials + (size_t)slot * PARTIAL_STRIDE;
            if (lane < 16) reinterpret_cast<float4 *>(p)[lane] = acc;
            if (kSc && lane == 16) p[D] = sc;
            __threadfence();
            const int first = slot - part;
            int old = 0;
            if (lane == 0) old = atomicAdd(counters + first, 1);
            old = __shfl_sync(FULL, old, 0);
            if (old == nparts - 1) {
                __threadfence();
                acc = f4zero();
                sc = 0.f;
                // four slots per round trip, summed in slot order (deterministic)
                for (int i = 0; i < nparts; i += 4) {
                    float4 q4[4];
                    float qs[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float *q = partials + (size_t)(first + min(i + u, nparts - 1)) * PARTIAL_STRIDE;
                        q4[u] = __ldcg(reinterpret_cast<const float4 *>(q) + (lane & 15));
                        qs[u] = kSc ? __ldcg(q + D) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i + u < nparts) { f4add(acc, q4[u]); sc += qs[u]; }
                }
                if (lane == 0) counters[first] = 0;
                run = true;
            }
        }
        if (run) epilogue(row, tc.z, tc.w, acc, sc);
        if (ci == 0) EP_FINE(fb, fslot, gw, 4);
    }
    EP_FINE(fb, fslot, gw, 5);
}

// Fill this warp's cache for the step described by d (launch-immutable data only; called between the arrive
// and the wait of a barrier).  Loads are issued stage by stage for all slots so the dependent round trips
// (descriptor -> indices -> dis / negative) overlap across slots.
__device__ __forceinline__ void fill_cache(WarpCache &wc, const StepDesc &d, int gw, int nw, int lane, int U) {
    int4 ta[2][EP_CT], tc[2][EP_CT];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
        const lgcn_task *tasks = dir == DIR_IN ? d.in_tasks : d.out_tasks;
        const int nt = dir == DIR_IN ? d.n_in_tasks : d.n_out_tasks;
#pragma unroll
        for (int ci = 0; ci < EP_CT; ++ci) {
            const int t = gw + ci * nw;
            ta[dir][ci] = make_int4(0, 0, 0, -1);
            tc[dir][ci] = make_int4(0, 1, 0, 0);
            if (t < nt) {
                ta[dir][ci] = __ldg(reinterpret_cast<const int4 *>(tasks + t));
                tc[dir][ci] = __ldg(reinterpret_cast<const int4 *>(tasks + t) + 1);
            }
        }
    }
    int nb[2][EP_CT][EP_CE / 32], tr[2][EP_CT][EP_CE / 32];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
        const int32_t *nbr = dir == DIR_IN ? d.in_nbr : d.out_nbr;
        const int32_t *trp = dir == DIR_IN ? d.in_trip : d.out_trip;
#pragma unroll
        for (int ci = 0; ci < EP_CT; ++ci) {
            const int begin = ta[dir][ci].y, len = ta[dir][ci].z - begin;
            // triplet numbers: BPR walks the user rows of the by-source list
            const bool want_trip = dir == DIR_OUT && ta[dir][ci].x < U;
#pragma unroll
            for (int h = 0; h < EP_CE / 32; ++h) {
                const int e = lane + 32 * h;
                const bool ok = len <= EP_CE && e < len;
                nb[dir][ci][h] = ok ? __ldg(nbr + begin + e) : -1;
                tr[dir][ci][h] = ok && want_trip ? __ldg(trp + begin + e) : -1;
            }
        }
    }
#pragma unroll
    for (int ci = 0; ci < EP_CT; ++ci) {
#pragma unroll
        for (int h = 0; h < EP_CE / 32; ++h) {
            const int e = lane + 32 * h;
            const int src = nb[DIR_IN][ci][h], dst = nb[DIR_OUT][ci][h], t_out = tr[DIR_OUT][ci][h];
            wc.wgt[DIR_IN][ci][e] = src >= 0 ? __ldg(d.dis + src) : 0.f;
            wc.wgt[DIR_OUT][ci][e] = dst >= 0 ? __ldg(d.dis + dst) : 0.f;
            wc.ng[ci][e] = t_out >= 0 ? U + (int)__ldg(d.neg + t_out) : 0;
        }
    }
#pragma unroll
    for (int dir = 0; dir < 2; ++dir)
#pragma unroll
        for (int ci = 0; ci < EP_CT; ++ci) {
            if (lane == 0) { wc.ta[dir][ci] = ta[dir][ci]; wc.tc[dir][ci] = tc[dir][ci]; }
#pragma unroll
            for (int h = 0; h < EP_CE / 32; ++h) {
                wc.nbr[dir][ci][lane + 32 * h] = nb[dir][ci][h];
                if (dir == DIR_OUT) wc.trip[dir][ci][lane + 32 * h] = tr[dir][ci][h];
            }
        }
    __syncwarp();
}

// Bring `row` up to optimiser step `target` by replaying zero-gradient Adam steps (adam_replay_kernel's
// arithmetic, element by element).  ONE WARP per row, two elements per lane: the chain of a row is bound by the
// instructions one warp can issue (an IEEE sqrt and two divisions per element and step), so a row spread over
// 32 lanes finishes in half the time of the half-warp / float4 layout.
__device__ __forceinline__ void replay_row(int row, int lane, int target, const Tab &w, float4 *m, float4 *v,
                                           int32_t *row_step, const AdamHyper &h) {
    const int from = __ldcg(row_step + row);
    if (from >= target) return;                                       // warp-uniform
    float2 *pp = reinterpret_cast<float2 *>(w.row4(row)) + lane;
    float2 *mp = reinterpret_cast<float2 *>(m + (size_t)row * D4) + lane;
    float2 *vp = reinterpret_cast<float2 *>(v + (size_t)row * D4) + lane;
    float2 p2 = __ldcg(pp), m2 = __ldcg(mp), v2 = __ldcg(vp);
    const bool live = m2.x != 0.f || m2.y != 0.f || v2.x != 0.f || v2.y != 0.f;
    if (__any_sync(FULL, live)) {                                     // all-zero moments: the row does not move
        // iterations only chain through one fma each for p, m, v: unrolling lets the sqrt / division
        // sequences of neighbouring steps overlap
#pragma unroll 4
        for (int t = from + 1; t <= target; ++t) {
            const AdamScalars a = adam_scalars_tab(h, t);
            adam_elem(p2.x, m2.x, v2.x, 0.f, a);
            adam_elem(p2.y, m2.y, v2.y, 0.f, a);
        }
        *pp = p2; *mp = m2; *vp = v2;
    }
    if (lane == 0) row_step[row] = target;
}

// Touched by the step with number `tp` (its stamps live in act_p / flag_p)?  Such a row is left alone by
// whoever prepares the following step: phase J of step tp brings it to step tp itself.
__device__ __forceinline__ bool touched_by(int row, int num_users, int tp, const int32_t *act_p, const int32_t *flag_p) {
    if (!act_p) return false;
    if (__ldcg(act_p + row) == tp) return true;
    return row >= num_users && __ldcg(flag_p + row - num_users) == tp;
}

// Replay the pending zero-gradient steps of a row set up to step `target`, behind ONE copy of the
// (unrolled) Adam loop.  RP_ACTIVE: the first-part in-tasks' rows, which are also stamped as active in
// step target+1;  RP_LIST: rows num_users + list[i].  Rows touched by step `target` itself are skipped.
// Rows differ widely in pending steps (1 .. steps per epoch), so after its first (static) item a warp pulls
// further items from a device queue (`queue`, zero on entry; the pull is issued before the current row is
// processed so its latency hides behind the replay).
enum { RP_ACTIVE = 0, RP_LIST = 1 };

__device__ __noinline__ void replay_rows(int mode, const lgcn_task *__restrict__ tasks, const int32_t *list, int count,
                                         int target, Tab w, float4 *m, float4 *v, int32_t *row_step, int32_t *act_next,
                                         const int32_t *act_p, const int32_t *flag_p, AdamHyper h, int gw, int nw, int lane,
                                         int32_t *queue) {
    int i = gw;
    while (i < count) {                                               // warp-uniform
        int nxt = 0;
        if (lane == 0) nxt = nw + atomicAdd(queue, 1);
        int row;
        bool valid = true;
        if (mode == RP_LIST) {
            row = w.num_users + __ldcg(list + i);
        } else {
            row = __ldg(&tasks[i].row);
            valid = __ldg(&tasks[i].part) == 0;                      // first part of a split row speaks for the row
            if (valid && lane == 0) act_next[row] = target + 1;
        }
        if (valid) valid = !touched_by(row, w.num_users, target, act_p, flag_p);
        if (valid) replay_row(row, lane, target, w, m, v, row_step, h);
        i = __shfl_sync(FULL, nxt, 0);
    }
}

// The list rows (inactive negatives) have short chains (an item is sampled every ~17 steps) and cold state, so a
// row costs mostly its dependent loads: two rows per warp (half-warp / float4 layout, adam_replay_kernel's) halve
// the number of rounds.  Same queue protocol as replay_rows, in pairs.
__device__ __noinline__ void replay_list(const int32_t *list, int count, int target, Tab w, float4 *m, float4 *v,
                                         int32_t *row_step, const int32_t *act_p, const int32_t *flag_p, AdamHyper h, int gw,
                                         int nw, int lane, int32_t *queue) {
    const int half = lane >> 4, l16 = lane & 15;
    const unsigned half_mask = 0xffffu << (lane & 16);
    int i = 2 * gw;
    while (i < count) {                                               // warp-uniform
        int nxt = 0;
        if (lane == 0) nxt = 2 * nw + 2 * atomicAdd(queue, 1);
        const int mine = i + half;
        bool valid = mine < count;
        int row = 0;
        if (valid) {
            row = w.num_users + __ldcg(list + mine);
            valid = !touched_by(row, w.num_users, target, act_p, flag_p);
        }
        const int from = valid ? __ldcg(row_step + row) : target;
        valid = valid && from < target;
        float4 *pp = w.row4(row) + l16;
        const size_t o = (size_t)row * D4 + l16;
        float4 p4 = f4zero(), m4 = f4zero(), v4 = f4zero();
        if (valid) { p4 = ldcg4(pp); m4 = ldcg4(m + o); v4 = ldcg4(v + o); }
        const bool live = m4.x != 0.f || m4.y != 0.f || m4.z != 0.f || m4.w != 0.f ||
                          v4.x != 0.f || v4.y != 0.f || v4.z != 0.f || v4.w != 0.f;
        const bool any_live = (__ballot_sync(FULL, live) & half_mask) != 0u;
        if (valid && any_live) {
            const float4 zero = f4zero();
#pragma unroll 2
            for (int t = from + 1; t <= target; ++t) {
                const AdamScalars a = adam_scalars_tab(h, t);
                adam_vec(p4, m4, v4, zero, 1.0f, a);
            }
            *pp = p4; m[o] = m4; v[o] = v4;
        }
        if (valid && l16 == 0) row_step[row] = target;
        i = __shfl_sync(FULL, nxt, 0);
    }
}

__device__ __noinline__ void adam_row(int row, int lane, int t, const Tab &w, float4 *m, float4 *v, const float4 *grad,
                                         float4 *G, int32_t *row_step, int32_t *neg_count, float clip,
                                         const AdamScalars &a) {
    const int l16 = lane & 15;
    float4 *pp = w.row4(row) + l16;
    const size_t o = (size_t)row * D4 + l16;
    float4 p4 = ldcg4(pp), m4 = ldcg4(m + o), v4 = ldcg4(v + o);
    adam_vec(p4, m4, v4, ldcg4(grad + o), clip, a);
    *pp = p4; m[o] = m4; v[o] = v4;
    G[o] = f4zero();
    if (l16 == 0) {
        row_step[row] = t;
        if (row >= w.num_users) neg_count[row - w.num_users] = 0;
    }
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

__device__ __forceinline__ void prefetch_range(const void *base, size_t bytes, int gtid, int nthreads) {
    const char *p = reinterpret_cast<const char *>(base);
    for (size_t o = (size_t)gtid * 128; o < bytes; o += (size_t)nthreads * 128) prefetch_l2(p + o);
}

struct TripA {
    int dst, t, ng;
    float rp, rn;
    float4 vp, vn;
    __device__ __forceinline__ TripA shfl(int src) const {
        TripA r;
        r.dst = __shfl_sync(FULL, dst, src);
        r.t = __shfl_sync(FULL, t, src);
        r.ng = __shfl_sync(FULL, ng, src);
        r.rp = __shfl_sync(FULL, rp, src);
        r.rn = __shfl_sync(FULL, rn, src);
        r.vp = f4zero();
        r.vn = f4zero();
        return r;
    }
};

// BPR arithmetic of one triplet, evaluated by a half-warp that holds the three rows (utils/train_test.py:18-64 for one
// (u, i+, i-)): fu = normalised user row, vp / vn = final rows of the items, rp = 1/||vp||, rn = 1/||vn|| or < 0 when
// the negative is inactive in this batch (vn is then c0 * e0[neg] and its norm is formed here).
struct BprOut {
    float s, cp, cn, rn, sp;
};
__device__ __forceinline__ BprOut bpr_math(const float4 &fu, const float4 &vp, const float4 &vn, float rp, float rn) {
    BprOut o;
    const float n2 = half_sum(f4dot(vn, vn));
    o.rn = rn < 0.f ? 1.0f / sqrtf(n2) : rn;
    o.cp = half_sum(f4dot(fu, vp)) * rp;
    o.cn = half_sum(f4dot(fu, vn)) * o.rn;
    const float x = 10.f * (o.cp - o.cn);
    o.sp = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
    o.s = -1.f / (1.f + expf(-x));
    return o;
}
// dL/dfinal of both items of the triplet, added where they live: a cluster batch has at most a few dozen triplets per
// item, so the vector atomics do not pile up (the full-graph kernels walk the item rows instead, bpr.cu)
__device__ __forceinline__ void bpr_item_grads(float4 *G4, int dst, int ng, const float4 &fu, const float4 &vp,
                                               const float4 &vn, float rp, const BprOut &o, float invP, int l16) {
    const float kp = o.s * rp * invP;
    float4 gp = fu;
    f4fma(gp, -o.cp * rp, vp);
    gp = f4scale(kp, gp);
    atomicAdd(G4 + (size_t)dst * D4 + l16, gp);
    const float kk = -o.s * o.rn * invP;
    float4 g = fu;
    f4fma(g, -o.cn * o.rn, vn);
    g = f4scale(kk, g);
    atomicAdd(G4 + (size_t)ng * D4 + l16, g);
}

struct TripP {                       // a pooled triplet: user row and accumulator slot travel with it
    int row, dst, ng, j;
    float ru, rp, rn;
    float4 vu, vp, vn;
};

static_assert(EP_WARPS * EP_CT <= 16, "the pooled BPR pass scans the CTA's task slots with a half-warp");

__global__ void __launch_bounds__(EP_THREADS, EP_CTAS_PER_SM) epoch_kernel(const EpochArgs a) {
    __shared__ WarpCache s_cache[EP_WARPS];
    __shared__ float s_red[EP_WARPS][2];
    __shared__ __align__(16) float s_acc[EP_WARPS * EP_CT][D + 4];   // BPR: user-row sums of the CTA's resident tasks; [D] = the row scalar
    __shared__ int s_cnt[16];
    WarpCache &wc = s_cache[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4;
    const int gw_all = blockIdx.x * EP_WARPS + (threadIdx.x >> 5), nw_all = gridDim.x * EP_WARPS;
    const unsigned nblocks = gridDim.x;
    unsigned tgt_all = 0, tgt_main = 0, tgt_help = 0;
    unsigned *const bar_all = a.bar, *const bar_main = a.bar + 32, *const bar_help = a.bar + 64;

    // Roles by SM: the CTAs of the first `main_sms` SMs run the phases of a step (MAIN), the others prepare the next
    // step (HELPERS).  A main warp's phase is a few hundred instructions around one or two dependent memory round
    // trips; sharing a scheduler with helper warps (long ready-to-issue sqrt / division chains) stretched every one
    // of them, and both code paths fought over the instruction cache.  Every CTA takes a number within its role;
    // the counts are read after a barrier, so nothing is assumed about how the hardware numbers or fills SMs.
    __shared__ int s_role[2];
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const int r = (a.num_helpers > 0 && (int)smid >= a.main_sms) ? 1 : 0;
        s_role[0] = r;
        s_role[1] = atomicAdd(a.counts + 8 + r, 1);
    }
    grid_barrier(bar_all, tgt_all, nblocks);             // (__syncthreads inside publishes s_role)
    unsigned nmain = (unsigned)__ldcg(a.counts + 8), nhelp = (unsigned)__ldcg(a.counts + 9);
    bool helper = s_role[0] != 0;
    int cidx = s_role[1];
    if (a.num_helpers > 0 && (nmain == 0 || nhelp == 0)) {   // unexpected SM numbering: roles by block index
        nhelp = (unsigned)a.num_helpers;
        nmain = nblocks - nhelp;
        helper = blockIdx.x >= nmain;
        cidx = helper ? (int)(blockIdx.x - nmain) : (int)blockIdx.x;
    }
    // main: consecutive tasks (the parts of a long row, the long rows of the most active users) go to
    // different CTAs -- with gw = cta * warps + warp the slowest CTA set the pace of every phase
    const int gw = helper ? cidx * EP_WARPS + (int)(threadIdx.x >> 5) : (int)(threadIdx.x >> 5) * (int)nmain + cidx;
    const int nw = helper ? (int)nhelp * EP_WARPS : (int)nmain * EP_WARPS;
    const int gtid_all = blockIdx.x * EP_THREADS + threadIdx.x, nthreads_all = (int)nblocks * EP_THREADS;
    const int gtid = helper ? -1 : cidx * EP_THREADS + (int)threadIdx.x;       // 0: the main CTAs' lead thread
    const int htid = helper ? cidx * EP_THREADS + (int)threadIdx.x : -1, nhthreads = (int)nhelp * EP_THREADS;
    const int K = a.K, U = a.num_users;
    const size_t n = (size_t)a.num_users + (size_t)a.num_items;
    const Tab w{a.user_w, a.item_w, U};
    const float c0 = 1.0f / (float)((K + 1) * (K + 1));
    float *const Z0 = a.work, *const Z1 = a.work + n * D;   // backward tables (the forward's y tables are dead then)
    float4 *const G4 = reinterpret_cast<float4 *>(a.G);
    const float4 *const F4 = reinterpret_cast<const float4 *>(a.final_emb);
    const long long t0 = *a.step;                           // only rewritten after the first barrier of a step

    // Pull a step's launch-immutable arrays into L2 (two steps ahead of their use by the main CTAs' cache fill).
    auto prefetch_step = [&](int s_, int tid, int nth) {
        const StepDesc nx = a.steps[s_];
        prefetch_range(nx.in_tasks, sizeof(lgcn_task) * (size_t)nx.n_in_tasks, tid, nth);
        prefetch_range(nx.out_tasks, sizeof(lgcn_task) * (size_t)nx.n_out_tasks, tid, nth);
        prefetch_range(nx.in_nbr, 4 * (size_t)nx.num_edges, tid, nth);
        prefetch_range(nx.out_nbr, 4 * (size_t)nx.num_edges, tid, nth);
        prefetch_range(nx.in_trip, 4 * (size_t)nx.num_edges, tid, nth);
        prefetch_range(nx.out_trip, 4 * (size_t)nx.num_edges, tid, nth);
        prefetch_range(nx.neg, 8 * (size_t)nx.P, tid, nth);
        prefetch_range(nx.dis, 4 * n, tid, nth);
    };
#ifdef EP_TRACE
    if (g_trace_smid && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_trace_smid[blockIdx.x] = (int)smid;
        g_trace_smid[gridDim.x + blockIdx.x] = (helper ? 65536 : 0) + cidx;
    }
#endif
    prefetch_step(0, gtid_all, nthreads_all);
    if (a.num_steps > 1) prefetch_step(1, gtid_all, nthreads_all);

    // Prepare step s (number ts = t0+1+s): stamp its active rows, list its distinct inactive negatives, bring
    // both row sets to step ts-1.  Rows touched by the previous step are skipped (has_prev).  Executed by the
    // warps (pgw of pnw) of `pcount` CTAs that synchronise on `pbar`.
    auto prepare = [&](int s_, int pgw, int pnw, unsigned *pbar, unsigned &ptgt, unsigned pcount) {
        const StepDesc sd = a.steps[s_];
        const int ts = (int)(t0 + 1 + s_), par = s_ & 1;
        int32_t *act_s = a.act_stamp + (size_t)par * n, *flag_s = a.neg_flag + (size_t)par * a.num_items;
        int32_t *list_s = a.neg_list + (size_t)par * a.num_items, *cnt_s = a.counts + par;
        const int32_t *act_p = s_ > 0 ? a.act_stamp + (size_t)(par ^ 1) * n : nullptr;
        const int32_t *flag_p = s_ > 0 ? a.neg_flag + (size_t)(par ^ 1) * a.num_items : nullptr;
        int32_t *const q_act = a.counts + 4, *const q_list = a.counts + 5;     // work queues, zero between uses
        const int tb = s_ > 0 ? s_ - 1 : 0, ts0 = s_ > 0 ? 10 : 16;              // (trace slots of the diagnostic build)
        if (pgw == 0 && lane == 0) { *cnt_s = 0; *q_list = 0; }
        replay_rows(RP_ACTIVE, sd.in_tasks, nullptr, sd.n_in_tasks, ts - 1, w, a.m, a.v, a.row_step, act_s, act_p, flag_p, a.h,
                    pgw, pnw, lane, q_act);
        grid_barrier(pbar, ptgt, pcount, tb, ts0);
        if (pgw == 0 && lane == 0) *q_act = 0;                                    // everybody has left the queue
        for (long long base = (long long)pgw * 32; base < sd.P; base += (long long)pnw * 32) {
            const long long tt = base + lane;
            bool add = false;
            int i = 0;
            if (tt < sd.P) {
                i = (int)__ldg(sd.neg + tt);
                add = __ldcg(act_s + U + i) != ts && atomicExch(flag_s + i, ts) != ts;
            }
            const unsigned mk = __ballot_sync(FULL, add);                         // one counter update per warp
            if (mk) {
                const int leader = __ffs(mk) - 1;
                int pos = 0;
                if (lane == leader) pos = atomicAdd(cnt_s, __popc(mk));
                pos = __shfl_sync(FULL, pos, leader);
                if (add) list_s[pos + __popc(mk & ((1u << lane) - 1u))] = i;
            }
        }
        grid_barrier(pbar, ptgt, pcount, tb, ts0 + 1);
        replay_list(list_s, __ldcg(cnt_s), ts - 1, w, a.m, a.v, a.row_step, act_p, flag_p, a.h, pgw, pnw, lane, q_list);
    };
    prepare(0, gw_all, nw_all, bar_all, tgt_all, nblocks);
    barrier_arrive(bar_all);
    if (!helper) fill_cache(wc, a.steps[0], gw, nw, lane, U);
    barrier_wait(bar_all, tgt_all, nblocks);

    for (int b = 0; b < a.num_steps; ++b) {
        const StepDesc d = a.steps[b];
        const int t = (int)(t0 + 1 + b);
        double *const acc_cur = a.accum + (b & 1) * 4;
        int32_t *const cnt_cur = a.counts + (b & 1);
        const float invP = 1.0f / (float)d.P;
        const float reg_coef = 2.0f * a.bpr_coeff / (64.0f * (float)d.P);
        int ps = 0;
        stamp(a.prof, b, ps, gtid);

        const int par = b & 1;
        const int32_t *const act_cur = a.act_stamp + (size_t)par * n;
        const int32_t *const list_cur = a.neg_list + (size_t)par * a.num_items;

        if (helper) {
            int hs = 12;
            stamp(a.prof, b, hs, htid);
            if (b + 2 < a.num_steps) prefetch_step(b + 2, htid, nhthreads);
            if (b + 1 < a.num_steps) prepare(b + 1, gw, nw, bar_help, tgt_help, nhelp);
            stamp(a.prof, b, hs, htid);
            grid_barrier(bar_all, tgt_all, nblocks, b, 2 * K + 2);     // end of the step
            continue;
        }
        if (gtid == 0) {                                 // the other parity's sums: last read in step b-1, next used in b+1
            double *nx = a.accum + ((b + 1) & 1) * 4;
            nx[0] = 0.0; nx[1] = 0.0; nx[2] = 0.0; nx[3] = 0.0;
        }

        // ---- B: forward layers 1..K (layer 1 reads the weight tables of the active rows: all at step t-1) ----
        auto fwd_layer = [&](int k) {
            const float *src = k == 1 ? nullptr : a.work + (size_t)(k - 2) * n * D;
            float *dst = a.work + (size_t)(k - 1) * n * D;
            const bool last = k == K;
            run_tasks<ST_WGT, false>(wc, DIR_IN, d, 0, d.n_in_tasks, gw, nw, lane, U,
                      [&](int, int n, int sl, float4 &acc, float &) {
                          const int32_t *cn = wc.nbr[DIR_IN][sl];
                          if (k == 1) {        // raw = sum dis[r] * e0[r] straight from the weight tables (no pre-scaled copy)
                              const float *cw = wc.wgt[DIR_IN][sl];
                              for_each_edge_smem<Nbr, EP_GATHER_UNROLL>(
                                  n, half,
                                  [&](int e) { Nbr it; it.nbr = e >= 0 ? cn[e] : -1; it.w = e >= 0 ? cw[e] : 0.f; it.v = f4zero(); return it; },
                                  [&](int, Nbr &it) { if (it.nbr >= 0) it.v = ldcg4(w.row4(it.nbr) + l16); },
                                  [&](int, Nbr &it) { f4fma(acc, it.w, it.v); });
                          } else {
                              const float4 *x4 = reinterpret_cast<const float4 *>(src);
                              for_each_edge_smem<Nbr, EP_GATHER_UNROLL>(
                                  n, half,
                                  [&](int e) { Nbr it; it.nbr = e >= 0 ? cn[e] : -1; it.w = 0.f; it.v = f4zero(); return it; },
                                  [&](int, Nbr &it) { if (it.nbr >= 0) it.v = ldcg4(x4 + (size_t)it.nbr * D4 + l16); },
                                  [&](int, Nbr &it) { f4add(acc, it.v); });
                          }
                      },
                      [&](int row, int din, int, const float4 &raw, float) {
                          if (!last) {
                              const float inv = din > 0 ? 1.0f / (float)din : 0.f;
                              if (lane < 16) reinterpret_cast<float4 *>(dst)[(size_t)row * D4 + l16] = f4scale(inv, raw);
                          } else {
                              float4 s = f4zero();
                              for (int i = 0; i < K - 1; ++i)
                                  f4add(s, ldcg4(reinterpret_cast<const float4 *>(a.work + (size_t)i * n * D) + (size_t)row * D4 + l16));
                              float4 f = ldcg4(w.row4(row) + l16);
                              f4fma(f, sqrtf((float)din), s);
                              f4fma(f, dis_of(din), raw);
                              f = f4scale(c0, f);
                              if (lane < 16) reinterpret_cast<float4 *>(a.final_emb)[(size_t)row * D4 + l16] = f;
                              const float n2 = half_sum(f4dot(f, f));
                              if (lane == 0) a.rnorm[row] = 1.0f / sqrtf(n2);
                          }
                      }, b, ps);
        };
        for (int k = 1; k <= K; ++k) {
            fwd_layer(k);
            EP_FINE(b, ps, gw, 6);
            grid_barrier(bar_main, tgt_main, nmain, b, ps);
            EP_FINE(b, ps + 1, gw, 7);
            stamp(a.prof, b, ps, gtid);
        }

        // ---- E: BPR over user rows -------------------------------------------------------------
        // The triplets of the CTA's resident user tasks are POOLED: every half-warp takes triplets round-robin, whoever
        // owns the user, and adds the user-row terms into shared-memory accumulators -- a user with many triplets no longer
        // keeps one warp busy for several rounds while the rest of the CTA waits at the barrier.  The owner then runs the
        // row epilogue.  Non-resident tasks (hub cluster) are walked by their owner as before.
        float ex0 = 0.f, ex1 = 0.f;
        float ru_row = 0.f;                 // 1/||final|| and normalised final row of the user this warp works on:
        float4 fu_row = f4zero();           // loaded by accumulate, reused by the epilogue of the same row
        bool pooled = false;                // CTA-uniform: the resident user tasks went through the pooled pass
        {
            const int wid = threadIdx.x >> 5;
            if (lane < EP_CT) {
                const int4 ta = wc.ta[DIR_OUT][lane];
                const int len = ta.z - ta.y;
                s_cnt[wid * EP_CT + lane] = (gw + lane * nw < d.n_out_user_tasks && len <= EP_CE) ? len : 0;
            }
            for (int i = lane; i < EP_CT * (D + 4); i += 32) (&s_acc[wid * EP_CT][0])[i] = 0.f;
            __syncthreads();
            constexpr int NS = EP_WARPS * EP_CT;
            const int cnt_l = l16 < NS ? s_cnt[l16] : 0;       // both halves scan the slots
            int incl = cnt_l;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o, 16);
                if (l16 >= o) incl += v;
            }
            // Pooling pays when it shortens the longest chain; with hundreds of triplets per CTA (hub cluster) every warp
            // is busy anyway and the extra user-row load per triplet only costs.
            const int total_all = __shfl_sync(FULL, incl, 15, 16);
            pooled = total_all <= 16 * EP_BPR_A_UNROLL * 4;
            const int total = pooled ? total_all : 0;
            const int excl = incl - cnt_l;
            const int hh = wid * 2 + half;
            float loss = 0.f;
#pragma unroll 1
            for (int base = 0; base < total; base += 16 * EP_BPR_A_UNROLL) {       // CTA-uniform trip count
                TripP it[EP_BPR_A_UNROLL];
#pragma unroll
                for (int u = 0; u < EP_BPR_A_UNROLL; ++u) {
                    const int k = base + u * 16 + hh;
                    const bool ok = k < total;
                    // slot of pooled triplet k = number of slots whose inclusive prefix is <= k (empty slots drop out)
                    const unsigned m = __ballot_sync(FULL, ok && l16 < NS && incl <= k);
                    const int j = __popc((m >> (16 * half)) & 0xffffu);
                    const int e = k - __shfl_sync(FULL, excl, j & 15, 16);
                    TripP &q = it[u];
                    q.row = 0; q.dst = -1; q.ng = 0; q.j = j & 15; q.ru = 0.f; q.rp = 0.f; q.rn = 0.f;
                    q.vu = f4zero(); q.vp = f4zero(); q.vn = f4zero();
                    if (ok) {
                        const WarpCache &oc = s_cache[j / EP_CT];
                        const int ci = j % EP_CT;
                        q.row = oc.ta[DIR_OUT][ci].x;
                        q.dst = oc.nbr[DIR_OUT][ci][e];
                        q.ng = oc.ng[ci][e];
                        q.ru = __ldcg(a.rnorm + q.row);
                        q.rp = __ldcg(a.rnorm + q.dst);
                        const int st_ng = __ldcg(act_cur + q.ng);
                        const float rn_ng = __ldcg(a.rnorm + q.ng);
                        q.rn = st_ng != t ? -1.f : rn_ng;                                // inactive: formed in bpr_math
                        q.vu = ldcg4(F4 + (size_t)q.row * D4 + l16);
                        q.vp = ldcg4(F4 + (size_t)q.dst * D4 + l16);
                        if (q.rn < 0.f) q.vn = f4scale(c0, ldcg4(w.row4(q.ng) + l16));
                        else q.vn = ldcg4(F4 + (size_t)q.ng * D4 + l16);
                    }
                }
#pragma unroll
                for (int u = 0; u < EP_BPR_A_UNROLL; ++u) {
                    TripP &q = it[u];
                    const float4 fu = f4scale(q.ru, q.vu);
                    const BprOut o = bpr_math(fu, q.vp, q.vn, q.rp, q.rn);
                    if (q.dst >= 0) {
                        float *sa = &s_acc[q.j][0];
                        const float kp = o.s * q.rp, kn = -o.s * o.rn;
                        atomicAdd(sa + 4 * l16 + 0, fmaf(kp, q.vp.x, kn * q.vn.x));
                        atomicAdd(sa + 4 * l16 + 1, fmaf(kp, q.vp.y, kn * q.vn.y));
                        atomicAdd(sa + 4 * l16 + 2, fmaf(kp, q.vp.z, kn * q.vn.z));
                        atomicAdd(sa + 4 * l16 + 3, fmaf(kp, q.vp.w, kn * q.vn.w));
                        if (l16 == 0) {
                            loss += o.sp;
                            atomicAdd(sa + D, o.s * (o.cp - o.cn));
                            atomicAdd(a.neg_count + (q.ng - U), 1);
                        }
                        bpr_item_grads(G4, q.dst, q.ng, fu, q.vp, q.vn, q.rp, o, invP, l16);
                    }
                }
            }
            ex0 += warp_sum(loss);
            __syncthreads();
        }
        run_tasks<ST_TRIP | ST_NG, true>(wc, DIR_OUT, d, 0, d.n_out_user_tasks, gw, nw, lane, U,
                  [&](int row, int n, int sl, float4 &acc, float &sc) {
                      const float ru = __ldcg(a.rnorm + row);
                      const float4 fu = f4scale(ru, ldcg4(F4 + (size_t)row * D4 + l16));
                      ru_row = ru;
                      fu_row = fu;
                      if (pooled && sl != EP_SCRATCH) { // resident: summed by the pooled pass (half 0 / lane 0 carry the value)
                          const float *sa = &s_acc[(threadIdx.x >> 5) * EP_CT + sl][0];
                          if (lane < 16) acc = *reinterpret_cast<const float4 *>(sa + 4 * l16);
                          if (lane == 0) sc = sa[D];
                          return;
                      }
                      float loss = 0.f;
                      for_each_edge_smem<TripA, EP_BPR_A_UNROLL>(
                          n, half,
                          [&](int e) {
                              TripA it;
                              it.dst = -1; it.t = 0; it.ng = 0; it.rp = 0.f; it.rn = 0.f;
                              it.vp = f4zero(); it.vn = f4zero();
                              if (e >= 0) {
                                  it.dst = wc.nbr[DIR_OUT][sl][e];
                                  it.ng = wc.ng[sl][e];
                                  // three independent loads: rnorm of an inactive negative is stale and not used
                                  it.rp = __ldcg(a.rnorm + it.dst);
                                  const int st_ng = __ldcg(act_cur + it.ng);
                                  const float rn_ng = __ldcg(a.rnorm + it.ng);
                                  it.rn = st_ng != t ? -1.f : rn_ng;                    // inactive: formed in bpr_math
                              }
                              return it;
                          },
                          [&](int, TripA &it) {
                              if (it.dst >= 0) {
                                  it.vp = ldcg4(F4 + (size_t)it.dst * D4 + l16);
                                  if (it.rn < 0.f) it.vn = f4scale(c0, ldcg4(w.row4(it.ng) + l16));
                                  else it.vn = ldcg4(F4 + (size_t)it.ng * D4 + l16);
                              }
                          },
                          [&](int, TripA &it) {
                              const BprOut o = bpr_math(fu, it.vp, it.vn, it.rp, it.rn);
                              f4fma(acc, o.s * it.rp, it.vp);
                              f4fma(acc, -o.s * o.rn, it.vn);
                              if (it.dst >= 0) {
                                  if (l16 == 0) {
                                      loss += o.sp;
                                      sc += o.s * (o.cp - o.cn);
                                      atomicAdd(a.neg_count + (it.ng - U), 1);
                                  }
                                  bpr_item_grads(G4, it.dst, it.ng, fu, it.vp, it.vn, it.rp, o, invP, l16);
                              }
                          });
                      ex0 += warp_sum(loss);
                  },
                  [&](int row, int, int, const float4 &A, float B) {
                      float4 g = A;
                      f4fma(g, -B, fu_row);
                      g = f4scale(ru_row * invP, g);
                      if (lane < 16) G4[(size_t)row * D4 + l16] = g;
                  }, b, ps);
        cta_add2(s_red, lane, ex0, 0.f, acc_cur + 0, nullptr);
        grid_barrier(bar_main, tgt_main, nmain, b, ps);
        stamp(a.prof, b, ps, gtid);

        // ---- G: backward layers 1..K (Horner) ------------------------------------------------------
        // Layer 1 gathers dL/dfinal itself with the weight dis[target] per edge (the item rows' dL/dfinal was summed by
        // atomics in E, so there is no owner who could have written a pre-scaled copy); layers 2..K gather z = dis (.) h.
        ex0 = 0.f;
        for (int j = 1; j <= K; ++j) {
            const float *src = j == 1 ? a.G : ((j & 1) ? Z1 : Z0);
            float *dst = (j & 1) ? Z0 : Z1;
            const bool last = j == K;
            run_tasks<ST_WGT, false>(wc, DIR_OUT, d, 0, d.n_out_tasks, gw, nw, lane, U,
                      [&](int, int n, int sl, float4 &acc, float &) {
                          const int32_t *cn = wc.nbr[DIR_OUT][sl];
                          const float4 *x4 = reinterpret_cast<const float4 *>(src);
                          if (j == 1) {
                              const float *cw = wc.wgt[DIR_OUT][sl];
                              for_each_edge_smem<Nbr, EP_GATHER_UNROLL>(
                                  n, half,
                                  [&](int e) { Nbr it; it.nbr = e >= 0 ? cn[e] : -1; it.w = e >= 0 ? cw[e] : 0.f; it.v = f4zero(); return it; },
                                  [&](int, Nbr &it) { if (it.nbr >= 0) it.v = ldcg4(x4 + (size_t)it.nbr * D4 + l16); },
                                  [&](int, Nbr &it) { f4fma(acc, it.w, it.v); });
                          } else {
                              for_each_edge_smem<Nbr, EP_GATHER_UNROLL>(
                                  n, half,
                                  [&](int e) { Nbr it; it.nbr = e >= 0 ? cn[e] : -1; it.w = 0.f; it.v = f4zero(); return it; },
                                  [&](int, Nbr &it) { if (it.nbr >= 0) it.v = ldcg4(x4 + (size_t)it.nbr * D4 + l16); },
                                  [&](int, Nbr &it) { f4add(acc, it.v); });
                          }
                      },
                      [&](int row, int din, int dout, const float4 &S, float) {
                          const float dd = dis_of(din);
                          float4 hh = ldcg4(G4 + (size_t)row * D4 + l16);
                          f4fma(hh, dd, S);
                          if (!last) {
                              if (lane < 16) reinterpret_cast<float4 *>(dst)[(size_t)row * D4 + l16] = f4scale(dd, hh);
                          } else {
                              float4 g = f4scale(c0, hh);
                              const int cnt = row < U ? dout : din + __ldcg(a.neg_count + row - U);
                              const float4 e = ldcg4(w.row4(row) + l16);
                              f4fma(g, reg_coef * (float)cnt, e);
                              ex0 += (float)cnt * half_sum(f4dot(e, e));
                              if (lane < 16) reinterpret_cast<float4 *>(a.grad)[(size_t)row * D4 + l16] = g;
                              ex1 += half_sum(f4dot(g, g));
                          }
                      }, b, ps);
            if (j == 1) {
                // the INACTIVE negatives' gradient rows (nothing propagates to a row without edges): grad = G/(K+1)^2 + reg.
                // Their dL/dfinal rows and the histogram are complete since the barrier after E; no task gathers them.
                const int cnt = __ldcg(cnt_cur);
                for (int base = gw * 2; base < cnt; base += nw * 2) {
                    const int idx = base + half;
                    float4 g = f4zero();
                    float reg = 0.f;
                    if (idx < cnt) {
                        const int item = __ldcg(list_cur + idx), row = U + item;
                        const float4 gf = ldcg4(G4 + (size_t)row * D4 + l16);        // three independent loads
                        const int c = __ldcg(a.neg_count + item);
                        const float4 e = ldcg4(w.row4(row) + l16);
                        g = f4scale(c0, gf);
                        G4[(size_t)row * D4 + l16] = f4zero();
                        if (c) {
                            f4fma(g, reg_coef * (float)c, e);
                            reg = (float)c * f4dot(e, e);
                        }
                        reinterpret_cast<float4 *>(a.grad)[(size_t)row * D4 + l16] = g;
                    }
                    ex1 += warp_sum(f4dot(g, g));
                    ex0 += warp_sum(reg);
                }
            }
            if (last) cta_add2(s_red, lane, ex0, ex1, acc_cur + 1, acc_cur + 2);
            grid_barrier(bar_main, tgt_main, nmain, b, ps);
            stamp(a.prof, b, ps, gtid);
        }

        // ---- J: clip + Adam step t on the touched rows, loss; then fetch the next step's work --------
        {
            const AdamScalars as = adam_scalars_tab(a.h, t);
            const float clip = clip_coef(a.h, __ldcg(acc_cur + 2));
            const float4 *gr = reinterpret_cast<const float4 *>(a.grad);
            // one half-warp per active row, found through this warp's own in-tasks (first part speaks for the row)
            int ci = half;
            for (int ti = gw + half * nw; ti < d.n_in_tasks; ti += 2 * nw, ci += 2) {
                int row, part;
                if (ci < EP_CT) { row = wc.ta[DIR_IN][ci].x; part = wc.tc[DIR_IN][ci].x; }
                else { row = __ldg(&d.in_tasks[ti].row); part = __ldg(&d.in_tasks[ti].part); }
                if (part == 0) adam_row(row, lane, t, w, a.m, a.v, gr, G4, a.row_step, a.neg_count, clip, as);
            }
            const int cnt = __ldcg(cnt_cur);
            for (int idx = gw * 2 + half; idx < cnt; idx += nw * 2)
                adam_row(U + __ldcg(list_cur + idx), lane, t, w, a.m, a.v, gr, G4, a.row_step, a.neg_count, clip, as);
            if (gtid == 0) {
                const double p = (double)d.P;
                d.loss_out[0] = (float)(-__ldcg(acc_cur + 0) / (10.0 * p) + (double)a.bpr_coeff * __ldcg(acc_cur + 1) / (64.0 * p));
                *a.step = (long long)t;
            }
        }
        barrier_arrive(bar_all, b, ps);
        if (b + 1 < a.num_steps) fill_cache(wc, a.steps[b + 1], gw, nw, lane, U);   // overlaps the wait
        EP_TRACE_AT(b, ps + 1, 0);                                                   // cache filled
        barrier_wait(bar_all, tgt_all, nblocks, b, ps);
        stamp(a.prof, b, ps, gtid);
    }
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace ep
}  // namespace lgcn

extern "C" size_t lgcn_train_steps_workspace_bytes(int64_t num_steps) {
    using namespace lgcn::ep;
    const size_t steps = (size_t)(num_steps > 0 ? num_steps : 1);
    return align256(sizeof(StepDesc) * steps) + 512 + 128 * steps;       // descriptors, state, phase stamps
}

extern "C" int lgcn_train_steps_sparse(const lgcn_graph *graphs, int64_t num_steps, float *user_w, float *item_w, int K,
                                       const int64_t *neg, float bpr_coeff, const lgcn_adam *opt,
                                       const lgcn_step_buffers *buf, float *loss_out, void *workspace,
                                       size_t workspace_bytes, void *stream) {
    using namespace lgcn;
    using namespace lgcn::ep;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(graphs && user_w && item_w && neg && opt && buf && loss_out && workspace, LGCN_E_INVALID,
                 "train_steps_sparse: null argument");
    LGCN_REQUIRE(num_steps >= 1 && num_steps < (1 << 24), LGCN_E_INVALID, "train_steps_sparse: %lld steps", (long long)num_steps);
    LGCN_REQUIRE(K >= 1 && K <= 4, LGCN_E_INVALID, "train_steps_sparse: num_layers %d outside [1,4]", K);
    LGCN_REQUIRE(opt->row_step && opt->m && opt->v && opt->step && opt->bc_table && opt->bc_len > num_steps, LGCN_E_INVALID,
                 "train_steps_sparse: optimiser state (row_step, m, v, step, bc_table covering the run) missing");
    LGCN_REQUIRE(buf->final_emb && buf->rnorm && buf->grad_final && buf->grad_e0 && buf->neg_count && buf->trip_scratch &&
                 buf->act_stamp, LGCN_E_INVALID, "train_steps_sparse: step buffers missing");
    LGCN_REQUIRE(workspace_bytes >= lgcn_train_steps_workspace_bytes(num_steps), LGCN_E_WORKSPACE,
                 "train_steps_sparse: workspace %zu < %zu", workspace_bytes, lgcn_train_steps_workspace_bytes(num_steps));
    const int N = graphs[0].num_nodes, U = graphs[0].num_users;
    const size_t need = (size_t)(K - 1 > 2 ? K - 1 : 2) * (size_t)N * D * sizeof(float);
    LGCN_REQUIRE(buf->work && buf->work_bytes >= need, LGCN_E_WORKSPACE, "train_steps_sparse: work %zu < %zu",
                 buf->work_bytes, need);
    std::vector<StepDesc> descs((size_t)num_steps);
    const int64_t *np = neg;
    for (int64_t b = 0; b < num_steps; ++b) {
        const lgcn_graph &g = graphs[b];
        LGCN_REQUIRE(g.num_nodes == N && g.num_users == U, LGCN_E_INVALID, "train_steps_sparse: graph %lld has another shape",
                     (long long)b);
        LGCN_REQUIRE(g.num_triplets > 0, LGCN_E_INVALID, "train_steps_sparse: batch %lld has no user->movie edge", (long long)b);
        LGCN_REQUIRE(g.in_tasks && g.out_tasks && g.partials && g.slot_counters && g.dis, LGCN_E_INVALID,
                     "train_steps_sparse: graph %lld not built", (long long)b);
        StepDesc &d = descs[(size_t)b];
        d.in_tasks = g.in_tasks; d.out_tasks = g.out_tasks;
        d.in_nbr = g.in_nbr; d.in_trip = g.in_trip; d.out_nbr = g.out_nbr; d.out_trip = g.out_trip;
        d.dis = g.dis;
        d.partials = g.partials; d.slot_counters = g.slot_counters;
        d.neg = np; d.loss_out = loss_out + b; d.P = g.num_triplets;
        d.n_in_tasks = g.n_in_tasks; d.n_out_tasks = g.n_out_tasks;
        d.n_in_user_tasks = g.n_in_user_tasks; d.n_out_user_tasks = g.n_out_user_tasks;
        d.num_edges = g.num_edges;
        np += g.num_triplets;
    }
    char *ws = (char *)workspace;
    const size_t desc_bytes = sizeof(StepDesc) * (size_t)num_steps;
    char *state = ws + align256(desc_bytes);
    LGCN_CUDA(cudaMemcpyAsync(ws, descs.data(), desc_bytes, cudaMemcpyHostToDevice, st));   // pageable: staged before return
    LGCN_CUDA(cudaMemsetAsync(state, 0, 512, st));

    EpochArgs a{};
    a.steps = (const StepDesc *)ws;
    a.num_steps = (int)num_steps; a.K = K;
    a.user_w = user_w; a.item_w = item_w; a.num_users = U; a.num_items = N - U;
    a.m = reinterpret_cast<float4 *>(opt->m); a.v = reinterpret_cast<float4 *>(opt->v);
    a.row_step = opt->row_step; a.step = opt->step; a.h = make_hyper(opt);
    a.final_emb = buf->final_emb; a.rnorm = buf->rnorm; a.G = buf->grad_final; a.grad = buf->grad_e0; a.work = buf->work;
    a.neg_count = buf->neg_count;
    a.act_stamp = buf->act_stamp; a.neg_flag = buf->act_stamp + 2 * (size_t)N; a.neg_list = a.neg_flag + 2 * (size_t)(N - U);
    a.scratch = buf->trip_scratch;
    a.accum = (double *)state; a.counts = (int32_t *)(state + 64); a.bar = (unsigned *)(state + 128);   // [0], [32], [64]: 128 bytes apart
    a.bpr_coeff = bpr_coeff;
    static const bool want_prof = getenv("LGCN_EPOCH_PROF") != nullptr;   // tools/epoch_breakdown.py
    a.prof = want_prof ? (long long *)(state + 512) : nullptr;

    // launch geometry, cached per device (a process may drive several GPUs)
    static int grid_of[64], sms_of[64], per_sm_of[64];
    int dev = 0;
    LGCN_CUDA(cudaGetDevice(&dev));
    LGCN_REQUIRE(dev >= 0 && dev < 64, LGCN_E_INVALID, "train_steps_sparse: device ordinal %d", dev);
    if (grid_of[dev] == 0) {
        int sms = 0, per_sm = 0, coop = 0;
        LGCN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        LGCN_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        LGCN_REQUIRE(coop, LGCN_E_CUDA, "train_steps_sparse: device does not support cooperative launches");
        LGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, epoch_kernel, EP_THREADS, 0));
        LGCN_REQUIRE(per_sm >= 1, LGCN_E_CUDA, "train_steps_sparse: kernel does not fit on an SM");
        sms_of[dev] = sms;
        per_sm_of[dev] = per_sm >= EP_CTAS_PER_SM ? EP_CTAS_PER_SM : per_sm;
        grid_of[dev] = sms * per_sm_of[dev];             // the cooperative launch fills every SM with ctas_per_sm CTAs
    }
    const int grid = grid_of[dev], num_sms = sms_of[dev], ctas_per_sm = per_sm_of[dev];
    // helpers: whole SMs (LGCN_EPOCH_HELPER_SMS: tuning aid)
    static const int hsms_env = getenv("LGCN_EPOCH_HELPER_SMS") ? atoi(getenv("LGCN_EPOCH_HELPER_SMS")) : -1;
    int helper_sms = hsms_env >= 1 ? hsms_env : (num_sms * 7) / 20;
    if (helper_sms > num_sms - 1) helper_sms = num_sms - 1;
    if (helper_sms < 1) helper_sms = 1;
    a.num_helpers = num_steps > 1 && num_sms > 1 ? helper_sms * ctas_per_sm : 0;
    a.main_sms = num_sms - helper_sms;
#ifdef EP_TRACE
    {
        const char *tp = getenv("LGCN_EPOCH_TRACE_PTR");
        long long *ptr = tp ? (long long *)strtoull(tp, nullptr, 0) : nullptr;
        LGCN_CUDA(cudaMemcpyToSymbolAsync(g_trace, &ptr, sizeof(ptr), 0, cudaMemcpyHostToDevice, st));
        const char *tw = getenv("LGCN_EPOCH_TRACE_WARP_PTR");
        long long *wptr = tw ? (long long *)strtoull(tw, nullptr, 0) : nullptr;
        LGCN_CUDA(cudaMemcpyToSymbolAsync(g_trace_warp, &wptr, sizeof(wptr), 0, cudaMemcpyHostToDevice, st));
        const char *tf = getenv("LGCN_EPOCH_TRACE_FINE_PTR");
        long long *fptr = tf ? (long long *)strtoull(tf, nullptr, 0) : nullptr;
        LGCN_CUDA(cudaMemcpyToSymbolAsync(g_trace_fine, &fptr, sizeof(fptr), 0, cudaMemcpyHostToDevice, st));
        const char *tsm = getenv("LGCN_EPOCH_TRACE_SMID_PTR");
        int *sptr = tsm ? (int *)strtoull(tsm, nullptr, 0) : nullptr;
        LGCN_CUDA(cudaMemcpyToSymbolAsync(g_trace_smid, &sptr, sizeof(sptr), 0, cudaMemcpyHostToDevice, st));
    }
#endif
    void *params[] = {(void *)&a};
    LGCN_CUDA(cudaLaunchCooperativeKernel((const void *)epoch_kernel, dim3(grid), dim3(EP_THREADS), params, 0, st));
    return LGCN_OK;
}
