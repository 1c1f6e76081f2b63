// A whole RUN of sparse (touched-rows) training steps in ONE persistent cooperative launch.
//
// Replaces, for a sequence of Cluster-GCN batches, the loop of /root/reference/utils/train_test.py:
// 86-101 (`for batch in train_loader:` zero_grad, forward, bpr_loss, backward, clip_grad_norm_,
// Adam.step).  Arithmetic per step is that of lgcn_train_step_sparse (sparse_step.cu): only the rows
// a batch touches are visited, the zero-gradient Adam updates of the others are replayed exactly when
// they are next touched.
//
// Why one kernel: a median ML-25M cluster batch has 2.2 k active rows and 7 k edges; as 16 launches
// per step the epoch is bound by launch/drain latency and by dependent COLD misses (task -> index ->
// row), ~185 us per step measured.  Here one CTA per SM stays resident for the whole run, phases are
// separated by device-wide barriers, per-node metadata rides in the task descriptors (deg_in /
// deg_out) so no O(N) array is touched, layer 1 gathers a PRE-SCALED table (no per-edge normalisation
// gather), and the next step's descriptors and index arrays are prefetched into L2.
//
// Two CTA roles.  The replay of pending zero-gradient Adam steps is a long SEQUENTIAL chain per row (a
// user row is touched once per epoch: ~100 steps, each with an IEEE sqrt and two divisions per element;
// ~20 us for one warp) -- the same arithmetic dense Adam performs, but on the critical path if done when
// the row is needed.  So:
//   HELPER CTAs work one step AHEAD: during step b they prepare step b+1 -- stamp its active rows,
//     collect its distinct inactive negatives (the run's negatives are sampled up front), and bring all
//     those rows up to date, except rows step b touches itself (phase J leaves them up to date).
//   MAIN CTAs run the phases of step b, separated by barriers among themselves:
//     A  y0 = dis (.) e0 for the active rows                                                    |
//     B  forward layers 1..K (barrier after each; the last forms the layer mean and 1/||.||)    |
//     E  BPR over user rows (loss, user-row gradient, negative-item gradient by vector atomics) |
//     F  BPR over item rows (positive-item gradient, owner computes); both write dis (.) G too;
//        + the gradient rows of the inactive negatives                                          |
//     G  backward layers 1..K (barrier after each)                                              |
//     J  clip + Adam step on the touched rows, restore the all-zero invariants, loss, prefetch
//   Everybody meets at the end of the step.  Step 0 is prepared by all CTAs before the loop.
// Stamp / list arrays are double-buffered by step parity so that preparing b+1 never disturbs b.
//
// Memory rules inside the kernel: everything another SM may have written earlier in the launch is read
// with ld.global.cg (L2, the coherence point); __ldg only for data that is immutable for the whole
// launch (task lists, index arrays, negatives, the bias-correction table).
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
#ifndef EP_BPR_B_UNROLL
#define EP_BPR_B_UNROLL 4
#endif
#ifndef EP_CTAS_PER_SM
#define EP_CTAS_PER_SM 3
#endif
constexpr int EP_WARPS = 8;        // EP_CTAS_PER_SM CTAs per SM: one of them a HELPER, the others MAIN
constexpr int EP_THREADS = EP_WARPS * 32;

struct StepDesc {
    const lgcn_task *in_tasks, *out_tasks;
    const int32_t *in_nbr, *in_trip, *out_nbr, *out_trip;
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
    int num_helpers;     // CTAs [gridDim.x - num_helpers, gridDim.x) prepare the next step
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

// Arrive (release) on a monotonic counter, wait (acquire) until `count` more CTAs have arrived than at
// the previous barrier on it.  bar.sync before/after extends the ordering to the whole CTA.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned &target, unsigned count) {
    __syncthreads();
    target += count;
    if (threadIdx.x == 0) {
        unsigned cur;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(bar), "r"(1u) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(bar) : "memory");
        } while ((int)(cur - target) < 0);
    }
    __syncthreads();
}

// for_each_edge (rowtask.cuh) with the 32-edge chunk loop left ROLLED: the persistent kernel holds every
// phase's code at once, and sixteen inlined copies of each edge body made it 220 KB -- every phase then
// started with instruction-cache misses.  Same traversal order, same arithmetic.
template <class Item, int kUnroll, class Fetch, class Load, class Apply>
__device__ __forceinline__ void for_each_edge_rolled(int begin, int end, int lane, Fetch fetch, Load load, Apply apply) {
    const int half = lane >> 4;
    Item mine = fetch(begin + lane < end ? begin + lane : -1);
    for (int base = begin; base < end; base += 32) {
        const int n = min(32, end - base);
        const int nb = base + 32 + lane;
        Item next = fetch(nb < end ? nb : -1);
#pragma unroll 1
        for (int j = 0; j < n; j += 2 * kUnroll) {
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

struct Tab {                        // mutable e0 = (user_w, item_w)
    float *user, *item;
    int num_users;
    __device__ __forceinline__ float4 *row4(int r) const {
        float *p = r < num_users ? user + (size_t)r * D : item + (size_t)(r - num_users) * D;
        return reinterpret_cast<float4 *>(p);
    }
};

struct Nbr {
    int nbr;
    float4 v;
    __device__ __forceinline__ Nbr shfl(int src_lane) const {
        Nbr r;
        r.nbr = __shfl_sync(FULL, nbr, src_lane);
        r.v = f4zero();
        return r;
    }
};

__device__ __forceinline__ float dis_of(int deg_in) { return deg_in > 0 ? 1.0f / sqrtf((float)deg_in) : 0.f; }

// raw[c] = sum over the task's edges of x[nbr]   (x written earlier in this launch -> ld.cg)
__device__ __noinline__ void gather_sum(const int32_t *__restrict__ nbr, const float *x, int begin, int end, int lane,
                                           float4 &acc) {
    const int l16 = lane & 15;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    for_each_edge_rolled<Nbr, EP_GATHER_UNROLL>(
        begin, end, lane,
        [&](int e) { Nbr it; it.nbr = e >= 0 ? __ldg(nbr + e) : -1; it.v = f4zero(); return it; },
        [&](int, Nbr &it) { if (it.nbr >= 0) it.v = ldcg4(x4 + (size_t)it.nbr * D4 + l16); },
        [&](int, Nbr &it) { f4add(acc, it.v); });
}

// Walk tasks [tb,te) round-robin over all warps of the grid; same split-row protocol as rowtask_kernel.
template <class Acc, class Epi>
__device__ __forceinline__ void run_tasks(const lgcn_task *__restrict__ tasks, int tb, int te, float *partials,
                                          int *counters, int gw, int nw, int lane, Acc accumulate, Epi epilogue) {
    for (int t = tb + gw; t < te; t += nw) {
        const int4 ta = __ldg(reinterpret_cast<const int4 *>(tasks + t));
        const int4 tc = __ldg(reinterpret_cast<const int4 *>(tasks + t) + 1);
        const int row = ta.x, begin = ta.y, end = ta.z, slot = ta.w, part = tc.x, nparts = tc.y;
        float4 acc = f4zero();
        float sc = 0.f;
        accumulate(row, begin, end, acc, sc);
        f4add(acc, f4shfl_xor16(acc));
        sc = warp_sum(sc);
        bool run = slot < 0;
        if (slot >= 0) {
            float *p = partials + (size_t)slot * PARTIAL_STRIDE;
            if (lane < 16) reinterpret_cast<float4 *>(p)[lane] = acc;
            if (lane == 16) p[D] = sc;
            __threadfence();
            const int first = slot - part;
            int old = 0;
            if (lane == 0) old = atomicAdd(counters + first, 1);
            old = __shfl_sync(FULL, old, 0);
            if (old == nparts - 1) {
                __threadfence();
                acc = f4zero();
                sc = 0.f;
                for (int i = 0; i < nparts; ++i) {
                    const float *q = partials + (size_t)(first + i) * PARTIAL_STRIDE;
                    f4add(acc, __ldcg(reinterpret_cast<const float4 *>(q) + (lane & 15)));
                    sc += __ldcg(q + D);
                }
                if (lane == 0) counters[first] = 0;
                run = true;
            }
        }
        if (run) epilogue(row, tc.z, tc.w, acc, sc);
    }
}

// Bring `row` up to optimiser step `target` by replaying zero-gradient Adam steps (adam_replay_kernel's
// arithmetic).  One half-warp per row; returns this lane's 4 weights (zero for invalid rows).
__device__ __forceinline__ float4 replay_row(bool valid, int row, int lane, int target, const Tab &w, float4 *m, float4 *v,
                                             int32_t *row_step, const AdamHyper &h) {
    const int l16 = lane & 15;
    const int from = valid ? __ldcg(row_step + row) : target;
    const bool need = valid && from < target;
    float4 *pp = w.row4(valid ? row : 0) + l16;
    const size_t o = (size_t)(valid ? row : 0) * D4 + l16;
    float4 p4 = f4zero(), m4 = f4zero(), v4 = f4zero();
    if (valid) p4 = ldcg4(pp);
    if (need) { m4 = ldcg4(m + o); v4 = ldcg4(v + o); }
    const bool live = m4.x != 0.f || m4.y != 0.f || m4.z != 0.f || m4.w != 0.f ||
                      v4.x != 0.f || v4.y != 0.f || v4.z != 0.f || v4.w != 0.f;
    const unsigned half_mask = 0xffffu << (lane & 16);
    const bool any_live = (__ballot_sync(FULL, live) & half_mask) != 0u;
    if (need && any_live) {
        // iterations only chain through one fma each for p, m, v: unrolling lets the sqrt / division
        // sequences of neighbouring steps overlap
        const float4 zero = f4zero();
#pragma unroll 2
        for (int t = from + 1; t <= target; ++t) {
            const AdamScalars a = adam_scalars_tab(h, t);
            adam_vec(p4, m4, v4, zero, 1.0f, a);
        }
        *pp = p4; m[o] = m4; v[o] = v4;
    }
    if (need && l16 == 0) row_step[row] = target;
    return p4;
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
enum { RP_ACTIVE = 0, RP_LIST = 1 };

__device__ __noinline__ void replay_rows(int mode, const lgcn_task *__restrict__ tasks, const int32_t *list, int count,
                                         int target, Tab w, float4 *m, float4 *v, int32_t *row_step, int32_t *act_next,
                                         const int32_t *act_p, const int32_t *flag_p, AdamHyper h, int gw, int nw, int lane) {
    const int l16 = lane & 15, half = lane >> 4;
    for (int base = gw * 2; base < count; base += nw * 2) {
        const int i = base + half;
        bool valid = i < count;
        int row = 0;
        if (valid) {
            if (mode == RP_LIST) {
                row = w.num_users + __ldcg(list + i);
            } else {
                row = __ldg(&tasks[i].row);
                valid = __ldg(&tasks[i].part) == 0;          // first part of a split row speaks for the row
                if (valid && l16 == 0) act_next[row] = target + 1;
            }
            if (valid) valid = !touched_by(row, w.num_users, target, act_p, flag_p);
        }
        replay_row(valid, row, lane, target, w, m, v, row_step, h);
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

struct TripB {
    int u;
    float s, scp, ru;
    float4 vu;
    __device__ __forceinline__ TripB shfl(int src) const {
        TripB r;
        r.u = __shfl_sync(FULL, u, src);
        r.s = __shfl_sync(FULL, s, src);
        r.scp = __shfl_sync(FULL, scp, src);
        r.ru = __shfl_sync(FULL, ru, src);
        r.vu = f4zero();
        return r;
    }
};

__global__ void __launch_bounds__(EP_THREADS, EP_CTAS_PER_SM) epoch_kernel(const EpochArgs a) {
    const int lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4;
    const int gw_all = blockIdx.x * EP_WARPS + (threadIdx.x >> 5), nw_all = gridDim.x * EP_WARPS;
    const unsigned nblocks = gridDim.x, nhelp = a.num_helpers, nmain = gridDim.x - a.num_helpers;
    const bool helper = blockIdx.x >= nmain;
    // main CTAs index their work among themselves; helpers among themselves
    const int gw = helper ? gw_all - (int)nmain * EP_WARPS : gw_all;
    const int nw = helper ? (int)nhelp * EP_WARPS : (int)nmain * EP_WARPS;
    const int gtid = blockIdx.x * EP_THREADS + threadIdx.x, nthreads = (int)nmain * EP_THREADS;
    unsigned tgt_all = 0, tgt_main = 0, tgt_help = 0;
    unsigned *const bar_all = a.bar, *const bar_main = a.bar + 32, *const bar_help = a.bar + 64;
    const int K = a.K, U = a.num_users;
    const size_t n = (size_t)a.num_users + (size_t)a.num_items;
    const Tab w{a.user_w, a.item_w, U};
    const float c0 = 1.0f / (float)((K + 1) * (K + 1));
    float *const y0 = a.grad;                               // dead until the last backward layer writes grad
    float *const Z0 = a.work, *const Z1 = a.work + n * D;   // backward tables (the forward's y tables are dead then)
    float4 *const G4 = reinterpret_cast<float4 *>(a.G);
    const float4 *const F4 = reinterpret_cast<const float4 *>(a.final_emb);
    const long long t0 = *a.step;                           // only rewritten after the first barrier of a step

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
        if (pgw == 0 && lane == 0) *cnt_s = 0;
        replay_rows(RP_ACTIVE, sd.in_tasks, nullptr, sd.n_in_tasks, ts - 1, w, a.m, a.v, a.row_step, act_s, act_p, flag_p, a.h,
                    pgw, pnw, lane);
        grid_barrier(pbar, ptgt, pcount);
        for (long long base = (long long)pgw * 32; base < sd.P; base += (long long)pnw * 32) {
            const long long tt = base + lane;
            if (tt < sd.P) {
                const int i = (int)__ldg(sd.neg + tt);
                if (__ldcg(act_s + U + i) != ts && atomicExch(flag_s + i, ts) != ts) list_s[atomicAdd(cnt_s, 1)] = i;
            }
        }
        grid_barrier(pbar, ptgt, pcount);
        replay_rows(RP_LIST, nullptr, list_s, __ldcg(cnt_s), ts - 1, w, a.m, a.v, a.row_step, nullptr, act_p, flag_p, a.h, pgw,
                    pnw, lane);
    };
    prepare(0, gw_all, nw_all, bar_all, tgt_all, nblocks);
    grid_barrier(bar_all, tgt_all, nblocks);

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
            if (b + 1 < a.num_steps) prepare(b + 1, gw, nw, bar_help, tgt_help, nhelp);
            grid_barrier(bar_all, tgt_all, nblocks);     // end of the step
            continue;
        }

        // ---- A: y0 = dis (.) e0 for the active rows (already at step t-1) ------------------------
        if (gtid == 0) {
            double *nx = a.accum + ((b + 1) & 1) * 4;
            nx[0] = 0.0; nx[1] = 0.0; nx[2] = 0.0; nx[3] = 0.0;
        }
        for (int ti = gw * 2 + half; ti < d.n_in_tasks; ti += nw * 2) {
            const int4 ta = __ldg(reinterpret_cast<const int4 *>(d.in_tasks + ti));
            const int4 tc = __ldg(reinterpret_cast<const int4 *>(d.in_tasks + ti) + 1);
            if (tc.x == 0)
                reinterpret_cast<float4 *>(y0)[(size_t)ta.x * D4 + l16] = f4scale(dis_of(tc.z), ldcg4(w.row4(ta.x) + l16));
        }
        grid_barrier(bar_main, tgt_main, nmain);
        stamp(a.prof, b, ps, gtid);

        // ---- B: forward layers 1..K --------------------------------------------------------------
        auto fwd_layer = [&](int k) {
            const float *src = k == 1 ? y0 : a.work + (size_t)(k - 2) * n * D;
            float *dst = a.work + (size_t)(k - 1) * n * D;
            const bool last = k == K;
            run_tasks(d.in_tasks, 0, d.n_in_tasks, d.partials, d.slot_counters, gw, nw, lane,
                      [&](int, int begin, int end, float4 &acc, float &) { gather_sum(d.in_nbr, src, begin, end, lane, acc); },
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
                      });
        };
        for (int k = 1; k <= K; ++k) {
            fwd_layer(k);
            grid_barrier(bar_main, tgt_main, nmain);
            stamp(a.prof, b, ps, gtid);
        }

        // ---- E: BPR over user rows -------------------------------------------------------------
        float ex0 = 0.f, ex1 = 0.f;
        run_tasks(d.out_tasks, 0, d.n_out_user_tasks, d.partials, d.slot_counters, gw, nw, lane,
                  [&](int row, int begin, int end, float4 &acc, float &sc) {
                      const float ru = __ldcg(a.rnorm + row);
                      const float4 fu = f4scale(ru, ldcg4(F4 + (size_t)row * D4 + l16));
                      float loss = 0.f;
                      for_each_edge_rolled<TripA, EP_BPR_A_UNROLL>(
                          begin, end, lane,
                          [&](int e) {
                              TripA it;
                              it.dst = -1; it.t = 0; it.ng = 0; it.rp = 0.f; it.rn = 0.f;
                              it.vp = f4zero(); it.vn = f4zero();
                              if (e >= 0) {
                                  it.dst = __ldg(d.out_nbr + e);
                                  it.t = __ldg(d.out_trip + e);
                                  it.ng = (int)__ldg(d.neg + it.t) + U;
                                  it.rp = __ldcg(a.rnorm + it.dst);
                                  if (__ldcg(act_cur + it.ng) != t) it.rn = -1.f;      // inactive: formed below
                                  else it.rn = __ldcg(a.rnorm + it.ng);
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
                              const bool valid = it.dst >= 0;
                              const float n2 = half_sum(f4dot(it.vn, it.vn));
                              if (it.rn < 0.f) it.rn = 1.0f / sqrtf(n2);
                              const float cp = half_sum(f4dot(fu, it.vp)) * it.rp;
                              const float cn = half_sum(f4dot(fu, it.vn)) * it.rn;
                              const float x = 10.f * (cp - cn);
                              const float sp = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
                              if (valid && l16 == 0) loss += sp;
                              const float s = -1.f / (1.f + expf(-x));
                              f4fma(acc, s * it.rp, it.vp);
                              f4fma(acc, -s * it.rn, it.vn);
                              if (valid) {
                                  if (l16 == 0) {
                                      sc += s * (cp - cn);
                                      reinterpret_cast<float2 *>(a.scratch)[it.t] = make_float2(s, cp);
                                      atomicAdd(a.neg_count + (it.ng - U), 1);
                                  }
                                  const float kk = -s * it.rn * invP;
                                  float4 g = fu;
                                  f4fma(g, -cn * it.rn, it.vn);
                                  g = f4scale(kk, g);
                                  atomicAdd(G4 + (size_t)it.ng * D4 + l16, g);
                              }
                          });
                      ex0 += warp_sum(loss);
                  },
                  [&](int row, int din, int, const float4 &A, float B) {
                      const float ru = __ldcg(a.rnorm + row);
                      const float4 fu = f4scale(ru, ldcg4(F4 + (size_t)row * D4 + l16));
                      float4 g = A;
                      f4fma(g, -B, fu);
                      g = f4scale(ru * invP, g);
                      if (lane < 16) {
                          G4[(size_t)row * D4 + l16] = g;
                          reinterpret_cast<float4 *>(Z1)[(size_t)row * D4 + l16] = f4scale(dis_of(din), g);
                      }
                  });
        if (lane == 0 && ex0 != 0.f) atomicAdd(acc_cur + 0, (double)ex0);
        grid_barrier(bar_main, tgt_main, nmain);
        stamp(a.prof, b, ps, gtid);

        // ---- F: BPR over item rows -------------------------------------------------------------
        run_tasks(d.in_tasks, d.n_in_user_tasks, d.n_in_tasks, d.partials, d.slot_counters, gw, nw, lane,
                  [&](int, int begin, int end, float4 &acc, float &sc) {
                      for_each_edge_rolled<TripB, EP_BPR_B_UNROLL>(
                          begin, end, lane,
                          [&](int e) {
                              TripB it;
                              it.u = -1; it.s = 0.f; it.scp = 0.f; it.ru = 0.f; it.vu = f4zero();
                              if (e >= 0) it.u = __ldg(d.in_nbr + e);
                              if (it.u >= U) it.u = -1;
                              if (it.u >= 0) {
                                  const float2 sc2 = __ldcg(reinterpret_cast<const float2 *>(a.scratch) + __ldg(d.in_trip + e));
                                  it.ru = __ldcg(a.rnorm + it.u);
                                  it.s = sc2.x * it.ru;
                                  it.scp = sc2.x * sc2.y;
                              }
                              return it;
                          },
                          [&](int, TripB &it) { if (it.u >= 0) it.vu = ldcg4(F4 + (size_t)it.u * D4 + l16); },
                          [&](int, TripB &it) {
                              f4fma(acc, it.s, it.vu);
                              if (l16 == 0) sc += it.scp;
                          });
                  },
                  [&](int row, int din, int, const float4 &A, float B) {
                      const float rp = __ldcg(a.rnorm + row);
                      const float4 fp = f4scale(rp, ldcg4(F4 + (size_t)row * D4 + l16));
                      float4 g = A;
                      f4fma(g, -B, fp);
                      g = f4scale(rp * invP, g);
                      if (lane < 16) {
                          float4 cur = ldcg4(G4 + (size_t)row * D4 + l16);      // negative-sample contributions (E)
                          f4add(cur, g);
                          G4[(size_t)row * D4 + l16] = cur;
                          reinterpret_cast<float4 *>(Z1)[(size_t)row * D4 + l16] = f4scale(dis_of(din), cur);
                      }
                  });
        // the INACTIVE negatives' gradient rows (nothing propagates to a row without edges): grad = G/(K+1)^2 + reg.
        // Their dL/dfinal rows and the histogram are complete since the barrier after E; this phase has only the
        // item rows' tasks, so the list fills otherwise idle warps.
        {
            float n0 = 0.f, n1 = 0.f;
            const int cnt = __ldcg(cnt_cur);
            for (int base = gw * 2; base < cnt; base += nw * 2) {
                const int idx = base + half;
                float4 g = f4zero();
                float reg = 0.f;
                if (idx < cnt) {
                    const int item = __ldcg(list_cur + idx), row = U + item;
                    g = f4scale(c0, ldcg4(G4 + (size_t)row * D4 + l16));
                    G4[(size_t)row * D4 + l16] = f4zero();
                    const int c = __ldcg(a.neg_count + item);
                    if (c) {
                        const float4 e = ldcg4(w.row4(row) + l16);
                        f4fma(g, reg_coef * (float)c, e);
                        reg = (float)c * f4dot(e, e);
                    }
                    reinterpret_cast<float4 *>(a.grad)[(size_t)row * D4 + l16] = g;
                }
                n1 += warp_sum(f4dot(g, g));
                n0 += warp_sum(reg);
            }
            if (lane == 0 && n0 != 0.f) atomicAdd(acc_cur + 1, (double)n0);
            if (lane == 0 && n1 != 0.f) atomicAdd(acc_cur + 2, (double)n1);
        }
        grid_barrier(bar_main, tgt_main, nmain);
        stamp(a.prof, b, ps, gtid);

        // ---- G: backward layers 1..K (Horner) ------------------------------------------------------
        ex0 = 0.f;
        for (int j = 1; j <= K; ++j) {
            const float *src = (j & 1) ? Z1 : Z0;
            float *dst = (j & 1) ? Z0 : Z1;
            const bool last = j == K;
            run_tasks(d.out_tasks, 0, d.n_out_tasks, d.partials, d.slot_counters, gw, nw, lane,
                      [&](int, int begin, int end, float4 &acc, float &) { gather_sum(d.out_nbr, src, begin, end, lane, acc); },
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
                      });
            if (last) {
                if (lane == 0 && ex0 != 0.f) atomicAdd(acc_cur + 1, (double)ex0);
                if (lane == 0 && ex1 != 0.f) atomicAdd(acc_cur + 2, (double)ex1);
            }
            grid_barrier(bar_main, tgt_main, nmain);
        stamp(a.prof, b, ps, gtid);
        }

        // ---- J: clip + Adam step t on the touched rows, loss, prefetch the next step -------------
        {
            const AdamScalars as = adam_scalars_tab(a.h, t);
            const float clip = clip_coef(a.h, __ldcg(acc_cur + 2));
            const float4 *gr = reinterpret_cast<const float4 *>(a.grad);
            for (int base = gw * 2; base < d.n_in_tasks; base += nw * 2) {
                const int ti = base + half;
                if (ti < d.n_in_tasks) {
                    const int4 ta = __ldg(reinterpret_cast<const int4 *>(d.in_tasks + ti));
                    const int part = __ldg(&d.in_tasks[ti].part);
                    if (part == 0) adam_row(ta.x, lane, t, w, a.m, a.v, gr, G4, a.row_step, a.neg_count, clip, as);
                }
            }
            const int cnt = __ldcg(cnt_cur);
            for (int idx = gw * 2 + half; idx < cnt; idx += nw * 2)
                adam_row(U + __ldcg(list_cur + idx), lane, t, w, a.m, a.v, gr, G4, a.row_step, a.neg_count, clip, as);
            if (gtid == 0) {
                const double p = (double)d.P;
                d.loss_out[0] = (float)(-__ldcg(acc_cur + 0) / (10.0 * p) + (double)a.bpr_coeff * __ldcg(acc_cur + 1) / (64.0 * p));
                *a.step = (long long)t;
            }
            if (b + 1 < a.num_steps) {
                const StepDesc nx = a.steps[b + 1];
                prefetch_range(nx.in_tasks, sizeof(lgcn_task) * (size_t)nx.n_in_tasks, gtid, nthreads);
                prefetch_range(nx.out_tasks, sizeof(lgcn_task) * (size_t)nx.n_out_tasks, gtid, nthreads);
                prefetch_range(nx.in_nbr, 4 * (size_t)nx.num_edges, gtid, nthreads);
                prefetch_range(nx.out_nbr, 4 * (size_t)nx.num_edges, gtid, nthreads);
                prefetch_range(nx.in_trip, 4 * (size_t)nx.num_edges, gtid, nthreads);
                prefetch_range(nx.out_trip, 4 * (size_t)nx.num_edges, gtid, nthreads);
                prefetch_range(nx.neg, 8 * (size_t)nx.P, gtid, nthreads);
            }
        }
        grid_barrier(bar_all, tgt_all, nblocks);
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
        LGCN_REQUIRE(g.in_tasks && g.out_tasks && g.partials && g.slot_counters, LGCN_E_INVALID,
                     "train_steps_sparse: graph %lld not built", (long long)b);
        StepDesc &d = descs[(size_t)b];
        d.in_tasks = g.in_tasks; d.out_tasks = g.out_tasks;
        d.in_nbr = g.in_nbr; d.in_trip = g.in_trip; d.out_nbr = g.out_nbr; d.out_trip = g.out_trip;
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

    static int grid = 0;
    if (grid == 0) {
        int dev = 0, sms = 0, per_sm = 0, coop = 0;
        LGCN_CUDA(cudaGetDevice(&dev));
        LGCN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        LGCN_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        LGCN_REQUIRE(coop, LGCN_E_CUDA, "train_steps_sparse: device does not support cooperative launches");
        LGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, epoch_kernel, EP_THREADS, 0));
        LGCN_REQUIRE(per_sm >= 1, LGCN_E_CUDA, "train_steps_sparse: kernel does not fit on an SM");
        grid = sms * (per_sm >= EP_CTAS_PER_SM ? EP_CTAS_PER_SM : per_sm);
    }
    static const int helpers_env = getenv("LGCN_EPOCH_HELPERS") ? atoi(getenv("LGCN_EPOCH_HELPERS")) : -1;   // tuning aid
    a.num_helpers = num_steps > 1 ? (helpers_env >= 1 ? helpers_env : (grid * 9) / 20) : 0;
    if (num_steps > 1 && a.num_helpers < 1) a.num_helpers = 1;
    if (a.num_helpers > grid - 1) a.num_helpers = grid - 1;
    void *params[] = {(void *)&a};
    LGCN_CUDA(cudaLaunchCooperativeKernel((const void *)epoch_kernel, dim3(grid), dim3(EP_THREADS), params, 0, st));
    return LGCN_OK;
}
