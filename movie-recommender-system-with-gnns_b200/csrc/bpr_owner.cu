// K3 (full-graph form): BPR-cosine loss and its gradient w.r.t. the final embeddings, OWNER-COMPUTES.
//
// Same arithmetic as bpr.cu (compute_embeddings' six gathers + bpr_loss + autograd,
// /root/reference/utils/train_test.py:18-64,128-132), restructured so that every row of dL/dfinal is
// produced by exactly one warp task and no float atomics are used -- neither for the positive items
// (as before) nor for the sampled negatives.  That is what lets the step shard by node range with no
// reduction of dL/dfinal across GPUs: a rank computes the rows it owns, nothing else.
//
//   user pass   one task per user row of the CSR by source: cos+/cos-, softplus sum, the user-row gradient
//   buckets     the step's negatives (utils/helpers.py:79-80: neg[t] pairs with the t-th user->movie edge)
//               grouped by item: histogram -> scan -> fill -> per-bucket sort by t (=> a fixed summation
//               order: the result is bit-stable run to run, unlike an atomic scatter)
//   neg pass    one warp per owned item: the negative-role gradient from its bucket
//   item pass   one task per item row of the CSR by target: the positive-role gradient, added to the above
//
// A triplet's scalars (s_t = dL/dcos+, cos+_t, cos-_t) are either written by the user pass and read by the
// item passes (kScalars, single GPU: 4 row gathers per triplet) or RECOMPUTED where they are needed from the
// same rows with the same instruction sequence (sharded: 6 row gathers per triplet, bit-identical scalars,
// nothing but embedding rows ever crosses NVLink).
//
// Besides G = dL/dfinal (local rows) the epilogues store zG = dis (.) G into every rank's copy (fused
// all-gather, see common.cuh push4): the first backward layer is then a pure gather-sum of zG.
#include "rowtask.cuh"
#include <limits.h>

namespace lgcn {

__device__ __forceinline__ float bpr_s(float x) { return -1.f / (1.f + expf(-x)); }   // dL_t/dcos+ (x P)

// ---------------------------------------------------------------------------------------------------
// The triplet batch engine.  All tables hold L2-NORMALISED final rows (Fh = final / ||final||, written by the
// last forward layer), so a cosine is a plain dot product and no per-row norm is gathered.
//
// A warp walks entries [begin,end) of an index list, 4 entries per half-warp and batch; per entry it gathers TWO
// rows (i0, i1) -- 16 row loads in flight per warp -- and forms the entry's two cosines against/among them:
//     kMode USER : own = u^,  i0 = p, i1 = n      cos+ = own.Fh[i0]     cos- = own.Fh[i1]
//     kMode POS  : own = p^,  i0 = u, i1 = n      cos+ = Fh[i0].own     cos- = Fh[i0].Fh[i1]
//     kMode NEG  : own = n^,  i0 = u, i1 = p      cos+ = Fh[i0].Fh[i1]  cos- = Fh[i0].own
// The 8 partial dot products of a batch are reduced over the 16 lanes with a TRANSPOSED butterfly (8 shuffles
// instead of 32): afterwards lanes 2k, 2k+1 (mod 8) hold (cos+, cos-) of entry k, so softplus / sigmoid are
// evaluated once per batch, one entry per lane pair, instead of once per entry by every lane.  The same reduction
// tree and the same fma order are used in all three modes => s_t, cos+_t, cos-_t are bit-identical wherever a
// triplet is met (user pass, positive item's pass, negative item's pass), on whatever rank.
// ---------------------------------------------------------------------------------------------------

enum { MODE_USER = 0, MODE_POS = 1, MODE_NEG = 2 };

struct TripEnt {
    int i0, i1, t;       // t < 0: padding
};

template <int kMode, bool kWrite, class Fetch>
__device__ __forceinline__ void trip_batches(int begin, int end, int lane, const float4 *__restrict__ Fh4,
                                             const float4 &own, Fetch fetch, float4 &acc, float &sc, float &loss,
                                             float2 *s_cp, float2 *s_cn) {
    const int half = lane >> 4, l16 = lane & 15;
    const int kk = (l16 >> 1) & 3;                        // the entry of a batch this lane does the scalar math for
    const bool holder = (l16 & 9) == 0;                   // lanes 0,2,4,6 of a half-warp account for entries 0..3
    TripEnt mine = fetch(begin + lane < end ? begin + lane : -1);
    for (int base = begin; base < end; base += 32) {
        const int n = min(32, end - base);
        const int nb = base + 32 + lane;
        TripEnt next = fetch(nb < end ? nb : -1);         // issued before this chunk's gathers (software pipeline)
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            if (j >= n) break;                            // warp-uniform
            int r0[4], r1[4];
            float4 a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int src = j + 2 * k + half;
                r0[k] = __shfl_sync(FULL, mine.i0, src);
                r1[k] = __shfl_sync(FULL, mine.i1, src);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                a[k] = f4zero();
                b[k] = f4zero();
                if (r0[k] >= 0) {
                    a[k] = ldg4(Fh4 + (size_t)r0[k] * D4 + l16);
                    b[k] = ldg4(Fh4 + (size_t)r1[k] * D4 + l16);
                }
            }
            float v[8];                                   // [cos+ 0..3, cos- 0..3], this lane's 4 columns
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if constexpr (kMode == MODE_USER) { v[k] = f4dot(own, a[k]); v[4 + k] = f4dot(own, b[k]); }
                else if constexpr (kMode == MODE_POS) { v[k] = f4dot(a[k], own); v[4 + k] = f4dot(a[k], b[k]); }
                else { v[k] = f4dot(a[k], b[k]); v[4 + k] = f4dot(a[k], own); }
            }
            // transposed butterfly: 8 values x 16 lanes -> lane l holds the total of value (l >> 1)
            float w[4], u2[2], z;
            {
                const bool hi = (l16 & 8) != 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float send = hi ? v[i] : v[i + 4], keep = hi ? v[i + 4] : v[i];
                    w[i] = keep + __shfl_xor_sync(FULL, send, 8);
                }
            }
            {
                const bool hi = (l16 & 4) != 0;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float send = hi ? w[i] : w[i + 2], keep = hi ? w[i + 2] : w[i];
                    u2[i] = keep + __shfl_xor_sync(FULL, send, 4);
                }
            }
            {
                const bool hi = (l16 & 2) != 0;
                const float send = hi ? u2[0] : u2[1], keep = hi ? u2[1] : u2[0];
                z = keep + __shfl_xor_sync(FULL, send, 2);
            }
            z += __shfl_xor_sync(FULL, z, 1);
            const float other = __shfl_xor_sync(FULL, z, 8);
            const float cp = (l16 & 8) ? other : z, cn = (l16 & 8) ? z : other;     // of entry kk, in all 16 lanes
            const int t = __shfl_sync(FULL, mine.t, j + 2 * kk + half);
            const float x = 10.f * (cp - cn);
            const float s = bpr_s(x);
            if (holder && t >= 0) {
                if constexpr (kMode == MODE_USER) {
                    loss += fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));                // softplus
                    sc += s * (cp - cn);
                    if constexpr (kWrite) {
                        s_cp[t] = make_float2(s, cp);
                        s_cn[t] = make_float2(s, cn);
                    }
                } else if constexpr (kMode == MODE_POS) sc += s * cp;
                else sc -= s * cn;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float sk = __shfl_sync(FULL, s, (half << 4) + 2 * k);
                if constexpr (kMode == MODE_USER) { f4fma(acc, sk, a[k]); f4fma(acc, -sk, b[k]); }   // s (p^ - n^)
                else if constexpr (kMode == MODE_POS) f4fma(acc, sk, a[k]);                          // s u^
                else f4fma(acc, -sk, a[k]);                                                          // -s u^
            }
        }
        mine = next;
    }
}

// ---------------------------------------------------------------------------------------------------
// user pass: A = sum_t s_t (p^_t - n^_t), B = sum_t s_t (cos+_t - cos-_t);  dL/dfinal[u] = (A - B u^) / ||u|| / P
// ---------------------------------------------------------------------------------------------------

template <bool kScalars>
struct BprUserOwnOp {
    static constexpr bool kExtras = true;
    double *extra0, *extra1;            // extra0: sum_t softplus(10 (cos+ - cos-))
    const int32_t *out_nbr, *out_trip;
    const int64_t *neg;
    const float *Fh, *rnorm, *dis;
    int num_users;
    float invP;
    float *G, *zG;
    float2 *s_cp, *s_cn;                // kScalars: (s_t, cos+_t) and (s_t, cos-_t) by triplet
    Peers peers;

    __device__ __forceinline__ void accumulate(int row, int begin, int end, int lane, float4 &acc,
                                               float &sc, float &ex0, float &) const {
        const float4 *Fh4 = reinterpret_cast<const float4 *>(Fh);
        const float4 own = ldg4(Fh4 + (size_t)row * D4 + (lane & 15));
        float loss = 0.f;
        trip_batches<MODE_USER, kScalars>(
            begin, end, lane, Fh4, own,
            [&](int e) {
                TripEnt it{-1, -1, -1};
                if (e >= 0) {
                    it.i0 = __ldg(out_nbr + e);
                    it.t = __ldg(out_trip + e);
                    it.i1 = (int)__ldg(neg + it.t) + num_users;
                }
                return it;
            },
            acc, sc, loss, s_cp, s_cn);
        ex0 += warp_sum(loss);
    }
    __device__ __forceinline__ void epilogue(int row, int lane, const float4 &A, float B, float &, float &) const {
        const int l16 = lane & 15;
        const float4 uh = ldg4(reinterpret_cast<const float4 *>(Fh) + (size_t)row * D4 + l16);
        float4 g = A;
        f4fma(g, -B, uh);
        g = f4scale(__ldg(rnorm + row) * invP, g);
        if (lane < 16) {
            reinterpret_cast<float4 *>(G)[(size_t)row * D4 + l16] = g;
            push4(reinterpret_cast<float4 *>(zG) + (size_t)row * D4 + l16, f4scale(__ldg(dis + row), g), peers);
        }
    }
};

// Resident warps per SM the user pass is compiled for (32 = a 64-register cap instead of 110 registers / 16 warps):
// the whole BPR stage 1938 -> 1705 us at ML-25M shape (profiles/r2b_variants.txt), bit-identical results.
#ifndef LGCN_BPR_USER_MINWARPS
#define LGCN_BPR_USER_MINWARPS 32
#endif
template <bool kScalars>
struct MinBlocks<BprUserOwnOp<kScalars>> { static constexpr int value = LGCN_BPR_USER_MINWARPS / WARPS_PER_CTA; };

// ---------------------------------------------------------------------------------------------------
// item passes: A = sum_t sigma s_t u^_t, B = sum_t sigma s_t c_t over the triplets in which the item plays a
// role (sigma = +1, c = cos+ as the positive; sigma = -1, c = cos- as the negative);
// dL/dfinal[i] = (A - B i^) / ||i|| / P.
// ---------------------------------------------------------------------------------------------------

struct ItemTrip {
    int u;
    float w, wc;         // sigma s_t, sigma s_t c_t
    float4 vu;
    __device__ __forceinline__ ItemTrip shfl(int src) const {
        ItemTrip r;
        r.u = __shfl_sync(FULL, u, src);
        r.w = __shfl_sync(FULL, w, src);
        r.wc = __shfl_sync(FULL, wc, src);
        r.vu = f4zero();
        return r;
    }
};

// Walks [begin,end) of an index list: kBucket = false: CSR-by-target positions (u = in_nbr[e], t = in_trip[e], the row
// is the triplet's POSITIVE item); kBucket = true: bucket entries (t = bucket[e], u = trip_user[t], the row is the
// NEGATIVE).  kScalars: (s_t, c_t) come from the user pass; else they are recomputed (trip_batches).
template <bool kScalars, bool kBucket>
struct ItemGather {
    const int32_t *in_nbr, *in_trip;      // !kBucket
    const int32_t *bucket, *trip_user, *trip_pos;   // kBucket
    const int64_t *neg;
    const float2 *s_c;                    // kScalars: (s_t, c_t) for this role
    const float *Fh;
    int num_users;

    // own = this row's normalised final row (this lane's 4 columns)
    __device__ __forceinline__ void run(int begin, int end, int lane, const float4 &own, float4 &acc, float &sc) const {
        const int l16 = lane & 15;
        const float4 *Fh4 = reinterpret_cast<const float4 *>(Fh);
        constexpr float sigma = kBucket ? -1.f : 1.f;
        if constexpr (kScalars) {
            for_each_edge<ItemTrip, UNROLL>(
                begin, end, lane,
                [&](int e) {
                    ItemTrip it;
                    it.u = -1; it.w = 0.f; it.wc = 0.f; it.vu = f4zero();
                    if (e >= 0) {
                        int t;
                        if constexpr (kBucket) { t = __ldg(bucket + e); it.u = __ldg(trip_user + t); }
                        else { it.u = __ldg(in_nbr + e); t = __ldg(in_trip + e); }
                        const float2 q = __ldg(s_c + t);
                        it.w = sigma * q.x;
                        it.wc = sigma * q.x * q.y;
                    }
                    return it;
                },
                [&](int, ItemTrip &it) { if (it.u >= 0) it.vu = ldg4(Fh4 + (size_t)it.u * D4 + l16); },
                [&](int, ItemTrip &it) {
                    f4fma(acc, it.w, it.vu);
                    if (l16 == 0) sc += it.wc;
                });
        } else {
            float unused = 0.f;
            trip_batches<kBucket ? MODE_NEG : MODE_POS, false>(
                begin, end, lane, Fh4, own,
                [&](int e) {
                    TripEnt it{-1, -1, -1};
                    if (e >= 0) {
                        if constexpr (kBucket) {
                            it.t = __ldg(bucket + e);
                            it.i0 = __ldg(trip_user + it.t);
                            it.i1 = __ldg(trip_pos + it.t);
                        } else {
                            it.i0 = __ldg(in_nbr + e);
                            it.t = __ldg(in_trip + e);
                            it.i1 = (int)__ldg(neg + it.t) + num_users;
                        }
                    }
                    return it;
                },
                acc, sc, unused, nullptr, nullptr);
        }
    }
};

#ifndef LGCN_BPR_ITEM_MINWARPS
#define LGCN_BPR_ITEM_MINWARPS 64     /* resident warps per SM the item passes are compiled for (0 = no hint): 64 = a 32-register
                                         cap, BPR stage 1636 -> 1590 us (profiles/r2c_variants.txt);
                                         (single-GPU form that reads the per-triplet scalars; the recomputing form of the
                                         sharded step needs its 64 registers) */
#endif
template <bool kScalars>
struct BprItemOwnOp;
template <bool kScalars>
struct MinBlocks<BprItemOwnOp<kScalars>> { static constexpr int value = kScalars ? LGCN_BPR_ITEM_MINWARPS / WARPS_PER_CTA : 0; };

template <bool kScalars>
struct BprItemOwnOp {
    static constexpr bool kExtras = false;
    double *extra0, *extra1;
    ItemGather<kScalars, false> gather;
    const float *rnorm, *dis;
    float invP;
    float *G, *zG;
    Peers peers;

    __device__ __forceinline__ void accumulate(int row, int begin, int end, int lane, float4 &acc, float &sc,
                                               float &, float &) const {
        float4 own = f4zero();
        if constexpr (!kScalars) own = ldg4(reinterpret_cast<const float4 *>(gather.Fh) + (size_t)row * D4 + (lane & 15));
        gather.run(begin, end, lane, own, acc, sc);
    }
    __device__ __forceinline__ void epilogue(int row, int lane, const float4 &A, float B, float &, float &) const {
        const int l16 = lane & 15;
        const float4 ph = ldg4(reinterpret_cast<const float4 *>(gather.Fh) + (size_t)row * D4 + l16);
        float4 g = A;
        f4fma(g, -B, ph);
        g = f4scale(__ldg(rnorm + row) * invP, g);
        if (lane < 16) {
            float4 *dst = reinterpret_cast<float4 *>(G) + (size_t)row * D4 + l16;
            float4 cur = *dst;                             // the negative-role part (neg pass, earlier launch)
            f4add(cur, g);
            *dst = cur;
            push4(reinterpret_cast<float4 *>(zG) + (size_t)row * D4 + l16, f4scale(__ldg(dis + row), cur), peers);
        }
    }
};

// One warp per owned item row [ib,ie) (dynamic): the negative-role gradient from the row's bucket.  WRITES G[row]
// (zero for an empty bucket), so G needs no clearing; rows without in-edges keep zG = 0 (dis = 0).
template <bool kScalars>
__global__ void __launch_bounds__(CTA_THREADS, kScalars ? LGCN_BPR_ITEM_MINWARPS / WARPS_PER_CTA : 0)
bpr_neg_kernel(ItemGather<kScalars, true> gather, const float *__restrict__ rnorm,
               const int32_t *__restrict__ bucket_ptr, int ib, int ie, float invP, float *__restrict__ G,
               int *__restrict__ sched) {
    const int lane = threadIdx.x & 31, l16 = lane & 15;
    const int total = ie - ib;
    int r = 0;
    if (lane == 0) r = atomicAdd(sched, 1);
    r = __shfl_sync(FULL, r, 0);
    while (r < total) {
        int nxt = 0;
        if (lane == 0) nxt = atomicAdd(sched, 1);
        const int row = gather.num_users + ib + r;
        const int begin = __ldg(bucket_ptr + r), end = __ldg(bucket_ptr + r + 1);
        float4 g = f4zero();
        if (end > begin) {
            const float4 own = ldg4(reinterpret_cast<const float4 *>(gather.Fh) + (size_t)row * D4 + l16);
            float4 acc = f4zero();
            float sc = 0.f;
            gather.run(begin, end, lane, own, acc, sc);
            f4add(acc, f4shfl_xor16(acc));
            sc = warp_sum(sc);
            g = acc;
            f4fma(g, -sc, own);
            g = f4scale(__ldg(rnorm + row) * invP, g);
        }
        if (lane < 16) reinterpret_cast<float4 *>(G)[(size_t)row * D4 + l16] = g;
        r = __shfl_sync(FULL, nxt, 0);
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }
}

// ---------------------------------------------------------------------------------------------------
// buckets of the step's negatives, by item, for items [ib,ie)
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
neg_hist_kernel(const int64_t *__restrict__ neg, int64_t P, int ib, int ie, int32_t *__restrict__ neg_count,
                long long *bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < P; t += stride) {
        const int64_t i = __ldg(neg + t);
        if (i >= ib && i < ie) atomicAdd(neg_count + i, 1);
    }
    (void)bad;
}

// bucket_ptr[0..n] = exclusive scan of neg_count[ib..ie), cursor = copy; one CTA, 4 elements per thread and round
__global__ void __launch_bounds__(1024)
bucket_scan_kernel(const int32_t *__restrict__ neg_count, int ib, int n, int32_t *__restrict__ bucket_ptr,
                   int32_t *__restrict__ cursor) {
    __shared__ int warp_tot[32];
    __shared__ int carry_sh;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    int nv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) nv[q] = (int)threadIdx.x * 4 + q < n ? neg_count[ib + threadIdx.x * 4 + q] : 0;
    for (int base = 0; base < n; base += 4096) {
        const int i0 = base + threadIdx.x * 4;
        int v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = nv[q];
#pragma unroll
        for (int q = 0; q < 4; ++q) nv[q] = i0 + 4096 + q < n ? neg_count[ib + i0 + 4096 + q] : 0;   // next round's loads in flight
        const int mine = v[0] + v[1] + v[2] + v[3];
        int x = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(FULL, w, o);
                if (lane >= o) w += y;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int carry = carry_sh;
        int excl = carry + (wid ? warp_tot[wid - 1] : 0) + x - mine;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (i0 + q < n) { bucket_ptr[i0 + q] = excl; cursor[i0 + q] = excl; }
            excl += v[q];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_sh = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) bucket_ptr[n] = carry_sh;
}

// (several independent atomics in flight per thread do not help: 4 / 8-way unrolling moved the BPR stage by < 1 %,
// profiles/r2k_variants.txt -- the two kernels are bound by the L2 atomic rate, not by latency)
__global__ void __launch_bounds__(256)
bucket_fill_kernel(const int64_t *__restrict__ neg, int64_t P, int ib, int ie, int32_t *__restrict__ cursor,
                   int32_t *__restrict__ bucket, int64_t cap) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < P; t += stride) {
        const int64_t i = __ldg(neg + t);
        if (i >= ib && i < ie) {
            const int pos = atomicAdd(cursor + (i - ib), 1);
            if (pos < cap) bucket[pos] = (int)t;
        }
    }
}

// ascending t inside every bucket, one warp per bucket: a bitonic network over 256 keys held in REGISTERS (8 per lane,
// key e = r*32 + lane: partners at distance < 32 are reached with one shuffle, the others are registers of the same
// lane) -- uniform negatives give buckets of ~P/I entries (190 at ML-25M).  Buckets of 257..SORT_CAP entries go through
// shared memory; larger ones (never the case for uniform negatives) stay in arrival order -- their sum is then
// correct but not bit-stable.
constexpr int SORT_CAP = 1024;

__device__ __forceinline__ void sort256_regs(int (&v)[8], int lane) {
#pragma unroll
    for (int k = 2; k <= 256; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if ((r & jr) == 0) {
                        const bool up = (((r << 5) | lane) & k) == 0;
                        const int a = v[r], c = v[r | jr];
                        const bool sw = (a > c) == up;
                        v[r] = sw ? c : a;
                        v[r | jr] = sw ? a : c;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int other = __shfl_xor_sync(FULL, v[r], j);
                    const bool up = (((r << 5) | lane) & k) == 0;
                    const bool lower = (lane & j) == 0;
                    v[r] = (lower == up) ? min(v[r], other) : max(v[r], other);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128)
bucket_sort_kernel(const int32_t *__restrict__ bucket_ptr, int n, int32_t *__restrict__ bucket) {
    __shared__ int buf[4][SORT_CAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int *s = buf[wid];
    for (int r = blockIdx.x * 4 + wid; r < n; r += gridDim.x * 4) {
        const int b = bucket_ptr[r], len = bucket_ptr[r + 1] - b;
        if (len < 2 || len > SORT_CAP) continue;
        if (len <= 256) {
            int v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = (q << 5) + lane < len ? bucket[b + (q << 5) + lane] : INT_MAX;
            sort256_regs(v, lane);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if ((q << 5) + lane < len) bucket[b + (q << 5) + lane] = v[q];
            continue;
        }
        int m = 512;
        while (m < len) m <<= 1;
        for (int i = lane; i < m; i += 32) s[i] = i < len ? bucket[b + i] : INT_MAX;
        __syncwarp();
        for (int k = 2; k <= m; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = lane; i < m; i += 32) {
                    const int p = i ^ j;
                    if (p > i) {
                        const int a = s[i], c = s[p];
                        const bool up = (i & k) == 0;
                        if ((a > c) == up) { s[i] = c; s[p] = a; }
                    }
                }
                __syncwarp();
            }
        for (int i = lane; i < len; i += 32) bucket[b + i] = s[i];
        __syncwarp();
    }
}

// trip_user[t] / trip_pos[t] for the triplets of the user rows covered by out-tasks [tb,te): written into every rank's
// copy (the neg pass of ANY rank may meet the triplet).
__global__ void __launch_bounds__(CTA_THREADS)
triplet_index_kernel(const lgcn_task *__restrict__ tasks, int tb, int te, const int32_t *__restrict__ out_nbr,
                     const int32_t *__restrict__ out_trip, int32_t *__restrict__ trip_user,
                     int32_t *__restrict__ trip_pos, Peers peers) {
    const int lane = threadIdx.x & 31;
    for (int w = tb + blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5); w < te; w += gridDim.x * WARPS_PER_CTA) {
        const lgcn_task ta = tasks[w];
        for (int e = ta.begin + lane; e < ta.end; e += 32) {
            const int t = __ldg(out_trip + e);
            if (t < 0) continue;
            push1(reinterpret_cast<float *>(trip_user) + t, __int_as_float(ta.row), peers);
            push1(reinterpret_cast<float *>(trip_pos) + t, __int_as_float(__ldg(out_nbr + e)), peers);
        }
    }
}

__global__ void remap_triplets_kernel(int32_t *__restrict__ trip, int64_t E, const int32_t *__restrict__ map, int64_t P) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int t = trip[e];
    if (t >= 0 && t < P) trip[e] = map[t];
}

static inline int grid_for(int64_t n, int threads, int cap) {
    int64_t want = (n + threads - 1) / threads;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace lgcn

extern "C" int lgcn_triplet_index(const lgcn_graph *g, int user_task_begin, int user_task_end, int32_t *trip_user,
                                  int32_t *trip_pos, const lgcn_peers *peers, void *stream) {
    using namespace lgcn;
    LGCN_REQUIRE(g && trip_user && trip_pos && user_task_begin >= 0 && user_task_end <= g->n_out_tasks &&
                 user_task_begin <= user_task_end, LGCN_E_INVALID, "triplet_index: bad argument");
    const int n = user_task_end - user_task_begin;
    if (n == 0) return LGCN_OK;
    triplet_index_kernel<<<grid_for(n, WARPS_PER_CTA, 148 * 8), CTA_THREADS, 0, (cudaStream_t)stream>>>(
        g->out_tasks, user_task_begin, user_task_end, g->out_nbr, g->out_trip, trip_user, trip_pos, make_peers(peers));
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

extern "C" int lgcn_graph_remap_triplets(lgcn_graph *g, const int32_t *trip_global, void *stream) {
    using namespace lgcn;
    LGCN_REQUIRE(g && trip_global, LGCN_E_INVALID, "remap_triplets: null argument");
    const int64_t E = g->num_edges, P = g->num_triplets;
    if (E == 0 || P == 0) return LGCN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    remap_triplets_kernel<<<cdiv(E, 256), 256, 0, st>>>(const_cast<int32_t *>(g->in_trip), E, trip_global, P);
    LGCN_LAUNCH_CHECK();
    remap_triplets_kernel<<<cdiv(E, 256), 256, 0, st>>>(const_cast<int32_t *>(g->out_trip), E, trip_global, P);
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

extern "C" int lgcn_bpr_buckets(const int64_t *neg, int64_t num_triplets, int64_t item_begin, int64_t item_end,
                                int32_t *neg_count, const lgcn_bpr_owner_ws *ws, void *stream) {
    using namespace lgcn;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(neg && neg_count && ws && ws->bucket_ptr && ws->bucket_cursor && ws->bucket, LGCN_E_INVALID,
                 "bpr_buckets: null argument");
    LGCN_REQUIRE(num_triplets > 0 && num_triplets < ((int64_t)1 << 31), LGCN_E_INVALID, "bpr_buckets: %lld triplets",
                 (long long)num_triplets);
    LGCN_REQUIRE(item_begin >= 0 && item_begin <= item_end && item_end < ((int64_t)1 << 31), LGCN_E_INVALID,
                 "bpr_buckets: item range [%lld,%lld)", (long long)item_begin, (long long)item_end);
    LGCN_REQUIRE(ws->bucket_cap >= num_triplets, LGCN_E_WORKSPACE, "bpr_buckets: bucket array holds %lld < %lld entries",
                 (long long)ws->bucket_cap, (long long)num_triplets);
    const int ib = (int)item_begin, ie = (int)item_end, ni = ie - ib;
    if (ni <= 0) return LGCN_OK;
    // buckets of this step's negatives for the owned items (also the histogram the regulariser needs)
    LGCN_CUDA(cudaMemsetAsync(neg_count + ib, 0, sizeof(int32_t) * (size_t)ni, st));
    const int grid = grid_for(num_triplets, 256, 148 * 8);
    neg_hist_kernel<<<grid, 256, 0, st>>>(neg, num_triplets, ib, ie, neg_count, nullptr);
    LGCN_LAUNCH_CHECK();
    bucket_scan_kernel<<<1, 1024, 0, st>>>(neg_count, ib, ni, ws->bucket_ptr, ws->bucket_cursor);
    LGCN_LAUNCH_CHECK();
    bucket_fill_kernel<<<grid, 256, 0, st>>>(neg, num_triplets, ib, ie, ws->bucket_cursor, ws->bucket, ws->bucket_cap);
    LGCN_LAUNCH_CHECK();
    bucket_sort_kernel<<<grid_for(ni, 4, 148 * 16), 128, 0, st>>>(ws->bucket_ptr, ni, ws->bucket);
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

static int bpr_owner_impl(bool build_buckets, const lgcn_graph *g, const float *final_hat, const float *rnorm,
                          const int64_t *neg, int64_t num_triplets, float *G, float *zG, int32_t *neg_count, double *accum,
                          const lgcn_bpr_owner_ws *ws, int user_task_begin, int user_task_end, int item_task_begin,
                          int item_task_end, int64_t item_begin, int64_t item_end, const lgcn_peers *peers, void *stream) {

    using namespace lgcn;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(g && final_hat && rnorm && neg && G && zG && neg_count && accum && ws, LGCN_E_INVALID,
                 "bpr_owner: null argument");
    LGCN_REQUIRE(ws->trip_user && ws->trip_pos && ws->bucket_ptr && ws->bucket_cursor && ws->bucket && ws->sched,
                 LGCN_E_INVALID, "bpr_owner: workspace arrays missing");
    const int num_items = g->num_nodes - g->num_users;
    LGCN_REQUIRE(num_triplets > 0 && num_triplets < ((int64_t)1 << 31), LGCN_E_INVALID, "bpr_owner: %lld triplets",
                 (long long)num_triplets);
    LGCN_REQUIRE(item_begin >= 0 && item_end <= num_items && item_begin <= item_end, LGCN_E_INVALID,
                 "bpr_owner: item range [%lld,%lld) outside [0,%d)", (long long)item_begin, (long long)item_end, num_items);
    LGCN_REQUIRE(user_task_begin >= 0 && user_task_end <= g->n_out_tasks && user_task_begin <= user_task_end &&
                 item_task_begin >= 0 && item_task_end <= g->n_in_tasks && item_task_begin <= item_task_end,
                 LGCN_E_INVALID, "bpr_owner: task range outside the lists");
    LGCN_REQUIRE(ws->bucket_cap >= num_triplets, LGCN_E_WORKSPACE, "bpr_owner: bucket array holds %lld < %lld entries",
                 (long long)ws->bucket_cap, (long long)num_triplets);
    const int ib = (int)item_begin, ie = (int)item_end, ni = ie - ib;
    const float invP = 1.0f / (float)num_triplets;
    const Peers P = make_peers(peers);
    const bool scalars = ws->scalars != nullptr;
    float2 *s_cp = reinterpret_cast<float2 *>(ws->scalars);
    float2 *s_cn = s_cp ? s_cp + num_triplets : nullptr;

    if (build_buckets && ni > 0) {
        const int rc = lgcn_bpr_buckets(neg, num_triplets, item_begin, item_end, neg_count, ws, stream);
        if (rc != LGCN_OK) return rc;
    }
    // user rows
    auto users = [&](auto op) {
        return launch_rowtasks(op, g->out_tasks, user_task_begin, user_task_end, g->partials, g->slot_counters, g->sched, st);
    };
    if (scalars)
        LGCN_CUDA(users(BprUserOwnOp<true>{accum, nullptr, g->out_nbr, g->out_trip, neg, final_hat, rnorm, g->dis,
                                           g->num_users, invP, G, zG, s_cp, s_cn, P}));
    else
        LGCN_CUDA(users(BprUserOwnOp<false>{accum, nullptr, g->out_nbr, g->out_trip, neg, final_hat, rnorm, g->dis,
                                            g->num_users, invP, G, zG, nullptr, nullptr, P}));
    // item rows: negative role (writes G), then positive role (adds, stores zG everywhere)
    if (ni > 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int grid = grid_for(ni, WARPS_PER_CTA, sms * 6);
        if (scalars)
            bpr_neg_kernel<true><<<grid, CTA_THREADS, 0, st>>>(
                ItemGather<true, true>{nullptr, nullptr, ws->bucket, ws->trip_user, ws->trip_pos, neg, s_cn, final_hat,
                                       g->num_users}, rnorm, ws->bucket_ptr, ib, ie, invP, G, ws->sched);
        else
            bpr_neg_kernel<false><<<grid, CTA_THREADS, 0, st>>>(
                ItemGather<false, true>{nullptr, nullptr, ws->bucket, ws->trip_user, ws->trip_pos, neg, nullptr,
                                        final_hat, g->num_users}, rnorm, ws->bucket_ptr, ib, ie, invP, G, ws->sched);
        LGCN_LAUNCH_CHECK();
    }
    auto items = [&](auto op) {
        return launch_rowtasks(op, g->in_tasks, item_task_begin, item_task_end, g->partials, g->slot_counters, g->sched, st);
    };
    if (scalars)
        LGCN_CUDA(items(BprItemOwnOp<true>{nullptr, nullptr,
                                           ItemGather<true, false>{g->in_nbr, g->in_trip, nullptr, nullptr, nullptr, neg, s_cp,
                                                                   final_hat, g->num_users},
                                           rnorm, g->dis, invP, G, zG, P}));
    else
        LGCN_CUDA(items(BprItemOwnOp<false>{nullptr, nullptr,
                                            ItemGather<false, false>{g->in_nbr, g->in_trip, nullptr, nullptr, nullptr, neg,
                                                                     nullptr, final_hat, g->num_users},
                                            rnorm, g->dis, invP, G, zG, P}));
    return LGCN_OK;
}

extern "C" int lgcn_bpr_owner(const lgcn_graph *g, const float *final_hat, const float *rnorm, const int64_t *neg,
                              int64_t num_triplets, float *G, float *zG, int32_t *neg_count, double *accum,
                              const lgcn_bpr_owner_ws *ws, int user_task_begin, int user_task_end,
                              int item_task_begin, int item_task_end, int64_t item_begin, int64_t item_end,
                              const lgcn_peers *peers, void *stream) {
    return bpr_owner_impl(true, g, final_hat, rnorm, neg, num_triplets, G, zG, neg_count, accum, ws, user_task_begin,
                          user_task_end, item_task_begin, item_task_end, item_begin, item_end, peers, stream);
}

extern "C" int lgcn_bpr_owner_passes(const lgcn_graph *g, const float *final_hat, const float *rnorm, const int64_t *neg,
                                     int64_t num_triplets, float *G, float *zG, int32_t *neg_count, double *accum,
                                     const lgcn_bpr_owner_ws *ws, int user_task_begin, int user_task_end,
                                     int item_task_begin, int item_task_end, int64_t item_begin, int64_t item_end,
                                     const lgcn_peers *peers, void *stream) {
    return bpr_owner_impl(false, g, final_hat, rnorm, neg, num_triplets, G, zG, neg_count, accum, ws, user_task_begin,
                          user_task_end, item_task_begin, item_task_end, item_begin, item_end, peers, stream);
}
