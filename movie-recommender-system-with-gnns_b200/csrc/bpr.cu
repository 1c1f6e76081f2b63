// K3: BPR-cosine loss, forward value and gradient w.r.t. the FINAL embeddings, fused.
//
// Replaces compute_embeddings' six [P,64] row gathers (/root/reference/utils/train_test.py:
// 128-132), bpr_loss/normalize_embedding (:18-64) and their autograd (index_put scatter-adds).
//
//   triplet t = (u, p, n): the t-th edge with source < U in edge order gives (u, p)
//   (utils/helpers.py:98-99); n = neg[t] (utils/helpers.py:79-80).
//   cos+ = <u^,p^>, cos- = <u^,n^> with x^ = x/||x||_2 (no epsilon, :64)
//   loss = -mean_t softplus(10 (cos+ - cos-)) / 10 + coeff * mean_{P x 64}(u0^2 + p0^2 + n0^2)
//
// Pass A (one warp task per user row of the CSR by source): cos+/cos-, softplus sum, the
//   user-row gradient (owner computes, no atomics), the per-triplet scalars (s_t, cos+_t) and the
//   negative-item gradient (the only scattered target: uniform-random rows, vector red.add).
// Pass B (one warp task per item row of the CSR by target): positive-item gradient, owner
//   computes, reading (s_t, cos+_t).  Popular items (10^4-10^5 in-edges) therefore never
//   serialise on atomics.
// The regulariser only needs per-row role COUNTS (out-degree, in-degree, negative histogram), so
// its value and gradient are folded into the last backward layer's epilogue (propagate.cu).
#include "rowtask.cuh"

namespace lgcn {

constexpr int BPR_UNROLL = 4;

struct TripItemA {
    int dst, t, ng;
    float rp, rn;
    float4 vp, vn;
    __device__ __forceinline__ TripItemA shfl(int src) const {
        TripItemA r;
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

// kSparse: F / rnorm only hold the ACTIVE rows; a sampled negative without any incident edge has
// final = e0 / (K+1)^2 (every propagated layer is zero), formed on the fly.
template <bool kGrad, bool kSparse = false>
struct BprUserOp {
    static constexpr bool kExtras = true;
    double *extra0, *extra1;            // extra0: sum_t softplus(10 (cos+ - cos-))
    const int32_t *out_nbr, *out_trip;
    const int64_t *neg;
    const float *F, *rnorm;
    int num_users;
    float invP;
    float *G;
    int32_t *neg_count;
    float *scratch;
    const uint8_t *active;
    Table e0;
    float c0;

    __device__ __forceinline__ void accumulate(int row, int begin, int end, int lane, float4 &acc,
                                               float &sc, float &ex0, float &) const {
        const int l16 = lane & 15;
        const float4 *F4 = reinterpret_cast<const float4 *>(F);
        const float ru = __ldg(rnorm + row);
        const float4 fu = f4scale(ru, ldg4(F4 + (size_t)row * D4 + l16));     // u^
        float loss = 0.f;
        for_each_edge<TripItemA, BPR_UNROLL>(
            begin, end, lane,
            [&](int e) {
                TripItemA it;
                it.dst = -1; it.t = 0; it.ng = 0; it.rp = 0.f; it.rn = 0.f;
                it.vp = f4zero(); it.vn = f4zero();
                if (e >= 0) {
                    it.dst = __ldg(out_nbr + e);
                    it.t = __ldg(out_trip + e);
                    it.ng = (int)__ldg(neg + it.t) + num_users;
                    it.rp = __ldg(rnorm + it.dst);
                    if (kSparse && !__ldg(active + it.ng)) it.rn = -1.f;      // formed in `apply`
                    else it.rn = __ldg(rnorm + it.ng);
                }
                return it;
            },
            [&](int, TripItemA &it) {
                if (it.dst >= 0) {
                    it.vp = ldg4(F4 + (size_t)it.dst * D4 + l16);
                    if (kSparse && it.rn < 0.f) it.vn = f4scale(c0, ldg4(e0.row4(it.ng) + l16));
                    else it.vn = ldg4(F4 + (size_t)it.ng * D4 + l16);
                }
            },
            [&](int, TripItemA &it) {
                const bool valid = it.dst >= 0;
                if constexpr (kSparse) {
                    const float n2 = half_sum(f4dot(it.vn, it.vn));
                    if (it.rn < 0.f) it.rn = 1.0f / sqrtf(n2);
                }
                const float cp = half_sum(f4dot(fu, it.vp)) * it.rp;
                const float cn = half_sum(f4dot(fu, it.vn)) * it.rn;
                const float x = 10.f * (cp - cn);
                const float sp = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));      // softplus
                if (valid && l16 == 0) loss += sp;
                if constexpr (kGrad) {
                    const float s = -1.f / (1.f + expf(-x));                    // dL_t/dcos+ (x P)
                    f4fma(acc, s * it.rp, it.vp);                               // s (p^ - n^)
                    f4fma(acc, -s * it.rn, it.vn);
                    if (valid) {
                        if (l16 == 0) {
                            sc += s * (cp - cn);
                            reinterpret_cast<float2 *>(scratch)[it.t] = make_float2(s, cp);
                            atomicAdd(neg_count + (it.ng - num_users), 1);
                        }
                        // d/dn: -s (u^ - cos- n^)/||n|| / P   -> scattered row, vector red
                        const float k = -s * it.rn * invP;
                        float4 g = fu;
                        f4fma(g, -cn * it.rn, it.vn);
                        g = f4scale(k, g);
                        atomicAdd(reinterpret_cast<float4 *>(G) + (size_t)it.ng * D4 + l16, g);
                    }
                } else {
                    if (valid && l16 == 0) atomicAdd(neg_count + (it.ng - num_users), 1);
                }
            });
        ex0 += warp_sum(loss);
    }
    __device__ __forceinline__ void epilogue(int row, int lane, const float4 &A, float B, float &, float &) const {
        if constexpr (kGrad) {
            const int l16 = lane & 15;
            const float ru = __ldg(rnorm + row);
            const float4 fu = f4scale(ru, ldg4(reinterpret_cast<const float4 *>(F) + (size_t)row * D4 + l16));
            float4 g = A;                                  // (A - B u^) / ||u|| / P
            f4fma(g, -B, fu);
            g = f4scale(ru * invP, g);
            if (lane < 16) reinterpret_cast<float4 *>(G)[(size_t)row * D4 + l16] = g;
        }
    }
};

struct TripItemB {
    int u;
    float s, scp, ru;
    float4 vu;
    __device__ __forceinline__ TripItemB shfl(int src) const {
        TripItemB r;
        r.u = __shfl_sync(FULL, u, src);
        r.s = __shfl_sync(FULL, s, src);
        r.scp = __shfl_sync(FULL, scp, src);
        r.ru = __shfl_sync(FULL, ru, src);
        r.vu = f4zero();
        return r;
    }
};

struct BprItemOp {
    static constexpr bool kExtras = false;
    double *extra0, *extra1;
    const int32_t *in_nbr, *in_trip;
    const float *F, *rnorm, *scratch;
    float invP;
    float *G;
    int ub, ue;                 // only triplets of users in [ub,ue) are accumulated (sharded path)
    int sorted;                 // the row's source ids ascend: the slice of users in [ub,ue) is found by bisection

    // first position in [lo,hi) whose source id is >= key: 32-ary search, one probe per lane and round
    __device__ __forceinline__ int lower_bound32(int lo, int hi, int key, int lane) const {
        while (hi - lo > 32) {
            const int step = (hi - lo + 31) / 32;                       // probes at lo + step*(lane+1) - 1
            const int pos = min(lo + step * (lane + 1) - 1, hi - 1);
            const unsigned less = __ballot_sync(FULL, __ldg(in_nbr + pos) < key);
            const int k = __popc(less);                                  // probes 0..k-1 are < key (ascending row)
            const int nlo = k == 0 ? lo : min(lo + step * k, hi);
            const int nhi = k == 32 ? hi : min(lo + step * (k + 1) - 1, hi - 1) + 1;
            lo = nlo; hi = nhi;
        }
        const int pos = lo + lane;
        const unsigned less = __ballot_sync(FULL, pos < hi && __ldg(in_nbr + pos) < key);
        return lo + __popc(less);
    }

    __device__ __forceinline__ void accumulate(int, int begin, int end, int lane, float4 &acc, float &sc,
                                               float &, float &) const {
        const int l16 = lane & 15;
        const float4 *F4 = reinterpret_cast<const float4 *>(F);
        if (sorted && end - begin > 8) {                                 // this rank's users form one contiguous slice
            const int b2 = lower_bound32(begin, end, ub, lane);
            end = lower_bound32(b2, end, ue, lane);
            begin = b2;
        }
        for_each_edge<TripItemB, UNROLL>(
            begin, end, lane,
            [&](int e) {
                TripItemB it;
                it.u = -1; it.s = 0.f; it.scp = 0.f; it.ru = 0.f; it.vu = f4zero();
                if (e >= 0) it.u = __ldg(in_nbr + e);
                if (it.u < ub || it.u >= ue) it.u = -1;
                if (it.u >= 0) {
                    const float2 sc2 = __ldg(reinterpret_cast<const float2 *>(scratch) + __ldg(in_trip + e));
                    it.ru = __ldg(rnorm + it.u);
                    it.s = sc2.x * it.ru;                 // s_t / ||u||
                    it.scp = sc2.x * sc2.y;               // s_t cos+_t
                }
                return it;
            },
            [&](int, TripItemB &it) { if (it.u >= 0) it.vu = ldg4(F4 + (size_t)it.u * D4 + l16); },
            [&](int, TripItemB &it) {
                f4fma(acc, it.s, it.vu);                   // sum_t s_t u^_t
                if (l16 == 0) sc += it.scp;                // sum_t s_t cos+_t (0 for padding)
            });
    }
    __device__ __forceinline__ void epilogue(int row, int lane, const float4 &A, float B, float &, float &) const {
        const int l16 = lane & 15;
        const float rp = __ldg(rnorm + row);
        const float4 fp = f4scale(rp, ldg4(reinterpret_cast<const float4 *>(F) + (size_t)row * D4 + l16));
        float4 g = A;                                      // (A - B p^) / ||p|| / P
        f4fma(g, -B, fp);
        g = f4scale(rp * invP, g);
        if (lane < 16) {
            float4 *dst = reinterpret_cast<float4 *>(G) + (size_t)row * D4 + l16;
            float4 cur = *dst;                             // negative-sample contributions (pass A)
            f4add(cur, g);
            *dst = cur;
        }
    }
};

// Regulariser value for the loss-only path: sum_r cnt[r] * ||e0[r]||^2.
__global__ void __launch_bounds__(CTA_THREADS)
reg_value_kernel(Table e0, const int32_t *__restrict__ in_ptr, const int32_t *__restrict__ out_ptr,
                 const int32_t *__restrict__ neg_count, int n, int num_users, double *out) {
    const int lane = threadIdx.x & 31, l16 = lane & 15, wid = threadIdx.x >> 5;
    const int row = (blockIdx.x * WARPS_PER_CTA + wid) * 2 + (lane >> 4);
    float v = 0.f;
    if (row < n) {
        int cnt;
        if (row < num_users) cnt = out_ptr[row + 1] - out_ptr[row];
        else cnt = in_ptr[row + 1] - in_ptr[row] + neg_count[row - num_users];
        if (cnt) {
            const float4 e = ldg4(e0.row4(row) + l16);
            v = (float)cnt * f4dot(e, e);
        }
    }
    v = warp_sum(v);
    __shared__ float s[WARPS_PER_CTA];
    if (lane == 0) s[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int w = 0; w < WARPS_PER_CTA; ++w) a += s[w];
        if (a != 0.0) atomicAdd(out, a);
    }
}

__global__ void loss_finalize_kernel(const double *accum, int64_t P, float coeff, float *loss_out) {
    // -mean softplus / 10 + coeff * mean_{P x 64}(...)   (utils/train_test.py:38-51)
    const double p = (double)P;
    loss_out[0] = (float)(-accum[0] / (10.0 * p) + (double)coeff * accum[1] / (64.0 * p));
}

// Triplets of the users covered by out-tasks [utb,ute) / rows [urb,ure) only (everything: 0,
// n_out_user_tasks, 0, num_users).  G and neg_count are zeroed first, so per-rank results can be summed.
int bpr_impl(const lgcn_graph *g, const float *F, const float *rnorm, const int64_t *neg, float *G,
             int32_t *neg_count, float *scratch, double *accum, bool grad, int utb, int ute, int urb, int ure,
             cudaStream_t st) {
    LGCN_REQUIRE(g && F && rnorm && neg && neg_count && accum, LGCN_E_INVALID, "bpr: null argument");
    LGCN_REQUIRE(!grad || (G && scratch), LGCN_E_INVALID, "bpr: gradient buffers missing");
    const int num_items = g->num_nodes - g->num_users;
    LGCN_CUDA(cudaMemsetAsync(neg_count, 0, sizeof(int32_t) * (size_t)num_items, st));
    if (g->num_triplets == 0) {
        if (grad) LGCN_CUDA(cudaMemsetAsync(G, 0, sizeof(float) * (size_t)g->num_nodes * D, st));
        return LGCN_OK;
    }
    const float invP = 1.0f / (float)g->num_triplets;
    if (grad) {
        LGCN_CUDA(cudaMemsetAsync(G, 0, sizeof(float) * (size_t)g->num_nodes * D, st));
        BprUserOp<true> a{accum, nullptr, g->out_nbr, g->out_trip, neg, F, rnorm, g->num_users, invP, G,
                          neg_count, scratch, nullptr, Table{}, 0.f};
        LGCN_CUDA(launch_rowtasks(a, g->out_tasks, utb, ute, g->partials, g->slot_counters, g->sched, st));
        BprItemOp b{nullptr, nullptr, g->in_nbr, g->in_trip, F, rnorm, scratch, invP, G, urb, ure,
                    (g->in_src_sorted && (urb > 0 || ure < g->num_users)) ? 1 : 0};
        LGCN_CUDA(launch_rowtasks(b, g->in_tasks, g->n_in_user_tasks, g->n_in_tasks, g->partials,
                                  g->slot_counters, g->sched, st));
    } else {
        BprUserOp<false> a{accum, nullptr, g->out_nbr, g->out_trip, neg, F, rnorm, g->num_users, invP,
                           nullptr, neg_count, nullptr, nullptr, Table{}, 0.f};
        LGCN_CUDA(launch_rowtasks(a, g->out_tasks, utb, ute, g->partials, g->slot_counters, g->sched, st));
    }
    return LGCN_OK;
}

// Sparse-step BPR: G and neg_count are all-zero on entry (invariant kept by the sparse step), F / rnorm
// are only valid for active rows.
int bpr_sparse_impl(const lgcn_graph *g, const float *F, const float *rnorm, const int64_t *neg, float *G,
                    int32_t *neg_count, float *scratch, double *accum, const float *user_w, const float *item_w,
                    int K, cudaStream_t st) {
    const float invP = 1.0f / (float)g->num_triplets;
    const float c0 = 1.0f / (float)((K + 1) * (K + 1));
    BprUserOp<true, true> a{accum, nullptr, g->out_nbr, g->out_trip, neg, F, rnorm, g->num_users, invP, G,
                            neg_count, scratch, g->active, Table{user_w, item_w, g->num_users}, c0};
    LGCN_CUDA(launch_rowtasks(a, g->out_tasks, 0, g->n_out_user_tasks, g->partials, g->slot_counters, g->sched, st));
    BprItemOp b{nullptr, nullptr, g->in_nbr, g->in_trip, F, rnorm, scratch, invP, G, 0, g->num_users, 0};
    LGCN_CUDA(launch_rowtasks(b, g->in_tasks, g->n_in_user_tasks, g->n_in_tasks, g->partials, g->slot_counters, g->sched, st));
    return LGCN_OK;
}

int reg_value_impl(const lgcn_graph *g, const float *user_w, const float *item_w,
                   const int32_t *neg_count, double *out, cudaStream_t st) {
    const Table e0{user_w, item_w, g->num_users};
    reg_value_kernel<<<cdiv(g->num_nodes, 2 * WARPS_PER_CTA), CTA_THREADS, 0, st>>>(
        e0, g->in_ptr, g->out_ptr, neg_count, g->num_nodes, g->num_users, out);
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

int loss_finalize_impl(const double *accum, int64_t P, float coeff, float *loss_out, cudaStream_t st) {
    loss_finalize_kernel<<<1, 1, 0, st>>>(accum, P, coeff, loss_out);
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

}  // namespace lgcn

extern "C" int lgcn_bpr_fwd_bwd(const lgcn_graph *g, const float *final_emb, const float *rnorm,
                                const int64_t *neg, float *grad_final, int32_t *neg_count,
                                float *trip_scratch, double *accum, void *stream) {
    LGCN_REQUIRE(g, LGCN_E_INVALID, "bpr: null graph");
    return lgcn::bpr_impl(g, final_emb, rnorm, neg, grad_final, neg_count, trip_scratch, accum, true, 0,
                          g->n_out_user_tasks, 0, g->num_users, (cudaStream_t)stream);
}

extern "C" int lgcn_bpr_fwd_bwd_range(const lgcn_graph *g, const float *final_emb, const float *rnorm,
                                      const int64_t *neg, float *grad_final, int32_t *neg_count,
                                      float *trip_scratch, double *accum, int user_task_begin, int user_task_end,
                                      int64_t user_row_begin, int64_t user_row_end, void *stream) {
    LGCN_REQUIRE(g && user_task_begin >= 0 && user_task_end <= g->n_out_user_tasks && user_task_begin <= user_task_end &&
                 user_row_begin >= 0 && user_row_end <= g->num_users, LGCN_E_INVALID, "bpr_range: bad range");
    return lgcn::bpr_impl(g, final_emb, rnorm, neg, grad_final, neg_count, trip_scratch, accum, true, user_task_begin,
                          user_task_end, (int)user_row_begin, (int)user_row_end, (cudaStream_t)stream);
}
