// K1 / K2: K-layer symmetric-normalised propagation, forward and backward.
//
// Replaces `for conv in self.convs: emb = conv(emb, edge_index)` + stack/mean/split
// (/root/reference/models/light_gcn.py:29-38; LGConv = PyG gcn_norm + gather/mul/scatter_add)
// and its autograd (utils/train_test.py:94).
//
// Pre-scaled formulation.  With dis = in_degree^-1/2 (0 where the degree is 0) the layer is
//     x_{k+1}[c] = dis[c] * sum_{r->c} dis[r] * x_k[r].
// Instead of x_k the kernels store y_k = dis (.) x_k, so a layer is a pure row gather-sum
//     raw[c] = sum_{r->c} y_k[r];   x_{k+1}[c] = dis[c]*raw[c];   y_{k+1}[c] = raw[c]/deg[c]
// with no per-edge weight to load (layer 1 reads e0 and applies dis[r] per edge, the 4-byte dis
// gather rides along with the coalesced index load).  The layer mean is fused into the LAST
// layer's epilogue: final = (e0 + sqrt(deg)*(y_1+..+y_{K-1}) + dis*raw_K) / (K+1)^2, so the
// N x (K+1) x 64 stack of the reference is never materialised.
//
// Backward is the same gather-sum over the CSR by source, Horner-evaluated:
//     h_0 = G;  h_j = G + A^T h_{j-1};  grad_e0 = h_K / (K+1)^2  (+ BPR regulariser gradient),
// storing z_j = dis (.) h_j between layers.  Owner-computes rows => no float atomics, results
// are run-to-run bit-stable (the reference's CUDA scatter_add_ is not).
#include "rowtask.cuh"

namespace lgcn {

struct NbrItem {
    int nbr;
    float w;
    float4 v;
    __device__ __forceinline__ NbrItem shfl(int src_lane) const {
        NbrItem r;
        r.nbr = __shfl_sync(FULL, nbr, src_lane);
        r.w = __shfl_sync(FULL, w, src_lane);
        r.v = f4zero();
        return r;
    }
};

// Gather-sum of rows of a flat [N,64] table (layers >= 2, both directions).
struct GatherFlat {
    const int32_t *nbr;
    const float *x;
    __device__ __forceinline__ void run(int begin, int end, int lane, float4 &acc) const {
        const int l16 = lane & 15;
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        for_each_edge<NbrItem>(
            begin, end, lane,
            [&](int e) { NbrItem it; it.nbr = e >= 0 ? __ldg(nbr + e) : -1; it.w = 0.f; it.v = f4zero(); return it; },
            [&](int, NbrItem &it) { if (it.nbr >= 0) it.v = ldg4(x4 + (size_t)it.nbr * D4 + l16); },
            [&](int, NbrItem &it) { f4add(acc, it.v); });
    }
};

// Layer-1 gather: sum_r dis[r] * T[r] where T is either the (user,item) weight pair (forward)
// or a flat table (backward, T = G).
template <bool kTable>
struct GatherScaled {
    const int32_t *nbr;
    const float *dis;
    Table tab;
    const float *x;
    __device__ __forceinline__ void run(int begin, int end, int lane, float4 &acc) const {
        const int l16 = lane & 15;
        for_each_edge<NbrItem>(
            begin, end, lane,
            [&](int e) {
                NbrItem it;
                it.nbr = e >= 0 ? __ldg(nbr + e) : -1;
                it.w = e >= 0 ? __ldg(dis + it.nbr) : 0.f;
                it.v = f4zero();
                return it;
            },
            [&](int, NbrItem &it) {
                if (it.nbr >= 0) {
                    if constexpr (kTable) it.v = ldg4(tab.row4(it.nbr) + l16);
                    else it.v = ldg4(reinterpret_cast<const float4 *>(x) + (size_t)it.nbr * D4 + l16);
                }
            },
            [&](int, NbrItem &it) { f4fma(acc, it.w, it.v); });
    }
};

// ---------------------------------------------------------------------------------------
// forward ops
// ---------------------------------------------------------------------------------------

template <bool kFirst, bool kLast>
struct FwdOp {
    static constexpr bool kExtras = false;
    double *extra0, *extra1;
    const int32_t *ptr, *nbr;
    const float *dis;
    Table e0;
    const float *yin;        // y_{k-1} (unused when kFirst)
    float *yout;             // y_k     (unused when kLast)
    const float *ysum[3];    // y_1..y_{K-1} for the fused mean (kLast)
    int nsum;
    float c0;                // 1/(K+1)^2
    float *final_out;
    float *rnorm;
    Peers peers;             // where produced rows go (all ranks' copies in the sharded path)
    int norm_out;            // kLast: final_out receives final / ||final|| (every copy), rnorm 1/||final|| (LOCAL only)

    __device__ __forceinline__ void accumulate(int, int begin, int end, int lane, float4 &acc, float &, float &, float &) const {
        if constexpr (kFirst) GatherScaled<true>{nbr, dis, e0, nullptr}.run(begin, end, lane, acc);
        else GatherFlat{nbr, yin}.run(begin, end, lane, acc);
    }
    __device__ __forceinline__ void epilogue(int row, int lane, const float4 &raw, float, float &, float &) const {
        const int l16 = lane & 15;
        const int deg = __ldg(ptr + row + 1) - __ldg(ptr + row);
        if constexpr (!kLast) {
            const float inv = deg > 0 ? 1.0f / (float)deg : 0.f;
            if (lane < 16) push4(reinterpret_cast<float4 *>(yout) + (size_t)row * D4 + l16, f4scale(inv, raw), peers);
        } else {
            float4 s = f4zero();
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (i < nsum) f4add(s, ldg4(reinterpret_cast<const float4 *>(ysum[i]) + (size_t)row * D4 + l16));
            const float sq = sqrtf((float)deg);
            const float d = __ldg(dis + row);
            float4 f = ldg4(e0.row4(row) + l16);
            f4fma(f, sq, s);
            f4fma(f, d, raw);
            f = f4scale(c0, f);
            if (norm_out) {
                const float rn = 1.0f / sqrtf(half_sum(f4dot(f, f)));
                if (lane < 16) push4(reinterpret_cast<float4 *>(final_out) + (size_t)row * D4 + l16, f4scale(rn, f), peers);
                if (lane == 0) rnorm[row] = rn;
            } else {
                if (lane < 16) push4(reinterpret_cast<float4 *>(final_out) + (size_t)row * D4 + l16, f, peers);
                if (rnorm) {
                    const float n2 = half_sum(f4dot(f, f));
                    if (lane == 0) push1(rnorm + row, 1.0f / sqrtf(n2), peers);
                }
            }
        }
    }
};

// Rows without any incident edge: final = e0 / (K+1)^2 (every propagated layer is zero).
__global__ void __launch_bounds__(CTA_THREADS)
fwd_inactive_kernel(Table e0, const uint8_t *__restrict__ active, int row0, int n, float c0,
                    float *__restrict__ final_out, float *__restrict__ rnorm, Peers peers, int norm_out) {
    const int lane = threadIdx.x & 31, l16 = lane & 15;
    int row = row0 + (blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5)) * 2 + (lane >> 4);
    const bool ok = row < n && !active[row];
    float4 f = f4zero();
    if (ok) f = f4scale(c0, ldg4(e0.row4(row) + l16));
    const float rn = 1.0f / sqrtf(half_sum(f4dot(f, f)));
    if (ok) {
        push4(reinterpret_cast<float4 *>(final_out) + (size_t)row * D4 + l16, norm_out ? f4scale(rn, f) : f, peers);
        if (l16 == 0 && rnorm) {
            if (norm_out) rnorm[row] = rn;
            else push1(rnorm + row, rn, peers);
        }
    }
}

// ---------------------------------------------------------------------------------------
// backward ops
// ---------------------------------------------------------------------------------------

template <bool kFirst, bool kLast>
struct BwdOp {
    static constexpr bool kExtras = kLast;
    double *extra0, *extra1;   // extra0: sum cnt*||e0||^2 (reg loss), extra1: ||grad||^2
    const int32_t *nbr;        // out_nbr
    const int32_t *in_ptr, *out_ptr;
    const float *dis;
    const float *G;            // grad w.r.t. final
    const float *zin;
    float *zout;
    float c0;
    float *grad;
    Table e0;
    const int32_t *neg_count;
    float reg_coef;
    int num_users;
    Peers peers;

    __device__ __forceinline__ void accumulate(int, int begin, int end, int lane, float4 &acc, float &, float &, float &) const {
        if constexpr (kFirst) GatherScaled<false>{nbr, dis, Table{}, G}.run(begin, end, lane, acc);
        else GatherFlat{nbr, zin}.run(begin, end, lane, acc);
    }
    __device__ __forceinline__ void epilogue(int row, int lane, const float4 &S, float, float &ex0, float &ex1) const {
        const int l16 = lane & 15;
        const float d = __ldg(dis + row);
        float4 h = ldg4(reinterpret_cast<const float4 *>(G) + (size_t)row * D4 + l16);
        f4fma(h, d, S);
        if constexpr (!kLast) {
            if (lane < 16) push4(reinterpret_cast<float4 *>(zout) + (size_t)row * D4 + l16, f4scale(d, h), peers);
        } else {
            float4 g = f4scale(c0, h);
            if (reg_coef != 0.f) {
                // BPR regulariser: d/de0 of coef/(64P) * sum_t (|u0|^2+|p0|^2+|n0|^2); a row is
                // counted once per triplet role it plays (utils/train_test.py:38-40).
                int cnt;
                if (row < num_users) cnt = __ldg(out_ptr + row + 1) - __ldg(out_ptr + row);
                else cnt = __ldg(in_ptr + row + 1) - __ldg(in_ptr + row) + __ldg(neg_count + row - num_users);
                const float4 e = ldg4(e0.row4(row) + l16);
                f4fma(g, reg_coef * (float)cnt, e);
                const float e2 = half_sum(f4dot(e, e));
                ex0 += (float)cnt * e2;
            }
            if (lane < 16) reinterpret_cast<float4 *>(grad)[(size_t)row * D4 + l16] = g;
            ex1 += half_sum(f4dot(g, g));
        }
    }
};

// Rows without incident edges: grad = G/(K+1)^2 + reg (only sampled negatives reach them).
__global__ void __launch_bounds__(CTA_THREADS)
bwd_inactive_kernel(const float *__restrict__ G, const uint8_t *__restrict__ active, int row0, int n, float c0,
                    Table e0, const int32_t *__restrict__ neg_count, float reg_coef, int num_users,
                    float *__restrict__ grad, double *extra0, double *extra1) {
    const int lane = threadIdx.x & 31, l16 = lane & 15, wid = threadIdx.x >> 5;
    int row = row0 + (blockIdx.x * WARPS_PER_CTA + wid) * 2 + (lane >> 4);
    const bool ok = row < n && !active[row];
    float4 g = f4zero();
    float reg = 0.f;
    if (ok) {
        g = f4scale(c0, ldg4(reinterpret_cast<const float4 *>(G) + (size_t)row * D4 + l16));
        if (reg_coef != 0.f && row >= num_users) {
            const int cnt = __ldg(neg_count + row - num_users);
            if (cnt) {
                const float4 e = ldg4(e0.row4(row) + l16);
                f4fma(g, reg_coef * (float)cnt, e);
                reg = (float)cnt * f4dot(e, e);
            }
        }
        reinterpret_cast<float4 *>(grad)[(size_t)row * D4 + l16] = g;
    }
    float n2 = warp_sum(f4dot(g, g));
    reg = warp_sum(reg);
    __shared__ float s[WARPS_PER_CTA][2];
    if (lane == 0) { s[wid][0] = reg; s[wid][1] = n2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < WARPS_PER_CTA; ++w) { a += s[w][0]; b += s[w][1]; }
        if (extra0 && a != 0.0) atomicAdd(extra0, a);
        if (extra1 && b != 0.0) atomicAdd(extra1, b);
    }
}

// Resident warps per SM the pure gather-sum layer kernels are compiled for (a register cap + a scheduling hint for
// ptxas): 48 warps = 6 CTAs = at most 40 registers.  Measured on B200 at ML-25M shape, bit-identical results
// (profiles/r2b_variants.txt, r2c_variants.txt): no hint (40 registers) fwd 442 / bwd 470-500 us per layer;
// hint 64 warps (32 registers) 383 / 376 us; hint 48 warps 376 / 367 us.  (The layer-1 kernels with per-edge weights
// spill under a cap and keep ptxas' own choice.)
#ifndef LGCN_SPMM_MINWARPS
#define LGCN_SPMM_MINWARPS 48
#endif
template <bool kLast>
struct MinBlocks<FwdOp<false, kLast>> { static constexpr int value = LGCN_SPMM_MINWARPS / WARPS_PER_CTA; };
template <bool kLast>
struct MinBlocks<BwdOp<false, kLast>> { static constexpr int value = LGCN_SPMM_MINWARPS / WARPS_PER_CTA; };

// ---------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------

// y0[r] = dis[r] * e0[r] for rows [row_begin,row_end): lets layer 1 run as a pure gather-sum when the
// table is assembled from per-rank slabs (sharded path).
__global__ void __launch_bounds__(CTA_THREADS)
prescale_kernel(Table e0, const float *__restrict__ dis, int row0, int n, float *__restrict__ y0, Peers peers) {
    const int lane = threadIdx.x & 31, l16 = lane & 15;
    const int row = row0 + (blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5)) * 2 + (lane >> 4);
    if (row < n)
        push4(reinterpret_cast<float4 *>(y0) + (size_t)row * D4 + l16, f4scale(__ldg(dis + row), ldg4(e0.row4(row) + l16)), peers);
}

// One forward layer k of K over the tasks of `r`.  scaled_first: layer 1 reads e0 and applies dis[r]
// per edge (single-GPU path); otherwise yin is the pre-scaled table y_{k-1} for every layer.
int fwd_layer_impl(const lgcn_graph *g, const Table &e0, int k, int K, bool scaled_first, const float *yin,
                   float *yout, const float *y1, const float *y2, const float *y3, float *final_out,
                   float *rnorm, Range r, cudaStream_t st, const Peers &peers, int norm_out) {
    const float c0 = 1.0f / (float)((K + 1) * (K + 1));
    const bool first = k == 1 && scaled_first, last = k == K;
    if (last && r.inactive_rows && g->num_active < g->num_nodes && r.re > r.rb) {
        fwd_inactive_kernel<<<cdiv(r.re - r.rb, 2 * WARPS_PER_CTA), CTA_THREADS, 0, st>>>(
            e0, g->active, r.rb, r.re, c0, final_out, rnorm, peers, norm_out);
        LGCN_LAUNCH_CHECK();
    }
    auto go = [&](auto op) { return launch_rowtasks(op, g->in_tasks, r.tb, r.te, g->partials, g->slot_counters, g->sched, st); };
    if (first && last)
        LGCN_CUDA(go(FwdOp<true, true>{nullptr, nullptr, g->in_ptr, g->in_nbr, g->dis, e0, nullptr, nullptr,
                                       {nullptr, nullptr, nullptr}, 0, c0, final_out, rnorm, peers, norm_out}));
    else if (first)
        LGCN_CUDA(go(FwdOp<true, false>{nullptr, nullptr, g->in_ptr, g->in_nbr, g->dis, e0, nullptr, yout,
                                        {nullptr, nullptr, nullptr}, 0, c0, nullptr, nullptr, peers, 0}));
    else if (!last)
        LGCN_CUDA(go(FwdOp<false, false>{nullptr, nullptr, g->in_ptr, g->in_nbr, g->dis, e0, yin, yout,
                                         {nullptr, nullptr, nullptr}, 0, c0, nullptr, nullptr, peers, 0}));
    else
        LGCN_CUDA(go(FwdOp<false, true>{nullptr, nullptr, g->in_ptr, g->in_nbr, g->dis, e0, yin, nullptr,
                                        {y1, y2, y3}, K - 1, c0, final_out, rnorm, peers, norm_out}));
    return LGCN_OK;
}

int fwd_layer_impl(const lgcn_graph *g, const Table &e0, int k, int K, bool scaled_first, const float *yin,
                   float *yout, const float *y1, const float *y2, const float *y3, float *final_out,
                   float *rnorm, Range r, cudaStream_t st, const Peers &peers) {
    return fwd_layer_impl(g, e0, k, K, scaled_first, yin, yout, y1, y2, y3, final_out, rnorm, r, st, peers, 0);
}

// One backward (Horner) layer j of K over the tasks of `r`; layer 1 gathers dis (.) G itself unless the caller
// hands it the pre-scaled table z_0 = dis (.) G as `zin` (sharded path: the BPR epilogues store z_0 into
// every rank's copy, G itself stays local).
int bwd_layer_impl(const lgcn_graph *g, const float *G, int j, int K, const float *zin, float *zout,
                   const Table &e0, const int32_t *neg_count, float reg_coef, float *grad, double *accum,
                   Range r, cudaStream_t st, const Peers &peers) {
    const float c0 = 1.0f / (float)((K + 1) * (K + 1));
    const bool first = j == 1 && zin == nullptr, last = j == K;
    double *ex0 = accum ? accum + 1 : nullptr, *ex1 = accum ? accum + 2 : nullptr;
    if (last && r.inactive_rows && g->num_active < g->num_nodes && r.re > r.rb) {
        bwd_inactive_kernel<<<cdiv(r.re - r.rb, 2 * WARPS_PER_CTA), CTA_THREADS, 0, st>>>(
            G, g->active, r.rb, r.re, c0, e0, neg_count, reg_coef, g->num_users, grad, ex0, ex1);
        LGCN_LAUNCH_CHECK();
    }
    auto go = [&](auto op) { return launch_rowtasks(op, g->out_tasks, r.tb, r.te, g->partials, g->slot_counters, g->sched, st); };
    if (first && last)
        LGCN_CUDA(go(BwdOp<true, true>{ex0, ex1, g->out_nbr, g->in_ptr, g->out_ptr, g->dis, G, zin, zout, c0, grad, e0,
                                       neg_count, reg_coef, g->num_users, peers}));
    else if (first)
        LGCN_CUDA(go(BwdOp<true, false>{nullptr, nullptr, g->out_nbr, g->in_ptr, g->out_ptr, g->dis, G, zin, zout, c0,
                                        grad, e0, neg_count, reg_coef, g->num_users, peers}));
    else if (!last)
        LGCN_CUDA(go(BwdOp<false, false>{nullptr, nullptr, g->out_nbr, g->in_ptr, g->out_ptr, g->dis, G, zin, zout, c0,
                                         grad, e0, neg_count, reg_coef, g->num_users, peers}));
    else
        LGCN_CUDA(go(BwdOp<false, true>{ex0, ex1, g->out_nbr, g->in_ptr, g->out_ptr, g->dis, G, zin, zout, c0, grad, e0,
                                        neg_count, reg_coef, g->num_users, peers}));
    return LGCN_OK;
}

int propagate_fwd_impl(const lgcn_graph *g, const float *user_w, const float *item_w, int K,
                       float *final_out, float *rnorm, float *work, size_t work_bytes, cudaStream_t st) {
    LGCN_REQUIRE(g && user_w && item_w && final_out, LGCN_E_INVALID, "propagate_fwd: null argument");
    LGCN_REQUIRE(K >= 1 && K <= 4, LGCN_E_INVALID, "propagate_fwd: num_layers %d outside [1,4]", K);
    const size_t n = (size_t)g->num_nodes;
    LGCN_REQUIRE(work_bytes >= (size_t)(K - 1) * n * D * sizeof(float) && (K == 1 || work),
                 LGCN_E_WORKSPACE, "propagate_fwd: workspace %zu < %zu bytes", work_bytes,
                 (size_t)(K - 1) * n * D * sizeof(float));
    const Table e0{user_w, item_w, g->num_users};
    float *y[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int k = 1; k < K; ++k) y[k] = work + (size_t)(k - 1) * n * D;
    const Range all{0, g->n_in_tasks, 0, g->num_nodes, true};
    for (int k = 1; k <= K; ++k) {
        int rc = fwd_layer_impl(g, e0, k, K, true, y[k - 1], k < K ? y[k] : nullptr, y[1], y[2], y[3], final_out,
                                rnorm, all, st, local_only());
        if (rc) return rc;
    }
    return LGCN_OK;
}

int propagate_bwd_impl(const lgcn_graph *g, const float *G, int K, const float *user_w,
                       const float *item_w, const int32_t *neg_count, float reg_coef, float *grad,
                       double *accum, float *work, size_t work_bytes, cudaStream_t st) {
    LGCN_REQUIRE(g && G && grad, LGCN_E_INVALID, "propagate_bwd: null argument");
    LGCN_REQUIRE(K >= 1 && K <= 4, LGCN_E_INVALID, "propagate_bwd: num_layers %d outside [1,4]", K);
    LGCN_REQUIRE(reg_coef == 0.f || (user_w && item_w && neg_count && accum), LGCN_E_INVALID,
                 "propagate_bwd: regulariser needs weights, neg_count and accum");
    const size_t n = (size_t)g->num_nodes;
    const size_t need = (K >= 3 ? 2 : (K == 2 ? 1 : 0)) * n * D * sizeof(float);
    LGCN_REQUIRE(work_bytes >= need && (need == 0 || work), LGCN_E_WORKSPACE,
                 "propagate_bwd: workspace %zu < %zu bytes", work_bytes, need);
    const Table e0{user_w, item_w, g->num_users};
    float *z[2] = {work, work ? work + n * D : nullptr};
    const Range all{0, g->n_out_tasks, 0, g->num_nodes, true};
    for (int j = 1; j <= K; ++j) {
        const float *zin = j == 1 ? nullptr : z[j & 1];          // written by layer j-1
        float *zout = j == K ? nullptr : z[(j - 1) & 1];
        int rc = bwd_layer_impl(g, G, j, K, zin, zout, e0, neg_count, reg_coef, grad, accum, all, st, local_only());
        if (rc) return rc;
    }
    return LGCN_OK;
}

}  // namespace lgcn

extern "C" int lgcn_propagate_fwd(const lgcn_graph *g, const float *user_w, const float *item_w,
                                  int num_layers, float *final_out, float *rnorm, float *work,
                                  size_t work_bytes, void *stream) {
    return lgcn::propagate_fwd_impl(g, user_w, item_w, num_layers, final_out, rnorm, work, work_bytes,
                                    (cudaStream_t)stream);
}

extern "C" int lgcn_propagate_bwd(const lgcn_graph *g, const float *grad_final, int num_layers,
                                  const float *user_w, const float *item_w, const int32_t *neg_count,
                                  float reg_coef, float *grad_e0, double *accum, float *work,
                                  size_t work_bytes, void *stream) {
    return lgcn::propagate_bwd_impl(g, grad_final, num_layers, user_w, item_w, neg_count, reg_coef,
                                    grad_e0, accum, work, work_bytes, (cudaStream_t)stream);
}

// ---- sharded (owner-computes-by-row-range) entry points -------------------------------------------

extern "C" int lgcn_prescale(const lgcn_graph *g, const float *user_w, const float *item_w, int64_t row_begin,
                             int64_t row_end, float *y0, const lgcn_peers *peers, void *stream) {
    using namespace lgcn;
    LGCN_REQUIRE(g && user_w && item_w && y0 && row_begin >= 0 && row_end <= g->num_nodes, LGCN_E_INVALID,
                 "prescale: bad argument");
    if (row_end <= row_begin) return LGCN_OK;
    prescale_kernel<<<cdiv(row_end - row_begin, 2 * WARPS_PER_CTA), CTA_THREADS, 0, (cudaStream_t)stream>>>(
        Table{user_w, item_w, g->num_users}, g->dis, (int)row_begin, (int)row_end, y0, make_peers(peers));
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

extern "C" int lgcn_fwd_layer_ex(const lgcn_graph *g, const float *user_w, const float *item_w, int k, int num_layers,
                                 const float *yin, float *yout, const float *y1, const float *y2, const float *y3,
                                 float *final_out, float *rnorm, int task_begin, int task_end, int64_t row_begin,
                                 int64_t row_end, int flags, const lgcn_peers *peers, void *stream) {
    using namespace lgcn;
    LGCN_REQUIRE(g && user_w && item_w && yin && k >= 1 && k <= num_layers && num_layers <= 4, LGCN_E_INVALID,
                 "fwd_layer: bad argument");
    LGCN_REQUIRE(k == num_layers ? final_out != nullptr : yout != nullptr, LGCN_E_INVALID, "fwd_layer: missing output");
    LGCN_REQUIRE(task_begin >= 0 && task_end <= g->n_in_tasks && task_begin <= task_end, LGCN_E_INVALID,
                 "fwd_layer: task range [%d,%d) outside [0,%d)", task_begin, task_end, g->n_in_tasks);
    const int norm_out = (flags & LGCN_FWD_NORMALIZED) ? 1 : 0;
    LGCN_REQUIRE(!norm_out || k < num_layers || rnorm, LGCN_E_INVALID, "fwd_layer: LGCN_FWD_NORMALIZED needs rnorm");
    return fwd_layer_impl(g, Table{user_w, item_w, g->num_users}, k, num_layers, false, yin, yout, y1, y2, y3, final_out,
                          rnorm, Range{task_begin, task_end, (int)row_begin, (int)row_end, true}, (cudaStream_t)stream,
                          make_peers(peers), norm_out);
}

extern "C" int lgcn_fwd_layer(const lgcn_graph *g, const float *user_w, const float *item_w, int k, int num_layers,
                              const float *yin, float *yout, const float *y1, const float *y2, const float *y3,
                              float *final_out, float *rnorm, int task_begin, int task_end, int64_t row_begin,
                              int64_t row_end, const lgcn_peers *peers, void *stream) {
    return lgcn_fwd_layer_ex(g, user_w, item_w, k, num_layers, yin, yout, y1, y2, y3, final_out, rnorm, task_begin,
                             task_end, row_begin, row_end, 0, peers, stream);
}

extern "C" int lgcn_bwd_layer(const lgcn_graph *g, const float *grad_final, int j, int num_layers, const float *zin,
                              float *zout, const float *user_w, const float *item_w, const int32_t *neg_count,
                              float reg_coef, float *grad_e0, double *accum, int task_begin, int task_end,
                              int64_t row_begin, int64_t row_end, const lgcn_peers *peers, void *stream) {
    using namespace lgcn;
    LGCN_REQUIRE(g && grad_final && j >= 1 && j <= num_layers && num_layers <= 4, LGCN_E_INVALID, "bwd_layer: bad argument");
    LGCN_REQUIRE(j == 1 || zin, LGCN_E_INVALID, "bwd_layer: zin missing");   // j == 1: zin = dis (.) G (optional)
    LGCN_REQUIRE(j == num_layers ? grad_e0 != nullptr : zout != nullptr, LGCN_E_INVALID, "bwd_layer: missing output");
    LGCN_REQUIRE(task_begin >= 0 && task_end <= g->n_out_tasks && task_begin <= task_end, LGCN_E_INVALID,
                 "bwd_layer: task range outside the list");
    LGCN_REQUIRE(reg_coef == 0.f || (user_w && item_w && neg_count && accum), LGCN_E_INVALID,
                 "bwd_layer: regulariser needs weights, neg_count and accum");
    return bwd_layer_impl(g, grad_final, j, num_layers, zin, zout, Table{user_w, item_w, g->num_users}, neg_count,
                          reg_coef, grad_e0, accum, Range{task_begin, task_end, (int)row_begin, (int)row_end, true},
                          (cudaStream_t)stream, make_peers(peers));
}
