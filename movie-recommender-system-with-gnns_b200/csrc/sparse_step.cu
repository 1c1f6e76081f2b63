// Sparse training step: the loop body of /root/reference/utils/train_test.py:88-96 with work
// proportional to the rows a Cluster-GCN batch TOUCHES instead of the table size N.
//
// The reference pays full-table passes for every batch: the forward
// rebuilds all N rows, autograd materialises dense [N,64] gradients, clip_grad_norm_ and Adam stream
// both tables and their moments.  At ML-25M shape a batch has ~2 k nodes with edges and a few
// thousand sampled negatives out of N = 221,588 rows, so > 95 % of that traffic moves rows whose
// gradient is exactly zero.  For such a row Adam's update is a pure function of its own (p, m, v)
// and the step number, so it can be REPLAYED later with bit-identical arithmetic:
//
//   row_step[r] = last optimiser step applied to row r.
//   before a step reads row r (forward: active rows; BPR: sampled negatives) the pending
//   zero-gradient steps row_step[r]+1 .. t-1 are replayed in registers (no extra HBM traffic);
//   after the backward pass Adam step t is applied to the touched rows only.
//   lgcn_adam_flush replays everything that is still pending (end of epoch / before evaluation).
//
// Touched rows = the graph's active list (static, built by K0) + the distinct INACTIVE negatives
// of the step (deduplicated on the device with a per-item step stamp).  dL/dfinal (G) and the
// negative histogram are kept all-zero BETWEEN steps by zeroing exactly the rows that were used.
#include "adam.cuh"
#include "rowtask.cuh"

namespace lgcn {

int fwd_layer_impl(const lgcn_graph *, const Table &, int, int, bool, const float *, float *, const float *,
                   const float *, const float *, float *, float *, Range, cudaStream_t, const Peers &);
int bwd_layer_impl(const lgcn_graph *, const float *, int, int, const float *, float *, const Table &,
                   const int32_t *, float, float *, double *, Range, cudaStream_t, const Peers &);
int bpr_sparse_impl(const lgcn_graph *, const float *, const float *, const int64_t *, float *, int32_t *, float *,
                    double *, const float *, const float *, int, cudaStream_t);
__global__ void step_begin_kernel(int64_t *step, double *accum, int32_t *list_count);

struct MutTable {
    float *user, *item;
    int num_users;
    __device__ __forceinline__ float4 *row4(int r) const {
        float *p = r < num_users ? user + (size_t)r * D : item + (size_t)(r - num_users) * D;
        return reinterpret_cast<float4 *>(p);
    }
};

// distinct inactive negatives of this step -> neg_list (order irrelevant: rows are independent)
__global__ void mark_negs_kernel(const int64_t *__restrict__ neg, int64_t P, const int64_t *__restrict__ step,
                                 const uint8_t *__restrict__ active, int num_users, int32_t *__restrict__ neg_flag,
                                 int32_t *__restrict__ neg_list, int32_t *__restrict__ count) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P) return;
    const int i = (int)neg[t];
    const int stamp = (int)step[0];
    if (active[num_users + i]) return;                       // handled through the active list
    if (atomicExch(neg_flag + i, stamp) != stamp) neg_list[atomicAdd(count, 1)] = i;
}

// Replay the pending zero-gradient Adam steps of the listed rows up to step (t + delta).
// One half-warp per row; list == null: rows [0,n) (flush).  offset: added to list entries (items).
__global__ void __launch_bounds__(CTA_THREADS)
adam_replay_kernel(const int32_t *__restrict__ list, const int32_t *__restrict__ count_ptr, int n, int offset,
                   MutTable w, float4 *__restrict__ m, float4 *__restrict__ v, int32_t *__restrict__ row_step,
                   const int64_t *__restrict__ step, int delta, AdamHyper h) {
    const int lane = threadIdx.x & 31, l16 = lane & 15;
    const int idx = (blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5)) * 2 + (lane >> 4);
    const int cnt = count_ptr ? *count_ptr : n;
    bool valid = idx < cnt;
    const int row = valid ? (list ? list[idx] + offset : idx) : 0;
    const int target = (int)step[0] + delta;
    const int from = valid ? row_step[row] : target;
    valid = valid && from < target;
    float4 *pp = w.row4(row) + l16;
    const size_t o = (size_t)row * D4 + l16;
    float4 p4 = f4zero(), m4 = f4zero(), v4 = f4zero();
    if (valid) { p4 = *pp; m4 = m[o]; v4 = v[o]; }
    // a row whose moments are all zero (never received a gradient) does not move
    const bool live = m4.x != 0.f || m4.y != 0.f || m4.z != 0.f || m4.w != 0.f ||
                      v4.x != 0.f || v4.y != 0.f || v4.z != 0.f || v4.w != 0.f;
    const unsigned half_mask = 0xffffu << (lane & 16);
    const bool any_live = (__ballot_sync(FULL, live) & half_mask) != 0u;      // no lane has exited yet
    if (valid && any_live) {
        const float4 zero = f4zero();
        if (h.bc_table && target < h.bc_len) {
            // steps chain only through one fma each for p, m, v: unrolling lets the sqrt / division sequences of
            // neighbouring steps overlap (the flush is MUFU- and latency-bound, not bandwidth-bound)
#pragma unroll 4
            for (int t = from + 1; t <= target; ++t) {
                const AdamScalars a = adam_scalars_tab(h, t);
                adam_vec(p4, m4, v4, zero, 1.0f, a);
            }
        } else {
            for (int t = from + 1; t <= target; ++t) {
                const AdamScalars a = adam_scalars(h, t);
                adam_vec(p4, m4, v4, zero, 1.0f, a);
            }
        }
        *pp = p4; m[o] = m4; v[o] = v4;
    }
    if (valid && l16 == 0) row_step[row] = target;
}

// Inactive sampled negatives: grad = G/(K+1)^2 + reg (nothing propagates to a row without edges).
__global__ void __launch_bounds__(CTA_THREADS)
neg_rows_grad_kernel(const int32_t *__restrict__ neg_list, const int32_t *__restrict__ count_ptr, int num_users,
                     float *__restrict__ G, float c0, Table e0, const int32_t *__restrict__ neg_count, float reg_coef,
                     float *__restrict__ grad, double *extra0, double *extra1) {
    const int lane = threadIdx.x & 31, l16 = lane & 15, wid = threadIdx.x >> 5;
    const int idx = (blockIdx.x * WARPS_PER_CTA + wid) * 2 + (lane >> 4);
    const bool ok = idx < *count_ptr;
    float4 g = f4zero();
    float reg = 0.f;
    if (ok) {
        const int item = neg_list[idx], row = num_users + item;
        float4 *gp = reinterpret_cast<float4 *>(G) + (size_t)row * D4 + l16;
        g = f4scale(c0, *gp);
        *gp = f4zero();                                       // keep G all-zero between steps
        const int cnt = neg_count[item];
        if (reg_coef != 0.f && cnt) {
            const float4 e = ldg4(e0.row4(row) + l16);
            f4fma(g, reg_coef * (float)cnt, e);
            reg = (float)cnt * f4dot(e, e);
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
        if (a != 0.0) atomicAdd(extra0, a);
        if (b != 0.0) atomicAdd(extra1, b);
    }
}

// Adam step t on the listed rows; also restores the all-zero invariants for the rows it visits.
__global__ void __launch_bounds__(CTA_THREADS)
adam_rows_kernel(const int32_t *__restrict__ list, const int32_t *__restrict__ count_ptr, int n, int offset,
                 MutTable w, const float4 *__restrict__ grad, float4 *__restrict__ m, float4 *__restrict__ v,
                 int32_t *__restrict__ row_step, float4 *__restrict__ G, int32_t *__restrict__ neg_count,
                 const double *__restrict__ accum, const int64_t *__restrict__ step, AdamHyper h, int64_t P,
                 float coeff, float *loss_out) {
    const int lane = threadIdx.x & 31, l16 = lane & 15;
    const int idx = (blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5)) * 2 + (lane >> 4);
    if (loss_out && blockIdx.x == 0 && threadIdx.x == 0) {
        const double p = (double)P;
        loss_out[0] = (float)(-accum[0] / (10.0 * p) + (double)coeff * accum[1] / (64.0 * p));
    }
    const int cnt = count_ptr ? *count_ptr : n;
    if (idx >= cnt) return;
    const int row = list[idx] + offset;
    const long long t = step[0];
    const AdamScalars a = adam_scalars(h, t);
    const float clip = clip_coef(h, accum[2]);
    float4 *pp = w.row4(row) + l16;
    const size_t o = (size_t)row * D4 + l16;
    float4 p4 = *pp, m4 = m[o], v4 = v[o];
    adam_vec(p4, m4, v4, grad[o], clip, a);
    *pp = p4; m[o] = m4; v[o] = v4;
    G[o] = f4zero();
    if (l16 == 0) {
        row_step[row] = (int)t;
        if (row >= w.num_users) neg_count[row - w.num_users] = 0;
    }
}

static inline int grid_rows(int64_t rows) { return cdiv(rows > 0 ? rows : 1, 2 * WARPS_PER_CTA); }

}  // namespace lgcn

extern "C" int lgcn_train_step_sparse(const lgcn_graph *g, float *user_w, float *item_w, int K, const int64_t *neg,
                                      float bpr_coeff, const lgcn_adam *opt, const lgcn_step_buffers *buf,
                                      float *loss_out, void *stream) {
    using namespace lgcn;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(g && user_w && item_w && neg && opt && buf, LGCN_E_INVALID, "train_step_sparse: null argument");
    LGCN_REQUIRE(K >= 1 && K <= 4, LGCN_E_INVALID, "train_step_sparse: num_layers %d outside [1,4]", K);
    LGCN_REQUIRE(opt->row_step && opt->m && opt->v && opt->step, LGCN_E_INVALID,
                 "train_step_sparse: optimiser state (row_step, m, v, step) missing");
    LGCN_REQUIRE(buf->final_emb && buf->rnorm && buf->grad_final && buf->grad_e0 && buf->neg_count && buf->accum &&
                 buf->trip_scratch && buf->neg_flag && buf->neg_list && buf->neg_list_count,
                 LGCN_E_INVALID, "train_step_sparse: step buffers missing");
    LGCN_REQUIRE(g->num_triplets > 0, LGCN_E_INVALID, "train_step_sparse: batch has no user->movie edge");
    const size_t n = (size_t)g->num_nodes;
    const size_t need = (size_t)(K - 1 > 2 ? K - 1 : 2) * n * D * sizeof(float);
    LGCN_REQUIRE(buf->work && buf->work_bytes >= need, LGCN_E_WORKSPACE, "train_step_sparse: work %zu < %zu",
                 buf->work_bytes, need);
    const int num_items = g->num_nodes - g->num_users;
    const int64_t P = g->num_triplets;
    const int64_t max_negs = P < num_items ? P : num_items;
    const AdamHyper h = make_hyper(opt);
    const MutTable w{user_w, item_w, g->num_users};
    const Table e0{user_w, item_w, g->num_users};
    float4 *m4 = reinterpret_cast<float4 *>(opt->m), *v4 = reinterpret_cast<float4 *>(opt->v);

    step_begin_kernel<<<1, 1, 0, st>>>(opt->step, buf->accum, buf->neg_list_count);
    LGCN_LAUNCH_CHECK();
    mark_negs_kernel<<<cdiv(P, 256), 256, 0, st>>>(neg, P, opt->step, g->active, g->num_users, buf->neg_flag,
                                                   buf->neg_list, buf->neg_list_count);
    LGCN_LAUNCH_CHECK();
    // bring the rows this step reads up to step t-1
    adam_replay_kernel<<<grid_rows(g->num_active), CTA_THREADS, 0, st>>>(g->active_list, nullptr, g->num_active, 0, w, m4,
                                                                        v4, opt->row_step, opt->step, -1, h);
    LGCN_LAUNCH_CHECK();
    adam_replay_kernel<<<grid_rows(max_negs), CTA_THREADS, 0, st>>>(buf->neg_list, buf->neg_list_count, 0, g->num_users, w,
                                                                   m4, v4, opt->row_step, opt->step, -1, h);
    LGCN_LAUNCH_CHECK();
    // forward over the active rows only
    float *y[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int k = 1; k < K; ++k) y[k] = buf->work + (size_t)(k - 1) * n * D;
    const Range fin{0, g->n_in_tasks, 0, g->num_nodes, false};
    int rc;
    for (int k = 1; k <= K; ++k)
        if ((rc = fwd_layer_impl(g, e0, k, K, true, y[k - 1], k < K ? y[k] : nullptr, y[1], y[2], y[3], buf->final_emb,
                                 buf->rnorm, fin, st, local_only()))) return rc;
    if ((rc = bpr_sparse_impl(g, buf->final_emb, buf->rnorm, neg, buf->grad_final, buf->neg_count, buf->trip_scratch,
                              buf->accum, user_w, item_w, K, st))) return rc;
    // backward over the active rows, then the inactive negatives
    const float reg_coef = 2.0f * bpr_coeff / (64.0f * (float)P);
    const float c0 = 1.0f / (float)((K + 1) * (K + 1));
    float *z[2] = {buf->work, buf->work + n * D};
    const Range fout{0, g->n_out_tasks, 0, g->num_nodes, false};
    for (int j = 1; j <= K; ++j)
        if ((rc = bwd_layer_impl(g, buf->grad_final, j, K, j == 1 ? nullptr : z[j & 1], j == K ? nullptr : z[(j - 1) & 1],
                                 e0, buf->neg_count, reg_coef, buf->grad_e0, buf->accum, fout, st, local_only()))) return rc;
    neg_rows_grad_kernel<<<grid_rows(max_negs), CTA_THREADS, 0, st>>>(buf->neg_list, buf->neg_list_count, g->num_users,
                                                                     buf->grad_final, c0, e0, buf->neg_count, reg_coef,
                                                                     buf->grad_e0, buf->accum + 1, buf->accum + 2);
    LGCN_LAUNCH_CHECK();
    // clip + Adam step t on the touched rows
    const float4 *gr = reinterpret_cast<const float4 *>(buf->grad_e0);
    float4 *G4 = reinterpret_cast<float4 *>(buf->grad_final);
    adam_rows_kernel<<<grid_rows(g->num_active), CTA_THREADS, 0, st>>>(g->active_list, nullptr, g->num_active, 0, w, gr, m4,
                                                                      v4, opt->row_step, G4, buf->neg_count, buf->accum,
                                                                      opt->step, h, P, bpr_coeff, loss_out);
    LGCN_LAUNCH_CHECK();
    adam_rows_kernel<<<grid_rows(max_negs), CTA_THREADS, 0, st>>>(buf->neg_list, buf->neg_list_count, 0, g->num_users, w, gr,
                                                                 m4, v4, opt->row_step, G4, buf->neg_count, buf->accum,
                                                                 opt->step, h, P, bpr_coeff, nullptr);
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

extern "C" int lgcn_adam_flush(const lgcn_adam *opt, float *user_w, float *item_w, int64_t num_users, int64_t num_items,
                               void *stream) {
    using namespace lgcn;
    LGCN_REQUIRE(opt && user_w && item_w && opt->row_step && opt->m && opt->v && opt->step, LGCN_E_INVALID,
                 "adam_flush: null argument");
    const int n = (int)(num_users + num_items);
    adam_replay_kernel<<<grid_rows(n), CTA_THREADS, 0, (cudaStream_t)stream>>>(
        nullptr, nullptr, n, 0, MutTable{user_w, item_w, (int)num_users}, reinterpret_cast<float4 *>(opt->m),
        reinterpret_cast<float4 *>(opt->v), opt->row_step, opt->step, 0, make_hyper(opt));
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}
