// Operator-level entry points used by the drop-in Python API when callers compose the reference's
// own pieces instead of the fused training step:
//   lgcn_spmm      one LGConv layer  out = A x  (or A^T x): `conv(x=emb, edge_index=edge_index)`
//                  (/root/reference/models/light_gcn.py:33)
//   lgcn_bpr_rows  bpr_loss on six already-gathered [P,64] tensors, value and gradients
//                  (/root/reference/utils/train_test.py:18-64)
#include "rowtask.cuh"

namespace lgcn {

int loss_finalize_impl(const double *, int64_t, float, float *, cudaStream_t);

struct SpmmItem {
    int nbr; float w; float4 v;
    __device__ __forceinline__ SpmmItem shfl(int src) const {
        SpmmItem r; r.nbr = __shfl_sync(FULL, nbr, src); r.w = __shfl_sync(FULL, w, src); r.v = f4zero(); return r;
    }
};

struct SpmmOp {
    static constexpr bool kExtras = false;
    double *extra0, *extra1;
    const int32_t *nbr;
    const float *dis, *x;
    float *out;
    __device__ __forceinline__ void accumulate(int, int begin, int end, int lane, float4 &acc, float &, float &, float &) const {
        const int l16 = lane & 15;
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        for_each_edge<SpmmItem>(
            begin, end, lane,
            [&](int e) { SpmmItem it; it.nbr = e >= 0 ? __ldg(nbr + e) : -1; it.w = e >= 0 ? __ldg(dis + it.nbr) : 0.f; it.v = f4zero(); return it; },
            [&](int, SpmmItem &it) { if (it.nbr >= 0) it.v = ldg4(x4 + (size_t)it.nbr * D4 + l16); },
            [&](int, SpmmItem &it) { f4fma(acc, it.w, it.v); });
    }
    __device__ __forceinline__ void epilogue(int row, int lane, const float4 &raw, float, float &, float &) const {
        if (lane < 16) reinterpret_cast<float4 *>(out)[(size_t)row * D4 + lane] = f4scale(__ldg(dis + row), raw);
    }
};

// One half-warp per triplet row.
__global__ void __launch_bounds__(CTA_THREADS)
bpr_rows_kernel(const float4 *__restrict__ uf, const float4 *__restrict__ u0, const float4 *__restrict__ pf,
                const float4 *__restrict__ p0, const float4 *__restrict__ nf, const float4 *__restrict__ n0,
                int64_t P, float coeff, double *accum, const float *grad_scale, float4 *g_uf, float4 *g_u0,
                float4 *g_pf, float4 *g_p0, float4 *g_nf, float4 *g_n0) {
    const int lane = threadIdx.x & 31, l16 = lane & 15, wid = threadIdx.x >> 5;
    const int64_t t = ((int64_t)blockIdx.x * WARPS_PER_CTA + wid) * 2 + (lane >> 4);
    const bool ok = t < P;
    const size_t o = (size_t)(ok ? t : 0) * D4 + l16;
    float4 a = f4zero(), b = f4zero(), c = f4zero(), a0 = f4zero(), b0 = f4zero(), c0 = f4zero();
    if (ok) { a = ldg4(uf + o); b = ldg4(pf + o); c = ldg4(nf + o); a0 = ldg4(u0 + o); b0 = ldg4(p0 + o); c0 = ldg4(n0 + o); }
    const float ra = 1.0f / sqrtf(half_sum(f4dot(a, a))), rb = 1.0f / sqrtf(half_sum(f4dot(b, b))),
                rc = 1.0f / sqrtf(half_sum(f4dot(c, c)));
    const float cp = half_sum(f4dot(a, b)) * ra * rb, cn = half_sum(f4dot(a, c)) * ra * rc;
    const float x = 10.f * (cp - cn);
    float sp = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
    float reg = half_sum(f4dot(a0, a0) + f4dot(b0, b0) + f4dot(c0, c0));
    if (!ok || l16 != 0) { sp = 0.f; reg = 0.f; }
    if (accum) {
        sp = warp_sum(sp);
        reg = warp_sum(reg);
        __shared__ float s[WARPS_PER_CTA][2];
        if (lane == 0) { s[wid][0] = sp; s[wid][1] = reg; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double x0 = 0.0, x1 = 0.0;
            for (int w = 0; w < WARPS_PER_CTA; ++w) { x0 += s[w][0]; x1 += s[w][1]; }
            atomicAdd(accum, x0);
            atomicAdd(accum + 1, x1);
        }
    }
    if (g_uf && ok) {
        const float gs = grad_scale ? __ldg(grad_scale) : 1.0f;
        const float invP = 1.0f / (float)P;
        const float s = -gs * invP / (1.f + expf(-x));           // dL/dcos+
        const float4 ah = f4scale(ra, a), bh = f4scale(rb, b), ch = f4scale(rc, c);
        float4 gu = f4scale(s, bh); f4fma(gu, -s, ch); f4fma(gu, -s * (cp - cn), ah);
        g_uf[o] = f4scale(ra, gu);
        float4 gp = f4scale(s, ah); f4fma(gp, -s * cp, bh);
        g_pf[o] = f4scale(rb, gp);
        float4 gn = f4scale(-s, ah); f4fma(gn, s * cn, ch);
        g_nf[o] = f4scale(rc, gn);
        const float r = gs * 2.0f * coeff * invP / 64.0f;
        g_u0[o] = f4scale(r, a0); g_p0[o] = f4scale(r, b0); g_n0[o] = f4scale(r, c0);
    }
}

// Cross-rank barrier for the fused all-gather: every rank bumps its private epoch, publishes it into
// slot [rank] of every rank's flag array (symmetric memory; release at system scope AFTER a system
// fence, so the rows stored by the preceding kernels of this stream are visible first), then waits
// until all slots of its own array have reached the epoch.  State lives on the device => the launch
// can be replayed from a CUDA graph.  Each GPU runs its own instance; nothing waits on a co-resident
// kernel of the same GPU.
__global__ void peer_barrier_kernel(int *flags_local, int *epoch, Peers P, int rank) {
    __shared__ int e_sh;
    if (threadIdx.x == 0) {
        const int e = epoch[0] + 1;
        epoch[0] = e;
        e_sh = e;
        __threadfence_system();
    }
    __syncthreads();
    const int e = e_sh;
    if ((int)threadIdx.x < P.world) {
        int *dst = reinterpret_cast<int *>(reinterpret_cast<char *>(flags_local + rank) + P.delta[threadIdx.x]);
        asm volatile("st.release.sys.global.s32 [%0], %1;" :: "l"(dst), "r"(e) : "memory");
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags_local + threadIdx.x) : "memory");
        } while (v < e);
    }
    __syncthreads();
    __threadfence_system();
}

// The step's three global sums (softplus sum, regulariser sum, ||grad||^2; accum[0..3], doubles) across ranks, riding
// on the same flag protocol as the barrier: every rank stores its four doubles into slot [rank] of every rank's slot
// table (symmetric memory), publishes its epoch, waits for all, then sums the slots IN RANK ORDER (every rank gets the
// bit-identical total).  Replaces an NCCL all-reduce of 32 bytes; also a full barrier.
__global__ void peer_allreduce4_kernel(int *flags_local, int *epoch, Peers P, int rank, double *slots_local,
                                       double *accum) {
    __shared__ int e_sh;
    if ((int)threadIdx.x < P.world * 4) {
        const int peer = threadIdx.x >> 2, j = threadIdx.x & 3;
        double *dst = reinterpret_cast<double *>(reinterpret_cast<char *>(slots_local + rank * 4 + j) + P.delta[peer]);
        const double v = accum[j];
        asm volatile("st.relaxed.sys.global.f64 [%0], %1;" :: "l"(dst), "d"(v) : "memory");
        __threadfence_system();
    }
    if (threadIdx.x == 0) {
        const int e = epoch[0] + 1;
        epoch[0] = e;
        e_sh = e;
        __threadfence_system();
    }
    __syncthreads();
    const int e = e_sh;
    if ((int)threadIdx.x < P.world) {
        int *dst = reinterpret_cast<int *>(reinterpret_cast<char *>(flags_local + rank) + P.delta[threadIdx.x]);
        asm volatile("st.release.sys.global.s32 [%0], %1;" :: "l"(dst), "r"(e) : "memory");
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags_local + threadIdx.x) : "memory");
        } while (v < e);
    }
    __syncthreads();
    __threadfence_system();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int r = 0; r < P.world; ++r) {
            double v;
            asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(slots_local + r * 4 + threadIdx.x) : "memory");
            s += v;
        }
        accum[threadIdx.x] = s;
    }
}

}  // namespace lgcn

extern "C" int lgcn_peer_allreduce4(const lgcn_peers *peers, int32_t *flags_local, int32_t *epoch, double *slots_local,
                                    double *accum, void *stream) {
    using namespace lgcn;
    LGCN_REQUIRE(peers && flags_local && epoch && slots_local && accum && peers->world >= 1 && peers->world <= 8,
                 LGCN_E_INVALID, "peer_allreduce4: bad argument");
    if (peers->world == 1) return LGCN_OK;
    peer_allreduce4_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags_local, epoch, make_peers(peers), peers->rank,
                                                              slots_local, accum);
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

extern "C" int lgcn_peer_barrier(const lgcn_peers *peers, int32_t *flags_local, int32_t *epoch, void *stream) {
    using namespace lgcn;
    LGCN_REQUIRE(peers && flags_local && epoch && peers->world >= 1 && peers->world <= 8, LGCN_E_INVALID,
                 "peer_barrier: bad argument");
    if (peers->world == 1) return LGCN_OK;
    Peers P = make_peers(peers);
    // the flag array must be addressed through the same symmetric region as the tables
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags_local, epoch, P, peers->rank);
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

extern "C" int lgcn_spmm(const lgcn_graph *g, const float *x, float *out, int transpose, void *stream) {
    using namespace lgcn;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(g && x && out, LGCN_E_INVALID, "spmm: null argument");
    if (g->num_active < g->num_nodes)
        LGCN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)g->num_nodes * D, st));
    SpmmOp op{nullptr, nullptr, transpose ? g->out_nbr : g->in_nbr, g->dis, x, out};
    LGCN_CUDA(launch_rowtasks(op, transpose ? g->out_tasks : g->in_tasks, 0, transpose ? g->n_out_tasks : g->n_in_tasks,
                              g->partials, g->slot_counters, g->sched, st));
    return LGCN_OK;
}

extern "C" int lgcn_bpr_rows(const float *uf, const float *u0, const float *pf, const float *p0, const float *nf,
                             const float *n0, int64_t P, float coeff, double *accum, float *loss_out,
                             const float *grad_scale, float *g_uf, float *g_u0, float *g_pf, float *g_p0,
                             float *g_nf, float *g_n0, void *stream) {
    using namespace lgcn;
    cudaStream_t st = (cudaStream_t)stream;
    LGCN_REQUIRE(uf && u0 && pf && p0 && nf && n0 && P > 0, LGCN_E_INVALID, "bpr_rows: null argument or P == 0");
    LGCN_REQUIRE(!g_uf || (g_u0 && g_pf && g_p0 && g_nf && g_n0), LGCN_E_INVALID, "bpr_rows: partial gradient set");
    LGCN_REQUIRE(!loss_out || accum, LGCN_E_INVALID, "bpr_rows: loss_out needs accum");
    if (accum) LGCN_CUDA(cudaMemsetAsync(accum, 0, 2 * sizeof(double), st));
    auto f4 = [](const float *p) { return reinterpret_cast<const float4 *>(p); };
    auto m4 = [](float *p) { return reinterpret_cast<float4 *>(p); };
    bpr_rows_kernel<<<cdiv(P, 2 * WARPS_PER_CTA), CTA_THREADS, 0, st>>>(f4(uf), f4(u0), f4(pf), f4(p0), f4(nf), f4(n0), P,
                                                                      coeff, accum, grad_scale, m4(g_uf), m4(g_u0),
                                                                      m4(g_pf), m4(g_p0), m4(g_nf), m4(g_n0));
    LGCN_LAUNCH_CHECK();
    if (loss_out) return loss_finalize_impl(accum, P, coeff, loss_out, st);
    return LGCN_OK;
}
