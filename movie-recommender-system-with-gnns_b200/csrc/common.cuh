// Shared device/host helpers for the sm_100a LightGCN kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "lgcn_b200.h"

namespace lgcn {

constexpr int D = LGCN_DIM;                 // 64 fp32 = 256 B per embedding row
constexpr int D4 = D / 4;                   // float4 per row (one per lane of a half-warp)
#ifndef LGCN_WARPS
#define LGCN_WARPS 8
#endif
#ifndef LGCN_UNROLL
#define LGCN_UNROLL 8
#endif
constexpr int WARPS_PER_CTA = LGCN_WARPS;
constexpr int CTA_THREADS = WARPS_PER_CTA * 32;
constexpr int PARTIAL_STRIDE = 80;          // floats per partial slot: 64 + scalar, 64 B aligned
constexpr unsigned FULL = 0xffffffffu;

void set_error(const char *fmt, ...);

#define LGCN_REQUIRE(cond, code, ...)                                                         \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ::lgcn::set_error(__VA_ARGS__);                                                   \
            return (code);                                                                    \
        }                                                                                     \
    } while (0)

#define LGCN_CUDA(expr)                                                                       \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::lgcn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,                    \
                              cudaGetErrorString(_e));                                        \
            return LGCN_E_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define LGCN_LAUNCH_CHECK() LGCN_CUDA(cudaGetLastError())

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// The embedding table e0 = cat(user_embedding.weight, item_embedding.weight)
// (models/light_gcn.py:29) without materialising the concatenation.
struct Table {
    const float *user;
    const float *item;
    int num_users;
    __device__ __forceinline__ const float4 *row4(int r) const {
        const float *p = r < num_users ? user + (size_t)r * D : item + (size_t)(r - num_users) * D;
        return reinterpret_cast<const float4 *>(p);
    }
};

// Fused compute + all-gather: where to store a produced row so that EVERY rank's copy of the table
// receives it.  The tables live in symmetric memory (same offset on every rank), so a peer address is
// the local address plus a per-peer delta; with NVLS a single store to the multicast alias is
// replicated by the NVSwitch.  world <= 1: plain local store.
struct Peers {
    int world;
    int use_mc;
    long long mc_delta;          // multicast alias = local address + mc_delta
    long long delta[8];          // peer p = local address + delta[p]   (delta[own rank] = 0)
};

static inline Peers local_only() {
    Peers p{};
    p.world = 1;
    return p;
}

static inline Peers make_peers(const lgcn_peers *q) {
    if (!q || q->world <= 1) return local_only();
    Peers p{};
    p.world = q->world;
    const long long self = (long long)(uintptr_t)q->base[q->rank];
    p.use_mc = q->mc_base != nullptr;
    p.mc_delta = p.use_mc ? (long long)(uintptr_t)q->mc_base - self : 0;
    for (int i = 0; i < q->world && i < 8; ++i) p.delta[i] = (long long)(uintptr_t)q->base[i] - self;
    return p;
}

__device__ __forceinline__ void push4(float4 *local, const float4 &v, const Peers &P) {
    if (P.world <= 1) { *local = v; return; }
    if (P.use_mc) {
        float4 *mc = reinterpret_cast<float4 *>(reinterpret_cast<char *>(local) + P.mc_delta);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < P.world) *reinterpret_cast<float4 *>(reinterpret_cast<char *>(local) + P.delta[i]) = v;
}

__device__ __forceinline__ void push1(float *local, float v, const Peers &P) {
    if (P.world <= 1) { *local = v; return; }
    if (P.use_mc) {
        float *mc = reinterpret_cast<float *>(reinterpret_cast<char *>(local) + P.mc_delta);
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(mc), "f"(v) : "memory");
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < P.world) *reinterpret_cast<float *>(reinterpret_cast<char *>(local) + P.delta[i]) = v;
}

// tasks [tb,te) of a row-sorted task list and the node rows [rb,re) they cover; inactive_rows: also
// produce the rows without any incident edge (dense semantics) -- false in the sparse step.
struct Range {
    int tb, te, rb, re;
    bool inactive_rows;
};

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4add(float4 &a, const float4 &b) {
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
__device__ __forceinline__ void f4fma(float4 &a, float s, const float4 &b) {
    a.x = fmaf(s, b.x, a.x); a.y = fmaf(s, b.y, a.y); a.z = fmaf(s, b.z, a.z); a.w = fmaf(s, b.w, a.w);
}
__device__ __forceinline__ float4 f4scale(float s, const float4 &b) {
    return make_float4(s * b.x, s * b.y, s * b.z, s * b.w);
}
__device__ __forceinline__ float f4dot(const float4 &a, const float4 &b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
// sum over the 16 lanes of a half-warp (xor 8,4,2,1 never crosses the half boundary)
__device__ __forceinline__ float half_sum(float v) {
    v += __shfl_xor_sync(FULL, v, 8);
    v += __shfl_xor_sync(FULL, v, 4);
    v += __shfl_xor_sync(FULL, v, 2);
    v += __shfl_xor_sync(FULL, v, 1);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(FULL, v, 16);
    return half_sum(v);
}
__device__ __forceinline__ float4 f4shfl_xor16(const float4 &a) {
    return make_float4(__shfl_xor_sync(FULL, a.x, 16), __shfl_xor_sync(FULL, a.y, 16),
                       __shfl_xor_sync(FULL, a.z, 16), __shfl_xor_sync(FULL, a.w, 16));
}

}  // namespace lgcn
