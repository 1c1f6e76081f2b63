// Diagnostics: the speed of light of the propagation kernels' access pattern -- independent random
// 256-byte row gathers, 16 lanes x 128-bit per row, as many rows in flight as registers allow, no
// index array and no dependent address.  tools/gather_peak.py times it over a table that fits the L2
// (ML-25M: 56.7 MB) and over one that does not (10x: 563 MB); bench.py quotes the propagation layer's
// gather-model GB/s against these measured ceilings next to the compulsory-byte HBM roofline.
#include "common.cuh"

namespace lgcn {

__global__ void __launch_bounds__(256)
probe_gather_kernel(const float4 *__restrict__ table, unsigned nrows, int rows_per_halfwarp, float *__restrict__ sink) {
    const int lane = threadIdx.x & 31, l16 = lane & 15;
    const unsigned hw = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;            // global half-warp id
    unsigned state = hw * 2654435761u + 12345u;
    float4 acc = f4zero();
    for (int i = 0; i < rows_per_halfwarp; i += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            state = state * 1664525u + 1013904223u;                              // LCG: same value in all 16 lanes
            const unsigned row = (unsigned)(((unsigned long long)(state >> 4) * nrows) >> 28);
            v[u] = ldg4(table + (size_t)row * D4 + l16);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) f4add(acc, v[u]);
    }
    if (acc.x == 123.456f) sink[hw] = acc.y + acc.z + acc.w;                     // keeps the loads alive
}

}  // namespace lgcn

// Gathers ctas * 16 * rows_per_halfwarp random rows of table [nrows,64] (fp32).  The caller times it.
extern "C" int lgcn_probe_gather(const float *table, int64_t nrows, int ctas, int rows_per_halfwarp, float *sink,
                                 void *stream) {
    using namespace lgcn;
    LGCN_REQUIRE(table && sink && nrows > 0 && nrows < (1ll << 28) && ctas > 0 && rows_per_halfwarp > 0, LGCN_E_INVALID,
                 "probe_gather: bad argument");
    probe_gather_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(table), (unsigned)nrows,
                                                                rows_per_halfwarp, sink);
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}
