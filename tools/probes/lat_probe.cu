// Dependent-load latency on this GPU, in the access shapes the persistent step kernel uses.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probes/lat_probe tools/probes/lat_probe.cu
// A chain of row indices is laid out in a table of 256-byte rows (row r holds next(r) in all 64 floats); warps chase it
// with half-warp float4 row loads (ld.global.cg / ld.global.nc), so one hop = one dependent L2 (or DRAM) round trip.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <numeric>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void fill(float4 *tab, const int *next, int rows) {
    const int r = blockIdx.x * (blockDim.x / 16) + threadIdx.x / 16, l = threadIdx.x & 15;
    if (r < rows) { const float v = __int_as_float(next[r]); tab[(size_t)r * 16 + l] = make_float4(v, v, v, v); }
}

template <int MODE>   // 0: ld.global.cg   1: ld.global.nc (__ldg)
__global__ void chase(const float4 *tab, int rows, int hops, int warps_per_sm_active, long long *out, int *sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= warps_per_sm_active) return;
    int idx = (int)(((size_t)blockIdx.x * 131 + warp * 977 + (lane >> 4) * 7919) % rows);
    long long t0 = clock64();
    for (int h = 0; h < hops; ++h) {
        const float4 *p = tab + (size_t)idx * 16 + (lane & 15);
        float4 v = MODE == 0 ? __ldcg(p) : __ldg(p);
        idx = __float_as_int(v.x);
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
    if (idx == -12345) sink[0] = idx;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs, clock %.0f MHz\n", prop.name, sms, prop.clockRate / 1e3);
    long long *out; int *sink;
    CK(cudaMalloc(&out, sizeof(long long) * sms * 32));
    CK(cudaMalloc(&sink, 4));
    for (size_t mb : {1, 32, 512}) {
        const int rows = (int)(mb * 1024 * 1024 / 256);
        std::vector<int> perm(rows), next(rows);
        std::iota(perm.begin(), perm.end(), 0);
        std::mt19937 rng(1);
        std::shuffle(perm.begin(), perm.end(), rng);
        for (int i = 0; i < rows; ++i) next[perm[i]] = perm[(i + 1) % rows];
        float4 *tab; int *dnext;
        CK(cudaMalloc(&tab, (size_t)rows * 256));
        CK(cudaMalloc(&dnext, (size_t)rows * 4));
        CK(cudaMemcpy(dnext, next.data(), (size_t)rows * 4, cudaMemcpyHostToDevice));
        for (int mode = 0; mode < 2; ++mode)
            for (int w : {1, 8, 24}) {
                fill<<<(rows + 15) / 16, 256>>>(tab, dnext, rows);     // freshly written by other SMs
                CK(cudaDeviceSynchronize());
                const int hops = 2000;
                for (int rep = 0; rep < 2; ++rep) {
                    if (mode == 0) chase<0><<<sms, 32 * 24>>>(tab, rows, hops, w, out, sink);
                    else chase<1><<<sms, 32 * 24>>>(tab, rows, hops, w, out, sink);
                    CK(cudaDeviceSynchronize());
                }
                std::vector<long long> h(sms * 32);
                CK(cudaMemcpy(h.data(), out, sizeof(long long) * sms * 32, cudaMemcpyDeviceToHost));
                double sum = 0, mx = 0; int n = 0;
                for (int b = 0; b < sms; ++b) for (int i = 0; i < w; ++i) { double c = (double)h[b * 32 + i] / hops; sum += c; mx = std::max(mx, c); ++n; }
                printf("table %4zu MB  %s  %2d warps/SM (x2 half-warp chains): %.0f cycles/hop mean, %.0f max  (= %.0f ns at %.0f MHz)\n",
                       mb, mode == 0 ? "ld.cg" : "ld.nc", w, sum / n, mx, sum / n / (prop.clockRate / 1e6), prop.clockRate / 1e3);
            }
        CK(cudaFree(tab)); CK(cudaFree(dnext));
    }
    return 0;
}
