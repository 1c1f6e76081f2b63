// f4: the voting kernel of the GPU graph partitioner (lgcn_b200/data/partition_gpu.py), the
// alternative to the host METIS call of PyG's ClusterData (/root/reference/data/dataset_handler.py:273).
//
// Balanced label propagation: every node looks at the part labels of its out-neighbours (CSR by source, the same
// adjacency METIS is given) and reports the most frequent one.  One warp per row: a per-warp histogram over the
// P labels in shared memory (shared-memory atomics), then a warp arg-max with ties going to the SMALLEST label so
// that the result does not depend on the order of the atomics -- integer work, deterministic, bit-exact against the
// torch restatement in tests/partition_ref.py.  The capacity-constrained acceptance of the proposed moves (a sort by
// (target part, gain, id)) is host-side orchestration of device sorts in partition_gpu.py.
#include "common.cuh"

namespace lgcn {

constexpr int VOTE_WARPS = 8;

// want[v]  = label with the most out-neighbours of v (v's own label if it has none)
// best[v]  = that count;  own[v] = out-neighbours carrying v's current label      (v in [row_begin, row_end))
__global__ void __launch_bounds__(VOTE_WARPS * 32)
label_vote_kernel(const int32_t *__restrict__ ptr, const int32_t *__restrict__ nbr, const int32_t *__restrict__ labels,
                  int row_begin, int row_end, int P, int32_t *__restrict__ want, int32_t *__restrict__ best,
                  int32_t *__restrict__ own) {
    extern __shared__ int hist_all[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int *hist = hist_all + wid * P;
    for (int v = row_begin + blockIdx.x * VOTE_WARPS + wid; v < row_end; v += gridDim.x * VOTE_WARPS) {
        for (int p = lane; p < P; p += 32) hist[p] = 0;
        __syncwarp();
        const int b = __ldg(ptr + v), e = __ldg(ptr + v + 1);
        for (int i = b + lane; i < e; i += 32) atomicAdd(hist + __ldg(labels + __ldg(nbr + i)), 1);
        __syncwarp();
        const int cur = __ldg(labels + v);
        // arg-max under (count desc, label asc): key = count * 2^20 + (2^20 - 1 - label)
        long long key = -1;
        for (int p = lane; p < P; p += 32) {
            const long long k = ((long long)hist[p] << 20) | (long long)((1 << 20) - 1 - p);
            key = k > key ? k : key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const long long other = __shfl_xor_sync(FULL, key, o);
            key = other > key ? other : key;
        }
        if (lane == 0) {
            const int cnt = (int)(key >> 20), lab = (1 << 20) - 1 - (int)(key & ((1 << 20) - 1));
            want[v] = e > b ? lab : cur;
            best[v] = e > b ? cnt : 0;
            own[v] = hist[cur];
        }
        __syncwarp();
    }
}

}  // namespace lgcn

extern "C" int lgcn_label_vote(const int32_t *ptr, const int32_t *nbr, const int32_t *labels, int64_t row_begin,
                               int64_t row_end, int num_parts, int32_t *want, int32_t *best, int32_t *own, void *stream) {
    using namespace lgcn;
    LGCN_REQUIRE(ptr && labels && want && best && own, LGCN_E_INVALID, "label_vote: null argument");
    LGCN_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end < INT32_MAX, LGCN_E_INVALID, "label_vote: bad row range");
    LGCN_REQUIRE(num_parts >= 1 && num_parts <= 4096, LGCN_E_INVALID, "label_vote: num_parts %d outside [1,4096]", num_parts);
    const int64_t rows = row_end - row_begin;
    if (rows == 0) return LGCN_OK;
    const size_t smem = sizeof(int) * (size_t)VOTE_WARPS * (size_t)num_parts;
    if (smem > 48 * 1024)
        LGCN_CUDA(cudaFuncSetAttribute(label_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = (rows + VOTE_WARPS - 1) / VOTE_WARPS;
    if (grid > 148 * 16) grid = 148 * 16;
    label_vote_kernel<<<(int)grid, VOTE_WARPS * 32, smem, (cudaStream_t)stream>>>(ptr, nbr, labels, (int)row_begin,
                                                                               (int)row_end, num_parts, want, best, own);
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}
