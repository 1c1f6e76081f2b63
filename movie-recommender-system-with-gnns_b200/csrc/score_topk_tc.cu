// K5 on the 5th-generation tensor cores: user x item scoring as a tcgen05 TF32 GEMM with the
// accumulator in TMEM, fused with train-edge masking and a per-user top-k, so that the score matrix
// (9.6 G scores at ML-25M) never leaves the SM.
//
// Replaces (same contract as score_topk.cu, the FFMA version):
//   utils/recommend.py:39-61      normalise rows, matmul, sort, skip excluded, first 10
//   utils/train_test.py:191-197   torch.mm(user_normalized, items.t()) + topk(k)
//
// fp32 parity on TF32 tensor cores: every fp32 operand x is split as x = hi + lo with hi = x with
// the 13 low mantissa bits cleared (exactly a TF32 number) and lo = x - hi (exact in fp32, <= 13
// significant bits, of which the MMA keeps 11).  score = hi.hi + hi.lo + lo.hi accumulated in fp32
// in TMEM; the dropped lo.lo term and the truncation of lo are O(2^-21) relative -- well inside the
// 1e-5 budget (measured in tests/test_gpu_cluster_score.py).  Cost: 3 MMAs per K-step.
//
// CTA = 128 users (TMEM lanes) x all items, streamed as 64-item tiles:
//   warps 0-3  epilogue (packed variant: warps 0-7, two groups that split each tile's 64 columns 32 / 32 and
//              merge their sorted lists at the end): thread t owns user t: tcgen05.ld of its 64 scores, compare with an admission
//              threshold kept in a register, mask train items with a 64-bit mask built from a cursor
//              into the user's sorted exclusion row, APPEND survivors to the user's candidate buffer
//              (shared memory, column per user => conflict-free).  When a buffer runs full the warp
//              prunes it together: every lane ranks up to three entries against all of them (broadcast
//              reads), the k best are written back in rank order and the k-th becomes the new threshold.
//              A user is pruned ~6 times over 59 k items instead of paying ~100 serial sorted insertions
//              (ncu: those were 43 % of the epilogue warps' samples and made them the bottleneck).
//   warp  4    one elected thread issues 24 tcgen05.mma (8 K-steps x 3 split terms, M=128 N=64 K=8)
//              per tile and commits to mbarriers
//   warps 5-12 producers (two groups of four, one per stage): gather 64 item rows, L2-normalise, split
//              hi/lo, store in the UMMA K-major no-swizzle ("interleaved") core-matrix layout,
//              fence.proxy.async, arrive
// Two shared-memory stages for B and two TMEM accumulator stages (2 x 64 columns) decouple the roles.
//
// PACKED variant (lgcn_score_topk_ex with a workspace): a pre-pass normalises and splits the item table
// ONCE into the shared-memory image of every 64-item tile (hi tile, lo tile: 32 KB, back to back in global
// memory), and the producer is a single thread issuing two 16 KB TMA bulk copies per tile
// (cp.async.bulk ... mbarrier::complete_tx) -- instead of every CTA re-normalising all items with eight
// warps whose dependent L2 gathers set a floor of ~1900 cycles per tile (round-1 ncu / timing experiments).
#include "common.cuh"
#include <limits.h>
#include <stdint.h>
#include <math_constants.h>

namespace lgcn {
namespace tc {

constexpr int BM = 128, BN = 64, BK = 64;
constexpr int GROUP_BYTES = 2048;                 // 8 rows x 256 B: 16 core matrices of 8 x 16 B
constexpr int A_BYTES = (BM / 8) * GROUP_BYTES;   // 32 KB per split half
constexpr int B_BYTES = (BN / 8) * GROUP_BYTES;   // 16 KB per split half
constexpr int NUM_EPI = 128, NUM_PROD = 128;
constexpr int PROD_GROUPS = 2;                    // producer group g fills stage g for tiles t = g (mod 2)
constexpr int THREADS = NUM_EPI + 32 + PROD_GROUPS * NUM_PROD;  // 416
constexpr int THREADS_PACKED = 2 * NUM_EPI + 32 + 32;  // 320: two epilogue warpgroups, MMA issuer, TMA issuer
constexpr int TMEM_COLS = 128;                    // 2 accumulator stages x 64 fp32 columns
constexpr int TMEM_COLS_PACKED = 512;             // 4 accumulator stages (columns 0..255) + the user tile: hi at 256, lo at 320
constexpr int KMAX = 32;                          // k <= 32

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// K-major, no swizzle: start address, LBO = 128 B between the two 16-byte K chunks of one MMA,
// SBO = 2048 B between 8-row groups; descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((128 >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((GROUP_BYTES >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// kind::tf32, D = F32, A/B = TF32 K-major, N = 64, M = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate) : "memory");
}
// A operand from tensor memory: row m of the 128 x 8 tf32 slice in lane m, one 32-bit column per element
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(IDESC), "r"(accumulate) : "memory");
}
// Same with the accumulate flag known at compile time (no setp per MMA: the single issuing thread's instruction
// stream is on the critical path -- 24 MMAs per tile with descriptor arithmetic cost ~1800 cycles, round 1).
template <bool kAcc>
__device__ __forceinline__ void umma_tf32_ts_imm(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc) {
    if constexpr (kAcc)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(IDESC) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(IDESC) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[N]);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld_cols<64>(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
          "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
          "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
}

__device__ __forceinline__ float4 split_hi(const float4 &v) {
    return make_float4(__uint_as_float(__float_as_uint(v.x) & 0xffffe000u), __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                       __uint_as_float(__float_as_uint(v.z) & 0xffffe000u), __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
}

// 8 rows x 256 B per warp pass: lane = (row r8 = lane % 8, chunk quarter c4 = lane / 8); each lane
// loads the 4 chunks c = cb*4 + c4 of its row, the 4 lanes of a row reduce the squared norm.
__device__ __forceinline__ void load_rows8(const float *__restrict__ tab, int64_t row0, int64_t nrows, bool normalize,
                                           unsigned char *hi_base, unsigned char *lo_base, int group, int lane) {
    const int r8 = lane & 7, c4 = lane >> 3;
    const int64_t row = row0 + group * 8 + r8;
    float4 v[4];
    float n2 = 0.f;
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
        v[cb] = row < nrows ? ldg4(reinterpret_cast<const float4 *>(tab) + row * D4 + cb * 4 + c4) : f4zero();
        n2 += f4dot(v[cb], v[cb]);
    }
    n2 += __shfl_xor_sync(FULL, n2, 8);
    n2 += __shfl_xor_sync(FULL, n2, 16);
    const float inv = normalize ? 1.0f / sqrtf(n2) : 1.0f;
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
        float4 x = row < nrows ? f4scale(inv, v[cb]) : f4zero();
        const float4 h = split_hi(x);
        const float4 l = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
        const int off = group * GROUP_BYTES + (cb * 4 + c4) * 128 + r8 * 16;
        *reinterpret_cast<float4 *>(hi_base + off) = h;
        *reinterpret_cast<float4 *>(lo_base + off) = l;
    }
}

// Producer-warp variant: the user tile (hi / lo) and two item stages in shared memory.
// Packed variant: the STATIONARY user tile lives in TMEM (128 columns: hi, lo) -- every MMA then reads only its
// 2 KB item operand from shared memory instead of 6 KB -- and the freed 64 KB hold two more item stages, which
// is what covers the latency of the TMA copies (two stages bounded a tile at ~1800 cycles, round-1 experiments).
template <bool kPacked>
struct __align__(16) SmemT {
    static constexpr int NSTAGE = 2, ASTAGE = 2, EPI_GROUPS = 1, CANDN = 96;
    unsigned char a_hi[A_BYTES], a_lo[A_BYTES];
    unsigned char b[NSTAGE][2 * B_BYTES];             // hi tile, lo tile
    float cand_v[EPI_GROUPS][CANDN][BM];              // candidate buffers: a column per user => conflict-free
    int cand_i[EPI_GROUPS][CANDN][BM];
    int cand_cnt[EPI_GROUPS][BM];
    uint64_t full[NSTAGE], empty[NSTAGE], tfull[ASTAGE], tempty[ASTAGE];
    uint32_t tmem_base;
};
template <>
struct __align__(16) SmemT<true> {
    // two epilogue warpgroups split every 64-column accumulator tile 32 / 32, each with its own buffers
    // four accumulator stages (256 TMEM columns) cover the MMA -> commit -> tcgen05.ld -> release round trip
    static constexpr int NSTAGE = 3, ASTAGE = 4, EPI_GROUPS = 2, CANDN = 64;
    unsigned char b[NSTAGE][2 * B_BYTES];
    float cand_v[EPI_GROUPS][CANDN][BM];
    int cand_i[EPI_GROUPS][CANDN][BM];
    int cand_cnt[EPI_GROUPS][BM];
    uint64_t full[NSTAGE], empty[NSTAGE], tfull[ASTAGE], tempty[ASTAGE];
    uint32_t tmem_base;
};

// The shared-memory image of item tile t (rows [64 t, 64 t + 64)): normalised, split, core-matrix layout.
__global__ void __launch_bounds__(256)
pack_items_kernel(const float *__restrict__ item_emb, int64_t num_items, int normalize, unsigned char *__restrict__ packed) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;      // warp = 8-row group of the tile
    unsigned char *tile = packed + (size_t)blockIdx.x * (2 * B_BYTES);
    load_rows8(item_emb, (int64_t)blockIdx.x * BN, num_items, normalize != 0, tile, tile + B_BYTES, warp, lane);
}

template <bool kPacked>
__global__ void __launch_bounds__(kPacked ? THREADS_PACKED : THREADS, 1)
score_topk_tc_kernel(const float *__restrict__ user_emb, const float *__restrict__ item_emb,
                     const unsigned char *__restrict__ packed, int64_t num_items,
                     int64_t u_begin, int64_t u_end, int normalize, const int64_t *__restrict__ excl_ptr,
                     const int32_t *__restrict__ excl_idx, int k, int32_t *__restrict__ topk_idx,
                     float *__restrict__ topk_val) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    using Smem = SmemT<kPacked>;
    constexpr int ASTAGE = Smem::ASTAGE;
    constexpr uint32_t A_HI_COL = ASTAGE * BN, A_LO_COL = ASTAGE * BN + BK;
    constexpr int NSTAGE = Smem::NSTAGE, EPI_GROUPS = Smem::EPI_GROUPS, CANDN = Smem::CANDN;
    constexpr int COLS = BN / EPI_GROUPS;       // accumulator columns per epilogue thread and tile
    constexpr int SUB = COLS / 2;               // columns scanned between two room checks
    constexpr int RPL = CANDN / 32;             // buffer entries per lane in a prune
    constexpr int MMA_WARP = 4 * EPI_GROUPS;    // warps [0, MMA_WARP): epilogue, then the MMA issuer, then producers / TMA
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t u0 = u_begin + (int64_t)blockIdx.x * BM;
    const int num_tiles = (int)((num_items + BN - 1) / BN);
    // Every CTA walks ALL item tiles; staggered starting points keep the 148 resident CTAs from asking the
    // same L2 lines for the same tile at the same moment (iteration i handles tile (tile0 + i) mod num_tiles).
    const int tile0 = (int)(((long long)blockIdx.x * 61) % num_tiles);
#define TILE_OF(i) ((tile0 + (i)) < num_tiles ? (tile0 + (i)) : (tile0 + (i)) - num_tiles)

    // ---- one-time setup ------------------------------------------------------------------------
    if constexpr (!kPacked) {
        for (int grp = warp; grp < BM / 8; grp += (int)(blockDim.x >> 5))
            load_rows8(user_emb, u0, u_end, normalize != 0, S.a_hi, S.a_lo, grp, lane);
    }
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&S.full[s], kPacked ? 1 : NUM_PROD);
            mbar_init(&S.empty[s], 1);
        }
        for (int s = 0; s < ASTAGE; ++s) {
            mbar_init(&S.tfull[s], 1);
            mbar_init(&S.tempty[s], MMA_WARP);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&S.tmem_base)), "n"(kPacked ? TMEM_COLS_PACKED : TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if constexpr (kPacked) {
        __syncthreads();                                               // tmem_base is visible
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp < 4) {
            // thread = user = TMEM lane: normalise the row, split it, store hi / lo as 64 + 64 columns
            const int64_t row = u0 + tid;
            const float4 *src = reinterpret_cast<const float4 *>(user_emb) + (size_t)(row < u_end ? row : u_begin) * D4;
            float n2 = 0.f;
#pragma unroll
            for (int c = 0; c < D4; ++c) { const float4 v = ldg4(src + c); n2 += f4dot(v, v); }
            const float inv = row < u_end ? (normalize ? 1.0f / sqrtf(n2) : 1.0f) : 0.f;
            const uint32_t lane_base = S.tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int c = 0; c < 4; ++c) {                              // 16 columns at a time
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 x = f4scale(inv, ldg4(src + c * 4 + q));
                    const float4 h = split_hi(x);
                    hi[4 * q + 0] = __float_as_uint(h.x); hi[4 * q + 1] = __float_as_uint(h.y);
                    hi[4 * q + 2] = __float_as_uint(h.z); hi[4 * q + 3] = __float_as_uint(h.w);
                    lo[4 * q + 0] = __float_as_uint(x.x - h.x); lo[4 * q + 1] = __float_as_uint(x.y - h.y);
                    lo[4 * q + 2] = __float_as_uint(x.z - h.z); lo[4 * q + 3] = __float_as_uint(x.w - h.w);
                }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                             :: "r"(lane_base + A_HI_COL + (uint32_t)(c * 16)), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]),
                                "r"(hi[5]), "r"(hi[6]), "r"(hi[7]), "r"(hi[8]), "r"(hi[9]), "r"(hi[10]), "r"(hi[11]), "r"(hi[12]),
                                "r"(hi[13]), "r"(hi[14]), "r"(hi[15]) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                             :: "r"(lane_base + A_LO_COL + (uint32_t)(c * 16)), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]),
                                "r"(lo[5]), "r"(lo[6]), "r"(lo[7]), "r"(lo[8]), "r"(lo[9]), "r"(lo[10]), "r"(lo[11]), "r"(lo[12]),
                                "r"(lo[13]), "r"(lo[14]), "r"(lo[15]) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // A tile (generic stores) -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;

    if (kPacked && warp > MMA_WARP) {
        // ===== TMA issuer: one 32 KB bulk copy per tile, completion counted in bytes on full[s] =====
        if (warp == MMA_WARP + 1 && lane == 0) {
            for (int t = 0; t < num_tiles; ++t) {
                const int s = t % NSTAGE;
                mbar_wait(&S.empty[s], ((t / NSTAGE) & 1) ^ 1);
                const unsigned char *src = packed + (size_t)TILE_OF(t) * (2 * B_BYTES);
                const uint32_t bar = smem_u32(&S.full[s]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(2 * B_BYTES) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(S.b[s])), "l"(src), "r"(2 * B_BYTES), "r"(bar) : "memory");
            }
        }
    } else if (warp > MMA_WARP) {
        // ===== producers: item tiles -> shared memory (hi / lo halves) ============================
        // two groups of four warps, one per shared-memory stage: a group has two tile periods to cover the
        // L2 latency of its row gathers (with one group the loads were exposed: ncu, round 1)
        const int pw = (warp - MMA_WARP - 1) & 3, grp = (warp - MMA_WARP - 1) >> 2;
        for (int t = grp; t < num_tiles; t += PROD_GROUPS) {
            const int s = t & 1;
            mbar_wait(&S.empty[s], ((t >> 1) & 1) ^ 1);
            const int64_t n0 = (int64_t)TILE_OF(t) * BN;
#pragma unroll
            for (int it = 0; it < 2; ++it)
                load_rows8(item_emb, n0, num_items, normalize != 0, S.b[s], S.b[s] + B_BYTES, pw * 2 + it, lane);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&S.full[s]);
        }
    } else if (warp == MMA_WARP) {
        // ===== MMA issuer ===========================================================================
        if (lane == 0) {
            for (int t = 0; t < num_tiles; ++t) {
                const int s = t % NSTAGE, as = t % ASTAGE;      // item stage, accumulator stage
                mbar_wait(&S.full[s], (t / NSTAGE) & 1);
                mbar_wait(&S.tempty[as], ((t / ASTAGE) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem + (uint32_t)(as * BN);
                const uint32_t b_hi = smem_u32(S.b[s]), b_lo = b_hi + B_BYTES;
                // one descriptor per operand tile; a K-step advances its 14-bit start-address field by 256 B >> 4
                const uint64_t dhi = umma_desc(b_hi), dlo = umma_desc(b_lo);
#pragma unroll
                for (int j = 0; j < BK / 8; ++j) {             // one MMA K-step = 8 tf32 = two 16-byte chunks
                    const uint32_t ko = (uint32_t)j * 256u;
                    if constexpr (kPacked) {
                        const uint32_t a_hi = tmem + A_HI_COL + (uint32_t)(j * 8), a_lo = tmem + A_LO_COL + (uint32_t)(j * 8);
                        if (j == 0) umma_tf32_ts_imm<false>(d, a_hi, dhi);
                        else umma_tf32_ts_imm<true>(d, a_hi, dhi + (uint64_t)(j * 16));
                        umma_tf32_ts_imm<true>(d, a_hi, dlo + (uint64_t)(j * 16));
                        umma_tf32_ts_imm<true>(d, a_lo, dhi + (uint64_t)(j * 16));
                    } else {
                        const uint32_t a_hi = smem_u32(S.a_hi), a_lo = smem_u32(S.a_lo);
                        umma_tf32(d, umma_desc(a_hi + ko), umma_desc(b_hi + ko), j > 0 ? 1u : 0u);
                        umma_tf32(d, umma_desc(a_hi + ko), umma_desc(b_lo + ko), 1u);
                        umma_tf32(d, umma_desc(a_lo + ko), umma_desc(b_hi + ko), 1u);
                    }
                }
                umma_commit(&S.empty[s]);                      // item stage free once these MMAs have read it
                umma_commit(&S.tfull[as]);                     // accumulator stage ready for the epilogue
            }
        }
    } else {
        // ===== epilogue: thread = (user = TMEM lane, column group) =================================
        const int eg = warp >> 2, quarter = warp & 3;           // column group, TMEM lane quarter of this warp
        const int m = quarter * 32 + lane;                      // user row inside the CTA
        const bool live = u0 + m < u_end;
        float (*cv)[BM] = S.cand_v[eg];
        int (*ci)[BM] = S.cand_i[eg];
        int64_t ex_cur = 0, ex_end = 0, ex_begin = 0;
        if (live && excl_ptr) {
            ex_begin = excl_ptr[u0 + m];
            ex_end = excl_ptr[u0 + m + 1];
            int64_t lo = ex_begin, hi = ex_end;                // first excluded item at or after the first tile
            const int64_t first = (int64_t)tile0 * BN;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (excl_idx[mid] < first) lo = mid + 1; else hi = mid;
            }
            ex_cur = lo;
        }
        // next excluded item kept in a register: a tile without one costs a compare, not a global load
        int64_t ex_next = ex_cur < ex_end ? (int64_t)excl_idx[ex_cur] : INT64_MAX;
        // admission: a score enters the buffer iff it orders before (thr, thr_id) under (score desc, id asc);
        // at least k kept candidates do.  thr_id = INT_MAX admits every score equal to thr.
        float thr = -CUDART_INF_F;
        int thr_id = INT_MAX;
        int cnt = 0;                                           // candidates in the buffer

        // EXACT prune (once per user, at the end): all 32 lanes rank up to RPL entries each against every
        // entry under (score desc, id asc); the k best move to their rank, so the buffer comes out sorted.
        auto prune_exact = [&](int src) {
            const int um = quarter * 32 + src;
            const int n = __shfl_sync(FULL, cnt, src);
            float mv[RPL];
            int mi[RPL], rank[RPL];
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int pos = lane + 32 * r;
                mv[r] = pos < n ? cv[pos][um] : -CUDART_INF_F;
                mi[r] = pos < n ? ci[pos][um] : INT_MAX;
                rank[r] = 0;
            }
            for (int t = 0; t < n; ++t) {
                const float v = cv[t][um];                     // same address in every lane: broadcast
                const int i = ci[t][um];
#pragma unroll
                for (int r = 0; r < RPL; ++r) rank[r] += (v > mv[r] || (v == mv[r] && i < mi[r])) ? 1 : 0;
            }
            __syncwarp();
#pragma unroll
            for (int r = 0; r < RPL; ++r)
                if (lane + 32 * r < n && rank[r] < k) { cv[rank[r]][um] = mv[r]; ci[rank[r]][um] = mi[r]; }
            __syncwarp();
            if (lane == src) {
                cnt = n < k ? n : k;
                if (n >= k) { thr = cv[k - 1][um]; thr_id = ci[k - 1][um]; }
            }
        };

        // CHEAP prune (whenever a buffer runs full, ~8 times per user over 59 k items): bisect, on the
        // order-preserving integer image of the scores, for a pivot that at least k and at most k + 8 kept
        // entries reach (or as close as ties allow); entries below it can never be in the final top-k and
        // are dropped, the pivot becomes the admission threshold.  ~15 warp instructions per bisection step
        // instead of a full ranking.
        auto prune = [&](int src) {
            const int um = quarter * 32 + src;
            const int n = __shfl_sync(FULL, cnt, src);
            const float told = __shfl_sync(FULL, thr, src);
            float mv[RPL];
            int mi[RPL];
            unsigned key[RPL], kmax = 0u;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int pos = lane + 32 * r;
                mv[r] = pos < n ? cv[pos][um] : -CUDART_INF_F;
                mi[r] = pos < n ? ci[pos][um] : INT_MAX;
                const unsigned bits = __float_as_uint(mv[r]);
                key[r] = pos < n ? ((bits & 0x80000000u) ? ~bits : (bits | 0x80000000u)) : 0u;
                kmax = max(kmax, key[r]);
            }
            kmax = __reduce_max_sync(FULL, kmax);
            const unsigned tb = __float_as_uint(told);
            unsigned lo = told == -CUDART_INF_F ? 1u : ((tb & 0x80000000u) ? ~tb : (tb | 0x80000000u));   // every entry reaches lo
            unsigned hi = kmax + 1u;                           // no entry reaches hi (kmax < 2^32 - 1 for finite scores)
            if (hi == 0u) hi = 0xffffffffu;
            for (int it = 0; it < 40 && hi - lo > 1u; ++it) {
                const unsigned mid = lo + ((hi - lo) >> 1);
                int c = 0;
#pragma unroll
                for (int r = 0; r < RPL; ++r) c += __popc(__ballot_sync(FULL, key[r] >= mid));
                if (c >= k) {
                    lo = mid;
                    if (c <= k + 8) break;
                } else {
                    hi = mid;
                }
            }
            __syncwarp();
            int base = 0;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const bool keep = key[r] >= lo && key[r] != 0u;
                const unsigned bal = __ballot_sync(FULL, keep);
                if (keep) {
                    const int pos = base + __popc(bal & ((1u << lane) - 1u));
                    cv[pos][um] = mv[r];
                    ci[pos][um] = mi[r];
                }
                base += __popc(bal);
            }
            __syncwarp();
            if (lane == src) {
                cnt = base;
                const unsigned pb = (lo & 0x80000000u) ? (lo & 0x7fffffffu) : ~lo;
                thr = lo <= 1u ? -CUDART_INF_F : __uint_as_float(pb);
                thr_id = INT_MAX;
            }
            return base;
        };

        for (int t = 0; t < num_tiles; ++t) {
            const int s = t % ASTAGE;
            mbar_wait(&S.tfull[s], (t / ASTAGE) & 1);
            __syncwarp();                                       // reconverge: tcgen05.ld is .sync.aligned
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[COLS];
            tmem_ld_cols<COLS>(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * BN + eg * COLS), r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.tempty[s]);           // this warp is done with the accumulator stage
            const int tile = TILE_OF(t);
            const int64_t n0 = (int64_t)tile * BN;
            if (tile == 0 && t > 0) {                           // wrapped around: back to the head of the exclusion row
                ex_cur = ex_begin;
                ex_next = ex_cur < ex_end ? (int64_t)excl_idx[ex_cur] : INT64_MAX;
            }
            unsigned long long mask = 0ull;                     // train items of this user inside the tile
            while (ex_next < n0 + BN) {
                if (ex_next >= n0) mask |= 1ull << (int)(ex_next - n0);
                ++ex_cur;
                ex_next = ex_cur < ex_end ? (int64_t)excl_idx[ex_cur] : INT64_MAX;
            }
            const int valid = (int)min((int64_t)BN, num_items - n0);
            const unsigned long long adm64 = live ? (valid >= 64 ? ~0ull : ((1ull << valid) - 1ull)) & ~mask : 0ull;
            const unsigned long long admissible = adm64 >> (eg * COLS);          // bit c = this thread's column c
            const int64_t nc0 = n0 + eg * COLS;
#pragma unroll
            for (int h = 0; h < COLS / SUB; ++h) {
                // room for SUB more candidates, or prune first (warp-uniform decision)
                unsigned need = __ballot_sync(FULL, cnt > CANDN - SUB);
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    if (prune(src) > CANDN - SUB) prune_exact(src);   // too many exact ties at the pivot
                }
                // fast path: the maximum of each 8-column group against the threshold (a max tree has no
                // loop-carried dependence; a lone epilogue warp per scheduler is latency-, not issue-bound)
                float gm[SUB / 8];
                float hm = -CUDART_INF_F;
#pragma unroll
                for (int g = 0; g < SUB / 8; ++g) {
                    const int c0 = SUB * h + 8 * g;
                    const float a0 = fmaxf(__uint_as_float(r[c0 + 0]), __uint_as_float(r[c0 + 1]));
                    const float a1 = fmaxf(__uint_as_float(r[c0 + 2]), __uint_as_float(r[c0 + 3]));
                    const float a2 = fmaxf(__uint_as_float(r[c0 + 4]), __uint_as_float(r[c0 + 5]));
                    const float a3 = fmaxf(__uint_as_float(r[c0 + 6]), __uint_as_float(r[c0 + 7]));
                    gm[g] = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
                    hm = fmaxf(hm, gm[g]);
                }
                if (hm >= thr) {
#pragma unroll
                    for (int g = 0; g < SUB / 8; ++g) {
                        if (gm[g] >= thr) {
                            // which of the 8 columns pass: a bit per column, then only the set bits are visited
                            // (usually one lane, one bit) instead of eight predicated bodies
                            const int c0 = SUB * h + 8 * g;
                            unsigned pass = 0u;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                pass |= (__uint_as_float(r[c0 + j]) >= thr ? 1u : 0u) << j;
                            pass &= (unsigned)(admissible >> c0) & 0xffu;
                            while (pass) {
                                const int j = __ffs(pass) - 1;
                                pass &= pass - 1;
                                float sc = 0.f;
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    if (q == j) sc = __uint_as_float(r[c0 + q]);          // registers cannot be indexed
                                sc += 0.0f;                                              // -0 -> +0: one image per value
                                const int id = (int)(nc0 + c0 + j);
                                if (sc > thr || id < thr_id) {
                                    cv[cnt][m] = sc;
                                    ci[cnt][m] = id;
                                    ++cnt;
                                }
                            }
                        }
                    }
                }
            }
        }
        // final exact prune of every buffer -> sorted top-min(cnt,k)
        {
            unsigned need = __ballot_sync(FULL, live);
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1;
                prune_exact(src);
            }
        }
        if constexpr (EPI_GROUPS == 2) {
            // the two column groups of a user hold sorted top-k lists of disjoint item sets: merge
            S.cand_cnt[eg][m] = cnt;
            asm volatile("bar.sync 1, %0;" :: "n"(32 * MMA_WARP) : "memory");          // epilogue warps only
            if (live && eg == 0) {
                const int64_t out = (u0 - u_begin + m) * (int64_t)k;
                const int na = cnt, nb = S.cand_cnt[1][m];
                int ia = 0, ib = 0;
                for (int e = 0; e < k; ++e) {
                    const bool ha = ia < na, hb = ib < nb;
                    const float va = ha ? S.cand_v[0][ia][m] : 0.f, vb = hb ? S.cand_v[1][ib][m] : 0.f;
                    const int xa = ha ? S.cand_i[0][ia][m] : 0, xb = hb ? S.cand_i[1][ib][m] : 0;
                    const bool take_a = ha && (!hb || va > vb || (va == vb && xa < xb));
                    if (take_a) { topk_val[out + e] = va; topk_idx[out + e] = xa; ++ia; }
                    else if (hb) { topk_val[out + e] = vb; topk_idx[out + e] = xb; ++ib; }
                    else { topk_val[out + e] = -CUDART_INF_F; topk_idx[out + e] = -1; }
                }
            }
        } else if (live) {
            const int64_t out = (u0 - u_begin + m) * (int64_t)k;
            for (int e = 0; e < k; ++e) {
                const bool have = e < cnt;
                topk_val[out + e] = have ? cv[e][m] : -CUDART_INF_F;
                topk_idx[out + e] = have ? ci[e][m] : -1;
            }
        }
    }
#undef TILE_OF
    // ---- teardown ------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(kPacked ? TMEM_COLS_PACKED : TMEM_COLS) : "memory");
    }
}

}  // namespace tc

int score_topk_tc_impl(const float *ue, const float *ie, int64_t I, int64_t ub, int64_t uend, int normalize,
                       const int64_t *ep, const int32_t *ex, int k, int32_t *ti, float *tv, void *workspace,
                       size_t workspace_bytes, cudaStream_t st) {
    const size_t smem = (workspace ? sizeof(tc::SmemT<true>) : sizeof(tc::SmemT<false>)) + 1024;
    const int grid = cdiv(uend - ub, tc::BM);
    const int64_t tiles = (I + tc::BN - 1) / tc::BN;
    if (workspace) {
        LGCN_REQUIRE(workspace_bytes >= (size_t)tiles * 2 * tc::B_BYTES && ((uintptr_t)workspace & 127) == 0, LGCN_E_WORKSPACE,
                     "score_topk: workspace %zu < %zu bytes (or not 128-byte aligned)", workspace_bytes,
                     (size_t)tiles * 2 * tc::B_BYTES);
        tc::pack_items_kernel<<<(unsigned)tiles, 256, 0, st>>>(ie, I, normalize, (unsigned char *)workspace);
        LGCN_LAUNCH_CHECK();
        LGCN_CUDA(cudaFuncSetAttribute(tc::score_topk_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::score_topk_tc_kernel<true><<<grid, tc::THREADS_PACKED, smem, st>>>(ue, ie, (const unsigned char *)workspace, I, ub, uend,
                                                                                normalize, ep, ex, k, ti, tv);
    } else {
        LGCN_CUDA(cudaFuncSetAttribute(tc::score_topk_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::score_topk_tc_kernel<false><<<grid, tc::THREADS, smem, st>>>(ue, ie, nullptr, I, ub, uend, normalize, ep, ex, k, ti, tv);
    }
    LGCN_LAUNCH_CHECK();
    return LGCN_OK;
}

size_t score_topk_tc_workspace_bytes(int64_t I) {
    return (size_t)((I + tc::BN - 1) / tc::BN) * 2 * tc::B_BYTES;
}

}  // namespace lgcn
