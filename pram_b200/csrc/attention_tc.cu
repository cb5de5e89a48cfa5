// tcgen05 flash attention for sm_100a (head dim 64): out = softmax(Q K^T * scale) V without ever
// materialising the Nq x Nk matrix (the reference writes S as [B,4,N,N] fp32: 16.8 MB per layer at
// N=1024, 268 MB at N=4096 -- nets/segnetvit.py:73-76, nets/gml.py:104-107, 175-181).
//
// persistent CTAs, work item = (batch*head, 128-query tile); per 128-key tile:
//   warp 0  : TMA producer   Q {64 x 128}, K {64 x 128}, V^T {2 x (64 keys x 64 d)} -> SWIZZLE_128B smem
//   warp 1  : MMA issuer     S = Q K^T  (SS, fp32 in TMEM, double buffered)
//                            O += P V   (TS: P read from TMEM as the A operand, V^T from smem)
//   warps 2-17: softmax      one thread per query row and a quarter of the keys: tcgen05.ld S, running max with lazy
//                            rescale of O (only when the max moved by > 8 in log2 units), exp2,
//                            P packed to bf16 and written back to TMEM (over S, in place) with tcgen05.st; epilogue O / l
// SPLIT=3 keeps ~fp32 accuracy with bf16 tensor-core operands: Q,K,V and P are hi/lo split and each
// product is three MMAs (hi*hi + lo*hi + hi*lo) into the same fp32 accumulator.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

namespace fa {

constexpr int BQ = 128, HD = 64;

// Geometry by key-tile size.  BKV = 128: one CTA per SM, 16 softmax warps.  BKV = 64: half the shared memory and TMEM
// per CTA, 8 softmax warps, TWO CTAs per SM -- the two CTAs run out of phase, so the per-tile handshake chain of one
// (TMEM load -> row max -> exchange -> exp -> TMEM store -> mbarrier round trip) overlaps the MMAs of the other.
template <int BKV>
struct Geo {
    static constexpr int NPART = BKV / 32;             // softmax warps per TMEM lane quarter: each takes 32 key columns
    static constexpr int CPT = 32;                     // key columns per thread
    static constexpr int OPT = HD / NPART;             // O columns per thread (rescale + epilogue): 16 or 32
    static constexpr int NUM_SM_WARPS = 4 * NPART;
    static constexpr int NUM_THREADS = 64 + 32 * NUM_SM_WARPS;  // TMA warp, MMA warp, softmax warps
    // TMEM: two S buffers (BKV fp32 columns each) and O.  P(g) is written IN PLACE over S(g) (bf16 hi plane in columns
    // [0, BKV/2) of the buffer, lo plane in [BKV/2, BKV)) once every warp of the lane quarter holds its scores in
    // registers, so P is double-buffered for free: softmax(g+1) never waits for PV(g), and the tensor pipe's in-order
    // execution (PV(g) is issued before S(g+2)) is the only "buffer free" signal an S buffer needs.
    static constexpr uint32_t COL_S0 = 0, COL_S1 = BKV, COL_O = 2 * BKV;
    static constexpr uint32_t TMEM_COLS = (BKV == 128) ? 512 : 256;
    static constexpr int CTAS_PER_SM = (BKV == 128) ? 1 : 2;
};

struct Args {
    int BH, heads, Nq, Nk;
    float scale_log2;  // softmax scale * log2(e)
    float* out_f32; __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; int out_ld;
    int v_mn;          // 1: V given as [BH][Nk][64] (MN-major B operand), 0: V^T [BH][64][nk_pad] (K-major)
    const int* nk_counts;  // optional [B]: valid keys of batch element b (keys >= nk_counts[b] are padding and masked)
    int kv_shift;          // keys / values (and nk_counts) of (batch, head) index bh come from index (bh + kv_shift) mod BH:
                           // B/2 * heads turns ONE launch over [set 0 | set 1] into both directions of the cross attention
    // MODE 0, optional: log2-domain log-sum-exp of every query row, [BH][ld_lse] (softmax prob = exp2(s * scale_log2 - lse))
    float* lse_out;
    // MODE 1 (column sums of the attention matrix, AdaGML's per-token mean attention, nets/adagml.py:148, 229): the kernel's
    // ROWS are the keys, its streamed COLUMNS the queries; lse_in [BH][ld_lse] = the row statistics of the queries written
    // by the MODE 0 launch, colsum [BH][ld_colsum] receives sum_i softmax(i, key) over the valid queries
    const float* lse_in; int ld_lse;
    float* colsum; int ld_colsum;
    const int* pred;   // launch predicate (common.cuh)
    int pdl_early;    // 1: release the dependent launch right after this grid's own wait (common.cuh)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait suspends the thread for a hardware-bounded time per call; the spin counter (2 instructions per poll
    // instead of a 64-bit clock comparison) turns a protocol bug into a trap after seconds instead of a hang
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) {
            printf("pram attention_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major SWIZZLE_128B, see gemm_tc.cu
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major SWIZZLE_128B operand (rows = K index, 64 MN elements = 128 B contiguous per row):
// SBO = 1024 B between 8-row K groups, LBO = stride between 64-element MN blocks (single block here)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(16384 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ldN(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_ldN(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
__device__ __forceinline__ void tmem_stN(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st16(taddr, r); }
__device__ __forceinline__ void tmem_stN(uint32_t taddr, const uint32_t (&r)[32]) { tmem_st32(taddr, r); }
// explicit shared-space accessors for the row max / row sum exchange (the pointer is derived from the manually aligned
// dynamic shared-memory base, so the compiler would otherwise emit generic LD / ST)
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int SPLIT, int BKV>
struct Cfg {
    static constexpr int NPL = (SPLIT == 3) ? 2 : 1;
    static constexpr int Q_BYTES = BQ * HD * 2;           // 16 KB per plane
    static constexpr int K_BYTES = BKV * HD * 2;          // per plane
    static constexpr int V_BYTES = BKV * HD * 2;          // per plane (V^T form: BKV / 64 boxes of 64 d x 64 keys)
    static constexpr int VBOX_BYTES = 64 * HD * 2;
    // K and V live in separate rings: S runs two tiles ahead of PV, so K(g+2) is needed while V(g) / V(g+1) are still
    // being consumed
    static constexpr int K_STAGES = (BKV == 128) ? 3 : 2, V_STAGES = 2;
    static constexpr int K_STAGE = NPL * K_BYTES, V_STAGE = NPL * V_BYTES;
    static constexpr int SMEM_BYTES = NPL * Q_BYTES + K_STAGES * K_STAGE + V_STAGES * V_STAGE + 1024 + 256 +
                                      2 * Geo<BKV>::NPART * 128 * 4 /*row max/sum exchange*/;
    static_assert(SMEM_BYTES <= (BKV == 128 ? 227 : 113) * 1024, "dynamic shared memory budget of sm_100a exceeded");
};

// P16: the probabilities go back to TMEM as ONE plane of IEEE fp16 (11-bit mantissas; P lies in [0, 1], far inside fp16's
// range) instead of bf16 hi / lo planes, and PV is P16 . V_hi + P16 . V_lo with V given as fp16 hi / lo planes (the
// hardware traps on an fp16 A operand against a bf16 B operand, so the producer of V -- the qkv GEMM epilogue -- writes
// fp16 planes for this mode).  The row sum is accumulated from the ROUNDED probabilities, so O / l
// is an exact convex combination of the V rows with weights that are off by <= 2^-12 relative -- the error that matters
// for a softmax average -- while the softmax warps execute about half the instructions per score and PV needs two MMAs
// per k-step instead of three.
template <int SPLIT, int BKV, int MODE, int P16>
__global__ void __launch_bounds__(Geo<BKV>::NUM_THREADS, Geo<BKV>::CTAS_PER_SM) attention_tc_kernel(
    const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
    const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo,
    const __grid_constant__ CUtensorMap map_v_hi, const __grid_constant__ CUtensorMap map_v_lo, const Args p) {
    using C = Cfg<SPLIT, BKV>;
    using Gm = Geo<BKV>;
    constexpr int NPART = Gm::NPART, CPT = Gm::CPT, OPT = Gm::OPT, NUM_SM_WARPS = Gm::NUM_SM_WARPS;
    constexpr uint32_t COL_S0 = Gm::COL_S0, COL_S1 = Gm::COL_S1, COL_O = Gm::COL_O, TMEM_COLS = Gm::TMEM_COLS;
    if (p.pred) { pram_pdl_wait(); if (pram_pred_skip(p.pred)) return; }  // the flag is written by a predecessor kernel
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* q_s = smem;                                  // [NPL][16 KB]
    uint8_t* k_s = smem + C::NPL * C::Q_BYTES;            // [K_STAGES][planes]
    uint8_t* v_s = k_s + C::K_STAGES * C::K_STAGE;        // [V_STAGES][planes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + C::V_STAGES * C::V_STAGE);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;    // [3]
    uint64_t* k_empty = bars + 5;   // [3]
    uint64_t* v_full = bars + 8;    // [2]
    uint64_t* v_empty = bars + 10;  // [2]
    uint64_t* s_full = bars + 12;   // [2]  S(g) accumulated
    uint64_t* p_full = bars + 14;   // [2]  P(g) stored by all softmax warps
    uint64_t* o_done = bars + 16;   // [2]  PV(g) retired
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tiles = (p.Nq + BQ - 1) / BQ;
    const int kv_tiles = (p.Nk + BKV - 1) / BKV;
    const int total = p.BH * q_tiles;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        for (int s = 0; s < 3; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1); mbar_init(&p_full[s], NUM_SM_WARPS); mbar_init(&o_done[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    pram_pdl_wait();     // programmatic dependent launch: everything above overlapped the predecessor's drain
    if (p.pdl_early) pram_pdl_trigger();  // all CTAs of a persistent grid are resident: the successor may be scheduled as SMs free up

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        // this CTA's tiles in issue order: t -> (local item t / kv_tiles, key tile t % kv_tiles)
        const int n_items = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int n_tiles = n_items * kv_tiles;
        auto item_of = [&](int li) { return (int)blockIdx.x + li * (int)gridDim.x; };
        auto load_q = [&](int li) {
            const int item = item_of(li), bh = item / q_tiles, q0 = (item % q_tiles) * BQ;
            mbar_wait(q_empty, (li & 1) ^ 1);
            mbar_expect_tx(q_full, C::NPL * C::Q_BYTES);
            tma_load_3d(q_s, &map_q_hi, q_full, 0, q0, bh);
            if (SPLIT == 3) tma_load_3d(q_s + C::Q_BYTES, &map_q_lo, q_full, 0, q0, bh);
        };
        auto kv_of = [&](int bh) { const int s = bh + p.kv_shift; return s >= p.BH ? s - p.BH : s; };
        auto load_k = [&](int t) {
            const int bh = kv_of(item_of(t / kv_tiles) / q_tiles), k0 = (t % kv_tiles) * BKV;
            const int st = t % C::K_STAGES;
            mbar_wait(&k_empty[st], ((t / C::K_STAGES) & 1) ^ 1);
            uint8_t* ks = k_s + st * C::K_STAGE;
            mbar_expect_tx(&k_full[st], C::K_STAGE);
            tma_load_3d(ks, &map_k_hi, &k_full[st], 0, k0, bh);
            if (SPLIT == 3) tma_load_3d(ks + C::K_BYTES, &map_k_lo, &k_full[st], 0, k0, bh);
        };
        auto load_v = [&](int t) {
            const int bh = kv_of(item_of(t / kv_tiles) / q_tiles), k0 = (t % kv_tiles) * BKV;
            const int st = t & 1;
            mbar_wait(&v_empty[st], ((t >> 1) & 1) ^ 1);
            uint8_t* vs = v_s + st * C::V_STAGE;
            mbar_expect_tx(&v_full[st], C::V_STAGE);
            if (p.v_mn) {
                tma_load_3d(vs, &map_v_hi, &v_full[st], 0, k0, bh);  // one {64 d x 128 keys} box
                if (SPLIT == 3) tma_load_3d(vs + C::V_BYTES, &map_v_lo, &v_full[st], 0, k0, bh);
            } else {
#pragma unroll
                for (int hb = 0; hb < BKV / 64; ++hb) {
                    tma_load_3d(vs + hb * C::VBOX_BYTES, &map_v_hi, &v_full[st], k0 + 64 * hb, 0, bh);
                    if (SPLIT == 3) tma_load_3d(vs + C::V_BYTES + hb * C::VBOX_BYTES, &map_v_lo, &v_full[st], k0 + 64 * hb, 0, bh);
                }
            }
        };
        if (n_tiles > 0) {
            load_q(0);
            load_k(0);
            for (int t = 0; t < n_tiles; ++t) {
                if (t + 1 < n_tiles) load_k(t + 1);  // K runs one tile ahead of V
                if constexpr (MODE == 0) load_v(t);
                if ((t + 1) % kv_tiles == 0 && t + 1 < n_tiles) load_q((t + 1) / kv_tiles);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, one elected lane per instruction; see gemm_tc.cu) =====================
        constexpr uint32_t idesc_s = make_idesc(BQ, BKV);
        // bit 16: B is MN-major; P16: A and B formats (bits 7..9, 10..12) = F16 (0) instead of BF16 (1)
        const uint32_t idesc_o = (make_idesc(BQ, HD) & ~(P16 ? ((7u << 7) | (7u << 10)) : 0u)) | (p.v_mn ? (1u << 16) : 0u);
        uint32_t g = 0, w = 0;
        const uint32_t q_hi = smem_u32(q_s), q_lo = q_hi + C::Q_BYTES;
        auto issue_S = [&](uint32_t gg) {
            const int st = gg & 1, kst = gg % C::K_STAGES;
            mbar_wait(&k_full[kst], (gg / C::K_STAGES) & 1);
            tc_fence_after();
            const uint32_t k_hi = smem_u32(k_s + kst * C::K_STAGE), k_lo = k_hi + C::K_BYTES;
            const uint32_t d = tmem_base + (st ? COL_S1 : COL_S0);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
                const uint32_t ko = k * 32;
                umma_ss(d, make_desc(q_hi + ko), make_desc(k_hi + ko), idesc_s, k != 0);
                if (SPLIT == 3) {
                    umma_ss(d, make_desc(q_lo + ko), make_desc(k_hi + ko), idesc_s, 1);
                    umma_ss(d, make_desc(q_hi + ko), make_desc(k_lo + ko), idesc_s, 1);
                }
            }
            umma_commit(&k_empty[kst]);  // K slot frees when these MMAs retire
            umma_commit(&s_full[st]);
        };
        for (int item = blockIdx.x; item < total; item += gridDim.x, ++w) {
            mbar_wait(q_full, w & 1);
            tc_fence_after();
            // S runs two tiles ahead of PV.  S(g) reuses the buffer of tile g-2, whose PV was issued earlier: in-order
            // execution of the tensor pipe orders the overwrite after PV(g-2) has read P(g-2) from it.
            issue_S(g);
            if (kv_tiles > 1) issue_S(g + 1);
            if (kv_tiles <= 2) umma_commit(q_empty);  // every S MMA of this item has been issued: Q frees when they retire
            for (int j = 0; j < kv_tiles; ++j, ++g) {
                const int st = g & 1;
                mbar_wait(&p_full[st], (g >> 1) & 1);
                if constexpr (MODE == 1) {
                    // column-sum mode: no P, no PV.  p_full(g) here means "every softmax warp holds S(g) in registers",
                    // which is what frees the S buffer for tile g + 2
                    tc_fence_after();
                    if (j + 2 < kv_tiles) {
                        issue_S(g + 2);
                        if (j + 3 == kv_tiles) umma_commit(q_empty);
                    }
                    continue;
                }
                mbar_wait(&v_full[st], (g >> 1) & 1);
                tc_fence_after();
                const uint32_t v_hi = smem_u32(v_s + st * C::V_STAGE), v_lo = v_hi + C::V_BYTES;
                const uint32_t d = tmem_base + COL_O;
                const uint32_t p_col = tmem_base + (st ? COL_S1 : COL_S0);
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks) {
                    const uint32_t vo = p.v_mn ? ks * 2048 : (ks >> 2) * C::VBOX_BYTES + (ks & 3) * 32;
                    const uint32_t a_hi = p_col + ks * 8, a_lo = p_col + BKV / 2 + ks * 8;
                    const uint64_t bh_d = p.v_mn ? make_desc_mn(v_hi + vo) : make_desc(v_hi + vo);
                    umma_ts(d, a_hi, bh_d, idesc_o, (j | ks) != 0);
                    if (SPLIT == 3) {
                        if (!P16) umma_ts(d, a_lo, bh_d, idesc_o, 1);
                        umma_ts(d, a_hi, p.v_mn ? make_desc_mn(v_lo + vo) : make_desc(v_lo + vo), idesc_o, 1);
                    }
                }
                umma_commit(&v_empty[st]);
                umma_commit(&o_done[st]);
                if (j + 2 < kv_tiles) {
                    issue_S(g + 2);
                    if (j + 3 == kv_tiles) umma_commit(q_empty);
                }
            }
        }
    } else if (warp >= 2) {
        // ===================== softmax + epilogue (16 warps) =====================
        // thread <-> query row (TMEM lane) x one quarter of the 128 key columns of a tile (32 scores held in
        // registers -> single pass over S).  The four warps of a lane quarter exchange their row maxima / row sums
        // through shared memory and split the 64 O columns for rescaling and the epilogue.
        const int qd = warp & 3, part = (warp - 2) >> 2;
        const int r = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint32_t xch = smem_u32(tmem_base_s + 4);  // float [2 parity][NPART][128 rows]
        const uint32_t bar_id = 1 + qd;
        uint32_t g = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            const int bh = item / q_tiles, q0 = (item % q_tiles) * BQ;
            // keys of this batch element that are real tokens: the fixed [B, K] layout of the batched pipeline pads
            // frames with fewer keypoints; a padded key must not receive attention (the reference runs n[b] tokens)
            const int kvbh = (bh + p.kv_shift >= p.BH) ? bh + p.kv_shift - p.BH : bh + p.kv_shift;
            const int nk_b = p.nk_counts ? max(1, min(p.Nk, __ldg(p.nk_counts + kvbh / p.heads))) : p.Nk;
            if constexpr (MODE == 1) {
                // ---- column sums: this thread's row is a KEY, the streamed columns are the queries whose row statistics are
                // known (lse_in), so every probability is final the moment its score is read: no running max, no P, no O
                float acc0 = 0.f, acc1 = 0.f;
                const float* lrow = p.lse_in + (long long)bh * p.ld_lse + part * CPT;
                for (int j = 0; j < kv_tiles; ++j, ++g) {
                    const int st = g & 1;
                    const uint32_t s_addr = tmem_base + lane_off + (st ? COL_S1 : COL_S0) + part * CPT;
                    const int nvalid = min(BKV, nk_b - j * BKV) - part * CPT;
                    float4 L[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) L[i] = __ldg(reinterpret_cast<const float4*>(lrow + j * BKV) + i);
                    mbar_wait(&s_full[st], (g >> 1) & 1);
                    tc_fence_after();
                    uint32_t v0[32];
                    tmem_ld32(s_addr, v0);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_full[st]);  // S(g) is in registers: its buffer may be overwritten
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 lv = L[i >> 2];
                        float e0, e1, e2, e3;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(v0[i]), p.scale_log2, -lv.x)));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(v0[i + 1]), p.scale_log2, -lv.y)));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(__uint_as_float(v0[i + 2]), p.scale_log2, -lv.z)));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fmaf(__uint_as_float(v0[i + 3]), p.scale_log2, -lv.w)));
                        if (nvalid < CPT) {  // padded / out-of-range queries contribute nothing (their statistics may be garbage)
                            e0 = (i < nvalid) ? e0 : 0.f; e1 = (i + 1 < nvalid) ? e1 : 0.f;
                            e2 = (i + 2 < nvalid) ? e2 : 0.f; e3 = (i + 3 < nvalid) ? e3 : 0.f;
                        }
                        acc0 += e0 + e1; acc1 += e2 + e3;
                    }
                }
                const uint32_t x = xch + (uint32_t)r * 4u;
                sts_f32(x + part * 512, acc0 + acc1);
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NPART) : "memory");
                if (part == 0 && q0 + r < p.Nq) {
                    float tot = lds_f32(x);
#pragma unroll
                    for (int q2 = 1; q2 < NPART; ++q2) tot += lds_f32(x + q2 * 512);  // fixed order: bit-reproducible
                    p.colsum[(long long)bh * p.ld_colsum + q0 + r] = tot;
                }
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NPART) : "memory");
                continue;
            }
            float m_used = -INFINITY, l = 0.f;  // l: this thread's partial row sum (its 32 columns)
            for (int j = 0; j < kv_tiles; ++j, ++g) {
                const int st = g & 1;
                const uint32_t s_addr = tmem_base + lane_off + (st ? COL_S1 : COL_S0) + part * CPT;
                const int nvalid = min(BKV, nk_b - j * BKV) - part * CPT;  // valid columns of this part (may be <= 0)
                mbar_wait(&s_full[st], (g >> 1) & 1);
                tc_fence_after();
                uint32_t v0[32];
                tmem_ld32(s_addr, v0);
                if (nvalid < CPT) {  // last, partial tile: mask the tail (warp-uniform branch)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i >= nvalid) v0[i] = 0xff800000u;  // -inf
                }
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    mx4[0] = fmaxf(mx4[0], __uint_as_float(v0[i]));
                    mx4[1] = fmaxf(mx4[1], __uint_as_float(v0[i + 1]));
                    mx4[2] = fmaxf(mx4[2], __uint_as_float(v0[i + 2]));
                    mx4[3] = fmaxf(mx4[3], __uint_as_float(v0[i + 3]));
                }
                float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
                mx *= p.scale_log2;  // scale > 0
                const uint32_t x = xch + (uint32_t)((g & 1) * (NPART * 128) + r) * 4u;
                sts_f32(x + part * 512, mx);
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NPART) : "memory");
                mx = lds_f32(x);
#pragma unroll
                for (int q2 = 1; q2 < NPART; ++q2) mx = fmaxf(mx, lds_f32(x + q2 * 512));
                // lazy rescale: keep the stale reference max unless it moved by more than 8 (log2 units)
                const bool need = (mx > m_used + 8.f);
                const bool warp_need = __any_sync(0xffffffffu, need) || (j == 0);
                float corr = 1.f;
                if (warp_need) {
                    const float m_new = fmaxf(m_used, mx);
                    corr = (m_used == -INFINITY) ? 0.f : exp2f(m_used - m_new);
                    m_used = m_new;
                    l *= corr;
                }
                // p = exp2(s*scale - m) for this thread's 32 columns, packed to bf16 hi / lo
                uint32_t ph[16], pl[16];
                float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0, p1;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(v0[i]), p.scale_log2, -m_used)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(v0[i + 1]), p.scale_log2, -m_used)));
                    if constexpr (P16) {
                        const __half2 hh = __floats2half2_rn(p0, p1);
                        ph[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                        const float2 pr = __half22float2(hh);  // the row sum follows the rounded weights
                        ls0 += pr.x; ls1 += pr.y;
                        continue;
                    }
                    ls0 += p0; ls1 += p1;
                    __nv_bfloat162 ha = __floats2bfloat162_rn(p0, p1);
                    const uint32_t ua = *reinterpret_cast<uint32_t*>(&ha);
                    ph[i >> 1] = ua;
                    if (SPLIT == 3) {
                        __nv_bfloat162 la = __floats2bfloat162_rn(p0 - __uint_as_float(ua << 16), p1 - __uint_as_float(ua & 0xffff0000u));
                        pl[i >> 1] = *reinterpret_cast<uint32_t*>(&la);
                    }
                }
                l += ls0 + ls1;
                // O may be touched only after the previous PV retired (rare: the reference max moved by > 2^8)
                if (j > 0 && warp_need) {
                    mbar_wait(&o_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
                    tc_fence_after();
                    const uint32_t o_addr = tmem_base + lane_off + COL_O + part * OPT;
                    uint32_t o[OPT];
                    tmem_ldN(o_addr, o);
#pragma unroll
                    for (int i = 0; i < OPT; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                    tmem_stN(o_addr, o);
                }
                // P over S, in place: all four warps of this lane quarter loaded their scores before the bar.sync above
                const uint32_t p_addr = tmem_base + lane_off + (st ? COL_S1 : COL_S0) + part * (CPT / 2);
                tmem_st16(p_addr, ph);
                if (SPLIT == 3 && !P16) tmem_st16(p_addr + BKV / 2, pl);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[st]);
            }
            // epilogue: O / l (row sum over the four parts), each warp stores 16 of the 64 head dims
            const uint32_t x = xch + (uint32_t)((g & 1) * (NPART * 128) + r) * 4u;
            sts_f32(x + part * 512, l);
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NPART) : "memory");
            float lsum_all = lds_f32(x);
#pragma unroll
            for (int q2 = 1; q2 < NPART; ++q2) lsum_all += lds_f32(x + q2 * 512);
            const float inv = 1.f / lsum_all;
            const int qn = q0 + r;
            if (p.lse_out && part == 0 && qn < p.Nq) p.lse_out[(long long)bh * p.ld_lse + qn] = m_used + log2f(lsum_all);
            mbar_wait(&o_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
            tc_fence_after();
            const int b = bh / p.heads, hh = bh - b * p.heads;
            const long long orow = ((long long)b * p.Nq + qn) * p.out_ld + hh * HD + part * OPT;
            {
                uint32_t v[OPT];
                tmem_ldN(tmem_base + lane_off + COL_O + part * OPT, v);
                if (qn < p.Nq) {
                    float f[OPT];
#pragma unroll
                    for (int i = 0; i < OPT; ++i) f[i] = __uint_as_float(v[i]) * inv;
                    if (p.out_f32) {
#pragma unroll
                        for (int i = 0; i < OPT; i += 4)
                            *reinterpret_cast<float4*>(p.out_f32 + orow + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
                    }
                    if (p.out_hi) {
                        uint32_t hi[OPT / 2], lo[OPT / 2];
#pragma unroll
                        for (int i = 0; i < OPT; i += 2) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[i], f[i + 1]);
                            const uint32_t u = *reinterpret_cast<uint32_t*>(&h2);
                            hi[i >> 1] = u;
                            __nv_bfloat162 l2 = __floats2bfloat162_rn(f[i] - __uint_as_float(u << 16), f[i + 1] - __uint_as_float(u & 0xffff0000u));
                            lo[i >> 1] = *reinterpret_cast<uint32_t*>(&l2);
                        }
                        uint4* oh = reinterpret_cast<uint4*>(p.out_hi + orow);
#pragma unroll
                        for (int i = 0; i < OPT / 8; ++i) oh[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                        if (p.out_lo) {
                            uint4* ol = reinterpret_cast<uint4*>(p.out_lo + orow);
#pragma unroll
                            for (int i = 0; i < OPT / 8; ++i) ol[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                        }
                    }
                }
            }
            tc_fence_before();  // O reads ordered before the next item's first PV (gated by p_full)
            // the partners must have read this item's row sums before the exchange slot is reused
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NPART) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}
static int encode3(CUtensorMap* m, const void* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2, cuuint64_t s1_bytes,
                   cuuint64_t s2_bytes, cuuint32_t b0, cuuint32_t b1) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return PRAM_ERR_CUDA;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t str[2] = {s1_bytes, s2_bytes};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, str, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PRAM_OK : PRAM_ERR_CUDA;
}

}  // namespace fa

// q/k: bf16 [BH][N][64]; v_mn = 0: vt bf16 [BH][64][nk_pad] (keys contiguous, nk_pad % 8 == 0, padding zeroed);
// v_mn = 1: vt is V itself, bf16 [BH][Nk][64] (consumed as an MN-major UMMA operand, no transposition);
// *_lo may be NULL when split == 1.  out: f32 and/or split bf16 [B][Nq][out_ld] at column head*64.
// mode 1 (column sums): q_* = the KEY operand (rows), k_* = the QUERY operand (columns), no V / out.
static int attention_launch(int mode, const void* q_hi, const void* q_lo, const void* k_hi, const void* k_lo, const void* vt_hi,
                            const void* vt_lo, int B, int heads, int Nq, int Nk, int nk_pad, float scale, float* out_f32,
                            void* out_hi, void* out_lo, int out_ld, int split, int p_swap, int v_mn, const int* nk_counts,
                            float* lse_out, const float* lse_in, int ld_lse, float* colsum, int ld_colsum, cudaStream_t stream,
                            int kv_shift = 0) {
    using namespace fa;
    const bool p16 = (v_mn & 2) != 0;  // V planes hold IEEE fp16 hi / lo: probabilities as one fp16 plane (kernel variant P16)
    v_mn &= 1;
    if (!q_hi || !k_hi || (mode == 0 && !vt_hi) || B <= 0 || heads <= 0 || Nq <= 0 || Nk <= 0) return PRAM_ERR_ARG;
    if (split != 1 && split != 3) return PRAM_ERR_ARG;
    if (p_swap != 0 && p_swap != 64 && p_swap != 128) return PRAM_ERR_ARG;  // key-tile variant: 0 = auto, 64, 128
    if (split == 3 && (!q_lo || !k_lo || (mode == 0 && !vt_lo))) return PRAM_ERR_ARG;
    if (mode == 0 && ((!v_mn && ((nk_pad % 8) || nk_pad < Nk)) || (out_ld % 8))) return PRAM_ERR_UNSUPPORTED;
    if (mode == 1 && (!lse_in || !colsum || ld_colsum < Nq)) return PRAM_ERR_ARG;
    // the statistics rows are read / written in whole key tiles with 16-byte loads
    if ((lse_out || lse_in) && ((ld_lse % 4) || ld_lse < (mode == 1 ? (Nk + 127) / 128 * 128 : Nq))) return PRAM_ERR_ARG;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        PRAM_CUDA(cudaGetDevice(&dev));
        PRAM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int BH = B * heads;
    CUtensorMap mq[2], mk[2], mv[2];
    const void* qs[2] = {q_hi, q_lo ? q_lo : q_hi};
    const void* ks[2] = {k_hi, k_lo ? k_lo : k_hi};
    const void* vs[2] = {vt_hi, vt_lo ? vt_lo : vt_hi};
    // key-tile variant: 64 (two CTAs per SM, out of phase) unless asked otherwise
    const int bkv = (p_swap == 128) ? 128 : 64;
    for (int i = 0; i < 2; ++i) {
        int rc = encode3(&mq[i], qs[i], HD, Nq, BH, HD * 2, (cuuint64_t)Nq * HD * 2, HD, BQ);
        if (rc) return rc;
        rc = encode3(&mk[i], ks[i], HD, Nk, BH, HD * 2, (cuuint64_t)Nk * HD * 2, HD, bkv);
        if (rc) return rc;
        if (mode == 1) { mv[i] = mk[i]; continue; }
        if (v_mn) rc = encode3(&mv[i], vs[i], HD, Nk, BH, HD * 2, (cuuint64_t)Nk * HD * 2, HD, bkv);
        else rc = encode3(&mv[i], vs[i], nk_pad, HD, BH, (cuuint64_t)nk_pad * 2, (cuuint64_t)nk_pad * HD * 2, 64, HD);
        if (rc) return rc;
    }
    Args a;
    a.BH = BH; a.heads = heads; a.Nq = Nq; a.Nk = Nk;
    a.scale_log2 = scale * 1.4426950408889634f;
    a.out_f32 = out_f32; a.out_hi = (__nv_bfloat16*)out_hi; a.out_lo = (__nv_bfloat16*)out_lo; a.out_ld = out_ld;
    a.v_mn = v_mn;
    if (kv_shift < 0 || kv_shift >= B || (kv_shift && mode != 0)) return PRAM_ERR_ARG;
    a.nk_counts = nk_counts;
    a.kv_shift = kv_shift * heads;
    a.pred = g_pram_pred;
    a.pdl_early = g_pram_pdl >= 2;
    a.lse_out = lse_out; a.lse_in = lse_in; a.ld_lse = ld_lse; a.colsum = colsum; a.ld_colsum = ld_colsum;
    const int total = BH * ((Nq + BQ - 1) / BQ);
#define PRAM_ATT_LAUNCH(SPLIT_, BKV_, MODE_, P16_)                                                                      \
    do {                                                                                                                \
        auto kern = attention_tc_kernel<SPLIT_, BKV_, MODE_, P16_>;                                                     \
        static bool attr = false;                                                                                       \
        if (!attr) { PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<SPLIT_, BKV_>::SMEM_BYTES)); attr = true; } \
        const int cap = Geo<BKV_>::CTAS_PER_SM * num_sms;                                                               \
        const int grid = total < cap ? total : cap;                                                                     \
        PRAM_CUDA(pram_launch_pdl(kern, dim3(grid), dim3(Geo<BKV_>::NUM_THREADS), Cfg<SPLIT_, BKV_>::SMEM_BYTES, stream, mq[0], mq[1], mk[0], mk[1], mv[0], mv[1], a)); \
    } while (0)
#define PRAM_ATT_MODE(MODE_, P16_)                                                                                      \
    do {                                                                                                                \
        if (split == 3) { if (bkv == 128) PRAM_ATT_LAUNCH(3, 128, MODE_, P16_); else PRAM_ATT_LAUNCH(3, 64, MODE_, P16_); } \
        else { if (bkv == 128) PRAM_ATT_LAUNCH(1, 128, MODE_, P16_); else PRAM_ATT_LAUNCH(1, 64, MODE_, P16_); }        \
    } while (0)
    if (mode == 1) PRAM_ATT_MODE(1, 0); else if (p16) PRAM_ATT_MODE(0, 1); else PRAM_ATT_MODE(0, 0);
#undef PRAM_ATT_MODE
#undef PRAM_ATT_LAUNCH
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

PRAM_API int pram_attention_tc(const void* q_hi, const void* q_lo, const void* k_hi, const void* k_lo, const void* vt_hi,
                               const void* vt_lo, int B, int heads, int Nq, int Nk, int nk_pad, float scale,
                               float* out_f32, void* out_hi, void* out_lo, int out_ld, int split, int p_swap,
                               int v_mn, const int* nk_counts, cudaStream_t stream) {
    return attention_launch(0, q_hi, q_lo, k_hi, k_lo, vt_hi, vt_lo, B, heads, Nq, Nk, nk_pad, scale, out_f32, out_hi, out_lo,
                            out_ld, split, p_swap, v_mn, nk_counts, nullptr, nullptr, 0, nullptr, 0, stream);
}

// same with rotated key / value batches: query batch element b attends to the keys / values (and nk_counts) of batch element
// (b + kv_shift) mod B.  B = 2 x pairs, kv_shift = pairs, q = k = [set 0 | set 1]: both directions of the GML cross attention
// (nets/gml.py:175-181) in ONE launch -- 2048 work items fill 296 CTA slots 6.9 times instead of 2 x 3.5 (rounded up to 2 x 4).
PRAM_API int pram_attention_tc_shift(const void* q_hi, const void* q_lo, const void* k_hi, const void* k_lo, const void* vt_hi,
                                     const void* vt_lo, int B, int heads, int Nq, int Nk, int nk_pad, float scale,
                                     float* out_f32, void* out_hi, void* out_lo, int out_ld, int split, int p_swap,
                                     int v_mn, const int* nk_counts, int kv_shift, cudaStream_t stream) {
    return attention_launch(0, q_hi, q_lo, k_hi, k_lo, vt_hi, vt_lo, B, heads, Nq, Nk, nk_pad, scale, out_f32, out_hi, out_lo,
                            out_ld, split, p_swap, v_mn, nk_counts, nullptr, nullptr, 0, nullptr, 0, stream, kv_shift);
}

// same, additionally writing the log2-domain log-sum-exp of every query row to lse_out [B*heads][ld_lse]
PRAM_API int pram_attention_tc_lse(const void* q_hi, const void* q_lo, const void* k_hi, const void* k_lo, const void* vt_hi,
                                   const void* vt_lo, int B, int heads, int Nq, int Nk, int nk_pad, float scale,
                                   float* out_f32, void* out_hi, void* out_lo, int out_ld, int split, int p_swap,
                                   int v_mn, const int* nk_counts, float* lse_out, int ld_lse, cudaStream_t stream) {
    return attention_launch(0, q_hi, q_lo, k_hi, k_lo, vt_hi, vt_lo, B, heads, Nq, Nk, nk_pad, scale, out_f32, out_hi, out_lo,
                            out_ld, split, p_swap, v_mn, nk_counts, lse_out, nullptr, ld_lse, nullptr, 0, stream);
}

// AdaGML's per-key attention mass (nets/adagml.py:148, 229) on the tensor cores: colsum[bh][key] = sum over the valid queries
// of softmax(query, key), from S^T = K Q^T tiles and the queries' row statistics of the pram_attention_tc_lse launch.
PRAM_API int pram_attention_colsum_tc(const void* key_hi, const void* key_lo, const void* qry_hi, const void* qry_lo, int B,
                                      int heads, int Nkeys, int Nqueries, float scale, const float* lse, int ld_lse,
                                      float* colsum, int ld_colsum, int split, int kv_tile, const int* nq_counts,
                                      cudaStream_t stream) {
    return attention_launch(1, key_hi, key_lo, qry_hi, qry_lo, nullptr, nullptr, B, heads, Nkeys, Nqueries, 0, scale, nullptr,
                            nullptr, nullptr, 0, split, kv_tile, 1, nq_counts, nullptr, lse, ld_lse, colsum, ld_colsum, stream);
}

// ------------------------------------------------------------------------------------------
// qkv [tokens][nparts*heads*64] fp32 -> split-bf16 operands of the attention kernel:
//   Q, K [B][heads][N][64] (rotary on adjacent pairs, scale_qk applied), V^T [B][heads][64][n_pad]
// nparts = 3: (q | k | v);  nparts = 2: (qk | v) (cross attention: only Q=qk and V^T are produced)
// ------------------------------------------------------------------------------------------
__global__ void qk_prep_kernel(const float* __restrict__ qkv, int nparts, int B, int N, int heads,
                               const float* __restrict__ cosb, const float* __restrict__ sinb, float scale_qk,
                               __nv_bfloat16* __restrict__ q_hi, __nv_bfloat16* __restrict__ q_lo,
                               __nv_bfloat16* __restrict__ k_hi, __nv_bfloat16* __restrict__ k_lo) {
    const int nqk = nparts - 1;  // number of rotated parts
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * N * nqk * heads * 32;
    if (i >= total) return;
    const int pr = (int)(i & 31);
    const int h = (int)((i >> 5) % heads);
    const int part = (int)((i / (32 * heads)) % nqk);
    const long long t = i / ((long long)32 * heads * nqk);
    const int b = (int)(t / N), n = (int)(t - (long long)b * N);
    const float* src = qkv + t * (long long)(nparts * heads * 64) + (long long)part * heads * 64 + h * 64 + 2 * pr;
    float x0 = src[0], x1 = src[1];
    if (cosb) {
        const float c = cosb[t * 32 + pr], s = sinb[t * 32 + pr];
        const float y0 = x0 * c + (-x1) * s, y1 = x1 * c + x0 * s;
        x0 = y0; x1 = y1;
    }
    x0 *= scale_qk; x1 *= scale_qk;
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const long long o = (((long long)b * heads + h) * N + n) * 64 + 2 * pr;
    __nv_bfloat16* dh = part == 0 ? q_hi : k_hi;
    __nv_bfloat16* dl = part == 0 ? q_lo : k_lo;
    *reinterpret_cast<uint32_t*>(dh + o) = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    if (dl) {
        const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
        *reinterpret_cast<uint32_t*>(dl + o) = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
}

// V part -> V^T [B][heads][64][n_pad] through a shared-memory transpose (coalesced on both sides)
__global__ void __launch_bounds__(256) vt_prep_kernel(const float* __restrict__ qkv, int row_stride, int v_off, int B, int N,
                                                      int heads, int n_pad, __nv_bfloat16* __restrict__ vt_hi,
                                                      __nv_bfloat16* __restrict__ vt_lo) {
    __shared__ float tile[64][65];
    const int n0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
        const int tk = i >> 6, d = i & 63;
        const int n = n0 + tk;
        tile[tk][d] = (n < N) ? qkv[((long long)b * N + n) * row_stride + v_off + h * 64 + d] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
        const int d = i >> 6, tk = i & 63;
        const int n = n0 + tk;
        if (n < n_pad) {
            const float v = tile[tk][d];
            const __nv_bfloat16 hv = __float2bfloat16_rn(v);
            const long long o = (((long long)b * heads + h) * 64 + d) * n_pad + n;
            vt_hi[o] = hv;
            if (vt_lo) vt_lo[o] = __float2bfloat16_rn(v - __bfloat162float(hv));
        }
    }
}

PRAM_API int pram_attention_prep(const float* qkv, int nparts, int B, int N, int heads, const float* cosb,
                                 const float* sinb, float scale_qk, void* q_hi, void* q_lo, void* k_hi, void* k_lo,
                                 void* vt_hi, void* vt_lo, int n_pad, cudaStream_t stream) {
    if (!qkv || !q_hi || !vt_hi || (nparts != 2 && nparts != 3) || (nparts == 3 && !k_hi) || n_pad < N) return PRAM_ERR_ARG;
    const long long total = (long long)B * N * (nparts - 1) * heads * 32;
    qk_prep_kernel<<<cdiv(total, 256), 256, 0, stream>>>(qkv, nparts, B, N, heads, cosb, sinb, scale_qk,
                                                        (__nv_bfloat16*)q_hi, (__nv_bfloat16*)q_lo, (__nv_bfloat16*)k_hi,
                                                        (__nv_bfloat16*)k_lo);
    PRAM_CHECK_LAUNCH();
    dim3 grid(cdiv(n_pad, 64), heads, B);
    vt_prep_kernel<<<grid, 256, 0, stream>>>(qkv, nparts * heads * 64, (nparts - 1) * heads * 64, B, N, heads, n_pad,
                                             (__nv_bfloat16*)vt_hi, (__nv_bfloat16*)vt_lo);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}
