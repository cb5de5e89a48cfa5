// tcgen05 tensor-core implicit GEMM for sm_100a: the 3x3 / 1x1 convolutions of SFD2 (K1-K4) and every
// Linear layer / batched A.B^T of SegNetViT and GML (K10-K14) -- reference nets/sfd2.py:141-170,
// nets/segnetvit.py:88-106, nets/gml.py:119-186,278-282 (cuDNN / cuBLAS fp32 there).
//
//   D[pixel, n] = sum_{tap, c} A[pixel + offset(tap), c] * W[tap][n][c]       (fp32 accumulate in TMEM)
//
// * persistent kernel, one CTA per SM, static round-robin tile scheduler
// * warp 0 (one lane): TMA producer -- A tiles are 4-D boxes {64 ch, TW, TH, 1} of the NHWC activation
//   tensor shifted by the filter tap (out-of-bounds = zero fill = the convolution's padding), which land
//   in shared memory as a 128 x 64 K-major SWIZZLE_128B operand; W tiles are {64, BN, 1} boxes
// * warp 1 (one lane): tcgen05.mma issuer, accumulators in TMEM, double-buffered (2 x BN columns) so
//   the epilogue of tile i overlaps the main loop of tile i+1
// * warps 2-9: epilogue -- tcgen05.ld -> smem transpose -> bias / residual / ReLU -> coalesced fp32
//   and/or split-bf16 stores
//   (NHWC, and optionally a 2x2 phase-split copy that feeds a following stride-2 convolution with
//   unit-stride TMA boxes)
// * SPLIT=3: error-compensated bf16x3 (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, all into the same fp32
//   accumulator) -- ~16 mantissa bits, needed for bit-stable keypoint selection;
//   SPLIT=1: plain bf16.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdio.h>

namespace tc {

constexpr int BM = 128;       // UMMA_M
constexpr int BK = 64;        // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int MAX_TAPS = 9;

struct Args {
    // geometry of the output
    int B, Ho, Wo, N;
    int tw_log2;         // tile = (128 >> tw_log2) rows x (1 << tw_log2) columns of output pixels
    int ntaps, kblocks;  // kblocks = ceil(Cin / 64)
    int planes_per_image;  // 1, or 4 when A is a phase-split tensor
    int tap_dx[MAX_TAPS], tap_dy[MAX_TAPS], tap_plane[MAX_TAPS];
    int w_batch_mult;    // weight plane = tap + b * w_batch_mult
    // epilogue
    const float* bias;
    const float* res; long long res_ld;
    const __nv_bfloat16* res_hi; const __nv_bfloat16* res_lo;  // residual given as split-bf16 planes (res == NULL): r = hi + lo
    int relu;
    float* out_f32; long long ld_f32;
    __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; long long ld_bf;
    __nv_bfloat16* ps_hi; __nv_bfloat16* ps_lo; long long ld_ps;  // phase-split copy
    int l2norm;          // normalise each output row (requires N <= BN)
    // fused attention-operand epilogue (qkv_mode 1: columns (q|k|v), rotary on q,k; 2: columns (qk|v)):
    // output goes to split-bf16 tensors laid out [B][heads][n][64] per token segment instead of rows
    int qkv_mode;
    const float* cosb; const float* sinb; float qk_scale;
    __nv_bfloat16* q_hi; __nv_bfloat16* q_lo; __nv_bfloat16* k_hi; __nv_bfloat16* k_lo; __nv_bfloat16* v_hi; __nv_bfloat16* v_lo;
    int seg_split, seg_n0, seg_n1, heads;
    int l2_prefetch;     // 1: pull the next tile's activation boxes into L2 one tile ahead (single-tap layers)
    const int* pred;     // launch predicate (common.cuh): the whole grid returns when *pred == 0
    int pdl_early;    // 1: release the dependent launch right after this grid's own wait (common.cuh)
    int v_f16;           // qkv epilogue: the V planes receive IEEE fp16 hi / lo instead of bf16 hi / lo
    __half* out_h16;     // optional: the output once more as ONE IEEE fp16 plane (row stride ld_bf) -- feeds a single-pass fp16 layer
    int f16;             // operands are IEEE fp16 (one MMA per k-step, 11-bit mantissas): A / W planes hold fp16 bits and the
                         // out_hi plane receives fp16 (out_lo unused) -- the single-pass mode of the descriptor head
};

// ---------------------------------------------------------------------------------------- PTX
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
// bounded wait: a protocol bug must trap, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait suspends the thread for a hardware-bounded time per call; the spin counter (2 instructions per poll
    // instead of a 64-bit clock comparison) turns a protocol bug into a trap after seconds instead of a hang
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) {
            printf("pram gemm_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// multicast variants for 2-CTA clusters: the box lands at the same CTA-relative offset in every CTA of `mask`
// and completes bytes on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6}], [%2], %3;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// L2 prefetch of a future operand box (no shared memory, no barrier): shortens the latency of the later TMA load
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 UMMA): start>>4, LBO=0, SBO=1024 B
// (8 rows x 128 B), version 1, layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major both,
// N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with A = B = fp16 (format code 0 in bits 7-9 / 10-12)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// The MMA-issuing WARP runs its loops convergently and elects one lane per instruction (elect.sync): operands that
// are warp-uniform then stay in uniform registers and each MMA costs a handful of uniform-datapath instructions.
// Issued from inside an `if (lane == 0)` region instead, the compiler wraps every tcgen05.mma in an
// elect / R2UR / branch loop (~10 dependent instructions, ~100 cycles per MMA -- measured with ncu: the issue
// thread, not the tensor pipe, bounded every kernel whose MMAs are shorter than that).
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Explicit shared-space accessors for the epilogue staging tile and row tables.  The staging pointers are derived
// from the manually 1024-byte-aligned dynamic shared-memory base (integer arithmetic), so the compiler no longer
// knows their address space and emits GENERIC LD / ST (ncu: long-scoreboard stalls on every staged value, lg-throttle
// on the stores); these force LDS / STS.
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ int lds32(uint32_t a) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// per-lane-quarter row table of the epilogue (output pixel index per tile row, -1 = outside the map)
struct EpiQuarter {
    int pix[32];
    int pix_ps[32];
    float inv[32];
    // qkv_mode (never combined with the phase-split copy or the L2 norm) reuses two of the tables:
    //   pix_ps -> element offset of (token, head 0, dim 0) in the [B][heads][n][64] tensors
    //   inv    -> (as int) element stride between heads for this token's segment (n * 64)
};
constexpr int EPI_BYTES = 8 * 4096 + 4 * (int)sizeof(EpiQuarter);  // 8 swizzled 32x32 fp32 tiles + tables

template <int BN, int SPLIT>
struct Cfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int NPLANES = (SPLIT == 3) ? 2 : 1;
    static constexpr int STAGE_BYTES = NPLANES * (A_BYTES + B_BYTES);
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES >= 6 ? 6 : (200 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_BYTES;
    static_assert(SMEM_BYTES <= 227 * 1024, "dynamic shared memory budget of sm_100a exceeded");
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

// MODE (epilogue specialisation, keeps registers down): 0 plain, 1 residual add, 2 fused attention operands
// CL = 2: the kernel runs as 2-CTA clusters.  The pair works on two neighbouring M tiles of the SAME N tile, so the
// weight tile is identical for both: each CTA fetches one half of it (BN/2 rows) and TMA-multicasts it into both
// shared memories -- per-CTA L2->SM operand traffic drops from A + W to A + W/2 (the kernel is L2-bandwidth bound:
// 96 KB per k-block at bf16x3 / BN = 256).  A stage of a CTA is written by the peer too, so the "slot free" barrier
// collects the MMA commits of BOTH CTAs (tcgen05.commit.multicast).
template <int BN, int SPLIT, int MODE, int CL>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(
    const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
    const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const Args p) {
    using C = Cfg<BN, SPLIT>;
    if (p.pred) { pram_pdl_wait(); if (pram_pred_skip(p.pred)) return; }  // the flag is written by a predecessor kernel
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* full = bars;                       // [STAGES]
    uint64_t* empty = bars + C::STAGES;          // [STAGES]
    uint64_t* tfull = bars + 2 * C::STAGES;      // [2]
    uint64_t* tempty = bars + 2 * C::STAGES + 2; // [2]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TW = 1 << p.tw_log2, TH = BM >> p.tw_log2;
    const int tiles_x = (p.Wo + TW - 1) / TW, tiles_y = (p.Ho + TH - 1) / TH;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int m_tiles = p.B * tiles_y * tiles_x;
    const int crank = (CL == 2) ? (int)cluster_ctarank() : 0;
    const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;          // cluster index / number of clusters
    const int total = ((m_tiles + CL - 1) / CL) * n_tiles;           // work items per cluster: (M-tile group, N tile)
    const int num_kb = p.ntaps * p.kblocks;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a_hi);
        prefetch_tmap(&map_w_hi);
        if (SPLIT == 3) { prefetch_tmap(&map_a_lo); prefetch_tmap(&map_w_lo); }
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(C::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    pram_pdl_wait();     // programmatic dependent launch: everything above overlapped the predecessor's drain
    if (p.pdl_early) pram_pdl_trigger();  // all CTAs of a persistent grid are resident: the successor may be scheduled as SMs free up

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int stage = 0; uint32_t phase = 0;
        for (int tile = cid; tile < total; tile += ncl) {
            const int nt = tile % n_tiles, mt = (tile / n_tiles) * CL + crank;  // mt >= m_tiles: all-zero A (OOB fill)
            const int txi = mt % tiles_x, tyi = (mt / tiles_x) % tiles_y, b = mt / (tiles_x * tiles_y);
            const int x0 = txi * TW, y0 = tyi * TH, n0 = nt * BN;
            // optional experiment (off by default, no gain measured): Linear / 1x1 layers stream their activations from
            // HBM exactly once and only two 96 KB stages fit in shared memory; pull the next tile's boxes into L2 early
            if (p.l2_prefetch && tile + ncl < total) {
                const int nxt = tile + ncl;
                const int nnt = nxt % n_tiles, nmt = (nxt / n_tiles) * CL + crank;
                if (nnt == 0 || n_tiles == 1 || nmt != mt) {  // a new M tile: its boxes have not been requested yet
                    const int ntx = nmt % tiles_x, nty = (nmt / tiles_x) % tiles_y, nb = nmt / (tiles_x * tiles_y);
                    for (int kc = 0; kc < p.kblocks; ++kc) {
                        tma_prefetch_4d(&map_a_hi, kc * BK, ntx * TW + p.tap_dx[0], nty * TH + p.tap_dy[0], nb * p.planes_per_image + p.tap_plane[0]);
                        if (SPLIT == 3)
                            tma_prefetch_4d(&map_a_lo, kc * BK, ntx * TW + p.tap_dx[0], nty * TH + p.tap_dy[0], nb * p.planes_per_image + p.tap_plane[0]);
                    }
                }
            }
            for (int kb = 0; kb < num_kb; ++kb) {
                const int tap = kb / p.kblocks, kc = kb - tap * p.kblocks;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = smem + stage * C::STAGE_BYTES;
                mbar_expect_tx(&full[stage], C::STAGE_BYTES);
                const int ax = x0 + p.tap_dx[tap], ay = y0 + p.tap_dy[tap];
                const int ap = b * p.planes_per_image + p.tap_plane[tap];
                const int wp = tap + b * p.w_batch_mult;
                tma_load_4d(st, &map_a_hi, &full[stage], kc * BK, ax, ay, ap);
                if (SPLIT == 3) tma_load_4d(st + C::A_BYTES, &map_a_lo, &full[stage], kc * BK, ax, ay, ap);
                if (CL == 1) {
                    tma_load_3d(st + C::NPLANES * C::A_BYTES, &map_w_hi, &full[stage], kc * BK, n0, wp);
                    if (SPLIT == 3) tma_load_3d(st + 2 * C::A_BYTES + C::B_BYTES, &map_w_lo, &full[stage], kc * BK, n0, wp);
                } else {  // this CTA's half of the weight tile, multicast into both CTAs of the pair
                    const int hb = crank * (C::B_BYTES / 2), hn = n0 + crank * (BN / 2);
                    tma_load_3d_mc(st + C::NPLANES * C::A_BYTES + hb, &map_w_hi, &full[stage], kc * BK, hn, wp, (uint16_t)3);
                    if (SPLIT == 3)
                        tma_load_3d_mc(st + 2 * C::A_BYTES + C::B_BYTES + hb, &map_w_lo, &full[stage], kc * BK, hn, wp, (uint16_t)3);
                }
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, one elected lane per instruction) =====================
        const uint32_t idesc = p.f16 ? make_idesc_f16(BM, BN) : make_idesc(BM, BN);
        int stage = 0; uint32_t phase = 0;
        int as = 0; uint32_t aphase = 0;
        for (int tile = cid; tile < total; tile += ncl) {
            mbar_wait(&tempty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                const uint32_t a_hi = sa, a_lo = sa + C::A_BYTES;
                const uint32_t b_hi = sa + C::NPLANES * C::A_BYTES, b_lo = b_hi + C::B_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t koff = k * UMMA_K * 2;
                    umma(d_tmem, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, (kb | k) != 0);
                    if (SPLIT == 3) {
                        umma(d_tmem, make_desc(a_lo + koff), make_desc(b_hi + koff), idesc, 1);
                        umma(d_tmem, make_desc(a_hi + koff), make_desc(b_lo + koff), idesc, 1);
                    }
                }
                // frees the smem slot when these MMAs retire (in both CTAs of a pair: the peer writes half of W here)
                if (CL == 1) umma_commit(&empty[stage]);
                else umma_commit_mc(&empty[stage], (uint16_t)3);
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tfull[as]);  // accumulator ready for the epilogue
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    } else if (warp >= 2) {
        // ===================== epilogue (8 warps) =====================
        // TMEM lane quarter = warp % 4; the two warps of a quarter take alternate 32-column chunks.
        // tcgen05.ld hands each thread one output ROW (32 consecutive channels).  Storing that directly
        // makes every warp-level store touch 32 different 128-byte lines, so each warp transposes its
        // 32x32 chunk through an XOR-swizzled shared-memory tile: 8 (fp32) / 4 (bf16) lanes then cover one
        // row segment and every store / residual load moves whole lines.  Residual loads of the next
        // chunk are issued before the current chunk is processed (latency hidden).
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int r = q * 32 + lane;  // row of the tile == TMEM lane
        uint8_t* epi_base = smem + C::STAGES * C::STAGE_BYTES + 256;
        const uint32_t tile_a = smem_u32(epi_base) + (uint32_t)(warp - 2) * 4096u;          // [32][32] fp32, swizzled
        // per-quarter row tables (struct EpiQuarter): pix / pix_ps / inv, 32 entries each
        const uint32_t eq_pix = smem_u32(epi_base + 8 * 4096) + (uint32_t)q * (uint32_t)sizeof(EpiQuarter), eq_ps = eq_pix + 128u, eq_inv = eq_pix + 256u;
        const float* __restrict__ resp = p.res;
        const float* __restrict__ biasp = p.bias;
        const bool res_vec = ((p.res_ld & 3) == 0);
        const int g1 = lane >> 3, col1 = (lane & 7) * 4;   // pass-1 mapping
        const int g2 = lane >> 2, col2 = (lane & 3) * 8;   // pass-2 mapping
        // one-pass epilogue (chunk loop below) needs 16-byte accesses everywhere: checked once per kernel
        auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
        const bool fast_ok = al16(biasp) && (!p.out_f32 || (al16(p.out_f32) && (p.ld_f32 & 3) == 0)) &&
                             (!resp || (al16(resp) && (p.res_ld & 3) == 0)) &&
                             (resp || !p.res_hi || (al16(p.res_hi) && al16(p.res_lo) && (p.res_ld & 7) == 0));
        int as = 0; uint32_t aphase = 0;
        for (int tile = cid; tile < total; tile += ncl) {
            const int nt = tile % n_tiles, mt = (tile / n_tiles) * CL + crank;
            const int txi = mt % tiles_x, tyi = (mt / tiles_x) % tiles_y, b = mt / (tiles_x * tiles_y);
            const int n0 = nt * BN;
            if (half == 0) {
                const int y = tyi * TH + (r >> p.tw_log2), x = txi * TW + (r & (TW - 1));
                const bool valid = (y < p.Ho) && (x < p.Wo) && (mt < m_tiles);
                const int Hp = (p.Ho + 1) >> 1, Wp = (p.Wo + 1) >> 1;
                sts32(eq_pix + 4 * lane, valid ? (int)(((long long)b * p.Ho + y) * p.Wo + x) : -1);
                sts32(eq_ps + 4 * lane, (int)((((long long)(b * 4 + (y & 1) * 2 + (x & 1))) * Hp + (y >> 1)) * Wp + (x >> 1)));
                if (MODE == 2) {  // rows are tokens (Linear): token -> (segment, batch element, position)
                    const int t = x;
                    const bool s1 = t >= p.seg_split;
                    const int ns = s1 ? p.seg_n1 : p.seg_n0, tt = s1 ? t - p.seg_split : t;
                    const int bb = tt / ns, nn = tt - bb * ns;
                    sts32(eq_ps + 4 * lane, (s1 ? p.seg_split * p.heads * 64 : 0) + (bb * p.heads * ns + nn) * 64);
                    sts32(eq_inv + 4 * lane, ns * 64);
                }
            }
            // the two warps of a quarter exchange the row table through a named barrier (id 1 + q, 64 threads)
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            // Residual of one 32 x 32 chunk, kept as RAW bits until it is consumed: converting the bf16 planes right after the
            // loads made every chunk wait for its own DRAM round trip (ncu: conv4.x.conv3 ran at 0.95 ms against 0.29 ms for the
            // same GEMM without a residual); with the raw words held in registers the loads of chunk c + 1 stay in flight
            // while chunk c is processed.  fp32 residual: the four floats; bf16 planes: {hi.x, hi.y, lo.x, lo.y}.
            auto load_res = [&](int nb, uint4 (&rv)[MODE == 1 ? 8 : 1]) {
                if constexpr (MODE != 1) return;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int pixr = lds32(eq_pix + 4 * (i * 4 + g1));
                    uint4 t = make_uint4(0u, 0u, 0u, 0u);
                    if (pixr >= 0 && nb < p.N) {
                        if (resp) {
                            const float* rp = resp + (long long)pixr * p.res_ld + nb + col1;
                            if (res_vec && nb + 32 <= p.N) t = __ldg(reinterpret_cast<const uint4*>(rp));
                            else {
                                if (nb + col1 + 0 < p.N) t.x = __float_as_uint(__ldg(rp));
                                if (nb + col1 + 1 < p.N) t.y = __float_as_uint(__ldg(rp + 1));
                                if (nb + col1 + 2 < p.N) t.z = __float_as_uint(__ldg(rp + 2));
                                if (nb + col1 + 3 < p.N) t.w = __float_as_uint(__ldg(rp + 3));
                            }
                        } else if (p.res_hi) {
                            // residual from the producer's split-bf16 planes (host guarantees N % 32 == 0 and res_ld % 4 == 0):
                            // the activation then never needs an fp32 copy in HBM
                            const long long o = (long long)pixr * p.res_ld + nb + col1;
                            const uint2 h = __ldg(reinterpret_cast<const uint2*>(p.res_hi + o));
                            uint2 l = make_uint2(0u, 0u);
                            if (p.res_lo) l = __ldg(reinterpret_cast<const uint2*>(p.res_lo + o));
                            t = make_uint4(h.x, h.y, l.x, l.y);
                        }
                    }
                    rv[i] = t;
                }
            };
            auto res_value = [&](const uint4& t) -> float4 {
                if (resp) return make_float4(__uint_as_float(t.x), __uint_as_float(t.y), __uint_as_float(t.z), __uint_as_float(t.w));
                return make_float4(__uint_as_float(t.x << 16) + __uint_as_float(t.z << 16),
                                   __uint_as_float(t.x & 0xffff0000u) + __uint_as_float(t.z & 0xffff0000u),
                                   __uint_as_float(t.y << 16) + __uint_as_float(t.w << 16),
                                   __uint_as_float(t.y & 0xffff0000u) + __uint_as_float(t.w & 0xffff0000u));
            };
            // The same residual in the ONE-PASS mapping (4 lanes per row, 8 columns per lane; see the chunk loop): two 16-byte
            // words per row group -- fp32: columns 0-3 / 4-7; bf16 planes: the hi words / the lo words of the 8 columns.
            auto load_res_fast = [&](int nb, uint4 (&rf)[MODE == 1 ? 8 : 1]) {
                if constexpr (MODE != 1) return;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int pixr = lds32(eq_pix + 4 * (i * 8 + g2));
                    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
                    if (pixr >= 0) {
                        if (resp) {
                            const uint4* rp = reinterpret_cast<const uint4*>(resp + (long long)pixr * p.res_ld + nb + col2);
                            a = __ldg(rp);
                            b = __ldg(rp + 1);
                        } else {
                            const long long o = (long long)pixr * p.res_ld + nb + col2;
                            a = __ldg(reinterpret_cast<const uint4*>(p.res_hi + o));
                            if (p.res_lo) b = __ldg(reinterpret_cast<const uint4*>(p.res_lo + o));
                        }
                    }
                    rf[2 * i] = a;
                    rf[2 * i + 1] = b;
                }
            };
            uint4 rv[MODE == 1 ? 8 : 1];
            if constexpr (MODE == 1) {  // independent of the accumulator: overlaps the MMA tail
                if (fast_ok && n0 + half * 32 + 32 <= p.N) load_res_fast(n0 + half * 32, rv);
            }
            // MODE 2: the rotary factors of a row depend on (token, dim pair) only -- not on the head -- and this
            // warp's chunks (c = half*32 + 64 m) all cover dims half*32..+31 of head m: ONE load per tile serves
            // all four chunks and is issued before the accumulator wait (latency hidden behind the MMA).
            float4 rot_c[MODE == 2 ? 4 : 1], rot_s[MODE == 2 ? 4 : 1];
            if constexpr (MODE == 2) {
                const bool rot = (p.qkv_mode == 1) && (nt != 2);
                const int pr0 = (half * 32 + col2) >> 1;   // 4 consecutive (even, odd) pairs of this lane's 8 columns
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int pixr = lds32(eq_pix + 4 * (i * 8 + g2));
                    rot_c[i] = make_float4(1.f, 1.f, 1.f, 1.f);
                    rot_s[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rot && pixr >= 0) {
                        rot_c[i] = __ldg(reinterpret_cast<const float4*>(p.cosb + (long long)pixr * 32 + pr0));
                        rot_s[i] = __ldg(reinterpret_cast<const float4*>(p.sinb + (long long)pixr * 32 + pr0));
                    }
                }
            }
            // bias of the first chunk, loaded before the accumulator wait; every chunk then loads the NEXT chunk's bias while it
            // works (ncu: the first FADD of a chunk waited for this load -- the top stall of the qkv GEMM, 10 % of its samples)
            float4 bn0 = make_float4(0.f, 0.f, 0.f, 0.f), bn1 = bn0;
            auto load_bias = [&](int nb) {
                bn0 = make_float4(0.f, 0.f, 0.f, 0.f); bn1 = bn0;
                if (biasp && (MODE == 2 || fast_ok) && nb + 32 <= p.N) {
                    bn0 = __ldg(reinterpret_cast<const float4*>(biasp + nb + col2));
                    bn1 = __ldg(reinterpret_cast<const float4*>(biasp + nb + col2 + 4));
                }
            };
            load_bias(n0 + half * 32);
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            if (p.l2norm) {
                // Row norms (N <= BN <= 256): each warp of the quarter sums the squares of ITS chunks (the same chunks it
                // stores below), the partials meet through the staging tiles (idle until the chunk loop), and half 0 publishes
                // max(sqrt(sum), eps) per row.  The bias comes in as float4 (the per-element predicated LDG chain of the first
                // version cost 0.33 ms of a 0.44 ms launch at the descriptor head: ncu, long-scoreboard on every FADD).
                float ss = 0.f;
                const bool bias_vec = biasp && ((reinterpret_cast<uintptr_t>(biasp) & 15) == 0);
                for (int c = half * 32; c < BN; c += 64) {
                    if (n0 + c >= p.N) break;
                    uint32_t v[32];
                    tmem_ld32(taddr + c, v);
                    if (n0 + c + 32 <= p.N && (bias_vec || !biasp)) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 bb = biasp ? __ldg(reinterpret_cast<const float4*>(biasp + n0 + c) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float f0 = __uint_as_float(v[4 * j4]) + bb.x, f1 = __uint_as_float(v[4 * j4 + 1]) + bb.y;
                            const float f2 = __uint_as_float(v[4 * j4 + 2]) + bb.z, f3 = __uint_as_float(v[4 * j4 + 3]) + bb.w;
                            ss += f0 * f0; ss += f1 * f1; ss += f2 * f2; ss += f3 * f3;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = n0 + c + j;
                            if (n < p.N) {
                                const float f = __uint_as_float(v[j]) + (biasp ? __ldg(biasp + n) : 0.f);
                                ss += f * f;
                            }
                        }
                    }
                }
                sts32(tile_a + 4 * lane, __float_as_int(ss));
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                if (half == 0) {
                    const float other = __int_as_float(lds32(tile_a + 4u * 4096u + 4 * lane));  // the partner warp's tile
                    sts32(eq_inv + 4 * lane, __float_as_int(fmaxf(sqrtf(ss + other), 1e-12f)));
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            }
            if constexpr (MODE == 2) {
                // ---- dedicated qkv epilogue: ONE pass in the bf16 mapping (4 lanes per row, 8 columns per lane): bias, rotary on
                // adjacent pairs, scale, hi / lo split, 16-byte stores into the [B][heads][n][64] operand tensors.  The generic
                // two-pass loop below spends ~2000 SASS instructions per 32 x 32 chunk on runtime feature checks (ReLU, L2 norm,
                // fp32 / row / phase-split outputs, ragged N) that never apply here; this path needs ~300 (ncu: the kernel was
                // bound by the issue of its own epilogue, tensor pipe 30 %).  N is 768 or 512: every chunk is full.
                const bool is_v = (p.qkv_mode == 1) ? (nt == 2) : (nt == 1);
                const bool rot = (p.qkv_mode == 1) && !is_v;
                __nv_bfloat16* qh = is_v ? p.v_hi : (nt == 0 ? p.q_hi : p.k_hi);
                __nv_bfloat16* ql = is_v ? p.v_lo : (nt == 0 ? p.q_lo : p.k_lo);
                const float sc = is_v ? 1.f : p.qk_scale;
                for (int c = half * 32; c < BN; c += 64) {
                    const int nb = n0 + c;
                    {
                        uint32_t v[32];
                        tmem_ld32(taddr + c, v);
                        const uint32_t trow = tile_a + (uint32_t)lane * 128u;
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4)
                            sts128(trow + (uint32_t)((j4 ^ (lane & 7)) << 4), make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                                                                                  __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])));
                    }
                    const float4 b0 = bn0, b1 = bn1;
                    if (c + 64 < BN) load_bias(nb + 64);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = i * 8 + g2;
                        const int pixr = lds32(eq_pix + 4 * row);
                        const uint32_t tr = tile_a + (uint32_t)row * 128u;
                        const float4 a = lds128(tr + (uint32_t)((((col2 >> 2)) ^ (row & 7)) << 4)), bq = lds128(tr + (uint32_t)((((col2 >> 2) + 1) ^ (row & 7)) << 4));
                        float f[8] = {a.x + b0.x, a.y + b0.y, a.z + b0.z, a.w + b0.w, bq.x + b1.x, bq.y + b1.y, bq.z + b1.z, bq.w + b1.w};
                        if (rot) {
                            const float cs[4] = {rot_c[i].x, rot_c[i].y, rot_c[i].z, rot_c[i].w};
                            const float sn[4] = {rot_s[i].x, rot_s[i].y, rot_s[i].z, rot_s[i].w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float e = f[2 * k], o = f[2 * k + 1];
                                f[2 * k] = e * cs[k] + (-o) * sn[k];
                                f[2 * k + 1] = o * cs[k] + e * sn[k];
                            }
                        }
                        uint32_t hi[4], lo[4];
                        if (is_v && p.v_f16) {  // V as IEEE fp16 hi / lo planes: operand of the fp16-probability PV (attention_tc.cu, P16)
#pragma unroll
                            for (int k = 0; k < 8; k += 2) {
                                const __half2 h2 = __floats2half2_rn(f[k], f[k + 1]);
                                const float2 hf = __half22float2(h2);
                                const __half2 l2 = __floats2half2_rn(f[k] - hf.x, f[k + 1] - hf.y);
                                hi[k >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                                lo[k >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; k += 2) {
                                const float x0 = f[k] * sc, x1 = f[k + 1] * sc;
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
                                const uint32_t u = *reinterpret_cast<uint32_t*>(&h2);
                                __nv_bfloat162 l2 = __floats2bfloat162_rn(x0 - __uint_as_float(u << 16), x1 - __uint_as_float(u & 0xffff0000u));
                                hi[k >> 1] = u;
                                lo[k >> 1] = *reinterpret_cast<uint32_t*>(&l2);
                            }
                        }
                        if (pixr >= 0) {
                            const long long qo = (long long)lds32(eq_ps + 4 * row) + (long long)(c >> 6) * lds32(eq_inv + 4 * row) + (c & 63) + col2;
                            *reinterpret_cast<uint4*>(qh + qo) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            if (SPLIT == 3) *reinterpret_cast<uint4*>(ql + qo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                    __syncwarp();
                }
            } else
            for (int c = half * 32; c < BN; c += 64) {
                if (n0 + c >= p.N) break;  // warp-uniform
                const int nb = n0 + c;
                const bool full32 = (nb + 32 <= p.N);
                {
                    uint32_t v[32];
                    tmem_ld32(taddr + c, v);
                    const uint32_t trow = tile_a + (uint32_t)lane * 128u;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)  // 16-byte chunk j4 of row `lane` lives at chunk position j4 ^ (lane & 7)
                        sts128(trow + (uint32_t)((j4 ^ (lane & 7)) << 4), make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                                                                              __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])));
                }
                if (fast_ok && full32) {
                    // ---- ONE pass in the 16-byte-store mapping (4 lanes per row, 8 columns per lane): bias, residual, ReLU, fp32 /
                    // split-bf16 / phase-split / fp16 stores.  Covers every layer of the networks except ragged N;
                    // the generic two-pass loop below costs ~3x the instructions per chunk (ncu: the 1x1 ResBlock convolutions
                    // were bound by the issue of their own epilogue, short-scoreboard chains on the staging tile).
                    uint4 rcur[MODE == 1 ? 8 : 1];
                    if constexpr (MODE == 1) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) rcur[i] = rv[i];
                        if (c + 64 < BN && nb + 64 + 32 <= p.N) load_res_fast(nb + 64, rv);  // next chunk's residual stays in flight
                    }
                    const float4 b0 = bn0, b1 = bn1;
                    if (c + 64 < BN) load_bias(nb + 64);
                    const bool want_planes = p.out_hi || p.ps_hi;
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = i * 8 + g2;
                        const int pixr = lds32(eq_pix + 4 * row);
                        const uint32_t tr = tile_a + (uint32_t)row * 128u;
                        const float4 a = lds128(tr + (uint32_t)((((col2 >> 2)) ^ (row & 7)) << 4)), bq = lds128(tr + (uint32_t)((((col2 >> 2) + 1) ^ (row & 7)) << 4));
                        float f[8] = {a.x + b0.x, a.y + b0.y, a.z + b0.z, a.w + b0.w, bq.x + b1.x, bq.y + b1.y, bq.z + b1.z, bq.w + b1.w};
                        if constexpr (MODE == 1) {
                            const uint4 ra = rcur[2 * i], rb = rcur[2 * i + 1];
                            if (resp) {
                                f[0] += __uint_as_float(ra.x); f[1] += __uint_as_float(ra.y); f[2] += __uint_as_float(ra.z); f[3] += __uint_as_float(ra.w);
                                f[4] += __uint_as_float(rb.x); f[5] += __uint_as_float(rb.y); f[6] += __uint_as_float(rb.z); f[7] += __uint_as_float(rb.w);
                            } else {
                                const uint32_t hw[4] = {ra.x, ra.y, ra.z, ra.w}, lw[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    f[2 * k] += __uint_as_float(hw[k] << 16) + __uint_as_float(lw[k] << 16);
                                    f[2 * k + 1] += __uint_as_float(hw[k] & 0xffff0000u) + __uint_as_float(lw[k] & 0xffff0000u);
                                }
                            }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
                        }
                        if (p.l2norm) {
                            const float d = __int_as_float(lds32(eq_inv + 4 * row));
#pragma unroll
                            for (int k = 0; k < 8; ++k) f[k] = f[k] / d;
                        }
                        if (pixr < 0) continue;
                        if (p.out_f32) {
                            float4* op = reinterpret_cast<float4*>(p.out_f32 + (long long)pixr * p.ld_f32 + nb + col2);
                            op[0] = make_float4(f[0], f[1], f[2], f[3]);
                            op[1] = make_float4(f[4], f[5], f[6], f[7]);
                        }
                        if (p.out_h16 || (SPLIT == 1 && p.f16 && want_planes)) {
                            uint32_t hh[4];
#pragma unroll
                            for (int k = 0; k < 8; k += 2) {
                                const __half2 h2 = __floats2half2_rn(f[k], f[k + 1]);
                                hh[k >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                            }
                            const uint4 hv = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                            if (p.out_h16) *reinterpret_cast<uint4*>(p.out_h16 + (long long)pixr * p.ld_bf + nb + col2) = hv;
                            if (SPLIT == 1 && p.f16) {  // single-pass fp16 layer: the hi plane carries fp16, no lo plane
                                if (p.out_hi) *reinterpret_cast<uint4*>(p.out_hi + (long long)pixr * p.ld_bf + nb + col2) = hv;
                                if (p.ps_hi) *reinterpret_cast<uint4*>(p.ps_hi + (long long)lds32(eq_ps + 4 * row) * p.ld_ps + nb + col2) = hv;
                                continue;
                            }
                        }
                        if (want_planes) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int k = 0; k < 8; k += 2) {
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(f[k], f[k + 1]);
                                const uint32_t u = *reinterpret_cast<uint32_t*>(&h2);
                                __nv_bfloat162 l2 = __floats2bfloat162_rn(f[k] - __uint_as_float(u << 16), f[k + 1] - __uint_as_float(u & 0xffff0000u));
                                hi[k >> 1] = u;
                                lo[k >> 1] = *reinterpret_cast<uint32_t*>(&l2);
                            }
                            const uint4 hv = make_uint4(hi[0], hi[1], hi[2], hi[3]), lv = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            if (p.out_hi) {
                                *reinterpret_cast<uint4*>(p.out_hi + (long long)pixr * p.ld_bf + nb + col2) = hv;
                                if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + (long long)pixr * p.ld_bf + nb + col2) = lv;
                            }
                            if (p.ps_hi) {
                                const long long pp = lds32(eq_ps + 4 * row);
                                *reinterpret_cast<uint4*>(p.ps_hi + pp * p.ld_ps + nb + col2) = hv;
                                if (p.ps_lo) *reinterpret_cast<uint4*>(p.ps_lo + pp * p.ld_ps + nb + col2) = lv;
                            }
                        }
                    }
                    __syncwarp();
                    continue;
                }
                uint4 rcur[MODE == 1 ? 8 : 1];
                if constexpr (MODE == 1) load_res(nb, rcur);  // generic path (ragged N, L2 norm, unaligned rows): loaded in place
                __syncwarp();
                // ---- pass 1 (fp32 mapping: 8 lanes per row): bias, residual, ReLU, L2 norm, fp32 store ----
                {
                    float bz[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) bz[k] = (biasp && nb + col1 + k < p.N) ? __ldg(biasp + nb + col1 + k) : 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = i * 4 + g1;
                        const int pixr = lds32(eq_pix + 4 * row);
                        const uint32_t tp = tile_a + (uint32_t)row * 128u + (uint32_t)((((col1 >> 2) ^ (row & 7))) << 4);
                        const float4 tv = lds128(tp);
                        float f[4] = {tv.x + bz[0], tv.y + bz[1], tv.z + bz[2], tv.w + bz[3]};
                        if constexpr (MODE == 1) { const float4 rr = res_value(rcur[i]); f[0] += rr.x; f[1] += rr.y; f[2] += rr.z; f[3] += rr.w; }
                        if (p.relu) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) f[k] = fmaxf(f[k], 0.f);
                        }
                        if (p.l2norm) {
                            const float d = __int_as_float(lds32(eq_inv + 4 * row));
#pragma unroll
                            for (int k = 0; k < 4; ++k) f[k] = f[k] / d;
                        }
                        if (p.out_f32 && pixr >= 0) {
                            float* op = p.out_f32 + (long long)pixr * p.ld_f32 + nb + col1;
                            if (full32 && ((p.ld_f32 & 3) == 0)) {
                                *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
                            } else {
#pragma unroll
                                for (int k = 0; k < 4; ++k) if (nb + col1 + k < p.N) op[k] = f[k];
                            }
                        }
                        if (p.out_hi || p.ps_hi || p.out_h16) sts128(tp, make_float4(f[0], f[1], f[2], f[3]));
                    }
                }
                // ---- pass 2 (bf16 mapping: 4 lanes per row): split into hi / lo planes, 16-byte stores ----
                if (p.out_hi || p.ps_hi || p.out_h16) {  // requires N % 32 == 0 (checked on the host)
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = i * 8 + g2;
                        const int pixr = lds32(eq_pix + 4 * row);
                        uint32_t hi[4], lo[4], hh[4] = {0u, 0u, 0u, 0u};
                        {
                            const uint32_t tr = tile_a + (uint32_t)row * 128u;
                            const float4 a = lds128(tr + (uint32_t)((((col2 >> 2)) ^ (row & 7)) << 4)), bq = lds128(tr + (uint32_t)((((col2 >> 2) + 1) ^ (row & 7)) << 4));
                            const float fv[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                            for (int k = 0; k < 8; k += 2) {
                                if (SPLIT == 1 && p.f16) {  // single fp16 plane
                                    __half2 h2 = __floats2half2_rn(fv[k], fv[k + 1]);
                                    hi[k >> 1] = *reinterpret_cast<uint32_t*>(&h2);
                                    lo[k >> 1] = 0u;
                                    continue;
                                }
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(fv[k], fv[k + 1]);
                                const uint32_t u = *reinterpret_cast<uint32_t*>(&h2);
                                __nv_bfloat162 l2 = __floats2bfloat162_rn(fv[k] - __uint_as_float(u << 16),
                                                                          fv[k + 1] - __uint_as_float(u & 0xffff0000u));
                                hi[k >> 1] = u;
                                lo[k >> 1] = *reinterpret_cast<uint32_t*>(&l2);
                            }
                            if (p.out_h16) {
#pragma unroll
                                for (int k = 0; k < 8; k += 2) {
                                    const __half2 h2 = __floats2half2_rn(fv[k], fv[k + 1]);
                                    hh[k >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                                }
                            }
                        }
                        if (pixr >= 0) {
                            if (p.out_h16) {
                                *reinterpret_cast<uint4*>(p.out_h16 + (long long)pixr * p.ld_bf + nb + col2) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                            }
                            if (p.out_hi) {
                                *reinterpret_cast<uint4*>(p.out_hi + (long long)pixr * p.ld_bf + nb + col2) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                                if (p.out_lo)
                                    *reinterpret_cast<uint4*>(p.out_lo + (long long)pixr * p.ld_bf + nb + col2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            }
                            if (p.ps_hi) {
                                const long long pp = lds32(eq_ps + 4 * row);
                                *reinterpret_cast<uint4*>(p.ps_hi + pp * p.ld_ps + nb + col2) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                                if (p.ps_lo)
                                    *reinterpret_cast<uint4*>(p.ps_lo + pp * p.ld_ps + nb + col2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            // both warps of the quarter are done with the row table and the accumulator
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            if (lane == 0) mbar_arrive(&tempty[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // no CTA exits while the peer may still multicast into it / arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS));
    }
}

// ---------------------------------------------------------------------------------------- host
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

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return PRAM_ERR_CUDA;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PRAM_OK : PRAM_ERR_CUDA;
}

static int g_num_sms = 0;

template <int BN, int SPLIT, int MODE, int CL>
static int launch_mode(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl,
                       const Args& a, int m_tiles, int n_tiles, cudaStream_t stream) {
    using C = Cfg<BN, SPLIT>;
    auto kern = gemm_tc_kernel<BN, SPLIT, MODE, CL>;
    static bool attr = false;
    static int max_clusters = 0;
    if (!attr) {
        PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr = true;
    }
    const int items = ((m_tiles + CL - 1) / CL) * n_tiles;
    if (CL == 1) {
        int grid = items < g_num_sms ? items : g_num_sms;
        PRAM_CUDA(pram_launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), C::SMEM_BYTES, stream, ah, al, wh, wl, a));
    } else {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.blockDim = dim3(NUM_THREADS, 1, 1);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cfg.stream = stream;
        cfg.attrs = at; cfg.numAttrs = g_pram_pdl ? 2 : 1;
        if (!max_clusters) {
            cfg.gridDim = dim3(g_num_sms / CL * CL, 1, 1);
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = g_num_sms / CL; }
            max_clusters = n;
        }
        const int ncl = items < max_clusters ? items : max_clusters;
        cfg.gridDim = dim3(ncl * CL, 1, 1);
        PRAM_CUDA(cudaLaunchKernelEx(&cfg, kern, ah, al, wh, wl, a));
    }
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

template <int BN, int SPLIT, int CL>
static int launch(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl,
                  const Args& a, int m_tiles, int n_tiles, cudaStream_t stream) {
    if (a.qkv_mode) {
        if constexpr (BN == 256) return launch_mode<BN, SPLIT, 2, CL>(ah, al, wh, wl, a, m_tiles, n_tiles, stream);
        else return PRAM_ERR_UNSUPPORTED;
    }
    if (a.res || a.res_hi) return launch_mode<BN, SPLIT, 1, CL>(ah, al, wh, wl, a, m_tiles, n_tiles, stream);
    return launch_mode<BN, SPLIT, 0, CL>(ah, al, wh, wl, a, m_tiles, n_tiles, stream);
}

}  // namespace tc

// Public argument block of pram_gemm_tc (mirrored by ctypes in pram_b200/_lib.py).
struct pram_tc_args {
    const void* a_hi; const void* a_lo;   // bf16 activations [planes][in_H][in_W][a_ld], lo may be NULL (split=1)
    long long a_ld;                       // elements between consecutive pixels
    int in_W, in_H, in_planes, Cin;
    const void* w_hi; const void* w_lo;   // bf16 weights [w_planes][N][Cin]
    int w_planes;
    int B, Ho, Wo, N;
    int tw_log2;
    int ntaps;
    int tap_dx[9], tap_dy[9], tap_plane[9];
    int planes_per_image;
    int w_batch_mult;
    const float* bias; const float* res; long long res_ld; int relu;
    float* out_f32; long long ld_f32;
    void* out_hi; void* out_lo; long long ld_bf;
    void* ps_hi; void* ps_lo; long long ld_ps;
    int l2norm;
    int split;                            // 1: bf16, 3: error-compensated bf16x3
    int bn;                               // 0 = auto
    // fused attention-operand epilogue (see tc::Args)
    int qkv_mode; const float* cosb; const float* sinb; float qk_scale;
    void* q_hi; void* q_lo; void* k_hi; void* k_lo; void* v_hi; void* v_lo;
    int seg_split, seg_n0, seg_n1, heads;
    int cluster;                          // 0 = auto, 1 = single CTAs, 2 = 2-CTA clusters with a multicast weight tile
    int l2_prefetch;                      // 1 = pull the next tile's activation boxes into L2 one tile ahead (single-tap layers); default off
    int f16;                              // 1 (split == 1 only): a / w planes hold IEEE fp16, out_hi receives fp16 -- single-pass fp16 mode
    const void* res_hi; const void* res_lo;  // residual as split-bf16 planes (used when res == NULL; row stride res_ld, N % 32 == 0)
    int v_f16;                            // qkv epilogue: v_hi / v_lo receive IEEE fp16 planes
    void* out_h16;                        // optional extra output: one IEEE fp16 plane, row stride ld_bf (N % 32 == 0)
};

PRAM_API int pram_gemm_tc(const pram_tc_args* a, cudaStream_t stream) {
    using namespace tc;
    if (!a || !a->a_hi || !a->w_hi || a->B <= 0 || a->N <= 0 || a->ntaps <= 0 || a->ntaps > MAX_TAPS) return PRAM_ERR_ARG;
    if (a->split != 1 && a->split != 3) return PRAM_ERR_ARG;
    if (a->split == 3 && (!a->a_lo || !a->w_lo)) return PRAM_ERR_ARG;
    if ((a->a_ld % 8) || (a->Cin % 8)) return PRAM_ERR_UNSUPPORTED;  // 16-byte TMA strides
    if ((a->out_hi || a->ps_hi) && ((a->N % 32) || (a->ld_bf % 8) || (a->ps_hi && (a->ld_ps % 8)))) return PRAM_ERR_UNSUPPORTED;
    if (g_num_sms == 0) {
        int dev = 0;
        PRAM_CUDA(cudaGetDevice(&dev));
        PRAM_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int bn = a->bn;
    if (bn == 0) bn = (a->N > 128) ? 256 : (a->N > 64 ? 128 : 64);
    const int TW = 1 << a->tw_log2, TH = BM >> a->tw_log2;
    if (TW > 256 || TH < 1) return PRAM_ERR_ARG;
    if (a->bn == 0 && bn == 256 && !a->qkv_mode && !a->l2norm) {
        // Wave quantisation of small problems: one 640x480 frame has 150 output tiles per 256-wide layer at 1/4 resolution --
        // two waves on 148 SMs, the second one with 2 tiles.  Half-width tiles (0.55 of the cost: half the MMAs, the same
        // activation tile) come out ahead whenever they need fewer weighted waves; large batches keep the 256-wide tile.
        const long long mt = (long long)a->B * ((a->Wo + TW - 1) / TW) * ((a->Ho + TH - 1) / TH);
        const long long w256 = (mt * ((a->N + 255) / 256) + g_num_sms - 1) / g_num_sms;
        const long long w128 = (mt * ((a->N + 127) / 128) + g_num_sms - 1) / g_num_sms;
        if (0.55 * (double)w128 < (double)w256 && w256 <= 8) bn = 128;
    }
    if (a->l2norm && a->N > bn) return PRAM_ERR_UNSUPPORTED;

    // 2-CTA clusters with a multicast weight tile: shared weights (no per-batch weight planes), at least one full
    // wave of tiles, and a weight tile whose halves are whole swizzle atoms
    const int tiles_x = (a->Wo + TW - 1) / TW, tiles_y = (a->Ho + TH - 1) / TH;
    const int m_tiles = a->B * tiles_x * tiles_y, n_tiles = (a->N + bn - 1) / bn;
    int cl = 1;  // measured on B200 (profiles/README.md): no gain from the multicast pair -- the kernel is tensor-issue bound, not L2 bound
    if (a->cluster == 1 || a->cluster == 2) cl = (a->w_batch_mult == 0 && bn >= 128) ? a->cluster : 1;
    CUtensorMap ah, al, wh, wl;
    {
        cuuint64_t dims[4] = {(cuuint64_t)a->Cin, (cuuint64_t)a->in_W, (cuuint64_t)a->in_H, (cuuint64_t)a->in_planes};
        cuuint64_t str[3] = {(cuuint64_t)a->a_ld * 2, (cuuint64_t)a->a_ld * 2 * a->in_W,
                             (cuuint64_t)a->a_ld * 2 * a->in_W * a->in_H};
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
        int rc = encode(&ah, a->a_hi, 4, dims, str, box);
        if (rc) return rc;
        rc = encode(&al, a->a_lo ? a->a_lo : a->a_hi, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)a->Cin, (cuuint64_t)a->N, (cuuint64_t)a->w_planes};
        cuuint64_t str[2] = {(cuuint64_t)a->Cin * 2, (cuuint64_t)a->Cin * 2 * a->N};
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(bn / cl), 1};
        int rc = encode(&wh, a->w_hi, 3, dims, str, box);
        if (rc) return rc;
        rc = encode(&wl, a->w_lo ? a->w_lo : a->w_hi, 3, dims, str, box);
        if (rc) return rc;
    }
    Args k;
    k.B = a->B; k.Ho = a->Ho; k.Wo = a->Wo; k.N = a->N; k.tw_log2 = a->tw_log2; k.ntaps = a->ntaps;
    k.kblocks = (a->Cin + BK - 1) / BK;
    k.planes_per_image = a->planes_per_image;
    for (int i = 0; i < MAX_TAPS; ++i) { k.tap_dx[i] = a->tap_dx[i]; k.tap_dy[i] = a->tap_dy[i]; k.tap_plane[i] = a->tap_plane[i]; }
    k.w_batch_mult = a->w_batch_mult;
    k.bias = a->bias; k.res = a->res; k.res_ld = a->res_ld; k.relu = a->relu;
    k.out_f32 = a->out_f32; k.ld_f32 = a->ld_f32;
    k.out_hi = (__nv_bfloat16*)a->out_hi; k.out_lo = (__nv_bfloat16*)a->out_lo; k.ld_bf = a->ld_bf;
    k.ps_hi = (__nv_bfloat16*)a->ps_hi; k.ps_lo = (__nv_bfloat16*)a->ps_lo; k.ld_ps = a->ld_ps;
    k.l2norm = a->l2norm;
    k.qkv_mode = a->qkv_mode; k.cosb = a->cosb; k.sinb = a->sinb; k.qk_scale = a->qk_scale;
    k.q_hi = (__nv_bfloat16*)a->q_hi; k.q_lo = (__nv_bfloat16*)a->q_lo; k.k_hi = (__nv_bfloat16*)a->k_hi;
    k.k_lo = (__nv_bfloat16*)a->k_lo; k.v_hi = (__nv_bfloat16*)a->v_hi; k.v_lo = (__nv_bfloat16*)a->v_lo;
    k.seg_split = a->seg_split; k.seg_n0 = a->seg_n0; k.seg_n1 = a->seg_n1; k.heads = a->heads;
    if (a->f16 && a->split != 1) return PRAM_ERR_ARG;
    k.f16 = a->f16;
    k.pred = g_pram_pred;
    k.pdl_early = g_pram_pdl >= 2;
    k.v_f16 = a->v_f16;
    k.out_h16 = (__half*)a->out_h16;
    if (a->out_h16 && ((a->N % 32) || (a->ld_bf % 8))) return PRAM_ERR_UNSUPPORTED;
    k.res_hi = (const __nv_bfloat16*)a->res_hi; k.res_lo = (const __nv_bfloat16*)a->res_lo;
    if (!a->res && a->res_hi && ((a->N % 32) || (a->res_ld % 4))) return PRAM_ERR_UNSUPPORTED;
    k.l2_prefetch = (a->ntaps == 1) && (a->l2_prefetch == 1);  // measured on B200: no gain (the thin GEMMs are store-bound), off unless asked for
    if (a->qkv_mode) {
        if (bn != 256 || a->heads * 64 != 256 || !a->q_hi || !a->v_hi || a->ps_hi || a->l2norm) return PRAM_ERR_UNSUPPORTED;
        if ((a->qkv_mode == 1 && (a->N != 768 || !a->k_hi || !a->cosb || !a->sinb)) || (a->qkv_mode == 2 && a->N != 512)) return PRAM_ERR_ARG;
        if (a->seg_n0 <= 0 || a->seg_n1 <= 0) return PRAM_ERR_ARG;
    }
    if (a->split == 3) {
        if (bn == 256) return cl == 2 ? launch<256, 3, 2>(ah, al, wh, wl, k, m_tiles, n_tiles, stream) : launch<256, 3, 1>(ah, al, wh, wl, k, m_tiles, n_tiles, stream);
        if (bn == 128) return cl == 2 ? launch<128, 3, 2>(ah, al, wh, wl, k, m_tiles, n_tiles, stream) : launch<128, 3, 1>(ah, al, wh, wl, k, m_tiles, n_tiles, stream);
        return launch<64, 3, 1>(ah, al, wh, wl, k, m_tiles, n_tiles, stream);
    }
    if (bn == 256) return cl == 2 ? launch<256, 1, 2>(ah, al, wh, wl, k, m_tiles, n_tiles, stream) : launch<256, 1, 1>(ah, al, wh, wl, k, m_tiles, n_tiles, stream);
    if (bn == 128) return cl == 2 ? launch<128, 1, 2>(ah, al, wh, wl, k, m_tiles, n_tiles, stream) : launch<128, 1, 1>(ah, al, wh, wl, k, m_tiles, n_tiles, stream);
    return launch<64, 1, 1>(ah, al, wh, wl, k, m_tiles, n_tiles, stream);
}

// fp32 -> split bf16 planes (hi = bf16(x), lo = bf16(x - hi)); used for weights (once) and for tensors
// produced by CUDA-core kernels that feed a tensor-core GEMM.
__global__ void split_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    if (i + 4 <= n) {
        float4 v = *reinterpret_cast<const float4*>(in + i);
        __nv_bfloat16 h[4], l[4];
        tc::split_bf16(v.x, h[0], l[0]); tc::split_bf16(v.y, h[1], l[1]);
        tc::split_bf16(v.z, h[2], l[2]); tc::split_bf16(v.w, h[3], l[3]);
        *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<uint2*>(h);
        if (lo) *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<uint2*>(l);
    } else {
        for (; i < n; ++i) {
            __nv_bfloat16 h, l;
            tc::split_bf16(in[i], h, l);
            hi[i] = h;
            if (lo) lo[i] = l;
        }
    }
}

__global__ void cast_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    if (i + 8 <= n) {
        const float4 a = *reinterpret_cast<const float4*>(in + i), b = *reinterpret_cast<const float4*>(in + i + 4);
        __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y), __floats2half2_rn(b.z, b.w)};
        *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<uint4*>(h);
    } else {
        for (; i < n; ++i) out[i] = __float2half_rn(in[i]);
    }
}

// fp32 -> IEEE fp16 (operand plane of the single-pass fp16 GEMM mode); n % 8 == 0 rows are moved 32 B -> 16 B per thread
PRAM_API int pram_cast_f16(const float* in, void* out, long long n, cudaStream_t stream) {
    if (!in || !out || n <= 0) return PRAM_ERR_ARG;
    cast_f16_kernel<<<cdiv((n + 7) / 8, 256), 256, 0, stream>>>(in, (__half*)out, n);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

__global__ void split_f16_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = in[i];
    const __half h = __float2half_rn(x);
    hi[i] = h;
    if (lo) lo[i] = __float2half_rn(x - __half2float(h));
}

// fp32 -> IEEE fp16 hi / lo planes: hi = fp16(x), lo = fp16(x - hi) (lo may be NULL)
PRAM_API int pram_split_f16(const float* in, void* hi, void* lo, long long n, cudaStream_t stream) {
    if (!in || !hi || n <= 0) return PRAM_ERR_ARG;
    split_f16_kernel<<<cdiv(n, 256), 256, 0, stream>>>(in, (__half*)hi, (__half*)lo, n);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

PRAM_API int pram_split_bf16(const float* in, void* hi, void* lo, long long n, cudaStream_t stream) {
    if (!in || !hi || n <= 0 || (n % 4 && false)) return PRAM_ERR_ARG;
    split_bf16_kernel<<<cdiv((n + 3) / 4, 256), 256, 0, stream>>>(in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}
