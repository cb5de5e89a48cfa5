// One persistent tcgen05 kernel for the tail of every transformer block of SegNetViT / GML / AdaGML
// (reference nets/segnetvit.py:97-106, nets/gml.py:128-137, 164-186):
//
//     message = proj(ctx);  x_new = x + mlp.3( GELU( LayerNorm( mlp.0( [x | message] ) ) ) )
//
// `proj` is folded into mlp.0 on the host once (W1 = [W0x | W0m . Wp], b1 = b0 + W0m . bp, float64), so the kernel
// computes, per 128-token tile, with the tile resident on chip from the first MMA to the residual add:
//
//   phase A   h[128 x 512]  = [x | ctx] . W1^T          two N halves of 256, fp32 accumulators = ALL 512 TMEM columns
//   phase B   LayerNorm statistics (two exact passes over TMEM), then per 64-column k-block:
//             y = GELU(LN(h + b1)) -> split bf16 (hi, lo) written as a SWIZZLE_128B K-major UMMA operand into a 4-slot
//             shared-memory ring; the MMA warp consumes the ring against TMA-streamed W3 tiles:
//             out[128 x 256] += y_kblock . W3_kblock^T   (N = 256 per MMA; accumulator re-uses TMEM columns 0..255 once k-blocks 0..3,
//             which live there, have been converted)
//   phase C   out + b3 + residual -> fp32 and split-bf16 rows of the next activation buffer
//
// Nothing of the 512-wide hidden activation ever touches HBM (the unfused path wrote it as fp32, re-read it for
// LayerNorm, wrote hi/lo planes and re-read those: 8 KB per token per block), and four launches become one.
//
// Warp roles (576 threads, one CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer (elect.sync per instruction),
// warps 2-17 = 16 epilogue warps (TMEM lane quarter = warp % 4, column part = (warp - 2) / 4).
// Shared memory is time-multiplexed: phase A uses two 96 KB operand stages, phase B the same bytes as ring + W3 stages.
// SPLIT = 3: error-compensated bf16x3 (hi.hi + lo.hi + hi.lo); SPLIT = 1: plain bf16.
#include "common.cuh"
#include <cuda.h>
#include <stdio.h>
#include <string.h>

namespace mb {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int D_IN = 512, D_HID = 512, D_OUT = 256;
constexpr int KB1 = D_IN / BK, KB2 = D_HID / BK;            // 8, 8
constexpr int EPI_WARPS = 16;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;             // 576
constexpr int A_BYTES = BM * BK * 2;                         // 16 KB: 128 rows x 128 B
constexpr int W1_BYTES = 256 * BK * 2;                       // 32 KB: one N half of W1, one k-block
constexpr int BK3 = 32;                                      // W3 is streamed in half k-blocks: N = 256 rows x 32 K columns
constexpr int W3_BYTES = 256 * BK3 * 2;                      // 16 KB per plane (SWIZZLE_64B rows of 64 B)
constexpr int STAGE_A = 2 * A_BYTES + 2 * W1_BYTES;          // 96 KB (both planes; SPLIT = 1 uses the hi halves only)
constexpr int NSTAGE_A = 2;
constexpr int SLOT = 2 * A_BYTES;                            // 32 KB ring slot: hi | lo
constexpr int NSLOT = 4;
constexpr int STAGE_3 = 2 * W3_BYTES;                        // 32 KB
constexpr int NSTAGE_3 = 3;
constexpr int RING_OFF = 0, W3_OFF = NSLOT * SLOT;           // 0, 128 KB
constexpr int REGION = W3_OFF + NSTAGE_3 * STAGE_3;          // 224 KB  (>= NSTAGE_A * STAGE_A = 192 KB)
constexpr int RED_OFF = RING_OFF + 3 * SLOT;                 // LN partial sums live in ring slot 3 (written last)
constexpr int BAR_OFF = REGION;                              // barriers (256 B)
constexpr int SMEM_BYTES = 1024 + REGION + 256;
static_assert(NSTAGE_A * STAGE_A <= REGION, "phase A stages must fit the multiplexed region");
static_assert(SMEM_BYTES <= 227 * 1024, "dynamic shared memory budget of sm_100a exceeded");

struct Args {
    int T;
    const float* res; long long res_ld;                          // fp32 residual rows, or NULL:
    const __nv_bfloat16* xa_hi; const __nv_bfloat16* xa_lo; long long lda;  // ... residual = hi + lo of the A rows (x)
    float* out_f32; long long ld_f32;
    __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; long long ld_bf;
    const int* pred;  // launch predicate (common.cuh)
    int pdl_early;    // 1: release the dependent launch right after this grid's own wait (common.cuh)
    long long* dbg;   // optional timeline: [CTA][tile iteration (<= 8)][32] SM clock stamps (tools/bench_block.py --timeline)
};
__device__ __forceinline__ void stamp(const Args& p, uint32_t it, int slot) {
    if (p.dbg && it < 8) p.dbg[((long long)blockIdx.x * 8 + it) * 32 + slot] = clock64();
}
// Per-column tables travel as a KERNEL PARAMETER (constant bank): every epilogue access is warp-uniform, so it becomes
// one LDC broadcast instead of a global load (ncu on the first version: the 12 LDG per 16 elements in front of each
// k-block were the longest stall of the LayerNorm pass).
struct alignas(16) Tables { float b1[D_HID]; float g[D_HID]; float be[D_HID]; float b3[D_OUT]; };

// ---------------------------------------------------------------------------------------- PTX (see gemm_tc.cu)
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int what) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            printf("pram mlp_block_tc: mbarrier timeout (block %d thread %d wait %d)\n", blockIdx.x, threadIdx.x, what);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major SWIZZLE_128B, SBO = 1024 B
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// K-major SWIZZLE_64B operand (rows of 64 B = 32 bf16): 8-row atoms of 512 B -> SBO = 512 B, layout type 4
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<uint32_t*>(&h);
    __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// exact (erf) GELU in 13 instructions: with erf(|x|) = 1 - P(t) exp(-x^2), t = 1 / (1 + p |x|) (Abramowitz-Stegun 7.1.26,
// abs error <= 1.5e-7) and x = y / sqrt(2):   GELU(y) = 0.5 y (1 + erf(x)) = max(y, 0) - |y| * [0.5 P(t)] * exp(-y^2 / 2)
// (both signs).  1/sqrt(2), 0.5 and log2(e) are folded into the constants; rcp / ex2 are single MUFU instructions.
__device__ __forceinline__ float gelu_fast(float y) {
    const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(y), 1.f));
    float pl = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    pl = fmaf(pl, t, 0.5f * 1.421413741f);
    pl = fmaf(pl, t, 0.5f * -0.284496736f);
    pl = fmaf(pl, t, 0.5f * 0.254829592f);
    pl *= t;
    const float ex = ex2_approx((y * -0.72134752044448170368f) * y);   // exp(-y^2 / 2)
    return fmaf(-fabsf(y), pl * ex, fmaxf(y, 0.f));
}
// ---- packed fp32 (FFMA2 / FMUL2 / FADD2: two lanes per instruction, sm_100) ---------------------------------------------
// The LayerNorm + GELU pass is bound by the issue of its own fp32 arithmetic (20 instructions per element); the packed
// forms halve every FFMA / FMUL / FADD of it.  Each lane is rounded exactly like the scalar instruction, and negating the
// polynomial's constants instead of |y| is exact too (round-to-nearest is symmetric), so the results are those of gelu_fast.
#ifndef PRAM_MLP_PACKED
#define PRAM_MLP_PACKED 1
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 pk2u(uint32_t a, uint32_t b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// GELU of two values (same formula and constants as gelu_fast)
__device__ __forceinline__ void gelu_fast2(f32x2 y, float& o0, float& o1) {
    const f32x2 ay = y & 0x7fffffff7fffffffull;
    float t0, t1;
    upk2(fma2(ay, pk2(0.3275911f * 0.70710678118654752440f, 0.3275911f * 0.70710678118654752440f), pk2(1.f, 1.f)), t0, t1);
    const f32x2 t = pk2(rcp_approx(t0), rcp_approx(t1));
    f32x2 pl = fma2(pk2(-0.5f * 1.061405429f, -0.5f * 1.061405429f), t, pk2(-0.5f * -1.453152027f, -0.5f * -1.453152027f));
    pl = fma2(pl, t, pk2(-0.5f * 1.421413741f, -0.5f * 1.421413741f));
    pl = fma2(pl, t, pk2(-0.5f * -0.284496736f, -0.5f * -0.284496736f));
    pl = fma2(pl, t, pk2(-0.5f * 0.254829592f, -0.5f * 0.254829592f));
    pl = mul2(pl, t);                                                   // -(0.5 P(t))
    float e0, e1;
    upk2(mul2(mul2(y, pk2(-0.72134752044448170368f, -0.72134752044448170368f)), y), e0, e1);
    const f32x2 ex = pk2(ex2_approx(e0), ex2_approx(e1));             // exp(-y^2 / 2)
    float y0, y1;
    upk2(y, y0, y1);
    upk2(fma2(ay, mul2(pl, ex), pk2(fmaxf(y0, 0.f), fmaxf(y1, 0.f))), o0, o1);
}
// erf by Abramowitz-Stegun 7.1.26 (abs error <= 1.5e-7), same routine as the split-bf16 LayerNorm kernel (nn_simt.cu)
__device__ __forceinline__ float erf_as(float x) {
    const float z = fabsf(x);
    const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float y = 1.f - p * t * __expf(-z * z);
    return copysignf(y, x);
}

// barrier block layout (uint64 each)
enum : int { B_FULLA = 0, B_EMPTYA = 2, B_HFULL = 4, B_SLOTF = 5, B_SLOTE = 9, B_W3F = 13, B_W3E = 16, B_ACCF = 19, B_ACCE = 20,
             B_H1FULL = 21, B_CDONE = 22, B_COUNT = 23 };

template <int SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_block_kernel(
    const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
    const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
    const __grid_constant__ CUtensorMap map_w3_hi, const __grid_constant__ CUtensorMap map_w3_lo, const Args p,
    const __grid_constant__ Tables tb) {
    if (p.pred) { pram_pdl_wait(); if (pram_pred_skip(p.pred)) return; }  // the flag is written by a predecessor kernel
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (p.T + BM - 1) / BM;
    constexpr uint32_t A_TX = (SPLIT == 3 ? 2u : 1u) * (A_BYTES + W1_BYTES);
    constexpr uint32_t W3_TX = (SPLIT == 3 ? 2u : 1u) * W3_BYTES;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a_hi); prefetch_tmap(&map_w1_hi); prefetch_tmap(&map_w3_hi);
        if (SPLIT == 3) { prefetch_tmap(&map_a_lo); prefetch_tmap(&map_w1_lo); prefetch_tmap(&map_w3_lo); }
        for (int s = 0; s < NSTAGE_A; ++s) { mbar_init(&bars[B_FULLA + s], 1); mbar_init(&bars[B_EMPTYA + s], 1); }
        mbar_init(&bars[B_HFULL], 1);
        mbar_init(&bars[B_H1FULL], 1);
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&bars[B_SLOTF + s], EPI_WARPS); mbar_init(&bars[B_SLOTE + s], 1); }
        for (int s = 0; s < NSTAGE_3; ++s) { mbar_init(&bars[B_W3F + s], 1); mbar_init(&bars[B_W3E + s], 1); }
        mbar_init(&bars[B_ACCF], 1);
        mbar_init(&bars[B_ACCE], EPI_WARPS);
        mbar_init(&bars[B_CDONE], EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    pram_pdl_wait();     // programmatic dependent launch: everything above overlapped the predecessor's drain
    if (p.pdl_early) pram_pdl_trigger();  // all CTAs of a persistent grid are resident: the successor may be scheduled as SMs free up

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int sa = 0; uint32_t pha = 0;   // phase A ring position (runs across tiles)
            int s3 = 0; uint32_t ph3 = 0;   // W3 ring position
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                // the previous tile's second GEMM has retired: ring slots and W3 stages (which alias these stages) are dead
                mbar_wait(&bars[B_ACCF], (it & 1) ^ 1, 0);
                const int row0 = tile * BM;
                for (int s = 0; s < 2 * KB1; ++s) {
                    const int half = (s < KB1) ? 1 : 0, kb = s & (KB1 - 1);  // N half 1 first: its TMEM columns free up first
                    mbar_wait(&bars[B_EMPTYA + sa], pha ^ 1, 1);
                    const uint32_t st = sbase + sa * STAGE_A;
                    mbar_expect_tx(&bars[B_FULLA + sa], A_TX);
                    tma_load_2d(st, &map_a_hi, &bars[B_FULLA + sa], kb * BK, row0);
                    tma_load_2d(st + 2 * A_BYTES, &map_w1_hi, &bars[B_FULLA + sa], 0, kb * D_HID + half * 256);
                    if (SPLIT == 3) {
                        tma_load_2d(st + A_BYTES, &map_a_lo, &bars[B_FULLA + sa], kb * BK, row0);
                        tma_load_2d(st + 2 * A_BYTES + W1_BYTES, &map_w1_lo, &bars[B_FULLA + sa], 0, kb * D_HID + half * 256);
                    }
                    if (++sa == NSTAGE_A) { sa = 0; pha ^= 1; }
                }
                // the first GEMM has retired: its stages may be overwritten by the W3 tiles
                mbar_wait(&bars[B_HFULL], it & 1, 2);
                // pull the NEXT tile's activation rows into L2 while this tile's epilogue / second GEMM run: its first operand
                // stages then land in ~1.5k cycles instead of a DRAM round trip on the critical path between two tiles
                if (tile + (int)gridDim.x < ntiles) {
                    const int nrow0 = (tile + (int)gridDim.x) * BM;
                    for (int kb = 0; kb < KB1; ++kb) {
                        tma_prefetch_2d(&map_a_hi, kb * BK, nrow0);
                        if (SPLIT == 3) tma_prefetch_2d(&map_a_lo, kb * BK, nrow0);
                    }
                }
                // the last W3 stage doubles as the staging area of the previous tile's phase C (2 KB per epilogue warp)
                mbar_wait(&bars[B_CDONE], (it & 1) ^ 1, 13);
                for (int s = 0; s < 2 * KB2; ++s) {   // 16 half k-blocks of W3, all 256 output rows each
                    mbar_wait(&bars[B_W3E + s3], ph3 ^ 1, 3);
                    const uint32_t st = sbase + W3_OFF + s3 * STAGE_3;
                    mbar_expect_tx(&bars[B_W3F + s3], W3_TX);
                    tma_load_2d(st, &map_w3_hi, &bars[B_W3F + s3], 0, s * D_OUT);
                    if (SPLIT == 3) tma_load_2d(st + W3_BYTES, &map_w3_lo, &bars[B_W3F + s3], 0, s * D_OUT);
                    if (++s3 == NSTAGE_3) { s3 = 0; ph3 ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, one elected lane per instruction) =====================
        constexpr uint32_t idesc1 = make_idesc(BM, 256);
        int sa = 0; uint32_t pha = 0;
        int s3 = 0; uint32_t ph3 = 0;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            // ---- phase A: h = [x | ctx] . W1^T, N half 1 -> columns 256..511, then N half 0 -> columns 0..255 ----
            for (int s = 0; s < 2 * KB1; ++s) {
                const int half = (s < KB1) ? 1 : 0, kb = s & (KB1 - 1);
                if (s == 1 && lane == 0) stamp(p, it, 8);   // first operand stage of this tile has landed
                if (s == KB1) {  // columns 0..255 still hold the previous tile's output accumulator until its epilogue read it
                    mbar_wait(&bars[B_ACCE], (it & 1) ^ 1, 4);
                    tc_fence_after();
                }
                mbar_wait(&bars[B_FULLA + sa], pha, 5);
                tc_fence_after();
                const uint32_t st = sbase + sa * STAGE_A;
                const uint32_t a_hi = st, a_lo = st + A_BYTES, b_hi = st + 2 * A_BYTES, b_lo = b_hi + W1_BYTES;
                const uint32_t d_tmem = tmem_base + (uint32_t)(half * 256);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t koff = k * UMMA_K * 2;
                    umma(d_tmem, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc1, (kb | k) != 0);
                    if (SPLIT == 3) {
                        umma(d_tmem, make_desc(a_lo + koff), make_desc(b_hi + koff), idesc1, 1);
                        umma(d_tmem, make_desc(a_hi + koff), make_desc(b_lo + koff), idesc1, 1);
                    }
                }
                umma_commit(&bars[B_EMPTYA + sa]);
                if (s == KB1 - 1) umma_commit(&bars[B_H1FULL]);  // columns 256..511 complete: their LN partial sums can start
                if (++sa == NSTAGE_A) { sa = 0; pha ^= 1; }
            }
            umma_commit(&bars[B_HFULL]);
            if (lane == 0) stamp(p, it, 9);                 // all first-GEMM MMAs issued
            // ---- phase B: out = GELU(LN(h)) . W3^T from the ring; accumulator in columns 0..255 (two N halves of 128) ----
            for (int j = 0; j < KB2; ++j) {
                const int slot = j & (NSLOT - 1);
                if (j == 0) {  // the accumulator overwrites h columns 0..255: k-blocks 0..3 must have been converted
                    for (int s = NSLOT - 1; s >= 0; --s) mbar_wait(&bars[B_SLOTF + s], 0, 6);
                } else if (j >= NSLOT) {
                    mbar_wait(&bars[B_SLOTF + slot], 1, 7);
                }
                tc_fence_after();
                const uint32_t ring = sbase + RING_OFF + slot * SLOT;
                // N = 256 per instruction: with two N = 128 halves the A operand was read from shared memory twice and the
                // tensor pipe became shared-memory-bandwidth bound (8 KB per 64-cycle MMA = 128 B/clk); the W3 tiles are
                // therefore streamed as half k-blocks {256 rows x 32 K columns} (SWIZZLE_64B), two UMMA_K steps each
                for (int kh = 0; kh < 2; ++kh) {
                    mbar_wait(&bars[B_W3F + s3], ph3, 8);
                    tc_fence_after();
                    const uint32_t w = sbase + W3_OFF + s3 * STAGE_3;
#pragma unroll
                    for (int k = 0; k < BK3 / UMMA_K; ++k) {
                        const uint32_t aoff = (uint32_t)((kh * 2 + k) * UMMA_K * 2), boff = (uint32_t)(k * UMMA_K * 2);
                        umma(tmem_base, make_desc(ring + aoff), make_desc_sw64(w + boff), idesc1, (j | kh | k) != 0);
                        if (SPLIT == 3) {
                            umma(tmem_base, make_desc(ring + A_BYTES + aoff), make_desc_sw64(w + boff), idesc1, 1);
                            umma(tmem_base, make_desc(ring + aoff), make_desc_sw64(w + W3_BYTES + boff), idesc1, 1);
                        }
                    }
                    umma_commit(&bars[B_W3E + s3]);
                    if (++s3 == NSTAGE_3) { s3 = 0; ph3 ^= 1; }
                }
                umma_commit(&bars[B_SLOTE + slot]);
                if (lane == 0) stamp(p, it, 24 + j);
            }
            umma_commit(&bars[B_ACCF]);
            if (lane == 0) stamp(p, it, 10);                // all second-GEMM MMAs issued
        }
    } else {
        // ===================== epilogue warps =====================
        const int e = warp - 2;
        const int q = warp & 3;          // TMEM lane quarter this warp may access
        const int part = e >> 2;         // column part 0..3
        const int r = q * 32 + lane;     // row of the tile == TMEM lane
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t red = sbase + RED_OFF;              // float red[2][4][128]
        const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const long long grow = (long long)tile * BM + r;
            const bool row_ok = grow < p.T;
            if (e == 0 && lane == 0) stamp(p, it, 0);
            // ---- LayerNorm statistics in ONE pass over TMEM with shifted data: pivot c = one element of the row,
            //      S1 = sum(x - c), S2 = sum((x - c)^2); mean = c + S1/n, var = S2/n - (S1/n)^2.  The pivot is a sample of
            //      the row, so |mean - c| ~ std and the subtraction loses no more than a couple of bits.  Each thread sums
            //      64 columns of N half 1 as soon as that half has retired (while the MMA warp computes half 0), then 64
            //      columns of half 0.
            float mean, rstd;
            {
                float s1 = 0.f, s2 = 0.f, piv = 0.f;
#pragma unroll 1
                for (int hh = 1; hh >= 0; --hh) {
                    if (hh == 1) {
                        mbar_wait(&bars[B_H1FULL], it & 1, 12);
                        tc_fence_after();
                        uint32_t v0[1];
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v0[0]) : "r"(tq + 256));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        piv = __uint_as_float(v0[0]) + tb.b1[256];
                    } else {
                        mbar_wait(&bars[B_HFULL], it & 1, 9);
                        tc_fence_after();
                        if (e == 0 && lane == 0) stamp(p, it, 1);
                    }
#pragma unroll 1
                    for (int cc = 0; cc < 2; ++cc) {
                        const int c0 = hh * 256 + part * 64 + cc * 32;
                        uint32_t v[32];
                        tmem_ld32(tq + c0, v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 b = *reinterpret_cast<const float4*>(&tb.b1[c0 + 4 * i]);
                            const float d0 = __uint_as_float(v[4 * i]) + (b.x - piv), d1 = __uint_as_float(v[4 * i + 1]) + (b.y - piv);
                            const float d2 = __uint_as_float(v[4 * i + 2]) + (b.z - piv), d3 = __uint_as_float(v[4 * i + 3]) + (b.w - piv);
                            s1 += (d0 + d1) + (d2 + d3);
                            s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2); s2 = fmaf(d2, d2, s2); s2 = fmaf(d3, d3, s2);
                        }
                    }
                }
                // The scratch lives in ring slot 3, which the previous tile's conversion pass wrote last (k-block 7).  Those writes
                // are already ordered before this point through SLOTF -> MMA -> tcgen05.commit(ACCF) -> every warp's ACCF wait
                // in phase C, but that chain runs through the async proxy, which compute-sanitizer's racecheck cannot follow
                // (it reported the pair); one named barrier per tile (~100 cycles of ~60k) makes the order explicit.
                asm volatile("bar.sync 5, 512;" ::: "memory");
                sts_f32(red + 4 * (part * 128 + r), s1);
                sts_f32(red + 4 * (512 + part * 128 + r), s2);
                asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
                const float t1 = (lds_f32(red + 4 * r) + lds_f32(red + 4 * (128 + r)) + lds_f32(red + 4 * (256 + r)) + lds_f32(red + 4 * (384 + r))) *
                                 (1.f / D_HID);
                const float t2 = (lds_f32(red + 4 * (512 + r)) + lds_f32(red + 4 * (640 + r)) + lds_f32(red + 4 * (768 + r)) +
                                  lds_f32(red + 4 * (896 + r))) * (1.f / D_HID);
                mean = piv + t1;
                rstd = rsqrtf(fmaxf(t2 - t1 * t1, 0.f) + 1e-5f);
            }
            if (e == 0 && lane == 0) stamp(p, it, 2);
            // ---- per k-block: LN + GELU + split -> ring slot (SWIZZLE_128B K-major rows), 16 columns per thread ----
#pragma unroll 1
            for (int j = 0; j < KB2; ++j) {
                const int slot = j & (NSLOT - 1);
                const int c0 = j * BK + part * 16;
                uint32_t v[16];
                tmem_ld16(tq + c0, v);
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b = *reinterpret_cast<const float4*>(&tb.b1[c0 + 4 * i]);
                    const float4 g = *reinterpret_cast<const float4*>(&tb.g[c0 + 4 * i]);
                    const float4 be = *reinterpret_cast<const float4*>(&tb.be[c0 + 4 * i]);
#if PRAM_MLP_PACKED
                    float y[4];
                    {
                        const f32x2 nm = pk2(-mean, -mean), rs = pk2(rstd, rstd);
                        const f32x2 ya = fma2(add2(pk2u(v[4 * i], v[4 * i + 1]), add2(pk2(b.x, b.y), nm)), mul2(rs, pk2(g.x, g.y)), pk2(be.x, be.y));
                        const f32x2 yb = fma2(add2(pk2u(v[4 * i + 2], v[4 * i + 3]), add2(pk2(b.z, b.w), nm)), mul2(rs, pk2(g.z, g.w)), pk2(be.z, be.w));
                        gelu_fast2(ya, y[0], y[1]);
                        gelu_fast2(yb, y[2], y[3]);
                    }
#else
                    float y[4] = {fmaf(__uint_as_float(v[4 * i]) + (b.x - mean), rstd * g.x, be.x),
                                  fmaf(__uint_as_float(v[4 * i + 1]) + (b.y - mean), rstd * g.y, be.y),
                                  fmaf(__uint_as_float(v[4 * i + 2]) + (b.z - mean), rstd * g.z, be.z),
                                  fmaf(__uint_as_float(v[4 * i + 3]) + (b.w - mean), rstd * g.w, be.w)};
#pragma unroll
                    for (int k = 0; k < 4; ++k) y[k] = gelu_fast(y[k]);
#endif
                    split2(y[0], y[1], hi[2 * i], lo[2 * i]);
                    split2(y[2], y[3], hi[2 * i + 1], lo[2 * i + 1]);
                }
                if (j >= NSLOT) mbar_wait(&bars[B_SLOTE + slot], 0, 10);  // k-block j-4 has been consumed by the MMA
                if (j == NSLOT - 1) {
                    // slot 3 doubles as the LN scratch: every warp of this quarter has read its statistics (program order +
                    // the second bar.sync above); the other quarters use disjoint rows of the scratch but the SAME slot
                    // bytes -> all 16 warps must be past the statistics before anyone writes operand data there
                    asm volatile("bar.sync 5, 512;" ::: "memory");
                }
                const uint32_t sl = sbase + RING_OFF + slot * SLOT + row_off;
                const int ch = part * 2;
                sts_u4(sl + (uint32_t)(((ch) ^ (r & 7)) << 4), make_uint4(hi[0], hi[1], hi[2], hi[3]));
                sts_u4(sl + (uint32_t)(((ch + 1) ^ (r & 7)) << 4), make_uint4(hi[4], hi[5], hi[6], hi[7]));
                if (SPLIT == 3) {
                    sts_u4(sl + A_BYTES + (uint32_t)(((ch) ^ (r & 7)) << 4), make_uint4(lo[0], lo[1], lo[2], lo[3]));
                    sts_u4(sl + A_BYTES + (uint32_t)(((ch + 1) ^ (r & 7)) << 4), make_uint4(lo[4], lo[5], lo[6], lo[7]));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
                tc_fence_before();                                            // and this warp's TMEM reads are complete
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[B_SLOTF + slot]);
                if (e == 0 && lane == 0) stamp(p, it, 16 + j);
            }
            // ---- phase C: out = acc + b3 + residual -> split bf16 (+ fp32); 64 columns per thread in two chunks ----
            // residual: fp32 rows when given, else x = hi + lo of this tile's own A rows (just streamed by phase A, L2-hot):
            // the activation stream then lives in HBM as its two bf16 planes only -- 1 KB per token written per block
            // instead of 3 KB moved (fp32 read + fp32 write + planes), which was the DRAM burst that stalled all SMs at once
            // Global accesses of phase C go through a 2 KB staging tile per warp ([32 rows][32 columns] of one bf16 plane,
            // 16-byte pieces XOR-swizzled by the row pair): the thread-per-row accesses of the first version touched 32
            // different 128-byte lines per instruction (1024 L1 wavefronts per warp and tile, which is what bounded the phase:
            // 21k cycles, exposed on the last tile of a CTA and slowing the next tile's first GEMM otherwise); four lanes per row
            // move whole 64-byte segments, 8 lines per instruction.  The staging bytes are the last W3 stage, idle between the
            // second GEMM of this tile and the W3 stream of the next one (B_CDONE orders the two).
            const uint32_t stg = sbase + W3_OFF + (NSTAGE_3 - 1) * STAGE_3 + (uint32_t)e * 2048u;
            auto swz = [](int row, int piece) { return (uint32_t)(row * 64 + ((piece ^ ((row >> 1) & 3)) << 4)); };
            const int crow = lane >> 2, cpiece = lane & 3;                      // coalesced mapping: row = 8 t + crow
            const long long tile_row0 = (long long)tile * BM + q * 32;
            // one plane tile (columns n0 .. n0 + 31 of this warp's 32 rows) -> this lane's row as 16 words
            auto load_plane = [&](const __nv_bfloat16* plane, int n0, uint32_t (&w)[16]) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int row = t * 8 + crow;
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (tile_row0 + row < p.T) v = __ldg(reinterpret_cast<const uint4*>(plane + (tile_row0 + row) * p.lda + n0 + cpiece * 8));
                    sts_u4(stg + swz(row, cpiece), v);
                }
                __syncwarp();
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) {
                    const uint4 v = lds_u4(stg + swz(lane, pc));
                    w[4 * pc] = v.x; w[4 * pc + 1] = v.y; w[4 * pc + 2] = v.z; w[4 * pc + 3] = v.w;
                }
                __syncwarp();
            };
            // this lane's row of one output plane (16 words = 32 bf16) -> global, through the staging tile
            auto store_plane = [&](__nv_bfloat16* plane, int n0, const uint4 (&w)[4]) {
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) sts_u4(stg + swz(lane, pc), w[pc]);
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int row = t * 8 + crow;
                    const uint4 v = lds_u4(stg + swz(row, cpiece));
                    if (tile_row0 + row < p.T) *reinterpret_cast<uint4*>(plane + (tile_row0 + row) * p.ld_bf + n0 + cpiece * 8) = v;
                }
                __syncwarp();
            };
            auto load_res = [&](int n0, float4 (&rv)[8]) {
                if (p.res) {   // fp32 residual rows (AdaGML keeps the fp32 stream): thread-per-row loads as before
                    if (!row_ok) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) rv[i] = __ldg(reinterpret_cast<const float4*>(p.res + grow * p.res_ld + n0) + i);
                    }
                    return;
                }
                uint32_t h[16], l[16];
                load_plane(p.xa_hi, n0, h);
                if (SPLIT == 3) load_plane(p.xa_lo, n0, l);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t h0 = h[2 * i], h1 = h[2 * i + 1];
                    const uint32_t l0 = (SPLIT == 3) ? l[2 * i] : 0u, l1 = (SPLIT == 3) ? l[2 * i + 1] : 0u;
                    rv[i] = make_float4(__uint_as_float(h0 << 16) + __uint_as_float(l0 << 16),
                                        __uint_as_float(h0 & 0xffff0000u) + __uint_as_float(l0 & 0xffff0000u),
                                        __uint_as_float(h1 << 16) + __uint_as_float(l1 << 16),
                                        __uint_as_float(h1 & 0xffff0000u) + __uint_as_float(l1 & 0xffff0000u));
                }
            };
            // first chunk's residual BEFORE the accumulator wait (the warps idle through the tail of the second GEMM anyway) --
            // thread-per-row loads, because the staging bytes are a W3 stage the MMA may still be reading until B_ACCF
            auto load_res_direct = [&](int n0, float4 (&rv)[8]) {
                if (p.res) { load_res(n0, rv); return; }
                if (!row_ok) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    return;
                }
                const uint4* ph = reinterpret_cast<const uint4*>(p.xa_hi + grow * p.lda + n0);
                const uint4* pl = reinterpret_cast<const uint4*>(p.xa_lo + grow * p.lda + n0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 h = __ldg(ph + i);
                    uint4 l = make_uint4(0u, 0u, 0u, 0u);
                    if (SPLIT == 3) l = __ldg(pl + i);
                    rv[2 * i] = make_float4(__uint_as_float(h.x << 16) + __uint_as_float(l.x << 16),
                                            __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u),
                                            __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16),
                                            __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u));
                    rv[2 * i + 1] = make_float4(__uint_as_float(h.z << 16) + __uint_as_float(l.z << 16),
                                                __uint_as_float(h.z & 0xffff0000u) + __uint_as_float(l.z & 0xffff0000u),
                                                __uint_as_float(h.w << 16) + __uint_as_float(l.w << 16),
                                                __uint_as_float(h.w & 0xffff0000u) + __uint_as_float(l.w & 0xffff0000u));
                }
            };
            if (e == 0 && lane == 0) stamp(p, it, 3);
            float4 rv[8];
            load_res_direct(part * 64, rv);
            mbar_wait(&bars[B_ACCF], it & 1, 11);
            tc_fence_after();
            if (e == 0 && lane == 0) stamp(p, it, 4);
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int n0 = part * 64 + cc * 32;
                uint32_t v[32];
                tmem_ld32(tq + n0, v);
                if (cc == 1) {
                    // the accumulator is in registers: hand TMEM columns 0..255 back to the MMA warp before the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[B_ACCE]);
                    load_res(n0, rv);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float f0 = __uint_as_float(v[4 * i]) + tb.b3[n0 + 4 * i] + rv[i].x;
                    const float f1 = __uint_as_float(v[4 * i + 1]) + tb.b3[n0 + 4 * i + 1] + rv[i].y;
                    const float f2 = __uint_as_float(v[4 * i + 2]) + tb.b3[n0 + 4 * i + 2] + rv[i].z;
                    const float f3 = __uint_as_float(v[4 * i + 3]) + tb.b3[n0 + 4 * i + 3] + rv[i].w;
                    if (p.out_f32 && row_ok) *reinterpret_cast<float4*>(p.out_f32 + grow * p.ld_f32 + n0 + 4 * i) = make_float4(f0, f1, f2, f3);
                    split2(f0, f1, v[4 * i], v[4 * i + 2]);       // re-use v: [4i] = hi01, [4i+1] = hi23, [4i+2] = lo01, [4i+3] = lo23
                    uint32_t h23, l23;
                    split2(f2, f3, h23, l23);
                    v[4 * i + 1] = h23; v[4 * i + 3] = l23;
                }
                if (p.out_hi) {   // warp-uniform
                    uint4 w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[i] = make_uint4(v[8 * i], v[8 * i + 1], v[8 * i + 4], v[8 * i + 5]);
                    store_plane(p.out_hi, n0, w);
                    if (SPLIT == 3 && p.out_lo) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) w[i] = make_uint4(v[8 * i + 2], v[8 * i + 3], v[8 * i + 6], v[8 * i + 7]);
                        store_plane(p.out_lo, n0, w);
                    }
                }
            }
            // phase C no longer touches the staging tile: the producer may stream the next tile's W3 into that stage
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_CDONE]);
            if (e == 0 && lane == 0) stamp(p, it, 5);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
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

// 2-D bf16 row-major matrix [rows][ld] -> boxes of {64 columns, box_rows rows}, SWIZZLE_128B
static int encode2d(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows,
                    int box_cols = BK) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return PRAM_ERR_CUDA;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, str, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PRAM_OK : PRAM_ERR_CUDA;
}

}  // namespace mb

// Public argument block of pram_mlp_block_tc (mirrored by ctypes in pram_b200/_lib.py).
struct pram_mlp_block_args {
    const void* a_hi; const void* a_lo;   // bf16 [T][lda]: columns 0..255 = x, 256..511 = attention context (pre-projection)
    long long lda;
    int T;
    const void* w1_hi; const void* w1_lo; // bf16 [8][512][64]: mlp.0 with proj folded into its right half, tiled k-block major
                                          // (tile kb = W1[:, 64 kb : 64 kb + 64])
    const void* w3_hi; const void* w3_lo; // bf16 [16][256][32]: mlp.3, tiled half-k-block major (tile s = W3[:, 32 s : 32 s + 32])
    const float* tables_host;             // HOST pointer, fp32 [b1 512 | ln gamma 512 | ln beta 512 | b3 256]: copied into the
                                          // kernel's parameter block (constant bank) at launch
    const float* res; long long res_ld;   // fp32 residual rows, or NULL: the residual is x = hi + lo of columns 0..255 of `a`
    float* out_f32; long long ld_f32;     // may be NULL
    void* out_hi; void* out_lo; long long ld_bf;  // may be NULL
    int split;                            // 1: bf16, 3: bf16x3
    long long* dbg;                       // optional device buffer [grid][8][32] of SM clock stamps (profiling aid), NULL = off
};

PRAM_API int pram_mlp_block_tc(const pram_mlp_block_args* a, cudaStream_t stream) {
    using namespace mb;
    if (!a || !a->a_hi || !a->w1_hi || !a->w3_hi || !a->tables_host || a->T <= 0) return PRAM_ERR_ARG;
    if (a->split != 1 && a->split != 3) return PRAM_ERR_ARG;
    if (a->split == 3 && (!a->a_lo || !a->w1_lo || !a->w3_lo)) return PRAM_ERR_ARG;
    if (!a->out_f32 && !a->out_hi) return PRAM_ERR_ARG;
    if ((a->lda % 8) || a->lda < D_IN) return PRAM_ERR_UNSUPPORTED;                      // 16-byte TMA strides
    if (a->res && (a->res_ld % 4)) return PRAM_ERR_UNSUPPORTED;                        // float4 residual rows
    if (a->out_f32 && (a->ld_f32 % 4)) return PRAM_ERR_UNSUPPORTED;
    if (a->out_hi && (a->ld_bf % 8)) return PRAM_ERR_UNSUPPORTED;
    static int sms = 0;
    if (!sms) { int dev = 0; PRAM_CUDA(cudaGetDevice(&dev)); PRAM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)); }
    CUtensorMap ah, al, w1h, w1l, w3h, w3l;
    int rc;
    if ((rc = encode2d(&ah, a->a_hi, a->T, D_IN, a->lda, BM))) return rc;
    if ((rc = encode2d(&al, a->a_lo ? a->a_lo : a->a_hi, a->T, D_IN, a->lda, BM))) return rc;
    // weights arrive pre-tiled (k-block major): every TMA box is one contiguous 32 KB / 16 KB run of memory
    if ((rc = encode2d(&w1h, a->w1_hi, (long long)KB1 * D_HID, BK, BK, 256))) return rc;
    if ((rc = encode2d(&w1l, a->w1_lo ? a->w1_lo : a->w1_hi, (long long)KB1 * D_HID, BK, BK, 256))) return rc;
    if ((rc = encode2d(&w3h, a->w3_hi, (long long)(D_HID / BK3) * D_OUT, BK3, BK3, 256, BK3))) return rc;
    if ((rc = encode2d(&w3l, a->w3_lo ? a->w3_lo : a->w3_hi, (long long)(D_HID / BK3) * D_OUT, BK3, BK3, 256, BK3))) return rc;
    Args k;
    k.T = a->T;
    k.res = a->res; k.res_ld = a->res_ld; k.out_f32 = a->out_f32; k.ld_f32 = a->ld_f32;
    k.xa_hi = (const __nv_bfloat16*)a->a_hi; k.xa_lo = (const __nv_bfloat16*)(a->a_lo ? a->a_lo : a->a_hi); k.lda = a->lda;
    k.dbg = a->dbg;
    k.pred = g_pram_pred;
    k.pdl_early = g_pram_pdl >= 2;
    Tables tb;
    memcpy(&tb, a->tables_host, sizeof(Tables));
    k.out_hi = (__nv_bfloat16*)a->out_hi; k.out_lo = (__nv_bfloat16*)a->out_lo; k.ld_bf = a->ld_bf;
    const int ntiles = (a->T + BM - 1) / BM;
    const int grid = ntiles < sms ? ntiles : sms;
    if (a->split == 3) {
        auto kern = mlp_block_kernel<3>;
        static bool attr = false;
        if (!attr) { PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr = true; }
        PRAM_CUDA(pram_launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, ah, al, w1h, w1l, w3h, w3l, k, tb));
    } else {
        auto kern = mlp_block_kernel<1>;
        static bool attr = false;
        if (!attr) { PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr = true; }
        PRAM_CUDA(pram_launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, ah, al, w1h, w1l, w3h, w3l, k, tb));
    }
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}
