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
//             out[128 x 256] += y_kblock . W3_kblock^T   (accumulator re-uses TMEM columns 0..255 once k-blocks 0..3,
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

namespace mb {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int D_IN = 512, D_HID = 512, D_OUT = 256;
constexpr int KB1 = D_IN / BK, KB2 = D_HID / BK;            // 8, 8
constexpr int EPI_WARPS = 16;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;             // 576
constexpr int A_BYTES = BM * BK * 2;                         // 16 KB: 128 rows x 128 B
constexpr int W1_BYTES = 256 * BK * 2;                       // 32 KB: one N half of W1, one k-block
constexpr int W3_BYTES = 128 * BK * 2;                       // 16 KB: one N half of W3, one k-block
constexpr int STAGE_A = 2 * A_BYTES + 2 * W1_BYTES;          // 96 KB (both planes; SPLIT = 1 uses the hi halves only)
constexpr int NSTAGE_A = 2;
constexpr int SLOT = 2 * A_BYTES;                            // 32 KB ring slot: hi | lo
constexpr int NSLOT = 4;
constexpr int STAGE_3 = 2 * W3_BYTES;                        // 32 KB
constexpr int NSTAGE_3 = 3;
constexpr int RING_OFF = 0, W3_OFF = NSLOT * SLOT;           // 0, 128 KB
constexpr int REGION = W3_OFF + NSTAGE_3 * STAGE_3;          // 224 KB  (>= NSTAGE_A * STAGE_A = 192 KB)
constexpr int RED_OFF = RING_OFF + 3 * SLOT;                 // LN partial sums live in ring slot 3 (written last)
constexpr int BAR_OFF = REGION;                              // barriers (256 B) then the b3 table (1 KB)
constexpr int SMEM_BYTES = 1024 + REGION + 256 + D_OUT * 4;
static_assert(NSTAGE_A * STAGE_A <= REGION, "phase A stages must fit the multiplexed region");
static_assert(SMEM_BYTES <= 227 * 1024, "dynamic shared memory budget of sm_100a exceeded");

struct Args {
    int T;
    const float* b1; const float* ln_g; const float* ln_b; const float* b3;
    const float* res; long long res_ld;
    float* out_f32; long long ld_f32;
    __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; long long ld_bf;
};

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
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<uint32_t*>(&h);
    __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<uint32_t*>(&l);
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
             B_COUNT = 21 };

template <int SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_block_kernel(
    const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
    const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
    const __grid_constant__ CUtensorMap map_w3_hi, const __grid_constant__ CUtensorMap map_w3_lo, const Args p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + B_COUNT);
    float* b3_s = reinterpret_cast<float*>(smem + BAR_OFF + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (p.T + BM - 1) / BM;
    constexpr uint32_t A_TX = (SPLIT == 3 ? 2u : 1u) * (A_BYTES + W1_BYTES);
    constexpr uint32_t W3_TX = (SPLIT == 3 ? 2u : 1u) * W3_BYTES;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a_hi); prefetch_tmap(&map_w1_hi); prefetch_tmap(&map_w3_hi);
        if (SPLIT == 3) { prefetch_tmap(&map_a_lo); prefetch_tmap(&map_w1_lo); prefetch_tmap(&map_w3_lo); }
        for (int s = 0; s < NSTAGE_A; ++s) { mbar_init(&bars[B_FULLA + s], 1); mbar_init(&bars[B_EMPTYA + s], 1); }
        mbar_init(&bars[B_HFULL], 1);
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&bars[B_SLOTF + s], EPI_WARPS); mbar_init(&bars[B_SLOTE + s], 1); }
        for (int s = 0; s < NSTAGE_3; ++s) { mbar_init(&bars[B_W3F + s], 1); mbar_init(&bars[B_W3E + s], 1); }
        mbar_init(&bars[B_ACCF], 1);
        mbar_init(&bars[B_ACCE], EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + D_OUT) b3_s[threadIdx.x - 64] = p.b3 ? p.b3[threadIdx.x - 64] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

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
                    tma_load_2d(st + 2 * A_BYTES, &map_w1_hi, &bars[B_FULLA + sa], kb * BK, half * 256);
                    if (SPLIT == 3) {
                        tma_load_2d(st + A_BYTES, &map_a_lo, &bars[B_FULLA + sa], kb * BK, row0);
                        tma_load_2d(st + 2 * A_BYTES + W1_BYTES, &map_w1_lo, &bars[B_FULLA + sa], kb * BK, half * 256);
                    }
                    if (++sa == NSTAGE_A) { sa = 0; pha ^= 1; }
                }
                // the first GEMM has retired: its stages may be overwritten by the W3 tiles
                mbar_wait(&bars[B_HFULL], it & 1, 2);
                for (int s = 0; s < 2 * KB2; ++s) {
                    const int j = s >> 1, nh = s & 1;
                    mbar_wait(&bars[B_W3E + s3], ph3 ^ 1, 3);
                    const uint32_t st = sbase + W3_OFF + s3 * STAGE_3;
                    mbar_expect_tx(&bars[B_W3F + s3], W3_TX);
                    tma_load_2d(st, &map_w3_hi, &bars[B_W3F + s3], j * BK, nh * 128);
                    if (SPLIT == 3) tma_load_2d(st + W3_BYTES, &map_w3_lo, &bars[B_W3F + s3], j * BK, nh * 128);
                    if (++s3 == NSTAGE_3) { s3 = 0; ph3 ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, one elected lane per instruction) =====================
        constexpr uint32_t idesc1 = make_idesc(BM, 256), idesc2 = make_idesc(BM, 128);
        int sa = 0; uint32_t pha = 0;
        int s3 = 0; uint32_t ph3 = 0;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            // ---- phase A: h = [x | ctx] . W1^T, N half 1 -> columns 256..511, then N half 0 -> columns 0..255 ----
            for (int s = 0; s < 2 * KB1; ++s) {
                const int half = (s < KB1) ? 1 : 0, kb = s & (KB1 - 1);
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
                if (++sa == NSTAGE_A) { sa = 0; pha ^= 1; }
            }
            umma_commit(&bars[B_HFULL]);
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
                for (int nh = 0; nh < 2; ++nh) {
                    mbar_wait(&bars[B_W3F + s3], ph3, 8);
                    tc_fence_after();
                    const uint32_t w = sbase + W3_OFF + s3 * STAGE_3;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(nh * 128);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t koff = k * UMMA_K * 2;
                        umma(d_tmem, make_desc(ring + koff), make_desc(w + koff), idesc2, (j | k) != 0);
                        if (SPLIT == 3) {
                            umma(d_tmem, make_desc(ring + A_BYTES + koff), make_desc(w + koff), idesc2, 1);
                            umma(d_tmem, make_desc(ring + koff), make_desc(w + W3_BYTES + koff), idesc2, 1);
                        }
                    }
                    umma_commit(&bars[B_W3E + s3]);
                    if (++s3 == NSTAGE_3) { s3 = 0; ph3 ^= 1; }
                }
                umma_commit(&bars[B_SLOTE + slot]);
            }
            umma_commit(&bars[B_ACCF]);
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
        const float4* __restrict__ b1v = reinterpret_cast<const float4*>(p.b1);
        const float4* __restrict__ gv = reinterpret_cast<const float4*>(p.ln_g);
        const float4* __restrict__ bv = reinterpret_cast<const float4*>(p.ln_b);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const long long grow = (long long)tile * BM + r;
            const bool row_ok = grow < p.T;
            mbar_wait(&bars[B_HFULL], it & 1, 9);
            tc_fence_after();
            // ---- LayerNorm statistics: two exact passes (mean, then centred sum of squares), 128 columns per thread ----
            float mean, rstd;
            {
                float s = 0.f;
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int c0 = part * 128 + cc * 32;
                    uint32_t v[32];
                    tmem_ld32(tq + c0, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = __ldg(b1v + (c0 >> 2) + i);
                        s += (__uint_as_float(v[4 * i]) + b.x) + (__uint_as_float(v[4 * i + 1]) + b.y) +
                             (__uint_as_float(v[4 * i + 2]) + b.z) + (__uint_as_float(v[4 * i + 3]) + b.w);
                    }
                }
                sts_f32(red + 4 * (part * 128 + r), s);
                asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
                mean = (lds_f32(red + 4 * r) + lds_f32(red + 4 * (128 + r)) + lds_f32(red + 4 * (256 + r)) + lds_f32(red + 4 * (384 + r))) *
                       (1.f / D_HID);
                float qs = 0.f;
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int c0 = part * 128 + cc * 32;
                    uint32_t v[32];
                    tmem_ld32(tq + c0, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = __ldg(b1v + (c0 >> 2) + i);
                        const float d0 = __uint_as_float(v[4 * i]) + b.x - mean, d1 = __uint_as_float(v[4 * i + 1]) + b.y - mean;
                        const float d2 = __uint_as_float(v[4 * i + 2]) + b.z - mean, d3 = __uint_as_float(v[4 * i + 3]) + b.w - mean;
                        qs += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
                    }
                }
                sts_f32(red + 4 * (512 + part * 128 + r), qs);
                asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
                const float var = (lds_f32(red + 4 * (512 + r)) + lds_f32(red + 4 * (640 + r)) + lds_f32(red + 4 * (768 + r)) +
                                   lds_f32(red + 4 * (896 + r))) * (1.f / D_HID);
                rstd = rsqrtf(var + 1e-5f);
            }
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
                    const float4 b = __ldg(b1v + (c0 >> 2) + i), g = __ldg(gv + (c0 >> 2) + i), be = __ldg(bv + (c0 >> 2) + i);
                    float y[4] = {(__uint_as_float(v[4 * i]) + b.x - mean) * rstd * g.x + be.x,
                                  (__uint_as_float(v[4 * i + 1]) + b.y - mean) * rstd * g.y + be.y,
                                  (__uint_as_float(v[4 * i + 2]) + b.z - mean) * rstd * g.z + be.z,
                                  (__uint_as_float(v[4 * i + 3]) + b.w - mean) * rstd * g.w + be.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) y[k] = 0.5f * y[k] * (1.f + erf_as(y[k] * 0.70710678118654752440f));
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
            }
            // ---- phase C: out = acc + b3 + residual -> fp32 + split bf16; 64 columns per thread in two chunks ----
            float4 rv[8];
            {
                const int n0 = part * 64;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    rv[i] = (p.res && row_ok) ? __ldg(reinterpret_cast<const float4*>(p.res + grow * p.res_ld + n0) + i)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(&bars[B_ACCF], it & 1, 11);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int n0 = part * 64 + cc * 32;
                uint32_t v[32];
                tmem_ld32(tq + n0, v);
                if (cc == 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        rv[i] = (p.res && row_ok) ? __ldg(reinterpret_cast<const float4*>(p.res + grow * p.res_ld + n0) + i)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t ba = smem_u32(b3_s) + 4 * (n0 + 4 * i);
                        const float f0 = __uint_as_float(v[4 * i]) + lds_f32(ba) + rv[i].x;
                        const float f1 = __uint_as_float(v[4 * i + 1]) + lds_f32(ba + 4) + rv[i].y;
                        const float f2 = __uint_as_float(v[4 * i + 2]) + lds_f32(ba + 8) + rv[i].z;
                        const float f3 = __uint_as_float(v[4 * i + 3]) + lds_f32(ba + 12) + rv[i].w;
                        if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + grow * p.ld_f32 + n0 + 4 * i) = make_float4(f0, f1, f2, f3);
                        split2(f0, f1, v[4 * i], v[4 * i + 2]);       // re-use v: [4i] = hi01, [4i+1] = hi23, [4i+2] = lo01, [4i+3] = lo23
                        uint32_t h23, l23;
                        split2(f2, f3, h23, l23);
                        v[4 * i + 1] = h23; v[4 * i + 3] = l23;
                    }
                    if (p.out_hi) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            *reinterpret_cast<uint4*>(p.out_hi + grow * p.ld_bf + n0 + 8 * i) =
                                make_uint4(v[8 * i], v[8 * i + 1], v[8 * i + 4], v[8 * i + 5]);
                            if (SPLIT == 3 && p.out_lo)
                                *reinterpret_cast<uint4*>(p.out_lo + grow * p.ld_bf + n0 + 8 * i) =
                                    make_uint4(v[8 * i + 2], v[8 * i + 3], v[8 * i + 6], v[8 * i + 7]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_ACCE]);
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
static int encode2d(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return PRAM_ERR_CUDA;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, str, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PRAM_OK : PRAM_ERR_CUDA;
}

}  // namespace mb

// Public argument block of pram_mlp_block_tc (mirrored by ctypes in pram_b200/_lib.py).
struct pram_mlp_block_args {
    const void* a_hi; const void* a_lo;   // bf16 [T][lda]: columns 0..255 = x, 256..511 = attention context (pre-projection)
    long long lda;
    int T;
    const void* w1_hi; const void* w1_lo; // bf16 [512][512]: mlp.0 with proj folded into its right half
    const float* b1;                      // [512]
    const float* ln_g; const float* ln_b; // [512]
    const void* w3_hi; const void* w3_lo; // bf16 [256][512]: mlp.3
    const float* b3;                      // [256]
    const float* res; long long res_ld;   // fp32 residual rows (x), may be NULL
    float* out_f32; long long ld_f32;     // may be NULL
    void* out_hi; void* out_lo; long long ld_bf;  // may be NULL
    int split;                            // 1: bf16, 3: bf16x3
};

PRAM_API int pram_mlp_block_tc(const pram_mlp_block_args* a, cudaStream_t stream) {
    using namespace mb;
    if (!a || !a->a_hi || !a->w1_hi || !a->w3_hi || !a->b1 || !a->ln_g || !a->ln_b || a->T <= 0) return PRAM_ERR_ARG;
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
    if ((rc = encode2d(&w1h, a->w1_hi, D_HID, D_IN, D_IN, 256))) return rc;
    if ((rc = encode2d(&w1l, a->w1_lo ? a->w1_lo : a->w1_hi, D_HID, D_IN, D_IN, 256))) return rc;
    if ((rc = encode2d(&w3h, a->w3_hi, D_OUT, D_HID, D_HID, 128))) return rc;
    if ((rc = encode2d(&w3l, a->w3_lo ? a->w3_lo : a->w3_hi, D_OUT, D_HID, D_HID, 128))) return rc;
    Args k;
    k.T = a->T; k.b1 = a->b1; k.ln_g = a->ln_g; k.ln_b = a->ln_b; k.b3 = a->b3;
    k.res = a->res; k.res_ld = a->res_ld; k.out_f32 = a->out_f32; k.ld_f32 = a->ld_f32;
    k.out_hi = (__nv_bfloat16*)a->out_hi; k.out_lo = (__nv_bfloat16*)a->out_lo; k.ld_bf = a->ld_bf;
    const int ntiles = (a->T + BM - 1) / BM;
    const int grid = ntiles < sms ? ntiles : sms;
    if (a->split == 3) {
        auto kern = mlp_block_kernel<3>;
        static bool attr = false;
        if (!attr) { PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr = true; }
        kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ah, al, w1h, w1l, w3h, w3l, k);
    } else {
        auto kern = mlp_block_kernel<1>;
        static bool attr = false;
        if (!attr) { PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr = true; }
        kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ah, al, w1h, w1l, w3h, w3l, k);
    }
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}
