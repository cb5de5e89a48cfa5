// conv1a of SFD2 (3 -> 64 channels, 3x3, BN folded, ReLU; reference nets/sfd2.py:141) on tcgen05 tensor cores.
//
// As a GEMM it is M = pixels, N = 64, K = 27: far too thin for TMA-fed implicit GEMM (3 input channels = 6 bytes per
// pixel), and as an FFMA kernel it is bound by the FP32 pipe (1728 FMAs per pixel; the CUDA-core version ran at
// 1.5 ms per 32 frames, 27 % of the HBM roofline it should sit on).  Here each CTA builds the im2col operand
// itself: thread t gathers the 27 taps of pixel t from the NCHW fp32 image (coalesced along x), splits them into bf16
// hi / lo and writes its row of a 128 x 32 K-major SWIZZLE_128B shared-memory tile -- exactly the layout TMA would
// produce -- then one elected lane issues the 2 k-steps x 3 (bf16x3) tcgen05.mma into a 64-column TMEM
// accumulator.  The epilogue (tcgen05.ld -> bias -> ReLU -> hi/lo split) is staged through shared memory so that 8
// lanes write one pixel's 128-byte line of the 2x2 PHASE-SPLIT output tensor that feeds conv1b's stride-2 TMA boxes.
// Phases of one tile are serialised inside a CTA; four CTAs per SM overlap them.  HBM-bound by design: 12 B read and
// 256 B written per pixel.
#include "common.cuh"
#include <stdio.h>

namespace c1 {

constexpr int TM = 128;            // pixels per tile (one row segment) == UMMA_M == threads per CTA
constexpr int NOUT = 64;           // output channels == UMMA_N
constexpr int KPAD = 32;           // 27 taps padded to two UMMA_K = 16 steps
constexpr int ROW_BYTES = 128;     // one swizzle row: 64 bf16 (only the first 32 are used)
constexpr int A_BYTES = TM * ROW_BYTES;      // 16 KB per plane
constexpr int B_BYTES = NOUT * ROW_BYTES;    // 8 KB per plane
constexpr int SMEM_BYTES = 2 * A_BYTES + 2 * B_BYTES + 512 + 1024 /*alignment*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) {
            printf("pram conv1a_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
// K-major SWIZZLE_128B descriptor, identical to gemm_tc.cu: start >> 4, SBO = 1024 B (8 rows x 128 B), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
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
// explicit shared-space 16-byte accessors (the tile pointers come from the manually aligned dynamic shared-memory base)
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<uint32_t*>(&h);
    __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<uint32_t*>(&l);
}

template <int SPLIT>
__global__ void __launch_bounds__(TM, 4) conv1a_tc_kernel(const float* __restrict__ img, const float* __restrict__ w /*[27][64]*/,
                                                          const float* __restrict__ bias, int B, int H, int W,
                                                          __nv_bfloat16* __restrict__ ps_hi, __nv_bfloat16* __restrict__ ps_lo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_hi = smem;
    uint8_t* a_lo = smem + A_BYTES;
    uint8_t* b_hi = smem + 2 * A_BYTES;
    uint8_t* b_lo = b_hi + B_BYTES;
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(b_lo + B_BYTES);
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(mma_done + 1);
    float* bias_s = reinterpret_cast<float*>(mma_done + 2);  // [64]
    const int t = threadIdx.x, warp = t >> 5;

    if (t == 0) {
        mbar_init(mma_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(NOUT));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (t < NOUT) bias_s[t] = bias[t];
    // weight operand B[n][k] (K-major rows of 128 B, 16-byte chunk j of row n at chunk position j ^ (n & 7))
    for (int idx = t; idx < NOUT * (KPAD / 8); idx += TM) {
        const int n = idx >> 2, j = idx & 3;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const int k0 = 8 * j + i;
            const float w0 = (k0 < 27) ? __ldg(w + k0 * NOUT + n) : 0.f;
            const float w1 = (k0 + 1 < 27) ? __ldg(w + (k0 + 1) * NOUT + n) : 0.f;
            split2(w0, w1, hi[i >> 1], lo[i >> 1]);
        }
        const int off = n * ROW_BYTES + ((j ^ (n & 7)) << 4);
        sts_u4(smem_u32(b_hi) + off, make_uint4(hi[0], hi[1], hi[2], hi[3]));
        sts_u4(smem_u32(b_lo) + off, make_uint4(lo[0], lo[1], lo[2], lo[3]));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    constexpr uint32_t idesc = make_idesc(TM, NOUT);
    const uint32_t sa_hi_g = smem_u32(a_hi), sa_lo_g = smem_u32(a_lo), bias_a = smem_u32(bias_s);

    const int tiles_x = (W + TM - 1) / TM;
    const long long total = (long long)B * H * tiles_x;
    const int Hp = (H + 1) >> 1, Wp = (W + 1) >> 1;
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int tx = (int)(tile % tiles_x);
        const int y = (int)((tile / tiles_x) % H);
        const int b = (int)(tile / ((long long)tiles_x * H));
        const int x = tx * TM + t;
        // ---- im2col row of this thread's pixel: k = (ry * 3 + rx) * 3 + c ----
        {
            float v[KPAD];
#pragma unroll
            for (int k = 27; k < KPAD; ++k) v[k] = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int ry = 0; ry < 3; ++ry)
#pragma unroll
                    for (int rx = 0; rx < 3; ++rx) {
                        const int iy = y + ry - 1, ix = x + rx - 1;
                        v[(ry * 3 + rx) * 3 + c] = (iy >= 0 && iy < H && ix >= 0 && ix < W)
                                                       ? __ldg(img + (((long long)b * 3 + c) * H + iy) * W + ix) : 0.f;
                    }
#pragma unroll
            for (int j = 0; j < KPAD / 8; ++j) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split2(v[8 * j + 2 * i], v[8 * j + 2 * i + 1], hi[i], lo[i]);
                const int off = t * ROW_BYTES + ((j ^ (t & 7)) << 4);
                sts_u4(sa_hi_g + off, make_uint4(hi[0], hi[1], hi[2], hi[3]));
                if (SPLIT == 3) sts_u4(sa_lo_g + off, make_uint4(lo[0], lo[1], lo[2], lo[3]));
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo), sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
#pragma unroll
            for (int k = 0; k < KPAD / 16; ++k) {
                const uint32_t ko = k * 32;
                umma(tmem_base, make_desc(sa_hi + ko), make_desc(sb_hi + ko), idesc, k != 0);
                if (SPLIT == 3) {
                    umma(tmem_base, make_desc(sa_lo + ko), make_desc(sb_hi + ko), idesc, 1);
                    umma(tmem_base, make_desc(sa_hi + ko), make_desc(sb_lo + ko), idesc, 1);
                }
            }
            umma_commit(mma_done);
        }
        mbar_wait(mma_done, it & 1);
        tc_fence_after();
        // ---- epilogue: thread t owns TMEM lane t = pixel t; the A tiles are free again -> output staging ----
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + hc * 32, acc);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int co = hc * 32 + 8 * j + 2 * i;
                    const float f0 = fmaxf(__uint_as_float(acc[8 * j + 2 * i]) + lds_f32(bias_a + 4 * co), 0.f);
                    const float f1 = fmaxf(__uint_as_float(acc[8 * j + 2 * i + 1]) + lds_f32(bias_a + 4 * co + 4), 0.f);
                    split2(f0, f1, hi[i], lo[i]);
                }
                const int ch = hc * 4 + j;  // 16-byte chunk (8 channels) of this pixel's 128-byte line
                const int off = t * ROW_BYTES + ((ch ^ (t & 7)) << 4);
                sts_u4(sa_hi_g + off, make_uint4(hi[0], hi[1], hi[2], hi[3]));
                if (SPLIT == 3) sts_u4(sa_lo_g + off, make_uint4(lo[0], lo[1], lo[2], lo[3]));
            }
        }
        tc_fence_before();  // accumulator reads ordered before the next tile's MMA
        __syncthreads();
        const int x_base = tx * TM;
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {
            const int px = i8 * 16 + (t >> 3), ch = t & 7;
            const int xx = x_base + px;
            if (xx < W) {
                const long long off = ((((long long)(b * 4 + (y & 1) * 2 + (xx & 1))) * Hp + (y >> 1)) * Wp + (xx >> 1)) * 64 + ch * 8;
                const int so = px * ROW_BYTES + ((ch ^ (px & 7)) << 4);
                *reinterpret_cast<uint4*>(ps_hi + off) = lds_u4(sa_hi_g + so);
                if (SPLIT == 3) *reinterpret_cast<uint4*>(ps_lo + off) = lds_u4(sa_lo_g + so);
            }
        }
        __syncthreads();  // staging reads done before the next tile's im2col rows overwrite the region
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NOUT));
    }
}

}  // namespace c1

// img fp32 NCHW [B][3][H][W]; w fp32 [27][64] (k = (ry*3 + rx)*3 + c); bias [64];
// ps_hi / ps_lo: bf16 [B*4][ceil(H/2)][ceil(W/2)][64] phase-split planes (plane = (y&1)*2 + (x&1)); ps_lo NULL when split == 1.
PRAM_API int pram_conv1a_tc(const float* img, const float* w, const float* bias, int B, int H, int W, void* ps_hi, void* ps_lo,
                            int split, cudaStream_t stream) {
    using namespace c1;
    if (!img || !w || !bias || !ps_hi || B <= 0 || H <= 0 || W <= 0) return PRAM_ERR_ARG;
    if (split != 1 && split != 3) return PRAM_ERR_ARG;
    if (split == 3 && !ps_lo) return PRAM_ERR_ARG;
    static int sms = 0;
    if (!sms) { int dev = 0; PRAM_CUDA(cudaGetDevice(&dev)); PRAM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)); }
    const long long total = (long long)B * H * cdiv(W, TM);
    const int grid = (int)(total < 4LL * sms ? total : 4LL * sms);
    if (split == 3) {
        auto kern = conv1a_tc_kernel<3>;
        static bool attr = false;
        if (!attr) { PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr = true; }
        kern<<<grid, TM, SMEM_BYTES, stream>>>(img, w, bias, B, H, W, (__nv_bfloat16*)ps_hi, (__nv_bfloat16*)ps_lo);
    } else {
        auto kern = conv1a_tc_kernel<1>;
        static bool attr = false;
        if (!attr) { PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr = true; }
        kern<<<grid, TM, SMEM_BYTES, stream>>>(img, w, bias, B, H, W, (__nv_bfloat16*)ps_hi, nullptr);
    }
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}
