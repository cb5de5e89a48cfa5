// Grouped 3x3 convolution (32 groups x 8 channels, stride 1, pad 1, + bias + ReLU) of the SFD2 ResBlocks
// (reference nets/sfd2.py:100-124, `conv2 = conv3x3(width, width, stride, groups=32)`) on tensor cores.
//
// A group is an 8 -> 8 channel convolution: as a GEMM it is M = pixels, N = 8, K = 9 taps x 8 = 72 -- exactly the
// m16n8k16 warp-level MMA shape (two taps per K step, no zero padding along N), whereas a 128 x 64 tcgen05 tile
// would spend 8x its FLOPs on the zero blocks of the block-diagonal weight and re-fetch the activation tile from L2
// for every tap.  So this kernel stages a (8+2) x (32+2) pixel x 64-channel tile (split-bf16 planes) in shared
// memory ONCE, and each of the 8 warps owns one group of the 64-channel slab: A fragments come straight out of
// the staged tile with ldmatrix (a fragment row = one pixel's 8 input channels of one tap = 16 bytes; 16-byte
// chunks are XOR-swizzled by the pixel column so the 8 rows of a matrix hit 8 different bank groups), B
// fragments (the group's 72 x 8 weights) live in registers for the whole CTA.  SPLIT = 3 is the same
// error-compensated bf16x3 product as the tcgen05 kernels (hi*hi + lo*hi + hi*lo into one fp32 accumulator).
// The result is staged through shared memory and written as whole 128-byte lines.
// HBM-bound by design: reads each activation once (+ halo from L2), writes once.
#include "common.cuh"

namespace gc {

constexpr int TY = 8, TX = 32;                  // output pixels per CTA
constexpr int PY = TY + 2, PX = TX + 2;         // staged tile with the 1-pixel halo
constexpr int SLAB = 64;                        // channels per CTA = 8 groups = 8 warps
constexpr int PIX_BYTES = SLAB * 2;             // 128 B per pixel per plane
constexpr int PLANE_BYTES = PY * PX * PIX_BYTES;  // 43520
constexpr int THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&a)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<uint32_t*>(&h);
    __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<uint32_t*>(&l);
}

template <int SPLIT>
__global__ void __launch_bounds__(THREADS, 2) gconv_mma_kernel(
    const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
    const float* __restrict__ w /*[9][8 ci][8 co][G]*/, const float* __restrict__ bias,
    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int H, int W, int C, int relu) {
    constexpr int NPL = (SPLIT == 3) ? 2 : 1;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = C / 8;
    const int slab = blockIdx.y;
    const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
    const int b = blockIdx.x / (tiles_x * tiles_y);
    const int trem = blockIdx.x - b * tiles_x * tiles_y;
    const int ty0 = (trem / tiles_x) * TY, tx0 = (trem % tiles_x) * TX;

    // ---- stage the input tile: (pixel, 16-byte chunk, plane) -> swizzled shared memory, zero fill outside ----
    {
        // thread = (16-byte chunk c of the 64-channel slab, pixel p0 + 32 k): one constant division per copy and
        // 32-bit offsets inside the frame (the index arithmetic was 26 % of the kernel's instructions)
        const uint32_t sbase = smem_u32(smem);
        const int c = threadIdx.x & 7, p0 = threadIdx.x >> 3;
        const long long frame = (long long)b * H * W * C + slab * SLAB + c * 8;
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) {
            const __nv_bfloat16* base = (pl ? in_lo : in_hi) + frame;
#pragma unroll
            for (int it = 0; it < (PY * PX + 31) / 32; ++it) {
                const int pp = p0 + it * 32;
                if (pp < PY * PX) {
                    const int py = pp / PX, px = pp - py * PX;
                    const int gy = ty0 + py - 1, gx = tx0 + px - 1;
                    const bool inb = ((unsigned)gy < (unsigned)H) && ((unsigned)gx < (unsigned)W);
                    const __nv_bfloat16* src = inb ? base + (gy * W + gx) * C : base;
                    cp_async16(sbase + pl * PLANE_BYTES + pp * PIX_BYTES + ((c ^ (px & 7)) << 4), src, inb ? 16 : 0);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- this warp's group weights as B fragments (registers), split into hi / lo ----
    const int g = slab * 8 + warp;
    const int gid = lane >> 2, tig = lane & 3;
    uint32_t bh[5][2], bl[5][2];
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
            const int tap = 2 * s + hlf;
            float w0 = 0.f, w1 = 0.f;
            if (tap < 9) {
                w0 = __ldg(w + (((long long)tap * 8 + 2 * tig) * 8 + gid) * G + g);
                w1 = __ldg(w + (((long long)tap * 8 + 2 * tig + 1) * 8 + gid) * G + g);
            }
            split2(w0, w1, bh[s][hlf], bl[s][hlf]);
        }
    // per-lane ldmatrix row offsets (constant across the 16-pixel tiles): lane = 8 i + r supplies row r of matrix i
    //   matrix 0: pixels 0-7, first tap of the step   matrix 1: pixels 8-15, first tap
    //   matrix 2: pixels 0-7, second tap              matrix 3: pixels 8-15, second tap
    uint32_t aoff[5];
    {
        const int i = lane >> 3, r = lane & 7;
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            int tap = 2 * s + (i >> 1);
            if (tap > 8) tap = 8;  // K padding: B is zero there, any finite A will do
            const int dy = tap / 3, dx = tap - dy * 3;
            const int col = (i & 1) * 8 + r + dx;
            aoff[s] = (uint32_t)((dy * PX + col) * PIX_BYTES + ((warp ^ (col & 7)) << 4));
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    float acc[16][4];
#pragma unroll
    for (int mt = 0; mt < 16; ++mt) {
        acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
        const uint32_t tb = smem_u32(smem) + (uint32_t)(((mt >> 1) * PX + (mt & 1) * 16) * PIX_BYTES);
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            uint32_t ah[4];
            ldsm4(tb + aoff[s], ah);
            mma16816(acc[mt], ah, bh[s][0], bh[s][1]);
            if (SPLIT == 3) {
                uint32_t al[4];
                ldsm4(tb + PLANE_BYTES + aoff[s], al);
                mma16816(acc[mt], al, bh[s][0], bh[s][1]);
                mma16816(acc[mt], ah, bl[s][0], bl[s][1]);
            }
        }
    }
    __syncthreads();  // every warp is done reading the staged input: reuse it as the output staging tile
    {
        const float b0 = bias ? __ldg(bias + g * 8 + 2 * tig) : 0.f, b1 = bias ? __ldg(bias + g * 8 + 2 * tig + 1) : 0.f;
#pragma unroll
        for (int mt = 0; mt < 16; ++mt)
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                const int op = (mt >> 1) * TX + (mt & 1) * 16 + gid + hlf * 8;  // output pixel inside the tile
                float v0 = acc[mt][2 * hlf] + b0, v1 = acc[mt][2 * hlf + 1] + b1;
                if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                uint32_t hi, lo;
                split2(v0, v1, hi, lo);
                uint8_t* d = smem + op * PIX_BYTES + ((warp ^ (op & 7)) << 4) + tig * 4;
                *reinterpret_cast<uint32_t*>(d) = hi;
                if (SPLIT == 3) *reinterpret_cast<uint32_t*>(d + TY * TX * PIX_BYTES) = lo;
            }
    }
    __syncthreads();
    {
        const int c = threadIdx.x & 7, o0 = threadIdx.x >> 3;
        const long long frame = (long long)b * H * W * C + slab * SLAB + c * 8;
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) {
            __nv_bfloat16* base = (pl ? out_lo : out_hi) + frame;
#pragma unroll
            for (int it = 0; it < TY * TX / 32; ++it) {
                const int op = o0 + it * 32;
                const int gy = ty0 + (op >> 5), gx = tx0 + (op & 31);  // TX == 32
                if (gy < H && gx < W)
                    *reinterpret_cast<uint4*>(base + (gy * W + gx) * C) =
                        *reinterpret_cast<const uint4*>(smem + pl * (TY * TX * PIX_BYTES) + op * PIX_BYTES + ((c ^ (op & 7)) << 4));
            }
        }
    }
}

}  // namespace gc

// in / out: split-bf16 NHWC planes [B][H][W][C] (lo planes NULL when split == 1); w fp32 [9][8][8][C/8]
// (tap, ci, co, group), bias fp32 [C].
PRAM_API int pram_gconv3x3_tc(const void* in_hi, const void* in_lo, const float* w, const float* bias, void* out_hi,
                              void* out_lo, int B, int H, int W, int C, int relu, int split, cudaStream_t stream) {
    using namespace gc;
    if (!in_hi || !w || !out_hi || B <= 0 || H <= 0 || W <= 0) return PRAM_ERR_ARG;
    if (split != 1 && split != 3) return PRAM_ERR_ARG;
    if (split == 3 && (!in_lo || !out_lo)) return PRAM_ERR_ARG;
    if (C % SLAB || (long long)H * W * C >= (1LL << 31)) return PRAM_ERR_UNSUPPORTED;  // 32-bit offsets inside a frame
    static_assert(TX == 32, "the store loop decodes pixels with shifts");
    const int tiles = cdiv(W, TX) * cdiv(H, TY);
    dim3 grid(B * tiles, C / SLAB);
    if (split == 3) {
        auto kern = gconv_mma_kernel<3>;
        static bool attr = false;
        if (!attr) { PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PLANE_BYTES)); attr = true; }
        kern<<<grid, THREADS, 2 * PLANE_BYTES, stream>>>((const __nv_bfloat16*)in_hi, (const __nv_bfloat16*)in_lo, w, bias,
                                                        (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, H, W, C, relu);
    } else {
        auto kern = gconv_mma_kernel<1>;
        static bool attr = false;
        if (!attr) { PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PLANE_BYTES)); attr = true; }
        kern<<<grid, THREADS, PLANE_BYTES, stream>>>((const __nv_bfloat16*)in_hi, nullptr, w, bias, (__nv_bfloat16*)out_hi,
                                                    nullptr, H, W, C, relu);
    }
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}
