// K5 / K5b / K6 / K7 of SURVEY.md section 2b: detector-logit softmax + pixel shuffle, optional bilinear
// resize, 2-round non-maximum suppression with candidate emission, and per-frame keypoint selection.
// Replaces reference nets/sfd2.py:294-329 (score map, simple_nms :20-35, threshold/border/top-k
// :306-329).  All HBM-bound: coalesced vector loads, shared-memory staging, no host sync.
#include "common.cuh"
#include "nms_tile.cuh"

// ------------------------------------------------------------------------------------------
// K5: softmax over 65 detector channels, drop the dustbin, 8x8 pixel shuffle.
//   block = 256 threads = 8 warps; a block owns 32 consecutive coarse cells of one coarse row
//   (each warp 4 cells), stages the 8 x 256 probabilities in shared memory and writes full lines.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) score_map_kernel(
    const float* __restrict__ logits, long long bs, long long ys, long long xs, long long cs,
    int Hc, int Wc, float* __restrict__ out) {
    __shared__ float tile[8][32 * 8 + 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hc = blockIdx.y, b = blockIdx.z;
    const int wc0 = blockIdx.x * 32;
    const float* base = logits + (long long)b * bs + (long long)hc * ys;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int cell = warp * 4 + k;
        const int wc = wc0 + cell;
        if (wc < Wc) {  // warp-uniform
            const float* p = base + (long long)wc * xs;
            float a0 = p[(long long)lane * cs];
            float a1 = p[(long long)(lane + 32) * cs];
            float a2 = p[64 * cs];  // dustbin, same value in all lanes
            float m = warp_max(fmaxf(fmaxf(a0, a1), a2));
            float e0 = expf(a0 - m), e1 = expf(a1 - m), e2 = expf(a2 - m);
            float s = warp_sum(e0 + e1) + e2;
            // channel c -> (row c>>3, col c&7) inside the 8x8 patch
            tile[lane >> 3][cell * 8 + (lane & 7)] = e0 / s;
            tile[(lane >> 3) + 4][cell * 8 + (lane & 7)] = e1 / s;
        }
    }
    __syncthreads();
    const int W = Wc * 8;
    const int r = threadIdx.x >> 5;  // 0..7
    float* orow = out + ((long long)b * Hc * 8 + hc * 8 + r) * (long long)W + wc0 * 8;
    const int ncols = min(32, Wc - wc0) * 8;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        int c = (threadIdx.x & 31) * 4 + k * 128;
        if (c < ncols) {
            float4 v = make_float4(tile[r][c], tile[r][c + 1], tile[r][c + 2], tile[r][c + 3]);
            *reinterpret_cast<float4*>(orow + c) = v;
        }
    }
}

PRAM_API int pram_score_map(const float* logits, long long batch_stride, long long y_stride,
                            long long x_stride, long long ch_stride, int B, int Hc, int Wc,
                            float* score, cudaStream_t stream) {
    if (!logits || !score || B <= 0 || Hc <= 0 || Wc <= 0) return PRAM_ERR_ARG;
    dim3 grid(cdiv(Wc, 32), Hc, B);
    score_map_kernel<<<grid, 256, 0, stream>>>(logits, batch_stride, y_stride, x_stride, ch_stride,
                                               Hc, Wc, score);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// K5b: bilinear resize, align_corners=True (reference nets/sfd2.py:301-303).
__global__ void resize_bilinear_kernel(const float* __restrict__ in, int B, int Hi, int Wi,
                                       float* __restrict__ out, int Ho, int Wo, float sh, float sw) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * Wo;
    if (i >= total) return;
    int x = (int)(i % Wo);
    int y = (int)((i / Wo) % Ho);
    int b = (int)(i / ((long long)Wo * Ho));
    float fy = sh * y, fx = sw * x;
    int y0 = (int)fy, x0 = (int)fx;
    int y1 = y0 + (y0 < Hi - 1), x1 = x0 + (x0 < Wi - 1);
    float ly = fy - y0, lx = fx - x0;
    float hy = 1.f - ly, hx = 1.f - lx;
    const float* p = in + (long long)b * Hi * Wi;
    out[i] = hy * (hx * p[(long long)y0 * Wi + x0] + lx * p[(long long)y0 * Wi + x1]) +
             ly * (hx * p[(long long)y1 * Wi + x0] + lx * p[(long long)y1 * Wi + x1]);
}

PRAM_API int pram_resize_bilinear(const float* in, int B, int Hi, int Wi, float* out, int Ho, int Wo,
                                  cudaStream_t stream) {
    if (!in || !out || B <= 0) return PRAM_ERR_ARG;
    float sh = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
    float sw = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
    long long total = (long long)B * Ho * Wo;
    resize_bilinear_kernel<<<cdiv(total, 256), 256, 0, stream>>>(in, B, Hi, Wi, out, Ho, Wo, sh, sw);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// ------------------------------------------------------------------------------------------
// K6: simple_nms (reference nets/sfd2.py:20-35), exact semantics:
//   keep0 = s == maxpool(s);  2x { supp = dilate(keep); rest = supp ? 0 : s;
//                                  keep |= (rest == maxpool(rest)) & ~supp }
//   windows are (2r+1)^2 and pixels outside the image behave as -inf / false.
// The result at a pixel depends on raw scores within 5r, so a TH x TW output tile is computed from a
// (TH+10r) x (TW+10r) shared-memory tile in one launch (the reference runs 5 max-pools + ~12
// elementwise passes).  Max-pools are separable (row pass into T, column pass fused with the
// equality test).  Candidates (score >= th_lo) are emitted as 64-bit keys (score bits << 32 | index)
// with warp-aggregated atomics; the number of pixels >= th_hi is counted for the reference's
// "too few keypoints -> halve the threshold" rule (nets/sfd2.py:311-315).
// ------------------------------------------------------------------------------------------
// Implementation: nms_tile.cuh (phase functions shared with the host emulation in tests/nms_host.cu).
// ------------------------------------------------------------------------------------------
constexpr int NMS_MAXR = 4;
// threads per CTA: 24 warps on the tall tile (1 CTA / SM; the passes are latency-bound chains of shared-memory loads),
// 12 on the short one (2 CTAs / SM)
template <int TH> struct NmsThreads { static constexpr int value = (TH >= 96) ? 768 : 384; };

template <int R, int TH>
__global__ void __launch_bounds__(NmsThreads<TH>::value) nms_kernel(
    const float* __restrict__ score, int H, int W, float th_lo, float th_hi,
    float* __restrict__ nms_out, unsigned long long* __restrict__ cand, int cap,
    int* __restrict__ cand_count, int* __restrict__ count_hi) {
    using G = NmsGeom<R, TH>;
    constexpr int NMS_THREADS = NmsThreads<TH>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int block_count, block_base, block_hi;
    NmsTile t;
    t.S = reinterpret_cast<float*>(smem_raw);
    t.T = t.S + G::S_FLOATS;
    t.keep = reinterpret_cast<unsigned char*>(t.T + G::T_FLOATS);
    t.supp = t.keep + G::MASK_BYTES;
    t.tmpb = t.supp + G::MASK_BYTES;
    t.x0 = blockIdx.x * G::TW - G::HALO;
    t.y0 = blockIdx.y * G::TH - G::HALO;
    t.H = H; t.W = W;
    const int b = blockIdx.z;
    t.score = score + (long long)b * H * W;
    const int tid = threadIdx.x;
    if (tid == 0) { block_count = 0; block_hi = 0; }
    constexpr int N_INIT = (8 * G::SW / 4 > G::MASK_BYTES / 4) ? 8 * G::SW / 4 : G::MASK_BYTES / 4;
    for (int i = tid; i < N_INIT; i += NMS_THREADS) nms_init<G>(t, i);
    {   // tile load, 4 independent 16-byte global loads in flight per thread
        constexpr int NLOAD = G::SH * (G::NG + 2);
        for (int i0 = tid; i0 < NLOAD; i0 += 4 * NMS_THREADS) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * NMS_THREADS < NLOAD) v[u] = nms_load_value<G>(t, i0 + u * NMS_THREADS);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * NMS_THREADS < NLOAD) nms_store_value<G>(t, i0 + u * NMS_THREADS, v[u]);
        }
    }
    __syncthreads();
    for (int i = tid; i < G::SH * (G::SW / 8); i += NMS_THREADS) nms_rowmax<G, R, false>(t, i);
    __syncthreads();
    for (int i = tid; i < (G::SH / 8) * G::NG; i += NMS_THREADS) nms_colmax<G, R, false>(t, i);
    __syncthreads();
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        for (int i = tid; i < G::SH * G::NG; i += NMS_THREADS) nms_dilate_h<G, R>(t, i);
        __syncthreads();
        for (int i = tid; i < G::SH * (G::GW / 4); i += NMS_THREADS) nms_dilate_v<G, R>(t, i);
        __syncthreads();
        for (int i = tid; i < G::SH * (G::SW / 8); i += NMS_THREADS) nms_rowmax<G, R, true>(t, i);
        __syncthreads();
        for (int i = tid; i < (G::SH / 8) * G::NG; i += NMS_THREADS) nms_colmax<G, R, true>(t, i);
        __syncthreads();
    }
    // ---- epilogue over the inner tile: nms map, counts, candidate emission (two passes over the nibbles) ----
    int my_n = 0, my_hi = 0;
    for (int i = tid; i < G::TH * 32; i += NMS_THREADS) {
        int gy, gx;
        const float4 v4 = nms_result<G>(t, i, gy, gx);
        if (gy >= H || gx >= W) continue;
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
        if (nms_out) {
            float* op = nms_out + ((long long)b * H + gy) * W + gx;
            if (gx + 3 < W && (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
                *reinterpret_cast<float4*>(op) = v4;
            } else {
#pragma unroll
                for (int x = 0; x < 4; ++x)
                    if (gx + x < W) op[x] = v[x];
            }
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
            if (gx + x < W) {
                my_hi += (v[x] >= th_hi);
                my_n += (v[x] >= th_lo && v[x] > 0.f);
            }
    }
    if (my_hi) atomicAdd(&block_hi, my_hi);
    int my_off = my_n ? atomicAdd(&block_count, my_n) : 0;
    __syncthreads();
    if (tid == 0) {
        block_base = block_count ? atomicAdd(&cand_count[b], block_count) : 0;
        if (block_hi) atomicAdd(&count_hi[b], block_hi);
    }
    __syncthreads();
    if (my_n) {
        int pos = block_base + my_off;
        for (int i = tid; i < G::TH * 32; i += NMS_THREADS) {
            int gy, gx;
            const float4 v4 = nms_result<G>(t, i, gy, gx);
            if (gy >= H || gx >= W) continue;
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
                if (gx + x < W && v[x] >= th_lo && v[x] > 0.f) {
                    if (pos < cap)
                        cand[(long long)b * cap + pos] = ((unsigned long long)__float_as_uint(v[x]) << 32) |
                                                        (unsigned int)(gy * W + gx + x);
                    ++pos;
                }
        }
    }
}

template <int R, int TH>
static int nms_launch(const float* score, int B, int H, int W, float th_lo, float th_hi, float* nms_out,
                      unsigned long long* cand, int cap, int* cand_count, int* count_hi, cudaStream_t stream) {
    using G = NmsGeom<R, TH>;
    auto kern = nms_kernel<R, TH>;
    static bool attr_set = false;
    if (!attr_set) {
        PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
        attr_set = true;
    }
    dim3 grid(cdiv(W, G::TW), cdiv(H, G::TH), B);
    kern<<<grid, NmsThreads<TH>::value, G::SMEM, stream>>>(score, H, W, th_lo, th_hi, nms_out, cand, cap, cand_count, count_hi);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

template <int R>
static int nms_dispatch(const float* score, int B, int H, int W, float th_lo, float th_hi, float* nms_out,
                        unsigned long long* cand, int cap, int* cand_count, int* count_hi, cudaStream_t stream) {
    // tall tiles (1.9x halo redundancy, 1 CTA/SM) once they fill the GPU twice over; short tiles (2 CTAs/SM,
    // 4x more CTAs) for small batches / single-frame latency
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const long long tall = (long long)cdiv(W, 128) * cdiv(H, 96) * B;
    if (tall >= 2LL * sms)
        return nms_launch<R, 96>(score, B, H, W, th_lo, th_hi, nms_out, cand, cap, cand_count, count_hi, stream);
    return nms_launch<R, 24>(score, B, H, W, th_lo, th_hi, nms_out, cand, cap, cand_count, count_hi, stream);
}

PRAM_API int pram_nms_candidates(const float* score, int B, int H, int W, int radius, float th_lo,
                                 float th_hi, float* nms_out, unsigned long long* cand, int cap,
                                 int* cand_count, int* count_hi, cudaStream_t stream) {
    if (!score || !cand || !cand_count || !count_hi || B <= 0 || radius < 0 || radius > NMS_MAXR)
        return PRAM_ERR_ARG;
    PRAM_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(int) * B, stream));
    PRAM_CUDA(cudaMemsetAsync(count_hi, 0, sizeof(int) * B, stream));
    switch (radius) {
        case 0: return nms_dispatch<0>(score, B, H, W, th_lo, th_hi, nms_out, cand, cap, cand_count, count_hi, stream);
        case 1: return nms_dispatch<1>(score, B, H, W, th_lo, th_hi, nms_out, cand, cap, cand_count, count_hi, stream);
        case 2: return nms_dispatch<2>(score, B, H, W, th_lo, th_hi, nms_out, cand, cap, cand_count, count_hi, stream);
        case 3: return nms_dispatch<3>(score, B, H, W, th_lo, th_hi, nms_out, cand, cap, cand_count, count_hi, stream);
        default: return nms_dispatch<4>(score, B, H, W, th_lo, th_hi, nms_out, cand, cap, cand_count, count_hi, stream);
    }
}

// ------------------------------------------------------------------------------------------
// K7: per-frame selection (reference nets/sfd2.py:306-329), one CTA per frame, no host sync.
//   eff_th = (count(nms >= th_hi) <= min_keypoints) ? th_lo : th_hi          (:311-315)
//   valid  = score >= eff_th  &&  border <= y < H-border  &&  border <= x < W-border   (:38-43)
//   n <= K : row-major (y,x) order                                              (:47-48)
//   n >  K : top-K by score, descending (ties: lower index first)               (:49-50)
//   output (x,y) float32                                                        (:329)
// Top-K = 8-pass MSB radix select of the K-th largest unique 64-bit key, then a bitonic sort of the
// K survivors in shared memory.
// ------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 1024;
constexpr int SEL_MAXK = 4096;

// border window: border <= y < y_hi, border <= x < x_hi.  The upper bounds are separate arguments because the export path
// (nets/sfd2.py:447-451) tests the keypoints of a RESCALED image against the ORIGINAL width / height.
struct SelWin { int border, y_hi, x_hi, W; };
__device__ __forceinline__ bool sel_valid(unsigned long long key, float eff_th, const SelWin& w) {
    float v = __uint_as_float((unsigned int)(key >> 32));
    unsigned int idx = (unsigned int)key;
    int y = idx / w.W, x = idx - y * w.W;
    return v >= eff_th && y >= w.border && y < w.y_hi && x >= w.border && x < w.x_hi;
}

__global__ void __launch_bounds__(SEL_THREADS) select_kernel(
    const unsigned long long* __restrict__ cand, int cap, const int* __restrict__ cand_count,
    const int* __restrict__ count_hi, float th_lo, float th_hi, int min_kp, int K, SelWin win, int H,
    int W, float* __restrict__ kpts, float* __restrict__ scores, int* __restrict__ n_out, int kpad,
    int* __restrict__ n_valid_out) {
    __shared__ unsigned long long keys[SEL_MAXK];
    __shared__ int hist[256];
    __shared__ int s_n, s_pos;
    __shared__ unsigned long long s_prefix;
    __shared__ int s_remaining;
    const int b = blockIdx.x;
    const unsigned long long* c = cand + (long long)b * cap;
    const int n_all = min(cand_count[b], cap);
    const float eff_th = (count_hi[b] <= min_kp) ? th_lo : th_hi;
    if (threadIdx.x == 0) { s_n = 0; s_pos = 0; }
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < n_all; i += SEL_THREADS) cnt += sel_valid(c[i], eff_th, win);
    cnt = (int)warp_sum((float)cnt);  // exact for counts < 2^24
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_n, cnt);
    __syncthreads();
    const int n_valid = s_n;
    if (threadIdx.x == 0 && n_valid_out) n_valid_out[b] = n_valid;   // lets the host see "more valid keypoints than slots"
    // "unlimited" (K < 0) with more valid keypoints than output slots: fall back to the best `kpad` by score -- a
    // deterministic subset (the gather below would otherwise keep whichever candidates won the atomics)
    if (K < 0 && n_valid > min(kpad, SEL_MAXK)) K = min(kpad, SEL_MAXK);
    const bool take_all = (K < 0) || (n_valid <= K);
    int n_sel = take_all ? n_valid : K;
    if (n_sel > SEL_MAXK) n_sel = SEL_MAXK;  // host guarantees K <= SEL_MAXK and kpad <= SEL_MAXK
    if (n_sel > kpad) n_sel = kpad;

    unsigned long long kth = 0;  // selection threshold on the composite key (score desc, idx asc)
    if (!take_all) {
        // composite key: score bits in the high word, ~idx in the low word -> unique, larger = better
        if (threadIdx.x == 0) { s_prefix = 0; s_remaining = K; }
        __syncthreads();
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            for (int i = threadIdx.x; i < 256; i += SEL_THREADS) hist[i] = 0;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
            for (int i = threadIdx.x; i < n_all; i += SEL_THREADS) {
                unsigned long long k = c[i];
                if (!sel_valid(k, eff_th, win)) continue;
                unsigned long long ck = (k & 0xffffffff00000000ull) | (unsigned int)(~(unsigned int)k);
                if ((ck & himask) == prefix) atomicAdd(&hist[(ck >> shift) & 0xff], 1);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int rem = s_remaining, d = 255;
                for (; d > 0; --d) {
                    if (hist[d] >= rem) break;
                    rem -= hist[d];
                }
                s_remaining = rem;
                s_prefix = prefix | ((unsigned long long)d << shift);
            }
            __syncthreads();
        }
        kth = s_prefix;
    }
    // gather
    for (int i = threadIdx.x; i < n_all; i += SEL_THREADS) {
        unsigned long long k = c[i];
        if (!sel_valid(k, eff_th, win)) continue;
        unsigned long long ck = (k & 0xffffffff00000000ull) | (unsigned int)(~(unsigned int)k);
        if (take_all) {
            // ascending index == descending ~idx: reuse the descending sort on the low word alone
            int pos = atomicAdd(&s_pos, 1);
            if (pos < SEL_MAXK) keys[pos] = (unsigned long long)(unsigned int)(~(unsigned int)k);
        } else if (ck >= kth) {
            int pos = atomicAdd(&s_pos, 1);
            if (pos < SEL_MAXK) keys[pos] = ck;
        }
    }
    __syncthreads();
    int m = min(s_pos, SEL_MAXK);
    int P = 1;
    while (P < m) P <<= 1;
    for (int i = m + threadIdx.x; i < P; i += SEL_THREADS) keys[i] = 0ull;  // pad: sorts last
    __syncthreads();
    // bitonic sort, descending
    for (int k2 = 2; k2 <= P; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += SEL_THREADS) {
                int l = i ^ j;
                if (l > i) {
                    unsigned long long a = keys[i], bb = keys[l];
                    bool desc = ((i & k2) == 0);
                    if (desc ? (a < bb) : (a > bb)) { keys[i] = bb; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    // write out (x,y); scores are gathered from the score map by fill_scores_kernel (an NMS-kept
    // value IS the score-map value at that pixel)
    for (int i = threadIdx.x; i < kpad; i += SEL_THREADS) {
        float x = 0.f, y = 0.f;
        if (i < n_sel && i < m) {
            unsigned int idx = ~(unsigned int)keys[i];
            y = (float)(idx / W);
            x = (float)(idx - (idx / W) * W);
        }
        kpts[((long long)b * kpad + i) * 2 + 0] = x;
        kpts[((long long)b * kpad + i) * 2 + 1] = y;
        scores[(long long)b * kpad + i] = 0.f;
    }
    if (threadIdx.x == 0) n_out[b] = min(n_sel, m);
}

// Scores of the selected keypoints: a gather from the score map.
__global__ void fill_scores_kernel(const float* __restrict__ score, int H, int W,
                                   const float* __restrict__ kpts, const int* __restrict__ n_out,
                                   int kpad, float* __restrict__ scores, int total) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int b = i / kpad, j = i - b * kpad;
    if (j >= n_out[b]) return;
    int x = (int)kpts[2 * (long long)i], y = (int)kpts[2 * (long long)i + 1];
    scores[i] = score[((long long)b * H + y) * W + x];
}

PRAM_API int pram_select_keypoints(const unsigned long long* cand, int cap, const int* cand_count,
                                   const int* count_hi, const float* score, int B, int H, int W,
                                   float th_lo, float th_hi, int min_keypoints, int max_keypoints,
                                   int border, int y_hi, int x_hi, float* kpts, float* scores, int* n_out, int kpad,
                                   int* n_valid_out, cudaStream_t stream) {
    if (!cand || !cand_count || !count_hi || !score || !kpts || !scores || !n_out) return PRAM_ERR_ARG;
    if (kpad <= 0 || kpad > SEL_MAXK || max_keypoints > SEL_MAXK) return PRAM_ERR_UNSUPPORTED;
    SelWin win;
    win.border = border; win.W = W;
    win.y_hi = y_hi > 0 ? y_hi : H - border;   // <= 0: the symmetric window of extract_local_global (nets/sfd2.py:38-43)
    win.x_hi = x_hi > 0 ? x_hi : W - border;
    select_kernel<<<B, SEL_THREADS, 0, stream>>>(cand, cap, cand_count, count_hi, th_lo, th_hi,
                                                 min_keypoints, max_keypoints, win, H, W, kpts,
                                                 scores, n_out, kpad, n_valid_out);
    PRAM_CHECK_LAUNCH();
    fill_scores_kernel<<<cdiv((long long)B * kpad, 256), 256, 0, stream>>>(score, H, W, kpts, n_out,
                                                                          kpad, scores, B * kpad);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}
