// K5 / K5b / K6 / K7 of SURVEY.md section 2b: detector-logit softmax + pixel shuffle, optional bilinear
// resize, 2-round non-maximum suppression with candidate emission, and per-frame keypoint selection.
// Replaces reference nets/sfd2.py:294-329 (score map, simple_nms :20-35, threshold/border/top-k
// :306-329).  All HBM-bound: coalesced vector loads, shared-memory staging, no host sync.
#include "common.cuh"

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
constexpr int NMS_TW = 64, NMS_TH = 32, NMS_MAXR = 4;
constexpr int NMS_THREADS = 512;

struct NmsSmem {
    // sized for the largest halo (r = 4 -> 20)
    static constexpr int HALO = 5 * NMS_MAXR;
    static constexpr int SW = NMS_TW + 2 * HALO, SH = NMS_TH + 2 * HALO;
    float S[SH * SW];
    float T[SH * SW];
    unsigned char keep[SH * SW];
    unsigned char supp[SH * SW];
    unsigned char tb[SH * SW];
    int block_count;
    int block_base;
    int block_hi;
};

__global__ void __launch_bounds__(NMS_THREADS) nms_kernel(
    const float* __restrict__ score, int H, int W, int r, float th_lo, float th_hi,
    float* __restrict__ nms_out, unsigned long long* __restrict__ cand, int cap,
    int* __restrict__ cand_count, int* __restrict__ count_hi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem& sm = *reinterpret_cast<NmsSmem*>(smem_raw);
    const int halo = 5 * r;
    const int SW = NMS_TW + 2 * halo, SH = NMS_TH + 2 * halo;
    const int n = SW * SH;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * NMS_TW - halo, y0 = blockIdx.y * NMS_TH - halo;
    const float* sc = score + (long long)b * H * W;
    const float NEG = -INFINITY;
    if (threadIdx.x == 0) { sm.block_count = 0; sm.block_hi = 0; }
    for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
        int ly = i / SW, lx = i - ly * SW;
        int gy = y0 + ly, gx = x0 + lx;
        sm.S[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? sc[(long long)gy * W + gx] : NEG;
    }
    __syncthreads();
    // ---- keep0 ----
    for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
        int ly = i / SW, lx = i - ly * SW;
        int a = max(lx - r, 0), e = min(lx + r, SW - 1);
        float m = NEG;
        for (int x = a; x <= e; ++x) m = fmaxf(m, sm.S[ly * SW + x]);
        sm.T[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
        int ly = i / SW, lx = i - ly * SW;
        int a = max(ly - r, 0), e = min(ly + r, SH - 1);
        float m = NEG;
        for (int y = a; y <= e; ++y) m = fmaxf(m, sm.T[y * SW + lx]);
        float s = sm.S[i];
        sm.keep[i] = (s == m) && (s > NEG);
    }
    __syncthreads();
    for (int round = 0; round < 2; ++round) {
        // dilate keep -> supp
        for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
            int ly = i / SW, lx = i - ly * SW;
            int a = max(lx - r, 0), e = min(lx + r, SW - 1);
            unsigned char v = 0;
            for (int x = a; x <= e; ++x) v |= sm.keep[ly * SW + x];
            sm.tb[i] = v;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
            int ly = i / SW, lx = i - ly * SW;
            int a = max(ly - r, 0), e = min(ly + r, SH - 1);
            unsigned char v = 0;
            for (int y = a; y <= e; ++y) v |= sm.tb[y * SW + lx];
            sm.supp[i] = v;
        }
        __syncthreads();
        // rest = supp ? 0 : s  (outside the image stays -inf); row max
        for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
            int ly = i / SW, lx = i - ly * SW;
            int a = max(lx - r, 0), e = min(lx + r, SW - 1);
            float m = NEG;
            for (int x = a; x <= e; ++x) {
                float s = sm.S[ly * SW + x];
                float rest = (sm.supp[ly * SW + x] && s > NEG) ? 0.f : s;
                m = fmaxf(m, rest);
            }
            sm.T[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
            int ly = i / SW, lx = i - ly * SW;
            int a = max(ly - r, 0), e = min(ly + r, SH - 1);
            float m = NEG;
            for (int y = a; y <= e; ++y) m = fmaxf(m, sm.T[y * SW + lx]);
            float s = sm.S[i];
            bool sp = sm.supp[i];
            float rest = (sp && s > NEG) ? 0.f : s;
            if ((rest == m) && !sp && (s > NEG)) sm.keep[i] = 1;
        }
        __syncthreads();
    }
    // ---- epilogue over the inner tile: nms map, candidate emission, counts ----
    int my_hi = 0;
    unsigned long long my_keys[(NMS_TW * NMS_TH + NMS_THREADS - 1) / NMS_THREADS];
    int my_n = 0;
    for (int i = threadIdx.x; i < NMS_TW * NMS_TH; i += NMS_THREADS) {
        int ty = i / NMS_TW, tx = i - ty * NMS_TW;
        int gy = blockIdx.y * NMS_TH + ty, gx = blockIdx.x * NMS_TW + tx;
        if (gy < H && gx < W) {
            int li = (ty + halo) * SW + tx + halo;
            float v = sm.keep[li] ? sm.S[li] : 0.f;
            if (nms_out) nms_out[((long long)b * H + gy) * W + gx] = v;
            if (v >= th_hi) ++my_hi;
            if (v >= th_lo && v > 0.f)
                my_keys[my_n++] = ((unsigned long long)__float_as_uint(v) << 32) |
                                  (unsigned int)(gy * W + gx);
        }
    }
    if (my_hi) atomicAdd(&sm.block_hi, my_hi);
    int my_off = my_n ? atomicAdd(&sm.block_count, my_n) : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        sm.block_base = sm.block_count ? atomicAdd(&cand_count[b], sm.block_count) : 0;
        if (sm.block_hi) atomicAdd(&count_hi[b], sm.block_hi);
    }
    __syncthreads();
    for (int k = 0; k < my_n; ++k) {
        int pos = sm.block_base + my_off + k;
        if (pos < cap) cand[(long long)b * cap + pos] = my_keys[k];
    }
}

PRAM_API int pram_nms_candidates(const float* score, int B, int H, int W, int radius, float th_lo,
                                 float th_hi, float* nms_out, unsigned long long* cand, int cap,
                                 int* cand_count, int* count_hi, cudaStream_t stream) {
    if (!score || !cand || !cand_count || !count_hi || B <= 0 || radius < 0 || radius > NMS_MAXR)
        return PRAM_ERR_ARG;
    static bool attr_set = false;
    if (!attr_set) {
        PRAM_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(NmsSmem)));
        attr_set = true;
    }
    PRAM_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(int) * B, stream));
    PRAM_CUDA(cudaMemsetAsync(count_hi, 0, sizeof(int) * B, stream));
    dim3 grid(cdiv(W, NMS_TW), cdiv(H, NMS_TH), B);
    nms_kernel<<<grid, NMS_THREADS, sizeof(NmsSmem), stream>>>(score, H, W, radius, th_lo, th_hi,
                                                              nms_out, cand, cap, cand_count, count_hi);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
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

__device__ __forceinline__ bool sel_valid(unsigned long long key, float eff_th, int border, int H,
                                          int W) {
    float v = __uint_as_float((unsigned int)(key >> 32));
    unsigned int idx = (unsigned int)key;
    int y = idx / W, x = idx - y * W;
    return v >= eff_th && y >= border && y < H - border && x >= border && x < W - border;
}

__global__ void __launch_bounds__(SEL_THREADS) select_kernel(
    const unsigned long long* __restrict__ cand, int cap, const int* __restrict__ cand_count,
    const int* __restrict__ count_hi, float th_lo, float th_hi, int min_kp, int K, int border, int H,
    int W, float* __restrict__ kpts, float* __restrict__ scores, int* __restrict__ n_out, int kpad) {
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
    for (int i = threadIdx.x; i < n_all; i += SEL_THREADS) cnt += sel_valid(c[i], eff_th, border, H, W);
    cnt = (int)warp_sum((float)cnt);  // exact for counts < 2^24
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_n, cnt);
    __syncthreads();
    const int n_valid = s_n;
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
                if (!sel_valid(k, eff_th, border, H, W)) continue;
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
        if (!sel_valid(k, eff_th, border, H, W)) continue;
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
                                   int border, float* kpts, float* scores, int* n_out, int kpad,
                                   cudaStream_t stream) {
    if (!cand || !cand_count || !count_hi || !score || !kpts || !scores || !n_out) return PRAM_ERR_ARG;
    if (kpad <= 0 || kpad > SEL_MAXK || max_keypoints > SEL_MAXK) return PRAM_ERR_UNSUPPORTED;
    select_kernel<<<B, SEL_THREADS, 0, stream>>>(cand, cap, cand_count, count_hi, th_lo, th_hi,
                                                 min_keypoints, max_keypoints, border, H, W, kpts,
                                                 scores, n_out, kpad);
    PRAM_CHECK_LAUNCH();
    fill_scores_kernel<<<cdiv((long long)B * kpad, 256), 256, 0, stream>>>(score, H, W, kpts, n_out,
                                                                          kpad, scores, B * kpad);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}
