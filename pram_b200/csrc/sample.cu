// K8 / K9 of SURVEY.md section 2b: bilinear gather of per-keypoint features from NHWC maps, score gather,
// keypoint normalisation + learnable Fourier positional encoding, rotary application.
// Replaces reference nets/sfd2.py:53-64 and :348-369 (sample_descriptors / ResNet4x.sample),
// nets/utils.py:17-24 (normalize_keypoints), nets/segnetvit.py:15-40 (rotary, Fourier encoding).
#include "common.cuh"

// One warp per keypoint; lanes stride the channel axis with 16-byte loads from the 4 taps (each tap
// is C contiguous floats in an NHWC map).  grid_sample(align_corners=True, zeros padding) semantics.
__global__ void __launch_bounds__(256) sample_kernel(
    const float* __restrict__ fmap, int C, int h, int w, const float* __restrict__ kpts,
    const int* __restrict__ counts, int kpad, float off, float divx, float divy, int normalize,
    float* __restrict__ out, int total) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= total) return;
    const int b = gw / kpad, j = gw - b * kpad;
    float* o = out + (long long)gw * C;
    if (counts && j >= counts[b]) {
        for (int c = lane * 4; c < C; c += 128) *reinterpret_cast<float4*>(o + c) = make_float4(0, 0, 0, 0);
        return;
    }
    const float kx = kpts[2 * (long long)gw], ky = kpts[2 * (long long)gw + 1];
    // reference: k = k - s/2 + 0.5 ; k /= (w*s - s/2 - 0.5, h*s - s/2 - 0.5) ; k = k*2 - 1
    float gx = ((kx - off) / divx) * 2.f - 1.f;
    float gy = ((ky - off) / divy) * 2.f - 1.f;
    // grid_sample unnormalise, align_corners=True
    float ix = ((gx + 1.f) / 2.f) * (float)(w - 1);
    float iy = ((gy + 1.f) / 2.f) * (float)(h - 1);
    float fx0 = floorf(ix), fy0 = floorf(iy);
    int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    float wx1 = ix - fx0, wy1 = iy - fy0;
    float wx0 = (fx0 + 1.f) - ix, wy0 = (fy0 + 1.f) - iy;
    float w_nw = wx0 * wy0, w_ne = wx1 * wy0, w_sw = wx0 * wy1, w_se = wx1 * wy1;
    bool vx0 = x0 >= 0 && x0 < w, vx1 = x1 >= 0 && x1 < w, vy0 = y0 >= 0 && y0 < h, vy1 = y1 >= 0 && y1 < h;
    const float* base = fmap + (long long)b * h * w * C;
    const float* p_nw = base + ((long long)y0 * w + x0) * C;
    const float* p_ne = base + ((long long)y0 * w + x1) * C;
    const float* p_sw = base + ((long long)y1 * w + x0) * C;
    const float* p_se = base + ((long long)y1 * w + x1) * C;
    float4 acc[2];
    float ss = 0.f;
    int it = 0;
    for (int c = lane * 4; c < C; c += 128, ++it) {
        float4 z = make_float4(0, 0, 0, 0);
        float4 a = (vy0 && vx0) ? *reinterpret_cast<const float4*>(p_nw + c) : z;
        float4 bq = (vy0 && vx1) ? *reinterpret_cast<const float4*>(p_ne + c) : z;
        float4 cq = (vy1 && vx0) ? *reinterpret_cast<const float4*>(p_sw + c) : z;
        float4 d = (vy1 && vx1) ? *reinterpret_cast<const float4*>(p_se + c) : z;
        float4 r;
        r.x = a.x * w_nw + bq.x * w_ne + cq.x * w_sw + d.x * w_se;
        r.y = a.y * w_nw + bq.y * w_ne + cq.y * w_sw + d.y * w_se;
        r.z = a.z * w_nw + bq.z * w_ne + cq.z * w_sw + d.z * w_se;
        r.w = a.w * w_nw + bq.w * w_ne + cq.w * w_sw + d.w * w_se;
        ss += r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
        if (it < 2) acc[it] = r;
    }
    float denom = 1.f;
    if (normalize) denom = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);  // F.normalize: x / max(||x||, eps)
    it = 0;
    for (int c = lane * 4; c < C; c += 128, ++it) {
        float4 r = acc[it];
        if (normalize) { r.x /= denom; r.y /= denom; r.z /= denom; r.w /= denom; }
        *reinterpret_cast<float4*>(o + c) = r;
    }
}

PRAM_API int pram_sample_features(const float* fmap, int B, int C, int h, int w, const float* kpts,
                                  const int* counts, int kpad, int s, int normalize, float* out,
                                  cudaStream_t stream) {
    if (!fmap || !kpts || !out || B <= 0 || kpad <= 0) return PRAM_ERR_ARG;
    if (C % 4 != 0 || C > 256) return PRAM_ERR_UNSUPPORTED;
    // the divisors are formed in double precision by the reference (python floats) and then cast
    float off = (float)(s / 2.0 - 0.5);
    float divx = (float)(w * (double)s - s / 2.0 - 0.5), divy = (float)(h * (double)s - s / 2.0 - 0.5);
    int total = B * kpad;
    sample_kernel<<<cdiv((long long)total * 32, 256), 256, 0, stream>>>(fmap, C, h, w, kpts, counts, kpad,
                                                                       off, divx, divy, normalize, out, total);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// scores[b][j] = score_map[b][(int)y][(int)x]   (reference nets/sfd2.py:367)
__global__ void gather_scores_kernel(const float* __restrict__ score, int H, int W,
                                     const float* __restrict__ kpts, const int* __restrict__ counts,
                                     int kpad, float* __restrict__ out, int total, int batch_stride_on) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int b = i / kpad, j = i - b * kpad;
    if (counts && j >= counts[b]) { out[i] = 0.f; return; }
    int x = (int)kpts[2 * (long long)i], y = (int)kpts[2 * (long long)i + 1];
    x = min(max(x, 0), W - 1);
    y = min(max(y, 0), H - 1);
    out[i] = score[((long long)(batch_stride_on ? b : 0) * H + y) * W + x];
}

PRAM_API int pram_gather_scores(const float* score, int B, int H, int W, const float* kpts,
                                const int* counts, int kpad, float* out, cudaStream_t stream) {
    if (!score || !kpts || !out) return PRAM_ERR_ARG;
    int total = B * kpad;
    gather_scores_kernel<<<cdiv(total, 256), 256, 0, stream>>>(score, H, W, kpts, counts, kpad, out, total, 1);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// K9: normalize_keypoints + Fourier encoding.  One thread per (token, frequency).
//   nk = (k - (W,H)/2) / (0.7*max(W,H));  proj_f = nk . Wr[f];  cos_out/sin_out [tokens][32]
__global__ void posenc_kernel(const float* __restrict__ kpts, int total, float cx, float cy, float scale,
                              int prenormalized, const float* __restrict__ Wr,
                              float* __restrict__ cos_out, float* __restrict__ sin_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total * 32) return;
    int t = i >> 5, f = i & 31;
    float x = kpts[2 * (long long)t], y = kpts[2 * (long long)t + 1];
    if (!prenormalized) { x = (x - cx) / scale; y = (y - cy) / scale; }
    // torch Linear(2,32) on CPU: dot of length 2 (x*w0 + y*w1)
    float p = x * Wr[2 * f] + y * Wr[2 * f + 1];
    float s, c;
    sincosf(p, &s, &c);
    cos_out[i] = c;
    sin_out[i] = s;
}

PRAM_API int pram_posenc(const float* kpts, int tokens, float width, float height, int prenormalized,
                         const float* Wr, float* cos_out, float* sin_out, cudaStream_t stream) {
    if (!kpts || !Wr || !cos_out || !sin_out || tokens <= 0) return PRAM_ERR_ARG;
    float cx = width / 2.f, cy = height / 2.f;
    float scale = fmaxf(width, height) * 0.7f;
    posenc_kernel<<<cdiv((long long)tokens * 32, 256), 256, 0, stream>>>(kpts, tokens, cx, cy, scale,
                                                                        prenormalized, Wr, cos_out, sin_out);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}
