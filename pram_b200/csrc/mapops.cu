// "Next" rows of SURVEY.md section 8f: the device-side glue between recognition and matching, and the
// projection-based pose refinement (K18).
//   segmentation_kernel   : Frame.add_segmentations (reference localization/frame.py:96-121): softmax over the
//                           landmark logits, background probability, arg-max label - 1, background pre-filter mask
//   rank_landmarks_kernel : MultiMap3D.process_segmentations (reference localization/multimap3d.py:348-379):
//                           greedy ranking of candidate landmarks by (rank of the label in each keypoint's
//                           sorted logits, number of keypoints voting for it), mean logit as score
//   project_points_kernel + proj_top2_kernel : SingleMap3D.refine_pose_by_projection (reference
//                           localization/singlemap3d.py:405-440): project the covisible map points with the
//                           current pose, descriptor distance sqrt(2 - 2 q.d + 1e-6) (+100 outside a 2*th
//                           reprojection window), top-2 + ratio test.  The M x N similarity comes from the
//                           tcgen05 GEMM; this kernel fuses mask + distance + top-2 + ratio (no M x N temporaries).
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// one warp per keypoint; logits [T][C]
// ------------------------------------------------------------------------------------------
__global__ void segmentation_kernel(const float* __restrict__ logits, int T, int C, float bg_th,
                                    float* __restrict__ probs /*optional [T][C]*/, float* __restrict__ bg_prob,
                                    int* __restrict__ seg_id, unsigned char* __restrict__ non_bg) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= T) return;
    const float* p = logits + (long long)t * C;
    float mx = -INFINITY;
    int am = 0;
    for (int c = lane; c < C; c += 32) {
        const float v = p[c];
        if (v > mx) { mx = v; am = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
    }
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(p[c] - mx);
    s = warp_sum(s);
    if (probs)
        for (int c = lane; c < C; c += 32) probs[(long long)t * C + c] = expf(p[c] - mx) / s;
    if (lane == 0) {
        const float b = expf(p[0] - mx) / s;
        bg_prob[t] = b;
        seg_id[t] = am - 1;  // labels start from 0, class 0 is background (frame.py:121)
        non_bg[t] = b < bg_th;
    }
}

PRAM_API int pram_segmentation(const float* logits, int T, int C, float bg_threshold, float* probs, float* bg_prob,
                               int* seg_id, unsigned char* non_bg, cudaStream_t stream) {
    if (!logits || !bg_prob || !seg_id || !non_bg || T <= 0 || C <= 0) return PRAM_ERR_ARG;
    segmentation_kernel<<<cdiv((long long)T * 32, 256), 256, 0, stream>>>(logits, T, C, bg_threshold, probs, bg_prob, seg_id, non_bg);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// ------------------------------------------------------------------------------------------
// process_segmentations, two launches.
//  (1) top_classes_kernel: one warp per keypoint reads its C logits once (coalesced, register-resident) and
//      extracts its max_ranks best classes in (value descending, class index ascending) order ->
//      label_at_rank [max_ranks][N] (-1 for keypoints filtered out by the background mask).
//  (2) rank_entries_kernel: one CTA per frame.  rank k = 0,1,...: votes per class at this rank (count, sum of the
//      logits); classes that are not background and not yet used are appended in order of (vote count
//      descending, class id ascending) until topk entries exist.
// Outputs: entry_sid / entry_rank / entry_count / entry_score [topk], n_entries, and label_at_rank
// (the keypoint ids of entry e are { i : label_at_rank[entry_rank[e]][i] == entry_sid[e] }).
// ------------------------------------------------------------------------------------------
constexpr int RK_THREADS = 256;
constexpr int RK_MAXC = 1024;

template <int NV>  // classes per lane: C <= 32 * NV
__global__ void __launch_bounds__(256) top_classes_kernel(const float* __restrict__ logits,
                                                          const unsigned char* __restrict__ keep /*optional [B][N]*/, int N,
                                                          int C, int max_ranks, int* __restrict__ label_at_rank) {
    const int b = blockIdx.y;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= N) return;
    const float* p = logits + ((long long)b * N + i) * C;
    float v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = lane + 32 * j;
        v[j] = (c < C) ? p[c] : -INFINITY;
    }
    const bool kept = !keep || keep[(long long)b * N + i];
    unsigned taken = 0;  // bit j: class lane + 32 j already emitted
    for (int k = 0; k < max_ranks; ++k) {
        float bv = -INFINITY;
        int bc = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = lane + 32 * j;
            if (c < C && !((taken >> j) & 1u) && (bc == 0x7fffffff || v[j] > bv)) { bv = v[j]; bc = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            const bool better = (oc != 0x7fffffff) && (bc == 0x7fffffff || ov > bv || (ov == bv && oc < bc));
            if (better) { bv = ov; bc = oc; }
        }
        if (bc == 0x7fffffff) bc = -1;  // fewer than k+1 classes
        if (bc >= 0 && (bc & 31) == lane) taken |= 1u << (bc >> 5);
        if (lane == 0) label_at_rank[((long long)b * max_ranks + k) * N + i] = kept ? bc : -1;
    }
}

__global__ void __launch_bounds__(RK_THREADS) rank_entries_kernel(
    const float* __restrict__ logits, int N, int C, int topk, int max_ranks, int* __restrict__ entry_sid,
    int* __restrict__ entry_rank, int* __restrict__ entry_count, float* __restrict__ entry_score, int* __restrict__ n_entries,
    const int* __restrict__ label_at_rank) {
    __shared__ int cnt[RK_MAXC];
    __shared__ unsigned char used[RK_MAXC];
    __shared__ float wsum[RK_THREADS / 32];
    __shared__ int s_n, s_best, s_done, s_e;
    const int b = blockIdx.x;
    const float* L = logits + (long long)b * N * C;
    for (int c = threadIdx.x; c < C; c += RK_THREADS) used[c] = 0;
    if (threadIdx.x == 0) { s_n = 0; s_done = 0; }
    __syncthreads();
    for (int k = 0; k < max_ranks && k < C; ++k) {
        for (int c = threadIdx.x; c < C; c += RK_THREADS) cnt[c] = 0;
        __syncthreads();
        const int* lab = label_at_rank + ((long long)b * max_ranks + k) * N;
        for (int i = threadIdx.x; i < N; i += RK_THREADS) {
            const int c = lab[i];
            if (c >= 0) atomicAdd(&cnt[c], 1);  // integer counts: order independent
        }
        __syncthreads();
        // append unused, non-background classes of this rank by (count desc, class asc); warp 0 does the arg-max
        while (true) {
            if (threadIdx.x < 32) {
                int best = -1, bcnt = 0;
                for (int c = 1 + threadIdx.x; c < C; c += 32)
                    if (cnt[c] > 0 && !used[c] && cnt[c] > bcnt) { best = c; bcnt = cnt[c]; }  // ascending c: first wins ties
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int ob = __shfl_xor_sync(0xffffffffu, best, o), oc = __shfl_xor_sync(0xffffffffu, bcnt, o);
                    if (ob >= 0 && (best < 0 || oc > bcnt || (oc == bcnt && ob < best))) { best = ob; bcnt = oc; }
                }
                if (threadIdx.x == 0) {
                    s_best = best;
                    if (best >= 0) {
                        const int e = s_n;
                        entry_sid[(long long)b * topk + e] = best;
                        entry_rank[(long long)b * topk + e] = k;
                        entry_count[(long long)b * topk + e] = bcnt;
                        used[best] = 1;
                        s_e = e;
                        s_n = e + 1;
                        if (s_n >= topk) s_done = 1;
                    }
                }
            }
            __syncthreads();
            if (s_best >= 0) {
                // mean logit of the appended class over its keypoints, summed in a FIXED order (strided partials, shuffle
                // tree, warp partials in index order): float atomics made the score depend on the scheduling, so a graph
                // replay or a second stream could differ from the eager run in the last bit
                const int best = s_best;
                float part = 0.f;
                for (int i = threadIdx.x; i < N; i += RK_THREADS)
                    if (lab[i] == best) part += L[(long long)i * C + best];
                part = warp_sum(part);
                if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = part;
                __syncthreads();
                if (threadIdx.x == 0) {
                    float tot = 0.f;
                    for (int w = 0; w < RK_THREADS / 32; ++w) tot += wsum[w];
                    entry_score[(long long)b * topk + s_e] = tot / (float)cnt[best];
                }
            }
            const bool stop = (s_best < 0 || s_done);
            __syncthreads();  // everyone has read s_best before warp 0 overwrites it
            if (stop) break;
        }
        if (s_done) break;
    }
    if (threadIdx.x == 0) n_entries[b] = s_n;
}

PRAM_API int pram_rank_landmarks(const float* logits, const unsigned char* keep, int B, int N, int C, int topk,
                                 int max_ranks, int* entry_sid, int* entry_rank, int* entry_count, float* entry_score,
                                 int* n_entries, int* label_at_rank, cudaStream_t stream) {
    if (!logits || !entry_sid || !entry_rank || !entry_count || !entry_score || !n_entries || !label_at_rank) return PRAM_ERR_ARG;
    if (C > RK_MAXC || topk <= 0 || max_ranks <= 0) return PRAM_ERR_UNSUPPORTED;
    dim3 grid(cdiv((long long)N * 32, 256), B);
    if (C <= 128) top_classes_kernel<4><<<grid, 256, 0, stream>>>(logits, keep, N, C, max_ranks, label_at_rank);
    else if (C <= 256) top_classes_kernel<8><<<grid, 256, 0, stream>>>(logits, keep, N, C, max_ranks, label_at_rank);
    else if (C <= 512) top_classes_kernel<16><<<grid, 256, 0, stream>>>(logits, keep, N, C, max_ranks, label_at_rank);
    else top_classes_kernel<32><<<grid, 256, 0, stream>>>(logits, keep, N, C, max_ranks, label_at_rank);
    PRAM_CHECK_LAUNCH();
    rank_entries_kernel<<<B, RK_THREADS, 0, stream>>>(logits, N, C, topk, max_ranks, entry_sid, entry_rank, entry_count,
                                                     entry_score, n_entries, label_at_rank);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// ------------------------------------------------------------------------------------------
// K18: projection of map points (float64, like the reference's numpy/cuda double path)
//   uv = K (R X + t); valid = 0 < z < 100 and 0 <= u < width and 0 <= v < height
// ------------------------------------------------------------------------------------------
__global__ void project_points_kernel(const float* __restrict__ xyz, int n, const double* __restrict__ pose /*R[9],t[3]*/,
                                      double fx, double fy, double cx, double cy, double width, double height,
                                      float* __restrict__ uv, unsigned char* __restrict__ valid) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double X = xyz[3 * (long long)j], Y = xyz[3 * (long long)j + 1], Z = xyz[3 * (long long)j + 2];
    const double xc = pose[0] * X + pose[1] * Y + pose[2] * Z + pose[9];
    const double yc = pose[3] * X + pose[4] * Y + pose[5] * Z + pose[10];
    const double zc = pose[6] * X + pose[7] * Y + pose[8] * Z + pose[11];
    const double u = (fx * xc + cx * zc) / zc, v = (fy * yc + cy * zc) / zc;
    const bool ok = (zc > 0) && (zc < 100) && (u >= 0) && (u < width) && (v >= 0) && (v < height);
    uv[2 * (long long)j] = (float)u;
    uv[2 * (long long)j + 1] = (float)v;
    valid[j] = ok;
}

// one warp per query keypoint: top-2 of the masked descriptor distance over all valid map points
__global__ void __launch_bounds__(256) proj_top2_kernel(const float* __restrict__ sim /*[M][ld]*/, int ld, int M, int N,
                                                        const float* __restrict__ kpts, const float* __restrict__ uv,
                                                        const unsigned char* __restrict__ valid, float window,
                                                        float ratio, long long* __restrict__ match,
                                                        float* __restrict__ d0_out, float* __restrict__ d1_out) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= M) return;
    const float kx = kpts[2 * (long long)i], ky = kpts[2 * (long long)i + 1];
    float b0 = INFINITY, b1 = INFINITY;
    int j0 = -1;
    const float* s = sim + (long long)i * ld;
    for (int j = lane; j < N; j += 32) {
        if (!valid[j]) continue;
        const float dx = kx - uv[2 * (long long)j], dy = ky - uv[2 * (long long)j + 1];
        const float err = sqrtf(dx * dx + dy * dy);
        float d = sqrtf(2.f - 2.f * s[j] + 1e-6f);
        if (err >= window) d += 100.f;
        if (d < b0) { b1 = b0; b0 = d; j0 = j; }
        else if (d < b1) b1 = d;
    }
    // merge the per-lane top-2 lists (ties: lower index first)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob0 = __shfl_xor_sync(0xffffffffu, b0, o), ob1 = __shfl_xor_sync(0xffffffffu, b1, o);
        const int oj0 = __shfl_xor_sync(0xffffffffu, j0, o);
        if (ob0 < b0 || (ob0 == b0 && oj0 >= 0 && (j0 < 0 || oj0 < j0))) {
            b1 = fminf(b0, ob1); b0 = ob0; j0 = oj0;
        } else {
            b1 = fminf(b1, ob0);
        }
    }
    if (lane == 0) {
        const bool ok = (j0 >= 0) && (b0 / b1 <= ratio) && (b0 < 100.f);
        match[i] = ok ? (long long)j0 : -1ll;
        d0_out[i] = b0;
        d1_out[i] = b1;
    }
}

PRAM_API int pram_project_points(const float* xyz, int n, const double* pose, double fx, double fy, double cx, double cy,
                                 double width, double height, float* uv, unsigned char* valid, cudaStream_t stream) {
    if (!xyz || !pose || !uv || !valid || n <= 0) return PRAM_ERR_ARG;
    project_points_kernel<<<cdiv(n, 256), 256, 0, stream>>>(xyz, n, pose, fx, fy, cx, cy, width, height, uv, valid);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

PRAM_API int pram_projection_top2(const float* sim, int ld, int M, int N, const float* kpts, const float* uv,
                                  const unsigned char* valid, float window, float ratio, long long* match, float* d0,
                                  float* d1, cudaStream_t stream) {
    if (!sim || !kpts || !uv || !valid || !match || !d0 || !d1 || M <= 0 || N <= 0) return PRAM_ERR_ARG;
    proj_top2_kernel<<<cdiv((long long)M * 32, 256), 256, 0, stream>>>(sim, ld, M, N, kpts, uv, valid, window, ratio, match, d0, d1);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// ------------------------------------------------------------------------------------------
// NearestNeighbor matcher (reference localization/matchers/nearest_neighbor.py:5-56; section 8f row 4):
//   find_nn: top-2 of each similarity row, dist = 2 (1 - sim), optional ratio test d0 <= ratio^2 d1 and distance
//   test d0 <= dist_th^2; matches = arg-max or -1, scores = (sim0 + 1) / 2 or 0.  One warp per row; the M x N
//   similarity comes from the tcgen05 GEMM.  mutual_check: keep i -> j only if j -> i.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nn_top2_kernel(const float* __restrict__ sim, long long batch_stride, int ld, int rows,
                                                      int cols, int B, float ratio2 /*< 0: off*/, float dist2 /*< 0: off*/,
                                                      long long* __restrict__ matches, float* __restrict__ scores) {
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= (long long)B * rows) return;
    const int b = (int)(wid / rows), i = (int)(wid - (long long)b * rows);
    const float* p = sim + (long long)b * batch_stride + (long long)i * ld;
    float v0 = -INFINITY, v1 = -INFINITY;
    int i0 = 0x7fffffff;
    for (int j = lane; j < cols; j += 32) {
        const float v = p[j];
        if (v > v0) { v1 = v0; v0 = v; i0 = j; }   // ascending j: the first occurrence wins ties
        else if (v > v1) v1 = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov0 = __shfl_xor_sync(0xffffffffu, v0, o), ov1 = __shfl_xor_sync(0xffffffffu, v1, o);
        const int oi0 = __shfl_xor_sync(0xffffffffu, i0, o);
        const bool other_wins = (ov0 > v0) || (ov0 == v0 && oi0 < i0);
        const float loser = other_wins ? v0 : ov0;
        v1 = fmaxf(fmaxf(v1, ov1), loser);
        if (other_wins) { v0 = ov0; i0 = oi0; }
    }
    if (lane == 0) {
        const float d0 = 2.f * (1.f - v0), d1 = 2.f * (1.f - v1);
        bool ok = true;
        if (ratio2 >= 0.f) ok = ok && (d0 <= ratio2 * d1);
        if (dist2 >= 0.f) ok = ok && (d0 <= dist2);
        matches[wid] = ok ? (long long)i0 : -1ll;
        scores[wid] = ok ? (v0 + 1.f) * 0.5f : 0.f;
    }
}

__global__ void nn_mutual_kernel(long long* __restrict__ m0, const long long* __restrict__ m1, int N, int M, int B) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * N) return;
    const int b = (int)(t / N), i = (int)(t - (long long)b * N);
    const long long j = m0[t];
    if (j > -1 && m1[(long long)b * M + j] != i) m0[t] = -1;
}

// sim [B][N][M] (row stride ld = M), simT [B][M][N] (may be NULL when mutual == 0)
PRAM_API int pram_nn_match(const float* sim, const float* simT, int B, int N, int M, float ratio_threshold, float distance_threshold,
                           int mutual, long long* matches0, float* scores0, long long* matches1_ws, float* scores1_ws,
                           cudaStream_t stream) {
    if (!sim || !matches0 || !scores0 || B <= 0 || N <= 0 || M <= 0) return PRAM_ERR_ARG;
    if (mutual && (!simT || !matches1_ws || !scores1_ws)) return PRAM_ERR_ARG;
    const float r2 = ratio_threshold > 0.f ? ratio_threshold * ratio_threshold : -1.f;
    const float d2 = distance_threshold > 0.f ? distance_threshold * distance_threshold : -1.f;
    nn_top2_kernel<<<cdiv((long long)B * N * 32, 256), 256, 0, stream>>>(sim, (long long)N * M, M, N, M, B, r2, d2, matches0, scores0);
    PRAM_CHECK_LAUNCH();
    if (mutual) {
        nn_top2_kernel<<<cdiv((long long)B * M * 32, 256), 256, 0, stream>>>(simT, (long long)N * M, N, M, N, B, r2, d2, matches1_ws, scores1_ws);
        PRAM_CHECK_LAUNCH();
        nn_mutual_kernel<<<cdiv((long long)B * N, 256), 256, 0, stream>>>(matches0, matches1_ws, N, M, B);
        PRAM_CHECK_LAUNCH();
    }
    return PRAM_OK;
}
