/* libpram_b200 -- C ABI of the B200-native (sm_100a) PRAM localization hot path.
 *
 * The reference (feixue94/pram) is 100 % Python/PyTorch and has no FFI of its own; its operator
 * boundary is the Python API in nets/{sfd2,segnetvit,gml,adagml}.py and localization/matchers/*.py
 * (SURVEY.md section 8b).  pram_b200 keeps that Python API and implements it on top of the entry
 * points below, which a reference maintainer could equally bind with ctypes from the original modules
 * (INTEGRATION.md shows the stubs).  Every entry point cites the reference code it replaces.
 *
 * Conventions: plain device pointers and sizes, no allocation and no host synchronisation inside,
 * work is enqueued on the given cudaStream_t (passed as void*), return 0 on success or a negative
 * PRAM_ERR_* code.  All floating-point buffers are fp32 unless stated; feature maps are NHWC.
 */
#ifndef PRAM_B200_H
#define PRAM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PRAM_OK 0
#define PRAM_ERR_ARG (-1)
#define PRAM_ERR_CUDA (-2)
#define PRAM_ERR_WORKSPACE (-3)
#define PRAM_ERR_UNSUPPORTED (-4)

typedef void* pram_stream_t; /* cudaStream_t */

int pram_version(void);
unsigned long long pram_launch_count(void); /* kernels launched by this library so far */
const char* pram_error_string(int code);
const char* pram_last_cuda_error(void);

/* Device-side early exit (AdaGML's stop test, nets/adagml.py:370-372 `break`): while `flag` (device int) is registered, the
 * GEMM / attention / block-tail / Linear / LayerNorm / AdaGML-control kernels launched by this thread return immediately when
 * *flag == 0 at execution time.  NULL clears it.  Host-side state only (no launch, no synchronisation). */
int pram_set_launch_predicate(const int* flag);

/* Programmatic dependent launch of the persistent tensor-core kernels (GEMM, attention, block tail, conv1a, grouped conv):
 * each is scheduled while its predecessor in the stream drains and waits on the device (griddepcontrol.wait) before touching
 * global memory.  On by default (environment PRAM_PDL=0 or pram_set_pdl(0) turns it off). */
int pram_set_pdl(int enable);
int pram_get_pdl(void);

/* K5: softmax(65) -> drop dustbin -> 8x8 pixel shuffle.  nets/sfd2.py:294-300.
 * logits addressed as base + b*batch_stride + hc*y_stride + wc*x_stride + c*ch_stride (floats). */
int pram_score_map(const float* logits, long long batch_stride, long long y_stride, long long x_stride,
                   long long ch_stride, int B, int Hc, int Wc, float* score, pram_stream_t stream);

/* K5b: bilinear resize, align_corners=True.  nets/sfd2.py:301-303. */
int pram_resize_bilinear(const float* in, int B, int Hi, int Wi, float* out, int Ho, int Wo,
                         pram_stream_t stream);

/* K6: simple_nms(radius) + candidate emission.  nets/sfd2.py:20-35, 305-315.
 * cand[B][cap] receives 64-bit keys (score bits << 32 | y*W+x) of survivors >= th_lo;
 * cand_count[B] their number (may exceed cap: overflow), count_hi[B] the number >= th_hi.
 * nms_out (optional) receives the dense NMS map. */
int pram_nms_candidates(const float* score, int B, int H, int W, int radius, float th_lo, float th_hi,
                        float* nms_out, unsigned long long* cand, int cap, int* cand_count, int* count_hi,
                        pram_stream_t stream);

/* K7: threshold fallback, border removal, top-k / row-major ordering, (x,y) float output.
 * nets/sfd2.py:38-50, 306-329.  kpts[B][kpad][2], scores[B][kpad], n_out[B].  Border window: border <= y < y_hi,
 * border <= x < x_hi (y_hi / x_hi <= 0 = H - border / W - border; the export path passes the ORIGINAL image's bounds for
 * a rescaled map, nets/sfd2.py:447-451).  max_keypoints < 0 = unlimited; with more valid keypoints than kpad slots the
 * best kpad by score are returned and n_valid_out[B] (optional) reports the true count. */
int pram_select_keypoints(const unsigned long long* cand, int cap, const int* cand_count, const int* count_hi,
                          const float* score, int B, int H, int W, float th_lo, float th_hi, int min_keypoints,
                          int max_keypoints, int border, int y_hi, int x_hi, float* kpts, float* scores, int* n_out,
                          int kpad, int* n_valid_out, pram_stream_t stream);

/* K8: bilinear sampling of an NHWC map at keypoints (+ optional L2 norm).  nets/sfd2.py:53-64, 348-363.
 * fmap[B][h][w][C], kpts[B][kpad][2], counts[B] or NULL, out[B][kpad][C]. */
int pram_sample_features(const float* fmap, int B, int C, int h, int w, const float* kpts, const int* counts,
                         int kpad, int s, int normalize, float* out, pram_stream_t stream);

/* score_map[b][(int)y][(int)x] gather.  nets/sfd2.py:367. */
int pram_gather_scores(const float* score, int B, int H, int W, const float* kpts, const int* counts, int kpad,
                       float* out, pram_stream_t stream);

/* K9: normalize_keypoints + learnable Fourier encoding -> cos/sin [tokens][32].
 * nets/utils.py:17-24, nets/segnetvit.py:35-40 (== nets/gml.py:69-74). */
int pram_posenc(const float* kpts, int tokens, float width, float height, int prenormalized, const float* Wr,
                float* cos_out, float* sin_out, pram_stream_t stream);

/* K1-K4 (fp32 CUDA-core path): 3x3 / 1x1 convolution, NHWC, BN pre-folded, fused bias/residual/ReLU.
 * nets/sfd2.py:141-170.  w[taps][Cin][Cout]. */
int pram_conv_f32(const float* in, long long in_pix_stride, const float* w, const float* bias, const float* res,
                  long long res_pix_stride, float* out, long long out_pix_stride, int B, int H, int W, int Cin,
                  int Cout, int ksize, int stride, int relu, pram_stream_t stream);

/* nn.Linear / batched A.B^T with row strides (concat-free MLPs).  nets/segnetvit.py:88-95,
 * nets/gml.py:119-126, 152-159, 278-282. */
int pram_linear_f32(const float* a, long long lda, const float* w, const float* bias, const float* res,
                    long long ldres, float* out, long long ldo, long long rows, int K, int N, int relu, int batch,
                    long long a_batch_stride, long long w_batch_stride, long long out_batch_stride,
                    pram_stream_t stream);

/* K2: grouped 3x3 convolution (32 groups x 8 channels) + bias + ReLU.  nets/sfd2.py:101. */
int pram_gconv3x3_f32(const float* in, const float* w, const float* bias, float* out, int B, int H, int W,
                      int groups, int relu, pram_stream_t stream);

int pram_gconv3x3_split(const float* in, const float* w, const float* bias, float* out_f32, void* out_hi,
                        void* out_lo, int B, int H, int W, int groups, int relu, pram_stream_t stream);

/* same convolution on tensor cores (warp-level bf16 MMA, m16n8k16 = 16 pixels x one 8-channel group x two taps),
 * split-bf16 NHWC planes in and out; replaces the cuDNN grouped convolution of nets/sfd2.py:100-124 (ResBlock.conv2,
 * groups = 32).  w fp32 [9][8 ci][8 co][C/8], C % 64 == 0, lo planes NULL when split == 1. */
int pram_gconv3x3_tc(const void* in_hi, const void* in_lo, const float* w, const float* bias, void* out_hi, void* out_lo,
                     int B, int H, int W, int C, int relu, int split, pram_stream_t stream);

/* K1 first layer: conv1a 3->64 + BN + ReLU straight from the NCHW image, output as split-bf16 NHWC in
 * the 2x2 phase-split layout consumed by the stride-2 conv1b (and/or fp32 NHWC).  nets/sfd2.py:141.
 * w[27][64] (tap-major (r,s,c)), bias[64]. */
int pram_conv1a(const float* img_nchw, const float* w, const float* bias, int B, int H, int W, void* ps_hi,
                void* ps_lo, float* out_f32, pram_stream_t stream);

/* conv1a on tcgen05: the CTA builds the 128 x 32 im2col operand (bf16 hi/lo, SWIZZLE_128B K-major) in shared memory
 * itself, 2 k-steps x 3 MMAs per 128 pixels, phase-split split-bf16 output as above.  nets/sfd2.py:141. */
int pram_conv1a_tc(const float* img_nchw, const float* w, const float* bias, int B, int H, int W, void* ps_hi, void* ps_lo,
                   int split, pram_stream_t stream);

/* F.normalize over the channel axis of an NHWC map.  nets/sfd2.py:333. */
int pram_l2norm_rows(const float* in, float* out, long long rows, int C, pram_stream_t stream);

/* LayerNorm + exact GELU.  nets/segnetvit.py:92-93. */
int pram_layernorm_gelu(const float* in, const float* gamma, const float* beta, float* out, long long rows, int C,
                        int gelu, pram_stream_t stream);

int pram_layernorm_gelu_split(const float* in, const float* gamma, const float* beta, float* out_f32, void* out_hi,
                              void* out_lo, long long rows, int C, int gelu, pram_stream_t stream);

/* qkv split + rotary embedding -> q,k,v [B][heads][N][64].  nets/segnetvit.py:15-23, 98-103. */
int pram_rotary_split(const float* qkv, int nparts, int B, int N, int heads, const float* cosb, const float* sinb,
                      float scale_qk, float* q, float* k, float* v, pram_stream_t stream);

/* K10/K12/K13: softmax(QK^T*scale)V without materialising the N x N matrix; optional per-key mean
 * attention (AdaGML).  nets/segnetvit.py:73-76, nets/gml.py:175-181, nets/adagml.py:148. */
int pram_attention_f32(const float* Q, const float* K, const float* V, int B, int heads, int Nq, int Nk,
                       float scale, float* out, int out_stride, float* colmean,
                       float* colmean_ws /* pram_attention_f32_colmean_ws_floats() floats when colmean != NULL: per-CTA partials,
                                            reduced in a fixed order (bit-reproducible pruning decisions in AdaGML) */,
                       pram_stream_t stream);
long long pram_attention_f32_colmean_ws_floats(int B, int heads, int Nq, int Nk);

/* K15+K16: dustbin Sinkhorn (probability domain) + mutual arg-max matches in one launch.
 * nets/gml.py:27-46, 304-319.  pws: pram_sinkhorn_workspace_floats() floats (holds P on return),
 * iws: B*(M+N) ints, fws: B*M floats. */
long long pram_sinkhorn_workspace_floats(int B, int M, int N);
int pram_sinkhorn_match(const float* dist, int B, int M, int N, const float* bin_score, int iters,
                        float threshold, float* pws, int* iws, float* fws, long long* matches0,
                        long long* matches1, float* mscores0, float* mscores1, int cluster,
                        const int* m_counts, const int* n_counts /* optional [B]: pair b is the (m_counts[b]+1) x (n_counts[b]+1)
                        problem of its first rows / columns; the rest of the [M, N] block is padding (matches -1, scores 0) */,
                        pram_stream_t stream);

/* K1-K4, K10-K14 (tensor-core path): tcgen05 implicit GEMM, TMA-fed, accumulators in TMEM.
 *   D[pixel][n] = sum_{tap,c} A[pixel + offset(tap)][c] * W[tap][n][c]  (+bias)(+res)(ReLU)(L2 norm)
 * A: bf16 NHWC activations (optionally hi/lo split planes; optionally 2x2 phase-split for stride 2),
 * W: bf16 [planes][N][Cin].  Replaces cuDNN/cuBLAS under nets/sfd2.py:141-170 and the nn.Linear /
 * einsum calls of nets/segnetvit.py:88-106, nets/gml.py:119-186, 278-282. */
typedef struct pram_tc_args {
    const void* a_hi; const void* a_lo; long long a_ld;
    int in_W, in_H, in_planes, Cin;
    const void* w_hi; const void* w_lo; int w_planes;
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
    int split; /* 1: bf16, 3: error-compensated bf16x3 */
    int bn;    /* N tile (64/128/256), 0 = auto */
    /* fused attention-operand epilogue: qkv_mode 1 = output columns (q|k|v), rotary on q,k (nets/segnetvit.py:98-103);
     * 2 = (qk|v) of the cross block (nets/gml.py:165-174).  Results go to split-bf16 tensors [B][heads][n][64]
     * (tokens < seg_split form segment 0 with n = seg_n0 per batch element, the rest segment 1 with seg_n1). */
    int qkv_mode; const float* cosb; const float* sinb; float qk_scale;
    void* q_hi; void* q_lo; void* k_hi; void* k_lo; void* v_hi; void* v_lo;
    int seg_split, seg_n0, seg_n1, heads;
    int cluster;                          /* 0 = auto, 1 = single CTAs, 2 = 2-CTA clusters, weight tile TMA-multicast */
    int l2_prefetch;                      /* 1 = L2-prefetch the next tile's activation boxes (single-tap layers); default off */
    int f16;                              /* 1 (split == 1): a / w planes hold IEEE fp16 and out_hi receives fp16 (single-pass fp16 mode) */
    const void* res_hi; const void* res_lo; /* residual as split-bf16 planes (when res == NULL): r = hi + lo, row stride res_ld */
    int v_f16;                            /* qkv epilogue: v_hi / v_lo receive IEEE fp16 planes (V operand of pram_attention_tc with v_f16) */
    void* out_h16;                        /* optional extra copy of the output as ONE IEEE fp16 plane (row stride ld_bf, N % 32 == 0): operand of a following f16 layer */
} pram_tc_args;
int pram_gemm_tc(const pram_tc_args* args, pram_stream_t stream);

/* K11/K13 block tail (tensor-core path), ONE persistent kernel per transformer block:
 *   message = proj(ctx);  x_new = x + mlp.3(GELU(LayerNorm(mlp.0([x | message]))))
 * (SelfMultiHeadAttention / CrossMultiHeadAttention tails: nets/segnetvit.py:104-106, nets/gml.py:135-137, 182-186).
 * proj is folded into mlp.0 by the caller (w1 = [W0x | W0m.Wp], b1 = b0 + W0m.bp); the 128-token tile stays in
 * TMEM / shared memory from the first MMA to the residual add (h = all 512 TMEM columns, LayerNorm + GELU in the
 * epilogue warps, hidden activations fed to the second GEMM through a shared-memory operand ring). */
typedef struct pram_mlp_block_args {
    const void* a_hi; const void* a_lo;   /* bf16 [T][lda]: columns 0..255 = x, 256..511 = attention context */
    long long lda;
    int T;
    const void* w1_hi; const void* w1_lo; /* bf16 [8][512][64]: tile kb = W1[:, 64 kb : 64 kb + 64] (every TMA box contiguous) */
    const void* w3_hi; const void* w3_lo; /* bf16 [16][256][32]: tile s = W3[:, 32 s : 32 s + 32] */
    const float* tables_host;             /* HOST fp32 [b1 512 | ln gamma 512 | ln beta 512 | b3 256] -> kernel parameter block */
    const float* res; long long res_ld;   /* fp32 residual rows, or NULL: residual = hi + lo of columns 0..255 of a */
    float* out_f32; long long ld_f32;     /* may be NULL */
    void* out_hi; void* out_lo; long long ld_bf;  /* may be NULL */
    int split;                            /* 1: bf16, 3: bf16x3 */
    long long* dbg;                       /* optional device buffer [grid][8][32] of SM clock stamps (profiling aid), NULL = off */
} pram_mlp_block_args;
int pram_mlp_block_tc(const pram_mlp_block_args* args, pram_stream_t stream);

/* K10/K12/K13 (tensor-core path): flash attention on tcgen05, head dim 64.  S = QK^T and O += PV on the
 * tensor cores (P is fed back from TMEM as the A operand), softmax on one thread per query row.
 * q/k: bf16 [B*heads][N][64]; vt: bf16 [B*heads][64][nk_pad] (keys contiguous) when v_mn == 0, or V itself
 * [B*heads][Nk][64] when v_mn & 1 (MN-major UMMA operand); *_lo NULL when split == 1.  v_mn & 2: the V planes hold IEEE
 * fp16 hi / lo (pram_split_f16, pram_tc_args.v_f16) and the probabilities are fed back as ONE fp16 plane (11-bit mantissas,
 * row sums taken from the rounded values): two PV MMAs per k-step instead of three and half the softmax instructions.
 * Replaces Attention.forward, nets/segnetvit.py:73-76 and the two einsum+softmax pairs of
 * nets/gml.py:175-181. */
int pram_attention_tc(const void* q_hi, const void* q_lo, const void* k_hi, const void* k_lo, const void* vt_hi,
                      const void* vt_lo, int B, int heads, int Nq, int Nk, int nk_pad, float scale, float* out_f32,
                      void* out_hi, void* out_lo, int out_ld, int split, int kv_tile /* 0 = auto, 64 (two CTAs per SM) or 128 */, int v_mn,
                      const int* nk_counts /* optional [B]: keys >= nk_counts[b] of batch element b are padding and masked */,
                      pram_stream_t stream);

/* Same launch with rotated key / value batches: query batch element b attends to the keys / values (and nk_counts entry) of
 * batch element (b + kv_shift) mod B.  With B = 2 x pairs, kv_shift = pairs and q = k = [set 0 | set 1] this is both directions
 * of the bidirectional cross attention (nets/gml.py:175-181) in one launch. */
int pram_attention_tc_shift(const void* q_hi, const void* q_lo, const void* k_hi, const void* k_lo, const void* vt_hi,
                            const void* vt_lo, int B, int heads, int Nq, int Nk, int nk_pad, float scale, float* out_f32,
                            void* out_hi, void* out_lo, int out_ld, int split, int kv_tile, int v_mn, const int* nk_counts,
                            int kv_shift, pram_stream_t stream);

/* Same launch, additionally writing the log2-domain log-sum-exp of every query row (softmax prob = exp2(s * scale * log2 e -
 * lse)) to lse_out [B*heads][ld_lse] (ld_lse % 4 == 0, >= Nq). */
int pram_attention_tc_lse(const void* q_hi, const void* q_lo, const void* k_hi, const void* k_lo, const void* vt_hi,
                          const void* vt_lo, int B, int heads, int Nq, int Nk, int nk_pad, float scale, float* out_f32,
                          void* out_hi, void* out_lo, int out_ld, int split, int kv_tile, int v_mn, const int* nk_counts,
                          float* lse_out, int ld_lse, pram_stream_t stream);

/* K17, Attention.forward's second output (nets/adagml.py:145-148; cross block :229): the attention mass every KEY receives,
 * colsum[b*heads + h][key] = sum over the valid queries of softmax(query, key), on tcgen05: S^T = K Q^T tiles (one thread per
 * key row) against the queries' row statistics `lse` of the pram_attention_tc_lse launch (ld_lse >= Nqueries rounded up to
 * 128).  nq_counts optional [B]: queries >= nq_counts[b] are padding.  Mean attention = pram_colmean_reduce. */
int pram_attention_colsum_tc(const void* key_hi, const void* key_lo, const void* qry_hi, const void* qry_lo, int B, int heads,
                             int Nkeys, int Nqueries, float scale, const float* lse, int ld_lse, float* colsum, int ld_colsum,
                             int split, int kv_tile, const int* nq_counts, pram_stream_t stream);
/* out[(b*N + j) * out_stride] = sum_h colsum[(b*heads + h) * ld + j] / (heads * queries_b), fixed order (bit-reproducible). */
int pram_colmean_reduce(const float* colsum, int ld, int B, int heads, int N, int nq, const int* nq_counts, float* out,
                        int out_stride, pram_stream_t stream);

/* K17, AdaGML's pruning / early-exit control flow on the device (nets/adagml.py:344-372, 516-531) for a BATCH of pairs in the
 * fixed token layout rows [0, B*M) = set 0, [B*M, B*(M+N)) = set 1, with per-pair token counts cnt0 / cnt1 [B]:
 *  prune : keep = sigmoid(conf_logits) > threshold for sets with >= n_min_tokens tokens (do_prune), stable compaction
 *          positions dest[row] (-1 = dropped), counts updated in place, stop test `1 - #(conf < th) / (m + n) > 0.95` ->
 *          stop_layer[b] = layer (also at the last layer), active[l] (pairs still running at the START of layer l, [n_layers],
 *          initialised to B: the launch predicate of layer l's kernels) decremented for l > layer, trace[layer][set][b] = counts,
 *          *err = 1 when a set is pruned to zero tokens (the reference raises);
 *  move  : kept rows -> their compacted position in the other buffer set (bf16 planes, fp32 rows, rotary cos / sin, index map);
 *  latch : pairs with stop_layer[b] == layer keep out_proj[layer](tokens) (src planes [T][256]), index map and counts;
 *  scatter: matches of the compacted problems -> full-size matches0 / matching_scores0 (nets/adagml.py:383-394). */
int pram_adagml_prune(const float* conf_logits, int B, int M, int N, int* cnt0, int* cnt1, float threshold, int n_min_tokens,
                      int do_prune, int layer, int n_layers, int* dest, int* stop_layer, int* active, int* trace, int* err,
                      const int* full0, const int* full1 /* optional [B]: the pairs' original token counts (padded batches) */,
                      pram_stream_t stream);
int pram_adagml_move(const int* dest, int B, int M, int N, const void* src_hi, const void* src_lo, void* dst_hi, void* dst_lo,
                     long long ld_bf, const float* src_f32, float* dst_f32, long long ld_f32, const float* src_cos,
                     const float* src_sin, float* dst_cos, float* dst_sin, const int* src_ind, int* dst_ind,
                     pram_stream_t stream);
int pram_adagml_latch(const int* stop_layer, int layer, int B, int M, int N, const void* src_hi, const void* src_lo, void* dst_hi,
                      void* dst_lo, const int* src_ind, int* dst_ind, const int* cnt0, const int* cnt1, int* final_cnt0,
                      int* final_cnt1, pram_stream_t stream);
int pram_adagml_scatter(const long long* matches0, const float* mscores0, const int* ind, const int* final_cnt0,
                        const int* final_cnt1, int B, int M, int N, long long* full_matches0, float* full_scores0,
                        pram_stream_t stream);

/* qkv fp32 rows -> the attention kernel's operands (rotary + scale on q,k; V transposed per head).
 * nets/segnetvit.py:98-103, nets/gml.py:169-174. */
int pram_attention_prep(const float* qkv, int nparts, int B, int N, int heads, const float* cosb, const float* sinb,
                        float scale_qk, void* q_hi, void* q_lo, void* k_hi, void* k_lo, void* vt_hi, void* vt_lo,
                        int n_pad, pram_stream_t stream);

/* K19: batched absolute pose, P3P + RANSAC + local optimisation + robust refinement, float64.
 * Replaces pycolmap.absolute_pose_estimation at localization/singlemap3d.py:168-175 (and :324, :454,
 * tracker.py:211, pose_estimator.py:213,338,452).  kpts [B][n][2] f32 pixels, matches [B][n] i64 into
 * xyz [B][nref][3] f32 (-1 = unmatched).  Outputs qvec (wxyz) / tvec f64, inlier mask in keypoint order. */
long long pram_ransac_workspace_bytes(int B, int cap, int num_hypotheses);
int pram_ransac_pnp(const float* kpts, const long long* matches, const float* xyz, int B, int n, int nref, double fx,
                    double fy, double cx, double cy, double pixel_shift, double max_error, int num_hypotheses,
                    int lo_iters, int final_iters, int min_inliers, unsigned int seed, void* workspace, double* qvec,
                    double* tvec, int* num_inliers, unsigned char* inliers, int* success, pram_stream_t stream);

/* Same estimator on float64 correspondences normalised by the caller: corr [B][n][5] = (x, y, X, Y, Z), (x, y) = undistorted
 * camera-plane coordinates.  Used by the pycolmap-compatible host call (float64 numpy inputs, COLMAP camera models with
 * distortion: the call sites above pass SIMPLE_RADIAL cameras on Aachen).  counts [B] or NULL; focal_mean converts max_error
 * (pixels) into the camera plane like COLMAP's CamFromImgThreshold.  Workspace: pram_ransac_workspace_bytes(B, n, hyp). */
int pram_ransac_pnp_corr(const double* corr, const int* counts, int B, int n, double focal_mean, double max_error,
                         int num_hypotheses, int lo_iters, int final_iters, int min_inliers, unsigned int seed, void* workspace,
                         double* qvec, double* tvec, int* num_inliers, unsigned char* inliers, int* success,
                         pram_stream_t stream);

/* ---- "next" rows ---- */
/* Frame.add_segmentations (localization/frame.py:96-121): softmax, background probability, argmax-1, pre-filter mask. */
int pram_segmentation(const float* logits, int T, int C, float bg_threshold, float* probs, float* bg_prob, int* seg_id,
                      unsigned char* non_bg, pram_stream_t stream);
/* MultiMap3D.process_segmentations (localization/multimap3d.py:348-379): greedy landmark ranking, one CTA per frame. */
int pram_rank_landmarks(const float* logits, const unsigned char* keep, int B, int N, int C, int topk, int max_ranks,
                        int* entry_sid, int* entry_rank, int* entry_count, float* entry_score, int* n_entries,
                        int* label_at_rank, pram_stream_t stream);
/* K18, SingleMap3D.refine_pose_by_projection (localization/singlemap3d.py:405-440): projection + masked top-2. */
int pram_project_points(const float* xyz, int n, const double* pose, double fx, double fy, double cx, double cy,
                        double width, double height, float* uv, unsigned char* valid, pram_stream_t stream);
int pram_projection_top2(const float* sim, int ld, int M, int N, const float* kpts, const float* uv,
                         const unsigned char* valid, float window, float ratio, long long* match, float* d0, float* d1,
                         pram_stream_t stream);

/* NearestNeighbor matcher (localization/matchers/nearest_neighbor.py:5-56): top-2 per similarity row, ratio / distance
 * tests (thresholds <= 0 disable them), optional mutual check.  sim [B][N][M], simT [B][M][N] (NULL when mutual == 0);
 * matches0 int64 [B][N] (-1 = none), scores0 = (sim+1)/2 or 0; *_ws: [B][M] scratch for the reverse direction. */
int pram_nn_match(const float* sim, const float* simT, int B, int N, int M, float ratio_threshold, float distance_threshold,
                  int mutual, long long* matches0, float* scores0, long long* matches1_ws, float* scores1_ws,
                  pram_stream_t stream);

/* fp32 -> one plane of IEEE fp16 (operand of the single-pass fp16 GEMM mode, pram_tc_args.f16). */
int pram_cast_f16(const float* in, void* out, long long n, pram_stream_t stream);

/* fp32 -> IEEE fp16 hi / lo planes (lo may be NULL). */
int pram_split_f16(const float* in, void* hi, void* lo, long long n, pram_stream_t stream);

/* fp32 -> split bf16 planes: hi = bf16(x), lo = bf16(x - hi) (lo may be NULL). */
int pram_split_bf16(const float* in, void* hi, void* lo, long long n, pram_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PRAM_B200_H */
