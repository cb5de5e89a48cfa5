// AdaGML's data-dependent control flow (K17, reference nets/adagml.py:344-372, 516-531) evaluated ON THE DEVICE.
//
// The reference prunes tokens with boolean-mask indexing and decides the early exit with a host read per layer
// (`pos > 0.95`), which pins it to batch 1 and to one device->host synchronisation per layer.  Here a batch of pairs
// keeps its fixed [B, M | N] token layout; every pair carries device-side token counts, pruning COMPACTS the surviving
// rows of each set to the front of its segment (stable order, so results equal the reference's masked tensors), and the
// kernels downstream treat rows >= count as non-existent (attention key masking, per-pair Sinkhorn sizes).  The early
// exit becomes a per-pair `stop_layer`; the state a pair exits with (projected descriptors, index maps, counts) is
// LATCHED at that layer, so whatever the later layers compute for it is irrelevant, and a per-layer `active[l]` counter
// (pairs still running at the start of layer l) lets every launch of a layer be skipped once all pairs have stopped.
//
// token row layout (as in the batched GML path): rows [0, B*M) = set 0 of pair 0..B-1, rows [B*M, B*(M+N)) = set 1.
#include "common.cuh"

namespace adagml {

__device__ __forceinline__ void row_to_pair(int row, int B, int M, int N, int& set, int& b, int& r, int& base) {
    if (row < B * M) { set = 0; b = row / M; r = row - b * M; base = b * M; }
    else { set = 1; const int q = row - B * M; b = q / N; r = q - b * N; base = B * M + b * N; }
}

// att[(b*N + j) * out_stride] = (sum_h colsum[(b*heads + h) * ld + j]) / (heads * queries_b): mean over heads and queries of
// the attention a key receives (nets/adagml.py:148: torch.mean(torch.mean(attn, dim=1), dim=1)), fixed summation order.
__global__ void colmean_reduce_kernel(const float* __restrict__ colsum, int ld, int B, int heads, int N, int nq,
                                      const int* __restrict__ nq_counts, float* __restrict__ out, int out_stride,
                                      const int* __restrict__ pred) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N || pram_pred_skip(pred)) return;
    const int b = (int)(i / N), j = (int)(i - (long long)b * N);
    float s = 0.f;
    for (int h = 0; h < heads; ++h) s += colsum[(long long)(b * heads + h) * ld + j];
    const int q = nq_counts ? max(1, min(nq, nq_counts[b])) : nq;
    out[i * out_stride] = s / (float)(heads * q);
}

// One CTA per pair.  For both sets: keep flag of every valid row (confidence > threshold when the set still has at least
// n_min tokens, else everything: nets/adagml.py:345-355), stable exclusive scan -> dest[row] (position inside the set's
// segment, -1 = dropped / padding), new counts; the stop statistic of nets/adagml.py:522-531 over the confidences of the
// CURRENT tokens of both sets with the FULL sizes as the denominator.
__global__ void __launch_bounds__(1024) prune_kernel(const float* __restrict__ z, int B, int M, int N, int* __restrict__ cnt0,
                                                     int* __restrict__ cnt1, float th, int n_min, int do_prune, int layer,
                                                     int is_last, int n_layers, int* __restrict__ dest,
                                                     int* __restrict__ stop_layer, int* __restrict__ active,
                                                     int* __restrict__ trace, int* __restrict__ err,
                                                     const int* __restrict__ full0, const int* __restrict__ full1,
                                                     const int* __restrict__ pred) {
    __shared__ int warp_tot[32];
    __shared__ int running, below_tot;
    if (pram_pred_skip(pred)) return;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool stopped = stop_layer[b] >= 0;
    int newc[2];
    if (tid == 0) below_tot = 0;
    for (int set = 0; set < 2; ++set) {
        const int cap = set ? N : M;
        const int base = set ? B * M + b * N : b * M;
        const int len = min(cap, set ? cnt1[b] : cnt0[b]);
        const bool prune = do_prune && !stopped && len >= n_min;
        if (tid == 0) running = 0;
        __syncthreads();
        int below = 0;
        for (int c0 = 0; c0 < cap; c0 += 1024) {
            const int r = c0 + tid;
            const bool valid = r < len;
            bool keep = valid;
            if (valid && do_prune && !stopped) {
                const float conf = 1.f / (1.f + expf(-z[base + r]));  // torch.sigmoid
                below += conf < th;
                if (prune) keep = conf > th;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int pre = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) warp_tot[wid] = __popc(bal);
            __syncthreads();
            int off = running;
            for (int w = 0; w < wid; ++w) off += warp_tot[w];
            if (r < cap) dest[base + r] = keep ? off + pre : -1;
            __syncthreads();
            if (tid == 0) {
                int t = 0;
                for (int w = 0; w < 32; ++w) t += warp_tot[w];
                running += t;
            }
            __syncthreads();
        }
        for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        if (lane == 0 && below) atomicAdd(&below_tot, below);
        __syncthreads();
        newc[set] = running;
        __syncthreads();
    }
    if (tid == 0) {
        cnt0[b] = newc[0];
        cnt1[b] = newc[1];
        if (trace) { trace[(layer * 2 + 0) * B + b] = newc[0]; trace[(layer * 2 + 1) * B + b] = newc[1]; }
        if (newc[0] == 0 || newc[1] == 0) *err = 1;  // the reference raises when a set is pruned to nothing
        if (!stopped) {
            bool stop = is_last != 0;
            if (do_prune) {
                // num_points = the pair's ORIGINAL m + n (nets/adagml.py:370), not the padded capacity of the batch layout
                const int num_points = (full0 ? min(M, full0[b]) : M) + (full1 ? min(N, full1[b]) : N);
                const float pos = 1.0f - (float)below_tot / (float)num_points;
                stop = stop || (pos > 0.95f);
            }
            if (stop) {  // active[l] = pairs still running at the START of layer l: this pair is gone from layer + 1 on
                stop_layer[b] = layer;
                for (int l = layer + 1; l < n_layers; ++l) atomicSub(active + l, 1);
            }
        }
    }
}

// One warp per token row: kept rows move to their compacted position in the OTHER buffer set (activation planes, the fp32
// copy the pooling MLPs read, rotary factors, original-index map).  Replaces desc[mask][None], enc[:, :, mask][:, None],
// ind[mask][None] of nets/adagml.py:347-355.
__global__ void __launch_bounds__(256) move_kernel(const int* __restrict__ dest, int B, int M, int N,
                                                   const __nv_bfloat16* __restrict__ s_hi, const __nv_bfloat16* __restrict__ s_lo,
                                                   __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo, long long ld_bf,
                                                   const float* __restrict__ s_f32, float* __restrict__ d_f32, long long ld_f32,
                                                   const float* __restrict__ s_cos, const float* __restrict__ s_sin,
                                                   float* __restrict__ d_cos, float* __restrict__ d_sin,
                                                   const int* __restrict__ s_ind, int* __restrict__ d_ind,
                                                   const int* __restrict__ pred) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B * (M + N) || pram_pred_skip(pred)) return;  // (once every pair has exited, the live buffers are dead state)
    int set, b, r, base;
    row_to_pair(row, B, M, N, set, b, r, base);
    const int d = dest[row];
    if (d < 0) return;
    const long long so = (long long)row, to = (long long)(base + d);
    reinterpret_cast<uint4*>(d_hi + to * ld_bf)[lane] = __ldg(reinterpret_cast<const uint4*>(s_hi + so * ld_bf) + lane);
    if (s_lo) reinterpret_cast<uint4*>(d_lo + to * ld_bf)[lane] = __ldg(reinterpret_cast<const uint4*>(s_lo + so * ld_bf) + lane);
    if (s_f32) {
        const float4* sp = reinterpret_cast<const float4*>(s_f32 + so * ld_f32);
        float4* dp = reinterpret_cast<float4*>(d_f32 + to * ld_f32);
        dp[lane] = __ldg(sp + lane);
        dp[lane + 32] = __ldg(sp + lane + 32);
    }
    d_cos[to * 32 + lane] = s_cos[so * 32 + lane];
    d_sin[to * 32 + lane] = s_sin[so * 32 + lane];
    if (lane == 0) d_ind[to] = s_ind[so];
}

// Pairs that stop at this layer keep the state they stop with: projected descriptors (out_proj[layer] of the pruned tokens,
// nets/adagml.py:375-377), index maps and counts.
__global__ void __launch_bounds__(256) latch_kernel(const int* __restrict__ stop_layer, int layer, int B, int M, int N,
                                                    const __nv_bfloat16* __restrict__ s_hi, const __nv_bfloat16* __restrict__ s_lo,
                                                    __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo,
                                                    const int* __restrict__ s_ind, int* __restrict__ d_ind,
                                                    const int* __restrict__ cnt0, const int* __restrict__ cnt1,
                                                    int* __restrict__ fcnt0, int* __restrict__ fcnt1,
                                                    const int* __restrict__ pred) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B * (M + N) || pram_pred_skip(pred)) return;
    int set, b, r, base;
    row_to_pair(row, B, M, N, set, b, r, base);
    if (stop_layer[b] != layer) return;
    const long long o = (long long)row * 256;
    reinterpret_cast<uint4*>(d_hi + o)[lane] = __ldg(reinterpret_cast<const uint4*>(s_hi + o) + lane);
    if (s_lo) reinterpret_cast<uint4*>(d_lo + o)[lane] = __ldg(reinterpret_cast<const uint4*>(s_lo + o) + lane);
    if (lane == 0) {
        d_ind[row] = s_ind[row];
        if (r == 0) { if (set == 0) fcnt0[b] = cnt0[b]; else fcnt1[b] = cnt1[b]; }
    }
}

// matches of the compacted problem scattered back to the full keypoint set (nets/adagml.py:383-394).
// phase 0: fill (-1, 0); phase 1: scatter.
__global__ void scatter_kernel(int phase, const long long* __restrict__ m0, const float* __restrict__ s0,
                               const int* __restrict__ ind0, const int* __restrict__ ind1, const int* __restrict__ fcnt0,
                               const int* __restrict__ fcnt1, int B, int M, int N, long long* __restrict__ full_i,
                               float* __restrict__ full_s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * M) return;
    if (phase == 0) { full_i[i] = -1; full_s[i] = 0.f; return; }
    const int b = (int)(i / M), r = (int)(i - (long long)b * M);
    if (r >= fcnt0[b]) return;
    const long long o = (long long)b * M + ind0[i];
    full_s[o] = s0[i];
    const long long j = m0[i];
    if (j >= 0 && j < fcnt1[b]) full_i[o] = ind1[(long long)B * M + (long long)b * N + j];
}

}  // namespace adagml

PRAM_API int pram_colmean_reduce(const float* colsum, int ld, int B, int heads, int N, int nq, const int* nq_counts,
                                 float* out, int out_stride, cudaStream_t stream) {
    if (!colsum || !out || B <= 0 || heads <= 0 || N <= 0 || nq <= 0 || ld < N || out_stride <= 0) return PRAM_ERR_ARG;
    adagml::colmean_reduce_kernel<<<cdiv((long long)B * N, 256), 256, 0, stream>>>(colsum, ld, B, heads, N, nq, nq_counts, out,
                                                                                 out_stride, g_pram_pred);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

PRAM_API int pram_adagml_prune(const float* conf_logits, int B, int M, int N, int* cnt0, int* cnt1, float threshold,
                               int n_min_tokens, int do_prune, int layer, int n_layers, int* dest, int* stop_layer,
                               int* active, int* trace, int* err, const int* full0, const int* full1, cudaStream_t stream) {
    const int is_last = layer == n_layers - 1;
    if (!conf_logits || !cnt0 || !cnt1 || !dest || !stop_layer || !active || !err || B <= 0 || M <= 0 || N <= 0)
        return PRAM_ERR_ARG;
    adagml::prune_kernel<<<B, 1024, 0, stream>>>(conf_logits, B, M, N, cnt0, cnt1, threshold, n_min_tokens, do_prune, layer,
                                                 is_last, n_layers, dest, stop_layer, active, trace, err, full0, full1,
                                                 g_pram_pred);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

PRAM_API int pram_adagml_move(const int* dest, int B, int M, int N, const void* src_hi, const void* src_lo, void* dst_hi,
                              void* dst_lo, long long ld_bf, const float* src_f32, float* dst_f32, long long ld_f32,
                              const float* src_cos, const float* src_sin, float* dst_cos, float* dst_sin, const int* src_ind,
                              int* dst_ind, cudaStream_t stream) {
    if (!dest || !src_hi || !dst_hi || !src_cos || !src_sin || !dst_cos || !dst_sin || !src_ind || !dst_ind || B <= 0 ||
        M <= 0 || N <= 0 || (ld_bf % 8) || (src_f32 && (!dst_f32 || (ld_f32 % 4))) || (src_lo && !dst_lo))
        return PRAM_ERR_ARG;
    adagml::move_kernel<<<cdiv((long long)B * (M + N), 8), 256, 0, stream>>>(
        dest, B, M, N, (const __nv_bfloat16*)src_hi, (const __nv_bfloat16*)src_lo, (__nv_bfloat16*)dst_hi, (__nv_bfloat16*)dst_lo,
        ld_bf, src_f32, dst_f32, ld_f32, src_cos, src_sin, dst_cos, dst_sin, src_ind, dst_ind, g_pram_pred);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

PRAM_API int pram_adagml_latch(const int* stop_layer, int layer, int B, int M, int N, const void* src_hi, const void* src_lo,
                               void* dst_hi, void* dst_lo, const int* src_ind, int* dst_ind, const int* cnt0, const int* cnt1,
                               int* final_cnt0, int* final_cnt1, cudaStream_t stream) {
    if (!stop_layer || !src_hi || !dst_hi || !src_ind || !dst_ind || !cnt0 || !cnt1 || !final_cnt0 || !final_cnt1 || B <= 0 ||
        M <= 0 || N <= 0 || (src_lo && !dst_lo))
        return PRAM_ERR_ARG;
    adagml::latch_kernel<<<cdiv((long long)B * (M + N), 8), 256, 0, stream>>>(
        stop_layer, layer, B, M, N, (const __nv_bfloat16*)src_hi, (const __nv_bfloat16*)src_lo, (__nv_bfloat16*)dst_hi,
        (__nv_bfloat16*)dst_lo, src_ind, dst_ind, cnt0, cnt1, final_cnt0, final_cnt1, g_pram_pred);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

PRAM_API int pram_adagml_scatter(const long long* matches0, const float* mscores0, const int* ind, const int* final_cnt0,
                                 const int* final_cnt1, int B, int M, int N, long long* full_matches0, float* full_scores0,
                                 cudaStream_t stream) {
    if (!matches0 || !mscores0 || !ind || !final_cnt0 || !final_cnt1 || !full_matches0 || !full_scores0 || B <= 0 || M <= 0 ||
        N <= 0)
        return PRAM_ERR_ARG;
    const int blocks = cdiv((long long)B * M, 256);
    for (int phase = 0; phase < 2; ++phase) {
        adagml::scatter_kernel<<<blocks, 256, 0, stream>>>(phase, matches0, mscores0, ind, ind, final_cnt0, final_cnt1, B, M, N,
                                                          full_matches0, full_scores0);
        PRAM_CHECK_LAUNCH();
    }
    return PRAM_OK;
}
