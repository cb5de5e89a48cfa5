// K15 / K16 of SURVEY.md section 2b: dustbin-augmented probability-domain Sinkhorn + mutual-arg-max match
// extraction.  Replaces reference nets/gml.py:27-46 (sinkhorn / sink_algorithm, ~125 launches each
// streaming an M x N temporary) and nets/gml.py:304-319 (compute_matches, ~13 launches) with ONE
// launch: a thread-block cluster per (M+1) x (N+1) problem.
//
//   p = softmax_rows([dist, bin; bin, bin])                       written once to the workspace
//   20 x { u_i = r_i / (sum_j p_ij v_j + 1e-8) ;  v_j = c_j / (sum_i p_ij u_i + 1e-8) }
//        -> ONE sweep over p per iteration: a warp finishes row i's dot product, gets u_i, and
//           immediately accumulates p_ij * u_i into its register-resident column partials
//   P_ij = (p_ij * u_i) * v_j ; row / column arg-max over the inner M x N ; mutual check ; threshold
//
// The rows of p are split across the G CTAs of the cluster; column partials are reduced in a fixed
// order (warp order inside a CTA, CTA-rank order across the cluster through distributed shared
// memory), so results are bit-reproducible run to run.  u, v, r, c never touch HBM.
// Same algebra and eps placement as the reference (probability domain, eps inside the divisor).
#include "common.cuh"
#include <stdlib.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

constexpr int SK_MAXG = 16;  // 16 = non-portable cluster size (opt-in per kernel); used when few problems leave most SMs idle

template <int NV, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) sinkhorn_match_kernel(
    const float* __restrict__ dist, int M, int N, const float* __restrict__ bin_ptr, int iters,
    float th, float* __restrict__ pws, int ldp, long long* __restrict__ matches0,
    long long* __restrict__ matches1, float* __restrict__ mscores0, float* __restrict__ mscores1,
    int* __restrict__ idx0_ws, int* __restrict__ idx1_ws, float* __restrict__ max0_ws, int G,
    const int* __restrict__ m_counts, const int* __restrict__ n_counts) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float smem[];
    const int NC = N + 1, MR = M + 1;
    float* v_s = smem;                // [ldp]   column scaling, replicated in every CTA
    float* col_s = smem + ldp;        // [ldp]   this CTA's partial column sums
    float* u_s = smem + 2 * ldp;      // [rows_local]
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.y;
    const int rows_per = (MR + G - 1) / G;
    float* part_s = u_s + ((rows_per + 3) & ~3);   // [WARPS][ldp] per-warp column partials of one sweep
    const int r0 = min(rank * rows_per, MR), r1 = min(r0 + rows_per, MR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float bin = *bin_ptr;
    const float* D = dist + (long long)b * M * N;
    float* P = pws + (long long)b * MR * ldp;
    const float eps = 1e-8f;
    // per-pair problem size inside the padded [M, N] block (the batched pipeline pads frames with fewer keypoints): rows
    // mb..M-1 and columns nb..N-1 are absent; the dustbin row / column stay at index M / N of the padded layout
    const int mb = m_counts ? max(1, min(M, m_counts[b])) : M;
    const int nb = n_counts ? max(1, min(N, n_counts[b])) : N;
    auto row_ok = [&](int i) { return i < mb || i == M; };
    auto col_ok = [&](int j) { return j < nb || j == N; };

    // ---- phase 0: p = softmax over each augmented row ----
    for (int i = r0 + warp; i < r1; i += WARPS) {
        float z[NV * 4];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int j = k * 128 + lane * 4 + q;
                float val = -INFINITY;
                if (j < NC && col_ok(j) && row_ok(i)) val = (i < M && j < N) ? D[(long long)i * N + j] : bin;
                z[k * 4 + q] = val;
                mx = fmaxf(mx, val);
            }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < NV * 4; ++k) {
            z[k] = (z[k] == -INFINITY) ? 0.f : expf(z[k] - mx);
            sum += z[k];
        }
        sum = warp_sum(sum);
        if (!row_ok(i)) sum = 1.f;  // absent row: all zeros
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            int j = k * 128 + lane * 4;
            if (j < ldp)
                *reinterpret_cast<float4*>(P + (long long)i * ldp + j) =
                    make_float4(z[k * 4] / sum, z[k * 4 + 1] / sum, z[k * 4 + 2] / sum, z[k * 4 + 3] / sum);
        }
    }
    for (int j = threadIdx.x; j < ldp; j += WARPS * 32) v_s[j] = (j < NC && col_ok(j)) ? 1.f : 0.f;
    __syncthreads();

    const int chunk = (ldp / 4 + G - 1) / G * 4;  // columns reduced by each rank (multiple of 4)
    // ---- Sinkhorn iterations ----
    for (int it = 0; it < iters; ++it) {
        float4 cp[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) cp[k] = make_float4(0, 0, 0, 0);
        // rows are streamed from L2; the next row of this warp is fetched while the current one is reduced (the
        // warp_sum in the middle of the body would otherwise leave a single row of loads in flight per warp)
        constexpr bool kCacheRow = (NV <= 17);  // wider rows are re-read (L1-resident) instead of cached in registers
        float4 nx[kCacheRow ? NV : 1];
        auto load_row = [&](int i, float4 (&dst)[kCacheRow ? NV : 1]) {
            if (!kCacheRow) return;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int j = k * 128 + lane * 4;
                dst[kCacheRow ? k : 0] = (j < ldp && i < r1) ? *reinterpret_cast<const float4*>(P + (long long)i * ldp + j)
                                                              : make_float4(0, 0, 0, 0);
            }
        };
        load_row(r0 + warp, nx);
        for (int i = r0 + warp; i < r1; i += WARPS) {
            float4 pr[kCacheRow ? NV : 1];
            float s = 0.f;
            if (kCacheRow) {
#pragma unroll
                for (int k = 0; k < NV; ++k) pr[kCacheRow ? k : 0] = nx[kCacheRow ? k : 0];
                load_row(i + WARPS, nx);  // prefetch (zeros past the last row)
            }
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                int j = k * 128 + lane * 4;
                float4 pv = make_float4(0, 0, 0, 0);
                if (j < ldp) {
                    pv = kCacheRow ? pr[kCacheRow ? k : 0] : *reinterpret_cast<const float4*>(P + (long long)i * ldp + j);
                    float4 vv = *reinterpret_cast<const float4*>(v_s + j);
                    s += pv.x * vv.x + pv.y * vv.y + pv.z * vv.z + pv.w * vv.w;
                }
            }
            s = warp_sum(s);
            const float ri = (i == M) ? (float)(mb + 1) : 1.f;
            const float ui = row_ok(i) ? ri / (s + eps) : 0.f;
            if (lane == 0) u_s[i - r0] = ui;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                float4 pv;
                if (kCacheRow) {
                    pv = pr[kCacheRow ? k : 0];
                } else {
                    int j = k * 128 + lane * 4;
                    pv = (j < ldp) ? *reinterpret_cast<const float4*>(P + (long long)i * ldp + j)
                                   : make_float4(0, 0, 0, 0);
                }
                cp[k].x += pv.x * ui; cp[k].y += pv.y * ui;
                cp[k].z += pv.z * ui; cp[k].w += pv.w * ui;
            }
        }
        // fixed-order reduction of the per-warp partials into col_s: every warp parks its partials, then each thread adds the
        // WARPS values of its four columns in warp order -- the same order (and bits) as the former one-warp-at-a-time
        // accumulation, with one block barrier instead of WARPS of them
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int j = k * 128 + lane * 4;
            if (j < ldp) *reinterpret_cast<float4*>(part_s + (size_t)warp * ldp + j) = cp[k];
        }
        __syncthreads();
        for (int j = threadIdx.x * 4; j < ldp; j += WARPS * 32 * 4) {
            float4 a = *reinterpret_cast<const float4*>(part_s + j);
#pragma unroll
            for (int w = 1; w < WARPS; ++w) {
                const float4 o = *reinterpret_cast<const float4*>(part_s + (size_t)w * ldp + j);
                a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
            }
            *reinterpret_cast<float4*>(col_s + j) = a;
        }
        cluster.sync();
        // rank-ordered reduction of this rank's column chunk over all CTAs, then broadcast v
        {
            const int c0 = rank * chunk, c1 = min(c0 + chunk, ldp);
            for (int j = c0 + threadIdx.x; j < c1; j += WARPS * 32) {
                float s = 0.f;
                for (int g = 0; g < G; ++g) s += cluster.map_shared_rank(col_s, g)[j];
                const float cj = (j == N) ? (float)(nb + 1) : 1.f;
                const float vj = (j < NC && col_ok(j)) ? cj / (s + eps) : 0.f;
                for (int g = 0; g < G; ++g) cluster.map_shared_rank(v_s, g)[j] = vj;
            }
        }
        cluster.sync();
    }

    // ---- final scaling: P = (p*u)*v, row arg-max over j < N for rows i < M; P written back ----
    for (int i = r0 + warp; i < r1; i += WARPS) {
        const float ui = (iters > 0) ? u_s[i - r0] : 1.f;
        float best = -1.f;
        int bj = 0;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            int j = k * 128 + lane * 4;
            if (j < ldp) {
                float4 pv = *reinterpret_cast<const float4*>(P + (long long)i * ldp + j);
                float4 vv = *reinterpret_cast<const float4*>(v_s + j);
                pv.x = (pv.x * ui) * vv.x; pv.y = (pv.y * ui) * vv.y;
                pv.z = (pv.z * ui) * vv.z; pv.w = (pv.w * ui) * vv.w;
                *reinterpret_cast<float4*>(P + (long long)i * ldp + j) = pv;
                if (j < nb && pv.x > best) { best = pv.x; bj = j; }
                if (j + 1 < nb && pv.y > best) { best = pv.y; bj = j + 1; }
                if (j + 2 < nb && pv.z > best) { best = pv.z; bj = j + 2; }
                if (j + 3 < nb && pv.w > best) { best = pv.w; bj = j + 3; }
            }
        }
        // warp arg-max, first index wins ties
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
        if (lane == 0 && i < M) { idx0_ws[(long long)b * M + i] = bj; max0_ws[(long long)b * M + i] = best; }
    }
    __threadfence();
    cluster.sync();
    // ---- column arg-max over i < mb for this rank's columns j < N ----
    // The rank's columns are taken in blocks of 32 float4 groups (128 columns): lane = group, warp = a band of rows, four
    // running maxima per thread from 16-byte loads; the bands meet in shared memory and are merged in row order with a
    // strict comparison, so the first maximal row wins exactly as in a top-to-bottom scan.  (One thread per column walking
    // all rows -- 1024 dependent-latency loads on half the threads -- was a 45 us serial tail of every launch.)
    {
        const int ngroups = (N + 3) / 4;                       // float4 groups that hold columns j < N
        const int gper = (ngroups + G - 1) / G;
        const int g0 = rank * gper, g1 = min(g0 + gper, ngroups);
        const int rows_pw = (mb + WARPS - 1) / WARPS;
        const int ra = min(warp * rows_pw, mb), rb = min(ra + rows_pw, mb);
        float* sb = part_s;                                     // [WARPS][128] best values
        int* si = reinterpret_cast<int*>(part_s + WARPS * 128);  // [WARPS][128] their rows   (8 KB + 8 KB <= WARPS * ldp floats)
        for (int gb = g0; gb < g1; gb += 32) {
            const int g = gb + lane;
            float best[4] = {-1.f, -1.f, -1.f, -1.f};
            int bi[4] = {0, 0, 0, 0};
            if (g < g1) {
                const float* col = P + 4 * g;
                for (int i = ra; i < rb; ++i) {
                    const float4 pv = *reinterpret_cast<const float4*>(col + (long long)i * ldp);
                    if (pv.x > best[0]) { best[0] = pv.x; bi[0] = i; }
                    if (pv.y > best[1]) { best[1] = pv.y; bi[1] = i; }
                    if (pv.z > best[2]) { best[2] = pv.z; bi[2] = i; }
                    if (pv.w > best[3]) { best[3] = pv.w; bi[3] = i; }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { sb[warp * 128 + lane * 4 + q] = best[q]; si[warp * 128 + lane * 4 + q] = bi[q]; }
            __syncthreads();
            if (threadIdx.x < 128) {
                const int j = 4 * gb + threadIdx.x;
                float bbest = -1.f;
                int bbi = 0;
                for (int w = 0; w < WARPS; ++w) {
                    const float v = sb[w * 128 + threadIdx.x];
                    if (v > bbest) { bbest = v; bbi = si[w * 128 + threadIdx.x]; }
                }
                if (j < N && 4 * gb + threadIdx.x < 4 * g1) idx1_ws[(long long)b * N + j] = bbi;
            }
            __syncthreads();
        }
    }
    __threadfence();
    cluster.sync();
    // ---- mutual check + threshold (reference nets/gml.py:304-319) ----
    const int* i0 = idx0_ws + (long long)b * M;
    const int* i1 = idx1_ws + (long long)b * N;
    const float* m0 = max0_ws + (long long)b * M;
    const int gthreads = G * WARPS * 32, gtid = rank * WARPS * 32 + threadIdx.x;
    for (int i = gtid; i < M; i += gthreads) {
        const int j = i0[i];
        const bool mutual = (i < mb) && (j < nb) && (i1[j] == i);
        const float s0 = mutual ? m0[i] : 0.f;
        matches0[(long long)b * M + i] = (mutual && s0 > th) ? (long long)j : -1ll;
        mscores0[(long long)b * M + i] = s0;
    }
    for (int j = gtid; j < N; j += gthreads) {
        const int i = i1[j];
        const bool mutual1 = (j < nb) && (i < mb) && (i0[i] == j);
        // mscores0[i] recomputed locally: s0_i = (i1[i0[i]] == i) ? max0[i] : 0
        const bool mutual0_i = (i < mb) && (i0[i] < nb) && (i1[i0[i]] == i);
        const float s0_i = mutual0_i ? m0[i] : 0.f;
        const bool valid0_i = mutual0_i && s0_i > th;
        if (matches1) matches1[(long long)b * N + j] = (mutual1 && valid0_i) ? (long long)i : -1ll;
        if (mscores1) mscores1[(long long)b * N + j] = mutual1 ? s0_i : 0.f;
    }
}

template <int NV, int WARPS>
static int launch_sinkhorn(const float* dist, int B, int M, int N, const float* bin, int iters, float th,
                           float* pws, int ldp, long long* m0, long long* m1, float* s0, float* s1,
                           int* idx0, int* idx1, float* max0, int G, const int* mc, const int* nc, cudaStream_t stream) {
    auto kern = sinkhorn_match_kernel<NV, WARPS>;
    const int rows_per = (M + 1 + G - 1) / G;
    // v | column partials | u | per-warp partials of a sweep, re-used by the column arg-max (2 x WARPS x 128 words)
    const size_t part = (size_t)WARPS * ((size_t)ldp > 256 ? (size_t)ldp : 256);
    size_t smem = sizeof(float) * (2 * (size_t)ldp + ((rows_per + 3) & ~3) + part);
    PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (G > 8) PRAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G, B, 1);
    cfg.blockDim = dim3(WARPS * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PRAM_CUDA(cudaLaunchKernelEx(&cfg, kern, dist, M, N, bin, iters, th, pws, ldp, m0, m1, s0, s1, idx0,
                                 idx1, max0, G, mc, nc));
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// workspace: pws  float [B][(M+1)][ldp], ldp = round_up(N+1, 4)   (holds P on return)
//            iws  int   [B][M + N] , fws float [B][M]
PRAM_API long long pram_sinkhorn_workspace_floats(int B, int M, int N) {
    long long ldp = (N + 1 + 3) / 4 * 4;
    return (long long)B * (M + 1) * ldp;
}

PRAM_API int pram_sinkhorn_match(const float* dist, int B, int M, int N, const float* bin_score, int iters,
                                 float threshold, float* pws, int* iws, float* fws, long long* matches0,
                                 long long* matches1, float* mscores0, float* mscores1, int cluster,
                                 const int* m_counts, const int* n_counts, cudaStream_t stream) {
    if (!dist || !bin_score || !pws || !iws || !fws || !matches0 || !mscores0 || B <= 0 || M <= 0 || N <= 0)
        return PRAM_ERR_ARG;
    int G = cluster;
    if (G <= 0) {
        // largest cluster that still runs all B problems in ONE wave of CTAs (a second, partial wave costs a full
        // 20-iteration pass); a single problem keeps the portable maximum of 8 CTAs
        static int sms = 0;
        if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
        // (16-CTA clusters for up to 4 problems: a single 1024 x 1024 problem is bound by the latency of its row sweeps,
        // 8 rows per warp and sweep with 8 CTAs; PRAM_SINKHORN_MAXG=8 keeps the portable size)
        static int maxg = 0;
        if (!maxg) { const char* e = getenv("PRAM_SINKHORN_MAXG"); maxg = e ? atoi(e) : SK_MAXG; if (maxg < 1 || maxg > SK_MAXG) maxg = SK_MAXG; }
        G = maxg;
        while (G > 1 && ((long long)B * G > sms || (G > 8 && B > 4))) G >>= 1;
    }
    if (G > SK_MAXG || (G & (G - 1))) return PRAM_ERR_ARG;
    const int ldp = (N + 1 + 3) / 4 * 4;
    int* idx0 = iws;
    int* idx1 = iws + (long long)B * M;
    const int nv = (ldp + 127) / 128;
    if (nv <= 9) {
        int rc = launch_sinkhorn<9, 16>(dist, B, M, N, bin_score, iters, threshold, pws, ldp, matches0, matches1,
                                        mscores0, mscores1, idx0, idx1, fws, G, m_counts, n_counts, stream);
        if (rc != PRAM_OK && G > 8 && cluster <= 0) {  // the opt-in 16-CTA cluster could not be placed on this part: portable size
            cudaGetLastError();
            rc = launch_sinkhorn<9, 16>(dist, B, M, N, bin_score, iters, threshold, pws, ldp, matches0, matches1,
                                        mscores0, mscores1, idx0, idx1, fws, 8, m_counts, n_counts, stream);
        }
        return rc;
    }
    if (nv <= 17)
        return launch_sinkhorn<17, 8>(dist, B, M, N, bin_score, iters, threshold, pws, ldp, matches0, matches1,
                                       mscores0, mscores1, idx0, idx1, fws, G, m_counts, n_counts, stream);
    if (nv <= 33)
        return launch_sinkhorn<33, 8>(dist, B, M, N, bin_score, iters, threshold, pws, ldp, matches0, matches1,
                                      mscores0, mscores1, idx0, idx1, fws, G, m_counts, n_counts, stream);
    return PRAM_ERR_UNSUPPORTED;
}
