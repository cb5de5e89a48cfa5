// FP32 CUDA-core kernels: the exact-arithmetic path of the networks (and the on-device reference the
// tcgen05 path is validated against).  NHWC activations, BN folded into weights/bias on the host.
//   * conv_igemm_kernel : implicit-GEMM 3x3 / 1x1 convolution == Linear layer (H=1), fused
//                         bias + residual + ReLU, strided input/output rows (writes straight into a
//                         concat buffer)                         -> reference nets/sfd2.py:141-170,
//                                                                   nn.Linear in segnetvit.py / gml.py
//   * gconv3x3_kernel   : grouped 3x3 convolution, 32 groups x 8 channels (ResBlock.conv2,
//                         reference nets/sfd2.py:101)
//   * l2norm_rows       : channel L2 normalisation (F.normalize, reference nets/sfd2.py:333)
//   * layernorm_gelu    : LayerNorm + exact (erf) GELU, reference nets/segnetvit.py:92-93
//   * rotary_split      : de-interleaved qkv -> rotary(q), rotary(k), v in [B,h,N,64]
//                         (reference nets/segnetvit.py:98-103)
//   * attention_kernel  : softmax(QK^T * scale) V, flash-style (no N x N matrix in HBM), fp32,
//                         optional per-key column mean of the attention (AdaGML, adagml.py:148)
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// implicit GEMM:  out[m][n] = act( sum_{tap,ci} in[pix(m,tap)][ci] * w[tap][ci][n] + bias[n] + res[m][n] )
// ------------------------------------------------------------------------------------------
constexpr int IG_BM = 128, IG_BN = 64, IG_BK = 16, IG_THREADS = 256;

struct ConvParams {
    const float* in; const float* w; const float* bias; const float* res; float* out;
    int B, H, W, Cin, Ho, Wo, Cout, ksize, stride;
    long long in_pix_stride;   // floats between consecutive input pixels (>= Cin)
    long long out_pix_stride;  // floats between consecutive output pixels (>= Cout)
    long long res_pix_stride;
    int relu;
    int w_layout;              // 0: w[tap][Cin][Cout]   1: w[Cout][Cin] (ksize 1; torch Linear layout)
    long long in_batch_stride, w_batch_stride, out_batch_stride;  // blockIdx.z batching (bmm)
    const int* pred;           // launch predicate (common.cuh), NULL = always run
};

__global__ void __launch_bounds__(IG_THREADS) conv_igemm_kernel(ConvParams p) {
    __shared__ float As[IG_BK][IG_BM + 4];
    __shared__ float Bs[IG_BK][IG_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, micro-tile 8 (m) x 4 (n)
    if (pram_pred_skip(p.pred)) return;
    p.in += blockIdx.z * p.in_batch_stride;
    p.w += blockIdx.z * p.w_batch_stride;
    p.out += blockIdx.z * p.out_batch_stride;
    const long long M = (long long)p.B * p.Ho * p.Wo;
    const long long m0 = (long long)blockIdx.x * IG_BM;
    const int n0 = blockIdx.y * IG_BN;
    const int pad = p.ksize / 2;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // A-load assignment: 128 pixels x 16 channels -> thread handles pixel (tid>>1), channels (tid&1)*8..+8
    const int a_m = tid >> 1, a_k = (tid & 1) * 8;
    const long long am = m0 + a_m;
    int ab = 0, aoy = 0, aox = 0;
    const bool a_valid_m = am < M;
    if (a_valid_m) {
        ab = (int)(am / ((long long)p.Ho * p.Wo));
        int rem = (int)(am - (long long)ab * p.Ho * p.Wo);
        aoy = rem / p.Wo;
        aox = rem - aoy * p.Wo;
    }
    // B-load assignment: 16 k x 64 n -> thread handles k = tid>>4, n = (tid&15)*4..+4
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;
    const bool vec_a = (p.Cin % 4 == 0) && (p.in_pix_stride % 4 == 0);
    const bool vec_b = (p.Cout % 4 == 0);

    const int taps = p.ksize * p.ksize;
    for (int tap = 0; tap < taps; ++tap) {
        const int dy = tap / p.ksize - pad, dx = tap % p.ksize - pad;
        const int iy = aoy * p.stride + dy, ix = aox * p.stride + dx;
        const bool pix_ok = a_valid_m && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        const float* ain = p.in + (((long long)ab * p.H + iy) * p.W + ix) * p.in_pix_stride;
        const float* wt = p.w + (long long)tap * p.Cin * p.Cout;
        for (int k0 = 0; k0 < p.Cin; k0 += IG_BK) {
            // ---- stage A ----
            float av[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) av[q] = 0.f;
            if (pix_ok) {
                int c = k0 + a_k;
                if (vec_a && c + 8 <= p.Cin) {
                    float4 v0 = *reinterpret_cast<const float4*>(ain + c);
                    float4 v1 = *reinterpret_cast<const float4*>(ain + c + 4);
                    av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
                    av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) if (c + q < p.Cin) av[q] = ain[c + q];
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) As[a_k + q][a_m] = av[q];
            // ---- stage B ----
            if (p.w_layout == 0) {
                int k = k0 + b_k, n = n0 + b_n;
                float bv[4] = {0.f, 0.f, 0.f, 0.f};
                if (k < p.Cin) {
                    const float* wp = wt + (long long)k * p.Cout + n;
                    if (vec_b && n + 4 <= p.Cout) {
                        float4 v = *reinterpret_cast<const float4*>(wp);
                        bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) if (n + q < p.Cout) bv[q] = wp[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) Bs[b_k][b_n + q] = bv[q];
            } else {
                // [Cout][Cin]: thread handles n = tid>>2, k = (tid&3)*4..+4 (contiguous along Cin)
                const int n = n0 + (tid >> 2), kk = (tid & 3) * 4, k = k0 + kk;
                float bv[4] = {0.f, 0.f, 0.f, 0.f};
                if (n < p.Cout) {
                    const float* wp = p.w + (long long)n * p.Cin + k;
                    if ((p.Cin % 4 == 0) && k + 4 <= p.Cin) {
                        float4 v = *reinterpret_cast<const float4*>(wp);
                        bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) if (k + q < p.Cin) bv[q] = wp[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) Bs[kk + q][tid >> 2] = bv[q];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < IG_BK; ++k) {
                float a[8], b[4];
                float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
                float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    // ---- epilogue ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= p.Cout) continue;
            float v = acc[i][j];
            if (p.bias) v += p.bias[n];
            if (p.res) v += p.res[m * p.res_pix_stride + n];
            if (p.relu) v = fmaxf(v, 0.f);
            p.out[m * p.out_pix_stride + n] = v;
        }
    }
}

PRAM_API int pram_conv_f32(const float* in, long long in_pix_stride, const float* w, const float* bias,
                           const float* res, long long res_pix_stride, float* out,
                           long long out_pix_stride, int B, int H, int W, int Cin, int Cout, int ksize,
                           int stride, int relu, cudaStream_t stream) {
    if (!in || !w || !out || B <= 0 || (ksize != 1 && ksize != 3) || (stride != 1 && stride != 2))
        return PRAM_ERR_ARG;
    ConvParams p;
    p.in = in; p.w = w; p.bias = bias; p.res = res; p.out = out;
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ksize = ksize; p.stride = stride;
    const int pad = ksize / 2;
    p.Ho = (H + 2 * pad - ksize) / stride + 1;
    p.Wo = (W + 2 * pad - ksize) / stride + 1;
    p.in_pix_stride = in_pix_stride; p.out_pix_stride = out_pix_stride; p.res_pix_stride = res_pix_stride;
    p.relu = relu; p.w_layout = 0;
    p.in_batch_stride = p.w_batch_stride = p.out_batch_stride = 0;
    p.pred = nullptr;
    long long M = (long long)B * p.Ho * p.Wo;
    dim3 grid(cdiv(M, IG_BM), cdiv(Cout, IG_BN));
    conv_igemm_kernel<<<grid, IG_THREADS, 0, stream>>>(p);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// out[z][m][n] = act( sum_k a[z][m][k] * w[z][n][k] + bias[n] + res[m][n] )   (torch Linear / bmm-NT)
PRAM_API int pram_linear_f32(const float* a, long long lda, const float* w, const float* bias,
                             const float* res, long long ldres, float* out, long long ldo, long long rows,
                             int K, int N, int relu, int batch, long long a_batch_stride,
                             long long w_batch_stride, long long out_batch_stride, cudaStream_t stream) {
    if (!a || !w || !out || rows <= 0 || K <= 0 || N <= 0 || batch <= 0) return PRAM_ERR_ARG;
    ConvParams p;
    p.in = a; p.w = w; p.bias = bias; p.res = res; p.out = out;
    p.B = 1; p.H = 1; p.W = (int)rows; p.Cin = K; p.Cout = N; p.ksize = 1; p.stride = 1;
    p.Ho = 1; p.Wo = (int)rows;
    p.in_pix_stride = lda; p.out_pix_stride = ldo; p.res_pix_stride = ldres;
    p.relu = relu; p.w_layout = 1;
    p.in_batch_stride = a_batch_stride; p.w_batch_stride = w_batch_stride; p.out_batch_stride = out_batch_stride;
    p.pred = g_pram_pred;
    dim3 grid(cdiv(rows, IG_BM), cdiv(N, IG_BN), batch);
    conv_igemm_kernel<<<grid, IG_THREADS, 0, stream>>>(p);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// ------------------------------------------------------------------------------------------
// grouped 3x3 conv, G groups of 8 channels, stride 1, pad 1, + bias + ReLU.
// lanes <-> groups (a warp reads one pixel's 256 channels as one 1 KB line), 4 consecutive output
// pixels per thread so each shared-memory weight read feeds 4 FMAs.
// weights in smem as [tap][ci][co][group].
// ------------------------------------------------------------------------------------------
constexpr int GC_PX = 8;   // output pixels per thread: each shared-memory weight read feeds 8 FMAs (FMA-bound)
__global__ void __launch_bounds__(256) gconv3x3_kernel(const float* __restrict__ in,
                                                       const float* __restrict__ w,  // [9][8][8][G]
                                                       const float* __restrict__ bias,
                                                       float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                                                       __nv_bfloat16* __restrict__ out_lo, int B, int H, int W,
                                                       int G, int relu) {
    extern __shared__ float ws[];
    const int C = G * 8;
    for (int i = threadIdx.x; i < 9 * 64 * G; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    const int g = threadIdx.x % G;
    const int slot = threadIdx.x / G;            // which pixel group inside the block
    const int quads_per_block = blockDim.x / G;  // 8 for G=32
    const int wq = cdiv(W, GC_PX);
    const long long total = (long long)B * H * wq;
    // persistent: the 72 KB weight tile is loaded once per CTA, then the CTA strides over pixel groups
    for (long long quad = (long long)blockIdx.x * quads_per_block + slot; quad < total;
         quad += (long long)gridDim.x * quads_per_block) {
        const int b = (int)(quad / ((long long)H * wq));
        const int rem = (int)(quad - (long long)b * H * wq);
        const int y = rem / wq, x0 = (rem - y * wq) * GC_PX;
        float acc[GC_PX][8];
#pragma unroll
        for (int p = 0; p < GC_PX; ++p)
#pragma unroll
            for (int co = 0; co < 8; ++co) acc[p][co] = 0.f;
#pragma unroll 1
        for (int r = 0; r < 3; ++r) {
            const int iy = y + r - 1;
            if (iy < 0 || iy >= H) continue;
            float v[GC_PX + 2][8];
#pragma unroll
            for (int q = 0; q < GC_PX + 2; ++q) {
                const int ix = x0 + q - 1;
                if (ix >= 0 && ix < W) {
                    const float* ip = in + (((long long)b * H + iy) * W + ix) * C + g * 8;
                    float4 a = *reinterpret_cast<const float4*>(ip);
                    float4 c = *reinterpret_cast<const float4*>(ip + 4);
                    v[q][0] = a.x; v[q][1] = a.y; v[q][2] = a.z; v[q][3] = a.w;
                    v[q][4] = c.x; v[q][5] = c.y; v[q][6] = c.z; v[q][7] = c.w;
                } else {
#pragma unroll
                    for (int ci = 0; ci < 8; ++ci) v[q][ci] = 0.f;
                }
            }
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int ci = 0; ci < 8; ++ci)
#pragma unroll
                    for (int co = 0; co < 8; ++co) {
                        const float wv = ws[(((r * 3 + s) * 8 + ci) * 8 + co) * G + g];
#pragma unroll
                        for (int p = 0; p < GC_PX; ++p) acc[p][co] = fmaf(v[p + s][ci], wv, acc[p][co]);
                    }
        }
#pragma unroll
        for (int p = 0; p < GC_PX; ++p) {
            const int x = x0 + p;
            if (x >= W) break;
            float o[8];
#pragma unroll
            for (int co = 0; co < 8; ++co) {
                float t = acc[p][co] + (bias ? bias[g * 8 + co] : 0.f);
                o[co] = relu ? fmaxf(t, 0.f) : t;
            }
            const long long off = (((long long)b * H + y) * W + x) * C + g * 8;
            if (out) {
                *reinterpret_cast<float4*>(out + off) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(out + off + 4) = make_float4(o[4], o[5], o[6], o[7]);
            }
            if (out_hi) {
                __nv_bfloat16 h[8], l[8];
#pragma unroll
                for (int co = 0; co < 8; ++co) {
                    h[co] = __float2bfloat16_rn(o[co]);
                    l[co] = __float2bfloat16_rn(o[co] - __bfloat162float(h[co]));
                }
                *reinterpret_cast<uint4*>(out_hi + off) = *reinterpret_cast<uint4*>(h);
                if (out_lo) *reinterpret_cast<uint4*>(out_lo + off) = *reinterpret_cast<uint4*>(l);
            }
        }
    }
}

static int gconv_launch(const float* in, const float* w, const float* bias, float* out, void* out_hi, void* out_lo,
                        int B, int H, int W, int groups, int relu, cudaStream_t stream);

PRAM_API int pram_gconv3x3_f32(const float* in, const float* w, const float* bias, float* out, int B,
                               int H, int W, int groups, int relu, cudaStream_t stream) {
    if (!out) return PRAM_ERR_ARG;
    return gconv_launch(in, w, bias, out, nullptr, nullptr, B, H, W, groups, relu, stream);
}

// same, with the result also (or only) emitted as split bf16 planes for a following tensor-core GEMM
PRAM_API int pram_gconv3x3_split(const float* in, const float* w, const float* bias, float* out_f32, void* out_hi,
                                 void* out_lo, int B, int H, int W, int groups, int relu, cudaStream_t stream) {
    if (!out_f32 && !out_hi) return PRAM_ERR_ARG;
    return gconv_launch(in, w, bias, out_f32, out_hi, out_lo, B, H, W, groups, relu, stream);
}

static int gconv_launch(const float* in, const float* w, const float* bias, float* out, void* out_hi, void* out_lo,
                        int B, int H, int W, int groups, int relu, cudaStream_t stream) {
    if (!in || !w || groups != 32) return PRAM_ERR_UNSUPPORTED;
    size_t smem = sizeof(float) * 9 * 64 * groups;
    static bool attr_set = false;
    if (!attr_set) {
        PRAM_CUDA(cudaFuncSetAttribute(gconv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        attr_set = true;
    }
    long long quads = (long long)B * H * cdiv(W, GC_PX);
    int qpb = 256 / groups;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const long long want = cdiv(quads, qpb);
    const int grid = (int)(want < 2LL * sms ? want : 2LL * sms);
    gconv3x3_kernel<<<grid, 256, smem, stream>>>(in, w, bias, out, (__nv_bfloat16*)out_hi,
                                                             (__nv_bfloat16*)out_lo, B, H, W, groups, relu);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// ------------------------------------------------------------------------------------------
// row-wise L2 normalisation (F.normalize over channels of an NHWC map), warp per row, in place ok
// ------------------------------------------------------------------------------------------
__global__ void l2norm_rows_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows,
                                   int C) {
    long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* p = in + row * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { float v = p[c]; ss += v * v; }
    float d = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
    for (int c = lane; c < C; c += 32) out[row * C + c] = p[c] / d;
}

PRAM_API int pram_l2norm_rows(const float* in, float* out, long long rows, int C, cudaStream_t stream) {
    if (!in || !out) return PRAM_ERR_ARG;
    l2norm_rows_kernel<<<cdiv(rows * 32, 256), 256, 0, stream>>>(in, out, rows, C);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// ------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5, biased variance) + exact GELU, warp per row, in place ok
// ------------------------------------------------------------------------------------------
// single-pass variant for C % 128 == 0 (256 / 512 / 1024): the row lives in registers (NV float4 per
// lane), one 16-byte load and one 16-byte (fp32) / 8-byte (bf16 plane) store per element group
// erf by Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7 + fp32 rounding): one rcp, one ex2, five FMAs instead of the
// ~25-instruction erff.  Used only by the split-bf16 (tensor-core path) variant, whose outputs are rounded to 2 x bf16
// (2^-17 relative) anyway; the fp32 exact-arithmetic path keeps erff.
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

template <int NV, bool FAST_ERF>
__global__ void __launch_bounds__(256) layernorm_gelu_vec_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float* __restrict__ out,
                                                                __nv_bfloat16* __restrict__ out_hi,
                                                                __nv_bfloat16* __restrict__ out_lo, long long rows, int gelu,
                                                                const int* __restrict__ pred) {
    constexpr int C = NV * 128;
    if (pram_pred_skip(pred)) return;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* p = reinterpret_cast<const float4*>(in + row * C);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        v[k] = p[lane + 32 * k];
        s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) / C + 1e-5f);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int c4 = lane + 32 * k;
        const float4 g = reinterpret_cast<const float4*>(gamma)[c4], bt = reinterpret_cast<const float4*>(beta)[c4];
        float o[4] = {(v[k].x - mean) * rstd * g.x + bt.x, (v[k].y - mean) * rstd * g.y + bt.y,
                      (v[k].z - mean) * rstd * g.z + bt.z, (v[k].w - mean) * rstd * g.w + bt.w};
        if (gelu) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                o[j] = 0.5f * o[j] * (1.f + (FAST_ERF ? erf_as(o[j] * 0.70710678118654752440f) : erff(o[j] * 0.70710678118654752440f)));
        }
        if (out) reinterpret_cast<float4*>(out + row * C)[c4] = make_float4(o[0], o[1], o[2], o[3]);
        if (out_hi) {
            __nv_bfloat16 h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[j] = __float2bfloat16_rn(o[j]);
                l[j] = __float2bfloat16_rn(o[j] - __bfloat162float(h[j]));
            }
            reinterpret_cast<uint2*>(out_hi + row * C)[c4] = *reinterpret_cast<uint2*>(h);
            if (out_lo) reinterpret_cast<uint2*>(out_lo + row * C)[c4] = *reinterpret_cast<uint2*>(l);
        }
    }
}

static bool layernorm_vec_launch(const float* in, const float* gamma, const float* beta, float* out, void* hi, void* lo,
                                 long long rows, int C, int gelu, cudaStream_t stream) {
    const int grid = cdiv(rows * 32, 256);
    __nv_bfloat16* h = (__nv_bfloat16*)hi;
    __nv_bfloat16* l = (__nv_bfloat16*)lo;
    // the fp32 output path is the exact-arithmetic mode (erff); split-bf16-only output may use the cheaper erf
    const bool fast = (out == nullptr);
    switch (C) {
        case 256: if (fast) layernorm_gelu_vec_kernel<2, true><<<grid, 256, 0, stream>>>(in, gamma, beta, out, h, l, rows, gelu, g_pram_pred);
                  else layernorm_gelu_vec_kernel<2, false><<<grid, 256, 0, stream>>>(in, gamma, beta, out, h, l, rows, gelu, g_pram_pred);
                  return true;
        case 512: if (fast) layernorm_gelu_vec_kernel<4, true><<<grid, 256, 0, stream>>>(in, gamma, beta, out, h, l, rows, gelu, g_pram_pred);
                  else layernorm_gelu_vec_kernel<4, false><<<grid, 256, 0, stream>>>(in, gamma, beta, out, h, l, rows, gelu, g_pram_pred);
                  return true;
        case 1024: if (fast) layernorm_gelu_vec_kernel<8, true><<<grid, 256, 0, stream>>>(in, gamma, beta, out, h, l, rows, gelu, g_pram_pred);
                   else layernorm_gelu_vec_kernel<8, false><<<grid, 256, 0, stream>>>(in, gamma, beta, out, h, l, rows, gelu, g_pram_pred);
                   return true;
        default: return false;
    }
}

__global__ void layernorm_gelu_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float* __restrict__ out,
                                      __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                      long long rows, int C, int gelu, const int* __restrict__ pred) {
    long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= rows || pram_pred_skip(pred)) return;
    const float* p = in + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += p[c];
    const float mean = warp_sum(s) / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { float d = p[c] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / C + 1e-5f);
    for (int c = lane; c < C; c += 32) {
        float v = (p[c] - mean) * rstd * gamma[c] + beta[c];
        if (gelu) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        if (out) out[row * C + c] = v;
        if (out_hi) {
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            out_hi[row * C + c] = h;
            if (out_lo) out_lo[row * C + c] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
}

PRAM_API int pram_layernorm_gelu(const float* in, const float* gamma, const float* beta, float* out,
                                 long long rows, int C, int gelu, cudaStream_t stream) {
    if (!in || !out || !gamma || !beta) return PRAM_ERR_ARG;
    if (!layernorm_vec_launch(in, gamma, beta, out, nullptr, nullptr, rows, C, gelu, stream))
        layernorm_gelu_kernel<<<cdiv(rows * 32, 256), 256, 0, stream>>>(in, gamma, beta, out, nullptr, nullptr, rows, C, gelu, g_pram_pred);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

PRAM_API int pram_layernorm_gelu_split(const float* in, const float* gamma, const float* beta, float* out_f32,
                                       void* out_hi, void* out_lo, long long rows, int C, int gelu,
                                       cudaStream_t stream) {
    if (!in || !gamma || !beta || (!out_f32 && !out_hi)) return PRAM_ERR_ARG;
    if (!layernorm_vec_launch(in, gamma, beta, out_f32, out_hi, out_lo, rows, C, gelu, stream))
        layernorm_gelu_kernel<<<cdiv(rows * 32, 256), 256, 0, stream>>>(in, gamma, beta, out_f32, (__nv_bfloat16*)out_hi,
                                                                       (__nv_bfloat16*)out_lo, rows, C, gelu, g_pram_pred);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// ------------------------------------------------------------------------------------------
// qkv [tokens][3*H*64] laid out (q | k | v), each (head, dim) -- the host permutes the reference's
// interleaved (head, dim, {q,k,v}) weight rows once at load time -- to q,k,v [B][H][N][64] with the
// rotary embedding applied to q and k on adjacent pairs (2i,2i+1).
// If cosb == nullptr no rotation is applied (cross attention: qk and v only re-laid out).
// ------------------------------------------------------------------------------------------
__global__ void rotary_split_kernel(const float* __restrict__ qkv, int nparts, int B, int N, int heads,
                                    const float* __restrict__ cosb, const float* __restrict__ sinb,
                                    float scale_qk, float* __restrict__ q, float* __restrict__ k,
                                    float* __restrict__ v) {
    // one thread per (token, part, head, pair)
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int pairs = 32;
    long long total = (long long)B * N * nparts * heads * pairs;
    if (i >= total) return;
    int pr = (int)(i % pairs);
    int h = (int)((i / pairs) % heads);
    int part = (int)((i / (pairs * heads)) % nparts);
    long long t = i / ((long long)pairs * heads * nparts);
    int b = (int)(t / N), n = (int)(t - (long long)b * N);
    const float* src = qkv + t * (long long)(nparts * heads * 64) + (long long)part * heads * 64 + h * 64 + 2 * pr;
    float x0 = src[0], x1 = src[1];
    float* dst = (part == 0 ? q : (part == 1 ? k : v));
    const bool is_v = (nparts == 3 && part == 2) || (nparts == 2 && part == 1);
    if (nparts == 2 && part == 1) dst = v;
    if (!is_v) {
        if (cosb) {
            float c = cosb[t * 32 + pr], s = sinb[t * 32 + pr];
            float y0 = x0 * c + (-x1) * s;
            float y1 = x1 * c + x0 * s;
            x0 = y0; x1 = y1;
        }
        x0 *= scale_qk; x1 *= scale_qk;
    }
    float* o = dst + (((long long)b * heads + h) * N + n) * 64 + 2 * pr;
    o[0] = x0; o[1] = x1;
}

PRAM_API int pram_rotary_split(const float* qkv, int nparts, int B, int N, int heads, const float* cosb,
                               const float* sinb, float scale_qk, float* q, float* k, float* v,
                               cudaStream_t stream) {
    if (!qkv || !q || !v || (nparts != 2 && nparts != 3) || (nparts == 3 && !k)) return PRAM_ERR_ARG;
    long long total = (long long)B * N * nparts * heads * 32;
    rotary_split_kernel<<<cdiv(total, 256), 256, 0, stream>>>(qkv, nparts, B, N, heads, cosb, sinb, scale_qk,
                                                             q, k, v);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// ------------------------------------------------------------------------------------------
// attention: out[b][n][h*64+d] = sum_j softmax_j(q_n . k_j * scale) v_j[d]   (head dim 64)
// CTA = 64 queries x 1 (b,h); K/V streamed through shared memory in 64-row tiles; online softmax.
// thread layout: 256 threads = 64 queries x 4 lanes; each lane owns 16 of the 64 dims.
// Two passes when colmean != nullptr (needs the final row max / sum to form normalised weights).
// ------------------------------------------------------------------------------------------
constexpr int AT_BQ = 64, AT_BK = 32;
__global__ void __launch_bounds__(256) attention_kernel(
    const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, int B, int heads,
    int Nq, int Nk, float scale, float* __restrict__ out, int out_stride, float* __restrict__ colmean /* partials */) {
    __shared__ float Ks[AT_BK][64 + 1];
    __shared__ float Vs[AT_BK][64 + 1];
    __shared__ float Ps[AT_BQ][AT_BK + 1];
    const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads;
    const int q0 = blockIdx.x * AT_BQ;
    const int tid = threadIdx.x;
    const int qi = tid >> 2, part = tid & 3;  // query row in tile, 16-dim slice
    const int qn = q0 + qi;
    const float* Qb = Q + ((long long)bh * Nq) * 64;
    const float* Kb = K + ((long long)bh * Nk) * 64;
    const float* Vb = V + ((long long)bh * Nk) * 64;
    float qreg[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) qreg[d] = (qn < Nq) ? Qb[(long long)qn * 64 + part * 16 + d] * scale : 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    float acc[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) acc[d] = 0.f;
    for (int k0 = 0; k0 < Nk; k0 += AT_BK) {
        __syncthreads();
        for (int i = tid; i < AT_BK * 16; i += 256) {
            int r = i >> 4, c4 = (i & 15) * 4;
            float4 kv = make_float4(0, 0, 0, 0), vv = make_float4(0, 0, 0, 0);
            if (k0 + r < Nk) {
                kv = *reinterpret_cast<const float4*>(Kb + (long long)(k0 + r) * 64 + c4);
                vv = *reinterpret_cast<const float4*>(Vb + (long long)(k0 + r) * 64 + c4);
            }
            Ks[r][c4] = kv.x; Ks[r][c4 + 1] = kv.y; Ks[r][c4 + 2] = kv.z; Ks[r][c4 + 3] = kv.w;
            Vs[r][c4] = vv.x; Vs[r][c4 + 1] = vv.y; Vs[r][c4 + 2] = vv.z; Vs[r][c4 + 3] = vv.w;
        }
        __syncthreads();
        // scores for this query row against 64 keys: each of the 4 lanes does a partial dot, then
        // quad-reduce.  Lane `part` keeps keys j with j%4==part for the softmax bookkeeping.
        float tile_max = -INFINITY;
        for (int j = 0; j < AT_BK; ++j) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < 16; ++d) s = fmaf(qreg[d], Ks[j][part * 16 + d], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (k0 + j >= Nk) s = -INFINITY;
            if ((j & 3) == part) Ps[qi][j] = s;  // the quad shares row qi; lane `part` owns j%4==part
            tile_max = fmaxf(tile_max, s);
        }
        __syncwarp();
        const float m_new = fmaxf(m_run, tile_max);
        const float corr = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
        float psum = 0.f;
#pragma unroll
        for (int jj = 0; jj < AT_BK / 4; ++jj) {
            const float sv = Ps[qi][jj * 4 + part];
            const float pv = (sv == -INFINITY) ? 0.f : expf(sv - m_new);
            Ps[qi][jj * 4 + part] = pv;
            psum += pv;
        }
        psum += __shfl_xor_sync(0xffffffffu, psum, 1);
        psum += __shfl_xor_sync(0xffffffffu, psum, 2);
        l_run = l_run * corr + psum;
        m_run = m_new;
        __syncwarp();
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[d] *= corr;
        for (int j = 0; j < AT_BK; ++j) {
            const float pv = Ps[qi][j];
#pragma unroll
            for (int d = 0; d < 16; ++d) acc[d] = fmaf(pv, Vs[j][part * 16 + d], acc[d]);
        }
    }
    if (qn < Nq) {
        float* o = out + ((long long)b * Nq + qn) * out_stride + h * 64 + part * 16;
        const float inv = 1.f / l_run;
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] = acc[d] * inv;
    }
    if (colmean) {
        // second sweep: normalised attention weights, summed over this tile's queries and scaled by 1/(heads*Nq), written
        // as this CTA's partial row [bh][query tile][Nk]; colmean_reduce_kernel adds the partials of one batch element in a
        // FIXED order (float atomics made AdaGML's pruning decisions near the confidence threshold irreproducible)
        const float inv = (qn < Nq) ? 1.f / l_run : 0.f;
        const float wscale = 1.f / ((float)heads * (float)Nq);
        for (int k0 = 0; k0 < Nk; k0 += AT_BK) {
            __syncthreads();
            for (int i = tid; i < AT_BK * 16; i += 256) {
                int r = i >> 4, c4 = (i & 15) * 4;
                float4 kv = make_float4(0, 0, 0, 0);
                if (k0 + r < Nk) kv = *reinterpret_cast<const float4*>(Kb + (long long)(k0 + r) * 64 + c4);
                Ks[r][c4] = kv.x; Ks[r][c4 + 1] = kv.y; Ks[r][c4 + 2] = kv.z; Ks[r][c4 + 3] = kv.w;
            }
            __syncthreads();
            for (int j = 0; j < AT_BK; ++j) {
                float s = 0.f;
#pragma unroll
                for (int d = 0; d < 16; ++d) s = fmaf(qreg[d], Ks[j][part * 16 + d], s);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (part == 0) Ps[qi][j] = (qn < Nq && k0 + j < Nk) ? expf(s - m_run) * inv : 0.f;
            }
            __syncthreads();
            if (tid < AT_BK && k0 + tid < Nk) {
                float cs = 0.f;
                for (int r = 0; r < AT_BQ; ++r) cs += Ps[r][tid];
                colmean[((long long)bh * gridDim.x + blockIdx.x) * Nk + k0 + tid] = cs * wscale;
            }
        }
    }
}

__global__ void colmean_reduce_kernel(const float* __restrict__ parts, int B, int nparts /* heads * query tiles */, int Nk,
                                      float* __restrict__ colmean) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * Nk) return;
    const int b = (int)(i / Nk), k = (int)(i - (long long)b * Nk);
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += parts[((long long)b * nparts + p) * Nk + k];
    colmean[i] = s;
}

PRAM_API long long pram_attention_f32_colmean_ws_floats(int B, int heads, int Nq, int Nk) {
    return (long long)B * heads * cdiv(Nq, AT_BQ) * Nk;
}

PRAM_API int pram_attention_f32(const float* Q, const float* K, const float* V, int B, int heads, int Nq,
                                int Nk, float scale, float* out, int out_stride, float* colmean, float* colmean_ws,
                                cudaStream_t stream) {
    if (!Q || !K || !V || !out || B <= 0 || Nq <= 0 || Nk <= 0) return PRAM_ERR_ARG;
    if (colmean && !colmean_ws) return PRAM_ERR_WORKSPACE;
    dim3 grid(cdiv(Nq, AT_BQ), B * heads);
    attention_kernel<<<grid, 256, 0, stream>>>(Q, K, V, B, heads, Nq, Nk, scale, out, out_stride, colmean ? colmean_ws : nullptr);
    PRAM_CHECK_LAUNCH();
    if (colmean) {
        colmean_reduce_kernel<<<cdiv((long long)B * Nk, 256), 256, 0, stream>>>(colmean_ws, B, heads * (int)grid.x, Nk, colmean);
        PRAM_CHECK_LAUNCH();
    }
    return PRAM_OK;
}

// ------------------------------------------------------------------------------------------
// conv1a: 3 -> 64 channels, 3x3, BN folded, ReLU (reference nets/sfd2.py:141).  13 FLOP/B, i.e.
// HBM-bound: reads the NCHW fp32 image directly (coalesced along x), keeps the 27x64 weights in shared
// memory (broadcast reads), and writes split-bf16 NHWC output in the 2x2 PHASE-SPLIT layout that lets
// the following stride-2 convolution (conv1b) fetch its taps with unit-stride TMA boxes.
// One thread = one pixel x 64 channels (each output line of 128 B written whole by one thread).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv1a_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                     const float* __restrict__ bias, int B, int H, int W,
                                                     __nv_bfloat16* __restrict__ ps_hi, __nv_bfloat16* __restrict__ ps_lo,
                                                     float* __restrict__ out_f32) {
    __shared__ __align__(16) float ws[27 * 64];
    __shared__ float bs[64];
    for (int i = threadIdx.x; i < 27 * 64; i += 128) ws[i] = w[i];
    if (threadIdx.x < 64) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    // threads beyond the row end compute on zero-padded taps and store nothing (they must reach the barrier)
    float in[27];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int iy = y + r - 1, ix = x + q - 1;
                // tap-major order (r, q, c) to match w[tap][ci][co]
                in[(r * 3 + q) * 3 + c] = (iy >= 0 && iy < H && ix >= 0 && ix < W)
                                              ? __ldg(img + (((long long)b * 3 + c) * H + iy) * W + ix) : 0.f;
            }
    float acc[64];
#pragma unroll
    for (int co = 0; co < 64; ++co) acc[co] = 0.f;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
        const float a = in[t];
#pragma unroll
        for (int co = 0; co < 64; co += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(&ws[t * 64 + co]);
            acc[co] = fmaf(a, wv.x, acc[co]);
            acc[co + 1] = fmaf(a, wv.y, acc[co + 1]);
            acc[co + 2] = fmaf(a, wv.z, acc[co + 2]);
            acc[co + 3] = fmaf(a, wv.w, acc[co + 3]);
        }
    }
#pragma unroll
    for (int co = 0; co < 64; ++co) acc[co] = fmaxf(acc[co] + bs[co], 0.f);
    if (out_f32 && x < W) {
        float* op = out_f32 + (((long long)b * H + y) * W + x) * 64;
#pragma unroll
        for (int co = 0; co < 64; co += 4) *reinterpret_cast<float4*>(op + co) = make_float4(acc[co], acc[co + 1], acc[co + 2], acc[co + 3]);
    }
    if (ps_hi) {
        // stage the 128 pixels x 64 channels of this block in shared memory (XOR-swizzled 16-byte chunks) so
        // that 8 lanes write one pixel's 128-byte line together instead of one thread writing it alone
        __shared__ __align__(16) uint4 st_hi[128 * 8];
        __shared__ __align__(16) uint4 st_lo[128 * 8];
        const int pl = threadIdx.x;
#pragma unroll
        for (int co = 0; co < 64; co += 8) {
            __nv_bfloat16 h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                h[j] = __float2bfloat16_rn(acc[co + j]);
                l[j] = __float2bfloat16_rn(acc[co + j] - __bfloat162float(h[j]));
            }
            const int ch = (co >> 3) ^ (pl & 7);
            st_hi[pl * 8 + ch] = *reinterpret_cast<uint4*>(h);
            st_lo[pl * 8 + ch] = *reinterpret_cast<uint4*>(l);
        }
        __syncthreads();
        const int Hp = (H + 1) >> 1, Wp = (W + 1) >> 1;
        const int x_base = blockIdx.x * 128;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int px = it * 16 + (threadIdx.x >> 3), ch = threadIdx.x & 7;
            const int xx = x_base + px;
            if (xx < W) {
                const long long off = ((((long long)(b * 4 + (y & 1) * 2 + (xx & 1))) * Hp + (y >> 1)) * Wp + (xx >> 1)) * 64 + ch * 8;
                *reinterpret_cast<uint4*>(ps_hi + off) = st_hi[px * 8 + (ch ^ (px & 7))];
                if (ps_lo) *reinterpret_cast<uint4*>(ps_lo + off) = st_lo[px * 8 + (ch ^ (px & 7))];
            }
        }
    }
}

PRAM_API int pram_conv1a(const float* img_nchw, const float* w, const float* bias, int B, int H, int W, void* ps_hi,
                         void* ps_lo, float* out_f32, cudaStream_t stream) {
    if (!img_nchw || !w || !bias || (!ps_hi && !out_f32)) return PRAM_ERR_ARG;
    dim3 grid(cdiv(W, 128), H, B);
    conv1a_kernel<<<grid, 128, 0, stream>>>(img_nchw, w, bias, B, H, W, (__nv_bfloat16*)ps_hi, (__nv_bfloat16*)ps_lo, out_f32);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}
