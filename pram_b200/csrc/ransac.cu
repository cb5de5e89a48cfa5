// K19 of SURVEY.md section 2b: absolute pose from 2D-3D matches -- P3P + (LO-)RANSAC + non-linear refinement,
// batched over frames, all in float64 like the library it replaces.
// The reference calls pycolmap.absolute_pose_estimation (CPU C++; localization/singlemap3d.py:168-175,
// :324, :454; tracker.py:211; pose_estimator.py:213,338,452) once per frame, after a D2H copy of the
// matches.  Here the matches stay on the device:
//   compact_matches : matches0 / keypoints / reference xyz -> ordered list of correspondences per frame
//   hypotheses      : one thread per minimal sample (counter-based RNG), Grunert P3P in registers,
//                     every solution scored against all correspondences (squared reprojection error in
//                     normalised coordinates vs (max_error/f)^2, points behind the camera rejected),
//                     block-level arg-max (inlier count, then residual sum) with warp shuffles
//   finalize        : one CTA per frame: global arg-max, local optimisation (Gauss-Newton on the inlier
//                     set, re-scored), final Cauchy-weighted LM refinement, inlier mask, quaternion
// Parity is UNPINNED against pycolmap (absent): RANSAC is randomised, so tests compare the
// pose with known answers / the CPU restatement within a tolerance and the inlier mask recomputed
// from the final pose.
#include "common.cuh"

namespace rs {

struct Pose { double R[9]; double t[3]; };

__device__ __forceinline__ unsigned int hash32(unsigned int x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// real roots of a4 x^4 + a3 x^3 + a2 x^2 + a1 x + a0 (Ferrari + Newton polish); returns count
__device__ int solve_quartic(double a4, double a3, double a2, double a1, double a0, double* roots) {
    if (fabs(a4) < 1e-14 * (fabs(a3) + fabs(a2) + fabs(a1) + fabs(a0) + 1e-300)) return 0;
    const double b = a3 / a4, c = a2 / a4, d = a1 / a4, e = a0 / a4;
    // depressed quartic y^4 + p y^2 + q y + r, x = y - b/4
    const double b2 = b * b;
    const double p = c - 0.375 * b2;
    const double q = d - 0.5 * b * c + 0.125 * b2 * b;
    const double r = e - 0.25 * b * d + 0.0625 * b2 * c - 0.01171875 * b2 * b2;
    int n = 0;
    double ys[4];
    if (fabs(q) < 1e-14) {
        // biquadratic
        const double disc = p * p - 4 * r;
        if (disc >= 0) {
            const double s = sqrt(disc);
            const double z1 = 0.5 * (-p + s), z2 = 0.5 * (-p - s);
            if (z1 >= 0) { ys[n++] = sqrt(z1); ys[n++] = -sqrt(z1); }
            if (z2 >= 0) { ys[n++] = sqrt(z2); ys[n++] = -sqrt(z2); }
        }
    } else {
        // resolvent cubic z^3 + 2p z^2 + (p^2 - 4r) z - q^2 = 0 has a positive real root
        const double A = 2 * p, Bc = p * p - 4 * r, Cc = -q * q;
        // depressed cubic: z = w - A/3
        const double P = Bc - A * A / 3.0;
        const double Q = 2.0 * A * A * A / 27.0 - A * Bc / 3.0 + Cc;
        double z;
        const double disc = 0.25 * Q * Q + P * P * P / 27.0;
        if (disc >= 0) {
            const double sd = sqrt(disc);
            z = cbrt(-0.5 * Q + sd) + cbrt(-0.5 * Q - sd) - A / 3.0;
        } else {
            const double rr = sqrt(-P * P * P / 27.0);
            const double phi = acos(fmin(1.0, fmax(-1.0, -0.5 * Q / rr)));
            const double mag = 2.0 * sqrt(-P / 3.0);
            // the largest of the three real roots
            z = mag * cos(phi / 3.0) - A / 3.0;
        }
        // polish z on the cubic
        for (int it = 0; it < 3; ++it) {
            const double f = ((z + A) * z + Bc) * z + Cc, df = (3 * z + 2 * A) * z + Bc;
            if (fabs(df) > 1e-300) z -= f / df;
        }
        if (z <= 0) return 0;
        const double s = sqrt(z);
        // y^2 + s y + (p + z - q/s)/2 = 0   and   y^2 - s y + (p + z + q/s)/2 = 0
        const double t1 = 0.5 * (p + z - q / s), t2 = 0.5 * (p + z + q / s);
        double dsc = s * s - 4 * t1;
        if (dsc >= 0) { const double sq = sqrt(dsc); ys[n++] = 0.5 * (-s + sq); ys[n++] = 0.5 * (-s - sq); }
        dsc = s * s - 4 * t2;
        if (dsc >= 0) { const double sq = sqrt(dsc); ys[n++] = 0.5 * (s + sq); ys[n++] = 0.5 * (s - sq); }
    }
    for (int i = 0; i < n; ++i) {
        double x = ys[i] - 0.25 * b;
        for (int it = 0; it < 2; ++it) {  // Newton polish on the original polynomial
            const double f = (((a4 * x + a3) * x + a2) * x + a1) * x + a0;
            const double df = ((4 * a4 * x + 3 * a3) * x + 2 * a2) * x + a1;
            if (fabs(df) > 1e-300) x -= f / df;
        }
        roots[i] = x;
    }
    return n;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// orthonormal frame from 3 points: e1 along P1-P0, e3 normal to the triangle
__device__ bool frame3(const double* P0, const double* P1, const double* P2, double* E /*3x3 columns e1,e2,e3*/) {
    double d1[3] = {P1[0] - P0[0], P1[1] - P0[1], P1[2] - P0[2]};
    double d2[3] = {P2[0] - P0[0], P2[1] - P0[1], P2[2] - P0[2]};
    double n1 = norm3(d1);
    if (n1 < 1e-12) return false;
    double e1[3] = {d1[0] / n1, d1[1] / n1, d1[2] / n1}, e3[3], e2[3];
    cross3(e1, d2, e3);
    double n3 = norm3(e3);
    if (n3 < 1e-12) return false;
    e3[0] /= n3; e3[1] /= n3; e3[2] /= n3;
    cross3(e3, e1, e2);
    for (int i = 0; i < 3; ++i) { E[i * 3 + 0] = e1[i]; E[i * 3 + 1] = e2[i]; E[i * 3 + 2] = e3[i]; }
    return true;
}

// Grunert P3P: x[3][2] normalised image points, X[3][3] world points -> up to 4 poses (X_cam = R X + t)
__device__ int p3p(const double (*x)[2], const double (*X)[3], Pose* out) {
    double f[3][3];
    for (int i = 0; i < 3; ++i) {
        const double n = sqrt(x[i][0] * x[i][0] + x[i][1] * x[i][1] + 1.0);
        f[i][0] = x[i][0] / n; f[i][1] = x[i][1] / n; f[i][2] = 1.0 / n;
    }
    double d12[3] = {X[1][0] - X[2][0], X[1][1] - X[2][1], X[1][2] - X[2][2]};
    double d02[3] = {X[0][0] - X[2][0], X[0][1] - X[2][1], X[0][2] - X[2][2]};
    double d01[3] = {X[0][0] - X[1][0], X[0][1] - X[1][1], X[0][2] - X[1][2]};
    const double a = norm3(d12), b = norm3(d02), c = norm3(d01);
    if (a < 1e-12 || b < 1e-12 || c < 1e-12) return 0;
    const double ca = f[1][0] * f[2][0] + f[1][1] * f[2][1] + f[1][2] * f[2][2];
    const double cb = f[0][0] * f[2][0] + f[0][1] * f[2][1] + f[0][2] * f[2][2];
    const double cg = f[0][0] * f[1][0] + f[0][1] * f[1][1] + f[0][2] * f[1][2];
    const double a2 = a * a, b2 = b * b, c2 = c * c;
    const double q = (a2 - c2) / b2, p = (a2 + c2) / b2;
    const double A4 = (q - 1) * (q - 1) - 4 * c2 / b2 * ca * ca;
    const double A3 = 4 * (q * (1 - q) * cb - (1 - p) * ca * cg + 2 * c2 / b2 * ca * ca * cb);
    const double A2 = 2 * (q * q - 1 + 2 * q * q * cb * cb + 2 * (b2 - c2) / b2 * ca * ca - 4 * p * ca * cb * cg +
                           2 * (b2 - a2) / b2 * cg * cg);
    const double A1 = 4 * (-q * (1 + q) * cb + 2 * a2 / b2 * cg * cg * cb - (1 - p) * ca * cg);
    const double A0 = (1 + q) * (1 + q) - 4 * a2 / b2 * cg * cg;
    double roots[4];
    const int nr = solve_quartic(A4, A3, A2, A1, A0, roots);
    double EX[9];
    if (!frame3(X[0], X[1], X[2], EX)) return 0;
    int ns = 0;
    for (int i = 0; i < nr; ++i) {
        const double v = roots[i];
        if (!(v > 0)) continue;
        const double den = 2 * (cg - v * ca);
        if (fabs(den) < 1e-12) continue;
        const double u = ((-1 + q) * v * v - 2 * q * cb * v + 1 + q) / den;
        if (!(u > 0)) continue;
        const double s1sq = c2 / (1 + u * u - 2 * u * cg);
        if (!(s1sq > 0)) continue;
        const double s1 = sqrt(s1sq), s2 = u * s1, s3 = v * s1;
        double Y0[3] = {s1 * f[0][0], s1 * f[0][1], s1 * f[0][2]};
        double Y1[3] = {s2 * f[1][0], s2 * f[1][1], s2 * f[1][2]};
        double Y2[3] = {s3 * f[2][0], s3 * f[2][1], s3 * f[2][2]};
        double EY[9];
        if (!frame3(Y0, Y1, Y2, EY)) continue;
        Pose& P = out[ns];
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc)
                P.R[r * 3 + cc] = EY[r * 3 + 0] * EX[cc * 3 + 0] + EY[r * 3 + 1] * EX[cc * 3 + 1] + EY[r * 3 + 2] * EX[cc * 3 + 2];
        for (int r = 0; r < 3; ++r)
            P.t[r] = Y0[r] - (P.R[r * 3] * X[0][0] + P.R[r * 3 + 1] * X[0][1] + P.R[r * 3 + 2] * X[0][2]);
        ++ns;
    }
    return ns;
}

__device__ __forceinline__ double reproj_err2(const Pose& P, const double* c /*x,y,X,Y,Z*/) {
    const double X = c[2], Y = c[3], Z = c[4];
    const double zc = P.R[6] * X + P.R[7] * Y + P.R[8] * Z + P.t[2];
    if (!(zc > 1e-12)) return 1e300;
    const double xc = P.R[0] * X + P.R[1] * Y + P.R[2] * Z + P.t[0];
    const double yc = P.R[3] * X + P.R[4] * Y + P.R[5] * Z + P.t[1];
    const double dx = xc / zc - c[0], dy = yc / zc - c[1];
    return dx * dx + dy * dy;
}

}  // namespace rs

// correspondences: corr[B][cap][5] doubles (x_norm, y_norm, X, Y, Z), count[B]
__global__ void compact_matches_kernel(const float* __restrict__ kpts, const long long* __restrict__ matches,
                                       const float* __restrict__ xyz, int n, int nref, double fx, double fy,
                                       double cx, double cy, double pixel_shift, double* __restrict__ corr,
                                       int* __restrict__ src_index, int cap, int* __restrict__ count) {
    // one CTA per frame, ordered compaction (deterministic)
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        long long m = (i < n) ? matches[(long long)b * n + i] : -1;
        const bool ok = (m >= 0 && m < nref);
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int pre = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int w = 0; w < wid; ++w) off += warp_tot[w];
        if (ok) {
            const int pos = off + pre;
            if (pos < cap) {
                double* c = corr + ((long long)b * cap + pos) * 5;
                c[0] = ((double)kpts[((long long)b * n + i) * 2] + pixel_shift - cx) / fx;
                c[1] = ((double)kpts[((long long)b * n + i) * 2 + 1] + pixel_shift - cy) / fy;
                const float* X = xyz + ((long long)b * nref + m) * 3;
                c[2] = X[0]; c[3] = X[1]; c[4] = X[2];
                src_index[(long long)b * cap + pos] = i;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < nw; ++w) t += warp_tot[w]; base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) count[b] = min(base, cap);
}

struct HypBest { int count; double res; rs::Pose pose; };

constexpr int HYP_THREADS = 128;           // hypotheses per block
constexpr int HYP_LANES = 4;               // threads per hypothesis: one per P3P solution
constexpr int HYP_BLOCK = HYP_THREADS * HYP_LANES;
constexpr int HYP_CHUNK = 1024;  // correspondences staged per pass (20 KB of shared memory)

// One hypothesis per QUAD of threads (hq = thread / 4): the quad leader draws the minimal sample and solves P3P in float64,
// then each lane scores ONE of the up-to-four solutions against all correspondences (the first version looped over the
// solutions in one thread: 7 warps per SM at batch 32, 4 warps on 8 SMs for a single frame, a latency-bound 0.26 / 0.23 ms).
// Per (hypothesis, solution) the arithmetic is unchanged, and the leader picks among its four lanes in solution order with
// the same strict-improvement rule, so the selected pose is the one the single-thread loop selected.
__global__ void __launch_bounds__(HYP_BLOCK) ransac_hyp_kernel(const double* __restrict__ corr, const int* __restrict__ count,
                                                               int cap, double thr2, unsigned int seed,
                                                               HypBest* __restrict__ block_best) {
    const int b = blockIdx.y;
    const int m = count[b];
    const double* C = corr + (long long)b * cap * 5;
    const int hq = threadIdx.x / HYP_LANES, sidx = threadIdx.x % HYP_LANES;
    const int hyp = blockIdx.x * HYP_THREADS + hq;
    const int lane = threadIdx.x & 31, leader = lane & ~(HYP_LANES - 1);
    rs::Pose sol[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        for (int i = 0; i < 9; ++i) sol[q].R[i] = 0;
        sol[q].t[0] = sol[q].t[1] = sol[q].t[2] = 0;
    }
    int ns = 0;
    __shared__ float s_pts[HYP_CHUNK * 5];
    if (m >= 3 && sidx == 0) {
        // three distinct indices from a counter-based hash
        unsigned int h = rs::hash32(seed ^ rs::hash32((unsigned)b * 0x9e3779b9u + (unsigned)hyp));
        int i0 = h % m;
        h = rs::hash32(h + 0x68bc21ebu);
        int i1 = h % (m - 1); if (i1 >= i0) ++i1;
        h = rs::hash32(h + 0x02e5be93u);
        int i2 = h % (m - 2);
        const int lo = min(i0, i1), hi = max(i0, i1);
        if (i2 >= lo) ++i2;
        if (i2 >= hi) ++i2;
        double x[3][2], X[3][3];
        const int idx[3] = {i0, i1, i2};
        for (int k = 0; k < 3; ++k) {
            const double* c = C + (long long)idx[k] * 5;
            x[k][0] = c[0]; x[k][1] = c[1]; X[k][0] = c[2]; X[k][1] = c[3]; X[k][2] = c[4];
        }
        ns = rs::p3p(x, X, sol);
    }
    ns = __shfl_sync(0xffffffffu, ns, leader);
    // this lane's solution as fp32 (R row-major, then t), handed over by the quad leader
    float R[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        R[k] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float v = __shfl_sync(0xffffffffu, (float)(k < 9 ? sol[q].R[k] : sol[q].t[k - 9]), leader);
            if (q == sidx) R[k] = v;
        }
    }
    // Scoring: this lane's solution against all correspondences.  The correspondences are staged in shared memory as
    // fp32 (one broadcast read feeds every thread of the block) and the inlier test is division-free:
    //   |x_c/z_c - x|^2 <= thr^2   <=>   (x_c - x z_c)^2 + (y_c - y z_c)^2 <= thr^2 z_c^2,   z_c > 0
    // fp32 only RANKS hypotheses; the winner is re-scored, locally optimised and refined in fp64 by the finalize
    // kernel, which also produces the inlier mask.
    int cnt = 0;
    float res = 0.f;
    const float thr2f = (float)thr2;
    const bool live = sidx < ns;
    for (int c0 = 0; c0 < m; c0 += HYP_CHUNK) {
        const int nc = min(HYP_CHUNK, m - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nc * 5; i += HYP_BLOCK) s_pts[i] = (float)C[(long long)c0 * 5 + i];
        __syncthreads();
        if (live) {
            int c = 0;
            float rs_ = 0.f;
            for (int i = 0; i < nc; ++i) {
                const float x = s_pts[5 * i], y = s_pts[5 * i + 1], X = s_pts[5 * i + 2], Y = s_pts[5 * i + 3], Z = s_pts[5 * i + 4];
                const float zc = fmaf(R[6], X, fmaf(R[7], Y, fmaf(R[8], Z, R[11])));
                const float xc = fmaf(R[0], X, fmaf(R[1], Y, fmaf(R[2], Z, R[9])));
                const float yc = fmaf(R[3], X, fmaf(R[4], Y, fmaf(R[5], Z, R[10])));
                const float dx = fmaf(-x, zc, xc), dy = fmaf(-y, zc, yc);
                const float num = fmaf(dx, dx, dy * dy), z2 = zc * zc;
                if (zc > 1e-12f && num <= thr2f * z2) { ++c; rs_ += __fdividef(num, z2); }
            }
            cnt += c;
            res += rs_;
        }
    }
    // quad leader: best of its (up to four) solutions, in solution order, strict improvement only
    int best_cnt = -1, best_s = -1;
    double best_res = 1e300;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int cq = __shfl_sync(0xffffffffu, cnt, leader + q);
        const float rq = __shfl_sync(0xffffffffu, res, leader + q);
        if (q < ns && (cq > best_cnt || (cq == best_cnt && (double)rq < best_res))) { best_cnt = cq; best_res = (double)rq; best_s = q; }
    }
    // block arg-max over the hypotheses: (count desc, residual asc, hypothesis index asc) -- deterministic
    __shared__ int s_cnt[HYP_THREADS];
    __shared__ double s_res[HYP_THREADS];
    if (sidx == 0) { s_cnt[hq] = best_cnt; s_res[hq] = best_res; }
    __syncthreads();
    __shared__ int winner;
    if (threadIdx.x == 0) {
        int w = 0;
        for (int i = 1; i < HYP_THREADS; ++i)
            if (s_cnt[i] > s_cnt[w] || (s_cnt[i] == s_cnt[w] && s_res[i] < s_res[w])) w = i;
        winner = w;
    }
    __syncthreads();
    if (hq == winner && sidx == 0) {
        HypBest& o = block_best[(long long)b * gridDim.x + blockIdx.x];
        o.count = best_cnt; o.res = best_res;
        rs::Pose bp;   // no solution at all (m < 3 or a degenerate sample): identity, count -1 as before
        for (int i = 0; i < 9; ++i) bp.R[i] = (i % 4 == 0);
        bp.t[0] = bp.t[1] = bp.t[2] = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q == best_s) bp = sol[q];
        o.pose = bp;
    }
}

// 6x6 symmetric solve (Gaussian elimination with partial pivoting); returns false if singular
__device__ bool solve6(double* H /*36*/, double* g /*6*/, double* x) {
    double A[6][7];
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) A[i][j] = H[i * 6 + j]; A[i][6] = g[i]; }
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        for (int r = c + 1; r < 6; ++r) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (fabs(A[piv][c]) < 1e-300) return false;
        if (piv != c) for (int j = 0; j < 7; ++j) { double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
        for (int r = c + 1; r < 6; ++r) {
            const double f = A[r][c] / A[c][c];
            for (int j = c; j < 7; ++j) A[r][j] -= f * A[c][j];
        }
    }
    for (int r = 5; r >= 0; --r) {
        double s = A[r][6];
        for (int j = r + 1; j < 6; ++j) s -= A[r][j] * x[j];
        x[r] = s / A[r][r];
    }
    return true;
}

constexpr int FIN_THREADS = 256;

// block-wide sum of NV doubles per thread -> result in out[] (shared), valid after return for all threads
template <int NV>
__device__ void block_sum(double (&v)[NV], double* out, double* scratch /*[NV][8]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = v[k];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) scratch[k * 8 + wid] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < FIN_THREADS / 32; ++w) s += scratch[threadIdx.x * 8 + w];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(FIN_THREADS) ransac_finalize_kernel(
    const double* __restrict__ corr, const int* __restrict__ src_index, const int* __restrict__ count, int cap,
    const HypBest* __restrict__ block_best, int nblocks, double thr2, double cauchy_scale2, int lo_iters, int final_iters,
    int min_inliers, int n_out, double* __restrict__ qvec, double* __restrict__ tvec, int* __restrict__ num_inliers,
    unsigned char* __restrict__ inlier_mask, int* __restrict__ success) {
    const int b = blockIdx.x;
    const int m = count[b];
    const double* C = corr + (long long)b * cap * 5;
    __shared__ rs::Pose pose, trial;
    __shared__ double red[32];
    __shared__ double scratch[28 * 8];
    __shared__ int s_best_cnt, s_ok, s_stop;
    __shared__ double s_lambda, s_cost;
    for (int i = threadIdx.x; i < n_out; i += FIN_THREADS) inlier_mask[(long long)b * n_out + i] = 0;
    if (threadIdx.x == 0) {
        int w = 0;
        for (int i = 1; i < nblocks; ++i) {
            const HypBest& a = block_best[(long long)b * nblocks + i];
            const HypBest& c = block_best[(long long)b * nblocks + w];
            if (a.count > c.count || (a.count == c.count && a.res < c.res)) w = i;
        }
        pose = block_best[(long long)b * nblocks + w].pose;
        s_best_cnt = block_best[(long long)b * nblocks + w].count;
        s_ok = (m >= 3 && s_best_cnt >= 3);
        s_lambda = 1e-4;
    }
    __syncthreads();
    if (!s_ok) {
        if (threadIdx.x == 0) {
            success[b] = 0; num_inliers[b] = 0;
            qvec[b * 4] = 1; qvec[b * 4 + 1] = qvec[b * 4 + 2] = qvec[b * 4 + 3] = 0;
            tvec[b * 3] = tvec[b * 3 + 1] = tvec[b * 3 + 2] = 0;
        }
        return;
    }
    // Two refinement phases, each a Levenberg-Marquardt loop over a FIXED correspondence set (the inliers
    // of the pose the phase starts from), steps accepted on cost decrease only:
    //   phase 0  local optimisation: plain L2 on the RANSAC winner's inliers; the result replaces the
    //            winner only if its support (inlier count) does not shrink            (LO-RANSAC)
    //   phase 1  final refinement: Cauchy-weighted, always kept                       (COLMAP RefineAbsolutePose)
    __shared__ rs::Pose ref, cur;
    for (int phase = 0; phase < 2; ++phase) {
        const bool robust = (phase == 1);
        const int iters = robust ? final_iters : lo_iters;
        if (threadIdx.x == 0) { ref = pose; cur = pose; s_lambda = 1e-4; s_stop = 0; }
        __syncthreads();
        for (int it = 0; it < iters; ++it) {
            double acc[28];  // 21 upper-triangular H entries, 6 gradient entries, cost
#pragma unroll
            for (int k = 0; k < 28; ++k) acc[k] = 0;
            for (int i = threadIdx.x; i < m; i += FIN_THREADS) {
                const double* c = C + (long long)i * 5;
                if (!(rs::reproj_err2(ref, c) <= thr2)) continue;  // fixed set
                const double X = c[2], Y = c[3], Z = c[4];
                const double xr = cur.R[0] * X + cur.R[1] * Y + cur.R[2] * Z;
                const double yr = cur.R[3] * X + cur.R[4] * Y + cur.R[5] * Z;
                const double zr = cur.R[6] * X + cur.R[7] * Y + cur.R[8] * Z;
                const double xc = xr + cur.t[0], yc = yr + cur.t[1], zc = fmax(zr + cur.t[2], 1e-9);
                const double iz = 1.0 / zc;
                const double rx = xc * iz - c[0], ry = yc * iz - c[1];
                const double e = rx * rx + ry * ry;
                const double w = robust ? 1.0 / (1.0 + e / cauchy_scale2) : 1.0;
                const double dx[3] = {iz, 0, -xc * iz * iz}, dy[3] = {0, iz, -yc * iz * iz};
                const double Xr[3] = {xr, yr, zr};
                double Jx[6], Jy[6];
                // d(Xc)/d(omega) = -[Xr]_x  ->  row = Xr x d
                Jx[0] = Xr[1] * dx[2] - Xr[2] * dx[1]; Jx[1] = Xr[2] * dx[0] - Xr[0] * dx[2]; Jx[2] = Xr[0] * dx[1] - Xr[1] * dx[0];
                Jy[0] = Xr[1] * dy[2] - Xr[2] * dy[1]; Jy[1] = Xr[2] * dy[0] - Xr[0] * dy[2]; Jy[2] = Xr[0] * dy[1] - Xr[1] * dy[0];
                Jx[3] = dx[0]; Jx[4] = dx[1]; Jx[5] = dx[2];
                Jy[3] = dy[0]; Jy[4] = dy[1]; Jy[5] = dy[2];
                int k = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int cc = r; cc < 6; ++cc) acc[k++] += w * (Jx[r] * Jx[cc] + Jy[r] * Jy[cc]);
#pragma unroll
                for (int r = 0; r < 6; ++r) acc[21 + r] += w * (Jx[r] * rx + Jy[r] * ry);
                acc[27] += robust ? cauchy_scale2 * log1p(e / cauchy_scale2) : e;
            }
            block_sum<28>(acc, red, scratch);
            if (threadIdx.x == 0) {
                double H[36], g[6], d[6];
                int k = 0;
                for (int r = 0; r < 6; ++r)
                    for (int cc = r; cc < 6; ++cc) { H[r * 6 + cc] = red[k]; H[cc * 6 + r] = red[k]; ++k; }
                for (int r = 0; r < 6; ++r) { g[r] = -red[21 + r]; H[r * 6 + r] += s_lambda * (H[r * 6 + r] + 1e-12); }
                s_cost = red[27];
                trial = cur;
                if (solve6(H, g, d)) {
                    const double th = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                    double dR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
                    if (th > 1e-15) {
                        const double kx = d[0] / th, ky = d[1] / th, kz = d[2] / th;
                        const double sn = sin(th), c1 = 1 - cos(th);
                        const double K[9] = {0, -kz, ky, kz, 0, -kx, -ky, kx, 0};
                        double K2[9];
                        for (int r = 0; r < 3; ++r)
                            for (int cc = 0; cc < 3; ++cc) K2[r * 3 + cc] = K[r * 3] * K[cc] + K[r * 3 + 1] * K[3 + cc] + K[r * 3 + 2] * K[6 + cc];
                        for (int i = 0; i < 9; ++i) dR[i] += sn * K[i] + c1 * K2[i];
                    }
                    for (int r = 0; r < 3; ++r)
                        for (int cc = 0; cc < 3; ++cc)
                            trial.R[r * 3 + cc] = dR[r * 3] * cur.R[cc] + dR[r * 3 + 1] * cur.R[3 + cc] + dR[r * 3 + 2] * cur.R[6 + cc];
                    trial.t[0] = cur.t[0] + d[3]; trial.t[1] = cur.t[1] + d[4]; trial.t[2] = cur.t[2] + d[5];
                    // converged: the step is below what float64 resolves on a pose of unit scale (rotation in radians, translation
                    // in scene units); the remaining iterations of the budget would only re-evaluate the same pose.  A typical
                    // phase stops after 3 - 6 of its 10 / 20 iterations (each one is two block-wide float64 reductions).
                    double dm = 0;
                    for (int r = 0; r < 6; ++r) dm = fmax(dm, fabs(d[r]));
                    const double scale = 1.0 + fmax(fabs(cur.t[0]), fmax(fabs(cur.t[1]), fabs(cur.t[2])));
                    if (dm < 1e-13 * scale) s_stop = 1;
                }
            }
            __syncthreads();
            if (s_stop) break;   // block-uniform (shared flag read after the barrier)
            // cost of the trial pose on the same fixed set
            double ev[1] = {0};
            for (int i = threadIdx.x; i < m; i += FIN_THREADS) {
                const double* c = C + (long long)i * 5;
                if (!(rs::reproj_err2(ref, c) <= thr2)) continue;
                const double X = c[2], Y = c[3], Z = c[4];
                const double zc = fmax(trial.R[6] * X + trial.R[7] * Y + trial.R[8] * Z + trial.t[2], 1e-9);
                const double rx = (trial.R[0] * X + trial.R[1] * Y + trial.R[2] * Z + trial.t[0]) / zc - c[0];
                const double ry = (trial.R[3] * X + trial.R[4] * Y + trial.R[5] * Z + trial.t[1]) / zc - c[1];
                const double e = rx * rx + ry * ry;
                ev[0] += robust ? cauchy_scale2 * log1p(e / cauchy_scale2) : e;
            }
            block_sum<1>(ev, red, scratch);
            if (threadIdx.x == 0) {
                if (red[0] < s_cost) { cur = trial; s_lambda = fmax(s_lambda * 0.3, 1e-12); }
                else s_lambda = fmin(s_lambda * 10.0, 1e8);
            }
            __syncthreads();
        }
        // support of the refined pose
        double sv[1] = {0};
        for (int i = threadIdx.x; i < m; i += FIN_THREADS)
            if (rs::reproj_err2(cur, C + (long long)i * 5) <= thr2) sv[0] += 1.0;
        block_sum<1>(sv, red, scratch);
        if (threadIdx.x == 0) {
            const int cnt = (int)(red[0] + 0.5);
            if (robust || cnt >= s_best_cnt) { pose = cur; s_best_cnt = cnt; }
        }
        __syncthreads();
    }
    // final inlier mask (scattered to the original keypoint order) and outputs
    int cnt = 0;
    for (int i = threadIdx.x; i < m; i += FIN_THREADS) {
        const double e = rs::reproj_err2(pose, C + (long long)i * 5);
        if (e <= thr2) { ++cnt; inlier_mask[(long long)b * n_out + src_index[(long long)b * cap + i]] = 1; }
    }
    double cv[1] = {(double)cnt};
    block_sum<1>(cv, red, scratch);
    if (threadIdx.x == 0) {
        const int total = (int)(red[0] + 0.5);
        num_inliers[b] = total;
        success[b] = total >= min_inliers && total >= 3;
        // rotation matrix -> wxyz quaternion (w >= 0)
        const double* R = pose.R;
        double q[4];
        const double tr = R[0] + R[4] + R[8];
        if (tr > 0) {
            const double s = sqrt(tr + 1.0) * 2; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
        } else if (R[0] > R[4] && R[0] > R[8]) {
            const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
        } else if (R[4] > R[8]) {
            const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
        } else {
            const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
        }
        const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        const double sg = (q[0] < 0 ? -1.0 : 1.0) / nq;
        for (int i = 0; i < 4; ++i) qvec[b * 4 + i] = q[i] * sg;
        for (int i = 0; i < 3; ++i) tvec[b * 3 + i] = pose.t[i];
    }
}

// workspace (bytes): corr B*cap*5 doubles + src_index B*cap ints + count B ints + block_best B*nblocks
PRAM_API long long pram_ransac_workspace_bytes(int B, int cap, int num_hypotheses) {
    const int nblocks = cdiv(num_hypotheses, HYP_THREADS);
    long long bytes = (long long)B * cap * 5 * 8 + (long long)B * cap * 4 + (long long)B * 4;
    bytes = (bytes + 15) / 16 * 16;
    bytes += (long long)B * nblocks * sizeof(HypBest);
    return bytes + 64;
}

__global__ void corr_identity_kernel(int* __restrict__ src_index, int* __restrict__ count, int B, int cap,
                                     const int* __restrict__ counts_in) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * cap) return;
    src_index[i] = (int)(i % cap);
    if (i % cap == 0) { const int b = (int)(i / cap); count[b] = counts_in ? min(max(counts_in[b], 0), cap) : cap; }
}

// Same estimator on correspondences the caller has already normalised in float64: corr [B][n][5] = (x, y, X, Y, Z) with
// (x, y) = camera-plane coordinates of the (undistorted) keypoint and (X, Y, Z) the world point.  This is the entry the
// pycolmap-compatible host call uses (localization/pose_estimator.py): the reference passes float64 numpy arrays and a
// COLMAP camera with distortion parameters (SIMPLE_RADIAL for Aachen), so pixels -> camera plane happens in float64
// on the host and nothing is rounded to float32 on the way in.  counts [B] (NULL = n valid rows each), focal_mean scales
// max_error (pixels) into the camera plane like COLMAP's CamFromImgThreshold.
PRAM_API int pram_ransac_pnp_corr(const double* corr, const int* counts, int B, int n, double focal_mean, double max_error,
                                  int num_hypotheses, int lo_iters, int final_iters, int min_inliers, unsigned int seed,
                                  void* workspace, double* qvec, double* tvec, int* num_inliers, unsigned char* inliers,
                                  int* success, cudaStream_t stream) {
    if (!corr || !workspace || !qvec || !tvec || !num_inliers || !inliers || !success) return PRAM_ERR_ARG;
    if (B <= 0 || n <= 0 || num_hypotheses <= 0 || max_error <= 0 || focal_mean <= 0) return PRAM_ERR_ARG;
    const int cap = n;
    const int nblocks = cdiv(num_hypotheses, HYP_THREADS);
    char* ws = (char*)workspace;
    int* src_index = (int*)(ws + (long long)B * cap * 5 * 8);
    int* count = src_index + (long long)B * cap;
    long long off = (long long)B * cap * 5 * 8 + (long long)B * cap * 4 + (long long)B * 4;
    off = (off + 15) / 16 * 16;
    HypBest* bb = (HypBest*)(ws + off);
    corr_identity_kernel<<<cdiv((long long)B * cap, 256), 256, 0, stream>>>(src_index, count, B, cap, counts);
    PRAM_CHECK_LAUNCH();
    const double thr2 = (max_error / focal_mean) * (max_error / focal_mean);
    dim3 grid(nblocks, B);
    ransac_hyp_kernel<<<grid, HYP_BLOCK, 0, stream>>>(corr, count, cap, thr2, seed, bb);
    PRAM_CHECK_LAUNCH();
    const double cs = 1.0 / focal_mean;
    ransac_finalize_kernel<<<B, FIN_THREADS, 0, stream>>>(corr, src_index, count, cap, bb, nblocks, thr2, cs * cs, lo_iters,
                                                         final_iters, min_inliers, n, qvec, tvec, num_inliers, inliers, success);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}

// kpts [B][n][2] f32 (pixels), matches [B][n] i64 (index into xyz, -1 = none), xyz [B][nref][3] f32.
// outputs: qvec [B][4] (wxyz) f64, tvec [B][3] f64, num_inliers [B] i32, inliers [B][n] u8, success [B] i32
PRAM_API int pram_ransac_pnp(const float* kpts, const long long* matches, const float* xyz, int B, int n, int nref,
                             double fx, double fy, double cx, double cy, double pixel_shift, double max_error,
                             int num_hypotheses, int lo_iters, int final_iters, int min_inliers, unsigned int seed,
                             void* workspace, double* qvec, double* tvec, int* num_inliers, unsigned char* inliers,
                             int* success, cudaStream_t stream) {
    if (!kpts || !matches || !xyz || !workspace || !qvec || !tvec || !num_inliers || !inliers || !success) return PRAM_ERR_ARG;
    if (B <= 0 || n <= 0 || nref <= 0 || num_hypotheses <= 0 || max_error <= 0) return PRAM_ERR_ARG;
    const int cap = n;
    const int nblocks = cdiv(num_hypotheses, HYP_THREADS);
    char* ws = (char*)workspace;
    double* corr = (double*)ws;
    int* src_index = (int*)(ws + (long long)B * cap * 5 * 8);
    int* count = src_index + (long long)B * cap;
    long long off = (long long)B * cap * 5 * 8 + (long long)B * cap * 4 + (long long)B * 4;
    off = (off + 15) / 16 * 16;
    HypBest* bb = (HypBest*)(ws + off);
    compact_matches_kernel<<<B, 256, 0, stream>>>(kpts, matches, xyz, n, nref, fx, fy, cx, cy, pixel_shift, corr, src_index, cap, count);
    PRAM_CHECK_LAUNCH();
    const double f = 0.5 * (fx + fy);
    const double thr2 = (max_error / f) * (max_error / f);
    dim3 grid(nblocks, B);
    ransac_hyp_kernel<<<grid, HYP_BLOCK, 0, stream>>>(corr, count, cap, thr2, seed, bb);
    PRAM_CHECK_LAUNCH();
    const double cs = 1.0 / f;  // Cauchy scale of 1 px in normalised units
    ransac_finalize_kernel<<<B, FIN_THREADS, 0, stream>>>(corr, src_index, count, cap, bb, nblocks, thr2, cs * cs, lo_iters,
                                                         final_iters, min_inliers, n, qvec, tvec, num_inliers, inliers, success);
    PRAM_CHECK_LAUNCH();
    return PRAM_OK;
}
