// Tile-local exact simple_nms (reference nets/sfd2.py:20-35), written as per-item phase functions so that the
// CUDA kernel (score.cu) and the host emulation used by the CPU tests (tests/nms_host.cu) run the SAME code.
//
//   keep0 = s == maxpool(s);  2x { supp = dilate(keep); rest = supp ? 0 : s;
//                                  keep |= (rest == maxpool(rest)) & ~supp }
//   windows (2R+1)^2, pixels outside the image behave as -inf / false, comparisons are exact.
//
// A TW x TH output tile depends on raw scores within 5R, so the CTA stages a (TW+2*HALO) x (TH+2*HALO) tile in
// shared memory and runs every pool on it.  Max-pools are separable and register-blocked: a thread produces 8
// consecutive outputs from 16 inputs fetched with 16-byte shared-memory loads and a log-step (2,4,8) sliding
// maximum (5.25 max ops per output instead of 2R); keep / suppression masks are 4-pixel nibbles, so the two
// dilations are a handful of byte operations.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#ifndef NMS_HD
#define NMS_HD __host__ __device__ __forceinline__
#endif

template <int R, int TH_>
struct NmsGeom {
    static constexpr int TW = 128, TH = TH_;
    static constexpr int HALO = (5 * R + 3) / 4 * 4;      // >= 5R, multiple of 4 (aligned 16-byte loads)
    static constexpr int SW = TW + 2 * HALO, SH = TH + 2 * HALO;
    static constexpr int SWP = SW + 8;                     // S row stride: 4 columns of -inf on each side
    static constexpr int TROWS = SH + 8;                   // T: 4 rows of -inf above and below
    static constexpr int NG = SW / 4;                      // 4-pixel groups per row
    static constexpr int GW = (NG + 2 + 3) / 4 * 4;        // mask row stride in bytes (1 zero group each side)
    static constexpr int S_FLOATS = SH * SWP, T_FLOATS = TROWS * SW;
    static constexpr int MASK_BYTES = SH * GW;
    static constexpr size_t SMEM = sizeof(float) * (S_FLOATS + T_FLOATS) + 3 * MASK_BYTES;
    static_assert(SW % 8 == 0 && SH % 8 == 0, "tile must be a multiple of the 8-wide register blocks");
};

struct NmsTile {
    float* S;             // [SH][SWP]   raw scores, -inf outside the image and in the pad columns
    float* T;             // [SH+8][SW]  row maxima, -inf in the pad rows
    unsigned char* keep;  // [SH][GW]    nibble per 4 pixels (bit x = pixel 4c+x), real group c at byte c+1
    unsigned char* supp;
    unsigned char* tmpb;
    int x0, y0;           // image coordinates of tile pixel (0,0)
    int H, W;
    const float* score;   // this frame's score map
};

// out[i] = max(v[i+4-R .. i+4+R]), i = 0..7, from 16 consecutive inputs
template <int R>
NMS_HD void nms_slide8(const float* v, float* out) {
    constexpr int b = 4 - R;
    if constexpr (R == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = v[4 + i];
    } else {
        float p2[7 + 2 * R];  // window 2
#pragma unroll
        for (int j = 0; j < 7 + 2 * R; ++j) p2[j] = fmaxf(v[b + j], v[b + j + 1]);
        if constexpr (R == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) out[i] = fmaxf(p2[i], v[b + i + 2]);
        } else {
            float p4[5 + 2 * R];  // window 4
#pragma unroll
            for (int j = 0; j < 5 + 2 * R; ++j) p4[j] = fmaxf(p2[j], p2[j + 2]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if constexpr (R == 2) out[i] = fmaxf(p4[i], v[b + i + 4]);
                else if constexpr (R == 3) out[i] = fmaxf(p4[i], p4[i + 3]);
                else out[i] = fmaxf(fmaxf(p4[i], p4[i + 4]), v[b + i + 8]);
            }
        }
    }
}

// ---- phase 0: stage the tile; item = row * (NG + 2) + (c + 1), c in [-1, NG] (pad groups included) ----
// split into value / store so that the kernel can keep several global loads in flight per thread
template <class G>
NMS_HD float4 nms_load_value(const NmsTile& t, int item) {
    const int row = item / (G::NG + 2), c = item - row * (G::NG + 2) - 1;
    const float NEG = -INFINITY;
    float4 v = make_float4(NEG, NEG, NEG, NEG);
    const int gy = t.y0 + row, gx = t.x0 + 4 * c;
    if (c >= 0 && c < G::NG && gy >= 0 && gy < t.H && gx + 3 >= 0 && gx < t.W) {
        const float* p = t.score + (long long)gy * t.W + gx;
        if (gx >= 0 && gx + 3 < t.W && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            v = *reinterpret_cast<const float4*>(p);
        } else {
            if (gx >= 0 && gx < t.W) v.x = p[0];
            if (gx + 1 >= 0 && gx + 1 < t.W) v.y = p[1];
            if (gx + 2 >= 0 && gx + 2 < t.W) v.z = p[2];
            if (gx + 3 >= 0 && gx + 3 < t.W) v.w = p[3];
        }
    }
    return v;
}
template <class G>
NMS_HD void nms_store_value(const NmsTile& t, int item, float4 v) {
    const int row = item / (G::NG + 2), c1 = item - row * (G::NG + 2);
    *reinterpret_cast<float4*>(t.S + row * G::SWP + 4 * c1) = v;
}
template <class G>
NMS_HD void nms_load(const NmsTile& t, int item) { nms_store_value<G>(t, item, nms_load_value<G>(t, item)); }

// pad rows of T and the mask arrays; item over max(8 * SW / 4, 3 * MASK_BYTES / 4) words
template <class G>
NMS_HD void nms_init(const NmsTile& t, int item) {
    if (item < 8 * G::SW / 4) {
        const int pr = item / (G::SW / 4), c = item - pr * (G::SW / 4);
        const int row = pr < 4 ? pr : G::SH + pr;  // rows 0..3 and SH+4..SH+7
        const float NEG = -INFINITY;
        *reinterpret_cast<float4*>(t.T + row * G::SW + 4 * c) = make_float4(NEG, NEG, NEG, NEG);
    }
    if (item < G::MASK_BYTES / 4) {
        reinterpret_cast<uint32_t*>(t.keep)[item] = 0u;
        reinterpret_cast<uint32_t*>(t.supp)[item] = 0u;
        reinterpret_cast<uint32_t*>(t.tmpb)[item] = 0u;
    }
}

// ---- row maxima of s (REST = false) or of rest = supp ? 0 : s (REST = true); item = row * (SW/8) + seg ----
template <class G, int R, bool REST>
NMS_HD void nms_rowmax(const NmsTile& t, int item) {
    const int row = item / (G::SW / 8), seg = item - row * (G::SW / 8);
    const int lx = 8 * seg;
    float v[16];
    const float4* sp = reinterpret_cast<const float4*>(t.S + row * G::SWP + lx);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 a = sp[q];
        v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
    }
    if (REST) {
        // v[j] <-> tile column lx + j - 4 <-> real group lx/4 - 1 + j/4 <-> mask byte lx/4 + j/4
        const uint16_t* mp = reinterpret_cast<const uint16_t*>(t.supp + row * G::GW + lx / 4);
        const uint32_t m = (uint32_t)mp[0] | ((uint32_t)mp[1] << 16);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const bool sp_ = (m >> (8 * (j >> 2) + (j & 3))) & 1u;
            if (sp_ && v[j] > -INFINITY) v[j] = 0.f;
        }
    }
    float o[8];
    nms_slide8<R>(v, o);
    float4* tp = reinterpret_cast<float4*>(t.T + (row + 4) * G::SW + lx);
    tp[0] = make_float4(o[0], o[1], o[2], o[3]);
    tp[1] = make_float4(o[4], o[5], o[6], o[7]);
}

// ---- column maxima of T + equality test -> keep nibbles; item = rb * NG + c (8 rows x 4 columns) ----
template <class G, int R, bool REST>
NMS_HD void nms_colmax(const NmsTile& t, int item) {
    const int rb = item / G::NG, c = item - rb * G::NG;
    const int ly = 8 * rb;
    float col[4][16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(t.T + (ly + j) * G::SW + 4 * c);
        col[0][j] = a.x; col[1][j] = a.y; col[2][j] = a.z; col[3][j] = a.w;
    }
    float m[4][8];
#pragma unroll
    for (int x = 0; x < 4; ++x) nms_slide8<R>(col[x], m[x]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 s = *reinterpret_cast<const float4*>(t.S + (ly + i) * G::SWP + 4 * (c + 1));
        const float sv[4] = {s.x, s.y, s.z, s.w};
        unsigned char* kp = t.keep + (ly + i) * G::GW + c + 1;
        const unsigned sup = REST ? (unsigned)t.supp[(ly + i) * G::GW + c + 1] : 0u;
        unsigned nib = 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const bool ok = (sv[x] > -INFINITY) && (sv[x] == m[x][i]) && !((sup >> x) & 1u);
            nib |= ok ? (1u << x) : 0u;
        }
        if (REST) { if (nib) *kp = (unsigned char)(*kp | nib); }
        else *kp = (unsigned char)nib;
    }
}

// ---- dilation of keep: horizontal (item = row * NG + c) then vertical (item = row * (GW/4) + word) ----
template <class G, int R>
NMS_HD void nms_dilate_h(const NmsTile& t, int item) {
    const int row = item / G::NG, c = item - row * G::NG;
    const unsigned char* k = t.keep + row * G::GW + c;  // bytes c, c+1, c+2 = real groups c-1, c, c+1
    const unsigned w = (unsigned)k[0] | ((unsigned)k[1] << 4) | ((unsigned)k[2] << 8);
    constexpr unsigned WIN = (1u << (2 * R + 1)) - 1u;
    unsigned nib = 0;
#pragma unroll
    for (int x = 0; x < 4; ++x) nib |= ((w >> (x + 4 - R)) & WIN) ? (1u << x) : 0u;
    t.tmpb[row * G::GW + c + 1] = (unsigned char)nib;
}

template <class G, int R>
NMS_HD void nms_dilate_v(const NmsTile& t, int item) {
    const int row = item / (G::GW / 4), wd = item - row * (G::GW / 4);
    const int a = row - R < 0 ? 0 : row - R, e = row + R > G::SH - 1 ? G::SH - 1 : row + R;
    uint32_t o = 0;
    for (int y = a; y <= e; ++y) o |= reinterpret_cast<const uint32_t*>(t.tmpb + y * G::GW)[wd];
    reinterpret_cast<uint32_t*>(t.supp + row * G::GW)[wd] = o;
}

// ---- epilogue over the inner tile; item = ty * 32 + cg.  Returns the 4 kept values (0 where not kept). ----
template <class G>
NMS_HD float4 nms_result(const NmsTile& t, int item, int& gy, int& gx) {
    const int ty = item >> 5, cg = item & 31;
    const int row = ty + G::HALO, c = cg + G::HALO / 4;
    gy = t.y0 + row;
    gx = t.x0 + 4 * c;
    const unsigned nib = t.keep[row * G::GW + c + 1];
    const float4 s = *reinterpret_cast<const float4*>(t.S + row * G::SWP + 4 * (c + 1));
    return make_float4((nib & 1u) ? s.x : 0.f, (nib & 2u) ? s.y : 0.f, (nib & 4u) ? s.z : 0.f, (nib & 8u) ? s.w : 0.f);
}
