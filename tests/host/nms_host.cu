// Host emulation of the NMS tile kernel: runs the phase functions of pram_b200/csrc/nms_tile.cuh sequentially
// (one "thread" per item, phases in kernel order) so that the exact kernel logic is checked against the oracle on
// CPU (tests/test_nms_host.py).  Test infrastructure only -- never loaded by the product.
#include "../../pram_b200/csrc/nms_tile.cuh"
#include <stdlib.h>
#include <string.h>

template <int R, int TH>
static void run(const float* score, int B, int H, int W, float* out) {
    using G = NmsGeom<R, TH>;
    unsigned char* raw = (unsigned char*)aligned_alloc(16, (G::SMEM + 15) / 16 * 16);
    for (int b = 0; b < B; ++b)
        for (int by = 0; by < (H + G::TH - 1) / G::TH; ++by)
            for (int bx = 0; bx < (W + G::TW - 1) / G::TW; ++bx) {
                memset(raw, 0xCD, G::SMEM);  // garbage: every byte that is read must have been written by a phase
                NmsTile t;
                t.S = (float*)raw;
                t.T = t.S + G::S_FLOATS;
                t.keep = (unsigned char*)(t.T + G::T_FLOATS);
                t.supp = t.keep + G::MASK_BYTES;
                t.tmpb = t.supp + G::MASK_BYTES;
                t.x0 = bx * G::TW - G::HALO;
                t.y0 = by * G::TH - G::HALO;
                t.H = H; t.W = W;
                t.score = score + (long long)b * H * W;
                const int n_init = (8 * G::SW / 4 > G::MASK_BYTES / 4) ? 8 * G::SW / 4 : G::MASK_BYTES / 4;
                for (int i = 0; i < n_init; ++i) nms_init<G>(t, i);
                for (int i = 0; i < G::SH * (G::NG + 2); ++i) nms_load<G>(t, i);
                for (int i = 0; i < G::SH * (G::SW / 8); ++i) nms_rowmax<G, R, false>(t, i);
                for (int i = 0; i < (G::SH / 8) * G::NG; ++i) nms_colmax<G, R, false>(t, i);
                for (int round = 0; round < 2; ++round) {
                    for (int i = 0; i < G::SH * G::NG; ++i) nms_dilate_h<G, R>(t, i);
                    for (int i = 0; i < G::SH * (G::GW / 4); ++i) nms_dilate_v<G, R>(t, i);
                    for (int i = 0; i < G::SH * (G::SW / 8); ++i) nms_rowmax<G, R, true>(t, i);
                    for (int i = 0; i < (G::SH / 8) * G::NG; ++i) nms_colmax<G, R, true>(t, i);
                }
                for (int i = 0; i < G::TH * 32; ++i) {
                    int gy, gx;
                    float4 v = nms_result<G>(t, i, gy, gx);
                    if (gy >= H || gx >= W) continue;
                    const float vv[4] = {v.x, v.y, v.z, v.w};
                    for (int x = 0; x < 4; ++x)
                        if (gx + x < W) out[((long long)b * H + gy) * W + gx + x] = vv[x];
                }
            }
    free(raw);
}

template <int R>
static void pick(const float* s, int B, int H, int W, int th, float* out) {
    if (th == 96) run<R, 96>(s, B, H, W, out); else run<R, 24>(s, B, H, W, out);
}

extern "C" int nms_host(const float* score, int B, int H, int W, int radius, int th, float* out) {
    switch (radius) {
        case 0: pick<0>(score, B, H, W, th, out); break;
        case 1: pick<1>(score, B, H, W, th, out); break;
        case 2: pick<2>(score, B, H, W, th, out); break;
        case 3: pick<3>(score, B, H, W, th, out); break;
        case 4: pick<4>(score, B, H, W, th, out); break;
        default: return -1;
    }
    return 0;
}
