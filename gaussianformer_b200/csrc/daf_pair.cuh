// Pieces shared by the deformable-aggregation kernels (daf.cu, daf_fused.cu): the per-(camera, level)
// bilinear setup that one lane computes for its whole warp.
#pragma once
#include "common.cuh"

namespace gf {

constexpr int kDafThreads = 256;
constexpr int kMaxLevels = 8;

struct PairSetup {     // 64 bytes per (camera, level) pair
    int row[4];        // element offset of each corner row inside the batch's feature block (row * C)
    float w[4];        // bilinear corner weights, 0 where the corner is outside the map
    float lh, lw;      // fractional parts (backward only)
    float fh, fw;      // level height / width as floats (backward only)
    int ok;            // bit k: corner k lies inside the map
    int cam;           // camera of this pair
    int pad0, pad1;
};

__device__ __forceinline__ bool pair_setup(const gf_daf_desc &d, const float *loc, const int *lh, const int *lw, const int *ls,
                                           long long bp, int m, int lv, PairSetup &o) {
    const int M = d.num_cams, F = d.num_feat, C = d.num_embeds;
    const float lx = __ldg(loc + (bp * M + m) * 2), ly = __ldg(loc + (bp * M + m) * 2 + 1);
    const bool gate = lx > 0.f && lx < 1.f && ly > 0.f && ly < 1.f;
    const int h = lh[lv], w = lw[lv];
    const float y_im = ly * static_cast<float>(h) - 0.5f, x_im = lx * static_cast<float>(w) - 0.5f;
    const float yf = floorf(y_im), xf = floorf(x_im);
    const int y0 = static_cast<int>(yf), x0 = static_cast<int>(xf);
    o.lh = y_im - yf; o.lw = x_im - xf;
    o.fh = static_cast<float>(h); o.fw = static_cast<float>(w);
    const float hh = 1.f - o.lh, hw = 1.f - o.lw;
    const bool oky0 = y0 >= 0, oky1 = y0 + 1 <= h - 1, okx0 = x0 >= 0, okx1 = x0 + 1 <= w - 1;
    const int cy0 = max(y0, 0), cy1 = min(y0 + 1, h - 1), cx0 = max(x0, 0), cx1 = min(x0 + 1, w - 1);
    const int base = m * F + ls[lv];
    o.row[0] = (base + cy0 * w + cx0) * C;
    o.row[1] = (base + cy0 * w + cx1) * C;
    o.row[2] = (base + cy1 * w + cx0) * C;
    o.row[3] = (base + cy1 * w + cx1) * C;
    o.w[0] = (oky0 && okx0) ? hh * hw : 0.f;
    o.w[1] = (oky0 && okx1) ? hh * o.lw : 0.f;
    o.w[2] = (oky1 && okx0) ? o.lh * hw : 0.f;
    o.w[3] = (oky1 && okx1) ? o.lh * o.lw : 0.f;
    o.ok = (oky0 && okx0 ? 1 : 0) | (oky0 && okx1 ? 2 : 0) | (oky1 && okx0 ? 4 : 0) | (oky1 && okx1 ? 8 : 0);
    o.cam = m;
    o.pad0 = o.pad1 = 0;
    return gate;
}

// conditions of the one-warp-per-point fast kernels: C % 128 == 0, a group = a power-of-two run of whole
// lanes, at most 32 (camera, level) pairs, 32-bit row offsets
inline bool daf_vec4_ok(const gf_daf_desc &d) {
    if (d.num_embeds % 128 != 0) return false;
    if (d.num_cams * d.num_scale > 32) return false;
    if (static_cast<long long>(d.num_cams) * d.num_feat * d.num_embeds >= (1ll << 31)) return false;
    const int gdim = d.num_embeds / d.num_groups;
    if (gdim % 4 != 0) return false;
    const int lpg = gdim / 4;
    return lpg <= 32 && (lpg & (lpg - 1)) == 0;
}

}  // namespace gf
