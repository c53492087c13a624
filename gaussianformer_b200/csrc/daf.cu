// Deformable multi-camera / multi-scale aggregation (replaces deformable_aggregation_kernel and
// deformable_aggregation_grad_kernel, model/encoder/gaussian_encoder/ops/src/deformable_aggregation_cuda.cu:125-259).
//
// The reference runs one thread per output scalar: every thread re-reads the sampling location, the
// level table and its weight, and the backward issues one scalar atomic per corner, per weight and
// per location from every channel thread.  Here one WARP owns one sampling point:
//   * lanes span the channel vector in float4 slices, so each of the 4 bilinear corners is one
//     coalesced 16-byte-per-lane read of a 512-byte feature row (channels-last layout);
//   * the camera gate and all bilinear coefficients are warp-uniform;
//   * backward: feature gradients go out as 16-byte vector atomics (one per lane per corner),
//     weight gradients are reduced over the lanes of a group with shuffles and written by one
//     lane, location gradients are reduced over the warp — no scalar atomic storms.
#include "daf_pair.cuh"

namespace gf {

struct DafParams {
    gf_daf_desc d;
    const float *feat;
    const int32_t *shape;
    const int32_t *start;
    const float *loc;
    const float *weights;
    float *out;               // forward
    const float *grad_out;    // backward
    float *grad_feat, *grad_loc, *grad_weights;
};

struct Bilinear {
    float w[4];          // corner weights (tl, tr, bl, br)
    long long row[4];    // feature row of each corner inside the level
    bool ok[4];
    float lh, lw, hh, hw;
};

__device__ __forceinline__ Bilinear bilinear_setup(float lx, float ly, int h, int w) {
    Bilinear b;
    const float y_im = ly * static_cast<float>(h) - 0.5f;
    const float x_im = lx * static_cast<float>(w) - 0.5f;
    const float yf = floorf(y_im), xf = floorf(x_im);
    const int y0 = static_cast<int>(yf), x0 = static_cast<int>(xf);
    b.lh = y_im - yf; b.lw = x_im - xf;
    b.hh = 1.f - b.lh; b.hw = 1.f - b.lw;
    b.ok[0] = y0 >= 0 && x0 >= 0;
    b.ok[1] = y0 >= 0 && x0 + 1 <= w - 1;
    b.ok[2] = y0 + 1 <= h - 1 && x0 >= 0;
    b.ok[3] = y0 + 1 <= h - 1 && x0 + 1 <= w - 1;
    b.row[0] = static_cast<long long>(y0) * w + x0;
    b.row[1] = b.row[0] + 1;
    b.row[2] = b.row[0] + w;
    b.row[3] = b.row[2] + 1;
    b.w[0] = b.hh * b.hw; b.w[1] = b.hh * b.lw; b.w[2] = b.lh * b.hw; b.w[3] = b.lh * b.lw;
    return b;
}

template <int VEC> struct Vec;
template <> struct Vec<4> {
    using T = float4;
    static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ T load(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
    static __device__ __forceinline__ void fma(T &a, float s, const T &v) {
        a.x = fmaf(s, v.x, a.x); a.y = fmaf(s, v.y, a.y); a.z = fmaf(s, v.z, a.z); a.w = fmaf(s, v.w, a.w);
    }
    static __device__ __forceinline__ float dot(const T &a, const T &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
    static __device__ __forceinline__ void store(float *p, const T &v) { *reinterpret_cast<float4 *>(p) = v; }
    static __device__ __forceinline__ void atomic_add_scaled(float *p, float s, const T &v) {
        atomicAdd(reinterpret_cast<float4 *>(p), make_float4(s * v.x, s * v.y, s * v.z, s * v.w));
    }
};
template <> struct Vec<1> {
    using T = float;
    static __device__ __forceinline__ T zero() { return 0.f; }
    static __device__ __forceinline__ T load(const float *p) { return __ldg(p); }
    static __device__ __forceinline__ void fma(T &a, float s, const T &v) { a = fmaf(s, v, a); }
    static __device__ __forceinline__ float dot(const T &a, const T &b) { return a * b; }
    static __device__ __forceinline__ void store(float *p, const T &v) { *p = v; }
    static __device__ __forceinline__ void atomic_add_scaled(float *p, float s, const T &v) { atomicAdd(p, s * v); }
};

// VEC = 4: C % 128 == 0 and (C/Gr)/4 a power of two <= 32 (a group is a run of whole lanes).
// VEC = 1: any C, Gr.
template <int VEC, bool BACKWARD>
__global__ void __launch_bounds__(kDafThreads) daf_kernel(const DafParams p) {
    using V = Vec<VEC>;
    using T = typename V::T;
    const int lane = threadIdx.x & 31;
    const int C = p.d.num_embeds, M = p.d.num_cams, L = p.d.num_scale, Gr = p.d.num_groups, F = p.d.num_feat;
    const int gdim = C / Gr;
    const long long npts = static_cast<long long>(p.d.batch) * p.d.num_pts;
    const long long warps = static_cast<long long>(gridDim.x) * (kDafThreads / 32);
    // level table in shared memory (dynamically indexed registers would spill to local memory)
    __shared__ int lh[kMaxLevels], lw[kMaxLevels], ls[kMaxLevels];
    if (threadIdx.x < L) {
        lh[threadIdx.x] = p.shape[2 * threadIdx.x];
        lw[threadIdx.x] = p.shape[2 * threadIdx.x + 1];
        ls[threadIdx.x] = p.start[threadIdx.x];
    }
    __syncthreads();

    for (long long bp = static_cast<long long>(blockIdx.x) * (kDafThreads / 32) + (threadIdx.x >> 5); bp < npts; bp += warps) {
        const int b = static_cast<int>(bp / p.d.num_pts);
        const float *locp = p.loc + bp * M * 2;  // (x,y) per camera; broadcast loads
        for (int c0 = lane * VEC; c0 < C; c0 += 32 * VEC) {
            const int grp = c0 / gdim;
            T acc = V::zero();
            T gout = V::zero();
            if (BACKWARD) gout = V::load(p.grad_out + bp * C + c0);
            for (int m = 0; m < M; ++m) {
                const float lx = __ldg(locp + 2 * m), ly = __ldg(locp + 2 * m + 1);
                if (!(lx > 0.f && lx < 1.f && ly > 0.f && ly < 1.f)) continue;  // warp-uniform
                const float *fcam = p.feat + (static_cast<long long>(b) * M + m) * F * C + c0;
                const long long wbase = (bp * M + m) * static_cast<long long>(L) * Gr + grp;
                float gx = 0.f, gy = 0.f;
#pragma unroll 4
                for (int l = 0; l < L; ++l) {
                    const Bilinear bl = bilinear_setup(lx, ly, lh[l], lw[l]);
                    const float *flev = fcam + static_cast<long long>(ls[l]) * C;
                    T v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = bl.ok[k] ? V::load(flev + bl.row[k] * C) : V::zero();
                    const float wt = __ldg(p.weights + wbase + static_cast<long long>(l) * Gr);
                    if (!BACKWARD) {
                        T val = V::zero();
#pragma unroll
                        for (int k = 0; k < 4; ++k) V::fma(val, bl.w[k], v[k]);
                        V::fma(acc, wt, val);
                    } else {
                        // d(out)/d(feat corner) = bilinear weight * aggregation weight
                        float *gf = p.grad_feat + (fcam - p.feat) + static_cast<long long>(ls[l]) * C;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (bl.ok[k]) V::atomic_add_scaled(gf + bl.row[k] * C, bl.w[k] * wt, gout);
                        // d(out)/d(weight) = sampled value
                        T val = V::zero();
#pragma unroll
                        for (int k = 0; k < 4; ++k) V::fma(val, bl.w[k], v[k]);
                        float gw = V::dot(gout, val);
                        // d(out)/d(x_im), d(out)/d(y_im)   (cuda.cu:85-121)
                        T dvx = V::zero(), dvy = V::zero();
                        V::fma(dvx, -bl.hh, v[0]); V::fma(dvx, bl.hh, v[1]); V::fma(dvx, -bl.lh, v[2]); V::fma(dvx, bl.lh, v[3]);
                        V::fma(dvy, -bl.hw, v[0]); V::fma(dvy, -bl.lw, v[1]); V::fma(dvy, bl.hw, v[2]); V::fma(dvy, bl.lw, v[3]);
                        gx = fmaf(static_cast<float>(lw[l]) * wt, V::dot(gout, dvx), gx);
                        gy = fmaf(static_cast<float>(lh[l]) * wt, V::dot(gout, dvy), gy);
                        if (VEC == 4) {
                            const int lanes_per_group = gdim / 4;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1)
                                if (o < lanes_per_group) gw += __shfl_xor_sync(0xffffffffu, gw, o);
                            if ((lane & (lanes_per_group - 1)) == 0) p.grad_weights[wbase + static_cast<long long>(l) * Gr] += gw;
                        } else {
                            atomicAdd(p.grad_weights + wbase + static_cast<long long>(l) * Gr, gw);
                        }
                    }
                }
                if (BACKWARD) {
                    float *gl = p.grad_loc + (bp * M + m) * 2;
                    if (VEC == 4) {  // all 32 lanes are here (C % 128 == 0)
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            gx += __shfl_xor_sync(0xffffffffu, gx, o);
                            gy += __shfl_xor_sync(0xffffffffu, gy, o);
                        }
                        if (lane == 0) {
                            // with several channel slices (C > 128) each slice adds its share in turn
                            gl[0] += gx;
                            gl[1] += gy;
                        }
                    } else {
                        atomicAdd(gl, gx);
                        atomicAdd(gl + 1, gy);
                    }
                }
            }
            if (!BACKWARD) V::store(p.out + bp * C + c0, acc);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fast path (C % 128 == 0, a group = whole lanes, M*L <= 32): one warp per sampling point, and the
// bilinear setup of each (camera, level) pair is computed by ONE lane instead of all 32 — the r01
// profile showed the first-generation kernel to be instruction-bound (224 M warp instructions,
// L2 at 27 %), almost all of it redundant warp-uniform setup arithmetic.
//   lane p < M*L      : pair p = (camera p / L, level p % L): gate, 4 corner rows (clamped to a valid
//                       row), 4 corner weights (zero for corners outside the map) -> per-warp smem
//   ballot of the gate: only visible pairs are visited, in ascending (camera, level) order
//   every visit       : 2-3 uniform LDS.128 for the setup, one weight load, 4 coalesced 16-byte row
//                       loads per lane, 16 FMAs
// ------------------------------------------------------------------------------------------------
template <bool BACKWARD>
__global__ void __launch_bounds__(kDafThreads, BACKWARD ? 3 : 4) daf_fast_kernel(const DafParams p) {
    __shared__ int lh[kMaxLevels], lw[kMaxLevels], ls[kMaxLevels];
    __shared__ __align__(16) PairSetup s_pair[kDafThreads / 32][32];
    if (threadIdx.x < p.d.num_scale) {
        lh[threadIdx.x] = p.shape[2 * threadIdx.x];
        lw[threadIdx.x] = p.shape[2 * threadIdx.x + 1];
        ls[threadIdx.x] = p.start[threadIdx.x];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = p.d.num_embeds, M = p.d.num_cams, L = p.d.num_scale, Gr = p.d.num_groups, F = p.d.num_feat;
    const int npair = M * L;
    const int gdim = C / Gr;
    const long long npts = static_cast<long long>(p.d.batch) * p.d.num_pts;
    const long long warps = static_cast<long long>(gridDim.x) * (kDafThreads / 32);
    PairSetup *mine = s_pair[warp];
    const int my_cam = lane / L, my_lv = lane - my_cam * L;   // lane <-> (camera, level) pair, fixed for the kernel
    const bool small_index = npts < (1ll << 31);

    for (long long bp = static_cast<long long>(blockIdx.x) * (kDafThreads / 32) + warp; bp < npts; bp += warps) {
        const int b = small_index ? static_cast<int>(bp) / p.d.num_pts : static_cast<int>(bp / p.d.num_pts);
        bool gate = false;
        if (lane < npair) {
            PairSetup ps;
            gate = pair_setup(p.d, p.loc, lh, lw, ls, bp, my_cam, my_lv, ps);
            mine[lane] = ps;
        }
        const uint32_t visible = __ballot_sync(0xffffffffu, gate);
        __syncwarp();
        const float *featb = p.feat + static_cast<long long>(b) * M * F * C;
        const float *wpt = p.weights + bp * npair * Gr;
        for (int c0 = lane * 4; c0 < C; c0 += 128) {
            const int grp = c0 / gdim;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 gout = make_float4(0.f, 0.f, 0.f, 0.f);
            if (BACKWARD) gout = __ldg(reinterpret_cast<const float4 *>(p.grad_out + bp * C + c0));
            float gx = 0.f, gy = 0.f;
            uint32_t todo = visible;
            while (todo) {
                const int pr = __ffs(todo) - 1;
                todo &= todo - 1;
                const int4 rows = *reinterpret_cast<const int4 *>(mine[pr].row);      // warp-uniform loads
                const float4 cw = *reinterpret_cast<const float4 *>(mine[pr].w);
                const float wt = __ldg(wpt + pr * Gr + grp);
                const float4 v0 = __ldg(reinterpret_cast<const float4 *>(featb + rows.x + c0));
                const float4 v1 = __ldg(reinterpret_cast<const float4 *>(featb + rows.y + c0));
                const float4 v2 = __ldg(reinterpret_cast<const float4 *>(featb + rows.z + c0));
                const float4 v3 = __ldg(reinterpret_cast<const float4 *>(featb + rows.w + c0));
                if (!BACKWARD) {
                    const float a0 = cw.x * wt, a1 = cw.y * wt, a2 = cw.z * wt, a3 = cw.w * wt;
                    acc.x = fmaf(a0, v0.x, acc.x); acc.y = fmaf(a0, v0.y, acc.y); acc.z = fmaf(a0, v0.z, acc.z); acc.w = fmaf(a0, v0.w, acc.w);
                    acc.x = fmaf(a1, v1.x, acc.x); acc.y = fmaf(a1, v1.y, acc.y); acc.z = fmaf(a1, v1.z, acc.z); acc.w = fmaf(a1, v1.w, acc.w);
                    acc.x = fmaf(a2, v2.x, acc.x); acc.y = fmaf(a2, v2.y, acc.y); acc.z = fmaf(a2, v2.z, acc.z); acc.w = fmaf(a2, v2.w, acc.w);
                    acc.x = fmaf(a3, v3.x, acc.x); acc.y = fmaf(a3, v3.y, acc.y); acc.z = fmaf(a3, v3.z, acc.z); acc.w = fmaf(a3, v3.w, acc.w);
                } else {
                    const float4 fr = *reinterpret_cast<const float4 *>(&mine[pr].lh);
                    const int2 okcam = *reinterpret_cast<const int2 *>(&mine[pr].ok);
                    const float2 frac = make_float2(fr.x, fr.y), dims = make_float2(fr.z, fr.w);
                    const int ok = okcam.x;
                    // d(out)/d(feat corner) = corner weight * aggregation weight, for corners inside the map
                    float *gfb = p.grad_feat + static_cast<long long>(b) * M * F * C + c0;
                    if (ok & 1) atomicAdd(reinterpret_cast<float4 *>(gfb + rows.x), make_float4(cw.x * wt * gout.x, cw.x * wt * gout.y, cw.x * wt * gout.z, cw.x * wt * gout.w));
                    if (ok & 2) atomicAdd(reinterpret_cast<float4 *>(gfb + rows.y), make_float4(cw.y * wt * gout.x, cw.y * wt * gout.y, cw.y * wt * gout.z, cw.y * wt * gout.w));
                    if (ok & 4) atomicAdd(reinterpret_cast<float4 *>(gfb + rows.z), make_float4(cw.z * wt * gout.x, cw.z * wt * gout.y, cw.z * wt * gout.z, cw.z * wt * gout.w));
                    if (ok & 8) atomicAdd(reinterpret_cast<float4 *>(gfb + rows.w), make_float4(cw.w * wt * gout.x, cw.w * wt * gout.y, cw.w * wt * gout.z, cw.w * wt * gout.w));
                    // g . v_k for the four corners (outside corners carry weight 0; their clamped values are masked below)
                    const float m0 = (ok & 1) ? 1.f : 0.f, m1 = (ok & 2) ? 1.f : 0.f, m2 = (ok & 4) ? 1.f : 0.f, m3 = (ok & 8) ? 1.f : 0.f;
                    const float d0 = m0 * (gout.x * v0.x + gout.y * v0.y + gout.z * v0.z + gout.w * v0.w);
                    const float d1 = m1 * (gout.x * v1.x + gout.y * v1.y + gout.z * v1.z + gout.w * v1.w);
                    const float d2 = m2 * (gout.x * v2.x + gout.y * v2.y + gout.z * v2.z + gout.w * v2.w);
                    const float d3 = m3 * (gout.x * v3.x + gout.y * v3.y + gout.z * v3.z + gout.w * v3.w);
                    const float lhh = frac.x, lww = frac.y, hh = 1.f - lhh, hw = 1.f - lww;
                    // d(out)/d(weight) = sampled value . g   (cuda.cu:117-119)
                    float gw = hh * hw * d0 + hh * lww * d1 + lhh * hw * d2 + lhh * lww * d3;
                    // d(val)/d(x_im), d(val)/d(y_im)   (cuda.cu:85-121)
                    gx = fmaf(dims.y * wt, -hh * d0 + hh * d1 - lhh * d2 + lhh * d3, gx);
                    gy = fmaf(dims.x * wt, -hw * d0 - lww * d1 + hw * d2 + lww * d3, gy);
                    const int lanes_per_group = gdim / 4;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                        if (o < lanes_per_group) gw += __shfl_xor_sync(0xffffffffu, gw, o);
                    if ((lane & (lanes_per_group - 1)) == 0) p.grad_weights[(bp * npair + pr) * Gr + grp] += gw;
                    // the location gradient of a camera is complete after its last visible level
                    const int m = okcam.y;
                    const bool cam_done = todo == 0 || mine[__ffs(todo) - 1].cam != m;
                    if (cam_done) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            gx += __shfl_xor_sync(0xffffffffu, gx, o);
                            gy += __shfl_xor_sync(0xffffffffu, gy, o);
                        }
                        if (lane == 0) {
                            float *gl = p.grad_loc + (bp * M + m) * 2;
                            gl[0] += gx;
                            gl[1] += gy;
                        }
                        gx = 0.f; gy = 0.f;
                    }
                }
            }
            if (!BACKWARD) *reinterpret_cast<float4 *>(p.out + bp * C + c0) = acc;
        }
        __syncwarp();   // the next point overwrites this warp's setup slots
    }
}

int launch_daf(const gf_daf_desc &d, const DafParams &dp, bool backward, int num_sms, cudaStream_t stream) {
    const long long npts = static_cast<long long>(d.batch) * d.num_pts;
    if (npts == 0) return GF_OK;
    const int per_cta = kDafThreads / 32;
    long long want = (npts + per_cta - 1) / per_cta;
    const long long cap = static_cast<long long>(num_sms) * 64;
    const int grid = static_cast<int>(want < cap ? want : cap);
    const bool v4 = daf_vec4_ok(d);
    if (backward) {
        if (v4) daf_fast_kernel<true><<<grid, kDafThreads, 0, stream>>>(dp);
        else daf_kernel<1, true><<<grid, kDafThreads, 0, stream>>>(dp);
    } else {
        if (v4) daf_fast_kernel<false><<<grid, kDafThreads, 0, stream>>>(dp);
        else daf_kernel<1, false><<<grid, kDafThreads, 0, stream>>>(dp);
    }
    GF_CUDA_TRY(cudaGetLastError());
    return GF_OK;
}

// ------------------------------------------------------------------------------------------------
// feature_maps_format: [B*M, C, hw_l] maps <-> channels-last table [B*M, F, C], one tiled transpose
// ------------------------------------------------------------------------------------------------
struct FormatParams {
    float *maps[GF_DAF_MAX_LEVELS];
    float *table;
    int hw[GF_DAF_MAX_LEVELS], start[GF_DAF_MAX_LEVELS], tile0[GF_DAF_MAX_LEVELS + 1];   // rows, first row, first tile of a level
    int L, C, F;
};

// 64 (rows of the table) x 64 (channels) tiles through shared memory; both global sides move 16 bytes
// per thread along their contiguous axis (rows of a map, channels of the table).
template <bool INVERSE>
__global__ void __launch_bounds__(256) daf_format_kernel(const FormatParams p) {
    __shared__ float tile[64][65];
    int lv = 0;
    while (lv + 1 < p.L && static_cast<int>(blockIdx.x) >= p.tile0[lv + 1]) ++lv;
    const int r0 = (blockIdx.x - p.tile0[lv]) * 64, c0 = blockIdx.y * 64;
    const int hw = p.hw[lv];
    const size_t bm = blockIdx.z;
    float *map = p.maps[lv] + bm * p.C * hw;                                // [C][hw]
    float *tab = p.table + (bm * p.F + p.start[lv]) * p.C;                  // [hw][C]
    const int t = threadIdx.x;
    const bool vec_rows = (hw & 3) == 0 && (reinterpret_cast<uintptr_t>(map) & 15) == 0;
    const bool vec_ch = (p.C & 3) == 0 && (reinterpret_cast<uintptr_t>(tab) & 15) == 0;
    if (!INVERSE) {
        // map -> tile[row][channel]
        for (int i = t; i < 64 * 16; i += 256) {
            const int c = i >> 4, r4 = (i & 15) * 4;
            if (c0 + c >= p.C) continue;
            const float *src = map + static_cast<size_t>(c0 + c) * hw + r0 + r4;
            if (vec_rows && r0 + r4 + 3 < hw) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(src));
                tile[r4][c] = v.x; tile[r4 + 1][c] = v.y; tile[r4 + 2][c] = v.z; tile[r4 + 3][c] = v.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (r0 + r4 + k < hw) tile[r4 + k][c] = __ldg(src + k);
            }
        }
        __syncthreads();
        for (int i = t; i < 64 * 16; i += 256) {
            const int r = i >> 4, c4 = (i & 15) * 4;
            if (r0 + r >= hw || c0 + c4 >= p.C) continue;
            float *dst = tab + static_cast<size_t>(r0 + r) * p.C + c0 + c4;
            if (vec_ch && c0 + c4 + 3 < p.C) {
                *reinterpret_cast<float4 *>(dst) = make_float4(tile[r][c4], tile[r][c4 + 1], tile[r][c4 + 2], tile[r][c4 + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c0 + c4 + k < p.C) dst[k] = tile[r][c4 + k];
            }
        }
    } else {
        // table -> tile[row][channel] -> map
        for (int i = t; i < 64 * 16; i += 256) {
            const int r = i >> 4, c4 = (i & 15) * 4;
            if (r0 + r >= hw || c0 + c4 >= p.C) continue;
            const float *src = tab + static_cast<size_t>(r0 + r) * p.C + c0 + c4;
            if (vec_ch && c0 + c4 + 3 < p.C) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(src));
                tile[r][c4] = v.x; tile[r][c4 + 1] = v.y; tile[r][c4 + 2] = v.z; tile[r][c4 + 3] = v.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c0 + c4 + k < p.C) tile[r][c4 + k] = __ldg(src + k);
            }
        }
        __syncthreads();
        for (int i = t; i < 64 * 16; i += 256) {
            const int c = i >> 4, r4 = (i & 15) * 4;
            if (c0 + c >= p.C) continue;
            float *dst = map + static_cast<size_t>(c0 + c) * hw + r0 + r4;
            if (vec_rows && r0 + r4 + 3 < hw) {
                *reinterpret_cast<float4 *>(dst) = make_float4(tile[r4][c], tile[r4 + 1][c], tile[r4 + 2][c], tile[r4 + 3][c]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (r0 + r4 + k < hw) dst[k] = tile[r4 + k][c];
            }
        }
    }
}

int launch_daf_format(const gf_daf_format_desc &d, float *const *maps, float *table, bool inverse, cudaStream_t stream) {
    FormatParams fp{};
    fp.L = d.num_scale;
    fp.C = d.num_embeds;
    fp.table = table;
    int rows = 0, tiles = 0;
    for (int l = 0; l < d.num_scale; ++l) {
        fp.maps[l] = maps[l];
        fp.hw[l] = d.hw[l];
        fp.start[l] = rows;
        fp.tile0[l] = tiles;
        rows += d.hw[l];
        tiles += (d.hw[l] + 63) / 64;
    }
    fp.tile0[d.num_scale] = tiles;
    fp.F = rows;
    if (tiles == 0 || d.batch_cams == 0 || d.num_embeds == 0) return GF_OK;
    const dim3 grid(tiles, (d.num_embeds + 63) / 64, d.batch_cams);
    GF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, GF_ERR_UNSUPPORTED, "daf format: too many channels or cameras");
    if (inverse) daf_format_kernel<true><<<grid, 256, 0, stream>>>(fp);
    else daf_format_kernel<false><<<grid, 256, 0, stream>>>(fp);
    GF_CUDA_TRY(cudaGetLastError());
    return GF_OK;
}

int launch_daf_forward(const gf_daf_desc &d, const float *feat, const int32_t *shape, const int32_t *start,
                       const float *loc, const float *weights, float *out, int num_sms, cudaStream_t stream) {
    DafParams dp{};
    dp.d = d; dp.feat = feat; dp.shape = shape; dp.start = start; dp.loc = loc; dp.weights = weights; dp.out = out;
    return launch_daf(d, dp, false, num_sms, stream);
}

int launch_daf_backward(const gf_daf_desc &d, const float *feat, const int32_t *shape, const int32_t *start,
                        const float *loc, const float *weights, const float *grad_out, float *grad_feat,
                        float *grad_loc, float *grad_weights, int num_sms, cudaStream_t stream) {
    DafParams dp{};
    dp.d = d; dp.feat = feat; dp.shape = shape; dp.start = start; dp.loc = loc; dp.weights = weights;
    dp.grad_out = grad_out; dp.grad_feat = grad_feat; dp.grad_loc = grad_loc; dp.grad_weights = grad_weights;
    return launch_daf(d, dp, true, num_sms, stream);
}

}  // namespace gf
