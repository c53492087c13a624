// Splat stage 2 — the render kernels (replaces FORWARD::renderCUDA, model/head/localagg/src/forward.cu:35-82
// and the prob variant model/head/localagg_prob/src/forward.cu:35-102).
//
// render_tile_kernel    the default tile kernel for points in canonical voxel order (see its header
//                       comment below); stray points are detected per voxel and re-evaluated exactly.
// render_points_kernel  arbitrary points (N != H*W*D, several per voxel): one thread per point walks
//                       the supertile list with the exact box test.
// The tcgen05 variant of the tile kernel lives in splat_forward_tc.cu.
#include <cstdlib>

#include "splat_tile.cuh"

namespace gf {

extern thread_local cudaEvent_t g_ev_before, g_ev_after;  // measurement hooks (cabi.cu)

// SIMT render kernel: bin = 8 x 4 columns x 16 z, VOX consecutive z voxels per thread with VOX x C
// accumulators in registers; a warp covers 4 x 4 columns x 2*VOX z.  VOX = 4: 128 threads, 128
// registers, 4 CTAs per SM; VOX = 2: 256 threads, fewer registers per thread, more resident warps and
// tighter warp footprints at the price of more per-record overhead.
//   * Phase A resolves the bin's ordered Gaussian list; each entry carries the box clipped to the bin
//     as bit masks (x: 8 bits, y: 4 bits, z: 16 bits) and one "this warp's footprint is touched" bit
//     per warp, so Phase B never unpacks coordinates.
//   * Phase B streams the records (cp.async, double buffered); every warp walks only the records that
//     touch its footprint (ballot -> bit loop), a lane tests its column with one AND and its four
//     voxels with one shift, and the class accumulation runs on packed fp32 pairs (FFMA2).
template <int C, bool PROB, int VOX>
__global__ void __launch_bounds__(512 / VOX, VOX == 4 ? (PROB ? 3 : GF_RENDER_CTAS) : (PROB ? 2 : 3)) render_tile_kernel(const RenderParams p) {
    constexpr int REC = rec_floats(C);
    constexpr int CP2 = (C + 1) / 2;   // packed class pairs
    static_assert(REC == 32, "one record = 128 bytes");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RenderSmem<C, VOX> &sm = *reinterpret_cast<RenderSmem<C, VOX> *>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.d.H, W = p.d.W, D = p.d.D;

    // ---- which voxels are mine -------------------------------------------------------------------
    const int binX0 = blockIdx.z * kBinX, binY0 = blockIdx.y * kBinY, binZ0 = blockIdx.x * kBinZ;
#if GF_TILE_MAP == 1
    const int lx = lane & 7, ly = lane >> 3, lq = warp;                                                    // z group
#else
    const int lx = (warp & 1) * 4 + (lane >> 3), ly = (lane >> 1) & 3, lq = (warp >> 1) * 2 + (lane & 1);  // z group
#endif
    const int X = binX0 + lx, Y = binY0 + ly, Z0 = binZ0 + VOX * lq;
    const bool col_ok = X < H && Y < W;
    const long long n0 = (static_cast<long long>(X) * W + Y) * D + Z0;
    const bool vec_ok = (D & 3) == 0;  // then n0 % VOX == 0 and Z0+VOX-1 < D whenever Z0 < D

    float px[VOX], py[VOX], pz[VOX];
    bool vox_ok[VOX];
#pragma unroll
    for (int v = 0; v < VOX; ++v) {
        vox_ok[v] = col_ok && (Z0 + v) < D;
        px[v] = py[v] = pz[v] = 0.f;
    }
    uint32_t stray = 0;   // bit v: point n0+v does not sit in voxel n0+v
    if (col_ok && Z0 < D) {
        if (vec_ok) {
            // 3*VOX contiguous floats, 8-byte aligned (n0 is even)
            float raw[3 * VOX];
            const float2 *src = reinterpret_cast<const float2 *>(p.pts + 3 * n0);
#pragma unroll
            for (int i = 0; i < 3 * VOX / 2; ++i) {
                const float2 t = __ldg(src + i);
                raw[2 * i] = t.x; raw[2 * i + 1] = t.y;
            }
#pragma unroll
            for (int v = 0; v < VOX; ++v) { px[v] = raw[3 * v]; py[v] = raw[3 * v + 1]; pz[v] = raw[3 * v + 2]; }
        } else {
#pragma unroll
            for (int v = 0; v < VOX; ++v)
                if (vox_ok[v]) {
                    px[v] = __ldg(p.pts + 3 * (n0 + v));
                    py[v] = __ldg(p.pts + 3 * (n0 + v) + 1);
                    pz[v] = __ldg(p.pts + 3 * (n0 + v) + 2);
                }
        }
        // canonical-order check: point n must lie in voxel n.  A cheap reciprocal estimate settles every
        // point that is not within 1e-3 cells of a voxel face; only those pay the exact IEEE divisions.
        const float inv = __frcp_rn(p.d.grid_size);
#pragma unroll
        for (int v = 0; v < VOX; ++v)
            if (vox_ok[v]) {
                int ix, iy, iz;
                if (p.points_int) {
                    ix = p.points_int[3 * (n0 + v)];
                    iy = p.points_int[3 * (n0 + v) + 1];
                    iz = p.points_int[3 * (n0 + v) + 2];
                } else {
                    const float fx = (px[v] - p.d.pc_min[0]) * inv - static_cast<float>(X);
                    const float fy = (py[v] - p.d.pc_min[1]) * inv - static_cast<float>(Y);
                    const float fz = (pz[v] - p.d.pc_min[2]) * inv - static_cast<float>(Z0 + v);
                    const float lo_m = 1e-3f, hi_m = 1.f - 1e-3f;
                    if (fx > lo_m && fx < hi_m && fy > lo_m && fy < hi_m && fz > lo_m && fz < hi_m) {
                        ix = X; iy = Y; iz = Z0 + v;
                    } else {
                        ix = voxel_coord(px[v], p.d.pc_min[0], p.d.grid_size);
                        iy = voxel_coord(py[v], p.d.pc_min[1], p.d.grid_size);
                        iz = voxel_coord(pz[v], p.d.pc_min[2], p.d.grid_size);
                    }
                }
                if (!(ix == X && iy == Y && iz == Z0 + v)) stray |= 1u << v;
            }
    }
    // do my VOX points share x and y exactly?  (decided once; the branch on it is warp-uniform in practice)
    bool column = true;
#pragma unroll
    for (int v = 1; v < VOX; ++v) column = column && px[v] == px[0] && py[v] == py[0];
    // entry word: x mask [0,8) | y mask [8,12) | z mask [16,32)
    const uint32_t my_xy = (1u << lx) | (1u << (8 + ly));
    const int my_zshift = 16 + VOX * lq;

    float2 acc[VOX][CP2];
    float zsum[VOX], dens[VOX], keep[VOX];
#pragma unroll
    for (int v = 0; v < VOX; ++v) {
#pragma unroll
        for (int c = 0; c < CP2; ++c) acc[v][c] = make_float2(0.f, 0.f);
        zsum[v] = 0.f; dens[v] = 0.f; keep[v] = 1.f;
    }

    // Everything above depends on the caller's inputs only; the records, boxes, lists and the status word
    // come from the two preparation kernels, which this grid may have been launched ahead of.
    pdl_wait();
    if (stray) atomicOr(p.flags, GF_FLAG_GENERIC_PATH);
    // One step of a lane = one (record, my VOX voxels) evaluation, in two stages so that the walker can overlap the
    // loads of the next step with the arithmetic of the current one:
    //   stage_e    exponent and weight of my voxels from the record's geometry chunks; starts the loads of its class chunks
    //   stage_acc  the class accumulation (VOX x C/2 packed FMAs) with the weights / class chunks of the last stage_e
#if GF_TILE_PIPE == 2
    // Branch-free fused step (splat_tile.cuh): weights of the CURRENT hit in wv[], its record in cur_rec.
    float wv[VOX];
    RecView cur_rec;
    cur_rec.addr = 0;
    // exponent + weights of one hit, no branches: COLUMN is decided once per CTA (every thread's points share x, y)
    auto weights_of = [&](auto column_tag, const RecView rec, uint32_t zb, float (&w)[VOX]) {
        constexpr bool COLUMN = decltype(column_tag)::value;
        const float4 g0 = rec.chunk(0), g1 = rec.chunk(1), g2c = rec.chunk(2);
        if constexpr (COLUMN) {
            const float dx = g0.x - px[0], dy = g0.y - py[0];
            float t1 = g1.x * dx;
            t1 = fmaf(g1.w, dy, t1);
            float A = t1 * dx;
            A = fmaf(g1.y * dy, dy, A);
            const float B = fmaf(g2c.x, dy, g2c.y * dx);
#pragma unroll
            for (int v = 0; v < VOX; ++v) {
                const float dz = g0.z - pz[v];
                const float q = fmaf(fmaf(g1.z, dz, B), dz, A);
                const float E = ((zb >> v) & 1u) ? ex2_approx(q) : 0.f;
                w[v] = PROB ? g0.w * E : E;
                if (PROB) { zsum[v] += w[v]; dens[v] += E; keep[v] *= (1.f - E); }
            }
        } else {
#pragma unroll
            for (int v = 0; v < VOX; ++v) {
                const float dx = g0.x - px[v], dy = g0.y - py[v], dz = g0.z - pz[v];
                float t1 = g1.x * dx;
                t1 = fmaf(g1.w, dy, t1);
                t1 = fmaf(g2c.y, dz, t1);
                float t2 = g1.y * dy;
                t2 = fmaf(g2c.x, dz, t2);
                float q = t1 * dx;
                q = fmaf(t2, dy, q);
                q = fmaf(g1.z * dz, dz, q);
                const float E = ((zb >> v) & 1u) ? ex2_approx(q) : 0.f;
                w[v] = PROB ? g0.w * E : E;
                if (PROB) { zsum[v] += w[v]; dens[v] += E; keep[v] *= (1.f - E); }
            }
        }
    };
    auto run_walk = [&](auto column_tag) {
        auto prime = [&](const RecView rec, uint32_t zb, bool) {
            weights_of(column_tag, rec, zb, wv);
            cur_rec = rec;
        };
        auto fused = [&](const RecView next, uint32_t zb_next, bool) {
            float wn[VOX];
            weights_of(column_tag, next, zb_next, wn);          // same basic block as the accumulation below
#pragma unroll
            for (int c4 = 0; c4 < (C + 3) / 4; ++c4) {
                const float4 s4 = cur_rec.chunk(3 + c4);
#pragma unroll
                for (int v = 0; v < VOX; ++v) {
                    const float2 ww = make_float2(wv[v], wv[v]);
                    acc[v][2 * c4] = __ffma2_rn(make_float2(s4.x, s4.y), ww, acc[v][2 * c4]);
                    if (2 * c4 + 1 < CP2) acc[v][2 * c4 + 1] = __ffma2_rn(make_float2(s4.z, s4.w), ww, acc[v][2 * c4 + 1]);
                }
            }
#pragma unroll
            for (int v = 0; v < VOX; ++v) wv[v] = wn[v];
            cur_rec = next;
        };
        walk_tile<C, VOX>(p, sm, binX0, binY0, binZ0, my_xy, my_zshift, prime, fused);
    };
    if (__syncthreads_and(column ? 1 : 0)) run_walk(std::true_type{});
    else run_walk(std::false_type{});
#else
    float wv[VOX];
#if GF_TILE_PIPE
    float4 cls[(C + 3) / 4];
#else
    RecView cur_rec;   // without the pipeline the class chunks are read where they are used (fewer live registers)
    cur_rec.addr = 0;
#endif
    auto stage_e = [&](const float4 g0, const float4 g1, const float4 g2c, const RecView rec, uint32_t zb, bool active) {
        if (!active) return;
#if GF_TILE_PIPE
#pragma unroll
        for (int c4 = 0; c4 < (C + 3) / 4; ++c4) cls[c4] = rec.chunk(3 + c4);
#else
        cur_rec = rec;
#endif
        const float2 g2 = make_float2(g2c.x, g2c.y);
        if (column) {
            // My VOX points share x and y (voxel centres of one z column — every shipped config,
            // dataset/transform_3d.py:484-499 without perturbation): the exponent is a quadratic
            // in dz alone, q = (cc*dz + B)*dz + A, with A and B evaluated once per record.
            const float dx = g0.x - px[0], dy = g0.y - py[0];
            float t1 = g1.x * dx;
            t1 = fmaf(g1.w, dy, t1);
            float A = t1 * dx;
            A = fmaf(g1.y * dy, dy, A);
            const float B = fmaf(g2.x, dy, g2.y * dx);
#pragma unroll
            for (int v = 0; v < VOX; ++v) {
                const float dz = g0.z - pz[v];
                const float q = fmaf(fmaf(g1.z, dz, B), dz, A);
                const float E = ((zb >> v) & 1u) ? ex2_approx(q) : 0.f;
                wv[v] = PROB ? g0.w * E : E;      // base: the opacity is folded into the class vector (pack kernel)
                if (PROB) { zsum[v] += wv[v]; dens[v] += E; keep[v] *= (1.f - E); }
            }
        } else {
            // general points: quadratic form on packed fp32 pairs, voxels (0,1) and (2,3) share each instruction
#pragma unroll
            for (int h2 = 0; h2 < VOX / 2; ++h2) {
                const int v0 = 2 * h2, v1 = v0 + 1;
                const float2 dx = __fadd2_rn(make_float2(g0.x, g0.x), make_float2(-px[v0], -px[v1]));
                const float2 dy = __fadd2_rn(make_float2(g0.y, g0.y), make_float2(-py[v0], -py[v1]));
                const float2 dz = __fadd2_rn(make_float2(g0.z, g0.z), make_float2(-pz[v0], -pz[v1]));
                float2 t1 = __fmul2_rn(make_float2(g1.x, g1.x), dx);
                t1 = __ffma2_rn(make_float2(g1.w, g1.w), dy, t1);
                t1 = __ffma2_rn(make_float2(g2.y, g2.y), dz, t1);
                float2 t2 = __fmul2_rn(make_float2(g1.y, g1.y), dy);
                t2 = __ffma2_rn(make_float2(g2.x, g2.x), dz, t2);
                float2 q = __fmul2_rn(t1, dx);
                q = __ffma2_rn(t2, dy, q);
                q = __ffma2_rn(__fmul2_rn(make_float2(g1.z, g1.z), dz), dz, q);
                const float E0 = ((zb >> v0) & 1u) ? ex2_approx(q.x) : 0.f;
                const float E1 = ((zb >> v1) & 1u) ? ex2_approx(q.y) : 0.f;
                wv[v0] = PROB ? g0.w * E0 : E0;
                wv[v1] = PROB ? g0.w * E1 : E1;
                if (PROB) {
                    zsum[v0] += wv[v0]; dens[v0] += E0; keep[v0] *= (1.f - E0);
                    zsum[v1] += wv[v1]; dens[v1] += E1; keep[v1] *= (1.f - E1);
                }
            }
        }
    };
    auto stage_acc = [&](bool active) {
        if (!active) return;
#pragma unroll
        for (int c4 = 0; c4 < (C + 3) / 4; ++c4) {
#if GF_TILE_PIPE
            const float4 s4 = cls[c4];
#else
            const float4 s4 = cur_rec.chunk(3 + c4);
#endif
#pragma unroll
            for (int v = 0; v < VOX; ++v) {
                const float2 ww = make_float2(wv[v], wv[v]);
                acc[v][2 * c4] = __ffma2_rn(make_float2(s4.x, s4.y), ww, acc[v][2 * c4]);
                if (2 * c4 + 1 < CP2) acc[v][2 * c4 + 1] = __ffma2_rn(make_float2(s4.z, s4.w), ww, acc[v][2 * c4 + 1]);
            }
        }
    };
    walk_tile<C, VOX>(p, sm, binX0, binY0, binZ0, my_xy, my_zshift, stage_e, stage_acc);

#endif

    // ---- epilogue ----------------------------------------------------------------------------------
    if (!(col_ok && Z0 < D)) return;
    float out[VOX][C];
#pragma unroll
    for (int v = 0; v < VOX; ++v)
#pragma unroll
        for (int c = 0; c < C; ++c) out[v][c] = (c & 1) ? acc[v][c >> 1].y : acc[v][c >> 1].x;
    if (PROB) {
#pragma unroll
        for (int v = 0; v < VOX; ++v) {
            if (zsum[v] > 1e-9f) {
#pragma unroll
                for (int c = 0; c < C; ++c) out[v][c] = __fdiv_rn(out[v][c], zsum[v]);
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) out[v][c] = (c < C - 1) ? static_cast<float>(1.0 / (C - 1)) : 0.f;
            }
        }
    }
    if (p.out.argmax) {   // fused occupancy prediction (lowest index on ties)
        uint32_t packed = 0;
#pragma unroll
        for (int v = 0; v < VOX; ++v) packed |= static_cast<uint32_t>(argmax_of<C>(out[v])) << (8 * v);
        if (vec_ok && VOX == 4) {
            *reinterpret_cast<uint32_t *>(p.out.argmax + n0) = packed;
        } else {
#pragma unroll
            for (int v = 0; v < VOX; ++v)
                if (vox_ok[v]) p.out.argmax[n0 + v] = static_cast<uint8_t>(packed >> (8 * v));
        }
    }
    float *dst = p.out.logits + n0 * C;
    if (vec_ok) {
        float flat[VOX * C];
#pragma unroll
        for (int v = 0; v < VOX; ++v)
#pragma unroll
            for (int c = 0; c < C; ++c) flat[v * C + c] = out[v][c];
        if constexpr ((VOX * C) % 4 == 0) {      // n0 * C * 4 bytes is then a multiple of 16
#pragma unroll
            for (int i = 0; i < VOX * C / 4; ++i)
                __stcs(reinterpret_cast<float4 *>(dst) + i,
                       make_float4(flat[4 * i], flat[4 * i + 1], flat[4 * i + 2], flat[4 * i + 3]));
        } else {                                  // n0 is even, so rows start 8-byte aligned
#pragma unroll
            for (int i = 0; i < VOX * C / 2; ++i)
                __stcs(reinterpret_cast<float2 *>(dst) + i, make_float2(flat[2 * i], flat[2 * i + 1]));
        }
        if (PROB) {
#pragma unroll
            for (int v = 0; v < VOX; v += 2) {
                __stcs(reinterpret_cast<float2 *>(p.out.bin_logits + n0 + v), make_float2(1.f - keep[v], 1.f - keep[v + 1]));
                __stcs(reinterpret_cast<float2 *>(p.out.density + n0 + v), make_float2(dens[v], dens[v + 1]));
                __stcs(reinterpret_cast<float2 *>(p.out.probability + n0 + v), make_float2(zsum[v], zsum[v + 1]));
            }
        }
    } else {
#pragma unroll
        for (int v = 0; v < VOX; ++v)
            if (vox_ok[v]) {
#pragma unroll
                for (int c = 0; c < C; ++c) dst[v * C + c] = out[v][c];
                if (PROB) {
                    p.out.bin_logits[n0 + v] = 1.f - keep[v];
                    p.out.density[n0 + v] = dens[v];
                    p.out.probability[n0 + v] = zsum[v];
                }
            }
    }
    // points that are not in canonical voxel order: exact per-point evaluation overwrites their rows
    if (stray) {
#pragma unroll
        for (int v = 0; v < VOX; ++v)
            if (vox_ok[v] && ((stray >> v) & 1u)) render_one_point<C, PROB>(p, n0 + v, px[v], py[v], pz[v]);
    }
}

// ------------------------------------------------------------------------------------------------
// generic path: arbitrary points
// ------------------------------------------------------------------------------------------------
template <int C, bool PROB>
__global__ void __launch_bounds__(256) render_points_kernel(const RenderParams p) {
    if (!(*reinterpret_cast<volatile uint32_t *>(p.flags) & GF_FLAG_GENERIC_PATH)) return;
    for (long long n = blockIdx.x * 256ll + threadIdx.x; n < p.d.N; n += 256ll * gridDim.x)
        render_one_point<C, PROB>(p, n, p.pts[3 * n], p.pts[3 * n + 1], p.pts[3 * n + 2]);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int launch_render_tc(const RenderParams &rp, cudaStream_t stream);  // splat_forward_tc.cu

// Two tile kernels exist for canonical-order points.  The SIMT quad kernel (this file) is the default
// because it is the faster one on B200 today (profiles/README.md); GF_B200_RENDER=tc selects the
// tcgen05 kernel (splat_forward_tc.cu), kept for A/B measurements.  Both evaluate stray points inline.
static bool use_simt_render() {
    static int cached = -1;
    if (cached < 0) {
        const char *e = getenv("GF_B200_RENDER");
        cached = (e && e[0] == 't') ? 0 : 1;
    }
    return cached == 1;
}

#ifdef GF_ENABLE_VOX2
// voxels per thread of the tile kernel: GF_B200_VOX=2|4 overrides the default
static int render_vox() {
    static int cached = 0;
    if (cached == 0) {
        const char *e = getenv("GF_B200_VOX");
        cached = (e && e[0] == '2') ? 2 : GF_RENDER_VOX;
    }
    return cached;
}
#endif

template <int C, bool PROB>
static int launch_render_t(const RenderParams &rp, bool tile_path, int num_sms, cudaStream_t stream) {
    if (tile_path && !use_simt_render()) return launch_render_tc(rp, stream);
    if (tile_path) {
        const int nbx = (rp.d.H + kBinX - 1) / kBinX;
        GF_REQUIRE(rp.nby <= 65535 && nbx <= 65535, GF_ERR_UNSUPPORTED, "splat: grid too large for the render launch");
        const dim3 grid(rp.nzc, rp.nby, nbx);
        if (g_ev_before && g_ev_after) GF_CUDA_TRY(cudaEventRecord(g_ev_before, stream));
#ifdef GF_ENABLE_VOX2   // experiment kept in the source: 2 voxels per thread (measured equal to 4 on B200)
        if (render_vox() == 2)
            render_tile_kernel<C, PROB, 2><<<grid, 256, sizeof(RenderSmem<C, 2>), stream>>>(rp);
        else
#endif
        if (g_ev_before && g_ev_after) {   // kernel timed alone: plain stream order
            render_tile_kernel<C, PROB, 4><<<grid, 128, sizeof(RenderSmem<C, 4>), stream>>>(rp);
        } else {
            GF_CUDA_TRY(launch_chained(render_tile_kernel<C, PROB, 4>, grid, dim3(128), sizeof(RenderSmem<C, 4>), stream, rp));
        }
        GF_CUDA_TRY(cudaGetLastError());
        if (g_ev_before && g_ev_after) GF_CUDA_TRY(cudaEventRecord(g_ev_after, stream));
        return GF_OK;   // stray points were handled inside the tile kernel
    }
    const long long want = (static_cast<long long>(rp.d.N) + 255) / 256;
    const int grid = static_cast<int>(want < 8ll * num_sms ? (want > 0 ? want : 1) : 8ll * num_sms);
    render_points_kernel<C, PROB><<<grid, 256, 0, stream>>>(rp);
    GF_CUDA_TRY(cudaGetLastError());
    return GF_OK;
}

extern const int kSupportedClasses[] = {16, 17, 18, 19, 20};
extern const int kNumSupportedClasses = 5;

int launch_render(const gf_splat_desc &d, const gf_splat_inputs &in, const gf_splat_outputs &out,
                  const SplatWorkspace &ws, bool tile_path, int num_sms, cudaStream_t stream) {
    RenderParams rp;
    rp.d = d;
    rp.pts = in.pts;
    rp.points_int = in.points_int;
    rp.out = out;
    rp.records = ws.records;
    rp.boxes = ws.boxes;
    rp.lists = ws.lists;
    rp.counts = ws.counts;
    rp.flags = ws.flags;
    rp.st = ws.st;
    rp.nsy = ws.nsy;
    rp.nby = (d.W + kBinY - 1) / kBinY;
    rp.nzc = (d.D + kBinZ - 1) / kBinZ;
    const bool prob = d.variant == GF_SPLAT_PROB;
#define GF_CASE(CC)                                                                  \
    case CC:                                                                         \
        return prob ? launch_render_t<CC, true>(rp, tile_path, num_sms, stream)      \
                    : launch_render_t<CC, false>(rp, tile_path, num_sms, stream);
    switch (d.C) {
        GF_CASE(16)
        GF_CASE(17)
        GF_CASE(18)
        GF_CASE(19)
        GF_CASE(20)
        default:
            set_error("splat: class count C=%d is not compiled in (supported: 16..20)", d.C);
            return GF_ERR_UNSUPPORTED;
    }
#undef GF_CASE
}

}  // namespace gf
