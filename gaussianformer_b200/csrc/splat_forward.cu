// Splat stage 2 — the render kernels (replaces FORWARD::renderCUDA, model/head/localagg/src/forward.cu:35-82
// and the prob variant model/head/localagg_prob/src/forward.cu:35-102).
//
// render_tile_kernel (fast path, points in canonical voxel order):
//   one 128-thread CTA per bin of 8x4 columns x 16 z.  Each thread owns a z-quad (4 consecutive
//   voxels = 288 contiguous output bytes) and keeps 4 x C accumulators in registers.
//   Phase A: the CTA resolves its own Gaussian list from the supertile list (ordered ballot
//            compaction of packed boxes into shared memory; ascending index == reference order).
//   Phase B: records of the listed Gaussians are staged 32 at a time into a double-buffered shared
//            ring by per-record 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx); every
//            lane then reads the record by shared-memory broadcast, applies the exact integer-box
//            test per voxel, evaluates exp2 of the pre-scaled quadratic form and accumulates.
//   The kernel also verifies that its points really are in canonical order; if any thread finds
//   a mismatch it raises GF_FLAG_GENERIC_PATH and the generic kernel (launched right after, a
//   no-op otherwise) recomputes every output.
//
// render_points_kernel (generic path): one thread per point, arbitrary points (several per voxel,
//   N != H*W*D), walks the supertile list with the exact box test.
#include <cstdlib>

#include "splat_render.cuh"

namespace gf {

extern thread_local cudaEvent_t g_ev_before, g_ev_after;  // measurement hooks (cabi.cu)

template <int C>
struct RenderSmem {
    static constexpr int REC = rec_floats(C);
    alignas(128) float stage[2][kChunk * REC];
    alignas(16) uint4 list[kSeg];  // x,y,z packed bounds + Gaussian index
    alignas(8) uint64_t bar[2];
    int warp_count[kRenderThreads / 32];
    int nlist;
};

template <int C, bool PROB>
__global__ void __launch_bounds__(kRenderThreads, PROB ? 3 : 4) render_tile_kernel(const RenderParams p) {
    constexpr int REC = rec_floats(C);
    constexpr int CP = REC - kGeomFloats;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RenderSmem<C> &sm = *reinterpret_cast<RenderSmem<C> *>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.d.H, W = p.d.W, D = p.d.D;

    // ---- which voxels are mine -------------------------------------------------------------------
    const int bin = blockIdx.x / p.nzc, zc = blockIdx.x % p.nzc;
    const int bxi = bin / p.nby, byi = bin % p.nby;
    const int binX0 = bxi * kBinX, binY0 = byi * kBinY, binZ0 = zc * kBinZ;
    const int wX0 = binX0 + (warp & 1) * 4;       // warp footprint: 4 x, 4 y, 8 z
    const int wZ0 = binZ0 + (warp >> 1) * 8;
    const int X = wX0 + (lane >> 3);
    const int Y = binY0 + ((lane >> 1) & 3);
    const int Z0 = wZ0 + (lane & 1) * 4;
    const bool col_ok = X < H && Y < W;
    const long long n0 = (static_cast<long long>(X) * W + Y) * D + Z0;
    const bool vec_ok = (D & 3) == 0;  // then n0 % 4 == 0 and Z0+3 < D whenever Z0 < D

    float px[kVox], py[kVox], pz[kVox];
    bool vox_ok[kVox];
#pragma unroll
    for (int v = 0; v < kVox; ++v) {
        vox_ok[v] = col_ok && (Z0 + v) < D;
        px[v] = py[v] = pz[v] = 0.f;
    }
    if (col_ok && Z0 < D) {
        if (vec_ok) {
            const float4 *src = reinterpret_cast<const float4 *>(p.pts + 3 * n0);
            const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
            px[0] = a.x; py[0] = a.y; pz[0] = a.z; px[1] = a.w;
            py[1] = b.x; pz[1] = b.y; px[2] = b.z; py[2] = b.w;
            pz[2] = c.x; px[3] = c.y; py[3] = c.z; pz[3] = c.w;
        } else {
#pragma unroll
            for (int v = 0; v < kVox; ++v)
                if (vox_ok[v]) {
                    px[v] = __ldg(p.pts + 3 * (n0 + v));
                    py[v] = __ldg(p.pts + 3 * (n0 + v) + 1);
                    pz[v] = __ldg(p.pts + 3 * (n0 + v) + 2);
                }
        }
        // canonical-order check: point n must lie in voxel n
        bool canon = true;
#pragma unroll
        for (int v = 0; v < kVox; ++v)
            if (vox_ok[v]) {
                int ix, iy, iz;
                if (p.points_int) {
                    ix = p.points_int[3 * (n0 + v)];
                    iy = p.points_int[3 * (n0 + v) + 1];
                    iz = p.points_int[3 * (n0 + v) + 2];
                } else {
                    ix = voxel_coord(px[v], p.d.pc_min[0], p.d.grid_size);
                    iy = voxel_coord(py[v], p.d.pc_min[1], p.d.grid_size);
                    iz = voxel_coord(pz[v], p.d.pc_min[2], p.d.grid_size);
                }
                canon = canon && ix == X && iy == Y && iz == Z0 + v;
            }
        if (!canon) atomicOr(p.flags, GF_FLAG_GENERIC_PATH);
    }

    float acc[kVox][C];
    float zsum[kVox], dens[kVox], keep[kVox];
#pragma unroll
    for (int v = 0; v < kVox; ++v) {
#pragma unroll
        for (int c = 0; c < C; ++c) acc[v][c] = 0.f;
        zsum[v] = 0.f; dens[v] = 0.f; keep[v] = 1.f;
    }

    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t use[2] = {0, 0};  // how many times each stage's barrier has completed (parity source)

    // ---- candidates: the ascending list of this bin's supertile ------------------------------------
    const int s = (binX0 / p.st) * p.nsy + (binY0 / p.st);
    const int ncand = p.counts[s];
    const int32_t *cand = p.lists + static_cast<size_t>(s) * p.d.G;
    const uint32_t bX1 = min(binX0 + kBinX, H) - 1, bY1 = min(binY0 + kBinY, W) - 1, bZ1 = min(binZ0 + kBinZ, D) - 1;

    int cpos = 0;
    while (cpos < ncand) {
        // ======================= Phase A: fill sm.list with up to kSeg survivors =====================
        if (tid == 0) sm.nlist = 0;
        __syncthreads();
        int nlist = 0;
        while (cpos < ncand && nlist + kRenderThreads <= kSeg) {
            const int i = cpos + tid;
            uint4 entry = make_uint4(1u, 1u, 1u, 0u);
            bool hit = false;
            if (i < ncand) {
                const int g = __ldg(cand + i);
                const uint4 b = __ldg(reinterpret_cast<const uint4 *>(p.boxes) + g);
                hit = (b.x & 0xffffu) <= bX1 && (b.x >> 16) >= static_cast<uint32_t>(binX0) &&
                      (b.y & 0xffffu) <= bY1 && (b.y >> 16) >= static_cast<uint32_t>(binY0) &&
                      (b.z & 0xffffu) <= bZ1 && (b.z >> 16) >= static_cast<uint32_t>(binZ0) && b.w == 0u;
                entry = make_uint4(b.x, b.y, b.z, static_cast<uint32_t>(g));
            }
            const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) sm.warp_count[warp] = __popc(ballot);
            __syncthreads();
            int off = nlist;
#pragma unroll
            for (int k = 0; k < kRenderThreads / 32; ++k)
                if (k < warp) off += sm.warp_count[k];
            int total = 0;
#pragma unroll
            for (int k = 0; k < kRenderThreads / 32; ++k) total += sm.warp_count[k];
            if (hit) sm.list[off + __popc(ballot & lanemask_lt())] = entry;
            nlist += total;
            cpos += kRenderThreads;
            __syncthreads();
        }

        // ======================= Phase B: stream records and accumulate ==============================
        const int nchunks = (nlist + kChunk - 1) / kChunk;
        auto issue = [&](int k) {  // warp 0 stages chunk k into ring slot k&1
            const int slot = k & 1;
            const int cnt = min(kChunk, nlist - k * kChunk);
            if (lane == 0) mbar_expect_tx(&sm.bar[slot], cnt * REC * 4);
            __syncwarp();
            if (lane < cnt) {
                const uint32_t g = sm.list[k * kChunk + lane].w;
                tma_load_1d(&sm.stage[slot][lane * REC], p.records + static_cast<size_t>(g) * REC, REC * 4,
                            &sm.bar[slot]);
            }
        };
        if (warp == 0 && nchunks > 0) issue(0);
        for (int k = 0; k < nchunks; ++k) {
            const int slot = k & 1;
            if (warp == 0 && k + 1 < nchunks) issue(k + 1);
            mbar_wait(&sm.bar[slot], use[slot] & 1);
            use[slot]++;
            const int cnt = min(kChunk, nlist - k * kChunk);
            for (int j = 0; j < cnt; ++j) {
                const uint4 b = sm.list[k * kChunk + j];
                const int x0 = b.x & 0xffff, x1 = b.x >> 16, z0 = b.z & 0xffff, z1 = b.z >> 16;
                // warp-uniform cull against this warp's 4x4x8 footprint
                if (x1 < wX0 || x0 > wX0 + 3 || z1 < wZ0 || z0 > wZ0 + 7) continue;
                const int y0 = b.y & 0xffff, y1 = b.y >> 16;
                const bool act = X >= x0 && X <= x1 && Y >= y0 && Y <= y1 && z1 >= Z0 && z0 <= Z0 + 3;
                if (act) {
                    const float4 *r4 = reinterpret_cast<const float4 *>(&sm.stage[slot][j * REC]);
                    const float4 g0 = r4[0], g1 = r4[1];
                    const float2 g2 = *reinterpret_cast<const float2 *>(r4 + 2);
                    float wv[kVox];
#pragma unroll
                    for (int v = 0; v < kVox; ++v) {
                        const float dx = g0.x - px[v], dy = g0.y - py[v], dz = g0.z - pz[v];
                        float t1 = g1.x * dx;
                        t1 = fmaf(g1.w, dy, t1);
                        t1 = fmaf(g2.y, dz, t1);
                        float t2 = g1.y * dy;
                        t2 = fmaf(g2.x, dz, t2);
                        float q = t1 * dx;
                        q = fmaf(t2, dy, q);
                        q = fmaf(g1.z * dz, dz, q);
                        const bool in = (Z0 + v) >= z0 && (Z0 + v) <= z1;
                        const float E = in ? ex2_approx(q) : 0.f;
                        wv[v] = g0.w * E;
                        if (PROB) {
                            zsum[v] += wv[v];
                            dens[v] += E;
                            keep[v] *= (1.f - E);
                        }
                    }
#pragma unroll
                    for (int c4 = 0; c4 < CP / 4; ++c4) {
                        const float4 s4 = r4[3 + c4];
                        const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int c = c4 * 4 + i;
                            if (c < C) {
#pragma unroll
                                for (int v = 0; v < kVox; ++v) acc[v][c] = fmaf(sv[i], wv[v], acc[v][c]);
                            }
                        }
                    }
                }
            }
            __syncthreads();  // everyone is done with this ring slot before it is refilled
        }
    }

    // ---- epilogue ----------------------------------------------------------------------------------
    if (!(col_ok && Z0 < D)) return;
    if (PROB) {
#pragma unroll
        for (int v = 0; v < kVox; ++v) {
            if (zsum[v] > 1e-9f) {
                const float inv = 1.f / zsum[v];
#pragma unroll
                for (int c = 0; c < C; ++c) acc[v][c] = __fdiv_rn(acc[v][c], zsum[v]);
                (void)inv;
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) acc[v][c] = (c < C - 1) ? static_cast<float>(1.0 / (C - 1)) : 0.f;
            }
        }
    }
    float *dst = p.out.logits + n0 * C;
    if (vec_ok) {
        float flat[kVox * C];
#pragma unroll
        for (int v = 0; v < kVox; ++v)
#pragma unroll
            for (int c = 0; c < C; ++c) flat[v * C + c] = acc[v][c];
#pragma unroll
        for (int i = 0; i < kVox * C / 4; ++i)
            __stcs(reinterpret_cast<float4 *>(dst) + i,
                   make_float4(flat[4 * i], flat[4 * i + 1], flat[4 * i + 2], flat[4 * i + 3]));
        if (PROB) {
            __stcs(reinterpret_cast<float4 *>(p.out.bin_logits + n0),
                   make_float4(1.f - keep[0], 1.f - keep[1], 1.f - keep[2], 1.f - keep[3]));
            __stcs(reinterpret_cast<float4 *>(p.out.density + n0), make_float4(dens[0], dens[1], dens[2], dens[3]));
            __stcs(reinterpret_cast<float4 *>(p.out.probability + n0),
                   make_float4(zsum[0], zsum[1], zsum[2], zsum[3]));
        }
    } else {
#pragma unroll
        for (int v = 0; v < kVox; ++v)
            if (vox_ok[v]) {
#pragma unroll
                for (int c = 0; c < C; ++c) dst[v * C + c] = acc[v][c];
                if (PROB) {
                    p.out.bin_logits[n0 + v] = 1.f - keep[v];
                    p.out.density[n0 + v] = dens[v];
                    p.out.probability[n0 + v] = zsum[v];
                }
            }
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

// GF_B200_RENDER=simt selects the first-generation SIMT tile kernel (kept for A/B measurements);
// the default is the tcgen05 kernel, which also handles non-canonical points inline.
static bool use_simt_render() {
    static int cached = -1;
    if (cached < 0) {
        const char *e = getenv("GF_B200_RENDER");
        cached = (e && e[0] == 's') ? 1 : 0;
    }
    return cached == 1;
}

template <int C, bool PROB>
static int launch_render_t(const RenderParams &rp, bool tile_path, int num_sms, cudaStream_t stream) {
    if (tile_path && !use_simt_render()) return launch_render_tc(rp, stream);
    if (tile_path) {
        const size_t smem = sizeof(RenderSmem<C>);
        const int nbx = (rp.d.H + kBinX - 1) / kBinX;
        const int grid = nbx * rp.nby * rp.nzc;
        if (g_ev_before && g_ev_after) GF_CUDA_TRY(cudaEventRecord(g_ev_before, stream));
        render_tile_kernel<C, PROB><<<grid, kRenderThreads, smem, stream>>>(rp);
        GF_CUDA_TRY(cudaGetLastError());
        if (g_ev_before && g_ev_after) GF_CUDA_TRY(cudaEventRecord(g_ev_after, stream));
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
