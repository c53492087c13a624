// Splat backward, bin-centric form (replaces BACKWARD::renderCUDA, model/head/localagg/src/backward.cu:24-103 and
// model/head/localagg_prob/src/backward.cu:24-123, for points in canonical voxel order — every shipped config).
//
// The Gaussian-centric kernels (splat_backward.cu) gather one 4C-byte row of the upstream gradient per (Gaussian,
// voxel) pair from L1/L2; ncu shows them waiting on that gather (long scoreboard 3.9 of 8.3 stall cycles per issue, L1
// data pipe 54 %, DRAM 3 %).  Here a CTA owns a bin of 8 x 4 columns x 16 z (the render kernel's bin) and
//
//   * stages the bin's tile of the upstream gradient ([8][4][16][C] fp32 = 36.9 KB for C = 18), its points (6 KB) and,
//     for the prob variant, the per-point terms of prob_aux_kernel (8 KB) into shared memory with TMA tensor copies
//     (cp.async.bulk.tensor.3d, boxes {8C, 4, 8} / {48, 4, 8} / {64, 4, 8} over [B*H][W][D*C]-shaped tensor maps;
//     out-of-grid parts of a box are zero-filled by the hardware) -- every byte of the 46 MB gradient is read from
//     global memory once, coalesced;
//   * resolves the bin's ordered Gaussian list from the supertile list while the copies are in flight (Phase A of
//     the render kernel);
//   * gives every listed Gaussian to EIGHT lanes (four Gaussians per warp at a time; warps take groups of four list
//     entries from a shared counter): the lanes stride over the box clipped to the bin (92 voxels on average for the
//     nuScenes workload -- a whole warp per Gaussian would spend more on its reduction than on the pairs), read their
//     pair's point, gradient row and aux terms from shared memory, accumulate the 28 (+1) sums of GaussAcc, reduce them
//     with a transposing reduction over the eight lanes (28 shuffles for four Gaussians) and add the raw sums to a
//     [G, 32] accumulator with one atomic per value;
//   * bin_finish_kernel applies the per-Gaussian linear maps (finish(), splat_bwd_common.cuh) once per Gaussian.
//
// The accumulator is zero-filled by the host first.  Summation order across bins is not fixed (fp32 atomics), so two
// runs may differ in the last bits; GF_B200_BWD=gauss selects the deterministic Gaussian-centric kernels, which also
// serve every case this kernel does not cover (points not in canonical order, odd strides, misaligned tensors).
#include <cuda.h>
#include <cstdlib>

#include "splat_bwd_common.cuh"

namespace gf {

constexpr int kBinThreads = 256;
constexpr int kBinSeg = 512;      // list entries resolved per segment
// Phase B balance: four list entries share a warp (8 lanes each) and run in lock step, so they should need the same
// number of passes; entries are therefore ordered by key = ceil(volume / 8) with a counting sort, and boxes of
// kWideVol voxels or more (the whole-grid "empty" Gaussian; most Gaussians of the prob config) get a whole warp.
// Host-side model on the nuScenes workload: 667 k -> 372 k warp passes (ideal 339 k).
constexpr int kVolKeys = 64;      // keys 1..64 (a clipped box has at most 512 voxels)
constexpr int kWideVol = 192;

struct alignas(64) BinMaps {
    CUtensorMap grad;   // logits_grad as [B*H][W][D*C]
    CUtensorMap pts;    // pts as [B*H or H][W][D*3]
    CUtensorMap aux;    // prob: aux as [B*H][W][D*4]
};

struct BinParams {
    const float *records;     // raw records of the pack kernel: [B*G, 32]
    float *sums;              // [B, 32, G] raw sums (value-major), zero-filled by the host
    const PackedBox *boxes;
    const int32_t *lists;
    const int32_t *counts;
    int st, nsy, nsuper;
    int nbx;
};

template <int C, bool PROB>
struct BinSmem {
    // two TMA boxes {8C, 4, 8}: z half h at h*256*C, voxel (x, y, z & 7) at ((x*4 + y)*8 + (z & 7))*C
    alignas(128) float grad[512 * C];
    alignas(128) float pts[512 * 3];                 // one box {48, 4, 8}: (x*4 + y)*48 + z*3
    alignas(128) float4 aux[PROB ? 512 : 1];         // one box {64, 4, 8}: (x*4 + y)*16 + z
    alignas(8) uint2 list[kBinSeg];                  // x: box relative to the bin as bit masks, y: Gaussian index
    alignas(8) uint2 sorted[kBinSeg];                // the same entries ordered by clipped-box volume (counting sort)
    alignas(8) uint64_t bar;
    int warp_count[2][kBinThreads / 32];
    int hist[kVolKeys + 1];                          // counting sort: entries per key, then running offsets
    int n_narrow;                                    // entries served by 8-lane groups (the rest: one warp each)
    int next;                                        // next unclaimed ticket of the segment
};

// Sum of x[i] over the eight lanes of a group for all i at once: afterwards the lane with low bits (b2 b1 b0) holds the
// group totals of x[16*b2 + 8*b1 + 4*b0 + i] in x[i], i = 0..3.  28 shuffles.
__device__ __forceinline__ void group8_transpose_reduce(float x[32], int lane) {
#pragma unroll
    for (int st = 0; st < 3; ++st) {
        const int h = 4 >> st, half = 16 >> st;
        const bool upper = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? x[i] : x[i + half];
            const float keep = upper ? x[i + half] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
}

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// floor(i / n) for 0 <= i < 2048 and 1 <= n <= 16 without a division: i * ceil(2^16 / n) >> 16 (exact: i * (ceil - 2^16/n) < 2^16 / n)
__constant__ uint32_t kRcp16[17] = {0u, 65536u, 32768u, 21846u, 16384u, 13108u, 10923u, 9363u, 8192u,
                                    7282u, 6554u, 5958u, 5462u, 5042u, 4682u, 4370u, 4096u};
__device__ __forceinline__ int small_div(int i, int n) {
    return static_cast<int>((static_cast<uint32_t>(i) * kRcp16[n]) >> 16);
}

template <int C, bool PROB>
__global__ void __launch_bounds__(kBinThreads, 2) backward_bin_kernel(const BwdParams pb, const BinParams bp,
                                                                       const __grid_constant__ BinMaps maps) {
    constexpr int NT = kBinThreads, NWARP = NT / 32, CP2 = (C + 1) / 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BinSmem<C, PROB> &sm = *reinterpret_cast<BinSmem<C, PROB> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sample = blockIdx.z / bp.nbx;
    const BwdParams p = sample_bwd(pb, sample);
    const int H = p.d.H, W = p.d.W, D = p.d.D;
    const int binX0 = (blockIdx.z - sample * bp.nbx) * kBinX, binY0 = blockIdx.y * kBinY, binZ0 = blockIdx.x * kBinZ;
    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        mbar_fence_init();
    }
    pdl_launch_dependents();
    __syncthreads();
    pdl_wait();   // boxes / lists of the preparation kernels, and (transitively) the aux terms of the prob variant

    if (tid == 0) {
        constexpr uint32_t kBytes = 512u * C * 4u + 512u * 3u * 4u + (PROB ? 512u * 16u : 0u);
        mbar_expect_tx(&sm.bar, kBytes);
        const int xr = sample * H + binX0;
        tma_load_3d(&sm.grad[0], &maps.grad, binZ0 * C, binY0, xr, &sm.bar);
        tma_load_3d(&sm.grad[256 * C], &maps.grad, binZ0 * C + 8 * C, binY0, xr, &sm.bar);
        tma_load_3d(&sm.pts[0], &maps.pts, binZ0 * 3, binY0, (p.d.pts_shared ? 0 : sample * H) + binX0, &sm.bar);
        if (PROB) tma_load_3d(&sm.aux[0], &maps.aux, binZ0 * 4, binY0, xr, &sm.bar);
    }

    // ---- candidates: the ascending list of this bin's supertile (as walk_tile, splat_tile.cuh) ------------------
    const int st_shift = 31 - __clz(bp.st);
    const int s = sample * bp.nsuper + (binX0 >> st_shift) * bp.nsy + (binY0 >> st_shift);
    const int ncand = bp.counts[s];
    const int32_t *cand = bp.lists + static_cast<size_t>(s) * p.d.G;
    const PackedBox *boxes = bp.boxes + static_cast<size_t>(sample) * p.d.G;
    const uint32_t bX1 = min(binX0 + kBinX, H) - 1, bY1 = min(binY0 + kBinY, W) - 1, bZ1 = min(binZ0 + kBinZ, D) - 1;
    bool tiles_ready = false;

    int cpos = 0;
    while (cpos < ncand) {
        __syncthreads();   // previous segment fully consumed
        if (tid == 0) sm.next = 0;   // published by the barriers of Phase A
        if (tid <= kVolKeys) sm.hist[tid] = 0;
        // ======================= Phase A: ordered survivors of the box test ==========================
        int nlist = 0;
        while (cpos < ncand && nlist + NT <= kBinSeg) {
            constexpr int kPre = 2;   // rounds fetched together (memory-level parallelism)
            int gg[kPre];
            uint4 bb[kPre];
#pragma unroll
            for (int u = 0; u < kPre; ++u) {
                const int i = cpos + u * NT + tid;
                gg[u] = i < ncand ? __ldg(cand + i) : -1;
            }
#pragma unroll
            for (int u = 0; u < kPre; ++u)
                bb[u] = gg[u] >= 0 ? __ldg(reinterpret_cast<const uint4 *>(boxes) + gg[u]) : make_uint4(1u, 1u, 1u, 1u);
#pragma unroll
            for (int u = 0; u < kPre; ++u) {
                if (cpos >= ncand || nlist + NT > kBinSeg) break;   // uniform
                const uint4 b = bb[u];
                const uint32_t x0 = b.x & 0xffffu, x1 = b.x >> 16, y0 = b.y & 0xffffu, y1 = b.y >> 16,
                               z0 = b.z & 0xffffu, z1 = b.z >> 16;
                const bool hit = gg[u] >= 0 && x0 <= bX1 && x1 >= static_cast<uint32_t>(binX0) && y0 <= bY1 &&
                                 y1 >= static_cast<uint32_t>(binY0) && z0 <= bZ1 && z1 >= static_cast<uint32_t>(binZ0) &&
                                 b.w == 0u;
                const int rx0 = max(static_cast<int>(x0) - binX0, 0), rx1 = min(static_cast<int>(x1) - binX0, kBinX - 1);
                const int ry0 = max(static_cast<int>(y0) - binY0, 0), ry1 = min(static_cast<int>(y1) - binY0, kBinY - 1);
                const int rz0 = max(static_cast<int>(z0) - binZ0, 0), rz1 = min(static_cast<int>(z1) - binZ0, kBinZ - 1);
                const uint32_t xm = ((2u << rx1) - 1u) & ~((1u << rx0) - 1u);
                const uint32_t ym = ((2u << ry1) - 1u) & ~((1u << ry0) - 1u);
                const uint32_t zm = ((2u << rz1) - 1u) & ~((1u << rz0) - 1u);
                const uint2 entry = make_uint2(xm | (ym << 8) | (zm << 16), static_cast<uint32_t>(gg[u]));
                const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) sm.warp_count[u & 1][warp] = __popc(ballot);
                __syncthreads();
                int off = nlist, total = 0;
#pragma unroll
                for (int k = 0; k < NWARP; ++k) {
                    const int c = sm.warp_count[u & 1][k];
                    if (k < warp) off += c;
                    total += c;
                }
                if (hit) sm.list[off + __popc(ballot & lanemask_lt())] = entry;
                nlist += total;
                cpos += NT;
            }
            __syncthreads();
        }
        // ---- order the entries by clipped-box volume (counting sort; the order of the sum does not matter here) ----
        int mykey[kBinSeg / NT];
#pragma unroll
        for (int r = 0; r < kBinSeg / NT; ++r) {
            const int i = tid + r * NT;
            mykey[r] = 0;
            if (i < nlist) {
                const uint32_t m = sm.list[i].x;
                const int vol = __popc(m & 0xffu) * __popc((m >> 8) & 0xfu) * __popc(m >> 16);
                mykey[r] = (vol + 7) >> 3;
                atomicAdd(&sm.hist[mykey[r]], 1);
            }
        }
        __syncthreads();
        if (warp == 0) {   // exclusive prefix over the keys: lane L owns keys 2L+1, 2L+2 (key 0 is unused)
            const int c0 = sm.hist[2 * lane + 1], c1 = sm.hist[2 * lane + 2];
            int incl = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int base = incl - c0 - c1;
            sm.hist[2 * lane + 1] = base;
            sm.hist[2 * lane + 2] = base + c0;
            // first entry of the wide class: keys >= kWideVol / 8
            constexpr int kWideKey = kWideVol / 8;
            if (2 * lane + 1 == kWideKey) sm.n_narrow = base;
            if (2 * lane + 2 == kWideKey) sm.n_narrow = base + c0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kBinSeg / NT; ++r) {
            const int i = tid + r * NT;
            if (i < nlist) sm.sorted[atomicAdd(&sm.hist[mykey[r]], 1)] = sm.list[i];
        }
        if (!tiles_ready) {   // the tiles were in flight during the first Phase A
            mbar_wait(&sm.bar, 0);
            tiles_ready = true;
            // canonical-order check of MY bin's points (point n must lie in voxel n): this kernel assumed it; a
            // violation clears the flag, bin_finish_kernel then discards the sums and the Gaussian-centric kernels
            // (which build the voxel -> point map) produce the gradients instead
#pragma unroll
            for (int r = 0; r < 512 / NT; ++r) {
                const int i = tid + r * NT, col = i >> 4, z = i & 15;
                const int X = binX0 + (col >> 2), Y = binY0 + (col & 3), Z = binZ0 + z;
                if (X < H && Y < W && Z < D) {
                    int vx, vy, vz;
                    if (p.in.points_int) {
                        const long long n = (static_cast<long long>(X) * W + Y) * D + Z;
                        vx = p.in.points_int[3 * n]; vy = p.in.points_int[3 * n + 1]; vz = p.in.points_int[3 * n + 2];
                    } else {
                        const float *pp = &sm.pts[col * (kBinZ * 3) + z * 3];
                        vx = voxel_coord(pp[0], p.d.pc_min[0], p.d.grid_size);
                        vy = voxel_coord(pp[1], p.d.pc_min[1], p.d.grid_size);
                        vz = voxel_coord(pp[2], p.d.pc_min[2], p.d.grid_size);
                    }
                    if (vx != X || vy != Y || vz != Z) *p.canon = 0;   // benign race: every writer stores 0
                }
            }
        }
        __syncthreads();

        // ======================= Phase B: eight lanes (or a whole warp) per listed Gaussian ===========
        const float *records = bp.records + static_cast<size_t>(sample) * p.d.G * 32;
        float *sums = bp.sums + static_cast<size_t>(sample) * p.d.G * 32;
        const int n_narrow = sm.n_narrow, narrow_tickets = (n_narrow + 3) >> 2;
        const int tickets = narrow_tickets + (nlist - n_narrow);
#pragma unroll 1
        while (true) {
            int ticket = 0;
            if (lane == 0) ticket = atomicAdd(&sm.next, 1);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            if (ticket >= tickets) break;
            const bool wide = ticket >= narrow_tickets;     // warp-uniform
            const int grp = lane >> 3, sub = lane & 7;
            const int e = wide ? n_narrow + (ticket - narrow_tickets) : 4 * ticket + grp;
            const bool have = wide || e < n_narrow;
            const uint2 ent = have ? sm.sorted[e] : make_uint2(0u, 0u);
            const int g = static_cast<int>(ent.y);
            const uint32_t xm = ent.x & 0xffu, ym = (ent.x >> 8) & 0xfu, zm = ent.x >> 16;
            const int x0 = __ffs(xm) - 1, nx = __popc(xm), y0 = __ffs(ym) - 1, ny = __popc(ym), z0 = __ffs(zm) - 1, nz = __popc(zm);
            const int vol = nx * ny * nz;   // 0 for a lane group without an entry
            GaussAcc<C, PROB> acc;
            acc.load_record(reinterpret_cast<const float4 *>(records + static_cast<size_t>(g) * 32));
            // My lanes stride over the clipped box, z fastest; the stride is decomposed once into box steps.  A whole-warp
            // entry walks the two z halves of the box one after the other: the halves of the gradient tile are two TMA
            // boxes whose distance is a multiple of 128 bytes, so lanes on the same column and z & 7 of different halves
            // would hit the same banks (ncu: 39 % of the shared wavefronts of the prob config were such replays), while
            // neighbouring columns of ONE half sit 16 banks apart.
            const int first = wide ? lane : sub, stride = wide ? 32 : 8;
            const int nseg = wide ? 2 : 1;
#pragma unroll 1
            for (int sg = 0; sg < nseg; ++sg) {
                int zs = z0, nzh = nz;
                if (wide) {
                    zs = sg == 0 ? z0 : max(z0, 8);
                    nzh = sg == 0 ? min(z0 + nz, 8) - z0 : z0 + nz - zs;
                    if (nzh <= 0) continue;
                }
                const int volh = nx * ny * nzh;
                int t = small_div(first, nzh);
                int iz = first - t * nzh, ix = small_div(t, ny), iy = t - ix * ny;
                const int t2 = small_div(stride, nzh), sz = stride - t2 * nzh, sx = small_div(t2, ny), sy = t2 - sx * ny;
#pragma unroll 1
                for (int idx = first; idx < volh; idx += stride) {
                    const int col = (x0 + ix) * kBinY + (y0 + iy), Z = zs + iz;
                    PairData<C, PROB> pd;
                    pd.ok = true;
                    pd.shift = false;
                    const float *pp = &sm.pts[col * (kBinZ * 3) + Z * 3];
                    pd.px = pp[0]; pd.py = pp[1]; pd.pz = pp[2];
                    const float *row = &sm.grad[(Z >> 3) * (256 * C) + (col * 8 + (Z & 7)) * C];
                    if constexpr ((C & 1) == 0) {   // rows start 8-byte aligned
#pragma unroll
                        for (int k = 0; k < CP2; ++k) {
                            const float2 v = *reinterpret_cast<const float2 *>(row + 2 * k);
                            if (k & 1) { pd.raw[k >> 1].z = v.x; pd.raw[k >> 1].w = v.y; } else { pd.raw[k >> 1].x = v.x; pd.raw[k >> 1].y = v.y; }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < CP2; ++k) {
                            const float a = row[2 * k], b = (2 * k + 1 < C) ? row[2 * k + 1] : 0.f;
                            if (k & 1) { pd.raw[k >> 1].z = a; pd.raw[k >> 1].w = b; } else { pd.raw[k >> 1].x = a; pd.raw[k >> 1].y = b; }
                        }
                    }
                    if (PROB) pd.ax = sm.aux[col * kBinZ + Z];
                    acc.template consume<false>(pd);
                    iz += sz;
                    const int cz = iz >= nzh;
                    iz -= cz ? nzh : 0;
                    iy += sy + cz;
                    const int cy = iy >= ny;
                    iy -= cy ? ny : 0;
                    ix += sx + cy;
                }
            }
            __syncwarp();
            float x[32];
            acc.to_vector(x);
            group8_transpose_reduce(x, lane);
            if (wide) {   // the four lane groups worked on the same Gaussian
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    x[i] += __shfl_xor_sync(0xffffffffu, x[i], 8);
                    x[i] += __shfl_xor_sync(0xffffffffu, x[i], 16);
                }
            }
            if (have && (!wide || grp == 0)) {
                const int v0 = ((sub & 4) ? 16 : 0) + ((sub & 2) ? 8 : 0) + ((sub & 1) ? 4 : 0);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (v0 + i < GaussAcc<C, PROB>::kVals) atomicAdd(sums + static_cast<size_t>(v0 + i) * p.d.G + g, x[i]);
            }
        }
    }
}

// One thread per Gaussian: the raw sums of all bins ([32][G], value-major: coalesced reads) -> the gradients, i.e. the
// per-Gaussian linear maps of finish() (splat_bwd_common.cuh): d(mean) = -A * sum(w d), d(cov) = -(1/2 | 1) * sum(w d d^T)
// (+ the determinant terms of the prob variant), d(sem) = opacity * sum(E g) for the base variant.
template <int C, bool PROB>
__global__ void __launch_bounds__(128) bin_finish_kernel(const BwdParams pb, const float *sums_all) {
    pdl_launch_dependents();
    pdl_wait();
    const BwdParams p = sample_bwd(pb, blockIdx.y);
    if (*p.canon == 0) return;
    const int g = blockIdx.x * 128 + threadIdx.x;
    const int G = p.d.G;
    if (g >= G) return;
    const float *sums = sums_all + static_cast<size_t>(blockIdx.y) * 32 * G + g;
    float v[32];
#pragma unroll
    for (int i = 0; i < GaussAcc<C, PROB>::kVals; ++i) v[i] = __ldg(sums + static_cast<size_t>(i) * G);
    float c6[6];
    load_cov6_in(p.d, p.in, g, c6);
    const float a = c6[0], b = c6[1], c = c6[2], d = c6[3], e = c6[4], f = c6[5];
    p.gr.means_grad[3 * g + 0] = -(a * v[0] + d * v[1] + f * v[2]);
    p.gr.means_grad[3 * g + 1] = -(d * v[0] + b * v[1] + e * v[2]);
    p.gr.means_grad[3 * g + 2] = -(f * v[0] + e * v[1] + c * v[2]);
    p.gr.opacity_grad[g] = v[3];
    float gc[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) gc[i] = (i < 3) ? -0.5f * v[4 + i] : -v[4 + i];
    if (PROB) {
        const float sg = v[10 + C];
        const float m[6] = {b * c - e * e, a * c - f * f, a * b - d * d, 2.f * (e * f - c * d), 2.f * (d * f - a * e), 2.f * (d * e - b * f)};
#pragma unroll
        for (int i = 0; i < 6; ++i) gc[i] = fmaf(sg, m[i], gc[i]);
    }
    if (p.d.cov_stride == 9) {   // the six gathered entries [0,4,8,1,5,2] of the 3x3 carry the gradient, the lower triangle gets 0
        float *o = p.gr.cov_grad + 9 * static_cast<size_t>(g);
        o[0] = gc[0]; o[1] = gc[3]; o[2] = gc[5]; o[3] = 0.f; o[4] = gc[1]; o[5] = gc[4]; o[6] = 0.f; o[7] = 0.f; o[8] = gc[2];
    } else {
        float *o = p.gr.cov_grad + 6 * static_cast<size_t>(g);
#pragma unroll
        for (int i = 0; i < 6; ++i) o[i] = gc[i];
    }
    const float scale = PROB ? 1.f : p.in.opacities[g];
#pragma unroll
    for (int k = 0; k < C; ++k) p.gr.semantics_grad[static_cast<size_t>(g) * C + k] = scale * v[10 + k];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// fp32 tensor [rows][W][inner] (contiguous) with boxes {box_inner, 4, 8}
static bool make_map(CUtensorMap *m, const void *base, long long rows, int W, int inner, int box_inner) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(inner) * 4, static_cast<cuuint64_t>(inner) * 4 * W};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(box_inner), kBinY, kBinX};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Can the bin-centric kernel serve this call (if the points turn out to be in canonical order on the device)?
bool backward_bin_eligible(const gf_splat_desc &d, const gf_splat_inputs &in, const gf_splat_grads &gr) {
    static const bool forced_off = [] {
        const char *e = getenv("GF_B200_BWD");
        return e && e[0] == 'g';   // GF_B200_BWD=gauss: the deterministic Gaussian-centric kernels
    }();
    if (forced_off || encode_tiled() == nullptr) return false;
    if (static_cast<long long>(d.N) != static_cast<long long>(d.H) * d.W * d.D || d.N == 0 || d.G == 0) return false;
    if ((d.D & 3) != 0 || 8 * d.C > 256) return false;   // 16-byte strides of the tensor maps, box <= 256 elements
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    return al16(in.pts) && al16(gr.logits_grad);
}

template <int C, bool PROB>
static int launch_bin_t(const BwdParams &bp, const BinParams &np, const BinMaps &maps, cudaStream_t stream) {
    const gf_splat_desc &d = bp.d;
    const int B = batch_of(d);
    const int nzc = (d.D + kBinZ - 1) / kBinZ, nby = (d.W + kBinY - 1) / kBinY;
    GF_REQUIRE(nby <= 65535 && static_cast<long long>(np.nbx) * B <= 65535, GF_ERR_UNSUPPORTED,
               "splat backward: grid x batch too large for the bin launch");
    const size_t smem = sizeof(BinSmem<C, PROB>);
    // per call: the attribute is per device, and a process may drive several
    GF_CUDA_TRY(cudaFuncSetAttribute(backward_bin_kernel<C, PROB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nzc, nby, np.nbx * B);
    cfg.blockDim = dim3(kBinThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    GF_CUDA_TRY(cudaLaunchKernelEx(&cfg, backward_bin_kernel<C, PROB>, bp, np, maps));
    const float *sums = np.sums;
    GF_CUDA_TRY(launch_chained(bin_finish_kernel<C, PROB>, dim3((d.G + 127) / 128, B), dim3(128), 0, stream, bp, sums));
    return GF_OK;
}

// bp: the launch parameters of the backward (batch-level pointers); fws: the supertile lists built for this call
int launch_backward_bin(const BwdParams &bp, const SplatWorkspace &fws, float *sums, cudaStream_t stream) {
    const gf_splat_desc &d = bp.d;
    const long long B = batch_of(d);
    BinMaps maps;
    bool ok = make_map(&maps.grad, bp.gr.logits_grad, B * d.H, d.W, d.D * d.C, 8 * d.C);
    ok = ok && make_map(&maps.pts, bp.in.pts, (d.pts_shared ? 1 : B) * d.H, d.W, d.D * 3, kBinZ * 3);
    if (d.variant == GF_SPLAT_PROB) ok = ok && make_map(&maps.aux, bp.aux, B * d.H, d.W, d.D * 4, kBinZ * 4);
    else maps.aux = maps.pts;
    GF_REQUIRE(ok, GF_ERR_CUDA, "splat backward: cuTensorMapEncodeTiled failed");
    BinParams np;
    np.records = fws.records;
    np.sums = sums;
    np.boxes = fws.boxes;
    np.lists = fws.lists;
    np.counts = fws.counts;
    np.st = fws.st;
    np.nsy = fws.nsy;
    np.nsuper = fws.nsuper;
    np.nbx = (d.H + kBinX - 1) / kBinX;
    const bool prob = d.variant == GF_SPLAT_PROB;
#define GF_CASE(CC)                                                      \
    case CC:                                                             \
        return prob ? launch_bin_t<CC, true>(bp, np, maps, stream)       \
                    : launch_bin_t<CC, false>(bp, np, maps, stream);
    switch (d.C) {
        GF_CASE(16)
        GF_CASE(17)
        GF_CASE(18)
        GF_CASE(19)
        GF_CASE(20)
        default:
            set_error("splat backward: class count C=%d is not compiled in (supported: 16..20)", d.C);
            return GF_ERR_UNSUPPORTED;
    }
#undef GF_CASE
}

}  // namespace gf
