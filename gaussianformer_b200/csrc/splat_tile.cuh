// The tile walker of the render kernel: a CTA owns a bin of 8 x 4 columns x 16 z; it resolves its own ordered
// Gaussian list from the supertile list (Phase A), streams the records through a shared-memory ring (Phase B)
// and lets every lane walk ITS OWN hits of a batch in ascending Gaussian order (lane-private traversal).
//
// Measured and removed (profiles/README.md has the numbers): warp-uniform traversal (first generation), a warp =
// the bin's 8 x 4 columns at one z quad, 64-record batches, the software-pipelined and the branch-free fused step,
// 2 voxels per thread, 3 CTAs per SM, a K split of the last-round bins, early candidate prefetch.
#pragma once
#include "splat_render.cuh"

namespace gf {

#ifndef GF_RENDER_CTAS
#define GF_RENDER_CTAS 4   // resident CTAs per SM the base tile kernel is compiled for
#endif
constexpr int kVoxT = 4;        // voxels per thread: a z quad
constexpr int kQuadSeg = 512;   // list entries resolved per segment
constexpr int kBatch = 32;      // records staged per ring slot; a lane keeps its hit mask of a batch in one word
#ifndef GF_TILE_RING
#define GF_TILE_RING 5   // measured: 5, 6, 7 slots 72.0 us per step, 4 and 8 slots 73.4 us
#endif
constexpr int kRing = GF_TILE_RING;   // ring slots: a warp may run up to kRing-2 batches ahead of the slowest one

// Shared-memory row of a staged record: the 128-byte record followed by 16 bytes of padding, so that chunk c of row j
// starts at bank group (j + c) mod 8 and lanes of one warp reading the same chunk of DIFFERENT records fall into
// different bank groups -- the effect of an XOR swizzle, but a chunk's address is row + 16*c, an immediate offset of the
// load (the XOR cost one LOP3 per chunk and step: 8 of 92 instructions).
template <int C>
struct RenderSmem {
    static constexpr int REC = rec_floats(C);
    static constexpr int ROW = REC + 4;        // floats per staged row
    static constexpr int NT = kRenderThreads;
    alignas(128) float stage[kRing][kBatch * ROW];
    alignas(8) uint2 list[kQuadSeg + kBatch];  // x: box relative to the bin as bit masks, y: Gaussian index
    uint32_t hitw[kQuadSeg / kBatch + 1][16];  // per batch: bit j of word i = entry j covers x position i (0..7), y position i-8 (8..11), z quad i-12 (12..15)
    alignas(8) uint64_t bar_full[kRing];       // records of the slot have landed (one cp.async arrival per thread)
    alignas(8) uint64_t bar_empty[kRing];      // every warp is done with the slot (one arrival per warp)
    int warp_count[2][NT / 32];
    float ce[NT / 32][2];                      // per-warp cross-entropy partials of the epilogue
#ifdef GF_RENDER_TIMING
    unsigned long long t_phase[8];             // cycle sums: prologue, Phase A, Phase B, epilogue (thread 0); all warps: wait-full, wait-empty + issue, hit masks, walk
#endif
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// the mbarrier receives one arrival from this thread when all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive_on(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_one(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One staged record as a lane sees it: `addr` is the shared-space address of its row, chunk i sits 16*i bytes further.
struct RecView {
    uint32_t addr;
    __device__ __forceinline__ float4 chunk(int i) const {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr + 16u * static_cast<uint32_t>(i)));
        return v;
    }
};

// The caller supplies one callable:
//   step(RecView record, uint32_t zbits, bool active)   evaluate + accumulate one (record, my z quad) pair; `active`
//       says whether this lane has a record in this step, bit v of zbits whether its voxel v lies inside the box.
//       Inactive lanes must not touch `record`.
//
// Lane-private traversal: a Gaussian's box covers only part of a warp's 4 x 4 x 8 footprint (17.7 of 32 lanes on the
// nuScenes workload), so marching all lanes through every record that touches the footprint leaves almost half of them
// idle in every step.  Instead each lane gets the bit mask of the records of the batch that cover ITS column and z
// quad -- the box masks are separable, so 16 ballots per batch and CTA (8 x bits, 4 y bits, 4 z quads; three word loads
// and two ANDs per lane) produce all 128 masks -- and walks its own bits in ascending order (the reference's summation order per voxel); the warp iterates
// max-over-lanes popcount times instead of once per touching record.
// Two step functors: `step_fast` is used when `fast` (CTA-uniform) says every thread qualifies for it, `step` otherwise;
// only the innermost loop exists twice.
template <int C, class StepFast, class Step>
__device__ __forceinline__ void walk_tile(const RenderParams &p, RenderSmem<C> &sm, int binX0, int binY0, int binZ0,
                                          int my_zshift, bool fast, StepFast &&step_fast, Step &&step) {
    constexpr int REC = rec_floats(C), ROW = RenderSmem<C>::ROW;
    constexpr int NT = kRenderThreads, NWARP = NT / 32, VOX = kVoxT;
    constexpr uint32_t VMASK = (1u << VOX) - 1u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.d.H, W = p.d.W, D = p.d.D;
    if (tid == 0) {
#pragma unroll
        for (int r = 0; r < kRing; ++r) {
            mbar_init(&sm.bar_full[r], NT);
            mbar_init(&sm.bar_empty[r], NWARP);
        }
        mbar_fence_init();
    }
    uint32_t gb = 0;   // batches consumed so far by this CTA: drives ring slots and barrier parities
    // (the __syncthreads of Phase A below publishes the barrier initialisation)

    // ---- candidates: the ascending list of this bin's supertile ------------------------------------
    const int st_shift = 31 - __clz(p.st);
    const int s = (binX0 >> st_shift) * p.nsy + (binY0 >> st_shift);
    const int ncand = p.counts[s];
    const int32_t *cand = p.lists + static_cast<size_t>(s) * p.d.G;
    const uint32_t bX1 = min(binX0 + kBinX, H) - 1, bY1 = min(binY0 + kBinY, W) - 1, bZ1 = min(binZ0 + kBinZ, D) - 1;

    int cpos = 0;
    while (cpos < ncand) {
        __syncthreads();   // previous segment fully consumed (and, the first time, barriers initialised)
#ifdef GF_RENDER_TIMING
        const long long tA0 = clock64();
#endif
        // ======================= Phase A: ordered survivors of the box test ==========================
        int nlist = 0;
        while (cpos < ncand && nlist + NT <= kQuadSeg) {
            constexpr int kPre = 4;   // rounds fetched together (memory-level parallelism)
            int gg[kPre];
            uint4 bb[kPre];
#pragma unroll
            for (int u = 0; u < kPre; ++u) {
                const int i = cpos + u * NT + tid;
                gg[u] = i < ncand ? __ldg(cand + i) : -1;
            }
#pragma unroll
            for (int u = 0; u < kPre; ++u)
                bb[u] = gg[u] >= 0 ? __ldg(reinterpret_cast<const uint4 *>(p.boxes) + gg[u]) : make_uint4(1u, 1u, 1u, 1u);
#pragma unroll
            for (int u = 0; u < kPre; ++u) {
                if (cpos >= ncand || nlist + NT > kQuadSeg) break;   // uniform
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
                for (int k = 0; k < NT / 32; ++k) {
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
        // pad the last batch with empty entries (all-zero masks, so nobody visits them)
        if (tid < kBatch && nlist + tid < ((nlist + kBatch - 1) / kBatch) * kBatch) sm.list[nlist + tid] = make_uint2(0u, 0u);
        __syncthreads();
        const int nchunks = (nlist + kBatch - 1) / kBatch;
        auto issue = [&](int k, uint32_t b_index) {  // batch k of this segment -> ring slot b_index % kRing
            const int slot = b_index % kRing;
            const uint32_t use = b_index / kRing;
            if (use > 0) mbar_wait(&sm.bar_empty[slot], (use - 1) & 1);   // previous occupant released by all warps
#pragma unroll
            for (int q = 0; q < (kBatch * 8 + NT - 1) / NT; ++q) {     // 32 records x 8 x 16 B = 256 copies
                const int piece = tid + NT * q, row = piece >> 3, col = (piece & 7) * 4;
                if (piece < kBatch * 8 && k * kBatch + row < nlist) {
                    const uint32_t g = sm.list[k * kBatch + row].y;
                    cp_async16(&sm.stage[slot][row * ROW + col], p.records + static_cast<size_t>(g) * REC + col);
                }
            }
            cp_async_arrive_on(&sm.bar_full[slot]);
        };
#pragma unroll 1
        for (int k = 0; k < kRing - 1 && k < nchunks; ++k) issue(k, gb + k);
        // The box masks are separable, so 16 ballots per batch -- taken ONCE per CTA, the batches dealt to the four warps --
        // transpose them into 16 words (which entries cover x position i / y position i / z quad i); a lane's hit mask of a
        // batch is then the AND of three of those words (ten ballots per warp and batch before: 9 % of Phase B).
        for (int b = warp; b < (nlist + kBatch - 1) / kBatch; b += NWARP) {
            const uint32_t ex = sm.list[b * kBatch + lane].x;
            uint32_t mine = 0u;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const bool covers = i < 12 ? ((ex >> i) & 1u) != 0u : ((ex >> (16 + VOX * (i - 12))) & VMASK) != 0u;
                const uint32_t w = __ballot_sync(0xffffffffu, covers);
                if (lane == i) mine = w;
            }
            if (lane < 16) sm.hitw[b][lane] = mine;
        }
        __syncthreads();
#ifdef GF_RENDER_TIMING
        const long long tB0 = clock64();
        if (tid == 0) sm.t_phase[1] += static_cast<unsigned long long>(tB0 - tA0);
#endif

        // ======================= Phase B: stream records and accumulate ==============================
        // No CTA-wide barrier in this loop: full[] / empty[] mbarriers let the four warps drift apart by
        // up to kRing-2 batches, so a warp whose footprint is touched by few records does not wait.
#pragma unroll 1
        for (int k = 0; k < nchunks; ++k, ++gb) {
#ifdef GF_RENDER_TIMING
            const long long tk0 = clock64();
#endif
            const int slot = gb % kRing;
            const uint32_t stage_base = smem_u32(&sm.stage[slot][0]);
            // bit j of `hits`: record j of this batch covers my column and my z quad (padded entries are all-zero)
            uint32_t hits;
            {
                const uint32_t *hw = sm.hitw[k];
                hits = hw[4 * (warp & 1) + (lane >> 3)] & hw[8 + ((lane >> 1) & 3)] & hw[12 + 2 * (warp >> 1) + (lane & 1)];
            }
#ifdef GF_RENDER_TIMING
            const long long tw0 = clock64();
#endif
            mbar_wait(&sm.bar_full[slot], (gb / kRing) & 1);
#ifdef GF_RENDER_TIMING
            const long long tw1 = clock64();
#endif
            // the warp takes as many steps as its busiest lane has hits in this batch
            const int nsteps = __reduce_max_sync(0xffffffffu, __popc(hits));
            auto pop = [&](bool &act, RecView &rv, uint32_t &zb) {
                act = hits != 0;
                const int jraw = __ffs(static_cast<int>(hits)) - 1;          // my lowest remaining hit; -1 when I have none
                hits &= hits - 1;                                             // 0 stays 0
                rv.addr = stage_base + static_cast<uint32_t>(jraw * (ROW * 4));   // an inactive lane never dereferences it
                const uint32_t e = sm.list[k * kBatch + (act ? jraw : 0)].x;
                zb = (e >> my_zshift) & VMASK;
            };
            if (fast) {
#pragma unroll 2
                for (int st = 0; st < nsteps; ++st) {
                    bool act; RecView rv; uint32_t zb;
                    pop(act, rv, zb);
                    step_fast(rv, zb, act);
                }
            } else {
#pragma unroll 2
                for (int st = 0; st < nsteps; ++st) {
                    bool act; RecView rv; uint32_t zb;
                    pop(act, rv, zb);
                    step(rv, zb, act);
                }
            }
            __syncwarp();
#ifdef GF_RENDER_TIMING
            const long long tw2 = clock64();
#endif
            if (lane == 0) mbar_arrive_one(&sm.bar_empty[slot]);       // my warp is done with this slot
            if (k + kRing - 1 < nchunks) issue(k + kRing - 1, gb + kRing - 1);
#ifdef GF_RENDER_TIMING
            if (lane == 0) {
                atomicAdd(&sm.t_phase[4], static_cast<unsigned long long>(tw1 - tw0));
                atomicAdd(&sm.t_phase[5], static_cast<unsigned long long>(clock64() - tw2));
                atomicAdd(&sm.t_phase[6], static_cast<unsigned long long>(tw0 - tk0));
                atomicAdd(&sm.t_phase[7], static_cast<unsigned long long>(tw2 - tw1));
            }
#endif
        }
#ifdef GF_RENDER_TIMING
        if (tid == 0) sm.t_phase[2] += static_cast<unsigned long long>(clock64() - tB0);
#endif
    }
}

}  // namespace gf
