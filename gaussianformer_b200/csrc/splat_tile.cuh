// The tile walker shared by the forward and backward tile kernels: a CTA owns a bin of 8 x 4 columns x
// 16 z; it resolves its own ordered Gaussian list from the supertile list (Phase A), streams the
// records through a shared-memory ring (Phase B) and calls `visit` for every record that touches the
// calling warp's footprint, in ascending Gaussian order, with all 32 lanes converged.
#pragma once
#include <type_traits>

#include "splat_render.cuh"

namespace gf {

#ifndef GF_RENDER_CTAS
#define GF_RENDER_CTAS 4   // resident CTAs per SM the base tile kernel is compiled for
#endif
#ifndef GF_RENDER_VOX
#define GF_RENDER_VOX 4   // voxels per thread of the tile kernel (2 or 4)
#endif
constexpr int kQuadSeg = 512;   // list entries resolved per segment
// lane -> voxel mapping of a warp: 0 = 4 x 4 columns x two z groups (first generation), 1 = the bin's 8 x 4 columns
// at ONE z group per warp (lanes of a warp then see the same z statistics: ~6 % fewer walk steps)
#ifndef GF_TILE_MAP
#define GF_TILE_MAP 0
#endif
#ifndef GF_TILE_PIPE
#define GF_TILE_PIPE 0
#endif
#ifndef GF_TILE_BATCH
#define GF_TILE_BATCH 32
#endif
constexpr int kBatch = GF_TILE_BATCH;   // records staged per ring slot (32 or 64; 64 needs GF_TILE_LANEWALK)
static_assert(kBatch == 32 || kBatch == 64, "a lane keeps its hit mask of a batch in one 32- or 64-bit word");
#ifndef GF_TILE_RING
#define GF_TILE_RING 5   // measured: 5, 6, 7 slots 72.0 us per step, 4 and 8 slots 73.4 us
#endif
constexpr int kRing = GF_TILE_RING;   // ring slots: a warp may run up to kRing-2 batches ahead of the slowest one
// 1: every lane walks its OWN hits of a batch (lane-private traversal, see walk_tile); 0: the warp visits every
// record that touches its footprint with all lanes on the same record (first generation)
#ifndef GF_TILE_LANEWALK
#define GF_TILE_LANEWALK 1
#endif

template <int C, int VOX>
struct RenderSmem {
    static constexpr int REC = rec_floats(C);
    static constexpr int NT = 512 / VOX;   // threads per CTA: 8 x 4 columns x (16 / VOX) z groups
    alignas(128) float stage[kRing][kBatch * REC];
    alignas(8) uint2 list[kQuadSeg + kBatch];  // x: box relative to the bin as bit masks, y: index | warp-hit bits
    alignas(8) uint64_t bar_full[kRing];       // records of the slot have landed (one cp.async arrival per thread)
    alignas(8) uint64_t bar_empty[kRing];      // every warp is done with the slot (one arrival per warp)
    int warp_count[2][NT / 32];
#if GF_TILE_PIPE == 2
    alignas(128) float zero_row[32];           // the "record" of a lane without a hit: weights 0, classes 0
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


// One staged record as a lane sees it: a 128-byte slot whose eight 16-byte chunks are stored XOR-swizzled by the
// slot's row (chunk c of row j sits at position c ^ (j & 7)), so that lanes of one warp reading the same chunk of
// DIFFERENT records fall into different bank groups.  `addr` is the shared-space address of the row with the
// swizzle already folded in (the row is 128-byte aligned, so the fold is an OR), chunk i is one XOR away.
struct RecView {
    uint32_t addr;
    __device__ __forceinline__ float4 chunk(int i) const {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr ^ (static_cast<uint32_t>(i) << 4)));
        return v;
    }
};

// The caller supplies the two stages of a lane's step:
//   stage_e(float4 g0, float4 g1, float4 g2, RecView record, uint32_t zbits, bool active)   geometry chunks 0..2 (already
//       loaded) -> weights; `active` says whether this lane has a record to evaluate in this step (its column lies
//       inside the Gaussian's box and at least one of its voxels does); bit v of zbits whether its voxel v does.
//       Inactive lanes must not touch `record`.
//   stage_acc(bool active)   accumulate with what the last stage_e left behind.
// GF_TILE_PIPE = 2 (unmeasured, prepared for the next round): branch-free fused step.  The two callables become
//   prime(RecView, zbits, active)            evaluate a lane's first hit of the batch (weights only)
//   fused(RecView next, zbits, active_next)  accumulate the CURRENT hit and evaluate the NEXT one in ONE basic block,
// so that the compiler can interleave the 36 independent FFMA2 of the accumulation with the dependent chain
// (shared load -> quadratic form -> ex2) of the next exponent.  A lane without a hit points at an all-zero row
// (weights 0, classes 0) instead of branching, so no lane ever multiplies a record it is not entitled to.
// GF_TILE_PIPE = 1 software-pipelines the two: the geometry of a lane's NEXT hit is requested before stage_acc of the
// current one, so its shared-memory latency hides behind the accumulation instead of stalling the next exponent.
//
// Lane-private traversal (GF_TILE_LANEWALK): a Gaussian's box covers only part of a warp's 4 x 4 x 2*VOX
// footprint (17.7 of 32 lanes on the nuScenes workload), so marching all lanes through every record that touches
// the footprint leaves almost half of them idle in every step.  Instead each lane gets the bit mask of the
// records of the batch that cover ITS column and z group -- the box masks are separable, so ten ballots
// (4 x bits, 4 y bits, 2 z groups) and three selects produce all 32 masks -- and walks its own bits in ascending
// order (the reference's summation order per voxel); the warp iterates max-over-lanes popcount times instead of
// once per touching record, and every step does useful work in every lane that still has hits.
template <int C, int VOX, class StageE, class StageAcc>
__device__ __forceinline__ void walk_tile(const RenderParams &p, RenderSmem<C, VOX> &sm, int binX0, int binY0, int binZ0,
                                          uint32_t my_xy, int my_zshift, StageE &&stage_e, StageAcc &&stage_acc) {
    constexpr int REC = rec_floats(C);
    constexpr int NT = 512 / VOX, NWARP = NT / 32;
    constexpr uint32_t VMASK = (1u << VOX) - 1u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.d.H, W = p.d.W, D = p.d.D;
    if (tid == 0) {
        if (smem_u32(&sm.stage[0][0]) & 127u) __trap();   // RecView folds the swizzle into the address with an OR
#pragma unroll
        for (int r = 0; r < kRing; ++r) {
            mbar_init(&sm.bar_full[r], NT);
            mbar_init(&sm.bar_empty[r], NWARP);
        }
        mbar_fence_init();
    }
#if GF_TILE_PIPE == 2
    if (tid < 32) sm.zero_row[tid] = 0.f;   // published by the first __syncthreads of Phase A
#endif
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
                // which warp footprints (x half, z half) does the clipped box touch?
                uint32_t wh = 0;
#pragma unroll
                for (int wq = 0; wq < NWARP; ++wq)
                    if ((xm & (0xFu << (4 * (wq & 1)))) && (zm & (((1u << (2 * VOX)) - 1u) << (2 * VOX * (wq >> 1))))) wh |= 1u << wq;
                const uint2 entry = make_uint2(xm | (ym << 8) | (zm << 16), static_cast<uint32_t>(gg[u]) | (wh << 24));
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
        // pad the last batch with empty entries (no warp-hit bits, so nobody visits them)
        if (tid < kBatch && nlist + tid < ((nlist + kBatch - 1) / kBatch) * kBatch) sm.list[nlist + tid] = make_uint2(0u, 0u);
        __syncthreads();

        // ======================= Phase B: stream records and accumulate ==============================
        // No CTA-wide barrier in this loop: full[] / empty[] mbarriers let the four warps drift apart by
        // up to kRing-2 batches, so a warp whose footprint is touched by few records does not wait.
        const int nchunks = (nlist + kBatch - 1) / kBatch;
        auto issue = [&](int k, uint32_t b_index) {  // batch k of this segment -> ring slot b_index % kRing
            const int slot = b_index % kRing;
            const uint32_t use = b_index / kRing;
            if (use > 0) mbar_wait(&sm.bar_empty[slot], (use - 1) & 1);   // previous occupant released by all warps
#pragma unroll
            for (int q = 0; q < (kBatch * 8 + NT - 1) / NT; ++q) {     // 32 records x 8 x 16 B = 256 copies
                const int piece = tid + NT * q, row = piece >> 3, col = (piece & 7) * 4;
                if (piece < kBatch * 8 && k * kBatch + row < nlist) {
                    const uint32_t g = sm.list[k * kBatch + row].y & 0x00FFFFFFu;
                    const int dcol = (((piece & 7) ^ (row & 7)) * 4);      // swizzled chunk position (see RecView)
                    cp_async16(&sm.stage[slot][row * REC + dcol], p.records + static_cast<size_t>(g) * REC + col);
                }
            }
            cp_async_arrive_on(&sm.bar_full[slot]);
        };
#pragma unroll 1
        for (int k = 0; k < kRing - 1 && k < nchunks; ++k) issue(k, gb + k);
#pragma unroll 1
        for (int k = 0; k < nchunks; ++k, ++gb) {
            const int slot = gb % kRing;
            const uint32_t stage_base = smem_u32(&sm.stage[slot][0]);
#if GF_TILE_LANEWALK
            // bit j of `hits`: record j of this batch covers my column and my z group (padded entries are all-zero)
            using HitMask = typename std::conditional<kBatch == 64, unsigned long long, uint32_t>::type;
            HitMask hits = 0;
#pragma unroll
            for (int h = 0; h < kBatch / 32; ++h) {
                const uint32_t ex = sm.list[k * kBatch + 32 * h + lane].x;
#if GF_TILE_MAP == 1
                uint32_t bx[8], by[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) bx[i] = __ballot_sync(0xffffffffu, (ex >> i) & 1u);
#pragma unroll
                for (int i = 0; i < 4; ++i) by[i] = __ballot_sync(0xffffffffu, (ex >> (8 + i)) & 1u);
                const uint32_t bzq = __ballot_sync(0xffffffffu, ((ex >> (16 + VOX * warp)) & VMASK) != 0u);
                const int sx = lane & 7, sy = lane >> 3;
                const uint32_t x03 = (sx & 2) ? ((sx & 1) ? bx[3] : bx[2]) : ((sx & 1) ? bx[1] : bx[0]);
                const uint32_t x47 = (sx & 2) ? ((sx & 1) ? bx[7] : bx[6]) : ((sx & 1) ? bx[5] : bx[4]);
                const uint32_t word = ((sx & 4) ? x47 : x03) &
                                      (sy == 0 ? by[0] : sy == 1 ? by[1] : sy == 2 ? by[2] : by[3]) & bzq;
#else
                const int xh = 4 * (warp & 1), zg = 16 + 2 * VOX * (warp >> 1);
                const int sx = lane >> 3, sy = (lane >> 1) & 3;
                uint32_t bx[4], by[4], bz[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    bx[i] = __ballot_sync(0xffffffffu, (ex >> (xh + i)) & 1u);
                    by[i] = __ballot_sync(0xffffffffu, (ex >> (8 + i)) & 1u);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) bz[q] = __ballot_sync(0xffffffffu, ((ex >> (zg + VOX * q)) & VMASK) != 0u);
                const uint32_t word = (sx == 0 ? bx[0] : sx == 1 ? bx[1] : sx == 2 ? bx[2] : bx[3]) &
                                      (sy == 0 ? by[0] : sy == 1 ? by[1] : sy == 2 ? by[2] : by[3]) & ((lane & 1) ? bz[1] : bz[0]);
#endif
                hits |= static_cast<HitMask>(word) << (32 * h);
            }
            mbar_wait(&sm.bar_full[slot], (gb / kRing) & 1);
            auto next_hit = [&](bool &act, RecView &rv, uint32_t &zb) {   // pops my lowest remaining hit
                act = hits != 0;
                int j = 0;
                if (kBatch == 64) j = act ? __ffsll(static_cast<long long>(hits)) - 1 : 0;
                else j = act ? __ffs(static_cast<int>(hits)) - 1 : 0;
                hits &= hits - 1;                                    // 0 stays 0
                const uint32_t e = sm.list[k * kBatch + j].x;
                rv.addr = stage_base + static_cast<uint32_t>(j) * (REC * 4) + ((static_cast<uint32_t>(j) & 7u) << 4);
                zb = (e >> my_zshift) & VMASK;
            };
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#if GF_TILE_PIPE == 2
            {
                const uint32_t zero_addr = smem_u32(&sm.zero_row[0]);
                bool act_c, act_n;
                RecView rv;
                uint32_t zb;
                next_hit(act_c, rv, zb);
                if (!act_c) { rv.addr = zero_addr; zb = 0u; }
                stage_e(rv, zb, act_c);                       // prime
                while (__any_sync(0xffffffffu, act_c)) {
                    next_hit(act_n, rv, zb);
                    if (!act_n) { rv.addr = zero_addr; zb = 0u; }
                    stage_acc(rv, zb, act_n);                 // fused: accumulate current, evaluate next
                    act_c = act_n;
                }
                (void)zero4;
            }
#elif GF_TILE_PIPE
            bool act_c, act_n;
            RecView rv;
            uint32_t zb;
            float4 g0 = zero4, g1 = zero4, g2 = zero4;
            next_hit(act_c, rv, zb);
            if (act_c) { g0 = rv.chunk(0); g1 = rv.chunk(1); g2 = rv.chunk(2); }
            stage_e(g0, g1, g2, rv, zb, act_c);
            while (__any_sync(0xffffffffu, act_c)) {
                next_hit(act_n, rv, zb);
                if (act_n) { g0 = rv.chunk(0); g1 = rv.chunk(1); g2 = rv.chunk(2); }   // in flight during the accumulation
                stage_acc(act_c);
                stage_e(g0, g1, g2, rv, zb, act_n);
                act_c = act_n;
            }
#else
            while (__any_sync(0xffffffffu, hits != 0)) {
                bool act;
                RecView rv;
                uint32_t zb;
                next_hit(act, rv, zb);
                float4 g0 = zero4, g1 = zero4, g2 = zero4;
                if (act) { g0 = rv.chunk(0); g1 = rv.chunk(1); g2 = rv.chunk(2); }
                stage_e(g0, g1, g2, rv, zb, act);
                stage_acc(act);
            }
#endif
#else
            static_assert(kBatch == 32 && GF_TILE_MAP == 0, "the first-generation walk: one hit word per batch, warp-hit bits of mapping 0");
            // records that touch my warp's footprint, in ascending order (the hit bits are gathered before the
            // wait for the records).  Fetching the next hit's entry / geometry ahead of the current visit was
            // measured slower (75.3 / 81.4 vs 73.4 us per step): the registers it takes cost more than the
            // latency it hides.
            const uint2 mine = sm.list[k * kBatch + lane];
            uint32_t todo = __ballot_sync(0xffffffffu, (mine.y >> (24 + warp)) & 1u);
            mbar_wait(&sm.bar_full[slot], (gb / kRing) & 1);
            while (todo) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint32_t e = sm.list[k * kBatch + j].x;
                const uint32_t zb = (e >> my_zshift) & VMASK;
                RecView rv;
                rv.addr = stage_base + static_cast<uint32_t>(j) * (REC * 4) + ((static_cast<uint32_t>(j) & 7u) << 4);
                const bool act = (e & my_xy) == my_xy && zb != 0u;
                const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 g0 = zero4, g1 = zero4, g2 = zero4;
                if (act) { g0 = rv.chunk(0); g1 = rv.chunk(1); g2 = rv.chunk(2); }
                stage_e(g0, g1, g2, rv, zb, act);
                stage_acc(act);
            }
#endif
            __syncwarp();
            if (lane == 0) mbar_arrive_one(&sm.bar_empty[slot]);       // my warp is done with this slot
            if (k + kRing - 1 < nchunks) issue(k + kRing - 1, gb + kRing - 1);
        }
    }

}

}  // namespace gf
