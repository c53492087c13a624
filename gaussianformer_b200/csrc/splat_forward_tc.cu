// Splat render kernel, second generation: the per-voxel class accumulation as a dense contraction
// on the 5th-generation tensor cores.
//
//   out[v, c] = sum_g  w[v, g] * s[g, c]      w[v,g] = k_g * exp2(q_g(x_v))  if voxel(v) in box(g) else 0
//
// (reference: FORWARD::renderCUDA, model/head/localagg/src/forward.cu:61-81; prob variant
// model/head/localagg_prob/src/forward.cu:63-101.)  One 128-thread CTA owns a bin of 4x4 columns x
// 8 z = 128 voxels, one voxel per thread = one row of the MMA; a warp covers a compact 4x4x2 block
// so that most of its lanes fall inside a Gaussian's box together.  Per batch of 16 listed Gaussians:
//
//   * records arrive by per-record 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//   * every thread evaluates its row of W (exact integer-box test, exp2 of the pre-scaled quadratic
//     form) on the CUDA cores, splits each value into two TF32 terms (hi = top 19 bits, lo = w - hi)
//     and stores them into the K-major, un-swizzled canonical UMMA layout in shared memory;
//   * the class-matrix tile S (hi / lo) is written in the same K-major canonical layout;
//   * one thread issues tcgen05.mma (kind::tf32, M=128, N=32, K=8): hi*hi + lo*hi + hi*lo, i.e. the
//     "3xTF32" scheme, ~2^-21 relative error per product, accumulating in fp32 in TENSOR MEMORY;
//     tcgen05.commit -> mbarrier tells the CTA when the operand tiles may be overwritten.
//
// The epilogue reads the 128x32 fp32 accumulator with tcgen05.ld (one row per thread) and writes
// the C logits of each voxel.  For the prob variant an extra all-ones class column makes the tensor
// core produce Z = sum_g P as well; density and the product of complements stay in registers.
//
// Points that are not in canonical voxel order are detected per thread (W row forced to zero) and
// evaluated afterwards by render_one_point(), so no second kernel is needed when N == H*W*D.
#include <cstdlib>

#include "splat_render.cuh"

namespace gf {

extern thread_local cudaEvent_t g_ev_before, g_ev_after;  // measurement hooks (cabi.cu)

constexpr int kTcThreads = 128;
constexpr int kTcK = 16;    // Gaussians per operand tile
constexpr int kTcN = 32;    // MMA N (classes padded)
constexpr int kTcSeg = 512; // list entries resolved per segment
constexpr int kTcBinX = 4, kTcBinY = 4, kTcBinZ = 8;
constexpr uint32_t kTmemCols = 32;

// canonical UMMA layouts (byte offsets), cf. cute/atom/mma_traits_sm100.hpp "make_umma_desc":
//   A  (K-major,  SWIZZLE_NONE): (m%8)*16 + (m/8)*SBO_A + (k/4)*LBO_A + (k%4)*4,  SBO_A = 128, LBO_A = 2048
//   B  (K-major,  SWIZZLE_NONE): (n%8)*16 + (n/8)*SBO_B + (k/4)*LBO_B + (k%4)*4,  SBO_B = 128, LBO_B = 512
// (LBO = byte distance between the two 16-byte K chunks of one MMA, SBO = distance between 8-row
// groups; verified on hardware by tools/umma_test.cu.  MN-major operands are NOT usable with
// kind::tf32 + SWIZZLE_NONE: the same probe returns all zeros for them.)
constexpr uint32_t kSboA = 128, kLboA = 2048, kSboB = 128, kLboB = (kTcN / 8) * 128;

template <int C>
struct TcSmem {
    static constexpr int REC = rec_floats(C);
    alignas(128) float rec[2][kTcK * REC];
    alignas(128) uint32_t a_hi[128 * kTcK];
    alignas(128) uint32_t a_lo[128 * kTcK];
    alignas(128) uint32_t b_hi[kTcN * kTcK];
    alignas(128) uint32_t b_lo[kTcN * kTcK];
    alignas(8) uint2 list[kTcSeg + kTcK];  // x: box relative to the bin (bit masks), y: Gaussian index
    alignas(8) uint64_t bar_rec[2];
    alignas(8) uint64_t bar_full;
    alignas(8) uint64_t bar_free;
    uint32_t tmem_base;
    int warp_count[2][kTcThreads / 32];   // double-buffered: one barrier per compaction round
    int nlist, last;
};

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;  // descriptor version of sm_100
    return d;         // base offset 0, layout type 0 = SWIZZLE_NONE
}

// kind::tf32, D = F32, A = TF32 K-major, B = TF32 K-major, N = 32, M = 128
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) |
                                ((kTcN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(kInstrDesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_load_32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;\n"   // same asm statement: the registers are not readable before the wait
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- small PTX helpers of the warp-specialised pipeline ---------------------------------------------
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// the mbarrier receives one arrival from this thread when all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void compute_warps_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTcThreads) : "memory"); }

// Thread roles: threads 0..127 (4 warps) own one voxel each and produce the operand tiles; warp 4 is
// the control warp: it owns tensor memory and its lane 0 issues the MMAs, so no compute warp ever
// serialises the others behind descriptor arithmetic.  All hand-offs inside the batch loop are
// mbarriers (no CTA-wide barrier):
//   rec_full[2]   records of a batch have landed (cp.async arrivals of the control warp's 32 lanes)
//   tiles_full    all compute threads have written their part of the A / B tiles
//   tiles_free    the MMAs that read those tiles have completed (tcgen05.commit)
template <int C, bool PROB>
__global__ void __launch_bounds__(kTcThreads + 32, 6) render_tc_kernel(const RenderParams p) {
    constexpr int REC = rec_floats(C);
    constexpr int CP = REC - kGeomFloats;
    static_assert(REC == 32, "one record = 128 bytes = 8 cp.async chunks");
    static_assert(C + (PROB ? 1 : 0) <= kTcN, "class count exceeds the MMA N tile");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TcSmem<C> &sm = *reinterpret_cast<TcSmem<C> *>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_compute = warp < kTcThreads / 32;
    const int H = p.d.H, W = p.d.W, D = p.d.D;

    // ---- my voxel (compute threads) ---------------------------------------------------------------------
    const int binX0 = blockIdx.z * kTcBinX, binY0 = blockIdx.y * kTcBinY, binZ0 = blockIdx.x * kTcBinZ;
    const int lx = lane >> 3, ly = (lane >> 1) & 3, lz = 2 * (warp & 3) + (lane & 1);   // warp = 4 x 4 x 2 voxels
    const int X = binX0 + lx, Y = binY0 + ly, Z = binZ0 + lz;
    const bool valid = is_compute && X < H && Y < W && Z < D;
    const long long n = (static_cast<long long>(X) * W + Y) * D + Z;
    float px = 0.f, py = 0.f, pz = 0.f;
    bool canon = true;
    if (valid) {
        px = __ldg(p.pts + 3 * n); py = __ldg(p.pts + 3 * n + 1); pz = __ldg(p.pts + 3 * n + 2);
        int ix, iy, iz;
        if (p.points_int) {
            ix = p.points_int[3 * n]; iy = p.points_int[3 * n + 1]; iz = p.points_int[3 * n + 2];
        } else {
            // exact trunc((p - origin) / cell) costs three IEEE divisions; a point that lies well inside
            // voxel (X,Y,Z) by a cheap reciprocal estimate needs none (the estimate is off by < 1e-5 cells)
            const float inv = __frcp_rn(p.d.grid_size);
            const float fx = (px - p.d.pc_min[0]) * inv - static_cast<float>(X);
            const float fy = (py - p.d.pc_min[1]) * inv - static_cast<float>(Y);
            const float fz = (pz - p.d.pc_min[2]) * inv - static_cast<float>(Z);
            const float lo_m = 1e-3f, hi_m = 1.f - 1e-3f;
            if (fx > lo_m && fx < hi_m && fy > lo_m && fy < hi_m && fz > lo_m && fz < hi_m) {
                ix = X; iy = Y; iz = Z;
            } else {
                ix = voxel_coord(px, p.d.pc_min[0], p.d.grid_size);
                iy = voxel_coord(py, p.d.pc_min[1], p.d.grid_size);
                iz = voxel_coord(pz, p.d.pc_min[2], p.d.grid_size);
            }
        }
        canon = ix == X && iy == Y && iz == Z;
        if (!canon) atomicOr(p.flags, GF_FLAG_GENERIC_PATH);
    }
    const bool live = valid && canon;   // rows of W that may be non-zero
    // entry word of a listed Gaussian: x mask [0,4) | y mask [4,8) | z mask [8,16); I am inside its
    // box iff all three of my bits are set (dead rows get a pattern no entry can match)
    const uint32_t my_bits = live ? ((1u << lx) | (1u << (4 + ly)) | (1u << (8 + lz))) : 0xFFFFFFFFu;

    // ---- one-time setup: barriers, tensor memory, zeroed S tiles ---------------------------------
    if (tid == 0) {
        mbar_init(&sm.bar_rec[0], 32);
        mbar_init(&sm.bar_rec[1], 32);
        mbar_init(&sm.bar_full, kTcThreads);
        mbar_init(&sm.bar_free, 1);
        mbar_fence_init();
    }
    if (!is_compute) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < kTcN * kTcK / 4; i += kTcThreads + 32) {
        reinterpret_cast<uint4 *>(sm.b_hi)[i] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4 *>(sm.b_lo)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    float dens = 0.f, keep = 1.f;
    uint32_t gc = 0;   // batches processed so far by this CTA (drives every ring slot / parity)

    // byte offset of my row inside an A tile
    const uint32_t a_row = (tid & 7) * 16 + ((tid & 127) >> 3) * kSboA;

    // ---- candidates: the ascending list of this bin's supertile ------------------------------------
    const int st_shift = 31 - __clz(p.st);   // the supertile edge is a power of two
    const int s = (binX0 >> st_shift) * p.nsy + (binY0 >> st_shift);
    const int ncand = p.counts[s];
    const int32_t *cand = p.lists + static_cast<size_t>(s) * p.d.G;
    const uint32_t bX1 = min(binX0 + kTcBinX, H) - 1, bY1 = min(binY0 + kTcBinY, W) - 1, bZ1 = min(binZ0 + kTcBinZ, D) - 1;

    int cpos = 0;
    bool last;
    do {
        // ======================= Phase A (compute warps): ordered survivors of the box test ===========
        if (is_compute) {
            int nlist = 0;
            while (cpos < ncand && nlist + kTcThreads <= kTcSeg) {
                // fetch up to kPre rounds of candidates and their boxes before touching any of them
                // (two dependent L2/HBM round trips per super-round instead of per round)
                constexpr int kPre = 4;
                int gg[kPre];
                uint4 bb[kPre];
#pragma unroll
                for (int u = 0; u < kPre; ++u) {
                    const int i = cpos + u * kTcThreads + tid;
                    gg[u] = i < ncand ? __ldg(cand + i) : -1;
                }
#pragma unroll
                for (int u = 0; u < kPre; ++u)
                    bb[u] = gg[u] >= 0 ? __ldg(reinterpret_cast<const uint4 *>(p.boxes) + gg[u]) : make_uint4(1u, 1u, 1u, 1u);
#pragma unroll
                for (int u = 0; u < kPre; ++u) {
                    if (cpos >= ncand || nlist + kTcThreads > kTcSeg) break;   // uniform
                    const uint4 b = bb[u];
                    const uint32_t x0 = b.x & 0xffffu, x1 = b.x >> 16, y0 = b.y & 0xffffu, y1 = b.y >> 16,
                                   z0 = b.z & 0xffffu, z1 = b.z >> 16;
                    const bool hit = gg[u] >= 0 && x0 <= bX1 && x1 >= static_cast<uint32_t>(binX0) && y0 <= bY1 &&
                                     y1 >= static_cast<uint32_t>(binY0) && z0 <= bZ1 && z1 >= static_cast<uint32_t>(binZ0) &&
                                     b.w == 0u;
                    // box relative to the bin as three bit masks
                    const int rx0 = max(static_cast<int>(x0) - binX0, 0), rx1 = min(static_cast<int>(x1) - binX0, kTcBinX - 1);
                    const int ry0 = max(static_cast<int>(y0) - binY0, 0), ry1 = min(static_cast<int>(y1) - binY0, kTcBinY - 1);
                    const int rz0 = max(static_cast<int>(z0) - binZ0, 0), rz1 = min(static_cast<int>(z1) - binZ0, kTcBinZ - 1);
                    const uint32_t xm = ((2u << rx1) - 1u) & ~((1u << rx0) - 1u);
                    const uint32_t ym = ((2u << ry1) - 1u) & ~((1u << ry0) - 1u);
                    const uint32_t zm = ((2u << rz1) - 1u) & ~((1u << rz0) - 1u);
                    const uint2 entry = make_uint2(xm | (ym << 4) | (zm << 8), static_cast<uint32_t>(gg[u]));
                    const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
                    if (lane == 0) sm.warp_count[u & 1][warp] = __popc(ballot);
                    compute_warps_sync();
                    int off = nlist, total = 0;
#pragma unroll
                    for (int k = 0; k < kTcThreads / 32; ++k) {
                        const int c = sm.warp_count[u & 1][k];
                        if (k < warp) off += c;
                        total += c;
                    }
                    if (hit) sm.list[off + __popc(ballot & lanemask_lt())] = entry;
                    nlist += total;
                    cpos += kTcThreads;
                }
                compute_warps_sync();
            }
            // pad the last batch with empty entries (mask 0 never matches)
            if (tid < kTcK && nlist + tid < ((nlist + kTcK - 1) / kTcK) * kTcK) sm.list[nlist + tid] = make_uint2(0u, 0u);
            if (tid == 0) {
                sm.nlist = nlist;
                sm.last = cpos >= ncand ? 1 : 0;
            }
        }
        __syncthreads();
        const int nlist = sm.nlist;
        last = sm.last != 0;
        const int nchunks = (nlist + kTcK - 1) / kTcK;

        if (!is_compute) {
            // ======================= control warp: record ring + tensor-core issue ========================
            auto load_records = [&](int k, uint32_t g_index) {   // batch k of this segment -> ring slot g_index & 1
                const int slot = g_index & 1;
#pragma unroll
                for (int q = 0; q < 4; ++q) {                      // 16 records x 8 chunks of 16 bytes = 128 copies
                    const int piece = lane + 32 * q, row = piece >> 3, col = (piece & 7) * 4;
                    if (k * kTcK + row < nlist) {
                        const uint32_t g = sm.list[k * kTcK + row].y;
                        cp_async_16(&sm.rec[slot][row * REC + col], p.records + static_cast<size_t>(g) * REC + col);
                    }
                }
                cp_async_arrive(&sm.bar_rec[slot]);
            };
            const uint32_t a_hi = smem_u32(sm.a_hi), a_lo = smem_u32(sm.a_lo);
            const uint32_t b_hi = smem_u32(sm.b_hi), b_lo = smem_u32(sm.b_lo);
            if (nchunks > 0) load_records(0, gc);
            if (nchunks > 1) load_records(1, gc + 1);
            for (int k = 0; k < nchunks; ++k, ++gc) {
                // all compute threads have written the tiles of this batch (and are done with its records)
                uint32_t spins = 0;
                while (!mbar_try_wait(&sm.bar_full, gc & 1)) {
                    __nanosleep(64);
                    if (++spins > (1u << 22)) __trap();
                }
                if (lane == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < kTcK / 8; ++ks) {
                        const uint64_t dah = umma_smem_desc(a_hi + ks * 2 * kLboA, kLboA, kSboA);
                        const uint64_t dal = umma_smem_desc(a_lo + ks * 2 * kLboA, kLboA, kSboA);
                        const uint64_t dbh = umma_smem_desc(b_hi + ks * 2 * kLboB, kLboB, kSboB);
                        const uint64_t dbl = umma_smem_desc(b_lo + ks * 2 * kLboB, kLboB, kSboB);
                        umma_tf32(tmem, dah, dbh, (gc > 0 || ks > 0) ? 1u : 0u);
                        umma_tf32(tmem, dal, dbh, 1u);
                        umma_tf32(tmem, dah, dbl, 1u);
                    }
                    umma_commit(&sm.bar_free);
                }
                __syncwarp();
                if (k + 2 < nchunks) load_records(k + 2, gc + 2);   // ring slot gc & 1 is free again
            }
        } else {
            // ======================= compute warps: W / S tiles ===========================================
            for (int k = 0; k < nchunks; ++k, ++gc) {
                const int slot = gc & 1;
                const int cnt = min(kTcK, nlist - k * kTcK);
                mbar_wait(&sm.bar_rec[slot], (gc >> 1) & 1);

                // ---- my row of W for these 16 Gaussians: branch-free so the 16 evaluations overlap --------
                float w[kTcK];
#pragma unroll
                for (int j = 0; j < kTcK; ++j) {
                    const uint32_t e = sm.list[k * kTcK + j].x;   // warp-uniform
                    const float4 *r4 = reinterpret_cast<const float4 *>(&sm.rec[slot][j * REC]);
                    const float4 g0 = r4[0], g1 = r4[1];
                    const float2 g2 = *reinterpret_cast<const float2 *>(r4 + 2);
                    const float dx = g0.x - px, dy = g0.y - py, dz = g0.z - pz;
                    float t1 = g1.x * dx;
                    t1 = fmaf(g1.w, dy, t1);
                    t1 = fmaf(g2.y, dz, t1);
                    float t2 = g1.y * dy;
                    t2 = fmaf(g2.x, dz, t2);
                    float q = t1 * dx;
                    q = fmaf(t2, dy, q);
                    q = fmaf(g1.z * dz, dz, q);
                    const bool in = (e & my_bits) == my_bits;
                    const float Eraw = ex2_approx(q);
                    const float E = in ? Eraw : 0.f;          // select AFTER the arithmetic: padded slots hold
                    w[j] = in ? (PROB ? g0.w * Eraw : Eraw) : 0.f;   // stale bytes and must never leak a NaN into W (base: opacity is in S)
                    if (PROB) {
                        dens += E;
                        keep *= (1.f - E);
                    }
                }
                // ---- my classes of Gaussian kk for the S tile: lane -> (class mod 8, kk mod 4) makes both the
                //      scalar stores below bank-conflict free --------------------------------------------------
                const int kk = (lane & 3) + 4 * warp, nn_low = lane >> 2;
                constexpr int kNg = (C + (PROB ? 1 : 0) + 7) / 8;   // groups of 8 classes actually used
                float sv[kNg];
#pragma unroll
                for (int e8 = 0; e8 < kNg; ++e8) {
                    const int nn = nn_low + 8 * e8;
                    float v = 0.f;
                    if (kk < cnt && nn < C) v = sm.rec[slot][kk * REC + kGeomFloats + nn];
                    if (PROB && kk < cnt && nn == C) v = 1.f;       // all-ones column -> Z
                    sv[e8] = v;
                }

                // ---- the operand tiles are free once the previous batch's MMAs have completed -----------
                if (gc > 0) mbar_wait(&sm.bar_free, (gc - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int kc = 0; kc < kTcK / 4; ++kc) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float v = w[kc * 4 + i];
                        hi[i] = __float_as_uint(v) & 0xFFFFE000u;
                        lo[i] = __float_as_uint(v - __uint_as_float(hi[i]));
                    }
                    *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(sm.a_hi) + a_row + kc * kLboA) =
                        make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(sm.a_lo) + a_row + kc * kLboA) =
                        make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
#pragma unroll
                for (int e8 = 0; e8 < kNg; ++e8) {
                    const int nn = nn_low + 8 * e8;
                    const uint32_t off = (nn & 7) * 16 + (nn >> 3) * kSboB + (kk >> 2) * kLboB + (kk & 3) * 4;
                    const uint32_t hi = __float_as_uint(sv[e8]) & 0xFFFFE000u;
                    const uint32_t lo = __float_as_uint(sv[e8] - __uint_as_float(hi));
                    *reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(sm.b_hi) + off) = hi;
                    *reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(sm.b_lo) + off) = lo;
                }
                fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async proxy
                tc_fence_before();
                mbar_arrive(&sm.bar_full);
            }
        }
        __syncthreads();
    } while (!last);

    // ---- epilogue: accumulator row -> logits ------------------------------------------------------
    if (is_compute) {
        float acc[32];
        if (gc > 0) {
            mbar_wait(&sm.bar_free, (gc - 1) & 1);
            tc_fence_after();
            tmem_load_32(tmem + (static_cast<uint32_t>(warp * 32) << 16), acc);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        }
        if (live) {
            float *dst = p.out.logits + n * C;
            if (PROB) {
                const float zsum = acc[C];
                if (zsum > 1e-9f) {
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[c] = __fdiv_rn(acc[c], zsum);
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[c] = (c < C - 1) ? static_cast<float>(1.0 / (C - 1)) : 0.f;
                }
                p.out.bin_logits[n] = 1.f - keep;
                p.out.density[n] = dens;
                p.out.probability[n] = zsum;
            }
            if (p.out.argmax) {
                int best = 0;
                float bv = acc[0];
#pragma unroll
                for (int c = 1; c < C; ++c)
                    if (acc[c] > bv) { bv = acc[c]; best = c; }
                p.out.argmax[n] = static_cast<uint8_t>(best);
            }
            if ((C & 1) == 0) {
#pragma unroll
                for (int c = 0; c < C; c += 2) __stcs(reinterpret_cast<float2 *>(dst + c), make_float2(acc[c], acc[c + 1]));
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) dst[c] = acc[c];
            }
        } else if (valid) {
            render_one_point<C, PROB>(p, n, px, py, pz);   // point n does not sit in voxel n
        }
    }

    tc_fence_before();
    __syncthreads();
    if (!is_compute) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    }
}

// ------------------------------------------------------------------------------------------------
// Third-generation tcgen05 render kernel (GF_B200_RENDER=tc2; base variant; prepared for measurement, see
// DESIGN.md 7): same UMMA / TMEM / mbarrier plumbing as render_tc_kernel above, different W producer.
//
// render_tc_kernel evaluates one voxel per thread: every thread re-reads every record and pays the whole quadratic
// form per slot (~25 instructions), which is why it loses to the SIMT kernel.  Here a producer thread owns one
// COLUMN of the 4 x 4 x 8 tile (8 z voxels that share x and y) and 4 consecutive Gaussians of a 32-Gaussian batch:
// the x/y part of the exponent is evaluated once per (column, Gaussian), each slot then costs one FMA pair in dz,
// one ex2, the box mask and the TF32 split, and the four Gaussians of a row go out as ONE 16-byte store per operand
// (the K-major canonical layout keeps 4 consecutive k of a row contiguous).  Rows are ordered r = 16*z + column and
// lane <-> (column mod 8) so that the eight lanes of a quarter warp hit eight different 16-byte bank groups.
// CTAs whose 128 points are not voxel centres of their columns in canonical order take the exact per-point path.
// ------------------------------------------------------------------------------------------------
constexpr int kT2K = 32;   // Gaussians per operand tile (4 MMA k-steps)

template <int C>
struct Tc2Smem {
    static constexpr int REC = rec_floats(C);
    alignas(128) float rec[2][kT2K * REC];
    alignas(128) uint32_t a_hi[128 * kT2K];
    alignas(128) uint32_t a_lo[128 * kT2K];
    alignas(128) uint32_t b_hi[kTcN * kT2K];
    alignas(128) uint32_t b_lo[kTcN * kT2K];
    alignas(16) float4 pts[128];               // (x, y, z, live) of row r = 16*z + column
    alignas(8) uint2 list[kTcSeg + kT2K];
    alignas(8) uint64_t bar_rec[2];
    alignas(8) uint64_t bar_full;
    alignas(8) uint64_t bar_free;
    uint32_t tmem_base;
    int warp_count[2][kTcThreads / 32];
    int nlist, last;
};

template <int C>
__global__ void __launch_bounds__(kTcThreads + 32, 4) render_tc2_kernel(const RenderParams p) {
    constexpr int REC = rec_floats(C);
    static_assert(REC == 32, "one record = 128 bytes = 8 cp.async chunks");
    static_assert(C <= kTcN, "class count exceeds the MMA N tile");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Tc2Smem<C> &sm = *reinterpret_cast<Tc2Smem<C> *>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_compute = warp < kTcThreads / 32;
    const int H = p.d.H, W = p.d.W, D = p.d.D;
    const int binX0 = blockIdx.z * kTcBinX, binY0 = blockIdx.y * kTcBinY, binZ0 = blockIdx.x * kTcBinZ;

    // ---- epilogue role: thread tid < 128 owns row r = tid = 16*z + column, column = 4*cy + cx ----------------
    const int ez = (tid >> 4) & 7, ecol = tid & 15;
    const int X = binX0 + (ecol & 3), Y = binY0 + (ecol >> 2), Z = binZ0 + ez;
    const bool valid = is_compute && X < H && Y < W && Z < D;
    const long long n = (static_cast<long long>(X) * W + Y) * D + Z;
    float px = 0.f, py = 0.f, pz = 0.f;
    bool canon = false;
    if (valid) {
        px = __ldg(p.pts + 3 * n); py = __ldg(p.pts + 3 * n + 1); pz = __ldg(p.pts + 3 * n + 2);
        int ix, iy, iz;
        if (p.points_int) {
            ix = p.points_int[3 * n]; iy = p.points_int[3 * n + 1]; iz = p.points_int[3 * n + 2];
        } else {
            ix = voxel_coord(px, p.d.pc_min[0], p.d.grid_size);
            iy = voxel_coord(py, p.d.pc_min[1], p.d.grid_size);
            iz = voxel_coord(pz, p.d.pc_min[2], p.d.grid_size);
        }
        canon = ix == X && iy == Y && iz == Z;
        if (!canon) atomicOr(p.flags, GF_FLAG_GENERIC_PATH);
    }
    if (is_compute) sm.pts[tid] = make_float4(px, py, pz, (valid && canon) ? 1.f : 0.f);

    // ---- one-time setup: barriers, tensor memory, zeroed S tiles --------------------------------------------
    if (tid == 0) {
        mbar_init(&sm.bar_rec[0], 32);
        mbar_init(&sm.bar_rec[1], 32);
        mbar_init(&sm.bar_full, kTcThreads);
        mbar_init(&sm.bar_free, 1);
        mbar_fence_init();
    }
    if (!is_compute) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < kTcN * kT2K / 4; i += kTcThreads + 32) {
        reinterpret_cast<uint4 *>(sm.b_hi)[i] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4 *>(sm.b_lo)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    // ---- producer role: column pcol of the tile, Gaussians 4*kgrp .. 4*kgrp+3 of every batch --------------------
    const int pcol = (lane & 7) + 8 * ((lane >> 3) & 1), kgrp = (warp & 3) * 2 + (lane >> 4);
    float cpx = 0.f, cpy = 0.f, cpz[kTcBinZ];
    bool col_fast = true;
#pragma unroll
    for (int z = 0; z < kTcBinZ; ++z) {
        const float4 q = sm.pts[16 * z + pcol];
        if (z == 0) { cpx = q.x; cpy = q.y; }
        cpz[z] = q.z;
        col_fast = col_fast && q.w != 0.f && q.x == cpx && q.y == cpy;
    }
    const int cta_fast = __syncthreads_and((!is_compute || col_fast) ? 1 : 0);
    const uint32_t col_bits = (1u << (pcol & 3)) | (1u << (4 + (pcol >> 2)));   // entry word: x mask [0,4) | y mask [4,8) | z mask [8,16)

    uint32_t gc = 0;   // batches processed so far by this CTA
    if (cta_fast) {
    const int st_shift = 31 - __clz(p.st);
    const int s = (binX0 >> st_shift) * p.nsy + (binY0 >> st_shift);
    const int ncand = p.counts[s];
    const int32_t *cand = p.lists + static_cast<size_t>(s) * p.d.G;
    const uint32_t bX1 = min(binX0 + kTcBinX, H) - 1, bY1 = min(binY0 + kTcBinY, W) - 1, bZ1 = min(binZ0 + kTcBinZ, D) - 1;
    // byte offset of my column's rows inside an A tile: row r = 16*z + pcol -> (r%8)*16 + (r/8)*SBO + (k/4)*LBO
    const uint32_t a_col = (pcol & 7) * 16 + (pcol >> 3) * kSboA + kgrp * kLboA;

    int cpos = 0;
    bool last;
    do {
        // ======================= Phase A (compute warps): ordered survivors of the box test ===========
        if (is_compute) {
            int nlist = 0;
            while (cpos < ncand && nlist + kTcThreads <= kTcSeg) {
                constexpr int kPre = 4;
                int gg[kPre];
                uint4 bb[kPre];
#pragma unroll
                for (int u = 0; u < kPre; ++u) {
                    const int i = cpos + u * kTcThreads + tid;
                    gg[u] = i < ncand ? __ldg(cand + i) : -1;
                }
#pragma unroll
                for (int u = 0; u < kPre; ++u)
                    bb[u] = gg[u] >= 0 ? __ldg(reinterpret_cast<const uint4 *>(p.boxes) + gg[u]) : make_uint4(1u, 1u, 1u, 1u);
#pragma unroll
                for (int u = 0; u < kPre; ++u) {
                    if (cpos >= ncand || nlist + kTcThreads > kTcSeg) break;   // uniform
                    const uint4 b = bb[u];
                    const uint32_t x0 = b.x & 0xffffu, x1 = b.x >> 16, y0 = b.y & 0xffffu, y1 = b.y >> 16,
                                   z0 = b.z & 0xffffu, z1 = b.z >> 16;
                    const bool hit = gg[u] >= 0 && x0 <= bX1 && x1 >= static_cast<uint32_t>(binX0) && y0 <= bY1 &&
                                     y1 >= static_cast<uint32_t>(binY0) && z0 <= bZ1 && z1 >= static_cast<uint32_t>(binZ0) &&
                                     b.w == 0u;
                    const int rx0 = max(static_cast<int>(x0) - binX0, 0), rx1 = min(static_cast<int>(x1) - binX0, kTcBinX - 1);
                    const int ry0 = max(static_cast<int>(y0) - binY0, 0), ry1 = min(static_cast<int>(y1) - binY0, kTcBinY - 1);
                    const int rz0 = max(static_cast<int>(z0) - binZ0, 0), rz1 = min(static_cast<int>(z1) - binZ0, kTcBinZ - 1);
                    const uint32_t xm = ((2u << rx1) - 1u) & ~((1u << rx0) - 1u);
                    const uint32_t ym = ((2u << ry1) - 1u) & ~((1u << ry0) - 1u);
                    const uint32_t zm = ((2u << rz1) - 1u) & ~((1u << rz0) - 1u);
                    const uint2 entry = make_uint2(xm | (ym << 4) | (zm << 8), static_cast<uint32_t>(gg[u]));
                    const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
                    if (lane == 0) sm.warp_count[u & 1][warp] = __popc(ballot);
                    compute_warps_sync();
                    int off = nlist, total = 0;
#pragma unroll
                    for (int k = 0; k < kTcThreads / 32; ++k) {
                        const int c = sm.warp_count[u & 1][k];
                        if (k < warp) off += c;
                        total += c;
                    }
                    if (hit) sm.list[off + __popc(ballot & lanemask_lt())] = entry;
                    nlist += total;
                    cpos += kTcThreads;
                }
                compute_warps_sync();
            }
            // pad the last batch with empty entries (mask 0 never matches)
            if (tid < kT2K && nlist + tid < ((nlist + kT2K - 1) / kT2K) * kT2K) sm.list[nlist + tid] = make_uint2(0u, 0u);
            if (tid == 0) {
                sm.nlist = nlist;
                sm.last = cpos >= ncand ? 1 : 0;
            }
        }
        __syncthreads();
        const int nlist = sm.nlist;
        last = sm.last != 0;
        const int nchunks = (nlist + kT2K - 1) / kT2K;

        if (!is_compute) {
            // ======================= control warp: record ring + tensor-core issue ========================
            auto load_records = [&](int k, uint32_t g_index) {   // batch k of this segment -> ring slot g_index & 1
                const int slot = g_index & 1;
#pragma unroll
                for (int q = 0; q < kT2K * 8 / 32; ++q) {          // 32 records x 8 chunks of 16 bytes = 256 copies
                    const int piece = lane + 32 * q, row = piece >> 3, col = (piece & 7) * 4;
                    if (k * kT2K + row < nlist) {
                        const uint32_t g = sm.list[k * kT2K + row].y;
                        cp_async_16(&sm.rec[slot][row * REC + col], p.records + static_cast<size_t>(g) * REC + col);
                    }
                }
                cp_async_arrive(&sm.bar_rec[slot]);
            };
            const uint32_t a_hi = smem_u32(sm.a_hi), a_lo = smem_u32(sm.a_lo);
            const uint32_t b_hi = smem_u32(sm.b_hi), b_lo = smem_u32(sm.b_lo);
            if (nchunks > 0) load_records(0, gc);
            if (nchunks > 1) load_records(1, gc + 1);
            for (int k = 0; k < nchunks; ++k, ++gc) {
                uint32_t spins = 0;
                while (!mbar_try_wait(&sm.bar_full, gc & 1)) {
                    __nanosleep(64);
                    if (++spins > (1u << 22)) __trap();
                }
                if (lane == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < kT2K / 8; ++ks) {
                        const uint64_t dah = umma_smem_desc(a_hi + ks * 2 * kLboA, kLboA, kSboA);
                        const uint64_t dal = umma_smem_desc(a_lo + ks * 2 * kLboA, kLboA, kSboA);
                        const uint64_t dbh = umma_smem_desc(b_hi + ks * 2 * kLboB, kLboB, kSboB);
                        const uint64_t dbl = umma_smem_desc(b_lo + ks * 2 * kLboB, kLboB, kSboB);
                        umma_tf32(tmem, dah, dbh, (gc > 0 || ks > 0) ? 1u : 0u);
                        umma_tf32(tmem, dal, dbh, 1u);
                        umma_tf32(tmem, dah, dbl, 1u);
                    }
                    umma_commit(&sm.bar_free);
                }
                __syncwarp();
                if (k + 2 < nchunks) load_records(k + 2, gc + 2);   // ring slot gc & 1 is free again
            }
        } else {
            // ======================= compute warps: W / S tiles ===========================================
            for (int k = 0; k < nchunks; ++k, ++gc) {
                const int slot = gc & 1;
                const int cnt = min(kT2K, nlist - k * kT2K);
                mbar_wait(&sm.bar_rec[slot], (gc >> 1) & 1);

                // ---- my column x my four Gaussians: E for 8 z each; selects AFTER the arithmetic (padded slots hold
                //      stale bytes that must never leak a NaN into W) ------------------------------------------------
                float e[4][kTcBinZ];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = 4 * kgrp + i;
                    const uint32_t ent = sm.list[k * kT2K + j].x;
                    const float4 *r4 = reinterpret_cast<const float4 *>(&sm.rec[slot][j * REC]);
                    const float4 g0 = r4[0], g1 = r4[1];
                    const float2 g2 = *reinterpret_cast<const float2 *>(r4 + 2);
                    const uint32_t zm = ((ent & col_bits) == col_bits) ? (ent >> 8) & 0xffu : 0u;
                    const float dx = g0.x - cpx, dy = g0.y - cpy;
                    float t1 = g1.x * dx;
                    t1 = fmaf(g1.w, dy, t1);
                    float A = t1 * dx;
                    A = fmaf(g1.y * dy, dy, A);
                    const float B = fmaf(g2.x, dy, g2.y * dx);
#pragma unroll
                    for (int z = 0; z < kTcBinZ; ++z) {
                        const float dz = g0.z - cpz[z];
                        const float q = fmaf(fmaf(g1.z, dz, B), dz, A);
                        const float Eraw = ex2_approx(q);
                        e[i][z] = ((zm >> z) & 1u) ? Eraw : 0.f;     // base variant: the opacity rides in S
                    }
                }
                // ---- S tile values: lane -> (class mod 8, kk mod 4) keeps the scalar stores conflict-free ---------
                const int nn_low = lane >> 2;
                constexpr int kNg = (C + 7) / 8;
                float sv[2][kNg];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int kk = (lane & 3) + 4 * (warp & 3) + 16 * h2;
#pragma unroll
                    for (int e8 = 0; e8 < kNg; ++e8) {
                        const int nn = nn_low + 8 * e8;
                        sv[h2][e8] = (kk < cnt && nn < C) ? sm.rec[slot][kk * REC + kGeomFloats + nn] : 0.f;
                    }
                }

                // ---- the operand tiles are free once the previous batch's MMAs have completed -----------
                if (gc > 0) mbar_wait(&sm.bar_free, (gc - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int z = 0; z < kTcBinZ; ++z) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float v = e[i][z];
                        hi[i] = __float_as_uint(v) & 0xFFFFE000u;
                        lo[i] = __float_as_uint(v - __uint_as_float(hi[i]));
                    }
                    const uint32_t off = a_col + 2 * z * kSboA;   // rows 16*z + pcol: (r/8) = 2*z + pcol/8
                    *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(sm.a_hi) + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(sm.a_lo) + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int kk = (lane & 3) + 4 * (warp & 3) + 16 * h2;
#pragma unroll
                    for (int e8 = 0; e8 < kNg; ++e8) {
                        const int nn = nn_low + 8 * e8;
                        const uint32_t off = (nn & 7) * 16 + (nn >> 3) * kSboB + (kk >> 2) * kLboB + (kk & 3) * 4;
                        const uint32_t hi = __float_as_uint(sv[h2][e8]) & 0xFFFFE000u;
                        const uint32_t lo = __float_as_uint(sv[h2][e8] - __uint_as_float(hi));
                        *reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(sm.b_hi) + off) = hi;
                        *reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(sm.b_lo) + off) = lo;
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&sm.bar_full);
            }
        }
        __syncthreads();
    } while (!last);
    }   // cta_fast

    // ---- epilogue: accumulator row -> logits ------------------------------------------------------
    if (is_compute) {
        if (cta_fast) {
            float acc[32];
            if (gc > 0) {
                mbar_wait(&sm.bar_free, (gc - 1) & 1);
                tc_fence_after();
                tmem_load_32(tmem + (static_cast<uint32_t>(warp * 32) << 16), acc);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = 0.f;
            }
            if (valid) {
                float *dst = p.out.logits + n * C;
                if (p.out.argmax) {
                    int best = 0;
                    float bv = acc[0];
#pragma unroll
                    for (int c = 1; c < C; ++c)
                        if (acc[c] > bv) { bv = acc[c]; best = c; }
                    p.out.argmax[n] = static_cast<uint8_t>(best);
                }
                if ((C & 1) == 0) {
#pragma unroll
                    for (int c = 0; c < C; c += 2) __stcs(reinterpret_cast<float2 *>(dst + c), make_float2(acc[c], acc[c + 1]));
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) dst[c] = acc[c];
                }
            }
        } else if (valid) {
            render_one_point<C, false>(p, n, px, py, pz);   // exact per-point path for this CTA's voxels
        }
    }

    tc_fence_before();
    __syncthreads();
    if (!is_compute) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    }
}

template <int C>
static int launch_render_tc2_t(const RenderParams &rp_in, cudaStream_t stream) {
    RenderParams rp = rp_in;
    rp.nby = (rp.d.W + kTcBinY - 1) / kTcBinY;
    rp.nzc = (rp.d.D + kTcBinZ - 1) / kTcBinZ;
    const int nbx = (rp.d.H + kTcBinX - 1) / kTcBinX;
    GF_REQUIRE(rp.nby <= 65535 && nbx <= 65535, GF_ERR_UNSUPPORTED, "splat: grid too large for the render launch");
    const dim3 grid(rp.nzc, rp.nby, nbx);
    const size_t smem = sizeof(Tc2Smem<C>);
    static bool configured = false;   // more than 48 KB of dynamic shared memory needs the opt-in
    if (!configured) {
        GF_CUDA_TRY(cudaFuncSetAttribute(render_tc2_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured = true;
    }
    if (g_ev_before && g_ev_after) GF_CUDA_TRY(cudaEventRecord(g_ev_before, stream));
    render_tc2_kernel<C><<<grid, kTcThreads + 32, smem, stream>>>(rp);
    GF_CUDA_TRY(cudaGetLastError());
    if (g_ev_before && g_ev_after) GF_CUDA_TRY(cudaEventRecord(g_ev_after, stream));
    return GF_OK;
}

static bool use_tc2() {
    static int cached = -1;
    if (cached < 0) {
        const char *e = getenv("GF_B200_RENDER");
        cached = (e && e[0] == 't' && e[1] == 'c' && e[2] == '2') ? 1 : 0;
    }
    return cached == 1;
}

// ------------------------------------------------------------------------------------------------
// Fourth tcgen05 render kernel (GF_B200_RENDER=tc3; base variant; prepared for measurement, DESIGN.md 7 "tc3"):
// the SIMT kernel's bin (8 x 4 columns x 16 z, one Phase A, one record ring) split into FOUR M = 128 tiles of
// 4 x 4 columns x 8 z, one producer warp per tile.  A warp compacts, per batch of 32 listed Gaussians, the ones
// that touch ITS tile into dense K positions (a tile therefore multiplies only its own Gaussians: 52 of the bin's
// 98 on the nuScenes workload), evaluates W with the column-per-thread producer of render_tc2_kernel, writes its
// own A / S operand tiles (K = 16) and issues its own tcgen05.mma into its own 32 TMEM columns; tcgen05.commit on a
// per-warp mbarrier tells the warp when the tiles may be overwritten.  Rows of a tile: r = 16*z + column,
// column = 4*cy + cx; lane = 16*khalf + column; a lane evaluates its column for the 4 consecutive K positions
// 8*step + 4*khalf .. +3 and stores them as one 16-byte word per operand and z.
// ------------------------------------------------------------------------------------------------
// GF_TC3_HALF = 1 (unmeasured): the two 8-position halves of a warp's K = 16 operand tile are flushed separately, each
// with its own completion barrier, so the warp fills one half while the tensor core reads the other instead of waiting
// for its MMAs after every second step.
#ifndef GF_TC3_HALF
#define GF_TC3_HALF 0
#endif
constexpr int kT3K = 16;          // K positions per operand tile = two steps of 8 hits
constexpr int kT3Batch = 32;      // list entries / records per ring slot
constexpr int kT3Ring = 3;
constexpr int kT3Seg = 512;
constexpr uint32_t kT3TmemCols = 128;   // four accumulators of 32 columns

template <int C>
struct Tc3Smem {
    static constexpr int REC = rec_floats(C);
    alignas(128) float stage[kT3Ring][kT3Batch * REC];
    alignas(128) uint32_t a_hi[4][128 * kT3K];
    alignas(128) uint32_t a_lo[4][128 * kT3K];
    alignas(128) uint32_t b_hi[4][kTcN * kT3K];
    alignas(128) uint32_t b_lo[4][kTcN * kT3K];
    alignas(16) float4 pts[512];                   // (x, y, z, live) of tile t, row r at [128*t + r]
    alignas(8) uint2 list[kT3Seg + kT3Batch];      // x: x mask [0,8) | y mask [8,12) | z mask [16,32); y: Gaussian index
    alignas(8) uint64_t bar_full[kT3Ring];
    alignas(8) uint64_t bar_empty[kT3Ring];
    alignas(8) uint64_t bar_mma[4][2];             // [tile][half]; GF_TC3_HALF = 0 uses [tile][0] only
    uint32_t tmem_base;
    int warp_count[2][4];
    int hits[4][kT3Batch];
    int nflush[4][2];
};

template <int C>
__global__ void __launch_bounds__(128, 2) render_tc3_kernel(const RenderParams p) {
    constexpr int REC = rec_floats(C);
    constexpr int NT = 128;
    static_assert(REC == 32, "one record = 128 bytes = 8 cp.async chunks");
    static_assert(C <= 24, "the S tile writer covers 24 class rows");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Tc3Smem<C> &sm = *reinterpret_cast<Tc3Smem<C> *>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.d.H, W = p.d.W, D = p.d.D;
    const int binX0 = blockIdx.z * kBinX, binY0 = blockIdx.y * kBinY, binZ0 = blockIdx.x * kBinZ;

    // ---- my four voxels (one row of every tile): points, canonical check -> sm.pts -----------------------------------
    const int rz = (tid >> 4) & 7, rcol = tid & 15;
    auto voxel_of = [&](int t, int &X, int &Y, int &Z) {       // row tid of tile t
        X = binX0 + 4 * (t & 1) + (rcol & 3);
        Y = binY0 + (rcol >> 2);
        Z = binZ0 + 8 * (t >> 1) + rz;
    };
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int X, Y, Z;
        voxel_of(t, X, Y, Z);
        const bool valid = X < H && Y < W && Z < D;
        const long long n = (static_cast<long long>(X) * W + Y) * D + Z;
        float x = 0.f, y = 0.f, z = 0.f;
        bool canon = false;
        if (valid) {
            x = __ldg(p.pts + 3 * n); y = __ldg(p.pts + 3 * n + 1); z = __ldg(p.pts + 3 * n + 2);
            int ix, iy, iz;
            if (p.points_int) {
                ix = p.points_int[3 * n]; iy = p.points_int[3 * n + 1]; iz = p.points_int[3 * n + 2];
            } else {
                ix = voxel_coord(x, p.d.pc_min[0], p.d.grid_size);
                iy = voxel_coord(y, p.d.pc_min[1], p.d.grid_size);
                iz = voxel_coord(z, p.d.pc_min[2], p.d.grid_size);
            }
            canon = ix == X && iy == Y && iz == Z;
            if (!canon) atomicOr(p.flags, GF_FLAG_GENERIC_PATH);
        }
        sm.pts[128 * t + tid] = make_float4(x, y, z, (valid && canon) ? 1.f : 0.f);
    }

    // ---- one-time setup ----------------------------------------------------------------------------------------------
    if (tid == 0) {
#pragma unroll
        for (int r = 0; r < kT3Ring; ++r) {
            mbar_init(&sm.bar_full[r], NT);
            mbar_init(&sm.bar_empty[r], 4);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            mbar_init(&sm.bar_mma[t][0], 1);
            mbar_init(&sm.bar_mma[t][1], 1);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "r"(kT3TmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid < 8) sm.nflush[tid >> 1][tid & 1] = 0;
    // S tiles: class rows 24..31 are never written by the producers and must read as zero
    for (int i = lane; i < kTcN * kT3K / 4; i += 32) {
        reinterpret_cast<uint4 *>(sm.b_hi[warp])[i] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4 *>(sm.b_lo[warp])[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    // ---- producer role: warp = tile, lane = 16*khalf + column ---------------------------------------------------------
    const int pcol = lane & 15, khalf = lane >> 4;
    float cpx = 0.f, cpy = 0.f, cpz[8];
    bool col_fast = true;
#pragma unroll
    for (int z = 0; z < 8; ++z) {
        const float4 q = sm.pts[128 * warp + 16 * z + pcol];
        if (z == 0) { cpx = q.x; cpy = q.y; }
        cpz[z] = q.z;
        col_fast = col_fast && q.w != 0.f && q.x == cpx && q.y == cpy;
    }
    const int cta_fast = __syncthreads_and(col_fast ? 1 : 0);
    // my tile inside the bin: x bits 4*(warp&1) .. +3, all four y, z bits 8*(warp>>1) .. +7
    const uint32_t tile_x = 0xFu << (4 * (warp & 1)), tile_z = 0xFFu << (16 + 8 * (warp >> 1));
    const uint32_t col_x = 1u << (4 * (warp & 1) + (pcol & 3)), col_y = 1u << (8 + (pcol >> 2));
    const int zshift = 16 + 8 * (warp >> 1);
    const uint32_t a_col = (pcol & 7) * 16 + (pcol >> 3) * kSboA + khalf * kLboA;   // + (kfill/4)*LBO + 2*z*SBO
    unsigned char *const my_a_hi = reinterpret_cast<unsigned char *>(sm.a_hi[warp]);
    unsigned char *const my_a_lo = reinterpret_cast<unsigned char *>(sm.a_lo[warp]);
    unsigned char *const my_b_hi = reinterpret_cast<unsigned char *>(sm.b_hi[warp]);
    unsigned char *const my_b_lo = reinterpret_cast<unsigned char *>(sm.b_lo[warp]);
    const uint32_t d_tmem = tmem + 32u * warp;      // my accumulator: columns 32*warp .. +31, all 128 lanes
    int kfill = 0, nflush = 0;
#if GF_TC3_HALF
    int nfl[2] = {0, 0};   // flushes of each half so far

    // issue the 3 MMAs of the half tile that was just filled (K positions kfill .. kfill+7) and commit them on its barrier
    auto flush_half = [&]() {
        const int half = kfill >> 3;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tc_fence_after();
            const uint32_t ko = half * 2 * kLboA, kob = half * 2 * kLboB;
            const uint64_t dah = umma_smem_desc(smem_u32(my_a_hi) + ko, kLboA, kSboA);
            const uint64_t dal = umma_smem_desc(smem_u32(my_a_lo) + ko, kLboA, kSboA);
            const uint64_t dbh = umma_smem_desc(smem_u32(my_b_hi) + kob, kLboB, kSboB);
            const uint64_t dbl = umma_smem_desc(smem_u32(my_b_lo) + kob, kLboB, kSboB);
            umma_tf32(d_tmem, dah, dbh, nflush > 0 ? 1u : 0u);
            umma_tf32(d_tmem, dal, dbh, 1u);
            umma_tf32(d_tmem, dah, dbl, 1u);
            umma_commit(&sm.bar_mma[warp][half]);
        }
        __syncwarp();
        ++nflush;
        if (half) ++nfl[1]; else ++nfl[0];
        kfill = (kfill + 8) & (kT3K - 1);
    };
#else

    // issue the MMAs of my tile for `ksteps` (1 or 2) k-steps of 8 and commit them on my barrier
    auto flush = [&](int ksteps) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tc_fence_after();
            const uint32_t ah = smem_u32(my_a_hi), al = smem_u32(my_a_lo), bh = smem_u32(my_b_hi), bl = smem_u32(my_b_lo);
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t dah = umma_smem_desc(ah + ks * 2 * kLboA, kLboA, kSboA);
                const uint64_t dal = umma_smem_desc(al + ks * 2 * kLboA, kLboA, kSboA);
                const uint64_t dbh = umma_smem_desc(bh + ks * 2 * kLboB, kLboB, kSboB);
                const uint64_t dbl = umma_smem_desc(bl + ks * 2 * kLboB, kLboB, kSboB);
                umma_tf32(d_tmem, dah, dbh, (nflush > 0 || ks > 0) ? 1u : 0u);
                umma_tf32(d_tmem, dal, dbh, 1u);
                umma_tf32(d_tmem, dah, dbl, 1u);
            }
            umma_commit(&sm.bar_mma[warp][0]);
        }
        __syncwarp();
        ++nflush;
        kfill = 0;
    };
#endif

    if (cta_fast) {
    uint32_t gb = 0;   // batches consumed so far by this CTA (ring slots / parities)
    const int st_shift = 31 - __clz(p.st);
    const int s = (binX0 >> st_shift) * p.nsy + (binY0 >> st_shift);
    const int ncand = p.counts[s];
    const int32_t *cand = p.lists + static_cast<size_t>(s) * p.d.G;
    const uint32_t bX1 = min(binX0 + kBinX, H) - 1, bY1 = min(binY0 + kBinY, W) - 1, bZ1 = min(binZ0 + kBinZ, D) - 1;

    int cpos = 0;
    while (cpos < ncand) {
        __syncthreads();   // previous segment fully consumed
        // ======================= Phase A: ordered survivors of the box test (as walk_tile, splat_tile.cuh) ==========
        int nlist = 0;
        while (cpos < ncand && nlist + NT <= kT3Seg) {
            constexpr int kPre = 4;
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
                if (cpos >= ncand || nlist + NT > kT3Seg) break;   // uniform
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
        if (tid < kT3Batch && nlist + tid < ((nlist + kT3Batch - 1) / kT3Batch) * kT3Batch) sm.list[nlist + tid] = make_uint2(0u, 0u);
        __syncthreads();

        // ======================= Phase B: record ring (full / empty mbarriers), one tile per warp ====================
        const int nchunks = (nlist + kT3Batch - 1) / kT3Batch;
        auto issue = [&](int k, uint32_t b_index) {
            const int slot = b_index % kT3Ring;
            const uint32_t use = b_index / kT3Ring;
            if (use > 0) mbar_wait(&sm.bar_empty[slot], (use - 1) & 1);
#pragma unroll
            for (int q = 0; q < kT3Batch * 8 / NT; ++q) {
                const int piece = tid + NT * q, row = piece >> 3, col = (piece & 7) * 4;
                if (k * kT3Batch + row < nlist) {
                    const uint32_t g = sm.list[k * kT3Batch + row].y;
                    cp_async_16(&sm.stage[slot][row * REC + col], p.records + static_cast<size_t>(g) * REC + col);
                }
            }
            cp_async_arrive(&sm.bar_full[slot]);
        };
#pragma unroll 1
        for (int k = 0; k < kT3Ring - 1 && k < nchunks; ++k) issue(k, gb + k);
#pragma unroll 1
        for (int k = 0; k < nchunks; ++k, ++gb) {
            const int slot = gb % kT3Ring;
            // the entries of this batch that touch my tile, compacted in ascending order
            const uint32_t ex = sm.list[k * kT3Batch + lane].x;
            const bool touch = (ex & tile_x) != 0u && (ex & tile_z) != 0u;
            const uint32_t tm = __ballot_sync(0xffffffffu, touch);
            if (touch) sm.hits[warp][__popc(tm & lanemask_lt())] = lane;
            const int nh = __popc(tm);
            mbar_wait(&sm.bar_full[slot], (gb / kT3Ring) & 1);
            __syncwarp();
            const float *stg = sm.stage[slot];
            for (int base = 0; base < nh; base += 8) {
#if GF_TC3_HALF
                {                                          // this half is free once ITS previous MMAs have completed
                    const int half = kfill >> 3, done = half ? nfl[1] : nfl[0];
                    if (done > 0) {
                        mbar_wait(&sm.bar_mma[warp][half], (done - 1) & 1);
                        tc_fence_after();
                    }
                }
#else
                if (kfill == 0 && nflush > 0) {            // my operand tiles are free once my previous MMAs have completed
                    mbar_wait(&sm.bar_mma[warp][0], (nflush - 1) & 1);
                    tc_fence_after();
                }
#endif
                // ---- W: my column x my 4 K positions (hits base + 4*khalf + i), 8 z each --------------------------------
                float e[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int hq = base + 4 * khalf + i;
                    const bool live = hq < nh;
                    const int j = live ? sm.hits[warp][hq] : 0;
                    const uint32_t ent = sm.list[k * kT3Batch + j].x;
                    const float4 *r4 = reinterpret_cast<const float4 *>(stg + j * REC);
                    const float4 g0 = r4[0], g1 = r4[1];
                    const float2 g2 = *reinterpret_cast<const float2 *>(r4 + 2);
                    const uint32_t zm = (live && (ent & col_x) && (ent & col_y)) ? (ent >> zshift) & 0xffu : 0u;
                    const float dx = g0.x - cpx, dy = g0.y - cpy;
                    float t1 = g1.x * dx;
                    t1 = fmaf(g1.w, dy, t1);
                    float A = t1 * dx;
                    A = fmaf(g1.y * dy, dy, A);
                    const float B = fmaf(g2.x, dy, g2.y * dx);
#pragma unroll
                    for (int z = 0; z < 8; ++z) {
                        const float dz = g0.z - cpz[z];
                        const float q = fmaf(fmaf(g1.z, dz, B), dz, A);
                        const float Eraw = ex2_approx(q);
                        e[i][z] = ((zm >> z) & 1u) ? Eraw : 0.f;     // select after the arithmetic; the opacity rides in S
                    }
                }
#pragma unroll
                for (int z = 0; z < 8; ++z) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float v = e[i][z];
                        hi[i] = __float_as_uint(v) & 0xFFFFE000u;
                        lo[i] = __float_as_uint(v - __uint_as_float(hi[i]));
                    }
                    const uint32_t off = a_col + (kfill >> 2) * kLboA + 2 * z * kSboA;
                    *reinterpret_cast<uint4 *>(my_a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(my_a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                // ---- S: K positions kfill + kq (+4), class rows nn_low + 8*e8; lane -> (nn_low, kq) is conflict-free --
                {
                    const int kq = lane & 3, nn_low = lane >> 2;
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int hq = base + kq + 4 * h2;
                        const bool live = hq < nh;
                        const int j = live ? sm.hits[warp][hq] : 0;
                        const int kp = kfill + kq + 4 * h2;
#pragma unroll
                        for (int e8 = 0; e8 < 3; ++e8) {
                            const int nn = nn_low + 8 * e8;
                            const float v = (live && nn < C) ? stg[j * REC + kGeomFloats + nn] : 0.f;
                            const uint32_t off = (nn & 7) * 16 + (nn >> 3) * kSboB + (kp >> 2) * kLboB + (kp & 3) * 4;
                            const uint32_t hi = __float_as_uint(v) & 0xFFFFE000u;
                            const uint32_t lo = __float_as_uint(v - __uint_as_float(hi));
                            *reinterpret_cast<uint32_t *>(my_b_hi + off) = hi;
                            *reinterpret_cast<uint32_t *>(my_b_lo + off) = lo;
                        }
                    }
                }
#if GF_TC3_HALF
                flush_half();
#else
                kfill += 8;
                if (kfill == kT3K) flush(2);
#endif
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.bar_empty[slot]);
            if (k + kT3Ring - 1 < nchunks) issue(k + kT3Ring - 1, gb + kT3Ring - 1);
        }
    }
#if GF_TC3_HALF
    if (lane == 0) { sm.nflush[warp][0] = nfl[0]; sm.nflush[warp][1] = nfl[1]; }
#else
    if (kfill == 8) flush(1);          // half-filled tile: positions 0..7 only
    if (lane == 0) { sm.nflush[warp][0] = nflush; sm.nflush[warp][1] = 0; }
#endif
    }   // cta_fast
    __syncthreads();

    // ---- epilogue: thread tid reads row tid (TMEM lane tid) of all four accumulators -------------------------------
#pragma unroll 1
    for (int t = 0; t < 4; ++t) {
        int X, Y, Z;
        voxel_of(t, X, Y, Z);
        const bool valid = X < H && Y < W && Z < D;
        const long long n = (static_cast<long long>(X) * W + Y) * D + Z;
        if (!cta_fast) {
            const float4 q = sm.pts[128 * t + tid];
            if (valid) render_one_point<C, false>(p, n, q.x, q.y, q.z);   // exact per-point path
            continue;
        }
        float acc[32];
        const int nf0 = sm.nflush[t][0], nf1 = sm.nflush[t][1], nf = nf0 + nf1;
        if (nf > 0) {
            if (nf0 > 0) mbar_wait(&sm.bar_mma[t][0], (nf0 - 1) & 1);
            if (nf1 > 0) mbar_wait(&sm.bar_mma[t][1], (nf1 - 1) & 1);
            tc_fence_after();
            tmem_load_32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 32u * t, acc);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        }
        if (valid) {
            float *dst = p.out.logits + n * C;
            if (p.out.argmax) {
                int best = 0;
                float bv = acc[0];
#pragma unroll
                for (int c = 1; c < C; ++c)
                    if (acc[c] > bv) { bv = acc[c]; best = c; }
                p.out.argmax[n] = static_cast<uint8_t>(best);
            }
            if ((C & 1) == 0) {
#pragma unroll
                for (int c = 0; c < C; c += 2) __stcs(reinterpret_cast<float2 *>(dst + c), make_float2(acc[c], acc[c + 1]));
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) dst[c] = acc[c];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kT3TmemCols));
    }
}

template <int C>
static int launch_render_tc3_t(const RenderParams &rp_in, cudaStream_t stream) {
    RenderParams rp = rp_in;
    rp.nby = (rp.d.W + kBinY - 1) / kBinY;
    rp.nzc = (rp.d.D + kBinZ - 1) / kBinZ;
    const int nbx = (rp.d.H + kBinX - 1) / kBinX;
    GF_REQUIRE(rp.nby <= 65535 && nbx <= 65535, GF_ERR_UNSUPPORTED, "splat: grid too large for the render launch");
    const dim3 grid(rp.nzc, rp.nby, nbx);
    const size_t smem = sizeof(Tc3Smem<C>);
    static bool configured = false;
    if (!configured) {
        GF_CUDA_TRY(cudaFuncSetAttribute(render_tc3_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured = true;
    }
    if (g_ev_before && g_ev_after) GF_CUDA_TRY(cudaEventRecord(g_ev_before, stream));
    render_tc3_kernel<C><<<grid, 128, smem, stream>>>(rp);
    GF_CUDA_TRY(cudaGetLastError());
    if (g_ev_before && g_ev_after) GF_CUDA_TRY(cudaEventRecord(g_ev_after, stream));
    return GF_OK;
}

static bool use_tc3() {
    static int cached = -1;
    if (cached < 0) {
        const char *e = getenv("GF_B200_RENDER");
        cached = (e && e[0] == 't' && e[1] == 'c' && e[2] == '3') ? 1 : 0;
    }
    return cached == 1;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int C, bool PROB>
static int launch_render_tc_t(const RenderParams &rp_in, cudaStream_t stream) {
    RenderParams rp = rp_in;
    rp.nby = (rp.d.W + kTcBinY - 1) / kTcBinY;
    rp.nzc = (rp.d.D + kTcBinZ - 1) / kTcBinZ;
    const int nbx = (rp.d.H + kTcBinX - 1) / kTcBinX;
    GF_REQUIRE(rp.nby <= 65535 && nbx <= 65535, GF_ERR_UNSUPPORTED, "splat: grid too large for the render launch");
    const dim3 grid(rp.nzc, rp.nby, nbx);
    const size_t smem = sizeof(TcSmem<C>);
    if (g_ev_before && g_ev_after) GF_CUDA_TRY(cudaEventRecord(g_ev_before, stream));
    render_tc_kernel<C, PROB><<<grid, kTcThreads + 32, smem, stream>>>(rp);
    GF_CUDA_TRY(cudaGetLastError());
    if (g_ev_before && g_ev_after) GF_CUDA_TRY(cudaEventRecord(g_ev_after, stream));
    return GF_OK;
}

int launch_render_tc(const RenderParams &rp, cudaStream_t stream) {
    const bool prob = rp.d.variant == GF_SPLAT_PROB;
    if (!prob && use_tc3()) {
        switch (rp.d.C) {
            case 16: return launch_render_tc3_t<16>(rp, stream);
            case 17: return launch_render_tc3_t<17>(rp, stream);
            case 18: return launch_render_tc3_t<18>(rp, stream);
            case 19: return launch_render_tc3_t<19>(rp, stream);
            case 20: return launch_render_tc3_t<20>(rp, stream);
            default: break;
        }
    }
    if (!prob && use_tc2()) {
        switch (rp.d.C) {
            case 16: return launch_render_tc2_t<16>(rp, stream);
            case 17: return launch_render_tc2_t<17>(rp, stream);
            case 18: return launch_render_tc2_t<18>(rp, stream);
            case 19: return launch_render_tc2_t<19>(rp, stream);
            case 20: return launch_render_tc2_t<20>(rp, stream);
            default: break;
        }
    }
#define GF_CASE(CC)                                                   \
    case CC:                                                          \
        return prob ? launch_render_tc_t<CC, true>(rp, stream)        \
                    : launch_render_tc_t<CC, false>(rp, stream);
    switch (rp.d.C) {
        GF_CASE(16)
        GF_CASE(17)
        GF_CASE(18)
        GF_CASE(19)
        GF_CASE(20)
        default:
            set_error("splat: class count C=%d is not compiled in (supported: 16..20)", rp.d.C);
            return GF_ERR_UNSUPPORTED;
    }
#undef GF_CASE
}

}  // namespace gf
