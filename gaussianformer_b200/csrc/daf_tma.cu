// Deformable aggregation forward with TMA-staged corner fetches (experiment behind include/gf_b200_debug.h; the
// product kernel is daf_fast_kernel, daf.cu) and the L2 -> SM gather probe that gives the sampling op its roof.
//
// One warp per sampling point, as daf_fast_kernel; but the 2 x 2 x C corner block of every visible (camera, level) pair
// arrives as ONE cp.async.bulk.tensor.4d box {C, 2, 2, 1} from a per-level tensor map [B*M][h_l][w_l][C] over the
// channels-last feature table.  Corners outside the map are zero-filled by the hardware, which is exactly the
// reference's rule (a corner contributes only if it lies inside, deformable_aggregation_cuda.cu:31-49), so no clamping
// and no validity masks remain.  Boxes land in a per-warp ring of kSlots x 4C floats guarded by one mbarrier per slot.
#include <cuda.h>

#include "daf_pair.cuh"
#include "gf_b200_debug.h"

namespace gf {

constexpr int kTmaSlots = 4;       // boxes in flight per warp
constexpr int kTmaWarps = 8;       // warps per CTA
constexpr int kTmaMaxC = 256;      // a TMA box dimension holds at most 256 elements

struct alignas(64) DafMaps {
    CUtensorMap level[kMaxLevels];
};

struct DafTmaParams {
    gf_daf_desc d;
    const float *loc;
    const float *weights;
    float *out;
    int lh[kMaxLevels], lw[kMaxLevels];
};

struct TmaPair {       // per (camera, level) pair of the current point
    int x0, y0;        // top-left corner (may be -1)
    float lh, lw;      // fractional parts
};

__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(kTmaWarps * 32) daf_tma_kernel(const DafTmaParams p, const __grid_constant__ DafMaps maps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = p.d.num_embeds, M = p.d.num_cams, L = p.d.num_scale, Gr = p.d.num_groups;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npair = M * L, gdim = C / Gr;
    // per warp: kTmaSlots boxes of 4*C floats, then the pair table, then the barriers
    float *ring = reinterpret_cast<float *>(smem_raw) + static_cast<size_t>(warp) * kTmaSlots * 4 * C;
    TmaPair *pairs = reinterpret_cast<TmaPair *>(reinterpret_cast<float *>(smem_raw) + static_cast<size_t>(kTmaWarps) * kTmaSlots * 4 * C) + warp * 32;
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<TmaPair *>(reinterpret_cast<float *>(smem_raw) + static_cast<size_t>(kTmaWarps) * kTmaSlots * 4 * C) + kTmaWarps * 32) + warp * kTmaSlots;
    if (lane < kTmaSlots) mbar_init(&bars[lane], 1);
    mbar_fence_init();
    __syncwarp();
    const uint32_t box_bytes = 4u * C * 4u;
    const long long npts = static_cast<long long>(p.d.batch) * p.d.num_pts;
    const long long warps = static_cast<long long>(gridDim.x) * kTmaWarps;
    const int my_cam = lane / L, my_lv = lane - my_cam * L;
    uint32_t uses = 0;   // boxes issued so far by this warp: slot = uses % kTmaSlots, parity = (uses / kTmaSlots) & 1

    for (long long bp = static_cast<long long>(blockIdx.x) * kTmaWarps + warp; bp < npts; bp += warps) {
        const int b = static_cast<int>(bp / p.d.num_pts);
        bool gate = false;
        if (lane < npair) {
            const float lx = __ldg(p.loc + (bp * M + my_cam) * 2), ly = __ldg(p.loc + (bp * M + my_cam) * 2 + 1);
            gate = lx > 0.f && lx < 1.f && ly > 0.f && ly < 1.f;
            const float y_im = ly * static_cast<float>(p.lh[my_lv]) - 0.5f, x_im = lx * static_cast<float>(p.lw[my_lv]) - 0.5f;
            const float yf = floorf(y_im), xf = floorf(x_im);
            TmaPair t;
            t.x0 = static_cast<int>(xf); t.y0 = static_cast<int>(yf);
            t.lh = y_im - yf; t.lw = x_im - xf;
            pairs[lane] = t;
        }
        const uint32_t visible = __ballot_sync(0xffffffffu, gate);
        __syncwarp();
        const float *wpt = p.weights + bp * npair * Gr;
        // issue side: lane 0 keeps up to kTmaSlots boxes in flight, in ascending pair order
        uint32_t to_issue = visible, to_use = visible;
        uint32_t issued = uses, used = uses;
        auto issue_one = [&]() {
            const int pr = __ffs(to_issue) - 1;
            to_issue &= to_issue - 1;
            if (lane == 0) {
                const int slot = issued % kTmaSlots;
                const int cam = pr / L, lv = pr - cam * L;
                mbar_expect_tx(&bars[slot], box_bytes);
                tma_load_4d(ring + static_cast<size_t>(slot) * 4 * C, &maps.level[lv], 0, pairs[pr].x0, pairs[pr].y0, b * M + cam, &bars[slot]);
            }
            ++issued;
        };
#pragma unroll 1
        for (int i = 0; i < kTmaSlots && to_issue; ++i) issue_one();
        float4 acc[kTmaMaxC / 128];
#pragma unroll
        for (int s = 0; s < kTmaMaxC / 128; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        while (to_use) {
            const int pr = __ffs(to_use) - 1;
            to_use &= to_use - 1;
            const int slot = used % kTmaSlots;
            mbar_wait(&bars[slot], (used / kTmaSlots) & 1);
            const TmaPair t = pairs[pr];
            const float hh = 1.f - t.lh, hw = 1.f - t.lw;
            const float *box = ring + static_cast<size_t>(slot) * 4 * C;
#pragma unroll
            for (int s = 0; s < kTmaMaxC / 128; ++s) {
                const int c0 = lane * 4 + 128 * s;
                if (c0 < C) {
                    const float wt = __ldg(wpt + pr * Gr + c0 / gdim);
                    const float a0 = hh * hw * wt, a1 = hh * t.lw * wt, a2 = t.lh * hw * wt, a3 = t.lh * t.lw * wt;
                    const float4 v0 = *reinterpret_cast<const float4 *>(box + c0);
                    const float4 v1 = *reinterpret_cast<const float4 *>(box + C + c0);
                    const float4 v2 = *reinterpret_cast<const float4 *>(box + 2 * C + c0);
                    const float4 v3 = *reinterpret_cast<const float4 *>(box + 3 * C + c0);
                    acc[s].x = fmaf(a0, v0.x, acc[s].x); acc[s].y = fmaf(a0, v0.y, acc[s].y); acc[s].z = fmaf(a0, v0.z, acc[s].z); acc[s].w = fmaf(a0, v0.w, acc[s].w);
                    acc[s].x = fmaf(a1, v1.x, acc[s].x); acc[s].y = fmaf(a1, v1.y, acc[s].y); acc[s].z = fmaf(a1, v1.z, acc[s].z); acc[s].w = fmaf(a1, v1.w, acc[s].w);
                    acc[s].x = fmaf(a2, v2.x, acc[s].x); acc[s].y = fmaf(a2, v2.y, acc[s].y); acc[s].z = fmaf(a2, v2.z, acc[s].z); acc[s].w = fmaf(a2, v2.w, acc[s].w);
                    acc[s].x = fmaf(a3, v3.x, acc[s].x); acc[s].y = fmaf(a3, v3.y, acc[s].y); acc[s].z = fmaf(a3, v3.z, acc[s].z); acc[s].w = fmaf(a3, v3.w, acc[s].w);
                }
            }
            ++used;
            __syncwarp();                                                     // every lane has read the slot ...
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // ... before the async proxy may overwrite it
            if (to_issue) issue_one();
        }
        uses = used;
#pragma unroll
        for (int s = 0; s < kTmaMaxC / 128; ++s) {
            const int c0 = lane * 4 + 128 * s;
            if (c0 < C) *reinterpret_cast<float4 *>(p.out + bp * C + c0) = acc[s];
        }
        __syncwarp();   // the next point overwrites this warp's pair table
    }
}

// Random row gather: every warp sums rows of `row_floats` floats (a multiple of 128) picked by idx[] from a table -- the
// access pattern of the sampling op without its arithmetic.  A warp reads 32 row numbers with one coalesced load and
// keeps eight 512-byte row reads in flight (the op keeps four per visible pair).  With a table that fits L2 this
// measures the L2 -> SM gather bandwidth the op can reach at most.
__global__ void __launch_bounds__(256) gather_probe_kernel(const float *table, const int32_t *idx, long long n, int row_floats, float *sink) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) >> 5, warps = static_cast<long long>(gridDim.x) * 8;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = warp * 32; i < n; i += warps * 32) {
        const long long j = i + lane < n ? i + lane : n - 1;
        const int mine = __ldg(idx + j);
#pragma unroll 1
        for (int k0 = 0; k0 < 32; k0 += 8) {
            const float *row[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) row[k] = table + static_cast<long long>(__shfl_sync(0xffffffffu, mine, k0 + k)) * row_floats;
            for (int c0 = lane * 4; c0 < row_floats; c0 += 128) {
                float4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const float4 *>(row[k] + c0));   // eight loads in flight
#pragma unroll
                for (int k = 0; k < 8; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
            }
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123456.789f) sink[0] = acc.x;   // keeps the loads alive
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn daf_encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

}  // namespace gf

using namespace gf;

extern "C" {

int gf_debug_daf_forward_tma(const gf_daf_desc *d, const float *feat, const int32_t *host_shape, const int32_t *host_start,
                             const float *loc, const float *weights, float *out, gf_stream_t stream_) {
    GF_REQUIRE(d && feat && host_shape && host_start && loc && weights && out, GF_ERR_INVALID_ARG, "daf tma: NULL argument");
    GF_REQUIRE(d->num_embeds % 128 == 0 && d->num_embeds <= kTmaMaxC && d->num_cams * d->num_scale <= 32 &&
                   d->num_scale <= kMaxLevels && (d->num_embeds / d->num_groups) % 4 == 0,
               GF_ERR_UNSUPPORTED, "daf tma: needs C in {128, 256}, cams x levels <= 32, whole-lane groups");
    EncodeTiledFn enc = daf_encode_tiled();
    GF_REQUIRE(enc != nullptr, GF_ERR_CUDA, "daf tma: cuTensorMapEncodeTiled is not available");
    const long long npts = static_cast<long long>(d->batch) * d->num_pts;
    if (npts == 0) return GF_OK;
    DafMaps maps;
    DafTmaParams p;
    p.d = *d; p.loc = loc; p.weights = weights; p.out = out;
    const int C = d->num_embeds;
    for (int l = 0; l < d->num_scale; ++l) {
        const int h = host_shape[2 * l], w = host_shape[2 * l + 1];
        p.lh[l] = h; p.lw[l] = w;
        const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                                    static_cast<cuuint64_t>(d->batch) * d->num_cams};
        const cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(C) * 4 * w,
                                       static_cast<cuuint64_t>(d->num_feat) * C * 4};
        const cuuint32_t box[4] = {static_cast<cuuint32_t>(C), 2, 2, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = enc(&maps.level[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                               const_cast<float *>(feat) + static_cast<long long>(host_start[l]) * C, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GF_REQUIRE(r == CUDA_SUCCESS, GF_ERR_CUDA, "daf tma: cuTensorMapEncodeTiled failed for level %d (%d)", l, static_cast<int>(r));
    }
    for (int l = d->num_scale; l < kMaxLevels; ++l) maps.level[l] = maps.level[0];
    int dev = 0, num_sms = 1;
    GF_CUDA_TRY(cudaGetDevice(&dev));
    GF_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = static_cast<size_t>(kTmaWarps) * kTmaSlots * 4 * C * 4 + kTmaWarps * 32 * sizeof(TmaPair) + kTmaWarps * kTmaSlots * 8;
    GF_CUDA_TRY(cudaFuncSetAttribute(daf_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    long long want = (npts + kTmaWarps - 1) / kTmaWarps;
    const long long cap = static_cast<long long>(num_sms) * 16;
    const int grid = static_cast<int>(want < cap ? want : cap);
    daf_tma_kernel<<<grid, kTmaWarps * 32, smem, static_cast<cudaStream_t>(stream_)>>>(p, maps);
    GF_CUDA_TRY(cudaGetLastError());
    return GF_OK;
}

int gf_debug_gather_probe(const float *table, const int32_t *idx, int64_t n, int32_t row_floats, float *sink, gf_stream_t stream_) {
    GF_REQUIRE(table && idx && sink && n > 0 && row_floats > 0 && row_floats % 128 == 0, GF_ERR_INVALID_ARG, "gather probe: bad argument");
    int dev = 0, num_sms = 1;
    GF_CUDA_TRY(cudaGetDevice(&dev));
    GF_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    gather_probe_kernel<<<num_sms * 8, 256, 0, static_cast<cudaStream_t>(stream_)>>>(table, idx, n, row_floats, sink);
    GF_CUDA_TRY(cudaGetLastError());
    return GF_OK;
}

}  // extern "C"
