// Shared device/host helpers for the sm_100a kernels (splat + deformable aggregation).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "gf_b200.h"

namespace gf {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define GF_CUDA_TRY(expr)                                         \
    do {                                                          \
        cudaError_t _e = (expr);                                  \
        if (_e != cudaSuccess) return ::gf::cuda_fail(_e, #expr); \
    } while (0)

#define GF_REQUIRE(cond, code, ...)      \
    do {                                 \
        if (!(cond)) {                   \
            ::gf::set_error(__VA_ARGS__); \
            return (code);               \
        }                                \
    } while (0)

// ---------------------------------------------------------------------------------------------
// geometry of the splat pipeline
// ---------------------------------------------------------------------------------------------
constexpr int kBinX = 8;        // render CTA footprint in columns: 8 (x) by 4 (y) ...
constexpr int kBinY = 4;
constexpr int kBinZ = 16;       // ... by 16 voxels in z (one z-chunk)
constexpr int kRenderThreads = 128;
constexpr int kVox = 4;         // voxels per thread (a z-quad)
constexpr int kChunk = 32;      // Gaussian records staged per TMA batch
constexpr int kSeg = 512;       // bin-list entries resolved per segment
constexpr int kGeomFloats = 12; // mu(3) k(1) c6(6) pad(2)

__host__ __device__ constexpr int round_up(int v, int m) { return (v + m - 1) / m * m; }
// one record = geometry + class vector, padded to a whole number of 128-byte lines.  Forward records (render kernels), as
// 16-byte chunks: 0 = mu(3), amplitude (base variant: z-level mask of the box) | 1 = exponent coefficients a b c d |
// 2 = e f, classes 16 17 | 3..6 = classes 0..15 | 7 = classes 18 19, pad(2).  Classes 16 and 17 share the chunk of the
// last two coefficients so that an 18-class step is six 16-byte shared loads instead of five and two 8-byte ones.  The raw
// records of the backward keep the plain order: chunk 2 = e f 0 0, chunks 3..7 = classes 0..19.
__host__ __device__ constexpr int rec_floats(int C) { return round_up(kGeomFloats + round_up(C, 4), 32); }

// Integer box of one Gaussian, inclusive bounds packed lo | hi << 16 per axis; w = 1 when the
// clipped box is empty.  (reference: getRect, model/head/localagg/src/auxiliary.h:8-20)
struct __align__(16) PackedBox {
    uint32_t x, y, z, empty;
};

struct SplatWorkspace {
    // all pointers carved out of the caller's workspace by carve_forward_workspace()
    uint32_t *flags;       // [16]  status word(s)
    uint32_t *pack_flags;  // [pack_ctas] per-CTA error bits of the pack kernel
    float *records;        // [G, rec_floats(C)]
    PackedBox *boxes;      // [G]
    uint32_t *masks;       // [nsuper, nwords] bit g of supertile s: box(g) overlaps s
    int32_t *lists;        // [nsuper, G] ascending Gaussian indices per supertile
    int32_t *counts;       // [nsuper]
    int st;                // supertile edge (columns)
    int nsx, nsy, nsuper;  // supertile grid
    int nwords;            // ceil(G/32)
    int pack_ctas;         // pack CTAs per sample
    int batch;             // samples; every region above except `flags` holds `batch` consecutive per-sample blocks
    size_t bytes;
};

__host__ __device__ inline int batch_of(const gf_splat_desc &d) { return d.batch > 0 ? d.batch : 1; }

constexpr int kPackThreads = 64;   // small CTAs: ~400 of them cover the 148 SMs several times over

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// Launches `kernel` so that it may become resident while the preceding kernel of the stream is still
// running (programmatic stream serialization).  The kernel must call pdl_wait() before it touches
// anything the preceding kernel writes.  GF_B200_PDL=0 falls back to plain stream order.
bool pdl_enabled();
template <class... Params>
cudaError_t launch_chained(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const Params &...p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, p...);
}


__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x4000;\n"   // suspend-time hint (ns): sleep, don't poll
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// Bounded spin: a protocol bug must surface as a launch failure (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
// 1-D bulk async copy global -> shared through the TMA unit, completion counted on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Programmatic dependent launch (the three kernels of a forward call are chained with it, see
// launch_chained): launch_dependents lets the next kernel of the stream become resident while this grid
// is still running; wait blocks until the preceding grid has completed and its writes are visible.
// Both are no-ops for a kernel launched without the attribute / without a dependent.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ((v - origin) / cell) truncated toward zero, with the exact fp32 operation order of
// `((x - pc_min) / grid_size).to(torch.int)` (local_aggregate/__init__.py:137,139).
__device__ __forceinline__ int voxel_coord(float v, float origin, float cell) {
    return static_cast<int>(__fdiv_rn(__fsub_rn(v, origin), cell));
}

// Integer box of Gaussian g (inclusive bounds, clipped to the grid) and the error bits the
// reference asserts on.  Fuses the Python host preparation (trunc voxel index of the mean, ceil
// radius; local_aggregate/__init__.py:139-142, prob: :151-153, prob_fast: :151) with getRect
// (src/auxiliary.h:8-20).  Returns true when the clipped box is empty.
__device__ __forceinline__ bool gaussian_box(const gf_splat_desc &d, const gf_splat_inputs &in, int g,
                                             const float mu[3], int lo[3], int hi[3], uint32_t &err) {
    const int dims[3] = {d.H, d.W, d.D};
    int m[3], r[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
        m[a] = in.means_int ? in.means_int[3 * g + a] : voxel_coord(mu[a], d.pc_min[a], d.grid_size);
    if (in.radii) {
#pragma unroll
        for (int a = 0; a < 3; ++a) r[a] = (d.radii_axes == 3) ? in.radii[3 * g + a] : in.radii[g];
    } else {
        const float s[3] = {in.scales[3 * g], in.scales[3 * g + 1], in.scales[3 * g + 2]};
        if (d.radii_axes == 3) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
                r[a] = static_cast<int>(ceilf(__fdiv_rn(__fmul_rn(s[a], d.scale_multiplier), d.grid_size)));
        } else {
            const float smax = fmaxf(s[0], fmaxf(s[1], s[2]));
            r[0] = r[1] = r[2] = static_cast<int>(ceilf(__fdiv_rn(__fmul_rn(smax, d.scale_multiplier), d.grid_size)));
        }
        if (d.radii_min > 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) r[a] = max(r[a], d.radii_min);
        }
    }
    bool empty = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (m[a] < 0 || m[a] >= dims[a]) err |= GF_FLAG_MEAN_OUT_OF_GRID;
        if (r[a] < 1) err |= GF_FLAG_RADIUS_LT_1;
        // inclusive form of [min(dim,max(0,m-r)), min(dim,max(0,m+r+1)))
        const long long l = static_cast<long long>(m[a]) - r[a];
        const long long h = static_cast<long long>(m[a]) + r[a];
        lo[a] = static_cast<int>(l < 0 ? 0 : l);
        hi[a] = static_cast<int>(h > dims[a] - 1 ? dims[a] - 1 : h);
        if (l > dims[a] - 1 || h < 0 || lo[a] > hi[a]) empty = true;
    }
    return empty;
}

// ---- batch: the tensors of sample b (dense strides; NULL stays NULL; pts may be shared by the batch) ----------------
template <class T>
__device__ __forceinline__ T *adv(T *p, long long off) { return p ? p + off : p; }

__device__ __forceinline__ gf_splat_inputs sample_inputs(const gf_splat_desc &d, const gf_splat_inputs &in, int b) {
    gf_splat_inputs o;
    const long long G = d.G, N = d.N;
    const long long np = d.pts_shared ? 0 : b * N * 3;
    o.pts = adv(in.pts, np);
    o.points_int = adv(in.points_int, np);
    o.means = adv(in.means, b * G * 3);
    o.means_int = adv(in.means_int, b * G * 3);
    o.opacities = adv(in.opacities, b * G);
    o.semantics = adv(in.semantics, b * G * d.C);
    o.cov = adv(in.cov, b * G * d.cov_stride);
    o.radii = adv(in.radii, b * G * (d.radii_axes == 3 ? 3 : 1));
    o.scales = adv(in.scales, b * G * 3);
    o.rotations = adv(in.rotations, b * G * 4);
    return o;
}

__device__ __forceinline__ gf_splat_outputs sample_outputs(const gf_splat_desc &d, const gf_splat_outputs &out, int b, int ce_rows) {
    gf_splat_outputs o;
    const long long N = d.N;
    o.logits = adv(out.logits, b * N * d.C);
    o.bin_logits = adv(out.bin_logits, b * N);
    o.density = adv(out.density, b * N);
    o.probability = adv(out.probability, b * N);
    o.argmax = adv(out.argmax, b * N);
    o.logits_cn = adv(out.logits_cn, b * N * d.C);
    o.labels = adv(out.labels, b * N);
    o.class_weights = out.class_weights;
    o.ce_partials = adv(out.ce_partials, 2ll * b * ce_rows);
    return o;
}

__device__ __forceinline__ gf_splat_grads sample_grads(const gf_splat_desc &d, const gf_splat_grads &gr, int b) {
    gf_splat_grads o;
    const long long G = d.G, N = d.N;
    o.logits_grad = adv(gr.logits_grad, b * N * d.C);
    o.bin_logits_grad = adv(gr.bin_logits_grad, b * N);
    o.density_grad = adv(gr.density_grad, b * N);
    o.logits = adv(gr.logits, b * N * d.C);
    o.bin_logits = adv(gr.bin_logits, b * N);
    o.probability = adv(gr.probability, b * N);
    o.means_grad = adv(gr.means_grad, b * G * 3);
    o.opacity_grad = adv(gr.opacity_grad, b * G);
    o.semantics_grad = adv(gr.semantics_grad, b * G * d.C);
    o.cov_grad = adv(gr.cov_grad, b * G * d.cov_stride);
    o.scales_grad = adv(gr.scales_grad, b * G * 3);
    o.rotations_grad = adv(gr.rotations_grad, b * G * 4);
    return o;
}

// Rotation matrix of the NORMALISED quaternion (w,x,y,z) -- get_rotation_matrix, model/utils/utils.py:20-66
// (F.normalize: q / max(|q|, 1e-12)).  R[k][l], row-major.
__device__ __forceinline__ void quat_rotation(const float q[4], float R[3][3], float &inv_norm) {
    const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    inv_norm = __fdiv_rn(1.f, fmaxf(n, 1e-12f));
    const float w = q[0] * inv_norm, x = q[1] * inv_norm, y = q[2] * inv_norm, z = q[3] * inv_norm;
    R[0][0] = w * w + x * x - y * y - z * z; R[0][1] = 2.f * (x * y - w * z);         R[0][2] = 2.f * (x * z + w * y);
    R[1][0] = 2.f * (x * y + w * z);         R[1][1] = w * w - x * x + y * y - z * z; R[1][2] = 2.f * (y * z - w * x);
    R[2][0] = 2.f * (x * z - w * y);         R[2][1] = 2.f * (y * z + w * x);         R[2][2] = w * w - x * x - y * y + z * z;
}

// Sigma^-1 = R^T diag(1/s^2) R as (xx,yy,zz,xy,yz,xz): the closed form of GaussianHead.prepare_gaussian_args'
// Cov = (S R)^T (S R); Cov.cpu().inverse() (model/head/gaussian_head.py:111-119), R orthogonal.
__device__ __forceinline__ void cov6_from_srt(const float s[3], const float q[4], float c6[6]) {
    float R[3][3], inv_norm;
    quat_rotation(q, R, inv_norm);
    const float d0 = __fdiv_rn(1.f, s[0] * s[0]), d1 = __fdiv_rn(1.f, s[1] * s[1]), d2 = __fdiv_rn(1.f, s[2] * s[2]);
    auto a = [&](int i, int j) { return R[0][i] * d0 * R[0][j] + R[1][i] * d1 * R[1][j] + R[2][i] * d2 * R[2][j]; };
    c6[0] = a(0, 0); c6[1] = a(1, 1); c6[2] = a(2, 2); c6[3] = a(0, 1); c6[4] = a(1, 2); c6[5] = a(0, 2);
}

// the six inverse-covariance entries (xx,yy,zz,xy,yz,xz) of Gaussian g (in: the tensors of ONE sample)
__device__ __forceinline__ void load_cov6_in(const gf_splat_desc &d, const gf_splat_inputs &in, int g, float c6[6]) {
    if (in.cov == nullptr) {
        const float s[3] = {__ldg(in.scales + 3 * g), __ldg(in.scales + 3 * g + 1), __ldg(in.scales + 3 * g + 2)};
        const float q[4] = {__ldg(in.rotations + 4 * g), __ldg(in.rotations + 4 * g + 1), __ldg(in.rotations + 4 * g + 2),
                            __ldg(in.rotations + 4 * g + 3)};
        cov6_from_srt(s, q, c6);
        return;
    }
    const float *cv = in.cov + static_cast<size_t>(g) * d.cov_stride;
    if (d.cov_stride == 9) {  // flat entries [0,4,8,1,5,2] of the row-major 3x3 (__init__.py:143)
        c6[0] = __ldg(cv); c6[1] = __ldg(cv + 4); c6[2] = __ldg(cv + 8); c6[3] = __ldg(cv + 1); c6[4] = __ldg(cv + 5); c6[5] = __ldg(cv + 2);
    } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) c6[i] = __ldg(cv + i);
    }
}

// the six inverse-covariance entries (xx,yy,zz,xy,yz,xz) of Gaussian g
__device__ __forceinline__ void load_cov6(const gf_splat_desc &d, const float *cov, int g, float c6[6]) {
    const float *cv = cov + static_cast<size_t>(g) * d.cov_stride;
    if (d.cov_stride == 9) {  // flat entries [0,4,8,1,5,2] of the row-major 3x3 (__init__.py:143)
        c6[0] = cv[0]; c6[1] = cv[4]; c6[2] = cv[8]; c6[3] = cv[1]; c6[4] = cv[5]; c6[5] = cv[2];
    } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) c6[i] = cv[i];
    }
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kKappa = 0.06349363593424097f;  // powf(2 * 3.1415926535f, -1.5f), localagg_prob/src/forward.cu:78

#endif  // __CUDACC__

}  // namespace gf
