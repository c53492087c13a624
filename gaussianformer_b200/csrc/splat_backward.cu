// Splat backward (replaces BACKWARD::preprocess / BACKWARD::renderCUDA,
// model/head/localagg/src/backward.cu:8-20,24-103 and model/head/localagg_prob/src/backward.cu:24-123).
//
// The reference walks each Gaussian's box serially in ONE thread (640 000 iterations for the
// "empty" Gaussian of the solid config).  Here the walk is parallel and needs nothing saved by the
// forward pass:
//
//   voxel_map_kernel      voxel -> point index (largest index wins; the reference's write is a race)
//   backward_small_kernel one warp per Gaussian whose clipped box holds <= kBigBox voxels: lanes
//                         stride over the box, 28 (+1) partial sums in registers, warp-shuffle
//                         reduction, plain stores (deterministic).  Larger boxes are queued.
//   backward_big_kernel   queued Gaussians are split over teams of CTAs sized from the queue
//                         length (the single whole-grid Gaussian gets every CTA; thousands of large
//                         Gaussians get one CTA each); block reduction + one atomicAdd per scalar.
#include "common.cuh"

namespace gf {

constexpr int kBigBox = 2048;
constexpr int kBwdThreads = 256;

struct BwdParams {
    gf_splat_desc d;
    gf_splat_inputs in;
    gf_splat_grads gr;
    int32_t *v2p;      // [H*W*D]
    int32_t *big_list; // [G]
    int32_t *big_count;
    int32_t *canon;    // non-zero after voxel_map_kernel iff N == H*W*D and point n sits in voxel n for all n
    float4 *aux;       // [N] prob only: per-point terms that do not depend on the Gaussian
};

__global__ void __launch_bounds__(256) voxel_map_kernel(const BwdParams p) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.big_count = 0;
    const int H = p.d.H, W = p.d.W, D = p.d.D;
    for (long long n = blockIdx.x * 256ll + threadIdx.x; n < p.d.N; n += 256ll * gridDim.x) {
        int ix, iy, iz;
        if (p.in.points_int) {
            ix = p.in.points_int[3 * n]; iy = p.in.points_int[3 * n + 1]; iz = p.in.points_int[3 * n + 2];
        } else {
            ix = voxel_coord(p.in.pts[3 * n], p.d.pc_min[0], p.d.grid_size);
            iy = voxel_coord(p.in.pts[3 * n + 1], p.d.pc_min[1], p.d.grid_size);
            iz = voxel_coord(p.in.pts[3 * n + 2], p.d.pc_min[2], p.d.grid_size);
        }
        if (ix < 0 || ix >= H || iy < 0 || iy >= W || iz < 0 || iz >= D) { *p.canon = 0; continue; }
        const long long v = (static_cast<long long>(ix) * W + iy) * D + iz;
        if (v != n) *p.canon = 0;   // benign race: every writer stores 0 (the memset left it non-zero)
        atomicMax(p.v2p + v, static_cast<int>(n));
    }
}

// Prob variant: everything in localagg_prob/src/backward.cu:76-100 that depends on the point only is
// folded into one float4 per point, read once per (Gaussian, point) pair instead of 2C+4 scalars:
//   x = sum_k dL/dlogits[n,k] * logits[n,k]     (so that sum_k up_k (sem_k - logits_k) = up.sem - x)
//   y = (1 - bin_logits[n]) * dL/dbin[n]
//   z = dL/ddensity[n]
//   w = 1 / probability[n]  if probability[n] > 1e-9 else 0   (0 switches the logits branch off)
__global__ void __launch_bounds__(256) prob_aux_kernel(const BwdParams p, int C) {
    for (long long n = blockIdx.x * 256ll + threadIdx.x; n < p.d.N; n += 256ll * gridDim.x) {
        float x = 0.f;
        for (int k = 0; k < C; ++k) x = fmaf(__ldg(p.gr.logits_grad + n * C + k), __ldg(p.gr.logits + n * C + k), x);
        const float Z = __ldg(p.gr.probability + n);
        p.aux[n] = make_float4(x, (1.f - __ldg(p.gr.bin_logits + n)) * __ldg(p.gr.bin_logits_grad + n),
                               __ldg(p.gr.density_grad + n), Z > 1e-9f ? __fdiv_rn(1.f, Z) : 0.f);
    }
}

// Per-Gaussian constants and running sums of one thread.
template <int C, bool PROB>
struct GaussAcc {
    // constants
    float mu[3], c6[6], q6[6];  // q6: exponent coefficients pre-scaled by log2(e)
    float opa, det, norm;
    float sem[C];
    // sums
    float sm[3];   // sum w * (A d)
    float so;      // opacity gradient
    float ss[C];   // semantics gradient (without the common per-Gaussian factor for the base variant)
    float sq[6];   // sum w * (dx^2, dy^2, dz^2, dx dy, dy dz, dx dz)
    float sg;      // prob: sum of gamma

    __device__ __forceinline__ void load(const BwdParams &p, int g) {
#pragma unroll
        for (int a = 0; a < 3; ++a) mu[a] = p.in.means[3 * g + a];
        load_cov6(p.d, p.in.cov, g, c6);
        q6[0] = -0.5f * kLog2e * c6[0]; q6[1] = -0.5f * kLog2e * c6[1]; q6[2] = -0.5f * kLog2e * c6[2];
        q6[3] = -kLog2e * c6[3]; q6[4] = -kLog2e * c6[4]; q6[5] = -kLog2e * c6[5];
        opa = p.in.opacities[g];
        det = c6[0] * c6[1] * c6[2] + 2.f * c6[3] * c6[4] * c6[5] - c6[0] * c6[4] * c6[4] -
              c6[1] * c6[5] * c6[5] - c6[2] * c6[3] * c6[3];
        norm = kKappa * sqrtf(det);
#pragma unroll
        for (int k = 0; k < C; ++k) sem[k] = p.in.semantics[static_cast<size_t>(g) * C + k];
#pragma unroll
        for (int a = 0; a < 3; ++a) sm[a] = 0.f;
        so = 0.f; sg = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) ss[k] = 0.f;
#pragma unroll
        for (int a = 0; a < 6; ++a) sq[a] = 0.f;
    }

    // contribution of point n (which lies inside the box)
    __device__ __forceinline__ void visit(const BwdParams &p, long long n) {
        const float dx = mu[0] - __ldg(p.in.pts + 3 * n), dy = mu[1] - __ldg(p.in.pts + 3 * n + 1),
                    dz = mu[2] - __ldg(p.in.pts + 3 * n + 2);
        float t1 = q6[0] * dx;
        t1 = fmaf(q6[3], dy, t1);
        t1 = fmaf(q6[5], dz, t1);
        float t2 = q6[1] * dy;
        t2 = fmaf(q6[4], dz, t2);
        float q = t1 * dx;
        q = fmaf(t2, dy, q);
        q = fmaf(q6[2] * dz, dz, q);
        const float E = ex2_approx(q);
        float w;  // weight of the geometric terms: d(loss)/d(power) * E
        const float2 *up2 = reinterpret_cast<const float2 *>(p.gr.logits_grad + n * C);  // rows are 8B aligned for even C
        float up[C];
        if ((C & 1) == 0) {
#pragma unroll
            for (int k = 0; k < C / 2; ++k) {
                const float2 u = __ldg(up2 + k);
                up[2 * k] = u.x; up[2 * k + 1] = u.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < C; ++k) up[k] = __ldg(p.gr.logits_grad + n * C + k);
        }
        if (!PROB) {
            // backward.cu:72-87 with t = sum_k sem_k * E * up_k
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const float eu = E * up[k];
                ss[k] += eu;               // * opa at the end
                t = fmaf(sem[k], eu, t);
            }
            so += t;
            w = opa * t;
        } else {
            // localagg_prob/src/backward.cu:76-100 with the point-only terms pre-folded (prob_aux_kernel)
            const float4 ax = __ldg(p.aux + n);
            const float Pt = norm * E;
            float pi = 0.f;
            if (ax.w > 0.f) {
                float u = -ax.x;
                const float sfac = Pt * opa * ax.w;
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    u = fmaf(up[k], sem[k], u);
                    ss[k] = fmaf(up[k], sfac, ss[k]);
                }
                pi = u * opa * ax.w;
                so = fmaf(u * Pt, ax.w, so);
            }
            const float eps = pi * norm + __fdiv_rn(ax.y, 1.f - E + 1e-9f) + ax.z;
            sg += __fdiv_rn(pi * Pt * 0.5f, det);
            w = eps * E;
        }
        sm[0] = fmaf(w, c6[0] * dx + c6[3] * dy + c6[5] * dz, sm[0]);
        sm[1] = fmaf(w, c6[3] * dx + c6[1] * dy + c6[4] * dz, sm[1]);
        sm[2] = fmaf(w, c6[5] * dx + c6[4] * dy + c6[2] * dz, sm[2]);
        sq[0] = fmaf(w * dx, dx, sq[0]);
        sq[1] = fmaf(w * dy, dy, sq[1]);
        sq[2] = fmaf(w * dz, dz, sq[2]);
        sq[3] = fmaf(w * dx, dy, sq[3]);
        sq[4] = fmaf(w * dy, dz, sq[4]);
        sq[5] = fmaf(w * dx, dz, sq[5]);
    }

    // number of scalars that travel through reductions: 3 + 1 + C + 6 (+1)
    static constexpr int kVals = 3 + 1 + C + 6 + (PROB ? 1 : 0);
    __device__ __forceinline__ float &val(int i) {
        if (i < 3) return sm[i];
        if (i == 3) return so;
        if (i < 4 + C) return ss[i - 4];
        if (i < 10 + C) return sq[i - 4 - C];
        return sg;
    }

    // final gradients from the (fully reduced) sums: means[3], opacity, sem[C], cov[6]
    __device__ __forceinline__ float grad_means(int a) const { return -sm[a]; }
    __device__ __forceinline__ float grad_opa() const { return so; }
    __device__ __forceinline__ float grad_sem(int k) const { return PROB ? ss[k] : opa * ss[k]; }
    __device__ __forceinline__ float grad_cov(int i) const {
        float gq = (i < 3) ? -0.5f * sq[i] : -sq[i];
        if (PROB) {
            const float a = c6[0], b = c6[1], c = c6[2], d = c6[3], e = c6[4], f = c6[5];
            const float m[6] = {b * c - e * e, a * c - f * f, a * b - d * d,
                                2.f * (e * f - c * d), 2.f * (d * f - a * e), 2.f * (d * e - b * f)};
            gq = fmaf(sg, m[i], gq);
        }
        return gq;
    }
};

// exact floor(i / n) for 0 <= i < 2^20 given inv_n = fl(1/n): (i+0.5)/n is at least 0.5/n away
// from an integer while the two roundings perturb it by < 2^20/n * 2^-23 = 0.125/n.
__device__ __forceinline__ int div_small(int i, float inv_n) {
    return __float2int_rz((static_cast<float>(i) + 0.5f) * inv_n);
}

struct BoxWalk {
    int lo[3], nx, ny, nz;
    long long vol;
    float inv_nz, inv_ny;
    bool small_idx;
    __device__ __forceinline__ void init(const int l[3], const int h[3], bool empty) {
        lo[0] = l[0]; lo[1] = l[1]; lo[2] = l[2];
        nx = empty ? 0 : h[0] - l[0] + 1;
        ny = empty ? 0 : h[1] - l[1] + 1;
        nz = empty ? 0 : h[2] - l[2] + 1;
        vol = static_cast<long long>(nx) * ny * nz;
        inv_nz = nz ? __fdiv_rn(1.f, static_cast<float>(nz)) : 0.f;
        inv_ny = ny ? __fdiv_rn(1.f, static_cast<float>(ny)) : 0.f;
        small_idx = vol < (1ll << 20);
    }
    // flat box index -> voxel index of the grid
    __device__ __forceinline__ long long voxel(long long i, int W, int D) const {
        int ix, iy, iz;
        if (small_idx) {
            const int ii = static_cast<int>(i);
            const int t = div_small(ii, inv_nz);
            iz = ii - t * nz;
            ix = div_small(t, inv_ny);
            iy = t - ix * ny;
        } else {
            iz = static_cast<int>(i % nz);
            const long long t = i / nz;
            iy = static_cast<int>(t % ny);
            ix = static_cast<int>(t / ny);
        }
        return (static_cast<long long>(lo[0] + ix) * W + (lo[1] + iy)) * D + (lo[2] + iz);
    }
};

template <int C, bool PROB>
__device__ __forceinline__ void store_grads(const BwdParams &p, int g, GaussAcc<C, PROB> &acc, int lane, bool atomic) {
    // lanes 0..2 means, 3 opacity, 4..9 cov, then semantics over lanes (C <= 32)
    float v = 0.f;
    float *dst = nullptr;
    if (lane < 3) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (lane == a) v = acc.grad_means(a);
        dst = p.gr.means_grad + 3 * g + lane;
    } else if (lane == 3) {
        v = acc.grad_opa();
        dst = p.gr.opacity_grad + g;
    } else if (lane < 10) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (lane == 4 + i) v = acc.grad_cov(i);
        dst = p.gr.cov_grad + 6 * g + (lane - 4);
    }
    if (dst) {
        if (atomic) atomicAdd(dst, v); else *dst = v;
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k)
        if (lane == k) s = acc.grad_sem(k);
    if (lane < C) {
        float *d2 = p.gr.semantics_grad + static_cast<size_t>(g) * C + lane;
        if (atomic) atomicAdd(d2, s); else *d2 = s;
    }
}

template <int C, bool PROB>
__global__ void __launch_bounds__(kBwdThreads) backward_small_kernel(const BwdParams p) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (kBwdThreads / 32) + (threadIdx.x >> 5);
    if (g >= p.d.G) return;
    GaussAcc<C, PROB> acc;
    acc.load(p, g);
    int lo[3], hi[3];
    uint32_t err = 0;
    const bool empty = gaussian_box(p.d, p.in, g, acc.mu, lo, hi, err);
    BoxWalk box;
    box.init(lo, hi, empty);
    const bool big = box.vol > kBigBox;
    const bool canon = *p.canon != 0;   // then voxel index == point index and the map need not be read
    if (!big) {
        for (int i = lane; i < static_cast<int>(box.vol); i += 32) {
            const long long v = box.voxel(i, p.d.W, p.d.D);
            const long long n = canon ? v : __ldg(p.v2p + v);
            if (n >= 0) acc.visit(p, n);
        }
#pragma unroll
        for (int i = 0; i < GaussAcc<C, PROB>::kVals; ++i) {
            float v = acc.val(i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc.val(i) = v;
        }
    } else if (lane == 0) {
        p.big_list[atomicAdd(p.big_count, 1)] = g;
    }
    // small boxes: final values; queued boxes: zeros (the big kernel accumulates atomically)
    store_grads<C, PROB>(p, g, acc, lane, false);
}

template <int C, bool PROB>
__global__ void __launch_bounds__(kBwdThreads) backward_big_kernel(const BwdParams p) {
    const int nbig = *p.big_count;
    if (nbig == 0) return;
    const int NB = gridDim.x;
    const int T = max(1, NB / nbig);        // CTAs per Gaussian
    const int nteams = NB / T;
    const int team = blockIdx.x / T, part = blockIdx.x % T;
    if (team >= nteams) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ float s_part[kBwdThreads / 32][GaussAcc<C, PROB>::kVals];
    for (int bi = team; bi < nbig; bi += nteams) {
        const int g = p.big_list[bi];
        GaussAcc<C, PROB> acc;
        acc.load(p, g);
        int lo[3], hi[3];
        uint32_t err = 0;
        const bool empty = gaussian_box(p.d, p.in, g, acc.mu, lo, hi, err);
        BoxWalk box;
        box.init(lo, hi, empty);
        const long long beg = box.vol * part / T, end = box.vol * (part + 1) / T;
        const bool canon = *p.canon != 0;
        for (long long i = beg + threadIdx.x; i < end; i += kBwdThreads) {
            const long long v = box.voxel(i, p.d.W, p.d.D);
            const long long n = canon ? v : __ldg(p.v2p + v);
            if (n >= 0) acc.visit(p, n);
        }
#pragma unroll
        for (int i = 0; i < GaussAcc<C, PROB>::kVals; ++i) {
            float v = acc.val(i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_part[warp][i] = v;
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int i = 0; i < GaussAcc<C, PROB>::kVals; ++i) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < kBwdThreads / 32; ++w) v += s_part[w][i];
                acc.val(i) = v;
            }
            store_grads<C, PROB>(p, g, acc, lane, true);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct BwdWorkspace {
    int32_t *v2p, *big_list, *big_count, *canon;
    float4 *aux;
    size_t bytes;
};

static size_t align_up_b(size_t v, size_t a) { return (v + a - 1) / a * a; }

void plan_backward_workspace(const gf_splat_desc &d, void *base, BwdWorkspace *ws) {
    size_t off = 0;
    char *b = static_cast<char *>(base);
    auto take = [&](size_t bytes) {
        char *p = b ? b + off : nullptr;
        off = align_up_b(off + bytes, 256);
        return p;
    };
    ws->big_count = reinterpret_cast<int32_t *>(take(64));
    ws->canon = reinterpret_cast<int32_t *>(take(256));   // directly in front of v2p: one memset covers both
    ws->v2p = reinterpret_cast<int32_t *>(take(size_t(d.H) * d.W * d.D * 4));
    ws->big_list = reinterpret_cast<int32_t *>(take(size_t(d.G) * 4));
    ws->aux = reinterpret_cast<float4 *>(take(d.variant == GF_SPLAT_PROB ? size_t(d.N) * 16 : 0));
    ws->bytes = off;
}

size_t backward_workspace_bytes(const gf_splat_desc &d) {
    BwdWorkspace ws;
    plan_backward_workspace(d, nullptr, &ws);
    return ws.bytes;
}

template <int C, bool PROB>
static int launch_backward_t(const BwdParams &bp, int num_sms, cudaStream_t stream) {
    const gf_splat_desc &d = bp.d;
    // canon word (non-zero = "still canonical") and the voxel->point map (-1 = empty) in one fill
    GF_CUDA_TRY(cudaMemsetAsync(bp.canon, 0xFF, 256 + size_t(d.H) * d.W * d.D * 4, stream));
    if (static_cast<long long>(d.N) != static_cast<long long>(d.H) * d.W * d.D)
        GF_CUDA_TRY(cudaMemsetAsync(bp.canon, 0, 4, stream));
    const long long want = (static_cast<long long>(d.N) + 255) / 256;
    const int grid0 = static_cast<int>(want < 16ll * num_sms ? (want > 0 ? want : 1) : 16ll * num_sms);
    voxel_map_kernel<<<grid0, 256, 0, stream>>>(bp);
    GF_CUDA_TRY(cudaGetLastError());
    if (PROB) {
        prob_aux_kernel<<<grid0, 256, 0, stream>>>(bp, C);
        GF_CUDA_TRY(cudaGetLastError());
    }
    const int per_cta = kBwdThreads / 32;
    backward_small_kernel<C, PROB><<<(d.G + per_cta - 1) / per_cta, kBwdThreads, 0, stream>>>(bp);
    GF_CUDA_TRY(cudaGetLastError());
    backward_big_kernel<C, PROB><<<num_sms * 4, kBwdThreads, 0, stream>>>(bp);
    GF_CUDA_TRY(cudaGetLastError());
    return GF_OK;
}

int launch_backward(const gf_splat_desc &d, const gf_splat_inputs &in, const gf_splat_grads &gr, void *workspace,
                    int num_sms, cudaStream_t stream) {
    BwdWorkspace ws;
    plan_backward_workspace(d, workspace, &ws);
    BwdParams bp;
    bp.d = d;
    bp.in = in;
    bp.gr = gr;
    bp.v2p = ws.v2p;
    bp.big_list = ws.big_list;
    bp.big_count = ws.big_count;
    bp.canon = ws.canon;
    bp.aux = ws.aux;
    const bool prob = d.variant == GF_SPLAT_PROB;
#define GF_CASE(CC)                                                       \
    case CC:                                                              \
        return prob ? launch_backward_t<CC, true>(bp, num_sms, stream)    \
                    : launch_backward_t<CC, false>(bp, num_sms, stream);
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
