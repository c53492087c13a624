// Splat backward (replaces BACKWARD::preprocess / BACKWARD::renderCUDA,
// model/head/localagg/src/backward.cu:8-20,24-103 and model/head/localagg_prob/src/backward.cu:24-123).
//
// The reference walks each Gaussian's box serially in ONE thread (640 000 iterations for the
// "empty" Gaussian of the solid config).  Here the walk is parallel and needs nothing saved by the
// forward pass:
//
//   voxel_map_kernel      voxel -> point index (largest index wins; the reference's write is a race)
//   backward_small_kernel one warp per Gaussian whose clipped box holds <= kBigBox voxels: lanes
//                         stride over the box (division-free walk, loads one pair ahead, packed
//                         fp32 pairs for the class sums), 28 (+1) partial sums in registers, one
//                         transposing warp reduction, plain stores (deterministic).  Larger boxes
//                         are queued.
//   backward_big_kernel   the queued boxes, measured in chunks of kBigChunk box voxels, form one long
//                         line of work; every CTA takes an equal contiguous span of it (balanced by
//                         volume: the whole-grid Gaussian is spread over the whole GPU, while a span
//                         of small queued boxes is walked Gaussian by Gaussian); block reduction +
//                         one atomicAdd per scalar and (CTA, Gaussian) segment.
#include "common.cuh"

namespace gf {

constexpr int kBigBox = 2048;
// CTA shapes of the two pair kernels: the prob variant carries more per-thread state, so it trades resident
// warps for registers (128 threads x 3 CTAs -> 168 registers) instead of spilling at 128.
#ifndef GF_BWD_PROB_THREADS
#define GF_BWD_PROB_THREADS 128
#define GF_BWD_PROB_CTAS 3
#endif
constexpr int bwd_threads(bool prob) { return prob ? GF_BWD_PROB_THREADS : 256; }
constexpr int bwd_ctas(bool prob) { return prob ? GF_BWD_PROB_CTAS : 2; }

struct BwdParams {
    gf_splat_desc d;
    gf_splat_inputs in;
    gf_splat_grads gr;
    int32_t *v2p;      // [H*W*D]
    uint2 *big;        // [G] queue of the boxes larger than kBigBox: (Gaussian, first chunk), ascending in both
    unsigned long long *big_ctr;   // queue length << 40 | chunks queued (one atomic keeps the two in step)
    int chunk;         // box voxels per chunk
    int32_t *canon;    // non-zero after voxel_map_kernel iff N == H*W*D and point n sits in voxel n for all n
    float4 *aux;       // [N] prob only: per-point terms that do not depend on the Gaussian
    size_t cv_block;   // bytes between the (canon, v2p) blocks of consecutive samples
};

// The launch parameters narrowed to sample b of the batch (the kernels run with blockIdx.y = sample).
__device__ __forceinline__ BwdParams sample_bwd(const BwdParams &p, int b) {
    BwdParams q = p;
    q.in = sample_inputs(p.d, p.in, b);
    q.gr = sample_grads(p.d, p.gr, b);
    q.canon = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(p.canon) + b * p.cv_block);
    q.v2p = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(p.v2p) + b * p.cv_block);
    q.big = p.big + static_cast<size_t>(b) * p.d.G;
    q.big_ctr = p.big_ctr + b;
    q.aux = adv(p.aux, static_cast<long long>(b) * p.d.N);
    return q;
}

__global__ void __launch_bounds__(256) voxel_map_kernel(const BwdParams pb) {
    pdl_launch_dependents();   // the pair kernels may become resident; they wait before using the map
    const BwdParams p = sample_bwd(pb, blockIdx.y);
    const int H = p.d.H, W = p.d.W, D = p.d.D;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *p.big_ctr = 0ull;
        if (static_cast<long long>(p.d.N) != static_cast<long long>(H) * W * D) *p.canon = 0;   // can only be generic
    }
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
__global__ void __launch_bounds__(256) prob_aux_kernel(const BwdParams pb, int C) {
    pdl_launch_dependents();
    const BwdParams p = sample_bwd(pb, blockIdx.y);
    for (long long n = blockIdx.x * 256ll + threadIdx.x; n < p.d.N; n += 256ll * gridDim.x) {
        float x = 0.f;
        for (int k = 0; k < C; ++k) x = fmaf(__ldg(p.gr.logits_grad + n * C + k), __ldg(p.gr.logits + n * C + k), x);
        const float Z = __ldg(p.gr.probability + n);
        p.aux[n] = make_float4(x, (1.f - __ldg(p.gr.bin_logits + n)) * __ldg(p.gr.bin_logits_grad + n),
                               __ldg(p.gr.density_grad + n), Z > 1e-9f ? __fdiv_rn(1.f, Z) : 0.f);
    }
    // runs beside voxel_map_kernel (nothing above depends on it) but must not COMPLETE before it: the pair
    // kernel's wait only covers the grid launched immediately before it
    pdl_wait();
}

// One (Gaussian, point) pair's point-side data, fetched one iteration ahead of its use.
template <int C, bool PROB>
struct PairData {
    static constexpr int CP2 = (C + 1) / 2;
    static constexpr int kVec = (C + 2 + 3) / 4;   // float4 loads that cover a row starting 0 or 8 bytes into the first
    float px, py, pz;
    // dL/dlogits[n, :].  Wide mode (even C, 16-byte aligned base): the kVec aligned float4 that contain the
    // row, which starts at float `2*shift` of them -- a warp's 32 rows of 4C bytes then cost kVec
    // L1 passes instead of C/2.  Otherwise: packed pairs (zero padded for odd C) in raw[k/2].
    float4 raw[kVec];
    bool shift;
    float4 ax;                // prob: prob_aux_kernel's per-point terms
    bool ok;
    template <bool WIDE>
    __device__ __forceinline__ float2 up(int k) const {   // k is a compile-time constant after unrolling
        if (WIDE) {
            const float2 a = (k & 1) ? make_float2(raw[k >> 1].z, raw[k >> 1].w) : make_float2(raw[k >> 1].x, raw[k >> 1].y);
            const int k1 = k + 1;
            const float2 b = (k1 & 1) ? make_float2(raw[k1 >> 1].z, raw[k1 >> 1].w) : make_float2(raw[k1 >> 1].x, raw[k1 >> 1].y);
            return shift ? b : a;
        }
        return (k & 1) ? make_float2(raw[k >> 1].z, raw[k >> 1].w) : make_float2(raw[k >> 1].x, raw[k >> 1].y);
    }
};

// Per-Gaussian constants and running sums of one thread.  The sums are kept in the form that needs the
// fewest instructions per pair; the linear maps to the actual gradients are applied once per Gaussian
// (finish()):  d(mean) = -A * sum(w d),  d(cov) = -(1/2 | 1) * sum(w d d^T) (+ det terms for prob).
template <int C, bool PROB>
struct GaussAcc {
    static constexpr int CP2 = (C + 1) / 2;
    // constants
    float mu[3], q6[6];       // q6: exponent coefficients pre-scaled by log2(e)
    float opa, norm, inv2det; // prob: kappa*sqrt(det), 0.5/det
    float2 sem[CP2];
    // sums
    float sd[3];      // sum w * d
    float so;         // opacity gradient
    float2 ss[CP2];   // semantics gradient (base: without the common factor opa)
    float sq[6];      // sum w * (dx^2, dy^2, dz^2, dx dy, dy dz, dx dz)
    float sg;         // prob: sum of gamma

    __device__ __forceinline__ void load(const BwdParams &p, int g) {
#pragma unroll
        for (int a = 0; a < 3; ++a) mu[a] = p.in.means[3 * g + a];
        float c6[6];
        load_cov6_in(p.d, p.in, g, c6);
        q6[0] = -0.5f * kLog2e * c6[0]; q6[1] = -0.5f * kLog2e * c6[1]; q6[2] = -0.5f * kLog2e * c6[2];
        q6[3] = -kLog2e * c6[3]; q6[4] = -kLog2e * c6[4]; q6[5] = -kLog2e * c6[5];
        opa = p.in.opacities[g];
        if (PROB) {
            const float det = c6[0] * c6[1] * c6[2] + 2.f * c6[3] * c6[4] * c6[5] - c6[0] * c6[4] * c6[4] -
                              c6[1] * c6[5] * c6[5] - c6[2] * c6[3] * c6[3];
            norm = kKappa * sqrtf(det);
            inv2det = __fdiv_rn(0.5f, det);
        } else {
            norm = 0.f; inv2det = 0.f;
        }
#pragma unroll
        for (int k = 0; k < CP2; ++k) {
            sem[k].x = p.in.semantics[static_cast<size_t>(g) * C + 2 * k];
            sem[k].y = (2 * k + 1 < C) ? p.in.semantics[static_cast<size_t>(g) * C + 2 * k + 1] : 0.f;
            ss[k] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) sd[a] = 0.f;
        so = 0.f; sg = 0.f;
#pragma unroll
        for (int a = 0; a < 6; ++a) sq[a] = 0.f;
    }

    // issue the loads of point n (n < 0: no point in that voxel)
    template <bool WIDE>
    __device__ __forceinline__ void fetch(const BwdParams &p, long long n, PairData<C, PROB> &o) const {
        o.ok = n >= 0;
        if (!o.ok) return;
        o.px = __ldg(p.in.pts + 3 * n); o.py = __ldg(p.in.pts + 3 * n + 1); o.pz = __ldg(p.in.pts + 3 * n + 2);
        const float *row = p.gr.logits_grad + n * C;
        if (WIDE) {
            // 4C bytes starting 8-byte aligned: the enclosing 16-byte aligned window of kVec float4
            o.shift = (reinterpret_cast<uintptr_t>(row) & 8) != 0;
            const float4 *w = reinterpret_cast<const float4 *>(row - (o.shift ? 2 : 0));
#pragma unroll
            for (int k = 0; k < PairData<C, PROB>::kVec - 1; ++k) o.raw[k] = __ldg(w + k);
            constexpr int last = PairData<C, PROB>::kVec - 1;
            // the last float4 of an unshifted row may reach past the row; past the tensor for the last row
            if ((C % 4) == 2 && !o.shift && n + 1 >= p.d.N) {
                const float2 t = __ldg(reinterpret_cast<const float2 *>(w + last));
                o.raw[last] = make_float4(t.x, t.y, 0.f, 0.f);
            } else if ((C % 4) == 0 && !o.shift) {
                o.raw[last] = make_float4(0.f, 0.f, 0.f, 0.f);   // C % 4 == 0: the unshifted row ends with float4 last-1
            } else {
                o.raw[last] = __ldg(w + last);
            }
        } else if ((C & 1) == 0 && (reinterpret_cast<uintptr_t>(p.gr.logits_grad) & 7) == 0) {   // rows are 8-byte aligned
#pragma unroll
            for (int k = 0; k < CP2; ++k) {
                const float2 t = __ldg(reinterpret_cast<const float2 *>(row) + k);
                if (k & 1) { o.raw[k >> 1].z = t.x; o.raw[k >> 1].w = t.y; } else { o.raw[k >> 1].x = t.x; o.raw[k >> 1].y = t.y; }
            }
        } else {
#pragma unroll
            for (int k = 0; k < CP2; ++k) {
                const float a = __ldg(row + 2 * k);
                const float b = (2 * k + 1 < C) ? __ldg(row + 2 * k + 1) : 0.f;
                if (k & 1) { o.raw[k >> 1].z = a; o.raw[k >> 1].w = b; } else { o.raw[k >> 1].x = a; o.raw[k >> 1].y = b; }
            }
        }
        if (PROB) o.ax = __ldg(p.aux + n);
    }

    // contribution of one pair (the point lies inside the box)
    template <bool WIDE>
    __device__ __forceinline__ void consume(const PairData<C, PROB> &d) {
        if (!d.ok) return;
        const float dx = mu[0] - d.px, dy = mu[1] - d.py, dz = mu[2] - d.pz;
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
        if (!PROB) {
            // backward.cu:72-87 with t = sum_k sem_k * up_k
            float2 t = make_float2(0.f, 0.f);
            const float2 EE = make_float2(E, E);
#pragma unroll
            for (int k = 0; k < CP2; ++k) {
                const float2 u = d.template up<WIDE>(k);
                t = __ffma2_rn(sem[k], u, t);
                ss[k] = __ffma2_rn(u, EE, ss[k]);   // * opa at the end
            }
            const float et = E * (t.x + t.y);
            so += et;
            w = opa * et;
        } else {
            // localagg_prob/src/backward.cu:76-100 with the point-only terms pre-folded (prob_aux_kernel)
            const float Pt = norm * E;
            float pi = 0.f;
            if (d.ax.w > 0.f) {
                float2 u2 = make_float2(-d.ax.x, 0.f);
                const float sfac = Pt * opa * d.ax.w;
                const float2 ff = make_float2(sfac, sfac);
#pragma unroll
                for (int k = 0; k < CP2; ++k) {
                    const float2 uk = d.template up<WIDE>(k);
                    u2 = __ffma2_rn(uk, sem[k], u2);
                    ss[k] = __ffma2_rn(uk, ff, ss[k]);
                }
                const float u = u2.x + u2.y;
                pi = u * opa * d.ax.w;
                so = fmaf(u * Pt, d.ax.w, so);
            }
            const float eps = pi * norm + __fdividef(d.ax.y, 1.f - E + 1e-9f) + d.ax.z;
            sg = fmaf(pi * Pt, inv2det, sg);
            w = eps * E;
        }
        const float wx = w * dx, wy = w * dy, wz = w * dz;
        sd[0] += wx; sd[1] += wy; sd[2] += wz;
        sq[0] = fmaf(wx, dx, sq[0]);
        sq[1] = fmaf(wy, dy, sq[1]);
        sq[2] = fmaf(wz, dz, sq[2]);
        sq[3] = fmaf(wx, dy, sq[3]);
        sq[4] = fmaf(wy, dz, sq[4]);
        sq[5] = fmaf(wx, dz, sq[5]);
    }

    // The sums as a 32-vector in the order of the output lanes: [0,3) mean, 3 opacity, [4,10) cov,
    // [10,10+C) semantics, 10+C gamma (prob).  C <= 21.
    static constexpr int kVals = 10 + C + (PROB ? 1 : 0);
    static_assert(kVals <= 32, "one lane per reduced value");
    __device__ __forceinline__ void to_vector(float x[32]) const {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) x[a] = sd[a];
        x[3] = so;
#pragma unroll
        for (int a = 0; a < 6; ++a) x[4 + a] = sq[a];
#pragma unroll
        for (int k = 0; k < C; ++k) x[10 + k] = (k & 1) ? ss[k >> 1].y : ss[k >> 1].x;
        if (PROB) x[10 + C] = sg;
    }
};

// Sum of x[i] over the warp for all i at once: afterwards lane L holds the total of x[L] (in x[0]).
// 31 shuffles instead of 5 per value.
__device__ __forceinline__ float warp_transpose_reduce(float x[32], int lane) {
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
        const bool upper = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = upper ? x[i] : x[i + h];
            const float keep = upper ? x[i + h] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return x[0];
}

// Lane L holds the warp/CTA total of value L (see to_vector).  Applies the per-Gaussian linear maps and
// writes (or atomically adds) the gradients of Gaussian g.
template <int C, bool PROB>
__device__ __forceinline__ void finish(const BwdParams &p, int g, float v, int lane, bool atomic) {
    const float s0 = __shfl_sync(0xffffffffu, v, 0), s1 = __shfl_sync(0xffffffffu, v, 1), s2 = __shfl_sync(0xffffffffu, v, 2);
    const float sg = PROB ? __shfl_sync(0xffffffffu, v, 10 + C) : 0.f;
    float c6[6];
    load_cov6_in(p.d, p.in, g, c6);
    const float a = c6[0], b = c6[1], c = c6[2], d = c6[3], e = c6[4], f = c6[5];
    float out = 0.f;
    float *dst = nullptr;   // lane >= 10 + C: nothing to store
    if (lane < 3) {
        // -(A * sum w d): rows (a d f), (d b e), (f e c)
        const float r0 = lane == 0 ? a : (lane == 1 ? d : f);
        const float r1 = lane == 0 ? d : (lane == 1 ? b : e);
        const float r2 = lane == 0 ? f : (lane == 1 ? e : c);
        out = -(r0 * s0 + r1 * s1 + r2 * s2);
        dst = p.gr.means_grad + 3 * g + lane;
    } else if (lane == 3) {
        out = v;
        dst = p.gr.opacity_grad + g;
    } else if (lane < 10) {
        const int i = lane - 4;
        out = (i < 3) ? -0.5f * v : -v;
        if (PROB) {
            const float m[6] = {b * c - e * e, a * c - f * f, a * b - d * d,
                                2.f * (e * f - c * d), 2.f * (d * f - a * e), 2.f * (d * e - b * f)};
            float mi = m[0];
#pragma unroll
            for (int j = 1; j < 6; ++j) mi = (i == j) ? m[j] : mi;
            out = fmaf(sg, mi, out);
        }
        if (p.d.cov_stride == 9) {
            // gradient in the layout of the 3x3 input: the six gathered entries [0,4,8,1,5,2] receive it, the
            // lower triangle gets zero (what indexing autograd produces in the reference, __init__.py:143)
            const int flat = (i < 3) ? 4 * i : (i == 3 ? 1 : (i == 4 ? 5 : 2));
            dst = p.gr.cov_grad + 9 * static_cast<size_t>(g) + flat;
            if (!atomic && i < 3) p.gr.cov_grad[9 * static_cast<size_t>(g) + (i == 0 ? 3 : 5 + i)] = 0.f;
        } else {
            dst = p.gr.cov_grad + 6 * static_cast<size_t>(g) + i;
        }
    } else if (lane < 10 + C) {
        out = PROB ? v : p.in.opacities[g] * v;
        dst = p.gr.semantics_grad + static_cast<size_t>(g) * C + (lane - 10);
    }
    if (dst) {
        if (atomic) atomicAdd(dst, out); else *dst = out;
    }
}

// Walks the flat indices first, first+stride, ... < end of a box (z fastest) without a division per
// step: the stride is decomposed once into (sx, sy, sz) box steps and applied with two carries.
// Grid voxel indices fit 32 bits (H*W*D < 2^31 is checked at the C ABI).
struct BoxWalk {
    int nx, ny, nz;
    long long vol;
    int vbase;   // grid voxel index of the box corner
    int WD, D;
    int ix, iy, iz, sx, sy, sz;
    int left;    // steps this thread still has to take
    __device__ __forceinline__ void init(const int l[3], const int h[3], bool empty, int W, int D_) {
        nx = empty ? 0 : h[0] - l[0] + 1;
        ny = empty ? 0 : h[1] - l[1] + 1;
        nz = empty ? 0 : h[2] - l[2] + 1;
        vol = static_cast<long long>(nx) * ny * nz;
        D = D_; WD = W * D_;
        vbase = empty ? 0 : (l[0] * W + l[1]) * D_ + l[2];
    }
    __device__ __forceinline__ void start(long long first, long long end, int stride) {
        if (end > vol) end = vol;
        left = first < end ? static_cast<int>((end - first + stride - 1) / stride) : 0;
        if (left == 0) { ix = iy = iz = sx = sy = sz = 0; return; }
        iz = static_cast<int>(first % nz);
        const long long t = first / nz;
        iy = static_cast<int>(t % ny);
        ix = static_cast<int>(t / ny);
        sz = stride % nz;
        const int t2 = stride / nz;
        sy = t2 % ny;
        sx = t2 / ny;
    }
    __device__ __forceinline__ bool valid() const { return left > 0; }
    __device__ __forceinline__ int voxel() const { return vbase + ix * WD + iy * D + iz; }
    __device__ __forceinline__ void step() {
        --left;
        iz += sz;
        const int cz = iz >= nz;
        iz -= cz ? nz : 0;
        iy += sy + cz;
        const int cy = iy >= ny;
        iy -= cy ? ny : 0;
        ix += sx + cy;
    }
};

// The pair loop of one thread, software pipelined by one iteration (the loads of pair i+1 are in
// flight while pair i is evaluated).
template <int C, bool PROB, bool WIDE>
__device__ __forceinline__ void walk_pairs_t(const BwdParams &p, GaussAcc<C, PROB> &acc, BoxWalk &box, bool canon) {
    PairData<C, PROB> pa, pb;
    auto next = [&](PairData<C, PROB> &o) -> bool {   // false: the walk is over
        if (!box.valid()) { o.ok = false; return false; }
        const int v = box.voxel();
        acc.template fetch<WIDE>(p, canon ? v : __ldg(p.v2p + v), o);
        box.step();
        return true;
    };
    bool more = next(pa);
    while (more) {
        more = next(pb);
        acc.template consume<WIDE>(pa);
        if (!more) { acc.template consume<WIDE>(pb); break; }
        more = next(pa);
        acc.template consume<WIDE>(pb);
        if (!more) acc.template consume<WIDE>(pa);
    }
}

#ifndef GF_BWD_WIDE
#define GF_BWD_WIDE 1
#endif
template <int C, bool PROB>
__device__ __forceinline__ void walk_pairs(const BwdParams &p, GaussAcc<C, PROB> &acc, BoxWalk &box, bool canon) {
    // wide rows need an even class count and a 16-byte aligned gradient tensor (uniform for the launch)
    if (GF_BWD_WIDE && (C & 1) == 0 && (reinterpret_cast<uintptr_t>(p.gr.logits_grad) & 15) == 0)
        walk_pairs_t<C, PROB, true>(p, acc, box, canon);
    else
        walk_pairs_t<C, PROB, false>(p, acc, box, canon);
}

template <int C, bool PROB>
__global__ void __launch_bounds__(bwd_threads(PROB), bwd_ctas(PROB)) backward_small_kernel(const BwdParams pb) {
    constexpr int kBwdThreads = bwd_threads(PROB);
    const BwdParams p = sample_bwd(pb, blockIdx.y);
    const int lane = threadIdx.x & 31;
    // no early exit for the warps past G: they redo the last Gaussian and skip the stores, which keeps
    // every warp provably converged at the shuffles below
    const int g_raw = blockIdx.x * (kBwdThreads / 32) + (threadIdx.x >> 5);
    const bool live = g_raw < p.d.G;
    const int g = live ? g_raw : p.d.G - 1;
    GaussAcc<C, PROB> acc;
    acc.load(p, g);
    int lo[3], hi[3];
    uint32_t err = 0;
    const bool empty = gaussian_box(p.d, p.in, g, acc.mu, lo, hi, err) || !live;
    BoxWalk box;
    box.init(lo, hi, empty, p.d.W, p.d.D);
    const bool big = box.vol > kBigBox;
    pdl_launch_dependents();
    pdl_wait();   // everything above read the caller's inputs only; the map, the canonical flag and the queue follow
    const bool canon = *p.canon != 0;   // then voxel index == point index and the map need not be read
    if (!big) {
        box.start(lane, box.vol, 32);
        walk_pairs<C, PROB>(p, acc, box, canon);
    } else {
        if (lane == 0) {
            const unsigned long long nch = static_cast<unsigned long long>((box.vol + p.chunk - 1) / p.chunk);
            const unsigned long long old = atomicAdd(p.big_ctr, (1ull << 40) | nch);
            p.big[old >> 40] = make_uint2(static_cast<uint32_t>(g), static_cast<uint32_t>(old & ((1ull << 40) - 1)));
        }
    }
    // small boxes: final values; queued boxes: zeros (the big kernel accumulates atomically)
    float x[32];
    acc.to_vector(x);
    const float v = warp_transpose_reduce(x, lane);
    finish<C, PROB>(p, g, v, live ? lane : 32, false);
}

template <int C, bool PROB>
__global__ void __launch_bounds__(bwd_threads(PROB), bwd_ctas(PROB)) backward_big_kernel(const BwdParams pb) {
    constexpr int kBwdThreads = bwd_threads(PROB);
    const BwdParams p = sample_bwd(pb, blockIdx.y);
    pdl_wait();
    const unsigned long long ctr = *p.big_ctr;
    const int nbig = static_cast<int>(ctr >> 40);
    const long long total = static_cast<long long>(ctr & ((1ull << 40) - 1));
    if (nbig == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ float s_part[kBwdThreads / 32][32];
    const bool canon = *p.canon != 0;
    // my span of the chunk line, and the queue entry its first chunk belongs to (last entry with first <= lo)
    long long lo = total * blockIdx.x / gridDim.x;
    const long long hi = total * (blockIdx.x + 1) / gridDim.x;
    if (lo >= hi) return;
    int e = 0;
    for (int step = 1 << (31 - __clz(nbig)); step > 0; step >>= 1)
        if (e + step < nbig && static_cast<long long>(p.big[e + step].y) <= lo) e += step;
    for (; lo < hi; ++e) {
        const uint2 q = p.big[e];
        const int g = static_cast<int>(q.x);
        const long long next = (e + 1 < nbig) ? static_cast<long long>(p.big[e + 1].y) : total;
        const long long seg_end = next < hi ? next : hi;
        GaussAcc<C, PROB> acc;
        acc.load(p, g);
        int blo[3], bhi[3];
        uint32_t err = 0;
        const bool empty = gaussian_box(p.d, p.in, g, acc.mu, blo, bhi, err);
        BoxWalk box;
        box.init(blo, bhi, empty, p.d.W, p.d.D);
        const long long beg = (lo - q.y) * p.chunk;
        box.start(beg + threadIdx.x, (seg_end - q.y) * p.chunk, kBwdThreads);
        walk_pairs<C, PROB>(p, acc, box, canon);
        float x[32];
        acc.to_vector(x);
        const float v = warp_transpose_reduce(x, lane);
        s_part[warp][lane] = v;
        __syncthreads();
        if (warp == 0) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < kBwdThreads / 32; ++w) tot += s_part[w][lane];
            finish<C, PROB>(p, g, tot, lane, true);
        }
        __syncthreads();
        lo = seg_end;
    }
}

// Gradients of the fused pre-op (in.cov == NULL): the [G,6] gradient of Sigma^-1 = R^T diag(1/s^2) R mapped to the
// scales and the (un-normalised) quaternion.  With U the upper-triangular matrix of the six gradients (only the six
// gathered entries of the 3x3 carry gradient, __init__.py:143) and Gs = U + U^T:
//   dL/dD_k = 1/2 r_k^T Gs r_k (r_k = row k of R),  dL/ds_k = -2/s_k^3 dL/dD_k,  dL/dR_kl = D_k (Gs r_k)_l,
// then through the quaternion -> matrix map (model/utils/utils.py:20-66) and F.normalize.
__global__ void __launch_bounds__(128) srt_grad_kernel(const BwdParams pb) {
    pdl_wait();
    const BwdParams p = sample_bwd(pb, blockIdx.y);
    const int g = blockIdx.x * 128 + threadIdx.x;
    if (g >= p.d.G) return;
    const float *g6 = p.gr.cov_grad + 6 * static_cast<size_t>(g);
    const float s[3] = {p.in.scales[3 * g], p.in.scales[3 * g + 1], p.in.scales[3 * g + 2]};
    const float q[4] = {p.in.rotations[4 * g], p.in.rotations[4 * g + 1], p.in.rotations[4 * g + 2], p.in.rotations[4 * g + 3]};
    float R[3][3], inv_norm;
    quat_rotation(q, R, inv_norm);
    const float Gs[3][3] = {{2.f * g6[0], g6[3], g6[5]}, {g6[3], 2.f * g6[1], g6[4]}, {g6[5], g6[4], 2.f * g6[2]}};
    float M[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float Dk = __fdiv_rn(1.f, s[k] * s[k]);
        float t[3];
#pragma unroll
        for (int l = 0; l < 3; ++l) t[l] = Gs[l][0] * R[k][0] + Gs[l][1] * R[k][1] + Gs[l][2] * R[k][2];
        const float dD = 0.5f * (R[k][0] * t[0] + R[k][1] * t[1] + R[k][2] * t[2]);
        p.gr.scales_grad[3 * g + k] = -2.f * dD * __fdiv_rn(Dk, s[k]);
#pragma unroll
        for (int l = 0; l < 3; ++l) M[k][l] = Dk * t[l];
    }
    const float w = q[0] * inv_norm, x = q[1] * inv_norm, y = q[2] * inv_norm, z = q[3] * inv_norm;
    const float gw = 2.f * (w * (M[0][0] + M[1][1] + M[2][2]) - z * M[0][1] + y * M[0][2] + z * M[1][0] - x * M[1][2] - y * M[2][0] + x * M[2][1]);
    const float gx = 2.f * (x * (M[0][0] - M[1][1] - M[2][2]) + y * M[0][1] + z * M[0][2] + y * M[1][0] - w * M[1][2] + z * M[2][0] + w * M[2][1]);
    const float gy = 2.f * (y * (-M[0][0] + M[1][1] - M[2][2]) + x * M[0][1] + w * M[0][2] + x * M[1][0] + z * M[1][2] - w * M[2][0] + z * M[2][1]);
    const float gz = 2.f * (z * (-M[0][0] - M[1][1] + M[2][2]) - w * M[0][1] + x * M[0][2] + w * M[1][0] + y * M[1][2] + x * M[2][0] + y * M[2][1]);
    const float dot = w * gw + x * gx + y * gy + z * gz;
    p.gr.rotations_grad[4 * g] = inv_norm * (gw - w * dot);
    p.gr.rotations_grad[4 * g + 1] = inv_norm * (gx - x * dot);
    p.gr.rotations_grad[4 * g + 2] = inv_norm * (gy - y * dot);
    p.gr.rotations_grad[4 * g + 3] = inv_norm * (gz - z * dot);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct BwdWorkspace {
    int32_t *v2p, *canon;
    unsigned long long *big_ctr;
    uint2 *big;
    float4 *aux;
    float *cov6;       // [B,G,6] gradient of Sigma^-1 when the caller passed scales + rotations instead of cov
    size_t cv_block;   // bytes of one sample's (canon word, voxel -> point map) block
    size_t cv_bytes;   // all of them: one memset
    size_t bytes;
};

// Box voxels per chunk of the big-box kernel: 4 passes of a 256-thread CTA, grown when needed so that
// the chunk line of the worst case (every Gaussian covers the whole grid) stays below 2^31 chunks.
static int chunk_voxels(const gf_splat_desc &d) {
    const long long vox = static_cast<long long>(d.H) * d.W * d.D;
    long long c = 1024;
    while ((vox + c - 1) / c * static_cast<long long>(d.G > 0 ? d.G : 1) >= (1ll << 31)) c *= 2;
    return static_cast<int>(c);
}

static size_t align_up_b(size_t v, size_t a) { return (v + a - 1) / a * a; }

void plan_backward_workspace(const gf_splat_desc &d, void *base, BwdWorkspace *ws) {
    size_t off = 0;
    char *b = static_cast<char *>(base);
    const size_t B = static_cast<size_t>(batch_of(d));
    auto take = [&](size_t bytes) {
        char *p = b ? b + off : nullptr;
        off = align_up_b(off + bytes, 256);
        return p;
    };
    ws->big_ctr = reinterpret_cast<unsigned long long *>(take(B * 8));
    // per sample: the canon word (256 bytes) directly in front of its voxel -> point map; one memset covers all
    ws->cv_block = 256 + align_up_b(size_t(d.H) * d.W * d.D * 4, 256);
    ws->cv_bytes = B * ws->cv_block;
    ws->canon = reinterpret_cast<int32_t *>(take(ws->cv_bytes));
    ws->v2p = ws->canon ? ws->canon + 64 : nullptr;
    ws->big = reinterpret_cast<uint2 *>(take(B * size_t(d.G) * sizeof(uint2)));
    ws->aux = reinterpret_cast<float4 *>(take(d.variant == GF_SPLAT_PROB ? B * size_t(d.N) * 16 : 0));
    ws->cov6 = reinterpret_cast<float *>(take(B * size_t(d.G) * 6 * 4));   // only used with scales + rotations input
    ws->bytes = off;
}

size_t backward_workspace_bytes(const gf_splat_desc &d) {
    BwdWorkspace ws;
    plan_backward_workspace(d, nullptr, &ws);
    return ws.bytes;
}

template <int C, bool PROB>
static int launch_backward_t(const BwdParams &bp, const BwdWorkspace &ws, bool srt, int num_sms, cudaStream_t stream) {
    const gf_splat_desc &d = bp.d;
    const int B = batch_of(d);
    GF_REQUIRE(B <= 65535, GF_ERR_UNSUPPORTED, "splat backward: batch above 65535");
    // canon words (non-zero = "still canonical") and the voxel->point maps (-1 = empty) in one fill
    GF_CUDA_TRY(cudaMemsetAsync(bp.canon, 0xFF, ws.cv_bytes, stream));
    const long long want = (static_cast<long long>(d.N) + 255) / 256;
    const int grid0 = static_cast<int>(want < 16ll * num_sms ? (want > 0 ? want : 1) : 16ll * num_sms);
    voxel_map_kernel<<<dim3(grid0, B), 256, 0, stream>>>(bp);   // plain stream order: never overlaps the previous call
    GF_CUDA_TRY(cudaGetLastError());
    if (PROB) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid0, B); cfg.blockDim = dim3(256); cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
        GF_CUDA_TRY(cudaLaunchKernelEx(&cfg, prob_aux_kernel, bp, static_cast<int>(C)));
    }
    constexpr int kBwdThreads = bwd_threads(PROB);
    const int per_cta = kBwdThreads / 32;
    GF_CUDA_TRY(launch_chained(backward_small_kernel<C, PROB>, dim3((d.G + per_cta - 1) / per_cta, B), dim3(kBwdThreads), 0, stream, bp));
    int big_ctas = (num_sms * bwd_ctas(PROB) * 2 + B - 1) / B;   // per sample: together they fill the GPU twice over
    if (big_ctas < 32) big_ctas = 32;
    GF_CUDA_TRY(launch_chained(backward_big_kernel<C, PROB>, dim3(big_ctas, B), dim3(kBwdThreads), 0, stream, bp));
    if (srt) GF_CUDA_TRY(launch_chained(srt_grad_kernel, dim3((d.G + 127) / 128, B), dim3(128), 0, stream, bp));
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
    const bool srt = in.cov == nullptr;
    if (srt) {   // the pair kernels write the [G,6] gradient of Sigma^-1 into the workspace; srt_grad_kernel maps it on
        bp.d.cov_stride = 6;
        bp.gr.cov_grad = ws.cov6;
    }
    bp.v2p = ws.v2p;
    bp.big = ws.big;
    bp.big_ctr = ws.big_ctr;
    bp.chunk = chunk_voxels(d);
    bp.canon = ws.canon;
    bp.aux = ws.aux;
    bp.cv_block = ws.cv_block;
    const bool prob = d.variant == GF_SPLAT_PROB;
#define GF_CASE(CC)                                                                \
    case CC:                                                                       \
        return prob ? launch_backward_t<CC, true>(bp, ws, srt, num_sms, stream)    \
                    : launch_backward_t<CC, false>(bp, ws, srt, num_sms, stream);
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
