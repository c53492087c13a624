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
#include "splat_bwd_common.cuh"

namespace gf {

__global__ void __launch_bounds__(256) voxel_map_kernel(const BwdParams pb) {
    pdl_launch_dependents();   // the pair kernels may become resident; they wait before using the map
    const BwdParams p = sample_bwd(pb, blockIdx.y);
    if (p.bin_active) {   // launched behind the bin-centric kernel, which verified the point order: canonical -> done
        pdl_wait();
        if (*p.canon != 0) return;
    }
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

template <int C, bool PROB, bool STRIDED>
__global__ void __launch_bounds__(bwd_threads(PROB), bwd_ctas(PROB)) backward_small_kernel(const BwdParams pb) {
    constexpr int kBwdThreads = bwd_threads(PROB);
    const BwdParams p = sample_bwd(pb, blockIdx.y);
    const int lane = threadIdx.x & 31;
    if (p.bin_active) {   // canonical points: the bin-centric kernel has produced the gradients; nothing to do here
        pdl_wait();
        if (*p.canon != 0) return;
    }
    // One Gaussian per warp and pass; the grid normally covers all Gaussians in one pass.  Behind the bin-centric kernel
    // this kernel is only its (rare) fallback and is launched with a small grid that strides over the Gaussians
    // (STRIDED; a separate instantiation, so that the one-pass kernel keeps its register allocation).
    const int per_cta = kBwdThreads / 32;
    const int passes = STRIDED ? (p.d.G + gridDim.x * per_cta - 1) / (gridDim.x * per_cta) : 1;
    bool waited = false;
#pragma unroll 1
    for (int pass = 0; pass < passes; ++pass) {
        // no early exit for the warps past G: they redo the last Gaussian and skip the stores, which keeps
        // every warp provably converged at the shuffles below
        const int g_raw = (pass * gridDim.x + blockIdx.x) * per_cta + (threadIdx.x >> 5);
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
        if (!waited) {
            pdl_launch_dependents();
            pdl_wait();   // everything above read the caller's inputs only; the map, the canonical flag and the queue follow
            waited = true;
        }
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
}

template <int C, bool PROB>
__global__ void __launch_bounds__(bwd_threads(PROB), bwd_ctas(PROB)) backward_big_kernel(const BwdParams pb) {
    constexpr int kBwdThreads = bwd_threads(PROB);
    const BwdParams p = sample_bwd(pb, blockIdx.y);
    pdl_wait();
    if (p.bin_active && *p.canon != 0) return;   // (the queue is empty then anyway)
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
int plan_forward_workspace(const gf_splat_desc &d, void *base, SplatWorkspace *ws);   // splat_prep.cu
int launch_prep(const gf_splat_desc &d, const gf_splat_inputs &in, const SplatWorkspace &ws, uint32_t initial_flags,
                cudaStream_t stream, bool raw_records = false);
bool backward_bin_eligible(const gf_splat_desc &d, const gf_splat_inputs &in, const gf_splat_grads &gr);   // splat_backward_bin.cu
int launch_backward_bin(const BwdParams &bp, const SplatWorkspace &fws, float *sums, cudaStream_t stream);

struct BwdWorkspace {
    int32_t *v2p, *canon;
    unsigned long long *big_ctr;
    uint2 *big;
    float4 *aux;
    float *cov6;       // [B,G,6] gradient of Sigma^-1 when the caller passed scales + rotations instead of cov
    void *fwd;         // boxes + supertile lists for the bin-centric kernel (a forward-style workspace; 256-byte aligned)
    size_t fwd_bytes;
    float *sums;       // [B,G,32] raw sums of the bin-centric kernel
    size_t sums_bytes;
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
    SplatWorkspace fws;
    plan_forward_workspace(d, nullptr, &fws);
    ws->fwd_bytes = fws.bytes;
    ws->fwd = take(fws.bytes);
    ws->sums_bytes = B * size_t(d.G) * 32 * 4;
    ws->sums = reinterpret_cast<float *>(take(ws->sums_bytes));
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
    auto launch_aux = [&](bool chained) -> int {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid0, B); cfg.blockDim = dim3(256); cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = (chained && pdl_enabled()) ? 1 : 0;
        GF_CUDA_TRY(cudaLaunchKernelEx(&cfg, prob_aux_kernel, bp, static_cast<int>(C)));
        return GF_OK;
    };
    if (bp.bin_active) {
        // Canonical points (checked by the bin kernel itself): memset -> [prob_aux] -> pack -> list -> bin -> finish, and
        // the Gaussian-centric chain below finds the flag set and returns.  pack is launched in plain stream order (it
        // starts when everything above has completed); the kernels after it are chained to their predecessor.
        GF_CUDA_TRY(cudaMemsetAsync(ws.sums, 0, ws.sums_bytes, stream));   // accumulated with atomics
        if (PROB) { int rc = launch_aux(false); if (rc != GF_OK) return rc; }
        SplatWorkspace fws;
        plan_forward_workspace(d, ws.fwd, &fws);
        int rc = launch_prep(d, bp.in, fws, 0u, stream, true);
        if (rc != GF_OK) return rc;
        rc = launch_backward_bin(bp, fws, ws.sums, stream);
        if (rc != GF_OK) return rc;
        // the Gaussian-centric chain is only the fallback now: small grids (they return at once when the points are canonical)
        GF_CUDA_TRY(launch_chained(voxel_map_kernel, dim3(num_sms, B), dim3(256), 0, stream, bp));
    } else {
        voxel_map_kernel<<<dim3(grid0, B), 256, 0, stream>>>(bp);   // plain stream order: never overlaps the previous call
        GF_CUDA_TRY(cudaGetLastError());
        if (PROB) { int rc = launch_aux(true); if (rc != GF_OK) return rc; }
    }
    constexpr int kBwdThreads = bwd_threads(PROB);
    const int per_cta = kBwdThreads / 32;
    int small_ctas = (d.G + per_cta - 1) / per_cta;
    int big_ctas = (num_sms * bwd_ctas(PROB) * 2 + B - 1) / B;   // per sample: together they fill the GPU twice over
    if (big_ctas < 32) big_ctas = 32;
    if (bp.bin_active) {
        if (small_ctas > num_sms) small_ctas = num_sms;
        big_ctas = num_sms;
    }
    if (bp.bin_active) GF_CUDA_TRY(launch_chained(backward_small_kernel<C, PROB, true>, dim3(small_ctas, B), dim3(kBwdThreads), 0, stream, bp));
    else GF_CUDA_TRY(launch_chained(backward_small_kernel<C, PROB, false>, dim3(small_ctas, B), dim3(kBwdThreads), 0, stream, bp));
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
    bp.bin_active = backward_bin_eligible(bp.d, bp.in, bp.gr) ? 1 : 0;
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
