/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path
 * (gaussianformer_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * CPU restatement of the GaussianFormer Gaussian->voxel splat ("local aggregate"),
 * written from the semantics of the reference, not from its code layout:
 *
 *   inclusion rule   model/head/localagg/src/auxiliary.h:8-20   (clipped integer AABB;
 *                    per-axis radii: model/head/localagg_prob_fast/src/auxiliary.h:8-20)
 *   per-voxel lists  model/head/localagg/src/aggregator_impl.cu:55-86,91-115,219-224
 *                    (pairs keyed by voxel, stable radix sort => ascending Gaussian index)
 *   base forward     model/head/localagg/src/forward.cu:46-82
 *   base backward    model/head/localagg/src/backward.cu:8-20,44-102
 *   prob forward     model/head/localagg_prob/src/forward.cu:56-101
 *   prob backward    model/head/localagg_prob/src/backward.cu:51-123
 *
 * Compiled twice (REAL=float: same precision as the reference; REAL=double: a
 * high-precision truth used for tolerances and finite differences).  Inputs are
 * always float32 / int32 arrays exactly as the reference's native boundary takes
 * them (model/head/localagg/src/aggregator.h:24-61).
 *
 * Parity pin: see oracle/README.md — pinned against outputs of the reference CUDA
 * op itself (oracle/_ref, built by oracle/build_ref.py, run on a B200) stored in
 * tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#define SUF _f32
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

#if defined(_OPENMP)
#include <omp.h>
#endif

static inline REAL r_exp(REAL x) { return (sizeof(REAL) == 4) ? (REAL)expf((float)x) : (REAL)exp((double)x); }
static inline REAL r_sqrt(REAL x) { return (sizeof(REAL) == 4) ? (REAL)sqrtf((float)x) : (REAL)sqrt((double)x); }

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Half-open clipped box of Gaussian g: [lo, hi) per axis (auxiliary.h:8-20). */
static inline void gaussian_box(const int *m, const int *rad, int axes, int g,
                                const int dims[3], int lo[3], int hi[3]) {
    for (int a = 0; a < 3; ++a) {
        int r = (axes == 3) ? rad[3 * g + a] : rad[g];
        lo[a] = clampi(m[3 * g + a] - r, 0, dims[a]);
        hi[a] = clampi(m[3 * g + a] + r + 1, 0, dims[a]);
    }
}

#ifdef ORACLE_SHARED_ONCE
/*
 * Per-voxel Gaussian lists in CSR form: start[v]..start[v+1] index into `list`,
 * Gaussians in ascending order (what the reference's stable sort produces).
 * Returns the number of (Gaussian, voxel) pairs; caller frees *start_out / *list_out.
 */
int64_t gfo_build_voxel_lists(int G, int H, int W, int D, const int *means_int,
                              const int *radii, int radii_axes,
                              int64_t **start_out, int32_t **list_out) {
    const int dims[3] = {H, W, D};
    const int64_t V = (int64_t)H * W * D;
    int64_t *start = (int64_t *)calloc((size_t)V + 1, sizeof(int64_t));
    for (int g = 0; g < G; ++g) {
        int lo[3], hi[3];
        gaussian_box(means_int, radii, radii_axes, g, dims, lo, hi);
        for (int x = lo[0]; x < hi[0]; ++x)
            for (int y = lo[1]; y < hi[1]; ++y)
                for (int z = lo[2]; z < hi[2]; ++z)
                    start[((int64_t)x * W + y) * D + z + 1]++;
    }
    for (int64_t v = 0; v < V; ++v) start[v + 1] += start[v];
    const int64_t R = start[V];
    int32_t *list = (int32_t *)malloc((size_t)(R > 0 ? R : 1) * sizeof(int32_t));
    int64_t *cursor = (int64_t *)malloc((size_t)V * sizeof(int64_t));
    memcpy(cursor, start, (size_t)V * sizeof(int64_t));
    for (int g = 0; g < G; ++g) { /* ascending g => each voxel's list is ascending */
        int lo[3], hi[3];
        gaussian_box(means_int, radii, radii_axes, g, dims, lo, hi);
        for (int x = lo[0]; x < hi[0]; ++x)
            for (int y = lo[1]; y < hi[1]; ++y)
                for (int z = lo[2]; z < hi[2]; ++z)
                    list[cursor[((int64_t)x * W + y) * D + z]++] = g;
    }
    free(cursor);
    *start_out = start;
    *list_out = list;
    return R;
}

void gfo_free(void *p) { free(p); }

/* OpenMP threads of the following calls made by this thread (bench.py: all host cores, whatever OMP_NUM_THREADS a
 * launcher such as torchrun exported) */
void gfo_set_num_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
}

int gfo_num_threads(void) {
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}
#else
int64_t gfo_build_voxel_lists(int, int, int, int, const int *, const int *, int, int64_t **, int32_t **);
#endif

/* exponent of the anisotropic Gaussian; cov = (a,b,c,d,e,f) = (A00,A11,A22,A01,A12,A02), delta = mu - x */
static inline REAL gauss_power(const float *cov, REAL dx, REAL dy, REAL dz) {
    REAL quad = (REAL)cov[0] * dx * dx + (REAL)cov[1] * dy * dy + (REAL)cov[2] * dz * dz;
    REAL cross = (REAL)cov[3] * dx * dy + (REAL)cov[4] * dy * dz + (REAL)cov[5] * dx * dz;
    return (REAL)-0.5 * quad - cross;
}

/* ---------------------------------------------------------------- base forward */
/* out[n,k] = sum_{g in list(voxel(n))} opa_g * sem_{g,k} * exp(power)   (forward.cu:61-81) */
int64_t FN(gfo_splat_forward)(int G, int N, int C, int H, int W, int D,
                              const float *pts, const int *points_int,
                              const float *means, const int *means_int,
                              const float *opa, const float *sem, const float *cov6,
                              const int *radii, int radii_axes, REAL *out) {
    int64_t *start;
    int32_t *list;
    int64_t R = gfo_build_voxel_lists(G, H, W, D, means_int, radii, radii_axes, &start, &list);
#pragma omp parallel for schedule(dynamic, 256)
    for (int n = 0; n < N; ++n) {
        const int64_t v = ((int64_t)points_int[3 * n] * W + points_int[3 * n + 1]) * D + points_int[3 * n + 2];
        REAL acc[64];
        for (int k = 0; k < C; ++k) acc[k] = 0;
        for (int64_t i = start[v]; i < start[v + 1]; ++i) {
            const int g = list[i];
            REAL dx = (REAL)means[3 * g] - (REAL)pts[3 * n];
            REAL dy = (REAL)means[3 * g + 1] - (REAL)pts[3 * n + 1];
            REAL dz = (REAL)means[3 * g + 2] - (REAL)pts[3 * n + 2];
            REAL w = (REAL)opa[g] * r_exp(gauss_power(cov6 + 6 * g, dx, dy, dz));
            for (int k = 0; k < C; ++k) acc[k] += (REAL)sem[(int64_t)C * g + k] * w;
        }
        for (int k = 0; k < C; ++k) out[(int64_t)n * C + k] = acc[k];
    }
    free(start);
    free(list);
    return R;
}

/* voxel -> point map of the backward pass (backward.cu:8-20).  The reference's write is
 * racy when several points share a voxel; the oracle keeps the largest point index. */
static int32_t *voxel_to_point(int N, int H, int W, int D, const int *points_int) {
    const int64_t V = (int64_t)H * W * D;
    int32_t *v2p = (int32_t *)malloc((size_t)V * sizeof(int32_t));
    for (int64_t v = 0; v < V; ++v) v2p[v] = -1;
    for (int n = 0; n < N; ++n) {
        int64_t v = ((int64_t)points_int[3 * n] * W + points_int[3 * n + 1]) * D + points_int[3 * n + 2];
        v2p[v] = n;
    }
    return v2p;
}

/* ---------------------------------------------------------------- base backward */
/* backward.cu:62-102: per Gaussian, walk its box; gradients for means, opacity, semantics, cov6 */
void FN(gfo_splat_backward)(int G, int N, int C, int H, int W, int D,
                            const float *pts, const int *points_int,
                            const float *means, const int *means_int,
                            const float *opa, const float *sem, const float *cov6,
                            const int *radii, int radii_axes, const float *out_grad,
                            REAL *g_means, REAL *g_opa, REAL *g_sem, REAL *g_cov6) {
    const int dims[3] = {H, W, D};
    int32_t *v2p = voxel_to_point(N, H, W, D, points_int);
#pragma omp parallel for schedule(dynamic, 16)
    for (int g = 0; g < G; ++g) {
        int lo[3], hi[3];
        gaussian_box(means_int, radii, radii_axes, g, dims, lo, hi);
        const float *cv = cov6 + 6 * g;
        const REAL a = cv[0], b = cv[1], c = cv[2], d = cv[3], e = cv[4], f = cv[5];
        const REAL o = opa[g];
        REAL gm[3] = {0, 0, 0}, go = 0, gc[6] = {0, 0, 0, 0, 0, 0}, gs[64];
        for (int k = 0; k < C; ++k) gs[k] = 0;
        for (int x = lo[0]; x < hi[0]; ++x)
            for (int y = lo[1]; y < hi[1]; ++y)
                for (int z = lo[2]; z < hi[2]; ++z) {
                    const int n = v2p[((int64_t)x * W + y) * D + z];
                    if (n < 0) continue;
                    REAL dx = (REAL)means[3 * g] - (REAL)pts[3 * n];
                    REAL dy = (REAL)means[3 * g + 1] - (REAL)pts[3 * n + 1];
                    REAL dz = (REAL)means[3 * g + 2] - (REAL)pts[3 * n + 2];
                    REAL E = r_exp(gauss_power(cv, dx, dy, dz));
                    REAL t = 0; /* sum_k sem_k * dL/dout[n,k] */
                    for (int k = 0; k < C; ++k) {
                        REAL up = E * (REAL)out_grad[(int64_t)n * C + k];
                        gs[k] += o * up;
                        t += (REAL)sem[(int64_t)C * g + k] * up;
                    }
                    go += t;
                    REAL coef = o * t;
                    gc[0] += (REAL)-0.5 * coef * dx * dx;
                    gc[1] += (REAL)-0.5 * coef * dy * dy;
                    gc[2] += (REAL)-0.5 * coef * dz * dz;
                    gc[3] -= coef * dx * dy;
                    gc[4] -= coef * dy * dz;
                    gc[5] -= coef * dx * dz;
                    gm[0] -= coef * (a * dx + d * dy + f * dz);
                    gm[1] -= coef * (d * dx + b * dy + e * dz);
                    gm[2] -= coef * (f * dx + e * dy + c * dz);
                }
        for (int i = 0; i < 3; ++i) g_means[3 * g + i] = gm[i];
        g_opa[g] = go;
        for (int k = 0; k < C; ++k) g_sem[(int64_t)C * g + k] = gs[k];
        for (int i = 0; i < 6; ++i) g_cov6[6 * g + i] = gc[i];
    }
    free(v2p);
}

/* ---------------------------------------------------------------- prob forward */
/* localagg_prob/src/forward.cu:63-101.  kappa = powf(2*3.1415926535, -1.5). */
static inline REAL kappa_const(void) {
    return (sizeof(REAL) == 4) ? (REAL)powf((float)(2 * 3.1415926535), -1.5f) : (REAL)pow(2 * 3.1415926535, -1.5);
}
static inline REAL cov_det(const float *cv) {
    const REAL a = cv[0], b = cv[1], c = cv[2], d = cv[3], e = cv[4], f = cv[5];
    return a * b * c + 2 * d * e * f - a * e * e - b * f * f - c * d * d;
}

int64_t FN(gfo_splat_prob_forward)(int G, int N, int C, int H, int W, int D,
                                   const float *pts, const int *points_int,
                                   const float *means, const int *means_int,
                                   const float *opa, const float *sem, const float *cov6,
                                   const int *radii, int radii_axes,
                                   REAL *logits, REAL *bin_logits, REAL *density, REAL *probability) {
    int64_t *start;
    int32_t *list;
    int64_t R = gfo_build_voxel_lists(G, H, W, D, means_int, radii, radii_axes, &start, &list);
    const REAL kappa = kappa_const();
#pragma omp parallel for schedule(dynamic, 256)
    for (int n = 0; n < N; ++n) {
        const int64_t v = ((int64_t)points_int[3 * n] * W + points_int[3 * n + 1]) * D + points_int[3 * n + 2];
        REAL acc[64];
        for (int k = 0; k < C; ++k) acc[k] = 0;
        REAL keep = 1, dens = 0, Z = 0;
        for (int64_t i = start[v]; i < start[v + 1]; ++i) {
            const int g = list[i];
            REAL dx = (REAL)means[3 * g] - (REAL)pts[3 * n];
            REAL dy = (REAL)means[3 * g + 1] - (REAL)pts[3 * n + 1];
            REAL dz = (REAL)means[3 * g + 2] - (REAL)pts[3 * n + 2];
            REAL E = r_exp(gauss_power(cov6 + 6 * g, dx, dy, dz));
            REAL P = kappa * r_sqrt(cov_det(cov6 + 6 * g)) * E * (REAL)opa[g];
            for (int k = 0; k < C; ++k) acc[k] += (REAL)sem[(int64_t)C * g + k] * P;
            keep = ((REAL)1 - E) * keep;
            dens = E + dens;
            Z = P + Z;
        }
        if (Z > (REAL)1e-9) {
            for (int k = 0; k < C; ++k) logits[(int64_t)n * C + k] = acc[k] / Z;
        } else { /* uniform over the first C-1 classes; the last channel keeps its zero fill */
            for (int k = 0; k < C - 1; ++k) logits[(int64_t)n * C + k] = (REAL)(1.0 / (C - 1));
            logits[(int64_t)n * C + C - 1] = 0;
        }
        bin_logits[n] = (REAL)1 - keep;
        density[n] = dens;
        probability[n] = Z;
    }
    free(start);
    free(list);
    return R;
}

/* ---------------------------------------------------------------- prob backward */
/* localagg_prob/src/backward.cu:62-123 (not autograd-exact: +1e-9 in the bin term,
 * logits branch skipped when Z <= 1e-9).  Saved forward outputs are inputs, as in the reference. */
void FN(gfo_splat_prob_backward)(int G, int N, int C, int H, int W, int D,
                                 const float *pts, const int *points_int,
                                 const float *means, const int *means_int,
                                 const float *opa, const float *sem, const float *cov6,
                                 const int *radii, int radii_axes,
                                 const float *logits, const float *bin_logits, const float *probability,
                                 const float *g_logits, const float *g_bin, const float *g_density,
                                 REAL *g_means, REAL *g_opa, REAL *g_sem, REAL *g_cov6) {
    const int dims[3] = {H, W, D};
    int32_t *v2p = voxel_to_point(N, H, W, D, points_int);
    const REAL kappa = kappa_const();
#pragma omp parallel for schedule(dynamic, 16)
    for (int g = 0; g < G; ++g) {
        int lo[3], hi[3];
        gaussian_box(means_int, radii, radii_axes, g, dims, lo, hi);
        const float *cv = cov6 + 6 * g;
        const REAL a = cv[0], b = cv[1], c = cv[2], d = cv[3], e = cv[4], f = cv[5];
        const REAL o = opa[g];
        const REAL det = cov_det(cv);
        const REAL norm = kappa * r_sqrt(det);
        REAL gm[3] = {0, 0, 0}, go = 0, gc[6] = {0, 0, 0, 0, 0, 0}, gs[64];
        for (int k = 0; k < C; ++k) gs[k] = 0;
        for (int x = lo[0]; x < hi[0]; ++x)
            for (int y = lo[1]; y < hi[1]; ++y)
                for (int z = lo[2]; z < hi[2]; ++z) {
                    const int n = v2p[((int64_t)x * W + y) * D + z];
                    if (n < 0) continue;
                    REAL dx = (REAL)means[3 * g] - (REAL)pts[3 * n];
                    REAL dy = (REAL)means[3 * g + 1] - (REAL)pts[3 * n + 1];
                    REAL dz = (REAL)means[3 * g + 2] - (REAL)pts[3 * n + 2];
                    REAL E = r_exp(gauss_power(cv, dx, dy, dz));
                    REAL Pt = norm * E; /* probability density without opacity */
                    REAL Z = probability[n];
                    REAL pi = 0;
                    if (Z > (REAL)1e-9) {
                        for (int k = 0; k < C; ++k) {
                            REAL up = g_logits[(int64_t)n * C + k];
                            REAL dev = (REAL)sem[(int64_t)C * g + k] - (REAL)logits[(int64_t)n * C + k];
                            gs[k] += up * Pt * o / Z;
                            pi += up * dev * o / Z;
                            go += up * dev * Pt / Z;
                        }
                    }
                    REAL eps = pi * norm
                             + ((REAL)1 - (REAL)bin_logits[n]) / ((REAL)1 - E + (REAL)1e-9) * (REAL)g_bin[n]
                             + (REAL)g_density[n];
                    REAL gam = pi * Pt / 2 / det;
                    REAL eE = eps * E;
                    gm[0] -= eE * (a * dx + d * dy + f * dz);
                    gm[1] -= eE * (d * dx + b * dy + e * dz);
                    gm[2] -= eE * (f * dx + e * dy + c * dz);
                    gc[0] += eE * ((REAL)-0.5 * dx * dx) + gam * (b * c - e * e);
                    gc[1] += eE * ((REAL)-0.5 * dy * dy) + gam * (a * c - f * f);
                    gc[2] += eE * ((REAL)-0.5 * dz * dz) + gam * (a * b - d * d);
                    gc[3] += eE * (-dx * dy) + 2 * gam * (e * f - c * d);
                    gc[4] += eE * (-dy * dz) + 2 * gam * (d * f - a * e);
                    gc[5] += eE * (-dx * dz) + 2 * gam * (d * e - b * f);
                }
        for (int i = 0; i < 3; ++i) g_means[3 * g + i] = gm[i];
        g_opa[g] = go;
        for (int k = 0; k < C; ++k) g_sem[(int64_t)C * g + k] = gs[k];
        for (int i = 0; i < 6; ++i) g_cov6[6 * g + i] = gc[i];
    }
    free(v2p);
}
