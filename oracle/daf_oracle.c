/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see splat_oracle.c header for the import rule).
 *
 * CPU restatement of the deformable multi-camera / multi-scale aggregation op:
 *
 *   forward   model/encoder/gaussian_encoder/ops/src/deformable_aggregation_cuda.cu:9-55,125-187
 *   backward  model/encoder/gaussian_encoder/ops/src/deformable_aggregation_cuda.cu:58-122,190-259
 *
 * Layouts (deformable_aggregation.cpp:41-71): feat[B,M,F,C] channels-last with the L pyramid
 * levels concatenated along F (level l starts at row start[l], is shape[l] = (h,w) row-major),
 * loc[B,P,M,2] = (x,y) normalised to (0,1), weights[B,P,M,L,Gr], out[B,P,C].
 * A camera contributes only if 0<x<1 and 0<y<1 (strict); sampling is bilinear with
 * pixel centres at (i+0.5)/size and zero padding (a corner counts iff it is inside the map).
 */
#include <math.h>
#include <stdint.h>

#ifndef REAL
#define REAL float
#define SUF _f32
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

typedef struct {
    int ok[4];        /* corner validity: (top,left) (top,right) (bottom,left) (bottom,right) */
    int64_t row[4];   /* row index inside the level */
    REAL wgt[4];      /* bilinear corner weights */
    REAL lh, lw, hh, hw;
} Corners;

static inline void corners_at(REAL y_im, REAL x_im, int h, int w, Corners *c) {
    const int y0 = (int)floor((double)y_im), x0 = (int)floor((double)x_im);
    const int y1 = y0 + 1, x1 = x0 + 1;
    c->lh = y_im - (REAL)y0;
    c->lw = x_im - (REAL)x0;
    c->hh = (REAL)1 - c->lh;
    c->hw = (REAL)1 - c->lw;
    c->ok[0] = (y0 >= 0 && x0 >= 0);
    c->ok[1] = (y0 >= 0 && x1 <= w - 1);
    c->ok[2] = (y1 <= h - 1 && x0 >= 0);
    c->ok[3] = (y1 <= h - 1 && x1 <= w - 1);
    c->row[0] = (int64_t)y0 * w + x0;
    c->row[1] = (int64_t)y0 * w + x1;
    c->row[2] = (int64_t)y1 * w + x0;
    c->row[3] = (int64_t)y1 * w + x1;
    c->wgt[0] = c->hh * c->hw;
    c->wgt[1] = c->hh * c->lw;
    c->wgt[2] = c->lh * c->hw;
    c->wgt[3] = c->lh * c->lw;
}

void FN(gfo_daf_forward)(int B, int M, int F, int C, int L, int P, int Gr,
                         const float *feat, const int *shape, const int *start,
                         const float *loc, const float *weights, REAL *out) {
    const int gdim = C / Gr;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < P; ++p) {
            REAL *o = out + ((int64_t)b * P + p) * C;
            for (int ch = 0; ch < C; ++ch) o[ch] = 0;
            for (int m = 0; m < M; ++m) {
                const float *lp = loc + (((int64_t)b * P + p) * M + m) * 2;
                const float lx = lp[0], ly = lp[1];
                if (!(lx > 0 && lx < 1 && ly > 0 && ly < 1)) continue;
                for (int l = 0; l < L; ++l) {
                    const int h = shape[2 * l], w = shape[2 * l + 1];
                    /* the reference forms loc*size in fp32 and subtracts 0.5 (cuda.cu:174-175) */
                    REAL y_im = (REAL)ly * (REAL)h - (REAL)0.5;
                    REAL x_im = (REAL)lx * (REAL)w - (REAL)0.5;
                    Corners cn;
                    corners_at(y_im, x_im, h, w, &cn);
                    const float *base = feat + (((int64_t)b * M + m) * F + start[l]) * C;
                    const float *wp = weights + ((((int64_t)b * P + p) * M + m) * L + l) * Gr;
                    for (int ch = 0; ch < C; ++ch) {
                        REAL v = 0;
                        for (int k = 0; k < 4; ++k)
                            if (cn.ok[k]) v += cn.wgt[k] * (REAL)base[cn.row[k] * C + ch];
                        o[ch] += v * (REAL)wp[ch / gdim];
                    }
                }
            }
        }
}

/* Serial over (b,p) because grad_feat rows collide; the oracle favours clarity. */
void FN(gfo_daf_backward)(int B, int M, int F, int C, int L, int P, int Gr,
                          const float *feat, const int *shape, const int *start,
                          const float *loc, const float *weights, const float *g_out,
                          REAL *g_feat, REAL *g_loc, REAL *g_weights) {
    const int gdim = C / Gr;
    for (int64_t i = 0; i < (int64_t)B * M * F * C; ++i) g_feat[i] = 0;
    for (int64_t i = 0; i < (int64_t)B * P * M * 2; ++i) g_loc[i] = 0;
    for (int64_t i = 0; i < (int64_t)B * P * M * L * Gr; ++i) g_weights[i] = 0;
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < P; ++p) {
            const float *go = g_out + ((int64_t)b * P + p) * C;
            for (int m = 0; m < M; ++m) {
                const int64_t li = (((int64_t)b * P + p) * M + m) * 2;
                const float lx = loc[li], ly = loc[li + 1];
                if (!(lx > 0 && lx < 1 && ly > 0 && ly < 1)) continue;
                for (int l = 0; l < L; ++l) {
                    const int h = shape[2 * l], w = shape[2 * l + 1];
                    REAL y_im = (REAL)ly * (REAL)h - (REAL)0.5;
                    REAL x_im = (REAL)lx * (REAL)w - (REAL)0.5;
                    Corners cn;
                    corners_at(y_im, x_im, h, w, &cn);
                    const int64_t fbase = (((int64_t)b * M + m) * F + start[l]) * C;
                    const int64_t wbase = ((((int64_t)b * P + p) * M + m) * L + l) * Gr;
                    for (int ch = 0; ch < C; ++ch) {
                        const REAL wt = weights[wbase + ch / gdim];
                        const REAL g = go[ch];
                        const REAL top = g * wt;
                        REAL v[4] = {0, 0, 0, 0};
                        for (int k = 0; k < 4; ++k)
                            if (cn.ok[k]) {
                                v[k] = feat[fbase + cn.row[k] * C + ch];
                                g_feat[fbase + cn.row[k] * C + ch] += cn.wgt[k] * top;
                            }
                        /* d(val)/d(y_im), d(val)/d(x_im): cuda.cu:85-121 */
                        REAL dy = -cn.hw * v[0] - cn.lw * v[1] + cn.hw * v[2] + cn.lw * v[3];
                        REAL dx = -cn.hh * v[0] + cn.hh * v[1] - cn.lh * v[2] + cn.lh * v[3];
                        REAL val = cn.wgt[0] * v[0] + cn.wgt[1] * v[1] + cn.wgt[2] * v[2] + cn.wgt[3] * v[3];
                        g_weights[wbase + ch / gdim] += g * val;
                        g_loc[li] += (REAL)w * dx * top;
                        g_loc[li + 1] += (REAL)h * dy * top;
                    }
                }
            }
        }
}
