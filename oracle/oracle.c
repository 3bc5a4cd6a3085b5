/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * CPU restatement (plain C, per-voxel loops) of the hot path of
 * balbasty/torch-interpol: grid_pull / grid_push / grid_count / grid_grad /
 * grid_pushgrad / grid_hess and the spline prefilter.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may build, load or call this file.  The product (torch-interpol_b200/) never
 * links or imports it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * here against fixtures in tests/golden/ that were produced by importing the
 * unmodified reference (tests/golden/make_golden.py, run in the build
 * container where /root/reference is mounted).
 *
 * The file is compiled twice, with -DREAL=double -DSFX=_f64 and with
 * -DREAL=float -DSFX=_f32 (see oracle/Makefile).  All arithmetic is carried
 * out in REAL, like the reference which computes in the tensors' dtype
 * (interpol/nd.py:45-46).
 *
 * Citations are to files under the reference tree (interpol/...).
 *
 * Deliberate deviations from the reference (documented in DESIGN.md):
 *   - iso0.pull2d returns mask*mask for extrapolate in {0,2} (iso0.py:155);
 *     here the 2-D nearest pull multiplies data by the mask like 1-D/3-D.
 *   - the derivative of the order-1 spline on an axis of a mixed-order call
 *     has the wrong sign in splines.py:96-97; `quirk_linear_grad` != 0
 *     reproduces the reference, 0 gives the true derivative (-sign(x)).
 *   - nd.hess broadcasts its mask to the wrong axes when C > 1
 *     (nd.py:455-456); here the mask multiplies every channel.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL double
#define SFX _f64
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

typedef long long i64;

/* ------------------------------------------------------------------ */
/* Boundary index / sign maps -- interpol/bounds.py:30-89              */
/* ------------------------------------------------------------------ */

/* python-style remainder (non-negative for positive divisor) */
static i64 pymod(i64 a, i64 n) {
    i64 r = a % n;
    return r < 0 ? r + n : r;
}

/* Bound.index, bounds.py:30-60 */
i64 FN(orc_bound_index)(int bound, i64 i, i64 n) {
    switch (bound) {
    case 0: /* zero */
    case 1: /* replicate -- bounds.py:31-32 */
        return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
    case 3: /* dct2 */
    case 5: { /* dst2 -- bounds.py:33-38 */
        i64 n2 = n * 2;
        i = i < 0 ? (n2 - 1) - pymod(-i - 1, n2) : pymod(i, n2);
        if (i >= n) i = n2 - 1 - i;
        return i;
    }
    case 2: { /* dct1 -- bounds.py:39-46 */
        if (n == 1) return 0;
        i64 n2 = (n - 1) * 2;
        i = pymod(i < 0 ? -i : i, n2);
        if (i >= n) i = n2 - i;
        return i;
    }
    case 4: { /* dst1 -- bounds.py:47-56 */
        i64 n2 = 2 * (n + 1);
        if (i < 0) i = -i - 2;
        i = pymod(i, n2);
        if (i > n) i = n2 - 2 - i;
        if (i == -1) i = 0;
        if (i == n) i = n - 1;
        return i;
    }
    case 6: /* dft -- bounds.py:57-58 */
        return pymod(i, n);
    default:
        return i;
    }
}

/* Bound.transform, bounds.py:62-89.  Returns -1, 0 or +1 (None -> +1). */
int FN(orc_bound_sign)(int bound, i64 i, i64 n) {
    switch (bound) {
    case 4: { /* dst1 -- bounds.py:63-75 (zeroes i == 0 mod 2(n+1): quirk Q1) */
        if (n == 1) return 1;
        i64 n2 = 2 * (n + 1);
        if (i < 0) i = -i + (n - 1);
        i = pymod(i, n2);
        int x = (i == 0) ? 0 : 1;
        if (pymod(i, n + 1) == n) x = 0;
        i = i / (n + 1);
        return (i % 2 > 0) ? -x : x;
    }
    case 5: { /* dst2 -- bounds.py:76-81 */
        if (i < 0) i = n - 1 - i;
        i = i / n;
        return (i % 2 > 0) ? -1 : 1;
    }
    case 0: /* zero -- bounds.py:82-87 */
        return (i < 0 || i >= n) ? 0 : 1;
    default:
        return 1;
    }
}

/* ------------------------------------------------------------------ */
/* B-spline weight / first / second derivative -- interpol/splines.py */
/* ------------------------------------------------------------------ */

static REAL sq(REAL x) { return x * x; }
static REAL cube(REAL x) { return x * x * x; }
static REAL p4(REAL x) { x = x * x; return x * x; }
static REAL p5(REAL x) { return p4(x) * x; }
static REAL p6(REAL x) { return sq(cube(x)); }
static REAL p7(REAL x) { return p6(x) * x; }

/* Spline.fastweight, splines.py:30-80 */
REAL FN(orc_weight)(int order, REAL x) {
    const REAL one = 1;
    if (order == 0) return one;
    x = (REAL)fabs((double)x);
    switch (order) {
    case 1:
        return 1 - x;
    case 2:
        return x < (REAL)0.5 ? (REAL)0.75 - sq(x) : (REAL)0.5 * sq((REAL)1.5 - x);
    case 3:
        return x < 1 ? (x * x * (x - 2) * 3 + 4) / 6 : cube(2 - x) / 6;
    case 4: {
        if (x < (REAL)0.5) {
            REAL y = sq(x);
            return y * (y * (REAL)0.25 - (REAL)0.625) + (REAL)(115. / 192.);
        } else if (x < (REAL)1.5) {
            return x * (x * (x * (5 - x) / 6 - (REAL)1.25) + (REAL)(5. / 24.)) + (REAL)(55. / 96.);
        }
        return p4(x - (REAL)2.5) / 24;
    }
    case 5: {
        if (x < 1) {
            REAL y = sq(x);
            return y * (y * ((REAL)0.25 - x / 12) - (REAL)0.5) + (REAL)0.55;
        } else if (x < 2) {
            return x * (x * (x * (x * (x / 24 - (REAL)0.375) + (REAL)1.25) - (REAL)1.75) + (REAL)0.625) + (REAL)0.425;
        }
        return p5(3 - x) / 120;
    }
    case 6: {
        if (x < (REAL)0.5) {
            REAL y = sq(x);
            return y * (y * ((REAL)(7. / 48.) - y / 36) - (REAL)(77. / 192.)) + (REAL)(5887. / 11520.);
        } else if (x < (REAL)1.5) {
            return x * (x * (x * (x * (x * (x / 48 - (REAL)(7. / 48.)) + (REAL)0.328125) - (REAL)(35. / 288.)) - (REAL)(91. / 256.)) - (REAL)(7. / 768.)) + (REAL)(7861. / 15360.);
        } else if (x < (REAL)2.5) {
            return x * (x * (x * (x * (x * ((REAL)(7. / 60.) - x / 120) - (REAL)0.65625) + (REAL)(133. / 72.)) - (REAL)2.5703125) + (REAL)(1267. / 960.)) + (REAL)(1379. / 7680.);
        }
        return p6(x - (REAL)3.5) / 720;
    }
    case 7: {
        if (x < 1) {
            REAL y = sq(x);
            return y * (y * (y * (x / 144 - (REAL)(1. / 36.)) + (REAL)(1. / 9.)) - (REAL)(1. / 3.)) + (REAL)(151. / 315.);
        } else if (x < 2) {
            return x * (x * (x * (x * (x * (x * ((REAL)0.05 - x / 240) - (REAL)(7. / 30.)) + (REAL)0.5) - (REAL)(7. / 18.)) - (REAL)0.1) - (REAL)(7. / 90.)) + (REAL)(103. / 210.);
        } else if (x < 3) {
            return x * (x * (x * (x * (x * (x * (x / 720 - (REAL)(1. / 36.)) + (REAL)(7. / 30.)) - (REAL)(19. / 18.)) + (REAL)(49. / 18.)) - (REAL)(23. / 6.)) + (REAL)(217. / 90.)) - (REAL)(139. / 630.);
        }
        return p7(4 - x) / 5040;
    }
    }
    return 0;
}

/* Spline._fastgrad on |x|, splines.py:95-139 */
static REAL fastgrad_abs(int order, REAL x, int quirk_linear_grad) {
    switch (order) {
    case 1:
        /* splines.py:96-97 returns +1; the true derivative of 1-|x| is -1 */
        return quirk_linear_grad ? (REAL)1 : (REAL)-1;
    case 2:
        return x < (REAL)0.5 ? -2 * x : x - (REAL)1.5;
    case 3:
        return x < 1 ? x * (x * (REAL)1.5 - 2) : (REAL)-0.5 * sq(2 - x);
    case 4:
        if (x < (REAL)0.5) return x * (sq(x) - (REAL)1.25);
        if (x < (REAL)1.5) return x * (x * (x * (REAL)(-2. / 3.) + (REAL)2.5) - (REAL)2.5) + (REAL)(5. / 24.);
        return cube(2 * x - 5) / 48;
    case 5:
        if (x < 1) return x * (x * (x * (x * (REAL)(-5. / 12.) + 1)) - 1);
        if (x < 2) return x * (x * (x * (x * (REAL)(5. / 24.) - (REAL)1.5) + (REAL)3.75) - (REAL)3.5) + (REAL)0.625;
        return p4(x - 3) / (-24);
    case 6:
        if (x < (REAL)0.5) {
            REAL y = sq(x);
            return x * (y * (REAL)(7. / 12.) - sq(y) / 6 - (REAL)(77. / 96.));
        }
        if (x < (REAL)1.5)
            return x * (x * (x * (x * (x * (REAL)0.125 - (REAL)(35. / 48.)) + (REAL)1.3125) - (REAL)(35. / 96.)) - (REAL)0.7109375) - (REAL)(7. / 768.);
        if (x < (REAL)2.5)
            return x * (x * (x * (x * (x / (-20) + (REAL)(7. / 12.)) - (REAL)2.625) + (REAL)(133. / 24.)) - (REAL)5.140625) + (REAL)(1267. / 960.);
        return p5(2 * x - 7) / 3840;
    case 7:
        if (x < 1) {
            REAL y = sq(x);
            return x * (y * (y * (x * (REAL)(7. / 144.) - (REAL)(1. / 6.)) + (REAL)(4. / 9.)) - (REAL)(2. / 3.));
        }
        if (x < 2)
            return x * (x * (x * (x * (x * (x * (REAL)(-7. / 240.) + (REAL)(3. / 10.)) - (REAL)(7. / 6.)) + 2) - (REAL)(7. / 6.)) - (REAL)(1. / 5.)) - (REAL)(7. / 90.);
        if (x < 3)
            return x * (x * (x * (x * (x * (x * (REAL)(7. / 720.) - (REAL)(1. / 6.)) + (REAL)(7. / 6.)) - (REAL)(38. / 9.)) + (REAL)(49. / 6.)) - (REAL)(23. / 3.)) + (REAL)(217. / 90.);
        return p6(x - 4) / (-720);
    }
    return 0;
}

/* Spline.fastgrad, splines.py:90-93: _fastgrad(|x|) * sign(x) */
REAL FN(orc_grad)(int order, REAL x, int quirk_linear_grad) {
    if (order == 0) return 0;
    REAL s = (x > 0) ? (REAL)1 : ((x < 0) ? (REAL)-1 : (REAL)0);
    return fastgrad_abs(order, (REAL)fabs((double)x), quirk_linear_grad) * s;
}

/* Spline.fasthess, splines.py:149-195 */
REAL FN(orc_hess)(int order, REAL x) {
    if (order == 0 || order == 1) return 0;
    x = (REAL)fabs((double)x);
    switch (order) {
    case 2:
        return x < (REAL)0.5 ? (REAL)-2 : (REAL)1;
    case 3:
        return x < 1 ? 3 * x - 2 : 2 - x;
    case 4:
        if (x < (REAL)0.5) return 3 * sq(x) - (REAL)1.25;
        if (x < (REAL)1.5) return x * (-2 * x + 5) - (REAL)2.5;
        return sq(2 * x - 5) / 8;
    case 5:
        if (x < 1) {
            REAL y = sq(x);
            return -y * (x * (REAL)(5. / 3.) - 3) - 1;
        }
        if (x < 2) return x * (x * (x * (REAL)(5. / 6.) - (REAL)(9. / 2.)) + (REAL)(15. / 2.)) - (REAL)(7. / 2.);
        return (REAL)(9. / 2.) - x * (x * (x / 6 - (REAL)(3. / 2.)) + (REAL)(9. / 2.));
    case 6:
        if (x < (REAL)0.5) {
            REAL y = sq(x);
            return -y * (y * (REAL)(5. / 6) - (REAL)(7. / 4.)) - (REAL)(77. / 96.);
        }
        if (x < (REAL)1.5)
            return x * (x * (x * (x * (REAL)(5. / 8.) - (REAL)(35. / 12.)) + (REAL)(63. / 16.)) - (REAL)(35. / 48.)) - (REAL)(91. / 128.);
        if (x < (REAL)2.5)
            return -(x * (x * (x * (x / 4 - (REAL)(7. / 3.)) + (REAL)(63. / 8.)) - (REAL)(133. / 12.)) + (REAL)(329. / 64.));
        return x * (x * (x * (x / 24 - (REAL)(7. / 12.)) + (REAL)(49. / 16.)) - (REAL)(343. / 48.)) + (REAL)(2401. / 384.);
    case 7:
        if (x < 1) {
            REAL y = sq(x);
            return y * (y * (x * (REAL)(7. / 24.) - (REAL)(5. / 6.)) + (REAL)(4. / 3.)) - (REAL)(2. / 3.);
        }
        if (x < 2)
            return -(x * (x * (x * (x * (x * (REAL)(7. / 40.) - (REAL)(3. / 2.)) + (REAL)(14. / 3.)) - 6) + (REAL)(7. / 3.)) + (REAL)(1. / 5.));
        if (x < 3)
            return x * (x * (x * (x * (x * (REAL)(7. / 120.) - (REAL)(5. / 6.)) + (REAL)(14. / 3.)) - (REAL)(38. / 3.)) + (REAL)(49. / 3.)) - (REAL)(23. / 3.);
        return -(x * (x * (x * (x * (x / 120 - (REAL)(1. / 6.)) + (REAL)(4. / 3.)) - (REAL)(16. / 3.)) + (REAL)(32. / 3.)) - (REAL)(128. / 15.));
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* Per-voxel support: nd.get_weights (nd.py:31-77), iso0.get_indices  */
/* (iso0.py:11-15), nd.inbounds_mask (nd.py:11-27)                    */
/* ------------------------------------------------------------------ */

typedef struct {
    int nnodes[3];
    i64 idx[3][8];
    REAL w[3][8];  /* weight * sign */
    REAL g[3][8];  /* first derivative * sign */
    REAL h[3][8];  /* second derivative * sign */
    int inb;       /* inbounds mask (1 when extrapolate == 1) */
} support_t;

static void make_support(support_t *s, const REAL *coord, int dim,
                         const i64 *shape, const int *bound, const int *order,
                         int extrapolate, int quirk_linear_grad) {
    int all0 = 1, all1 = 1;
    for (int d = 0; d < dim; ++d) {
        if (order[d] != 0) all0 = 0;
        if (order[d] != 1) all1 = 0;
    }
    s->inb = 1;
    if (extrapolate == 0 || extrapolate == 2) { /* nd.py:15-26 */
        /* thresholds are python doubles, cast to the grid dtype by the compare */
        double thr = 5e-2;
        if (extrapolate == 2) thr = 0.5 + 5e-2;
        for (int d = 0; d < dim; ++d) {
            if (!(coord[d] > (REAL)(-thr))) s->inb = 0;
            if (!(coord[d] < (REAL)((double)(shape[d] - 1) + thr))) s->inb = 0;
        }
    }
    for (int d = 0; d < 3; ++d) {
        s->nnodes[d] = 1;
        s->idx[d][0] = 0;
        s->w[d][0] = 1; s->g[d][0] = 0; s->h[d][0] = 0;
    }
    for (int d = 0; d < dim; ++d) {
        const int o = order[d];
        const i64 n = shape[d];
        REAL g = coord[d];
        REAL g0;
        if (all0) {
            /* iso0.py:12: round() is round-half-to-even */
            g0 = (REAL)nearbyint((double)g);
        } else {
            /* nd.py:45 (order 1 == iso1.py:13): floor(g - (order-1)/2) */
            g0 = (REAL)floor((double)(g - (REAL)(o - 1) / 2));
        }
        REAL dist0 = g - g0; /* nd.py:46 */
        if (!isfinite((double)g0)) { g0 = 0; dist0 = 0; s->inb = 0; }
        i64 i0 = (i64)g0;
        s->nnodes[d] = o + 1;
        for (int k = 0; k <= o; ++k) {
            i64 i1 = i0 + k;
            int sign = FN(orc_bound_sign)(bound[d], i1, n);
            s->idx[d][k] = FN(orc_bound_index)(bound[d], i1, n);
            REAL x = dist0 - (REAL)k; /* nd.py:60 */
            REAL w, gr, he;
            if (all0) {
                w = 1; gr = 0; he = 0;
            } else if (all1) {
                /* iso1.py:19 and the closed forms used by iso1 grad/hess */
                w = (k == 0) ? 1 - dist0 : dist0;
                gr = (k == 0) ? (REAL)-1 : (REAL)1;
                he = 0;
            } else {
                w = FN(orc_weight)(o, x);
                gr = FN(orc_grad)(o, x, quirk_linear_grad);
                he = FN(orc_hess)(o, x);
            }
            s->w[d][k] = w * (REAL)sign;
            s->g[d][k] = gr * (REAL)sign;
            s->h[d][k] = he * (REAL)sign;
        }
    }
}

static i64 prod(const i64 *s, int dim) {
    i64 p = 1;
    for (int d = 0; d < dim; ++d) p *= s[d];
    return p;
}

static void strides3(const i64 *shape, int dim, i64 *st) {
    /* jit_utils.py:164-190: last axis fastest */
    st[0] = st[1] = st[2] = 0;
    i64 acc = 1;
    for (int d = dim - 1; d >= 0; --d) { st[d] = acc; acc *= shape[d]; }
}

/* ------------------------------------------------------------------ */
/* pull -- nd.py:81-143 (iso1.py:29-140, iso0.py:24-62)               */
/* inp (B,C,*ishape), grid (B,*oshape,dim) -> out (B,C,*oshape)       */
/* ------------------------------------------------------------------ */
void FN(orc_pull)(const REAL *inp, const REAL *grid, REAL *out, i64 B, i64 C,
                  int dim, const i64 *ishape, const i64 *oshape,
                  const int *bound, const int *order, int extrapolate) {
    const i64 Ni = prod(ishape, dim), No = prod(oshape, dim);
    i64 st[3];
    strides3(ishape, dim, st);
#pragma omp parallel for collapse(2) schedule(static)
    for (i64 b = 0; b < B; ++b)
        for (i64 v = 0; v < No; ++v) {
            support_t s;
            make_support(&s, grid + (b * No + v) * dim, dim, ishape, bound, order, extrapolate, 0);
            for (i64 c = 0; c < C; ++c) {
                const REAL *src = inp + (b * C + c) * Ni;
                REAL acc = 0;
                for (int i = 0; i < s.nnodes[0]; ++i)
                    for (int j = 0; j < s.nnodes[1]; ++j)
                        for (int k = 0; k < s.nnodes[2]; ++k) {
                            i64 idx = s.idx[0][i] * st[0] + s.idx[1][j] * st[1] + s.idx[2][k] * st[2];
                            acc += src[idx] * s.w[0][i] * s.w[1][j] * s.w[2][k];
                        }
                out[(b * C + c) * No + v] = s.inb ? acc : (REAL)0;
            }
        }
}

/* ------------------------------------------------------------------ */
/* grad -- nd.py:217-288; out (B,C,*oshape,dim)                       */
/* ------------------------------------------------------------------ */
void FN(orc_grad_pull)(const REAL *inp, const REAL *grid, REAL *out, i64 B, i64 C,
                       int dim, const i64 *ishape, const i64 *oshape,
                       const int *bound, const int *order, int extrapolate,
                       int quirk_linear_grad) {
    const i64 Ni = prod(ishape, dim), No = prod(oshape, dim);
    i64 st[3];
    strides3(ishape, dim, st);
#pragma omp parallel for collapse(2) schedule(static)
    for (i64 b = 0; b < B; ++b)
        for (i64 v = 0; v < No; ++v) {
            support_t s;
            make_support(&s, grid + (b * No + v) * dim, dim, ishape, bound, order, extrapolate, quirk_linear_grad);
            for (i64 c = 0; c < C; ++c) {
                const REAL *src = inp + (b * C + c) * Ni;
                REAL acc[3] = {0, 0, 0};
                for (int i = 0; i < s.nnodes[0]; ++i)
                    for (int j = 0; j < s.nnodes[1]; ++j)
                        for (int k = 0; k < s.nnodes[2]; ++k) {
                            i64 idx = s.idx[0][i] * st[0] + s.idx[1][j] * st[1] + s.idx[2][k] * st[2];
                            REAL val = src[idx];
                            acc[0] += val * s.g[0][i] * s.w[1][j] * s.w[2][k];
                            acc[1] += val * s.w[0][i] * s.g[1][j] * s.w[2][k];
                            acc[2] += val * s.w[0][i] * s.w[1][j] * s.g[2][k];
                        }
                for (int d = 0; d < dim; ++d)
                    out[((b * C + c) * No + v) * dim + d] = s.inb ? acc[d] : (REAL)0;
            }
        }
}

/* ------------------------------------------------------------------ */
/* hess -- nd.py:368-464; out (B,C,*oshape,dim,dim)                   */
/* ------------------------------------------------------------------ */
void FN(orc_hess_pull)(const REAL *inp, const REAL *grid, REAL *out, i64 B, i64 C,
                       int dim, const i64 *ishape, const i64 *oshape,
                       const int *bound, const int *order, int extrapolate,
                       int quirk_linear_grad) {
    const i64 Ni = prod(ishape, dim), No = prod(oshape, dim);
    i64 st[3];
    strides3(ishape, dim, st);
#pragma omp parallel for collapse(2) schedule(static)
    for (i64 b = 0; b < B; ++b)
        for (i64 v = 0; v < No; ++v) {
            support_t s;
            make_support(&s, grid + (b * No + v) * dim, dim, ishape, bound, order, extrapolate, quirk_linear_grad);
            for (i64 c = 0; c < C; ++c) {
                const REAL *src = inp + (b * C + c) * Ni;
                REAL acc[3][3] = {{0}};
                for (int i = 0; i < s.nnodes[0]; ++i)
                    for (int j = 0; j < s.nnodes[1]; ++j)
                        for (int k = 0; k < s.nnodes[2]; ++k) {
                            i64 idx = s.idx[0][i] * st[0] + s.idx[1][j] * st[1] + s.idx[2][k] * st[2];
                            REAL val = src[idx];
                            const int n[3] = {i, j, k};
                            for (int d = 0; d < dim; ++d)
                                for (int e = d; e < dim; ++e) {
                                    REAL t = val;
                                    for (int q = 0; q < dim; ++q) {
                                        if (d == e && q == d) t *= s.h[q][n[q]];
                                        else if (q == d || q == e) t *= s.g[q][n[q]];
                                        else t *= s.w[q][n[q]];
                                    }
                                    acc[d][e] += t;
                                }
                        }
                for (int d = 0; d < dim; ++d)
                    for (int e = 0; e < dim; ++e) {
                        REAL t = (d <= e) ? acc[d][e] : acc[e][d]; /* nd.py:459-461 */
                        out[(((b * C + c) * No + v) * dim + d) * dim + e] = s.inb ? t : (REAL)0;
                    }
            }
        }
}

/* ------------------------------------------------------------------ */
/* push -- nd.py:147-213.  inp (B,C,*gshape) or NULL (count: ones,    */
/* pushpull.py:123-124), grid (B,*gshape,dim) -> out (B,C,*oshape)    */
/* ------------------------------------------------------------------ */
void FN(orc_push)(const REAL *inp, const REAL *grid, REAL *out, i64 B, i64 C,
                  int dim, const i64 *gshape, const i64 *oshape,
                  const int *bound, const int *order, int extrapolate) {
    const i64 Ng = prod(gshape, dim), No = prod(oshape, dim);
    i64 st[3];
    strides3(oshape, dim, st);
    memset(out, 0, sizeof(REAL) * (size_t)(B * C * No));
    /* parallel over (b, c): every target volume has a single writer */
#pragma omp parallel for collapse(2) schedule(static)
    for (i64 b = 0; b < B; ++b)
        for (i64 c = 0; c < C; ++c) {
            REAL *dst = out + (b * C + c) * No;
            for (i64 v = 0; v < Ng; ++v) {
                support_t s;
                make_support(&s, grid + (b * Ng + v) * dim, dim, oshape, bound, order, extrapolate, 0);
                if (!s.inb) continue;
                REAL val = inp ? inp[(b * C + c) * Ng + v] : (REAL)1;
                for (int i = 0; i < s.nnodes[0]; ++i)
                    for (int j = 0; j < s.nnodes[1]; ++j)
                        for (int k = 0; k < s.nnodes[2]; ++k) {
                            i64 idx = s.idx[0][i] * st[0] + s.idx[1][j] * st[1] + s.idx[2][k] * st[2];
                            dst[idx] += val * s.w[0][i] * s.w[1][j] * s.w[2][k];
                        }
            }
        }
}

/* push with the source volume split in slabs over threads: used by the
 * CPU-baseline timing when B*C is smaller than the number of cores.  Each
 * thread scatters into a private copy which is then summed. */
void FN(orc_push_mt)(const REAL *inp, const REAL *grid, REAL *out, i64 B, i64 C,
                     int dim, const i64 *gshape, const i64 *oshape,
                     const int *bound, const int *order, int extrapolate,
                     int nthreads) {
    const i64 Ng = prod(gshape, dim), No = prod(oshape, dim);
    i64 st[3];
    strides3(oshape, dim, st);
    if (nthreads < 1) nthreads = 1;
    REAL *priv = (REAL *)calloc((size_t)nthreads * (size_t)No, sizeof(REAL));
    for (i64 b = 0; b < B; ++b)
        for (i64 c = 0; c < C; ++c) {
            memset(priv, 0, sizeof(REAL) * (size_t)nthreads * (size_t)No);
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (int t = 0; t < nthreads; ++t) {
                REAL *dst = priv + (size_t)t * (size_t)No;
                i64 lo = Ng * t / nthreads, hi = Ng * (t + 1) / nthreads;
                for (i64 v = lo; v < hi; ++v) {
                    support_t s;
                    make_support(&s, grid + (b * Ng + v) * dim, dim, oshape, bound, order, extrapolate, 0);
                    if (!s.inb) continue;
                    REAL val = inp ? inp[(b * C + c) * Ng + v] : (REAL)1;
                    for (int i = 0; i < s.nnodes[0]; ++i)
                        for (int j = 0; j < s.nnodes[1]; ++j)
                            for (int k = 0; k < s.nnodes[2]; ++k) {
                                i64 idx = s.idx[0][i] * st[0] + s.idx[1][j] * st[1] + s.idx[2][k] * st[2];
                                dst[idx] += val * s.w[0][i] * s.w[1][j] * s.w[2][k];
                            }
                }
            }
            REAL *o = out + (b * C + c) * No;
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (i64 v = 0; v < No; ++v) {
                REAL acc = 0;
                for (int t = 0; t < nthreads; ++t) acc += priv[(size_t)t * (size_t)No + v];
                o[v] = acc;
            }
        }
    free(priv);
}

/* ------------------------------------------------------------------ */
/* pushgrad -- nd.py:292-364.  inp (B,C,*gshape,dim) -> out (B,C,*oshape) */
/* ------------------------------------------------------------------ */
void FN(orc_pushgrad)(const REAL *inp, const REAL *grid, REAL *out, i64 B, i64 C,
                      int dim, const i64 *gshape, const i64 *oshape,
                      const int *bound, const int *order, int extrapolate,
                      int quirk_linear_grad) {
    const i64 Ng = prod(gshape, dim), No = prod(oshape, dim);
    i64 st[3];
    strides3(oshape, dim, st);
    memset(out, 0, sizeof(REAL) * (size_t)(B * C * No));
#pragma omp parallel for collapse(2) schedule(static)
    for (i64 b = 0; b < B; ++b)
        for (i64 c = 0; c < C; ++c) {
            REAL *dst = out + (b * C + c) * No;
            for (i64 v = 0; v < Ng; ++v) {
                support_t s;
                make_support(&s, grid + (b * Ng + v) * dim, dim, oshape, bound, order, extrapolate, quirk_linear_grad);
                if (!s.inb) continue;
                REAL val[3] = {0, 0, 0};
                for (int d = 0; d < dim; ++d) val[d] = inp[((b * C + c) * Ng + v) * dim + d];
                for (int i = 0; i < s.nnodes[0]; ++i)
                    for (int j = 0; j < s.nnodes[1]; ++j)
                        for (int k = 0; k < s.nnodes[2]; ++k) {
                            i64 idx = s.idx[0][i] * st[0] + s.idx[1][j] * st[1] + s.idx[2][k] * st[2];
                            REAL t = val[0] * s.g[0][i] * s.w[1][j] * s.w[2][k];
                            if (dim > 1) t += val[1] * s.w[0][i] * s.g[1][j] * s.w[2][k];
                            if (dim > 2) t += val[2] * s.w[0][i] * s.w[1][j] * s.g[2][k];
                            dst[idx] += t;
                        }
            }
        }
}

/* ------------------------------------------------------------------ */
/* Spline prefilter -- interpol/coeff.py                               */
/* ------------------------------------------------------------------ */

/* get_poles, coeff.py:35-65 */
int FN(orc_get_poles)(int order, double *poles) {
    switch (order) {
    case 0: case 1: return 0;
    case 2: poles[0] = sqrt(8.) - 3.; return 1;
    case 3: poles[0] = sqrt(3.) - 2.; return 1;
    case 4:
        poles[0] = sqrt(664. - sqrt(438976.)) + sqrt(304.) - 19.;
        poles[1] = sqrt(664. + sqrt(438976.)) - sqrt(304.) - 19.;
        return 2;
    case 5:
        poles[0] = sqrt(67.5 - sqrt(4436.25)) + sqrt(26.25) - 6.5;
        poles[1] = sqrt(67.5 + sqrt(4436.25)) - sqrt(26.25) - 6.5;
        return 2;
    case 6:
        poles[0] = -0.488294589303044755130118038883789062112279161239377608394;
        poles[1] = -0.081679271076237512597937765737059080653379610398148178525368;
        poles[2] = -0.00141415180832581775108724397655859252786416905534669851652709;
        return 3;
    case 7:
        poles[0] = -0.5352804307964381655424037816816460718339231523426924148812;
        poles[1] = -0.122554615192326690515272264359357343605486549427295558490763;
        poles[2] = -0.0091486948096082769285930216516478534156925639545994482648003;
        return 3;
    }
    return -1;
}

/* pole ** k as a tensor of the line's dtype.  Under TorchScript
 * torch.as_tensor(pole, dtype=...) rounds the double pole through float32
 * first (coeff.py:88,116,134,167,189), whatever REAL is. */
static REAL polepow(double pole, i64 k) {
    REAL p = (REAL)(float)pole;
    return (REAL)pow((double)p, (double)k);
}

#define X(i) line[(i) * stride]

/* dct1_initial, coeff.py:109-149 */
static REAL dct1_initial(const REAL *line, i64 n, i64 stride, double pole) {
    i64 max_iter = (i64)ceil(-30. / log(fabs(pole)));
    if (max_iter < n) {
        REAL acc = 0;
        for (i64 k = 0; k < max_iter; ++k) acc += X(k) * polepow(pole, k);
        return acc;
    }
    double polen = pow(pole, (double)(n - 1));
    REAL out = X(0) + (REAL)polen * X(n - 1);
    if (n > 2) {
        REAL acc = 0;
        for (i64 k = 1; k < n - 1; ++k) {
            REAL pk = polepow(pole, k);
            pk = pk + (REAL)(polen * polen) / pk;
            acc += X(k) * pk;
        }
        out = out + acc;
    }
    double pp = pow(pole, (double)(n - 1));
    return out / (REAL)(1 - pp * pp);
}

/* dct1_final, coeff.py:210-216 */
static REAL dct1_final(const REAL *line, i64 n, i64 stride, double pole) {
    REAL out = (REAL)pole * X(n - 2) + X(n - 1);
    return out * (REAL)(pole / (pole * pole - 1));
}

/* dct2_initial, coeff.py:153-179 */
static REAL dct2_initial(const REAL *line, i64 n, i64 stride, double pole) {
    double polen = pow(pole, (double)n);
    REAL acc = 0;
    for (i64 k = 0; k < n; ++k) {
        REAL pk = polepow(pole, k) + (REAL)polen * polepow(pole, n - 1 - k);
        acc += X(k) * pk;
    }
    acc = acc * (REAL)(pole / (1 - polen * polen));
    return acc + X(0);
}

/* dct2_final, coeff.py:220-227 */
static REAL dct2_final(const REAL *line, i64 n, i64 stride, double pole) {
    return X(n - 1) * (REAL)(pole / (pole - 1));
}

/* dft_initial, coeff.py:82-105 */
static REAL dft_initial(const REAL *line, i64 n, i64 stride, double pole) {
    i64 max_iter = (i64)ceil(-30. / log(fabs(pole)));
    if (max_iter > n) max_iter = n;
    REAL acc = 0;
    for (i64 j = 1; j < max_iter; ++j) acc += X(n - j) * polepow(pole, j);
    acc = acc + X(0);
    double pp = pow(pole, (double)max_iter);
    return acc / (REAL)(1 - pp);
}

/* dft_final, coeff.py:183-206 */
static REAL dft_final(const REAL *line, i64 n, i64 stride, double pole) {
    i64 max_iter = (i64)ceil(-30. / log(fabs(pole)));
    if (max_iter > n) max_iter = n;
    REAL acc = 0;
    for (i64 k = 0; k < max_iter - 1; ++k) acc += X(k) * polepow(pole, k + 2);
    acc = acc + (REAL)pole * X(n - 1);
    double pp = pow(pole, (double)max_iter);
    return acc / (REAL)(pp - 1);
}

/* filter one line in place, coeff.py:258-284 (bound map coeff.py:237-254) */
static int filter_line(REAL *line, i64 n, i64 stride, int bound, int order) {
    double poles[3];
    int np = FN(orc_get_poles)(order, poles);
    if (np <= 0 || n == 1) return 0;
    int kind;
    if (bound == 0 || bound == 2) kind = 1;       /* zero, dct1 */
    else if (bound == 1 || bound == 3) kind = 2;  /* replicate, dct2 */
    else if (bound == 6) kind = 6;                /* dft */
    else return -1;                               /* NotImplementedError */
    double gain = 1.;
    for (int p = 0; p < np; ++p) gain *= (1. - poles[p]) * (1. - 1. / poles[p]); /* coeff.py:69-73 */
    for (i64 i = 0; i < n; ++i) X(i) = X(i) * (REAL)gain;
    for (int p = 0; p < np; ++p) {
        double pole = poles[p];
        REAL init = kind == 1 ? dct1_initial(line, n, stride, pole)
                  : kind == 2 ? dct2_initial(line, n, stride, pole)
                              : dft_initial(line, n, stride, pole);
        X(0) = init;
        for (i64 i = 1; i < n; ++i) X(i) = X(i) + (REAL)pole * X(i - 1);
        REAL fin = kind == 1 ? dct1_final(line, n, stride, pole)
                 : kind == 2 ? dct2_final(line, n, stride, pole)
                             : dft_final(line, n, stride, pole);
        X(n - 1) = fin;
        for (i64 i = n - 2; i >= 0; --i) X(i) = (REAL)pole * (X(i + 1) - X(i));
    }
    return 0;
}
#undef X

/* spline_coeff along one axis of a contiguous (outer, n, inner) view, in
 * place.  coeff.py:288-313.  Returns 0, or -1 for an unsupported bound. */
int FN(orc_spline_coeff)(REAL *x, i64 outer, i64 n, i64 inner, int bound, int order) {
    if (order <= 1) return 0;
    if (!(bound == 0 || bound == 1 || bound == 2 || bound == 3 || bound == 6)) return -1;
    int status = 0;
#pragma omp parallel for collapse(2) schedule(static)
    for (i64 o = 0; o < outer; ++o)
        for (i64 k = 0; k < inner; ++k) {
            int r = filter_line(x + o * n * inner + k, n, inner, bound, order);
            if (r) status = r;
        }
    return status;
}
