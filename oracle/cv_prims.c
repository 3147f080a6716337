/*
 * oracle/cv_prims.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see cv_prims.h).
 *
 * Plain-C restatement of the OpenCV 4.13 8-bit primitives and glibc 2.39 sinf/cosf
 * used by the reference's ORB front end.  Pinned against
 * Python cv2 4.13.0 / libm by tests/test_cv_prims.py and tests/golden/.
 *
 * Build with -ffp-contract=off (no FMA contraction): every float/double operation
 * below must round individually.
 */
#include "cv_prims.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

int cvp_round(double v) { return (int)nearbyint(v); } /* default mode = half-to-even */

static inline int refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

/* ------------------------------------------------------------------ resize */
/* OpenCV imgproc/resize.cpp, INTER_LINEAR, uchar: ialpha/ibeta are
 * saturate_cast<short>(w * 2048); HResizeLinear accumulates in int;
 * VResizeLinear<uchar,int,short> does ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2)>>2. */
void cvp_resize_axis_table(int n_src, int n_dst, int* ofs, short* c0, short* c1)
{
    double inv_scale = (double)n_dst / (double)n_src;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < n_dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }
        ofs[d] = s;
        c0[d] = (short)cvp_round((double)((1.f - f) * 2048.f));
        c1[d] = (short)cvp_round((double)(f * 2048.f));
    }
}

void cvp_resize_linear_8u(const uint8_t* src, int sw, int sh, size_t sstep,
                          uint8_t* dst, int dw, int dh, size_t dstep)
{
    int* xofs = (int*)malloc(sizeof(int) * (size_t)dw);
    int* yofs = (int*)malloc(sizeof(int) * (size_t)dh);
    short* xa0 = (short*)malloc(sizeof(short) * (size_t)dw * 2);
    short* xa1 = xa0 + dw;
    short* ya0 = (short*)malloc(sizeof(short) * (size_t)dh * 2);
    short* ya1 = ya0 + dh;
    int* rowbuf = (int*)malloc(sizeof(int) * (size_t)dw * 2);
    int* rows[2] = { rowbuf, rowbuf + dw };
    int rowid[2] = { -1, -1 };

    cvp_resize_axis_table(sw, dw, xofs, xa0, xa1);
    cvp_resize_axis_table(sh, dh, yofs, ya0, ya1);

    for (int y = 0; y < dh; ++y) {
        int sy0 = yofs[y];
        int sy1 = sy0 + 1 < sh ? sy0 + 1 : sh - 1;
        /* two-slot cache of horizontally resized source rows */
        if (rowid[1] == sy0 && rowid[0] != sy0) {
            int* t = rows[0]; rows[0] = rows[1]; rows[1] = t;
            rowid[0] = rowid[1]; rowid[1] = -1;
        }
        for (int k = 0; k < 2; ++k) {
            int want = k ? sy1 : sy0;
            if (rowid[k] == want) continue;
            if (k == 1 && want == sy0) { memcpy(rows[1], rows[0], sizeof(int) * (size_t)dw); rowid[1] = want; continue; }
            const uint8_t* S = src + (size_t)want * sstep;
            int* D = rows[k];
            for (int x = 0; x < dw; ++x) {
                int sx = xofs[x];
                int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
                D[x] = S[sx] * xa0[x] + S[sx1] * xa1[x];
            }
            rowid[k] = want;
        }
        uint8_t* D = dst + (size_t)y * dstep;
        int b0 = ya0[y], b1 = ya1[y];
        const int* H0 = rows[0];
        const int* H1 = rows[1];
        for (int x = 0; x < dw; ++x)
            D[x] = (uint8_t)((((b0 * (H0[x] >> 4)) >> 16) + ((b1 * (H1[x] >> 4)) >> 16) + 2) >> 2);
    }
    free(rowbuf); free(ya0); free(xa0); free(yofs); free(xofs);
}

/* ------------------------------------------------------------ border */
void cvp_border_reflect101(const uint8_t* src, int w, int h, size_t sstep,
                           uint8_t* dst, size_t dstep, int b)
{
    /* interior first (memmove: src may be the interior of dst), then the sides of
     * every interior row, then whole border rows copied from finished rows. */
    for (int y = 0; y < h; ++y) {
        uint8_t* drow = dst + (size_t)(y + b) * dstep;
        const uint8_t* srow = src + (size_t)y * sstep;
        if (drow + b != srow) memmove(drow + b, srow, (size_t)w);
    }
    for (int y = 0; y < h; ++y) {
        uint8_t* drow = dst + (size_t)(y + b) * dstep;
        for (int i = 0; i < b; ++i) {
            drow[i] = drow[b + refl101(i - b, w)];
            drow[b + w + i] = drow[b + refl101(w + i, w)];
        }
    }
    for (int i = 0; i < b; ++i) {
        memcpy(dst + (size_t)i * dstep, dst + (size_t)(b + refl101(i - b, h)) * dstep, (size_t)(w + 2 * b));
        memcpy(dst + (size_t)(b + h + i) * dstep, dst + (size_t)(b + refl101(h + i, h)) * dstep, (size_t)(w + 2 * b));
    }
}

/* ------------------------------------------------------------ GaussianBlur */
/* OpenCV 4.13 8U fixed-point Gaussian (smooth.dispatch/simd): ksize 7, sigma 2 gives
 * the integer kernel {18,34,48,56,48,34,18}/256 per axis; single rounding
 * (acc + 2^15) >> 16 after the vertical pass. */
void cvp_gaussian7x7_s2(const uint8_t* src, int w, int h, size_t sstep,
                        uint8_t* dst, size_t dstep)
{
    static const int K[7] = { 18, 34, 48, 56, 48, 34, 18 };
    uint16_t* hbuf = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)w * (size_t)h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* S = src + (size_t)y * sstep;
        uint16_t* H = hbuf + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            if (x >= 3 && x + 3 < w) {
                for (int i = -3; i <= 3; ++i) acc += K[i + 3] * S[x + i];
            } else {
                for (int i = -3; i <= 3; ++i) acc += K[i + 3] * S[refl101(x + i, w)];
            }
            H[x] = (uint16_t)acc; /* <= 255*256 */
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* R[7];
        for (int j = -3; j <= 3; ++j) R[j + 3] = hbuf + (size_t)refl101(y + j, h) * w;
        uint8_t* D = dst + (size_t)y * dstep;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 32768u;
            for (int j = 0; j < 7; ++j) acc += (uint32_t)K[j] * R[j][x];
            D[x] = (uint8_t)(acc >> 16);
        }
    }
    free(hbuf);
}

/* ------------------------------------------------------------------ FAST */
/* OpenCV features2d/fast.cpp (FAST_t<16>) + fast_score.cpp (cornerScore<16>).
 * Ring in OpenCV order; a pixel is a corner at threshold t iff 9 contiguous ring
 * pixels are all > v+t or all < v-t, iff score >= t where
 * score = max(max_k min_{m<9} d[k+m], max_k min_{m<9} -d[k+m]) - 1, d = v - ring. */
static const int RING_DX[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
static const int RING_DY[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

static inline void ring_offsets(size_t step, ptrdiff_t off[16])
{
    for (int k = 0; k < 16; ++k) off[k] = (ptrdiff_t)RING_DY[k] * (ptrdiff_t)step + RING_DX[k];
}

static inline int score_from_d(const int d[16])
{
    int e[32], m2[32], m4[32], m8;
    for (int k = 0; k < 16; ++k) e[k] = e[k + 16] = d[k];
    int best_pos = -1000, best_neg = 1000;
    /* sliding min / max of window 9 over the circular ring */
    for (int k = 0; k < 24; ++k) m2[k] = e[k] < e[k + 1] ? e[k] : e[k + 1];
    for (int k = 0; k < 22; ++k) m4[k] = m2[k] < m2[k + 2] ? m2[k] : m2[k + 2];
    for (int k = 0; k < 16; ++k) {
        m8 = m4[k] < m4[k + 4] ? m4[k] : m4[k + 4];
        int m9 = m8 < e[k + 8] ? m8 : e[k + 8];
        if (m9 > best_pos) best_pos = m9;
    }
    for (int k = 0; k < 24; ++k) m2[k] = e[k] > e[k + 1] ? e[k] : e[k + 1];
    for (int k = 0; k < 22; ++k) m4[k] = m2[k] > m2[k + 2] ? m2[k] : m2[k + 2];
    for (int k = 0; k < 16; ++k) {
        m8 = m4[k] > m4[k + 4] ? m4[k] : m4[k + 4];
        int m9 = m8 > e[k + 8] ? m8 : e[k + 8];
        if (m9 < best_neg) best_neg = m9;
    }
    int s = best_pos > -best_neg ? best_pos : -best_neg;
    return s - 1;
}

int cvp_fast9_score(const uint8_t* p, size_t step)
{
    ptrdiff_t off[16];
    int d[16];
    ring_offsets(step, off);
    for (int k = 0; k < 16; ++k) d[k] = (int)p[0] - (int)p[off[k]];
    return score_from_d(d);
}

/* 9-contiguous test on a 16-bit ring mask */
static inline int has_arc9(unsigned m)
{
    m |= m << 16;
    unsigned x = m & (m >> 1);
    x &= x >> 2;
    x &= x >> 4;
    x &= m >> 8;
    return (x & 0xffffu) != 0;
}

int cvp_fast9_nms(const uint8_t* img, int w, int h, size_t step, int threshold,
                  cvp_corner* out, int cap)
{
    if (w < 7 || h < 7) return 0;
    ptrdiff_t off[16];
    ring_offsets(step, off);
    /* score buffer with a 1-pixel zero frame around the evaluated rectangle
     * [3,w-3) x [3,h-3): index (y-2, x-2) */
    const int bw = w - 4, bh = h - 4;
    uint8_t* buf = (uint8_t*)calloc((size_t)bw * (size_t)bh, 1);
    const int t = threshold;
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t* row = img + (size_t)y * step;
        uint8_t* brow = buf + (size_t)(y - 2) * bw - 2;
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* p = row + x;
            const int v = p[0];
            const int hi = v + t, lo = v - t;
            /* antipodal early exit: a 9-arc contains one pixel of every opposite pair */
            int a = p[off[0]], b = p[off[8]];
            int br = (a > hi) | (b > hi), dk = (a < lo) | (b < lo);
            if (!(br | dk)) continue;
            a = p[off[4]]; b = p[off[12]];
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            if (!(br | dk)) continue;
            a = p[off[2]]; b = p[off[10]];
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            a = p[off[6]]; b = p[off[14]];
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            if (!(br | dk)) continue;
            unsigned mb = 0, md = 0;
            int d[16];
            for (int k = 0; k < 16; ++k) {
                int r = p[off[k]];
                d[k] = v - r;
                mb |= (unsigned)(r > hi) << k;
                md |= (unsigned)(r < lo) << k;
            }
            if (!((br && has_arc9(mb)) || (dk && has_arc9(md)))) continue;
            brow[x] = (uint8_t)score_from_d(d);
        }
    }
    int n = 0;
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t* c = buf + (size_t)(y - 2) * bw - 2;
        const uint8_t* u = c - bw;
        const uint8_t* l = c + bw;
        for (int x = 3; x < w - 3; ++x) {
            int s = c[x];
            if (!s) continue;
            if (s > c[x - 1] && s > c[x + 1] && s > u[x - 1] && s > u[x] && s > u[x + 1] &&
                s > l[x - 1] && s > l[x] && s > l[x + 1]) {
                if (n < cap) { out[n].x = x; out[n].y = y; out[n].score = s; }
                ++n;
            }
        }
    }
    free(buf);
    return n;
}

/* ------------------------------------------------------------ fastAtan2 */
/* OpenCV core/mathfuncs_core.simd.hpp atanImpl<float> (scalar), degrees. */
float cvp_fast_atan2(float y, float x)
{
    const float S = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * S;
    const float p3 = -0.3258083974640975f * S;
    const float p5 = 0.1555786518463281f * S;
    const float p7 = -0.04432655554792128f * S;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ------------------------------------------------------------ sinf / cosf */
/* glibc 2.39 sysdeps/ieee754/flt-32/{s_sinf.c,s_cosf.c,sincosf.h}: argument reduction
 * by 2/pi * 2^24 and degree-7/8 polynomials evaluated in double. */
static const double HPI_INV = 0x1.45F306DC9C883p+23;
static const double HPI = 0x1.921FB54442D18p0;
static const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
                    C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
static const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7,
                    S3 = -0x1.994eb3774cf24p-13;

static inline uint32_t abstop12(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    return (u >> 20) & 0x7ff;
}

static inline float sincos_poly(double x, double x2, int negcos, int n)
{
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = S2 + x2 * S3;
        double x7 = x3 * x2;
        double s = x + x3 * S1;
        return (float)(s + x7 * s1);
    } else {
        double sg = negcos ? -1.0 : 1.0;
        double x4 = x2 * x2;
        double c2 = sg * C3 + x2 * (sg * C4);
        double c1 = sg * C0 + x2 * (sg * C1);
        double x6 = x4 * x2;
        double c = c1 + x4 * (sg * C2);
        return (float)(c + x6 * c2);
    }
}

static float sincosf_impl(float y, int want_cos)
{
    double x = y;
    if (abstop12(y) < 0x3f4) { /* |y| < pi/4 */
        double s = x * x;
        if (abstop12(y) < 0x398) /* |y| < 2^-12 */
            return want_cos ? 1.0f : y;
        return sincos_poly(x, s, 0, want_cos);
    }
    if (abstop12(y) < 0x42f) { /* |y| < 120 */
        double r = x * HPI_INV;
        int n = ((int32_t)r + 0x800000) >> 24;
        x = x - n * HPI;
        static const double sign[4] = { 1.0, -1.0, -1.0, 1.0 };
        double s = sign[n & 3];
        return sincos_poly(x * s, x * x, (n & 2) != 0, want_cos ? (n ^ 1) : n);
    }
    return want_cos ? cosf(y) : sinf(y); /* never reached on the ORB path (angles in [0, 2pi)) */
}

float cvp_sinf(float x) { return sincosf_impl(x, 0); }
float cvp_cosf(float x) { return sincosf_impl(x, 1); }

/* ------------------------------------------------------------------------------------------------
 * glibc 2.39 logf (sysdeps/ieee754/flt-32/e_logf.c, e_logf_data.c; table bits = 4, degree-3
 * polynomial, double arithmetic, operations in the source's order, no FMA).  Called by
 * MapPoint::PredictScale (src/MapPoint.cc:450, :468: log(float) resolves to the float overload)
 * and for mfLogScaleFactor (src/Frame.cc:75).  The table was checked against the bytes of this
 * image's libm.so.6; tests compare the function with libm's logf. */
static const double logf_tab[16][2] = {
    { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 }, { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
    { 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2 },  { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
    { 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 }, { 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3 },
    { 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 }, { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
    { 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 }, { 0x1p+0, 0x0p+0 },
    { 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 },  { 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4 },
    { 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 },  { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3 },
    { 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 },  { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 },
};
static const double logf_ln2 = 0x1.62e42fefa39efp-1;
static const double logf_poly[3] = { -0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2 };

float cvp_logf(float x)
{
    uint32_t ix;
    memcpy(&ix, &x, 4);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2u == 0) return -INFINITY;                        /* log(+-0) */
        if (ix == 0x7f800000u) return x;                           /* log(inf) */
        if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return NAN;
        x = x * 0x1p23f;                                           /* subnormal: normalise */
        memcpy(&ix, &x, 4);
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> (23 - 4)) % 16u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & (0x1ffu << 23));
    float zf;
    memcpy(&zf, &iz, 4);
    const double invc = logf_tab[i][0], logc = logf_tab[i][1], z = (double)zf;
    const double r = z * invc - 1;
    const double y0 = logc + (double)k * logf_ln2;
    const double r2 = r * r;
    double y = logf_poly[1] * r + logf_poly[2];
    y = logf_poly[0] * r2 + y;
    y = y * r2 + (y0 + r);
    return (float)y;
}
