/*
 * TEST INFRASTRUCTURE ONLY (see compv_oracle.c).  CPU restatement of CompVGradientFast and the S-HOG descriptor (scalar C paths).
 *   gradients [-1 0 1], zero border               base/compv_gradient_fast.cxx:88-99, 243-253, 396-433
 *   hypot_naive, fastAtan2 (degrees)             base/math/compv_math_trig.cxx:411-446, 496-510; constants base/math/compv_math.cxx:39-43
 *   cell binning (nearest / bilinear / LUT)      core/features/hog/compv_core_feature_hog_std.cxx:564-743; LUT: core/include/.../compv_core_feature_hog_std.h:50-88
 *   cell grid, blocks, descriptor size           hog_std.cxx:196-393; base/compv_features.cxx:274-299
 *   block norms with 8-lane partial sums         core/include/compv/core/features/hog/compv_core_feature_hog_common_norm.h:22-143
 * Float tolerance: the compiled reference runs AVX2/FMA leaves for magnitude/direction/binning, so reference vs oracle is compared at 1e-4 (the
 * reference keeps separate `md5_fma` goldens for the same reason, unittests/hog_s.cxx:24-41); oracle vs CUDA is bit-exact (same scalar order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define V volatile float /* every float operation rounds to fp32 on its own */

static float atan2_deg(float y, float x)
{
    const float eps = (float)2.2204460492503131e-016, p1 = 57.2836266f, p3 = -18.6674461f, p5 = 8.91400051f, p7 = -2.53972459f;
    const float ax = fabsf(x), ay = fabsf(y);
    V a, c, c2, t;
    if (ax >= ay) { t = ax + eps; c = ay / t; c2 = c * c; t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1; a = t * c; }
    else { t = ay + eps; c = ax / t; c2 = c * c; t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1; t = t * c; a = 90.f - t; }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

static void grad_px(const uint8_t* in, size_t stride, size_t w, size_t h, size_t x, size_t y, float* gx, float* gy)
{
    const uint8_t* p = in + y * stride + x;
    *gx = (x >= 1 && x + 1 < w) ? (float)((int)p[1] - (int)p[-1]) : 0.f;
    *gy = (y >= 1 && y + 1 < h) ? (float)((int)p[stride] - (int)p[-(long)stride]) : 0.f;
}

ORC_API int orc_gradient_fast_8u(const uint8_t* in, size_t w, size_t h, size_t stride, int16_t* gx16, int16_t* gy16, float* gx32, float* gy32, float* mag, float* dir)
{
    if (!in || !w || !h || stride < w) return 20006;
    for (size_t y = 0; y < h; ++y) for (size_t x = 0; x < w; ++x) {
        float gx, gy;
        grad_px(in, stride, w, h, x, y, &gx, &gy);
        const size_t o = y * stride + x;
        if (gx16) gx16[o] = (int16_t)gx;
        if (gy16) gy16[o] = (int16_t)gy;
        if (gx32) gx32[o] = gx;
        if (gy32) gy32[o] = gy;
        if (mag) { V a = gx * gx, b = gy * gy, c = a + b; mag[o] = sqrtf(c); }
        if (dir) dir[o] = atan2_deg(gy, gx);
    }
    return 0;
}

static float den8(const float* v, size_t n, int sq)
{
    const size_t n8 = n & ~(size_t)7, n4 = n & ~(size_t)3;
    V d[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    size_t i;
    for (i = 0; i < n8; i += 8) for (int k = 0; k < 8; ++k) { V t = sq ? v[i + k] * v[i + k] : v[i + k]; d[k] = d[k] + t; }
    for (; i < n4; i += 4) for (int k = 0; k < 4; ++k) { V t = sq ? v[i + k] * v[i + k] : v[i + k]; d[k] = d[k] + t; }
    d[0] = d[0] + d[4]; d[1] = d[1] + d[5]; d[2] = d[2] + d[6]; d[3] = d[3] + d[7];
    d[0] = d[0] + d[2]; d[1] = d[1] + d[3];
    d[0] = d[0] + d[1];
    for (; i < n; ++i) { V t = sq ? v[i] * v[i] : v[i]; d[0] = d[0] + t; }
    return d[0];
}
static void norm_l1(float* v, size_t n, float eps) { V s = den8(v, n, 0) + eps; V den = 1.f / s; for (size_t i = 0; i < n; ++i) { V t = v[i] * den; v[i] = t; } }
static void norm_l2(float* v, size_t n, float eps2) { V s = den8(v, n, 1) + eps2; V q = sqrtf(s); V den = 1.f / q; for (size_t i = 0; i < n; ++i) { V t = v[i] * den; v[i] = t; } }

/* blockNorm / interp use the reference enum values (48..52, 53..55). Returns the descriptor length through *size. */
ORC_API int orc_hog(const uint8_t* in, size_t w, size_t h, size_t stride, size_t bw, size_t bh, size_t sw, size_t sh, size_t cw, size_t ch, size_t nbins,
                    int blockNorm, int gradientSigned, int interp, float* out, size_t capacity, size_t* size)
{
    if (!in || !size || !bw || !bh || !sw || !sh || !cw || !ch || nbins < 2 || nbins > 360 || (bw % cw) || (bh % ch) || (cw % sw) || (ch % sh) || w < bw || h < bh) return 20006;
    const float sx = sw / (float)cw, sy = sh / (float)ch;
    const int ncx = (int)(w / cw / sx), ncy = (int)(h / ch / sy);
    const size_t xo = (size_t)(cw * sx), yo = (size_t)(ch * sy);
    const int xg = (((size_t)(ncx - 1) * xo) + cw) > w, yg = (((size_t)(ncy - 1) * yo) + ch) > h;
    const size_t pitch = (size_t)ncx * nbins;
    float* map = (float*)calloc(pitch * ncy + 1, sizeof(float));
    if (!map) return 20013;
    const float thetaMax = gradientSigned ? 360.f : 180.f;
    const int binWidth = (gradientSigned ? 360 : 180) / (int)nbins;
    V scale = 1.f / (float)binWidth;
    const int binMax = (int)nbins - 1;
    /* LUT (0.1 degree) */
    struct { float diff; int bin, next; } *lut = NULL;
    if (interp == 54) {
        const size_t cnt = (size_t)((thetaMax + 1) * 10) + 16;
        lut = calloc(cnt, sizeof(*lut));
        size_t k = 0;
        for (float th = 0.f; th <= thetaMax + 1 && k < cnt; th += 0.1f, ++k) {
            const int b = (int)((th * scale) - 0.5f);
            const float df = ((th - (b * binWidth)) * scale) - 0.5f;
            const int nx = b + ((df >= 0) ? 1 : -1);
            lut[k].bin = b; lut[k].next = nx < 0 ? binMax : (nx > binMax ? 0 : nx); lut[k].diff = df;
        }
    }
    for (int cj = 0; cj < ncy - yg; ++cj) for (int ci = 0; ci < ncx - xg; ++ci) {
        float* hist = map + (size_t)cj * pitch + (size_t)ci * nbins;
        for (size_t j = 0; j < ch; ++j) for (size_t i = 0; i < cw; ++i) {
            float gx, gy;
            grad_px(in, stride, w, h, ci * xo + i, cj * yo + j, &gx, &gy);
            V a = gx * gx, b2 = gy * gy, c = a + b2;
            const float m = sqrtf(c);
            const float d = atan2_deg(gy, gx);
            V theta = (d > thetaMax) ? (d - thetaMax) : d;
            if (interp == 53) { V t = theta * scale; V r = hist[(int)t] + m; hist[(int)t] = r; }
            else if (interp == 55) {
                V t = theta * scale; t = t - 0.5f;
                const int b = (int)t;
                V u = theta - (float)(b * binWidth); u = u * scale; V diff = u - 0.5f;
                V vv = m * diff;
                if (diff >= 0) { float* q = &hist[b == binMax ? 0 : b + 1]; V r = *q + vv; *q = r; V s = m - vv; V r2 = hist[b] + s; hist[b] = r2; }
                else { float* q = &hist[b ? b - 1 : binMax]; V r = *q - vv; *q = r; V s = m + vv; V r2 = hist[b] + s; hist[b] = r2; }
            }
            else {
                V t = theta * 10.f; t = t + 0.5f;
                const int k = (int)t;
                V av = m * lut[k].diff; av = fabsf(av);
                V r = hist[lut[k].next] + av; hist[lut[k].next] = r;
                V s = m - av; V r2 = hist[lut[k].bin] + s; hist[lut[k].bin] = r2;
            }
        }
    }
    size_t nbx = 0, nby = 0;
    for (size_t x = 0; x <= w - bw; x += sw) ++nbx;
    for (size_t y = 0; y <= h - bh; y += sh) ++nby;
    const size_t cbx = bw / cw, cby = bh / ch, binsX = cbx * nbins, n = cby * binsX;
    *size = n * nbx * nby;
    int rc = 0;
    if (out) {
        if (capacity < *size) rc = 20014;
        else {
            const size_t xBinOff = nbins * (size_t)(sx + 0.5), yStep = (size_t)(sy + 0.5);
            const float eps = 1e-6f; V eps2 = eps * eps;
            for (size_t by = 0; by < nby; ++by) for (size_t bx = 0; bx < nbx; ++bx) {
                float* o = out + (by * nbx + bx) * n;
                const float* src = map + by * yStep * pitch + bx * xBinOff;
                for (size_t cy = 0; cy < cby; ++cy) memcpy(o + cy * binsX, src + cy * pitch, binsX * sizeof(float));
                switch (blockNorm) {
                case 49: norm_l1(o, n, eps); break;
                case 50: norm_l1(o, n, eps); for (size_t i = 0; i < n; ++i) o[i] = sqrtf(o[i]); break;
                case 51: norm_l2(o, n, eps2); break;
                case 52: norm_l2(o, n, eps2); for (size_t i = 0; i < n; ++i) if (o[i] > 0.2f) o[i] = 0.2f; norm_l2(o, n, eps2); break;
                default: break;
                }
            }
        }
    }
    free(map); free(lut);
    return rc;
}
