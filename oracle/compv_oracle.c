/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the reference algorithms on CompV's image hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library; the product (compv_b200/) never does.
 *
 * Pinning: every function here is checked (tests/test_oracle_vs_ref.py) against the UNMODIFIED reference compiled into
 * oracle/_ref/libcompv_ref.so, and the convolution functions against the reference's own golden MD5s
 * (unittests/math_convlt.cxx:17-26, reproduced in tests/test_golden_convlt.py).
 *
 * Paths are relative to /root/reference.  Each function names the reference lines it restates; the code is written
 * from the arithmetic contract (scalar loops, no SIMD, no threading), it is not a copy of the reference text.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static int g_use_fma = 1;

/* The reference's AVX2 float leaves are compiled with -mfma and GCC contracts mul+add (they reproduce the `md5_fma` goldens,
 * unittests/math_convlt.cxx:17-26).  use_fma=0 gives the plain C++ path (compv_math_convlt.h:358-384) for triangulation. */
ORC_API void orc_set_fma(int use_fma) { g_use_fma = use_fma; }

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------------------------------
 * K1/K2/K3 separable correlation.  base/include/compv/base/math/compv_math_convlt.h
 *   driver + ordering (hz pass then vt pass, intermediate plane of OutputType)   :98-173
 *   horizontal border handling (zero / replicate / ignore)                       :176-229
 *   vertical border handling                                                     :231-292
 *   integer arithmetic  (int accumulate, clip to OutputType)                     :332-353
 *   float arithmetic    (tap order, clip, truncation)                            :358-384
 *   fixed point         (sum of (v*k)>>16, clip 0..255)                          :386-405
 * border: 0 zero, 1 ignore, 2 replicate (COMPV_BORDER_TYPE).
 * ---------------------------------------------------------------------------------------------- */
#define ORC_CONV_IMPL(NAME, IN_T, K_T, OUT_T, SAMPLE_EXPR_DECL)                                                            \
ORC_API int orc_convlt1_##NAME(const IN_T* in, size_t w, size_t h, size_t stride, const K_T* vt, const K_T* hz, size_t ks,   \
                               OUT_T* out, int border)                                                                     \
{                                                                                                                          \
    if (!in || !out || !vt || !hz || !(ks & 1) || w < ks || h < ks || stride < w) return 20006;                            \
    const size_t r = ks >> 1;                                                                                              \
    OUT_T* mid = (OUT_T*)calloc(stride * h, sizeof(OUT_T));                                                                \
    if (!mid) return 20013;                                                                                                \
    for (size_t y = 0; y < h; ++y) {                                                                                       \
        for (size_t x = 0; x < w; ++x) {                                                                                   \
            OUT_T m = 0;                                                                                                   \
            if (x >= r && x < w - r) { const IN_T* p = &in[y * stride + x - r]; const size_t step = 1; const K_T* taps = hz; SAMPLE_EXPR_DECL }  \
            else if (border == 2) m = (OUT_T)in[y * stride + x];                                                           \
            mid[y * stride + x] = m;                                                                                       \
        }                                                                                                                  \
    }                                                                                                                      \
    for (size_t y = 0; y < h; ++y) {                                                                                       \
        for (size_t x = 0; x < w; ++x) {                                                                                   \
            OUT_T m = 0;                                                                                                   \
            if (y >= r && y < h - r) {                                                                                     \
                if (border == 1 && (x < r || x >= w - r)) continue;                                                        \
                const OUT_T* p = &mid[(y - r) * stride + x]; const size_t step = stride; const K_T* taps = vt; SAMPLE_EXPR_DECL  \
                out[y * stride + x] = m;                                                                                   \
            }                                                                                                              \
            else if (border == 0) out[y * stride + x] = 0;                                                                 \
            else if (border == 2) out[y * stride + x] = mid[y * stride + x];                                               \
        }                                                                                                                  \
    }                                                                                                                      \
    free(mid);                                                                                                             \
    return 0;                                                                                                              \
}

#define ORC_SAMPLE_INT16 { int acc = 0; for (size_t k = 0; k < ks; ++k) acc += (int)p[k * step] * (int)taps[k]; m = (int16_t)clampi(acc, -32768, 32767); }
#define ORC_SAMPLE_F32_ACC float acc = 0.f; for (size_t k = 0; k < ks; ++k) { const float v = (float)p[k * step]; if (g_use_fma) acc = fmaf(v, taps[k], acc); else { volatile float prod = v * taps[k]; acc = k ? (acc + prod) : prod; } }
#define ORC_SAMPLE_F32_U8 { ORC_SAMPLE_F32_ACC acc = acc < 0.f ? 0.f : (acc > 255.f ? 255.f : acc); m = (uint8_t)(int)acc; }
#define ORC_SAMPLE_F32_F32 { ORC_SAMPLE_F32_ACC m = acc; }
#define ORC_SAMPLE_FXP { unsigned acc = 0; for (size_t k = 0; k < ks; ++k) acc += ((unsigned)p[k * step] * (unsigned)taps[k]) >> 16; m = (uint8_t)(acc > 255u ? 255u : acc); }

ORC_CONV_IMPL(8u16s16s, uint8_t, int16_t, int16_t, ORC_SAMPLE_INT16)
ORC_CONV_IMPL(16s16s16s, int16_t, int16_t, int16_t, ORC_SAMPLE_INT16)
ORC_CONV_IMPL(8u32f8u, uint8_t, float, uint8_t, ORC_SAMPLE_F32_U8)
ORC_CONV_IMPL(8u32f32f, uint8_t, float, float, ORC_SAMPLE_F32_F32)
ORC_CONV_IMPL(32f32f32f, float, float, float, ORC_SAMPLE_F32_F32)
ORC_CONV_IMPL(32f32f8u, float, float, uint8_t, ORC_SAMPLE_F32_U8)
ORC_CONV_IMPL(fxp_8u16u8u, uint8_t, uint16_t, uint8_t, ORC_SAMPLE_FXP)

/* CompVMathGauss::kernelDim1<float> -- base/include/compv/base/math/compv_math_gauss.h:23-56 */
ORC_API int orc_gauss_kernel_dim1_32f(size_t size, float sigma, float* kernel)
{
    if (!kernel || !(size & 1)) return 20006;
    const size_t c = size >> 1;
    const float two_sigma2 = (float)(2 * (sigma * sigma));
    const float peak = (float)(1 / sqrt(M_PI * two_sigma2));
    float total = peak;
    kernel[c] = peak;
    for (size_t x = 1; x <= c; ++x) {
        const float k = (float)(peak * exp(-(double)((x * x) / two_sigma2)));
        kernel[c + x] = k;
        kernel[c - x] = k;
        total += (k + k);
    }
    total = 1 / total;
    for (size_t x = 0; x < size; ++x) kernel[x] *= total;
    return 0;
}

/* CompVMathGauss::kernelDim1FixedPoint (base/math/compv_math_gauss.cxx:11-17) + fixedPointKernel (compv_math_convlt.h:76-92) */
ORC_API int orc_gauss_kernel_dim1_fxp(size_t size, float sigma, uint16_t* kernel)
{
    float tmp[1024];
    if (!kernel || size > 1024) return 20006;
    int rc = orc_gauss_kernel_dim1_32f(size, sigma, tmp);
    if (rc) return rc;
    for (size_t x = 0; x < size; ++x) kernel[x] = (uint16_t)(tmp[x] * 0xffff);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Kernel tables: base/include/compv/base/compv_features.h:124-133.  which: 0 Sobel 1 Scharr 2 Prewitt 3 Canny(=Sobel 3 or 5)
 * ---------------------------------------------------------------------------------------------- */
static int edge_tables(int which, int ks, const int16_t** vt, const int16_t** hz)
{
    static const int16_t sobel3_vt[3] = { 1, 2, 1 }, sobel3_hz[3] = { -1, 0, 1 };
    static const int16_t sobel5_vt[5] = { 1, 4, 6, 4, 1 }, sobel5_hz[5] = { 1, 2, 0, -2, -1 };
    static const int16_t scharr_vt[3] = { 3, 10, 3 }, prewitt_vt[3] = { 1, 1, 1 };
    switch (which) {
    case 0: case 3:
        if (ks == 5) { *vt = sobel5_vt; *hz = sobel5_hz; return 5; }
        *vt = sobel3_vt; *hz = sobel3_hz; return 3;
    case 1: *vt = scharr_vt; *hz = sobel3_hz; return 3;
    case 2: *vt = prewitt_vt; *hz = sobel3_hz; return 3;
    default: return 0;
    }
}

/* K1+K4: Gx = convlt1(vt=tableVt, hz=tableHz), Gy = convlt1(vt=tableHz, hz=tableVt), G = |Gx|+|Gy| saturated to u16.
 * core/features/edges/compv_core_feature_canny_dete.cxx:236-240, base/include/compv/base/math/compv_math_utils.h:173-186,
 * SIMD leaf saturation: base/math/intrin/x86/compv_math_utils_intrin_avx2.cxx:21-48 */
ORC_API int orc_sobel_g(const uint8_t* img, size_t w, size_t h, size_t stride, int which, int ks, int16_t* gx, int16_t* gy, uint16_t* g)
{
    const int16_t *vt, *hz;
    const int k = edge_tables(which, ks, &vt, &hz);
    if (!k) return 20006;
    int rc = orc_convlt1_8u16s16s(img, w, h, stride, vt, hz, (size_t)k, gx, 0);
    if (rc) return rc;
    rc = orc_convlt1_8u16s16s(img, w, h, stride, hz, vt, (size_t)k, gy, 0);
    if (rc) return rc;
    for (size_t y = 0; y < h; ++y) {
        for (size_t x = 0; x < w; ++x) {
            const int s = abs((int)gx[y * stride + x]) + abs((int)gy[y * stride + x]);
            g[y * stride + x] = (uint16_t)(s > 65535 ? 65535 : s);
        }
    }
    return 0;
}

/* a3: Sobel/Scharr/Prewitt detector = G -> gmax (>=1) -> u8(trunc(G * (255.f/gmax))) saturated.
 * core/features/edges/compv_core_feature_edge_dete.cxx:93,192-202; base/math/compv_math_utils.cxx:336-364 (+SSE2 leaf: cvtt, packs, packus) */
/* sse41_gmax_lanes != 0 reproduces a defect of the reference's x86 SSE4.1 leaf CompVMathUtilsMax_16u_Intrin_SSE41
 * (base/math/intrin/x86/compv_math_utils_intrin_sse41.cxx:55-63): its final horizontal reduction only folds lanes 0,1,2,4 of the
 * 8-lane running maximum, i.e. gmax is taken over the columns x with x%8 in {0,1,2,4}.  The plain C++ path (compv_math_utils.h:158-170)
 * and this function with sse41_gmax_lanes == 0 take the true maximum. */
ORC_API int orc_edge_normalized(const uint8_t* img, size_t w, size_t h, size_t stride, int which, int sse41_gmax_lanes, uint8_t* edges)
{
    const size_t n = stride * h;
    int16_t* gx = (int16_t*)malloc(n * 2); int16_t* gy = (int16_t*)malloc(n * 2); uint16_t* g = (uint16_t*)malloc(n * 2);
    if (!gx || !gy || !g) { free(gx); free(gy); free(g); return 20013; }
    int rc = orc_sobel_g(img, w, h, stride, which, 3, gx, gy, g);
    if (!rc) {
        unsigned gmax = 1;
        for (size_t y = 0; y < h; ++y) for (size_t x = 0; x < w; ++x) {
            if (sse41_gmax_lanes && !((0x17u >> (x & 7)) & 1u)) continue; /* lanes 0,1,2,4 */
            if (g[y * stride + x] > gmax) gmax = g[y * stride + x];
        }
        const float scale = 255.f / (float)gmax;
        for (size_t y = 0; y < h; ++y) {
            for (size_t x = 0; x < w; ++x) {
                const int v = (int)((float)g[y * stride + x] * scale);
                edges[y * stride + x] = (uint8_t)clampi(v, 0, 255);
            }
        }
    }
    free(gx); free(gy); free(g);
    return rc;
}

/* a5: Canny.  core/features/edges/compv_core_feature_canny_dete.cxx
 *   thresholds (compare-to-gradient / percent-of-mean, tHigh >= tLow+2)            :251-266
 *   NMS gather on unsuppressed g, integer direction tests, strict '>'             :566-598 (consts compv_core_feature_canny_dete.h:58-61)
 *   NMS apply (g=0 where gathered)                                                :414-460
 *   hysteresis: seeds g>tHigh in the interior, 8-neighbour growth through g>tLow,
 *               growth only FROM interior pixels                                  :600-680
 * thresholdType: 0 = compare to gradient, 1 = percent of mean.  The edge map is 0/255; rows 0,H-1 and cols 0,W-1 are 0. */
ORC_API int orc_canny(const uint8_t* img, size_t w, size_t h, size_t stride, float tLowF, float tHighF, int ks, int thresholdType, uint8_t* edges)
{
    if (!img || !edges || w < 3 || h < 3 || stride < w) return 20006;
    if (!(tLowF < tHighF)) return 20005;
    const size_t n = stride * h;
    int16_t* gx = (int16_t*)malloc(n * 2); int16_t* gy = (int16_t*)malloc(n * 2); uint16_t* g = (uint16_t*)malloc(n * 2);
    uint8_t* sup = (uint8_t*)calloc(n, 1);
    uint32_t* stack = (uint32_t*)malloc(n * sizeof(uint32_t) + 16);
    if (!gx || !gy || !g || !sup || !stack) { free(gx); free(gy); free(g); free(sup); free(stack); return 20013; }
    int rc = orc_sobel_g(img, w, h, stride, 3, ks == 3 ? 3 : 5, gx, gy, g);
    if (rc) goto done;
    {
        int tLow, tHigh;
        if (thresholdType == 1) {
            uint32_t sum = 0;
            for (size_t y = 0; y < h; ++y) for (size_t x = 0; x < w; ++x) sum += img[y * stride + x];
            int mean = (uint8_t)(sum / (uint32_t)(w * h));
            mean = clampi(mean, 1, 255);
            tLow = (uint16_t)(mean * tLowF);
            tHigh = (uint16_t)(mean * tHighF);
        }
        else {
            tLow = (uint16_t)(tLowF < 1.f ? 1.f : (tLowF > 65535.f ? 65535.f : tLowF));
            tHigh = (uint16_t)(tHighF < 1.f ? 1.f : (tHighF > 65535.f ? 65535.f : tHighF));
        }
        if (tLow < 1) tLow = 1;
        if (tHigh < tLow + 2) tHigh = tLow + 2;

        memset(edges, 0, n);
        const long s = (long)stride;
        for (size_t y = 1; y + 1 < h; ++y) {
            for (size_t x = 1; x + 1 < w; ++x) {
                const size_t i = y * stride + x;
                const int gc = g[i];
                if (gc <= tLow) continue;
                const int ax = abs((int)gx[i]), ay = abs((int)gy[i]) << 16;
                long d;
                if (ay < 27145 * ax) d = 1;                                          /* 0 deg   */
                else if (ay < 158217 * ax) d = ((gx[i] ^ gy[i]) < 0) ? (1 - s) : (1 + s); /* 45 / 135 */
                else d = s;                                                          /* 90 deg  */
                if (g[(long)i - d] > gc || g[(long)i + d] > gc) sup[i] = 1;
            }
        }
        for (size_t i = 0; i < n; ++i) if (sup[i]) g[i] = 0;

        for (size_t y = 1; y + 1 < h; ++y) {
            for (size_t x = 1; x + 1 < w; ++x) {
                if (!(g[y * stride + x] > tHigh) || edges[y * stride + x]) continue;
                size_t top = 0;
                edges[y * stride + x] = 0xff;
                stack[top++] = (uint32_t)((y << 16) | x);
                while (top) {
                    const uint32_t e = stack[--top];
                    const size_t cx = e & 0xffff, cy = e >> 16;
                    if (!(cy && cx && cy < h - 1 && cx < w - 1)) continue; /* growth only from interior pixels */
                    for (int dy = -1; dy <= 1; ++dy) {
                        for (int dx = -1; dx <= 1; ++dx) {
                            if (!dx && !dy) continue;
                            const size_t ny = cy + dy, nx = cx + dx, ni = ny * stride + nx;
                            if (g[ni] > tLow && !edges[ni]) { edges[ni] = 0xff; stack[top++] = (uint32_t)((ny << 16) | nx); }
                        }
                    }
                }
            }
        }
    }
done:
    free(gx); free(gy); free(g); free(sup); free(stack);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * a10 thresholding.  base/image/compv_image_threshold.cxx
 *   Otsu: histogram, sumA256[i] = i*h[i], fp32 between-class variance scan, strict '>' keeps the first maximum   :52-104, :349-366
 *   global: out = in > uint8(clip(T)+0.5) ? 255 : 0                                                               :118-180, :319-347
 *   adaptive: mean = fixed-point separable convolution with CompVKernel::mean taps (base/compv_kernel.cxx:12-25), out = lut[in - mean + 255]  :200-317
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_histogram_8u(const uint8_t* in, size_t w, size_t h, size_t stride, uint32_t* hist)
{
    if (!in || !hist || !w || !h || stride < w) return 20006;
    memset(hist, 0, 256 * sizeof(uint32_t));
    for (size_t y = 0; y < h; ++y) for (size_t x = 0; x < w; ++x) hist[in[y * stride + x]]++;
    return 0;
}

ORC_API int orc_threshold_global(const uint8_t* in, size_t w, size_t h, size_t stride, double threshold, uint8_t* out)
{
    if (!in || !out || !w || !h || stride < w || threshold < 0) return 20006;
    const double t = threshold > 255.0 ? 255.0 : threshold;
    const uint8_t T = (uint8_t)(t + 0.5);
    for (size_t y = 0; y < h; ++y) for (size_t x = 0; x < w; ++x) out[y * stride + x] = in[y * stride + x] > T ? 0xff : 0;
    return 0;
}

ORC_API int orc_threshold_otsu(const uint8_t* in, size_t w, size_t h, size_t stride, double* threshold, uint8_t* out)
{
    uint32_t hist[256];
    int rc = orc_histogram_8u(in, w, h, stride, hist);
    if (rc) return rc;
    uint32_t sum = 0;
    for (uint32_t i = 0; i < 256; ++i) sum += i * hist[i];
    const float sumf = (float)sum;
    const int N = (int)(w * h);
    volatile float sumB = 0.f, varMax = 0.f; /* volatile: every float operation rounds to fp32 individually */
    int q1 = 0, thr = 0;
    for (int i = 0; i < 256; ++i) {
        q1 += (int)hist[i];
        if (!q1) continue;
        const int q2 = N - q1;
        if (!q2) break;
        const float q1f = (float)q1, q2f = (float)q2;
        sumB += (float)(i * hist[i]);
        volatile float a = sumB / q1f, b = (sumf - sumB), c = b / q2f;
        volatile float mf = a - c;
        volatile float v1 = q1f * q2f, v2 = v1 * mf, varB = v2 * mf;
        if (varB > varMax) { varMax = varB; thr = i; }
    }
    *threshold = (double)thr;
    return out ? orc_threshold_global(in, w, h, stride, *threshold, out) : 0;
}

ORC_API int orc_threshold_adaptive(const uint8_t* in, size_t w, size_t h, size_t stride, size_t blockSize, double delta, double maxVal, int invert, uint8_t* out)
{
    if (!in || !out || !(blockSize & 1) || maxVal < 0 || blockSize > 1024) return 20006;
    uint16_t k[1024];
    const float vvv = 1.f / (float)blockSize;
    for (size_t i = 0; i < blockSize; ++i) k[i] = (uint16_t)(vvv * 0xffff);
    uint8_t* mean = (uint8_t*)calloc(stride * h, 1);
    if (!mean) return 20013;
    int rc = orc_convlt1_fxp_8u16u8u(in, w, h, stride, k, k, blockSize, mean, 0);
    if (!rc) {
        const double dc = delta < 0.0 ? 0.0 : (delta > 255.0 ? 255.0 : delta), mc = maxVal > 255.0 ? 255.0 : maxVal;
        const int deltaInt = (int)(dc + 0.5);
        const uint8_t mv = (uint8_t)(mc + 0.5);
        uint8_t lut[768];
        const size_t offCount = (size_t)(255 - deltaInt + 1);
        memset(lut, invert ? mv : 0, offCount);
        memset(lut + offCount, invert ? 0 : mv, 768 - offCount);
        for (size_t y = 0; y < h; ++y) for (size_t x = 0; x < w; ++x) out[y * stride + x] = lut[(int)in[y * stride + x] - (int)mean[y * stride + x] + 255];
    }
    free(mean);
    return rc;
}
