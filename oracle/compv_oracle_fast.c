/*
 * TEST INFRASTRUCTURE ONLY (see compv_oracle.c).  CPU restatement of the FAST9/FAST12 detector.
 * Reference: core/features/fast/compv_core_feature_fast_dete.cxx
 *   circle offsets                                                  :221-238
 *   per-pixel strength (CompVFastDataRow_C)                         :658-771
 *   NMS: suppressed when any 8-neighbour >= own strength            :773-831
 *   point list, raster order, strength + threshold - 1              :490-585
 * Pinned against the compiled reference in tests/test_oracle_vs_ref.py (points with and without NMS, FAST9 and FAST12).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct { float x, y, strength, orient; int32_t level; float size; } orc_interest_point;

static const int kDx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
static const int kDy[16] = { -3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3 };

/* strengths[y*stride+x]: 0 outside [3,W-3)x[3,H-3).  Definition: with darker = sat(p-t), brighter = sat(p+t), the strength is the maximum over the
 * 16 arcs of N contiguous circle pixels that are ALL strictly darker (only considered when at least N of the 16 are darker) of the minimum
 * (darker - c) in the arc; otherwise the same with brighter (c - brighter) when at least N are brighter; 0 when no arc qualifies. */
ORC_API int orc_fast_scores(const uint8_t* img, size_t w, size_t h, size_t stride, int N, int threshold, uint8_t* strengths)
{
    if (!img || !strengths || w < 4 || h < 4 || stride < w || (N != 9 && N != 12)) return 20006;
    memset(strengths, 0, stride * h);
    for (size_t y = 3; y + 3 < h; ++y) {
        for (size_t x = 3; x + 3 < w; ++x) {
            const int p = img[y * stride + x];
            const int br = p + threshold > 255 ? 255 : p + threshold;
            const int dk = p - threshold < 0 ? 0 : p - threshold;
            int c[16], nd = 0, nb = 0;
            for (int k = 0; k < 16; ++k) {
                c[k] = img[(y + kDy[k]) * stride + (x + kDx[k])];
                nd += c[k] < dk;
                nb += c[k] > br;
            }
            int best = 0;
            const int dark = nd >= N;
            if (dark || nb >= N) {
                for (int s = 0; s < 16; ++s) {
                    int mn = 255, ok = 1;
                    for (int k = 0; k < N && ok; ++k) {
                        const int v = c[(s + k) & 15];
                        const int d = dark ? (dk - v) : (v - br);
                        if (d <= 0) ok = 0; else if (d < mn) mn = d;
                    }
                    if (ok && mn > best) best = mn;
                }
            }
            strengths[y * stride + x] = (uint8_t)best;
        }
    }
    return 0;
}

/* Points in raster order.  Returns the number found through *count; writes at most `capacity`. */
ORC_API int orc_fast_detect(const uint8_t* img, size_t w, size_t h, size_t stride, int N, int threshold, int nms, orc_interest_point* pts, size_t capacity, size_t* count)
{
    uint8_t* s = (uint8_t*)malloc(stride * h);
    uint8_t* sup = (uint8_t*)calloc(stride * h, 1);
    if (!s || !sup) { free(s); free(sup); return 20013; }
    int rc = orc_fast_scores(img, w, h, stride, N, threshold, s);
    size_t n = 0;
    if (!rc) {
        if (nms) {
            for (size_t y = 3; y + 3 < h; ++y) {
                for (size_t x = 3; x + 3 < w; ++x) {
                    const int v = s[y * stride + x];
                    if (!v) continue;
                    for (int dy = -1; dy <= 1 && !sup[y * stride + x]; ++dy)
                        for (int dx = -1; dx <= 1; ++dx)
                            if ((dx || dy) && s[(y + dy) * stride + (x + dx)] >= v) { sup[y * stride + x] = 1; break; }
                }
            }
        }
        for (size_t y = 3; y + 3 < h; ++y) {
            for (size_t x = 0; x < w; ++x) {
                if (s[y * stride + x] && !sup[y * stride + x]) {
                    if (n < capacity) {
                        orc_interest_point* p = &pts[n];
                        p->x = (float)x; p->y = (float)y; p->strength = (float)(s[y * stride + x] + threshold - 1);
                        p->orient = -1.f; p->level = 0; p->size = 0.f;
                    }
                    ++n;
                }
            }
        }
    }
    *count = n;
    free(s); free(sup);
    return rc;
}
