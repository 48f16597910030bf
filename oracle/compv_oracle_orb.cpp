// TEST INFRASTRUCTURE ONLY -- restatement of the reference's ORB DETECTOR: CompVCornerDeteORB::process / processLevelAt
// (core/features/orb/compv_core_feature_orb_dete.cxx:148-358) with its defaults (:35-44): 8 pyramid levels of scale 0.83^level, every level scaled from the ORIGINAL
// image with the 8-bit fixed-point bilinear kernel (base/image/compv_image_scale_pyramid.cxx:37-45,150-155; base/image/compv_image_scale_bilinear.cxx:50-86,163-176),
// FAST9 + NMS per level (its own maxFeatures 2000, fast_dete.cxx:79,417-420), per-level quota (orb_dete.cxx:312-320), CompVInterestPoint::selectBest /
// eraseTooCloseToBorder (base/include/compv/base/compv_common.h:641-663), intensity-centroid orientation from the circular patch moments
// (base/compv_patch.cxx:62-140, abscissas :199-203) and the scaling of the coordinates back to level 0 (orb_dete.cxx:330-355).
// C++ because selectBest is libstdc++'s std::nth_element + std::partition (the tie behaviour at the cut is the library's).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

typedef struct { float x, y, strength, orient; int32_t level; float size; } orc_interest_point;
extern "C" int orc_fast_detect(const uint8_t* img, size_t w, size_t h, size_t stride, int N, int threshold, int nms, orc_interest_point* pts, size_t capacity, size_t* count);

static void select_best(std::vector<orc_interest_point>& v, size_t max) // compv_common.h:641-656
{
	if (max > 1) {
		std::nth_element(v.begin(), v.begin() + max, v.end(), [](const orc_interest_point& i, const orc_interest_point& j) { return i.strength > j.strength; });
		const float pivot = v.at(max - 1).strength;
		v.resize(std::partition(v.begin() + max, v.end(), [pivot](orc_interest_point i) { return i.strength >= pivot; }) - v.begin());
	}
}

// compv_image_scale_bilinear.cxx:50-86 with the factors of :163-176
ORC_API int orc_scale_bilinear(const uint8_t* in, size_t inW, size_t inH, size_t inStride, uint8_t* out, size_t outW, size_t outH, size_t outStride)
{
	if (!in || !out || !inW || !inH || !outW || !outH) return 20006;
	const float fsx = static_cast<float>(inW) / outW, fsy = static_cast<float>(inH) / outH;
	const size_t sfx = static_cast<size_t>(static_cast<long>(fsx * 256.f)), sfy = static_cast<size_t>(static_cast<long>(fsy * 256.f));
	size_t oy = 0;
	for (size_t j = 0; j < outH; ++j, oy += sfy) {
		const size_t ny = oy >> 8;
		const uint8_t* p = in + ny * inStride;
		const unsigned int y0 = oy & 255, y1 = 255 - y0;
		size_t x = 0;
		for (size_t i = 0; i < outW; ++i, x += sfx) {
			const size_t nx = x >> 8;
			const unsigned int n0 = p[nx], n1 = p[nx + 1], n2 = p[nx + inStride], n3 = p[nx + 1 + inStride];
			const unsigned int x0 = x & 255, x1 = 255 - x0;
			out[j * outStride + i] = static_cast<uint8_t>((y1 * ((n0 * x1) + (n1 * x0)) >> 16) + (y0 * ((n2 * x1) + (n3 * x0)) >> 16));
		}
	}
	return 0;
}

ORC_API int orc_orb_detect(const uint8_t* img, size_t w, size_t h, size_t stride, int fastThreshold, int nms, int fastN, int maxFeatures, orc_interest_point* out, size_t capacity, size_t* count)
{
	if (!img || !count || w < 4 || h < 4 || stride < w) return 20006;
	const int levels = 8, patchDiameter = 31, radius = patchDiameter >> 1;
	const float sf0 = 0.83f;
	float sfTab[8]; float sfs = 1.f;          // scale_pyramid.cxx:37-45
	sfTab[0] = 1.f;
	{ float s = sf0; for (int l = 1; l < levels; ++l, s *= sf0) { sfTab[l] = s; sfs += s; } }
	int16_t maxAbs[64];
	for (int i = 0; i <= radius; ++i) maxAbs[i] = static_cast<int16_t>(sqrt(static_cast<double>(radius * radius - (i * i)))); // compv_patch.cxx:199-203
	size_t n = 0;
	std::vector<uint8_t> lvl;
	for (int level = 0; level < levels; ++level) {
		const float sf = sfTab[level];
		const uint8_t* p = img; size_t lw = w, lh = h, ls = stride;
		if (level) {
			lw = static_cast<size_t>(w * sf); lh = static_cast<size_t>(h * sf); ls = lw;
			if (lw < 4 || lh < 4) continue;
			lvl.assign(ls * lh, 0);
			orc_scale_bilinear(img, w, h, stride, lvl.data(), lw, lh, ls);
			p = lvl.data();
		}
		std::vector<orc_interest_point> pts(lw * lh);
		size_t cnt = 0;
		int rc = orc_fast_detect(p, lw, lh, ls, fastN, fastThreshold, nms, pts.data(), pts.size(), &cnt);
		if (rc) return rc;
		pts.resize(cnt);
		if (pts.size() > 2000) select_best(pts, 2000);                     // the internal FAST detector's own default (fast_dete.cxx:79,417-420)
		if (maxFeatures > 0 && !pts.empty()) {                            // orb_dete.cxx:312-320
			const float nf = ((maxFeatures / sfs) * sf);
			int32_t mf = static_cast<int32_t>(nf + 0.5);
			mf = mf < 10 ? 10 : mf;
			if (pts.size() > static_cast<size_t>(mf)) select_best(pts, static_cast<size_t>(mf));
		}
		{                                                                 // eraseTooCloseToBorder, border (31 + 5) >> 1 = 18
			const float fw = static_cast<float>(lw), fh = static_cast<float>(lh), b = static_cast<float>((patchDiameter + 5) >> 1);
			pts.erase(std::remove_if(pts.begin(), pts.end(), [&](const orc_interest_point& q) { return (q.x < b || (q.x + b) >= fw || (q.y < b) || (q.y + b) >= fh); }), pts.end());
		}
		const float sfi = 1.f / sf, patchSize = patchDiameter / sf;
		for (orc_interest_point& q : pts) {
			q.level = level; q.size = patchSize;
			const int cx = static_cast<int>(static_cast<int>(q.x >= 0.0 ? (q.x + 0.5) : (q.x - 0.5))), cy = static_cast<int>(static_cast<int>(q.y >= 0.0 ? (q.y + 0.5) : (q.y - 0.5)));
			int m10 = 0, m01 = 0;                                        // compv_patch.cxx:96-140: the whole disc (never close to the border after the erase above)
			for (int j = -radius; j <= radius; ++j) {
				const int dX = maxAbs[j < 0 ? -j : j];
				const uint8_t* row = p + static_cast<size_t>(cy + j) * ls + cx;
				for (int i = -dX; i <= dX; ++i) { m10 += i * row[i]; m01 += j * row[i]; }
			}
			const float rad = std::atan2(static_cast<float>(m01), static_cast<float>(m10));
			q.orient = static_cast<float>(rad * (180.f / 3.1415926535897932384626433f)); // COMPV_MATH_RADIAN_TO_DEGREE_FLOAT, kfMathTrig180OverPi = 180.f / kfMathTrigPi (compv_math.cxx:31)
			if (q.orient < 0) q.orient += 360;
			if (level != 0) { q.x *= sfi; q.y *= sfi; }
			if (out && n < capacity) out[n] = q;
			++n;
		}
	}
	*count = n;
	return 0;
}
