// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's Standard Hough Transform (row a6 of SURVEY.md section 8).
// Nothing under compv_b200/ links or calls this file; tests/, __graft_entry__.smoke() and bench.py's CPU legs are its only users.
//
// Follows /root/reference/core/features/hough/compv_core_feature_houghsht.cxx:
//   ctor :41-51 (theta in degrees -> radians in float), process :96-262, initCoords :320-348, acc_gather :350-481 (+ row kernel :594-605),
//   nms_gather :483-531 (+ row kernel :607-626, SSE2 kernel core/features/hough/intrin/x86/compv_core_feature_houghsht_intrin_sse2.cxx:16-50),
//   nms_apply :533-564 (+ row kernel :651-668).
// Pinned against the compiled reference (oracle/_ref) by tests/test_sht.py, for both the x86 SIMD path and the plain C++ path.
//
// Two behaviours of the reference that a reader would not expect are restated, selected by `x86Simd`:
//   x86Simd != 0 (what any x86 build runs): nms_gather's SSE2 row kernel walks the theta columns in groups of 4 starting at column 1 up to
//     (maxCols & -4) and the scalar tail that should finish the row is never entered (:524-530 passes an already offset pointer together with
//     colStart = consumed, so its loop range is empty): columns > 4*ceil(((maxCols & -4) - 1)/4) are never suppressed.
//   x86Simd == 0 (generic C++): the scalar kernel is started at column 0 (:526, consumed = 0) and reads column -1, i.e. the zeroed row padding of
//     the CompVMemZero accumulator (stride > cols for every theta count that is not a multiple of the allocator alignment); restated as "reads 0".
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <vector>
#include <algorithm>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {
struct Line { float rho, theta; size_t strength; }; // CompVHoughLine (compv_common.h:686-692)
}

// Returns 0 on success, the reference's error code otherwise. `count` receives the number of lines the reference would return (after maxLines);
// min(count, capacity) lines are written.
ORC_API int orc_hough_sht(const uint8_t* edges, size_t width, size_t height, size_t stride, float rho, float thetaDeg, size_t threshold, int maxLines,
	int x86Simd, void* linesOut, size_t capacity, size_t* count)
{
	if (!edges || !width || !height || stride < width || !count) return 20006;
	if (rho != 1.f || thetaDeg <= 0.f) return 20006; // newObj :310-314, set :71
	static const float kPi = 3.1415926535897932384626433f; // base/math/compv_math.cxx:27
	static const float kPiOver180 = kPi / 180.f; // :30
	const float fRho = rho * 1.f;
	const float fTheta = thetaDeg * kPiOver180; // ctor :44
	// initCoords :326-327 -- COMPV_MATH_ROUNDFU_2_NEAREST_INT(f) = (size_t)(f + 0.5) with the sum in double (compv_math.h:60)
	const size_t rows = static_cast<size_t>((static_cast<float>(((width + height) << 1) + 1) / fRho) + 0.5);
	const size_t cols = static_cast<size_t>((kPi / fTheta) + 0.5);
	if (!rows || !cols) return 20006;
	std::vector<int32_t> sinRho(cols), cosRho(cols);
	{
		float tt = 0.f;
		for (size_t t = 0; t < cols; ++t, tt += fTheta) { // :337-340, angle accumulated in float
			sinRho[t] = static_cast<int32_t>((sinf(tt) * fRho) * 65535.f);
			cosRho[t] = static_cast<int32_t>((cosf(tt) * fRho) * 65535.f);
		}
	}
	const int32_t barrier = static_cast<int32_t>(width + height); // :346
	std::vector<int32_t> acc(rows * cols, 0);
	// acc_gather :414-452 (directions are never supplied by the public API; the plain vote is taken for every theta)
	for (size_t j = 0; j < height; ++j) {
		const uint8_t* e = edges + j * stride;
		for (size_t i = 0; i < width; ++i) {
			if (!e[i]) continue;
			const int32_t row = static_cast<int32_t>(j), col = static_cast<int32_t>(i);
			for (size_t t = 0; t < cols; ++t) {
				const int32_t r = (col * cosRho[t] + sinRho[t] * row) >> 16; // :601
				acc[static_cast<size_t>(barrier - r) * cols + t]++;
			}
		}
	}
	// nms_gather :483-531
	std::vector<uint8_t> nms(rows * cols, 0);
	const int32_t thr = static_cast<int32_t>(threshold);
	const size_t maxCols = cols - 1;
	size_t cBegin, cEnd; // columns the suppression test is applied to
	if (x86Simd && maxCols >= 4) {
		const size_t m4 = maxCols & ~static_cast<size_t>(3);
		cBegin = 1;
		cEnd = 1 + 4 * ((m4 - 1 + 3) / 4); // sse2.cxx:24-26 (whole groups of 4 starting at column 1)
		if (cEnd > cols) cEnd = cols; // cannot happen (m4 <= cols - 1), kept as a guard
	}
	else {
		cBegin = 0; // :526 with consumed == 0
		cEnd = maxCols;
	}
	auto at = [&](size_t r, ptrdiff_t c) -> int32_t {
		if (c < 0 || c >= static_cast<ptrdiff_t>(cols)) return 0; // row padding of the zeroed accumulator
		return acc[r * cols + static_cast<size_t>(c)];
	};
	for (size_t r = 1; r + 1 < rows; ++r) { // :487-488
		for (size_t c = cBegin; c < cEnd; ++c) {
			const int32_t t = acc[r * cols + c];
			if (t <= thr) continue;
			const ptrdiff_t cc = static_cast<ptrdiff_t>(c);
			if (at(r, cc - 1) > t || at(r, cc + 1) > t || at(r - 1, cc - 1) > t || at(r - 1, cc) > t || at(r - 1, cc + 1) > t
				|| at(r + 1, cc - 1) > t || at(r + 1, cc) > t || at(r + 1, cc + 1) > t) {
				nms[r * cols + c] = 0xf;
			}
		}
	}
	// nms_apply :533-564: accumulator row-major order, rho = barrier - row, theta = col * fTheta
	std::vector<Line> lines;
	for (size_t r = 0; r < rows; ++r) {
		for (size_t c = 0; c < cols; ++c) {
			if (nms[r * cols + c]) continue;
			if (acc[r * cols + c] > thr) {
				Line l;
				l.rho = static_cast<float>(barrier - static_cast<int32_t>(r));
				l.theta = c * fTheta;
				l.strength = static_cast<size_t>(acc[r * cols + c]);
				lines.push_back(l);
			}
		}
	}
	// :241-247
	if (!lines.empty()) {
		std::sort(lines.begin(), lines.end(), [](const Line& a, const Line& b) -> bool { return a.strength > b.strength; });
		const size_t m = static_cast<size_t>(maxLines <= 0 ? 2147483647 : maxLines); // set :84
		if (lines.size() > m) lines.resize(m);
	}
	*count = lines.size();
	Line* out = static_cast<Line*>(linesOut);
	for (size_t i = 0; i < lines.size() && i < capacity; ++i) out[i] = lines[i];
	return 0;
}
