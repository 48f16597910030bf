/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's mathematical morphology (first "next" row).
 * Nothing under compv_b200/ links or calls this file; tests/, __graft_entry__.smoke() and bench.py's CPU legs are its only users.
 *
 * Follows /root/reference/base/math/compv_math_morph.cxx: process :95-126, basicOper :128-240 (interior = min / max over the non-zero cells of the
 * structuring element, then addBordersVt :585-629 and addBordersHz :631-694), openCloseOper :242-337 (two basic operations, the second on the first's
 * full output), buildStructuringElementGeneric :513-583 (RECT / CROSS / DIAMOND).  Pinned against the compiled reference by tests/test_morph.py.
 * Restated as is: the vertical border is (strelHeight + 1) / 2 rows high (:596), i.e. one row MORE than the half height for odd sizes -- for a 3x3
 * element rows 0-1 and H-2..H-1 are copies of the input (REPLICATE) or zero, although rows 1 and H-2 were computed. */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* type: 0 RECT, 1 DIAMOND, 2 CROSS (COMPV_MATH_MORPH_STREL_TYPE, compv_common.h:402-406); strel is height x width, tightly packed */
ORC_API int orc_morph_strel(uint8_t* strel, size_t width, size_t height, int type)
{
	if (!strel || !width || !height) return 20006;
	memset(strel, 0, width * height);
	if (type == 0) memset(strel, 255, width * height);
	else if (type == 2) {
		for (size_t i = 0; i < width; ++i) strel[(height >> 1) * width + i] = 255;
		for (size_t j = 0; j < height; ++j) strel[j * width + (width >> 1)] = 255;
	}
	else if (type == 1) { /* :552-575; cells outside the element (non square sizes) are skipped here, the reference writes through them */
		const size_t hd = height >> 1, wd = width >> 1;
		ptrdiff_t col = (ptrdiff_t)wd; size_t row = 0, count = 1;
		for (size_t j = 0; j < hd; ++j, count += 2, ++row, --col)
			for (size_t i = 0; i < count; ++i) if (col + (ptrdiff_t)i >= 0 && col + (ptrdiff_t)i < (ptrdiff_t)width && row < height) strel[row * width + (size_t)(col + (ptrdiff_t)i)] = 255;
		for (size_t j = 0; j <= hd; ++j, count -= 2, ++row, ++col)
			for (size_t i = 0; i < count; ++i) if (col + (ptrdiff_t)i >= 0 && col + (ptrdiff_t)i < (ptrdiff_t)width && row < height) strel[row * width + (size_t)(col + (ptrdiff_t)i)] = 255;
	}
	else return 20001;
	return 0;
}

/* one basic operation (erode = min, dilate = max) with the reference's border handling; border: 0 ZERO, 1 IGNORE, 2 REPLICATE */
static int basic(const uint8_t* in, size_t W, size_t H, size_t stride, const uint8_t* strel, size_t sw, size_t sh, size_t sstride, uint8_t* out, int erode, int border)
{
	const size_t rw = sw >> 1, rh = sh >> 1;
	size_t count = 0;
	for (size_t j = 0; j < sh; ++j) for (size_t i = 0; i < sw; ++i) if (strel[j * sstride + i]) ++count;
	if (!count) return 20006; /* :483 */
	const size_t opH = H - (rh << 1), opW = W - (rw << 1);
	for (size_t y = 0; y < opH; ++y) {
		for (size_t x = 0; x < opW; ++x) {
			int v = erode ? 255 : 0;
			for (size_t j = 0; j < sh; ++j) for (size_t i = 0; i < sw; ++i) {
				if (!strel[j * sstride + i]) continue;
				const int s = in[(y + j) * stride + x + i];
				v = erode ? (s < v ? s : v) : (s > v ? s : v);
			}
			out[(y + rh) * stride + x + rw] = (uint8_t)v;
		}
	}
	/* addBordersVt :585-629 */
	const size_t bh = (sh + 1) >> 1;
	for (size_t k = 0; k < bh && k < H; ++k) {
		if (border == 0) { memset(out + k * stride, 0, W); memset(out + (H - bh + k) * stride, 0, W); }
		else if (border == 2) { memcpy(out + k * stride, in + k * stride, W); memcpy(out + (H - bh + k) * stride, in + (H - bh + k) * stride, W); }
	}
	/* addBordersHz :631-694 */
	for (size_t y = 0; y < H; ++y) {
		for (size_t c = 0; c < rw; ++c) {
			if (border == 0) { out[y * stride + c] = 0; out[y * stride + W - rw + c] = 0; }
			else if (border == 2) { out[y * stride + c] = in[y * stride + c]; out[y * stride + W - rw + c] = in[y * stride + W - rw + c]; }
		}
	}
	return 0;
}

/* op: 0 ERODE, 1 DILATE, 2 OPEN, 3 CLOSE (COMPV_MATH_MORPH_OP_TYPE, compv_common.h:410-419); out must not alias in. With border IGNORE the cells the
 * reference leaves untouched are whatever `out` held (callers pre-fill it). */
ORC_API int orc_morph_process(const uint8_t* in, size_t W, size_t H, size_t stride, const uint8_t* strel, size_t sw, size_t sh, size_t sstride, uint8_t* out, int op, int border)
{
	if (!in || !out || !strel || !W || !H || stride < W || !sw || !sh || sstride < sw || W < sw || H < sh) return 20006; /* :131-137 */
	if (border < 0 || border > 2) return 20001;
	if (op == 0 || op == 1) return basic(in, W, H, stride, strel, sw, sh, sstride, out, op == 0, border);
	if (op == 2 || op == 3) {
		uint8_t* tmp = (uint8_t*)malloc(stride * H);
		if (!tmp) return 20013;
		memcpy(tmp, out, stride * H); /* IGNORE: the intermediate starts from the same bytes the output holds */
		int r = basic(in, W, H, stride, strel, sw, sh, sstride, tmp, op == 2, border);
		if (!r) r = basic(tmp, W, H, stride, strel, sw, sh, sstride, out, op != 2, border);
		free(tmp);
		return r;
	}
	return 20001; /* :119-122 */
}
