/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's Parallel Light Speed Labeling (row a11 of SURVEY.md section 8).
 * Nothing under compv_b200/ links or calls this file; tests/, __graft_entry__.smoke() and bench.py's CPU legs are its only users.
 *
 * Follows /root/reference/core/ccl/compv_core_ccl_lsl.cxx: step 1 (relative labels ER, run-length table RLC, ner) :153-214,
 * step 2.0 (range of previous-row relative labels each segment touches, 8-connectivity) :341-374, step 2.1 (absolute labels + equivalence table EQ)
 * :430-479, step 4 (EQ -> ancestors A, final numbering) :481-505, build_LEA :545-576, process :579-751; and the result class
 * core/ccl/compv_core_ccl_lsl_result.cxx: debugFlatten :51-98, boundingBoxes :136-185.
 * Pinned against the compiled reference (oracle/_ref) by tests/test_lsl.py.
 *
 * NOTE (restated, not corrected): step 2.1 writes the new minimum into EQ[eak] of the *segment's* label, not into the root of its class; when a
 * previous-row segment had already been attached to another class during the same row the older class keeps its own root.  On noise-like inputs the
 * reference therefore returns a few more labels than there are 8-connected components (39 of 300 random frames in scripts/lsl_probe.py).  The
 * restatement keeps the exact sequence of EQ reads and writes so that the labels match the reference's, including in those cases. */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct { int32_t a; int16_t start, end; } orc_lsl_range_t; /* compv_ccl_range_t (compv_core_ccl_lsl_result.h:32-36) */

/* labels: height x width int32 (strideless) or NULL; boxes: {left, top, right, bottom} int16 per label or NULL (boxCap labels);
 * rowOffsets: height+1 entries or NULL; ranges: the LEA rows concatenated, up to rangeCap, or NULL. Returns 0, or the reference's error code. */
ORC_API int orc_ccl_lsl(const uint8_t* img, size_t width, size_t height, size_t stride, int32_t* labels, int32_t* naOut, int16_t* boxes, size_t boxCap,
	uint32_t* rowOffsets, orc_lsl_range_t* ranges, size_t rangeCap, size_t* rangeCount)
{
	if (!img || !width || !height || stride < width || width > 32767 || height > 32767) return 20006;
	const int16_t w = (int16_t)width;
	int16_t* ER = (int16_t*)malloc(width * height * sizeof(int16_t));
	int16_t* RLC = (int16_t*)malloc((width + 1) * height * sizeof(int16_t));
	int16_t* ner = (int16_t*)malloc(height * sizeof(int16_t));
	if (!ER || !RLC || !ner) { free(ER); free(RLC); free(ner); return 20013; }
	int16_t ner_max = 0; int32_t ner_sum = 0;
	/* step 1 :153-214 */
	for (size_t j = 0; j < height; ++j) {
		const uint8_t* Xi = img + j * stride;
		int16_t* ERi = ER + j * width;
		int16_t* RLCi = RLC + j * (width + 1);
		int16_t er = (Xi[0] & 1);
		ERi[0] = er;
		for (int16_t i = 1; i < w; ++i) { er += ((Xi[i - 1] ^ Xi[i]) & 1); ERi[i] = er; }
		er += (Xi[w - 1] & 1);
		ner[j] = er; ner_sum += er; if (ner_max < er) ner_max = er;
		er = (Xi[0] & 1);
		RLCi[0] = 0;
		for (int16_t i = 1; i < w; ++i) if (ERi[i - 1] != ERi[i]) RLCi[er++] = i;
		RLCi[er] = w - ((Xi[w - 1] & 1) ^ 1);
	}
	if (naOut) *naOut = 0;
	if (rangeCount) *rangeCount = 0;
	if (labels) memset(labels, 0, width * height * sizeof(int32_t));
	if (rowOffsets) memset(rowOffsets, 0, (height + 1) * sizeof(uint32_t));
	if (!ner_max) { free(ER); free(RLC); free(ner); return 0; } /* black image :676-680 */
	const size_t eraStride = (size_t)ner_max + 1;
	int32_t* ERA = (int32_t*)calloc(eraStride * height, sizeof(int32_t));
	int32_t* EQ = (int32_t*)malloc(((size_t)ner_sum + 1) * sizeof(int32_t));
	if (!ERA || !EQ) { free(ER); free(RLC); free(ner); free(ERA); free(EQ); return 20013; }
	for (int32_t i = 0; i < ner_sum; ++i) EQ[i] = i; /* build_EQ :531-541 */
	/* step 2.0 :341-374 */
	const int16_t wminus1 = (int16_t)(w - 1);
	for (size_t j = 1; j < height; ++j) {
		const int16_t* ERiminus1 = ER + (j - 1) * width;
		const int16_t* RLCi = RLC + j * (width + 1);
		int32_t* ERAi = ERA + j * eraStride;
		for (int16_t er = 1; er < ner[j]; er += 2) {
			int16_t j0 = RLCi[er - 1], j1 = (int16_t)(RLCi[er] - 1);
			j0 -= (j0 > 0); j1 += (j1 < wminus1);
			int16_t er0 = ERiminus1[j0], er1 = ERiminus1[j1];
			er0 += ((er0 & 1) ^ 1); er1 -= ((er1 & 1) ^ 1);
			ERAi[er] = (er1 >= er0) ? (er0 | er1 << 16) : 0;
		}
	}
	/* step 2.1 :430-479 */
	int32_t nea = 0;
	for (int16_t er = 1; er < ner[0]; er += 2) ERA[er] = ++nea;
	for (size_t j = 1; j < height; ++j) {
		int32_t* ERAi = ERA + j * eraStride;
		const int32_t* ERAiminus1 = ERA + (j - 1) * eraStride;
		for (int16_t er = 1; er < ner[j]; er += 2) {
			if (ERAi[er]) {
				const int32_t er0 = ERAi[er] & 0xffff, er1 = (ERAi[er] >> 16) & 0xffff;
				int32_t ea = ERAiminus1[er0];
				int32_t a = EQ[ea];
				for (int32_t erk = er0 + 2; erk <= er1; erk += 2) {
					const int32_t eak = ERAiminus1[erk];
					const int32_t ak = EQ[eak];
					if (a < ak) EQ[eak] = a;
					else { a = ak; EQ[ea] = a; ea = eak; }
				}
				ERAi[er] = a;
			}
			else ERAi[er] = ++nea;
		}
	}
	/* step 4 :481-505 (the 4-way unrolled loop and its tail visit ea = 1 .. nea in order) */
	int32_t* A = (int32_t*)malloc(((size_t)nea + 1) * sizeof(int32_t));
	if (!A) { free(ER); free(RLC); free(ner); free(ERA); free(EQ); return 20013; }
	A[0] = 0;
	int32_t na = 0;
	for (int32_t ea = 1; ea <= nea; ++ea) { const int32_t a1 = EQ[ea]; A[ea] = (a1 != ea) ? A[a1] : ++na; }
	if (naOut) *naOut = na;
	/* build_LEA :545-576 + debugFlatten / boundingBoxes of the result */
	if (boxes) for (int32_t k = 0; k < na && (size_t)k < boxCap; ++k) { boxes[4 * k] = (int16_t)width; boxes[4 * k + 1] = (int16_t)height; boxes[4 * k + 2] = 0; boxes[4 * k + 3] = 0; }
	size_t nr = 0;
	for (size_t j = 0; j < height; ++j) {
		const int32_t* ERAi = ERA + j * eraStride;
		const int16_t* RLCi = RLC + j * (width + 1);
		if (rowOffsets) rowOffsets[j] = (uint32_t)nr;
		for (int16_t er = 1; er < ner[j]; er += 2) {
			const int32_t a = A[ERAi[er]];
			if (!a) continue;
			const int16_t s = RLCi[er - 1], e = RLCi[er];
			if (ranges && nr < rangeCap) { ranges[nr].a = a; ranges[nr].start = s; ranges[nr].end = e; }
			++nr;
			if (labels) for (int16_t x = s; x < e; ++x) labels[j * width + (size_t)x] = a;
			if (boxes && (size_t)(a - 1) < boxCap) {
				int16_t* bb = boxes + 4 * (a - 1);
				if (s < bb[0]) bb[0] = s;
				if ((int16_t)j < bb[1]) bb[1] = (int16_t)j;
				if (e > bb[2]) bb[2] = e;
				bb[3] = (int16_t)j;
			}
		}
	}
	if (rowOffsets) rowOffsets[height] = (uint32_t)nr;
	if (rangeCount) *rangeCount = nr;
	free(ER); free(RLC); free(ner); free(ERA); free(EQ); free(A);
	return 0;
}
