// The objects behind cvb200_ccl_t / cvb200_ccl_result_t: one pair for both labelers (CompVConnectedComponentLabeling::newObj picks by id, base/compv_ccl.cxx:69-97).
#pragma once
#include "common.cuh"

#include <new>
#include <vector>

struct cvb200_ccl_result {
	int id = 0;                // CVB200_PLSL_ID or CVB200_LMSER_ID
	size_t width = 0, height = 0;
	int32_t na = 0;            // PLSL: labels; LMSER: regions (labelsCount, lmser_result.cxx:30-33)
	// PLSL: the LEA in CSR form
	std::vector<uint32_t> rowOffsets;
	std::vector<cvb200_ccl_range_t> ranges;
	// LMSER: regions
	std::vector<int32_t> regionSizes;
	std::vector<cvb200_rect16_t> regionBoxes;
	std::vector<int16_t> regionPoints; // (x, y) pairs, regions back to back
};

struct cvb200_ccl {
	int id;
	int type;
	bool sortSegments;
	int connectivity;
	int delta; double minArea, maxArea, maxVariation, minDiversity; // compv_ccl.h:23-28, 229-236
	// PLSL scratch (ccl_lsl.cu)
	cvb::DevBuf fg, spre, rowCnt, rowOff, frames, segStart, segEnd, ov, label, eq, a, ranges, hostIn;
	cvb::HostBuf hFrames, hRowOff, hRanges;
	// LMSER scratch (ccl_lmser.cu)
	cvb::DevBuf mUf, mStamp, mPending, mCompSize, mAddSize, mOwnCnt, mTopNode, mPixNode, mOrder, mAbsorbed, mNodeRoot, mNodeParent, mNodeArea, mNodeOwn, mNodeLevel,
		mChild, mSister, mOff, mCursor, mOwnCursor, mVar, mFlags, mDfsPix, mCounters, mRegions, mOutOff, mPoints, mBoxes;
	std::mutex mutex;
};

namespace cvb {
int mser_process_dev(cvb200_ccl* c, const uint8_t* img, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch, cvb200_ccl_result_t** results, cudaStream_t stream);
}
