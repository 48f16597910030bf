// The object behind cvb200_hough_t: one struct for both line detectors (CompVHough::newObj picks the implementation by id, base/compv_features.cxx:176-191).
#pragma once
#include "common.cuh"

struct cvb200_hough {
	int id;
	float rho, theta;           // as handed to newObj
	size_t threshold;
	int maxLines;
	float clusterMinDeviation; int clusterMinSize; float kernelMinHeight;
	bool x86Simd;
	double lastGs;
	// KHT scratch (hough_kht.cu)
	int sortItemsHint = 8192;  // cells of the largest frame seen by the previous call (sizes the shared memory of the peak sort)
	cvb::DevBuf sortItems, sortLists, sortRanges, dLines, dCounts, tabs;
	cvb::HostBuf hTabs;
	size_t posCapEl = 0, strCapEl = 0, voteCapEl = 0;   // element capacities of the shared pools (grow-only)
	double tabRho = 0, tabTheta = 0, tabR = 0; size_t tabNRho = 0;
	int traceSlot = 0;
	bool bitsPrepared = false; // kht_prepare_bits was called: the next kht_enqueue finds the bitmap and the edge counts written by the producer of the edge map
	// SHT row-strip mode (hough_sht.cu): set around one call by the cvb200_hough_sht_* entry points
	size_t shtFullHeight = 0, shtYOffset = 0; int* shtExternalAcc = nullptr; int shtStage = 0; size_t* shtAccElems = nullptr;
	size_t pendBatch = 0, pendCapacity = 0; cudaStream_t pendStream = nullptr; // what kht_enqueue left for kht_finish
	cvb::DevBuf bits, poss, strings, strRev, clus, clusOrd, nClusStr, stack, kern, acc, rowCount, votes, frames, edgeCount, hostIn;
	cvb::HostBuf hFrames, hVotes, hCounts;
	// SHT scratch (hough_sht.cu)
	cvb::DevBuf shtTables, shtList, shtCursor, shtMask, shtPool, shtDesc;
	std::mutex mutex;
};

namespace cvb {
int kht_enqueue(cvb200_hough* h, const uint8_t* edges, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch, size_t capacity, cudaStream_t stream);
int kht_prepare_bits(cvb200_hough* h, size_t width, size_t height, size_t batch, cudaStream_t stream, unsigned int** bits, unsigned int** edgeCount, int* wordsPerRow);
int kht_finish(cvb200_hough* h, cvb200_hough_line_t* lines, size_t capacity, size_t* counts, bool* again);
int kht_process_dev(cvb200_hough* h, const uint8_t* edges, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch,
	cvb200_hough_line_t* lines, size_t capacity, size_t* counts, cudaStream_t stream);
int sht_process_dev(cvb200_hough* h, const uint8_t* edges, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch,
	cvb200_hough_line_t* lines, size_t capacity, size_t* counts, cudaStream_t stream);
}
