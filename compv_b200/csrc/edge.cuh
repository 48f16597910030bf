// The object behind cvb200_edge_dete_t (Canny / Sobel / Scharr / Prewitt; CompVEdgeDete::newObj picks by id, base/compv_features.cxx:146-161).
// Caches its scratch like the reference objects do (canny_dete.cxx:133-147).
#pragma once
#include "common.cuh"

namespace cvb {
// kernel tables: base/include/compv/base/compv_features.h:124-133
struct EdgeTaps {
	int16_t vt[5];
	int16_t hz[5];
	int ks;
};
struct BlurTaps {
	float k[7];
	int ks; // 0 = no blur; 3, 5 or 7
};
}

struct cvb200_edge_dete {
	int id;
	float tLow, tHigh;
	int thresholdType;
	cvb::EdgeTaps taps;
	cvb::BlurTaps blur;
	cvb::DevBuf dirty;      // hysteresis: per-tile epoch + two tile lists + per-round counters
	cvb::DevBuf counters;   // per-frame gmax / sums / thresholds
	cvb::HostBuf hostFlag;
	cvb::DevBuf hostIn, hostOut; // staging for the host-buffer entry point
	bool gmaxLanes;
	bool genericKernel;     // CVB200_EDGE_SET_BOOL_GENERIC_KERNEL: force the generic front kernel (tests)
	int hystRounds = 8;     // list-driven hysteresis rounds issued per call after round 0 (raised when a call did not converge)
	// pipeline.cu: the finalize pass also writes the KHT linking bitmap (kht_prepare_bits) and the per-frame edge counts, saving the KHT stage a pass over the edge map
	unsigned int* khtBits = nullptr; unsigned int* khtEdgeCount = nullptr; int khtWW = 0;
	int stages = 0; unsigned int* externalGmax = nullptr; // row-strip mode: set around one call by cvb200_edge_dete_process_stages_dev
	cudaStream_t pendStream = nullptr; bool pendCheck = false; // what edge_enqueue left for edge_finish
	std::mutex mutex;
};

namespace cvb {
int edge_enqueue(cvb200_edge_dete* d, const uint8_t* image, size_t width, size_t height, size_t stride, uint8_t* edges, size_t batch, size_t framePitch, cudaStream_t stream);
int edge_finish(cvb200_edge_dete* d, bool* again);
}
