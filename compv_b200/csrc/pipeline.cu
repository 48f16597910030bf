// Host-buffer pipeline for the headline workload (BASELINE.json: Canny + HoughKHT): frames in host memory -> lines in host memory.
// The reference runs CompVEdgeDete::process then CompVHough::process on the CPU and the edge map travels through host memory between them
// (samples/hough_lines/main.cxx:59,106).  Here the edge map never leaves the device: chunks of frames are copied in on a copy stream while the
// previous chunk is in the Canny / KHT kernels, and only the detected lines come back.
#include "hough.cuh"

#include <climits>
#include <thread>

using namespace cvb;

namespace {
struct PipeState {
	cudaStream_t sIn = nullptr, sCompute[2] = { nullptr, nullptr };
	cudaEvent_t evIn[2] = { nullptr, nullptr }, evDone[2] = { nullptr, nullptr };
	DevBuf in[2], edges; // input chunks are double buffered; the edge maps of the whole batch stay resident for the Hough calls
	cvb200_hough* twin = nullptr; // second Hough object: the two halves of a batch are linked concurrently
};
thread_local PipeState t_pipe;

void copy_settings(const cvb200_hough* src, cvb200_hough* dst)
{
	dst->id = src->id; dst->rho = src->rho; dst->theta = src->theta; dst->threshold = src->threshold; dst->maxLines = src->maxLines;
	dst->clusterMinDeviation = src->clusterMinDeviation; dst->clusterMinSize = src->clusterMinSize; dst->kernelMinHeight = src->kernelMinHeight; dst->x86Simd = src->x86Simd;
}
}

extern "C" int cvb200_canny_kht_process_batch(cvb200_edge_dete_t* canny, cvb200_hough_t* hough, const uint8_t* images, size_t width, size_t height, size_t stride,
	size_t batch, size_t framePitch, cvb200_hough_line_t* lines, size_t capacity, size_t* counts)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(canny && hough && images && counts && width && height && stride >= width && (lines || !capacity), CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	CVB_REQUIRE(framePitch >= stride * height, CVB200_E_INVALID_PARAMETER);
	PipeState& st = t_pipe;
	if (!st.sIn) {
		CVB_CUDA(cudaStreamCreateWithFlags(&st.sIn, cudaStreamNonBlocking));
		for (int i = 0; i < 2; ++i) {
			CVB_CUDA(cudaStreamCreateWithFlags(&st.sCompute[i], cudaStreamNonBlocking));
			CVB_CUDA(cudaEventCreateWithFlags(&st.evIn[i], cudaEventDisableTiming));
			CVB_CUDA(cudaEventCreateWithFlags(&st.evDone[i], cudaEventDisableTiming));
		}
	}
	const size_t frameBytes = stride * height;
	size_t chunk = (16u << 20) / frameBytes; // ~16 MiB H2D chunks, overlapped with the Canny kernels of the previous chunk
	if (chunk < 1) chunk = 1;
	if (chunk > batch) chunk = batch;
	for (int i = 0; i < 2; ++i) CVB_CHECK(st.in[i].ensure(chunk * frameBytes));
	CVB_CHECK(st.edges.ensure(batch * frameBytes));
	// The linking stage of the Hough transform is latency bound with one warp per frame: its launch time does not depend on the number of frames, and two
	// launches on different streams overlap.  A large batch is therefore cut in two halves: the first half is handed to a helper thread (Hough on stream 0)
	// as soon as its edge maps exist, while this thread keeps uploading and edge-detecting the second half (stream 1) and then runs its Hough stage.
	size_t nHalves = (batch >= 64) ? 2 : 1;
	size_t half0 = (nHalves == 2) ? (batch / 2 / chunk) * chunk : batch; // a whole number of chunks
	if (half0 == 0 || half0 >= batch) { nHalves = 1; half0 = batch; }
	if (nHalves == 2 && !st.twin) { st.twin = new (std::nothrow) cvb200_hough(); CVB_REQUIRE(st.twin, CVB200_E_OUT_OF_MEMORY); st.twin->lastGs = 1.0; }
	const size_t nChunks = div_up(batch, chunk);
	auto h2d = [&](size_t c) -> int {
		const int slot = static_cast<int>(c & 1);
		const size_t f0 = c * chunk, nf = (f0 + chunk <= batch) ? chunk : (batch - f0);
		if (c >= 2) CVB_CUDA(cudaStreamWaitEvent(st.sIn, st.evDone[slot], 0));
		CVB_CUDA(cudaMemcpy2DAsync(st.in[slot].p, frameBytes, images + f0 * framePitch, framePitch, frameBytes, nf, cudaMemcpyHostToDevice, st.sIn));
		CVB_CUDA(cudaEventRecord(st.evIn[slot], st.sIn));
		return CVB200_S_OK;
	};
	int rcHelper = CVB200_S_OK;
	std::thread helper;
	CVB_CHECK(h2d(0));
	for (size_t c = 0; c < nChunks; ++c) {
		const int slot = static_cast<int>(c & 1);
		const size_t f0 = c * chunk, nf = (f0 + chunk <= batch) ? chunk : (batch - f0);
		const int which = (nHalves == 2 && f0 >= half0) ? 1 : 0;
		if (c + 1 < nChunks) { const int rc = h2d(c + 1); if (rc != CVB200_S_OK) { if (helper.joinable()) helper.join(); return rc; } }
		cudaStream_t sc = st.sCompute[which];
		int rc = CVB200_S_OK;
		if (cudaStreamWaitEvent(sc, st.evIn[slot], 0) != cudaSuccess) rc = CVB200_E_CUDA;
		if (rc == CVB200_S_OK) rc = cvb200_edge_dete_process_dev(canny, st.in[slot].as<uint8_t>(), width, height, stride, st.edges.as<uint8_t>() + f0 * frameBytes, nf, frameBytes,
			reinterpret_cast<cvb200_stream_t>(sc));
		if (rc == CVB200_S_OK && cudaEventRecord(st.evDone[slot], sc) != cudaSuccess) rc = CVB200_E_CUDA;
		if (rc != CVB200_S_OK) { if (helper.joinable()) helper.join(); return rc; }
		if (nHalves == 2 && f0 + nf == half0) { // the first half's edge maps are queued on stream 0: its Hough stage starts now, on the helper thread
			const int device = g_device.load();
			helper = std::thread([&, device]() {
				cudaSetDevice(device);
				rcHelper = cvb200_hough_process_dev(hough, st.edges.as<uint8_t>(), width, height, stride, half0, frameBytes, lines, capacity, counts,
					reinterpret_cast<cvb200_stream_t>(st.sCompute[0]));
			});
		}
	}
	int rc = CVB200_S_OK;
	if (nHalves == 2) {
		copy_settings(hough, st.twin);
		rc = cvb200_hough_process_dev(st.twin, st.edges.as<uint8_t>() + half0 * frameBytes, width, height, stride, batch - half0, frameBytes, lines ? lines + half0 * capacity : nullptr, capacity,
			counts + half0, reinterpret_cast<cvb200_stream_t>(st.sCompute[1]));
		if (helper.joinable()) helper.join();
		if (rc == CVB200_S_OK) rc = rcHelper;
		if (rc == CVB200_S_OK) hough->lastGs = st.twin->lastGs; // Gs of the last frame processed (houghkht.cxx:194-206)
	}
	else {
		rc = cvb200_hough_process_dev(hough, st.edges.as<uint8_t>(), width, height, stride, batch, frameBytes, lines, capacity, counts, reinterpret_cast<cvb200_stream_t>(st.sCompute[0]));
	}
	CVB_CHECK(rc);
	CVB_CUDA(cudaStreamSynchronize(st.sCompute[0]));
	CVB_CUDA(cudaStreamSynchronize(st.sCompute[1]));
	return CVB200_S_OK;
}
