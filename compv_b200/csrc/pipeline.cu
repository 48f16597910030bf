// Host-buffer pipeline for the headline workload (BASELINE.json: Canny + HoughKHT): frames in host memory -> lines in host memory.
// The reference runs CompVEdgeDete::process then CompVHough::process on the CPU and the edge map travels through host memory between them
// (samples/hough_lines/main.cxx:59,106).  Here the edge map never leaves the device: chunks of frames are copied in on a copy stream while the
// previous chunk is in the Canny / KHT kernels, and only the detected lines come back.
#include "common.cuh"

using namespace cvb;

namespace {
struct PipeState {
	cudaStream_t sIn = nullptr, sCompute = nullptr;
	cudaEvent_t evIn[2] = { nullptr, nullptr }, evDone[2] = { nullptr, nullptr };
	DevBuf in[2], edges; // input chunks are double buffered; the edge maps of the whole batch stay resident for one KHT call
};
thread_local PipeState t_pipe;
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
		CVB_CUDA(cudaStreamCreateWithFlags(&st.sCompute, cudaStreamNonBlocking));
		for (int i = 0; i < 2; ++i) {
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
	const size_t nChunks = div_up(batch, chunk);
	auto h2d = [&](size_t c) -> int {
		const int slot = static_cast<int>(c & 1);
		const size_t f0 = c * chunk, nf = (f0 + chunk <= batch) ? chunk : (batch - f0);
		if (c >= 2) CVB_CUDA(cudaStreamWaitEvent(st.sIn, st.evDone[slot], 0));
		CVB_CUDA(cudaMemcpy2DAsync(st.in[slot].p, frameBytes, images + f0 * framePitch, framePitch, frameBytes, nf, cudaMemcpyHostToDevice, st.sIn));
		CVB_CUDA(cudaEventRecord(st.evIn[slot], st.sIn));
		return CVB200_S_OK;
	};
	CVB_CHECK(h2d(0));
	for (size_t c = 0; c < nChunks; ++c) {
		const int slot = static_cast<int>(c & 1);
		const size_t f0 = c * chunk, nf = (f0 + chunk <= batch) ? chunk : (batch - f0);
		if (c + 1 < nChunks) CVB_CHECK(h2d(c + 1));
		CVB_CUDA(cudaStreamWaitEvent(st.sCompute, st.evIn[slot], 0));
		CVB_CHECK(cvb200_edge_dete_process_dev(canny, st.in[slot].as<uint8_t>(), width, height, stride, st.edges.as<uint8_t>() + f0 * frameBytes, nf, frameBytes,
			reinterpret_cast<cvb200_stream_t>(st.sCompute)));
		CVB_CUDA(cudaEventRecord(st.evDone[slot], st.sCompute));
	}
	// the linking stage is latency bound with one warp per frame: its launch time does not depend on the number of frames, so the whole batch goes in one call
	CVB_CHECK(cvb200_hough_process_dev(hough, st.edges.as<uint8_t>(), width, height, stride, batch, frameBytes, lines, capacity, counts, reinterpret_cast<cvb200_stream_t>(st.sCompute)));
	CVB_CUDA(cudaStreamSynchronize(st.sCompute));
	return CVB200_S_OK;
}
