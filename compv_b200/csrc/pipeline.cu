// The headline workload (BASELINE.json: Canny + HoughKHT) as one pipelined call: frames (host or device memory) -> lines in host memory.
// The reference runs CompVEdgeDete::process then CompVHough::process per frame on the CPU and the edge map travels through host memory between them
// (samples/hough_lines/main.cxx:59,106).  Here the frames of a batch are cut into sub-batches that flow through a ring of SLOTS: every slot owns a stream,
// a private copy of the two detector objects (their scratch is per object, as in the reference) and its edge maps.  Nothing inside a slot talks to the host
// (device-side pool offsets, hysteresis rounds and peak ordering), so the host only enqueues; the sub-batches of neighbouring slots overlap on the GPU:
// the KHT linking kernel of one sub-batch (one warp per frame: a latency chain that leaves the SMs almost empty) runs next to the Canny and voting kernels
// of the others and, for host frames, next to the H2D copy of the next sub-batch.
#include "edge.cuh"
#include "hough.cuh"

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>

using namespace cvb;

namespace {

struct PipeSlot {
	cudaStream_t stream = nullptr;    // the stream of the current call: one of the two below
	cudaStream_t streamDev = nullptr, streamHost = nullptr; // device frames: earlier slots have the higher stream priority; host frames: later slots (see run_pipeline)
	cudaEvent_t evIn = nullptr;       // this slot's upload has landed
	cvb200_edge_dete canny;
	cvb200_hough hough;
	DevBuf in, edges, raw;
	size_t f0 = 0, nf = 0;            // frames in flight
	bool busy = false;
	const uint8_t* dIn = nullptr;     // where the sub-batch's frames are on the device
	size_t inPitch = 0;
};

struct PipeState {
	std::vector<PipeSlot*> slots;
	cudaStream_t sCopy = nullptr;
	cudaEvent_t evStart = nullptr;
};
thread_local PipeState t_pipe;

int env_int(const char* name, int dflt) { const char* v = getenv(name); return (v && *v) ? atoi(v) : dflt; }

void copy_params(cvb200_edge_dete& dst, const cvb200_edge_dete& src)
{
	dst.id = src.id; dst.tLow = src.tLow; dst.tHigh = src.tHigh; dst.thresholdType = src.thresholdType; dst.taps = src.taps; dst.blur = src.blur;
	dst.gmaxLanes = src.gmaxLanes; dst.genericKernel = src.genericKernel;
	if (dst.hystRounds < src.hystRounds) dst.hystRounds = src.hystRounds;
}
void copy_params(cvb200_hough& dst, const cvb200_hough& src)
{
	dst.id = src.id; dst.rho = src.rho; dst.theta = src.theta; dst.threshold = src.threshold; dst.maxLines = src.maxLines;
	dst.clusterMinDeviation = src.clusterMinDeviation; dst.clusterMinSize = src.clusterMinSize; dst.kernelMinHeight = src.kernelMinHeight; dst.x86Simd = src.x86Simd;
}

struct Job {
	cvb200_edge_dete* canny; cvb200_hough* hough;
	const uint8_t* images; bool onHost;
	size_t width, height, stride, batch, framePitch;
	cvb200_hough_line_t* lines; size_t capacity; size_t* counts;
	size_t frameBytes, sub;
	int pixelFormat = 0; size_t bpp = 1; // host frames in a camera format: uploaded as they are (luma-carrying plane only), made gray on the device (gray.cu)
};

int slot_enqueue(PipeState& st, PipeSlot& s, const Job& j, size_t f0, size_t nf)
{
	s.f0 = f0; s.nf = nf;
	if (j.onHost) {
		CVB_CHECK(s.in.ensure(j.sub * j.frameBytes));
		// uploads go through one copy stream in order; the copy may start as soon as this slot's previous kernels are done (finish() has synchronised them)
		if (j.bpp == 1) { // gray, or a planar / semi-planar YUV format: the Y plane is the gray image, the chroma planes never cross PCIe
			CVB_CUDA(cudaMemcpy2DAsync(s.in.p, j.frameBytes, j.images + f0 * j.framePitch, j.framePitch, j.frameBytes, nf, cudaMemcpyHostToDevice, st.sCopy));
			CVB_CUDA(cudaEventRecord(s.evIn, st.sCopy));
			CVB_CUDA(cudaStreamWaitEvent(s.stream, s.evIn, 0));
		}
		else {
			const size_t rawBytes = j.frameBytes * j.bpp;
			CVB_CHECK(s.raw.ensure(j.sub * rawBytes));
			CVB_CUDA(cudaMemcpy2DAsync(s.raw.p, rawBytes, j.images + f0 * j.framePitch, j.framePitch, rawBytes, nf, cudaMemcpyHostToDevice, st.sCopy));
			CVB_CUDA(cudaEventRecord(s.evIn, st.sCopy));
			CVB_CUDA(cudaStreamWaitEvent(s.stream, s.evIn, 0));
			CVB_CHECK(cvb200_image_to_grayscale_dev(j.pixelFormat, s.raw.as<uint8_t>(), j.width, j.height, j.stride, s.in.as<uint8_t>(), j.stride, nf, rawBytes, j.frameBytes,
				reinterpret_cast<cvb200_stream_t>(s.stream)));
		}
		s.dIn = s.in.as<uint8_t>(); s.inPitch = j.frameBytes;
	}
	else {
		CVB_CUDA(cudaStreamWaitEvent(s.stream, st.evStart, 0)); // ordered after what the caller had queued on its stream
		s.dIn = j.images + f0 * j.framePitch; s.inPitch = j.framePitch;
	}
	CVB_CHECK(s.edges.ensure(j.sub * j.frameBytes));
	const int slotId = static_cast<int>(f0 / j.sub);
	trace_mark(s.stream, "canny>", slotId);
	// the Canny finalize pass writes the linking bitmap and the edge counts of this sub-batch straight into the KHT object's buffers
	CVB_CHECK(kht_prepare_bits(&s.hough, j.width, j.height, nf, s.stream, &s.canny.khtBits, &s.canny.khtEdgeCount, &s.canny.khtWW));
	const int rcEdge = edge_enqueue(&s.canny, s.dIn, j.width, j.height, j.stride, s.edges.as<uint8_t>(), nf, s.inPitch, s.stream);
	s.canny.khtBits = nullptr; s.canny.khtEdgeCount = nullptr;
	if (rcEdge != CVB200_S_OK) { s.hough.bitsPrepared = false; return rcEdge; }
	trace_mark(s.stream, "canny<", slotId);
	s.hough.traceSlot = slotId;
	CVB_CHECK(kht_enqueue(&s.hough, s.edges.as<uint8_t>(), j.width, j.height, j.stride, nf, j.frameBytes, j.capacity, s.stream));
	trace_mark(s.stream, "kht<", slotId);
	s.busy = true;
	return CVB200_S_OK;
}

// waits for the slot, hands its lines out; a sub-batch whose pools were too small / whose hysteresis needed more rounds is simply run again (rare: first call, odd frames)
int slot_finish(PipeSlot& s, const Job& j)
{
	if (!s.busy) return CVB200_S_OK;
	s.busy = false;
	for (int attempt = 0; attempt < 16; ++attempt) {
		bool againE = false, againH = false;
		CVB_CHECK(edge_finish(&s.canny, &againE));
		if (!againE) CVB_CHECK(kht_finish(&s.hough, j.lines + s.f0 * j.capacity, j.capacity, j.counts + s.f0, &againH));
		if (!againE && !againH) return CVB200_S_OK;
		if (againE) CVB_CHECK(edge_enqueue(&s.canny, s.dIn, j.width, j.height, j.stride, s.edges.as<uint8_t>(), s.nf, s.inPitch, s.stream));
		CVB_CHECK(kht_enqueue(&s.hough, s.edges.as<uint8_t>(), j.width, j.height, j.stride, s.nf, j.frameBytes, j.capacity, s.stream));
	}
	return CVB200_E_INVALID_STATE;
}

int run_pipeline(Job& j, cudaStream_t callerStream)
{
	PipeState& st = t_pipe;
	if (!st.sCopy) {
		CVB_CUDA(cudaStreamCreateWithFlags(&st.sCopy, cudaStreamNonBlocking));
		CVB_CUDA(cudaEventCreateWithFlags(&st.evStart, cudaEventDisableTiming));
	}
	j.frameBytes = j.stride * j.height;
	// Sub-batches.  The linking kernel's duration hardly depends on how many frames it holds, and a slot's chain is Canny -> linking -> voting/peaks:
	//   * device frames: as many sub-batches as slots, all enqueued at once (a slot that has to be waited for and refilled leaves the GPU idle for one linking latency);
	//   * host frames: the upload paces everything, so a ring of smaller sub-batches keeps the tail (the last sub-batch's chain after its upload) short.
	// host frames: 256 frames take ~10 ms to upload and ~30 ms to flow through a slot while the GPU is busy with the others -> 8 slots keep the copy engine fed
	const size_t nSlotsWanted = static_cast<size_t>(std::max(1, env_int("CVB200_PIPE_SLOTS", j.onHost ? 8 : 6)));
	size_t sub;
	if (j.batch <= 8) sub = j.batch; // a handful of frames: one sub-batch, no ring
	else if (j.onHost) { sub = static_cast<size_t>(std::max(1, env_int("CVB200_PIPE_SUB", 256))); if (sub * 3 > j.batch) sub = std::max<size_t>(1, div_up(j.batch, 3)); }
	else { const int forced = env_int("CVB200_PIPE_SUB", 0); sub = forced > 0 ? static_cast<size_t>(forced) : std::max<size_t>(div_up(j.batch, nSlotsWanted), std::min<size_t>(j.batch, 16)); }
	j.sub = sub;
	const size_t nSub = div_up(j.batch, sub);
	const size_t nSlots = std::min(nSlotsWanted, nSub);
	// Stream priorities are fixed when a stream is created, so every slot owns one stream per mode (scratch and detectors are shared: a call drains all slots before it returns):
	//   * device frames: everything is queued at once on streams of EQUAL priority.  Measured (B200, 2048 frames, 6 slots): 44.2 ms per step with equal priorities against
	//     49.3 ms with "earlier slots first" and 49.7 ms with "later slots first".  The linking kernel is a latency chain per warp: sharing an SM with Canny blocks slows
	//     both down by more than the overlap wins, so the best schedule is the one where the Canny kernels of all slots finish together and the linking kernels then
	//     run side by side (more linking warps per scheduler = better issue utilisation), the short voting / peak stages filling the gaps;
	//   * host frames: the upload paces the call and the GPU is not full; what counts is how soon after the LAST upload the last sub-batch is done, so the newest
	//     sub-batch must not queue behind its predecessors: later slots get the higher priority and the ring is rotated so that the last sub-batch runs on the last slot
	//     (trace of 2048 host frames before this: the last sub-batch's Canny took 10 ms instead of 2.8).
	std::vector<PipeSlot*>& slots = st.slots;
	while (slots.size() < nSlots) {
		PipeSlot* s = new (std::nothrow) PipeSlot();
		CVB_REQUIRE(s, CVB200_E_OUT_OF_MEMORY);
		int prLo = 0, prHi = 0;
		CVB_CUDA(cudaDeviceGetStreamPriorityRange(&prLo, &prHi)); // numerically lower = higher priority
		const int idx = static_cast<int>(slots.size());
		const bool flat = env_int("CVB200_PIPE_PRIORITIES", 1) == 0;
		const int devMode = env_int("CVB200_PIPE_PRIO_DEV", 0); // 0 = equal (default), 1 = earlier slots higher, 2 = later slots higher
		CVB_CUDA(cudaStreamCreateWithPriority(&s->streamDev, cudaStreamNonBlocking, (flat || devMode == 0) ? prLo : devMode == 1 ? std::min(prLo, prHi + idx) : std::max(prHi, prLo - idx)));
		CVB_CUDA(cudaStreamCreateWithPriority(&s->streamHost, cudaStreamNonBlocking, flat ? prLo : std::max(prHi, prLo - idx)));
		CVB_CUDA(cudaEventCreateWithFlags(&s->evIn, cudaEventDisableTiming));
		s->hough.lastGs = 1.0;
		slots.push_back(s);
	}
	for (size_t i = 0; i < nSlots; ++i) slots[i]->stream = j.onHost ? slots[i]->streamHost : slots[i]->streamDev;
	for (size_t i = 0; i < nSlots; ++i) { copy_params(slots[i]->canny, *j.canny); copy_params(slots[i]->hough, *j.hough); slots[i]->busy = false; }
	if (!j.onHost) CVB_CUDA(cudaEventRecord(st.evStart, callerStream));
	// sub-batch k runs on slot (k + rot) % nSlots; host frames: rot puts the last sub-batch on the last (highest-priority) slot
	const size_t rot = j.onHost ? (nSlots - 1 + nSlots - ((nSub - 1) % nSlots)) % nSlots : 0;
	int rc = CVB200_S_OK;
	for (size_t k = 0; k < nSub && rc == CVB200_S_OK; ++k) {
		PipeSlot& s = *slots[(k + rot) % nSlots];
		rc = slot_finish(s, j); // the sub-batch that used this slot nSlots steps ago
		if (rc == CVB200_S_OK) { const size_t f0 = k * sub; rc = slot_enqueue(st, s, j, f0, std::min(sub, j.batch - f0)); }
	}
	// drain in submission order (also on errors: nothing may stay in flight on the cached streams)
	for (size_t k = 0; k < nSlots; ++k) {
		PipeSlot& s = *slots[(nSub + k + rot) % nSlots];
		const int r2 = (rc == CVB200_S_OK) ? slot_finish(s, j) : (cudaStreamSynchronize(s.stream), CVB200_S_OK);
		if (rc == CVB200_S_OK) rc = r2;
		s.busy = false;
	}
	trace_dump(j.onHost ? "canny+kht pipeline, host frames" : "canny+kht pipeline, device frames");
	if (rc == CVB200_S_OK) {
		j.hough->lastGs = slots[(nSub - 1 + rot) % nSlots]->hough.lastGs;
		for (size_t i = 0; i < nSlots; ++i) if (j.canny->hystRounds < slots[i]->canny.hystRounds) j.canny->hystRounds = slots[i]->canny.hystRounds;
	}
	return rc;
}

// ---- several devices in one process: one persistent worker thread per device, bound to it for life (cudaSetDevice is per thread) with its own ring of slots
// (the thread_local PipeState above).  Frames are independent, so a batch is cut into contiguous shards, one per device, with no exchange between them.
struct DeviceWorker {
	int device = 0;
	std::thread th;
	std::mutex m;
	std::condition_variable cv;
	std::function<int()> job;
	bool hasJob = false, done = false, stop = false;
	int rc = CVB200_S_OK;
	void loop() {
		bind_thread_to_device(device);
		for (;;) {
			std::function<int()> f;
			{
				std::unique_lock<std::mutex> lk(m);
				cv.wait(lk, [&] { return stop || hasJob; });
				if (stop) return;
				f = job; hasJob = false;
			}
			int r = ensure_device();
			if (r == CVB200_S_OK) r = f();
			{ std::lock_guard<std::mutex> lk(m); rc = r; done = true; }
			cv.notify_all();
		}
	}
	void submit(const std::function<int()>& f) { { std::lock_guard<std::mutex> lk(m); job = f; hasJob = true; done = false; } cv.notify_all(); }
	int wait() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return done; }); return rc; }
};
struct DevicePool {
	std::vector<DeviceWorker*> workers;
	std::mutex m;
	~DevicePool() { for (DeviceWorker* w : workers) { { std::lock_guard<std::mutex> lk(w->m); w->stop = true; } w->cv.notify_all(); if (w->th.joinable()) w->th.join(); delete w; } }
	DeviceWorker* get(int device) {
		std::lock_guard<std::mutex> lk(m);
		while (static_cast<int>(workers.size()) <= device) {
			DeviceWorker* w = new DeviceWorker();
			w->device = static_cast<int>(workers.size());
			w->th = std::thread([w] { w->loop(); });
			workers.push_back(w);
		}
		return workers[device];
	}
};
DevicePool& device_pool() { static DevicePool p; return p; }

} // namespace

extern "C" {

int cvb200_canny_kht_process_batch_fmt(cvb200_edge_dete_t* canny, cvb200_hough_t* hough, int pixelFormat, const uint8_t* frames, size_t width, size_t height, size_t stride,
	size_t batch, size_t framePitchBytes, cvb200_hough_line_t* lines, size_t capacity, size_t* counts)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(canny && hough && frames && counts && width && height && stride >= width && (lines || !capacity), CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(canny->id == CVB200_CANNY_ID && hough->id == CVB200_HOUGHKHT_ID, CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	size_t bpp = 0;
	CVB_CHECK(cvb200_image_bytes_per_sample(pixelFormat, &bpp));
	CVB_REQUIRE(framePitchBytes >= stride * bpp * height, CVB200_E_INVALID_PARAMETER); // planar formats: the pitch includes the chroma planes, it cannot be defaulted
	std::lock_guard<std::mutex> l1(canny->mutex);
	std::lock_guard<std::mutex> l2(hough->mutex);
	Job j = { canny, hough, frames, true, width, height, stride, batch, framePitchBytes, lines, capacity, counts, 0, 0 };
	j.pixelFormat = pixelFormat; j.bpp = bpp;
	return run_pipeline(j, nullptr);
}

int cvb200_canny_kht_process_batch_multi(cvb200_edge_dete_t* canny, cvb200_hough_t* hough, const uint8_t* images, size_t width, size_t height, size_t stride,
	size_t batch, size_t framePitch, cvb200_hough_line_t* lines, size_t capacity, size_t* counts)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(canny && hough && images && counts && width && height && stride >= width && (lines || !capacity), CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(canny->id == CVB200_CANNY_ID && hough->id == CVB200_HOUGHKHT_ID, CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	CVB_REQUIRE(framePitch >= stride * height, CVB200_E_INVALID_PARAMETER);
	const size_t nDev = std::min<size_t>(static_cast<size_t>(std::max(1, g_device_count.load())), batch);
	std::lock_guard<std::mutex> l1(canny->mutex);
	std::lock_guard<std::mutex> l2(hough->mutex);
	std::vector<Job> jobs(nDev);
	std::vector<cvb200_hough> houghCopies(nDev); // each shard reports its own lastGs; the caller's object gets the last shard's
	std::vector<cvb200_edge_dete> cannyCopies(nDev);
	const size_t per = div_up(batch, nDev);
	size_t used = 0;
	for (size_t d = 0; d < nDev; ++d) {
		const size_t f0 = d * per;
		if (f0 >= batch) break;
		const size_t nf = std::min(per, batch - f0);
		copy_params(houghCopies[d], *hough);
		houghCopies[d].lastGs = 1.0;
		copy_params(cannyCopies[d], *canny);
		jobs[d] = Job{ &cannyCopies[d], &houghCopies[d], images + f0 * framePitch, true, width, height, stride, nf, framePitch, lines ? lines + f0 * capacity : nullptr, capacity, counts + f0, 0, 0 };
		Job* jp = &jobs[d];
		device_pool().get(static_cast<int>(d))->submit([jp] { return run_pipeline(*jp, nullptr); });
		++used;
	}
	int rc = CVB200_S_OK;
	for (size_t d = 0; d < used; ++d) { const int r = device_pool().get(static_cast<int>(d))->wait(); if (rc == CVB200_S_OK) rc = r; }
	if (rc == CVB200_S_OK) {
		hough->lastGs = houghCopies[used - 1].lastGs;
		for (size_t d = 0; d < used; ++d) if (canny->hystRounds < cannyCopies[d].hystRounds) canny->hystRounds = cannyCopies[d].hystRounds;
	}
	return rc;
}

int cvb200_canny_kht_process_batch(cvb200_edge_dete_t* canny, cvb200_hough_t* hough, const uint8_t* images, size_t width, size_t height, size_t stride,
	size_t batch, size_t framePitch, cvb200_hough_line_t* lines, size_t capacity, size_t* counts)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(canny && hough && images && counts && width && height && stride >= width && (lines || !capacity), CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(canny->id == CVB200_CANNY_ID && hough->id == CVB200_HOUGHKHT_ID, CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	CVB_REQUIRE(framePitch >= stride * height, CVB200_E_INVALID_PARAMETER);
	std::lock_guard<std::mutex> l1(canny->mutex);
	std::lock_guard<std::mutex> l2(hough->mutex);
	Job j = { canny, hough, images, true, width, height, stride, batch, framePitch, lines, capacity, counts, 0, 0 };
	return run_pipeline(j, nullptr);
}

int cvb200_canny_kht_process_batch_dev(cvb200_edge_dete_t* canny, cvb200_hough_t* hough, const uint8_t* images, size_t width, size_t height, size_t stride,
	size_t batch, size_t framePitch, cvb200_hough_line_t* lines, size_t capacity, size_t* counts, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(canny && hough && images && counts && width && height && stride >= width && (lines || !capacity), CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(canny->id == CVB200_CANNY_ID && hough->id == CVB200_HOUGHKHT_ID, CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	CVB_REQUIRE(framePitch >= stride * height, CVB200_E_INVALID_PARAMETER);
	std::lock_guard<std::mutex> l1(canny->mutex);
	std::lock_guard<std::mutex> l2(hough->mutex);
	Job j = { canny, hough, images, false, width, height, stride, batch, framePitch, lines, capacity, counts, 0, 0 };
	return run_pipeline(j, as_stream(stream));
}

} // extern "C"
