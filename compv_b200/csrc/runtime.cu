// Runtime plumbing of libcompv_b200.so: device binding, error strings, memory/stream helpers.
// Replaces the reference's CompVGpu::init probe (gpu/compv_gpu.cxx:36-62), which only dlopen()s libcuda and sets a flag.
#include "common.cuh"

#include <condition_variable>
#include <thread>
#include <vector>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>

namespace cvb {

std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_device{-1};
std::atomic<int> g_device_count{0};
static int g_num_sms = 0;
static thread_local std::string t_last_cuda_error;

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
	char buf[512];
	snprintf(buf, sizeof(buf), "%s: %s (%s:%d)", cudaGetErrorName(e), what, file, line);
	t_last_cuda_error = buf;
	cudaGetLastError(); // clear sticky-less errors so that the next call starts clean
	return e == cudaErrorMemoryAllocation ? CVB200_E_OUT_OF_MEMORY : CVB200_E_CUDA;
}

int num_sms() { return g_num_sms > 0 ? g_num_sms : 148; }

static thread_local int t_wanted_device = -1; // >= 0: this thread is a multi-device worker
static thread_local int t_bound_device = -1;  // what cudaSetDevice was last called with on this thread
int bound_device() { return t_wanted_device >= 0 ? t_wanted_device : g_device.load(); }
void bind_thread_to_device(int device) { t_wanted_device = device; }
int ensure_device()
{
	const int want = bound_device();
	if (want != t_bound_device) {
		CVB_CUDA(cudaSetDevice(want));
		t_bound_device = want;
	}
	return CVB200_S_OK;
}
int set_max_smem_once(const void* func, int bytes, std::atomic<unsigned int>& flags)
{
	static std::mutex m;
	const unsigned int bit = 1u << (bound_device() & 31);
	if (flags.load(std::memory_order_acquire) & bit) return CVB200_S_OK;
	std::lock_guard<std::mutex> lock(m);
	if (flags.load(std::memory_order_relaxed) & bit) return CVB200_S_OK;
	CVB_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
	flags.fetch_or(bit, std::memory_order_release);
	return CVB200_S_OK;
}

// ---- per-kernel timing ----
std::atomic<int> g_profiling{0};
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static std::mutex g_prof_mutex;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;

static cudaEvent_t prof_event()
{
	if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
	cudaEvent_t e = nullptr;
	cudaEventCreate(&e);
	return e;
}

void profile_mark(const char* name, cudaStream_t stream, bool begin)
{
	std::lock_guard<std::mutex> lock(g_prof_mutex);
	if (begin) {
		ProfRec r; r.name = name; r.e0 = prof_event(); r.e1 = nullptr;
		cudaEventRecord(r.e0, stream);
		g_prof_recs.push_back(r);
	}
	else {
		for (size_t i = g_prof_recs.size(); i-- > 0;) {
			if (g_prof_recs[i].name == name && !g_prof_recs[i].e1) {
				g_prof_recs[i].e1 = prof_event();
				cudaEventRecord(g_prof_recs[i].e1, stream);
				break;
			}
		}
	}
}

// ---- timeline tracing ----
struct TraceRec { const char* tag; int slot; cudaEvent_t ev; };
static std::vector<TraceRec> g_trace;
static std::mutex g_trace_mutex;
bool trace_on() { static const bool on = getenv("CVB200_TRACE") != nullptr; return on; }
void trace_mark(cudaStream_t stream, const char* tag, int slot)
{
	if (!trace_on()) return;
	TraceRec r; r.tag = tag; r.slot = slot; r.ev = nullptr;
	cudaEventCreate(&r.ev);
	cudaEventRecord(r.ev, stream);
	std::lock_guard<std::mutex> lock(g_trace_mutex);
	g_trace.push_back(r);
}
void trace_dump(const char* title)
{
	if (!trace_on()) return;
	cudaDeviceSynchronize();
	std::lock_guard<std::mutex> lock(g_trace_mutex);
	if (g_trace.empty()) return;
	fprintf(stderr, "[cvb200 trace] %s\n", title);
	for (auto& r : g_trace) {
		float ms = 0.f;
		cudaEventElapsedTime(&ms, g_trace[0].ev, r.ev);
		fprintf(stderr, "[cvb200 trace]   slot %d %-14s %8.3f ms\n", r.slot, r.tag, ms);
	}
	for (auto& r : g_trace) cudaEventDestroy(r.ev);
	g_trace.clear();
}

} // namespace cvb

using namespace cvb;

extern "C" {

int cvb200_init(int device)
{
	CVB_REQUIRE(device >= 0, CVB200_E_INVALID_PARAMETER);
	// the pipelined entry points keep up to a dozen streams busy: ask for enough hardware work queues (only effective when this is the process's first CUDA call;
	// an application that initialises CUDA itself sets CUDA_DEVICE_MAX_CONNECTIONS=32 in its environment)
	setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
	int count = 0;
	CVB_CUDA(cudaGetDeviceCount(&count));
	CVB_REQUIRE(device < count, CVB200_E_INVALID_PARAMETER);
	CVB_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	CVB_CUDA(cudaGetDeviceProperties(&prop, device));
	// sm_100a cubins only: there is deliberately no PTX/other-arch fallback in this library
	if (prop.major != 10) {
		char buf[256];
		snprintf(buf, sizeof(buf), "device %d is sm_%d%d; libcompv_b200 is built for sm_100a only", device, prop.major, prop.minor);
		// stored through cuda_fail's thread-local so that cvb200_last_cuda_error() explains the refusal
		cvb::cuda_fail(cudaErrorNoKernelImageForDevice, buf, __FILE__, __LINE__);
		return CVB200_E_CUDA;
	}
	g_num_sms = prop.multiProcessorCount;
	CVB_CUDA(cudaFree(0)); // force primary-context creation
	g_device.store(device);
	if (g_device_count.load() < 1) g_device_count.store(1);
	return CVB200_S_OK;
}

// cvb200_init(int device_count).  Initialises devices 0 .. count-1 (count <= 0: every device of the process) for the *_multi batch entry points, which
// spread the frames of a batch over them (one worker thread + its own streams and scratch per device); every other entry point keeps running on device 0.
int cvb200_init_devices(int device_count)
{
	int count = 0;
	setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
	CVB_CUDA(cudaGetDeviceCount(&count));
	CVB_REQUIRE(count > 0, CVB200_E_CUDA);
	if (device_count <= 0 || device_count > count) device_count = count;
	for (int d = device_count - 1; d >= 0; --d) { // device 0 last: it stays the calling thread's current device
		CVB_CHECK(cvb200_init(d));
	}
	g_device_count.store(device_count);
	return CVB200_S_OK;
}

int cvb200_active_device_count(void) { return g_device.load() >= 0 ? g_device_count.load() : 0; }

int cvb200_deinit(void)
{
	g_device.store(-1);
	g_device_count.store(0);
	return CVB200_S_OK;
}

int cvb200_is_active(void) { return g_device.load() >= 0 ? 1 : 0; }

int cvb200_device_count(int* count)
{
	CVB_REQUIRE(count, CVB200_E_INVALID_PARAMETER);
	*count = 0;
	CVB_CUDA(cudaGetDeviceCount(count));
	return CVB200_S_OK;
}

const char* cvb200_error_string(int code)
{
	switch (code) {
	case CVB200_S_OK: return "S_OK";
	case CVB200_E_NOT_IMPLEMENTED: return "E_NOT_IMPLEMENTED";
	case CVB200_E_NOT_INITIALIZED: return "E_NOT_INITIALIZED";
	case CVB200_E_INVALID_CALL: return "E_INVALID_CALL";
	case CVB200_E_INVALID_STATE: return "E_INVALID_STATE";
	case CVB200_E_INVALID_PARAMETER: return "E_INVALID_PARAMETER";
	case CVB200_E_INVALID_SUBTYPE: return "E_INVALID_SUBTYPE";
	case CVB200_E_OUT_OF_MEMORY: return "E_OUT_OF_MEMORY";
	case CVB200_E_OUT_OF_BOUND: return "E_OUT_OF_BOUND";
	case CVB200_E_MEMORY_NOT_ALIGNED: return "E_MEMORY_NOT_ALIGNED";
	case CVB200_E_CUDA: return "E_CUDA";
	default: return "E_UNKNOWN";
	}
}

const char* cvb200_last_cuda_error(void) { return t_last_cuda_error.c_str(); }

uint64_t cvb200_launch_count(void) { return g_launches.load(); }

int cvb200_profile_begin(void)
{
	CVB_REQUIRE_INIT();
	g_profiling.store(1);
	return CVB200_S_OK;
}

// Writes "name count total_ms\n" lines (one per kernel name) into buf; stops profiling and recycles the events.
int cvb200_profile_end(char* buf, size_t bufSize)
{
	CVB_REQUIRE(buf && bufSize, CVB200_E_INVALID_PARAMETER);
	g_profiling.store(0);
	CVB_CUDA(cudaDeviceSynchronize());
	std::lock_guard<std::mutex> lock(g_prof_mutex);
	std::map<std::string, std::pair<uint64_t, double> > acc;
	for (auto& r : g_prof_recs) {
		if (r.e0 && r.e1) {
			float ms = 0.f;
			if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { auto& a = acc[r.name]; a.first += 1; a.second += ms; }
		}
		if (r.e0) g_prof_pool.push_back(r.e0);
		if (r.e1) g_prof_pool.push_back(r.e1);
	}
	g_prof_recs.clear();
	std::string out;
	char line[256];
	for (auto& kv : acc) {
		snprintf(line, sizeof(line), "%s %llu %.6f\n", kv.first.c_str(), static_cast<unsigned long long>(kv.second.first), kv.second.second);
		out += line;
	}
	CVB_REQUIRE(out.size() + 1 <= bufSize, CVB200_E_OUT_OF_BOUND);
	memcpy(buf, out.c_str(), out.size() + 1);
	return CVB200_S_OK;
}

int cvb200_malloc(void** dptr, size_t bytes)
{
	CVB_REQUIRE(dptr, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE_INIT();
	*dptr = nullptr;
	CVB_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
	return CVB200_S_OK;
}

int cvb200_free(void* dptr)
{
	if (dptr) CVB_CUDA(cudaFree(dptr));
	return CVB200_S_OK;
}

int cvb200_host_alloc(void** hptr, size_t bytes)
{
	CVB_REQUIRE(hptr, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE_INIT();
	*hptr = nullptr;
	CVB_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
	return CVB200_S_OK;
}

int cvb200_host_free(void* hptr)
{
	if (hptr) CVB_CUDA(cudaFreeHost(hptr));
	return CVB200_S_OK;
}

int cvb200_memcpy_h2d(void* dptr, const void* hptr, size_t bytes, cvb200_stream_t stream)
{
	CVB_REQUIRE(dptr && hptr, CVB200_E_INVALID_PARAMETER);
	CVB_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
	return CVB200_S_OK;
}

int cvb200_memcpy_d2h(void* hptr, const void* dptr, size_t bytes, cvb200_stream_t stream)
{
	CVB_REQUIRE(dptr && hptr, CVB200_E_INVALID_PARAMETER);
	CVB_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
	return CVB200_S_OK;
}

int cvb200_memset(void* dptr, int value, size_t bytes, cvb200_stream_t stream)
{
	CVB_REQUIRE(dptr, CVB200_E_INVALID_PARAMETER);
	CVB_CUDA(cudaMemsetAsync(dptr, value, bytes, as_stream(stream)));
	return CVB200_S_OK;
}

int cvb200_stream_create(cvb200_stream_t* stream)
{
	CVB_REQUIRE(stream, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE_INIT();
	cudaStream_t s;
	CVB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
	*stream = reinterpret_cast<cvb200_stream_t>(s);
	return CVB200_S_OK;
}

int cvb200_stream_destroy(cvb200_stream_t stream)
{
	if (stream) CVB_CUDA(cudaStreamDestroy(as_stream(stream)));
	return CVB200_S_OK;
}

int cvb200_stream_sync(cvb200_stream_t stream)
{
	CVB_CUDA(cudaStreamSynchronize(as_stream(stream)));
	return CVB200_S_OK;
}

} // extern "C"

// ---- host worker pool ----
namespace cvb {
namespace {
// One job at a time (callMutex).  Every job names its participants explicitly: a worker takes part in job `generation` only if it was alive when the job
// was published (its `seen` starts at the generation current at its creation, so a worker created while growing the pool never mistakes an older, finished
// job for a new one), and the job's function pointer is cleared before host_parallel_for returns.
struct HostPool {
	std::vector<std::thread> workers;
	std::mutex m;
	std::condition_variable cvWork, cvDone;
	const std::function<void(size_t)>* fn = nullptr;
	std::atomic<size_t> next{0};
	size_t n = 0, generation = 0, active = 0;
	bool stop = false;
	std::mutex callMutex; // one parallel_for at a time
	void run(size_t seen) {
		for (;;) {
			const std::function<void(size_t)>* f;
			size_t count;
			{
				std::unique_lock<std::mutex> lk(m);
				cvWork.wait(lk, [&] { return stop || generation != seen; });
				if (stop) return;
				seen = generation; f = fn; count = n;
			}
			if (f) for (size_t i; (i = next.fetch_add(1, std::memory_order_relaxed)) < count;) (*f)(i);
			{
				std::lock_guard<std::mutex> lk(m);
				if (--active == 0) cvDone.notify_all();
			}
		}
	}
	~HostPool() {
		{ std::lock_guard<std::mutex> lk(m); stop = true; }
		cvWork.notify_all();
		for (auto& t : workers) t.join();
	}
};
HostPool& host_pool() { static HostPool p; return p; }
} // namespace

static size_t g_host_threads_cap = 0; // 0 = hardware_concurrency (capped at 128); cvb200_set_host_threads

void host_parallel_for(size_t n, const std::function<void(size_t)>& fn)
{
	if (n == 0) return;
	HostPool& p = host_pool();
	size_t want = g_host_threads_cap ? g_host_threads_cap : std::thread::hardware_concurrency();
	if (want > 128) want = 128;
	if (want > n) want = n;
	if (want <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
	std::lock_guard<std::mutex> call(p.callMutex);
	while (p.workers.size() + 1 < want) {
		size_t gen;
		{ std::lock_guard<std::mutex> lk(p.m); gen = p.generation; } // no job is in flight here (callMutex): a new worker must wait for the NEXT generation
		p.workers.emplace_back([&p, gen] { p.run(gen); });
	}
	{
		std::lock_guard<std::mutex> lk(p.m);
		p.fn = &fn; p.n = n; p.next.store(0); p.active = p.workers.size(); ++p.generation;
	}
	p.cvWork.notify_all();
	for (size_t i; (i = p.next.fetch_add(1, std::memory_order_relaxed)) < n;) fn(i);
	std::unique_lock<std::mutex> lk(p.m);
	p.cvDone.wait(lk, [&] { return p.active == 0; });
	p.fn = nullptr; p.n = 0;
}
} // namespace cvb

extern "C" int cvb200_set_host_threads(int n)
{
	CVB_REQUIRE(n >= 0, CVB200_E_INVALID_PARAMETER);
	cvb::g_host_threads_cap = static_cast<size_t>(n);
	return CVB200_S_OK;
}

// test hook: fn(i) = out[i] += 1 through the pool (tests/test_abi.py grows the pool between calls)
extern "C" int cvb200_selftest_host_pool(size_t n, unsigned int* out)
{
	CVB_REQUIRE(out || !n, CVB200_E_INVALID_PARAMETER);
	cvb::host_parallel_for(n, [&](size_t i) { out[i] += 1; });
	return CVB200_S_OK;
}
