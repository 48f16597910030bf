// Shared device/host helpers for libcompv_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include <functional>
#include <mutex>

#include "../../include/cvb200.h"

namespace cvb {

// ---- runtime state (runtime.cu) ----
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_device;          // -1 when not initialised
extern std::atomic<int> g_device_count;    // devices initialised by cvb200_init_devices (1 after cvb200_init)
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int num_sms();

#define CVB_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return ::cvb::cuda_fail(e__, #x, __FILE__, __LINE__); } while (0)
#define CVB_CHECK(x) do { int c__ = (x); if (c__ != CVB200_S_OK) return c__; } while (0)
#define CVB_REQUIRE(cond, code) do { if (!(cond)) return (code); } while (0)
// Every entry point runs on the device its THREAD is bound to: the one given to cvb200_init (any thread of the application), or the one a multi-device worker
// thread was started for (pipeline.cu).  cudaSetDevice is per thread, so it is (re)applied here whenever the calling thread is not on that device yet.
int ensure_device();
int bound_device();
void bind_thread_to_device(int device);  // multi-device workers
#define CVB_REQUIRE_INIT() do { if (::cvb::g_device.load() < 0) return CVB200_E_NOT_INITIALIZED; CVB_CHECK(::cvb::ensure_device()); } while (0)
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per device and kernel: `flags` is a per-call-site bit set indexed by device
int set_max_smem_once(const void* func, int bytes, std::atomic<unsigned int>& flags);
// call after every kernel launch: counts it and surfaces launch-configuration errors immediately
#define CVB_LAUNCHED() do { ::cvb::g_launches.fetch_add(1, std::memory_order_relaxed); CVB_CUDA(cudaGetLastError()); } while (0)

static inline cudaStream_t as_stream(cvb200_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Per-kernel device timing (cvb200_profile_begin/_end): when enabled, every launch site brackets its kernel with CUDA events on the
// launching stream.  Disabled (the default) it costs one relaxed atomic load per launch.
extern std::atomic<int> g_profiling;
void profile_mark(const char* name, cudaStream_t stream, bool begin);
struct KernelScope {
	const char* name; cudaStream_t stream; bool on;
	KernelScope(const char* n, cudaStream_t s) : name(n), stream(s), on(g_profiling.load(std::memory_order_relaxed) != 0) { if (on) profile_mark(name, stream, true); }
	KernelScope(const char* n, cudaStream_t s, bool enable) : name(n), stream(s), on(enable && g_profiling.load(std::memory_order_relaxed) != 0) { if (on) profile_mark(name, stream, true); }
	~KernelScope() { if (on) profile_mark(name, stream, false); }
};

// Timeline tracing for pipeline tuning (environment CVB200_TRACE=1): trace_mark records an event on `stream` under `tag`; trace_dump (after a device
// synchronisation) prints every mark's time relative to the first one and forgets them.  Costs one getenv-cached flag test when off.
bool trace_on();
void trace_mark(cudaStream_t stream, const char* tag, int slot);
void trace_dump(const char* title);

// Grow-only device scratch buffer (the reference caches its scratch per object the same way, e.g. canny_dete.cxx:133-147)
struct DevBuf {
	void* p = nullptr;
	size_t bytes = 0;
	int ensure(size_t n) {
		if (n <= bytes && p) return CVB200_S_OK;
		if (p) { cudaFree(p); p = nullptr; bytes = 0; }
		if (n == 0) n = 256;
		cudaError_t e = cudaMalloc(&p, n);
		if (e != cudaSuccess) { p = nullptr; return e == cudaErrorMemoryAllocation ? (cudaGetLastError(), CVB200_E_OUT_OF_MEMORY) : cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); }
		bytes = n;
		return CVB200_S_OK;
	}
	void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
	template <typename T> T* as() const { return static_cast<T*>(p); }
};

// Pinned host scratch (for small read-backs: flags, counts, thresholds)
struct HostBuf {
	void* p = nullptr;
	size_t bytes = 0;
	int ensure(size_t n) {
		if (n <= bytes && p) return CVB200_S_OK;
		if (p) { cudaFreeHost(p); p = nullptr; bytes = 0; }
		cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocDefault);
		if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__); }
		bytes = n;
		return CVB200_S_OK;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
	template <typename T> T* as() const { return static_cast<T*>(p); }
};

static inline size_t div_up(size_t a, size_t b) { return (a + b - 1) / b; }

// Persistent host worker pool for the per-frame host stages (the std::sort that fixes the Hough detectors' tie order): fn(i) for i in [0, n), the calling
// thread takes part.  Workers are created once (at most 127) instead of per call; frames are handed out one at a time.
void host_parallel_for(size_t n, const std::function<void(size_t)>& fn);

// ---- device helpers ----
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

} // namespace cvb
