// a10 -- thresholding: global, Otsu (histogram + between-class variance scan) and adaptive (fixed-point mean + LUT).
// Replaces CompVImageThreshold::global / otsu / adaptive (base/image/compv_image_threshold.cxx:52-116,118-180,183-317,339-366),
// CompVMathHistogram::build for 8-bit data (base/math/compv_math_histogram.cxx:44-61) and CompVKernel::mean (base/compv_kernel.cxx:12-25).
//   histogram  : per-block shared-memory bins, one global atomicAdd per non-empty bin        HBM: 1 B/px read
//   otsu scan  : 256 steps in fp32 in the reference's operation order (explicit _rn intrinsics, no contraction), one thread per frame
//   global     : 16 px per thread, u8 compare                                                  HBM: 1 B/px read + 1 B/px written
//   adaptive   : fused kernel -- tile in shared memory, fixed-point horizontal then vertical mean (u8 intermediate, exactly K3 of convlt.cu),
//                then out = (in - mean > -delta) ? maxVal : 0 (the reference's 768-entry LUT)    HBM: 1 B/px read + 1 B/px written
#include "common.cuh"
#include "tma.cuh"

#include <cstring>

namespace cvb {

__global__ void __launch_bounds__(256)
histogram_kernel(const uint8_t* __restrict__ in, int W, int H, size_t stride, size_t framePitch, unsigned int* __restrict__ hist /* [batch][256] */)
{
	__shared__ unsigned int sh[256];
	sh[threadIdx.x] = 0;
	__syncthreads();
	const uint8_t* f = in + blockIdx.z * framePitch;
	for (int y = blockIdx.y; y < H; y += gridDim.y) {
		const uint8_t* row = f + static_cast<size_t>(y) * stride;
		if ((reinterpret_cast<uintptr_t>(row) & 3) == 0) {
			const int w4 = W >> 2;
			for (int i = threadIdx.x; i < w4; i += 256) {
				const unsigned int v = reinterpret_cast<const unsigned int*>(row)[i];
				atomicAdd(&sh[v & 0xff], 1u); atomicAdd(&sh[(v >> 8) & 0xff], 1u); atomicAdd(&sh[(v >> 16) & 0xff], 1u); atomicAdd(&sh[v >> 24], 1u);
			}
			for (int x = (w4 << 2) + threadIdx.x; x < W; x += 256) atomicAdd(&sh[row[x]], 1u);
		}
		else {
			for (int x = threadIdx.x; x < W; x += 256) atomicAdd(&sh[row[x]], 1u);
		}
	}
	__syncthreads();
	const unsigned int c = sh[threadIdx.x];
	if (c) atomicAdd(&hist[blockIdx.z * 256 + threadIdx.x], c);
}

// compv_image_threshold.cxx:77-104 (scan) + :349-366 (sumA256 / sum)
__global__ void otsu_scan_kernel(const unsigned int* __restrict__ hist, int N, double* __restrict__ thresholds, int batch)
{
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= batch) return;
	const unsigned int* h = hist + f * 256;
	unsigned int sum32 = 0;
	for (unsigned int i = 0; i < 256; ++i) sum32 += i * h[i];
	const float sumf = static_cast<float>(sum32);
	float sumB = 0.f, varMax = 0.f;
	int q1 = 0, q2 = 0, thr = 0;
	for (int i = 0; i < 256; ++i) {
		q1 += static_cast<int>(h[i]);
		if (q1) {
			q2 = N - q1;
			if (!q2) break;
			const float q1f = static_cast<float>(q1), q2f = static_cast<float>(q2);
			sumB = __fadd_rn(sumB, static_cast<float>(static_cast<unsigned int>(i) * h[i]));
			const float mf = __fsub_rn(__fdiv_rn(sumB, q1f), __fdiv_rn(__fsub_rn(sumf, sumB), q2f));
			const float varB = __fmul_rn(__fmul_rn(__fmul_rn(q1f, q2f), mf), mf);
			if (varB > varMax) { varMax = varB; thr = i; }
		}
	}
	thresholds[f] = static_cast<double>(thr);
}

// out = in > T ? 255 : 0 (compv_image_threshold.cxx:319-347). thresholds: per-frame doubles on the device, or nullptr -> tConst
__global__ void threshold_global_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int W, int H, size_t stride, size_t framePitch, const double* thresholds, int tConst)
{
	const int y = blockIdx.y;
	const size_t off = blockIdx.z * framePitch + static_cast<size_t>(y) * stride;
	int T = tConst;
	if (thresholds) {
		// thresholdUInt8 = ROUNDFU(clip(0, 255, threshold)) (compv_image_threshold.cxx:131-134)
		double t = thresholds[blockIdx.z];
		t = t < 0.0 ? 0.0 : (t > 255.0 ? 255.0 : t);
		T = static_cast<uint8_t>(t + 0.5);
	}
	const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
	if (x >= W) return;
	if (x + 16 <= W && (((reinterpret_cast<uintptr_t>(in + off + x) | reinterpret_cast<uintptr_t>(out + off + x)) & 15) == 0)) {
		const uint4 v = *reinterpret_cast<const uint4*>(in + off + x);
		auto cmp = [T](unsigned int w) {
			unsigned int r = 0;
#pragma unroll
			for (int j = 0; j < 4; ++j) if (static_cast<int>((w >> (8 * j)) & 0xff) > T) r |= 0xffu << (8 * j);
			return r;
		};
		uint4 o; o.x = cmp(v.x); o.y = cmp(v.y); o.z = cmp(v.z); o.w = cmp(v.w);
		*reinterpret_cast<uint4*>(out + off + x) = o;
	}
	else {
		for (int k = 0; k < 16 && x + k < W; ++k) out[off + x + k] = (in[off + x + k] > T) ? 0xff : 0;
	}
}

constexpr int AD_TW = 64, AD_TH = 32, AD_THREADS = 256, AD_MAX_TAPS = 63;
struct AdaptTaps { uint16_t vt[AD_MAX_TAPS]; uint16_t hz[AD_MAX_TAPS]; };

// KS > 0: compile-time kernel size (the block sizes callers use: 3, 5, 7), taps in registers, loops unrolled; KS == 0: any odd size
template <int KS>
__global__ void __launch_bounds__(AD_THREADS)
threshold_adaptive_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int W, int H, size_t stride, size_t framePitch, const AdaptTaps taps, int ksRuntime,
	int deltaInt, int maxVal, int invert)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int ks = KS > 0 ? KS : ksRuntime;
	unsigned int hzR[KS > 0 ? KS : 1], vtR[KS > 0 ? KS : 1];
	if (KS > 0) {
		#pragma unroll
		for (int k = 0; k < KS; ++k) { hzR[k] = taps.hz[k]; vtR[k] = taps.vt[k]; }
	}
	const int r = ks >> 1;
	const int tw = AD_TW + 2 * r, th = AD_TH + 2 * r;
	uint8_t* sIn = smem_raw;
	uint8_t* sMid = smem_raw + ((tw * th + 15) & ~15);
	const int x0 = blockIdx.x * AD_TW, y0 = blockIdx.y * AD_TH;
	in += blockIdx.z * framePitch; out += blockIdx.z * framePitch;
	const int tid = threadIdx.x;
	for (int ly = tid >> 5; ly < th; ly += AD_THREADS / 32) { // one warp per tile row: no division
		const int gy = y0 - r + ly;
		const bool rowIn = (gy >= 0 && gy < H);
		const uint8_t* src = in + static_cast<size_t>(rowIn ? gy : 0) * stride;
		for (int lx = tid & 31; lx < tw; lx += 32) {
			const int gx = x0 - r + lx;
			sIn[ly * tw + lx] = (rowIn && gx >= 0 && gx < W) ? src[gx] : 0;
		}
	}
	__syncthreads();
	// horizontal fixed-point mean (compv_math_convlt.h:386-405), zero on the r-wide column border
	for (int i = tid; i < AD_TW * th; i += AD_THREADS) {
		const int ly = i / AD_TW, lx = i - ly * AD_TW;
		const int gx = x0 + lx, gy = y0 - r + ly;
		unsigned int sum = 0;
		if (gy >= 0 && gy < H && gx >= r && gx < W - r) {
			const uint8_t* p = &sIn[ly * tw + lx];
			#pragma unroll
			for (int k = 0; k < ks; ++k) sum += (static_cast<unsigned int>(p[k]) * (KS > 0 ? hzR[KS > 0 ? k : 0] : static_cast<unsigned int>(taps.hz[k]))) >> 16;
			sum = sum > 255u ? 255u : sum;
		}
		sMid[i] = static_cast<uint8_t>(sum);
	}
	__syncthreads();
	// vertical pass (zero on the r-high row border) + LUT: lut[in - mean + 255], first (255 - delta + 1) entries "off" (compv_image_threshold.cxx:221-225, 283-286)
	const int onVal = invert ? 0 : maxVal, offVal = invert ? maxVal : 0;
	for (int i = tid; i < AD_TW * AD_TH; i += AD_THREADS) {
		const int ly = i / AD_TW, lx = i - ly * AD_TW;
		const int gx = x0 + lx, gy = y0 + ly;
		if (gx >= W || gy >= H) continue;
		unsigned int mean = 0;
		if (gy >= r && gy < H - r) {
			const uint8_t* p = &sMid[ly * AD_TW + lx];
			#pragma unroll
			for (int k = 0; k < ks; ++k) mean += (static_cast<unsigned int>(p[k * AD_TW]) * (KS > 0 ? vtR[KS > 0 ? k : 0] : static_cast<unsigned int>(taps.vt[k]))) >> 16;
			mean = mean > 255u ? 255u : mean;
		}
		const int idx = static_cast<int>(sIn[(ly + r) * tw + lx + r]) - static_cast<int>(mean) + 255;
		out[static_cast<size_t>(gy) * stride + gx] = static_cast<uint8_t>(idx >= (255 - deltaInt + 1) ? onVal : offVal);
	}
}

// ---- adaptive threshold, fast path: TMA-staged tile, 4 px per lane (same structure as convlt_fast.cuh), block sizes 3 / 5 / 7 ----
constexpr int AF_TW = 120, AF_TH = 60, AF_THREADS = 256, AF_WARPS = 8, AF_ROWW = 32, AF_INW = 36;

template <int KS>
__global__ void __launch_bounds__(AF_THREADS, 3)
threshold_adaptive_fast_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t* __restrict__ outAll, int W, int H, size_t stride, size_t framePitch, const AdaptTaps taps,
	int deltaInt, int maxVal, int invert, int vecStore)
{
	constexpr int R = KS >> 1;
	constexpr int IN_ROWS = AF_TH + 2 * R;
	extern __shared__ __align__(128) unsigned char af_smem[];
	const unsigned int pad = (128u - (static_cast<unsigned int>(__cvta_generic_to_shared(af_smem)) & 127u)) & 127u;
	unsigned int* sA = reinterpret_cast<unsigned int*>(af_smem + pad);
	unsigned int* sM = sA + IN_ROWS * AF_INW + 4;
	uint64_t* bar = reinterpret_cast<uint64_t*>(sM + IN_ROWS * AF_ROWW + 2);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int x0 = blockIdx.x * AF_TW, y0 = blockIdx.y * AF_TH, frame = blockIdx.z;
	const int xl = x0 - 4 + 4 * lane;
	const int yIn0 = y0 - R;
	const int xTma = (x0 - 4) & ~15;
	const int woff = ((x0 - 4) - xTma) >> 2;
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		mbar_expect_tx(bar, IN_ROWS * AF_INW * 4);
		tma_load_3d(sA, &tmap, bar, xTma, yIn0, frame);
	}
	__syncthreads();
	mbar_wait(bar, 0);
	unsigned int hz[KS], vt[KS];
#pragma unroll
	for (int k = 0; k < KS; ++k) { hz[k] = taps.hz[k]; vt[k] = taps.vt[k]; }
	// horizontal fixed-point mean (compv_math_convlt.h:386-405), zero on the R-wide column border
	unsigned int convMask = 0;
#pragma unroll
	for (int i = 0; i < 4; ++i) if (xl + i >= R && xl + i < W - R) convMask |= 0xffu << (8 * i);
	for (int r = warp; r < IN_ROWS; r += AF_WARPS) {
		const int y = yIn0 + r;
		unsigned int outw = 0;
		if (y >= 0 && y < H && convMask) {
			const unsigned int* q = &sA[r * AF_INW + woff + lane];
			const unsigned int wl = q[-1], wc = q[0], wr = q[1];
			unsigned int v[4 + 2 * R];
#pragma unroll
			for (int j = 0; j < R; ++j) v[j] = (wl >> (8 * (4 - R + j))) & 0xffu;
#pragma unroll
			for (int j = 0; j < 4; ++j) v[R + j] = (wc >> (8 * j)) & 0xffu;
#pragma unroll
			for (int j = 0; j < R; ++j) v[R + 4 + j] = (wr >> (8 * j)) & 0xffu;
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				unsigned int sum = 0;
#pragma unroll
				for (int k = 0; k < KS; ++k) sum += (v[i + k] * hz[k]) >> 16;
				outw |= min(sum, 255u) << (8 * i);
			}
			outw &= convMask;
		}
		sM[r * AF_ROWW + lane] = outw;
	}
	__syncthreads();
	// vertical pass (zero on the R-high row border) + LUT compare: lut[in - mean + 255], first (255 - delta + 1) entries "off" (compv_image_threshold.cxx:221-225, 283-286)
	const unsigned int onVal = invert ? 0u : static_cast<unsigned int>(maxVal), offVal = invert ? static_cast<unsigned int>(maxVal) : 0u;
	const int cut = 255 - deltaInt + 1;
	constexpr int RPW = (AF_TH + AF_WARPS - 1) / AF_WARPS;
	const int ro0 = warp * RPW;
	const bool laneOut = (lane >= 1 && lane <= 30) && xl < W;
	uint8_t* __restrict__ out = outAll + frame * framePitch;
#pragma unroll 1
	for (int j = 0; j < RPW; ++j) {
		const int ro = ro0 + j;
		const int y = y0 + ro;
		if (ro >= AF_TH || y >= H || !laneOut) continue;
		unsigned int mean[4] = { 0, 0, 0, 0 };
		if (y >= R && y < H - R) {
#pragma unroll
			for (int k = 0; k < KS; ++k) {
				const unsigned int w = sM[(ro + k) * AF_ROWW + lane];
#pragma unroll
				for (int i = 0; i < 4; ++i) mean[i] += (((w >> (8 * i)) & 0xffu) * vt[k]) >> 16;
			}
		}
		const unsigned int wc = sA[(ro + R) * AF_INW + woff + lane];
		unsigned int outw = 0;
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int idx = static_cast<int>((wc >> (8 * i)) & 0xffu) - static_cast<int>(min(mean[i], 255u)) + 255;
			outw |= (idx >= cut ? onVal : offVal) << (8 * i);
		}
		uint8_t* o8 = out + static_cast<size_t>(y) * stride + xl;
		if (vecStore && xl + 4 <= W) *reinterpret_cast<unsigned int*>(o8) = outw;
		else {
#pragma unroll
			for (int i = 0; i < 4; ++i) if (xl + i < W) o8[i] = static_cast<uint8_t>(outw >> (8 * i));
		}
	}
}

template <int KS>
static int adaptive_fast_launch(const uint8_t* in, uint8_t* out, size_t W, size_t H, size_t stride, size_t framePitch, const AdaptTaps& taps, int deltaInt, int maxVal, int invert,
	size_t batch, cudaStream_t stream)
{
	constexpr int IN_ROWS = AF_TH + 2 * (KS >> 1);
	alignas(64) CUtensorMap map;
	memset(&map, 0, sizeof(map));
	if (!make_u8_tile_map(&map, in, W, H, stride, framePitch, batch, AF_INW * 4, IN_ROWS)) return 1; // not addressable by the TMA unit: the caller takes the generic kernel
	const size_t smem = (static_cast<size_t>(IN_ROWS) * (AF_INW + AF_ROWW) + 8) * 4 + 128 + 16;
	dim3 grid(static_cast<unsigned>(div_up(W, AF_TW)), static_cast<unsigned>(div_up(H, AF_TH)), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	const int vecStore = (((reinterpret_cast<uintptr_t>(out) | stride | framePitch) & 3) == 0) ? 1 : 0;
	{ KernelScope ks_("threshold_adaptive", stream);
	  threshold_adaptive_fast_kernel<KS><<<grid, AF_THREADS, smem, stream>>>(map, out, static_cast<int>(W), static_cast<int>(H), stride, framePitch, taps, deltaInt, maxVal, invert, vecStore); }
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

static int launch_histogram(const uint8_t* in, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch, unsigned int* hist, cudaStream_t stream)
{
	CVB_CUDA(cudaMemsetAsync(hist, 0, batch * 256 * sizeof(unsigned int), stream));
	dim3 grid(1, static_cast<unsigned>(height < 128 ? height : 128), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	{ KernelScope ks_("histogram", stream);
	  histogram_kernel<<<grid, 256, 0, stream>>>(in, static_cast<int>(width), static_cast<int>(height), stride, framePitch, hist); }
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

static int launch_global(const uint8_t* in, uint8_t* out, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch, const double* dThr, int tConst, cudaStream_t stream)
{
	dim3 grid(static_cast<unsigned>(div_up(div_up(width, 16), 128)), static_cast<unsigned>(height), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	{ KernelScope ks_("threshold_global", stream);
	  threshold_global_kernel<<<grid, 128, 0, stream>>>(in, out, static_cast<int>(width), static_cast<int>(height), stride, framePitch, dThr, tConst); }
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

static std::mutex g_thr_mutex;
static DevBuf g_thr_in, g_thr_out, g_thr_hist;

} // namespace cvb

using namespace cvb;

extern "C" {

int cvb200_histogram_8u_dev(const uint8_t* in, size_t width, size_t height, size_t stride, unsigned int* hist, size_t batch, size_t framePitch, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(in && hist && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	return launch_histogram(in, width, height, stride, batch, framePitch ? framePitch : stride * height, hist, as_stream(stream));
}

int cvb200_histogram_8u(const uint8_t* in, size_t width, size_t height, size_t stride, unsigned int* hist)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(in && hist && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	std::lock_guard<std::mutex> lock(g_thr_mutex);
	const size_t n = stride * height;
	CVB_CHECK(g_thr_in.ensure(n));
	CVB_CHECK(g_thr_hist.ensure(256 * 4 + 8));
	CVB_CUDA(cudaMemcpyAsync(g_thr_in.p, in, n, cudaMemcpyHostToDevice, 0));
	CVB_CHECK(launch_histogram(g_thr_in.as<uint8_t>(), width, height, stride, 1, n, g_thr_hist.as<unsigned int>(), 0));
	CVB_CUDA(cudaMemcpyAsync(hist, g_thr_hist.p, 256 * 4, cudaMemcpyDeviceToHost, 0));
	CVB_CUDA(cudaStreamSynchronize(0));
	return CVB200_S_OK;
}

int cvb200_threshold_global_dev(const uint8_t* in, size_t width, size_t height, size_t stride, double threshold, uint8_t* out, size_t batch, size_t framePitch, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(in && out && width && height && stride >= width && !(threshold < 0), CVB200_E_INVALID_PARAMETER); // compv_image_threshold.cxx:120
	if (!batch) return CVB200_S_OK;
	const double t = threshold > 255.0 ? 255.0 : threshold;
	return launch_global(in, out, width, height, stride, batch, framePitch ? framePitch : stride * height, nullptr, static_cast<uint8_t>(t + 0.5), as_stream(stream));
}

int cvb200_threshold_otsu_dev(const uint8_t* in, size_t width, size_t height, size_t stride, double* thresholds /* device, [batch] */, uint8_t* out /* may be NULL */,
	unsigned int* histScratch /* device, [batch*256] */, size_t batch, size_t framePitch, cvb200_stream_t stream_)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(in && thresholds && histScratch && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(width * height <= 0x7fffffff, CVB200_E_OUT_OF_BOUND);
	if (!batch) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	cudaStream_t stream = as_stream(stream_);
	CVB_CHECK(launch_histogram(in, width, height, stride, batch, framePitch, histScratch, stream));
	{ KernelScope ks_("otsu_scan", stream);
	  otsu_scan_kernel<<<static_cast<unsigned>(div_up(batch, 64)), 64, 0, stream>>>(histScratch, static_cast<int>(width * height), thresholds, static_cast<int>(batch)); }
	CVB_LAUNCHED();
	if (out) CVB_CHECK(launch_global(in, out, width, height, stride, batch, framePitch, thresholds, 0, stream));
	return CVB200_S_OK;
}

static int adaptive_launch(const uint8_t* in, size_t width, size_t height, size_t stride, const uint16_t* kernelVt, const uint16_t* kernelHz, size_t kernSize,
	double delta, double maxVal, int invert, uint8_t* out, size_t batch, size_t framePitch, cudaStream_t stream)
{
	CVB_REQUIRE(in && out && in != out && kernelVt && kernelHz && (kernSize & 1) && !(maxVal < 0) && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(kernSize <= AD_MAX_TAPS, CVB200_E_NOT_IMPLEMENTED);
	CVB_REQUIRE(width >= kernSize && height >= kernSize, CVB200_E_INVALID_PARAMETER); // convlt1 precondition (compv_math_convlt.h:101)
	if (!batch) return CVB200_S_OK;
	AdaptTaps taps;
	memset(&taps, 0, sizeof(taps));
	for (size_t i = 0; i < kernSize; ++i) { taps.vt[i] = kernelVt[i]; taps.hz[i] = kernelHz[i]; }
	// compv_image_threshold.cxx:210-217
	const double dc = delta < 0.0 ? 0.0 : (delta > 255.0 ? 255.0 : delta);
	const double mc = maxVal > 255.0 ? 255.0 : maxVal;
	const int deltaInt = static_cast<int>(dc + 0.5);
	const int maxValU8 = static_cast<uint8_t>(mc + 0.5);
	const int r = static_cast<int>(kernSize >> 1);
	const size_t tw = AD_TW + 2 * r, th = AD_TH + 2 * r;
	const size_t smem = ((tw * th + 15) & ~static_cast<size_t>(15)) + AD_TW * th;
	{
		const size_t fp = framePitch ? framePitch : stride * height;
		int rc = 1;
		if (kernSize == 3) rc = adaptive_fast_launch<3>(in, out, width, height, stride, fp, taps, deltaInt, maxValU8, invert ? 1 : 0, batch, stream);
		else if (kernSize == 5) rc = adaptive_fast_launch<5>(in, out, width, height, stride, fp, taps, deltaInt, maxValU8, invert ? 1 : 0, batch, stream);
		else if (kernSize == 7) rc = adaptive_fast_launch<7>(in, out, width, height, stride, fp, taps, deltaInt, maxValU8, invert ? 1 : 0, batch, stream);
		if (rc != 1) return rc;
	}
	auto kern = threshold_adaptive_kernel<0>;
	if (kernSize == 3) kern = threshold_adaptive_kernel<3>;
	else if (kernSize == 5) kern = threshold_adaptive_kernel<5>;
	else if (kernSize == 7) kern = threshold_adaptive_kernel<7>;
	if (smem > 48 * 1024) CVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
	dim3 grid(static_cast<unsigned>(div_up(width, AD_TW)), static_cast<unsigned>(div_up(height, AD_TH)), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	{ KernelScope ks_("threshold_adaptive", stream);
	  kern<<<grid, AD_THREADS, smem, stream>>>(in, out, static_cast<int>(width), static_cast<int>(height), stride, framePitch ? framePitch : stride * height,
		taps, static_cast<int>(kernSize), deltaInt, maxValU8, invert ? 1 : 0); }
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

int cvb200_threshold_adaptive_kernel_dev(const uint8_t* in, size_t width, size_t height, size_t stride, const uint16_t* kernelVt, const uint16_t* kernelHz, size_t kernSize,
	double delta, double maxVal, int invert, uint8_t* out, size_t batch, size_t framePitch, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	return adaptive_launch(in, width, height, stride, kernelVt, kernelHz, kernSize, delta, maxVal, invert, out, batch, framePitch, as_stream(stream));
}

// CompVKernel::mean (base/compv_kernel.cxx:12-25): uint16(1.f/blockSize * 0xffff) on every tap
int cvb200_kernel_mean_fxp(size_t blockSize, uint16_t* kernel)
{
	CVB_REQUIRE(kernel && (blockSize & 1), CVB200_E_INVALID_PARAMETER);
	const float vvv = 1.f / static_cast<float>(blockSize);
	for (size_t i = 0; i < blockSize; ++i) kernel[i] = static_cast<uint16_t>(vvv * 0xffff);
	return CVB200_S_OK;
}

int cvb200_threshold_adaptive_dev(const uint8_t* in, size_t width, size_t height, size_t stride, size_t blockSize, double delta, double maxVal, int invert, uint8_t* out,
	size_t batch, size_t framePitch, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE((blockSize & 1) && blockSize <= AD_MAX_TAPS, blockSize > AD_MAX_TAPS ? CVB200_E_NOT_IMPLEMENTED : CVB200_E_INVALID_PARAMETER);
	uint16_t k[AD_MAX_TAPS];
	CVB_CHECK(cvb200_kernel_mean_fxp(blockSize, k));
	return adaptive_launch(in, width, height, stride, k, k, blockSize, delta, maxVal, invert, out, batch, framePitch, as_stream(stream));
}

// ---- host-buffer entry points (synchronous) ----
int cvb200_threshold_global(const uint8_t* in, size_t width, size_t height, size_t stride, double threshold, uint8_t* out)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(in && out && width && height && stride >= width && !(threshold < 0), CVB200_E_INVALID_PARAMETER);
	std::lock_guard<std::mutex> lock(g_thr_mutex);
	const size_t n = stride * height;
	CVB_CHECK(g_thr_in.ensure(n)); CVB_CHECK(g_thr_out.ensure(n));
	CVB_CUDA(cudaMemcpyAsync(g_thr_in.p, in, n, cudaMemcpyHostToDevice, 0));
	CVB_CHECK(cvb200_threshold_global_dev(g_thr_in.as<uint8_t>(), width, height, stride, threshold, g_thr_out.as<uint8_t>(), 1, n, nullptr));
	CVB_CUDA(cudaMemcpy2DAsync(out, stride, g_thr_out.p, stride, width, height, cudaMemcpyDeviceToHost, 0));
	CVB_CUDA(cudaStreamSynchronize(0));
	return CVB200_S_OK;
}

int cvb200_threshold_otsu(const uint8_t* in, size_t width, size_t height, size_t stride, double* threshold, uint8_t* out /* may be NULL */)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(in && threshold && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	std::lock_guard<std::mutex> lock(g_thr_mutex);
	const size_t n = stride * height;
	CVB_CHECK(g_thr_in.ensure(n)); CVB_CHECK(g_thr_out.ensure(n)); CVB_CHECK(g_thr_hist.ensure(256 * 4 + 8));
	double* dThr = reinterpret_cast<double*>(g_thr_hist.as<unsigned char>() + 256 * 4);
	CVB_CUDA(cudaMemcpyAsync(g_thr_in.p, in, n, cudaMemcpyHostToDevice, 0));
	CVB_CHECK(cvb200_threshold_otsu_dev(g_thr_in.as<uint8_t>(), width, height, stride, dThr, out ? g_thr_out.as<uint8_t>() : nullptr, g_thr_hist.as<unsigned int>(), 1, n, nullptr));
	CVB_CUDA(cudaMemcpyAsync(threshold, dThr, sizeof(double), cudaMemcpyDeviceToHost, 0));
	if (out) CVB_CUDA(cudaMemcpy2DAsync(out, stride, g_thr_out.p, stride, width, height, cudaMemcpyDeviceToHost, 0));
	CVB_CUDA(cudaStreamSynchronize(0));
	return CVB200_S_OK;
}

int cvb200_threshold_adaptive(const uint8_t* in, size_t width, size_t height, size_t stride, size_t blockSize, double delta, double maxVal, int invert, uint8_t* out)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(in && out && width && height && stride >= width && (blockSize & 1), CVB200_E_INVALID_PARAMETER);
	std::lock_guard<std::mutex> lock(g_thr_mutex);
	const size_t n = stride * height;
	CVB_CHECK(g_thr_in.ensure(n)); CVB_CHECK(g_thr_out.ensure(n));
	CVB_CUDA(cudaMemcpyAsync(g_thr_in.p, in, n, cudaMemcpyHostToDevice, 0));
	CVB_CHECK(cvb200_threshold_adaptive_dev(g_thr_in.as<uint8_t>(), width, height, stride, blockSize, delta, maxVal, invert, g_thr_out.as<uint8_t>(), 1, n, nullptr));
	CVB_CUDA(cudaMemcpy2DAsync(out, stride, g_thr_out.p, stride, width, height, cudaMemcpyDeviceToHost, 0));
	CVB_CUDA(cudaStreamSynchronize(0));
	return CVB200_S_OK;
}

// The Otsu scan on the host for a histogram that was summed elsewhere (row-strip mode: every GPU histograms its strip, the 256 counters are all-reduced).  Same
// operation order and fp32 arithmetic as otsu_scan_kernel / compv_image_threshold.cxx:77-104 (separate multiplies and divides; x86-64 without -mfma does not contract).
int cvb200_otsu_threshold_from_histogram(const uint32_t* h, size_t pixelCount, double* threshold)
{
	CVB_REQUIRE(h && threshold && pixelCount && pixelCount < (1ull << 31), CVB200_E_INVALID_PARAMETER);
	const int N = static_cast<int>(pixelCount);
	unsigned int sum32 = 0;
	for (unsigned int i = 0; i < 256; ++i) sum32 += i * h[i];
	const volatile float sumf = static_cast<float>(sum32);
	volatile float sumB = 0.f, varMax = 0.f;
	int q1 = 0, q2 = 0, thr = 0;
	for (int i = 0; i < 256; ++i) {
		q1 += static_cast<int>(h[i]);
		if (q1) {
			q2 = N - q1;
			if (!q2) break;
			const volatile float q1f = static_cast<float>(q1), q2f = static_cast<float>(q2);
			sumB = sumB + static_cast<float>(static_cast<unsigned int>(i) * h[i]);
			const volatile float a = sumB / q1f, b = (sumf - sumB) / q2f;
			const volatile float mf = a - b;
			const volatile float t0 = q1f * q2f;
			const volatile float t1 = t0 * mf;
			const volatile float varB = t1 * mf;
			if (varB > varMax) { varMax = varB; thr = i; }
		}
	}
	*threshold = static_cast<double>(thr);
	return CVB200_S_OK;
}

} // extern "C"
