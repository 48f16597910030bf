// a2 -- separable convolution (K1, K2, K3).
// Replaces CompVMathConvlt::convlt1<In,Kern,Out> / convlt1FixedPoint
// (reference: base/include/compv/base/math/compv_math_convlt.h:98-173 driver, :176-292 border handling, :332-405 arithmetic).
//
// One fused launch does both passes: a (TW+2r)x(TH+2r) input tile is staged in shared memory, the horizontal pass writes
// a TWx(TH+2r) intermediate tile of OutputType (the reference stores its intermediate plane as OutputType too, so the
// saturation / truncation between the passes is part of the arithmetic contract), the vertical pass writes the output.
// HBM traffic = 1 read of the input (+ halo re-reads that hit L2) + 1 write of the output; the intermediate plane of the
// CPU implementation (sizeof(Out) B/px written and re-read) never exists.
#include "common.cuh"
#include "convlt_fast.cuh"

#include <type_traits>
#include <cmath>
#include <limits>
#include <cstring>

namespace cvb {

constexpr int CONV_TW = 64;
constexpr int CONV_TH = 32;
constexpr int CONV_THREADS = 256;
constexpr int CONV_MAX_TAPS = 63;

template <typename K>
struct Taps {
	K vt[CONV_MAX_TAPS];
	K hz[CONV_MAX_TAPS];
};

// Arithmetic of one output sample. `p` points at the first tap's sample, `step` is the distance between taps.
// int kernels : compv_math_convlt.h:332-353 (int accumulate, clip to OutputType range)
// float kernels: compv_math_convlt.h:358-384 + the AVX2 leaf the oracle actually runs (intrin/x86/compv_math_convlt_intrin_avx2.cxx:65-310),
//               which GCC contracts to sum = fma(v, c, sum) in tap order starting from 0 (reproduces the reference's md5_fma goldens,
//               unittests/math_convlt.cxx:17-26); conversion to u8 truncates (cvttps) and saturates.
// fixed point  : compv_math_convlt.h:386-405 (sum of (v*k)>>16, clip 0..255)
// KS > 0: the kernel size is a compile-time constant (loops unroll, taps live in registers); KS == 0: any odd size up to CONV_MAX_TAPS.
template <typename In, typename K, typename Out, bool FXP, int KS>
__device__ __forceinline__ Out conv_sample(const In* p, int step, const K* taps, int ksRuntime)
{
	const int ks = KS > 0 ? KS : ksRuntime;
	if constexpr (FXP) {
		unsigned int sum = 0;
		#pragma unroll
		for (int k = 0; k < ks; ++k) sum += (static_cast<unsigned int>(p[k * step]) * static_cast<unsigned int>(taps[k])) >> 16;
		return static_cast<Out>(sum > 255u ? 255u : sum);
	}
	else if constexpr (std::is_floating_point<K>::value) {
		float sum = 0.f;
		#pragma unroll
		for (int k = 0; k < ks; ++k) sum = __fmaf_rn(static_cast<float>(p[k * step]), taps[k], sum);
		if constexpr (std::is_same<Out, uint8_t>::value) {
			sum = fminf(fmaxf(sum, 0.f), 255.f);
			return static_cast<uint8_t>(__float2int_rz(sum));
		}
		else {
			return static_cast<Out>(sum);
		}
	}
	else {
		int sum = 0;
		#pragma unroll
		for (int k = 0; k < ks; ++k) sum += static_cast<int>(p[k * step]) * static_cast<int>(taps[k]);
		constexpr int lo = std::is_signed<Out>::value ? -(1 << (8 * sizeof(Out) - 1)) : 0;
		constexpr int hi = std::is_signed<Out>::value ? (1 << (8 * sizeof(Out) - 1)) - 1 : (1 << (8 * sizeof(Out))) - 1;
		return static_cast<Out>(clampi(sum, lo, hi));
	}
}

template <typename In, typename K, typename Out, bool FXP, int KS>
__global__ void __launch_bounds__(CONV_THREADS)
convlt1_kernel(const In* __restrict__ in, Out* __restrict__ out, int W, int H, size_t stride, size_t framePitch, const Taps<K> taps, int ksRuntime, int border)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int ks = KS > 0 ? KS : ksRuntime;
	const int r = ks >> 1;
	K hz[KS > 0 ? KS : 1], vt[KS > 0 ? KS : 1]; // register copies of the taps when the size is known
	if (KS > 0) {
		#pragma unroll
		for (int k = 0; k < KS; ++k) { hz[k] = taps.hz[k]; vt[k] = taps.vt[k]; }
	}
	const K* hzTaps = KS > 0 ? hz : taps.hz;
	const K* vtTaps = KS > 0 ? vt : taps.vt;
	const int tw = CONV_TW + 2 * r, th = CONV_TH + 2 * r;
	In* sIn = reinterpret_cast<In*>(smem_raw);
	const size_t inBytes = (static_cast<size_t>(tw) * th * sizeof(In) + 15) & ~static_cast<size_t>(15);
	Out* sMid = reinterpret_cast<Out*>(smem_raw + inBytes);

	const int x0 = blockIdx.x * CONV_TW, y0 = blockIdx.y * CONV_TH;
	in += blockIdx.z * framePitch;
	out += blockIdx.z * framePitch;
	const int tid = threadIdx.x;

	// stage the input tile (+halo), one warp per tile row; samples outside the image are never used by a valid output, store 0
	for (int ly = tid >> 5; ly < th; ly += CONV_THREADS / 32) {
		const int gy = y0 - r + ly;
		const bool rowIn = (gy >= 0 && gy < H);
		const In* src = in + static_cast<size_t>(rowIn ? gy : 0) * stride;
		for (int lx = tid & 31; lx < tw; lx += 32) {
			const int gx = x0 - r + lx;
			In v = 0;
			if (rowIn && gx >= 0 && gx < W) v = src[gx];
			sIn[ly * tw + lx] = v;
		}
	}
	__syncthreads();

	// horizontal pass -> intermediate tile (compv_math_convlt.h:176-229)
	for (int i = tid; i < CONV_TW * th; i += CONV_THREADS) {
		const int ly = i / CONV_TW, lx = i - ly * CONV_TW;
		const int gx = x0 + lx, gy = y0 - r + ly;
		Out m = 0;
		if (gy >= 0 && gy < H && gx < W) {
			if (gx >= r && gx < W - r) {
				m = conv_sample<In, K, Out, FXP, KS>(&sIn[ly * tw + lx], 1, hzTaps, ks);
			}
			else if (border == CVB200_BORDER_TYPE_REPLICATE) {
				m = static_cast<Out>(sIn[ly * tw + lx + r]);
			}
		}
		sMid[i] = m;
	}
	__syncthreads();

	// vertical pass -> output (compv_math_convlt.h:231-292)
	for (int i = tid; i < CONV_TW * CONV_TH; i += CONV_THREADS) {
		const int ly = i / CONV_TW, lx = i - ly * CONV_TW;
		const int gx = x0 + lx, gy = y0 + ly;
		if (gx >= W || gy >= H) continue;
		Out* o = &out[static_cast<size_t>(gy) * stride + gx];
		if (gy >= r && gy < H - r) {
			if (border == CVB200_BORDER_TYPE_IGNORE && (gx < r || gx >= W - r)) continue;
			*o = conv_sample<Out, K, Out, FXP, KS>(&sMid[ly * CONV_TW + lx], CONV_TW, vtTaps, ks);
		}
		else if (border == CVB200_BORDER_TYPE_ZERO) {
			*o = 0;
		}
		else if (border == CVB200_BORDER_TYPE_REPLICATE) {
			*o = sMid[(ly + r) * CONV_TW + lx];
		}
	}
}

template <typename In, typename K, typename Out, bool FXP>
static int convlt1_launch(const In* in, size_t width, size_t height, size_t stride, const K* vtKern, const K* hzKern, size_t kernSize, Out* out, int borderType,
	size_t batch, size_t framePitch, cudaStream_t stream)
{
	CVB_REQUIRE_INIT();
	// same parameter checks as compv_math_convlt.h:101
	CVB_REQUIRE(in && out && vtKern && hzKern && (kernSize & 1) && width >= kernSize && height >= kernSize && stride >= width, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(borderType == CVB200_BORDER_TYPE_ZERO || borderType == CVB200_BORDER_TYPE_REPLICATE || borderType == CVB200_BORDER_TYPE_IGNORE, CVB200_E_NOT_IMPLEMENTED);
	CVB_REQUIRE(kernSize <= CONV_MAX_TAPS, CVB200_E_NOT_IMPLEMENTED);
	CVB_REQUIRE(static_cast<const void*>(in) != static_cast<const void*>(out), CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(width <= 0x7fffffff / 4 && height <= 0x7fffffff / 4, CVB200_E_OUT_OF_BOUND);
	if (batch == 0) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	Taps<K> taps;
	memset(&taps, 0, sizeof(taps));
	for (size_t i = 0; i < kernSize; ++i) { taps.vt[i] = vtKern[i]; taps.hz[i] = hzKern[i]; }
	if constexpr (std::is_same<In, uint8_t>::value && std::is_same<K, float>::value && std::is_same<Out, uint8_t>::value && !FXP) {
		// the Gaussian-blur case: TMA-staged tile, 4 px per lane (convlt_fast.cuh); frames the TMA unit cannot address fall through to the generic kernel
		int rc = 1;
		if (kernSize == 3) rc = launch_convlt_fast<3>(in, out, width, height, stride, framePitch, vtKern, hzKern, borderType, batch, stream);
		else if (kernSize == 5) rc = launch_convlt_fast<5>(in, out, width, height, stride, framePitch, vtKern, hzKern, borderType, batch, stream);
		if (rc != 1) return rc;
	}
	const int r = static_cast<int>(kernSize >> 1);
	const size_t tw = CONV_TW + 2 * r, th = CONV_TH + 2 * r;
	const size_t smem = ((tw * th * sizeof(In) + 15) & ~static_cast<size_t>(15)) + CONV_TW * th * sizeof(Out);
	auto kern = convlt1_kernel<In, K, Out, FXP, 0>;
	if (kernSize == 3) kern = convlt1_kernel<In, K, Out, FXP, 3>;       // the sizes the reference's callers use (Sobel 3, Gaussian 5 / 7, adaptive mean 5)
	else if (kernSize == 5) kern = convlt1_kernel<In, K, Out, FXP, 5>;
	else if (kernSize == 7) kern = convlt1_kernel<In, K, Out, FXP, 7>;
	if (smem > 48 * 1024) CVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
	dim3 grid(static_cast<unsigned>(div_up(width, CONV_TW)), static_cast<unsigned>(div_up(height, CONV_TH)), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	{
		KernelScope ks_("convlt1", stream);
		kern<<<grid, CONV_THREADS, smem, stream>>>(in, out, static_cast<int>(width), static_cast<int>(height), stride, framePitch, taps, static_cast<int>(kernSize), borderType);
	}
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

// Host-buffer front end: H2D, launch, D2H, synchronous (shape of the reference hook gpu_convlt1VtHz_8u8u32f,
// gpu/include/compv/gpu/base/math/compv_gpu_math_convlt.h:21-27)
static std::mutex g_host_mutex;
static DevBuf g_host_in, g_host_out;

template <typename In, typename K, typename Out, bool FXP>
static int convlt1_host(const In* in, size_t width, size_t height, size_t stride, const K* vtKern, const K* hzKern, size_t kernSize, Out* out, int borderType)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(in && out && stride >= width && width && height, CVB200_E_INVALID_PARAMETER);
	std::lock_guard<std::mutex> lock(g_host_mutex);
	const size_t n = stride * height;
	CVB_CHECK(g_host_in.ensure(n * sizeof(In)));
	CVB_CHECK(g_host_out.ensure(n * sizeof(Out)));
	CVB_CUDA(cudaMemcpyAsync(g_host_in.p, in, n * sizeof(In), cudaMemcpyHostToDevice, 0));
	if (borderType == CVB200_BORDER_TYPE_IGNORE) { // untouched samples must keep the caller's values
		CVB_CUDA(cudaMemcpyAsync(g_host_out.p, out, n * sizeof(Out), cudaMemcpyHostToDevice, 0));
	}
	CVB_CHECK((convlt1_launch<In, K, Out, FXP>(g_host_in.as<In>(), width, height, stride, vtKern, hzKern, kernSize, g_host_out.as<Out>(), borderType, 1, 0, 0)));
	// copy back only the rows' `width` samples: the stride padding of `out` belongs to the caller
	CVB_CUDA(cudaMemcpy2DAsync(out, stride * sizeof(Out), g_host_out.p, stride * sizeof(Out), width * sizeof(Out), height, cudaMemcpyDeviceToHost, 0));
	CVB_CUDA(cudaStreamSynchronize(0));
	return CVB200_S_OK;
}

} // namespace cvb

using namespace cvb;

extern "C" {

#define CVB_CONVLT_ENTRY(NAME, IN, KERN, OUT, FXP) \
int cvb200_convlt1_##NAME(const IN* in, size_t width, size_t height, size_t stride, const KERN* vtKern, const KERN* hzKern, size_t kernSize, OUT* out, int borderType) \
{ return convlt1_host<IN, KERN, OUT, FXP>(in, width, height, stride, vtKern, hzKern, kernSize, out, borderType); } \
int cvb200_convlt1_##NAME##_dev(const IN* in, size_t width, size_t height, size_t stride, const KERN* vtKern, const KERN* hzKern, size_t kernSize, OUT* out, int borderType, size_t batch, size_t framePitch, cvb200_stream_t stream) \
{ return convlt1_launch<IN, KERN, OUT, FXP>(in, width, height, stride, vtKern, hzKern, kernSize, out, borderType, batch, framePitch, as_stream(stream)); }

CVB_CONVLT_ENTRY(8u16s16s, uint8_t, int16_t, int16_t, false)
CVB_CONVLT_ENTRY(16s16s16s, int16_t, int16_t, int16_t, false)
CVB_CONVLT_ENTRY(8u32f8u, uint8_t, float, uint8_t, false)
CVB_CONVLT_ENTRY(8u32f32f, uint8_t, float, float, false)
CVB_CONVLT_ENTRY(32f32f32f, float, float, float, false)
CVB_CONVLT_ENTRY(32f32f8u, float, float, uint8_t, false)
CVB_CONVLT_ENTRY(fxp_8u16u8u, uint8_t, uint16_t, uint8_t, true)

// CompVMathGauss::kernelDim1<float> (base/include/compv/base/math/compv_math_gauss.h:23-56): same operation order and types
int cvb200_gauss_kernel_dim1_32f(size_t size, float sigma, float* kernel)
{
	CVB_REQUIRE(kernel && (size & 1), CVB200_E_INVALID_PARAMETER);
	const size_t size_div2 = size >> 1;
	const float sigma2_times2 = static_cast<float>(2 * (sigma * sigma));
	const float one_over = static_cast<float>(1 / sqrt(3.14159265358979323846 * sigma2_times2));
	float sum, k;
	kernel[size_div2] = one_over;
	sum = one_over;
	for (size_t x = 1; x <= size_div2; ++x) {
		k = static_cast<float>(one_over * exp(-static_cast<double>((x * x) / sigma2_times2)));
		kernel[x + size_div2] = k;
		kernel[size_div2 - x] = k;
		sum += (k + k);
	}
	sum = 1 / sum;
	for (size_t x = 0; x < size; ++x) kernel[x] *= sum;
	return CVB200_S_OK;
}

// CompVMathGauss::kernelDim1FixedPoint (base/math/compv_math_gauss.cxx) = kernelDim1<float> then CompVMathConvlt::fixedPointKernel
// (compv_math_convlt.h:76-92): k16 = uint16(k * 0xffff)
int cvb200_gauss_kernel_dim1_fxp(size_t size, float sigma, uint16_t* kernel)
{
	CVB_REQUIRE(kernel && (size & 1) && size <= 1024, CVB200_E_INVALID_PARAMETER);
	float tmp[1024];
	CVB_CHECK(cvb200_gauss_kernel_dim1_32f(size, sigma, tmp));
	for (size_t x = 0; x < size; ++x) kernel[x] = static_cast<uint16_t>(tmp[x] * 0xffff);
	return CVB200_S_OK;
}

} // extern "C"
