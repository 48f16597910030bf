// TEST INFRASTRUCTURE ONLY -- never linked or loaded by the product path (compv_b200/).
//
// extern "C" shim over the UNMODIFIED CompV reference library (oracle/_ref/libcompv_ref.so, built by
// oracle/build_ref.sh from the sources where they lie under /root/reference). It lets the Python tests and
// bench.py's CPU legs call the reference's own public C++ API through ctypes. It contains no algorithm of its
// own: every function below is a thin call into the reference API named in its comment.
#include "compv/base/compv_base.h"
#include "compv/base/compv_cpu.h"
#include "compv/base/compv_features.h"
#include "compv/base/compv_ccl.h"
#include "compv/base/image/compv_image.h"
#include "compv/base/math/compv_math_convlt.h"
#include "compv/base/math/compv_math_gauss.h"
#include "compv/base/math/compv_math_utils.h"
#include "compv/base/math/compv_math_morph.h"
#include "compv/base/compv_gradient_fast.h"
#include "compv/base/parallel/compv_parallel.h"
#include "compv/core/compv_core.h"

#include <chrono>
#include <cstring>
#include <string>
#include <vector>

using namespace compv;

#define SHIM_CHECK(x) do { COMPV_ERROR_CODE e__ = (x); if (COMPV_ERROR_CODE_IS_NOK(e__)) return static_cast<int>(e__); } while (0)

static double now_ms()
{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int wrap8u(const uint8_t* img, size_t w, size_t h, size_t stride, CompVMatPtr* out)
{
	// CompVImage::wrap copies into a CompVMat with the requested stride (base/image/compv_image.cxx:381-404)
	SHIM_CHECK(CompVImage::wrap(COMPV_SUBTYPE_PIXELS_Y, img, w, h, stride, &(*out), stride));
	return 0;
}

static void copy_rows(const CompVMatPtr& m, void* dst, size_t dstStrideBytes)
{
	const size_t rowBytes = m->rowInBytes();
	for (size_t j = 0; j < m->rows(); ++j) {
		memcpy(static_cast<uint8_t*>(dst) + j * dstStrideBytes, m->ptr<const uint8_t>(j), rowBytes);
	}
}

extern "C" {

// CompVBase::init + CompVCore::init (base/compv_base.cxx, core/compv_core.cxx:149-164). numThreads=-1 -> one per core.
int ref_init(int numThreads)
{
	CompVDebugMgr::setLevel(COMPV_DEBUG_LEVEL_ERROR);
	SHIM_CHECK(CompVBase::init(numThreads));
	SHIM_CHECK(CompVCore::init());
	return 0;
}

int ref_deinit()
{
	SHIM_CHECK(CompVCore::deInit());
	SHIM_CHECK(CompVBase::deInit());
	return 0;
}

int ref_set_max_threads(int n)
{
	SHIM_CHECK(CompVParallel::multiThreadingSetMaxThreads(n));
	return 0;
}

int ref_threads_count()
{
	CompVThreadDispatcherPtr d = CompVParallel::threadDispatcher();
	return d ? static_cast<int>(d->threadsCount()) : 1;
}

// CompVCpu::flagsDisable / flagsEnable: select the plain C++ path (all flags off) or the SIMD path
int ref_cpu_simd(int enable)
{
	if (enable) {
		SHIM_CHECK(CompVCpu::flagsEnable(kCpuFlagAll));
		SHIM_CHECK(CompVCpu::setIntrinsicsEnabled(true));
	}
	else {
		SHIM_CHECK(CompVCpu::flagsDisable(kCpuFlagAll));
		SHIM_CHECK(CompVCpu::setIntrinsicsEnabled(false));
	}
	return 0;
}

const char* ref_cpu_flags()
{
	static std::string s;
	s = CompVCpu::flagsAsString(CompVCpu::getFlags());
	return s.c_str();
}

// ---- K1..K3: CompVMathConvlt::convlt1<...> (base/include/compv/base/math/compv_math_convlt.h:25-55) ----
int ref_convlt1_8u16s16s(const uint8_t* in, size_t w, size_t h, size_t stride, const int16_t* vt, const int16_t* hz, size_t k, int16_t* out)
{
	SHIM_CHECK((CompVMathConvlt::convlt1<uint8_t, int16_t, int16_t>(in, w, h, stride, vt, hz, k, out)));
	return 0;
}
int ref_convlt1_16s16s16s(const int16_t* in, size_t w, size_t h, size_t stride, const int16_t* vt, const int16_t* hz, size_t k, int16_t* out)
{
	SHIM_CHECK((CompVMathConvlt::convlt1<int16_t, int16_t, int16_t>(in, w, h, stride, vt, hz, k, out)));
	return 0;
}
int ref_convlt1_8u32f8u(const uint8_t* in, size_t w, size_t h, size_t stride, const float* vt, const float* hz, size_t k, uint8_t* out)
{
	SHIM_CHECK((CompVMathConvlt::convlt1<uint8_t, compv_float32_t, uint8_t>(in, w, h, stride, vt, hz, k, out)));
	return 0;
}
int ref_convlt1_8u32f32f(const uint8_t* in, size_t w, size_t h, size_t stride, const float* vt, const float* hz, size_t k, float* out)
{
	SHIM_CHECK((CompVMathConvlt::convlt1<uint8_t, compv_float32_t, compv_float32_t>(in, w, h, stride, vt, hz, k, out)));
	return 0;
}
int ref_convlt1_32f32f32f(const float* in, size_t w, size_t h, size_t stride, const float* vt, const float* hz, size_t k, float* out)
{
	SHIM_CHECK((CompVMathConvlt::convlt1<compv_float32_t, compv_float32_t, compv_float32_t>(in, w, h, stride, vt, hz, k, out)));
	return 0;
}
int ref_convlt1_32f32f8u(const float* in, size_t w, size_t h, size_t stride, const float* vt, const float* hz, size_t k, uint8_t* out)
{
	SHIM_CHECK((CompVMathConvlt::convlt1<compv_float32_t, compv_float32_t, uint8_t>(in, w, h, stride, vt, hz, k, out)));
	return 0;
}
int ref_convlt1_fxp_8u16u8u(const uint8_t* in, size_t w, size_t h, size_t stride, const uint16_t* vt, const uint16_t* hz, size_t k, uint8_t* out)
{
	SHIM_CHECK(CompVMathConvlt::convlt1FixedPoint(in, w, h, stride, vt, hz, k, out));
	return 0;
}

// CompVMathGauss::kernelDim1<float> (base/include/compv/base/math/compv_math_gauss.h:23-56)
int ref_gauss_kernel_dim1_32f(size_t size, float sigma, float* out)
{
	CompVMatPtr kernel;
	SHIM_CHECK(CompVMathGauss::kernelDim1<compv_float32_t>(&kernel, size, sigma));
	memcpy(out, kernel->ptr<const compv_float32_t>(), size * sizeof(float));
	return 0;
}
// CompVMathGauss::kernelDim1FixedPoint (base/math/compv_math_gauss.cxx)
int ref_gauss_kernel_dim1_fxp(size_t size, float sigma, uint16_t* out)
{
	CompVMatPtr kernel;
	SHIM_CHECK(CompVMathGauss::kernelDim1FixedPoint(&kernel, size, sigma));
	memcpy(out, kernel->ptr<const uint16_t>(), size * sizeof(uint16_t));
	return 0;
}

// K4: CompVMathUtils::sumAbs<int16_t,uint16_t> (base/math/compv_math_utils.cxx:211-246)
int ref_sum_abs_16s16u(const int16_t* a, const int16_t* b, uint16_t* r, size_t w, size_t h, size_t stride)
{
	SHIM_CHECK((CompVMathUtils::sumAbs<int16_t, uint16_t>(a, b, r, w, h, stride)));
	return 0;
}

// ---- a3/a5: CompVEdgeDete (Sobel/Scharr/Prewitt/Canny) through the factory (base/compv_features.cxx:146-161) ----
// which: 0=Sobel 1=Scharr 2=Prewitt 3=Canny
// thresholdType: 0 = COMPV_CANNY_THRESHOLD_TYPE_COMPARE_TO_GRADIENT (default), 1 = ..._PERCENT_OF_MEAN (Canny only)
int ref_edge_dete(int which, const uint8_t* img, size_t w, size_t h, size_t stride, float tLow, float tHigh, int kernSize, int thresholdType, uint8_t* edges /* h*stride */)
{
	static const int ids[4] = { COMPV_SOBEL_ID, COMPV_SCHARR_ID, COMPV_PREWITT_ID, COMPV_CANNY_ID };
	CompVMatPtr image, out;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	CompVEdgeDetePtr dete;
	SHIM_CHECK(CompVEdgeDete::newObj(&dete, ids[which & 3], tLow, tHigh, static_cast<size_t>(kernSize)));
	if ((which & 3) == 3 && thresholdType == 1) {
		SHIM_CHECK(dete->setInt(COMPV_CANNY_SET_INT_THRESHOLD_TYPE, COMPV_CANNY_THRESHOLD_TYPE_PERCENT_OF_MEAN));
	}
	SHIM_CHECK(dete->process(image, &out));
	copy_rows(out, edges, stride);
	return 0;
}

// Timed variant of the above: object created once, 1 warm-up, `iters` timed iterations (steady_clock), per-iteration ms written to msOut[iters].
// Optional Gaussian pre-blur (blurSize>0): CompVMathGauss::kernelDim1<float> + CompVMathConvlt::convlt1<u8,f32,u8> in the timed loop (BASELINE config 2).
int ref_time_edge_dete(int which, const uint8_t* img, size_t w, size_t h, size_t stride, float tLow, float tHigh, int kernSize,
	int blurSize, float blurSigma, int iters, double* msOut, uint8_t* edges)
{
	static const int ids[4] = { COMPV_SOBEL_ID, COMPV_SCHARR_ID, COMPV_PREWITT_ID, COMPV_CANNY_ID };
	CompVMatPtr image, blurred, out, kernel;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	CompVEdgeDetePtr dete;
	SHIM_CHECK(CompVEdgeDete::newObj(&dete, ids[which & 3], tLow, tHigh, static_cast<size_t>(kernSize)));
	if (blurSize > 0) {
		SHIM_CHECK(CompVMathGauss::kernelDim1<compv_float32_t>(&kernel, static_cast<size_t>(blurSize), blurSigma));
		SHIM_CHECK(CompVImage::newObj8u(&blurred, COMPV_SUBTYPE_PIXELS_Y, w, h, stride));
	}
	for (int it = -1; it < iters; ++it) {
		const double t0 = now_ms();
		if (blurSize > 0) {
			uint8_t* bptr = blurred->ptr<uint8_t>();
			SHIM_CHECK((CompVMathConvlt::convlt1<uint8_t, compv_float32_t, uint8_t>(image->ptr<const uint8_t>(), w, h, stride,
				kernel->ptr<const compv_float32_t>(), kernel->ptr<const compv_float32_t>(), static_cast<size_t>(blurSize), bptr)));
			SHIM_CHECK(dete->process(blurred, &out));
		}
		else {
			SHIM_CHECK(dete->process(image, &out));
		}
		const double t1 = now_ms();
		if (it >= 0) msOut[it] = t1 - t0;
	}
	if (edges) copy_rows(out, edges, stride);
	return 0;
}

// ---- a8: CompVCornerDete (FAST) through the factory (base/compv_features.cxx:80-95; core/features/fast/compv_core_feature_fast_dete.cxx:163-422) ----
// fastType: 9 or 12. maxFeatures <= 1 disables selectBest. Points are copied out in the order the reference produced them.
// pts layout == CompVInterestPoint (x, y, strength, orient, level, size). iters > 0: timed loop, ms per iteration written to msOut (may be NULL).
int ref_fast_detect(const uint8_t* img, size_t w, size_t h, size_t stride, int fastType, int threshold, int nms, int maxFeatures,
	void* pts, size_t capacity, size_t* count, int iters, double* msOut)
{
	CompVMatPtr image;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	CompVCornerDetePtr dete;
	SHIM_CHECK(CompVCornerDete::newObj(&dete, COMPV_FAST_ID));
	SHIM_CHECK(dete->setInt(COMPV_FAST_SET_INT_THRESHOLD, threshold));
	SHIM_CHECK(dete->setInt(COMPV_FAST_SET_INT_FAST_TYPE, fastType == 12 ? COMPV_FAST_TYPE_12 : COMPV_FAST_TYPE_9));
	SHIM_CHECK(dete->setBool(COMPV_FAST_SET_BOOL_NON_MAXIMA_SUPP, nms != 0));
	SHIM_CHECK(dete->setInt(COMPV_FAST_SET_INT_MAX_FEATURES, maxFeatures));
	CompVInterestPointVector points;
	SHIM_CHECK(dete->process(image, points));
	for (int it = 0; it < iters; ++it) {
		const double t0 = now_ms();
		SHIM_CHECK(dete->process(image, points));
		if (msOut) msOut[it] = now_ms() - t0;
	}
	*count = points.size();
	static_assert(sizeof(CompVInterestPoint) == 24, "CompVInterestPoint layout");
	if (pts && capacity) memcpy(pts, points.data(), (points.size() < capacity ? points.size() : capacity) * sizeof(CompVInterestPoint));
	return 0;
}

// ---- section 8f-2: the ORB detector through the factory (core/features/orb/compv_core_feature_orb_dete.cxx:148-358), defaults except what is passed ----
int ref_orb_detect(const uint8_t* img, size_t w, size_t h, size_t stride, int fastThreshold, int nms, int maxFeatures, void* pts, size_t capacity, size_t* count, int iters, double* msOut)
{
	CompVMatPtr image;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	CompVCornerDetePtr dete;
	SHIM_CHECK(CompVCornerDete::newObj(&dete, COMPV_ORB_ID));
	SHIM_CHECK(dete->setInt(COMPV_ORB_SET_INT_FAST_THRESHOLD, fastThreshold));
	SHIM_CHECK(dete->setBool(COMPV_ORB_SET_BOOL_FAST_NON_MAXIMA_SUPP, nms != 0));
	SHIM_CHECK(dete->setInt(COMPV_ORB_SET_INT_MAX_FEATURES, maxFeatures));
	CompVInterestPointVector points;
	SHIM_CHECK(dete->process(image, points));
	for (int it = 0; it < iters; ++it) {
		const double t0 = now_ms();
		SHIM_CHECK(dete->process(image, points));
		if (msOut) msOut[it] = now_ms() - t0;
	}
	*count = points.size();
	if (pts && capacity) memcpy(pts, points.data(), (points.size() < capacity ? points.size() : capacity) * sizeof(CompVInterestPoint));
	return 0;
}

// Reference defect probe: ONE CompVCornerDeteFAST object run on image A and then on image B (what CompVCornerDeteORB does across pyramid levels, orb_dete.cxx:239-246).
// The detector keeps its strengths / NMS maps while the image STRIDE does not change (fast_dete.cxx:186-197) and only rewrites the positions of the new, smaller image,
// so B's result can contain corners left over from A.  Returns B's point count through *count (and the points).
int ref_fast_detect_after(const uint8_t* imgA, size_t wA, size_t hA, const uint8_t* imgB, size_t wB, size_t hB, void* pts, size_t capacity, size_t* count)
{
	CompVMatPtr a, b;
	SHIM_CHECK(CompVImage::wrap(COMPV_SUBTYPE_PIXELS_Y, imgA, wA, hA, wA, &a));
	SHIM_CHECK(CompVImage::wrap(COMPV_SUBTYPE_PIXELS_Y, imgB, wB, hB, wB, &b));
	CompVCornerDetePtr dete;
	SHIM_CHECK(CompVCornerDete::newObj(&dete, COMPV_FAST_ID));
	SHIM_CHECK(dete->setInt(COMPV_FAST_SET_INT_MAX_FEATURES, -1));
	CompVInterestPointVector points;
	SHIM_CHECK(dete->process(a, points));
	SHIM_CHECK(dete->process(b, points));
	*count = points.size();
	if (pts && capacity) memcpy(pts, points.data(), (points.size() < capacity ? points.size() : capacity) * sizeof(CompVInterestPoint));
	return (a->stride() == b->stride()) ? 0 : -7; // -7: the two images did not share a stride, the maps were re-allocated, nothing to see
}

// CompVImage::scale, bilinear (base/image/compv_image.cxx:840-905, base/image/compv_image_scale_bilinear.cxx)
int ref_scale_bilinear(const uint8_t* img, size_t w, size_t h, size_t stride, uint8_t* out, size_t outW, size_t outH)
{
	CompVMatPtr image, scaled;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	SHIM_CHECK(CompVImage::scale(image, &scaled, outW, outH, COMPV_INTERPOLATION_TYPE_BILINEAR));
	for (size_t j = 0; j < outH; ++j) memcpy(out + j * outW, scaled->ptr<const uint8_t>(j), outW);
	return 0;
}

// ---- a6/a7: CompVHough (SHT / KHT) through the factory (base/compv_features.cxx:176-191) ----
// which: 0 = COMPV_HOUGHSHT_ID, 1 = COMPV_HOUGHKHT_ID. theta is what CompVHough::newObj receives (KHT: degrees; SHT: see houghsht.cxx).
// lines layout == CompVHoughLine {float rho; float theta; size_t strength}. iters > 0 adds a timed loop (ms per iteration in msOut, may be NULL).
int ref_hough(int which, const uint8_t* edges, size_t w, size_t h, size_t stride, float rho, float theta, size_t threshold, int maxLines,
	float khtClusterMinDeviation, int khtClusterMinSize, float khtKernelMinHeight,
	void* lines, size_t capacity, size_t* count, double* gsOut, int iters, double* msOut)
{
	CompVMatPtr image;
	int r = wrap8u(edges, w, h, stride, &image);
	if (r) return r;
	CompVHoughPtr hough;
	SHIM_CHECK(CompVHough::newObj(&hough, which ? COMPV_HOUGHKHT_ID : COMPV_HOUGHSHT_ID, rho, theta, threshold));
	if (maxLines > 0) SHIM_CHECK(hough->setInt(COMPV_HOUGH_SET_INT_MAXLINES, maxLines));
	if (which) {
		SHIM_CHECK(hough->setFloat32(COMPV_HOUGHKHT_SET_FLT32_CLUSTER_MIN_DEVIATION, khtClusterMinDeviation));
		SHIM_CHECK(hough->setInt(COMPV_HOUGHKHT_SET_INT_CLUSTER_MIN_SIZE, khtClusterMinSize));
		SHIM_CHECK(hough->setFloat32(COMPV_HOUGHKHT_SET_FLT32_KERNEL_MIN_HEIGTH, khtKernelMinHeight));
	}
	CompVHoughLineVector out;
	SHIM_CHECK(hough->process(image, out));
	for (int it = 0; it < iters; ++it) {
		const double t0 = now_ms();
		SHIM_CHECK(hough->process(image, out));
		if (msOut) msOut[it] = now_ms() - t0;
	}
	if (which && gsOut) {
		compv_float64_t gs = 0;
		SHIM_CHECK(hough->getFloat64(COMPV_HOUGHKHT_GET_FLT64_GS, &gs));
		*gsOut = gs;
	}
	*count = out.size();
	static_assert(sizeof(CompVHoughLine) == 16, "CompVHoughLine layout");
	if (lines && capacity) memcpy(lines, out.data(), (out.size() < capacity ? out.size() : capacity) * sizeof(CompVHoughLine));
	return 0;
}

// ---- a10: CompVImage::thresholdOtsu / thresholdGlobal / thresholdAdaptive (base/include/compv/base/image/compv_image.h:63-67) ----
// mode 0: global(threshold) ; 1: otsu (threshold written to *thrOut) ; 2: adaptive(blockSize, delta, maxVal, invert)
int ref_threshold(int mode, const uint8_t* img, size_t w, size_t h, size_t stride, double threshold, size_t blockSize, double delta, double maxVal, int invert,
	double* thrOut, uint8_t* outPtr, int iters, double* msOut)
{
	CompVMatPtr image, out;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	for (int it = -1; it < iters; ++it) {
		const double t0 = now_ms();
		if (mode == 0) SHIM_CHECK(CompVImage::thresholdGlobal(image, &out, threshold));
		else if (mode == 1) { double t = 0; SHIM_CHECK(CompVImage::thresholdOtsu(image, t, &out)); if (thrOut) *thrOut = t; }
		else SHIM_CHECK(CompVImage::thresholdAdaptive(image, &out, blockSize, delta, maxVal, invert != 0));
		if (it >= 0 && msOut) msOut[it] = now_ms() - t0;
	}
	if (outPtr) copy_rows(out, outPtr, stride);
	return 0;
}

// ---- a4: CompVImage::gradientX/Y (base/image/compv_image.cxx:710-730) + CompVGradientFast::magnitude/direction ----
int ref_gradient_fast(const uint8_t* img, size_t w, size_t h, size_t stride, int16_t* gx16, int16_t* gy16, float* gx32, float* gy32, float* mag, float* dir)
{
	CompVMatPtr image, a, b, c, d, m, dd;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	SHIM_CHECK(CompVImage::gradientX(image, &a, false));
	SHIM_CHECK(CompVImage::gradientY(image, &b, false));
	SHIM_CHECK(CompVImage::gradientX(image, &c, true));
	SHIM_CHECK(CompVImage::gradientY(image, &d, true));
	SHIM_CHECK(CompVGradientFast::magnitude(c, d, &m));
	SHIM_CHECK(CompVGradientFast::direction(c, d, &dd, true));
	if (gx16) copy_rows(a, gx16, stride * 2);
	if (gy16) copy_rows(b, gy16, stride * 2);
	if (gx32) copy_rows(c, gx32, stride * 4);
	if (gy32) copy_rows(d, gy32, stride * 4);
	if (mag) copy_rows(m, mag, stride * 4);
	if (dir) copy_rows(dd, dir, stride * 4);
	return 0;
}

// ---- a9: CompVHOG through the factory (base/compv_features.cxx:210-235; core/features/hog/compv_core_feature_hog_std.cxx:196-393) ----
int ref_hog(const uint8_t* img, size_t w, size_t h, size_t stride, size_t bw, size_t bh, size_t sw, size_t sh, size_t cw, size_t ch, size_t nbins, int blockNorm, int gradientSigned, int interp,
	float* out, size_t capacity, size_t* size, int iters, double* msOut)
{
	CompVMatPtr image, desc;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	CompVHOGPtr hog;
	SHIM_CHECK(CompVHOG::newObj(&hog, COMPV_HOGS_ID, CompVSizeSz(bw, bh), CompVSizeSz(sw, sh), CompVSizeSz(cw, ch), nbins, blockNorm, gradientSigned != 0, interp));
	for (int it = -1; it < iters; ++it) {
		const double t0 = now_ms();
		SHIM_CHECK(hog->process(image, &desc));
		if (it >= 0 && msOut) msOut[it] = now_ms() - t0;
	}
	*size = desc->cols();
	if (out) memcpy(out, desc->ptr<const float>(), (desc->cols() < capacity ? desc->cols() : capacity) * sizeof(float));
	return 0;
}

// Persistent edge-detection session for bench.py's CPU legs: frames are wrapped once (CompVImage::wrap), the detector, the Gaussian kernel and the
// output matrices are created once, then ref_edge_session_run() times CompVMathConvlt::convlt1<u8,f32,u8> (optional) + CompVEdgeDete::process per frame.
struct RefEdgeSession {
	std::vector<CompVMatPtr> frames;
	CompVMatPtr blurred, out, kernel;
	CompVEdgeDetePtr dete;
	CompVHoughPtr hough;        // optional: KHT on the edge map (BASELINE metric "Canny+HoughKHT")
	CompVHoughLineVector lines;
	size_t linesTotal;
	int blurSize;
};

void* ref_edge_session_new(int which, const uint8_t* frames, size_t count, size_t w, size_t h, size_t stride, float tLow, float tHigh, int kernSize, int blurSize, float blurSigma)
{
	static const int ids[4] = { COMPV_SOBEL_ID, COMPV_SCHARR_ID, COMPV_PREWITT_ID, COMPV_CANNY_ID };
	RefEdgeSession* s = new RefEdgeSession();
	s->blurSize = blurSize;
	s->linesTotal = 0;
	for (size_t i = 0; i < count; ++i) {
		CompVMatPtr m;
		if (wrap8u(frames + i * stride * h, w, h, stride, &m)) { delete s; return NULL; }
		s->frames.push_back(m);
	}
	if (COMPV_ERROR_CODE_IS_NOK(CompVEdgeDete::newObj(&s->dete, ids[which & 3], tLow, tHigh, static_cast<size_t>(kernSize)))) { delete s; return NULL; }
	if (blurSize > 0) {
		if (COMPV_ERROR_CODE_IS_NOK(CompVMathGauss::kernelDim1<compv_float32_t>(&s->kernel, static_cast<size_t>(blurSize), blurSigma))
			|| COMPV_ERROR_CODE_IS_NOK(CompVImage::newObj8u(&s->blurred, COMPV_SUBTYPE_PIXELS_Y, w, h, stride))) { delete s; return NULL; }
	}
	return s;
}

// Adds CompVHough (KHT, rho 1, theta 1 degree) after the edge detector of the session; returns 0 on success.
int ref_edge_session_add_kht(void* session, size_t threshold)
{
	RefEdgeSession* s = static_cast<RefEdgeSession*>(session);
	if (!s) return -1;
	SHIM_CHECK(CompVHough::newObj(&s->hough, COMPV_HOUGHKHT_ID, 1.f, 1.f, threshold));
	return 0;
}

size_t ref_edge_session_lines_total(void* session) { return session ? static_cast<RefEdgeSession*>(session)->linesTotal : 0; }

// Processes frames[first .. first+count) (indices wrap around); returns elapsed milliseconds (steady_clock) or a negative error code.
double ref_edge_session_run(void* session, size_t first, size_t count, uint8_t* lastEdges, size_t stride)
{
	RefEdgeSession* s = static_cast<RefEdgeSession*>(session);
	if (!s || s->frames.empty()) return -1.0;
	const double t0 = now_ms();
	for (size_t i = 0; i < count; ++i) {
		const CompVMatPtr& image = s->frames[(first + i) % s->frames.size()];
		if (s->blurSize > 0) {
			uint8_t* bptr = s->blurred->ptr<uint8_t>();
			if (COMPV_ERROR_CODE_IS_NOK((CompVMathConvlt::convlt1<uint8_t, compv_float32_t, uint8_t>(image->ptr<const uint8_t>(), image->cols(), image->rows(), image->stride(),
				s->kernel->ptr<const compv_float32_t>(), s->kernel->ptr<const compv_float32_t>(), static_cast<size_t>(s->blurSize), bptr)))) return -2.0;
			if (COMPV_ERROR_CODE_IS_NOK(s->dete->process(s->blurred, &s->out))) return -3.0;
		}
		else {
			if (COMPV_ERROR_CODE_IS_NOK(s->dete->process(image, &s->out))) return -3.0;
		}
		if (s->hough) {
			if (COMPV_ERROR_CODE_IS_NOK(s->hough->process(s->out, s->lines))) return -4.0;
			s->linesTotal += s->lines.size();
		}
	}
	const double t1 = now_ms();
	if (lastEdges && s->out) copy_rows(s->out, lastEdges, stride);
	return t1 - t0;
}

void ref_edge_session_free(void* session)
{
	delete static_cast<RefEdgeSession*>(session);
}

// ---- a11: CompVConnectedComponentLabeling (COMPV_PLSL_ID) through the factory (base/compv_ccl.cxx:69-97) ----
// labels: height x width int32 (strideless, CompVConnectedComponentLabelingResultLSL::debugFlatten); boxes: labelsCount x {left, top, right, bottom} int16
// (CompVRectInt16, boundingBoxes()). boxCap in boxes. iters > 0 adds a timed loop over process() only.
int ref_ccl_lsl(const uint8_t* img, size_t w, size_t h, size_t stride, int32_t* labels, int32_t* naOut, int16_t* boxes, size_t boxCap, int iters, double* msOut)
{
	CompVMatPtr image;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	CompVConnectedComponentLabelingPtr ccl;
	SHIM_CHECK(CompVConnectedComponentLabeling::newObj(&ccl, COMPV_PLSL_ID));
	CompVConnectedComponentLabelingResultPtr result;
	SHIM_CHECK(ccl->process(image, &result));
	for (int it = 0; it < iters; ++it) {
		const double t0 = now_ms();
		SHIM_CHECK(ccl->process(image, &result));
		if (msOut) msOut[it] = now_ms() - t0;
	}
	const CompVConnectedComponentLabelingResultLSL* lsl = CompVConnectedComponentLabeling::reinterpret_castr<CompVConnectedComponentLabelingResultLSL>(result);
	if (!lsl) return -2;
	if (naOut) *naOut = static_cast<int32_t>(lsl->labelsCount());
	if (labels) {
		if (lsl->labelsCount()) {
			CompVMatPtr flat;
			SHIM_CHECK(lsl->debugFlatten(&flat));
			for (size_t j = 0; j < h; ++j) memcpy(labels + j * w, flat->ptr<const int32_t>(j), w * sizeof(int32_t));
		}
		else memset(labels, 0, w * h * sizeof(int32_t)); // black image: the result is empty (ccl_lsl.cxx:676-680), debugFlatten refuses it
	}
	if (boxes) {
		CompVConnectedComponentBoundingBoxesVector bb;
		SHIM_CHECK(lsl->boundingBoxes(bb));
		static_assert(sizeof(CompVConnectedComponentBoundingBox) == 8, "CompVRectInt16 layout");
		memcpy(boxes, bb.data(), (bb.size() < boxCap ? bb.size() : boxCap) * sizeof(CompVConnectedComponentBoundingBox));
	}
	return 0;
}

// ---- section 8f-3: CompVImage::wrap + CompVImage::convertGrayscale (base/image/compv_image.cxx:381-404, 687-692) ----
// data: a whole frame in `subtype` layout with `stride` samples per row (planar formats: planes back to back as CompVImage lays them out); gray out with the image's own Y stride.
int ref_to_grayscale(int subtype, const uint8_t* data, size_t w, size_t h, size_t stride, uint8_t* gray, size_t grayStride)
{
	CompVMatPtr image, out;
	SHIM_CHECK(CompVImage::wrap(static_cast<COMPV_SUBTYPE>(subtype), data, w, h, stride, &image, stride));
	SHIM_CHECK(CompVImage::convertGrayscale(image, &out));
	if (out->cols() != w || out->rows() != h) return -2;
	for (size_t j = 0; j < h; ++j) memcpy(gray + j * grayStride, out->ptr<const uint8_t>(j), w);
	return 0;
}

// ---- section 8f-4: the consumers of the PLSL / Hough results ----
// CompVConnectedComponentLabelingResultLSL::extract (core/ccl/compv_core_ccl_lsl_result.cxx:100-134, 308-416): counts[label] points per label, points (x, y) concatenated in label order.
// type: 0 = COMPV_CCL_EXTRACT_TYPE_SEGMENT, 1 = COMPV_CCL_EXTRACT_TYPE_BLOB.  Also the boxes computed FROM the extracted segments (ccl_lsl_result.cxx:187-230) when boxesFromSegments != NULL.
int ref_ccl_lsl_extract(const uint8_t* img, size_t w, size_t h, size_t stride, int type, int32_t* counts, size_t countCap, int16_t* points, size_t pointCap, size_t* nLabels, size_t* nPoints,
	int16_t* boxesFromSegments)
{
	CompVMatPtr image;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	CompVConnectedComponentLabelingPtr ccl;
	SHIM_CHECK(CompVConnectedComponentLabeling::newObj(&ccl, COMPV_PLSL_ID));
	CompVConnectedComponentLabelingResultPtr result;
	SHIM_CHECK(ccl->process(image, &result));
	const CompVConnectedComponentLabelingResultLSL* lsl = CompVConnectedComponentLabeling::reinterpret_castr<CompVConnectedComponentLabelingResultLSL>(result);
	if (!lsl) return -2;
	CompVConnectedComponentPointsVector pts;
	SHIM_CHECK(lsl->extract(pts, type ? COMPV_CCL_EXTRACT_TYPE_BLOB : COMPV_CCL_EXTRACT_TYPE_SEGMENT));
	size_t np = 0;
	for (size_t a = 0; a < pts.size(); ++a) {
		if (counts && a < countCap) counts[a] = static_cast<int32_t>(pts[a].size());
		for (size_t k = 0; k < pts[a].size(); ++k, ++np) if (points && np < pointCap) { points[2 * np] = pts[a][k].x; points[2 * np + 1] = pts[a][k].y; }
	}
	if (nLabels) *nLabels = pts.size();
	if (nPoints) *nPoints = np;
	if (boxesFromSegments && !type) {
		CompVConnectedComponentBoundingBoxesVector bb;
		SHIM_CHECK(lsl->boundingBoxes(pts, bb));
		memcpy(boxesFromSegments, bb.data(), (bb.size() < countCap ? bb.size() : countCap) * sizeof(CompVConnectedComponentBoundingBox));
	}
	return 0;
}

// CompVHough::toCartesian (houghkht.cxx:1249-1280, houghsht.cxx:566-592): out = n x {a.x, a.y, b.x, b.y}
int ref_hough_to_cartesian(int kht, size_t w, size_t h, size_t n, const float* rho, const float* theta, float* out)
{
	CompVHoughPtr hough;
	SHIM_CHECK(CompVHough::newObj(&hough, kht ? COMPV_HOUGHKHT_ID : COMPV_HOUGHSHT_ID, 1.f, 1.f, 1));
	CompVHoughLineVector polar(n);
	for (size_t i = 0; i < n; ++i) { polar[i].rho = rho[i]; polar[i].theta = theta[i]; polar[i].strength = 1; }
	CompVLineFloat32Vector cart;
	SHIM_CHECK(hough->toCartesian(w, h, polar, cart));
	if (cart.size() != n) return -2;
	for (size_t i = 0; i < n; ++i) { out[4 * i] = cart[i].a.x; out[4 * i + 1] = cart[i].a.y; out[4 * i + 2] = cart[i].b.x; out[4 * i + 3] = cart[i].b.y; }
	return 0;
}

// ---- a12: CompVConnectedComponentLabeling (COMPV_LMSER_ID) ----
// regions are returned in the reference's order: regionSizes[i] = number of points, regionBoxes[4*i..] = {left, top, right, bottom},
// points (int16 x, y pairs) concatenated in region order up to pointCap points. *regionCount / *pointCount receive the totals.
int ref_ccl_lmser(const uint8_t* img, size_t w, size_t h, size_t stride, int delta, double minArea, double maxArea, double maxVariation, double minDiversity, int connectivity,
	int32_t* regionSizes, int16_t* regionBoxes, size_t regionCap, int16_t* points, size_t pointCap, size_t* regionCount, size_t* pointCount, int iters, double* msOut)
{
	CompVMatPtr image;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	CompVConnectedComponentLabelingPtr ccl;
	SHIM_CHECK(CompVConnectedComponentLabeling::newObj(&ccl, COMPV_LMSER_ID, delta, minArea, maxArea, maxVariation, minDiversity, connectivity));
	CompVConnectedComponentLabelingResultPtr result;
	SHIM_CHECK(ccl->process(image, &result));
	for (int it = 0; it < iters; ++it) {
		const double t0 = now_ms();
		SHIM_CHECK(ccl->process(image, &result));
		if (msOut) msOut[it] = now_ms() - t0;
	}
	const CompVConnectedComponentLabelingResultLMSER* mser = CompVConnectedComponentLabeling::reinterpret_castr<CompVConnectedComponentLabelingResultLMSER>(result);
	if (!mser) return -2;
	const CompVConnectedComponentLabelingRegionMserVector& regions = mser->boundingBoxes(); // same vector as points(), with the boxes filled (lmser_result.cxx:50-88)
	size_t np = 0;
	for (size_t i = 0; i < regions.size(); ++i) {
		const CompVConnectedComponentLabelingRegionMser& reg = regions[i];
		if (i < regionCap) {
			if (regionSizes) regionSizes[i] = static_cast<int32_t>(reg.points.size());
			if (regionBoxes) { regionBoxes[4 * i] = reg.boundingBox.left; regionBoxes[4 * i + 1] = reg.boundingBox.top; regionBoxes[4 * i + 2] = reg.boundingBox.right; regionBoxes[4 * i + 3] = reg.boundingBox.bottom; }
		}
		for (size_t k = 0; k < reg.points.size(); ++k, ++np) {
			if (points && np < pointCap) { points[2 * np] = reg.points[k].x; points[2 * np + 1] = reg.points[k].y; }
		}
	}
	if (regionCount) *regionCount = regions.size();
	if (pointCount) *pointCount = np;
	return 0;
}

// ---- section 8f "next" row 1: CompVMathMorph (base/math/compv_math_morph.cxx:85-126) ----
// strelType >= 0: CompVMathMorph::buildStructuringElement(size sw x sh, type); strelType < 0: `strel` (sh x sw, packed) is used as is.
// out is pre-filled by the caller (border IGNORE leaves cells untouched); op / border are the reference's enum values.
int ref_morph(const uint8_t* img, size_t w, size_t h, size_t stride, int strelType, const uint8_t* strel, size_t sw, size_t sh, uint8_t* strelOut, int op, int border, uint8_t* outPtr,
	int iters, double* msOut)
{
	CompVMatPtr image, se, out;
	int r = wrap8u(img, w, h, stride, &image);
	if (r) return r;
	if (strelType >= 0) SHIM_CHECK(CompVMathMorph::buildStructuringElement(&se, CompVSizeSz(sw, sh), static_cast<COMPV_MATH_MORPH_STREL_TYPE>(strelType)));
	else {
		SHIM_CHECK(CompVMat::newObjAligned<uint8_t>(&se, sh, sw));
		for (size_t j = 0; j < sh; ++j) memcpy(se->ptr<uint8_t>(j), strel + j * sw, sw);
	}
	if (strelOut) for (size_t j = 0; j < sh; ++j) memcpy(strelOut + j * sw, se->ptr<const uint8_t>(j), sw);
	if (!outPtr) return 0;
	r = wrap8u(outPtr, w, h, stride, &out); // pre-filled output with the input's geometry: reused by basicOper (:144-148)
	if (r) return r;
	for (int it = -1; it < iters; ++it) {
		const double t0 = now_ms();
		SHIM_CHECK(CompVMathMorph::process(image, se, &out, static_cast<COMPV_MATH_MORPH_OP_TYPE>(op), static_cast<COMPV_BORDER_TYPE>(border)));
		if (it >= 0 && msOut) msOut[it] = now_ms() - t0;
	}
	copy_rows(out, outPtr, stride);
	return 0;
}

} // extern "C"
