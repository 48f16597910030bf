// Host-side C++ mirror of the reference's operator interface for the hot path, over the C ABI of cvb200.h.
// Same class names, factory signatures, argument meaning and error codes as the reference, so that code written against CompV reads the same:
//
//     CompVEdgeDetePtr dete;   COMPV_CHECK_CODE_RETURN(CompVEdgeDete::newObj(&dete, COMPV_CANNY_ID, 59.f, 119.f));
//     CompVMatPtr edges;       COMPV_CHECK_CODE_RETURN(dete->process(image, &edges));
//     CompVHoughPtr hough;     COMPV_CHECK_CODE_RETURN(CompVHough::newObj(&hough, COMPV_HOUGHKHT_ID, 1.f, 1.f, 100));
//     CompVHoughLineVector lines; COMPV_CHECK_CODE_RETURN(hough->process(edges, lines));          (samples/hough_lines/main.cxx:59-106)
//
// Mirrors: CompVMat (base/include/compv/base/compv_mat.h:21-588, the single-plane subset the path uses), CompVCaps (compv_caps.h), CompVEdgeDete / CompVCornerDete /
// CompVHough / CompVHOG (base/include/compv/base/compv_features.h:160-240), CompVConnectedComponentLabeling + results (base/include/compv/base/compv_ccl.h:105-241),
// CompVImage::threshold* (base/include/compv/base/image/compv_image.h:63-67), CompVMathConvlt::convlt1 (base/include/compv/base/math/compv_math_convlt.h:25-55).
// Header only; link with -lcompv_b200.  Everything computes on the GPU: there is no CPU path behind these classes (calls fail with E_NOT_INITIALIZED / E_CUDA).
// Host-side helpers of the results are here too (CompVHough::toCartesian, CompVConnectedComponentLabelingResult::extract, CompVMathMorph).
// Not mirrored: the threading runtime, image formats other than 8-bit gray, ORB / pyramids.
#pragma once

#include "cvb200.h"

#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

namespace compv {

typedef int COMPV_ERROR_CODE; // numeric values of the reference's enum (compv_common.h), see cvb200.h
enum {
	COMPV_ERROR_CODE_S_OK = CVB200_S_OK,
	COMPV_ERROR_CODE_E_NOT_IMPLEMENTED = CVB200_E_NOT_IMPLEMENTED,
	COMPV_ERROR_CODE_E_NOT_INITIALIZED = CVB200_E_NOT_INITIALIZED,
	COMPV_ERROR_CODE_E_INVALID_STATE = CVB200_E_INVALID_STATE,
	COMPV_ERROR_CODE_E_INVALID_PARAMETER = CVB200_E_INVALID_PARAMETER,
	COMPV_ERROR_CODE_E_OUT_OF_MEMORY = CVB200_E_OUT_OF_MEMORY,
	COMPV_ERROR_CODE_E_OUT_OF_BOUND = CVB200_E_OUT_OF_BOUND,
	COMPV_ERROR_CODE_E_CUDA = CVB200_E_CUDA,
};
#define COMPV_ERROR_CODE_IS_OK(code) ((code) == ::compv::COMPV_ERROR_CODE_S_OK)
#define COMPV_ERROR_CODE_IS_NOK(code) ((code) != ::compv::COMPV_ERROR_CODE_S_OK)
#define COMPV_CHECK_CODE_RETURN(expr) do { const ::compv::COMPV_ERROR_CODE compv_rc_ = (expr); if (compv_rc_ != ::compv::COMPV_ERROR_CODE_S_OK) return compv_rc_; } while (0)
#define COMPV_CHECK_EXP_RETURN(exp, code) do { if ((exp)) return (code); } while (0)

// feature / capability ids (compv_features.h:47-121, compv_ccl.h:63-103)
enum {
	COMPV_FAST_ID = CVB200_FAST_ID, COMPV_FAST_SET_INT_THRESHOLD = CVB200_FAST_SET_INT_THRESHOLD, COMPV_FAST_SET_INT_MAX_FEATURES = CVB200_FAST_SET_INT_MAX_FEATURES,
	COMPV_FAST_SET_INT_FAST_TYPE = CVB200_FAST_SET_INT_FAST_TYPE, COMPV_FAST_SET_BOOL_NON_MAXIMA_SUPP = CVB200_FAST_SET_BOOL_NON_MAXIMA_SUPP,
	COMPV_FAST_TYPE_9 = CVB200_FAST_TYPE_9, COMPV_FAST_TYPE_12 = CVB200_FAST_TYPE_12,
	COMPV_CANNY_ID = CVB200_CANNY_ID, COMPV_CANNY_SET_FLT32_THRESHOLD_LOW = CVB200_CANNY_SET_FLT32_THRESHOLD_LOW, COMPV_CANNY_SET_FLT32_THRESHOLD_HIGH = CVB200_CANNY_SET_FLT32_THRESHOLD_HIGH,
	COMPV_CANNY_SET_INT_KERNEL_SIZE = CVB200_CANNY_SET_INT_KERNEL_SIZE,
	COMPV_SOBEL_ID = CVB200_SOBEL_ID, COMPV_SCHARR_ID = CVB200_SCHARR_ID, COMPV_PREWITT_ID = CVB200_PREWITT_ID,
	COMPV_HOUGHSHT_ID = CVB200_HOUGHSHT_ID, COMPV_HOUGHKHT_ID = CVB200_HOUGHKHT_ID,
	COMPV_HOUGH_SET_FLT32_RHO = CVB200_HOUGH_SET_FLT32_RHO, COMPV_HOUGH_SET_FLT32_THETA = CVB200_HOUGH_SET_FLT32_THETA,
	COMPV_HOUGH_SET_INT_THRESHOLD = CVB200_HOUGH_SET_INT_THRESHOLD, COMPV_HOUGH_SET_INT_MAXLINES = CVB200_HOUGH_SET_INT_MAXLINES,
	COMPV_HOUGHKHT_SET_FLT32_CLUSTER_MIN_DEVIATION = CVB200_HOUGHKHT_SET_FLT32_CLUSTER_MIN_DEVIATION, COMPV_HOUGHKHT_SET_INT_CLUSTER_MIN_SIZE = CVB200_HOUGHKHT_SET_INT_CLUSTER_MIN_SIZE,
	COMPV_HOUGHKHT_SET_FLT32_KERNEL_MIN_HEIGTH = CVB200_HOUGHKHT_SET_FLT32_KERNEL_MIN_HEIGTH, COMPV_HOUGHKHT_GET_FLT64_GS = CVB200_HOUGHKHT_GET_FLT64_GS,
	COMPV_HOGS_ID = CVB200_HOGS_ID, COMPV_HOG_BLOCK_NORM_NONE = CVB200_HOG_BLOCK_NORM_NONE, COMPV_HOG_BLOCK_NORM_L1 = CVB200_HOG_BLOCK_NORM_L1,
	COMPV_HOG_BLOCK_NORM_L1SQRT = CVB200_HOG_BLOCK_NORM_L1SQRT, COMPV_HOG_BLOCK_NORM_L2 = CVB200_HOG_BLOCK_NORM_L2, COMPV_HOG_BLOCK_NORM_L2HYS = CVB200_HOG_BLOCK_NORM_L2HYS,
	COMPV_HOG_INTERPOLATION_NEAREST = CVB200_HOG_INTERPOLATION_NEAREST, COMPV_HOG_INTERPOLATION_BILINEAR = CVB200_HOG_INTERPOLATION_BILINEAR,
	COMPV_HOG_INTERPOLATION_BILINEAR_LUT = CVB200_HOG_INTERPOLATION_BILINEAR_LUT,
};
enum { // separate id space (compv_ccl.h)
	COMPV_CCL_SET_INT_CONNECTIVITY = CVB200_CCL_SET_INT_CONNECTIVITY, COMPV_PLSL_ID = CVB200_PLSL_ID, COMPV_PLSL_SET_INT_TYPE = CVB200_PLSL_SET_INT_TYPE,
	COMPV_PLSL_SET_BOOL_SORT_SEGMENTS = CVB200_PLSL_SET_BOOL_SORT_SEGMENTS, COMPV_PLSL_TYPE_XRLEZ = CVB200_PLSL_TYPE_XRLEZ, COMPV_LMSER_ID = CVB200_LMSER_ID,
};
enum COMPV_BORDER_TYPE { COMPV_BORDER_TYPE_ZERO = CVB200_BORDER_TYPE_ZERO, COMPV_BORDER_TYPE_IGNORE = CVB200_BORDER_TYPE_IGNORE, COMPV_BORDER_TYPE_REPLICATE = CVB200_BORDER_TYPE_REPLICATE };

// ---- CompVBase / CompVGpu: initialisation (base/compv_base.cxx, gpu/compv_gpu.cxx:36-62) ----
struct CompVBase {
	static COMPV_ERROR_CODE init(int device = 0) { return cvb200_init(device); }
	static COMPV_ERROR_CODE deInit() { return cvb200_deinit(); }
	static bool isInitialized() { return cvb200_is_active() != 0; }
};

// ---- CompVMat: one plane, row-major, stride in samples, 64-byte aligned rows (compv_mat.h; CompVMem alignment) ----
class CompVMat;
typedef std::shared_ptr<CompVMat> CompVMatPtr;
typedef CompVMatPtr* CompVMatPtrPtr;
class CompVMat {
public:
	~CompVMat() { if (m_own) free(m_ptr); }
	template <typename T>
	static COMPV_ERROR_CODE newObj(CompVMatPtrPtr mat, size_t rows, size_t cols, size_t stride = 0) {
		COMPV_CHECK_EXP_RETURN(!mat || !rows || !cols || (stride && stride < cols), COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		const size_t alignElts = 64 / sizeof(T);
		const size_t s = stride ? stride : ((cols + alignElts - 1) / alignElts) * alignElts;
		if (*mat && (*mat)->m_own && (*mat)->m_rows == rows && (*mat)->m_cols == cols && (*mat)->m_stride == s && (*mat)->m_elmt == sizeof(T)) return COMPV_ERROR_CODE_S_OK; // reuse, like CompVMat::newObj
		void* p = NULL;
		COMPV_CHECK_EXP_RETURN(posix_memalign(&p, 64, rows * s * sizeof(T) + 64) != 0, COMPV_ERROR_CODE_E_OUT_OF_MEMORY);
		CompVMatPtr m(new CompVMat());
		m->m_ptr = p; m->m_own = true; m->m_rows = rows; m->m_cols = cols; m->m_stride = s; m->m_elmt = sizeof(T);
		*mat = m;
		return COMPV_ERROR_CODE_S_OK;
	}
	// CompVImage::wrap-like: copies rows of a caller buffer into an aligned CompVMat (base/image/compv_image.cxx:381-404)
	static COMPV_ERROR_CODE wrap8u(CompVMatPtrPtr mat, const uint8_t* data, size_t width, size_t height, size_t dataStride) {
		COMPV_CHECK_EXP_RETURN(!data || dataStride < width, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		COMPV_CHECK_CODE_RETURN(newObj<uint8_t>(mat, height, width));
		for (size_t j = 0; j < height; ++j) memcpy((*mat)->ptr<uint8_t>(j), data + j * dataStride, width);
		return COMPV_ERROR_CODE_S_OK;
	}
	template <typename T> T* ptr(size_t row = 0, size_t col = 0) { return reinterpret_cast<T*>(m_ptr) + row * m_stride + col; }
	template <typename T> const T* ptr(size_t row = 0, size_t col = 0) const { return reinterpret_cast<const T*>(m_ptr) + row * m_stride + col; }
	size_t rows() const { return m_rows; }
	size_t cols() const { return m_cols; }
	size_t stride() const { return m_stride; }
	size_t strideInBytes() const { return m_stride * m_elmt; }
	size_t elmtInBytes() const { return m_elmt; }
	size_t planeCount() const { return 1; }
	bool isEmpty() const { return !m_rows || !m_cols; }
private:
	CompVMat() : m_ptr(NULL), m_own(false), m_rows(0), m_cols(0), m_stride(0), m_elmt(0) {}
	void* m_ptr; bool m_own; size_t m_rows, m_cols, m_stride, m_elmt;
};

// ---- CompVCaps (compv_caps.h): typed helpers over set(id, ptr, size) ----
class CompVCaps {
public:
	virtual ~CompVCaps() {}
	virtual COMPV_ERROR_CODE set(int id, const void* valuePtr, size_t valueSize) = 0;
	virtual COMPV_ERROR_CODE get(int, void*, size_t) { return COMPV_ERROR_CODE_E_NOT_IMPLEMENTED; }
	COMPV_ERROR_CODE setInt(int id, int v) { return set(id, &v, sizeof(v)); }
	COMPV_ERROR_CODE setFloat32(int id, float v) { return set(id, &v, sizeof(v)); }
	COMPV_ERROR_CODE setBool(int id, bool v) { return set(id, &v, sizeof(v)); }
	COMPV_ERROR_CODE getFloat64(int id, double* v) { return get(id, v, sizeof(*v)); }
};

static inline bool compv_is_8u1(const CompVMatPtr& m) { return m && !m->isEmpty() && m->planeCount() == 1 && m->elmtInBytes() == sizeof(uint8_t); }

// ---- CompVEdgeDete (compv_features.h:199-215) ----
class CompVEdgeDete;
typedef std::shared_ptr<CompVEdgeDete> CompVEdgeDetePtr;
typedef CompVEdgeDetePtr* CompVEdgeDetePtrPtr;
class CompVEdgeDete : public CompVCaps {
public:
	~CompVEdgeDete() { cvb200_edge_dete_free(&m_h); }
	int id() const { return m_id; }
	COMPV_ERROR_CODE set(int id, const void* valuePtr, size_t valueSize) override { return cvb200_edge_dete_set(m_h, id, valuePtr, valueSize); }
	// edges gets the input's stride (canny_dete.cxx:249); `directions` is not produced (the reference's detectors ignore it as well unless asked by the SHT, disabled :104-108)
	COMPV_ERROR_CODE process(const CompVMatPtr& image, CompVMatPtrPtr edges, CompVMatPtrPtr directions = NULL) {
		(void)directions;
		COMPV_CHECK_EXP_RETURN(!compv_is_8u1(image) || !edges, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		COMPV_CHECK_CODE_RETURN(CompVMat::newObj<uint8_t>(edges, image->rows(), image->cols(), image->stride()));
		return cvb200_edge_dete_process(m_h, image->ptr<uint8_t>(), image->cols(), image->rows(), image->stride(), (*edges)->ptr<uint8_t>());
	}
	static COMPV_ERROR_CODE newObj(CompVEdgeDetePtrPtr dete, int id, float tLow = 0.68f, float tHigh = 0.68f * 2.f, size_t kernSize = 3) {
		COMPV_CHECK_EXP_RETURN(!dete, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		cvb200_edge_dete_t* h = NULL;
		COMPV_CHECK_CODE_RETURN(cvb200_edge_dete_new(&h, id, tLow, tHigh, kernSize));
		dete->reset(new CompVEdgeDete(h, id));
		return COMPV_ERROR_CODE_S_OK;
	}
	cvb200_edge_dete_t* handle() { return m_h; } // for the batched / device-pointer entry points of cvb200.h
private:
	CompVEdgeDete(cvb200_edge_dete_t* h, int id) : m_h(h), m_id(id) {}
	cvb200_edge_dete_t* m_h; int m_id;
};

// ---- CompVCornerDete (compv_features.h:160-174) ----
typedef cvb200_interest_point_t CompVInterestPoint; // {x, y, strength, orient, level, size} (compv_common.h:629-656)
typedef std::vector<CompVInterestPoint> CompVInterestPointVector;
class CompVCornerDete;
typedef std::shared_ptr<CompVCornerDete> CompVCornerDetePtr;
typedef CompVCornerDetePtr* CompVCornerDetePtrPtr;
class CompVCornerDete : public CompVCaps {
public:
	~CompVCornerDete() { cvb200_corner_dete_free(&m_h); }
	COMPV_ERROR_CODE set(int id, const void* valuePtr, size_t valueSize) override { return cvb200_corner_dete_set(m_h, id, valuePtr, valueSize); }
	COMPV_ERROR_CODE process(const CompVMatPtr& image, CompVInterestPointVector& interestPoints) {
		COMPV_CHECK_EXP_RETURN(!compv_is_8u1(image), COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		size_t capacity = 4096, count = 0;
		for (int attempt = 0; attempt < 2; ++attempt) {
			interestPoints.resize(capacity);
			const COMPV_ERROR_CODE rc = cvb200_corner_dete_process(m_h, image->ptr<uint8_t>(), image->cols(), image->rows(), image->stride(), interestPoints.data(), capacity, &count);
			if (rc == COMPV_ERROR_CODE_S_OK) break;
			if (rc != COMPV_ERROR_CODE_E_OUT_OF_BOUND || attempt) return rc;
			capacity = count; // the call reports the full count: second pass with room for all of them
		}
		interestPoints.resize(count < capacity ? count : capacity);
		return COMPV_ERROR_CODE_S_OK;
	}
	static COMPV_ERROR_CODE newObj(CompVCornerDetePtrPtr dete, int id) {
		COMPV_CHECK_EXP_RETURN(!dete, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		cvb200_corner_dete_t* h = NULL;
		COMPV_CHECK_CODE_RETURN(cvb200_corner_dete_new(&h, id));
		dete->reset(new CompVCornerDete(h));
		return COMPV_ERROR_CODE_S_OK;
	}
private:
	explicit CompVCornerDete(cvb200_corner_dete_t* h) : m_h(h) {}
	cvb200_corner_dete_t* m_h;
};

// ---- CompVHough (compv_features.h:217-227) ----
typedef cvb200_hough_line_t CompVHoughLine; // {rho, theta, strength} (compv_common.h:686-692)
typedef std::vector<CompVHoughLine> CompVHoughLineVector;
struct CompVPointFloat32 { float x, y, z; };
struct CompVLineFloat32 { CompVPointFloat32 a, b; };
typedef std::vector<CompVLineFloat32> CompVLineFloat32Vector;
class CompVHough;
typedef std::shared_ptr<CompVHough> CompVHoughPtr;
typedef CompVHoughPtr* CompVHoughPtrPtr;
class CompVHough : public CompVCaps {
public:
	~CompVHough() { cvb200_hough_free(&m_h); }
	COMPV_ERROR_CODE set(int id, const void* valuePtr, size_t valueSize) override { return cvb200_hough_set(m_h, id, valuePtr, valueSize); }
	COMPV_ERROR_CODE get(int id, void* valuePtr, size_t valueSize) override { return cvb200_hough_get(m_h, id, valuePtr, valueSize); }
	COMPV_ERROR_CODE process(const CompVMatPtr& edges, CompVHoughLineVector& lines, const CompVMatPtr& directions = CompVMatPtr()) {
		(void)directions;
		COMPV_CHECK_EXP_RETURN(!compv_is_8u1(edges), COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		size_t capacity = 4096, count = 0;
		for (int attempt = 0; attempt < 2; ++attempt) {
			lines.resize(capacity);
			COMPV_CHECK_CODE_RETURN(cvb200_hough_process(m_h, edges->ptr<uint8_t>(), edges->cols(), edges->rows(), edges->stride(), lines.data(), capacity, &count));
			if (count <= capacity) break;
			capacity = count;
		}
		lines.resize(count < capacity ? count : capacity);
		return COMPV_ERROR_CODE_S_OK;
	}
	static COMPV_ERROR_CODE newObj(CompVHoughPtrPtr hough, int id, float rho = 1.f, float theta = 1.f, size_t threshold = 1) {
		COMPV_CHECK_EXP_RETURN(!hough, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		cvb200_hough_t* h = NULL;
		COMPV_CHECK_CODE_RETURN(cvb200_hough_new(&h, id, rho, theta, threshold));
		hough->reset(new CompVHough(h, id));
		return COMPV_ERROR_CODE_S_OK;
	}
	// Polar (rho, theta) -> the two points where the line crosses x = 0 and x = width (or a vertical line when theta == 0).  Host arithmetic, as in the reference:
	// KHT measures rho from the image centre (houghkht.cxx:1249-1280), SHT from the origin (houghsht.cxx:566-592).
	COMPV_ERROR_CODE toCartesian(const size_t imageWidth, const size_t imageHeight, const CompVHoughLineVector& polar, CompVLineFloat32Vector& cartesian) const {
		COMPV_CHECK_EXP_RETURN(!imageWidth || !imageHeight, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		cartesian.resize(polar.size());
		const float widthF = static_cast<float>(imageWidth), heightF = static_cast<float>(imageHeight);
		const float r = std::sqrt((widthF * widthF) + (heightF * heightF));
		const bool centred = (m_id == COMPV_HOUGHKHT_ID);
		const float ox = centred ? widthF * 0.5f : 0.f, oy = centred ? heightF * 0.5f : 0.f;
		for (size_t i = 0; i < polar.size(); ++i) {
			const float rho = polar[i].rho, theta = polar[i].theta;
			CompVLineFloat32& l = cartesian[i];
			if (theta == 0.f) { l.a.x = l.b.x = rho + ox; l.a.y = r; l.b.y = -r; }
			else if (centred) {
				const float a = std::cos(theta) * ox, b = 1.f / std::sin(theta);
				l.a.x = 0.f; l.a.y = ((rho + a) * b) + oy;
				l.b.x = widthF; l.b.y = ((rho - a) * b) + oy;
			}
			else {
				const float a = std::cos(theta), b = 1.f / std::sin(theta);
				l.a.x = 0.f; l.a.y = rho * b;
				l.b.x = widthF; l.b.y = (rho - (widthF * a)) * b;
			}
			l.a.z = l.b.z = 1.f;
		}
		return COMPV_ERROR_CODE_S_OK;
	}
	cvb200_hough_t* handle() { return m_h; }
private:
	CompVHough(cvb200_hough_t* h, int id) : m_h(h), m_id(id) {}
	cvb200_hough_t* m_h; int m_id;
};

// ---- CompVHOG (compv_features.h:229-240) ----
struct CompVSizeSz { size_t width, height; CompVSizeSz(size_t w = 0, size_t h = 0) : width(w), height(h) {} };
class CompVHOG;
typedef std::shared_ptr<CompVHOG> CompVHOGPtr;
typedef CompVHOGPtr* CompVHOGPtrPtr;
class CompVHOG : public CompVCaps {
public:
	~CompVHOG() { cvb200_hog_free(&m_h); }
	COMPV_ERROR_CODE set(int id, const void* valuePtr, size_t valueSize) override { return cvb200_hog_set(m_h, id, valuePtr, valueSize); }
	// output: 1 x N float row vector (hog_std.cxx:337)
	COMPV_ERROR_CODE process(const CompVMatPtr& input, CompVMatPtrPtr output) {
		COMPV_CHECK_EXP_RETURN(!compv_is_8u1(input) || !output, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		size_t n = 0;
		COMPV_CHECK_CODE_RETURN(cvb200_hog_descriptor_size(m_h, input->cols(), input->rows(), &n));
		COMPV_CHECK_CODE_RETURN(CompVMat::newObj<float>(output, 1, n));
		return cvb200_hog_process(m_h, input->ptr<uint8_t>(), input->cols(), input->rows(), input->stride(), (*output)->ptr<float>(), n, &n);
	}
	static COMPV_ERROR_CODE newObj(CompVHOGPtrPtr hog, int id, const CompVSizeSz& blockSize = CompVSizeSz(16, 16), const CompVSizeSz& blockStride = CompVSizeSz(8, 8),
		const CompVSizeSz& cellSize = CompVSizeSz(8, 8), size_t nbins = 9, int blockNorm = COMPV_HOG_BLOCK_NORM_L2HYS, bool gradientSigned = true, int interp = COMPV_HOG_INTERPOLATION_BILINEAR) {
		COMPV_CHECK_EXP_RETURN(!hog, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		cvb200_hog_t* h = NULL;
		COMPV_CHECK_CODE_RETURN(cvb200_hog_new(&h, id, blockSize.width, blockSize.height, blockStride.width, blockStride.height, cellSize.width, cellSize.height, nbins, blockNorm,
			gradientSigned ? 1 : 0, interp));
		hog->reset(new CompVHOG(h));
		return COMPV_ERROR_CODE_S_OK;
	}
private:
	explicit CompVHOG(cvb200_hog_t* h) : m_h(h) {}
	cvb200_hog_t* m_h;
};

// ---- CompVConnectedComponentLabeling + results (compv_ccl.h:105-241) ----
typedef int32_t CompVConnectedComponentId;
typedef std::vector<CompVConnectedComponentId> CompVConnectedComponentIdsVector;
typedef cvb200_rect16_t CompVConnectedComponentBoundingBox; // CompVRectInt16 {left, top, right, bottom}
typedef std::vector<CompVConnectedComponentBoundingBox> CompVConnectedComponentBoundingBoxesVector;
struct CompVPoint2DInt16 { int16_t x, y; };
typedef std::vector<CompVPoint2DInt16> CompVConnectedComponentPoints;
typedef std::vector<CompVConnectedComponentPoints> CompVConnectedComponentPointsVector;
enum COMPV_CCL_EXTRACT_TYPE { COMPV_CCL_EXTRACT_TYPE_SEGMENT, COMPV_CCL_EXTRACT_TYPE_BLOB };
struct CompVConnectedComponentLabelingRegionMser { CompVConnectedComponentPoints points; CompVConnectedComponentBoundingBox boundingBox; };
typedef std::vector<CompVConnectedComponentLabelingRegionMser> CompVConnectedComponentLabelingRegionMserVector;

class CompVConnectedComponentLabelingResult;
typedef std::shared_ptr<CompVConnectedComponentLabelingResult> CompVConnectedComponentLabelingResultPtr;
typedef CompVConnectedComponentLabelingResultPtr* CompVConnectedComponentLabelingResultPtrPtr;
// One class for both result kinds; the PLSL accessors fail on an MSER result and vice versa, as reinterpret_castr<> returns NULL for the wrong kind (compv_ccl.h:206-218).
class CompVConnectedComponentLabelingResult {
public:
	~CompVConnectedComponentLabelingResult() { cvb200_ccl_result_free(&m_h); }
	int id() const { return m_id; }
	int32_t backgroundLabelId() const { return 0; }
	size_t labelsCount() const { return cvb200_ccl_result_labels_count(m_h); }
	// CompVConnectedComponentLabelingResultLSL
	COMPV_ERROR_CODE labelIds(CompVConnectedComponentIdsVector& ids) const {
		COMPV_CHECK_EXP_RETURN(m_id != COMPV_PLSL_ID, COMPV_ERROR_CODE_E_NOT_IMPLEMENTED);
		ids.resize(labelsCount());
		for (size_t i = 0; i < ids.size(); ++i) ids[i] = static_cast<int32_t>(i + 1); // std::iota(1) (ccl_lsl.cxx:745-746)
		return COMPV_ERROR_CODE_S_OK;
	}
	COMPV_ERROR_CODE boundingBoxes(CompVConnectedComponentBoundingBoxesVector& boxes) const {
		size_t n = 0;
		boxes.resize(labelsCount());
		return cvb200_ccl_result_bounding_boxes(m_h, boxes.data(), boxes.size(), &n);
	}
	COMPV_ERROR_CODE debugFlatten(CompVMatPtrPtr ptr32sLabels) const {
		COMPV_CHECK_EXP_RETURN(!ptr32sLabels, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		COMPV_CHECK_EXP_RETURN(m_id != COMPV_PLSL_ID, COMPV_ERROR_CODE_E_NOT_IMPLEMENTED);
		COMPV_CHECK_CODE_RETURN(CompVMat::newObj<int32_t>(ptr32sLabels, m_height, m_width, m_width)); // strideless, like the reference (ccl_lsl_result.cxx:62)
		return cvb200_ccl_result_flatten(m_h, (*ptr32sLabels)->ptr<int32_t>(), m_width);
	}
	// extract (ccl_lsl_result.cxx:100-134, 308-416): per label, every pixel (BLOB) or the two end points {start, y}, {end, y} of every run (SEGMENT), rows top-down, runs
	// left to right -- the order of the reference's single-threaded fill (its threaded fill orders rows by arrival).
	COMPV_ERROR_CODE extract(CompVConnectedComponentPointsVector& points, COMPV_CCL_EXTRACT_TYPE type = COMPV_CCL_EXTRACT_TYPE_BLOB) const {
		points.clear();
		COMPV_CHECK_EXP_RETURN(m_id != COMPV_PLSL_ID, COMPV_ERROR_CODE_E_NOT_IMPLEMENTED); // lmser_result.cxx:41-45
		COMPV_CHECK_EXP_RETURN(type != COMPV_CCL_EXTRACT_TYPE_SEGMENT && type != COMPV_CCL_EXTRACT_TYPE_BLOB, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		const uint32_t* rowOffsets; const cvb200_ccl_range_t* ranges; size_t n = 0;
		COMPV_CHECK_CODE_RETURN(cvb200_ccl_result_segments(m_h, &rowOffsets, &ranges, &n));
		if (!labelsCount()) return COMPV_ERROR_CODE_S_OK;
		points.resize(labelsCount());
		std::vector<size_t> counts(points.size(), 0);
		for (size_t s = 0; s < n; ++s) counts[static_cast<size_t>(ranges[s].a - 1)] += (type == COMPV_CCL_EXTRACT_TYPE_BLOB) ? static_cast<size_t>(ranges[s].end - ranges[s].start) : 2;
		for (size_t a = 0; a < points.size(); ++a) points[a].reserve(counts[a]);
		for (size_t j = 0; j < m_height; ++j) {
			for (uint32_t s = rowOffsets[j]; s < rowOffsets[j + 1]; ++s) {
				CompVConnectedComponentPoints& pp = points[static_cast<size_t>(ranges[s].a - 1)];
				CompVPoint2DInt16 pt; pt.y = static_cast<int16_t>(j);
				if (type == COMPV_CCL_EXTRACT_TYPE_BLOB) for (int16_t x = ranges[s].start; x < ranges[s].end; ++x) { pt.x = x; pp.push_back(pt); }
				else { pt.x = ranges[s].start; pp.push_back(pt); pt.x = ranges[s].end; pp.push_back(pt); }
			}
		}
		return COMPV_ERROR_CODE_S_OK;
	}
	// CompVConnectedComponentLabelingResultLMSER::points() / boundingBoxes()
	COMPV_ERROR_CODE points(CompVConnectedComponentLabelingRegionMserVector& regions) const {
		const int32_t* sizes; const cvb200_rect16_t* boxes; const int16_t* pts; size_t nr = 0, np = 0;
		COMPV_CHECK_CODE_RETURN(cvb200_ccl_result_regions(m_h, &sizes, &boxes, &pts, &nr, &np));
		regions.resize(nr);
		for (size_t i = 0, o = 0; i < nr; ++i) {
			regions[i].boundingBox = boxes[i];
			regions[i].points.resize(static_cast<size_t>(sizes[i]));
			memcpy(regions[i].points.data(), pts + 2 * o, static_cast<size_t>(sizes[i]) * sizeof(CompVPoint2DInt16));
			o += static_cast<size_t>(sizes[i]);
		}
		return COMPV_ERROR_CODE_S_OK;
	}
private:
	friend class CompVConnectedComponentLabeling;
	CompVConnectedComponentLabelingResult(cvb200_ccl_result_t* h, int id, size_t w, size_t hh) : m_h(h), m_id(id), m_width(w), m_height(hh) {}
	cvb200_ccl_result_t* m_h; int m_id; size_t m_width, m_height;
};

class CompVConnectedComponentLabeling;
typedef std::shared_ptr<CompVConnectedComponentLabeling> CompVConnectedComponentLabelingPtr;
typedef CompVConnectedComponentLabelingPtr* CompVConnectedComponentLabelingPtrPtr;
class CompVConnectedComponentLabeling : public CompVCaps {
public:
	~CompVConnectedComponentLabeling() { cvb200_ccl_free(&m_h); }
	int id() const { return m_id; }
	COMPV_ERROR_CODE set(int id, const void* valuePtr, size_t valueSize) override { return cvb200_ccl_set(m_h, id, valuePtr, valueSize); }
	COMPV_ERROR_CODE process(const CompVMatPtr& ptr8uData, CompVConnectedComponentLabelingResultPtrPtr result) const {
		COMPV_CHECK_EXP_RETURN(!compv_is_8u1(ptr8uData) || !result, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		cvb200_ccl_result_t* r = NULL;
		if (*result && (*result)->id() == m_id) { r = (*result)->m_h; (*result)->m_h = NULL; } // reuse the result of the same kind (ccl_lsl.cxx:585-592)
		const COMPV_ERROR_CODE rc = cvb200_ccl_process(m_h, ptr8uData->ptr<uint8_t>(), ptr8uData->cols(), ptr8uData->rows(), ptr8uData->stride(), &r);
		if (rc != COMPV_ERROR_CODE_S_OK) { cvb200_ccl_result_free(&r); return rc; }
		result->reset(new CompVConnectedComponentLabelingResult(r, m_id, ptr8uData->cols(), ptr8uData->rows()));
		return COMPV_ERROR_CODE_S_OK;
	}
	static COMPV_ERROR_CODE newObj(CompVConnectedComponentLabelingPtrPtr ccl, int id, int delta = 5, double min_area = 0.0002, double max_area = 0.5, double max_variation = 0.5,
		double min_diversity = 0.5, int connectivity = 8) { // defaults: compv_ccl.h:23-28
		COMPV_CHECK_EXP_RETURN(!ccl, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		cvb200_ccl_t* h = NULL;
		COMPV_CHECK_CODE_RETURN(cvb200_ccl_new_ex(&h, id, delta, min_area, max_area, max_variation, min_diversity, connectivity));
		ccl->reset(new CompVConnectedComponentLabeling(h, id));
		return COMPV_ERROR_CODE_S_OK;
	}
private:
	CompVConnectedComponentLabeling(cvb200_ccl_t* h, int id) : m_h(h), m_id(id) {}
	cvb200_ccl_t* m_h; int m_id;
};

// ---- CompVImage thresholds (compv_image.h:63-67) ----
struct CompVImage {
	static COMPV_ERROR_CODE thresholdGlobal(const CompVMatPtr& input, CompVMatPtrPtr output, const double threshold) {
		COMPV_CHECK_EXP_RETURN(!compv_is_8u1(input) || !output, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		COMPV_CHECK_CODE_RETURN(CompVMat::newObj<uint8_t>(output, input->rows(), input->cols(), input->stride()));
		return cvb200_threshold_global(input->ptr<uint8_t>(), input->cols(), input->rows(), input->stride(), threshold, (*output)->ptr<uint8_t>());
	}
	static COMPV_ERROR_CODE thresholdOtsu(const CompVMatPtr& input, double& threshold, CompVMatPtrPtr output = NULL) {
		COMPV_CHECK_EXP_RETURN(!compv_is_8u1(input), COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		if (output) COMPV_CHECK_CODE_RETURN(CompVMat::newObj<uint8_t>(output, input->rows(), input->cols(), input->stride()));
		return cvb200_threshold_otsu(input->ptr<uint8_t>(), input->cols(), input->rows(), input->stride(), &threshold, output ? (*output)->ptr<uint8_t>() : NULL);
	}
	static COMPV_ERROR_CODE thresholdAdaptive(const CompVMatPtr& input, CompVMatPtrPtr output, const size_t blockSize, const double delta, const double maxVal = 255.0, bool invert = false) {
		COMPV_CHECK_EXP_RETURN(!compv_is_8u1(input) || !output, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		COMPV_CHECK_CODE_RETURN(CompVMat::newObj<uint8_t>(output, input->rows(), input->cols(), input->stride()));
		return cvb200_threshold_adaptive(input->ptr<uint8_t>(), input->cols(), input->rows(), input->stride(), blockSize, delta, maxVal, invert ? 1 : 0, (*output)->ptr<uint8_t>());
	}
};

// ---- CompVMathConvlt::convlt1 (compv_math_convlt.h:25-55): the type combinations the reference instantiates ----
struct CompVMathConvlt {
	static COMPV_ERROR_CODE convlt1(const uint8_t* in, size_t w, size_t h, size_t stride, const int16_t* vt, const int16_t* hz, size_t ks, int16_t* out, COMPV_BORDER_TYPE b = COMPV_BORDER_TYPE_ZERO) { return cvb200_convlt1_8u16s16s(in, w, h, stride, vt, hz, ks, out, b); }
	static COMPV_ERROR_CODE convlt1(const int16_t* in, size_t w, size_t h, size_t stride, const int16_t* vt, const int16_t* hz, size_t ks, int16_t* out, COMPV_BORDER_TYPE b = COMPV_BORDER_TYPE_ZERO) { return cvb200_convlt1_16s16s16s(in, w, h, stride, vt, hz, ks, out, b); }
	static COMPV_ERROR_CODE convlt1(const uint8_t* in, size_t w, size_t h, size_t stride, const float* vt, const float* hz, size_t ks, uint8_t* out, COMPV_BORDER_TYPE b = COMPV_BORDER_TYPE_ZERO) { return cvb200_convlt1_8u32f8u(in, w, h, stride, vt, hz, ks, out, b); }
	static COMPV_ERROR_CODE convlt1(const uint8_t* in, size_t w, size_t h, size_t stride, const float* vt, const float* hz, size_t ks, float* out, COMPV_BORDER_TYPE b = COMPV_BORDER_TYPE_ZERO) { return cvb200_convlt1_8u32f32f(in, w, h, stride, vt, hz, ks, out, b); }
	static COMPV_ERROR_CODE convlt1(const float* in, size_t w, size_t h, size_t stride, const float* vt, const float* hz, size_t ks, float* out, COMPV_BORDER_TYPE b = COMPV_BORDER_TYPE_ZERO) { return cvb200_convlt1_32f32f32f(in, w, h, stride, vt, hz, ks, out, b); }
	static COMPV_ERROR_CODE convlt1(const float* in, size_t w, size_t h, size_t stride, const float* vt, const float* hz, size_t ks, uint8_t* out, COMPV_BORDER_TYPE b = COMPV_BORDER_TYPE_ZERO) { return cvb200_convlt1_32f32f8u(in, w, h, stride, vt, hz, ks, out, b); }
	static COMPV_ERROR_CODE convlt1FixedPoint(const uint8_t* in, size_t w, size_t h, size_t stride, const uint16_t* vt, const uint16_t* hz, size_t ks, uint8_t* out, COMPV_BORDER_TYPE b = COMPV_BORDER_TYPE_ZERO) { return cvb200_convlt1_fxp_8u16u8u(in, w, h, stride, vt, hz, ks, out, b); }
};

// ---- CompVMathMorph (base/include/compv/base/math/compv_math_morph.h) ----
enum COMPV_MATH_MORPH_STREL_TYPE { COMPV_MATH_MORPH_STREL_TYPE_RECT = CVB200_MATH_MORPH_STREL_TYPE_RECT, COMPV_MATH_MORPH_STREL_TYPE_DIAMOND = CVB200_MATH_MORPH_STREL_TYPE_DIAMOND,
	COMPV_MATH_MORPH_STREL_TYPE_CROSS = CVB200_MATH_MORPH_STREL_TYPE_CROSS };
enum COMPV_MATH_MORPH_OP_TYPE { COMPV_MATH_MORPH_OP_TYPE_ERODE = CVB200_MATH_MORPH_OP_TYPE_ERODE, COMPV_MATH_MORPH_OP_TYPE_DILATE = CVB200_MATH_MORPH_OP_TYPE_DILATE,
	COMPV_MATH_MORPH_OP_TYPE_OPEN = CVB200_MATH_MORPH_OP_TYPE_OPEN, COMPV_MATH_MORPH_OP_TYPE_CLOSE = CVB200_MATH_MORPH_OP_TYPE_CLOSE };
struct CompVMathMorph {
	static COMPV_ERROR_CODE buildStructuringElement(CompVMatPtrPtr strel, const CompVSizeSz size, COMPV_MATH_MORPH_STREL_TYPE type = COMPV_MATH_MORPH_STREL_TYPE_RECT) {
		COMPV_CHECK_EXP_RETURN(!strel || !size.width || !size.height, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		COMPV_CHECK_CODE_RETURN(CompVMat::newObj<uint8_t>(strel, size.height, size.width));
		return cvb200_morph_build_strel((*strel)->ptr<uint8_t>(), size.width, size.height, (*strel)->stride(), type);
	}
	static COMPV_ERROR_CODE process(const CompVMatPtr& input, const CompVMatPtr& strel, CompVMatPtrPtr output, COMPV_MATH_MORPH_OP_TYPE opType, COMPV_BORDER_TYPE borderType = COMPV_BORDER_TYPE_REPLICATE) {
		COMPV_CHECK_EXP_RETURN(!compv_is_8u1(input) || !compv_is_8u1(strel) || !output || input == *output, COMPV_ERROR_CODE_E_INVALID_PARAMETER); // compv_math_morph.cxx:131-142
		COMPV_CHECK_CODE_RETURN(CompVMat::newObj<uint8_t>(output, input->rows(), input->cols(), input->stride()));
		return cvb200_morph_process(input->ptr<uint8_t>(), input->cols(), input->rows(), input->stride(), strel->ptr<uint8_t>(), strel->cols(), strel->rows(), strel->stride(),
			(*output)->ptr<uint8_t>(), opType, borderType);
	}
};

} // namespace compv
