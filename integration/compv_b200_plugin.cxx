// The adapter INTEGRATION.md describes, as real code: built against the UNMODIFIED reference headers and linked with libcompv_b200.so, it registers
// B200-backed factories through the reference's own seam, CompVFeature::addFactory (base/compv_features.cxx:30-40), which REPLACES the entry of an id.
// After compv_b200_register(device) unmodified application code -- CompVEdgeDete::newObj(&d, COMPV_CANNY_ID, 59.f, 119.f); d->process(image, &edges);
// (samples/hough_lines/main.cxx:59-106) -- runs on the GPU.  Built by oracle/build_ref.sh into oracle/_ref/libcompv_b200_plugin.so where /root/reference exists;
// tests/test_plugin.py drives it through the reference's public API (via oracle/ref_shim.cxx).  No reference source is modified or copied.
#include "compv/base/compv_base.h"
#include "compv/base/compv_features.h"
#include "compv/base/image/compv_image.h"

#include "cvb200.h"

COMPV_NAMESPACE_BEGIN()

#define B200_RC(expr) static_cast<COMPV_ERROR_CODE>(expr)
static bool b200_is_8u1(const CompVMatPtr& m) { return m && !m->isEmpty() && m->planeCount() == 1 && m->elmtInBytes() == sizeof(uint8_t); }

// ---- CompVEdgeDete: Canny / Sobel / Scharr / Prewitt (replaces core/features/edges/compv_core_feature_canny_dete.cxx, ..._edge_dete.cxx) ----
class CompVEdgeDeteB200 : public CompVEdgeDete {
	cvb200_edge_dete_t* m_h;
	CompVEdgeDeteB200(int id, cvb200_edge_dete_t* h) : CompVEdgeDete(id), m_h(h) {}
public:
	virtual ~CompVEdgeDeteB200() { cvb200_edge_dete_free(&m_h); }
	COMPV_OBJECT_GET_ID(CompVEdgeDeteB200);
	COMPV_ERROR_CODE set(int id, const void* valuePtr, size_t valueSize) override { return B200_RC(cvb200_edge_dete_set(m_h, id, valuePtr, valueSize)); }
	COMPV_ERROR_CODE process(const CompVMatPtr& image, CompVMatPtrPtr edges, CompVMatPtrPtr directions = NULL) override {
		COMPV_CHECK_EXP_RETURN(!b200_is_8u1(image) || !edges, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		COMPV_CHECK_CODE_RETURN(CompVImage::newObj8u(edges, COMPV_SUBTYPE_PIXELS_Y, image->cols(), image->rows(), image->stride())); // canny_dete.cxx:249
		return B200_RC(cvb200_edge_dete_process(m_h, image->ptr<const uint8_t>(), image->cols(), image->rows(), image->stride(), (*edges)->ptr<uint8_t>()));
	}
	template <int ID>
	static COMPV_ERROR_CODE newObj(CompVEdgeDetePtrPtr dete, float tLow, float tHigh, size_t kernSize) {
		COMPV_CHECK_EXP_RETURN(!dete, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		cvb200_edge_dete_t* h = NULL;
		COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_edge_dete_new(&h, ID, tLow, tHigh, kernSize)));
		CompVPtr<CompVEdgeDeteB200*> d = new CompVEdgeDeteB200(ID, h);
		COMPV_CHECK_EXP_RETURN(!d, COMPV_ERROR_CODE_E_OUT_OF_MEMORY);
		*dete = *d;
		return COMPV_ERROR_CODE_S_OK;
	}
};

// ---- CompVHough: SHT / KHT (replaces core/features/hough/compv_core_feature_hough{sht,kht}.cxx) ----
class CompVHoughB200 : public CompVHough {
	cvb200_hough_t* m_h; int m_id;
	CompVHoughB200(int id, cvb200_hough_t* h) : CompVHough(id), m_h(h), m_id(id) {}
public:
	virtual ~CompVHoughB200() { cvb200_hough_free(&m_h); }
	COMPV_OBJECT_GET_ID(CompVHoughB200);
	COMPV_ERROR_CODE set(int id, const void* valuePtr, size_t valueSize) override { return B200_RC(cvb200_hough_set(m_h, id, valuePtr, valueSize)); }
	COMPV_ERROR_CODE get(int id, const void** valuePtrPtr, size_t valueSize) override { // the reference passes the address of the destination pointer (houghkht.cxx:194-206)
		COMPV_CHECK_EXP_RETURN(!valuePtrPtr || !*valuePtrPtr, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		return B200_RC(cvb200_hough_get(m_h, id, const_cast<void*>(*valuePtrPtr), valueSize));
	}
	COMPV_ERROR_CODE process(const CompVMatPtr& edges, CompVHoughLineVector& lines, const CompVMatPtr& directions = NULL) override {
		COMPV_CHECK_EXP_RETURN(!b200_is_8u1(edges), COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		static_assert(sizeof(CompVHoughLine) == sizeof(cvb200_hough_line_t), "CompVHoughLine layout");
		size_t capacity = 4096, count = 0;
		for (int attempt = 0; attempt < 2; ++attempt) {
			lines.resize(capacity);
			COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_hough_process(m_h, edges->ptr<const uint8_t>(), edges->cols(), edges->rows(), edges->stride(),
				reinterpret_cast<cvb200_hough_line_t*>(lines.data()), capacity, &count)));
			if (count <= capacity) break;
			capacity = count;
		}
		lines.resize(count < capacity ? count : capacity);
		return COMPV_ERROR_CODE_S_OK;
	}
	// host-side arithmetic on a handful of lines: the formulas of houghkht.cxx:1249-1280 / houghsht.cxx:566-592
	COMPV_ERROR_CODE toCartesian(const size_t imageWidth, const size_t imageHeight, const CompVHoughLineVector& polar, CompVLineFloat32Vector& cartesian) override {
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
			else if (centred) { const float a = std::cos(theta) * ox, b = 1.f / std::sin(theta); l.a.x = 0.f; l.a.y = ((rho + a) * b) + oy; l.b.x = widthF; l.b.y = ((rho - a) * b) + oy; }
			else { const float a = std::cos(theta), b = 1.f / std::sin(theta); l.a.x = 0.f; l.a.y = rho * b; l.b.x = widthF; l.b.y = (rho - (widthF * a)) * b; }
			l.a.z = l.b.z = 1.f;
		}
		return COMPV_ERROR_CODE_S_OK;
	}
	template <int ID>
	static COMPV_ERROR_CODE newObj(CompVHoughPtrPtr hough, float rho, float theta, size_t threshold) {
		COMPV_CHECK_EXP_RETURN(!hough, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		cvb200_hough_t* h = NULL;
		COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_hough_new(&h, ID, rho, theta, threshold)));
		CompVPtr<CompVHoughB200*> d = new CompVHoughB200(ID, h);
		COMPV_CHECK_EXP_RETURN(!d, COMPV_ERROR_CODE_E_OUT_OF_MEMORY);
		*hough = *d;
		return COMPV_ERROR_CODE_S_OK;
	}
};

// ---- CompVCornerDete: FAST (replaces core/features/fast/compv_core_feature_fast_dete.cxx) ----
class CompVCornerDeteB200 : public CompVCornerDete {
	cvb200_corner_dete_t* m_h;
	explicit CompVCornerDeteB200(cvb200_corner_dete_t* h) : CompVCornerDete(COMPV_FAST_ID), m_h(h) {}
public:
	virtual ~CompVCornerDeteB200() { cvb200_corner_dete_free(&m_h); }
	COMPV_OBJECT_GET_ID(CompVCornerDeteB200);
	COMPV_ERROR_CODE set(int id, const void* valuePtr, size_t valueSize) override { return B200_RC(cvb200_corner_dete_set(m_h, id, valuePtr, valueSize)); }
	COMPV_ERROR_CODE process(const CompVMatPtr& image, CompVInterestPointVector& interestPoints) override {
		COMPV_CHECK_EXP_RETURN(!b200_is_8u1(image), COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		static_assert(sizeof(CompVInterestPoint) == sizeof(cvb200_interest_point_t), "CompVInterestPoint layout");
		size_t capacity = 4096, count = 0;
		for (int attempt = 0; attempt < 2; ++attempt) {
			interestPoints.resize(capacity);
			const int rc = cvb200_corner_dete_process(m_h, image->ptr<const uint8_t>(), image->cols(), image->rows(), image->stride(),
				reinterpret_cast<cvb200_interest_point_t*>(interestPoints.data()), capacity, &count);
			if (rc == CVB200_S_OK) break;
			if (rc != CVB200_E_OUT_OF_BOUND || attempt) return B200_RC(rc);
			capacity = count;
		}
		interestPoints.resize(count < capacity ? count : capacity);
		return COMPV_ERROR_CODE_S_OK;
	}
	static COMPV_ERROR_CODE newObj(CompVCornerDetePtrPtr dete) {
		COMPV_CHECK_EXP_RETURN(!dete, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		cvb200_corner_dete_t* h = NULL;
		COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_corner_dete_new(&h, CVB200_FAST_ID)));
		CompVPtr<CompVCornerDeteB200*> d = new CompVCornerDeteB200(h);
		COMPV_CHECK_EXP_RETURN(!d, COMPV_ERROR_CODE_E_OUT_OF_MEMORY);
		*dete = *d;
		return COMPV_ERROR_CODE_S_OK;
	}
};

// one static table per id: the factory map stores the raw pointer (compv_features.cxx:38), so the tables must outlive every use
static const CompVFeatureFactory kB200Factories[] = {
	{ COMPV_CANNY_ID, "Canny edge detector (B200)", nullptr, nullptr, CompVEdgeDeteB200::newObj<COMPV_CANNY_ID>, nullptr, nullptr },
	{ COMPV_SOBEL_ID, "Sobel edge detector (B200)", nullptr, nullptr, CompVEdgeDeteB200::newObj<COMPV_SOBEL_ID>, nullptr, nullptr },
	{ COMPV_SCHARR_ID, "Scharr edge detector (B200)", nullptr, nullptr, CompVEdgeDeteB200::newObj<COMPV_SCHARR_ID>, nullptr, nullptr },
	{ COMPV_PREWITT_ID, "Prewitt edge detector (B200)", nullptr, nullptr, CompVEdgeDeteB200::newObj<COMPV_PREWITT_ID>, nullptr, nullptr },
	{ COMPV_HOUGHKHT_ID, "Kernel-based Hough transform (B200)", nullptr, nullptr, nullptr, CompVHoughB200::newObj<COMPV_HOUGHKHT_ID>, nullptr },
	{ COMPV_HOUGHSHT_ID, "Standard Hough transform (B200)", nullptr, nullptr, nullptr, CompVHoughB200::newObj<COMPV_HOUGHSHT_ID>, nullptr },
	{ COMPV_FAST_ID, "FAST corner detector (B200)", CompVCornerDeteB200::newObj, nullptr, nullptr, nullptr, nullptr },
};

COMPV_NAMESPACE_END()

// Call after CompVBase::init() + CompVCore::init().  Returns 0 or a COMPV_ERROR_CODE value; on failure (no GPU) nothing is registered: the reference keeps its own path.
extern "C" __attribute__((visibility("default"))) int compv_b200_register(int device)
{
	COMPV_NAMESPACE::COMPV_ERROR_CODE rc = static_cast<COMPV_NAMESPACE::COMPV_ERROR_CODE>(cvb200_init(device));
	if (rc != COMPV_NAMESPACE::COMPV_ERROR_CODE_S_OK) return static_cast<int>(rc);
	for (size_t i = 0; i < sizeof(COMPV_NAMESPACE::kB200Factories) / sizeof(COMPV_NAMESPACE::kB200Factories[0]); ++i) {
		rc = COMPV_NAMESPACE::CompVFeature::addFactory(&COMPV_NAMESPACE::kB200Factories[i]);
		if (rc != COMPV_NAMESPACE::COMPV_ERROR_CODE_S_OK) return static_cast<int>(rc);
	}
	return 0;
}
