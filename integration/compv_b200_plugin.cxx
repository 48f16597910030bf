// The adapter INTEGRATION.md describes, as real code: built against the UNMODIFIED reference headers and linked with libcompv_b200.so, it registers
// B200-backed factories through the reference's own seam, CompVFeature::addFactory (base/compv_features.cxx:30-40), which REPLACES the entry of an id.
// After compv_b200_register(device) unmodified application code -- CompVEdgeDete::newObj(&d, COMPV_CANNY_ID, 59.f, 119.f); d->process(image, &edges);
// (samples/hough_lines/main.cxx:59-106) -- runs on the GPU.  Built by oracle/build_ref.sh into oracle/_ref/libcompv_b200_plugin.so where /root/reference exists;
// tests/test_plugin.py drives it through the reference's public API (via oracle/ref_shim.cxx).  No reference source is modified or copied.
#include "compv/base/compv_base.h"
#include "compv/base/compv_features.h"
#include "compv/base/image/compv_image.h"
#include "compv/base/compv_ccl.h"

#include <numeric>

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

// ---- CompVHOG: S-HOG (replaces core/features/hog/compv_core_feature_hog_std.cxx:196-393) ----
class CompVHOGB200 : public CompVHOG {
	cvb200_hog_t* m_h;
	explicit CompVHOGB200(cvb200_hog_t* h) : CompVHOG(COMPV_HOGS_ID), m_h(h) {}
public:
	virtual ~CompVHOGB200() { cvb200_hog_free(&m_h); }
	COMPV_OBJECT_GET_ID(CompVHOGB200);
	COMPV_ERROR_CODE set(int id, const void* valuePtr, size_t valueSize) override { return B200_RC(cvb200_hog_set(m_h, id, valuePtr, valueSize)); }
	COMPV_ERROR_CODE process(const CompVMatPtr& input, CompVMatPtrPtr output) override {
		COMPV_CHECK_EXP_RETURN(!b200_is_8u1(input) || !output, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		size_t n = 0;
		COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_hog_descriptor_size(m_h, input->cols(), input->rows(), &n)));
		CompVMatPtr output_ = *output;
		COMPV_CHECK_CODE_RETURN(CompVMat::newObjAligned<compv_float32_t>(&output_, 1, n)); // hog_std.cxx:337: a 1 x N row vector
		COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_hog_process(m_h, input->ptr<const uint8_t>(), input->cols(), input->rows(), input->stride(), output_->ptr<compv_float32_t>(), n, &n)));
		*output = output_;
		return COMPV_ERROR_CODE_S_OK;
	}
	static COMPV_ERROR_CODE newObj(CompVHOGPtrPtr hog, const CompVSizeSz& blockSize, const CompVSizeSz& blockStride, const CompVSizeSz& cellSize, const size_t nbins,
		const int blockNorm, const bool gradientSigned, const int interp) {
		COMPV_CHECK_EXP_RETURN(!hog, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		cvb200_hog_t* h = NULL;
		COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_hog_new(&h, CVB200_HOGS_ID, blockSize.width, blockSize.height, blockStride.width, blockStride.height, cellSize.width, cellSize.height,
			nbins, blockNorm, gradientSigned ? 1 : 0, interp)));
		CompVPtr<CompVHOGB200*> d = new CompVHOGB200(h);
		COMPV_CHECK_EXP_RETURN(!d, COMPV_ERROR_CODE_E_OUT_OF_MEMORY);
		*hog = *d;
		return COMPV_ERROR_CODE_S_OK;
	}
};

// ---- CompVConnectedComponentLabeling: PLSL and LMSER (replaces core/ccl/compv_core_ccl_lsl.cxx, compv_core_ccl_lmser.cxx and their result classes) ----
class CompVCclResultLSLB200 : public CompVConnectedComponentLabelingResultLSL {
	cvb200_ccl_result_t* m_r; size_t m_w, m_hh; CompVConnectedComponentIdsVector m_ids;
	CompVCclResultLSLB200(cvb200_ccl_result_t* r, size_t w, size_t h) : m_r(r), m_w(w), m_hh(h) {
		m_ids.resize(cvb200_ccl_result_labels_count(r));
		std::iota(m_ids.begin(), m_ids.end(), 1); // ccl_lsl.cxx:745-746
	}
public:
	virtual ~CompVCclResultLSLB200() { cvb200_ccl_result_free(&m_r); }
	COMPV_OBJECT_GET_ID(CompVCclResultLSLB200);
	size_t labelsCount() const override { return cvb200_ccl_result_labels_count(m_r); }
	const CompVConnectedComponentIdsVector& labelIds() const override { return m_ids; }
	COMPV_ERROR_CODE debugFlatten(CompVMatPtrPtr ptr32sLabels) const override {
		COMPV_CHECK_EXP_RETURN(!ptr32sLabels, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		COMPV_CHECK_EXP_RETURN(!labelsCount(), COMPV_ERROR_CODE_E_INVALID_STATE); // an all-background image has no LEA (ccl_lsl.cxx:676-680, ccl_lsl_result.cxx:54)
		CompVMatPtr ea = *ptr32sLabels;
		COMPV_CHECK_CODE_RETURN(CompVMat::newObjStrideless<int32_t>(&ea, m_hh, m_w));
		COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_ccl_result_flatten(m_r, ea->ptr<int32_t>(), ea->stride())));
		*ptr32sLabels = ea;
		return COMPV_ERROR_CODE_S_OK;
	}
	COMPV_ERROR_CODE boundingBoxes(CompVConnectedComponentBoundingBoxesVector& boxes) const override {
		static_assert(sizeof(CompVConnectedComponentBoundingBox) == sizeof(cvb200_rect16_t), "CompVRectInt16 layout");
		size_t n = 0;
		boxes.resize(labelsCount());
		return B200_RC(cvb200_ccl_result_bounding_boxes(m_r, reinterpret_cast<cvb200_rect16_t*>(boxes.data()), boxes.size(), &n));
	}
	// boxes of caller-supplied segment lists: host arithmetic on the caller's vectors (ccl_lsl_result.cxx:187-230)
	COMPV_ERROR_CODE boundingBoxes(const CompVConnectedComponentPointsVector& segments, CompVConnectedComponentBoundingBoxesVector& boxes) const override {
		boxes.clear();
		if (segments.empty()) return COMPV_ERROR_CODE_S_OK;
		boxes = CompVConnectedComponentBoundingBoxesVector(segments.size(), CompVConnectedComponentBoundingBox(static_cast<int16_t>(m_w), static_cast<int16_t>(m_hh), 0, 0));
		for (size_t j = 0; j < segments.size(); ++j) {
			const CompVConnectedComponentPoints& pts = segments[j];
			COMPV_CHECK_EXP_RETURN(pts.size() & 1, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
			CompVConnectedComponentBoundingBox& bb = boxes[j];
			for (size_t k = 0; k + 1 < pts.size(); k += 2) {
				bb.left = COMPV_MATH_MIN(bb.left, pts[k].x); bb.top = COMPV_MATH_MIN(bb.top, pts[k].y);
				bb.right = COMPV_MATH_MAX(bb.right, pts[k + 1].x); bb.bottom = COMPV_MATH_MAX(bb.bottom, pts[k + 1].y);
			}
		}
		return COMPV_ERROR_CODE_S_OK;
	}
	COMPV_ERROR_CODE remove(CompVConnectedComponentCallbackRemoveLabel, size_t& removedCount) override { removedCount = 0; return COMPV_ERROR_CODE_E_NOT_IMPLEMENTED; } // deprecated in the reference
	// extract (ccl_lsl_result.cxx:100-134, 308-416): per label every pixel (BLOB) or the end points {start, y}, {end, y} of every run (SEGMENT); rows top-down, runs left
	// to right: the order of the reference's single-threaded fill
	COMPV_ERROR_CODE extract(CompVConnectedComponentPointsVector& points, COMPV_CCL_EXTRACT_TYPE type = COMPV_CCL_EXTRACT_TYPE_BLOB) const override {
		points.clear();
		if (!labelsCount()) return COMPV_ERROR_CODE_S_OK;
		COMPV_CHECK_EXP_RETURN(type != COMPV_CCL_EXTRACT_TYPE_SEGMENT && type != COMPV_CCL_EXTRACT_TYPE_BLOB, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		const uint32_t* rowOffsets; const cvb200_ccl_range_t* ranges; size_t n = 0;
		COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_ccl_result_segments(m_r, &rowOffsets, &ranges, &n)));
		points.resize(labelsCount());
		std::vector<size_t> counts(points.size(), 0);
		for (size_t s = 0; s < n; ++s) counts[static_cast<size_t>(ranges[s].a - 1)] += (type == COMPV_CCL_EXTRACT_TYPE_BLOB) ? static_cast<size_t>(ranges[s].end - ranges[s].start) : 2;
		for (size_t a = 0; a < points.size(); ++a) points[a].reserve(counts[a]);
		for (size_t j = 0; j < m_hh; ++j) {
			for (uint32_t s = rowOffsets[j]; s < rowOffsets[j + 1]; ++s) {
				CompVConnectedComponentPoints& pp = points[static_cast<size_t>(ranges[s].a - 1)];
				if (type == COMPV_CCL_EXTRACT_TYPE_BLOB) for (int16_t x = ranges[s].start; x < ranges[s].end; ++x) pp.push_back(CompVConnectedComponentPoint(x, static_cast<int16_t>(j)));
				else { pp.push_back(CompVConnectedComponentPoint(ranges[s].start, static_cast<int16_t>(j))); pp.push_back(CompVConnectedComponentPoint(ranges[s].end, static_cast<int16_t>(j))); }
			}
		}
		return COMPV_ERROR_CODE_S_OK;
	}
	static COMPV_ERROR_CODE newObj(CompVConnectedComponentLabelingResultPtrPtr result, cvb200_ccl_result_t* r, size_t w, size_t h) {
		CompVPtr<CompVCclResultLSLB200*> o = new CompVCclResultLSLB200(r, w, h);
		COMPV_CHECK_EXP_RETURN(!o, COMPV_ERROR_CODE_E_OUT_OF_MEMORY);
		*result = *o;
		return COMPV_ERROR_CODE_S_OK;
	}
};

class CompVCclResultLMSERB200 : public CompVConnectedComponentLabelingResultLMSER {
	CompVConnectedComponentLabelingRegionMserVector m_regions;
	CompVCclResultLMSERB200() {}
public:
	COMPV_OBJECT_GET_ID(CompVCclResultLMSERB200);
	size_t labelsCount() const override { return m_regions.size(); }
	const CompVConnectedComponentLabelingRegionMserVector& points() const override { return m_regions; }
	const CompVConnectedComponentLabelingRegionMserVector& boundingBoxes() const override { return m_regions; } // the same vector, boxes filled (lmser_result.cxx:50-88)
	static COMPV_ERROR_CODE newObj(CompVConnectedComponentLabelingResultPtrPtr result, const cvb200_ccl_result_t* r) {
		CompVPtr<CompVCclResultLMSERB200*> o = new CompVCclResultLMSERB200();
		COMPV_CHECK_EXP_RETURN(!o, COMPV_ERROR_CODE_E_OUT_OF_MEMORY);
		const int32_t* sizes; const cvb200_rect16_t* boxes; const int16_t* pts; size_t nr = 0, np = 0;
		COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_ccl_result_regions(r, &sizes, &boxes, &pts, &nr, &np)));
		o->m_regions.resize(nr);
		for (size_t i = 0, off = 0; i < nr; ++i) {
			CompVConnectedComponentLabelingRegionMser& reg = o->m_regions[i];
			reg.boundingBox = CompVConnectedComponentBoundingBox(boxes[i].left, boxes[i].top, boxes[i].right, boxes[i].bottom);
			reg.points.resize(static_cast<size_t>(sizes[i]));
			static_assert(sizeof(CompVConnectedComponentPoint) == 2 * sizeof(int16_t), "CompVPoint2DInt16 layout");
			memcpy(reinterpret_cast<void*>(reg.points.data()), pts + 2 * off, static_cast<size_t>(sizes[i]) * sizeof(CompVConnectedComponentPoint));
			off += static_cast<size_t>(sizes[i]);
		}
		*result = *o;
		return COMPV_ERROR_CODE_S_OK;
	}
};

// The reference pushes delta / areas / variation / diversity / connectivity into the object AFTER the factory built it (base/compv_ccl.cxx:90-95), so the device object
// is (re)created from the getters when process() runs.
template <int ID>
class CompVCclB200 : public CompVConnectedComponentLabeling {
	mutable cvb200_ccl_t* m_h;
	mutable int m_delta, m_conn; mutable double m_minA, m_maxA, m_maxV, m_minD;
	int m_plslType; bool m_hasType;
	CompVCclB200() : CompVConnectedComponentLabeling(ID), m_h(NULL), m_delta(-1), m_conn(-1), m_minA(-1), m_maxA(-1), m_maxV(-1), m_minD(-1), m_plslType(0), m_hasType(false) {}
	COMPV_ERROR_CODE ensure() const {
		if (m_h && m_delta == delta() && m_conn == connectivity() && m_minA == minArea() && m_maxA == maxArea() && m_maxV == maxVariation() && m_minD == minDiversity()) return COMPV_ERROR_CODE_S_OK;
		cvb200_ccl_free(&m_h);
		COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_ccl_new_ex(&m_h, ID, delta(), minArea(), maxArea(), maxVariation(), minDiversity(), connectivity())));
		if (m_hasType) COMPV_CHECK_CODE_RETURN(B200_RC(cvb200_ccl_set(m_h, CVB200_PLSL_SET_INT_TYPE, &m_plslType, sizeof(m_plslType))));
		m_delta = delta(); m_conn = connectivity(); m_minA = minArea(); m_maxA = maxArea(); m_maxV = maxVariation(); m_minD = minDiversity();
		return COMPV_ERROR_CODE_S_OK;
	}
public:
	virtual ~CompVCclB200() { cvb200_ccl_free(&m_h); }
	COMPV_OBJECT_GET_ID(CompVCclB200);
	COMPV_ERROR_CODE set(int id, const void* valuePtr, size_t valueSize) override {
		if (ID == COMPV_PLSL_ID && id == COMPV_PLSL_SET_INT_TYPE) { // ccl_lsl.cxx: the LSL flavour; every flavour yields the same labels, the device path has one
			COMPV_CHECK_EXP_RETURN(!valuePtr || valueSize != sizeof(int), COMPV_ERROR_CODE_E_INVALID_PARAMETER);
			m_plslType = *static_cast<const int*>(valuePtr); m_hasType = true;
			return m_h ? B200_RC(cvb200_ccl_set(m_h, id, valuePtr, valueSize)) : COMPV_ERROR_CODE_S_OK;
		}
		return CompVConnectedComponentLabeling::set(id, valuePtr, valueSize); // COMPV_CCL_SET_INT_CONNECTIVITY (compv_ccl.cxx:24-40)
	}
	COMPV_ERROR_CODE process(const CompVMatPtr& ptr8uData, CompVConnectedComponentLabelingResultPtrPtr result) const override {
		COMPV_CHECK_EXP_RETURN(!b200_is_8u1(ptr8uData) || !result, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		COMPV_CHECK_CODE_RETURN(ensure());
		cvb200_ccl_result_t* r = NULL;
		COMPV_ERROR_CODE rc = B200_RC(cvb200_ccl_process(m_h, ptr8uData->ptr<const uint8_t>(), ptr8uData->cols(), ptr8uData->rows(), ptr8uData->stride(), &r));
		if (rc != COMPV_ERROR_CODE_S_OK) { cvb200_ccl_result_free(&r); return rc; }
		if (ID == COMPV_PLSL_ID) return CompVCclResultLSLB200::newObj(result, r, ptr8uData->cols(), ptr8uData->rows()); // owns r
		rc = CompVCclResultLMSERB200::newObj(result, r);
		cvb200_ccl_result_free(&r);
		return rc;
	}
	static COMPV_ERROR_CODE newObj(CompVConnectedComponentLabelingPtrPtr ccl) {
		COMPV_CHECK_EXP_RETURN(!ccl, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
		CompVPtr<CompVCclB200<ID>*> o = new CompVCclB200<ID>();
		COMPV_CHECK_EXP_RETURN(!o, COMPV_ERROR_CODE_E_OUT_OF_MEMORY);
		*ccl = *o;
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
	{ COMPV_HOGS_ID, "Standard HOG descriptor (B200)", nullptr, nullptr, nullptr, nullptr, CompVHOGB200::newObj },
};
// the CCL twin of the seam: CompVConnectedComponentLabeling::addFactory (base/compv_ccl.cxx:42-52)
static const CompVConnectedComponentLabelingFactory kB200CclFactories[] = {
	{ COMPV_PLSL_ID, "Parallel Light Speed Labeling (B200)", CompVCclB200<COMPV_PLSL_ID>::newObj },
	{ COMPV_LMSER_ID, "Linear time MSER (B200)", CompVCclB200<COMPV_LMSER_ID>::newObj },
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
	for (size_t i = 0; i < sizeof(COMPV_NAMESPACE::kB200CclFactories) / sizeof(COMPV_NAMESPACE::kB200CclFactories[0]); ++i) {
		rc = COMPV_NAMESPACE::CompVConnectedComponentLabeling::addFactory(&COMPV_NAMESPACE::kB200CclFactories[i]);
		if (rc != COMPV_NAMESPACE::COMPV_ERROR_CODE_S_OK) return static_cast<int>(rc);
	}
	return 0;
}
