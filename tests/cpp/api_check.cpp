// Exercises the C++ mirror of the reference interface (include/compv_b200.hpp) the way the reference's samples do (samples/hough_lines/main.cxx:59-106,
// unittests/feature_fast.cxx, unittests/ccl_binar.cxx) and dumps the results for tests/test_cpp_api.py, which compares them with the oracle.
//   api_check <width> <height> <frame.u8> <outdir>
#include "compv_b200.hpp"

#include <cstdio>
#include <string>

using namespace compv;

template <typename T>
static bool dump(const std::string& path, const T* data, size_t count)
{
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) return false;
	const bool ok = fwrite(data, sizeof(T), count, f) == count;
	fclose(f);
	return ok;
}

static COMPV_ERROR_CODE run(size_t width, size_t height, const char* framePath, const std::string& out)
{
	std::vector<uint8_t> raw(width * height);
	FILE* f = fopen(framePath, "rb");
	COMPV_CHECK_EXP_RETURN(!f, COMPV_ERROR_CODE_E_INVALID_PARAMETER);
	const size_t got = fread(raw.data(), 1, raw.size(), f);
	fclose(f);
	COMPV_CHECK_EXP_RETURN(got != raw.size(), COMPV_ERROR_CODE_E_INVALID_PARAMETER);

	// every call fails cleanly before init: there is no CPU path to fall back to
	CompVEdgeDetePtr dete;
	COMPV_CHECK_EXP_RETURN(CompVEdgeDete::newObj(&dete, COMPV_CANNY_ID, 59.f, 119.f) != COMPV_ERROR_CODE_E_NOT_INITIALIZED, COMPV_ERROR_CODE_E_INVALID_STATE);
	COMPV_CHECK_CODE_RETURN(CompVBase::init(0));

	CompVMatPtr image;
	COMPV_CHECK_CODE_RETURN(CompVMat::wrap8u(&image, raw.data(), width, height, width));

	// Canny -> KHT (samples/hough_lines/main.cxx)
	COMPV_CHECK_CODE_RETURN(CompVEdgeDete::newObj(&dete, COMPV_CANNY_ID, 59.f, 119.f));
	CompVMatPtr edges;
	COMPV_CHECK_CODE_RETURN(dete->process(image, &edges));
	std::vector<uint8_t> packed(width * height);
	for (size_t j = 0; j < height; ++j) memcpy(&packed[j * width], edges->ptr<uint8_t>(j), width);
	COMPV_CHECK_EXP_RETURN(!dump(out + "/edges.u8", packed.data(), packed.size()), COMPV_ERROR_CODE_E_INVALID_STATE);
	CompVHoughPtr hough;
	COMPV_CHECK_CODE_RETURN(CompVHough::newObj(&hough, COMPV_HOUGHKHT_ID, 1.f, 1.f, 50));
	COMPV_CHECK_CODE_RETURN(hough->setInt(COMPV_HOUGH_SET_INT_MAXLINES, 20));
	CompVHoughLineVector lines;
	COMPV_CHECK_CODE_RETURN(hough->process(edges, lines));
	COMPV_CHECK_EXP_RETURN(!dump(out + "/kht_lines.bin", lines.data(), lines.size()), COMPV_ERROR_CODE_E_INVALID_STATE);
	CompVLineFloat32Vector cart;
	COMPV_CHECK_CODE_RETURN(hough->toCartesian(width, height, lines, cart));
	COMPV_CHECK_EXP_RETURN(!dump(out + "/kht_cartesian.f32", reinterpret_cast<const float*>(cart.data()), cart.size() * 6), COMPV_ERROR_CODE_E_INVALID_STATE);
	COMPV_CHECK_EXP_RETURN(hough->setFloat32(COMPV_HOUGH_SET_FLT32_RHO, 2.f) != COMPV_ERROR_CODE_E_INVALID_PARAMETER, COMPV_ERROR_CODE_E_INVALID_STATE); // houghkht.cxx:144

	// FAST9, threshold 20, NMS, every corner (unittests/feature_fast.cxx)
	CompVCornerDetePtr fast;
	COMPV_CHECK_CODE_RETURN(CompVCornerDete::newObj(&fast, COMPV_FAST_ID));
	COMPV_CHECK_CODE_RETURN(fast->setInt(COMPV_FAST_SET_INT_THRESHOLD, 20));
	COMPV_CHECK_CODE_RETURN(fast->setInt(COMPV_FAST_SET_INT_FAST_TYPE, COMPV_FAST_TYPE_9));
	COMPV_CHECK_CODE_RETURN(fast->setInt(COMPV_FAST_SET_INT_MAX_FEATURES, -1));
	COMPV_CHECK_CODE_RETURN(fast->setBool(COMPV_FAST_SET_BOOL_NON_MAXIMA_SUPP, true));
	CompVInterestPointVector points;
	COMPV_CHECK_CODE_RETURN(fast->process(image, points));
	COMPV_CHECK_EXP_RETURN(!dump(out + "/fast_points.bin", points.data(), points.size()), COMPV_ERROR_CODE_E_INVALID_STATE);

	// Otsu -> PLSL (samples/text_recognition/main.cxx:93-104 without the morphology step)
	double thr = 0;
	CompVMatPtr binar;
	COMPV_CHECK_CODE_RETURN(CompVImage::thresholdOtsu(image, thr, &binar));
	CompVConnectedComponentLabelingPtr ccl;
	COMPV_CHECK_CODE_RETURN(CompVConnectedComponentLabeling::newObj(&ccl, COMPV_PLSL_ID));
	CompVConnectedComponentLabelingResultPtr result;
	COMPV_CHECK_CODE_RETURN(ccl->process(binar, &result));
	CompVMatPtr labels;
	COMPV_CHECK_CODE_RETURN(result->debugFlatten(&labels));
	COMPV_CHECK_EXP_RETURN(!dump(out + "/plsl_labels.i32", labels->ptr<int32_t>(), width * height), COMPV_ERROR_CODE_E_INVALID_STATE);
	for (size_t j = 0; j < height; ++j) memcpy(&packed[j * width], binar->ptr<uint8_t>(j), width);
	COMPV_CHECK_EXP_RETURN(!dump(out + "/otsu.u8", packed.data(), packed.size()), COMPV_ERROR_CODE_E_INVALID_STATE);
	CompVConnectedComponentBoundingBoxesVector boxes;
	COMPV_CHECK_CODE_RETURN(result->boundingBoxes(boxes));
	COMPV_CHECK_EXP_RETURN(boxes.size() != result->labelsCount(), COMPV_ERROR_CODE_E_INVALID_STATE);
	CompVConnectedComponentPointsVector blobs, segs;
	COMPV_CHECK_CODE_RETURN(result->extract(blobs, COMPV_CCL_EXTRACT_TYPE_BLOB));
	COMPV_CHECK_CODE_RETURN(result->extract(segs, COMPV_CCL_EXTRACT_TYPE_SEGMENT));
	COMPV_CHECK_EXP_RETURN(blobs.size() != result->labelsCount() || segs.size() != blobs.size(), COMPV_ERROR_CODE_E_INVALID_STATE);
	{
		std::vector<int32_t> flat; // label, count, then (x, y) pairs, for the three largest-index labels
		for (size_t a = blobs.size() > 3 ? blobs.size() - 3 : 0; a < blobs.size(); ++a) {
			flat.push_back(static_cast<int32_t>(a + 1)); flat.push_back(static_cast<int32_t>(blobs[a].size()));
			for (size_t k = 0; k < blobs[a].size(); ++k) { flat.push_back(blobs[a][k].x); flat.push_back(blobs[a][k].y); }
		}
		COMPV_CHECK_EXP_RETURN(!dump(out + "/plsl_blobs.i32", flat.data(), flat.size()), COMPV_ERROR_CODE_E_INVALID_STATE);
	}
	for (int t = 0; t < 2; ++t) { // every label: count, then (x, y) pairs
		const CompVConnectedComponentPointsVector& v = t ? segs : blobs;
		std::vector<int32_t> flat;
		for (size_t a = 0; a < v.size(); ++a) {
			flat.push_back(static_cast<int32_t>(v[a].size()));
			for (size_t k = 0; k < v[a].size(); ++k) { flat.push_back(v[a][k].x); flat.push_back(v[a][k].y); }
		}
		COMPV_CHECK_EXP_RETURN(!dump(out + (t ? "/plsl_all_segments.i32" : "/plsl_all_blobs.i32"), flat.data(), flat.size()), COMPV_ERROR_CODE_E_INVALID_STATE);
	}
	// the text pipeline's clean-up step (samples/text_recognition/main.cxx:93-104)
	CompVMatPtr strel, closed;
	COMPV_CHECK_CODE_RETURN(CompVMathMorph::buildStructuringElement(&strel, CompVSizeSz(3, 3), COMPV_MATH_MORPH_STREL_TYPE_RECT));
	COMPV_CHECK_CODE_RETURN(CompVMathMorph::process(binar, strel, &closed, COMPV_MATH_MORPH_OP_TYPE_CLOSE));
	for (size_t j = 0; j < height; ++j) memcpy(&packed[j * width], closed->ptr<uint8_t>(j), width);
	COMPV_CHECK_EXP_RETURN(!dump(out + "/closed.u8", packed.data(), packed.size()), COMPV_ERROR_CODE_E_INVALID_STATE);
	for (size_t j = 0; j < height; ++j) memcpy(&packed[j * width], binar->ptr<uint8_t>(j), width);
	COMPV_CHECK_CODE_RETURN(ccl->process(binar, &result)); // the result object is reused (ccl_lsl.cxx:585-592)

	// MSER with the unit test's parameters (unittests/ccl_mser.cxx:26-46)
	CompVConnectedComponentLabelingPtr mser;
	COMPV_CHECK_CODE_RETURN(CompVConnectedComponentLabeling::newObj(&mser, COMPV_LMSER_ID, 2, (0.0055 * 0.0055), (0.8 * 0.15), 0.3, 0.2, 8));
	CompVConnectedComponentLabelingResultPtr mresult;
	COMPV_CHECK_CODE_RETURN(mser->process(image, &mresult));
	CompVConnectedComponentLabelingRegionMserVector regions;
	COMPV_CHECK_CODE_RETURN(mresult->points(regions));
	std::vector<int32_t> sizes;
	for (size_t i = 0; i < regions.size(); ++i) sizes.push_back(static_cast<int32_t>(regions[i].points.size()));
	COMPV_CHECK_EXP_RETURN(!dump(out + "/mser_sizes.i32", sizes.data(), sizes.size()), COMPV_ERROR_CODE_E_INVALID_STATE);
	COMPV_CHECK_EXP_RETURN(mresult->debugFlatten(&labels) != COMPV_ERROR_CODE_E_NOT_IMPLEMENTED, COMPV_ERROR_CODE_E_INVALID_STATE); // lmser_result.cxx:35-39

	// HOG-S, default geometry
	CompVHOGPtr hog;
	COMPV_CHECK_CODE_RETURN(CompVHOG::newObj(&hog, COMPV_HOGS_ID));
	CompVMatPtr desc;
	COMPV_CHECK_CODE_RETURN(hog->process(image, &desc));
	COMPV_CHECK_EXP_RETURN(!dump(out + "/hog.f32", desc->ptr<float>(), desc->cols()), COMPV_ERROR_CODE_E_INVALID_STATE);

	// parameter errors are the reference's
	CompVHoughPtr bad;
	COMPV_CHECK_EXP_RETURN(CompVHough::newObj(&bad, COMPV_HOUGHSHT_ID, 0.5f, 1.f, 1) != COMPV_ERROR_CODE_E_INVALID_PARAMETER, COMPV_ERROR_CODE_E_INVALID_STATE); // houghsht.cxx:312
	CompVMatPtr none;
	COMPV_CHECK_EXP_RETURN(dete->process(none, &edges) != COMPV_ERROR_CODE_E_INVALID_PARAMETER, COMPV_ERROR_CODE_E_INVALID_STATE);
	printf("api_check OK: %zu edge lines, %zu corners, otsu %.1f, %zu labels, %zu mser regions, hog %zu floats\n", lines.size(), points.size(), thr, result->labelsCount(), regions.size(),
		desc->cols());
	return COMPV_ERROR_CODE_S_OK;
}

int main(int argc, char** argv)
{
	if (argc != 5) { fprintf(stderr, "usage: api_check <width> <height> <frame.u8> <outdir>\n"); return 2; }
	const COMPV_ERROR_CODE rc = run(static_cast<size_t>(atoi(argv[1])), static_cast<size_t>(atoi(argv[2])), argv[3], argv[4]);
	if (rc != COMPV_ERROR_CODE_S_OK) { fprintf(stderr, "api_check FAILED: code %d (%s) cuda: %s\n", rc, cvb200_error_string(rc), cvb200_last_cuda_error()); return 1; }
	return 0;
}
