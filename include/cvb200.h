/*
 * cvb200.h -- C ABI of libcompv_b200.so: the B200-native (sm_100a) implementation of CompV's per-pixel hot path.
 *
 * Conventions (they mirror the reference so that a CompV maintainer can bind these 1:1):
 *   - every function returns an int that is a COMPV_ERROR_CODE numeric value
 *     (reference: base/include/compv/base/compv_common.h:226-279): 0 = S_OK, >= 20000 = error.
 *   - images are single-plane, row-major; `stride` is in SAMPLES (not bytes), exactly like CompVMat::stride()
 *     (reference: base/include/compv/base/compv_mat.h:21-588).
 *   - functions without suffix take caller-owned HOST buffers (the shape of the reference's dormant GPU hook
 *     CompVGpuCornerDeteFAST::processData, gpu/include/compv/gpu/core/features/fast/compv_gpu_feature_fast_dete.h:23-47)
 *     and perform H2D / kernels / D2H internally, synchronously.
 *   - functions with the `_dev` suffix take DEVICE pointers, a batch of `batch` frames spaced `framePitch` samples apart
 *     (0 -> stride*height) and a cudaStream_t passed as void*; they are asynchronous with respect to the host unless noted.
 *   - setters/getters use the reference's CompVCaps convention set(id, valuePtr, valueSize)
 *     (base/include/compv/base/compv_caps.h:15-35) with the same ids and the same size checks.
 *   - no function falls back to a CPU implementation: if CUDA is unavailable the call returns CVB200_E_CUDA.
 */
#ifndef CVB200_H_
#define CVB200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#  define CVB200_API __attribute__((visibility("default")))
#else
#  define CVB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes: numeric values of COMPV_ERROR_CODE (compv_common.h:226-279) ---- */
#define CVB200_S_OK                    0
#define CVB200_E_NOT_IMPLEMENTED       20001
#define CVB200_E_NOT_INITIALIZED       20002
#define CVB200_E_INVALID_CALL          20004
#define CVB200_E_INVALID_STATE         20005
#define CVB200_E_INVALID_PARAMETER     20006
#define CVB200_E_INVALID_SUBTYPE       20009
#define CVB200_E_OUT_OF_MEMORY         20013
#define CVB200_E_OUT_OF_BOUND          20014
#define CVB200_E_MEMORY_NOT_ALIGNED    20021
#define CVB200_E_CUDA                  20035

/* ---- feature ids and capability ids: numeric values of the anonymous enum in
 *      base/include/compv/base/compv_features.h:47-121 and base/include/compv/base/compv_ccl.h:61-103 ---- */
#define CVB200_FAST_ID                                  1
/* 8f-2: the ORB detector (core/features/orb/compv_core_feature_orb_dete.cxx): same object type and calls as FAST (cvb200_corner_dete_new(&d, CVB200_ORB_ID), _set, _process);
 * defaults of the reference (8 levels at 0.83, FAST9 threshold 20 + NMS, 2000 features, patch 31). Values = COMPV_ORB_* (compv_features.h:60-72). */
#define CVB200_ORB_ID 8
#define CVB200_ORB_SET_INT_INTERNAL_DETE_ID 9
#define CVB200_ORB_SET_INT_FAST_THRESHOLD 10
#define CVB200_ORB_SET_BOOL_FAST_NON_MAXIMA_SUPP 11
#define CVB200_ORB_SET_INT_PYRAMID_LEVELS 12
#define CVB200_ORB_SET_INT_PYRAMID_SCALE_TYPE 13
#define CVB200_ORB_SET_FLT32_PYRAMID_SCALE_FACTOR 14
#define CVB200_ORB_SET_INT_MAX_FEATURES 15
#define CVB200_FAST_SET_INT_THRESHOLD                   2
#define CVB200_FAST_SET_INT_MAX_FEATURES                3
#define CVB200_FAST_SET_INT_FAST_TYPE                   4
#define CVB200_FAST_SET_BOOL_NON_MAXIMA_SUPP            5
#define CVB200_FAST_TYPE_9                              6
#define CVB200_FAST_TYPE_12                             7
#define CVB200_CANNY_ID                                 20
#define CVB200_CANNY_SET_INT_KERNEL_SIZE                21
#define CVB200_CANNY_SET_INT_THRESHOLD_TYPE             22
#define CVB200_CANNY_SET_FLT32_THRESHOLD_LOW            23
#define CVB200_CANNY_SET_FLT32_THRESHOLD_HIGH           24
#define CVB200_CANNY_THRESHOLD_TYPE_PERCENT_OF_MEAN     25
#define CVB200_CANNY_THRESHOLD_TYPE_COMPARE_TO_GRADIENT 26
#define CVB200_SOBEL_ID                                 27
#define CVB200_SCHARR_ID                                28
#define CVB200_PREWITT_ID                               29
#define CVB200_HOUGHSHT_ID                              30
#define CVB200_HOUGHKHT_ID                              31
#define CVB200_HOUGH_SET_FLT32_RHO                      32
#define CVB200_HOUGH_SET_FLT32_THETA                    33
#define CVB200_HOUGH_SET_INT_THRESHOLD                  34
#define CVB200_HOUGH_SET_INT_MAXLINES                   35
#define CVB200_HOUGHKHT_SET_FLT32_CLUSTER_MIN_DEVIATION 36
#define CVB200_HOUGHKHT_SET_INT_CLUSTER_MIN_SIZE        37
#define CVB200_HOUGHKHT_SET_FLT32_KERNEL_MIN_HEIGTH     38
#define CVB200_HOUGHKHT_SET_BOOL_OVERRIDE_INPUT_EDGES   39
#define CVB200_HOUGHKHT_GET_FLT64_GS                    40
#define CVB200_HOGS_ID                                  41
#define CVB200_HOG_SET_BOOL_GRADIENT_SIGNED             44
#define CVB200_HOG_SET_INT_BLOCK_NORM                   45
#define CVB200_HOG_SET_INT_NBINS                        46
#define CVB200_HOG_SET_INT_INTERPOLATION                47
#define CVB200_HOG_BLOCK_NORM_NONE                      48
#define CVB200_HOG_BLOCK_NORM_L1                        49
#define CVB200_HOG_BLOCK_NORM_L1SQRT                    50
#define CVB200_HOG_BLOCK_NORM_L2                        51
#define CVB200_HOG_BLOCK_NORM_L2HYS                     52
#define CVB200_HOG_INTERPOLATION_NEAREST                53
#define CVB200_HOG_INTERPOLATION_BILINEAR_LUT           54
#define CVB200_HOG_INTERPOLATION_BILINEAR               55
#define CVB200_CCL_SET_INT_CONNECTIVITY                 0   /* compv_ccl.h:63-103 enum (separate id space) */
#define CVB200_PLSL_ID                                  1
#define CVB200_PLSL_SET_INT_TYPE                        2
#define CVB200_PLSL_SET_BOOL_SORT_SEGMENTS              3
#define CVB200_PLSL_TYPE_XRLEZ                          10
#define CVB200_LMSER_ID                                 19

/* Extension (not a reference id): Sobel/Scharr/Prewitt objects only, bool. When true, gmax is taken over the columns x with
 * x%8 in {0,1,2,4} -- what the reference's x86 SSE4.1 leaf CompVMathUtilsMax_16u_Intrin_SSE41 actually computes
 * (base/math/intrin/x86/compv_math_utils_intrin_sse41.cxx:55-63 folds only lanes 0,1,2,4). Default false = the true maximum,
 * i.e. the reference's plain C++ path (compv_math_utils.h:158-170). */
#define CVB200_EDGE_SET_BOOL_X86_SSE41_GMAX_LANES       1000

/* Extension, bool: force the generic (any kernel size / any stride) Canny front kernel instead of the TMA fast path. Test hook. */
#define CVB200_EDGE_SET_BOOL_GENERIC_KERNEL             1001

/* Extension, bool, Hough objects: reproduce what the reference's x86 SIMD build computes where it differs from its own scalar path (default true):
 * kernel-height association of the SSE2/AVX leaves and the peak-scan coverage of the SSE2 leaf (see oracle/compv_oracle_kht.cpp header). */
#define CVB200_HOUGH_SET_BOOL_X86_SIMD_SCAN             1002

/* COMPV_BORDER_TYPE (compv_common.h) */
#define CVB200_BORDER_TYPE_ZERO      0
#define CVB200_BORDER_TYPE_IGNORE    1
#define CVB200_BORDER_TYPE_REPLICATE 2

typedef void* cvb200_stream_t; /* cudaStream_t */

/* ================================================================================================
 * Runtime. Replaces CompVGpu::init's dlopen("libcuda")/cuInit probe (gpu/compv_gpu.cxx:36-62).
 * ============================================================================================== */
/* Binds the calling process to CUDA device `device` (>=0). Must be called before anything else. Idempotent. */
CVB200_API int cvb200_init(int device);
/* SURVEY 8(b) `cvb200_init(int device_count)`: initialises devices 0 .. device_count-1 (<= 0: every device of the process) for the *_multi batch entry points, which
 * spread the frames of a batch over them in-process (one worker thread, its streams and scratch per device; no torchrun needed). Device 0 serves every other entry point. */
CVB200_API int cvb200_init_devices(int device_count);
CVB200_API int cvb200_active_device_count(void);
CVB200_API int cvb200_deinit(void);
/* 1 when cvb200_init succeeded (reference: CompVGpu::isActiveAndEnabled, gpu/include/compv/gpu/compv_gpu.h:31-33) */
CVB200_API int cvb200_is_active(void);
CVB200_API int cvb200_device_count(int* count);
CVB200_API const char* cvb200_error_string(int code);
/* Last CUDA runtime error string seen by the library on this thread ("" if none) */
CVB200_API const char* cvb200_last_cuda_error(void);
/* Number of kernel launches issued by this library since load (all threads); used by bench.py's gpu_launches */
CVB200_API uint64_t cvb200_launch_count(void);
/* Per-kernel device timing: between _begin and _end every kernel launched by the library is bracketed by CUDA events on its stream.
 * _end synchronises the device and writes one "name count total_ms\n" line per kernel into buf. Used by bench.py's roofline block. */
CVB200_API int cvb200_profile_begin(void);
CVB200_API int cvb200_profile_end(char* buf, size_t bufSize);
/* Memory helpers so that a host without the CUDA runtime headers can own device / pinned buffers */
CVB200_API int cvb200_malloc(void** dptr, size_t bytes);
CVB200_API int cvb200_free(void* dptr);
CVB200_API int cvb200_host_alloc(void** hptr, size_t bytes);   /* pinned */
CVB200_API int cvb200_host_free(void* hptr);
CVB200_API int cvb200_memcpy_h2d(void* dptr, const void* hptr, size_t bytes, cvb200_stream_t stream);
CVB200_API int cvb200_memcpy_d2h(void* hptr, const void* dptr, size_t bytes, cvb200_stream_t stream);
CVB200_API int cvb200_memset(void* dptr, int value, size_t bytes, cvb200_stream_t stream);
CVB200_API int cvb200_stream_create(cvb200_stream_t* stream);
CVB200_API int cvb200_stream_destroy(cvb200_stream_t stream);
CVB200_API int cvb200_stream_sync(cvb200_stream_t stream);
/* Caps the host worker pool used by the few per-frame host stages that remain (SHT line ordering); 0 = one per core (at most 128).
 * A multi-process launch (one rank per GPU) sets it to cores / local ranks.  Reference: CompVBase::init(numThreads), base/compv_base.cxx. */
CVB200_API int cvb200_set_host_threads(int n);
/* Test hook (no device needed): out[i] += 1 for i in [0, n) through the host worker pool. */
CVB200_API int cvb200_selftest_host_pool(size_t n, unsigned int* out);
/* mismatches[0..1]: the HOG cells pass' division / square-root sequences against IEEE division / square root over all of their integer operands (test hook) */
CVB200_API int cvb200_selftest_hog_math(unsigned int* mismatches);

/* ================================================================================================
 * a2 -- separable convolution. Replaces CompVMathConvlt::convlt1<In,Kern,Out> / convlt1FixedPoint
 * (base/include/compv/base/math/compv_math_convlt.h:25-55, leaves :332-405). Correlation (no kernel flip),
 * horizontal pass first, intermediate stored as OutputType with the reference's saturation/truncation,
 * then vertical pass. borderType applies to the r=kernSize/2 outer ring: ZERO writes 0 there, REPLICATE copies
 * the (saturated) input there, IGNORE leaves it untouched.
 * kernSize must be odd, <= 63, <= width and <= height. out may not alias in.
 * ============================================================================================== */
CVB200_API int cvb200_convlt1_8u16s16s(const uint8_t* in, size_t width, size_t height, size_t stride, const int16_t* vtKern, const int16_t* hzKern, size_t kernSize, int16_t* out, int borderType);
CVB200_API int cvb200_convlt1_16s16s16s(const int16_t* in, size_t width, size_t height, size_t stride, const int16_t* vtKern, const int16_t* hzKern, size_t kernSize, int16_t* out, int borderType);
CVB200_API int cvb200_convlt1_8u32f8u(const uint8_t* in, size_t width, size_t height, size_t stride, const float* vtKern, const float* hzKern, size_t kernSize, uint8_t* out, int borderType);
CVB200_API int cvb200_convlt1_8u32f32f(const uint8_t* in, size_t width, size_t height, size_t stride, const float* vtKern, const float* hzKern, size_t kernSize, float* out, int borderType);
CVB200_API int cvb200_convlt1_32f32f32f(const float* in, size_t width, size_t height, size_t stride, const float* vtKern, const float* hzKern, size_t kernSize, float* out, int borderType);
CVB200_API int cvb200_convlt1_32f32f8u(const float* in, size_t width, size_t height, size_t stride, const float* vtKern, const float* hzKern, size_t kernSize, uint8_t* out, int borderType);
CVB200_API int cvb200_convlt1_fxp_8u16u8u(const uint8_t* in, size_t width, size_t height, size_t stride, const uint16_t* vtKern, const uint16_t* hzKern, size_t kernSize, uint8_t* out, int borderType);
/* device-pointer, batched, asynchronous variants (kernels are HOST pointers: they are passed by value to the launch) */
CVB200_API int cvb200_convlt1_8u16s16s_dev(const uint8_t* in, size_t width, size_t height, size_t stride, const int16_t* vtKern, const int16_t* hzKern, size_t kernSize, int16_t* out, int borderType, size_t batch, size_t framePitch, cvb200_stream_t stream);
CVB200_API int cvb200_convlt1_16s16s16s_dev(const int16_t* in, size_t width, size_t height, size_t stride, const int16_t* vtKern, const int16_t* hzKern, size_t kernSize, int16_t* out, int borderType, size_t batch, size_t framePitch, cvb200_stream_t stream);
CVB200_API int cvb200_convlt1_8u32f8u_dev(const uint8_t* in, size_t width, size_t height, size_t stride, const float* vtKern, const float* hzKern, size_t kernSize, uint8_t* out, int borderType, size_t batch, size_t framePitch, cvb200_stream_t stream);
CVB200_API int cvb200_convlt1_8u32f32f_dev(const uint8_t* in, size_t width, size_t height, size_t stride, const float* vtKern, const float* hzKern, size_t kernSize, float* out, int borderType, size_t batch, size_t framePitch, cvb200_stream_t stream);
CVB200_API int cvb200_convlt1_32f32f32f_dev(const float* in, size_t width, size_t height, size_t stride, const float* vtKern, const float* hzKern, size_t kernSize, float* out, int borderType, size_t batch, size_t framePitch, cvb200_stream_t stream);
CVB200_API int cvb200_convlt1_32f32f8u_dev(const float* in, size_t width, size_t height, size_t stride, const float* vtKern, const float* hzKern, size_t kernSize, uint8_t* out, int borderType, size_t batch, size_t framePitch, cvb200_stream_t stream);
CVB200_API int cvb200_convlt1_fxp_8u16u8u_dev(const uint8_t* in, size_t width, size_t height, size_t stride, const uint16_t* vtKern, const uint16_t* hzKern, size_t kernSize, uint8_t* out, int borderType, size_t batch, size_t framePitch, cvb200_stream_t stream);
/* CompVMathGauss::kernelDim1<float> / kernelDim1FixedPoint (base/include/compv/base/math/compv_math_gauss.h:23-56,
 * base/math/compv_math_gauss.cxx) -- host-side helpers (5 taps of arithmetic, no pixel work). */
CVB200_API int cvb200_gauss_kernel_dim1_32f(size_t size, float sigma, float* kernel);
CVB200_API int cvb200_gauss_kernel_dim1_fxp(size_t size, float sigma, uint16_t* kernel);

/* ================================================================================================
 * a3/a5 -- edge detectors. Replaces CompVEdgeDete::newObj(&d, id, tLow, tHigh, kernSize) + d->process(image,&edges)
 * (base/compv_features.cxx:146-161; Sobel/Scharr/Prewitt: core/features/edges/compv_core_feature_edge_dete.cxx:55-206;
 * Canny: core/features/edges/compv_core_feature_canny_dete.cxx:123-331).
 * id in {CVB200_SOBEL_ID, CVB200_SCHARR_ID, CVB200_PREWITT_ID, CVB200_CANNY_ID}.
 * ============================================================================================== */
typedef struct cvb200_edge_dete cvb200_edge_dete_t;
CVB200_API int cvb200_edge_dete_new(cvb200_edge_dete_t** dete, int id, float tLow, float tHigh, size_t kernSize);
CVB200_API int cvb200_edge_dete_free(cvb200_edge_dete_t** dete);
/* CompVCaps::set (canny_dete.cxx:77-117): CANNY_SET_INT_THRESHOLD_TYPE (int32), CANNY_SET_FLT32_THRESHOLD_LOW/HIGH (float), CANNY_SET_INT_KERNEL_SIZE (int) */
CVB200_API int cvb200_edge_dete_set(cvb200_edge_dete_t* dete, int id, const void* valuePtr, size_t valueSize);
/* Optional fused Gaussian pre-blur (BASELINE config 2: CompVMathGauss::kernelDim1<float>(size,sigma) + convlt1<u8,f32,u8> before process()).
 * size 0 disables. Bit-identical to running cvb200_convlt1_8u32f8u first, but the blurred frame never touches HBM. Canny only. */
CVB200_API int cvb200_edge_dete_set_preblur(cvb200_edge_dete_t* dete, size_t size, float sigma);
/* edges: height*stride bytes, same stride as image (canny_dete.cxx:249). edges == image is allowed for the host variant (canny_dete.cxx:122). */
CVB200_API int cvb200_edge_dete_process(cvb200_edge_dete_t* dete, const uint8_t* image, size_t width, size_t height, size_t stride, uint8_t* edges);
/* Many frames in host memory (pinned memory recommended): chunks are pipelined H2D / kernels / D2H on three streams. Synchronous. */
CVB200_API int cvb200_edge_dete_process_batch(cvb200_edge_dete_t* dete, const uint8_t* image, size_t width, size_t height, size_t stride, uint8_t* edges, size_t batch, size_t framePitch);
CVB200_API int cvb200_edge_dete_process_dev(cvb200_edge_dete_t* dete, const uint8_t* image, size_t width, size_t height, size_t stride, uint8_t* edges, size_t batch, size_t framePitch, cvb200_stream_t stream);
/* Intermediate gradient planes of the Canny/Sobel front end (K1+K4: convlt1<u8,int16,int16> x2 + CompVMathUtils::sumAbs,
 * canny_dete.cxx:236-240): gx, gy int16 and g uint16, each height*stride samples. Any of the three outputs may be NULL. */
CVB200_API int cvb200_sobel_g(const uint8_t* image, size_t width, size_t height, size_t stride, int id, size_t kernSize, int16_t* gx, int16_t* gy, uint16_t* g);
CVB200_API int cvb200_sobel_g_dev(const uint8_t* image, size_t width, size_t height, size_t stride, int id, size_t kernSize, int16_t* gx, int16_t* gy, uint16_t* g, size_t batch, size_t framePitch, cvb200_stream_t stream);

/* ================================================================================================
 * a8 -- FAST9/FAST12 corners. Replaces CompVCornerDete::newObj(&f, COMPV_FAST_ID) + f->process(image, points)
 * (core/features/fast/compv_core_feature_fast_dete.cxx:163-422; leaves :658-831, point list :490-585).
 * ============================================================================================== */
/* Binary layout of CompVInterestPoint (base/include/compv/base/compv_common.h:629-656): the adapter can memcpy into the std::vector */
typedef struct cvb200_interest_point {
	float x, y, strength, orient;
	int32_t level;
	float size;
} cvb200_interest_point_t;
typedef struct cvb200_corner_dete cvb200_corner_dete_t;
CVB200_API int cvb200_corner_dete_new(cvb200_corner_dete_t** dete, int id /* CVB200_FAST_ID */);
CVB200_API int cvb200_corner_dete_free(cvb200_corner_dete_t** dete);
/* fast_dete.cxx:128-160: FAST_SET_INT_THRESHOLD (int, clipped 0..255), FAST_SET_INT_MAX_FEATURES (int), FAST_SET_INT_FAST_TYPE (int: FAST_TYPE_9/12),
 * FAST_SET_BOOL_NON_MAXIMA_SUPP (bool). Defaults: threshold 20, FAST9, NMS on, maxFeatures 2000 (fast_dete.cxx:74-80). */
CVB200_API int cvb200_corner_dete_set(cvb200_corner_dete_t* dete, int id, const void* valuePtr, size_t valueSize);
/* Host frame in, points out in raster order (x, y, strength + threshold - 1). *count receives the number of points found (after selectBest when
 * maxFeatures > 1 applies, compv_common.h:641-655); if it exceeds `capacity` the first `capacity` points are written and E_OUT_OF_BOUND is returned.
 * capacity == 0 (points may be NULL) only counts. */
CVB200_API int cvb200_corner_dete_process(cvb200_corner_dete_t* dete, const uint8_t* image, size_t width, size_t height, size_t stride, cvb200_interest_point_t* points, size_t capacity, size_t* count);
/* Device frames in; points[frame*capacity + k] (device) in raster order, counts[frame] (device) = points found (may exceed capacity: excess dropped).
 * maxFeatures/selectBest is NOT applied here (it is a host-side std::nth_element in the reference). Asynchronous. */
CVB200_API int cvb200_corner_dete_process_dev(cvb200_corner_dete_t* dete, const uint8_t* image, size_t width, size_t height, size_t stride, cvb200_interest_point_t* points, size_t capacity, unsigned int* counts, size_t batch, size_t framePitch, cvb200_stream_t stream);
/* K11 strength map, the shape of the reference's dormant hook CompVGpuCornerDeteFAST::processData(IP, width, height, stride, N, threshold, strengths)
 * (gpu/include/compv/gpu/core/features/fast/compv_gpu_feature_fast_dete.h:23-47): strengths[y*stride+x] (before NMS), 0 on the 3-pixel border. N = 9 or 12. */
CVB200_API int cvb200_fast_scores(const uint8_t* image, size_t width, size_t height, size_t stride, int N, int threshold, uint8_t* strengths);
CVB200_API int cvb200_fast_scores_dev(const uint8_t* image, size_t width, size_t height, size_t stride, int N, int threshold, uint8_t* strengths, size_t batch, size_t framePitch, cvb200_stream_t stream);

/* ================================================================================================
 * a6/a7 -- Hough line detectors. Replaces CompVHough::newObj(&h, id, rho, theta, threshold) + h->process(edges, lines)
 * (base/compv_features.cxx:176-191; KHT: core/features/hough/compv_core_feature_houghkht.cxx:208-447; SHT: compv_core_feature_houghsht.cxx:96-262).
 * id in {CVB200_HOUGHKHT_ID, CVB200_HOUGHSHT_ID}. theta is the value the reference's newObj receives (KHT: degrees).
 * ============================================================================================== */
/* Binary layout of CompVHoughLine (base/include/compv/base/compv_common.h:686-692) */
typedef struct cvb200_hough_line {
	float rho;
	float theta;
	size_t strength;
} cvb200_hough_line_t;
typedef struct cvb200_hough cvb200_hough_t;
CVB200_API int cvb200_hough_new(cvb200_hough_t** hough, int id, float rho, float theta, size_t threshold);
CVB200_API int cvb200_hough_free(cvb200_hough_t** hough);
/* houghkht.cxx:140-192: HOUGH_SET_FLT32_RHO / _THETA (float), HOUGH_SET_INT_THRESHOLD / _MAXLINES (int), HOUGHKHT_SET_FLT32_CLUSTER_MIN_DEVIATION (float),
 * HOUGHKHT_SET_INT_CLUSTER_MIN_SIZE (int, >= 2 here), HOUGHKHT_SET_FLT32_KERNEL_MIN_HEIGTH (float), HOUGHKHT_SET_BOOL_OVERRIDE_INPUT_EDGES (bool, accepted, no effect) */
CVB200_API int cvb200_hough_set(cvb200_hough_t* hough, int id, const void* valuePtr, size_t valueSize);
/* houghkht.cxx:194-206: HOUGHKHT_GET_FLT64_GS -> double */
CVB200_API int cvb200_hough_get(cvb200_hough_t* hough, int id, void* valuePtr, size_t valueSize);
/* Host edge map (non-zero = edge) in; lines out, strongest first (the reference's order). *count = lines found (<= maxLines); at most `capacity` written. Synchronous. */
CVB200_API int cvb200_hough_process(cvb200_hough_t* hough, const uint8_t* edges, size_t width, size_t height, size_t stride, cvb200_hough_line_t* lines, size_t capacity, size_t* count);
/* Device edge maps (batch) in; lines[frame*capacity + k] and counts[frame] are HOST arrays. Everything, including the reference's std::sort tie order and the
 * sweep of the peak stage, runs on the device; the call returns when the lines are in host memory. */
CVB200_API int cvb200_hough_process_dev(cvb200_hough_t* hough, const uint8_t* edges, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch, cvb200_hough_line_t* lines, size_t capacity, size_t* counts, cvb200_stream_t stream);

/* ================================================================================================
 * a10 -- thresholding. Replaces CompVImageThreshold::global / otsu / adaptive (base/image/compv_image_threshold.cxx:52-317, reached through
 * CompVImage::thresholdGlobal / thresholdOtsu / thresholdAdaptive, base/include/compv/base/image/compv_image.h:63-67),
 * CompVMathHistogram::build for 8-bit data (base/math/compv_math_histogram.cxx:44-61) and CompVKernel::mean (base/compv_kernel.cxx:12-25).
 * ============================================================================================== */
CVB200_API int cvb200_histogram_8u(const uint8_t* in, size_t width, size_t height, size_t stride, unsigned int* hist /* [256] */);
CVB200_API int cvb200_histogram_8u_dev(const uint8_t* in, size_t width, size_t height, size_t stride, unsigned int* hist /* device [batch*256] */, size_t batch, size_t framePitch, cvb200_stream_t stream);
/* out = in > uint8(clip(threshold)+0.5) ? 255 : 0. out may equal in. */
CVB200_API int cvb200_threshold_global(const uint8_t* in, size_t width, size_t height, size_t stride, double threshold, uint8_t* out);
CVB200_API int cvb200_threshold_global_dev(const uint8_t* in, size_t width, size_t height, size_t stride, double threshold, uint8_t* out, size_t batch, size_t framePitch, cvb200_stream_t stream);
/* *threshold receives the Otsu threshold (an integer value, as a double like the reference); out may be NULL (threshold only). */
CVB200_API int cvb200_threshold_otsu(const uint8_t* in, size_t width, size_t height, size_t stride, double* threshold, uint8_t* out);
/* thresholds: device doubles [batch]; histScratch: device [batch*256] uint32 */
CVB200_API int cvb200_threshold_otsu_dev(const uint8_t* in, size_t width, size_t height, size_t stride, double* thresholds, uint8_t* out, unsigned int* histScratch, size_t batch, size_t framePitch, cvb200_stream_t stream);
/* mean kernel of `blockSize` (odd, <= 63) taps in fixed point, out = (in - mean > -delta) ? maxVal : 0 (inverted when invert != 0). out may equal in only for the host variant. */
CVB200_API int cvb200_threshold_adaptive(const uint8_t* in, size_t width, size_t height, size_t stride, size_t blockSize, double delta, double maxVal, int invert, uint8_t* out);
CVB200_API int cvb200_threshold_adaptive_dev(const uint8_t* in, size_t width, size_t height, size_t stride, size_t blockSize, double delta, double maxVal, int invert, uint8_t* out, size_t batch, size_t framePitch, cvb200_stream_t stream);
/* the overload taking separable fixed-point kernels (compv_image_threshold.cxx:200): kernels are HOST pointers */
CVB200_API int cvb200_threshold_adaptive_kernel_dev(const uint8_t* in, size_t width, size_t height, size_t stride, const uint16_t* kernelVt, const uint16_t* kernelHz, size_t kernSize, double delta, double maxVal, int invert, uint8_t* out, size_t batch, size_t framePitch, cvb200_stream_t stream);
CVB200_API int cvb200_kernel_mean_fxp(size_t blockSize, uint16_t* kernel);

/* ================================================================================================
 * a4 -- CompVGradientFast (base/compv_gradient_fast.cxx:58-433): gx = in[x+1]-in[x-1], gy = in[y+1]-in[y-1] (0 on the border), magnitude = sqrt(gx^2+gy^2)
 * (CompVMathTrig::hypot_naive), direction = CompVMathTrig::fastAtan2 in DEGREES [0,360]. Every output pointer may be NULL.
 * ============================================================================================== */
CVB200_API int cvb200_gradient_fast_8u(const uint8_t* in, size_t width, size_t height, size_t stride, int16_t* gx16, int16_t* gy16, float* gx32, float* gy32, float* magnitude, float* direction);
CVB200_API int cvb200_gradient_fast_8u_dev(const uint8_t* in, size_t width, size_t height, size_t stride, int16_t* gx16, int16_t* gy16, float* gx32, float* gy32, float* magnitude, float* direction, size_t batch, size_t framePitch, cvb200_stream_t stream);
CVB200_API int cvb200_gradient_fast_32f_dev(const float* in, size_t width, size_t height, size_t stride, float* gx32, float* gy32, float* magnitude, float* direction, size_t batch, size_t framePitch, cvb200_stream_t stream);

/* ================================================================================================
 * a9 -- S-HOG. Replaces CompVHOG::newObj(&hog, COMPV_HOGS_ID, blockSize, blockStride, cellSize, nbins, blockNorm, gradientSigned, interp) + hog->process(input, &output)
 * (base/compv_features.cxx:210-299; core/features/hog/compv_core_feature_hog_std.cxx:196-393). The whole image is one window.
 * ============================================================================================== */
typedef struct cvb200_hog cvb200_hog_t;
CVB200_API int cvb200_hog_new(cvb200_hog_t** hog, int id /* CVB200_HOGS_ID */, size_t blockW, size_t blockH, size_t strideW, size_t strideH, size_t cellW, size_t cellH, size_t nbins, int blockNorm, int gradientSigned, int interp);
CVB200_API int cvb200_hog_free(cvb200_hog_t** hog);
/* hog_std.cxx:124-178: HOG_SET_BOOL_GRADIENT_SIGNED (bool), HOG_SET_INT_BLOCK_NORM / _NBINS / _INTERPOLATION (int) */
CVB200_API int cvb200_hog_set(cvb200_hog_t* hog, int id, const void* valuePtr, size_t valueSize);
/* CompVHOG::descriptorSize (base/compv_features.cxx:274-299) */
CVB200_API int cvb200_hog_descriptor_size(cvb200_hog_t* hog, size_t width, size_t height, size_t* size);
/* *size = descriptor length; out (capacity floats) receives it (E_OUT_OF_BOUND when too small; out == NULL only queries the size) */
CVB200_API int cvb200_hog_process(cvb200_hog_t* hog, const uint8_t* in, size_t width, size_t height, size_t stride, float* out, size_t capacity, size_t* size);
CVB200_API int cvb200_hog_process_32f(cvb200_hog_t* hog, const float* in, size_t width, size_t height, size_t stride, float* out, size_t capacity, size_t* size);
/* out: device, descriptor_size floats per frame, frames back to back */
CVB200_API int cvb200_hog_process_dev(cvb200_hog_t* hog, const uint8_t* in, size_t width, size_t height, size_t stride, float* out, size_t batch, size_t framePitch, cvb200_stream_t stream);
CVB200_API int cvb200_hog_process_32f_dev(cvb200_hog_t* hog, const float* in, size_t width, size_t height, size_t stride, float* out, size_t batch, size_t framePitch, cvb200_stream_t stream);

/* ================================================================================================
 * a11 -- connected component labeling, Parallel Light Speed Labeling. Replaces CompVConnectedComponentLabeling::newObj(&ccl, COMPV_PLSL_ID) + ccl->process(binar, &result)
 * (base/compv_ccl.cxx:69-97; core/ccl/compv_core_ccl_lsl.cxx:579-751) and the result class CompVConnectedComponentLabelingResultLSLImpl
 * (core/include/compv/core/ccl/compv_core_ccl_lsl_result.h:43-83, core/ccl/compv_core_ccl_lsl_result.cxx).
 * The input must be binary (0x00 / 0x01 / 0xff, ccl_lsl.cxx:578: only bit 0 is looked at); 8-connectivity; labels 1..na exactly as the reference numbers them.
 * width, height <= 32767 (the reference stores positions as int16).
 * ============================================================================================== */
typedef struct cvb200_ccl cvb200_ccl_t;
typedef struct cvb200_ccl_result cvb200_ccl_result_t;
typedef struct cvb200_ccl_range { int32_t a; int16_t start; int16_t end; } cvb200_ccl_range_t;          /* == compv_ccl_range_t (ccl_lsl_result.h:32-36): label, [start, end) */
typedef struct cvb200_rect16 { int16_t left, top, right, bottom; } cvb200_rect16_t;                     /* == CompVRectInt16 (compv_common.h) */
CVB200_API int cvb200_ccl_new(cvb200_ccl_t** ccl, int id /* CVB200_PLSL_ID or CVB200_LMSER_ID */);
/* the full CompVConnectedComponentLabeling::newObj signature (compv_ccl.h:229-236); the MSER parameters are ignored by the PLSL */
CVB200_API int cvb200_ccl_new_ex(cvb200_ccl_t** ccl, int id, int delta, double minArea, double maxArea, double maxVariation, double minDiversity, int connectivity);
CVB200_API int cvb200_ccl_free(cvb200_ccl_t** ccl);
/* ccl_lsl.cxx:129-151: PLSL_SET_INT_TYPE (int, only XRLEZ), PLSL_SET_BOOL_SORT_SEGMENTS (bool); compv_ccl.cxx:25-40: CCL_SET_INT_CONNECTIVITY (int, 4 or 8; the LSL ignores it) */
CVB200_API int cvb200_ccl_set(cvb200_ccl_t* ccl, int id, const void* valuePtr, size_t valueSize);
/* binar: host. *result is reused when non-NULL (as the reference reuses *result of the same id, ccl_lsl.cxx:585-592), else created; free it with cvb200_ccl_result_free. */
CVB200_API int cvb200_ccl_process(cvb200_ccl_t* ccl, const uint8_t* binar, size_t width, size_t height, size_t stride, cvb200_ccl_result_t** result);
/* binar: device, `batch` frames. labels (device, may be NULL): batch x height x width int32, the flattened label image (debugFlatten); na (host, may be NULL): labels per frame;
 * results (host array of `batch` handles, may be NULL): filled like cvb200_ccl_process does. */
CVB200_API int cvb200_ccl_process_dev(cvb200_ccl_t* ccl, const uint8_t* binar, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch,
	int32_t* labels, int32_t* na, cvb200_ccl_result_t** results, cvb200_stream_t stream);
CVB200_API int cvb200_ccl_result_free(cvb200_ccl_result_t** result);
/* labelsCount() (ccl_lsl_result.cxx:45-48); labelIds() are 1..count */
CVB200_API size_t cvb200_ccl_result_labels_count(const cvb200_ccl_result_t* result);
/* vecLEA() in CSR form: rowOffsets has height + 1 entries, ranges[rowOffsets[j] .. rowOffsets[j+1]) are row j's segments left to right. Pointers stay valid until the result is reused or freed. */
CVB200_API int cvb200_ccl_result_segments(const cvb200_ccl_result_t* result, const uint32_t** rowOffsets, const cvb200_ccl_range_t** ranges, size_t* count);
/* debugFlatten (ccl_lsl_result.cxx:51-98): labels is height rows of labelsStride int32 (host) */
CVB200_API int cvb200_ccl_result_flatten(const cvb200_ccl_result_t* result, int32_t* labels, size_t labelsStride);
/* boundingBoxes (ccl_lsl_result.cxx:136-185): right is the exclusive end column, bottom the last row */
CVB200_API int cvb200_ccl_result_bounding_boxes(const cvb200_ccl_result_t* result, cvb200_rect16_t* boxes, size_t capacity, size_t* count);

/* ================================================================================================
 * a12 -- maximally stable extremal regions. Replaces CompVConnectedComponentLabeling::newObj(&ccl, COMPV_LMSER_ID, delta, min_area, max_area, max_variation, min_diversity, connectivity)
 * + ccl->process(gray, &result) (core/ccl/compv_core_ccl_lmser.cxx:148-410) and CompVConnectedComponentLabelingResultLMSER::points() / boundingBoxes() (compv_ccl.h:159-170).
 * Same object and calls as a11 (cvb200_ccl_new_ex, cvb200_ccl_process[_dev] with labels == NULL); the stride is part of the input: like the reference, pixel indices
 * i and i +- 1 are neighbours wherever both are pixels, so with stride == width the end of a row touches the start of the next (ccl_lmser.cxx:214-231).
 * The same REGIONS as the reference (same pixel sets, same boxes); they are returned sorted by (grey level, smallest pixel index) and the order of the points inside a
 * region is unspecified -- the reference's orders are those of its serial flood.
 * ============================================================================================== */
CVB200_API int cvb200_ccl_result_regions(const cvb200_ccl_result_t* result, const int32_t** sizes, const cvb200_rect16_t** boxes, const int16_t** points, size_t* regionCount, size_t* pointCount);

/* ================================================================================================
 * Section 8f "next" row 1 -- mathematical morphology. Replaces CompVMathMorph::buildStructuringElement / ::process (base/math/compv_math_morph.cxx:85-126):
 * erode / dilate = min / max over the non-zero cells of the structuring element, open = erode then dilate, close = dilate then erode, with the reference's
 * border handling (basicOper :128-240, addBordersVt :585-629 -- (strelHeight+1)/2 rows --, addBordersHz :631-694). 8-bit only, like the reference (:57 CompVMathMorphT).
 * strel: HOST pointer, strelHeight rows of strelStride bytes; out may not alias in (:142). borderType: CVB200_BORDER_TYPE_* (IGNORE leaves the border cells of `out` untouched).
 * ============================================================================================== */
#define CVB200_MATH_MORPH_STREL_TYPE_RECT    0   /* COMPV_MATH_MORPH_STREL_TYPE (compv_common.h:402-406) */
#define CVB200_MATH_MORPH_STREL_TYPE_DIAMOND 1
#define CVB200_MATH_MORPH_STREL_TYPE_CROSS   2
#define CVB200_MATH_MORPH_OP_TYPE_ERODE      0   /* COMPV_MATH_MORPH_OP_TYPE (compv_common.h:410-419); the others return E_NOT_IMPLEMENTED as in the reference (:119-122) */
#define CVB200_MATH_MORPH_OP_TYPE_DILATE     1
#define CVB200_MATH_MORPH_OP_TYPE_OPEN       2
#define CVB200_MATH_MORPH_OP_TYPE_CLOSE      3
CVB200_API int cvb200_morph_build_strel(uint8_t* strel, size_t width, size_t height, size_t strelStride, int type);
CVB200_API int cvb200_morph_process(const uint8_t* in, size_t width, size_t height, size_t stride, const uint8_t* strel, size_t strelWidth, size_t strelHeight, size_t strelStride,
	uint8_t* out, int opType, int borderType);
CVB200_API int cvb200_morph_process_dev(const uint8_t* in, size_t width, size_t height, size_t stride, const uint8_t* strel, size_t strelWidth, size_t strelHeight, size_t strelStride,
	uint8_t* out, int opType, int borderType, size_t batch, size_t framePitch, cvb200_stream_t stream);

/* ================================================================================================
 * 8e -- row-strip mode: one frame cut into horizontal strips, one per GPU (compv_b200/strips.py drives these over torch.distributed / NCCL).
 * The reference splits the same stages across threads by rows: convolution with overlap rows (compv_math_convlt.h:129-159), Canny NMS + hysteresis per row band
 * (canny_dete.cxx:175-234, 282-306), SHT accumulation into per-thread accumulators that are summed (houghsht.cxx:455-477).
 * ============================================================================================== */
/* Stages of one edge-detector call apart (see edges.cu). Canny: 1 = front -> class map (0 / 0x80 weak / 0xff strong), 2 = closure of the strong pixels inside the map as it
 * stands, 4 = weak -> 0. Sobel / Scharr / Prewitt: 1 = frame maximum into gmax[frame] (device uint32), 2 = normalisation with gmax[frame] as given. Synchronous. */
CVB200_API int cvb200_edge_dete_process_stages_dev(cvb200_edge_dete_t* dete, const uint8_t* image, size_t width, size_t height, size_t stride, uint8_t* edges, size_t batch, size_t framePitch, int stages, uint32_t* gmax, cvb200_stream_t stream);
/* SHT: number of int32 cells of the accumulator of a width x fullHeight frame; votes of a strip (stripHeight rows from row yOffset) into acc (device, zeroed or summed by the
 * caller afterwards: the kernel OVERWRITES acc with the strip's votes); lines from a (summed) accumulator. */
CVB200_API int cvb200_hough_sht_acc_size(cvb200_hough_t* hough, size_t width, size_t fullHeight, size_t* elems);
CVB200_API int cvb200_hough_sht_accumulate_dev(cvb200_hough_t* hough, const uint8_t* edges, size_t width, size_t stripHeight, size_t stride, size_t fullHeight, size_t yOffset, int32_t* acc, cvb200_stream_t stream);
CVB200_API int cvb200_hough_sht_lines_dev(cvb200_hough_t* hough, int32_t* acc, size_t width, size_t fullHeight, cvb200_hough_line_t* lines, size_t capacity, size_t* count, cvb200_stream_t stream);
/* Otsu's threshold from a (summed) 256-bin histogram: the scan of CompVImageThreshold::otsu (base/image/compv_image_threshold.cxx:52-100) on the host. */
CVB200_API int cvb200_otsu_threshold_from_histogram(const uint32_t* histogram256, size_t pixelCount, double* threshold);

/* ================================================================================================
 * 8f-3 -- device-side grayscale of camera frames. Replaces CompVImage::convertGrayscale (base/image/compv_image.cxx:687-692,
 * base/image/compv_image_conv_to_grayscale.cxx:35-93, compv_image_conv_rgbfamily.cxx:93-120,243-270,400-425).
 * pixelFormat = the reference's COMPV_SUBTYPE_PIXELS_* value (compv_common.h:347-367); stride in SAMPLES (pixels), as CompVMat::stride().
 * ============================================================================================== */
#define CVB200_SUBTYPE_PIXELS_RGB24 13
#define CVB200_SUBTYPE_PIXELS_BGR24 14
#define CVB200_SUBTYPE_PIXELS_RGBA32 15
#define CVB200_SUBTYPE_PIXELS_BGRA32 16
#define CVB200_SUBTYPE_PIXELS_ABGR32 17
#define CVB200_SUBTYPE_PIXELS_ARGB32 18
#define CVB200_SUBTYPE_PIXELS_RGB565LE 19
#define CVB200_SUBTYPE_PIXELS_RGB565BE 20
#define CVB200_SUBTYPE_PIXELS_BGR565LE 21
#define CVB200_SUBTYPE_PIXELS_BGR565BE 22
#define CVB200_SUBTYPE_PIXELS_Y 25
#define CVB200_SUBTYPE_PIXELS_NV12 26
#define CVB200_SUBTYPE_PIXELS_NV21 27
#define CVB200_SUBTYPE_PIXELS_YUV420P 28
#define CVB200_SUBTYPE_PIXELS_YVU420P 29
#define CVB200_SUBTYPE_PIXELS_YUV422P 30
#define CVB200_SUBTYPE_PIXELS_YUYV422 31
#define CVB200_SUBTYPE_PIXELS_UYVY422 32
#define CVB200_SUBTYPE_PIXELS_YUV444P 33
/* Bytes per sample of the plane that carries luma (3, 4, 2 for the packed formats; 1 for Y / planar / semi-planar YUV, whose Y plane is all that is read).
 * E_NOT_IMPLEMENTED for the formats the reference cannot convert to gray either (ABGR32, HSV, HSL). */
CVB200_API int cvb200_image_bytes_per_sample(int pixelFormat, size_t* bytesPerSample);
/* Host frame in, host gray plane out, same stride in samples (conv_to_grayscale.cxx:213). */
CVB200_API int cvb200_image_to_grayscale(int pixelFormat, const uint8_t* data, size_t width, size_t height, size_t stride, uint8_t* gray);
/* Device to device, batched: frame f starts at data + f*framePitchBytes (0 = stride*bytesPerSample*height), its gray plane at gray + f*grayPitch (0 = grayStride*height). */
CVB200_API int cvb200_image_to_grayscale_dev(int pixelFormat, const uint8_t* data, size_t width, size_t height, size_t stride, uint8_t* gray, size_t grayStride,
	size_t batch, size_t framePitchBytes, size_t grayPitch, cvb200_stream_t stream);

/* Headline pipeline on host buffers: (optional fused Gaussian) Canny then Hough on `batch` frames; the edge maps stay on the device, only lines return.
 * Equivalent to cvb200_edge_dete_process + cvb200_hough_process per frame (the two calls samples/hough_lines/main.cxx:59,106 makes), pipelined H2D/compute. */
CVB200_API int cvb200_canny_kht_process_batch(cvb200_edge_dete_t* canny, cvb200_hough_t* hough, const uint8_t* images, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch, cvb200_hough_line_t* lines, size_t capacity, size_t* counts);
/* The same call on frames in a camera format (pixelFormat = COMPV_SUBTYPE_PIXELS_*): each frame crosses PCIe once, as it is -- for Y / NV12 / NV21 / I420 / YV12 / 4:2:2 / 4:4:4
 * planar only its Y plane -- and is made gray on the device (cvb200_image_to_grayscale_dev) in front of the Canny kernels; the reference converts on the CPU first
 * (CompVImage::convertGrayscale, samples/hough_lines/main.cxx). framePitchBytes = distance between frames in bytes (whole frame incl. chroma planes); stride in samples. */
CVB200_API int cvb200_canny_kht_process_batch_fmt(cvb200_edge_dete_t* canny, cvb200_hough_t* hough, int pixelFormat, const uint8_t* frames, size_t width, size_t height, size_t stride, size_t batch, size_t framePitchBytes, cvb200_hough_line_t* lines, size_t capacity, size_t* counts);
/* The same call spread over every device initialised by cvb200_init_devices: frames are independent, the batch is cut into one contiguous shard per device and each
 * shard runs the pipeline above on its device from its own host thread (no collective). Results land in the caller's arrays exactly as with one device. */
CVB200_API int cvb200_canny_kht_process_batch_multi(cvb200_edge_dete_t* canny, cvb200_hough_t* hough, const uint8_t* images, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch, cvb200_hough_line_t* lines, size_t capacity, size_t* counts);
/* Same pipeline on frames that are already in device memory (`images` is a device pointer; `lines` / `counts` are host arrays). The work is ordered after what was
 * queued on `stream` before the call; the call returns when the lines are in host memory. Sub-batches of the frames run on several internal streams so that the
 * linking stage of one sub-batch overlaps the other stages of its neighbours (environment: CVB200_PIPE_SUB frames per sub-batch, CVB200_PIPE_SLOTS in flight). */
CVB200_API int cvb200_canny_kht_process_batch_dev(cvb200_edge_dete_t* canny, cvb200_hough_t* hough, const uint8_t* images, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch, cvb200_hough_line_t* lines, size_t capacity, size_t* counts, cvb200_stream_t stream);

#ifdef __cplusplus
}
#endif

#endif /* CVB200_H_ */
