// SURVEY 8f-3 -- the step before every stage: camera frames arrive as YUV / RGB, the detectors want 8-bit gray.
// Replaces CompVImage::convertGrayscale (base/image/compv_image.cxx:687-692) -> CompVImageConvToGrayscale::process (base/image/compv_image_conv_to_grayscale.cxx:35-93):
//   * RGB family (compv_image_conv_rgbfamily.cxx:93-120, 243-270, 400-425; coefficients compv_image_conv_common.cxx:20-41):
//       Y = clampPixel8((33 R + 65 G + 13 B) >> 7) + 16), 565 samples widened to 8 bits by bit replication;
//   * packed YUV 4:2:2 (compv_image_conv_to_grayscale.cxx:260-280): the Y byte of every sample pair;
//   * planar / semi-planar YUV: the Y plane as it is (the reference re-wraps it, :57-78).
// One pass, device to device: raw frame in (bytes per pixel 1..4), gray plane out.  With it a frame crosses PCIe once in its native format -- for the planar
// formats only the Y plane has to be uploaded at all -- instead of being converted on the CPU first.  HBM: bpp B/px read + 1 B/px written.
#include "common.cuh"

namespace cvb {

struct GrayParams {
	const uint8_t* in; uint8_t* out;
	int W, H;
	size_t inStrideBytes, inPitchBytes, outStride, outPitch;
	int bpp;                // bytes per input sample
	int o0, o1, o2;         // byte offsets of the three colour channels inside a sample (RGB 24/32) or -1
	int c0, c1, c2;         // their coefficients
	int mode;               // 0 = RGB 24/32, 1 = 565 little endian, 2 = 565 big endian, 3 = packed 4:2:2 (Y at byte o0 of each 2-byte sample), 4 = plane copy
};

__device__ __forceinline__ uint8_t gray_of(const uint8_t* s, const GrayParams& p)
{
	if (p.mode == 0) {
		const int v = (((p.c0 * s[p.o0]) + (p.c1 * s[p.o1]) + (p.c2 * s[p.o2])) >> 7) + 16;
		return static_cast<uint8_t>(v > 255 ? 255 : v);
	}
	if (p.mode <= 2) {
		unsigned int k = (p.mode == 1) ? (s[0] | (s[1] << 8)) : ((s[0] << 8) | s[1]);
		unsigned int r = (k & 0xF800u) >> 8; r |= (r >> 5);
		unsigned int g = (k & 0x07E0u) >> 3; g |= (g >> 6);
		unsigned int b = (k & 0x001Fu) << 3; b |= (b >> 5);
		const int v = static_cast<int>(((p.c0 * r) + (p.c1 * g) + (p.c2 * b)) >> 7) + 16;
		return static_cast<uint8_t>(v > 255 ? 255 : v);
	}
	return s[p.o0];
}

// 4 output pixels per thread (one 32-bit store where the output row allows it)
__global__ void __launch_bounds__(256) to_gray_kernel(const GrayParams p)
{
	const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
	const int y = blockIdx.y;
	if (x0 >= p.W) return;
	const uint8_t* src = p.in + blockIdx.z * p.inPitchBytes + static_cast<size_t>(y) * p.inStrideBytes + static_cast<size_t>(x0) * p.bpp;
	uint8_t* dst = p.out + blockIdx.z * p.outPitch + static_cast<size_t>(y) * p.outStride + x0;
	if (x0 + 4 <= p.W && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
		unsigned int v = 0;
#pragma unroll
		for (int k = 0; k < 4; ++k) v |= static_cast<unsigned int>(gray_of(src + k * p.bpp, p)) << (8 * k);
		*reinterpret_cast<unsigned int*>(dst) = v;
	}
	else {
		for (int k = 0; k < 4 && x0 + k < p.W; ++k) dst[k] = gray_of(src + k * p.bpp, p);
	}
}

static int gray_params(int fmt, GrayParams* p)
{
	p->o0 = p->o1 = p->o2 = 0; p->c0 = 33; p->c1 = 65; p->c2 = 13; p->mode = 0;
	switch (fmt) {
	case CVB200_SUBTYPE_PIXELS_RGB24: p->bpp = 3; p->o0 = 0; p->o1 = 1; p->o2 = 2; return CVB200_S_OK;
	case CVB200_SUBTYPE_PIXELS_BGR24: p->bpp = 3; p->o0 = 0; p->o1 = 1; p->o2 = 2; p->c0 = 13; p->c2 = 33; return CVB200_S_OK;
	case CVB200_SUBTYPE_PIXELS_RGBA32: p->bpp = 4; p->o0 = 0; p->o1 = 1; p->o2 = 2; return CVB200_S_OK;
	case CVB200_SUBTYPE_PIXELS_BGRA32: p->bpp = 4; p->o0 = 0; p->o1 = 1; p->o2 = 2; p->c0 = 13; p->c2 = 33; return CVB200_S_OK;
	case CVB200_SUBTYPE_PIXELS_ARGB32: p->bpp = 4; p->o0 = 1; p->o1 = 2; p->o2 = 3; return CVB200_S_OK;
	case CVB200_SUBTYPE_PIXELS_RGB565LE: p->bpp = 2; p->mode = 1; return CVB200_S_OK;
	case CVB200_SUBTYPE_PIXELS_RGB565BE: p->bpp = 2; p->mode = 2; return CVB200_S_OK;
	case CVB200_SUBTYPE_PIXELS_BGR565LE: p->bpp = 2; p->mode = 1; p->c0 = 13; p->c2 = 33; return CVB200_S_OK;
	case CVB200_SUBTYPE_PIXELS_BGR565BE: p->bpp = 2; p->mode = 2; p->c0 = 13; p->c2 = 33; return CVB200_S_OK;
	case CVB200_SUBTYPE_PIXELS_YUYV422: p->bpp = 2; p->mode = 3; p->o0 = 0; return CVB200_S_OK;
	case CVB200_SUBTYPE_PIXELS_UYVY422: p->bpp = 2; p->mode = 3; p->o0 = 1; return CVB200_S_OK;
	case CVB200_SUBTYPE_PIXELS_Y: case CVB200_SUBTYPE_PIXELS_NV12: case CVB200_SUBTYPE_PIXELS_NV21: case CVB200_SUBTYPE_PIXELS_YUV420P:
	case CVB200_SUBTYPE_PIXELS_YVU420P: case CVB200_SUBTYPE_PIXELS_YUV422P: case CVB200_SUBTYPE_PIXELS_YUV444P:
		p->bpp = 1; p->mode = 4; p->o0 = 0; return CVB200_S_OK; // the Y plane comes first in all of them
	default:
		return CVB200_E_NOT_IMPLEMENTED; // compv_image_conv_to_grayscale.cxx:88-91
	}
}

} // namespace cvb

using namespace cvb;

extern "C" {

int cvb200_image_bytes_per_sample(int pixelFormat, size_t* bytesPerSample)
{
	CVB_REQUIRE(bytesPerSample, CVB200_E_INVALID_PARAMETER);
	GrayParams p;
	CVB_CHECK(gray_params(pixelFormat, &p));
	*bytesPerSample = static_cast<size_t>(p.bpp);
	return CVB200_S_OK;
}

int cvb200_image_to_grayscale_dev(int pixelFormat, const uint8_t* data, size_t width, size_t height, size_t stride, uint8_t* gray, size_t grayStride,
	size_t batch, size_t framePitchBytes, size_t grayPitch, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(data && gray && width && height && stride >= width && grayStride >= width && data != gray, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(width <= 0x3fffffff && height <= 65535 && batch <= 65535, CVB200_E_OUT_OF_BOUND);
	if (!batch) return CVB200_S_OK;
	GrayParams p;
	memset(&p, 0, sizeof(p));
	CVB_CHECK(gray_params(pixelFormat, &p));
	p.in = data; p.out = gray; p.W = static_cast<int>(width); p.H = static_cast<int>(height);
	p.inStrideBytes = stride * p.bpp;
	p.inPitchBytes = framePitchBytes ? framePitchBytes : p.inStrideBytes * height;
	p.outStride = grayStride;
	p.outPitch = grayPitch ? grayPitch : grayStride * height;
	dim3 grid(static_cast<unsigned>(div_up(div_up(width, 4), 256)), static_cast<unsigned>(height), static_cast<unsigned>(batch));
	{
		KernelScope ks_("to_gray", as_stream(stream));
		to_gray_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
	}
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

// Host frame in, host gray plane out (same stride in samples, like the reference: "in and out images must have same stride", conv_to_grayscale.cxx:213)
int cvb200_image_to_grayscale(int pixelFormat, const uint8_t* data, size_t width, size_t height, size_t stride, uint8_t* gray)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(data && gray && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	size_t bpp = 0;
	CVB_CHECK(cvb200_image_bytes_per_sample(pixelFormat, &bpp));
	DevBuf dIn, dOut;
	int rc = dIn.ensure(stride * bpp * height);
	if (!rc) rc = dOut.ensure(stride * height);
	if (!rc) rc = cvb200_memcpy_h2d(dIn.p, data, stride * bpp * height, nullptr);
	if (!rc) rc = cvb200_image_to_grayscale_dev(pixelFormat, dIn.as<uint8_t>(), width, height, stride, dOut.as<uint8_t>(), stride, 1, 0, 0, nullptr);
	if (!rc) { cudaError_t e = cudaMemcpy2DAsync(gray, stride, dOut.p, stride, width, height, cudaMemcpyDeviceToHost, 0); if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy2DAsync", __FILE__, __LINE__); }
	if (!rc) rc = cvb200_stream_sync(nullptr);
	dIn.release(); dOut.release();
	return rc;
}

} // extern "C"
