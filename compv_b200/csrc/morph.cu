// Section 8f "next" row 1 -- mathematical morphology (erode / dilate / open / close), 8-bit.
// Replaces CompVMathMorph::process (base/math/compv_math_morph.cxx:95-126): basicOper :128-240 (the reference gathers one input pointer per non-zero cell of the
// structuring element, :471-511, and takes the running min / max of those rows), openCloseOper :242-337, borders :585-694, buildStructuringElementGeneric :513-583.
//
// One launch per basic operation: a (64 + sw - 1) x (32 + sh - 1) input tile is staged in shared memory; a full rectangle (the common element) is reduced
// separably -- running min / max along rows into a second tile, then along columns: sw + sh operations per pixel instead of sw * sh --, any other element
// walks the list of its non-zero cells.  The reference's border rule is applied in the same kernel, so a basic operation reads the frame once and writes it once.
#include "common.cuh"
#include "tma.cuh"

#include <cstring>
#include <vector>

namespace cvb {

constexpr int MORPH_TW = 64;
constexpr int MORPH_TH = 32;
constexpr int MORPH_THREADS = 256;

struct MorphParams {
	int W, H;
	size_t stride, framePitch;
	int sw, sh, rw, rh, bh; // element size, half sizes, vertical border height (sh + 1) / 2
	int border;
	int nTaps;          // 0: full rectangle (separable path)
};

template <bool ERODE> __device__ __forceinline__ int morph_op(int a, int b) { return ERODE ? min(a, b) : max(a, b); }

template <bool ERODE>
__global__ void __launch_bounds__(MORPH_THREADS) morph_basic_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const short2* __restrict__ taps, MorphParams p)
{
	extern __shared__ uint8_t smem[];
	const int tw = MORPH_TW + p.sw - 1, th = MORPH_TH + p.sh - 1;
	uint8_t* sIn = smem;
	uint8_t* sMid = smem + ((tw * th + 15) & ~15); // MORPH_TW x th (rect path)
	const int x0 = blockIdx.x * MORPH_TW, y0 = blockIdx.y * MORPH_TH; // first output column / row of the tile
	in += blockIdx.z * p.framePitch;
	out += blockIdx.z * p.framePitch;
	const int tid = threadIdx.x;
	// stage rows y0 - rh ..., columns x0 - rw ...; cells outside the image are never used by a computed output
	for (int ly = tid >> 5; ly < th; ly += MORPH_THREADS / 32) {
		const int gy = y0 - p.rh + ly;
		const bool rowIn = (gy >= 0 && gy < p.H);
		const uint8_t* src = in + static_cast<size_t>(rowIn ? gy : 0) * p.stride;
		for (int lx = tid & 31; lx < tw; lx += 32) {
			const int gx = x0 - p.rw + lx;
			sIn[ly * tw + lx] = (rowIn && gx >= 0 && gx < p.W) ? src[gx] : 0;
		}
	}
	__syncthreads();
	if (p.nTaps == 0) { // rectangle: rows first
		for (int i = tid; i < MORPH_TW * th; i += MORPH_THREADS) {
			const int ly = i >> 6, lx = i & 63;
			const uint8_t* s = &sIn[ly * tw + lx];
			int v = s[0];
			for (int k = 1; k < p.sw; ++k) v = morph_op<ERODE>(v, s[k]);
			sMid[i] = static_cast<uint8_t>(v);
		}
		__syncthreads();
	}
	for (int i = tid; i < MORPH_TW * MORPH_TH; i += MORPH_THREADS) {
		const int ly = i >> 6, lx = i & 63;
		const int gx = x0 + lx, gy = y0 + ly;
		if (gx >= p.W || gy >= p.H) continue;
		const size_t o = static_cast<size_t>(gy) * p.stride + gx;
		// the reference computes the interior, then overwrites the border rows (addBordersVt), then the border columns (addBordersHz)
		const bool borderCell = (gx < p.rw || gx >= p.W - p.rw) || (gy < p.bh || gy >= p.H - p.bh);
		if (borderCell && p.border != CVB200_BORDER_TYPE_IGNORE) {
			out[o] = (p.border == CVB200_BORDER_TYPE_ZERO) ? 0 : sIn[(ly + p.rh) * tw + lx + p.rw];
			continue;
		}
		if (gx < p.rw || gx >= p.W - p.rw || gy < p.rh || gy >= p.H - p.rh) continue; // not computed, IGNORE: left as is
		int v;
		if (p.nTaps == 0) {
			const uint8_t* s = &sMid[ly * MORPH_TW + lx];
			v = s[0];
			for (int k = 1; k < p.sh; ++k) v = morph_op<ERODE>(v, s[k * MORPH_TW]);
		}
		else {
			v = ERODE ? 255 : 0;
			const uint8_t* s = &sIn[ly * tw + lx];
			for (int k = 0; k < p.nTaps; ++k) { const short2 t = taps[k]; v = morph_op<ERODE>(v, s[t.y * tw + t.x]); }
		}
		out[o] = static_cast<uint8_t>(v);
	}
}

// ---- fast path: full rectangles of 3 or 5 rows / columns (the elements the text pipeline uses) ----
// Same staging as convlt_fast.cuh (one TMA bulk-tensor copy per CTA, 4 pixels per lane); min / max are byte-wise SIMD-in-a-word (__vminu4 / __vmaxu4) on the packed
// pixels: the horizontal neighbours of a word are two byte permutes of it and its neighbours, so a 3x3 basic operation costs about a dozen instructions per 4 pixels.
constexpr int MF_TW = 120, MF_TH = 60, MF_THREADS = 256, MF_WARPS = 8, MF_ROWW = 32, MF_INW = 36;

template <bool ERODE> __device__ __forceinline__ unsigned int morph_op4(unsigned int a, unsigned int b) { return ERODE ? __vminu4(a, b) : __vmaxu4(a, b); }

template <bool ERODE, int SW, int SH>
__global__ void __launch_bounds__(MF_THREADS, 3)
morph_rect_fast_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t* __restrict__ outAll, int W, int H, size_t stride, size_t framePitch, int border, int vecStore)
{
	constexpr int RW = SW >> 1, RH = SH >> 1, BH = (SH + 1) >> 1;
	constexpr int IN_ROWS = MF_TH + 2 * RH;
	extern __shared__ __align__(128) unsigned char mf_smem[];
	const unsigned int pad = (128u - (static_cast<unsigned int>(__cvta_generic_to_shared(mf_smem)) & 127u)) & 127u;
	unsigned int* sA = reinterpret_cast<unsigned int*>(mf_smem + pad);
	unsigned int* sM = sA + IN_ROWS * MF_INW + 4;
	uint64_t* bar = reinterpret_cast<uint64_t*>(sM + IN_ROWS * MF_ROWW + 2);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int x0 = blockIdx.x * MF_TW, y0 = blockIdx.y * MF_TH, frame = blockIdx.z;
	const int xl = x0 - 4 + 4 * lane;
	const int yIn0 = y0 - RH;
	const int xTma = (x0 - 4) & ~15;
	const int woff = ((x0 - 4) - xTma) >> 2;
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		mbar_expect_tx(bar, IN_ROWS * MF_INW * 4);
		tma_load_3d(sA, &tmap, bar, xTma, yIn0, frame);
	}
	__syncthreads();
	mbar_wait(bar, 0);
	// rows: for each of my 4 pixels the min / max over columns x-RW .. x+RW (bytes shifted in from the neighbouring words)
	for (int r = warp; r < IN_ROWS; r += MF_WARPS) {
		const unsigned int* q = &sA[r * MF_INW + woff + lane];
		const unsigned int wl = q[-1], wc = q[0], wr = q[1];
		unsigned int v = wc;
		v = morph_op4<ERODE>(v, __byte_perm(wl, wc, 0x6543)); // x-1
		v = morph_op4<ERODE>(v, __byte_perm(wc, wr, 0x4321)); // x+1
		if (RW == 2) {
			v = morph_op4<ERODE>(v, __byte_perm(wl, wc, 0x5432)); // x-2
			v = morph_op4<ERODE>(v, __byte_perm(wc, wr, 0x5432)); // x+2
		}
		sM[r * MF_ROWW + lane] = v;
	}
	__syncthreads();
	unsigned int inImage = 0, colBand = 0;
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		if (xl + i < W) inImage |= 0xffu << (8 * i);
		if (xl + i < RW || xl + i >= W - RW) colBand |= 0xffu << (8 * i);
	}
	const bool laneOut = (lane >= 1 && lane <= 30) && xl < W;
	uint8_t* __restrict__ out = outAll + frame * framePitch;
	for (int ro = warp; ro < MF_TH; ro += MF_WARPS) {
		const int y = y0 + ro;
		if (y >= H) break; // warp-uniform
		if (!laneOut) continue;
		unsigned int v = sM[ro * MF_ROWW + lane];
#pragma unroll
		for (int k = 1; k < SH; ++k) v = morph_op4<ERODE>(v, sM[(ro + k) * MF_ROWW + lane]);
		const bool rowComputed = (y >= RH && y < H - RH), rowBand = (y < BH || y >= H - BH);
		// the reference computes the interior, then overwrites the border rows (addBordersVt), then the border columns (addBordersHz)
		const unsigned int band = rowBand ? 0xffffffffu : colBand;
		const unsigned int computed = rowComputed ? ~colBand : 0u;
		unsigned int value, storeMask;
		if (border == CVB200_BORDER_TYPE_IGNORE) { value = v; storeMask = computed & inImage; }
		else {
			const unsigned int b = (border == CVB200_BORDER_TYPE_ZERO) ? 0u : sA[(ro + RH) * MF_INW + woff + lane];
			value = (b & band) | (v & ~band);
			storeMask = inImage;
		}
		if (!storeMask) continue;
		uint8_t* o8 = out + static_cast<size_t>(y) * stride + xl;
		if (vecStore && storeMask == 0xffffffffu) *reinterpret_cast<unsigned int*>(o8) = value;
		else {
#pragma unroll
			for (int i = 0; i < 4; ++i) if ((storeMask >> (8 * i)) & 0xffu) o8[i] = static_cast<uint8_t>(value >> (8 * i));
		}
	}
}

// CVB200_S_OK when the fast path ran, 1 when it does not apply (element, alignment): the caller takes the generic kernel
template <bool ERODE, int SW, int SH>
static int morph_rect_fast_launch_t(const uint8_t* in, uint8_t* out, const MorphParams& p, size_t batch, cudaStream_t stream)
{
	constexpr int IN_ROWS = MF_TH + 2 * (SH >> 1);
	alignas(64) CUtensorMap map;
	memset(&map, 0, sizeof(map));
	if (!make_u8_tile_map(&map, in, p.W, p.H, p.stride, p.framePitch, batch, MF_INW * 4, IN_ROWS)) return 1;
	const size_t smem = (static_cast<size_t>(IN_ROWS) * (MF_INW + MF_ROWW) + 8) * 4 + 128 + 16;
	dim3 grid(static_cast<unsigned>(div_up(p.W, MF_TW)), static_cast<unsigned>(div_up(p.H, MF_TH)), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	const int vecStore = (((reinterpret_cast<uintptr_t>(out) | p.stride | p.framePitch) & 3) == 0) ? 1 : 0;
	{ KernelScope ks_(ERODE ? "morph_erode" : "morph_dilate", stream);
	  morph_rect_fast_kernel<ERODE, SW, SH><<<grid, MF_THREADS, smem, stream>>>(map, out, p.W, p.H, p.stride, p.framePitch, p.border, vecStore); }
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

template <bool ERODE>
static int morph_rect_fast_launch(const uint8_t* in, uint8_t* out, const MorphParams& p, size_t batch, cudaStream_t stream)
{
	if (p.nTaps != 0) return 1;
	if (p.sw == 3 && p.sh == 3) return morph_rect_fast_launch_t<ERODE, 3, 3>(in, out, p, batch, stream);
	if (p.sw == 5 && p.sh == 5) return morph_rect_fast_launch_t<ERODE, 5, 5>(in, out, p, batch, stream);
	if (p.sw == 3 && p.sh == 5) return morph_rect_fast_launch_t<ERODE, 3, 5>(in, out, p, batch, stream);
	if (p.sw == 5 && p.sh == 3) return morph_rect_fast_launch_t<ERODE, 5, 3>(in, out, p, batch, stream);
	return 1;
}

static int morph_launch(const uint8_t* in, uint8_t* out, const short2* dTaps, const MorphParams& p, bool erode, size_t batch, cudaStream_t stream)
{
	{
		const int rc = erode ? morph_rect_fast_launch<true>(in, out, p, batch, stream) : morph_rect_fast_launch<false>(in, out, p, batch, stream);
		if (rc != 1) return rc;
	}
	const size_t tw = MORPH_TW + p.sw - 1, th = MORPH_TH + p.sh - 1;
	const size_t smem = ((tw * th + 15) & ~static_cast<size_t>(15)) + MORPH_TW * th;
	CVB_REQUIRE(smem <= 200 * 1024, CVB200_E_OUT_OF_BOUND);
	auto kern = erode ? morph_basic_kernel<true> : morph_basic_kernel<false>;
	if (smem > 48 * 1024) CVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
	dim3 grid(static_cast<unsigned>(div_up(p.W, MORPH_TW)), static_cast<unsigned>(div_up(p.H, MORPH_TH)), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	{ KernelScope ks_(erode ? "morph_erode" : "morph_dilate", stream);
	  kern<<<grid, MORPH_THREADS, smem, stream>>>(in, out, dTaps, p); }
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

static std::mutex g_morph_mutex, g_morph_host_mutex;
static DevBuf g_morph_taps, g_morph_tmp, g_morph_in, g_morph_out;
static cudaEvent_t g_morph_done = nullptr; // the tap list and the open/close intermediate are shared scratch: a call waits for the previous one, whatever its stream

} // namespace cvb

using namespace cvb;

extern "C" {

// buildStructuringElementGeneric (compv_math_morph.cxx:513-583); non-zero cells are 255 (:523)
int cvb200_morph_build_strel(uint8_t* strel, size_t width, size_t height, size_t strelStride, int type)
{
	CVB_REQUIRE(strel && width && height && strelStride >= width, CVB200_E_INVALID_PARAMETER);
	for (size_t j = 0; j < height; ++j) memset(strel + j * strelStride, type == CVB200_MATH_MORPH_STREL_TYPE_RECT ? 255 : 0, width);
	switch (type) {
	case CVB200_MATH_MORPH_STREL_TYPE_RECT: break;
	case CVB200_MATH_MORPH_STREL_TYPE_CROSS:
		for (size_t i = 0; i < width; ++i) strel[(height >> 1) * strelStride + i] = 255;
		for (size_t j = 0; j < height; ++j) strel[j * strelStride + (width >> 1)] = 255;
		break;
	case CVB200_MATH_MORPH_STREL_TYPE_DIAMOND: { // rows grow by two cells down to the middle row, then shrink; cells that would fall outside a non-square element are skipped
		const size_t hd = height >> 1;
		ptrdiff_t col = static_cast<ptrdiff_t>(width >> 1);
		size_t row = 0, count = 1;
		for (size_t j = 0; j < hd; ++j, count += 2, ++row, --col)
			for (size_t i = 0; i < count; ++i) { const ptrdiff_t c = col + static_cast<ptrdiff_t>(i); if (c >= 0 && c < static_cast<ptrdiff_t>(width) && row < height) strel[row * strelStride + c] = 255; }
		for (size_t j = 0; j <= hd; ++j, count -= 2, ++row, ++col)
			for (size_t i = 0; i < count; ++i) { const ptrdiff_t c = col + static_cast<ptrdiff_t>(i); if (c >= 0 && c < static_cast<ptrdiff_t>(width) && row < height) strel[row * strelStride + c] = 255; }
		break;
	}
	default: return CVB200_E_NOT_IMPLEMENTED;
	}
	return CVB200_S_OK;
}

int cvb200_morph_process_dev(const uint8_t* in, size_t width, size_t height, size_t stride, const uint8_t* strel, size_t strelWidth, size_t strelHeight, size_t strelStride,
	uint8_t* out, int opType, int borderType, size_t batch, size_t framePitch, cvb200_stream_t stream_)
{
	CVB_REQUIRE_INIT();
	// compv_math_morph.cxx:131-137
	CVB_REQUIRE(in && out && strel && width && height && stride >= width && strelWidth && strelHeight && strelStride >= strelWidth && width >= strelWidth && height >= strelHeight,
		CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(in != out, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(opType >= CVB200_MATH_MORPH_OP_TYPE_ERODE && opType <= CVB200_MATH_MORPH_OP_TYPE_CLOSE, CVB200_E_NOT_IMPLEMENTED);
	CVB_REQUIRE(borderType == CVB200_BORDER_TYPE_ZERO || borderType == CVB200_BORDER_TYPE_REPLICATE || borderType == CVB200_BORDER_TYPE_IGNORE, CVB200_E_NOT_IMPLEMENTED);
	CVB_REQUIRE(strelWidth <= 255 && strelHeight <= 255, CVB200_E_OUT_OF_BOUND);
	if (!batch) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	cudaStream_t stream = as_stream(stream_);
	std::vector<short2> taps;
	for (size_t j = 0; j < strelHeight; ++j) for (size_t i = 0; i < strelWidth; ++i)
		if (strel[j * strelStride + i]) taps.push_back(make_short2(static_cast<short>(i), static_cast<short>(j)));
	CVB_REQUIRE(!taps.empty(), CVB200_E_INVALID_PARAMETER); // :483 "Structured element is full of zeros"
	MorphParams p;
	memset(&p, 0, sizeof(p));
	p.W = static_cast<int>(width); p.H = static_cast<int>(height); p.stride = stride; p.framePitch = framePitch;
	p.sw = static_cast<int>(strelWidth); p.sh = static_cast<int>(strelHeight); p.rw = p.sw >> 1; p.rh = p.sh >> 1; p.bh = (p.sh + 1) >> 1;
	p.border = borderType;
	p.nTaps = (taps.size() == strelWidth * strelHeight) ? 0 : static_cast<int>(taps.size());
	std::lock_guard<std::mutex> lock(g_morph_mutex);
	if (!g_morph_done) CVB_CUDA(cudaEventCreateWithFlags(&g_morph_done, cudaEventDisableTiming));
	else CVB_CUDA(cudaStreamWaitEvent(stream, g_morph_done, 0));
	struct Done { cudaStream_t s; ~Done() { cudaEventRecord(g_morph_done, s); } } done_{ stream };
	CVB_CHECK(g_morph_taps.ensure(taps.size() * sizeof(short2)));
	CVB_CUDA(cudaMemcpyAsync(g_morph_taps.p, taps.data(), taps.size() * sizeof(short2), cudaMemcpyHostToDevice, stream));
	const short2* dTaps = g_morph_taps.as<short2>();
	if (opType == CVB200_MATH_MORPH_OP_TYPE_ERODE || opType == CVB200_MATH_MORPH_OP_TYPE_DILATE)
		return morph_launch(in, out, dTaps, p, opType == CVB200_MATH_MORPH_OP_TYPE_ERODE, batch, stream);
	// open = erode then dilate, close = dilate then erode (:104-111); the second operation works on the first one's complete output, borders included
	const size_t bytes = (batch - 1) * framePitch + stride * height;
	CVB_CHECK(g_morph_tmp.ensure(bytes));
	if (borderType == CVB200_BORDER_TYPE_IGNORE) CVB_CUDA(cudaMemcpyAsync(g_morph_tmp.p, out, bytes, cudaMemcpyDeviceToDevice, stream));
	const bool first = (opType == CVB200_MATH_MORPH_OP_TYPE_OPEN);
	CVB_CHECK(morph_launch(in, g_morph_tmp.as<uint8_t>(), dTaps, p, first, batch, stream));
	return morph_launch(g_morph_tmp.as<uint8_t>(), out, dTaps, p, !first, batch, stream);
}

int cvb200_morph_process(const uint8_t* in, size_t width, size_t height, size_t stride, const uint8_t* strel, size_t strelWidth, size_t strelHeight, size_t strelStride,
	uint8_t* out, int opType, int borderType)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(in && out && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	const size_t n = stride * height;
	std::lock_guard<std::mutex> hostLock(g_morph_host_mutex); // the staging buffers are shared: one host-buffer call at a time
	CVB_CHECK(g_morph_in.ensure(n));
	CVB_CHECK(g_morph_out.ensure(n));
	CVB_CUDA(cudaMemcpyAsync(g_morph_in.p, in, n, cudaMemcpyHostToDevice, 0));
	CVB_CUDA(cudaMemcpyAsync(g_morph_out.p, out, n, cudaMemcpyHostToDevice, 0)); // IGNORE keeps the caller's border cells
	CVB_CHECK(cvb200_morph_process_dev(g_morph_in.as<uint8_t>(), width, height, stride, strel, strelWidth, strelHeight, strelStride, g_morph_out.as<uint8_t>(), opType, borderType, 1, n, nullptr));
	CVB_CUDA(cudaMemcpyAsync(out, g_morph_out.p, n, cudaMemcpyDeviceToHost, 0));
	CVB_CUDA(cudaStreamSynchronize(0));
	return CVB200_S_OK;
}

} // extern "C"
