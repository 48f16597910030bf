// a6 -- standard Hough transform (SHT) line detector.
// Replaces CompVHoughSht::process (core/features/hough/compv_core_feature_houghsht.cxx:96-262) and its helpers: initCoords :320-348 (16.16 fixed-point
// sin/cos tables), acc_gather :350-481 (row kernel :594-605), nms_gather :483-531 (row kernels :607-626 and intrin/x86/..._houghsht_intrin_sse2.cxx:16-50),
// nms_apply :533-564 (row kernel :651-668), std::sort + maxLines :241-247.
//
// The reference walks the edge list once and scatters theta-count increments per edge into a (2(W+H)+1) x thetaCount int32 accumulator in memory.
// Here the loop nest is turned inside out so that no vote ever leaves the SM:
//   sht_list   edge bytes -> per-frame list of packed (x | y<<16) coordinates (order free: integer votes commute)             HBM: 1 B/px read
//   sht_vote   ONE CTA PER (theta, frame): the rho histogram of that theta (2(W+H)+1 ints, 24 KB at 1080p) lives in shared memory, the CTA streams the
//              L2-resident edge list through it with shared-memory atomics and writes the finished accumulator column out once, coalesced (theta-major layout)
//   sht_nms    3x3 strict-greater suppression exactly as nms_gather + nms_apply, one bit per cell in the reference's row-major order
//   sht_emit   one CTA per frame: block scan over the rows, lines written in accumulator row-major order (the order nms_apply pushes them in)
// Frames are processed in chunks whose accumulators fit the L2 (64 MiB), with no host synchronisation until the whole batch is queued.
// Host: std::sort (same libstdc++, same comparator, same input order => the reference's tie order) + maxLines.
#include "hough.cuh"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

namespace cvb {

struct ShtGeom {
	int W, H;
	size_t stride, framePitch;
	int R, RP, T, TW;          // accumulator rows (rho), rows padded to 4, theta count, theta count in 32-bit mask words
	int barrier, thr;
	int cBegin, cEnd;          // theta columns nms_gather applies the suppression test to
	float fTheta;
	unsigned int listCap;      // edge list capacity per frame (W*H)
	unsigned int poolCap;
	int yOff;                  // row-strip mode: the strip's first row in the full frame (added to y before voting)
};
struct ShtDesc { unsigned int base, total; };

// ---- edge list ----
template <bool kAligned4>
__global__ void __launch_bounds__(256) sht_list_kernel(const uint8_t* __restrict__ edges, unsigned int* __restrict__ list, unsigned int* __restrict__ cursor, ShtGeom g)
{
	__shared__ unsigned int sWarp[8];
	__shared__ unsigned int sBase;
	const int f = blockIdx.z, y = blockIdx.y;
	const int x = (blockIdx.x * 256 + threadIdx.x) * 4;
	const uint8_t* row = edges + static_cast<size_t>(f) * g.framePitch + static_cast<size_t>(y) * g.stride;
	unsigned int px = 0; // the 4 pixels, one per byte
	if (x < g.W) {
		if (kAligned4 && x + 4 <= g.W) px = *reinterpret_cast<const unsigned int*>(row + x);
		else {
			for (int k = 0; k < 4 && x + k < g.W; ++k) px |= static_cast<unsigned int>(row[x + k]) << (8 * k);
		}
	}
	// non-zero bytes -> 4-bit mask
	unsigned int nz = 0;
	#pragma unroll
	for (int k = 0; k < 4; ++k) nz |= ((px >> (8 * k)) & 0xffu) ? (1u << k) : 0u;
	const unsigned int n = __popc(nz);
	// block-wide exclusive scan of n
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned int inc = n;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
	if (lane == 31) sWarp[warp] = inc;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned int tot = 0;
		#pragma unroll
		for (int w = 0; w < 8; ++w) { const unsigned int v = sWarp[w]; sWarp[w] = tot; tot += v; }
		sBase = tot ? atomicAdd(&cursor[f], tot) : 0u;
	}
	__syncthreads();
	if (n) {
		unsigned int o = sBase + sWarp[warp] + (inc - n);
		unsigned int* L = list + static_cast<size_t>(f) * g.listCap;
		const unsigned int yy = static_cast<unsigned int>(y + g.yOff) << 16;
		#pragma unroll
		for (int k = 0; k < 4; ++k) if (nz & (1u << k)) L[o++] = yy | static_cast<unsigned int>(x + k);
	}
}

// ---- voting: one CTA per (theta, frame), histogram over rho in shared memory ----
__global__ void __launch_bounds__(256) sht_vote_kernel(const unsigned int* __restrict__ list, const unsigned int* __restrict__ cursor,
	const int* __restrict__ cosT, const int* __restrict__ sinT, int* __restrict__ acc, ShtGeom g)
{
	extern __shared__ __align__(16) int hist[];
	const int t = blockIdx.x, f = blockIdx.y;
	for (int i = threadIdx.x; i < g.RP; i += 256) hist[i] = 0;
	__syncthreads();
	const int c = cosT[t], s = sinT[t];
	const unsigned int n = cursor[f];
	const unsigned int* L = list + static_cast<size_t>(f) * g.listCap;
	unsigned int i = threadIdx.x;
	// rho = (x*cos + y*sin) >> 16 in wrapping int32 arithmetic (houghsht.cxx:601); accumulator row = barrier - rho
	for (; i + 768 < n; i += 1024) {
		const unsigned int p0 = L[i], p1 = L[i + 256], p2 = L[i + 512], p3 = L[i + 768];
		const int r0 = (static_cast<int>(p0 & 0xffffu) * c + static_cast<int>(p0 >> 16) * s) >> 16;
		const int r1 = (static_cast<int>(p1 & 0xffffu) * c + static_cast<int>(p1 >> 16) * s) >> 16;
		const int r2 = (static_cast<int>(p2 & 0xffffu) * c + static_cast<int>(p2 >> 16) * s) >> 16;
		const int r3 = (static_cast<int>(p3 & 0xffffu) * c + static_cast<int>(p3 >> 16) * s) >> 16;
		atomicAdd(&hist[g.barrier - r0], 1);
		atomicAdd(&hist[g.barrier - r1], 1);
		atomicAdd(&hist[g.barrier - r2], 1);
		atomicAdd(&hist[g.barrier - r3], 1);
	}
	for (; i < n; i += 256) {
		const unsigned int p = L[i];
		const int r = (static_cast<int>(p & 0xffffu) * c + static_cast<int>(p >> 16) * s) >> 16;
		atomicAdd(&hist[g.barrier - r], 1);
	}
	__syncthreads();
	int4* out = reinterpret_cast<int4*>(acc + (static_cast<size_t>(f) * g.T + t) * g.RP);
	const int4* h4 = reinterpret_cast<const int4*>(hist);
	for (int k = threadIdx.x; k < (g.RP >> 2); k += 256) out[k] = h4[k];
}

// ---- non-maxima suppression -> one bit per accumulator cell, stored in the reference's row-major (rho row, theta column) order ----
__global__ void __launch_bounds__(128) sht_nms_kernel(const int* __restrict__ acc, unsigned int* __restrict__ mask, ShtGeom g)
{
	const int r = blockIdx.x * 128 + threadIdx.x;
	if (r >= g.R) return;
	const int w = blockIdx.y, f = blockIdx.z;
	const int* A = acc + static_cast<size_t>(f) * g.T * g.RP;
	const bool rowInside = (r >= 1) && (r <= g.R - 2); // nms_gather :487-488
	unsigned int m = 0;
	const int tEnd = min(g.T, w * 32 + 32);
	for (int t = w * 32; t < tEnd; ++t) {
		const int v = A[static_cast<size_t>(t) * g.RP + r];
		if (v > g.thr) {
			bool sup = false;
			if (rowInside && t >= g.cBegin && t < g.cEnd) {
				const int* cc = A + static_cast<size_t>(t) * g.RP + r;
				sup = (cc[-1] > v) || (cc[1] > v);
				if (t > 0) { const int* cl = cc - g.RP; sup = sup || (cl[-1] > v) || (cl[0] > v) || (cl[1] > v); }
				if (t + 1 < g.T) { const int* cr = cc + g.RP; sup = sup || (cr[-1] > v) || (cr[0] > v) || (cr[1] > v); }
			}
			if (!sup) m |= 1u << (t & 31);
		}
	}
	mask[(static_cast<size_t>(f) * g.R + r) * g.TW + w] = m;
}

// ---- ordered emission: one CTA per frame ----
__global__ void __launch_bounds__(1024) sht_emit_kernel(const int* __restrict__ acc, const unsigned int* __restrict__ mask, cvb200_hough_line_t* __restrict__ pool,
	unsigned int* __restrict__ poolCursor, ShtDesc* __restrict__ desc, int frame0, ShtGeom g)
{
	__shared__ unsigned int sWarp[32];
	__shared__ unsigned int sBase, sTotal;
	const int f = blockIdx.x;
	const int* A = acc + static_cast<size_t>(f) * g.T * g.RP;
	const unsigned int* M = mask + static_cast<size_t>(f) * g.R * g.TW;
	const int rowsPer = (g.R + 1023) / 1024;
	const int r0 = min(g.R, static_cast<int>(threadIdx.x) * rowsPer), r1 = min(g.R, r0 + rowsPer);
	unsigned int n = 0;
	for (int k = r0 * g.TW; k < r1 * g.TW; ++k) n += __popc(M[k]);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned int inc = n;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
	if (lane == 31) sWarp[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		unsigned int v = sWarp[lane], s = v;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += u; }
		sWarp[lane] = s - v;
		if (lane == 31) {
			sTotal = s;
			sBase = atomicAdd(poolCursor, s);
			desc[frame0 + f].base = sBase; desc[frame0 + f].total = s;
		}
	}
	__syncthreads();
	const unsigned int base = sBase;
	if (static_cast<unsigned long long>(base) + sTotal > g.poolCap) return; // the host sees cursor > poolCap, grows the pool and runs the batch again
	unsigned int o = base + sWarp[warp] + (inc - n);
	for (int r = r0; r < r1; ++r) {
		for (int w = 0; w < g.TW; ++w) {
			unsigned int m = M[r * g.TW + w];
			while (m) {
				const int b = __ffs(m) - 1; m &= m - 1;
				const int t = w * 32 + b;
				cvb200_hough_line_t L;
				L.rho = static_cast<float>(g.barrier - r);                  // nms_apply :659-663
				L.theta = __fmul_rn(static_cast<float>(t), g.fTheta);
				L.strength = static_cast<size_t>(A[static_cast<size_t>(t) * g.RP + r]);
				pool[o++] = L;
			}
		}
	}
}

int sht_process_dev(cvb200_hough* h, const uint8_t* edges, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch,
	cvb200_hough_line_t* lines, size_t capacity, size_t* counts, cudaStream_t stream)
{
	CVB_REQUIRE(h->rho == 1.f && h->theta > 0.f, CVB200_E_INVALID_PARAMETER);
	// Row-strip mode (cvb200_hough_sht_accumulate_dev / _lines_dev): `height` rows starting at row shtYOffset of a frame of shtFullHeight rows; the accumulator geometry is
	// the full frame's, stage 1 stops after the voting (the caller sums the strips' accumulators, e.g. with an all-reduce), stage 2 starts from a given accumulator.
	const size_t stripHeight = height;
	const int stage = h->shtStage;
	if (h->shtFullHeight) { CVB_REQUIRE(h->shtYOffset + height <= h->shtFullHeight && batch == 1, CVB200_E_INVALID_PARAMETER); height = h->shtFullHeight; }
	CVB_REQUIRE(stage == 0 || stage == 3 || (h->shtExternalAcc && batch == 1), CVB200_E_INVALID_PARAMETER);
	// x*cos16 + y*sin16 must not leave int32 and the rho histogram must fit one SM's shared memory (227 KB)
	CVB_REQUIRE(width <= 65535 && height <= 65535 && (width + height) <= 29000, CVB200_E_OUT_OF_BOUND);
	CVB_REQUIRE(batch < 65536 && width * height < (1ull << 31), CVB200_E_OUT_OF_BOUND);
	static const float kPi = 3.1415926535897932384626433f;
	static const float kPiOver180 = kPi / 180.f;
	const float fRho = h->rho * 1.f;
	const float fTheta = h->theta * kPiOver180; // ctor :44
	const size_t R = static_cast<size_t>((static_cast<float>(((width + height) << 1) + 1) / fRho) + 0.5); // initCoords :326 (ROUNDFU: + 0.5 in double)
	const size_t T = static_cast<size_t>((kPi / fTheta) + 0.5);                                            // :327
	CVB_REQUIRE(R >= 3 && T >= 1 && T < (1u << 20), CVB200_E_INVALID_PARAMETER);
	ShtGeom g;
	memset(&g, 0, sizeof(g));
	g.W = static_cast<int>(width); g.H = static_cast<int>(stripHeight); g.stride = stride; g.framePitch = framePitch;
	g.yOff = static_cast<int>(h->shtYOffset);
	g.R = static_cast<int>(R); g.RP = static_cast<int>((R + 3) & ~static_cast<size_t>(3)); g.T = static_cast<int>(T); g.TW = static_cast<int>(div_up(T, 32));
	g.barrier = static_cast<int>(width + height);
	g.thr = static_cast<int>(h->threshold);
	g.fTheta = fTheta;
	{
		const size_t maxCols = T - 1;
		if (h->x86Simd && maxCols >= 4) { // the SSE2 row kernel: groups of 4 columns from column 1; the scalar tail is never entered (see oracle/compv_oracle_sht.cpp)
			const size_t m4 = maxCols & ~static_cast<size_t>(3);
			g.cBegin = 1; g.cEnd = static_cast<int>(std::min(T, 1 + 4 * ((m4 - 1 + 3) / 4)));
		}
		else { g.cBegin = 0; g.cEnd = static_cast<int>(maxCols); } // generic C++: columns [0, cols-1), column -1 reads the zeroed row padding
	}
	g.listCap = static_cast<unsigned int>(width * stripHeight);
	if (h->shtAccElems) *h->shtAccElems = T * static_cast<size_t>((R + 3) & ~static_cast<size_t>(3));
	if (stage == 3) return CVB200_S_OK; // size query only

	// fixed-point tables (initCoords :337-340): same libm, same float accumulation of the angle as the reference
	std::vector<int32_t> tab(2 * T);
	{
		float tt = 0.f;
		for (size_t t = 0; t < T; ++t, tt += fTheta) {
			tab[t] = static_cast<int32_t>((cosf(tt) * fRho) * 65535.f);
			tab[T + t] = static_cast<int32_t>((sinf(tt) * fRho) * 65535.f);
		}
	}
	CVB_CHECK(h->shtTables.ensure(2 * T * 4));
	CVB_CUDA(cudaMemcpyAsync(h->shtTables.p, tab.data(), 2 * T * 4, cudaMemcpyHostToDevice, stream));
	const int* dCos = h->shtTables.as<int>();
	const int* dSin = dCos + T;

	const size_t accFrame = T * static_cast<size_t>(g.RP) * 4;
	size_t chunk = (64u << 20) / accFrame;
	if (chunk < 1) chunk = 1;
	if (chunk > batch) chunk = batch;
	if (!h->shtExternalAcc) CVB_CHECK(h->acc.ensure(chunk * accFrame));
	int* dAcc = h->shtExternalAcc ? h->shtExternalAcc : h->acc.as<int>();
	CVB_CHECK(h->shtList.ensure(chunk * static_cast<size_t>(g.listCap) * 4));
	CVB_CHECK(h->shtCursor.ensure((chunk + 1) * 4));
	CVB_CHECK(h->shtMask.ensure(chunk * R * g.TW * 4));
	CVB_CHECK(h->shtDesc.ensure(batch * sizeof(ShtDesc)));
	CVB_CHECK(h->hFrames.ensure(batch * sizeof(ShtDesc) + 16));
	{
		size_t want = std::max<size_t>(size_t(1) << 16, batch * 4096) * sizeof(cvb200_hough_line_t);
		if (h->shtPool.bytes < want) CVB_CHECK(h->shtPool.ensure(want));
	}
	const size_t smem = static_cast<size_t>(g.RP) * 4;
	if (smem > 48 * 1024) CVB_CUDA(cudaFuncSetAttribute(sht_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
	const bool aligned4 = ((reinterpret_cast<uintptr_t>(edges) | stride | framePitch) & 3) == 0;
	unsigned int* dCursor = h->shtCursor.as<unsigned int>();
	unsigned int* dPoolCursor = dCursor + chunk;
	ShtDesc* hDesc = h->hFrames.as<ShtDesc>();
	unsigned int* hPoolCursor = reinterpret_cast<unsigned int*>(hDesc + batch);
	CVB_REQUIRE(static_cast<size_t>(g.TW) <= 65535, CVB200_E_OUT_OF_BOUND);

	for (int attempt = 0; ; ++attempt) {
		g.poolCap = static_cast<unsigned int>(std::min<size_t>(h->shtPool.bytes / sizeof(cvb200_hough_line_t), 0xffffffffu));
		CVB_CUDA(cudaMemsetAsync(dPoolCursor, 0, 4, stream));
		for (size_t f0 = 0; f0 < batch; f0 += chunk) {
			const unsigned int F = static_cast<unsigned int>(std::min(chunk, batch - f0));
			const uint8_t* e = edges + f0 * framePitch;
			if (stage != 2) {
				CVB_CUDA(cudaMemsetAsync(dCursor, 0, F * 4, stream));
				{
					dim3 grid(static_cast<unsigned>(div_up(width, 1024)), static_cast<unsigned>(stripHeight), F);
					KernelScope ks_("sht_list", stream);
					if (aligned4) sht_list_kernel<true><<<grid, 256, 0, stream>>>(e, h->shtList.as<unsigned int>(), dCursor, g);
					else sht_list_kernel<false><<<grid, 256, 0, stream>>>(e, h->shtList.as<unsigned int>(), dCursor, g);
				}
				CVB_LAUNCHED();
				{ KernelScope ks_("sht_vote", stream);
				  sht_vote_kernel<<<dim3(static_cast<unsigned>(T), F), 256, smem, stream>>>(h->shtList.as<unsigned int>(), dCursor, dCos, dSin, dAcc, g); }
				CVB_LAUNCHED();
			}
			if (stage == 1) { CVB_CUDA(cudaStreamSynchronize(stream)); return CVB200_S_OK; } // the strip's votes are in the caller's accumulator
			{ KernelScope ks_("sht_nms", stream);
			  sht_nms_kernel<<<dim3(static_cast<unsigned>(div_up(R, 128)), static_cast<unsigned>(g.TW), F), 128, 0, stream>>>(dAcc, h->shtMask.as<unsigned int>(), g); }
			CVB_LAUNCHED();
			{ KernelScope ks_("sht_emit", stream);
			  sht_emit_kernel<<<F, 1024, 0, stream>>>(dAcc, h->shtMask.as<unsigned int>(), h->shtPool.as<cvb200_hough_line_t>(), dPoolCursor, h->shtDesc.as<ShtDesc>(),
				static_cast<int>(f0), g); }
			CVB_LAUNCHED();
		}
		CVB_CUDA(cudaMemcpyAsync(hDesc, h->shtDesc.p, batch * sizeof(ShtDesc), cudaMemcpyDeviceToHost, stream));
		CVB_CUDA(cudaMemcpyAsync(hPoolCursor, dPoolCursor, 4, cudaMemcpyDeviceToHost, stream));
		CVB_CUDA(cudaStreamSynchronize(stream));
		size_t need = 0;
		for (size_t f = 0; f < batch; ++f) need += hDesc[f].total;
		CVB_REQUIRE(need < 0xffffffffull, CVB200_E_OUT_OF_BOUND);
		if (need <= g.poolCap) break;
		CVB_REQUIRE(attempt == 0, CVB200_E_INVALID_STATE);
		CVB_CHECK(h->shtPool.ensure(need * sizeof(cvb200_hough_line_t)));
	}
	const size_t used = *hPoolCursor;
	CVB_CHECK(h->hVotes.ensure(std::max<size_t>(used, 1) * sizeof(cvb200_hough_line_t)));
	cvb200_hough_line_t* hp = h->hVotes.as<cvb200_hough_line_t>();
	if (used) {
		CVB_CUDA(cudaMemcpyAsync(hp, h->shtPool.p, used * sizeof(cvb200_hough_line_t), cudaMemcpyDeviceToHost, stream));
		CVB_CUDA(cudaStreamSynchronize(stream));
	}

	// ---- host: std::sort by strength + maxLines (houghsht.cxx:241-247) ----
	const size_t lim = (h->maxLines <= 0) ? static_cast<size_t>(INT_MAX) : static_cast<size_t>(h->maxLines);
	host_parallel_for(batch, [&](size_t f) {
		cvb200_hough_line_t* v = hp + hDesc[f].base;
		size_t n = hDesc[f].total;
		std::sort(v, v + n, [](const cvb200_hough_line_t& a, const cvb200_hough_line_t& b) -> bool { return a.strength > b.strength; });
		if (n > lim) n = lim;
		if (capacity) memcpy(lines + f * capacity, v, std::min(n, capacity) * sizeof(cvb200_hough_line_t));
		counts[f] = n;
	});
	return CVB200_S_OK;
}

} // namespace cvb

// ---- row-strip mode of the SHT: every GPU votes its strip's edge pixels into its own accumulator of the FULL frame's geometry, the accumulators are summed
// (all-reduce of int32, done by the caller over NCCL), one call turns the sum into lines.  The reference's own thread split is the same sum (houghsht.cxx:455-477). ----
using namespace cvb;
extern "C" {

static int sht_strip_call(cvb200_hough_t* h, const uint8_t* edges, size_t width, size_t stripHeight, size_t stride, size_t fullHeight, size_t yOffset, int* acc, int stage, size_t* accElems,
	cvb200_hough_line_t* lines, size_t capacity, size_t* count, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(h && h->id == CVB200_HOUGHSHT_ID && width && stripHeight && fullHeight, CVB200_E_INVALID_PARAMETER);
	std::lock_guard<std::mutex> lock(h->mutex);
	h->shtFullHeight = fullHeight; h->shtYOffset = yOffset; h->shtExternalAcc = acc; h->shtStage = stage; h->shtAccElems = accElems;
	size_t dummy = 0;
	const int rc = sht_process_dev(h, edges, width, stripHeight, stride, 1, stride * stripHeight, lines, capacity, count ? count : &dummy, as_stream(stream));
	h->shtFullHeight = 0; h->shtYOffset = 0; h->shtExternalAcc = nullptr; h->shtStage = 0; h->shtAccElems = nullptr;
	return rc;
}

int cvb200_hough_sht_acc_size(cvb200_hough_t* h, size_t width, size_t fullHeight, size_t* elems)
{
	CVB_REQUIRE(elems, CVB200_E_INVALID_PARAMETER);
	return sht_strip_call(h, nullptr, width, fullHeight, width, fullHeight, 0, nullptr, 3, elems, nullptr, 0, nullptr, nullptr);
}

int cvb200_hough_sht_accumulate_dev(cvb200_hough_t* h, const uint8_t* edges, size_t width, size_t stripHeight, size_t stride, size_t fullHeight, size_t yOffset, int32_t* acc, cvb200_stream_t stream)
{
	CVB_REQUIRE(edges && acc, CVB200_E_INVALID_PARAMETER);
	return sht_strip_call(h, edges, width, stripHeight, stride, fullHeight, yOffset, acc, 1, nullptr, nullptr, 0, nullptr, stream);
}

int cvb200_hough_sht_lines_dev(cvb200_hough_t* h, int32_t* acc, size_t width, size_t fullHeight, cvb200_hough_line_t* lines, size_t capacity, size_t* count, cvb200_stream_t stream)
{
	CVB_REQUIRE(acc && count && (lines || !capacity), CVB200_E_INVALID_PARAMETER);
	return sht_strip_call(h, nullptr, width, fullHeight, width, fullHeight, 0, acc, 2, nullptr, lines, capacity, count, stream);
}

} // extern "C"
