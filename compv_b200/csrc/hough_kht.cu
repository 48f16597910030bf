// a7 -- kernel-based Hough transform (KHT) line detector.
// Replaces CompVHoughKht::process (core/features/hough/compv_core_feature_houghkht.cxx:208-447) and its helpers:
//   initCoords :501-541, linking_AppendixA / link_Algorithm5 / Algorithm 6 :544-760, clusters_subdivision :774-832,
//   voting_Algorithm2_Kernels :885-1026 (+ CompVMathEigen::find2x2, base/math/compv_math_eigen.cxx:285-342), DiscardShortKernels :1029-1041,
//   Gmin :1044-1062 (+ __gauss_Eq15 :834-847), voting_Algorithm2_Count / vote_Algorithm4 :1065-1148,
//   peaks_Section3_4 (3x3 smoothing + threshold :1282-1308, std::sort :1195-1204, sweep :1207-1247).
// This translation unit is compiled with -fmad=false: every double operation is an individually rounded IEEE op in the reference's order
// (the reference's KHT_TYP is double and its x86 build has no FMA in this file), so kernels, Gs and the integer votes reproduce bit for bit.
//
// Device pipeline (one launch each for the whole batch; frames are independent):
//   kht_bits      edge bytes -> 1 bit/px bitmap (+ per-frame edge count)                                  HBM: 1 B/px read, 1/8 B/px written
//   kht_link      the linking procedure.  It is a raster scan that erases pixels as it walks, i.e. inherently ordered: ONE WARP PER FRAME
//                 runs it (32 lanes scan 32 bitmap words per step with a ballot, lane 0 walks a string through the L1-resident bitmap);
//                 frames of the batch run concurrently on different SMs.  Same strings, same order as the reference.
//   kht_subdivide one thread per string: the recursive segmentation as an explicit-stack post-order walk
//   kht_scan / kht_kernels / kht_hmax / kht_gmin / kht_vote (one thread per kernel quadrant, integer atomicAdd) / kht_peaks (+ rank prefix)
// Host: the thresholded, smoothed cells (a few thousand per frame) are sorted with the same libstdc++ std::sort as the reference and swept.
#include "hough.cuh"
#include "kht_walk.cuh"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace cvb {

struct KhtGeom {
	int W, H, WW;              // WW = bitmap words per row = ceil(W/32) + 2: one zero word before and after the pixels; a frame's bitmap has H + 2 rows (a zero row above and below)
	size_t stride, framePitch;
	unsigned int nRho, nTheta, cs; // cs = accumulator pitch (nRho + 2)
	double dRho, dThetaDeg, rhoMaxNeg, halfW, halfH;
	double minDeviation, minHeight;
	unsigned int minSize;
	int threshold;
	int x86Simd;
};

struct KhtKernel { double rho, theta, h, sts, srs, m2, srt; unsigned int alive; unsigned int pad; };
struct KhtFrame {           // per-frame offsets into the batch-wide pools + device-side counters
	unsigned int posOff, posCap;      // positions / clusters / stack share this index space
	unsigned int strOff, strCap;
	unsigned int nPos, nStr, nClus, nVotes;
	unsigned long long hmaxBits, gminBits;
	double gs;
};
struct KhtStack { unsigned int a, b, mi, nclus0; double ratio, ratioLeft; unsigned int state, pad; };

// ---- bitmap -------------------------------------------------------------------------------------
template <bool REV>
__global__ void kht_bits_kernel(const uint8_t* __restrict__ edges, unsigned int* __restrict__ bits, KhtGeom g, unsigned int* edgeCount)
{
	const int frame = blockIdx.z, y = blockIdx.y;
	const int wi = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned int word = 0;
	if (wi < g.WW - 2) {
		const uint8_t* row = edges + frame * g.framePitch + static_cast<size_t>(y) * g.stride;
		const int x0 = wi * 32;
		if (x0 + 32 <= g.W && ((reinterpret_cast<uintptr_t>(row + x0) & 15) == 0)) {
			const uint4 a = *reinterpret_cast<const uint4*>(row + x0), b = *reinterpret_cast<const uint4*>(row + x0 + 16);
			const unsigned int v[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
#pragma unroll
			for (int k = 0; k < 8; ++k) {
#pragma unroll
				for (int j = 0; j < 4; ++j) if ((v[k] >> (8 * j)) & 0xffu) word |= 1u << (4 * k + j);
			}
		}
		else {
			for (int j = 0; j < 32 && x0 + j < g.W; ++j) if (row[x0 + j]) word |= 1u << j;
		}
		bits[(static_cast<size_t>(frame) * (g.H + 2 * KHT_PADR) + y + KHT_PADR) * g.WW + wi + 1] = REV ? __brev(word) : word;
	}
	unsigned int c = __popc(word);
	for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(&edgeCount[frame], c);
}

// ---- linking ------------------------------------------------------------------------------------
// The walk is a single dependent instruction chain on one lane: its cost is (instructions per step) x (issue latency, ~4 cycles).  Measured on frame G
// (82k edge px, 5.9k walks per 1080p frame): three bitmap row loads per step 20.8 ms (~110 instr/step); shared-memory band caches 30-47 ms; a 3-row x 64-column
// register window with bounds checks 16.0 ms (91 instr/step, ncu r1c); without bounds checks, running pointers 13.3 ms (~80 instr/step).  This version is written
// for instruction count: the bitmap is zero padded on all four sides so the walker has no border case at all, the window lives in six 32-bit registers,
// the position is one packed register (x | y << 16, the value that gets stored), the 9-bit neighbourhood is assembled in the order of Algorithm 6
// (houghkht.cxx:666-703: TL, T, TR, L, [centre, already erased], R, BL, B, BR) so that one find-first-set yields k with dx = k % 3 - 1, dy = k / 3 - 1.
// Every access to a frame's bitmap comes from this one warp, so its SM's L1 stays coherent with the stores.
#define KHT_AHEAD 96
struct KhtWalk {
	unsigned int t0, t1, c0, c1, b0, b1; // bitmap rows y-1, y, y+1: padded columns [32*wb, 32*wb + 64)
	unsigned int* p;                     // &paddedRow(y)[wb]
	int rel;                             // padded column - 32*wb, kept in [1, 62]
	unsigned int xy;                     // x | y << 16 (image coordinates)
};

__device__ __forceinline__ void kht_walk_load(KhtWalk& w, unsigned int* bits /* padded row 0 of the frame */, int WW)
{
	const int X = static_cast<int>(w.xy & 0xffffu) + 32, y = static_cast<int>(w.xy >> 16);
	const int wi = X >> 5; // >= 1
	const int wb = ((X & 31) < 16) ? wi - 1 : wi; // the pixel sits in columns [16, 47] of the window; wb + 1 <= WW - 1
	w.rel = X - (wb << 5);
	w.p = bits + y * WW + wb;
	w.t0 = w.p[-WW]; w.t1 = w.p[1 - WW];
	w.c0 = w.p[0]; w.c1 = w.p[1];
	w.b0 = w.p[WW]; w.b1 = w.p[WW + 1];
}

__device__ __forceinline__ unsigned int kht_shr64(unsigned int lo, unsigned int hi, int sh) // low word of (hi:lo) >> sh, sh in [0, 63]
{
	return static_cast<unsigned int>(((static_cast<unsigned long long>(hi) << 32) | lo) >> sh);
}

// erase the current pixel in the registers and in memory
__device__ __forceinline__ void kht_walk_erase(KhtWalk& w)
{
	const unsigned int mask = 1u << (w.rel & 31);
	const int hi = w.rel >> 5;
	if (hi) w.c1 &= ~mask; else w.c0 &= ~mask;
	w.p[hi] = hi ? w.c1 : w.c0;
}

// move to the next pixel of the string; false when the current (already erased) pixel has no neighbour left
__device__ __forceinline__ bool kht_walk_next(KhtWalk& w, unsigned int* bits, int WW)
{
	const int sh = w.rel - 1;
	unsigned int m = kht_shr64(w.t0, w.t1, sh) & 7u;
	m |= (kht_shr64(w.c0, w.c1, sh) & 7u) << 3;
	m |= (kht_shr64(w.b0, w.b1, sh) & 7u) << 6;
	if (!m) return false;
	const int k = __ffs(m) - 1;           // 0..8 (never 4: the centre is erased)
	const int dy1 = (k * 11) >> 5;        // k / 3
	const int dx1 = k - 3 * dy1;          // k % 3
	w.rel += dx1 - 1;
	w.xy += static_cast<unsigned int>(dx1 + (dy1 << 16) - 65537);
	if (static_cast<unsigned int>(w.rel - 1) > 61u) { kht_walk_load(w, bits, WW); return true; } // left the window sideways: re-centre
	if (dy1 == 0) { w.p -= WW; w.b0 = w.c0; w.b1 = w.c1; w.c0 = w.t0; w.c1 = w.t1; w.t0 = w.p[-WW]; w.t1 = w.p[1 - WW]; }
	else if (dy1 == 2) { w.p += WW; w.t0 = w.c0; w.t1 = w.c1; w.c0 = w.b0; w.c1 = w.b1; w.b0 = w.p[WW]; w.b1 = w.p[WW + 1]; }
	return true;
}

__global__ void __launch_bounds__(32)
kht_link_old_kernel(unsigned int* bitsAll /* read and written through several derived pointers: no __restrict__ */, ushort2* __restrict__ possAll, uint2* __restrict__ stringsAll, KhtFrame* frames, KhtGeom g)
{
	const int frame = blockIdx.x, lane = threadIdx.x;
	const int W = g.W, H = g.H, WW = g.WW;
	unsigned int* bits = bitsAll + (static_cast<size_t>(frame) * (H + 2 * KHT_PADR) + KHT_PADR) * WW; // padded word 0 of image row 0
	KhtFrame& fr = frames[frame];
	unsigned int* poss = reinterpret_cast<unsigned int*>(possAll + fr.posOff); // ushort2 {x, y} written as x | y << 16
	uint2* strings = stringsAll + fr.strOff;
	unsigned int nPos = 0, nStr = 0; // meaningful on lane 0
	const int lastWord = (W - 1) >> 5;

	// The seed scan reads rows in order, so rows above the scan line are in this SM's L1; walks mostly head DOWN into rows nobody has read yet and
	// would pay an L2 round trip per new row (the chain's dominant latency).  The idle lanes therefore keep KHT_AHEAD rows below the scan line prefetched.
	const char* bytes0 = reinterpret_cast<const char*>(bits - WW);                       // padded row -1
	const size_t bytesEnd = static_cast<size_t>(H + 3) * WW * 4;
	for (size_t o = static_cast<size_t>(lane) * 128; o < bytesEnd && o < static_cast<size_t>(KHT_AHEAD + 2) * WW * 4; o += 32 * 128)
		asm volatile("prefetch.global.L1 [%0];" :: "l"(bytes0 + o));
	for (int y = 1; y < H - 1; ++y) {
		const unsigned int* row = bits + static_cast<size_t>(y) * WW + 1; // word 0 of the image row
		{
			const size_t o = static_cast<size_t>(y + 1 + KHT_AHEAD) * WW * 4 + static_cast<size_t>(lane) * 128; // row y + KHT_AHEAD, one 128-byte line per lane
			if (o < bytesEnd && lane * 128 < WW * 4 + 128) asm volatile("prefetch.global.L1 [%0];" :: "l"(bytes0 + o));
		}
		for (int wb = 0; wb <= lastWord; wb += 32) {
			while (true) {
				const int wi = wb + lane;
				unsigned int w = (wi <= lastWord) ? row[wi] : 0u; // plain load: served by this SM's L1, which the walker's stores keep current
				// seeds are interior columns only: x in [1, W-2]
				if (wi == 0) w &= ~1u;
				if (wi == lastWord) w &= ~(1u << ((W - 1) & 31));
				const unsigned int any = __ballot_sync(0xffffffffu, w != 0);
				if (!any) break;
				const int src = __ffs(any) - 1;
				const unsigned int sw = __shfl_sync(0xffffffffu, w, src);
				const int xr = (wb + src) * 32 + (__ffs(sw) - 1);
				unsigned int begin = 0, rev = 0, end = 0;
				if (lane == 0) {
					// Algorithm 5 (houghkht.cxx:706-760)
					begin = nPos;
					unsigned int* out = poss + nPos;
					KhtWalk wk;
					wk.xy = static_cast<unsigned int>(xr) | (static_cast<unsigned int>(y) << 16);
					kht_walk_load(wk, bits, WW);
					do {
						*out++ = wk.xy;
						kht_walk_erase(wk);
					} while (kht_walk_next(wk, bits, WW));
					rev = static_cast<unsigned int>(out - poss);
					wk.xy = static_cast<unsigned int>(xr) | (static_cast<unsigned int>(y) << 16);
					kht_walk_load(wk, bits, WW);
					if (kht_walk_next(wk, bits, WW)) {
						do {
							*out++ = wk.xy;
							kht_walk_erase(wk);
						} while (kht_walk_next(wk, bits, WW));
					}
					nPos = static_cast<unsigned int>(out - poss);
					end = nPos;
					if (end - begin < g.minSize) { nPos = begin; end = begin; }
					else strings[nStr++] = make_uint2(begin, end);
				}
				__syncwarp();
				begin = __shfl_sync(0xffffffffu, begin, 0);
				rev = __shfl_sync(0xffffffffu, rev, 0);
				end = __shfl_sync(0xffffffffu, end, 0);
				if (end > begin) { // the first walk is stored reversed (std::reverse, houghkht.cxx:752-755)
					const unsigned int n = rev - begin;
					for (unsigned int i = lane; i < n / 2; i += 32) {
						const unsigned int a = poss[begin + i], b = poss[begin + n - 1 - i];
						poss[begin + i] = b; poss[begin + n - 1 - i] = a;
					}
				}
				__syncwarp();
			}
		}
	}
	if (lane == 0) { fr.nPos = nPos; fr.nStr = nStr; }
}

// ---- linking (new): see kht_walk.cuh for the walker ----------------------------------------------
// One warp per frame: the 32 lanes find the next seed in raster order (32 bitmap words per ballot), lane 0 runs Algorithm 5 for it.  The first walk of a
// string is stored in walk order; the reversal the reference applies (std::reverse, houghkht.cxx:752-755) is left to kht_reverse_kernel, which is parallel
// over strings, instead of a load-after-store round trip through L2 between two walks.
template <bool REV, bool BF>
__global__ void __launch_bounds__(32)
kht_link_kernel(unsigned int* bitsAll /* read and written through derived pointers: no __restrict__ */, ushort2* __restrict__ possAll, uint2* __restrict__ stringsAll,
	unsigned int* __restrict__ revAll, KhtFrame* frames, KhtGeom g)
{
	const int frame = blockIdx.x, lane = threadIdx.x;
	const int W = g.W, H = g.H, WW = g.WW;
	unsigned int* base = bitsAll + (static_cast<size_t>(frame) * (H + 2 * KHT_PADR) + KHT_PADR) * WW; // padded word 0 of image row 0
	KhtFrame& fr = frames[frame];
	unsigned int* poss = reinterpret_cast<unsigned int*>(possAll + fr.posOff); // ushort2 {x, y} written as x | y << 16
	uint2* strings = stringsAll + fr.strOff;
	unsigned int* revs = revAll + fr.strOff;
	unsigned int nPos = 0, nStr = 0; // meaningful on lane 0
	const int lastWord = (W - 1) >> 5;

	// The seed scan reads rows in order, so rows above the scan line are in this SM's L1; walks mostly head DOWN into rows nobody has read yet and
	// would pay an L2 round trip per new row.  The idle lanes therefore keep KHT_AHEAD rows below the scan line prefetched.
	const char* bytes0 = reinterpret_cast<const char*>(base - KHT_PADR * WW);
	const size_t bytesEnd = static_cast<size_t>(H + 2 * KHT_PADR) * WW * 4;
	for (size_t o = static_cast<size_t>(lane) * 128; o < bytesEnd && o < static_cast<size_t>(KHT_AHEAD + KHT_PADR + 1) * WW * 4; o += 32 * 128)
		asm volatile("prefetch.global.L1 [%0];" :: "l"(bytes0 + o));
	for (int y = 1; y < H - 1; ++y) {
		const unsigned int* row = base + static_cast<size_t>(y) * WW + 1; // word 0 of the image row
		{
			const size_t o = static_cast<size_t>(y + KHT_PADR + KHT_AHEAD) * WW * 4 + static_cast<size_t>(lane) * 128; // row y + KHT_AHEAD, one 128-byte line per lane
			if (o < bytesEnd && lane * 128 < WW * 4 + 128) asm volatile("prefetch.global.L1 [%0];" :: "l"(bytes0 + o));
		}
		for (int wb = 0; wb <= lastWord; wb += 32) {
			while (true) {
				const int wi = wb + lane;
				unsigned int w = (wi <= lastWord) ? row[wi] : 0u; // plain load: served by this SM's L1, which the walker's stores keep current
				// seeds are interior columns only: x in [1, W-2]
				if (wi == 0) w &= ~kw_colbit<REV>(0);
				if (wi == lastWord) w &= ~kw_colbit<REV>((W - 1) & 31);
				const unsigned int any = __ballot_sync(0xffffffffu, w != 0);
				if (!any) break;
				const int src = __ffs(any) - 1;
				const unsigned int sw = __shfl_sync(0xffffffffu, w, src);
				if (lane == 0) {
					const int xr = (wb + src) * 32 + kw_first_col<REV>(sw);
					unsigned int rev;
					const unsigned int n = kht_link_string<REV, BF>(base, WW, static_cast<unsigned int>(xr) | (static_cast<unsigned int>(y) << 16), poss + nPos, &rev);
					if (n >= g.minSize) {
						strings[nStr] = make_uint2(nPos, nPos + n);
						revs[nStr] = rev;
						++nStr; nPos += n;
					}
				}
				__syncwarp();
			}
		}
	}
	if (lane == 0) { fr.nPos = nPos; fr.nStr = nStr; }
}

// the first walk of every string is stored in walk order: reverse it (houghkht.cxx:752-755).  One warp per string.
__global__ void kht_reverse_kernel(ushort2* __restrict__ possAll, const uint2* __restrict__ stringsAll, const unsigned int* __restrict__ revAll, const KhtFrame* frames)
{
	const int frame = blockIdx.y, lane = threadIdx.x & 31;
	const KhtFrame& fr = frames[frame];
	unsigned int* poss = reinterpret_cast<unsigned int*>(possAll + fr.posOff);
	const unsigned int warpsPerGrid = gridDim.x * (blockDim.x >> 5);
	for (unsigned int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < fr.nStr; s += warpsPerGrid) {
		const unsigned int begin = stringsAll[fr.strOff + s].x, n = revAll[fr.strOff + s];
		for (unsigned int i = lane; i < n / 2; i += 32) {
			const unsigned int a = poss[begin + i], b = poss[begin + n - 1 - i];
			poss[begin + i] = b; poss[begin + n - 1 - i] = a;
		}
	}
}

// ---- cluster subdivision (houghkht.cxx:774-832) -------------------------------------------------
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; } // std::max semantics (NaN in `a` is returned)

__global__ void kht_subdivide_kernel(const ushort2* __restrict__ possAll, const uint2* __restrict__ stringsAll, uint2* __restrict__ clusAll, unsigned int* __restrict__ nClusStrAll,
	KhtStack* __restrict__ stackAll, const KhtFrame* frames, KhtGeom g)
{
	const int frame = blockIdx.y;
	const KhtFrame& fr = frames[frame];
	const ushort2* poss = possAll + fr.posOff;
	for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < fr.nStr; s += gridDim.x * blockDim.x) {
		const uint2 str = stringsAll[fr.strOff + s];
		const ushort2* p = poss + str.x;
		uint2* clus = clusAll + fr.posOff + str.x;       // at most (len-1) clusters: the string's own index range is enough room
		KhtStack* st = stackAll + fr.posOff + str.x;
		unsigned int nclus = 0;
		int top = 0;
		double ret = 0.0;
		st[0].a = 0; st[0].b = (str.y - str.x) - 1; st[0].state = 0;
		while (top >= 0) {
			KhtStack& f = st[top];
			if (f.state == 0) {
				const unsigned int a = f.a, b = f.b;
				const int dxx = static_cast<int>(p[a].x) - static_cast<int>(p[b].x), dyy = static_cast<int>(p[a].y) - static_cast<int>(p[b].y);
				const double length = sqrt(static_cast<double>((dxx * dxx) + (dyy * dyy)));
				unsigned int mi = a; int md = 0;
				for (unsigned int i = a + 1; i < b; ++i) {
					const int d = abs(((static_cast<int>(p[a].x) - static_cast<int>(p[i].x)) * dyy) - ((static_cast<int>(p[a].y) - static_cast<int>(p[i].y)) * dxx));
					if (d > md) { mi = i; md = d; }
				}
				f.ratio = length / std_max((static_cast<double>(md) / length), g.minDeviation);
				f.mi = mi; f.nclus0 = nclus;
				if ((mi - a + 1) >= g.minSize && (b - mi + 1) >= g.minSize) {
					f.state = 1;
					KhtStack& c = st[top + 1];
					c.a = a; c.b = mi; c.state = 0;
					++top;
					continue;
				}
				nclus = f.nclus0;
				clus[nclus++] = make_uint2(str.x + a, str.x + b + 1);
				ret = f.ratio; --top;
			}
			else if (f.state == 1) {
				f.ratioLeft = ret; f.state = 2;
				KhtStack& c = st[top + 1];
				c.a = f.mi; c.b = f.b; c.state = 0;
				++top;
			}
			else {
				const double rl = f.ratioLeft, rr = ret;
				if (rl > f.ratio || rr > f.ratio) { ret = (rl > rr) ? rl : rr; --top; }
				else {
					nclus = f.nclus0;
					clus[nclus++] = make_uint2(str.x + f.a, str.x + f.b + 1);
					ret = f.ratio; --top;
				}
			}
		}
		nClusStrAll[fr.strOff + s] = nclus;
	}
}

// exclusive scan of the per-string cluster counts + ordered compaction of the clusters (one block per frame)
__global__ void __launch_bounds__(256)
kht_gather_clusters_kernel(const uint2* __restrict__ stringsAll, const uint2* __restrict__ clusAll, const unsigned int* __restrict__ nClusStrAll,
	uint2* __restrict__ clusOrdAll, KhtFrame* frames)
{
	__shared__ unsigned int sScan[256];
	__shared__ unsigned int sCarry;
	const int frame = blockIdx.x;
	KhtFrame& fr = frames[frame];
	if (threadIdx.x == 0) sCarry = 0;
	__syncthreads();
	for (unsigned int base = 0; base < fr.nStr; base += 256) {
		const unsigned int s = base + threadIdx.x;
		const unsigned int c = (s < fr.nStr) ? nClusStrAll[fr.strOff + s] : 0;
		sScan[threadIdx.x] = c;
		__syncthreads();
		for (int o = 1; o < 256; o <<= 1) {
			const unsigned int v = (threadIdx.x >= o) ? sScan[threadIdx.x - o] : 0;
			__syncthreads();
			sScan[threadIdx.x] += v;
			__syncthreads();
		}
		const unsigned int off = sCarry + sScan[threadIdx.x] - c;
		if (s < fr.nStr) {
			const uint2 str = stringsAll[fr.strOff + s];
			for (unsigned int j = 0; j < c; ++j) clusOrdAll[fr.posOff + off + j] = clusAll[fr.posOff + str.x + j];
		}
		__syncthreads();
		if (threadIdx.x == 255) sCarry += sScan[255];
		__syncthreads();
	}
	if (threadIdx.x == 0) fr.nClus = sCarry;
}

// ---- Algorithm 2: kernels (houghkht.cxx:885-1026) -----------------------------------------------
__device__ void eigen2x2(const double A[4], double D[4], double Q[4]) // base/math/compv_math_eigen.cxx:285-342
{
	bool norm = true;
	const double trace = A[0] + A[3];
	const double half = trace / 2.0;
	const double det = (A[0] * A[3]) - (A[1] * A[2]);
	const double s = sqrt(((trace * trace) / 4.0) - det);
	D[1] = D[2] = 0.0;
	D[0] = half + s;
	D[3] = half - s;
	if (A[2] != 0) { Q[0] = D[0] - A[3]; Q[2] = A[2]; Q[1] = D[3] - A[3]; Q[3] = A[2]; }
	else if (A[1] != 0) { Q[0] = A[1]; Q[2] = D[0] - A[0]; Q[1] = A[1]; Q[3] = D[3] - A[0]; }
	else {
		norm = false;
		if (A[3] != 0.0) { Q[0] = 0.0; Q[2] = 1.0; Q[1] = 1.0; Q[3] = 0.0; }
		else { Q[0] = 1.0; Q[2] = 0.0; Q[1] = 0.0; Q[3] = 1.0; }
	}
	if (norm) {
		const double m02 = 1.0 / sqrt(Q[0] * Q[0] + Q[2] * Q[2]);
		const double m13 = 1.0 / sqrt(Q[1] * Q[1] + Q[3] * Q[3]);
		Q[0] *= m02; Q[2] *= m02; Q[1] *= m13; Q[3] *= m13;
	}
	if (D[0] < D[3]) {
		double t = Q[0]; Q[0] = Q[1]; Q[1] = t;
		t = Q[2]; Q[2] = Q[3]; Q[3] = t;
		t = D[0]; D[0] = D[3]; D[3] = t;
	}
}

__device__ __forceinline__ double exp_small(double x) // houghkht.cxx:79-88
{
	x = 1.0 + (x * (1.0 / 1024.0));
#pragma unroll
	for (int i = 0; i < 10; ++i) x *= x;
	return x;
}

#define KHT_PI 3.14159265358979323846

// double -> int32 as the reference's x86 build does it (cvttsd2si): NaN and out-of-range values give INT_MIN ("integer indefinite"), which is what
// ends the voting loop when a degenerate kernel produces a huge or NaN density (houghkht.cxx:1124 `(votes = ...) > 0`); CUDA's cast would saturate.
__device__ __forceinline__ int x86_double_to_int32(double v)
{
	return (v > -2147483649.0 && v < 2147483648.0) ? static_cast<int>(v) : INT_MIN;
}

__global__ void kht_kernels_kernel(const ushort2* __restrict__ possAll, const uint2* __restrict__ clusOrdAll, KhtKernel* __restrict__ kernAll, KhtFrame* frames, KhtGeom g)
{
	const int frame = blockIdx.y;
	KhtFrame& fr = frames[frame];
	const ushort2* poss = possAll + fr.posOff;
	const unsigned int n = fr.nClus;
	const unsigned int pack = g.x86Simd ? (n >= 4 ? 4u : (n >= 2 ? 2u : 1u)) : 1u;
	const unsigned int simdCount = (pack > 1) ? (n & ~(pack - 1)) : 0;
	const double RAD2DEG = 180.0 / KHT_PI, TWOPI = 2.0 * KHT_PI;
	for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
		const uint2 cl = clusOrdAll[fr.posOff + c];
		const ushort2* pb = poss + cl.x;
		const unsigned int np = cl.y - cl.x;
		const double ns = 1.0 / static_cast<double>(np);
		double mx = 0, my = 0;
		for (unsigned int i = 0; i < np; ++i) { mx += (static_cast<int>(pb[i].x) - g.halfW); my += (static_cast<int>(pb[i].y) - g.halfH); }
		mx *= ns; my *= ns;
		double cxx = 0, cyy = 0, cxy = 0;
		for (unsigned int i = 0; i < np; ++i) {
			const double cx = (static_cast<int>(pb[i].x) - g.halfW) - mx, cy = (static_cast<int>(pb[i].y) - g.halfH) - my;
			cxx += (cx * cx); cyy += (cy * cy); cxy += (cx * cy);
		}
		const double M[4] = { cxx, cxy, cxy, cyy };
		double D[4], Q[4];
		eigen2x2(M, D, Q);
		const double ux = Q[0], uy = Q[2];
		double vx = Q[1], vy = Q[3];
		if (vy < 0.0) { vx = -vx; vy = -vy; }
		KhtKernel k;
		k.rho = (vx * mx) + (vy * my);
		k.theta = acos(vx) * RAD2DEG;
		const double s1 = sqrt(1.0 - (vx * vx));
		const double m0 = -(ux * mx) - (uy * my);
		const double m2e = (s1 == 0.0) ? 0.0 : ((ux / s1) * RAD2DEG);
		double r0 = 0.0;
		for (unsigned int i = 0; i < np; ++i) {
			const double t = (ux * ((static_cast<int>(pb[i].x) - g.halfW) - mx)) + (uy * ((static_cast<int>(pb[i].y) - g.halfH) - my));
			r0 += (t * t);
		}
		// heights (houghkht.cxx:849-884 and the SSE2/AVX leaves)
		const double q0 = 1.0 / r0;
		const double q1 = m0 * q0, q2 = m2e * q0;
		double srs = q1 * m0 + ns;
		const double srt = q1 * m2e;
		const double m2 = q2 * m0;
		double sts = q2 * m2e;
		if (sts == 0.0) sts = 0.1;
		srs *= 4.0; sts *= 4.0;
		const double sst = sqrt(srs) * sqrt(sts);
		const double rr = srt / sst;
		const double omr = 1.0 - (rr * rr);
		k.h = (c < simdCount) ? (1.0 / ((sqrt(omr) * sst) * TWOPI)) : (1.0 / (TWOPI * sst * sqrt(omr)));
		k.srs = srs; k.srt = srt; k.m2 = m2; k.sts = sts; k.alive = 1; k.pad = 0;
		kernAll[fr.posOff + c] = k;
		if (k.h > 0.0) atomicMax(&fr.hmaxBits, static_cast<unsigned long long>(__double_as_longlong(k.h))); // positive doubles order like their bit patterns
	}
}

// discard short kernels (houghkht.cxx:1029-1041) + Gmin (houghkht.cxx:1044-1062, Eq15 :834-847)
__global__ void kht_gmin_kernel(KhtKernel* __restrict__ kernAll, KhtFrame* frames, KhtGeom g)
{
	const int frame = blockIdx.y;
	KhtFrame& fr = frames[frame];
	const double hmax = __longlong_as_double(static_cast<long long>(fr.hmaxBits));
	const double scale = 1.0 / hmax;
	const double TWOPI = 2.0 * KHT_PI;
	for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < fr.nClus; c += gridDim.x * blockDim.x) {
		KhtKernel& q = kernAll[fr.posOff + c];
		if ((q.h * scale) < g.minHeight) { q.alive = 0; continue; }
		const double M[4] = { q.srs, q.srt, q.m2, q.sts };
		double D[4], Q[4];
		eigen2x2(M, D, Q);
		const double r1 = sqrt(D[3]);
		const double rh = Q[1] * r1, th = Q[3] * r1;
		const double sst = sqrt(q.srs) * sqrt(q.sts);
		const double sc = 1.0 / sst;
		const double rr = q.srt * sc;
		const double omr = 1.0 - (rr * rr);
		const double x = 1.0 / (TWOPI * sst * sqrt(omr));
		const double y = 1.0 / (2.0 * omr);
		const double z = ((rh * rh) / q.srs) - (((rr * 2.0) * rh * th) * sc) + ((th * th) / q.sts);
		const double gv = x * exp_small(-z * y);
		// `if (r2 < Gmin) Gmin = r2` with Gmin starting at DBL_MAX: NaN never wins, negative values cannot occur (x > 0 or NaN, exp_small >= 0)
		if (gv >= 0.0) atomicMin(&fr.gminBits, static_cast<unsigned long long>(__double_as_longlong(gv)));
	}
}

__global__ void kht_gs_kernel(KhtFrame* frames, int batch)
{
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= batch) return;
	const double gmin = __longlong_as_double(static_cast<long long>(frames[f].gminBits));
	frames[f].gs = (gmin == 0.0) ? 1.0 : std_max((1.0 / gmin), 1.0); // houghkht.cxx:377
}

// ---- voting (houghkht.cxx:1065-1148): one thread per (kernel, quadrant) --------------------------
__global__ void kht_vote_kernel(const KhtKernel* __restrict__ kernAll, int* __restrict__ accAll, const KhtFrame* frames, KhtGeom g)
{
	const int frame = blockIdx.y;
	const KhtFrame& fr = frames[frame];
	int* acc = accAll + static_cast<size_t>(frame) * (g.nTheta + 2) * g.cs;
	const double Gs = fr.gs;
	const double rhoScale = 1.0 / g.dRho, thetaScale = 1.0 / g.dThetaDeg;
	const unsigned long long nRho = g.nRho, nTheta = g.nTheta;
	for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < fr.nClus * 4u; t += gridDim.x * blockDim.x) {
		const KhtKernel q = kernAll[fr.posOff + (t >> 2)];
		if (!q.alive) continue;
		const int quad = t & 3;
		const unsigned long long ri = static_cast<unsigned long long>(__double2ull_rz(fabs((q.rho - g.rhoMaxNeg) * rhoScale) + 0.5)) + 1ull;
		const unsigned long long ti = static_cast<unsigned long long>(__double2ull_rz(fabs(q.theta * thetaScale) + 0.5)) + 1ull;
		unsigned long long rhoStartIdx = (quad & 2) ? ri - 1 : ri;
		unsigned long long thetaIdx = (quad & 1) ? ti - 1 : ti;
		const double rhoStart = (quad & 2) ? -g.dRho : 0.0;
		const double thetaStart = (quad & 1) ? -g.dThetaDeg : 0.0;
		long long incRhoIdx = (quad & 2) ? -1 : 1;
		const long long incThetaIdx = (quad & 1) ? -1 : 1;
		const double incRho = g.dRho * static_cast<double>((quad & 2) ? -1 : 1), incTheta = g.dThetaDeg * static_cast<double>(incThetaIdx);
		const double srsS = 1.0 / q.srs, stsS = 1.0 / q.sts;
		const double sst = sqrt(q.srs) * sqrt(q.sts);
		const double sc = 1.0 / sst;
		const double rr = q.srt * sc;
		const double omr = 1.0 - (rr * rr);
		const double r2 = rr * 2.0;
		const double x = 1.0 / ((2.0 * KHT_PI) * sst * sqrt(omr));
		const double y = 1.0 / (2.0 * omr);
		unsigned long long thetaCount = 0;
		double th = thetaStart, rh = rhoStart;
		do {
			if (!thetaIdx || thetaIdx > nTheta) {
				rhoStartIdx = (nRho - rhoStartIdx) + 1ull;
				thetaIdx = thetaIdx ? 1ull : nTheta;
				incRhoIdx = -incRhoIdx;
			}
			if (rhoStartIdx >= 1ull) {
				int* pc = acc + thetaIdx * g.cs;
				unsigned long long rhoIdx = rhoStartIdx;
				rh = rhoStart;
				const double wv = (th * th) * stsS;
				const double kk = r2 * th * sc;
				double krho = kk * rh;
				const double ki = kk * incRho;
				double z = ((rh * rh) * srsS) - krho + wv;
				int votes;
				while (rhoIdx <= nRho && (votes = x86_double_to_int32(((x * exp_small(-z * y)) * Gs) + 0.5)) > 0) {
					atomicAdd(&pc[rhoIdx], votes);
					rhoIdx += static_cast<unsigned long long>(incRhoIdx);
					rh += incRho;
					krho += ki;
					z = ((rh * rh) * srsS) - krho + wv;
				}
				thetaIdx += static_cast<unsigned long long>(incThetaIdx);
				th += incTheta;
			}
			else break;
		} while ((rh != rhoStart) && (++thetaCount < nTheta));
	}
}

// ---- peaks: 3x3 smoothing + threshold (houghkht.cxx:1282-1308), cells kept in the reference's scan order ----------------
// cell (ti, ri) qualifies when count > 0 and smoothed >= threshold, and is scanned at all (see the SSE2 coverage note in oracle/compv_oracle_kht.cpp)
__device__ __forceinline__ bool kht_cell(const int* acc, const KhtGeom& g, unsigned int ti, unsigned int ri, int& v)
{
	const int* c = acc + static_cast<size_t>(ti) * g.cs + ri;
	if (!(*c > 0)) return false;
	const int* t = c - g.cs; const int* b = c + g.cs;
	v = t[-1] + (t[0] << 1) + t[1] + b[-1] + (b[0] << 1) + b[1] + (c[-1] << 1) + (c[0] << 2) + (c[1] << 1);
	return v >= g.threshold;
}

__device__ __forceinline__ bool kht_scanned(const KhtGeom& g, unsigned int ri, unsigned int& reported)
{
	reported = ri;
	if (ri < 1 || ri >= g.nRho) return false;
	if (!(g.x86Simd && g.nRho > 4)) return true;
	const unsigned int sseEnd = g.nRho - 3;                  // SSE starts 1, 5, ... < sseEnd, each covering 4 cells
	const unsigned int lastStart = (sseEnd > 1) ? (1 + 4 * ((sseEnd - 2) / 4)) : 0;
	if (lastStart && ri <= lastStart + 3) return true;
	const unsigned int consumed = (g.nRho & ~3u) + 1;
	if (g.nRho > consumed && ri > consumed && ri < g.nRho) { reported = ri - consumed; return true; } // scalar tail reports pointer-relative indices
	return false;
}

// pass 1: per (frame, theta row) count of qualifying cells; pass 2 (after a host-side-free device scan) writes them in order
__global__ void kht_peaks_count_kernel(const int* __restrict__ accAll, unsigned int* __restrict__ rowCount, KhtGeom g)
{
	const int frame = blockIdx.y;
	const unsigned int ti = blockIdx.x + 1;
	if (ti >= g.nTheta) return;
	const int* acc = accAll + static_cast<size_t>(frame) * (g.nTheta + 2) * g.cs;
	unsigned int n = 0;
	for (unsigned int ri = 1 + threadIdx.x; ri < g.nRho; ri += blockDim.x) {
		unsigned int rep; int v;
		if (kht_scanned(g, ri, rep) && kht_cell(acc, g, ti, ri, v)) ++n;
	}
	__shared__ unsigned int sSum;
	if (threadIdx.x == 0) sSum = 0;
	__syncthreads();
	for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
	if ((threadIdx.x & 31) == 0 && n) atomicAdd(&sSum, n);
	__syncthreads();
	if (threadIdx.x == 0) rowCount[frame * (g.nTheta + 2) + ti] = sSum;
}

struct KhtVote { unsigned int rho_index, theta_index; int count; };

__global__ void __launch_bounds__(256)
kht_peaks_emit_kernel(const int* __restrict__ accAll, const unsigned int* __restrict__ rowCount, KhtVote* __restrict__ votesAll, unsigned int votesCap, KhtFrame* frames, KhtGeom g)
{
	// one block per frame: rows in order, cells of a row in ascending rho by chunks of 256 with a block scan
	__shared__ unsigned int sScan[256];
	__shared__ unsigned int sBase;
	const int frame = blockIdx.x;
	const int* acc = accAll + static_cast<size_t>(frame) * (g.nTheta + 2) * g.cs;
	KhtVote* votes = votesAll + static_cast<size_t>(frame) * votesCap;
	if (threadIdx.x == 0) sBase = 0;
	__syncthreads();
	for (unsigned int ti = 1; ti < g.nTheta; ++ti) {
		if (rowCount[frame * (g.nTheta + 2) + ti] == 0) continue; // block-uniform
		for (unsigned int base = 1; base < g.nRho; base += 256) {
			const unsigned int ri = base + threadIdx.x;
			unsigned int rep = 0; int v = 0;
			const unsigned int ok = (kht_scanned(g, ri, rep) && kht_cell(acc, g, ti, ri, v)) ? 1u : 0u;
			sScan[threadIdx.x] = ok;
			__syncthreads();
			for (int o = 1; o < 256; o <<= 1) {
				const unsigned int u = (threadIdx.x >= o) ? sScan[threadIdx.x - o] : 0;
				__syncthreads();
				sScan[threadIdx.x] += u;
				__syncthreads();
			}
			const unsigned int pos = sBase + sScan[threadIdx.x] - ok;
			if (ok && pos < votesCap) { KhtVote o; o.rho_index = rep; o.theta_index = ti; o.count = v; votes[pos] = o; }
			__syncthreads();
			if (threadIdx.x == 255) sBase += sScan[255];
			__syncthreads();
		}
	}
	if (threadIdx.x == 0) frames[frame].nVotes = sBase;
}

} // namespace cvb

using namespace cvb;


int cvb::kht_process_dev(cvb200_hough* h, const uint8_t* edges, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch,
	cvb200_hough_line_t* lines, size_t capacity, size_t* counts, cudaStream_t stream)
{
	const auto tCall0 = std::chrono::steady_clock::now();
	CVB_REQUIRE(width <= 65535 && height <= 65535, CVB200_E_OUT_OF_BOUND); // positions are stored as 16-bit coordinates
	CVB_REQUIRE(h->clusterMinSize >= 2, CVB200_E_INVALID_PARAMETER);       // 1 makes the reference's recursion endless
	KhtGeom g;
	memset(&g, 0, sizeof(g));
	g.W = static_cast<int>(width); g.H = static_cast<int>(height); g.WW = static_cast<int>(div_up(width, 32) + 2);
	g.stride = stride; g.framePitch = framePitch;
	// ctor + initCoords (houghkht.cxx:113-116, 501-541)
	const float kPi = 3.1415926535897932384626433f;
	const float kPiOver180 = kPi / 180.f;
	g.dRho = static_cast<double>(h->rho * 1.f);
	const double dThetaRad = static_cast<double>(h->theta * kPiOver180);
	g.dThetaDeg = (dThetaRad * 180.0) / M_PI;
	const double r = std::sqrt(static_cast<double>((width * width) + (height * height)));
	const size_t nRho = static_cast<size_t>((r + 1.0) / g.dRho);
	const size_t nTheta = static_cast<size_t>(180.0 / g.dThetaDeg);
	CVB_REQUIRE(nRho >= 2 && nTheta >= 2 && nRho < (1u << 24) && nTheta < (1u << 16), CVB200_E_INVALID_PARAMETER);
	g.nRho = static_cast<unsigned int>(nRho); g.nTheta = static_cast<unsigned int>(nTheta); g.cs = g.nRho + 2;
	std::vector<double> rho(nRho + 1, 0.0), theta(nTheta + 1, 0.0);
	{ double v = -(r * 0.5); for (size_t i = 1; i < nRho; ++i, v += g.dRho) rho[i] = v; }
	{ double v = 0.0; for (size_t i = 1; i < nTheta; ++i, v += g.dThetaDeg) theta[i] = v; }
	g.rhoMaxNeg = rho[1];
	g.halfW = static_cast<double>(width) * 0.5; g.halfH = static_cast<double>(height) * 0.5;
	g.minDeviation = static_cast<double>(h->clusterMinDeviation); g.minHeight = static_cast<double>(h->kernelMinHeight);
	g.minSize = static_cast<unsigned int>(h->clusterMinSize);
	g.threshold = static_cast<int>(h->threshold);
	g.x86Simd = h->x86Simd ? 1 : 0;

	static const int linkVariant = getenv("CVB200_KHT_LINK") ? atoi(getenv("CVB200_KHT_LINK")) : 2; // TEMPORARY A/B switch: 0 = round-1 walker, 1 = window walker, 2 = window walker on bit-reversed words
	// ---- phase 1: bitmap + edge counts ----
	const size_t bitWords = static_cast<size_t>(g.H + 2 * KHT_PADR) * g.WW;
	CVB_CHECK(h->bits.ensure(batch * bitWords * 4));
	CVB_CUDA(cudaMemsetAsync(h->bits.p, 0, batch * bitWords * 4, stream)); // the zero border rows / words the walker relies on
	CVB_CHECK(h->edgeCount.ensure(batch * 4));
	CVB_CHECK(h->hCounts.ensure(batch * 4));
	CVB_CUDA(cudaMemsetAsync(h->edgeCount.p, 0, batch * 4, stream));
	{
		dim3 grid(static_cast<unsigned>(div_up(g.WW, 64)), static_cast<unsigned>(g.H), static_cast<unsigned>(batch));
		CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
		KernelScope ks_("kht_bits", stream);
		if (linkVariant == 2 || linkVariant == 3) kht_bits_kernel<true><<<grid, 64, 0, stream>>>(edges, h->bits.as<unsigned int>(), g, h->edgeCount.as<unsigned int>());
		else kht_bits_kernel<false><<<grid, 64, 0, stream>>>(edges, h->bits.as<unsigned int>(), g, h->edgeCount.as<unsigned int>());
	}
	CVB_LAUNCHED();
	unsigned int* hCounts = h->hCounts.as<unsigned int>();
	CVB_CUDA(cudaMemcpyAsync(hCounts, h->edgeCount.p, batch * 4, cudaMemcpyDeviceToHost, stream));
	CVB_CUDA(cudaStreamSynchronize(stream));

	// ---- pools sized by the edge counts ----
	CVB_CHECK(h->hFrames.ensure(batch * sizeof(KhtFrame)));
	KhtFrame* hf = h->hFrames.as<KhtFrame>();
	size_t posTotal = 0, strTotal = 0;
	const unsigned long long dblMaxBits = static_cast<unsigned long long>(0x7FEFFFFFFFFFFFFFull);
	for (size_t f = 0; f < batch; ++f) {
		memset(&hf[f], 0, sizeof(KhtFrame));
		hf[f].posOff = static_cast<unsigned int>(posTotal); hf[f].posCap = hCounts[f];
		hf[f].strOff = static_cast<unsigned int>(strTotal); hf[f].strCap = hCounts[f] / g.minSize + 1;
		hf[f].gminBits = dblMaxBits; hf[f].gs = 1.0;
		posTotal += hCounts[f] + 2; strTotal += hf[f].strCap;
		CVB_REQUIRE(posTotal < (1ull << 31) && strTotal < (1ull << 31), CVB200_E_OUT_OF_BOUND);
	}
	const size_t accCells = static_cast<size_t>(g.nTheta + 2) * g.cs;
	// every accumulator cell may qualify (threshold 1 on a busy frame); only when that worst case gets large is the list capped (overflow is detected below)
	size_t votesCap = accCells;
	if (batch * accCells * sizeof(KhtVote) > (256u << 20)) { votesCap = accCells / 4; if (votesCap < 65536) votesCap = 65536; if (votesCap > accCells) votesCap = accCells; }
	CVB_CHECK(h->frames.ensure(batch * sizeof(KhtFrame)));
	CVB_CHECK(h->poss.ensure((posTotal + 1) * sizeof(ushort2)));
	CVB_CHECK(h->strings.ensure((strTotal + 1) * sizeof(uint2)));
	CVB_CHECK(h->nClusStr.ensure((strTotal + 1) * 4));
	CVB_CHECK(h->clus.ensure((posTotal + 1) * sizeof(uint2)));
	CVB_CHECK(h->clusOrd.ensure((posTotal + 1) * sizeof(uint2)));
	CVB_CHECK(h->stack.ensure((posTotal + 1) * sizeof(KhtStack)));
	CVB_CHECK(h->kern.ensure((posTotal + 1) * sizeof(KhtKernel)));
	CVB_CHECK(h->acc.ensure(batch * accCells * 4));
	CVB_CHECK(h->rowCount.ensure(batch * (g.nTheta + 2) * 4));
	CVB_CHECK(h->votes.ensure(batch * votesCap * sizeof(KhtVote)));
	CVB_CHECK(h->hVotes.ensure(batch * votesCap * sizeof(KhtVote)));
	CVB_CUDA(cudaMemcpyAsync(h->frames.p, hf, batch * sizeof(KhtFrame), cudaMemcpyHostToDevice, stream));
	CVB_CUDA(cudaMemsetAsync(h->acc.p, 0, batch * accCells * 4, stream));
	KhtFrame* dFrames = h->frames.as<KhtFrame>();
	const unsigned int B = static_cast<unsigned int>(batch);

	CVB_CHECK(h->strRev.ensure((strTotal + 1) * 4));
	if (linkVariant == 0) {
		KernelScope ks_("kht_link", stream);
		kht_link_old_kernel<<<B, 32, 0, stream>>>(h->bits.as<unsigned int>(), h->poss.as<ushort2>(), h->strings.as<uint2>(), dFrames, g);
		CVB_LAUNCHED();
	}
	else {
		{ KernelScope ks_("kht_link", stream);
		  unsigned int* bp = h->bits.as<unsigned int>(); ushort2* pp = h->poss.as<ushort2>(); uint2* sp = h->strings.as<uint2>(); unsigned int* rp = h->strRev.as<unsigned int>();
		  if (linkVariant == 1) kht_link_kernel<false, false><<<B, 32, 0, stream>>>(bp, pp, sp, rp, dFrames, g);
		  else if (linkVariant == 2) kht_link_kernel<true, false><<<B, 32, 0, stream>>>(bp, pp, sp, rp, dFrames, g);
		  else if (linkVariant == 3) kht_link_kernel<true, true><<<B, 32, 0, stream>>>(bp, pp, sp, rp, dFrames, g);
		  else kht_link_kernel<false, true><<<B, 32, 0, stream>>>(bp, pp, sp, rp, dFrames, g); }
		CVB_LAUNCHED();
		{ KernelScope ks_("kht_reverse", stream);
		  kht_reverse_kernel<<<dim3(8, B), 128, 0, stream>>>(h->poss.as<ushort2>(), h->strings.as<uint2>(), h->strRev.as<unsigned int>(), dFrames); }
		CVB_LAUNCHED();
	}
	{ KernelScope ks_("kht_subdivide", stream);
	  kht_subdivide_kernel<<<dim3(32, B), 64, 0, stream>>>(h->poss.as<ushort2>(), h->strings.as<uint2>(), h->clus.as<uint2>(), h->nClusStr.as<unsigned int>(), h->stack.as<KhtStack>(), dFrames, g); }
	CVB_LAUNCHED();
	{ KernelScope ks_("kht_gather_clusters", stream);
	  kht_gather_clusters_kernel<<<B, 256, 0, stream>>>(h->strings.as<uint2>(), h->clus.as<uint2>(), h->nClusStr.as<unsigned int>(), h->clusOrd.as<uint2>(), dFrames); }
	CVB_LAUNCHED();
	{ KernelScope ks_("kht_kernels", stream);
	  kht_kernels_kernel<<<dim3(32, B), 64, 0, stream>>>(h->poss.as<ushort2>(), h->clusOrd.as<uint2>(), h->kern.as<KhtKernel>(), dFrames, g); }
	CVB_LAUNCHED();
	{ KernelScope ks_("kht_gmin", stream);
	  kht_gmin_kernel<<<dim3(32, B), 64, 0, stream>>>(h->kern.as<KhtKernel>(), dFrames, g); }
	CVB_LAUNCHED();
	{ KernelScope ks_("kht_gs", stream);
	  kht_gs_kernel<<<static_cast<unsigned>(div_up(batch, 64)), 64, 0, stream>>>(dFrames, static_cast<int>(batch)); }
	CVB_LAUNCHED();
	{ KernelScope ks_("kht_vote", stream);
	  kht_vote_kernel<<<dim3(64, B), 64, 0, stream>>>(h->kern.as<KhtKernel>(), h->acc.as<int>(), dFrames, g); }
	CVB_LAUNCHED();
	{ KernelScope ks_("kht_peaks_count", stream);
	  kht_peaks_count_kernel<<<dim3(g.nTheta, B), 128, 0, stream>>>(h->acc.as<int>(), h->rowCount.as<unsigned int>(), g); }
	CVB_LAUNCHED();
	{ KernelScope ks_("kht_peaks_emit", stream);
	  kht_peaks_emit_kernel<<<B, 256, 0, stream>>>(h->acc.as<int>(), h->rowCount.as<unsigned int>(), h->votes.as<KhtVote>(), static_cast<unsigned int>(votesCap), dFrames, g); }
	CVB_LAUNCHED();
	CVB_CUDA(cudaMemcpyAsync(hf, dFrames, batch * sizeof(KhtFrame), cudaMemcpyDeviceToHost, stream));
	CVB_CUDA(cudaStreamSynchronize(stream));
	for (size_t f = 0; f < batch; ++f) CVB_REQUIRE(hf[f].nVotes <= votesCap, CVB200_E_OUT_OF_BOUND);
	KhtVote* hv = h->hVotes.as<KhtVote>();
	{
		size_t maxVotes = 0; // one strided copy for the whole batch: the leading maxVotes cells of every frame's list
		for (size_t f = 0; f < batch; ++f) maxVotes = std::max<size_t>(maxVotes, hf[f].nVotes);
		if (maxVotes) CVB_CUDA(cudaMemcpy2DAsync(hv, votesCap * sizeof(KhtVote), h->votes.p, votesCap * sizeof(KhtVote), maxVotes * sizeof(KhtVote), batch, cudaMemcpyDeviceToHost, stream));
	}
	CVB_CUDA(cudaStreamSynchronize(stream));
	const auto tHost0 = std::chrono::steady_clock::now();

	// ---- host: sort + sweep (houghkht.cxx:1195-1247). std::sort of the same libstdc++ on the same input order = the reference's tie order ----
	// Frames are independent: a few host threads share them (the sort of a few thousand cells per frame is the only per-frame host work).
	const size_t lim = (h->maxLines <= 0) ? static_cast<size_t>(INT_MAX) : static_cast<size_t>(h->maxLines);
	host_parallel_for(batch, [&](size_t f) {
		static thread_local std::vector<uint8_t> visited; // all zero between frames (the sweep clears what it marked)
		if (visited.size() < accCells) visited.assign(accCells, 0);

		KhtVote* v = hv + f * votesCap;
		const size_t nv = hf[f].nVotes;
		std::sort(v, v + nv, [](const KhtVote& a, const KhtVote& b) { return a.count > b.count; });
		size_t n = 0;
		for (size_t i = 0; i < nv; ++i) {
			uint8_t* pv = &visited[static_cast<size_t>(v[i].theta_index) * g.cs + v[i].rho_index];
			const uint8_t* t = pv - g.cs; const uint8_t* b = pv + g.cs;
			const bool seen = t[-1] || t[0] || t[1] || pv[-1] || pv[1] || b[-1] || b[0] || b[1];
			if (!seen) {
				if (n < lim) {
					if (n < capacity) {
						cvb200_hough_line_t& L = lines[f * capacity + n];
						L.rho = static_cast<float>(rho[v[i].rho_index]);
						L.theta = static_cast<float>((theta[v[i].theta_index] * M_PI) / 180.0);
						L.strength = static_cast<size_t>(v[i].count);
					}
					++n;
				}
			}
			*pv = 0xff;
		}
		for (size_t i = 0; i < nv; ++i) visited[static_cast<size_t>(v[i].theta_index) * g.cs + v[i].rho_index] = 0;
		counts[f] = n;
	});
	{
		const size_t f = batch - 1;
		h->lastGs = (hf[f].nStr && hf[f].nClus) ? hf[f].gs : 1.0;
	}
	if (getenv("CVB200_TRACE")) {
		const auto t1 = std::chrono::steady_clock::now();
		size_t nv = 0; for (size_t f = 0; f < batch; ++f) nv += hf[f].nVotes;
		fprintf(stderr, "[cvb200] kht batch %zu: device+copies %.3f ms, host peaks %.3f ms (%zu cells)\n", batch,
			std::chrono::duration<double, std::milli>(tHost0 - tCall0).count(), std::chrono::duration<double, std::milli>(t1 - tHost0).count(), nv);
	}
	return CVB200_S_OK;
}

extern "C" {

int cvb200_hough_new(cvb200_hough_t** hough, int id, float rho, float theta, size_t threshold)
{
	CVB_REQUIRE(hough, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(id == CVB200_HOUGHKHT_ID || id == CVB200_HOUGHSHT_ID, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(rho > 0.f && rho <= 1.f, CVB200_E_INVALID_PARAMETER); // houghkht.cxx:493, houghsht.cxx:310
	CVB_REQUIRE(id == CVB200_HOUGHKHT_ID || rho == 1.f, CVB200_E_INVALID_PARAMETER); // houghsht.cxx:312: the SHT requires rho == 1
	cvb200_hough* h = new (std::nothrow) cvb200_hough();
	CVB_REQUIRE(h, CVB200_E_OUT_OF_MEMORY);
	h->id = id; h->rho = rho; h->theta = theta; h->threshold = threshold;
	h->maxLines = INT_MAX;
	h->clusterMinDeviation = 2.0f; h->clusterMinSize = 10; h->kernelMinHeight = 0.002f; // houghkht.cxx:36-38
	h->x86Simd = true; h->lastGs = 1.0;
	*hough = h;
	return CVB200_S_OK;
}

int cvb200_hough_free(cvb200_hough_t** hough)
{
	if (hough && *hough) {
		cvb200_hough* h = *hough;
		DevBuf* bufs[] = { &h->bits, &h->poss, &h->strings, &h->strRev, &h->clus, &h->clusOrd, &h->nClusStr, &h->stack, &h->kern, &h->acc, &h->rowCount, &h->votes, &h->frames, &h->edgeCount, &h->hostIn,
			&h->shtTables, &h->shtList, &h->shtCursor, &h->shtMask, &h->shtPool, &h->shtDesc };
		for (DevBuf* b : bufs) b->release();
		h->hFrames.release(); h->hVotes.release(); h->hCounts.release();
		delete h;
		*hough = nullptr;
	}
	return CVB200_S_OK;
}

// houghkht.cxx:140-192
int cvb200_hough_set(cvb200_hough_t* h, int id, const void* valuePtr, size_t valueSize)
{
	CVB_REQUIRE(h && valuePtr && valueSize, CVB200_E_INVALID_PARAMETER);
	switch (id) {
	case CVB200_HOUGH_SET_FLT32_RHO: {
		CVB_REQUIRE(valueSize == sizeof(float), CVB200_E_INVALID_PARAMETER);
		const float v = *static_cast<const float*>(valuePtr);
		CVB_REQUIRE(v > 0.f && v <= 1.f, CVB200_E_INVALID_PARAMETER);
		CVB_REQUIRE(h->id == CVB200_HOUGHKHT_ID || v == 1.f, CVB200_E_INVALID_PARAMETER); // houghsht.cxx:71
		h->rho = v; return CVB200_S_OK;
	}
	case CVB200_HOUGH_SET_FLT32_THETA: {
		CVB_REQUIRE(valueSize == sizeof(float) && *static_cast<const float*>(valuePtr) > 0.f, CVB200_E_INVALID_PARAMETER);
		h->theta = *static_cast<const float*>(valuePtr); return CVB200_S_OK;
	}
	case CVB200_HOUGH_SET_INT_THRESHOLD: {
		CVB_REQUIRE(valueSize == sizeof(int) && *static_cast<const int*>(valuePtr) > 0, CVB200_E_INVALID_PARAMETER);
		h->threshold = static_cast<size_t>(*static_cast<const int*>(valuePtr)); return CVB200_S_OK;
	}
	case CVB200_HOUGH_SET_INT_MAXLINES: {
		CVB_REQUIRE(valueSize == sizeof(int), CVB200_E_INVALID_PARAMETER);
		const int v = *static_cast<const int*>(valuePtr);
		h->maxLines = v <= 0 ? INT_MAX : v; return CVB200_S_OK;
	}
	case CVB200_HOUGHKHT_SET_FLT32_CLUSTER_MIN_DEVIATION:
		CVB_REQUIRE(h->id == CVB200_HOUGHKHT_ID, CVB200_E_NOT_IMPLEMENTED);
		CVB_REQUIRE(valueSize == sizeof(float), CVB200_E_INVALID_PARAMETER);
		h->clusterMinDeviation = *static_cast<const float*>(valuePtr); return CVB200_S_OK;
	case CVB200_HOUGHKHT_SET_INT_CLUSTER_MIN_SIZE:
		CVB_REQUIRE(h->id == CVB200_HOUGHKHT_ID, CVB200_E_NOT_IMPLEMENTED);
		CVB_REQUIRE(valueSize == sizeof(int) && *static_cast<const int*>(valuePtr) > 0, CVB200_E_INVALID_PARAMETER);
		h->clusterMinSize = *static_cast<const int*>(valuePtr); return CVB200_S_OK;
	case CVB200_HOUGHKHT_SET_FLT32_KERNEL_MIN_HEIGTH:
		CVB_REQUIRE(h->id == CVB200_HOUGHKHT_ID, CVB200_E_NOT_IMPLEMENTED);
		CVB_REQUIRE(valueSize == sizeof(float) && *static_cast<const float*>(valuePtr) >= 0.f, CVB200_E_INVALID_PARAMETER);
		h->kernelMinHeight = *static_cast<const float*>(valuePtr); return CVB200_S_OK;
	case CVB200_HOUGHKHT_SET_BOOL_OVERRIDE_INPUT_EDGES:
		CVB_REQUIRE(valueSize == sizeof(bool), CVB200_E_INVALID_PARAMETER);
		return CVB200_S_OK; // the device path never modifies the caller's edges: nothing to override
	case CVB200_HOUGH_SET_BOOL_X86_SIMD_SCAN:
		CVB_REQUIRE(valueSize == sizeof(bool), CVB200_E_INVALID_PARAMETER);
		h->x86Simd = *static_cast<const bool*>(valuePtr); return CVB200_S_OK;
	default:
		return CVB200_E_NOT_IMPLEMENTED;
	}
}

// houghkht.cxx:194-206
int cvb200_hough_get(cvb200_hough_t* h, int id, void* valuePtr, size_t valueSize)
{
	CVB_REQUIRE(h && valuePtr && valueSize, CVB200_E_INVALID_PARAMETER);
	if (id == CVB200_HOUGHKHT_GET_FLT64_GS && h->id == CVB200_HOUGHKHT_ID) {
		CVB_REQUIRE(valueSize == sizeof(double), CVB200_E_INVALID_PARAMETER);
		*static_cast<double*>(valuePtr) = h->lastGs;
		return CVB200_S_OK;
	}
	return CVB200_E_NOT_IMPLEMENTED;
}

int cvb200_hough_process_dev(cvb200_hough_t* h, const uint8_t* edges, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch,
	cvb200_hough_line_t* lines, size_t capacity, size_t* counts, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(h && edges && counts && width && height && stride >= width && (lines || !capacity), CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	std::lock_guard<std::mutex> lock(h->mutex);
	for (size_t f = 0; f < batch; ++f) counts[f] = 0;
	if (h->id == CVB200_HOUGHKHT_ID) return kht_process_dev(h, edges, width, height, stride, batch, framePitch, lines, capacity, counts, as_stream(stream));
	return sht_process_dev(h, edges, width, height, stride, batch, framePitch, lines, capacity, counts, as_stream(stream));
}

int cvb200_hough_process(cvb200_hough_t* h, const uint8_t* edges, size_t width, size_t height, size_t stride, cvb200_hough_line_t* lines, size_t capacity, size_t* count)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(h && edges && count && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	const size_t n = stride * height;
	{
		std::lock_guard<std::mutex> lock(h->mutex);
		CVB_CHECK(h->hostIn.ensure(n));
	}
	CVB_CUDA(cudaMemcpyAsync(h->hostIn.p, edges, n, cudaMemcpyHostToDevice, 0));
	return cvb200_hough_process_dev(h, h->hostIn.as<uint8_t>(), width, height, stride, 1, n, lines, capacity, count, nullptr);
}

} // extern "C"
