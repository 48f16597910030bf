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
#include "std_sort_emu.cuh"

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
	unsigned long long voteOff;       // this frame's cells in the vote / sort pools
	unsigned int skip, pad;           // set when the position pools are too small: the linking kernel does nothing
};
struct KhtStack { unsigned int a, b, mi, nclus0; double ratio, ratioLeft; unsigned int state, pad; };

// ---- bitmap -------------------------------------------------------------------------------------
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
		bits[(static_cast<size_t>(frame) * (g.H + 2 * KHT_PADR) + y + KHT_PADR) * g.WW + wi + 1] = __brev(word); // bit 31 = leftmost column (kht_walk.cuh)
	}
	unsigned int c = __popc(word);
	for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(&edgeCount[frame], c);
}

#define KHT_AHEAD 96
// ---- linking: see kht_walk.cuh for the walker ----------------------------------------------
// One warp per frame: the 32 lanes find the next seed in raster order (32 bitmap words per ballot), lane 0 runs Algorithm 5 for it.  The first walk of a
// string is stored in walk order; the reversal the reference applies (std::reverse, houghkht.cxx:752-755) is left to kht_reverse_kernel, which is parallel
// over strings, instead of a load-after-store round trip through L2 between two walks.
__global__ void __launch_bounds__(32)
kht_link_kernel(unsigned int* bitsAll /* read and written through derived pointers: no __restrict__ */, ushort2* __restrict__ possAll, uint2* __restrict__ stringsAll,
	unsigned int* __restrict__ revAll, KhtFrame* frames, KhtGeom g, int ahead)
{
	const int frame = blockIdx.x, lane = threadIdx.x;
	const int W = g.W, H = g.H, WW = g.WW;
	unsigned int* base = bitsAll + (static_cast<size_t>(frame) * (H + 2 * KHT_PADR) + KHT_PADR) * WW; // padded word 0 of image row 0
	KhtFrame& fr = frames[frame];
	if (fr.skip) return;
	unsigned int* poss = reinterpret_cast<unsigned int*>(possAll + fr.posOff); // ushort2 {x, y} written as x | y << 16
	uint2* strings = stringsAll + fr.strOff;
	unsigned int* revs = revAll + fr.strOff;
	// the walker forms every address as pointer + 32-bit offset: keep the two pointers in registers (otherwise they are re-derived from the kernel parameters
	// with four 64-bit instructions and a constant-bank load per access)
	asm volatile("" : "+l"(base));
	asm volatile("" : "+l"(poss));
	unsigned int nPos = 0, nStr = 0; // meaningful on lane 0
	const int lastWord = (W - 1) >> 5;

	// The seed scan reads rows in order, so rows above the scan line are in this SM's L1; walks mostly head DOWN into rows nobody has read yet and
	// would pay an L2 round trip per new row.  The idle lanes therefore keep KHT_AHEAD rows below the scan line prefetched.
	const char* bytes0 = reinterpret_cast<const char*>(base - KHT_PADR * WW);
	const size_t bytesEnd = static_cast<size_t>(H + 2 * KHT_PADR) * WW * 4;
	for (size_t o = static_cast<size_t>(lane) * 128; o < bytesEnd && o < static_cast<size_t>(ahead + KHT_PADR + 1) * WW * 4; o += 32 * 128)
		asm volatile("prefetch.global.L1 [%0];" :: "l"(bytes0 + o));
	for (int y = 1; y < H - 1; ++y) {
		const unsigned int* row = base + static_cast<size_t>(y) * WW + 1; // word 0 of the image row
		{
			const size_t o = static_cast<size_t>(y + KHT_PADR + ahead) * WW * 4 + static_cast<size_t>(lane) * 128; // row y + ahead, one 128-byte line per lane
			if (o < bytesEnd && lane * 128 < WW * 4 + 128) asm volatile("prefetch.global.L1 [%0];" :: "l"(bytes0 + o));
		}
		for (int wb = 0; wb <= lastWord; wb += 32) {
			while (true) {
				const int wi = wb + lane;
				unsigned int w = (wi <= lastWord) ? row[wi] : 0u; // plain load: served by this SM's L1, which the walker's stores keep current
				// seeds are interior columns only: x in [1, W-2]
				if (wi == 0) w &= ~kw_colbit(0);
				if (wi == lastWord) w &= ~kw_colbit((W - 1) & 31);
				const unsigned int any = __ballot_sync(0xffffffffu, w != 0);
				if (!any) break;
				const int src = __ffs(any) - 1;
				const unsigned int sw = __shfl_sync(0xffffffffu, w, src);
				if (lane == 0) {
					const int xr = (wb + src) * 32 + kw_first_col(sw);
					unsigned int rev;
					const unsigned int n = kht_link_string(base, WW, static_cast<unsigned int>(xr) | (static_cast<unsigned int>(y) << 16), poss + nPos, &rev);
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

// ---- linking, 32 frames per warp (kht_walk.cuh: KhtLane) ----
// Every lane links its own frame; the warp goes round one loop whose body is "scan two bitmap words" or "one step of the walk".  Used when a launch holds enough frames
// for the batch to be the parallelism: per frame it issues ~1/20 of the instructions of the kernel above (whose 31 idle lanes still cost an issue slot each instruction),
// at about the same latency per launch.  Same strings in the same order: the scan order, the walker and the records are the ones above.
__global__ void __launch_bounds__(32)
kht_link_lanes_kernel(unsigned int* bitsAll, ushort2* __restrict__ possAll, uint2* __restrict__ stringsAll, unsigned int* __restrict__ revAll, KhtFrame* frames, KhtGeom g, int batch)
{
	const int frame = blockIdx.x * 32 + threadIdx.x;
	if (frame >= batch) return;
	const int W = g.W, H = g.H, WW = g.WW;
	KhtFrame& fr = frames[frame];
	if (fr.skip) return;
	unsigned int* base = bitsAll + (static_cast<size_t>(frame) * (H + 2 * KHT_PADR) + KHT_PADR) * WW; // padded word 0 of image row 0
	unsigned int* poss = reinterpret_cast<unsigned int*>(possAll + fr.posOff);
	unsigned long long* strs = reinterpret_cast<unsigned long long*>(stringsAll + fr.strOff);
	unsigned int* revs = revAll + fr.strOff;
	asm volatile("" : "+l"(base));
	asm volatile("" : "+l"(poss));
	KhtLane L;
	L.start(H);
	while (L.phase != 3) L.iterate(base, WW, W, H, g.minSize, poss, strs, revs);
	fr.nPos = L.nPos; fr.nStr = L.nStr;
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

// Peak detection without the host (the reference: peaks_Section3_4, houghkht.cxx:1195-1247 and :1282-1308):
//   kht_peaks_count   per (frame, theta row): number of qualifying cells
//   kht_row_offsets   per frame: exclusive scan of its rows + the frame's total
//   kht_offsets2      exclusive scan over the frames of the batch -> where each frame's cells go in the shared pools (overflow -> flag, the host grows the pools and retries)
//   kht_peaks_emit    per (frame, theta row): the cells in the reference's scan order (row-major, ascending rho)
//   kht_peaks_sort    per frame: libstdc++'s std::sort permutation (std_sort_emu.cuh) evaluated by the warps of one CTA, then the sweep as a rank comparison, then the lines
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

// block-wide exclusive scan helper (blockDim.x == 256): returns the exclusive prefix of v, *total = the block's sum
__device__ __forceinline__ unsigned int block_scan_256(unsigned int v, unsigned int* sWarp /* [9] shared */, unsigned int* total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned int inc = v;
	for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
	__syncthreads(); // protects sWarp against the previous use
	if (lane == 31) sWarp[warp] = inc;
	__syncthreads();
	if (threadIdx.x == 0) { unsigned int run = 0; for (int w = 0; w < 8; ++w) { const unsigned int t = sWarp[w]; sWarp[w] = run; run += t; } sWarp[8] = run; }
	__syncthreads();
	*total = sWarp[8];
	return sWarp[warp] + inc - v;
}

// rowCount[frame][ti] (counts) -> exclusive offsets inside the frame; frames[frame].nVotes = the frame's total
__global__ void __launch_bounds__(256) kht_row_offsets_kernel(unsigned int* __restrict__ rowCount, KhtFrame* frames, KhtGeom g)
{
	__shared__ unsigned int sWarp[9];
	const int frame = blockIdx.x;
	unsigned int* rc = rowCount + frame * (g.nTheta + 2);
	unsigned int carry = 0;
	for (unsigned int base = 1; base < g.nTheta; base += 256) {
		const unsigned int ti = base + threadIdx.x;
		const unsigned int c = (ti < g.nTheta) ? rc[ti] : 0u;
		unsigned int total;
		const unsigned int ex = block_scan_256(c, sWarp, &total);
		if (ti < g.nTheta) rc[ti] = carry + ex;
		carry += total;
	}
	if (threadIdx.x == 0) frames[frame].nVotes = carry;
}

struct KhtMeta { unsigned int overflow; unsigned int maxVotes /* largest per-frame cell count: sizes the next call's shared-memory sort */; unsigned long long needPos, needStr, needVotes; };

// per-frame offsets into the position / string pools from the edge counts (one block; batch frames)
__global__ void __launch_bounds__(256) kht_offsets1_kernel(const unsigned int* __restrict__ edgeCount, KhtFrame* frames, KhtMeta* meta, int batch, unsigned int minSize,
	unsigned long long posCap, unsigned long long strCap)
{
	__shared__ unsigned int sWarp[9];
	unsigned long long posRun = 0, strRun = 0;
	for (int base = 0; base < batch; base += 256) {
		const int f = base + threadIdx.x;
		const unsigned int c = (f < batch) ? edgeCount[f] : 0u;
		const unsigned int pc = (f < batch) ? c + 2u : 0u, sc = (f < batch) ? c / minSize + 1u : 0u;
		unsigned int pt, st;
		const unsigned int pe = block_scan_256(pc, sWarp, &pt);
		const unsigned int se = block_scan_256(sc, sWarp, &st);
		if (f < batch) {
			KhtFrame fr;
			memset(&fr, 0, sizeof(fr));
			fr.posOff = static_cast<unsigned int>(posRun + pe); fr.posCap = c;
			fr.strOff = static_cast<unsigned int>(strRun + se); fr.strCap = sc;
			fr.gminBits = 0x7FEFFFFFFFFFFFFFull; fr.gs = 1.0;
			frames[f] = fr;
		}
		posRun += pt; strRun += st;
	}
	__syncthreads();
	const bool over = posRun + 1 > posCap || strRun + 1 > strCap || posRun >= (1ull << 31) || strRun >= (1ull << 31);
	if (threadIdx.x == 0) { meta->needPos = posRun + 1; meta->needStr = strRun + 1; meta->needVotes = 0; meta->overflow = over ? 1u : 0u; }
	if (over) for (int f = threadIdx.x; f < batch; f += 256) frames[f].skip = 1; // nothing downstream runs: every per-frame count stays 0
}

__global__ void __launch_bounds__(256) kht_offsets2_kernel(KhtFrame* frames, KhtMeta* meta, int batch, unsigned long long voteCap)
{
	__shared__ unsigned int sWarp[9];
	unsigned long long run = 0;
	for (int base = 0; base < batch; base += 256) {
		const int f = base + threadIdx.x;
		const unsigned int c = (f < batch) ? frames[f].nVotes : 0u;
		unsigned int t;
		const unsigned int e = block_scan_256(c, sWarp, &t);
		if (f < batch) frames[f].voteOff = run + e;
		run += t;
		if (c) atomicMax(&meta->maxVotes, c);
	}
	if (threadIdx.x == 0) { meta->needVotes = run + 1; if (run + 1 > voteCap) meta->overflow |= 2u; }
}

struct KhtVote { unsigned int rho_index, theta_index; int count; };

__global__ void __launch_bounds__(256)
kht_peaks_emit_kernel(const int* __restrict__ accAll, const unsigned int* __restrict__ rowOff, KhtVote* __restrict__ votesAll, const KhtFrame* frames, const KhtMeta* meta, KhtGeom g)
{
	__shared__ unsigned int sWarp[9];
	if (meta->overflow) return;
	const int frame = blockIdx.y;
	const unsigned int ti = blockIdx.x + 1;
	if (ti >= g.nTheta) return;
	const unsigned int* ro = rowOff + frame * (g.nTheta + 2);
	const unsigned int next = (ti + 1 < g.nTheta) ? ro[ti + 1] : frames[frame].nVotes;
	if (next == ro[ti]) return; // block-uniform: nothing in this row
	const int* acc = accAll + static_cast<size_t>(frame) * (g.nTheta + 2) * g.cs;
	KhtVote* votes = votesAll + frames[frame].voteOff + ro[ti];
	unsigned int carry = 0;
	for (unsigned int base = 1; base < g.nRho; base += 256) {
		const unsigned int ri = base + threadIdx.x;
		unsigned int rep = 0; int v = 0;
		const unsigned int ok = (kht_scanned(g, ri, rep) && kht_cell(acc, g, ti, ri, v)) ? 1u : 0u;
		unsigned int total;
		const unsigned int ex = block_scan_256(ok, sWarp, &total);
		if (ok) { KhtVote o; o.rho_index = rep; o.theta_index = ti; o.count = v; votes[carry + ex] = o; }
		carry += total;
	}
}

// ---- std::sort's permutation + the sweep, one CTA per frame ----
#define KSORT_THREADS 256
#define KSORT_WARPS (KSORT_THREADS / 32)
#define KSORT_SMEM_ITEMS 8192 // up to this many cells the whole sort runs in shared memory: 8192 * (8 + 2 + 2) bytes = 96 KB (16-bit position lists)

// one partition step of a[first, last) by one warp: the closed form of std_sort_emu.cuh (sse_partition_closed_form) with ballots for the two ordered compactions
template <typename IDX>
__device__ int ksort_warp_partition(sse_item* a, int first, int last, IDX* Ls, IDX* Rs)
{
	const int lane = threadIdx.x & 31;
	const unsigned int lt = (1u << lane) - 1u;
	if (lane == 0) sse_move_median_to_first(a, first, first + 1, first + (last - first) / 2, last - 1);
	__syncwarp();
	const unsigned int pk = sse_key(a[first]);
	int nL = 0, nR = 0;
	for (int base = first + 1; base < last; base += 32) {
		const int i = base + lane;
		const bool flag = i < last && !(sse_key(a[i]) > pk);
		const unsigned int bal = __ballot_sync(0xffffffffu, flag);
		if (flag) Ls[first + nL + __popc(bal & lt)] = static_cast<IDX>(i);
		nL += __popc(bal);
	}
	for (int base = last - 1; base > first; base -= 32) {
		const int i = base - lane;
		const bool flag = i > first && !(pk > sse_key(a[i]));
		const unsigned int bal = __ballot_sync(0xffffffffu, flag);
		if (flag) Rs[first + nR + __popc(bal & lt)] = static_cast<IDX>(i);
		nR += __popc(bal);
	}
	__syncwarp();
	const int m = nL < nR ? nL : nR;
	int K = 0;
	for (int base = 0; base < m; base += 32) { // L ascends and R descends: the predicate is monotone, stop at the first chunk that contains a false
		const int k = base + lane;
		const bool ok = k < m && static_cast<unsigned int>(Ls[first + k]) < static_cast<unsigned int>(Rs[first + k]);
		const unsigned int bal = __ballot_sync(0xffffffffu, ok);
		K += __popc(bal);
		if (bal != 0xffffffffu) break;
	}
	for (int k = lane; k < K; k += 32) { const unsigned int i = Ls[first + k], j = Rs[first + k]; const sse_item t = a[i]; a[i] = a[j]; a[j] = t; }
	const unsigned int cl = (K < nL) ? static_cast<unsigned int>(Ls[first + K]) : 0xffffffffu;
	const unsigned int cr = (K > 0) ? static_cast<unsigned int>(Rs[first + K - 1]) : static_cast<unsigned int>(last);
	__syncwarp();
	return static_cast<int>(cl < cr ? cl : cr);
}

// The same partition step by the whole CTA, for the large ranges at the top of the recursion tree (the first three levels of a 5.7 k-cell frame are 1, 2 and 4 ranges:
// one warp each would leave most of the CTA waiting at the level barrier).  Every warp scans a contiguous chunk; the two ordered lists are the concatenation of the
// chunks' lists (L ascending: chunks in order; R descending: chunks in reverse order), so they are exactly the lists of ksort_warp_partition.
#define KSORT_BIG 1024
template <typename IDX>
__device__ int ksort_block_partition(sse_item* a, int first, int last, IDX* Ls, IDX* Rs, int* sPart /* 24 ints of shared memory */)
{
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const unsigned int lt = (1u << lane) - 1u;
	if (tid == 0) sse_move_median_to_first(a, first, first + 1, first + (last - first) / 2, last - 1);
	__syncthreads();
	const unsigned int pk = sse_key(a[first]);
	const int n = last - first - 1;
	const int per = ((n + KSORT_WARPS * 32 - 1) / (KSORT_WARPS * 32)) * 32;
	const int cb = min(first + 1 + warp * per, last), ce = min(cb + per, last); // my chunk of first+1 .. last-1
	int cL = 0, cR = 0;
	for (int base = cb; base < ce; base += 32) {
		const int i = base + lane;
		const bool in = i < ce;
		const unsigned int k = in ? sse_key(a[i]) : 0u;
		cL += __popc(__ballot_sync(0xffffffffu, in && !(k > pk)));
		cR += __popc(__ballot_sync(0xffffffffu, in && !(pk > k)));
	}
	if (lane == 0) { sPart[warp] = cL; sPart[8 + warp] = cR; }
	__syncthreads();
	int offL = 0, offR = 0, nL = 0, nR = 0;
	for (int w = 0; w < KSORT_WARPS; ++w) {
		const int l_ = sPart[w], r_ = sPart[8 + w];
		if (w < warp) offL += l_;
		if (w > warp) offR += r_;
		nL += l_; nR += r_;
	}
	int pos = offL;
	for (int base = cb; base < ce; base += 32) {
		const int i = base + lane;
		const bool flag = i < ce && !(sse_key(a[i]) > pk);
		const unsigned int bal = __ballot_sync(0xffffffffu, flag);
		if (flag) Ls[first + pos + __popc(bal & lt)] = static_cast<IDX>(i);
		pos += __popc(bal);
	}
	pos = offR;
	for (int base = ce - 1; base >= cb; base -= 32) {
		const int i = base - lane;
		const bool flag = i >= cb && !(pk > sse_key(a[i]));
		const unsigned int bal = __ballot_sync(0xffffffffu, flag);
		if (flag) Rs[first + pos + __popc(bal & lt)] = static_cast<IDX>(i);
		pos += __popc(bal);
	}
	__syncthreads();
	const int m = nL < nR ? nL : nR;
	int cnt = 0; // L ascends and R descends: L[k] < R[k] holds for k < K and for no other k
	for (int k = tid; k < m; k += KSORT_THREADS) cnt += (static_cast<unsigned int>(Ls[first + k]) < static_cast<unsigned int>(Rs[first + k])) ? 1 : 0;
	for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	if (lane == 0) sPart[16 + warp] = cnt;
	__syncthreads();
	int K = 0;
	for (int w = 0; w < KSORT_WARPS; ++w) K += sPart[16 + w];
	for (int k = tid; k < K; k += KSORT_THREADS) { const unsigned int i = Ls[first + k], j = Rs[first + k]; const sse_item t = a[i]; a[i] = a[j]; a[j] = t; }
	const unsigned int cl = (K < nL) ? static_cast<unsigned int>(Ls[first + K]) : 0xffffffffu;
	const unsigned int cr = (K > 0) ? static_cast<unsigned int>(Rs[first + K - 1]) : static_cast<unsigned int>(last);
	__syncthreads();
	return static_cast<int>(cl < cr ? cl : cr);
}

__global__ void __launch_bounds__(KSORT_THREADS)
kht_peaks_sort_kernel(const KhtVote* __restrict__ votesAll, sse_item* __restrict__ itemsAll, unsigned int* __restrict__ listsAll, int* __restrict__ rangesAll, int* __restrict__ accAll,
	const double* __restrict__ rhoTab, const double* __restrict__ thetaTab, cvb200_hough_line_t* __restrict__ lines, unsigned long long capacity, unsigned long long* __restrict__ counts,
	const KhtFrame* frames, const KhtMeta* meta, KhtGeom g, unsigned int lim, int smemItems)
{
	extern __shared__ __align__(16) unsigned char ksortSmem[];
	__shared__ unsigned int sWarp[9];
	__shared__ int sCnt[2], sLeaves, sBigN, sBig[72][3], sPart[24];
	const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (meta->overflow) { if (tid == 0) counts[frame] = 0; return; }
	const KhtFrame& fr = frames[frame];
	const int nv = static_cast<int>(fr.nVotes);
	if (nv == 0) { if (tid == 0) counts[frame] = 0; return; }
	const KhtVote* votes = votesAll + fr.voteOff;
	const bool inSmem = nv <= smemItems;               // block-uniform
	sse_item* a; unsigned int* Ls = nullptr; unsigned int* Rs = nullptr; unsigned short* Ls16 = nullptr; unsigned short* Rs16 = nullptr;
	if (inSmem) { a = reinterpret_cast<sse_item*>(ksortSmem); Ls16 = reinterpret_cast<unsigned short*>(a + smemItems); Rs16 = Ls16 + smemItems; }
	else { a = itemsAll + fr.voteOff; Ls = listsAll + 2 * fr.voteOff; Rs = Ls + nv; }
	for (int i = tid; i < nv; i += KSORT_THREADS) a[i] = (static_cast<sse_item>(static_cast<unsigned int>(votes[i].count)) << 32) | static_cast<unsigned int>(i);
	__syncthreads();

	if (nv <= 96) { if (tid == 0) sse_sort_serial(a, nv); }
	else {
		// level-synchronous evaluation of the recursion tree: ranges of one level are disjoint and independent
		int* leaves = rangesAll + 2 * fr.voteOff;             // (first, last) pairs of ranges with 2..16 elements: at most nv / 2 of them
		int* lvl[2] = { leaves + nv, leaves + nv + nv / 2 };   // (first, last, depth) triples of ranges > 16 elements: at most nv / 17 per level, room for nv / 6
		// phase 1: ranges above KSORT_BIG cells, one at a time by the whole CTA (a stack: at most one pending sibling per level of the depth limit)
		if (tid == 0) { sBig[0][0] = 0; sBig[0][1] = nv; sBig[0][2] = sse_lg(static_cast<unsigned int>(nv)) * 2; sBigN = 1; sCnt[0] = 0; sCnt[1] = 0; sLeaves = 0; }
		__syncthreads();
		for (;;) {
			const int nb = sBigN;
			if (nb == 0) break;
			const int f = sBig[nb - 1][0], l = sBig[nb - 1][1], d = sBig[nb - 1][2];
			__syncthreads(); // everybody has read the top of the stack
			int cut = -1;
			if (l - f <= KSORT_BIG) { /* small from the start: straight to the level lists */ }
			else if (d == 0) { if (tid == 0) sse_heap_sort(a + f, l - f); } // the depth-limit fallback of introsort
			else cut = inSmem ? ksort_block_partition<unsigned short>(a, f, l, Ls16, Rs16, sPart) : ksort_block_partition<unsigned int>(a, f, l, Ls, Rs, sPart);
			if (tid == 0) {
				int top = nb - 1;
				const int cf[2] = { cut >= 0 ? f : f, cut >= 0 ? cut : 0 }, cl[2] = { cut >= 0 ? cut : l, cut >= 0 ? l : 0 };
				const int nChild = cut >= 0 ? 2 : ((l - f <= KSORT_BIG) ? 1 : 0), dc = cut >= 0 ? d - 1 : d;
				for (int c = 0; c < nChild; ++c) {
					const int len = cl[c] - cf[c];
					if (len > KSORT_BIG && cut >= 0) { sBig[top][0] = cf[c]; sBig[top][1] = cl[c]; sBig[top][2] = dc; ++top; }
					else if (len > SSE_THRESHOLD) { const int k = sCnt[0]++; lvl[0][3 * k] = cf[c]; lvl[0][3 * k + 1] = cl[c]; lvl[0][3 * k + 2] = dc; }
					else if (len > 1) { const int k = sLeaves++; leaves[2 * k] = cf[c]; leaves[2 * k + 1] = cl[c]; }
				}
				sBigN = top;
			}
			__syncthreads();
		}
		for (int cur = 0;; cur ^= 1) {
			const int cnt = sCnt[cur];
			if (cnt == 0) break;
			for (int r = warp; r < cnt; r += KSORT_WARPS) {
				const int f = lvl[cur][3 * r], l = lvl[cur][3 * r + 1], d = lvl[cur][3 * r + 2];
				if (d == 0) { if (lane == 0) sse_heap_sort(a + f, l - f); __syncwarp(); continue; } // the depth-limit fallback of introsort
				const int cut = inSmem ? ksort_warp_partition<unsigned short>(a, f, l, Ls16, Rs16) : ksort_warp_partition<unsigned int>(a, f, l, Ls, Rs);
				if (lane < 2) {
					const int cf = lane ? cut : f, cl = lane ? l : cut;
					if (cl - cf > SSE_THRESHOLD) { const int k = atomicAdd(&sCnt[cur ^ 1], 1); lvl[cur ^ 1][3 * k] = cf; lvl[cur ^ 1][3 * k + 1] = cl; lvl[cur ^ 1][3 * k + 2] = d - 1; }
					else if (cl - cf > 1) { const int k = atomicAdd(&sLeaves, 1); leaves[2 * k] = cf; leaves[2 * k + 1] = cl; }
				}
			}
			__syncthreads();
			if (tid == 0) sCnt[cur] = 0;
			__syncthreads();
		}
		for (int k = tid; k < sLeaves; k += KSORT_THREADS) sse_insertion_sort(a, leaves[2 * k], leaves[2 * k + 1]);
	}
	__syncthreads();

	// ---- the sweep (houghkht.cxx:1207-1247).  The reference marks EVERY cell it visits, line or not, so "a neighbour was visited" == "a neighbour comes earlier in
	// the sorted order": cell p is a line iff no vote of smaller sorted rank sits on one of its 8 neighbouring (theta, reported rho) positions.  The accumulator of this
	// frame has been consumed by kht_peaks_emit and is reused as the rank map.
	int* map = accAll + static_cast<size_t>(frame) * (g.nTheta + 2) * g.cs;
	const int cs = static_cast<int>(g.cs);
	for (int p = tid; p < nv; p += KSORT_THREADS) {
		const KhtVote v = votes[static_cast<unsigned int>(a[p])];
		int* c = map + static_cast<size_t>(v.theta_index) * cs + v.rho_index;
		c[-cs - 1] = INT_MAX; c[-cs] = INT_MAX; c[-cs + 1] = INT_MAX; c[-1] = INT_MAX; c[0] = INT_MAX; c[1] = INT_MAX; c[cs - 1] = INT_MAX; c[cs] = INT_MAX; c[cs + 1] = INT_MAX;
	}
	__syncthreads();
	for (int p = tid; p < nv; p += KSORT_THREADS) {
		const KhtVote v = votes[static_cast<unsigned int>(a[p])];
		atomicMin(map + static_cast<size_t>(v.theta_index) * cs + v.rho_index, p);
	}
	__syncthreads();
	unsigned int nLines = 0;
	for (int base = 0; base < nv; base += KSORT_THREADS) {
		const int p = base + tid;
		unsigned int isLine = 0;
		KhtVote v; v.rho_index = 0; v.theta_index = 0; v.count = 0;
		if (p < nv) {
			v = votes[static_cast<unsigned int>(a[p])];
			const int* c = map + static_cast<size_t>(v.theta_index) * cs + v.rho_index;
			const bool seen = c[-cs - 1] < p || c[-cs] < p || c[-cs + 1] < p || c[-1] < p || c[1] < p || c[cs - 1] < p || c[cs] < p || c[cs + 1] < p;
			isLine = seen ? 0u : 1u;
		}
		unsigned int total;
		const unsigned int ex = block_scan_256(isLine, sWarp, &total);
		const unsigned long long k = static_cast<unsigned long long>(nLines) + ex;
		if (isLine && k < lim && k < capacity) {
			cvb200_hough_line_t L;
			L.rho = static_cast<float>(rhoTab[v.rho_index]);
			L.theta = static_cast<float>((thetaTab[v.theta_index] * KHT_PI) / 180.0);
			L.strength = static_cast<size_t>(v.count);
			lines[static_cast<unsigned long long>(frame) * capacity + k] = L;
		}
		nLines += total;
	}
	if (tid == 0) counts[frame] = nLines < lim ? nLines : lim;
}

} // namespace cvb

using namespace cvb;

// Everything up to the lines runs on the device with no host interaction: pools are shared by the frames of the batch, their per-frame offsets come from
// device-side scans, and a pool that turns out too small raises a flag (nothing is written out of bounds) that kht_finish answers by growing it and running again.
int cvb::kht_enqueue(cvb200_hough* h, const uint8_t* edges, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch, size_t capacity, cudaStream_t stream)
{
	CVB_REQUIRE(width <= 65535 && height <= 65535, CVB200_E_OUT_OF_BOUND); // positions are stored as 16-bit coordinates
	CVB_REQUIRE(h->clusterMinSize >= 2, CVB200_E_INVALID_PARAMETER);       // 1 makes the reference's recursion endless
	CVB_REQUIRE(batch < (1u << 20), CVB200_E_OUT_OF_BOUND);
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
	g.rhoMaxNeg = -(r * 0.5);
	g.halfW = static_cast<double>(width) * 0.5; g.halfH = static_cast<double>(height) * 0.5;
	g.minDeviation = static_cast<double>(h->clusterMinDeviation); g.minHeight = static_cast<double>(h->kernelMinHeight);
	g.minSize = static_cast<unsigned int>(h->clusterMinSize);
	g.threshold = static_cast<int>(h->threshold);
	g.x86Simd = h->x86Simd ? 1 : 0;
	// the rho / theta tables (accumulated additions, houghkht.cxx:519-537) live on the device; re-made only when the geometry changes
	if (h->tabRho != g.dRho || h->tabTheta != g.dThetaDeg || h->tabR != r || !h->tabs.p) {
		CVB_CUDA(cudaStreamSynchronize(stream)); // the previous upload (if any) has left the pinned staging buffer
		CVB_CHECK(h->hTabs.ensure((nRho + nTheta + 2) * sizeof(double)));
		CVB_CHECK(h->tabs.ensure((nRho + nTheta + 2) * sizeof(double)));
		double* rho = h->hTabs.as<double>(); double* theta = rho + nRho + 1;
		rho[0] = 0.0; theta[0] = 0.0;
		{ double v = -(r * 0.5); for (size_t i = 1; i <= nRho; ++i, v += g.dRho) rho[i] = (i < nRho) ? v : 0.0; }
		{ double v = 0.0; for (size_t i = 1; i <= nTheta; ++i, v += g.dThetaDeg) theta[i] = (i < nTheta) ? v : 0.0; }
		CVB_CUDA(cudaMemcpyAsync(h->tabs.p, rho, (nRho + nTheta + 2) * sizeof(double), cudaMemcpyHostToDevice, stream));
		h->tabRho = g.dRho; h->tabTheta = g.dThetaDeg; h->tabR = r; h->tabNRho = nRho;
	}
	const double* dRhoTab = h->tabs.as<double>(); const double* dThetaTab = dRhoTab + nRho + 1;

	// ---- pools (grow-only; first call: a guess, afterwards whatever the largest call needed) ----
	const size_t px = width * height * batch;
	if (h->posCapEl < batch * 4) h->posCapEl = std::max<size_t>(px / 16, batch * 4 + 1024);      // ~6 % edge pixels
	if (h->strCapEl < batch * 2) h->strCapEl = std::max<size_t>(h->posCapEl / g.minSize + batch, batch * 2 + 1024);
	if (h->voteCapEl < 1024) h->voteCapEl = std::max<size_t>(batch * 8192, 65536);
	const size_t bitWords = static_cast<size_t>(g.H + 2 * KHT_PADR) * g.WW;
	const size_t accCells = static_cast<size_t>(g.nTheta + 2) * g.cs;
	CVB_CHECK(h->bits.ensure(batch * bitWords * 4));
	CVB_CHECK(h->edgeCount.ensure(batch * 4 + sizeof(KhtMeta)));
	CVB_CHECK(h->frames.ensure(batch * sizeof(KhtFrame)));
	CVB_CHECK(h->poss.ensure(h->posCapEl * sizeof(ushort2)));
	CVB_CHECK(h->strings.ensure(h->strCapEl * sizeof(uint2)));
	CVB_CHECK(h->strRev.ensure(h->strCapEl * 4));
	CVB_CHECK(h->nClusStr.ensure(h->strCapEl * 4));
	CVB_CHECK(h->clus.ensure(h->posCapEl * sizeof(uint2)));
	CVB_CHECK(h->clusOrd.ensure(h->posCapEl * sizeof(uint2)));
	CVB_CHECK(h->stack.ensure(h->posCapEl * sizeof(KhtStack)));
	CVB_CHECK(h->kern.ensure(h->posCapEl * sizeof(KhtKernel)));
	CVB_CHECK(h->acc.ensure(batch * accCells * 4));
	CVB_CHECK(h->rowCount.ensure(batch * (g.nTheta + 2) * 4));
	CVB_CHECK(h->votes.ensure(h->voteCapEl * sizeof(KhtVote)));
	CVB_CHECK(h->sortItems.ensure(h->voteCapEl * sizeof(sse_item)));
	CVB_CHECK(h->sortLists.ensure(h->voteCapEl * 2 * 4));
	CVB_CHECK(h->sortRanges.ensure(h->voteCapEl * 2 * 4));
	CVB_CHECK(h->dLines.ensure(std::max<size_t>(batch * capacity, 1) * sizeof(cvb200_hough_line_t)));
	CVB_CHECK(h->dCounts.ensure(batch * 8));
	CVB_CHECK(h->hFrames.ensure(batch * sizeof(KhtFrame) + sizeof(KhtMeta)));
	CVB_CHECK(h->hCounts.ensure(batch * 8));
	unsigned int* dEdgeCount = h->edgeCount.as<unsigned int>();
	KhtMeta* dMeta = reinterpret_cast<KhtMeta*>(h->edgeCount.as<unsigned char>() + ((batch * 4 + 15) & ~static_cast<size_t>(15)));
	CVB_CHECK(h->edgeCount.ensure(((batch * 4 + 15) & ~static_cast<size_t>(15)) + sizeof(KhtMeta)));
	dEdgeCount = h->edgeCount.as<unsigned int>();
	dMeta = reinterpret_cast<KhtMeta*>(h->edgeCount.as<unsigned char>() + ((batch * 4 + 15) & ~static_cast<size_t>(15)));
	KhtFrame* dFrames = h->frames.as<KhtFrame>();
	const unsigned int B = static_cast<unsigned int>(batch);

	const bool bitsReady = h->bitsPrepared; // the producer of the edge map (canny_finalize in the pipeline) has already written the bitmap and the edge counts
	h->bitsPrepared = false;
	if (!bitsReady) {
		CVB_CUDA(cudaMemsetAsync(h->bits.p, 0, batch * bitWords * 4, stream)); // the zero border rows / words the walker relies on
		CVB_CUDA(cudaMemsetAsync(dEdgeCount, 0, ((batch * 4 + 15) & ~static_cast<size_t>(15)) + sizeof(KhtMeta), stream));
	}
	CVB_CUDA(cudaMemsetAsync(h->acc.p, 0, batch * accCells * 4, stream));
	if (!bitsReady) {
		dim3 grid(static_cast<unsigned>(div_up(g.WW, 64)), static_cast<unsigned>(g.H), B);
		CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
		KernelScope ks_("kht_bits", stream);
		kht_bits_kernel<<<grid, 64, 0, stream>>>(edges, h->bits.as<unsigned int>(), g, dEdgeCount);
		CVB_LAUNCHED();
	}
	{ KernelScope ks_("kht_offsets", stream);
	  kht_offsets1_kernel<<<1, 256, 0, stream>>>(dEdgeCount, dFrames, dMeta, static_cast<int>(batch), g.minSize, h->posCapEl, h->strCapEl); }
	CVB_LAUNCHED();
	trace_mark(stream, "link>", h->traceSlot);
	{
		// enough frames in one launch: one LANE per frame (issue-efficient); few frames: one WARP per frame (its 32 lanes share the seed scan: lower latency)
		const char* ev = getenv("CVB200_KHT_LANES_MIN"); // read per call: tests switch it
		const int lanesMin = (ev && *ev) ? atoi(ev) : INT_MAX; // measured (B200, 2048 1080p frames): 114 ms per launch against 18.2 ms for one warp per frame -- off unless asked for
		const char* ea = getenv("CVB200_KHT_AHEAD");
		const int ahead = (ea && *ea) ? atoi(ea) : KHT_AHEAD;
		KernelScope ks_("kht_link", stream);
		if (static_cast<long long>(batch) >= lanesMin)
			kht_link_lanes_kernel<<<static_cast<unsigned>(div_up(batch, 32)), 32, 0, stream>>>(h->bits.as<unsigned int>(), h->poss.as<ushort2>(), h->strings.as<uint2>(), h->strRev.as<unsigned int>(), dFrames, g, static_cast<int>(batch));
		else
			kht_link_kernel<<<B, 32, 0, stream>>>(h->bits.as<unsigned int>(), h->poss.as<ushort2>(), h->strings.as<uint2>(), h->strRev.as<unsigned int>(), dFrames, g, ahead);
	}
	CVB_LAUNCHED();
	trace_mark(stream, "link<", h->traceSlot);
	{ KernelScope ks_("kht_reverse", stream);
	  kht_reverse_kernel<<<dim3(8, B), 128, 0, stream>>>(h->poss.as<ushort2>(), h->strings.as<uint2>(), h->strRev.as<unsigned int>(), dFrames); }
	CVB_LAUNCHED();
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
	{ KernelScope ks_("kht_offsets", stream);
	  kht_row_offsets_kernel<<<B, 256, 0, stream>>>(h->rowCount.as<unsigned int>(), dFrames, g); }
	CVB_LAUNCHED();
	{ KernelScope ks_("kht_offsets", stream);
	  kht_offsets2_kernel<<<1, 256, 0, stream>>>(dFrames, dMeta, static_cast<int>(batch), h->voteCapEl); }
	CVB_LAUNCHED();
	{ KernelScope ks_("kht_peaks_emit", stream);
	  kht_peaks_emit_kernel<<<dim3(g.nTheta, B), 256, 0, stream>>>(h->acc.as<int>(), h->rowCount.as<unsigned int>(), h->votes.as<KhtVote>(), dFrames, dMeta, g); }
	CVB_LAUNCHED();
	{
		static std::atomic<unsigned int> attrSet{0};
		// shared memory per CTA follows the largest frame of the previous call on this object (5.7 k cells for the 1080p bench frames: 3 CTAs per SM instead of the 2 that
		// the full 96 KB allow); a frame with more cells than that sorts in global memory
		const int smemItems = std::min(KSORT_SMEM_ITEMS, std::max(1024, h->sortItemsHint));
		const int smem = smemItems * (8 + 2 + 2);
		CVB_CHECK(set_max_smem_once(reinterpret_cast<const void*>(kht_peaks_sort_kernel), KSORT_SMEM_ITEMS * (8 + 2 + 2), attrSet));
		const unsigned int lim = (h->maxLines <= 0) ? static_cast<unsigned int>(INT_MAX) : static_cast<unsigned int>(h->maxLines);
		KernelScope ks_("kht_peaks_sort", stream);
		kht_peaks_sort_kernel<<<B, KSORT_THREADS, smem, stream>>>(h->votes.as<KhtVote>(), h->sortItems.as<sse_item>(), h->sortLists.as<unsigned int>(), h->sortRanges.as<int>(), h->acc.as<int>(),
			dRhoTab, dThetaTab, h->dLines.as<cvb200_hough_line_t>(), static_cast<unsigned long long>(capacity), h->dCounts.as<unsigned long long>(), dFrames, dMeta, g, lim, smemItems);
	}
	CVB_LAUNCHED();
	// results that the host always needs: the per-frame line counts + the overflow flag + the last frame's record (Gs)
	unsigned char* hMeta = h->hFrames.as<unsigned char>();
	CVB_CUDA(cudaMemcpyAsync(hMeta, dMeta, sizeof(KhtMeta), cudaMemcpyDeviceToHost, stream));
	CVB_CUDA(cudaMemcpyAsync(hMeta + sizeof(KhtMeta), dFrames + (batch - 1), sizeof(KhtFrame), cudaMemcpyDeviceToHost, stream));
	CVB_CUDA(cudaMemcpyAsync(h->hCounts.p, h->dCounts.p, batch * 8, cudaMemcpyDeviceToHost, stream));
	h->pendBatch = batch; h->pendCapacity = capacity; h->pendStream = stream;
	return CVB200_S_OK;
}

// For a producer that can write the linking bitmap itself (pipeline.cu: canny_finalize): sizes and zeroes the bitmap and the edge counters of the NEXT kht_enqueue on
// this object and hands out where they are.  Bitmap layout: (H + 2*KHT_PADR) rows of WW = ceil(W/32) + 2 words per frame, image row y at row y + KHT_PADR, pixel x in
// word 1 + x/32 at bit 31 - x%32 (kht_walk.cuh).
int cvb::kht_prepare_bits(cvb200_hough* h, size_t width, size_t height, size_t batch, cudaStream_t stream, unsigned int** bits, unsigned int** edgeCount, int* wordsPerRow)
{
	const size_t WW = div_up(width, 32) + 2;
	const size_t bitWords = (height + 2 * KHT_PADR) * WW;
	const size_t countBytes = ((batch * 4 + 15) & ~static_cast<size_t>(15)) + sizeof(KhtMeta);
	CVB_CHECK(h->bits.ensure(batch * bitWords * 4));
	CVB_CHECK(h->edgeCount.ensure(countBytes));
	CVB_CUDA(cudaMemsetAsync(h->bits.p, 0, batch * bitWords * 4, stream));
	CVB_CUDA(cudaMemsetAsync(h->edgeCount.p, 0, countBytes, stream));
	*bits = h->bits.as<unsigned int>(); *edgeCount = h->edgeCount.as<unsigned int>(); *wordsPerRow = static_cast<int>(WW);
	h->bitsPrepared = true;
	return CVB200_S_OK;
}

// Waits for kht_enqueue's work, hands out the lines.  Returns CVB200_E_PENDING-like internal code 1 when a pool was too small and has been grown: the caller enqueues again.
int cvb::kht_finish(cvb200_hough* h, cvb200_hough_line_t* lines, size_t capacity, size_t* counts, bool* again)
{
	*again = false;
	cudaStream_t stream = h->pendStream;
	const size_t batch = h->pendBatch;
	CVB_CUDA(cudaStreamSynchronize(stream));
	const KhtMeta* m = h->hFrames.as<KhtMeta>();
	if (m->maxVotes) h->sortItemsHint = static_cast<int>(std::min<unsigned int>(KSORT_SMEM_ITEMS, ((m->maxVotes + m->maxVotes / 16 + 255u) / 256u) * 256u));
	if (m->overflow) {
		if (m->overflow & 1u) { h->posCapEl = std::max<size_t>(h->posCapEl, static_cast<size_t>(m->needPos + m->needPos / 4 + 1024)); h->strCapEl = std::max<size_t>(h->strCapEl, static_cast<size_t>(m->needStr + m->needStr / 4 + 1024)); }
		if (m->overflow & 2u) h->voteCapEl = std::max<size_t>(h->voteCapEl, static_cast<size_t>(m->needVotes + m->needVotes / 4 + 1024));
		CVB_REQUIRE(h->posCapEl < (1ull << 31) && h->strCapEl < (1ull << 31), CVB200_E_OUT_OF_BOUND);
		*again = true;
		return CVB200_S_OK;
	}
	const unsigned long long* hc = h->hCounts.as<unsigned long long>();
	size_t maxLines = 0;
	for (size_t f = 0; f < batch; ++f) { counts[f] = static_cast<size_t>(hc[f]); maxLines = std::max<size_t>(maxLines, std::min<size_t>(counts[f], capacity)); }
	if (maxLines) { // one strided copy: the leading maxLines lines of every frame
		CVB_CUDA(cudaMemcpy2DAsync(lines, capacity * sizeof(cvb200_hough_line_t), h->dLines.p, capacity * sizeof(cvb200_hough_line_t), maxLines * sizeof(cvb200_hough_line_t), batch, cudaMemcpyDeviceToHost, stream));
		CVB_CUDA(cudaStreamSynchronize(stream));
	}
	const KhtFrame* last = reinterpret_cast<const KhtFrame*>(h->hFrames.as<unsigned char>() + sizeof(KhtMeta));
	h->lastGs = (last->nStr && last->nClus) ? last->gs : 1.0;
	return CVB200_S_OK;
}

int cvb::kht_process_dev(cvb200_hough* h, const uint8_t* edges, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch,
	cvb200_hough_line_t* lines, size_t capacity, size_t* counts, cudaStream_t stream)
{
	for (int attempt = 0; attempt < 4; ++attempt) {
		CVB_CHECK(kht_enqueue(h, edges, width, height, stride, batch, framePitch, capacity, stream));
		bool again = false;
		CVB_CHECK(kht_finish(h, lines, capacity, counts, &again));
		if (!again) return CVB200_S_OK;
	}
	return CVB200_E_OUT_OF_BOUND;
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
		DevBuf* bufs[] = { &h->bits, &h->poss, &h->strings, &h->strRev, &h->sortItems, &h->sortLists, &h->sortRanges, &h->dLines, &h->dCounts, &h->tabs, &h->clus, &h->clusOrd, &h->nClusStr, &h->stack, &h->kern, &h->acc, &h->rowCount, &h->votes, &h->frames, &h->edgeCount, &h->hostIn,
			&h->shtTables, &h->shtList, &h->shtCursor, &h->shtMask, &h->shtPool, &h->shtDesc };
		for (DevBuf* b : bufs) b->release();
		h->hFrames.release(); h->hVotes.release(); h->hCounts.release(); h->hTabs.release();
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
