// a8 -- FAST9 / FAST12 corner detector: strength map (K11), 3x3 non-maximum suppression and raster-ordered point list (K12).
// Replaces CompVCornerDeteFAST::process (core/features/fast/compv_core_feature_fast_dete.cxx:163-422), the leaves
// CompVFastDataRow_C (:658-771), CompVFastNmsGather_C/_Apply_C (:773-831) and CompVFastBuildInterestPoints (:490-585).
// The reference author already put the GPU seam at "produce the strength map" (CompVGpuCornerDeteFAST::processData,
// gpu/include/compv/gpu/core/features/fast/compv_gpu_feature_fast_dete.h:23-47, call site compiled out at fast_dete.cxx:252-283);
// cvb200_fast_scores has exactly that shape.
//
// One kernel per batch does the pixel work (HBM traffic: 1 B/px read + 1 bit/px mask + a few bytes per corner):
//   stage 0  TMA tile load (u8, 144 x (TH+8) box, zero filled outside the image)
//   stage A  compass filter on 4 px per lane: a contiguous arc of N>=9 of the 16 circle pixels always contains at least 2 (N=9) / 3 (N=12) of the 4 compass
//            pixels, so a corner has that many compass pixels with |c - centre| > t; byte-wise SIMD (VABSDIFF4 + a carry into bit 7) tests the lane's four pixels at once.
//            Survivors are marked in a per-lane bit mask (8 rows x 4 px) and compacted into a queue private to the warp once, after the warp's rows
//   stage B  full segment test on the queue (one candidate per lane): 16-bit darker/brighter masks, run-of-N test by shift-and,
//            strength = max over qualifying arcs of the minimum |difference| in the arc; written to a shared strength tile (exact whatever stage A let through)
//   stage C  NMS (suppressed when any 8-neighbour is >= own strength, ties kill both) and emission: a bit in the per-frame corner
//            mask (atomicOr) + (key,strength) appended to an unordered list
// then   fast_rank_*: exclusive scan of the popcounts of the mask words (chunk sums, scan of the chunk sums, chunk-local scans)
//        fast_emit_points: every list entry finds its raster rank = prefix[word] + popc(lower bits) and writes its point there.
#include "common.cuh"
#include "tma.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace cvb {

constexpr int FA_TW = 120, FA_TH = 56, FA_THREADS = 256, FA_WARPS = 8;
constexpr int FA_INW = 36;                 // words per staged row (144-byte TMA box, see canny_fast.cuh)
constexpr int FA_IN_ROWS = FA_TH + 8;      // image rows y0-4 .. y0+TH+3
constexpr int FA_S_ROWS = FA_TH + 2;       // strength rows y0-1 .. y0+TH
constexpr int FA_S_PITCH = 128;            // strength tile: one byte per staged column
constexpr int FA_QWARP = 8 * 128;          // candidate queue of one warp: every pixel of its (at most 8) rows may be a candidate
constexpr int FA_QCAP = FA_QWARP * FA_WARPS;

struct FastKParams {
	const uint8_t* in;
	uint8_t* scores;            // optional dense strength map (cvb200_fast_scores); pre-NMS values, 0 elsewhere
	unsigned int* mask;         // [batch][maskWordsPerFrame] corner bits (nullptr when only scores are wanted)
	unsigned long long* list;   // [batch][listCap] (key << 8) | strength
	unsigned int* listCount;    // [batch]
	int W, H;
	size_t stride, framePitch;
	int threshold, N, nms;
	int useTma;
	unsigned int maskWordsPerFrame, listCap;
};

// circle offsets (dx, dy) in the reference's order (fast_dete.cxx:221-238); compile-time so that an unrolled loop turns them into immediate load offsets
__host__ __device__ constexpr int circle_dx(int k) { return (k <= 3) ? k : (k <= 5) ? 3 : (k <= 8) ? (8 - k) : (k <= 11) ? (8 - k) : (k <= 13) ? -3 : (k - 16); }
__host__ __device__ constexpr int circle_dy(int k) { return circle_dx((k + 12) & 15); }
static_assert(circle_dx(0) == 0 && circle_dx(1) == 1 && circle_dx(2) == 2 && circle_dx(3) == 3 && circle_dx(4) == 3 && circle_dx(5) == 3 && circle_dx(6) == 2 && circle_dx(7) == 1
	&& circle_dx(8) == 0 && circle_dx(9) == -1 && circle_dx(10) == -2 && circle_dx(11) == -3 && circle_dx(12) == -3 && circle_dx(13) == -3 && circle_dx(14) == -2 && circle_dx(15) == -1, "circle dx");
static_assert(circle_dy(0) == -3 && circle_dy(1) == -3 && circle_dy(2) == -2 && circle_dy(3) == -1 && circle_dy(4) == 0 && circle_dy(5) == 1 && circle_dy(6) == 2 && circle_dy(7) == 3
	&& circle_dy(8) == 3 && circle_dy(9) == 3 && circle_dy(10) == 2 && circle_dy(11) == 1 && circle_dy(12) == 0 && circle_dy(13) == -1 && circle_dy(14) == -2 && circle_dy(15) == -3, "circle dy");

// bit i of the result is set when bits i..i+n-1 (circularly, 16 positions) of m are all set
__device__ __forceinline__ unsigned int arc_starts(unsigned int m, int n)
{
	const unsigned int x = m | (m << 16);
	unsigned int a = x & (x >> 1);   // runs >= 2
	unsigned int b = a & (a >> 2);   // runs >= 4
	unsigned int c = b & (b >> 4);   // runs >= 8
	unsigned int d = (n == 9) ? (c & (x >> 8)) : (c & (b >> 8)); // 8+1, or 8 followed by 4 starting at +8 = 12
	return d & 0xffffu;
}

__global__ void __launch_bounds__(FA_THREADS, 5)
fast_detect_kernel(const __grid_constant__ CUtensorMap tmap, const FastKParams p)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const unsigned int pad = (128u - (static_cast<unsigned int>(__cvta_generic_to_shared(smem_raw)) & 127u)) & 127u;
	unsigned int* sIn = reinterpret_cast<unsigned int*>(smem_raw + pad) + 32;           // FA_IN_ROWS x FA_INW words (128-byte aligned, 32 pad words in front)
	uint8_t* sS = reinterpret_cast<uint8_t*>(sIn + FA_IN_ROWS * FA_INW + 4);            // FA_S_ROWS x 128 bytes, +1 row of slack on each side
	unsigned short* sQ = reinterpret_cast<unsigned short*>(sS + (FA_S_ROWS + 2) * FA_S_PITCH); // candidate queue
	uint64_t* bar = reinterpret_cast<uint64_t*>(sQ + ((FA_QCAP + 3) & ~3));

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int W = p.W, H = p.H;
	const int x0 = blockIdx.x * FA_TW, y0 = blockIdx.y * FA_TH;
	const int frame = blockIdx.z;
	const int xl = x0 - 4 + 4 * lane;
	const int yIn0 = y0 - 4;
	const int xTma = (x0 - 4) & ~15;
	const int woff = ((x0 - 4) - xTma) >> 2;
	const int t = p.threshold;

	if (p.useTma) {
		if (threadIdx.x == 0) {
			mbar_init(bar, 1);
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			mbar_expect_tx(bar, FA_IN_ROWS * FA_INW * 4);
			tma_load_3d(sIn, &tmap, bar, xTma, yIn0, frame);
		}
	}
	else {
		const uint8_t* __restrict__ in = p.in + frame * p.framePitch;
		for (int r = warp; r < FA_IN_ROWS; r += FA_WARPS) {
			const int y = yIn0 + r;
			unsigned int w = 0;
			if (y >= 0 && y < H) {
				const uint8_t* row = in + static_cast<size_t>(y) * p.stride;
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const int x = xl + i;
					if (x >= 0 && x < W) w |= static_cast<unsigned int>(row[x]) << (8 * i);
				}
			}
			sIn[r * FA_INW + woff + lane] = w;
		}
	}
	// zero the strength tile (with its slack rows) while the copy is in flight
	for (int i = threadIdx.x; i < (FA_S_ROWS + 2) * FA_S_PITCH / 4; i += FA_THREADS) reinterpret_cast<unsigned int*>(sS)[i] = 0;
	__syncthreads();
	if (p.useTma) mbar_wait(bar, 0);
	uint8_t* sStr = sS + FA_S_PITCH; // row 0 of the strength region (image y0-1); one slack row above and below

	// ---- stage A: compass filter, 4 px per lane, rows of the strength region; stage B: full segment test on the survivors ----
	// A corner has at least `need` of its 4 compass pixels darker than pc - t, or as many brighter than pc + t (a contiguous arc of N of the 16 circle pixels always
	// holds 2 (N = 9) / 3 (N = 12) of them).  The filter keeps the weaker condition "|c - pc| > t for at least `need` compass pixels", which byte-wise SIMD evaluates
	// for the lane's four pixels at once (one absolute-difference instruction per compass word, the comparison with t through the carry into bit 7 of each byte);
	// the segment test of stage B is exact whatever the filter lets through.  Survivors go to a queue private to the warp (its rows interleave with the other
	// warps', so the queues balance).
	const bool need3 = (p.N == 12), tBig = (t >= 128);
	const unsigned int K = static_cast<unsigned int>((tBig ? 0xff : 0x7f) - t) * 0x01010101u;
	unsigned int xOk = 0; // which of the lane's 4 columns may hold a corner of this tile's strength region
#pragma unroll
	for (int i = 0; i < 4; ++i) { const int x = xl + i; if (x >= 3 && x < W - 3 && x >= x0 - 1 && x <= x0 + FA_TW) xOk |= 1u << i; }
	unsigned short* myQ = sQ + warp * FA_QWARP;
	const uint8_t* sInB = reinterpret_cast<const uint8_t*>(sIn) + woff * 4; // byte (r, c): column c <-> image x0-4+c
	// A warp owns rows warp, warp + 8, ... of the strength region: at most 8 rows x 4 px per lane = one 32-bit candidate mask per lane, no queue traffic inside the row loop.
	static_assert((FA_S_ROWS + FA_WARPS - 1) / FA_WARPS <= 8, "candidate mask: 8 rows x 4 px per lane");
	unsigned int candAll = 0;
	{
		const unsigned int* q = &sIn[(warp + 3) * FA_INW + woff + lane]; // staged row of image y (yIn0 = y0-4 -> row index y - yIn0 = rs + 3)
		int sh = 0;
		for (int rs = warp; rs < FA_S_ROWS; rs += FA_WARPS, q += FA_WARPS * FA_INW, sh += 4) {
			const int y = y0 - 1 + rs;
			if (y < 3 || y >= H - 3) continue;
			const unsigned int wc = q[0], wl = q[-1], wr = q[1];
			const unsigned int wu = q[-3 * FA_INW], wd = q[3 * FA_INW];
			const unsigned int wL = __byte_perm(wl, wc, 0x4321);   // bytes x-3..x   -> [wl.b1 wl.b2 wl.b3 wc.b0]
			const unsigned int wR = __byte_perm(wc, wr, 0x6543);   // bytes x+3..x+6 -> [wc.b3 wr.b0 wr.b1 wr.b2]
			const unsigned int dU = __vabsdiffu4(wu, wc), dD = __vabsdiffu4(wd, wc), dL = __vabsdiffu4(wL, wc), dR = __vabsdiffu4(wR, wc);
			// bit 7 of every byte: d > t.  (d & 0x7f) + K carries into bit 7 exactly when the low seven bits exceed t (t < 128: OR with d's own bit 7) or t - 128 (AND with it)
			const unsigned int sU = (dU & 0x7f7f7f7fu) + K, sD = (dD & 0x7f7f7f7fu) + K, sL = (dL & 0x7f7f7f7fu) + K, sR = (dR & 0x7f7f7f7fu) + K;
			const unsigned int h0 = tBig ? (dU & sU) : (dU | sU), h1 = tBig ? (dD & sD) : (dD | sD), h2 = tBig ? (dL & sL) : (dL | sL), h3 = tBig ? (dR & sR) : (dR | sR);
			const unsigned int m3 = (h0 & h1) | (h0 & h2) | (h1 & h2);                                 // at least two of the first three
			const unsigned int hit = need3 ? ((h0 & h1 & h2) | (m3 & h3)) : (m3 | (h3 & (h0 | h1 | h2)));
			const unsigned int cand = ((((hit >> 7) & 0x01010101u) * 0x01020408u) >> 24) & xOk;       // bits 7/15/23/31 -> bits 0..3
			candAll |= cand << sh;
		}
	}
	// queue positions: exclusive prefix of the lanes' candidate counts
	const int myCnt = __popc(candAll);
	int incl = myCnt;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
	const unsigned int qn = static_cast<unsigned int>(__shfl_sync(0xffffffffu, incl, 31));
	{
		unsigned int pos = static_cast<unsigned int>(incl - myCnt);
		while (candAll) {
			const int bidx = __ffs(candAll) - 1;
			candAll &= candAll - 1;
			myQ[pos++] = static_cast<unsigned short>((warp + FA_WARPS * (bidx >> 2)) * FA_S_PITCH + lane * 4 + (bidx & 3));
		}
	}
	__syncwarp();

	// ---- stage B: full segment test on the warp's own candidates ----
	for (unsigned int qi = lane; qi < qn; qi += 32) {
		const int code = myQ[qi];
		const int rq = code >> 7, c = code & 127;
		const uint8_t* ctr = sInB + (rq + 3) * (FA_INW * 4) + c;
		const int pc = ctr[0];
		const int br = min(pc + t, 255), dk = max(pc - t, 0);
		unsigned long long vlo = 0, vhi = 0; // the 16 circle bytes, k = 0..7 in vlo, 8..15 in vhi (registers, no local-memory array)
		unsigned int md = 0, mb = 0;
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			const int vk = ctr[circle_dy(k) * (FA_INW * 4) + circle_dx(k)];
			if (k < 8) vlo |= static_cast<unsigned long long>(vk) << (8 * k); else vhi |= static_cast<unsigned long long>(vk) << (8 * (k - 8));
			md |= (vk < dk ? 1u : 0u) << k;
			mb |= (vk > br ? 1u : 0u) << k;
		}
		int strength = 0;
		// fast_dete.cxx:733-764: darker arcs when >= N pixels are darker, else brighter arcs when >= N are brighter
		unsigned int starts = 0;
		bool dark = false;
		if (__popc(md) >= p.N) { starts = arc_starts(md, p.N); dark = true; }
		else if (__popc(mb) >= p.N) { starts = arc_starts(mb, p.N); }
		while (starts) {
			const int s = __ffs(starts) - 1;
			starts &= starts - 1;
			int mn = 255;
			for (int k = 0; k < p.N; ++k) {
				const int idx = (s + k) & 15;
				const int val = static_cast<int>(((idx < 8 ? vlo : vhi) >> (8 * (idx & 7))) & 0xff);
				const int dif = dark ? (dk - val) : (val - br);
				mn = min(mn, dif);
			}
			strength = max(strength, mn);
		}
		if (strength) sStr[rq * FA_S_PITCH + c] = static_cast<uint8_t>(strength);
	}
	__syncthreads();

	// ---- stage C: NMS + emission, output rows y0 .. y0+TH-1 (strength rows 1..TH), lanes 1..30 ----
	const bool laneOut = (lane >= 1 && lane <= 30);
	for (int ro = warp; ro < FA_TH; ro += FA_WARPS) {
		const int y = y0 + ro;
		if (y >= H) break;
		const int rs = ro + 1;
		const unsigned int sw = *reinterpret_cast<const unsigned int*>(&sStr[rs * FA_S_PITCH + lane * 4]);
		if (p.scores && laneOut && xl < W) {
			uint8_t* o = p.scores + frame * p.framePitch + static_cast<size_t>(y) * p.stride + xl;
#pragma unroll
			for (int i = 0; i < 4; ++i) if (xl + i < W) o[i] = static_cast<uint8_t>(sw >> (8 * i));
		}
		if (!p.mask || !laneOut || !sw) continue;
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int s = (sw >> (8 * i)) & 0xff;
			if (!s) continue;
			const int x = xl + i;
			if (p.nms) {
				const uint8_t* q = &sStr[rs * FA_S_PITCH + lane * 4 + i];
				// fast_dete.cxx:773-812: suppressed when ANY of the 8 neighbours is >= own strength
				if (q[-1] >= s || q[1] >= s || q[-FA_S_PITCH - 1] >= s || q[-FA_S_PITCH] >= s || q[-FA_S_PITCH + 1] >= s
					|| q[FA_S_PITCH - 1] >= s || q[FA_S_PITCH] >= s || q[FA_S_PITCH + 1] >= s) continue;
			}
			const unsigned int key = static_cast<unsigned int>(y) * static_cast<unsigned int>(W) + static_cast<unsigned int>(x);
			atomicOr(&p.mask[frame * static_cast<size_t>(p.maskWordsPerFrame) + (key >> 5)], 1u << (key & 31));
			const unsigned int slot = atomicAdd(&p.listCount[frame], 1u);
			if (slot < p.listCap) p.list[frame * static_cast<size_t>(p.listCap) + slot] = (static_cast<unsigned long long>(key) << 8) | static_cast<unsigned long long>(s);
		}
	}
}

// Exclusive prefix of popc(mask word) per frame (the raster rank of every corner).  Three small passes over 4096-word chunks -- chunk sums, a scan of the chunk sums per
// frame, chunk-local scans + offset -- instead of one 1024-thread block walking a whole frame (round 1: 41 % of the FAST time at 3840x2160).
constexpr int FR_WORDS_PER_THREAD = 16, FR_THREADS = 256, FR_CHUNK = FR_WORDS_PER_THREAD * FR_THREADS;

__device__ __forceinline__ unsigned int fr_block_exclusive(unsigned int v, unsigned int* sWarp, unsigned int* total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned int inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
	if (lane == 31) sWarp[warp] = inc;
	__syncthreads();
	if (threadIdx.x == 0) { unsigned int run = 0; for (int w = 0; w < FR_THREADS / 32; ++w) { const unsigned int t = sWarp[w]; sWarp[w] = run; run += t; } sWarp[FR_THREADS / 32] = run; }
	__syncthreads();
	*total = sWarp[FR_THREADS / 32];
	return sWarp[warp] + inc - v;
}

__global__ void __launch_bounds__(FR_THREADS) fast_rank_chunksum_kernel(const unsigned int* __restrict__ mask, unsigned int* __restrict__ chunkSums, unsigned int wordsPerFrame, unsigned int chunksPerFrame)
{
	__shared__ unsigned int sWarp[FR_THREADS / 32 + 1];
	const unsigned int* m = mask + blockIdx.y * static_cast<size_t>(wordsPerFrame);
	const unsigned int i0 = (blockIdx.x * FR_THREADS + threadIdx.x) * FR_WORDS_PER_THREAD;
	unsigned int c = 0;
#pragma unroll
	for (int k = 0; k < FR_WORDS_PER_THREAD; ++k) if (i0 + k < wordsPerFrame) c += __popc(m[i0 + k]);
	unsigned int total;
	fr_block_exclusive(c, sWarp, &total);
	if (threadIdx.x == 0) chunkSums[blockIdx.y * chunksPerFrame + blockIdx.x] = total;
}

__global__ void __launch_bounds__(FR_THREADS) fast_rank_chunkscan_kernel(unsigned int* __restrict__ chunkSums, unsigned int chunksPerFrame)
{
	__shared__ unsigned int sWarp[FR_THREADS / 32 + 1];
	unsigned int* cs = chunkSums + blockIdx.x * chunksPerFrame;
	unsigned int carry = 0;
	for (unsigned int base = 0; base < chunksPerFrame; base += FR_THREADS) {
		const unsigned int i = base + threadIdx.x;
		const unsigned int v = (i < chunksPerFrame) ? cs[i] : 0u;
		unsigned int total;
		const unsigned int ex = fr_block_exclusive(v, sWarp, &total);
		if (i < chunksPerFrame) cs[i] = carry + ex;
		carry += total;
		__syncthreads();
	}
}

__global__ void __launch_bounds__(FR_THREADS) fast_rank_apply_kernel(const unsigned int* __restrict__ mask, const unsigned int* __restrict__ chunkOffsets, unsigned int* __restrict__ prefix,
	unsigned int wordsPerFrame, unsigned int chunksPerFrame)
{
	__shared__ unsigned int sWarp[FR_THREADS / 32 + 1];
	const unsigned int* m = mask + blockIdx.y * static_cast<size_t>(wordsPerFrame);
	unsigned int* out = prefix + blockIdx.y * static_cast<size_t>(wordsPerFrame);
	const unsigned int i0 = (blockIdx.x * FR_THREADS + threadIdx.x) * FR_WORDS_PER_THREAD;
	unsigned int pc[FR_WORDS_PER_THREAD];
	unsigned int c = 0;
#pragma unroll
	for (int k = 0; k < FR_WORDS_PER_THREAD; ++k) { pc[k] = (i0 + k < wordsPerFrame) ? __popc(m[i0 + k]) : 0u; c += pc[k]; }
	unsigned int total;
	unsigned int run = chunkOffsets[blockIdx.y * chunksPerFrame + blockIdx.x] + fr_block_exclusive(c, sWarp, &total);
#pragma unroll
	for (int k = 0; k < FR_WORDS_PER_THREAD; ++k) { if (i0 + k < wordsPerFrame) out[i0 + k] = run; run += pc[k]; }
}

// CompVInterestPoint layout (base/include/compv/base/compv_common.h:629-656)
__global__ void fast_emit_points_kernel(const unsigned long long* list, const unsigned int* listCount, unsigned int listCap, const unsigned int* mask,
	const unsigned int* prefix, unsigned int wordsPerFrame, int W, int thresholdMinus1, cvb200_interest_point_t* points, unsigned int capacity, unsigned int* counts)
{
	const int frame = blockIdx.y;
	const unsigned int n = min(listCount[frame], listCap);
	if (blockIdx.x == 0 && threadIdx.x == 0) counts[frame] = listCount[frame];
	for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const unsigned long long e = list[frame * static_cast<size_t>(listCap) + i];
		const unsigned int key = static_cast<unsigned int>(e >> 8), s = static_cast<unsigned int>(e & 0xff);
		const unsigned int word = key >> 5, bit = key & 31;
		const unsigned int rank = prefix[frame * static_cast<size_t>(wordsPerFrame) + word] + __popc(mask[frame * static_cast<size_t>(wordsPerFrame) + word] & ((1u << bit) - 1u));
		if (rank >= capacity) continue;
		cvb200_interest_point_t pt;
		pt.x = static_cast<float>(key % static_cast<unsigned int>(W));
		pt.y = static_cast<float>(key / static_cast<unsigned int>(W));
		pt.strength = static_cast<float>(s + thresholdMinus1); // fast_dete.cxx:515-519: strength + (threshold - 1)
		pt.orient = -1.f; pt.level = 0; pt.size = 0.f;           // CompVInterestPoint ctor defaults
		points[frame * static_cast<size_t>(capacity) + rank] = pt;
	}
}

constexpr size_t FA_SMEM = 128 + 128 + (FA_IN_ROWS * FA_INW + 4) * 4 + (FA_S_ROWS + 2) * FA_S_PITCH + ((FA_QCAP + 3) & ~3) * 2 + 16;

} // namespace cvb

using namespace cvb;

struct cvb200_corner_dete {
	int id;
	int threshold;       // COMPV_FEATURE_DETE_FAST_THRESHOLD_DEFAULT 20
	int type;            // FAST_TYPE_9
	int N;
	int maxFeatures;     // 2000
	bool nms;            // true
	// ORB detector (id CVB200_ORB_ID): its own quota + the internal FAST object (which keeps FAST's defaults, maxFeatures 2000 included)
	int orbMaxFeatures = 2000;
	cvb200_corner_dete* orbFast = nullptr;
	DevBuf orbLevel, orbPts, orbMom;
	DevBuf mask, prefix, list, counters, hostIn, points, rankSums;
	HostBuf hostCount;
	std::mutex mutex;
};

static int fast_launch(cvb200_corner_dete* d, const uint8_t* image, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch,
	uint8_t* scores, cvb200_interest_point_t* points, size_t capacity, unsigned int* counts, int threshold, int N, bool nms, cudaStream_t stream)
{
	CVB_REQUIRE(width >= 4 && height >= 4, CVB200_E_INVALID_PARAMETER); // fast_dete.cxx:181
	CVB_REQUIRE(width * height < (1ull << 32) && width <= 0x3fffffff && height <= 0x3fffffff, CVB200_E_OUT_OF_BOUND);
	FastKParams p;
	memset(&p, 0, sizeof(p));
	p.in = image; p.scores = scores;
	p.W = static_cast<int>(width); p.H = static_cast<int>(height);
	p.stride = stride; p.framePitch = framePitch;
	p.threshold = threshold; p.N = N; p.nms = nms ? 1 : 0;
	const unsigned int words = static_cast<unsigned int>(div_up(width * height, 32));
	p.maskWordsPerFrame = words;
	if (points) {
		size_t cap = (width * height) / 4 + 1024;     // entries per frame; NMS'd corners are never 8-adjacent (<= W*H/4), without NMS every pixel may score
		if (!nms) cap = width * height;
		CVB_REQUIRE(cap < (1ull << 32), CVB200_E_OUT_OF_BOUND);
		p.listCap = static_cast<unsigned int>(cap);
		CVB_CHECK(d->mask.ensure(batch * words * sizeof(unsigned int)));
		CVB_CHECK(d->prefix.ensure(batch * words * sizeof(unsigned int)));
		CVB_CHECK(d->list.ensure(batch * cap * sizeof(unsigned long long)));
		CVB_CHECK(d->counters.ensure(batch * sizeof(unsigned int)));
		p.mask = d->mask.as<unsigned int>();
		p.list = d->list.as<unsigned long long>();
		p.listCount = d->counters.as<unsigned int>();
		CVB_CUDA(cudaMemsetAsync(p.mask, 0, batch * words * sizeof(unsigned int), stream));
		CVB_CUDA(cudaMemsetAsync(p.listCount, 0, batch * sizeof(unsigned int), stream));
	}
	alignas(64) CUtensorMap map;
	memset(&map, 0, sizeof(map));
	p.useTma = make_u8_tile_map(&map, image, width, height, stride, framePitch, batch, FA_INW * 4, FA_IN_ROWS) ? 1 : 0;
	static std::atomic<unsigned int> attrSet{0};
	CVB_CHECK(set_max_smem_once(reinterpret_cast<const void*>(fast_detect_kernel), static_cast<int>(FA_SMEM), attrSet));
	dim3 grid(static_cast<unsigned>(div_up(width, FA_TW)), static_cast<unsigned>(div_up(height, FA_TH)), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	{
		KernelScope ks_("fast_detect", stream);
		fast_detect_kernel<<<grid, FA_THREADS, FA_SMEM, stream>>>(map, p);
	}
	CVB_LAUNCHED();
	if (points) {
		{
			const unsigned int chunks = static_cast<unsigned int>(div_up(words, FR_CHUNK));
			CVB_CHECK(d->rankSums.ensure(batch * chunks * sizeof(unsigned int)));
			KernelScope ks_("fast_rank_prefix", stream);
			fast_rank_chunksum_kernel<<<dim3(chunks, static_cast<unsigned>(batch)), FR_THREADS, 0, stream>>>(p.mask, d->rankSums.as<unsigned int>(), words, chunks);
			fast_rank_chunkscan_kernel<<<static_cast<unsigned>(batch), FR_THREADS, 0, stream>>>(d->rankSums.as<unsigned int>(), chunks);
			fast_rank_apply_kernel<<<dim3(chunks, static_cast<unsigned>(batch)), FR_THREADS, 0, stream>>>(p.mask, d->rankSums.as<unsigned int>(), d->prefix.as<unsigned int>(), words, chunks);
			g_launches.fetch_add(2, std::memory_order_relaxed);
		}
		CVB_LAUNCHED();
		{
			KernelScope ks_("fast_emit_points", stream);
			fast_emit_points_kernel<<<dim3(32, static_cast<unsigned>(batch)), 256, 0, stream>>>(p.list, p.listCount, p.listCap, p.mask,
				d->prefix.as<unsigned int>(), words, p.W, threshold - 1, points, static_cast<unsigned int>(capacity), counts);
		}
		CVB_LAUNCHED();
	}
	return CVB200_S_OK;
}

// ---- the ORB detector (core/features/orb/compv_core_feature_orb_dete.cxx:148-358) on top of the FAST kernels ----
// Pyramid level: CompVImage::scale from the ORIGINAL image with the reference's 8-bit fixed-point bilinear kernel (base/image/compv_image_scale_bilinear.cxx:50-86, factors :163-176).
__global__ void orb_scale_bilinear_kernel(const uint8_t* __restrict__ in, int inStride, uint8_t* __restrict__ out, int outW, int outH, int outStride, unsigned int sfx, unsigned int sfy)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
	if (i >= outW) return;
	const unsigned int oy = static_cast<unsigned int>(j) * sfy, x = static_cast<unsigned int>(i) * sfx;
	const uint8_t* p = in + static_cast<size_t>(oy >> 8) * inStride + (x >> 8);
	const unsigned int y0 = oy & 255u, y1 = 255u - y0, x0 = x & 255u, x1 = 255u - x0;
	const unsigned int n0 = p[0], n1 = p[1], n2 = p[inStride], n3 = p[inStride + 1];
	out[static_cast<size_t>(j) * outStride + i] = static_cast<uint8_t>(((y1 * ((n0 * x1) + (n1 * x0))) >> 16) + ((y0 * ((n2 * x1) + (n3 * x0))) >> 16));
}

// Intensity-centroid moments of the circular patch (base/compv_patch.cxx:96-140; abscissas :199-203): m10 = sum i*I, m01 = sum j*I over the disc.  One warp per point,
// one lane per patch row; integer sums, so the order is free.
struct OrbAbscissas { short dx[64]; };
__global__ void orb_moments_kernel(const uint8_t* __restrict__ img, int stride, const int2* __restrict__ centres, int n, int radius, OrbAbscissas ab, int2* __restrict__ moments)
{
	const int pt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (pt >= n) return;
	const int2 c = centres[pt];
	int m10 = 0, m01 = 0;
	for (int j = -radius + lane; j <= radius; j += 32) {
		const int dX = ab.dx[j < 0 ? -j : j];
		const uint8_t* row = img + static_cast<size_t>(c.y + j) * stride + c.x;
		int s = 0, si = 0;
		for (int i = -dX; i <= dX; ++i) { const int v = row[i]; s += v; si += i * v; }
		m10 += si; m01 += j * s;
	}
	for (int o = 16; o; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
	if (lane == 0) moments[pt] = make_int2(m10, m01);
}

// CompVInterestPoint::selectBest (compv_common.h:641-655): the same libstdc++ nth_element + partition on the same list, so ties at the cut resolve as in the reference
static void select_best(std::vector<cvb200_interest_point_t>& v, size_t max)
{
	if (max > 1) {
		std::nth_element(v.begin(), v.begin() + max, v.end(), [](const cvb200_interest_point_t& i, const cvb200_interest_point_t& j) { return i.strength > j.strength; });
		const float pivot = v.at(max - 1).strength;
		v.resize(std::partition(v.begin() + max, v.end(), [pivot](cvb200_interest_point_t i) { return i.strength >= pivot; }) - v.begin());
	}
}

// FAST on a device image, points to the host with FAST's own maxFeatures applied (what CompVCornerDeteFAST::process returns, fast_dete.cxx:163-422)
static int fast_points_from_dev(cvb200_corner_dete* d, const uint8_t* dImage, size_t width, size_t height, size_t stride, std::vector<cvb200_interest_point_t>& v)
{
	const size_t devCap = d->nms ? (width * height) / 4 + 16 : width * height;
	CVB_CHECK(d->points.ensure(devCap * sizeof(cvb200_interest_point_t) + sizeof(unsigned int)));
	CVB_CHECK(d->hostCount.ensure(sizeof(unsigned int)));
	cvb200_interest_point_t* dPts = d->points.as<cvb200_interest_point_t>();
	unsigned int* dCount = reinterpret_cast<unsigned int*>(dPts + devCap);
	CVB_CHECK(fast_launch(d, dImage, width, height, stride, 1, stride * height, nullptr, dPts, devCap, dCount, d->threshold, d->N, d->nms, 0));
	unsigned int* hCount = d->hostCount.as<unsigned int>();
	CVB_CUDA(cudaMemcpyAsync(hCount, dCount, sizeof(unsigned int), cudaMemcpyDeviceToHost, 0));
	CVB_CUDA(cudaStreamSynchronize(0));
	const size_t found = *hCount;
	CVB_REQUIRE(found <= devCap, CVB200_E_OUT_OF_BOUND);
	v.resize(found);
	if (found) CVB_CUDA(cudaMemcpy(v.data(), dPts, found * sizeof(cvb200_interest_point_t), cudaMemcpyDeviceToHost));
	if (d->maxFeatures > 1 && found > static_cast<size_t>(d->maxFeatures)) select_best(v, static_cast<size_t>(d->maxFeatures));
	return CVB200_S_OK;
}

static int orb_process(cvb200_corner_dete* d, const uint8_t* image, size_t width, size_t height, size_t stride, cvb200_interest_point_t* points, size_t capacity, size_t* count)
{
	constexpr int kLevels = 8, kPatchDiameter = 31, kRadius = kPatchDiameter >> 1; // orb_dete.cxx:35-44
	const float kSf = 0.83f;
	float sfTab[kLevels], sfs = 1.f;                                             // compv_image_scale_pyramid.cxx:37-45
	sfTab[0] = 1.f;
	{ float s = kSf; for (int l = 1; l < kLevels; ++l, s *= kSf) { sfTab[l] = s; sfs += s; } }
	OrbAbscissas ab;
	memset(&ab, 0, sizeof(ab));
	for (int i = 0; i <= kRadius; ++i) ab.dx[i] = static_cast<short>(sqrt(static_cast<double>(kRadius * kRadius - (i * i)))); // compv_patch.cxx:199-203
	cvb200_corner_dete* f = d->orbFast;
	f->threshold = d->threshold; f->nms = d->nms; f->type = d->type; f->N = d->N;   // initDetector (orb_dete.cxx:263-273); FAST's own maxFeatures stays at its default
	const size_t n = stride * height;
	CVB_CHECK(d->hostIn.ensure(n + stride + 16));
	CVB_CUDA(cudaMemcpyAsync(d->hostIn.p, image, n, cudaMemcpyHostToDevice, 0));
	std::vector<cvb200_interest_point_t> all, pts;
	for (int level = 0; level < kLevels; ++level) {
		const float sf = sfTab[level];
		const uint8_t* dImg = d->hostIn.as<uint8_t>();
		size_t lw = width, lh = height, ls = stride;
		if (level) {
			lw = static_cast<size_t>(width * sf); lh = static_cast<size_t>(height * sf); ls = (lw + 15) & ~static_cast<size_t>(15);
			if (lw < 4 || lh < 4) continue;
			CVB_CHECK(d->orbLevel.ensure(ls * lh));
			const float fsx = static_cast<float>(width) / lw, fsy = static_cast<float>(height) / lh;
			const unsigned int sfx = static_cast<unsigned int>(static_cast<long>(fsx * 256.f)), sfy = static_cast<unsigned int>(static_cast<long>(fsy * 256.f));
			{
				KernelScope ks_("orb_scale", 0);
				orb_scale_bilinear_kernel<<<dim3(static_cast<unsigned>(div_up(lw, 256)), static_cast<unsigned>(lh)), 256>>>(d->hostIn.as<uint8_t>(), static_cast<int>(stride), d->orbLevel.as<uint8_t>(),
					static_cast<int>(lw), static_cast<int>(lh), static_cast<int>(ls), sfx, sfy);
			}
			CVB_LAUNCHED();
			dImg = d->orbLevel.as<uint8_t>();
		}
		CVB_CHECK(fast_points_from_dev(f, dImg, lw, lh, ls, pts));
		if (d->orbMaxFeatures > 0 && !pts.empty()) {                              // orb_dete.cxx:312-320
			const float nf = ((d->orbMaxFeatures / sfs) * sf);
			int32_t mf = static_cast<int32_t>(nf + 0.5);
			mf = mf < 10 ? 10 : mf;
			if (pts.size() > static_cast<size_t>(mf)) select_best(pts, static_cast<size_t>(mf));
		}
		{                                                                         // eraseTooCloseToBorder (compv_common.h:657-663), border (31 + 5) >> 1
			const float fw = static_cast<float>(lw), fh = static_cast<float>(lh), b = static_cast<float>((kPatchDiameter + 5) >> 1);
			pts.erase(std::remove_if(pts.begin(), pts.end(), [&](const cvb200_interest_point_t& q) { return (q.x < b || (q.x + b) >= fw || (q.y < b) || (q.y + b) >= fh); }), pts.end());
		}
		if (pts.empty()) continue;
		// moments on the device, on the level image that is already there
		std::vector<int2> centres(pts.size()), mom(pts.size());
		for (size_t k = 0; k < pts.size(); ++k) {
			centres[k].x = static_cast<int>(static_cast<int>(pts[k].x >= 0.0 ? (pts[k].x + 0.5) : (pts[k].x - 0.5)));
			centres[k].y = static_cast<int>(static_cast<int>(pts[k].y >= 0.0 ? (pts[k].y + 0.5) : (pts[k].y - 0.5)));
		}
		CVB_CHECK(d->orbPts.ensure(pts.size() * sizeof(int2)));
		CVB_CHECK(d->orbMom.ensure(pts.size() * sizeof(int2)));
		CVB_CUDA(cudaMemcpyAsync(d->orbPts.p, centres.data(), pts.size() * sizeof(int2), cudaMemcpyHostToDevice, 0));
		{
			KernelScope ks_("orb_moments", 0);
			orb_moments_kernel<<<static_cast<unsigned>(div_up(pts.size(), 8)), 256>>>(dImg, static_cast<int>(ls), d->orbPts.as<int2>(), static_cast<int>(pts.size()), kRadius, ab, d->orbMom.as<int2>());
		}
		CVB_LAUNCHED();
		CVB_CUDA(cudaMemcpy(mom.data(), d->orbMom.p, pts.size() * sizeof(int2), cudaMemcpyDeviceToHost));
		const float sfi = 1.f / sf, patchSize = kPatchDiameter / sf;
		for (size_t k = 0; k < pts.size(); ++k) {                                 // orb_dete.cxx:330-355; atan2f of the host's libm, as in the reference
			cvb200_interest_point_t& q = pts[k];
			q.level = level; q.size = patchSize;
			const float rad = std::atan2(static_cast<float>(mom[k].y), static_cast<float>(mom[k].x));
			q.orient = static_cast<float>(rad * (180.f / 3.1415926535897932384626433f));
			if (q.orient < 0) q.orient += 360;
			if (level != 0) { q.x *= sfi; q.y *= sfi; }
		}
		all.insert(all.end(), pts.begin(), pts.end());
	}
	*count = all.size();
	if (capacity && !all.empty()) memcpy(points, all.data(), std::min(capacity, all.size()) * sizeof(cvb200_interest_point_t));
	return (capacity && all.size() > capacity) ? CVB200_E_OUT_OF_BOUND : CVB200_S_OK;
}

extern "C" {

int cvb200_corner_dete_new(cvb200_corner_dete_t** dete, int id)
{
	CVB_REQUIRE(dete, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(id == CVB200_FAST_ID || id == CVB200_ORB_ID, CVB200_E_INVALID_PARAMETER);
	cvb200_corner_dete* d = new (std::nothrow) cvb200_corner_dete();
	CVB_REQUIRE(d, CVB200_E_OUT_OF_MEMORY);
	d->id = id; d->threshold = 20; d->type = CVB200_FAST_TYPE_9; d->N = 9; d->maxFeatures = 2000; d->nms = true; // fast_dete.cxx:74-80,111-125 (ORB: orb_dete.cxx:35-44)
	if (id == CVB200_ORB_ID) {
		d->orbFast = new (std::nothrow) cvb200_corner_dete();
		if (!d->orbFast) { delete d; return CVB200_E_OUT_OF_MEMORY; }
		d->orbFast->id = CVB200_FAST_ID; d->orbFast->threshold = 20; d->orbFast->type = CVB200_FAST_TYPE_9; d->orbFast->N = 9; d->orbFast->maxFeatures = 2000; d->orbFast->nms = true;
	}
	*dete = d;
	return CVB200_S_OK;
}

int cvb200_corner_dete_free(cvb200_corner_dete_t** dete)
{
	if (dete && *dete) {
		cvb200_corner_dete* d = *dete;
		if (d->orbFast) { cvb200_corner_dete_t* f = d->orbFast; cvb200_corner_dete_free(&f); }
		d->orbLevel.release(); d->orbPts.release(); d->orbMom.release();
		d->mask.release(); d->prefix.release(); d->list.release(); d->counters.release(); d->hostIn.release(); d->points.release(); d->hostCount.release(); d->rankSums.release();
		delete d;
		*dete = nullptr;
	}
	return CVB200_S_OK;
}

// fast_dete.cxx:128-160
int cvb200_corner_dete_set(cvb200_corner_dete_t* d, int id, const void* valuePtr, size_t valueSize)
{
	CVB_REQUIRE(d && valuePtr && valueSize, CVB200_E_INVALID_PARAMETER);
	if (d->id == CVB200_ORB_ID) { // orb_dete.cxx:58-134
		switch (id) {
		case CVB200_ORB_SET_INT_FAST_THRESHOLD: CVB_REQUIRE(valueSize == sizeof(int), CVB200_E_INVALID_PARAMETER); { const int t = *static_cast<const int*>(valuePtr); d->threshold = t < 0 ? 0 : (t > 255 ? 255 : t); } return CVB200_S_OK;
		case CVB200_ORB_SET_BOOL_FAST_NON_MAXIMA_SUPP: CVB_REQUIRE(valueSize == sizeof(bool), CVB200_E_INVALID_PARAMETER); d->nms = *static_cast<const bool*>(valuePtr); return CVB200_S_OK;
		case CVB200_ORB_SET_INT_MAX_FEATURES: CVB_REQUIRE(valueSize == sizeof(int), CVB200_E_INVALID_PARAMETER); d->orbMaxFeatures = *static_cast<const int*>(valuePtr); return CVB200_S_OK;
		case CVB200_ORB_SET_INT_INTERNAL_DETE_ID: {
			CVB_REQUIRE(valueSize == sizeof(int), CVB200_E_INVALID_PARAMETER);
			const int t = *static_cast<const int*>(valuePtr);
			CVB_REQUIRE(t == CVB200_FAST_TYPE_9 || t == CVB200_FAST_TYPE_12, CVB200_E_INVALID_PARAMETER);
			d->type = t; d->N = (t == CVB200_FAST_TYPE_12) ? 12 : 9;
			return CVB200_S_OK;
		}
		default: return CVB200_E_NOT_IMPLEMENTED; // pyramid levels / scale factor / scale type: the reference rebuilds its pyramid with scaleFactor() of LEVEL 0, i.e. 1.0 (orb_dete.cxx:101,121): not reproduced
		}
	}
	switch (id) {
	case CVB200_FAST_SET_INT_THRESHOLD: {
		CVB_REQUIRE(valueSize == sizeof(int), CVB200_E_INVALID_PARAMETER);
		const int t = *static_cast<const int*>(valuePtr);
		d->threshold = t < 0 ? 0 : (t > 255 ? 255 : t);
		return CVB200_S_OK;
	}
	case CVB200_FAST_SET_INT_MAX_FEATURES:
		CVB_REQUIRE(valueSize == sizeof(int), CVB200_E_INVALID_PARAMETER);
		d->maxFeatures = *static_cast<const int*>(valuePtr);
		return CVB200_S_OK;
	case CVB200_FAST_SET_INT_FAST_TYPE: {
		CVB_REQUIRE(valueSize == sizeof(int), CVB200_E_INVALID_PARAMETER);
		const int t = *static_cast<const int*>(valuePtr);
		CVB_REQUIRE(t == CVB200_FAST_TYPE_9 || t == CVB200_FAST_TYPE_12, CVB200_E_INVALID_PARAMETER);
		d->type = t; d->N = (t == CVB200_FAST_TYPE_12) ? 12 : 9;
		return CVB200_S_OK;
	}
	case CVB200_FAST_SET_BOOL_NON_MAXIMA_SUPP:
		CVB_REQUIRE(valueSize == sizeof(bool), CVB200_E_INVALID_PARAMETER);
		d->nms = *static_cast<const bool*>(valuePtr);
		return CVB200_S_OK;
	default:
		return CVB200_E_NOT_IMPLEMENTED;
	}
}

int cvb200_corner_dete_process_dev(cvb200_corner_dete_t* d, const uint8_t* image, size_t width, size_t height, size_t stride,
	cvb200_interest_point_t* points, size_t capacity, unsigned int* counts, size_t batch, size_t framePitch, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(d && image && points && counts && capacity && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(d->id == CVB200_FAST_ID, CVB200_E_NOT_IMPLEMENTED); // the ORB detector has the per-frame entry point only
	CVB_REQUIRE(capacity < (1ull << 32), CVB200_E_OUT_OF_BOUND);
	if (!batch) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	std::lock_guard<std::mutex> lock(d->mutex);
	return fast_launch(d, image, width, height, stride, batch, framePitch, nullptr, points, capacity, counts, d->threshold, d->N, d->nms, as_stream(stream));
}

int cvb200_corner_dete_process(cvb200_corner_dete_t* d, const uint8_t* image, size_t width, size_t height, size_t stride,
	cvb200_interest_point_t* points, size_t capacity, size_t* count)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(d && image && count && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(!capacity || points, CVB200_E_INVALID_PARAMETER);
	*count = 0;
	CVB_REQUIRE(width >= 4 && height >= 4, CVB200_E_INVALID_PARAMETER);
	std::lock_guard<std::mutex> lock(d->mutex);
	if (d->id == CVB200_ORB_ID) return orb_process(d, image, width, height, stride, points, capacity, count);
	const size_t n = stride * height;
	// device capacity: enough for every possible corner so that selectBest sees the full list like the reference does
	size_t devCap = d->nms ? (width * height) / 4 + 16 : width * height;
	CVB_CHECK(d->hostIn.ensure(n));
	CVB_CHECK(d->points.ensure(devCap * sizeof(cvb200_interest_point_t) + sizeof(unsigned int)));
	CVB_CHECK(d->hostCount.ensure(sizeof(unsigned int)));
	cvb200_interest_point_t* dPts = d->points.as<cvb200_interest_point_t>();
	unsigned int* dCount = reinterpret_cast<unsigned int*>(dPts + devCap);
	CVB_CUDA(cudaMemcpyAsync(d->hostIn.p, image, n, cudaMemcpyHostToDevice, 0));
	CVB_CHECK(fast_launch(d, d->hostIn.as<uint8_t>(), width, height, stride, 1, n, nullptr, dPts, devCap, dCount, d->threshold, d->N, d->nms, 0));
	unsigned int* hCount = d->hostCount.as<unsigned int>();
	CVB_CUDA(cudaMemcpyAsync(hCount, dCount, sizeof(unsigned int), cudaMemcpyDeviceToHost, 0));
	CVB_CUDA(cudaStreamSynchronize(0));
	size_t found = *hCount;
	CVB_REQUIRE(found <= devCap, CVB200_E_OUT_OF_BOUND);
	if (d->maxFeatures > 1 && found > static_cast<size_t>(d->maxFeatures)) {
		// CompVInterestPoint::selectBest (compv_common.h:641-655): the same libstdc++ nth_element + partition on the same raster-ordered list,
		// so ties at the cut resolve exactly as in the reference built with this toolchain
		std::vector<cvb200_interest_point_t> v(found);
		CVB_CUDA(cudaMemcpy(v.data(), dPts, found * sizeof(cvb200_interest_point_t), cudaMemcpyDeviceToHost));
		const size_t max = static_cast<size_t>(d->maxFeatures);
		std::nth_element(v.begin(), v.begin() + max, v.end(), [](const cvb200_interest_point_t& i, const cvb200_interest_point_t& j) { return i.strength > j.strength; });
		const float pivot = v.at(max - 1).strength;
		v.resize(std::partition(v.begin() + max, v.end(), [pivot](cvb200_interest_point_t i) { return i.strength >= pivot; }) - v.begin());
		*count = v.size();
		if (capacity) memcpy(points, v.data(), std::min(capacity, v.size()) * sizeof(cvb200_interest_point_t));
		return v.size() > capacity && capacity ? CVB200_E_OUT_OF_BOUND : CVB200_S_OK;
	}
	*count = found;
	if (capacity && found) CVB_CUDA(cudaMemcpy(points, dPts, std::min(capacity, found) * sizeof(cvb200_interest_point_t), cudaMemcpyDeviceToHost));
	return (capacity && found > capacity) ? CVB200_E_OUT_OF_BOUND : CVB200_S_OK;
}

// K11 strength map == CompVGpuCornerDeteFAST::processData(IP, width, height, stride, N, threshold, strengths)
int cvb200_fast_scores_dev(const uint8_t* image, size_t width, size_t height, size_t stride, int N, int threshold, uint8_t* strengths, size_t batch, size_t framePitch, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(image && strengths && image != strengths && width && height && stride >= width && (N == 9 || N == 12) && threshold >= 0 && threshold <= 255, CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	return fast_launch(nullptr, image, width, height, stride, batch, framePitch, strengths, nullptr, 0, nullptr, threshold, N, false, as_stream(stream));
}

int cvb200_fast_scores(const uint8_t* image, size_t width, size_t height, size_t stride, int N, int threshold, uint8_t* strengths)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(image && strengths && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	const size_t n = stride * height;
	DevBuf dIn, dOut;
	int rc = dIn.ensure(n);
	if (!rc) rc = dOut.ensure(n);
	if (!rc) rc = cvb200_memcpy_h2d(dIn.p, image, n, nullptr);
	if (!rc) rc = cvb200_memset(dOut.p, 0, n, nullptr);
	if (!rc) rc = cvb200_fast_scores_dev(dIn.as<uint8_t>(), width, height, stride, N, threshold, dOut.as<uint8_t>(), 1, 0, nullptr);
	if (!rc) rc = cvb200_memcpy_d2h(strengths, dOut.p, n, nullptr);
	if (!rc) rc = cvb200_stream_sync(nullptr);
	dIn.release(); dOut.release();
	return rc;
}

} // extern "C"
