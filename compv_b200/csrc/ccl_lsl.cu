// a11 -- connected component labeling (Parallel Light Speed Labeling).
// Replaces CompVConnectedComponentLabelingLSL::process (core/ccl/compv_core_ccl_lsl.cxx:579-751): step 1 :153-214, step 2.0 :341-374, step 2.1 :430-479,
// step 4 :481-505, build_LEA :545-576; result accessors of core/ccl/compv_core_ccl_lsl_result.cxx (debugFlatten :51-98, boundingBoxes :136-185).
//
// The reference keeps a relative-label image ER (2 B/px), a run-length table RLC (2 B/px) and an association table ERA, all per pixel column.
// Here a frame is reduced once to a 1 bit/px bitmap and everything else is indexed by SEGMENT (a maximal run of foreground pixels in a row):
//   lsl_bits      bytes -> bitmap + per-word prefix count of segment starts + segments per row                      HBM: 1 B/px read, 3/16 B/px written
//   lsl_rowscan   exclusive scan of the row counts -> CSR row offsets, segments per frame
//   lsl_emit      one thread per bitmap word: segment start / end columns, and step 2.0 in closed form: with S(x) = number of segment starts in
//                 columns [0, x] of the previous row, the previous-row segments an 8-connected run [s, e) touches are k0 .. k1 with
//                 k0 = S(max(s-1,0)) - fg(max(s-1,0)), k1 = S(min(e, W-1)) - 1 (the odd relative labels er0 .. er1 of :361-370 are 2k+1)
//   lsl_equiv     step 2.1, the equivalence table.  It is order dependent (each merge reads the table the previous ones wrote, and the reference's
//                 update rule is not a plain union-find, see oracle/compv_oracle_lsl.c), so ONE WARP PER FRAME replays it in raster order: lanes fetch
//                 32 segments at a time, lane 0 walks them against label rings of the previous/current row and the EQ table held in shared memory.
//                 Frames of a batch run concurrently on different SMs.
//   lsl_resolve   step 4 (EQ -> final numbering): roots ranked with a block scan, other labels chase EQ to their root (equal to the serial A[ea] = A[EQ[ea]])
//   lsl_lea       final label per segment -> the LEA {a, start, end}; lsl_flatten: label image straight from the bitmap rank (no per-run loops)
#include "ccl.cuh"

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

namespace cvb {

struct LslGeom {
	int W, H, WW;              // WW = W/32 + 1 bitmap words per row: column W always has a (zero) bit so that a run reaching the border ends inside the row
	size_t stride, framePitch;
	int eqCap;                 // EQ entries cached in shared memory by lsl_equiv
	int ringCap;               // label ring capacity = max segments per row
};
struct LslFrame { unsigned int segBase, nseg; int nea, na; };
#define LSL_STAGE 1024

__device__ __forceinline__ unsigned int lsl_pack4(unsigned int v) // bit 0 of each of 4 bytes -> 4-bit nibble
{
	return (((v & 0x01010101u) * 0x00204081u) >> 21) & 0xfu;
}

// ---- bitmap + starts prefix: one warp per row ----
template <bool kAligned4>
__global__ void __launch_bounds__(128) lsl_bits_kernel(const uint8_t* __restrict__ img, unsigned int* __restrict__ fg, unsigned short* __restrict__ spre,
	unsigned int* __restrict__ rowCnt, LslGeom g)
{
	const int lane = threadIdx.x & 31;
	const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
	const int f = blockIdx.y;
	if (j >= g.H) return;
	const uint8_t* row = img + static_cast<size_t>(f) * g.framePitch + static_cast<size_t>(j) * g.stride;
	unsigned int* fgRow = fg + (static_cast<size_t>(f) * g.H + j) * g.WW;
	unsigned short* spRow = spre + (static_cast<size_t>(f) * g.H + j) * g.WW;
	unsigned int carryBit = 0, carryCnt = 0;
	for (int k0 = 0; k0 < g.WW; k0 += 32) {
		const int k = k0 + lane;
		const int x = k * 32;
		unsigned int w = 0;
		if (k < g.WW && x < g.W) {
			if (kAligned4 && x + 32 <= g.W) {
				const unsigned int* p = reinterpret_cast<const unsigned int*>(row + x);
				#pragma unroll
				for (int i = 0; i < 8; ++i) w |= lsl_pack4(p[i]) << (4 * i);
			}
			else {
				const int n = min(32, g.W - x);
				for (int i = 0; i < n; ++i) w |= static_cast<unsigned int>(row[x + i] & 1u) << i;
			}
		}
		unsigned int left = __shfl_up_sync(0xffffffffu, w, 1) >> 31;
		if (lane == 0) left = carryBit;
		const unsigned int starts = w & ~((w << 1) | left);
		const unsigned int n = __popc(starts);
		unsigned int inc = n;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
		if (k < g.WW) { fgRow[k] = w; spRow[k] = static_cast<unsigned short>(carryCnt + inc - n); }
		carryCnt += __shfl_sync(0xffffffffu, inc, 31);
		carryBit = __shfl_sync(0xffffffffu, w, 31) >> 31;
	}
	if (lane == 0) rowCnt[static_cast<size_t>(f) * g.H + j] = carryCnt;
}

// ---- exclusive scan of the row counts: one CTA per frame ----
__global__ void __launch_bounds__(1024) lsl_rowscan_kernel(const unsigned int* __restrict__ rowCnt, unsigned int* __restrict__ rowOff, LslFrame* __restrict__ frames, LslGeom g)
{
	__shared__ unsigned int sWarp[32];
	__shared__ unsigned int sCarry;
	const int f = blockIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) sCarry = 0;
	__syncthreads();
	for (int j0 = 0; j0 < g.H; j0 += 1024) {
		const int j = j0 + threadIdx.x;
		const unsigned int n = (j < g.H) ? rowCnt[static_cast<size_t>(f) * g.H + j] : 0u;
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
		}
		__syncthreads();
		const unsigned int carry = sCarry;
		const unsigned int excl = carry + sWarp[warp] + inc - n;
		if (j < g.H) rowOff[static_cast<size_t>(f) * (g.H + 1) + j] = excl;
		__syncthreads();
		if (threadIdx.x == 1023) sCarry = excl + n;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		rowOff[static_cast<size_t>(f) * (g.H + 1) + g.H] = sCarry;
		frames[f].nseg = sCarry;
	}
}

// ---- segments + step 2.0: one thread per bitmap word ----
__device__ __forceinline__ unsigned int lsl_starts_upto(const unsigned int* __restrict__ fgRow, const unsigned short* __restrict__ spRow, int x)
{
	const int kk = x >> 5;
	const unsigned int pw = fgRow[kk];
	const unsigned int pl = kk ? (fgRow[kk - 1] >> 31) : 0u;
	const unsigned int st = pw & ~((pw << 1) | pl);
	return spRow[kk] + __popc(st & (0xffffffffu >> (31 - (x & 31))));
}

__global__ void __launch_bounds__(64) lsl_emit_kernel(const unsigned int* __restrict__ fg, const unsigned short* __restrict__ spre, const unsigned int* __restrict__ rowOff,
	const LslFrame* __restrict__ frames, short* __restrict__ segStart, short* __restrict__ segEnd, ushort2* __restrict__ ov, LslGeom g)
{
	const int k = blockIdx.x * 64 + threadIdx.x;
	const int j = blockIdx.y, f = blockIdx.z;
	if (k >= g.WW) return;
	const size_t rowIdx = static_cast<size_t>(f) * g.H + j;
	const unsigned int* fgRow = fg + rowIdx * g.WW;
	const unsigned int w = fgRow[k];
	const unsigned int left = k ? (fgRow[k - 1] >> 31) : 0u;
	const unsigned int prevMask = (w << 1) | left;
	unsigned int starts = w & ~prevMask, ends = ~w & prevMask;
	if (!(starts | ends)) return;
	const unsigned int base = frames[f].segBase;
	const unsigned int rowFirst = rowOff[static_cast<size_t>(f) * (g.H + 1) + j];
	const unsigned int sp = spre[rowIdx * g.WW + k];
	unsigned int sIdx = rowFirst + sp, eIdx = rowFirst + sp - left;
	const size_t prevIdx = (j > 0) ? rowIdx - 1 : rowIdx; // row 0 has no previous row: never dereferenced below
	const unsigned int* fgPrev = fg + prevIdx * g.WW;
	const unsigned short* spPrev = spre + prevIdx * g.WW;
	while (starts) {
		const int b = __ffs(starts) - 1; starts &= starts - 1;
		const int x = k * 32 + b;
		unsigned int k0 = 0;
		if (j > 0) {
			const int j0 = max(x - 1, 0);
			k0 = lsl_starts_upto(fgPrev, spPrev, j0) - ((fgPrev[j0 >> 5] >> (j0 & 31)) & 1u);
		}
		segStart[base + sIdx] = static_cast<short>(x);
		ov[base + sIdx].x = static_cast<unsigned short>(k0 | (sIdx == rowFirst ? 0x8000u : 0u));
		++sIdx;
	}
	while (ends) {
		const int b = __ffs(ends) - 1; ends &= ends - 1;
		const int x = k * 32 + b; // exclusive end column (<= W)
		unsigned int k1p = 0;
		if (j > 0) k1p = lsl_starts_upto(fgPrev, spPrev, min(x, g.W - 1));
		segEnd[base + eIdx] = static_cast<short>(x);
		ov[base + eIdx].y = static_cast<unsigned short>(k1p);
		++eIdx;
	}
}

// ---- step 2.1: one warp per frame, raster order ----
// The reference visits the segments one by one (:441-473).  Three kinds: NEW (touches no previous-row segment: takes the next label), SIMPLE (touches exactly
// one: copies EQ[label of that segment]) and MERGING (touches several: reads and rewrites EQ).  Only MERGING segments change what later segments read, and NEW
// ones only append to EQ, so a run of consecutive NEW / SIMPLE segments of one row gives the same result processed all at once: 32 segments are fetched, the
// warp walks the runs between "barriers" (a MERGING segment, handled by its own lane alone, or the first segment of a row, where the label rings swap).
__device__ __forceinline__ int lsl_eq_get(const int* eqS, const int* eqG, int eqCap, int ea) { return (ea < eqCap) ? eqS[ea] : eqG[ea]; }
__device__ __forceinline__ void lsl_eq_set(int* eqS, int* eqG, int eqCap, int ea, int v) { if (ea < eqCap) eqS[ea] = v; else eqG[ea] = v; }

__global__ void __launch_bounds__(32) lsl_equiv_kernel(const ushort2* __restrict__ ov, int* __restrict__ label, int* eqAll, LslFrame* frames, LslGeom g)
{
	extern __shared__ int sm[];
	int* prev = sm;
	int* cur = sm + g.ringCap;
	int* eqS = sm + 2 * g.ringCap;
	const int f = blockIdx.x, lane = threadIdx.x;
	const unsigned int base = frames[f].segBase, nseg = frames[f].nseg;
	int* eqG = eqAll + base + f; // nseg + 1 entries per frame
	const int eqCap = g.eqCap;
	const unsigned int ltMask = (1u << lane) - 1u;
	int curN = 0, nea = 0; // uniform across the warp
	// The descriptors were written by another kernel: every fetch is an L2 round trip.  They are staged 1024 at a time (32 independent loads per lane in flight)
	// so that the round trip is paid once per 1024 segments instead of once per group of 32.
	unsigned int* ovS = reinterpret_cast<unsigned int*>(sm + 2 * g.ringCap + g.eqCap); // LSL_STAGE entries
	const unsigned int* ov32 = reinterpret_cast<const unsigned int*>(ov + base);
	for (unsigned int s0 = 0; s0 < nseg; s0 += 32) {
		if ((s0 & (LSL_STAGE - 1)) == 0) {
			__syncwarp();
			#pragma unroll 8
			for (int i = 0; i < LSL_STAGE / 32; ++i) {
				const unsigned int k = s0 + i * 32 + lane;
				if (k < nseg) ovS[i * 32 + lane] = ov32[k];
			}
			__syncwarp();
		}
		const int cnt = static_cast<int>(min(32u, nseg - s0));
		const bool valid = lane < cnt;
		ushort2 mine = make_ushort2(0, 0);
		if (valid) { const unsigned int v = ovS[(s0 & (LSL_STAGE - 1)) + lane]; mine.x = static_cast<unsigned short>(v & 0xffffu); mine.y = static_cast<unsigned short>(v >> 16); }
		const int k0 = mine.x & 0x7fff, k1p = mine.y;
		const int n = k1p - k0;                                  // previous-row segments touched
		const unsigned int fMask = __ballot_sync(0xffffffffu, valid && (mine.x & 0x8000u)); // first segment of a row
		const unsigned int cMask = __ballot_sync(0xffffffffu, valid && n >= 2);             // merging segments
		int lab = 0;
		int pos = 0;
		while (pos < cnt) {
			if ((fMask >> pos) & 1u) { int* tmp = prev; prev = cur; cur = tmp; curN = 0; }
			if ((cMask >> pos) & 1u) {
				if (lane == pos) { // :449-465, alone
					int ea = prev[k0];
					int a = lsl_eq_get(eqS, eqG, eqCap, ea);
					for (int kk = k0 + 1; kk < k1p; ++kk) {
						const int eak = prev[kk];
						const int ak = lsl_eq_get(eqS, eqG, eqCap, eak);
						if (a < ak) lsl_eq_set(eqS, eqG, eqCap, eak, a);
						else { a = ak; lsl_eq_set(eqS, eqG, eqCap, ea, a); ea = eak; }
					}
					lab = a;
					cur[curN] = a;
				}
				__syncwarp();
				++curN; ++pos;
				continue;
			}
			// run of NEW / SIMPLE segments [pos, next): up to the next barrier
			const unsigned int later = (fMask | cMask) & ~((2u << pos) - 1u);
			const int next = later ? min(cnt, __ffs(later) - 1) : cnt;
			const bool mineNow = (lane >= pos) && (lane < next);
			const unsigned int newMask = __ballot_sync(0xffffffffu, mineNow && n <= 0);
			if (mineNow) {
				if (n <= 0) { // :467-469; EQ[ea] = ea is written when the label is born instead of pre-filling the table (build_EQ :531-541)
					lab = nea + __popc(newMask & ltMask) + 1;
					lsl_eq_set(eqS, eqG, eqCap, lab, lab);
				}
				else lab = lsl_eq_get(eqS, eqG, eqCap, prev[k0]); // :449-451 with an empty loop, :466
				cur[curN + lane - pos] = lab;
			}
			__syncwarp();
			nea += __popc(newMask);
			curN += next - pos;
			pos = next;
		}
		if (valid) label[base + s0 + lane] = lab;
	}
	__syncwarp();
	for (int ea = 1 + lane; ea <= nea && ea < eqCap; ea += 32) eqG[ea] = eqS[ea];
	if (lane == 0) { eqG[0] = 0; frames[f].nea = nea; }
}

// ---- step 4: one CTA per frame ----
__global__ void __launch_bounds__(1024) lsl_resolve_kernel(const int* __restrict__ eqAll, int* __restrict__ aAll, LslFrame* __restrict__ frames)
{
	__shared__ unsigned int sWarp[32];
	__shared__ unsigned int sCarry;
	const int f = blockIdx.x;
	const unsigned int base = frames[f].segBase;
	const int nea = frames[f].nea;
	const int* eq = eqAll + base + f;
	int* A = aAll + base + f;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) { sCarry = 0; A[0] = 0; }
	__syncthreads();
	// roots (EQ[ea] == ea) are numbered in increasing ea: A[root] = rank
	for (int e0 = 1; e0 <= nea; e0 += 1024) {
		const int ea = e0 + threadIdx.x;
		const unsigned int n = (ea <= nea && eq[ea] == ea) ? 1u : 0u;
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
		}
		__syncthreads();
		const unsigned int incl = sCarry + sWarp[warp] + inc;
		if (n) A[ea] = static_cast<int>(incl);
		__syncthreads();
		if (threadIdx.x == 1023) sCarry = incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) frames[f].na = static_cast<int>(sCarry);
	__syncthreads();
	// every other label takes its root's number: the serial loop's A[ea] = A[EQ[ea]] with EQ[ea] < ea unrolls to exactly this chase
	for (int ea = 1 + threadIdx.x; ea <= nea; ea += 1024) {
		int r = eq[ea];
		if (r == ea) continue;
		int rr;
		while ((rr = eq[r]) != r) r = rr;
		A[ea] = A[r];
	}
}

// ---- LEA ----
__global__ void __launch_bounds__(256) lsl_lea_kernel(const int* __restrict__ label, const int* __restrict__ aAll, const short* __restrict__ segStart, const short* __restrict__ segEnd,
	const LslFrame* __restrict__ frames, cvb200_ccl_range_t* __restrict__ ranges)
{
	const int f = blockIdx.y;
	const unsigned int base = frames[f].segBase, nseg = frames[f].nseg;
	const int* A = aAll + base + f;
	for (unsigned int s = blockIdx.x * 256 + threadIdx.x; s < nseg; s += gridDim.x * 256) {
		cvb200_ccl_range_t r;
		r.a = A[label[base + s]];
		r.start = segStart[base + s];
		r.end = segEnd[base + s];
		ranges[base + s] = r;
	}
}

// ---- flattened label image from the bitmap rank ----
__global__ void __launch_bounds__(256) lsl_flatten_kernel(const unsigned int* __restrict__ fg, const unsigned short* __restrict__ spre, const unsigned int* __restrict__ rowOff,
	const LslFrame* __restrict__ frames, const cvb200_ccl_range_t* __restrict__ ranges, int* __restrict__ labels, LslGeom g)
{
	const int x = blockIdx.x * 256 + threadIdx.x;
	const int j = blockIdx.y, f = blockIdx.z;
	if (x >= g.W) return;
	const size_t rowIdx = static_cast<size_t>(f) * g.H + j;
	const unsigned int* fgRow = fg + rowIdx * g.WW;
	int a = 0;
	if ((fgRow[x >> 5] >> (x & 31)) & 1u) {
		const unsigned int s = lsl_starts_upto(fgRow, spre + rowIdx * g.WW, x) - 1;
		a = ranges[frames[f].segBase + rowOff[static_cast<size_t>(f) * (g.H + 1) + j] + s].a;
	}
	labels[(static_cast<size_t>(f) * g.H + j) * g.W + x] = a;
}

} // namespace cvb

using namespace cvb;

static int lsl_process_dev(cvb200_ccl* c, const uint8_t* binar, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch,
	int32_t* labels, int32_t* na, cvb200_ccl_result_t** results, cudaStream_t stream)
{
	CVB_REQUIRE(width <= 32767 && height <= 32767 && batch <= 65535, CVB200_E_OUT_OF_BOUND); // int16 columns / rows as in the reference
	LslGeom g;
	memset(&g, 0, sizeof(g));
	g.W = static_cast<int>(width); g.H = static_cast<int>(height); g.WW = static_cast<int>(width / 32 + 1);
	g.stride = stride; g.framePitch = framePitch;
	g.ringCap = static_cast<int>((width + 1) / 2 + 1);
	g.eqCap = 20480; // 80 KB of EQ per frame-warp: two frames per SM; labels beyond it live in global memory
	const size_t smemMax = 200 * 1024;
	if ((static_cast<size_t>(g.ringCap) * 2 + g.eqCap + LSL_STAGE) * 4 > smemMax) g.eqCap = static_cast<int>(smemMax / 4 - static_cast<size_t>(g.ringCap) * 2 - LSL_STAGE);
	const size_t words = static_cast<size_t>(g.H) * g.WW;
	CVB_CHECK(c->fg.ensure(batch * words * 4));
	CVB_CHECK(c->spre.ensure(batch * words * 2));
	CVB_CHECK(c->rowCnt.ensure(batch * height * 4));
	CVB_CHECK(c->rowOff.ensure(batch * (height + 1) * 4));
	CVB_CHECK(c->frames.ensure(batch * sizeof(LslFrame)));
	CVB_CHECK(c->hFrames.ensure(batch * sizeof(LslFrame)));
	LslFrame* dFrames = c->frames.as<LslFrame>();
	LslFrame* hf = c->hFrames.as<LslFrame>();
	const unsigned int B = static_cast<unsigned int>(batch);
	const bool aligned4 = ((reinterpret_cast<uintptr_t>(binar) | stride | framePitch) & 3) == 0;
	{
		dim3 grid(static_cast<unsigned>(div_up(height, 4)), B);
		KernelScope ks_("lsl_bits", stream);
		if (aligned4) lsl_bits_kernel<true><<<grid, 128, 0, stream>>>(binar, c->fg.as<unsigned int>(), c->spre.as<unsigned short>(), c->rowCnt.as<unsigned int>(), g);
		else lsl_bits_kernel<false><<<grid, 128, 0, stream>>>(binar, c->fg.as<unsigned int>(), c->spre.as<unsigned short>(), c->rowCnt.as<unsigned int>(), g);
	}
	CVB_LAUNCHED();
	{ KernelScope ks_("lsl_rowscan", stream);
	  lsl_rowscan_kernel<<<B, 1024, 0, stream>>>(c->rowCnt.as<unsigned int>(), c->rowOff.as<unsigned int>(), dFrames, g); }
	CVB_LAUNCHED();
	CVB_CUDA(cudaMemcpyAsync(hf, dFrames, batch * sizeof(LslFrame), cudaMemcpyDeviceToHost, stream));
	CVB_CUDA(cudaStreamSynchronize(stream));
	size_t total = 0;
	for (size_t f = 0; f < batch; ++f) {
		hf[f].segBase = static_cast<unsigned int>(total); hf[f].nea = 0; hf[f].na = 0;
		total += hf[f].nseg;
		CVB_REQUIRE(total < (1ull << 31), CVB200_E_OUT_OF_BOUND);
	}
	if (total) {
		CVB_CHECK(c->segStart.ensure(total * 2));
		CVB_CHECK(c->segEnd.ensure(total * 2));
		CVB_CHECK(c->ov.ensure(total * sizeof(ushort2)));
		CVB_CHECK(c->label.ensure(total * 4));
		CVB_CHECK(c->eq.ensure((total + batch) * 4));
		CVB_CHECK(c->a.ensure((total + batch) * 4));
		CVB_CHECK(c->ranges.ensure(total * sizeof(cvb200_ccl_range_t)));
		CVB_CUDA(cudaMemcpyAsync(dFrames, hf, batch * sizeof(LslFrame), cudaMemcpyHostToDevice, stream));
		{
			dim3 grid(static_cast<unsigned>(div_up(g.WW, 64)), static_cast<unsigned>(height), B);
			KernelScope ks_("lsl_emit", stream);
			lsl_emit_kernel<<<grid, 64, 0, stream>>>(c->fg.as<unsigned int>(), c->spre.as<unsigned short>(), c->rowOff.as<unsigned int>(), dFrames,
				c->segStart.as<short>(), c->segEnd.as<short>(), c->ov.as<ushort2>(), g);
		}
		CVB_LAUNCHED();
		{
			const size_t smem = (static_cast<size_t>(g.ringCap) * 2 + g.eqCap + LSL_STAGE) * 4;
			CVB_CUDA(cudaFuncSetAttribute(lsl_equiv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
			KernelScope ks_("lsl_equiv", stream);
			lsl_equiv_kernel<<<B, 32, smem, stream>>>(c->ov.as<ushort2>(), c->label.as<int>(), c->eq.as<int>(), dFrames, g);
		}
		CVB_LAUNCHED();
		{ KernelScope ks_("lsl_resolve", stream);
		  lsl_resolve_kernel<<<B, 1024, 0, stream>>>(c->eq.as<int>(), c->a.as<int>(), dFrames); }
		CVB_LAUNCHED();
		{
			size_t maxSeg = 0;
			for (size_t f = 0; f < batch; ++f) maxSeg = std::max<size_t>(maxSeg, hf[f].nseg);
			dim3 grid(static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(div_up(maxSeg, 256), 1024))), B);
			KernelScope ks_("lsl_lea", stream);
			lsl_lea_kernel<<<grid, 256, 0, stream>>>(c->label.as<int>(), c->a.as<int>(), c->segStart.as<short>(), c->segEnd.as<short>(), dFrames, c->ranges.as<cvb200_ccl_range_t>());
		}
		CVB_LAUNCHED();
	}
	if (labels) {
		if (total) {
			dim3 grid(static_cast<unsigned>(div_up(width, 256)), static_cast<unsigned>(height), B);
			KernelScope ks_("lsl_flatten", stream);
			lsl_flatten_kernel<<<grid, 256, 0, stream>>>(c->fg.as<unsigned int>(), c->spre.as<unsigned short>(), c->rowOff.as<unsigned int>(), dFrames,
				c->ranges.as<cvb200_ccl_range_t>(), labels, g);
			CVB_LAUNCHED();
		}
		else CVB_CUDA(cudaMemsetAsync(labels, 0, batch * width * height * 4, stream));
	}
	if (!na && !results) return CVB200_S_OK;
	if (total) CVB_CUDA(cudaMemcpyAsync(hf, dFrames, batch * sizeof(LslFrame), cudaMemcpyDeviceToHost, stream));
	if (results) {
		CVB_CHECK(c->hRowOff.ensure(batch * (height + 1) * 4));
		CVB_CUDA(cudaMemcpyAsync(c->hRowOff.p, c->rowOff.p, batch * (height + 1) * 4, cudaMemcpyDeviceToHost, stream));
		if (total) {
			CVB_CHECK(c->hRanges.ensure(total * sizeof(cvb200_ccl_range_t)));
			CVB_CUDA(cudaMemcpyAsync(c->hRanges.p, c->ranges.p, total * sizeof(cvb200_ccl_range_t), cudaMemcpyDeviceToHost, stream));
		}
	}
	CVB_CUDA(cudaStreamSynchronize(stream));
	for (size_t f = 0; f < batch; ++f) {
		if (na) na[f] = hf[f].na;
		if (results) {
			if (!results[f]) { results[f] = new (std::nothrow) cvb200_ccl_result(); CVB_REQUIRE(results[f], CVB200_E_OUT_OF_MEMORY); }
			cvb200_ccl_result* r = results[f];
			r->id = CVB200_PLSL_ID; r->width = width; r->height = height; r->na = hf[f].na;
			r->regionSizes.clear(); r->regionBoxes.clear(); r->regionPoints.clear();
			const uint32_t* ro = c->hRowOff.as<uint32_t>() + f * (height + 1);
			r->rowOffsets.assign(ro, ro + height + 1);
			const cvb200_ccl_range_t* rg = total ? c->hRanges.as<cvb200_ccl_range_t>() + hf[f].segBase : nullptr;
			r->ranges.assign(rg, rg + hf[f].nseg);
		}
	}
	return CVB200_S_OK;
}

extern "C" {

// CompVConnectedComponentLabeling::newObj (base/compv_ccl.cxx:69-97): same parameter checks, same defaults (compv_ccl.h:23-28)
int cvb200_ccl_new_ex(cvb200_ccl_t** ccl, int id, int delta, double minArea, double maxArea, double maxVariation, double minDiversity, int connectivity)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(ccl && delta > 0 && delta <= 255 && minArea >= 0.0 && minArea <= 1.0 && minArea <= maxArea && maxArea >= 0.0 && maxArea <= 1.0
		&& maxVariation >= 0.0 && maxVariation <= 1.0 && minDiversity >= 0.0 && minDiversity <= 1.0 && (connectivity == 4 || connectivity == 8), CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(id == CVB200_PLSL_ID || id == CVB200_LMSER_ID, CVB200_E_INVALID_PARAMETER); // :83-87: unknown factory id
	cvb200_ccl* c = new (std::nothrow) cvb200_ccl();
	CVB_REQUIRE(c, CVB200_E_OUT_OF_MEMORY);
	c->id = id; c->type = CVB200_PLSL_TYPE_XRLEZ; c->sortSegments = false; // ccl_lsl.cxx:118-124
	c->delta = delta; c->minArea = minArea; c->maxArea = maxArea; c->maxVariation = maxVariation; c->minDiversity = minDiversity; c->connectivity = connectivity;
	*ccl = c;
	return CVB200_S_OK;
}

int cvb200_ccl_new(cvb200_ccl_t** ccl, int id)
{
	return cvb200_ccl_new_ex(ccl, id, 5, 0.0002, 0.5, 0.5, 0.5, 8);
}

int cvb200_ccl_free(cvb200_ccl_t** ccl)
{
	if (ccl && *ccl) {
		cvb200_ccl* c = *ccl;
		DevBuf* bufs[] = { &c->fg, &c->spre, &c->rowCnt, &c->rowOff, &c->frames, &c->segStart, &c->segEnd, &c->ov, &c->label, &c->eq, &c->a, &c->ranges, &c->hostIn,
			&c->mUf, &c->mStamp, &c->mPending, &c->mCompSize, &c->mAddSize, &c->mOwnCnt, &c->mTopNode, &c->mPixNode, &c->mOrder, &c->mAbsorbed, &c->mNodeRoot, &c->mNodeParent,
			&c->mNodeArea, &c->mNodeOwn, &c->mNodeLevel, &c->mChild, &c->mSister, &c->mOff, &c->mCursor, &c->mOwnCursor, &c->mVar, &c->mFlags, &c->mDfsPix, &c->mCounters, &c->mRegions,
			&c->mOutOff, &c->mPoints, &c->mBoxes };
		for (DevBuf* b : bufs) b->release();
		c->hFrames.release(); c->hRowOff.release(); c->hRanges.release();
		delete c;
		*ccl = nullptr;
	}
	return CVB200_S_OK;
}

// ccl_lsl.cxx:129-151, compv_ccl.cxx:25-40
int cvb200_ccl_set(cvb200_ccl_t* c, int id, const void* valuePtr, size_t valueSize)
{
	CVB_REQUIRE(c && valuePtr && valueSize, CVB200_E_INVALID_PARAMETER);
	switch (id) {
	case CVB200_PLSL_SET_INT_TYPE:
		CVB_REQUIRE(c->id == CVB200_PLSL_ID, CVB200_E_NOT_IMPLEMENTED); // the LMSER forwards everything to the base class (ccl_lmser.cxx:138-146)
		CVB_REQUIRE(valueSize == sizeof(int), CVB200_E_INVALID_PARAMETER);
		CVB_REQUIRE(*static_cast<const int*>(valuePtr) == CVB200_PLSL_TYPE_XRLEZ, CVB200_E_NOT_IMPLEMENTED);
		c->type = *static_cast<const int*>(valuePtr); return CVB200_S_OK;
	case CVB200_PLSL_SET_BOOL_SORT_SEGMENTS:
		CVB_REQUIRE(c->id == CVB200_PLSL_ID, CVB200_E_NOT_IMPLEMENTED);
		CVB_REQUIRE(valueSize == sizeof(bool), CVB200_E_INVALID_PARAMETER);
		c->sortSegments = *static_cast<const bool*>(valuePtr); return CVB200_S_OK;
	case CVB200_CCL_SET_INT_CONNECTIVITY: {
		CVB_REQUIRE(valueSize == sizeof(int), CVB200_E_INVALID_PARAMETER);
		const int v = *static_cast<const int*>(valuePtr);
		CVB_REQUIRE(v == 4 || v == 8, CVB200_E_NOT_IMPLEMENTED);
		c->connectivity = v; return CVB200_S_OK; // the LSL is 8-connected whatever the value (ccl_lsl.cxx:361-363); the LMSER uses it (ccl_lmser.cxx:172-192)
	}
	default:
		return CVB200_E_NOT_IMPLEMENTED;
	}
}

int cvb200_ccl_process_dev(cvb200_ccl_t* c, const uint8_t* binar, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch,
	int32_t* labels, int32_t* na, cvb200_ccl_result_t** results, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(c && binar && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	CVB_REQUIRE(framePitch >= stride * height, CVB200_E_INVALID_PARAMETER);
	std::lock_guard<std::mutex> lock(c->mutex);
	if (c->id == CVB200_LMSER_ID) {
		CVB_REQUIRE(!labels, CVB200_E_NOT_IMPLEMENTED); // debugFlatten is not implemented for MSER results (lmser_result.cxx:35-39)
		CVB_CHECK(mser_process_dev(c, binar, width, height, stride, batch, framePitch, results, as_stream(stream)));
		if (na) for (size_t f = 0; f < batch; ++f) na[f] = results[f]->na;
		return CVB200_S_OK;
	}
	return lsl_process_dev(c, binar, width, height, stride, batch, framePitch, labels, na, results, as_stream(stream));
}

int cvb200_ccl_process(cvb200_ccl_t* c, const uint8_t* binar, size_t width, size_t height, size_t stride, cvb200_ccl_result_t** result)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(c && binar && width && height && stride >= width && result, CVB200_E_INVALID_PARAMETER); // ccl_lsl.cxx:581-582
	const size_t n = stride * height;
	{
		std::lock_guard<std::mutex> lock(c->mutex);
		CVB_CHECK(c->hostIn.ensure(n));
	}
	CVB_CUDA(cudaMemcpyAsync(c->hostIn.p, binar, n, cudaMemcpyHostToDevice, 0));
	return cvb200_ccl_process_dev(c, c->hostIn.as<uint8_t>(), width, height, stride, 1, n, nullptr, nullptr, result, nullptr);
}

int cvb200_ccl_result_free(cvb200_ccl_result_t** result)
{
	if (result && *result) { delete *result; *result = nullptr; }
	return CVB200_S_OK;
}

size_t cvb200_ccl_result_labels_count(const cvb200_ccl_result_t* result) { return result ? static_cast<size_t>(result->na) : 0; }

int cvb200_ccl_result_segments(const cvb200_ccl_result_t* result, const uint32_t** rowOffsets, const cvb200_ccl_range_t** ranges, size_t* count)
{
	CVB_REQUIRE(result, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(result->id == CVB200_PLSL_ID, CVB200_E_NOT_IMPLEMENTED);
	if (rowOffsets) *rowOffsets = result->rowOffsets.data();
	if (ranges) *ranges = result->ranges.data();
	if (count) *count = result->ranges.size();
	return CVB200_S_OK;
}

// ccl_lsl_result.cxx:51-98 (host loop over the runs, as the reference does: "for visual debugging only")
int cvb200_ccl_result_flatten(const cvb200_ccl_result_t* result, int32_t* labels, size_t labelsStride)
{
	CVB_REQUIRE(result && labels && labelsStride >= result->width, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(result->id == CVB200_PLSL_ID, CVB200_E_NOT_IMPLEMENTED); // lmser_result.cxx:35-39
	CVB_REQUIRE(result->width && result->height && result->rowOffsets.size() == result->height + 1, CVB200_E_INVALID_STATE);
	for (size_t j = 0; j < result->height; ++j) {
		int32_t* row = labels + j * labelsStride;
		memset(row, 0, result->width * sizeof(int32_t));
		for (uint32_t s = result->rowOffsets[j]; s < result->rowOffsets[j + 1]; ++s) {
			const cvb200_ccl_range_t& r = result->ranges[s];
			for (int x = r.start; x < r.end; ++x) row[x] = r.a;
		}
	}
	return CVB200_S_OK;
}

// ccl_lsl_result.cxx:136-185
int cvb200_ccl_result_bounding_boxes(const cvb200_ccl_result_t* result, cvb200_rect16_t* boxes, size_t capacity, size_t* count)
{
	CVB_REQUIRE(result && count && (boxes || !capacity), CVB200_E_INVALID_PARAMETER);
	const size_t na = static_cast<size_t>(result->na);
	*count = na;
	if (!na || !capacity) return CVB200_S_OK;
	const size_t n = std::min(na, capacity);
	if (result->id == CVB200_LMSER_ID) { // lmser_result.cxx:50-88: inclusive min / max of the region's points
		memcpy(boxes, result->regionBoxes.data(), n * sizeof(cvb200_rect16_t));
		return CVB200_S_OK;
	}
	for (size_t k = 0; k < n; ++k) { boxes[k].left = static_cast<int16_t>(result->width); boxes[k].top = static_cast<int16_t>(result->height); boxes[k].right = 0; boxes[k].bottom = 0; }
	for (size_t j = 0; j < result->height; ++j) {
		const int16_t y = static_cast<int16_t>(j);
		for (uint32_t s = result->rowOffsets[j]; s < result->rowOffsets[j + 1]; ++s) {
			const cvb200_ccl_range_t& r = result->ranges[s];
			const size_t a = static_cast<size_t>(r.a - 1);
			if (a >= n) continue;
			cvb200_rect16_t& bb = boxes[a];
			bb.left = std::min(bb.left, r.start);
			bb.top = std::min(bb.top, y);
			bb.right = std::max(bb.right, r.end);
			bb.bottom = y;
		}
	}
	return CVB200_S_OK;
}

// CompVConnectedComponentLabelingResultLMSER::points() (compv_ccl.h:159-170): sizes[i] points per region, boxes[i], points = (x, y) int16 pairs, regions back to back
int cvb200_ccl_result_regions(const cvb200_ccl_result_t* result, const int32_t** sizes, const cvb200_rect16_t** boxes, const int16_t** points, size_t* regionCount, size_t* pointCount)
{
	CVB_REQUIRE(result, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(result->id == CVB200_LMSER_ID, CVB200_E_NOT_IMPLEMENTED);
	if (sizes) *sizes = result->regionSizes.data();
	if (boxes) *boxes = result->regionBoxes.data();
	if (points) *points = result->regionPoints.data();
	if (regionCount) *regionCount = result->regionSizes.size();
	if (pointCount) *pointCount = result->regionPoints.size() / 2;
	return CVB200_S_OK;
}

} // extern "C"
