// a12 -- maximally stable extremal regions (MSER).
// Replaces CompVConnectedComponentLabelingLMSER::process (core/ccl/compv_core_ccl_lmser.cxx:148-410) and the component arithmetic of
// core/include/compv/core/ccl/compv_core_ccl_lmser_result.h (merge :86-91, computeVariation :252-262, computeStability :287-307,
// collectStableRegions :94-119, checkCrit :192-207, computeFinalPoints :122-156) and the boxes of core/ccl/compv_core_ccl_lmser_result.cxx:50-88.
//
// The reference floods the image serially from pixel 0 (Nister & Stewenius) and the component tree falls out of its stack discipline.  The tree
// itself does not depend on the flood: its nodes are the connected components of {I <= t} that own at least one pixel of level exactly t.  Here it is
// built level-synchronously with a lock-free union-find over ALL frames of a batch at once:
//   mser_hist / mser_scatter   counting sort of the pixels by grey level
//   per level t (4 small launches, grid-stride over that level's pixels; empty levels cost only the launches):
//     union     every pixel of level t is united with its neighbours of level <= t (atomicCAS linking, smallest pixel index is the root); roots that
//               lose their root status are appended to an "absorbed" list
//     claim     one node per distinct root among the level's pixels (atomicExch stamp), own-pixel counts
//     attach    pixel -> node; absorbed old roots hand their size and their top node to the component that swallowed them (tree edges)
//     finalize  area = old size + own pixels + absorbed sizes (merge :86-91); the root's previous top node becomes a child
//   then on the node arrays: child lists (atomicExch), variation, stability, and one top-down sweep by grey level that (a) applies the diversity filter of
//   collectStableRegions in the reference's ancestor-before-descendant order and (b) lays the pixels out so that every subtree is one contiguous slice:
//   a region's point list is a slice copy.
// Neighbourhood: the reference indexes its accessibility mask with the image stride (:217-231), so the neighbours of index i are i + {+-1, +-stride, ...}
// wherever that index is a pixel; when stride == width the last pixel of a row is adjacent to the first of the next.  The same rule is used here.
// Regions are returned sorted by (frame, grey level, smallest pixel index); point order inside a region is unspecified (the reference's is its flood order).
#include "ccl.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace cvb {

struct MserGeom {
	int W, H, S;               // width, height, stride: pixel index = y*S + x, FP = H*S indices per frame
	size_t framePitch;
	int B;
	long long total;           // B * H * S
	int conn8;
	int delta, minArea, maxArea;
	double maxVariation, oneMinusDiv, oneMinusDivScale;
	float strideScale;
};

struct MserCounters { unsigned int nodeCount, absorbedCount, regionCount, pad; unsigned int levelStart[257], levelNodeStart[257], levelAbsStart[257]; };
struct MserRegion { int frame, level, root, off, area, node; };

#define MSER_GRID 1184 // 8 CTAs of 256 threads per SM: the level kernels are chains of dependent L2 / HBM accesses (ncu: 27 warps stalled on the long scoreboard per issue), more of them in flight is what helps
#define MSER_BLOCK 256

__device__ __forceinline__ bool mser_valid(const MserGeom& g, int idx) { return idx >= 0 && idx < g.H * g.S && (idx % g.S) < g.W; }

__device__ __forceinline__ int uf_find(int* __restrict__ uf, int x)
{
	int p = __ldcg(&uf[x]);
	while (p != x) {
		const int gp = __ldcg(&uf[p]);
		if (gp != p) uf[x] = gp; // path halving: only non-root entries are rewritten, roots change through atomicCAS alone
		x = p; p = gp;
	}
	return x;
}

__global__ void mser_init_kernel(int* stamp, int* topNode, int* compSize, int* addSize, int* ownCnt, MserGeom g)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < g.total; i += (long long)gridDim.x * blockDim.x) {
		stamp[i] = -1; topNode[i] = -1; compSize[i] = 0; addSize[i] = 0; ownCnt[i] = 0;
	}
}

// Union-find initialisation, one block per image row: every pixel starts attached to the first pixel of its maximal run of equal grey level in the row.
// Those pixels become active at the same level and are 4-connected, so the horizontal unions of flat areas are done before the level loop starts
// (the run start is the smallest index of the run: it stays the root).  Until their level is reached nobody looks at them.
__global__ void __launch_bounds__(256) mser_runstart_kernel(const uint8_t* __restrict__ img, int* __restrict__ uf, MserGeom g)
{
	__shared__ int sWarp[8];
	__shared__ int sCarry;
	const int y = blockIdx.x, f = blockIdx.y;
	const uint8_t* row = img + static_cast<size_t>(f) * g.framePitch + static_cast<size_t>(y) * g.S;
	int* ufRow = uf + (static_cast<size_t>(f) * g.H + y) * g.S;
	const int rowBase = (f * g.H + y) * g.S;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) sCarry = 0;
	__syncthreads();
	for (int x0 = 0; x0 < g.S; x0 += 256) {
		const int x = x0 + threadIdx.x;
		int v = -1; // column where a run starts, else -1
		if (x < g.W && (x == 0 || row[x - 1] != row[x])) v = x;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v = max(v, u); }
		if (lane == 31) sWarp[warp] = v;
		__syncthreads();
		int pre = sCarry;
		for (int w = 0; w < warp; ++w) pre = max(pre, sWarp[w]);
		v = max(v, pre);
		if (x < g.S) ufRow[x] = rowBase + ((x < g.W) ? v : x);
		__syncthreads();
		if (threadIdx.x == 255) sCarry = v;
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256) mser_hist_kernel(const uint8_t* __restrict__ img, MserCounters* cnt, MserGeom g)
{
	__shared__ unsigned int h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	const int FP = g.H * g.S;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < g.total; i += gridDim.x * 256ll) {
		const int f = static_cast<int>(i / FP), idx = static_cast<int>(i % FP);
		if ((idx % g.S) < g.W) atomicAdd(&h[img[static_cast<size_t>(f) * g.framePitch + idx]], 1u);
	}
	__syncthreads();
	if (h[threadIdx.x]) atomicAdd(&cnt->levelStart[threadIdx.x + 1], h[threadIdx.x]);
}

__global__ void mser_levelscan_kernel(MserCounters* cnt, unsigned int* cursor)
{
	if (threadIdx.x == 0) {
		unsigned int run = 0;
		cnt->levelStart[0] = 0;
		for (int t = 1; t <= 256; ++t) { run += cnt->levelStart[t]; cnt->levelStart[t] = run; }
		for (int t = 0; t < 256; ++t) cursor[t] = cnt->levelStart[t];
	}
}

__global__ void __launch_bounds__(256) mser_scatter_kernel(const uint8_t* __restrict__ img, unsigned int* cursor, int* __restrict__ order, MserGeom g)
{
	__shared__ unsigned int h[256], base[256];
	const int FP = g.H * g.S;
	// each block owns a contiguous span of indices: local histogram -> one reservation per level -> scatter
	const long long per = (g.total + gridDim.x - 1) / gridDim.x;
	const long long b0 = blockIdx.x * per, b1 = min(g.total, b0 + per);
	h[threadIdx.x] = 0;
	__syncthreads();
	for (long long i = b0 + threadIdx.x; i < b1; i += 256) {
		const int f = static_cast<int>(i / FP), idx = static_cast<int>(i % FP);
		if ((idx % g.S) < g.W) atomicAdd(&h[img[static_cast<size_t>(f) * g.framePitch + idx]], 1u);
	}
	__syncthreads();
	base[threadIdx.x] = h[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], h[threadIdx.x]) : 0u;
	h[threadIdx.x] = 0;
	__syncthreads();
	for (long long i = b0 + threadIdx.x; i < b1; i += 256) {
		const int f = static_cast<int>(i / FP), idx = static_cast<int>(i % FP);
		if ((idx % g.S) < g.W) {
			const int l = img[static_cast<size_t>(f) * g.framePitch + idx];
			order[base[l] + atomicAdd(&h[l], 1u)] = static_cast<int>(i);
		}
	}
}

// ---- level t: union ----
__global__ void __launch_bounds__(MSER_BLOCK) mser_union_kernel(const uint8_t* __restrict__ img, const int* __restrict__ order, int* __restrict__ uf, int* __restrict__ absorbed,
	MserCounters* cnt, int t, MserGeom g)
{
	const unsigned int i0 = cnt->levelStart[t], i1 = cnt->levelStart[t + 1];
	const int FP = g.H * g.S;
	const int nEdges = g.conn8 ? 8 : 4;
	const int offs8[8] = { 1, 1 - g.S, -g.S, -g.S - 1, -1, g.S - 1, g.S + 1, g.S };
	const int offs4[4] = { 1, -g.S, -1, g.S };
	const unsigned int lane = threadIdx.x & 31u, ltMask = (1u << lane) - 1u;
	// warp-uniform trip count: the append to the absorbed list below is aggregated per warp (one atomicAdd on the shared counter per warp and edge instead of one per
	// union: ~250 k same-address atomics per level were the bulk of this kernel's time)
	for (unsigned int base = i0 + blockIdx.x * MSER_BLOCK; base < i1; base += gridDim.x * MSER_BLOCK) {
		const unsigned int i = base + threadIdx.x;
		const bool active = i < i1;
		const int p = active ? order[i] : 0;
		const int f = p / FP, idx = p - f * FP;
		const uint8_t* frame = img + static_cast<size_t>(f) * g.framePitch;
		for (int e = 0; e < nEdges; ++e) {
			int gone = -1; // the root this thread's union removed, if any
			const int q = idx + (g.conn8 ? offs8[e] : offs4[e]);
			if (active && mser_valid(g, q)) {
				const int lq = frame[q];
				// higher levels later; equal levels are united once, from the larger index; the left neighbour of the same level was united by mser_runstart
				if (!(lq > t || (lq == t && q > idx)) && !(lq == t && q == idx - 1 && (idx % g.S) != 0)) {
					int ra = uf_find(uf, p), rb = uf_find(uf, f * FP + q);
					while (ra != rb) {
						if (ra < rb) { const int tmp = ra; ra = rb; rb = tmp; }
						const int old = atomicCAS(&uf[ra], ra, rb);
						if (old == ra) { gone = ra; break; }
						ra = uf_find(uf, old);
						rb = uf_find(uf, rb);
					}
				}
			}
			__syncwarp();
			const unsigned int bal = __ballot_sync(0xffffffffu, gone >= 0);
			if (bal) {
				unsigned int at = 0;
				if (lane == static_cast<unsigned int>(__ffs(bal) - 1)) at = atomicAdd(&cnt->absorbedCount, static_cast<unsigned int>(__popc(bal)));
				at = __shfl_sync(0xffffffffu, at, __ffs(bal) - 1);
				if (gone >= 0) absorbed[at + __popc(bal & ltMask)] = gone;
			}
		}
	}
}

// ---- level t: one node per root, own-pixel counts ----
__global__ void __launch_bounds__(MSER_BLOCK) mser_claim_kernel(const int* __restrict__ order, int* __restrict__ uf, int* __restrict__ stamp, int* __restrict__ pendingNode,
	int* __restrict__ ownCnt, int* __restrict__ nodeRoot, MserCounters* cnt, int t)
{
	const unsigned int i0 = cnt->levelStart[t], i1 = cnt->levelStart[t + 1];
	const unsigned int lane = threadIdx.x & 31u, ltMask = (1u << lane) - 1u;
	for (unsigned int base = i0 + blockIdx.x * MSER_BLOCK; base < i1; base += gridDim.x * MSER_BLOCK) { // warp-uniform trip count: node numbers are reserved once per warp
		const unsigned int i = base + threadIdx.x;
		const bool active = i < i1;
		const int r = active ? uf_find(uf, order[i]) : -1 - static_cast<int>(lane); // inactive lanes: distinct keys nobody shares
		// lanes of the warp that found the same root speak once (flat areas put whole warps on one root)
		const unsigned int peers = __match_any_sync(0xffffffffu, r);
		bool fresh = false;
		if (active && (__ffs(peers) - 1) == static_cast<int>(lane)) {
			atomicAdd(&ownCnt[r], __popc(peers));
			fresh = atomicExch(&stamp[r], t) != t;
		}
		const unsigned int bal = __ballot_sync(0xffffffffu, fresh);
		if (bal) {
			unsigned int at = 0;
			if (lane == static_cast<unsigned int>(__ffs(bal) - 1)) at = atomicAdd(&cnt->nodeCount, static_cast<unsigned int>(__popc(bal)));
			at = __shfl_sync(0xffffffffu, at, __ffs(bal) - 1);
			if (fresh) {
				const int n = static_cast<int>(at + __popc(bal & ltMask));
				nodeRoot[n] = r;
				pendingNode[r] = n;
			}
		}
	}
}

// ---- level t: pixel -> node; absorbed components -> their new owner ----
__global__ void __launch_bounds__(MSER_BLOCK) mser_attach_kernel(const int* __restrict__ order, int* __restrict__ uf, const int* __restrict__ pendingNode, int* __restrict__ pixNode,
	const int* __restrict__ absorbed, const int* __restrict__ compSize, int* __restrict__ addSize, const int* __restrict__ topNode, int* __restrict__ nodeParent, MserCounters* cnt, int t)
{
	const unsigned int i0 = cnt->levelStart[t], i1 = cnt->levelStart[t + 1];
	for (unsigned int i = i0 + blockIdx.x * MSER_BLOCK + threadIdx.x; i < i1; i += gridDim.x * MSER_BLOCK) {
		const int p = order[i];
		pixNode[p] = pendingNode[uf_find(uf, p)];
	}
	const unsigned int a0 = cnt->levelAbsStart[t], a1 = cnt->absorbedCount;
	for (unsigned int i = a0 + blockIdx.x * MSER_BLOCK + threadIdx.x; i < a1; i += gridDim.x * MSER_BLOCK) {
		const int a = absorbed[i];
		const int sz = compSize[a];
		if (sz > 0) { // a component that existed before this level: it becomes a child of the node of the component that absorbed it
			const int r = uf_find(uf, a);
			atomicAdd(&addSize[r], sz);
			nodeParent[topNode[a]] = pendingNode[r];
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) { cnt->levelNodeStart[t + 1] = cnt->nodeCount; cnt->levelAbsStart[t + 1] = a1; }
}

// ---- level t: areas, tree edges of persisting roots ----
__global__ void __launch_bounds__(MSER_BLOCK) mser_finalize_kernel(const int* __restrict__ nodeRoot, int* __restrict__ compSize, int* __restrict__ addSize, int* __restrict__ ownCnt,
	int* __restrict__ topNode, int* __restrict__ nodeParent, int* __restrict__ nodeArea, int* __restrict__ nodeOwn, int* __restrict__ nodeLevel, MserCounters* cnt, int t)
{
	const unsigned int n0 = cnt->levelNodeStart[t], n1 = cnt->levelNodeStart[t + 1];
	for (unsigned int n = n0 + blockIdx.x * MSER_BLOCK + threadIdx.x; n < n1; n += gridDim.x * MSER_BLOCK) {
		const int r = nodeRoot[n];
		const int own = ownCnt[r];
		const int area = compSize[r] + own + addSize[r];
		compSize[r] = area; addSize[r] = 0; ownCnt[r] = 0;
		const int prev = topNode[r];
		if (prev >= 0) nodeParent[prev] = static_cast<int>(n);
		topNode[r] = static_cast<int>(n);
		nodeParent[n] = -1;
		nodeArea[n] = area; nodeOwn[n] = own; nodeLevel[n] = t;
	}
}

// ---- tree analytics ----
__global__ void mser_children_kernel(const int* __restrict__ nodeParent, int* __restrict__ child, int* __restrict__ sister, unsigned int M)
{
	for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < M; c += gridDim.x * blockDim.x) {
		const int p = nodeParent[c];
		sister[c] = (p >= 0) ? atomicExch(&child[p], static_cast<int>(c)) : -1;
	}
}

__global__ void mser_variation_kernel(const int* __restrict__ nodeParent, const int* __restrict__ nodeArea, const int* __restrict__ nodeLevel, double* __restrict__ var, unsigned int M, MserGeom g)
{
	for (unsigned int n = blockIdx.x * blockDim.x + threadIdx.x; n < M; n += gridDim.x * blockDim.x) {
		const int deltaPlus = nodeLevel[n] + g.delta; // lmser_result.h:255-260
		int a = static_cast<int>(n), p;
		while ((p = nodeParent[a]) >= 0 && nodeLevel[p] <= deltaPlus) a = p;
		var[n] = (nodeArea[a] - nodeArea[n]) / static_cast<double>(nodeArea[n]);
	}
}

__global__ void mser_childflags_kernel(const int* __restrict__ nodeParent, const double* __restrict__ var, unsigned char* __restrict__ flags, unsigned int M)
{
	for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < M; c += gridDim.x * blockDim.x) {
		const int p = nodeParent[c];
		if (p >= 0 && var[p] < var[c]) flags[p] = 1; // some child is less stable than its parent (:296-301)
	}
}

__global__ void mser_stability_kernel(const int* __restrict__ nodeParent, const int* __restrict__ nodeArea, const int* __restrict__ child, const double* __restrict__ var,
	const unsigned char* __restrict__ flags, unsigned char* __restrict__ stable0, unsigned int M, MserGeom g)
{
	for (unsigned int n = blockIdx.x * blockDim.x + threadIdx.x; n < M; n += gridDim.x * blockDim.x) {
		const int p = nodeParent[n];
		const double v = var[n];
		const bool s = (p < 0 || var[p] >= v) && (v <= g.maxVariation) && (g.minArea <= nodeArea[n] && nodeArea[n] <= g.maxArea); // :291-292
		stable0[n] = (child[n] >= 0) ? (s && flags[n]) : s;
	}
}

// checkCrit (:192-207) as a stackless walk over child / sister / parent links
__device__ bool mser_check_crit(int n, int area_, double v, const int* __restrict__ nodeParent, const int* __restrict__ nodeArea, const int* __restrict__ child, const int* __restrict__ sister,
	const double* __restrict__ var, const unsigned char* __restrict__ stable0)
{
	if (nodeArea[n] <= area_) return true;
	int cur = child[n];
	if (cur < 0) return true;
	for (;;) {
		bool descend = false;
		if (nodeArea[cur] > area_) {
			if (stable0[cur] && var[cur] < v) return false;
			if (child[cur] >= 0) descend = true;
		}
		if (descend) { cur = child[cur]; continue; }
		while (sister[cur] < 0) {
			cur = nodeParent[cur];
			if (cur == n) return true;
		}
		cur = sister[cur];
	}
}

// one grey level of the top-down sweep: subtree slices + the diversity filter of collectStableRegions (:94-119)
__global__ void mser_topdown_kernel(const int* __restrict__ nodeParent, const int* __restrict__ nodeArea, const int* __restrict__ nodeOwn, const int* __restrict__ nodeRoot,
	const int* __restrict__ child, const int* __restrict__ sister, const double* __restrict__ var, const unsigned char* __restrict__ stable0, unsigned char* __restrict__ fin,
	int* __restrict__ off, int* __restrict__ cursor, unsigned int n0, unsigned int n1, MserGeom g)
{
	for (unsigned int n = n0 + blockIdx.x * blockDim.x + threadIdx.x; n < n1; n += gridDim.x * blockDim.x) {
		const int p = nodeParent[n];
		const int area = nodeArea[n];
		off[n] = (p < 0) ? (nodeRoot[n] / (g.H * g.S)) * (g.W * g.H) : off[p] + nodeOwn[p] + atomicAdd(&cursor[p], area);
		bool st = stable0[n] != 0;
		if (st) {
			const double v = var[n];
			const int minParentArea = static_cast<int>((area * g.oneMinusDivScale) + 0.5);
			for (int a = p; a >= 0 && nodeArea[a] < minParentArea && (st = (!fin[a] || var[a] > v)); a = nodeParent[a]) { }
			st = st && mser_check_crit(static_cast<int>(n), static_cast<int>((area * g.oneMinusDiv) + 0.5), v, nodeParent, nodeArea, child, sister, var, stable0);
		}
		fin[n] = st ? 1 : 0;
	}
}

__global__ void mser_layout_kernel(const int* __restrict__ order, const int* __restrict__ pixNode, const int* __restrict__ off, int* __restrict__ ownCursor, short2* __restrict__ dfsPix,
	unsigned int count, MserGeom g)
{
	const int FP = g.H * g.S;
	for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
		const int p = order[i];
		const int n = pixNode[p];
		const int idx = p % FP;
		const short y = static_cast<short>(__fmul_rn(static_cast<float>(idx), g.strideScale)); // computeFinalPoints :145-146, float arithmetic included
		const short x = static_cast<short>(idx - (y * g.S));
		const unsigned int peers = __match_any_sync(__activemask(), n);
		const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
		int base = 0;
		if (lane == leader) base = atomicAdd(&ownCursor[n], __popc(peers));
		base = __shfl_sync(peers, base, leader);
		dfsPix[off[n] + base + __popc(peers & ((1u << lane) - 1u))] = make_short2(x, y);
	}
}

__global__ void mser_regions_kernel(const unsigned char* __restrict__ fin, const int* __restrict__ nodeRoot, const int* __restrict__ nodeLevel, const int* __restrict__ nodeArea,
	const int* __restrict__ off, MserRegion* __restrict__ regions, MserCounters* cnt, unsigned int M, MserGeom g)
{
	for (unsigned int n = blockIdx.x * blockDim.x + threadIdx.x; n < M; n += gridDim.x * blockDim.x) {
		if (!fin[n]) continue;
		MserRegion r;
		r.frame = nodeRoot[n] / (g.H * g.S); r.level = nodeLevel[n]; r.root = nodeRoot[n]; r.off = off[n]; r.area = nodeArea[n]; r.node = static_cast<int>(n);
		regions[atomicAdd(&cnt->regionCount, 1u)] = r;
	}
}

// one block per region: slice copy + bounding box (lmser_result.cxx:60-74: inclusive min / max)
__global__ void __launch_bounds__(256) mser_gather_kernel(const MserRegion* __restrict__ regions, const unsigned long long* __restrict__ outOff, const short2* __restrict__ dfsPix,
	short2* __restrict__ points, cvb200_rect16_t* __restrict__ boxes)
{
	__shared__ int sMin[2][8], sMax[2][8];
	const MserRegion r = regions[blockIdx.x];
	const short2* src = dfsPix + r.off;
	short2* dst = points + outOff[blockIdx.x];
	int xmin = 32767, xmax = -32768, ymin = 32767, ymax = -32768;
	for (int i = threadIdx.x; i < r.area; i += 256) {
		const short2 v = src[i];
		dst[i] = v;
		xmin = min(xmin, (int)v.x); xmax = max(xmax, (int)v.x); ymin = min(ymin, (int)v.y); ymax = max(ymax, (int)v.y);
	}
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, d)); xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, d));
		ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, d)); ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, d));
	}
	const int warp = threadIdx.x >> 5;
	if ((threadIdx.x & 31) == 0) { sMin[0][warp] = xmin; sMax[0][warp] = xmax; sMin[1][warp] = ymin; sMax[1][warp] = ymax; }
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < 8; ++w) { xmin = min(xmin, sMin[0][w]); xmax = max(xmax, sMax[0][w]); ymin = min(ymin, sMin[1][w]); ymax = max(ymax, sMax[1][w]); }
		cvb200_rect16_t b; b.left = static_cast<int16_t>(xmin); b.top = static_cast<int16_t>(ymin); b.right = static_cast<int16_t>(xmax); b.bottom = static_cast<int16_t>(ymax);
		boxes[blockIdx.x] = b;
	}
}

static int mser_chunk(cvb200_ccl* c, const uint8_t* img, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch, cvb200_ccl_result_t** results, cudaStream_t stream)
{
	MserGeom g;
	memset(&g, 0, sizeof(g));
	g.W = static_cast<int>(width); g.H = static_cast<int>(height); g.S = static_cast<int>(stride); g.framePitch = framePitch; g.B = static_cast<int>(batch);
	g.total = static_cast<long long>(batch) * height * stride;
	g.conn8 = (c->connectivity == 8);
	g.delta = c->delta;
	const int inputArea = static_cast<int>(width * height);
	g.minArea = static_cast<int>(inputArea * c->minArea); g.maxArea = static_cast<int>(inputArea * c->maxArea); // ccl_lmser.cxx:374-375
	g.maxVariation = c->maxVariation;
	g.oneMinusDiv = 1.0 - c->minDiversity; g.oneMinusDivScale = 1.0 / g.oneMinusDiv; // :376-377
	g.strideScale = 1.f / static_cast<float>(stride); // :387
	const size_t T = static_cast<size_t>(g.total), NP = batch * width * height; // index space, pixels (= node capacity)
	DevBuf* ib[] = { &c->mUf, &c->mStamp, &c->mPending, &c->mCompSize, &c->mAddSize, &c->mOwnCnt, &c->mTopNode, &c->mPixNode };
	for (DevBuf* b : ib) CVB_CHECK(b->ensure(T * 4));
	DevBuf* nb[] = { &c->mOrder, &c->mAbsorbed, &c->mNodeRoot, &c->mNodeParent, &c->mNodeArea, &c->mNodeOwn, &c->mNodeLevel, &c->mChild, &c->mSister, &c->mOff, &c->mCursor, &c->mOwnCursor };
	for (DevBuf* b : nb) CVB_CHECK(b->ensure((NP + 1) * 4));
	CVB_CHECK(c->mVar.ensure((NP + 1) * 8));
	CVB_CHECK(c->mFlags.ensure((NP + 1) * 3));
	CVB_CHECK(c->mDfsPix.ensure((NP + 1) * sizeof(short2)));
	CVB_CHECK(c->mCounters.ensure(sizeof(MserCounters) + 256 * 4));
	CVB_CHECK(c->hFrames.ensure(sizeof(MserCounters)));
	MserCounters* dCnt = c->mCounters.as<MserCounters>();
	unsigned int* dLevelCursor = reinterpret_cast<unsigned int*>(dCnt + 1);
	MserCounters* hCnt = c->hFrames.as<MserCounters>();
	int* uf = c->mUf.as<int>();
	unsigned char* flags = c->mFlags.as<unsigned char>();
	unsigned char* stable0 = flags + (NP + 1);
	unsigned char* fin = stable0 + (NP + 1);

	CVB_CUDA(cudaMemsetAsync(dCnt, 0, sizeof(MserCounters) + 256 * 4, stream));
	{ KernelScope ks_("mser_init", stream);
	  mser_init_kernel<<<MSER_GRID, 256, 0, stream>>>(c->mStamp.as<int>(), c->mTopNode.as<int>(), c->mCompSize.as<int>(), c->mAddSize.as<int>(), c->mOwnCnt.as<int>(), g);
	  mser_runstart_kernel<<<dim3(static_cast<unsigned>(height), static_cast<unsigned>(batch)), 256, 0, stream>>>(img, uf, g); }
	CVB_LAUNCHED(); g_launches.fetch_add(1, std::memory_order_relaxed);
	{ KernelScope ks_("mser_sort", stream);
	  mser_hist_kernel<<<MSER_GRID, 256, 0, stream>>>(img, dCnt, g);
	  mser_levelscan_kernel<<<1, 32, 0, stream>>>(dCnt, dLevelCursor);
	  mser_scatter_kernel<<<MSER_GRID, 256, 0, stream>>>(img, dLevelCursor, c->mOrder.as<int>(), g); }
	CVB_LAUNCHED(); g_launches.fetch_add(2, std::memory_order_relaxed);
	{
		static const bool split = getenv("CVB200_MSER_SPLIT") != nullptr; // profiling aid: one scope per kernel of a level instead of one for the loop
		KernelScope ks_(split ? "mser_levels_all" : "mser_levels", stream);
		for (int t = 0; t < 256; ++t) {
			{ KernelScope k1(split ? "mser_union" : nullptr, split ? stream : nullptr, split);
			mser_union_kernel<<<MSER_GRID, MSER_BLOCK, 0, stream>>>(img, c->mOrder.as<int>(), uf, c->mAbsorbed.as<int>(), dCnt, t, g); }
			{ KernelScope k2(split ? "mser_claim" : nullptr, split ? stream : nullptr, split);
			mser_claim_kernel<<<MSER_GRID, MSER_BLOCK, 0, stream>>>(c->mOrder.as<int>(), uf, c->mStamp.as<int>(), c->mPending.as<int>(), c->mOwnCnt.as<int>(), c->mNodeRoot.as<int>(), dCnt, t); }
			{ KernelScope k3(split ? "mser_attach" : nullptr, split ? stream : nullptr, split);
			mser_attach_kernel<<<MSER_GRID, MSER_BLOCK, 0, stream>>>(c->mOrder.as<int>(), uf, c->mPending.as<int>(), c->mPixNode.as<int>(), c->mAbsorbed.as<int>(), c->mCompSize.as<int>(),
				c->mAddSize.as<int>(), c->mTopNode.as<int>(), c->mNodeParent.as<int>(), dCnt, t); }
			{ KernelScope k4(split ? "mser_finalize" : nullptr, split ? stream : nullptr, split);
			mser_finalize_kernel<<<MSER_GRID, MSER_BLOCK, 0, stream>>>(c->mNodeRoot.as<int>(), c->mCompSize.as<int>(), c->mAddSize.as<int>(), c->mOwnCnt.as<int>(), c->mTopNode.as<int>(),
				c->mNodeParent.as<int>(), c->mNodeArea.as<int>(), c->mNodeOwn.as<int>(), c->mNodeLevel.as<int>(), dCnt, t); }
		}
	}
	CVB_LAUNCHED(); g_launches.fetch_add(1023, std::memory_order_relaxed);
	CVB_CUDA(cudaMemcpyAsync(hCnt, dCnt, sizeof(MserCounters), cudaMemcpyDeviceToHost, stream));
	CVB_CUDA(cudaStreamSynchronize(stream));
	const unsigned int M = hCnt->nodeCount;
	CVB_REQUIRE(M >= 1 && M <= NP && hCnt->levelStart[256] == NP, CVB200_E_INVALID_STATE);
	const unsigned int gridM = static_cast<unsigned int>(std::min<size_t>(div_up(M, 256), 4096));
	CVB_CUDA(cudaMemsetAsync(c->mChild.p, 0xff, static_cast<size_t>(M) * 4, stream));
	CVB_CUDA(cudaMemsetAsync(c->mCursor.p, 0, static_cast<size_t>(M) * 4, stream));
	CVB_CUDA(cudaMemsetAsync(c->mOwnCursor.p, 0, static_cast<size_t>(M) * 4, stream));
	CVB_CUDA(cudaMemsetAsync(flags, 0, static_cast<size_t>(M), stream));
	{ KernelScope ks_("mser_tree", stream);
	  mser_children_kernel<<<gridM, 256, 0, stream>>>(c->mNodeParent.as<int>(), c->mChild.as<int>(), c->mSister.as<int>(), M);
	  mser_variation_kernel<<<gridM, 256, 0, stream>>>(c->mNodeParent.as<int>(), c->mNodeArea.as<int>(), c->mNodeLevel.as<int>(), c->mVar.as<double>(), M, g);
	  mser_childflags_kernel<<<gridM, 256, 0, stream>>>(c->mNodeParent.as<int>(), c->mVar.as<double>(), flags, M);
	  mser_stability_kernel<<<gridM, 256, 0, stream>>>(c->mNodeParent.as<int>(), c->mNodeArea.as<int>(), c->mChild.as<int>(), c->mVar.as<double>(), flags, stable0, M, g); }
	CVB_LAUNCHED(); g_launches.fetch_add(3, std::memory_order_relaxed);
	{
		KernelScope ks_("mser_topdown", stream);
		for (int t = 255; t >= 0; --t) {
			const unsigned int n0 = hCnt->levelNodeStart[t], n1 = hCnt->levelNodeStart[t + 1];
			if (n1 <= n0) continue;
			const unsigned int grid = static_cast<unsigned int>(std::min<size_t>(div_up(n1 - n0, 128), 2048));
			mser_topdown_kernel<<<grid, 128, 0, stream>>>(c->mNodeParent.as<int>(), c->mNodeArea.as<int>(), c->mNodeOwn.as<int>(), c->mNodeRoot.as<int>(), c->mChild.as<int>(), c->mSister.as<int>(),
				c->mVar.as<double>(), stable0, fin, c->mOff.as<int>(), c->mCursor.as<int>(), n0, n1, g);
			g_launches.fetch_add(1, std::memory_order_relaxed);
		}
	}
	CVB_CUDA(cudaGetLastError());
	CVB_CHECK(c->mRegions.ensure(static_cast<size_t>(M) * sizeof(MserRegion)));
	{ KernelScope ks_("mser_layout", stream);
	  mser_layout_kernel<<<MSER_GRID, 256, 0, stream>>>(c->mOrder.as<int>(), c->mPixNode.as<int>(), c->mOff.as<int>(), c->mOwnCursor.as<int>(), c->mDfsPix.as<short2>(), static_cast<unsigned int>(NP), g);
	  mser_regions_kernel<<<gridM, 256, 0, stream>>>(fin, c->mNodeRoot.as<int>(), c->mNodeLevel.as<int>(), c->mNodeArea.as<int>(), c->mOff.as<int>(), c->mRegions.as<MserRegion>(), dCnt, M, g); }
	CVB_LAUNCHED(); g_launches.fetch_add(1, std::memory_order_relaxed);
	CVB_CUDA(cudaMemcpyAsync(hCnt, dCnt, 16, cudaMemcpyDeviceToHost, stream));
	CVB_CUDA(cudaStreamSynchronize(stream));
	const unsigned int R = hCnt->regionCount;
	std::vector<MserRegion> regs(R);
	if (R) CVB_CUDA(cudaMemcpy(regs.data(), c->mRegions.p, static_cast<size_t>(R) * sizeof(MserRegion), cudaMemcpyDeviceToHost));
	std::sort(regs.begin(), regs.end(), [](const MserRegion& a, const MserRegion& b) {
		if (a.frame != b.frame) return a.frame < b.frame;
		if (a.level != b.level) return a.level < b.level;
		return a.root < b.root;
	});
	std::vector<unsigned long long> outOff(R + 1, 0);
	for (unsigned int i = 0; i < R; ++i) outOff[i + 1] = outOff[i] + static_cast<unsigned long long>(regs[i].area);
	const size_t totalPts = static_cast<size_t>(outOff[R]);
	std::vector<short2> hPts(totalPts);
	std::vector<cvb200_rect16_t> hBoxes(R);
	if (R) {
		CVB_CHECK(c->mOutOff.ensure((R + 1) * 8));
		CVB_CHECK(c->mPoints.ensure(std::max<size_t>(totalPts, 1) * sizeof(short2)));
		CVB_CHECK(c->mBoxes.ensure(static_cast<size_t>(R) * sizeof(cvb200_rect16_t)));
		CVB_CUDA(cudaMemcpyAsync(c->mRegions.p, regs.data(), static_cast<size_t>(R) * sizeof(MserRegion), cudaMemcpyHostToDevice, stream));
		CVB_CUDA(cudaMemcpyAsync(c->mOutOff.p, outOff.data(), (R + 1) * 8, cudaMemcpyHostToDevice, stream));
		{ KernelScope ks_("mser_gather", stream);
		  mser_gather_kernel<<<R, 256, 0, stream>>>(c->mRegions.as<MserRegion>(), c->mOutOff.as<unsigned long long>(), c->mDfsPix.as<short2>(), c->mPoints.as<short2>(), c->mBoxes.as<cvb200_rect16_t>()); }
		CVB_LAUNCHED();
		CVB_CUDA(cudaMemcpyAsync(hPts.data(), c->mPoints.p, totalPts * sizeof(short2), cudaMemcpyDeviceToHost, stream));
		CVB_CUDA(cudaMemcpyAsync(hBoxes.data(), c->mBoxes.p, static_cast<size_t>(R) * sizeof(cvb200_rect16_t), cudaMemcpyDeviceToHost, stream));
		CVB_CUDA(cudaStreamSynchronize(stream));
	}
	// per-frame results
	size_t i = 0;
	for (size_t f = 0; f < batch; ++f) {
		if (!results[f]) { results[f] = new (std::nothrow) cvb200_ccl_result(); CVB_REQUIRE(results[f], CVB200_E_OUT_OF_MEMORY); }
		cvb200_ccl_result* r = results[f];
		r->id = CVB200_LMSER_ID; r->width = width; r->height = height; r->na = 0;
		r->rowOffsets.clear(); r->ranges.clear(); r->regionSizes.clear(); r->regionBoxes.clear(); r->regionPoints.clear();
		const size_t first = i;
		while (i < R && regs[i].frame == static_cast<int>(f)) ++i;
		for (size_t k = first; k < i; ++k) { r->regionSizes.push_back(regs[k].area); r->regionBoxes.push_back(hBoxes[k]); }
		if (i > first) {
			const int16_t* p0 = reinterpret_cast<const int16_t*>(hPts.data() + outOff[first]);
			r->regionPoints.assign(p0, p0 + 2 * (outOff[i] - outOff[first]));
		}
		r->na = static_cast<int32_t>(i - first);
	}
	return CVB200_S_OK;
}

int mser_process_dev(cvb200_ccl* c, const uint8_t* img, size_t width, size_t height, size_t stride, size_t batch, size_t framePitch, cvb200_ccl_result_t** results, cudaStream_t stream)
{
	CVB_REQUIRE(results, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(width <= 32767 && height <= 32767 && height * stride < (1ull << 28), CVB200_E_OUT_OF_BOUND); // int16 points; the reference packs the pixel index in 28 bits (:313)
	// frames per pass: the per-pixel tables take ~100 B/px
	size_t chunk = (size_t(3) << 30) / (height * stride * 100);
	if (chunk < 1) chunk = 1;
	if (chunk > batch) chunk = batch;
	while (chunk * height * stride >= (1ull << 31)) --chunk;
	for (size_t f0 = 0; f0 < batch; f0 += chunk) {
		const size_t nf = std::min(chunk, batch - f0);
		CVB_CHECK(mser_chunk(c, img + f0 * framePitch, width, height, stride, nf, framePitch, results + f0, stream));
	}
	return CVB200_S_OK;
}

} // namespace cvb
