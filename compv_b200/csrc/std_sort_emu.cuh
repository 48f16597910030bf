// Reproduces, element for element, the permutation that libstdc++'s std::sort produces for a given input order -- i.e. the tie order the reference's Hough
// detectors get from `std::sort(votes.begin(), votes.end(), [](a, b) { return a.count > b.count; })` (core/features/hough/compv_core_feature_houghkht.cxx:1195-1204,
// compv_core_feature_houghsht.cxx) -- without running on the host.  The algorithm restated here is the published libstdc++ introsort (bits/stl_algo.h,
// GCC 13: __introsort_loop with median-of-three to *first, Hoare-style __unguarded_partition, depth limit 2*floor(log2 n) with a heap-sort fallback,
// _S_threshold = 16, then __final_insertion_sort), written from its behaviour:
//
//   * a partition step only touches its own range and is a pure function of that range's contents, so the recursion tree can be evaluated in any order and
//     its nodes in parallel;
//   * the Hoare partition has a closed form: with L = ascending positions of elements NOT less than the pivot and R = descending positions of elements NOT
//     greater than it (both inside [first+1, last)), the loop swaps the pairs (L[k], R[k]) for k < K, K = #{k : L[k] < R[k]}, and returns
//     cut = min(L[K], R[K-1]) (L[K] = +inf when absent, R[-1] = last).  That turns a partition into two ordered compactions + K independent swaps;
//   * the final insertion sort is a stable sort, and after the introsort loop every unsorted stretch lies inside one leaf range (<= 16 elements) whose
//     neighbours are already on the correct side: it is the same as insertion-sorting every leaf on its own.
//
// Elements are 64-bit: key in the high word, payload (original index) in the low word; "less" is `key(a) > key(b)` (descending keys), as in the reference.
// Everything here is host+device so that tests/cpp/sort_check.cpp can pin it on std::sort itself without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SSE_FN __host__ __device__ __forceinline__
#else
#define SSE_FN inline
#endif

namespace cvb {

typedef unsigned long long sse_item;
SSE_FN unsigned int sse_key(sse_item v) { return static_cast<unsigned int>(v >> 32); }
SSE_FN bool sse_less(sse_item a, sse_item b) { return sse_key(a) > sse_key(b); }   // the reference's comparator: a.count > b.count

#define SSE_THRESHOLD 16

SSE_FN int sse_lg(unsigned int n) { int k = 0; while (n >>= 1) ++k; return k; } // std::__lg

SSE_FN void sse_move_median_to_first(sse_item* a, int result, int ia, int ib, int ic)
{
	const sse_item va = a[ia], vb = a[ib], vc = a[ic];
	int m;
	if (sse_less(va, vb)) m = sse_less(vb, vc) ? ib : (sse_less(va, vc) ? ic : ia);
	else m = sse_less(va, vc) ? ia : (sse_less(vb, vc) ? ic : ib);
	const sse_item t = a[result]; a[result] = a[m]; a[m] = t;
}

// ---- heap-sort fallback (std::__partial_sort(first, last, last) = make_heap + sort_heap, bits/stl_heap.h) ----
SSE_FN void sse_push_heap(sse_item* a, int hole, int top, sse_item value)
{
	int parent = (hole - 1) / 2;
	while (hole > top && sse_less(a[parent], value)) { a[hole] = a[parent]; hole = parent; parent = (hole - 1) / 2; }
	a[hole] = value;
}
SSE_FN void sse_adjust_heap(sse_item* a, int hole, int len, sse_item value)
{
	const int top = hole;
	int child = hole;
	while (child < (len - 1) / 2) {
		child = 2 * (child + 1);
		if (sse_less(a[child], a[child - 1])) --child;
		a[hole] = a[child]; hole = child;
	}
	if ((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); a[hole] = a[child - 1]; hole = child - 1; }
	sse_push_heap(a, hole, top, value);
}
SSE_FN void sse_heap_sort(sse_item* a, int n) // a[0, n)
{
	if (n >= 2) {
		for (int parent = (n - 2) / 2;; --parent) { sse_adjust_heap(a, parent, n, a[parent]); if (parent == 0) break; }
	}
	for (int last = n; last > 1;) { --last; const sse_item v = a[last]; a[last] = a[0]; sse_adjust_heap(a, 0, last, v); }
}

// stable insertion sort of a leaf range a[first, last)
SSE_FN void sse_insertion_sort(sse_item* a, int first, int last)
{
	for (int i = first + 1; i < last; ++i) {
		const sse_item v = a[i];
		int j = i;
		while (j > first && sse_less(v, a[j - 1])) { a[j] = a[j - 1]; --j; }
		a[j] = v;
	}
}

// one partition step of range [first, last), last - first > 16: returns the cut.  Serial form (the loop as published).
SSE_FN int sse_partition_serial(sse_item* a, int first, int last)
{
	sse_move_median_to_first(a, first, first + 1, first + (last - first) / 2, last - 1);
	const sse_item pivot = a[first];
	int f = first + 1, l = last;
	for (;;) {
		while (sse_less(a[f], pivot)) ++f;
		--l;
		while (sse_less(pivot, a[l])) --l;
		if (!(f < l)) return f;
		const sse_item t = a[f]; a[f] = a[l]; a[l] = t;
		++f;
	}
}

// the whole sort of a[first, last) with `depth` partition levels left, by one thread (used for small ranges and for the heap-sort fallback)
SSE_FN void sse_sort_range_serial(sse_item* a, int first, int last, int depth)
{
	// explicit stack instead of the recursion on the right part: at most one entry per level
	int stF[64], stL[64], stD[64], top = 0;
	for (;;) {
		while (last - first > SSE_THRESHOLD) {
			if (depth == 0) { sse_heap_sort(a + first, last - first); first = last; break; }
			--depth;
			const int cut = sse_partition_serial(a, first, last);
			stF[top] = cut; stL[top] = last; stD[top] = depth; ++top;
			last = cut;
		}
		if (last - first > 1) sse_insertion_sort(a, first, last);
		if (!top) return;
		--top; first = stF[top]; last = stL[top]; depth = stD[top];
	}
}

SSE_FN void sse_sort_serial(sse_item* a, int n)
{
	if (n > 1) sse_sort_range_serial(a, 0, n, sse_lg(static_cast<unsigned int>(n)) * 2);
}

} // namespace cvb

// ---- the closed form of the partition step (what the warp-cooperative device code evaluates with ballots); serial statement for the CPU check ----
namespace cvb {
// Ls / Rs: scratch for last - first entries each.  Returns the cut; leaves a[first, last) exactly as sse_partition_serial does.
SSE_FN int sse_partition_closed_form(sse_item* a, int first, int last, unsigned int* Ls, unsigned int* Rs)
{
	sse_move_median_to_first(a, first, first + 1, first + (last - first) / 2, last - 1);
	const sse_item pivot = a[first];
	int nL = 0, nR = 0;
	for (int i = first + 1; i < last; ++i) if (!sse_less(a[i], pivot)) Ls[nL++] = static_cast<unsigned int>(i);
	for (int i = last - 1; i > first; --i) if (!sse_less(pivot, a[i])) Rs[nR++] = static_cast<unsigned int>(i);
	int K = 0;
	while (K < nL && K < nR && Ls[K] < Rs[K]) ++K;
	for (int k = 0; k < K; ++k) { const sse_item t = a[Ls[k]]; a[Ls[k]] = a[Rs[k]]; a[Rs[k]] = t; }
	const unsigned int cl = (K < nL) ? Ls[K] : 0xffffffffu;
	const unsigned int cr = (K > 0) ? Rs[K - 1] : static_cast<unsigned int>(last);
	return static_cast<int>(cl < cr ? cl : cr);
}

// The same step as the CTA-wide device code builds it (ksort_block_partition in hough_kht.cu): `chunks` workers scan contiguous chunks of first+1 .. last-1 whose size is a
// multiple of 32; L is the chunks' ascending lists in chunk order, R the chunks' descending lists in REVERSE chunk order; K is the NUMBER of k < min(nL, nR) with
// L[k] < R[k] (L ascends and R descends, so the predicate is monotone and the count equals the first failing index).  Serial statement for the CPU check.
SSE_FN int sse_partition_closed_form_chunked(sse_item* a, int first, int last, unsigned int* Ls, unsigned int* Rs, int chunks)
{
	sse_move_median_to_first(a, first, first + 1, first + (last - first) / 2, last - 1);
	const sse_item pivot = a[first];
	const int n = last - first - 1;
	const int per = ((n + chunks * 32 - 1) / (chunks * 32)) * 32;
	int nL = 0, nR = 0;
	for (int w = 0; w < chunks; ++w) { // L: chunk order
		const int cb = (first + 1 + w * per < last) ? first + 1 + w * per : last, ce = (cb + per < last) ? cb + per : last;
		for (int i = cb; i < ce; ++i) if (!sse_less(a[i], pivot)) Ls[nL++] = static_cast<unsigned int>(i);
	}
	for (int w = chunks - 1; w >= 0; --w) { // R: reverse chunk order, descending inside a chunk
		const int cb = (first + 1 + w * per < last) ? first + 1 + w * per : last, ce = (cb + per < last) ? cb + per : last;
		for (int i = ce - 1; i >= cb; --i) if (!sse_less(pivot, a[i])) Rs[nR++] = static_cast<unsigned int>(i);
	}
	const int m = nL < nR ? nL : nR;
	int K = 0;
	for (int k = 0; k < m; ++k) K += (Ls[k] < Rs[k]) ? 1 : 0;
	for (int k = 0; k < K; ++k) { const sse_item t = a[Ls[k]]; a[Ls[k]] = a[Rs[k]]; a[Rs[k]] = t; }
	const unsigned int cl = (K < nL) ? Ls[K] : 0xffffffffu;
	const unsigned int cr = (K > 0) ? Rs[K - 1] : static_cast<unsigned int>(last);
	return static_cast<int>(cl < cr ? cl : cr);
}
} // namespace cvb
