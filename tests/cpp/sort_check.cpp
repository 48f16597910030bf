// Pins compv_b200/csrc/std_sort_emu.cuh (the device-side replacement of the Hough detectors' host std::sort) on libstdc++'s std::sort itself: same input
// order in, same permutation out, ties included.  TEST INFRASTRUCTURE, no GPU; run by tests/test_kht_walk_cpu.py.
#include "../../compv_b200/csrc/std_sort_emu.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace cvb;

static unsigned int lcg(unsigned int& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// the recursion tree evaluated breadth first with the closed-form partition (any order is allowed: ranges are disjoint)
static void sort_closed_form(std::vector<sse_item>& a, int depth0, bool reverseOrder, int bigChunks = 0)
{
	struct R { int f, l, d; };
	std::vector<R> work, next;
	const int n = static_cast<int>(a.size());
	if (n < 2) return;
	std::vector<unsigned int> Ls(n), Rs(n);
	work.push_back({ 0, n, depth0 });
	while (!work.empty()) {
		next.clear();
		if (reverseOrder) std::reverse(work.begin(), work.end());
		for (const R& r : work) {
			if (r.l - r.f <= SSE_THRESHOLD) { sse_insertion_sort(a.data(), r.f, r.l); continue; }
			if (r.d == 0) { sse_heap_sort(a.data() + r.f, r.l - r.f); continue; }
			// ranges above 1024 elements: the CTA-wide formulation (8 chunks), as the device code does (KSORT_BIG)
			const int cut = (bigChunks && r.l - r.f > 1024) ? sse_partition_closed_form_chunked(a.data(), r.f, r.l, Ls.data() + r.f, Rs.data() + r.f, bigChunks)
			                                                  : sse_partition_closed_form(a.data(), r.f, r.l, Ls.data() + r.f, Rs.data() + r.f);
			next.push_back({ r.f, cut, r.d - 1 });
			next.push_back({ cut, r.l, r.d - 1 });
		}
		work.swap(next);
	}
}

int main(int argc, char** argv)
{
	const int cases = argc > 1 ? atoi(argv[1]) : 400;
	unsigned int seed = 99;
	int bad = 0;
	auto comp = [](sse_item x, sse_item y) { return sse_less(x, y); };
	for (int c = 0; c < cases; ++c) {
		static const int sizes[] = { 0, 1, 2, 3, 15, 16, 17, 18, 31, 33, 64, 100, 257, 1000, 4097, 5700, 20000, 60000 };
		const int n = sizes[c % 18];
		const int kind = (c / 18) % 6;
		std::vector<sse_item> in(n);
		for (int i = 0; i < n; ++i) {
			unsigned int key;
			switch (kind) {
			case 0: key = lcg(seed) % 7; break;                       // massive ties
			case 1: key = lcg(seed) % 300 + 100; break;               // vote-count like
			case 2: key = static_cast<unsigned int>(i); break;        // already "ascending" = worst order for a descending sort
			case 3: key = static_cast<unsigned int>(n - i); break;    // already sorted
			case 4: key = (i & 1) ? 5u : static_cast<unsigned int>(i % 50); break; // organ-pipe-ish with ties
			default: key = lcg(seed); break;                          // (almost) distinct
			}
			in[i] = (static_cast<sse_item>(key) << 32) | static_cast<unsigned int>(i);
		}
		std::vector<sse_item> want = in; std::sort(want.begin(), want.end(), comp);
		std::vector<sse_item> g1 = in; sse_sort_serial(g1.data(), n);
		std::vector<sse_item> g2 = in; sort_closed_form(g2, n > 1 ? sse_lg(n) * 2 : 0, false);
		std::vector<sse_item> g3 = in; sort_closed_form(g3, n > 1 ? sse_lg(n) * 2 : 0, true);
		std::vector<sse_item> g4 = in; sort_closed_form(g4, n > 1 ? sse_lg(n) * 2 : 0, false, 8);
		std::vector<sse_item> g5 = in; sort_closed_form(g5, n > 1 ? sse_lg(n) * 2 : 0, true, 3);
		if (g1 != want || g2 != want || g3 != want || g4 != want || g5 != want) { ++bad; fprintf(stderr, "MISMATCH std::sort n=%d kind=%d (%d %d %d %d %d)\n", n, kind, g1 != want, g2 != want, g3 != want, g4 != want, g5 != want); }
		// shallow depth limits force the heap-sort fallback: compare with libstdc++'s own loop run with the same limit
		for (int d = 0; d <= 3 && n > 1; ++d) {
			std::vector<sse_item> w2 = in;
			std::__introsort_loop(w2.begin(), w2.end(), d, __gnu_cxx::__ops::__iter_comp_iter(comp));
			std::__final_insertion_sort(w2.begin(), w2.end(), __gnu_cxx::__ops::__iter_comp_iter(comp));
			std::vector<sse_item> h1 = in; sse_sort_range_serial(h1.data(), 0, n, d);
			std::vector<sse_item> h2 = in; sort_closed_form(h2, d, false);
			std::vector<sse_item> h3 = in; sort_closed_form(h3, d, false, 8);
			if (h3 != w2) { ++bad; fprintf(stderr, "MISMATCH depth-limited chunked n=%d kind=%d d=%d\n", n, kind, d); }
			if (h1 != w2 || h2 != w2) { ++bad; fprintf(stderr, "MISMATCH depth-limited n=%d kind=%d d=%d (%d %d)\n", n, kind, d, h1 != w2, h2 != w2); }
		}
	}
	printf("sort_check: %d cases, %d mismatches\n", cases, bad);
	return bad ? 1 : 0;
}
