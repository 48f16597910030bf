// CPU replay of the KHT string walker (compv_b200/csrc/kht_walk.cuh, the code the CUDA linking kernel runs on one lane) against a byte-map restatement of the
// reference's linking procedure (core/features/hough/compv_core_feature_houghkht.cxx:544-760: raster scan over interior seeds, Algorithm 5 with its two walks
// and the reversal of the first, Algorithm 6's neighbour order).  TEST INFRASTRUCTURE: no GPU involved; run by tests/test_kht_walk_cpu.py.
#include "../../compv_b200/csrc/kht_walk.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace cvb;

struct Strings { std::vector<unsigned int> poss; std::vector<unsigned int> begin, end; };

// ---- byte-map restatement ----
static bool next_a6(std::vector<unsigned char>& e, int W, int H, int& x, int& y)
{
	static const int dx[8] = { -1, 0, 1, -1, 1, -1, 0, 1 }, dy[8] = { -1, -1, -1, 0, 0, 1, 1, 1 };
	for (int k = 0; k < 8; ++k) {
		const int nx = x + dx[k], ny = y + dy[k];
		if (nx < 0 || ny < 0 || nx >= W || ny >= H) continue;
		if (e[static_cast<size_t>(ny) * W + nx]) { x = nx; y = ny; return true; }
	}
	return false;
}

static void link_bytes(std::vector<unsigned char> e, int W, int H, unsigned int minSize, Strings& out)
{
	for (int yr = 1; yr < H - 1; ++yr) for (int xr = 1; xr < W - 1; ++xr) {
		if (!e[static_cast<size_t>(yr) * W + xr]) continue;
		const size_t b = out.poss.size();
		int x = xr, y = yr;
		do { out.poss.push_back(static_cast<unsigned int>(x) | (static_cast<unsigned int>(y) << 16)); e[static_cast<size_t>(y) * W + x] = 0; } while (next_a6(e, W, H, x, y));
		const size_t r = out.poss.size();
		x = xr; y = yr;
		if (next_a6(e, W, H, x, y)) {
			do { out.poss.push_back(static_cast<unsigned int>(x) | (static_cast<unsigned int>(y) << 16)); e[static_cast<size_t>(y) * W + x] = 0; } while (next_a6(e, W, H, x, y));
		}
		const size_t n = out.poss.size();
		if (n - b >= minSize) { std::reverse(out.poss.begin() + b, out.poss.begin() + r); out.begin.push_back(static_cast<unsigned int>(b)); out.end.push_back(static_cast<unsigned int>(n)); }
		else out.poss.resize(b);
	}
}

// ---- the walker on the padded bitmap, seeds found by a plain raster scan of the bitmap words ----
static void link_bits(const std::vector<unsigned char>& e, int W, int H, unsigned int minSize, Strings& out)
{
	const int WW = (W + 31) / 32 + 2;
	std::vector<unsigned int> bits(static_cast<size_t>(H + 2 * KHT_PADR) * WW, 0u);
	unsigned int* base = bits.data() + static_cast<size_t>(KHT_PADR) * WW;
	size_t edges = 0;
	for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) if (e[static_cast<size_t>(y) * W + x]) { base[y * WW + 1 + (x >> 5)] |= kw_colbit(x & 31); ++edges; }
	out.poss.assign(edges + 1, 0u);
	unsigned int nPos = 0;
	const int lastWord = (W - 1) >> 5;
	for (int y = 1; y < H - 1; ++y) {
		for (int wi = 0; wi <= lastWord; ++wi) {
			for (;;) {
				unsigned int w = base[y * WW + 1 + wi];
				if (wi == 0) w &= ~kw_colbit(0);
				if (wi == lastWord) w &= ~kw_colbit((W - 1) & 31);
				if (!w) break;
				const int xr = wi * 32 + kw_first_col(w);
				unsigned int rev = 0;
				const unsigned int n = kht_link_string(base, WW, static_cast<unsigned int>(xr) | (static_cast<unsigned int>(y) << 16), out.poss.data() + nPos, &rev);
				if (n >= minSize) {
					std::reverse(out.poss.begin() + nPos, out.poss.begin() + nPos + rev); // kht_reverse_kernel on the device
					out.begin.push_back(nPos); out.end.push_back(nPos + n);
					nPos += n;
				}
			}
		}
	}
	out.poss.resize(nPos);
	// whatever is left in the bitmap must be what the byte-map version leaves: checked through the strings only (every seedable pixel was consumed)
}


// ---- the same pass through the per-lane state machine (KhtLane: what kht_link_lanes_kernel runs, 32 frames per warp) ----
static void link_lane(const std::vector<unsigned char>& e, int W, int H, unsigned int minSize, Strings& out)
{
	const int WW = (W + 31) / 32 + 2;
	std::vector<unsigned int> bits(static_cast<size_t>(H + 2 * KHT_PADR) * WW, 0u);
	unsigned int* base = bits.data() + static_cast<size_t>(KHT_PADR) * WW;
	size_t edges = 0;
	for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) if (e[static_cast<size_t>(y) * W + x]) { base[y * WW + 1 + (x >> 5)] |= kw_colbit(x & 31); ++edges; }
	out.poss.assign(edges + 1, 0u);
	std::vector<unsigned long long> strs(edges + 1, 0ull);
	std::vector<unsigned int> revs(edges + 1, 0u);
	KhtLane L;
	L.start(H);
	while (L.phase != 3) L.iterate(base, WW, W, H, minSize, out.poss.data(), strs.data(), revs.data());
	for (unsigned int s = 0; s < L.nStr; ++s) {
		const unsigned int b = static_cast<unsigned int>(strs[s]), en = static_cast<unsigned int>(strs[s] >> 32);
		std::reverse(out.poss.begin() + b, out.poss.begin() + b + revs[s]); // kht_reverse_kernel on the device
		out.begin.push_back(b); out.end.push_back(en);
	}
	out.poss.resize(L.nPos);
}

static unsigned int lcg(unsigned int& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

static bool same(const Strings& a, const Strings& b) { return a.poss == b.poss && a.begin == b.begin && a.end == b.end; }

int main(int argc, char** argv)
{
	const int cases = argc > 1 ? atoi(argv[1]) : 200;
	unsigned int seed = 2024;
	int bad = 0;
	size_t totalStrings = 0, totalPos = 0;
	for (int c = 0; c < cases; ++c) {
		static const int sizes[][2] = { { 3, 3 }, { 5, 4 }, { 31, 9 }, { 32, 17 }, { 33, 20 }, { 64, 48 }, { 65, 33 }, { 97, 61 }, { 160, 120 }, { 321, 77 }, { 640, 200 } };
		const int W = sizes[c % 11][0], H = sizes[c % 11][1];
		std::vector<unsigned char> e(static_cast<size_t>(W) * H, 0);
		const int kind = (c / 11) % 5;
		const unsigned int density = 5 + lcg(seed) % 60;      // percent
		if (kind == 0) { for (auto& v : e) v = (lcg(seed) % 100 < density) ? 255 : 0; }                   // noise of any density: junction-heavy
		else if (kind == 1) { for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) e[static_cast<size_t>(y) * W + x] = ((x % 7 == 3) || (y % 5 == 2)) ? 255 : 0; } // grid: one giant component, border pixels
		else if (kind == 2) { for (int i = 0; i < 40; ++i) { int x = lcg(seed) % W, y = lcg(seed) % H; const int dxs = static_cast<int>(lcg(seed) % 3) - 1, dys = static_cast<int>(lcg(seed) % 3) - 1;
			for (int s = 0; s < 300 && x >= 0 && y >= 0 && x < W && y < H; ++s) { e[static_cast<size_t>(y) * W + x] = 255; x += dxs; y += dys; if (lcg(seed) % 16 == 0) x += 1; } } } // long thin strokes in all 8 directions
		else if (kind == 3) { for (auto& v : e) v = 255; }                                               // everything set: walks sweep whole rows and hit all four borders
		else { for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) e[static_cast<size_t>(y) * W + x] = (((x + y) % 9 == 0) || ((x - y + 1000) % 11 == 0) || lcg(seed) % 100 < 3) ? 255 : 0; } // diagonals + specks
		const unsigned int minSize = (c % 3 == 0) ? 2 : 10;
		Strings want; link_bytes(e, W, H, minSize, want);
		Strings got;
		link_bits(e, W, H, minSize, got);
		if (!same(want, got)) { ++bad; fprintf(stderr, "MISMATCH case %d: %dx%d kind %d minSize %u: strings %zu vs %zu\n", c, W, H, kind, minSize, want.begin.size(), got.begin.size()); }
		Strings lane;
		link_lane(e, W, H, minSize, lane);
		if (!same(want, lane)) { ++bad; fprintf(stderr, "MISMATCH (lane state machine) case %d: %dx%d kind %d minSize %u: strings %zu vs %zu\n", c, W, H, kind, minSize, want.begin.size(), lane.begin.size()); }
		totalStrings += want.begin.size(); totalPos += want.poss.size();
	}
	printf("link_check: %d cases, %zu strings, %zu positions, %d mismatches\n", cases, totalStrings, totalPos, bad);
	return bad ? 1 : 0;
}
