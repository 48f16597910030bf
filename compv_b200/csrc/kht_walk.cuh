// The string walker of the KHT linking stage: Algorithm 5 / Algorithm 6 of the reference (core/features/hough/compv_core_feature_houghkht.cxx:666-760)
// on a 1 bit/px bitmap.  Written once for the device (hough_kht.cu, one lane per frame walks) and for the host (tests/cpp/link_check.cpp replays it
// against a byte-map restatement of the reference's procedure without a GPU): plain C++ plus four bit intrinsics.
//
// The walk is one dependent instruction chain, so the design goal is chain LENGTH, not instruction count:
//   * the bitmap has PADR zero rows above/below and one zero word left/right of the pixels: no border case anywhere;
//   * five bitmap rows (y-2 .. y+2) x 32 columns live in registers, the window may start at ANY column (two aligned words + one funnel shift), so a
//     re-centred pixel always sits in the middle; rows y-2 / y+2 are there so that the row a vertical move needs is already in a register and the
//     load of the row after it has a whole step to arrive;
//   * the 8-neighbourhood is gathered with three rotates + three ANDs + one OR into bit fields 8 apart (row | row << 8 | row << 16) so that ONE
//     find-first-set gives the reference's priority order (TL, T, TR, L, R, BL, B, BR) and the move decodes as k & 3, k >> 3;
//   * the bitmap words are bit-reversed (bit 31 = leftmost column) and the TOP row goes to the HIGH field: the priority encoder is then a single
//     find-leading-one (FLO) instead of bit-reverse + FLO;
//   * every load has a whole step between its issue and its first use (the machine is in-order: a consumer placed right behind its load stalls the chain
//     for the full L1 / L2 latency, which is what ncu showed for the first version of this walker).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KW_FN __host__ __device__ __forceinline__
#else
#define KW_FN inline
#endif

namespace cvb {

#define KHT_PADR 2 // zero rows above and below the image rows of a frame's bitmap

KW_FN unsigned int kw_rotr(unsigned int v, int n) // rotate right by n mod 32
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_r(v, v, n);
#else
	n &= 31;
	return n ? ((v >> n) | (v << (32 - n))) : v;
#endif
}
KW_FN unsigned int kw_funnel_r(unsigned int lo, unsigned int hi, int s) // low word of (hi:lo) >> s, s in [0, 31]
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_r(lo, hi, s);
#else
	return static_cast<unsigned int>(((static_cast<unsigned long long>(hi) << 32) | lo) >> (s & 31));
#endif
}
KW_FN unsigned int kw_funnel_l(unsigned int lo, unsigned int hi, int s) // high word of (hi:lo) << s, s in [0, 31]
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_l(lo, hi, s);
#else
	return static_cast<unsigned int>((((static_cast<unsigned long long>(hi) << 32) | lo) << (s & 31)) >> 32);
#endif
}
KW_FN int kw_lowest(unsigned int v) // index of the lowest set bit, v != 0
{
#if defined(__CUDA_ARCH__)
	return __ffs(v) - 1;
#else
	return __builtin_ctz(v);
#endif
}
KW_FN int kw_highest(unsigned int v) // index of the highest set bit, v != 0
{
#if defined(__CUDA_ARCH__)
	int r;
	asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v)); // FLO.U32: one instruction (31 - __clz(v) compiles to three)
	return r;
#else
	return 31 - __builtin_clz(v);
#endif
}

// Bitmap words are stored bit-reversed: bit 31 = leftmost of the word's 32 columns, so that "first in raster order" and "highest priority" are both
// "highest set bit" (one FLO instruction).
KW_FN unsigned int kw_colbit(int j) { return 0x80000000u >> j; }              // bit of column j (0..31) inside a bitmap word
KW_FN int kw_first_col(unsigned int w) { return 31 - kw_highest(w); }         // column of the first pixel, in raster order, of a non-zero word

KW_FN void kw_prefetch_l1(const void* p)
{
#if defined(__CUDA_ARCH__)
	asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#else
	(void)p;
#endif
}

#define KHT_WALK_PREFETCH_ROWS 6 // a vertical move prefetches the row this many rows further on into L1: the far-row load two steps later then hits

// The walker's state.  Window index i (0..31) <-> bit 31 - i; the current pixel sits at bit q = c + 1, c kept in [6, 23] (window index 7..24) so that the
// aligned 8-column byte that holds the pixel always lies inside the window: the memory erase is then ONE BYTE STORE taken from the window row, with no load.
// (ncu on the first two versions: a read-modify-write of the bitmap word costs an L2 round trip per step, because the previous step's store has just
// invalidated that line in L1 -- every version that loaded before storing ran at the same ~300 cycles per pixel whatever else it did.)
// The three rows of the 8-neighbourhood are held PRE-ROTATED (T = row(y-1) rotated by 16, M = row(y) rotated left by 8, B = row(y+1) as is) so that one
// rotation count c brings the three 3-bit fields to bits 16-18 / 8-10 / 0-2: top row in the high field, left neighbour in the high bit of each field.
struct KhtWalker {
	unsigned int T, M, B;            // rows y-1, y, y+1 (pre-rotated, see above)
	unsigned int Fu, Fd;             // rows y-2, y+2 (not rotated)
	unsigned int plo, phi; int pend; // a far row still in flight: raw words, 1 = it is Fu, 2 = it is Fd, 0 = none.  Consumed one step later at the earliest,
	                                 // so that the load's latency is off the dependency chain
	int c;                           // rotation count = bit of the current pixel - 1
	int ro;                          // word offset of padded word 0 of row y from `base`
	int w0, s;                       // the window starts at padded column 32 * w0 + s
	unsigned int xy;                 // x | y << 16 (image coordinates): the value that is stored as the position

	KW_FN unsigned int load_row(const unsigned int* base, int off) const
	{
		const unsigned int* row = base + (off + w0);
		return kw_funnel_l(row[1], row[0], s);
	}
	// (re)load the window with the current pixel at window index ip (7..24); `base` = padded word 0 of image row 0
	KW_FN void centre(const unsigned int* base, int WW, int ip)
	{
		const int X = static_cast<int>(xy & 0xffffu) + 32, y = static_cast<int>(xy >> 16);
		const int c0 = X - ip;
		w0 = c0 >> 5; s = c0 & 31;
		ro = y * WW;
		Fu = load_row(base, ro - 2 * WW);
		T = kw_rotr(load_row(base, ro - WW), 16);
		M = kw_rotr(load_row(base, ro), 24);
		B = load_row(base, ro + WW);
		Fd = load_row(base, ro + 2 * WW);
		pend = 0; // plo / phi are deliberately left alone: writing them here would have to wait for a far-row load that may still be in flight
		c = 30 - ip;
	}
	// Erase the current pixel (window + memory), then Algorithm 6: move to the first remaining neighbour in the order TL, T, TR, L, R, BL, B, BR.
	// false when there is none.
	KW_FN bool step(unsigned int* base, int WW)
	{
		const unsigned int m = (kw_rotr(T, c) & 0x70000u) | (kw_rotr(M, c) & 0x500u) | (kw_rotr(B, c) & 7u); // the centre bit is masked out: no need to wait for the erase
		M &= ~kw_rotr(0x200u, 32 - c); // bit c + 1 of row y sits at bit c + 9 of M
		{
			// the aligned byte (8 columns) that holds the pixel, taken from the erased window row: column X & ~7 is its bit 7.  Bitmap words are bit-reversed,
			// so inside a little-endian word the byte of columns 8k..8k+7 is byte 3 - k.
			const int X = static_cast<int>(xy & 0xffffu) + 32;
			const unsigned int v = kw_rotr(M, c + (X & 7) + 2);
			reinterpret_cast<unsigned char*>(base + ro)[(X >> 3) ^ 3] = static_cast<unsigned char>(v);
		}
		if (!m) return false;
		const int k = kw_highest(m);
		const int kx = k & 3, ky = k >> 3;   // kx: 2 = left, 1 = same column, 0 = right;  ky: 2 = up, 1 = same row, 0 = down
		c += kx - 1;
		xy += static_cast<unsigned int>(65537 - kx - (ky << 16));
		if (static_cast<unsigned int>(c - 6) > 17u) { centre(base, WW, c < 6 ? 9 : 22); return true; } // left the window sideways: re-enter it with room ahead in the direction of travel
		if (ky != 1) {
			// the far row that was requested by the previous vertical move (at least one step ago) is taken out of its raw words now
			const unsigned int v = kw_funnel_l(phi, plo, s);
			if (pend == 1) Fu = v;
			if (pend == 2) Fd = v;
			int far;
			if (ky == 2) { ro -= WW; Fd = B; B = kw_rotr(M, 8); M = kw_rotr(T, 8); T = kw_rotr(Fu, 16); far = ro - 2 * WW; pend = 1; kw_prefetch_l1(base + (ro - KHT_WALK_PREFETCH_ROWS * WW + w0)); }
			else { ro += WW; Fu = kw_rotr(T, 16); T = kw_rotr(M, 24); M = kw_rotr(B, 24); B = Fd; far = ro + 2 * WW; pend = 2; kw_prefetch_l1(base + (ro + KHT_WALK_PREFETCH_ROWS * WW + w0)); }
			const unsigned int* row = base + (far + w0);
			plo = row[0]; phi = row[1];
		}
		return true;
	}
};

// Algorithm 5 for one seed: appends the string's positions at `out` (first walk in walk order -- the caller reverses out[0, rev) afterwards, the
// reference's std::reverse, houghkht.cxx:752-755 -- then the second walk) and returns the number of positions; *rev = length of the first walk.
KW_FN unsigned int kht_link_string(unsigned int* base, int WW, unsigned int seedXY, unsigned int* out, unsigned int* rev)
{
	int n = 0; // a signed 32-bit index: one IMAD.WIDE per address
	KhtWalker wk;
	wk.plo = 0; wk.phi = 0;
	wk.xy = seedXY;
	wk.centre(base, WW, 12); // a seed is the raster-first pixel left: nothing above it or to its left, the walk starts rightwards or downwards
	do {
		out[n++] = wk.xy;
	} while (wk.step(base, WW));
	*rev = static_cast<unsigned int>(n);
	wk.xy = seedXY;
	wk.centre(base, WW, 16);
	if (wk.step(base, WW)) { // the seed is erased already: this only looks for what the first walk left around it
		do {
			out[n++] = wk.xy;
		} while (wk.step(base, WW));
	}
	return static_cast<unsigned int>(n);
}

// ---- the whole linking pass of ONE frame as a state machine (one lane per frame: kht_link_lanes_kernel runs 32 frames per warp) ----
// The one-warp-per-frame kernel spends 31 of its 32 lanes idle while lane 0 walks, so with many frames in flight it is bound by instruction issue (5.4 M
// warp instructions per 1080p frame).  Here every lane owns a frame and all of them go round ONE loop whose body is "scan two bitmap words for the next seed" or
// "take one step of the walk", so that one issued instruction serves up to 32 frames.  Same scan order, same walker, same output as kht_link_string driven by a raster scan.
struct KhtLane {
	KhtWalker wk;
	unsigned int seedXY;
	unsigned int n, rev;       // positions of the current string so far, length of its first walk
	unsigned int nPos, nStr;   // positions / strings kept so far
	int sy, sw;                // scan position: image row, (even) word of the row
	int phase;                 // 0 = looking for a seed, 1 = first walk, 2 = second walk, 3 = frame finished
	int emit;                  // store the current position before the next step (0 only for the step that opens the second walk)

	KW_FN void start(int H)
	{
		wk.plo = 0; wk.phi = 0; wk.pend = 0; wk.T = wk.M = wk.B = wk.Fu = wk.Fd = 0; wk.c = 0; wk.ro = 0; wk.w0 = 0; wk.s = 0; wk.xy = 0;
		seedXY = 0; n = 0; rev = 0; nPos = 0; nStr = 0; sy = 1; sw = 0; emit = 1;
		phase = (H > 2) ? 0 : 3;
	}
	// One round of the loop.  base = padded word 0 of image row 0, poss = the frame's position pool (positions as x | y << 16), strs = its string records
	// (begin | end << 32: a little-endian uint2 {begin, end}), revs = the length of every string's first walk.
	KW_FN void iterate(unsigned int* base, int WW, int W, int H, unsigned int minSize, unsigned int* poss, unsigned long long* strs, unsigned int* revs)
	{
		if (phase == 0) {
			const int lastWord = (W - 1) >> 5;
			const unsigned int* row = base + sy * WW + 1;
			unsigned int a = row[sw], b = row[sw + 1]; // sw + 1 may be the zero pad word right of the row
			// seeds are interior columns only: x in [1, W-2]
			if (sw == 0) a &= ~kw_colbit(0);
			if (sw == lastWord) a &= ~kw_colbit((W - 1) & 31);
			if (sw + 1 == lastWord) b &= ~kw_colbit((W - 1) & 31);
			if (a | b) {
				const int xr = sw * 32 + (a ? kw_first_col(a) : 32 + kw_first_col(b));
				seedXY = static_cast<unsigned int>(xr) | (static_cast<unsigned int>(sy) << 16);
				wk.xy = seedXY;
				wk.centre(base, WW, 12); // a seed is the raster-first pixel left: nothing above it or to its left
				n = 0; emit = 1; phase = 1;
			}
			else {
				sw += 2;
				if (sw > lastWord) { sw = 0; ++sy; if (sy >= H - 1) phase = 3; }
			}
		}
		if (phase == 1 || phase == 2) {
			if (emit) poss[nPos + n++] = wk.xy;
			emit = 1;
			if (!wk.step(base, WW)) {
				if (phase == 1) { // the first walk is over: look around the seed for what it left (the seed itself is erased already)
					rev = n;
					wk.xy = seedXY;
					wk.centre(base, WW, 16);
					emit = 0; phase = 2;
				}
				else {
					if (n >= minSize) { strs[nStr] = static_cast<unsigned long long>(nPos) | (static_cast<unsigned long long>(nPos + n) << 32); revs[nStr] = rev; ++nStr; nPos += n; }
					phase = 0;
				}
			}
		}
	}
};

} // namespace cvb
