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
//   * REV = the bitmap words are bit-reversed (bit 31 = leftmost column) and the TOP row goes to the HIGH field: the priority encoder is then a single
//     find-leading-one (FLO) instead of bit-reverse + FLO.
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

// bit of column j (0..31) inside a bitmap word
template <bool REV> KW_FN unsigned int kw_colbit(int j) { return REV ? (0x80000000u >> j) : (1u << j); }
// column (0..31) of the first pixel, in raster order, of a non-zero word
template <bool REV> KW_FN int kw_first_col(unsigned int w) { return REV ? (31 - kw_highest(w)) : kw_lowest(w); }

// BF = the row rotation of a vertical move is done with selects instead of branches
template <bool REV, bool BF>
struct KhtWalker {
	unsigned int r0, r1, r2, r3, r4; // rows y-2 .. y+2, 32 columns from padded column c0 (window index i <-> bit i, or bit 31 - i when REV)
	int qa, qb, qc;                  // q - 1, q - 9, q - 17 where q = bit of the current pixel, kept in [1, 30]: the three rotation counts
	int ro;                          // word offset of padded word 0 of row y from `base`
	int w0, s;                       // c0 = 32 * w0 + s
	unsigned int xy;                 // x | y << 16 (image coordinates): the value that is stored as the position

	KW_FN unsigned int load_row(const unsigned int* base, int off) const
	{
		const unsigned int* row = base + (off + w0);
		return REV ? kw_funnel_l(row[1], row[0], s) : kw_funnel_r(row[0], row[1], s);
	}
	// (re)load the window with the current pixel in the middle; `base` = padded word 0 of image row 0
	KW_FN void centre(const unsigned int* base, int WW)
	{
		const int X = static_cast<int>(xy & 0xffffu) + 32, y = static_cast<int>(xy >> 16);
		const int c0 = X - 16;
		w0 = c0 >> 5; s = c0 & 31;
		ro = y * WW;
		r0 = load_row(base, ro - 2 * WW); r1 = load_row(base, ro - WW); r2 = load_row(base, ro); r3 = load_row(base, ro + WW); r4 = load_row(base, ro + 2 * WW);
		const int q = REV ? 15 : 16;
		qa = q - 1; qb = q - 9; qc = q - 17;
	}
	// Erase the current pixel (window + memory), then Algorithm 6: move to the first remaining neighbour in the order TL, T, TR, L, R, BL, B, BR.
	// false when there is none.  The centre bit is masked out of the neighbourhood instead of waiting for the erase, so the erase is off the chain.
	KW_FN bool step(unsigned int* base, int WW)
	{
		unsigned int m;
		if (REV) m = (kw_rotr(r1, qc) & 0x70000u) | (kw_rotr(r2, qb) & 0x500u) | (kw_rotr(r3, qa) & 7u);
		else m = (kw_rotr(r1, qa) & 7u) | (kw_rotr(r2, qb) & 0x500u) | (kw_rotr(r3, qc) & 0x70000u);
		r2 &= ~(2u << qa);
		{
			const int X = static_cast<int>(xy & 0xffffu) + 32;
			unsigned int* wp = base + (ro + (X >> 5));
			*wp &= ~kw_colbit<REV>(X & 31);
		}
		if (!m) return false;
		const int k = REV ? kw_highest(m) : kw_lowest(m);
		const int kx = k & 3, ky = k >> 3;   // 0..2 each
		qa += kx - 1; qb += kx - 1; qc += kx - 1;
		// x moves with the bit index (or against it when REV); y moves down when the field is the bottom one
		xy += REV ? static_cast<unsigned int>(65537 - kx - (ky << 16)) : static_cast<unsigned int>(kx + (ky << 16) - 65537);
		if (static_cast<unsigned int>(qa) > 29u) { centre(base, WW); return true; }
		const int up = REV ? 2 : 0;
		if (BF) {
			const bool isUp = (ky == up), isDn = (ky == 2 - up);
			const int dro = isUp ? -WW : (isDn ? WW : 0);
			ro += dro;
			const unsigned int n0 = r0, n1 = r1, n2 = r2, n3 = r3, n4 = r4;
			unsigned int far = 0;
			if (ky != 1) far = load_row(base, ro + 2 * dro);
			r0 = isUp ? far : (isDn ? n1 : n0);
			r1 = isUp ? n0 : (isDn ? n2 : n1);
			r2 = isUp ? n1 : (isDn ? n3 : n2);
			r3 = isUp ? n2 : (isDn ? n4 : n3);
			r4 = isUp ? n3 : (isDn ? far : n4);
		}
		else {
			if (ky == up) { ro -= WW; r4 = r3; r3 = r2; r2 = r1; r1 = r0; r0 = load_row(base, ro - 2 * WW); }
			else if (ky == 2 - up) { ro += WW; r0 = r1; r1 = r2; r2 = r3; r3 = r4; r4 = load_row(base, ro + 2 * WW); }
		}
		return true;
	}
};

// Algorithm 5 for one seed: appends the string's positions at `out` (first walk in walk order -- the caller reverses out[0, rev) afterwards, the
// reference's std::reverse, houghkht.cxx:752-755 -- then the second walk) and returns the number of positions; *rev = length of the first walk.
template <bool REV, bool BF>
KW_FN unsigned int kht_link_string(unsigned int* base, int WW, unsigned int seedXY, unsigned int* out, unsigned int* rev)
{
	unsigned int n = 0;
	KhtWalker<REV, BF> wk;
	wk.xy = seedXY;
	wk.centre(base, WW);
	do {
		out[n++] = wk.xy;
	} while (wk.step(base, WW));
	*rev = n;
	wk.xy = seedXY;
	wk.centre(base, WW);
	if (wk.step(base, WW)) { // the seed is erased already: this only looks for what the first walk left around it
		do {
			out[n++] = wk.xy;
		} while (wk.step(base, WW));
	}
	return n;
}

} // namespace cvb
