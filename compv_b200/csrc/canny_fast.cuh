// Fast path of the Canny front end for the common configuration: Sobel 3x3 with an optional fused 3- or 5-tap Gaussian.
// Same arithmetic contract as edge_front_kernel<3,0> in edges.cu (that generic kernel stays as the fallback for Sobel 5x5,
// 7-tap blurs and strides TMA cannot address), restructured for issue rate:
//   * the (120+8) x (60+2*(rb+2)) u8 input tile lands in shared memory through ONE TMA bulk-tensor copy per CTA
//     (cp.async.bulk.tensor.3d, out-of-image elements zero-filled by the hardware) -- no per-pixel load instructions;
//   * every lane owns 4 consecutive pixels (one 32-bit shared-memory word per row): all shared-memory traffic is 32/64-bit and conflict free;
//   * u8 <-> f32 conversions use the 2^23 magic-number trick (PRMT + FADD / FADD.RZ) instead of the quarter-rate I2F/F2I pipe;
//   * the Gaussian runs on PACKED fp32 pairs (packed.cuh: fma.rn.f32x2 / add.rz.f32x2 = FFMA2 / FADD2 of sm_100a): each half rounds exactly like the scalar
//     instruction, so the reference's FMA chain is reproduced bit for bit with half the issue slots (two rows per step horizontally, two columns vertically);
//   * the Sobel stage works on two 16-bit pixels per 32-bit integer instruction (values biased so that no borrow crosses the halves);
//   * gx/gy never reach shared memory: the gradient stage stores g (11 bits for Sobel 3x3); NMS candidates (g > tLow) are found by one ballot per pixel slot,
//     kept one mask per lane and expanded once per warp into a queue that all 32 lanes work off (direction + the two neighbours only for candidates).
// Tile geometry: a warp's 32 lanes x 4 px = 128 columns [x0-4, x0+124); blurred columns valid on [x0-2, x0+122), g on [x0-1, x0+121),
// output on [x0, x0+120) (lanes 1..30).  120 divides 1920 and 3840, 60 divides 1080, 2160 and 480.
#pragma once

#include "tma.cuh"
#include "packed.cuh"

namespace cvb {

constexpr int CF_TW = 120, CF_TH = 60, CF_THREADS = 256, CF_WARPS = 8;
constexpr int CF_ROWW = 32;           // words per row of the intermediate tiles (128 bytes)
constexpr int CF_INW = 36;            // words per row of the TMA-staged input tile: the box is 144 bytes wide because the innermost TMA
                                      // coordinate must be a multiple of 16 bytes: origin = (x0-4) & ~15, my word = woff + lane, woff in {1, 3}
constexpr int CF_PAD = 4;             // pad words before each array: lane-1 / lane+1 over-reads stay inside the allocation

template <int BKS>
struct CFGeom {
	static constexpr int RB = BKS >> 1;
	static constexpr int IN_ROWS = CF_TH + 2 * (RB + 2);   // image rows y0-RB-2 .. y0+TH+RB+1
	static constexpr int B_ROWS = CF_TH + 4;                // y0-2 .. y0+TH+1
	static constexpr int G_ROWS = CF_TH + 2;                // y0-1 .. y0+TH
	// word offsets inside dynamic shared memory
	static constexpr int OFF_A = CF_PAD;                                  // input tile, later the blurred tile
	static constexpr int OFF_M = OFF_A + IN_ROWS * CF_INW + CF_PAD;       // horizontally blurred tile
	static constexpr int OFF_G = OFF_M + (BKS ? IN_ROWS * CF_ROWW : 0) + CF_PAD; // g|dir (u16 x 128 per row = 64 words)
	static constexpr int OFF_Q = OFF_G + G_ROWS * 64 + CF_PAD;            // per-warp candidate queues of stage S4: 8 warps x (8 rows x 128 px) u16 entries + 8 counters
	static constexpr int Q_WORDS_PER_WARP = 8 * 128 / 2;
	static constexpr int WORDS = OFF_Q + CF_WARPS * Q_WORDS_PER_WARP + CF_WARPS + CF_PAD;
	static constexpr size_t SMEM = WORDS * 4 + 512; // + 128-byte alignment slack, the 32 pad words in front of sA and the mbarrier
};

struct FastParams {
	const uint8_t* in;      // used only by the non-TMA loader
	uint8_t* cls;
	const ushort2* thr;
	int W, H;
	size_t stride, framePitch;
	int tLow, tHigh;
	float k[5];
	int useTma;
	int vecStore;           // cls rows are 4-byte aligned
};

__device__ __forceinline__ float u8_to_f32(unsigned int w, int byteIdx)
{
	// [b, 0, 0, 0x4B] = 2^23 + b as a float; subtracting 2^23 is exact
	return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | byteIdx)) - 8388608.f;
}

// trunc(s) in the low byte of the returned bit pattern (2^23 magic add, round toward zero).  Valid for 0 <= s < 256: the fast path is only
// taken for non-negative taps whose sum is <= 1.003 (a normalised Gaussian), so s <= 255 * 1.003 < 256 and the reference's clamp is a no-op.
__device__ __forceinline__ unsigned int f32_to_u8_bits(float s)
{
	return __float_as_uint(__fadd_rz(s, 8388608.f));
}

__device__ __forceinline__ unsigned int pack4(unsigned int a, unsigned int b, unsigned int c, unsigned int d)
{
	return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}



template <int BKS>
__global__ void __launch_bounds__(CF_THREADS, 3)
canny_front_fast_kernel(const __grid_constant__ CUtensorMap tmap, const FastParams p)
{
	using G = CFGeom<BKS>;
	constexpr int RB = G::RB;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	// 128-byte aligned base (TMA destination alignment); the pad is computed on the shared-window address so that the compiler keeps
	// the pointers in the shared address space (LDS/STS, not generic LD/ST)
	const unsigned int pad = (128u - (static_cast<unsigned int>(__cvta_generic_to_shared(smem_raw)) & 127u)) & 127u;
	unsigned int* base = reinterpret_cast<unsigned int*>(smem_raw + pad);
	// OFF_A = CF_PAD words = 16 bytes into the aligned block: shift so that sA itself is 128-byte aligned
	unsigned int* sA = base + 32;                         // input tile rows, then blurred rows
	unsigned int* sM = sA + (G::OFF_M - G::OFF_A);
	unsigned int* sGw = sA + (G::OFF_G - G::OFF_A);
	uint64_t* bar = reinterpret_cast<uint64_t*>(sA + (G::WORDS - G::OFF_A) + 2);

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int W = p.W, H = p.H;
	const int x0 = blockIdx.x * CF_TW, y0 = blockIdx.y * CF_TH;
	const int frame = blockIdx.z;
	const int xl = x0 - 4 + 4 * lane;                     // first of my 4 columns
	const int yIn0 = y0 - RB - 2;                         // image row of staged row 0
	const int xTma = (x0 - 4) & ~15;                      // 16-byte aligned origin of the staged tile
	const int woff = ((x0 - 4) - xTma) >> 2;              // word of lane 0 inside a staged row

	// ---- S0: stage the input tile ----
	if (p.useTma) {
		if (threadIdx.x == 0) {
			mbar_init(bar, 1);
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			mbar_expect_tx(bar, G::IN_ROWS * CF_INW * 4);
			tma_load_3d(sA, &tmap, bar, xTma, yIn0, frame);
		}
		__syncthreads();
		mbar_wait(bar, 0);
	}
	else {
		const uint8_t* __restrict__ in = p.in + frame * p.framePitch;
		for (int r = warp; r < G::IN_ROWS; r += CF_WARPS) {
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
			sA[r * CF_INW + woff + lane] = w;
		}
		__syncthreads();
	}

	if (BKS) {
		const f32x2 negMagic = pk2(0xCB000000u, 0xCB000000u), magic = pk2(0x4B000000u, 0x4B000000u); // -2^23, 2^23
		f32x2 kk[BKS ? BKS : 1];
#pragma unroll
		for (int k = 0; k < BKS; ++k) kk[k] = pk2(__float_as_uint(p.k[k]), __float_as_uint(p.k[k]));
		// ---- S1: horizontal blur -> sM (same rows as the input tile).  mid(y,x) = 0 outside [RB, W-RB) x [0, H): rows outside the image are staged as zeros
		// and blur to zero by themselves.  Two rows per step: the float pair is (row r, row r+1) of one column, so every FFMA2 serves both rows. ----
		unsigned int colMask = 0;
#pragma unroll
		for (int i = 0; i < 4; ++i) if (xl + i >= RB && xl + i < W - RB) colMask |= 0xffu << (8 * i);
		static_assert(G::IN_ROWS % 2 == 0, "rows are blurred in pairs");
		for (int r = 2 * warp; r < G::IN_ROWS; r += 2 * CF_WARPS) {
			const unsigned int* q0 = &sA[r * CF_INW + woff + lane];
			const unsigned int* q1 = q0 + CF_INW;
			const unsigned int l0 = q0[-1], c0 = q0[0], r0 = q0[1], l1 = q1[-1], c1 = q1[0], r1 = q1[1];
			f32x2 v[4 + 2 * RB];
#pragma unroll
			for (int j = 0; j < RB; ++j) v[j] = u8x2_to_f32x2(l0, 4 - RB + j, l1, 4 - RB + j, negMagic);
#pragma unroll
			for (int j = 0; j < 4; ++j) v[RB + j] = u8x2_to_f32x2(c0, j, c1, j, negMagic);
#pragma unroll
			for (int j = 0; j < RB; ++j) v[RB + 4 + j] = u8x2_to_f32x2(r0, j, r1, j, negMagic);
			unsigned int oa[4], ob[4];
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				f32x2 s = fmul2(v[i], kk[0]); // == fma(v, k, 0)
#pragma unroll
				for (int k = 1; k < BKS; ++k) s = ffma2(v[i + k], kk[k], s);
				unpk2(fadd2_rz(s, magic), oa[i], ob[i]);
			}
			sM[r * CF_ROWW + lane] = pack4(oa[0], oa[1], oa[2], oa[3]) & colMask;
			sM[(r + 1) * CF_ROWW + lane] = pack4(ob[0], ob[1], ob[2], ob[3]) & colMask;
		}
		__syncthreads();

		// ---- S2: vertical blur -> sA (blurred rows: image y0-2 .. y0+TH+1).  B(y,x) = 0 outside [RB, H-RB).  The float pairs are columns (0,1) and (2,3) of the lane's word. ----
		{
			constexpr int RPW = G::B_ROWS / CF_WARPS; // 8 rows per warp
			static_assert(G::B_ROWS % CF_WARPS == 0, "B_ROWS must split evenly across warps");
			const int rb0 = warp * RPW;
			f32x2 win[RPW + 2 * RB][2];
#pragma unroll
			for (int r = 0; r < RPW + 2 * RB; ++r) {
				const unsigned int w = sM[(rb0 + r) * CF_ROWW + lane];
				win[r][0] = u8x2_to_f32x2(w, 0, w, 1, negMagic);
				win[r][1] = u8x2_to_f32x2(w, 2, w, 3, negMagic);
			}
#pragma unroll
			for (int j = 0; j < RPW; ++j) {
				const int y = y0 - 2 + rb0 + j;
				unsigned int outw = 0;
				if (y >= RB && y < H - RB) {
					unsigned int o[4];
#pragma unroll
					for (int h = 0; h < 2; ++h) {
						f32x2 s = fmul2(win[j][h], kk[0]);
#pragma unroll
						for (int k = 1; k < BKS; ++k) s = ffma2(win[j + k][h], kk[k], s);
						unpk2(fadd2_rz(s, magic), o[2 * h], o[2 * h + 1]);
					}
					outw = pack4(o[0], o[1], o[2], o[3]);
				}
				sA[(rb0 + j) * CF_ROWW + lane] = outw;
			}
		}
		__syncthreads();
	}
	// here sA rows hold the image the gradient runs on: row rb <-> image y0-2+rb, lane word <-> columns xl..xl+3

	// ---- S3: Sobel 3x3 + L1 magnitude + direction code -> sG ----
	int tLow = p.tLow, tHigh = p.tHigh;
	if (p.thr) { const ushort2 t = p.thr[frame]; tLow = t.x; tHigh = t.y; }
	// per-warp candidate queue of stage S4 (entries: row * 128 + column inside the g tile); every warp queues, tests and writes out the rows it computes g for
	unsigned short* q = reinterpret_cast<unsigned short*>(sA + (G::OFF_Q - G::OFF_A)) + warp * (G::Q_WORDS_PER_WARP * 2);
	unsigned int candWord = 0; // lane 4j + i: which lanes hold a candidate at pixel slot i of the warp's j-th g row (RPW = 8 rows x 4 slots = 32 lanes)
	unsigned int* sOut = BKS ? sM : (sA + (G::OFF_M - G::OFF_A)); // class tile, 32 words per output row (only with a blur: without one there is no spare tile)
	{
		constexpr int RPW = (G::G_ROWS + CF_WARPS - 1) / CF_WARPS; // 8
		const int rg0 = warp * RPW;
		// Two pixels per 32-bit integer instruction: 16-bit halves hold pixels (0, 2) and (1, 3) of the lane's word, every intermediate value is kept non-negative by a bias
		// (so no borrow crosses the halves).  Per source row: hs = p[i-1] + 2p[i] + p[i+1] (<= 1020), hd = p[i+1] - p[i-1] + 256.
		unsigned int hs[3][2], hd[3][2];
		// the gradient source is the blurred tile (pitch 32, word = lane) or, without blur, the staged input itself (pitch 36, word = woff+lane)
		const int srcPitch = BKS ? CF_ROWW : CF_INW;
		const unsigned int* src = sA + (BKS ? 0 : woff) + lane;
		auto loadRow = [&](int rb, int slot) {
			const unsigned int* sw = src + rb * srcPitch;
			const unsigned int wl = sw[-1], wc = sw[0], wr = sw[1];
			const unsigned int B = __byte_perm(wc, 0u, 0x4240);  // (p0, p2)
			const unsigned int Cc = __byte_perm(wc, 0u, 0x4341); // (p1, p3)
			const unsigned int A = __byte_perm(wl, Cc, 0x5453);  // (p-1, p1)
			const unsigned int D = __byte_perm(B, wr, 0x1412);   // (p2, p4)
			hs[slot][0] = A + 2u * B + Cc;
			hs[slot][1] = B + 2u * Cc + D;
			hd[slot][0] = Cc + 0x01000100u - A;
			hd[slot][1] = D + 0x01000100u - B;
		};
		// columns 1 <= x < W-1 of my four pixels, as masks over the (0, 2) and (1, 3) pairs
		unsigned int cm[2] = { 0u, 0u };
#pragma unroll
		for (int i = 0; i < 4; ++i) if (xl + i >= 1 && xl + i < W - 1) cm[i & 1] |= 0xffffu << (16 * (i >> 1));
		// NMS candidates (g > tLow) are queued right here, one ballot per pixel slot: adding 0x8000 - (tLow + 1) to a 16-bit half sets its top bit exactly when g > tLow
		// (g <= 2040: nothing carries into the other half).  Lanes 0 and 31 hold halo columns only, rows 0 and G_ROWS-1 are halo rows.
		const unsigned int candAdd = (0x8000u - static_cast<unsigned int>(min(tLow, 0x7ffe) + 1)) * 0x10001u;
		const unsigned int candLane = (lane >= 1 && lane <= 30) ? 0x80008000u : 0u;
		loadRow(rg0, 0);
		loadRow(rg0 + 1, 1);
#pragma unroll
		for (int j = 0; j < RPW; ++j) {
			const int rg = rg0 + j;
			if (rg < G::G_ROWS) { // warp-uniform
				loadRow(rg + 2, (j + 2) % 3);
				const int a = j % 3, b = (j + 1) % 3, c = (j + 2) % 3;
				const int y = y0 - 1 + rg;
				uint2 o = make_uint2(0u, 0u);
				if (y >= 1 && y < H - 1) {
					unsigned int g[2];
#pragma unroll
					for (int h = 0; h < 2; ++h) {
						const unsigned int gx = hd[a][h] + 2u * hd[b][h] + hd[c][h];    // gx + 1024 in [4, 2044]
						const unsigned int gy = hs[c][h] + 0x04000400u - hs[a][h];      // gy + 1024
						const unsigned int ax = __vmaxu2(gx, 0x08000800u - gx);         // |gx| + 1024
						const unsigned int ay = __vmaxu2(gy, 0x08000800u - gy);
						// the NMS direction is NOT computed here: only the few per cent of the pixels with g > tLow need it, stage S4 recomputes gx / gy for those
						g[h] = (ax + ay - 0x08000800u) & cm[h];
					}
					o.x = __byte_perm(g[0], g[1], 0x5410); // g0 | g1 << 16
					o.y = __byte_perm(g[0], g[1], 0x7632); // g2 | g3 << 16
					if (rg >= 1 && rg <= CF_TH) { // warp-uniform
						const unsigned int f[2] = { (g[0] + candAdd) & candLane, (g[1] + candAdd) & candLane };
#pragma unroll
						for (int i = 0; i < 4; ++i) { // lane 4j + i keeps the candidate mask of (my j-th row, pixel slot i): expanded into the queue once, in S4
							const unsigned int bal = __ballot_sync(0xffffffffu, f[i & 1] & (0x8000u << (16 * (i >> 1))));
							if (lane == 4 * j + i) candWord = bal;
						}
					}
				}
				*reinterpret_cast<uint2*>(&sGw[rg * 64 + lane * 2]) = o;
				if (BKS && rg >= 1 && rg <= CF_TH) sOut[(rg - 1) * CF_ROWW + lane] = 0; // the class tile starts out empty (sM is dead since the barrier before this stage)
			}
		}
	}
	__syncthreads();

	// ---- S4: NMS on the unsuppressed g + classification -> global ----
	// Only ~13 % of the pixels pass g > tLow, and they cluster on a few lanes: testing them where they lie keeps 1-2 lanes of a warp busy for tens of
	// instructions per row.  Instead every warp (A) turns the candidate masks of its rows (ballots taken in S3) into a queue, (B) works the queue off with all 32
	// lanes -- gradient recomputed from the blurred tile that is still resident, direction, the two neighbours along it -- writing the class byte into the class
	// tile (the dead sM), and (C) streams its rows of that tile to global memory.
	{
		constexpr int RPW = (G::G_ROWS + CF_WARPS - 1) / CF_WARPS; // 8: the same row ownership as S3
		const unsigned short* sG = reinterpret_cast<const unsigned short*>(sGw);
		uint8_t* __restrict__ cls = p.cls + frame * p.framePitch;
		// (A) expand the candidate masks into the queue: an exclusive prefix of the per-lane counts, then every lane writes out the set bits of its word
		unsigned int n;
		{
			const unsigned int c = __popc(candWord);
			unsigned int pre = c;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, pre, d); if (lane >= d) pre += t; }
			n = __shfl_sync(0xffffffffu, pre, 31);
			unsigned int pos = pre - c;
			const unsigned int eBase = static_cast<unsigned int>((warp * RPW + (lane >> 2)) * 128 + (lane & 3)); // row of the word, pixel slot
			for (unsigned int w = candWord; w; w &= w - 1) q[pos++] = static_cast<unsigned short>(eBase + 4u * static_cast<unsigned int>(__ffs(w) - 1));
			__syncwarp();
		}
		// (B)
		constexpr int PB = (BKS ? CF_ROWW : CF_INW) * 4;
		const unsigned char* tile = reinterpret_cast<const unsigned char*>(sA + (BKS ? 0 : woff));
		unsigned char* outBytes = reinterpret_cast<unsigned char*>(sOut);
		for (unsigned int k = lane; k < n; k += 32) {
			const int idx = q[k];
			const int rg = idx >> 7, col = idx & 127;
			const int gc = sG[idx];
			const unsigned char* t = tile + rg * PB + col; // row y-1, my column
			const int a0 = t[-1], a1 = t[0], a2 = t[1], b0 = t[PB - 1], b2 = t[PB + 1], c0 = t[2 * PB - 1], c1 = t[2 * PB], c2 = t[2 * PB + 1];
			const int gx = (a2 - a0) + 2 * (b2 - b0) + (c2 - c0);
			const int gy = (c0 + 2 * c1 + c2) - (a0 + 2 * a1 + a2);
			const int ax = abs(gx), ays = abs(gy) << 16;
			int off;
			if (ays < kTangentPiOver8Int * ax) off = 1;
			else if (ays < kTangentPiTimes3Over8Int * ax) off = ((gx ^ gy) < 0) ? -127 : 129;
			else off = 128;
			const int n0 = sG[idx - off], n1 = sG[idx + off];
			if (!(n0 > gc || n1 > gc)) {
				const uint8_t c = gc > tHigh ? CLS_STRONG : CLS_WEAK;
				if (BKS) outBytes[(rg - 1) * (CF_ROWW * 4) + col] = c;
				else { // no blurred tile, hence no spare tile: the few survivors go straight to global memory, after this warp's zero fill below (same warp, ordered by the __syncwarp + fence)
					const int y = y0 + rg - 1, x = x0 - 4 + col;
					if (y < H && x < W) q[k] = static_cast<unsigned short>(idx | (c == CLS_STRONG ? 0x8000 : 0x4000));
					else q[k] = 0;
					continue;
				}
			}
			else if (!BKS) q[k] = 0;
		}
		__syncwarp();
		// (C)
		const bool laneOut = (lane >= 1 && lane <= 30);
		const int roBegin = max(warp * RPW - 1, 0), roEnd = min(warp * RPW + RPW - 1, CF_TH); // output rows ro = rg - 1 of my g rows
		if (laneOut && xl < W) {
			uint8_t* o = cls + static_cast<size_t>(y0 + roBegin) * p.stride + xl;
			for (int ro = roBegin; ro < roEnd && y0 + ro < H; ++ro, o += p.stride) {
				const unsigned int outw = BKS ? sOut[ro * CF_ROWW + lane] : 0u;
				if (p.vecStore && xl + 4 <= W) {
					*reinterpret_cast<unsigned int*>(o) = outw;
				}
				else {
#pragma unroll
					for (int i = 0; i < 4; ++i) if (xl + i < W) o[i] = static_cast<uint8_t>(outw >> (8 * i));
				}
			}
		}
		if (!BKS) {
			__threadfence_block();
			__syncwarp();
			for (unsigned int k = lane; k < n; k += 32) {
				const unsigned int e = q[k];
				if (e & 0xc000u) {
					const int idx = e & 0x3fff, rg = idx >> 7, col = idx & 127;
					cls[static_cast<size_t>(y0 + rg - 1) * p.stride + (x0 - 4 + col)] = (e & 0x8000u) ? CLS_STRONG : CLS_WEAK;
				}
			}
		}
	}
}

// ---- host side: launch ----
template <int BKS>
static int launch_canny_fast_t(const FastParams& p0, size_t batch, cudaStream_t stream)
{
	using G = CFGeom<BKS>;
	FastParams p = p0;
	alignas(64) CUtensorMap map;
	memset(&map, 0, sizeof(map));
	p.useTma = make_u8_tile_map(&map, p.in, p.W, p.H, p.stride, p.framePitch, batch, CF_INW * 4, G::IN_ROWS) ? 1 : 0;
	p.vecStore = (((reinterpret_cast<uintptr_t>(p.cls) | p.stride | p.framePitch) & 3) == 0) ? 1 : 0;
	auto kern = canny_front_fast_kernel<BKS>;
	static std::atomic<unsigned int> attrSet{0};
	CVB_CHECK(set_max_smem_once(reinterpret_cast<const void*>(kern), static_cast<int>(G::SMEM), attrSet));
	dim3 grid(static_cast<unsigned>(div_up(p.W, CF_TW)), static_cast<unsigned>(div_up(p.H, CF_TH)), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	{
		KernelScope ks_("canny_front", stream);
		kern<<<grid, CF_THREADS, G::SMEM, stream>>>(map, p);
	}
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

// ---- Sobel detector (edge_dete.cxx:55-206) on the same staged tile: pass 1 (MODE 2) = per-frame maximum of g = |gx| + |gy|, pass 2 (MODE 3) = u8(trunc(g * 255.f / gmax)).
// Same arithmetic as edge_front_kernel<3, 2|3> in edges.cu; 4 pixels per lane, no intermediate tile: the gradient never leaves the registers.
template <int MODE>
__global__ void __launch_bounds__(CF_THREADS, 3)
sobel_fast_kernel(const __grid_constant__ CUtensorMap tmap, const FastParams p, unsigned int* __restrict__ gmaxOut, const unsigned int* __restrict__ gmaxIn, int gmaxLanes)
{
	using G = CFGeom<0>;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const unsigned int pad = (128u - (static_cast<unsigned int>(__cvta_generic_to_shared(smem_raw)) & 127u)) & 127u;
	unsigned int* base = reinterpret_cast<unsigned int*>(smem_raw + pad);
	unsigned int* sA = base + 32;
	uint64_t* bar = reinterpret_cast<uint64_t*>(sA + (G::WORDS - G::OFF_A) + 2);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int W = p.W, H = p.H;
	const int x0 = blockIdx.x * CF_TW, y0 = blockIdx.y * CF_TH;
	const int frame = blockIdx.z;
	const int xl = x0 - 4 + 4 * lane;
	const int yIn0 = y0 - 2;                              // image row of staged row 0 (CFGeom<0>: rows y0-2 .. y0+TH+1)
	const int xTma = (x0 - 4) & ~15;
	const int woff = ((x0 - 4) - xTma) >> 2;
	if (p.useTma) {
		if (threadIdx.x == 0) {
			mbar_init(bar, 1);
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			mbar_expect_tx(bar, G::IN_ROWS * CF_INW * 4);
			tma_load_3d(sA, &tmap, bar, xTma, yIn0, frame);
		}
		__syncthreads();
		mbar_wait(bar, 0);
	}
	else {
		const uint8_t* __restrict__ in = p.in + frame * p.framePitch;
		for (int r = warp; r < G::IN_ROWS; r += CF_WARPS) {
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
			sA[r * CF_INW + woff + lane] = w;
		}
		__syncthreads();
	}
	constexpr int RPW = (CF_TH + CF_WARPS - 1) / CF_WARPS; // 8 output rows per warp
	const int ro0 = warp * RPW;
	// Two pixels per 32-bit integer instruction, as in canny_front_fast_kernel: 16-bit halves hold pixels (0, 2) and (1, 3) of the lane's word, biased to stay non-negative.
	unsigned int hs[3][2], hd[3][2]; // per source row: hs = p[i-1] + 2p[i] + p[i+1], hd = p[i+1] - p[i-1] + 256
	const unsigned int* src = sA + woff + lane;
	auto loadRow = [&](int rb, int slot) {
		const unsigned int* sw = src + rb * CF_INW;
		const unsigned int wl = sw[-1], wc = sw[0], wr = sw[1];
		const unsigned int B = __byte_perm(wc, 0u, 0x4240);  // (p0, p2)
		const unsigned int Cc = __byte_perm(wc, 0u, 0x4341); // (p1, p3)
		const unsigned int A = __byte_perm(wl, Cc, 0x5453);  // (p-1, p1)
		const unsigned int D = __byte_perm(B, wr, 0x1412);   // (p2, p4)
		hs[slot][0] = A + 2u * B + Cc;
		hs[slot][1] = B + 2u * Cc + D;
		hd[slot][0] = Cc + 0x01000100u - A;
		hd[slot][1] = D + 0x01000100u - B;
	};
	const bool laneOut = (lane >= 1 && lane <= 30);
	// the r = 1 border ring of the convolutions is zero: columns 1 <= x < W-1 of my four pixels as masks over the (0, 2) and (1, 3) pairs;
	// pass 1 additionally looks only at the columns that count for the frame maximum (all of the image, or the x86 lanes of defect 1)
	unsigned int cm[2] = { 0u, 0u }, mm[2] = { 0u, 0u };
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int x = xl + i;
		if (x >= 1 && x < W - 1) cm[i & 1] |= 0xffffu << (16 * (i >> 1));
		if (laneOut && x < W && (!gmaxLanes || ((0x17u >> (x & 7)) & 1u))) mm[i & 1] |= 0xffffu << (16 * (i >> 1));
	}
	unsigned int localMax2 = 0; // packed pair of running maxima
	float scale = 0.f;
	if (MODE == 3) scale = __fdiv_rn(255.f, static_cast<float>(max(gmaxIn[frame], 1u))); // scaleAndClip (compv_math_utils.cxx:336-364)
	const f32x2 scale2 = pk2(__float_as_uint(scale), __float_as_uint(scale));
	const f32x2 negMagic = pk2(0xCB000000u, 0xCB000000u), magic = pk2(0x4B000000u, 0x4B000000u);
	uint8_t* __restrict__ cls = p.cls + frame * p.framePitch;
	// output row y0 + ro reads staged rows ro + 1, ro + 2, ro + 3 (image rows y-1, y, y+1)
	loadRow(ro0 + 1, 0);
	loadRow(ro0 + 2, 1);
#pragma unroll
	for (int j = 0; j < RPW; ++j) {
		const int ro = ro0 + j;
		if (ro < CF_TH) { // warp-uniform
			loadRow(ro + 3, (j + 2) % 3);
			const int a = j % 3, b = (j + 1) % 3, c = (j + 2) % 3;
			const int y = y0 + ro;
			unsigned int g[2] = { 0u, 0u };
			if (y >= 1 && y < H - 1) {
#pragma unroll
				for (int h = 0; h < 2; ++h) {
					const unsigned int gx = hd[a][h] + 2u * hd[b][h] + hd[c][h];    // gx + 1024 in [4, 2044]
					const unsigned int gy = hs[c][h] + 0x04000400u - hs[a][h];      // gy + 1024
					const unsigned int ax = __vmaxu2(gx, 0x08000800u - gx);         // |gx| + 1024
					const unsigned int ay = __vmaxu2(gy, 0x08000800u - gy);
					g[h] = (ax + ay - 0x08000800u) & cm[h];
				}
			}
			if (MODE == 2) {
				if (y < H) localMax2 = __vmaxu2(localMax2, __vmaxu2(g[0] & mm[0], g[1] & mm[1]));
			}
			else {
				// u8(trunc(g * scale)) clamped to 255: (float)g by the 2^23 trick (g <= 2040), the product is < 2^23, adding 2^23 toward zero leaves trunc() in the low mantissa bits
				unsigned int v[4];
#pragma unroll
				for (int h = 0; h < 2; ++h) {
					const f32x2 gf = fadd2(pk2(__byte_perm(g[h], 0x4B000000u, 0x7410), __byte_perm(g[h], 0x4B000000u, 0x7432)), negMagic); // pixels h and h + 2
					unsigned int lo, hi;
					unpk2(fadd2_rz(fmul2(gf, scale2), magic), lo, hi);
					v[h] = min(lo & 0x7fffffu, 255u); v[h + 2] = min(hi & 0x7fffffu, 255u);
				}
				const unsigned int outw = pack4(v[0], v[1], v[2], v[3]);
				if (laneOut && y < H && xl < W) {
					uint8_t* o = cls + static_cast<size_t>(y) * p.stride + xl;
					if (p.vecStore && xl + 4 <= W) *reinterpret_cast<unsigned int*>(o) = outw;
					else {
#pragma unroll
						for (int i = 0; i < 4; ++i) if (xl + i < W) o[i] = static_cast<uint8_t>(outw >> (8 * i));
					}
				}
			}
		}
	}
	unsigned int localMax = max(localMax2 & 0xffffu, localMax2 >> 16);
	if (MODE == 2) {
		for (int o = 16; o; o >>= 1) localMax = max(localMax, __shfl_xor_sync(0xffffffffu, localMax, o));
		if (lane == 0 && localMax) atomicMax(&gmaxOut[frame], localMax);
	}
}

static int launch_sobel_fast(const FastParams& p0, unsigned int* gmax, int gmaxLanes, size_t batch, cudaStream_t stream)
{
	using G = CFGeom<0>;
	FastParams p = p0;
	alignas(64) CUtensorMap map;
	memset(&map, 0, sizeof(map));
	p.useTma = make_u8_tile_map(&map, p.in, p.W, p.H, p.stride, p.framePitch, batch, CF_INW * 4, G::IN_ROWS) ? 1 : 0;
	p.vecStore = (((reinterpret_cast<uintptr_t>(p.cls) | p.stride | p.framePitch) & 3) == 0) ? 1 : 0;
	static std::atomic<unsigned int> attrSet2{0}, attrSet3{0};
	CVB_CHECK(set_max_smem_once(reinterpret_cast<const void*>(sobel_fast_kernel<2>), static_cast<int>(G::SMEM), attrSet2));
	CVB_CHECK(set_max_smem_once(reinterpret_cast<const void*>(sobel_fast_kernel<3>), static_cast<int>(G::SMEM), attrSet3));
	dim3 grid(static_cast<unsigned>(div_up(p.W, CF_TW)), static_cast<unsigned>(div_up(p.H, CF_TH)), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	{ KernelScope ks_("edge_gmax", stream);
	  sobel_fast_kernel<2><<<grid, CF_THREADS, G::SMEM, stream>>>(map, p, gmax, nullptr, gmaxLanes); }
	CVB_LAUNCHED();
	{ KernelScope ks_("edge_normalize", stream);
	  sobel_fast_kernel<3><<<grid, CF_THREADS, G::SMEM, stream>>>(map, p, nullptr, gmax, gmaxLanes); }
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

} // namespace cvb
