// a3 / a5 -- Sobel / Scharr / Prewitt normalised-gradient detectors and the Canny detector.
// Replaces CompVCornerDeteEdgeBase::process (core/features/edges/compv_core_feature_edge_dete.cxx:55-206) and
// CompVEdgeDeteCanny::process / nms_gather / nms_apply / hysteresis (core/features/edges/compv_core_feature_canny_dete.cxx:123-528,566-680).
//
// Data flow on the device (per frame, batched along grid.z / tile index):
//   canny_front : u8 image tile (+halo) in shared memory -> [optional fused Gaussian blur, u8 intermediate] -> gx, gy (int16), g = |gx|+|gy| (u16)
//                 in shared memory -> NMS on the UNSUPPRESSED g (the reference gathers first and applies later, canny_dete.cxx:278-299)
//                 -> one byte per pixel: 0 (nothing), WEAK (tLow < g <= tHigh, kept by NMS), STRONG (g > tHigh, kept by NMS).
//                 The CPU path's gx/gy/g/nms planes (2+2+2+1 B/px written and re-read) never touch HBM.
//   hysteresis  : 8-connected closure of STRONG over WEAK. The reference's DFS (canny_dete.cxx:600-680) is order independent:
//                 its result is exactly the set of {g>tLow} pixels 8-connected to a {g>tHigh} seed. Tiles are relaxed to a fixed point in shared
//                 memory; a tile whose border pixels changed marks its neighbours dirty; rounds repeat until no tile is dirty.
//   finalize    : WEAK -> 0, STRONG -> 0xff (in place).
#include "edge.cuh"

#include <cooperative_groups.h>
#include <cstring>

namespace cg = cooperative_groups;

namespace cvb {

// ---- kernel tables: base/include/compv/base/compv_features.h:124-133 (EdgeTaps / BlurTaps: edge.cuh) ----

static int edge_taps(int id, size_t kernSize, EdgeTaps* t)
{
	memset(t, 0, sizeof(*t));
	switch (id) {
	case CVB200_SOBEL_ID:
	case CVB200_CANNY_ID:
		if (kernSize == 5) { // CompVSobel5x5Gx_vt/_hz
			const int16_t vt[5] = { 1, 4, 6, 4, 1 }, hz[5] = { 1, 2, 0, -2, -1 };
			memcpy(t->vt, vt, sizeof(vt)); memcpy(t->hz, hz, sizeof(hz)); t->ks = 5;
		}
		else { // CompVSobel3x3Gx_vt/_hz
			const int16_t vt[3] = { 1, 2, 1 }, hz[3] = { -1, 0, 1 };
			memcpy(t->vt, vt, sizeof(vt)); memcpy(t->hz, hz, sizeof(hz)); t->ks = 3;
		}
		return CVB200_S_OK;
	case CVB200_SCHARR_ID: {
		const int16_t vt[3] = { 3, 10, 3 }, hz[3] = { -1, 0, 1 };
		memcpy(t->vt, vt, sizeof(vt)); memcpy(t->hz, hz, sizeof(hz)); t->ks = 3;
		return CVB200_S_OK;
	}
	case CVB200_PREWITT_ID: {
		const int16_t vt[3] = { 1, 1, 1 }, hz[3] = { -1, 0, 1 };
		memcpy(t->vt, vt, sizeof(vt)); memcpy(t->hz, hz, sizeof(hz)); t->ks = 3;
		return CVB200_S_OK;
	}
	default:
		return CVB200_E_INVALID_PARAMETER;
	}
}


constexpr uint8_t CLS_WEAK = 0x80;
constexpr uint8_t CLS_STRONG = 0xff;

// canny_dete.h:58-61 : tan(pi/8) and tan(3pi/8) in Q16
constexpr int kTangentPiOver8Int = 27145;
constexpr int kTangentPiTimes3Over8Int = 158217;

constexpr int FT_W = 128, FT_H = 32, FT_THREADS = 256;

// Gradient of the (optionally blurred) image at (gxp, gyp) given a shared-memory u8 tile whose (0,0) is image (ox, oy).
// Equals convlt1<u8,int16,int16>(vt,hz) / (hz,vt) (compv_math_convlt.h:332-353): for the tap tables above the int16 saturation
// between and after the passes can never trigger (|sum| <= 12240), so the 2-D form is bit-identical to the separable one.
template <int KS>
__device__ __forceinline__ void grad_at(const uint8_t* __restrict__ s, int pitch, int lx, int ly, const EdgeTaps& t, int& gx, int& gy)
{
	constexpr int R = KS >> 1;
	int sx = 0, sy = 0;
#pragma unroll
	for (int j = 0; j < KS; ++j) {
		int rowh = 0, rowv = 0;
#pragma unroll
		for (int i = 0; i < KS; ++i) {
			const int p = s[(ly + j - R) * pitch + (lx + i - R)];
			rowh += p * t.hz[i];
			rowv += p * t.vt[i];
		}
		sx += rowh * t.vt[j];
		sy += rowv * t.hz[j];
	}
	gx = sx; gy = sy;
}

struct FrontParams {
	const uint8_t* in;
	uint8_t* cls;           // canny: class map; may be null
	int16_t* gxOut;         // sobel_g outputs; may be null
	int16_t* gyOut;
	uint16_t* gOut;
	unsigned int* gmax;     // per-frame max of g (Sobel detector pass 1); may be null
	const unsigned int* gmaxIn; // per-frame max (Sobel detector pass 2) -> normalised u8 written to cls
	const ushort2* thr;     // per-frame (tLow,tHigh) on the device, or null -> tLow/tHigh below
	int W, H;
	size_t stride, framePitch;
	int tLow, tHigh;
	int gmaxLanes;          // !=0: CVB200_EDGE_SET_BOOL_X86_SSE41_GMAX_LANES (see cvb200.h)
	EdgeTaps taps;
	BlurTaps blur;
};

// MODE 0: Canny class map. MODE 1: gx/gy/g planes. MODE 2: per-frame gmax. MODE 3: normalised u8 = trunc(g*255/gmax).
template <int KS, int MODE>
__global__ void __launch_bounds__(FT_THREADS)
edge_front_kernel(const FrontParams p)
{
	constexpr int RS = KS >> 1;
	constexpr int GH = (MODE == 0) ? 1 : 0;       // g is needed on a 1-px ring around the tile for NMS
	const int rb = p.blur.ks >> 1;                // blur radius (0 when disabled)
	const int halo = rb + RS + GH;
	const int inW = FT_W + 2 * halo, inH = FT_H + 2 * halo;
	const int bW = FT_W + 2 * (RS + GH), bH = FT_H + 2 * (RS + GH); // blurred region needed by the gradient
	constexpr int gW = FT_W + 2 * GH, gH = FT_H + 2 * GH;

	extern __shared__ __align__(16) unsigned char smem[];
	uint8_t* sIn = smem;                                             // inW x inH
	uint8_t* sMid = sIn + ((inW * inH + 15) & ~15);                  // bW x (bH + 2rb)   (blur only)
	uint8_t* sB = sMid + ((bW * (bH + 2 * rb) + 15) & ~15);          // bW x bH           (blur only)
	uint16_t* sG = reinterpret_cast<uint16_t*>(sB + ((bW * bH + 15) & ~15)); // gW x gH
	int16_t* sGx = reinterpret_cast<int16_t*>(sG + ((gW * gH + 7) & ~7));    // FT_W x FT_H (MODE 0 only)
	int16_t* sGy = sGx + FT_W * FT_H;

	const int W = p.W, H = p.H;
	const int x0 = blockIdx.x * FT_W, y0 = blockIdx.y * FT_H;
	const size_t frameOff = blockIdx.z * p.framePitch;
	const uint8_t* __restrict__ in = p.in + frameOff;
	const int tid = threadIdx.x;

	// ---- stage the input tile; outside the image -> 0 (never contributes to a valid sample) ----
	{
		const int ox = x0 - halo, oy = y0 - halo;
		for (int i = tid; i < inW * inH; i += FT_THREADS) {
			const int ly = i / inW, lx = i - ly * inW;
			const int gx = ox + lx, gy = oy + ly;
			uint8_t v = 0;
			if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = in[static_cast<size_t>(gy) * p.stride + gx];
			sIn[i] = v;
		}
	}
	__syncthreads();

	const uint8_t* sSrc = sIn; // image the gradient runs on
	int srcPitch = inW;
	if (rb) {
		// ---- fused Gaussian blur = convlt1<u8,f32,u8> (compv_math_convlt.h:358-384; fma chain, truncation, see convlt.cu) ----
		// horizontal: sMid(y, x) for y in [y0-RS-GH-rb, +bH+2rb), x in [x0-RS-GH, +bW)
		const int mH = bH + 2 * rb;
		const int mox = x0 - RS - GH, moy = y0 - RS - GH - rb;
		for (int i = tid; i < bW * mH; i += FT_THREADS) {
			const int ly = i / bW, lx = i - ly * bW;
			const int gx = mox + lx, gy = moy + ly;
			uint8_t m = 0;
			if (gy >= 0 && gy < H && gx >= rb && gx < W - rb) {
				float sum = 0.f;
				const uint8_t* q = &sIn[ly * inW + lx]; // sIn col of (gx - rb) == lx
				for (int k = 0; k < p.blur.ks; ++k) sum = __fmaf_rn(static_cast<float>(q[k]), p.blur.k[k], sum);
				m = static_cast<uint8_t>(__float2int_rz(fminf(fmaxf(sum, 0.f), 255.f)));
			}
			sMid[i] = m;
		}
		__syncthreads();
		// vertical: sB(y, x) for y in [y0-RS-GH, +bH)
		const int boy = y0 - RS - GH;
		for (int i = tid; i < bW * bH; i += FT_THREADS) {
			const int ly = i / bW, lx = i - ly * bW;
			const int gx = mox + lx, gy = boy + ly;
			uint8_t b = 0;
			if (gy >= rb && gy < H - rb && gx >= 0 && gx < W) {
				float sum = 0.f;
				const uint8_t* q = &sMid[ly * bW + lx];
				for (int k = 0; k < p.blur.ks; ++k) sum = __fmaf_rn(static_cast<float>(q[k * bW]), p.blur.k[k], sum);
				b = static_cast<uint8_t>(__float2int_rz(fminf(fmaxf(sum, 0.f), 255.f)));
			}
			sB[i] = b;
		}
		__syncthreads();
		sSrc = sB;
		srcPitch = bW;
	}

	// ---- gradient on the tile (+1 ring for NMS): zero on the RS-wide image border ring (compv_math_convlt.h:176-292) ----
	unsigned int localMax = 0;
	for (int i = tid; i < gW * gH; i += FT_THREADS) {
		const int ly = i / gW, lx = i - ly * gW;
		const int gxp = x0 - GH + lx, gyp = y0 - GH + ly;
		int gx = 0, gy = 0;
		if (gxp >= RS && gxp < W - RS && gyp >= RS && gyp < H - RS) {
			grad_at<KS>(sSrc, srcPitch, lx + RS, ly + RS, p.taps, gx, gy);
		}
		// K4: CompVMathUtils::sumAbs (compv_math_utils.h:173-186); the SIMD leaves saturate to u16 -- unreachable here (g <= 24480)
		const unsigned int g = static_cast<unsigned int>(abs(gx) + abs(gy));
		const bool inTile = (lx >= GH && lx < GH + FT_W && ly >= GH && ly < GH + FT_H);
		const bool inImg = (gxp >= 0 && gxp < W && gyp >= 0 && gyp < H);
		if (MODE == 0) {
			sG[i] = static_cast<uint16_t>(g);
			if (inTile) {
				const int ti = (ly - GH) * FT_W + (lx - GH);
				sGx[ti] = static_cast<int16_t>(gx);
				sGy[ti] = static_cast<int16_t>(gy);
			}
		}
		else if (inTile && inImg) {
			const size_t o = frameOff + static_cast<size_t>(gyp) * p.stride + gxp;
			if (MODE == 1) {
				if (p.gxOut) p.gxOut[o] = static_cast<int16_t>(gx);
				if (p.gyOut) p.gyOut[o] = static_cast<int16_t>(gy);
				if (p.gOut) p.gOut[o] = static_cast<uint16_t>(g);
			}
			else if (MODE == 2) {
				if (!p.gmaxLanes || ((0x17u >> (gxp & 7)) & 1u)) localMax = max(localMax, g);
			}
			else { // MODE 3: edge_dete.cxx:199-202 + scaleAndClip (compv_math_utils.cxx:336-364): trunc(g * (255.f / gmax)), saturate
				const float scale = __fdiv_rn(255.f, static_cast<float>(max(p.gmaxIn[blockIdx.z], 1u)));
				const int v = __float2int_rz(__fmul_rn(static_cast<float>(g), scale));
				p.cls[o] = static_cast<uint8_t>(min(max(v, 0), 255));
			}
		}
	}
	if (MODE == 2) {
		for (int o = 16; o; o >>= 1) localMax = max(localMax, __shfl_xor_sync(0xffffffffu, localMax, o));
		if ((tid & 31) == 0 && localMax) atomicMax(&p.gmax[blockIdx.z], localMax);
		return;
	}
	if (MODE != 0) return;
	__syncthreads();

	// ---- K6 NMS (canny_dete.cxx:566-598) fused with the apply step K7 (:414-460) and the threshold classification ----
	int tLow = p.tLow, tHigh = p.tHigh;
	if (p.thr) { const ushort2 t = p.thr[blockIdx.z]; tLow = t.x; tHigh = t.y; }
	uint8_t* __restrict__ cls = p.cls + frameOff;
	for (int i = tid; i < FT_W * FT_H; i += FT_THREADS) {
		const int ly = i / FT_W, lx = i - ly * FT_W;
		const int gxp = x0 + lx, gyp = y0 + ly;
		if (gxp >= W || gyp >= H) continue;
		uint8_t c = 0;
		if (gxp >= 1 && gxp < W - 1 && gyp >= 1 && gyp < H - 1) {
			const uint16_t* g = &sG[(ly + 1) * gW + (lx + 1)];
			const int gc = g[0];
			if (gc > tLow) {
				const int gxi = sGx[i], gyi = sGy[i];
				const int absgy = abs(gyi) << 16, absgx = abs(gxi);
				int n0, n1;
				if (absgy < kTangentPiOver8Int * absgx) { n0 = g[-1]; n1 = g[1]; }
				else if (absgy < kTangentPiTimes3Over8Int * absgx) {
					const int c = ((gxi ^ gyi) < 0) ? (1 - gW) : (1 + gW);
					n0 = g[-c]; n1 = g[c];
				}
				else { n0 = g[-gW]; n1 = g[gW]; }
				if (!(n0 > gc || n1 > gc)) c = (gc > tHigh) ? CLS_STRONG : CLS_WEAK;
			}
		}
		cls[static_cast<size_t>(gyp) * p.stride + gxp] = c;
	}
}

// ---- hysteresis -------------------------------------------------------------------------------
constexpr int HT = 64;              // tile edge
constexpr int HT_THREADS = 256;     // each thread owns a 16-px row segment
constexpr int HP = HT + 32;         // shared pitch: 16 pad bytes | 64 tile bytes | 16 pad bytes, so that a thread's 16-byte segment is one aligned vector
constexpr int HX0 = 16;             // column of tile pixel 0 inside a shared row (the 1-px halo sits at HX0 - 1 and HX0 + HT)
constexpr int HROWS = HT + 2;       // shared rows: halo, 64 tile rows, halo

struct HystParams {
	uint8_t* cls;
	int W, H;
	size_t stride, framePitch;
	int tilesX, tilesY, nTiles;     // nTiles = tilesX*tilesY*batch
	int round;                      // 0: every tile (tile = blockIdx.x); r > 0: the tiles queued by round r-1
	const int* listIn;              // round > 0: tile ids to visit
	const unsigned int* countIn;
	int* listOut;                   // tiles whose neighbourhood changed in this round: visited by round + 1
	unsigned int* countOut;
	int* epoch;                     // per tile: the last round it was queued for (a tile is queued at most once per round)
};

// One tile: load with a 1-px halo, relax to the fixed point inside the tile, write the promotions back, queue the neighbours that can see a change.
__device__ __forceinline__ void hysteresis_tile(const HystParams& p, int tile, uint8_t* s, int* sBorder)
{
	const int tid = threadIdx.x;
	if (tid == 0) *sBorder = 0;
	const int frame = tile / (p.tilesX * p.tilesY);
	const int t2 = tile - frame * (p.tilesX * p.tilesY);
	const int ty = t2 / p.tilesX, tx = t2 - ty * p.tilesX;
	const int x0 = tx * HT, y0 = ty * HT;
	uint8_t* __restrict__ cls = p.cls + frame * p.framePitch;

	// my segment: row ry, columns [16*seg, 16*seg+16): one aligned 16-byte load where the frame allows it
	const int ry = tid >> 2, seg = tid & 3;
	uint8_t* row = &s[(ry + 1) * HP + HX0 + seg * 16];
	unsigned int weak = 0;
	{
		const int gy = y0 + ry, gx = x0 + seg * 16;
		const uint8_t* src = cls + static_cast<size_t>(gy) * p.stride + gx;
		uint4 v = make_uint4(0, 0, 0, 0);
		if (gy < p.H && gx + 16 <= p.W && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) v = *reinterpret_cast<const uint4*>(src);
		else if (gy < p.H) {
			unsigned int w[4] = { 0, 0, 0, 0 };
			for (int k = 0; k < 16 && gx + k < p.W; ++k) w[k >> 2] |= static_cast<unsigned int>(src[k]) << (8 * (k & 3));
			v = make_uint4(w[0], w[1], w[2], w[3]);
		}
		*reinterpret_cast<uint4*>(row) = v;
		// a byte is 0x00, 0x80 (weak) or 0xff (strong): weak <=> bit 7 set and bit 0 clear
		auto weakBits = [](unsigned int w) { const unsigned int m = (w >> 7) & ~w & 0x01010101u; return (m * 0x01020408u) >> 24; };
		weak = (weakBits(v.x) & 15u) | ((weakBits(v.y) & 15u) << 4) | ((weakBits(v.z) & 15u) << 8) | ((weakBits(v.w) & 15u) << 12);
	}
	// the 1-px halo: top and bottom rows (66 bytes each), left and right columns (64 each)
	for (int i = tid; i < 2 * (HT + 2) + 2 * HT; i += HT_THREADS) {
		int ly, lx;
		if (i < 2 * (HT + 2)) { ly = (i < HT + 2) ? -1 : HT; lx = ((i < HT + 2) ? i : i - (HT + 2)) - 1; }
		else { const int k = i - 2 * (HT + 2); ly = k & (HT - 1); lx = (k < HT) ? -1 : HT; }
		const int gx = x0 + lx, gy = y0 + ly;
		uint8_t v = 0;
		if (gx >= 0 && gx < p.W && gy >= 0 && gy < p.H) v = cls[static_cast<size_t>(gy) * p.stride + gx];
		s[(ly + 1) * HP + HX0 + lx] = v;
	}
	__syncthreads();
	unsigned int promoted = 0;

	// Chaotic relaxation: a thread reads its neighbours' cells while their owners may be promoting them (compute-sanitizer racecheck reports exactly this
	// read/write pair, profiles/r1_compute_sanitizer_racecheck.log).  It is deliberate: cells only ever change weak -> strong (one byte store), a stale read merely
	// postpones a promotion to the next sweep, and the loop runs until a full sweep changes nothing, so the fixed point -- the 8-connected closure -- is unique.
	while (true) {
		bool changed = false;
		// forward then backward sweep over my weak pixels (Gauss-Seidel inside the segment)
		for (int pass = 0; pass < 2 && weak; ++pass) {
			unsigned int m = weak;
			while (m) {
				const int k = pass ? (31 - __clz(m)) : (__ffs(m) - 1);
				m &= ~(1u << k);
				const uint8_t* q = row + k;
				const bool hit = (q[-1] == CLS_STRONG) | (q[1] == CLS_STRONG)
					| (q[-HP - 1] == CLS_STRONG) | (q[-HP] == CLS_STRONG) | (q[-HP + 1] == CLS_STRONG)
					| (q[HP - 1] == CLS_STRONG) | (q[HP] == CLS_STRONG) | (q[HP + 1] == CLS_STRONG);
				if (hit) {
					row[k] = CLS_STRONG;
					weak &= ~(1u << k);
					promoted |= (1u << k);
					changed = true;
				}
			}
		}
		if (!__syncthreads_or(changed)) break;
	}

	if (promoted) {
		const int gy = y0 + ry;
		uint8_t* o = &cls[static_cast<size_t>(gy) * p.stride + x0 + seg * 16];
		unsigned int m = promoted;
		int b = 0;
		while (m) {
			const int k = __ffs(m) - 1;
			m &= ~(1u << k);
			o[k] = CLS_STRONG;
			const int lx = seg * 16 + k;
			if (lx == 0) b |= 4;
			if (lx == HT - 1) b |= 8;
		}
		if (ry == 0) b |= 1;
		if (ry == HT - 1) b |= 2;
		if (b) atomicOr(sBorder, b);
	}
	__syncthreads();
	if (tid < 9 && tid != 4 && *sBorder) {
		// bit0 top, bit1 bottom, bit2 left, bit3 right: a neighbour tile is queued for the next round when a promoted pixel lies on the border it shares with this tile
		const int b = *sBorder;
		const int dy = tid / 3 - 1, dx = tid % 3 - 1;
		const bool sees = !((dy < 0 && !(b & 1)) || (dy > 0 && !(b & 2)) || (dx < 0 && !(b & 4)) || (dx > 0 && !(b & 8)));
		const int nx = tx + dx, ny = ty + dy;
		if (sees && nx >= 0 && nx < p.tilesX && ny >= 0 && ny < p.tilesY) {
			const int nb = frame * (p.tilesX * p.tilesY) + ny * p.tilesX + nx;
			if (atomicMax(&p.epoch[nb], p.round + 1) < p.round + 1) p.listOut[atomicAdd(p.countOut, 1u)] = nb;
		}
	}
	__syncthreads(); // s / sBorder are reused by the next tile of this CTA
}

// Round 0 visits every tile (grid = nTiles).  Later rounds run on a fixed small grid over the list the previous round queued: a round with nothing queued costs one
// empty launch, so a fixed number of rounds can be issued without asking the host whether the previous one changed anything.
template <bool FIRST>
__global__ void __launch_bounds__(HT_THREADS)
canny_hysteresis_kernel(const HystParams p)
{
	__shared__ __align__(16) uint8_t s[HROWS * HP];
	__shared__ int sBorder;
	if (FIRST) { hysteresis_tile(p, blockIdx.x, s, &sBorder); return; }
	const unsigned int n = *p.countIn;
	for (unsigned int i = blockIdx.x; i < n; i += gridDim.x) hysteresis_tile(p, p.listIn[i], s, &sBorder);
}

// WEAK -> 0 (in place). 16 pixels per thread; stores only where something changes.
__global__ void canny_finalize_kernel(uint8_t* cls, int W, int H, size_t stride, size_t framePitch)
{
	const int y = blockIdx.y;
	uint8_t* row = cls + blockIdx.z * framePitch + static_cast<size_t>(y) * stride;
	const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
	if (x >= W) return;
	if (x + 16 <= W && ((reinterpret_cast<uintptr_t>(row + x) & 15) == 0)) {
		uint4 v = *reinterpret_cast<const uint4*>(row + x);
		// a byte is 0x00, 0x80 or 0xff: keep it iff bit0 is set
		const unsigned int any80 = ((v.x ^ (v.x << 7)) | (v.y ^ (v.y << 7)) | (v.z ^ (v.z << 7)) | (v.w ^ (v.w << 7))) & 0x80808080u;
		if (any80) {
			auto fix = [](unsigned int w) { const unsigned int keep = (w & 0x01010101u) * 0xffu; return w & keep; };
			v.x = fix(v.x); v.y = fix(v.y); v.z = fix(v.z); v.w = fix(v.w);
			*reinterpret_cast<uint4*>(row + x) = v;
		}
	}
	else {
		for (int k = 0; k < 16 && x + k < W; ++k) if (row[x + k] == CLS_WEAK) row[x + k] = 0;
	}
}

// The same pass that also emits the KHT linking bitmap (1 bit per pixel, bit 31 = leftmost column of a word, two zero rows / one zero word of padding: kht_walk.cuh)
// and counts the edge pixels of each frame.  32 pixels per thread.
// Block = 32 x 8 threads: a warp covers 32 words (1024 pixels) of one row, a block 8 rows (one block per row left most of a 128-thread block idle at 1920 columns
// and made 4.4 M blocks per 4096 frames); one atomic per block on the frame's edge counter.
__global__ void __launch_bounds__(256) canny_finalize_bits_kernel(uint8_t* cls, int W, int H, size_t stride, size_t framePitch, unsigned int* __restrict__ bits, int WW, int padRows,
	unsigned int* __restrict__ edgeCount)
{
	__shared__ unsigned int sCount;
	const int y = blockIdx.y * blockDim.y + threadIdx.y, frame = blockIdx.z;
	if (threadIdx.x == 0 && threadIdx.y == 0) sCount = 0;
	__syncthreads();
	uint8_t* row = cls + frame * framePitch + static_cast<size_t>(y) * stride;
	const int wi = blockIdx.x * blockDim.x + threadIdx.x;
	const int x = wi * 32;
	unsigned int word = 0;
	if (x < W && y < H) {
		auto fix = [](unsigned int w) { const unsigned int keep = (w & 0x01010101u) * 0xffu; return w & keep; };
		auto nz4 = [](unsigned int w) { return ((w & 0x01010101u) * 0x01020408u) >> 24; }; // after fix() a byte is 0x00 or 0xff: bit 0 of each byte -> 4 bits
		if (x + 32 <= W && ((reinterpret_cast<uintptr_t>(row + x) & 15) == 0)) {
			uint4 a = *reinterpret_cast<const uint4*>(row + x), b = *reinterpret_cast<const uint4*>(row + x + 16);
			const unsigned int anyA = ((a.x ^ (a.x << 7)) | (a.y ^ (a.y << 7)) | (a.z ^ (a.z << 7)) | (a.w ^ (a.w << 7))) & 0x80808080u;
			const unsigned int anyB = ((b.x ^ (b.x << 7)) | (b.y ^ (b.y << 7)) | (b.z ^ (b.z << 7)) | (b.w ^ (b.w << 7))) & 0x80808080u;
			a.x = fix(a.x); a.y = fix(a.y); a.z = fix(a.z); a.w = fix(a.w);
			b.x = fix(b.x); b.y = fix(b.y); b.z = fix(b.z); b.w = fix(b.w);
			if (anyA) *reinterpret_cast<uint4*>(row + x) = a;
			if (anyB) *reinterpret_cast<uint4*>(row + x + 16) = b;
			word = (nz4(a.x) & 15u) | ((nz4(a.y) & 15u) << 4) | ((nz4(a.z) & 15u) << 8) | ((nz4(a.w) & 15u) << 12)
				| ((nz4(b.x) & 15u) << 16) | ((nz4(b.y) & 15u) << 20) | ((nz4(b.z) & 15u) << 24) | ((nz4(b.w) & 15u) << 28);
		}
		else {
			for (int k = 0; k < 32 && x + k < W; ++k) {
				if (row[x + k] == CLS_WEAK) row[x + k] = 0;
				if (row[x + k]) word |= 1u << k;
			}
		}
		bits[(static_cast<size_t>(frame) * (H + 2 * padRows) + y + padRows) * WW + wi + 1] = __brev(word);
	}
	unsigned int c = __popc(word);
	for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if (threadIdx.x == 0 && c) atomicAdd(&sCount, c);
	__syncthreads();
	if (threadIdx.x == 0 && threadIdx.y == 0 && sCount) atomicAdd(&edgeCount[frame], sCount);
}

// sum of a u8 frame (CompVMathUtils::sum<uint8_t,uint32_t>, canny_dete.cxx:243) -> PERCENT_OF_MEAN thresholds (:253-258)
__global__ void frame_sum_kernel(const uint8_t* in, int W, int H, size_t stride, size_t framePitch, unsigned int* sums)
{
	const uint8_t* f = in + blockIdx.z * framePitch;
	unsigned int acc = 0;
	for (int y = blockIdx.y; y < H; y += gridDim.y) {
		for (int x = threadIdx.x; x < W; x += blockDim.x) acc += f[static_cast<size_t>(y) * stride + x];
	}
	for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&sums[blockIdx.z], acc);
}

__global__ void mean_thresholds_kernel(const unsigned int* sums, unsigned int count, float fLow, float fHigh, ushort2* thr, int batch)
{
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= batch) return;
	int mean = static_cast<uint8_t>(sums[f] / count);
	mean = clampi(mean, 1, 255);
	int tLow = static_cast<uint16_t>(__float2int_rz(__fmul_rn(static_cast<float>(mean), fLow)));
	int tHigh = static_cast<uint16_t>(__float2int_rz(__fmul_rn(static_cast<float>(mean), fHigh)));
	tLow = max(1, tLow);
	tHigh = max(tLow + 2, tHigh);
	thr[f] = make_ushort2(static_cast<unsigned short>(tLow), static_cast<unsigned short>(min(tHigh, 65535)));
}

// ---- launch helpers ----
static size_t front_smem(int ks, int mode, int blurKs)
{
	const int RS = ks >> 1, GH = (mode == 0) ? 1 : 0, rb = blurKs >> 1;
	const int halo = rb + RS + GH;
	const int inW = FT_W + 2 * halo, inH = FT_H + 2 * halo;
	const int bW = FT_W + 2 * (RS + GH), bH = FT_H + 2 * (RS + GH);
	const int gW = FT_W + 2 * GH, gH = FT_H + 2 * GH;
	size_t n = (inW * inH + 15) & ~15;
	n += (bW * (bH + 2 * rb) + 15) & ~15;
	n += (bW * bH + 15) & ~15;
	n += static_cast<size_t>((gW * gH + 7) & ~7) * 2;
	if (mode == 0) n += static_cast<size_t>(FT_W) * FT_H * 2 * 2;
	return n;
}

template <int KS, int MODE>
static int launch_front_t(const FrontParams& p, size_t batch, cudaStream_t stream)
{
	const size_t smem = front_smem(KS, MODE, p.blur.ks);
	auto kern = edge_front_kernel<KS, MODE>;
	if (smem > 48 * 1024) CVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
	dim3 grid(static_cast<unsigned>(div_up(p.W, FT_W)), static_cast<unsigned>(div_up(p.H, FT_H)), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	{
		static const char* const names[4] = { "canny_front", "sobel_g", "edge_gmax", "edge_normalize" };
		KernelScope ks_(names[MODE], stream);
		kern<<<grid, FT_THREADS, smem, stream>>>(p);
	}
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

static int launch_front(const FrontParams& p, int mode, size_t batch, cudaStream_t stream)
{
	if (p.taps.ks == 3) {
		switch (mode) {
		case 0: return launch_front_t<3, 0>(p, batch, stream);
		case 1: return launch_front_t<3, 1>(p, batch, stream);
		case 2: return launch_front_t<3, 2>(p, batch, stream);
		default: return launch_front_t<3, 3>(p, batch, stream);
		}
	}
	switch (mode) {
	case 0: return launch_front_t<5, 0>(p, batch, stream);
	case 1: return launch_front_t<5, 1>(p, batch, stream);
	case 2: return launch_front_t<5, 2>(p, batch, stream);
	default: return launch_front_t<5, 3>(p, batch, stream);
	}
}

} // namespace cvb

#include "canny_fast.cuh"

using namespace cvb;

// The detector object: caches its scratch like the reference objects do (canny_dete.cxx:133-147)

extern "C" {

int cvb200_edge_dete_new(cvb200_edge_dete_t** dete, int id, float tLow, float tHigh, size_t kernSize)
{
	CVB_REQUIRE(dete, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE_INIT();
	EdgeTaps taps;
	// canny ctor: kernSize==3 ? 3x3 : 5x5 (canny_dete.cxx:50-67); Sobel/Scharr/Prewitt objects always use their 3x3 tables
	// whatever kernSize says (edge_dete.cxx:226-255)
	CVB_CHECK(edge_taps(id, id == CVB200_CANNY_ID ? (kernSize == 3 ? 3 : 5) : 3, &taps));
	cvb200_edge_dete* d = new (std::nothrow) cvb200_edge_dete();
	CVB_REQUIRE(d, CVB200_E_OUT_OF_MEMORY);
	d->id = id;
	d->tLow = tLow;
	d->tHigh = tHigh;
	d->thresholdType = CVB200_CANNY_THRESHOLD_TYPE_COMPARE_TO_GRADIENT;
	d->taps = taps;
	d->gmaxLanes = false;
	d->genericKernel = false;
	memset(&d->blur, 0, sizeof(d->blur));
	*dete = d;
	return CVB200_S_OK;
}

int cvb200_edge_dete_free(cvb200_edge_dete_t** dete)
{
	if (dete && *dete) {
		cvb200_edge_dete* d = *dete;
		d->dirty.release(); d->counters.release(); d->hostFlag.release(); d->hostIn.release(); d->hostOut.release();
		delete d;
		*dete = nullptr;
	}
	return CVB200_S_OK;
}

// canny_dete.cxx:77-117
int cvb200_edge_dete_set(cvb200_edge_dete_t* d, int id, const void* valuePtr, size_t valueSize)
{
	CVB_REQUIRE(d && valuePtr && valueSize, CVB200_E_INVALID_PARAMETER);
	if (id == CVB200_EDGE_SET_BOOL_GENERIC_KERNEL) {
		CVB_REQUIRE(valueSize == sizeof(bool), CVB200_E_INVALID_PARAMETER);
		d->genericKernel = *static_cast<const bool*>(valuePtr);
		return CVB200_S_OK;
	}
	if (id == CVB200_EDGE_SET_BOOL_X86_SSE41_GMAX_LANES) {
		CVB_REQUIRE(valueSize == sizeof(bool) && d->id != CVB200_CANNY_ID, CVB200_E_INVALID_PARAMETER);
		d->gmaxLanes = *static_cast<const bool*>(valuePtr);
		return CVB200_S_OK;
	}
	CVB_REQUIRE(d->id == CVB200_CANNY_ID, CVB200_E_NOT_IMPLEMENTED); // CompVCaps::set default (edge_dete.cxx:45-52)
	switch (id) {
	case CVB200_CANNY_SET_INT_THRESHOLD_TYPE: {
		CVB_REQUIRE(valueSize == sizeof(int32_t), CVB200_E_INVALID_PARAMETER);
		const int32_t t = *static_cast<const int32_t*>(valuePtr);
		CVB_REQUIRE(t == CVB200_CANNY_THRESHOLD_TYPE_PERCENT_OF_MEAN || t == CVB200_CANNY_THRESHOLD_TYPE_COMPARE_TO_GRADIENT, CVB200_E_INVALID_PARAMETER);
		d->thresholdType = t;
		return CVB200_S_OK;
	}
	case CVB200_CANNY_SET_FLT32_THRESHOLD_LOW: {
		CVB_REQUIRE(valueSize == sizeof(float) && *static_cast<const float*>(valuePtr) > 0.f, CVB200_E_INVALID_PARAMETER);
		d->tLow = *static_cast<const float*>(valuePtr);
		return CVB200_S_OK;
	}
	case CVB200_CANNY_SET_FLT32_THRESHOLD_HIGH: {
		CVB_REQUIRE(valueSize == sizeof(float) && *static_cast<const float*>(valuePtr) > 0.f, CVB200_E_INVALID_PARAMETER);
		d->tHigh = *static_cast<const float*>(valuePtr);
		return CVB200_S_OK;
	}
	case CVB200_CANNY_SET_INT_KERNEL_SIZE: {
		CVB_REQUIRE(valueSize == sizeof(int), CVB200_E_INVALID_PARAMETER);
		const int k = *static_cast<const int*>(valuePtr);
		CVB_REQUIRE(k == 3 || k == 5, CVB200_E_INVALID_PARAMETER);
		return edge_taps(CVB200_CANNY_ID, static_cast<size_t>(k), &d->taps);
	}
	default:
		return CVB200_E_NOT_IMPLEMENTED;
	}
}

int cvb200_edge_dete_set_preblur(cvb200_edge_dete_t* d, size_t size, float sigma)
{
	CVB_REQUIRE(d, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(d->id == CVB200_CANNY_ID, CVB200_E_INVALID_CALL);
	CVB_REQUIRE(size == 0 || size == 3 || size == 5 || size == 7, CVB200_E_INVALID_PARAMETER);
	memset(&d->blur, 0, sizeof(d->blur));
	if (size) {
		CVB_CHECK(cvb200_gauss_kernel_dim1_32f(size, sigma, d->blur.k));
		d->blur.ks = static_cast<int>(size);
	}
	return CVB200_S_OK;
}

} // extern "C"

// All the kernels of one call, issued on `stream` without host interaction (the caller holds d->mutex).  edge_finish completes the call.
int cvb::edge_enqueue(cvb200_edge_dete* d, const uint8_t* image, size_t width, size_t height, size_t stride, uint8_t* edges, size_t batch, size_t framePitch, cudaStream_t stream)
{
	CVB_REQUIRE_INIT();
	const int stages = d->stages ? d->stages : 7; // 1 = front (or the gmax pass), 2 = hysteresis (or the normalisation pass), 4 = finalize: row-strip mode runs them apart
	CVB_REQUIRE(d && edges && (image || !(stages & 1) || d->id != CVB200_CANNY_ID) && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(width <= 0x3fffffff && height <= 0x3fffffff, CVB200_E_OUT_OF_BOUND);
	if (!framePitch) framePitch = stride * height;
	d->pendStream = stream; d->pendCheck = false;

	FrontParams p;
	memset(&p, 0, sizeof(p));
	p.in = image; p.cls = edges;
	p.W = static_cast<int>(width); p.H = static_cast<int>(height);
	p.stride = stride; p.framePitch = framePitch;
	p.taps = d->taps;

	if (d->id != CVB200_CANNY_ID) {
		// Sobel / Scharr / Prewitt: edge_dete.cxx:55-206. pass 1: gmax = max(1, max g); pass 2: u8(trunc(g * 255.f/gmax))
		CVB_REQUIRE(image != edges, CVB200_E_INVALID_PARAMETER);
		CVB_REQUIRE(static_cast<size_t>(d->taps.ks) <= width && static_cast<size_t>(d->taps.ks) <= height, CVB200_E_INVALID_PARAMETER);
		CVB_CHECK(d->counters.ensure(batch * sizeof(unsigned int)));
		unsigned int* gmax = d->externalGmax ? d->externalGmax : d->counters.as<unsigned int>();
		// `uint16_t gmax = 1` (edge_dete.cxx:93) then max over the frame: the normalisation pass uses max(gmax, 1)
		if (stages & 1) CVB_CUDA(cudaMemsetAsync(gmax, 0, batch * sizeof(unsigned int), stream));
		if (stages != 7) { // row-strip mode: the two passes apart, the frame maximum is all-reduced (MAX) between them by the caller
			p.gmaxLanes = d->gmaxLanes ? 1 : 0;
			if (stages & 1) { p.gmax = gmax; CVB_CHECK(launch_front(p, 2, batch, stream)); }
			if (stages & 2) { p.gmax = nullptr; p.gmaxIn = gmax; CVB_CHECK(launch_front(p, 3, batch, stream)); }
			return CVB200_S_OK;
		}
		if (d->id == CVB200_SOBEL_ID && d->taps.ks == 3 && !d->genericKernel) { // fast path (canny_fast.cuh): TMA-staged tile, 4 px per lane, both passes
			FastParams f;
			memset(&f, 0, sizeof(f));
			f.in = image; f.cls = edges; f.W = p.W; f.H = p.H; f.stride = stride; f.framePitch = framePitch;
			return launch_sobel_fast(f, gmax, d->gmaxLanes ? 1 : 0, batch, stream);
		}
		p.gmax = gmax;
		p.gmaxLanes = d->gmaxLanes ? 1 : 0;
		CVB_CHECK(launch_front(p, 2, batch, stream));
		p.gmax = nullptr; p.gmaxIn = gmax;
		CVB_CHECK(launch_front(p, 3, batch, stream));
		return CVB200_S_OK;
	}

	// ---- Canny ----
	CVB_REQUIRE(image != edges, CVB200_E_INVALID_PARAMETER); // in place only through the host entry point (staged)
	CVB_REQUIRE(d->tLow < d->tHigh, CVB200_E_INVALID_STATE); // canny_dete.cxx:126
	CVB_REQUIRE(static_cast<size_t>(d->taps.ks) <= width && static_cast<size_t>(d->taps.ks) <= height, CVB200_E_INVALID_PARAMETER);
	if (d->blur.ks) CVB_REQUIRE(static_cast<size_t>(d->blur.ks) <= width && static_cast<size_t>(d->blur.ks) <= height, CVB200_E_INVALID_PARAMETER);
	p.blur = d->blur;

	const int tilesX = static_cast<int>(div_up(width, HT)), tilesY = static_cast<int>(div_up(height, HT));
	const size_t nTiles = static_cast<size_t>(tilesX) * tilesY * batch;
	CVB_REQUIRE(nTiles <= 0x7fffffff, CVB200_E_OUT_OF_BOUND);
	const size_t countersBytes = batch * (sizeof(unsigned int) + sizeof(ushort2));
	CVB_CHECK(d->counters.ensure(countersBytes));
	CVB_CHECK(d->hostFlag.ensure(64 * sizeof(unsigned int)));
	unsigned int* sums = d->counters.as<unsigned int>();                 // [batch]
	ushort2* thr = reinterpret_cast<ushort2*>(sums + batch);             // [batch]

	if (d->thresholdType == CVB200_CANNY_THRESHOLD_TYPE_PERCENT_OF_MEAN) {
		// mean of the image handed to process() (canny_dete.cxx:243,253-258). With the fused pre-blur that image never exists in HBM:
		// not supported together (use cvb200_convlt1_8u32f8u_dev first).
		CVB_REQUIRE(!d->blur.ks, CVB200_E_NOT_IMPLEMENTED);
		CVB_CUDA(cudaMemsetAsync(sums, 0, batch * sizeof(unsigned int), stream));
		dim3 g(1, static_cast<unsigned>(height < 64 ? height : 64), static_cast<unsigned>(batch));
		frame_sum_kernel<<<g, 256, 0, stream>>>(image, p.W, p.H, stride, framePitch, sums);
		CVB_LAUNCHED();
		mean_thresholds_kernel<<<static_cast<unsigned>(div_up(batch, 128)), 128, 0, stream>>>(sums, static_cast<unsigned int>(width * height), d->tLow, d->tHigh, thr, static_cast<int>(batch));
		CVB_LAUNCHED();
		p.thr = thr;
	}
	else {
		// canny_dete.cxx:260-266
		float fl = d->tLow < 1.f ? 1.f : (d->tLow > 65535.f ? 65535.f : d->tLow);
		float fh = d->tHigh < 1.f ? 1.f : (d->tHigh > 65535.f ? 65535.f : d->tHigh);
		int tLow = static_cast<uint16_t>(fl), tHigh = static_cast<uint16_t>(fh);
		tLow = tLow < 1 ? 1 : tLow;
		tHigh = tHigh < tLow + 2 ? tLow + 2 : tHigh;
		p.tLow = tLow; p.tHigh = tHigh;
	}

	bool tapsOk = true; // the fast path drops the (then redundant) 0..255 clamp of the blur: needs non-negative taps summing to <= 1.003
	{
		float sum = 0.f;
		for (int i = 0; i < p.blur.ks; ++i) { if (!(p.blur.k[i] >= 0.f)) tapsOk = false; sum += p.blur.k[i]; }
		if (!(sum <= 1.003f)) tapsOk = false;
	}
	if (!(stages & 1)) { /* the class map is already in `edges` */ }
	else if (!d->genericKernel && tapsOk && p.taps.ks == 3 && (p.blur.ks == 0 || p.blur.ks == 3 || p.blur.ks == 5)) {
		// fast path (canny_fast.cuh): TMA-staged tile, 4 px per lane
		FastParams f;
		memset(&f, 0, sizeof(f));
		f.in = image; f.cls = edges; f.thr = p.thr;
		f.W = p.W; f.H = p.H; f.stride = stride; f.framePitch = framePitch;
		f.tLow = p.tLow; f.tHigh = p.tHigh;
		for (int i = 0; i < p.blur.ks; ++i) f.k[i] = p.blur.k[i];
		if (p.blur.ks == 0) CVB_CHECK(launch_canny_fast_t<0>(f, batch, stream));
		else if (p.blur.ks == 3) CVB_CHECK(launch_canny_fast_t<3>(f, batch, stream));
		else CVB_CHECK(launch_canny_fast_t<5>(f, batch, stream));
	}
	else {
		CVB_CHECK(launch_front(p, 0, batch, stream));
	}

	// hysteresis: round 0 over every tile, then d->hystRounds list-driven rounds, all issued without host interaction.  Whether the last round still queued
	// something (not converged: only for edges that snake through more tiles than there were rounds) is read back with the results; edge_finish then
	// raises the number of rounds and the call is run again.
	HystParams h;
	h.cls = edges; h.W = p.W; h.H = p.H; h.stride = stride; h.framePitch = framePitch;
	h.tilesX = tilesX; h.tilesY = tilesY; h.nTiles = static_cast<int>(nTiles);
	const int rounds = d->hystRounds;
	if (stages & 2) {
	CVB_CHECK(d->dirty.ensure(nTiles * 3 * sizeof(int) + (static_cast<size_t>(rounds) + 2) * sizeof(unsigned int)));
	int* epoch = d->dirty.as<int>();
	int* lists[2] = { epoch + nTiles, epoch + 2 * nTiles };
	unsigned int* roundCount = reinterpret_cast<unsigned int*>(epoch + 3 * nTiles); // [rounds + 2]: roundCount[r] = tiles queued for round r
	CVB_CUDA(cudaMemsetAsync(epoch, 0, nTiles * sizeof(int), stream));
	CVB_CUDA(cudaMemsetAsync(roundCount, 0, (static_cast<size_t>(rounds) + 2) * sizeof(unsigned int), stream));
	h.epoch = epoch;
	for (int round = 0; round <= rounds; ++round) {
		h.round = round;
		h.listIn = lists[round & 1]; h.countIn = roundCount + round;
		h.listOut = lists[(round + 1) & 1]; h.countOut = roundCount + round + 1;
		KernelScope ks_("canny_hysteresis", stream);
		if (round == 0) canny_hysteresis_kernel<true><<<static_cast<unsigned>(nTiles), HT_THREADS, 0, stream>>>(h);
		else canny_hysteresis_kernel<false><<<static_cast<unsigned>(std::min<size_t>(nTiles, static_cast<size_t>(num_sms()) * 4)), HT_THREADS, 0, stream>>>(h);
		CVB_LAUNCHED();
	}
	CVB_CUDA(cudaMemcpyAsync(d->hostFlag.p, roundCount + rounds + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
	d->pendCheck = true;
	}
	if (!(stages & 4)) return CVB200_S_OK;

	dim3 fg(static_cast<unsigned>(div_up(div_up(width, 16), 128)), static_cast<unsigned>(height), static_cast<unsigned>(batch));
	CVB_REQUIRE(fg.y <= 65535, CVB200_E_OUT_OF_BOUND);
	if (d->khtBits) {
		dim3 bg(static_cast<unsigned>(div_up(div_up(width, 32), 32)), static_cast<unsigned>(div_up(height, 8)), static_cast<unsigned>(batch));
		KernelScope ks_("canny_finalize", stream);
		canny_finalize_bits_kernel<<<bg, dim3(32, 8), 0, stream>>>(edges, p.W, p.H, stride, framePitch, d->khtBits, d->khtWW, 2, d->khtEdgeCount);
	}
	else {
		KernelScope ks_("canny_finalize", stream);
		canny_finalize_kernel<<<fg, 128, 0, stream>>>(edges, p.W, p.H, stride, framePitch);
	}
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

// Waits for edge_enqueue's work.  *again = the hysteresis had not converged within the rounds issued: their number has been raised, enqueue again.
int cvb::edge_finish(cvb200_edge_dete* d, bool* again)
{
	*again = false;
	CVB_CUDA(cudaStreamSynchronize(d->pendStream));
	if (d->pendCheck) {
		d->pendCheck = false;
		if (d->hostFlag.as<unsigned int>()[0] != 0) {
			CVB_REQUIRE(d->hystRounds < (1 << 20), CVB200_E_INVALID_STATE);
			d->hystRounds *= 4;
			*again = true;
		}
	}
	return CVB200_S_OK;
}

extern "C" {

// Row-strip mode: the stages of one detector call apart, so that a caller that owns a strip of a frame (plus halo rows) can exchange what crosses the
// seams between them.  Canny: 1 = Gaussian/Sobel/NMS -> class map (0, 0x80 weak, 0xff strong) in `edges`; 2 = 8-connected closure of the strong pixels inside `edges` as it
// stands (halo rows received from the neighbours included); 4 = weak -> 0.  Sobel/Scharr/Prewitt: 1 = frame maximum of |gx|+|gy| into gmax[frame] (device u32),
// 2 = normalisation with gmax[frame] as given (after the caller's MAX all-reduce).  Synchronous.
int cvb200_edge_dete_process_stages_dev(cvb200_edge_dete_t* d, const uint8_t* image, size_t width, size_t height, size_t stride, uint8_t* edges,
	size_t batch, size_t framePitch, int stages, uint32_t* gmax, cvb200_stream_t stream_)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(d && edges && width && height && stride >= width && stages > 0 && stages <= 7, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(d->id == CVB200_CANNY_ID || (gmax && image && stages <= 3), CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	std::lock_guard<std::mutex> lock(d->mutex);
	d->stages = stages; d->externalGmax = gmax;
	int rc = CVB200_S_OK;
	for (int attempt = 0; attempt < 12; ++attempt) {
		rc = edge_enqueue(d, image, width, height, stride, edges, batch, framePitch, as_stream(stream_));
		bool again = false;
		if (rc == CVB200_S_OK) rc = edge_finish(d, &again);
		if (rc != CVB200_S_OK || !again) break;
		// hysteresis did not converge within the rounds issued: promotions are monotone, so simply continue from the map as it stands (never re-run the front here)
		d->stages = stages & ~1;
		if (attempt == 11) rc = CVB200_E_INVALID_STATE;
	}
	d->stages = 0; d->externalGmax = nullptr;
	return rc;
}

int cvb200_edge_dete_process_dev(cvb200_edge_dete_t* d, const uint8_t* image, size_t width, size_t height, size_t stride, uint8_t* edges,
	size_t batch, size_t framePitch, cvb200_stream_t stream_)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(d && image && edges && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	std::lock_guard<std::mutex> lock(d->mutex);
	for (int attempt = 0; attempt < 12; ++attempt) {
		CVB_CHECK(edge_enqueue(d, image, width, height, stride, edges, batch, framePitch, as_stream(stream_)));
		bool again = false;
		CVB_CHECK(edge_finish(d, &again));
		if (!again) return CVB200_S_OK;
	}
	return CVB200_E_INVALID_STATE;
}

int cvb200_edge_dete_process(cvb200_edge_dete_t* d, const uint8_t* image, size_t width, size_t height, size_t stride, uint8_t* edges)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(d && image && edges && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	const size_t n = stride * height;
	// the object's staging buffers are used for the whole call: one caller at a time per object (the reference's detector objects are not thread-safe either)
	std::lock_guard<std::mutex> lock(d->mutex);
	CVB_CHECK(d->hostIn.ensure(n));
	CVB_CHECK(d->hostOut.ensure(n));
	// row by row: the caller's last row need not be padded to the stride
	CVB_CUDA(cudaMemcpy2DAsync(d->hostIn.p, stride, image, stride, width, height, cudaMemcpyHostToDevice, 0));
	int rc = CVB200_S_OK;
	for (int attempt = 0; attempt < 12; ++attempt) {
		rc = edge_enqueue(d, d->hostIn.as<uint8_t>(), width, height, stride, d->hostOut.as<uint8_t>(), 1, 0, nullptr);
		bool again = false;
		if (rc == CVB200_S_OK) rc = edge_finish(d, &again);
		if (rc != CVB200_S_OK || !again) break;
		if (attempt == 11) rc = CVB200_E_INVALID_STATE;
	}
	if (rc != CVB200_S_OK) { cudaStreamSynchronize(0); return rc; }
	CVB_CUDA(cudaMemcpy2DAsync(edges, stride, d->hostOut.p, stride, width, height, cudaMemcpyDeviceToHost, 0));
	CVB_CUDA(cudaStreamSynchronize(0));
	return CVB200_S_OK;
}

// Host buffers, many frames: the reference-facing call for throughput use.  Frames are cut into chunks that flow through a
// 3-deep ring (pinned or pageable host memory -> H2D on a copy stream -> kernels on the compute stream -> D2H on a copy stream), so
// the PCIe transfers of neighbouring chunks overlap the kernels of the current one.
int cvb200_edge_dete_process_batch(cvb200_edge_dete_t* d, const uint8_t* image, size_t width, size_t height, size_t stride, uint8_t* edges,
	size_t batch, size_t framePitch)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(d && image && edges && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	if (!framePitch) framePitch = stride * height;
	CVB_REQUIRE(framePitch >= stride * height, CVB200_E_INVALID_PARAMETER);
	const size_t frameBytes = stride * height;
	size_t chunk = (8u << 20) / frameBytes; // ~8 MiB per chunk
	if (chunk < 1) chunk = 1;
	if (chunk > batch) chunk = batch;
	constexpr int RING = 3;
	static thread_local cudaStream_t sIn = nullptr, sCompute = nullptr, sOut = nullptr;
	static thread_local cudaEvent_t evIn[RING], evDone[RING], evOut[RING];
	if (!sIn) {
		CVB_CUDA(cudaStreamCreateWithFlags(&sIn, cudaStreamNonBlocking));
		CVB_CUDA(cudaStreamCreateWithFlags(&sCompute, cudaStreamNonBlocking));
		CVB_CUDA(cudaStreamCreateWithFlags(&sOut, cudaStreamNonBlocking));
		for (int i = 0; i < RING; ++i) {
			CVB_CUDA(cudaEventCreateWithFlags(&evIn[i], cudaEventDisableTiming));
			CVB_CUDA(cudaEventCreateWithFlags(&evDone[i], cudaEventDisableTiming));
			CVB_CUDA(cudaEventCreateWithFlags(&evOut[i], cudaEventDisableTiming));
		}
	}
	{
		std::lock_guard<std::mutex> lock(d->mutex);
		CVB_CHECK(d->hostIn.ensure(RING * chunk * frameBytes));
		CVB_CHECK(d->hostOut.ensure(RING * chunk * frameBytes));
	}
	uint8_t* dIn = d->hostIn.as<uint8_t>();
	uint8_t* dOut = d->hostOut.as<uint8_t>();
	const size_t nChunks = div_up(batch, chunk);
	auto issue_h2d = [&](size_t c) -> int {
		const int slot = static_cast<int>(c % RING);
		const size_t f0 = c * chunk, nf = (f0 + chunk <= batch) ? chunk : (batch - f0);
		if (c >= RING) CVB_CUDA(cudaStreamWaitEvent(sIn, evDone[slot], 0)); // slot's previous kernels have consumed the input
		CVB_CUDA(cudaMemcpy2DAsync(dIn + slot * chunk * frameBytes, frameBytes, image + f0 * framePitch, framePitch, frameBytes, nf, cudaMemcpyHostToDevice, sIn));
		CVB_CUDA(cudaEventRecord(evIn[slot], sIn));
		return CVB200_S_OK;
	};
	CVB_CHECK(issue_h2d(0));
	for (size_t c = 0; c < nChunks; ++c) {
		const int slot = static_cast<int>(c % RING);
		const size_t f0 = c * chunk, nf = (f0 + chunk <= batch) ? chunk : (batch - f0);
		if (c + 1 < nChunks) CVB_CHECK(issue_h2d(c + 1)); // prefetch the next chunk before this one's kernels block the host
		CVB_CUDA(cudaStreamWaitEvent(sCompute, evIn[slot], 0));
		if (c >= RING) CVB_CUDA(cudaStreamWaitEvent(sCompute, evOut[slot], 0)); // slot's previous result has left the device
		CVB_CHECK(cvb200_edge_dete_process_dev(d, dIn + slot * chunk * frameBytes, width, height, stride, dOut + slot * chunk * frameBytes, nf, frameBytes,
			reinterpret_cast<cvb200_stream_t>(sCompute)));
		CVB_CUDA(cudaEventRecord(evDone[slot], sCompute));
		CVB_CUDA(cudaStreamWaitEvent(sOut, evDone[slot], 0));
		CVB_CUDA(cudaMemcpy2DAsync(edges + f0 * framePitch, framePitch, dOut + slot * chunk * frameBytes, frameBytes, frameBytes, nf, cudaMemcpyDeviceToHost, sOut));
		CVB_CUDA(cudaEventRecord(evOut[slot], sOut));
	}
	CVB_CUDA(cudaStreamSynchronize(sOut));
	return CVB200_S_OK;
}

int cvb200_sobel_g_dev(const uint8_t* image, size_t width, size_t height, size_t stride, int id, size_t kernSize, int16_t* gx, int16_t* gy, uint16_t* g,
	size_t batch, size_t framePitch, cvb200_stream_t stream)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(image && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(kernSize == 3 || (kernSize == 5 && (id == CVB200_SOBEL_ID || id == CVB200_CANNY_ID)), CVB200_E_INVALID_PARAMETER);
	CVB_REQUIRE(width >= kernSize && height >= kernSize, CVB200_E_INVALID_PARAMETER);
	if (!batch) return CVB200_S_OK;
	FrontParams p;
	memset(&p, 0, sizeof(p));
	CVB_CHECK(edge_taps(id, kernSize, &p.taps));
	p.in = image; p.gxOut = gx; p.gyOut = gy; p.gOut = g;
	p.W = static_cast<int>(width); p.H = static_cast<int>(height);
	p.stride = stride; p.framePitch = framePitch ? framePitch : stride * height;
	return launch_front(p, 1, batch, as_stream(stream));
}

int cvb200_sobel_g(const uint8_t* image, size_t width, size_t height, size_t stride, int id, size_t kernSize, int16_t* gx, int16_t* gy, uint16_t* g)
{
	CVB_REQUIRE_INIT();
	CVB_REQUIRE(image && width && height && stride >= width, CVB200_E_INVALID_PARAMETER);
	const size_t n = stride * height;
	DevBuf dIn, dGx, dGy, dG;
	int rc = dIn.ensure(n);
	if (!rc) rc = dGx.ensure(n * 2);
	if (!rc) rc = dGy.ensure(n * 2);
	if (!rc) rc = dG.ensure(n * 2);
	if (!rc) rc = cvb200_memcpy_h2d(dIn.p, image, n, nullptr);
	if (!rc) rc = cvb200_sobel_g_dev(dIn.as<uint8_t>(), width, height, stride, id, kernSize, dGx.as<int16_t>(), dGy.as<int16_t>(), dG.as<uint16_t>(), 1, 0, nullptr);
	if (!rc && gx) rc = cvb200_memcpy_d2h(gx, dGx.p, n * 2, nullptr);
	if (!rc && gy) rc = cvb200_memcpy_d2h(gy, dGy.p, n * 2, nullptr);
	if (!rc && g) rc = cvb200_memcpy_d2h(g, dG.p, n * 2, nullptr);
	if (!rc) rc = cvb200_stream_sync(nullptr);
	dIn.release(); dGx.release(); dGy.release(); dG.release();
	return rc;
}

} // extern "C"
