// Fast path of CompVMathConvlt::convlt1<uint8_t, float, uint8_t> (the Gaussian blur of the text / edge pipelines) for 3- and 5-tap kernels.
// Same arithmetic contract as convlt1_kernel in convlt.cu (float FMA chain in tap order starting from 0, clamp to 0..255, truncate; u8 intermediate between the
// horizontal and the vertical pass; the reference's three border types), restructured like canny_front_fast_kernel:
//   * the (120 + 24) x (60 + 2r) u8 input tile lands in shared memory through ONE TMA bulk-tensor copy per CTA (out-of-image elements zero-filled by the hardware);
//   * every lane owns 4 consecutive pixels (one 32-bit shared-memory word per row): all shared-memory traffic is 32-bit and conflict free;
//   * u8 <-> f32 conversions use the 2^23 magic-number trick (PRMT + FADD / FADD.RZ) instead of the quarter-rate I2F / F2I pipe.
// Frames whose base / stride / pitch are not 16-byte aligned cannot be described to the TMA unit: they take the generic kernel.
#pragma once

#include "tma.cuh"
#include "packed.cuh"

namespace cvb {

constexpr int VF_TW = 120, VF_TH = 60, VF_THREADS = 256, VF_WARPS = 8;
constexpr int VF_ROWW = 32;   // words per row of the intermediate tile
constexpr int VF_INW = 36;    // words per row of the staged input tile (144-byte TMA box: the innermost coordinate must be a multiple of 16 bytes)

struct ConvFastParams {
	uint8_t* out;
	int W, H;
	size_t stride, framePitch;
	float vt[5], hz[5];
	int border;
	int vecStore; // out rows are 4-byte aligned
};

__device__ __forceinline__ float vf_u8_to_f32(unsigned int w, int byteIdx)
{
	return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | byteIdx)) - 8388608.f; // [b, 0, 0, 0x4B] = 2^23 + b
}
// clamp to 0..255 then truncate (compv_math_convlt.h:358-384 with OutputType = uint8_t): the low byte of the returned pattern
__device__ __forceinline__ unsigned int vf_f32_to_u8_bits(float s)
{
	return __float_as_uint(__fadd_rz(fminf(fmaxf(s, 0.f), 255.f), 8388608.f));
}
__device__ __forceinline__ unsigned int vf_pack4(unsigned int a, unsigned int b, unsigned int c, unsigned int d)
{
	return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

template <int KS, bool CLAMP>
__global__ void __launch_bounds__(VF_THREADS, 3)
convlt_fast_8u32f8u_kernel(const __grid_constant__ CUtensorMap tmap, const ConvFastParams p)
{
	constexpr int R = KS >> 1;
	constexpr int IN_ROWS = VF_TH + 2 * R;
	extern __shared__ __align__(128) unsigned char vf_smem[];
	const unsigned int pad = (128u - (static_cast<unsigned int>(__cvta_generic_to_shared(vf_smem)) & 127u)) & 127u;
	unsigned int* sA = reinterpret_cast<unsigned int*>(vf_smem + pad);          // staged input rows (pitch VF_INW), 128-byte aligned
	unsigned int* sM = sA + IN_ROWS * VF_INW + 4;                                 // horizontal pass output (pitch VF_ROWW)
	uint64_t* bar = reinterpret_cast<uint64_t*>(sM + IN_ROWS * VF_ROWW + 2);

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int W = p.W, H = p.H;
	const int x0 = blockIdx.x * VF_TW, y0 = blockIdx.y * VF_TH;
	const int frame = blockIdx.z;
	const int xl = x0 - 4 + 4 * lane;                     // first of my 4 columns; lanes 1..30 own output columns
	const int yIn0 = y0 - R;                              // image row of staged row 0
	const int xTma = (x0 - 4) & ~15;
	const int woff = ((x0 - 4) - xTma) >> 2;              // word of lane 0 inside a staged row (1 or 3)

	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		mbar_expect_tx(bar, IN_ROWS * VF_INW * 4);
		tma_load_3d(sA, &tmap, bar, xTma, yIn0, frame);
	}
	__syncthreads();
	mbar_wait(bar, 0);

	float hz[KS], vt[KS];
#pragma unroll
	for (int k = 0; k < KS; ++k) { hz[k] = p.hz[k]; vt[k] = p.vt[k]; }

	// ---- horizontal pass (compv_math_convlt.h:176-229): mid(y, x) = conv for x in [R, W-R); on the R-wide column border 0, or the input sample for REPLICATE ----
	unsigned int convMask = 0;
#pragma unroll
	for (int i = 0; i < 4; ++i) if (xl + i >= R && xl + i < W - R) convMask |= 0xffu << (8 * i);
	if (!CLAMP) {
		// Non-negative taps summing to <= 1.003 (a normalised Gaussian): 0 <= s < 256, the clamp is a no-op.  Two rows per step, the float pair is (row r, row r+1) of
		// one column: FFMA2 / FADD2 round each half exactly like the scalar instruction.  Rows outside the image are staged as zeros and blur to zero by themselves.
		static_assert(IN_ROWS % 2 == 0, "rows are filtered in pairs");
		const f32x2 negMagic = pk2(0xCB000000u, 0xCB000000u), magic = pk2(0x4B000000u, 0x4B000000u);
		f32x2 kk[KS];
#pragma unroll
		for (int k = 0; k < KS; ++k) kk[k] = pk2(__float_as_uint(hz[k]), __float_as_uint(hz[k]));
		for (int r = 2 * warp; r < IN_ROWS; r += 2 * VF_WARPS) {
			const unsigned int* q0 = &sA[r * VF_INW + woff + lane];
			const unsigned int* q1 = q0 + VF_INW;
			const unsigned int l0 = q0[-1], c0 = q0[0], r0 = q0[1], l1 = q1[-1], c1 = q1[0], r1 = q1[1];
			unsigned int out0 = 0, out1 = 0;
			if (convMask) {
				f32x2 v[4 + 2 * R];
#pragma unroll
				for (int j = 0; j < R; ++j) v[j] = u8x2_to_f32x2(l0, 4 - R + j, l1, 4 - R + j, negMagic);
#pragma unroll
				for (int j = 0; j < 4; ++j) v[R + j] = u8x2_to_f32x2(c0, j, c1, j, negMagic);
#pragma unroll
				for (int j = 0; j < R; ++j) v[R + 4 + j] = u8x2_to_f32x2(r0, j, r1, j, negMagic);
				unsigned int oa[4], ob[4];
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					f32x2 s = fmul2(v[i], kk[0]); // == fma(v, k, 0)
#pragma unroll
					for (int k = 1; k < KS; ++k) s = ffma2(v[i + k], kk[k], s);
					unpk2(fadd2_rz(s, magic), oa[i], ob[i]);
				}
				out0 = vf_pack4(oa[0], oa[1], oa[2], oa[3]) & convMask;
				out1 = vf_pack4(ob[0], ob[1], ob[2], ob[3]) & convMask;
			}
			if (p.border == CVB200_BORDER_TYPE_REPLICATE) { out0 |= c0 & ~convMask; out1 |= c1 & ~convMask; } // columns / rows outside the image were zero-filled by the TMA unit
			sM[r * VF_ROWW + lane] = out0;
			sM[(r + 1) * VF_ROWW + lane] = out1;
		}
	}
	else
	for (int r = warp; r < IN_ROWS; r += VF_WARPS) {
		const int y = yIn0 + r;
		unsigned int outw = 0;
		if (y >= 0 && y < H) {
			const unsigned int* q = &sA[r * VF_INW + woff + lane];
			const unsigned int wl = q[-1], wc = q[0], wr = q[1];
			if (convMask) {
				float v[4 + 2 * R];
#pragma unroll
				for (int j = 0; j < R; ++j) v[j] = vf_u8_to_f32(wl, 4 - R + j);
#pragma unroll
				for (int j = 0; j < 4; ++j) v[R + j] = vf_u8_to_f32(wc, j);
#pragma unroll
				for (int j = 0; j < R; ++j) v[R + 4 + j] = vf_u8_to_f32(wr, j);
				unsigned int o[4];
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					float s = 0.f;
#pragma unroll
					for (int k = 0; k < KS; ++k) s = __fmaf_rn(v[i + k], hz[k], s);
					o[i] = vf_f32_to_u8_bits(s);
				}
				outw = vf_pack4(o[0], o[1], o[2], o[3]) & convMask;
			}
			if (p.border == CVB200_BORDER_TYPE_REPLICATE) outw |= wc & ~convMask; // columns outside the image were zero-filled by the TMA unit
		}
		sM[r * VF_ROWW + lane] = outw;
	}
	__syncthreads();

	// ---- vertical pass (compv_math_convlt.h:231-292) ----
	{
		constexpr int RPW = (VF_TH + VF_WARPS - 1) / VF_WARPS; // 8 output rows per warp
		const int ro0 = warp * RPW;
		float win[CLAMP ? RPW + 2 * R : 1][4];
		f32x2 win2[CLAMP ? 1 : RPW + 2 * R][2], vk[KS];
#pragma unroll
		for (int k = 0; k < KS; ++k) vk[k] = pk2(__float_as_uint(vt[k]), __float_as_uint(vt[k]));
#pragma unroll
		for (int r = 0; r < RPW + 2 * R; ++r) {
			const int rr = min(ro0 + r, IN_ROWS - 1);
			const unsigned int w = sM[rr * VF_ROWW + lane];
			if (!CLAMP) {
				const f32x2 negMagic = pk2(0xCB000000u, 0xCB000000u);
				win2[r][0] = u8x2_to_f32x2(w, 0, w, 1, negMagic);
				win2[r][1] = u8x2_to_f32x2(w, 2, w, 3, negMagic);
			}
			else {
#pragma unroll
				for (int i = 0; i < 4; ++i) win[r][i] = vf_u8_to_f32(w, i);
			}
		}
		const bool laneOut = (lane >= 1 && lane <= 30) && xl < W;
		uint8_t* __restrict__ out = p.out + frame * p.framePitch;
#pragma unroll
		for (int j = 0; j < RPW; ++j) {
			const int ro = ro0 + j;
			const int y = y0 + ro;
			if (ro >= VF_TH || y >= H || !laneOut) continue;
			unsigned int outw = 0, storeMask = 0;
#pragma unroll
			for (int i = 0; i < 4; ++i) if (xl + i < W) storeMask |= 0xffu << (8 * i);
			if (y >= R && y < H - R) {
				unsigned int o[4];
				if (!CLAMP) { // column pairs (0, 1) and (2, 3) of the lane's word
#pragma unroll
					for (int h = 0; h < 2; ++h) {
						f32x2 s = fmul2(win2[j][h], vk[0]);
#pragma unroll
						for (int k = 1; k < KS; ++k) s = ffma2(win2[j + k][h], vk[k], s);
						unpk2(fadd2_rz(s, pk2(0x4B000000u, 0x4B000000u)), o[2 * h], o[2 * h + 1]);
					}
				}
				else {
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						float s = 0.f;
#pragma unroll
						for (int k = 0; k < KS; ++k) s = __fmaf_rn(win[j + k][i], vt[k], s);
						o[i] = vf_f32_to_u8_bits(s);
					}
				}
				outw = vf_pack4(o[0], o[1], o[2], o[3]);
				if (p.border == CVB200_BORDER_TYPE_IGNORE) storeMask &= convMask; // the column border is left as is
			}
			else if (p.border == CVB200_BORDER_TYPE_ZERO) outw = 0;
			else if (p.border == CVB200_BORDER_TYPE_REPLICATE) outw = sM[(ro + R) * VF_ROWW + lane]; // the intermediate sample (generic kernel: sMid[(ly + r) ...])
			else storeMask = 0; // IGNORE on the row border
			if (!storeMask) continue;
			uint8_t* o8 = out + static_cast<size_t>(y) * p.stride + xl;
			if (p.vecStore && storeMask == 0xffffffffu) *reinterpret_cast<unsigned int*>(o8) = outw;
			else {
#pragma unroll
				for (int i = 0; i < 4; ++i) if ((storeMask >> (8 * i)) & 0xffu) o8[i] = static_cast<uint8_t>(outw >> (8 * i));
			}
		}
	}
}

// returns CVB200_S_OK when the fast path ran, 1 when the frames cannot be described to the TMA unit (the caller falls back)
template <int KS, bool CLAMP>
static int launch_convlt_fast_t(const uint8_t* in, uint8_t* out, size_t W, size_t H, size_t stride, size_t framePitch, const float* vt, const float* hz, int border, size_t batch, cudaStream_t stream)
{
	constexpr int R = KS >> 1;
	constexpr int IN_ROWS = VF_TH + 2 * R;
	alignas(64) CUtensorMap map;
	memset(&map, 0, sizeof(map));
	if (!make_u8_tile_map(&map, in, W, H, stride, framePitch, batch, VF_INW * 4, IN_ROWS)) return 1;
	ConvFastParams p;
	memset(&p, 0, sizeof(p));
	p.out = out; p.W = static_cast<int>(W); p.H = static_cast<int>(H); p.stride = stride; p.framePitch = framePitch; p.border = border;
	for (int k = 0; k < KS; ++k) { p.vt[k] = vt[k]; p.hz[k] = hz[k]; }
	p.vecStore = (((reinterpret_cast<uintptr_t>(out) | stride | framePitch) & 3) == 0) ? 1 : 0;
	const size_t smem = (static_cast<size_t>(IN_ROWS) * (VF_INW + VF_ROWW) + 8) * 4 + 128 + 16;
	auto kern = convlt_fast_8u32f8u_kernel<KS, CLAMP>;
	dim3 grid(static_cast<unsigned>(div_up(W, VF_TW)), static_cast<unsigned>(div_up(H, VF_TH)), static_cast<unsigned>(batch));
	CVB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CVB200_E_OUT_OF_BOUND);
	{
		KernelScope ks_("convlt1", stream);
		kern<<<grid, VF_THREADS, smem, stream>>>(map, p);
	}
	CVB_LAUNCHED();
	return CVB200_S_OK;
}

// the clamp to 0..255 is a no-op for non-negative taps summing to <= 1.003 (every normalised Gaussian): those take the packed (FFMA2) instance
template <int KS>
static int launch_convlt_fast(const uint8_t* in, uint8_t* out, size_t W, size_t H, size_t stride, size_t framePitch, const float* vt, const float* hz, int border, size_t batch, cudaStream_t stream)
{
	bool noClamp = true;
	float sv = 0.f, sh = 0.f;
	for (int k = 0; k < KS; ++k) { if (!(vt[k] >= 0.f) || !(hz[k] >= 0.f)) noClamp = false; sv += vt[k]; sh += hz[k]; }
	if (!(sv <= 1.003f) || !(sh <= 1.003f)) noClamp = false;
	return noClamp ? launch_convlt_fast_t<KS, false>(in, out, W, H, stride, framePitch, vt, hz, border, batch, stream)
	               : launch_convlt_fast_t<KS, true>(in, out, W, H, stride, framePitch, vt, hz, border, batch, stream);
}

} // namespace cvb
