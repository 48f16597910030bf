// TMA (cp.async.bulk.tensor) + mbarrier helpers shared by the tile-staging kernels (sm_100a).
#pragma once

#include <cuda.h>

namespace cvb {

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase)
{
	const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(bar));
	unsigned ok;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(phase) : "memory");
	} while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* smemDst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2)
{
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
		:: "r"(static_cast<unsigned>(__cvta_generic_to_shared(smemDst))), "l"(reinterpret_cast<uint64_t>(map)),
		   "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(c0), "r"(c1), "r"(c2)
		: "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
	CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled()
{
	static PFN_encodeTiled fn = nullptr;
	static bool tried = false;
	if (!tried) {
		tried = true;
		void* p = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
		else cudaGetLastError();
	}
	return fn;
}

// u8 frames as a 3-D tensor {W, H, batch} with byte strides {stride, framePitch}; box = {boxBytes, boxRows, 1}.
// TMA constraints: base and strides 16-byte aligned; the innermost start coordinate must be a multiple of 16 bytes (callers round down).
static bool make_u8_tile_map(CUtensorMap* map, const uint8_t* base, size_t W, size_t H, size_t stride, size_t framePitch, size_t batch, int boxBytes, int boxRows)
{
	PFN_encodeTiled enc = get_encode_tiled();
	if (!enc) return false;
	if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride & 15) || (framePitch & 15)) return false;
	const cuuint64_t dims[3] = { W, H, batch };
	const cuuint64_t strides[2] = { stride, framePitch };
	const cuuint32_t box[3] = { static_cast<cuuint32_t>(boxBytes), static_cast<cuuint32_t>(boxRows), 1 };
	const cuuint32_t estr[3] = { 1, 1, 1 };
	return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
		CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace cvb
