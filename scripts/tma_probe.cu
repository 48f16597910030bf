// Stand-alone probe used while bringing up the TMA tile loader (not part of the product build).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe scripts/tma_probe.cu && /tmp/tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
	CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, int useGlobal, int rows, int c0, int c1, unsigned* out, int boxw)
{
	extern __shared__ __align__(128) unsigned char smem[];
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + rows * 256);
	const unsigned barAddr = static_cast<unsigned>(__cvta_generic_to_shared(bar));
	const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem));
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(barAddr));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(barAddr), "r"(rows * boxw) : "memory");
		const uint64_t desc = useGlobal ? reinterpret_cast<uint64_t>(gmap) : reinterpret_cast<uint64_t>(&tmap);
		if (RANK == 2) {
			asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
				:: "r"(dst), "l"(desc), "r"(barAddr), "r"(c0), "r"(c1) : "memory");
		}
		else {
			asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
				:: "r"(dst), "l"(desc), "r"(barAddr), "r"(c0), "r"(c1), "r"(0) : "memory");
		}
	}
	__syncthreads();
	unsigned ok;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(barAddr) : "memory");
	} while (!ok);
	unsigned sum = 0;
	for (int i = threadIdx.x; i < rows * boxw; i += blockDim.x) sum += smem[i];
	atomicAdd(out, sum);
}

int main(int argc, char** argv)
{
	const int variant = argc > 1 ? atoi(argv[1]) : 0;
	const int rank = (variant & 1) ? 2 : 3;
	const int useGlobal = (variant & 2) ? 1 : 0;
	const int W = 1920, H = 1080, rows = 64;
	uint8_t* img;
	cudaMalloc(&img, W * H);
	cudaMemset(img, 1, W * H);
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult q;
	cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
	printf("variant %d rank %d global %d: entry point err=%d q=%d fn=%p\n", variant, rank, useGlobal, (int)e, (int)q, fn);
	alignas(64) CUtensorMap map;
	memset(&map, 0, sizeof(map));
	const cuuint64_t dims[3] = { (cuuint64_t)W, (cuuint64_t)H, 1 };
	const cuuint64_t strides[2] = { (cuuint64_t)W, (cuuint64_t)W * H };
	const cuuint32_t box[3] = { (cuuint32_t)(argc > 4 ? atoi(argv[4]) : 128), (cuuint32_t)rows, 1 };
	const cuuint32_t estr[3] = { 1, 1, 1 };
	CUresult r = reinterpret_cast<PFN_encodeTiled>(fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
		CU_TENSOR_MAP_SWIZZLE_NONE, (variant & 4) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode result %d\n", (int)r);
	const unsigned long long* w = reinterpret_cast<const unsigned long long*>(&map);
	for (int i = 0; i < 16; ++i) printf("%016llx%s", w[i], (i % 4 == 3) ? "\n" : " ");
	CUtensorMap* gmap;
	cudaMalloc(&gmap, sizeof(map));
	cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
	unsigned* out;
	cudaMalloc(&out, 4);
	cudaMemset(out, 0, 4);
	int c0 = (variant & 8) ? 0 : -4, c1 = (variant & 8) ? 0 : -2;
	if (argc > 3) { c0 = atoi(argv[2]); c1 = atoi(argv[3]); }
	if (rank == 2) probe<2><<<1, 128, rows * 256 + 64>>>(map, gmap, useGlobal, rows, c0, c1, out, (int)box[0]);
	else probe<3><<<1, 128, rows * 256 + 64>>>(map, gmap, useGlobal, rows, c0, c1, out, (int)box[0]);
	e = cudaDeviceSynchronize();
	unsigned h = 0;
	cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
	printf("sync: %s, sum=%u (expect %d)\n", cudaGetErrorString(e), h, (variant & 8) ? 128 * rows : 124 * 62);
	return 0;
}
