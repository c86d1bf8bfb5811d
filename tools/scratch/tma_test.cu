// scratch: isolate the TMA 3-D box load
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef unsigned long long u64;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__constant__ CUtensorMap c_tmap;
__global__ void k(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, int c0, int c1, int c2, int bytes, uint8_t *out, int variant) {
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ u64 bar;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
		const CUtensorMap* mp = variant == 2 ? gmap : (variant == 3 ? &c_tmap : &tmap);
		if (variant == 2 || variant == 3)
			asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem)),
			             "l"(mp), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar))
			             : "memory");
		else if (variant == 0)
			asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem)),
			             "l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar))
			             : "memory");
		else
			asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem)),
			             "l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar))
			             : "memory");
	}
	asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
	for (int i = threadIdx.x; i < bytes; i += blockDim.x)
		out[i] = smem[i];
}
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
	int bx = atoi(argv[1]), bz = atoi(argv[2]), by = atoi(argv[3]), variant = atoi(argv[4]), X = atoi(argv[5]); int C0 = atoi(argv[6]), C1 = atoi(argv[7]), C2 = atoi(argv[8]);
	int Z = X, Y = X;
	uint8_t *g;
	cudaMalloc(&g, (size_t)X * Z * Y);
	uint8_t *h = (uint8_t *)malloc((size_t)X * Z * Y);
	for (size_t i = 0; i < (size_t)X * Z * Y; ++i)
		h[i] = (uint8_t)(i * 7 + (i >> 8));
	cudaMemcpy(g, h, (size_t)X * Z * Y, cudaMemcpyHostToDevice);
	void *fn = nullptr;
	cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
	CUtensorMap tm;
	cuuint64_t dims[3] = {(cuuint64_t)X, (cuuint64_t)Z, (cuuint64_t)Y}, strides[2] = {(cuuint64_t)X, (cuuint64_t)X * Z};
	cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)bz, (cuuint32_t)by}, es[3] = {1, 1, 1};
	CUresult r = ((PFN)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
	                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode box %dx%dx%d variant %d X %d -> %d\n", bx, bz, by, variant, X, (int)r);
	int bytes = bx * bz * by;
	uint8_t *out;
	cudaMalloc(&out, bytes);
	cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
	CUtensorMap* gm; cudaMalloc(&gm, 128); cudaMemcpy(gm,&tm,128,cudaMemcpyHostToDevice); cudaMemcpyToSymbol(c_tmap,&tm,128);
	k<<<1, 128, bytes + 128>>>(tm, gm, C0, C1, C2, bytes, out, variant);
	cudaError_t e = cudaDeviceSynchronize();
	printf("  kernel: %s\n", cudaGetErrorString(e));
	if (e == cudaSuccess) {
		uint8_t *ho = (uint8_t *)malloc(bytes);
		cudaMemcpy(ho, out, bytes, cudaMemcpyDeviceToHost);
		int bad = 0;
		for (int y = 0; y < by; ++y)
			for (int z = 0; z < bz; ++z)
				for (int x = 0; x < bx; ++x) {
					int gx = C0 + x, gz = C1 + z, gy = C2 + y;
					uint8_t want = (gx >= 0 && gz >= 0 && gy >= 0 && gx < X && gz < Z && gy < Y) ? h[((size_t)gy * Z + gz) * X + gx] : 0;
					bad += ho[(y * bz + z) * bx + x] != want;
				}
		printf("  mismatches %d\n", bad);
	}
	return 0;
}
