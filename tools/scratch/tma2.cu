#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef unsigned long long u64;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k2(const __grid_constant__ CUtensorMap tmap, int bytes, float *out) {
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ __align__(8) u64 bar;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem)),
		             "l"(&tmap), "r"(0), "r"(0), "r"(smem_u32(&bar))
		             : "memory");
	}
	asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
	for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x)
		out[i] = ((float *)smem)[i];
}
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
	float *g;
	cudaMalloc(&g, 256 * 256 * 4);
	cudaMemset(g, 0, 256 * 256 * 4);
	void *fn = nullptr;
	cudaDriverEntryPointQueryResult q;
	cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
	printf("entry %p q %d err %d\n", fn, (int)q, (int)ge);
	alignas(64) CUtensorMap tm;
	cuuint64_t dims[2] = {256, 256}, strides[1] = {1024};
	cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
	CUresult r = ((PFN)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
	                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode -> %d\n", (int)r);
	const uint32_t *w = (const uint32_t *)&tm;
	for (int i = 0; i < 32; ++i)
		printf("%08x%c", w[i], i % 8 == 7 ? '\n' : ' ');
	float *out;
	cudaMalloc(&out, 4096);
	k2<<<1, 128, 4096 + 128>>>(tm, 4096, out);
	printf("kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	int drv = 0, rt = 0;
	cudaDriverGetVersion(&drv);
	cudaRuntimeGetVersion(&rt);
	printf("driver %d runtime %d\n", drv, rt);
	return 0;
}
