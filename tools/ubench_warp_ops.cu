// Developer tool: latency (dependent chain, one warp) and throughput (32 warps per SM, independent chains) of the warp-level
// and integer operations the mesh kernel's merge and flood phases are made of.  Prints cycles per operation.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench tools/ubench_warp_ops.cu && /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;
constexpr int N = 2048;

template <int OP> __device__ __forceinline__ u32 op(u32 x, u32 y, u32 *sm) {
	if (OP == 0) return (x ^ y) + 0x9e37u;                             // LOP3 + IADD
	if (OP == 1) return __funnelshift_l(x, y, 7);                      // SHF
	if (OP == 2) return __popc(x) + y;                                 // POPC
	if (OP == 3) return __ffs(x | 1u) + y;                             // BREV + FLO
	if (OP == 4) return __ballot_sync(0xffffffffu, x & 1u) + y;        // VOTE
	if (OP == 5) return __shfl_sync(0xffffffffu, x, y & 31) + 1u;      // SHFL.IDX
	if (OP == 6) return __reduce_or_sync(0xffffffffu, x) + y;          // REDUX.OR
	if (OP == 7) return __match_any_sync(0xffffffffu, x & 3u) + y;     // MATCH.ANY
	if (OP == 8) return sm[(x + threadIdx.x) & 1023] + y;              // LDS (dependent address)
	if (OP == 9) return atomicOr(&sm[(x + threadIdx.x) & 1023], y) + 1u; // ATOMS.OR
	if (OP == 10) return x * 0x00204081u + y;                          // IMAD
	if (OP == 11) return __clz(x) + y;                                 // FLO
	if (OP == 12) return __reduce_min_sync(0xffffffffu, x) + y;        // REDUX.MIN
	if (OP == 13) return (u32)__any_sync(0xffffffffu, x & 1u) + y;     // VOTE.ANY
	return x;
}

template <int OP, int ILP> __global__ void bench(u32 *out, long long *cyc, u32 seed) {
	__shared__ u32 sm[1024];
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 7u + seed;
	__syncthreads();
	u32 v[ILP];
	for (int i = 0; i < ILP; ++i) v[i] = seed + threadIdx.x * 17u + i;
	const long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < N; ++it) {
#pragma unroll
		for (int i = 0; i < ILP; ++i) v[i] = op<OP>(v[i], (u32)it, sm);
	}
	const long long t1 = clock64();
	u32 s = 0;
	for (int i = 0; i < ILP; ++i) s += v[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char *name, u32 *out, long long *cyc) {
	long long h[4];
	bench<OP, 1><<<1, 32>>>(out, cyc, 1);
	cudaMemcpy(&h[0], cyc, 8, cudaMemcpyDeviceToHost);
	bench<OP, 4><<<1, 32>>>(out, cyc, 1);
	cudaMemcpy(&h[1], cyc, 8, cudaMemcpyDeviceToHost);
	bench<OP, 1><<<1, 1024>>>(out, cyc, 1);
	cudaMemcpy(&h[2], cyc, 8, cudaMemcpyDeviceToHost);
	bench<OP, 4><<<1, 1024>>>(out, cyc, 1);
	cudaMemcpy(&h[3], cyc, 8, cudaMemcpyDeviceToHost);
	// latency: one warp, dependent chain.  throughput: 32 warps x 4 independent chains -> cycles per warp-instruction per SM sub-partition
	printf("%-22s latency %6.1f cyc | 1 warp ILP4 %6.1f cyc/op | 32 warps: %6.2f cyc/op/warp (ILP1) %6.2f (ILP4) -> %5.2f cyc per warp-op per SMSP\n", name,
	       (double)h[0] / N, (double)h[1] / N / 4, (double)h[2] / N, (double)h[3] / N / 4, (double)h[3] / N / 4 / 8);
}

int main() {
	u32 *out;
	long long *cyc;
	cudaMalloc(&out, 1 << 20);
	cudaMalloc(&cyc, 1024);
	printf("(each figure includes the +1 IADD/IMAD that chains the result)\n");
	run<0>("LOP3+IADD", out, cyc);
	run<1>("SHF", out, cyc);
	run<10>("IMAD", out, cyc);
	run<2>("POPC+IADD", out, cyc);
	run<3>("FFS (BREV+FLO)+IADD", out, cyc);
	run<11>("CLZ (FLO)+IADD", out, cyc);
	run<4>("VOTE.ballot+IADD", out, cyc);
	run<13>("VOTE.any+IADD", out, cyc);
	run<5>("SHFL.idx+IADD", out, cyc);
	run<6>("REDUX.or+IADD", out, cyc);
	run<12>("REDUX.min+IADD", out, cyc);
	run<7>("MATCH.any+IADD", out, cyc);
	run<8>("LDS+IADD", out, cyc);
	run<9>("ATOMS.or+IADD", out, cyc);
	cudaError_t e = cudaDeviceSynchronize();
	printf("%s\n", cudaGetErrorString(e));
	return 0;
}
