// tools/ubench4.cu — round 2, question behind scope_fused_v3: does a shared-memory RED whose addend is NOT the
// constant 1 (SASS: ATOMS.ADD RZ instead of ATOMS.POPC.INC) still merge the lanes of a warp that hit one word?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench4 tools/ubench4.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

enum Mode { RED_ONE = 0, RED_REG, RED_64K, RED_MIXED_HALVES, RED_SPLIT_PRED, FFMA2_PAIR, FUNNEL };

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k4(int iters, int param, uint32_t addend, uint32_t *sink, long long *cycles)
{
	extern __shared__ __align__(16) uint32_t sm[];
	for (int i = threadIdx.x; i < 32768; i += blockDim.x)
		sm[i] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
	uint32_t addr[8], val[8];
	unsigned long long w[8];
#pragma unroll
	for (int j = 0; j < 8; j++) {
		// param lanes share a word; the words of one instruction lie in distinct banks
		addr[j] = base + ((((threadIdx.x >> 5) * 8 + j) & 1023u) << 7) + (lane / param) * 4;
		val[j] = addend;
		if (MODE == RED_MIXED_HALVES || MODE == RED_SPLIT_PRED)
			val[j] = (lane & 1) ? (addend << 16) : addend; // neighbours add to different halves of the shared word
		w[j] = threadIdx.x * 2654435761u + j;
	}
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int j = 0; j < 8; j++) {
			if (MODE == RED_ONE)
				asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr[j]));
			else if (MODE == RED_64K)
				asm volatile("red.shared.add.u32 [%0], 65536;" ::"r"(addr[j]));
			else if (MODE == RED_REG || MODE == RED_MIXED_HALVES)
				asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr[j]), "r"(val[j]));
			else if (MODE == RED_SPLIT_PRED) {
				asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 1; @!p red.shared.add.u32 [%0], 1; @p red.shared.add.u32 [%0], 65536; }"
					     ::"r"(addr[j]), "r"(val[j]));
			} else if (MODE == FFMA2_PAIR)
				asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(w[j]) : "l"(w[(j + 1) & 7]));
			else if (MODE == FUNNEL)
				asm volatile("shf.r.clamp.b32 %0, %0, %1, 6;" : "+r"(val[j]) : "r"(addr[j]));
		}
	}
	long long t1 = clock64();
	uint32_t acc = 0;
#pragma unroll
	for (int j = 0; j < 8; j++)
		acc ^= val[j] ^ (uint32_t)w[j];
	if (acc == 0x12345u)
		sink[0] = acc;
	if (threadIdx.x == 0)
		cycles[blockIdx.x] = t1 - t0;
	__syncthreads();
	if (threadIdx.x == 0 && blockIdx.x == 0)
		sink[1] = sm[0] + sm[32];
}

template <int MODE>
void run(const char *name, int warps, int param, uint32_t addend)
{
	const int iters = 2000;
	uint32_t *sink;
	long long *cyc;
	CK(cudaMalloc(&sink, 64));
	CK(cudaMalloc(&cyc, 8 * 256));
	CK(cudaFuncSetAttribute(k4<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
	k4<MODE><<<148, warps * 32, 131072>>>(iters, param, addend, sink, cyc);
	CK(cudaDeviceSynchronize());
	k4<MODE><<<148, warps * 32, 131072>>>(iters, param, addend, sink, cyc);
	CK(cudaDeviceSynchronize());
	long long h[148];
	CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
	double avg = 0;
	for (int i = 0; i < 148; i++)
		avg += (double)h[i];
	avg /= 148;
	printf("{\"bench\": \"%s\", \"warps\": %d, \"lanes_per_word\": %d, \"cycles_per_warp_instr_per_SM\": %.3f}\n", name, warps, param,
	       avg / ((double)iters * 8 * warps));
	cudaFree(sink);
	cudaFree(cyc);
}

int main()
{
	for (int warps : {8, 24}) {
		for (int k : {1, 2, 4, 8, 32}) {
			run<RED_ONE>("red_add_const_1 (POPC.INC)", warps, k, 1);
			run<RED_REG>("red_add_register (ATOMS.ADD RZ)", warps, k, 1);
			run<RED_64K>("red_add_const_65536", warps, k, 1);
			run<RED_MIXED_HALVES>("red_add_register_alternating_halves", warps, k, 1);
			run<RED_SPLIT_PRED>("split: @!p POPC.INC + @p ADD 65536 (2 instr)", warps, k, 1);
		}
		run<FFMA2_PAIR>("ffma2", warps, 1, 1);
		run<FUNNEL>("shf.r.clamp (funnel)", warps, 1, 1);
	}
	return 0;
}
