// tools/ubench6.cu — round 2: what one warp-wide shared atomic on RANDOM words costs, with no address arithmetic in the
// loop (16 addresses per thread, hashed once, kept in registers).  32 random words fall into 32 banks with an expected
// maximum of ~3.5 per bank: 3.5 cycles per instruction if a conflict pass costs one cycle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench6 tools/ubench6.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t h)
{
	h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
	return h;
}
enum Mode { ATOM_RANDOM = 0, RED_RANDOM_REG, POPC_RANDOM, ATOM_OWN_BANK, PIXEL_MIX, ATOM_RANDOM_SAMEWORD_PAIRS };

template <int MODE>
__global__ void __launch_bounds__(768, 1) k6(int iters, uint32_t *sink, long long *cycles, float *avg_ways)
{
	extern __shared__ __align__(16) uint32_t sm[];
	for (int i = threadIdx.x; i < 49152; i += blockDim.x)
		sm[i] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
	uint32_t a[16], v[16], w[12];
	float ways = 0;
#pragma unroll
	for (int j = 0; j < 16; j++) {
		uint32_t r = mix(threadIdx.x * 16u + j + blockIdx.x * 12345u + 1u);
		if (MODE == ATOM_RANDOM_SAMEWORD_PAIRS)
			r = mix((threadIdx.x >> 1) * 16u + j + blockIdx.x * 12345u + 1u); // lanes 2k, 2k+1 on one word
		a[j] = base + (MODE == ATOM_OWN_BANK ? (((r & 0x3FFu) << 7) + lane * 4) : ((r & 0x7FFFu) * 4));
		v[j] = 1u + ((r >> 20) & 1u) * 65535u;
		// conflict ways of this instruction (max lanes per bank), measured with ballots
		uint32_t bank = (a[j] >> 2) & 31u, mx = 0;
		for (int b = 0; b < 32; b++)
			mx = max(mx, (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, bank == (uint32_t)b)));
		ways += (float)mx;
	}
#pragma unroll
	for (int j = 0; j < 12; j++) { // conflict-free column-bin addresses (lane = bank), random rows
		uint32_t r = mix(threadIdx.x * 12u + j + 777u);
		w[j] = base + 131072 + (j % 3 == 2 ? 32768 : 0) + ((r & 0xFFu) << 7) + lane * 4;
	}
	uint32_t acc = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int j = 0; j < 16; j++) {
			if (MODE == PIXEL_MIX) {
				asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(w[(3 * j) % 12]));
				asm volatile("red.shared.add.u32 [%0], 65536;" ::"r"(w[(3 * j + 1) % 12]));
				asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(w[(3 * j + 2) % 12]));
			}
			if (MODE == RED_RANDOM_REG)
				asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[j]), "r"(v[j]));
			else if (MODE == POPC_RANDOM)
				asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a[j]));
			else {
				uint32_t old;
				asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a[j]), "r"(v[j]));
				acc |= old;
			}
		}
	}
	long long t1 = clock64();
	if (acc == 0x12345u)
		sink[0] = acc;
	if (threadIdx.x == 0) {
		cycles[blockIdx.x] = t1 - t0;
		avg_ways[blockIdx.x] = ways / 16.0f;
	}
}

template <int MODE>
void run(const char *name, int warps)
{
	const int iters = 2000;
	uint32_t *sink; long long *cyc; float *ways;
	CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&cyc, 8 * 256)); CK(cudaMalloc(&ways, 4 * 256));
	CK(cudaFuncSetAttribute(k6<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608));
	for (int rep = 0; rep < 2; rep++) {
		k6<MODE><<<148, warps * 32, 196608>>>(iters, sink, cyc, ways);
		CK(cudaDeviceSynchronize());
	}
	long long h[148]; float hw[148];
	CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(hw, ways, sizeof hw, cudaMemcpyDeviceToHost));
	double avg = 0, w = 0;
	for (int i = 0; i < 148; i++) { avg += (double)h[i]; w += hw[i]; }
	avg /= 148; w /= 148;
	const double per = avg / ((double)iters * 16 * warps);
	printf("{\"bench\": \"%s\", \"warps\": %d, \"cycles_per_pixel_row_per_SM\": %.3f, \"conflict_ways_of_warp0\": %.2f}\n", name, warps, per, w);
	cudaFree(sink); cudaFree(cyc); cudaFree(ways);
}

int main()
{
	for (int warps : {8, 24}) {
		run<ATOM_OWN_BANK>("1 ATOM (returned), random rows, own bank", warps);
		run<ATOM_RANDOM>("1 ATOM (returned) on random words of 128 KB", warps);
		run<RED_RANDOM_REG>("1 RED (register addend) on random words", warps);
		run<POPC_RANDOM>("1 RED add 1 (POPC.INC) on random words", warps);
		run<ATOM_RANDOM_SAMEWORD_PAIRS>("1 ATOM, lane pairs on one word, pairs random", warps);
		run<PIXEL_MIX>("3 conflict-free REDs + 1 ATOM random (a pixel row)", warps);
	}
	return 0;
}
