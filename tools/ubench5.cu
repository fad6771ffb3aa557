// tools/ubench5.cu — round 2: what does the vectorscope atomic really cost?  Shared atomics on addresses that CHANGE
// every iteration (an LCG per thread, like pixel content), alone and mixed with the three conflict-free column-bin REDs
// of a pixel; 24 warps per SM like scope_fused_kernel_v3.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench5 tools/ubench5.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

enum Mode { VS_ONLY_ATOM = 0, VS_ONLY_RED, WAVE_ONLY, PIXEL_ATOM, PIXEL_RED, VS_DISTINCT_BANKS, VS_HALF_TABLE, VS_2WAY };

template <int MODE>
__global__ void __launch_bounds__(768, 1) k5(int iters, uint32_t *sink, long long *cycles)
{
	extern __shared__ __align__(16) uint32_t sm[];
	for (int i = threadIdx.x; i < 49152; i += blockDim.x)
		sm[i] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
	const uint32_t wave = base + 131072 + lane * 4; // 64 KB of column bins behind the 128 KB table
	uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
	uint32_t acc = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int j = 0; j < 4; j++) {
			x = x * 1664525u + 1013904223u;
			const uint32_t r = x >> 8;
			uint32_t vs;
			if (MODE == VS_DISTINCT_BANKS)
				vs = base + ((r & 0x3FFu) << 7) + lane * 4;       // random row, own bank: conflict-free
			else if (MODE == VS_HALF_TABLE)
				vs = base + (r & 0x3FFCu) * 4;                     // random words of 64 KB
			else if (MODE == VS_2WAY)
				vs = base + ((r & 0x3FFu) << 7) + (lane >> 1) * 4 + ((lane & 1) << 16); // 2 lanes per bank, distinct words
			else
				vs = base + (r & 0x7FFFu) * 4;                     // random words of the 128 KB table
			if (MODE == WAVE_ONLY || MODE == PIXEL_ATOM || MODE == PIXEL_RED) {
				asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(wave + ((r >> 3) & 0xFFu) * 128));
				asm volatile("red.shared.add.u32 [%0], 65536;" ::"r"(wave + ((r >> 11) & 0xFFu) * 128));
				asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(wave + 32768 + ((r >> 5) & 0xFFu) * 128));
			}
			if (MODE == VS_ONLY_ATOM || MODE == PIXEL_ATOM || MODE == VS_DISTINCT_BANKS || MODE == VS_HALF_TABLE || MODE == VS_2WAY) {
				uint32_t old;
				asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(vs), "r"(1u + (r & 1u) * 65535u));
				acc |= old;
			} else if (MODE == VS_ONLY_RED || MODE == PIXEL_RED) {
				asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(vs), "r"(1u + (r & 1u) * 65535u));
			}
		}
	}
	long long t1 = clock64();
	if (acc == 0x12345u)
		sink[0] = acc;
	if (threadIdx.x == 0)
		cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps)
{
	const int iters = 4000;
	uint32_t *sink;
	long long *cyc;
	CK(cudaMalloc(&sink, 64));
	CK(cudaMalloc(&cyc, 8 * 256));
	CK(cudaFuncSetAttribute(k5<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608));
	for (int rep = 0; rep < 2; rep++) {
		k5<MODE><<<148, warps * 32, 196608>>>(iters, sink, cyc);
		CK(cudaDeviceSynchronize());
	}
	long long h[148];
	CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
	double avg = 0;
	for (int i = 0; i < 148; i++)
		avg += (double)h[i];
	avg /= 148;
	printf("{\"bench\": \"%s\", \"warps\": %d, \"cycles_per_pixel_row_per_SM\": %.3f}\n", name, warps, avg / ((double)iters * 4 * warps));
	cudaFree(sink);
	cudaFree(cyc);
}

int main()
{
	for (int warps : {8, 24}) {
		run<VS_DISTINCT_BANKS>("1 ATOM, random rows, own bank (conflict-free)", warps);
		run<VS_2WAY>("1 ATOM, 2 lanes per bank, distinct words", warps);
		run<VS_ONLY_ATOM>("1 ATOM on random words of 128 KB", warps);
		run<VS_ONLY_RED>("1 RED on random words of 128 KB", warps);
		run<VS_HALF_TABLE>("1 ATOM on random words of 64 KB", warps);
		run<WAVE_ONLY>("3 conflict-free REDs (column bins)", warps);
		run<PIXEL_ATOM>("3 REDs + 1 ATOM random (one pixel row)", warps);
		run<PIXEL_RED>("3 REDs + 1 RED random (one pixel row)", warps);
	}
	return 0;
}
