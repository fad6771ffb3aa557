// tools/ubench.cu — sm_100a micro-benchmarks that size the scope kernel's design
// choices (shared-memory atomics, LDS, MATCH, streaming-read shapes).  Not part of
// the product; results are summarised in profiles/ubench_r01.md.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench tools/ubench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

enum Mode { ATOM_OWN_BANK = 0, ATOM_RANDOM = 1, ATOM_SAME = 2, ATOM_KWAY = 3, LDS32 = 4, LDS128 = 5,
            MATCH_ANY = 6, ATOM_OWN_BANK_PRED = 7, LDS_STS_U8 = 8, ATOM_PAIR = 9, ATOM_RET = 10, REDUX = 11 };

// One CTA per SM, `warps` warps; every warp runs `iters` iterations of UNROLL ops.
template <int MODE>
__global__ void __launch_bounds__(1024, 1) smem_kernel(int iters, int kway, uint32_t *sink, long long *cycles)
{
	extern __shared__ uint32_t sm[];
	const int NW = 32768; // 128 KB of words
	for (int i = threadIdx.x; i < NW; i += blockDim.x)
		sm[i] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x * 97u + 12345u;
	uint32_t acc = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			uint32_t r = lcg(seed);
			if (MODE == ATOM_OWN_BANK) {
				atomicAdd(&sm[((r & 1023) << 5) | lane], 1u);
			} else if (MODE == ATOM_OWN_BANK_PRED) {
				if (r & 0x10000)
					atomicAdd(&sm[((r & 1023) << 5) | lane], 1u);
			} else if (MODE == ATOM_RANDOM) {
				atomicAdd(&sm[r & (NW - 1)], 1u);
			} else if (MODE == ATOM_SAME) {
				atomicAdd(&sm[(it * 8 + u) & (NW - 1)], 1u);
			} else if (MODE == ATOM_KWAY) {
				// kway lanes share one address; groups are on distinct banks
				int g = lane / kway;
				atomicAdd(&sm[(((r >> 10) & 1023) << 5) | g], 1u);
			} else if (MODE == ATOM_PAIR) {
				// two atomics whose addresses differ but hit the lane's own bank
				atomicAdd(&sm[((r & 1023) << 5) | lane], 0x10001u);
			} else if (MODE == ATOM_RET) {
				acc += atomicAdd(&sm[((r & 1023) << 5) | lane], 1u);
			} else if (MODE == LDS32) {
				acc += sm[((r & 1023) << 5) | lane];
			} else if (MODE == LDS128) {
				uint4 v = reinterpret_cast<uint4 *>(sm)[((r & 255) << 5) | lane];
				acc += v.x ^ v.y ^ v.z ^ v.w;
			} else if (MODE == MATCH_ANY) {
				acc += __match_any_sync(0xffffffffu, (r >> 4) & (kway - 1));
			} else if (MODE == REDUX) {
				acc += __reduce_add_sync(0xffffffffu, r & 0xff);
			} else if (MODE == LDS_STS_U8) {
				uint8_t *b = reinterpret_cast<uint8_t *>(sm);
				uint32_t a = ((r & 1023) << 7) | (lane << 2) | (u & 3);
				b[a] = b[a] + 1;
			}
		}
	}
	long long t1 = clock64();
	__syncthreads();
	if (threadIdx.x == 0)
		cycles[blockIdx.x] = t1 - t0;
	if (acc == 0xdeadbeef)
		sink[0] = acc;
	if (threadIdx.x < 32)
		atomicAdd(&sink[1], sm[threadIdx.x]);
}

template <int MODE>
static void run_smem(const char *name, int warps, int kway = 1)
{
	int dev_sms = 0;
	CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
	uint32_t *sink;
	long long *cyc;
	CK(cudaMalloc(&sink, 64));
	CK(cudaMemset(sink, 0, 64));
	CK(cudaMalloc(&cyc, sizeof(long long) * dev_sms));
	CK(cudaFuncSetAttribute(smem_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
	const int iters = 2000;
	smem_kernel<MODE><<<dev_sms, warps * 32, 131072>>>(10, kway, sink, cyc);
	smem_kernel<MODE><<<dev_sms, warps * 32, 131072>>>(iters, kway, sink, cyc);
	CK(cudaDeviceSynchronize());
	std::vector<long long> h(dev_sms);
	CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * dev_sms, cudaMemcpyDeviceToHost));
	double avg = 0;
	for (auto v : h)
		avg += (double)v;
	avg /= dev_sms;
	double warp_instr = (double)iters * 8 * warps;
	printf("{\"bench\": \"%s\", \"warps\": %d, \"kway\": %d, \"cycles_per_warp_instr_per_SM\": %.3f}\n", name, warps,
	       kway, avg / warp_instr);
	CK(cudaFree(sink));
	CK(cudaFree(cyc));
}

// ---------------- streaming-read shapes ----------------
// (a) linear LDG.128 over the whole buffer; (b) column strips: each CTA walks a
// strip of `strip_bytes` per row down all rows of a frame (pitch = width*4).
__global__ void __launch_bounds__(512) read_linear(const uint4 *__restrict__ p, size_t n, uint32_t *sink)
{
	uint32_t acc = 0;
	size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * 4) {
		uint4 v[4];
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (i + k * stride < n)
				asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
					     : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w)
					     : "l"(p + i + k * stride));
			else
				v[k] = make_uint4(0, 0, 0, 0);
#pragma unroll
		for (int k = 0; k < 4; k++)
			acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
	}
	if (acc == 0xdeadbeef)
		sink[0] = acc;
}

// strip read with plain loads: CTA item = (frame, strip); lanes cover strip_px pixels
// (4 B each) per row, warps cover different rows, UNR rows in flight per thread.
template <int STRIP_PX, int UNR>
__global__ void __launch_bounds__(512) read_strips(const uint32_t *__restrict__ p, int width, int height, int frames,
						    uint32_t *sink)
{
	const int strips = width / STRIP_PX;
	const int items = strips * frames;
	const int lanes_per_row = STRIP_PX;                  // one pixel per lane
	const int rows_per_pass = blockDim.x / lanes_per_row;
	const int lx = threadIdx.x % lanes_per_row, ly = threadIdx.x / lanes_per_row;
	uint32_t acc = 0;
	for (int item = blockIdx.x; item < items; item += gridDim.x) {
		const int f = item / strips, s = item % strips;
		const uint32_t *base = p + (size_t)f * width * height + (size_t)s * STRIP_PX + lx;
		for (int y = ly; y < height; y += rows_per_pass * UNR) {
			uint32_t v[UNR];
#pragma unroll
			for (int k = 0; k < UNR; k++) {
				int yy = y + k * rows_per_pass;
				v[k] = yy < height ? __ldg(base + (size_t)yy * width) : 0;
			}
#pragma unroll
			for (int k = 0; k < UNR; k++)
				acc += v[k];
		}
	}
	if (acc == 0xdeadbeef)
		sink[0] = acc;
}

template <typename F>
static float time_ms(F f, int reps)
{
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a));
	CK(cudaEventCreate(&b));
	f();
	CK(cudaDeviceSynchronize());
	CK(cudaEventRecord(a));
	for (int i = 0; i < reps; i++)
		f();
	CK(cudaEventRecord(b));
	CK(cudaEventSynchronize(b));
	float ms;
	CK(cudaEventElapsedTime(&ms, a, b));
	return ms / reps;
}

int main()
{
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	printf("{\"device\": \"%s\", \"sms\": %d, \"smem_optin\": %zu, \"l2\": %d}\n", prop.name, prop.multiProcessorCount,
	       prop.sharedMemPerBlockOptin, prop.l2CacheSize);

	for (int warps : {8, 16, 32}) {
		run_smem<ATOM_OWN_BANK>("atoms_own_bank", warps);
		run_smem<ATOM_OWN_BANK_PRED>("atoms_own_bank_pred50", warps);
		run_smem<ATOM_PAIR>("atoms_own_bank_add10001", warps);
		run_smem<ATOM_RET>("atoms_own_bank_ret", warps);
		run_smem<ATOM_RANDOM>("atoms_random_addr", warps);
		run_smem<ATOM_SAME>("atoms_same_addr_compiler", warps);
		for (int k : {2, 4, 8, 32})
			run_smem<ATOM_KWAY>("atoms_kway_same_addr", warps, k);
		run_smem<LDS32>("lds32_own_bank", warps);
		run_smem<LDS128>("lds128", warps);
		run_smem<LDS_STS_U8>("lds_sts_u8_rmw", warps);
		for (int k : {1, 4, 32})
			run_smem<MATCH_ANY>("match_any", warps, k);
		run_smem<REDUX>("redux_add", warps);
	}

	// streaming reads: 16 distinct 4K frames (531 MB) > 2x L2
	const int W = 3840, H = 2160, F = 16;
	const size_t bytes = (size_t)W * H * 4 * F;
	uint32_t *buf, *sink;
	CK(cudaMalloc(&buf, bytes));
	CK(cudaMemset(buf, 1, bytes));
	CK(cudaMalloc(&sink, 64));
	const int sms = prop.multiProcessorCount;
	for (int mult : {2, 4, 8}) {
		float ms = time_ms([&] { read_linear<<<sms * mult, 512>>>((const uint4 *)buf, bytes / 16, sink); }, 5);
		printf("{\"bench\": \"read_linear_ldg128\", \"ctas_per_sm\": %d, \"GBps\": %.1f}\n", mult, bytes / ms * 1e-6);
	}
	for (int mult : {1, 2, 4}) {
		float ms = time_ms([&] { read_strips<32, 8><<<sms * mult, 512>>>(buf, W, H, F, sink); }, 5);
		printf("{\"bench\": \"read_strips_32px_unr8\", \"ctas_per_sm\": %d, \"GBps\": %.1f}\n", mult, bytes / ms * 1e-6);
		ms = time_ms([&] { read_strips<32, 16><<<sms * mult, 512>>>(buf, W, H, F, sink); }, 5);
		printf("{\"bench\": \"read_strips_32px_unr16\", \"ctas_per_sm\": %d, \"GBps\": %.1f}\n", mult, bytes / ms * 1e-6);
		ms = time_ms([&] { read_strips<16, 16><<<sms * mult, 512>>>(buf, W, H, F, sink); }, 5);
		printf("{\"bench\": \"read_strips_16px_unr16\", \"ctas_per_sm\": %d, \"GBps\": %.1f}\n", mult, bytes / ms * 1e-6);
		ms = time_ms([&] { read_strips<64, 8><<<sms * mult, 512>>>(buf, W, H, F, sink); }, 5);
		printf("{\"bench\": \"read_strips_64px_unr8\", \"ctas_per_sm\": %d, \"GBps\": %.1f}\n", mult, bytes / ms * 1e-6);
	}
	return 0;
}
