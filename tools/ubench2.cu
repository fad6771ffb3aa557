// tools/ubench2.cu — second set of sm_100a micro-benchmarks: the questions DESIGN.md section 8.1 leaves
// open, to be answered by ONE short GPU call at the start of round 2.  Not part of the product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench2 tools/ubench2.cu
//   (under gpurun)  bash tools/run_ubench2.sh      -> gpurun_out/ubench2.txt
//
// Q1  lanes:   what does a conflict-free shared atomic cost, and does it depend on the active lanes?  (addresses
//              precomputed: no ALU work in the loop; warps whose upper lanes have exited)
// Q2  mix:     K arithmetic instructions per atomic, atomics issued as a BURST after the arithmetic or SPREAD
//              through it, with 4 / 8 / 12 / 16 warps: how well do LSU and issue overlap, and with how many warps?
// Q3  wide:    64-bit shared atomics (one lane-op updating two adjacent words) - same 16 lanes per clock?
// Q4  bank:    2- / 4- / 8-way BANK conflicts on distinct words versus k lanes on the SAME word
// Q5  pipes:   IMAD (register and immediate multiplier), PRMT, FFMA-immediate, IDP.4A / IDP.2A issue rates
//              (8 independent chains per thread)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void red_add(uint32_t addr, uint32_t v)
{
	asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ void red_add64(uint32_t addr, unsigned long long v)
{
	asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(addr), "l"(v));
}

enum Mode { LANES = 0, MIX_BURST = 1, MIX_SPREAD = 2, WIDE64 = 3, BANK_KWAY = 4, IDP4A = 5, IDP2A = 6, IMAD = 7, ALU_ONLY = 8,
            SAME_KWAY = 9, ATOM_RET = 10, PRMT = 11, FFMA_IMM = 12, IMAD_IMM = 13 };

// One CTA per SM.  `param` = active lanes (LANES), arithmetic instructions per atomic (MIX_*), k (BANK_KWAY).
template <int MODE, int K = 0>
__global__ void __launch_bounds__(1024, 1) k2(int iters, int param, uint32_t *sink, long long *cycles)
{
	extern __shared__ __align__(16) uint32_t sm[];
	for (int i = threadIdx.x; i < 32768; i += blockDim.x)
		sm[i] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
	if (MODE == LANES && lane >= param)
		return; // the warp goes on with `param` active lanes and no per-iteration branch
	uint32_t x[8], addr[8];
	uint32_t zero; // a 0 the compiler cannot see through: ties an atomic to a chain without arithmetic
	asm volatile("mov.u32 %0, 0;" : "=r"(zero));
#pragma unroll
	for (int j = 0; j < 8; j++) {
		x[j] = threadIdx.x * 2654435761u + j * 40503u + blockIdx.x;
		// fixed per-thread addresses, computed ONCE: the loops below spend no ALU instruction on them
		// (round 1's ubench spent ~4 per atomic and probably measured the ALU pipe, not the LSU).
		// default: lane's own bank, 8 different rows per warp
		addr[j] = base + ((((threadIdx.x >> 5) * 8 + j) & 1023u) << 7) + lane * 4;
		if (MODE == WIDE64)
			addr[j] = base + ((((threadIdx.x >> 5) * 8 + j) & 511u) << 8) + lane * 8;
		if (MODE == BANK_KWAY) // k lanes share a BANK but not a word
			addr[j] = base + (((((threadIdx.x >> 5) * 8 + j) * 8 + lane % param) & 1023u) << 7) + (lane / param) * 4;
		if (MODE == SAME_KWAY) // k lanes share a WORD
			addr[j] = base + ((((threadIdx.x >> 5) * 8 + j) & 1023u) << 7) + (lane / param) * 4;
	}
	const uint32_t a = 1664525u + 2 * lane, c = 1013904223u;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
		if (MODE == LANES || MODE == BANK_KWAY || MODE == SAME_KWAY) {
#pragma unroll
			for (int j = 0; j < 8; j++)
				red_add(addr[j], 1u);
		} else if (MODE == ATOM_RET) {
#pragma unroll
			for (int j = 0; j < 8; j++) {
				uint32_t old;
				asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr[j]), "r"(1u));
				x[j] ^= old;
			}
		} else if (MODE == WIDE64) {
#pragma unroll
			for (int j = 0; j < 8; j++)
				red_add64(addr[j], 0x0000000100000001ull);
		} else if (MODE == MIX_BURST || MODE == MIX_SPREAD || MODE == ALU_ONLY) {
			// 8 independent chains advance together; per atomic K arithmetic instructions, half multiply-adds
			// and half byte permutes (the two half-rate pipes of the real loop), plus one LOP3 that makes the
			// atomic's address depend on its chain.  BURST: all arithmetic first, then the 8 atomics.
			// SPREAD: chain j's atomic sits behind sub-step j mod (K/2) of the arithmetic.
			constexpr int S = K / 2;
#pragma unroll
			for (int st = 0; st < (S ? S : 1); st++) {
#pragma unroll
				for (int j = 0; j < 8; j++) {
					if (S) {
						x[j] = x[j] * a + c;
						x[j] = __byte_perm(x[j], c, 0x2103);
					}
					if (MODE == MIX_SPREAD && j % S == st)
						red_add(addr[j] + (x[j] & zero), 1u);
				}
			}
			if (MODE == MIX_BURST) {
#pragma unroll
				for (int j = 0; j < 8; j++)
					red_add(addr[j] + (x[j] & zero), 1u);
			}
		} else if (MODE == IDP4A) {
#pragma unroll
			for (int j = 0; j < 8; j++)
				x[j] = __dp4a(x[j], a, x[j]);
		} else if (MODE == IDP2A) {
#pragma unroll
			for (int j = 0; j < 8; j++)
				x[j] = __dp2a_lo(x[j], a, x[j]);
		} else if (MODE == IMAD) {
#pragma unroll
			for (int j = 0; j < 8; j++)
				x[j] = x[j] * a + c;
		} else if (MODE == IMAD_IMM) {
#pragma unroll
			for (int j = 0; j < 8; j++)
				x[j] = x[j] * 439216u + c; // multiplier as an immediate (SCOPE_IMMCOEF's form), addend in a register
		} else if (MODE == PRMT) {
#pragma unroll
			for (int j = 0; j < 8; j++)
				x[j] = __byte_perm(x[j], c, 0x2103);
		} else if (MODE == FFMA_IMM) {
#pragma unroll
			for (int j = 0; j < 8; j++)
				x[j] = __float_as_uint(fmaf(__uint_as_float(x[j]), 1.0009765625f, 0.5f));
		}
	}
	long long t1 = clock64();
	uint32_t acc = 0;
#pragma unroll
	for (int j = 0; j < 8; j++)
		acc ^= x[j];
	if (threadIdx.x == 0)
		cycles[blockIdx.x] = t1 - t0;
	if (acc == 0xdeadbeef)
		sink[0] = acc;
}

template <int MODE, int K = 0>
static void run(const char *name, int warps, int param)
{
	int sms = 0;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	uint32_t *sink;
	long long *cyc;
	CK(cudaMalloc(&sink, 64));
	CK(cudaMemset(sink, 0, 64));
	CK(cudaMalloc(&cyc, sizeof(long long) * sms));
	CK(cudaFuncSetAttribute(k2<MODE, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
	const int iters = 2000;
	k2<MODE, K><<<sms, warps * 32, 131072>>>(10, param, sink, cyc);
	k2<MODE, K><<<sms, warps * 32, 131072>>>(iters, param, sink, cyc);
	CK(cudaDeviceSynchronize());
	std::vector<long long> h(sms);
	CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
	double avg = 0;
	for (auto v : h)
		avg += (double)v;
	avg /= sms;
	// "op" = one chain step group: one atomic (LANES/WIDE/BANK/MIX) or one arithmetic instruction (IDP/IMAD)
	printf("{\"bench\": \"%s\", \"warps\": %d, \"param\": %d, \"cycles_per_op_per_SM\": %.3f}\n", name, warps, param,
	       avg / ((double)iters * 8 * warps));
	CK(cudaFree(sink));
	CK(cudaFree(cyc));
}

int main()
{
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	printf("{\"device\": \"%s\", \"sms\": %d}\n", prop.name, prop.multiProcessorCount);
	for (int warps : {4, 8, 16})
		for (int lanes : {1, 8, 16, 24, 32})
			run<LANES>("atoms_active_lanes", warps, lanes);
	for (int warps : {8, 16}) {
		run<ATOM_RET>("atoms_value_returned", warps, 0);
		run<WIDE64>("atoms_u64_own_bank_pair", warps, 0);
		for (int k : {2, 4, 8}) {
			run<BANK_KWAY>("atoms_kway_bank_distinct_words", warps, k);
			run<SAME_KWAY>("atoms_kway_same_word", warps, k);
		}
	}
	// the real loop: ~8 arithmetic instructions per atomic (27 core + overhead over 4 atomics)
	for (int warps : {4, 8, 12, 16}) {
		run<MIX_BURST, 0>("mix_burst", warps, 0);
		run<MIX_BURST, 4>("mix_burst", warps, 4);
		run<MIX_SPREAD, 4>("mix_spread", warps, 4);
		run<ALU_ONLY, 4>("mix_alu_only", warps, 4);
		run<MIX_BURST, 8>("mix_burst", warps, 8);
		run<MIX_SPREAD, 8>("mix_spread", warps, 8);
		run<ALU_ONLY, 8>("mix_alu_only", warps, 8);
		run<MIX_BURST, 12>("mix_burst", warps, 12);
		run<MIX_SPREAD, 12>("mix_spread", warps, 12);
		run<ALU_ONLY, 12>("mix_alu_only", warps, 12);
	}
	for (int warps : {4, 8, 16}) {
		run<IMAD>("imad", warps, 0);
		run<IMAD_IMM>("imad_immediate_multiplier", warps, 0);
		run<PRMT>("prmt", warps, 0);
		run<FFMA_IMM>("ffma_imm", warps, 0);
		run<IDP4A>("idp4a", warps, 0);
		run<IDP2A>("idp2a", warps, 0);
	}
	return 0;
}
