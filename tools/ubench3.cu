// tools/ubench3.cu — third set of sm_100a micro-benchmarks (round 2), questions left after ubench2:
//   Q1  same-word shared atomics with 16 / 32 lanes on ONE word, RED and ATOMS (value returned)
//   Q2  issue rates of IMAD.HI, IMAD.WIDE, LOP3, SHF, ISETP+SEL, VIMNMX, IADD3, LEA, FFMA (3-register form)
//   Q3  do the ALU pipe and the FMA pipe run side by side?  IMAD + PRMT interleaved, FFMA-imm + IMAD, ...
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench3 tools/ubench3.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

enum Mode { RED_SAME = 0, ATOM_SAME, IMADHI, IMADWIDE, LOP3, SHF, ISETP_SEL, VIMNMX, IADD3, LEA, FFMA3, IMAD_PRMT, FFMAIMM_IMAD, FFMAIMM_PRMT,
            IMAD_PRMT_FFMAIMM, RED_RANDOM, ATOM_RANDOM };

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k3(int iters, int param, uint32_t *sink, long long *cycles)
{
	extern __shared__ __align__(16) uint32_t sm[];
	for (int i = threadIdx.x; i < 32768; i += blockDim.x)
		sm[i] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
	uint32_t x[8], addr[8];
	unsigned long long w[8];
#pragma unroll
	for (int j = 0; j < 8; j++) {
		x[j] = threadIdx.x * 2654435761u + j * 40503u + blockIdx.x + 12345u;
		w[j] = x[j];
		// param lanes share a word; words of one instruction lie in distinct banks
		addr[j] = base + ((((threadIdx.x >> 5) * 8 + j) & 1023u) << 7) + (lane / param) * 4;
		if (MODE == RED_RANDOM || MODE == ATOM_RANDOM) // 32 pseudo-random words of a 128 KB table (fixed per thread)
			addr[j] = base + ((x[j] >> 7) & 0x7FFFu) * 4;
	}
	const uint32_t a = 1664525u + 2 * lane, c = 1013904223u;
	const float fa = 1.0009765625f + lane * 1e-6f, fc = 0.5f;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int j = 0; j < 8; j++) {
			if (MODE == RED_SAME || MODE == RED_RANDOM)
				asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr[j]), "r"(1u));
			else if (MODE == ATOM_SAME || MODE == ATOM_RANDOM) {
				uint32_t old;
				asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr[j]), "r"(1u));
				x[j] ^= old;
			} else if (MODE == IMADHI)
				x[j] = __umulhi(x[j], a) + c;
			else if (MODE == IMADWIDE)
				w[j] = (unsigned long long)(uint32_t)w[j] * a + w[j];
			else if (MODE == LOP3)
				asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(a), "r"(c));
			else if (MODE == SHF)
				asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[j]) : "r"(a));
			else if (MODE == ISETP_SEL)
				x[j] = x[j] > a ? x[j] - 1u : c; // ISETP + SEL / IADD
			else if (MODE == VIMNMX)
				x[j] = min(x[j] + 1u, a);
			else if (MODE == IADD3)
				asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(a));
			else if (MODE == LEA)
				x[j] = (x[j] << 3) + a;
			else if (MODE == FFMA3)
				x[j] = __float_as_uint(fmaf(__uint_as_float(x[j]), fa, fc));
			else if (MODE == IMAD_PRMT) {
				x[j] = x[j] * a + c;
				x[j] = __byte_perm(x[j], c, 0x2103);
			} else if (MODE == FFMAIMM_IMAD) {
				x[j] = __float_as_uint(fmaf(__uint_as_float(x[j]), 1.0009765625f, 0.5f));
				x[j] = x[j] * a + c;
			} else if (MODE == FFMAIMM_PRMT) {
				x[j] = __float_as_uint(fmaf(__uint_as_float(x[j]), 1.0009765625f, 0.5f));
				x[j] = __byte_perm(x[j], c, 0x2103);
			} else if (MODE == IMAD_PRMT_FFMAIMM) {
				x[j] = x[j] * a + c;
				x[j] = __byte_perm(x[j], c, 0x2103);
				x[j] = __float_as_uint(fmaf(__uint_as_float(x[j]), 1.0009765625f, 0.5f));
			}
		}
	}
	long long t1 = clock64();
	uint32_t acc = 0;
#pragma unroll
	for (int j = 0; j < 8; j++)
		acc ^= x[j] ^ (uint32_t)w[j] ^ (uint32_t)(w[j] >> 32);
	if (threadIdx.x == 0)
		cycles[blockIdx.x] = t1 - t0;
	if (acc == 0xdeadbeef)
		sink[0] = acc;
}

template <int MODE>
static void run(const char *name, int warps, int param, int ops_per_step)
{
	int sms = 0;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	uint32_t *sink;
	long long *cyc;
	CK(cudaMalloc(&sink, 64));
	CK(cudaMemset(sink, 0, 64));
	CK(cudaMalloc(&cyc, sizeof(long long) * sms));
	CK(cudaFuncSetAttribute(k3<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
	const int iters = 2000;
	k3<MODE><<<sms, warps * 32, 131072>>>(10, param, sink, cyc);
	k3<MODE><<<sms, warps * 32, 131072>>>(iters, param, sink, cyc);
	CK(cudaDeviceSynchronize());
	std::vector<long long> h(sms);
	CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
	double avg = 0;
	for (auto v : h)
		avg += (double)v;
	avg /= sms;
	printf("{\"bench\": \"%s\", \"warps\": %d, \"param\": %d, \"cycles_per_warp_instr_per_SM\": %.3f}\n", name, warps, param,
	       avg / ((double)iters * 8 * warps * ops_per_step));
	CK(cudaFree(sink));
	CK(cudaFree(cyc));
}

int main()
{
	for (int warps : {8, 16}) {
		for (int k : {1, 2, 8, 16, 32}) {
			run<RED_SAME>("red_k_lanes_one_word", warps, k, 1);
			run<ATOM_SAME>("atom_returned_k_lanes_one_word", warps, k, 1);
		}
		run<RED_RANDOM>("red_32_random_words", warps, 1, 1);
		run<ATOM_RANDOM>("atom_returned_32_random_words", warps, 1, 1);
	}
	for (int warps : {4, 16}) {
		run<IMADHI>("imad_hi(+iadd)", warps, 1, 1);
		run<IMADWIDE>("imad_wide", warps, 1, 1);
		run<LOP3>("lop3", warps, 1, 1);
		run<SHF>("shf", warps, 1, 1);
		run<ISETP_SEL>("isetp+sel(2 instr)", warps, 1, 2);
		run<VIMNMX>("iadd+vimnmx(2 instr)", warps, 1, 2);
		run<IADD3>("iadd3", warps, 1, 1);
		run<LEA>("lea", warps, 1, 1);
		run<FFMA3>("ffma_3reg", warps, 1, 1);
		run<IMAD_PRMT>("imad+prmt(2 instr)", warps, 1, 2);
		run<FFMAIMM_IMAD>("ffma_imm+imad(2 instr)", warps, 1, 2);
		run<FFMAIMM_PRMT>("ffma_imm+prmt(2 instr)", warps, 1, 2);
		run<IMAD_PRMT_FFMAIMM>("imad+prmt+ffma_imm(3 instr)", warps, 1, 3);
	}
	return 0;
}
