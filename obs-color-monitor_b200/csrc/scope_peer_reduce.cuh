// scope_peer_reduce.cuh — tile-sharded frames (one frame split into row / column bands over the
// GPUs of a node, BASELINE config 4): sum the ranks' partial bins over PEER MEMORY and apply the
// reference's saturation in ONE kernel, instead of an NCCL all-reduce followed by a clamp kernel.
//
// Every rank holds partial accumulators (struct scope_partial_device, include/scope_ffi.h) that all
// ranks can address (NVLink peer mappings: torch symmetric memory, cudaIpcOpenMemHandle, ...).
// A rank takes a SLICE of the bins, loads that slice from every peer's partials (16-byte peer
// loads), adds, saturates - min(sum, 255) is what the reference's inc_uint8 (waveform.c:201-205)
// and `if (*c < 255) ++*c` (vectorscope.c:233-234) give for the whole frame - and stores the u8
// result into every rank's output image (peer stores).  With slice = everything and one output it
// is the one-shot form (each rank reads all partials, writes only locally).  The histogram is
// 4 KB: every rank sums all of it for itself (hi_max needs the whole table anyway).
//
// NVLS form (scope_finalize_multicast): with the partials and images bound to a multicast object the
// switch does the sum (multimem.ld_reduce) and the distribution (multimem.st): one response and one
// store per bin instead of N.
//
// The per-thread body is plain C++ so that tests/test_peer_reduce_host.py can run exactly this
// code on the CPU (tools/simt/peer_reduce_host.cpp) against numpy; the kernel below is a thin
// wrapper.  Synchronisation between the ranks (partials complete before, outputs complete after)
// is the caller's: see scope_finalize_peers in include/scope_ffi.h.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SCOPE_PEER_FN __device__ __forceinline__
#else
#define SCOPE_PEER_FN static inline
#endif

namespace scope {

constexpr int kMaxPeers = 16;

struct alignas(16) Quad {
	uint32_t x, y, z, w;
};

struct PeerReduceParams {
	const uint32_t *hist[kMaxPeers];   // [1024] per rank
	const uint32_t *pairs[kMaxPeers];  // [2][n_px] per rank (u16 pairs, scope_partial_device.wave_pairs)
	const uint32_t *vscope[kMaxPeers]; // [65536] per rank
	uint8_t *wave[kMaxPeers];          // outputs: [n_px][4] per receiving rank (NULL entries are skipped)
	uint8_t *wave_display[kMaxPeers];
	uint8_t *vs_out[kMaxPeers];        // [65536]
	uint8_t *vs_display[kMaxPeers];
	uint32_t *hist_out;                // [1024], local only
	unsigned long long n_px;           // 256 * full_width (a multiple of 4)
	uint32_t n_partials, n_outs;
	uint32_t wave_planes;              // 0 = no waveform, 1 = plane 0 only (no R|V channel), 2 = both
	uint32_t wave_q0, wave_q1;         // this slice: quads (4 consecutive bins) [q0, q1) of the n_px / 4
	uint32_t vs_q0, vs_q1;             // and of the 16384 vectorscope quads
	uint32_t wave_blocks, vs_blocks, hist_blocks; // grid = the sum of the three
	uint32_t multicast;                // 1: entry [0] of hist / pairs / vscope is an NVLS multicast address and the
	                                   // switch adds the ranks' copies (multimem.ld_reduce); n_partials is ignored
	uint32_t multicast_out;            // 1: the image outputs (entry [0]) are multicast addresses: one multimem.st
	                                   // puts the slice into every rank's image
	float wave_intensity, vs_intensity; // display mapping when the *_display outputs are set
};

// display mapping, same arithmetic as intensity_u8 (scope_kernels.cuh) and oracle/scope_oracle.c
SCOPE_PEER_FN uint32_t peer_intensity_u8(uint32_t c, float k)
{
#ifdef __CUDA_ARCH__
	float r = __fmul_rn(__fdiv_rn((float)c, 255.0f), k);
	r = fminf(r, 1.0f);
	return (uint32_t)floorf(__fmaf_rn(r, 255.0f, 0.5f));
#else
	volatile float q = (float)c / 255.0f; // (volatile: no contraction / reassociation on the host either)
	volatile float r = q * k;
	float m = fminf(r, 1.0f);
	return (uint32_t)floorf(fmaf(m, 255.0f, 0.5f));
#endif
}

SCOPE_PEER_FN uint32_t peer_min255(uint32_t v)
{
	return v < 255u ? v : 255u;
}

// Sum of one quad (4 consecutive u32 bins) over the ranks.  Peer form: one 16-byte load per rank.  Multicast
// form (device only): two multimem.ld_reduce.add.u64 on the NVLS multicast address - the NVSwitch adds the
// ranks' copies and returns the sum, one response per rank instead of N.  Adding the bins as u64 pairs is
// exact because no u32 lane's total reaches 2^32 (histogram and vectorscope totals are bounded by the
// pixel count of the frame, the u16 halves of the waveform pairs by its height), so no carry crosses.
SCOPE_PEER_FN Quad peer_sum_quad(const PeerReduceParams &P, const uint32_t *const *src, unsigned long long word)
{
#ifdef __CUDA_ARCH__
	if (P.multicast) {
		const uint32_t *a = src[0] + word;
		unsigned long long lo, hi;
		asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(lo) : "l"(a) : "memory");
		asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(hi) : "l"(a + 2) : "memory");
		const Quad m = {(uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)};
		return m;
	}
#endif
	Quad s = {0, 0, 0, 0};
#pragma unroll 4
	for (uint32_t p = 0; p < P.n_partials; p++) {
		const Quad a = *reinterpret_cast<const Quad *>(src[p] + word);
		s.x += a.x;
		s.y += a.y;
		s.z += a.z;
		s.w += a.w;
	}
	return s;
}

SCOPE_PEER_FN void peer_store_quad(const PeerReduceParams &P, uint8_t *dst, const Quad &o)
{
#ifdef __CUDA_ARCH__
	if (P.multicast_out) {
		const unsigned long long lo = (unsigned long long)o.x | ((unsigned long long)o.y << 32);
		const unsigned long long hi = (unsigned long long)o.z | ((unsigned long long)o.w << 32);
		asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(lo) : "memory");
		asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 8), "l"(hi) : "memory");
		return;
	}
#endif
	*reinterpret_cast<Quad *>(dst) = o;
}

SCOPE_PEER_FN void peer_store_word(const PeerReduceParams &P, uint8_t *dst, uint32_t o)
{
#ifdef __CUDA_ARCH__
	if (P.multicast_out) {
		asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(o) : "memory");
		return;
	}
#endif
	*reinterpret_cast<uint32_t *>(dst) = o;
}

// one waveform word: summed pairs (B|U, G|Y in plane 0; R|V in plane 1) -> saturated BGRX
SCOPE_PEER_FN uint32_t peer_wave_word(uint32_t w0, uint32_t w1)
{
	return peer_min255(w0 & 0xFFFFu) | (peer_min255(w0 >> 16) << 8) | (peer_min255(w1 & 0xFFFFu) << 16);
}

SCOPE_PEER_FN uint32_t peer_display_word(uint32_t w, float k)
{
	return peer_intensity_u8(w & 0xFF, k) | (peer_intensity_u8((w >> 8) & 0xFF, k) << 8) |
	       (peer_intensity_u8((w >> 16) & 0xFF, k) << 16);
}

// What thread `tid` of block `block` (256 threads per block) does.
SCOPE_PEER_FN void peer_reduce_thread(const PeerReduceParams &P, uint32_t block, uint32_t tid)
{
	if (block < P.wave_blocks) {
		for (unsigned long long q = P.wave_q0 + (unsigned long long)block * 256 + tid; q < P.wave_q1;
		     q += (unsigned long long)P.wave_blocks * 256) {
			const Quad s0 = peer_sum_quad(P, P.pairs, q * 4);
			Quad s1 = {0, 0, 0, 0};
			if (P.wave_planes > 1)
				s1 = peer_sum_quad(P, P.pairs, P.n_px + q * 4);
			const Quad o = {peer_wave_word(s0.x, s1.x), peer_wave_word(s0.y, s1.y), peer_wave_word(s0.z, s1.z),
					peer_wave_word(s0.w, s1.w)};
			for (uint32_t r = 0; r < P.n_outs; r++) {
				if (P.wave[r])
					peer_store_quad(P, P.wave[r] + q * 16, o);
				if (P.wave_display[r]) {
					const Quad d = {peer_display_word(o.x, P.wave_intensity), peer_display_word(o.y, P.wave_intensity),
							peer_display_word(o.z, P.wave_intensity), peer_display_word(o.w, P.wave_intensity)};
					peer_store_quad(P, P.wave_display[r] + q * 16, d);
				}
			}
		}
		return;
	}
	block -= P.wave_blocks;
	if (block < P.vs_blocks) {
		for (uint32_t q = P.vs_q0 + block * 256 + tid; q < P.vs_q1; q += P.vs_blocks * 256) {
			const Quad s = peer_sum_quad(P, P.vscope, (unsigned long long)q * 4);
			const uint32_t c0 = peer_min255(s.x), c1 = peer_min255(s.y), c2 = peer_min255(s.z), c3 = peer_min255(s.w);
			const uint32_t o = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
			for (uint32_t r = 0; r < P.n_outs; r++) {
				if (P.vs_out[r])
					peer_store_word(P, P.vs_out[r] + (size_t)q * 4, o);
				if (P.vs_display[r])
					peer_store_word(P, P.vs_display[r] + (size_t)q * 4,
							peer_intensity_u8(c0, P.vs_intensity) | (peer_intensity_u8(c1, P.vs_intensity) << 8) |
								(peer_intensity_u8(c2, P.vs_intensity) << 16) |
								(peer_intensity_u8(c3, P.vs_intensity) << 24));
			}
		}
		return;
	}
	block -= P.vs_blocks;
	if (block < P.hist_blocks && tid < 256) {
		// 1024 counts = 256 quads, one per thread
		const Quad s = peer_sum_quad(P, P.hist, (unsigned long long)tid * 4);
		*reinterpret_cast<Quad *>(P.hist_out + tid * 4) = s;
	}
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(256) peer_reduce_finalize_kernel(const __grid_constant__ PeerReduceParams P)
{
	peer_reduce_thread(P, blockIdx.x, threadIdx.x);
}
#endif

} // namespace scope
