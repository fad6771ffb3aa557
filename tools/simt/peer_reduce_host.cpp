// tools/simt/peer_reduce_host.cpp — TEST INFRASTRUCTURE (no GPU): runs the per-thread body of the
// peer-memory reduce + saturate kernel (csrc/scope_peer_reduce.cuh) on the CPU, every (block, thread)
// of the grid in turn, with the same slice / grid arithmetic as scope_finalize_peers (csrc/scope_ffi.cu).
// tests/test_peer_reduce_host.py compares the outputs with numpy.
//   g++ -std=c++17 -O1 -ffp-contract=off -fPIC -shared -Iobs-color-monitor_b200/csrc tools/simt/peer_reduce_host.cpp
#include "scope_peer_reduce.cuh"

#include <string.h>

using namespace scope;

extern "C" int peer_reduce_host(const uint32_t *const *hist, const uint32_t *const *pairs, const uint32_t *const *vscope,
				uint32_t n_partials, uint8_t *const *wave, uint8_t *const *wave_display,
				uint8_t *const *vs_out, uint8_t *const *vs_display, uint32_t n_outs, uint32_t *hist_out,
				uint32_t full_width, uint32_t wave_planes, int do_vscope, uint32_t slice_index,
				uint32_t slice_count, uint32_t max_blocks, float wave_intensity, float vs_intensity)
{
	if (n_partials == 0 || n_partials > (uint32_t)kMaxPeers || n_outs == 0 || n_outs > (uint32_t)kMaxPeers)
		return 1;
	PeerReduceParams P;
	memset(&P, 0, sizeof P);
	P.n_partials = n_partials;
	P.n_outs = n_outs;
	P.n_px = 256ull * full_width;
	for (uint32_t i = 0; i < n_partials; i++) {
		P.hist[i] = hist ? hist[i] : nullptr;
		P.pairs[i] = pairs ? pairs[i] : nullptr;
		P.vscope[i] = vscope ? vscope[i] : nullptr;
	}
	for (uint32_t r = 0; r < n_outs; r++) {
		P.wave[r] = wave ? wave[r] : nullptr;
		P.wave_display[r] = wave_display ? wave_display[r] : nullptr;
		P.vs_out[r] = vs_out ? vs_out[r] : nullptr;
		P.vs_display[r] = vs_display ? vs_display[r] : nullptr;
	}
	P.wave_intensity = wave_intensity;
	P.vs_intensity = vs_intensity;
	P.wave_planes = wave_planes;
	const unsigned long long wave_quads = P.n_px / 4;
	if (wave_planes) {
		P.wave_q0 = (uint32_t)(wave_quads * slice_index / slice_count);
		P.wave_q1 = (uint32_t)(wave_quads * (slice_index + 1) / slice_count);
		const uint32_t n = P.wave_q1 - P.wave_q0;
		const uint32_t want = (n + 255u) / 256u;
		P.wave_blocks = n ? (want < max_blocks ? want : max_blocks) : 0u;
	}
	if (do_vscope) {
		P.vs_q0 = (uint32_t)(16384ull * slice_index / slice_count);
		P.vs_q1 = (uint32_t)(16384ull * (slice_index + 1) / slice_count);
		P.vs_blocks = (P.vs_q1 - P.vs_q0 + 255u) / 256u;
	}
	if (hist_out) {
		P.hist_out = hist_out;
		P.hist_blocks = 1;
	}
	const uint32_t grid = P.wave_blocks + P.vs_blocks + P.hist_blocks;
	for (uint32_t b = 0; b < grid; b++)
		for (uint32_t t = 0; t < 256; t++)
			peer_reduce_thread(P, b, t);
	return 0;
}
