// scope_kernels_experiments.cuh — strip kernels that were measured and NOT adopted, kept
// because they are selectable for A/B runs and covered by the tests: the row-group kernel (SCOPE_KERNEL=group,
// tests/test_gpu_parity.py::test_row_group_kernel_parity, tests/test_kernel_emulation.py) and the warp-specialised
// kernel (SCOPE_SPLIT=1).  Included by scope_kernels.cuh inside namespace scope; not a stand-alone header.
// ---------------------------------------------------------------------------
// Row-group consumer.  In the kernel above every consumer warp owns kTileRows / NW rows of EVERY
// tile, which ties the warp count to the tile height (16 warps x 4 rows = 64-row tiles) and makes
// the 16 warps walk the ring in lock step.  Here the unit of work is a GROUP of 4 rows (what one
// ldmatrix.x4 reads): the groups of a strip are numbered top to bottom, warp w takes groups
// w, w + NW, w + 2 NW, ... whatever tile they fall in; a warp with no group in a tile just passes
// it (see "barrier discipline" below).  The warp count is then free
// (24 warps at <= 80 registers fill the register file), warps drift apart by up to the ring
// depth instead of meeting at every tile, and the stage/phase of a group follow from the number
// of tiles this CTA has consumed so far, which every warp can compute on its own.
// ---------------------------------------------------------------------------
template <class L, int R_SRC, bool R_VS, bool SURFACE, int NWORK, bool K_BINS, bool K_VS>
__device__ __forceinline__ void tma_consume_groups(const StripParams &P, uint8_t *smem, uint32_t smem_base,
						   volatile uint32_t *chunk_q, uint32_t bar_full, uint32_t bar_empty,
						   int warp, int lane, int tid)
{
	constexpr int kStages = L::kStages;
	constexpr int N = kGroupRows;
	constexpr uint32_t GPT = L::kTileRows / kGroupRows; // groups per tile
	static_assert(L::kTileRows % kGroupRows == 0, "tile height must be a multiple of the group height");
	constexpr bool kNeedP = R_SRC == SRC_RGB || (!SURFACE && (R_VS || R_SRC == SRC_YUV));
	constexpr bool kNeedQ = SURFACE && (R_SRC == SRC_YUV || R_VS);
	uint32_t *vs = reinterpret_cast<uint32_t *>(smem + L::kVsOff);
	uint32_t *wave0 = reinterpret_cast<uint32_t *>(smem + L::kWave0Off);
	const uint32_t tiles = (P.height + L::kTileRows - 1) / L::kTileRows;
	const uint32_t groups = tiles * GPT;          // per strip, including the ones below the frame
	const uint32_t groups_inside = P.height / N;  // groups [0, groups_inside) have all 4 rows in the frame
	const uint32_t wave_lane_addr = smem_base + L::kWave0Off + lane * 4;
	uint32_t magic;
	magic = opaque_carrier_bias();
	const TileCtx tc{smem_base + L::kVsOff, wave_lane_addr - kCarrierBias * 128u,
			 wave_lane_addr - kCarrierBias * 128u + kWaveWords * 4, magic, P.bins_mask, lane};
	const Coef coef = P.coef;
	uint32_t zero;
	zero = opaque_zero();
	uint32_t tile_seq = 0;  // tiles this CTA consumed before the current strip (same in every warp)
	// Barrier discipline.  An mbarrier wait names a phase by its parity only, so a waiter must be
	// neither two phases ahead of the barrier nor two behind.  Every warp therefore visits EVERY
	// tile in order, also the tiles it has no group in: it waits for the tile ("full") and then
	// either reads its group and releases, or simply passes - one arrival per warp per tile on the
	// "empty" barrier either way (NWORK arrivals complete a phase).  Ahead: tile n - kStages (same
	// stage, previous phase) was waited for before tile n is asked about.  Behind: tile n + kStages
	// cannot be loaded before this warp has arrived for tile n, which it does after its wait.
#if !defined(SCOPE_EXPERIMENT) && !SCOPE_WIDE_FUSED // (the row-group kernel is not offered in those builds)
	static_assert(NWORK >= (int)GPT, "a warp must own at most one group per tile");
#endif
	uint32_t next_tile = 0; // first tile this warp has not finished (read + released, or passed)
	uint32_t waited = 0;    // tiles [0, waited) have been waited for; next_tile <= waited <= next_tile + 1
	uint32_t landed = 0;    // early answer of mbar_test for tile `waited`
	auto ensure_waited = [&](uint32_t m) {
		if (waited <= m) { // (then waited == m: tiles are waited for strictly in order)
			if (!landed)
				mbar_wait(bar_full + 8 * (m % kStages), (m / kStages) & 1u);
			landed = 0;
			waited = m + 1;
		}
	};
	// finish every tile before n without reading it, then wait for tile n
	auto advance_to = [&](uint32_t n) {
		while (next_tile < n) {
			ensure_waited(next_tile);
			if (lane == 0)
				mbar_arrive(bar_empty + 8 * (next_tile % kStages));
			next_tile++;
		}
		ensure_waited(n);
	};
	// ask early whether the next tile has landed, so the answer's latency hides behind arithmetic
	auto peek = [&]() {
		if (waited == next_tile)
			landed = mbar_test(bar_full + 8 * (waited % kStages), (waited / kStages) & 1u);
	};
	uint32_t qr = 0;
	uint32_t cur_frame = 0xFFFFFFFFu;

	for (;;) {
		// the chunk id becomes readable once the chunk's first tile (or the end marker) lands
		advance_to(tile_seq);
		const uint32_t first = chunk_q[2 * (qr % kQueue)], count = chunk_q[2 * (qr % kQueue) + 1];
		qr++;
		if (count == 0u)
			break;
		const uint32_t last = first + count;
		for (uint32_t item = first; item < last; item++) {
			const uint32_t frame = item / P.strips, strip = item - frame * P.strips;
			if (K_VS && frame != cur_frame && cur_frame != 0xFFFFFFFFu)
				flush_vscope<NWORK>(P, vs, cur_frame, tid);
			cur_frame = frame;
			const uint32_t x = strip * kStripPx + lane;
			const bool lane_ok = x < P.width;
			const bool strip_full = strip * kStripPx + kStripPx <= P.width;
			const uint32_t n_fast = strip_full ? groups_inside : 0u; // groups with no validity logic

			// wait for group g's tile, read this thread's 4 pixels; returns the barrier to release
			auto fetch = [&](uint32_t g, uint32_t(&p)[N], uint32_t(&q)[N]) -> uint32_t {
				const uint32_t n = tile_seq + g / GPT, stage = n % kStages;
				advance_to(n);
				next_tile = n + 1; // (the caller releases right after)
				const uint32_t rows = smem_base + L::kStageOff + stage * L::kStageBytes +
						      (g % GPT) * (N * kStripPx * 4);
				if (kNeedP)
					ldsm_rows<N>(rows, lane, p);
				if (kNeedQ)
					ldsm_rows<N>(rows + (L::kLoadRgb ? L::kTileBytes : 0), lane, q);
#pragma unroll
				for (int k = 0; k < N; k++) {
					if (!kNeedP)
						p[k] = 0u;
					if (!kNeedQ)
						q[k] = 0u;
				}
				return bar_empty + 8 * stage;
			};
			// hand the group back (data-dependent on the loaded pixels: see release_tile above)
			auto release = [&](uint32_t bar, const uint32_t(&p)[N], const uint32_t(&q)[N]) {
				uint32_t dep = 0;
#pragma unroll
				for (int k = 0; k < N; k++)
					dep |= p[k] | q[k];
				__syncwarp();
				if (lane == 0)
					mbar_arrive(bar + (dep & zero));
			};

			uint32_t g = warp;
			if (g < n_fast) {
				// software pipeline over this warp's interior groups (ping-pong A / B)
				uint32_t p[N], q[N], pend[N];
				Prep<N> A, B;
				auto issue = [&](const Prep<N> &o) {
					commit_issue<R_SRC, R_VS, SURFACE, N>(tc, o, pend);
					if (!SCOPE_DEFER)
						commit_resolve<R_SRC, R_VS, SURFACE, N>(tc, o, pend);
				};
				auto resolve = [&](const Prep<N> &o) {
					if (SCOPE_DEFER)
						commit_resolve<R_SRC, R_VS, SURFACE, N>(tc, o, pend);
				};
				uint32_t bar = fetch(g, p, q);
				release(bar, p, q);
				peek();
				prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, A);
				g += NWORK;
				for (; g + NWORK < n_fast; g += 2 * NWORK) {
					bar = fetch(g, p, q);
					issue(A);
					release(bar, p, q);
					peek();
					prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, B);
					resolve(A);
					bar = fetch(g + NWORK, p, q);
					issue(B);
					release(bar, p, q);
					peek();
					prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, A);
					resolve(B);
				}
				if (g < n_fast) {
					bar = fetch(g, p, q);
					issue(A);
					release(bar, p, q);
					peek();
					prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, B);
					resolve(A);
					commit_tile<R_SRC, R_VS, SURFACE, N>(tc, B);
					g += NWORK;
				} else {
					commit_tile<R_SRC, R_VS, SURFACE, N>(tc, A);
				}
			}
			for (; g < groups; g += NWORK) {
				// edge groups (last rows, last strip, rows below the frame): per-pixel validity
				uint32_t p[N], q[N];
				bool ok[N];
				const uint32_t bar = fetch(g, p, q);
				release(bar, p, q);
				const uint32_t y0 = g * N;
#pragma unroll
				for (int k = 0; k < N; k++)
					ok[k] = lane_ok && (y0 + k < P.height);
				if (y0 < P.height)
					process_tile<R_SRC, R_VS, SURFACE, false, N>(tc, coef, p, q, ok);
			}
			// pass the strip's remaining tiles (the ones after this warp's last group)
			tile_seq += tiles;
			while (next_tile < tile_seq) {
				ensure_waited(next_tile);
				if (lane == 0)
					mbar_arrive(bar_empty + 8 * (next_tile % kStages));
				next_tile++;
			}
			if (K_BINS)
				emit_strip<NWORK>(P, wave0, frame, x, lane_ok, warp, lane);
		}
	}
	if (K_VS && cur_frame != 0xFFFFFFFFu)
		flush_vscope<NWORK>(P, vs, cur_frame, tid);
}

template <int SRC, bool VSCOPE, bool SURFACE>
__global__ void __launch_bounds__(kGroupWarps * 32 + 32, 1)
	scope_strip_kernel_tmag(const __grid_constant__ StripParams P, const __grid_constant__ CUtensorMap map_rgb,
				const __grid_constant__ CUtensorMap map_yuv)
{
	using L = SmemLayout<SRC, VSCOPE, SURFACE, true>;
	constexpr int NW = kGroupWarps;
	SCOPE_DYNAMIC_SMEM(smem);
	volatile uint32_t *chunk_q = reinterpret_cast<volatile uint32_t *>(smem + L::kQueueOff);
	const uint32_t smem_base = smem_u32(smem);
	const uint32_t bar_full = smem_base + L::kBarOff;
	const uint32_t bar_empty = bar_full + kMaxStages * 8;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const bool is_producer = warp == NW;

	tma_setup<L, NW>(smem, bar_full, bar_empty, VSCOPE, SRC != SRC_NONE, !is_producer, tid);
	if (is_producer) {
		if (lane == 0)
			tma_produce<L>(P, &map_rgb, &map_yuv, smem_base, chunk_q, bar_full, bar_empty);
		return;
	}
	tma_consume_groups<L, SRC, VSCOPE, SURFACE, NW, SRC != SRC_NONE, VSCOPE>(P, smem, smem_base, chunk_q, bar_full,
										  bar_empty, warp, lane, tid);
}

// ---------------------------------------------------------------------------
// strip kernel, TMA loader, SPECIALISED warps for the headline combination (column bins on the
// RGB plane + vectorscope): kSplitVsWarps warps only do transform + vectorscope, kSplitBinWarps
// warps only do the waveform/histogram bins, both reading the same TMA tiles.  The vectorscope's
// 128 KB of bins allow one CTA per SM, so the only way to more resident warps is a wider CTA;
// the two roles have complementary instruction mixes (FP32 + 1 atomic vs 3 atomics per pixel).
// ---------------------------------------------------------------------------
template <bool SURFACE>
__global__ void __launch_bounds__((kSplitVsWarps + kSplitBinWarps) * 32 + 32, 1)
	scope_strip_kernel_split(const __grid_constant__ StripParams P, const __grid_constant__ CUtensorMap map_rgb,
				 const __grid_constant__ CUtensorMap map_yuv)
{
	using L = SmemLayout<SRC_RGB, true, SURFACE, true>;
	constexpr int NV = kSplitVsWarps, NB = kSplitBinWarps, NW = NV + NB;
	constexpr int RV = L::kTileRows / NV, RB = L::kTileRows / NB;
	SCOPE_DYNAMIC_SMEM(smem);
	volatile uint32_t *chunk_q = reinterpret_cast<volatile uint32_t *>(smem + L::kQueueOff);
	const uint32_t smem_base = smem_u32(smem);
	const uint32_t bar_full = smem_base + L::kBarOff;
	const uint32_t bar_empty = bar_full + kMaxStages * 8;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const bool is_producer = warp == NW;

	tma_setup<L, NW>(smem, bar_full, bar_empty, true, true, !is_producer, tid);
	if (is_producer) {
		if (lane == 0)
			tma_produce<L>(P, &map_rgb, &map_yuv, smem_base, chunk_q, bar_full, bar_empty);
		return;
	}
	if (warp < NV)
		tma_consume<L, SRC_NONE, true, SURFACE, RV, NW, true, true>(P, smem, smem_base, chunk_q, bar_full, bar_empty,
									    warp * RV, warp, lane, tid);
	else
		tma_consume<L, SRC_RGB, false, SURFACE, RB, NW, true, true>(P, smem, smem_base, chunk_q, bar_full, bar_empty,
									    (warp - NV) * RB, warp, lane, tid);
}

