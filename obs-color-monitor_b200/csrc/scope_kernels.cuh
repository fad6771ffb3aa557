// scope_kernels.cuh — sm_100a device code for the fused scope accumulation pass.
//
// One persistent kernel reads every BGRA pixel once and accumulates
//   * the waveform  (reference: wvs_draw_waveform,   src/waveform.c:220-257)
//   * the histogram (reference: his_draw_histogram,  src/histogram.c:357-395) — derived
//     from the waveform's per-column bins, so it costs no per-pixel work
//   * the vectorscope (reference: vss_draw_vectorscope, src/vectorscope.c:217-238)
// with the BT.601/709 transform of data/common.effect:23-43 evaluated in registers.
//
// Decomposition: a work item is a STRIP = 32 pixel columns (128 B per
// row) x all rows of one frame.  Lane l of every warp owns column l of the strip, so
//   - a waveform bin [level][column] lives in shared-memory bank l: conflict-free
//     atomics by construction, and no two lanes of a warp ever share an address;
//   - the CTA owns its 32 output columns exclusively and writes the final saturated
//     u8 waveform directly (no zero-fill pass, no global atomics for the waveform).
// Pixels arrive through TMA (cp.async.bulk.tensor) into a 4-stage shared-memory ring filled
// by a producer warp; consumers read one 32-bit pixel per lane per row.  A plain-load kernel
// covers planes TMA cannot describe.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace scope {

constexpr int kStripPx = 32;          // columns per strip == lanes per warp
#ifndef SCOPE_TILE_ROWS
#define SCOPE_TILE_ROWS 64
#endif
#ifndef SCOPE_TMA_WARPS
#define SCOPE_TMA_WARPS 16
#endif
constexpr int kTileRows = SCOPE_TILE_ROWS; // rows per TMA tile
constexpr int kTmaWarps = SCOPE_TMA_WARPS; // consumer warps of the TMA kernel (kTileRows / kTmaWarps rows each)
constexpr int kRingBytes = 32768;          // shared memory the bins leave for the tile ring
constexpr int kMaxStages = 8;              // upper bound of the ring depth (barrier storage)
constexpr int kTileBytes = kStripPx * 4 * kTileRows; // 8 KB per plane per stage
constexpr int kMaxChunkItems = 10;    // upper bound of strips per dynamically claimed chunk
constexpr int kQueue = 4;             // chunk-id mailbox entries (producer is < kQueue chunks ahead)
constexpr int kLdgWarps = 16;              // plain-load fallback kernel
constexpr int kLdgRows = 4;
constexpr int kVsWords = 32768;       // 65536 vectorscope bins, two u16 per word
constexpr int kWaveWords = 256 * 32;  // one plane: [level][lane]

enum : int { SRC_NONE = 0, SRC_RGB = 1, SRC_YUV = 2 };

struct Coef {
	float u0, u1, u2, y0, y1, y2, v0, v1, v2;
};

struct StripParams {
	const uint8_t *rgb;       // frame 0, device
	const uint8_t *yuv;       // frame 0, device (surface mode only)
	unsigned long long frame_stride; // bytes
	uint32_t linesize, width, height, n_frames;
	uint32_t strips;          // per frame
	uint32_t items;           // n_frames * strips
	uint32_t items_per_cta;   // plain-load kernel: static share of each CTA
	uint32_t chunk_items;     // TMA kernel: strips per dynamically claimed chunk
	uint32_t bins_mask;       // channels accumulated into the column bins: bit0 B|U, bit1 G|Y, bit2 R|V
	uint32_t hist_mask;       // channels the histogram output wants
	uint32_t wave_mask;       // channels the waveform output wants
	uint32_t x_offset;        // first output column (tile-sharded frames)
	uint32_t out_width;       // row length of the waveform output in pixels
	uint32_t partial;         // 1: add u16 pairs into wave_pairs instead of writing u8
	uint32_t *chunk_counter;  // global work counter of this launch (zeroed by the host)
	uint32_t *hist;           // [n][1024] u32, zeroed
	uint8_t *wave;            // [n][256][out_width][4]
	uint32_t *wave_pairs;     // partial: [2][256][out_width]: plane 0 = (B|U : lo16, G|Y : hi16), plane 1 = R|V
	uint32_t *vscope_acc;     // [n][65536] u32, zeroed
	unsigned long long hist_stride, wave_stride, vscope_stride; // elements between frames
	Coef coef;
};

// ---------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile("{\n"
		     ".reg .pred p;\n"
		     "WAIT_%=:\n"
		     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
		     "@p bra DONE_%=;\n"
		     "bra WAIT_%=;\n"
		     "DONE_%=:\n"
		     "}" ::"r"(bar),
		     "r"(parity)
		     : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y, int z)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
		     " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
		     "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z)
		     : "memory");
}
__device__ __forceinline__ uint32_t atom_shared_add(uint32_t addr, uint32_t val)
{
	uint32_t old;
	asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(val));
	return old;
}
__device__ __forceinline__ uint32_t ld_nc_u32(const void *p)
{
	uint32_t v;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

// ---- packed fp32x2 (sm_100a FMUL2 / FFMA2 / FADD2): two pixels per instruction ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
	f32x2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, uint32_t &lo, uint32_t &hi)
{
	asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
	f32x2 r;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
	f32x2 r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
	return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
	f32x2 r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f32x2 add2_rz(f32x2 a, f32x2 b)
{
	f32x2 r;
	asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f32x2 splat2(float c)
{
	return pack2(c, c);
}

// ---------------------------------------------------------------------------
// "carriers": a byte b travels as the bit pattern 0x4B0000bb, i.e. the float 2^23 + b.
//   * PRMT makes one from a pixel byte in a single instruction;
//   * the colour transform ends in such a pattern (add.rz with 2^23);
//   * (carrier << 7) + (base - 0x80000000) == base + 128*b   (0x4B000000 << 7 = 0x80000000
//     mod 2^32), so a waveform bin address is ONE multiply-add away from a carrier.
// ---------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ uint32_t carrier(uint32_t pixel, uint32_t magic)
{
	uint32_t r;
	if (K == 0)
		asm("prmt.b32 %0, %1, %2, 0x7650;" : "=r"(r) : "r"(pixel), "r"(magic));
	else if (K == 1)
		asm("prmt.b32 %0, %1, %2, 0x7651;" : "=r"(r) : "r"(pixel), "r"(magic));
	else
		asm("prmt.b32 %0, %1, %2, 0x7652;" : "=r"(r) : "r"(pixel), "r"(magic));
	return r;
}

// carriers of two pixels -> exact floats b (subtract 2^23), packed
__device__ __forceinline__ f32x2 carriers_to_f32x2(uint32_t ca, uint32_t cb)
{
	return add2(pack2(__uint_as_float(ca), __uint_as_float(cb)), splat2(-8388608.0f));
}

// x / 255 correctly rounded for x in {0..255}: fma(x, k0, rn(x*k1)), k0 = rn(1/255),
// k1 = rn(1/255 - k0).  Checked against IEEE division for all 256 inputs
// (tests/test_oracle.py::test_div255_constants) and, through the kernel, for all 2^24 colours.
__device__ __forceinline__ f32x2 div255(f32x2 x)
{
	const f32x2 k0 = splat2(__uint_as_float(0x3B808081u)); // 0x1.010102p-8
	const f32x2 k1 = splat2(__uint_as_float(0xAF7EFEFFu)); // -0x1.fdfdfep-33
	return fma2(x, k0, mul2(x, k1));
}

// one output channel for two pixels: p = c0*r; p = fma(c1,g,p); p = fma(c2,b,p); t = p + off;
// q = floor(fma(t, 255, 0.5)).  Returns the two CARRIERS of q (floats 2^23 + q).  The [0,1]
// clamp of the definition never acts (exhaustively verified, tests/test_oracle.py).
__device__ __forceinline__ void yuv_channel(f32x2 r, f32x2 g, f32x2 b, float c0, float c1, float c2, float off,
					    uint32_t &qa, uint32_t &qb)
{
	f32x2 p = mul2(splat2(c0), r);
	p = fma2(splat2(c1), g, p);
	p = fma2(splat2(c2), b, p);
	p = add2(p, splat2(off));
	p = fma2(p, splat2(255.0f), splat2(0.5f));
	unpack2(add2_rz(p, splat2(8388608.0f)), qa, qb);
}

// RGB carriers of two pixels -> U/(Y)/V carriers (data/common.effect:23-43 as pinned in
// oracle/scope_oracle.c).
template <bool NEED_Y>
__device__ __forceinline__ void rgb_to_yuv_pair(const uint32_t (&ca)[3], const uint32_t (&cb)[3], const Coef &c,
						 uint32_t (&ya)[3], uint32_t (&yb)[3])
{
	const f32x2 b = div255(carriers_to_f32x2(ca[0], cb[0]));
	const f32x2 g = div255(carriers_to_f32x2(ca[1], cb[1]));
	const f32x2 r = div255(carriers_to_f32x2(ca[2], cb[2]));
	yuv_channel(r, g, b, c.u0, c.u1, c.u2, 0.5f - 1.0f / 256.0f, ya[0], yb[0]);
	if (NEED_Y)
		yuv_channel(r, g, b, c.y0, c.y1, c.y2, 0.0f, ya[1], yb[1]);
	yuv_channel(r, g, b, c.v0, c.v1, c.v2, 0.5f, ya[2], yb[2]);
}

// ---------------------------------------------------------------------------
// shared-memory layout
// ---------------------------------------------------------------------------
template <int SRC, bool VSCOPE, bool SURFACE, bool USE_TMA>
struct SmemLayout {
	static constexpr bool kLoadRgb = !SURFACE || SRC == SRC_RGB;
	static constexpr bool kLoadYuv = SURFACE && (SRC == SRC_YUV || VSCOPE);
	static constexpr int kPlanes = (kLoadRgb ? 1 : 0) + (kLoadYuv ? 1 : 0);
	static constexpr int kVsOff = 0;
	static constexpr int kVsBytes = VSCOPE ? kVsWords * 4 : 0;
	static constexpr int kWave0Off = kVsOff + kVsBytes;
	static constexpr int kWaveBytes = SRC != SRC_NONE ? 2 * kWaveWords * 4 : 0;
	static constexpr int kStageOff = kWave0Off + kWaveBytes;
	static constexpr int kStageBytes = USE_TMA ? kPlanes * kTileBytes : 0;
	// two planes per stage (surface mode) leave room for a 2-deep ring only
	// as many stages as fit (two planes per stage in surface mode halve the depth)
	static constexpr int kStagesFit = USE_TMA ? kRingBytes / (kPlanes * kTileBytes) : 1;
	static constexpr int kStages = kStagesFit > kMaxStages ? kMaxStages : (kStagesFit < 2 ? 2 : kStagesFit);
#ifndef SCOPE_EXPERIMENT
	static_assert(!USE_TMA || kStagesFit >= 2, "tile too large for the ring");
#endif
	static constexpr int kBarOff = kStageOff + kStages * kStageBytes;
	static constexpr int kQueueOff = kBarOff + (USE_TMA ? 2 * kMaxStages * 8 : 0);
	static constexpr int kTotal = kQueueOff + (USE_TMA ? kQueue * 4 : 0) + 16;
};

// ---------------------------------------------------------------------------
// per-pixel accumulation primitives
// ---------------------------------------------------------------------------
__device__ __forceinline__ void red_shared(uint32_t addr, uint32_t val)
{
	asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(val));
}

// Waveform / histogram column bins: plane0[level][lane] = (count B|U : lo16, count G|Y : hi16),
// plane1[level][lane] = (count R|V : lo16).  A strip has <= 65535 rows, so no half overflows.
// wb0 / wb1 = lane's plane-0 / plane-1 base address minus 0x80000000 (see "carriers").
// `one` is 1 for a counted pixel and 0 for a skipped one (alpha == 0: histogram.c:385-387,
// waveform.c:246-248; or outside the frame): adding 0 needs no predication and no branch.
template <bool DO_B, bool DO_G, bool DO_R>
__device__ __forceinline__ void bins_add(uint32_t cb, uint32_t cg, uint32_t cr, uint32_t wb0, uint32_t wb1,
					 uint32_t one)
{
	if (DO_B)
		red_shared(cb * 128u + wb0, one);
	if (DO_G)
		red_shared(cg * 128u + wb0, one << 16);
	if (DO_R)
		red_shared(cr * 128u + wb1, one);
}

// Vectorscope bins in shared memory are indexed by idx = U | V << 8 (NOT yet flipped to the
// reference's row = 255 - V; the flush does that) and are u16 halves, two per 32-bit word:
// word = idx & 0x7FFF, half = V >> 7.  vs_add returns the bit of the OLD word that says "this
// half already held >= 0x8000"; the caller ORs those over its pixels and, only if any is set,
// calls vs_undo.  An add that found its bin at >= 0x8000 is taken back, so a half can never
// wrap 16 bits, and a bin that ever reached 0x8000 keeps a value far above 255: it saturates
// to 255 at the end exactly like the reference's `if (*c < 255) ++*c`.
struct VsAdd {
	uint32_t addr, add, sat;
};
__device__ __forceinline__ VsAdd vs_add(uint32_t vs_base, uint32_t idx, uint32_t k)
{
	VsAdd r;
	const uint32_t h = idx >> 15;            // 0: lower half (V < 128), 1: upper half
	r.addr = (idx & 0x7FFFu) * 4u + vs_base;
	r.add = h * (k * 0xFFFFu) + k;           // k << 16 for the upper half, k for the lower
	const uint32_t old = atom_shared_add(r.addr, r.add);
	r.sat = old & (h * 0x7FFF8000u + 0x8000u);
	return r;
}
__device__ __forceinline__ void vs_undo(const VsAdd &a)
{
	if (a.sat)
		red_shared(a.addr, 0u - a.add);
}

// ---------------------------------------------------------------------------
// one thread's share of a tile: N vertically adjacent pixels of its column.
// FAST = the tile lies completely inside the frame: no per-pixel validity logic at all.
// ---------------------------------------------------------------------------
struct TileCtx {
	uint32_t vs_base, wb0, wb1, magic, bins_mask;
	int lane;
};

template <int SRC, bool VSCOPE, bool SURFACE, bool FAST, int N>
__device__ __forceinline__ void process_tile(const TileCtx &c, const Coef &coef, const uint32_t (&p)[N],
					     const uint32_t (&q)[N], const bool (&ok)[N])
{
	static_assert(N % 2 == 0, "pixels are transformed in packed pairs");
	constexpr bool kTransform = !SURFACE && (VSCOPE || SRC == SRC_YUV);
	uint32_t crgb[N][3]; // carriers of B, G, R
	uint32_t cyuv[N][3]; // carriers of U, Y, V
	if (SRC == SRC_RGB || kTransform) {
#pragma unroll
		for (int k = 0; k < N; k++) {
			crgb[k][0] = carrier<0>(p[k], c.magic);
			crgb[k][1] = carrier<1>(p[k], c.magic);
			crgb[k][2] = carrier<2>(p[k], c.magic);
		}
	}
	if (kTransform) {
#pragma unroll
		for (int k = 0; k < N; k += 2)
			rgb_to_yuv_pair<SRC == SRC_YUV>(crgb[k], crgb[k + 1], coef, cyuv[k], cyuv[k + 1]);
	} else if (SURFACE && SRC == SRC_YUV) {
#pragma unroll
		for (int k = 0; k < N; k++) {
			cyuv[k][0] = carrier<0>(q[k], c.magic);
			cyuv[k][1] = carrier<1>(q[k], c.magic);
			cyuv[k][2] = carrier<2>(q[k], c.magic);
		}
	}

	// ---- waveform / histogram column bins ----
	if (SRC != SRC_NONE) {
		const uint32_t(*cs)[3] = SRC == SRC_RGB ? crgb : cyuv;
		// the word whose alpha byte decides whether a pixel counts (the fused YUV plane has
		// alpha 255 everywhere, common.effect:30,41); pixels outside the frame arrive as 0
		bool all_counted;
		if (SRC == SRC_RGB || SURFACE) {
			const uint32_t *a = SRC == SRC_RGB ? p : q;
			uint32_t m = a[0];
#pragma unroll
			for (int k = 1; k < N; k++)
				m = min(m, a[k]);
			all_counted = m > 0x00FFFFFFu;
		} else {
			all_counted = FAST || (ok[0] && ok[N - 1]);
		}
		if ((FAST || c.bins_mask == 7u) && __all_sync(0xFFFFFFFFu, all_counted)) {
#pragma unroll
			for (int k = 0; k < N; k++)
				bins_add<true, true, true>(cs[k][0], cs[k][1], cs[k][2], c.wb0, c.wb1, 1u);
		} else {
#pragma unroll
			for (int k = 0; k < N; k++) {
				uint32_t one;
				if (SRC == SRC_RGB)
					one = p[k] > 0x00FFFFFFu ? 1u : 0u;
				else if (SURFACE)
					one = q[k] > 0x00FFFFFFu ? 1u : 0u;
				else
					one = ok[k] ? 1u : 0u;
				if (FAST || (c.bins_mask & 1u))
					bins_add<true, false, false>(cs[k][0], cs[k][1], cs[k][2], c.wb0, c.wb1, one);
				if (FAST || (c.bins_mask & 2u))
					bins_add<false, true, false>(cs[k][0], cs[k][1], cs[k][2], c.wb0, c.wb1, one);
				if (FAST || (c.bins_mask & 4u))
					bins_add<false, false, true>(cs[k][0], cs[k][1], cs[k][2], c.wb0, c.wb1, one);
			}
		}
	}

	// ---- vectorscope ----
	if (VSCOPE) {
		uint32_t idx[N]; // U | V << 8
#pragma unroll
		for (int k = 0; k < N; k++) {
			if (SURFACE)
				idx[k] = __byte_perm(q[k], 0u, 0x4420);
			else
				idx[k] = __byte_perm(cyuv[k][0], cyuv[k][2], 0x1140);
		}
		const bool full = FAST || (ok[0] && ok[N - 1]); // this lane's N pixels all valid
		bool same = full;
#pragma unroll
		for (int k = 1; k < N; k++)
			same = same && (idx[k] == idx[0]);
		// (the shuffle must be executed by every lane: no short-circuit around it)
		const uint32_t idx_lane0 = __shfl_sync(0xFFFFFFFFu, idx[0], 0);
		same = same && (idx[0] == idx_lane0);
		if (__all_sync(0xFFFFFFFFu, same)) {
			// flat block: all N x 32 pixels hit one bin -> one atomic
			if (c.lane == 0)
				vs_undo(vs_add(c.vs_base, idx[0], 32u * N));
		} else if (FAST) {
			// N adds in flight, one combined overflow check
			VsAdd a[N];
			uint32_t any = 0;
#pragma unroll
			for (int k = 0; k < N; k++) {
				a[k] = vs_add(c.vs_base, idx[k], 1u);
				any |= a[k].sat;
			}
			if (any) {
#pragma unroll
				for (int k = 0; k < N; k++)
					vs_undo(a[k]);
			}
		} else {
#pragma unroll
			for (int k = 0; k < N; k++)
				if (ok[k])
					vs_undo(vs_add(c.vs_base, idx[k], 1u));
		}
	}
}

// ---------------------------------------------------------------------------
// pieces shared by the two strip kernels (NW = accumulating warps per CTA)
// ---------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void workers_bar()
{
	asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
}

template <int NW>
__device__ __forceinline__ void zero_bins(uint32_t *vs, uint32_t *wave0, bool vscope, bool bins, int tid)
{
	if (vscope)
		for (int i = tid; i < kVsWords / 4; i += NW * 32)
			reinterpret_cast<uint4 *>(vs)[i] = make_uint4(0, 0, 0, 0);
	if (bins)
		for (int i = tid; i < 2 * kWaveWords / 4; i += NW * 32)
			reinterpret_cast<uint4 *>(wave0)[i] = make_uint4(0, 0, 0, 0);
}

// every worker thread: move its slice of the u16 pairs to the frame's u32 accumulators,
// flipping V into the reference's row order (row = 255 - V, vectorscope.c:232)
template <int NW>
__device__ __forceinline__ void flush_vscope(const StripParams &P, uint32_t *vs, uint32_t frame, int tid)
{
	workers_bar<NW>();
	uint32_t *acc = P.vscope_acc + (size_t)frame * P.vscope_stride;
	for (int i = tid; i < kVsWords / 4; i += NW * 32) {
		uint4 w = reinterpret_cast<uint4 *>(vs)[i];
		if ((w.x | w.y | w.z | w.w) != 0u) {
			const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
			const uint32_t word = i * 4; // = U | (V & 0x7F) << 8
			const uint32_t u = word & 0xFFu, v7 = word >> 8;
			uint32_t *lo = acc + (255u - v7) * 256u + u; // V = v7
			uint32_t *hi = acc + (127u - v7) * 256u + u; // V = v7 | 0x80
#pragma unroll
			for (int j = 0; j < 4; j++) {
				if (ww[j] & 0xFFFFu)
					atomicAdd(lo + j, ww[j] & 0xFFFFu);
				if (ww[j] >> 16)
					atomicAdd(hi + j, ww[j] >> 16);
			}
			reinterpret_cast<uint4 *>(vs)[i] = make_uint4(0, 0, 0, 0);
		}
	}
	workers_bar<NW>();
}

// end of a strip: write this strip's 32 waveform columns (final, saturated) and add the
// strip's share of the histogram (= column bins summed over the 32 columns); re-zero the bins
template <int NW>
__device__ __forceinline__ void emit_strip(const StripParams &P, uint32_t *wave0, uint32_t frame, uint32_t x,
					   bool lane_ok, int warp, int lane)
{
	workers_bar<NW>();
	uint32_t *hist = P.hist + (size_t)frame * P.hist_stride;
	const uint32_t xo = P.x_offset + x;
	for (int v = warp; v < 256; v += NW) {
		const uint32_t w0 = wave0[v * 32 + lane];
		const uint32_t w1 = wave0[kWaveWords + v * 32 + lane];
		wave0[v * 32 + lane] = 0;
		wave0[kWaveWords + v * 32 + lane] = 0;
		const uint32_t cb = w0 & 0xFFFFu, cg = w0 >> 16, cr = w1 & 0xFFFFu;
		if (P.hist_mask) {
			const uint32_t sb = __reduce_add_sync(0xFFFFFFFFu, cb);
			const uint32_t sg = __reduce_add_sync(0xFFFFFFFFu, cg);
			const uint32_t sr = __reduce_add_sync(0xFFFFFFFFu, cr);
			if (lane == 0) {
				if ((P.hist_mask & 4u) && sr)
					atomicAdd(hist + v * 4 + 0, sr);
				if ((P.hist_mask & 2u) && sg)
					atomicAdd(hist + v * 4 + 1, sg);
				if ((P.hist_mask & 1u) && sb)
					atomicAdd(hist + v * 4 + 2, sb);
			}
		}
		if (P.wave_mask && lane_ok) {
			const uint32_t mb = (P.wave_mask & 1u) ? cb : 0u;
			const uint32_t mg = (P.wave_mask & 2u) ? cg : 0u;
			const uint32_t mr = (P.wave_mask & 4u) ? cr : 0u;
			const size_t o = (size_t)(255 - v) * P.out_width + xo;
			if (P.partial) {
				if (mb | mg)
					atomicAdd(P.wave_pairs + o, mb | (mg << 16));
				if (mr)
					atomicAdd(P.wave_pairs + (size_t)256 * P.out_width + o, mr);
			} else {
				uint32_t *dst = reinterpret_cast<uint32_t *>(P.wave + (size_t)frame * P.wave_stride);
				dst[o] = min(mb, 255u) | (min(mg, 255u) << 8) | (min(mr, 255u) << 16);
			}
		}
	}
	workers_bar<NW>();
}

// ---------------------------------------------------------------------------
// Two-phase form of the interior-tile body (tile completely inside the frame, all three
// channels on): prepare_tile = carriers, colour transform, indices and the two warp votes;
// commit_tile = the shared-memory atomics.  The TMA kernel runs commit(tile t) and
// prepare(tile t+1) back to back so that a warp overlaps the LSU work of one tile with the
// FP32 work of the next instead of alternating between the two pipes.
// ---------------------------------------------------------------------------
template <int N>
struct Prep {
	uint32_t cs[N][3]; // carriers of the three bytes the column bins look at
	uint32_t a[N];     // the word whose alpha decides whether the pixel counts
	uint32_t idx[N];   // vectorscope bin, U | V << 8
	bool all_counted;  // warp-uniform: every pixel of the warp's N x 32 block counts
	bool flat;         // warp-uniform: the whole block hits one vectorscope bin
};

template <int SRC, bool VSCOPE, bool SURFACE, int N>
__device__ __forceinline__ void prepare_tile(const TileCtx &c, const Coef &coef, const uint32_t (&p)[N],
					     const uint32_t (&q)[N], Prep<N> &o)
{
	constexpr bool kTransform = !SURFACE && (VSCOPE || SRC == SRC_YUV);
	uint32_t crgb[N][3], cyuv[N][3];
	if (SRC == SRC_RGB || kTransform) {
#pragma unroll
		for (int k = 0; k < N; k++) {
			crgb[k][0] = carrier<0>(p[k], c.magic);
			crgb[k][1] = carrier<1>(p[k], c.magic);
			crgb[k][2] = carrier<2>(p[k], c.magic);
		}
	}
	if (kTransform) {
#pragma unroll
		for (int k = 0; k < N; k += 2)
			rgb_to_yuv_pair<SRC == SRC_YUV>(crgb[k], crgb[k + 1], coef, cyuv[k], cyuv[k + 1]);
	} else if (SURFACE && SRC == SRC_YUV) {
#pragma unroll
		for (int k = 0; k < N; k++) {
			cyuv[k][0] = carrier<0>(q[k], c.magic);
			cyuv[k][1] = carrier<1>(q[k], c.magic);
			cyuv[k][2] = carrier<2>(q[k], c.magic);
		}
	}
	o.all_counted = true;
	if (SRC != SRC_NONE) {
#pragma unroll
		for (int k = 0; k < N; k++) {
			o.cs[k][0] = SRC == SRC_RGB ? crgb[k][0] : cyuv[k][0];
			o.cs[k][1] = SRC == SRC_RGB ? crgb[k][1] : cyuv[k][1];
			o.cs[k][2] = SRC == SRC_RGB ? crgb[k][2] : cyuv[k][2];
		}
		if (SRC == SRC_RGB || SURFACE) {
			uint32_t m = 0xFFFFFFFFu;
#pragma unroll
			for (int k = 0; k < N; k++) {
				o.a[k] = SRC == SRC_RGB ? p[k] : q[k];
				m = min(m, o.a[k]);
			}
			o.all_counted = __all_sync(0xFFFFFFFFu, m > 0x00FFFFFFu);
		}
	}
	o.flat = false;
	if (VSCOPE) {
#pragma unroll
		for (int k = 0; k < N; k++)
			o.idx[k] = SURFACE ? __byte_perm(q[k], 0u, 0x4420) : __byte_perm(cyuv[k][0], cyuv[k][2], 0x1140);
		bool same = true;
#pragma unroll
		for (int k = 1; k < N; k++)
			same = same && (o.idx[k] == o.idx[0]);
		const uint32_t idx_lane0 = __shfl_sync(0xFFFFFFFFu, o.idx[0], 0);
		o.flat = __all_sync(0xFFFFFFFFu, same && (o.idx[0] == idx_lane0));
	}
}

template <int SRC, bool VSCOPE, bool SURFACE, int N>
__device__ __forceinline__ void commit_tile(const TileCtx &c, const Prep<N> &o)
{
	if (SRC != SRC_NONE) {
		if (o.all_counted && c.bins_mask == 7u) {
#pragma unroll
			for (int k = 0; k < N; k++)
				bins_add<true, true, true>(o.cs[k][0], o.cs[k][1], o.cs[k][2], c.wb0, c.wb1, 1u);
		} else {
			// some pixels transparent, or only some channels wanted (uniform branches)
#pragma unroll
			for (int k = 0; k < N; k++) {
				const uint32_t one = (o.all_counted || o.a[k] > 0x00FFFFFFu) ? 1u : 0u;
				if (c.bins_mask & 1u)
					bins_add<true, false, false>(o.cs[k][0], o.cs[k][1], o.cs[k][2], c.wb0, c.wb1, one);
				if (c.bins_mask & 2u)
					bins_add<false, true, false>(o.cs[k][0], o.cs[k][1], o.cs[k][2], c.wb0, c.wb1, one);
				if (c.bins_mask & 4u)
					bins_add<false, false, true>(o.cs[k][0], o.cs[k][1], o.cs[k][2], c.wb0, c.wb1, one);
			}
		}
	}
	if (VSCOPE) {
		if (o.flat) {
			if (c.lane == 0)
				vs_undo(vs_add(c.vs_base, o.idx[0], 32u * N));
		} else {
			// N adds in flight; their old words are OR-ed and tested once for "some half
			// already >= 0x8000" (either half: a false alarm only costs the exact re-check)
			uint32_t old[N], any = 0;
#pragma unroll
			for (int k = 0; k < N; k++) {
				const uint32_t h = o.idx[k] >> 15;
				old[k] = atom_shared_add((o.idx[k] & 0x7FFFu) * 4u + c.vs_base, h * 0xFFFFu + 1u);
				any |= old[k];
			}
			if (any & 0x80008000u) {
#pragma unroll
				for (int k = 0; k < N; k++) {
					const uint32_t h = o.idx[k] >> 15;
					if (old[k] & (h * 0x7FFF8000u + 0x8000u))
						red_shared((o.idx[k] & 0x7FFFu) * 4u + c.vs_base, 0u - (h * 0xFFFFu + 1u));
				}
			}
		}
	}
}

// ---------------------------------------------------------------------------
// strip kernel, TMA loader.  A producer warp (one elected lane) walks CHUNKS of P.chunk_items
// consecutive strips, claimed from a global counter so that fast and slow frame content
// balances across CTAs, and fills a kStages-deep shared-memory ring of 64-row x 128-byte
// tiles with cp.async.bulk.tensor.  16 consumer warps take 4 rows of every tile each.
// The chunk id travels to the consumers through a small shared-memory queue that is written
// before the chunk's first tile is armed (mbarrier release/acquire orders it).
// ---------------------------------------------------------------------------
template <int SRC, bool VSCOPE, bool SURFACE>
__global__ void __launch_bounds__(kTmaWarps * 32 + 32, 1)
	scope_strip_kernel_tma(const __grid_constant__ StripParams P, const __grid_constant__ CUtensorMap map_rgb,
			       const __grid_constant__ CUtensorMap map_yuv)
{
	using L = SmemLayout<SRC, VSCOPE, SURFACE, true>;
	constexpr int NW = kTmaWarps, RPW = kTileRows / NW;
	constexpr int kStages = L::kStages;
	extern __shared__ __align__(128) uint8_t smem[];
	uint32_t *vs = reinterpret_cast<uint32_t *>(smem + L::kVsOff);
	uint32_t *wave0 = reinterpret_cast<uint32_t *>(smem + L::kWave0Off);
	volatile uint32_t *chunk_q = reinterpret_cast<volatile uint32_t *>(smem + L::kQueueOff); // kQueue entries
	const uint32_t smem_base = smem_u32(smem);
	const uint32_t bar_full = smem_base + L::kBarOff;     // kStages x 8 B
	const uint32_t bar_empty = bar_full + kMaxStages * 8; // kStages x 8 B

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const bool is_producer = warp == NW;

	if (!is_producer)
		zero_bins<NW>(vs, wave0, VSCOPE, SRC != SRC_NONE, tid);
	if (tid == 0) {
		for (int s = 0; s < kStages; s++) {
			mbar_init(bar_full + 8 * s, 1);
			mbar_init(bar_empty + 8 * s, NW);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const uint32_t tiles = (P.height + kTileRows - 1) / kTileRows;
	const uint32_t n_chunks = (P.items + P.chunk_items - 1) / P.chunk_items;

	if (is_producer) {
		// ================= TMA producer (one elected lane) =================
		if (lane == 0) {
			uint32_t stage = 0, phase = 0, qw = 0;
			for (;;) {
				const uint32_t chunk = atomicAdd(P.chunk_counter, 1u);
				const bool done = chunk >= n_chunks;
				// announce the chunk (or the end) before its first tile can complete
				mbar_wait(bar_empty + 8 * stage, phase ^ 1);
				chunk_q[qw % kQueue] = done ? 0xFFFFFFFFu : chunk;
				qw++;
				if (done) {
					mbar_arrive(bar_full + 8 * stage); // wake the consumers with no data
					break;
				}
				const uint32_t first = chunk * P.chunk_items, last = min(first + P.chunk_items, P.items);
				for (uint32_t item = first; item < last; item++) {
					const uint32_t frame = item / P.strips, strip = item - frame * P.strips;
					const int x = (int)(strip * kStripPx);
					for (uint32_t t = 0; t < tiles; t++) {
						if (!(item == first && t == 0))
							mbar_wait(bar_empty + 8 * stage, phase ^ 1);
						const uint32_t dst = smem_base + L::kStageOff + stage * L::kStageBytes;
						mbar_expect_tx(bar_full + 8 * stage, L::kStageBytes);
						if (L::kLoadRgb)
							tma_load_3d(dst, &map_rgb, bar_full + 8 * stage, x, (int)(t * kTileRows),
								    (int)frame);
						if (L::kLoadYuv)
							tma_load_3d(dst + (L::kLoadRgb ? kTileBytes : 0), &map_yuv,
								    bar_full + 8 * stage, x, (int)(t * kTileRows), (int)frame);
						if (++stage == kStages) {
							stage = 0;
							phase ^= 1;
						}
					}
				}
			}
		}
		return;
	}

	// ================= consumers =================
	const uint32_t wave_lane_addr = smem_base + L::kWave0Off + lane * 4;
	uint32_t magic; // 0x4B000000 kept in a register so PRMT can take the selector as its immediate
	asm volatile("mov.u32 %0, 0x4B000000;" : "=r"(magic));
	const TileCtx tc{smem_base + L::kVsOff, wave_lane_addr - 0x80000000u,
			 wave_lane_addr - 0x80000000u + kWaveWords * 4, magic, P.bins_mask, lane};
	const Coef coef = P.coef;
	uint32_t zero; // a 0 the compiler cannot see through (used to build data dependencies)
	asm volatile("mov.u32 %0, 0;" : "=r"(zero));
	uint32_t stage = 0, phase = 0, qr = 0;
	uint32_t cur_frame = 0xFFFFFFFFu;

	for (;;) {
		// the chunk id becomes readable once the chunk's first tile (or the end marker) lands
		mbar_wait(bar_full + 8 * stage, phase);
		const uint32_t chunk = chunk_q[qr % kQueue];
		qr++;
		if (chunk == 0xFFFFFFFFu)
			break;
		const uint32_t first = chunk * P.chunk_items, last = min(first + P.chunk_items, P.items);
		for (uint32_t item = first; item < last; item++) {
			const uint32_t frame = item / P.strips, strip = item - frame * P.strips;
			if (VSCOPE && frame != cur_frame && cur_frame != 0xFFFFFFFFu)
				flush_vscope<NW>(P, vs, cur_frame, tid);
			cur_frame = frame;
			const uint32_t x = strip * kStripPx + lane;
			const bool lane_ok = x < P.width;
			const bool strip_full = strip * kStripPx + kStripPx <= P.width;

			// tiles [0, n_full) lie completely inside the frame (uniform over the CTA)
			const uint32_t n_full = strip_full ? P.height / kTileRows : 0u;
			bool skip_wait = item == first; // the chunk announcement already waited for tile 0
			// fetch = wait for the tile, read this thread's pixels, remember which stage to hand back
			auto fetch_tile = [&](uint32_t(&p)[RPW], uint32_t(&q)[RPW]) -> uint32_t {
				if (!skip_wait)
					mbar_wait(bar_full + 8 * stage, phase);
				skip_wait = false;
				const uint32_t *tile =
					reinterpret_cast<const uint32_t *>(smem + L::kStageOff + stage * L::kStageBytes) +
					warp * RPW * kStripPx + lane;
#pragma unroll
				for (int k = 0; k < RPW; k++) {
					p[k] = L::kLoadRgb ? tile[k * kStripPx] : 0u;
					q[k] = L::kLoadYuv ? tile[k * kStripPx + (L::kLoadRgb ? kTileBytes / 4 : 0)] : 0u;
				}
				const uint32_t bar = bar_empty + 8 * stage;
				if (++stage == kStages) {
					stage = 0;
					phase ^= 1;
				}
				return bar;
			};
			// release = hand the stage back to the producer.  The barrier address is made to
			// depend on the loaded pixels (`& zero`, an opaque 0) so the arrive cannot be issued
			// before the LDS results are in registers: under a backlog of serialised atomics the
			// LSU can otherwise still be holding those reads when the TMA refill lands (seen as
			// rare wrong-bin pixels on smooth content; profiles/ubench_r01.md, "WAR on the ring").
			auto release_tile = [&](uint32_t bar, const uint32_t(&p)[RPW], const uint32_t(&q)[RPW]) {
				uint32_t dep = 0;
#pragma unroll
				for (int k = 0; k < RPW; k++)
					dep |= p[k] | q[k];
				__syncwarp();
				if (lane == 0)
					mbar_arrive(bar + (dep & zero));
			};
			uint32_t t = 0;
			if (n_full > 0) {
				// software pipeline over the interior tiles: atomics of tile t next to the
				// arithmetic of tile t+1 (two Prep register sets, ping-pong)
				uint32_t p[RPW], q[RPW];
				Prep<RPW> A, B;
				uint32_t bar = fetch_tile(p, q);
				release_tile(bar, p, q);
				prepare_tile<SRC, VSCOPE, SURFACE, RPW>(tc, coef, p, q, A);
				for (t = 1; t + 1 < n_full; t += 2) {
					bar = fetch_tile(p, q);
					commit_tile<SRC, VSCOPE, SURFACE, RPW>(tc, A);
					release_tile(bar, p, q);
					prepare_tile<SRC, VSCOPE, SURFACE, RPW>(tc, coef, p, q, B);
					bar = fetch_tile(p, q);
					commit_tile<SRC, VSCOPE, SURFACE, RPW>(tc, B);
					release_tile(bar, p, q);
					prepare_tile<SRC, VSCOPE, SURFACE, RPW>(tc, coef, p, q, A);
				}
				if (t < n_full) {
					bar = fetch_tile(p, q);
					commit_tile<SRC, VSCOPE, SURFACE, RPW>(tc, A);
					release_tile(bar, p, q);
					prepare_tile<SRC, VSCOPE, SURFACE, RPW>(tc, coef, p, q, B);
					commit_tile<SRC, VSCOPE, SURFACE, RPW>(tc, B);
				} else {
					commit_tile<SRC, VSCOPE, SURFACE, RPW>(tc, A);
				}
				t = n_full;
			}
			for (; t < tiles; t++) {
				// edge tiles (last rows, last strip) and partial channel masks: generic body
				uint32_t p[RPW], q[RPW];
				bool ok[RPW];
				const uint32_t bar = fetch_tile(p, q);
				release_tile(bar, p, q);
				const uint32_t y0 = t * kTileRows + warp * RPW;
#pragma unroll
				for (int k = 0; k < RPW; k++)
					ok[k] = lane_ok && (y0 + k < P.height);
				process_tile<SRC, VSCOPE, SURFACE, false, RPW>(tc, coef, p, q, ok);
			}
			if (SRC != SRC_NONE)
				emit_strip<NW>(P, wave0, frame, x, lane_ok, warp, lane);
		}
	}
	if (VSCOPE && cur_frame != 0xFFFFFFFFu)
		flush_vscope<NW>(P, vs, cur_frame, tid);
}

// ---------------------------------------------------------------------------
// strip kernel, plain-load fallback for planes TMA cannot describe (base or pitch not a
// multiple of 16 bytes, e.g. an ROI crop at an odd column).  Same accumulation code; each
// thread simply loads its own pixels (128 B per warp-row) right before using them.
// ---------------------------------------------------------------------------
template <int SRC, bool VSCOPE, bool SURFACE>
__global__ void __launch_bounds__(kLdgWarps * 32, 1) scope_strip_kernel_ldg(const __grid_constant__ StripParams P)
{
	using L = SmemLayout<SRC, VSCOPE, SURFACE, false>;
	constexpr int NW = kLdgWarps, RPW = kLdgRows;
	extern __shared__ __align__(128) uint8_t smem[];
	uint32_t *vs = reinterpret_cast<uint32_t *>(smem + L::kVsOff);
	uint32_t *wave0 = reinterpret_cast<uint32_t *>(smem + L::kWave0Off);
	const uint32_t smem_base = smem_u32(smem);
	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;

	zero_bins<NW>(vs, wave0, VSCOPE, SRC != SRC_NONE, tid);
	__syncthreads();

	const uint32_t first = blockIdx.x * P.items_per_cta;
	const uint32_t last = min(first + P.items_per_cta, P.items);
	const uint32_t groups = (P.height + RPW - 1) / RPW;

	const uint32_t wave_lane_addr = smem_base + L::kWave0Off + lane * 4;
	uint32_t magic;
	asm volatile("mov.u32 %0, 0x4B000000;" : "=r"(magic));
	const TileCtx tc{smem_base + L::kVsOff, wave_lane_addr - 0x80000000u,
			 wave_lane_addr - 0x80000000u + kWaveWords * 4, magic, P.bins_mask, lane};
	const Coef coef = P.coef;
	uint32_t cur_frame = 0xFFFFFFFFu;

	for (uint32_t item = first; item < last; item++) {
		const uint32_t frame = item / P.strips, strip = item - frame * P.strips;
		if (VSCOPE && frame != cur_frame && cur_frame != 0xFFFFFFFFu)
			flush_vscope<NW>(P, vs, cur_frame, tid);
		cur_frame = frame;
		const uint32_t x = strip * kStripPx + lane;
		const bool lane_ok = x < P.width;
		const bool strip_full = strip * kStripPx + kStripPx <= P.width;
		const size_t col = (size_t)frame * P.frame_stride + (size_t)(lane_ok ? x : 0) * 4;

		for (uint32_t g = warp; g < groups; g += NW) {
			const uint32_t y0 = g * RPW;
			const bool full = strip_full && (y0 + RPW <= P.height);
			uint32_t p[RPW], q[RPW];
			bool ok[RPW];
#pragma unroll
			for (int k = 0; k < RPW; k++) {
				ok[k] = full || (lane_ok && (y0 + k < P.height));
				p[k] = 0;
				q[k] = 0;
				if (ok[k]) {
					const size_t o = col + (size_t)(y0 + k) * P.linesize;
					if (L::kLoadRgb)
						p[k] = ld_nc_u32(P.rgb + o);
					if (L::kLoadYuv)
						q[k] = ld_nc_u32(P.yuv + o);
				}
			}
			if (full && P.bins_mask == 7u)
				process_tile<SRC, VSCOPE, SURFACE, true, RPW>(tc, coef, p, q, ok);
			else
				process_tile<SRC, VSCOPE, SURFACE, false, RPW>(tc, coef, p, q, ok);
		}
		if (SRC != SRC_NONE)
			emit_strip<NW>(P, wave0, frame, x, lane_ok, warp, lane);
	}
	if (VSCOPE && cur_frame != 0xFFFFFFFFu)
		flush_vscope<NW>(P, vs, cur_frame, tid);
}

// ---------------------------------------------------------------------------
// finalize kernels
// ---------------------------------------------------------------------------
// display mapping of a bin image (vectorscope.effect:30-31, waveform.effect:33-36) as
// pinned in oracle/scope_oracle.c: r = (c/255)*k, min(r,1), floor(fma(r,255,0.5))
__device__ __forceinline__ uint32_t intensity_u8(uint32_t c, float k)
{
	float r = __fmul_rn(__fdiv_rn((float)c, 255.0f), k);
	r = fminf(r, 1.0f);
	return (uint32_t)floorf(__fmaf_rn(r, 255.0f, 0.5f));
}

// vectorscope: u32 accumulators -> saturated u8 (+ optional display image)
__global__ void __launch_bounds__(256) vscope_finalize_kernel(const uint32_t *acc, unsigned long long acc_stride,
							      uint8_t *out, uint8_t *display,
							      unsigned long long out_stride, float intensity)
{
	const size_t f = blockIdx.y;
	const int i = (blockIdx.x * 256 + threadIdx.x) * 4;
	const uint4 a = *reinterpret_cast<const uint4 *>(acc + f * acc_stride + i);
	const uint32_t c0 = min(a.x, 255u), c1 = min(a.y, 255u), c2 = min(a.z, 255u), c3 = min(a.w, 255u);
	if (out)
		*reinterpret_cast<uint32_t *>(out + f * out_stride + i) = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
	if (display)
		*reinterpret_cast<uint32_t *>(display + f * out_stride + i) =
			intensity_u8(c0, intensity) | (intensity_u8(c1, intensity) << 8) |
			(intensity_u8(c2, intensity) << 16) | (intensity_u8(c3, intensity) << 24);
}

// histogram hi_max (histogram.c:330-355,397-402); one block of 256 threads per frame
__global__ void __launch_bounds__(256) hist_max_kernel(const uint32_t *hist, unsigned long long hist_stride,
						       uint32_t *hi_max, uint32_t components, uint32_t width,
						       uint32_t height, int level_fixed, int level_ratio)
{
	__shared__ uint32_t red[3][8];
	const size_t f = blockIdx.x;
	const uint32_t *h = hist + f * hist_stride;
	const uint32_t mask[3] = {0x44u, 0x22u, 0x11u};
#pragma unroll
	for (int j = 0; j < 3; j++) {
		uint32_t v = (components & mask[j]) ? h[threadIdx.x * 4 + j] : 0u;
		v = __reduce_max_sync(0xFFFFFFFFu, v);
		if ((threadIdx.x & 31) == 0)
			red[j][threadIdx.x >> 5] = v;
	}
	__syncthreads();
	if (threadIdx.x < 3) {
		uint32_t v = 1;
		for (int w = 0; w < 8; w++)
			v = max(v, red[threadIdx.x][w]);
		if (level_fixed > 0)
			v = (uint32_t)level_fixed;
		else if (level_ratio > 0) {
			v = (uint32_t)((unsigned long long)width * height * (unsigned long long)level_ratio / 1000ull);
			if (v == 0)
				v = 1;
		}
		hi_max[f * 4 + threadIdx.x] = v;
	}
	if (threadIdx.x == 3)
		hi_max[f * 4 + 3] = 0;
}

// waveform display image (intensity applied) from the final u8 waveform
__global__ void __launch_bounds__(256) wave_display_kernel(const uint8_t *wave, uint8_t *display, size_t n_words,
							   float intensity)
{
	const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n_words)
		return;
	const uint32_t w = reinterpret_cast<const uint32_t *>(wave)[i];
	reinterpret_cast<uint32_t *>(display)[i] = intensity_u8(w & 0xFF, intensity) |
						   (intensity_u8((w >> 8) & 0xFF, intensity) << 8) |
						   (intensity_u8((w >> 16) & 0xFF, intensity) << 16);
}

// partial (tile-sharded) waveform: summed u16 pairs (two planes) -> saturated u8 BGRX
__global__ void __launch_bounds__(256) wave_pairs_finalize_kernel(const uint32_t *pairs, uint8_t *wave, size_t n_px)
{
	const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n_px)
		return;
	const uint32_t w0 = pairs[i], w1 = pairs[n_px + i];
	reinterpret_cast<uint32_t *>(wave)[i] =
		min(w0 & 0xFFFFu, 255u) | (min(w0 >> 16, 255u) << 8) | (min(w1 & 0xFFFFu, 255u) << 16);
}

// test hook: the kernel's own transform for all 2^24 colours (index r<<16|g<<8|b)
__global__ void __launch_bounds__(256) yuv_table_kernel(Coef coef, uint32_t *out)
{
	const uint32_t i = (blockIdx.x * 256 + threadIdx.x) * 2;
	// index r<<16|g<<8|b is already the little-endian BGRA word b | g<<8 | r<<16
	uint32_t magic;
	asm volatile("mov.u32 %0, 0x4B000000;" : "=r"(magic));
	const uint32_t ca[3] = {carrier<0>(i, magic), carrier<1>(i, magic), carrier<2>(i, magic)};
	const uint32_t cb[3] = {carrier<0>(i + 1, magic), carrier<1>(i + 1, magic), carrier<2>(i + 1, magic)};
	uint32_t ya[3], yb[3];
	rgb_to_yuv_pair<true>(ca, cb, coef, ya, yb);
	out[i] = (ya[0] & 0xFFu) | ((ya[1] & 0xFFu) << 8) | ((ya[2] & 0xFFu) << 16);
	out[i + 1] = (yb[0] & 0xFFu) | ((yb[1] & 0xFFu) << 8) | ((yb[2] & 0xFFu) << 16);
}

} // namespace scope
