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
// Pixels arrive through TMA (cp.async.bulk.tensor) into a 4-stage shared-memory ring
// filled by a producer warp; consumers read one 32-bit pixel per lane per row.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace scope {

constexpr int kStripPx = 32;          // columns per strip == lanes per warp
constexpr int kConsumerWarps = 16;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kTileRows = 64;         // rows per TMA stage (4 rows per consumer warp)
constexpr int kRowsPerWarp = kTileRows / kConsumerWarps;
constexpr int kStages = 4;
constexpr int kTileBytes = kStripPx * 4 * kTileRows; // 8 KB per plane per stage
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
	uint32_t items_per_cta;
	uint32_t bins_mask;       // channels accumulated into the column bins: bit0 B|U, bit1 G|Y, bit2 R|V
	uint32_t hist_mask;       // channels the histogram output wants
	uint32_t wave_mask;       // channels the waveform output wants
	uint32_t x_offset;        // first output column (tile-sharded frames)
	uint32_t out_width;       // row length of the waveform output in pixels
	uint32_t partial;         // 1: add u16 pairs into wave_pairs instead of writing u8
	uint32_t tma_x0;          // pixel offset of column 0 inside the tensor map
	uint32_t *hist;           // [n][1024] u32, zeroed
	uint8_t *wave;            // [n][256][out_width][4]
	uint32_t *wave_pairs;     // partial: [256][out_width][2]
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
		     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
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
__device__ __forceinline__ void consumer_bar()
{
	asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
}
// predicated shared-memory reduction: no branch, no return value
__device__ __forceinline__ void red_shared_if(uint32_t addr, uint32_t val, bool pred)
{
	asm volatile("{\n"
		     ".reg .pred p;\n"
		     "setp.ne.u32 p, %2, 0;\n"
		     "@p red.shared.add.u32 [%0], %1;\n"
		     "}" ::"r"(addr),
		     "r"(val), "r"((uint32_t)pred)
		     : "memory");
}
__device__ __forceinline__ uint32_t atom_shared_add(uint32_t addr, uint32_t val)
{
	uint32_t old;
	asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(val) : "memory");
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

// byte k of the two pixels -> exact floats (2^23 + b, then - 2^23)
template <int K>
__device__ __forceinline__ f32x2 bytes_to_f32x2(uint32_t pa, uint32_t pb)
{
	const uint32_t magic = 0x4B000000u; // 8388608.0f
	const uint32_t fa = __byte_perm(pa, magic, 0x7650 + K);
	const uint32_t fb = __byte_perm(pb, magic, 0x7650 + K);
	return add2(pack2(__uint_as_float(fa), __uint_as_float(fb)), splat2(-8388608.0f));
}

// x / 255 correctly rounded for x in {0..255}: fma(x, k0, rn(x*k1)), k0 = rn(1/255),
// k1 = rn(1/255 - k0).  Checked against IEEE division for all 256 inputs
// (tests/test_transform_math.py) and, through the kernel, for all 2^24 colours.
__device__ __forceinline__ f32x2 div255(f32x2 x)
{
	const f32x2 k0 = splat2(__uint_as_float(0x3B808081u)); // 0x1.010102p-8
	const f32x2 k1 = splat2(__uint_as_float(0xAF7EFEFFu)); // -0x1.fdfdfep-33
	return fma2(x, k0, mul2(x, k1));
}

// one output channel for two pixels: p = c0*r; p = fma(c1,g,p); p = fma(c2,b,p); t = p + off;
// q = floor(fma(t, 255, 0.5)).  Returns the two floats 2^23 + q (low byte of the bit
// pattern = q).  The [0,1] clamp of the definition never acts (exhaustively verified).
__device__ __forceinline__ f32x2 yuv_channel(f32x2 r, f32x2 g, f32x2 b, float c0, float c1, float c2, float off)
{
	f32x2 p = mul2(splat2(c0), r);
	p = fma2(splat2(c1), g, p);
	p = fma2(splat2(c2), b, p);
	p = add2(p, splat2(off));
	p = fma2(p, splat2(255.0f), splat2(0.5f));
	return add2_rz(p, splat2(8388608.0f));
}

// BGRA pixel pair -> [U,Y,V,255] pixel pair (data/common.effect:23-43 as pinned in
// oracle/scope_oracle.c).  NEED_Y = false leaves the Y byte 0 (vectorscope only).
template <bool NEED_Y>
__device__ __forceinline__ void rgb_to_yuv_pair(uint32_t pa, uint32_t pb, const Coef &c, uint32_t &qa, uint32_t &qb)
{
	const f32x2 b = div255(bytes_to_f32x2<0>(pa, pb));
	const f32x2 g = div255(bytes_to_f32x2<1>(pa, pb));
	const f32x2 r = div255(bytes_to_f32x2<2>(pa, pb));
	uint32_t ua, ub, va, vb;
	unpack2(yuv_channel(r, g, b, c.u0, c.u1, c.u2, 0.5f - 1.0f / 256.0f), ua, ub);
	unpack2(yuv_channel(r, g, b, c.v0, c.v1, c.v2, 0.5f), va, vb);
	if (NEED_Y) {
		uint32_t ya, yb;
		unpack2(yuv_channel(r, g, b, c.y0, c.y1, c.y2, 0.0f), ya, yb);
		// byte0 = u.b0, byte1 = y.b0 ; then byte2 = v.b0, byte3 = 0xFF
		const uint32_t ta = __byte_perm(ua, ya, 0x0040), tb = __byte_perm(ub, yb, 0x0040);
		qa = __byte_perm(ta, va | 0xFF00u, 0x5410);
		qb = __byte_perm(tb, vb | 0xFF00u, 0x5410);
	} else {
		// 0x4B0000vv has zero bytes 1,2: byte0 = u.b0, byte1 = 0, byte2 = v.b0, byte3 = 0xFF
		qa = __byte_perm(ua, va | 0xFF00u, 0x5420);
		qb = __byte_perm(ub, vb | 0xFF00u, 0x5420);
	}
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
	static constexpr int kBarOff = kStageOff + kStages * kStageBytes;
	static constexpr int kTotal = kBarOff + (USE_TMA ? 2 * kStages * 8 : 0) + 16;
};

// ---------------------------------------------------------------------------
// per-pixel accumulation
// ---------------------------------------------------------------------------
// waveform/histogram column bins.  plane0[level][lane] = (count B|U : lo16, count G|Y : hi16),
// plane1[level][lane] = (count R|V : lo16).  A strip has <= 65535 rows, so no half overflows.
__device__ __forceinline__ void bins_add(uint32_t s, uint32_t wave_lane_addr, bool ok, uint32_t mask)
{
	ok = ok && (s > 0x00FFFFFFu); // alpha != 0 (histogram.c:385-387, waveform.c:246-248)
	const uint32_t ab = wave_lane_addr + ((s & 0xFFu) << 7);
	const uint32_t ag = wave_lane_addr + ((s >> 1) & 0x7F80u);
	const uint32_t ar = wave_lane_addr + kWaveWords * 4 + ((s >> 9) & 0x7F80u);
	red_shared_if(ab, 1u, ok && (mask & 1u));
	red_shared_if(ag, 0x10000u, ok && (mask & 2u));
	red_shared_if(ar, 1u, ok && (mask & 4u));
}

// vectorscope bin index of a [U,Y,V,A] pixel: row = 255 - V, column = U (vectorscope.c:232)
__device__ __forceinline__ uint32_t vs_index(uint32_t q)
{
	return __byte_perm(q, ~q, 0x4460) & 0xFFFFu; // byte0 = q.b0 (U), byte1 = (~q).b2 (255-V)
}

// add `k` to bin idx (u16 halves, two bins per word).  When a half crosses 0x8000 the
// add that crossed subtracts 0x4000 again: the bin stays > 255 (it saturates to 255 at
// the end, like the reference's `if (*c < 255) ++*c`) and can never wrap 16 bits.
__device__ __forceinline__ void vs_add(uint32_t vs_base, uint32_t idx, uint32_t k)
{
	const uint32_t addr = vs_base + ((idx << 1) & ~3u);
	const uint32_t sh = (idx & 1u) << 4;
	const uint32_t add = k << sh;
	const uint32_t old = atom_shared_add(addr, add);
	if (((old ^ (old + add)) & (0x8000u << sh)) != 0u)
		atom_shared_add(addr, 0u - (0x4000u << sh));
}

// ---------------------------------------------------------------------------
// the strip kernel
// ---------------------------------------------------------------------------
template <int SRC, bool VSCOPE, bool SURFACE, bool USE_TMA>
__global__ void __launch_bounds__(kConsumerThreads + (USE_TMA ? 32 : 0), 1)
	scope_strip_kernel(const __grid_constant__ StripParams P, const __grid_constant__ CUtensorMap map_rgb,
			   const __grid_constant__ CUtensorMap map_yuv)
{
	using L = SmemLayout<SRC, VSCOPE, SURFACE, USE_TMA>;
	extern __shared__ __align__(128) uint8_t smem[];
	uint32_t *vs = reinterpret_cast<uint32_t *>(smem + L::kVsOff);
	uint32_t *wave0 = reinterpret_cast<uint32_t *>(smem + L::kWave0Off);
	const uint32_t smem_base = smem_u32(smem);
	const uint32_t bar_full = smem_base + L::kBarOff;          // kStages x 8 B
	const uint32_t bar_empty = bar_full + kStages * 8;         // kStages x 8 B

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const bool is_producer = USE_TMA && warp == kConsumerWarps;

	// ---- one-time setup: zero the bins, init barriers ----
	if (!is_producer) {
		if (VSCOPE)
			for (int i = tid; i < kVsWords / 4; i += kConsumerThreads)
				reinterpret_cast<uint4 *>(vs)[i] = make_uint4(0, 0, 0, 0);
		if (SRC != SRC_NONE)
			for (int i = tid; i < 2 * kWaveWords / 4; i += kConsumerThreads)
				reinterpret_cast<uint4 *>(wave0)[i] = make_uint4(0, 0, 0, 0);
	}
	if (USE_TMA && tid == 0) {
		for (int s = 0; s < kStages; s++) {
			mbar_init(bar_full + 8 * s, 1);
			mbar_init(bar_empty + 8 * s, kConsumerWarps);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const uint32_t first = blockIdx.x * P.items_per_cta;
	const uint32_t last = min(first + P.items_per_cta, P.items);
	const uint32_t tiles = (P.height + kTileRows - 1) / kTileRows;

	if (is_producer) {
		// ================= TMA producer (one elected lane) =================
		if (lane == 0) {
			uint32_t stage = 0, phase = 0;
			for (uint32_t item = first; item < last; item++) {
				const uint32_t frame = item / P.strips, strip = item - frame * P.strips;
				const int x = (int)(P.tma_x0 + strip * kStripPx);
				for (uint32_t t = 0; t < tiles; t++) {
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
		return;
	}

	// ================= consumers =================
	const uint32_t vs_base = smem_base + L::kVsOff;
	const uint32_t wave_lane_addr = smem_base + L::kWave0Off + lane * 4;
	uint32_t stage = 0, phase = 0;
	uint32_t cur_frame = 0xFFFFFFFFu;

	auto flush_vscope = [&](uint32_t frame) {
		// every consumer thread: move its slice of the u16 pairs to the frame's u32 accumulators
		consumer_bar();
		uint32_t *acc = P.vscope_acc + (size_t)frame * P.vscope_stride;
		for (int i = tid; i < kVsWords / 4; i += kConsumerThreads) {
			uint4 w = reinterpret_cast<uint4 *>(vs)[i];
			if ((w.x | w.y | w.z | w.w) != 0u) {
				const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
				for (int j = 0; j < 4; j++) {
					if (ww[j] & 0xFFFFu)
						atomicAdd(acc + (i * 4 + j) * 2, ww[j] & 0xFFFFu);
					if (ww[j] >> 16)
						atomicAdd(acc + (i * 4 + j) * 2 + 1, ww[j] >> 16);
				}
				reinterpret_cast<uint4 *>(vs)[i] = make_uint4(0, 0, 0, 0);
			}
		}
		consumer_bar();
	};

	for (uint32_t item = first; item < last; item++) {
		const uint32_t frame = item / P.strips, strip = item - frame * P.strips;
		if (VSCOPE && frame != cur_frame && cur_frame != 0xFFFFFFFFu)
			flush_vscope(cur_frame);
		cur_frame = frame;
		const uint32_t x = strip * kStripPx + lane;
		const bool lane_ok = x < P.width;
		const uint8_t *rgb_px = nullptr, *yuv_px = nullptr;
		if (!USE_TMA) {
			rgb_px = P.rgb + (size_t)frame * P.frame_stride + (size_t)(lane_ok ? x : 0) * 4;
			yuv_px = P.yuv + (size_t)frame * P.frame_stride + (size_t)(lane_ok ? x : 0) * 4;
		}

		for (uint32_t t = 0; t < tiles; t++) {
			const uint32_t y0 = t * kTileRows + warp * kRowsPerWarp;
			uint32_t p[kRowsPerWarp], q[kRowsPerWarp];
			bool ok[kRowsPerWarp];
#pragma unroll
			for (int k = 0; k < kRowsPerWarp; k++) {
				ok[k] = lane_ok && (y0 + k < P.height);
				p[k] = 0;
				q[k] = 0;
			}
			if (USE_TMA) {
				mbar_wait(bar_full + 8 * stage, phase);
				const uint32_t *tile = reinterpret_cast<const uint32_t *>(
					smem + L::kStageOff + stage * L::kStageBytes);
#pragma unroll
				for (int k = 0; k < kRowsPerWarp; k++) {
					const int o = (warp * kRowsPerWarp + k) * kStripPx + lane;
					if (L::kLoadRgb)
						p[k] = tile[o];
					if (L::kLoadYuv)
						q[k] = tile[o + (L::kLoadRgb ? kTileBytes / 4 : 0)];
				}
				__syncwarp();
				if (lane == 0)
					mbar_arrive(bar_empty + 8 * stage);
				if (++stage == kStages) {
					stage = 0;
					phase ^= 1;
				}
			} else {
#pragma unroll
				for (int k = 0; k < kRowsPerWarp; k++) {
					if (ok[k]) {
						if (L::kLoadRgb)
							p[k] = ld_nc_u32(rgb_px + (size_t)(y0 + k) * P.linesize);
						if (L::kLoadYuv)
							q[k] = ld_nc_u32(yuv_px + (size_t)(y0 + k) * P.linesize);
					}
				}
			}

			// ---- colour transform in registers (fused mode) ----
			if (!SURFACE && (VSCOPE || SRC == SRC_YUV)) {
#pragma unroll
				for (int k = 0; k < kRowsPerWarp; k += 2)
					rgb_to_yuv_pair<SRC == SRC_YUV>(p[k], p[k + 1], P.coef, q[k], q[k + 1]);
			}

			// ---- waveform / histogram column bins ----
			if (SRC != SRC_NONE) {
#pragma unroll
				for (int k = 0; k < kRowsPerWarp; k++)
					bins_add(SRC == SRC_RGB ? p[k] : q[k], wave_lane_addr, ok[k], P.bins_mask);
			}

			// ---- vectorscope ----
			if (VSCOPE) {
				uint32_t idx[kRowsPerWarp];
#pragma unroll
				for (int k = 0; k < kRowsPerWarp; k++)
					idx[k] = vs_index(q[k]);
				bool same = ok[0];
#pragma unroll
				for (int k = 1; k < kRowsPerWarp; k++)
					same = same && ok[k] && (idx[k] == idx[0]);
				// (the shuffle must be executed by every lane: no short-circuit around it)
				const uint32_t idx_lane0 = __shfl_sync(0xFFFFFFFFu, idx[0], 0);
				same = same && (idx[0] == idx_lane0);
				if (__all_sync(0xFFFFFFFFu, same)) {
					// flat block: the whole 4x32 block hits one bin -> one atomic
					if (lane == 0)
						vs_add(vs_base, idx[0], 32u * kRowsPerWarp);
				} else {
#pragma unroll
					for (int k = 0; k < kRowsPerWarp; k++)
						if (ok[k])
							vs_add(vs_base, idx[k], 1u);
				}
			}
		}

		// ---- end of strip: emit this strip's waveform columns + histogram share ----
		if (SRC != SRC_NONE) {
			consumer_bar();
			uint32_t *hist = P.hist + (size_t)frame * P.hist_stride;
			const uint32_t xo = P.x_offset + x;
			for (int v = warp; v < 256; v += kConsumerWarps) {
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
							atomicAdd(P.wave_pairs + o * 2, mb | (mg << 16));
						if (mr)
							atomicAdd(P.wave_pairs + o * 2 + 1, mr);
					} else {
						uint32_t *dst = reinterpret_cast<uint32_t *>(
							P.wave + (size_t)frame * P.wave_stride);
						dst[o] = min(mb, 255u) | (min(mg, 255u) << 8) | (min(mr, 255u) << 16);
					}
				}
			}
			consumer_bar();
		}
	}
	if (VSCOPE && cur_frame != 0xFFFFFFFFu)
		flush_vscope(cur_frame);
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

// partial (tile-sharded) waveform: summed u16 pairs -> saturated u8 BGRX
__global__ void __launch_bounds__(256) wave_pairs_finalize_kernel(const uint32_t *pairs, uint8_t *wave, size_t n_px)
{
	const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n_px)
		return;
	const uint2 w = reinterpret_cast<const uint2 *>(pairs)[i];
	reinterpret_cast<uint32_t *>(wave)[i] =
		min(w.x & 0xFFFFu, 255u) | (min(w.x >> 16, 255u) << 8) | (min(w.y & 0xFFFFu, 255u) << 16);
}

// test hook: the kernel's own transform for all 2^24 colours (index r<<16|g<<8|b)
__global__ void __launch_bounds__(256) yuv_table_kernel(Coef coef, uint32_t *out)
{
	const uint32_t i = (blockIdx.x * 256 + threadIdx.x) * 2;
	// index r<<16|g<<8|b is already the little-endian BGRA word b | g<<8 | r<<16
	const uint32_t pa = i | 0xFF000000u, pb = (i + 1) | 0xFF000000u;
	uint32_t qa, qb;
	rgb_to_yuv_pair<true>(pa, pb, coef, qa, qb);
	out[i] = qa & 0xFFFFFFu;
	out[i + 1] = qb & 0xFFFFFFu;
}

} // namespace scope
