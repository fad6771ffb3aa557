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
// by a producer warp; a consumer warp reads its 4 rows of a tile with one ldmatrix.x4 (one pixel
// per lane per row).  A plain-load kernel covers planes TMA cannot describe.
// The pass is bound by instruction issue, not by HBM or the LSU: every
// instruction in the steady-state loop of tma_consume counts.
#pragma once
#include <cstdint>
// SCOPE_EMULATE: this file compiled for the HOST on top of tools/simt/cuda_emul.h (every CUDA thread a
// coroutine; tests/test_kernel_emulation.py checks the kernels' logic against the oracle without a GPU).
// The product never defines it; each PTX helper below has its emulated twin next to it.
#ifndef SCOPE_EMULATE
#include <cuda.h>
#include <cuda_runtime.h>
#define SCOPE_DYNAMIC_SMEM(name) extern __shared__ __align__(128) uint8_t name[]
#else
#define SCOPE_DYNAMIC_SMEM(name) uint8_t *name = emul::W->cur->cta->smem.data()
#endif

namespace scope {

constexpr int kStripPx = 32;          // columns per strip == lanes per warp
// Kernel features that can be switched off for A/B builds (tools/run_ab.sh builds the variants):
//   SCOPE_LDSM    consumers read their 4 rows of a tile with ONE ldmatrix.x4 instead of four LDS.32
//   SCOPE_XORSWZ  vectorscope bank swizzle by XOR (2 instructions) instead of add-mod-32 (4)
#ifndef SCOPE_LDSM
#define SCOPE_LDSM 1
#endif
#ifndef SCOPE_XORSWZ
#define SCOPE_XORSWZ 1
#endif
//   SCOPE_DEFER   the vectorscope's overflow check of tile t is evaluated after the arithmetic of
//                 tile t+1 instead of right behind the atomics (no wait for their return values)
#ifndef SCOPE_DEFER
#define SCOPE_DEFER 1
#endif
//   SCOPE_FADDR   bin addresses and the vectorscope addend are formed by FFMA on denormals (the
//                 FMA-lite pipe) instead of IMAD / LEA (the half-rate FMA-heavy and ALU pipes)
#ifndef SCOPE_FADDR
#define SCOPE_FADDR 1
#endif
//   SCOPE_FAST_EMIT  end-of-strip write-out with 43 (empty level: 12) instead of 62 instructions per level
#ifndef SCOPE_FAST_EMIT
#define SCOPE_FAST_EMIT 1
#endif
//   SCOPE_RAWFLAT (experiment for round 2, OFF: written after this round's GPU budget was spent, not yet
//                 run on a GPU) a 4 x 32 block whose 128 pixel words are all equal (solid regions,
//                 letterbox bars) is accumulated from ONE pixel: one transform, three column-bin atomics
//                 of +4 per lane, one vectorscope atomic of +128 per warp
#ifndef SCOPE_RAWFLAT
#define SCOPE_RAWFLAT 0
#endif
//   SCOPE_DEEP_RING (experiment for round 2, OFF) the tile ring takes all the shared memory the bins leave
//                 instead of a fixed 32 KB: 8 stages = 64 KB in flight for the vectorscope-only pass
//                 (one CTA per SM, no column bins), 6 stages = 48 KB per CTA for the passes without
//                 the vectorscope (two CTAs per SM); the fused pass has no room to spare.  32 KB in
//                 flight per SM caps a strip reader at ~3.9 TB/s (profiles/ubench_r01.md).
#ifndef SCOPE_DEEP_RING
#define SCOPE_DEEP_RING 0
#endif
//   SCOPE_PIPELINE 0 = no software pipeline over the interior tiles (prepare and commit of a tile back to
//                 back, like the edge tiles): the A/B partner of the shipped commit(t) || prepare(t+1)
#ifndef SCOPE_PIPELINE
#define SCOPE_PIPELINE 1
#endif
//   SCOPE_STRAIGHT (experiment for round 2, OFF) the steady state of an ordinary tile (all pixels counted,
//                 all three channels, block not flat) is ONE basic block: stage handed back first, then
//                 the atomics of tile t and the arithmetic of tile t+1 with no branch between them, so
//                 that ptxas can spread the 16 atomics over the ~100 arithmetic instructions instead of
//                 issuing them as a burst that fills the LSU queue while the other pipes idle
#ifndef SCOPE_STRAIGHT
#define SCOPE_STRAIGHT 0
#endif
//   SCOPE_BALLOT (experiment for round 2, OFF) almost-flat blocks - screen content: a flat background with a little
//                 text in it - put nearly every lane of a warp on ONE vectorscope bin, and k lanes on one word
//                 cost k cycles.  The flat test's vote becomes a ballot; when at least kBallotLanes lanes hold
//                 nothing but the reference bin (lane 0's first pixel), every pixel of the block that hits that
//                 bin is counted with per-row ballots and added by one lane, the rest go their usual way.
#ifndef SCOPE_BALLOT
#define SCOPE_BALLOT 0
#endif
//   SCOPE_IMMCOEF the TMA tile kernels that evaluate the transform are instantiated per colour space and take its
//                 coefficients as compile-time constants (IMAD with an immediate operand) instead of twelve registers
//                 loaded from the launch parameters: 9 registers fewer and no per-visit reload (measured +2 %,
//                 profiles/r02/ab_round2.md; it is what lets 20 warps run the pipelined loop in 80 registers)
#ifndef SCOPE_IMMCOEF
#define SCOPE_IMMCOEF 1
#endif
//   SCOPE_WIDE_FUSED (experiment for round 2, OFF) tile geometry per kernel family: the kernels that hold the
//                 vectorscope (one CTA per SM, 120 registers through __maxnreg__) take tiles of twice the height,
//                 i.e. 8 rows per warp and visit, the kernels that run two CTAs per SM keep 4 rows (56 registers).
//                 What `w16n8_immcoef_r120_x` measures for the fused pass, in a form that could ship.
#ifndef SCOPE_WIDE_FUSED
#define SCOPE_WIDE_FUSED 0
#endif
#if SCOPE_WIDE_FUSED && !defined(SCOPE_MAXNREG)
#define SCOPE_MAXNREG 120
#endif
//   SCOPE_DEPHASE (experiment for round 2, OFF) the measured time of the fused pass is close to the SUM of its
//                 issue cycles and its shared-memory cycles: the warps of a CTA start every strip
//                 together after emit_strip's barrier and then all do arithmetic, then all do atomics.  With this
//                 flag consumer warps 4-7 and 12-15 begin each strip one arithmetic phase late (one named-barrier
//                 handshake per strip), so that half the warps issue atomics while the other half computes.  (Bit 2
//                 of the warp number, not bit 0: warp w runs on scheduler w mod 4, and every scheduler must keep
//                 warps of both halves or it would sit idle during the other half's arithmetic.)
#ifndef SCOPE_DEPHASE
#define SCOPE_DEPHASE 0
#endif
//   SCOPE_FUSED_WARPS consumer warps of the kernels that hold the vectorscope bins and read ONE plane (one CTA per
//                 SM; the fused headline pass is one of them).  The register file is per scheduler (16 K registers for
//                 the warps w with w mod 4 == s): 17 warps (16 + producer) put 5 on one scheduler = at most 96
//                 registers per thread, 21 put 6 = at most 80 - and the loop needs 80 with SCOPE_IMMCOEF.  Measured on
//                 the mixed 4K batch (profiles/r02/ab_round2.md): 16 warps 33.9 %, 18: 33.8 %, 19: 34.7 %, 20: 36.1 %,
//                 21: 35.0 %, 22: 35.8 %, 23: 37.1 % of the HBM peak (23 + the producer = 6 warps on every scheduler;
//                 with a two-stage ring 23 warps fell to 31.5 %: the ring must hold three tiles).  The pass is bound
//                 by per-warp latency, not by instruction count: 8 rows per visit (33 instead of 39 instructions per
//                 pixel-warp, 15 warps, two stages) measured 32.1 %.
#ifndef SCOPE_TMA_WARPS
#define SCOPE_TMA_WARPS 16
#define SCOPE_TMA_WARPS_DEFAULTED 1
#endif
#ifndef SCOPE_FUSED_WARPS
#if defined(SCOPE_TMA_WARPS_DEFAULTED) && !defined(SCOPE_TILE_ROWS)
#define SCOPE_FUSED_WARPS 23
#else
#define SCOPE_FUSED_WARPS SCOPE_TMA_WARPS // an A/B build that names a warp count or a tile height means it for every kernel
#endif
#endif
#ifndef SCOPE_TILE_ROWS
#define SCOPE_TILE_ROWS 64
#endif
constexpr int kTileRows = SCOPE_TILE_ROWS; // rows per TMA tile of the 16-warp kernels (per kernel family: SmemLayout::kTileRows)
constexpr int kTmaWarps = SCOPE_TMA_WARPS; // consumer warps of the TMA kernels that run two CTAs per SM or read two planes
// SCOPE_GROUP_WARPS > 0 selects the row-group kernel (scope_strip_kernel_tmag) with that many
// consumer warps for every TMA launch; 0 keeps the tile-synchronous kernel above
#ifndef SCOPE_GROUP_WARPS
#define SCOPE_GROUP_WARPS 0
#endif
#ifndef SCOPE_SPLIT_VS_WARPS
#define SCOPE_SPLIT_VS_WARPS 16
#endif
#ifndef SCOPE_SPLIT_BIN_WARPS
#define SCOPE_SPLIT_BIN_WARPS 8
#endif
constexpr int kGroupWarps = SCOPE_GROUP_WARPS > 0 ? SCOPE_GROUP_WARPS : 24; // consumer warps of the row-group kernel
constexpr int kGroupRows = 4;                    // rows per group == rows one ldmatrix.x4 reads
constexpr int kSplitVsWarps = SCOPE_SPLIT_VS_WARPS;   // specialised kernel: warps doing transform + vectorscope
constexpr int kSplitBinWarps = SCOPE_SPLIT_BIN_WARPS; // specialised kernel: warps doing the column bins
constexpr int kBallotLanes = 8;            // SCOPE_BALLOT: lanes that must agree before a block takes the aggregating path
constexpr int kRingBytes = 35328;          // shared memory the bins leave for the tile ring: 227 KB - 128 KB (vectorscope) - 64 KB
                                           // (column bins) - barriers and mailbox; 4 stages of 64-row tiles, 3 of 80- to 92-row tiles
constexpr int kMaxStages = 8;              // upper bound of the ring depth (barrier storage)
constexpr int kTileBytes = kStripPx * 4 * kTileRows; // 8 KB per plane per stage
#ifndef SCOPE_MAX_CHUNK
#define SCOPE_MAX_CHUNK 10
#endif
constexpr int kMaxChunkItems = SCOPE_MAX_CHUNK; // upper bound of strips per dynamically claimed chunk
// chunk mailbox entries {first strip, count}.  The producer announces a chunk only after the stage of
// its first tile was handed back, so it is at most kStages chunks ahead of the slowest consumer (chunks
// of one single-tile strip): kQueue >= kStages (tools/ring_model.py checks the mailbox as well).
constexpr int kQueue = SCOPE_DEEP_RING ? 8 : 4;
#ifndef SCOPE_L2_AHEAD
#define SCOPE_L2_AHEAD 0
#endif
constexpr int kL2Ahead = SCOPE_L2_AHEAD;   // tiles the general producer's L2 prefetch runs ahead of its TMA loads.  0: none -
                                           // measured with 6: waveform-only 71.0 -> 64.8 %, histogram-only 66.4 -> 58.8 %,
                                           // vectorscope-only 46.4 -> 43.6 % of the HBM peak (these kernels already keep
                                           // 64 KB in flight per SM, or are bound elsewhere); only scope_fused_kernel_v3,
                                           // whose ring is 35 KB, gains from it (SCOPE_V3_L2_AHEAD)
constexpr int kLdgWarps = 16;              // plain-load fallback kernel
constexpr int kLdgRows = 4;
constexpr int kMaxWaveCopies = 15;    // extra destinations of the final waveform (scope_accumulate_band: <= 16 ranks)
constexpr int kVsWords = 32768;       // 65536 vectorscope bins, two u16 per word
constexpr int kWaveWords = 256 * 32;  // one plane: [level][lane]

enum : int { SRC_NONE = 0, SRC_RGB = 1, SRC_YUV = 2 };

// colour transform constants: 10^6 x the effect file's coefficients, order R, G, B, and the
// rounding/offset constants with the carrier bias folded in (scope_ffi.cu: coef_for)
struct Coef {
	uint32_t u[3], y[3], v[3];
	uint32_t ku, ky, kv;
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
	uint32_t partial;         // 0: write the final u8 waveform; 1: ADD u16 pairs into wave_pairs (several tiles share
	                          // the columns: caller-zeroed accumulators, global atomics); 2: STORE the u16 pairs (this
	                          // launch is the only writer of its columns: no zero-fill, no atomics)
	uint32_t n_wave_copies;   // final u8 waveform also goes to wave_copies[0 .. n): the other ranks' images (peer stores)
	uint32_t tma_x0_rgb, tma_x0_yuv; // TMA kernels: pixel column of the plane's first pixel inside its tensor map
	                                 // (the map starts at the plane pointer rounded down to 16 bytes)
	uint32_t *chunk_counter;  // global work counter of this launch (zeroed by the host)
	uint32_t frame_affine;    // scope_fused_kernel_v3: chunk_counter points at ONE COUNTER PER FRAME (all zeroed by the host) and a
	                          // CTA keeps claiming strips of the frame it is on until that frame is used up: the vectorscope
	                          // table is flushed when the frame changes, and with one counter for the whole batch that is
	                          // after nearly every chunk (0: one counter over all strips of the batch)
	uint32_t *hist;           // [n][1024] u32, zeroed
	uint8_t *wave;            // [n][256][out_width][4]
	uint32_t *wave_pairs;     // partial: [2][256][out_width]: plane 0 = (B|U : lo16, G|Y : hi16), plane 1 = R|V
	uint32_t *vscope_acc;     // [n][65536] u32, zeroed
	unsigned long long hist_stride, wave_stride, vscope_stride; // elements between frames
	uint8_t *wave_copies[kMaxWaveCopies]; // column-band sharding: every rank's image gets this rank's columns
	uint32_t scale_x, scale_y; // plain-load kernel: point-downsample (0 or 1: none): pixel (x, y) of the pass is the source
	                          // pixel (x * scale_x + scale_x / 2, y * scale_y + scale_y / 2); width / height are the
	                          // SCALED size (target size / target_scale, common.c:249-250).  The host entry points
	                          // drop the rows while copying (scale_y = 1 on the device then)
	uint32_t xform_strict;    // plain-load kernel, fused mode: evaluate the transform in fp32, every product and sum
	                          // rounded separately (SCOPE_XFORM_FP32_STRICT) instead of the exact integer form
	int32_t colorspace;       // 1 = BT.601, else BT.709 (only the strict transform looks at it; the exact one has `coef`)
	uint32_t v3c[8];          // scope_fused_kernel_v3: its float / integer constants as LAUNCH PARAMETERS (v3_param_consts):
	                          // ptxas re-materialised them as immediates with ~18 MOVs per visit (10 % of all executed
	                          // instructions, profiles/ncu_lines_r02e.md); from the constant bank they are free operands
	uint32_t rt_zero;         // always 0, but only known at run time: the consumers AND it with the pixels they loaded
	                          // and add it to the address of the "stage is free" arrive, so that ptxas must keep the
	                          // arrive behind the arrival of the data (a `mov 0` inside inline PTX is folded by ptxas)
	Coef coef;
};

// ---------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
#ifdef SCOPE_EMULATE
	return emul::kSmemBase + (uint32_t)(static_cast<const uint8_t *>(p) - emul::W->cur->cta->smem.data());
#else
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
#endif
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
#ifdef SCOPE_EMULATE
	return emul::mbar_init(bar, count);
#else
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
#endif
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
#ifdef SCOPE_EMULATE
	return emul::mbar_expect_tx(bar, bytes);
#else
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
#ifdef SCOPE_EMULATE
	return emul::mbar_arrive(bar);
#else
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
#ifdef SCOPE_EMULATE
	return emul::mbar_wait(bar, parity);
#else
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
#endif
}
// non-blocking: has the phase with this parity completed?  (acquire, like try_wait)
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity)
{
#ifdef SCOPE_EMULATE
	return emul::mbar_test(bar, parity) ? 1u : 0u;
#else
	uint32_t ok;
	asm volatile("{\n"
		     ".reg .pred p;\n"
		     "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		     "selp.u32 %0, 1, 0, p;\n"
		     "}"
		     : "=r"(ok)
		     : "r"(bar), "r"(parity)
		     : "memory");
	return ok;
#endif
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y, int z)
{
#ifdef SCOPE_EMULATE
	return emul::tma_issue(dst, map, bar, x, y, z);
#else
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
		     " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
		     "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z)
		     : "memory");
#endif
}
// L2 prefetch of a tile (no shared-memory destination, no completion): issued a few tiles ahead of the load so that
// the load itself finds its lines in L2 - the ring then covers the L2 latency instead of the HBM latency
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int x, int y, int z)
{
#ifdef SCOPE_EMULATE
	(void)map; (void)x; (void)y; (void)z;
#else
	asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z)
		     : "memory");
#endif
}
__device__ __forceinline__ uint32_t atom_shared_add(uint32_t addr, uint32_t val)
{
#ifdef SCOPE_EMULATE
	return atomicAdd(reinterpret_cast<uint32_t *>(emul::smem_ptr(addr)), val);
#else
	uint32_t old;
	asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(val));
	return old;
#endif
}
__device__ __forceinline__ uint32_t ld_nc_u32(const void *p)
{
#ifdef SCOPE_EMULATE
	return *static_cast<const uint32_t *>(p);
#else
	uint32_t v;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
#endif
}

// N consecutive 128-byte tile rows -> one pixel per lane per row, i.e. exactly what
// `p[k] = tile[k * 32 + lane]` reads, but with one LSU instruction per FOUR rows.
// ldmatrix (m8n8, b16) hands thread T the 32-bit word T % 4 of the 16-byte row whose address
// lane 8 j + T / 4 supplied, for j = 0..3.  With lane L supplying rows_addr + 16 L that is
// the word at rows_addr + 128 j + 4 T: row j, pixel column T.  512 bytes per instruction at the
// full 128 B/clk of the shared-memory crossbar (LDS.32 is issue-bound at half of that).
template <int N>
__device__ __forceinline__ void ldsm_rows(uint32_t rows_addr, int lane, uint32_t (&p)[N])
{
#ifdef SCOPE_EMULATE
	for (int k = 0; k + 3 < N; k += 4) {
		uint32_t r[4];
		emul::ldmatrix<4>(rows_addr + (uint32_t)k * 128u + (uint32_t)lane * 16u, r);
		for (int j = 0; j < 4; j++)
			p[k + j] = r[j];
	}
	if (N % 4 == 2) {
		uint32_t r[2];
		emul::ldmatrix<2>(rows_addr + (uint32_t)(N - 2) * 128u + (uint32_t)(lane & 15) * 16u, r);
		p[N - 2] = r[0];
		p[N - 1] = r[1];
	}
	return;
#else
	// (callers only come here with N even: groups of four rows, then one pair if N % 4 == 2)
#pragma unroll
	for (int k = 0; k + 3 < N; k += 4)
		asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
			     : "=r"(p[k]), "=r"(p[k + 1]), "=r"(p[k + 2]), "=r"(p[k + 3])
			     : "r"(rows_addr + (uint32_t)k * 128u + (uint32_t)lane * 16u)
			     : "memory");
	if (N % 4 == 2)
		asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];"
			     : "=r"(p[N - 2]), "=r"(p[N - 1])
			     : "r"(rows_addr + (uint32_t)(N - 2) * 128u + (uint32_t)(lane & 15) * 16u)
			     : "memory");
#endif
}

// ---------------------------------------------------------------------------
// "carriers": a byte b travels as the bit pattern 0x4B0000bb, i.e. the float 2^23 + b.
//   * PRMT makes one from a pixel byte in a single instruction;
//   * integer multiply-adds on carriers differ from those on the bytes by a constant that
//     folds into the addend (the colour transform below);
//   * (carrier << 7) + (base - 0x80000000) == base + 128*b   (0x4B000000 << 7 = 0x80000000
//     mod 2^32), so a waveform bin address is ONE multiply-add away from a carrier.
// ---------------------------------------------------------------------------
// With SCOPE_FADDR the bias is 0: a byte b travels as the plain integer b, which read as a float
// is the DENORMAL b * 2^-149.  FFMA without .ftz is exact on denormals, so
//   fma(b, 128.0, base) = (128 b + base) * 2^-149, whose bit pattern is the integer 128 b + base:
// the bin address comes out of the FMA pipe (full rate, shared with nothing else in this kernel)
// instead of the half-rate FMA-heavy (IMAD) or ALU (LEA) pipes that bound the inner loop.
constexpr uint32_t kCarrierBias = SCOPE_FADDR ? 0u : 0x4B000000u;

// Division by 10^6 on the FMA pipe (round 2: IMAD.HI occupies the FMA-heavy pipe like four IMADs, tools/ubench3.cu).
// S < 2^28 is exact; 10^6 = 64 * 15625, so floor(S / 10^6) = floor(T / 15625) with T = S >> 6 < 2^22.  The sums carry
// kDivExpBits = 0xC0000000 in their constant: a funnel shift by 6 with 0x12 in the upper word then yields the bit
// pattern 0x4B000000 | T, i.e. the float 2^23 + T, without a conversion.  q = round((T - 7812) / 15625) = floor(T / 15625)
// comes out of one add (exact) and one fused multiply-add whose addend 1.5 * 2^23 puts the sum where the ulp is 1: the
// result's bit pattern is 0x4B400000 + q.  The product's rounding error is < 1.7e-5, the distance of (T - 7812) / 15625
// from a half-integer >= 3.2e-5.  Checked for all 2^24 colours and both colour spaces against the oracle.
constexpr uint32_t kDivExpBits = 0xC0000000u;
constexpr uint32_t kDivFunnelHi = 0x12u;
// (hi : lo) >> n, low 32 bits
__device__ __forceinline__ uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t n)
{
#ifdef SCOPE_EMULATE
	return (uint32_t)((((uint64_t)hi << 32) | lo) >> n);
#else
	uint32_t d;
	asm("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(n));
	return d;
#endif
}

// 0x4B400000 + floor(S / 10^6) from a sum S that carries kDivExpBits; `round_to` = 1.5 * 2^23 (+ a bias the caller wants
// in the result)
__device__ __forceinline__ uint32_t div_1e6_bits(uint32_t s, float round_to)
{
#ifdef SCOPE_EMULATE
	uint32_t t = funnel_r(s, kDivFunnelHi, 6u);
	float tf;
	memcpy(&tf, &t, 4);
	const float d = std::fmaf(tf + -(8388608.0f + 7812.0f), 1.0f / 15625.0f, round_to);
	uint32_t r;
	memcpy(&r, &d, 4);
	return r;
#else
	const float tf = __uint_as_float(funnel_r(s, kDivFunnelHi, 6u));
	return __float_as_uint(__fmaf_rn(__fadd_rn(tf, -(8388608.0f + 7812.0f)), 1.0f / 15625.0f, round_to));
#endif
}

// (host side)
// 10^6 x the coefficients printed in data/common.effect:27-29 (BT.601) and :38-40 (BT.709), as
// integers, order R, G, B; K = floor(10^6 * (255 * off + 1/2)) with off = 1/2 - 1/256 (U), 0 (Y),
// 1/2 (V).  The kernel multiplies carriers (kCarrierBias + byte), so the bias they add is taken out
// of K here (mod 2^32; the bias is 0 in SCOPE_FADDR builds).
inline Coef coef_for(int colorspace)
{
	static const int32_t k601[3][3] = {{-147643, -289855, +437500}, {+299000, +587000, +114000},
					   {+437500, -366351, -71147}};
	static const int32_t k709[3][3] = {{-100643, -338571, +439216}, {+212600, +715200, +72200},
					   {+439216, -398941, -40273}};
	static const uint32_t k_add[3] = {127003906u + kDivExpBits, 500000u + kDivExpBits, 128000000u + kDivExpBits};
	const int32_t(*m)[3] = colorspace == 1 ? k601 : k709;
	Coef c;
	uint32_t *rows[3] = {c.u, c.y, c.v};
	uint32_t *adds[3] = {&c.ku, &c.ky, &c.kv};
	for (int ch = 0; ch < 3; ch++) {
		uint32_t sum = 0;
		for (int i = 0; i < 3; i++) {
			rows[ch][i] = (uint32_t)m[ch][i];
			sum += (uint32_t)m[ch][i];
		}
		*adds[ch] = k_add[ch] - sum * kCarrierBias;
	}
	return c;
}

// the same numbers as compile-time constants (SCOPE_IMMCOEF); CS = 1: BT.601, 2: BT.709
template <int CS>
__device__ __forceinline__ constexpr Coef const_coef()
{
	constexpr int32_t m[3][3] = {{CS == 1 ? -147643 : -100643, CS == 1 ? -289855 : -338571, CS == 1 ? 437500 : 439216},
				     {CS == 1 ? 299000 : 212600, CS == 1 ? 587000 : 715200, CS == 1 ? 114000 : 72200},
				     {CS == 1 ? 437500 : 439216, CS == 1 ? -366351 : -398941, CS == 1 ? -71147 : -40273}};
	constexpr uint32_t k_add[3] = {127003906u + kDivExpBits, 500000u + kDivExpBits, 128000000u + kDivExpBits};
	Coef c{};
	for (int i = 0; i < 3; i++) {
		c.u[i] = (uint32_t)m[0][i];
		c.y[i] = (uint32_t)m[1][i];
		c.v[i] = (uint32_t)m[2][i];
	}
	c.ku = k_add[0] - ((uint32_t)m[0][0] + (uint32_t)m[0][1] + (uint32_t)m[0][2]) * kCarrierBias;
	c.ky = k_add[1] - ((uint32_t)m[1][0] + (uint32_t)m[1][1] + (uint32_t)m[1][2]) * kCarrierBias;
	c.kv = k_add[2] - ((uint32_t)m[2][0] + (uint32_t)m[2][1] + (uint32_t)m[2][2]) * kCarrierBias;
	return c;
}


// a * b + c on the bit patterns of small non-negative integers (all < 2^23), b a float constant
__device__ __forceinline__ uint32_t fma_bits(uint32_t a, float b, uint32_t c)
{
#ifdef SCOPE_EMULATE
	return emul::fma_bits(a, b, c);
#else
	float d;
	asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(__uint_as_float(a)), "f"(b), "f"(__uint_as_float(c)));
	return __float_as_uint(d);
#endif
}

// constants the compiler must not see through: the carrier bias stays in a register so that PRMT can take
// its selector as the immediate; the zero is used to build data dependencies
__device__ __forceinline__ uint32_t opaque_carrier_bias()
{
#ifdef SCOPE_EMULATE
	return kCarrierBias;
#else
	uint32_t v;
	asm volatile("mov.u32 %0, %1;" : "=r"(v) : "n"(kCarrierBias));
	return v;
#endif
}
__device__ __forceinline__ uint32_t opaque_zero()
{
#ifdef SCOPE_EMULATE
	return 0u;
#else
	uint32_t v;
	asm volatile("mov.u32 %0, 0;" : "=r"(v));
	return v;
#endif
}

template <int K>
__device__ __forceinline__ uint32_t carrier(uint32_t pixel, uint32_t magic)
{
#ifdef SCOPE_EMULATE
	return emul::prmt(pixel, magic, 0x7650u + (uint32_t)K);
#else
	uint32_t r;
	if (K == 0)
		asm("prmt.b32 %0, %1, %2, 0x7650;" : "=r"(r) : "r"(pixel), "r"(magic));
	else if (K == 1)
		asm("prmt.b32 %0, %1, %2, 0x7651;" : "=r"(r) : "r"(pixel), "r"(magic));
	else
		asm("prmt.b32 %0, %1, %2, 0x7652;" : "=r"(r) : "r"(pixel), "r"(magic));
	return r;
#endif
}

// RGB -> YUV (data/common.effect:23-43), pinned as the EXACT value of the shader's expression
// rounded the way a UNORM8 render target rounds (oracle/scope_oracle.c):
//   q = floor(c0*R + c1*G + c2*B + 255*off + 1/2)       (bytes in, real arithmetic)
// The effect file's coefficients have six decimals, so with c' = 10^6 c the sum
//   S = c0'*R + c1'*G + c2'*B + K'    (K' = floor(10^6 (255 off + 1/2)), 0 <= S < 2^28)
// is an exact integer and q = floor(S / 10^6) = (S * ceil(2^48 / 10^6)) >> 48 for every S < 2^28
// (the usual multiply-high division; checked for all 2^24 colours in tests/test_oracle.py).
// The high word of that product is q * 2^16 + fraction: q sits in BYTE 2, where one PRMT picks it
// up.  The three multiply-adds run on the CARRIERS (0x4B000000 + byte): the extra
// 0x4B000000 * (c0'+c1'+c2') is folded into K' on the host, mod 2^32.
constexpr uint32_t kDivMagic = 281474977u; // ceil(2^48 / 10^6)

__device__ __forceinline__ uint32_t yuv_channel(const uint32_t (&bgr)[3], const uint32_t (&c)[3], uint32_t k)
{
	uint32_t s = bgr[2] * c[0] + k;
	s = bgr[1] * c[1] + s;
	s = bgr[0] * c[2] + s;
	// q in BYTE 2 (bytes 0, 1, 3 zero): (0x4B400000 + q) << 16 = q << 16 mod 2^32
	return div_1e6_bits(s, 12582912.0f) << 16;
}

// B, G, R carriers of one pixel -> U, (Y), V, each in byte 2 of its word (bytes 3 = 0)
// `need`: bit 0 U, bit 1 Y, bit 2 V (launch-uniform: a luma-only waveform - BASELINE config 4 - evaluates one channel, not three)
template <bool NEED_Y>
__device__ __forceinline__ void rgb_to_yuv_hi(const uint32_t (&bgr)[3], const Coef &c, uint32_t (&hi)[3], uint32_t need = 7u)
{
	hi[0] = (need & 1u) ? yuv_channel(bgr, c.u, c.ku) : 0u;
	hi[1] = (NEED_Y && (need & 2u)) ? yuv_channel(bgr, c.y, c.ky) : 0u;
	hi[2] = (need & 4u) ? yuv_channel(bgr, c.v, c.kv) : 0u;
}

// the same, as carriers of U, Y, V
template <bool NEED_Y>
__device__ __forceinline__ void rgb_to_yuv_carriers(const uint32_t (&bgr)[3], const Coef &c, uint32_t magic,
						     uint32_t (&yuv)[3], uint32_t need = 7u)
{
	uint32_t hi[3];
	rgb_to_yuv_hi<NEED_Y>(bgr, c, hi, need);
	yuv[0] = carrier<2>(hi[0], magic);
	yuv[1] = NEED_Y ? carrier<2>(hi[1], magic) : 0u;
	yuv[2] = carrier<2>(hi[2], magic);
}

// SCOPE_XFORM_FP32_STRICT: the fp32 reading of data/common.effect:23-43 that SURVEY.md 8(c) drafted and
// oracle/scope_oracle.c keeps as variant 1 ("strict"): xf = (float)X / 255.0f, every product and sum rounded
// separately, left to right, no FMA; q = floor(min(max(t, 0), 1) * 255 + 0.5) with the product and the sum
// rounded separately as well.  It differs from the exact value on 0-387 of the 2^24 colours per channel, by one
// step.  pixel = the BGRA word; result bytes U, Y, V as plain integers.
#ifndef SCOPE_EMULATE
__device__ __forceinline__ uint32_t strict_channel(float r, float g, float b, float c0, float c1, float c2, float off)
{
	float s = __fadd_rn(__fmul_rn(c0, r), __fmul_rn(c1, g));
	s = __fadd_rn(s, __fmul_rn(c2, b));
	s = __fadd_rn(s, off);
	s = fminf(fmaxf(s, 0.0f), 1.0f);
	return (uint32_t)floorf(__fadd_rn(__fmul_rn(s, 255.0f), 0.5f));
}
__device__ __forceinline__ void rgb_to_yuv_strict(uint32_t pixel, int colorspace, uint32_t (&uyv)[3])
{
	const float b = __fdiv_rn((float)(pixel & 0xFFu), 255.0f), g = __fdiv_rn((float)((pixel >> 8) & 0xFFu), 255.0f),
		    r = __fdiv_rn((float)((pixel >> 16) & 0xFFu), 255.0f);
	const float off_u = 0.5f - 1.0f / 256.0f;
	if (colorspace == 1) {
		uyv[0] = strict_channel(r, g, b, -0.147643f, -0.289855f, +0.437500f, off_u);
		uyv[1] = strict_channel(r, g, b, +0.299000f, +0.587000f, +0.114000f, 0.0f);
		uyv[2] = strict_channel(r, g, b, +0.437500f, -0.366351f, -0.071147f, 0.5f);
	} else {
		uyv[0] = strict_channel(r, g, b, -0.100643f, -0.338571f, +0.439216f, off_u);
		uyv[1] = strict_channel(r, g, b, +0.212600f, +0.715200f, +0.072200f, 0.0f);
		uyv[2] = strict_channel(r, g, b, +0.439216f, -0.398941f, -0.040273f, 0.5f);
	}
}
#else
__device__ __forceinline__ void rgb_to_yuv_strict(uint32_t, int, uint32_t (&uyv)[3]) // (the emulator runs the exact form only)
{
	uyv[0] = uyv[1] = uyv[2] = 0;
}
#endif

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
	// consumer warps and tile height of this kernel family: the one-plane kernels with the vectorscope (one CTA per
	// SM) run SCOPE_FUSED_WARPS warps; the two-CTA kernels and the two-plane surface-mode ring (no room for taller
	// tiles) keep kTmaWarps.  Every warp takes kRowsPerWarp rows of a tile.
	static constexpr bool kOnePlaneVs = VSCOPE && kPlanes == 1;
	static constexpr int kWarps = (USE_TMA && kOnePlaneVs && !SCOPE_WIDE_FUSED) ? SCOPE_FUSED_WARPS : kTmaWarps;
	static constexpr int kRowsPerWarp = (scope::kTileRows / kTmaWarps) * ((SCOPE_WIDE_FUSED && kOnePlaneVs) ? 2 : 1);
	static constexpr int kTileRows = kWarps * kRowsPerWarp;
	static constexpr int kTileBytes = kStripPx * 4 * kTileRows;
	static constexpr int kStageBytes = USE_TMA ? kPlanes * kTileBytes : 0;
	// two planes per stage (surface mode) leave room for a 2-deep ring only
	// as many stages as fit (two planes per stage in surface mode halve the depth)
#if SCOPE_DEEP_RING
	// what the bins leave of the SM's shared memory: 227 KB per CTA alone on an SM, half of 228 KB minus
	// the 1 KB the system reserves per CTA when two CTAs share it; 512 B for barriers and the mailbox
	static constexpr int kRing =
		((VSCOPE || !(SRC == SRC_RGB || SURFACE)) ? 227 * 1024 : 113 * 1024) - kStageOff - 512;
	static constexpr int kStagesFit = USE_TMA ? kRing / (kPlanes * kTileBytes) : 1;
#else
	static constexpr int kStagesFit = USE_TMA ? kRingBytes / (kPlanes * kTileBytes) : 1;
#endif
	static constexpr int kStages = kStagesFit > kMaxStages ? kMaxStages : (kStagesFit < 2 ? 2 : kStagesFit);
#ifndef SCOPE_EXPERIMENT
	static_assert(!USE_TMA || kStagesFit >= 2, "tile too large for the ring");
#endif
	static_assert(!USE_TMA || kStages <= kQueue || kStages <= 4, "chunk mailbox shorter than the ring");
	static constexpr int kBarOff = kStageOff + kStages * kStageBytes;
	static constexpr int kQueueOff = kBarOff + (USE_TMA ? 2 * kMaxStages * 8 : 0);
	static constexpr int kTotal = kQueueOff + (USE_TMA ? kQueue * 8 : 0) + 16;
};

// ---------------------------------------------------------------------------
// per-pixel accumulation primitives
// ---------------------------------------------------------------------------
__device__ __forceinline__ void red_shared(uint32_t addr, uint32_t val)
{
#ifdef SCOPE_EMULATE
	atomicAdd(reinterpret_cast<uint32_t *>(emul::smem_ptr(addr)), val);
	return;
#else
	asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(val));
#endif
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
#if SCOPE_FADDR
	if (DO_B)
		red_shared(fma_bits(cb, 128.0f, wb0), one);
	if (DO_G)
		red_shared(fma_bits(cg, 128.0f, wb0), one << 16);
	if (DO_R)
		red_shared(fma_bits(cr, 128.0f, wb1), one);
#else
	if (DO_B)
		red_shared(cb * 128u + wb0, one);
	if (DO_G)
		red_shared(cg * 128u + wb0, one << 16);
	if (DO_R)
		red_shared(cr * 128u + wb1, one);
#endif
}

// Vectorscope bins in shared memory are indexed by idx = U | V << 8 (NOT yet flipped to the
// reference's row = 255 - V; the flush does that) and are u16 halves, two per 32-bit word:
// word = idx & 0x7FFF, half = V >> 7.  vs_add returns the bit of the OLD word that says "this
// half already held >= 0x8000"; the caller ORs those over its pixels and, only if any is set,
// calls vs_undo.  An add that found its bin at >= 0x8000 is taken back, so a half can never
// wrap 16 bits, and a bin that ever reached 0x8000 keeps a value far above 255: it saturates
// to 255 at the end exactly like the reference's `if (*c < 255) ++*c`.
//
// Bank swizzle.  The plain word index U + 256 (V & 127) puts every bin of one U column in the
// same bank (bank = U mod 32), and picture content has few distinct U values per 32-pixel row
// segment: simulated on the "natural" test frames one warp-wide atomic then costs 11.2
// conflict passes (tools/vs_conflicts.py).  Mixing the low bits of V into the bank spreads
// neighbouring (U, V) pairs over the banks (4.9 passes, the floor set by lanes that hit the very
// SAME bin); random content is unchanged (3.5).  Two forms, both bijections on the 15-bit word
// index that flush_vscope undoes: XOR (shipped) and the original (U + 4 V) mod 32.
__device__ __forceinline__ uint32_t vs_word(uint32_t idx)
{
#if SCOPE_XORSWZ
	// bits 2..4 of U are XOR-ed with the low three bits of V: same spreading (simulated passes
	// per warp-wide atomic: tools/vs_conflicts.py), two instructions instead of four, and still
	// a permutation of 4-word groups for fixed V (what flush_vscope relies on)
	return (idx ^ ((idx >> 6) & 0x1Cu)) & 0x7FFFu;
#else
	const uint32_t t = (idx >> 8) * 4u + idx; // low 5 bits: (U + 4 V) mod 32
	return (idx & 0x7FE0u) | (t & 31u);
#endif
}
// inverse of vs_word for a 4-aligned word index: the U of the group's first word (V = word >> 8)
__device__ __forceinline__ uint32_t vs_unswizzle_u(uint32_t word)
{
	const uint32_t v7 = word >> 8;
#if SCOPE_XORSWZ
	return (word & 0xFFu) ^ ((v7 & 7u) << 2);
#else
	return (word & 0xE0u) | ((word - v7 * 4u) & 31u);
#endif
}

// byte address of a bin's word, and the addend that puts k into the bin's half (h = idx >> 15)
__device__ __forceinline__ uint32_t vs_addr(uint32_t vs_base, uint32_t idx)
{
#if SCOPE_FADDR
	return fma_bits(vs_word(idx), 4.0f, vs_base);
#else
	return vs_word(idx) * 4u + vs_base;
#endif
}
__device__ __forceinline__ uint32_t vs_one(uint32_t idx)
{
#if SCOPE_FADDR
	return fma_bits(idx >> 15, 65535.0f, 1u);
#else
	return (idx >> 15) * 0xFFFFu + 1u;
#endif
}

struct VsAdd {
	uint32_t addr, add, sat;
};
__device__ __forceinline__ VsAdd vs_add(uint32_t vs_base, uint32_t idx, uint32_t k)
{
	VsAdd r;
	const uint32_t h = idx >> 15;            // 0: lower half (V < 128), 1: upper half
	r.addr = vs_addr(vs_base, idx);
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
	uint32_t strict = 0; // SCOPE_XFORM_FP32_STRICT (plain-load kernel only)
	int colorspace = 2;
};

template <int SRC, bool VSCOPE, bool SURFACE, bool FAST, int N>
__device__ __forceinline__ void process_tile(const TileCtx &c, const Coef &coef, const uint32_t (&p)[N],
					     const uint32_t (&q)[N], const bool (&ok)[N])
{
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
		if (c.strict) {
#pragma unroll
			for (int k = 0; k < N; k++) {
				uint32_t uyv[3];
				rgb_to_yuv_strict(p[k], c.colorspace, uyv);
#pragma unroll
				for (int j = 0; j < 3; j++)
					cyuv[k][j] = uyv[j] + kCarrierBias;
			}
		} else {
			const uint32_t need = VSCOPE ? (c.bins_mask | 5u) : c.bins_mask;
#pragma unroll
			for (int k = 0; k < N; k++)
				rgb_to_yuv_carriers<SRC == SRC_YUV>(crgb[k], coef, c.magic, cyuv[k], need);
		}
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
#ifdef SCOPE_EMULATE
	return emul::bar_sync(1, NW * 32);
#else
	asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
#endif
}

// SCOPE_DEPHASE: all NW consumer warps meet here once per strip, the late half (warp & 4) BEFORE their first
// tile's arithmetic, the others AFTER theirs
template <int NW>
__device__ __forceinline__ void dephase_bar()
{
#ifdef SCOPE_EMULATE
	return emul::bar_sync(2, NW * 32);
#else
	asm volatile("bar.sync 2, %0;" ::"n"(NW * 32) : "memory");
#endif
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
	constexpr int kStep = NW * 32, kBatch = 4; // kBatch loads in flight per thread
	for (int i0 = tid; i0 < kVsWords / 4; i0 += kBatch * kStep) {
		uint4 wv[kBatch];
#pragma unroll
		for (int b = 0; b < kBatch; b++) {
			const int i = i0 + b * kStep;
			wv[b] = i < kVsWords / 4 ? reinterpret_cast<uint4 *>(vs)[i] : make_uint4(0, 0, 0, 0);
		}
#pragma unroll
		for (int b = 0; b < kBatch; b++) {
			const int i = i0 + b * kStep;
			const uint4 w = wv[b];
			if ((w.x | w.y | w.z | w.w) != 0u) {
				const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
				// word = vs_word(U | V << 8): undo the bank swizzle (the four words of this
				// uint4 stay four consecutive U values: the swizzle permutes groups of 4)
				const uint32_t word = i * 4;
				const uint32_t v7 = word >> 8, u = vs_unswizzle_u(word);
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
	if (SCOPE_FAST_EMIT && NW >= 8 && !P.partial && !P.n_wave_copies && P.wave_mask == 7u &&
	    (P.hist_mask == 7u || P.hist_mask == 0u)) {
		// The common case (all three channels to the waveform, all or none to the histogram),
		// without per-level tests of launch-uniform flags.  Levels whose 32 columns are all empty
		// (most levels, for picture content) only get their zero row written.  The warp's column
		// sums stay in registers - lane i keeps those of the warp's i-th level - and go out as one
		// atomic per channel per warp at the end instead of three predicated ones per level.
		uint32_t *dst = reinterpret_cast<uint32_t *>(P.wave + (size_t)frame * P.wave_stride) + xo;
		const bool do_hist = P.hist_mask != 0u;
		uint32_t keep_b = 0, keep_g = 0, keep_r = 0;
		int i = 0;
#pragma unroll 4
		for (int v = warp; v < 256; v += NW, i++) {
			const uint32_t w0 = wave0[v * 32 + lane];
			const uint32_t w1 = wave0[kWaveWords + v * 32 + lane];
			uint32_t *row = dst + (size_t)(255 - v) * P.out_width;
			if (!__any_sync(0xFFFFFFFFu, (w0 | w1) != 0u)) {
				if (lane_ok)
					*row = 0u;
				continue;
			}
			wave0[v * 32 + lane] = 0;
			wave0[kWaveWords + v * 32 + lane] = 0;
			const uint32_t cb = w0 & 0xFFFFu, cg = w0 >> 16, cr = w1 & 0xFFFFu;
			if (do_hist) {
				const uint32_t sb = __reduce_add_sync(0xFFFFFFFFu, cb);
				const uint32_t sg = __reduce_add_sync(0xFFFFFFFFu, cg);
				const uint32_t sr = __reduce_add_sync(0xFFFFFFFFu, cr);
				if (lane == i) {
					keep_b = sb;
					keep_g = sg;
					keep_r = sr;
				}
			}
			if (lane_ok)
				*row = min(cb, 255u) | (min(cg, 255u) << 8) | (min(cr, 255u) << 16);
		}
		const int v = warp + lane * NW; // the level whose sums this lane kept
		if (do_hist && v < 256) {
			if (keep_r)
				atomicAdd(hist + v * 4 + 0, keep_r);
			if (keep_g)
				atomicAdd(hist + v * 4 + 1, keep_g);
			if (keep_b)
				atomicAdd(hist + v * 4 + 2, keep_b);
		}
		workers_bar<NW>();
		return;
	}
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
			if (P.partial == 2u) {
				// the only writer of these columns (a row band accumulated in one launch): plain stores
				P.wave_pairs[o] = mb | (mg << 16);
				if (P.wave_mask & 4u) // plane 1 only exists for the R|V channel
					P.wave_pairs[(size_t)256 * P.out_width + o] = mr;
			} else if (P.partial) {
				if (mb | mg)
					atomicAdd(P.wave_pairs + o, mb | (mg << 16));
				if (mr)
					atomicAdd(P.wave_pairs + (size_t)256 * P.out_width + o, mr);
			} else {
				const uint32_t word = min(mb, 255u) | (min(mg, 255u) << 8) | (min(mr, 255u) << 16);
				uint32_t *dst = reinterpret_cast<uint32_t *>(P.wave + (size_t)frame * P.wave_stride);
				dst[o] = word;
				// column bands over several GPUs: the same columns of every other rank's image
				// (NVLink peer stores, 128 bytes per warp and row: the all-gather happens here)
				for (uint32_t c = 0; c < P.n_wave_copies; c++)
					reinterpret_cast<uint32_t *>(P.wave_copies[c])[o] = word;
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
#if SCOPE_RAWFLAT
	bool rawflat;      // warp-uniform: all N x 32 pixel words are equal; only element [0] is filled in
#endif
#if SCOPE_BALLOT
	bool dominant;     // warp-uniform: not flat, but many lanes hold only the reference bin `ref`
	uint32_t ref;      // warp-uniform: lane 0's first vectorscope bin
#endif
};

template <int SRC, bool VSCOPE, bool SURFACE, int N>
__device__ __forceinline__ void prepare_tile(const TileCtx &c, const Coef &coef, const uint32_t (&p)[N],
					     const uint32_t (&q)[N], Prep<N> &o)
{
	constexpr bool kTransform = !SURFACE && (VSCOPE || SRC == SRC_YUV);
#if SCOPE_RAWFLAT
	o.rawflat = false;
	if (!SURFACE) { // everything below is a function of the pixel word alone
		uint32_t d = 0;
#pragma unroll
		for (int k = 1; k < N; k++)
			d |= p[0] ^ p[k];
		const uint32_t p_lane0 = __shfl_sync(0xFFFFFFFFu, p[0], 0);
		if (__all_sync(0xFFFFFFFFu, (d | (p[0] ^ p_lane0)) == 0u)) {
			// sass-cold{
			const uint32_t bgr[3] = {carrier<0>(p[0], c.magic), carrier<1>(p[0], c.magic), carrier<2>(p[0], c.magic)};
			uint32_t h3[3] = {0u, 0u, 0u};
			if (kTransform)
				rgb_to_yuv_hi<SRC == SRC_YUV>(bgr, coef, h3);
#pragma unroll
			for (int j = 0; j < 3; j++)
				o.cs[0][j] = SRC == SRC_RGB ? bgr[j] : carrier<2>(h3[j], c.magic);
			o.a[0] = p[0];
			// the fused YUV plane has alpha 255 everywhere (common.effect:30,41): only an RGB source looks at it
			o.all_counted = SRC != SRC_RGB || p[0] > 0x00FFFFFFu;
			o.idx[0] = __byte_perm(h3[0], h3[2], 0x3362);
			o.flat = true;
			o.rawflat = true;
			return;
			// sass-cold}
		}
	}
#endif
	uint32_t crgb[N][3]; // carriers of B, G, R
	uint32_t hi[N][3];   // transform results: U, Y, V in byte 2
	if (SRC == SRC_RGB || kTransform) {
#pragma unroll
		for (int k = 0; k < N; k++) {
			crgb[k][0] = carrier<0>(p[k], c.magic);
			crgb[k][1] = carrier<1>(p[k], c.magic);
			crgb[k][2] = carrier<2>(p[k], c.magic);
		}
	}
	if (kTransform) {
		const uint32_t need = VSCOPE ? (c.bins_mask | 5u) : c.bins_mask;
#pragma unroll
		for (int k = 0; k < N; k++)
			rgb_to_yuv_hi<SRC == SRC_YUV>(crgb[k], coef, hi[k], need);
	}
	o.all_counted = true;
	if (SRC != SRC_NONE) {
#pragma unroll
		for (int k = 0; k < N; k++) {
#pragma unroll
			for (int j = 0; j < 3; j++) {
				if (SRC == SRC_RGB)
					o.cs[k][j] = crgb[k][j];
				else if (SURFACE)
					o.cs[k][j] = j == 0   ? carrier<0>(q[k], c.magic)
						     : j == 1 ? carrier<1>(q[k], c.magic)
							      : carrier<2>(q[k], c.magic);
				else
					o.cs[k][j] = carrier<2>(hi[k][j], c.magic);
			}
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
			o.idx[k] = SURFACE ? __byte_perm(q[k], 0u, 0x4420) : __byte_perm(hi[k][0], hi[k][2], 0x3362);
		bool same = true;
#pragma unroll
		for (int k = 1; k < N; k++)
			same = same && (o.idx[k] == o.idx[0]);
		const uint32_t idx_lane0 = __shfl_sync(0xFFFFFFFFu, o.idx[0], 0);
#if SCOPE_BALLOT
		const uint32_t agree = __ballot_sync(0xFFFFFFFFu, same && (o.idx[0] == idx_lane0));
		o.flat = agree == 0xFFFFFFFFu;
		o.dominant = !o.flat && __popc(agree) >= kBallotLanes;
		o.ref = idx_lane0;
#else
		o.flat = __all_sync(0xFFFFFFFFu, same && (o.idx[0] == idx_lane0));
#endif
	}
#if SCOPE_BALLOT
	else {
		o.dominant = false;
		o.ref = 0;
	}
#endif
}

// commit_issue: every shared-memory atomic of the tile.  The vectorscope adds return the old
// words into `pend`; commit_resolve looks at them LATER (after the next tile's arithmetic), so
// the warp never sits waiting for the LSU queue to hand the old values back.  Deferring the
// check does not change the saturation argument of vs_add: a thread still has at most N
// unchecked adds in flight, because it resolves one tile before issuing the next.
template <int SRC, bool VSCOPE, bool SURFACE, int N>
__device__ __forceinline__ void commit_issue(const TileCtx &c, const Prep<N> &o, uint32_t (&pend)[N])
{
#if SCOPE_RAWFLAT
	if (o.rawflat) {
		// sass-cold{
		// N equal pixels per lane: each column bin gets +N once; a fully transparent block (RGB
		// source, alpha 0) adds nothing to the bins but still counts for the vectorscope
		if (SRC != SRC_NONE) {
			const uint32_t n = o.all_counted ? (uint32_t)N : 0u;
			if (c.bins_mask & 1u)
				bins_add<true, false, false>(o.cs[0][0], o.cs[0][1], o.cs[0][2], c.wb0, c.wb1, n);
			if (c.bins_mask & 2u)
				bins_add<false, true, false>(o.cs[0][0], o.cs[0][1], o.cs[0][2], c.wb0, c.wb1, n);
			if (c.bins_mask & 4u)
				bins_add<false, false, true>(o.cs[0][0], o.cs[0][1], o.cs[0][2], c.wb0, c.wb1, n);
		}
		if (VSCOPE && c.lane == 0)
			vs_undo(vs_add(c.vs_base, o.idx[0], 32u * N));
		return;
		// sass-cold}
	}
#endif
	if (SRC != SRC_NONE) {
		if (o.all_counted && c.bins_mask == 7u) {
#pragma unroll
			for (int k = 0; k < N; k++)
				bins_add<true, true, true>(o.cs[k][0], o.cs[k][1], o.cs[k][2], c.wb0, c.wb1, 1u);
		} else {
			// some pixels transparent, or only some channels wanted (uniform branches)
			// sass-cold{ (tools/sass_budget.py leaves these lines out of the fast-path count)
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
			// sass-cold}
		}
	}
	if (VSCOPE) {
		if (o.flat) {
			// sass-cold{
			if (c.lane == 0)
				vs_undo(vs_add(c.vs_base, o.idx[0], 32u * N));
			// sass-cold}
#if SCOPE_BALLOT
		} else if (o.dominant) {
			// sass-cold{
			// every pixel of the block that hits the reference bin is counted by ballot and added once;
			// the others are added one by one, checked right away (nothing is left for commit_resolve)
			uint32_t n_ref = 0;
#pragma unroll
			for (int k = 0; k < N; k++) {
				const bool hit = o.idx[k] == o.ref;
				n_ref += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, hit));
				if (!hit)
					vs_undo(vs_add(c.vs_base, o.idx[k], 1u));
				pend[k] = 0u;
			}
			if (c.lane == 0)
				vs_undo(vs_add(c.vs_base, o.ref, n_ref));
			// sass-cold}
#endif
		} else {
#pragma unroll
			for (int k = 0; k < N; k++) {
				pend[k] = atom_shared_add(vs_addr(c.vs_base, o.idx[k]), vs_one(o.idx[k]));
			}
		}
	}
}

// the atomics of an ordinary tile and nothing else (no branch): the caller has checked all_counted,
// !flat (and !rawflat) and bins_mask == 7
template <int SRC, bool VSCOPE, bool SURFACE, int N>
__device__ __forceinline__ void commit_issue_fast(const TileCtx &c, const Prep<N> &o, uint32_t (&pend)[N])
{
#pragma unroll
	for (int k = 0; k < N; k++) {
		if (SRC != SRC_NONE)
			bins_add<true, true, true>(o.cs[k][0], o.cs[k][1], o.cs[k][2], c.wb0, c.wb1, 1u);
		if (VSCOPE)
			pend[k] = atom_shared_add(vs_addr(c.vs_base, o.idx[k]), vs_one(o.idx[k]));
	}
}

template <int SRC, bool VSCOPE, bool SURFACE, int N>
__device__ __forceinline__ void commit_resolve(const TileCtx &c, const Prep<N> &o, const uint32_t (&pend)[N])
{
	if (VSCOPE && !o.flat) {
		// the old words are OR-ed and tested once for "some half already >= 0x8000" (either
		// half: a false alarm only costs the exact re-check)
		uint32_t any = 0;
#pragma unroll
		for (int k = 0; k < N; k++)
			any |= pend[k];
		if (any & 0x80008000u) {
			// sass-cold{
#pragma unroll
			for (int k = 0; k < N; k++) {
				const uint32_t h = o.idx[k] >> 15;
				if (pend[k] & (h * 0x7FFF8000u + 0x8000u))
					red_shared(vs_addr(c.vs_base, o.idx[k]), 0u - vs_one(o.idx[k]));
			}
			// sass-cold}
		}
	}
}

template <int SRC, bool VSCOPE, bool SURFACE, int N>
__device__ __forceinline__ void commit_tile(const TileCtx &c, const Prep<N> &o)
{
	uint32_t pend[N];
	commit_issue<SRC, VSCOPE, SURFACE, N>(c, o, pend);
	commit_resolve<SRC, VSCOPE, SURFACE, N>(c, o, pend);
}

// ---------------------------------------------------------------------------
// TMA loader, shared by the two TMA kernels below.
// A producer warp (one elected lane) walks CHUNKS of up to P.chunk_items consecutive strips,
// claimed from a global counter so that fast and slow frame content balances across CTAs, and fills a
// kStages-deep shared-memory ring of 64-row x 128-byte tiles with cp.async.bulk.tensor.  The
// chunk id travels to the consumers through a small shared-memory queue that is written before
// the chunk's first tile is armed (mbarrier release/acquire orders it).
// ---------------------------------------------------------------------------
template <class L>
__device__ __forceinline__ void tma_produce(const StripParams &P, const CUtensorMap *map_rgb,
					    const CUtensorMap *map_yuv, uint32_t smem_base,
					    volatile uint32_t *chunk_q, uint32_t bar_full, uint32_t bar_empty)
{
	constexpr int kStages = L::kStages;
	const uint32_t tiles = (P.height + L::kTileRows - 1) / L::kTileRows;
	uint32_t stage = 0, phase = 0, qw = 0;
	for (;;) {
		// guided self-scheduling: take 1/(2 x grid) of what is left (at most chunk_items
		// strips, at least one), so chunks shrink towards the end of the batch and the CTAs
		// finish together.  `seen` may be stale; only the size of the claim depends on it.
		const uint32_t seen = *reinterpret_cast<volatile const uint32_t *>(P.chunk_counter);
		uint32_t want = seen < P.items ? (P.items - seen) / (2u * gridDim.x) : 1u;
		want = min(max(want, 1u), P.chunk_items);
		const uint32_t first = atomicAdd(P.chunk_counter, want);
		const bool done = first >= P.items;
		const uint32_t last = min(first + want, P.items);
		// announce the chunk (or the end) before its first tile can complete
		mbar_wait(bar_empty + 8 * stage, phase ^ 1);
		chunk_q[2 * (qw % kQueue)] = first;
		chunk_q[2 * (qw % kQueue) + 1] = done ? 0u : last - first;
		qw++;
		if (done) {
			mbar_arrive(bar_full + 8 * stage); // wake the consumers with no data
			break;
		}
		// L2 prefetch cursor (off by default, see kL2Ahead): runs kL2Ahead tiles ahead of the loads inside this chunk
		uint32_t pf_item = first, pf_t = 0;
		auto prefetch_to = [&](uint32_t item, uint32_t t) {
			const uint32_t goal = (item - first) * tiles + t + (uint32_t)kL2Ahead;
			while (pf_item < last && (pf_item - first) * tiles + pf_t <= goal) {
				const uint32_t f = pf_item / P.strips, st = pf_item - f * P.strips;
				if (L::kLoadRgb)
					tma_prefetch_3d(map_rgb, (int)(st * kStripPx + P.tma_x0_rgb), (int)(pf_t * L::kTileRows), (int)f);
				if (L::kLoadYuv)
					tma_prefetch_3d(map_yuv, (int)(st * kStripPx + P.tma_x0_yuv), (int)(pf_t * L::kTileRows), (int)f);
				if (++pf_t == tiles) {
					pf_t = 0;
					pf_item++;
				}
			}
		};
		for (uint32_t item = first; item < last; item++) {
			const uint32_t frame = item / P.strips, strip = item - frame * P.strips;
			const int x = (int)(strip * kStripPx);
			for (uint32_t t = 0; t < tiles; t++) {
				if (kL2Ahead > 0)
					prefetch_to(item, t);
				if (!(item == first && t == 0))
					mbar_wait(bar_empty + 8 * stage, phase ^ 1);
				const uint32_t dst = smem_base + L::kStageOff + stage * L::kStageBytes;
				mbar_expect_tx(bar_full + 8 * stage, L::kStageBytes);
				if (L::kLoadRgb)
					tma_load_3d(dst, map_rgb, bar_full + 8 * stage, x + (int)P.tma_x0_rgb, (int)(t * L::kTileRows),
						    (int)frame);
				if (L::kLoadYuv)
					tma_load_3d(dst + (L::kLoadRgb ? L::kTileBytes : 0), map_yuv, bar_full + 8 * stage,
						    x + (int)P.tma_x0_yuv, (int)(t * L::kTileRows), (int)frame);
				if (++stage == kStages) {
					stage = 0;
					phase ^= 1;
				}
			}
		}
	}
}

// One consumer warp's walk over the chunks.  R_SRC / R_VS = what THIS warp accumulates (its
// role), N = rows of every tile it takes starting at `row0`; K_BINS / K_VS = what the kernel as
// a whole holds in shared memory (all NWORK consumer warps meet in emit_strip / flush_vscope).
template <class L, int R_SRC, bool R_VS, bool SURFACE, int N, int NWORK, bool K_BINS, bool K_VS, int CS = 0>
__device__ __forceinline__ void tma_consume(const StripParams &P, uint8_t *smem, uint32_t smem_base,
					    volatile uint32_t *chunk_q, uint32_t bar_full, uint32_t bar_empty,
					    int row0, int warp, int lane, int tid)
{
	constexpr int kStages = L::kStages;
	constexpr bool kNeedP = R_SRC == SRC_RGB || (!SURFACE && (R_VS || R_SRC == SRC_YUV));
	constexpr bool kNeedQ = SURFACE && (R_SRC == SRC_YUV || R_VS);
	uint32_t *vs = reinterpret_cast<uint32_t *>(smem + L::kVsOff);
	uint32_t *wave0 = reinterpret_cast<uint32_t *>(smem + L::kWave0Off);
	const uint32_t tiles = (P.height + L::kTileRows - 1) / L::kTileRows;
	const uint32_t wave_lane_addr = smem_base + L::kWave0Off + lane * 4;
	uint32_t magic; // 0x4B000000 kept in a register so PRMT can take the selector as its immediate
	magic = opaque_carrier_bias();
	const TileCtx tc{smem_base + L::kVsOff, wave_lane_addr - kCarrierBias * 128u,
			 wave_lane_addr - kCarrierBias * 128u + kWaveWords * 4, magic, P.bins_mask, lane};
	// CS != 0 (SCOPE_IMMCOEF): the nine multipliers are compile-time constants; the three addends stay in
	// registers (an IMAD takes one immediate, and with both constant ptxas spends an extra move per pixel)
	Coef coef = CS ? const_coef<CS ? CS : 2>() : P.coef;
	coef.ku = P.coef.ku;
	coef.ky = P.coef.ky;
	coef.kv = P.coef.kv;
	const uint32_t zero = P.rt_zero; // a 0 the compiler cannot see through (used to build data dependencies)
	uint32_t stage = 0, phase = 0, qr = 0;
	uint32_t cur_frame = 0xFFFFFFFFu;

	for (;;) {
		// the chunk id becomes readable once the chunk's first tile (or the end marker) lands
		mbar_wait(bar_full + 8 * stage, phase);
		// (lane 0 reads the mailbox for the warp: it is also the lane whose arrive hands stages back, so the
		// read is ordered before the producer's next write of this entry by one thread's program order -
		// which is also all that compute-sanitizer's racecheck can follow)
		uint32_t first = 0, count = 0;
		if (lane == 0) {
			first = chunk_q[2 * (qr % kQueue)];
			count = chunk_q[2 * (qr % kQueue) + 1];
		}
		first = __shfl_sync(0xFFFFFFFFu, first, 0);
		count = __shfl_sync(0xFFFFFFFFu, count, 0);
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

			// tiles [0, n_full) lie completely inside the frame (uniform over the CTA)
			const uint32_t n_full = strip_full ? P.height / L::kTileRows : 0u;
			bool skip_wait = item == first; // the chunk announcement already waited for tile 0
			// peek = ask early whether the NEXT tile has landed, so that the answer's latency
			// hides behind arithmetic instead of stalling the warp at the top of fetch
			uint32_t landed = 0;
			auto peek_tile = [&]() { landed = mbar_test(bar_full + 8 * stage, phase); };
			// fetch = wait for the tile, read this thread's pixels, remember which stage to hand back
			auto fetch_tile = [&](uint32_t(&p)[N], uint32_t(&q)[N]) -> uint32_t {
				if (!skip_wait && !landed)
					mbar_wait(bar_full + 8 * stage, phase);
				skip_wait = false;
				landed = 0;
				if (SCOPE_LDSM && N % 2 == 0) {
					const uint32_t rows =
						smem_base + L::kStageOff + stage * L::kStageBytes + row0 * (kStripPx * 4);
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
				} else {
					const uint32_t *tile =
						reinterpret_cast<const uint32_t *>(smem + L::kStageOff + stage * L::kStageBytes) +
						row0 * kStripPx + lane;
#pragma unroll
					for (int k = 0; k < N; k++) {
						p[k] = kNeedP ? tile[k * kStripPx] : 0u;
						q[k] = kNeedQ ? tile[k * kStripPx + (L::kLoadRgb ? L::kTileBytes / 4 : 0)] : 0u;
					}
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
			auto release_tile = [&](uint32_t bar, const uint32_t(&p)[N], const uint32_t(&q)[N]) {
				uint32_t dep = 0;
#pragma unroll
				for (int k = 0; k < N; k++)
					dep |= p[k] | q[k];
				__syncwarp();
				if (lane == 0)
					mbar_arrive(bar + (dep & zero));
			};
			uint32_t t = 0;
			if (!SCOPE_PIPELINE) {
				for (; t < n_full; t++) { // sass-loop (only in SCOPE_PIPELINE=0 builds)
					uint32_t p[N], q[N];
					Prep<N> E;
					const uint32_t bar = fetch_tile(p, q);
					release_tile(bar, p, q);
					peek_tile();
					prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, E);
					commit_tile<R_SRC, R_VS, SURFACE, N>(tc, E);
				}
			} else if (n_full > 0) {
				// software pipeline over the interior tiles: atomics of tile t next to the
				// arithmetic of tile t+1 (two Prep register sets, ping-pong); the vectorscope's
				// overflow check of tile t is looked at after that arithmetic (SCOPE_DEFER)
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
				uint32_t bar = fetch_tile(p, q);
				release_tile(bar, p, q);
				peek_tile();
#if SCOPE_DEPHASE
				// (n_full is uniform over the CTA: every consumer warp comes through here once per strip)
				if (warp & 4)
					dephase_bar<NWORK>();
#endif
				prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, A);
#if SCOPE_DEPHASE
				if (!(warp & 4))
					dephase_bar<NWORK>();
#endif
#if SCOPE_STRAIGHT
				// one step: the atomics of `cur`, the arithmetic of `nxt`.  Ordinary tile: one basic block.
				auto step = [&](const Prep<N> &cur, Prep<N> &nxt) {
					const uint32_t bar2 = fetch_tile(p, q);
					bool ordinary = !cur.flat && (R_SRC == SRC_NONE || (cur.all_counted && tc.bins_mask == 7u));
#if SCOPE_RAWFLAT
					ordinary = ordinary && !cur.rawflat;
#endif
#if SCOPE_BALLOT
					ordinary = ordinary && !cur.dominant;
#endif
					if (ordinary) {
						release_tile(bar2, p, q);
						peek_tile();
						commit_issue_fast<R_SRC, R_VS, SURFACE, N>(tc, cur, pend);
						if (!SCOPE_DEFER)
							commit_resolve<R_SRC, R_VS, SURFACE, N>(tc, cur, pend);
						prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, nxt);
					} else {
						// sass-cold{
						commit_issue<R_SRC, R_VS, SURFACE, N>(tc, cur, pend);
						if (!SCOPE_DEFER)
							commit_resolve<R_SRC, R_VS, SURFACE, N>(tc, cur, pend);
						release_tile(bar2, p, q);
						peek_tile();
						prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, nxt);
						// sass-cold}
					}
					resolve(cur);
				};
				for (t = 1; t + 1 < n_full; t += 2) { // sass-loop
					step(A, B);
					step(B, A);
				}
#else
				for (t = 1; t + 1 < n_full; t += 2) { // sass-loop (tools/sass_budget.py: the steady state)
					bar = fetch_tile(p, q);
					issue(A);
					release_tile(bar, p, q);
					peek_tile();
					prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, B);
					resolve(A);
					bar = fetch_tile(p, q);
					issue(B);
					release_tile(bar, p, q);
					peek_tile();
					prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, A);
					resolve(B);
				}
#endif
				if (t < n_full) {
					bar = fetch_tile(p, q);
					issue(A);
					release_tile(bar, p, q);
					prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, B);
					resolve(A);
					commit_tile<R_SRC, R_VS, SURFACE, N>(tc, B);
				} else {
					commit_tile<R_SRC, R_VS, SURFACE, N>(tc, A);
				}
				t = n_full;
			}
			for (; t < tiles; t++) {
				// edge tiles (last rows, last strip): generic body with per-pixel validity
				uint32_t p[N], q[N];
				bool ok[N];
				const uint32_t bar = fetch_tile(p, q);
				release_tile(bar, p, q);
				const uint32_t y0 = t * L::kTileRows + row0;
				if (strip_full && y0 + N <= P.height) {
					// this warp's rows of the partial tile are all inside the frame
					Prep<N> E;
					prepare_tile<R_SRC, R_VS, SURFACE, N>(tc, coef, p, q, E);
					commit_tile<R_SRC, R_VS, SURFACE, N>(tc, E);
				} else if (y0 < P.height) {
#pragma unroll
					for (int k = 0; k < N; k++)
						ok[k] = lane_ok && (y0 + k < P.height);
					process_tile<R_SRC, R_VS, SURFACE, false, N>(tc, coef, p, q, ok);
				}
			}
			if (K_BINS)
				emit_strip<NWORK>(P, wave0, frame, x, lane_ok, warp, lane);
		}
	}
	if (K_VS && cur_frame != 0xFFFFFFFFu)
		flush_vscope<NWORK>(P, vs, cur_frame, tid);
}

template <class L, int NWORK, int EMPTY_ARRIVALS = NWORK>
__device__ __forceinline__ void tma_setup(uint8_t *smem, uint32_t bar_full, uint32_t bar_empty, bool vscope, bool bins,
					  bool worker, int tid)
{
	if (worker)
		zero_bins<NWORK>(reinterpret_cast<uint32_t *>(smem + L::kVsOff),
				 reinterpret_cast<uint32_t *>(smem + L::kWave0Off), vscope, bins, tid);
	if (tid == 0) {
		for (int s = 0; s < L::kStages; s++) {
			mbar_init(bar_full + 8 * s, 1);
			mbar_init(bar_empty + 8 * s, EMPTY_ARRIVALS);
		}
#ifndef SCOPE_EMULATE
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
	}
	__syncthreads();
}

// ---------------------------------------------------------------------------
// strip kernel, TMA loader, every consumer warp does everything: 16 consumer warps take 4 rows
// of each 64-row tile; 1 producer warp.  Used for all scope combinations except the one below.
// ---------------------------------------------------------------------------
// Without the vectorscope's 128 KB of bins two CTAs fit in an SM's shared memory; the register
// budget is then 56 per thread (34 warps).  Since round 2 the kernels WITH a colour transform meet it as well
// (coefficients as immediates, division by 10^6 on the FMA pipe, only the channels the launch needs): the
// luma-only waveform of BASELINE config 4 runs two CTAs per SM = 64 KB of tiles in flight instead of 32.
template <int SRC, bool VSCOPE, bool SURFACE>
constexpr int kTmaMinCtas = !VSCOPE ? 2 : 1;

// SCOPE_MAXNREG (experiment): an explicit register cap instead of the one ptxas derives from the launch bounds
// (for 544 threads it stops at 96, not at the 120 that fit: it seems to round the block up to 640 threads);
// the kernels that run two CTAs per SM keep their 56
#if defined(SCOPE_MAXNREG) && !defined(SCOPE_EMULATE)
#define SCOPE_TMA_BOUNDS __maxnreg__((kTmaMinCtas<SRC, VSCOPE, SURFACE> == 1 ? SCOPE_MAXNREG : 56))
#else
#define SCOPE_TMA_BOUNDS __launch_bounds__((SmemLayout<SRC, VSCOPE, SURFACE, true>::kWarps * 32 + 32), kTmaMinCtas<SRC, VSCOPE, SURFACE>)
#endif
template <int SRC, bool VSCOPE, bool SURFACE, int CS = 0>
__global__ void SCOPE_TMA_BOUNDS
	scope_strip_kernel_tma(const __grid_constant__ StripParams P, const __grid_constant__ CUtensorMap map_rgb,
			       const __grid_constant__ CUtensorMap map_yuv)
{
	using L = SmemLayout<SRC, VSCOPE, SURFACE, true>;
	constexpr int NW = L::kWarps, RPW = L::kRowsPerWarp;
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
	tma_consume<L, SRC, VSCOPE, SURFACE, RPW, NW, SRC != SRC_NONE, VSCOPE, CS>(P, smem, smem_base, chunk_q, bar_full,
										  bar_empty, warp * RPW, warp, lane, tid);
}

} // namespace scope
#include "scope_fused_v3.cuh" // the headline combination (fused mode, RGB bins + vectorscope) as its own kernel
namespace scope {

// the kernels that were measured and not adopted (row-group consumer, warp-specialised kernel): A/B builds
// (-DSCOPE_EXPERIMENT) and the CPU emulator only; the shipped library does not carry them
#if defined(SCOPE_EXPERIMENT) || defined(SCOPE_EMULATE)
#include "scope_kernels_experiments.cuh"
#endif

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
	SCOPE_DYNAMIC_SMEM(smem);
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
	magic = opaque_carrier_bias();
	const TileCtx tc{smem_base + L::kVsOff, wave_lane_addr - kCarrierBias * 128u,
			 wave_lane_addr - kCarrierBias * 128u + kWaveWords * 4, magic, P.bins_mask, lane,
			 P.xform_strict, P.colorspace};
	const Coef coef = P.coef;
	uint32_t cur_frame = 0xFFFFFFFFu;
	// target_scale: this pass's pixel (x, y) is the source pixel (x s + s / 2, y s + s / 2)
	const uint32_t scx = P.scale_x ? P.scale_x : 1u, scy = P.scale_y ? P.scale_y : 1u;

	for (uint32_t item = first; item < last; item++) {
		const uint32_t frame = item / P.strips, strip = item - frame * P.strips;
		if (VSCOPE && frame != cur_frame && cur_frame != 0xFFFFFFFFu)
			flush_vscope<NW>(P, vs, cur_frame, tid);
		cur_frame = frame;
		const uint32_t x = strip * kStripPx + lane;
		const bool lane_ok = x < P.width;
		const bool strip_full = strip * kStripPx + kStripPx <= P.width;
		const size_t col = (size_t)frame * P.frame_stride + ((size_t)(lane_ok ? x : 0) * scx + scx / 2u) * 4;

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
					const size_t o = col + ((size_t)(y0 + k) * scy + scy / 2u) * P.linesize;
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

#ifndef SCOPE_EMULATE
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
// (plane 1 = the R|V channel is only read when the waveform has that channel: otherwise it need not hold valid data)
__global__ void __launch_bounds__(256) wave_pairs_finalize_kernel(const uint32_t *pairs, uint8_t *wave, size_t n_px,
								  int plane1)
{
	const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n_px)
		return;
	const uint32_t w0 = pairs[i], w1 = plane1 ? pairs[n_px + i] : 0u;
	reinterpret_cast<uint32_t *>(wave)[i] =
		min(w0 & 0xFFFFu, 255u) | (min(w0 >> 16, 255u) << 8) | (min(w1 & 0xFFFFu, 255u) << 16);
}

// test hook: the kernel's own transform for all 2^24 colours (index r<<16|g<<8|b)
__global__ void __launch_bounds__(256) yuv_table_strict_kernel(int colorspace, uint32_t *out)
{
	const uint32_t i = (blockIdx.x * 256 + threadIdx.x) * 2;
#pragma unroll
	for (uint32_t j = i; j < i + 2; j++) {
		uint32_t uyv[3];
		rgb_to_yuv_strict(j, colorspace, uyv);
		out[j] = uyv[0] | (uyv[1] << 8) | (uyv[2] << 16);
	}
}

__global__ void __launch_bounds__(256) yuv_table_kernel(Coef coef, uint32_t *out)
{
	const uint32_t i = (blockIdx.x * 256 + threadIdx.x) * 2;
	// index r<<16|g<<8|b is already the little-endian BGRA word b | g<<8 | r<<16
	uint32_t magic;
	magic = opaque_carrier_bias();
#pragma unroll
	for (uint32_t j = i; j < i + 2; j++) {
		const uint32_t bgr[3] = {carrier<0>(j, magic), carrier<1>(j, magic), carrier<2>(j, magic)};
		uint32_t hi[3];
		rgb_to_yuv_hi<true>(bgr, coef, hi);
		out[j] = __byte_perm(__byte_perm(hi[0], hi[1], 0x3362), hi[2], 0x3610);
	}
}

#endif // !SCOPE_EMULATE

} // namespace scope
