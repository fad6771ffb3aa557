// scope_fused_v3.cuh — the headline pass as its own kernel: histogram RGB + waveform RGB + vectorscope from ONE
// BGRA plane (SCOPE_MODE_FUSED, all three channels), i.e. BASELINE configs 2 and 5.  Included by scope_kernels.cuh
// (namespace scope); every other scope combination keeps scope_strip_kernel_tma / _ldg.
//
// Same decomposition as the general kernel (strips of 32 columns, lane = column, TMA ring filled by a producer warp,
// one ldmatrix.x4 per warp and visit, histogram derived from the column bins), but the per-pixel work is rebuilt
// around what the round-2 profile showed (profiles/r02/ncu_r02b.md): all four of issue slots (72 %), the shared-memory
// data pipe (62 %), the FMA-heavy pipe (49 %: IMAD.HI costs four IMADs) and the ALU pipe (49 %) were loaded, so all
// four had to shrink together.
//
//  * vectorscope bins: u16 halves with the take-back rule of the general kernel (an add that finds its half at
//    >= 0x8000 is undone), the check of a visit's four adds deferred behind the next visit's loads.
//    (A return-less RED formulation with a sweeping producer warp was built and measured first, profiles/r02/v3_red.md:
//    only `red.add 1` - SASS ATOMS.POPC.INC - merges the lanes of a warp that hit one word; an add of 65536 for the
//    upper half is ATOMS.ADD and serialises them whether or not it returns a value (tools/ubench4.cu), and the sweep's
//    shared-memory loads sat between the producer's TMA issues: 18 % of the HBM peak.)
//  * the division S / 10^6 of the exact transform is done on the FMA pipe: T = S >> 6 (10^6 = 64 * 15625) enters a
//    float through a funnel shift that also supplies the exponent (bits = 0x4B000000 | T, value 2^23 + T),
//    q = round((T - 7812) / 15625) = floor(T / 15625) comes out of one add and one fused multiply-add in the low byte
//    of the result's bit pattern (the addend 1.5 * 2^23 keeps the sum in the binade whose ulp is 1; the rounding error
//    of the product, < 1.7e-5, is below the 3.2e-5 that separates (T - 7812) / 15625 from a half-integer).
//    Checked for all 2^24 colours against the oracle (tests/test_gpu_parity.py::test_transform_exhaustive_v3).
//  * U and V of a pixel go through the add and the multiply-add as ONE f32x2 instruction each (FADD2 / FFMA2), the six
//    column-bin addresses of two pixels as three FFMA2, the vectorscope's address and addend as one.
//  * vectorscope bin index = U + 260 (V - 16): ONE IMAD on the two float bit patterns (their biases cancel mod 2^16).
//    V of the fused transform lies in [16, 240], so the index is below 58 496 < 2^16 and the map is injective; bank =
//    (U + 4 V) mod 32, the additive swizzle that round 1 measured best but could not afford in four instructions.
//  * end of a strip: clamps with one packed min per plane and one PRMT per output word.
#pragma once

namespace scope {

#ifndef SCOPE_V3
#define SCOPE_V3 1 // 0: the general kernel serves the headline combination as well (A/B builds)
#endif
#ifndef SCOPE_V3_WARPS
#define SCOPE_V3_WARPS 27
#endif
#ifndef SCOPE_V3_STAGES
#define SCOPE_V3_STAGES 3
#endif
#ifndef SCOPE_V3_L2_AHEAD
#define SCOPE_V3_L2_AHEAD 6 // tiles the producer's L2 prefetch runs ahead of its loads (0: none)
#endif
#ifndef SCOPE_V3_PARAM_CONSTS
#define SCOPE_V3_PARAM_CONSTS 1 // the kernel's constants come from the launch parameters (0: immediates, A/B partner)
#endif
#ifndef SCOPE_V3_FFMA2
#define SCOPE_V3_FFMA2 1 // 0: scalar FFMA / FADD (A/B partner)
#endif
#ifndef SCOPE_V3_RESOLVE_NOW
#define SCOPE_V3_RESOLVE_NOW 1 // the ordinary block issues its vectorscope adds FIRST and looks at their old values at its own
                               // end, behind the twelve column-bin adds (0: the look is deferred to the next visit, which
                               // keeps twelve registers alive across the visit boundary)
#endif
#ifndef SCOPE_V3_SMEM_CONSTS
#define SCOPE_V3_SMEM_CONSTS 1 // the per-visit constants that must sit in registers (IMAD addends, PRMT carrier, f32x2 addends) are
                               // read ONCE per thread from a shared-memory copy: ptxas re-materialises anything it can trace
                               // to the constant bank with an LDC per visit (nine of them), a loaded value it has to keep
#endif
#ifndef SCOPE_V3_WIDE_EMIT
#define SCOPE_V3_WIDE_EMIT 1 // end of a full strip: 16-byte loads / stores, four levels per warp and step (0: one level per step)
#endif
#ifndef SCOPE_V3_FRAME_AFFINE
#define SCOPE_V3_FRAME_AFFINE 1 // strips are claimed frame by frame (StripParams::frame_affine; 0: one counter over the batch)
#endif
#ifndef SCOPE_V3_DIAG
#define SCOPE_V3_DIAG 0 // diagnostic builds (never shipped, results wrong by construction): bit 0 no end-of-strip write-out
                        // (its two barriers stay), bit 1 not even the barriers, bit 2 no vectorscope flush
#endif
#ifndef SCOPE_V3_BG_SKIP
#define SCOPE_V3_BG_SKIP 1 // blocks with lanes whose four pixels are equal (screen content: text on a flat background) go through
                           // v3_block_mixed: the lanes that hold the background colour leave their vectorscope adds to ONE lane
#endif
#ifndef SCOPE_V3_LEAN
#define SCOPE_V3_LEAN 1 // visits that lie inside the frame and have a successor run without the per-visit checks (0: A/B partner)
#endif

struct V3 {
	static constexpr int kWarps = SCOPE_V3_WARPS; // consumer warps; + 1 producer warp
	static constexpr int kRows = 4;               // rows per warp and visit (one ldmatrix.x4)
	static constexpr int kTileRows = kWarps * kRows;
	static constexpr int kTileBytes = kTileRows * kStripPx * 4;
	static constexpr int kStages = SCOPE_V3_STAGES;
	static constexpr int kThreads = (kWarps + 1) * 32;
	// vectorscope table: bin (U, V) lives in word (U & 127) + 132 V, half U >> 7.  V of the fused transform lies in
	// [16, 240] for every colour (tests/test_oracle.py), so only words [132 * 16, 132 * 241) exist: 29 700 words =
	// 116 KB instead of 128 KB, which is what pays for the ring's fourth stage.  The stride 132 = 128 + 4 makes the
	// bank (U + 4 V) mod 32 - the additive swizzle - and keeps a word's two bins 128 U apart (neighbouring colours
	// never share a word: same-word lanes serialise like bank conflicts do).
	static constexpr uint32_t kVStride = 132, kVMin = 16, kVMax = 240;
	static constexpr int kVsFirstWord = kVStride * kVMin;                   // 2112
	static constexpr int kVsTableWords = (kVStride * (kVMax + 1) - kVsFirstWord + 3) / 4 * 4; // 29 700
	static constexpr int kWaveOff = 0;
	static constexpr int kVsOff = 2 * kWaveWords * 4;                       // behind the column bins
	static_assert(kVsOff >= kVsFirstWord * 4, "the table's virtual base (kVsOff - 4 * kVsFirstWord) must not be negative");
	static constexpr int kStageOff = (kVsOff + kVsTableWords * 4 + 127) / 128 * 128;
	static constexpr int kBarOff = kStageOff + kStages * kTileBytes;
	static constexpr int kQueueOff = kBarOff + 2 * 8 * kStages + 16;
	static constexpr int kConstOff = kQueueOff + kQueue * 8 + 16; // 16 words: SCOPE_V3_SMEM_CONSTS
	static constexpr int kTotal = kConstOff + 64;
	static_assert(kTotal <= 227 * 1024, "shared memory");
#ifndef SCOPE_EMULATE // (the emulation tests build a copy with a one-entry mailbox on purpose)
	static_assert(kStages <= kQueue, "chunk mailbox shorter than the ring");
#endif
};

// ---------------------------------------------------------------------------
// helpers (each with its emulated twin, like the ones in scope_kernels.cuh)
// ---------------------------------------------------------------------------
// two fp32 lanes in one 64-bit value
struct F2 {
#ifdef SCOPE_EMULATE
	uint32_t lo, hi;
#else
	unsigned long long v;
#endif
};
__device__ __forceinline__ F2 f2_pack(uint32_t lo, uint32_t hi)
{
	F2 r;
#ifdef SCOPE_EMULATE
	r.lo = lo;
	r.hi = hi;
#else
	asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "r"(lo), "r"(hi));
#endif
	return r;
}
__device__ __forceinline__ void f2_unpack(const F2 &a, uint32_t &lo, uint32_t &hi)
{
#ifdef SCOPE_EMULATE
	lo = a.lo;
	hi = a.hi;
#else
	asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(a.v));
#endif
}
__device__ __forceinline__ uint32_t f32_bits(float f)
{
#ifdef SCOPE_EMULATE
	uint32_t r;
	memcpy(&r, &f, 4);
	return r;
#else
	return __float_as_uint(f);
#endif
}
// a * b + c per lane on bit patterns (round to nearest, denormals kept: no .ftz)
__device__ __forceinline__ F2 f2_fma(const F2 &a, const F2 &b, const F2 &c)
{
	F2 d;
#ifdef SCOPE_EMULATE
	float fb0, fb1;
	memcpy(&fb0, &b.lo, 4);
	memcpy(&fb1, &b.hi, 4);
	d.lo = emul::fma_bits(a.lo, fb0, c.lo);
	d.hi = emul::fma_bits(a.hi, fb1, c.hi);
#elif SCOPE_V3_FFMA2
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
#else
	uint32_t a0, a1, b0, b1, c0, c1;
	f2_unpack(a, a0, a1);
	f2_unpack(b, b0, b1);
	f2_unpack(c, c0, c1);
	d = f2_pack(fma_bits(a0, __uint_as_float(b0), c0), fma_bits(a1, __uint_as_float(b1), c1));
#endif
	return d;
}
__device__ __forceinline__ F2 f2_add(const F2 &a, const F2 &b)
{
	F2 d;
#ifdef SCOPE_EMULATE
	float x0, x1, y0, y1;
	memcpy(&x0, &a.lo, 4);
	memcpy(&x1, &a.hi, 4);
	memcpy(&y0, &b.lo, 4);
	memcpy(&y1, &b.hi, 4);
	const float s0 = x0 + y0, s1 = x1 + y1;
	memcpy(&d.lo, &s0, 4);
	memcpy(&d.hi, &s1, 4);
#elif SCOPE_V3_FFMA2
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
#else
	uint32_t a0, a1, b0, b1;
	f2_unpack(a, a0, a1);
	f2_unpack(b, b0, b1);
	d = f2_pack(__float_as_uint(__fadd_rn(__uint_as_float(a0), __uint_as_float(b0))),
		    __float_as_uint(__fadd_rn(__uint_as_float(a1), __uint_as_float(b1))));
#endif
	return d;
}

// per-halfword unsigned minimum
__device__ __forceinline__ uint32_t min_u16x2(uint32_t a, uint32_t b)
{
#ifdef SCOPE_EMULATE
	const uint32_t lo = (a & 0xFFFFu) < (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
	const uint32_t hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
	return lo | (hi << 16);
#else
	return __vminu2(a, b);
#endif
}

// ---------------------------------------------------------------------------
// the transform: B, G, R bytes (plain integers) -> bit patterns 0x4B400000 + 61376 + U and 0x4B400000 + V
// ---------------------------------------------------------------------------
constexpr uint32_t kV3UBias = 0u; // (no bias on U in the float result any more)
struct V3Coef {
	uint32_t u[3], v[3]; // 10^6 x the effect file's coefficients, order R, G, B
	uint32_t ku, kv;     // rounding/offset constants incl. kDivExpBits (Coef)
};
template <int CS>
__device__ __forceinline__ constexpr V3Coef v3_coef()
{
	constexpr Coef c = const_coef<CS>();
	V3Coef r{};
	for (int i = 0; i < 3; i++) {
		r.u[i] = c.u[i];
		r.v[i] = c.v[i];
	}
	r.ku = c.ku;
	r.kv = c.kv;
	return r;
}
inline V3Coef v3_coef_for(int colorspace)
{
	const Coef c = coef_for(colorspace);
	V3Coef r{};
	for (int i = 0; i < 3; i++) {
		r.u[i] = c.u[i];
		r.v[i] = c.v[i];
	}
	r.ku = c.ku;
	r.kv = c.kv;
	return r;
}

struct V3Consts {
	F2 neg_bias; // (-(2^23 + 7812)) x 2
	F2 inv;      // fl(1 / 15625) x 2
	F2 round;    // (1.5 * 2^23 + 61376, 1.5 * 2^23)
	F2 k128;     // 128.0 x 2
	F2 wb0, wb1; // this lane's plane-0 / plane-1 base address x 2
	F2 vs_mul;   // (4.0, 65535 / 128)
	F2 vs_add;   // (address word 0 of the vectorscope table WOULD have, 1)
	uint32_t exp_hi; // 0x12: the upper bits of the funnel shift
};
// the values of StripParams::v3c (host side)
inline void v3_param_consts(uint32_t (&c)[8])
{
	auto bits = [](float f) {
		uint32_t r;
		memcpy(&r, &f, 4);
		return r;
	};
	c[0] = bits(12582912.0f + (float)kV3UBias);
	c[1] = bits(12582912.0f);
	c[2] = bits(4.0f);
	c[3] = bits(65535.0f / 128.0f);
	c[4] = bits(1.0f / 15625.0f);
	c[5] = bits(-(8388608.0f + 7812.0f));
	c[6] = bits(128.0f);
	c[7] = 0x12u;
}
__device__ __forceinline__ V3Consts v3_consts_from(const uint32_t (&pc)[8], uint32_t smem_base, int lane)
{
	V3Consts c;
	c.neg_bias = f2_pack(pc[5], pc[5]);
	c.inv = f2_pack(pc[4], pc[4]);
	c.round = f2_pack(pc[0], pc[1]);
	c.k128 = f2_pack(pc[6], pc[6]);
	const uint32_t w0 = smem_base + V3::kWaveOff + lane * 4;
	c.wb0 = f2_pack(w0, w0);
	c.wb1 = f2_pack(w0 + kWaveWords * 4, w0 + kWaveWords * 4);
	c.vs_mul = f2_pack(pc[2], pc[3]);
	c.vs_add = f2_pack(smem_base + V3::kVsOff - 4u * V3::kVsFirstWord, 1u);
	c.exp_hi = pc[7];
	return c;
}
__device__ __forceinline__ V3Consts v3_consts(uint32_t smem_base, int lane)
{
	V3Consts c;
	const uint32_t nb = f32_bits(-(8388608.0f + 7812.0f));
	c.neg_bias = f2_pack(nb, nb);
	const uint32_t inv = f32_bits(1.0f / 15625.0f);
	c.inv = f2_pack(inv, inv);
	c.round = f2_pack(f32_bits(12582912.0f + (float)kV3UBias), f32_bits(12582912.0f));
	const uint32_t k128 = f32_bits(128.0f);
	c.k128 = f2_pack(k128, k128);
	const uint32_t w0 = smem_base + V3::kWaveOff + lane * 4;
	c.wb0 = f2_pack(w0, w0);
	c.wb1 = f2_pack(w0 + kWaveWords * 4, w0 + kWaveWords * 4);
	c.vs_mul = f2_pack(f32_bits(4.0f), f32_bits(65535.0f / 128.0f));
	c.vs_add = f2_pack(smem_base + V3::kVsOff - 4u * V3::kVsFirstWord, 1u);
	c.exp_hi = 0x12u;
	return c;
}

// U, V of one pixel as float bit patterns (see the header comment); cb, cg, cr = the bytes as integers
template <int CS>
__device__ __forceinline__ F2 v3_uv(uint32_t cb, uint32_t cg, uint32_t cr, const V3Consts &k, uint32_t ku, uint32_t kv)
{
	constexpr V3Coef c = v3_coef<CS>();
	uint32_t su = cr * c.u[0] + ku;
	su = cg * c.u[1] + su;
	su = cb * c.u[2] + su;
	uint32_t sv = cr * c.v[0] + kv;
	sv = cg * c.v[1] + sv;
	sv = cb * c.v[2] + sv;
	// bits 0x4B000000 | (S >> 6): the float 2^23 + T
	const F2 t = f2_pack(funnel_r(su, k.exp_hi, 6u), funnel_r(sv, k.exp_hi, 6u));
	return f2_fma(f2_add(t, k.neg_bias), k.inv, k.round);
}

// the vectorscope word address and addend of one pixel from v3_uv's result
__device__ __forceinline__ void v3_vs_target(const F2 &uv, const V3Consts &k, uint32_t &addr, uint32_t &add)
{
	uint32_t ru, rv;
	f2_unpack(uv, ru, rv);
	// word = (U & 127) + 132 V: ONE IMAD on the two float bit patterns (0x4B400000 + U with bit 7 cleared, 0x4B400000 + V;
	// the biases add up to a multiple of 2^22); half = bit 7 of U
	const uint32_t idx = rv * V3::kVStride + (ru & ~0x80u);
	const F2 t = f2_fma(f2_pack(idx & 0xFFFFu, ru & 0x80u), k.vs_mul, k.vs_add);
	f2_unpack(t, addr, add);
}

// test hook: (U | V << 8) of the kernel's own transform
template <int CS>
__device__ __forceinline__ uint32_t v3_uv_bytes(uint32_t pixel, uint32_t smem_base_unused)
{
	const V3Consts k = v3_consts(smem_base_unused, 0);
	constexpr V3Coef c = v3_coef<CS>();
	const uint32_t cb = pixel & 0xFFu, cg = (pixel >> 8) & 0xFFu, cr = (pixel >> 16) & 0xFFu;
	uint32_t ru, rv;
	f2_unpack(v3_uv<CS>(cb, cg, cr, k, c.ku, c.kv), ru, rv);
	return ((ru - kV3UBias) & 0xFFu) | ((rv & 0xFFu) << 8);
}

// ---------------------------------------------------------------------------
// accumulation of one warp's 4 x 32 block
// ---------------------------------------------------------------------------
struct V3Px {
	uint32_t cb, cg, cr;
};
__device__ __forceinline__ V3Px v3_bytes(uint32_t p, uint32_t zero_reg)
{
	V3Px r;
	r.cb = carrier<0>(p, zero_reg);
	r.cg = carrier<1>(p, zero_reg);
	r.cr = carrier<2>(p, zero_reg);
	return r;
}

// the vectorscope adds of one visit whose old values have not been looked at yet
struct V3Pend {
	uint32_t addr[4], add[4], old[4];
};
// an add that found its half at >= 0x8000 is taken back: add << 15 is the half's top bit
__device__ __forceinline__ void v3_resolve(const V3Pend &q)
{
	const uint32_t any = q.old[0] | q.old[1] | q.old[2] | q.old[3];
	if (any & 0x80008000u) {
		// sass-cold{
#pragma unroll
		for (int i = 0; i < 4; i++)
			if (q.old[i] & (q.add[i] << 15))
				red_shared(q.addr[i], 0u - q.add[i]);
		// sass-cold}
	}
}
__device__ __forceinline__ void v3_pend_clear(V3Pend &q)
{
#pragma unroll
	for (int i = 0; i < 4; i++)
		q.old[i] = 0u;
}

// every pixel counted, all rows inside the frame, block not flat
// (SCOPE_V3_SKIP: diagnostic builds, never shipped - bit 0 leaves out the column-bin adds, bit 1 the vectorscope's;
// the addresses are still computed and kept alive)
#ifndef SCOPE_V3_SKIP
#define SCOPE_V3_SKIP 0
#endif
__device__ __forceinline__ void v3_keep(uint32_t a, uint32_t b)
{
#ifndef SCOPE_EMULATE
	asm volatile("" ::"r"(a), "r"(b));
#endif
}
template <int CS>
__device__ __forceinline__ void v3_block_fast(const uint32_t (&p)[4], const V3Consts &k, uint32_t ku, uint32_t kv,
					      uint32_t zero_reg, V3Pend &q)
{
	V3Px px[4];
#pragma unroll
	for (int i = 0; i < 4; i++)
		px[i] = v3_bytes(p[i], zero_reg);
#pragma unroll
	for (int i = 0; i < 4; i += 2) {
		uint32_t a0, a1;
		f2_unpack(f2_fma(f2_pack(px[i].cb, px[i + 1].cb), k.k128, k.wb0), a0, a1);
		if (SCOPE_V3_SKIP & 1) {
			v3_keep(a0, a1);
		} else {
			red_shared(a0, 1u);
			red_shared(a1, 1u);
		}
		f2_unpack(f2_fma(f2_pack(px[i].cg, px[i + 1].cg), k.k128, k.wb0), a0, a1);
		if (SCOPE_V3_SKIP & 1) {
			v3_keep(a0, a1);
		} else {
			red_shared(a0, 0x10000u);
			red_shared(a1, 0x10000u);
		}
		f2_unpack(f2_fma(f2_pack(px[i].cr, px[i + 1].cr), k.k128, k.wb1), a0, a1);
		if (SCOPE_V3_SKIP & 1) {
			v3_keep(a0, a1);
		} else {
			red_shared(a0, 1u);
			red_shared(a1, 1u);
		}
	}
	uint32_t addr[4], add[4];
#pragma unroll
	for (int i = 0; i < 4; i++)
		v3_vs_target(v3_uv<CS>(px[i].cb, px[i].cg, px[i].cr, k, ku, kv), k, addr[i], add[i]);
#pragma unroll
	for (int i = 0; i < 4; i++) {
		q.addr[i] = addr[i];
		q.add[i] = add[i];
		if (SCOPE_V3_SKIP & 2) {
			v3_keep(q.addr[i], q.add[i]);
			q.old[i] = 0;
		} else if (SCOPE_V3_SKIP & 4) { // diagnostic: no return value (saturation of flat content is then wrong)
			red_shared(q.addr[i], q.add[i]);
			q.old[i] = 0;
		} else {
			q.old[i] = atom_shared_add(q.addr[i], q.add[i]);
		}
	}
}

// the same block with the vectorscope adds in front: their old values have arrived by the time the column-bin adds
// have been issued, so the take-back needs no state that outlives the block
// MIXED (cold paths only): lanes with `bg` leave out their vectorscope adds, and a lane with lead_count != 0 adds that
// many to the bin of its first pixel at the end (see v3_block_mixed)
template <int CS, bool MIXED = false>
__device__ __forceinline__ void v3_block_fast_now(const uint32_t (&p)[4], const V3Consts &k, uint32_t ku, uint32_t kv,
						  uint32_t zero_reg, bool bg = false, uint32_t lead_count = 0u)
{
	V3Px px[4];
#pragma unroll
	for (int i = 0; i < 4; i++)
		px[i] = v3_bytes(p[i], zero_reg);
	V3Pend q;
#pragma unroll
	for (int i = 0; i < 4; i++)
		v3_vs_target(v3_uv<CS>(px[i].cb, px[i].cg, px[i].cr, k, ku, kv), k, q.addr[i], q.add[i]);
#pragma unroll
	for (int i = 0; i < 4; i++) {
		if (SCOPE_V3_SKIP & 2) {
			v3_keep(q.addr[i], q.add[i]);
			q.old[i] = 0;
		} else if (MIXED) {
			q.old[i] = 0u;
			if (!bg)
				q.old[i] = atom_shared_add(q.addr[i], q.add[i]);
		} else {
			q.old[i] = atom_shared_add(q.addr[i], q.add[i]);
		}
	}
#pragma unroll
	for (int i = 0; i < 4; i += 2) {
		uint32_t a0, a1;
		f2_unpack(f2_fma(f2_pack(px[i].cb, px[i + 1].cb), k.k128, k.wb0), a0, a1);
		if (SCOPE_V3_SKIP & 1) {
			v3_keep(a0, a1);
		} else {
			red_shared(a0, 1u);
			red_shared(a1, 1u);
		}
		f2_unpack(f2_fma(f2_pack(px[i].cg, px[i + 1].cg), k.k128, k.wb0), a0, a1);
		if (SCOPE_V3_SKIP & 1) {
			v3_keep(a0, a1);
		} else {
			red_shared(a0, 0x10000u);
			red_shared(a1, 0x10000u);
		}
		f2_unpack(f2_fma(f2_pack(px[i].cr, px[i + 1].cr), k.k128, k.wb1), a0, a1);
		if (SCOPE_V3_SKIP & 1) {
			v3_keep(a0, a1);
		} else {
			red_shared(a0, 1u);
			red_shared(a1, 1u);
		}
	}
	v3_resolve(q);
	if (MIXED && lead_count != 0u) { // the flat block's rule (v3_block_flat): an add that finds its half at >= 0x8000 is undone
		const uint32_t old = atom_shared_add(q.addr[0], q.add[0] * lead_count);
		if (old & (q.add[0] << 15))
			red_shared(q.addr[0], 0u - q.add[0] * lead_count);
	}
}

template <int CS>
__device__ __forceinline__ void v3_block_ordinary(const uint32_t (&p)[4], const V3Consts &k, uint32_t ku, uint32_t kv,
						  uint32_t zero_reg, V3Pend &q)
{
#if SCOPE_V3_RESOLVE_NOW
	(void)q;
	v3_block_fast_now<CS>(p, k, ku, kv, zero_reg);
#else
	v3_block_fast<CS>(p, k, ku, kv, zero_reg, q);
#endif
}

// all 4 x 32 pixel words equal (solid regions, letterbox bars): 32 lanes on one vectorscope word would take 32
// cycles per add; here lane 0 adds 128 at once.  The column bins have no such problem (lane = column = bank).
template <int CS>
__device__ __forceinline__ void v3_block_flat(uint32_t p0, const V3Consts &k, uint32_t ku, uint32_t kv, uint32_t zero_reg,
					      int lane)
{
	const V3Px px = v3_bytes(p0, zero_reg);
	uint32_t a0, a1;
	f2_unpack(f2_fma(f2_pack(px.cb, px.cg), k.k128, k.wb0), a0, a1);
	red_shared(a0, 4u);
	red_shared(a1, 4u << 16);
	f2_unpack(f2_fma(f2_pack(px.cr, px.cr), k.k128, k.wb1), a0, a1);
	red_shared(a0, 4u);
	if (lane == 0) {
		uint32_t addr, add;
		v3_vs_target(v3_uv<CS>(px.cb, px.cg, px.cr, k, ku, kv), k, addr, add);
		const uint32_t old = atom_shared_add(addr, add * 128u);
		if (old & (add << 15))
			red_shared(addr, 0u - add * 128u);
	}
}

// a consumer warp's standing state (used from here on)
struct V3Warp {
	V3Consts k;
	uint32_t ku, kv, zero;
	uint32_t rows_base; // this lane's ldmatrix address inside stage 0
	uint32_t base;      // the CTA's shared-memory window
	uint32_t bar_full, bar_empty;
	uint32_t y_warp;
	int lane;
};

// An opaque block inside the frame in which SOME lanes hold four equal pixels but not all lanes the same one: screen
// content - text, window edges - on a flat background.  Left to the ordinary block, most of the 32 lanes of every
// vectorscope add would sit on the background's word and be served one after the other (32-way: 16 % of the HBM peak
// on `--content ui`).  Here the background is the colour of the first such lane; the lanes that hold nothing else
// (`bg`) skip their four vectorscope adds and that first lane adds 4 x their number at once, with the flat block's
// take-back rule.  Everything else is the ordinary block - the SAME copy of it (v3_block_fast_now<CS, true>): with
// fewer than eight such lanes no lane is `bg` and the call is the ordinary block.  (Two earlier forms cost the
// headline more than screen content gained, profiles/r02/c39 and c40: one out-of-line copy per kernel, `__noinline__`
// - the call's ABI took 3 % off every content and 10 % off solid frames; a second inlined block per cold path - 1.5 %
// off the mixed batch through code size alone, although picture-like content never enters it.)
template <int CS>
__device__ __forceinline__ void v3_block_mixed(const uint32_t (&p)[4], bool flat, const V3Warp &w)
{
	bool bg = false;
	uint32_t lead_count = 0u;
#if SCOPE_V3_BG_SKIP
	const uint32_t fm = __ballot_sync(0xFFFFFFFFu, flat);
	if (__popc(fm) >= 8) {
		const int leader = __ffs((int)fm) - 1;
		const uint32_t key = __shfl_sync(0xFFFFFFFFu, p[0], leader);
		bg = flat && p[0] == key;
		const uint32_t n_bg = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, bg));
#ifdef SCOPE_EMULATE
		const int lane = w.lane;
#else
		int lane; // (read where it is needed: a longer life of w.lane costs the visits' loop six instructions)
		asm("mov.u32 %0, %%laneid;" : "=r"(lane));
#endif
		lead_count = lane == leader ? 4u * n_bg : 0u;
	}
#else
	(void)flat;
#endif
	v3_block_fast_now<CS, true>(p, w.k, w.ku, w.kv, w.zero, bg, lead_count);
}

// rows partly outside the frame, columns outside it, transparent pixels: per-pixel conditions
// (`rows_ok` bit i: row i of the block is inside the frame)
template <int CS>
__device__ __forceinline__ void v3_block_slow(const uint32_t (&p)[4], const V3Consts &k, uint32_t ku, uint32_t kv,
					      uint32_t zero_reg, bool lane_ok, uint32_t rows_ok)
{
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const bool ok = lane_ok && ((rows_ok >> i) & 1u);
		if (!ok)
			continue;
		const V3Px px = v3_bytes(p[i], zero_reg);
		if (p[i] > 0x00FFFFFFu) { // alpha != 0 (histogram.c:385-387, waveform.c:246-248)
			uint32_t a0, a1;
			f2_unpack(f2_fma(f2_pack(px.cb, px.cg), k.k128, k.wb0), a0, a1);
			red_shared(a0, 1u);
			red_shared(a1, 0x10000u);
			f2_unpack(f2_fma(f2_pack(px.cr, px.cr), k.k128, k.wb1), a0, a1);
			red_shared(a0, 1u);
		}
		// the vectorscope never looks at alpha (vectorscope.c:228-231)
		uint32_t addr, add;
		v3_vs_target(v3_uv<CS>(px.cb, px.cg, px.cr, k, ku, kv), k, addr, add);
		const uint32_t old = atom_shared_add(addr, add);
		if (old & (add << 15))
			red_shared(addr, 0u - add);
	}
}

// ---------------------------------------------------------------------------
// vectorscope flush: the consumer warps move the u16 halves to the frame's u32 accumulators
// ---------------------------------------------------------------------------
// table word (0-based inside the table) + half -> offset in the frame's u32 accumulators (row = 255 - V,
// vectorscope.c:232); the four slots per 132 that hold no bin are never written
__device__ __forceinline__ uint32_t v3_acc_offset(uint32_t word, uint32_t half)
{
	const uint32_t w = word + V3::kVsFirstWord;
	const uint32_t v = ((w >> 2) * 1986u) >> 16; // w / 132 for w < 32768
	const uint32_t u = w - v * V3::kVStride + 128u * half;
	return (255u - v) * 256u + u;
}

__device__ __forceinline__ void v3_flush(const StripParams &P, uint32_t *vs, uint32_t frame, int tid)
{
	constexpr int kStep = V3::kWarps * 32, kBatch = 4;
	workers_bar<V3::kWarps>();
	uint32_t *acc = P.vscope_acc + (size_t)frame * P.vscope_stride;
	for (int i0 = tid; i0 < V3::kVsTableWords / 4; i0 += kBatch * kStep) {
		uint4 wv[kBatch];
#pragma unroll
		for (int b = 0; b < kBatch; b++) {
			const int i = i0 + b * kStep;
			wv[b] = i < V3::kVsTableWords / 4 ? reinterpret_cast<uint4 *>(vs)[i] : make_uint4(0, 0, 0, 0);
		}
#pragma unroll
		for (int b = 0; b < kBatch; b++) {
			const int i = i0 + b * kStep;
			const uint4 w = wv[b];
			if ((w.x | w.y | w.z | w.w) != 0u) {
				const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const uint32_t word = (uint32_t)i * 4u + (uint32_t)j;
					if (ww[j] & 0xFFFFu)
						atomicAdd(acc + v3_acc_offset(word, 0u), ww[j] & 0xFFFFu);
					if (ww[j] >> 16)
						atomicAdd(acc + v3_acc_offset(word, 1u), ww[j] >> 16);
				}
				reinterpret_cast<uint4 *>(vs)[i] = make_uint4(0, 0, 0, 0);
			}
		}
	}
	workers_bar<V3::kWarps>();
}

// ---------------------------------------------------------------------------
// end of a strip: final saturated waveform rows of the CTA's 32 columns, the strip's share of the histogram,
// bins back to zero (reference layouts: waveform.c:240-256, histogram.c:379-395)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void v3_emit_strip(const StripParams &P, uint32_t *wave0, uint32_t frame, uint32_t x,
					      bool lane_ok, int warp, int lane)
{
	constexpr int NW = V3::kWarps;
	workers_bar<NW>();
	uint32_t *hist = P.hist + (size_t)frame * P.hist_stride;
	const bool do_hist = P.hist_mask != 0u;
	// row 255 - v of the frame's BGRX image, this lane's column; walked with a constant stride
	uint32_t *row = reinterpret_cast<uint32_t *>(P.wave + (size_t)frame * P.wave_stride) + (P.x_offset + x) +
			(size_t)(255 - warp) * P.out_width;
	const size_t step = (size_t)NW * P.out_width;
	uint32_t keep_b = 0, keep_g = 0, keep_r = 0;
	int i = 0;
#pragma unroll 4
	for (int v = warp; v < 256; v += NW, i++, row -= step) {
		const uint32_t w0 = wave0[v * 32 + lane];
		const uint32_t w1 = wave0[kWaveWords + v * 32 + lane];
		if (!__any_sync(0xFFFFFFFFu, (w0 | w1) != 0u)) {
			if (lane_ok)
				*row = 0u;
			continue;
		}
		wave0[v * 32 + lane] = 0;
		wave0[kWaveWords + v * 32 + lane] = 0;
		if (do_hist) {
			const uint32_t sb = __reduce_add_sync(0xFFFFFFFFu, w0 & 0xFFFFu);
			const uint32_t sg = __reduce_add_sync(0xFFFFFFFFu, w0 >> 16);
			const uint32_t sr = __reduce_add_sync(0xFFFFFFFFu, w1);
			if (lane == i) {
				keep_b = sb;
				keep_g = sg;
				keep_r = sr;
			}
		}
		if (lane_ok) // bytes B, G, R, 0: plane 0 holds (B : lo16, G : hi16), plane 1 holds R
			*row = __byte_perm(min_u16x2(w0, 0x00FF00FFu), min(w1, 255u), 0x5420);
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
}

// The same for a strip whose 32 columns all lie inside the frame and whose output rows can be written 16 bytes at a
// time: a lane takes FOUR neighbouring columns of one level (one 16-byte load per plane, conflict-free: eight lanes
// cover a level's 128 bytes), a warp four levels per step - 64 steps per strip instead of 256, and an empty group of
// four levels costs a vote and one store.  A warp's (up to three) steps are loaded up front, so that their latencies
// overlap.  The level's histogram share is summed over the lane's four columns first (B and G still packed: 8 x
// height <= 65 535, checked by the caller) and then over the eight lanes of the level by a TRANSPOSING butterfly:
// after the first exchange the even lanes carry B|G and the odd lanes R, after the second the lanes 0 / 2 / 1 of
// each eight carry B / G / R - three shuffles instead of nine, and one atomic per level and channel.
__device__ __forceinline__ void v3_emit_strip_wide(const StripParams &P, uint32_t *wave0, uint32_t frame, uint32_t x_strip,
						   int warp, int lane_in)
{
	constexpr int NW = V3::kWarps;
	constexpr int kSteps = (64 + NW - 1) / NW;
	workers_bar<NW>();
#ifdef SCOPE_EMULATE
	const int lane = lane_in;
#else
	int lane; // (read here: a lane id that ptxas traces back to %tid is re-read with S2R inside the loop)
	asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
	(void)lane_in;
#endif
	uint32_t *hist = P.hist + (size_t)frame * P.hist_stride;
	const bool do_hist = P.hist_mask != 0u;
	const int sub = lane >> 3, quad = lane & 7; // level inside the group of four; which four columns
	uint4 *w4 = reinterpret_cast<uint4 *>(wave0);
	uint32_t *img = reinterpret_cast<uint32_t *>(P.wave + (size_t)frame * P.wave_stride) + (P.x_offset + x_strip) + 4 * quad;
	uint4 a[kSteps], b[kSteps];
#pragma unroll
	for (int k = 0; k < kSteps; k++) {
		const int g = warp + k * NW;
		if (g < 64) {
			const int idx = (4 * g + sub) * 8 + quad;
			a[k] = w4[idx];
			b[k] = w4[kWaveWords / 4 + idx];
		} else {
			a[k] = b[k] = make_uint4(0, 0, 0, 0);
		}
	}
#pragma unroll
	for (int k = 0; k < kSteps; k++) {
		const int g = warp + k * NW;
		if (g >= 64)
			break;
		const int v = 4 * g + sub;
		const int idx = v * 8 + quad;
		uint4 *dst = reinterpret_cast<uint4 *>(img + (size_t)(255 - v) * P.out_width);
		const uint32_t any = (a[k].x | a[k].y | a[k].z | a[k].w) | (b[k].x | b[k].y | b[k].z | b[k].w);
		if (!__any_sync(0xFFFFFFFFu, any != 0u)) {
			*dst = make_uint4(0, 0, 0, 0);
			continue;
		}
		w4[idx] = make_uint4(0, 0, 0, 0);
		w4[kWaveWords / 4 + idx] = make_uint4(0, 0, 0, 0);
		uint4 o; // bytes B, G, R, 0 per column
		o.x = __byte_perm(min_u16x2(a[k].x, 0x00FF00FFu), min(b[k].x, 255u), 0x5420);
		o.y = __byte_perm(min_u16x2(a[k].y, 0x00FF00FFu), min(b[k].y, 255u), 0x5420);
		o.z = __byte_perm(min_u16x2(a[k].z, 0x00FF00FFu), min(b[k].z, 255u), 0x5420);
		o.w = __byte_perm(min_u16x2(a[k].w, 0x00FF00FFu), min(b[k].w, 255u), 0x5420);
		*dst = o;
		if (do_hist) {
			const uint32_t bg = (a[k].x + a[k].y) + (a[k].z + a[k].w); // two u16 sums side by side
			const uint32_t sr = (b[k].x + b[k].y) + (b[k].z + b[k].w);
			const bool odd = (quad & 1) != 0, up = (quad & 2) != 0;
			// 1: even lanes collect B|G of the pair, odd lanes R
			const uint32_t v1 = (odd ? sr : bg) + __shfl_xor_sync(0xFFFFFFFFu, odd ? bg : sr, 1);
			// 2: of the even lanes, 0 and 4 collect B, 2 and 6 collect G; the odd lanes go on with R
			const uint32_t lo = v1 & 0xFFFFu, hi = v1 >> 16;
			const uint32_t keep = odd ? v1 : (up ? hi : lo), give = odd ? v1 : (up ? lo : hi);
			const uint32_t v2 = keep + __shfl_xor_sync(0xFFFFFFFFu, give, 2);
			// 3: lanes 0 / 2 / 1 hold B / G / R of the level's 32 columns
			const uint32_t v3 = v2 + __shfl_xor_sync(0xFFFFFFFFu, v2, 4);
			if (quad < 3 && v3 != 0u) // histogram.c:379-395: counts in the order R, G, B
				atomicAdd(hist + v * 4 + (quad == 0 ? 2 : (quad == 2 ? 1 : 0)), v3);
		}
	}
	workers_bar<NW>();
}

// ---------------------------------------------------------------------------
// The ring.  Tile t of EVERY strip goes to stage t mod kStages (a strip restarts at stage 0), so that the consumers'
// loop, unrolled kStages times, knows its stage at compile time: barrier and tile addresses are immediates and
// there is no stage arithmetic per visit.  Each stage's pair of mbarriers keeps its own phase bit on both sides.
// ---------------------------------------------------------------------------
struct V3Phase {
	uint32_t ph[V3::kStages];
};

// producer (one lane): claims chunks of strips and fills the ring; chunk mailbox as in tma_produce
__device__ __forceinline__ void v3_produce(const StripParams &P, const CUtensorMap *map, uint32_t smem_base,
					   volatile uint32_t *chunk_q, uint32_t bar_full, uint32_t bar_empty)
{
	const uint32_t tiles = (P.height + V3::kTileRows - 1) / V3::kTileRows;
	uint32_t phases = 0; // bit s: phase of stage s
	uint32_t qw = 0;
	// frame-affine claiming: this CTA's home frame, how many CTAs share a frame, frames found used up so far
	const uint32_t n_frames = P.items / P.strips;
	uint32_t cur_f = (uint32_t)(((unsigned long long)blockIdx.x * n_frames) / gridDim.x), used_up = 0;
	const uint32_t share = gridDim.x / n_frames + 2u;
	for (;;) {
		uint32_t first, last;
		bool done;
		if (P.frame_affine) {
			// strips of the frame this CTA is on, as long as there are any (guided: 1 / (2 x share) of what the
			// frame has left, at most chunk_items); then the next frame that still has strips.  A frame that was
			// found used up stays used up, so after n_frames of them in a row the batch is done.
			done = true;
			first = last = 0;
			while (used_up < n_frames) {
				const uint32_t seen = *reinterpret_cast<volatile const uint32_t *>(P.chunk_counter + cur_f);
				if (seen < P.strips) {
					const uint32_t want = min(max((P.strips - seen) / (2u * share), 1u), P.chunk_items);
					const uint32_t got = atomicAdd(P.chunk_counter + cur_f, want);
					if (got < P.strips) {
						first = cur_f * P.strips + got;
						last = cur_f * P.strips + min(got + want, P.strips);
						done = false;
						break;
					}
				}
				cur_f = cur_f + 1u == n_frames ? 0u : cur_f + 1u;
				used_up++;
			}
		} else {
			// guided self-scheduling over the whole batch (see tma_produce)
			const uint32_t seen = *reinterpret_cast<volatile const uint32_t *>(P.chunk_counter);
			uint32_t want = seen < P.items ? (P.items - seen) / (2u * gridDim.x) : 1u;
			want = min(max(want, 1u), P.chunk_items);
			first = atomicAdd(P.chunk_counter, want);
			done = first >= P.items;
			last = min(first + want, P.items);
		}
		// announce the chunk (or the end) before its first tile (stage 0) can complete
		mbar_wait(bar_empty, (phases & 1u) ^ 1u);
		chunk_q[2 * (qw % kQueue)] = first;
		chunk_q[2 * (qw % kQueue) + 1] = done ? 0u : last - first;
		qw++;
		if (done) {
			mbar_arrive(bar_full); // wake the consumers with no data
			break;
		}
#ifdef SCOPE_V3_NOLOAD // diagnostic build (never shipped): no TMA at all, the consumers work on whatever the stages hold
		mbar_arrive(bar_full);
		phases ^= 1u;
		continue;
#endif
		// L2 prefetch cursor: runs SCOPE_V3_L2_AHEAD tiles ahead of the loads, inside this chunk
		uint32_t pf_item = first, pf_t = 0;
		auto prefetch_to = [&](uint32_t item, uint32_t t) { // everything up to (item, t + AHEAD)
			const uint32_t goal = (item - first) * tiles + t + (uint32_t)SCOPE_V3_L2_AHEAD;
			while (pf_item < last && (pf_item - first) * tiles + pf_t <= goal) {
				const uint32_t f = pf_item / P.strips, st = pf_item - f * P.strips;
				tma_prefetch_3d(map, (int)(st * kStripPx + P.tma_x0_rgb), (int)(pf_t * V3::kTileRows), (int)f);
				if (++pf_t == tiles) {
					pf_t = 0;
					pf_item++;
				}
			}
		};
		for (uint32_t item = first; item < last; item++) {
			const uint32_t frame = item / P.strips, strip = item - frame * P.strips;
			const int x = (int)(strip * kStripPx);
			uint32_t s = 0;
			for (uint32_t t = 0; t < tiles; t++) {
				if (SCOPE_V3_L2_AHEAD > 0)
					prefetch_to(item, t);
				if (!(item == first && t == 0))
					mbar_wait(bar_empty + 8 * s, ((phases >> s) & 1u) ^ 1u);
				phases ^= 1u << s;
				const uint32_t dst = smem_base + V3::kStageOff + s * V3::kTileBytes;
				mbar_expect_tx(bar_full + 8 * s, V3::kTileBytes);
				tma_load_3d(dst, map, bar_full + 8 * s, x + (int)P.tma_x0_rgb, (int)(t * V3::kTileRows), (int)frame);
				if (++s == V3::kStages)
					s = 0;
			}
		}
	}
}

// ---------------------------------------------------------------------------
// consumer warps
// ---------------------------------------------------------------------------

// wait for the tile in stage S and read this warp's four rows of it
template <int S>
__device__ __forceinline__ void v3_load(const V3Warp &w, V3Phase &ph, uint32_t (&p)[4], bool wait)
{
#ifndef SCOPE_V3_NOLOAD
	if (wait)
		mbar_wait(w.bar_full + 8 * S, ph.ph[S]);
	ph.ph[S] ^= 1;
#endif
#ifdef SCOPE_EMULATE
	emul::ldmatrix<4>(w.rows_base + S * V3::kTileBytes, p);
#else
	asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
		     : "=r"(p[0]), "=r"(p[1]), "=r"(p[2]), "=r"(p[3])
		     : "r"(w.rows_base + S * V3::kTileBytes)
		     : "memory");
#endif
}
// hand stage S back.  The barrier address depends on the loaded data (AND with a zero that only the launch
// parameters know): the arrive cannot be issued before ldmatrix has delivered, i.e. before the stage has been
// read.  Without it the arrive overtook the loads and the producer's next TMA tile landed on rows still to be
// read (profiles/r02/v3_war.md: wrong-tile pixels on flat content, where the consumers run far ahead of the loader).
template <int S>
__device__ __forceinline__ void v3_release(const V3Warp &w, const uint32_t (&p)[4])
{
#ifndef SCOPE_V3_NOLOAD
	const uint32_t dep = (p[0] & p[1] & p[2] & p[3]) & w.zero;
	if (w.lane == 0)
		mbar_arrive((w.bar_empty + 8 * S) | dep);
#endif
}

// one visit: the block of tile t (in `p`, read from stage S during the previous visit) is accumulated while the
// rows of tile t + 1 travel from stage S + 1 to `pn`
template <int CS, int S>
__device__ __forceinline__ void v3_visit(const StripParams &P, const V3Warp &w, V3Phase &ph, uint32_t t, uint32_t tiles,
					 uint32_t n_fast, bool lane_ok, const uint32_t (&p)[4], uint32_t (&pn)[4], V3Pend &pend)
{
	constexpr int S1 = (S + 1) % V3::kStages;
	const bool more = t + 1 < tiles;
	if (more)
		v3_load<S1>(w, ph, pn, true);
	// the previous visit's vectorscope adds have long returned
	v3_resolve(pend);
	v3_pend_clear(pend);
	if (more)
		v3_release<S1>(w, pn); // right away: the loader's round trip is what the three stages have to cover
	const uint32_t m_and = p[0] & p[1] & p[2] & p[3], m_or = p[0] | p[1] | p[2] | p[3];
#ifdef SCOPE_V3_NOP // diagnostic build (never shipped): the ring alone, no accumulation - how fast can tiles be consumed?
	if (m_and == 0x12345678u && m_or == 0x9ABCDEF0u)
		red_shared(w.rows_base, 1u);
	return;
#endif
	// ONE vote finds the ordinary block: every pixel counted (alpha != 0: histogram.c:385-387, waveform.c:246-248)
	// and no lane whose four words are equal (a flat block needs every lane like that)
	const bool special = m_and <= 0x00FFFFFFu || m_and == m_or;
	if (t < n_fast && !__any_sync(0xFFFFFFFFu, special)) {
		v3_block_ordinary<CS>(p, w.k, w.ku, w.kv, w.zero, pend);
	} else {
		// sass-cold{
		bool done = false;
		if (t < n_fast && __all_sync(0xFFFFFFFFu, m_and > 0x00FFFFFFu)) {
			const uint32_t p_lane0 = __shfl_sync(0xFFFFFFFFu, p[0], 0);
			if (__all_sync(0xFFFFFFFFu, ((m_and ^ m_or) | (p[0] ^ p_lane0)) == 0u))
				v3_block_flat<CS>(p[0], w.k, w.ku, w.kv, w.zero, w.lane);
			else
#if SCOPE_V3_RESOLVE_NOW
				v3_block_mixed<CS>(p, m_and == m_or, w);
#else
				v3_block_ordinary<CS>(p, w.k, w.ku, w.kv, w.zero, pend);
#endif
			done = true;
		}
		if (!done) {
			const uint32_t y0 = t * V3::kTileRows + w.y_warp;
			uint32_t rows_ok = 0;
#pragma unroll
			for (int i = 0; i < 4; i++)
				rows_ok |= (y0 + i < P.height ? 1u : 0u) << i;
			v3_block_slow<CS>(p, w.k, w.ku, w.kv, w.zero, lane_ok, rows_ok);
		}
		// sass-cold}
	}
}

// hand stage S back, lean form: the arrive's address register is (shared-memory base | pn[0] & 0) and the barrier's
// offset an immediate.  ONE of ldmatrix's four destination registers is enough for the dependency: the scoreboard
// that the arrive has to wait for belongs to the instruction, not to a register.
template <int S>
__device__ __forceinline__ void v3_release_lean(const V3Warp &w, uint32_t p0)
{
#ifndef SCOPE_V3_NOLOAD
	if (w.lane == 0) {
#ifdef SCOPE_EMULATE
		mbar_arrive((w.base | (p0 & w.zero)) + V3::kBarOff + 8 * (V3::kStages + S));
#else
		asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0+%1];" ::"r"(w.base | (p0 & w.zero)),
			     "n"(V3::kBarOff + 8 * (V3::kStages + S))
			     : "memory");
#endif
	}
#endif
}

// A visit whose block lies completely inside the frame and that has a successor (t + 1 < tiles, t < n_fast): no
// per-visit range checks, and the look at the previous visit's vectorscope adds (v3_resolve) is part of the ONE vote -
// an old value with its top bit set is as rare as a flat block and takes the same way out.
template <int CS, int S>
__device__ __forceinline__ void v3_visit_lean(const V3Warp &w, V3Phase &ph, const uint32_t (&p)[4], uint32_t (&pn)[4],
					      V3Pend &pend)
{
	constexpr int S1 = (S + 1) % V3::kStages;
	v3_load<S1>(w, ph, pn, true);
	v3_release_lean<S1>(w, pn[0]);
	const uint32_t m_and = p[0] & p[1] & p[2] & p[3], m_or = p[0] | p[1] | p[2] | p[3];
#if SCOPE_V3_RESOLVE_NOW
	const uint32_t olds = 0u; // (`pend` is empty throughout: every block settles its own adds)
#else
	const uint32_t olds = pend.old[0] | pend.old[1] | pend.old[2] | pend.old[3];
#endif
#ifdef SCOPE_V3_NOP
	if (m_and == 0x12345678u && m_or == 0x9ABCDEF0u)
		red_shared(w.rows_base, olds);
	return;
#endif
	const bool special = m_and <= 0x00FFFFFFu || m_and == m_or || (olds & 0x80008000u) != 0u;
	if (!__any_sync(0xFFFFFFFFu, special)) {
		v3_block_ordinary<CS>(p, w.k, w.ku, w.kv, w.zero, pend);
	} else {
		// sass-cold{
		v3_resolve(pend);
		v3_pend_clear(pend);
		if (__all_sync(0xFFFFFFFFu, m_and > 0x00FFFFFFu)) {
			const uint32_t p_lane0 = __shfl_sync(0xFFFFFFFFu, p[0], 0);
			if (__all_sync(0xFFFFFFFFu, ((m_and ^ m_or) | (p[0] ^ p_lane0)) == 0u))
				v3_block_flat<CS>(p[0], w.k, w.ku, w.kv, w.zero, w.lane);
			else
#if SCOPE_V3_RESOLVE_NOW
				v3_block_mixed<CS>(p, m_and == m_or, w);
#else
				v3_block_ordinary<CS>(p, w.k, w.ku, w.kv, w.zero, pend);
#endif
		} else {
			v3_block_slow<CS>(p, w.k, w.ku, w.kv, w.zero, true, 0xFu);
		}
		// sass-cold}
	}
}

template <int CS>
__device__ __forceinline__ void v3_consume(const StripParams &P, uint8_t *smem, uint32_t smem_base,
					   volatile uint32_t *chunk_q, uint32_t bar_full, uint32_t bar_empty, int warp, int lane,
					   int tid)
{
	static_assert(V3::kStages == 3 || V3::kStages == 4, "v3_consume is unrolled for three or four stages");
	uint32_t *vs = reinterpret_cast<uint32_t *>(smem + V3::kVsOff);
	uint32_t *wave0 = reinterpret_cast<uint32_t *>(smem + V3::kWaveOff);
	const uint32_t tiles = (P.height + V3::kTileRows - 1) / V3::kTileRows;
	constexpr V3Coef coef = v3_coef<CS>();
	V3Warp w;
	// (the addends of the two sums stay in registers: an IMAD takes one immediate, the multiplier)
#if SCOPE_V3_PARAM_CONSTS
	w.k = v3_consts_from(P.v3c, smem_base, lane);
	w.ku = P.coef.ku;
	w.kv = P.coef.kv;
	(void)coef;
#else
	w.k = v3_consts(smem_base, lane);
	w.ku = coef.ku;
	w.kv = coef.kv;
#endif
	w.zero = P.rt_zero; // 0, known only at run time (see StripParams)
#if SCOPE_V3_SMEM_CONSTS
	{
		// (written by thread 0 before the CTA's first barrier, see the kernel)
		const volatile uint32_t *kc = reinterpret_cast<const volatile uint32_t *>(smem + V3::kConstOff);
		w.ku = kc[0];
		w.kv = kc[1];
		w.zero = kc[2];
		w.k.round = f2_pack(kc[3], kc[4]);
		w.k.k128 = f2_pack(kc[5], kc[5]);
		w.k.exp_hi = kc[6];
		w.k.vs_add = f2_pack(kc[7], kc[8]);
	}
#endif
	w.rows_base = smem_base + V3::kStageOff + (uint32_t)warp * (V3::kRows * kStripPx * 4) + (uint32_t)lane * 16u;
	w.base = smem_base;
	w.bar_full = bar_full;
	w.bar_empty = bar_empty;
	w.y_warp = (uint32_t)warp * V3::kRows;
	w.lane = lane;
	V3Phase ph;
	for (int s = 0; s < V3::kStages; s++)
		ph.ph[s] = 0;
	uint32_t qr = 0;
	uint32_t cur_frame = 0xFFFFFFFFu;
	V3Pend pend;
	v3_pend_clear(pend);

	for (;;) {
		// the chunk id becomes readable once the chunk's first tile (or the end marker) lands in stage 0
		mbar_wait(bar_full, ph.ph[0]);
		uint32_t first = 0, count = 0;
		if (lane == 0) { // (the lane that hands stages back reads the mailbox for the warp, see tma_consume)
			first = chunk_q[2 * (qr % kQueue)];
			count = chunk_q[2 * (qr % kQueue) + 1];
		}
		first = __shfl_sync(0xFFFFFFFFu, first, 0);
		count = __shfl_sync(0xFFFFFFFFu, count, 0);
		qr++;
		if (count == 0u)
			break;
#ifdef SCOPE_V3_NOLOAD
		ph.ph[0] ^= 1;
		if (lane == 0)
			mbar_arrive(bar_empty);
#endif
		const uint32_t last = first + count;
		for (uint32_t item = first; item < last; item++) {
			const uint32_t frame = item / P.strips, strip = item - frame * P.strips;
			if (frame != cur_frame && cur_frame != 0xFFFFFFFFu && !(SCOPE_V3_DIAG & 4))
				v3_flush(P, vs, cur_frame, tid);
			cur_frame = frame;
			const uint32_t x = strip * kStripPx + lane;
			const bool lane_ok = x < P.width;
			const bool strip_full = strip * kStripPx + kStripPx <= P.width;
			// visits [0, n_fast) of this warp lie completely inside the frame
			uint32_t n_fast = 0;
			if (strip_full && P.height >= w.y_warp + V3::kRows)
				n_fast = (P.height - w.y_warp - V3::kRows) / V3::kTileRows + 1u;
			uint32_t pa[4], pb[4], pc[4], pd[4];
			// tile 0 (the chunk announcement already waited for the first item's)
			v3_load<0>(w, ph, pa, item != first);
			v3_release<0>(w, pa);
			uint32_t t = 0;
#if SCOPE_V3_LEAN
			// visits [0, n_lean): inside the frame and with a successor
			const uint32_t n_lean = min(n_fast, tiles - 1u);
#endif
			if (V3::kStages == 3) {
#if SCOPE_V3_LEAN
				for (; t + 3 <= n_lean; t += 3) { // sass-loop-v3
					v3_visit_lean<CS, 0>(w, ph, pa, pb, pend);
					v3_visit_lean<CS, 1>(w, ph, pb, pc, pend);
					v3_visit_lean<CS, 2>(w, ph, pc, pa, pend);
				}
#endif
				for (; t < tiles; t += 3) {
					v3_visit<CS, 0>(P, w, ph, t, tiles, n_fast, lane_ok, pa, pb, pend);
					if (t + 1 < tiles)
						v3_visit<CS, 1>(P, w, ph, t + 1, tiles, n_fast, lane_ok, pb, pc, pend);
					if (t + 2 < tiles)
						v3_visit<CS, 2>(P, w, ph, t + 2, tiles, n_fast, lane_ok, pc, pa, pend);
				}
			} else {
#if SCOPE_V3_LEAN
				for (; t + 4 <= n_lean; t += 4) { // sass-loop-v3
					v3_visit_lean<CS, 0>(w, ph, pa, pb, pend);
					v3_visit_lean<CS, 1>(w, ph, pb, pc, pend);
					v3_visit_lean<CS, 2>(w, ph, pc, pd, pend);
					v3_visit_lean<CS, 3 % V3::kStages>(w, ph, pd, pa, pend);
				}
#endif
				for (; t < tiles; t += 4) {
					v3_visit<CS, 0>(P, w, ph, t, tiles, n_fast, lane_ok, pa, pb, pend);
					if (t + 1 < tiles)
						v3_visit<CS, 1>(P, w, ph, t + 1, tiles, n_fast, lane_ok, pb, pc, pend);
					if (t + 2 < tiles)
						v3_visit<CS, 2>(P, w, ph, t + 2, tiles, n_fast, lane_ok, pc, pd, pend);
					if (t + 3 < tiles)
						v3_visit<CS, 3 % V3::kStages>(P, w, ph, t + 3, tiles, n_fast, lane_ok, pd, pa, pend);
				}
			}
			v3_resolve(pend);
			v3_pend_clear(pend);
#if SCOPE_V3_DIAG & 3
			if (!(SCOPE_V3_DIAG & 2)) {
				workers_bar<V3::kWarps>();
				workers_bar<V3::kWarps>();
			}
			continue;
#endif
#if SCOPE_V3_WIDE_EMIT
			// (from the launch parameters alone, every time: nothing to keep in a register across the strip)
			// 16-byte stores into the waveform rows, packed u16 sums of eight columns
			const bool wide_ok = ((P.x_offset | P.out_width) & 3u) == 0u && (P.wave_stride & 15u) == 0u &&
					     (reinterpret_cast<uintptr_t>(P.wave) & 15u) == 0u && P.height <= 8191u;
			if (wide_ok && strip * kStripPx + kStripPx <= P.width)
				v3_emit_strip_wide(P, wave0, frame, strip * kStripPx, warp, lane);
			else
#endif
				v3_emit_strip(P, wave0, frame, x, lane_ok, warp, lane);
		}
	}
	if (cur_frame != 0xFFFFFFFFu)
		v3_flush(P, vs, cur_frame, tid);
}

template <int CS>
__global__ void __launch_bounds__(V3::kThreads, 1)
	scope_fused_kernel_v3(const __grid_constant__ StripParams P, const __grid_constant__ CUtensorMap map_rgb)
{
	SCOPE_DYNAMIC_SMEM(smem);
	volatile uint32_t *chunk_q = reinterpret_cast<volatile uint32_t *>(smem + V3::kQueueOff);
	const uint32_t smem_base = smem_u32(smem);
	const uint32_t bar_full = smem_base + V3::kBarOff;
	const uint32_t bar_empty = bar_full + V3::kStages * 8;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

	// zero the bins (all threads), set up the ring
	for (int i = tid; i < (V3::kVsOff + V3::kVsTableWords * 4) / 16; i += V3::kThreads)
		reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
#ifdef SCOPE_V3_NOLOAD
	for (int i = tid; i < V3::kStages * V3::kTileBytes / 4; i += V3::kThreads)
		reinterpret_cast<uint32_t *>(smem + V3::kStageOff)[i] = ((uint32_t)i * 2654435761u + blockIdx.x * 40503u) | 0xFF000000u;
#endif
	if (tid == 0) {
#if SCOPE_V3_SMEM_CONSTS
		{
			volatile uint32_t *kc = reinterpret_cast<volatile uint32_t *>(smem + V3::kConstOff);
#if SCOPE_V3_PARAM_CONSTS
			const V3Consts k0 = v3_consts_from(P.v3c, smem_base, 0);
			kc[0] = P.coef.ku;
			kc[1] = P.coef.kv;
#else
			const V3Consts k0 = v3_consts(smem_base, 0);
			kc[0] = v3_coef<CS>().ku;
			kc[1] = v3_coef<CS>().kv;
#endif
			uint32_t lo, hi;
			kc[2] = P.rt_zero;
			f2_unpack(k0.round, lo, hi);
			kc[3] = lo;
			kc[4] = hi;
			f2_unpack(k0.k128, lo, hi);
			kc[5] = lo;
			kc[6] = k0.exp_hi;
			f2_unpack(k0.vs_add, lo, hi);
			kc[7] = lo;
			kc[8] = hi;
		}
#endif
		for (int s = 0; s < V3::kStages; s++) {
			mbar_init(bar_full + 8 * s, 1);
			mbar_init(bar_empty + 8 * s, V3::kWarps);
		}
#ifndef SCOPE_EMULATE
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
	}
	__syncthreads();
	if (warp == V3::kWarps) {
		if (lane == 0)
			v3_produce(P, &map_rgb, smem_base, chunk_q, bar_full, bar_empty);
		return;
	}
	v3_consume<CS>(P, smem, smem_base, chunk_q, bar_full, bar_empty, warp, lane, tid);
}

#ifndef SCOPE_EMULATE
// test hook: the v3 transform for all 2^24 colours (index r<<16|g<<8|b), U | V << 8 per colour
template <int CS>
__global__ void __launch_bounds__(256) uv_table_kernel_v3(uint32_t *out)
{
	const uint32_t i = (blockIdx.x * 256 + threadIdx.x) * 2;
#pragma unroll
	for (uint32_t j = i; j < i + 2; j++)
		out[j] = v3_uv_bytes<CS>(j, 0u);
}
#endif

} // namespace scope
