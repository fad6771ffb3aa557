// tools/simt/cuda_emul.h — a small SIMT emulator for the strip kernels (TEST INFRASTRUCTURE; no GPU needed).
//
// csrc/scope_kernels.cuh is compiled for the HOST with -DSCOPE_EMULATE -include this file.  Every CUDA thread
// of a CTA becomes a coroutine (ucontext); a seeded scheduler picks the next runnable one at random, so the
// interleaving of lanes, warps and CTAs is far more adversarial than on the hardware.  What is modelled:
//   * warp collectives (shfl, vote, redux, syncwarp, ldmatrix) as rendezvous of the warp's 32 lanes,
//   * named CTA barriers (bar.sync id, n) and __syncthreads,
//   * mbarriers with phases, pending arrivals and pending transaction bytes,
//   * cp.async.bulk.tensor (TMA tile loads) as DEFERRED copies that complete at random later points and
//     out of order, zero-filling outside the tensor, then complete_tx on the tile's mbarrier,
//   * shared / global atomics (plain read-modify-write: coroutines never preempt each other),
//   * the arithmetic helpers (prmt, fma on denormal bit patterns, mul.hi).
// What is NOT modelled: timing, bank conflicts, memory-ordering subtleties below the mbarrier level.
// tests/test_kernel_emulation.py runs every kernel variant through this against the CPU oracle.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <random>
#include <vector>
#include <ucontext.h>

#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__

using std::max;
using std::min;

struct uint4 {
	uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3e {
	unsigned x, y, z;
};

// what cuTensorMapEncodeTiled describes, for the emulated TMA: a 3-D tensor of 4-byte elements
struct CUtensorMap {
	const uint8_t *base;
	uint64_t dims[3];    // elements: x (pixels), y (rows), z (frames)
	uint64_t strides[2]; // bytes: row pitch, frame stride
	uint32_t box[3];     // elements
};

namespace emul {

constexpr uint32_t kSmemBase = 0x400; // what smem_u32 returns for the first byte of dynamic shared memory
constexpr size_t kSmemBytes = 232448;

struct MBar {
	uint32_t count = 0, phase = 0;
	int pending = 0;
	long tx = 0;
	void maybe_complete()
	{
		if (pending == 0 && tx == 0) {
			phase++;
			pending = (int)count;
		}
	}
};

struct PendingTma {
	uint32_t dst, bar;
	const CUtensorMap *map;
	int x, y, z;
};

struct Rendezvous {
	int gen = 0, arrived = 0;
};

struct Warp {
	Rendezvous rv;
	uint32_t vals[2][32];
	uint32_t res[2][32][4];
};

struct Cta {
	std::vector<uint8_t> smem;
	std::map<uint32_t, MBar> mbar;
	std::vector<PendingTma> tma;
	std::map<int, Rendezvous> bars;
	std::vector<Warp> warps;
	unsigned block_idx = 0, block_dim = 0;
};

struct Thread {
	ucontext_t ctx;
	void *stack = nullptr;
	Cta *cta = nullptr;
	unsigned tid = 0;
	bool done = false;
};

struct World {
	std::vector<Cta> ctas;
	std::vector<Thread> threads;
	ucontext_t sched;
	Thread *cur = nullptr;
	std::mt19937 rng;
	unsigned grid_dim = 0;
	long steps = 0, progress_mark = 0;
	const char *error = nullptr;
};

extern World *W;

inline void fail(const char *msg)
{
	if (!W->error)
		W->error = msg;
	// unwind this coroutine for good: mark done and go back to the scheduler
	W->cur->done = true;
	swapcontext(&W->cur->ctx, &W->sched);
}

inline void yield()
{
	swapcontext(&W->cur->ctx, &W->sched);
}

inline void progress() { W->progress_mark = W->steps; }

inline uint8_t *smem_ptr(uint32_t addr, size_t bytes = 4)
{
	if (addr < kSmemBase || addr - kSmemBase + bytes > W->cur->cta->smem.size())
		fail("shared-memory address out of range");
	return W->cur->cta->smem.data() + (addr - kSmemBase);
}

// ---- rendezvous helpers ----
template <class F> inline void rendezvous(Rendezvous &rv, int n, F on_complete)
{
	const int g = rv.gen;
	if (++rv.arrived == n) {
		on_complete();
		rv.arrived = 0;
		rv.gen++;
		progress();
	} else {
		while (rv.gen == g)
			yield();
	}
}

inline Warp &my_warp() { return W->cur->cta->warps[W->cur->tid >> 5]; }
inline int my_lane() { return (int)(W->cur->tid & 31); }

enum Op { OP_SHFL, OP_ALL, OP_ANY, OP_ADD, OP_SYNC, OP_BALLOT };

// full-mask warp collective on one 32-bit value per lane (src_lane only for OP_SHFL)
inline uint32_t collective(Op op, uint32_t v, int src_lane = 0)
{
	Warp &w = my_warp();
	const int lane = my_lane(), b = w.rv.gen & 1;
	w.vals[b][lane] = v;
	rendezvous(w.rv, 32, [&] {
		uint32_t r = 0;
		if (op == OP_ALL) {
			r = 1;
			for (int i = 0; i < 32; i++)
				r &= w.vals[b][i] ? 1u : 0u;
		} else if (op == OP_ANY) {
			for (int i = 0; i < 32; i++)
				r |= w.vals[b][i] ? 1u : 0u;
		} else if (op == OP_ADD) {
			for (int i = 0; i < 32; i++)
				r += w.vals[b][i];
		} else if (op == OP_SHFL) {
			r = w.vals[b][src_lane & 31];
		} else if (op == OP_BALLOT) {
			for (int i = 0; i < 32; i++)
				r |= (w.vals[b][i] ? 1u : 0u) << i;
		}
		for (int i = 0; i < 32; i++)
			w.res[b][i][0] = r;
	});
	return w.res[b][lane][0];
}

// butterfly exchange: lane i receives the value of lane i ^ mask (full mask, uniform `mask`)
inline uint32_t shfl_xor(uint32_t v, int mask)
{
	Warp &w = my_warp();
	const int lane = my_lane(), b = w.rv.gen & 1;
	w.vals[b][lane] = v;
	rendezvous(w.rv, 32, [&] {
		for (int i = 0; i < 32; i++)
			w.res[b][i][0] = w.vals[b][(i ^ mask) & 31];
	});
	return w.res[b][lane][0];
}

// ldmatrix m8n8 b16: matrix j's eight 16-byte rows come from the addresses of lanes 8j..8j+7; thread T gets
// word T % 4 of row T / 4 of every matrix
template <int NMAT> inline void ldmatrix(uint32_t addr, uint32_t (&out)[NMAT])
{
	Warp &w = my_warp();
	const int lane = my_lane(), b = w.rv.gen & 1;
	w.vals[b][lane] = addr;
	rendezvous(w.rv, 32, [&] {
		for (int t = 0; t < 32; t++)
			for (int j = 0; j < NMAT; j++) {
				const uint32_t row = w.vals[b][8 * j + t / 4];
				if (row & 15u)
					fail("ldmatrix row address not 16-byte aligned");
				uint32_t v;
				memcpy(&v, smem_ptr(row + 4u * (t % 4)), 4);
				w.res[b][t][j] = v;
			}
	});
	for (int j = 0; j < NMAT; j++)
		out[j] = w.res[b][lane][j];
}

inline void bar_sync(int id, int n)
{
	rendezvous(W->cur->cta->bars[id], n, [] {});
}

// ---- mbarrier ----
inline MBar &mbar(uint32_t addr)
{
	(void)smem_ptr(addr, 8);
	return W->cur->cta->mbar[addr];
}
inline void mbar_init(uint32_t addr, uint32_t count)
{
	MBar &m = mbar(addr);
	m = MBar();
	m.count = count;
	m.pending = (int)count;
}
inline void mbar_arrive(uint32_t addr)
{
	MBar &m = mbar(addr);
	if (m.count == 0)
		fail("arrive on an uninitialised mbarrier");
	if (m.pending == 0)
		fail("arrival on an mbarrier whose phase already has all its arrivals");
	m.pending--;
	m.maybe_complete();
	progress();
}
inline void mbar_expect_tx(uint32_t addr, uint32_t bytes)
{
	mbar(addr).tx += bytes;
	mbar_arrive(addr);
}
inline bool mbar_test(uint32_t addr, uint32_t parity)
{
	MBar &m = mbar(addr);
	if (m.count == 0)
		fail("wait on an uninitialised mbarrier");
	return (m.phase & 1u) != (parity & 1u);
}
inline void mbar_wait(uint32_t addr, uint32_t parity)
{
	while (!mbar_test(addr, parity))
		yield();
}

// ---- TMA ----
inline void tma_issue(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y, int z)
{
	if (dst & 127u)
		fail("TMA destination not 128-byte aligned");
	if (((uintptr_t)map->base & 15u) || ((uintptr_t)(map->base + (size_t)x * 4) & 15u))
		fail("TMA box does not start on a 16-byte boundary in global memory"); // sm_100a traps
	W->cur->cta->tma.push_back(PendingTma{dst, bar, map, x, y, z});
	progress();
}
// scheduler side: one pending load of `c` lands
inline void tma_land(Cta &c, size_t k)
{
	PendingTma t = c.tma[k];
	c.tma.erase(c.tma.begin() + (long)k);
	const CUtensorMap &m = *t.map;
	const size_t bytes = (size_t)m.box[0] * m.box[1] * m.box[2] * 4;
	uint8_t *dst = c.smem.data() + (t.dst - kSmemBase);
	if (t.dst - kSmemBase + bytes > c.smem.size()) {
		W->error = "TMA tile does not fit the shared-memory window";
		return;
	}
	for (uint32_t bz = 0; bz < m.box[2]; bz++)
		for (uint32_t by = 0; by < m.box[1]; by++)
			for (uint32_t bx = 0; bx < m.box[0]; bx++) {
				const int64_t gx = t.x + (int64_t)bx, gy = t.y + (int64_t)by, gz = t.z + (int64_t)bz;
				uint32_t v = 0; // outside the tensor: zero fill
				if (gx >= 0 && gy >= 0 && gz >= 0 && (uint64_t)gx < m.dims[0] && (uint64_t)gy < m.dims[1] &&
				    (uint64_t)gz < m.dims[2])
					memcpy(&v, m.base + (size_t)gz * m.strides[1] + (size_t)gy * m.strides[0] + (size_t)gx * 4, 4);
				memcpy(dst + (((size_t)bz * m.box[1] + by) * m.box[0] + bx) * 4, &v, 4);
			}
	MBar &mb = c.mbar[t.bar];
	mb.tx -= (long)bytes;
	mb.maybe_complete();
}

// ---- arithmetic ----
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
	const uint64_t src = ((uint64_t)b << 32) | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; i++) {
		const uint32_t n = (sel >> (4 * i)) & 0xF;
		uint32_t byte = (uint32_t)(src >> (8 * (n & 7))) & 0xFF;
		if (n & 8)
			byte = (byte & 0x80) ? 0xFF : 0x00;
		r |= byte << (8 * i);
	}
	return r;
}
inline uint32_t fma_bits(uint32_t a, float b, uint32_t c)
{
	float fa, fc;
	memcpy(&fa, &a, 4);
	memcpy(&fc, &c, 4);
	const float d = std::fmaf(fa, b, fc); // exact on denormals, like fma.rn.f32 without .ftz
	uint32_t r;
	memcpy(&r, &d, 4);
	return r;
}

} // namespace emul

// ---- the CUDA spellings the kernels use ----
#define threadIdx (dim3e{emul::W->cur->tid, 0, 0})
#define blockIdx (dim3e{emul::W->cur->cta->block_idx, 0, 0})
#define gridDim (dim3e{emul::W->grid_dim, 1, 1})
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) { return emul::prmt(a, b, s); }
static inline uint32_t __shfl_sync(uint32_t, uint32_t v, int lane) { return emul::collective(emul::OP_SHFL, v, lane); }
static inline uint32_t __shfl_xor_sync(uint32_t, uint32_t v, int mask) { return emul::shfl_xor(v, mask); }
static inline bool __all_sync(uint32_t, bool p) { return emul::collective(emul::OP_ALL, p) != 0; }
static inline bool __any_sync(uint32_t, bool p) { return emul::collective(emul::OP_ANY, p) != 0; }
static inline uint32_t __reduce_add_sync(uint32_t, uint32_t v) { return emul::collective(emul::OP_ADD, v); }
static inline uint32_t __ballot_sync(uint32_t, bool p) { return emul::collective(emul::OP_BALLOT, p); }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline void __syncwarp() { emul::collective(emul::OP_SYNC, 0); }
static inline void __syncthreads() { emul::bar_sync(0, (int)emul::W->cur->cta->block_dim); }
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v)
{
	const uint32_t old = *p;
	*p = old + v;
	emul::progress();
	return old;
}
