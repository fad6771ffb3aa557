// tools/simt/emul_main.cpp — runs the strip kernels of csrc/scope_kernels.cuh on the SIMT emulator
// (cuda_emul.h; TEST INFRASTRUCTURE, no GPU).  Built by tools/simt/build.sh into libscope_emul*.so, one
// library per set of kernel build flags; tests/test_kernel_emulation.py calls emul_run through ctypes and
// compares the outputs with the CPU oracle.
//
// emul_run does what launch_strip (csrc/scope_ffi.cu) does on the host - StripParams, the tensor maps,
// chunk size, grid - then runs the CTAs as coroutines under a seeded random scheduler, with TMA loads
// landing late and out of order.
#include "scope_kernels.cuh"

#include <string>

namespace emul {
World *W = nullptr;
}
using namespace scope;

extern "C" {
struct EmulRequest {
	const uint8_t *rgb, *yuv;
	uint32_t linesize, width, height, n_frames;
	uint64_t frame_stride;
	int32_t colorspace, surface, src, vscope;
	uint32_t bins_mask, hist_mask, wave_mask;
	uint32_t *hist;       // [n][1024], zeroed by the caller
	uint8_t *wave;        // [n][256][out_width][4]
	uint32_t *vs_acc;     // [n][65536], zeroed by the caller
	uint32_t *wave_pairs; // partial mode: [2][256][out_width], zeroed by the caller
	uint32_t out_width, x_offset, partial;
	int32_t kernel;       // 0 = TMA tile kernel, 1 = plain-load kernel, 2 = TMA row-group kernel
	int32_t ctas;         // CTAs to run concurrently (the "grid")
	uint32_t seed;
	int32_t tma_land_percent; // chance per scheduler round that one pending TMA load of a CTA lands
	int64_t steps;            // out: scheduler steps taken
	char error[256];          // out
};
}

namespace {

struct Launch {
	StripParams P;
	CUtensorMap map_rgb, map_yuv;
	int kernel;
	int src, vs, surf, colorspace;
};

template <int SRC, bool VS, bool SURF> void call_kernel(const Launch &l)
{
	if (l.kernel == 0) {
#if SCOPE_IMMCOEF
		if constexpr (!SURF && (VS || SRC == SRC_YUV)) { // what kernel_entry (csrc/scope_ffi.cu) picks in these builds
			if (l.colorspace == 1)
				scope_strip_kernel_tma<SRC, VS, SURF, 1>(l.P, l.map_rgb, l.map_yuv);
			else
				scope_strip_kernel_tma<SRC, VS, SURF, 2>(l.P, l.map_rgb, l.map_yuv);
			return;
		}
#endif
		scope_strip_kernel_tma<SRC, VS, SURF>(l.P, l.map_rgb, l.map_yuv);
	}
	else if (l.kernel == 2)
		scope_strip_kernel_tmag<SRC, VS, SURF>(l.P, l.map_rgb, l.map_yuv);
	else
		scope_strip_kernel_ldg<SRC, VS, SURF>(l.P);
}

template <bool SURF> bool dispatch2(const Launch &l)
{
	if (l.src == SRC_NONE && l.vs)
		call_kernel<SRC_NONE, true, SURF>(l);
	else if (l.src == SRC_RGB && l.vs)
		call_kernel<SRC_RGB, true, SURF>(l);
	else if (l.src == SRC_RGB && !l.vs)
		call_kernel<SRC_RGB, false, SURF>(l);
	else if (l.src == SRC_YUV && l.vs)
		call_kernel<SRC_YUV, true, SURF>(l);
	else if (l.src == SRC_YUV && !l.vs)
		call_kernel<SRC_YUV, false, SURF>(l);
	else
		return false;
	return true;
}

const Launch *g_launch;

void thread_entry()
{
	const Launch &l = *g_launch;
	if (l.kernel == 3) { // scope_fused_kernel_v3: the headline combination's own kernel
		if (l.colorspace == 1)
			scope_fused_kernel_v3<1>(l.P, l.map_rgb);
		else
			scope_fused_kernel_v3<2>(l.P, l.map_rgb);
	} else if (l.surf)
		dispatch2<true>(l);
	else
		dispatch2<false>(l);
	emul::W->cur->done = true;
	emul::progress();
	swapcontext(&emul::W->cur->ctx, &emul::W->sched);
}

template <int SRC, bool VS, bool SURF> int tile_rows_total() { return SmemLayout<SRC, VS, SURF, true>::kTileRows; }
template <int SRC, bool VS, bool SURF> int tma_warps_total() { return SmemLayout<SRC, VS, SURF, true>::kWarps; }
template <bool SURF> int tma_warps_for(int src, bool vs)
{
	if (src == SRC_NONE)
		return tma_warps_total<SRC_NONE, true, SURF>();
	if (src == SRC_RGB)
		return vs ? tma_warps_total<SRC_RGB, true, SURF>() : tma_warps_total<SRC_RGB, false, SURF>();
	return vs ? tma_warps_total<SRC_YUV, true, SURF>() : tma_warps_total<SRC_YUV, false, SURF>();
}
template <bool SURF> int tile_rows_for(int src, bool vs)
{
	if (src == SRC_NONE)
		return tile_rows_total<SRC_NONE, true, SURF>();
	if (src == SRC_RGB)
		return vs ? tile_rows_total<SRC_RGB, true, SURF>() : tile_rows_total<SRC_RGB, false, SURF>();
	return vs ? tile_rows_total<SRC_YUV, true, SURF>() : tile_rows_total<SRC_YUV, false, SURF>();
}

template <int SRC, bool VS, bool SURF> int smem_total(int kernel)
{
	return kernel == 1 ? SmemLayout<SRC, VS, SURF, false>::kTotal : SmemLayout<SRC, VS, SURF, true>::kTotal;
}
template <bool SURF> int smem_for(int src, bool vs, int kernel)
{
	if (src == SRC_NONE)
		return smem_total<SRC_NONE, true, SURF>(kernel);
	if (src == SRC_RGB)
		return vs ? smem_total<SRC_RGB, true, SURF>(kernel) : smem_total<SRC_RGB, false, SURF>(kernel);
	return vs ? smem_total<SRC_YUV, true, SURF>(kernel) : smem_total<SRC_YUV, false, SURF>(kernel);
}

void make_map(CUtensorMap &m, const uint8_t *base, uint32_t width, uint32_t linesize, uint32_t height, uint32_t n,
	      uint64_t frame_stride, int tile_rows)
{
	m.base = base;
	m.dims[0] = width;
	m.dims[1] = height;
	m.dims[2] = n;
	m.strides[0] = linesize;
	m.strides[1] = n > 1 ? frame_stride : (((uint64_t)linesize * height + 15) & ~(uint64_t)15);
	m.box[0] = kStripPx;
	m.box[1] = (uint32_t)tile_rows;
	m.box[2] = 1;
}

} // namespace

extern "C" int emul_run(EmulRequest *rq)
{
	rq->error[0] = 0;
	rq->steps = 0;
	if (rq->src == SRC_NONE && !rq->vscope)
		return 0;
	Launch l{};
	StripParams &P = l.P;
	P.rgb = rq->rgb;
	P.yuv = rq->yuv;
	P.frame_stride = rq->frame_stride;
	P.linesize = rq->linesize;
	P.width = rq->width;
	P.height = rq->height;
	P.n_frames = rq->n_frames;
	P.strips = (rq->width + kStripPx - 1) / kStripPx;
	P.items = P.strips * rq->n_frames;
	P.bins_mask = rq->bins_mask;
	P.hist_mask = rq->hist_mask;
	P.wave_mask = rq->wave_mask;
	P.x_offset = rq->x_offset;
	P.out_width = rq->out_width;
	P.partial = rq->partial;
	P.hist = rq->hist;
	P.hist_stride = 1024;
	P.wave = rq->wave;
	P.wave_stride = (unsigned long long)256 * rq->out_width * 4;
	P.wave_pairs = rq->wave_pairs;
	P.vscope_acc = rq->vs_acc;
	P.vscope_stride = 65536;
	P.coef = coef_for(rq->colorspace);
	v3_param_consts(P.v3c);
	l.kernel = rq->kernel;
	l.src = rq->src;
	l.vs = rq->vscope != 0;
	l.surf = rq->surface != 0;
	l.colorspace = rq->colorspace;

	const bool need_rgb = !l.surf || l.src == SRC_RGB;
	const bool need_yuv = l.surf && (l.src == SRC_YUV || l.vs);
	if (rq->kernel == 3 && !(l.src == SRC_RGB && l.vs && !l.surf && P.bins_mask == 7u && P.wave_mask == 7u &&
				 (P.hist_mask == 7u || P.hist_mask == 0u) && P.partial == 0u)) {
		snprintf(rq->error, sizeof rq->error, "kernel 3 serves fused RGB bins + vectorscope only");
		return 1;
	}
	const int tile_rows = rq->kernel == 3 ? V3::kTileRows
			      : l.surf      ? tile_rows_for<true>(l.src, l.vs)
					    : tile_rows_for<false>(l.src, l.vs);
	if (need_rgb)
		make_map(l.map_rgb, rq->rgb, rq->width, rq->linesize, rq->height, rq->n_frames, rq->frame_stride, tile_rows);
	if (need_yuv)
		make_map(l.map_yuv, rq->yuv, rq->width, rq->linesize, rq->height, rq->n_frames, rq->frame_stride, tile_rows);

	uint32_t grid = (uint32_t)std::max(1, rq->ctas);
	if (grid > P.items)
		grid = P.items;
	static std::vector<uint32_t> chunk_counters;
	chunk_counters.assign(std::max<size_t>(1, rq->n_frames), 0u);
	unsigned threads;
	if (rq->kernel == 1) {
		P.items_per_cta = (P.items + grid - 1) / grid;
		grid = (P.items + P.items_per_cta - 1) / P.items_per_cta;
		threads = kLdgWarps * 32;
	} else {
		uint32_t ch = P.items / (grid * 6u);
		P.chunk_items = ch < 1u ? 1u : (ch > (uint32_t)kMaxChunkItems ? (uint32_t)kMaxChunkItems : ch);
		P.chunk_counter = chunk_counters.data();
		P.frame_affine = rq->kernel == 3 && SCOPE_V3_FRAME_AFFINE ? 1u : 0u; // (seeds alternate it off below)
		if (rq->kernel == 3 && (rq->seed & 4u))
			P.frame_affine = 0u;
		threads = (rq->kernel == 3   ? V3::kWarps
			   : rq->kernel == 2 ? kGroupWarps
					     : (l.surf ? tma_warps_for<true>(l.src, l.vs) : tma_warps_for<false>(l.src, l.vs))) * 32 + 32;
	}
	const int smem = rq->kernel == 3 ? V3::kTotal
			 : l.surf       ? smem_for<true>(l.src, l.vs, rq->kernel)
					: smem_for<false>(l.src, l.vs, rq->kernel);
	if (smem > (int)emul::kSmemBytes) {
		snprintf(rq->error, sizeof rq->error, "kernel needs %d bytes of shared memory: does not fit an SM", smem);
		return 1;
	}

	emul::World world;
	emul::W = &world;
	world.rng.seed(rq->seed);
	world.grid_dim = grid;
	world.ctas.resize(grid);
	world.threads.resize((size_t)grid * threads);
	g_launch = &l;
	constexpr size_t kStack = 256 * 1024;
	for (uint32_t c = 0; c < grid; c++) {
		emul::Cta &cta = world.ctas[c];
		cta.smem.assign((size_t)smem, 0xA5); // garbage: the kernels must zero what they use
		cta.block_idx = c;
		cta.block_dim = threads;
		cta.warps.resize(threads / 32);
		for (unsigned t = 0; t < threads; t++) {
			emul::Thread &th = world.threads[(size_t)c * threads + t];
			th.cta = &cta;
			th.tid = t;
			th.stack = malloc(kStack);
			getcontext(&th.ctx);
			th.ctx.uc_stack.ss_sp = th.stack;
			th.ctx.uc_stack.ss_size = kStack;
			th.ctx.uc_link = nullptr;
			makecontext(&th.ctx, thread_entry, 0);
		}
	}

	std::vector<size_t> order(world.threads.size());
	for (size_t i = 0; i < order.size(); i++)
		order[i] = i;
	size_t alive = order.size();
	std::vector<int> sleep(order.size() / 32 + 1, 0); // per warp: scheduler rounds it still sits out
	const long stall_limit = 60L * (long)order.size();
	while (alive > 0 && !world.error) {
		// TMA completions: late, and in any order
		for (auto &cta : world.ctas)
			if (!cta.tma.empty() && (int)(world.rng() % 100) < rq->tma_land_percent) {
				emul::tma_land(cta, world.rng() % cta.tma.size());
				emul::progress();
			}
		std::shuffle(order.begin(), order.end(), world.rng);
		// skew: now and then a whole warp (or a single lane) is not scheduled for a while, so that warps drift
		// apart by whole tiles - the situations in which a protocol error shows
		for (auto &z : sleep) {
			if (z > 0)
				z--;
			else if (world.rng() % 97 == 0)
				z = (int)(world.rng() % 64);
		}
		for (size_t k = 0; k < order.size() && !world.error; k++) {
			emul::Thread &th = world.threads[order[k]];
			if (th.done)
				continue;
			if (sleep[order[k] / 32] > 0 && world.steps - world.progress_mark < stall_limit / 2)
				continue;
			world.cur = &th;
			world.steps++;
			swapcontext(&world.sched, &th.ctx);
			if (th.done)
				alive--;
		}
		if (world.steps - world.progress_mark > stall_limit) {
			// nothing moved for a long time: a load that has not landed yet is not a deadlock
			bool landed = false;
			for (auto &cta : world.ctas)
				if (!cta.tma.empty()) {
					emul::tma_land(cta, world.rng() % cta.tma.size());
					emul::progress();
					landed = true;
				}
			if (!landed)
				world.error = "deadlock: no progress";
		}
	}
	for (auto &th : world.threads)
		free(th.stack);
	rq->steps = world.steps;
	emul::W = nullptr;
	if (world.error) {
		snprintf(rq->error, sizeof rq->error, "%s", world.error);
		return 1;
	}
	for (auto &cta : world.ctas)
		if (!cta.tma.empty()) {
			snprintf(rq->error, sizeof rq->error, "TMA loads still in flight at exit");
			return 1;
		}
	return 0;
}

extern "C" const char *emul_build_flags()
{
	static std::string s = std::string("warps=") + std::to_string(kTmaWarps) + " fused_warps=" + std::to_string(SCOPE_FUSED_WARPS) + " tile_rows=" + std::to_string(kTileRows) +
			       " straight=" + std::to_string(SCOPE_STRAIGHT) + " rawflat=" + std::to_string(SCOPE_RAWFLAT) +
			       " wide_fused=" + std::to_string(SCOPE_WIDE_FUSED) + " immcoef=" + std::to_string(SCOPE_IMMCOEF) + " ballot=" + std::to_string(SCOPE_BALLOT) + " deep_ring=" + std::to_string(SCOPE_DEEP_RING) + " pipeline=" + std::to_string(SCOPE_PIPELINE) +
			       " faddr=" + std::to_string(SCOPE_FADDR) + " ldsm=" + std::to_string(SCOPE_LDSM) +
			       " defer=" + std::to_string(SCOPE_DEFER) + " fast_emit=" + std::to_string(SCOPE_FAST_EMIT);
	return s.c_str();
}
