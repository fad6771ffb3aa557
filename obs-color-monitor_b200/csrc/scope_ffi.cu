// scope_ffi.cu — host side of libscope_b200.so: the C-ABI declared in include/scope_ffi.h.
//
// Replaces, behind the reference's cm_surface_cb_t seam (src/common.h:32), the work of
// his_surface_cb / wvs_surface_cb / vss_surface_cb (src/histogram.c:432-450,
// src/waveform.c:272-289, src/vectorscope.c:248-265) and of the RGB->YUV shader pass
// (src/common.c:170-221, data/common.effect).  There is no CPU fallback in here.
#include <nvtx3/nvToolsExt.h> // header-only; ranges named like the reference's profile scopes (common.c:10-21)
#include "scope_kernels.cuh"
#include "scope_peer_reduce.cuh"
#include "../../include/scope_ffi.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

using namespace scope;

namespace {

thread_local std::string g_create_error;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
				    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
				    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct RingSlot {
	cudaStream_t stream = nullptr;
	cudaEvent_t done = nullptr;
	bool in_flight = false;
	// device staging for the host entry points
	uint8_t *d_in = nullptr;
	size_t d_in_bytes = 0;
	uint32_t *d_hist = nullptr;    // 1024
	uint32_t *d_hist_max = nullptr; // 4
	uint8_t *d_wave = nullptr;
	uint8_t *d_wave_disp = nullptr;
	size_t d_wave_bytes = 0;
	uint8_t *d_vscope = nullptr;      // 65536
	uint8_t *d_vscope_disp = nullptr; // 65536
	uint32_t *d_vs_acc = nullptr;     // 65536 u32: this slot's own vectorscope accumulators (the slots run on
					  // different streams, so they must not share the context's scratch)
	// pinned input staging handed out by scope_ring_input() (the "stagesurface" of this slot)
	uint8_t *h_in = nullptr;
	size_t h_in_bytes = 0;
	// pinned result staging
	uint8_t *h_res = nullptr;
	size_t h_res_bytes = 0;
	// what the in-flight submission asked for
	scope_params params{};
	uint32_t width = 0, height = 0;
};

} // namespace

struct scope_ctx {
	int device = 0;
	int sm_count = 0;
	std::string err;
	uint64_t launches = 0;
	PFN_encodeTiled encode = nullptr;
	bool default_tma = true; // which loader launch_strip prefers when both are possible
	// scratch of the DEVICE entry points: u32 vectorscope accumulators [n][65536] and a histogram for
	// callers that only want hist_max.  One set per context, used by launches on whatever stream the
	// caller names: `scratch_free` is recorded behind the last kernel that reads them and every new
	// user waits for it on its own stream first, so two streams never interleave on the scratch.
	uint32_t *d_vs_acc = nullptr;
	size_t vs_acc_frames = 0;
	uint32_t *d_hist_scratch = nullptr;
	size_t hist_scratch_frames = 0;
	cudaEvent_t scratch_free = nullptr;
	bool scratch_used = false;
	// per-kernel launch facts that never change (cudaFuncSetAttribute done, CTAs per SM)
	struct KernelFacts {
		const void *fn;
		int ctas_per_sm;
	};
	std::vector<KernelFacts> kernel_facts;
	// work counters of the dynamically scheduled launches (one slot per launch, round robin)
	static constexpr uint32_t kCounterPool = 16384;
	uint32_t *d_counters = nullptr; // pool of kCounterPool counters, handed out in slices, round robin
	uint32_t counter_next = 0;
	RingSlot ring[SCOPE_RING_SLOTS];
	std::mutex mu;
	// optional per-launch timing of the strip kernel (scope_profile_*)
	bool profiling = false;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pool;
};

namespace {

int fail(scope_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess)
{
	char buf[512];
	if (e != cudaSuccess)
		snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
	else
		snprintf(buf, sizeof buf, "%s", what);
	if (ctx)
		ctx->err = buf;
	else
		g_create_error = buf;
	return code;
}

#define CU_TRY(ctx, call)                                              \
	do {                                                           \
		cudaError_t e_ = (call);                               \
		if (e_ != cudaSuccess)                                 \
			return fail(ctx, SCOPE_ERR_CUDA, #call, e_);   \
	} while (0)

// components -> (source plane, channel mask) exactly like histogram.c:367-377 / waveform.c:228-238
void decode_components(uint32_t comp, int &src, uint32_t &mask)
{
	if (comp & 0x07u)
		src = SRC_RGB;
	else if (comp & 0x70u)
		src = SRC_YUV;
	else
		src = SRC_NONE;
	mask = ((comp & 0x11u) ? 1u : 0u) | ((comp & 0x22u) ? 2u : 0u) | ((comp & 0x44u) ? 4u : 0u);
	if (src == SRC_NONE)
		mask = 0;
}

typedef void (*TmaKernel)(const StripParams, const CUtensorMap, const CUtensorMap);
typedef void (*LdgKernel)(const StripParams);

typedef void (*V3Kernel)(const StripParams, const CUtensorMap);

struct KernelChoice {
	V3Kernel v3 = nullptr; // the headline combination's own kernel (scope_fused_v3.cuh)
	TmaKernel tma = nullptr;
	LdgKernel ldg = nullptr;
	int smem = 0, threads = 0;
	int tile_rows = kTileRows; // rows of the TMA box this kernel expects (SCOPE_WIDE_FUSED: per kernel family)
};

// A/B builds only (-DSCOPE_EXPERIMENT, variants_tmp/): the kernels that were measured and not adopted
// (scope_kernels_experiments.cuh) and the environment switches that select them.  The shipped
// library carries neither: it has ONE kernel family and reads no environment variable.
#ifdef SCOPE_EXPERIMENT
bool use_group_kernel()
{
	const char *e = getenv("SCOPE_KERNEL");
	if (e && !strcmp(e, "group"))
		return !SCOPE_WIDE_FUSED; // (the row-group kernel is not offered in those builds)
	if (e && !strcmp(e, "tile"))
		return false;
	return SCOPE_GROUP_WARPS > 0 && !SCOPE_WIDE_FUSED;
}
#endif

template <int SRC, bool VS, bool SURF>
void kernel_entry(bool tma, int colorspace, KernelChoice &k)
{
#ifdef SCOPE_EXPERIMENT
	if (tma && use_group_kernel()) {
		k.tma = scope_strip_kernel_tmag<SRC, VS, SURF>;
		k.smem = SmemLayout<SRC, VS, SURF, true>::kTotal;
		k.threads = kGroupWarps * 32 + 32;
		return;
	}
#endif
	if (tma) {
		k.tile_rows = SmemLayout<SRC, VS, SURF, true>::kTileRows;
		// the kernels that evaluate the transform exist once per colour space, coefficients as immediates
		if constexpr (SCOPE_IMMCOEF && !SURF && (VS || SRC == SRC_YUV))
			k.tma = colorspace == 1 ? scope_strip_kernel_tma<SRC, VS, SURF, 1> : scope_strip_kernel_tma<SRC, VS, SURF, 2>;
		else
			k.tma = scope_strip_kernel_tma<SRC, VS, SURF>;
		k.smem = SmemLayout<SRC, VS, SURF, true>::kTotal;
		k.threads = SmemLayout<SRC, VS, SURF, true>::kWarps * 32 + 32;
	} else {
		k.ldg = scope_strip_kernel_ldg<SRC, VS, SURF>;
		k.smem = SmemLayout<SRC, VS, SURF, false>::kTotal;
		k.threads = kLdgWarps * 32;
	}
}

template <bool SURF>
bool pick_kernel2(int src, bool vs, bool tma, int colorspace, KernelChoice &k)
{
	if (src == SRC_NONE && vs)
		kernel_entry<SRC_NONE, true, SURF>(tma, colorspace, k);
	else if (src == SRC_RGB && vs)
		kernel_entry<SRC_RGB, true, SURF>(tma, colorspace, k);
	else if (src == SRC_RGB && !vs)
		kernel_entry<SRC_RGB, false, SURF>(tma, colorspace, k);
	else if (src == SRC_YUV && vs)
		kernel_entry<SRC_YUV, true, SURF>(tma, colorspace, k);
	else if (src == SRC_YUV && !vs)
		kernel_entry<SRC_YUV, false, SURF>(tma, colorspace, k);
	else
		return false;
	return true;
}

bool pick_kernel(int src, bool vs, bool surface, bool tma, int colorspace, KernelChoice &k)
{
#ifdef SCOPE_EXPERIMENT
	// experiment kept for A/B runs (profiles/ubench_r01.md): SCOPE_SPLIT=1 selects the
	// warp-specialised kernel for the headline combination; it measured no faster
	const char *split = getenv("SCOPE_SPLIT");
	if (tma && vs && src == SRC_RGB && split && split[0] == '1') {
		if (surface) {
			k.tma = scope_strip_kernel_split<true>;
			k.smem = SmemLayout<SRC_RGB, true, true, true>::kTotal;
			k.tile_rows = SmemLayout<SRC_RGB, true, true, true>::kTileRows;
		} else {
			k.tma = scope_strip_kernel_split<false>;
			k.smem = SmemLayout<SRC_RGB, true, false, true>::kTotal;
			k.tile_rows = SmemLayout<SRC_RGB, true, false, true>::kTileRows;
		}
		k.threads = (kSplitVsWarps + kSplitBinWarps) * 32 + 32;
		return true;
	}
#endif
	return surface ? pick_kernel2<true>(src, vs, tma, colorspace, k) : pick_kernel2<false>(src, vs, tma, colorspace, k);
}

int make_map(scope_ctx *ctx, CUtensorMap *map, const uint8_t *base16, uint32_t x_extent_px, uint32_t linesize,
	     uint32_t height, uint32_t n_frames, size_t frame_stride, int tile_rows)
{
	cuuint64_t dims[3] = {x_extent_px, height, n_frames};
	cuuint64_t strides[2] = {linesize, n_frames > 1 ? (cuuint64_t)frame_stride
							  : (cuuint64_t)(((size_t)linesize * height + 15) & ~(size_t)15)};
	cuuint32_t box[3] = {(cuuint32_t)kStripPx, (cuuint32_t)tile_rows, 1};
	cuuint32_t estr[3] = {1, 1, 1};
	CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)base16, dims, strides, box, estr,
				 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
				 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		char buf[128];
		snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
		return fail(ctx, SCOPE_ERR_CUDA, buf);
	}
	return SCOPE_OK;
}

struct Request {
	const uint8_t *rgb, *yuv; // device
	uint32_t linesize, width, height, n_frames;
	size_t frame_stride;
	int colorspace;
	bool surface;
	// which column bins / vectorscope this launch serves
	int src;
	uint32_t bins_mask, hist_mask, wave_mask;
	bool vscope;
	// outputs
	uint32_t *hist;
	size_t hist_stride;
	uint8_t *wave;
	size_t wave_stride;
	uint32_t *wave_pairs;
	uint32_t x_offset, out_width;
	int partial; // StripParams::partial: 0 final u8, 1 add into wave_pairs, 2 store into wave_pairs
	uint8_t *const *wave_copies = nullptr; // further images that get the final waveform columns (peer stores)
	uint32_t n_wave_copies = 0;
	uint32_t *vs_acc;
	size_t vs_stride;
	// target_scale (width / height above are the SCALED size) and the transform mode: both plain-load kernel only
	uint32_t scale_x = 1, scale_y = 1;
	bool strict = false;
};

int launch_strip(scope_ctx *ctx, const Request &rq, cudaStream_t stream)
{
	if (rq.src == SRC_NONE && !rq.vscope)
		return SCOPE_OK;
	const bool need_rgb = !rq.surface || rq.src == SRC_RGB;
	const bool need_yuv = rq.surface && (rq.src == SRC_YUV || rq.vscope);
	if ((need_rgb && !rq.rgb) || (need_yuv && !rq.yuv))
		return fail(ctx, SCOPE_ERR_INVALID, "a plane the request needs is NULL");
	if (rq.height > 65535u)
		return fail(ctx, SCOPE_ERR_UNSUPPORTED, "height > 65535 (u16 column bins)");
	if ((rq.linesize & 3u) || ((uintptr_t)rq.rgb & 3u) || ((uintptr_t)rq.yuv & 3u) || (rq.frame_stride & 3u))
		return fail(ctx, SCOPE_ERR_UNSUPPORTED, "planes must be 4-byte aligned with a 4-byte multiple pitch");

	StripParams P{};
	P.rgb = rq.rgb;
	P.yuv = rq.yuv;
	P.frame_stride = rq.frame_stride;
	P.linesize = rq.linesize;
	P.width = rq.width;
	P.height = rq.height;
	P.n_frames = rq.n_frames;
	P.strips = (rq.width + kStripPx - 1) / kStripPx;
	P.items = P.strips * rq.n_frames;
	P.bins_mask = rq.bins_mask;
	P.hist_mask = rq.hist_mask;
	P.wave_mask = rq.wave_mask;
	P.x_offset = rq.x_offset;
	P.out_width = rq.out_width;
	P.partial = (uint32_t)rq.partial;
	P.n_wave_copies = rq.n_wave_copies;
	for (uint32_t c = 0; c < rq.n_wave_copies; c++)
		P.wave_copies[c] = rq.wave_copies[c];
	P.hist = rq.hist;
	P.hist_stride = rq.hist_stride;
	P.wave = rq.wave;
	P.wave_stride = rq.wave_stride;
	P.wave_pairs = rq.wave_pairs;
	P.vscope_acc = rq.vs_acc;
	P.vscope_stride = rq.vs_stride;
	P.coef = coef_for(rq.colorspace);
	P.scale_x = rq.scale_x;
	P.scale_y = rq.scale_y;
	v3_param_consts(P.v3c);
	P.xform_strict = rq.strict ? 1u : 0u;
	P.colorspace = rq.colorspace;

	// Loader choice.  TMA can describe the planes if base pointers, pitch and frame stride are
	// multiples of 16 bytes; everything else (e.g. an ROI crop at an odd column, common.c:272-282)
	// takes the direct loader: sm_100a traps with "Illegal instruction" inside cp.async.bulk.tensor
	// when a box's first pixel is not 16-byte aligned in global memory, even if the tensor map's base
	// is (measured in round 1, gpurun_out/final2/x0_sanitizer.log).
	P.tma_x0_rgb = (uint32_t)(((uintptr_t)rq.rgb & 15u) / 4u);
	P.tma_x0_yuv = (uint32_t)(((uintptr_t)rq.yuv & 15u) / 4u);
	bool allow_x0 = false;
#ifdef SCOPE_EXPERIMENT
	const char *x0_env = getenv("SCOPE_TMA_X0");
	allow_x0 = x0_env && x0_env[0] == '1';
#endif
	// a scaled pass reads every scale-th pixel (a TMA box would have to start scale / 2 pixels into its row: not
	// 16-byte aligned), and the fp32-strict transform only exists in the plain-load kernel
	const bool plain_only = rq.scale_x > 1 || rq.scale_y > 1 || rq.strict;
	const bool tma_ok = !plain_only && ctx->encode != nullptr && (rq.linesize % 16u) == 0 &&
			    (rq.n_frames == 1 || (rq.frame_stride % 16u) == 0) &&
			    (allow_x0 || ((!need_rgb || P.tma_x0_rgb == 0) && (!need_yuv || P.tma_x0_yuv == 0)));
	bool use_tma = tma_ok && ctx->default_tma;
#ifdef SCOPE_EXPERIMENT
	const char *loader = getenv("SCOPE_LOADER"); // A/B builds: SCOPE_LOADER=tma|ldg
	if (loader && !strcmp(loader, "tma"))
		use_tma = tma_ok;
	else if (loader && !strcmp(loader, "ldg"))
		use_tma = false;
#endif

	KernelChoice k;
	// fused mode, RGB column bins for all three channels, final u8 waveform, vectorscope: scope_fused_kernel_v3
	const bool v3 = SCOPE_V3 && use_tma && !rq.surface && rq.src == SRC_RGB && rq.vscope && rq.bins_mask == 7u &&
			rq.wave_mask == 7u && (rq.hist_mask == 7u || rq.hist_mask == 0u) && rq.partial == 0 &&
			rq.n_wave_copies == 0 && rq.wave != nullptr;
	if (v3) {
		k.v3 = rq.colorspace == 1 ? scope_fused_kernel_v3<1> : scope_fused_kernel_v3<2>;
		k.smem = V3::kTotal;
		k.threads = V3::kThreads;
		k.tile_rows = V3::kTileRows;
	} else if (!pick_kernel(rq.src, rq.vscope, rq.surface, use_tma, rq.colorspace, k))
		return fail(ctx, SCOPE_ERR_INVALID, "no kernel for this scope combination");

	CUtensorMap map_rgb, map_yuv;
	memset(&map_rgb, 0, sizeof map_rgb);
	memset(&map_yuv, 0, sizeof map_yuv);
	if (use_tma) {
		if (need_rgb) {
			int r = make_map(ctx, &map_rgb, rq.rgb - 4u * P.tma_x0_rgb, rq.width + P.tma_x0_rgb, rq.linesize,
					 rq.height, rq.n_frames, rq.frame_stride, k.tile_rows);
			if (r)
				return r;
		}
		if (need_yuv) {
			int r = make_map(ctx, &map_yuv, rq.yuv - 4u * P.tma_x0_yuv, rq.width + P.tma_x0_yuv, rq.linesize,
					 rq.height, rq.n_frames, rq.frame_stride, k.tile_rows);
			if (r)
				return r;
		}
	}

	// the opt-in to large dynamic shared memory and the occupancy of a kernel never change: asked once
	const void *fn = k.v3 ? (const void *)k.v3 : use_tma ? (const void *)k.tma : (const void *)k.ldg;
	int ctas_per_sm = 0;
	for (const auto &f : ctx->kernel_facts)
		if (f.fn == fn)
			ctas_per_sm = f.ctas_per_sm;
	if (ctas_per_sm == 0) {
		CU_TRY(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem));
		CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, fn, k.threads, k.smem));
		if (ctas_per_sm < 1)
			return fail(ctx, SCOPE_ERR_CUDA, "kernel does not fit on an SM");
		ctx->kernel_facts.push_back({fn, ctas_per_sm});
	}
	uint32_t grid = (uint32_t)(ctx->sm_count * ctas_per_sm);
	if (grid > P.items)
		grid = P.items;
	if (use_tma) {
		// chunks of strips are claimed at run time from a zeroed counter (guided
		// self-scheduling in tma_produce): at most chunk_items strips each - big enough to keep
		// vectorscope flushes rare, small enough that every CTA gets >= ~6 of them - and
		// shrinking towards the end of the batch
		uint32_t ch = P.items / (grid * 6u);
		P.chunk_items = ch < 1u ? 1u : (ch > (uint32_t)kMaxChunkItems ? (uint32_t)kMaxChunkItems : ch);
		if (!ctx->d_counters)
			CU_TRY(ctx, cudaMalloc(&ctx->d_counters, ctx->kCounterPool * sizeof(uint32_t)));
		// the headline kernel claims strips frame by frame (one counter per frame): a CTA then changes frames - and
		// flushes its vectorscope table - a few times per launch instead of after nearly every chunk
		const uint32_t n_counters = (v3 && SCOPE_V3_FRAME_AFFINE && rq.n_frames <= ctx->kCounterPool / 4u) ? rq.n_frames : 1u;
		P.frame_affine = n_counters == rq.n_frames && v3 && SCOPE_V3_FRAME_AFFINE ? 1u : 0u;
		if (ctx->counter_next + n_counters > ctx->kCounterPool)
			ctx->counter_next = 0; // (a slice is reused only after kCounterPool counters' worth of later launches)
		P.chunk_counter = ctx->d_counters + ctx->counter_next;
		ctx->counter_next += n_counters;
		CU_TRY(ctx, cudaMemsetAsync(P.chunk_counter, 0, n_counters * sizeof(uint32_t), stream));
		P.items_per_cta = 0;
	} else {
		P.items_per_cta = (P.items + grid - 1) / grid;
		grid = (P.items + P.items_per_cta - 1) / P.items_per_cta;
	}

	std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
	if (ctx->profiling) {
		if (!ctx->prof_pool.empty()) {
			ev = ctx->prof_pool.back();
			ctx->prof_pool.pop_back();
		} else {
			CU_TRY(ctx, cudaEventCreate(&ev.first));
			CU_TRY(ctx, cudaEventCreate(&ev.second));
		}
		CU_TRY(ctx, cudaEventRecord(ev.first, stream));
	}
	if (k.v3)
		k.v3<<<grid, k.threads, k.smem, stream>>>(P, map_rgb);
	else if (use_tma)
		k.tma<<<grid, k.threads, k.smem, stream>>>(P, map_rgb, map_yuv);
	else
		k.ldg<<<grid, k.threads, k.smem, stream>>>(P);
	CU_TRY(ctx, cudaGetLastError());
	if (ctx->profiling) {
		CU_TRY(ctx, cudaEventRecord(ev.second, stream));
		ctx->prof_events.push_back(ev);
	}
	ctx->launches++;
	return SCOPE_OK;
}

int ensure_vs_acc(scope_ctx *ctx, size_t frames)
{
	if (ctx->vs_acc_frames >= frames)
		return SCOPE_OK;
	if (ctx->d_vs_acc) // (cudaFree waits for the device, so an earlier user is done before the memory goes away)
		cudaFree(ctx->d_vs_acc);
	ctx->d_vs_acc = nullptr;
	ctx->vs_acc_frames = 0;
	cudaError_t e = cudaMalloc(&ctx->d_vs_acc, frames * 65536 * sizeof(uint32_t));
	if (e != cudaSuccess)
		return fail(ctx, SCOPE_ERR_NOMEM, "cudaMalloc(vectorscope accumulators)", e);
	ctx->vs_acc_frames = frames;
	return SCOPE_OK;
}

int ensure_hist_scratch(scope_ctx *ctx, size_t frames)
{
	if (ctx->hist_scratch_frames >= frames)
		return SCOPE_OK;
	if (ctx->d_hist_scratch)
		cudaFree(ctx->d_hist_scratch);
	ctx->d_hist_scratch = nullptr;
	ctx->hist_scratch_frames = 0;
	cudaError_t e = cudaMalloc(&ctx->d_hist_scratch, frames * 1024 * sizeof(uint32_t));
	if (e != cudaSuccess)
		return fail(ctx, SCOPE_ERR_NOMEM, "cudaMalloc(histogram scratch)", e);
	ctx->hist_scratch_frames = frames;
	return SCOPE_OK;
}

// The whole device-side pass for n frames (used by both the device and the host entry points).
// slot_vs_acc: a ring slot's own vectorscope accumulators (host entry points, n_frames == 1), or NULL:
// then the context's shared scratch is used, ordered against its previous user through `scratch_free`.
int run_device(scope_ctx *ctx, const scope_params *pr, const scope_surface *s, uint32_t n_frames,
	       size_t frame_stride, const scope_out_device *out, cudaStream_t stream, uint32_t *slot_vs_acc = nullptr,
	       bool rows_prescaled = false)
{
	if (!pr || !s || !out)
		return fail(ctx, SCOPE_ERR_INVALID, "NULL argument");
	if (s->width == 0 || s->height == 0 || n_frames == 0)
		return fail(ctx, SCOPE_ERR_INVALID, "empty surface");
	if (s->linesize < s->width * 4u)
		return fail(ctx, SCOPE_ERR_INVALID, "linesize < width*4");
	const bool surface = pr->mode == SCOPE_MODE_SURFACE;
	// target_scale: the scopes see the scaled surface (common.c:249-250); scale_y_device = 1 when the caller has
	// already dropped the rows (the host entry points do that while copying)
	const uint32_t scale = pr->target_scale > 1 ? pr->target_scale : 1u;
	if (scale > 128u)
		return fail(ctx, SCOPE_ERR_INVALID, "target_scale > 128 (common.c:88-90)");
	const uint32_t sw = s->width / scale, sh = rows_prescaled ? s->height : s->height / scale;
	if (sw == 0 || sh == 0)
		return fail(ctx, SCOPE_ERR_INVALID, "surface smaller than target_scale");
	if (pr->xform != SCOPE_XFORM_EXACT && pr->xform != SCOPE_XFORM_FP32_STRICT)
		return fail(ctx, SCOPE_ERR_INVALID, "unknown xform");
	const bool want_hist = (pr->scopes & SCOPE_HIST) && (out->hist_counts || out->hist_max);
	const bool want_wave = (pr->scopes & SCOPE_WAVE) && (out->wave || out->wave_display);
	const bool want_vs = (pr->scopes & SCOPE_VSCOPE) && (out->vscope || out->vscope_display);
	if ((pr->scopes & SCOPE_WAVE) && out->wave_display && !out->wave)
		return fail(ctx, SCOPE_ERR_INVALID, "wave_display needs wave");

	int hsrc = SRC_NONE, wsrc = SRC_NONE;
	uint32_t hmask = 0, wmask = 0;
	if (want_hist)
		decode_components(pr->hist_components, hsrc, hmask);
	if (want_wave)
		decode_components(pr->wave_components, wsrc, wmask);

	uint32_t *hist = out->hist_counts;
	size_t hist_stride = 1024;
	uint32_t *vs_acc = slot_vs_acc;
	bool shared_scratch = false;
	if (want_hist && !hist) {
		int r = ensure_hist_scratch(ctx, n_frames);
		if (r)
			return r;
		hist = ctx->d_hist_scratch;
		shared_scratch = true;
	}
	if (want_vs && !vs_acc) {
		int r = ensure_vs_acc(ctx, n_frames);
		if (r)
			return r;
		vs_acc = ctx->d_vs_acc;
		shared_scratch = true;
	}
	if (shared_scratch) {
		if (!ctx->scratch_free)
			CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->scratch_free, cudaEventDisableTiming));
		if (ctx->scratch_used) // the previous user (possibly on another stream) must be done with it
			CU_TRY(ctx, cudaStreamWaitEvent(stream, ctx->scratch_free, 0));
	}
	if (want_hist)
		CU_TRY(ctx, cudaMemsetAsync(hist, 0, (size_t)n_frames * 1024 * sizeof(uint32_t), stream));
	if (want_wave && wsrc == SRC_NONE) // components select no plane: the reference leaves zeros
		CU_TRY(ctx, cudaMemsetAsync(out->wave, 0, (size_t)n_frames * scope_wave_bytes(sw), stream));
	if (want_vs)
		CU_TRY(ctx, cudaMemsetAsync(vs_acc, 0, (size_t)n_frames * 65536 * sizeof(uint32_t), stream));

	Request rq{};
	rq.rgb = s->rgb_data;
	rq.yuv = surface ? s->yuv_data : nullptr;
	rq.linesize = s->linesize;
	rq.width = sw;
	rq.height = sh;
	rq.scale_x = scale;
	rq.scale_y = rows_prescaled ? 1u : scale;
	rq.strict = !surface && pr->xform == SCOPE_XFORM_FP32_STRICT;
	rq.n_frames = n_frames;
	rq.frame_stride = frame_stride;
	rq.colorspace = s->colorspace;
	rq.surface = surface;
	rq.hist = hist;
	rq.hist_stride = hist_stride;
	rq.wave = out->wave;
	rq.wave_stride = scope_wave_bytes(sw);
	rq.wave_pairs = nullptr;
	rq.x_offset = 0;
	rq.out_width = sw;
	rq.partial = 0;
	rq.vs_acc = vs_acc;
	rq.vs_stride = 65536;

	// the reference's named profile scopes (histogram.c:10-19, waveform.c:8-18, vectorscope.c:10-20), as NVTX
	// ranges around the launches that do their work; one fused launch carries all the names it serves
	struct ScopeRanges {
		int n = 0;
		ScopeRanges(bool h, bool w, bool v)
		{
			if (h)
				nvtxRangePushA("draw_histogram"), n++;
			if (w)
				nvtxRangePushA("draw_waveform"), n++;
			if (v)
				nvtxRangePushA("draw_vectorscope"), n++;
		}
		~ScopeRanges()
		{
			while (n-- > 0)
				nvtxRangePop();
		}
	} ranges(want_hist, want_wave, want_vs);

	// One launch when histogram and waveform read the same plane (or only one of them is
	// on); otherwise the histogram gets its own launch with the vectorscope riding on the
	// waveform's.  Mirrors the ROI fan-out (roi.c:329-341): same surface, every scope.
	if (hsrc != SRC_NONE && wsrc != SRC_NONE && hsrc != wsrc) {
		Request a = rq;
		a.src = wsrc;
		a.bins_mask = wmask;
		a.wave_mask = wmask;
		a.hist_mask = 0;
		a.vscope = want_vs;
		int r = launch_strip(ctx, a, stream);
		if (r)
			return r;
		Request b = rq;
		b.src = hsrc;
		b.bins_mask = hmask;
		b.hist_mask = hmask;
		b.wave_mask = 0;
		b.vscope = false;
		r = launch_strip(ctx, b, stream);
		if (r)
			return r;
	} else {
		Request a = rq;
		a.src = hsrc != SRC_NONE ? hsrc : wsrc;
		a.hist_mask = hmask;
		a.wave_mask = wmask;
		a.bins_mask = hmask | wmask;
		a.vscope = want_vs;
		int r = launch_strip(ctx, a, stream);
		if (r)
			return r;
	}

	if (want_vs) {
		const float k = pr->vscope_intensity > 0 ? (float)pr->vscope_intensity : 1.0f;
		dim3 grid(65536 / 1024, n_frames);
		vscope_finalize_kernel<<<grid, 256, 0, stream>>>(vs_acc, 65536, out->vscope,
								  pr->vscope_intensity > 0 ? out->vscope_display : nullptr,
								  65536, k);
		CU_TRY(ctx, cudaGetLastError());
		ctx->launches++;
	}
	// (no component bit selects a plane: the reference returns before its level pass, histogram.c:366-373,
	// and hi_max keeps its previous contents - so nothing is written here either)
	if (want_hist && out->hist_max && (pr->hist_components & 0x77u)) {
		hist_max_kernel<<<n_frames, 256, 0, stream>>>(hist, hist_stride, out->hist_max, pr->hist_components,
							       sw, sh, pr->level_fixed_value,
							       pr->level_ratio_value);
		CU_TRY(ctx, cudaGetLastError());
		ctx->launches++;
	}
	if (want_wave && out->wave_display && pr->wave_intensity > 0) {
		const size_t words = (size_t)n_frames * scope_wave_bytes(sw) / 4;
		wave_display_kernel<<<(unsigned)((words + 255) / 256), 256, 0, stream>>>(out->wave, out->wave_display, words,
										       (float)pr->wave_intensity);
		CU_TRY(ctx, cudaGetLastError());
		ctx->launches++;
	}
	if (shared_scratch) {
		CU_TRY(ctx, cudaEventRecord(ctx->scratch_free, stream));
		ctx->scratch_used = true;
	}
	return SCOPE_OK;
}

// histogram post-pass on the host: exactly histogram.c:404-417 (same libm logf as the reference)
void hist_to_float(const scope_params *pr, const uint32_t *counts, uint32_t *hi_max, float *out)
{
	if (pr->logscale) {
		memcpy(out, counts, sizeof(uint32_t) * 1024);
		for (int j = 0, mask = 0x44; j < 3; j++, mask >>= 1) {
			if (!(pr->hist_components & (uint32_t)mask))
				continue;
			const float sc = 1.0f / logf((float)(hi_max[j] + 1));
			for (int i = 0; i < 256; i++)
				out[i * 4 + j] = counts[i * 4 + j] ? logf((float)(counts[i * 4 + j] + 1)) * sc : 0;
			hi_max[j] = 1;
		}
	} else {
		for (int i = 0; i < 1024; i++)
			out[i] = (float)counts[i];
	}
}

int ensure_slot(scope_ctx *ctx, RingSlot &sl, size_t in_bytes, uint32_t width)
{
	if (!sl.stream) {
		CU_TRY(ctx, cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
		CU_TRY(ctx, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
		CU_TRY(ctx, cudaMalloc(&sl.d_hist, 1024 * 4));
		CU_TRY(ctx, cudaMalloc(&sl.d_hist_max, 16));
		CU_TRY(ctx, cudaMalloc(&sl.d_vscope, 65536));
		CU_TRY(ctx, cudaMalloc(&sl.d_vscope_disp, 65536));
		CU_TRY(ctx, cudaMalloc(&sl.d_vs_acc, 65536 * sizeof(uint32_t)));
	}
	if (sl.d_in_bytes < in_bytes) {
		if (sl.d_in)
			cudaFree(sl.d_in);
		sl.d_in = nullptr;
		sl.d_in_bytes = 0;
		cudaError_t e = cudaMalloc(&sl.d_in, in_bytes);
		if (e != cudaSuccess)
			return fail(ctx, SCOPE_ERR_NOMEM, "cudaMalloc(input staging)", e);
		sl.d_in_bytes = in_bytes;
	}
	const size_t wb = scope_wave_bytes(width);
	if (sl.d_wave_bytes < wb) {
		if (sl.d_wave)
			cudaFree(sl.d_wave);
		if (sl.d_wave_disp)
			cudaFree(sl.d_wave_disp);
		sl.d_wave = sl.d_wave_disp = nullptr;
		sl.d_wave_bytes = 0;
		cudaError_t e = cudaMalloc(&sl.d_wave, wb);
		if (e == cudaSuccess)
			e = cudaMalloc(&sl.d_wave_disp, wb);
		if (e != cudaSuccess)
			return fail(ctx, SCOPE_ERR_NOMEM, "cudaMalloc(waveform staging)", e);
		sl.d_wave_bytes = wb;
	}
	const size_t rb = 4096 + 16 + 65536 * 2 + wb * 2;
	if (sl.h_res_bytes < rb) {
		if (sl.h_res)
			cudaFreeHost(sl.h_res);
		sl.h_res = nullptr;
		sl.h_res_bytes = 0;
		cudaError_t e = cudaHostAlloc(&sl.h_res, rb, cudaHostAllocDefault);
		if (e != cudaSuccess)
			return fail(ctx, SCOPE_ERR_NOMEM, "cudaHostAlloc(result staging)", e);
		sl.h_res_bytes = rb;
	}
	return SCOPE_OK;
}

int submit_host(scope_ctx *ctx, int slot, const scope_params *pr, const scope_surface *s)
{
	if (!pr || !s)
		return fail(ctx, SCOPE_ERR_INVALID, "NULL argument");
	if (slot < 0 || slot >= SCOPE_RING_SLOTS)
		return fail(ctx, SCOPE_ERR_INVALID, "bad ring slot");
	RingSlot &sl = ctx->ring[slot];
	if (sl.in_flight)
		return fail(ctx, SCOPE_ERR_BUSY, "ring slot still in flight");
	if (s->width == 0 || s->height == 0)
		return fail(ctx, SCOPE_ERR_INVALID, "empty surface");
	if (s->linesize < s->width * 4u)
		return fail(ctx, SCOPE_ERR_INVALID, "linesize < width*4");
	const bool surface = pr->mode == SCOPE_MODE_SURFACE;
	// same early-outs as the reference callbacks (histogram.c:436-441, waveform.c:276-281,
	// vectorscope.c:252-253): a missing plane means "do nothing, keep the previous result"
	int hsrc, wsrc;
	uint32_t m;
	decode_components(pr->hist_components, hsrc, m);
	decode_components(pr->wave_components, wsrc, m);
	const bool hist_on = pr->scopes & SCOPE_HIST, wave_on = pr->scopes & SCOPE_WAVE, vs_on = pr->scopes & SCOPE_VSCOPE;
	const bool need_rgb = !surface || (hist_on && hsrc == SRC_RGB) || (wave_on && wsrc == SRC_RGB);
	const bool need_yuv = surface && ((hist_on && hsrc == SRC_YUV) || (wave_on && wsrc == SRC_YUV) || vs_on);
	if ((need_rgb && !s->rgb_data) || (need_yuv && !s->yuv_data))
		return fail(ctx, SCOPE_ERR_INVALID, "a plane the request needs is NULL");

	// device copy: rows packed at a 16-byte-multiple pitch so the TMA path applies.  With target_scale only every
	// scale-th row is needed (row y of the scaled surface = source row y s + s / 2): the copy takes just those, so
	// the bus carries 1 / s of the frame; the columns are picked by the kernel
	const uint32_t scale = pr->target_scale > 1 ? pr->target_scale : 1u;
	if (scale > 128u || s->width / scale == 0 || s->height / scale == 0)
		return fail(ctx, SCOPE_ERR_INVALID, "target_scale out of range for this surface");
	const uint32_t rows = s->height / scale, out_w = s->width / scale;
	const size_t row0 = (size_t)(scale / 2u) * s->linesize, src_pitch = (size_t)s->linesize * scale;
	const uint32_t pitch = (s->width * 4u + 15u) & ~15u;
	const size_t plane_bytes = (size_t)pitch * rows;
	int r = ensure_slot(ctx, sl, plane_bytes * 2, out_w);
	if (r)
		return r;
	uint8_t *d_rgb = sl.d_in, *d_yuv = sl.d_in + plane_bytes;
	nvtxRangePushA("stage_surface"); // common.c:316-320: the copy of the frame towards the consumer
	if (need_rgb)
		CU_TRY(ctx, cudaMemcpy2DAsync(d_rgb, pitch, s->rgb_data + row0, src_pitch, (size_t)s->width * 4, rows,
					      cudaMemcpyHostToDevice, sl.stream));
	if (need_yuv)
		CU_TRY(ctx, cudaMemcpy2DAsync(d_yuv, pitch, s->yuv_data + row0, src_pitch, (size_t)s->width * 4, rows,
					      cudaMemcpyHostToDevice, sl.stream));
	nvtxRangePop();

	scope_surface ds = *s;
	ds.rgb_data = need_rgb ? d_rgb : nullptr;
	ds.yuv_data = need_yuv ? d_yuv : nullptr;
	ds.linesize = pitch;
	ds.height = rows; // (rows already dropped; run_device scales the width only)
	scope_out_device od{};
	od.hist_counts = sl.d_hist;
	od.hist_max = sl.d_hist_max;
	od.wave = sl.d_wave;
	od.wave_display = pr->wave_intensity > 0 ? sl.d_wave_disp : nullptr;
	od.vscope = sl.d_vscope;
	od.vscope_display = pr->vscope_intensity > 0 ? sl.d_vscope_disp : nullptr;
	r = run_device(ctx, pr, &ds, 1, plane_bytes, &od, sl.stream, sl.d_vs_acc, /*rows_prescaled=*/true);
	if (r)
		return r;

	// results -> pinned staging
	const size_t wb = scope_wave_bytes(out_w);
	uint8_t *h = sl.h_res;
	if (hist_on) {
		CU_TRY(ctx, cudaMemcpyAsync(h, sl.d_hist, 4096, cudaMemcpyDeviceToHost, sl.stream));
		CU_TRY(ctx, cudaMemcpyAsync(h + 4096, sl.d_hist_max, 16, cudaMemcpyDeviceToHost, sl.stream));
	}
	if (vs_on) {
		CU_TRY(ctx, cudaMemcpyAsync(h + 4112, sl.d_vscope, 65536, cudaMemcpyDeviceToHost, sl.stream));
		if (pr->vscope_intensity > 0)
			CU_TRY(ctx, cudaMemcpyAsync(h + 4112 + 65536, sl.d_vscope_disp, 65536, cudaMemcpyDeviceToHost,
						    sl.stream));
	}
	if (wave_on) {
		CU_TRY(ctx, cudaMemcpyAsync(h + 4112 + 131072, sl.d_wave, wb, cudaMemcpyDeviceToHost, sl.stream));
		if (pr->wave_intensity > 0)
			CU_TRY(ctx, cudaMemcpyAsync(h + 4112 + 131072 + wb, sl.d_wave_disp, wb, cudaMemcpyDeviceToHost,
						    sl.stream));
	}
	CU_TRY(ctx, cudaEventRecord(sl.done, sl.stream));
	sl.params = *pr;
	sl.width = out_w;
	sl.height = rows;
	sl.in_flight = true;
	return SCOPE_OK;
}

int wait_host(scope_ctx *ctx, int slot, const scope_out_host *out)
{
	if (slot < 0 || slot >= SCOPE_RING_SLOTS)
		return fail(ctx, SCOPE_ERR_INVALID, "bad ring slot");
	RingSlot &sl = ctx->ring[slot];
	if (!sl.in_flight)
		return fail(ctx, SCOPE_ERR_INVALID, "ring slot has no submission");
	cudaError_t e = cudaEventSynchronize(sl.done);
	sl.in_flight = false;
	if (e != cudaSuccess)
		return fail(ctx, SCOPE_ERR_CUDA, "cudaEventSynchronize", e);
	if (!out)
		return SCOPE_OK;
	const scope_params &pr = sl.params;
	const size_t wb = scope_wave_bytes(sl.width);
	const uint8_t *h = sl.h_res;
	if (pr.scopes & SCOPE_HIST) {
		const uint32_t *counts = reinterpret_cast<const uint32_t *>(h);
		uint32_t hi[3];
		memcpy(hi, h + 4096, sizeof hi);
		if (out->hist_counts)
			memcpy(out->hist_counts, counts, 4096);
		if (out->hist_float) {
			hist_to_float(&pr, counts, hi, out->hist_float);
		} else if (pr.logscale) {
			for (int j = 0, mask = 0x44; j < 3; j++, mask >>= 1)
				if (pr.hist_components & (uint32_t)mask)
					hi[j] = 1;
		}
		if (out->hist_max && (pr.hist_components & 0x77u)) // else untouched, like histogram.c:366-373
			memcpy(out->hist_max, hi, sizeof hi);
	}
	if (pr.scopes & SCOPE_VSCOPE) {
		if (out->vscope)
			memcpy(out->vscope, h + 4112, 65536);
		if (out->vscope_display && pr.vscope_intensity > 0)
			memcpy(out->vscope_display, h + 4112 + 65536, 65536);
	}
	if (pr.scopes & SCOPE_WAVE) {
		if (out->wave)
			memcpy(out->wave, h + 4112 + 131072, wb);
		if (out->wave_display && pr.wave_intensity > 0)
			memcpy(out->wave_display, h + 4112 + 131072 + wb, wb);
	}
	return SCOPE_OK;
}

struct DeviceGuard {
	int prev = -1;
	explicit DeviceGuard(int dev)
	{
		cudaGetDevice(&prev);
		if (prev != dev)
			cudaSetDevice(dev);
		else
			prev = -1;
	}
	~DeviceGuard()
	{
		if (prev >= 0)
			cudaSetDevice(prev);
	}
};

} // namespace

// ===========================================================================
// C-ABI
// ===========================================================================
extern "C" {

int scope_abi_version(void)
{
	return SCOPE_ABI_VERSION;
}

size_t scope_wave_bytes(uint32_t width)
{
	return (size_t)256 * width * 4;
}

size_t scope_partial_wave_words(uint32_t width)
{
	return (size_t)256 * width * 2;
}

int scope_ctx_create(int device, scope_ctx **out_ctx)
{
	if (!out_ctx)
		return fail(nullptr, SCOPE_ERR_INVALID, "out_ctx is NULL");
	*out_ctx = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(nullptr, SCOPE_ERR_NO_DEVICE,
			    "no CUDA device available (libscope_b200 has no CPU fallback)", e);
	if (device < 0) {
		e = cudaGetDevice(&device);
		if (e != cudaSuccess)
			return fail(nullptr, SCOPE_ERR_CUDA, "cudaGetDevice", e);
	}
	if (device >= count)
		return fail(nullptr, SCOPE_ERR_INVALID, "device index out of range");
	DeviceGuard guard(device);
	cudaDeviceProp prop;
	e = cudaGetDeviceProperties(&prop, device);
	if (e != cudaSuccess)
		return fail(nullptr, SCOPE_ERR_CUDA, "cudaGetDeviceProperties", e);
	if (prop.major != 10)
		return fail(nullptr, SCOPE_ERR_NO_DEVICE,
			    "device is not compute capability 10.x: this library carries sm_100a code only");
	scope_ctx *ctx = new (std::nothrow) scope_ctx();
	if (!ctx)
		return fail(nullptr, SCOPE_ERR_NOMEM, "out of host memory");
	ctx->device = device;
	ctx->sm_count = prop.multiProcessorCount;
	void *fn = nullptr;
	cudaDriverEntryPointQueryResult qres;
	e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
	if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess)
		ctx->encode = reinterpret_cast<PFN_encodeTiled>(fn);
	else
		(void)cudaGetLastError();
	*out_ctx = ctx;
	return SCOPE_OK;
}

void scope_ctx_destroy(scope_ctx *ctx)
{
	if (!ctx)
		return;
	DeviceGuard guard(ctx->device);
	for (RingSlot &sl : ctx->ring) {
		if (sl.stream)
			cudaStreamSynchronize(sl.stream);
		cudaFree(sl.d_in);
		cudaFree(sl.d_hist);
		cudaFree(sl.d_hist_max);
		cudaFree(sl.d_wave);
		cudaFree(sl.d_wave_disp);
		cudaFree(sl.d_vscope);
		cudaFree(sl.d_vscope_disp);
		cudaFree(sl.d_vs_acc);
		if (sl.h_in)
			cudaFreeHost(sl.h_in);
		if (sl.h_res)
			cudaFreeHost(sl.h_res);
		if (sl.done)
			cudaEventDestroy(sl.done);
		if (sl.stream)
			cudaStreamDestroy(sl.stream);
	}
	cudaFree(ctx->d_vs_acc);
	cudaFree(ctx->d_hist_scratch);
	cudaFree(ctx->d_counters);
	if (ctx->scratch_free)
		cudaEventDestroy(ctx->scratch_free);
	for (auto &ev : ctx->prof_events)
		ctx->prof_pool.push_back(ev);
	for (auto &ev : ctx->prof_pool) {
		cudaEventDestroy(ev.first);
		cudaEventDestroy(ev.second);
	}
	delete ctx;
}

const char *scope_last_error(const scope_ctx *ctx)
{
	return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

uint64_t scope_launch_count(const scope_ctx *ctx)
{
	return ctx ? ctx->launches : 0;
}

int scope_sm_count(const scope_ctx *ctx)
{
	return ctx ? ctx->sm_count : 0;
}

int scope_accumulate_device(scope_ctx *ctx, const struct scope_params *params, const struct scope_surface *surface,
			    uint32_t n_frames, size_t frame_stride, const struct scope_out_device *out, void *stream)
{
	if (!ctx)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	return run_device(ctx, params, surface, n_frames, frame_stride, out, (cudaStream_t)stream);
}

int scope_submit_host(scope_ctx *ctx, int slot, const struct scope_params *params, const struct scope_surface *surface)
{
	if (!ctx)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	return submit_host(ctx, slot, params, surface);
}

int scope_wait_host(scope_ctx *ctx, int slot, const struct scope_out_host *out)
{
	if (!ctx)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	return wait_host(ctx, slot, out);
}

int scope_ring_input(scope_ctx *ctx, int slot, size_t bytes, void **out_ptr)
{
	if (!ctx)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	if (!out_ptr || slot < 0 || slot >= SCOPE_RING_SLOTS)
		return fail(ctx, SCOPE_ERR_INVALID, "scope_ring_input: bad slot or NULL out_ptr");
	*out_ptr = nullptr;
	RingSlot &sl = ctx->ring[slot];
	if (sl.in_flight) // the DMA of the submission in flight may still be reading the buffer
		return fail(ctx, SCOPE_ERR_BUSY, "ring slot still in flight");
	if (sl.h_in_bytes < bytes) {
		if (sl.h_in)
			cudaFreeHost(sl.h_in);
		sl.h_in = nullptr;
		sl.h_in_bytes = 0;
		cudaError_t e = cudaHostAlloc(&sl.h_in, bytes ? bytes : 1, cudaHostAllocDefault);
		if (e != cudaSuccess)
			return fail(ctx, SCOPE_ERR_NOMEM, "cudaHostAlloc(input staging)", e);
		sl.h_in_bytes = bytes;
	}
	*out_ptr = sl.h_in;
	return SCOPE_OK;
}

int scope_accumulate_host(scope_ctx *ctx, const struct scope_params *params, const struct scope_surface *surface,
			  const struct scope_out_host *out)
{
	if (!ctx)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	// use the first idle slot so a pending stream submission is not disturbed
	int slot = -1;
	for (int i = 0; i < SCOPE_RING_SLOTS; i++)
		if (!ctx->ring[i].in_flight) {
			slot = i;
			break;
		}
	if (slot < 0)
		return fail(ctx, SCOPE_ERR_BUSY, "all ring slots in flight");
	int r = submit_host(ctx, slot, params, surface);
	if (r)
		return r;
	return wait_host(ctx, slot, out);
}

namespace {
// One band of a tile-sharded frame.  wave_outs == NULL: the waveform goes into partial->wave_pairs (u16 pairs, added,
// or stored when `exclusive`).  wave_outs != NULL: the band spans the full height of its columns, so its waveform
// columns are final: written as saturated u8 into every wave_outs[i] (local and peer images), no partial, no reduce.
int accumulate_band(scope_ctx *ctx, const struct scope_params *pr, const struct scope_surface *tile, uint32_t x_offset,
		    uint32_t full_width, const struct scope_partial_device *partial, uint8_t *const *wave_outs,
		    uint32_t n_wave_outs, bool exclusive, cudaStream_t st)
{
	if (!pr || !tile || (!partial && !wave_outs))
		return fail(ctx, SCOPE_ERR_INVALID, "NULL argument");
	if (tile->width == 0 || tile->height == 0 || x_offset + tile->width > full_width)
		return fail(ctx, SCOPE_ERR_INVALID, "bad tile geometry");
	if (wave_outs && (n_wave_outs == 0 || n_wave_outs > (uint32_t)kMaxWaveCopies + 1u))
		return fail(ctx, SCOPE_ERR_UNSUPPORTED, "scope_accumulate_band: 1..16 waveform outputs");
	if (pr->target_scale > 1)
		return fail(ctx, SCOPE_ERR_UNSUPPORTED, "tile-sharded frames take the scaled surface (target_scale <= 1 here)");
	const bool surface = pr->mode == SCOPE_MODE_SURFACE;
	const bool want_hist = (pr->scopes & SCOPE_HIST) && partial && partial->hist_counts;
	const bool want_wave = (pr->scopes & SCOPE_WAVE) && (wave_outs ? wave_outs[0] != nullptr : (partial && partial->wave_pairs));
	const bool want_vs = (pr->scopes & SCOPE_VSCOPE) && partial && partial->vscope_counts;
	int hsrc = SRC_NONE, wsrc = SRC_NONE;
	uint32_t hmask = 0, wmask = 0;
	if (want_hist)
		decode_components(pr->hist_components, hsrc, hmask);
	if (want_wave)
		decode_components(pr->wave_components, wsrc, wmask);
	Request rq{};
	rq.rgb = tile->rgb_data;
	rq.yuv = surface ? tile->yuv_data : nullptr;
	rq.linesize = tile->linesize;
	rq.width = tile->width;
	rq.height = tile->height;
	rq.n_frames = 1;
	rq.frame_stride = 0;
	rq.colorspace = tile->colorspace;
	rq.surface = surface;
	rq.strict = !surface && pr->xform == SCOPE_XFORM_FP32_STRICT;
	rq.hist = partial ? partial->hist_counts : nullptr;
	rq.hist_stride = 1024;
	rq.x_offset = x_offset;
	rq.out_width = full_width;
	if (wave_outs) {
		rq.wave = wave_outs[0];
		rq.wave_stride = scope_wave_bytes(full_width);
		rq.wave_copies = wave_outs + 1;
		rq.n_wave_copies = n_wave_outs - 1;
		rq.partial = 0;
		if (want_wave && wsrc == SRC_NONE) // components select no plane: the reference leaves zeros
			return fail(ctx, SCOPE_ERR_UNSUPPORTED, "scope_accumulate_band: wave_components select no plane");
	} else {
		rq.wave = nullptr;
		rq.wave_stride = 0;
		rq.wave_pairs = partial->wave_pairs;
		rq.partial = exclusive ? 2 : 1;
	}
	rq.vs_acc = partial ? partial->vscope_counts : nullptr;
	rq.vs_stride = 65536;
	if (hsrc != SRC_NONE && wsrc != SRC_NONE && hsrc != wsrc) {
		Request a = rq;
		a.src = wsrc;
		a.bins_mask = a.wave_mask = wmask;
		a.hist_mask = 0;
		a.vscope = want_vs;
		int r = launch_strip(ctx, a, st);
		if (r)
			return r;
		Request b = rq;
		b.src = hsrc;
		b.bins_mask = b.hist_mask = hmask;
		b.wave_mask = 0;
		b.vscope = false;
		return launch_strip(ctx, b, st);
	}
	Request a = rq;
	a.src = hsrc != SRC_NONE ? hsrc : wsrc;
	a.hist_mask = hmask;
	a.wave_mask = wmask;
	a.bins_mask = hmask | wmask;
	a.vscope = want_vs;
	return launch_strip(ctx, a, st);
}
} // namespace

int scope_accumulate_partial(scope_ctx *ctx, const struct scope_params *pr, const struct scope_surface *tile,
			     uint32_t x_offset, uint32_t full_width, const struct scope_partial_device *partial,
			     void *stream)
{
	if (!ctx)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	if (!partial)
		return fail(ctx, SCOPE_ERR_INVALID, "NULL argument");
	return accumulate_band(ctx, pr, tile, x_offset, full_width, partial, nullptr, 0, false, (cudaStream_t)stream);
}

int scope_accumulate_band(scope_ctx *ctx, const struct scope_params *pr, const struct scope_surface *tile,
			  uint32_t x_offset, uint32_t full_width, const struct scope_partial_device *partial,
			  uint8_t *const *wave_outs, uint32_t n_wave_outs, uint32_t flags, void *stream)
{
	if (!ctx)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	return accumulate_band(ctx, pr, tile, x_offset, full_width, partial, wave_outs, n_wave_outs,
			       (flags & SCOPE_BAND_EXCLUSIVE) != 0, (cudaStream_t)stream);
}

int scope_finalize_partial(scope_ctx *ctx, const struct scope_params *pr, uint32_t full_width, uint32_t full_height,
			   const struct scope_partial_device *partial, const struct scope_out_device *out, void *stream)
{
	if (!ctx)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	if (!pr || !partial || !out)
		return fail(ctx, SCOPE_ERR_INVALID, "NULL argument");
	// the summed u16 halves of wave_pairs count up to full_height: a taller frame would carry from the
	// B|U half into the G|Y half
	if (full_height > 65535u)
		return fail(ctx, SCOPE_ERR_UNSUPPORTED, "full_height > 65535 (u16 halves of the partial waveform)");
	cudaStream_t st = (cudaStream_t)stream;
	if ((pr->scopes & SCOPE_VSCOPE) && partial->vscope_counts && (out->vscope || out->vscope_display)) {
		const float k = pr->vscope_intensity > 0 ? (float)pr->vscope_intensity : 1.0f;
		vscope_finalize_kernel<<<dim3(64, 1), 256, 0, st>>>(partial->vscope_counts, 65536, out->vscope,
								     pr->vscope_intensity > 0 ? out->vscope_display : nullptr,
								     65536, k);
		CU_TRY(ctx, cudaGetLastError());
		ctx->launches++;
	}
	if ((pr->scopes & SCOPE_WAVE) && partial->wave_pairs && out->wave) {
		const size_t n_px = (size_t)256 * full_width;
		// plane 1 only ever holds the R|V channel: without it the plane need not even be reduced (scope_ffi.h)
		wave_pairs_finalize_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, st>>>(
			partial->wave_pairs, out->wave, n_px, (pr->wave_components & 0x44u) ? 1 : 0);
		CU_TRY(ctx, cudaGetLastError());
		ctx->launches++;
		if (out->wave_display && pr->wave_intensity > 0) {
			wave_display_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, st>>>(out->wave, out->wave_display, n_px,
										       (float)pr->wave_intensity);
			CU_TRY(ctx, cudaGetLastError());
			ctx->launches++;
		}
	}
	if ((pr->scopes & SCOPE_HIST) && partial->hist_counts) {
		if (out->hist_counts && out->hist_counts != partial->hist_counts)
			CU_TRY(ctx, cudaMemcpyAsync(out->hist_counts, partial->hist_counts, 4096, cudaMemcpyDeviceToDevice,
						    st));
		if (out->hist_max && (pr->hist_components & 0x77u)) {
			hist_max_kernel<<<1, 256, 0, st>>>(partial->hist_counts, 1024, out->hist_max, pr->hist_components,
							    full_width, full_height, pr->level_fixed_value,
							    pr->level_ratio_value);
			CU_TRY(ctx, cudaGetLastError());
			ctx->launches++;
		}
	}
	return SCOPE_OK;
}

// Tile-sharded frames without NCCL: the sum over the ranks' partials, the saturation and the
// distribution of the result in one kernel over peer memory (scope_peer_reduce.cuh).
namespace {
// outs[0] is the local output (histogram, hi_max).  images: where the u8 images of the slice go - the same
// array as outs (peer form) or ONE entry of multicast addresses (multicast_out).
int finalize_peers_impl(scope_ctx *ctx, const struct scope_params *pr, uint32_t full_width, uint32_t full_height,
			const struct scope_partial_device *partials, uint32_t n_partials, uint32_t slice_index,
			uint32_t slice_count, const struct scope_out_device *outs, const struct scope_out_device *images,
			uint32_t n_outs, void *stream, bool multicast_in, bool multicast_out)
{
	if (n_partials == 0 || n_partials > (uint32_t)kMaxPeers || n_outs == 0 || n_outs > (uint32_t)kMaxPeers)
		return fail(ctx, SCOPE_ERR_UNSUPPORTED, "scope_finalize_peers: 1..16 partials and 1..16 outputs");
	if (slice_count == 0 || slice_index >= slice_count)
		return fail(ctx, SCOPE_ERR_INVALID, "scope_finalize_peers: slice_index must be < slice_count");
	if (full_width == 0 || full_height == 0)
		return fail(ctx, SCOPE_ERR_INVALID, "scope_finalize_peers: empty frame");
	if (full_height > 65535u)
		return fail(ctx, SCOPE_ERR_UNSUPPORTED, "scope_finalize_peers: full_height > 65535 (u16 halves of the partial waveform)");
	cudaStream_t st = (cudaStream_t)stream;

	const bool want_vs = (pr->scopes & SCOPE_VSCOPE) != 0;
	const bool want_wave = (pr->scopes & SCOPE_WAVE) != 0;
	const bool want_hist = (pr->scopes & SCOPE_HIST) != 0 && outs[0].hist_counts;
	auto misaligned = [](const void *p) { return ((uintptr_t)p & 15u) != 0; };

	PeerReduceParams P;
	memset(&P, 0, sizeof P);
	P.n_partials = n_partials;
	P.n_outs = n_outs;
	P.multicast = multicast_in ? 1u : 0u;
	P.multicast_out = multicast_out ? 1u : 0u;
	P.n_px = (unsigned long long)256 * full_width;
	for (uint32_t i = 0; i < n_partials; i++) {
		P.hist[i] = partials[i].hist_counts;
		P.pairs[i] = partials[i].wave_pairs;
		P.vscope[i] = partials[i].vscope_counts;
		if ((want_hist && !P.hist[i]) || (want_wave && !P.pairs[i]) || (want_vs && !P.vscope[i]))
			return fail(ctx, SCOPE_ERR_INVALID, "scope_finalize_peers: a partial lacks an array a requested scope needs");
		if (misaligned(P.hist[i]) || misaligned(P.pairs[i]) || misaligned(P.vscope[i]))
			return fail(ctx, SCOPE_ERR_INVALID, "scope_finalize_peers: partial arrays must be 16-byte aligned");
	}
	bool any_wave = false, any_vs = false;
	for (uint32_t r = 0; r < n_outs; r++) {
		if (want_wave) {
			P.wave[r] = images[r].wave;
			P.wave_display[r] = pr->wave_intensity > 0 ? images[r].wave_display : nullptr;
			any_wave = any_wave || P.wave[r] || P.wave_display[r];
		}
		if (want_vs) {
			P.vs_out[r] = images[r].vscope;
			P.vs_display[r] = pr->vscope_intensity > 0 ? images[r].vscope_display : nullptr;
			any_vs = any_vs || P.vs_out[r] || P.vs_display[r];
		}
		if (misaligned(P.wave[r]) || misaligned(P.wave_display[r]) || misaligned(P.vs_out[r]) ||
		    misaligned(P.vs_display[r]))
			return fail(ctx, SCOPE_ERR_INVALID, "scope_finalize_peers: output images must be 16-byte aligned");
	}
	P.wave_intensity = (float)pr->wave_intensity;
	P.vs_intensity = (float)pr->vscope_intensity;
	// plane 1 of the pairs only ever holds the R|V channel (scope_partial_device)
	P.wave_planes = !any_wave ? 0u : (pr->wave_components & 0x44u) ? 2u : 1u;

	// slice k of n over q quads: [k*q/n, (k+1)*q/n)
	auto slice = [&](unsigned long long quads, uint32_t &q0, uint32_t &q1) {
		q0 = (uint32_t)(quads * slice_index / slice_count);
		q1 = (uint32_t)(quads * (slice_index + 1) / slice_count);
	};
	const unsigned long long wave_quads = P.n_px / 4;
	if (wave_quads > 0xFFFFFFFFull)
		return fail(ctx, SCOPE_ERR_UNSUPPORTED, "scope_finalize_peers: frame too wide");
	const uint32_t max_blocks = (uint32_t)(ctx->sm_count > 0 ? ctx->sm_count : 148) * 8u;
	if (P.wave_planes) {
		slice(wave_quads, P.wave_q0, P.wave_q1);
		const uint32_t n = P.wave_q1 - P.wave_q0;
		P.wave_blocks = n ? std::min((n + 255u) / 256u, max_blocks) : 0u;
	}
	if (any_vs) {
		slice(16384, P.vs_q0, P.vs_q1);
		const uint32_t n = P.vs_q1 - P.vs_q0;
		P.vs_blocks = (n + 255u) / 256u;
	}
	if (want_hist) {
		if (misaligned(outs[0].hist_counts))
			return fail(ctx, SCOPE_ERR_INVALID, "scope_finalize_peers: hist_counts must be 16-byte aligned");
		P.hist_out = outs[0].hist_counts;
		P.hist_blocks = 1;
	}
	const uint32_t grid = P.wave_blocks + P.vs_blocks + P.hist_blocks;
	if (grid) {
		peer_reduce_finalize_kernel<<<grid, 256, 0, st>>>(P);
		CU_TRY(ctx, cudaGetLastError());
		ctx->launches++;
	}
	if (want_hist && outs[0].hist_max && (pr->hist_components & 0x77u)) {
		hist_max_kernel<<<1, 256, 0, st>>>(outs[0].hist_counts, 1024, outs[0].hist_max, pr->hist_components,
						    full_width, full_height, pr->level_fixed_value, pr->level_ratio_value);
		CU_TRY(ctx, cudaGetLastError());
		ctx->launches++;
	}
	return SCOPE_OK;
}
} // namespace

int scope_finalize_peers(scope_ctx *ctx, const struct scope_params *pr, uint32_t full_width, uint32_t full_height,
			 const struct scope_partial_device *partials, uint32_t n_partials, uint32_t slice_index,
			 uint32_t slice_count, const struct scope_out_device *outs, uint32_t n_outs, void *stream)
{
	if (!ctx)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	if (!pr || !partials || !outs)
		return fail(ctx, SCOPE_ERR_INVALID, "NULL argument");
	return finalize_peers_impl(ctx, pr, full_width, full_height, partials, n_partials, slice_index, slice_count, outs,
				   outs, n_outs, stream, false, false);
}

// The NVLS form: the NVSwitch sums the ranks' partials (multimem.ld_reduce) and, when mc_images is given,
// replicates the result slice into every rank's images (multimem.st).
int scope_finalize_multicast(scope_ctx *ctx, const struct scope_params *pr, uint32_t full_width, uint32_t full_height,
			     const struct scope_partial_device *mc_partials, uint32_t slice_index, uint32_t slice_count,
			     const struct scope_out_device *local_out, const struct scope_out_device *mc_images,
			     void *stream)
{
	if (!ctx)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	if (!pr || !mc_partials || !local_out)
		return fail(ctx, SCOPE_ERR_INVALID, "NULL argument");
	return finalize_peers_impl(ctx, pr, full_width, full_height, mc_partials, 1, slice_index, slice_count, local_out,
				   mc_images ? mc_images : local_out, 1, stream, true, mc_images != nullptr);
}

int scope_profile_enable(scope_ctx *ctx, int on)
{
	if (!ctx)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	ctx->profiling = on != 0;
	return SCOPE_OK;
}

int scope_profile_read(scope_ctx *ctx, float *ms_out, int max_entries)
{
	if (!ctx)
		return -1;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	int n = 0;
	for (auto &ev : ctx->prof_events) {
		float ms = 0.0f;
		if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess &&
		    ms_out && n < max_entries)
			ms_out[n++] = ms;
		ctx->prof_pool.push_back(ev);
	}
	ctx->prof_events.clear();
	return n;
}

void *scope_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
		(void)cudaGetLastError();
		return nullptr;
	}
	return p;
}

void scope_host_free(void *p)
{
	if (p)
		cudaFreeHost(p);
}

// test hook (not part of the drop-in surface): the kernel's transform over all 2^24 colours
int scope_debug_yuv_table(scope_ctx *ctx, int colorspace, uint32_t *d_out /* device, 1<<24 u32 */, void *stream)
{
	if (!ctx || !d_out)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	yuv_table_kernel<<<(1u << 24) / 512, 256, 0, (cudaStream_t)stream>>>(coef_for(colorspace), d_out);
	CU_TRY(ctx, cudaGetLastError());
	ctx->launches++;
	return SCOPE_OK;
}

// test hook: SCOPE_XFORM_FP32_STRICT over all 2^24 colours, d_out[r<<16|g<<8|b] = u | y<<8 | v<<16
int scope_debug_yuv_table_strict(scope_ctx *ctx, int colorspace, uint32_t *d_out /* device, 1<<24 u32 */, void *stream)
{
	if (!ctx || !d_out)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	yuv_table_strict_kernel<<<(1u << 24) / 512, 256, 0, (cudaStream_t)stream>>>(colorspace, d_out);
	CU_TRY(ctx, cudaGetLastError());
	ctx->launches++;
	return SCOPE_OK;
}

// test hook: the transform of scope_fused_kernel_v3 (funnel shift + FADD2 / FFMA2 division) over all 2^24 colours,
// d_out[r<<16|g<<8|b] = u | v<<8
int scope_debug_uv_table_v3(scope_ctx *ctx, int colorspace, uint32_t *d_out /* device, 1<<24 u32 */, void *stream)
{
	if (!ctx || !d_out)
		return SCOPE_ERR_INVALID;
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	if (colorspace == 1)
		uv_table_kernel_v3<1><<<(1u << 24) / 512, 256, 0, (cudaStream_t)stream>>>(d_out);
	else
		uv_table_kernel_v3<2><<<(1u << 24) / 512, 256, 0, (cudaStream_t)stream>>>(d_out);
	CU_TRY(ctx, cudaGetLastError());
	ctx->launches++;
	return SCOPE_OK;
}

} // extern "C"
