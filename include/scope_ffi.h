/* scope_ffi.h — C-ABI of libscope_b200.so: the B200 (sm_100a) implementation of
 * obs-color-monitor's per-pixel scope accumulation path.
 *
 * This is the drop-in boundary.  Plain C types only: pointers, sizes, PODs.
 * Every entry point returns an int status (SCOPE_OK == 0) and never falls back
 * to a CPU implementation: if no CUDA device / kernel image is usable the call
 * fails loudly with SCOPE_ERR_NO_DEVICE or SCOPE_ERR_CUDA.
 *
 * Reference interfaces replaced (citations relative to the reference tree,
 * norihiro/obs-color-monitor @ e904d82):
 *
 *   struct scope_surface          <- struct cm_surface_data        src/common.h:24-30
 *   scope_accumulate_host()       <- the bodies of his_surface_cb  src/histogram.c:432-450
 *                                    (his_draw_histogram           src/histogram.c:357-418),
 *                                    wvs_surface_cb                src/waveform.c:272-289
 *                                    (wvs_draw_waveform            src/waveform.c:220-257),
 *                                    vss_surface_cb                src/vectorscope.c:248-265
 *                                    (vss_draw_vectorscope         src/vectorscope.c:217-238),
 *                                    called once per surface instead of once per scope:
 *                                    the fan-out of roi_surface_cb src/roi.c:329-341
 *   SCOPE_MODE_FUSED transform    <- PSConvertRGB_YUV601/709       data/common.effect:23-43
 *                                    (+ render_rgb_yuv             src/common.c:170-221)
 *   scope_submit_host()/scope_wait_host()
 *                                 <- the 3-deep stagesurface ring  src/common.h:46,
 *                                    src/common.c:260-268,316-329,375-403
 *   scope_params.colorspace       <- calc_colorspace               src/util.c:25-41
 *   hist post-pass fields         <- his_calculate_max / his_fix_max_level / float+log
 *                                    conversion                    src/histogram.c:330-355,397-417
 *   *_intensity / *_display       <- PSDrawBare / PSDrawOverlay    data/vectorscope.effect:27-33,
 *                                                                  data/waveform.effect:30-39
 *   scope_accumulate_partial() / scope_finalize_partial() / scope_finalize_peers() / scope_finalize_multicast()
 *                                 <- no counterpart (the reference is single threaded); they keep the
 *                                    saturating semantics of inc_uint8   src/waveform.c:201-205 and
 *                                    `if (*c < 255) ++*c`                src/vectorscope.c:233-234
 *                                    when one frame is split over GPUs: min(sum of partials, 255)
 *
 * Output layouts are byte-for-byte the reference's tex_buf layouts:
 *   histogram    uint32[256][4]   slot 0 = R|V, 1 = G|Y, 2 = B|U, 3 = 0       (histogram.c:379-395)
 *   waveform     uint8 [256][W][4] row 0 = value 255, bytes B|U,G|Y,R|V,0, saturating at 255
 *                                                                             (waveform.c:240-256)
 *   vectorscope  uint8 [256][256] row = 255 - V, column = U, saturating at 255 (vectorscope.c:226-237)
 */
#ifndef SCOPE_FFI_H
#define SCOPE_FFI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCOPE_ABI_VERSION 1

enum scope_status {
	SCOPE_OK = 0,
	SCOPE_ERR_INVALID = 1,     /* bad argument (NULL plane that the request needs, zero width, ...) */
	SCOPE_ERR_NO_DEVICE = 2,   /* no CUDA device / driver: there is NO CPU fallback */
	SCOPE_ERR_CUDA = 3,        /* a CUDA call failed; see scope_last_error() */
	SCOPE_ERR_UNSUPPORTED = 4, /* geometry outside what the kernels handle (height > 65535, ...) */
	SCOPE_ERR_NOMEM = 5,
	SCOPE_ERR_BUSY = 6,        /* ring slot still in flight (the caller drops the frame, common.c:260-268) */
};

/* which scopes to accumulate (bit mask) */
#define SCOPE_HIST 0x1u
#define SCOPE_WAVE 0x2u
#define SCOPE_VSCOPE 0x4u
#define SCOPE_ALL 0x7u

/* `components` masks, identical to the reference (histogram.c:27-30, waveform.c:26-29) */
#define SCOPE_COMP_RGB 0x07u
#define SCOPE_COMP_Y 0x20u
#define SCOPE_COMP_UV 0x50u
#define SCOPE_COMP_YUV 0x70u

/* where the YUV plane comes from */
#define SCOPE_MODE_FUSED 0   /* only rgb_data is read; BT.601/709 transform applied in registers */
#define SCOPE_MODE_SURFACE 1 /* rgb_data / yuv_data are used exactly as the reference's callbacks
                                use them (integer-only path, no transform) */

/* mirror of struct cm_surface_data (common.h:24-30); host OR device pointers
 * depending on the entry point */
struct scope_surface {
	const uint8_t *rgb_data; /* BGRA, may be NULL if no requested scope reads it */
	const uint8_t *yuv_data; /* [U,Y,V,A] bytes; ignored in SCOPE_MODE_FUSED */
	uint32_t linesize;       /* bytes between rows (>= width*4) */
	uint32_t width, height;
	int32_t colorspace;      /* 1 = BT.601, 2 = BT.709 (anything else: 709, util.c:25-41) */
};

struct scope_params {
	uint32_t scopes;          /* SCOPE_HIST | SCOPE_WAVE | SCOPE_VSCOPE */
	uint32_t mode;            /* SCOPE_MODE_FUSED | SCOPE_MODE_SURFACE */
	uint32_t hist_components; /* his_source.components */
	uint32_t wave_components; /* wvs_source.components */
	/* histogram post-pass (histogram.c:397-417) */
	int32_t level_fixed_value;
	int32_t level_ratio_value;
	int32_t logscale;
	/* display mapping; 0 = not requested */
	int32_t wave_intensity;
	int32_t vscope_intensity;
	/* point-downsample before the scopes, the reference's `target_scale` (1..128, default 2: common.c:88-90,
	 * histogram.c:166, waveform.c:113, vectorscope.c:157): 0 or 1 = none.  The surface is the FULL-SIZE target;
	 * the scopes see width / target_scale x height / target_scale pixels (common.c:249-250), pixel (x, y) of
	 * them being the source pixel (x s + s / 2, y s + s / 2) - the texel a point-sampled render of the target
	 * into the smaller texrender picks.  Output widths (waveform) are the scaled width. */
	uint32_t target_scale;
	/* how SCOPE_MODE_FUSED evaluates data/common.effect:23-43 (the transform is a float shader
	 * in the reference and nothing pins its rounding): */
	uint32_t xform;
	uint32_t reserved[1];
};
#define SCOPE_XFORM_EXACT 0       /* the exact value of the expression, rounded like a UNORM8 target (the default) */
#define SCOPE_XFORM_FP32_STRICT 1 /* fp32, every product and sum rounded separately, left to right, no FMA
                                     (SURVEY.md 8(c)'s draft; differs on < 0.003 % of the colours, by one step) */

/* host result buffers; NULL members are skipped */
struct scope_out_host {
	uint32_t *hist_counts;   /* [1024] raw counts */
	float *hist_float;       /* [1024] what the reference uploads as GS_RGBA32F (linear or log) */
	uint32_t *hist_max;      /* [3]    hi_max after the post-pass; left untouched when hist_components selects
				  * no plane (no bit of 0x77), like his_draw_histogram's early return, histogram.c:366-373 */
	uint8_t *wave;           /* [256*width*4] */
	uint8_t *vscope;         /* [65536] */
	uint8_t *wave_display;   /* [256*width*4] intensity applied (needs wave_intensity > 0) */
	uint8_t *vscope_display; /* [65536]       intensity applied (needs vscope_intensity > 0) */
};

/* device result buffers for the batched entry point; per frame f the arrays are
 * at base + f * stride (strides in ELEMENTS of the array type; 0 = densely packed) */
struct scope_out_device {
	uint32_t *hist_counts; /* [n][1024]          */
	uint32_t *hist_max;    /* [n][4] (3 used)    */
	uint8_t *wave;         /* [n][256*width*4]   */
	uint8_t *vscope;       /* [n][65536]         */
	uint8_t *vscope_display;
	uint8_t *wave_display;
};

/* partial (unclamped) accumulators for tile-sharded frames: sums over tiles /
 * GPUs are formed on these (e.g. NCCL all-reduce as int32), then
 * scope_finalize_partial() applies the saturation the reference applies per
 * increment.  min(sum of partials, 255) == the reference's saturating count. */
struct scope_partial_device {
	uint32_t *hist_counts; /* [1024]            additive */
	uint32_t *wave_pairs;  /* [2][256][width]   plane 0: u16 pair (B|U, G|Y) per (level, column),
	                                            plane 1: R|V; additive as int32 lanes while every
	                                            u16 stays < 65536 (a frame has <= 65535 rows).  A
	                                            waveform without the R|V channel never touches plane 1,
	                                            so only plane 0 has to be reduced */
	uint32_t *vscope_counts; /* [65536]         additive */
};

typedef struct scope_ctx scope_ctx;

int scope_abi_version(void);

/* device < 0: current device.  Fails with SCOPE_ERR_NO_DEVICE when there is no GPU. */
int scope_ctx_create(int device, scope_ctx **out_ctx);
void scope_ctx_destroy(scope_ctx *ctx);
const char *scope_last_error(const scope_ctx *ctx); /* ctx may be NULL: last create error */

/* number of kernel launches issued through this context so far */
uint64_t scope_launch_count(const scope_ctx *ctx);
/* SM count of the context's device (grid sizing), 0 on error */
int scope_sm_count(const scope_ctx *ctx);

/* ---- host buffers in, host buffers out: the drop-in for the surface callbacks ----
 * Synchronous.  surface pointers are HOST memory and only need to be valid during the call
 * (exactly like the mapped stagesurface, common.c:343-372). */
int scope_accumulate_host(scope_ctx *ctx, const struct scope_params *params, const struct scope_surface *surface,
			  const struct scope_out_host *out);

/* ---- 3-deep ring (CM_SURFACE_QUEUE_SIZE, common.h:46) for streams ----
 * scope_submit_host enqueues, on the stream of ring slot `slot` (0..2), the host->device copy of the
 * surface's planes, the kernels and the device->host copy of the results, and returns before the GPU
 * finishes; SCOPE_ERR_BUSY if the slot is still in flight (the caller drops the frame, common.c:260-268).
 * scope_wait_host blocks until the slot's results are in `out`.
 *
 * LIFETIME OF THE INPUT (the counterpart of "the mapped stagesurface is valid during the callback",
 * common.c:343-372): the planes are read by DMA, so
 *   - page-locked memory (scope_ring_input, scope_host_alloc) is read AFTER the call returns: it must stay
 *     valid and unmodified until scope_wait_host(slot) has returned.  This is the fast path, one copy;
 *   - pageable memory is copied to the driver's staging during the call (CUDA's pageable-copy rule) and may be
 *     reused as soon as the call returns; it is slower.
 * scope_accumulate_host is submit + wait in one call, so its pointers only need to live for the call.
 *
 * scope_ring_input returns the slot's own page-locked input buffer of at least `bytes` bytes - the stand-in
 * for the reference's stagesurface (gs_stage_texture target, common.c:316-320): the producer stages (or renders)
 * the frame straight into it and passes pointers into it to scope_submit_host.  SCOPE_ERR_BUSY while the slot
 * is in flight.  The buffer stays valid until a larger request for the same slot or scope_ctx_destroy. */
#define SCOPE_RING_SLOTS 3
int scope_submit_host(scope_ctx *ctx, int slot, const struct scope_params *params,
		      const struct scope_surface *surface);
int scope_wait_host(scope_ctx *ctx, int slot, const struct scope_out_host *out);
int scope_ring_input(scope_ctx *ctx, int slot, size_t bytes, void **out_ptr);

/* ---- device buffers in, device buffers out, batched ----
 * surface pointers are DEVICE memory; frame f lives at pointer + f*frame_stride
 * bytes.  Asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 * legacy default stream). */
int scope_accumulate_device(scope_ctx *ctx, const struct scope_params *params, const struct scope_surface *surface,
			    uint32_t n_frames, size_t frame_stride, const struct scope_out_device *out, void *stream);

/* ---- tile-sharded frames (ROI tiles / row bands across GPUs) ----
 * Accumulates ONE tile (surface = the tile's rows; `x_offset` = first column of
 * the tile inside the full-width accumulators, full_width = their width) into
 * caller-zeroed partial accumulators.  Asynchronous on `stream`. */
int scope_accumulate_partial(scope_ctx *ctx, const struct scope_params *params, const struct scope_surface *tile,
			     uint32_t x_offset, uint32_t full_width, const struct scope_partial_device *partial,
			     void *stream);
/* The same with two shortcuts for frames whose bands go to different GPUs:
 *   flags & SCOPE_BAND_EXCLUSIVE  this call is the ONLY writer of its columns of wave_pairs (a rank accumulates its
 *       whole row band in one call): the u16 pairs are stored instead of added, so wave_pairs needs no zero-fill and
 *       no global atomics.  (hist_counts / vscope_counts are still added to: zero them.)
 *   wave_outs != NULL  the tile spans the FULL HEIGHT of its columns (column bands): its waveform columns are final,
 *       so they are written as saturated u8 into wave_outs[0 .. n_wave_outs) - this rank's image and, through peer
 *       mappings, every other rank's - and the waveform needs neither partial sums nor a reduce step: the
 *       "all-gather" of the column bands is the kernel's own stores (128 bytes per warp and row over NVLink).
 *       partial may be NULL when only the waveform is requested; 1 <= n_wave_outs <= 16. */
#define SCOPE_BAND_EXCLUSIVE 1u
int scope_accumulate_band(scope_ctx *ctx, const struct scope_params *params, const struct scope_surface *tile,
			  uint32_t x_offset, uint32_t full_width, const struct scope_partial_device *partial,
			  uint8_t *const *wave_outs, uint32_t n_wave_outs, uint32_t flags, void *stream);
/* clamp the (summed) partials into the reference layouts */
int scope_finalize_partial(scope_ctx *ctx, const struct scope_params *params, uint32_t full_width,
			   uint32_t full_height, const struct scope_partial_device *partial,
			   const struct scope_out_device *out, void *stream);

/* ---- tile-sharded frames without a collective library: reduce + saturate over PEER MEMORY ----
 * Replaces "NCCL all-reduce of the partials, then scope_finalize_partial" by one kernel.
 * partials[0..n_partials) are the partial accumulators of ALL ranks, as addresses valid on this
 * context's device: local allocations or NVLink peer mappings (torch symmetric memory buffer_ptrs,
 * cudaIpcOpenMemHandle, cudaDeviceEnablePeerAccess ...).  The call sums slice `slice_index` of
 * `slice_count` of the waveform and vectorscope bins over all partials (16-byte peer loads), saturates
 * at 255 and stores the u8 images (and the intensity-mapped *_display images, if requested) into
 * EVERY outs[0..n_outs) - the receiving ranks' output images, again local or peer addresses.
 *   one-shot : slice 0 of 1, n_outs = 1 (own output): every rank reads everything, no peer stores;
 *   two-shot : slice = rank of world, outs = all ranks' outputs: 1/N of the reads per rank.
 * The histogram (4 KB) is always summed whole, into outs[0] only (the LOCAL output), followed by
 * hist_max like scope_finalize_partial.  All partial and output arrays must be 16-byte aligned.
 * Synchronisation across ranks is the caller's, on `stream`: every rank's scope_accumulate_partial
 * must have completed before any rank's kernel reads (a device-side barrier, e.g. the symmetric-memory
 * handle's barrier()), and in the two-shot form the outputs are complete after a second barrier.
 * 1 <= n_partials, n_outs <= 16. */
int scope_finalize_peers(scope_ctx *ctx, const struct scope_params *params, uint32_t full_width,
			 uint32_t full_height, const struct scope_partial_device *partials, uint32_t n_partials,
			 uint32_t slice_index, uint32_t slice_count, const struct scope_out_device *outs,
			 uint32_t n_outs, void *stream);

/* The same step through the NVSwitch (NVLS): mc_partials holds the MULTICAST addresses of the partial arrays
 * (one multicast object bound to every rank's buffer: torch symmetric memory's multicast_ptr,
 * cuMulticastCreate/cuMulticastBindMem).  The kernel reads each bin with multimem.ld_reduce.add - the switch
 * adds the ranks' copies, one response instead of N - saturates, and
 *   mc_images == NULL : stores the slice into local_out's images (use slice 0 of 1: every rank gets all);
 *   mc_images != NULL : stores it with multimem.st to the multicast addresses of the images, i.e. into every
 *                       rank's images at once (use slice = rank of world).
 * Histogram and hist_max go to local_out.  Same alignment and synchronisation rules as scope_finalize_peers. */
int scope_finalize_multicast(scope_ctx *ctx, const struct scope_params *params, uint32_t full_width,
			     uint32_t full_height, const struct scope_partial_device *mc_partials,
			     uint32_t slice_index, uint32_t slice_count, const struct scope_out_device *local_out,
			     const struct scope_out_device *mc_images, void *stream);

/* Per-launch device timing of the accumulation kernel (CUDA events recorded on the launch
 * stream around each launch while enabled).  scope_profile_read waits for the recorded
 * launches, writes their durations in milliseconds (oldest first) and forgets them; returns
 * how many were written. */
int scope_profile_enable(scope_ctx *ctx, int on);
int scope_profile_read(scope_ctx *ctx, float *ms_out, int max_entries);

/* page-locked host memory (what a caller should hand to the host entry points for
 * full PCIe speed; pageable pointers work too, just slower).  NULL on failure. */
void *scope_host_alloc(size_t bytes);
void scope_host_free(void *p);

/* test hook, not part of the drop-in surface: the kernels' own RGB->YUV transform for
 * all 2^24 colours, d_out[r<<16|g<<8|b] = u | y<<8 | v<<16 (device pointer). */
int scope_debug_yuv_table(scope_ctx *ctx, int colorspace, uint32_t *d_out, void *stream);
/* the same for the headline kernel's own form of the transform (scope_fused_v3.cuh): d_out[..] = u | v<<8 */
int scope_debug_uv_table_v3(scope_ctx *ctx, int colorspace, uint32_t *d_out, void *stream);
/* the same for SCOPE_XFORM_FP32_STRICT: d_out[..] = u | y<<8 | v<<16 */
int scope_debug_yuv_table_strict(scope_ctx *ctx, int colorspace, uint32_t *d_out, void *stream);

/* size helpers */
size_t scope_wave_bytes(uint32_t width);          /* 256*width*4 */
size_t scope_partial_wave_words(uint32_t width);  /* 256*width*2 */

#ifdef __cplusplus
}
#endif
#endif /* SCOPE_FFI_H */
