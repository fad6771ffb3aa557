/* cm_shim.h — C host side that mirrors obs-color-monitor's seam for the scope path, with the
 * per-pixel loops replaced by calls into libscope_b200 (include/scope_ffi.h).
 *
 * What is mirrored (reference file:line, tree @ e904d82):
 *   struct cm_surface_data, cm_surface_cb_t            src/common.h:24-32
 *   his_surface_cb + his_source's result double buffer  src/histogram.c:40-47,432-450
 *   wvs_surface_cb + ensure_tex_buf_size                src/waveform.c:34-41,207-218,272-289
 *   vss_surface_cb                                      src/vectorscope.c:42-48,248-265
 *   roi_register_source / roi_surface_cb fan-out        src/roi.c:313-341
 *   cm_request, the 3-slot queue, drop-on-busy producer and the "color-monitor" worker
 *                                                       src/common.h:46-68, src/common.c:260-268,
 *                                                       322-329,335-403,615-620
 *
 * Same names with a b200_ prefix, same argument meaning, same error behaviour: a callback
 * that cannot produce a result returns WITHOUT flipping w_tex_buf, so the reader keeps the
 * previous result (SURVEY.md §8(b) "error convention").
 *
 * libobs itself is not needed: where the reference receives a mapped stagesurface, the shim
 * receives the same bytes as a plain host pointer.
 */
#ifndef CM_SHIM_H
#define CM_SHIM_H

#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>

#include "scope_ffi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* identical layout to the reference's struct (gs_texture_t* is opaque here) */
struct cm_surface_data {
	uint8_t *rgb_data, *yuv_data;
	uint32_t linesize, width, height;
	int colorspace;
	void *tex; /* for bypass mode; unused on this path */
};

typedef void (*cm_surface_cb_t)(void *data, struct cm_surface_data *surface_data);

#define B200_HI_SIZE 256
#define B200_WV_SIZE 256
#define B200_VS_SIZE 256

/* ---- scope sources: the fields of his_source / wvs_source / vss_source the path uses ---- */
struct b200_his_source {
	scope_ctx *ctx;           /* worker's GPU context (not owned) */
	uint32_t mode;            /* SCOPE_MODE_SURFACE (strict drop-in) or SCOPE_MODE_FUSED */
	uint32_t components;
	int level_fixed_value, level_ratio_value;
	bool logscale;
	uint8_t *tex_buf[2];      /* float[1024] each, lazily allocated (histogram.c:443-444) */
	uint32_t hi_max[2][3];
	volatile int w_tex_buf;
};

struct b200_wvs_source {
	scope_ctx *ctx;
	uint32_t mode;
	uint32_t components;
	uint8_t *tex_buf[2];      /* width*256*4 each */
	uint32_t tex_buf_width[2];
	volatile int w_tex_buf;
};

struct b200_vss_source {
	scope_ctx *ctx;
	uint32_t mode;
	uint8_t *tex_buf[2];      /* 65536 each */
	int tex_cs[2];
	volatile int w_tex_buf;
};

void b200_his_init(struct b200_his_source *src, scope_ctx *ctx, uint32_t components);
void b200_his_destroy(struct b200_his_source *src);
void b200_wvs_init(struct b200_wvs_source *src, scope_ctx *ctx, uint32_t components);
void b200_wvs_destroy(struct b200_wvs_source *src);
void b200_vss_init(struct b200_vss_source *src, scope_ctx *ctx);
void b200_vss_destroy(struct b200_vss_source *src);

/* cm_surface_cb_t-compatible callbacks: what cm_request() installs in the reference */
void b200_his_surface_cb(void *data, struct cm_surface_data *surface_data);
void b200_wvs_surface_cb(void *data, struct cm_surface_data *surface_data);
void b200_vss_surface_cb(void *data, struct cm_surface_data *surface_data);

/* the callbacks' early-return rules (histogram.c:436-441, waveform.c:276-281, vectorscope.c:252-253): true =
 * the callback returns without touching its buffers and without flipping w_tex_buf */
bool b200_his_inputs_missing(const struct b200_his_source *src, const struct cm_surface_data *surface_data);
bool b200_wvs_inputs_missing(const struct b200_wvs_source *src, const struct cm_surface_data *surface_data);
bool b200_vss_inputs_missing(const struct b200_vss_source *src, const struct cm_surface_data *surface_data);

/* ---- ROI fan-out: one surface, every registered scope — as ONE fused GPU pass ---- */
#define B200_ROI_MAX_SOURCES 8
struct b200_roi_source {
	scope_ctx *ctx;
	uint32_t mode;
	pthread_mutex_t sources_mutex;
	struct b200_his_source *his[B200_ROI_MAX_SOURCES];
	struct b200_wvs_source *wvs[B200_ROI_MAX_SOURCES];
	struct b200_vss_source *vss[B200_ROI_MAX_SOURCES];
	int n_his, n_wvs, n_vss;
	/* scratch for the fused result */
	uint8_t *wave_tmp;
	uint32_t wave_tmp_width;
	/* frame interleave (roi.c:96-100,266-277,523-532; doc/dock.md:58-65): with n_interleave = 1
	 * (the reference's default) staging happens on every other tick */
	int n_interleave, i_interleave;
	bool interleave_rendered;
	/* GPU-attached capture core: the fan-out SUBMITS the surface to its ring slot and files the results of the
	 * surface before it, so that the copy of one frame overlaps the kernels and the read-back of the previous
	 * one; b200_roi_finish files the last one */
	struct {
		bool valid;
		int slot;
		struct b200_his_source *his;
		struct b200_wvs_source *wvs;
		struct b200_vss_source *vss;
		uint32_t width; /* of the scaled surface */
		int colorspace;
	} pending;
	unsigned long frames_filed; /* results filed so far (tests, bench) */
};

void b200_roi_init(struct b200_roi_source *roi, scope_ctx *ctx, uint32_t mode);
/* what the ROI's capture core has to stage for the scopes registered on it: roi_tick's
 * ROI | RAW_TEXTURE | OR of the consumers' CONVERT flags (roi.c:533-540), each consumer's flags by the rule of
 * its own update function (histogram.c:120-121, waveform.c:101-102, vectorscope.c:79).  In SCOPE_MODE_FUSED the
 * YUV plane is made on the GPU: any consumer needs the RGB plane and nobody needs CONVERT_YUV. */
uint32_t b200_roi_capture_flags(struct b200_roi_source *roi);
void b200_roi_destroy(struct b200_roi_source *roi);
int b200_roi_register_his(struct b200_roi_source *roi, struct b200_his_source *src);
int b200_roi_register_wvs(struct b200_roi_source *roi, struct b200_wvs_source *src);
int b200_roi_register_vss(struct b200_roi_source *roi, struct b200_vss_source *src);
void b200_roi_surface_cb(void *data, struct cm_surface_data *surface_data);
/* file the results of the last submitted surface (GPU-attached mode; no-op otherwise).  Call when the stream ends
 * or pauses; b200_roi_destroy does it too. */
void b200_roi_finish(struct b200_roi_source *roi);

/* ---- the capture core's queue + worker (struct cm_source, common.h:48-88) ---- */
#define B200_CM_SURFACE_QUEUE_SIZE 3
#define B200_CM_FLAG_CONVERT_RGB 1
#define B200_CM_FLAG_CONVERT_YUV 2
#define B200_CM_FLAG_RAW_TEXTURE 4 /* the ROI source's own texture, always set for an ROI (roi.c:35) */
#define B200_CM_FLAG_ROI 8 /* crop to (x0, y0)-(x1, y1) before staging (common.h:93, common.c:272-282) */

/* What the GPU-attached capture core tells a callback about the surface, through cm_surface_data.tex (a
 * gs_texture_t * that only the reference's bypass mode uses; NULL on this path in the reference). */
#define B200_CM_HINT_MAGIC 0xB200C0DEu
struct b200_cm_hint {
	uint32_t magic;
	int slot;              /* ring slot of the scope_ctx that belongs to this queue item (queue index == ring slot) */
	uint32_t target_scale; /* >= 1.  > 1: the planes are the FULL-SIZE target and the scopes see width / scale x
	                        * height / scale pixels of it (scope_params.target_scale: rows dropped by the copy,
	                        * columns picked by the kernel) */
};

struct b200_cm_queue_item {
	uint8_t *staged;          /* host copy of the surface: RGB rows then YUV rows (common.c:358-364); in GPU mode the
				   * ring slot's own page-locked input buffer (scope_ring_input), or NULL with zero_copy */
	size_t staged_bytes;
	const uint8_t *rgb, *yuv; /* zero_copy: the caller's planes (first pixel of the crop), read by DMA */
	uint32_t width, height, linesize;
	uint32_t flags;
	int colorspace;
	cm_surface_cb_t cb;
	void *cb_data;
};

struct b200_cm_source {
	struct b200_cm_queue_item queue[B200_CM_SURFACE_QUEUE_SIZE];
	volatile int i_write_queue, i_staging_queue, i_read_queue;
	bool rendered;
	pthread_t pipeline_thread;
	pthread_mutex_t pipeline_mutex;
	pthread_cond_t pipeline_cond;
	volatile bool pipeline_thread_running;
	volatile bool request_exit;
	volatile bool worker_busy; /* shim-only: lets b200_cm_drain see a callback in progress */
	cm_surface_cb_t callback;
	void *callback_data;
	uint32_t flags;
	int colorspace;
	/* statistics for tests */
	volatile unsigned long frames_dropped, frames_processed;
	/* ROI rectangle in pixels of the (already scaled) target, used when B200_CM_FLAG_ROI is set
	 * and 0 <= x0 < x1, 0 <= y0 < y1 (common.h:61, common.c:272-282); written by b200_cm_set_roi */
	int x0, x1, y0, y1;
	/* `target_scale` (common.c:88-90,124: 1..128): the staged surface is target size / target_scale
	 * (common.c:249-250).  Without a GPU attached the staging copy point-samples; with one, the full-size rows
	 * are handed on and the scale travels in the hint */
	int target_scale;
	/* GPU-attached mode (b200_cm_attach_gpu): queue slot i uses ring slot i of `gpu` */
	scope_ctx *gpu;
	bool zero_copy;
	struct b200_cm_hint hints[B200_CM_SURFACE_QUEUE_SIZE];
};

void b200_cm_create(struct b200_cm_source *src);
/* Put the capture core on the GPU library's 3-slot ring (scope_submit_host / scope_wait_host): queue slot i stages
 * into ring slot i's page-locked input buffer - ONE host copy, the stand-in for gs_stage_texture (common.c:316-320) -
 * and the callbacks find the slot in cm_surface_data.tex (struct b200_cm_hint).  zero_copy: no host copy at all; the
 * planes given to b200_cm_render_target are read by DMA and must then stay valid and unchanged until the worker
 * has finished the frame after them (page-locked memory: a mapped stagesurface, scope_host_alloc).  Call before
 * the first tick. */
void b200_cm_attach_gpu(struct b200_cm_source *src, scope_ctx *ctx, bool zero_copy);
void b200_cm_destroy(struct b200_cm_source *src);
void b200_cm_request(struct b200_cm_source *src, cm_surface_cb_t callback, void *data);
/* per-frame, in this order, like libobs calls video_tick then video_render */
void b200_cm_tick(struct b200_cm_source *src);
/* "render": stage one frame (rgb and/or yuv host planes) into the queue.  Returns false when
 * the frame was dropped because the worker still owns the slot (common.c:260-268) or the
 * frame was already rendered in this tick (common.c:225-227). */
bool b200_cm_render_target(struct b200_cm_source *src, const uint8_t *rgb, const uint8_t *yuv, uint32_t linesize,
			   uint32_t width, uint32_t height);
/* roi_send_range (roi.c:478-500): clamp the requested rectangle to the target (negative or
 * too large ends snap to the border) and hand it to the capture core, which then stages only
 * that sub-rectangle: the callbacks see width = x1 - x0, height = y1 - y0 (common.c:272-291). */
void b200_cm_set_roi(struct b200_cm_source *src, int x0in, int y0in, int x1in, int y1in, uint32_t target_width,
		     uint32_t target_height);
/* ROI pacing around the capture core: call b200_roi_tick then b200_roi_target_render once per frame.
 * roi_tick (roi.c:523-532) only ticks the capture core on the staging phase of the interleave;
 * roi_target_render (roi.c:266-277) only stages on that phase.  Returns what the reference returns
 * (true = nothing more to do this frame). */
void b200_roi_tick(struct b200_roi_source *roi, struct b200_cm_source *cm);
bool b200_roi_target_render(struct b200_roi_source *roi, struct b200_cm_source *cm, const uint8_t *rgb,
			    const uint8_t *yuv, uint32_t linesize, uint32_t width, uint32_t height);
/* test helper: block until the worker has consumed everything queued so far */
void b200_cm_drain(struct b200_cm_source *src);

/* ---- the outer plugin ABI's shape: what libobs would call ----
 * struct b200_source_info has the members of struct obs_source_info that the three scope sources fill in
 * (histogram.c:580-595, waveform.c:402-417, vectorscope.c:484-519), same names, same order, same signatures with
 * libobs's opaque types as void *: libobs calls video_tick(data, seconds) and then video_render(data, effect) on
 * the graphics thread once per frame.  libobs itself is absent here, so
 *   obs_data_t *settings  -> struct b200_settings   (the obs_data keys the path reads: common.c:88-90,124,
 *                                                     histogram.c:119-156,166-171, waveform.c:100-106,113-116,
 *                                                     vectorscope.c:129-131,157-158, util.c:15-41)
 *   obs_source_t *source  -> struct b200_target     (the target whose frame video_render captures:
 *                                                     render_target_to_texrender, common.c:141-168)
 * and get_properties / enum_active_sources (UI, scene graph) are NULL. */
struct b200_settings {
	scope_ctx *ctx;         /* GPU context of the source's worker (not owned) */
	uint32_t mode;          /* SCOPE_MODE_SURFACE | SCOPE_MODE_FUSED */
	int target_scale;       /* "target_scale", default 2 */
	int colorspace;         /* "colorspace": 0 auto (-> 709), 1 = 601, 2 = 709 (util.c:25-41) */
	uint32_t components;    /* "components", default 0x07 (histogram, waveform) */
	int intensity;          /* "intensity": waveform 51, vectorscope 25 (display side) */
	int level_mode;         /* "level_mode": 0 auto (per-channel maximum), 1 pixels (level_fixed_value), 2 ratio
	                         * (histogram.c:32-36,131-156) */
	int level_fixed_value;  /* "level_fixed_value", default 1000 */
	double level_ratio_value; /* "level_ratio_value" in percent, default 10.0; the source keeps (int)(v * 10 + 0.5) */
	bool logscale;
	bool gpu_ring;          /* true: b200_cm_attach_gpu(ctx, zero_copy) */
	bool zero_copy;
};
struct b200_target {
	/* the current frame of the target as host BGRA rows (and, in SCOPE_MODE_SURFACE, the [U,Y,V,A] plane the
	 * shader pass would have made); false = nothing to capture this frame */
	bool (*get_frame)(void *opaque, const uint8_t **rgb, const uint8_t **yuv, uint32_t *linesize, uint32_t *width,
			  uint32_t *height);
	void *opaque;
};
struct b200_source_info {
	const char *id;
	int type;              /* OBS_SOURCE_TYPE_INPUT */
	uint32_t output_flags; /* OBS_SOURCE_VIDEO | OBS_SOURCE_CUSTOM_DRAW (| OBS_SOURCE_INTERACTION) */
	const char *(*get_name)(void *type_data);
	void *(*create)(void *settings, void *source);
	void (*destroy)(void *data);
	void (*update)(void *data, void *settings);
	void (*get_defaults)(void *settings);
	void *(*get_properties)(void *data);
	uint32_t (*get_width)(void *data);
	uint32_t (*get_height)(void *data);
	void (*enum_active_sources)(void *data, void *enum_callback, void *param);
	void (*video_render)(void *data, void *effect);
	void (*video_tick)(void *data, float seconds);
};
#define B200_OBS_SOURCE_TYPE_INPUT 0
#define B200_OBS_SOURCE_VIDEO (1u << 0)
#define B200_OBS_SOURCE_CUSTOM_DRAW (1u << 3)
#define B200_OBS_SOURCE_INTERACTION (1u << 5)
extern const struct b200_source_info b200_colormonitor_histogram;   /* .id = "histogram_source" */
extern const struct b200_source_info b200_colormonitor_waveform;    /* .id = "waveform_source" */
extern const struct b200_source_info b200_colormonitor_vectorscope; /* .id = "vectorscope_source" */
/* cm_tick's outer signature (common.h:104) for a bare capture core */
void b200_cm_tick_obs(void *data, float seconds);
/* what video_render would upload and draw: the buffer the worker finished last (tex_buf[w_tex_buf ^ 1],
 * histogram.c:563-566, waveform.c:375-377, vectorscope.c:410-412) or NULL; *width = image width in pixels
 * (256 / scaled surface width / 256), *aux = hi_max[3] for the histogram, tex_cs for the vectorscope */
const uint8_t *b200_source_result(void *data, uint32_t *width, const uint32_t **aux);
/* test / bench helper: wait until the source's worker is idle */
void b200_source_drain(void *data);
/* sizeof of the structs above (0 his, 1 wvs, 2 vss, 3 roi, 4 queue item, 5 cm source, 6 settings, 7 source info,
 * 8 hint, 9 cm_surface_data, 10 target), for bindings that mirror them */
size_t b200_sizeof_struct(int which);

#ifdef __cplusplus
}
#endif
#endif /* CM_SHIM_H */
