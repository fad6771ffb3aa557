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

/* ---- the capture core's queue + worker (struct cm_source, common.h:48-88) ---- */
#define B200_CM_SURFACE_QUEUE_SIZE 3
#define B200_CM_FLAG_CONVERT_RGB 1
#define B200_CM_FLAG_CONVERT_YUV 2
#define B200_CM_FLAG_RAW_TEXTURE 4 /* the ROI source's own texture, always set for an ROI (roi.c:35) */
#define B200_CM_FLAG_ROI 8 /* crop to (x0, y0)-(x1, y1) before staging (common.h:93, common.c:272-282) */

struct b200_cm_queue_item {
	uint8_t *staged;          /* host copy of the surface: RGB rows then YUV rows (common.c:358-364) */
	size_t staged_bytes;
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
};

void b200_cm_create(struct b200_cm_source *src);
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

#ifdef __cplusplus
}
#endif
#endif /* CM_SHIM_H */
