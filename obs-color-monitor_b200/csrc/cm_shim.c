/* cm_shim.c — plain-C host code behind the reference's scope seam (see include/cm_shim.h for the
 * reference file:line each piece mirrors).  The three per-pixel loops of the reference are
 * gone: every callback forwards the surface to libscope_b200 and files the result into the
 * same double buffers, with the same flip / no-flip behaviour. */
#include "cm_shim.h"

#include <stdlib.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h> /* header-only; ranges named like the reference's profile scopes (common.c:10-21) */

/* the GPU-attached capture core's note about the surface (cm_shim.h), or NULL */
static const struct b200_cm_hint *surface_hint(const struct cm_surface_data *sd)
{
	const struct b200_cm_hint *h = sd->tex;
	return (h && h->magic == B200_CM_HINT_MAGIC) ? h : NULL;
}
static uint32_t hint_scale(const struct b200_cm_hint *h)
{
	return (h && h->target_scale > 1) ? h->target_scale : 1u;
}
/* one surface through the GPU, synchronously: on the item's own ring slot when the capture core names one (so that
 * slot i is busy exactly while queue item i is being worked on), else on any idle slot */
static int accumulate_sync(scope_ctx *ctx, const struct b200_cm_hint *h, struct scope_params *p,
			   const struct scope_surface *s, const struct scope_out_host *out)
{
	p->target_scale = hint_scale(h);
	if (!h)
		return scope_accumulate_host(ctx, p, s, out);
	int r = scope_submit_host(ctx, h->slot, p, s);
	if (r != SCOPE_OK)
		return r;
	return scope_wait_host(ctx, h->slot, out);
}

/* the reference allocates result buffers zero-filled (bzalloc) */
static void *zalloc(size_t n)
{
	return calloc(1, n ? n : 1);
}

static void fill_surface(struct scope_surface *s, const struct cm_surface_data *sd)
{
	s->rgb_data = sd->rgb_data;
	s->yuv_data = sd->yuv_data;
	s->linesize = sd->linesize;
	s->width = sd->width;
	s->height = sd->height;
	s->colorspace = sd->colorspace;
}

/* ------------------------------------------------------------------ */
/* histogram source                                                    */
/* ------------------------------------------------------------------ */
void b200_his_init(struct b200_his_source *src, scope_ctx *ctx, uint32_t components)
{
	memset(src, 0, sizeof(*src));
	src->ctx = ctx;
	src->mode = SCOPE_MODE_SURFACE;
	src->components = components;
}

void b200_his_destroy(struct b200_his_source *src)
{
	free(src->tex_buf[0]);
	free(src->tex_buf[1]);
	src->tex_buf[0] = src->tex_buf[1] = NULL;
}

bool b200_his_inputs_missing(const struct b200_his_source *src, const struct cm_surface_data *sd)
{
	/* histogram.c:436-441, including the case of both RGB and YUV bits set (a missing YUV plane then
	 * stops the callback although the RGB plane would be the one read).  In fused mode the YUV
	 * plane is made on the GPU from rgb_data. */
	const void *yuv = src->mode == SCOPE_MODE_FUSED ? sd->rgb_data : sd->yuv_data;
	if ((src->components & SCOPE_COMP_RGB) && !sd->rgb_data)
		return true;
	if ((src->components & SCOPE_COMP_YUV) && !yuv)
		return true;
	return sd->width == 0;
}

static void his_params(const struct b200_his_source *src, struct scope_params *p)
{
	memset(p, 0, sizeof(*p));
	p->scopes = SCOPE_HIST;
	p->mode = src->mode;
	p->hist_components = src->components;
	p->level_fixed_value = src->level_fixed_value;
	p->level_ratio_value = src->level_ratio_value;
	p->logscale = src->logscale;
}

void b200_his_surface_cb(void *data, struct cm_surface_data *sd)
{
	struct b200_his_source *src = data;
	if (b200_his_inputs_missing(src, sd))
		return;
	const int w = src->w_tex_buf;
	if (!src->tex_buf[w])
		src->tex_buf[w] = zalloc(sizeof(float) * B200_HI_SIZE * 4);
	if (!src->tex_buf[w])
		return;
	if (sd->height / hint_scale(surface_hint(sd)) == 0 || sd->width / hint_scale(surface_hint(sd)) == 0) {
		/* No rows: the reference's pixel loop does not iterate (histogram.c:379-395), so the
		 * buffer stays zero - as u32, as float and on the log scale alike - the level pass runs
		 * on zero counts (histogram.c:397-402, 412) and the buffer is flipped.  Nothing for the
		 * GPU to do. */
		memset(src->tex_buf[w], 0, sizeof(float) * B200_HI_SIZE * 4);
		if (src->components & (SCOPE_COMP_RGB | SCOPE_COMP_YUV)) {
			uint32_t v = 1; /* his_calculate_max on zero counts; W*H*ratio/1000 = 0 is raised to 1 too */
			if (src->level_fixed_value > 0)
				v = (uint32_t)src->level_fixed_value;
			for (int j = 0, mask = 0x44; j < 3; j++, mask >>= 1)
				src->hi_max[w][j] = (src->logscale && (src->components & (uint32_t)mask)) ? 1u : v;
		}
		src->w_tex_buf = w ^ 1;
		return;
	}

	struct scope_params p;
	his_params(src, &p);
	struct scope_surface s;
	fill_surface(&s, sd);
	struct scope_out_host out;
	memset(&out, 0, sizeof(out));
	out.hist_float = (float *)src->tex_buf[w];
	out.hist_max = src->hi_max[w];
	nvtxRangePushA("draw_histogram"); /* histogram.c:446-448 */
	const int r = accumulate_sync(src->ctx, surface_hint(sd), &p, &s, &out);
	nvtxRangePop();
	if (r != SCOPE_OK)
		return; /* keep showing the previous result */
	src->w_tex_buf = w ^ 1;
}

/* ------------------------------------------------------------------ */
/* waveform source                                                     */
/* ------------------------------------------------------------------ */
void b200_wvs_init(struct b200_wvs_source *src, scope_ctx *ctx, uint32_t components)
{
	memset(src, 0, sizeof(*src));
	src->ctx = ctx;
	src->mode = SCOPE_MODE_SURFACE;
	src->components = components;
}

void b200_wvs_destroy(struct b200_wvs_source *src)
{
	free(src->tex_buf[0]);
	free(src->tex_buf[1]);
	src->tex_buf[0] = src->tex_buf[1] = NULL;
}

/* waveform.c:207-218 */
static void wvs_ensure_tex_buf_size(struct b200_wvs_source *src, uint32_t width, int ix)
{
	if (src->tex_buf[ix] && src->tex_buf_width[ix] == width)
		return;
	if (!width)
		return;
	free(src->tex_buf[ix]);
	src->tex_buf[ix] = zalloc((size_t)width * B200_WV_SIZE * 4);
	src->tex_buf_width[ix] = width;
}

bool b200_wvs_inputs_missing(const struct b200_wvs_source *src, const struct cm_surface_data *sd)
{
	/* waveform.c:276-281 (same rule as the histogram's) */
	const void *yuv = src->mode == SCOPE_MODE_FUSED ? sd->rgb_data : sd->yuv_data;
	if ((src->components & SCOPE_COMP_RGB) && !sd->rgb_data)
		return true;
	if ((src->components & SCOPE_COMP_YUV) && !yuv)
		return true;
	return sd->width == 0;
}

void b200_wvs_surface_cb(void *data, struct cm_surface_data *sd)
{
	struct b200_wvs_source *src = data;
	if (b200_wvs_inputs_missing(src, sd))
		return;
	const int w = src->w_tex_buf;
	const struct b200_cm_hint *hint = surface_hint(sd);
	const uint32_t out_w = sd->width / hint_scale(hint); /* the scopes see the scaled surface */
	if (out_w == 0)
		return;
	wvs_ensure_tex_buf_size(src, out_w, w);
	if (!src->tex_buf[w])
		return;
	if (sd->height / hint_scale(hint) == 0) {
		/* no rows: zero-filled image and a flip (waveform.c:225-226, 240, 288) */
		memset(src->tex_buf[w], 0, (size_t)out_w * B200_WV_SIZE * 4);
		src->w_tex_buf = w ^ 1;
		return;
	}

	struct scope_params p;
	memset(&p, 0, sizeof(p));
	p.scopes = SCOPE_WAVE;
	p.mode = src->mode;
	p.wave_components = src->components;
	struct scope_surface s;
	fill_surface(&s, sd);
	struct scope_out_host out;
	memset(&out, 0, sizeof(out));
	out.wave = src->tex_buf[w];
	nvtxRangePushA("draw_waveform"); /* waveform.c:285-287 */
	const int r = accumulate_sync(src->ctx, hint, &p, &s, &out);
	nvtxRangePop();
	if (r != SCOPE_OK)
		return;
	src->w_tex_buf = w ^ 1;
}

/* ------------------------------------------------------------------ */
/* vectorscope source                                                  */
/* ------------------------------------------------------------------ */
void b200_vss_init(struct b200_vss_source *src, scope_ctx *ctx)
{
	memset(src, 0, sizeof(*src));
	src->ctx = ctx;
	src->mode = SCOPE_MODE_SURFACE;
}

void b200_vss_destroy(struct b200_vss_source *src)
{
	free(src->tex_buf[0]);
	free(src->tex_buf[1]);
	src->tex_buf[0] = src->tex_buf[1] = NULL;
}

bool b200_vss_inputs_missing(const struct b200_vss_source *src, const struct cm_surface_data *sd)
{
	/* vectorscope.c:252-253: only the plane is tested; an empty surface is still processed */
	return src->mode == SCOPE_MODE_FUSED ? !sd->rgb_data : !sd->yuv_data;
}

void b200_vss_surface_cb(void *data, struct cm_surface_data *sd)
{
	struct b200_vss_source *src = data;
	if (b200_vss_inputs_missing(src, sd))
		return;
	if (sd->width / hint_scale(surface_hint(sd)) == 0 || sd->height / hint_scale(surface_hint(sd)) == 0) {
		/* the reference zero-fills and flips even for an empty surface (vectorscope.c:219-236) */
		const int w0 = src->w_tex_buf;
		if (!src->tex_buf[w0])
			src->tex_buf[w0] = zalloc(B200_VS_SIZE * B200_VS_SIZE);
		if (!src->tex_buf[w0])
			return;
		memset(src->tex_buf[w0], 0, B200_VS_SIZE * B200_VS_SIZE);
		src->tex_cs[w0] = sd->colorspace;
		src->w_tex_buf = w0 ^ 1;
		return;
	}
	const int w = src->w_tex_buf;
	if (!src->tex_buf[w])
		src->tex_buf[w] = zalloc(B200_VS_SIZE * B200_VS_SIZE);
	if (!src->tex_buf[w])
		return;

	struct scope_params p;
	memset(&p, 0, sizeof(p));
	p.scopes = SCOPE_VSCOPE;
	p.mode = src->mode;
	struct scope_surface s;
	fill_surface(&s, sd);
	struct scope_out_host out;
	memset(&out, 0, sizeof(out));
	out.vscope = src->tex_buf[w];
	nvtxRangePushA("draw_vectorscope"); /* vectorscope.c:258-260 */
	const int r = accumulate_sync(src->ctx, surface_hint(sd), &p, &s, &out);
	nvtxRangePop();
	if (r != SCOPE_OK)
		return;
	src->tex_cs[w] = sd->colorspace;
	src->w_tex_buf = w ^ 1;
}

/* ------------------------------------------------------------------ */
/* ROI fan-out: the reference calls every registered callback in turn   */
/* (roi.c:329-341); here scopes that share settings share ONE fused      */
/* GPU pass over the surface, the rest fall back to their own callback.  */
/* ------------------------------------------------------------------ */
void b200_roi_init(struct b200_roi_source *roi, scope_ctx *ctx, uint32_t mode)
{
	memset(roi, 0, sizeof(*roi));
	roi->ctx = ctx;
	roi->mode = mode;
	pthread_mutex_init(&roi->sources_mutex, NULL);
}

void b200_roi_destroy(struct b200_roi_source *roi)
{
	b200_roi_finish(roi);
	pthread_mutex_destroy(&roi->sources_mutex);
	free(roi->wave_tmp);
	roi->wave_tmp = NULL;
}

#define ROI_REGISTER(kind)                                                                      \
	int b200_roi_register_##kind(struct b200_roi_source *roi, struct b200_##kind##_source *src) \
	{                                                                                       \
		int ok = -1;                                                                    \
		pthread_mutex_lock(&roi->sources_mutex);                                        \
		if (roi->n_##kind < B200_ROI_MAX_SOURCES) {                                     \
			roi->kind[roi->n_##kind++] = src;                                       \
			src->ctx = roi->ctx;                                                    \
			src->mode = roi->mode;                                                  \
			ok = 0;                                                                 \
		}                                                                               \
		pthread_mutex_unlock(&roi->sources_mutex);                                      \
		return ok;                                                                      \
	}
ROI_REGISTER(his)
ROI_REGISTER(wvs)
ROI_REGISTER(vss)

static uint32_t convert_flags(uint32_t components)
{
	return ((components & SCOPE_COMP_RGB) ? B200_CM_FLAG_CONVERT_RGB : 0u) |
	       ((components & SCOPE_COMP_YUV) ? B200_CM_FLAG_CONVERT_YUV : 0u);
}

uint32_t b200_roi_capture_flags(struct b200_roi_source *roi)
{
	uint32_t flags = 0;
	pthread_mutex_lock(&roi->sources_mutex);
	for (int i = 0; i < roi->n_his; i++)
		flags |= convert_flags(roi->his[i]->components);
	for (int i = 0; i < roi->n_wvs; i++)
		flags |= convert_flags(roi->wvs[i]->components);
	if (roi->n_vss)
		flags |= B200_CM_FLAG_CONVERT_YUV;
	const bool any = roi->n_his || roi->n_wvs || roi->n_vss;
	pthread_mutex_unlock(&roi->sources_mutex);
	if (roi->mode == SCOPE_MODE_FUSED)
		flags = any ? B200_CM_FLAG_CONVERT_RGB : 0u;
	return flags | B200_CM_FLAG_ROI | B200_CM_FLAG_RAW_TEXTURE;
}

/* result buffers of the sources that ride in a fused pass, sized for a surface `width` pixels wide; false if an
 * allocation failed.  Writes the buffer indices it chose. */
static bool roi_targets(struct b200_his_source *his, struct b200_wvs_source *wvs, struct b200_vss_source *vss,
			uint32_t width, struct scope_out_host *out, int *hw, int *ww, int *vw)
{
	bool ok = true;
	memset(out, 0, sizeof(*out));
	if (his) {
		*hw = his->w_tex_buf;
		if (!his->tex_buf[*hw])
			his->tex_buf[*hw] = zalloc(sizeof(float) * B200_HI_SIZE * 4);
		out->hist_float = (float *)his->tex_buf[*hw];
		out->hist_max = his->hi_max[*hw];
		ok = ok && his->tex_buf[*hw];
	}
	if (wvs) {
		*ww = wvs->w_tex_buf;
		wvs_ensure_tex_buf_size(wvs, width, *ww);
		out->wave = wvs->tex_buf[*ww];
		ok = ok && wvs->tex_buf[*ww];
	}
	if (vss) {
		*vw = vss->w_tex_buf;
		if (!vss->tex_buf[*vw])
			vss->tex_buf[*vw] = zalloc(B200_VS_SIZE * B200_VS_SIZE);
		out->vscope = vss->tex_buf[*vw];
		ok = ok && vss->tex_buf[*vw];
	}
	return ok;
}

static void roi_flip(struct b200_his_source *his, struct b200_wvs_source *wvs, struct b200_vss_source *vss, int hw,
		     int ww, int vw, int colorspace)
{
	if (his)
		his->w_tex_buf = hw ^ 1;
	if (wvs)
		wvs->w_tex_buf = ww ^ 1;
	if (vss) {
		vss->tex_cs[vw] = colorspace;
		vss->w_tex_buf = vw ^ 1;
	}
}

/* GPU-attached mode: wait for the surface submitted last and file its results (sources_mutex held) */
static void roi_file_pending(struct b200_roi_source *roi)
{
	if (!roi->pending.valid)
		return;
	roi->pending.valid = false;
	struct scope_out_host out;
	int hw = 0, ww = 0, vw = 0;
	const bool ok = roi_targets(roi->pending.his, roi->pending.wvs, roi->pending.vss, roi->pending.width, &out, &hw,
				    &ww, &vw);
	/* (the slot must be waited for even if there is nowhere to put the results) */
	if (scope_wait_host(roi->ctx, roi->pending.slot, ok ? &out : NULL) == SCOPE_OK && ok) {
		roi_flip(roi->pending.his, roi->pending.wvs, roi->pending.vss, hw, ww, vw, roi->pending.colorspace);
		roi->frames_filed++;
	}
}

void b200_roi_finish(struct b200_roi_source *roi)
{
	pthread_mutex_lock(&roi->sources_mutex);
	roi_file_pending(roi);
	pthread_mutex_unlock(&roi->sources_mutex);
}

void b200_roi_surface_cb(void *data, struct cm_surface_data *sd)
{
	struct b200_roi_source *roi = data;
	pthread_mutex_lock(&roi->sources_mutex);
	const struct b200_cm_hint *hint = surface_hint(sd);
	const uint32_t scale = hint_scale(hint);
	const uint32_t sw = sd->width / scale, sh = sd->height / scale; /* what the scopes see */

	/* the first source of each kind rides in the fused pass */
	struct b200_his_source *his = roi->n_his ? roi->his[0] : NULL;
	struct b200_wvs_source *wvs = roi->n_wvs ? roi->wvs[0] : NULL;
	struct b200_vss_source *vss = roi->n_vss ? roi->vss[0] : NULL;
	if (his && b200_his_inputs_missing(his, sd))
		his = NULL;
	if (wvs && b200_wvs_inputs_missing(wvs, sd))
		wvs = NULL;
	if (vss && b200_vss_inputs_missing(vss, sd))
		vss = NULL;

	if (sw == 0 || sh == 0) {
		/* an empty surface never reaches the GPU: each callback files its zeroed result itself */
		roi_file_pending(roi);
		if (his)
			b200_his_surface_cb(his, sd);
		if (wvs)
			b200_wvs_surface_cb(wvs, sd);
		if (vss)
			b200_vss_surface_cb(vss, sd);
	} else if (his || wvs || vss) {
		struct scope_params p;
		memset(&p, 0, sizeof(p));
		if (his)
			his_params(his, &p);
		p.mode = roi->mode;
		if (wvs)
			p.wave_components = wvs->components;
		p.scopes = (his ? SCOPE_HIST : 0) | (wvs ? SCOPE_WAVE : 0) | (vss ? SCOPE_VSCOPE : 0);
		p.target_scale = scale;
		struct scope_surface s;
		fill_surface(&s, sd);
		if (hint) {
			/* the surface goes to its ring slot and the call returns while the copy is still running;
			 * what comes back now are the results of the surface BEFORE it (one frame later than the
			 * synchronous form - the reference's reader is a frame behind its worker anyway) */
			nvtxRangePushA("draw_scopes_submit");
			const int r = scope_submit_host(roi->ctx, hint->slot, &p, &s);
			nvtxRangePop();
			roi_file_pending(roi);
			if (r == SCOPE_OK) {
				roi->pending.valid = true;
				roi->pending.slot = hint->slot;
				roi->pending.his = his;
				roi->pending.wvs = wvs;
				roi->pending.vss = vss;
				roi->pending.width = sw;
				roi->pending.colorspace = sd->colorspace;
			}
		} else {
			struct scope_out_host out;
			int hw = 0, ww = 0, vw = 0;
			const bool ok = roi_targets(his, wvs, vss, sw, &out, &hw, &ww, &vw);
			if (ok && scope_accumulate_host(roi->ctx, &p, &s, &out) == SCOPE_OK)
				roi_flip(his, wvs, vss, hw, ww, vw, sd->colorspace);
		}
	}

	/* any further sources of the same kind: their own pass, like the reference's loop */
	if (roi->n_his > 1 || roi->n_wvs > 1 || roi->n_vss > 1)
		roi_file_pending(roi); /* (they use the item's ring slot synchronously) */
	for (int i = 1; i < roi->n_his; i++)
		b200_his_surface_cb(roi->his[i], sd);
	for (int i = 1; i < roi->n_wvs; i++)
		b200_wvs_surface_cb(roi->wvs[i], sd);
	for (int i = 1; i < roi->n_vss; i++)
		b200_vss_surface_cb(roi->vss[i], sd);
	pthread_mutex_unlock(&roi->sources_mutex);
}

/* ------------------------------------------------------------------ */
/* capture core: 3-slot queue, drop-on-busy producer, worker thread     */
/* ------------------------------------------------------------------ */
void b200_cm_create(struct b200_cm_source *src)
{
	memset(src, 0, sizeof(*src));
	src->i_write_queue = 0;
	src->i_staging_queue = 0;
	src->i_read_queue = B200_CM_SURFACE_QUEUE_SIZE - 1; /* common.c:30-32 */
	src->colorspace = 2;
	src->target_scale = 1;
	pthread_mutex_init(&src->pipeline_mutex, NULL);
	pthread_cond_init(&src->pipeline_cond, NULL);
}

void b200_cm_attach_gpu(struct b200_cm_source *src, scope_ctx *ctx, bool zero_copy)
{
	src->gpu = ctx;
	src->zero_copy = ctx && zero_copy;
}

static void stop_pipeline_thread(struct b200_cm_source *src)
{
	if (!src->pipeline_thread_running)
		return;
	pthread_mutex_lock(&src->pipeline_mutex);
	src->request_exit = true;
	pthread_cond_broadcast(&src->pipeline_cond);
	pthread_mutex_unlock(&src->pipeline_mutex);
	pthread_join(src->pipeline_thread, NULL);
	src->pipeline_thread_running = false;
}

void b200_cm_destroy(struct b200_cm_source *src)
{
	stop_pipeline_thread(src);
	for (int i = 0; i < B200_CM_SURFACE_QUEUE_SIZE; i++)
		if (!src->gpu) /* (GPU mode: the staging buffers belong to the ring slots) */
			free(src->queue[i].staged);
	pthread_mutex_destroy(&src->pipeline_mutex);
	pthread_cond_destroy(&src->pipeline_cond);
}

void b200_cm_request(struct b200_cm_source *src, cm_surface_cb_t callback, void *data)
{
	src->callback = callback;
	src->callback_data = data;
}

/* common.c:335-373 with the stagesurface map replaced by the staged host copy */
static void pipeline_thread_loop(struct b200_cm_source *src, struct b200_cm_queue_item *item, int slot)
{
	if (!(item->flags & (B200_CM_FLAG_CONVERT_RGB | B200_CM_FLAG_CONVERT_YUV)) ||
	    (!item->staged && !item->rgb && !item->yuv))
		return;
	nvtxRangePushA("cm_pipeline_thread_loop"); /* common.c:10, 337 */
	struct cm_surface_data sd;
	memset(&sd, 0, sizeof(sd));
	sd.linesize = item->linesize;
	sd.width = item->width;
	sd.height = item->height;
	sd.colorspace = item->colorspace;
	if (item->staged) {
		uint8_t *video_data = item->staged;
		if (item->flags & B200_CM_FLAG_CONVERT_RGB) {
			sd.rgb_data = video_data;
			video_data += (size_t)item->linesize * item->height;
		}
		if (item->flags & B200_CM_FLAG_CONVERT_YUV)
			sd.yuv_data = video_data;
	} else { /* zero_copy: the caller's own (page-locked) planes */
		sd.rgb_data = (uint8_t *)item->rgb;
		sd.yuv_data = (uint8_t *)item->yuv;
	}
	if (src->gpu) {
		struct b200_cm_hint *h = &src->hints[slot];
		h->magic = B200_CM_HINT_MAGIC;
		h->slot = slot;
		h->target_scale = src->target_scale > 1 ? (uint32_t)src->target_scale : 1u;
		sd.tex = h;
	}
	if (item->cb)
		item->cb(item->cb_data, &sd);
	__sync_fetch_and_add(&src->frames_processed, 1);
	nvtxRangePop();
}

static void *pipeline_thread(void *data)
{
	struct b200_cm_source *src = data;
	pthread_mutex_lock(&src->pipeline_mutex);
	while (!src->request_exit) {
		const int next = (src->i_read_queue + 1) % B200_CM_SURFACE_QUEUE_SIZE;
		if (src->i_write_queue == next || src->i_staging_queue == next) {
			pthread_cond_wait(&src->pipeline_cond, &src->pipeline_mutex);
			continue;
		}
		src->i_read_queue = next;
		src->worker_busy = true;
		pthread_mutex_unlock(&src->pipeline_mutex);
		pipeline_thread_loop(src, &src->queue[next], next);
		pthread_mutex_lock(&src->pipeline_mutex);
		src->worker_busy = false;
		pthread_cond_broadcast(&src->pipeline_cond);
	}
	pthread_mutex_unlock(&src->pipeline_mutex);
	return NULL;
}

void b200_cm_tick(struct b200_cm_source *src)
{
	if (!src->pipeline_thread_running) {
		src->request_exit = false;
		if (pthread_create(&src->pipeline_thread, NULL, pipeline_thread, src) == 0)
			src->pipeline_thread_running = true;
	}
	src->rendered = 0;
}

bool b200_cm_render_target(struct b200_cm_source *src, const uint8_t *rgb, const uint8_t *yuv, uint32_t linesize,
			   uint32_t width, uint32_t height)
{
	if (src->rendered)
		return false; /* once per tick (common.c:225-227) */
	src->rendered = 1;
	if (width == 0 || height == 0)
		return false;
	const bool has_rgb = (src->flags & B200_CM_FLAG_CONVERT_RGB) && rgb;
	const bool has_yuv = (src->flags & B200_CM_FLAG_CONVERT_YUV) && yuv;

	/* back-pressure: the worker still owns the slot we would write -> drop (common.c:260-268) */
	if ((has_rgb || has_yuv) && src->i_write_queue == src->i_read_queue) {
		pthread_mutex_lock(&src->pipeline_mutex);
		src->i_staging_queue = -1;
		pthread_cond_broadcast(&src->pipeline_cond);
		pthread_mutex_unlock(&src->pipeline_mutex);
		__sync_fetch_and_add(&src->frames_dropped, 1);
		return false;
	}

	/* target_scale (common.c:249-250): the surface the scopes see is target size / scale.  The ROI rectangle is in
	 * pixels of that scaled surface (common.c:272-282). */
	const uint32_t scale = src->target_scale > 1 ? (uint32_t)src->target_scale : 1u;
	const uint32_t sw = width / scale, sh = height / scale;
	if (sw == 0 || sh == 0)
		return false;
	/* crop rectangle (common.c:272-282): the ROI if the flag is set and the rectangle is sane
	 * and inside the frame, the whole frame otherwise */
	uint32_t x = 0, y = 0, cx = sw, cy = sh;
	if ((src->flags & B200_CM_FLAG_ROI) && 0 <= src->x0 && src->x0 < src->x1 && 0 <= src->y0 && src->y0 < src->y1 &&
	    (uint32_t)src->x1 <= sw && (uint32_t)src->y1 <= sh) {
		x = (uint32_t)src->x0;
		y = (uint32_t)src->y0;
		cx = (uint32_t)src->x1 - x;
		cy = (uint32_t)src->y1 - y;
	}
	struct b200_cm_queue_item *item = &src->queue[src->i_write_queue];
	const uint8_t *planes[2] = {has_rgb ? rgb : NULL, has_yuv ? yuv : NULL};
	nvtxRangePushA("stage_surface"); /* common.c:14, 316-320 */
	uint32_t out_linesize, out_w, out_h;
	if (src->gpu) {
		/* GPU mode: the FULL-SIZE rows of the (scaled) crop travel; the scale itself is applied on the way to and
		 * on the device (struct b200_cm_hint).  Source rectangle in target pixels: */
		const uint32_t fx = x * scale, fy = y * scale;
		out_w = cx * scale;
		out_h = cy * scale;
		if (src->zero_copy) {
			item->staged = NULL;
			item->rgb = planes[0] ? planes[0] + (size_t)fy * linesize + (size_t)fx * 4u : NULL;
			item->yuv = planes[1] ? planes[1] + (size_t)fy * linesize + (size_t)fx * 4u : NULL;
			out_linesize = linesize;
		} else {
			out_linesize = out_w * 4u;
			const size_t plane = (size_t)out_linesize * out_h;
			const size_t need = plane * ((has_rgb ? 1 : 0) + (has_yuv ? 1 : 0));
			void *pinned = NULL;
			/* the ring slot of this queue index; still in flight = its results have not been filed yet:
			 * the frame is dropped like any other frame the consumer is not ready for */
			if (scope_ring_input(src->gpu, src->i_write_queue, need, &pinned) != SCOPE_OK || !pinned) {
				nvtxRangePop();
				__sync_fetch_and_add(&src->frames_dropped, 1);
				return false;
			}
			item->staged = pinned;
			item->staged_bytes = need;
			item->rgb = item->yuv = NULL;
			uint8_t *dst = item->staged;
			for (int p = 0; p < 2; p++) {
				if (!planes[p])
					continue;
				const uint8_t *from = planes[p] + (size_t)fy * linesize + (size_t)fx * 4u;
				if (out_linesize == linesize)
					memcpy(dst, from, plane);
				else
					for (uint32_t r = 0; r < out_h; r++)
						memcpy(dst + (size_t)r * out_linesize, from + (size_t)r * linesize, out_linesize);
				dst += plane;
			}
		}
	} else {
		/* no GPU attached: the staged copy IS the scaled, cropped surface (point-sampled at the texel centres,
		 * oracle/scope_oracle.c: orc_point_downsample); a whole unscaled frame keeps the caller's pitch so that it
		 * moves with one memcpy per plane */
		const bool whole = scale == 1 && cx == width && cy == height;
		out_linesize = whole ? linesize : cx * 4u;
		out_w = cx;
		out_h = cy;
		const size_t plane = (size_t)out_linesize * cy;
		const size_t need = plane * ((has_rgb ? 1 : 0) + (has_yuv ? 1 : 0));
		if (item->staged_bytes < need) {
			free(item->staged);
			item->staged = malloc(need ? need : 1);
			item->staged_bytes = item->staged ? need : 0;
			if (!item->staged) {
				nvtxRangePop();
				return false;
			}
		}
		item->rgb = item->yuv = NULL;
		uint8_t *dst = item->staged; /* "gs_stage_texture": RGB rows first, YUV rows below */
		for (int p = 0; p < 2; p++) {
			if (!planes[p])
				continue;
			if (whole) {
				memcpy(dst, planes[p], plane);
			} else if (scale == 1) {
				const uint8_t *from = planes[p] + (size_t)y * linesize + (size_t)x * 4u;
				for (uint32_t r = 0; r < cy; r++)
					memcpy(dst + (size_t)r * out_linesize, from + (size_t)r * linesize, out_linesize);
			} else {
				for (uint32_t r = 0; r < cy; r++) {
					const uint8_t *from = planes[p] + (size_t)((y + r) * scale + scale / 2u) * linesize;
					uint32_t *to = (uint32_t *)(dst + (size_t)r * out_linesize);
					for (uint32_t c = 0; c < cx; c++)
						memcpy(&to[c], from + (size_t)((x + c) * scale + scale / 2u) * 4u, 4);
				}
			}
			dst += plane;
		}
	}
	nvtxRangePop();
	cx = out_w;
	cy = out_h;
	item->width = cx;
	item->height = cy;
	item->linesize = out_linesize;
	item->flags = (has_rgb ? B200_CM_FLAG_CONVERT_RGB : 0) | (has_yuv ? B200_CM_FLAG_CONVERT_YUV : 0);
	item->colorspace = src->colorspace;
	item->cb = src->callback;
	item->cb_data = src->callback_data;

	pthread_mutex_lock(&src->pipeline_mutex);
	src->i_staging_queue = src->i_write_queue;
	src->i_write_queue = (src->i_write_queue + 1) % B200_CM_SURFACE_QUEUE_SIZE;
	pthread_cond_broadcast(&src->pipeline_cond);
	pthread_mutex_unlock(&src->pipeline_mutex);
	return true;
}

void b200_cm_set_roi(struct b200_cm_source *src, int x0in, int y0in, int x1in, int y1in, uint32_t target_width,
		     uint32_t target_height)
{
	/* roi.c:478-500 */
	const int w = (int)target_width, h = (int)target_height;
	int x0 = x0in, y0 = y0in, x1 = x1in, y1 = y1in;
	if (x0 < 0)
		x0 = 0;
	if (x1 < 0 || w < x1)
		x1 = w;
	if (y0 < 0)
		y0 = 0;
	if (y1 < 0 || h < y1)
		y1 = h;
	src->x0 = x0;
	src->y0 = y0;
	src->y1 = y1;
	src->x1 = x1;
	src->flags |= B200_CM_FLAG_ROI;
}

void b200_roi_tick(struct b200_roi_source *roi, struct b200_cm_source *cm)
{
	if (roi->interleave_rendered && roi->i_interleave++ >= roi->n_interleave)
		roi->i_interleave = 0;
	roi->interleave_rendered = false;
	if (roi->i_interleave == 0 || roi->n_interleave <= 0)
		b200_cm_tick(cm);
}

bool b200_roi_target_render(struct b200_roi_source *roi, struct b200_cm_source *cm, const uint8_t *rgb,
			    const uint8_t *yuv, uint32_t linesize, uint32_t width, uint32_t height)
{
	roi->interleave_rendered = true;
	if (roi->i_interleave != 0 && roi->n_interleave > 0)
		return true;
	b200_cm_render_target(cm, rgb, yuv, linesize, width, height);
	return roi->n_interleave <= 0;
}

void b200_cm_drain(struct b200_cm_source *src)
{
	pthread_mutex_lock(&src->pipeline_mutex);
	/* nothing staged and unread: the slot after i_read is the one the producer writes next */
	for (;;) {
		const int next = (src->i_read_queue + 1) % B200_CM_SURFACE_QUEUE_SIZE;
		const bool idle = (src->i_write_queue == next || src->i_staging_queue == next);
		if ((idle && !src->worker_busy) || !src->pipeline_thread_running)
			break;
		pthread_cond_wait(&src->pipeline_cond, &src->pipeline_mutex);
	}
	pthread_mutex_unlock(&src->pipeline_mutex);
}

/* ------------------------------------------------------------------ */
/* the outer plugin ABI's shape: what libobs would call                 */
/* (histogram.c:580-595, waveform.c:402-417, vectorscope.c:484-519)     */
/* ------------------------------------------------------------------ */
enum b200_scope_kind { KIND_HIS, KIND_WVS, KIND_VSS };

struct b200_scope_source {
	enum b200_scope_kind kind;
	struct b200_cm_source cm; /* every scope source embeds its capture core (histogram.c:39, waveform.c:33) */
	union {
		struct b200_his_source his;
		struct b200_wvs_source wvs;
		struct b200_vss_source vss;
	} u;
	const struct b200_target *target;
	int intensity;
	int level_height; /* histogram.c:170 */
};

/* util.c:25-41 with the OBS video-info lookup replaced by its default */
static int calc_colorspace(int colorspace)
{
	return (colorspace == 1 || colorspace == 2) ? colorspace : 2;
}

static void scope_update(void *data, void *settings_)
{
	struct b200_scope_source *src = data;
	const struct b200_settings *st = settings_;
	if (!st)
		return;
	/* cm_update (common.c:88-90): target_scale clamped to 1..128 */
	int scale = st->target_scale;
	if (scale < 1)
		scale = 1;
	if (scale > 128)
		scale = 128;
	src->cm.target_scale = scale;
	src->cm.colorspace = calc_colorspace(st->colorspace);
	src->intensity = st->intensity;
	const uint32_t mode = st->mode;
	switch (src->kind) {
	case KIND_HIS:
		src->u.his.mode = mode;
		src->u.his.components = st->components;
		/* his_update's level_mode switch (histogram.c:131-156): one of the two values is live, or none */
		src->u.his.level_fixed_value = st->level_mode == 1 ? st->level_fixed_value : 0;
		src->u.his.level_ratio_value = st->level_mode == 2 ? (int)(st->level_ratio_value * 10.0 + 0.5) : 0;
		src->u.his.logscale = st->logscale;
		break;
	case KIND_WVS:
		src->u.wvs.mode = mode;
		src->u.wvs.components = st->components;
		break;
	case KIND_VSS:
		src->u.vss.mode = mode;
		break;
	}
	/* the planes the capture core has to stage (histogram.c:120-121, waveform.c:101-102, vectorscope.c:79); in
	 * fused mode the YUV plane is made on the GPU from the RGB plane */
	uint32_t flags = src->kind == KIND_VSS ? B200_CM_FLAG_CONVERT_YUV : convert_flags(st->components);
	if (mode == SCOPE_MODE_FUSED)
		flags = flags ? B200_CM_FLAG_CONVERT_RGB : 0u;
	src->cm.flags = (src->cm.flags & B200_CM_FLAG_ROI) | flags;
}

static void *scope_create(enum b200_scope_kind kind, void *settings_, void *source)
{
	const struct b200_settings *st = settings_;
	if (!st || !st->ctx)
		return NULL;
	struct b200_scope_source *src = calloc(1, sizeof(*src));
	if (!src)
		return NULL;
	src->kind = kind;
	src->target = source;
	src->level_height = 200;
	b200_cm_create(&src->cm);
	if (st->gpu_ring)
		b200_cm_attach_gpu(&src->cm, st->ctx, st->zero_copy);
	switch (kind) {
	case KIND_HIS:
		b200_his_init(&src->u.his, st->ctx, st->components);
		b200_cm_request(&src->cm, b200_his_surface_cb, &src->u.his);
		break;
	case KIND_WVS:
		b200_wvs_init(&src->u.wvs, st->ctx, st->components);
		b200_cm_request(&src->cm, b200_wvs_surface_cb, &src->u.wvs);
		break;
	case KIND_VSS:
		b200_vss_init(&src->u.vss, st->ctx);
		b200_cm_request(&src->cm, b200_vss_surface_cb, &src->u.vss);
		break;
	}
	scope_update(src, settings_);
	return src;
}

static void *his_create(void *settings, void *source) { return scope_create(KIND_HIS, settings, source); }
static void *wvs_create(void *settings, void *source) { return scope_create(KIND_WVS, settings, source); }
static void *vss_create(void *settings, void *source) { return scope_create(KIND_VSS, settings, source); }

static void scope_destroy(void *data)
{
	struct b200_scope_source *src = data;
	if (!src)
		return;
	b200_cm_destroy(&src->cm); /* stops the worker first (common.c:42-58) */
	switch (src->kind) {
	case KIND_HIS:
		b200_his_destroy(&src->u.his);
		break;
	case KIND_WVS:
		b200_wvs_destroy(&src->u.wvs);
		break;
	case KIND_VSS:
		b200_vss_destroy(&src->u.vss);
		break;
	}
	free(src);
}

static void his_get_defaults(void *settings_)
{
	struct b200_settings *st = settings_; /* histogram.c:164-172 */
	st->target_scale = 2;
	st->components = SCOPE_COMP_RGB;
	st->level_fixed_value = 1000;
	st->level_ratio_value = 10.0;
}
static void wvs_get_defaults(void *settings_)
{
	struct b200_settings *st = settings_; /* waveform.c:111-117 */
	st->target_scale = 2;
	st->intensity = 51;
	st->components = SCOPE_COMP_RGB;
}
static void vss_get_defaults(void *settings_)
{
	struct b200_settings *st = settings_; /* vectorscope.c:155-161 */
	st->target_scale = 2;
	st->intensity = 25;
}

static const char *his_get_name(void *unused) { (void)unused; return "Histogram"; }
static const char *wvs_get_name(void *unused) { (void)unused; return "Waveform"; }
static const char *vss_get_name(void *unused) { (void)unused; return "Vectorscope"; }

/* histogram.c:304-321 (overlay display), waveform.c:181-199, vectorscope.c:205-215 */
static uint32_t scope_get_width(void *data)
{
	struct b200_scope_source *src = data;
	switch (src->kind) {
	case KIND_HIS:
		return B200_HI_SIZE;
	case KIND_WVS:
		return src->u.wvs.tex_buf_width[src->u.wvs.w_tex_buf ^ 1];
	default:
		return B200_VS_SIZE;
	}
}
static uint32_t scope_get_height(void *data)
{
	struct b200_scope_source *src = data;
	switch (src->kind) {
	case KIND_HIS:
		return (uint32_t)src->level_height;
	case KIND_WVS:
		return B200_WV_SIZE;
	default:
		return B200_VS_SIZE;
	}
}

void b200_cm_tick_obs(void *data, float seconds)
{
	(void)seconds;
	b200_cm_tick(data);
}
static void scope_video_tick(void *data, float seconds)
{
	struct b200_scope_source *src = data;
	b200_cm_tick_obs(&src->cm, seconds); /* `.video_tick = cm_tick`: the capture core is the struct's first member there */
}

/* his_render / wvs_render / vss_render (histogram.c:550-578, waveform.c:362-392, vectorscope.c:382-471) up to the
 * draw calls: capture the target for the worker; what they then upload and draw is b200_source_result() */
static void scope_video_render(void *data, void *effect)
{
	(void)effect;
	struct b200_scope_source *src = data;
	nvtxRangePushA("render_target"); /* common.c:12 */
	const uint8_t *rgb = NULL, *yuv = NULL;
	uint32_t linesize = 0, width = 0, height = 0;
	if (src->target && src->target->get_frame &&
	    src->target->get_frame(src->target->opaque, &rgb, &yuv, &linesize, &width, &height))
		b200_cm_render_target(&src->cm, rgb, yuv, linesize, width, height);
	nvtxRangePop();
}

const uint8_t *b200_source_result(void *data, uint32_t *width, const uint32_t **aux)
{
	struct b200_scope_source *src = data;
	static const uint32_t none[3] = {0, 0, 0};
	if (aux)
		*aux = none;
	switch (src->kind) {
	case KIND_HIS: {
		const int r = src->u.his.w_tex_buf ^ 1;
		if (width)
			*width = B200_HI_SIZE;
		if (aux)
			*aux = src->u.his.hi_max[r];
		return src->u.his.tex_buf[r];
	}
	case KIND_WVS: {
		const int r = src->u.wvs.w_tex_buf ^ 1;
		if (width)
			*width = src->u.wvs.tex_buf_width[r];
		return src->u.wvs.tex_buf[r];
	}
	default: {
		const int r = src->u.vss.w_tex_buf ^ 1;
		if (width)
			*width = B200_VS_SIZE;
		if (aux)
			*aux = (const uint32_t *)&src->u.vss.tex_cs[r];
		return src->u.vss.tex_buf[r];
	}
	}
}

void b200_source_drain(void *data)
{
	struct b200_scope_source *src = data;
	b200_cm_drain(&src->cm);
}

const struct b200_source_info b200_colormonitor_histogram = {
	.id = "histogram_source",
	.type = B200_OBS_SOURCE_TYPE_INPUT,
	.output_flags = B200_OBS_SOURCE_VIDEO | B200_OBS_SOURCE_CUSTOM_DRAW,
	.get_name = his_get_name,
	.create = his_create,
	.destroy = scope_destroy,
	.update = scope_update,
	.get_defaults = his_get_defaults,
	.get_properties = NULL,
	.get_width = scope_get_width,
	.get_height = scope_get_height,
	.enum_active_sources = NULL,
	.video_render = scope_video_render,
	.video_tick = scope_video_tick,
};
const struct b200_source_info b200_colormonitor_waveform = {
	.id = "waveform_source",
	.type = B200_OBS_SOURCE_TYPE_INPUT,
	.output_flags = B200_OBS_SOURCE_VIDEO | B200_OBS_SOURCE_CUSTOM_DRAW,
	.get_name = wvs_get_name,
	.create = wvs_create,
	.destroy = scope_destroy,
	.update = scope_update,
	.get_defaults = wvs_get_defaults,
	.get_properties = NULL,
	.get_width = scope_get_width,
	.get_height = scope_get_height,
	.enum_active_sources = NULL,
	.video_render = scope_video_render,
	.video_tick = scope_video_tick,
};
const struct b200_source_info b200_colormonitor_vectorscope = {
	.id = "vectorscope_source",
	.type = B200_OBS_SOURCE_TYPE_INPUT,
	.output_flags = B200_OBS_SOURCE_VIDEO | B200_OBS_SOURCE_CUSTOM_DRAW | B200_OBS_SOURCE_INTERACTION,
	.get_name = vss_get_name,
	.create = vss_create,
	.destroy = scope_destroy,
	.update = scope_update,
	.get_defaults = vss_get_defaults,
	.get_properties = NULL,
	.get_width = scope_get_width,
	.get_height = scope_get_height,
	.enum_active_sources = NULL,
	.video_render = scope_video_render,
	.video_tick = scope_video_tick,
};

/* sizes of the structs of cm_shim.h, for bindings that mirror them (obs-color-monitor_b200/shim.py checks its ctypes
 * layouts against these) */
size_t b200_sizeof_struct(int which)
{
	switch (which) {
	case 0:
		return sizeof(struct b200_his_source);
	case 1:
		return sizeof(struct b200_wvs_source);
	case 2:
		return sizeof(struct b200_vss_source);
	case 3:
		return sizeof(struct b200_roi_source);
	case 4:
		return sizeof(struct b200_cm_queue_item);
	case 5:
		return sizeof(struct b200_cm_source);
	case 6:
		return sizeof(struct b200_settings);
	case 7:
		return sizeof(struct b200_source_info);
	case 8:
		return sizeof(struct b200_cm_hint);
	case 9:
		return sizeof(struct cm_surface_data);
	case 10:
		return sizeof(struct b200_target);
	default:
		return 0;
	}
}
