/* cm_shim.c — plain-C host code behind the reference's scope seam (see include/cm_shim.h for the
 * reference file:line each piece mirrors).  The three per-pixel loops of the reference are
 * gone: every callback forwards the surface to libscope_b200 and files the result into the
 * same double buffers, with the same flip / no-flip behaviour. */
#include "cm_shim.h"

#include <stdlib.h>
#include <string.h>

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
	if (sd->height == 0) {
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
	if (scope_accumulate_host(src->ctx, &p, &s, &out) != SCOPE_OK)
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
	wvs_ensure_tex_buf_size(src, sd->width, w);
	if (!src->tex_buf[w])
		return;
	if (sd->height == 0) {
		/* no rows: zero-filled image and a flip (waveform.c:225-226, 240, 288) */
		memset(src->tex_buf[w], 0, (size_t)sd->width * B200_WV_SIZE * 4);
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
	if (scope_accumulate_host(src->ctx, &p, &s, &out) != SCOPE_OK)
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
	if (sd->width == 0 || sd->height == 0) {
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
	if (scope_accumulate_host(src->ctx, &p, &s, &out) != SCOPE_OK)
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

void b200_roi_surface_cb(void *data, struct cm_surface_data *sd)
{
	struct b200_roi_source *roi = data;
	pthread_mutex_lock(&roi->sources_mutex);

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

	if (sd->width == 0 || sd->height == 0) {
		/* an empty surface never reaches the GPU: each callback files its zeroed result itself */
		if (his)
			b200_his_surface_cb(his, sd);
		if (wvs)
			b200_wvs_surface_cb(wvs, sd);
		if (vss)
			b200_vss_surface_cb(vss, sd);
	} else if (his || wvs || vss) {
		struct scope_params p;
		memset(&p, 0, sizeof(p));
		p.mode = roi->mode;
		struct scope_out_host out;
		memset(&out, 0, sizeof(out));
		int hw = 0, ww = 0, vw = 0;
		bool ok = true;
		if (his) {
			his_params(his, &p);
			p.mode = roi->mode;
			hw = his->w_tex_buf;
			if (!his->tex_buf[hw])
				his->tex_buf[hw] = zalloc(sizeof(float) * B200_HI_SIZE * 4);
			out.hist_float = (float *)his->tex_buf[hw];
			out.hist_max = his->hi_max[hw];
			ok = ok && his->tex_buf[hw];
		}
		if (wvs) {
			ww = wvs->w_tex_buf;
			wvs_ensure_tex_buf_size(wvs, sd->width, ww);
			p.wave_components = wvs->components;
			out.wave = wvs->tex_buf[ww];
			ok = ok && wvs->tex_buf[ww];
		}
		if (vss) {
			vw = vss->w_tex_buf;
			if (!vss->tex_buf[vw])
				vss->tex_buf[vw] = zalloc(B200_VS_SIZE * B200_VS_SIZE);
			out.vscope = vss->tex_buf[vw];
			ok = ok && vss->tex_buf[vw];
		}
		p.scopes = (his ? SCOPE_HIST : 0) | (wvs ? SCOPE_WAVE : 0) | (vss ? SCOPE_VSCOPE : 0);
		struct scope_surface s;
		fill_surface(&s, sd);
		if (ok && scope_accumulate_host(roi->ctx, &p, &s, &out) == SCOPE_OK) {
			if (his)
				his->w_tex_buf = hw ^ 1;
			if (wvs)
				wvs->w_tex_buf = ww ^ 1;
			if (vss) {
				vss->tex_cs[vw] = sd->colorspace;
				vss->w_tex_buf = vw ^ 1;
			}
		}
	}

	/* any further sources of the same kind: their own pass, like the reference's loop */
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
	pthread_mutex_init(&src->pipeline_mutex, NULL);
	pthread_cond_init(&src->pipeline_cond, NULL);
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
static void pipeline_thread_loop(struct b200_cm_source *src, struct b200_cm_queue_item *item)
{
	if (!(item->flags & (B200_CM_FLAG_CONVERT_RGB | B200_CM_FLAG_CONVERT_YUV)) || !item->staged)
		return;
	uint8_t *video_data = item->staged;
	struct cm_surface_data sd;
	memset(&sd, 0, sizeof(sd));
	sd.linesize = item->linesize;
	sd.width = item->width;
	sd.height = item->height;
	sd.colorspace = item->colorspace;
	if (item->flags & B200_CM_FLAG_CONVERT_RGB) {
		sd.rgb_data = video_data;
		video_data += (size_t)item->linesize * item->height;
	}
	if (item->flags & B200_CM_FLAG_CONVERT_YUV)
		sd.yuv_data = video_data;
	if (item->cb)
		item->cb(item->cb_data, &sd);
	__sync_fetch_and_add(&src->frames_processed, 1);
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
		pipeline_thread_loop(src, &src->queue[next]);
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

	/* crop rectangle (common.c:272-282): the ROI if the flag is set and the rectangle is sane
	 * and inside the frame, the whole frame otherwise */
	uint32_t x = 0, y = 0, cx = width, cy = height;
	if ((src->flags & B200_CM_FLAG_ROI) && 0 <= src->x0 && src->x0 < src->x1 && 0 <= src->y0 && src->y0 < src->y1 &&
	    (uint32_t)src->x1 <= width && (uint32_t)src->y1 <= height) {
		x = (uint32_t)src->x0;
		y = (uint32_t)src->y0;
		cx = (uint32_t)src->x1 - x;
		cy = (uint32_t)src->y1 - y;
	}
	const bool whole = cx == width && cy == height;
	/* the staged surface is cx wide (prepare_stagesurface, common.c:130-139); a whole frame keeps
	 * the caller's pitch so that it moves with one memcpy per plane */
	const uint32_t out_linesize = whole ? linesize : cx * 4u;

	struct b200_cm_queue_item *item = &src->queue[src->i_write_queue];
	const size_t plane = (size_t)out_linesize * cy;
	const size_t need = plane * ((has_rgb ? 1 : 0) + (has_yuv ? 1 : 0));
	if (item->staged_bytes < need) {
		free(item->staged);
		item->staged = malloc(need ? need : 1);
		item->staged_bytes = item->staged ? need : 0;
		if (!item->staged)
			return false;
	}
	uint8_t *dst = item->staged; /* "gs_stage_texture": RGB rows first, YUV rows below */
	const uint8_t *planes[2] = {has_rgb ? rgb : NULL, has_yuv ? yuv : NULL};
	for (int p = 0; p < 2; p++) {
		if (!planes[p])
			continue;
		if (whole) {
			memcpy(dst, planes[p], plane);
		} else {
			const uint8_t *from = planes[p] + (size_t)y * linesize + (size_t)x * 4u;
			for (uint32_t r = 0; r < cy; r++)
				memcpy(dst + (size_t)r * out_linesize, from + (size_t)r * linesize, out_linesize);
		}
		dst += plane;
	}
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
