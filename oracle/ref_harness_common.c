/* oracle/_ref harness — TEST INFRASTRUCTURE, not product code.
 * Compiles the reference's capture core src/common.c unmodified (via -I$(REF)/src) on top of a tiny
 * SOFTWARE graphics layer, so that cm_tick / cm_render_target / the "color-monitor" worker thread
 * (common.c:223-403, 578-598) run for real: the 3-slot queue with its one-frame staging latency,
 * the drop when the worker still owns the slot, the ROI crop and the RGB-rows-then-YUV-rows layout
 * of the staged surface.  tests/test_shim_ring.py drives this and the C shim's b200_cm_* with the
 * same schedule and compares what the callbacks see.
 *
 * The fake layer implements exactly the calls that path makes: texrenders and stagesurfaces are
 * malloc'ed BGRA images, obs_source_video_render copies the current test frame, the "Draw"
 * technique copies a sub-rectangle, the ConvertRGB_YUV techniques (a GPU shader in the reference,
 * data/common.effect) write a recognisable stand-in: B, G, R inverted, alpha 255.  Built into its own
 * library (_ref/libref_common.so) because ref_harness_roi.c replaces cm_tick / cm_render_target. */
#include "common.c"

#define HARNESS_API __attribute__((visibility("default")))

/* ---- fake objects ---- */
struct gs_texture { uint32_t w, h; uint8_t *px; };
struct gs_texrender { struct gs_texture tex; };
struct gs_stage_surface { uint32_t w, h; uint8_t *px; };
struct obs_source { uint32_t w, h; const uint8_t *bgra; };

static struct obs_source g_target;
static struct gs_texrender *g_rt;      /* render target between begin and end */
static float g_translate_y;
static const char *g_technique;
static int g_loop_state;
static int g_effect_default, g_effect_common;

char *bstrdup(const char *s) { return s ? strdup(s) : NULL; }
void os_set_thread_name(const char *n) { (void)n; }
void obs_enter_graphics(void) {}
void obs_leave_graphics(void) {}
gs_effect_t *create_effect_from_module_file(const char *basename) { (void)basename; return (gs_effect_t *)&g_effect_common; }
gs_effect_t *obs_get_base_effect(int effect) { (void)effect; return (gs_effect_t *)&g_effect_default; }
gs_eparam_t *gs_effect_get_param_by_name(const gs_effect_t *effect, const char *name) { (void)name; return (gs_eparam_t *)effect; }
void gs_effect_set_texture(gs_eparam_t *param, gs_texture_t *tex) { (void)param; (void)tex; }
bool gs_effect_loop(gs_effect_t *effect, const char *name)
{
	(void)effect;
	if (!g_loop_state) {
		g_loop_state = 1;
		g_technique = name;
		return true;
	}
	g_loop_state = 0;
	return false;
}

/* the target source: found by name, never removed */
obs_source_t *obs_get_source_by_name(const char *name) { (void)name; return &g_target; }
obs_weak_source_t *obs_source_get_weak_source(obs_source_t *s) { return (obs_weak_source_t *)s; }
obs_source_t *obs_weak_source_get_source(obs_weak_source_t *w) { return (obs_source_t *)w; }
void obs_source_release(obs_source_t *s) { (void)s; }
void obs_weak_source_release(obs_weak_source_t *w) { (void)w; }
bool obs_source_removed(const obs_source_t *s) { (void)s; return false; }
const char *obs_source_get_name(const obs_source_t *s) { (void)s; return "target"; }
uint32_t obs_source_get_width(obs_source_t *s) { return s->w; }
uint32_t obs_source_get_height(obs_source_t *s) { return s->h; }
struct roi_source *roi_from_source(obs_source_t *s) { (void)s; return NULL; } /* the target is no ROI source */
void roi_register_source(struct roi_source *r, struct cm_source *c) { (void)r; (void)c; }
void roi_unregister_source(struct roi_source *r, struct cm_source *c) { (void)r; (void)c; }

/* texrender */
gs_texrender_t *gs_texrender_create(int format, int zs) { (void)format; (void)zs; return calloc(1, sizeof(struct gs_texrender)); }
void gs_texrender_destroy(gs_texrender_t *t)
{
	if (t)
		free(t->tex.px);
	free(t);
}
void gs_texrender_reset(gs_texrender_t *t) { (void)t; }
bool gs_texrender_begin(gs_texrender_t *t, uint32_t cx, uint32_t cy)
{
	if (!t || !cx || !cy)
		return false;
	if (t->tex.w != cx || t->tex.h != cy || !t->tex.px) {
		free(t->tex.px);
		t->tex.px = malloc((size_t)cx * cy * 4);
		t->tex.w = cx;
		t->tex.h = cy;
	}
	g_rt = t;
	g_translate_y = 0.0f;
	return true;
}
void gs_texrender_end(gs_texrender_t *t) { (void)t; g_rt = NULL; }
gs_texture_t *gs_texrender_get_texture(const gs_texrender_t *t) { return t && t->tex.px ? (gs_texture_t *)&t->tex : NULL; }
void gs_clear(uint32_t flags, const struct vec4 *color, float depth, uint8_t stencil)
{
	(void)flags; (void)color; (void)depth; (void)stencil;
	if (g_rt)
		memset(g_rt->tex.px, 0, (size_t)g_rt->tex.w * g_rt->tex.h * 4);
}
void gs_projection_push(void) {}
void gs_projection_pop(void) {}
void gs_ortho(float l, float r, float t, float b, float n, float f) { (void)l; (void)r; (void)t; (void)b; (void)n; (void)f; }
void gs_blend_state_push(void) {}
void gs_blend_state_pop(void) {}
void gs_blend_function(int s, int d) { (void)s; (void)d; }
void gs_matrix_translate3f(float x, float y, float z) { (void)x; (void)z; g_translate_y += y; }

/* the target draws itself 1:1 into the render target (target_scale = 1 in the tests) */
void obs_source_video_render(obs_source_t *s)
{
	if (!g_rt || !s->bgra)
		return;
	for (uint32_t y = 0; y < g_rt->tex.h && y < s->h; y++)
		memcpy(g_rt->tex.px + (size_t)y * g_rt->tex.w * 4, s->bgra + (size_t)y * s->w * 4,
		       (size_t)(g_rt->tex.w < s->w ? g_rt->tex.w : s->w) * 4);
}

/* sub-rectangle (x, y, cx, cy) of tex -> render target at (0, translate_y), through the current technique */
void gs_draw_sprite_subregion(gs_texture_t *tex, uint32_t flip, uint32_t x, uint32_t y, uint32_t cx, uint32_t cy)
{
	(void)flip;
	if (!g_rt || !tex)
		return;
	const bool convert = g_technique && strncmp(g_technique, "ConvertRGB_YUV", 14) == 0;
	const uint32_t oy = (uint32_t)g_translate_y;
	for (uint32_t r = 0; r < cy; r++) {
		if (oy + r >= g_rt->tex.h || y + r >= tex->h)
			continue;
		for (uint32_t c = 0; c < cx; c++) {
			if (c >= g_rt->tex.w || x + c >= tex->w)
				continue;
			const uint8_t *s = tex->px + ((size_t)(y + r) * tex->w + x + c) * 4;
			uint8_t *d = g_rt->tex.px + ((size_t)(oy + r) * g_rt->tex.w + c) * 4;
			if (convert) {
				d[0] = (uint8_t)~s[0];
				d[1] = (uint8_t)~s[1];
				d[2] = (uint8_t)~s[2];
				d[3] = 255;
			} else {
				memcpy(d, s, 4);
			}
		}
	}
}

/* stagesurface */
gs_stagesurf_t *gs_stagesurface_create(uint32_t w, uint32_t h, int format)
{
	(void)format;
	struct gs_stage_surface *s = calloc(1, sizeof(*s));
	s->w = w;
	s->h = h;
	s->px = calloc((size_t)w * h ? (size_t)w * h : 1, 4);
	return s;
}
void gs_stagesurface_destroy(gs_stagesurf_t *s)
{
	if (s)
		free(s->px);
	free(s);
}
void gs_stage_texture(gs_stagesurf_t *dst, gs_texture_t *src)
{
	if (!dst || !src)
		return;
	for (uint32_t y = 0; y < dst->h && y < src->h; y++)
		memcpy(dst->px + (size_t)y * dst->w * 4, src->px + (size_t)y * src->w * 4,
		       (size_t)(dst->w < src->w ? dst->w : src->w) * 4);
}
bool gs_stagesurface_map(gs_stagesurf_t *s, uint8_t **data, uint32_t *linesize)
{
	if (!s)
		return false;
	*data = s->px;
	*linesize = s->w * 4;
	return true;
}
void gs_stagesurface_unmap(gs_stagesurf_t *s) { (void)s; }

/* ---- harness API ---- */
struct refc {
	struct cm_source src;
	cm_surface_cb_t user_cb;
	void *user_data;
	volatile int in_callback;
	volatile long callbacks;
};

static void wrap_cb(void *data, struct cm_surface_data *sd)
{
	struct refc *r = data;
	r->in_callback = 1;
	if (r->user_cb)
		r->user_cb(r->user_data, sd);
	r->callbacks++;
	r->in_callback = 0;
}

HARNESS_API void *refc_new(uint32_t flags, int colorspace, uint32_t target_w, uint32_t target_h, cm_surface_cb_t cb,
			   void *cb_data)
{
	struct refc *r = calloc(1, sizeof(*r));
	cm_create(&r->src, NULL, NULL);
	r->src.flags = flags;
	r->src.colorspace = colorspace;
	r->src.target_scale = 1;
	r->src.target_name = bstrdup("target");
	g_target.w = target_w;
	g_target.h = target_h;
	g_target.bgra = NULL;
	r->user_cb = cb;
	r->user_data = cb_data;
	cm_request(&r->src, wrap_cb, r);
	return r;
}

HARNESS_API void refc_free(void *state)
{
	struct refc *r = state;
	cm_destroy(&r->src);
	free(r);
}

/* what roi_send_range leaves in the capture core for an ROI source (roi.c:494-497) + CM_FLAG_ROI */
HARNESS_API void refc_set_roi(void *state, int x0, int y0, int x1, int y1)
{
	struct refc *r = state;
	r->src.x0 = x0;
	r->src.y0 = y0;
	r->src.x1 = x1;
	r->src.y1 = y1;
	r->src.flags |= CM_FLAG_ROI;
}

HARNESS_API void refc_tick(void *state)
{
	cm_tick(&((struct refc *)state)->src, 0.0f);
}

/* video_render with `bgra` as the target's current frame; returns 1 if a slot was staged */
HARNESS_API int refc_render(void *state, const uint8_t *bgra)
{
	struct refc *r = state;
	g_target.bgra = bgra;
	const int before = r->src.i_write_queue;
	cm_render_target(&r->src);
	return r->src.i_write_queue != before;
}

HARNESS_API void refc_indices(void *state, int out[3])
{
	struct refc *r = state;
	out[0] = r->src.i_write_queue;
	out[1] = r->src.i_staging_queue;
	out[2] = r->src.i_read_queue;
}

/* nothing staged and unread, worker not inside a callback */
HARNESS_API int refc_idle(void *state)
{
	struct refc *r = state;
	pthread_mutex_lock(&r->src.pipeline_mutex);
	const int next = (r->src.i_read_queue + 1) % CM_SURFACE_QUEUE_SIZE;
	const int idle = (r->src.i_write_queue == next || r->src.i_staging_queue == next) && !r->in_callback;
	pthread_mutex_unlock(&r->src.pipeline_mutex);
	return idle;
}

HARNESS_API long refc_callbacks(void *state)
{
	return ((struct refc *)state)->callbacks;
}
