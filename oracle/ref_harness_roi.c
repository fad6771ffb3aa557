/* oracle/_ref harness — TEST INFRASTRUCTURE, not product code.
 * Compiles the reference's src/roi.c unmodified (via -I$(REF)/src) and exposes the three small
 * pieces of host logic the C shim mirrors:
 *   roi_tick / roi_target_render   frame-interleave pacing        (roi.c:266-277, 523-532)
 *   roi_send_range                 clamping of the ROI rectangle  (roi.c:478-500)
 *   roi_tick                       capture flags = OR of the consumers' flags (roi.c:533-540)
 * cm_tick and cm_render_target (src/common.c, not compiled here: all graphics) are replaced by
 * counters, which is exactly what the pacing logic is about: on which ticks they are reached. */
#include "roi.c"

#define HARNESS_API __attribute__((visibility("default")))

static int n_cm_tick, n_cm_render_target;
void cm_tick(void *data, float unused)
{
	(void)data;
	(void)unused;
	n_cm_tick++;
}
void cm_render_target(struct cm_source *src)
{
	(void)src;
	n_cm_render_target++;
}

HARNESS_API void *ref_roi_new(int n_interleave)
{
	struct roi_source *src = calloc(1, sizeof(*src));
	pthread_mutex_init(&src->sources_mutex, NULL);
	src->n_interleave = n_interleave;
	n_cm_tick = n_cm_render_target = 0;
	return src;
}

HARNESS_API void ref_roi_free(void *state)
{
	struct roi_source *src = state;
	da_free(src->sources);
	pthread_mutex_destroy(&src->sources_mutex);
	free(src);
}

/* one video_tick; returns the number of cm_tick calls so far; *out_flags = cm.flags after the tick */
HARNESS_API int ref_roi_tick(void *state, uint32_t *out_flags)
{
	struct roi_source *src = state;
	roi_tick(src, 0.0f);
	if (out_flags)
		*out_flags = src->cm.flags;
	return n_cm_tick;
}

/* one render; returns roi_target_render's result in bit 16 and the cm_render_target count below */
HARNESS_API int ref_roi_target_render(void *state)
{
	const bool r = roi_target_render(state);
	return (r ? 0x10000 : 0) | n_cm_render_target;
}

/* a consumer (scope source) with the given capture flags registers on the ROI (roi.c:314-320) */
HARNESS_API void *ref_roi_add_consumer(void *state, uint32_t flags)
{
	struct cm_source *cm = calloc(1, sizeof(*cm));
	cm->flags = flags;
	roi_register_source(state, cm);
	return cm;
}

HARNESS_API void ref_roi_remove_consumer(void *state, void *consumer)
{
	roi_unregister_source(state, consumer);
	free(consumer);
}

/* roi_send_range on a texrender of w x h with the requested rectangle; out = x0, y0, x1, y1 */
HARNESS_API void ref_roi_send_range(int x0in, int y0in, int x1in, int y1in, uint32_t w, uint32_t h, int out[4])
{
	struct roi_source src;
	memset(&src, 0, sizeof(src));
	src.cm.texrender_width = w;
	src.cm.texrender_height = h;
	src.x0in = x0in;
	src.y0in = y0in;
	src.x1in = x1in;
	src.y1in = y1in;
	roi_send_range(&src);
	out[0] = src.cm.x0;
	out[1] = src.cm.y0;
	out[2] = src.cm.x1;
	out[3] = src.cm.y1;
}
