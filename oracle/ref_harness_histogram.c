/* oracle/_ref harness — TEST INFRASTRUCTURE, not product code.
 *
 * Compiles the reference's src/histogram.c UNMODIFIED, where it lies under
 * /root/reference (found through -I$(REF)/src; nothing is copied into this
 * repo), inside this translation unit so that its file-static functions
 * his_draw_histogram (histogram.c:357-418) and his_surface_cb
 * (histogram.c:432-450) can be driven with plain host buffers.
 * libobs is replaced by oracle/mock_libobs (declarations only).
 */
#include "histogram.c"

#define HARNESS_API __attribute__((visibility("default")))

/* Run the reference accumulation + post-pass once.
 * out_buf: 4096 bytes, receives the reference's tex_buf content: float[1024]
 * after the in-place u32->float (or log) conversion (histogram.c:404-417).
 * Returns 0.  Either plane pointer may be NULL exactly as in the reference. */
HARNESS_API int ref_his_draw_histogram(uint32_t components, int level_fixed_value, int level_ratio_value,
				       int logscale, const uint8_t *rgb_data, const uint8_t *yuv_data,
				       uint32_t linesize, uint32_t width, uint32_t height, int colorspace,
				       uint8_t *out_buf, uint32_t *out_hi_max)
{
	struct his_source src;
	memset(&src, 0, sizeof(src));
	src.components = components;
	src.level_fixed_value = level_fixed_value;
	src.level_ratio_value = level_ratio_value;
	src.logscale = logscale ? true : false;
	struct cm_surface_data sd = {
		.rgb_data = (uint8_t *)rgb_data,
		.yuv_data = (uint8_t *)yuv_data,
		.linesize = linesize,
		.width = width,
		.height = height,
		.colorspace = colorspace,
	};
	his_draw_histogram(&src, out_buf, out_hi_max, &sd);
	return 0;
}

/* Drive the reference's callback (double buffer + flip semantics).
 * `state` is an opaque struct his_source* created by ref_his_new. */
HARNESS_API void *ref_his_new(uint32_t components, int level_fixed_value, int level_ratio_value, int logscale)
{
	struct his_source *src = calloc(1, sizeof(*src));
	src->components = components;
	src->level_fixed_value = level_fixed_value;
	src->level_ratio_value = level_ratio_value;
	src->logscale = logscale ? true : false;
	return src;
}

HARNESS_API void ref_his_free(void *state)
{
	struct his_source *src = state;
	free(src->tex_buf[0]);
	free(src->tex_buf[1]);
	free(src);
}

/* Returns w_tex_buf after the call; *out_buf (4096 B) / out_hi_max get the
 * buffer the graphics thread would read next (index w^1), if allocated. */
HARNESS_API int ref_his_surface_cb(void *state, const uint8_t *rgb_data, const uint8_t *yuv_data, uint32_t linesize,
				   uint32_t width, uint32_t height, int colorspace, uint8_t *out_buf,
				   uint32_t *out_hi_max)
{
	struct his_source *src = state;
	struct cm_surface_data sd = {
		.rgb_data = (uint8_t *)rgb_data,
		.yuv_data = (uint8_t *)yuv_data,
		.linesize = linesize,
		.width = width,
		.height = height,
		.colorspace = colorspace,
	};
	his_surface_cb(src, &sd);
	int r = src->w_tex_buf ^ 1;
	if (src->tex_buf[r]) {
		memcpy(out_buf, src->tex_buf[r], sizeof(float) * HI_SIZE * 4);
		memcpy(out_hi_max, src->hi_max[r], sizeof(uint32_t) * 3);
	}
	return src->w_tex_buf;
}
