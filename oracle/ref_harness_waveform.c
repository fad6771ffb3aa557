/* oracle/_ref harness — TEST INFRASTRUCTURE, not product code.
 * Compiles the reference's src/waveform.c unmodified (via -I$(REF)/src) and
 * exposes wvs_draw_waveform (waveform.c:220-257) and wvs_surface_cb
 * (waveform.c:272-289).  See ref_harness_histogram.c for the method. */
#include "waveform.c"

#define HARNESS_API __attribute__((visibility("default")))

/* out_buf: width*256*4 bytes (reference layout: BGRX rows, row 0 = value 255) */
HARNESS_API int ref_wvs_draw_waveform(uint32_t components, const uint8_t *rgb_data, const uint8_t *yuv_data,
				      uint32_t linesize, uint32_t width, uint32_t height, int colorspace,
				      uint8_t *out_buf)
{
	struct wvs_source src;
	memset(&src, 0, sizeof(src));
	src.components = components;
	struct cm_surface_data sd = {
		.rgb_data = (uint8_t *)rgb_data,
		.yuv_data = (uint8_t *)yuv_data,
		.linesize = linesize,
		.width = width,
		.height = height,
		.colorspace = colorspace,
	};
	wvs_draw_waveform(&src, out_buf, &sd);
	return 0;
}

HARNESS_API void *ref_wvs_new(uint32_t components)
{
	struct wvs_source *src = calloc(1, sizeof(*src));
	src->components = components;
	return src;
}

HARNESS_API void ref_wvs_free(void *state)
{
	struct wvs_source *src = state;
	free(src->tex_buf[0]);
	free(src->tex_buf[1]);
	free(src);
}

/* Returns w_tex_buf after the call; *out_width = tex_buf_width of the buffer
 * the graphics thread would read (w^1); out_buf must hold max_width*256*4. */
HARNESS_API int ref_wvs_surface_cb(void *state, const uint8_t *rgb_data, const uint8_t *yuv_data, uint32_t linesize,
				   uint32_t width, uint32_t height, int colorspace, uint8_t *out_buf,
				   uint32_t *out_width)
{
	struct wvs_source *src = state;
	struct cm_surface_data sd = {
		.rgb_data = (uint8_t *)rgb_data,
		.yuv_data = (uint8_t *)yuv_data,
		.linesize = linesize,
		.width = width,
		.height = height,
		.colorspace = colorspace,
	};
	wvs_surface_cb(src, &sd);
	int r = src->w_tex_buf ^ 1;
	*out_width = 0;
	if (src->tex_buf[r]) {
		*out_width = src->tex_buf_width[r];
		memcpy(out_buf, src->tex_buf[r], (size_t)src->tex_buf_width[r] * WV_SIZE * 4);
	}
	return src->w_tex_buf;
}
