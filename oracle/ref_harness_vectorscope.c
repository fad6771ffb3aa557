/* oracle/_ref harness — TEST INFRASTRUCTURE, not product code.
 * Compiles the reference's src/vectorscope.c unmodified (via -I$(REF)/src)
 * and exposes vss_draw_vectorscope (vectorscope.c:217-238) and vss_surface_cb
 * (vectorscope.c:248-265).  See ref_harness_histogram.c for the method. */
#include "vectorscope.c"

#define HARNESS_API __attribute__((visibility("default")))

/* out_buf: 65536 bytes, row 0 = V 255, column = U */
HARNESS_API int ref_vss_draw_vectorscope(const uint8_t *yuv_data, uint32_t linesize, uint32_t width,
					 uint32_t height, int colorspace, uint8_t *out_buf)
{
	struct cm_surface_data sd = {
		.rgb_data = NULL,
		.yuv_data = (uint8_t *)yuv_data,
		.linesize = linesize,
		.width = width,
		.height = height,
		.colorspace = colorspace,
	};
	vss_draw_vectorscope(out_buf, &sd);
	return 0;
}

HARNESS_API void *ref_vss_new(void)
{
	return calloc(1, sizeof(struct vss_source));
}

HARNESS_API void ref_vss_free(void *state)
{
	struct vss_source *src = state;
	free(src->tex_buf[0]);
	free(src->tex_buf[1]);
	free(src);
}

/* Returns w_tex_buf after the call; out_buf/out_cs get buffer w^1 if allocated. */
HARNESS_API int ref_vss_surface_cb(void *state, const uint8_t *rgb_data, const uint8_t *yuv_data, uint32_t linesize,
				   uint32_t width, uint32_t height, int colorspace, uint8_t *out_buf, int *out_cs)
{
	struct vss_source *src = state;
	struct cm_surface_data sd = {
		.rgb_data = (uint8_t *)rgb_data,
		.yuv_data = (uint8_t *)yuv_data,
		.linesize = linesize,
		.width = width,
		.height = height,
		.colorspace = colorspace,
	};
	vss_surface_cb(src, &sd);
	int r = src->w_tex_buf ^ 1;
	if (src->tex_buf[r]) {
		memcpy(out_buf, src->tex_buf[r], VS_SIZE * VS_SIZE);
		*out_cs = src->tex_cs[r];
	}
	return src->w_tex_buf;
}
