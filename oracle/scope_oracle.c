/* scope_oracle.c — CPU ORACLE for the scope-accumulation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from it.  The shipped path
 * (obs-color-monitor_b200/csrc) never links or calls anything in oracle/.
 *
 * It restates, in plain C with no libobs, the algorithm of the reference
 * (norihiro/obs-color-monitor @ e904d82; citations relative to that tree):
 *
 *   histogram   src/histogram.c:330-418   (his_calculate_max, his_fix_max_level,
 *                                          his_draw_histogram)
 *   waveform    src/waveform.c:201-257    (inc_uint8, wvs_draw_waveform)
 *   vectorscope src/vectorscope.c:217-238 (vss_draw_vectorscope)
 *   surface     src/common.h:24-30, src/common.c:352-364 (plane placement)
 *   fan-out     src/roi.c:329-341         (one surface, every scope in turn)
 *   transform   data/common.effect:23-43  (PSConvertRGB_YUV601 / 709)
 *   colourspace src/util.c:25-41          (1 = BT.601, 2 = BT.709)
 *   display     data/vectorscope.effect:27-33, data/waveform.effect:30-39
 *
 * PARITY PIN STATUS
 *   Integer accumulation (hist / waveform / vectorscope, and the histogram
 *   post-pass): PINNED — tests/test_oracle_vs_ref.py checks every function here
 *   bit-for-bit against the reference's own loops, compiled unmodified from
 *   /root/reference into oracle/_ref/libref.so (oracle/Makefile), and the
 *   resulting vectors are committed under tests/golden/.
 *   RGB->YUV transform: PARITY UNPINNED.  In the reference it is a float pixel
 *   shader whose evaluation order, FMA contraction and float->UNORM8 rounding
 *   belong to the GPU driver; the reference holds no test or
 *   vector for it.  The definition below is the exact (infinitely precise) value
 *   of the shader's expression, which every fp32 evaluation approximates.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define NBINS 256

/* mirror of struct cm_surface_data (common.h:24-30) without the texture */
struct orc_surface {
	const uint8_t *rgb_data, *yuv_data;
	uint32_t linesize, width, height;
	int colorspace;
};

/* ------------------------------------------------------------------ */
/* RGB -> YUV, pinned definition (the one the product implements;       */
/* DESIGN.md §3 explains the choice): the EXACT value of the shader's   */
/* expression (data/common.effect:27-29,38-40) on UNORM8 inputs,        */
/* converted to UNORM8 the way a render target does (x*255 + 0.5,       */
/* truncate):                                                           */
/*   q = floor(c0*R + c1*G + c2*B + 255*off + 1/2)   in real arithmetic */
/* with R,G,B the bytes, the coefficients exactly as printed in the     */
/* effect file (six decimals), off = 1/2 - 1/256 (U), 0 (Y), 1/2 (V).   */
/* With c' = 10^6 * c this is integer arithmetic:                       */
/*   q = (c0'*R + c1'*G + c2'*B + K) / 10^6,                            */
/*   K = floor(10^6 * (255*off + 1/2)) = 127003906 (U; the .25 dropped  */
/*       by the floor can never carry), 500000 (Y), 128000000 (V).      */
/* The value is always inside [0,255] (checked by the table function),  */
/* so the shader's implicit [0,1] clamp never acts.                     */
/* Output bytes [U, Y, V, 255] (B<-z, G<-y, R<-x, A<-1).                */
/*                                                                      */
/* Two fp32 evaluations of the same expression are kept for the tests,  */
/* which report on how many of the 2^24 colours a float pipeline        */
/* deviates from the exact value (a few hundred per channel, by 1):     */
/*   "strict"     xf = (float)X/255.0f; every product and sum rounded   */
/*                separately, left to right (SURVEY.md §8(c)'s draft);   */
/*   "contracted" p = c0*rf; p = fma(c1,gf,p); p = fma(c2,bf,p);        */
/*                t = p + off; q = floor(fma(t,255,0.5)) - the mul/mad/  */
/*                mad/add sequence a shader compiler emits.             */
/* Build with -ffp-contract=off so the compiler adds no contraction of  */
/* its own.                                                             */
/* ------------------------------------------------------------------ */
struct yuv_coeffs_i {
	int32_t u[3], y[3], v[3]; /* 10^6 x coefficient, order R, G, B */
};
static const struct yuv_coeffs_i ki_bt601 = {
	{-147643, -289855, +437500},
	{+299000, +587000, +114000},
	{+437500, -366351, -71147},
};
static const struct yuv_coeffs_i ki_bt709 = {
	{-100643, -338571, +439216},
	{+212600, +715200, +72200},
	{+439216, -398941, -40273},
};
#define K_U 127003906
#define K_Y 500000
#define K_V 128000000

static inline int32_t dot3_exact(const int32_t c[3], int r, int g, int b, int32_t k)
{
	return c[0] * r + c[1] * g + c[2] * b + k; /* |.| < 2^28: no overflow */
}

struct yuv_coeffs {
	float u[3], y[3], v[3];
};

static const struct yuv_coeffs k_bt601 = {
	{-0.147643f, -0.289855f, +0.437500f},
	{+0.299000f, +0.587000f, +0.114000f},
	{+0.437500f, -0.366351f, -0.071147f},
};
static const struct yuv_coeffs k_bt709 = {
	{-0.100643f, -0.338571f, +0.439216f},
	{+0.212600f, +0.715200f, +0.072200f},
	{+0.439216f, -0.398941f, -0.040273f},
};

static inline float clamp01(float t)
{
	return fminf(fmaxf(t, 0.0f), 1.0f);
}

static inline uint8_t quantise_contracted(float t)
{
	return (uint8_t)floorf(fmaf(clamp01(t), 255.0f, 0.5f));
}

static inline uint8_t quantise_strict(float t)
{
	float s = clamp01(t) * 255.0f;
	s = s + 0.5f;
	return (uint8_t)floorf(s);
}

static inline float dot3_contracted(const float c[3], float r, float g, float b, float off)
{
	float p = c[0] * r;
	p = fmaf(c[1], g, p);
	p = fmaf(c[2], b, p);
	return p + off;
}

static inline float dot3_strict(const float c[3], float r, float g, float b, float off)
{
	float a0 = c[0] * r;
	float a1 = c[1] * g;
	float a2 = c[2] * b;
	float s = a0 + a1;
	s = s + a2;
	s = s + off;
	return s;
}

#define OFF_U (0.5f - 1.0f / 256.0f)
#define OFF_Y 0.0f
#define OFF_V 0.5f

ORC_API void orc_rgb_to_yuv_pixel(int colorspace, uint8_t r8, uint8_t g8, uint8_t b8, uint8_t out_uyv[3])
{
	const struct yuv_coeffs_i *k = colorspace == 1 ? &ki_bt601 : &ki_bt709;
	out_uyv[0] = (uint8_t)(dot3_exact(k->u, r8, g8, b8, K_U) / 1000000);
	out_uyv[1] = (uint8_t)(dot3_exact(k->y, r8, g8, b8, K_Y) / 1000000);
	out_uyv[2] = (uint8_t)(dot3_exact(k->v, r8, g8, b8, K_V) / 1000000);
}

/* Whole plane: BGRA in, [U,Y,V,255] out.  Alpha of the source is ignored
 * (the shader writes a=1, common.effect:30,41). */
ORC_API void orc_rgb_to_yuv(const uint8_t *bgra, uint32_t linesize, uint32_t width, uint32_t height,
			    int colorspace, uint8_t *yuv, uint32_t yuv_linesize)
{
	for (uint32_t y = 0; y < height; y++) {
		const uint8_t *s = bgra + (size_t)linesize * y;
		uint8_t *d = yuv + (size_t)yuv_linesize * y;
		for (uint32_t x = 0; x < width; x++, s += 4, d += 4) {
			uint8_t uyv[3];
			orc_rgb_to_yuv_pixel(colorspace, s[2], s[1], s[0], uyv);
			d[0] = uyv[0];
			d[1] = uyv[1];
			d[2] = uyv[2];
			d[3] = 255;
		}
	}
}

/* Exhaustive helper for tests: all 2^24 (r,g,b) -> packed u | y<<8 | v<<16,
 * index = r<<16 | g<<8 | b.  variant 0 = the pinned exact definition,
 * 1 = fp32 "strict", 2 = fp32 "contracted".  Returns 1 if a value ever left
 * the UNORM range (exact: S outside [0, 256e6); fp32: the [0,1] clamp changed
 * something) - it must not: the GPU kernel relies on that. */
ORC_API int orc_rgb_to_yuv_table(int colorspace, int variant, uint32_t *out /* 1<<24 entries */)
{
	const struct yuv_coeffs *k = colorspace == 1 ? &k_bt601 : &k_bt709;
	const struct yuv_coeffs_i *ki = colorspace == 1 ? &ki_bt601 : &ki_bt709;
	int out_of_range = 0;
	for (uint32_t r8 = 0; r8 < 256; r8++)
		for (uint32_t g8 = 0; g8 < 256; g8++)
			for (uint32_t b8 = 0; b8 < 256; b8++) {
				uint32_t qu, qy, qv;
				if (variant == 0) {
					const int32_t su = dot3_exact(ki->u, r8, g8, b8, K_U);
					const int32_t sy = dot3_exact(ki->y, r8, g8, b8, K_Y);
					const int32_t sv = dot3_exact(ki->v, r8, g8, b8, K_V);
					if (su < 0 || su >= 256000000 || sy < 0 || sy >= 256000000 || sv < 0 ||
					    sv >= 256000000)
						out_of_range = 1;
					qu = (uint32_t)(su / 1000000), qy = (uint32_t)(sy / 1000000),
					qv = (uint32_t)(sv / 1000000);
				} else {
					const float r = (float)r8 / 255.0f, g = (float)g8 / 255.0f,
						    b = (float)b8 / 255.0f;
					float tu, ty, tv;
					if (variant == 1) {
						tu = dot3_strict(k->u, r, g, b, OFF_U);
						ty = dot3_strict(k->y, r, g, b, OFF_Y);
						tv = dot3_strict(k->v, r, g, b, OFF_V);
						qu = quantise_strict(tu), qy = quantise_strict(ty), qv = quantise_strict(tv);
					} else {
						tu = dot3_contracted(k->u, r, g, b, OFF_U);
						ty = dot3_contracted(k->y, r, g, b, OFF_Y);
						tv = dot3_contracted(k->v, r, g, b, OFF_V);
						qu = quantise_contracted(tu), qy = quantise_contracted(ty),
						qv = quantise_contracted(tv);
					}
					if (tu < 0.0f || tu > 1.0f || ty < 0.0f || ty > 1.0f || tv < 0.0f || tv > 1.0f)
						out_of_range = 1;
				}
				out[(r8 << 16) | (g8 << 8) | b8] = qu | (qy << 8) | (qv << 16);
			}
	return out_of_range;
}

/* target_scale (src/common.c:88-90: 1..128; :249-250: the scopes' surface is target size / target_scale, integer
 * division).  In the reference the target is DRAWN into the smaller texrender (render_target_to_texrender,
 * common.c:141-168: gs_ortho over the full target size), so how a texel is filled belongs to the target's own
 * sampler; the rule pinned here is point sampling: texel (x, y) takes the source pixel that contains the texel's
 * centre, ((x + 1/2) s, (y + 1/2) s) -> (x s + s / 2, y s + s / 2) (for even s the centre lies on a pixel corner
 * and the top-left fill rule selects the pixel to its lower right).  dst is (width / s) x (height / s). */
ORC_API void orc_point_downsample(const uint8_t *src, uint32_t linesize, uint32_t width, uint32_t height,
				  uint32_t scale, uint8_t *dst, uint32_t dst_linesize)
{
	if (scale == 0)
		scale = 1;
	const uint32_t w = width / scale, h = height / scale;
	for (uint32_t y = 0; y < h; y++) {
		const uint8_t *row = src + (size_t)linesize * (y * scale + scale / 2);
		uint8_t *d = dst + (size_t)dst_linesize * y;
		for (uint32_t x = 0; x < w; x++)
			memcpy(d + 4 * x, row + 4 * (size_t)(x * scale + scale / 2), 4);
	}
}

/* src/util.c:25-41 with the OBS video-info lookup replaced by its default */
ORC_API int orc_calc_colorspace(int colorspace)
{
	if (colorspace == 1 || colorspace == 2)
		return colorspace;
	return 2;
}

/* ------------------------------------------------------------------ */
/* plane selection shared by histogram and waveform                    */
/* (histogram.c:367-373, waveform.c:228-234): RGB bits win over YUV.    */
/* ------------------------------------------------------------------ */
static const uint8_t *pick_plane(uint32_t components, const struct orc_surface *s)
{
	if (components & 0x07)
		return s->rgb_data;
	if (components & 0x70)
		return s->yuv_data;
	return NULL;
}

/* ------------------------------------------------------------------ */
/* histogram: raw counts.  dbuf = 256 x {slot0 R|V, slot1 G|Y, slot2 B|U, 0} */
/* ------------------------------------------------------------------ */
ORC_API void orc_histogram_counts(uint32_t components, const struct orc_surface *s, uint32_t dbuf[NBINS * 4])
{
	memset(dbuf, 0, sizeof(uint32_t) * NBINS * 4);
	const uint8_t *plane = pick_plane(components, s);
	if (!plane)
		return;
	const int want_b = (components & 0x11) != 0;
	const int want_g = (components & 0x22) != 0;
	const int want_r = (components & 0x44) != 0;
	for (uint32_t y = 0; y < s->height; y++) {
		const uint8_t *p = plane + (size_t)s->linesize * y;
		for (uint32_t x = 0; x < s->width; x++, p += 4) {
			if (p[3] == 0)
				continue;
			if (want_r)
				dbuf[p[2] * 4 + 0] += 1;
			if (want_g)
				dbuf[p[1] * 4 + 1] += 1;
			if (want_b)
				dbuf[p[0] * 4 + 2] += 1;
		}
	}
}

/* histogram post-pass (histogram.c:330-355, 397-417): hi_max then in-place
 * u32 -> float (linear) or log-normalised float. */
ORC_API void orc_histogram_post(uint32_t components, uint32_t width, uint32_t height, int level_fixed_value,
				int level_ratio_value, int logscale, const uint32_t dbuf[NBINS * 4],
				float out[NBINS * 4], uint32_t hi_max[3])
{
	if (level_fixed_value > 0) {
		uint32_t v = (uint32_t)level_fixed_value;
		hi_max[0] = hi_max[1] = hi_max[2] = v ? v : 1;
	} else if (level_ratio_value > 0) {
		uint32_t v = (uint32_t)((uint64_t)width * height * (uint64_t)level_ratio_value / 1000);
		hi_max[0] = hi_max[1] = hi_max[2] = v ? v : 1;
	} else {
		static const uint32_t mask[3] = {0x44, 0x22, 0x11};
		for (int j = 0; j < 3; j++) {
			uint32_t m = 1;
			if (components & mask[j])
				for (int i = 0; i < NBINS; i++)
					if (dbuf[i * 4 + j] > m)
						m = dbuf[i * 4 + j];
			hi_max[j] = m;
		}
	}

	if (logscale) {
		/* untouched slots keep the integer bit pattern (in-place reinterpretation
		 * in the reference: only enabled channels are rewritten, histogram.c:405-413) */
		memcpy(out, dbuf, sizeof(uint32_t) * NBINS * 4);
		static const uint32_t mask[3] = {0x44, 0x22, 0x11};
		for (int j = 0; j < 3; j++) {
			if (!(components & mask[j]))
				continue;
			const float scale = 1.0f / logf((float)(hi_max[j] + 1));
			for (int i = 0; i < NBINS; i++) {
				const uint32_t c = dbuf[i * 4 + j];
				out[i * 4 + j] = c ? logf((float)(c + 1)) * scale : 0;
			}
			hi_max[j] = 1;
		}
	} else {
		for (int i = 0; i < NBINS * 4; i++)
			out[i] = (float)dbuf[i];
	}
}

/* his_draw_histogram as a whole (histogram.c:357-418), for callers that want the reference's exact
 * buffer state: when `components` selects no plane, or the selected plane is NULL, the reference
 * returns right after zeroing the buffer (histogram.c:366-373) - the level pass is never reached and
 * hi_max keeps whatever the caller had in it. */
ORC_API void orc_draw_histogram(uint32_t components, int level_fixed_value, int level_ratio_value, int logscale,
				const struct orc_surface *s, float out[NBINS * 4], uint32_t hi_max[3])
{
	uint32_t dbuf[NBINS * 4];
	orc_histogram_counts(components, s, dbuf);
	if (!pick_plane(components, s)) {
		memset(out, 0, sizeof(float) * NBINS * 4);
		return;
	}
	orc_histogram_post(components, s->width, s->height, level_fixed_value, level_ratio_value, logscale, dbuf, out,
			   hi_max);
}

/* ------------------------------------------------------------------ */
/* waveform: dbuf = u8 [256][width][4], row 0 = value 255, byte 3 = 0   */
/* ------------------------------------------------------------------ */
static inline void sat_inc(uint8_t *c)
{
	if (*c != 255)
		*c += 1;
}

ORC_API void orc_waveform(uint32_t components, const struct orc_surface *s, uint8_t *dbuf)
{
	const size_t row = (size_t)s->width * 4;
	memset(dbuf, 0, row * NBINS);
	const uint8_t *plane = pick_plane(components, s);
	if (!plane)
		return;
	const int want_b = (components & 0x11) != 0;
	const int want_g = (components & 0x22) != 0;
	const int want_r = (components & 0x44) != 0;
	for (uint32_t y = 0; y < s->height; y++) {
		const uint8_t *p = plane + (size_t)s->linesize * y;
		for (uint32_t x = 0; x < s->width; x++, p += 4) {
			if (p[3] == 0)
				continue;
			uint8_t *col = dbuf + (size_t)x * 4;
			if (want_b)
				sat_inc(col + (size_t)(NBINS - 1 - p[0]) * row + 0);
			if (want_g)
				sat_inc(col + (size_t)(NBINS - 1 - p[1]) * row + 1);
			if (want_r)
				sat_inc(col + (size_t)(NBINS - 1 - p[2]) * row + 2);
		}
	}
}

/* ------------------------------------------------------------------ */
/* vectorscope: dbuf = u8 [256][256], row = 255 - V, column = U; no     */
/* alpha test; reads the YUV plane only.                                */
/* ------------------------------------------------------------------ */
ORC_API void orc_vectorscope(const struct orc_surface *s, uint8_t dbuf[NBINS * NBINS])
{
	memset(dbuf, 0, NBINS * NBINS);
	if (!s->yuv_data)
		return;
	for (uint32_t y = 0; y < s->height; y++) {
		const uint8_t *p = s->yuv_data + (size_t)s->linesize * y;
		for (uint32_t x = 0; x < s->width; x++, p += 4)
			sat_inc(dbuf + p[0] + NBINS * (255 - p[2]));
	}
}

/* ------------------------------------------------------------------ */
/* ROI fan-out (roi.c:329-341): the same surface goes to every          */
/* registered scope one after the other.  mask bit0 hist, bit1 wave,    */
/* bit2 vectorscope; NULL outputs are skipped.                          */
/* ------------------------------------------------------------------ */
ORC_API void orc_fanout(const struct orc_surface *s, uint32_t hist_components, uint32_t wave_components,
			uint32_t *hist_dbuf, uint8_t *wave_dbuf, uint8_t *vscope_dbuf)
{
	if (vscope_dbuf)
		orc_vectorscope(s, vscope_dbuf);
	if (wave_dbuf)
		orc_waveform(wave_components, s, wave_dbuf);
	if (hist_dbuf)
		orc_histogram_counts(hist_components, s, hist_dbuf);
}

/* ------------------------------------------------------------------ */
/* display mapping of an 8-bit bin image (vectorscope.effect:30-31,      */
/* waveform.effect:33-36): r = (c/255) * intensity, clamped to 1, then   */
/* stored as UNORM8 with the same rounding rule as the transform.        */
/* PARITY UNPINNED (shader + ROP), definition from SURVEY.md §8(d) cfg 3. */
/* ------------------------------------------------------------------ */
ORC_API void orc_apply_intensity(const uint8_t *bins, size_t n, int intensity, uint8_t *out)
{
	if (intensity < 1)
		intensity = 1;
	const float k = (float)intensity;
	for (size_t i = 0; i < n; i++) {
		float r = (float)bins[i] / 255.0f;
		r = r * k;
		if (r > 1.0f)
			r = 1.0f;
		out[i] = (uint8_t)floorf(fmaf(r, 255.0f, 0.5f));
	}
}
