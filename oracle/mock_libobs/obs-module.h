/* mock libobs — TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 *
 * Just enough type, constant and macro declarations for the reference's
 * src/histogram.c, src/waveform.c and src/vectorscope.c to COMPILE unmodified
 * from /root/reference.  No libobs function is ever executed: the harness
 * (oracle/ref_harness_*.c) calls only the reference's pure-C accumulation
 * loops and *_surface_cb callbacks; every libobs function the translation
 * units mention is left undeclared (GNU C89-style implicit declaration) and
 * resolved at link time by oracle/_ref/trap_stubs.c, which abort()s if one is
 * reached.  Written from the call sites in the reference sources, not from
 * libobs headers (libobs is not installed in this image).
 */
#pragma once
#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

#define UNUSED_PARAMETER(x) ((void)(x))

/* ---- opaque handles ---- */
typedef struct obs_data obs_data_t;
typedef struct obs_source obs_source_t;
typedef struct obs_weak_source obs_weak_source_t;
typedef struct obs_property obs_property_t;
typedef struct obs_properties obs_properties_t;
typedef struct gs_effect gs_effect_t;
typedef struct gs_effect_param gs_eparam_t;
typedef struct gs_texture gs_texture_t;
typedef struct gs_texrender gs_texrender_t;
typedef struct gs_stage_surface gs_stagesurf_t;
typedef struct gs_vertex_buffer gs_vertbuffer_t;
typedef void (*obs_source_enum_proc_t)(obs_source_t *parent, obs_source_t *child, void *param);

/* ---- small math structs (only .ptr / brace-init shapes are used) ---- */
struct vec2 { union { struct { float x, y; }; float ptr[2]; }; };
struct vec3 { union { struct { float x, y, z, w; }; float ptr[4]; }; };
struct vec4 { union { struct { float x, y, z, w; }; float ptr[4]; }; };
struct matrix4 { struct vec4 x, y, z, t; };
static inline void vec2_set(struct vec2 *v, float x, float y) { v->x = x; v->y = y; }
static inline void vec3_set(struct vec3 *v, float x, float y, float z) { v->x = x; v->y = y; v->z = z; v->w = 0.0f; }

struct gs_tvertarray { size_t width; void *array; };
struct gs_vb_data {
	size_t num;
	struct vec3 *points, *normals, *tangents;
	uint32_t *colors;
	size_t num_tex;
	struct gs_tvertarray *tvarray;
};
typedef struct gs_image_file { gs_texture_t *texture; uint32_t cx, cy; bool loaded; } gs_image_file_t;

struct obs_mouse_event { uint32_t modifiers; int32_t x, y; };

/* ---- enums / flags named by the scope sources ---- */
enum { OBS_SOURCE_TYPE_INPUT, OBS_SOURCE_TYPE_FILTER };
enum { OBS_SOURCE_VIDEO = 1 << 0, OBS_SOURCE_CUSTOM_DRAW = 1 << 3, OBS_SOURCE_INTERACTION = 1 << 5,
       OBS_SOURCE_CAP_OBSOLETE = 1 << 8 };
enum { OBS_EFFECT_DEFAULT, OBS_EFFECT_SOLID };
enum { OBS_COMBO_TYPE_LIST = 2 };
enum { OBS_COMBO_FORMAT_INT = 1, OBS_COMBO_FORMAT_FLOAT = 2, OBS_COMBO_FORMAT_STRING = 3 };
enum { GS_POINTS, GS_LINES, GS_LINESTRIP, GS_TRIS, GS_TRISTRIP };
enum { GS_R8 = 3, GS_BGRX = 5, GS_RGBA32F = 11 };
enum { GS_DYNAMIC = 1 << 1 };
enum { LOG_ERROR = 100, LOG_WARNING = 200, LOG_INFO = 300, LOG_DEBUG = 400 };
enum { MOUSE_LEFT, MOUSE_MIDDLE, MOUSE_RIGHT };

/* ---- the plugin ABI struct: fields in the order the sources initialise ---- */
struct obs_source_info {
	const char *id;
	int type;
	uint32_t output_flags;
	uint32_t version;
	const char *(*get_name)(void *type_data);
	void *(*create)(obs_data_t *settings, obs_source_t *source);
	void (*destroy)(void *data);
	uint32_t (*get_width)(void *data);
	uint32_t (*get_height)(void *data);
	void (*get_defaults)(obs_data_t *settings);
	obs_properties_t *(*get_properties)(void *data);
	void (*update)(void *data, obs_data_t *settings);
	void (*video_tick)(void *data, float seconds);
	void (*video_render)(void *data, gs_effect_t *effect);
	void (*enum_active_sources)(void *data, obs_source_enum_proc_t enum_callback, void *param);
	void (*mouse_click)(void *data, const struct obs_mouse_event *event, int32_t type, bool mouse_up,
			    uint32_t click_count);
	void (*mouse_move)(void *data, const struct obs_mouse_event *event, bool mouse_leave);
	void (*mouse_wheel)(void *data, const struct obs_mouse_event *event, int x_delta, int y_delta);
};

/* pointer-returning libobs calls must not be truncated through implicit int */
void *bzalloc(size_t size);
void bfree(void *ptr);
const char *obs_module_text(const char *lookup_string);
char *obs_module_file(const char *file);
const char *obs_data_get_string(obs_data_t *data, const char *name);
long long obs_data_get_int(obs_data_t *data, const char *name);
double obs_data_get_double(obs_data_t *data, const char *name);
bool obs_data_get_bool(obs_data_t *data, const char *name);
const char *obs_source_get_name(const obs_source_t *source);
gs_effect_t *obs_get_base_effect(int effect);
gs_eparam_t *gs_effect_get_param_by_name(const gs_effect_t *effect, const char *name);
bool gs_effect_loop(gs_effect_t *effect, const char *name);
gs_texture_t *gs_texture_create(uint32_t width, uint32_t height, int color_format, uint32_t levels,
				const uint8_t **data, uint32_t flags);
gs_vertbuffer_t *gs_render_save(void);
struct gs_vb_data *gs_vertexbuffer_get_data(const gs_vertbuffer_t *vertbuffer);
obs_properties_t *obs_properties_create(void);
obs_property_t *obs_properties_get(obs_properties_t *props, const char *property);
obs_property_t *obs_properties_add_list(obs_properties_t *props, const char *name, const char *description,
					int type, int format);
obs_property_t *obs_properties_add_int(obs_properties_t *props, const char *name, const char *description,
				       int min, int max, int step);
obs_property_t *obs_properties_add_float(obs_properties_t *props, const char *name, const char *description,
					 double min, double max, double step);
obs_property_t *obs_properties_add_bool(obs_properties_t *props, const char *name, const char *description);
obs_property_t *obs_properties_add_color(obs_properties_t *props, const char *name, const char *description);

/* ---- for src/roi.c (ref_harness_roi.c): a working dynamic array, opaque proc-handler types ---- */
#define DARRAY(type) struct { type *array; size_t num; size_t capacity; }
#define da_free(v) do { free((v).array); (v).array = NULL; (v).num = (v).capacity = 0; } while (0)
#define da_push_back(v, item_ptr)                                                          \
	do {                                                                               \
		if ((v).num == (v).capacity) {                                             \
			(v).capacity = (v).capacity ? (v).capacity * 2 : 4;                \
			(v).array = realloc((v).array, (v).capacity * sizeof(*(v).array)); \
		}                                                                          \
		(v).array[(v).num++] = *(item_ptr);                                        \
	} while (0)
#define da_erase_item(v, item_ptr)                                                                     \
	do {                                                                                           \
		for (size_t i_ = 0; i_ < (v).num; i_++)                                                \
			if ((v).array[i_] == *(item_ptr)) {                                            \
				memmove(&(v).array[i_], &(v).array[i_ + 1],                            \
					((v).num - i_ - 1) * sizeof(*(v).array));                      \
				(v).num--;                                                             \
				break;                                                                 \
			}                                                                              \
	} while (0)
typedef struct calldata { void *stack; size_t size; size_t capacity; bool fixed; } calldata_t;
typedef struct proc_handler proc_handler_t;
enum { GS_BLEND_ZERO, GS_BLEND_ONE, GS_BLEND_SRCALPHA = 4, GS_BLEND_INVSRCALPHA = 5 };
enum { OBS_SOURCE_CAP_DISABLED = 1 << 10 };
proc_handler_t *obs_source_get_proc_handler(const obs_source_t *source);
obs_source_t *obs_get_source_by_name(const char *name);
gs_texture_t *gs_texrender_get_texture(const gs_texrender_t *texrender);

/* ---- for src/common.c (ref_harness_common.c; the functions themselves are FAKES defined there) ---- */
enum { GS_BGRA = 4 };
enum { GS_ZS_NONE = 0 };
enum { GS_CLEAR_COLOR = 1 };
struct obs_video_info { uint32_t base_width, base_height; int colorspace; };
static inline void vec4_zero(struct vec4 *v) { v->x = v->y = v->z = v->w = 0.0f; }
char *bstrdup(const char *str);
obs_source_t *obs_weak_source_get_source(obs_weak_source_t *weak);
obs_weak_source_t *obs_source_get_weak_source(obs_source_t *source);
obs_source_t *obs_get_output_source(uint32_t channel);
obs_source_t *obs_frontend_get_current_preview_scene(void);
bool obs_source_removed(const obs_source_t *source);
uint32_t obs_source_get_width(obs_source_t *source);
uint32_t obs_source_get_height(obs_source_t *source);
bool obs_get_video_info(struct obs_video_info *ovi);
gs_texrender_t *gs_texrender_create(int format, int zsformat);
bool gs_texrender_begin(gs_texrender_t *texrender, uint32_t cx, uint32_t cy);
gs_stagesurf_t *gs_stagesurface_create(uint32_t width, uint32_t height, int color_format);
bool gs_stagesurface_map(gs_stagesurf_t *stagesurf, uint8_t **data, uint32_t *linesize);
