/* oracle/vkr_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * C entry points of the oracle shared library (oracle/build/liboracle.so). Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it. */
#ifndef VKR_ORACLE_H
#define VKR_ORACLE_H
#include <stdint.h>
#include <stddef.h>

/* numeric values = the reference's sampling_strategies_t / mis_heuristic_t (src/main.h:45-92) */
enum { VKR_STRATEGY_DIFFUSE_ONLY = 0, VKR_STRATEGY_DIFFUSE_GGX_MIS = 1, VKR_STRATEGY_DIFFUSE_SPECULAR_SEPARATELY = 2,
	VKR_STRATEGY_DIFFUSE_SPECULAR_MIS = 3, VKR_STRATEGY_DIFFUSE_SPECULAR_RANDOM = 4 };
/* sample_polygon_technique_t (src/polygonal_light.h:30-66); the biased variant (12) is technique 11 + biased_sampling */
enum { VKR_TECHNIQUE_BASELINE = 0, VKR_TECHNIQUE_AREA_TURK = 1, VKR_TECHNIQUE_RECTANGLE_SOLID_ANGLE_URENA = 2, VKR_TECHNIQUE_SOLID_ANGLE_ARVO = 3,
	VKR_TECHNIQUE_SOLID_ANGLE = 4, VKR_TECHNIQUE_CLIPPED_SOLID_ANGLE = 5, VKR_TECHNIQUE_BILINEAR_COSINE_WARP_HART = 6, VKR_TECHNIQUE_BILINEAR_COSINE_WARP_CLIPPING_HART = 7,
	VKR_TECHNIQUE_BIQUADRATIC_COSINE_WARP_HART = 8, VKR_TECHNIQUE_BIQUADRATIC_COSINE_WARP_CLIPPING_HART = 9, VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO = 10,
	VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE = 11 };
enum { VKR_MIS_BALANCE = 0, VKR_MIS_POWER = 1, VKR_MIS_WEIGHTED = 2, VKR_MIS_OPTIMAL_CLAMPED = 3, VKR_MIS_OPTIMAL = 4 };

/* What the reference passes as -D defines (src/main.c:752-792) */
typedef struct {
	uint32_t width, height;
	uint32_t light_count;              /* POLYGONAL_LIGHT_COUNT */
	uint32_t max_light_vertex_count;   /* MAX_POLYGONAL_LIGHT_VERTEX_COUNT (3..7) */
	uint32_t min_light_vertex_count;   /* MIN_POLYGON_VERTEX_COUNT_BEFORE_CLIPPING */
	uint32_t sample_count;             /* SAMPLE_COUNT */
	uint32_t sampling_strategies;      /* SAMPLING_STRATEGIES_* */
	uint32_t mis_heuristic;            /* MIS_HEURISTIC_* */
	uint32_t biased_sampling;          /* USE_BIASED_PROJECTED_SOLID_ANGLE_SAMPLING */
	uint32_t trace_shadow_rays;        /* TRACE_SHADOW_RAYS */
	uint32_t show_polygonal_lights;    /* SHOW_POLYGONAL_LIGHTS */
	uint32_t row_begin, row_end;       /* shade rows [row_begin,row_end) only; row_end = 0 means height */
	uint32_t band_height, band_stride; /* if band_stride != 0: of those rows only the ones with (y - row_begin) % band_stride < band_height (bounded CPU samples) */
	uint32_t polygon_sampling_technique; /* SAMPLE_POLYGON_*: VKR_TECHNIQUE_* (the ctypes binding defaults to 11 = projected solid angle) */
	uint32_t error_display;            /* error_display_t (src/main.h:92-112): 0 none, 1-3 diffuse backward / backward scaled / forward, 4-6 specular */
	uint32_t output_srgb;              /* !OUTPUT_LINEAR_RGB: the shader itself converts to sRGB (UNORM swapchain); the half-bit split follows g_frame_bits in the constant block */
} vkr_oracle_config_t;

size_t vkr_oracle_light_stride(uint32_t max_light_vertex_count);
int vkr_oracle_shade(const vkr_oracle_config_t* cfg, const void* constants, const float* gbuffer,
	const uint16_t* noise, uint32_t noise_w, uint32_t noise_h, uint32_t noise_layers,
	const uint16_t* ltc0, const uint16_t* ltc1, uint32_t ltc_res, uint32_t ltc_layers,
	const float* tris, uint32_t tri_count, float* out_rgba, uint64_t* out_ray_count);
int vkr_oracle_shade_with_light_textures(const vkr_oracle_config_t* cfg, const void* constants, const float* gbuffer,
	const uint16_t* noise, uint32_t noise_w, uint32_t noise_h, uint32_t noise_layers,
	const uint16_t* ltc0, const uint16_t* ltc1, uint32_t ltc_res, uint32_t ltc_layers,
	const float* tris, uint32_t tri_count,
	uint32_t light_texture_count, const uint32_t* light_texture_dims, const uint64_t* light_texture_offsets, const float* light_texture_data,
	float* out_rgba, uint64_t* out_ray_count);
void vkr_oracle_dequantize_for_bvh(const uint32_t* quantized_positions, uint64_t vertex_count, const float* factor, const float* summand, float* out_vertices);
int vkr_oracle_visibility(uint32_t width, uint32_t height, const void* constants, const uint32_t* quantized_positions, uint64_t tri_count, uint32_t* out_visibility);
int vkr_oracle_gbuffer(uint32_t width, uint32_t height, const void* constants, const uint32_t* visibility,
	const uint32_t* quantized_positions, const uint16_t* normals_and_tex_coords, const uint8_t* material_indices,
	const float* material_params, float* out_gbuffer);
int vkr_oracle_related_work_batch(uint32_t technique, uint32_t maxv, const void* light_block, const float* position, const float* frame,
	uint32_t n, const float* random_numbers, float* out_dirs, float* out_densities, float* out_ggx_density_factor);
int vkr_oracle_error_display_batch(uint32_t technique, int biased, uint32_t maxv, uint32_t vertex_count, const float* vertices_xyz, uint32_t n, const float* rnd,
	float error_factor, float* out_errors, float* out_colors);
int vkr_oracle_gbuffer_textured(uint32_t width, uint32_t height, const void* constants, const uint32_t* visibility,
	const uint32_t* quantized_positions, const uint16_t* normals_and_tex_coords, const uint8_t* material_indices,
	uint32_t texture_count, const uint32_t* texture_dims, const uint64_t* texture_offsets, const float* texture_data, float* out_gbuffer);
void vkr_oracle_light_texture_batch(uint32_t width, uint32_t height, const float* texels, uint32_t n, const float* uv, float* out_rgba);
void vkr_oracle_texture_grad_batch(uint32_t width, uint32_t height, uint32_t mip_count, const float* texels, uint32_t n, const float* inputs, float* out_rgba);
uint32_t vkr_oracle_clip(uint32_t vertex_count, float* vertices_xyz, uint32_t maxp);
void vkr_oracle_psa_sample_batch(uint32_t vertex_count, const float* vertices_xyz, uint32_t maxp, int biased, int do_clip,
	uint32_t n, const float* random_numbers, float* out_dirs, float* out_errors, float* out_info);
void vkr_oracle_sort_network(uint32_t vertex_count, uint32_t maxp, float* vertices_xy, float* ellipses_xy);
void vkr_oracle_elementary_batch(int which, uint32_t n, const float* x, float* y);
void vkr_oracle_trace_any(const float* tris, uint32_t tri_count, uint32_t ray_count, const float* rays, uint8_t* out_bvh, uint8_t* out_brute);
int vkr_oracle_thread_count(void);
double vkr_oracle_last_shade_seconds(void);
#endif
