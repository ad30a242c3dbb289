/* oracle/glsl_compat/ref_args.h -- TEST INFRASTRUCTURE. Argument block of the reference-shader entry points. */
#ifndef VKR_REF_ARGS_H
#define VKR_REF_ARGS_H
#include <stdint.h>
typedef int (*ref_occluded_hook_t)(const void* user, const float* origin, const float* dir, float tmin, float tmax);
typedef struct ref_args_s {
	uint32_t width, height, light_count, sample_count, max_light_vertex_count, material_count;
	const void* constants;
	const uint32_t* visibility;
	const uint32_t* quantized_positions; const uint16_t* normals_and_tex_coords; const uint8_t* material_indices; const float* material_params;
	const uint16_t* noise; uint32_t noise_w, noise_h, noise_layers;
	const uint16_t* ltc0; const uint16_t* ltc1; uint32_t ltc_res, ltc_layers;
	ref_occluded_hook_t occluded_hook; const void* occluded_user;
	float* out_rgba;
	/* bounded samples for the CPU baseline (bench.py): rows [row_begin, row_end) (row_end = 0: height); if band_stride != 0 only rows
	   with (y - row_begin) % band_stride < band_height. shade_seconds: wall clock of the pixel loop (out) */
	uint32_t row_begin, row_end, band_height, band_stride;
	double shade_seconds;
	/* material textures as mip chains (NULL: constant materials from material_params): 3 per material {base colour, specular, normal};
	   texture_dims = {width, height, mip_count} per texture, texture_offsets = first float of level 0 in texture_data (oracle/texture_filter.h) */
	const uint32_t* texture_dims; const uint64_t* texture_offsets; const float* texture_data;
	/* light textures (g_light_textures; the lights' texture_index is in the constant block), same conventions; count <= LIGHT_TEXTURE_COUNT = 4 */
	uint32_t light_texture_count; const uint32_t* light_texture_dims; const uint64_t* light_texture_offsets; const float* light_texture_data;
} ref_args_t;
#endif
