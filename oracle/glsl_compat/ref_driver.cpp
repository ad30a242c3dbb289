// oracle/glsl_compat/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// One translation unit per shader configuration (the reference bakes its settings into the shader
// as -D defines, src/main.c:752-792). It includes the REFERENCE's shading_pass.frag.glsl (after the
// mechanical syntax pass of oracle/build_ref.py: qualifiers inout/out -> references, layout(...)
// stripped, uniform block members -> globals) inside a namespace and exports one C entry point
// that feeds the uniforms/resources and runs main() for every pixel.
#include "glsl_compat.hpp"
#include "ref_args.h"
#include <omp.h>

#ifndef REF_NS
#error "build with -DREF_NS=<namespace> -DREF_ENTRY=<symbol>"
#endif

namespace glsl {
namespace REF_NS {

// <cmath> defines M_PI as a double; the shaders define their own float M_PI under #ifndef (math_constants.glsl:15-17)
#undef M_PI
#define main shader_main
#include "shading_pass.frag.glsl"
#undef main

static float rdf(const uint8_t* p, size_t off) { float f; memcpy(&f, p + off, 4); return f; }
static uint32_t rdu(const uint8_t* p, size_t off) { uint32_t u; memcpy(&u, p + off, 4); return u; }

static void set_uniforms(const ref_args_t* a) {
	const uint8_t* c = (const uint8_t*) a->constants;
	// per_frame_constants (shared_constants.glsl:20-66, std140 row_major; offsets from src/main.h:488-505)
	g_mesh_dequantization_factor = vec3(rdf(c, 0), rdf(c, 4), rdf(c, 8));
	g_mesh_dequantization_summand = vec3(rdf(c, 16), rdf(c, 20), rdf(c, 24));
	g_error_factor = rdf(c, 28);
	for (int row = 0; row != 4; ++row) for (int col = 0; col != 4; ++col) g_world_to_projection_space[col][row] = rdf(c, 32 + 4 * (4 * row + col));
	for (int row = 0; row != 3; ++row) for (int col = 0; col != 3; ++col) g_pixel_to_ray_direction_world_space[col][row] = rdf(c, 96 + 4 * (4 * row + col));
	g_camera_position_world_space = vec3(rdf(c, 144), rdf(c, 148), rdf(c, 152));
	g_mis_visibility_estimate = rdf(c, 156);
	g_viewport_size = uvec2(rdu(c, 160), rdu(c, 164));
	g_cursor_position = ivec2((int) rdu(c, 168), (int) rdu(c, 172));
	g_exposure_factor = rdf(c, 176);
	g_roughness_factor = rdf(c, 180);
	g_noise_resolution_mask = uvec2(rdu(c, 184), rdu(c, 188));
	g_noise_texture_index_mask = rdu(c, 192);
	g_frame_bits = rdu(c, 196);
	g_noise_random_numbers = uvec4(rdu(c, 208), rdu(c, 212), rdu(c, 216), rdu(c, 220));
	g_ltc_constants.fresnel_index_factor = rdf(c, 224); g_ltc_constants.fresnel_index_summand = rdf(c, 228);
	g_ltc_constants.roughness_factor = rdf(c, 232); g_ltc_constants.roughness_summand = rdf(c, 236);
	g_ltc_constants.inclination_factor = rdf(c, 240); g_ltc_constants.inclination_summand = rdf(c, 244);
	// polygonal lights (polygonal_light_utility.glsl:26-83)
	const int V = MAX_POLYGONAL_LIGHT_VERTEX_COUNT;
	const size_t stride = 160 + 16 * (size_t) V * 2 + 16 * (size_t) (V - 2);
	for (int l = 0; l != POLYGONAL_LIGHT_COUNT; ++l) {
		const uint8_t* p = c + 256 + stride * l;
		polygonal_light_t& L = g_polygonal_lights[l];
		L.rotation_angles = vec3(rdf(p, 0), rdf(p, 4), rdf(p, 8)); L.scaling_x = rdf(p, 12);
		L.translation = vec3(rdf(p, 16), rdf(p, 20), rdf(p, 24)); L.scaling_y = rdf(p, 28);
		L.radiant_flux = vec3(rdf(p, 32), rdf(p, 36), rdf(p, 40)); L.inv_scaling_x = rdf(p, 44);
		L.surface_radiance = vec3(rdf(p, 48), rdf(p, 52), rdf(p, 56)); L.inv_scaling_y = rdf(p, 60);
		L.plane = vec4(rdf(p, 64), rdf(p, 68), rdf(p, 72), rdf(p, 76));
		L.vertex_count = rdu(p, 80); L.texturing_technique = rdu(p, 84); L.texture_index = rdu(p, 88);
		for (int row = 0; row != 3; ++row) for (int col = 0; col != 3; ++col) L.rotation[col][row] = rdf(p, 96 + 4 * (4 * row + col));
		L.area = rdf(p, 144); L.rcp_area = rdf(p, 148);
		const uint8_t* vp = p + 160; const uint8_t* vw = vp + 16 * V; const uint8_t* fa = vw + 16 * V;
		for (int i = 0; i != V; ++i) {
			L.vertices_plane_space[i] = vec2(rdf(vp, 16 * i), rdf(vp, 16 * i + 4));
			L.vertices_world_space[i] = vec3(rdf(vw, 16 * i), rdf(vw, 16 * i + 4), rdf(vw, 16 * i + 8));
		}
		for (int i = 0; i != V - 2; ++i) L.fan_areas[i] = vec2(rdf(fa, 16 * i), rdf(fa, 16 * i + 4));
	}
	// resources
	g_quantized_vertex_positions.data = a->quantized_positions; g_quantized_vertex_positions.channels = 2;
	g_packed_normals_and_tex_coords.data = a->normals_and_tex_coords;
	g_material_indices.data = nullptr; g_material_indices.channels = 1;
	g_material_index_bytes = a->material_indices;
	static thread_local vkr_texture_view_t views[3 * MATERIAL_COUNT];
	for (uint32_t m = 0; m != (uint32_t) MATERIAL_COUNT && m != a->material_count; ++m) {
		const float* mp = a->material_params + 8 * m;
		sampler2D base = {{mp[0], mp[1], mp[2], 1.0f}, nullptr}, spec = {{1.0f, mp[3], mp[4], 1.0f}, nullptr}, nrm = {{mp[5], mp[6], 1.0f, 1.0f}, nullptr};
		g_material_textures[3 * m + 0] = base; g_material_textures[3 * m + 1] = spec; g_material_textures[3 * m + 2] = nrm;
		if (a->texture_data) for (uint32_t k = 0; k != 3; ++k) {
			const uint32_t i = 3 * m + k;
			views[i].width = a->texture_dims[3 * i]; views[i].height = a->texture_dims[3 * i + 1]; views[i].mip_count = a->texture_dims[3 * i + 2];
			views[i].texels = a->texture_data + a->texture_offsets[i];
			g_material_textures[i].texture = &views[i];
		}
	}
	sampler2D white = {{1.0f, 1.0f, 1.0f, 1.0f}, nullptr};
	static thread_local vkr_texture_view_t light_views[LIGHT_TEXTURE_COUNT];
	for (int i = 0; i != LIGHT_TEXTURE_COUNT; ++i) {
		g_light_textures[i] = white;
		if ((uint32_t) i < a->light_texture_count) {
			light_views[i].width = a->light_texture_dims[3 * i]; light_views[i].height = a->light_texture_dims[3 * i + 1]; light_views[i].mip_count = a->light_texture_dims[3 * i + 2];
			light_views[i].texels = a->light_texture_data + a->light_texture_offsets[i];
			g_light_textures[i].texture = &light_views[i];
		}
	}
	g_noise_table.data = a->noise; g_noise_table.w = (int) a->noise_w; g_noise_table.h = (int) a->noise_h; g_noise_table.layers = (int) a->noise_layers;
	g_ltc_tables[0].data = a->ltc0; g_ltc_tables[0].res = (int) a->ltc_res; g_ltc_tables[0].layers = (int) a->ltc_layers; g_ltc_tables[0].channels = 4;
	g_ltc_tables[1].data = a->ltc1; g_ltc_tables[1].res = (int) a->ltc_res; g_ltc_tables[1].layers = (int) a->ltc_layers; g_ltc_tables[1].channels = 2;
	g_occluded_hook = a->occluded_hook; g_occluded_user = a->occluded_user;
}

extern "C" int REF_ENTRY(ref_args_t* a) {
	if (a->light_count != POLYGONAL_LIGHT_COUNT || a->sample_count != SAMPLE_COUNT || a->max_light_vertex_count != MAX_POLYGONAL_LIGHT_VERTEX_COUNT || a->material_count > MATERIAL_COUNT || a->light_texture_count > LIGHT_TEXTURE_COUNT) return 1;
	set_uniforms(a);
	const uint32_t y0 = a->row_begin, y1 = a->row_end ? a->row_end : a->height;
	const double begin = omp_get_wtime();
	// work items are 64-pixel pieces of rows: a banded sample of a frame has fewer rows than a big host has threads
	const uint32_t pieces = (a->width + 63) / 64;
	#pragma omp parallel for collapse(2) schedule(dynamic, 1)
	for (uint32_t y = y0; y < y1; ++y) for (uint32_t piece = 0; piece < pieces; ++piece) {
		if (a->band_stride && (y - y0) % a->band_stride >= a->band_height) continue;
		const uint32_t x_end = (piece * 64 + 64 < a->width) ? piece * 64 + 64 : a->width;
		for (uint32_t x = piece * 64; x != x_end; ++x) {
			gl_FragCoord = vec4((float) x + 0.5f, (float) y + 0.5f, 0.5f, 1.0f);
			g_current_visibility = a->visibility[(size_t) y * a->width + x];
			shader_main();
			float* o = a->out_rgba + 4 * ((size_t) y * a->width + x);
			o[0] = g_out_color.x; o[1] = g_out_color.y; o[2] = g_out_color.z; o[3] = g_out_color.w;
		}
	}
	a->shade_seconds = omp_get_wtime() - begin;
	return 0;
}

} // namespace REF_NS
} // namespace glsl
