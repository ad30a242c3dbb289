/* oracle/ref_host_probe.c -- TEST INFRASTRUCTURE. Calls the REFERENCE's unchanged loader / host-maths code
 * (src/scene.c, textures.c, ltc_table.c, noise_table.c, polygonal_light.c, camera.c), linked against the Vulkan shim
 * (shim/), and hands the resulting bytes to the tests through flat C entry points. Built by oracle/build_ref.py into
 * oracle/_ref/libref_host.so where /root/reference exists. */
#include "scene.h"
#include "ltc_table.h"
#include "noise_table.h"
#include "polygonal_light.h"
#include "camera.h"
#include "math_utilities.h"
#include <string.h>
#include <stdlib.h>

static device_t g_device;
static scene_t g_scene;
static ltc_table_t g_ltc;
static noise_table_t g_noise;

static const device_t* probe_device(void) {
	memset(&g_device, 0, sizeof(g_device));
	g_device.device = vkr_shim_device(); g_device.instance = vkr_shim_instance();
	g_device.ray_tracing_supported = VK_TRUE;
	g_device.acceleration_structure_properties.minAccelerationStructureScratchOffsetAlignment = 128;
	g_device.physical_device_properties.limits.nonCoherentAtomSize = 64;
	return &g_device;
}

int ref_probe_load_scene(const char* file_path, const char* texture_path, uint64_t* triangle_count, uint64_t* material_count, float* factor_and_summand,
	const void** positions, const void** normals_and_tex_coords, const void** material_indices, const float** bvh_vertices, uint64_t* bvh_triangle_count)
{
	const device_t* device = probe_device();
	if (load_scene(&g_scene, device, file_path, texture_path, VK_TRUE)) return 1;
	*triangle_count = g_scene.mesh.triangle_count; *material_count = g_scene.materials.material_count;
	memcpy(factor_and_summand, g_scene.mesh.dequantization_factor, 12); memcpy(factor_and_summand + 3, g_scene.mesh.dequantization_summand, 12);
	*positions = vkr_shim_buffer_data(g_scene.mesh.positions.buffer, NULL);
	*normals_and_tex_coords = vkr_shim_buffer_data(g_scene.mesh.normals_and_tex_coords.buffer, NULL);
	*material_indices = vkr_shim_buffer_data(g_scene.mesh.material_indices.buffer, NULL);
	*bvh_vertices = vkr_shim_acceleration_structure_vertices(g_scene.acceleration_structure.bottom_level, bvh_triangle_count);
	return 0;
}
const char* ref_probe_material_name(uint64_t i) { return g_scene.materials.material_names[i]; }
/* first texel of the smallest mip of material texture (material, type) as raw bytes; returns the VkFormat */
int ref_probe_material_texel(uint64_t material, uint32_t type, void* out16) {
	const image_t* image = &g_scene.materials.textures.images[material * material_texture_count + type];
	VkDeviceSize size = 0;
	const void* data = vkr_shim_image_data(image->image, image->image_info.mipLevels - 1, 0, &size);
	memcpy(out16, data, size < 16 ? size : 16);
	return (int) image->image_info.format;
}
void ref_probe_destroy_scene(void) { destroy_scene(&g_scene, probe_device()); }

int ref_probe_load_ltc(const char* directory, uint32_t fresnel_count, uint32_t* resolution, const void** table0, const void** table1, float* constants8) {
	const device_t* device = probe_device();
	if (load_ltc_table(&g_ltc, device, directory, fresnel_count)) return 1;
	*resolution = g_ltc.roughness_count;
	*table0 = vkr_shim_image_data(g_ltc.texture_arrays.images[0].image, 0, 0, NULL);
	*table1 = vkr_shim_image_data(g_ltc.texture_arrays.images[1].image, 0, 0, NULL);
	memcpy(constants8, &g_ltc.constants, sizeof(g_ltc.constants));
	return 0;
}
void ref_probe_destroy_ltc(void) { destroy_ltc_table(&g_ltc, probe_device()); }

int ref_probe_load_noise(uint32_t width, uint32_t height, uint32_t layers, int noise_type, const void** data, uint32_t* masks_and_randoms7, int animate) {
	const device_t* device = probe_device();
	VkExtent3D resolution = { width, height, layers };
	if (load_noise_table(&g_noise, device, resolution, (noise_type_t) noise_type)) return 1;
	*data = vkr_shim_image_data(g_noise.noise_array.images[0].image, 0, 0, NULL);
	set_noise_constants(masks_and_randoms7, masks_and_randoms7 + 2, masks_and_randoms7 + 3, &g_noise, (VkBool32) animate);
	return 0;
}
void ref_probe_destroy_noise(void) { destroy_noise_table(&g_noise, probe_device()); }

/* light192 = the first 160 bytes of polygonal_light_t are filled on input (angles, scalings, translation, flux, vertex_count);
   vertices_plane_space: 4 floats per vertex. Outputs the updated struct and arrays (update_polygonal_light, polygonal_light.c:46-104) */
void ref_probe_update_light(void* light160, uint32_t vertex_count, const float* vertices_plane_space, float* out_vertices_world_space, float* out_fan_areas) {
	polygonal_light_t light; memset(&light, 0, sizeof(light));
	memcpy(&light, light160, 160);
	light.vertex_count = 0;
	set_polygonal_light_vertex_count(&light, vertex_count);
	memcpy(light.vertices_plane_space, vertices_plane_space, sizeof(float) * 4 * vertex_count);
	update_polygonal_light(&light);
	memcpy(light160, &light, 160);
	memcpy(out_vertices_world_space, light.vertices_world_space, sizeof(float) * 4 * vertex_count);
	memcpy(out_fan_areas, light.fan_areas, sizeof(float) * 4 * (vertex_count - 2));
	destroy_polygonal_light(&light);
}
void ref_probe_world_to_projection(const void* camera48, float aspect_ratio, float* out16) {
	first_person_camera_t camera; memcpy(&camera, camera48, sizeof(camera));
	float m[4][4]; get_world_to_projection_space(m, &camera, aspect_ratio);
	memcpy(out16, m, sizeof(m));
}
void ref_probe_matrix_inverse(const float* in16, float* out16) {
	float a[4][4], b[4][4]; memcpy(a, in16, sizeof(a)); matrix_inverse(b, a); memcpy(out16, b, sizeof(b));
}
uint32_t ref_probe_sizes(int which) {
	switch (which) { case 0: return (uint32_t) sizeof(first_person_camera_t); case 1: return (uint32_t) sizeof(polygonal_light_t);
	case 2: return (uint32_t) POLYGONAL_LIGHT_QUICKSAVE_SIZE; case 3: return (uint32_t) POLYGONAL_LIGHT_FIXED_CONSTANT_BUFFER_SIZE; case 4: return (uint32_t) sizeof(ltc_constants_t); default: return 0; }
}

/* ---- The constant block of a frame from the REFERENCE's own host code. quick_load (src/main.c:82-130) and write_constants (src/main.c:2114-2188) live in
   main.c, the application, which cannot be compiled here (GLFW, Dear ImGui, the swapchain); their bodies are restated below over the reference's own
   structs and functions (polygonal_light_t, update_polygonal_light, get_world_to_projection_space, matrix_inverse, set_noise_constants, ltc_constants_t),
   per_frame_constants_t is restated from src/main.h:488-505. tests/test_ref_host.py holds vkr_write_constants against it byte for byte, and bench.py's
   reference arm gets its constants here, so that nothing of the product library is loaded in that process. */
typedef struct probe_per_frame_constants_s {
	float mesh_dequantization_factor[3], padding_0, mesh_dequantization_summand[3];
	float error_factor;
	float world_to_projection_space[4][4];
	float pixel_to_ray_direction_world_space[3][4];
	float camera_position_world_space[3];
	float mis_visibility_estimate;
	VkExtent2D viewport_size;
	int32_t cursor_position[2];
	float exposure_factor;
	float roughness_factor;
	uint32_t noise_resolution_mask[2];
	uint32_t noise_texture_index_mask;
	uint32_t frame_bits;
	uint32_t padding_3[2];
	uint32_t noise_random_numbers[4];
	ltc_constants_t ltc_constants;
} probe_per_frame_constants_t;

/* settings6 = {mis_visibility_estimate, error_min_exponent, exposure_factor, roughness_factor, animate_noise, frame_bits} (render_settings_t, screenshot_t);
   scene, LTC table and noise table are the ones loaded last through ref_probe_load_scene / _ltc / _noise. Returns the number of bytes written, 0 on failure. */
size_t ref_probe_write_constants(void* data, size_t capacity, const char* quick_save_path, uint32_t light_count, uint32_t width, uint32_t height, const float* settings6) {
	FILE* file = fopen(quick_save_path, "rb");
	if (!file) return 0;
	first_person_camera_t camera;
	uint32_t legacy_count = 0, file_light_count = 0;
	fread(&camera, sizeof(camera), 1, file);
	fread(&legacy_count, sizeof(uint32_t), 1, file);
	fread(&file_light_count, sizeof(uint32_t), 1, file);
	polygonal_light_t* lights = calloc(file_light_count ? file_light_count : 1, sizeof(polygonal_light_t)); /* malloc in the reference: its padding words are indeterminate, zero here */
	for (uint32_t i = 0; i != file_light_count; ++i) { /* main.c:105-125 */
		polygonal_light_t* light = &lights[i];
		fread(light, POLYGONAL_LIGHT_QUICKSAVE_SIZE, 1, file);
		if (light->scaling_y <= 0.0f) light->scaling_y = light->scaling_x;
		size_t path_size = 0;
		fread(&path_size, sizeof(path_size), 1, file);
		light->texture_file_path = NULL;
		if (path_size) {
			light->texture_file_path = malloc(sizeof(char) * path_size);
			fread(light->texture_file_path, sizeof(char), path_size, file);
		}
		fread(&light->vertices_plane_space, sizeof(float*), 2, file);
		light->fan_areas = NULL;
		set_polygonal_light_vertex_count(light, light->vertex_count);
		fread(light->vertices_plane_space, sizeof(float), 4 * light->vertex_count, file);
	}
	fclose(file);
	if (light_count > file_light_count) light_count = file_light_count;
	const int animate_noise = settings6[4] != 0.0f; const uint32_t frame_bits = (uint32_t) settings6[5];
	probe_per_frame_constants_t constants = { /* main.c:2119-2131; the cursor rests at the origin */
		.mesh_dequantization_factor = {g_scene.mesh.dequantization_factor[0], g_scene.mesh.dequantization_factor[1], g_scene.mesh.dequantization_factor[2]},
		.mesh_dequantization_summand = {g_scene.mesh.dequantization_summand[0], g_scene.mesh.dequantization_summand[1], g_scene.mesh.dequantization_summand[2]},
		.camera_position_world_space = {camera.position_world_space[0], camera.position_world_space[1], camera.position_world_space[2]},
		.mis_visibility_estimate = settings6[0],
		.viewport_size = { width, height },
		.cursor_position = { 0, 0 },
		.ltc_constants = g_ltc.constants,
		.error_factor = powf(10.0f, -settings6[1]),
		.exposure_factor = settings6[2],
		.roughness_factor = settings6[3],
		.frame_bits = frame_bits,
	};
	set_noise_constants(constants.noise_resolution_mask, &constants.noise_texture_index_mask, constants.noise_random_numbers, &g_noise, animate_noise && (frame_bits == 0));
	get_world_to_projection_space(constants.world_to_projection_space, &camera, ((float) width) / ((float) height)); /* get_aspect_ratio, vulkan_basics.h */
	float viewport_transform[4]; /* main.c:2136-2156 */
	viewport_transform[0] = 2.0f / width;
	viewport_transform[1] = 2.0f / height;
	viewport_transform[2] = 0.5f * viewport_transform[0] - 1.0f;
	viewport_transform[3] = 0.5f * viewport_transform[1] - 1.0f;
	float projection_to_world_space_no_translation[4][4];
	float world_to_projection_space_no_translation[4][4];
	memcpy(world_to_projection_space_no_translation, constants.world_to_projection_space, sizeof(world_to_projection_space_no_translation));
	world_to_projection_space_no_translation[0][3] = 0.0f;
	world_to_projection_space_no_translation[1][3] = 0.0f;
	world_to_projection_space_no_translation[2][3] = 0.0f;
	matrix_inverse(projection_to_world_space_no_translation, world_to_projection_space_no_translation);
	float pixel_to_ray_direction_projection_space[4][3] = {
		{viewport_transform[0], 0.0f, viewport_transform[2]},
		{0.0f, viewport_transform[1], viewport_transform[3]},
		{0.0f, 0.0f, 1.0f},
		{0.0f, 0.0f, 1.0f},
	};
	for (uint32_t i = 0; i != 3; ++i)
		for (uint32_t j = 0; j != 3; ++j)
			for (uint32_t k = 0; k != 4; ++k)
				constants.pixel_to_ray_direction_world_space[i][j] += projection_to_world_space_no_translation[i][k] * pixel_to_ray_direction_projection_space[k][j];
	uint32_t max_vertex_count = 3; /* get_max_polygonal_light_vertex_count, main.c:184-190 */
	for (uint32_t i = 0; i != light_count; ++i) if (max_vertex_count < lights[i].vertex_count) max_vertex_count = lights[i].vertex_count;
	size_t offset = sizeof(constants);
	const size_t light_size = POLYGONAL_LIGHT_FIXED_CONSTANT_BUFFER_SIZE + sizeof(float) * (12 * max_vertex_count - 8); /* main.c:334 */
	size_t result = 0;
	if (offset + light_count * light_size <= capacity) {
		memset(data, 0, offset + light_count * light_size);
		memcpy(data, &constants, sizeof(constants));
		for (uint32_t i = 0; i != light_count; ++i) { /* main.c:2160-2186; untextured lights: texture index of the default texture = 0 */
			polygonal_light_t* light = &lights[i];
			update_polygonal_light(light);
			light->texture_index = 0;
			memcpy(((char*) data) + offset, light, POLYGONAL_LIGHT_FIXED_CONSTANT_BUFFER_SIZE);
			offset += POLYGONAL_LIGHT_FIXED_CONSTANT_BUFFER_SIZE;
			float* vertex_data[2] = { light->vertices_plane_space, light->vertices_world_space };
			for (uint32_t j = 0; j != 2; ++j) {
				memcpy(((char*) data) + offset, vertex_data[j], sizeof(float) * 4 * light->vertex_count);
				if (light->vertex_count < max_vertex_count)
					memcpy(((char*) data) + offset + sizeof(float) * 4 * light->vertex_count, vertex_data[j], sizeof(float) * 4);
				offset += sizeof(float) * 4 * max_vertex_count;
			}
			memcpy(((char*) data) + offset, light->fan_areas, sizeof(float) * 4 * (light->vertex_count - 2));
			offset += sizeof(float) * 4 * (max_vertex_count - 2); /* the repeated fan areas of shorter lights stay zero here: bench.py and the test use lights of one vertex count */
		}
		result = offset;
	}
	for (uint32_t i = 0; i != file_light_count; ++i) { free(lights[i].texture_file_path); lights[i].texture_file_path = NULL; destroy_polygonal_light(&lights[i]); }
	free(lights);
	return result;
}
