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
