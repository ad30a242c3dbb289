/* oracle/shading_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Scalar fp32 restatement of the reference's per-pixel shading pass
 *   src/shaders/shading_pass.frag.glsl:120-138, 151-185, 203-231, 243-323, 329-711, 721-893
 *   src/shaders/ltc_utility.glsl:58-108, brdfs.glsl:42-224, noise_utility.glsl:63-103,
 *   polygonal_light_utility.glsl:93-112, mesh_quantization.glsl:19-45
 * for the projected-solid-angle technique (SAMPLE_POLYGON_PROJECTED_SOLID_ANGLE, incl. the
 * biased variant) with all five sampling strategies, and for the related-work techniques
 * (shading_pass.frag.glsl:332-481, related_work_oracle.h) with the two strategies they support. Used ONLY by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: the reference ships no golden vectors; texture filtering,
 * ray/triangle arithmetic and transcendental precision live in the un-vendored Vulkan
 * driver and are DEFINED here (vkr_math.h, bvh_oracle.h, ltc fetch below).
 * Pinning: analytic KATs + bit-for-bit agreement with the reference GLSL compiled as C++
 * (oracle/_ref) -- see DESIGN.md "Oracle".
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -shared -fPIC (oracle/Makefile).
 */
#include "psa_oracle.h"
#include "related_work_oracle.h"
#include "bvh_oracle.h"
#include "texture_filter.h"
#include "vkr_oracle.h"
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- per-frame constants (shared_constants.glsl:20-66 <-> main.h:488-505), byte offsets */
enum {
	OFF_DEQUANT_FACTOR = 0, OFF_DEQUANT_SUMMAND = 16, OFF_ERROR_FACTOR = 28, OFF_W2P = 32, OFF_PIXEL_TO_RAY = 96,
	OFF_CAMERA = 144, OFF_MIS_VIS = 156, OFF_VIEWPORT = 160, OFF_CURSOR = 168, OFF_EXPOSURE = 176, OFF_ROUGHNESS_FACTOR = 180,
	OFF_NOISE_RES_MASK = 184, OFF_NOISE_LAYER_MASK = 192, OFF_FRAME_BITS = 196, OFF_NOISE_RANDOM = 208, OFF_LTC = 224,
	CONSTANTS_FIXED_SIZE = 256,
	/* polygonal_light_t (polygonal_light_utility.glsl:26-83), offsets inside one light block */
	L_SCALING_X = 12, L_TRANSLATION = 16, L_SCALING_Y = 28, L_SURFACE_RADIANCE = 48, L_PLANE = 64, L_VERTEX_COUNT = 80, L_TEXTURING = 84, L_ROTATION = 96, L_AREA = 144,
	L_FIXED_SIZE = 160
};

typedef struct {
	v3 surface_radiance;
	float plane[4];
	uint32_t vertex_count;
	uint32_t texturing_technique, texture_index;
	float inv_scaling_x, inv_scaling_y;
	v3 vertices_world_space[PSA_MAXP];
	/* only the related-work techniques read these (polygonal_light_utility.glsl:26-83) */
	v3 translation, rotation_cols[3];
	float scaling_x, scaling_y, area;
	v2 fan_areas[PSA_MAXP];
} light_t;

typedef struct {
	const vkr_oracle_config_t* cfg;
	const uint8_t* constants;
	float pixel_to_ray[3][4]; /* row major */
	v3 camera;
	float mis_visibility_estimate, exposure, roughness_factor, error_factor;
	uint32_t noise_res_mask[2], noise_layer_mask, noise_random[4];
	float ltc_c[6];
	light_t* lights;
	const uint16_t* noise; uint32_t noise_w, noise_h;
	const uint16_t* ltc0; const uint16_t* ltc1; uint32_t ltc_res, ltc_layers;
	const obvh_t* bvh;
	const vkr_texture_view_t* light_textures; uint32_t light_texture_count; /* g_light_textures (shading_pass.frag.glsl:60) */
	uint32_t maxp;
} ctx_t;

static float rdf(const uint8_t* p, size_t off) { float f; memcpy(&f, p + off, 4); return f; }
static uint32_t rdu(const uint8_t* p, size_t off) { uint32_t u; memcpy(&u, p + off, 4); return u; }

size_t vkr_oracle_light_stride(uint32_t max_light_vertex_count) {
	return L_FIXED_SIZE + 16 * (size_t) max_light_vertex_count * 2 + 16 * (size_t) (max_light_vertex_count - 2);
}

/* ---- shading data (brdfs.glsl:21-38) */
typedef struct {
	v3 position, normal, outgoing;
	float lambert_outgoing;
	v3 diffuse_albedo, fresnel_0;
	float roughness;
} shading_data_t;

/* ---- LTC (ltc_utility.glsl:33-50). Matrices are GLSL column-major m[col][row]. */
typedef struct {
	float world_to_shading[4][3];
	float shading_to_cosine[3][3];
	float world_to_cosine[4][3];
	float cosine_to_shading[3][3];
	float albedo, determinant;
} ltc_t;

static inline v3 mat3_mul(const float m[3][3], v3 v) {
	return mk3(
		fmaf(m[2][0], v.z, fmaf(m[1][0], v.y, m[0][0] * v.x)),
		fmaf(m[2][1], v.z, fmaf(m[1][1], v.y, m[0][1] * v.x)),
		fmaf(m[2][2], v.z, fmaf(m[1][2], v.y, m[0][2] * v.x)));
}
/* mat4x3 * vec4(v, 1) */
static inline v3 mat43_mul_point(const float m[4][3], v3 v) {
	return mk3(
		fmaf(m[3][0], 1.0f, fmaf(m[2][0], v.z, fmaf(m[1][0], v.y, m[0][0] * v.x))),
		fmaf(m[3][1], 1.0f, fmaf(m[2][1], v.z, fmaf(m[1][1], v.y, m[0][1] * v.x))),
		fmaf(m[3][2], 1.0f, fmaf(m[2][2], v.z, fmaf(m[1][2], v.y, m[0][2] * v.x))));
}
/* mat4x3 * vec4(v, 0) */
static inline v3 mat43_mul_dir(const float m[4][3], v3 v) {
	return mk3(
		fmaf(m[3][0], 0.0f, fmaf(m[2][0], v.z, fmaf(m[1][0], v.y, m[0][0] * v.x))),
		fmaf(m[3][1], 0.0f, fmaf(m[2][1], v.z, fmaf(m[1][1], v.y, m[0][1] * v.x))),
		fmaf(m[3][2], 0.0f, fmaf(m[2][2], v.z, fmaf(m[1][2], v.y, m[0][2] * v.x))));
}
/* (transpose(mat4x3) * v).xyz */
static inline v3 mat43_transpose_mul(const float m[4][3], v3 v) {
	return mk3(
		fmaf(m[0][2], v.z, fmaf(m[0][1], v.y, m[0][0] * v.x)),
		fmaf(m[1][2], v.z, fmaf(m[1][1], v.y, m[1][0] * v.x)),
		fmaf(m[2][2], v.z, fmaf(m[2][1], v.y, m[2][0] * v.x)));
}

/* Software stand-in for textureLod on a bilinear, clamp-to-edge UNORM16 2D array
   (ltc_table.c:170-177): texel = u16/65535, layer = round-to-nearest-even, weights fp32. */
static void ltc_fetch(const ctx_t* c, const uint16_t* table, uint32_t channels, float u, float v, float layer_f, float* out) {
	int32_t res = (int32_t) c->ltc_res;
	float layer_r = rintf(layer_f);
	int32_t layer = (int32_t) vkr_clamp(layer_r, 0.0f, (float) (c->ltc_layers - 1));
	float x = u * (float) res - 0.5f, y = v * (float) res - 0.5f;
	float x0f = floorf(x), y0f = floorf(y);
	float fx = x - x0f, fy = y - y0f;
	int32_t x0 = (int32_t) x0f, y0 = (int32_t) y0f, x1 = x0 + 1, y1 = y0 + 1;
	x0 = x0 < 0 ? 0 : (x0 > res - 1 ? res - 1 : x0); x1 = x1 < 0 ? 0 : (x1 > res - 1 ? res - 1 : x1);
	y0 = y0 < 0 ? 0 : (y0 > res - 1 ? res - 1 : y0); y1 = y1 < 0 ? 0 : (y1 > res - 1 ? res - 1 : y1);
	const uint16_t* base = table + (size_t) layer * res * res * channels;
	for (uint32_t ch = 0; ch != channels; ++ch) {
		float t00 = (float) base[((size_t) y0 * res + x0) * channels + ch] / 65535.0f;
		float t10 = (float) base[((size_t) y0 * res + x1) * channels + ch] / 65535.0f;
		float t01 = (float) base[((size_t) y1 * res + x0) * channels + ch] / 65535.0f;
		float t11 = (float) base[((size_t) y1 * res + x1) * channels + ch] / 65535.0f;
		float a = fmaf(fx, t10 - t00, t00);
		float b = fmaf(fx, t11 - t01, t01);
		out[ch] = fmaf(fy, b - a, a);
	}
}

/* ltc_utility.glsl:58-91 */
static void get_ltc_coefficients(ltc_t* ltc, const ctx_t* c, float fresnel_0, float roughness, v3 position, v3 normal, v3 outgoing) {
	float normal_dot_outgoing = dot3(normal, outgoing);
	float inclination = vkr_acos01(vkr_clamp(normal_dot_outgoing, 0.0f, 1.0f));
	float tu = fmaf(sqrtf(vkr_clamp(roughness, 0.0f, 1.0f)), c->ltc_c[2], c->ltc_c[3]);
	float tv = fmaf(inclination, c->ltc_c[4], c->ltc_c[5]);
	float tw = fmaf(vkr_clamp(fresnel_0, 0.0f, 1.0f), c->ltc_c[0], c->ltc_c[1]);
	float d0[4], d1[2];
	ltc_fetch(c, c->ltc0, 4, tu, tv, tw, d0);
	ltc_fetch(c, c->ltc1, 2, tu, tv, tw, d1);
	float (*s2c)[3] = ltc->shading_to_cosine;
	s2c[0][0] = d0[0]; s2c[0][1] = 0.0f; s2c[0][2] = -d0[1];
	s2c[1][0] = 0.0f; s2c[1][1] = d0[2]; s2c[1][2] = 0.0f;
	s2c[2][0] = d0[3]; s2c[2][1] = 0.0f; s2c[2][2] = d1[0];
	ltc->albedo = d1[1];
	float determinant_2x2 = d0[0] * d1[0] + d0[1] * d0[3];
	ltc->determinant = d0[2] * determinant_2x2;
	float inv_determinant_2x2 = 1.0f / determinant_2x2;
	float (*c2s)[3] = ltc->cosine_to_shading;
	c2s[0][0] = d1[0] * inv_determinant_2x2; c2s[0][1] = 0.0f; c2s[0][2] = d0[1] * inv_determinant_2x2;
	c2s[1][0] = 0.0f; c2s[1][1] = 1.0f / d0[2]; c2s[1][2] = 0.0f;
	c2s[2][0] = -d0[3] * inv_determinant_2x2; c2s[2][1] = 0.0f; c2s[2][2] = d0[0] * inv_determinant_2x2;
	v3 x_axis = normalize3(mk3(
		fmaf(-normal_dot_outgoing, normal.x, outgoing.x),
		fmaf(-normal_dot_outgoing, normal.y, outgoing.y),
		fmaf(-normal_dot_outgoing, normal.z, outgoing.z)));
	v3 y_axis = cross3(normal, x_axis);
	/* rotation = transpose(mat3(x_axis, y_axis, normal)): column j = (x_axis[j], y_axis[j], normal[j]) */
	float (*w2s)[3] = ltc->world_to_shading;
	w2s[0][0] = x_axis.x; w2s[0][1] = y_axis.x; w2s[0][2] = normal.x;
	w2s[1][0] = x_axis.y; w2s[1][1] = y_axis.y; w2s[1][2] = normal.y;
	w2s[2][0] = x_axis.z; w2s[2][1] = y_axis.z; w2s[2][2] = normal.z;
	/* -rotation * position */
	w2s[3][0] = fmaf(-w2s[2][0], position.z, fmaf(-w2s[1][0], position.y, -w2s[0][0] * position.x));
	w2s[3][1] = fmaf(-w2s[2][1], position.z, fmaf(-w2s[1][1], position.y, -w2s[0][1] * position.x));
	w2s[3][2] = fmaf(-w2s[2][2], position.z, fmaf(-w2s[1][2], position.y, -w2s[0][2] * position.x));
	/* world_to_cosine = shading_to_cosine * world_to_shading, column by column */
	for (int j = 0; j != 4; ++j) {
		v3 col = mat3_mul(ltc->shading_to_cosine, mk3(w2s[j][0], w2s[j][1], w2s[j][2]));
		ltc->world_to_cosine[j][0] = col.x; ltc->world_to_cosine[j][1] = col.y; ltc->world_to_cosine[j][2] = col.z;
	}
}

/* ltc_utility.glsl:103-108 */
static float evaluate_ltc_density(const ltc_t* ltc, v3 dir_shading_space, float rcp_projected_solid_angle) {
	v3 dir_cosine_space = mat3_mul(ltc->shading_to_cosine, dir_shading_space);
	float l2 = dot3(dir_cosine_space, dir_cosine_space);
	float ltc_density = vkr_max(0.0f, dir_cosine_space.z) * ltc->determinant / (l2 * l2);
	return ltc_density * rcp_projected_solid_angle;
}

/* ---- noise (noise_utility.glsl:21-103) */
typedef struct {
	float noise[4];
	uint32_t available, pixel[2], sample_index;
} noise_accessor_t;

static void get_noise_sample(float out[4], const ctx_t* c, const uint32_t pixel[2], uint32_t sample_index) {
	uint32_t r[4];
	if (sample_index & 2) { r[0] = c->noise_random[2]; r[1] = c->noise_random[3]; r[2] = c->noise_random[0]; r[3] = c->noise_random[1]; }
	else { r[0] = c->noise_random[0]; r[1] = c->noise_random[1]; r[2] = c->noise_random[2]; r[3] = c->noise_random[3]; }
	if (sample_index & 1) { r[0] = r[1]; r[1] = r[2]; r[2] = r[3]; }
	uint32_t shift = (sample_index & 124) >> 2;
	uint32_t ox = r[0] >> shift, oy = r[1] >> shift;
	uint32_t layer = (r[2] + sample_index) & c->noise_layer_mask;
	uint32_t x = (pixel[0] + ox) & c->noise_res_mask[0];
	uint32_t y = (pixel[1] + oy) & c->noise_res_mask[1];
	const uint16_t* t = c->noise + (((size_t) layer * c->noise_h + y) * c->noise_w + x) * 4;
	for (int k = 0; k != 4; ++k) out[k] = (float) t[k] / 65535.0f; /* RGBA16_UNORM texel fetch */
}
static v2 get_noise_2(noise_accessor_t* a, const ctx_t* c) {
	if (a->available <= 1) {
		get_noise_sample(a->noise, c, a->pixel, a->sample_index);
		a->available = 4;
		++a->sample_index;
	}
	a->available -= 2;
	v2 result = mk2(a->noise[0], a->noise[1]);
	a->noise[0] = a->noise[2]; a->noise[1] = a->noise[3];
	return result;
}

/* ---- BRDF (brdfs.glsl:42-88) */
static inline float schlick_scalar(float f0, float f90, float cos_theta) {
	float flipped = 1.0f - cos_theta;
	float flipped_squared = flipped * flipped;
	return f0 + (f90 - f0) * (flipped_squared * flipped * flipped_squared);
}
static v3 evaluate_brdf(const shading_data_t* data, v3 incoming, int diffuse, int specular) {
	v3 half_vector = normalize3(add3(incoming, data->outgoing));
	float lambert_incoming = dot3(data->normal, incoming);
	float outgoing_dot_half = dot3(data->outgoing, half_vector);
	v3 brdf = mk3(0.0f, 0.0f, 0.0f);
	if (diffuse) {
		float fresnel_90 = fmaf(outgoing_dot_half * outgoing_dot_half, 2.0f * data->roughness, 0.5f);
		float fresnel_product = schlick_scalar(1.0f, fresnel_90, data->lambert_outgoing) * schlick_scalar(1.0f, fresnel_90, lambert_incoming);
		brdf = add3(brdf, scale3(data->diffuse_albedo, fresnel_product));
	}
	if (specular) {
		float normal_dot_half = dot3(data->normal, half_vector);
		float roughness_squared = data->roughness * data->roughness;
		float ggx = fmaf(fmaf(normal_dot_half, roughness_squared, -normal_dot_half), normal_dot_half, 1.0f);
		ggx = roughness_squared / (ggx * ggx);
		float masking = lambert_incoming * sqrtf(fmaf(fmaf(-data->lambert_outgoing, roughness_squared, data->lambert_outgoing), data->lambert_outgoing, roughness_squared));
		float shadowing = data->lambert_outgoing * sqrtf(fmaf(fmaf(-lambert_incoming, roughness_squared, lambert_incoming), lambert_incoming, roughness_squared));
		float smith = 0.5f / (masking + shadowing);
		float ct = vkr_clamp(outgoing_dot_half, 0.0f, 1.0f);
		float gs = ggx * smith;
		brdf.x += gs * schlick_scalar(data->fresnel_0.x, 1.0f, ct);
		brdf.y += gs * schlick_scalar(data->fresnel_0.y, 1.0f, ct);
		brdf.z += gs * schlick_scalar(data->fresnel_0.z, 1.0f, ct);
	}
	return scale3(brdf, VKR_INV_PI);
}

/* brdfs.glsl:127-224 (GGX VNDF sampling, only for SAMPLING_STRATEGIES_DIFFUSE_GGX_MIS) */
static v3 sample_ggx_vndf(v3 outgoing_ss, float rx, float ry, v2 rnd) {
	v3 e2 = normalize3(mk3(rx * outgoing_ss.x, ry * outgoing_ss.y, 1.0f * outgoing_ss.z));
	float length_sq = dot2(mk2(e2.x, e2.y), mk2(e2.x, e2.y));
	float rs = vkr_rsqrt(length_sq);
	v3 e0 = mk3(-e2.y * rs, e2.x * rs, 0.0f * rs);
	if (length_sq <= 0.0f) e0 = mk3(1.0f, 0.0f, 0.0f);
	v3 e1 = cross3(e2, e0);
	float radius = sqrtf(rnd.x);
	float azimuth = (2.0f * VKR_PI) * rnd.y;
	v2 disk = mk2(radius * vkr_cos(azimuth), radius * vkr_sin(azimuth));
	v3 s;
	s.x = disk.x;
	float lerp_factor = fmaf(0.5f, e2.z, 0.5f);
	/* mix(x, y, a) = x*(1-a) + y*a */
	float sx = sqrtf(fmaf(-disk.x, disk.x, 1.0f));
	s.y = sx * (1.0f - lerp_factor) + disk.y * lerp_factor;
	s.z = sqrtf(vkr_max(0.0f, 1.0f - dot2(mk2(s.x, s.y), mk2(s.x, s.y))));
	/* ellipse_to_hemi * s with columns e0,e1,e2 */
	v3 h = mk3(
		fmaf(e2.x, s.z, fmaf(e1.x, s.y, e0.x * s.x)),
		fmaf(e2.y, s.z, fmaf(e1.y, s.y, e0.y * s.x)),
		fmaf(e2.z, s.z, fmaf(e1.z, s.y, e0.z * s.x)));
	return normalize3(mk3(rx * h.x, ry * h.y, 1.0f * h.z));
}
static float ggx_visible_normal_density(float outgoing_dot_normal, float microfacet_dot_normal, float microfacet_dot_outgoing, float roughness) {
	float roughness_squared = roughness * roughness;
	float ggx = fmaf(fmaf(microfacet_dot_normal, roughness_squared, -microfacet_dot_normal), microfacet_dot_normal, 1.0f);
	ggx = roughness_squared / (ggx * ggx);
	ggx *= VKR_INV_PI;
	float masking_over_out_z = sqrtf(fmaf(fmaf(-outgoing_dot_normal, roughness_squared, outgoing_dot_normal), outgoing_dot_normal, roughness_squared));
	masking_over_out_z = 2.0f / (outgoing_dot_normal + masking_over_out_z);
	return masking_over_out_z * microfacet_dot_outgoing * ggx;
}
static v3 sample_ggx_reflected_direction(float* out_density, v3 outgoing_ss, float roughness, v2 rnd) {
	v3 micro_normal = sample_ggx_vndf(outgoing_ss, roughness, roughness, rnd);
	float micro_dot_out = dot3(micro_normal, outgoing_ss);
	float density = ggx_visible_normal_density(outgoing_ss.z, micro_normal.z, micro_dot_out, roughness);
	float two = 2.0f * micro_dot_out;
	v3 incoming = mk3(fmaf(two, micro_normal.x, -outgoing_ss.x), fmaf(two, micro_normal.y, -outgoing_ss.y), fmaf(two, micro_normal.z, -outgoing_ss.z));
	density /= 4.0f * micro_dot_out;
	*out_density = density;
	return incoming;
}
static float get_ggx_reflected_direction_density(float outgoing_dot_normal, v3 outgoing_dir, v3 incoming_dir, v3 surface_normal, float roughness) {
	v3 micro_normal = normalize3(add3(outgoing_dir, incoming_dir));
	float micro_dot_out = dot3(micro_normal, outgoing_dir);
	float micro_dot_normal = dot3(micro_normal, surface_normal);
	float density = ggx_visible_normal_density(outgoing_dot_normal, micro_dot_normal, micro_dot_out, roughness);
	density /= 4.0f * micro_dot_out;
	return density;
}

/* polygonal_light_utility.glsl:93-112 */
static int polygonal_light_ray_intersection(const light_t* light, uint32_t max_light_vertices, v3 ray_origin, v3 ray_end_xyz, float ray_end_w) {
	float d0 = fmaf(light->plane[3], 1.0f, fmaf(light->plane[2], ray_origin.z, fmaf(light->plane[1], ray_origin.y, light->plane[0] * ray_origin.x)));
	float d1 = fmaf(light->plane[3], ray_end_w, fmaf(light->plane[2], ray_end_xyz.z, fmaf(light->plane[1], ray_end_xyz.y, light->plane[0] * ray_end_xyz.x)));
	if (d0 * d1 > 0.0f) return 0;
	v3 ray_dir = mk3(ray_end_xyz.x - ray_end_w * ray_origin.x, ray_end_xyz.y - ray_end_w * ray_origin.y, ray_end_xyz.z - ray_end_w * ray_origin.z);
	float previous_sign = 0.0f;
	int result = 1;
	for (uint32_t i = 0; i != max_light_vertices; ++i) {
		float sign = det3(ray_dir, sub3(light->vertices_world_space[i], ray_origin), sub3(light->vertices_world_space[(i + 1) % max_light_vertices], ray_origin));
		result = result && ((i >= 3 && i >= light->vertex_count) || previous_sign * sign >= 0.0f);
		previous_sign = sign;
	}
	return result;
}

/* shading_pass.frag.glsl:120-138 */
static void get_polygon_visibility(int* visibility, v3 sampled_dir, v3 shading_position, const light_t* light, const ctx_t* c, uint64_t* ray_count) {
	if (!c->cfg->trace_shadow_rays) return;
	if (*visibility) {
		float num = fmaf(light->plane[3], 1.0f, fmaf(light->plane[2], shading_position.z, fmaf(light->plane[1], shading_position.y, light->plane[0] * shading_position.x)));
		float den = dot3(sampled_dir, mk3(light->plane[0], light->plane[1], light->plane[2]));
		float max_t = -num / den;
		float min_t = 1.0e-3f;
		++*ray_count;
		*visibility = !obvh_occluded(c->bvh, shading_position, sampled_dir, min_t, max_t);
	}
}

/* shading_pass.frag.glsl:151-185. polygon_texturing_technique_t (src/polygonal_light.h): 0 none, 1 area, 2 portal, 3 IES profile */
static v3 get_polygon_radiance(v3 sampled_dir, v3 shading_position, const light_t* light, const ctx_t* c) {
	v3 radiance = light->surface_radiance;
	const uint32_t technique = light->texturing_technique;
	if (technique != 0) {
		v2 tex_coord;
		if (technique == 1) {
			/* intersect the ray with the plane of the light, go to plane space */
			float num = fmaf(light->plane[3], 1.0f, fmaf(light->plane[2], shading_position.z, fmaf(light->plane[1], shading_position.y, light->plane[0] * shading_position.x)));
			float intersection_t = -num / dot3(sampled_dir, mk3(light->plane[0], light->plane[1], light->plane[2]));
			v3 intersection = add3(shading_position, scale3(sampled_dir, intersection_t));
			intersection = sub3(intersection, light->translation);
			tex_coord = mk2(dot3(light->rotation_cols[0], intersection), dot3(light->rotation_cols[1], intersection));
			tex_coord = mk2(tex_coord.x * light->inv_scaling_x, tex_coord.y * light->inv_scaling_y);
		}
		else {
			v3 lookup_dir;
			if (technique == 3) {
				lookup_dir = mk3(dot3(light->rotation_cols[0], sampled_dir), dot3(light->rotation_cols[1], sampled_dir), dot3(light->rotation_cols[2], sampled_dir));
				radiance = scale3(radiance, 1.0f / fabsf(lookup_dir.z)); /* IES profiles include the cosine already */
			}
			else
				lookup_dir = mk3(-sampled_dir.x, sampled_dir.y, sampled_dir.z);
			tex_coord.x = vkr_atan2(lookup_dir.y, lookup_dir.x) * (0.5f * VKR_INV_PI);
			tex_coord.y = vkr_acos(lookup_dir.z) * VKR_INV_PI;
		}
		float texel[4];
		vkr_texture_bilinear_repeat_clamp(texel, &c->light_textures[light->texture_index], tex_coord.x, tex_coord.y);
		radiance = mk3(radiance.x * texel[0], radiance.y * texel[1], radiance.z * texel[2]);
	}
	return radiance;
}

/* shading_pass.frag.glsl:203-231 */
static v3 radiance_visibility_brdf_product(float* out_lambert, int* out_visibility, v3 sampled_dir, const shading_data_t* sd, const light_t* light, int diffuse, int specular, const ctx_t* c, uint64_t* ray_count) {
	float lambert = dot3(sd->normal, sampled_dir);
	int visibility = lambert > 0.0f;
	get_polygon_visibility(&visibility, sampled_dir, sd->position, light, c, ray_count);
	if (out_lambert) *out_lambert = lambert;
	if (out_visibility) *out_visibility = visibility;
	if (visibility) {
		v3 radiance = get_polygon_radiance(sampled_dir, sd->position, light, c);
		v3 brdf = evaluate_brdf(sd, sampled_dir, diffuse, specular);
		return mk3(radiance.x * brdf.x, radiance.y * brdf.y, radiance.z * brdf.z);
	}
	return mk3(0.0f, 0.0f, 0.0f);
}

/* shading_pass.frag.glsl:243-252 */
static float get_mis_weight_over_density(float sampled_density, float other_density, int heuristic) {
	if (heuristic == VKR_MIS_BALANCE) return 1.0f / (sampled_density + other_density);
	if (heuristic == VKR_MIS_POWER) return sampled_density / (sampled_density * sampled_density + other_density * other_density);
	return 0.0f;
}

/* shading_pass.frag.glsl:270-293 */
static v3 get_mis_estimate(v3 integrand, v3 sampled_weight, float sampled_density, v3 other_weight, float other_density, float visibility_estimate, int heuristic) {
	if (heuristic == VKR_MIS_WEIGHTED) {
		v3 weighted_sum = mk3(sampled_weight.x * sampled_density + other_weight.x * other_density, sampled_weight.y * sampled_density + other_weight.y * other_density, sampled_weight.z * sampled_density + other_weight.z * other_density);
		return mk3((sampled_weight.x * integrand.x) / weighted_sum.x, (sampled_weight.y * integrand.y) / weighted_sum.y, (sampled_weight.z * integrand.z) / weighted_sum.z);
	}
	if (heuristic == VKR_MIS_OPTIMAL_CLAMPED || heuristic == VKR_MIS_OPTIMAL) {
		float balance = 1.0f / (sampled_density + other_density);
		v3 weighted_sum = mk3(sampled_weight.x * sampled_density + other_weight.x * other_density, sampled_weight.y * sampled_density + other_weight.y * other_density, sampled_weight.z * sampled_density + other_weight.z * other_density);
		if (heuristic == VKR_MIS_OPTIMAL_CLAMPED) {
			v3 wwod = mk3(sampled_weight.x / weighted_sum.x, sampled_weight.y / weighted_sum.y, sampled_weight.z / weighted_sum.z);
			float mixed = fmaf(-visibility_estimate, balance, balance);
			v3 m = mk3(fmaf(visibility_estimate, wwod.x, mixed), fmaf(visibility_estimate, wwod.y, mixed), fmaf(visibility_estimate, wwod.z, mixed));
			return mk3(m.x * integrand.x, m.y * integrand.y, m.z * integrand.z);
		}
		return mk3(
			visibility_estimate * sampled_weight.x + balance * (integrand.x - visibility_estimate * weighted_sum.x),
			visibility_estimate * sampled_weight.y + balance * (integrand.y - visibility_estimate * weighted_sum.y),
			visibility_estimate * sampled_weight.z + balance * (integrand.z - visibility_estimate * weighted_sum.z));
	}
	float w = get_mis_weight_over_density(sampled_density, other_density, heuristic);
	return scale3(integrand, w);
}

/* shading_pass.frag.glsl:305-323 */
static v3 get_polygonal_light_mis_estimate(v3 sampled_dir, float sampled_density, const shading_data_t* sd, const light_t* light, const ctx_t* c, uint64_t* ray_count) {
	float lambert;
	v3 rtb = radiance_visibility_brdf_product(&lambert, NULL, sampled_dir, sd, light, 1, 1, c, ray_count);
	if (c->cfg->sampling_strategies == VKR_STRATEGY_DIFFUSE_ONLY)
		return (sampled_density > 0.0f) ? scale3(rtb, lambert / sampled_density) : mk3(0.0f, 0.0f, 0.0f);
	if (c->cfg->sampling_strategies == VKR_STRATEGY_DIFFUSE_GGX_MIS) {
		float ggx_density = get_ggx_reflected_direction_density(sd->lambert_outgoing, sd->outgoing, sampled_dir, sd->normal, sd->roughness);
		float w = get_mis_weight_over_density(sampled_density, ggx_density, c->cfg->mis_heuristic);
		return mk3(rtb.x * lambert * w, rtb.y * lambert * w, rtb.z * lambert * w);
	}
	return mk3(0.0f, 0.0f, 0.0f);
}

/* shading_pass.frag.glsl:80-115: error magnitude -> matplotlib's tab20b colours (the table holds linear Rec. 709 values), one hue per power of ten.
   pow / log2 follow the arithmetic contract (vkr_math.h); an index outside the table (only a NaN error gets there; undefined in GLSL) selects the first colour */
static v3 error_to_color(float error, const ctx_t* c) {
	static const float tab20b_colors[20][3] = {
		{0.04092f, 0.04374f, 0.19120f}, {0.08438f, 0.08866f, 0.36625f}, {0.14703f, 0.15593f, 0.62396f}, {0.33245f, 0.34191f, 0.73046f},
		{0.12477f, 0.19120f, 0.04092f}, {0.26225f, 0.36131f, 0.08438f}, {0.46208f, 0.62396f, 0.14703f}, {0.61721f, 0.70838f, 0.33245f},
		{0.26225f, 0.15293f, 0.03071f}, {0.50888f, 0.34191f, 0.04092f}, {0.79910f, 0.49102f, 0.08438f}, {0.79910f, 0.59720f, 0.29614f},
		{0.23074f, 0.04519f, 0.04092f}, {0.41789f, 0.06663f, 0.06848f}, {0.67244f, 0.11954f, 0.14703f}, {0.79910f, 0.30499f, 0.33245f},
		{0.19807f, 0.05286f, 0.17144f}, {0.37626f, 0.08228f, 0.29614f}, {0.61721f, 0.15293f, 0.50888f}, {0.73046f, 0.34191f, 0.67244f},
	};
	const float min_exponent = 0.0f, max_exponent = 5.0f;
	const float min_error = vkr_pow(10.0f, min_exponent);
	const float max_error = vkr_pow(10.0f, max_exponent - 0.01f);
	const float color_count = 20.0f;
	error = vkr_clamp(fabsf(c->error_factor * error), min_error, max_error);
	float color_index = fmaf(vkr_log2(error), color_count / ((max_exponent - min_exponent) * vkr_log2(10.0f)), color_count * -min_exponent / (max_exponent - min_exponent));
	int index = (color_index >= 0.0f && color_index < 20.0f) ? (int) color_index : 0;
	return mk3(tab20b_colors[index][0], tab20b_colors[index][1], tab20b_colors[index][2]);
}
static v3 error_display_color(float error, const ctx_t* c) { /* "return error_to_color(error) / g_exposure_factor;" (:472, 493, 553, 560) */
	v3 color = error_to_color(error, c);
	return mk3(color.x / c->exposure, color.y / c->exposure, color.z / c->exposure);
}
static inline float error_component(v3 e, uint32_t error_display) { uint32_t i = (error_display - 1) % 3; return (i == 0) ? e.x : ((i == 1) ? e.y : e.z); }

/* shading_pass.frag.glsl:676-709 for the related-work techniques: GGX sampling with MIS against the polygon density.
   polygon_density_is_constant: every technique except our projected solid angle sampling passes density_factor as is (:702) */
/* :676-709. `result` is the running sum of the light: the shader keeps adding to the variable that already holds the samples of the polygon
   sampling loop (floating-point addition is not associative; tools/fuzz_parity.py found the frames where a separate partial sum shows) */
static v3 ggx_mis_samples(v3 result, float density_factor, int density_times_lambert, const shading_data_t* sd, const ltc_t* ltc, const light_t* light, noise_accessor_t* accessor, const ctx_t* c, uint64_t* ray_count) {
	const vkr_oracle_config_t* cfg = c->cfg;
	v3 outgoing_ss = mat43_mul_dir(ltc->world_to_shading, sd->outgoing);
	outgoing_ss.y = 0.0f;
	for (uint32_t s = 0; s != cfg->sample_count; ++s) {
		float ggx_density;
		v3 dir_ss = sample_ggx_reflected_direction(&ggx_density, outgoing_ss, sd->roughness, get_noise_2(accessor, c));
		v3 dir_ws = mat43_transpose_mul(ltc->world_to_shading, dir_ss);
		if (dir_ss.z > 0.0f && polygonal_light_ray_intersection(light, cfg->max_light_vertex_count, sd->position, dir_ws, 0.0f)) {
			float lambert;
			v3 rtb = radiance_visibility_brdf_product(&lambert, NULL, dir_ws, sd, light, 1, 1, c, ray_count);
			float polygon_density = density_times_lambert ? (lambert * density_factor) : density_factor;
			float w = get_mis_weight_over_density(ggx_density, polygon_density, cfg->mis_heuristic);
			result.x += rtb.x * lambert * w; result.y += rtb.y * lambert * w; result.z += rtb.z * lambert * w;
		}
	}
	return result;
}

/* shading_pass.frag.glsl:332-481: the related-work sampling techniques (SAMPLE_POLYGON_BASELINE .. SAMPLE_POLYGON_PROJECTED_SOLID_ANGLE_ARVO)
   as "prepare once per (pixel, light), then sample": what each #elif branch of evaluate_polygonal_light_shading() does. */
typedef struct {
	uint32_t technique, maxp;
	const light_t* light;
	v3 position, corner_offset;
	float world_to_shading[4][3]; /* with the y-row mirrored where the shader does that (:444-449) */
	rw_urena_t urena;
	rw_sa_arvo_t sa_arvo;
	rw_sa_t sa;
	rw_bilinear_hart_t bilinear;
	rw_biquadratic_hart_t biquadratic;
	rw_psa_arvo_t psa_arvo;
	float ggx_density_factor; /* 1 / solid angle (or 1 / projected solid angle), :683-687 */
} rw_sampler_t;

/* Returns 0 if the shader returns vec3(0.0f) before sampling (polygon clipped away, empty projected solid angle) */
static int rw_sampler_prepare(rw_sampler_t* sp, uint32_t technique, uint32_t maxp, uint32_t maxl, const light_t* light, v3 position, const float world_to_shading[4][3]) {
	const v3 zero = mk3(0.0f, 0.0f, 0.0f);
	sp->technique = technique; sp->maxp = maxp; sp->light = light; sp->position = position; sp->ggx_density_factor = 0.0f;
	memcpy(sp->world_to_shading, world_to_shading, sizeof(sp->world_to_shading));
	if (technique == VKR_TECHNIQUE_BASELINE) /* :335 */
		sp->corner_offset = sub3(light->translation, position);
	else if (technique == VKR_TECHNIQUE_RECTANGLE_SOLID_ANGLE_URENA) { /* :357-359 */
		sp->urena = rw_prepare_urena(light->translation, light->scaling_x, light->scaling_y, light->rotation_cols, position);
		sp->ggx_density_factor = 1.0f / sp->urena.solid_angle;
	}
	else if (technique == VKR_TECHNIQUE_SOLID_ANGLE_ARVO) { /* :370-371 */
		rw_prepare_sa_arvo(&sp->sa_arvo, light->vertex_count, light->vertices_world_space, position, maxp);
		sp->ggx_density_factor = 1.0f / sp->sa_arvo.solid_angle;
	}
	else if (technique == VKR_TECHNIQUE_SOLID_ANGLE) { /* :382-383 */
		rw_prepare_sa(&sp->sa, light->vertex_count, light->vertices_world_space, position, maxp, 0);
		sp->ggx_density_factor = 1.0f / sp->sa.solid_angle;
	}
	else if (technique >= VKR_TECHNIQUE_CLIPPED_SOLID_ANGLE && technique <= VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO) { /* :392-405, 439-457 */
		if (technique == VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO) {
			float side = fmaf(light->plane[3], 1.0f, fmaf(light->plane[2], position.z, fmaf(light->plane[1], position.y, light->plane[0] * position.x)));
			for (int i = 0; i != 4; ++i) sp->world_to_shading[i][1] = (side < 0.0f) ? -sp->world_to_shading[i][1] : sp->world_to_shading[i][1];
		}
		v3 verts[PSA_MAXP];
		memset(verts, 0, sizeof(verts));
		for (uint32_t i = 0; i != maxl; ++i) verts[i] = mat43_mul_point(sp->world_to_shading, light->vertices_world_space[i]);
		uint32_t cvc = light->vertex_count;
		if (technique != VKR_TECHNIQUE_BILINEAR_COSINE_WARP_HART && technique != VKR_TECHNIQUE_BIQUADRATIC_COSINE_WARP_HART) {
			cvc = psa_clip_polygon(light->vertex_count, verts, maxp);
			if (cvc == 0) return 0;
		}
		if (technique == VKR_TECHNIQUE_CLIPPED_SOLID_ANGLE) {
			rw_prepare_sa(&sp->sa, cvc, verts, zero, maxp, 0);
			sp->ggx_density_factor = 1.0f / sp->sa.solid_angle;
		}
		else if (technique == VKR_TECHNIQUE_BILINEAR_COSINE_WARP_HART || technique == VKR_TECHNIQUE_BILINEAR_COSINE_WARP_CLIPPING_HART)
			rw_prepare_bilinear_hart(&sp->bilinear, cvc, verts, maxp, 0);
		else if (technique == VKR_TECHNIQUE_BIQUADRATIC_COSINE_WARP_HART || technique == VKR_TECHNIQUE_BIQUADRATIC_COSINE_WARP_CLIPPING_HART)
			rw_prepare_biquadratic_hart(&sp->biquadratic, cvc, verts, maxp, 0);
		else {
			rw_prepare_psa_arvo(&sp->psa_arvo, cvc, verts, maxp);
			if (sp->psa_arvo.projected_solid_angle <= 0.0f) return 0;
			sp->ggx_density_factor = 1.0f / sp->psa_arvo.projected_solid_angle;
		}
	}
	return 1;
}

/* One sample: world-space direction and its density with respect to the solid angle measure */
static v3 rw_sampler_sample(const rw_sampler_t* sp, v2 rnd, float* density) {
	const light_t* light = sp->light;
	v3 dir;
	switch (sp->technique) {
	case VKR_TECHNIQUE_BASELINE: /* :341-343 */
		*density = 1.0f;
		return normalize3(add3(add3(sp->corner_offset, scale3(light->rotation_cols[0], rnd.x)), scale3(light->rotation_cols[1], rnd.y)));
	case VKR_TECHNIQUE_AREA_TURK: { /* :348-351 */
		v3 light_sample = rw_sample_area_polygon_turk(light->vertex_count, light->vertices_world_space, light->fan_areas, rnd, sp->maxp);
		*density = rw_get_area_sample_density(&dir, light_sample, sp->position, mk3(light->plane[0], light->plane[1], light->plane[2]), light->area);
		return dir;
	}
	case VKR_TECHNIQUE_RECTANGLE_SOLID_ANGLE_URENA: /* :362-363 */
		*density = 1.0f / sp->urena.solid_angle;
		return rw_sample_urena(&sp->urena, rnd);
	case VKR_TECHNIQUE_SOLID_ANGLE_ARVO: /* :374-375 */
		*density = 1.0f / sp->sa_arvo.solid_angle;
		return rw_sample_sa_arvo(&sp->sa_arvo, rnd, sp->maxp);
	case VKR_TECHNIQUE_SOLID_ANGLE: /* :386-387 */
		*density = 1.0f / sp->sa.solid_angle;
		return rw_sample_sa(&sp->sa, rnd, sp->maxp);
	case VKR_TECHNIQUE_CLIPPED_SOLID_ANGLE: /* :412-414 */
		dir = rw_sample_sa(&sp->sa, rnd, sp->maxp);
		*density = 1.0f / sp->sa.solid_angle;
		return mat43_transpose_mul(sp->world_to_shading, dir);
	case VKR_TECHNIQUE_BILINEAR_COSINE_WARP_HART: case VKR_TECHNIQUE_BILINEAR_COSINE_WARP_CLIPPING_HART: /* :423-425 */
		dir = rw_sample_bilinear_hart(density, &sp->bilinear, rnd, sp->maxp);
		return mat43_transpose_mul(sp->world_to_shading, dir);
	case VKR_TECHNIQUE_BIQUADRATIC_COSINE_WARP_HART: case VKR_TECHNIQUE_BIQUADRATIC_COSINE_WARP_CLIPPING_HART: /* :433-435 */
		dir = rw_sample_biquadratic_hart(density, &sp->biquadratic, rnd, sp->maxp);
		return mat43_transpose_mul(sp->world_to_shading, dir);
	default: /* VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO, :476-478 */
		dir = rw_sample_psa_arvo(&sp->psa_arvo, rnd, 3, sp->maxp);
		*density = dir.z / sp->psa_arvo.projected_solid_angle;
		return mat43_transpose_mul(sp->world_to_shading, dir);
	}
}

/* shading_pass.frag.glsl:332-481 + 676-711 for the related-work techniques, strategies DIFFUSE_ONLY and DIFFUSE_GGX_MIS */
static v3 evaluate_polygonal_light_shading_related_work(const shading_data_t* sd, ltc_t ltc, const light_t* light, noise_accessor_t* accessor, const ctx_t* c, uint64_t* ray_count) {
	const vkr_oracle_config_t* cfg = c->cfg;
	rw_sampler_t sampler;
	if (!rw_sampler_prepare(&sampler, cfg->polygon_sampling_technique, c->maxp, cfg->max_light_vertex_count, light, sd->position, ltc.world_to_shading))
		return mk3(0.0f, 0.0f, 0.0f);
	if (cfg->error_display >= 1 && cfg->error_display <= 3 && cfg->polygon_sampling_technique == VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO) { /* :468-472 */
		v2 rnd = get_noise_2(accessor, c);
		v3 sampled_dir = rw_sample_psa_arvo(&sampler.psa_arvo, rnd, 3, c->maxp);
		v2 e = rw_psa_arvo_sampling_error(&sampler.psa_arvo, rnd, sampled_dir, c->maxp);
		return error_display_color(error_component(mk3(e.x, e.y, 0.0f), cfg->error_display), c);
	}
	v3 result = mk3(0.0f, 0.0f, 0.0f);
	for (uint32_t s = 0; s != cfg->sample_count; ++s) {
		float density;
		v3 dir = rw_sampler_sample(&sampler, get_noise_2(accessor, c), &density);
		result = add3(result, get_polygonal_light_mis_estimate(dir, density, sd, light, c, ray_count));
	}
	if (cfg->sampling_strategies == VKR_STRATEGY_DIFFUSE_GGX_MIS) {
		memcpy(ltc.world_to_shading, sampler.world_to_shading, sizeof(ltc.world_to_shading));
		result = ggx_mis_samples(result, sampler.ggx_density_factor, 0, sd, &ltc, light, accessor, c, ray_count);
	}
	return scale3(result, 1.0f / (float) cfg->sample_count);
}

/* shading_pass.frag.glsl:329-711, PSA branches (:441-504 and :506-673, :676-709) */
static v3 evaluate_polygonal_light_shading(const shading_data_t* sd, ltc_t ltc, const light_t* light, noise_accessor_t* accessor, const ctx_t* c, uint64_t* ray_count) {
	const vkr_oracle_config_t* cfg = c->cfg;
	if (cfg->polygon_sampling_technique != VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE)
		return evaluate_polygonal_light_shading_related_work(sd, ltc, light, accessor, c, ray_count);
	const int biased = cfg->biased_sampling;
	const uint32_t S = cfg->sample_count;
	const uint32_t maxp = c->maxp;
	const uint32_t maxl = cfg->max_light_vertex_count;
	v3 result = mk3(0.0f, 0.0f, 0.0f);
	float side = fmaf(light->plane[3], 1.0f, fmaf(light->plane[2], sd->position.z, fmaf(light->plane[1], sd->position.y, light->plane[0] * sd->position.x)));
	for (int i = 0; i != 4; ++i) {
		ltc.world_to_shading[i][1] = (side < 0.0f) ? -ltc.world_to_shading[i][1] : ltc.world_to_shading[i][1];
		ltc.world_to_cosine[i][1] = (side < 0.0f) ? -ltc.world_to_cosine[i][1] : ltc.world_to_cosine[i][1];
	}
	psa_polygon_t polygon_diffuse, polygon_specular;
	memset(&polygon_specular, 0, sizeof(polygon_specular));
	const int strat = cfg->sampling_strategies;
	if (strat == VKR_STRATEGY_DIFFUSE_ONLY || strat == VKR_STRATEGY_DIFFUSE_GGX_MIS) {
		v3 verts[PSA_MAXP];
		memset(verts, 0, sizeof(verts));
		for (uint32_t i = 0; i != maxl; ++i) verts[i] = mat43_mul_point(ltc.world_to_shading, light->vertices_world_space[i]);
		uint32_t cvc = psa_clip_polygon(light->vertex_count, verts, maxp);
		if (cvc == 0) return mk3(0.0f, 0.0f, 0.0f);
		psa_prepare(&polygon_diffuse, cvc, verts, maxp, biased);
		if (polygon_diffuse.projected_solid_angle <= 0.0f) return mk3(0.0f, 0.0f, 0.0f);
		if (cfg->error_display >= 1 && cfg->error_display <= 3) { /* ERROR_DISPLAY_DIFFUSE, :489-493 */
			v2 rnd = get_noise_2(accessor, c);
			v3 sampled_dir = psa_sample(&polygon_diffuse, rnd, maxp, biased);
			return error_display_color(error_component(psa_sampling_error(&polygon_diffuse, rnd, sampled_dir, maxp, biased), cfg->error_display), c);
		}
		for (uint32_t s = 0; s != S; ++s) {
			v3 dir = psa_sample(&polygon_diffuse, get_noise_2(accessor, c), maxp, biased);
			float density = dir.z / polygon_diffuse.projected_solid_angle;
			v3 w = mat43_transpose_mul(ltc.world_to_shading, dir);
			result = add3(result, get_polygonal_light_mis_estimate(w, density, sd, light, c, ray_count));
		}
	}
	else {
		for (uint32_t i = 0; i != 2; ++i) {
			float (*w2l)[3] = (i == 0) ? ltc.world_to_shading : ltc.world_to_cosine;
			if (i > 0) polygon_diffuse = polygon_specular;
			v3 verts[PSA_MAXP];
			memset(verts, 0, sizeof(verts));
			for (uint32_t j = 0; j != maxl; ++j) verts[j] = mat43_mul_point(w2l, light->vertices_world_space[j]);
			uint32_t cvc = psa_clip_polygon(light->vertex_count, verts, maxp);
			if (cvc == 0 && i == 0) return mk3(0.0f, 0.0f, 0.0f);
			else if (cvc == 0) { polygon_specular.projected_solid_angle = 0.0f; break; }
			psa_prepare(&polygon_specular, cvc, verts, maxp, biased);
		}
		if (polygon_diffuse.projected_solid_angle == 0.0f) return mk3(0.0f, 0.0f, 0.0f);
		float specular_albedo = ltc.albedo;
		float specular_weight = specular_albedo * polygon_specular.projected_solid_angle;
		if (cfg->error_display >= 1 && cfg->error_display <= 3) { /* ERROR_DISPLAY_DIFFUSE, :549-553 */
			v2 rnd = get_noise_2(accessor, c);
			v3 sampled_dir = psa_sample(&polygon_diffuse, rnd, maxp, biased);
			return error_display_color(error_component(psa_sampling_error(&polygon_diffuse, rnd, sampled_dir, maxp, biased), cfg->error_display), c);
		}
		if (cfg->error_display >= 4) { /* ERROR_DISPLAY_SPECULAR, :555-563 */
			if (!(polygon_specular.projected_solid_angle > 0.0f)) return mk3(0.0f, 0.0f, 0.0f);
			v2 rnd = get_noise_2(accessor, c);
			v3 sampled_dir = psa_sample(&polygon_specular, rnd, maxp, biased);
			return error_display_color(error_component(psa_sampling_error(&polygon_specular, rnd, sampled_dir, maxp, biased), cfg->error_display), c);
		}
		if (strat == VKR_STRATEGY_DIFFUSE_SPECULAR_SEPARATELY) {
			for (uint32_t s = 0; s != S; ++s) {
				v3 diffuse_dir = psa_sample(&polygon_diffuse, get_noise_2(accessor, c), maxp, biased);
				diffuse_dir = mat43_transpose_mul(ltc.world_to_shading, diffuse_dir);
				v3 rtb = radiance_visibility_brdf_product(NULL, NULL, diffuse_dir, sd, light, 1, 0, c, ray_count);
				result = add3(result, scale3(rtb, polygon_diffuse.projected_solid_angle));
				if (polygon_specular.projected_solid_angle > 0.0f) {
					v3 dir_cosine_space = psa_sample(&polygon_specular, get_noise_2(accessor, c), maxp, biased);
					v3 dir_shading_space = normalize3(mat3_mul(ltc.cosine_to_shading, dir_cosine_space));
					float ltc_density = evaluate_ltc_density(&ltc, dir_shading_space, 1.0f);
					v3 rtb2 = radiance_visibility_brdf_product(NULL, NULL, mat43_transpose_mul(ltc.world_to_shading, dir_shading_space), sd, light, 0, 1, c, ray_count);
					if (!(dir_shading_space.z <= 0.0f || dir_cosine_space.z <= 0.0f)) {
						/* radiance_times_brdf * z * psa / density, left to right */
						result.x += rtb2.x * dir_shading_space.z * polygon_specular.projected_solid_angle / ltc_density;
						result.y += rtb2.y * dir_shading_space.z * polygon_specular.projected_solid_angle / ltc_density;
						result.z += rtb2.z * dir_shading_space.z * polygon_specular.projected_solid_angle / ltc_density;
					}
				}
			}
		}
		else if (strat == VKR_STRATEGY_DIFFUSE_SPECULAR_MIS) {
			v3 diffuse_albedo = mk3(vkr_max(sd->diffuse_albedo.x, 0.01f), vkr_max(sd->diffuse_albedo.y, 0.01f), vkr_max(sd->diffuse_albedo.z, 0.01f));
			v3 diffuse_weight = scale3(diffuse_albedo, polygon_diffuse.projected_solid_angle);
			uint32_t technique_count = (polygon_specular.projected_solid_angle > 0.0f) ? 2 : 1;
			float rcp_diffuse_psa = 1.0f / polygon_diffuse.projected_solid_angle;
			float rcp_specular_psa = 1.0f / polygon_specular.projected_solid_angle;
			v3 specular_weight_rgb = mk3(specular_weight, specular_weight, specular_weight);
			if (cfg->mis_heuristic == VKR_MIS_OPTIMAL) {
				v3 radiance_over_pi = scale3(light->surface_radiance, VKR_INV_PI);
				diffuse_weight = mk3(diffuse_weight.x * radiance_over_pi.x, diffuse_weight.y * radiance_over_pi.y, diffuse_weight.z * radiance_over_pi.z);
				specular_weight_rgb = mk3(specular_weight_rgb.x * radiance_over_pi.x, specular_weight_rgb.y * radiance_over_pi.y, specular_weight_rgb.z * radiance_over_pi.z);
			}
			for (uint32_t s = 0; s != S; ++s) {
				v3 dir_diffuse = psa_sample(&polygon_diffuse, get_noise_2(accessor, c), maxp, biased);
				v3 dir_specular = mk3(0.0f, 0.0f, 0.0f);
				if (polygon_specular.projected_solid_angle > 0.0f) {
					dir_specular = psa_sample(&polygon_specular, get_noise_2(accessor, c), maxp, biased);
					dir_specular = normalize3(mat3_mul(ltc.cosine_to_shading, dir_specular));
				}
				for (uint32_t j = 0; j != technique_count; ++j) {
					v3 dir = (j == 0) ? dir_diffuse : dir_specular;
					if (dir.z <= 0.0f) continue;
					float diffuse_density = dir.z * rcp_diffuse_psa;
					float specular_density = evaluate_ltc_density(&ltc, dir, rcp_specular_psa);
					int visibility;
					v3 rtb = radiance_visibility_brdf_product(NULL, &visibility, mat43_transpose_mul(ltc.world_to_shading, dir), sd, light, 1, 1, c, ray_count);
					v3 integrand = scale3(rtb, dir.z);
					if (j == 0 && polygon_specular.projected_solid_angle <= 0.0f) {
						if (visibility) result = add3(result, scale3(integrand, 1.0f / diffuse_density));
					}
					else if (j == 0)
						result = add3(result, get_mis_estimate(integrand, diffuse_weight, diffuse_density, specular_weight_rgb, specular_density, c->mis_visibility_estimate, cfg->mis_heuristic));
					else
						result = add3(result, get_mis_estimate(integrand, specular_weight_rgb, specular_density, diffuse_weight, diffuse_density, c->mis_visibility_estimate, cfg->mis_heuristic));
				}
			}
		}
		else if (strat == VKR_STRATEGY_DIFFUSE_SPECULAR_RANDOM) {
			float diffuse_albedo = vkr_max(dot3(sd->diffuse_albedo, mk3(0.21263901f, 0.71516868f, 0.07219232f)), 0.01f);
			float diffuse_weight = diffuse_albedo * polygon_diffuse.projected_solid_angle;
			float diffuse_ratio = diffuse_weight / (diffuse_weight + specular_weight);
			for (uint32_t s = 0; s != S; ++s) {
				v2 rnd = get_noise_2(accessor, c);
				int specular_selected = rnd.x >= diffuse_ratio;
				float random_number_offset = specular_selected ? 1.0f : 0.0f;
				rnd.x = (rnd.x - random_number_offset) / (diffuse_ratio - random_number_offset);
				const psa_polygon_t* selected = specular_selected ? &polygon_specular : &polygon_diffuse;
				v3 dir = psa_sample(selected, rnd, maxp, biased);
				if (specular_selected) dir = normalize3(mat3_mul(ltc.cosine_to_shading, dir));
				float lambert = dir.z;
				float diffuse_density = lambert * diffuse_albedo;
				float specular_density = evaluate_ltc_density(&ltc, dir, specular_albedo);
				float density = (diffuse_density + specular_density) / (diffuse_weight + specular_weight);
				v3 rtb = radiance_visibility_brdf_product(&lambert, NULL, mat43_transpose_mul(ltc.world_to_shading, dir), sd, light, 1, 1, c, ray_count);
				if (!(dir.z <= 0.0f)) {
					result.x += rtb.x * dir.z / density;
					result.y += rtb.y * dir.z / density;
					result.z += rtb.z * dir.z / density;
				}
			}
		}
	}
	if (strat == VKR_STRATEGY_DIFFUSE_GGX_MIS) {
		v3 outgoing_ss = mat43_mul_dir(ltc.world_to_shading, sd->outgoing);
		outgoing_ss.y = 0.0f;
		float density_factor = 1.0f / polygon_diffuse.projected_solid_angle;
		for (uint32_t s = 0; s != S; ++s) {
			float ggx_density;
			v3 dir_ss = sample_ggx_reflected_direction(&ggx_density, outgoing_ss, sd->roughness, get_noise_2(accessor, c));
			v3 dir_ws = mat43_transpose_mul(ltc.world_to_shading, dir_ss);
			if (dir_ss.z > 0.0f && polygonal_light_ray_intersection(light, maxl, sd->position, dir_ws, 0.0f)) {
				float lambert;
				v3 rtb = radiance_visibility_brdf_product(&lambert, NULL, dir_ws, sd, light, 1, 1, c, ray_count);
				float polygon_density = lambert * density_factor;
				float w = get_mis_weight_over_density(ggx_density, polygon_density, cfg->mis_heuristic);
				result.x += rtb.x * lambert * w; result.y += rtb.y * lambert * w; result.z += rtb.z * lambert * w;
			}
		}
	}
	return scale3(result, 1.0f / (float) S);
}

static void parse_light(light_t* L, const uint8_t* p, uint32_t V) { /* polygonal_light_utility.glsl:26-83, V = MAX_POLYGONAL_LIGHT_VERTEX_COUNT */
	L->surface_radiance = mk3(rdf(p, L_SURFACE_RADIANCE), rdf(p, L_SURFACE_RADIANCE + 4), rdf(p, L_SURFACE_RADIANCE + 8));
	for (int i = 0; i != 4; ++i) L->plane[i] = rdf(p, L_PLANE + 4 * i);
	L->vertex_count = rdu(p, L_VERTEX_COUNT);
	L->texturing_technique = rdu(p, L_TEXTURING); L->texture_index = rdu(p, L_TEXTURING + 4);
	L->inv_scaling_x = rdf(p, 44); L->inv_scaling_y = rdf(p, 60);
	const uint8_t* vw = p + L_FIXED_SIZE + 16 * (size_t) V;
	for (uint32_t i = 0; i != V; ++i) L->vertices_world_space[i] = mk3(rdf(vw, 16 * i), rdf(vw, 16 * i + 4), rdf(vw, 16 * i + 8));
	L->translation = mk3(rdf(p, L_TRANSLATION), rdf(p, L_TRANSLATION + 4), rdf(p, L_TRANSLATION + 8));
	L->scaling_x = rdf(p, L_SCALING_X); L->scaling_y = rdf(p, L_SCALING_Y); L->area = rdf(p, L_AREA);
	for (int col = 0; col != 3; ++col) L->rotation_cols[col] = mk3(rdf(p, L_ROTATION + 4 * col), rdf(p, L_ROTATION + 16 + 4 * col), rdf(p, L_ROTATION + 32 + 4 * col));
	const uint8_t* fa = vw + 16 * (size_t) V;
	for (uint32_t i = 0; i + 2 < V; ++i) L->fan_areas[i] = mk2(rdf(fa, 16 * i), rdf(fa, 16 * i + 4));
}

static uint32_t technique_maxp(uint32_t technique, uint32_t V) { /* get_max_polygon_vertex_count, main.c:194-216: the techniques that clip may gain one vertex */
	switch (technique) {
	case VKR_TECHNIQUE_CLIPPED_SOLID_ANGLE: case VKR_TECHNIQUE_BILINEAR_COSINE_WARP_CLIPPING_HART: case VKR_TECHNIQUE_BIQUADRATIC_COSINE_WARP_CLIPPING_HART:
	case VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO: case VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE: return V + 1;
	default: return V;
	}
}

static void parse_context(ctx_t* c, const vkr_oracle_config_t* cfg, const uint8_t* constants) {
	c->cfg = cfg; c->constants = constants;
	for (int i = 0; i != 3; ++i) for (int j = 0; j != 4; ++j) c->pixel_to_ray[i][j] = rdf(constants, OFF_PIXEL_TO_RAY + 4 * (4 * i + j));
	c->camera = mk3(rdf(constants, OFF_CAMERA), rdf(constants, OFF_CAMERA + 4), rdf(constants, OFF_CAMERA + 8));
	c->mis_visibility_estimate = rdf(constants, OFF_MIS_VIS);
	c->exposure = rdf(constants, OFF_EXPOSURE);
	c->error_factor = rdf(constants, OFF_ERROR_FACTOR);
	c->roughness_factor = rdf(constants, OFF_ROUGHNESS_FACTOR);
	c->noise_res_mask[0] = rdu(constants, OFF_NOISE_RES_MASK); c->noise_res_mask[1] = rdu(constants, OFF_NOISE_RES_MASK + 4);
	c->noise_layer_mask = rdu(constants, OFF_NOISE_LAYER_MASK);
	for (int i = 0; i != 4; ++i) c->noise_random[i] = rdu(constants, OFF_NOISE_RANDOM + 4 * i);
	for (int i = 0; i != 6; ++i) c->ltc_c[i] = rdf(constants, OFF_LTC + 4 * i);
	uint32_t V = cfg->max_light_vertex_count;
	size_t stride = vkr_oracle_light_stride(V);
	c->lights = (light_t*) calloc(cfg->light_count ? cfg->light_count : 1, sizeof(light_t));
	for (uint32_t l = 0; l != cfg->light_count; ++l) {
		const uint8_t* p = constants + CONSTANTS_FIXED_SIZE + stride * l;
		parse_light(&c->lights[l], p, V);
	}
	c->maxp = technique_maxp(cfg->polygon_sampling_technique, V);
}

/* Wall-clock seconds of the pixel loop of the last vkr_oracle_shade call (BVH build excluded), for the CPU baseline */
static double g_last_shade_seconds = 0.0;
/* vkr_texture_bilinear_repeat_clamp for n (u, v) pairs on a one-level RGBA32F texture (tests) */
void vkr_oracle_light_texture_batch(uint32_t width, uint32_t height, const float* texels, uint32_t n, const float* uv, float* out_rgba) {
	vkr_texture_view_t t; t.width = width; t.height = height; t.mip_count = 1; t.texels = texels;
	for (uint32_t i = 0; i != n; ++i) vkr_texture_bilinear_repeat_clamp(out_rgba + 4 * i, &t, uv[2 * i], uv[2 * i + 1]);
}

double vkr_oracle_last_shade_seconds(void) { return g_last_shade_seconds; }

int vkr_oracle_shade(const vkr_oracle_config_t* cfg, const void* constants, const float* gbuffer,
	const uint16_t* noise, uint32_t noise_w, uint32_t noise_h, uint32_t noise_layers,
	const uint16_t* ltc0, const uint16_t* ltc1, uint32_t ltc_res, uint32_t ltc_layers,
	const float* tris, uint32_t tri_count, float* out_rgba, uint64_t* out_ray_count)
{
	return vkr_oracle_shade_with_light_textures(cfg, constants, gbuffer, noise, noise_w, noise_h, noise_layers, ltc0, ltc1, ltc_res, ltc_layers, tris, tri_count, 0, NULL, NULL, NULL, out_rgba, out_ray_count);
}

/* shading_pass.frag.glsl:824-866 with the G-buffer standing in for get_shading_data(). Light textures (g_light_textures, the indices are in the
   light blocks): RGBA32F, light_texture_dims = {width, height, mip count} per texture, offsets in floats into light_texture_data; level 0 is used */
int vkr_oracle_shade_with_light_textures(const vkr_oracle_config_t* cfg, const void* constants, const float* gbuffer,
	const uint16_t* noise, uint32_t noise_w, uint32_t noise_h, uint32_t noise_layers,
	const uint16_t* ltc0, const uint16_t* ltc1, uint32_t ltc_res, uint32_t ltc_layers,
	const float* tris, uint32_t tri_count,
	uint32_t light_texture_count, const uint32_t* light_texture_dims, const uint64_t* light_texture_offsets, const float* light_texture_data,
	float* out_rgba, uint64_t* out_ray_count)
{
	(void) noise_layers;
	if (cfg->max_light_vertex_count < 3 || cfg->max_light_vertex_count > 7) return 1;
	if (cfg->polygon_sampling_technique > VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE) return 1;
	if (cfg->polygon_sampling_technique != VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE) {
		/* legality rules of the reference's user interface (user_interface.cpp:90-180) */
		const uint32_t t = cfg->polygon_sampling_technique;
		const int ggx_ok = t == VKR_TECHNIQUE_RECTANGLE_SOLID_ANGLE_URENA || t == VKR_TECHNIQUE_SOLID_ANGLE_ARVO || t == VKR_TECHNIQUE_SOLID_ANGLE || t == VKR_TECHNIQUE_CLIPPED_SOLID_ANGLE || t == VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO;
		if (!(cfg->sampling_strategies == VKR_STRATEGY_DIFFUSE_ONLY || (cfg->sampling_strategies == VKR_STRATEGY_DIFFUSE_GGX_MIS && ggx_ok)) || cfg->biased_sampling) {
			printf("oracle: sampling technique %u does not support sampling strategy %u.\n", t, cfg->sampling_strategies);
			return 1;
		}
	}
	if (cfg->error_display > 6) return 1;
	if (cfg->error_display) {
		/* the shader looks at ERROR_DISPLAY_* in the projected solid angle branches only (:468, 489, 549, 555); Arvo's error has two components */
		const uint32_t t = cfg->polygon_sampling_technique;
		const int combined = cfg->sampling_strategies >= VKR_STRATEGY_DIFFUSE_SPECULAR_SEPARATELY;
		if ((t != VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE && t != VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO) || (cfg->error_display >= 4 && !combined)
			|| (t == VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO && cfg->error_display == 3)) {
			printf("oracle: error display %u is not available with technique %u and strategy %u.\n", cfg->error_display, t, cfg->sampling_strategies);
			return 1;
		}
	}
	ctx_t c; memset(&c, 0, sizeof(c));
	parse_context(&c, cfg, (const uint8_t*) constants);
	for (uint32_t l = 0; l != cfg->light_count; ++l)
		if (c.lights[l].texturing_technique > 3 || (c.lights[l].texturing_technique != 0 && c.lights[l].texture_index >= light_texture_count)) {
			printf("oracle: light %u is textured (technique %u) with texture %u of %u.\n", l, c.lights[l].texturing_technique, c.lights[l].texture_index, light_texture_count); free(c.lights); return 1;
		}
	vkr_texture_view_t* light_views = (vkr_texture_view_t*) calloc(light_texture_count ? light_texture_count : 1, sizeof(vkr_texture_view_t));
	for (uint32_t i = 0; i != light_texture_count; ++i) {
		light_views[i].width = light_texture_dims[3 * i]; light_views[i].height = light_texture_dims[3 * i + 1]; light_views[i].mip_count = light_texture_dims[3 * i + 2];
		light_views[i].texels = light_texture_data + light_texture_offsets[i];
	}
	c.light_textures = light_views; c.light_texture_count = light_texture_count;
	c.noise = noise; c.noise_w = noise_w; c.noise_h = noise_h;
	c.ltc0 = ltc0; c.ltc1 = ltc1; c.ltc_res = ltc_res; c.ltc_layers = ltc_layers;
	obvh_t bvh; memset(&bvh, 0, sizeof(bvh));
	if (cfg->trace_shadow_rays) obvh_build(&bvh, tris, tri_count);
	c.bvh = &bvh;
	const uint32_t W = cfg->width, H = cfg->height;
	const uint32_t y0 = cfg->row_begin, y1 = cfg->row_end ? cfg->row_end : H;
	const size_t plane = (size_t) W * H * 4;
	uint64_t total_rays = 0;
#ifdef _OPENMP
	const double shade_begin = omp_get_wtime();
#endif
	/* work items are 64-pixel pieces of rows: a banded sample of a frame has fewer rows than a big host has threads */
	const uint32_t pieces = (W + 63) / 64;
	#pragma omp parallel for collapse(2) schedule(dynamic, 1) reduction(+:total_rays)
	for (uint32_t y = y0; y < y1; ++y) for (uint32_t piece = 0; piece < pieces; ++piece) {
		if (cfg->band_stride && (y - y0) % cfg->band_stride >= cfg->band_height) continue;
		const uint32_t x_end = (piece * 64 + 64 < W) ? piece * 64 + 64 : W;
		for (uint32_t x = piece * 64; x != x_end; ++x) {
			size_t pi = ((size_t) y * W + x) * 4;
			uint64_t rays = 0;
			v3 final_color = mk3(0.0f, 0.0f, 0.0f);
			int valid = gbuffer[plane + pi + 3] != 0.0f;
			v3 view_dir = mk3(
				fmaf(c.pixel_to_ray[0][2], 1.0f, fmaf(c.pixel_to_ray[0][1], (float) y, c.pixel_to_ray[0][0] * (float) x)),
				fmaf(c.pixel_to_ray[1][2], 1.0f, fmaf(c.pixel_to_ray[1][1], (float) y, c.pixel_to_ray[1][0] * (float) x)),
				fmaf(c.pixel_to_ray[2][2], 1.0f, fmaf(c.pixel_to_ray[2][1], (float) y, c.pixel_to_ray[2][0] * (float) x)));
			shading_data_t sd; memset(&sd, 0, sizeof(sd));
			v3 ray_end; float ray_end_w;
			if (!valid) { ray_end = view_dir; ray_end_w = 0.0f; }
			else {
				sd.position = mk3(gbuffer[pi], gbuffer[pi + 1], gbuffer[pi + 2]);
				sd.roughness = gbuffer[pi + 3];
				sd.normal = mk3(gbuffer[plane + pi], gbuffer[plane + pi + 1], gbuffer[plane + pi + 2]);
				sd.diffuse_albedo = mk3(gbuffer[2 * plane + pi], gbuffer[2 * plane + pi + 1], gbuffer[2 * plane + pi + 2]);
				sd.fresnel_0 = mk3(gbuffer[3 * plane + pi], gbuffer[3 * plane + pi + 1], gbuffer[3 * plane + pi + 2]);
				sd.outgoing = normalize3(sub3(c.camera, sd.position));
				sd.lambert_outgoing = dot3(sd.normal, sd.outgoing);
				ray_end = sd.position; ray_end_w = 1.0f;
			}
			if (cfg->show_polygonal_lights) {
				const v3 view_dir_normalized = normalize3(view_dir); /* :844 */
				for (uint32_t l = 0; l != cfg->light_count; ++l)
					if (polygonal_light_ray_intersection(&c.lights[l], cfg->max_light_vertex_count, c.camera, ray_end, ray_end_w))
						final_color = add3(final_color, get_polygon_radiance(view_dir_normalized, c.camera, &c.lights[l], &c));
			}
			if (valid) {
				float fresnel_luminance = dot3(sd.fresnel_0, mk3(0.2126f, 0.7152f, 0.0722f));
				ltc_t ltc;
				get_ltc_coefficients(&ltc, &c, fresnel_luminance, sd.roughness, sd.position, sd.normal, sd.outgoing);
				noise_accessor_t acc; memset(&acc, 0, sizeof(acc));
				acc.pixel[0] = x; acc.pixel[1] = y;
				for (uint32_t l = 0; l != cfg->light_count; ++l)
					final_color = add3(final_color, evaluate_polygonal_light_shading(&sd, ltc, &c.lights[l], &acc, &c, &rays));
			}
			if (isnan(final_color.x) || isnan(final_color.y) || isnan(final_color.z) || isinf(final_color.x) || isinf(final_color.y) || isinf(final_color.z))
				final_color = mk3(1.0f / c.exposure, 0.0f / c.exposure, 0.8f / c.exposure);
			v3 out_color = mk3(final_color.x * c.exposure, final_color.y * c.exposure, final_color.z * c.exposure);
			/* HDR screenshots: two LDR frames hold the low / high bytes of the half-precision colour (shading_pass.frag.glsl:871-887) */
			const uint32_t frame_bits = rdu(constants, OFF_FRAME_BITS);
			if (frame_bits > 0) {
				const uint32_t mask = (frame_bits == 1) ? 0xFFu : 0xFF00u, shift = (frame_bits == 1) ? 0u : 8u;
				const uint32_t h0 = vkr_pack_half_2x16(out_color.x, out_color.y), h1 = vkr_pack_half_2x16(out_color.z, 1.0f);
				out_color = mk3((float) ((h0 & mask) >> shift) * (1.0f / 255.0f), (float) ((((h0 & 0xFFFF0000u) >> 16) & mask) >> shift) * (1.0f / 255.0f), (float) ((h1 & mask) >> shift) * (1.0f / 255.0f));
				if (!cfg->output_srgb) out_color = mk3(vkr_srgb_to_linear(out_color.x), vkr_srgb_to_linear(out_color.y), vkr_srgb_to_linear(out_color.z));
			}
			else if (cfg->output_srgb) /* :888-892 */
				out_color = mk3(vkr_linear_to_srgb(out_color.x), vkr_linear_to_srgb(out_color.y), vkr_linear_to_srgb(out_color.z));
			out_rgba[pi + 0] = out_color.x;
			out_rgba[pi + 1] = out_color.y;
			out_rgba[pi + 2] = out_color.z;
			out_rgba[pi + 3] = 1.0f;
			total_rays += rays;
		}
	}
#ifdef _OPENMP
	g_last_shade_seconds = omp_get_wtime() - shade_begin;
#endif
	if (out_ray_count) *out_ray_count = total_rays;
	obvh_destroy(&bvh);
	free(c.lights); free(light_views);
	return 0;
}

/* ---- G-buffer producer: visibility stand-in + get_shading_data (shading_pass.frag.glsl:721-822) */

/* mesh_quantization.glsl:38-45 */
static v3 decode_position_64_bit(uint32_t q0, uint32_t q1, const float* factor, const float* summand) {
	float px = (float) (q0 & 0x1FFFFF);
	float py = (float) (((q0 & 0xFFE00000u) >> 21) | ((q1 & 0x3FF) << 11));
	float pz = (float) ((q1 & 0x7FFFFC00u) >> 10);
	return mk3(fmaf(px, factor[0], summand[0]), fmaf(py, factor[1], summand[1]), fmaf(pz, factor[2], summand[2]));
}
/* mesh_quantization.glsl:19-33; input = UNORM16 pair already divided by 65535 */
static v3 decode_normal_32_bit(float ox, float oy) {
	const float factor = 2.0f * (65534.0f / 65535.0f);
	const float summand = -(32768.0f / 65535.0f) * factor;
	ox = fmaf(ox, factor, summand); oy = fmaf(oy, factor, summand);
	v3 normal = mk3(ox, oy, 1.0f - fabsf(ox) - fabsf(oy));
	float sx = (ox >= 0.0f) ? 1.0f : -1.0f, sy = (oy >= 0.0f) ? 1.0f : -1.0f;
	if (normal.z < 0.0f) {
		float nx = (1.0f - fabsf(normal.y)) * sx;
		float ny = (1.0f - fabsf(normal.x)) * sy;
		normal.x = nx; normal.y = ny;
	}
	return normalize3(normal);
}

/* scene.c:175-187: the float soup handed to the acceleration-structure build (a*b+c, not fused) */
void vkr_oracle_dequantize_for_bvh(const uint32_t* quantized_positions, uint64_t vertex_count, const float* factor, const float* summand, float* out_vertices) {
	for (uint64_t i = 0; i != vertex_count; ++i) {
		uint32_t q0 = quantized_positions[2 * i], q1 = quantized_positions[2 * i + 1];
		float p[3] = { (float) (q0 & 0x1FFFFF), (float) (((q0 & 0xFFE00000u) >> 21) | ((q1 & 0x3FF) << 11)), (float) ((q1 & 0x7FFFFC00u) >> 10) };
		for (int j = 0; j != 3; ++j) out_vertices[3 * i + j] = p[j] * factor[j] + summand[j];
	}
}

/* Visibility-buffer stand-in: closest hit of the primary ray through each pixel centre against
   the shader-decoded (fma) vertices. The reference rasterises (visibility_pass.vert.glsl:27-33);
   this only has to give the SAME triangle to oracle and product. */
int vkr_oracle_visibility(uint32_t width, uint32_t height, const void* constants_v, const uint32_t* quantized_positions, uint64_t tri_count, uint32_t* out_visibility) {
	const uint8_t* constants = (const uint8_t*) constants_v;
	float factor[3], summand[3];
	for (int i = 0; i != 3; ++i) { factor[i] = rdf(constants, OFF_DEQUANT_FACTOR + 4 * i); summand[i] = rdf(constants, OFF_DEQUANT_SUMMAND + 4 * i); }
	float* tris = (float*) malloc(sizeof(float) * 9 * (size_t) tri_count);
	for (uint64_t i = 0; i != tri_count * 3; ++i) {
		v3 p = decode_position_64_bit(quantized_positions[2 * i], quantized_positions[2 * i + 1], factor, summand);
		tris[3 * i] = p.x; tris[3 * i + 1] = p.y; tris[3 * i + 2] = p.z;
	}
	obvh_t bvh; obvh_build(&bvh, tris, (uint32_t) tri_count);
	float p2r[3][4];
	for (int i = 0; i != 3; ++i) for (int j = 0; j != 4; ++j) p2r[i][j] = rdf(constants, OFF_PIXEL_TO_RAY + 4 * (4 * i + j));
	v3 cam = mk3(rdf(constants, OFF_CAMERA), rdf(constants, OFF_CAMERA + 4), rdf(constants, OFF_CAMERA + 8));
	#pragma omp parallel for schedule(dynamic, 4)
	for (uint32_t y = 0; y < height; ++y)
		for (uint32_t x = 0; x != width; ++x) {
			v3 d = mk3(
				fmaf(p2r[0][2], 1.0f, fmaf(p2r[0][1], (float) y, p2r[0][0] * (float) x)),
				fmaf(p2r[1][2], 1.0f, fmaf(p2r[1][1], (float) y, p2r[1][0] * (float) x)),
				fmaf(p2r[2][2], 1.0f, fmaf(p2r[2][1], (float) y, p2r[2][0] * (float) x)));
			float t;
			int32_t hit = obvh_closest(&bvh, cam, d, 0.0f, INFINITY, &t);
			out_visibility[(size_t) y * width + x] = (hit < 0) ? 0xFFFFFFFFu : (uint32_t) hit;
		}
	obvh_destroy(&bvh); free(tris);
	return 0;
}

/* get_shading_data (shading_pass.frag.glsl:721-822) with constant per-material textures:
   material_params = 8 floats per material {base.rgb, specular.g (linear roughness), specular.b (metalicity), normal.x, normal.y, pad}. */
/* textures: NULL (constant materials from material_params) or 3 views per material {base colour, specular, normal}, texture_filter.h */
static int gbuffer_impl(uint32_t width, uint32_t height, const void* constants_v, const uint32_t* visibility,
	const uint32_t* quantized_positions, const uint16_t* normals_and_tex_coords, const uint8_t* material_indices,
	const float* material_params, const vkr_texture_view_t* textures, float* out_gbuffer)
{
	const uint8_t* constants = (const uint8_t*) constants_v;
	float factor[3], summand[3];
	for (int i = 0; i != 3; ++i) { factor[i] = rdf(constants, OFF_DEQUANT_FACTOR + 4 * i); summand[i] = rdf(constants, OFF_DEQUANT_SUMMAND + 4 * i); }
	float p2r[3][4];
	for (int i = 0; i != 3; ++i) for (int j = 0; j != 4; ++j) p2r[i][j] = rdf(constants, OFF_PIXEL_TO_RAY + 4 * (4 * i + j));
	v3 cam = mk3(rdf(constants, OFF_CAMERA), rdf(constants, OFF_CAMERA + 4), rdf(constants, OFF_CAMERA + 8));
	float roughness_factor = rdf(constants, OFF_ROUGHNESS_FACTOR);
	const size_t plane = (size_t) width * height * 4;
	#pragma omp parallel for schedule(dynamic, 4)
	for (uint32_t y = 0; y < height; ++y)
		for (uint32_t x = 0; x != width; ++x) {
			size_t pi = ((size_t) y * width + x) * 4;
			uint32_t prim = visibility[(size_t) y * width + x];
			for (int k = 0; k != 4; ++k) { out_gbuffer[pi + k] = 0.0f; out_gbuffer[plane + pi + k] = 0.0f; out_gbuffer[2 * plane + pi + k] = 0.0f; out_gbuffer[3 * plane + pi + k] = 0.0f; }
			if (prim == 0xFFFFFFFFu) continue;
			v3 ray_direction = mk3(
				fmaf(p2r[0][2], 1.0f, fmaf(p2r[0][1], (float) y, p2r[0][0] * (float) x)),
				fmaf(p2r[1][2], 1.0f, fmaf(p2r[1][1], (float) y, p2r[1][0] * (float) x)),
				fmaf(p2r[2][2], 1.0f, fmaf(p2r[2][1], (float) y, p2r[2][0] * (float) x)));
			v3 positions[3], normals[3]; v2 tex_coords[3];
			for (int i = 0; i != 3; ++i) {
				size_t vi = (size_t) prim * 3 + i;
				positions[i] = decode_position_64_bit(quantized_positions[2 * vi], quantized_positions[2 * vi + 1], factor, summand);
				const uint16_t* nt = normals_and_tex_coords + 4 * vi;
				normals[i] = decode_normal_32_bit((float) nt[0] / 65535.0f, (float) nt[1] / 65535.0f);
				tex_coords[i] = mk2(fmaf((float) nt[2] / 65535.0f, 8.0f, 0.0f), fmaf((float) nt[3] / 65535.0f, -8.0f, 1.0f));
			}
			v3 edges[2] = { sub3(positions[1], positions[0]), sub3(positions[2], positions[0]) };
			v3 ray_cross_edge_1 = cross3(ray_direction, edges[1]);
			float rcp_det = 1.0f / dot3(edges[0], ray_cross_edge_1);
			v3 ray_to_0 = sub3(cam, positions[0]);
			float det_0_dir_edge_1 = dot3(ray_to_0, ray_cross_edge_1);
			float by = rcp_det * det_0_dir_edge_1;
			v3 edge_0_cross_0 = cross3(edges[0], ray_to_0);
			float det_dir_edge_0_0 = dot3(ray_direction, edge_0_cross_0);
			float bz = -rcp_det * det_dir_edge_0_0;
			float bx = 1.0f - (by + bz);
			v3 position = mk3(
				fmaf(bx, positions[0].x, fmaf(by, positions[1].x, bz * positions[2].x)),
				fmaf(bx, positions[0].y, fmaf(by, positions[1].y, bz * positions[2].y)),
				fmaf(bx, positions[0].z, fmaf(by, positions[1].z, bz * positions[2].z)));
			v3 interpolated_normal = normalize3(mk3(
				fmaf(bx, normals[0].x, fmaf(by, normals[1].x, bz * normals[2].x)),
				fmaf(bx, normals[0].y, fmaf(by, normals[1].y, bz * normals[2].y)),
				fmaf(bx, normals[0].z, fmaf(by, normals[1].z, bz * normals[2].z))));
			float tex_base[4], tex_specular[4], tex_normal[4];
			const uint32_t material_index = material_indices[prim];
			if (textures) {
				/* screen-space derivatives of the barycentrics and the texture coordinate (:754-777), then three textureGrad (:779-783) */
				v2 tex_coord = mk2(fmaf(bx, tex_coords[0].x, fmaf(by, tex_coords[1].x, bz * tex_coords[2].x)), fmaf(bx, tex_coords[0].y, fmaf(by, tex_coords[1].y, bz * tex_coords[2].y)));
				v2 tex_coord_derivs[2];
				for (int i = 0; i != 2; ++i) {
					v3 ray_direction_deriv = mk3(p2r[0][i], p2r[1][i], p2r[2][i]);
					v3 ray_cross_edge_1_deriv = cross3(ray_direction_deriv, edges[1]);
					float rcp_det_deriv = -dot3(edges[0], ray_cross_edge_1_deriv) * rcp_det * rcp_det;
					float det_0_dir_edge_1_deriv = dot3(ray_to_0, ray_cross_edge_1_deriv);
					float dy_ = rcp_det_deriv * det_0_dir_edge_1 + rcp_det * det_0_dir_edge_1_deriv;
					float det_dir_edge_0_0_deriv = dot3(ray_direction_deriv, edge_0_cross_0);
					float dz_ = -rcp_det_deriv * det_dir_edge_0_0 - rcp_det * det_dir_edge_0_0_deriv;
					float dx_ = -(dy_ + dz_);
					v2 d = mk2(0.0f, 0.0f);
					d = add2(d, scale2(tex_coords[0], dx_)); d = add2(d, scale2(tex_coords[1], dy_)); d = add2(d, scale2(tex_coords[2], dz_));
					tex_coord_derivs[i] = d;
				}
				vkr_texture_grad(tex_base, &textures[3 * material_index + 0], tex_coord, tex_coord_derivs[0], tex_coord_derivs[1]);
				vkr_texture_grad(tex_specular, &textures[3 * material_index + 1], tex_coord, tex_coord_derivs[0], tex_coord_derivs[1]);
				vkr_texture_grad(tex_normal, &textures[3 * material_index + 2], tex_coord, tex_coord_derivs[0], tex_coord_derivs[1]);
			}
			else { /* constant textures: the derivatives do not influence the fetch */
				const float* mp = material_params + 8 * (size_t) material_index;
				tex_base[0] = mp[0]; tex_base[1] = mp[1]; tex_base[2] = mp[2]; tex_specular[1] = mp[3]; tex_specular[2] = mp[4]; tex_normal[0] = mp[5]; tex_normal[1] = mp[6];
			}
			v3 base_color = mk3(tex_base[0], tex_base[1], tex_base[2]);
			float linear_roughness = tex_specular[1], metalicity = tex_specular[2];
			v3 nts; nts.x = fmaf(tex_normal[0], 2.0f, -1.0f); nts.y = fmaf(tex_normal[1], 2.0f, -1.0f);
			nts.z = sqrtf(vkr_max(0.0f, fmaf(-nts.x, nts.x, fmaf(-nts.y, nts.y, 1.0f))));
			v3 diffuse_albedo = mk3(fmaf(base_color.x, -metalicity, base_color.x), fmaf(base_color.y, -metalicity, base_color.y), fmaf(base_color.z, -metalicity, base_color.z));
			/* mix(x, y, a) = x*(1-a) + y*a */
			v3 fresnel_0 = mk3(0.02f * (1.0f - metalicity) + base_color.x * metalicity, 0.02f * (1.0f - metalicity) + base_color.y * metalicity, 0.02f * (1.0f - metalicity) + base_color.z * metalicity);
			float roughness = linear_roughness * linear_roughness;
			roughness = vkr_clamp(roughness * roughness_factor, 0.0064f, 1.0f);
			v2 tce[2] = { sub2(tex_coords[1], tex_coords[0]), sub2(tex_coords[2], tex_coords[0]) };
			v3 normal_cross_edge_0 = cross3(interpolated_normal, edges[0]);
			v3 edge1_cross_normal = cross3(edges[1], interpolated_normal);
			v3 tangent = add3(scale3(edge1_cross_normal, tce[0].x), scale3(normal_cross_edge_0, tce[1].x));
			v3 bitangent = add3(scale3(edge1_cross_normal, tce[0].y), scale3(normal_cross_edge_0, tce[1].y));
			float mean_tangent_length = sqrtf(0.5f * (dot3(tangent, tangent) + dot3(bitangent, bitangent)));
			nts.z *= vkr_max(1.0e-10f, mean_tangent_length);
			v3 n = normalize3(mk3(
				fmaf(interpolated_normal.x, nts.z, fmaf(bitangent.x, nts.y, tangent.x * nts.x)),
				fmaf(interpolated_normal.y, nts.z, fmaf(bitangent.y, nts.y, tangent.y * nts.x)),
				fmaf(interpolated_normal.z, nts.z, fmaf(bitangent.z, nts.y, tangent.z * nts.x))));
			v3 outgoing = normalize3(sub3(cam, position));
			float normal_offset = vkr_max(0.0f, 1.0e-3f - dot3(n, outgoing));
			n = mk3(fmaf(normal_offset, outgoing.x, n.x), fmaf(normal_offset, outgoing.y, n.y), fmaf(normal_offset, outgoing.z, n.z));
			n = normalize3(n);
			out_gbuffer[pi + 0] = position.x; out_gbuffer[pi + 1] = position.y; out_gbuffer[pi + 2] = position.z; out_gbuffer[pi + 3] = roughness;
			out_gbuffer[plane + pi + 0] = n.x; out_gbuffer[plane + pi + 1] = n.y; out_gbuffer[plane + pi + 2] = n.z; out_gbuffer[plane + pi + 3] = 1.0f;
			out_gbuffer[2 * plane + pi + 0] = diffuse_albedo.x; out_gbuffer[2 * plane + pi + 1] = diffuse_albedo.y; out_gbuffer[2 * plane + pi + 2] = diffuse_albedo.z;
			out_gbuffer[3 * plane + pi + 0] = fresnel_0.x; out_gbuffer[3 * plane + pi + 1] = fresnel_0.y; out_gbuffer[3 * plane + pi + 2] = fresnel_0.z;
		}
	return 0;
}

int vkr_oracle_gbuffer(uint32_t width, uint32_t height, const void* constants, const uint32_t* visibility,
	const uint32_t* quantized_positions, const uint16_t* normals_and_tex_coords, const uint8_t* material_indices,
	const float* material_params, float* out_gbuffer)
{
	return gbuffer_impl(width, height, constants, visibility, quantized_positions, normals_and_tex_coords, material_indices, material_params, NULL, out_gbuffer);
}

/* texture_dims: {width, height, mip_count} per texture, texture_offsets: first float of level 0 in texture_data; 3 textures per material */
int vkr_oracle_gbuffer_textured(uint32_t width, uint32_t height, const void* constants, const uint32_t* visibility,
	const uint32_t* quantized_positions, const uint16_t* normals_and_tex_coords, const uint8_t* material_indices,
	uint32_t texture_count, const uint32_t* texture_dims, const uint64_t* texture_offsets, const float* texture_data, float* out_gbuffer)
{
	vkr_texture_view_t* views = (vkr_texture_view_t*) calloc(texture_count ? texture_count : 1, sizeof(vkr_texture_view_t));
	for (uint32_t i = 0; i != texture_count; ++i) {
		views[i].width = texture_dims[3 * i]; views[i].height = texture_dims[3 * i + 1]; views[i].mip_count = texture_dims[3 * i + 2];
		views[i].texels = texture_data + texture_offsets[i];
	}
	int result = gbuffer_impl(width, height, constants, visibility, quantized_positions, normals_and_tex_coords, material_indices, NULL, views, out_gbuffer);
	free(views);
	return result;
}

/* one textureGrad per row of inputs {u, v, dudx, dvdx, dudy, dvdy} (tests) */
void vkr_oracle_texture_grad_batch(uint32_t width, uint32_t height, uint32_t mip_count, const float* texels, uint32_t n, const float* inputs, float* out_rgba) {
	vkr_texture_view_t view = { width, height, mip_count, texels };
	for (uint32_t i = 0; i != n; ++i) {
		const float* in = inputs + 6 * (size_t) i;
		vkr_texture_grad(out_rgba + 4 * (size_t) i, &view, mk2(in[0], in[1]), mk2(in[2], in[3]), mk2(in[4], in[5]));
	}
}

/* ---- small entry points for the known-answer tests */
/* Probe for tests/test_device_on_host.py: one (light, shading point) pair, n samples of a related-work technique.
   frame = rows x, y, z of world_to_shading_space and its translation column. Returns 0 if the light is culled. */
int vkr_oracle_related_work_batch(uint32_t technique, uint32_t maxv, const void* light_block, const float* position, const float* frame,
	uint32_t n, const float* random_numbers, float* out_dirs, float* out_densities, float* out_ggx_density_factor)
{
	light_t light; memset(&light, 0, sizeof(light));
	parse_light(&light, (const uint8_t*) light_block, maxv);
	float w2s[4][3];
	for (int row = 0; row != 3; ++row) {
		for (int col = 0; col != 3; ++col) w2s[col][row] = frame[3 * row + col];
		w2s[3][row] = frame[9 + row];
	}
	rw_sampler_t sampler;
	if (!rw_sampler_prepare(&sampler, technique, technique_maxp(technique, maxv), maxv, &light, mk3(position[0], position[1], position[2]), w2s)) return 0;
	*out_ggx_density_factor = sampler.ggx_density_factor;
	for (uint32_t i = 0; i != n; ++i) {
		v3 d = rw_sampler_sample(&sampler, mk2(random_numbers[2 * i], random_numbers[2 * i + 1]), &out_densities[i]);
		out_dirs[3 * i] = d.x; out_dirs[3 * i + 1] = d.y; out_dirs[3 * i + 2] = d.z;
	}
	return 1;
}

/* Probe for tests/test_device_on_host.py: clip + prepare + one sample and its error per random number pair, colour of error component 0 */
int vkr_oracle_error_display_batch(uint32_t technique, int biased, uint32_t maxv, uint32_t vertex_count, const float* vertices_xyz, uint32_t n, const float* rnd,
	float error_factor, float* out_errors, float* out_colors)
{
	const uint32_t maxp = maxv + 1;
	v3 verts[PSA_MAXP];
	memset(verts, 0, sizeof(verts));
	for (uint32_t i = 0; i != vertex_count; ++i) verts[i] = mk3(vertices_xyz[3 * i], vertices_xyz[3 * i + 1], vertices_xyz[3 * i + 2]);
	for (uint32_t i = vertex_count; i < maxv; ++i) verts[i] = verts[0];
	uint32_t cvc = psa_clip_polygon(vertex_count, verts, maxp);
	if (cvc == 0) return 0;
	psa_polygon_t polygon; rw_psa_arvo_t arvo;
	if (technique == VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO) { rw_prepare_psa_arvo(&arvo, cvc, verts, maxp); if (arvo.projected_solid_angle <= 0.0f) return 0; }
	else { psa_prepare(&polygon, cvc, verts, maxp, biased); if (polygon.projected_solid_angle <= 0.0f) return 0; }
	ctx_t c; memset(&c, 0, sizeof(c));
	c.error_factor = error_factor;
	for (uint32_t i = 0; i != n; ++i) {
		v2 r = mk2(rnd[2 * i], rnd[2 * i + 1]);
		v3 e;
		if (technique == VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO) {
			v2 e2 = rw_psa_arvo_sampling_error(&arvo, r, rw_sample_psa_arvo(&arvo, r, 3, maxp), maxp);
			e = mk3(e2.x, e2.y, 0.0f);
		}
		else e = psa_sampling_error(&polygon, r, psa_sample(&polygon, r, maxp, biased), maxp, biased);
		v3 col = error_to_color(e.x, &c);
		out_errors[3 * i] = e.x; out_errors[3 * i + 1] = e.y; out_errors[3 * i + 2] = e.z;
		out_colors[3 * i] = col.x; out_colors[3 * i + 1] = col.y; out_colors[3 * i + 2] = col.z;
	}
	return 1;
}

uint32_t vkr_oracle_clip(uint32_t vertex_count, float* vertices_xyz, uint32_t maxp) {
	v3 v[PSA_MAXP]; memset(v, 0, sizeof(v));
	for (uint32_t i = 0; i != maxp; ++i) v[i] = mk3(vertices_xyz[3 * i], vertices_xyz[3 * i + 1], vertices_xyz[3 * i + 2]);
	uint32_t vc = psa_clip_polygon(vertex_count, v, maxp);
	for (uint32_t i = 0; i != maxp; ++i) { vertices_xyz[3 * i] = v[i].x; vertices_xyz[3 * i + 1] = v[i].y; vertices_xyz[3 * i + 2] = v[i].z; }
	return vc;
}

/* Clips, prepares and draws n samples; out_dirs = 3n floats; out_info = {psa, central, vc, sectors[8]} */
void vkr_oracle_psa_sample_batch(uint32_t vertex_count, const float* vertices_xyz, uint32_t maxp, int biased, int do_clip,
	uint32_t n, const float* random_numbers, float* out_dirs, float* out_errors, float* out_info)
{
	v3 v[PSA_MAXP]; memset(v, 0, sizeof(v));
	for (uint32_t i = 0; i != maxp && i != vertex_count + 1; ++i) v[i] = mk3(vertices_xyz[3 * i], vertices_xyz[3 * i + 1], vertices_xyz[3 * i + 2]);
	uint32_t vc = vertex_count;
	if (do_clip) vc = psa_clip_polygon(vertex_count, v, maxp);
	else if (vc < maxp) v[vc] = v[0];
	psa_polygon_t p; memset(&p, 0, sizeof(p));
	if (vc) psa_prepare(&p, vc, v, maxp, biased);
	if (out_info) {
		out_info[0] = p.projected_solid_angle; out_info[1] = (float) (vc && psa_is_central_case(&p)); out_info[2] = (float) vc;
		for (int i = 0; i != 8; ++i) out_info[3 + i] = p.sector_projected_solid_angles[i];
	}
	if (!vc) return;
	for (uint32_t i = 0; i != n; ++i) {
		v2 r = mk2(random_numbers[2 * i], random_numbers[2 * i + 1]);
		v3 d = psa_sample(&p, r, maxp, biased);
		out_dirs[3 * i] = d.x; out_dirs[3 * i + 1] = d.y; out_dirs[3 * i + 2] = d.z;
		if (out_errors) { v3 e = psa_sampling_error(&p, r, d, maxp, biased); out_errors[2 * i] = e.x; out_errors[2 * i + 1] = e.y; }
	}
}

void vkr_oracle_sort_network(uint32_t vertex_count, uint32_t maxp, float* vertices_xy, float* ellipses_xy) {
	psa_polygon_t p; memset(&p, 0, sizeof(p));
	p.vertex_count = vertex_count;
	for (uint32_t i = 0; i != vertex_count; ++i) { p.vertices[i] = mk2(vertices_xy[2 * i], vertices_xy[2 * i + 1]); p.ellipses[i] = mk2(ellipses_xy[2 * i], ellipses_xy[2 * i + 1]); }
	psa_sort_convex_polygon_vertices(&p, maxp);
	for (uint32_t i = 0; i != vertex_count; ++i) { vertices_xy[2 * i] = p.vertices[i].x; vertices_xy[2 * i + 1] = p.vertices[i].y; ellipses_xy[2 * i] = p.ellipses[i].x; ellipses_xy[2 * i + 1] = p.ellipses[i].y; }
}

float vkr_oracle_kahan(float a, float b, float c, float d) { return psa_kahan(a, b, c, d); }
float vkr_oracle_atan(float x) { return vkr_atan(x); }
float vkr_oracle_sin(float x) { return vkr_sin(x); }
float vkr_oracle_cos(float x) { return vkr_cos(x); }
float vkr_oracle_acos01(float x) { return vkr_acos01(x); }
float vkr_oracle_fast_positive_atan(float x) { return psa_fast_positive_atan(x); }
void vkr_oracle_elementary_batch(int which, uint32_t n, const float* x, float* y) {
	for (uint32_t i = 0; i != n; ++i)
		y[i] = (which == 0) ? vkr_atan(x[i]) : (which == 1) ? vkr_sin(x[i]) : (which == 2) ? vkr_cos(x[i]) : (which == 3) ? vkr_acos01(x[i]) : (which == 4) ? vkr_rsqrt(x[i])
			: (which == 6) ? vkr_log2(x[i]) : (which == 7) ? vkr_exp2(x[i]) : (which == 8) ? vkr_linear_to_srgb(x[i]) : (which == 9) ? vkr_srgb_to_linear(x[i])
			: (which == 10) ? (float) vkr_float_to_half(x[i]) : (which == 11) ? vkr_acos(x[i]) : (which == 12) ? (vkr_atan2(x[i], 0.5f) + vkr_atan2(0.5f, x[i]))
			: (which == 13) ? vkr_pow(x[i], 1.0f / 3.0f) : psa_fast_positive_atan(x[i]);
}

/* Shadow predicate KATs: per ray {ox,oy,oz,dx,dy,dz,tmin,tmax} -> occluded bit via BVH and (optionally) brute force */
void vkr_oracle_trace_any(const float* tris, uint32_t tri_count, uint32_t ray_count, const float* rays, uint8_t* out_bvh, uint8_t* out_brute) {
	obvh_t bvh; obvh_build(&bvh, tris, tri_count);
	#pragma omp parallel for schedule(dynamic, 256)
	for (uint32_t i = 0; i < ray_count; ++i) {
		const float* r = rays + 8 * (size_t) i;
		v3 o = mk3(r[0], r[1], r[2]), d = mk3(r[3], r[4], r[5]);
		out_bvh[i] = (uint8_t) obvh_occluded(&bvh, o, d, r[6], r[7]);
		if (out_brute) out_brute[i] = (uint8_t) ((r[7] > r[6]) ? oracle_occluded_brute(tris, tri_count, o, d, r[6], r[7]) : 0);
	}
	obvh_destroy(&bvh);
}

int vkr_oracle_thread_count(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}
