// vkr_shade_common.cuh -- per-pixel building blocks of the shading megakernel: constant-block access,
// LTC set-up (ltc_utility.glsl:58-108), noise stream (noise_utility.glsl:63-103), Frostbite BRDF
// (brdfs.glsl:42-88), MIS estimators (shading_pass.frag.glsl:243-293), GGX VNDF sampling
// (brdfs.glsl:127-224) and the ray / light-polygon test (polygonal_light_utility.glsl:93-112).
// Compile with -fmad=false (see vkr_device_math.cuh).
#pragma once
#include "vkr_psa.cuh"
#include "vkr_trace.cuh"
#include "vkr_kernels.h"
#include "vkr_texture.cuh"
#ifndef VKR_DEVICE_CODE_ON_HOST
#include <cuda_fp16.h>
#endif

namespace vkr {

// Byte offsets in the per-frame constant block (src/main.h:488-505, shared_constants.glsl:20-66)
enum {
	OFF_PIXEL_TO_RAY = 96, OFF_CAMERA = 144, OFF_MIS_VIS = 156, OFF_EXPOSURE = 176,
	OFF_NOISE_RES_MASK = 184, OFF_NOISE_LAYER_MASK = 192, OFF_FRAME_BITS = 196, OFF_NOISE_RANDOM = 208, OFF_LTC = 224, CONSTANTS_FIXED = 256,
	// inside one light block (polygonal_light_utility.glsl:26-83)
	L_SURFACE_RADIANCE = 48, L_PLANE = 64, L_VERTEX_COUNT = 80, L_TRANSLATION = 16, L_INV_SCALING_X = 44, L_INV_SCALING_Y = 60, L_TEXTURING = 84, L_TEXTURE_INDEX = 88, L_ROTATION = 96, L_FIXED = 160
};

struct shading_point {
	f3 position, normal, outgoing;
	float lambert_outgoing;
	f3 diffuse_albedo, fresnel_0;
	float roughness;
};

// Linearly transformed cosine state of one pixel (ltc_utility.glsl:33-50), matrices as rows
struct ltc_state {
	f3 rx, ry;          // world_to_shading rows x, y (row z = shading normal)
	f3 t;               // world_to_shading translation column
	f3 cx, cy, cz;      // world_to_cosine rows (rotation part)
	f3 ct;              // world_to_cosine translation column
	float s00, s02, s11, s20, s22;   // shading_to_cosine entries [col][row] that are not zero
	float c00, c02, c11, c20, c22;   // cosine_to_shading
	float albedo, det;
};

VKR_DEV float ldf(const unsigned char* p, int off) { return *reinterpret_cast<const float*>(p + off); }
VKR_DEV uint32_t ldu(const unsigned char* p, int off) { return *reinterpret_cast<const uint32_t*>(p + off); }

VKR_DEV float dot4_point(const unsigned char* plane, f3 p) { // dot(vec4(p,1), plane)
	return fmaf(ldf(plane, 12), 1.0f, fmaf(ldf(plane, 8), p.z, fmaf(ldf(plane, 4), p.y, ldf(plane, 0) * p.x)));
}

// shading_to_cosine * v with the zero entries kept (0*x terms decide the sign of zero results)
VKR_DEV f3 s2c_mul(const ltc_state& l, f3 v) {
	return make3(
		fmaf(l.s20, v.z, fmaf(0.0f, v.y, l.s00 * v.x)),
		fmaf(0.0f, v.z, fmaf(l.s11, v.y, 0.0f * v.x)),
		fmaf(l.s22, v.z, fmaf(0.0f, v.y, l.s02 * v.x)));
}
VKR_DEV f3 c2s_mul(const ltc_state& l, f3 v) {
	return make3(
		fmaf(l.c20, v.z, fmaf(0.0f, v.y, l.c00 * v.x)),
		fmaf(0.0f, v.z, fmaf(l.c11, v.y, 0.0f * v.x)),
		fmaf(l.c22, v.z, fmaf(0.0f, v.y, l.c02 * v.x)));
}

// Bilinear fetch from a UNORM16 2D array with fp32 weights (stand-in for textureLod with the
// sampler of src/ltc_table.c:170-177; definition in DESIGN.md)
template <int CH>
VKR_DEV void ltc_fetch(const uint16_t* __restrict__ table, int res, int layers, float u, float v, float layer_f, float* out) {
	const float layer_r = rintf(layer_f);
	const int layer = (int) clamp_glsl(layer_r, 0.0f, (float) (layers - 1));
	const float x = u * (float) res - 0.5f, y = v * (float) res - 0.5f;
	const float x0f = floorf(x), y0f = floorf(y);
	const float fx = x - x0f, fy = y - y0f;
	int x0 = (int) x0f, y0 = (int) y0f, x1 = x0 + 1, y1 = y0 + 1;
	x0 = min(max(x0, 0), res - 1); x1 = min(max(x1, 0), res - 1);
	y0 = min(max(y0, 0), res - 1); y1 = min(max(y1, 0), res - 1);
	const uint16_t* base = table + (size_t) layer * res * res * CH;
#pragma unroll
	for (int ch = 0; ch != CH; ++ch) {
		const float t00 = (float) __ldg(base + ((size_t) y0 * res + x0) * CH + ch) / 65535.0f;
		const float t10 = (float) __ldg(base + ((size_t) y0 * res + x1) * CH + ch) / 65535.0f;
		const float t01 = (float) __ldg(base + ((size_t) y1 * res + x0) * CH + ch) / 65535.0f;
		const float t11 = (float) __ldg(base + ((size_t) y1 * res + x1) * CH + ch) / 65535.0f;
		const float a = fmaf(fx, t10 - t00, t00);
		const float b = fmaf(fx, t11 - t01, t01);
		out[ch] = fmaf(fy, b - a, a);
	}
}

VKR_DEV void get_ltc_coefficients(ltc_state& l, const shading_kernel_params& p, const unsigned char* cb, const shading_point& sp) {
	const float fresnel_luminance = dot(sp.fresnel_0, make3(0.2126f, 0.7152f, 0.0722f));
	const float ndo = dot(sp.normal, sp.outgoing);
	const float inclination = acos01(clamp_glsl(ndo, 0.0f, 1.0f));
	const float tu = fmaf(sqrtf(clamp_glsl(sp.roughness, 0.0f, 1.0f)), ldf(cb, OFF_LTC + 8), ldf(cb, OFF_LTC + 12));
	const float tv = fmaf(inclination, ldf(cb, OFF_LTC + 16), ldf(cb, OFF_LTC + 20));
	const float tw = fmaf(clamp_glsl(fresnel_luminance, 0.0f, 1.0f), ldf(cb, OFF_LTC + 0), ldf(cb, OFF_LTC + 4));
	float d0[4], d1[2];
	ltc_fetch<4>(p.ltc0, p.ltc_res, p.ltc_layers, tu, tv, tw, d0);
	ltc_fetch<2>(p.ltc1, p.ltc_res, p.ltc_layers, tu, tv, tw, d1);
	l.s00 = d0[0]; l.s02 = -d0[1]; l.s11 = d0[2]; l.s20 = d0[3]; l.s22 = d1[0];
	l.albedo = d1[1];
	const float det2 = d0[0] * d1[0] + d0[1] * d0[3];
	l.det = d0[2] * det2;
	const float inv_det2 = 1.0f / det2;
	l.c00 = d1[0] * inv_det2; l.c02 = d0[1] * inv_det2; l.c11 = 1.0f / d0[2];
	l.c20 = -d0[3] * inv_det2; l.c22 = d0[0] * inv_det2;
	const f3 x_axis = normalize(make3(fmaf(-ndo, sp.normal.x, sp.outgoing.x), fmaf(-ndo, sp.normal.y, sp.outgoing.y), fmaf(-ndo, sp.normal.z, sp.outgoing.z)));
	const f3 y_axis = cross(sp.normal, x_axis);
	l.rx = x_axis; l.ry = y_axis;
	const f3 n = sp.normal, pos = sp.position;
	l.t = make3(
		fmaf(-x_axis.z, pos.z, fmaf(-x_axis.y, pos.y, -x_axis.x * pos.x)),
		fmaf(-y_axis.z, pos.z, fmaf(-y_axis.y, pos.y, -y_axis.x * pos.x)),
		fmaf(-n.z, pos.z, fmaf(-n.y, pos.y, -n.x * pos.x)));
	// world_to_cosine = shading_to_cosine * world_to_shading, one column at a time
	const f3 c0 = s2c_mul(l, make3(x_axis.x, y_axis.x, n.x));
	const f3 c1 = s2c_mul(l, make3(x_axis.y, y_axis.y, n.y));
	const f3 c2 = s2c_mul(l, make3(x_axis.z, y_axis.z, n.z));
	const f3 c3 = s2c_mul(l, l.t);
	l.cx = make3(c0.x, c1.x, c2.x); l.cy = make3(c0.y, c1.y, c2.y); l.cz = make3(c0.z, c1.z, c2.z);
	l.ct = c3;
}

// M(4x3) * (v,1) for a matrix given by rows + translation column
VKR_DEV f3 affine(f3 rx, f3 ry, f3 rz, f3 t, f3 v) {
	return make3(
		fmaf(t.x, 1.0f, fmaf(rx.z, v.z, fmaf(rx.y, v.y, rx.x * v.x))),
		fmaf(t.y, 1.0f, fmaf(ry.z, v.z, fmaf(ry.y, v.y, ry.x * v.x))),
		fmaf(t.z, 1.0f, fmaf(rz.z, v.z, fmaf(rz.y, v.y, rz.x * v.x))));
}

struct noise_stream {
	float z, w;           // second half of the last texel
	uint32_t available;   // 0 or 2
	uint32_t sample_index;
};

VKR_DEV f2 next_noise_2(noise_stream& ns, const shading_kernel_params& p, const unsigned char* cb, uint32_t px, uint32_t py) { // noise_utility.glsl:63-103
	if (ns.available == 0) {
		const uint32_t si = ns.sample_index;
		uint32_t r0, r1, r2, r3;
		if (si & 2) { r0 = ldu(cb, OFF_NOISE_RANDOM + 8); r1 = ldu(cb, OFF_NOISE_RANDOM + 12); r2 = ldu(cb, OFF_NOISE_RANDOM); r3 = ldu(cb, OFF_NOISE_RANDOM + 4); }
		else { r0 = ldu(cb, OFF_NOISE_RANDOM); r1 = ldu(cb, OFF_NOISE_RANDOM + 4); r2 = ldu(cb, OFF_NOISE_RANDOM + 8); r3 = ldu(cb, OFF_NOISE_RANDOM + 12); }
		if (si & 1) { r0 = r1; r1 = r2; r2 = r3; }
		const uint32_t shift = (si & 124u) >> 2;
		const uint32_t layer = (r2 + si) & ldu(cb, OFF_NOISE_LAYER_MASK);
		const uint32_t x = (px + (r0 >> shift)) & ldu(cb, OFF_NOISE_RES_MASK);
		const uint32_t y = (py + (r1 >> shift)) & ldu(cb, OFF_NOISE_RES_MASK + 4);
		const uint2 texel = __ldg(reinterpret_cast<const uint2*>(p.noise) + ((size_t) layer * p.noise_h + y) * p.noise_w + x);
		ns.sample_index = si + 1;
		ns.available = 2;
		ns.z = (float) (texel.y & 0xffffu) / 65535.0f;
		ns.w = (float) (texel.y >> 16) / 65535.0f;
		return make2((float) (texel.x & 0xffffu) / 65535.0f, (float) (texel.x >> 16) / 65535.0f);
	}
	ns.available = 0;
	return make2(ns.z, ns.w);
}

VKR_DEV float schlick(float f0, float f90, float cos_theta) { // brdfs.glsl:42-46
	const float flipped = 1.0f - cos_theta;
	const float f2_ = flipped * flipped;
	return f0 + (f90 - f0) * (f2_ * flipped * f2_);
}

template <bool DIFFUSE, bool SPECULAR>
VKR_DEV f3 evaluate_brdf(const shading_point& sp, f3 incoming) { // brdfs.glsl:57-88
	const f3 h = normalize(incoming + sp.outgoing);
	const float lambert_incoming = dot(sp.normal, incoming);
	const float o_dot_h = dot(sp.outgoing, h);
	f3 brdf = make3(0.0f, 0.0f, 0.0f);
	if (DIFFUSE) {
		const float f90 = fmaf(o_dot_h * o_dot_h, 2.0f * sp.roughness, 0.5f);
		const float fp = schlick(1.0f, f90, sp.lambert_outgoing) * schlick(1.0f, f90, lambert_incoming);
		brdf = brdf + sp.diffuse_albedo * fp;
	}
	if (SPECULAR) {
		const float n_dot_h = dot(sp.normal, h);
		const float r2 = sp.roughness * sp.roughness;
		float ggx = fmaf(fmaf(n_dot_h, r2, -n_dot_h), n_dot_h, 1.0f);
		ggx = r2 / (ggx * ggx);
		const float masking = lambert_incoming * sqrtf(fmaf(fmaf(-sp.lambert_outgoing, r2, sp.lambert_outgoing), sp.lambert_outgoing, r2));
		const float shadowing = sp.lambert_outgoing * sqrtf(fmaf(fmaf(-lambert_incoming, r2, lambert_incoming), lambert_incoming, r2));
		const float smith = 0.5f / (masking + shadowing);
		const float ct = clamp_glsl(o_dot_h, 0.0f, 1.0f);
		const float gs = ggx * smith;
		brdf.x += gs * schlick(sp.fresnel_0.x, 1.0f, ct);
		brdf.y += gs * schlick(sp.fresnel_0.y, 1.0f, ct);
		brdf.z += gs * schlick(sp.fresnel_0.z, 1.0f, ct);
	}
	return brdf * kInvPi;
}

VKR_DEV float evaluate_ltc_density(const ltc_state& l, f3 dir_shading, float rcp_psa) { // ltc_utility.glsl:103-108
	const f3 dc = s2c_mul(l, dir_shading);
	const float l2 = dot(dc, dc);
	const float density = max_glsl(0.0f, dc.z) * l.det / (l2 * l2);
	return density * rcp_psa;
}

// Ray-vs-light-polygon test for light display and GGX MIS (polygonal_light_utility.glsl:93-112)
template <int MAXV>
VKR_DEV bool light_ray_intersection(const unsigned char* light, f3 origin, f3 end_xyz, float end_w) {
	const float d0 = dot4_point(light + L_PLANE, origin);
	const float d1 = fmaf(ldf(light, L_PLANE + 12), end_w, fmaf(ldf(light, L_PLANE + 8), end_xyz.z, fmaf(ldf(light, L_PLANE + 4), end_xyz.y, ldf(light, L_PLANE) * end_xyz.x)));
	if (d0 * d1 > 0.0f) return false;
	const f3 dir = make3(end_xyz.x - end_w * origin.x, end_xyz.y - end_w * origin.y, end_xyz.z - end_w * origin.z);
	const unsigned char* vw = light + L_FIXED + 16 * MAXV;
	const uint32_t n = ldu(light, L_VERTEX_COUNT);
	float previous_sign = 0.0f;
	bool result = true;
#pragma unroll
	for (int i = 0; i != MAXV; ++i) {
		const int j = (i + 1) % MAXV;
		const f3 a = make3(ldf(vw, 16 * i), ldf(vw, 16 * i + 4), ldf(vw, 16 * i + 8)) - origin;
		const f3 b = make3(ldf(vw, 16 * j), ldf(vw, 16 * j + 4), ldf(vw, 16 * j + 8)) - origin;
		const float sign = det3(dir, a, b);
		result = result && ((i >= 3 && i >= (int) n) || previous_sign * sign >= 0.0f);
		previous_sign = sign;
	}
	return result;
}


VKR_DEV f3 mis_estimate(int heuristic, f3 integrand, f3 sampled_weight, float sampled_density, f3 other_weight, float other_density, float visibility_estimate) { // :270-293
	if (heuristic == VKR_MIS_WEIGHTED) {
		const f3 ws = make3(sampled_weight.x * sampled_density + other_weight.x * other_density, sampled_weight.y * sampled_density + other_weight.y * other_density, sampled_weight.z * sampled_density + other_weight.z * other_density);
		return make3((sampled_weight.x * integrand.x) / ws.x, (sampled_weight.y * integrand.y) / ws.y, (sampled_weight.z * integrand.z) / ws.z);
	}
	if (heuristic == VKR_MIS_OPTIMAL_CLAMPED || heuristic == VKR_MIS_OPTIMAL) {
		const float balance = 1.0f / (sampled_density + other_density);
		const f3 ws = make3(sampled_weight.x * sampled_density + other_weight.x * other_density, sampled_weight.y * sampled_density + other_weight.y * other_density, sampled_weight.z * sampled_density + other_weight.z * other_density);
		if (heuristic == VKR_MIS_OPTIMAL_CLAMPED) {
			const float mixed = fmaf(-visibility_estimate, balance, balance);
			return make3(
				fmaf(visibility_estimate, sampled_weight.x / ws.x, mixed) * integrand.x,
				fmaf(visibility_estimate, sampled_weight.y / ws.y, mixed) * integrand.y,
				fmaf(visibility_estimate, sampled_weight.z / ws.z, mixed) * integrand.z);
		}
		return make3(
			visibility_estimate * sampled_weight.x + balance * (integrand.x - visibility_estimate * ws.x),
			visibility_estimate * sampled_weight.y + balance * (integrand.y - visibility_estimate * ws.y),
			visibility_estimate * sampled_weight.z + balance * (integrand.z - visibility_estimate * ws.z));
	}
	const float w = (heuristic == VKR_MIS_BALANCE) ? (1.0f / (sampled_density + other_density))
		: (sampled_density / (sampled_density * sampled_density + other_density * other_density));
	return integrand * w;
}

// Transforms the light's world-space vertices with rows (rx, ry*flip, rz) + t and clips to z >= 0
template <int MAXP>
VKR_DEV int transform_and_clip(f3 (&v)[MAXP], const unsigned char* light, f3 rx, f3 ry, f3 rz, f3 t, bool flip) {
	const unsigned char* vw = light + L_FIXED + 16 * (MAXP - 1);
#pragma unroll
	for (int i = 0; i != MAXP - 1; ++i) {
		f3 q = affine(rx, ry, rz, t, make3(ldf(vw, 16 * i), ldf(vw, 16 * i + 4), ldf(vw, 16 * i + 8)));
		q.y = flip ? -q.y : q.y;
		v[i] = q;
	}
	v[MAXP - 1] = make3(0.0f, 0.0f, 0.0f);
	return clip_polygon<MAXP>((int) ldu(light, L_VERTEX_COUNT), v);
}

VKR_DEV f3 shading_to_world(const ltc_state& l, f3 n, bool flip, f3 d) { // (transpose(world_to_shading) * d).xyz
	const float dy = flip ? -d.y : d.y;
	return make3(
		fmaf(n.x, d.z, fmaf(l.ry.x, dy, l.rx.x * d.x)),
		fmaf(n.y, d.z, fmaf(l.ry.y, dy, l.rx.y * d.x)),
		fmaf(n.z, d.z, fmaf(l.ry.z, dy, l.rx.z * d.x)));
}

// GGX VNDF sampling (brdfs.glsl:127-224), only for SAMPLING_STRATEGIES_DIFFUSE_GGX_MIS
VKR_DEV float ggx_visible_normal_density(float o_dot_n, float m_dot_n, float m_dot_o, float roughness) {
	const float r2 = roughness * roughness;
	float ggx = fmaf(fmaf(m_dot_n, r2, -m_dot_n), m_dot_n, 1.0f);
	ggx = r2 / (ggx * ggx);
	ggx *= kInvPi;
	float masking = sqrtf(fmaf(fmaf(-o_dot_n, r2, o_dot_n), o_dot_n, r2));
	masking = 2.0f / (o_dot_n + masking);
	return masking * m_dot_o * ggx;
}
VKR_DEV f3 sample_ggx_reflected_direction(float* out_density, f3 o, float roughness, f2 rnd) {
	const f3 e2 = normalize(make3(roughness * o.x, roughness * o.y, 1.0f * o.z));
	const float length_sq = dot(make2(e2.x, e2.y), make2(e2.x, e2.y));
	const float rs = rsqrt_ieee(length_sq);
	f3 e0 = make3(-e2.y * rs, e2.x * rs, 0.0f * rs);
	if (length_sq <= 0.0f) e0 = make3(1.0f, 0.0f, 0.0f);
	const f3 e1 = cross(e2, e0);
	const float radius = sqrtf(rnd.x);
	const float azimuth = (2.0f * kPi) * rnd.y;
	float sa, ca;
	sincos_cw(azimuth, &sa, &ca);
	const f2 disk = make2(radius * ca, radius * sa);
	f3 s;
	s.x = disk.x;
	const float lerp_factor = fmaf(0.5f, e2.z, 0.5f);
	const float sx = sqrtf(fmaf(-disk.x, disk.x, 1.0f));
	s.y = sx * (1.0f - lerp_factor) + disk.y * lerp_factor;
	s.z = sqrtf(max_glsl(0.0f, 1.0f - dot(make2(s.x, s.y), make2(s.x, s.y))));
	const f3 h = make3(
		fmaf(e2.x, s.z, fmaf(e1.x, s.y, e0.x * s.x)),
		fmaf(e2.y, s.z, fmaf(e1.y, s.y, e0.y * s.x)),
		fmaf(e2.z, s.z, fmaf(e1.z, s.y, e0.z * s.x)));
	const f3 m = normalize(make3(roughness * h.x, roughness * h.y, 1.0f * h.z));
	const float m_dot_o = dot(m, o);
	float density = ggx_visible_normal_density(o.z, m.z, m_dot_o, roughness);
	const float two = 2.0f * m_dot_o;
	const f3 incoming = make3(fmaf(two, m.x, -o.x), fmaf(two, m.y, -o.y), fmaf(two, m.z, -o.z));
	density /= 4.0f * m_dot_o;
	*out_density = density;
	return incoming;
}
VKR_DEV float ggx_reflected_direction_density(float o_dot_n, f3 o, f3 i, f3 n, float roughness) {
	const f3 m = normalize(o + i);
	const float m_dot_o = dot(m, o);
	const float m_dot_n = dot(m, n);
	float density = ggx_visible_normal_density(o_dot_n, m_dot_n, m_dot_o, roughness);
	density /= 4.0f * m_dot_o;
	return density;
}


VKR_DEV bool is_finite(float x) { return fabsf(x) <= 3.402823466e+38f; } // false for NaN too
VKR_DEV f3 not_a_number3() { const float n = __int_as_float(0x7fc00000); return make3(n, n, n); }
VKR_DEV bool is_finite(f3 v) { return fabsf(v.x) <= 3.402823466e+38f && fabsf(v.y) <= 3.402823466e+38f && fabsf(v.z) <= 3.402823466e+38f; } // false for NaN too

// Visibility pre-test and light-plane distance of a candidate direction (shading_pass.frag.glsl:120-124, 204-205)
VKR_DEV float light_plane_distance(const shading_point& sp, const unsigned char* light, f3 dir_world) {
	const float num = dot4_point(light + L_PLANE, sp.position);
	const float den = dot(dir_world, make3(ldf(light, L_PLANE), ldf(light, L_PLANE + 4), ldf(light, L_PLANE + 8)));
	return -num / den;
}
VKR_DEV f3 light_radiance(const unsigned char* light) { return make3(ldf(light, L_SURFACE_RADIANCE), ldf(light, L_SURFACE_RADIANCE + 4), ldf(light, L_SURFACE_RADIANCE + 8)); }

// get_polygon_radiance() (shading_pass.frag.glsl:151-185): radiance received at position from direction dir (normalised, hits the light's plane).
// LIGHT_TEXTURES = false is the untextured case every benchmark configuration runs; true adds the three texturing techniques
// (polygon_texturing_technique_t, src/polygonal_light.h: 1 area, 2 portal onto a light probe, 3 IES profile).
template <bool LIGHT_TEXTURES>
VKR_DEV f3 light_radiance(const shading_kernel_params& p, const unsigned char* light, f3 position, f3 dir) {
	f3 radiance = light_radiance(light);
	if constexpr (LIGHT_TEXTURES) {
		const uint32_t technique = ldu(light, L_TEXTURING);
		if (technique != 0u) {
			// plane space = transpose(rotation) * world; rotation is stored row major with a stride of four floats
			const f3 c0 = make3(ldf(light, L_ROTATION), ldf(light, L_ROTATION + 16), ldf(light, L_ROTATION + 32));
			const f3 c1 = make3(ldf(light, L_ROTATION + 4), ldf(light, L_ROTATION + 20), ldf(light, L_ROTATION + 36));
			float u, v;
			if (technique == 1u) {
				const float num = dot4_point(light + L_PLANE, position);
				const float t = -num / dot(dir, make3(ldf(light, L_PLANE), ldf(light, L_PLANE + 4), ldf(light, L_PLANE + 8)));
				f3 intersection = position + dir * t;
				intersection = intersection - make3(ldf(light, L_TRANSLATION), ldf(light, L_TRANSLATION + 4), ldf(light, L_TRANSLATION + 8));
				u = dot(c0, intersection) * ldf(light, L_INV_SCALING_X);
				v = dot(c1, intersection) * ldf(light, L_INV_SCALING_Y);
			}
			else {
				f3 lookup;
				if (technique == 3u) {
					const f3 c2 = make3(ldf(light, L_ROTATION + 8), ldf(light, L_ROTATION + 24), ldf(light, L_ROTATION + 40));
					lookup = make3(dot(c0, dir), dot(c1, dir), dot(c2, dir));
					radiance = radiance * (1.0f / fabsf(lookup.z)); // IES profiles include the cosine already
				}
				else lookup = make3(-dir.x, dir.y, dir.z); // the layout of HDRI Haven light probes
				u = atan2_poly(lookup.y, lookup.x) * (0.5f * kInvPi);
				v = acos_full(lookup.z) * kInvPi;
			}
			const uint32_t index = ldu(light, L_TEXTURE_INDEX);
			const uint4 dims = __ldg(p.light_texture_dims + index);
			texture_view view;
			view.width = dims.x; view.height = dims.y; view.mip_count = dims.z; view.texels = p.light_texture_texels + __ldg(p.light_texture_offsets + index);
			const float4 texel = texture_bilinear_repeat_clamp(view, u, v);
			radiance = make3(radiance.x * texel.x, radiance.y * texel.y, radiance.z * texel.z);
		}
	}
	return radiance;
}

// Output stage of the shader (shading_pass.frag.glsl:871-892) on the colour that is already multiplied by the exposure: the half-bit split for HDR
// screenshots (g_frame_bits 1 / 2: the low / high bytes of the three binary16 values, as 8-bit UNORM) and the sRGB conversion of !OUTPUT_LINEAR_RGB
VKR_DEV f3 output_stage(f3 out_color, uint32_t frame_bits, bool output_srgb) {
	if (frame_bits > 0u) {
		const uint32_t mask = (frame_bits == 1u) ? 0xFFu : 0xFF00u, shift = (frame_bits == 1u) ? 0u : 8u;
		const uint32_t h0 = (uint32_t) __half_as_ushort(__float2half_rn(out_color.x)) | ((uint32_t) __half_as_ushort(__float2half_rn(out_color.y)) << 16);
		const uint32_t h1 = (uint32_t) __half_as_ushort(__float2half_rn(out_color.z));
		out_color = make3((float) ((h0 & mask) >> shift) * (1.0f / 255.0f), (float) ((((h0 & 0xFFFF0000u) >> 16) & mask) >> shift) * (1.0f / 255.0f), (float) ((h1 & mask) >> shift) * (1.0f / 255.0f));
		if (!output_srgb) out_color = make3(srgb_to_linear(out_color.x), srgb_to_linear(out_color.y), srgb_to_linear(out_color.z));
	}
	else if (output_srgb) out_color = make3(linear_to_srgb(out_color.x), linear_to_srgb(out_color.y), linear_to_srgb(out_color.z));
	return out_color;
}

} // namespace vkr
