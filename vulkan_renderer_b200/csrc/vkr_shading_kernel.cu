// vkr_shading_kernel.cu -- the per-screen-tile shading megakernel (sm_100a).
//
// Replaces subpass 1 of the reference frame (src/main.c:1429-1434, the fragment shader
// src/shaders/shading_pass.frag.glsl:824-866 with everything it calls). One CTA shades one
// 16x8 pixel tile; a warp covers an 8x4 pixel patch so that shadow rays of a warp start close
// together and head for the same light. Per frame the kernel reads
//   - the G-buffer as four coalesced float4 planes (64 B/pixel),
//   - the per-frame constant block incl. all polygonal lights (bytes identical to what the
//     reference's write_constants() produces, src/main.c:2114-2188) -> staged once per CTA into
//     shared memory with a bulk async copy (cp.async.bulk + mbarrier, the TMA engine),
//   - 4 LTC texels per pixel (software bilinear, fp32 weights), one RGBA16 noise texel per two
//     2D random numbers,
//   - BVH node pairs and triangles along the shadow rays,
// and writes one float4 of linear radiance per pixel (16 B/pixel).
// Compile with -fmad=false (see vkr_device_math.cuh).
#include "vkr_psa.cuh"
#include "vkr_trace.cuh"
#include "vkr_kernels.h"

namespace vkr {

// Byte offsets in the per-frame constant block (src/main.h:488-505, shared_constants.glsl:20-66)
enum {
	OFF_PIXEL_TO_RAY = 96, OFF_CAMERA = 144, OFF_MIS_VIS = 156, OFF_EXPOSURE = 176,
	OFF_NOISE_RES_MASK = 184, OFF_NOISE_LAYER_MASK = 192, OFF_NOISE_RANDOM = 208, OFF_LTC = 224, CONSTANTS_FIXED = 256,
	// inside one light block (polygonal_light_utility.glsl:26-83)
	L_SURFACE_RADIANCE = 48, L_PLANE = 64, L_VERTEX_COUNT = 80, L_FIXED = 160
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

struct trace_context {
	bvh_view bvh;
	int* stack;
	int stride;
	bool enabled;
};

// radiance * BRDF * visibility for a world-space direction (shading_pass.frag.glsl:120-138, 203-231)
template <bool DIFFUSE, bool SPECULAR>
VKR_DEV f3 radiance_visibility_brdf(bool* out_visibility, float* out_lambert, f3 dir_world, const shading_point& sp, const unsigned char* light, const trace_context& tc) {
	const float lambert = dot(sp.normal, dir_world);
	bool visibility = lambert > 0.0f;
	if (tc.enabled && visibility) {
		const float num = dot4_point(light + L_PLANE, sp.position);
		const float den = dot(dir_world, make3(ldf(light, L_PLANE), ldf(light, L_PLANE + 4), ldf(light, L_PLANE + 8)));
		const float max_t = -num / den;
		visibility = !occluded(tc.bvh, sp.position, dir_world, 1.0e-3f, max_t, tc.stack, tc.stride);
	}
	*out_visibility = visibility;
	*out_lambert = lambert;
	if (!visibility) return make3(0.0f, 0.0f, 0.0f);
	const f3 radiance = make3(ldf(light, L_SURFACE_RADIANCE), ldf(light, L_SURFACE_RADIANCE + 4), ldf(light, L_SURFACE_RADIANCE + 8));
	return radiance * evaluate_brdf<DIFFUSE, SPECULAR>(sp, dir_world);
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

// One polygonal light for one pixel (shading_pass.frag.glsl:329-711, projected solid angle technique)
template <int STRATEGY, int MAXP, bool BIASED>
VKR_DEV f3 shade_light(const shading_point& sp, const ltc_state& l, const unsigned char* light, noise_stream& ns,
	const shading_kernel_params& p, const unsigned char* cb, uint32_t px, uint32_t py, const trace_context& tc)
{
	const int S = p.sample_count;
	const bool flip = dot4_point(light + L_PLANE, sp.position) < 0.0f;
	f3 result = make3(0.0f, 0.0f, 0.0f);
	bool vis; float lambert;
	psa_polygon<MAXP> pd;
	{
		f3 v[MAXP];
		const int vc = transform_and_clip<MAXP>(v, light, l.rx, l.ry, sp.normal, l.t, flip);
		if (vc == 0) return result;
		prepare_psa<MAXP, BIASED>(pd, vc, v);
	}
	if (STRATEGY == VKR_STRATEGY_DIFFUSE_ONLY || STRATEGY == VKR_STRATEGY_DIFFUSE_GGX_MIS) {
		if (pd.psa <= 0.0f) return result;
#pragma unroll 1
		for (int s = 0; s != S; ++s) {
			const f3 d = sample_psa<MAXP, BIASED>(pd, next_noise_2(ns, p, cb, px, py));
			const float density = d.z / pd.psa;
			const f3 w = shading_to_world(l, sp.normal, flip, d);
			const f3 rtb = radiance_visibility_brdf<true, true>(&vis, &lambert, w, sp, light, tc);
			if (STRATEGY == VKR_STRATEGY_DIFFUSE_ONLY) {
				if (density > 0.0f) result = result + rtb * (lambert / density);
			}
			else {
				const float ggx_density = ggx_reflected_direction_density(sp.lambert_outgoing, sp.outgoing, w, sp.normal, sp.roughness);
				const float wgt = (p.mis_heuristic == VKR_MIS_BALANCE) ? (1.0f / (density + ggx_density)) : (density / (density * density + ggx_density * ggx_density));
				result = make3(result.x + rtb.x * lambert * wgt, result.y + rtb.y * lambert * wgt, result.z + rtb.z * lambert * wgt);
			}
		}
		if (STRATEGY == VKR_STRATEGY_DIFFUSE_GGX_MIS) {
			f3 o_ss = make3(
				fmaf(l.t.x, 0.0f, fmaf(l.rx.z, sp.outgoing.z, fmaf(l.rx.y, sp.outgoing.y, l.rx.x * sp.outgoing.x))),
				0.0f,
				fmaf(l.t.z, 0.0f, fmaf(sp.normal.z, sp.outgoing.z, fmaf(sp.normal.y, sp.outgoing.y, sp.normal.x * sp.outgoing.x))));
			const float density_factor = 1.0f / pd.psa;
#pragma unroll 1
			for (int s = 0; s != S; ++s) {
				float ggx_density;
				const f3 d = sample_ggx_reflected_direction(&ggx_density, o_ss, sp.roughness, next_noise_2(ns, p, cb, px, py));
				const f3 w = shading_to_world(l, sp.normal, flip, d);
				if (d.z > 0.0f && light_ray_intersection<MAXP - 1>(light, sp.position, w, 0.0f)) {
					const f3 rtb = radiance_visibility_brdf<true, true>(&vis, &lambert, w, sp, light, tc);
					const float polygon_density = lambert * density_factor;
					const float wgt = (p.mis_heuristic == VKR_MIS_BALANCE) ? (1.0f / (ggx_density + polygon_density)) : (ggx_density / (ggx_density * ggx_density + polygon_density * polygon_density));
					result = make3(result.x + rtb.x * lambert * wgt, result.y + rtb.y * lambert * wgt, result.z + rtb.z * lambert * wgt);
				}
			}
		}
	}
	else {
		psa_polygon<MAXP> ps;
		ps.psa = 0.0f;
		{
			f3 v[MAXP];
			const int vc = transform_and_clip<MAXP>(v, light, l.cx, l.cy, l.cz, l.ct, flip);
			if (vc != 0) prepare_psa<MAXP, BIASED>(ps, vc, v);
		}
		if (pd.psa == 0.0f) return result;
		const float specular_albedo = l.albedo;
		const float specular_weight = specular_albedo * ps.psa;
		const bool has_specular = ps.psa > 0.0f;
		if (STRATEGY == VKR_STRATEGY_DIFFUSE_SPECULAR_SEPARATELY) {
#pragma unroll 1
			for (int s = 0; s != S; ++s) {
				const f3 dd = sample_psa<MAXP, BIASED>(pd, next_noise_2(ns, p, cb, px, py));
				const f3 rtb = radiance_visibility_brdf<true, false>(&vis, &lambert, shading_to_world(l, sp.normal, flip, dd), sp, light, tc);
				result = result + rtb * pd.psa;
				if (has_specular) {
					const f3 dc = sample_psa<MAXP, BIASED>(ps, next_noise_2(ns, p, cb, px, py));
					const f3 dsh = normalize(c2s_mul(l, dc));
					const float ltc_density = evaluate_ltc_density(l, dsh, 1.0f);
					const f3 rtb2 = radiance_visibility_brdf<false, true>(&vis, &lambert, shading_to_world(l, sp.normal, flip, dsh), sp, light, tc);
					if (!(dsh.z <= 0.0f || dc.z <= 0.0f)) {
						result.x += rtb2.x * dsh.z * ps.psa / ltc_density;
						result.y += rtb2.y * dsh.z * ps.psa / ltc_density;
						result.z += rtb2.z * dsh.z * ps.psa / ltc_density;
					}
				}
			}
		}
		else if (STRATEGY == VKR_STRATEGY_DIFFUSE_SPECULAR_MIS) {
			f3 diffuse_weight = make3(max_glsl(sp.diffuse_albedo.x, 0.01f), max_glsl(sp.diffuse_albedo.y, 0.01f), max_glsl(sp.diffuse_albedo.z, 0.01f)) * pd.psa;
			const float rcp_d = 1.0f / pd.psa;
			const float rcp_s = 1.0f / ps.psa;
			f3 specular_weight_rgb = make3(specular_weight, specular_weight, specular_weight);
			if (p.mis_heuristic == VKR_MIS_OPTIMAL) {
				const f3 radiance_over_pi = make3(ldf(light, L_SURFACE_RADIANCE), ldf(light, L_SURFACE_RADIANCE + 4), ldf(light, L_SURFACE_RADIANCE + 8)) * kInvPi;
				diffuse_weight = diffuse_weight * radiance_over_pi;
				specular_weight_rgb = specular_weight_rgb * radiance_over_pi;
			}
			const float v_est = ldf(cb, OFF_MIS_VIS);
#pragma unroll 1
			for (int s = 0; s != S; ++s) {
				const f3 dir_d = sample_psa<MAXP, BIASED>(pd, next_noise_2(ns, p, cb, px, py));
				f3 dir_s = make3(0.0f, 0.0f, 0.0f);
				if (has_specular) dir_s = normalize(c2s_mul(l, sample_psa<MAXP, BIASED>(ps, next_noise_2(ns, p, cb, px, py))));
#pragma unroll 1
				for (int j = 0; j != (has_specular ? 2 : 1); ++j) {
					const f3 d = (j == 0) ? dir_d : dir_s;
					if (d.z <= 0.0f) continue;
					const float diffuse_density = d.z * rcp_d;
					const float specular_density = evaluate_ltc_density(l, d, rcp_s);
					const f3 integrand = radiance_visibility_brdf<true, true>(&vis, &lambert, shading_to_world(l, sp.normal, flip, d), sp, light, tc) * d.z;
					if (j == 0 && !has_specular) {
						if (vis) result = result + integrand * (1.0f / diffuse_density);
					}
					else if (j == 0)
						result = result + mis_estimate(p.mis_heuristic, integrand, diffuse_weight, diffuse_density, specular_weight_rgb, specular_density, v_est);
					else
						result = result + mis_estimate(p.mis_heuristic, integrand, specular_weight_rgb, specular_density, diffuse_weight, diffuse_density, v_est);
				}
			}
		}
		else { // VKR_STRATEGY_DIFFUSE_SPECULAR_RANDOM
			const float diffuse_albedo = max_glsl(dot(sp.diffuse_albedo, make3(0.21263901f, 0.71516868f, 0.07219232f)), 0.01f);
			const float diffuse_weight = diffuse_albedo * pd.psa;
			const float diffuse_ratio = diffuse_weight / (diffuse_weight + specular_weight);
#pragma unroll 1
			for (int s = 0; s != S; ++s) {
				f2 rnd = next_noise_2(ns, p, cb, px, py);
				const bool specular_selected = rnd.x >= diffuse_ratio;
				const float offset = specular_selected ? 1.0f : 0.0f;
				rnd.x = (rnd.x - offset) / (diffuse_ratio - offset);
				f3 d = specular_selected ? sample_psa<MAXP, BIASED>(ps, rnd) : sample_psa<MAXP, BIASED>(pd, rnd);
				if (specular_selected) d = normalize(c2s_mul(l, d));
				const float diffuse_density = d.z * diffuse_albedo;
				const float specular_density = evaluate_ltc_density(l, d, specular_albedo);
				const float density = (diffuse_density + specular_density) / (diffuse_weight + specular_weight);
				const f3 rtb = radiance_visibility_brdf<true, true>(&vis, &lambert, shading_to_world(l, sp.normal, flip, d), sp, light, tc);
				if (!(d.z <= 0.0f)) {
					result.x += rtb.x * d.z / density;
					result.y += rtb.y * d.z / density;
					result.z += rtb.z * d.z / density;
				}
			}
		}
	}
	return result * (1.0f / (float) S);
}

constexpr int kTileW = 16, kTileH = 8, kThreads = kTileW * kTileH;

template <int STRATEGY, int MAXP, bool BIASED>
__global__ void __launch_bounds__(kThreads)
shading_kernel(const shading_kernel_params p) {
	extern __shared__ __align__(16) unsigned char smem[];
	unsigned char* cb = smem;                                   // constant block incl. lights
	int* stack = reinterpret_cast<int*>(smem + p.constants_smem_bytes);
	__shared__ __align__(8) unsigned long long mbar;
	// --- stage the constant block with one bulk async copy (TMA engine), completion on an mbarrier
	const uint32_t mbar_addr = (uint32_t) __cvta_generic_to_shared(&mbar);
	const uint32_t cb_addr = (uint32_t) __cvta_generic_to_shared(cb);
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar_addr));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_addr), "r"(p.constants_bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(cb_addr), "l"(p.constants), "r"(p.constants_bytes), "r"(mbar_addr) : "memory");
	}
	{
		uint32_t done = 0;
		while (!done) {
			asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(mbar_addr) : "memory");
		}
	}
	// --- pixel of this thread: warps cover 8x4 patches of the 16x8 tile
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
	const int tiles_x = (p.width + kTileW - 1) / kTileW;
	const int tile = blockIdx.x;
	const int x = (tile % tiles_x) * kTileW + lx;
	const int y = (p.tile_row_first + (tile / tiles_x) * p.tile_row_step) * kTileH + ly;
	if (x >= p.width || y >= p.height) return;
	const size_t pixel = (size_t) y * p.width + x;
	const size_t plane = (size_t) p.width * p.height;
	const float4 g0 = __ldg(p.gbuffer + pixel), g1 = __ldg(p.gbuffer + plane + pixel);
	const bool valid = g1.w != 0.0f;
	const f3 camera = make3(ldf(cb, OFF_CAMERA), ldf(cb, OFF_CAMERA + 4), ldf(cb, OFF_CAMERA + 8));
	const float exposure = ldf(cb, OFF_EXPOSURE);
	f3 color = make3(0.0f, 0.0f, 0.0f);
	shading_point sp;
	sp.position = make3(g0.x, g0.y, g0.z);
	sp.roughness = g0.w;
	sp.normal = make3(g1.x, g1.y, g1.z);
	const int light_stride = L_FIXED + 16 * (MAXP - 1) * 2 + 16 * (MAXP - 3);
	if (p.show_polygonal_lights) { // shading_pass.frag.glsl:841-850
		f3 end; float end_w;
		if (valid) { end = sp.position; end_w = 1.0f; }
		else {
			const float fx = (float) x, fy = (float) y;
			end = make3(
				fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 8), 1.0f, fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 4), fy, ldf(cb, OFF_PIXEL_TO_RAY) * fx)),
				fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 24), 1.0f, fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 20), fy, ldf(cb, OFF_PIXEL_TO_RAY + 16) * fx)),
				fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 40), 1.0f, fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 36), fy, ldf(cb, OFF_PIXEL_TO_RAY + 32) * fx)));
			end_w = 0.0f;
		}
		for (int li = 0; li != p.light_count; ++li) {
			const unsigned char* light = cb + CONSTANTS_FIXED + li * light_stride;
			if (light_ray_intersection<MAXP - 1>(light, camera, end, end_w))
				color = color + make3(ldf(light, L_SURFACE_RADIANCE), ldf(light, L_SURFACE_RADIANCE + 4), ldf(light, L_SURFACE_RADIANCE + 8));
		}
	}
	if (valid) {
		const float4 g2 = __ldg(p.gbuffer + 2 * plane + pixel), g3 = __ldg(p.gbuffer + 3 * plane + pixel);
		sp.diffuse_albedo = make3(g2.x, g2.y, g2.z);
		sp.fresnel_0 = make3(g3.x, g3.y, g3.z);
		sp.outgoing = normalize(camera - sp.position);
		sp.lambert_outgoing = dot(sp.normal, sp.outgoing);
		ltc_state l;
		get_ltc_coefficients(l, p, cb, sp);
		noise_stream ns;
		ns.z = 0.0f; ns.w = 0.0f; ns.available = 0; ns.sample_index = 0;
		trace_context tc;
		tc.bvh.nodes = p.bvh_nodes; tc.bvh.tris = p.bvh_tris; tc.bvh.tri_ids = nullptr; tc.bvh.tri_count = p.tri_count;
		tc.stack = stack + threadIdx.x; tc.stride = kThreads;
		tc.enabled = p.trace_shadow_rays != 0 && p.tri_count != 0;
#pragma unroll 1
		for (int li = 0; li != p.light_count; ++li) {
			const unsigned char* light = cb + CONSTANTS_FIXED + li * light_stride;
			color = color + shade_light<STRATEGY, MAXP, BIASED>(sp, l, light, ns, p, cb, (uint32_t) x, (uint32_t) y, tc);
		}
	}
	if (isnan(color.x) || isnan(color.y) || isnan(color.z) || isinf(color.x) || isinf(color.y) || isinf(color.z))
		color = make3(1.0f / exposure, 0.0f / exposure, 0.8f / exposure);
	p.out[pixel] = make_float4(color.x * exposure, color.y * exposure, color.z * exposure, 1.0f);
}

} // namespace vkr

using namespace vkr;

template <int STRATEGY, int MAXP, bool BIASED>
static cudaError_t launch_variant(const shading_kernel_params& p, cudaStream_t stream) {
	const int tiles_x = (p.width + kTileW - 1) / kTileW;
	const int tiles_y = p.tile_row_count;
	if (tiles_x <= 0 || tiles_y <= 0) return cudaSuccess;
	const size_t smem = p.constants_smem_bytes + sizeof(int) * kStackDepth * kThreads;
	auto kernel = shading_kernel<STRATEGY, MAXP, BIASED>;
	cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if (err != cudaSuccess) return err;
	kernel<<<tiles_x * tiles_y, kThreads, smem, stream>>>(p);
	return cudaGetLastError();
}

template <int MAXP, bool BIASED>
static cudaError_t launch_strategy(const shading_kernel_params& p, cudaStream_t stream) {
	switch (p.sampling_strategies) {
	case VKR_STRATEGY_DIFFUSE_ONLY: return launch_variant<VKR_STRATEGY_DIFFUSE_ONLY, MAXP, BIASED>(p, stream);
	case VKR_STRATEGY_DIFFUSE_GGX_MIS: return launch_variant<VKR_STRATEGY_DIFFUSE_GGX_MIS, MAXP, BIASED>(p, stream);
	case VKR_STRATEGY_DIFFUSE_SPECULAR_SEPARATELY: return launch_variant<VKR_STRATEGY_DIFFUSE_SPECULAR_SEPARATELY, MAXP, BIASED>(p, stream);
	case VKR_STRATEGY_DIFFUSE_SPECULAR_MIS: return launch_variant<VKR_STRATEGY_DIFFUSE_SPECULAR_MIS, MAXP, BIASED>(p, stream);
	case VKR_STRATEGY_DIFFUSE_SPECULAR_RANDOM: return launch_variant<VKR_STRATEGY_DIFFUSE_SPECULAR_RANDOM, MAXP, BIASED>(p, stream);
	default: return cudaErrorInvalidValue;
	}
}

cudaError_t vkr_launch_shading_kernel(const shading_kernel_params& p, cudaStream_t stream) {
	if (p.max_light_vertex_count == 4)
		return p.biased_sampling ? launch_strategy<5, true>(p, stream) : launch_strategy<5, false>(p, stream);
	if (p.max_light_vertex_count == 3)
		return p.biased_sampling ? launch_strategy<4, true>(p, stream) : launch_strategy<4, false>(p, stream);
	return cudaErrorInvalidValue;
}
