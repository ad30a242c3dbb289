// vkr_trace.cuh -- software BVH2 traversal for shadow (any-hit) and primary (closest-hit) rays.
//
// Replaces the VK_KHR_ray_query calls of the reference (src/shaders/shading_pass.frag.glsl:120-138)
// and the driver-built acceleration structure (src/scene.c:142-406). No RT cores, no OptiX.
//
// Layout in HBM (built once on the host, vkr_bvh.cpp):
//   node  = 64 B = 4 x float4: both children's boxes + both child references ("node pair"),
//           so one 64-B sector-aligned fetch decides both children.
//             q0 = (lo0.x, lo0.y, lo0.z, hi0.x)  q1 = (hi0.y, hi0.z, lo1.x, lo1.y)
//             q2 = (lo1.z, hi1.x, hi1.y, hi1.z)  q3 = (ref0, ref1, -, -) as int bits
//           ref >= 0: inner node index; ref < 0: leaf, (ref & 0x7fffffff) = first_triangle << 4 | count
//   tri   = 48 B = 3 x float4: v0.xyz e1.x | e1.yz e2.xy | e2.z - - -   (e1 = v1 - v0, e2 = v2 - v0)
//
// The triangle predicate is the arithmetic contract of DESIGN.md (Moeller-Trumbore, fp32, fixed
// operation order, no culling, open interval); hit/miss does not depend on traversal order.
#pragma once
#include "vkr_device_math.cuh"

namespace vkr {

struct bvh_view {
	const float4* nodes;
	const float4* tris;
	const uint32_t* tri_ids; // original triangle index per slot (closest-hit only)
	uint32_t tri_count;
};

constexpr int kStackDepth = 64;

VKR_DEV bool ray_triangle(const float4* __restrict__ tri, f3 o, f3 d, float tmin, float tmax, float* out_t) {
	const float4 a = __ldg(tri), b = __ldg(tri + 1), c = __ldg(tri + 2);
	const f3 p0 = make3(a.x, a.y, a.z);
	const f3 e1 = make3(a.w, b.x, b.y);
	const f3 e2 = make3(b.z, b.w, c.x);
	const f3 pv = cross(d, e2);
	const float det = dot(e1, pv);
	if (det == 0.0f) return false;
	const float inv_det = 1.0f / det;
	const f3 tv = o - p0;
	const float u = dot(tv, pv) * inv_det;
	if (!(u >= 0.0f && u <= 1.0f)) return false;
	const f3 qv = cross(tv, e1);
	const float v = dot(d, qv) * inv_det;
	if (!(v >= 0.0f && u + v <= 1.0f)) return false;
	const float t = dot(e2, qv) * inv_det;
	if (!(t > tmin && t < tmax)) return false;
	*out_t = t;
	return true;
}

VKR_DEV bool ray_box(float lox, float loy, float loz, float hix, float hiy, float hiz, f3 o, f3 id, float tmin, float tmax) {
	float t0 = (lox - o.x) * id.x, t1 = (hix - o.x) * id.x;
	float tn = (t0 < t1) ? t0 : t1, tf = (t0 > t1) ? t0 : t1;
	float t_near = (tn > tmin) ? tn : tmin, t_far = (tf < tmax) ? tf : tmax;
	t0 = (loy - o.y) * id.y; t1 = (hiy - o.y) * id.y;
	tn = (t0 < t1) ? t0 : t1; tf = (t0 > t1) ? t0 : t1;
	t_near = (tn > t_near) ? tn : t_near; t_far = (tf < t_far) ? tf : t_far;
	t0 = (loz - o.z) * id.z; t1 = (hiz - o.z) * id.z;
	tn = (t0 < t1) ? t0 : t1; tf = (t0 > t1) ? t0 : t1;
	t_near = (tn > t_near) ? tn : t_near; t_far = (tf < t_far) ? tf : t_far;
	return t_near <= t_far * 1.0000005f;
}

// Any-hit query. stack = this thread's column of the CTA's shared-memory stack (stride = blockDim).
VKR_DEV bool occluded(const bvh_view& bvh, f3 o, f3 d, float tmin, float tmax, int* stack, int stride) {
	if (!(tmax > tmin)) return false; // undefined in Vulkan; defined as "miss" 
	const f3 id = make3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	int sp = 0;
	int node = 0;
	float t;
	while (true) {
		const float4* n = bvh.nodes + 4 * (size_t) node;
		const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
		const int ref0 = __float_as_int(q3.x), ref1 = __float_as_int(q3.y);
		bool h0 = ray_box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, id, tmin, tmax);
		bool h1 = ray_box(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, id, tmin, tmax);
		if (h0 && ref0 < 0) {
			const int first = (ref0 & 0x7fffffff) >> 4, count = ref0 & 15;
			for (int i = 0; i != count; ++i)
				if (ray_triangle(bvh.tris + 3 * (size_t) (first + i), o, d, tmin, tmax, &t)) return true;
			h0 = false;
		}
		if (h1 && ref1 < 0) {
			const int first = (ref1 & 0x7fffffff) >> 4, count = ref1 & 15;
			for (int i = 0; i != count; ++i)
				if (ray_triangle(bvh.tris + 3 * (size_t) (first + i), o, d, tmin, tmax, &t)) return true;
			h1 = false;
		}
		if (h0 && h1) { stack[sp * stride] = ref1; ++sp; node = ref0; }
		else if (h0) node = ref0;
		else if (h1) node = ref1;
		else {
			if (sp == 0) return false;
			--sp; node = stack[sp * stride];
		}
	}
}

// Closest-hit query, ties in t resolve to the lowest original triangle index (order independent).
VKR_DEV int closest_hit(const bvh_view& bvh, f3 o, f3 d, float tmin, float tmax, int* stack, int stride) {
	const f3 id = make3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	int sp = 0;
	int node = 0;
	int best = -1;
	float best_t = tmax;
	float t;
	while (true) {
		const float4* n = bvh.nodes + 4 * (size_t) node;
		const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
		const int ref0 = __float_as_int(q3.x), ref1 = __float_as_int(q3.y);
		bool h0 = ray_box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, id, tmin, best_t);
		bool h1 = ray_box(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, id, tmin, best_t);
#pragma unroll
		for (int c = 0; c != 2; ++c) {
			const int ref = c ? ref1 : ref0;
			const bool h = c ? h1 : h0;
			if (h && ref < 0) {
				const int first = (ref & 0x7fffffff) >> 4, count = ref & 15;
				for (int i = 0; i != count; ++i) {
					if (ray_triangle(bvh.tris + 3 * (size_t) (first + i), o, d, tmin, __int_as_float(0x7f800000), &t)) {
						const int id_ = (int) __ldg(bvh.tri_ids + first + i);
						if (t < best_t || (t == best_t && best >= 0 && id_ < best)) { best_t = t; best = id_; }
					}
				}
				if (c) h1 = false; else h0 = false;
			}
		}
		if (h0 && h1) { stack[sp * stride] = ref1; ++sp; node = ref0; }
		else if (h0) node = ref0;
		else if (h1) node = ref1;
		else {
			if (sp == 0) return best;
			--sp; node = stack[sp * stride];
		}
	}
}

} // namespace vkr
