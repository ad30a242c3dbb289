// vkr_trace.cuh -- software BVH2 traversal for shadow (any-hit) and primary (closest-hit) rays.
//
// Replaces the VK_KHR_ray_query calls of the reference (src/shaders/shading_pass.frag.glsl:120-138)
// and the driver-built acceleration structure (src/scene.c:142-406). No RT cores, no OptiX.
//
// Layout in HBM (built once on the host, vkr_bvh.cpp):
//   node  = 64 B = 4 x float4: both children's boxes + both child references ("node pair"),
//           so one 64-B sector-aligned fetch decides both children.
//             q0 = (c0.x, c0.y, c0.z, h0.x)  q1 = (h0.y, h0.z, c1.x, c1.y)
//             q2 = (c1.z, h1.x, h1.y, h1.z)  q3 = (ref0, ref1, -, -) as int bits
//           c = box centre, h = half extent (rounded up)
//           ref >= 0: inner node index; ref < 0: leaf, (ref & 0x7fffffff) = first_triangle << 4 | count
//   tri   = 48 B = 3 x float4: v0.xyz e1.x | e1.yz e2.xy | e2.z - - -   (e1 = v1 - v0, e2 = v2 - v0)
//
// Arithmetic: the TRIANGLE predicate is part of the parity contract (Moeller-Trumbore,
// fp32, fixed operation order, no culling, open interval); hit/miss is the OR over all triangles and
// does not depend on traversal order. The BOX test only has to be conservative. With boxes stored as
// centre c and half extent h the near/far slab distances are (c - o)/d -+ h/|d|: three FFMAs per axis
// and no per-axis min/max, which moves the work from the ALU pipe (FMNMX, the busiest pipe of the
// traversal loop) to the FMA pipe. Its rounding error is below 1/64 of the padding the builder adds
// to every box (vkr_bvh.cpp), so no triangle the predicate accepts is culled.
#pragma once
#include "vkr_device_math.cuh"

namespace vkr {

struct bvh_view {
	const float4* nodes;
	const float4* tris;
	const uint32_t* tri_ids; // original triangle index per slot (closest-hit only)
	uint32_t tri_count;
};

constexpr int kMaxStackDepth = 64;      // the builder guarantees depth < 62 (vkr_host.cpp)
constexpr int kTraversalDone = 0x7fffffff;

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

// One 32-byte half of a node with a single 256-bit load (LDG.E.256, sm_100): a node pair is two of these instead of three 128-bit and
// one 64-bit load, which halves the wavefronts the L1 data pipe spends per visit -- the unit the trace warps keep busiest.
VKR_DEV void ldg_256(const float4* __restrict__ p, float4& a, float4& b) {
#if defined(__CUDA_ARCH__) && !defined(VKR_NO_LDG256)
	asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
	a = __ldg(p); b = __ldg(p + 1);
#endif
}

// Reciprocal for the slab test only: the box test has to be conservative, not exact (header), so the hardware's approximation does (MUFU.RCP: relative
// error 2^-23, i.e. one more rounding of the size the padding of the boxes is made for; denormal components flush to zero, whose reciprocal is infinite,
// and an infinite or NaN slab distance leaves the slab unconstrained). Saves three IEEE divisions (range check, refinement, slow path) per ray set-up.
VKR_DEV float slab_reciprocal(float x) {
#if defined(__CUDA_ARCH__) && !defined(VKR_EXACT_SLAB_RECIPROCAL)
	float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
	return 1.0f / x;
#endif
}

// Ray in the form the slab test wants: id = 1/d, oid = o/d
struct ray_slabs { f3 id, oid; };
VKR_DEV ray_slabs make_slabs(f3 o, f3 d) {
	ray_slabs r;
	r.id = make3(slab_reciprocal(d.x), slab_reciprocal(d.y), slab_reciprocal(d.z));
	r.oid = make3(o.x * r.id.x, o.y * r.id.y, o.z * r.id.z);
	return r;
}

// Conservative slab test (see header) of the box with centre c and half extent h. Returns the entry distance in
// *t_near. NaNs (inf - inf for axis-parallel rays) drop out of fminf/fmaxf, which leaves that slab unconstrained.
VKR_DEV bool ray_box(float cx, float cy, float cz, float hx, float hy, float hz, const ray_slabs& r, float tmin, float tmax, float* t_near) {
	const float mx = fmaf(cx, r.id.x, -r.oid.x), my = fmaf(cy, r.id.y, -r.oid.y), mz = fmaf(cz, r.id.z, -r.oid.z);
	const float ax = fabsf(r.id.x), ay = fabsf(r.id.y), az = fabsf(r.id.z);
	const float tn = fmaxf(fmaxf(fmaf(-hx, ax, mx), fmaf(-hy, ay, my)), fmaxf(fmaf(-hz, az, mz), tmin));
	const float tf = fminf(fminf(fmaf(hx, ax, mx), fmaf(hy, ay, my)), fminf(fmaf(hz, az, mz), tmax));
	*t_near = tn;
	return tn <= tf;
}

// Per-thread any-hit query (probe kernel vkr_trace_shadow_rays); the shading kernel's own traversal loop lives in
// vkr_ray_stream.cuh (trace warps), built from the same ray_box / ray_triangle.
VKR_DEV bool occluded(const bvh_view& bvh, f3 o, f3 d, float tmin, float tmax, int* stack, int stride) {
	if (!(tmax > tmin)) return false;
	const ray_slabs r = make_slabs(o, d);
	int sp = 0;
	int node = 0;
	float t, tn0, tn1;
	while (true) {
		const float4* n = bvh.nodes + 4 * (size_t) node;
		const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
		const int ref0 = __float_as_int(q3.x), ref1 = __float_as_int(q3.y);
		bool h0 = ray_box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, r, tmin, tmax, &tn0);
		bool h1 = ray_box(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, r, tmin, tmax, &tn1);
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

// One step through a 4-wide node (vkr_bvh.h: host_bvh4; 8 x float4 per node: child c has centre and half extent at floats [6c, 6c + 6), the four
// references as int bits in the seventh float4). Tests the four boxes, returns the reference of the nearest hit child and hands the other hit
// children (inner nodes and leaves alike) to push(); returns kTraversalDone if no child is hit. Shared by occluded4() below (tested on the CPU)
// and by the experimental 4-wide form of the trace warps' loop (vkr_ray_stream.cuh, VKR_BVH_WIDTH == 4).
template <class Push>
VKR_DEV int bvh4_descend_step(const float4* __restrict__ nodes4, int node, const ray_slabs& r, float tmin, float tmax, Push&& push) {
	const float4* nd = nodes4 + 8 * (size_t) node;
	float4 q0, q1, q2, q3, q4, q5, q6, q7;
	ldg_256(nd, q0, q1); ldg_256(nd + 2, q2, q3); ldg_256(nd + 4, q4, q5); ldg_256(nd + 6, q6, q7);
	const int ref0 = __float_as_int(q6.x), ref1 = __float_as_int(q6.y), ref2 = __float_as_int(q6.z), ref3 = __float_as_int(q6.w);
	float tn0, tn1, tn2, tn3;
	const bool h0 = ray_box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, r, tmin, tmax, &tn0);
	const bool h1 = ray_box(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, r, tmin, tmax, &tn1);
	const bool h2 = ray_box(q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, r, tmin, tmax, &tn2);
	const bool h3 = ray_box(q4.z, q4.w, q5.x, q5.y, q5.z, q5.w, r, tmin, tmax, &tn3);
	int best = -1; float best_t = 0.0f;
	if (h0) { best = 0; best_t = tn0; }
	if (h1 && (best < 0 || tn1 < best_t)) { best = 1; best_t = tn1; }
	if (h2 && (best < 0 || tn2 < best_t)) { best = 2; best_t = tn2; }
	if (h3 && (best < 0 || tn3 < best_t)) { best = 3; best_t = tn3; }
	if (h0 && best != 0) push(ref0);
	if (h1 && best != 1) push(ref1);
	if (h2 && best != 2) push(ref2);
	if (h3 && best != 3) push(ref3);
	return (best < 0) ? kTraversalDone : ((best == 0) ? ref0 : ((best == 1) ? ref1 : ((best == 2) ? ref2 : ref3)));
}

// Any-hit query over 4-wide nodes. Same predicate, same answer as occluded(); a step decides four children at once, so a ray takes about half as
// many steps. `steps` counts the nodes fetched (statistics for the tests, may be null).
VKR_DEV bool occluded4(const float4* __restrict__ nodes4, const float4* __restrict__ tris, f3 o, f3 d, float tmin, float tmax, int* stack, int stride, int* steps) {
	if (!(tmax > tmin)) return false;
	const ray_slabs r = make_slabs(o, d);
	int sp = 0;
	int node = 0;
	float t;
	auto push = [&](int ref) { stack[sp * stride] = ref; ++sp; };
	while (true) {
		if (node < 0) { // a leaf
			const int first = (node & 0x7fffffff) >> 4, count = node & 15;
			for (int i = 0; i != count; ++i)
				if (ray_triangle(tris + 3 * (size_t) (first + i), o, d, tmin, tmax, &t)) return true;
			node = kTraversalDone;
		}
		else {
			if (steps) ++*steps;
			node = bvh4_descend_step(nodes4, node, r, tmin, tmax, push);
		}
		if (node == kTraversalDone) {
			if (sp == 0) return false;
			--sp; node = stack[sp * stride];
		}
	}
}

// Closest-hit query, ties in t resolve to the lowest original triangle index (order independent).
VKR_DEV int closest_hit(const bvh_view& bvh, f3 o, f3 d, float tmin, float tmax, int* stack, int stride) {
	const ray_slabs r = make_slabs(o, d);
	int sp = 0;
	int node = 0;
	int best = -1;
	float best_t = tmax;
	float t, tn0, tn1;
	while (true) {
		const float4* n = bvh.nodes + 4 * (size_t) node;
		const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
		const int ref0 = __float_as_int(q3.x), ref1 = __float_as_int(q3.y);
		// closed upper bound so that equal-t candidates are still visited for the tie rule
		bool h0 = ray_box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, r, tmin, best_t, &tn0);
		bool h1 = ray_box(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, r, tmin, best_t, &tn1);
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
		if (h0 && h1) {
			const bool swap = tn1 < tn0;
			stack[sp * stride] = swap ? ref0 : ref1; ++sp; node = swap ? ref1 : ref0;
		}
		else if (h0) node = ref0;
		else if (h1) node = ref1;
		else {
			if (sp == 0) return best;
			--sp; node = stack[sp * stride];
		}
	}
}

} // namespace vkr
