// vkr_ray_queue.cuh -- per-warp shadow-ray queue of the shading megakernel.
//
// Rays are not traced where they are generated. Every lane pushes its candidate samples (ray
// direction, distance to the light plane, the radiance to add if the ray is unoccluded) into a ring
// buffer in shared memory owned by its warp, compacted with __ballot_sync. When the ring is nearly
// full (or a light is finished) the warp drains it: all 32 lanes traverse the BVH with a
// warp-synchronous shared-memory stack and REFILL themselves from the queue as soon as their ray
// terminates, so traversal runs at full warp width although the rays come from lanes that may be
// idle (background pixels, samples below the horizon, lights behind the surface) and although ray
// lengths differ. Afterwards every lane adds the resolved contributions of its own pixel strictly in
// submission order, which keeps the floating-point sums identical to the reference's sequential loop
// (shading_pass.frag.glsl:608-637).
#pragma once
#include "vkr_trace.cuh"

namespace vkr {

constexpr int kQueueCapacity = 256;        // rays per warp (power of two)
constexpr int kQueueDrainThreshold = kQueueCapacity - 64;  // a sample adds at most 2 x 32 rays
constexpr unsigned kFullMask = 0xffffffffu;

struct ray_queue {
	float* dx; float* dy; float* dz; float* tmax;   // [kQueueCapacity] ray direction (world), far end = light plane
	float* cx; float* cy; float* cz;                 // contribution if the ray is unoccluded
	float* ox_; float* oy_; float* oz_;              // contribution if it is occluded (MIS_HEURISTIC_OPTIMAL only)
	unsigned char* owner;                            // lane of the owning pixel; bit 7: known to be occluded (n.w <= 0)
	unsigned char* occluded;                         // result per entry
	const float* origin;                             // [3 * 32] ray origins = shading positions of the warp's lanes
	int* stack; int stack_stride;                    // this lane's column of the traversal stack
	bvh_view bvh;
	int count;                                       // warp-uniform; entries 0..count-1 are pending, oldest first
	int cached_triangle;                             // per lane: slot of the last triangle that occluded a ray (-1: none)
	bool enabled;                                    // TRACE_SHADOW_RAYS
};

// Floats of shared memory per warp: 7 (or 10) float arrays, owner + result bytes, 96 floats of origins
VKR_DEV constexpr size_t queue_floats_per_warp(bool optimal) { return (optimal ? 10 : 7) * kQueueCapacity + 2 * kQueueCapacity / 4 + 96; }

VKR_DEV void queue_bind(ray_queue& q, float* base, bool optimal) {
	q.dx = base; q.dy = base + kQueueCapacity; q.dz = base + 2 * kQueueCapacity; q.tmax = base + 3 * kQueueCapacity;
	q.cx = base + 4 * kQueueCapacity; q.cy = base + 5 * kQueueCapacity; q.cz = base + 6 * kQueueCapacity;
	float* rest = base + 7 * kQueueCapacity;
	q.ox_ = q.oy_ = q.oz_ = nullptr;
	if (optimal) { q.ox_ = rest; q.oy_ = rest + kQueueCapacity; q.oz_ = rest + 2 * kQueueCapacity; rest += 3 * kQueueCapacity; }
	q.owner = reinterpret_cast<unsigned char*>(rest);
	q.occluded = q.owner + kQueueCapacity;
	q.origin = rest + 2 * kQueueCapacity / 4;
	q.count = 0;
	q.cached_triangle = -1;
}

// Traces all pending rays (dynamic refill) and adds the resolved contributions to their owners, oldest first.
template <bool OPTIMAL>
VKR_DEV void drain(ray_queue& q, int lane, f3& result) {
	const int n = q.count;
	if (n == 0) return;
	const unsigned lt_mask = (1u << lane) - 1u;
	const float tmin = 1.0e-3f; // shading_pass.frag.glsl:124
	int next = 0;               // warp-uniform: first entry nobody has taken yet
	int entry = 0;
	int node = kTraversalDone, leaf = 0;
	// The stack is addressed through ONE loop-carried register (shared-memory address of the next free slot, 128 B
	// between levels = one slot per lane) with a kTraversalDone sentinel at the bottom, so a pop never needs an
	// "empty" test and the compiler cannot rematerialise base + sp * stride around every push and pop.
	const uint32_t stack_bottom = (uint32_t) __cvta_generic_to_shared(q.stack);
	uint32_t top = stack_bottom;
	const uint32_t level = (uint32_t) q.stack_stride * 4u;
	auto push = [&](int v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(top), "r"(v) : "memory"); top += level; };
	auto pop = [&]() { int v; top -= level; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(top) : "memory"); return v; };
	f3 o = make3(0.0f, 0.0f, 0.0f), d = make3(0.0f, 0.0f, 1.0f);
	float tmax = 0.0f;
	ray_slabs r = make_slabs(o, d);
	while (true) {
		// --- lanes whose ray has terminated take the next pending rays
		const bool idle = node == kTraversalDone && leaf == 0;
		const unsigned idle_mask = __ballot_sync(kFullMask, idle);
		if (next < n) {
			const int mine = next + __popc(idle_mask & lt_mask);
			if (idle && mine < n) {
				entry = mine;
				const int own = q.owner[entry];
				const int ol = own & 31;
				o = make3(q.origin[ol], q.origin[32 + ol], q.origin[64 + ol]);
				d = make3(q.dx[entry], q.dy[entry], q.dz[entry]);
				tmax = q.tmax[entry];
				bool occ = (own & 128) != 0;
				bool go = !occ && tmax > tmin; // tmax <= tmin / NaN: undefined in Vulkan, defined as "miss" 
				float t;
				if (go && q.cached_triangle >= 0 && ray_triangle(q.bvh.tris + 3 * (size_t) q.cached_triangle, o, d, tmin, tmax, &t)) { occ = true; go = false; }
				q.occluded[entry] = occ ? 1 : 0;
				if (go) { r = make_slabs(o, d); node = 0; top = stack_bottom; push(kTraversalDone); }
			}
			next = min(n, next + __popc(idle_mask));
		}
		else if (idle_mask == kFullMask) break;
		// --- descend until this lane holds two leaves or is out of nodes
		while (node >= 0 && node != kTraversalDone) {
			const float4* nd = q.bvh.nodes + 4 * (size_t) node;
			const float4 q0 = __ldg(nd), q1 = __ldg(nd + 1), q2 = __ldg(nd + 2), q3 = __ldg(nd + 3);
			const int ref0 = __float_as_int(q3.x), ref1 = __float_as_int(q3.y);
			float tn0, tn1;
			const bool h0 = ray_box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, r, tmin, tmax, &tn0);
			const bool h1 = ray_box(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, r, tmin, tmax, &tn1);
			if (h0 && h1) {
				const bool swap = tn1 < tn0;   // nearer child first: occluders close to the surface end the query early
				node = swap ? ref1 : ref0;
				push(swap ? ref0 : ref1);
			}
			else if (h0) node = ref0;
			else if (h1) node = ref1;
			else node = pop();
			if (node < 0 && leaf == 0) { // postpone the first leaf, keep descending
				leaf = node;
				node = pop();
			}
		}
		__syncwarp(kFullMask);
		// --- leaves: `leaf` and possibly `node` (a second leaf)
		while (leaf != 0) {
			const int first = (leaf & 0x7fffffff) >> 4, count = leaf & 15;
			bool hit = false;
			float t;
			for (int i = 0; i != count; ++i)
				if (ray_triangle(q.bvh.tris + 3 * (size_t) (first + i), o, d, tmin, tmax, &t)) { hit = true; q.cached_triangle = first + i; }
			leaf = 0;
			if (hit) { q.occluded[entry] = 1; node = kTraversalDone; }
			else if (node < 0) {
				leaf = node;
				node = pop();
			}
		}
		__syncwarp(kFullMask);
	}
	// --- every lane adds the contributions of its own pixel, oldest first, 32 entries at a time
	for (int base = 0; base < n; base += 32) {
		const int e = base + lane;
		const int own = (e < n) ? (q.owner[e] & 31) : 32;
		unsigned mine = __ballot_sync(kFullMask, e < n);
#pragma unroll
		for (int b = 0; b != 5; ++b) {
			const unsigned bits = __ballot_sync(kFullMask, (own >> b) & 1);
			mine &= ((lane >> b) & 1) ? bits : ~bits;
		}
		while (mine) {
			const int s = base + __ffs(mine) - 1;
			mine &= mine - 1;
			if (!q.occluded[s]) result = result + make3(q.cx[s], q.cy[s], q.cz[s]);
			else if (OPTIMAL) result = result + make3(q.ox_[s], q.oy_[s], q.oz_[s]);
		}
	}
	q.count = 0;
	__syncwarp(kFullMask);
}

// Warp-convergent: every lane calls it once per candidate sample. has = this lane contributes something.
// need_trace = visibility is not known yet (n.w > 0); otherwise the sample is known to be occluded.
// finish (warp-uniform) = drain even if the queue is not full (end of a light). Keeping the only call of drain() here
// gives each kernel ONE copy of the traversal loop (instruction-cache footprint).
template <bool OPTIMAL>
VKR_DEV void submit(ray_queue& q, int lane, bool has, bool need_trace, f3 dir_world, float tmax, f3 c_visible, f3 c_occluded, f3& result, bool finish) {
	if (!q.enabled) { // no shadow rays: visibility = (n.w > 0), nothing is ever pending, add in place
		if (has) {
			if (need_trace) result = result + c_visible;
			else if (OPTIMAL) result = result + c_occluded;
		}
		return;
	}
	const bool push = has && (need_trace || OPTIMAL);
	const unsigned mask = __ballot_sync(kFullMask, push);
	if (push) {
		const int s = q.count + __popc(mask & ((1u << lane) - 1u));
		q.dx[s] = dir_world.x; q.dy[s] = dir_world.y; q.dz[s] = dir_world.z; q.tmax[s] = tmax;
		q.cx[s] = c_visible.x; q.cy[s] = c_visible.y; q.cz[s] = c_visible.z;
		if (OPTIMAL) { q.ox_[s] = c_occluded.x; q.oy_[s] = c_occluded.y; q.oz_[s] = c_occluded.z; }
		q.owner[s] = (unsigned char) (need_trace ? lane : (lane | 128));
	}
	q.count += __popc(mask);
	__syncwarp(kFullMask);
	if (q.count > kQueueDrainThreshold || (finish && q.count > 0)) drain<OPTIMAL>(q, lane, result);
}

} // namespace vkr
