// Lock-step emulation of the trace warps' loop (vkr_ray_stream.cuh) on the CPU: 32 lanes, the same per-lane state machine, phases costed in warp instructions.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <vector>
#define VKR_DEVICE_CODE_ON_HOST 1
#define VKR_DEV inline
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
struct float4 { float x, y, z, w; };
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned) x) : 32; }
#include "vkr_anchor.cuh"
using namespace vkr;

struct params { int min_lanes, refill_min, leaf_once, anchored, pooled_tris; };
struct ray { uint32_t pixel; f3 d; float tmax; uint32_t mask; };
struct lane_state {
	bool active = false, hit = false; int node = kTraversalDone, leaf = 0, cached = -1; uint32_t pending = 0; int ray_index = -1;
	f3 o, d; float tmax = 0; ray_slabs r; std::vector<int> stack; const uint32_t* path = nullptr;
};
// instruction costs per phase (warp instructions when at least one lane is in the phase), from the SASS of shading_kernel<3,5,0,0,1>
static const int C_ROUND = 14, C_TICKET = 22, C_SETUP_FETCH = 28, C_CACHE_TEST = 62, C_SLABS = 44, C_ANCHOR_SETUP = 10, C_NODE = 58, C_SIBLING_EXTRA = 8, C_TRI = 58, C_LEAF_OVERHEAD = 8, C_FINISH = 6;

extern "C" void simulate(const float* nodes, const float* tris, const float* origins, uint32_t pixel_count, const uint32_t* paths /*[pixel][kPathLevels+2]*/, uint32_t ray_count, const float* rays /*pixel, dx,dy,dz,tmax, mask(bits as float)*/,
	const int* prm, double* out)
{
	const params P = { prm[0], prm[1], prm[2], prm[3], prm[4] };
	const float4* N = (const float4*) nodes; const float4* T = (const float4*) tris;
	const float tmin = 1e-3f;
	double warp_instr = 0, node_iters = 0, node_lane_visits = 0, leaf_iters = 0, leaf_lane_tests = 0, setup_rounds = 0, setup_lanes = 0, rounds = 0, hits = 0, lane_instr = 0;
	lane_state L[32];
	uint32_t next = 0;
	while (true) {
		++rounds; warp_instr += C_ROUND;
		// refill
		int busy = 0; for (int l = 0; l != 32; ++l) busy += L[l].active;
		const bool refill = busy == 0 || (32 - busy) >= P.refill_min;
		int starters = 0, cache_testers = 0, slab_makers = 0;
		if (refill && next < ray_count) {
			for (int l = 0; l != 32 && next < ray_count; ++l) if (!L[l].active) {
				lane_state& s = L[l];
				const float* r = rays + 6 * (size_t) next;
				s.ray_index = (int) next++;
				const uint32_t p = (uint32_t) r[0];
				s.o = make3(origins[3 * p], origins[3 * p + 1], origins[3 * p + 2]); s.d = make3(r[1], r[2], r[3]); s.tmax = r[4];
				s.active = true; s.hit = false; s.node = kTraversalDone; s.leaf = 0; s.pending = 0; s.stack.clear();
				++starters;
				if (s.tmax > tmin) {
					float t; bool cached_hit = false;
					if (s.cached >= 0) { ++cache_testers; cached_hit = ray_triangle(T + 3 * (size_t) s.cached, s.o, s.d, tmin, s.tmax, &t); }
					if (cached_hit) s.hit = true;
					else {
						++slab_makers;
						s.r = make_slabs(s.o, s.d);
						if (P.anchored) {
							const uint32_t* pp = paths + (size_t) p * (kPathLevels + 2);
							s.path = pp; const uint32_t count = pp[kPathLevels + 1];
							uint32_t mask; memcpy(&mask, &r[5], 4);
							s.pending = mask & ((1u << count) - 1u);
							s.node = (int) pp[kPathLevels];
							if (s.node < 0) { s.leaf = s.node; s.node = kTraversalDone; }
						}
						else s.node = 0;
					}
				}
			}
			if (starters) { warp_instr += C_TICKET + C_SETUP_FETCH; ++setup_rounds; setup_lanes += starters; lane_instr += starters * (C_TICKET + C_SETUP_FETCH); }
			if (cache_testers) { warp_instr += C_CACHE_TEST; lane_instr += cache_testers * C_CACHE_TEST; }
			if (slab_makers) { warp_instr += C_SLABS + (P.anchored ? C_ANCHOR_SETUP : 0); lane_instr += slab_makers * C_SLABS; }
		}
		busy = 0; for (int l = 0; l != 32; ++l) busy += L[l].active;
		if (!busy) { if (next >= ray_count) break; continue; }
		// node loop
		while (true) {
			int descending = 0; bool any_sibling = false;
			for (int l = 0; l != 32; ++l) { lane_state& s = L[l]; if (s.active && s.node >= 0 && (s.node != kTraversalDone || (P.anchored && s.pending && !s.hit))) ++descending; }
			if (descending == 0) break;
			for (int l = 0; l != 32; ++l) {
				lane_state& s = L[l];
				if (!(s.active && s.node >= 0 && (s.node != kTraversalDone || (P.anchored && s.pending && !s.hit)))) continue;
				int skip = 2;
				if (s.node == kTraversalDone) { const int k = 31 - __clz((int) s.pending); s.pending &= ~(1u << k); const uint32_t e = s.path[k]; s.node = (int) (e >> 1); skip = (int) (e & 1u); s.stack.clear(); any_sibling = true; }
				auto push = [&](int ref) { s.stack.push_back(ref); };
				auto pop = [&]() { if (s.stack.empty()) return (int) kTraversalDone; const int v = s.stack.back(); s.stack.pop_back(); return v; };
				s.node = visit_pair(N, s.node, skip, s.r, tmin, s.tmax, push);
				if (s.node == kTraversalDone) s.node = pop();
				if (s.node < 0 && s.leaf == 0) { s.leaf = s.node; s.node = pop(); }
			}
			++node_iters; node_lane_visits += descending; warp_instr += C_NODE + (any_sibling ? C_SIBLING_EXTRA : 0); lane_instr += descending * C_NODE;
			int still = 0;
			for (int l = 0; l != 32; ++l) { lane_state& s = L[l]; if (s.active && s.node >= 0 && (s.node != kTraversalDone || (P.anchored && s.pending && !s.hit))) ++still; }
			if (still < P.min_lanes) break;
		}
		// leaf phase
		while (true) {
			int with_leaf = 0, max_count = 0;
			for (int l = 0; l != 32; ++l) if (L[l].active && L[l].leaf != 0) { ++with_leaf; const int c = L[l].leaf & 15; if (c > max_count) max_count = c; }
			if (!with_leaf) break;
			if (P.pooled_tris) { // all (ray, triangle) pairs of the round go into one list that the warp works off 32 at a time
				int total = 0; for (int l = 0; l != 32; ++l) if (L[l].active && L[l].leaf != 0) total += L[l].leaf & 15;
				const int chunks = (total + 31) / 32;
				warp_instr += 24 + chunks * (C_TRI + 10); leaf_iters += chunks; leaf_lane_tests += total; lane_instr += total * C_TRI;
			}
			else {
				for (int i = 0; i != max_count; ++i) { int testers = 0; for (int l = 0; l != 32; ++l) if (L[l].active && L[l].leaf != 0 && (L[l].leaf & 15) > i) ++testers; warp_instr += C_TRI; ++leaf_iters; leaf_lane_tests += testers; lane_instr += testers * C_TRI; }
				warp_instr += C_LEAF_OVERHEAD;
			}
			for (int l = 0; l != 32; ++l) {
				lane_state& s = L[l]; if (!(s.active && s.leaf != 0)) continue;
				const int first = (s.leaf & 0x7fffffff) >> 4, c = s.leaf & 15; float t;
				for (int i = 0; i != c; ++i) if (ray_triangle(T + 3 * (size_t) (first + i), s.o, s.d, tmin, s.tmax, &t)) { s.hit = true; s.cached = first + i; }
				s.leaf = 0;
				if (s.hit) s.node = kTraversalDone;
				else if (s.node < 0) { s.leaf = s.node; if (s.stack.empty()) s.node = kTraversalDone; else { s.node = s.stack.back(); s.stack.pop_back(); } }
			}
			if (P.leaf_once) break;
		}
		// finish
		int finishing = 0;
		for (int l = 0; l != 32; ++l) { lane_state& s = L[l]; if (s.active && s.node == kTraversalDone && s.leaf == 0 && (s.hit || !P.anchored || s.pending == 0)) { s.active = false; hits += s.hit; ++finishing; } }
		if (finishing) warp_instr += C_FINISH;
	}
	out[0] = warp_instr / ray_count * 32; out[1] = node_lane_visits / node_iters; out[2] = leaf_lane_tests / leaf_iters; out[3] = setup_lanes / setup_rounds; out[4] = node_lane_visits / ray_count; out[5] = leaf_lane_tests / ray_count;
	out[6] = hits / ray_count; out[7] = lane_instr / ray_count; out[8] = rounds / ray_count * 32;
}

// helper: paths and masks for pixels / rays (mask per ray needs the light: rays carry light index in slot 5 on input, replaced by the mask)
extern "C" void prepare(const float* nodes, const float* origins, uint32_t pixel_count, const float* light_vertices, uint32_t ray_count, float* rays, uint32_t* paths) {
	const float4* N = (const float4*) nodes;
	for (uint32_t p = 0; p != pixel_count; ++p) {
		const f3 o = make3(origins[3 * p], origins[3 * p + 1], origins[3 * p + 2]);
		uint32_t* pp = paths + (size_t) p * (kPathLevels + 2); int tail = 0;
		const int count = find_origin_path(N, o, &tail, [&](int k, uint32_t e) { pp[k] = e; });
		pp[kPathLevels] = (uint32_t) tail; pp[kPathLevels + 1] = (uint32_t) count;
	}
	for (uint32_t i = 0; i != ray_count; ++i) {
		float* r = rays + 6 * (size_t) i; const uint32_t p = (uint32_t) r[0], l = (uint32_t) r[5];
		const f3 o = make3(origins[3 * p], origins[3 * p + 1], origins[3 * p + 2]);
		const uint32_t* pp = paths + (size_t) p * (kPathLevels + 2);
		const light_cone cone = make_light_cone(o, (const unsigned char*) (light_vertices + 16 * l), 4);
		uint32_t mask = cull_siblings(N, o, cone, (int) pp[kPathLevels + 1], [&](int k) { return pp[k]; });
		if (!ray_in_cone(cone, make3(r[1], r[2], r[3]), r[4])) mask = kAllSiblings;
		memcpy(&r[5], &mask, 4);
	}
}
