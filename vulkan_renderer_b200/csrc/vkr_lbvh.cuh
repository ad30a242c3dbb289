// vkr_lbvh.cuh -- the per-element steps of the GPU linear-BVH builder (vkr_lbvh_gpu.cu wraps each into a kernel).
//
// One function per step, one call per triangle / sorted position / internal node, no warp-level operations: the same source runs on the
// CPU (tests/device_on_host.cpp emulates the launches one element at a time) and must produce the arrays of the sequential reference
// vkr_lbvh.cpp byte for byte. Steps and their definitions are described there and in vkr_lbvh_gpu.cu.
#pragma once
#include "vkr_device_math.cuh"

namespace vkr {

constexpr int kLbvhLeafSize = 4;

// float <-> unsigned key with the same order, for atomicMin / atomicMax over bounds
VKR_DEV uint32_t float_to_ordered(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
VKR_DEV float ordered_to_float(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// 1. box and centroid of triangle t
VKR_DEV void lbvh_triangle_bounds(const float* __restrict__ vertices, uint32_t t, float* lo3, float* hi3, float* centroid3) {
	const float* v = vertices + 9 * (size_t) t;
#pragma unroll
	for (int a = 0; a != 3; ++a) {
		const float v0 = v[a], v1 = v[3 + a], v2 = v[6 + a];
		lo3[a] = fminf(v0, fminf(v1, v2)); hi3[a] = fmaxf(v0, fmaxf(v1, v2));
		centroid3[a] = 0.5f * (lo3[a] + hi3[a]);
	}
}

VKR_DEV uint64_t lbvh_expand_bits_21(uint32_t v) { // ...abc -> ..a00b00c
	uint64_t x = v & 0x1fffffu;
	x = (x | x << 32) & 0x1f00000000ffffull;
	x = (x | x << 16) & 0x1f0000ff0000ffull;
	x = (x | x << 8) & 0x100f00f00f00f00full;
	x = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}

// 2. 63-bit Morton code of a centroid inside the centroid bounds
VKR_DEV uint64_t lbvh_morton(const float* centroid3, f3 lo, f3 inv_extent) {
	const float l[3] = { lo.x, lo.y, lo.z }, ie[3] = { inv_extent.x, inv_extent.y, inv_extent.z };
	uint32_t q[3];
#pragma unroll
	for (int a = 0; a != 3; ++a) {
		const float f = ((centroid3[a] - l[a]) * ie[a]) * 2097152.0f;
		q[a] = (f > 0.0f) ? ((f < 2097151.0f) ? (uint32_t) f : 2097151u) : 0u;
	}
	return lbvh_expand_bits_21(q[0]) << 2 | lbvh_expand_bits_21(q[1]) << 1 | lbvh_expand_bits_21(q[2]);
}

// 3. triangle slot s (v0, e1, e2 as 12 floats) of the triangle at sorted position s
VKR_DEV void lbvh_slot(const float* __restrict__ vertices, uint32_t t, float* slot12) {
	const float* v = vertices + 9 * (size_t) t;
	slot12[0] = v[0]; slot12[1] = v[1]; slot12[2] = v[2];
	slot12[3] = v[3] - v[0]; slot12[4] = v[4] - v[1]; slot12[5] = v[5] - v[2];
	slot12[6] = v[6] - v[0]; slot12[7] = v[7] - v[1]; slot12[8] = v[8] - v[2];
	slot12[9] = slot12[10] = slot12[11] = 0.0f;
}

// Length of the common prefix of the keys at sorted positions i and j; the position breaks ties between equal codes
VKR_DEV int lbvh_delta(const uint64_t* __restrict__ codes, int64_t n, int64_t i, int64_t j) {
	if (j < 0 || j >= n) return -1;
	const uint64_t x = codes[i] ^ codes[j];
	return x ? __clzll((long long) x) : 64 + __clz((int) ((uint32_t) i ^ (uint32_t) j));
}

// 4. internal node i of the binary radix tree (Karras 2012): its range of sorted positions, its split, the parent links of its children
VKR_DEV void lbvh_radix_tree_node(const uint64_t* __restrict__ codes, int64_t count, int64_t i, int32_t* first, int32_t* last, int32_t* split, int32_t* parent, int32_t* leaf_parent) {
	const int d = (lbvh_delta(codes, count, i, i + 1) - lbvh_delta(codes, count, i, i - 1)) >= 0 ? 1 : -1;
	const int delta_min = lbvh_delta(codes, count, i, i - d);
	int64_t l_max = 2;
	while (lbvh_delta(codes, count, i, i + l_max * d) > delta_min) l_max *= 2;
	int64_t l = 0;
	for (int64_t t = l_max / 2; t >= 1; t /= 2)
		if (lbvh_delta(codes, count, i, i + (l + t) * d) > delta_min) l += t;
	const int64_t j = i + l * d;
	const int delta_node = lbvh_delta(codes, count, i, j);
	int64_t s = 0;
	for (int64_t t = (l + 1) / 2; ; t = (t + 1) / 2) {
		if (lbvh_delta(codes, count, i, i + (s + t) * d) > delta_node) s += t;
		if (t == 1) break;
	}
	const int64_t gamma = i + s * d + (d < 0 ? d : 0);
	const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
	first[i] = (int32_t) lo; last[i] = (int32_t) hi; split[i] = (int32_t) gamma;
	if (gamma == lo) leaf_parent[gamma] = (int32_t) i; else parent[gamma] = (int32_t) i;
	if (gamma + 1 == hi) leaf_parent[gamma + 1] = (int32_t) i; else parent[gamma + 1] = (int32_t) i;
}

// 5. boxes bottom-up, started from sorted position s: the first thread to arrive at a node leaves, the second one merges its children's boxes
VKR_DEV void lbvh_refit_from_leaf(uint32_t s, const uint32_t* __restrict__ order, const float* __restrict__ box_lo, const float* __restrict__ box_hi,
	const int32_t* __restrict__ first, const int32_t* __restrict__ last, const int32_t* __restrict__ split, const int32_t* __restrict__ parent, const int32_t* __restrict__ leaf_parent,
	float* node_lo, float* node_hi, uint32_t* arrivals)
{
	int32_t node = leaf_parent[s];
	while (node >= 0) {
		__threadfence(); // the boxes this thread has written below `node` are visible before its arrival is counted
		if (atomicAdd(&arrivals[node], 1u) == 0u) return; // the sibling subtree is not finished: its last thread continues
		__threadfence();
		const int32_t g = split[node];
		const volatile float* l_lo; const volatile float* l_hi; const volatile float* r_lo; const volatile float* r_hi;
		if (g == first[node]) { l_lo = box_lo + 3 * (size_t) order[g]; l_hi = box_hi + 3 * (size_t) order[g]; }
		else { l_lo = node_lo + 3 * (size_t) g; l_hi = node_hi + 3 * (size_t) g; }
		if (g + 1 == last[node]) { r_lo = box_lo + 3 * (size_t) order[g + 1]; r_hi = box_hi + 3 * (size_t) order[g + 1]; }
		else { r_lo = node_lo + 3 * (size_t) (g + 1); r_hi = node_hi + 3 * (size_t) (g + 1); }
#pragma unroll
		for (int a = 0; a != 3; ++a) {
			node_lo[3 * (size_t) node + a] = fminf(l_lo[a], r_lo[a]);
			node_hi[3 * (size_t) node + a] = fmaxf(l_hi[a], r_hi[a]);
		}
		node = parent[node];
	}
}

// 6. an internal node with more than four sorted positions keeps its node pair; the others collapse into leaves
VKR_DEV uint32_t lbvh_is_used(int32_t first, int32_t last) { return (last - first + 1 > kLbvhLeafSize) ? 1u : 0u; }

// Centre and half extent of a padded box, the half extent rounded up: the same double-precision steps as write_child() in vkr_lbvh.cpp
VKR_DEV void lbvh_encode_box(float* dst6, const float* lo3, const float* hi3, float pad) {
#pragma unroll
	for (int a = 0; a != 3; ++a) {
		const double lo = (double) lo3[a] - (double) pad, hi = (double) hi3[a] + (double) pad;
		const float ctr = (float) (0.5 * (lo + hi));
		const double up = (double) ctr - lo, down = hi - (double) ctr;
		dst6[a] = ctr;
		dst6[3 + a] = nextafterf((float) (up > down ? up : down), INFINITY);
	}
}

// 7. the node pair of used internal node i (16 floats at rank[i])
VKR_DEV void lbvh_emit_pair(uint32_t i, const uint32_t* __restrict__ order, const float* __restrict__ box_lo, const float* __restrict__ box_hi,
	const float* __restrict__ node_lo, const float* __restrict__ node_hi, const int32_t* __restrict__ first, const int32_t* __restrict__ last, const int32_t* __restrict__ split,
	const uint32_t* __restrict__ used, const uint32_t* __restrict__ rank, float pad, float* out16)
{
#pragma unroll
	for (int k = 0; k != 16; ++k) out16[k] = 0.0f;
	const int32_t g = split[i];
#pragma unroll
	for (int c = 0; c != 2; ++c) {
		const int32_t lo = c ? g + 1 : first[i], hi = c ? last[i] : g;
		int32_t ref;
		if (lo == hi) {
			lbvh_encode_box(out16 + 6 * c, box_lo + 3 * (size_t) order[lo], box_hi + 3 * (size_t) order[lo], pad);
			ref = (int32_t) (0x80000000u | ((uint32_t) lo << 4) | 1u);
		}
		else {
			const int32_t child = c ? g + 1 : g;
			lbvh_encode_box(out16 + 6 * c, node_lo + 3 * (size_t) child, node_hi + 3 * (size_t) child, pad);
			ref = used[child] ? (int32_t) rank[child] : (int32_t) (0x80000000u | ((uint32_t) lo << 4) | (uint32_t) (hi - lo + 1));
		}
		out16[12 + c] = __int_as_float(ref);
	}
}

// 8. depth of used internal node i in node pairs
VKR_DEV uint32_t lbvh_depth(uint32_t i, const int32_t* __restrict__ parent) {
	uint32_t depth = 1;
	for (int32_t p = parent[i]; p >= 0; p = parent[p]) ++depth;
	return depth;
}

} // namespace vkr
