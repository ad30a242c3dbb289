// vkr_lbvh.cpp -- linear BVH on the host: the reference of the GPU builder.
//
// The reference renderer hands its triangle soup to the driver (vkCmdBuildAccelerationStructuresKHR, src/scene.c:354-378),
// which builds on the GPU in milliseconds; the binned-SAH builder of vkr_bvh.cpp takes 1.3-1.8 s for 2.8 M triangles on the
// host. This file holds the fast alternative in its sequential form: Morton order of the centroids, the binary radix tree of
// Karras ("Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees", HPG 2012), subtrees of at most four
// triangles collapsed into leaves, output in the node-pair layout of vkr_trace.cuh. Every step is defined so that a parallel
// implementation gives the same bytes (exact min/max, unique sort keys, ranks by prefix sum): vkr_lbvh_gpu.cu is that
// implementation and is tested against this one array for array. Tree quality is below the SAH builder's (more node visits
// per ray), results are identical: the shadow predicate is an OR over all triangles.
//   select with VKR_BVH_BUILDER=lbvh (this file) or lbvh_gpu; the default stays "sah".
#include "vkr_bvh.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace vkr {

uint64_t lbvh_expand_bits_21(uint32_t v) { // ...abc -> ..a00b00c (21 bits to 61)
	uint64_t x = v & 0x1fffffu;
	x = (x | x << 32) & 0x1f00000000ffffull;
	x = (x | x << 16) & 0x1f0000ff0000ffull;
	x = (x | x << 8) & 0x100f00f00f00f00full;
	x = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}

uint64_t lbvh_morton_code(const float centroid[3], const float lo[3], const float inv_extent[3]) {
	uint32_t q[3];
	for (int a = 0; a != 3; ++a) {
		const float f = ((centroid[a] - lo[a]) * inv_extent[a]) * 2097152.0f;
		q[a] = (f > 0.0f) ? ((f < 2097151.0f) ? (uint32_t) f : 2097151u) : 0u;
	}
	return lbvh_expand_bits_21(q[0]) << 2 | lbvh_expand_bits_21(q[1]) << 1 | lbvh_expand_bits_21(q[2]);
}

namespace {

struct box3 { float lo[3], hi[3]; };

inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
inline float as_float(int32_t bits) { float f; std::memcpy(&f, &bits, 4); return f; }

// Length of the common prefix of the keys at sorted positions i and j (the position breaks ties between equal codes)
inline int delta(const std::vector<uint64_t>& codes, int64_t i, int64_t j) {
	if (j < 0 || j >= (int64_t) codes.size()) return -1;
	const uint64_t x = codes[i] ^ codes[j];
	return x ? clz64(x) : 64 + clz32((uint32_t) i ^ (uint32_t) j);
}

} // namespace

void build_lbvh(host_bvh& out, const float* vertices, uint64_t triangle_count) {
	out = host_bvh();
	const int64_t n = (int64_t) triangle_count;
	std::vector<box3> tri_box((size_t) n);
	std::vector<float> centroid(3 * (size_t) n);
	box3 scene, cbox;
	for (int a = 0; a != 3; ++a) { scene.lo[a] = cbox.lo[a] = std::numeric_limits<float>::infinity(); scene.hi[a] = cbox.hi[a] = -std::numeric_limits<float>::infinity(); }
	for (int64_t t = 0; t != n; ++t) {
		box3& b = tri_box[(size_t) t];
		for (int a = 0; a != 3; ++a) {
			const float v0 = vertices[9 * t + a], v1 = vertices[9 * t + 3 + a], v2 = vertices[9 * t + 6 + a];
			b.lo[a] = std::min(v0, std::min(v1, v2)); b.hi[a] = std::max(v0, std::max(v1, v2));
			const float c = 0.5f * (b.lo[a] + b.hi[a]);
			centroid[3 * (size_t) t + a] = c;
			scene.lo[a] = std::min(scene.lo[a], b.lo[a]); scene.hi[a] = std::max(scene.hi[a], b.hi[a]);
			cbox.lo[a] = std::min(cbox.lo[a], c); cbox.hi[a] = std::max(cbox.hi[a], c);
		}
	}
	float extent = 0.0f;
	for (int a = 0; a != 3; ++a) extent = std::max(extent, std::max(std::fabs(scene.lo[a]), std::fabs(scene.hi[a])));
	const float pad = n ? extent * (1.0f / 65536.0f) : 0.0f; // as in vkr_bvh.cpp: makes the slab test conservative
	// --- Morton order; the original index breaks ties, so the order is unique
	float inv_extent[3];
	for (int a = 0; a != 3; ++a) { const float e = cbox.hi[a] - cbox.lo[a]; inv_extent[a] = (e > 0.0f) ? 1.0f / e : 0.0f; }
	std::vector<uint64_t> codes((size_t) n);
	std::vector<uint32_t> order((size_t) n);
	{
		std::vector<uint64_t> unsorted((size_t) n);
		for (int64_t t = 0; t != n; ++t) { unsorted[(size_t) t] = lbvh_morton_code(&centroid[3 * (size_t) t], cbox.lo, inv_extent); order[(size_t) t] = (uint32_t) t; }
		std::stable_sort(order.begin(), order.end(), [&](uint32_t l, uint32_t r) { return unsorted[l] < unsorted[r]; });
		for (int64_t s = 0; s != n; ++s) codes[(size_t) s] = unsorted[order[(size_t) s]];
	}
	// --- triangle slots in Morton order
	out.tris.resize(12 * (size_t) n); out.tri_ids.resize((size_t) n);
	for (int64_t s = 0; s != n; ++s) {
		const uint32_t t = order[(size_t) s];
		const float* v = vertices + 9 * (size_t) t;
		float* o = &out.tris[12 * (size_t) s];
		o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
		o[3] = v[3] - v[0]; o[4] = v[4] - v[1]; o[5] = v[5] - v[2];
		o[6] = v[6] - v[0]; o[7] = v[7] - v[1]; o[8] = v[8] - v[2];
		o[9] = o[10] = o[11] = 0.0f;
		out.tri_ids[(size_t) s] = t;
	}
	auto write_child = [&](float* dst, int c, const box3* box, int32_t ref) { // centre + half extent rounded up, like vkr_bvh.cpp
		float ctr[3], half[3];
		for (int a = 0; a != 3; ++a) {
			const double lo = box ? (double) box->lo[a] - (double) pad : 0.0, hi = box ? (double) box->hi[a] + (double) pad : 0.0;
			ctr[a] = (float) (0.5 * (lo + hi));
			half[a] = std::nextafter((float) std::max((double) ctr[a] - lo, hi - (double) ctr[a]), std::numeric_limits<float>::infinity());
		}
		float* d = dst + 6 * c;
		d[0] = ctr[0]; d[1] = ctr[1]; d[2] = ctr[2]; d[3] = half[0]; d[4] = half[1]; d[5] = half[2];
		dst[12 + c] = as_float(ref);
	};
	auto leaf_ref = [](int64_t first, int64_t count) { return (int32_t) (0x80000000u | ((uint32_t) first << 4) | (uint32_t) count); };
	constexpr int64_t kLeafSize = 4;
	if (n <= kLeafSize) { // zero or one leaf: the root pair holds it and an empty leaf
		out.nodes.assign(16, 0.0f);
		box3 all = scene;
		write_child(out.nodes.data(), 0, n ? &all : nullptr, n ? leaf_ref(0, n) : (int32_t) 0x80000000u);
		write_child(out.nodes.data(), 1, nullptr, (int32_t) 0x80000000u);
		out.max_depth = 1; out.node_count = 1;
		return;
	}
	// --- binary radix tree: internal node i covers the sorted positions [first[i], last[i]] and splits after position split[i]
	const int64_t internal_count = n - 1;
	std::vector<int32_t> first((size_t) internal_count), last((size_t) internal_count), split((size_t) internal_count), parent((size_t) internal_count, -1);
	for (int64_t i = 0; i != internal_count; ++i) {
		const int d = (delta(codes, i, i + 1) - delta(codes, i, i - 1)) >= 0 ? 1 : -1;
		const int delta_min = delta(codes, i, i - d);
		int64_t l_max = 2;
		while (delta(codes, i, i + l_max * d) > delta_min) l_max *= 2;
		int64_t l = 0;
		for (int64_t t = l_max / 2; t >= 1; t /= 2)
			if (delta(codes, i, i + (l + t) * d) > delta_min) l += t;
		const int64_t j = i + l * d;
		const int delta_node = delta(codes, i, j);
		int64_t s = 0;
		for (int64_t t = (l + 1) / 2; ; t = (t + 1) / 2) {
			if (delta(codes, i, i + (s + t) * d) > delta_node) s += t;
			if (t == 1) break;
		}
		const int64_t gamma = i + s * d + std::min(d, 0);
		first[(size_t) i] = (int32_t) std::min(i, j); last[(size_t) i] = (int32_t) std::max(i, j); split[(size_t) i] = (int32_t) gamma;
	}
	for (int64_t i = 0; i != internal_count; ++i) { // children that are internal nodes: gamma if the left part has more than one position, gamma + 1 likewise
		const int64_t g = split[(size_t) i];
		if (g != first[(size_t) i]) parent[(size_t) g] = (int32_t) i;
		if (g + 1 != last[(size_t) i]) parent[(size_t) (g + 1)] = (int32_t) i;
	}
	// --- boxes of all internal nodes (union over their positions): children before parents = decreasing size of the range
	std::vector<box3> node_box((size_t) internal_count);
	{
		std::vector<int32_t> by_size((size_t) internal_count);
		for (int64_t i = 0; i != internal_count; ++i) by_size[(size_t) i] = (int32_t) i;
		std::sort(by_size.begin(), by_size.end(), [&](int32_t l, int32_t r) { const int32_t sl = last[(size_t) l] - first[(size_t) l], sr = last[(size_t) r] - first[(size_t) r]; return sl < sr || (sl == sr && l < r); });
		auto child_box = [&](int64_t node, bool right) -> const box3& {
			const int64_t g = split[(size_t) node];
			if (!right) return (g == first[(size_t) node]) ? tri_box[order[(size_t) g]] : node_box[(size_t) g];
			return (g + 1 == last[(size_t) node]) ? tri_box[order[(size_t) (g + 1)]] : node_box[(size_t) (g + 1)];
		};
		for (int32_t k : by_size) {
			const box3& l = child_box(k, false); const box3& r = child_box(k, true);
			for (int a = 0; a != 3; ++a) { node_box[(size_t) k].lo[a] = std::min(l.lo[a], r.lo[a]); node_box[(size_t) k].hi[a] = std::max(l.hi[a], r.hi[a]); }
		}
	}
	// --- collapse: an internal node with at most four positions becomes a leaf; the others are numbered in index order
	std::vector<uint32_t> rank((size_t) internal_count);
	uint32_t used_count = 0;
	auto used = [&](int64_t i) { return (int64_t) last[(size_t) i] - first[(size_t) i] + 1 > kLeafSize; };
	for (int64_t i = 0; i != internal_count; ++i) { rank[(size_t) i] = used_count; used_count += used(i) ? 1u : 0u; }
	out.nodes.assign(16 * (size_t) used_count, 0.0f);
	for (int64_t i = 0; i != internal_count; ++i) {
		if (!used(i)) continue;
		float* dst = &out.nodes[16 * (size_t) rank[(size_t) i]];
		const int64_t g = split[(size_t) i];
		for (int c = 0; c != 2; ++c) {
			const int64_t lo = c ? g + 1 : first[(size_t) i], hi = c ? last[(size_t) i] : g;
			if (lo == hi) write_child(dst, c, &tri_box[order[(size_t) lo]], leaf_ref(lo, 1));
			else {
				const int64_t child = c ? g + 1 : g;
				write_child(dst, c, &node_box[(size_t) child], used(child) ? (int32_t) rank[(size_t) child] : leaf_ref(lo, hi - lo + 1));
			}
		}
	}
	// --- depth in node pairs (the traversal stack is sized by it)
	uint32_t max_depth = 1;
	for (int64_t i = 0; i != internal_count; ++i) {
		if (!used(i)) continue;
		uint32_t depth = 1;
		for (int32_t p = parent[(size_t) i]; p >= 0; p = parent[(size_t) p]) ++depth;
		max_depth = std::max(max_depth, depth);
	}
	out.max_depth = max_depth; out.node_count = used_count;
}

} // namespace vkr
