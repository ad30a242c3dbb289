// vkr_bvh.cpp -- binned-SAH BVH2 build on the host, flattened into 64-byte node pairs.
//
// Quality matters more than build time here: the tree is built once per scene load and then
// traversed by billions of shadow rays per frame. Top-down build, 16 bins on each axis, leaves of
// at most 4 triangles, subtrees built in parallel (OpenMP tasks) with a deterministic result
// (every subtree owns a fixed range of the triangle order array; node numbering happens in a
// serial flattening pass afterwards). Leaf boxes are padded by 2^-16 of the scene extent so that
// the slab test is conservative w.r.t. the fp32 triangle predicate ("Shadow predicate").
// Boxes are stored as centre + half extent.
#include "vkr_bvh.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>

// builder parameters (tuning experiments may override them on the compiler's command line)
#ifndef VKR_BVH_BINS
#define VKR_BVH_BINS 16
#endif
#ifndef VKR_BVH_LEAF_SIZE
#define VKR_BVH_LEAF_SIZE 4
#endif

namespace vkr {
namespace {

struct box3 {
	float lo[3], hi[3];
	void clear() { for (int a = 0; a != 3; ++a) { lo[a] = std::numeric_limits<float>::infinity(); hi[a] = -std::numeric_limits<float>::infinity(); } }
	void grow(const float* p) { for (int a = 0; a != 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
	void merge(const box3& o) { for (int a = 0; a != 3; ++a) { lo[a] = std::min(lo[a], o.lo[a]); hi[a] = std::max(hi[a], o.hi[a]); } }
	float half_area() const {
		const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
		if (!(dx >= 0.0f)) return 0.0f;
		return dx * dy + dy * dz + dz * dx;
	}
};

struct build_node {
	box3 box;
	std::unique_ptr<build_node> child[2];
	uint32_t first = 0, count = 0; // leaf range in the order array
};

struct builder {
	const float* vertices;
	std::vector<box3> tri_box;
	std::vector<float> centroid; // 3 per triangle
	std::vector<uint32_t> order;
	static constexpr int kBins = VKR_BVH_BINS;
	static constexpr uint32_t kLeafSize = VKR_BVH_LEAF_SIZE;
	static constexpr uint32_t kMedianDepth = 32;

	void build(build_node* node, uint32_t first, uint32_t count, uint32_t depth) {
		box3 cbox; cbox.clear(); node->box.clear();
		for (uint32_t i = first; i != first + count; ++i) {
			const uint32_t t = order[i];
			node->box.merge(tri_box[t]);
			cbox.grow(&centroid[3 * (size_t) t]);
		}
		node->first = first; node->count = count;
		if (count <= kLeafSize) return;
		uint32_t mid = first + count / 2;
		int axis = -1; int split_bin = 0; float axis_lo = 0.0f, axis_scale = 0.0f;
		if (depth < kMedianDepth) {
			float best = std::numeric_limits<float>::infinity();
			for (int a = 0; a != 3; ++a) {
				const float lo = cbox.lo[a], ext = cbox.hi[a] - cbox.lo[a];
				if (!(ext > 0.0f)) continue;
				box3 bin_box[kBins]; uint32_t bin_count[kBins];
				for (int b = 0; b != kBins; ++b) { bin_box[b].clear(); bin_count[b] = 0; }
				const float scale = (float) kBins / ext;
				for (uint32_t i = first; i != first + count; ++i) {
					const uint32_t t = order[i];
					int b = (int) ((centroid[3 * (size_t) t + a] - lo) * scale);
					b = std::min(std::max(b, 0), kBins - 1);
					bin_box[b].merge(tri_box[t]); ++bin_count[b];
				}
				float right_area[kBins]; uint32_t right_count[kBins];
				box3 acc; acc.clear(); uint32_t n = 0;
				for (int b = kBins - 1; b > 0; --b) { acc.merge(bin_box[b]); n += bin_count[b]; right_area[b] = acc.half_area(); right_count[b] = n; }
				acc.clear(); n = 0;
				for (int b = 0; b + 1 < kBins; ++b) {
					acc.merge(bin_box[b]); n += bin_count[b];
					if (n == 0 || right_count[b + 1] == 0) continue;
					const float cost = acc.half_area() * (float) n + right_area[b + 1] * (float) right_count[b + 1];
					if (cost < best) { best = cost; axis = a; split_bin = b + 1; axis_lo = lo; axis_scale = scale; }
				}
			}
		}
		if (axis >= 0) {
			auto begin = order.begin() + first, end = begin + count;
			auto it = std::partition(begin, end, [&](uint32_t t) {
				int b = (int) ((centroid[3 * (size_t) t + axis] - axis_lo) * axis_scale);
				b = std::min(std::max(b, 0), kBins - 1);
				return b < split_bin;
			});
			mid = (uint32_t) (it - order.begin());
		}
		if (axis < 0 || mid == first || mid == first + count) {
			// object median along the widest centroid axis (also bounds the depth for hostile inputs)
			int a = 0;
			for (int k = 1; k != 3; ++k) if (cbox.hi[k] - cbox.lo[k] > cbox.hi[a] - cbox.lo[a]) a = k;
			mid = first + count / 2;
			std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
				[&](uint32_t l, uint32_t r) { const float cl = centroid[3 * (size_t) l + a], cr = centroid[3 * (size_t) r + a]; return cl < cr || (cl == cr && l < r); });
		}
		node->child[0].reset(new build_node());
		node->child[1].reset(new build_node());
		build_node* c0 = node->child[0].get(); build_node* c1 = node->child[1].get();
		const uint32_t n0 = mid - first, n1 = first + count - mid;
		if (count > 65536) {
			#pragma omp task default(shared) firstprivate(c0, first, n0, depth)
			build(c0, first, n0, depth + 1);
			#pragma omp task default(shared) firstprivate(c1, mid, n1, depth)
			build(c1, mid, n1, depth + 1);
			#pragma omp taskwait
		}
		else {
			build(c0, first, n0, depth + 1);
			build(c1, mid, n1, depth + 1);
		}
	}
};

inline float as_float(int32_t bits) { float f; std::memcpy(&f, &bits, 4); return f; }

} // namespace

void shadow_grid_from_root(const float* p, float grid[6]) {
	// root pair: child 0 = centre p[0..2], half extent p[3..5]; child 1 = p[6..8], p[9..11]
	float extent_max = 0.0f, lo[3], hi[3];
	for (int a = 0; a != 3; ++a) {
		const bool has1 = p[9 + a] >= 0.0f;   // an empty second child (single-leaf trees) has a negative half extent
		lo[a] = std::min(p[a] - p[3 + a], has1 ? p[6 + a] - p[9 + a] : p[a] - p[3 + a]);
		hi[a] = std::max(p[a] + p[3 + a], has1 ? p[6 + a] + p[9 + a] : p[a] + p[3 + a]);
		extent_max = std::max(extent_max, hi[a] - lo[a]);
	}
	for (int a = 0; a != 3; ++a) {
		float extent = hi[a] - lo[a];
		if (!(extent > 1.0e-6f * extent_max) || !(extent > 0.0f)) extent = (extent_max > 0.0f) ? 1.0e-6f * extent_max : 1.0f;   // flat scenes: any finite cell size does
		const float cells_per_unit = 65531.0f / extent;   // coordinates land in [2, 65533]: the rounding outwards plus one cell never leaves the 16 bits
		grid[3 + a] = cells_per_unit;
		grid[a] = lo[a] - 2.0f / cells_per_unit;
	}
}

void build_bvh(host_bvh& out, const float* vertices, uint64_t triangle_count) {
	out = host_bvh();
	const uint32_t n = (uint32_t) triangle_count;
	builder b;
	b.vertices = vertices;
	b.tri_box.resize(n); b.centroid.resize(3 * (size_t) n); b.order.resize(n);
	box3 scene; scene.clear();
	for (uint32_t t = 0; t != n; ++t) {
		b.tri_box[t].clear();
		for (int k = 0; k != 3; ++k) b.tri_box[t].grow(vertices + 9 * (size_t) t + 3 * k);
		for (int a = 0; a != 3; ++a) b.centroid[3 * (size_t) t + a] = 0.5f * (b.tri_box[t].lo[a] + b.tri_box[t].hi[a]);
		scene.merge(b.tri_box[t]);
		b.order[t] = t;
	}
	float extent = 0.0f;
	for (int a = 0; a != 3; ++a) extent = std::max(extent, std::max(std::fabs(scene.lo[a]), std::fabs(scene.hi[a])));
	const float pad = n ? extent * (1.0f / 65536.0f) : 0.0f;
	build_node root;
	if (n) {
		#pragma omp parallel
		#pragma omp single nowait
		b.build(&root, 0, n, 0);
	}
	// --- flatten: inner nodes in depth-first order, leaves become references into the slot arrays
	out.tris.resize(12 * (size_t) n); out.tri_ids.resize(n);
	for (uint32_t s = 0; s != n; ++s) {
		const uint32_t t = b.order[s];
		const float* v = vertices + 9 * (size_t) t;
		float* o = &out.tris[12 * (size_t) s];
		o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
		o[3] = v[3] - v[0]; o[4] = v[4] - v[1]; o[5] = v[5] - v[2];
		o[6] = v[6] - v[0]; o[7] = v[7] - v[1]; o[8] = v[8] - v[2];
		o[9] = o[10] = o[11] = 0.0f;
		out.tri_ids[s] = t;
	}
	struct item { const build_node* node; uint32_t index; uint32_t depth; };
	std::vector<item> stack;
	auto leaf_ref = [](const build_node* nd) { return (int32_t) (0x80000000u | (nd->first << 4) | nd->count); };
	// Child boxes are stored as centre and half extent (the slab test then needs no per-axis min/max, vkr_trace.cuh).
	// The half extent is rounded up so that [c - h, c + h] contains the padded box.
	auto write_child = [&](float* dst, int c, const build_node* nd, int32_t ref) {
		float ctr[3], half[3];
		for (int a = 0; a != 3; ++a) {
			const double lo = nd ? (double) nd->box.lo[a] - (double) pad : 0.0, hi = nd ? (double) nd->box.hi[a] + (double) pad : 0.0;
			ctr[a] = (float) (0.5 * (lo + hi));
			half[a] = std::nextafter((float) std::max((double) ctr[a] - lo, hi - (double) ctr[a]), std::numeric_limits<float>::infinity());
		}
		float* d = dst + 6 * c;
		d[0] = ctr[0]; d[1] = ctr[1]; d[2] = ctr[2]; d[3] = half[0]; d[4] = half[1]; d[5] = half[2];
		dst[12 + c] = as_float(ref);
	};
	out.nodes.assign(16, 0.0f);
	uint32_t next = 1;
	if (!n || !root.child[0]) {
		// zero or one leaf: the root pair holds the leaf and an empty leaf (count 0 is never tested)
		float* dst = out.nodes.data();
		write_child(dst, 0, n ? &root : nullptr, n ? leaf_ref(&root) : (int32_t) 0x80000000u);
		write_child(dst, 1, nullptr, (int32_t) 0x80000000u);
		out.max_depth = 1;
	}
	else {
		stack.push_back({&root, 0, 1});
		while (!stack.empty()) {
			const item it = stack.back(); stack.pop_back();
			out.max_depth = std::max(out.max_depth, it.depth);
			int32_t refs[2];
			for (int c = 0; c != 2; ++c) {
				const build_node* ch = it.node->child[c].get();
				if (ch->child[0]) {
					refs[c] = (int32_t) next++;
					out.nodes.resize(16 * (size_t) next, 0.0f);
					stack.push_back({ch, (uint32_t) refs[c], it.depth + 1});
				}
				else refs[c] = leaf_ref(ch);
			}
			float* dst = &out.nodes[16 * (size_t) it.index];
			for (int c = 0; c != 2; ++c) write_child(dst, c, it.node->child[c].get(), refs[c]);
		}
	}
	out.node_count = next;
}

// Collapses node pairs into 4-wide nodes: starting from the two children of a pair, the inner child with the largest surface area is
// replaced by its own two children until there are four (or only leaves are left). Boxes are copied as they are (already padded and
// rounded outwards), so the 4-wide tree is exactly as conservative as the binary one. Deterministic; nodes are numbered depth first.
void build_bvh4_from_bvh2(host_bvh4& out, const host_bvh& in) {
	out = host_bvh4();
	struct child { float box[6]; int32_t ref; };
	auto load = [&](uint32_t pair, int c) {
		child ch;
		const float* src = &in.nodes[16 * (size_t) pair];
		std::memcpy(ch.box, src + 6 * c, sizeof(ch.box));
		std::memcpy(&ch.ref, src + 12 + c, 4);
		return ch;
	};
	auto area = [](const child& ch) { return ch.box[3] * ch.box[4] + ch.box[4] * ch.box[5] + ch.box[5] * ch.box[3]; };
	struct item { uint32_t pair, index, depth; };
	std::vector<item> stack;
	out.nodes.assign(32, 0.0f);
	uint32_t next = 1;
	stack.push_back({0, 0, 1});
	while (!stack.empty()) {
		const item it = stack.back(); stack.pop_back();
		out.max_depth = std::max(out.max_depth, it.depth);
		child children[4]; int count = 2;
		children[0] = load(it.pair, 0); children[1] = load(it.pair, 1);
		while (count < 4) {
			int widest = -1;
			for (int c = 0; c != count; ++c)
				if (children[c].ref >= 0 && (widest < 0 || area(children[c]) > area(children[widest]))) widest = c;
			if (widest < 0) break;
			const uint32_t pair = (uint32_t) children[widest].ref;
			children[widest] = load(pair, 0);
			children[count++] = load(pair, 1);
		}
		int32_t refs[4];
		for (int c = 0; c != 4; ++c) {
			if (c >= count) { refs[c] = (int32_t) 0x80000000u; continue; }
			if (children[c].ref >= 0) {
				refs[c] = (int32_t) next++;
				out.nodes.resize(32 * (size_t) next, 0.0f);
				stack.push_back({(uint32_t) children[c].ref, (uint32_t) refs[c], it.depth + 1});
			}
			else refs[c] = children[c].ref;
		}
		float* dst = &out.nodes[32 * (size_t) it.index];
		for (int c = 0; c != 4; ++c) {
			if (c < count) std::memcpy(dst + 6 * c, children[c].box, sizeof(children[c].box));
			else { dst[6 * c] = dst[6 * c + 1] = dst[6 * c + 2] = 0.0f; dst[6 * c + 3] = dst[6 * c + 4] = dst[6 * c + 5] = -1.0f; }
			dst[24 + c] = as_float(refs[c]);
		}
	}
	out.node_count = next;
}

} // namespace vkr
