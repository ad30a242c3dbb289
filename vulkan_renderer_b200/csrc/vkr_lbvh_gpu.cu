// vkr_lbvh_gpu.cu -- linear BVH built on the GPU (replaces the driver's acceleration-structure build,
// vkCmdBuildAccelerationStructuresKHR at src/scene.c:354-378, which also runs on the GPU).
//
// The parallel form of vkr_lbvh.cpp, step for step, so that both produce the same bytes:
//   1. triangle boxes, centroids, scene / centroid bounds            one thread per triangle, atomicMin/Max on order-preserving keys
//   2. 63-bit Morton codes, stable radix sort by code                 cub::DeviceRadixSort (values start as 0..n-1, so ties keep the index order)
//   3. triangle slots (v0, e1, e2) and original indices in that order one thread per slot
//   4. binary radix tree (Karras, HPG 2012)                           one thread per internal node, no synchronisation
//   5. boxes bottom-up                                                 one thread per leaf walks up; the second arrival at a node merges
//   6. collapse subtrees of <= 4 triangles, number the rest           exclusive prefix sum over "has more than four triangles"
//   7. node pairs (centre + half extent rounded up, child references)  one thread per internal node
//   8. depth in node pairs                                             one thread per internal node walks up, atomicMax
// Everything is integer or exact min/max work except the box encoding of step 7, which repeats the host's double-precision
// arithmetic (-fmad=false: no contraction). HBM traffic ~ 0.4 KB per triangle; the radix sort (8 passes over 12-byte pairs) dominates.
// Selected with VKR_BVH_BUILDER=lbvh_gpu; the default builder stays the binned-SAH one (better trees for the benchmark).
#include "vkr_bvh.h"
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace vkr {
namespace {

constexpr int kThreads = 256;
constexpr int kLeafSize = 4;

// float <-> unsigned key with the same order, for atomicMin / atomicMax
__host__ __device__ inline uint32_t float_to_ordered(float f) { uint32_t u; memcpy(&u, &f, 4); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ inline float ordered_to_float(uint32_t k) { const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k; float f; memcpy(&f, &u, 4); return f; }

struct bounds_keys { uint32_t scene_lo[3], scene_hi[3], centroid_lo[3], centroid_hi[3]; };

__global__ void __launch_bounds__(kThreads) triangle_bounds_kernel(const float* __restrict__ vertices, uint32_t n, float* __restrict__ box_lo, float* __restrict__ box_hi, float* __restrict__ centroid, bounds_keys* keys) {
	const uint32_t t = blockIdx.x * kThreads + threadIdx.x;
	const bool valid = t < n;
	const float* v = vertices + 9 * (size_t) (valid ? t : 0);
#pragma unroll
	for (int a = 0; a != 3; ++a) {
		const float v0 = v[a], v1 = v[3 + a], v2 = v[6 + a];
		const float lo = fminf(v0, fminf(v1, v2)), hi = fmaxf(v0, fmaxf(v1, v2));
		const float c = 0.5f * (lo + hi);
		if (valid) { box_lo[3 * (size_t) t + a] = lo; box_hi[3 * (size_t) t + a] = hi; centroid[3 * (size_t) t + a] = c; }
		// one atomic per warp and bound: min / max over the warp first (lanes past the end contribute the neutral keys)
		const uint32_t k_lo = __reduce_min_sync(0xffffffffu, valid ? float_to_ordered(lo) : 0xffffffffu), k_hi = __reduce_max_sync(0xffffffffu, valid ? float_to_ordered(hi) : 0u);
		const uint32_t k_clo = __reduce_min_sync(0xffffffffu, valid ? float_to_ordered(c) : 0xffffffffu), k_chi = __reduce_max_sync(0xffffffffu, valid ? float_to_ordered(c) : 0u);
		if ((threadIdx.x & 31) == 0) {
			atomicMin(&keys->scene_lo[a], k_lo); atomicMax(&keys->scene_hi[a], k_hi);
			atomicMin(&keys->centroid_lo[a], k_clo); atomicMax(&keys->centroid_hi[a], k_chi);
		}
	}
}

__device__ inline uint64_t expand_bits_21(uint32_t v) {
	uint64_t x = v & 0x1fffffu;
	x = (x | x << 32) & 0x1f00000000ffffull;
	x = (x | x << 16) & 0x1f0000ff0000ffull;
	x = (x | x << 8) & 0x100f00f00f00f00full;
	x = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}

struct f3pod { float x, y, z; };

__global__ void __launch_bounds__(kThreads) morton_kernel(const float* __restrict__ centroid, uint32_t n, f3pod lo, f3pod inv_extent, uint64_t* __restrict__ codes, uint32_t* __restrict__ indices) {
	const uint32_t t = blockIdx.x * kThreads + threadIdx.x;
	if (t >= n) return;
	const float l[3] = { lo.x, lo.y, lo.z }, ie[3] = { inv_extent.x, inv_extent.y, inv_extent.z };
	uint32_t q[3];
#pragma unroll
	for (int a = 0; a != 3; ++a) {
		const float f = ((centroid[3 * (size_t) t + a] - l[a]) * ie[a]) * 2097152.0f;
		q[a] = (f > 0.0f) ? ((f < 2097151.0f) ? (uint32_t) f : 2097151u) : 0u;
	}
	codes[t] = expand_bits_21(q[0]) << 2 | expand_bits_21(q[1]) << 1 | expand_bits_21(q[2]);
	indices[t] = t;
}

__global__ void __launch_bounds__(kThreads) slots_kernel(const float* __restrict__ vertices, const uint32_t* __restrict__ order, uint32_t n, float4* __restrict__ tris, uint32_t* __restrict__ tri_ids) {
	const uint32_t s = blockIdx.x * kThreads + threadIdx.x;
	if (s >= n) return;
	const uint32_t t = order[s];
	const float* v = vertices + 9 * (size_t) t;
	tris[3 * (size_t) s] = make_float4(v[0], v[1], v[2], v[3] - v[0]);
	tris[3 * (size_t) s + 1] = make_float4(v[4] - v[1], v[5] - v[2], v[6] - v[0], v[7] - v[1]);
	tris[3 * (size_t) s + 2] = make_float4(v[8] - v[2], 0.0f, 0.0f, 0.0f);
	tri_ids[s] = t;
}

// Length of the common prefix of the keys at sorted positions i and j; the position breaks ties between equal codes
__device__ inline int delta(const uint64_t* __restrict__ codes, int64_t n, int64_t i, int64_t j) {
	if (j < 0 || j >= n) return -1;
	const uint64_t x = codes[i] ^ codes[j];
	return x ? __clzll((long long) x) : 64 + __clz((int) ((uint32_t) i ^ (uint32_t) j));
}

__global__ void __launch_bounds__(kThreads) radix_tree_kernel(const uint64_t* __restrict__ codes, uint32_t n, int32_t* __restrict__ first, int32_t* __restrict__ last, int32_t* __restrict__ split,
	int32_t* __restrict__ parent, int32_t* __restrict__ leaf_parent)
{
	const int64_t i = (int64_t) blockIdx.x * kThreads + threadIdx.x;
	const int64_t count = n;
	if (i >= count - 1) return;
	const int d = (delta(codes, count, i, i + 1) - delta(codes, count, i, i - 1)) >= 0 ? 1 : -1;
	const int delta_min = delta(codes, count, i, i - d);
	int64_t l_max = 2;
	while (delta(codes, count, i, i + l_max * d) > delta_min) l_max *= 2;
	int64_t l = 0;
	for (int64_t t = l_max / 2; t >= 1; t /= 2)
		if (delta(codes, count, i, i + (l + t) * d) > delta_min) l += t;
	const int64_t j = i + l * d;
	const int delta_node = delta(codes, count, i, j);
	int64_t s = 0;
	for (int64_t t = (l + 1) / 2; ; t = (t + 1) / 2) {
		if (delta(codes, count, i, i + (s + t) * d) > delta_node) s += t;
		if (t == 1) break;
	}
	const int64_t gamma = i + s * d + (d < 0 ? d : 0);
	const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
	first[i] = (int32_t) lo; last[i] = (int32_t) hi; split[i] = (int32_t) gamma;
	if (gamma == lo) leaf_parent[gamma] = (int32_t) i; else parent[gamma] = (int32_t) i;
	if (gamma + 1 == hi) leaf_parent[gamma + 1] = (int32_t) i; else parent[gamma + 1] = (int32_t) i;
}

__global__ void __launch_bounds__(kThreads) refit_kernel(uint32_t n, const uint32_t* __restrict__ order, const float* __restrict__ box_lo, const float* __restrict__ box_hi,
	const int32_t* __restrict__ first, const int32_t* __restrict__ last, const int32_t* __restrict__ split, const int32_t* __restrict__ parent, const int32_t* __restrict__ leaf_parent,
	float* node_lo, float* node_hi, uint32_t* arrivals)
{
	const uint32_t s = blockIdx.x * kThreads + threadIdx.x;
	if (s >= n) return;
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

__global__ void __launch_bounds__(kThreads) used_kernel(uint32_t internal_count, const int32_t* __restrict__ first, const int32_t* __restrict__ last, uint32_t* __restrict__ used) {
	const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
	if (i < internal_count) used[i] = (last[i] - first[i] + 1 > kLeafSize) ? 1u : 0u;
}

// Centre and half extent of a padded box, the half extent rounded up: the same double-precision steps as write_child() in vkr_lbvh.cpp
__device__ inline void encode_box(float* dst6, const float* lo3, const float* hi3, float pad) {
#pragma unroll
	for (int a = 0; a != 3; ++a) {
		const double lo = (double) lo3[a] - (double) pad, hi = (double) hi3[a] + (double) pad;
		const float ctr = (float) (0.5 * (lo + hi));
		const double up = (double) ctr - lo, down = hi - (double) ctr;
		dst6[a] = ctr;
		dst6[3 + a] = nextafterf((float) (up > down ? up : down), INFINITY);
	}
}

__global__ void __launch_bounds__(kThreads) emit_kernel(uint32_t internal_count, const uint32_t* __restrict__ order, const float* __restrict__ box_lo, const float* __restrict__ box_hi,
	const float* __restrict__ node_lo, const float* __restrict__ node_hi, const int32_t* __restrict__ first, const int32_t* __restrict__ last, const int32_t* __restrict__ split,
	const uint32_t* __restrict__ used, const uint32_t* __restrict__ rank, float pad, float* __restrict__ nodes)
{
	const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= internal_count || !used[i]) return;
	float out[16];
#pragma unroll
	for (int k = 0; k != 16; ++k) out[k] = 0.0f;
	const int32_t g = split[i];
#pragma unroll
	for (int c = 0; c != 2; ++c) {
		const int32_t lo = c ? g + 1 : first[i], hi = c ? last[i] : g;
		int32_t ref;
		if (lo == hi) {
			encode_box(out + 6 * c, box_lo + 3 * (size_t) order[lo], box_hi + 3 * (size_t) order[lo], pad);
			ref = (int32_t) (0x80000000u | ((uint32_t) lo << 4) | 1u);
		}
		else {
			const int32_t child = c ? g + 1 : g;
			encode_box(out + 6 * c, node_lo + 3 * (size_t) child, node_hi + 3 * (size_t) child, pad);
			ref = used[child] ? (int32_t) rank[child] : (int32_t) (0x80000000u | ((uint32_t) lo << 4) | (uint32_t) (hi - lo + 1));
		}
		out[12 + c] = __int_as_float(ref);
	}
	float4* dst = reinterpret_cast<float4*>(nodes + 16 * (size_t) rank[i]);
#pragma unroll
	for (int k = 0; k != 4; ++k) dst[k] = make_float4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
}

__global__ void __launch_bounds__(kThreads) depth_kernel(uint32_t internal_count, const uint32_t* __restrict__ used, const int32_t* __restrict__ parent, uint32_t* max_depth) {
	const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= internal_count || !used[i]) return;
	uint32_t depth = 1;
	for (int32_t p = parent[i]; p >= 0; p = parent[p]) ++depth;
	atomicMax(max_depth, depth);
}

struct scratch { // device allocations that live as long as one build
	static constexpr int kCapacity = 32;
	void* ptr[kCapacity] = {};
	int count = 0;
	template <class T> bool alloc(T** p, size_t elements) {
		*p = nullptr;
		if (count == kCapacity || cudaMalloc((void**) p, sizeof(T) * (elements ? elements : 1)) != cudaSuccess) { *p = nullptr; return false; }
		ptr[count++] = *p;
		return true;
	}
	~scratch() { for (int i = 0; i != count; ++i) cudaFree(ptr[i]); }
};

} // namespace

// Builds on the current device. vertices: HOST pointer, 9 floats per triangle. On success *d_nodes (16 floats per pair), *d_tris (12 floats per slot)
// and *d_tri_ids are device allocations owned by the caller (cudaFree). Returns non-zero (and allocates nothing) on failure or if there are
// fewer than five triangles (the host builder handles those).
int build_lbvh_device(const float* vertices, uint64_t triangle_count, void* stream_handle, void** d_nodes, void** d_tris, void** d_tri_ids, uint64_t* node_count, uint32_t* max_depth) {
	cudaStream_t stream = (cudaStream_t) stream_handle;
	*d_nodes = *d_tris = *d_tri_ids = nullptr; *node_count = 0; *max_depth = 0;
	if (triangle_count <= (uint64_t) kLeafSize || triangle_count >= (1ull << 27)) return 1;
	const uint32_t n = (uint32_t) triangle_count, internal_count = n - 1;
	const uint32_t blocks = (n + kThreads - 1) / kThreads;
	scratch s;
	float *d_vertices, *box_lo, *box_hi, *centroid, *node_lo, *node_hi;
	uint64_t *codes_in, *codes; uint32_t *index_in, *order, *arrivals, *used, *rank, *d_depth; bounds_keys* keys;
	int32_t *first, *last, *split, *parent, *leaf_parent;
	if (!s.alloc(&d_vertices, 9 * (size_t) n) || !s.alloc(&box_lo, 3 * (size_t) n) || !s.alloc(&box_hi, 3 * (size_t) n) || !s.alloc(&centroid, 3 * (size_t) n)
		|| !s.alloc(&codes_in, n) || !s.alloc(&codes, n) || !s.alloc(&index_in, n) || !s.alloc(&order, n) || !s.alloc(&keys, 1)) return 1;
	if (cudaMemcpyAsync(d_vertices, vertices, sizeof(float) * 9 * (size_t) n, cudaMemcpyHostToDevice, stream) != cudaSuccess) return 1;
	// 1. boxes and bounds
	bounds_keys init;
	for (int a = 0; a != 3; ++a) { init.scene_lo[a] = init.centroid_lo[a] = 0xffffffffu; init.scene_hi[a] = init.centroid_hi[a] = 0u; }
	cudaMemcpyAsync(keys, &init, sizeof(init), cudaMemcpyHostToDevice, stream);
	triangle_bounds_kernel<<<blocks, kThreads, 0, stream>>>(d_vertices, n, box_lo, box_hi, centroid, keys);
	bounds_keys found;
	if (cudaMemcpyAsync(&found, keys, sizeof(found), cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) return 1;
	float extent = 0.0f; f3pod lo, inv; float* lo_a[3] = { &lo.x, &lo.y, &lo.z }; float* inv_a[3] = { &inv.x, &inv.y, &inv.z };
	for (int a = 0; a != 3; ++a) {
		extent = fmaxf(extent, fmaxf(fabsf(ordered_to_float(found.scene_lo[a])), fabsf(ordered_to_float(found.scene_hi[a]))));
		const float c_lo = ordered_to_float(found.centroid_lo[a]), e = ordered_to_float(found.centroid_hi[a]) - c_lo;
		*lo_a[a] = c_lo; *inv_a[a] = (e > 0.0f) ? 1.0f / e : 0.0f;
	}
	const float pad = extent * (1.0f / 65536.0f);
	// 2. Morton codes and their order
	morton_kernel<<<blocks, kThreads, 0, stream>>>(centroid, n, lo, inv, codes_in, index_in);
	size_t sort_bytes = 0;
	if (cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, codes_in, codes, index_in, order, (int) n, 0, 63, stream) != cudaSuccess) return 1;
	size_t scan_bytes = 0;
	if (cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t*) nullptr, (uint32_t*) nullptr, (int) internal_count, stream) != cudaSuccess) return 1;
	unsigned char* cub_temp;
	if (!s.alloc(&cub_temp, sort_bytes > scan_bytes ? sort_bytes : scan_bytes)) return 1;
	if (cub::DeviceRadixSort::SortPairs(cub_temp, sort_bytes, codes_in, codes, index_in, order, (int) n, 0, 63, stream) != cudaSuccess) return 1;
	// 3. slots (outputs)
	float4* tris; uint32_t* tri_ids;
	if (cudaMalloc((void**) &tris, sizeof(float) * 12 * (size_t) n) != cudaSuccess) return 1;
	if (cudaMalloc((void**) &tri_ids, sizeof(uint32_t) * (size_t) n) != cudaSuccess) { cudaFree(tris); return 1; }
	auto fail = [&]() { cudaFree(tris); cudaFree(tri_ids); return 1; };
	slots_kernel<<<blocks, kThreads, 0, stream>>>(d_vertices, order, n, tris, tri_ids);
	// 4. radix tree
	if (!s.alloc(&first, internal_count) || !s.alloc(&last, internal_count) || !s.alloc(&split, internal_count) || !s.alloc(&parent, internal_count) || !s.alloc(&leaf_parent, n)
		|| !s.alloc(&node_lo, 3 * (size_t) internal_count) || !s.alloc(&node_hi, 3 * (size_t) internal_count) || !s.alloc(&arrivals, internal_count)
		|| !s.alloc(&used, internal_count) || !s.alloc(&rank, internal_count) || !s.alloc(&d_depth, 1)) return fail();
	cudaMemsetAsync(parent, 0xff, sizeof(int32_t) * (size_t) internal_count, stream);   // -1: the root has no parent
	cudaMemsetAsync(arrivals, 0, sizeof(uint32_t) * (size_t) internal_count, stream);
	cudaMemsetAsync(d_depth, 0, sizeof(uint32_t), stream);
	radix_tree_kernel<<<blocks, kThreads, 0, stream>>>(codes, n, first, last, split, parent, leaf_parent);
	// 5. boxes
	refit_kernel<<<blocks, kThreads, 0, stream>>>(n, order, box_lo, box_hi, first, last, split, parent, leaf_parent, node_lo, node_hi, arrivals);
	// 6. collapse + numbering
	used_kernel<<<blocks, kThreads, 0, stream>>>(internal_count, first, last, used);
	if (cub::DeviceScan::ExclusiveSum(cub_temp, scan_bytes, used, rank, (int) internal_count, stream) != cudaSuccess) return fail();
	uint32_t last_rank = 0, last_used = 0;
	if (cudaMemcpyAsync(&last_rank, rank + (internal_count - 1), 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess
		|| cudaMemcpyAsync(&last_used, used + (internal_count - 1), 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) return fail();
	const uint32_t used_count = last_rank + last_used;
	// 7. node pairs, 8. depth
	float* nodes;
	if (cudaMalloc((void**) &nodes, sizeof(float) * 16 * (size_t) (used_count ? used_count : 1)) != cudaSuccess) return fail();
	emit_kernel<<<blocks, kThreads, 0, stream>>>(internal_count, order, box_lo, box_hi, node_lo, node_hi, first, last, split, used, rank, pad, nodes);
	depth_kernel<<<blocks, kThreads, 0, stream>>>(internal_count, used, parent, d_depth);
	uint32_t depth = 0;
	if (cudaMemcpyAsync(&depth, d_depth, 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
		cudaFree(nodes); return fail();
	}
	*d_nodes = nodes; *d_tris = tris; *d_tri_ids = tri_ids; *node_count = used_count; *max_depth = depth;
	return 0;
}

} // namespace vkr
