// vkr_bvh.h -- host-side BVH2 builder producing the node-pair layout of vkr_trace.cuh.
// Replaces the driver's acceleration-structure build (vkCmdBuildAccelerationStructuresKHR,
// src/scene.c:354-378) for a triangle soup with stride 12 (scene.c:197-209).
#pragma once
#include <stdint.h>
#include <vector>

namespace vkr {

struct host_bvh {
	std::vector<float> nodes;       // 16 floats per inner node (4 x float4)
	std::vector<float> tris;        // 12 floats per triangle slot (3 x float4)
	std::vector<uint32_t> tri_ids;  // original triangle index per slot
	uint32_t max_depth = 0;
	uint64_t node_count = 0;
};

// vertices: 9 floats per triangle. The result is deterministic for a given input.
void build_bvh(host_bvh& out, const float* vertices, uint64_t triangle_count);

} // namespace vkr
