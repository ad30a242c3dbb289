// oracle/glsl_compat/ref_common.cpp -- TEST INFRASTRUCTURE. Globals of the GLSL-on-C++ layer and the
// ray-query hook that routes rayQueryInitializeEXT() to the oracle's BVH (bvh_oracle.h defines the predicate the
// un-pinned Vulkan driver would otherwise supply).
#include "glsl_compat.hpp"
#include <omp.h>
extern "C" {
#include "../bvh_oracle.h"
}
namespace glsl {
occluded_hook_t g_occluded_hook = nullptr;
const void* g_occluded_user = nullptr;
thread_local uint32_t g_current_visibility = 0xFFFFFFFFu;
const uint8_t* g_material_index_bytes = nullptr;
}
extern "C" void* ref_bvh_create(const float* tris, uint32_t tri_count) {
	obvh_t* bvh = (obvh_t*) malloc(sizeof(obvh_t));
	obvh_build(bvh, tris, tri_count);
	return bvh;
}
extern "C" void ref_bvh_destroy(void* bvh) { obvh_destroy((obvh_t*) bvh); free(bvh); }
extern "C" int ref_bvh_occluded(const void* user, const float* o, const float* d, float tmin, float tmax) {
	return obvh_occluded((const obvh_t*) user, mk3(o[0], o[1], o[2]), mk3(d[0], d[1], d[2]), tmin, tmax);
}
extern "C" int ref_thread_count(void) { return omp_get_max_threads(); }
