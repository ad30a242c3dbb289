// tests/device_on_host.cpp -- TEST INFRASTRUCTURE: the product's device sampling headers compiled for the CPU.
//
// The related-work samplers (vulkan_renderer_b200/csrc/vkr_related_work.cuh, on top of vkr_psa.cuh and vkr_device_math.cuh)
// use no warp intrinsics, so g++ can compile the very same source (-ffp-contract=off stands in for nvcc's -fmad=false; the
// headers are built only from IEEE add/mul/div/sqrt/fma). tests/test_device_on_host.py holds the result bit for bit against
// the CPU oracle, which in turn is pinned against the reference shader. This catches transcription errors in the device code
// without a GPU; the -m gpu tests then run the same functions inside the kernel against the reference-shader fixtures.
// Built by __graft_entry__.build() into tests/build/libdevice_on_host.so. Nothing in the product links against it.
#include <cmath>
#include <cstdint>
#include <cstring>

#define VKR_DEVICE_CODE_ON_HOST 1
#define VKR_DEV inline
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }

// the CUDA vector types the headers use, laid out as on the device
struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct ushort4 { uint16_t x, y, z, w; };
typedef int cudaError_t;
typedef void* cudaStream_t;
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
template <class T> static inline T __ldg(const T* p) { return *p; }
// what the builder's per-element functions use; the emulation below runs one element after the other, so plain reads and writes do
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned old = *p; *p += v; return old; }
static inline void __threadfence() {}
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long) x) : 64; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned) x) : 32; }

// binary16 conversion of the output stage: the CPU's own round-to-nearest-even conversion (F16C) stands in for the GPU's cvt.rn.f16.f32
#include <immintrin.h>
struct __half { uint16_t bits; };
static inline __half __float2half_rn(float x) { __half h; h.bits = (uint16_t) _cvtss_sh(x, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC); return h; }
static inline uint16_t __half_as_ushort(__half h) { return h.bits; }

// warp-level intrinsics: declared so that the ray-stream header parses; the TRACE = false instantiations used here never call them
unsigned __ballot_sync(unsigned, int); int __any_sync(unsigned, int); int __all_sync(unsigned, int); void __syncwarp(unsigned); int __popc(unsigned); int __ffs(unsigned);
void __nanosleep(unsigned); unsigned __activemask(); int __shfl_sync(unsigned, int, int); size_t __cvta_generic_to_shared(const void*);

#include "vkr_related_work.cuh"
#include "vkr_trace.cuh"
#include "vkr_anchor.cuh"
#include "vkr_texture.cuh"
#include "vkr_gbuffer.cuh"
#include <algorithm>
using std::min; using std::max;   // the integer min / max of the device headers
#include "vkr_error_display.cuh"
#include "vkr_shade_light.cuh"
#include "vkr_related_work_light.cuh"
#include "vkr_lbvh.cuh"
#include <vector>

using namespace vkr;

template <int TECHNIQUE, int MAXV>
static int run(const unsigned char* light_block, const float* position, const float* frame_rows, uint32_t n, const float* rnd, float* out_dirs, float* out_densities, float* out_ggx_density_factor) {
	rw_light<MAXV> light;
	rw_load_light<MAXV>(light, light_block);
	rw_frame frame;
	frame.rx = make3(frame_rows[0], frame_rows[1], frame_rows[2]); frame.ry = make3(frame_rows[3], frame_rows[4], frame_rows[5]);
	frame.rz = make3(frame_rows[6], frame_rows[7], frame_rows[8]); frame.t = make3(frame_rows[9], frame_rows[10], frame_rows[11]);
	rw_sampler<TECHNIQUE, MAXV> sampler;
	if (!sampler.prepare(light, make3(position[0], position[1], position[2]), frame)) return 0;
	*out_ggx_density_factor = sampler.ggx_density_factor();
	for (uint32_t i = 0; i != n; ++i) {
		const f3 d = sampler.sample(make2(rnd[2 * i], rnd[2 * i + 1]), &out_densities[i]);
		out_dirs[3 * i] = d.x; out_dirs[3 * i + 1] = d.y; out_dirs[3 * i + 2] = d.z;
	}
	return 1;
}

template <int MAXV>
static int run_technique(uint32_t technique, const unsigned char* light_block, const float* position, const float* frame, uint32_t n, const float* rnd, float* out_dirs, float* out_densities, float* out_ggx) {
	switch (technique) {
#define T(K) case K: return run<K, MAXV>(light_block, position, frame, n, rnd, out_dirs, out_densities, out_ggx);
	T(0) T(1) T(2) T(3) T(4) T(5) T(6) T(7) T(8) T(9) T(10)
#undef T
	default: return -1;
	}
}

// Same signature and meaning as vkr_oracle_related_work_batch (oracle/vkr_oracle.h)
extern "C" int vkr_device_on_host_related_work_batch(uint32_t technique, uint32_t maxv, const void* light_block, const float* position, const float* frame,
	uint32_t n, const float* random_numbers, float* out_dirs, float* out_densities, float* out_ggx_density_factor)
{
	const unsigned char* lb = (const unsigned char*) light_block;
	switch (maxv) {
	case 3: return run_technique<3>(technique, lb, position, frame, n, random_numbers, out_dirs, out_densities, out_ggx_density_factor);
	case 4: return run_technique<4>(technique, lb, position, frame, n, random_numbers, out_dirs, out_densities, out_ggx_density_factor);
	case 5: return run_technique<5>(technique, lb, position, frame, n, random_numbers, out_dirs, out_densities, out_ggx_density_factor);
	case 6: return run_technique<6>(technique, lb, position, frame, n, random_numbers, out_dirs, out_densities, out_ggx_density_factor);
	case 7: return run_technique<7>(technique, lb, position, frame, n, random_numbers, out_dirs, out_densities, out_ggx_density_factor);
	default: return -1;
	}
}

// Error display (shading_pass.frag.glsl:462-493, 549-563): clip + prepare + one sample + its error per random number pair, and the colour of
// error component 0. technique 10 = Arvo (two error components), 11 = projected solid angle sampling (biased or not).
template <int MAXV, bool BIASED>
static int run_error(uint32_t technique, uint32_t vertex_count, const float* vertices_xyz, uint32_t n, const float* rnd, float error_factor, float* out_errors, float* out_colors) {
	constexpr int MAXP = MAXV + 1;
	f3 v[MAXP];
	for (int i = 0; i != MAXP; ++i) v[i] = (i < (int) vertex_count) ? make3(vertices_xyz[3 * i], vertices_xyz[3 * i + 1], vertices_xyz[3 * i + 2]) : make3(0.0f, 0.0f, 0.0f);
	for (int i = (int) vertex_count; i < MAXV; ++i) v[i] = v[0]; // write_constants() repeats the first vertex (src/main.c:2176)
	const int vc = clip_polygon<MAXP>((int) vertex_count, v);
	if (vc == 0) return 0;
	psa_polygon<MAXP> polygon;
	psa_arvo_polygon<MAXP> arvo;
	if (technique == 10) { prepare_psa_arvo<MAXP>(arvo, vc, v); if (arvo.psa <= 0.0f) return 0; }
	else { prepare_psa<MAXP, BIASED>(polygon, vc, v); if (polygon.psa <= 0.0f) return 0; }
	for (uint32_t i = 0; i != n; ++i) {
		const f2 r = make2(rnd[2 * i], rnd[2 * i + 1]);
		f3 e;
		if (technique == 10) {
			const f2 e2 = sampling_error_arvo<MAXP>(arvo, r, sample_psa_arvo<MAXP>(arvo, r, 3));
			e = make3(e2.x, e2.y, 0.0f);
		}
		else e = sampling_error<MAXP, BIASED>(polygon, r, sample_psa<MAXP, BIASED>(polygon, r));
		const f3 c = error_to_color(e.x, error_factor);
		out_errors[3 * i] = e.x; out_errors[3 * i + 1] = e.y; out_errors[3 * i + 2] = e.z;
		out_colors[3 * i] = c.x; out_colors[3 * i + 1] = c.y; out_colors[3 * i + 2] = c.z;
	}
	return 1;
}

// Same signature and meaning as vkr_oracle_error_display_batch (oracle/vkr_oracle.h)
extern "C" int vkr_device_on_host_error_display_batch(uint32_t technique, int biased, uint32_t maxv, uint32_t vertex_count, const float* vertices_xyz, uint32_t n, const float* rnd,
	float error_factor, float* out_errors, float* out_colors)
{
	switch (maxv) {
#define V(K) case K: return biased ? run_error<K, true>(technique, vertex_count, vertices_xyz, n, rnd, error_factor, out_errors, out_colors) : run_error<K, false>(technique, vertex_count, vertices_xyz, n, rnd, error_factor, out_errors, out_colors);
	V(3) V(4) V(5) V(6) V(7)
#undef V
	default: return -1;
	}
}

// The any-hit traversal of the device (vkr_trace.cuh: occluded(), ray_box(), ray_triangle()) over a BVH in the node-pair layout, one ray after
// the other: rays = {ox, oy, oz, dx, dy, dz, tmin, tmax}. Lets the host-side tests hold every BVH builder against a brute-force loop over all
// triangles with the very traversal code the GPU runs.
extern "C" void vkr_device_on_host_trace_any(const float* nodes, const float* tris, uint32_t tri_count, uint32_t ray_count, const float* rays, uint8_t* out_bvh, uint8_t* out_brute) {
	bvh_view bvh;
	bvh.nodes = reinterpret_cast<const float4*>(nodes); bvh.tris = reinterpret_cast<const float4*>(tris); bvh.tri_ids = nullptr; bvh.tri_count = tri_count;
	int stack[kMaxStackDepth + 2];
	for (uint32_t i = 0; i != ray_count; ++i) {
		const float* r = rays + 8 * (size_t) i;
		const f3 o = make3(r[0], r[1], r[2]), d = make3(r[3], r[4], r[5]);
		out_bvh[i] = occluded(bvh, o, d, r[6], r[7], stack, 1) ? 1 : 0;
		if (out_brute) {
			bool hit = false; float t;
			if (r[7] > r[6]) for (uint32_t k = 0; k != tri_count && !hit; ++k) hit = ray_triangle(bvh.tris + 3 * (size_t) k, o, d, r[6], r[7], &t);
			out_brute[i] = hit ? 1 : 0;
		}
	}
}

// textureGrad as the G-buffer producer defines it (vkr_texture.cuh), one call per row of inputs {u, v, dudx, dvdx, dudy, dvdy}; same signature
// and meaning as vkr_oracle_texture_grad_batch
extern "C" void vkr_device_on_host_texture_grad_batch(uint32_t width, uint32_t height, uint32_t mip_count, const float* texels, uint32_t n, const float* inputs, float* out_rgba) {
	texture_view view;
	view.width = width; view.height = height; view.mip_count = mip_count; view.texels = reinterpret_cast<const float4*>(texels);
	for (uint32_t i = 0; i != n; ++i) {
		const float* in = inputs + 6 * (size_t) i;
		const float4 c = texture_grad(view, make2(in[0], in[1]), make2(in[2], in[3]), make2(in[4], in[5]));
		out_rgba[4 * i] = c.x; out_rgba[4 * i + 1] = c.y; out_rgba[4 * i + 2] = c.z; out_rgba[4 * i + 3] = c.w;
	}
}

// The same rays through the binary tree and through its 4-wide collapse (vkr_trace.cuh: occluded4): answers and nodes fetched per ray
extern "C" void vkr_device_on_host_trace_any_wide(const float* nodes2, const float* nodes4, const float* tris, uint32_t ray_count, const float* rays, uint8_t* out2, uint8_t* out4,
	uint64_t* out_steps2, uint64_t* out_steps4)
{
	bvh_view bvh;
	bvh.nodes = reinterpret_cast<const float4*>(nodes2); bvh.tris = reinterpret_cast<const float4*>(tris); bvh.tri_ids = nullptr; bvh.tri_count = 0;
	int stack[4 * kMaxStackDepth];
	*out_steps2 = *out_steps4 = 0;
	for (uint32_t i = 0; i != ray_count; ++i) {
		const float* r = rays + 8 * (size_t) i;
		const f3 o = make3(r[0], r[1], r[2]), d = make3(r[3], r[4], r[5]);
		out2[i] = occluded(bvh, o, d, r[6], r[7], stack, 1) ? 1 : 0;
		int steps = 0;
		out4[i] = occluded4(reinterpret_cast<const float4*>(nodes4), bvh.tris, o, d, r[6], r[7], stack, 1, &steps) ? 1 : 0;
		*out_steps4 += (uint64_t) steps;
	}
	(void) out_steps2;
}

// Shader-side vertex decode (vkr_gbuffer.cuh: decode_position, mesh_quantization.glsl:38-45) for all vertices: the input of the primary-ray BVH
// Quantised node pairs (vkr_trace.cuh): the float pairs of a tree quantised with quantise_node_pair() on the grid shadow_grid_from_root() chooses, then the same
// rays through occluded() on the float pairs and occluded_grid() on the quantised ones; grid = minimum xyz, cells per unit xyz. visits[2]: pairs fetched
extern "C" void vkr_device_on_host_trace_quantised(const float* nodes, uint64_t pair_count, const float* tris, const float* grid, uint32_t ray_count, const float* rays, uint8_t* out_float, uint8_t* out_grid, uint32_t* out_pairs8, uint64_t* visits) {
	bvh_view bvh; bvh.nodes = reinterpret_cast<const float4*>(nodes); bvh.tris = reinterpret_cast<const float4*>(tris); bvh.tri_ids = nullptr; bvh.tri_count = 0;
	for (uint64_t i = 0; i != pair_count; ++i) quantise_node_pair(bvh.nodes + 4 * i, grid, grid + 3, out_pairs8 + 8 * i);
	int stack[kMaxStackDepth + 2];
	visits[0] = visits[1] = 0;
	for (uint32_t i = 0; i != ray_count; ++i) {
		const float* r = rays + 8 * (size_t) i;
		const f3 o = make3(r[0], r[1], r[2]), d = make3(r[3], r[4], r[5]);
		int v = 0;
		out_float[i] = occluded(bvh, o, d, r[6], r[7], stack, 1) ? 1 : 0;
		out_grid[i] = occluded_grid(out_pairs8, bvh.tris, grid, grid + 3, o, d, r[6], r[7], stack, 1, &v) ? 1 : 0;
		visits[1] += (uint64_t) v;
	}
}

// Interleaved node pairs (vkr_trace.cuh): the float pairs of a tree re-arranged with interleave_node_pair(), then the same rays through occluded() on the float
// pairs and occluded_interleaved() (the packed-FMA form of the slab test, plain fmaf on the host) on the interleaved ones. visits[2]: pairs fetched by each
extern "C" void vkr_device_on_host_trace_interleaved(const float* nodes, uint64_t pair_count, const float* tris, uint32_t ray_count, const float* rays, uint8_t* out_float, uint8_t* out_interleaved, float* out_pairs16, uint64_t* visits) {
	bvh_view bvh; bvh.nodes = reinterpret_cast<const float4*>(nodes); bvh.tris = reinterpret_cast<const float4*>(tris); bvh.tri_ids = nullptr; bvh.tri_count = 0;
	for (uint64_t i = 0; i != pair_count; ++i) interleave_node_pair(bvh.nodes + 4 * i, out_pairs16 + 16 * i);
	int stack[kMaxStackDepth + 2];
	visits[0] = visits[1] = 0;
	for (uint32_t i = 0; i != ray_count; ++i) {
		const float* r = rays + 8 * (size_t) i;
		const f3 o = make3(r[0], r[1], r[2]), d = make3(r[3], r[4], r[5]);
		int v0 = 0, v1 = 0;
		out_float[i] = occluded_anchored(bvh, o, d, r[6], r[7], nullptr, 0, 0, 0u, stack, 1, &v0) ? 1 : 0;   // no origin path: the plain traversal, counting its visits
		out_interleaved[i] = occluded_interleaved(out_pairs16, bvh.tris, o, d, r[6], r[7], stack, 1, &v1) ? 1 : 0;
		visits[0] += (uint64_t) v0; visits[1] += (uint64_t) v1;
	}
}

// Anchored shadow rays (vkr_anchor.cuh): rays from `origins` towards points of a polygonal light through (a) the plain traversal, (b) the anchored one with
// all siblings of the origin path, (c) with the siblings the light's cone touches; rays = {origin index, dx, dy, dz, tmax}. out[3 * i + {0, 1, 2}] = answers,
// visits[3]: node pairs fetched in total, info = {rays outside their cone, siblings kept, siblings along the paths}
extern "C" void vkr_device_on_host_trace_anchored(const float* nodes, const float* tris, uint32_t origin_count, const float* origins, uint32_t vertex_count, const float* light_vertices_xyzw,
	uint32_t ray_count, const float* rays, uint8_t* out, uint64_t* visits, uint64_t* info)
{
	bvh_view bvh; bvh.nodes = reinterpret_cast<const float4*>(nodes); bvh.tris = reinterpret_cast<const float4*>(tris); bvh.tri_ids = nullptr; bvh.tri_count = 0;
	visits[0] = visits[1] = visits[2] = 0; info[0] = info[1] = info[2] = 0;
	int stack[kMaxStackDepth + 2];
	for (uint32_t i = 0; i != ray_count; ++i) {
		const float* r = rays + 5 * (size_t) i;
		const uint32_t pi = (uint32_t) r[0];
		const f3 o = make3(origins[3 * pi], origins[3 * pi + 1], origins[3 * pi + 2]), d = make3(r[1], r[2], r[3]);
		uint32_t path[kPathLevels]; int tail = 0;
		const int count = find_origin_path(bvh.nodes, o, &tail, [&](int k, uint32_t e) { path[k] = e; });
		const light_cone cone = make_light_cone(o, reinterpret_cast<const unsigned char*>(light_vertices_xyzw), (int) vertex_count);
		uint32_t mask = cull_siblings(bvh.nodes, o, cone, count, [&](int k) { return path[k]; });
		if (!ray_in_cone(cone, d, r[4])) { mask = kAllSiblings; ++info[0]; }
		info[1] += (uint64_t) __builtin_popcount(mask & ((1u << count) - 1u)); info[2] += (uint64_t) count;
		int v0 = 0, v1 = 0, v2 = 0;
		out[3 * i] = occluded_anchored(bvh, o, d, 1.0e-3f, r[4], path, 0, 0, 0u, stack, 1, &v0) ? 1 : 0;   // no path, tail = root: the plain traversal
		out[3 * i + 1] = occluded_anchored(bvh, o, d, 1.0e-3f, r[4], path, count, tail, kAllSiblings, stack, 1, &v1) ? 1 : 0;
		out[3 * i + 2] = occluded_anchored(bvh, o, d, 1.0e-3f, r[4], path, count, tail, mask, stack, 1, &v2) ? 1 : 0;
		if (out[3 * i] != (occluded(bvh, o, d, 1.0e-3f, r[4], stack, 1) ? 1 : 0)) out[3 * i] = 2;   // the counting edition must agree with the probe kernel's traversal
		visits[0] += v0; visits[1] += v1; visits[2] += v2;
	}
}

extern "C" void vkr_device_on_host_decode_positions(const void* constants, const uint32_t* quantized_positions, uint64_t vertex_count, float* out_xyz) {
	const uint2* q = reinterpret_cast<const uint2*>(quantized_positions);
	for (uint64_t i = 0; i != vertex_count; ++i) {
		const f3 v = decode_position(q[i], (const unsigned char*) constants);
		out_xyz[3 * i] = v.x; out_xyz[3 * i + 1] = v.y; out_xyz[3 * i + 2] = v.z;
	}
}

// The body of visibility_kernel (csrc/vkr_gbuffer_kernel.cu): primary ray through the pixel, closest hit (vkr_trace.cuh), ties -> lowest triangle index
extern "C" void vkr_device_on_host_visibility(uint32_t width, uint32_t height, const void* constants, const float* nodes, const float* tris, const uint32_t* tri_ids, uint32_t tri_count, uint32_t* out_visibility) {
	bvh_view bvh; bvh.nodes = reinterpret_cast<const float4*>(nodes); bvh.tris = reinterpret_cast<const float4*>(tris); bvh.tri_ids = tri_ids; bvh.tri_count = tri_count;
	const unsigned char* cb = (const unsigned char*) constants;
	const f3 camera = make3(gldf(cb, G_OFF_CAMERA), gldf(cb, G_OFF_CAMERA + 4), gldf(cb, G_OFF_CAMERA + 8));
	std::vector<int> stack(kMaxStackDepth);
	for (uint32_t y = 0; y != height; ++y) for (uint32_t x = 0; x != width; ++x) {
		const f3 d = pixel_ray(cb, (int) x, (int) y);
		int hit = -1;
		if (tri_count) hit = closest_hit(bvh, camera, d, 0.0f, __int_as_float(0x7f800000), stack.data(), 1);
		out_visibility[(size_t) y * width + x] = (hit < 0) ? 0xFFFFFFFFu : (uint32_t) hit;
	}
}

// The body of the G-buffer kernel (vkr_gbuffer.cuh: shade_gbuffer_pixel) for every pixel, on host arrays laid out like the device buffers.
// texture_dims = uint32[4] per texture, texture_offsets in texels, texture_data = RGBA32F texels; all three null for constant materials.
extern "C" void vkr_device_on_host_gbuffer(uint32_t width, uint32_t height, const void* constants, const uint32_t* visibility, const uint32_t* quantized_positions,
	const uint16_t* normals_and_tex_coords, const uint8_t* material_indices, const float* material_params, const uint32_t* texture_dims, const uint64_t* texture_offsets,
	const float* texture_data, float* out_gbuffer)
{
	gbuffer_kernel_params p;
	memset(&p, 0, sizeof(p));
	p.width = (int) width; p.height = (int) height; p.constants = (const unsigned char*) constants;
	p.quantized_positions = reinterpret_cast<const uint2*>(quantized_positions); p.normals_and_tex_coords = reinterpret_cast<const ushort4*>(normals_and_tex_coords);
	p.material_indices = material_indices; p.material_params = material_params;
	p.visibility = const_cast<uint32_t*>(visibility); p.gbuffer = reinterpret_cast<float4*>(out_gbuffer);
	p.texture_data = reinterpret_cast<const float4*>(texture_data); p.texture_dims = reinterpret_cast<const uint4*>(texture_dims);
	p.texture_offsets = reinterpret_cast<const unsigned long long*>(texture_offsets);
	for (size_t pixel = 0; pixel != (size_t) width * height; ++pixel) {
		if (texture_data) shade_gbuffer_pixel<true>(p, pixel);
		else shade_gbuffer_pixel<false>(p, pixel);
	}
}

// One frame of the error display modes on the CPU: per pixel the prologue and epilogue of the shading tile (csrc/vkr_shading_tile.cuh: G-buffer read, light
// display, LTC set-up, noise stream, NaN -> pink, exposure; restated here because that file is warp-level code) around the product's per-light function
// error_display_of_light() (csrc/vkr_error_display.cuh). Linear output, g_frame_bits = 0.
// light(sp, l, light_block, ns, x, y, acc): one light for one pixel; acc is the pixel's pixel_sum (vkr_ray_stream.cuh)
template <int MAXV, bool LIGHT_TEXTURES = false, class Light>
static void tile_frame(const shading_kernel_params& p, int show_lights, const Light& light_fn, float* out_rgba) {
	const unsigned char* cb = p.constants;
	const size_t plane = (size_t) p.width * p.height;
	const int light_stride = L_FIXED + 16 * MAXV * 2 + 16 * (MAXV - 2);
	const f3 camera = make3(ldf(cb, OFF_CAMERA), ldf(cb, OFF_CAMERA + 4), ldf(cb, OFF_CAMERA + 8));
	const float exposure = ldf(cb, OFF_EXPOSURE);
	for (int y = 0; y != p.height; ++y) for (int x = 0; x != p.width; ++x) {
		const size_t pixel = (size_t) y * p.width + x;
		const float4 g0 = p.gbuffer[pixel], g1 = p.gbuffer[plane + pixel];
		const bool valid = g1.w != 0.0f;
		f3 color = make3(0.0f, 0.0f, 0.0f);
		shading_point sp;
		sp.position = make3(g0.x, g0.y, g0.z); sp.roughness = g0.w; sp.normal = make3(g1.x, g1.y, g1.z);
		if (show_lights) {
			f3 end = sp.position; float end_w = 1.0f;
			const float fx = (float) x, fy = (float) y;
			f3 view_direction = make3(
				fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 8), 1.0f, fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 4), fy, ldf(cb, OFF_PIXEL_TO_RAY) * fx)),
				fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 24), 1.0f, fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 20), fy, ldf(cb, OFF_PIXEL_TO_RAY + 16) * fx)),
				fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 40), 1.0f, fmaf(ldf(cb, OFF_PIXEL_TO_RAY + 36), fy, ldf(cb, OFF_PIXEL_TO_RAY + 32) * fx)));
			if (!valid) { end = view_direction; end_w = 0.0f; }
			view_direction = normalize(view_direction);
			for (int li = 0; li != p.light_count; ++li) {
				const unsigned char* light = cb + CONSTANTS_FIXED + li * light_stride;
				if (light_ray_intersection<MAXV>(light, camera, end, end_w))
					color = color + light_radiance<LIGHT_TEXTURES>(p, light, camera, view_direction);
			}
		}
		pixel_sum acc;
		acc.color = color; acc.light = make3(0.0f, 0.0f, 0.0f); acc.inv_samples = 1.0f / (float) p.sample_count;
		acc.submit_parity = 0u; acc.resolve_parity = 0u; acc.pushed = false;
		if (valid) {
			const float4 g2 = p.gbuffer[2 * plane + pixel], g3 = p.gbuffer[3 * plane + pixel];
			sp.diffuse_albedo = make3(g2.x, g2.y, g2.z); sp.fresnel_0 = make3(g3.x, g3.y, g3.z);
			sp.outgoing = normalize(camera - sp.position);
			sp.lambert_outgoing = dot(sp.normal, sp.outgoing);
			ltc_state l = {};
			get_ltc_coefficients(l, p, cb, sp);
			noise_stream ns;
			ns.z = 0.0f; ns.w = 0.0f; ns.available = 0; ns.sample_index = 0;
			for (int li = 0; li != p.light_count; ++li) light_fn(sp, l, cb + CONSTANTS_FIXED + li * light_stride, ns, (uint32_t) x, (uint32_t) y, acc);
		}
		color = acc.color;
		f3 final_color = color;
		if (std::isnan(color.x) || std::isnan(color.y) || std::isnan(color.z) || std::isinf(color.x) || std::isinf(color.y) || std::isinf(color.z))
			final_color = make3(1.0f / exposure, 0.0f / exposure, 0.8f / exposure);
		float* o = out_rgba + 4 * pixel;
		const f3 out_color = output_stage(make3(final_color.x * exposure, final_color.y * exposure, final_color.z * exposure), ldu(cb, OFF_FRAME_BITS), p.output_srgb != 0);
		o[0] = out_color.x; o[1] = out_color.y; o[2] = out_color.z; o[3] = 1.0f;
	}
}

template <int MAXV>
static void error_display_frame(const shading_kernel_params& p, int show_lights, float* out_rgba) {
	tile_frame<MAXV>(p, show_lights, [&](const shading_point& sp, const ltc_state& l, const unsigned char* light, noise_stream& ns, uint32_t x, uint32_t y, pixel_sum& acc) {
		f3 contribution;
		if (error_display_of_light<MAXV>(&contribution, sp, l, light, ns, p, p.constants, x, y)) acc.color = acc.color + contribution;
	}, out_rgba);
}

// The shading pass without shadow rays (TRACE = false: every candidate sample is added in place) with the product's shade_light() (csrc/vkr_shade_light.cuh)
template <int STRATEGY, int MAXV, bool BIASED, bool OPTIMAL, bool LIGHT_TEXTURES>
static void shading_frame(const shading_kernel_params& p, int show_lights, float* out_rgba) {
	tile_frame<MAXV, LIGHT_TEXTURES>(p, show_lights, [&](const shading_point& sp, const ltc_state& l, const unsigned char* light, noise_stream& ns, uint32_t x, uint32_t y, pixel_sum& acc) {
		ray_producer q; q.base = 0; q.fill = 0; q.resolved = 0;
		shade_light<STRATEGY, MAXV + 1, BIASED, OPTIMAL, false, LIGHT_TEXTURES>(true, sp, l, light, ns, p, p.constants, x, y, q, acc, 0);
	}, out_rgba);
}

template <int MAXV, bool BIASED, bool T>
static int shading_frame_strategy(const shading_kernel_params& p, int show_lights, float* out_rgba) {
	switch (p.sampling_strategies) {
	case VKR_STRATEGY_DIFFUSE_ONLY: shading_frame<VKR_STRATEGY_DIFFUSE_ONLY, MAXV, BIASED, false, T>(p, show_lights, out_rgba); return 0;
	case VKR_STRATEGY_DIFFUSE_GGX_MIS: shading_frame<VKR_STRATEGY_DIFFUSE_GGX_MIS, MAXV, BIASED, false, T>(p, show_lights, out_rgba); return 0;
	case VKR_STRATEGY_DIFFUSE_SPECULAR_SEPARATELY: shading_frame<VKR_STRATEGY_DIFFUSE_SPECULAR_SEPARATELY, MAXV, BIASED, false, T>(p, show_lights, out_rgba); return 0;
	case VKR_STRATEGY_DIFFUSE_SPECULAR_MIS:
		if (p.mis_heuristic == VKR_MIS_OPTIMAL) shading_frame<VKR_STRATEGY_DIFFUSE_SPECULAR_MIS, MAXV, BIASED, true, T>(p, show_lights, out_rgba);
		else shading_frame<VKR_STRATEGY_DIFFUSE_SPECULAR_MIS, MAXV, BIASED, false, T>(p, show_lights, out_rgba);
		return 0;
	case VKR_STRATEGY_DIFFUSE_SPECULAR_RANDOM: shading_frame<VKR_STRATEGY_DIFFUSE_SPECULAR_RANDOM, MAXV, BIASED, false, T>(p, show_lights, out_rgba); return 0;
	default: return 1;
	}
}

// The same with the related-work techniques (csrc/vkr_related_work_light.cuh), technique = p.polygon_sampling_technique (0..10)
template <int STRATEGY, int MAXV, bool LIGHT_TEXTURES>
static void related_work_frame(const shading_kernel_params& p, int show_lights, float* out_rgba) {
	tile_frame<MAXV, LIGHT_TEXTURES>(p, show_lights, [&](const shading_point& sp, const ltc_state& l, const unsigned char* light, noise_stream& ns, uint32_t x, uint32_t y, pixel_sum& acc) {
		ray_producer q; q.base = 0; q.fill = 0; q.resolved = 0;
		related_work_light_shader<STRATEGY, MAXV, false, LIGHT_TEXTURES>()(true, sp, l, light, ns, p, p.constants, x, y, q, acc, 0);
	}, out_rgba);
}

// One frame of the shading pass WITHOUT shadow rays on the CPU (light vertex bounds 3 to 7). Output stage as the constant block (g_frame_bits) and output_srgb say. technique: sample_polygon_technique_t, 11 = projected solid angle (biased: 12),
// 0..10 = related work. light_texture_count != 0: the LIGHT_TEXTURES = true instantiation (csrc/vkr_textured_light_kernel.cu), dims = {width, height, mip count, 0} per texture, offsets in texels.
extern "C" int vkr_device_on_host_shade_frame(uint32_t width, uint32_t height, uint32_t maxv, uint32_t light_count, uint32_t technique, uint32_t strategy, uint32_t heuristic, int biased, uint32_t sample_count, int show_lights,
	const void* constants, const float* gbuffer, const uint16_t* noise, uint32_t noise_w, uint32_t noise_h, uint32_t noise_layers,
	const uint16_t* ltc0, const uint16_t* ltc1, uint32_t ltc_res, uint32_t ltc_layers, int output_srgb,
	uint32_t light_texture_count, const uint32_t* light_texture_dims4, const uint64_t* light_texture_offsets_texels, const float* light_texture_texels, float* out_rgba)
{
	shading_kernel_params p;
	memset(&p, 0, sizeof(p));
	p.width = (int) width; p.height = (int) height; p.gbuffer = reinterpret_cast<const float4*>(gbuffer); p.constants = (const unsigned char*) constants;
	p.light_count = (int) light_count; p.max_light_vertex_count = (int) maxv; p.sample_count = (int) sample_count;
	p.light_texture_count = light_texture_count; p.light_texture_dims = reinterpret_cast<const uint4*>(light_texture_dims4);
	p.light_texture_offsets = reinterpret_cast<const unsigned long long*>(light_texture_offsets_texels); p.light_texture_texels = reinterpret_cast<const float4*>(light_texture_texels);
	p.sampling_strategies = (int) strategy; p.mis_heuristic = (int) heuristic; p.biased_sampling = biased; p.polygon_sampling_technique = biased ? 12 : 11; p.output_srgb = output_srgb;
	p.noise = noise; p.noise_w = (int) noise_w; p.noise_h = (int) noise_h; p.noise_layers = (int) noise_layers;
	p.ltc0 = ltc0; p.ltc1 = ltc1; p.ltc_res = (int) ltc_res; p.ltc_layers = (int) ltc_layers;
	if (technique < 11) { // related work: diffuse only or GGX MIS
		p.polygon_sampling_technique = (int) technique;
		if (biased || strategy > 1) return 1;
		switch (maxv) {
#define V(K) case K: if (light_texture_count) { if (strategy) related_work_frame<1, K, true>(p, show_lights, out_rgba); else related_work_frame<0, K, true>(p, show_lights, out_rgba); } \
	else { if (strategy) related_work_frame<1, K, false>(p, show_lights, out_rgba); else related_work_frame<0, K, false>(p, show_lights, out_rgba); } return 0;
		V(3) V(4) V(5) V(6) V(7)
#undef V
		default: return 1;
		}
	}
	switch (maxv) {
#define V(K) case K: if (light_texture_count) return biased ? shading_frame_strategy<K, true, true>(p, show_lights, out_rgba) : shading_frame_strategy<K, false, true>(p, show_lights, out_rgba); \
	return biased ? shading_frame_strategy<K, true, false>(p, show_lights, out_rgba) : shading_frame_strategy<K, false, false>(p, show_lights, out_rgba);
	V(3) V(4) V(5) V(6) V(7)
#undef V
	default: return 1;
	}
}

extern "C" int vkr_device_on_host_error_display_frame(uint32_t width, uint32_t height, uint32_t maxv, uint32_t light_count, uint32_t technique, uint32_t error_display, int show_lights,
	const void* constants, const float* gbuffer, const uint16_t* noise, uint32_t noise_w, uint32_t noise_h, uint32_t noise_layers,
	const uint16_t* ltc0, const uint16_t* ltc1, uint32_t ltc_res, uint32_t ltc_layers, float* out_rgba)
{
	shading_kernel_params p;
	memset(&p, 0, sizeof(p));
	p.width = (int) width; p.height = (int) height; p.gbuffer = reinterpret_cast<const float4*>(gbuffer); p.constants = (const unsigned char*) constants;
	p.light_count = (int) light_count; p.max_light_vertex_count = (int) maxv; p.sample_count = 1;
	p.polygon_sampling_technique = (int) (technique == 12 ? 11 : technique); p.biased_sampling = technique == 12; p.error_display = (int) error_display;
	p.noise = noise; p.noise_w = (int) noise_w; p.noise_h = (int) noise_h; p.noise_layers = (int) noise_layers;
	p.ltc0 = ltc0; p.ltc1 = ltc1; p.ltc_res = (int) ltc_res; p.ltc_layers = (int) ltc_layers;
	switch (maxv) {
	case 3: error_display_frame<3>(p, show_lights, out_rgba); return 0;
	case 4: error_display_frame<4>(p, show_lights, out_rgba); return 0;
	case 5: error_display_frame<5>(p, show_lights, out_rgba); return 0;
	case 6: error_display_frame<6>(p, show_lights, out_rgba); return 0;
	case 7: error_display_frame<7>(p, show_lights, out_rgba); return 0;
	default: return 1;
	}
}

// The GPU linear-BVH builder (csrc/vkr_lbvh_gpu.cu) with every kernel launch replaced by a loop over its elements and the two library calls by their
// definitions (cub::DeviceRadixSort::SortPairs = stable sort by key, cub::DeviceScan::ExclusiveSum = running sum): the per-element device functions of
// csrc/vkr_lbvh.cuh do all the work. Output as the host probes deliver it; must equal builder 1 (csrc/vkr_lbvh.cpp) array for array.
extern "C" int vkr_device_on_host_lbvh(const float* vertices, uint64_t triangle_count, float* out_nodes, uint64_t* out_node_count, float* out_tris, uint32_t* out_tri_ids, uint32_t* out_max_depth) {
	if (triangle_count <= (uint64_t) kLbvhLeafSize) return 1;
	const uint32_t n = (uint32_t) triangle_count, internal_count = n - 1;
	std::vector<float> box_lo(3 * (size_t) n), box_hi(3 * (size_t) n), centroid(3 * (size_t) n);
	uint32_t keys[12]; // scene lo/hi, centroid lo/hi as ordered keys, reduced with min / max like the atomics do
	for (int a = 0; a != 3; ++a) { keys[a] = keys[6 + a] = 0xffffffffu; keys[3 + a] = keys[9 + a] = 0u; }
	for (uint32_t t = 0; t != n; ++t) {
		lbvh_triangle_bounds(vertices, t, &box_lo[3 * (size_t) t], &box_hi[3 * (size_t) t], &centroid[3 * (size_t) t]);
		for (int a = 0; a != 3; ++a) {
			keys[a] = std::min(keys[a], float_to_ordered(box_lo[3 * (size_t) t + a])); keys[3 + a] = std::max(keys[3 + a], float_to_ordered(box_hi[3 * (size_t) t + a]));
			keys[6 + a] = std::min(keys[6 + a], float_to_ordered(centroid[3 * (size_t) t + a])); keys[9 + a] = std::max(keys[9 + a], float_to_ordered(centroid[3 * (size_t) t + a]));
		}
	}
	float extent = 0.0f, lo[3], inv[3];
	for (int a = 0; a != 3; ++a) {
		extent = fmaxf(extent, fmaxf(fabsf(ordered_to_float(keys[a])), fabsf(ordered_to_float(keys[3 + a]))));
		lo[a] = ordered_to_float(keys[6 + a]);
		const float e = ordered_to_float(keys[9 + a]) - lo[a];
		inv[a] = (e > 0.0f) ? 1.0f / e : 0.0f;
	}
	const float pad = extent * (1.0f / 65536.0f);
	std::vector<uint64_t> codes_in(n), codes(n); std::vector<uint32_t> order(n);
	for (uint32_t t = 0; t != n; ++t) { codes_in[t] = lbvh_morton(&centroid[3 * (size_t) t], make3(lo[0], lo[1], lo[2]), make3(inv[0], inv[1], inv[2])); order[t] = t; }
	std::stable_sort(order.begin(), order.end(), [&](uint32_t l, uint32_t r) { return codes_in[l] < codes_in[r]; });
	for (uint32_t s2 = 0; s2 != n; ++s2) codes[s2] = codes_in[order[s2]];
	for (uint32_t s2 = 0; s2 != n; ++s2) { lbvh_slot(vertices, order[s2], out_tris + 12 * (size_t) s2); out_tri_ids[s2] = order[s2]; }
	std::vector<int32_t> first(internal_count), last(internal_count), split(internal_count), parent(internal_count, -1), leaf_parent(n, -1);
	for (int64_t i = 0; i != (int64_t) internal_count; ++i) lbvh_radix_tree_node(codes.data(), n, i, first.data(), last.data(), split.data(), parent.data(), leaf_parent.data());
	std::vector<float> node_lo(3 * (size_t) internal_count), node_hi(3 * (size_t) internal_count); std::vector<uint32_t> arrivals(internal_count, 0u), used(internal_count), rank(internal_count);
	for (uint32_t s2 = 0; s2 != n; ++s2) lbvh_refit_from_leaf(s2, order.data(), box_lo.data(), box_hi.data(), first.data(), last.data(), split.data(), parent.data(), leaf_parent.data(), node_lo.data(), node_hi.data(), arrivals.data());
	for (uint32_t i = 0; i != internal_count; ++i) if (arrivals[i] != 2u) return 2; // every internal node is reached by both subtrees
	uint32_t used_count = 0;
	for (uint32_t i = 0; i != internal_count; ++i) { used[i] = lbvh_is_used(first[i], last[i]); rank[i] = used_count; used_count += used[i]; }
	uint32_t depth = 0;
	for (uint32_t i = 0; i != internal_count; ++i) {
		if (!used[i]) continue;
		lbvh_emit_pair(i, order.data(), box_lo.data(), box_hi.data(), node_lo.data(), node_hi.data(), first.data(), last.data(), split.data(), used.data(), rank.data(), pad, out_nodes + 16 * (size_t) rank[i]);
		depth = std::max(depth, lbvh_depth(i, parent.data()));
	}
	*out_node_count = used_count; *out_max_depth = depth;
	return 0;
}

// Elementary functions of the device arithmetic contract: 0 atan, 1 sin, 2 cos, 3 acos on [-1,1], 4 atan2(x, 0.5), 5 pow(x, 1/3), 6 fast_positive_atan
extern "C" void vkr_device_on_host_elementary_batch(int which, uint32_t n, const float* x, float* y) {
	for (uint32_t i = 0; i != n; ++i) {
		switch (which) {
		case 0: y[i] = atan_poly(x[i]); break;
		case 1: y[i] = sin_cw(x[i]); break;
		case 2: y[i] = cos_cw(x[i]); break;
		case 3: y[i] = acos_full(x[i]); break;
		case 4: y[i] = atan2_poly(x[i], 0.5f) + atan2_poly(0.5f, x[i]); break;
		case 5: y[i] = pow_contract(x[i], 1.0f / 3.0f); break;
		default: y[i] = fast_positive_atan(x[i]); break;
		}
	}
}
