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

#include "vkr_related_work.cuh"
#include "vkr_trace.cuh"
#include "vkr_texture.cuh"
#include "vkr_gbuffer.cuh"

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
