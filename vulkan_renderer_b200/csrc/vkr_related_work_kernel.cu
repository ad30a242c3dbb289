// vkr_related_work_kernel.cu -- the shading megakernel with the related-work polygon sampling techniques (sm_100a).
//
// the samplers the reference compares projected solid angle sampling against (shading_pass.frag.glsl:332-481,
// polygon_sampling_related_work.glsl): baseline, Turk, Urena, Arvo (solid angle / projected solid angle), solid angle with and
// without clipping, Hart et al.'s bilinear and biquadratic cosine warps. The tile body, the warp-specialised ray streams and the
// output stage are those of the main kernel (vkr_shading_tile.cuh); only the light shader differs. Same launch shape, same
// arithmetic contract (-fmad=false), results bit-identical to the reference shader fixtures ("_q<technique>").
// The error display modes of the shader (colour-coded error of the sampling procedure) live here too: error_display_kernel.
// The technique is a warp-uniform run-time switch inside one kernel per (strategy, vertex bound, shadow rays): these are
// comparison baselines, not the benchmarked path, so code size matters more than the last register.
#include "vkr_shading_tile.cuh"
#include "vkr_related_work.cuh"
#include "vkr_error_display.cuh"

namespace vkr {

// One polygonal light for the warp's 32 pixels with sampling technique TECHNIQUE. The estimator is
// get_polygonal_light_mis_estimate() (shading_pass.frag.glsl:305-323): f * cos / p with SAMPLING_STRATEGIES_DIFFUSE_ONLY,
// MIS against GGX importance sampling (:676-709) with SAMPLING_STRATEGIES_DIFFUSE_GGX_MIS.
template <int TECHNIQUE, int STRATEGY, int MAXV, bool TRACE>
VKR_DEV void shade_light_related_work(bool on, const shading_point& sp, const ltc_state& l, const unsigned char* light, noise_stream& ns,
	const shading_kernel_params& p, const unsigned char* cb, uint32_t px, uint32_t py, ray_producer& q, pixel_sum& result, int lane)
{
	const int S = p.sample_count;
	const f3 zero = make3(0.0f, 0.0f, 0.0f);
	const float not_a_number = __int_as_float(0x7fc00000);
	rw_sampler<TECHNIQUE, MAXV> sampler;
	if (on) {
		rw_light<MAXV> view;
		rw_load_light<MAXV>(view, light);
		rw_frame frame;
		frame.rx = l.rx; frame.ry = l.ry; frame.rz = sp.normal; frame.t = l.t;
		on = sampler.prepare(view, sp.position, frame);
	}
#pragma unroll 1
	for (int s = 0; s != S; ++s) {
		bool has = false, pre_vis = false; f3 w = zero, c = zero; float tmax = 0.0f;
		if (on) {
			float density;
			w = sampler.sample(next_noise_2(ns, p, cb, px, py), &density);
			const float lambert = dot(sp.normal, w);
			pre_vis = lambert > 0.0f;
			// factor = what the shader multiplies radiance * BRDF (zero if the light is not visible) with. If it is not finite,
			// the product is NaN whether or not the light is visible and the pixel turns pink (:862-864)
			float factor;
			if (STRATEGY == VKR_STRATEGY_DIFFUSE_ONLY) {
				has = density > 0.0f;
				factor = lambert / density;
			}
			else {
				has = true;
				const float ggx_density = ggx_reflected_direction_density(sp.lambert_outgoing, sp.outgoing, w, sp.normal, sp.roughness);
				factor = (p.mis_heuristic == VKR_MIS_BALANCE) ? (1.0f / (density + ggx_density)) : (density / (density * density + ggx_density * ggx_density));
			}
			const bool poisoned = has && (!(fabsf(factor) <= 3.402823466e+38f) || (STRATEGY != VKR_STRATEGY_DIFFUSE_ONLY && !(fabsf(lambert) <= 3.402823466e+38f)));
			if (poisoned) { // tmax = 0 is a miss by definition (vkr_ray_stream.cuh), so NaN is added without a traversal
				pre_vis = true; tmax = 0.0f; w = zero;
				c = make3(not_a_number, not_a_number, not_a_number);
			}
			else if (has && pre_vis) {
				tmax = light_plane_distance(sp, light, w);
				const f3 rtb = light_radiance(light) * evaluate_brdf<true, true>(sp, w);
				if (STRATEGY == VKR_STRATEGY_DIFFUSE_ONLY) c = rtb * factor;
				else c = make3(rtb.x * lambert * factor, rtb.y * lambert * factor, rtb.z * lambert * factor);
			}
		}
		submit<TRACE, false>(q, lane, has, pre_vis, w, tmax, c, zero, result, false);
	}
	if (STRATEGY == VKR_STRATEGY_DIFFUSE_GGX_MIS) { // :676-709
		bool flip = false;
		if constexpr (TECHNIQUE == VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO) flip = on && sampler.flip;
		const f3 o_ss = make3(
			fmaf(l.t.x, 0.0f, fmaf(l.rx.z, sp.outgoing.z, fmaf(l.rx.y, sp.outgoing.y, l.rx.x * sp.outgoing.x))),
			0.0f,
			fmaf(l.t.z, 0.0f, fmaf(sp.normal.z, sp.outgoing.z, fmaf(sp.normal.y, sp.outgoing.y, sp.normal.x * sp.outgoing.x))));
		const float polygon_density = on ? sampler.ggx_density_factor() : 0.0f; // every technique but "ours" uses the factor as is (:702)
#pragma unroll 1
		for (int s = 0; s != S; ++s) {
			bool has = false; f3 w = zero, c = zero; float tmax = 0.0f;
			if (on) {
				float ggx_density;
				const f3 d = sample_ggx_reflected_direction(&ggx_density, o_ss, sp.roughness, next_noise_2(ns, p, cb, px, py));
				w = shading_to_world(l, sp.normal, flip, d);
				if (d.z > 0.0f && light_ray_intersection<MAXV>(light, sp.position, w, 0.0f)) {
					const float lambert = dot(sp.normal, w);
					if (lambert > 0.0f) {
						has = true;
						tmax = light_plane_distance(sp, light, w);
						const f3 rtb = light_radiance(light) * evaluate_brdf<true, true>(sp, w);
						const float wgt = (p.mis_heuristic == VKR_MIS_BALANCE) ? (1.0f / (ggx_density + polygon_density)) : (ggx_density / (ggx_density * ggx_density + polygon_density * polygon_density));
						c = make3(rtb.x * lambert * wgt, rtb.y * lambert * wgt, rtb.z * lambert * wgt);
					}
				}
			}
			submit<TRACE, false>(q, lane, has, true, w, tmax, c, zero, result, false);
		}
	}
	submit<TRACE, false>(q, lane, false, false, zero, 0.0f, zero, zero, result, true);
}

template <int STRATEGY, int MAXV, bool TRACE>
struct related_work_light_shader {
	VKR_DEV void operator()(bool on, const shading_point& sp, const ltc_state& l, const unsigned char* light, noise_stream& ns,
		const shading_kernel_params& p, const unsigned char* cb, uint32_t px, uint32_t py, ray_producer& q, pixel_sum& result, int lane) const
	{
		switch (p.polygon_sampling_technique) { // warp-uniform (kernel-uniform)
#define VKR_CASE(T) case T: shade_light_related_work<T, STRATEGY, MAXV, TRACE>(on, sp, l, light, ns, p, cb, px, py, q, result, lane); break;
		VKR_CASE(VKR_TECHNIQUE_RECTANGLE_SOLID_ANGLE_URENA) VKR_CASE(VKR_TECHNIQUE_SOLID_ANGLE_ARVO) VKR_CASE(VKR_TECHNIQUE_SOLID_ANGLE)
		VKR_CASE(VKR_TECHNIQUE_CLIPPED_SOLID_ANGLE) VKR_CASE(VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO)
		default:
			if constexpr (STRATEGY == VKR_STRATEGY_DIFFUSE_ONLY) { // the techniques without a stand-alone density (user_interface.cpp:133-140)
				switch (p.polygon_sampling_technique) {
				VKR_CASE(VKR_TECHNIQUE_BASELINE) VKR_CASE(VKR_TECHNIQUE_AREA_TURK)
				VKR_CASE(VKR_TECHNIQUE_BILINEAR_COSINE_WARP_HART) VKR_CASE(VKR_TECHNIQUE_BILINEAR_COSINE_WARP_CLIPPING_HART)
				VKR_CASE(VKR_TECHNIQUE_BIQUADRATIC_COSINE_WARP_HART) VKR_CASE(VKR_TECHNIQUE_BIQUADRATIC_COSINE_WARP_CLIPPING_HART)
				default: break;
				}
			}
			break;
#undef VKR_CASE
		}
	}
};

// Light shader of the error display modes (vkr_error_display.cuh)
template <int MAXV>
struct error_display_light_shader {
	VKR_DEV void operator()(bool on, const shading_point& sp, const ltc_state& l, const unsigned char* light, noise_stream& ns,
		const shading_kernel_params& p, const unsigned char* cb, uint32_t px, uint32_t py, ray_producer&, pixel_sum& result, int) const
	{
		if (!on) return; // no warp-level operations below: lanes are independent
		f3 color;
		if (error_display_of_light<MAXV>(&color, sp, l, light, ns, p, cb, px, py)) result.color = result.color + color;
	}
};

template <int MAXV>
__global__ void __launch_bounds__(kShadeThreads, 3)
error_display_kernel(const shading_kernel_params p) {
	shade_tile<MAXV + 1, false, false>(p, error_display_light_shader<MAXV>());
}

template <int STRATEGY, int MAXV, bool TRACE>
__global__ void __launch_bounds__(TRACE ? kShadeThreads + kTraceThreads : kShadeThreads, TRACE ? 2 : 3)
related_work_kernel(const shading_kernel_params p) {
	shade_tile<MAXV + 1, false, TRACE>(p, related_work_light_shader<STRATEGY, MAXV, TRACE>());
}

} // namespace vkr

using namespace vkr;

static constexpr size_t kStreamFloats = stream_floats_per_warp(false);

template <int STRATEGY, int MAXV, bool TRACE>
static cudaError_t launch_related_work(const shading_kernel_params& p, cudaStream_t stream) {
	const int tiles_x = (p.width + kTileW - 1) / kTileW;
	const int tiles_y = p.tile_row_count;
	if (tiles_x <= 0 || tiles_y <= 0) return cudaSuccess;
	const int threads = TRACE ? kShadeThreads + kTraceThreads : kShadeThreads;
	const size_t smem = p.constants_smem_bytes + (TRACE ? sizeof(float) * kStreamFloats * kShadeWarps + sizeof(int) * (size_t) p.stack_depth * kTraceThreads : 0);
	auto kernel = related_work_kernel<STRATEGY, MAXV, TRACE>;
	cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if (err != cudaSuccess) return err;
	int ctas = 0;
	err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, threads, smem);
	if (err != cudaSuccess) return err;
	const int carveout = (int) ((100 * ((smem + 1024) * (size_t) (ctas > 0 ? ctas : 1)) + 228 * 1024 - 1) / (228 * 1024));
	err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout > 100 ? 100 : carveout);
	if (err != cudaSuccess) return err;
	kernel<<<tiles_x * tiles_y, threads, smem, stream>>>(p);
	return cudaGetLastError();
}

template <int MAXV>
static cudaError_t launch_error_display(const shading_kernel_params& p, cudaStream_t stream) {
	const int tiles_x = (p.width + kTileW - 1) / kTileW;
	const int tiles_y = p.tile_row_count;
	if (tiles_x <= 0 || tiles_y <= 0) return cudaSuccess;
	auto kernel = error_display_kernel<MAXV>;
	cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) p.constants_smem_bytes);
	if (err != cudaSuccess) return err;
	kernel<<<tiles_x * tiles_y, kShadeThreads, p.constants_smem_bytes, stream>>>(p);
	return cudaGetLastError();
}

template <int STRATEGY, int MAXV>
static cudaError_t launch_related_work_traced(const shading_kernel_params& p, cudaStream_t stream) {
	if (p.trace_shadow_rays != 0 && p.tri_count != 0) return launch_related_work<STRATEGY, MAXV, true>(p, stream);
	return launch_related_work<STRATEGY, MAXV, false>(p, stream);
}

// One translation unit per light vertex bound (built with -DVKR_MAXV_TU=3 .. 7, __graft_entry__.py)
#ifndef VKR_MAXV_TU
#error "compile with -DVKR_MAXV_TU=<3..7>"
#endif
#define VKR_CONCAT2(a, b) a##b
#define VKR_CONCAT(a, b) VKR_CONCAT2(a, b)
cudaError_t VKR_CONCAT(vkr_launch_related_work_kernel_maxv, VKR_MAXV_TU)(const shading_kernel_params& p, cudaStream_t stream) {
	if (p.error_display != 0) { // the error display modes of projected solid angle sampling (techniques 10, 11, 12)
		if (p.error_display < 1 || p.error_display > 6 || p.polygon_sampling_technique < VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO) return cudaErrorInvalidValue;
		return launch_error_display<VKR_MAXV_TU>(p, stream);
	}
	if (p.stack_depth < 2 || p.stack_depth > kMaxStackDepth) return cudaErrorInvalidValue;
	if (p.trace_shadow_rays != 0 && p.tri_count != 0 && p.bvh_width != VKR_BVH_WIDTH) return cudaErrorInvalidValue;
	if (p.polygon_sampling_technique < VKR_TECHNIQUE_BASELINE || p.polygon_sampling_technique > VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO) return cudaErrorInvalidValue;
	switch (p.sampling_strategies) {
	case VKR_STRATEGY_DIFFUSE_ONLY: return launch_related_work_traced<VKR_STRATEGY_DIFFUSE_ONLY, VKR_MAXV_TU>(p, stream);
	case VKR_STRATEGY_DIFFUSE_GGX_MIS: return launch_related_work_traced<VKR_STRATEGY_DIFFUSE_GGX_MIS, VKR_MAXV_TU>(p, stream);
	default: return cudaErrorInvalidValue;
	}
}
