// vkr_related_work_light.cuh -- one polygonal light for one pixel with a related-work sampling technique.
//
// The per-(pixel, light) code of related_work_kernel (vkr_related_work_kernel.cu) and of its variant for textured lights
// (vkr_textured_related_work_kernel.cu): get_polygonal_light_mis_estimate() of the shader (shading_pass.frag.glsl:305-323, 332-437, 676-709).
// With TRACE = false there is no warp-level operation in here, so tests/device_on_host.cpp runs it on the CPU.
#pragma once
#include "vkr_shade_common.cuh"
#include "vkr_ray_stream.cuh"
#include "vkr_related_work.cuh"

namespace vkr {

// One polygonal light for the warp's 32 pixels with sampling technique TECHNIQUE. The estimator is
// get_polygonal_light_mis_estimate() (shading_pass.frag.glsl:305-323): f * cos / p with SAMPLING_STRATEGIES_DIFFUSE_ONLY,
// MIS against GGX importance sampling (:676-709) with SAMPLING_STRATEGIES_DIFFUSE_GGX_MIS.
template <int TECHNIQUE, int STRATEGY, int MAXV, bool TRACE, bool LIGHT_TEXTURES = false>
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
				const f3 rtb = light_radiance<LIGHT_TEXTURES>(p, light, sp.position, w) * evaluate_brdf<true, true>(sp, w);
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
					const float wgt = (p.mis_heuristic == VKR_MIS_BALANCE) ? (1.0f / (ggx_density + polygon_density)) : (ggx_density / (ggx_density * ggx_density + polygon_density * polygon_density));
					if (!is_finite(lambert) || !is_finite(wgt)) { has = true; tmax = 0.0f; w = zero; c = make3(not_a_number, not_a_number, not_a_number); } // 0 * inf: NaN whether visible or not
					else if (lambert > 0.0f) {
						has = true;
						tmax = light_plane_distance(sp, light, w);
						const f3 rtb = light_radiance<LIGHT_TEXTURES>(p, light, sp.position, w) * evaluate_brdf<true, true>(sp, w);
						c = make3(rtb.x * lambert * wgt, rtb.y * lambert * wgt, rtb.z * lambert * wgt);
					}
				}
			}
			submit<TRACE, false>(q, lane, has, true, w, tmax, c, zero, result, false);
		}
	}
	submit<TRACE, false>(q, lane, false, false, zero, 0.0f, zero, zero, result, true);
}

template <int STRATEGY, int MAXV, bool TRACE, bool LIGHT_TEXTURES = false>
struct related_work_light_shader {
	VKR_DEV void operator()(bool on, const shading_point& sp, const ltc_state& l, const unsigned char* light, noise_stream& ns,
		const shading_kernel_params& p, const unsigned char* cb, uint32_t px, uint32_t py, ray_producer& q, pixel_sum& result, int lane) const
	{
		switch (p.polygon_sampling_technique) { // warp-uniform (kernel-uniform)
#define VKR_CASE(T) case T: shade_light_related_work<T, STRATEGY, MAXV, TRACE, LIGHT_TEXTURES>(on, sp, l, light, ns, p, cb, px, py, q, result, lane); break;
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

} // namespace vkr
