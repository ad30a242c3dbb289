// vkr_error_display.cuh -- the error display modes of the shader for one (pixel, light) (ERROR_DISPLAY_DIFFUSE / ERROR_DISPLAY_SPECULAR,
// shading_pass.frag.glsl:462-493, 549-563): instead of shading, a light contributes the colour-coded error of ONE sample of projected solid angle
// sampling (ours: backward, scaled backward or forward error; Arvo's: backward or scaled backward error). No shadow rays, one noise fetch per light.
// Per-thread code without warp-level operations: vkr_related_work_kernel.cu calls it from its light shader, tests/device_on_host.cpp runs it on the
// CPU against the reference-shader fixtures. Compile with -fmad=false.
#pragma once
#include "vkr_shade_common.cuh"
#include "vkr_related_work.cuh"

namespace vkr {

constexpr int kOffsetErrorFactor = 28; // g_error_factor in the constant block (shared_constants.glsl:24, src/main.h:490)

template <int MAXV, bool BIASED>
VKR_DEV bool psa_sampling_error_of_light(float* out_error, const shading_point& sp, const ltc_state& l, const unsigned char* light, noise_stream& ns,
	const shading_kernel_params& p, const unsigned char* cb, uint32_t px, uint32_t py)
{
	constexpr int MAXP = MAXV + 1;
	const bool flip = dot4_point(light + L_PLANE, sp.position) < 0.0f;
	const bool specular = p.error_display >= 4;
	psa_polygon<MAXP> polygon;
	{
		f3 v[MAXP];
		const int vc = transform_and_clip<MAXP>(v, light, l.rx, l.ry, sp.normal, l.t, flip);
		if (vc == 0) return false; // the light is below the horizon (:481-482, :529-531)
		prepare_psa<MAXP, BIASED>(polygon, vc, v);
		if (polygon.psa <= 0.0f) return false; // (:486-487, :542-543)
	}
	if (specular) { // the polygon in cosine space (:506-540); nothing is displayed where the LTC vanishes on the light (:562-563)
		f3 v[MAXP];
		const int vc = transform_and_clip<MAXP>(v, light, l.cx, l.cy, l.cz, l.ct, flip);
		if (vc == 0) return false;
		prepare_psa<MAXP, BIASED>(polygon, vc, v);
		if (!(polygon.psa > 0.0f)) return false;
	}
	const f2 rnd = next_noise_2(ns, p, cb, px, py);
	const f3 sampled_dir = sample_psa<MAXP, BIASED>(polygon, rnd);
	const f3 error = sampling_error<MAXP, BIASED>(polygon, rnd, sampled_dir);
	const int component = (p.error_display - 1) % 3;
	*out_error = (component == 0) ? error.x : ((component == 1) ? error.y : error.z);
	return true;
}

// The whole contribution of one light: "return error_to_color(error) / g_exposure_factor;" (false: the light is culled and contributes nothing)
template <int MAXV>
VKR_DEV bool error_display_of_light(f3* out_color, const shading_point& sp, const ltc_state& l, const unsigned char* light, noise_stream& ns,
	const shading_kernel_params& p, const unsigned char* cb, uint32_t px, uint32_t py)
{
	float error = 0.0f;
	bool shown;
	if (p.polygon_sampling_technique == VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO) { // :462-472
		rw_light<MAXV> view;
		rw_load_light<MAXV>(view, light);
		rw_frame frame;
		frame.rx = l.rx; frame.ry = l.ry; frame.rz = sp.normal; frame.t = l.t;
		rw_sampler<VKR_TECHNIQUE_PROJECTED_SOLID_ANGLE_ARVO, MAXV> sampler;
		shown = sampler.prepare(view, sp.position, frame);
		if (shown) {
			const f2 rnd = next_noise_2(ns, p, cb, px, py);
			const f3 sampled_dir = sample_psa_arvo<MAXV + 1>(sampler.polygon, rnd, 3);
			const f2 e = sampling_error_arvo<MAXV + 1>(sampler.polygon, rnd, sampled_dir);
			error = ((p.error_display - 1) % 3 == 0) ? e.x : e.y;
		}
	}
	else if (p.biased_sampling) shown = psa_sampling_error_of_light<MAXV, true>(&error, sp, l, light, ns, p, cb, px, py);
	else shown = psa_sampling_error_of_light<MAXV, false>(&error, sp, l, light, ns, p, cb, px, py);
	if (shown) {
		const f3 color = error_to_color(error, ldf(cb, kOffsetErrorFactor));
		const float exposure = ldf(cb, OFF_EXPOSURE);
		*out_color = make3(color.x / exposure, color.y / exposure, color.z / exposure);
	}
	return shown;
}

} // namespace vkr
