// vkr_gbuffer.cuh -- one pixel of the G-buffer producer: the reference's get_shading_data() (src/shaders/shading_pass.frag.glsl:721-822).
//
// Reads the primitive index of the visibility buffer, decodes the triangle (21-bit positions, octahedral normals, 16-bit texture coordinates),
// intersects the pixel ray for barycentrics, fetches the material -- one texel where its textures are constant (TEXTURED = false), otherwise
// three textureGrad with the shader's screen-space derivatives (vkr_texture.cuh) -- and writes the 64 B of the G-buffer. A per-thread function
// without warp-level operations: vkr_gbuffer_kernel.cu wraps it into the kernel, tests/device_on_host.cpp runs it on the CPU against the oracle.
// Compile with -fmad=false.
#pragma once
#include "vkr_trace.cuh"
#include "vkr_texture.cuh"
#include "vkr_kernels.h"

namespace vkr {

enum { G_OFF_DEQUANT_FACTOR = 0, G_OFF_DEQUANT_SUMMAND = 16, G_OFF_PIXEL_TO_RAY = 96, G_OFF_CAMERA = 144, G_OFF_ROUGHNESS_FACTOR = 180 };

VKR_DEV float gldf(const unsigned char* p, int off) { return __ldg(reinterpret_cast<const float*>(p + off)); }

VKR_DEV f3 pixel_ray(const unsigned char* cb, int x, int y) { // g_pixel_to_ray_direction_world_space * vec3(pixel, 1)
	const float fx = (float) x, fy = (float) y;
	return make3(
		fmaf(gldf(cb, G_OFF_PIXEL_TO_RAY + 8), 1.0f, fmaf(gldf(cb, G_OFF_PIXEL_TO_RAY + 4), fy, gldf(cb, G_OFF_PIXEL_TO_RAY) * fx)),
		fmaf(gldf(cb, G_OFF_PIXEL_TO_RAY + 24), 1.0f, fmaf(gldf(cb, G_OFF_PIXEL_TO_RAY + 20), fy, gldf(cb, G_OFF_PIXEL_TO_RAY + 16) * fx)),
		fmaf(gldf(cb, G_OFF_PIXEL_TO_RAY + 40), 1.0f, fmaf(gldf(cb, G_OFF_PIXEL_TO_RAY + 36), fy, gldf(cb, G_OFF_PIXEL_TO_RAY + 32) * fx)));
}

VKR_DEV f3 decode_position(uint2 q, const unsigned char* cb) { // mesh_quantization.glsl:38-45
	const float px = (float) (q.x & 0x1FFFFFu);
	const float py = (float) (((q.x & 0xFFE00000u) >> 21) | ((q.y & 0x3FFu) << 11));
	const float pz = (float) ((q.y & 0x7FFFFC00u) >> 10);
	return make3(
		fmaf(px, gldf(cb, G_OFF_DEQUANT_FACTOR), gldf(cb, G_OFF_DEQUANT_SUMMAND)),
		fmaf(py, gldf(cb, G_OFF_DEQUANT_FACTOR + 4), gldf(cb, G_OFF_DEQUANT_SUMMAND + 4)),
		fmaf(pz, gldf(cb, G_OFF_DEQUANT_FACTOR + 8), gldf(cb, G_OFF_DEQUANT_SUMMAND + 8)));
}

VKR_DEV f3 decode_normal(float ox, float oy) { // mesh_quantization.glsl:19-33
	const float factor = 2.0f * (65534.0f / 65535.0f);
	const float summand = -(32768.0f / 65535.0f) * factor;
	ox = fmaf(ox, factor, summand); oy = fmaf(oy, factor, summand);
	f3 n = make3(ox, oy, 1.0f - fabsf(ox) - fabsf(oy));
	const float sx = (ox >= 0.0f) ? 1.0f : -1.0f, sy = (oy >= 0.0f) ? 1.0f : -1.0f;
	if (n.z < 0.0f) {
		const float nx = (1.0f - fabsf(n.y)) * sx;
		const float ny = (1.0f - fabsf(n.x)) * sy;
		n.x = nx; n.y = ny;
	}
	return normalize(n);
}

template <bool TEXTURED>
VKR_DEV void shade_gbuffer_pixel(const gbuffer_kernel_params& p, size_t pixel) {
	const size_t plane = (size_t) p.width * p.height;
	const int x = (int) (pixel % p.width), y = (int) (pixel / p.width);
	const uint32_t prim = p.visibility[pixel];
	const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	if (prim == 0xFFFFFFFFu) {
		p.gbuffer[pixel] = zero; p.gbuffer[plane + pixel] = zero; p.gbuffer[2 * plane + pixel] = zero; p.gbuffer[3 * plane + pixel] = zero;
		return;
	}
	const unsigned char* cb = p.constants;
	const f3 camera = make3(gldf(cb, G_OFF_CAMERA), gldf(cb, G_OFF_CAMERA + 4), gldf(cb, G_OFF_CAMERA + 8));
	const f3 ray = pixel_ray(cb, x, y);
	f3 pos[3], nrm[3]; f2 uv[3];
#pragma unroll
	for (int i = 0; i != 3; ++i) {
		const size_t vi = (size_t) prim * 3 + i;
		pos[i] = decode_position(__ldg(p.quantized_positions + vi), cb);
		const ushort4 nt = __ldg(p.normals_and_tex_coords + vi);
		nrm[i] = decode_normal((float) nt.x / 65535.0f, (float) nt.y / 65535.0f);
		uv[i] = make2(fmaf((float) nt.z / 65535.0f, 8.0f, 0.0f), fmaf((float) nt.w / 65535.0f, -8.0f, 1.0f));
	}
	const f3 e0 = pos[1] - pos[0], e1 = pos[2] - pos[0];
	const f3 ray_cross_e1 = cross(ray, e1);
	const float rcp_det = 1.0f / dot(e0, ray_cross_e1);
	const f3 ray_to_0 = camera - pos[0];
	const float by = rcp_det * dot(ray_to_0, ray_cross_e1);
	const f3 e0_cross_0 = cross(e0, ray_to_0);
	const float bz = -rcp_det * dot(ray, e0_cross_0);
	const float bx = 1.0f - (by + bz);
	const f3 position = make3(
		fmaf(bx, pos[0].x, fmaf(by, pos[1].x, bz * pos[2].x)),
		fmaf(bx, pos[0].y, fmaf(by, pos[1].y, bz * pos[2].y)),
		fmaf(bx, pos[0].z, fmaf(by, pos[1].z, bz * pos[2].z)));
	const f3 inormal = normalize(make3(
		fmaf(bx, nrm[0].x, fmaf(by, nrm[1].x, bz * nrm[2].x)),
		fmaf(bx, nrm[0].y, fmaf(by, nrm[1].y, bz * nrm[2].y)),
		fmaf(bx, nrm[0].z, fmaf(by, nrm[1].z, bz * nrm[2].z))));
	const uint32_t material_index = __ldg(p.material_indices + prim);
	f3 base; float linear_roughness, metalicity; f2 normal_texel;
	if (TEXTURED) {
		// screen-space derivatives of the barycentrics and of the texture coordinate (:754-777), then three textureGrad (:779-783)
		const f2 tex_coord = make2(fmaf(bx, uv[0].x, fmaf(by, uv[1].x, bz * uv[2].x)), fmaf(bx, uv[0].y, fmaf(by, uv[1].y, bz * uv[2].y)));
		const float det_0_dir_edge_1 = dot(ray_to_0, ray_cross_e1), det_dir_edge_0_0 = dot(ray, e0_cross_0);
		f2 tex_coord_derivs[2];
#pragma unroll
		for (int i = 0; i != 2; ++i) {
			const f3 ray_deriv = make3(gldf(cb, G_OFF_PIXEL_TO_RAY + 4 * i), gldf(cb, G_OFF_PIXEL_TO_RAY + 16 + 4 * i), gldf(cb, G_OFF_PIXEL_TO_RAY + 32 + 4 * i));
			const f3 ray_cross_e1_deriv = cross(ray_deriv, e1);
			const float rcp_det_deriv = -dot(e0, ray_cross_e1_deriv) * rcp_det * rcp_det;
			const float det_0_dir_edge_1_deriv = dot(ray_to_0, ray_cross_e1_deriv);
			const float dby = rcp_det_deriv * det_0_dir_edge_1 + rcp_det * det_0_dir_edge_1_deriv;
			const float det_dir_edge_0_0_deriv = dot(ray_deriv, e0_cross_0);
			const float dbz = -rcp_det_deriv * det_dir_edge_0_0 - rcp_det * det_dir_edge_0_0_deriv;
			const float dbx = -(dby + dbz);
			f2 d = make2(0.0f, 0.0f);
			d = d + uv[0] * dbx; d = d + uv[1] * dby; d = d + uv[2] * dbz;
			tex_coord_derivs[i] = d;
		}
		float4 texel[3];
#pragma unroll
		for (int k = 0; k != 3; ++k) {
			const uint4 dims = __ldg(p.texture_dims + 3 * material_index + k);
			texture_view view;
			view.width = dims.x; view.height = dims.y; view.mip_count = dims.z;
			view.texels = p.texture_data + __ldg(p.texture_offsets + 3 * material_index + k);
			texel[k] = texture_grad(view, tex_coord, tex_coord_derivs[0], tex_coord_derivs[1]);
		}
		base = make3(texel[0].x, texel[0].y, texel[0].z);
		linear_roughness = texel[1].y; metalicity = texel[1].z;
		normal_texel = make2(texel[2].x, texel[2].y);
	}
	else {
		const float* mp = p.material_params + 8 * (size_t) material_index;
		base = make3(__ldg(mp), __ldg(mp + 1), __ldg(mp + 2));
		linear_roughness = __ldg(mp + 3); metalicity = __ldg(mp + 4);
		normal_texel = make2(__ldg(mp + 5), __ldg(mp + 6));
	}
	f3 nts;
	nts.x = fmaf(normal_texel.x, 2.0f, -1.0f); nts.y = fmaf(normal_texel.y, 2.0f, -1.0f);
	nts.z = sqrtf(max_glsl(0.0f, fmaf(-nts.x, nts.x, fmaf(-nts.y, nts.y, 1.0f))));
	const f3 diffuse = make3(fmaf(base.x, -metalicity, base.x), fmaf(base.y, -metalicity, base.y), fmaf(base.z, -metalicity, base.z));
	const float om = 1.0f - metalicity;
	const f3 f0 = make3(0.02f * om + base.x * metalicity, 0.02f * om + base.y * metalicity, 0.02f * om + base.z * metalicity);
	float roughness = linear_roughness * linear_roughness;
	roughness = clamp_glsl(roughness * gldf(cb, G_OFF_ROUGHNESS_FACTOR), 0.0064f, 1.0f);
	const f2 te0 = uv[1] - uv[0], te1 = uv[2] - uv[0];
	const f3 n_cross_e0 = cross(inormal, e0);
	const f3 e1_cross_n = cross(e1, inormal);
	const f3 tangent = e1_cross_n * te0.x + n_cross_e0 * te1.x;
	const f3 bitangent = e1_cross_n * te0.y + n_cross_e0 * te1.y;
	const float mean_tangent_length = sqrtf(0.5f * (dot(tangent, tangent) + dot(bitangent, bitangent)));
	nts.z *= max_glsl(1.0e-10f, mean_tangent_length);
	f3 n = normalize(make3(
		fmaf(inormal.x, nts.z, fmaf(bitangent.x, nts.y, tangent.x * nts.x)),
		fmaf(inormal.y, nts.z, fmaf(bitangent.y, nts.y, tangent.y * nts.x)),
		fmaf(inormal.z, nts.z, fmaf(bitangent.z, nts.y, tangent.z * nts.x))));
	const f3 outgoing = normalize(camera - position);
	const float normal_offset = max_glsl(0.0f, 1.0e-3f - dot(n, outgoing));
	n = normalize(make3(fmaf(normal_offset, outgoing.x, n.x), fmaf(normal_offset, outgoing.y, n.y), fmaf(normal_offset, outgoing.z, n.z)));
	p.gbuffer[pixel] = make_float4(position.x, position.y, position.z, roughness);
	p.gbuffer[plane + pixel] = make_float4(n.x, n.y, n.z, 1.0f);
	p.gbuffer[2 * plane + pixel] = make_float4(diffuse.x, diffuse.y, diffuse.z, 0.0f);
	p.gbuffer[3 * plane + pixel] = make_float4(f0.x, f0.y, f0.z, 0.0f);
}

} // namespace vkr
