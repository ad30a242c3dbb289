// vkr_texture.cuh -- filtering of material textures in the G-buffer producer.
//
// textureGrad() with the reference's material sampler (src/scene.c:546-552: linear filters, 16x anisotropy, repeat) is left to the
// driver by Vulkan. This is the definition this library uses, the same one as oracle/texture_filter.h word for word (tests hold the
// two bit for bit, tests/test_device_on_host.py): texel-space derivatives at level 0, N = clamp(ceil(major / minor), 1, 16) taps along
// the major axis, level of detail log2(major / N) clamped to the chain, trilinear, bilinear taps with fp32 weights and wrapped indices.
// Textures are RGBA32F mip chains decoded on the host (csrc/vkr_textures.cpp), one float4 per texel: a tap is 8 x LDG.128 through the
// read-only path. No warp intrinsics here, so the header also compiles for the CPU.
#pragma once
#include "vkr_device_math.cuh"

namespace vkr {

struct texture_view {
	uint32_t width, height, mip_count;
	const float4* texels; // level 0 first; level l is max(width >> l, 1) x max(height >> l, 1)
};

VKR_DEV uint32_t texture_level_size(uint32_t size, uint32_t level) { const uint32_t s = size >> level; return s ? s : 1u; }
VKR_DEV int texture_wrap(int i, int n) { const int m = i % n; return (m < 0) ? m + n : m; }

VKR_DEV float4 texture_bilinear(const texture_view& t, uint32_t level, float u, float v) {
	const float4* texels = t.texels;
	for (uint32_t l = 0; l != level; ++l) texels += (size_t) texture_level_size(t.width, l) * texture_level_size(t.height, l);
	const uint32_t w = texture_level_size(t.width, level), h = texture_level_size(t.height, level);
	float x = u * (float) w - 0.5f, y = v * (float) h - 0.5f;
	if (!(fabsf(x) < 1.0e9f)) x = 0.0f; // NaN or absurdly far away: defined as the first texel
	if (!(fabsf(y) < 1.0e9f)) y = 0.0f;
	const float x0f = floorf(x), y0f = floorf(y);
	const float fx = x - x0f, fy = y - y0f;
	const int x0 = texture_wrap((int) x0f, (int) w), x1 = texture_wrap((int) x0f + 1, (int) w);
	const int y0 = texture_wrap((int) y0f, (int) h), y1 = texture_wrap((int) y0f + 1, (int) h);
	const float4 t00 = __ldg(texels + (size_t) y0 * w + x0), t10 = __ldg(texels + (size_t) y0 * w + x1);
	const float4 t01 = __ldg(texels + (size_t) y1 * w + x0), t11 = __ldg(texels + (size_t) y1 * w + x1);
	float4 out;
	{ const float a = fmaf(fx, t10.x - t00.x, t00.x), b = fmaf(fx, t11.x - t01.x, t01.x); out.x = fmaf(fy, b - a, a); }
	{ const float a = fmaf(fx, t10.y - t00.y, t00.y), b = fmaf(fx, t11.y - t01.y, t01.y); out.y = fmaf(fy, b - a, a); }
	{ const float a = fmaf(fx, t10.z - t00.z, t00.z), b = fmaf(fx, t11.z - t01.z, t01.z); out.z = fmaf(fy, b - a, a); }
	{ const float a = fmaf(fx, t10.w - t00.w, t00.w), b = fmaf(fx, t11.w - t01.w, t01.w); out.w = fmaf(fy, b - a, a); }
	return out;
}

// textureLod(..., 0.0f) with the sampler of the light textures (src/main.c:613-623: linear filters, repeat in u, clamp to edge in v):
// one bilinear tap on level 0; oracle/texture_filter.h: vkr_texture_bilinear_repeat_clamp, word for word
VKR_DEV float4 texture_bilinear_repeat_clamp(const texture_view& t, float u, float v) {
	const float4* texels = t.texels;
	const uint32_t w = t.width, h = t.height;
	float x = u * (float) w - 0.5f, y = v * (float) h - 0.5f;
	if (!(fabsf(x) < 1.0e9f)) x = 0.0f;
	if (!(fabsf(y) < 1.0e9f)) y = 0.0f;
	const float x0f = floorf(x), y0f = floorf(y);
	const float fx = x - x0f, fy = y - y0f;
	const int x0 = texture_wrap((int) x0f, (int) w), x1 = texture_wrap((int) x0f + 1, (int) w);
	int y0 = (int) y0f, y1 = (int) y0f + 1;
	y0 = (y0 < 0) ? 0 : ((y0 > (int) h - 1) ? (int) h - 1 : y0);
	y1 = (y1 < 0) ? 0 : ((y1 > (int) h - 1) ? (int) h - 1 : y1);
	const float4 t00 = __ldg(texels + (size_t) y0 * w + x0), t10 = __ldg(texels + (size_t) y0 * w + x1);
	const float4 t01 = __ldg(texels + (size_t) y1 * w + x0), t11 = __ldg(texels + (size_t) y1 * w + x1);
	float4 out;
	{ const float a = fmaf(fx, t10.x - t00.x, t00.x), b = fmaf(fx, t11.x - t01.x, t01.x); out.x = fmaf(fy, b - a, a); }
	{ const float a = fmaf(fx, t10.y - t00.y, t00.y), b = fmaf(fx, t11.y - t01.y, t01.y); out.y = fmaf(fy, b - a, a); }
	{ const float a = fmaf(fx, t10.z - t00.z, t00.z), b = fmaf(fx, t11.z - t01.z, t01.z); out.z = fmaf(fy, b - a, a); }
	{ const float a = fmaf(fx, t10.w - t00.w, t00.w), b = fmaf(fx, t11.w - t01.w, t01.w); out.w = fmaf(fy, b - a, a); }
	return out;
}

VKR_DEV float4 texture_grad(const texture_view& t, f2 uv, f2 ddx, f2 ddy) {
	const f2 px = make2(ddx.x * (float) t.width, ddx.y * (float) t.height), py = make2(ddy.x * (float) t.width, ddy.y * (float) t.height);
	const float lx2 = dot(px, px), ly2 = dot(py, py);
	const bool major_is_x = lx2 >= ly2;
	const float pmax2 = major_is_x ? lx2 : ly2, pmin2 = major_is_x ? ly2 : lx2;
	float ratio = (pmin2 > 0.0f) ? sqrtf(pmax2 / pmin2) : ((pmax2 > 0.0f) ? 16.0f : 1.0f);
	ratio = ceilf(ratio);
	const int taps = (ratio >= 1.0f) ? ((ratio <= 16.0f) ? (int) ratio : 16) : 1; // NaN -> 1
	const float footprint2 = pmax2 / (float) (taps * taps);
	float lod = (footprint2 > 1.0f) ? 0.5f * log2_poly(footprint2) : 0.0f; // minification only
	const float max_lod = (float) (t.mip_count - 1);
	lod = (lod < max_lod) ? lod : max_lod;
	const float l0f = floorf(lod);
	const float f = lod - l0f;
	const uint32_t l0 = (uint32_t) l0f, l1 = (l0 + 1 < t.mip_count) ? l0 + 1 : l0;
	const f2 major = major_is_x ? ddx : ddy;
	float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	for (int i = 0; i != taps; ++i) {
		const float s = ((float) i + 0.5f) / (float) taps - 0.5f;
		const float u = fmaf(major.x, s, uv.x), v = fmaf(major.y, s, uv.y);
		const float4 c0 = texture_bilinear(t, l0, u, v), c1 = texture_bilinear(t, l1, u, v);
		acc.x += fmaf(f, c1.x - c0.x, c0.x); acc.y += fmaf(f, c1.y - c0.y, c0.y); acc.z += fmaf(f, c1.z - c0.z, c0.z); acc.w += fmaf(f, c1.w - c0.w, c0.w);
	}
	const float rcp_taps = 1.0f / (float) taps;
	return make_float4(acc.x * rcp_taps, acc.y * rcp_taps, acc.z * rcp_taps, acc.w * rcp_taps);
}

} // namespace vkr
