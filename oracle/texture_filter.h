/* oracle/texture_filter.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * What textureGrad() does for the material textures (src/shaders/shading_pass.frag.glsl:779-783) with the reference's sampler
 * (src/scene.c:546-552: linear min / mag / mip filters, 16x anisotropy, repeat addressing). Vulkan leaves anisotropic filtering to
 * the implementation; the reference's values come out of an un-pinned driver. This header DEFINES one valid instance,
 * built from IEEE fp32 operations only, that the oracle, the reference shader compiled as C++ (oracle/glsl_compat) and the CUDA
 * producer (csrc/vkr_texture.cuh) share word for word:
 *   - texel-space derivatives Px = dUV/dx * size, Py = dUV/dy * size at level 0; the longer one is the major axis,
 *   - N = clamp(ceil(|Pmajor| / |Pminor|), 1, 16) taps, evenly spread along the major axis,
 *   - level of detail log2(|Pmajor| / N), clamped to the mip chain, trilinear,
 *   - every tap: bilinear with fp32 weights fma(fx, t10 - t00, t00), texel centres at (i + 0.5) / size, indices wrapped (repeat).
 * Textures are RGBA32F mip chains (whatever the file format was: the loader decodes RGBA16F, BC1 and BC5 first).
 */
#ifndef VKR_TEXTURE_FILTER_H
#define VKR_TEXTURE_FILTER_H
#include "vkr_math.h"

typedef struct {
	uint32_t width, height, mip_count;
	const float* texels; /* level 0 first, then level 1, ...; level l is max(width >> l, 1) x max(height >> l, 1) RGBA */
} vkr_texture_view_t;

static inline uint32_t vkr_texture_level_size(uint32_t size, uint32_t level) { uint32_t s = size >> level; return s ? s : 1u; }
static inline const float* vkr_texture_level(const vkr_texture_view_t* t, uint32_t level, uint32_t* w, uint32_t* h) {
	const float* p = t->texels;
	for (uint32_t l = 0; l != level; ++l) p += 4 * (size_t) vkr_texture_level_size(t->width, l) * vkr_texture_level_size(t->height, l);
	*w = vkr_texture_level_size(t->width, level); *h = vkr_texture_level_size(t->height, level);
	return p;
}
static inline int vkr_texture_wrap(int i, int n) { int m = i % n; return (m < 0) ? m + n : m; }

static inline void vkr_texture_bilinear(float out[4], const vkr_texture_view_t* t, uint32_t level, float u, float v) {
	uint32_t w, h;
	const float* texels = vkr_texture_level(t, level, &w, &h);
	float x = u * (float) w - 0.5f, y = v * (float) h - 0.5f;
	if (!(fabsf(x) < 1.0e9f)) x = 0.0f; /* NaN or absurdly far away: defined as the first texel */
	if (!(fabsf(y) < 1.0e9f)) y = 0.0f;
	const float x0f = floorf(x), y0f = floorf(y);
	const float fx = x - x0f, fy = y - y0f;
	const int x0 = vkr_texture_wrap((int) x0f, (int) w), x1 = vkr_texture_wrap((int) x0f + 1, (int) w);
	const int y0 = vkr_texture_wrap((int) y0f, (int) h), y1 = vkr_texture_wrap((int) y0f + 1, (int) h);
	for (int c = 0; c != 4; ++c) {
		const float t00 = texels[4 * ((size_t) y0 * w + x0) + c], t10 = texels[4 * ((size_t) y0 * w + x1) + c];
		const float t01 = texels[4 * ((size_t) y1 * w + x0) + c], t11 = texels[4 * ((size_t) y1 * w + x1) + c];
		const float a = fmaf(fx, t10 - t00, t00);
		const float b = fmaf(fx, t11 - t01, t01);
		out[c] = fmaf(fy, b - a, a);
	}
}

/* textureLod(..., 0.0f) with the sampler of the light textures (src/main.c:613-623: linear filters, REPEAT in u, CLAMP_TO_EDGE in v
 * -- "the way to go for theta-phi parametrizations"): one bilinear tap on level 0 with the conventions of vkr_texture_bilinear(). */
static inline void vkr_texture_bilinear_repeat_clamp(float out[4], const vkr_texture_view_t* t, float u, float v) {
	const uint32_t w = t->width, h = t->height;
	const float* texels = t->texels;
	float x = u * (float) w - 0.5f, y = v * (float) h - 0.5f;
	if (!(fabsf(x) < 1.0e9f)) x = 0.0f;
	if (!(fabsf(y) < 1.0e9f)) y = 0.0f;
	const float x0f = floorf(x), y0f = floorf(y);
	const float fx = x - x0f, fy = y - y0f;
	const int x0 = vkr_texture_wrap((int) x0f, (int) w), x1 = vkr_texture_wrap((int) x0f + 1, (int) w);
	int y0 = (int) y0f, y1 = (int) y0f + 1;
	y0 = (y0 < 0) ? 0 : ((y0 > (int) h - 1) ? (int) h - 1 : y0);
	y1 = (y1 < 0) ? 0 : ((y1 > (int) h - 1) ? (int) h - 1 : y1);
	for (int c = 0; c != 4; ++c) {
		const float t00 = texels[4 * ((size_t) y0 * w + x0) + c], t10 = texels[4 * ((size_t) y0 * w + x1) + c];
		const float t01 = texels[4 * ((size_t) y1 * w + x0) + c], t11 = texels[4 * ((size_t) y1 * w + x1) + c];
		const float a = fmaf(fx, t10 - t00, t00);
		const float b = fmaf(fx, t11 - t01, t01);
		out[c] = fmaf(fy, b - a, a);
	}
}

static inline void vkr_texture_grad(float out[4], const vkr_texture_view_t* t, v2 uv, v2 ddx, v2 ddy) {
	const v2 px = mk2(ddx.x * (float) t->width, ddx.y * (float) t->height), py = mk2(ddy.x * (float) t->width, ddy.y * (float) t->height);
	const float lx2 = dot2(px, px), ly2 = dot2(py, py);
	const int major_is_x = lx2 >= ly2;
	const float pmax2 = major_is_x ? lx2 : ly2, pmin2 = major_is_x ? ly2 : lx2;
	float ratio = (pmin2 > 0.0f) ? sqrtf(pmax2 / pmin2) : ((pmax2 > 0.0f) ? 16.0f : 1.0f);
	ratio = ceilf(ratio);
	const int taps = (ratio >= 1.0f) ? ((ratio <= 16.0f) ? (int) ratio : 16) : 1; /* NaN -> 1 */
	const float footprint2 = pmax2 / (float) (taps * taps);
	float lod = (footprint2 > 1.0f) ? 0.5f * vkr_log2(footprint2) : 0.0f; /* minification only; magnified or degenerate footprints stay on level 0 */
	const float max_lod = (float) (t->mip_count - 1);
	lod = (lod < max_lod) ? lod : max_lod; /* an infinite or NaN footprint never passes "> 1.0f" as NaN; +inf ends on the last level */
	const float l0f = floorf(lod);
	const float f = lod - l0f;
	const uint32_t l0 = (uint32_t) l0f, l1 = (l0 + 1 < t->mip_count) ? l0 + 1 : l0;
	const v2 major = major_is_x ? ddx : ddy;
	float acc[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
	for (int i = 0; i != taps; ++i) {
		const float s = ((float) i + 0.5f) / (float) taps - 0.5f;
		const float u = fmaf(major.x, s, uv.x), v = fmaf(major.y, s, uv.y);
		float c0[4], c1[4];
		vkr_texture_bilinear(c0, t, l0, u, v);
		vkr_texture_bilinear(c1, t, l1, u, v);
		for (int c = 0; c != 4; ++c) acc[c] += fmaf(f, c1[c] - c0[c], c0[c]);
	}
	const float rcp_taps = 1.0f / (float) taps;
	for (int c = 0; c != 4; ++c) out[c] = acc[c] * rcp_taps;
}

#endif
