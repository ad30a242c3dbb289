/* oracle/vkr_math.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Elementary fp32 functions of the oracle. GLSL leaves the precision of
 * atan/sin/cos/acos/inversesqrt/normalize implementation-defined (GLSL.std.450,
 * un-pinned driver code); this header pins ONE valid instance of
 * them built only from IEEE-754 correctly-rounded +,-,*,/,sqrt,fma so that the
 * same bits can be produced on x86 (gcc -ffp-contract=off -mfma) and on sm_100a
 * (nvcc -fmad=false, explicit fmaf). The CUDA path carries its OWN
 * implementation of the same definitions (vulkan_renderer_b200/csrc/
 * vkr_device_math.cuh); the two are compared bit-for-bit by tests/.
 *
 * Definitions (the "spec" both sides implement):
 *   rsqrt(x)      = 1.0f / sqrtf(x)
 *   dot3(a,b)     = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))
 *   dot2(a,b)     = fma(a.y,b.y, a.x*b.x)
 *   cross(a,b).x  = fma(a.y,b.z, -(a.z*b.y))  (cyclic)
 *   normalize(v)  = v * rsqrt(dot(v,v))
 *   M*v (n cols)  = fma chain: ((c0*v0 then fma c1,v1 ...)) in column order
 *   atan(x)       = odd minimax polynomial on [0,1] (9 coefficients in x^2),
 *                   1/x reflection for |x|>1 with a two-term pi/2
 *   sin/cos(x)    = Cody-Waite 3-term reduction by pi/2, Cephes sinf/cosf kernels
 *   acos(x)       = 2*atan(sqrt((1-x)/(1+x)))   for x in [0,1]
 */
#ifndef VKR_ORACLE_MATH_H
#define VKR_ORACLE_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float x, y; } v2;
typedef struct { float x, y, z; } v3;

#define VKR_PI 3.1415926535897932384626433832795f
#define VKR_INV_PI 0.31830988618379067153776752674503f
#define VKR_HALF_PI 1.5707963267948966192313216916398f

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline float vkr_rsqrt(float x) { return 1.0f / sqrtf(x); }
/* GLSL.std.450 FMax/FMin wording: max(x,y) = (x<y)?y:x, min(x,y) = (y<x)?y:x */
static inline float vkr_max(float x, float y) { return (x < y) ? y : x; }
static inline float vkr_min(float x, float y) { return (y < x) ? y : x; }
static inline float vkr_clamp(float x, float lo, float hi) { return vkr_min(vkr_max(x, lo), hi); }

static inline float vkr_atan(float x) {
	float ax = fabsf(x);
	int big = ax > 1.0f;
	float z = big ? (1.0f / ax) : ax;
	float s = z * z;
	float q = 0.002849885728210211f;
	q = fmaf(q, s, -0.016068613156676292f);
	q = fmaf(q, s, 0.042691491544246674f);
	q = fmaf(q, s, -0.07504292577505112f);
	q = fmaf(q, s, 0.10640932619571686f);
	q = fmaf(q, s, -0.14203643798828125f);
	q = fmaf(q, s, 0.1999261975288391f);
	q = fmaf(q, s, -0.3333307206630707f);
	float r = fmaf(z * s, q, z);
	/* pi/2 = 1.57079637050628662109375 (hi) - 4.37113882867379e-8 (lo) */
	if (big) r = (1.57079637050628662109375f - r) + (-4.37113882867379e-8f);
	return (x < 0.0f) ? -r : r;
}

/* Reduces x to r in [-pi/4,pi/4] and quadrant k (x = k*pi/2 + r). Accurate for |x| < ~1e4. */
static inline float vkr_reduce_pio2(float x, int* quadrant) {
	float k = rintf(x * 0.63661977236758134308f);
	float r = fmaf(-k, 1.5707962512969970703125f, x);
	r = fmaf(-k, 7.54978995489188216e-08f, r);
	r = fmaf(-k, 5.39030285815811905e-15f, r);
	*quadrant = (int) k;
	return r;
}
static inline float vkr_sin_kernel(float r) {
	float s = r * r;
	float p = -1.9515295891e-4f;
	p = fmaf(p, s, 8.3321608736e-3f);
	p = fmaf(p, s, -1.6666654611e-1f);
	return fmaf(r * s, p, r);
}
static inline float vkr_cos_kernel(float r) {
	float s = r * r;
	float p = 2.443315711809948e-5f;
	p = fmaf(p, s, -1.388731625493765e-3f);
	p = fmaf(p, s, 4.166664568298827e-2f);
	return fmaf(s * s, p, fmaf(-0.5f, s, 1.0f));
}
static inline float vkr_sin(float x) {
	int q; float r = vkr_reduce_pio2(x, &q);
	float v = (q & 1) ? vkr_cos_kernel(r) : vkr_sin_kernel(r);
	return (q & 2) ? -v : v;
}
static inline float vkr_cos(float x) {
	int q; float r = vkr_reduce_pio2(x, &q);
	float v = (q & 1) ? vkr_sin_kernel(r) : vkr_cos_kernel(r);
	return ((q + 1) & 2) ? -v : v;
}
/* acos for x in [0,1] (callers clamp first, ltc_utility.glsl:61) */
static inline float vkr_acos01(float x) {
	return 2.0f * vkr_atan(sqrtf((1.0f - x) / (1.0f + x)));
}

static inline v2 mk2(float x, float y) { v2 r = {x, y}; return r; }
static inline v3 mk3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline float dot2(v2 a, v2 b) { return fmaf(a.y, b.y, a.x * b.x); }
static inline float dot3(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline v3 cross3(v3 a, v3 b) {
	return mk3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
static inline v3 scale3(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline v2 scale2(v2 a, float s) { return mk2(a.x * s, a.y * s); }
static inline v3 add3(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v2 add2(v2 a, v2 b) { return mk2(a.x + b.x, a.y + b.y); }
static inline v2 sub2(v2 a, v2 b) { return mk2(a.x - b.x, a.y - b.y); }
static inline v3 normalize3(v3 a) { return scale3(a, vkr_rsqrt(dot3(a, a))); }
static inline v2 normalize2(v2 a) { return scale2(a, vkr_rsqrt(dot2(a, a))); }
/* det of the 3x3 matrix with COLUMNS a,b,c = dot(a, cross(b,c)) */
static inline float det3(v3 a, v3 b, v3 c) { return dot3(a, cross3(b, c)); }

#endif
