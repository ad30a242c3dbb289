// oracle/glsl_compat/glsl_compat.hpp -- TEST INFRASTRUCTURE, not product code.
//
// A small GLSL-on-C++ layer (vector/matrix types with swizzles, the GLSL built-ins the shading
// pass uses, texture/ray-query stand-ins) that lets g++ compile the REFERENCE's own shader sources
// (/root/reference/src/shaders/*.glsl, where they lie) into oracle/_ref/libref_shader.so. That
// binary is the reference's code for this path running on the CPU; the oracle is pinned against it
// bit for bit (tests/test_ref_shader.py, fixtures in tests/golden/).
//
// Everything GLSL leaves implementation-defined is bound here to the SAME definitions the oracle
// states in oracle/vkr_math.h (elementary functions, dot/cross/matrix products as fma chains,
// UNORM16 texel fetch, software bilinear LTC fetch, the ray-query predicate of bvh_oracle.h).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <utility>

extern "C" {
#include "../vkr_math.h"
#include "../texture_filter.h"
}

typedef unsigned int uint;

namespace glsl {

template <class T, int N> struct vec;

template <class T> struct is_vec : std::false_type {};
template <class T, int N> struct is_vec<vec<T, N>> : std::true_type {};

// lexicographic "pack A greater than pack B" for one-directional swizzle conversions
template <int... I> struct ipack {};
template <class A, class B> struct pack_greater;
template <> struct pack_greater<ipack<>, ipack<>> : std::false_type {};
template <int A0, int... A, int B0, int... B> struct pack_greater<ipack<A0, A...>, ipack<B0, B...>>
	: std::conditional_t<(A0 > B0), std::true_type, std::conditional_t<(A0 < B0), std::false_type, pack_greater<ipack<A...>, ipack<B...>>>> {};

// ---- swizzle proxy: lives in a union with the parent's component array ----
template <class T, int P, int... I>
struct swz {
	T d[P];
	static constexpr int N = sizeof...(I);
	using V = vec<T, N>;
	operator V() const { return V(d[I]...); }
	V v() const { return V(d[I]...); }
	swz() = default;
	// conversion from a different swizzle of the same length, one direction only (so that ?: has a common type)
	template <int P2, int... J, class = std::enable_if_t<sizeof...(J) == N && pack_greater<ipack<J...>, ipack<I...>>::value>>
	swz(const swz<T, P2, J...>& o) { V t = o; int k = 0; ((d[I] = t[k++]), ...); }
	swz& operator=(const V& t) { int k = 0; ((d[I] = t[k++]), ...); return *this; }
	swz& operator=(const swz& o) { V t = o; return *this = t; }
	template <int P2, int... J> swz& operator=(const swz<T, P2, J...>& o) { V t = o; return *this = t; }
	swz& operator+=(const V& t) { return *this = v() + t; }
	swz& operator-=(const V& t) { return *this = v() - t; }
	swz& operator*=(const V& t) { return *this = v() * t; }
	swz& operator/=(const V& t) { return *this = v() / t; }
	swz& operator*=(T s) { return *this = v() * s; }
	swz& operator/=(T s) { return *this = v() / s; }
	T& operator[](int i) { constexpr int idx[] = {I...}; return d[idx[i]]; }
	T operator[](int i) const { constexpr int idx[] = {I...}; return d[idx[i]]; }
	V operator-() const { return -v(); }
};

#define GLSL_SWZ_BINOP(op) \
	template <class T, int P, int... I> vec<T, sizeof...(I)> operator op(const swz<T, P, I...>& a, const vec<T, sizeof...(I)>& b) { return a.v() op b; } \
	template <class T, int P, int... I> vec<T, sizeof...(I)> operator op(const vec<T, sizeof...(I)>& a, const swz<T, P, I...>& b) { return a op b.v(); } \
	template <class T, int P, int... I, int P2, int... J> vec<T, sizeof...(I)> operator op(const swz<T, P, I...>& a, const swz<T, P2, J...>& b) { return a.v() op b.v(); } \
	template <class T, int P, int... I> vec<T, sizeof...(I)> operator op(const swz<T, P, I...>& a, std::common_type_t<T> s) { return a.v() op s; } \
	template <class T, int P, int... I> vec<T, sizeof...(I)> operator op(std::common_type_t<T> s, const swz<T, P, I...>& a) { return s op a.v(); }
GLSL_SWZ_BINOP(+) GLSL_SWZ_BINOP(-) GLSL_SWZ_BINOP(*) GLSL_SWZ_BINOP(/)

// ---- component flattening for GLSL-style constructors ----
template <class T> struct flat {
	template <class A> static std::enable_if_t<std::is_arithmetic<A>::value> put(T*& p, const A& a) { *p++ = (T) a; }
	template <class U, int M> static void put(T*& p, const vec<U, M>& a) { for (int i = 0; i != M; ++i) *p++ = (T) a[i]; }
	template <class U, int P, int... I> static void put(T*& p, const swz<U, P, I...>& a) { ((*p++ = (T) a.d[I]), ...); }
};
template <class A> struct comp_count : std::integral_constant<int, 1> {};
template <class U, int M> struct comp_count<vec<U, M>> : std::integral_constant<int, M> {};
template <class U, int P, int... I> struct comp_count<swz<U, P, I...>> : std::integral_constant<int, (int) sizeof...(I)> {};

#define GLSL_VEC_COMMON(N) \
	vec() { for (int i = 0; i != N; ++i) d[i] = T(); } \
	vec(const vec& o) { for (int i = 0; i != N; ++i) d[i] = o.d[i]; } \
	vec& operator=(const vec& o) { for (int i = 0; i != N; ++i) d[i] = o.d[i]; return *this; } \
	template <class A, class = std::enable_if_t<std::is_arithmetic<A>::value>> explicit vec(A s) { for (int i = 0; i != N; ++i) d[i] = (T) s; } \
	template <class U, class = std::enable_if_t<!std::is_same<U, T>::value>> vec(const vec<U, N>& o) { for (int i = 0; i != N; ++i) d[i] = (T) o[i]; } \
	template <class A0, class A1, class... A, class = std::enable_if_t<(((comp_count<A0>::value + comp_count<A1>::value)) + ... + comp_count<A>::value) >= N>> \
	vec(const A0& a0, const A1& a1, const A&... a) { T tmp[16]; T* p = tmp; flat<T>::put(p, a0); flat<T>::put(p, a1); (flat<T>::put(p, a), ...); for (int i = 0; i != N; ++i) d[i] = tmp[i]; } \
	template <class U, int M, class = std::enable_if_t<(M > N)>> explicit vec(const vec<U, M>& o) { for (int i = 0; i != N; ++i) d[i] = (T) o[i]; } \
	template <class U, int P, int... I, class = std::enable_if_t<(sizeof...(I) >= N) && !(std::is_same<U, T>::value && sizeof...(I) == N)>> explicit vec(const swz<U, P, I...>& o) { T tmp[8]; T* p = tmp; flat<T>::put(p, o); for (int i = 0; i != N; ++i) d[i] = tmp[i]; } \
	T& operator[](int i) { return d[i]; } \
	const T& operator[](int i) const { return d[i]; } \
	vec& operator+=(const vec& o) { for (int i = 0; i != N; ++i) d[i] = d[i] + o.d[i]; return *this; } \
	vec& operator-=(const vec& o) { for (int i = 0; i != N; ++i) d[i] = d[i] - o.d[i]; return *this; } \
	vec& operator*=(const vec& o) { for (int i = 0; i != N; ++i) d[i] = d[i] * o.d[i]; return *this; } \
	vec& operator/=(const vec& o) { for (int i = 0; i != N; ++i) d[i] = d[i] / o.d[i]; return *this; } \
	vec& operator*=(T s) { for (int i = 0; i != N; ++i) d[i] = d[i] * s; return *this; } \
	vec& operator/=(T s) { for (int i = 0; i != N; ++i) d[i] = d[i] / s; return *this; } \
	vec& operator+=(T s) { for (int i = 0; i != N; ++i) d[i] = d[i] + s; return *this; } \
	vec& operator-=(T s) { for (int i = 0; i != N; ++i) d[i] = d[i] - s; return *this; }

template <class T> struct vec<T, 2> {
	union {
		T d[2];
		struct { T x, y; };
		struct { T r, g; };
		swz<T, 2, 0, 1> xy, rg; swz<T, 2, 1, 0> yx;
	};
	GLSL_VEC_COMMON(2)
};
template <class T> struct vec<T, 3> {
	union {
		T d[3];
		struct { T x, y, z; };
		struct { T r, g, b; };
		swz<T, 3, 0, 1> xy, rg; swz<T, 3, 1, 0> yx; swz<T, 3, 1, 2> yz; swz<T, 3, 0, 2> xz;
		swz<T, 3, 0, 1, 2> xyz, rgb; swz<T, 3, 2, 1, 0> zyx;
	};
	GLSL_VEC_COMMON(3)
};
template <class T> struct vec<T, 4> {
	union {
		T d[4];
		struct { T x, y, z, w; };
		struct { T r, g, b, a; };
		swz<T, 4, 0, 1> xy, rg; swz<T, 4, 1, 0> yx; swz<T, 4, 1, 2> yz; swz<T, 4, 2, 3> zw, ba; swz<T, 4, 0, 2> xz;
		swz<T, 4, 0, 1, 2> xyz, rgb; swz<T, 4, 1, 2, 3> yzw;
		swz<T, 4, 0, 1, 2, 3> xyzw; swz<T, 4, 2, 3, 0, 1> zwxy;
	};
	GLSL_VEC_COMMON(4)
};

#define GLSL_VEC_BINOP(op) \
	template <class T, int N> vec<T, N> operator op(const vec<T, N>& a, const vec<T, N>& b) { vec<T, N> r; for (int i = 0; i != N; ++i) r[i] = a[i] op b[i]; return r; } \
	template <class T, int N> vec<T, N> operator op(const vec<T, N>& a, std::common_type_t<T> s) { vec<T, N> r; for (int i = 0; i != N; ++i) r[i] = a[i] op s; return r; } \
	template <class T, int N> vec<T, N> operator op(std::common_type_t<T> s, const vec<T, N>& a) { vec<T, N> r; for (int i = 0; i != N; ++i) r[i] = s op a[i]; return r; }
GLSL_VEC_BINOP(+) GLSL_VEC_BINOP(-) GLSL_VEC_BINOP(*) GLSL_VEC_BINOP(/)
template <class T, int N> vec<T, N> operator-(const vec<T, N>& a) { vec<T, N> r; for (int i = 0; i != N; ++i) r[i] = -a[i]; return r; }
template <class T, int N> vec<T, N> operator>>(const vec<T, N>& a, uint s) { vec<T, N> r; for (int i = 0; i != N; ++i) r[i] = a[i] >> s; return r; }
template <class T, int P, int... I> vec<T, sizeof...(I)> operator>>(const swz<T, P, I...>& a, uint s) { return a.v() >> s; }
template <class T, int N> vec<T, N> operator&(const vec<T, N>& a, const vec<T, N>& b) { vec<T, N> r; for (int i = 0; i != N; ++i) r[i] = a[i] & b[i]; return r; }

typedef vec<float, 2> vec2; typedef vec<float, 3> vec3; typedef vec<float, 4> vec4;
typedef vec<int, 2> ivec2; typedef vec<int, 3> ivec3; typedef vec<int, 4> ivec4;
typedef vec<uint, 2> uvec2; typedef vec<uint, 3> uvec3; typedef vec<uint, 4> uvec4;

// ---- matrices: column-major, mat<C,R> has C columns of R rows (GLSL matCxR) ----
template <int C, int R> struct mat {
	vec<float, R> c[C];
	mat() {}
	explicit mat(float s) { for (int i = 0; i != C; ++i) for (int j = 0; j != R; ++j) c[i][j] = (i == j) ? s : 0.0f; }
	template <class A0, class A1, class... A> mat(const A0& a0, const A1& a1, const A&... a) {
		float tmp[32]; float* p = tmp; flat<float>::put(p, a0); flat<float>::put(p, a1); (flat<float>::put(p, a), ...);
		for (int i = 0; i != C; ++i) for (int j = 0; j != R; ++j) c[i][j] = tmp[i * R + j];
	}
	vec<float, R>& operator[](int i) { return c[i]; }
	const vec<float, R>& operator[](int i) const { return c[i]; }
	mat& operator-=(const mat& o) { for (int i = 0; i != C; ++i) c[i] -= o.c[i]; return *this; }
	mat& operator+=(const mat& o) { for (int i = 0; i != C; ++i) c[i] += o.c[i]; return *this; }
};
typedef mat<2, 2> mat2; typedef mat<3, 3> mat3; typedef mat<4, 4> mat4; typedef mat<4, 3> mat4x3; typedef mat<3, 4> mat3x4;

template <int C, int R> mat<C, R> operator-(const mat<C, R>& a) { mat<C, R> r; for (int i = 0; i != C; ++i) r[i] = -a[i]; return r; }
template <int C, int R> mat<C, R> operator-(const mat<C, R>& a, const mat<C, R>& b) { mat<C, R> r; for (int i = 0; i != C; ++i) r[i] = a[i] - b[i]; return r; }
template <int C, int R> mat<C, R> operator+(const mat<C, R>& a, const mat<C, R>& b) { mat<C, R> r; for (int i = 0; i != C; ++i) r[i] = a[i] + b[i]; return r; }
template <int C, int R> mat<C, R> operator*(const mat<C, R>& a, float s) { mat<C, R> r; for (int i = 0; i != C; ++i) r[i] = a[i] * s; return r; }
template <int C, int R> mat<C, R> operator*(float s, const mat<C, R>& a) { mat<C, R> r; for (int i = 0; i != C; ++i) r[i] = s * a[i]; return r; }
// M * v: fma chain over the columns in order (vkr_math.h)
template <int C, int R> vec<float, R> operator*(const mat<C, R>& m, const vec<float, C>& v) {
	vec<float, R> r;
	for (int j = 0; j != R; ++j) {
		float acc = m[0][j] * v[0];
		for (int i = 1; i != C; ++i) acc = fmaf(m[i][j], v[i], acc);
		r[j] = acc;
	}
	return r;
}
template <int C, int R, int P, int... I> vec<float, R> operator*(const mat<C, R>& m, const swz<float, P, I...>& v) { return m * v.v(); }
// v * M = row vector times matrix: dot(v, column)
template <int C, int R> vec<float, C> operator*(const vec<float, R>& v, const mat<C, R>& m) {
	vec<float, C> r;
	for (int i = 0; i != C; ++i) { float acc = v[0] * m[i][0]; for (int j = 1; j != R; ++j) acc = fmaf(v[j], m[i][j], acc); r[i] = acc; }
	return r;
}
// A(K cols, R rows) * B(C cols, K rows): column by column
template <int K, int R, int C> mat<C, R> operator*(const mat<K, R>& a, const mat<C, K>& b) { mat<C, R> r; for (int i = 0; i != C; ++i) r[i] = a * b[i]; return r; }
template <int C, int R> mat<R, C> transpose(const mat<C, R>& m) { mat<R, C> r; for (int i = 0; i != C; ++i) for (int j = 0; j != R; ++j) r[j][i] = m[i][j]; return r; }
inline mat2 outerProduct(const vec2& c, const vec2& r) { mat2 m; m[0] = c * r[0]; m[1] = c * r[1]; return m; }
inline mat3 outerProduct(const vec3& c, const vec3& r) { mat3 m; m[0] = c * r[0]; m[1] = c * r[1]; m[2] = c * r[2]; return m; }
inline float determinant(const mat2& m) { return m[0][0] * m[1][1] - m[1][0] * m[0][1]; }

// ---- built-ins, bound to the oracle's definitions ----
inline float fma(float a, float b, float c) { return fmaf(a, b, c); }
inline vec2 fma(const vec2& a, const vec2& b, const vec2& c) { return vec2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline vec3 fma(const vec3& a, const vec3& b, const vec3& c) { return vec3(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z)); }
inline vec4 fma(const vec4& a, const vec4& b, const vec4& c) { return vec4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w)); }
inline float dot(const vec2& a, const vec2& b) { return fmaf(a.y, b.y, a.x * b.x); }
inline float dot(const vec3& a, const vec3& b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline float dot(const vec4& a, const vec4& b) { return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x))); }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))); }
inline float determinant(const mat3& m) { return dot(m[0], cross(m[1], m[2])); }
inline float inversesqrt(float x) { return vkr_rsqrt(x); }
inline float sqrt(float x) { return sqrtf(x); }
inline float abs(float x) { return fabsf(x); }
inline vec2 abs(const vec2& a) { return vec2(fabsf(a.x), fabsf(a.y)); }
inline vec3 abs(const vec3& a) { return vec3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
inline vec2 sqrt(const vec2& a) { return vec2(sqrtf(a.x), sqrtf(a.y)); }
inline vec3 sqrt(const vec3& a) { return vec3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
inline float length(const vec2& a) { return sqrtf(dot(a, a)); }
inline float length(const vec3& a) { return sqrtf(dot(a, a)); }
inline vec2 normalize(const vec2& a) { return a * vkr_rsqrt(dot(a, a)); }
inline vec3 normalize(const vec3& a) { return a * vkr_rsqrt(dot(a, a)); }
inline float max(float x, float y) { return vkr_max(x, y); }
inline float min(float x, float y) { return vkr_min(x, y); }
inline int max(int x, int y) { return (x < y) ? y : x; }
inline int min(int x, int y) { return (y < x) ? y : x; }
inline uint max(uint x, uint y) { return (x < y) ? y : x; }
inline uint min(uint x, uint y) { return (y < x) ? y : x; }
inline vec2 max(const vec2& a, const vec2& b) { return vec2(vkr_max(a.x, b.x), vkr_max(a.y, b.y)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(vkr_max(a.x, b.x), vkr_max(a.y, b.y), vkr_max(a.z, b.z)); }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(vkr_min(a.x, b.x), vkr_min(a.y, b.y), vkr_min(a.z, b.z)); }
inline vec3 max(const vec3& a, float b) { return vec3(vkr_max(a.x, b), vkr_max(a.y, b), vkr_max(a.z, b)); }
inline float clamp(float x, float lo, float hi) { return vkr_clamp(x, lo, hi); }
inline vec3 clamp(const vec3& a, float lo, float hi) { return vec3(vkr_clamp(a.x, lo, hi), vkr_clamp(a.y, lo, hi), vkr_clamp(a.z, lo, hi)); }
inline vec2 clamp(const vec2& a, float lo, float hi) { return vec2(vkr_clamp(a.x, lo, hi), vkr_clamp(a.y, lo, hi)); }
// mix(x, y, a) = x*(1-a) + y*a
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline vec2 mix(const vec2& x, const vec2& y, float a) { return vec2(mix(x.x, y.x, a), mix(x.y, y.y, a)); }
inline vec3 mix(const vec3& x, const vec3& y, float a) { return vec3(mix(x.x, y.x, a), mix(x.y, y.y, a), mix(x.z, y.z, a)); }
inline vec3 mix(const vec3& x, const vec3& y, const vec3& a) { return vec3(mix(x.x, y.x, a.x), mix(x.y, y.y, a.y), mix(x.z, y.z, a.z)); }
inline float sin(float x) { return vkr_sin(x); }
inline float cos(float x) { return vkr_cos(x); }
inline float tan(float x) { return vkr_sin(x) / vkr_cos(x); }
inline float atan(float x) { return vkr_atan(x); }
inline float atan(float y, float x) { return vkr_atan2(y, x); }   // quadrant-corrected, built on the same polynomial (vkr_math.h)
inline float acos(float x) { return vkr_acos(x); }
inline float asin(float x) { return VKR_HALF_PI - acos(x); }
inline float pow(float x, float y) { return vkr_pow(x, y); }   // the output stage's contract (vkr_math.h); GLSL leaves pow to the driver
inline vec3 pow(const vec3& x, const vec3& y) { return vec3(vkr_pow(x.x, y.x), vkr_pow(x.y, y.y), vkr_pow(x.z, y.z)); }
// index of error_to_color()'s colour table: valid errors give 0 <= x < 20; what a NaN error turns into is undefined in GLSL (clamp, log2 and int() of
// NaN) and would read out of bounds here -- defined as the first colour (build_ref.py)
inline int glsl_float_to_int(float x) { return (x >= 0.0f && x < 20.0f) ? (int) x : 0; }
inline float log2(float x) { return vkr_log2(x); }   // only error_to_color() uses it; the contract's log2 (vkr_math.h) like pow
inline float exp2(float x) { return exp2f(x); }
inline float floor(float x) { return floorf(x); }
inline float fract(float x) { return x - floorf(x); }
inline float sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
inline bool isinf(float x) { return std::isinf(x); }
inline bool isnan(float x) { return std::isnan(x); }
inline uint floatBitsToUint(float f) { return f2u(f); }
inline int floatBitsToInt(float f) { return (int) f2u(f); }
inline float uintBitsToFloat(uint u) { return u2f(u); }
inline uint packHalf2x16(const vec2& v) { return vkr_pack_half_2x16(v.x, v.y); } // IEEE round to nearest even
inline bool any(bool b) { return b; }
struct bvec3 { bool x, y, z; };
inline bvec3 lessThanEqual(const vec3& a, const vec3& b) { return bvec3{a.x <= b.x, a.y <= b.y, a.z <= b.z}; }
inline bvec3 lessThan(const vec3& a, const vec3& b) { return bvec3{a.x < b.x, a.y < b.y, a.z < b.z}; }
inline vec3 mix(const vec3& x, const vec3& y, const bvec3& a) { return vec3(a.x ? y.x : x.x, a.y ? y.y : x.y, a.z ? y.z : x.z); }

// ---- resources ----
struct utextureBuffer { const uint32_t* data; int channels; };   // R32G32_UINT / R8_UINT texel buffers
struct textureBuffer { const uint16_t* data; };                  // R16G16B16A16_UNORM texel buffer
struct usubpassInput { int dummy; };
struct texture2DArray { const uint16_t* data; int w, h, layers; };   // RGBA16_UNORM
struct sampler2DArray { const uint16_t* data; int res, layers, channels; };
struct sampler2D { float value[4]; const vkr_texture_view_t* texture; };   // constant texture, or a mip chain filtered as oracle/texture_filter.h defines
struct accelerationStructureEXT { int dummy; };
struct rayQueryEXT { bool hit; };
enum { gl_RayFlagsTerminateOnFirstHitEXT = 4, gl_RayFlagsOpaqueEXT = 1, gl_RayFlagsSkipClosestHitShaderEXT = 8, gl_RayQueryCommittedIntersectionNoneEXT = 0 };
#define nonuniformEXT(x) (x)

typedef int (*occluded_hook_t)(const void* user, const float* origin, const float* dir, float tmin, float tmax);
extern occluded_hook_t g_occluded_hook;
extern const void* g_occluded_user;
extern thread_local uint32_t g_current_visibility;
extern const uint8_t* g_material_index_bytes;

inline uvec4 texelFetch(const utextureBuffer& t, int i) {
	if (t.channels == 2) return uvec4(t.data[2 * (size_t) i], t.data[2 * (size_t) i + 1], 0u, 1u);
	return uvec4((uint) g_material_index_bytes[i], 0u, 0u, 1u);
}
inline vec4 texelFetch(const textureBuffer& t, int i) {
	const uint16_t* p = t.data + 4 * (size_t) i;
	return vec4((float) p[0] / 65535.0f, (float) p[1] / 65535.0f, (float) p[2] / 65535.0f, (float) p[3] / 65535.0f);
}
inline vec4 texelFetch(const texture2DArray& t, const ivec3& c, int) {
	const uint16_t* p = t.data + (((size_t) c.z * t.h + c.y) * t.w + c.x) * 4;
	return vec4((float) p[0] / 65535.0f, (float) p[1] / 65535.0f, (float) p[2] / 65535.0f, (float) p[3] / 65535.0f);
}
inline uvec4 subpassLoad(const usubpassInput&) { return uvec4(g_current_visibility, 0u, 0u, 0u); }
// bilinear, clamp-to-edge, layer = round-to-nearest-even, fp32 weights (the oracle's ltc_fetch definition)
inline vec4 textureLod(const sampler2DArray& s, const vec3& c, float) {
	int res = s.res;
	float layer_r = rintf(c.z);
	int layer = (int) vkr_clamp(layer_r, 0.0f, (float) (s.layers - 1));
	float x = c.x * (float) res - 0.5f, y = c.y * (float) res - 0.5f;
	float x0f = floorf(x), y0f = floorf(y);
	float fx = x - x0f, fy = y - y0f;
	int x0 = (int) x0f, y0 = (int) y0f, x1 = x0 + 1, y1 = y0 + 1;
	x0 = x0 < 0 ? 0 : (x0 > res - 1 ? res - 1 : x0); x1 = x1 < 0 ? 0 : (x1 > res - 1 ? res - 1 : x1);
	y0 = y0 < 0 ? 0 : (y0 > res - 1 ? res - 1 : y0); y1 = y1 < 0 ? 0 : (y1 > res - 1 ? res - 1 : y1);
	const uint16_t* base = s.data + (size_t) layer * res * res * s.channels;
	vec4 out(0.0f, 0.0f, 0.0f, 1.0f);
	for (int ch = 0; ch != s.channels; ++ch) {
		float t00 = (float) base[((size_t) y0 * res + x0) * s.channels + ch] / 65535.0f;
		float t10 = (float) base[((size_t) y0 * res + x1) * s.channels + ch] / 65535.0f;
		float t01 = (float) base[((size_t) y1 * res + x0) * s.channels + ch] / 65535.0f;
		float t11 = (float) base[((size_t) y1 * res + x1) * s.channels + ch] / 65535.0f;
		float a = fmaf(fx, t10 - t00, t00);
		float b = fmaf(fx, t11 - t01, t01);
		out[ch] = fmaf(fy, b - a, a);
	}
	return out;
}
inline vec4 textureGrad(const sampler2D& s, const vec2& uv, const vec2& ddx, const vec2& ddy) {
	if (!s.texture) return vec4(s.value[0], s.value[1], s.value[2], s.value[3]);
	float out[4];
	vkr_texture_grad(out, s.texture, mk2(uv.x, uv.y), mk2(ddx.x, ddx.y), mk2(ddy.x, ddy.y));
	return vec4(out[0], out[1], out[2], out[3]);
}
// only the light textures are read with textureLod() (shading_pass.frag.glsl:182, level 0; sampler src/main.c:613-623: repeat in u, clamp to edge in v)
inline vec4 textureLod(const sampler2D& s, const vec2& c, float) {
	if (!s.texture) return vec4(s.value[0], s.value[1], s.value[2], s.value[3]);
	float t[4];
	vkr_texture_bilinear_repeat_clamp(t, s.texture, c.x, c.y);
	return vec4(t[0], t[1], t[2], t[3]);
}

inline void rayQueryInitializeEXT(rayQueryEXT& q, const accelerationStructureEXT&, uint, uint, const vec3& origin, float tmin, const vec3& dir, float tmax) {
	q.hit = g_occluded_hook(g_occluded_user, origin.d, dir.d, tmin, tmax) != 0;
}
inline bool rayQueryProceedEXT(rayQueryEXT&) { return false; }
inline uint rayQueryGetIntersectionTypeEXT(const rayQueryEXT& q, bool) { return q.hit ? 1u : 0u; }

} // namespace glsl
