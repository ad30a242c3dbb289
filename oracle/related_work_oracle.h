/* oracle/related_work_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Scalar fp32 restatement of the reference's related-work polygon samplers:
 *   src/shaders/polygon_sampling_related_work.glsl:38-1048  (Turk, Urena, Arvo x2, Hart x2)
 *   src/shaders/polygon_sampling.glsl:120-229               (solid angle sampling, "ours")
 *   src/shaders/cubic_solver.glsl:29-76
 * Every function cites the GLSL lines it follows; expression order and the placement of fma() follow the GLSL
 * text, plain a*b+c is NOT contracted (-ffp-contract=off), elementary functions come from vkr_math.h.
 * maxp = MAX_POLYGON_VERTEX_COUNT of the shader configuration (src/main.c:194-216: light vertices + 1 for the
 * techniques that clip, light vertices otherwise).
 *
 * Parity status: pinned bit for bit against the reference's own GLSL compiled as C++ (oracle/_ref, configurations
 * "_q<technique>"), frozen in tests/golden/ref_shader.npz.
 */
#ifndef VKR_RELATED_WORK_ORACLE_H
#define VKR_RELATED_WORK_ORACLE_H
#include "psa_oracle.h"

#define RW_MAXP PSA_MAXP

static inline v3 rw_neg3(v3 a) { return mk3(-a.x, -a.y, -a.z); }
/* a*x + b*y (+ c*z) on vectors: componentwise products and sums, left to right, no contraction */
static inline v3 rw_lin2(float a, v3 x, float b, v3 y) { return add3(scale3(x, a), scale3(y, b)); }
static inline v3 rw_lin3(float a, v3 x, float b, v3 y, float c, v3 z) { return add3(add3(scale3(x, a), scale3(y, b)), scale3(z, c)); }

/* ---- polygon_sampling_related_work.glsl:38-66 (Turk) */
static inline v3 rw_sample_area_polygon_turk(uint32_t vertex_count, const v3* vertices, const v2* fan_areas, v2 rnd, uint32_t maxp) {
	float target_area = fan_areas[maxp - 3].y * rnd.x;
	float subtriangle_area = target_area;
	float triangle_area = fan_areas[0].x;
	v3 tv[3] = { vertices[1], vertices[0], vertices[2] };
	for (uint32_t i = 0; i != maxp - 3; ++i) {
		if (i + 3 >= vertex_count || fan_areas[i].y >= target_area) break;
		subtriangle_area = target_area - fan_areas[i].y;
		triangle_area = fan_areas[i + 1].x;
		tv[0] = vertices[i + 2];
		tv[2] = vertices[i + 3];
	}
	rnd.x = subtriangle_area / triangle_area;
	float sqrt_random_0 = sqrtf(rnd.x);
	float b0 = 1.0f - sqrt_random_0, b1 = sqrt_random_0 * rnd.y, b2 = fmaf(-sqrt_random_0, rnd.y, sqrt_random_0);
	return rw_lin3(b0, tv[0], b1, tv[1], b2, tv[2]);
}

/* :81-88 */
static inline float rw_get_area_sample_density(v3* out_normalized_dir, v3 light_sample, v3 shading_position, v3 light_normal, float light_area) {
	v3 dir = sub3(light_sample, shading_position);
	float distance_squared = dot3(dir, dir);
	float normalization = vkr_rsqrt(distance_squared);
	dir = scale3(dir, normalization);
	*out_normalized_dir = dir;
	float projected_area = fabsf(dot3(light_normal, dir)) * light_area;
	return distance_squared / projected_area;
}

/* ---- :100-112 (Urena) */
typedef struct {
	v3 o, x, y, z;
	float z0, z0sq, x0, y0, y0sq, x1, y1, y1sq, b0, b1, b0sq, k, solid_angle;
} rw_urena_t;

/* :127-170. rotation_cols = the columns of local_to_world_space */
static inline rw_urena_t rw_prepare_urena(v3 s, float exl, float eyl, const v3 rotation_cols[3], v3 o) {
	rw_urena_t q;
	q.o = o;
	q.x = rotation_cols[0]; q.y = rotation_cols[1]; q.z = rotation_cols[2];
	v3 d = sub3(s, o);
	q.z0 = dot3(d, q.z);
	q.z = (q.z0 > 0.0f) ? rw_neg3(q.z) : q.z;
	q.z0 = -fabsf(q.z0);
	q.z0sq = q.z0 * q.z0;
	q.x0 = dot3(d, q.x);
	q.y0 = dot3(d, q.y);
	q.x1 = q.x0 + exl;
	q.y1 = q.y0 + eyl;
	q.y0sq = q.y0 * q.y0;
	q.y1sq = q.y1 * q.y1;
	v3 v00 = mk3(q.x0, q.y0, q.z0), v01 = mk3(q.x0, q.y1, q.z0), v10 = mk3(q.x1, q.y0, q.z0), v11 = mk3(q.x1, q.y1, q.z0);
	v3 n0 = normalize3(cross3(v00, v10));
	v3 n1 = normalize3(cross3(v10, v11));
	v3 n2 = normalize3(cross3(v11, v01));
	v3 n3 = normalize3(cross3(v01, v00));
	float g0 = vkr_acos(-dot3(n0, n1));
	float g1 = vkr_acos(-dot3(n1, n2));
	float g2 = vkr_acos(-dot3(n2, n3));
	float g3 = vkr_acos(-dot3(n3, n0));
	q.b0 = n0.z;
	q.b1 = n2.z;
	q.b0sq = q.b0 * q.b0;
	q.k = 2.0f * VKR_PI - g2 - g3;
	q.solid_angle = g0 + g1 - q.k;
	return q;
}

/* :177-200 */
static inline v3 rw_sample_urena(const rw_urena_t* q, v2 rnd) {
	float u = rnd.x, v = rnd.y;
	float au = fmaf(u, q->solid_angle, q->k);
	float fu = fmaf(vkr_cos(au), q->b0, -q->b1) / vkr_sin(au);
	float cu = vkr_rsqrt(fmaf(fu, fu, q->b0sq));
	cu = (fu > 0.0f) ? cu : -cu;
	cu = vkr_clamp(cu, -1.0f, 1.0f);
	float xu = -(cu * q->z0) * vkr_rsqrt(fmaf(-cu, cu, 1.0f));
	xu = vkr_clamp(xu, q->x0, q->x1);
	float d = sqrtf(xu * xu + q->z0sq);
	float h0 = q->y0 * vkr_rsqrt(fmaf(d, d, q->y0sq));
	float h1 = q->y1 * vkr_rsqrt(fmaf(d, d, q->y1sq));
	float hv = h0 + v * (h1 - h0);
	float mhv2_1 = fmaf(-hv, hv, 1.0f);
	float yv = (mhv2_1 >= 0.0f) ? ((hv * d) * vkr_rsqrt(mhv2_1)) : q->y1;
	return normalize3(rw_lin3(xu, q->x, yv, q->y, q->z0, q->z));
}

/* ---- :209-224 (Arvo, solid angle) */
typedef struct {
	uint32_t vertex_count;
	v3 vertex_dirs[RW_MAXP];
	float fan_solid_angles[RW_MAXP];
	v2 opposite_dirs[RW_MAXP];
	float solid_angle;
} rw_sa_arvo_t;

/* :229-264 */
static inline void rw_prepare_sa_arvo(rw_sa_arvo_t* p, uint32_t vertex_count, const v3* vertices, v3 shading_position, uint32_t maxp) {
	memset(p, 0, sizeof(*p));
	for (uint32_t i = 0; i != maxp; ++i) p->vertex_dirs[i] = normalize3(sub3(vertices[i], shading_position));
	float solid_angle = 0.0f;
	for (uint32_t i = 0; i != maxp - 2; ++i) {
		if (i >= 1 && i + 2 >= vertex_count) break;
		v3 n0 = normalize3(cross3(sub3(p->vertex_dirs[i + 1], p->vertex_dirs[0]), p->vertex_dirs[0]));
		v3 n1 = normalize3(cross3(sub3(p->vertex_dirs[i + 2], p->vertex_dirs[i + 1]), p->vertex_dirs[i + 1]));
		p->opposite_dirs[i].x = -dot3(n0, n1);
		p->opposite_dirs[i].y = sqrtf(vkr_max(0.0f, fmaf(-p->opposite_dirs[i].x, p->opposite_dirs[i].x, 1.0f)));
		float dot_0_1 = dot3(p->vertex_dirs[0], p->vertex_dirs[i + 1]);
		float dot_0_2 = dot3(p->vertex_dirs[0], p->vertex_dirs[i + 2]);
		float dot_1_2 = dot3(p->vertex_dirs[i + 1], p->vertex_dirs[i + 2]);
		float simplex_volume = det3(p->vertex_dirs[0], p->vertex_dirs[i + 1], p->vertex_dirs[i + 2]);
		float tangent = fabsf(simplex_volume) / (1.0f + dot_0_1 + dot_0_2 + dot_1_2);
		solid_angle += 2.0f * psa_positive_atan(tangent, 0);
		p->fan_solid_angles[i] = solid_angle;
	}
	p->solid_angle = solid_angle;
	p->vertex_count = vertex_count;
}

/* :269-304 */
static inline v3 rw_sample_sa_arvo(const rw_sa_arvo_t* p, v2 rnd, uint32_t maxp) {
	float target_solid_angle = p->solid_angle * rnd.x;
	float subtriangle_solid_angle = target_solid_angle;
	v2 opposite_dir = p->opposite_dirs[0];
	v3 tv[3] = { p->vertex_dirs[1], p->vertex_dirs[0], p->vertex_dirs[2] };
	for (uint32_t i = 0; i != maxp - 3; ++i) {
		if (i + 3 >= p->vertex_count || p->fan_solid_angles[i] >= target_solid_angle) break;
		subtriangle_solid_angle = target_solid_angle - p->fan_solid_angles[i];
		tv[0] = p->vertex_dirs[i + 2];
		tv[2] = p->vertex_dirs[i + 3];
		opposite_dir = p->opposite_dirs[i + 1];
	}
	v2 sd = mk2(vkr_cos(subtriangle_solid_angle), vkr_sin(subtriangle_solid_angle));
	float pp = sd.y * opposite_dir.x - sd.x * opposite_dir.y;
	float qq = sd.y * opposite_dir.y + sd.x * opposite_dir.x;
	float u = qq - opposite_dir.x;
	float v = pp + opposite_dir.y * dot3(tv[0], tv[1]);
	float s = ((v * qq - u * pp) * opposite_dir.x - v) / ((v * pp + u * qq) * opposite_dir.y);
	v3 edge_tangent_2_0 = normalize3(sub3(tv[2], scale3(tv[0], dot3(tv[0], tv[2]))));
	v3 vertex_2 = rw_lin2(s, tv[0], sqrtf(vkr_clamp(fmaf(-s, s, 1.0f), 0.0f, 1.0f)), edge_tangent_2_0);
	float z = 1.0f - rnd.y * (1.0f - dot3(vertex_2, tv[1]));
	v3 edge_tangent_2_1 = normalize3(sub3(vertex_2, scale3(tv[1], dot3(tv[1], vertex_2))));
	return rw_lin2(z, tv[1], sqrtf(vkr_clamp(fmaf(-z, z, 1.0f), 0.0f, 1.0f)), edge_tangent_2_1);
}

/* ---- polygon_sampling.glsl:61-76 (solid angle sampling, ours) */
typedef struct {
	uint32_t vertex_count;
	v3 vertex_dirs[RW_MAXP];
	v3 triangle_parameters[RW_MAXP];
	float fan_solid_angles[RW_MAXP];
	float solid_angle;
} rw_sa_t;

/* polygon_sampling.glsl:120-175 */
static inline void rw_prepare_sa(rw_sa_t* p, uint32_t vertex_count, const v3* vertices, v3 shading_position, uint32_t maxp, int biased) {
	memset(p, 0, sizeof(*p));
	p->vertex_count = vertex_count;
	for (uint32_t i = 0; i != maxp; ++i) p->vertex_dirs[i] = normalize3(sub3(vertices[i], shading_position));
	float householder_sign = (p->vertex_dirs[0].x > 0.0f) ? -1.0f : 1.0f;
	float hs = 1.0f / (fabsf(p->vertex_dirs[0].x) + 1.0f);
	v2 householder_yz = mk2(p->vertex_dirs[0].y * hs, p->vertex_dirs[0].z * hs);
	p->solid_angle = 0.0f;
	float previous_dot_1_2 = dot3(p->vertex_dirs[0], p->vertex_dirs[1]);
	for (uint32_t i = 0; i != maxp - 2; ++i) {
		if (i >= 1 && i + 2 >= vertex_count) break;
		v3 vs[3] = { p->vertex_dirs[i + 1], p->vertex_dirs[0], p->vertex_dirs[i + 2] };
		float dot_0_1 = previous_dot_1_2;
		float dot_0_2 = dot3(vs[0], vs[2]);
		float dot_1_2 = dot3(vs[1], vs[2]);
		previous_dot_1_2 = dot_1_2;
		float dot_householder_0 = fmaf(-householder_sign, vs[0].x, dot_0_1);
		float dot_householder_2 = fmaf(-householder_sign, vs[2].x, dot_1_2);
		v2 c0 = mk2(fmaf(-dot_householder_0, householder_yz.x, vs[0].y), fmaf(-dot_householder_0, householder_yz.y, vs[0].z));
		v2 c1 = mk2(fmaf(-dot_householder_2, householder_yz.x, vs[2].y), fmaf(-dot_householder_2, householder_yz.y, vs[2].z));
		float simplex_volume = fabsf(c0.x * c1.y - c1.x * c0.y);
		float dot_0_2_plus_1_2 = dot_0_2 + dot_1_2;
		float one_plus_dot_0_1 = 1.0f + dot_0_1;
		float tangent = simplex_volume / (one_plus_dot_0_1 + dot_0_2_plus_1_2);
		float triangle_solid_angle = 2.0f * psa_positive_atan(tangent, biased);
		p->solid_angle += triangle_solid_angle;
		p->fan_solid_angles[i] = p->solid_angle;
		p->triangle_parameters[i] = mk3(simplex_volume, dot_0_2_plus_1_2, one_plus_dot_0_1);
	}
}

/* polygon_sampling.glsl:194-225 */
static inline v3 rw_sample_sa(const rw_sa_t* p, v2 rnd, uint32_t maxp) {
	float target_solid_angle = p->solid_angle * rnd.x;
	float subtriangle_solid_angle = target_solid_angle;
	v3 parameters = p->triangle_parameters[0];
	v3 vs[3] = { p->vertex_dirs[1], p->vertex_dirs[0], p->vertex_dirs[2] };
	for (uint32_t i = 0; i != maxp - 3; ++i) {
		if (i + 3 >= p->vertex_count || p->fan_solid_angles[i] >= target_solid_angle) break;
		subtriangle_solid_angle = target_solid_angle - p->fan_solid_angles[i];
		vs[0] = p->vertex_dirs[i + 2];
		vs[2] = p->vertex_dirs[i + 3];
		parameters = p->triangle_parameters[i + 1];
	}
	v2 cs = mk2(vkr_cos(0.5f * subtriangle_solid_angle), vkr_sin(0.5f * subtriangle_solid_angle));
	v3 offset = rw_lin2(parameters.x * cs.x - parameters.y * cs.y, vs[0], parameters.z * cs.y, vs[2]);
	float f = 2.0f * (dot3(vs[0], offset) / dot3(offset, offset));
	v3 new_vertex_2 = mk3(fmaf(f, offset.x, -vs[0].x), fmaf(f, offset.y, -vs[0].y), fmaf(f, offset.z, -vs[0].z));
	float s2 = dot3(vs[1], new_vertex_2);
	float s = psa_mix_fma(1.0f, s2, rnd.y);
	float denominator = fmaf(-s2, s2, 1.0f);
	float t_normed = sqrtf(fmaf(-s, s, 1.0f) / denominator);
	t_normed = (denominator > 0.0f) ? t_normed : rnd.y;
	return rw_lin2(fmaf(-t_normed, s2, s), vs[1], t_normed, new_vertex_2);
}

/* ---- polygon_sampling_related_work.glsl:311-319 (Hart, bilinear) */
typedef struct {
	rw_sa_t polygon;
	float density_0;
	v2 density_1;
} rw_bilinear_hart_t;

/* :327-354 */
static inline void rw_prepare_bilinear_hart(rw_bilinear_hart_t* h, uint32_t vertex_count, const v3* vertices, uint32_t maxp, int biased) {
	rw_prepare_sa(&h->polygon, vertex_count, vertices, mk3(0.0f, 0.0f, 0.0f), maxp, biased);
	h->density_0 = vkr_max(0.0f, h->polygon.vertex_dirs[0].z);
	h->density_1.x = vkr_max(0.0f, h->polygon.vertex_dirs[1].z);
	h->density_1.y = h->polygon.vertex_dirs[2].z;
	for (uint32_t i = 3; i < maxp; ++i)
		h->density_1.y = (i < vertex_count) ? h->polygon.vertex_dirs[i].z : h->density_1.y;
	h->density_1.y = vkr_max(0.0f, h->density_1.y);
	float density_sum = 2.0f * h->density_0 + h->density_1.x + h->density_1.y;
	float normalization = 4.0f / (h->polygon.solid_angle * density_sum);
	h->density_0 *= normalization;
	h->density_1 = scale2(h->density_1, normalization);
	float inv_solid_angle = 1.0f / h->polygon.solid_angle;
	h->density_0 = (density_sum <= 0.0f) ? inv_solid_angle : h->density_0;
	h->density_1 = (density_sum <= 0.0f) ? mk2(inv_solid_angle, inv_solid_angle) : h->density_1;
}

/* :360-374 */
static inline float rw_linear_warp(float random_number, float density_0, float density_1) {
	float lerped_density_sq = psa_mix_fma(density_0 * density_0, density_1 * density_1, random_number);
	float divisor = density_0 + sqrtf(lerped_density_sq);
	return random_number * (density_0 + density_1) / divisor;
}

/* :385-395 */
static inline v3 rw_sample_bilinear_hart(float* out_density, const rw_bilinear_hart_t* h, v2 rnd, uint32_t maxp) {
	rnd.y = rw_linear_warp(rnd.y, 2.0f * h->density_0, dot2(h->density_1, mk2(1.0f, 1.0f)));
	float density_0 = psa_mix_fma(h->density_0, h->density_1.x, rnd.y);
	float density_1 = psa_mix_fma(h->density_0, h->density_1.y, rnd.y);
	rnd.x = rw_linear_warp(rnd.x, density_0, density_1);
	*out_density = psa_mix_fma(density_0, density_1, rnd.x);
	return rw_sample_sa(&h->polygon, rnd, maxp);
}

/* ---- cubic_solver.glsl:29-76. coeffs = (c0, c1, c2, c3) */
static inline int rw_solve_cubic(float out_roots[3], float c0, float c1, float c2, float c3) {
	c0 /= c3; c1 /= c3; c2 /= c3;
	c1 /= 3.0f; c2 /= 3.0f;
	float delta0 = fmaf(-c2, c2, c1);
	float delta1 = fmaf(-c1, c2, c0);
	float delta2 = c2 * c0 - c1 * c1;
	float discriminant = 4.0f * delta0 * delta2 - delta1 * delta1;
	float sqrt_abs_discriminant = sqrtf(fabsf(discriminant));
	float depressed0 = fmaf(-2.0f * c2, delta0, delta1), depressed1 = delta0;
	if (discriminant >= 0.0f) {
		float theta = vkr_atan2(sqrt_abs_discriminant, -depressed0) * (1.0f / 3.0f);
		float cr0 = vkr_cos(theta), cr1 = vkr_sin(theta);
		const float sqrt_075 = sqrtf(0.75f);
		float r0 = cr0;
		float r1 = fmaf(-sqrt_075, cr1, -0.5f * cr0);
		float r2 = fmaf(+sqrt_075, cr1, -0.5f * cr0);
		float scale = 2.0f * sqrtf(-depressed1);
		out_roots[0] = fmaf(scale, r0, -c2);
		out_roots[1] = fmaf(scale, r1, -c2);
		out_roots[2] = fmaf(scale, r2, -c2);
		return 1;
	}
	else {
		float signed_sqrt_discriminant = (depressed0 < 0.0f) ? sqrt_abs_discriminant : -sqrt_abs_discriminant;
		float quadratic_root = 0.5f * (signed_sqrt_discriminant - depressed0);
		float cube_root_0 = vkr_pow(fabsf(quadratic_root), 1.0f / 3.0f);
		cube_root_0 = (quadratic_root < 0.0f) ? -cube_root_0 : cube_root_0;
		float cube_root_1 = -depressed1 / cube_root_0;
		float cubic_root = cube_root_0 + cube_root_1;
		out_roots[0] = cubic_root - c2;
		return 0;
	}
}

/* ---- polygon_sampling_related_work.glsl:400-412 (Hart, biquadratic) */
typedef struct {
	rw_sa_t polygon;
	float density_0;
	v3 density_1, density_2;
} rw_biquadratic_hart_t;

static inline float rw_get3(v3 a, int i) { return (i == 0) ? a.x : ((i == 1) ? a.y : a.z); }
static inline void rw_set3(v3* a, int i, float f) { if (i == 0) a->x = f; else if (i == 1) a->y = f; else a->z = f; }

/* :417-464 */
static inline void rw_prepare_biquadratic_hart(rw_biquadratic_hart_t* h, uint32_t vertex_count, const v3* vertices, uint32_t maxp, int biased) {
	rw_prepare_sa(&h->polygon, vertex_count, vertices, mk3(0.0f, 0.0f, 0.0f), maxp, biased);
	v3 last_vertex = h->polygon.vertex_dirs[2];
	for (uint32_t i = 3; i < maxp; ++i)
		last_vertex = (i < vertex_count) ? h->polygon.vertex_dirs[i] : last_vertex;
	v3 vertex_0 = h->polygon.vertex_dirs[0];
	h->density_0 = vkr_max(0.0f, vertex_0.z);
	h->density_2.x = vkr_max(0.0f, h->polygon.vertex_dirs[1].z);
	h->density_2.z = vkr_max(0.0f, last_vertex.z);
	v3 sample_2_1 = rw_sample_sa(&h->polygon, mk2(0.5f, 1.0f), maxp);
	h->density_2.y = vkr_max(0.0f, sample_2_1.z);
	v3 far_vertices[3] = { vertex_0, sample_2_1, last_vertex };
	for (int i = 0; i != 3; ++i) {
		float s2 = dot3(vertex_0, far_vertices[i]);
		float s = fmaf(0.5f, s2, 0.5f);
		float t = sqrtf(vkr_max(0.0f, fmaf(-s, s, 1.0f)));
		float t_axis_z = fmaf(-s2, vertex_0.z, far_vertices[i].z);
		float normalization_t_axis = vkr_rsqrt(2.0f * fmaf(-s2, s2, 1.0f));
		float sample_1_i_z = s * vertex_0.z + (t * normalization_t_axis) * t_axis_z;
		rw_set3(&h->density_1, i, vkr_max(0.0f, sample_1_i_z));
	}
	const v3 ones = mk3(1.0f, 1.0f, 1.0f);
	float density_sum = 3.0f * h->density_0 + dot3(h->density_1, ones) + dot3(h->density_2, ones);
	float normalization = 9.0f / (h->polygon.solid_angle * density_sum);
	h->density_0 *= normalization;
	h->density_1 = scale3(h->density_1, normalization);
	h->density_2 = scale3(h->density_2, normalization);
	float inv_solid_angle = 1.0f / h->polygon.solid_angle;
	h->density_0 = (density_sum <= 0.0f) ? inv_solid_angle : h->density_0;
	h->density_1 = (density_sum <= 0.0f) ? mk3(inv_solid_angle, inv_solid_angle, inv_solid_angle) : h->density_1;
	h->density_2 = (density_sum <= 0.0f) ? mk3(inv_solid_angle, inv_solid_angle, inv_solid_angle) : h->density_2;
}

/* :471-493 */
static inline float rw_quadratic_warp(float random_number, float density_0, float density_1, float density_2) {
	float q0 = density_0, q1 = 2.0f * (density_1 - density_0), q2 = density_0 - 2.0f * density_1 + density_2;
	float c1 = q0, c2 = 0.5f * q1, c3 = (1.0f / 3.0f) * q2;
	random_number *= dot3(mk3(c1, c2, c3), mk3(1.0f, 1.0f, 1.0f));
	float c0 = -random_number;
	float roots[3] = { 0.0f, 0.0f, 0.0f };
	if (rw_solve_cubic(roots, c0, c1, c2, c3)) {
		float result = roots[0];
		result = (roots[1] >= 0.0f && roots[1] <= 1.0f) ? roots[1] : result;
		result = (roots[2] >= 0.0f && roots[2] <= 1.0f) ? roots[2] : result;
		return result;
	}
	return roots[0];
}

/* :499-503 */
static inline float rw_quadratic_bezier(float b_0_0, float b_0_1, float b_0_2, float location) {
	float b_1_0 = psa_mix_fma(b_0_0, b_0_1, location);
	float b_1_1 = psa_mix_fma(b_0_1, b_0_2, location);
	return psa_mix_fma(b_1_0, b_1_1, location);
}

/* :508-520 */
static inline v3 rw_sample_biquadratic_hart(float* out_density, const rw_biquadratic_hart_t* h, v2 rnd, uint32_t maxp) {
	const v3 ones = mk3(1.0f, 1.0f, 1.0f);
	rnd.y = rw_quadratic_warp(rnd.y, 3.0f * h->density_0, dot3(h->density_1, ones), dot3(h->density_2, ones));
	float density_0 = rw_quadratic_bezier(h->density_0, h->density_1.x, h->density_2.x, rnd.y);
	float density_1 = rw_quadratic_bezier(h->density_0, h->density_1.y, h->density_2.y, rnd.y);
	float density_2 = rw_quadratic_bezier(h->density_0, h->density_1.z, h->density_2.z, rnd.y);
	rnd.x = rw_quadratic_warp(rnd.x, density_0, density_1, density_2);
	*out_density = rw_quadratic_bezier(density_0, density_1, density_2, rnd.x);
	return rw_sample_sa(&h->polygon, rnd, maxp);
}

/* ---- :525-540 (Arvo, projected solid angle) */
typedef struct {
	float cdf_factor;
	v2 length_coeffs;
	v2 elevations;
} rw_edge_arvo_t;

/* :551-576 */
typedef struct {
	uint32_t vertex_count;
	float vertex_azimuths[RW_MAXP];
	rw_edge_arvo_t edges[RW_MAXP];
	rw_edge_arvo_t inner_edge_0;
	float sector_projected_solid_angles[RW_MAXP];
	float projected_solid_angle;
} rw_psa_arvo_t;

/* :582-612 */
static inline rw_edge_arvo_t rw_prepare_edge_arvo(v3 vertex_0, v3 vertex_1) {
	rw_edge_arvo_t edge;
	v3 normal_a = normalize3(cross3(vertex_0, vertex_1));
	edge.cdf_factor = 0.5f * normal_a.z;
	v3 ccw_vertex = (edge.cdf_factor > 0.0f) ? vertex_0 : vertex_1;
	v2 normal_c = psa_rotate_90(normalize2(mk2(ccw_vertex.x, ccw_vertex.y)));
	float cos_beta = -dot2(mk2(normal_a.x, normal_a.y), normal_c);
	float sin_beta_sq = fmaf(-cos_beta, cos_beta, 1.0f);
	float csc_beta = vkr_rsqrt(vkr_max(0.0f, sin_beta_sq));
	float csc_c = vkr_rsqrt(vkr_max(0.0f, fmaf(-ccw_vertex.z, ccw_vertex.z, 1.0f)));
	edge.length_coeffs.x = sin_beta_sq;
	edge.length_coeffs.y = dot2(mk2(normal_a.x, normal_a.y), psa_rotate_90(normal_c)) * cos_beta;
	edge.length_coeffs = scale2(edge.length_coeffs, csc_beta * csc_c);
	edge.elevations.x = ccw_vertex.z;
	edge.elevations.y = cross3(ccw_vertex, normal_a).z;
	edge.elevations.y = (edge.cdf_factor > 0.0f) ? -edge.elevations.y : edge.elevations.y;
	return edge;
}

/* :624-638 */
static inline float rw_edge_psa_in_sector_arvo(const rw_edge_arvo_t* edge, float relative_azimuth_0, float relative_azimuth_1) {
	v2 dir_0 = mk2(vkr_cos(relative_azimuth_0), vkr_sin(relative_azimuth_0));
	v2 point_0 = mk2(dot2(edge->length_coeffs, dir_0), dir_0.y);
	v2 dir_1 = mk2(vkr_cos(relative_azimuth_1), vkr_sin(relative_azimuth_1));
	v2 point_1 = mk2(dot2(edge->length_coeffs, dir_1), dir_1.y);
	v2 rotated_point = mk2(point_0.x * point_1.x + point_0.y * point_1.y, point_0.x * point_1.y - point_0.y * point_1.x);
	float length = psa_positive_atan(fabsf(rotated_point.y) / rotated_point.x, 0);
	return edge->cdf_factor * length;
}

/* :644-668 */
static inline v2 rw_edge_psa_in_sector_derivative_arvo(const rw_edge_arvo_t* edge, float relative_azimuth_0, float relative_azimuth_1) {
	v2 dir_0 = mk2(vkr_cos(relative_azimuth_0), vkr_sin(relative_azimuth_0));
	v2 point_0 = mk2(dot2(edge->length_coeffs, dir_0), dir_0.y);
	v2 dir_1 = mk2(vkr_cos(relative_azimuth_1), vkr_sin(relative_azimuth_1));
	v2 point_1 = mk2(dot2(edge->length_coeffs, dir_1), dir_1.y);
	v2 rotated_point = mk2(point_0.x * point_1.x + point_0.y * point_1.y, point_0.x * point_1.y - point_0.y * point_1.x);
	float quotient = fabsf(rotated_point.y) / rotated_point.x;
	float length = psa_positive_atan(quotient, 0);
	v2 dir_1_deriv = psa_rotate_90(dir_1);
	v2 point_1_deriv = mk2(dot2(edge->length_coeffs, dir_1_deriv), dir_1_deriv.y);
	v2 rotated_point_deriv = mk2(point_0.x * point_1_deriv.x + point_0.y * point_1_deriv.y, point_0.x * point_1_deriv.y - point_0.y * point_1_deriv.x);
	float quotient_derivative = (rotated_point_deriv.y * rotated_point.x - rotated_point.y * rotated_point_deriv.x) / (rotated_point.x * rotated_point.x);
	quotient_derivative = (rotated_point.y < 0.0f) ? (-quotient_derivative) : quotient_derivative;
	float length_deriv = quotient_derivative / fmaf(quotient, quotient, 1.0f);
	return mk2(edge->cdf_factor * length, edge->cdf_factor * length_deriv);
}

/* :674-680 */
static inline float rw_edge_elevation_arvo(const rw_edge_arvo_t* edge, float relative_azimuth) {
	v2 dir = mk2(vkr_cos(relative_azimuth), vkr_sin(relative_azimuth));
	v2 point = mk2(dot2(edge->length_coeffs, dir), dir.y);
	point = normalize2(point);
	return dot2(point, edge->elevations);
}

/* :687-695 */
static inline void rw_compare_and_swap_arvo(rw_psa_arvo_t* p, uint32_t lhs, uint32_t rhs) {
	float lhs_azimuth = p->vertex_azimuths[lhs];
	float flip = p->vertex_azimuths[lhs] - p->vertex_azimuths[rhs];
	p->vertex_azimuths[lhs] = (flip > 0.0f) ? p->vertex_azimuths[rhs] : lhs_azimuth;
	p->vertex_azimuths[rhs] = (flip > 0.0f) ? lhs_azimuth : p->vertex_azimuths[rhs];
	rw_edge_arvo_t lhs_edge = p->edges[lhs];
	p->edges[lhs] = (flip > 0.0f) ? p->edges[rhs] : lhs_edge;
	p->edges[rhs] = (flip > 0.0f) ? lhs_edge : p->edges[rhs];
}

/* :700-769: the same comparator networks as polygon_sampling.glsl:440-505 */
static inline void rw_sort_convex_polygon_vertices_arvo(rw_psa_arvo_t* p, uint32_t maxp) {
	static const unsigned char net5[][2] = {{2,4},{1,3},{1,2},{0,3},{3,4}};
	static const unsigned char net6[][2] = {{3,5},{2,4},{1,5},{0,4},{4,5},{1,3}};
	static const unsigned char net7[][2] = {{2,5},{1,6},{5,6},{3,4},{0,4},{4,6},{1,3},{3,5},{4,5}};
	static const unsigned char net8[][2] = {{2,6},{3,7},{1,5},{0,4},{4,6},{5,7},{6,7},{4,5},{1,3}};
	uint32_t n = p->vertex_count;
	if (n == 3) rw_compare_and_swap_arvo(p, 1, 2);
	else if (maxp >= 4 && n == 4) rw_compare_and_swap_arvo(p, 1, 3);
	else if (maxp >= 5 && n == 5) for (int i = 0; i != 5; ++i) rw_compare_and_swap_arvo(p, net5[i][0], net5[i][1]);
	else if (maxp >= 6 && n == 6) for (int i = 0; i != 6; ++i) rw_compare_and_swap_arvo(p, net6[i][0], net6[i][1]);
	else if (maxp >= 7 && n == 7) for (int i = 0; i != 9; ++i) rw_compare_and_swap_arvo(p, net7[i][0], net7[i][1]);
	else if (maxp >= 8 && n == 8) for (int i = 0; i != 9; ++i) rw_compare_and_swap_arvo(p, net8[i][0], net8[i][1]);
	rw_compare_and_swap_arvo(p, 0, 2);
	if (maxp >= 4 && n >= 4) rw_compare_and_swap_arvo(p, 2, 3);
	rw_compare_and_swap_arvo(p, 0, 1);
}

/* :774-851. vertices is modified (normalised in place) like the GLSL's by-value copy */
static inline void rw_prepare_psa_arvo(rw_psa_arvo_t* p, uint32_t vertex_count, v3* vertices, uint32_t maxp) {
	memset(p, 0, sizeof(*p));
	for (uint32_t i = 0; i != maxp; ++i) vertices[i] = normalize3(vertices[i]);
	p->vertex_count = vertex_count;
	p->inner_edge_0.cdf_factor = 1.0f;
	p->inner_edge_0.length_coeffs = p->inner_edge_0.elevations = mk2(0.0f, 0.0f);
	p->vertex_azimuths[0] = vkr_atan2(vertices[0].y, vertices[0].x);
	p->edges[0] = rw_prepare_edge_arvo(vertices[0], vertices[1]);
	rw_edge_arvo_t previous_edge = p->edges[0];
	for (uint32_t i = 1; i != maxp; ++i) {
		p->vertex_azimuths[i] = vkr_atan2(vertices[i].y, vertices[i].x);
		p->vertex_azimuths[i] -= (p->vertex_azimuths[i] > p->vertex_azimuths[0] + VKR_PI) ? (2.0f * VKR_PI) : 0.0f;
		p->vertex_azimuths[i] += (p->vertex_azimuths[i] < p->vertex_azimuths[0] - VKR_PI) ? (2.0f * VKR_PI) : 0.0f;
		if (i > 2 && i == p->vertex_count) break;
		rw_edge_arvo_t edge = rw_prepare_edge_arvo(vertices[i], vertices[(i + 1) % maxp]);
		p->edges[i] = (edge.cdf_factor >= 0.0f) ? edge : previous_edge;
		p->inner_edge_0 = (previous_edge.cdf_factor < 0.0f && edge.cdf_factor >= 0.0f) ? previous_edge : p->inner_edge_0;
		previous_edge = edge;
	}
	rw_edge_arvo_t edge = p->edges[0];
	p->edges[0] = (edge.cdf_factor >= 0.0f) ? edge : previous_edge;
	p->inner_edge_0 = (previous_edge.cdf_factor < 0.0f && edge.cdf_factor >= 0.0f) ? previous_edge : p->inner_edge_0;
	p->projected_solid_angle = 0.0f;
	if (p->inner_edge_0.cdf_factor > 0.0f) {
		for (uint32_t i = 0; i != maxp; ++i) {
			if (i > 2 && i == p->vertex_count) break;
			p->sector_projected_solid_angles[i] = rw_edge_psa_in_sector_arvo(&p->edges[i], 0.0f, p->vertex_azimuths[(i + 1) % maxp] - p->vertex_azimuths[i]);
			p->projected_solid_angle += p->sector_projected_solid_angles[i];
		}
	}
	else {
		rw_sort_convex_polygon_vertices_arvo(p, maxp);
		rw_edge_arvo_t inner_edge = p->inner_edge_0;
		float inner_azimuth = p->vertex_azimuths[0];
		rw_edge_arvo_t outer_edge;
		memset(&outer_edge, 0, sizeof(outer_edge));
		float outer_azimuth = p->vertex_azimuths[0];
		for (uint32_t i = 0; i != maxp - 1; ++i) {
			if (i > 1 && i + 1 == p->vertex_count) break;
			rw_edge_arvo_t vertex_edge = p->edges[i];
			float vertex_azimuth = p->vertex_azimuths[i];
			if (i == 0) outer_edge = vertex_edge;
			else {
				inner_edge = (vertex_edge.cdf_factor >= 0.0f) ? inner_edge : vertex_edge;
				inner_azimuth = (vertex_edge.cdf_factor >= 0.0f) ? inner_azimuth : vertex_azimuth;
				outer_edge = (vertex_edge.cdf_factor >= 0.0f) ? vertex_edge : outer_edge;
				outer_azimuth = (vertex_edge.cdf_factor >= 0.0f) ? vertex_azimuth : outer_azimuth;
			}
			p->sector_projected_solid_angles[i] = rw_edge_psa_in_sector_arvo(&outer_edge, p->vertex_azimuths[i] - outer_azimuth, p->vertex_azimuths[i + 1] - outer_azimuth);
			p->sector_projected_solid_angles[i] += rw_edge_psa_in_sector_arvo(&inner_edge, p->vertex_azimuths[i] - inner_azimuth, p->vertex_azimuths[i + 1] - inner_azimuth);
			p->projected_solid_angle += p->sector_projected_solid_angles[i];
		}
	}
}

/* :856-868 */
static inline float rw_cubic_interpolation(float sample_x, const float x[4], const float y[4]) {
	float y01 = (y[0] - y[1]) / (x[0] - x[1]);
	float y12 = (y[1] - y[2]) / (x[1] - x[2]);
	float y23 = (y[2] - y[3]) / (x[2] - x[3]);
	float y012 = (y01 - y12) / (x[0] - x[2]);
	float y123 = (y12 - y23) / (x[1] - x[3]);
	float y0123 = (y012 - y123) / (x[0] - x[3]);
	return fmaf(sample_x - x[0], fmaf(sample_x - x[1], fmaf(sample_x - x[2], y0123, y012), y01), y[0]);
}

/* :872-908 (inner_edge == NULL) and :927-968 */
static inline v3 rw_sample_sector_arvo(v2 rnd, float target_psa, const rw_edge_arvo_t* inner_edge, float inner_azimuth, const rw_edge_arvo_t* outer_edge, float outer_azimuth, float azimuth_0, float azimuth_1, uint32_t iteration_count) {
	float azimuths[4] = { azimuth_0, psa_mix_fma(azimuth_0, azimuth_1, 1.0f / 3.0f), psa_mix_fma(azimuth_0, azimuth_1, 2.0f / 3.0f), azimuth_1 };
	float psas[4];
	for (uint32_t i = 0; i != 4; ++i) {
		psas[i] = rw_edge_psa_in_sector_arvo(outer_edge, azimuth_0 - outer_azimuth, azimuths[i] - outer_azimuth);
		if (inner_edge) psas[i] += rw_edge_psa_in_sector_arvo(inner_edge, azimuth_0 - inner_azimuth, azimuths[i] - inner_azimuth);
	}
	float sampled_azimuth = rw_cubic_interpolation(target_psa, psas, azimuths);
	for (uint32_t i = 0; i != iteration_count; ++i) {
		v2 outer_psa = rw_edge_psa_in_sector_derivative_arvo(outer_edge, azimuth_0 - outer_azimuth, sampled_azimuth - outer_azimuth);
		float error, derivative;
		if (inner_edge) {
			v2 inner_psa = rw_edge_psa_in_sector_derivative_arvo(inner_edge, azimuth_0 - inner_azimuth, sampled_azimuth - inner_azimuth);
			error = inner_psa.x + outer_psa.x - target_psa;
			derivative = inner_psa.y + outer_psa.y;
		}
		else {
			error = outer_psa.x - target_psa;
			derivative = outer_psa.y;
		}
		sampled_azimuth -= error / derivative;
		sampled_azimuth = vkr_clamp(sampled_azimuth, azimuth_0, azimuth_1);
	}
	v3 sampled_dir;
	sampled_dir.x = vkr_cos(sampled_azimuth);
	sampled_dir.y = vkr_sin(sampled_azimuth);
	float outer_z = rw_edge_elevation_arvo(outer_edge, sampled_azimuth - outer_azimuth);
	if (inner_edge) {
		float inner_z = rw_edge_elevation_arvo(inner_edge, sampled_azimuth - inner_azimuth);
		sampled_dir.z = sqrtf(psa_mix_fma(inner_z * inner_z, outer_z * outer_z, rnd.y));
	}
	else
		sampled_dir.z = sqrtf(psa_mix_fma(1.0f, outer_z * outer_z, rnd.y));
	float s = sqrtf(fmaf(-sampled_dir.z, sampled_dir.z, 1.0f));
	sampled_dir.x *= s; sampled_dir.y *= s;
	return sampled_dir;
}

/* :973-1030 */
static inline v3 rw_sample_psa_arvo(const rw_psa_arvo_t* p, v2 rnd, uint32_t iteration_count, uint32_t maxp) {
	float target = rnd.x * p->projected_solid_angle;
	float sector_psa = 0.0f;
	rw_edge_arvo_t outer_edge;
	memset(&outer_edge, 0, sizeof(outer_edge));
	float outer_azimuth = 0.0f, azimuth_1 = 0.0f;
	if (p->inner_edge_0.cdf_factor > 0.0f) {
		for (uint32_t i = 0; i != maxp; ++i) {
			if ((i > 2 && i == p->vertex_count) || (i > 0 && target < 0.0f)) break;
			sector_psa = p->sector_projected_solid_angles[i];
			target -= sector_psa;
			outer_edge = p->edges[i];
			outer_azimuth = p->vertex_azimuths[i];
			azimuth_1 = p->vertex_azimuths[(i + 1) % maxp];
		}
		azimuth_1 = (azimuth_1 < outer_azimuth) ? (azimuth_1 + 2.0f * VKR_PI) : azimuth_1;
		target += sector_psa;
		rnd.x = target / sector_psa;
		rnd.x = vkr_clamp(rnd.x, 0.0f, 1.0f);
		return rw_sample_sector_arvo(rnd, target, NULL, 0.0f, &outer_edge, outer_azimuth, outer_azimuth, azimuth_1, iteration_count);
	}
	else {
		rw_edge_arvo_t inner_edge = p->inner_edge_0;
		float inner_azimuth = p->vertex_azimuths[0];
		float azimuth_0 = 0.0f;
		for (uint32_t i = 0; i != maxp - 1; ++i) {
			if ((i > 1 && i + 1 == p->vertex_count) || (i > 0 && target < 0.0f)) break;
			sector_psa = p->sector_projected_solid_angles[i];
			target -= sector_psa;
			rw_edge_arvo_t vertex_edge = p->edges[i];
			float vertex_azimuth = p->vertex_azimuths[i];
			if (i == 0) {
				outer_edge = vertex_edge;
				outer_azimuth = vertex_azimuth;
			}
			else {
				inner_edge = (vertex_edge.cdf_factor >= 0.0f) ? inner_edge : vertex_edge;
				inner_azimuth = (vertex_edge.cdf_factor >= 0.0f) ? inner_azimuth : vertex_azimuth;
				outer_edge = (vertex_edge.cdf_factor >= 0.0f) ? vertex_edge : outer_edge;
				outer_azimuth = (vertex_edge.cdf_factor >= 0.0f) ? vertex_azimuth : outer_azimuth;
			}
			azimuth_0 = p->vertex_azimuths[i];
			azimuth_1 = p->vertex_azimuths[i + 1];
		}
		target += sector_psa;
		rnd.x = target / sector_psa;
		rnd.x = vkr_clamp(rnd.x, 0.0f, 1.0f);
		return rw_sample_sector_arvo(rnd, target, &inner_edge, inner_azimuth, &outer_edge, outer_azimuth, azimuth_0, azimuth_1, iteration_count);
	}
}

/* :1035-1087: backward error and backward error times projected solid angle of a sample of rw_sample_psa_arvo() */
static inline v2 rw_psa_arvo_sampling_error(const rw_psa_arvo_t* p, v2 rnd, v3 sampled_dir, uint32_t maxp) {
	float target = rnd.x * p->projected_solid_angle;
	if (p->inner_edge_0.cdf_factor > 0.0f) return mk2(0.0f, 0.0f);
	rw_edge_arvo_t outer_edge;
	memset(&outer_edge, 0, sizeof(outer_edge));
	rw_edge_arvo_t inner_edge = p->inner_edge_0;
	float inner_azimuth = p->vertex_azimuths[0], outer_azimuth = 0.0f, sector_psa = 0.0f, azimuth_0 = 0.0f;
	for (uint32_t i = 0; i != maxp - 1; ++i) {
		if ((i > 1 && i + 1 == p->vertex_count) || (i > 0 && target < 0.0f)) break;
		sector_psa = p->sector_projected_solid_angles[i];
		target -= sector_psa;
		rw_edge_arvo_t vertex_edge = p->edges[i];
		float vertex_azimuth = p->vertex_azimuths[i];
		if (i == 0) {
			outer_edge = vertex_edge;
			outer_azimuth = vertex_azimuth;
		}
		else {
			inner_edge = (vertex_edge.cdf_factor >= 0.0f) ? inner_edge : vertex_edge;
			inner_azimuth = (vertex_edge.cdf_factor >= 0.0f) ? inner_azimuth : vertex_azimuth;
			outer_edge = (vertex_edge.cdf_factor >= 0.0f) ? vertex_edge : outer_edge;
			outer_azimuth = (vertex_edge.cdf_factor >= 0.0f) ? vertex_azimuth : outer_azimuth;
		}
		azimuth_0 = p->vertex_azimuths[i];
	}
	target += sector_psa;
	float sampled_azimuth = vkr_atan2(sampled_dir.y, sampled_dir.x);
	float outer_psa = rw_edge_psa_in_sector_derivative_arvo(&outer_edge, azimuth_0 - outer_azimuth, sampled_azimuth - outer_azimuth).x;
	float inner_psa = rw_edge_psa_in_sector_derivative_arvo(&inner_edge, azimuth_0 - inner_azimuth, sampled_azimuth - inner_azimuth).x;
	float sampled_psa = outer_psa + inner_psa;
	return mk2((target - sampled_psa) / p->projected_solid_angle, target - sampled_psa);
}

#endif
