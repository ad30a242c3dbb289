/* oracle/psa_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Scalar fp32 restatement of the reference's projected-solid-angle polygon
 * sampling and polygon clipping:
 *   src/shaders/polygon_sampling.glsl:104-111, 183-185, 261-805
 *   src/shaders/polygon_clipping.glsl:19-25, 35-225
 * Every function cites the GLSL lines it follows. Expression order and the
 * placement of fma() follow the GLSL text; plain a*b+c is NOT contracted
 * (build with -ffp-contract=off). Elementary functions come from vkr_math.h.
 *
 * Parity status: the reference ships no golden vectors for this path (SURVEY 4,
 * 8c). This file is pinned two ways: (1) analytic known-answer tests in
 * tests/test_oracle_kats.py, (2) bit-for-bit comparison against the reference's
 * own GLSL sources compiled as C++ (oracle/_ref, see oracle/glsl_compat/), run
 * wherever /root/reference exists, with vectors frozen in tests/golden/.
 */
#ifndef VKR_PSA_ORACLE_H
#define VKR_PSA_ORACLE_H
#include "vkr_math.h"
#include "clip_table.inc"

#define PSA_MAXP 8 /* largest supported MAX_POLYGON_VERTEX_COUNT (7 light vertices + 1) */

typedef struct {
	uint32_t vertex_count;
	v2 vertices[PSA_MAXP];
	v2 ellipses[PSA_MAXP];
	v2 inner_ellipse_0;
	float sector_projected_solid_angles[PSA_MAXP];
	float projected_solid_angle;
} psa_polygon_t; /* projected_solid_angle_polygon_t, polygon_sampling.glsl:230-252 */

/* polygon_clipping.glsl:19-25 */
static inline v3 psa_iz0(v3 lhs, v3 rhs) {
	float lerp_factor = lhs.z / (lhs.z - rhs.z);
	return mk3(
		fmaf(lerp_factor, rhs.x, fmaf(-lerp_factor, lhs.x, lhs.x)),
		fmaf(lerp_factor, rhs.y, fmaf(-lerp_factor, lhs.y, lhs.y)),
		0.0f);
}

/* polygon_clipping.glsl:35-225. maxp = MAX_POLYGON_VERTEX_COUNT. The per-case
   vertex order comes from clip_table.inc (derived from the reference switch). */
static inline uint32_t psa_clip_polygon(uint32_t vertex_count, v3* v, uint32_t maxp) {
	uint32_t bits = 0;
	for (uint32_t i = 0; i + 1 < maxp; ++i)
		if (v[i].z > 0.0f && i < vertex_count) bits |= 1u << i;
	if (vertex_count < 3 || vertex_count > 7) return 0;
	const unsigned char* e = VKR_CLIP_TABLE[vertex_count - 3][bits];
	uint32_t vc = e[0];
	if (vc == 0) return 0;
	v3 in[PSA_MAXP];
	for (uint32_t i = 0; i < vertex_count; ++i) in[i] = v[i];
	for (uint32_t s = 0; s < vc; ++s) {
		uint32_t code = e[1 + s];
		if (code < 8) v[s] = in[code];
		else {
			uint32_t k = code - 8;
			v[s] = psa_iz0(in[k], in[(k + 1) % vertex_count]);
		}
	}
	if (vc < maxp) v[vc] = v[0];
	return vc;
}

/* polygon_sampling.glsl:104-111 (unbiased variant) and :83-97 (biased variant) */
static inline float psa_fast_positive_atan(float y) {
	float rx, ry, rz;
	rx = (fabsf(y) > 1.0f) ? (1.0f / fabsf(y)) : fabsf(y);
	ry = rx * rx;
	rz = fmaf(ry, 0.02083509974181652f, -0.08513300120830536f);
	rz = fmaf(ry, rz, 0.18014100193977356f);
	rz = fmaf(ry, rz, -0.3302994966506958f);
	ry = fmaf(ry, rz, 0.9998660087585449f);
	rz = fmaf(-2.0f * ry, rx, VKR_HALF_PI);
	rz = (fabsf(y) > 1.0f) ? rz : 0.0f;
	rx = fmaf(rx, ry, rz);
	return (y < 0.0f) ? (VKR_PI - rx) : rx;
}
static inline float psa_positive_atan(float tangent, int biased) {
	if (biased) return psa_fast_positive_atan(tangent);
	float offset = (tangent < 0.0f) ? VKR_PI : 0.0f;
	return vkr_atan(tangent) + offset;
}

/* :183-185 */
static inline float psa_mix_fma(float x, float y, float a) { return fmaf(a, y, fmaf(-a, x, x)); }

/* :261-268 */
static inline float psa_kahan(float a, float b, float c, float d) {
	float cd = c * d;
	float error = fmaf(c, d, -cd);
	float result = fmaf(a, b, -cd);
	return result - error;
}
/* :273-279 */
static inline v3 psa_cross_stable(v3 lhs, v3 rhs) {
	return mk3(
		psa_kahan(lhs.y, rhs.z, lhs.z, rhs.y),
		psa_kahan(lhs.z, rhs.x, lhs.x, rhs.z),
		psa_kahan(lhs.x, rhs.y, lhs.y, rhs.x));
}
/* :284-286 */
static inline v2 psa_rotate_90(v2 a) { return mk2(-a.y, a.x); }
/* :292-299 */
static inline int psa_is_inner_ellipse(v2 ellipse) { return (f2u(ellipse.x) & 0x80000000u) != 0; }
/* :304-306 */
static inline int psa_is_central_case(const psa_polygon_t* p) { return p->inner_ellipse_0.x > 0.0f; }

/* :317-326 */
static inline v2 psa_ellipse_from_edge(v3 vertex_0, v3 vertex_1) {
	v3 normal = psa_cross_stable(vertex_0, vertex_1);
	float scaling = 1.0f / normal.z;
	scaling = psa_is_inner_ellipse(mk2(normal.x, normal.y)) ? -scaling : scaling;
	v2 ellipse = mk2(normal.x * scaling, normal.y * scaling);
	ellipse.x = (normal.z != 0.0f) ? ellipse.x : INFINITY;
	return ellipse;
}
/* :332-334 */
static inline v2 psa_ellipse_transform(v2 ellipse, v2 point) {
	float d = dot2(ellipse, point);
	return mk2(fmaf(d, ellipse.x, point.x), fmaf(d, ellipse.y, point.y));
}
/* :340-342 */
static inline float psa_get_ellipse_det(v2 e) { return fmaf(e.x, e.x, fmaf(e.y, e.y, 1.0f)); }
/* :346-348 */
static inline float psa_get_ellipse_rsqrt_det(v2 e) { return vkr_rsqrt(psa_get_ellipse_det(e)); }
/* :351-355 */
static inline float psa_get_ellipse_direction_factor_rsq(v2 ellipse, v2 dir) {
	float ellipse_dot_dir = dot2(ellipse, dir);
	float dir_dot_dir = dot2(dir, dir);
	return fmaf(ellipse_dot_dir, ellipse_dot_dir, dir_dot_dir);
}
/* :363-365 */
static inline float psa_get_ellipse_direction_factor(v2 ellipse, v2 dir) {
	return vkr_rsqrt(psa_get_ellipse_direction_factor_rsq(ellipse, dir));
}
/* :369-372 */
static inline float psa_get_ellipse_normalized_direction_factor(v2 ellipse, v2 normalized_dir) {
	float ellipse_dot_dir = dot2(ellipse, normalized_dir);
	return vkr_rsqrt(fmaf(ellipse_dot_dir, ellipse_dot_dir, 1.0f));
}
/* :377-382 */
static inline float psa_area_between_from_tangents(float inner_rsqrt_det, float inner_tangent, float outer_rsqrt_det, float outer_tangent, int biased) {
	float inner_area = inner_rsqrt_det * psa_positive_atan(inner_tangent, biased);
	float result = fmaf(outer_rsqrt_det, psa_positive_atan(outer_tangent, biased), -inner_area);
	return (result > 0.0f) ? (0.5f * result) : 0.0f;
}
/* :390-397 */
static inline float psa_area_between_ellipses_in_sector(v2 inner_ellipse, float inner_rsqrt_det, v2 outer_ellipse, float outer_rsqrt_det, v2 dir_0, v2 dir_1, int biased) {
	float det_dirs = vkr_max(+0.0f, dot2(dir_1, psa_rotate_90(dir_0)));
	float inner_dot = inner_rsqrt_det * dot2(dir_0, psa_ellipse_transform(inner_ellipse, dir_1));
	float outer_dot = outer_rsqrt_det * dot2(dir_0, psa_ellipse_transform(outer_ellipse, dir_1));
	return psa_area_between_from_tangents(inner_rsqrt_det, det_dirs / inner_dot, outer_rsqrt_det, det_dirs / outer_dot, biased);
}
/* :405-412 */
static inline float psa_ellipse_area_in_sector(v2 ellipse, v2 dir_0, v2 dir_1, int biased) {
	float ellipse_rsqrt_det = psa_get_ellipse_rsqrt_det(ellipse);
	float det_dirs = vkr_max(+0.0f, dot2(dir_1, psa_rotate_90(dir_0)));
	float ellipse_dot = ellipse_rsqrt_det * dot2(dir_0, psa_ellipse_transform(ellipse, dir_1));
	float area = 0.5f * ellipse_rsqrt_det * psa_positive_atan(det_dirs / ellipse_dot, biased);
	return (ellipse_rsqrt_det > 0.0f) ? area : 0.0f;
}

/* :421-435 */
static inline void psa_compare_and_swap(psa_polygon_t* polygon, uint32_t lhs, uint32_t rhs) {
	v2 lhs_copy = polygon->vertices[lhs];
	float normal_z = psa_kahan(lhs_copy.x, -polygon->vertices[rhs].y, lhs_copy.y, -polygon->vertices[rhs].x);
	int swap = (normal_z == 0.0f) ? (isinf(polygon->ellipses[rhs].x) != 0) : (normal_z > 0.0f);
	polygon->vertices[lhs] = swap ? polygon->vertices[rhs] : lhs_copy;
	polygon->vertices[rhs] = swap ? lhs_copy : polygon->vertices[rhs];
	lhs_copy = polygon->ellipses[lhs];
	polygon->ellipses[lhs] = swap ? polygon->ellipses[rhs] : lhs_copy;
	polygon->ellipses[rhs] = swap ? lhs_copy : polygon->ellipses[rhs];
}

/* :440-505; the comparator pairs are the reference's sorting networks */
static inline void psa_sort_convex_polygon_vertices(psa_polygon_t* p, uint32_t maxp) {
	static const unsigned char net5[][2] = {{2,4},{1,3},{1,2},{0,3},{3,4}};
	static const unsigned char net6[][2] = {{3,5},{2,4},{1,5},{0,4},{4,5},{1,3}};
	static const unsigned char net7[][2] = {{2,5},{1,6},{5,6},{3,4},{0,4},{4,6},{1,3},{3,5},{4,5}};
	static const unsigned char net8[][2] = {{2,6},{3,7},{1,5},{0,4},{4,6},{5,7},{6,7},{4,5},{1,3}};
	uint32_t n = p->vertex_count;
	if (n == 3) psa_compare_and_swap(p, 1, 2);
	else if (maxp >= 4 && n == 4) psa_compare_and_swap(p, 1, 3);
	else if (maxp >= 5 && n == 5) for (int i = 0; i != 5; ++i) psa_compare_and_swap(p, net5[i][0], net5[i][1]);
	else if (maxp >= 6 && n == 6) for (int i = 0; i != 6; ++i) psa_compare_and_swap(p, net6[i][0], net6[i][1]);
	else if (maxp >= 7 && n == 7) for (int i = 0; i != 9; ++i) psa_compare_and_swap(p, net7[i][0], net7[i][1]);
	else if (maxp >= 8 && n == 8) for (int i = 0; i != 9; ++i) psa_compare_and_swap(p, net8[i][0], net8[i][1]);
	psa_compare_and_swap(p, 0, 2);
	if (maxp >= 4 && n >= 4) psa_compare_and_swap(p, 2, 3);
	psa_compare_and_swap(p, 0, 1);
}

/* :521-589 */
static inline void psa_prepare(psa_polygon_t* polygon, uint32_t vertex_count, const v3* vertices, uint32_t maxp, int biased) {
	memset(polygon, 0, sizeof(*polygon));
	polygon->vertex_count = vertex_count;
	polygon->inner_ellipse_0 = mk2(1.0f, 0.0f);
	polygon->vertices[0] = mk2(vertices[0].x, vertices[0].y);
	polygon->ellipses[0] = psa_ellipse_from_edge(vertices[0], vertices[1]);
	v2 previous_ellipse = polygon->ellipses[0];
	for (uint32_t i = 1; i != maxp; ++i) {
		polygon->vertices[i] = mk2(vertices[i].x, vertices[i].y);
		if (i > 2 && i == polygon->vertex_count) break;
		v2 ellipse = psa_ellipse_from_edge(vertices[i], vertices[(i + 1) % maxp]);
		int ellipse_inner = psa_is_inner_ellipse(ellipse);
		polygon->ellipses[i] = ellipse_inner ? previous_ellipse : ellipse;
		polygon->inner_ellipse_0 = (psa_is_inner_ellipse(previous_ellipse) && !ellipse_inner) ? previous_ellipse : polygon->inner_ellipse_0;
		previous_ellipse = ellipse;
	}
	v2 ellipse = polygon->ellipses[0];
	int ellipse_inner = psa_is_inner_ellipse(ellipse);
	polygon->ellipses[0] = ellipse_inner ? previous_ellipse : ellipse;
	polygon->inner_ellipse_0 = (psa_is_inner_ellipse(previous_ellipse) && !ellipse_inner) ? previous_ellipse : polygon->inner_ellipse_0;
	polygon->projected_solid_angle = 0.0f;
	if (psa_is_central_case(polygon)) {
		for (uint32_t i = 0; i != maxp; ++i) {
			if (i > 2 && i == polygon->vertex_count) break;
			polygon->sector_projected_solid_angles[i] = psa_ellipse_area_in_sector(polygon->ellipses[i], polygon->vertices[i], polygon->vertices[(i + 1) % maxp], biased);
			polygon->projected_solid_angle += polygon->sector_projected_solid_angles[i];
		}
	}
	else {
		psa_sort_convex_polygon_vertices(polygon, maxp);
		v2 inner_ellipse = polygon->inner_ellipse_0;
		float inner_rsqrt_det = psa_get_ellipse_rsqrt_det(inner_ellipse);
		v2 outer_ellipse = mk2(0.0f, 0.0f);
		float outer_rsqrt_det = 0.0f;
		for (uint32_t i = 0; i != maxp - 1; ++i) {
			if (i > 1 && i + 1 == polygon->vertex_count) break;
			v2 vertex_ellipse = polygon->ellipses[i];
			int vertex_inner = psa_is_inner_ellipse(vertex_ellipse);
			float vertex_rsqrt_det = psa_get_ellipse_rsqrt_det(vertex_ellipse);
			if (i == 0) {
				outer_ellipse = vertex_ellipse;
				outer_rsqrt_det = vertex_rsqrt_det;
			}
			else {
				inner_ellipse = vertex_inner ? vertex_ellipse : inner_ellipse;
				inner_rsqrt_det = vertex_inner ? vertex_rsqrt_det : inner_rsqrt_det;
				outer_ellipse = vertex_inner ? outer_ellipse : vertex_ellipse;
				outer_rsqrt_det = vertex_inner ? outer_rsqrt_det : vertex_rsqrt_det;
			}
			polygon->sector_projected_solid_angles[i] = psa_area_between_ellipses_in_sector(
				inner_ellipse, inner_rsqrt_det, outer_ellipse, outer_rsqrt_det, polygon->vertices[i], polygon->vertices[i + 1], biased);
			polygon->projected_solid_angle += polygon->sector_projected_solid_angles[i];
		}
	}
}

/* :599-611 */
static inline v2 psa_normalize_approx_and_flip(v2 rhs, v2 semi_circle) {
	float scaling = fabsf(rhs.x) + fabsf(rhs.y);
	scaling = u2f(f2u(scaling) ^ 0x7F800000u);
	scaling = (dot2(rhs, semi_circle) >= 0.0f) ? scaling : -scaling;
	return mk2(scaling * rhs.x, scaling * rhs.y);
}

/* :625-630. q is a GLSL mat2 in column-major order: q[col][row]. */
static inline v2 psa_solve_homogeneous_quadratic(const float q[2][2]) {
	float coeff_xy = 0.5f * (q[0][1] + q[1][0]);
	float sqrt_discriminant = sqrtf(vkr_max(0.0f, coeff_xy * coeff_xy - q[0][0] * q[1][1]));
	float scaled_root = fabsf(coeff_xy) + sqrt_discriminant;
	return (coeff_xy >= 0.0f) ? mk2(scaled_root, -q[0][0]) : mk2(q[1][1], scaled_root);
}

/* outerProduct(c, r)[col j][row i] = c[i] * r[j] */
static inline void psa_outer(float m[2][2], v2 c, v2 r) {
	m[0][0] = c.x * r.x; m[0][1] = c.y * r.x;
	m[1][0] = c.x * r.y; m[1][1] = c.y * r.y;
}

/* :645-739 */
static inline v2 psa_sample_sector_between_ellipses(v2 random_numbers, float target_area, v2 inner_ellipse, v2 outer_ellipse, v2 dir_0, v2 dir_1, uint32_t iteration_count, int biased) {
	v2 quad_dirs[3];
	quad_dirs[0] = normalize2(dir_0);
	quad_dirs[2] = normalize2(dir_1);
	quad_dirs[1] = add2(quad_dirs[0], quad_dirs[2]);
	float nf[2][3] = {
		{
			psa_get_ellipse_normalized_direction_factor(inner_ellipse, quad_dirs[0]),
			psa_get_ellipse_direction_factor(inner_ellipse, quad_dirs[1]),
			psa_get_ellipse_normalized_direction_factor(inner_ellipse, quad_dirs[2])
		},
		{
			psa_get_ellipse_normalized_direction_factor(outer_ellipse, quad_dirs[0]),
			psa_get_ellipse_direction_factor(outer_ellipse, quad_dirs[1]),
			psa_get_ellipse_normalized_direction_factor(outer_ellipse, quad_dirs[2])
		}
	};
	float sector_areas[2] = {
		nf[1][0] * nf[1][1] - nf[0][0] * nf[0][1],
		nf[1][1] * nf[1][2] - nf[0][1] * nf[0][2]
	};
	float target_quad_area = psa_mix_fma(-sector_areas[0], sector_areas[1], random_numbers.x);
	quad_dirs[2] = (target_quad_area <= 0.0f) ? quad_dirs[0] : quad_dirs[2];
	nf[0][2] = (target_quad_area <= 0.0f) ? nf[0][0] : nf[0][2];
	nf[1][2] = (target_quad_area <= 0.0f) ? nf[1][0] : nf[1][2];
	target_quad_area += (target_quad_area <= 0.0f) ? sector_areas[0] : -sector_areas[1];
	/* determinant(mat2(a, b)) with columns a,b = a.x*b.y - b.x*a.y */
	target_quad_area *= fabsf(quad_dirs[1].x * quad_dirs[2].y - quad_dirs[2].x * quad_dirs[1].y);
	v2 quad_normals[2] = {
		add2(scale2(quad_dirs[1], nf[0][1]), scale2(quad_dirs[2], nf[0][2])),
		add2(scale2(quad_dirs[1], nf[1][1]), scale2(quad_dirs[2], nf[1][2]))
	};
	quad_normals[0] = psa_ellipse_transform(inner_ellipse, quad_normals[0]);
	quad_normals[1] = psa_ellipse_transform(outer_ellipse, quad_normals[1]);
	float quad_offsets[2] = {
		dot2(quad_normals[0], quad_dirs[1]) * nf[0][1],
		dot2(quad_normals[1], quad_dirs[1]) * nf[1][1]
	};
	float quadratic[2][2], tmp[2][2];
	v2 r90 = psa_rotate_90(quad_dirs[2]);
	psa_outer(quadratic, scale2(r90, quad_offsets[1] * nf[1][2]), quad_normals[0]);
	psa_outer(tmp, add2(scale2(r90, quad_offsets[0] * nf[0][2]), scale2(quad_normals[0], target_quad_area)), quad_normals[1]);
	for (int c = 0; c != 2; ++c) for (int r = 0; r != 2; ++r) quadratic[c][r] -= tmp[c][r];
	v2 current_dir = psa_solve_homogeneous_quadratic(quadratic);

	if (!biased) {
		float acceptable_error = 1.0e-5f;
		iteration_count = (fabsf(random_numbers.x - 0.5f) <= 0.5f - acceptable_error) ? iteration_count : 0;
		float inner_rsqrt_det = psa_get_ellipse_rsqrt_det(inner_ellipse);
		float outer_rsqrt_det = psa_get_ellipse_rsqrt_det(outer_ellipse);
		for (uint32_t i = 0; i != iteration_count; ++i) {
			current_dir = psa_normalize_approx_and_flip(current_dir, quad_dirs[1]);
			v2 inner_dir = psa_ellipse_transform(inner_ellipse, current_dir);
			v2 outer_dir = psa_ellipse_transform(outer_ellipse, current_dir);
			float det_dirs = vkr_max(+0.0f, dot2(current_dir, psa_rotate_90(quad_dirs[0])));
			float error = target_area - psa_area_between_from_tangents(
				inner_rsqrt_det, det_dirs / (inner_rsqrt_det * dot2(quad_dirs[0], inner_dir)),
				outer_rsqrt_det, det_dirs / (outer_rsqrt_det * dot2(quad_dirs[0], outer_dir)), biased);
			psa_outer(quadratic, sub2(inner_dir, outer_dir), psa_rotate_90(current_dir));
			psa_outer(tmp, scale2(inner_dir, 2.0f * error), outer_dir);
			for (int c = 0; c != 2; ++c) for (int r = 0; r != 2; ++r) quadratic[c][r] -= tmp[c][r];
			current_dir = psa_solve_homogeneous_quadratic(quadratic);
		}
	}
	current_dir = (dot2(current_dir, quad_dirs[1]) >= 0.0f) ? current_dir : mk2(-current_dir.x, -current_dir.y);
	float inner_factor = 1.0f / psa_get_ellipse_direction_factor_rsq(inner_ellipse, current_dir);
	float outer_factor = 1.0f / psa_get_ellipse_direction_factor_rsq(outer_ellipse, current_dir);
	float s = sqrtf(psa_mix_fma(inner_factor, outer_factor, random_numbers.y));
	return mk2(current_dir.x * s, current_dir.y * s);
}

/* :749-805 */
static inline v3 psa_sample(const psa_polygon_t* polygon, v2 random_numbers, uint32_t maxp, int biased) {
	float target = random_numbers.x * polygon->projected_solid_angle;
	v2 sampled_xy;
	v2 outer_ellipse = mk2(0.0f, 0.0f);
	v2 dir_0 = mk2(0.0f, 0.0f);
	if (psa_is_central_case(polygon)) {
		for (uint32_t i = 0; i != maxp; ++i) {
			if (i > 0) target -= polygon->sector_projected_solid_angles[i - 1];
			outer_ellipse = polygon->ellipses[i];
			dir_0 = polygon->vertices[i];
			if ((i >= 2 && i + 1 == polygon->vertex_count) || target < polygon->sector_projected_solid_angles[i])
				break;
		}
		float sqrt_det = sqrtf(psa_get_ellipse_det(outer_ellipse));
		float angle = 2.0f * target * sqrt_det;
		float ca = vkr_cos(angle) * sqrt_det;
		float sa = vkr_sin(angle);
		v2 t = psa_rotate_90(psa_ellipse_transform(outer_ellipse, dir_0));
		sampled_xy = mk2(ca * dir_0.x + sa * t.x, ca * dir_0.y + sa * t.y);
		float s = sqrtf(random_numbers.y / psa_get_ellipse_direction_factor_rsq(outer_ellipse, sampled_xy));
		sampled_xy = mk2(sampled_xy.x * s, sampled_xy.y * s);
	}
	else {
		float sector_psa = 0.0f;
		v2 inner_ellipse = polygon->inner_ellipse_0;
		v2 dir_1 = mk2(0.0f, 0.0f);
		for (uint32_t i = 0; i != maxp - 1; ++i) {
			v2 vertex_ellipse = polygon->ellipses[i];
			if (i == 0) outer_ellipse = vertex_ellipse;
			else {
				target -= polygon->sector_projected_solid_angles[i - 1];
				int vertex_inner = psa_is_inner_ellipse(vertex_ellipse);
				inner_ellipse = vertex_inner ? vertex_ellipse : inner_ellipse;
				outer_ellipse = vertex_inner ? outer_ellipse : vertex_ellipse;
			}
			dir_0 = polygon->vertices[i];
			dir_1 = polygon->vertices[i + 1];
			sector_psa = polygon->sector_projected_solid_angles[i];
			if ((i >= 1 && i + 2 == polygon->vertex_count) || target < sector_psa)
				break;
		}
		random_numbers.x = target / sector_psa;
		sampled_xy = psa_sample_sector_between_ellipses(random_numbers, target, inner_ellipse, outer_ellipse, dir_0, dir_1, 2, biased);
	}
	v3 out;
	out.x = sampled_xy.x; out.y = sampled_xy.y;
	out.z = sqrtf(vkr_max(0.0f, fmaf(-sampled_xy.x, sampled_xy.x, fmaf(-sampled_xy.y, sampled_xy.y, 1.0f))));
	return out;
}

/* :823-883: backward error, backward error times projected solid angle, forward error in radians */
static inline v3 psa_sampling_error(const psa_polygon_t* polygon, v2 random_numbers, v3 sampled_dir, uint32_t maxp, int biased) {
	float target = random_numbers.x * polygon->projected_solid_angle;
	if (psa_is_central_case(polygon)) return mk3(0.0f, 0.0f, 0.0f);
	float sector_psa = 0.0f;
	v2 outer_ellipse = mk2(0.0f, 0.0f);
	v2 inner_ellipse = polygon->inner_ellipse_0;
	v2 dir_0 = mk2(0.0f, 0.0f);
	for (uint32_t i = 0; i != maxp - 1; ++i) {
		if ((i > 1 && i + 1 == polygon->vertex_count) || (i > 0 && target < 0.0f)) break;
		sector_psa = polygon->sector_projected_solid_angles[i];
		target -= sector_psa;
		v2 vertex_ellipse = polygon->ellipses[i];
		int vertex_inner = psa_is_inner_ellipse(vertex_ellipse);
		if (i == 0) outer_ellipse = vertex_ellipse;
		else {
			inner_ellipse = vertex_inner ? vertex_ellipse : inner_ellipse;
			outer_ellipse = vertex_inner ? outer_ellipse : vertex_ellipse;
		}
		dir_0 = polygon->vertices[i];
	}
	target += sector_psa;
	v2 sxy = mk2(sampled_dir.x, sampled_dir.y);
	float sampled_psa = psa_area_between_ellipses_in_sector(
		inner_ellipse, psa_get_ellipse_rsqrt_det(inner_ellipse), outer_ellipse, psa_get_ellipse_rsqrt_det(outer_ellipse), dir_0, sxy, biased);
	float scaled_backward_error = target - sampled_psa;
	float backward_error = scaled_backward_error / polygon->projected_solid_angle;
	/* derivative of the sampled direction with respect to the projected solid angle (:860-875); cm = constraint_matrix before the transpose */
	v2 inner_dir = psa_ellipse_transform(inner_ellipse, sxy);
	v2 outer_dir = psa_ellipse_transform(outer_ellipse, sxy);
	float inner_factor = 1.0f / dot2(sxy, inner_dir);
	float outer_factor = 1.0f / dot2(sxy, outer_dir);
	v2 cm0 = scale2(psa_rotate_90(sxy), 0.5f * (inner_factor - outer_factor));
	v2 cm1 = scale2(inner_dir, (1.0f - random_numbers.y) / (inner_factor * inner_factor));
	cm1 = add2(cm1, scale2(outer_dir, random_numbers.y / (outer_factor * outer_factor)));
	/* transpose: t[0] = (cm0.x, cm1.x), t[1] = (cm0.y, cm1.y); determinant(t) = t[0][0] * t[1][1] - t[1][0] * t[0][1] */
	float rcp_det = 1.0f / (cm0.x * cm1.y - cm0.y * cm1.x);
	v3 sample_derivative;
	sample_derivative.x = rcp_det * cm1.y;
	sample_derivative.y = rcp_det * -cm1.x;
	sample_derivative.z = -dot2(sxy, mk2(sample_derivative.x, sample_derivative.y)) / sampled_dir.z;
	float forward_error = sqrtf(dot3(sample_derivative, sample_derivative)) * scaled_backward_error;
	return mk3(backward_error, scaled_backward_error, forward_error);
}

#endif
