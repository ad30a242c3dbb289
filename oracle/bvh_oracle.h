/* oracle/bvh_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * The reference traces shadow rays with VK_KHR_ray_query against a driver-built
 * acceleration structure (src/shaders/shading_pass.frag.glsl:120-138,
 * src/scene.c:142-406). BVH construction and ray/triangle arithmetic live in the
 * un-vendored, un-pinned Vulkan driver -- "parity unpinned" for this
 * part by construction. The oracle therefore DEFINES the predicate:
 *
 *   hit(ray, tri)  = Moeller-Trumbore in fp32, operation order below, no culling
 *                    (scene.c:325 disables facing culls), open interval (tmin,tmax)
 *   occluded(ray)  = OR over all triangles of hit(ray, tri)
 *
 * and evaluates it through an independent binned-SAH BVH2 whose leaf boxes are
 * padded so that the box test is conservative w.r.t. the fp32 predicate; the
 * brute-force OR (oracle_occluded_brute) is the KAT for that.
 */
#ifndef VKR_BVH_ORACLE_H
#define VKR_BVH_ORACLE_H
#include "vkr_math.h"
#include <stdlib.h>

typedef struct {
	float lo[3], hi[3];
	int32_t left;   /* inner: index of first child (second = left+1); leaf: first triangle slot */
	int32_t count;  /* 0 for inner nodes, else triangle count of the leaf */
} obvh_node_t;

typedef struct {
	const float* tris;   /* 9 floats per triangle (v0,v1,v2), NOT reordered */
	uint32_t tri_count;
	uint32_t* order;     /* leaf slots -> triangle index */
	obvh_node_t* nodes;
	uint32_t node_count;
} obvh_t;

/* The shadow/primary-ray triangle predicate. Returns 1 on hit and writes t,u,v. */
static inline int oracle_ray_triangle(const float* tri, v3 o, v3 d, float tmin, float tmax, float* out_t, float* out_u, float* out_v) {
	v3 p0 = mk3(tri[0], tri[1], tri[2]);
	v3 e1 = mk3(tri[3] - tri[0], tri[4] - tri[1], tri[5] - tri[2]);
	v3 e2 = mk3(tri[6] - tri[0], tri[7] - tri[1], tri[8] - tri[2]);
	v3 pv = cross3(d, e2);
	float det = dot3(e1, pv);
	if (det == 0.0f) return 0;
	float inv_det = 1.0f / det;
	v3 tv = sub3(o, p0);
	float u = dot3(tv, pv) * inv_det;
	if (!(u >= 0.0f && u <= 1.0f)) return 0;
	v3 qv = cross3(tv, e1);
	float v = dot3(d, qv) * inv_det;
	if (!(v >= 0.0f && u + v <= 1.0f)) return 0;
	float t = dot3(e2, qv) * inv_det;
	if (!(t > tmin && t < tmax)) return 0;
	*out_t = t; *out_u = u; *out_v = v;
	return 1;
}

static inline int oracle_occluded_brute(const float* tris, uint32_t tri_count, v3 o, v3 d, float tmin, float tmax) {
	float t, u, v;
	for (uint32_t i = 0; i != tri_count; ++i)
		if (oracle_ray_triangle(tris + 9 * i, o, d, tmin, tmax, &t, &u, &v)) return 1;
	return 0;
}

/* Slab test. NaNs (0*inf) drop out of the min/max; tmax is widened by 4 ulp. */
static inline int obvh_ray_box(const obvh_node_t* n, v3 o, v3 inv_d, float tmin, float tmax) {
	float oo[3] = {o.x, o.y, o.z}, id[3] = {inv_d.x, inv_d.y, inv_d.z};
	float t_near = tmin, t_far = tmax;
	for (int a = 0; a != 3; ++a) {
		float t0 = (n->lo[a] - oo[a]) * id[a];
		float t1 = (n->hi[a] - oo[a]) * id[a];
		float tn = (t0 < t1) ? t0 : t1;
		float tf = (t0 > t1) ? t0 : t1;
		t_near = (tn > t_near) ? tn : t_near;
		t_far = (tf < t_far) ? tf : t_far;
	}
	return t_near <= t_far * 1.0000005f;
}

typedef struct { float lo[3], hi[3]; } obox_t;
static inline void obox_empty(obox_t* b) { for (int a = 0; a != 3; ++a) { b->lo[a] = INFINITY; b->hi[a] = -INFINITY; } }
static inline void obox_grow(obox_t* b, const float* p) { for (int a = 0; a != 3; ++a) { if (p[a] < b->lo[a]) b->lo[a] = p[a]; if (p[a] > b->hi[a]) b->hi[a] = p[a]; } }
static inline void obox_merge(obox_t* b, const obox_t* o) { for (int a = 0; a != 3; ++a) { if (o->lo[a] < b->lo[a]) b->lo[a] = o->lo[a]; if (o->hi[a] > b->hi[a]) b->hi[a] = o->hi[a]; } }
static inline float obox_area(const obox_t* b) {
	float dx = b->hi[0] - b->lo[0], dy = b->hi[1] - b->lo[1], dz = b->hi[2] - b->lo[2];
	if (!(dx >= 0.0f)) return 0.0f;
	return 2.0f * (dx * dy + dy * dz + dz * dx);
}

typedef struct { obox_t* tb; float* cen; float pad; } obvh_build_t;

static void obvh_build_rec(obvh_t* bvh, obvh_build_t* bd, uint32_t node_index, uint32_t first, uint32_t count) {
	obvh_node_t* node = &bvh->nodes[node_index];
	obox_t box, cbox; obox_empty(&box); obox_empty(&cbox);
	for (uint32_t i = first; i != first + count; ++i) {
		uint32_t t = bvh->order[i];
		obox_merge(&box, &bd->tb[t]);
		obox_grow(&cbox, bd->cen + 3 * t);
	}
	for (int a = 0; a != 3; ++a) { node->lo[a] = box.lo[a] - bd->pad; node->hi[a] = box.hi[a] + bd->pad; }
	node->left = (int32_t) first; node->count = (int32_t) count;
	if (count <= 4) return;
	/* binned SAH over the longest centroid axis candidates (all three axes) */
	enum { BINS = 16 };
	float best_cost = INFINITY; int best_axis = -1; int best_split = 0;
	for (int a = 0; a != 3; ++a) {
		float lo = cbox.lo[a], ext = cbox.hi[a] - cbox.lo[a];
		if (!(ext > 0.0f)) continue;
		obox_t bb[BINS]; uint32_t bc[BINS];
		for (int b = 0; b != BINS; ++b) { obox_empty(&bb[b]); bc[b] = 0; }
		float scale = (float) BINS / ext;
		for (uint32_t i = first; i != first + count; ++i) {
			uint32_t t = bvh->order[i];
			int b = (int) ((bd->cen[3 * t + a] - lo) * scale);
			if (b >= BINS) b = BINS - 1;
			if (b < 0) b = 0;
			obox_merge(&bb[b], &bd->tb[t]); ++bc[b];
		}
		float right_area[BINS]; uint32_t right_count[BINS];
		obox_t acc; obox_empty(&acc); uint32_t cnt = 0;
		for (int b = BINS - 1; b > 0; --b) { obox_merge(&acc, &bb[b]); cnt += bc[b]; right_area[b] = obox_area(&acc); right_count[b] = cnt; }
		obox_empty(&acc); cnt = 0;
		for (int b = 0; b + 1 < BINS; ++b) {
			obox_merge(&acc, &bb[b]); cnt += bc[b];
			if (cnt == 0 || right_count[b + 1] == 0) continue;
			float cost = obox_area(&acc) * (float) cnt + right_area[b + 1] * (float) right_count[b + 1];
			if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b + 1; }
		}
	}
	uint32_t mid;
	if (best_axis < 0) {
		if (count <= 16) return; /* all centroids coincide: keep a fat leaf */
		mid = first + count / 2;
	}
	else {
		float lo = cbox.lo[best_axis], ext = cbox.hi[best_axis] - cbox.lo[best_axis];
		float scale = (float) BINS / ext;
		uint32_t i = first, j = first + count;
		while (i < j) {
			uint32_t t = bvh->order[i];
			int b = (int) ((bd->cen[3 * t + best_axis] - lo) * scale);
			if (b >= BINS) b = BINS - 1;
			if (b < 0) b = 0;
			if (b < best_split) ++i;
			else { --j; bvh->order[i] = bvh->order[j]; bvh->order[j] = t; }
		}
		mid = i;
		if (mid == first || mid == first + count) mid = first + count / 2;
	}
	uint32_t child = bvh->node_count;
	bvh->node_count += 2;
	node->left = (int32_t) child; node->count = 0;
	obvh_build_rec(bvh, bd, child, first, mid - first);
	obvh_build_rec(bvh, bd, child + 1, mid, first + count - mid);
}

static inline int obvh_build(obvh_t* bvh, const float* tris, uint32_t tri_count) {
	memset(bvh, 0, sizeof(*bvh));
	bvh->tris = tris; bvh->tri_count = tri_count;
	bvh->order = (uint32_t*) malloc(sizeof(uint32_t) * (tri_count ? tri_count : 1));
	bvh->nodes = (obvh_node_t*) malloc(sizeof(obvh_node_t) * (2 * (size_t) tri_count + 2));
	obvh_build_t bd;
	bd.tb = (obox_t*) malloc(sizeof(obox_t) * (tri_count ? tri_count : 1));
	bd.cen = (float*) malloc(sizeof(float) * 3 * (tri_count ? tri_count : 1));
	obox_t scene; obox_empty(&scene);
	for (uint32_t t = 0; t != tri_count; ++t) {
		obox_empty(&bd.tb[t]);
		for (int k = 0; k != 3; ++k) obox_grow(&bd.tb[t], tris + 9 * t + 3 * k);
		for (int a = 0; a != 3; ++a) bd.cen[3 * t + a] = 0.5f * (bd.tb[t].lo[a] + bd.tb[t].hi[a]);
		obox_merge(&scene, &bd.tb[t]);
		bvh->order[t] = t;
	}
	float ext = 0.0f;
	for (int a = 0; a != 3; ++a) {
		float m = fabsf(scene.lo[a]) > fabsf(scene.hi[a]) ? fabsf(scene.lo[a]) : fabsf(scene.hi[a]);
		if (m > ext) ext = m;
	}
	bd.pad = ext * (1.0f / 65536.0f);
	bvh->node_count = 1;
	if (tri_count) obvh_build_rec(bvh, &bd, 0, 0, tri_count);
	else { bvh->nodes[0].count = 0; bvh->nodes[0].left = -1; for (int a = 0; a != 3; ++a) { bvh->nodes[0].lo[a] = INFINITY; bvh->nodes[0].hi[a] = -INFINITY; } }
	free(bd.tb); free(bd.cen);
	return 0;
}

static inline void obvh_destroy(obvh_t* bvh) { free(bvh->order); free(bvh->nodes); memset(bvh, 0, sizeof(*bvh)); }

/* Any-hit query (terminate on first hit), shading_pass.frag.glsl:128-135 */
static inline int obvh_occluded(const obvh_t* bvh, v3 o, v3 d, float tmin, float tmax) {
	if (!bvh->tri_count) return 0;
	if (!(tmax > tmin)) return 0; /* Vulkan leaves tmax<tmin / NaN undefined; the oracle defines "miss"  */
	v3 inv_d = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	int32_t stack[128]; int sp = 0;
	stack[sp++] = 0;
	float t, u, v;
	while (sp) {
		const obvh_node_t* n = &bvh->nodes[stack[--sp]];
		if (!obvh_ray_box(n, o, inv_d, tmin, tmax)) continue;
		if (n->count) {
			for (int32_t i = 0; i != n->count; ++i)
				if (oracle_ray_triangle(bvh->tris + 9 * (size_t) bvh->order[n->left + i], o, d, tmin, tmax, &t, &u, &v)) return 1;
		}
		else if (sp + 2 <= 128) { stack[sp++] = n->left; stack[sp++] = n->left + 1; }
	}
	return 0;
}

/* Closest-hit query for the visibility-buffer stand-in. Returns triangle index or -1.
   Ties in t resolve to the lowest triangle index so the result is order independent. */
static inline int32_t obvh_closest(const obvh_t* bvh, v3 o, v3 d, float tmin, float tmax, float* out_t) {
	if (!bvh->tri_count) return -1;
	v3 inv_d = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	int32_t stack[128]; int sp = 0;
	stack[sp++] = 0;
	int32_t best = -1; float best_t = tmax;
	float t, u, v;
	while (sp) {
		const obvh_node_t* n = &bvh->nodes[stack[--sp]];
		/* closed upper bound so that equal-t candidates are still visited for the tie rule */
		if (!obvh_ray_box(n, o, inv_d, tmin, best_t)) continue;
		if (n->count) {
			for (int32_t i = 0; i != n->count; ++i) {
				int32_t ti = (int32_t) bvh->order[n->left + i];
				if (oracle_ray_triangle(bvh->tris + 9 * (size_t) ti, o, d, tmin, INFINITY, &t, &u, &v)) {
					if (t < best_t || (t == best_t && best >= 0 && ti < best)) { best_t = t; best = ti; }
				}
			}
		}
		else if (sp + 2 <= 128) { stack[sp++] = n->left; stack[sp++] = n->left + 1; }
	}
	if (out_t) *out_t = best_t;
	return best;
}

#endif
