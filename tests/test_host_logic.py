"""CPU tests of the host side of libvkr_b200.so: the C-ABI exports every symbol of include/vkr_b200.h, the loaders
read the reference's file formats (checked against an independent numpy reading), the constant block has the
reference's layout, the BVH builder produces a valid tree, and error paths behave like the reference's (int codes,
zeroed structs). No compute entry point is called here (there is no GPU and no CPU fallback)."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

from tests import harness as H
from vulkan_renderer_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_library):
	header = open(os.path.join(ROOT, "include", "vkr_b200.h")).read()
	declared = set(re.findall(r"\b(vkr_[a-z0-9_]+)\s*\(", header))
	declared -= {n for n in declared if n.endswith("_t")}
	assert declared, "no declarations parsed"
	missing = [n for n in sorted(declared) if not hasattr(built_library, n)]
	assert not missing, missing
	assert set(api.EXPORTED_SYMBOLS) == declared
	assert built_library.vkr_abi_version() == 1


def test_struct_sizes_match_the_reference_layouts():
	# src/camera.h:24-44 (48 bytes), src/polygonal_light.h:100-137 (192 bytes, 88-byte quicksave prefix)
	assert C.sizeof(api.Camera) == 48
	assert C.sizeof(api.PolygonalLight) == 192
	assert api.PolygonalLight.vertex_count.offset == 80 and api.PolygonalLight.rotation.offset == 96
	assert api.PolygonalLight.texture_file_path.offset == 160
	assert C.sizeof(api.LtcConstants) == 32


def _host_objects(info):
	lib = api.load_library()
	scene = api.Scene(); ltc = api.LtcTable(); noise = api.NoiseTable(); spec = api.SceneSpecification()
	assert lib.vkr_load_scene(C.byref(scene), None, info["vks"].encode(), info["textures"].encode(), 1) == 0
	assert lib.vkr_load_ltc_table(C.byref(ltc), None, info["ltc"].encode(), 51) == 0
	assert lib.vkr_load_noise_table(C.byref(noise), None, 256, 256, 64, api.NOISE_WHITE) == 0
	assert lib.vkr_quick_load(C.byref(spec), info["save"].encode()) == 0
	return lib, scene, ltc, noise, spec


def test_loaders_agree_with_independent_numpy_reading():
	info = H.dataset("mini_city")
	lib, scene, ltc, noise, spec = _host_objects(info)
	vks = H.read_vks(info["vks"])
	assert scene.triangle_count == vks["triangle_count"] == info["triangle_count"]
	assert np.allclose(list(scene.dequantization_factor), vks["factor"], rtol=0, atol=0)
	assert np.allclose(list(scene.dequantization_summand), vks["summand"], rtol=0, atol=0)
	names = [scene.material_names[i].decode() for i in range(scene.material_count)]
	assert names == vks["names"]
	mp = np.ctypeslib.as_array(scene.material_params, (scene.material_count, 8))
	assert np.array_equal(mp, info["material_params"])     # RGBA16F *.vkt round trip
	t0, t1 = H.quantize_ltc(info["ltc"])
	assert np.array_equal(np.ctypeslib.as_array(ltc.h_table0, t0.shape), t0)   # src/ltc_table.c:82-116
	assert np.array_equal(np.ctypeslib.as_array(ltc.h_table1, t1.shape), t1)
	assert np.array_equal(np.ctypeslib.as_array(noise.h_noise, (64, 256, 256, 4)), H.white_noise_table())   # src/noise_table.c:73-75
	c = ltc.constants   # src/ltc_table.c:184-191
	assert c.fresnel_index_factor == 50.0 and c.roughness_summand == np.float32(0.5 / 64) and c.roughness_factor == np.float32(63 / 64)
	assert spec.polygonal_light_count == len(info["lights"])
	for i, L in enumerate(info["lights"]):
		light = spec.polygonal_lights[i]
		assert light.vertex_count == 4 and np.allclose(list(light.translation), L["translation"])
	assert abs(spec.camera.vertical_fov - 0.33 * np.pi) < 1e-6
	lib.vkr_destroy_scene_specification(C.byref(spec)); lib.vkr_destroy_noise_table(C.byref(noise), None); lib.vkr_destroy_ltc_table(C.byref(ltc), None); lib.vkr_destroy_scene(C.byref(scene), None)
	assert scene.triangle_count == 0 and not scene.material_params    # destroy zeroes the struct (SURVEY 8b conventions)


def test_quicksave_round_trip(tmp_path):
	info = H.dataset("mini_city")
	lib, scene, ltc, noise, spec = _host_objects(info)
	path = str(tmp_path / "copy.save").encode()
	assert lib.vkr_quick_save(C.byref(spec), path) == 0
	assert open(path, "rb").read() == open(info["save"], "rb").read()   # same bytes as the synthetic writer (layout src/main.c:49-130)
	assert lib.vkr_quick_load(C.byref(spec), b"/nonexistent/file.save") != 0
	assert spec.polygonal_light_count == len(info["lights"])            # a failed load leaves the old specification intact


def test_constant_block_layout_and_light_maths():
	info = H.dataset("mini_city")
	lib, scene, ltc, noise, spec = _host_objects(info)
	st = api.RenderSettings(); lib.vkr_specify_default_render_settings(C.byref(st))
	assert (st.exposure_factor, st.sample_count, st.sampling_strategies, st.mis_heuristic, st.mis_visibility_estimate) == (8.0, 1, 3, 3, 0.5)   # src/main.c:232-249
	st.animate_noise = 0
	size = lib.vkr_get_constants_size(C.byref(spec))
	assert size == 256 + 320 * spec.polygonal_light_count                 # quads: 320 bytes per light (src/main.c:334)
	buf = (C.c_uint8 * size)()
	assert lib.vkr_write_constants(buf, C.byref(spec), C.byref(st), C.byref(scene), C.byref(ltc), C.byref(noise), 320, 200) == size
	cb = bytes(buf)
	f = lambda off, n=1: np.frombuffer(cb[off:off + 4 * n], dtype=np.float32)
	u = lambda off, n=1: np.frombuffer(cb[off:off + 4 * n], dtype=np.uint32)
	assert tuple(u(160, 2)) == (320, 200) and f(176)[0] == 8.0 and f(180)[0] == 1.0
	assert tuple(u(184, 3)) == (255, 255, 63) and tuple(u(208, 4)) == (0, 0x123456, 0x2468AC, 0x369D02)   # src/noise_table.c:161-168
	assert np.allclose(f(144, 3), info["camera"]["position"])
	# pixel_to_ray * (w/2 - 0.5, h/2 - 0.5, 1) points along the view direction (camera looks at the dataset's target)
	p2r = f(96, 12).reshape(3, 4)[:, :3].astype(np.float64)
	d = p2r @ np.array([159.5, 99.5, 1.0]); d /= np.linalg.norm(d)
	w2p = f(32, 16).reshape(4, 4).astype(np.float64)
	target = np.array(info["camera"]["position"]) + d * 3.0
	clip = w2p @ np.append(target, 1.0)
	assert abs(clip[0] / clip[3]) < 1e-3 and abs(clip[1] / clip[3]) < 1e-3
	# light block: plane through the vertices, radiance = flux / (pi * area) (src/polygonal_light.c:46-104)
	for i in range(spec.polygonal_light_count):
		p = 256 + 320 * i
		plane = f(p + 64, 4).astype(np.float64); radiance = f(p + 48, 3); area = f(p + 144)[0]
		verts = f(p + 160 + 64, 16).reshape(4, 4)[:, :3].astype(np.float64)
		assert np.allclose(verts @ plane[:3] + plane[3], 0.0, atol=1e-4)
		assert abs(np.linalg.norm(plane[:3]) - 1.0) < 1e-5
		sx, sy = info["lights"][i]["scaling"]
		assert abs(area - sx * sy) < 1e-4 * sx * sy
		assert np.allclose(radiance, np.array(info["lights"][i]["flux"]) / (np.pi * area), rtol=1e-5)
		assert u(p + 80)[0] == 4 and u(p + 84)[0] == 0


BUILDERS = {"sah": 0, "lbvh": 1}   # vkr_bvh.cpp (default), vkr_lbvh.cpp (host reference of the GPU builder)


def _probe_bvh(lib, tris, builder=0):
	P = C.POINTER
	nodes = P(C.c_float)(); tri = P(C.c_float)(); ids = P(C.c_uint32)(); nc = C.c_uint64(); md = C.c_uint32()
	tris = np.ascontiguousarray(tris, dtype=np.float32)
	assert lib.vkr_bvh_build_probe_with(builder, tris.ctypes.data, len(tris), C.byref(nodes), C.byref(nc), C.byref(tri), C.byref(ids), C.byref(md)) == 0
	n = len(tris)
	out = (np.ctypeslib.as_array(nodes, (nc.value, 16)).copy(), np.ctypeslib.as_array(tri, (max(n, 1), 12)).copy()[:n], np.ctypeslib.as_array(ids, (max(n, 1),)).copy()[:n], md.value)
	lib.vkr_bvh_free_probe(nodes, tri, ids)
	return out


@pytest.mark.parametrize("builder", sorted(BUILDERS))
@pytest.mark.parametrize("name", ["cornell", "mini_city"])
def test_bvh_builder_structure(name, builder):
	info = H.dataset(name)
	lib = api.load_library()
	vks = H.read_vks(info["vks"])
	tris = H.oracle.dequantize_for_bvh(vks["positions"], vks["factor"], vks["summand"])
	nodes, slots, ids, depth = _probe_bvh(lib, tris, BUILDERS[builder])
	n = len(tris)
	assert sorted(ids.tolist()) == list(range(n))                                   # every triangle exactly once
	T = tris.reshape(-1, 3, 3)
	assert np.array_equal(slots[:, 0:3], T[ids][:, 0]) and np.array_equal(slots[:, 3:6], T[ids][:, 1] - T[ids][:, 0]) and np.array_equal(slots[:, 6:9], T[ids][:, 2] - T[ids][:, 0])
	refs = nodes[:, 12:14].copy().view(np.int32)
	seen_nodes = np.zeros(len(nodes), dtype=int); seen_slots = np.zeros(n, dtype=int)
	def walk(k, lo, hi, level):
		assert level <= depth
		for c in range(2):
			ctr = nodes[k, 6 * c:6 * c + 3].astype(np.float64); half = nodes[k, 6 * c + 3:6 * c + 6].astype(np.float64)   # centre, half extent
			clo = ctr - half; chi = ctr + half
			ref = int(refs[k, c])
			if ref >= 0:
				seen_nodes[ref] += 1
				assert np.all(clo >= lo - 1e-4) and np.all(chi <= hi + 1e-4)       # child boxes nest (up to the rounding of c, h)
				walk(ref, clo, chi, level + 1)
			else:
				first, count = (ref & 0x7FFFFFFF) >> 4, ref & 15
				assert count <= 4
				for s in range(first, first + count):
					seen_slots[s] += 1
					v = T[ids[s]]
					assert np.all(v >= clo) and np.all(v <= chi)                    # padded leaf boxes contain their triangles
	seen_nodes[0] = 1
	walk(0, np.full(3, -np.inf), np.full(3, np.inf), 1)
	assert np.all(seen_nodes == 1) and np.all(seen_slots == 1)
	assert depth < 62


@pytest.mark.parametrize("builder", sorted(BUILDERS))
@pytest.mark.parametrize("name", ["cornell", "mini_city", "roughness_planes"])
def test_bvh_traversal_equals_brute_force_for_every_builder(name, builder):
	"""The device's any-hit traversal (vkr_trace.cuh, compiled for the CPU by tests/device_on_host.cpp) over the BVH of each builder against a
	loop over all triangles with the same triangle predicate: the shadow result must not depend on the builder (shadow predicate)."""
	from tests.test_device_on_host import _lib
	dev = _lib()
	info = H.dataset(name); lib = api.load_library()
	vks = H.read_vks(info["vks"])
	tris = H.oracle.dequantize_for_bvh(vks["positions"], vks["factor"], vks["summand"])
	nodes, slots, ids, depth = _probe_bvh(lib, tris, BUILDERS[builder])
	assert depth + 2 <= 64
	rng = np.random.default_rng(11)
	T = tris.reshape(-1, 3, 3); lo = T.reshape(-1, 3).min(0); hi = T.reshape(-1, 3).max(0)
	n_rays = 1500
	origins = rng.uniform(lo, hi, (n_rays, 3)); targets = T[rng.integers(0, len(T), n_rays)].mean(1) + rng.normal(scale=0.05, size=(n_rays, 3))
	d = targets - origins; length = np.linalg.norm(d, axis=1, keepdims=True); d /= length
	rays = np.concatenate([origins, d, np.full((n_rays, 1), 1e-3), length * rng.uniform(0.3, 1.5, (n_rays, 1))], axis=1).astype(np.float32)
	rays[:20, 3:6] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 20)]      # axis-parallel directions: the slab test sees infinities
	rays[20:30, 7] = 0.0                                                     # empty intervals are misses by definition
	out_bvh = np.zeros(n_rays, dtype=np.uint8); out_brute = np.zeros(n_rays, dtype=np.uint8)
	nodes = np.ascontiguousarray(nodes, dtype=np.float32); slots = np.ascontiguousarray(slots, dtype=np.float32)
	dev.vkr_device_on_host_trace_any(nodes.ctypes.data_as(C.c_void_p), slots.ctypes.data_as(C.c_void_p), C.c_uint32(len(slots)), C.c_uint32(n_rays), rays.ctypes.data_as(C.c_void_p),
		out_bvh.ctypes.data_as(C.c_void_p), out_brute.ctypes.data_as(C.c_void_p))
	assert np.array_equal(out_bvh, out_brute)
	assert 0.05 < out_brute.mean() < 0.98 and not out_brute[20:30].any()


def _probe_bvh4(lib, tris):
	P = C.POINTER
	nodes4 = P(C.c_float)(); tri = P(C.c_float)(); ids = P(C.c_uint32)(); nc = C.c_uint64(); md = C.c_uint32(); nc2 = C.c_uint64(); md2 = C.c_uint32()
	tris = np.ascontiguousarray(tris, dtype=np.float32)
	assert lib.vkr_bvh4_build_probe(tris.ctypes.data, len(tris), C.byref(nodes4), C.byref(nc), C.byref(tri), C.byref(ids), C.byref(md), C.byref(nc2), C.byref(md2)) == 0
	n = len(tris)
	out = (np.ctypeslib.as_array(nodes4, (nc.value, 32)).copy(), np.ctypeslib.as_array(tri, (max(n, 1), 12)).copy()[:n], md.value, nc2.value, md2.value)
	lib.vkr_bvh_free_probe(nodes4, tri, ids)
	return out


@pytest.mark.parametrize("name", ["cornell", "mini_city", "roughness_planes"])
def test_four_wide_collapse_keeps_the_tree_and_the_answers(name):
	"""Groundwork for a 4-wide trace loop: the collapse of the SAH tree into 128-byte nodes references every leaf of the binary tree exactly once,
	roughly halves the depth, and the per-thread 4-wide traversal (vkr_trace.cuh: occluded4, compiled for the CPU) answers like the binary one."""
	from tests.test_device_on_host import _lib
	dev = _lib(); lib = api.load_library()
	info = H.dataset(name); vks = H.read_vks(info["vks"])
	tris = H.oracle.dequantize_for_bvh(vks["positions"], vks["factor"], vks["summand"])
	nodes2, slots, ids, depth2 = _probe_bvh(lib, tris, BUILDERS["sah"])
	nodes4, slots4, depth4, count2, d2 = _probe_bvh4(lib, tris)
	assert np.array_equal(slots, slots4) and count2 == len(nodes2) and d2 == depth2
	refs2 = nodes2[:, 12:14].copy().view(np.int32).reshape(-1); refs4 = nodes4[:, 24:28].copy().view(np.int32).reshape(-1)
	leaves2 = sorted(int(r) for r in refs2 if r < 0 and (r & 15)); leaves4 = sorted(int(r) for r in refs4 if r < 0 and (r & 15))
	assert leaves2 == leaves4                                                   # the same leaves, each once
	inner4 = sorted(int(r) for r in refs4 if r >= 0)
	assert inner4 == list(range(1, len(nodes4))) and len(nodes4) < 0.75 * len(nodes2) + 2 and depth4 <= (depth2 + 1) // 2 + 2
	empty = refs4[(refs4 < 0) & ((refs4 & 15) == 0)]
	assert (nodes4.reshape(-1, 32)[:, 3:24:6][refs4.reshape(-1, 4) == -2147483648] == -1.0).all() or len(empty) == 0   # unused children cannot be hit
	rng = np.random.default_rng(5)
	T = tris.reshape(-1, 3, 3); lo = T.reshape(-1, 3).min(0); hi = T.reshape(-1, 3).max(0)
	n_rays = 3000
	origins = rng.uniform(lo, hi, (n_rays, 3)); targets = T[rng.integers(0, len(T), n_rays)].mean(1) + rng.normal(scale=0.05, size=(n_rays, 3))
	d = targets - origins; length = np.linalg.norm(d, axis=1, keepdims=True); d /= length
	rays = np.concatenate([origins, d, np.full((n_rays, 1), 1e-3), length * rng.uniform(0.3, 1.5, (n_rays, 1))], axis=1).astype(np.float32)
	rays[:20, 3:6] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 20)]
	out2 = np.zeros(n_rays, dtype=np.uint8); out4 = np.zeros(n_rays, dtype=np.uint8); s2 = C.c_uint64(); s4 = C.c_uint64()
	n2 = np.ascontiguousarray(nodes2, dtype=np.float32); n4 = np.ascontiguousarray(nodes4, dtype=np.float32); sl = np.ascontiguousarray(slots, dtype=np.float32)
	dev.vkr_device_on_host_trace_any_wide(n2.ctypes.data_as(C.c_void_p), n4.ctypes.data_as(C.c_void_p), sl.ctypes.data_as(C.c_void_p), C.c_uint32(n_rays), rays.ctypes.data_as(C.c_void_p),
		out2.ctypes.data_as(C.c_void_p), out4.ctypes.data_as(C.c_void_p), C.byref(s2), C.byref(s4))
	assert np.array_equal(out2, out4) and 0.02 < out2.mean() < 0.99


@pytest.mark.parametrize("name", ["cornell", "mini_city", "mini_room", "roughness_planes"])
def test_quantised_node_pairs_answer_like_the_float_pairs(name):
	"""The trace warps walk 32-byte node pairs whose child boxes are 16-bit grid coordinates (vkr_trace.cuh). Quantisation must be conservative: every box
	contains its float box (checked by decoding), and rays -- axis-parallel ones, rays along box faces, rays from surface points -- get the answers of the
	float pairs, bit for bit, because hit / miss is the OR over the triangles the predicate accepts."""
	from tests.test_device_on_host import _lib
	dev = _lib(); lib = api.load_library()
	info = H.dataset(name); vks = H.read_vks(info["vks"])
	tris = H.oracle.dequantize_for_bvh(vks["positions"], vks["factor"], vks["summand"])
	nodes, slots, ids, depth = _probe_bvh(lib, tris, BUILDERS["sah"])
	n2 = np.ascontiguousarray(nodes, dtype=np.float32); sl = np.ascontiguousarray(slots, dtype=np.float32)
	# the grid as shadow_grid_from_root() makes it (vkr_bvh.cpp), restated: the root's boxes with two cells to spare, 65531 cells across
	c = n2[0, [0, 1, 2, 6, 7, 8]].reshape(2, 3); hx = n2[0, [3, 4, 5, 9, 10, 11]].reshape(2, 3)
	lo = (c - hx).min(0); hi = (c + hx).max(0); extent = hi - lo
	extent = np.where(extent > 1e-6 * extent.max(), extent, max(1e-6 * extent.max(), 1e-30)).astype(np.float32)
	scale = (np.float32(65531.0) / extent).astype(np.float32); gmin = (lo - np.float32(2.0) / scale).astype(np.float32)
	grid = np.concatenate([gmin, scale]).astype(np.float32)
	rng = np.random.default_rng(21)
	T = tris.reshape(-1, 3, 3); blo = T.reshape(-1, 3).min(0); bhi = T.reshape(-1, 3).max(0)
	n_rays = 6000
	origins = rng.uniform(blo, bhi, (n_rays, 3)); targets = T[rng.integers(0, len(T), n_rays)].mean(1) + rng.normal(scale=0.02, size=(n_rays, 3))
	origins[:2000] = T[rng.integers(0, len(T), 2000)].mean(1)                       # rays that start on a surface, like shadow rays do
	d = targets - origins; length = np.linalg.norm(d, axis=1, keepdims=True); length[length == 0] = 1.0; d /= length
	rays = np.concatenate([origins, d, np.full((n_rays, 1), 1e-3), length * rng.uniform(0.3, 1.5, (n_rays, 1))], axis=1).astype(np.float32)
	rays[:60, 3:6] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 60)] * rng.choice([-1.0, 1.0], (60, 1)).astype(np.float32)   # axis-parallel: infinite slab distances
	out_f = np.zeros(n_rays, dtype=np.uint8); out_q = np.zeros(n_rays, dtype=np.uint8); pairs8 = np.zeros((len(n2), 8), dtype=np.uint32); visits = (C.c_uint64 * 2)()
	dev.vkr_device_on_host_trace_quantised(n2.ctypes.data_as(C.c_void_p), C.c_uint64(len(n2)), sl.ctypes.data_as(C.c_void_p), grid.ctypes.data_as(C.c_void_p), C.c_uint32(n_rays), rays.ctypes.data_as(C.c_void_p),
		out_f.ctypes.data_as(C.c_void_p), out_q.ctypes.data_as(C.c_void_p), pairs8.ctypes.data_as(C.c_void_p), visits)
	assert np.array_equal(out_f, out_q) and 0.01 < out_f.mean() < 0.995
	# every quantised box contains its float box, with at least one and at most three cells to spare
	for k in range(2):
		cc = n2[:, 6 * k:6 * k + 3].astype(np.float64); hh = n2[:, 6 * k + 3:6 * k + 6].astype(np.float64)
		real = hh[:, 0] >= 0
		qlo = (pairs8[:, 3 * k:3 * k + 3] & 0xffff).astype(np.float64); qhi = (pairs8[:, 3 * k:3 * k + 3] >> 16).astype(np.float64)
		glo = ((cc - hh) - gmin) * scale; ghi = ((cc + hh) - gmin) * scale
		assert (qlo[real] <= glo[real] - 0.99).all() and (qhi[real] >= ghi[real] + 0.99).all()
		assert (qlo[real] >= glo[real] - 2.01).all() and (qhi[real] <= ghi[real] + 2.01).all()
		assert (qlo[real] >= 0).all() and (qhi[real] <= 65535).all()
	assert np.array_equal(pairs8[:, 6:8], n2[:, 12:14].copy().view(np.uint32))


@pytest.mark.parametrize("name", ["cornell", "mini_city", "mini_room"])
def test_interleaved_node_pairs_visit_and_answer_like_the_float_pairs(name):
	"""The packed-FMA edition of the trace warps walks node pairs whose two children are stored side by side (vkr_trace.cuh): the same numbers in another order and
	the same IEEE fma per slab plane, so a ray visits exactly the pairs it visits in the float layout and gets the same answer."""
	from tests.test_device_on_host import _lib
	dev = _lib(); lib = api.load_library()
	info = H.dataset(name); vks = H.read_vks(info["vks"])
	tris = H.oracle.dequantize_for_bvh(vks["positions"], vks["factor"], vks["summand"])
	nodes, slots, ids, depth = _probe_bvh(lib, tris, BUILDERS["sah"])
	n2 = np.ascontiguousarray(nodes, dtype=np.float32); sl = np.ascontiguousarray(slots, dtype=np.float32)
	rng = np.random.default_rng(23)
	T = tris.reshape(-1, 3, 3); blo = T.reshape(-1, 3).min(0); bhi = T.reshape(-1, 3).max(0)
	n_rays = 6000
	origins = rng.uniform(blo, bhi, (n_rays, 3)); targets = T[rng.integers(0, len(T), n_rays)].mean(1) + rng.normal(scale=0.02, size=(n_rays, 3))
	origins[:2000] = T[rng.integers(0, len(T), 2000)].mean(1)
	d = targets - origins; length = np.linalg.norm(d, axis=1, keepdims=True); length[length == 0] = 1.0; d /= length
	rays = np.concatenate([origins, d, np.full((n_rays, 1), 1e-3), length * rng.uniform(0.3, 1.5, (n_rays, 1))], axis=1).astype(np.float32)
	rays[:60, 3:6] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 60)] * rng.choice([-1.0, 1.0], (60, 1)).astype(np.float32)   # axis-parallel: infinite slab distances, NaN planes
	out_f = np.zeros(n_rays, dtype=np.uint8); out_i = np.zeros(n_rays, dtype=np.uint8); pairs16 = np.zeros((len(n2), 16), dtype=np.float32); visits = (C.c_uint64 * 2)()
	dev.vkr_device_on_host_trace_interleaved(n2.ctypes.data_as(C.c_void_p), C.c_uint64(len(n2)), sl.ctypes.data_as(C.c_void_p), C.c_uint32(n_rays), rays.ctypes.data_as(C.c_void_p),
		out_f.ctypes.data_as(C.c_void_p), out_i.ctypes.data_as(C.c_void_p), pairs16.ctypes.data_as(C.c_void_p), visits)
	assert np.array_equal(out_f, out_i) and 0.01 < out_f.mean() < 0.995
	assert visits[0] == visits[1] and visits[0] > n_rays
	# the layout: centres, then half extents, children side by side; references untouched
	assert np.array_equal(pairs16[:, 0:6:2], n2[:, 0:3]) and np.array_equal(pairs16[:, 1:6:2], n2[:, 6:9])
	assert np.array_equal(pairs16[:, 6:12:2], n2[:, 3:6]) and np.array_equal(pairs16[:, 7:12:2], n2[:, 9:12])
	assert np.array_equal(pairs16[:, 12:14].copy().view(np.uint32), n2[:, 12:14].copy().view(np.uint32))


@pytest.mark.parametrize("name,light", [("mini_city", 0), ("mini_city", 2), ("mini_room", 5), ("cornell", 0)])
def test_anchored_shadow_rays_answer_like_the_plain_traversal(name, light):
	"""vkr_anchor.cuh (compiled for the CPU): shadow rays that start at the siblings of their pixel's origin path -- all of them, or those the light's cone
	touches -- give the answers of the plain traversal from the root, with fewer node visits. Origins are surface points as the G-buffer holds them, the
	rays go to points of a light of the data set and a little beyond its rim (rays outside the cone must keep all siblings)."""
	from tests.test_device_on_host import _lib
	from tests.ref_frames import host_constants
	dev = _lib(); lib = api.load_library()
	info = H.dataset(name); oi = H.OracleInputs(info)
	w, h = 96, 64
	lights = len(info["lights"])
	constants = host_constants(info, w, h, lights)
	gb = oi.gbuffer(w, h, constants, oi.visibility(w, h, constants))
	nodes, slots, ids, depth = _probe_bvh(lib, oi.shadow_tris, BUILDERS["sah"])
	valid = np.argwhere(gb[1, :, :, 3] != 0)
	rng = np.random.default_rng(11)
	pick = valid[rng.choice(len(valid), min(400, len(valid)), replace=False)]
	origins = np.ascontiguousarray(gb[0, pick[:, 0], pick[:, 1], :3], dtype=np.float32)
	stride = 256 + (160 + 16 * 4 * 2 + 16 * 2) * light
	lv = np.frombuffer(constants[stride + 160 + 64:stride + 160 + 128], dtype=np.float32).reshape(4, 4).copy()
	rays = []
	for pi in range(len(origins)):
		uv = rng.uniform(-0.15, 1.15, (6, 2))   # some targets lie beyond the rim of the light
		pts = lv[0, :3] + uv[:, :1] * (lv[1, :3] - lv[0, :3]) + uv[:, 1:] * (lv[3, :3] - lv[0, :3])
		e = pts - origins[pi]; dist = np.linalg.norm(e, axis=1)
		for i in range(len(pts)):
			rays.append((pi, e[i, 0] / dist[i], e[i, 1] / dist[i], e[i, 2] / dist[i], dist[i] * (1.0 if i else 1.3)))   # the first of each pixel overshoots the light
	rays = np.ascontiguousarray(np.array(rays, dtype=np.float32))
	out = np.zeros((len(rays), 3), dtype=np.uint8); visits = (C.c_uint64 * 3)(); stats = (C.c_uint64 * 3)()
	n2 = np.ascontiguousarray(nodes, dtype=np.float32); sl = np.ascontiguousarray(slots, dtype=np.float32)
	dev.vkr_device_on_host_trace_anchored(n2.ctypes.data_as(C.c_void_p), sl.ctypes.data_as(C.c_void_p), C.c_uint32(len(origins)), origins.ctypes.data_as(C.c_void_p), C.c_uint32(4), lv.ctypes.data_as(C.c_void_p),
		C.c_uint32(len(rays)), rays.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), visits, stats)
	assert out.max() <= 1                                                      # (2 = the two plain traversals disagree)
	assert np.array_equal(out[:, 0], out[:, 1]) and np.array_equal(out[:, 0], out[:, 2])
	assert stats[0] > 0 and stats[0] < len(rays)                              # rays outside the cone were among them, and rays inside
	assert visits[2] <= visits[1]                                             # (visiting all siblings deepest first is not always cheaper than the plain order)
	if name != "cornell":
		assert stats[1] < stats[2] and visits[2] < visits[0]                   # the cone removes siblings, and visits with them


@pytest.mark.parametrize("source", ["cornell", "mini_city", "roughness_planes", "soup_5", "soup_4097", "soup_60000"])
def test_gpu_builder_steps_reproduce_the_reference_on_the_cpu(source):
	"""csrc/vkr_lbvh.cuh -- the per-element functions the GPU builder's kernels call -- run on the CPU with every launch replaced by a loop (tests/device_on_host.cpp)
	against the sequential reference csrc/vkr_lbvh.cpp: node pairs, slots, order and depth equal array for array. What this leaves to the GPU tests are the
	launches themselves, the warp-reduced bounds and the two cub calls."""
	from tests.test_device_on_host import _lib
	dev = _lib(); lib = api.load_library()
	if source.startswith("soup_"):
		n = int(source.split("_")[1]); rng = np.random.default_rng(7)
		centres = rng.uniform(-50.0, 50.0, (n, 1, 3)) * np.array([1.0, 1.0, 0.1])
		tris = (centres + rng.normal(scale=0.3, size=(n, 3, 3))).astype(np.float32).reshape(n, 9)
		tris[: n // 50] = tris[0]
	else:
		info = H.dataset(source); vks = H.read_vks(info["vks"])
		tris = H.oracle.dequantize_for_bvh(vks["positions"], vks["factor"], vks["summand"])
	tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
	ref_nodes, ref_slots, ref_ids, ref_depth = _probe_bvh(lib, tris, BUILDERS["lbvh"])
	n = len(tris)
	nodes = np.zeros((n, 16), dtype=np.float32); slots = np.zeros((n, 12), dtype=np.float32); ids = np.zeros(n, dtype=np.uint32); count = C.c_uint64(); depth = C.c_uint32()
	P = lambda a: a.ctypes.data_as(C.c_void_p)
	assert dev.vkr_device_on_host_lbvh(P(tris), C.c_uint64(n), P(nodes), C.byref(count), P(slots), P(ids), C.byref(depth)) == 0
	assert count.value == len(ref_nodes) and depth.value == ref_depth
	assert np.array_equal(ids, ref_ids) and np.array_equal(slots.view(np.uint32), ref_slots.view(np.uint32))
	assert np.array_equal(nodes[:count.value].view(np.uint32), ref_nodes.view(np.uint32))


def test_lbvh_follows_the_morton_order_of_the_centroids():
	"""What the GPU builder has to reproduce: slots in ascending (Morton code of the centroid, original index) order, leaves of up to four slots."""
	lib = api.load_library()
	info = H.dataset("mini_city"); vks = H.read_vks(info["vks"])
	tris = H.oracle.dequantize_for_bvh(vks["positions"], vks["factor"], vks["summand"])
	nodes, slots, ids, depth = _probe_bvh(lib, tris, BUILDERS["lbvh"])
	T = tris.reshape(-1, 3, 3)
	blo = T.min(1); bhi = T.max(1); c = (np.float32(0.5) * (blo + bhi)).astype(np.float32)
	clo = c.min(0); ext = c.max(0) - clo
	inv = np.where(ext > 0, np.float32(1.0) / ext, np.float32(0.0)).astype(np.float32)
	q = np.clip((((c - clo) * inv) * np.float32(2097152.0)).astype(np.float32), 0, 2097151).astype(np.uint64)
	def expand(v):
		x = v & np.uint64(0x1fffff)
		for shift, mask in ((32, 0x1f00000000ffff), (16, 0x1f0000ff0000ff), (8, 0x100f00f00f00f00f), (4, 0x10c30c30c30c30c3), (2, 0x1249249249249249)):
			x = (x | (x << np.uint64(shift))) & np.uint64(mask)
		return x
	code = (expand(q[:, 0]) << np.uint64(2)) | (expand(q[:, 1]) << np.uint64(1)) | expand(q[:, 2])
	expected = np.lexsort((np.arange(len(code)), code))
	assert np.array_equal(ids, expected.astype(np.uint32))
	refs = nodes[:, 12:14].copy().view(np.int32).reshape(-1)
	counts = refs[refs < 0] & 15
	assert counts.max() <= 4 and counts.sum() == len(tris) and (counts >= 1).all()
	sah_nodes, _, _, sah_depth = _probe_bvh(lib, tris, BUILDERS["sah"])
	assert len(nodes) < 1.3 * len(sah_nodes) + 8 and depth <= 2 * sah_depth + 8   # a comparable tree, not a degenerate one


@pytest.mark.parametrize("builder", sorted(BUILDERS))
def test_bvh_builder_degenerate_inputs(builder):
	import functools
	lib = api.load_library()
	_probe = functools.partial(_probe_bvh, builder=BUILDERS[builder])
	nodes, slots, ids, depth = _probe(lib, np.zeros((0, 9), dtype=np.float32))              # empty scene
	assert len(nodes) == 1 and (nodes[0, 12:14].view(np.int32) & 15).tolist() == [0, 0]
	one = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], dtype=np.float32)
	nodes, slots, ids, depth = _probe(lib, one)                                             # single leaf under the root pair
	assert len(nodes) == 1 and (int(nodes[0, 12:13].view(np.int32)[0]) & 15) == 1
	same = np.tile(one, (100, 1))                                                            # 100 coincident triangles: median splits
	nodes, slots, ids, depth = _probe(lib, same)                                            # (equal Morton codes: the position breaks the ties)
	assert sorted(ids.tolist()) == list(range(100)) and depth < 62
	five = np.tile(one, (5, 1)) + np.arange(5, dtype=np.float32)[:, None]                   # the smallest tree with an inner split
	nodes, slots, ids, depth = _probe(lib, five)
	refs = nodes[:, 12:14].copy().view(np.int32).reshape(-1)
	assert sorted(ids.tolist()) == list(range(5)) and int((refs[refs < 0] & 15).sum()) == 5


def test_error_paths_return_codes_and_leave_structs_zeroed(tmp_path, capfd):
	lib = api.load_library()
	scene = api.Scene()
	assert lib.vkr_load_scene(C.byref(scene), None, b"/nonexistent.vks", b"/tmp", 1) == 1
	assert scene.triangle_count == 0
	bad = tmp_path / "bad.vks"; bad.write_bytes(struct.pack("<II", 0x123, 1) + b"\0" * 64)
	assert lib.vkr_load_scene(C.byref(scene), None, str(bad).encode(), b"/tmp", 1) == 1          # wrong marker (src/scene.c:423)
	empty = tmp_path / "empty.vks"; empty.write_bytes(struct.pack("<IIQQ6f", 0xABCABC, 1, 0, 0, 1, 1, 1, 0, 0, 0))
	assert lib.vkr_load_scene(C.byref(scene), None, str(empty).encode(), b"/tmp", 1) == 1        # zero triangles (src/scene.c:435)
	info = H.dataset("cornell")
	trunc = tmp_path / "trunc.vks"; trunc.write_bytes(open(info["vks"], "rb").read()[:-4] + struct.pack("<I", 0))
	assert lib.vkr_load_scene(C.byref(scene), None, str(trunc).encode(), info["textures"].encode(), 1) == 1   # missing EOF marker (src/scene.c:479)
	assert lib.vkr_load_scene(C.byref(scene), None, info["vks"].encode(), str(tmp_path).encode(), 1) == 1      # textures missing
	ltc = api.LtcTable()
	assert lib.vkr_load_ltc_table(C.byref(ltc), None, str(tmp_path).encode(), 51) == 1 and ltc.roughness_count == 0
	noise = api.NoiseTable()
	assert lib.vkr_load_noise_table(C.byref(noise), None, 100, 256, 64, api.NOISE_WHITE) == 1    # not a power of two
	assert lib.vkr_load_noise_table(C.byref(noise), None, 256, 256, 64, api.NOISE_BLUE) == 1     # data/noise/*.blob absent
	dev = api.Device()
	rc = lib.vkr_create_device(C.byref(dev), 0, None)
	import torch
	if not torch.cuda.is_available():
		assert rc == 1 and dev.sm_count == 0      # no CUDA device: fails loudly, no CPU fallback
	else:
		lib.vkr_destroy_device(C.byref(dev))
	assert "Failed" in capfd.readouterr().out     # printf diagnostics like the reference


def test_shading_pass_rejects_illegal_technique_and_strategy_combinations(capfd):
	"""The legality rules of the reference's settings panel (src/user_interface.cpp:90-180) are checked before any CUDA call, so they can be
	tested without a GPU: the related-work techniques sample the diffuse lobe only, GGX MIS needs a stand-alone density."""
	lib = api.load_library()
	dev = api.Device()
	def create(technique, strategy, heuristic=api.MIS_BALANCE):
		p = api.ShadingPass(); d = api.ShadingPassDesc(width=64, height=32, polygonal_light_count=1, min_polygonal_light_vertex_count=4, max_polygonal_light_vertex_count=4,
			sample_count=1, sampling_strategies=strategy, mis_heuristic=heuristic, polygon_sampling_technique=technique, stripe_count=1)
		rc = lib.vkr_create_shading_pass(C.byref(p), C.byref(dev), C.byref(d))
		assert rc == 1 and p.constants_size == 0 and not p.d_constants     # no LTC / noise tables: every call fails, the message tells why
		return capfd.readouterr().out
	for technique in range(11):
		assert "does not support sampling strategy" in create(technique, api.STRATEGY_DIFFUSE_SPECULAR_MIS)
		assert "does not support sampling strategy" in create(technique, api.STRATEGY_DIFFUSE_SPECULAR_RANDOM)
		assert "missing LTC / noise tables" in create(technique, api.STRATEGY_DIFFUSE_ONLY)
	for technique in (0, 1, 6, 7, 8, 9):
		assert "does not support sampling strategy" in create(technique, api.STRATEGY_DIFFUSE_GGX_MIS)
	for technique in (2, 3, 4, 5, 10, 11, 12):
		assert "missing LTC / noise tables" in create(technique, api.STRATEGY_DIFFUSE_GGX_MIS)
	assert "balance and power heuristics only" in create(11, api.STRATEGY_DIFFUSE_GGX_MIS, api.MIS_OPTIMAL)
	assert "unknown polygon sampling technique" in create(13, api.STRATEGY_DIFFUSE_ONLY)
	# error display: in the projected solid angle branches of the shader only; the specular variants need a diffuse + specular strategy; Arvo's error has two measures
	def create_display(technique, strategy, error_display):
		p = api.ShadingPass(); d = api.ShadingPassDesc(width=64, height=32, polygonal_light_count=1, min_polygonal_light_vertex_count=4, max_polygonal_light_vertex_count=4,
			sample_count=1, sampling_strategies=strategy, mis_heuristic=api.MIS_BALANCE, polygon_sampling_technique=technique, stripe_count=1, error_display=error_display)
		assert lib.vkr_create_shading_pass(C.byref(p), C.byref(dev), C.byref(d)) == 1
		return capfd.readouterr().out
	for technique, strategy, error_display, legal in [(11, 0, 1, True), (12, 1, 3, True), (11, 3, 2, True), (11, 3, 4, True), (12, 2, 6, True), (10, 0, 1, True), (10, 0, 2, True),
			(10, 0, 3, False), (11, 0, 4, False), (11, 1, 5, False), (4, 0, 1, False), (10, 3, 4, False), (11, 3, 7, False)]:
		out = create_display(technique, strategy, error_display)
		assert ("missing LTC / noise tables" in out) if legal else ("is not available with" in out or "does not support" in out), (technique, strategy, error_display, out)


def test_white_noise_and_synthetic_formats_are_what_the_reference_loaders_expect():
	info = H.dataset("cornell")
	raw = open(info["vks"], "rb").read()
	assert struct.unpack_from("<II", raw, 0) == (0xABCABC, 1) and struct.unpack_from("<I", raw, len(raw) - 4)[0] == 0xE0FE0F
	tex = open(os.path.join(info["textures"], "white_BaseColor.vkt"), "rb").read()
	marker, version, mips, w, h, fmt, size = struct.unpack_from("<IIIIIIQ", tex, 0)
	assert (marker, version, mips, w, h, fmt) == (0xBC1BC1, 1, 3, 4, 4, 97) and struct.unpack_from("<I", tex, len(tex) - 4)[0] == 0xE0FE0F
	fit = open(os.path.join(info["ltc"], "fit0.dat"), "rb").read()
	assert struct.unpack_from("<Q", fit, 0)[0] == 64 and len(fit) == 8 + 64 * 64 * 20


def test_loaders_survive_corrupted_files(tmp_path, capfd):
	"""Truncated, bit-flipped and padded *.vks / *.vkt / *.save / fit*.dat files: every loader either rejects the file with a diagnostic (non-zero return, object
	zeroed, like the reference: src/scene.c:68-72, src/textures.c:117-121) or loads it; none may crash or leave a half-built object behind."""
	import shutil
	lib = api.load_library()
	info = H.dataset("mini_lit")
	rng = np.random.default_rng(17)
	light_textures = [l["texture_file_path"] for l in info["lights"]]
	def mutate(src, dst):
		data = bytearray(open(src, "rb").read())
		mode = int(rng.integers(4))
		if mode == 0: data = data[:int(rng.integers(0, len(data)))]
		elif mode == 1:
			for _ in range(int(rng.integers(1, 8))): data[int(rng.integers(0, min(len(data), 96)))] = int(rng.integers(0, 256))
		elif mode == 2:
			for _ in range(int(rng.integers(1, 30))): data[int(rng.integers(0, len(data)))] = int(rng.integers(0, 256))
		else: data = data + bytes(rng.integers(0, 256, int(rng.integers(1, 64)), dtype=np.uint8))
		open(dst, "wb").write(data)
	rejected = 0
	for i in range(160):
		kind = i % 4
		if kind == 0:
			dst = str(tmp_path / "s.vks"); mutate(info["vks"], dst)
			scene = api.Scene(); rc = lib.vkr_load_scene(C.byref(scene), None, dst.encode(), info["textures"].encode(), 1)
			if rc == 0: lib.vkr_destroy_scene(C.byref(scene), None)
			assert rc == 0 or (scene.triangle_count == 0 and not scene.material_params)
		elif kind == 1:
			dst = str(tmp_path / "t.vkt"); mutate(light_textures[i % 3], dst)
			t = api.Texture(); rc = lib.vkr_load_texture(C.byref(t), dst.encode())
			if rc == 0: lib.vkr_destroy_texture(C.byref(t))
			assert rc == 0 or not t.h_texels
		elif kind == 2:
			dst = str(tmp_path / "q.save"); mutate(info["save"], dst)
			spec = api.SceneSpecification(); rc = lib.vkr_quick_load(C.byref(spec), dst.encode())
			if rc == 0:
				lt = api.LightTextures()
				if lib.vkr_create_and_assign_light_textures(C.byref(lt), None, C.byref(spec)) == 0: lib.vkr_destroy_light_textures(C.byref(lt), None)
				lib.vkr_destroy_scene_specification(C.byref(spec))
			assert rc == 0 or spec.polygonal_light_count == 0
		else:
			d = str(tmp_path / "ltc"); shutil.rmtree(d, ignore_errors=True); shutil.copytree(info["ltc"], d)
			f = os.path.join(d, "fit%d.dat" % int(rng.integers(51))); mutate(f, f)
			ltc = api.LtcTable(); rc = lib.vkr_load_ltc_table(C.byref(ltc), None, d.encode(), 51)
			if rc == 0: lib.vkr_destroy_ltc_table(C.byref(ltc), None)
			assert rc == 0 or not ltc.h_table0
		rejected += rc != 0
	assert 30 < rejected < 130
	capfd.readouterr()
