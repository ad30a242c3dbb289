"""The linear BVH built on the GPU (vkr_lbvh_gpu.cu) against its sequential reference (vkr_lbvh.cpp): node pairs, triangle
slots and original indices must be equal array for array -- every step of the builder is defined so that the parallel and the
sequential form give the same bytes. The reference itself is tested on the CPU (tests/test_host_logic.py: structure, Morton order,
traversal == brute force). Then end to end: a scene loaded with VKR_BVH_BUILDER=lbvh_gpu shades to the same frame, bit for bit.
(Written after this round's GPU budget was spent: this file sorts last among the GPU tests.)"""
import ctypes as C
import time

import numpy as np
import pytest

from tests import harness as H
from tests.test_host_logic import _probe_bvh, BUILDERS
from vulkan_renderer_b200 import api

pytestmark = pytest.mark.gpu


def _probe_device(lib, device, tris):
	P = C.POINTER
	nodes = P(C.c_float)(); tri = P(C.c_float)(); ids = P(C.c_uint32)(); nc = C.c_uint64(); md = C.c_uint32()
	tris = np.ascontiguousarray(tris, dtype=np.float32)
	assert lib.vkr_bvh_build_probe_device(C.byref(device), tris.ctypes.data, len(tris), C.byref(nodes), C.byref(nc), C.byref(tri), C.byref(ids), C.byref(md)) == 0
	n = len(tris)
	out = (np.ctypeslib.as_array(nodes, (nc.value, 16)).copy(), np.ctypeslib.as_array(tri, (n, 12)).copy(), np.ctypeslib.as_array(ids, (n,)).copy(), md.value)
	lib.vkr_bvh_free_probe(nodes, tri, ids)
	return out


def _device():
	lib = api.load_library(); dev = api.Device()
	assert lib.vkr_create_device(C.byref(dev), 0, None) == 0
	return lib, dev


def _scene_triangles(name, **overrides):
	info = H.dataset(name, **overrides); vks = H.read_vks(info["vks"])
	return H.oracle.dequantize_for_bvh(vks["positions"], vks["factor"], vks["summand"])


def _random_soup(n, seed):
	rng = np.random.default_rng(seed)
	centres = rng.uniform(-50.0, 50.0, (n, 1, 3)) * np.array([1.0, 1.0, 0.1])
	tris = (centres + rng.normal(scale=0.3, size=(n, 3, 3))).astype(np.float32).reshape(n, 9)
	tris[: n // 50] = tris[0]          # coincident triangles: equal Morton codes, the index breaks the ties
	return tris


@pytest.mark.parametrize("source", ["cornell", "mini_city", "roughness_planes", "soup_5", "soup_4097", "soup_300000"])
def test_gpu_builder_equals_the_host_reference_array_for_array(source):
	lib, dev = _device()
	try:
		tris = _random_soup(int(source.split("_")[1]), 7) if source.startswith("soup_") else _scene_triangles(source)
		t0 = time.time(); nodes, slots, ids, depth = _probe_device(lib, dev, tris); t_gpu = time.time() - t0
		t0 = time.time(); ref_nodes, ref_slots, ref_ids, ref_depth = _probe_bvh(lib, tris, BUILDERS["lbvh"]); t_host = time.time() - t0
	finally:
		lib.vkr_destroy_device(C.byref(dev))
	print("%s: %d triangles, %d node pairs, depth %d; GPU %.1f ms incl. transfers, host reference %.1f ms" % (source, len(tris), len(nodes), depth, 1e3 * t_gpu, 1e3 * t_host))
	assert np.array_equal(ids, ref_ids), "Morton order differs"
	assert np.array_equal(slots.view(np.uint32), ref_slots.view(np.uint32))
	assert nodes.shape == ref_nodes.shape and depth == ref_depth
	assert np.array_equal(nodes.view(np.uint32), ref_nodes.view(np.uint32))


def test_gpu_builder_declines_what_the_host_handles():
	lib, dev = _device()
	try:
		P = C.POINTER
		nodes = P(C.c_float)(); tri = P(C.c_float)(); ids = P(C.c_uint32)(); nc = C.c_uint64(); md = C.c_uint32()
		tiny = _random_soup(4, 1)
		assert lib.vkr_bvh_build_probe_device(C.byref(dev), tiny.ctypes.data, 4, C.byref(nodes), C.byref(nc), C.byref(tri), C.byref(ids), C.byref(md)) == 1
		assert not nodes and nc.value == 0
	finally:
		lib.vkr_destroy_device(C.byref(dev))


def test_frames_do_not_depend_on_the_builder(monkeypatch):
	"""Shadow rays are an OR over all triangles: the same frame with the SAH tree, the host linear BVH and the GPU-built linear BVH."""
	from vulkan_renderer_b200 import api as A
	width, height = 128, 72
	info = H.dataset("mini_city")
	frames = {}
	for builder in ("sah", "lbvh", "lbvh_gpu"):
		monkeypatch.setenv("VKR_BVH_BUILDER", builder)
		frame = H.open_frame(info)
		try:
			frame.configure(sample_count=4, strategy=A.STRATEGY_DIFFUSE_SPECULAR_MIS, heuristic=A.MIS_OPTIMAL_CLAMPED, trace_shadow_rays=1)
			vis, gb = frame.gbuffer_host(width, height)
			frames[builder] = (frame.shade_host(width, height, gb), int(frame.scene.shadow_node_count), int(frame.scene.shadow_max_depth))
		finally:
			frame.close()
	assert frames["lbvh"][1:] == frames["lbvh_gpu"][1:] and frames["sah"][1] != frames["lbvh"][1]
	assert np.array_equal(frames["sah"][0].view(np.uint32), frames["lbvh"][0].view(np.uint32))
	assert np.array_equal(frames["sah"][0].view(np.uint32), frames["lbvh_gpu"][0].view(np.uint32))
