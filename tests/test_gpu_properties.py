"""Edge cases and size-independent properties of the CUDA shading path (through the C-ABI).

Small frames are checked against the oracle; at BASELINE.json's full size (config 3: 1920x1080, 8 lights, 64 spp, the
2.8 M triangle city) the oracle only checks a band of rows and the rest is covered by properties of the domain:
determinism (trace lanes pick rays up in a racy order, the sums must not depend on it), exact linearity in the radiant
flux (a power-of-two factor commutes with every rounding on the path), stripes that tile the frame.
"""
import ctypes as C

import numpy as np
import pytest

from tests import harness as H
from vulkan_renderer_b200 import api

pytestmark = pytest.mark.gpu


def _both(name, width, height, **settings):
	info = H.dataset(name)
	oi = H.OracleInputs(info)
	frame = H.open_frame(info)
	try:
		frame.configure(**settings)
		constants = frame.constants(width, height)
		vis = oi.visibility(width, height, constants)
		gb = oi.gbuffer(width, height, constants, vis)
		out_gpu = frame.shade_host(width, height, gb)
		out_cpu, _ = oi.shade(H.oracle_config(frame, width, height), constants, gb)
	finally:
		frame.close()
	return out_gpu, out_cpu


@pytest.mark.parametrize("width,height", [(1, 1), (15, 7), (17, 9), (150, 91), (33, 130)])
def test_ragged_frame_sizes(width, height):
	"""Frames that do not fill 16x8 tiles: out-of-frame lanes of a tile must neither write nor disturb their warp."""
	out_gpu, out_cpu = _both("mini_city", width, height, sample_count=2, strategy=api.STRATEGY_DIFFUSE_SPECULAR_MIS, heuristic=api.MIS_OPTIMAL_CLAMPED, trace_shadow_rays=1)
	assert np.array_equal(out_gpu.view(np.uint32), out_cpu.view(np.uint32)), H.compare_radiance(out_gpu, out_cpu)


@pytest.mark.parametrize("name", ["mini_tri", "mini_mixed", "mini_v5", "mini_v6", "mini_v7", "mini_poly"])
@pytest.mark.parametrize("strategy,heuristic", [(api.STRATEGY_DIFFUSE_ONLY, api.MIS_BALANCE), (api.STRATEGY_DIFFUSE_GGX_MIS, api.MIS_POWER),
	(api.STRATEGY_DIFFUSE_SPECULAR_SEPARATELY, api.MIS_BALANCE), (api.STRATEGY_DIFFUSE_SPECULAR_MIS, api.MIS_OPTIMAL), (api.STRATEGY_DIFFUSE_SPECULAR_RANDOM, api.MIS_BALANCE)])
def test_triangle_and_mixed_lights(name, strategy, heuristic):
	"""MAX_POLYGONAL_LIGHT_VERTEX_COUNT = 3, 5, 6, 7 and lights of different vertex counts in one frame (MIN < MAX)."""
	out_gpu, out_cpu = _both(name, 96, 64, sample_count=3, strategy=strategy, heuristic=heuristic, trace_shadow_rays=1)
	assert np.array_equal(out_gpu.view(np.uint32), out_cpu.view(np.uint32)), H.compare_radiance(out_gpu, out_cpu)
	assert float(out_cpu[..., :3].max()) > 0.0


def test_no_lights_and_background_only():
	"""Zero lights: black surfaces. A G-buffer without any surface: every warp skips shading, lights still show."""
	info = H.dataset("mini_city")
	oi = H.OracleInputs(info)
	frame = H.open_frame(info)
	try:
		frame.configure(sample_count=4, light_count=0, trace_shadow_rays=1)
		constants = frame.constants(64, 40)
		vis = oi.visibility(64, 40, constants)
		gb = oi.gbuffer(64, 40, constants, vis)
		out = frame.shade_host(64, 40, gb)
		assert np.all(out[..., :3] == 0.0) and np.all(out[..., 3] == 1.0)
		frame.configure(light_count=3)
		constants = frame.constants(64, 40)
		empty = np.zeros_like(gb)
		out = frame.shade_host(64, 40, empty)
		ref, _ = oi.shade(H.oracle_config(frame, 64, 40), constants, empty)
		assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
	finally:
		frame.close()


def test_stripes_tile_the_frame():
	"""Multi-GPU decomposition on one device: interleaved tile rows of 3 pass instances == the whole frame."""
	info = H.dataset("mini_city")
	oi = H.OracleInputs(info)
	frame = H.open_frame(info)
	try:
		frame.configure(sample_count=3, trace_shadow_rays=1)
		w, h = 100, 75   # 10 tile rows, the last one ragged
		constants = frame.constants(w, h)
		gb = oi.gbuffer(w, h, constants, oi.visibility(w, h, constants))
		whole = frame.shade_host(w, h, gb)
		assembled = np.full((h, w, 4), np.nan, dtype=np.float32)
		for k in range(3):
			frame.shade_host(w, h, gb, stripe_index=k, stripe_count=3, out=assembled)
		assert np.array_equal(assembled.view(np.uint32), whole.view(np.uint32))
	finally:
		frame.close()


def test_hdr_screenshot_halves_reassemble_to_the_half_precision_frame():
	"""Output stage (shading_pass.frag.glsl:871-887): the two LDR frames of an HDR screenshot carry the low and the high
	bytes of packHalf2x16(colour); put together they are the linear frame rounded to binary16, bit for bit."""
	info = H.dataset("mini_city")
	oi = H.OracleInputs(info)
	frame = H.open_frame(info)
	try:
		w, h = 160, 96
		frame.configure(sample_count=4, trace_shadow_rays=1, output_srgb=1, frame_bits=0)
		constants = frame.constants(w, h)
		gb = oi.gbuffer(w, h, constants, oi.visibility(w, h, constants))
		srgb = frame.shade_host(w, h, gb)
		ref_srgb, _ = oi.shade(H.oracle_config(frame, w, h), constants, gb)
		assert np.array_equal(srgb.view(np.uint32), ref_srgb.view(np.uint32))
		frame.configure(output_srgb=0)
		linear = frame.shade_host(w, h, gb)
		frame.configure(output_srgb=1, frame_bits=1); low = frame.shade_host(w, h, gb)
		frame.configure(frame_bits=2); high = frame.shade_host(w, h, gb)
		ref_high, _ = oi.shade(H.oracle_config(frame, w, h), frame.constants(w, h), gb)
		assert np.array_equal(high.view(np.uint32), ref_high.view(np.uint32))
	finally:
		frame.close()
	lo = np.rint(low[..., :3] * 255.0).astype(np.uint16); hi = np.rint(high[..., :3] * 255.0).astype(np.uint16)
	assert np.array_equal(low[..., :3], lo.astype(np.float32) * np.float32(1.0 / 255.0)) and np.array_equal(high[..., :3], hi.astype(np.float32) * np.float32(1.0 / 255.0))
	halves = (lo | (hi << 8)).astype(np.uint16)
	assert np.array_equal(halves, linear[..., :3].astype(np.float16).view(np.uint16))
	assert float(linear[..., :3].max()) > 0.0


class _FullSize:
	"""BASELINE config 3 on the device: 1920x1080, 8 quad lights, 64 spp, clamped optimal MIS, shadow rays on."""
	W, H_, SPP = 1920, 1080, 64

	def __init__(self):
		import torch
		self.torch = torch
		self.info = H.dataset("city")
		self.frame = H.open_frame(self.info)
		self.frame.configure(sample_count=self.SPP, strategy=api.STRATEGY_DIFFUSE_SPECULAR_MIS, heuristic=api.MIS_OPTIMAL_CLAMPED, trace_shadow_rays=1, show_lights=1)
		self.constants = self.frame.constants(self.W, self.H_)
		self.vis, self.gb = self.frame.gbuffer_host(self.W, self.H_)

	def shade(self, **kw):
		return self.frame.shade_host(self.W, self.H_, self.gb, **kw)


@pytest.fixture(scope="module")
def full():
	f = _FullSize()
	yield f
	f.frame.close()


def test_full_size_is_deterministic_and_stripes_tile_it(full):
	a = full.shade()
	b = full.shade()
	assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "radiance depends on the order in which trace lanes finish"
	assert (full.gb[1, :, :, 3] != 0).mean() > 0.5 and float(a[..., :3].mean()) > 0.01
	assembled = np.full_like(a, np.nan)
	for k in range(4):
		full.shade(stripe_index=k, stripe_count=4, out=assembled)
	assert np.array_equal(assembled.view(np.uint32), a.view(np.uint32))
	full.reference_frame = a


def test_full_size_band_against_the_oracle(full):
	"""The oracle shades rows 536..551 of the full-size frame (all 8 lights, 64 spp, the 2.8 M triangle BVH)."""
	a = getattr(full, "reference_frame", None)
	if a is None:
		a = full.shade()
	oi = H.OracleInputs(full.info)
	cfg = H.oracle_config(full.frame, full.W, full.H_)
	ref, rays = oi.shade(cfg, full.constants, full.gb, row_begin=536, row_end=552)
	assert rays > 1000000
	band_gpu, band_cpu = a[536:552], ref[536:552]
	assert np.array_equal(band_gpu.view(np.uint32), band_cpu.view(np.uint32)), H.compare_radiance(band_gpu, band_cpu)


def test_full_size_radiance_is_linear_in_the_flux(full):
	"""Doubling every light's radiant flux doubles every pixel exactly (NaN-marked pixels excepted)."""
	a = getattr(full, "reference_frame", None)
	if a is None:
		a = full.shade()
	lights = full.frame.spec.polygonal_lights
	for i in range(full.frame.light_count):
		for c in range(3):
			lights[i].radiant_flux[c] *= 2.0
		full.frame.lib.vkr_update_polygonal_light(C.byref(lights[i]))
	try:
		b = full.shade()
	finally:
		for i in range(full.frame.light_count):
			for c in range(3):
				lights[i].radiant_flux[c] *= 0.5
			full.frame.lib.vkr_update_polygonal_light(C.byref(lights[i]))
	marker = (a[..., 0] == 1.0) & (a[..., 1] == 0.0) & (np.abs(a[..., 2] - 0.8) < 1e-6)
	two_a = (a * np.float32(2.0)); two_a[..., 3] = 1.0
	ok = np.all(two_a.view(np.uint32) == b.view(np.uint32), axis=-1) | marker
	assert ok.all(), "%d pixels are not exactly doubled" % int((~ok).sum())
