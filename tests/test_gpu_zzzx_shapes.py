"""Parity at the SHAPES of the benchmark configurations that round 1 had only tested small (BASELINE.json configs 2, 3 and 4), bit for bit against the oracle:

  C2  city 1920x1080, 1 quad light, 4 spp, diffuse-only projected solid angle sampling, shadow rays: the WHOLE frame
  C3  city 1920x1080, 8 lights, 64 spp, clamped optimal MIS: 8-row bands every 128 rows over the whole frame (test_gpu_properties.py has the determinism,
      linearity and multi-GPU split of the same frame)
  C4  room 3840x2160, 32 lights: the whole frame at 1 spp, and one tile row of the 256 spp frame (SAMPLE_COUNT_CLAMPED = 33: the reference's shader
      loops instead of unrolling, src/shaders/unrolling.glsl:61; 10.5 KB of constants in shared memory, 8192 noise fetches per pixel)
The oracle runs on the host cores of the GPU box (OpenMP); sizes are chosen so that each comparison takes seconds."""
import numpy as np
import pytest

from tests import harness as H
from vulkan_renderer_b200 import api

pytestmark = pytest.mark.gpu


def _banded(cfg, stride):
	return dict(cfg, band_height=8, band_stride=stride)


def _band_rows(height, stride):
	return np.array([y for y in range(height) if y % stride < 8])


def test_c2_whole_frame_against_the_oracle():
	info = H.dataset("city")
	frame = H.open_frame(info)
	try:
		w, h = 1920, 1080
		frame.configure(sample_count=4, strategy=api.STRATEGY_DIFFUSE_ONLY, trace_shadow_rays=1, show_lights=1, light_count=1)
		constants = frame.constants(w, h)
		_, gb = frame.gbuffer_host(w, h)
		out = frame.shade_host(w, h, gb)
		oi = H.OracleInputs(info)
		ref, rays = oi.shade(H.oracle_config(frame, w, h), constants, gb)
		assert rays > 1000000
		assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), H.compare_radiance(out, ref)
	finally:
		frame.close()


def test_c3_bands_over_the_whole_frame_against_the_oracle():
	info = H.dataset("city")
	frame = H.open_frame(info)
	try:
		w, h = 1920, 1080
		frame.configure(sample_count=64, strategy=api.STRATEGY_DIFFUSE_SPECULAR_MIS, heuristic=api.MIS_OPTIMAL_CLAMPED, trace_shadow_rays=1, show_lights=1, light_count=8)
		constants = frame.constants(w, h)
		_, gb = frame.gbuffer_host(w, h)
		out = frame.shade_host(w, h, gb)
		oi = H.OracleInputs(info)
		ref, rays = oi.shade(_banded(H.oracle_config(frame, w, h), 128), constants, gb)
		rows = _band_rows(h, 128)
		assert len(rows) == 72 and rays > 10000000
		assert np.array_equal(out[rows].view(np.uint32), ref[rows].view(np.uint32)), H.compare_radiance(out[rows], ref[rows])
	finally:
		frame.close()


@pytest.fixture(scope="module")
def room():
	info = H.dataset("room")
	frame = H.open_frame(info)
	w, h = 3840, 2160
	frame.configure(sample_count=1, strategy=api.STRATEGY_DIFFUSE_SPECULAR_MIS, heuristic=api.MIS_OPTIMAL_CLAMPED, trace_shadow_rays=1, show_lights=1, light_count=32)
	_, gb = frame.gbuffer_host(w, h)
	yield info, frame, w, h, gb
	frame.close()


def test_c4_whole_frame_at_one_sample_against_the_oracle(room):
	info, frame, w, h, gb = room
	frame.configure(sample_count=1)
	constants = frame.constants(w, h)
	out = frame.shade_host(w, h, gb)
	oi = H.OracleInputs(info)
	ref, rays = oi.shade(H.oracle_config(frame, w, h), constants, gb)
	assert rays > 10000000
	assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), H.compare_radiance(out, ref)
	assert float(out[..., :3].mean()) > 0.001 and (gb[1, :, :, 3] != 0).mean() > 0.9


def test_c4_bands_at_256_samples_against_the_oracle(room):
	info, frame, w, h, gb = room
	frame.configure(sample_count=256)
	constants = frame.constants(w, h)
	out = frame.shade_host(w, h, gb)
	oi = H.OracleInputs(info)
	ref, rays = oi.shade(H.oracle_config(frame, w, h), constants, gb, row_begin=1080, row_end=1088)   # one tile row in the middle: 8 x 3840 pixels x 32 lights x 256 spp x 2 techniques
	assert rays > 10000000
	assert np.array_equal(out[1080:1088].view(np.uint32), ref[1080:1088].view(np.uint32)), H.compare_radiance(out[1080:1088], ref[1080:1088])
