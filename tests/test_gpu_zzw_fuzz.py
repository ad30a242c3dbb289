"""Random frames on the GPU (tools/fuzz_parity.py's generator): random shader configuration, camera, light transforms and fluxes, exposure, roughness factor,
MIS visibility estimate, frame size -- the kernels WITH shadow rays against the oracle, bit for bit, pink (NaN) pixels included. On the CPU the same
generator compares the reference shader, the oracle and the device code compiled for the CPU over thousands of frames; this is the part only a GPU can
do: the warp-level ray streams under degenerate inputs. Written after this round's GPU budget was spent (hence late in the order)."""
import os
import sys

import numpy as np
import pytest

from tests import harness as H
from tests.ref_frames import dataset_for
from vulkan_renderer_b200 import api

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fuzz_parity  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,wild,any_config", [(1, False, False), (2, True, False), (3, True, False), (4, False, True), (5, True, True)])
def test_kernels_match_the_oracle_on_random_frames(seed, wild, any_config):
	"""any_config: any legal combination of the run-time settings instead of the configurations a reference shader was compiled for."""
	rng = np.random.default_rng(1000 + seed)
	configs = [c for c in fuzz_parity.fixture_configs() if c["samples"] <= 8]
	pink_frames = 0
	for k in range(14):
		cfg = fuzz_parity.random_config(rng) if any_config else configs[int(rng.integers(len(configs)))]
		info = H.dataset(cfg.get("dataset") or dataset_for(cfg)); oi = H.OracleInputs(info)
		width, height = 48 + int(rng.integers(0, 40)), 32 + int(rng.integers(0, 24))
		frame = H.open_frame(info)
		try:
			fuzz_parity.perturb(frame.lib, frame.spec, frame.settings, info, cfg, rng, wild)
			frame.settings.animate_noise = 0
			technique = cfg["technique"] if cfg["technique"] != api.TECHNIQUE_PSA else (api.TECHNIQUE_PSA_BIASED if cfg["biased"] else api.TECHNIQUE_PSA)
			frame.configure(sample_count=cfg["samples"], strategy=cfg["strategy"], heuristic=cfg["heuristic"], technique=technique, trace_shadow_rays=cfg["trace"],
				show_lights=cfg["show_lights"], light_count=cfg["lights"], output_srgb=cfg["srgb"], frame_bits=cfg["frame_bits"], error_display=cfg["error_display"])
			constants = frame.constants(width, height)
			vis = oi.visibility(width, height, constants); gb = oi.gbuffer(width, height, constants, vis)
			out = frame.shade_host(width, height, gb)
			ref, _ = oi.shade(H.oracle_config(frame, width, height), constants, gb)
		finally:
			frame.close()
		assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), (seed, k, cfg["name"], width, height, H.compare_radiance(out, ref))
		pink_frames += int(((ref[..., 1] == 0) & (ref[..., 0] > 0) & (ref[..., 2] > 0)).any())
	assert pink_frames >= 0
