"""The experiment list: structure of the table against the reference's (src/experiment_list.c), legality of every
entry under the C-ABI's rules, the synthetic data it needs, and -- through the CPU oracle at a tiny resolution -- that the
timing scenes light what they are meant to light. Host side only; the run itself needs a GPU (tests/test_gpu_screenshots_experiments.py)."""
import ctypes as C
import re

import numpy as np
import pytest

from tests import harness as H
from tests.ref_frames import host_constants
from vulkan_renderer_b200 import api, experiments as E


def test_timing_matrix_has_the_shape_and_names_of_the_reference_list():
	t = E.timing_experiments()
	assert len(t) == 5 * 2 * 2 * 13                                   # src/experiment_list.c:383-406
	names = [e["name"] for e in t]
	assert len(set(names)) == len(names)
	assert names[0] == "timings_central_3_128_baseline" and names[-1] == "timings_decentral_7_projected_solid_angle_biased_ours"
	for e in t:
		m = re.fullmatch(r"timings_(central|decentral)_([3-7])(_128)?_([a-z_]+)", e["name"])
		assert m and m.group(4) in E.SAMPLE_POLYGON_NAME
		many = m.group(3) is not None
		assert e["screenshot_path"] == "data/experiments/%s_%%.3f.png" % e["name"]
		assert e["quick_save_path"] == "data/quicksaves/roughness_planes_%s_%s%s.save" % (m.group(1), m.group(2), "_128" if many else "")
		s = e["settings"]
		assert (e["width"], e["height"]) == (1920, 1080) and s["sampling_strategies"] == api.STRATEGY_DIFFUSE_ONLY
		assert s["sample_count"] == (1 if many else 128) and e["light_count"] == (128 if many else 1)
		assert s["exposure_factor"] == 8.0 / e["light_count"] and not s["trace_shadow_rays"] and not s["show_polygonal_lights"]
		assert s["polygon_sampling_technique"] == E.SAMPLE_POLYGON_NAME.index(m.group(4))
	assert (E.SAMPLE_POLYGON_NAME.index("projected_solid_angle_ours"), E.SAMPLE_POLYGON_NAME.index("projected_solid_angle_biased_ours")) == (api.TECHNIQUE_PSA, api.TECHNIQUE_PSA_BIASED)
	figures = E.experiment_list(all_timings=False)
	assert figures[0]["name"] == "attic_solid_angle_and_ggx_mis_2spp" and len(figures) == 5 + 2 + 2 * 12 + 8 + 16 + 2 + 1 + 3
	assert len({e["name"] for e in figures}) == len(figures)
	assert E.experiment_list(all_figs=False) == t and len(E.experiment_list()) == len(figures) + 260
	by_name = {e["name"]: e for e in figures}
	assert by_name["error_attic_backward_times_psa"]["settings"]["error_display"] == api.ERROR_DISPLAY_DIFFUSE_BACKWARD_SCALED
	assert "bistro_tiny_polygon_bilinear_cosine_warp_clipping_hart_1spp" not in by_name and by_name["bistro_small_polygon_reference_128spp"]["settings"]["sample_count"] == 128
	assert by_name["mis_plane_optimal_ours_2spp"]["settings"]["mis_heuristic"] == api.MIS_OPTIMAL and by_name["shadowed_plane_biased_4096spp"]["settings"]["sample_count"] == 2048
	assert by_name["cornell_box_projected_solid_angle_arvo_tilted_1spp"]["quick_save_path"] == "data/quicksaves/cornell_box_tilted_light.save"
	names = [e["name"] for e in figures]   # the textured-light figures sit where the reference has them (src/experiment_list.c:294-362)
	assert names.index("shadowed_plane_biased_4096spp") + 1 == names.index("ies_profile_attic_2spp") and names[-1] == "roughness_planes_screen_2spp"
	assert by_name["ies_profile_attic_2spp"]["scene_parameters"] == dict(ies_profile=1) and by_name["roughness_planes_screen_2spp"]["settings"]["mis_heuristic"] == api.MIS_OPTIMAL_CLAMPED


def test_every_experiment_is_a_legal_configuration(capfd):
	"""vkr_create_shading_pass checks technique / strategy / heuristic before it touches CUDA; without tables it must get as far as
	complaining about those, never about the combination."""
	lib = api.load_library(); dev = api.Device()
	for e in E.experiment_list():
		s = e["settings"]
		p = api.ShadingPass(); d = api.ShadingPassDesc(width=e["width"], height=e["height"], polygonal_light_count=1, min_polygonal_light_vertex_count=4, max_polygonal_light_vertex_count=4,
			sample_count=s["sample_count"], sampling_strategies=s["sampling_strategies"], mis_heuristic=s["mis_heuristic"], polygon_sampling_technique=s["polygon_sampling_technique"], stripe_count=1,
			error_display=s.get("error_display", 0))
		assert lib.vkr_create_shading_pass(C.byref(p), C.byref(dev), C.byref(d)) == 1
		assert "missing LTC / noise tables" in capfd.readouterr().out, e["name"]


@pytest.mark.parametrize("name", ["mis_plane_weighted_ours_2spp", "cornell_box_projected_solid_angle_arvo_tilted_1spp", "roughness_planes_lambertian_2spp",
	"ies_profile_attic_2spp", "roughness_planes_screen_2spp"])
def test_figure_scene_data_loads_and_is_lit(name):
	e = [x for x in E.experiment_list(all_timings=False) if x["name"] == name][0]
	parameters = dict(e["scene_parameters"])
	if e["scene"] == "room": parameters.update(detail=6, clutter=60, n_mat=8)   # the same room with few triangles: the oracle's visibility pass is brute force
	info = H.dataset(e["scene"], **parameters)
	oi = H.OracleInputs(info)
	textured = [l.get("texturing_technique", 0) for l in info["lights"]]
	assert textured == {"ies_profile_attic_2spp": [3], "roughness_planes_screen_2spp": [1]}.get(name, [0] * len(info["lights"]))
	assert (oi.light_textures is not None) == any(textured)
	width, height = 64, 48
	lights = len(info["lights"]); vertices = max(len(l["vertices"]) for l in info["lights"])
	constants = host_constants(info, width, height, lights)
	vis = oi.visibility(width, height, constants); gb = oi.gbuffer(width, height, constants, vis)
	s = e["settings"]
	cfg = dict(width=width, height=height, light_count=lights, max_light_vertex_count=vertices, min_light_vertex_count=vertices, sample_count=2, sampling_strategies=s["sampling_strategies"],
		mis_heuristic=s["mis_heuristic"], biased_sampling=0, trace_shadow_rays=1, show_polygonal_lights=1, polygon_sampling_technique=min(s["polygon_sampling_technique"], 11))
	out, rays = oi.shade(cfg, constants, gb)
	lit = out[..., :3].sum(-1) > 0
	assert (vis != 0xFFFFFFFF).mean() > 0.35 and 0.15 < lit.mean() and rays > 0
	if e["scene"] == "shadowed_plane":
		assert lit.mean() < (vis != 0xFFFFFFFF).mean()      # something is in shadow or faces away


@pytest.mark.parametrize("vertices,central,lights", [(3, 1, 128), (5, 0, 1), (7, 1, 1), (4, 0, 128)])
def test_timing_scene_data_loads_and_is_lit(vertices, central, lights):
	e = [x for x in E.timing_experiments() if x["scene_parameters"] == dict(vertices=vertices, central=central, lights=lights)][0]
	info = H.dataset(e["scene"], **e["scene_parameters"])
	lib = api.load_library(); spec = api.SceneSpecification()
	assert lib.vkr_quick_load(C.byref(spec), info["save"].encode()) == 0
	assert spec.polygonal_light_count == lights and all(spec.polygonal_lights[i].vertex_count == vertices for i in range(lights))
	lib.vkr_destroy_scene_specification(C.byref(spec))
	width, height = 64, 36
	constants = host_constants(info, width, height, lights)
	assert len(constants) == 256 + lights * (160 + 16 * vertices * 2 + 16 * (vertices - 2))
	oi = H.OracleInputs(info)
	vis = oi.visibility(width, height, constants); gb = oi.gbuffer(width, height, constants, vis)
	assert (vis != 0xFFFFFFFF).mean() > 0.8
	cfg = dict(width=width, height=height, light_count=lights, max_light_vertex_count=vertices, min_light_vertex_count=vertices, sample_count=2, sampling_strategies=0, mis_heuristic=0,
		biased_sampling=0, trace_shadow_rays=0, show_polygonal_lights=0)
	psa, _ = oi.shade(cfg, constants, gb)
	assert np.isfinite(psa).all() and (psa[..., :3].sum(-1) > 0).mean() > 0.5
	if lights == 1:   # "central": the surface normal passes through the light for a large part of the pixels (the zenith case of the sampler); "decentral": for none
		b = np.frombuffer(constants[256:], dtype=np.float32)
		plane = b[16:20]; verts = b[40 + 4 * vertices: 40 + 8 * vertices].reshape(vertices, 4)[:, :3].astype(np.float64)
		pos = gb[0][..., :3].reshape(-1, 3)[(vis != 0xFFFFFFFF).reshape(-1)].astype(np.float64)
		hit = pos + np.outer((-(pos @ plane[:3].astype(np.float64)) - plane[3]) / plane[2], [0.0, 0.0, 1.0])   # along the surface normal +z
		inside = np.ones(len(pos), dtype=bool); sign = None
		for k in range(vertices):
			a, c = verts[k], verts[(k + 1) % vertices]
			side = np.cross(c - a, hit - a)[:, 2]
			sign = np.sign(np.median(side)) if sign is None else sign
			inside &= side * sign >= 0
		assert (inside.mean() > 0.4) if central else (inside.mean() == 0.0)
	# another technique gives the same picture up to noise: the quicksave is usable by the whole matrix
	other, _ = oi.shade(dict(cfg, polygon_sampling_technique=4, sample_count=4), constants, gb)
	assert abs(float(other[..., :3].mean()) / float(psa[..., :3].mean()) - 1.0) < 0.35
