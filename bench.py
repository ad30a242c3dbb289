#!/usr/bin/env python3
"""Benchmark of the shading pass hot path (BASELINE.json: Msamples/s at 1920x1080x64spp).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one pass of the shading megakernel over one 1920x1080 frame of the synthetic
Bistro-like city (2.8 M triangles, 8 quad lights, 64 spp, diffuse+specular MIS with the clamped
optimal heuristic, shadow rays on) -- BASELINE config 3, the one the metric is quoted on.

  value  whole-job Msamples/s (pixels*spp / time), inputs resident in HBM, CUDA events on the
         launching stream, L2 flushed between steps, max over ranks
  e2e    same metric through the C-ABI call with HOST buffers (H2D of the G-buffer and D2H of the
         frame inside the timed region)
  N > 1  the frame is sharded by interleaved 8-pixel tile rows (strong scaling), one NCCL all-gather
         of the HDR stripes per step inside the timed region

--impl reference times the CPU restatement of the reference's path (oracle/, all host threads) on a
bounded sample of the same frame; the reference itself is GLSL + Vulkan and cannot run on this box.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
	# name: (dataset, dataset overrides, width, height, lights, spp)
	"C3": ("city", {}, 1920, 1080, 8, 64),
	"C2": ("city", {}, 1920, 1080, 1, 4),
	"C1": ("cornell", {}, 256, 256, 1, 1),
	"mini": ("mini_city", {}, 320, 192, 3, 8),
}


def log(*a):
	print(*a, file=sys.stderr, flush=True)


class ClockSampler:
	"""Samples nvidia-smi clocks and throttle reasons while the timed region runs."""
	QUERY = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

	def __init__(self, gpu_index):
		self.gpu_index = gpu_index; self.samples = []; self.proc = None; self.thread = None

	def start(self):
		try:
			self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
		except OSError:
			self.proc = None
			return
		def reader():
			for line in self.proc.stdout:
				parts = [p.strip() for p in line.split(",")]
				if len(parts) >= 9:
					self.samples.append(parts)
		self.thread = threading.Thread(target=reader, daemon=True); self.thread.start()

	def stop(self):
		if self.proc is None:
			return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
		time.sleep(0.15)
		self.proc.terminate()
		try:
			self.proc.wait(timeout=2)
		except subprocess.TimeoutExpired:
			self.proc.kill()
		clocks, max_clocks, reasons = [], [], set()
		for p in self.samples:
			try:
				clocks.append(float(p[1])); max_clocks.append(float(p[2]))
			except ValueError:
				continue
			for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
				if val.lower().startswith("active"):
					reasons.add(name)
		return {"sm_mhz": float(np.median(clocks)) if clocks else None, "sm_max_mhz": max(max_clocks) if max_clocks else None,
			"reasons": sorted(reasons), "samples": len(clocks), "power_w_max": max([float(p[3]) for p in self.samples if p[3].replace(".", "", 1).isdigit()] or [0.0])}


def algorithmic_bytes(width, height, tri_count, light_count, noise_fetches_per_pixel, ltc_res, ltc_layers_touched):
	"""compulsory bytes per frame, every byte counted once."""
	return (width * height * (64 + 16)
		+ 64 * (tri_count - 1) + 48 * tri_count
		+ 256 + 320 * light_count
		+ min(33554432, width * height * noise_fetches_per_pixel * 8)
		+ ltc_res * ltc_res * 12 * ltc_layers_touched)


def measured_peak():
	path = os.path.join(ROOT, "MEASURED_PEAKS.json")
	if os.path.exists(path):
		with open(path) as f:
			return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
	return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_counters(workload):
	"""Counters of the shading kernel from the committed `ncu --set full` capture of this workload (profiles/kernel_counters.json,
	written by tools/summarize_ncu.py): DRAM bytes per launch and the utilisation of the units that actually bound it."""
	path = os.path.join(ROOT, "profiles", "kernel_counters.json")
	if os.path.exists(path):
		with open(path) as f:
			return json.load(f).get(workload) or {}
	return {}


def build_frame(workload, cuda_device, stream, host_only=False):
	from vulkan_renderer_b200 import Frame, api, synth
	dataset, overrides, width, height, lights, spp = WORKLOADS[workload]
	data_root = os.environ.get("VKR_BENCH_DATA", os.path.join("/tmp", "vkr_b200_data"))
	t0 = time.time()
	info = synth.build_dataset(os.path.join(data_root, dataset), dataset, **overrides)
	log("[bench] dataset %s: %d triangles (%.1f s)" % (dataset, info["triangle_count"], time.time() - t0))
	return info, (width, height, lights, spp)


def run_b200(args):
	import torch
	import torch.distributed as dist
	from vulkan_renderer_b200 import Frame, api
	world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
	if world > 1:
		os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
		dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
	torch.cuda.set_device(local_rank)
	dev = torch.device("cuda", local_rank)
	# A dedicated stream shared by torch and the library: the default stream's handle is 0, which the C-ABI
	# reads as "create your own stream"; CUDA events must be recorded on the stream the kernels run on.
	stream = torch.cuda.Stream(dev)
	torch.cuda.set_stream(stream)
	assert stream.cuda_stream != 0
	if world > 1 and rank != 0:
		dist.barrier()  # rank 0 writes the dataset first
	info, (width, height, lights, spp) = build_frame(args.workload, local_rank, stream)
	if world > 1 and rank == 0:
		dist.barrier()
	frame = Frame(info["vks"], info["textures"], info["save"], info["ltc"], cuda_device=local_rank, stream=stream.cuda_stream)
	frame.configure(sample_count=spp, strategy=api.STRATEGY_DIFFUSE_SPECULAR_MIS, heuristic=api.MIS_OPTIMAL_CLAMPED, technique=api.TECHNIQUE_PSA, trace_shadow_rays=1, show_lights=1, light_count=lights)
	lib = frame.lib
	log("[bench] rank %d: BVH %d node pairs, depth %d, build %.2f s on the host" % (rank, frame.scene.shadow_node_count, frame.scene.shadow_max_depth, frame.scene.build_seconds))
	constants = frame.constants(width, height)
	# --- inputs: the G-buffer is produced on the device once, outside the timed region
	vis = torch.empty((height, width), dtype=torch.int32, device=dev)
	gb = torch.empty((4, height, width, 4), dtype=torch.float32, device=dev)
	out = torch.zeros((height, width, 4), dtype=torch.float32, device=dev)
	assert lib.vkr_run_visibility_pass(C.byref(frame.device), C.byref(frame.scene), constants, width, height, vis.data_ptr()) == 0
	assert lib.vkr_run_gbuffer_pass(C.byref(frame.device), C.byref(frame.scene), constants, width, height, vis.data_ptr(), gb.data_ptr()) == 0
	torch.cuda.synchronize()
	valid = gb[1, :, :, 3] != 0
	f0_lum = (gb[3, :, :, :3] * torch.tensor([0.2126, 0.7152, 0.0722], device=dev)).sum(-1)
	ltc_layers = int(torch.unique(torch.round(f0_lum[valid].clamp(0, 1) * 50.0)).numel()) if bool(valid.any()) else 0
	p = frame.create_pass(width, height, stripe_index=rank, stripe_count=world)
	from vulkan_renderer_b200.stripes import StripeGather, stripe_rows
	sg = StripeGather(height, width, rank, world, dev)
	row_idx = sg.my_rows
	flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

	def step_device():
		rc = lib.vkr_shading_pass_run(C.byref(p), C.byref(frame.device), constants, len(constants), gb.data_ptr(), out.data_ptr())
		assert rc == 0
		sg.gather_frame(out)   # one NCCL all-gather of the HDR stripes (no-op on a single GPU)

	def timed(step_fn, steps, warmup):
		for _ in range(warmup):
			flush.zero_(); step_fn()
		torch.cuda.synchronize()
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()
		events = []
		for _ in range(steps):
			flush.zero_()
			e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
			e0.record(stream); step_fn(); e1.record(stream)
			events.append((e0, e1))
		torch.cuda.synchronize()
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()
		total_ms = sum(a.elapsed_time(b) for a, b in events)
		t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
		if world > 1:
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
		return float(t.item())

	def launches_of(step_fn, steps, warmup):
		"""timed() plus the number of kernels of libvkr_b200.so launched inside the timed region (the library counts them)."""
		for _ in range(warmup):
			flush.zero_(); step_fn()
		before = int(p.kernel_launches)
		ms = timed(step_fn, steps, 0)
		return ms, int(p.kernel_launches) - before

	# kernel-only timing (per launch, CUDA events inside the library on the launching stream)
	p.timing_enabled = 1
	sampler = ClockSampler(local_rank)
	if rank == 0:
		sampler.start()
	total_ms, launches = launches_of(step_device, args.steps, args.warmup)
	clocks = sampler.stop() if rank == 0 else None
	launches *= world   # every rank launches its stripe's kernel
	# one more step to read the kernel's own duration
	flush.zero_(); step_device(); lib.vkr_shading_pass_wait(C.byref(p), C.byref(frame.device))
	kernel_ms = float(p.last_kernel_ms)
	ms_per_step = total_ms / args.steps
	samples = width * height * spp
	value = samples / (ms_per_step * 1e-3) / 1e6

	# --- e2e: host buffers in, host buffers out
	gb_host = torch.empty((4, height, width, 4), dtype=torch.float32).pin_memory(); gb_host.copy_(gb)
	out_host = torch.zeros((height, width, 4), dtype=torch.float32).pin_memory()
	stripe_row_count = len(stripe_rows(height, rank, world))
	row_idx_cpu = row_idx.cpu()

	def step_e2e():
		if world == 1:
			rc = lib.vkr_shading_pass_run_host(C.byref(p), C.byref(frame.device), constants, len(constants), gb_host.data_ptr(), out_host.data_ptr())
			assert rc == 0
		else:
			# stripe rows host->device, shade, gather over NVLink, rank 0 reads the frame back
			gb.index_copy_(1, row_idx, gb_host.index_select(1, row_idx_cpu).to(dev, non_blocking=True))
			step_device()
			if rank == 0:
				out_host.copy_(out, non_blocking=True)
	e2e_steps = max(1, min(args.steps, 5)); e2e_warm = max(1, min(args.warmup, 2))
	e2e_ms = timed(step_e2e, e2e_steps, e2e_warm) / e2e_steps
	e2e_value = samples / (e2e_ms * 1e-3) / 1e6
	h2d = (4 * stripe_row_count * width * 16) + len(constants) if world > 1 else 4 * height * width * 16 + len(constants)
	d2h = height * width * 16

	result = None
	if rank == 0:
		tri_count = int(frame.scene.triangle_count)
		fetches = min(lights * spp * 2 // 2, 128) if spp * lights > 0 else 0
		fetches = lights * spp  # one RGBA16 texel per diffuse+specular pair of 2D numbers
		bytes_alg = algorithmic_bytes(width, height, tri_count, lights, fetches, int(frame.ltc.roughness_count), ltc_layers)
		peak, peak_kind = measured_peak()
		counters = recorded_counters(args.workload)
		achieved = bytes_alg / (kernel_ms * 1e-3) / 1e9 if world == 1 else bytes_alg / (ms_per_step * 1e-3) / 1e9
		result = {
			"metric": "Msamples/s (pixels x spp) at 1920x1080x64spp; achieved HBM GB/s vs roofline",
			"value": round(value, 3), "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
			"higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
			"config": {"workload": "%s: Bistro-like synthetic city %dx%d, %d quad lights, %d spp, diffuse+specular MIS (clamped optimal), shadow rays on, %d triangles" % (args.workload, width, height, lights, spp, tri_count),
				"parallelism": "interleaved 8-px tile rows over %d GPU(s), NCCL all-gather of HDR stripes" % world if world > 1 else "single GPU",
				"l2": "flushed between steps (512 MiB memset); inputs 270 MB > 126 MB L2", "rays_per_sample_pair": 2, "sample_pairs": width * height * lights * spp},
			"e2e": {"value": round(e2e_value, 3), "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": round(e2e_ms, 4)},
			"gpu_launches": launches,
			"kernel_ms": round(kernel_ms, 4),
			"roofline": {"bound": "hbm", "achieved": round(achieved, 3), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 6), "traffic": counters.get("dram_bytes_per_launch"),
				"peak_source": peak_kind, "algorithmic_bytes": int(bytes_alg),
				"issue_active_frac": counters.get("issue_active_frac"), "l1_data_pipe_frac": counters.get("l1_data_pipe_frac"), "counters_from": counters.get("source"),
				"note": "not an HBM-bound path (SURVEY 8d): compulsory traffic is ~0.5 GB per frame against ~1.4 G shadow rays; the kernel is bound by instruction issue and the L1 data pipe (BVH node fetches that hit L1), see profiles/"},
			"clocks": clocks,
		}
		if world == 1 and not args.no_cpu_baseline:
			result["cpu_baseline"] = cpu_baseline(args, info, width, height, lights, spp, constants, gb.cpu().numpy(), visibility=vis.cpu().numpy().view(np.uint32))
	frame.destroy_pass(p)
	frame.close()
	if world > 1:
		dist.barrier()
		dist.destroy_process_group()
	if rank == 0:
		print(json.dumps(result), flush=True)


def cpu_baseline(args, info, width, height, lights, spp, constants, gbuffer, band_rows=None, visibility=None):
	"""Times the reference's path on the host cores on a bounded sample: 8-row bands spread over the frame, full light count and spp.
	kind "reference": the reference's own shader sources compiled for the CPU (oracle/_ref/libref_shader.so, built by
	oracle/build_ref.py where /root/reference exists and shipped prebuilt; it starts from the visibility buffer like the
	shader does, i.e. it includes get_shading_data). kind "port": the C restatement (oracle/) when that library or this
	configuration is not available. Both use OpenMP over rows with all host threads."""
	from tests import harness as H
	from vulkan_renderer_b200 import api
	oi = H.OracleInputs(info)
	band_stride = 8 * (args.cpu_band_stride if band_rows is None else band_rows)
	rows = sum(1 for y in range(height) if y % band_stride < 8)
	ref_cfg = None
	if not args.cpu_port:
		try:
			from oracle import ref_binding as R
			ref_cfg = R.find_config(strategy=api.STRATEGY_DIFFUSE_SPECULAR_MIS, heuristic=api.MIS_OPTIMAL_CLAMPED, biased=0, lights=lights, max_vertices=4, min_vertices=4,
				samples=spp, trace=1, show_lights=1, technique=11, srgb=0, frame_bits=0)
			if ref_cfg is not None and ref_cfg["materials"] < len(oi.material_params):
				ref_cfg = None
		except Exception as e:   # a broken prebuilt library must not take the bench down
			log("[bench] reference shader library unusable (%s); timing the C restatement instead" % e)
			ref_cfg = None
	t0 = time.time()
	if ref_cfg is not None:
		if visibility is None:
			visibility = oi.visibility(width, height, constants)
		t0 = time.time()
		R.shade(ref_cfg["entry"], width, height, ref_cfg, constants, visibility, oi.vks, oi.material_params, oi.noise, oi.ltc0, oi.ltc1, oi.shadow_tris, band_height=8, band_stride=band_stride)
		seconds = R.last_shade_seconds(); cores = R.thread_count(); kind = "reference"
		what = "the reference's shader sources (shading_pass.frag.glsl + includes) compiled as C++ with g++ -O2, fp32, OpenMP, ray queries on a CPU BVH"
		log("[bench] cpu reference shader: %d rows in %.2f s (+ %.1f s BVH build)" % (rows, seconds, time.time() - t0 - seconds))
	else:
		cfg = dict(width=width, height=height, light_count=lights, max_light_vertex_count=4, min_light_vertex_count=4, sample_count=spp,
			sampling_strategies=api.STRATEGY_DIFFUSE_SPECULAR_MIS, mis_heuristic=api.MIS_OPTIMAL_CLAMPED, biased_sampling=0, trace_shadow_rays=1, show_polygonal_lights=1,
			row_begin=0, row_end=0, band_height=8, band_stride=band_stride)
		_, rays = H.oracle.shade(cfg, constants, gbuffer, oi.noise, oi.ltc0, oi.ltc1, oi.shadow_tris)
		seconds = H.oracle.last_shade_seconds(); cores = H.oracle.thread_count(); kind = "port"
		what = "scalar fp32 C oracle, OpenMP"
		log("[bench] cpu oracle: %d rows in %.2f s (+ %.1f s BVH build), %d shadow rays" % (rows, seconds, time.time() - t0 - seconds, rays))
	value = rows * width * spp / seconds / 1e6
	return {"value": round(value, 4), "unit": "Msamples/s", "cores": cores, "kind": kind,
		"sample": "%d of %d rows (8-row bands every %d rows), all %d lights, %d spp, %s; BVH build excluded" % (rows, height, band_stride, lights, spp, what),
		"seconds": round(seconds, 3)}


def run_reference(args):
	"""The reference's own implementation of this path is a GLSL fragment shader driven through Vulkan (no ICD, no glslangValidator
	on this box). Its CPU-runnable form is that shader compiled as C++ (oracle/_ref, kind 'reference', see cpu_baseline); without the
	prebuilt library the C restatement is timed (kind 'port'). Rank 0 only."""
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return
	from tests import harness as H
	from vulkan_renderer_b200 import api
	info, (width, height, lights, spp) = build_frame(args.workload, 0, None)
	lib = api.load_library()
	scene = api.Scene(); ltc = api.LtcTable(); noise = api.NoiseTable(); spec = api.SceneSpecification(); st = api.RenderSettings()
	assert lib.vkr_load_scene(C.byref(scene), None, info["vks"].encode(), info["textures"].encode(), 0) == 0
	assert lib.vkr_load_ltc_table(C.byref(ltc), None, info["ltc"].encode(), 51) == 0
	assert lib.vkr_load_noise_table(C.byref(noise), None, 256, 256, 64, 0) == 0
	assert lib.vkr_quick_load(C.byref(spec), info["save"].encode()) == 0
	spec.polygonal_light_count = lights
	lib.vkr_specify_default_render_settings(C.byref(st)); st.animate_noise = 0; st.exposure_factor = 1.0; st.sample_count = spp
	size = lib.vkr_get_constants_size(C.byref(spec)); buf = (C.c_uint8 * size)()
	lib.vkr_write_constants(buf, C.byref(spec), C.byref(st), C.byref(scene), C.byref(ltc), C.byref(noise), width, height)
	constants = bytes(buf)
	oi = H.OracleInputs(info)
	t0 = time.time()
	vis = oi.visibility(width, height, constants)
	gb = oi.gbuffer(width, height, constants, vis)
	log("[bench] reference arm: oracle G-buffer in %.1f s" % (time.time() - t0))
	values = []
	for i in range(args.warmup + args.steps):
		r = cpu_baseline(args, info, width, height, lights, spp, constants, gb, band_rows=args.cpu_band_stride * 2, visibility=vis)
		if i >= args.warmup:
			values.append(r)
	seconds = sum(r["seconds"] for r in values)
	value = float(np.mean([r["value"] for r in values]))
	base = values[-1]
	print(json.dumps({
		"impl": "reference", "metric": "Msamples/s (pixels x spp) at 1920x1080x64spp; achieved HBM GB/s vs roofline", "value": round(value, 4), "unit": "Msamples/s",
		"n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * seconds / max(1, len(values)), 3),
		"higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
		"config": {"workload": "%s: Bistro-like synthetic city %dx%d, %d quad lights, %d spp, diffuse+specular MIS (clamped optimal), shadow rays on, %d triangles" % (args.workload, width, height, lights, spp, info["triangle_count"]),
			"note": "each step = a bounded sample of the frame on the host cores; the reference's GLSL/Vulkan path itself cannot run here (no Vulkan ICD / glslangValidator)"},
		"cpu_baseline": {"value": round(value, 4), "unit": "Msamples/s", "cores": base["cores"], "kind": base["kind"], "sample": base["sample"]},
		"e2e": {"value": round(value, 4), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
	}), flush=True)


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=5)
	ap.add_argument("--warmup", type=int, default=3)
	ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
	ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
	ap.add_argument("--no-cpu-baseline", action="store_true")
	ap.add_argument("--cpu-port", action="store_true", help="time the C restatement (oracle/) on the CPU legs even if the compiled reference shader is available")
	ap.add_argument("--cpu-band-stride", type=int, default=4, help="the CPU sample takes one 8-row band every this many tile rows (the reference arm: twice as many)")
	args = ap.parse_args()
	if args.impl == "reference":
		run_reference(args)
	else:
		run_b200(args)


if __name__ == "__main__":
	main()
