"""Shared test harness: synthetic datasets, an independent numpy reading of the reference's file
formats for the oracle side, and helpers that run the same frame through the oracle and through
the CUDA library (C-ABI). Used by tests/, __graft_entry__.smoke() and bench.py's CPU legs only.
"""
import ctypes as C
import os
import struct
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import sys
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)

from oracle import binding as oracle  # noqa: E402
from vulkan_renderer_b200 import api, synth  # noqa: E402

_DATA_ROOT = os.environ.get("VKR_TEST_DATA", os.path.join(tempfile.gettempdir(), "vkr_b200_data"))
_cache = {}


def dataset(name, **overrides):
	"""Builds (once per process and per parameter set) a synthetic dataset on disk."""
	key = (name, tuple(sorted(overrides.items())))
	if key not in _cache:
		tag = name + "".join("_%s%s" % (k, v) for k, v in sorted(overrides.items()))
		directory = os.path.join(_DATA_ROOT, tag)
		_cache[key] = synth.build_dataset(directory, name, **overrides)
	return _cache[key]


# ---------------------------------------------------------------------------------------------
# independent numpy readers of the reference's formats (oracle side; byte/integer work)
# ---------------------------------------------------------------------------------------------

def read_vks(path):
	"""src/scene.c:419-483"""
	with open(path, "rb") as f:
		marker, version = struct.unpack("<II", f.read(8))
		assert marker == 0xABCABC and version == 1
		n_mat, n_tri = struct.unpack("<QQ", f.read(16))
		factor = np.frombuffer(f.read(12), dtype="<f4").copy(); summand = np.frombuffer(f.read(12), dtype="<f4").copy()
		names = []
		for _ in range(n_mat):
			(length,) = struct.unpack("<Q", f.read(8))
			names.append(f.read(length + 1)[:-1].decode())
		positions = np.frombuffer(f.read(8 * 3 * n_tri), dtype="<u4").reshape(-1, 2).copy()
		normals_uv = np.frombuffer(f.read(8 * 3 * n_tri), dtype="<u2").reshape(-1, 4).copy()
		material_indices = np.frombuffer(f.read(n_tri), dtype=np.uint8).copy()
		(eof,) = struct.unpack("<I", f.read(4))
		assert eof == 0xE0FE0F
	return dict(triangle_count=n_tri, factor=factor, summand=summand, names=names, positions=positions, normals_uv=normals_uv, material_indices=material_indices)


def read_vkt(path):
	"""All mip levels of a *.vkt file as RGBA float32 arrays, decoded here in numpy (src/textures.c:111-169; formats of tools/texture_conversion)."""
	raw = open(path, "rb").read()
	marker, version, mips, width, height, vk_format, payload_size = struct.unpack_from("<IIIIIIQ", raw, 0)
	assert marker == 0xBC1BC1 and version == 1
	headers = [struct.unpack_from("<IIQQ", raw, 32 + 24 * k) for k in range(mips)]
	base = 32 + 24 * mips
	assert struct.unpack_from("<I", raw, base + payload_size)[0] == 0xE0FE0F
	levels = []
	for (w, h, size, offset) in headers:
		data = raw[base + offset: base + offset + size]
		out = np.zeros((h, w, 4), dtype=np.float32); out[..., 3] = 1.0
		if vk_format in (97, 90):
			c = 4 if vk_format == 97 else 3
			out[..., :c] = np.frombuffer(data, dtype="<f2", count=w * h * c).reshape(h, w, c).astype(np.float32)
		elif vk_format in (109, 106):
			c = 4 if vk_format == 109 else 3
			out[..., :c] = np.frombuffer(data, dtype="<f4", count=w * h * c).reshape(h, w, c)
		elif vk_format == 131:
			bw = (w + 3) // 4
			for by in range((h + 3) // 4):
				for bx in range(bw):
					c0, c1, bits = struct.unpack_from("<HHI", data, 8 * (by * bw + bx))
					un = lambda c: np.array([np.float32((c >> 11) & 31) / np.float32(31.0), np.float32((c >> 5) & 63) / np.float32(63.0), np.float32(c & 31) / np.float32(31.0)], dtype=np.float32)
					a, b = un(c0), un(c1)
					pal = [a, b, (np.float32(2.0) * a + b) / np.float32(3.0), (a + np.float32(2.0) * b) / np.float32(3.0)] if c0 > c1 else [a, b, np.float32(0.5) * (a + b), np.zeros(3, dtype=np.float32)]
					for t in range(16):
						x, y = 4 * bx + (t & 3), 4 * by + (t >> 2)
						if x < w and y < h: out[y, x, :3] = pal[(bits >> (2 * t)) & 3]
		elif vk_format == 141:
			bw = (w + 3) // 4
			def bc4(off):
				r0, r1 = np.float32(data[off]) / np.float32(255.0), np.float32(data[off + 1]) / np.float32(255.0)
				if data[off] > data[off + 1]: pal = [r0, r1] + [(np.float32(8 - i) * r0 + np.float32(i - 1) * r1) / np.float32(7.0) for i in range(2, 8)]
				else: pal = [r0, r1] + [(np.float32(6 - i) * r0 + np.float32(i - 1) * r1) / np.float32(5.0) for i in range(2, 6)] + [np.float32(0.0), np.float32(1.0)]
				bits = int.from_bytes(data[off + 2:off + 8], "little")
				return [pal[(bits >> (3 * t)) & 7] for t in range(16)]
			for by in range((h + 3) // 4):
				for bx in range(bw):
					red, green = bc4(16 * (by * bw + bx)), bc4(16 * (by * bw + bx) + 8)
					for t in range(16):
						x, y = 4 * bx + (t & 3), 4 * by + (t >> 2)
						if x < w and y < h: out[y, x, 0] = red[t]; out[y, x, 1] = green[t]
		else:
			raise ValueError("VkFormat %d" % vk_format)
		levels.append(out)
	return levels


def material_texture_set(info):
	"""(dims uint32 [T,3], offsets uint64 [T], data float32) of the 3 textures per material, in material order: oracle/texture_filter.h's input."""
	dims, offsets, chunks, at = [], [], [], 0
	for m in info["materials"]:
		for suffix in ("BaseColor", "Specular", "Normal"):
			levels = read_vkt(os.path.join(info["textures"], "%s_%s.vkt" % (m["name"], suffix)))
			dims.append((levels[0].shape[1], levels[0].shape[0], len(levels))); offsets.append(at)
			for l in levels: chunks.append(l.reshape(-1)); at += l.size
	return np.array(dims, dtype=np.uint32), np.array(offsets, dtype=np.uint64), np.concatenate(chunks).astype(np.float32)


def light_texture_set(lights, light_count=None):
	"""The light textures as create_and_assign_light_textures() orders them (src/main.c:371-418): unique paths in the order of first use, lights
	without a path share a white texture. Returns (texture index per light, (dims uint32 [T,3], offsets uint64 [T], data float32)), decoded in numpy."""
	paths, indices = [], []
	for light in lights[:light_count]:
		path = light.get("texture_file_path", "")
		if path not in paths: paths.append(path)
		indices.append(paths.index(path))
	if not paths: paths = [""]
	dims, offsets, chunks, at = [], [], [], 0
	for path in paths:
		levels = read_vkt(path) if path else [np.ones((1, 1, 4), dtype=np.float32)]
		dims.append((levels[0].shape[1], levels[0].shape[0], len(levels))); offsets.append(at)
		for l in levels: chunks.append(l.reshape(-1)); at += l.size
	return indices, (np.array(dims, dtype=np.uint32), np.array(offsets, dtype=np.uint64), np.concatenate(chunks).astype(np.float32))


def wang_hash(seed):
	"""src/math_utilities.h:50-57 on uint32 arrays"""
	seed = np.asarray(seed, dtype=np.uint32)
	seed = (seed ^ np.uint32(61)) ^ (seed >> np.uint32(16))
	seed = seed * np.uint32(9)
	seed = seed ^ (seed >> np.uint32(4))
	seed = seed * np.uint32(0x27D4EB2D)
	seed = seed ^ (seed >> np.uint32(15))
	return seed


def white_noise_table(width=256, height=256, layers=64):
	"""src/noise_table.c:73-75: RGBA16 [layer][y][x][4]"""
	n = width * height * layers * 4
	with np.errstate(over="ignore"):
		data = (wang_hash(np.arange(n, dtype=np.uint32) + np.uint32(243708)) & np.uint32(0xFFFF)).astype(np.uint16)
	return data.reshape(layers, height, width, 4)


def quantize_ltc(directory, fresnel_count=51):
	"""src/ltc_table.c:46-116 -> (table0 [F,res,res,4], table1 [F,res,res,2]) uint16"""
	t0, t1 = [], []
	for i in range(fresnel_count):
		with open(os.path.join(directory, "fit%d.dat" % i), "rb") as f:
			(res,) = struct.unpack("<Q", f.read(8))
			m = np.frombuffer(f.read(20 * res * res), dtype="<f4").reshape(-1, 5).astype(np.float32)
		m00, m02, m11, m20, albedo = (m[:, k] for k in range(5))
		inv = np.stack([m11, -m02 * m11, m00 - m02 * m20, -m11 * m20, m00 * m11], axis=1).astype(np.float32)
		# the reference scans all nine entries of the 3x3 inverse (zeros included) for the maximum magnitude
		mx = np.max(np.abs(inv), axis=1, keepdims=True)
		inv = (inv / mx).astype(np.float32)
		e = np.stack([inv[:, 0], -inv[:, 1], inv[:, 2], inv[:, 3], inv[:, 4], albedo], axis=1).astype(np.float32)
		e = np.clip(e, np.float32(0.0), np.float32(1.0))
		q = (e * np.float32(65535.0) + np.float32(0.5)).astype(np.float32).astype(np.uint16)
		t0.append(q[:, :4].reshape(res, res, 4)); t1.append(q[:, 4:].reshape(res, res, 2))
	return np.stack(t0), np.stack(t1)


# ---------------------------------------------------------------------------------------------
# oracle-side frame
# ---------------------------------------------------------------------------------------------

class OracleInputs:
	"""Everything the oracle needs for one dataset, read independently of the CUDA library's loaders."""

	def __init__(self, info, noise_shape=(256, 256, 64)):
		self.info = info
		self.vks = read_vks(info["vks"])
		self.noise = white_noise_table(*noise_shape)
		self.ltc0, self.ltc1 = quantize_ltc(info["ltc"])
		self.material_params = info["material_params"]
		self.textures = material_texture_set(info) if info.get("textured") else None   # mip chains that need filtering, else constant materials
		self._shadow_tris = None
		# textures of the lights (area / portal / IES, shading_pass.frag.glsl:151-185), None when no light is textured
		self.light_textures = light_texture_set(info["lights"])[1] if any(l.get("texturing_technique", 0) for l in info["lights"]) else None

	@property
	def shadow_tris(self):
		if self._shadow_tris is None:
			self._shadow_tris = oracle.dequantize_for_bvh(self.vks["positions"], self.vks["factor"], self.vks["summand"])
		return self._shadow_tris

	def visibility(self, width, height, constants):
		return oracle.visibility(width, height, constants, self.vks["positions"])

	def gbuffer(self, width, height, constants, vis):
		if self.textures is not None:
			return oracle.gbuffer_textured(width, height, constants, vis, self.vks["positions"], self.vks["normals_uv"], self.vks["material_indices"], self.textures)
		return oracle.gbuffer(width, height, constants, vis, self.vks["positions"], self.vks["normals_uv"], self.vks["material_indices"], self.material_params)

	def shade(self, frame_cfg, constants, gbuffer, row_begin=0, row_end=0):
		cfg = dict(frame_cfg); cfg["row_begin"] = row_begin; cfg["row_end"] = row_end
		tris = self.shadow_tris if cfg["trace_shadow_rays"] else np.zeros((0, 9), dtype=np.float32)
		return oracle.shade(cfg, constants, gbuffer, self.noise, self.ltc0, self.ltc1, tris, light_textures=self.light_textures)


def reference_constants(info, width, height, lights, sample_count=1, exposure=1.0, frame_bits=0):
	"""The constant block of a frame from the REFERENCE's own host code (oracle/_ref/libref_host.so: its unchanged loaders and host maths over the shim plus
	the restated bodies of quick_load / write_constants, oracle/ref_host_probe.c). Needs nothing of libvkr_b200.so; default render settings except the
	exposure, noise not animated (what bench.py and the fixtures use). Untextured lights only."""
	import ctypes as C
	lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_host.so"))
	lib.ref_probe_write_constants.restype = C.c_size_t
	lib.ref_probe_write_constants.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
	tri = C.c_uint64(); mat = C.c_uint64(); fs = (C.c_float * 6)(); pos = C.c_void_p(); nuv = C.c_void_p(); mi = C.c_void_p(); soup = C.POINTER(C.c_float)(); ntri = C.c_uint64()
	if lib.ref_probe_load_scene(info["vks"].encode(), info["textures"].encode(), C.byref(tri), C.byref(mat), fs, C.byref(pos), C.byref(nuv), C.byref(mi), C.byref(soup), C.byref(ntri)) != 0:
		raise RuntimeError("the reference's load_scene failed")
	res = C.c_uint32(); t0 = C.c_void_p(); t1 = C.c_void_p(); ltc_constants = (C.c_float * 8)()
	if lib.ref_probe_load_ltc(info["ltc"].encode(), 51, C.byref(res), C.byref(t0), C.byref(t1), ltc_constants) != 0:
		raise RuntimeError("the reference's load_ltc_table failed")
	data = C.c_void_p(); masks = (C.c_uint32 * 7)()
	if lib.ref_probe_load_noise(256, 256, 64, 0, C.byref(data), masks, 0) != 0:   # noise_type_white
		raise RuntimeError("the reference's load_noise_table failed")
	settings = (C.c_float * 6)(0.5, -7.0, exposure, 1.0, 0.0, float(frame_bits))   # specify_default_render_settings, src/main.c:232-249
	buf = (C.c_uint8 * (256 + 320 * 64 * 4))()
	size = lib.ref_probe_write_constants(buf, len(buf), info["save"].encode(), lights, width, height, settings)
	lib.ref_probe_destroy_noise(); lib.ref_probe_destroy_ltc(); lib.ref_probe_destroy_scene()
	if size == 0:
		raise RuntimeError("the reference's constants could not be written")
	return bytes(buf[:size])


def oracle_config(frame, width, height):
	"""The -D defines of the reference (src/main.c:752-792) as the oracle's config, from a vulkan_renderer_b200.Frame."""
	s = frame.settings
	counts = frame.light_vertex_counts() or [3]
	return dict(width=width, height=height, light_count=frame.light_count, max_light_vertex_count=max(max(counts), 3), min_light_vertex_count=min(counts),
		sample_count=s.sample_count, sampling_strategies=s.sampling_strategies, mis_heuristic=s.mis_heuristic,
		biased_sampling=int(s.polygon_sampling_technique == api.TECHNIQUE_PSA_BIASED), trace_shadow_rays=s.trace_shadow_rays, show_polygonal_lights=s.show_polygonal_lights,
		row_begin=0, row_end=0, output_srgb=getattr(frame, "output_srgb", 0), polygon_sampling_technique=min(int(s.polygon_sampling_technique), api.TECHNIQUE_PSA), error_display=getattr(frame, "error_display", 0))


def open_frame(info, cuda_device=0, **kw):
	from vulkan_renderer_b200 import Frame
	return Frame(info["vks"], info["textures"], info["save"], info["ltc"], cuda_device=cuda_device, **kw)


# ---------------------------------------------------------------------------------------------
# comparison
# ---------------------------------------------------------------------------------------------

def compare_radiance(test, ref, rel=1.0e-5, floor=1.0e-6):
	"""Per-pixel relative error |a-b| / max(|b|, floor) over RGB. Returns dict with the worst pixel and counts."""
	a = np.asarray(test, dtype=np.float64)[..., :3]; b = np.asarray(ref, dtype=np.float64)[..., :3]
	nan_mismatch = np.isnan(a) != np.isnan(b)
	err = np.abs(a - b) / np.maximum(np.abs(b), floor)
	err = np.where(np.isnan(err), 0.0, err)
	per_pixel = err.max(axis=-1)
	bad = per_pixel > rel
	return dict(max_rel=float(per_pixel.max()) if per_pixel.size else 0.0, bad_pixels=int(bad.sum()), pixels=int(per_pixel.size),
		bit_exact=bool(np.array_equal(np.asarray(test, dtype=np.float32).view(np.uint32), np.asarray(ref, dtype=np.float32).view(np.uint32))),
		nan_mismatch=int(nan_mismatch.sum()))
