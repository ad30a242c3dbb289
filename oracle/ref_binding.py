"""ctypes binding of oracle/_ref/libref_shader.so: the REFERENCE's shader sources compiled as C++ (TEST INFRASTRUCTURE).

Built by oracle/build_ref.py where /root/reference exists; travels prebuilt to the GPU box (never required there)."""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# VKR_REF_SET=<name>: a second set of configurations compiled by `python oracle/build_ref.py --random <count> <seed> <name>` (tools/fuzz_parity.py --ref-set)
_SET = os.environ.get("VKR_REF_SET", "")
LIB_PATH = os.path.join(_HERE, "_ref", "libref_shader%s.so" % (("_" + _SET) if _SET else ""))
CONFIGS_PATH = os.path.join(_HERE, "_ref", "configs%s.json" % (("_" + _SET) if _SET else ""))

HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float)


class RefArgs(C.Structure):
	_fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("light_count", C.c_uint32), ("sample_count", C.c_uint32), ("max_light_vertex_count", C.c_uint32), ("material_count", C.c_uint32),
		("constants", C.c_void_p), ("visibility", C.c_void_p),
		("quantized_positions", C.c_void_p), ("normals_and_tex_coords", C.c_void_p), ("material_indices", C.c_void_p), ("material_params", C.c_void_p),
		("noise", C.c_void_p), ("noise_w", C.c_uint32), ("noise_h", C.c_uint32), ("noise_layers", C.c_uint32),
		("ltc0", C.c_void_p), ("ltc1", C.c_void_p), ("ltc_res", C.c_uint32), ("ltc_layers", C.c_uint32),
		("occluded_hook", C.c_void_p), ("occluded_user", C.c_void_p), ("out_rgba", C.c_void_p),
		("row_begin", C.c_uint32), ("row_end", C.c_uint32), ("band_height", C.c_uint32), ("band_stride", C.c_uint32), ("shade_seconds", C.c_double),
		("texture_dims", C.c_void_p), ("texture_offsets", C.c_void_p), ("texture_data", C.c_void_p),
		("light_texture_count", C.c_uint32), ("light_texture_dims", C.c_void_p), ("light_texture_offsets", C.c_void_p), ("light_texture_data", C.c_void_p)]


def available():
	return os.path.exists(LIB_PATH) and os.path.exists(CONFIGS_PATH)


def configs():
	with open(CONFIGS_PATH) as f:
		return json.load(f)


_lib = None


def load():
	global _lib
	if _lib is None:
		_lib = C.CDLL(LIB_PATH)
		_lib.ref_bvh_create.restype = C.c_void_p
		_lib.ref_bvh_create.argtypes = [C.c_void_p, C.c_uint32]
		_lib.ref_bvh_destroy.argtypes = [C.c_void_p]
	return _lib


_last_shade_seconds = 0.0


def last_shade_seconds():
	"""Wall clock of the pixel loop of the last shade() call (BVH build and set-up excluded)."""
	return _last_shade_seconds


def set_threads(count):
	"""OpenMP threads of the next shade() call. The launcher's OMP_NUM_THREADS only sets the start value (torchrun exports 1)."""
	C.CDLL("libgomp.so.1").omp_set_num_threads(int(count))


def thread_count():
	return int(load().ref_thread_count())


def find_config(**wanted):
	"""The built configuration whose -D defines equal `wanted` (keys of oracle/build_ref.py), or None."""
	if not available():
		return None
	defaults = dict(technique=11, srgb=0, frame_bits=0, error_display=0)
	for c in configs():
		full = dict(defaults, min_vertices=c["max_vertices"]); full.update(c)
		if all(full.get(k) == v for k, v in wanted.items()):
			return c
	return None


def shade(entry, width, height, cfg, constants, visibility, vks, material_params, noise, ltc0, ltc1, shadow_tris, row_begin=0, row_end=0, band_height=0, band_stride=0, textures=None, light_textures=None):
	"""Runs the reference fragment shader (configuration `entry`) for every pixel (or the rows / bands asked for). Returns float32 [H, W, 4]."""
	global _last_shade_seconds
	lib = load()
	keep = []
	def arr(a, dtype):
		a = np.ascontiguousarray(a, dtype=dtype); keep.append(a); return a.ctypes.data
	out = np.zeros((height, width, 4), dtype=np.float32)
	cb = (C.c_uint8 * len(constants)).from_buffer_copy(constants)
	tris = np.ascontiguousarray(shadow_tris, dtype=np.float32).reshape(-1, 9)
	bvh = lib.ref_bvh_create(tris.ctypes.data, len(tris))
	a = RefArgs(width=width, height=height, light_count=cfg["lights"], sample_count=cfg["samples"], max_light_vertex_count=cfg["max_vertices"], material_count=len(material_params),
		constants=C.addressof(cb), visibility=arr(visibility, np.uint32),
		quantized_positions=arr(vks["positions"], np.uint32), normals_and_tex_coords=arr(vks["normals_uv"], np.uint16), material_indices=arr(vks["material_indices"], np.uint8),
		material_params=arr(material_params, np.float32), noise=arr(noise, np.uint16), noise_w=noise.shape[2], noise_h=noise.shape[1], noise_layers=noise.shape[0],
		ltc0=arr(ltc0, np.uint16), ltc1=arr(ltc1, np.uint16), ltc_res=ltc0.shape[1], ltc_layers=ltc0.shape[0],
		occluded_hook=C.cast(lib.ref_bvh_occluded, C.c_void_p), occluded_user=bvh, out_rgba=out.ctypes.data,
		row_begin=row_begin, row_end=row_end, band_height=band_height, band_stride=band_stride)
	if textures is not None:   # (dims uint32 [T,3], offsets uint64 [T], data float32): mip chains of 3 textures per material
		a.texture_dims = arr(textures[0], np.uint32); a.texture_offsets = arr(textures[1], np.uint64); a.texture_data = arr(textures[2], np.float32)
	if light_textures is not None:   # same triple for the textures of the lights
		a.light_texture_count = len(light_textures[0])
		a.light_texture_dims = arr(light_textures[0], np.uint32); a.light_texture_offsets = arr(light_textures[1], np.uint64); a.light_texture_data = arr(light_textures[2], np.float32)
	fn = getattr(lib, entry)
	rc = fn(C.byref(a))
	lib.ref_bvh_destroy(bvh)
	_last_shade_seconds = float(a.shade_seconds)
	if rc != 0:
		raise RuntimeError("%s rejected the arguments (configuration mismatch)" % entry)
	return out
