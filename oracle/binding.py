"""ctypes binding of the CPU oracle (oracle/build/liboracle.so). TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "build", "liboracle.so")


class Config(C.Structure):
	_fields_ = [(n, C.c_uint32) for n in ("width", "height", "light_count", "max_light_vertex_count", "min_light_vertex_count", "sample_count",
		"sampling_strategies", "mis_heuristic", "biased_sampling", "trace_shadow_rays", "show_polygonal_lights", "row_begin", "row_end", "band_height", "band_stride", "polygon_sampling_technique", "error_display", "output_srgb")]


_lib = None


def load():
	global _lib
	if _lib is None:
		if not os.path.exists(LIB_PATH):
			subprocess.check_call(["make", "-C", _HERE])
		lib = C.CDLL(LIB_PATH)
		lib.vkr_oracle_light_stride.restype = C.c_size_t
		lib.vkr_oracle_clip.restype = C.c_uint32
		_lib = lib
	return _lib


def _p(a):
	return a.ctypes.data_as(C.c_void_p)


def shade(cfg, constants, gbuffer, noise, ltc0, ltc1, tris, light_textures=None):
	"""cfg: dict of Config fields. light_textures: None or (dims uint32 [T,3], offsets uint64 [T] in floats, data float32), the textures the lights'
	texture_index refers to. Returns (rgba float32 [H,W,4], shadow ray count)."""
	lib = load()
	c = Config(**{"polygon_sampling_technique": 11, **cfg})
	gbuffer = np.ascontiguousarray(gbuffer, dtype=np.float32)
	noise = np.ascontiguousarray(noise, dtype=np.uint16); ltc0 = np.ascontiguousarray(ltc0, dtype=np.uint16); ltc1 = np.ascontiguousarray(ltc1, dtype=np.uint16)
	tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
	out = np.zeros((c.height, c.width, 4), dtype=np.float32)
	rays = C.c_uint64(0)
	cb = (C.c_uint8 * len(constants)).from_buffer_copy(constants)
	if light_textures is None:
		count, dims, offsets, data = 0, None, None, None
	else:
		dims = np.ascontiguousarray(light_textures[0], dtype=np.uint32); offsets = np.ascontiguousarray(light_textures[1], dtype=np.uint64); data = np.ascontiguousarray(light_textures[2], dtype=np.float32)
		count = len(dims)
	rc = lib.vkr_oracle_shade_with_light_textures(C.byref(c), cb, _p(gbuffer), _p(noise), C.c_uint32(noise.shape[2]), C.c_uint32(noise.shape[1]), C.c_uint32(noise.shape[0]),
		_p(ltc0), _p(ltc1), C.c_uint32(ltc0.shape[1]), C.c_uint32(ltc0.shape[0]), _p(tris), C.c_uint32(len(tris)),
		C.c_uint32(count), _p(dims) if count else None, _p(offsets) if count else None, _p(data) if count else None, _p(out), C.byref(rays))
	if rc != 0:
		raise RuntimeError("vkr_oracle_shade failed")
	return out, rays.value


def dequantize_for_bvh(quantized_positions, factor, summand):
	lib = load()
	q = np.ascontiguousarray(quantized_positions, dtype=np.uint32).reshape(-1, 2)
	f = np.ascontiguousarray(factor, dtype=np.float32); s = np.ascontiguousarray(summand, dtype=np.float32)
	out = np.zeros((len(q), 3), dtype=np.float32)
	lib.vkr_oracle_dequantize_for_bvh(_p(q), C.c_uint64(len(q)), _p(f), _p(s), _p(out))
	return out.reshape(-1, 9)


def visibility(width, height, constants, quantized_positions):
	lib = load()
	q = np.ascontiguousarray(quantized_positions, dtype=np.uint32).reshape(-1, 2)
	out = np.zeros((height, width), dtype=np.uint32)
	cb = (C.c_uint8 * len(constants)).from_buffer_copy(constants)
	lib.vkr_oracle_visibility(C.c_uint32(width), C.c_uint32(height), cb, _p(q), C.c_uint64(len(q) // 3), _p(out))
	return out


def gbuffer(width, height, constants, vis, quantized_positions, normals_and_tex_coords, material_indices, material_params):
	lib = load()
	q = np.ascontiguousarray(quantized_positions, dtype=np.uint32); nt = np.ascontiguousarray(normals_and_tex_coords, dtype=np.uint16)
	mi = np.ascontiguousarray(material_indices, dtype=np.uint8); mp = np.ascontiguousarray(material_params, dtype=np.float32)
	vis = np.ascontiguousarray(vis, dtype=np.uint32)
	out = np.zeros((4, height, width, 4), dtype=np.float32)
	cb = (C.c_uint8 * len(constants)).from_buffer_copy(constants)
	lib.vkr_oracle_gbuffer(C.c_uint32(width), C.c_uint32(height), cb, _p(vis), _p(q), _p(nt), _p(mi), _p(mp), _p(out))
	return out


def gbuffer_textured(width, height, constants, vis, quantized_positions, normals_and_tex_coords, material_indices, textures):
	"""textures = (dims uint32 [T,3] {width, height, mips}, offsets uint64 [T], data float32): RGBA32F mip chains, 3 per material."""
	lib = load()
	q = np.ascontiguousarray(quantized_positions, dtype=np.uint32); nt = np.ascontiguousarray(normals_and_tex_coords, dtype=np.uint16)
	mi = np.ascontiguousarray(material_indices, dtype=np.uint8); vis = np.ascontiguousarray(vis, dtype=np.uint32)
	dims = np.ascontiguousarray(textures[0], dtype=np.uint32); offsets = np.ascontiguousarray(textures[1], dtype=np.uint64); data = np.ascontiguousarray(textures[2], dtype=np.float32)
	out = np.zeros((4, height, width, 4), dtype=np.float32)
	cb = (C.c_uint8 * len(constants)).from_buffer_copy(constants)
	rc = lib.vkr_oracle_gbuffer_textured(C.c_uint32(width), C.c_uint32(height), cb, _p(vis), _p(q), _p(nt), _p(mi), C.c_uint32(len(dims)), _p(dims), _p(offsets), _p(data), _p(out))
	assert rc == 0
	return out


def texture_grad_batch(width, height, mip_count, texels, inputs):
	lib = load()
	texels = np.ascontiguousarray(texels, dtype=np.float32); inputs = np.ascontiguousarray(inputs, dtype=np.float32).reshape(-1, 6)
	out = np.zeros((len(inputs), 4), dtype=np.float32)
	lib.vkr_oracle_texture_grad_batch(C.c_uint32(width), C.c_uint32(height), C.c_uint32(mip_count), _p(texels), C.c_uint32(len(inputs)), _p(inputs), _p(out))
	return out


def clip(vertex_count, vertices, maxp):
	lib = load()
	v = np.zeros((8, 3), dtype=np.float32); v[:len(vertices)] = vertices
	vc = lib.vkr_oracle_clip(C.c_uint32(vertex_count), _p(v), C.c_uint32(maxp))
	return vc, v


def psa_sample_batch(vertices, maxp, random_numbers, biased=False, do_clip=True):
	"""Returns (dirs [n,3], errors [n,2], info dict)."""
	lib = load()
	vertices = np.asarray(vertices, dtype=np.float32)
	v = np.zeros((9, 3), dtype=np.float32); v[:len(vertices)] = vertices
	r = np.ascontiguousarray(random_numbers, dtype=np.float32).reshape(-1, 2)
	dirs = np.zeros((len(r), 3), dtype=np.float32); err = np.zeros((len(r), 2), dtype=np.float32); info = np.zeros(11, dtype=np.float32)
	lib.vkr_oracle_psa_sample_batch(C.c_uint32(len(vertices)), _p(v), C.c_uint32(maxp), C.c_int(int(biased)), C.c_int(int(do_clip)), C.c_uint32(len(r)), _p(r), _p(dirs), _p(err), _p(info))
	return dirs, err, dict(psa=float(info[0]), central=bool(info[1]), vc=int(info[2]), sectors=info[3:].copy())


def sort_network(vertices_xy, ellipses_xy, maxp):
	lib = load()
	v = np.ascontiguousarray(vertices_xy, dtype=np.float32).copy(); e = np.ascontiguousarray(ellipses_xy, dtype=np.float32).copy()
	lib.vkr_oracle_sort_network(C.c_uint32(len(v)), C.c_uint32(maxp), _p(v), _p(e))
	return v, e


ELEMENTARY = dict(atan=0, sin=1, cos=2, acos01=3, rsqrt=4, fast_positive_atan=5, log2=6, exp2=7, linear_to_srgb=8, srgb_to_linear=9, float_to_half=10, acos=11, atan2_pair=12, cbrt_pow=13)


def elementary(which, x):
	lib = load()
	x = np.ascontiguousarray(x, dtype=np.float32); y = np.zeros_like(x)
	lib.vkr_oracle_elementary_batch(C.c_int(ELEMENTARY[which]), C.c_uint32(x.size), _p(x), _p(y))
	return y


def kahan(a, b, c, d):
	lib = load()
	lib.vkr_oracle_kahan.restype = C.c_float
	lib.vkr_oracle_kahan.argtypes = [C.c_float] * 4
	return lib.vkr_oracle_kahan(a, b, c, d)


def trace_any(tris, rays, brute=True):
	lib = load()
	tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9); rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
	a = np.zeros(len(rays), dtype=np.uint8); b = np.zeros(len(rays), dtype=np.uint8)
	lib.vkr_oracle_trace_any(_p(tris), C.c_uint32(len(tris)), C.c_uint32(len(rays)), _p(rays), _p(a), _p(b) if brute else None)
	return a, (b if brute else None)


def set_threads(count):
	"""OpenMP threads of the next shade() call. The launcher's OMP_NUM_THREADS only sets the start value (torchrun exports 1)."""
	C.CDLL("libgomp.so.1").omp_set_num_threads(int(count))


def thread_count():
	return load().vkr_oracle_thread_count()


def last_shade_seconds():
	"""Wall-clock seconds of the pixel loop of the last shade() call (BVH build excluded)."""
	lib = load()
	lib.vkr_oracle_last_shade_seconds.restype = C.c_double
	return lib.vkr_oracle_last_shade_seconds()


def related_work_batch(technique, maxv, light_block, position, frame, random_numbers, symbol_library=None, symbol="vkr_oracle_related_work_batch"):
	"""One (light, shading point) pair, n samples of a related-work technique. frame = rows x, y, z of world_to_shading_space + translation (12 floats).
	Returns None if the light is culled, else (dirs [n,3], densities [n], ggx density factor). symbol_library/symbol let the same call run
	against tests/build/libdevice_on_host.so (the device code compiled for the CPU)."""
	lib = symbol_library if symbol_library is not None else load()
	rnd = np.ascontiguousarray(random_numbers, dtype=np.float32).reshape(-1, 2)
	n = len(rnd)
	pos = np.ascontiguousarray(position, dtype=np.float32); fr = np.ascontiguousarray(frame, dtype=np.float32).reshape(12)
	dirs = np.zeros((n, 3), dtype=np.float32); dens = np.zeros(n, dtype=np.float32); ggx = C.c_float(0.0)
	block = (C.c_uint8 * len(light_block)).from_buffer_copy(light_block)
	fn = getattr(lib, symbol); fn.restype = C.c_int
	on = fn(C.c_uint32(technique), C.c_uint32(maxv), block, _p(pos), _p(fr), C.c_uint32(n), _p(rnd), _p(dirs), _p(dens), C.byref(ggx))
	if on < 0:
		raise RuntimeError("unsupported technique / vertex bound")
	return None if on == 0 else (dirs, dens, float(ggx.value))


def light_texture_batch(texture, uv):
	"""textureLod(..., 0) with the light-texture sampler (repeat in u, clamp in v) on one RGBA32F level [H, W, 4] for uv float32 [n, 2]."""
	lib = load()
	texture = np.ascontiguousarray(texture, dtype=np.float32); uv = np.ascontiguousarray(uv, dtype=np.float32)
	out = np.zeros((len(uv), 4), dtype=np.float32)
	lib.vkr_oracle_light_texture_batch(C.c_uint32(texture.shape[1]), C.c_uint32(texture.shape[0]), _p(texture), C.c_uint32(len(uv)), _p(uv), _p(out))
	return out
