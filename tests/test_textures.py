"""Material textures: the *.vkt mip-chain loader with its format decoders against independent numpy decoders, and the
texture filter definition (oracle/texture_filter.h) against what it promises -- exact texel values at texel centres, the box-filtered
average under heavy minification, repeat addressing, anisotropic taps along the major axis. Host side only."""
import ctypes as C
import struct

import numpy as np
import pytest

from oracle import binding as O
from vulkan_renderer_b200 import api, synth


def _load(path):
	lib = api.load_library()
	t = api.Texture()
	assert lib.vkr_load_texture(C.byref(t), str(path).encode()) == 0
	levels = []; at = 0
	for l in range(t.mip_count):
		w, h = max(t.width >> l, 1), max(t.height >> l, 1)
		levels.append(np.ctypeslib.as_array(t.h_texels, (t.texel_float_count,))[at:at + 4 * w * h].reshape(h, w, 4).copy()); at += 4 * w * h
	assert at == t.texel_float_count
	info = dict(width=t.width, height=t.height, mips=t.mip_count, format=t.vk_format, constant=bool(t.is_constant))
	lib.vkr_destroy_texture(C.byref(t))
	assert not t.h_texels
	return info, levels


def _picture(h, w, seed):
	rng = np.random.default_rng(seed)
	yy, xx = np.mgrid[0:h, 0:w]
	return np.stack([0.1 + 0.8 * ((xx // 3 + yy // 5) % 2), rng.random((h, w)), 0.5 + 0.5 * np.sin(0.3 * xx + 0.2 * yy), np.ones((h, w))], -1).astype(np.float32)


def _decode_bc1(data, w, h):
	out = np.zeros((h, w, 3)); bw, bh = (w + 3) // 4, (h + 3) // 4
	for by in range(bh):
		for bx in range(bw):
			c0, c1, bits = struct.unpack_from("<HHI", data, 8 * (by * bw + bx))
			un = lambda c: np.array([((c >> 11) & 31) / 31.0, ((c >> 5) & 63) / 63.0, (c & 31) / 31.0])
			a, b = un(c0), un(c1)
			pal = [a, b, (2 * a + b) / 3, (a + 2 * b) / 3] if c0 > c1 else [a, b, (a + b) / 2, np.zeros(3)]
			for t in range(16):
				x, y = 4 * bx + (t & 3), 4 * by + (t >> 2)
				if x < w and y < h: out[y, x] = pal[(bits >> (2 * t)) & 3]
	return out


def _decode_bc4(data, offset):
	r0, r1 = data[offset], data[offset + 1]
	pal = [r0, r1] + ([((8 - i) * r0 + (i - 1) * r1) / 7.0 for i in range(2, 8)] if r0 > r1 else [((6 - i) * r0 + (i - 1) * r1) / 5.0 for i in range(2, 6)] + [0.0, 255.0])
	bits = int.from_bytes(data[offset + 2:offset + 8], "little")
	return [pal[(bits >> (3 * t)) & 7] / 255.0 for t in range(16)]


@pytest.mark.parametrize("vk_format", [97, 109, 131, 141])
@pytest.mark.parametrize("shape", [(16, 16), (20, 12), (5, 7)])
def test_loader_decodes_every_mip_level(tmp_path, vk_format, shape):
	h, w = shape
	picture = _picture(h, w, vk_format + h)
	path = tmp_path / "t.vkt"
	synth.write_vkt(str(path), picture, vk_format)
	info, levels = _load(path)
	expected = synth.mip_chain(picture)
	assert (info["width"], info["height"], info["mips"], info["format"]) == (w, h, len(expected), vk_format) and not info["constant"]
	raw = open(path, "rb").read()
	headers = [struct.unpack_from("<IIQQ", raw, 32 + 24 * k) for k in range(len(expected))]
	payload_at = 32 + 24 * len(expected)
	for level, want, (lw, lh, size, offset) in zip(levels, expected, headers):
		assert level.shape == (lh, lw, 4)
		data = raw[payload_at + offset: payload_at + offset + size]
		if vk_format == 97:
			assert np.array_equal(level[..., :3], want[..., :3].astype(np.float16).astype(np.float32)) and (level[..., 3] == 1).all()
		elif vk_format == 109:
			assert np.array_equal(level[..., :3], want[..., :3])
		elif vk_format == 131:
			assert np.allclose(level[..., :3], _decode_bc1(data, lw, lh), atol=1e-6) and (level[..., 3] == 1).all()
			assert np.abs(level[..., :3] - want[..., :3]).mean() < 0.25          # (the test picture has a noise channel BC1 cannot hold; this only guards against a broken encoder)
		else:
			bw = (lw + 3) // 4
			for y in range(lh):
				for x in range(lw):
					blk = 16 * ((y // 4) * bw + x // 4); t = (y % 4) * 4 + x % 4
					assert abs(level[y, x, 0] - _decode_bc4(data, blk)[t]) < 1e-6 and abs(level[y, x, 1] - _decode_bc4(data, blk + 8)[t]) < 1e-6
			assert (level[..., 2] == 0).all() and (level[..., 3] == 1).all() and np.abs(level[..., :2] - want[..., :2]).max() < 0.3


def test_constant_textures_are_recognised_and_bad_files_rejected(tmp_path, capfd):
	synth.write_vkt_constant(str(tmp_path / "c.vkt"), [0.25, 0.5, 0.75, 1.0])
	info, levels = _load(tmp_path / "c.vkt")
	assert info["constant"] and info["mips"] == 3 and np.array_equal(levels[0][0, 0], [0.25, 0.5, 0.75, 1.0])
	lib = api.load_library(); t = api.Texture()
	good = open(tmp_path / "c.vkt", "rb").read()
	for name, blob in (("marker", b"\\0\\0\\0\\0" + good[4:]), ("short", good[:-8]), ("eof", good[:-4] + b"\\0\\0\\0\\0"), ("format", good[:20] + struct.pack("<I", 999) + good[24:])):
		(tmp_path / name).write_bytes(blob)
		assert lib.vkr_load_texture(C.byref(t), str(tmp_path / name).encode()) == 1 and not t.h_texels and t.width == 0
	assert lib.vkr_load_texture(C.byref(t), b"/nonexistent.vkt") == 1
	assert "Failed to open" in capfd.readouterr().out


def _flat(levels):
	return np.concatenate([l.reshape(-1) for l in levels]).astype(np.float32)


def test_filter_definition_behaves_like_a_sampler():
	picture = _picture(32, 16, 3)
	levels = synth.mip_chain(picture)
	texels = _flat(levels)
	h, w = 32, 16
	grad = lambda rows: O.texture_grad_batch(w, h, len(levels), texels, rows)
	# (1) magnified, at texel centres: the texel itself; repeat addressing one period away
	xs, ys = np.meshgrid(np.arange(w), np.arange(h))
	centres = np.stack([(xs + 0.5) / w, (ys + 0.5) / h], -1).reshape(-1, 2)
	tiny = np.tile(np.array([1e-4 / w, 0, 0, 1e-4 / h], dtype=np.float32), (len(centres), 1))
	assert np.allclose(grad(np.concatenate([centres, tiny], 1)), picture.reshape(-1, 4), atol=1e-6)
	assert np.allclose(grad(np.concatenate([centres + [3.0, -2.0], tiny], 1)), picture.reshape(-1, 4), atol=1e-5)
	# (2) halfway between two texels: their mean
	mid = grad([[(4 + 1.0) / w, (7 + 0.5) / h, 1e-6, 0, 0, 1e-6]])[0]
	assert np.allclose(mid, 0.5 * (picture[7, 4] + picture[7, 5]), atol=1e-6)
	# (3) footprint of the whole texture: the last mip level = the mean of the picture
	assert np.allclose(grad([[0.3, 0.6, 1.0, 0, 0, 1.0]])[0], picture.reshape(-1, 4).mean(0), atol=1e-5)
	# (4) isotropic footprint of 2 texels -> level 1, exactly between level-1 texels' support; 4 texels -> level 2
	l1 = grad([[(2 + 1.0) / w, (4 + 1.0) / h, 2.0 / w, 0, 0, 2.0 / h]])[0]
	assert np.allclose(l1, levels[1][2, 1], atol=1e-6)
	# (5) anisotropic: 8 x 1 texels footprint along x = the mean of 8 neighbouring texels of level 0 in a row (8 taps, no mip blur along y)
	y, x0 = 9, 4
	aniso = grad([[(x0 + 4.0) / w, (y + 0.5) / h, 8.0 / w, 0, 0, 1.0 / h]])[0]
	assert np.allclose(aniso, picture[y, x0:x0 + 8].mean(0), atol=1e-5)
	iso = grad([[(x0 + 4.0) / w, (y + 0.5) / h, 8.0 / w, 0, 0, 8.0 / h]])[0]
	assert not np.allclose(iso, aniso, atol=1e-3)            # the isotropic footprint of the same width blurs in y as well
	# (6) degenerate inputs stay finite
	weird = grad([[np.nan, 0.5, 0, 0, 0, 0], [0.5, 0.5, np.inf, 0, 0, 0], [1e30, -1e30, 1e-3, 0, 0, 1e-3], [0.5, 0.5, 0, 0, 0, 0]])
	assert np.isfinite(weird).all()


# ---- textures of the polygonal lights (create_and_assign_light_textures, src/main.c:371-418; sampler src/main.c:613-623)

def test_light_textures_are_loaded_once_per_path_and_indexed_like_the_reference(tmp_path, capfd):
	from tests import harness as H
	lib = api.load_library()
	info = H.dataset("mini_lit")
	spec = api.SceneSpecification()
	assert lib.vkr_quick_load(C.byref(spec), info["save"].encode()) == 0
	assert [spec.polygonal_lights[i].texturing_technique for i in range(3)] == [1, 2, 3]
	lt = api.LightTextures()
	assert lib.vkr_create_and_assign_light_textures(C.byref(lt), None, C.byref(spec)) == 0
	indices, (dims, offsets, data) = H.light_texture_set(info["lights"])
	assert [spec.polygonal_lights[i].texture_index for i in range(3)] == indices == [0, 1, 2]
	assert lt.texture_count == 3 and not lt.d_texels
	for k in range(3):   # host copies == the numpy decoders, level by level
		t = lt.textures[k]
		assert (t.width, t.height, t.mip_count) == tuple(dims[k])
		mine = np.ctypeslib.as_array(t.h_texels, (t.texel_float_count,))
		assert np.array_equal(mine, data[offsets[k]:offsets[k] + t.texel_float_count])
	lib.vkr_destroy_light_textures(C.byref(lt), None)
	assert lt.texture_count == 0 and not lt.textures
	# lights 0 and 2 share a path, light 1 points at a missing file -> white, with the reference's message; indices only when no struct is passed
	shared = spec.polygonal_lights[0].texture_file_path
	missing = str(tmp_path / "nowhere.vkt").encode()
	saved = [spec.polygonal_lights[i].texture_file_path for i in range(3)]
	keep = C.create_string_buffer(missing)
	spec.polygonal_lights[2].texture_file_path = shared; spec.polygonal_lights[1].texture_file_path = C.cast(keep, C.c_void_p)
	assert lib.vkr_create_and_assign_light_textures(None, None, C.byref(spec)) == 0
	assert [spec.polygonal_lights[i].texture_index for i in range(3)] == [0, 1, 0]
	assert lib.vkr_create_and_assign_light_textures(C.byref(lt), None, C.byref(spec)) == 0
	assert "does not exist. Using a white texture instead." in capfd.readouterr().out
	assert lt.texture_count == 2 and (lt.textures[1].width, lt.textures[1].height, lt.textures[1].mip_count) == (1, 1, 1)
	assert list(np.ctypeslib.as_array(lt.textures[1].h_texels, (4,))) == [1.0, 1.0, 1.0, 1.0]
	lib.vkr_destroy_light_textures(C.byref(lt), None)
	for i in range(3): spec.polygonal_lights[i].texture_file_path = saved[i]
	lib.vkr_destroy_scene_specification(C.byref(spec))


def test_light_texture_sampler_repeats_in_u_and_clamps_in_v():
	"""oracle/texture_filter.h: vkr_texture_bilinear_repeat_clamp (what the compiled reference shader, the oracle and the kernel use for textureLod on
	light textures): texel centres are exact, u wraps around, v stops at the border rows."""
	from tests import harness as H
	rng = np.random.default_rng(11)
	tex = rng.random((8, 16, 4)).astype(np.float32)
	info = H.dataset("mini_lit"); oi = H.OracleInputs(info)
	def sample(u, v):
		uv = np.array([[u, v]], dtype=np.float32)
		return O.light_texture_batch(tex, uv)[0]
	assert np.array_equal(sample((3 + 0.5) / 16, (5 + 0.5) / 8), tex[5, 3])
	assert np.array_equal(sample((3 + 0.5) / 16 + 2.0, (5 + 0.5) / 8), sample((3 + 0.5) / 16, (5 + 0.5) / 8))   # repeat in u
	assert np.array_equal(sample((3 + 0.5) / 16, -0.7), tex[0, 3]) and np.array_equal(sample((3 + 0.5) / 16, 1.9), tex[7, 3])   # clamp in v
	edge = sample(0.0, (2 + 0.5) / 8)   # between the last and the first column
	assert np.allclose(edge, 0.5 * (tex[2, 15] + tex[2, 0]), atol=1e-6)
