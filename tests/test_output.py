"""The output side of the shading pass: 8-bit quantisation, *.png / *.hdr writers, HDR from the two half-bit
frames, the frame timer. Host code only -- no GPU needed. The files are read back with independent decoders written here
from the format specifications (zlib + PNG chunks, Radiance RGBE run-length scanlines)."""
import ctypes as C
import struct
import zlib

import numpy as np
import pytest

from vulkan_renderer_b200 import api


def _lib():
	lib = api.load_library()
	lib.vkr_get_frame_time.restype = C.c_float
	lib.vkr_record_frame_time.argtypes = [C.c_double]
	return lib


def _read_png(path):
	raw = open(path, "rb").read()
	assert raw[:8] == b"\x89PNG\r\n\x1a\n"
	pos = 8; chunks = []
	while pos < len(raw):
		n, = struct.unpack(">I", raw[pos:pos + 4]); kind = raw[pos + 4:pos + 8]; data = raw[pos + 8:pos + 8 + n]
		crc, = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
		assert zlib.crc32(kind + data) == crc, "bad chunk CRC"
		chunks.append((kind, data)); pos += 12 + n
	assert [k for k, _ in chunks][0] == b"IHDR" and chunks[-1][0] == b"IEND"
	w, h, depth, colour, comp, flt, lace = struct.unpack(">IIBBBBB", chunks[0][1])
	assert (depth, colour, comp, flt, lace) == (8, 2, 0, 0, 0)
	scan = zlib.decompress(b"".join(d for k, d in chunks if k == b"IDAT"))
	rows = np.frombuffer(scan, dtype=np.uint8).reshape(h, 1 + 3 * w)
	assert (rows[:, 0] == 0).all()
	return rows[:, 1:].reshape(h, w, 3)


def _read_hdr(path):
	raw = open(path, "rb").read()
	head, _, body = raw.partition(b"\n\n")
	assert head.startswith(b"#?RADIANCE") and b"FORMAT=32-bit_rle_rgbe" in head
	line, _, body = body.partition(b"\n")
	tok = line.split(); assert tok[0] == b"-Y" and tok[2] == b"+X"
	h, w = int(tok[1]), int(tok[3])
	out = np.zeros((h, w, 4), dtype=np.uint8); pos = 0
	for y in range(h):
		if 8 <= w < 32768:
			assert body[pos] == 2 and body[pos + 1] == 2 and (body[pos + 2] << 8 | body[pos + 3]) == w
			pos += 4
			for c in range(4):
				x = 0
				while x < w:
					n = body[pos]; pos += 1
					if n > 128:
						out[y, x:x + n - 128, c] = body[pos]; pos += 1; x += n - 128
					else:
						out[y, x:x + n, c] = np.frombuffer(body[pos:pos + n], dtype=np.uint8); pos += n; x += n
		else:
			out[y] = np.frombuffer(body[pos:pos + 4 * w], dtype=np.uint8).reshape(w, 4); pos += 4 * w
	assert pos == len(body)
	scale = np.where(out[..., 3:] == 0, 0.0, np.ldexp(1.0, out[..., 3:].astype(np.int32) - 136))
	return out[..., :3].astype(np.float64) * scale


@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (64, 48), (300, 200)])
def test_png_writer_round_trip(tmp_path, shape):
	lib = _lib(); w, h = shape
	rgb = np.random.default_rng(w * h).integers(0, 256, (h, w, 3), dtype=np.uint8)
	path = str(tmp_path / "frame.png").encode()
	assert lib.vkr_write_png(path, w, h, rgb.ctypes.data_as(C.c_void_p)) == 0
	assert np.array_equal(_read_png(path.decode()), rgb)


def test_png_writer_spans_several_stored_blocks(tmp_path):
	lib = _lib(); w, h = 640, 120    # 230 KB of scanlines: four stored deflate blocks
	rgb = np.random.default_rng(1).integers(0, 256, (h, w, 3), dtype=np.uint8)
	path = str(tmp_path / "big.png").encode()
	assert lib.vkr_write_png(path, w, h, rgb.ctypes.data_as(C.c_void_p)) == 0
	assert np.array_equal(_read_png(path.decode()), rgb)


@pytest.mark.parametrize("shape", [(5, 4), (8, 3), (200, 17), (300, 5)])
def test_hdr_writer_round_trip(tmp_path, shape):
	lib = _lib(); w, h = shape
	rng = np.random.default_rng(w)
	rgb = (rng.random((h, w, 3)) * np.exp(rng.uniform(-12, 12, (h, w, 1)))).astype(np.float32)
	rgb[0, 0] = 0.0; rgb[h - 1, w - 1] = (1.0, 0.5, 0.25)
	path = str(tmp_path / "frame.hdr").encode()
	assert lib.vkr_write_hdr(path, w, h, rgb.ctypes.data_as(C.c_void_p)) == 0
	back = _read_hdr(path.decode())
	peak = rgb.max(axis=-1, keepdims=True).astype(np.float64)
	assert (np.abs(back - rgb) <= peak / 128.0 + 1e-38).all()     # 8-bit mantissa relative to the largest channel, truncated
	assert np.array_equal(back[h - 1, w - 1], [1.0, 0.5, 0.25]) and (back[0, 0] == 0).all()


def test_writers_report_failure_like_the_reference(tmp_path, capfd):
	lib = _lib()
	rgb = np.zeros((2, 2, 3), dtype=np.uint8)
	assert lib.vkr_write_png(b"/nonexistent_directory/x.png", 2, 2, rgb.ctypes.data_as(C.c_void_p)) == 1
	assert lib.vkr_write_hdr(b"/nonexistent_directory/x.hdr", 2, 2, np.zeros((2, 2, 3), dtype=np.float32).ctypes.data_as(C.c_void_p)) == 1
	assert lib.vkr_write_png(str(tmp_path / "e.png").encode(), 0, 2, rgb.ctypes.data_as(C.c_void_p)) == 1
	out = capfd.readouterr().out
	assert "Please check path and permissions" in out


def test_unorm8_quantisation():
	lib = _lib()
	x = np.array([[-1.0, 0.0, 0.5 / 255, 1.0], [0.499 / 255, 0.501 / 255, 1.0, 1.0], [254.5 / 255 + 1e-6, 2.0, np.nan, 1.0], [np.inf, 0.2, 0.7, 0.0]], dtype=np.float32).reshape(2, 2, 4)
	out = np.zeros((2, 2, 3), dtype=np.uint8)
	lib.vkr_quantize_unorm8(x.ctypes.data_as(C.c_void_p), 2, 2, out.ctypes.data_as(C.c_void_p))
	assert out.reshape(4, 3).tolist() == [[0, 0, 1], [0, 1, 255], [255, 255, 0], [255, 51, 179]]
	ramp = (np.arange(256, dtype=np.float32) / np.float32(255.0)).reshape(16, 16)
	frame = np.stack([ramp, ramp, ramp, np.ones_like(ramp)], axis=-1).copy()
	q = np.zeros((16, 16, 3), dtype=np.uint8)
	lib.vkr_quantize_unorm8(frame.ctypes.data_as(C.c_void_p), 16, 16, q.ctypes.data_as(C.c_void_p))
	assert np.array_equal(q[..., 0].reshape(-1), np.arange(256))


def test_hdr_screenshot_combination_inverts_the_half_bit_split():
	"""The shader splits packHalf2x16(colour) into a low-byte and a high-byte frame (shading_pass.frag.glsl:871-887); the host puts
	them together again (src/main.c:1696-1707). All 65536 half patterns survive, Inf and NaN included."""
	lib = _lib()
	halves = np.arange(65536, dtype=np.uint16)
	low = (halves & 0xFF).astype(np.uint8); high = (halves >> 8).astype(np.uint8)
	out = np.zeros(65536, dtype=np.float32)
	lib.vkr_combine_ldr_screenshots_into_hdr(low.ctypes.data_as(C.c_void_p), high.ctypes.data_as(C.c_void_p), C.c_size_t(65536), out.ctypes.data_as(C.c_void_p))
	ref = halves.view(np.float16).astype(np.float32)
	finite = np.isfinite(ref)
	assert np.array_equal(out[finite].view(np.uint32), ref[finite].view(np.uint32))
	assert np.array_equal(np.isinf(out), np.isinf(ref)) and np.array_equal(np.isnan(out), np.isnan(ref))


def test_frame_timer_is_the_median_of_recent_frame_times():
	lib = _lib()
	lib.vkr_reset_frame_times()
	assert lib.vkr_get_frame_time() == 0.0
	lib.vkr_record_frame_time(10.0)
	assert lib.vkr_get_frame_time() == 0.0                    # one time stamp, no difference yet
	t = 10.0
	for dt in [0.010, 0.020, 0.030, 0.040, 1.000]:            # an outlier does not move the median
		t += dt; lib.vkr_record_frame_time(t)
	assert abs(lib.vkr_get_frame_time() - 0.030) < 1e-6
	for _ in range(150):                                      # the window holds the last 100 stamps
		t += 0.005; lib.vkr_record_frame_time(t)
	assert abs(lib.vkr_get_frame_time() - 0.005) < 1e-6
	lib.vkr_reset_frame_times()
	assert lib.vkr_get_frame_time() == 0.0


def test_render_targets_need_a_gpu_and_say_so(capfd):
	import torch
	if torch.cuda.is_available():
		pytest.skip("covered by the -m gpu tests on a GPU box")
	lib = _lib()
	targets = api.RenderTargets(); dev = api.Device()
	assert lib.vkr_create_render_targets(C.byref(targets), C.byref(dev), 64, 32) == 1
	assert targets.width == 0 and not targets.d_frame          # zeroed, like the reference's create_* on failure
	assert lib.vkr_create_render_targets(C.byref(targets), C.byref(dev), 0, 32) == 1
	assert "Failed to create render targets" in capfd.readouterr().out
	lib.vkr_destroy_render_targets(C.byref(targets), C.byref(dev))   # tolerates a zeroed object
