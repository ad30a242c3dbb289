"""One frame on several GPUs (include/vkr_b200.h, vkr_frame_exchange_t): every pass instance shades its tile columns and stores the pixels into the
frames of all instances from the kernel epilogue; two one-block kernels are the barrier. The exchange is exercised on ONE GPU (two or three instances with
their own streams and frames on the same device, connected with vkr_frame_exchange_connect_local: the same kernels, pointers and counters as between
GPUs, only the stores stay on the device) and, where the box has a second GPU, between two GPUs. Every instance's frame must equal the frame of a single
whole-frame pass bit for bit, frame after frame (the two frame buffers alternate)."""
import ctypes as C

import numpy as np
import pytest

from tests import harness as H
from vulkan_renderer_b200 import api

pytestmark = pytest.mark.gpu


def _device_count():
	import torch
	return torch.cuda.device_count()


def _run_exchanged(cuda_devices, width, height, frames=3, spp=2, dataset="mini_city"):
	"""Instance r lives on cuda_devices[r]. Returns (frames of every instance after the last exchange, the whole-frame render)."""
	import torch
	info = H.dataset(dataset)
	world = len(cuda_devices)
	rigs = []
	try:
		for r, dev in enumerate(cuda_devices):
			frame = H.open_frame(info, cuda_device=dev)   # stream = None: every instance gets a stream of its own from the library
			frame.configure(sample_count=spp, trace_shadow_rays=1)
			rigs.append(frame)
		lib = rigs[0].lib
		constants = rigs[0].constants(width, height)
		whole = None
		gbs, passes, exchanges = [], [], []
		for r, frame in enumerate(rigs):
			_, gb = frame.gbuffer_host(width, height)
			with torch.cuda.device(cuda_devices[r]):
				gbs.append(torch.from_numpy(gb).to(torch.device("cuda", cuda_devices[r])))
			if r == 0:
				whole = frame.shade_host(width, height, gb)
			passes.append(frame.create_pass(width, height, stripe_index=r, stripe_count=world))
			ex = api.FrameExchange()
			assert lib.vkr_create_frame_exchange(C.byref(ex), C.byref(frame.device), width, height, r, world) == 0
			ex.timeout_ns = 5 * 10 ** 9   # a barrier that never completes must fail this test quickly, not stall it
			exchanges.append(ex)
		blocks = (C.c_void_p * world)(*[ex.d_block for ex in exchanges])
		for r, frame in enumerate(rigs):
			assert lib.vkr_frame_exchange_connect_local(C.byref(exchanges[r]), C.byref(frame.device), blocks) == 0
		results = []
		for f in range(frames):
			for r, frame in enumerate(rigs):   # asynchronous: instance 0's wait kernel spins while instance 1's kernels run on another stream
				assert lib.vkr_shading_pass_run_exchange(C.byref(passes[r]), C.byref(frame.device), constants, len(constants), gbs[r].data_ptr(), C.byref(exchanges[r])) == 0
			results = []
			for r, frame in enumerate(rigs):
				host = np.full((height, width, 4), np.nan, dtype=np.float32)
				assert lib.vkr_frame_exchange_download(C.byref(exchanges[r]), C.byref(frame.device), host.ctypes.data) == 0
				results.append(host)
			for r in range(world):
				assert exchanges[r].frames_exchanged == f + 1
		for r, frame in enumerate(rigs):
			lib.vkr_destroy_frame_exchange(C.byref(exchanges[r]), C.byref(frame.device))
		return results, whole
	finally:
		for frame in rigs:
			frame.close()


@pytest.mark.parametrize("world,width,height", [(2, 96, 48), (3, 100, 75), (8, 64, 16)])   # even split; ragged last tile column and row; more instances than tile columns
def test_exchange_on_one_gpu_reproduces_the_whole_frame(world, width, height):
	results, whole = _run_exchanged([0] * world, width, height)
	for r, frame in enumerate(results):
		assert np.array_equal(frame.view(np.uint32), whole.view(np.uint32)), "instance %d of %d holds a different frame" % (r, world)


def test_exchange_between_two_gpus_reproduces_the_whole_frame():
	if _device_count() < 2:
		pytest.skip("needs two GPUs")
	results, whole = _run_exchanged([0, 1], 320, 192, frames=4, spp=4)
	for r, frame in enumerate(results):
		assert np.array_equal(frame.view(np.uint32), whole.view(np.uint32)), "GPU %d holds a different frame" % r


def test_host_exchange_entry_point_and_mismatch_errors():
	"""vkr_shading_pass_run_host_exchange with a single instance (world 1) == vkr_shading_pass_run_host; an exchange of another shape is refused."""
	info = H.dataset("mini_city")
	frame = H.open_frame(info)
	try:
		frame.configure(sample_count=2, trace_shadow_rays=1)
		lib = frame.lib
		w, h = 96, 48
		constants = frame.constants(w, h)
		_, gb = frame.gbuffer_host(w, h)
		whole = frame.shade_host(w, h, gb)
		p = frame.create_pass(w, h)
		ex = api.FrameExchange()
		assert lib.vkr_create_frame_exchange(C.byref(ex), C.byref(frame.device), w, h, 0, 1) == 0
		out = np.zeros((h, w, 4), dtype=np.float32)
		gbc = np.ascontiguousarray(gb, dtype=np.float32)
		assert lib.vkr_shading_pass_run_host_exchange(C.byref(p), C.byref(frame.device), constants, len(constants), gbc.ctypes.data, C.byref(ex), out.ctypes.data) == 0
		assert np.array_equal(out.view(np.uint32), whole.view(np.uint32))
		other = api.FrameExchange()
		assert lib.vkr_create_frame_exchange(C.byref(other), C.byref(frame.device), w, h + 8, 0, 1) == 0
		assert lib.vkr_shading_pass_run_host_exchange(C.byref(p), C.byref(frame.device), constants, len(constants), gbc.ctypes.data, C.byref(other), out.ctypes.data) != 0
		assert lib.vkr_create_frame_exchange(C.byref(api.FrameExchange()), C.byref(frame.device), w, h, 2, 2) != 0   # rank out of range
		lib.vkr_destroy_frame_exchange(C.byref(other), C.byref(frame.device)); lib.vkr_destroy_frame_exchange(C.byref(ex), C.byref(frame.device))
	finally:
		frame.close()


def test_tile_order_follows_the_measured_costs_and_never_changes_a_pixel():
	"""After a frame the pass reads back what every tile cost and launches the next frame dearest tile first (vkr_api.cu, reorder_tiles_by_cost)."""
	import torch
	info = H.dataset("mini_city")
	frame = H.open_frame(info)
	try:
		frame.configure(sample_count=4, trace_shadow_rays=1)
		lib = frame.lib
		w, h = 320, 192
		constants = frame.constants(w, h)
		_, gb = frame.gbuffer_host(w, h)
		dev = torch.device("cuda", 0)
		gbd = torch.from_numpy(gb).to(dev); out = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
		p = frame.create_pass(w, h)
		assert p.reorder_tiles == 1 and p.tile_count == (w // 16) * (h // 8)
		frames = []
		for f in range(4):
			out.zero_(); torch.cuda.synchronize()   # torch's stream is not the library's
			assert lib.vkr_shading_pass_run(C.byref(p), C.byref(frame.device), constants, len(constants), gbd.data_ptr(), out.data_ptr()) == 0
			assert lib.vkr_shading_pass_wait(C.byref(p), C.byref(frame.device)) == 0
			frames.append(out.cpu().numpy().copy())
		for f in frames[1:]:
			assert np.array_equal(f.view(np.uint32), frames[0].view(np.uint32))
		order = np.ctypeslib.as_array(C.cast(p.h_tile_list, C.POINTER(C.c_uint32)), (p.tile_count,)).copy()
		cost = np.ctypeslib.as_array(C.cast(p.h_tile_cost, C.POINTER(C.c_uint32)), (p.tile_count,)).copy()
		assert sorted(order.tolist()) == list(range(p.tile_count))        # a permutation of this instance's tiles
		assert cost.min() > 0                                              # every tile reported a cost
		assert order.tolist() != list(range(p.tile_count))                 # and the order has left row-major behind
	finally:
		frame.close()
